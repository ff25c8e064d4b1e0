// The index file of `malva-geno index` (same name as the reference's: <vcf>.c<ref_k>.k<k>.malvax.zst,
// main.cpp:406-412, read back at :455-461).  Both ends of the file are ours, so the content is our own layout
// (SURVEY 8f-3): the reference serialises the two raw bit vectors (2 x bf_bits/8 bytes: 8 GiB for -b 4) plus
// every ref key as text; here the filters travel as delta-coded sorted lists of set-bit indices and ref_bf as
// packed canonical keys -- what the device exports and imports (mg_export_set_bits / mg_import_set_bits /
// mg_export_ref_keys / mg_add_signatures_packed) -- in zstd-compressed chunks.
//
//   "MALVAGPUIDX1\0\0\0\0" | u32 k | u32 ref_k | u64 bf_bits
//   3 sections (context_bf bits, bf bits, ref_bf keys): u64 n_items, then chunks {u64 raw_bytes, u64 zstd_bytes,
//   data}, terminated by a chunk with raw_bytes == 0
//
// libzstd.so.1 is part of the image; its stable one-shot API is declared here (no zstd.h installed).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

extern "C" {
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
}

namespace mh {

constexpr char INDEX_MAGIC[16] = {'M', 'A', 'L', 'V', 'A', 'G', 'P', 'U', 'I', 'D', 'X', '1', 0, 0, 0, 0};
constexpr size_t INDEX_CHUNK = 8u << 20;  // chunks are compressed and decompressed independently, several at a time

inline int index_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return (int)(n < 1 ? 1 : n > 16 ? 16 : n);
}
// fn(i) for i in [0, n) on up to index_threads() threads
template <typename F>
inline void index_parallel(size_t n, F &&fn) {
  const int t = (int)std::min<size_t>((size_t)index_threads(), n);
  if (t <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int w = 0; w < t; ++w)
    pool.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
    });
  for (auto &th : pool) th.join();
}

class IndexWriter {
 public:
  IndexWriter(const std::string &path, uint32_t k, uint32_t ref_k, uint64_t bf_bits) : fp_(fopen(path.c_str(), "wb")) {
    if (!fp_) throw std::runtime_error("cannot write " + path);
    put(INDEX_MAGIC, 16);
    put(&k, 4);
    put(&ref_k, 4);
    put(&bf_bits, 8);
  }
  ~IndexWriter() {
    if (fp_) fclose(fp_);
  }
  // sorted set-bit indices -> deltas (small numbers compress to ~1-2 bytes each)
  void write_bits(std::vector<uint64_t> &idx) {
    uint64_t prev = 0;
    for (auto &x : idx) {
      uint64_t d = x - prev;
      prev = x;
      x = d;
    }
    write_section(idx.data(), idx.size(), 8);
  }
  void write_keys(const std::vector<uint64_t> &lohi) { write_section(lohi.data(), lohi.size() / 2, 16); }
  void close() {
    if (fp_ && fclose(fp_) != 0) {
      fp_ = nullptr;
      throw std::runtime_error("error closing the index file");
    }
    fp_ = nullptr;
  }

 private:
  void put(const void *p, size_t n) {
    if (fwrite(p, 1, n, fp_) != n) throw std::runtime_error("short write on the index file");
  }
  void write_section(const void *data, uint64_t n_items, size_t item) {
    put(&n_items, 8);
    const uint8_t *p = (const uint8_t *)data;
    const uint64_t total = n_items * item, n_chunks = (total + INDEX_CHUNK - 1) / INDEX_CHUNK;
    // a wave of chunks is compressed in parallel (level 1: the file is read back once, by `call`), then written in order
    const size_t wave = (size_t)index_threads() * 2, bound = ZSTD_compressBound(INDEX_CHUNK);
    std::vector<std::vector<uint8_t>> out(std::min<uint64_t>(wave, n_chunks));
    std::vector<size_t> zs(out.size());
    for (uint64_t c0 = 0; c0 < n_chunks; c0 += wave) {
      const size_t m = (size_t)std::min<uint64_t>(wave, n_chunks - c0);
      bool failed = false;
      index_parallel(m, [&](size_t j) {
        const uint64_t off = (c0 + j) * INDEX_CHUNK, raw = std::min<uint64_t>(INDEX_CHUNK, total - off);
        if (out[j].size() < bound) out[j].resize(bound);
        zs[j] = ZSTD_compress(out[j].data(), out[j].size(), p + off, (size_t)raw, 1);
        if (ZSTD_isError(zs[j])) failed = true;
      });
      if (failed) throw std::runtime_error("zstd compression failed");
      for (size_t j = 0; j < m; ++j) {
        const uint64_t off = (c0 + j) * INDEX_CHUNK, raw = std::min<uint64_t>(INDEX_CHUNK, total - off), z64 = zs[j];
        put(&raw, 8);
        put(&z64, 8);
        put(out[j].data(), zs[j]);
      }
    }
    uint64_t zero = 0;
    put(&zero, 8);
  }
  FILE *fp_;
};

class IndexReader {
 public:
  uint32_t k = 0, ref_k = 0;
  uint64_t bf_bits = 0;
  explicit IndexReader(const std::string &path) : fp_(fopen(path.c_str(), "rb")) {
    if (!fp_) throw std::runtime_error("cannot open index " + path + " (run `malva-geno index` first)");
    char magic[16];
    get(magic, 16);
    if (memcmp(magic, INDEX_MAGIC, 16) != 0)
      throw std::runtime_error(path + " is not an index written by this malva-geno (indexes of the CPU build hold raw "
                                      "sdsl bit vectors and are not read here): re-run `malva-geno index`");
    get(&k, 4);
    get(&ref_k, 4);
    get(&bf_bits, 8);
  }
  ~IndexReader() {
    if (fp_) fclose(fp_);
  }
  std::vector<uint64_t> read_bits() {
    std::vector<uint64_t> idx = read_section(8);
    uint64_t acc = 0;
    for (auto &x : idx) {
      acc += x;
      x = acc;
    }
    return idx;
  }
  std::vector<uint64_t> read_keys() { return read_section(16); }

 private:
  void get(void *p, size_t n) {
    if (fread(p, 1, n, fp_) != n) throw std::runtime_error("index file truncated");
  }
  std::vector<uint64_t> read_section(size_t item) {
    uint64_t n_items = 0;
    get(&n_items, 8);
    std::vector<uint64_t> data(n_items * item / 8);
    uint8_t *p = (uint8_t *)data.data();
    const uint64_t total = n_items * item;
    // the compressed chunks are read in file order, then decompressed in parallel into their places
    struct Chunk {
      uint64_t raw, off;
      std::vector<uint8_t> z;
    };
    std::vector<Chunk> chunks;
    uint64_t off = 0;
    while (true) {
      uint64_t raw = 0, z = 0;
      get(&raw, 8);
      if (raw == 0) break;
      get(&z, 8);
      if (raw > total - off) throw std::runtime_error("index file corrupt (section overflow)");
      chunks.push_back(Chunk{raw, off, std::vector<uint8_t>(z)});
      get(chunks.back().z.data(), z);
      off += raw;
    }
    if (off != total) throw std::runtime_error("index file corrupt (section short)");
    bool failed = false;
    index_parallel(chunks.size(), [&](size_t i) {
      const Chunk &c = chunks[i];
      size_t r = ZSTD_decompress(p + c.off, (size_t)c.raw, c.z.data(), c.z.size());
      if (ZSTD_isError(r) || r != c.raw) failed = true;
    });
    if (failed) throw std::runtime_error("index file corrupt (zstd)");
    return data;
  }
  FILE *fp_;
};

}  // namespace mh
