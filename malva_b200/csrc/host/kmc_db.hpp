// KMC database reader for the call side (replaces CKMCFile::OpenForListing / Info / ReadNextKmer + CKmerAPI::to_string,
// main.cpp:444-449, 482-490).  The KMC API is third party ("KMC >= v2.3", README.md:23) and not vendored by the
// reference; the on-disk layout is restated from the published format description (KMC1 "version 0" and KMC2
// "0x200" prefix files):
//   <db>.kmc_pre = "KMCP" | u64 LUT[...] (+ guard) | [0x200: u32 signature map] | header | u32 header_offset | "KMCP"
//   <db>.kmc_suf = "KMCS" | total_kmers x ((k - p)/4 suffix bytes + counter bytes) | "KMCS"
// Nothing is decoded on the host: the prefix LUT goes to the device once (mg_kmc_open) and the suffix records
// are streamed as they lie in the file (mg_scan_kmc_records), 10 bytes per 43-mer.
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace mh {

struct KmcDb {
  uint32_t kmer_len = 0, mode = 0, counter_size = 0, lut_prefix_len = 0, signature_len = 0, min_count = 0;
  uint64_t max_count = 0, total_kmers = 0;
  bool both_strands = true;
  std::vector<uint64_t> lut;  // n_lut entries (a multiple of 4^lut_prefix_len): records before each prefix
  uint32_t record_bytes = 0;
  int suf_fd = -1;  // <db>.kmc_suf; records start at byte 4

  ~KmcDb() {
    if (suf_fd >= 0) close(suf_fd);
  }
  KmcDb() = default;
  KmcDb(const KmcDb &) = delete;
  KmcDb &operator=(const KmcDb &) = delete;

  static uint32_t rd32(const unsigned char *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }
  static uint64_t rd64(const unsigned char *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

  // false if the database cannot be opened (the reference prints "ERROR: cannot open" and returns 1)
  bool open(const std::string &prefix, std::string &why) {
    FILE *fp = fopen((prefix + ".kmc_pre").c_str(), "rb");
    if (!fp) {
      why = "cannot open " + prefix + ".kmc_pre";
      return false;
    }
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> pre((size_t)(fsz > 0 ? fsz : 0));
    bool ok = fsz >= 52 && fread(pre.data(), 1, (size_t)fsz, fp) == (size_t)fsz;
    fclose(fp);
    if (!ok || memcmp(pre.data(), "KMCP", 4) != 0 || memcmp(pre.data() + fsz - 4, "KMCP", 4) != 0) {
      why = prefix + ".kmc_pre: not a KMC prefix file";
      return false;
    }
    const uint32_t version = rd32(&pre[(size_t)fsz - 12]), hoff = rd32(&pre[(size_t)fsz - 8]);
    if ((version != 0 && version != 0x200) || (size_t)hoff + 12 > (size_t)fsz) {
      why = prefix + ".kmc_pre: unsupported KMC version";
      return false;
    }
    const unsigned char *h = &pre[(size_t)fsz - 8 - hoff];
    size_t o = 0;
    kmer_len = rd32(h + o), o += 4;
    mode = rd32(h + o), o += 4;
    counter_size = rd32(h + o), o += 4;
    lut_prefix_len = rd32(h + o), o += 4;
    signature_len = 0;
    if (version == 0x200) signature_len = rd32(h + o), o += 4;
    min_count = rd32(h + o), o += 4;
    max_count = rd32(h + o), o += 4;
    total_kmers = rd64(h + o), o += 8;
    both_strands = !(h[o] & 1);  // (stored inverted)
    if (o + 5 + 4 <= hoff) max_count |= (uint64_t)rd32(h + o + 1) << 32;  // KMC 3: high word of max_count
    if (lut_prefix_len > 15 || lut_prefix_len >= kmer_len || (kmer_len - lut_prefix_len) % 4 != 0 || counter_size > 8) {
      why = prefix + ".kmc_pre: unsupported layout";
      return false;
    }
    const size_t sigmap = version == 0x200 ? (((size_t)1 << (2 * signature_len)) + 1) * 4 : 0;
    if ((size_t)fsz < 4 + 8 + (size_t)hoff + sigmap) {
      why = prefix + ".kmc_pre: truncated";
      return false;
    }
    const size_t n = ((size_t)fsz - 4 - 8 - hoff - sigmap) / 8, single = (size_t)1 << (2 * lut_prefix_len);
    const size_t n_lut = (n / single) * single;  // anything after that is a guard entry
    if (n_lut == 0) {
      why = prefix + ".kmc_pre: empty prefix table";
      return false;
    }
    lut.resize(n_lut);
    for (size_t i = 0; i < n_lut; ++i) lut[i] = rd64(&pre[4 + 8 * i]);
    record_bytes = (kmer_len - lut_prefix_len) / 4 + counter_size;
    suf_fd = ::open((prefix + ".kmc_suf").c_str(), O_RDONLY);
    char m[4];
    if (suf_fd < 0 || pread(suf_fd, m, 4, 0) != 4 || memcmp(m, "KMCS", 4) != 0) {
      why = "cannot open " + prefix + ".kmc_suf";
      return false;
    }
    return true;
  }

  // up to `max_records` whole records starting at record `first_record` into dst; returns the number read (0 at the
  // end).  Large requests are split over a few threads (pread): copying out of the page cache is the bottleneck of
  // the whole scan once the device side runs at PCIe speed.
  uint64_t read_records(uint8_t *dst, uint64_t first_record, uint64_t max_records, int threads = 4) {
    if (first_record >= total_kmers) return 0;
    const uint64_t want = total_kmers - first_record < max_records ? total_kmers - first_record : max_records;
    const uint64_t bytes = want * record_bytes, off0 = 4 + first_record * record_bytes;
    auto read_all = [&](uint64_t b, uint64_t e) -> uint64_t {  // bytes [b, e) of the request
      uint64_t done = b;
      while (done < e) {
        ssize_t r = pread(suf_fd, dst + done, (size_t)(e - done), (off_t)(off0 + done));
        if (r <= 0) break;
        done += (uint64_t)r;
      }
      return done - b;
    };
    if (threads < 2 || bytes < (8u << 20)) return read_all(0, bytes) / record_bytes;
    std::vector<uint64_t> got((size_t)threads, 0);
    std::vector<std::thread> pool;
    const uint64_t piece = (bytes + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t)
      pool.emplace_back([&, t] {
        uint64_t b = (uint64_t)t * piece, e = b + piece < bytes ? b + piece : bytes;
        if (b < e) got[(size_t)t] = read_all(b, e);
      });
    for (auto &th : pool) th.join();
    uint64_t total = 0;  // contiguous prefix actually read
    for (int t = 0; t < threads; ++t) {
      uint64_t b = (uint64_t)t * piece, e = b + piece < bytes ? b + piece : bytes;
      total += got[(size_t)t];
      if (b < e && got[(size_t)t] < e - b) break;
    }
    return total / record_bytes;
  }
};

// Writer of the same two files (KMC2 "0x200" layout, counter_size 1): what `malva-geno count` produces in place of
// `kmc`.  Records must arrive in ascending k-mer order (append() may be called several times).
class KmcWriter {
 public:
  // the prefix length the Python twin (malva_b200/kmc.py) picks: the largest p <= 13 with (k - p) % 4 == 0 and
  // 4^p <= max(64, n); the smallest valid p if there is none
  static uint32_t choose_prefix_len(uint32_t k, uint64_t n) {
    int best = -1;
    for (uint32_t p = 1; p <= (k < 13 ? k : 13); ++p) {
      if ((k - p) % 4) continue;
      if (best < 0 || (1ull << (2 * p)) <= (n > 64 ? n : 64)) best = (int)p;
    }
    if (best < 0) throw std::runtime_error("no LUT prefix length with (k-p)%4==0 for this k");
    return (uint32_t)best;
  }

  KmcWriter(const std::string &prefix, uint32_t k, uint32_t lut_prefix_len, uint32_t min_count, uint32_t counter_max)
      : prefix_(prefix), k_(k), p_(lut_prefix_len), min_count_(min_count), counter_max_(counter_max) {
    if ((k_ - p_) % 4 || p_ >= k_) throw std::runtime_error("(k - lut_prefix_len) must be a positive multiple of 4");
    suf_ = fopen((prefix + ".kmc_suf").c_str(), "wb");
    if (!suf_) throw std::runtime_error("cannot write " + prefix + ".kmc_suf");
    fwrite("KMCS", 1, 4, suf_);
    bins_.assign((((size_t)1) << (2 * p_)) + 1, 0);
  }
  ~KmcWriter() {
    if (suf_) fclose(suf_);
  }
  // n records {lo, hi} (packed canonical k-mers, right-aligned) + counts
  void append(const uint64_t *lohi, const uint32_t *counts, uint64_t n) {
    const uint32_t suf_syms = k_ - p_, sb = suf_syms / 4;
    std::vector<uint8_t> buf;
    buf.reserve((size_t)((n < (1u << 20) ? n : (1u << 20)) * (sb + 1)));
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t lo = lohi[2 * i], hi = lohi[2 * i + 1];
      // prefix = the top p symbols
      const uint32_t sh = 2 * suf_syms;
      const uint64_t pre = sh >= 64 ? (hi >> (sh - 64)) : ((lo >> sh) | (sh ? hi << (64 - sh) : 0));
      bins_[(size_t)pre + 1]++;
      for (uint32_t j = 0; j < sb; ++j) {  // suffix bytes, most significant first
        const uint32_t bit = 8 * (sb - 1 - j);
        uint64_t byte = bit >= 64 ? (hi >> (bit - 64)) : ((lo >> bit) | (bit && bit > 56 ? hi << (64 - bit) : 0));
        buf.push_back((uint8_t)(byte & 0xFF));
      }
      buf.push_back((uint8_t)(counts[i] > 255 ? 255 : counts[i]));
      if (buf.size() >= (1u << 24)) flush(buf);
    }
    flush(buf);
    total_ += n;
  }
  void close() {
    fwrite("KMCS", 1, 4, suf_);
    if (fclose(suf_) != 0) {
      suf_ = nullptr;
      throw std::runtime_error("error closing " + prefix_ + ".kmc_suf");
    }
    suf_ = nullptr;
    for (size_t i = 1; i < bins_.size(); ++i) bins_[i] += bins_[i - 1];  // bins_[i] = records with prefix < i
    FILE *fp = fopen((prefix_ + ".kmc_pre").c_str(), "wb");
    if (!fp) throw std::runtime_error("cannot write " + prefix_ + ".kmc_pre");
    fwrite("KMCP", 1, 4, fp);
    fwrite(bins_.data(), 8, bins_.size(), fp);
    const uint32_t sig_len = 5;
    std::vector<uint32_t> sigmap((((size_t)1) << (2 * sig_len)) + 1, 0);
    fwrite(sigmap.data(), 4, sigmap.size(), fp);
    uint8_t hdr[64] = {0};
    auto put32 = [&](size_t o, uint32_t v) { memcpy(hdr + o, &v, 4); };
    put32(0, k_), put32(4, 0), put32(8, 1), put32(12, p_), put32(16, sig_len), put32(20, min_count_), put32(24, counter_max_);
    memcpy(hdr + 28, &total_, 8);
    hdr[36] = 0;           // both strands (stored inverted)
    put32(60, 0x200);      // KMC2 layout
    fwrite(hdr, 1, 64, fp);
    const uint32_t hoff = 64;
    fwrite(&hoff, 4, 1, fp);
    fwrite("KMCP", 1, 4, fp);
    if (fclose(fp) != 0) throw std::runtime_error("error closing " + prefix_ + ".kmc_pre");
  }
  uint64_t total() const { return total_; }

 private:
  void flush(std::vector<uint8_t> &buf) {
    if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), suf_) != buf.size())
      throw std::runtime_error("short write on " + prefix_ + ".kmc_suf");
    buf.clear();
  }
  std::string prefix_;
  uint32_t k_, p_, min_count_, counter_max_;
  FILE *suf_ = nullptr;
  std::vector<uint64_t> bins_;
  uint64_t total_ = 0;
};

}  // namespace mh
