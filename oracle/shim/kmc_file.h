// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for the KMC API header
// <kmc_file.h> (third party, "KMC >= v2.3", README.md:23; not vendored in the
// reference and not installed here).  Call sites it serves: main.cpp:274-279,
// 444-449,482-490 and the `uint32` typedef used by bloom_filter.hpp:100,109 and
// kmap.hpp:119.
//
// It lists the (k-mer, count) records of a KMC database <prefix>.kmc_pre /
// <prefix>.kmc_suf.  The on-disk layout is restated from the published KMC
// format description (KMC1 "version 0" and KMC2 "0x200" prefix files):
//   .kmc_pre = "KMCP" | u64 LUT[...] (+ guard) | [0x200: u32 signature map] |
//              64-byte header | u32 header_offset | "KMCP"
//   .kmc_suf = "KMCS" | total_kmers x ((k-p)/4 suffix bytes + counter bytes) | "KMCS"
// No test of the reference pins this format: parity at the KMC-file boundary
// is UNPINNED; what the haploid golden pins is the semantics
// (canonical k-mers, count >= 2, counter cap 255).
#pragma once
#include <sys/types.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef unsigned int uint32;
typedef unsigned long long uint64;
typedef long long int64;
typedef unsigned char uchar;

class CKmerAPI {
 public:
  explicit CKmerAPI(uint32 klen = 0) : klen_(klen), s_(klen, 'A') {}
  void to_string(char *out) const {
    memcpy(out, s_.data(), klen_);
    out[klen_] = '\0';
  }
  std::string to_string() const { return s_; }
  uint32 klen_;
  std::string s_;
};

class CKMCFile {
 public:
  CKMCFile() {}
  ~CKMCFile() { Close(); }

  bool OpenForListing(const std::string &prefix) {
    Close();
    FILE *fp = fopen((prefix + ".kmc_pre").c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> pre((size_t)fsz);
    if (fsz < 84 || fread(pre.data(), 1, (size_t)fsz, fp) != (size_t)fsz) {
      fclose(fp);
      return false;
    }
    fclose(fp);
    if (memcmp(pre.data(), "KMCP", 4) != 0 || memcmp(pre.data() + fsz - 4, "KMCP", 4) != 0) return false;
    uint32 version = rd32(&pre[(size_t)fsz - 12]);
    uint32 hoff = rd32(&pre[(size_t)fsz - 8]);
    if (version != 0 && version != 0x200) return false;
    const unsigned char *h = &pre[(size_t)fsz - 8 - hoff];
    size_t o = 0;
    klen_ = rd32(h + o); o += 4;
    mode_ = rd32(h + o); o += 4;
    counter_size_ = rd32(h + o); o += 4;
    lut_prefix_len_ = rd32(h + o); o += 4;
    signature_len_ = 0;
    if (version == 0x200) { signature_len_ = rd32(h + o); o += 4; }
    min_count_ = rd32(h + o); o += 4;
    max_count_ = rd32(h + o); o += 4;
    total_kmers_ = rd64(h + o); o += 8;
    both_strands_ = !(h[o] & 1);
    if (o + 5 + 4 <= hoff) max_count_ |= (uint64)rd32(h + o + 1) << 32;  // KMC 3: high word of max_count
    size_t sigmap_bytes = version == 0x200 ? (((size_t)1 << (2 * signature_len_)) + 1) * 4 : 0;
    size_t lut_bytes = (size_t)fsz - 4 - 8 - hoff - sigmap_bytes;
    size_t n = lut_bytes / 8;
    lut_.resize(n);
    for (size_t i = 0; i < n; ++i) lut_[i] = rd64(&pre[4 + 8 * i]);
    single_lut_ = (size_t)1 << (2 * lut_prefix_len_);
    size_t n_bins = n / single_lut_;
    n_lut_ = n_bins * single_lut_;  // anything after that is a guard entry
    if (n_lut_ == 0) return false;
    suf_bytes_ = (klen_ - lut_prefix_len_) / 4;
    suf_ = fopen((prefix + ".kmc_suf").c_str(), "rb");
    if (!suf_) return false;
    char m[4];
    if (fread(m, 1, 4, suf_) != 4 || memcmp(m, "KMCS", 4) != 0) return false;
    rec_ = 0;
    prefix_index_ = 0;
    return true;
  }

  bool Info(uint32 &kmer_length, uint32 &mode, uint32 &counter_size, uint32 &lut_prefix_length,
            uint32 &signature_len, uint32 &min_count, uint64 &max_count, uint64 &total_kmers) {
    if (!suf_) return false;
    kmer_length = klen_;
    mode = mode_;
    counter_size = counter_size_;
    lut_prefix_length = lut_prefix_len_;
    signature_len = signature_len_;
    min_count = min_count_;
    max_count = max_count_;
    total_kmers = total_kmers_;
    return true;
  }

  bool ReadNextKmer(CKmerAPI &kmer, uint32 &count) {
    static const char SYM[4] = {'A', 'C', 'G', 'T'};
    if (!suf_) return false;
    unsigned char buf[80];
    while (rec_ < total_kmers_) {
      while (prefix_index_ + 1 < n_lut_ && lut_[prefix_index_ + 1] <= rec_) ++prefix_index_;
      if (fread(buf, 1, suf_bytes_ + counter_size_, suf_) != suf_bytes_ + counter_size_) return false;
      ++rec_;
      uint64 c = 0;
      for (uint32 b = 0; b < counter_size_; ++b) c |= (uint64)buf[suf_bytes_ + b] << (8 * b);
      if (counter_size_ == 0) c = 1;
      if (c < min_count_ || c > max_count_) continue;
      uint64 pfx = prefix_index_ % single_lut_;
      kmer.klen_ = klen_;
      kmer.s_.resize(klen_);
      for (uint32 i = 0; i < lut_prefix_len_; ++i)
        kmer.s_[i] = SYM[(pfx >> (2 * (lut_prefix_len_ - 1 - i))) & 3];
      for (uint32 i = 0; i < suf_bytes_ * 4; ++i)
        kmer.s_[lut_prefix_len_ + i] = SYM[(buf[i >> 2] >> (2 * (3 - (i & 3)))) & 3];
      count = (uint32)c;
      return true;
    }
    return false;
  }

  void Close() {
    if (suf_) fclose(suf_);
    suf_ = nullptr;
  }

 private:
  static uint32 rd32(const unsigned char *p) {
    return (uint32)p[0] | ((uint32)p[1] << 8) | ((uint32)p[2] << 16) | ((uint32)p[3] << 24);
  }
  static uint64 rd64(const unsigned char *p) { return (uint64)rd32(p) | ((uint64)rd32(p + 4) << 32); }

  FILE *suf_ = nullptr;
  uint32 klen_ = 0, mode_ = 0, counter_size_ = 0, lut_prefix_len_ = 0, signature_len_ = 0;
  uint32 min_count_ = 0;
  uint64 max_count_ = 0;
  uint64 total_kmers_ = 0, rec_ = 0;
  bool both_strands_ = true;
  std::vector<uint64> lut_;
  size_t single_lut_ = 0, n_lut_ = 0, prefix_index_ = 0, suf_bytes_ = 0;
};
