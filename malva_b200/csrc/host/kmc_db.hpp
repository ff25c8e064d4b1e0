// KMC database reader for the call side (replaces CKMCFile::OpenForListing / Info / ReadNextKmer + CKmerAPI::to_string,
// main.cpp:444-449, 482-490).  The KMC API is third party ("KMC >= v2.3", README.md:23) and not vendored by the
// reference; the on-disk layout is restated from the published format description (KMC1 "version 0" and KMC2
// "0x200" prefix files):
//   <db>.kmc_pre = "KMCP" | u64 LUT[...] (+ guard) | [0x200: u32 signature map] | header | u32 header_offset | "KMCP"
//   <db>.kmc_suf = "KMCS" | total_kmers x ((k - p)/4 suffix bytes + counter bytes) | "KMCS"
// Nothing is decoded on the host: the prefix LUT goes to the device once (mg_kmc_open) and the suffix records
// are streamed as they lie in the file (mg_scan_kmc_records), 10 bytes per 43-mer.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mh {

struct KmcDb {
  uint32_t kmer_len = 0, mode = 0, counter_size = 0, lut_prefix_len = 0, signature_len = 0, min_count = 0;
  uint64_t max_count = 0, total_kmers = 0;
  bool both_strands = true;
  std::vector<uint64_t> lut;  // n_lut entries (a multiple of 4^lut_prefix_len): records before each prefix
  uint32_t record_bytes = 0;
  FILE *suf = nullptr;

  ~KmcDb() {
    if (suf) fclose(suf);
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
    both_strands = !(h[o] & 1);
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
    suf = fopen((prefix + ".kmc_suf").c_str(), "rb");
    char m[4];
    if (!suf || fread(m, 1, 4, suf) != 4 || memcmp(m, "KMCS", 4) != 0) {
      why = "cannot open " + prefix + ".kmc_suf";
      return false;
    }
    return true;
  }

  // next `max_records` whole records into dst; returns the number read (0 at the end)
  uint64_t read_records(uint8_t *dst, uint64_t first_record, uint64_t max_records) {
    if (first_record >= total_kmers) return 0;
    uint64_t want = total_kmers - first_record < max_records ? total_kmers - first_record : max_records;
    size_t got = fread(dst, record_bytes, (size_t)want, suf);
    return (uint64_t)got;
  }
};

}  // namespace mh
