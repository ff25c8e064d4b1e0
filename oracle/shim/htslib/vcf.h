// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for the slice of htslib's
// <htslib/vcf.h> that the reference's main.cpp / variant.hpp call
// (main.cpp:190-219,261-272,309-312,374-376,505-515,522-524; variant.hpp:66-211).
// htslib is not installed in this image.  This is a text/gz VCF reader that
// reproduces the htslib conventions the reference's results depend on:
//   * header: hrec list with de-duplication by key(+ID), FILTER=PASS forced to
//     exist (inserted right after ##fileformat), sample subsetting;
//   * record: rid/pos/qual/n_allele/d.id/d.allele, missing QUAL = NaN pattern;
//   * INFO floats parsed with strtod then narrowed to float;
//   * GT encoded as ((allele+1)<<1)|phased, vector_end padding.
// Only plumbing is restated here; all MALVA arithmetic is the reference's own.
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#ifndef KSTRING_T
#define KSTRING_T kstring_t
typedef struct __kstring_t {
  size_t l, m;
  char *s;
} kstring_t;
#endif

#define BCF_UN_STR 1
#define BCF_UN_ALL 15
#define bcf_int32_vector_end (INT32_MIN + 1)
#define bcf_int32_missing INT32_MIN
#define bcf_gt_missing 0
#define bcf_gt_is_phased(idx) ((idx)&1)
#define bcf_gt_allele(val) (((val) >> 1) - 1)

struct htsFile {
  gzFile fp = nullptr;
  std::string pending;  // first non-header line, read ahead by bcf_hdr_read
  bool has_pending = false;
};

struct shim_hrec {
  std::string key;   // e.g. "INFO", "contig", "fileformat"
  std::string id;    // ID=... for structured lines, "" otherwise
  std::string line;  // full text without trailing newline
};

struct bcf_hdr_t {
  std::vector<shim_hrec> hrecs;
  std::vector<std::string> all_samples;  // as in the file (plus added ones)
  std::vector<int> keep;                 // indices into all_samples, header order
  bool has_format_col = false;
  std::vector<std::string> contigs;
  std::map<std::string, int> contig_id;
};

struct bcf_dec_t {
  char *id = nullptr;
  char **allele = nullptr;
};

struct bcf1_t {
  int32_t rid = 0;
  int64_t pos = 0;
  float qual = 0;
  uint32_t n_allele = 0;
  bcf_dec_t d;
  // backing storage
  std::string id_s;
  std::vector<std::string> allele_s;
  std::vector<char *> allele_p;
  std::string info_s, format_s;
  std::vector<std::string> sample_s;  // all sample columns of the line
};

static inline bool shim_gets(gzFile fp, std::string &out) {
  out.clear();
  char buf[1 << 16];
  bool any = false;
  while (gzgets(fp, buf, (int)sizeof(buf)) != nullptr) {
    any = true;
    size_t n = strlen(buf);
    if (n && buf[n - 1] == '\n') {
      out.append(buf, n - 1);
      if (!out.empty() && out.back() == '\r') out.pop_back();
      return true;
    }
    out.append(buf, n);
  }
  return any;
}

static inline bool shim_parse_hrec(const std::string &line, shim_hrec &h) {
  if (line.size() < 3 || line[0] != '#' || line[1] != '#') return false;
  size_t eq = line.find('=');
  if (eq == std::string::npos) return false;
  h.key = line.substr(2, eq - 2);
  h.line = line;
  h.id.clear();
  if (eq + 1 < line.size() && line[eq + 1] == '<') {
    size_t p = line.find("ID=", eq);
    if (p != std::string::npos) {
      size_t e = line.find_first_of(",>", p);
      h.id = line.substr(p + 3, e == std::string::npos ? std::string::npos : e - p - 3);
    }
  }
  return true;
}

static inline int shim_hdr_add(bcf_hdr_t *h, const shim_hrec &r) {
  for (const auto &o : h->hrecs) {
    if (o.key != r.key) continue;
    if (!r.id.empty() || !o.id.empty()) {
      if (o.id == r.id) return 0;  // structured duplicate (same key, same ID)
    } else if (o.line == r.line || r.key == "fileformat") {
      return 0;
    }
  }
  if (r.key == "fileformat") {
    h->hrecs.insert(h->hrecs.begin(), r);
  } else {
    h->hrecs.push_back(r);
  }
  if (r.key == "contig" && !r.id.empty() && !h->contig_id.count(r.id)) {
    h->contig_id[r.id] = (int)h->contigs.size();
    h->contigs.push_back(r.id);
  }
  return 0;
}

static inline htsFile *bcf_open(const char *path, const char * /*mode*/) {
  gzFile fp = gzopen(path, "r");
  if (!fp) return nullptr;
  htsFile *f = new htsFile();
  f->fp = fp;
  return f;
}

static inline int bcf_close(htsFile *f) {
  if (!f) return -1;
  gzclose(f->fp);
  delete f;
  return 0;
}

static inline bcf_hdr_t *bcf_hdr_read(htsFile *f) {
  bcf_hdr_t *h = new bcf_hdr_t();
  shim_hrec pass;
  shim_parse_hrec("##FILTER=<ID=PASS,Description=\"All filters passed\">", pass);
  h->hrecs.push_back(pass);
  std::string line;
  while (shim_gets(f->fp, line)) {
    if (line.size() >= 2 && line[0] == '#' && line[1] == '#') {
      shim_hrec r;
      if (shim_parse_hrec(line, r)) shim_hdr_add(h, r);
      continue;
    }
    if (!line.empty() && line[0] == '#') {
      // #CHROM POS ID REF ALT QUAL FILTER INFO [FORMAT sample...]
      std::vector<std::string> cols;
      size_t s = 0;
      while (true) {
        size_t t = line.find('\t', s);
        cols.push_back(line.substr(s, t == std::string::npos ? std::string::npos : t - s));
        if (t == std::string::npos) break;
        s = t + 1;
      }
      while (!cols.empty() && cols.back().empty()) cols.pop_back();  // trailing tab
      h->has_format_col = cols.size() > 8;
      for (size_t i = 9; i < cols.size(); ++i) h->all_samples.push_back(cols[i]);
      for (size_t i = 0; i < h->all_samples.size(); ++i) h->keep.push_back((int)i);
      return h;
    }
    f->pending = line;
    f->has_pending = true;
    return h;
  }
  return h;
}

static inline void bcf_hdr_destroy(bcf_hdr_t *h) { delete h; }

// samples: NULL = none, "-" = all, comma list (is_file=0) or file of names.
// Returns 0, or i+1 for the first listed sample that is absent (htslib rule).
static inline int bcf_hdr_set_samples(bcf_hdr_t *h, const char *samples, int is_file) {
  h->keep.clear();
  if (samples == nullptr) return 0;
  if (strcmp(samples, "-") == 0) {
    for (size_t i = 0; i < h->all_samples.size(); ++i) h->keep.push_back((int)i);
    return 0;
  }
  std::vector<std::string> names;
  if (is_file) {
    gzFile fp = gzopen(samples, "r");
    if (!fp) return -1;
    std::string l;
    while (shim_gets(fp, l)) {
      size_t e = l.find_first_of(" \t");
      if (e != std::string::npos) l = l.substr(0, e);
      if (!l.empty()) names.push_back(l);
    }
    gzclose(fp);
  } else {
    std::string s(samples);
    size_t p = 0;
    while (true) {
      size_t t = s.find(',', p);
      names.push_back(s.substr(p, t == std::string::npos ? std::string::npos : t - p));
      if (t == std::string::npos) break;
      p = t + 1;
    }
  }
  std::vector<char> want(h->all_samples.size(), 0);
  int ret = 0;
  for (size_t i = 0; i < names.size(); ++i) {
    bool found = false;
    for (size_t j = 0; j < h->all_samples.size(); ++j)
      if (h->all_samples[j] == names[i]) {
        want[j] = 1;
        found = true;
      }
    if (!found && ret == 0) ret = (int)i + 1;
  }
  for (size_t j = 0; j < want.size(); ++j)
    if (want[j]) h->keep.push_back((int)j);
  return ret;
}

static inline int bcf_hdr_nsamples(const bcf_hdr_t *h) { return (int)h->keep.size(); }

static inline int bcf_hdr_append(bcf_hdr_t *h, const char *line) {
  shim_hrec r;
  std::string l(line);
  while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
  if (!shim_parse_hrec(l, r)) return -1;
  return shim_hdr_add(h, r);
}

static inline int bcf_hdr_add_sample(bcf_hdr_t *h, const char *s) {
  if (!s) return 0;
  h->all_samples.push_back(s);
  h->keep.push_back((int)h->all_samples.size() - 1);
  h->has_format_col = true;
  return 0;
}

static inline int bcf_hdr_sync(bcf_hdr_t *) { return 0; }

static inline int bcf_hdr_format(const bcf_hdr_t *h, int /*is_bcf*/, kstring_t *str) {
  std::string out;
  for (const auto &r : h->hrecs) {
    out += r.line;
    out += '\n';
  }
  out += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO";
  if (!h->keep.empty()) {
    out += "\tFORMAT";
    for (int i : h->keep) {
      out += '\t';
      out += h->all_samples[(size_t)i];
    }
  }
  out += '\n';
  str->s = (char *)malloc(out.size() + 1);
  memcpy(str->s, out.c_str(), out.size() + 1);
  str->l = out.size();
  str->m = out.size() + 1;
  return 0;
}

static inline const char *bcf_hdr_id2name(const bcf_hdr_t *h, int rid) {
  if (rid < 0 || (size_t)rid >= h->contigs.size()) return nullptr;
  return h->contigs[(size_t)rid].c_str();
}

static inline bcf1_t *bcf_init() { return new bcf1_t(); }
static inline void bcf_destroy(bcf1_t *r) { delete r; }
static inline int bcf_unpack(bcf1_t *, int) { return 0; }

static inline int bcf_read(htsFile *f, bcf_hdr_t *h, bcf1_t *r) {
  std::string line;
  while (true) {
    if (f->has_pending) {
      line = f->pending;
      f->has_pending = false;
    } else if (!shim_gets(f->fp, line)) {
      return -1;
    }
    if (!line.empty()) break;
  }
  std::vector<std::string> c;
  size_t s = 0;
  while (true) {
    size_t t = line.find('\t', s);
    c.push_back(line.substr(s, t == std::string::npos ? std::string::npos : t - s));
    if (t == std::string::npos) break;
    s = t + 1;
  }
  if (c.size() < 8) return -2;
  auto it = h->contig_id.find(c[0]);
  if (it == h->contig_id.end()) {
    // htslib adds a dummy contig definition on the fly
    h->contig_id[c[0]] = (int)h->contigs.size();
    h->contigs.push_back(c[0]);
    it = h->contig_id.find(c[0]);
  }
  r->rid = it->second;
  r->pos = strtoll(c[1].c_str(), nullptr, 10) - 1;
  r->id_s = c[2];
  r->d.id = const_cast<char *>(r->id_s.c_str());
  r->allele_s.clear();
  r->allele_s.push_back(c[3]);
  if (c[4] != ".") {
    size_t p = 0;
    while (true) {
      size_t t = c[4].find(',', p);
      r->allele_s.push_back(c[4].substr(p, t == std::string::npos ? std::string::npos : t - p));
      if (t == std::string::npos) break;
      p = t + 1;
    }
  }
  r->n_allele = (uint32_t)r->allele_s.size();
  r->allele_p.clear();
  for (auto &a : r->allele_s) r->allele_p.push_back(const_cast<char *>(a.c_str()));
  r->d.allele = r->allele_p.data();
  if (c[5] == ".") {
    uint32_t bits = 0x7F800001u;  // bcf_float_missing
    memcpy(&r->qual, &bits, 4);
  } else {
    r->qual = (float)strtod(c[5].c_str(), nullptr);
  }
  r->info_s = c[7];
  r->format_s = c.size() > 8 ? c[8] : std::string();
  r->sample_s.clear();
  for (size_t i = 9; i < c.size(); ++i) r->sample_s.push_back(c[i]);
  return 0;
}

static inline int bcf_get_info_float(const bcf_hdr_t *h, bcf1_t *r, const char *tag, float **dst,
                                     int *ndst) {
  bool declared = false;
  for (const auto &hr : h->hrecs)
    if (hr.key == "INFO" && hr.id == tag) declared = true;
  if (!declared) return -1;
  const std::string &s = r->info_s;
  size_t tl = strlen(tag);
  size_t p = 0;
  while (p < s.size()) {
    size_t e = s.find(';', p);
    if (e == std::string::npos) e = s.size();
    if (e - p > tl && s.compare(p, tl, tag) == 0 && s[p + tl] == '=') {
      std::vector<float> vals;
      size_t q = p + tl + 1;
      while (q <= e) {
        size_t t = s.find(',', q);
        if (t == std::string::npos || t > e) t = e;
        std::string tok = s.substr(q, t - q);
        float fv;
        if (tok == "." || tok.empty()) {
          uint32_t bits = 0x7F800001u;
          memcpy(&fv, &bits, 4);
        } else {
          fv = (float)strtod(tok.c_str(), nullptr);
        }
        vals.push_back(fv);
        q = t + 1;
      }
      if (*ndst < (int)vals.size() || !*dst) {
        *dst = (float *)realloc(*dst, vals.size() * sizeof(float));
        *ndst = (int)vals.size();
      }
      memcpy(*dst, vals.data(), vals.size() * sizeof(float));
      return (int)vals.size();
    }
    p = e + 1;
  }
  return -3;
}

static inline int bcf_get_genotypes(const bcf_hdr_t *h, bcf1_t *r, int32_t **dst, int *ndst) {
  if (h->keep.empty()) return -1;
  // locate GT in FORMAT
  int gt_field = -1;
  {
    size_t p = 0;
    int idx = 0;
    const std::string &f = r->format_s;
    while (p <= f.size() && !f.empty()) {
      size_t t = f.find(':', p);
      if (t == std::string::npos) t = f.size();
      if (f.compare(p, t - p, "GT") == 0) {
        gt_field = idx;
        break;
      }
      ++idx;
      p = t + 1;
    }
  }
  if (gt_field < 0) return -3;
  std::vector<std::vector<int32_t>> per;
  size_t maxp = 0;
  for (int si : h->keep) {
    std::vector<int32_t> g;
    if ((size_t)si < r->sample_s.size()) {
      const std::string &col = r->sample_s[(size_t)si];
      size_t p = 0;
      for (int k = 0; k < gt_field; ++k) {
        size_t t = col.find(':', p);
        if (t == std::string::npos) {
          p = col.size();
          break;
        }
        p = t + 1;
      }
      size_t e = col.find(':', p);
      if (e == std::string::npos) e = col.size();
      int phased = 0;
      size_t q = p;
      while (q < e) {
        if (col[q] == '.') {
          g.push_back(0 | phased);
          ++q;
        } else {
          int v = 0;
          while (q < e && col[q] >= '0' && col[q] <= '9') v = v * 10 + (col[q++] - '0');
          g.push_back(((v + 1) << 1) | phased);
        }
        if (q < e) {
          phased = (col[q] == '|') ? 1 : 0;
          ++q;
        }
      }
    }
    if (g.empty()) g.push_back(0);
    maxp = std::max(maxp, g.size());
    per.push_back(g);
  }
  size_t n = per.size() * maxp;
  // one spare slot: variant.hpp:184 peeks at curr_gt[1] even when ploidy is 1
  *dst = (int32_t *)realloc(*dst, (n + 1) * sizeof(int32_t));
  *ndst = (int)n;
  for (size_t i = 0; i < per.size(); ++i)
    for (size_t j = 0; j < maxp; ++j)
      (*dst)[i * maxp + j] = j < per[i].size() ? per[i][j] : bcf_int32_vector_end;
  (*dst)[n] = bcf_int32_vector_end;
  return (int)n;
}
