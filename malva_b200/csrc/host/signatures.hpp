// Host-kept part of malva-geno (north star: "the C++ host keeps VCF parsing, FASTA reading and var_block
// signature enumeration").  This file: the variant record and the var_block signature enumerator, written
// from scratch to emit the packed batch (mg_packed_batch: 2-bit k-mer words, u32 CSR offsets) the C ABI consumes
// directly -- not the reference's map<int, map<int, vector<vector<string>>>> -- and to run in parallel over the
// variants of a batch, inside a block as well as across blocks.
//
// Behaviour follows the reference decision by decision:
//   block membership ("near")       var_block.hpp:417-423   (single-precision compare, see near())
//   overlap                         var_block.hpp:408-412
//   neighbour chains                var_block.hpp:436-677
//   haplotype allele combinations   var_block.hpp:709-786
//   signature construction          var_block.hpp:95-219
// What differs is only how the work is organised:
//   * genotypes are kept SPARSE: per variant the samples whose genotype is not the default (reference allele on
//     every haplotype, the variant's majority phasing flag); the haplotypes of a chain come from the chain members'
//     sparse lists -- scattered into per-sample keys / rows when every member defaults to phased genotypes (or in
//     haploid mode), merged sample by sample otherwise: work proportional to the carriers, not to the panel size
//     (27,934 samples in the SARS-CoV-2 example, a handful of carriers per record) -- plus one all-reference row for
//     everybody else;
//   * haplotypes are tuples of small allele ids; an allele id is the index of the first allele of the variant with
//     the same TEXT, which is exactly the identity the reference's unordered_set<vector<string_view>> and
//     Variant::get_allele_index (variant.hpp:228-240) use;
//   * identical genotype patterns are collapsed before unphased patterns are expanded into their 2^n haplotypes;
//   * duplicate signatures of an allele are dropped: the coverage of an allele is a max over its signatures
//     (main.cpp:176-177) and filter/table inserts are idempotent, so results cannot change;
//   * in a block sorted by position the walk that collects a variant's neighbours stops where nothing can be within
//     reach any more (the reference walks to the end of the block, to no effect).
// The order of the signatures of an allele is unspecified in the reference too (unordered_set iteration).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace mh {

// The handful of short lists every record carries (ALT alleles, frequencies, allele text ids, the few samples off the
// default genotype): std::vector's interface as far as it is used here, the first N elements inside the object.  A
// record of the usual kind then owns no heap block at all -- decoding allocated five per record, and giving them
// back cost 0.2 us per record on whichever thread dropped the batch (frees of blocks from other threads' arenas do
// not run side by side).
template <class T, size_t N>
class InlineVec {
 public:
  InlineVec() = default;
  InlineVec(const InlineVec &o) { copy_from(o); }
  InlineVec(InlineVec &&o) noexcept { move_from(o); }
  InlineVec &operator=(const InlineVec &o) {
    if (this != &o) {
      clear();
      copy_from(o);
    }
    return *this;
  }
  InlineVec &operator=(InlineVec &&o) noexcept {
    if (this != &o) {
      reset();
      move_from(o);
    }
    return *this;
  }
  ~InlineVec() { reset(); }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T &operator[](size_t i) { return p_[i]; }
  const T &operator[](size_t i) const { return p_[i]; }
  T *data() { return p_; }
  const T *data() const { return p_; }
  T *begin() { return p_; }
  T *end() { return p_ + n_; }
  const T *begin() const { return p_; }
  const T *end() const { return p_ + n_; }
  T &back() { return p_[n_ - 1]; }
  const T &back() const { return p_[n_ - 1]; }
  void clear() {
    for (uint32_t i = 0; i < n_; ++i) p_[i].~T();
    n_ = 0;
  }
  void reserve(size_t cap) {
    if (cap > cap_) grow(cap);
  }
  void push_back(const T &v) { emplace_back(v); }
  void push_back(T &&v) { emplace_back(std::move(v)); }
  template <class... A>
  T &emplace_back(A &&...a) {
    if (n_ == cap_) grow(2 * (size_t)cap_);
    T *q = new (p_ + n_) T(std::forward<A>(a)...);
    ++n_;
    return *q;
  }
  void resize(size_t n) {
    while (n_ > n) p_[--n_].~T();
    reserve(n);
    while (n_ < n) new (p_ + n_++) T();
  }
  void assign(size_t n, const T &v) {
    clear();
    reserve(n);
    while (n_ < n) new (p_ + n_++) T(v);
  }

 private:
  T *inline_ptr() { return reinterpret_cast<T *>(inl_); }
  void reset() {  // no elements, no heap block
    clear();
    if (p_ != inline_ptr()) ::operator delete(static_cast<void *>(p_));
    p_ = inline_ptr();
    cap_ = (uint32_t)N;
  }
  void copy_from(const InlineVec &o) {  // *this is empty
    reserve(o.n_);
    for (uint32_t i = 0; i < o.n_; ++i) new (p_ + i) T(o.p_[i]);
    n_ = o.n_;
  }
  void move_from(InlineVec &o) {  // *this is in the reset() state
    if (o.p_ == o.inline_ptr()) {
      for (uint32_t i = 0; i < o.n_; ++i) {
        new (p_ + i) T(std::move(o.p_[i]));
        o.p_[i].~T();
      }
    } else {
      p_ = o.p_;
      cap_ = o.cap_;
      o.p_ = o.inline_ptr();
      o.cap_ = (uint32_t)N;
    }
    n_ = o.n_;
    o.n_ = 0;
  }
  void grow(size_t cap) {
    T *q = static_cast<T *>(::operator new(cap * sizeof(T)));
    for (uint32_t i = 0; i < n_; ++i) {
      new (q + i) T(std::move(p_[i]));
      p_[i].~T();
    }
    if (p_ != inline_ptr()) ::operator delete(static_cast<void *>(p_));
    p_ = q;
    cap_ = (uint32_t)cap;
  }
  alignas(T) unsigned char inl_[N * sizeof(T)];
  T *p_ = inline_ptr();
  uint32_t n_ = 0, cap_ = (uint32_t)N;
};

struct GtEntry {  // genotype of one kept sample, as allele TEXT ids (see Variant::text_id)
  uint32_t sample;
  uint16_t h1, h2;
  uint8_t phased;
};

struct Variant {
  std::string seq_name;
  int ref_pos = 0;  // 0-based
  std::string idx;  // ID column
  std::string ref_sub;
  InlineVec<std::string, 2> alts;  // symbolic (<..>) alleles removed, upper-cased
  float quality = 0;
  // genotypes of the kept samples (variant.hpp:158-211): every sample without an entry is {0, 0, default_phased}
  uint32_t n_samples_ = 0;
  uint8_t default_phased = 1;
  InlineVec<GtEntry, 6> gts;  // ascending sample index
  int ref_size = 0, min_size = 0, max_size = 0;
  bool has_alts = true, is_present = true;
  InlineVec<float, 4> frequencies;
  InlineVec<uint16_t, 4> text_id;  // allele index -> first allele index with the same text

  int n_alleles() const { return (int)alts.size() + 1; }
  size_t n_samples() const { return n_samples_; }
  const std::string &allele(int i) const { return i == 0 ? ref_sub : alts[(size_t)i - 1]; }

  void set_sizes() {  // variant.hpp:108-124
    ref_size = (int)ref_sub.size();
    if (alts.empty()) {
      has_alts = false;
      return;
    }
    min_size = max_size = ref_size;
    for (const auto &a : alts) {
      min_size = std::min(min_size, (int)a.size());
      max_size = std::max(max_size, (int)a.size());
    }
    text_id.resize((size_t)n_alleles());
    for (int i = 0; i < n_alleles(); ++i) {
      int first = i;
      for (int j = 0; j < i; ++j)
        if (allele(j) == allele(i)) {
          first = j;
          break;
        }
      text_id[(size_t)i] = (uint16_t)first;
    }
  }
  // a GT index past the kept ALTs is undefined behaviour in the reference (variant.hpp:221); clamp it
  uint16_t text_id_of(int allele_index) const {
    if (allele_index >= n_alleles()) allele_index = n_alleles() - 1;
    return text_id[(size_t)allele_index];
  }
};

constexpr uint64_t SIG_IRREGULAR = 1ull << 63;  // mg_packed_batch: hi bit 63
constexpr uint64_t SIG_REF_ALLELE = 1ull << 62;  // hi bit 62

// Flattened signatures of a run of variants: variant -> allele slot -> signature -> k-mers (mg_packed_batch).
struct SignatureCsr {
  std::vector<uint64_t> kmers;  // {lo, hi} per k-mer
  std::vector<uint32_t> sig_kmer_off{0}, allele_sig_off{0}, var_allele_off{0};
  std::vector<float> freq;
  // k-mers that are not exactly k symbols of ACGT (var_block.hpp:178-193 at contig ends; N / IUPAC in the reference)
  std::string irr_pool;
  std::vector<uint64_t> irr_off{0};
  std::vector<uint32_t> irr_kmer;
  uint64_t n_variants() const { return var_allele_off.size() - 1; }
  uint64_t n_alleles() const { return allele_sig_off.size() - 1; }
  uint64_t n_sigs() const { return sig_kmer_off.size() - 1; }
  uint64_t n_kmers() const { return kmers.size() / 2; }
  uint64_t n_irregular() const { return irr_kmer.size(); }
  void clear() {
    kmers.clear();
    sig_kmer_off.assign(1, 0);
    allele_sig_off.assign(1, 0);
    var_allele_off.assign(1, 0);
    freq.clear();
    irr_pool.clear();
    irr_off.assign(1, 0);
    irr_kmer.clear();
  }
  bool is_ref_kmer(uint64_t i) const { return (kmers[2 * i + 1] & SIG_REF_ALLELE) != 0; }
  bool is_irregular(uint64_t i) const { return (kmers[2 * i + 1] & SIG_IRREGULAR) != 0; }
  // text of k-mer i (the `signatures` sub-command, the irregular inserts)
  std::string text(uint64_t i, int k) const {
    const uint64_t lo = kmers[2 * i], hi = kmers[2 * i + 1];
    if (hi & SIG_IRREGULAR) return irr_pool.substr(irr_off[lo], irr_off[lo + 1] - irr_off[lo]);
    std::string s((size_t)k, 'A');
    for (int j = 0; j < k; ++j) {
      const int sh = 2 * (k - 1 - j);
      const uint64_t code = sh >= 64 ? (hi >> (sh - 64)) : (lo >> sh);
      s[(size_t)j] = "ACGT"[code & 3];
    }
    return s;
  }
  // 8 symbols (first one in the low byte of `w`) -> their 2-bit codes, first symbol most significant, in 16 bits;
  // false if one of them is not A, C, G or T.  All eight at once: the code of a symbol is ((c >> 1) ^ (c >> 2)) & 3
  // (A 0, C 1, G 2, T 3), the symbol a code stands for is rebuilt bit by bit and compared with what was read.
  static bool pack8(uint64_t w, uint64_t &codes16) {
    constexpr uint64_t ONES = 0x0101010101010101ull;
    const uint64_t x = ((w >> 1) ^ (w >> 2)) & (3 * ONES);
    const uint64_t c0 = x & ONES, c1 = (x >> 1) & ONES, both = c0 & c1;
    const uint64_t expect = (0x40 * ONES) | (both ^ ONES) | ((c0 ^ c1) << 1) | (c1 << 2) | (both << 4);
    uint64_t r = __builtin_bswap64(x);  // first symbol in the high byte
    r = (r | (r >> 6)) & 0x000F000F000F000Full;
    r = (r | (r >> 12)) & 0x000000FF000000FFull;
    r = (r | (r >> 24)) & 0xFFFFull;
    codes16 = r;
    return expect == w;
  }
  void add_kmer(const char *s, size_t len, int k, bool is_ref) {
    uint64_t lo = 0, hi = 0;
    bool regular = (int)len == k && len <= 64;
    if (regular) {
      unsigned __int128 acc = 0;
      size_t j = 0;
      uint64_t w, c;
      for (; j + 8 <= len; j += 8) {
        memcpy(&w, s + j, 8);
        regular &= pack8(w, c);
        acc = (acc << 16) | c;
      }
      if (j < len) {  // the last 1..7 symbols, padded with 'A's that are shifted out again
        const size_t rest = len - j;
        w = 0x4141414141414141ull;
        memcpy(&w, s + j, rest);
        regular &= pack8(w, c);
        acc = (acc << (2 * rest)) | (c >> (16 - 2 * rest));
      }
      lo = regular ? (uint64_t)acc : 0;
      hi = regular ? (uint64_t)(acc >> 64) : 0;
    }
    if (!regular) {
      lo = irr_kmer.size();
      hi = SIG_IRREGULAR;
      irr_kmer.push_back((uint32_t)n_kmers());
      irr_pool.append(s, len);
      irr_off.push_back(irr_pool.size());
    }
    if (is_ref) hi |= SIG_REF_ALLELE;
    kmers.push_back(lo);
    kmers.push_back(hi);
  }
  // out = parts[0] ++ parts[1] ++ ... ; the copies (with their offset shifts) are independent and run through `par`
  template <class ParallelFor>
  static void concat(const std::vector<SignatureCsr> &parts, SignatureCsr &out, ParallelFor &&par) {
    const size_t n = parts.size();
    std::vector<uint64_t> v0(n + 1, 0), a0(n + 1, 0), s0(n + 1, 0), k0(n + 1, 0), i0(n + 1, 0), p0(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
      v0[i + 1] = v0[i] + parts[i].n_variants();
      a0[i + 1] = a0[i] + parts[i].n_alleles();
      s0[i + 1] = s0[i] + parts[i].n_sigs();
      k0[i + 1] = k0[i] + parts[i].n_kmers();
      i0[i + 1] = i0[i] + parts[i].n_irregular();
      p0[i + 1] = p0[i] + parts[i].irr_pool.size();
    }
    out.var_allele_off.resize(v0[n] + 1);
    out.allele_sig_off.resize(a0[n] + 1);
    out.sig_kmer_off.resize(s0[n] + 1);
    out.kmers.resize(2 * k0[n]);
    out.freq.resize(a0[n]);
    out.irr_kmer.resize(i0[n]);
    out.irr_off.resize(i0[n] + 1);
    out.irr_pool.resize(p0[n]);
    out.var_allele_off[0] = out.allele_sig_off[0] = out.sig_kmer_off[0] = 0;
    out.irr_off[0] = 0;
    par(n, [&](size_t i) {
      const SignatureCsr &p = parts[i];
      for (size_t j = 1; j < p.var_allele_off.size(); ++j) out.var_allele_off[v0[i] + j] = (uint32_t)(a0[i] + p.var_allele_off[j]);
      for (size_t j = 1; j < p.allele_sig_off.size(); ++j) out.allele_sig_off[a0[i] + j] = (uint32_t)(s0[i] + p.allele_sig_off[j]);
      for (size_t j = 1; j < p.sig_kmer_off.size(); ++j) out.sig_kmer_off[s0[i] + j] = (uint32_t)(k0[i] + p.sig_kmer_off[j]);
      if (!p.kmers.empty()) memcpy(out.kmers.data() + 2 * k0[i], p.kmers.data(), p.kmers.size() * 8);
      if (!p.freq.empty()) memcpy(out.freq.data() + a0[i], p.freq.data(), p.freq.size() * 4);
      for (size_t j = 0; j < p.irr_kmer.size(); ++j) {
        out.irr_kmer[i0[i] + j] = (uint32_t)(k0[i] + p.irr_kmer[j]);
        out.kmers[2 * (k0[i] + p.irr_kmer[j])] = i0[i] + j;  // index of the k-mer's text in the merged side pool
        out.irr_off[i0[i] + j + 1] = p0[i] + p.irr_off[j + 1];
      }
      if (!p.irr_pool.empty()) memcpy(&out.irr_pool[p0[i]], p.irr_pool.data(), p.irr_pool.size());
    });
  }
};

// var_block.hpp:417-423.  The reference adds ceil((float)k/2) to an int sum, so the whole comparison runs in
// single precision: beyond 2^24 bases positions are rounded before they are compared.  Kept on purpose.
inline bool variants_near(const Variant &a, const Variant &b, int k, int sum_to_add) {
  float lhs = (float)(a.ref_pos + a.ref_size - a.min_size - 1 + sum_to_add) + std::ceil((float)k / 2);
  return lhs >= (float)b.ref_pos;
}

// A chain of variant indices (block-local).  Chains are built and thrown away by the million; up to 14 members live
// inside the object, longer ones (dense panels) move to the heap.
class Chain {
 public:
  Chain() = default;
  explicit Chain(int first) { push_back(first); }
  Chain(const Chain &o) { assign(o.p_, o.n_); }
  Chain(Chain &&o) noexcept { take(o); }
  Chain &operator=(const Chain &o) {
    if (this != &o) assign(o.p_, o.n_);
    return *this;
  }
  Chain &operator=(Chain &&o) noexcept {
    if (this != &o) {
      release();
      take(o);
    }
    return *this;
  }
  ~Chain() { release(); }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  int operator[](size_t i) const { return p_[i]; }
  int back() const { return p_[n_ - 1]; }
  const int *data() const { return p_; }
  const int *begin() const { return p_; }
  const int *end() const { return p_ + n_; }
  void clear() { n_ = 0; }
  void pop_back() { --n_; }
  void push_back(int v) {
    if (n_ == cap_) grow(2 * cap_);
    p_[n_++] = v;
  }
  void append(const int *first, size_t n) {
    if (n_ + n > cap_) grow(std::max<size_t>(2 * cap_, n_ + n));
    memcpy(p_ + n_, first, n * sizeof(int));
    n_ += (uint32_t)n;
  }
  void append_reversed(const Chain &o) {
    for (size_t i = o.n_; i > 0; --i) push_back(o.p_[i - 1]);
  }

 private:
  static constexpr uint32_t INLINE = 14;
  void assign(const int *src, uint32_t n) {
    n_ = 0;
    append(src, n);
  }
  void take(Chain &o) {
    if (o.p_ == o.inl_) {
      p_ = inl_;
      cap_ = INLINE;
      memcpy(inl_, o.inl_, o.n_ * sizeof(int));
    } else {
      p_ = o.p_;
      cap_ = o.cap_;
      o.p_ = o.inl_;
      o.cap_ = INLINE;
    }
    n_ = o.n_;
    o.n_ = 0;
  }
  void release() {
    if (p_ != inl_) delete[] p_;
    p_ = inl_;
    cap_ = INLINE;
  }
  void grow(size_t cap) {
    int *q = new int[cap];
    memcpy(q, p_, n_ * sizeof(int));
    if (p_ != inl_) delete[] p_;
    p_ = q;
    cap_ = (uint32_t)cap;
  }
  int inl_[INLINE];
  int *p_ = inl_;
  uint32_t n_ = 0, cap_ = INLINE;
};

// A var_block (var_block.hpp:40-90) as a VIEW: n consecutive variants of the batch they were decoded into.  Grouping
// a batch into blocks moves no record and allocates nothing per block.
class VarBlock {
 public:
  VarBlock() = default;
  VarBlock(int k, const Variant *first, size_t n, const std::string *contig_name) : contig(contig_name), k_(k), vars_(first), n_(n) {
    // big blocks (a dense panel is ONE block of thousands of variants): if the records are sorted by position, the
    // walk away from a variant can stop where nothing can be within reach any more (side_chains)
    if (n_ >= 64) {
      can_cut_ = true;
      for (size_t i = 0; i < n_; ++i) {
        if (i && vars_[i].ref_pos < vars_[i - 1].ref_pos) can_cut_ = false;
        max_gain_ = std::max(max_gain_, vars_[i].ref_size - vars_[i].min_size);
      }
    }
  }
  bool empty() const { return n_ == 0; }
  size_t size() const { return n_; }
  const Variant &operator[](size_t i) const { return vars_[i]; }
  const std::string *contig = nullptr;  // name of the contig the block lies on (last_seq_name at flush time)

  // Work arrays reused from variant to variant (one per enumerating thread): no allocation per sample or per chain.
  struct SigRec {  // one signature of the current variant: its k-mers lie back to back in Scratch::text
    uint32_t allele, text_off, text_len, n_kmers;
  };
  struct Scratch {
    std::vector<uint16_t> pat, cand, haps;
    std::vector<uint32_t> order, table;
    std::vector<size_t> cursor;
    size_t n_haps = 0;
    std::string text, kmer;                // signature text of the current variant; the k-mer being built
    std::vector<SigRec> sigs;              // its signatures
    std::vector<Chain> left, right, forks, full;  // the chains around the current variant
    std::vector<int> reach_l, reach_r, fork_reach;
    std::vector<uint64_t> key1, key2;      // haplotypes_scatter: the two haplotypes of a sample, one member per byte
    std::vector<uint8_t> flag;             //   bit 0: the sample has an entry at some member, bit 1: an unphased one
    std::vector<uint32_t> touched;         //   the samples with bit 0 set
    std::vector<uint32_t> row_of;          // scatter_rows: sample -> its row + 1 (0: none yet)
    std::vector<uint64_t> row_hash;        //   hash of a row's non-zero entries, built up as they arrive
  };

  // Signatures of the variants [begin, end) of the block, appended to `out`: one variant entry per block member, in
  // order; allele slot a = allele index a (slots of duplicate-text alleles stay empty, like in the reference).
  void enumerate(const std::string &reference, bool haploid, size_t begin, size_t end, Scratch &sc, SignatureCsr &out) const {
    for (size_t vi = begin; vi < end && vi < n_; ++vi) {
      const Variant &v = vars_[vi];
      sc.text.clear();
      sc.sigs.clear();
      // var_block.hpp:104 -- no signatures for absent variants or variants within k of a contig end
      if (v.is_present && v.ref_pos >= k_ && v.ref_pos <= (int)reference.size() - k_)
        signatures_of((int)vi, reference, haploid, sc);
      // by allele, duplicates dropped (equal text + equal k-mer count = equal k-mer list: every k-mer of a
      // multi-k-mer signature is k long)
      auto text_of = [&](const SigRec &r) { return std::string_view(sc.text.data() + r.text_off, r.text_len); };
      if (sc.sigs.size() > 1)
        std::sort(sc.sigs.begin(), sc.sigs.end(), [&](const SigRec &x, const SigRec &y) {
          if (x.allele != y.allele) return x.allele < y.allele;
          if (x.n_kmers != y.n_kmers) return x.n_kmers < y.n_kmers;
          return text_of(x) < text_of(y);
        });
      size_t si = 0;
      for (int a = 0; a < v.n_alleles(); ++a) {
        const SigRec *prev = nullptr;
        for (; si < sc.sigs.size() && sc.sigs[si].allele == (uint32_t)a; ++si) {
          const SigRec &r = sc.sigs[si];
          if (prev && prev->n_kmers == r.n_kmers && text_of(*prev) == text_of(r)) continue;
          prev = &r;
          const size_t klen = r.n_kmers > 1 ? (size_t)k_ : r.text_len;
          for (uint32_t q = 0; q < r.n_kmers; ++q) out.add_kmer(sc.text.data() + r.text_off + q * klen, klen, k_, a == 0);
          out.sig_kmer_off.push_back((uint32_t)out.n_kmers());
        }
        out.allele_sig_off.push_back((uint32_t)out.n_sigs());
        out.freq.push_back((size_t)a < v.frequencies.size() ? v.frequencies[(size_t)a] : 0.0f);
      }
      out.var_allele_off.push_back((uint32_t)out.n_alleles());
    }
  }

 private:
  // one index per distinct row, in order of first occurrence.  Thousands of panel samples share a handful of
  // genotype patterns, so this is a small open-addressing hash set keyed by row content, not a sort.
  static void distinct_rows(const std::vector<uint16_t> &rows, size_t n_rows, size_t width, std::vector<uint32_t> &order,
                            std::vector<uint32_t> &table) {
    order.clear();
    const uint16_t *base = rows.data();
    const size_t bytes = width * sizeof(uint16_t);
    size_t cap = 64;
    table.assign(cap, 0xFFFFFFFFu);
    auto hash_row = [&](const uint16_t *r) {
      uint64_t h = 0xCBF29CE484222325ull;
      for (size_t i = 0; i < width; ++i) h = (h ^ r[i]) * 0x100000001B3ull;
      return h ^ (h >> 29);
    };
    for (size_t i = 0; i < n_rows; ++i) {
      const uint16_t *r = base + i * width;
      size_t slot = (size_t)hash_row(r) & (cap - 1);
      bool found = false;
      while (table[slot] != 0xFFFFFFFFu) {
        if (memcmp(base + (size_t)table[slot] * width, r, bytes) == 0) {
          found = true;
          break;
        }
        slot = (slot + 1) & (cap - 1);
      }
      if (found) continue;
      table[slot] = (uint32_t)i;
      order.push_back((uint32_t)i);
      if (order.size() * 2 > cap) {  // grow and re-insert the distinct rows
        cap *= 4;
        table.assign(cap, 0xFFFFFFFFu);
        for (uint32_t j : order) {
          size_t s2 = (size_t)hash_row(base + (size_t)j * width) & (cap - 1);
          while (table[s2] != 0xFFFFFFFFu) s2 = (s2 + 1) & (cap - 1);
          table[s2] = j;
        }
      }
    }
  }

  static bool overlapping(const Variant &a, const Variant &b) {  // var_block.hpp:408-412
    return a.ref_pos <= b.ref_pos && b.ref_pos < a.ref_pos + a.ref_size;
  }

  // var_block.hpp:436-525 (dir = +1) and 534-624 (dir = -1): every chain of mutually compatible neighbours that
  // stays within reach of the mid variant.  Pairs are always tested in genomic order.
  void side_chains(int i, int dir, std::vector<Chain> &chains, std::vector<int> &reach, Scratch &sc) const {
    const Variant &mid = vars_[(size_t)i];
    chains.clear();
    reach.clear();  // bases the chain's deletions add to the reach of the mid variant
    auto ovl = [&](int near_mid, int far) {
      return dir > 0 ? overlapping(vars_[(size_t)near_mid], vars_[(size_t)far])
                     : overlapping(vars_[(size_t)far], vars_[(size_t)near_mid]);
    };
    auto in_reach = [&](int j, int extra) {
      return dir > 0 ? variants_near(mid, vars_[(size_t)j], k_, extra) : variants_near(vars_[(size_t)j], mid, k_, extra);
    };
    auto gain = [&](int j) { return vars_[(size_t)j].ref_size - vars_[(size_t)j].min_size; };
    // Where the walk may stop (sorted blocks only; the reference walks to the end of the block, to no effect): a
    // variant joins or forks a chain only if it is within reach of the mid variant (variants_near), reaches only
    // grow by what joins, so once a variant lies beyond the largest reach -- with a margin for the single-precision
    // comparison of variants_near -- so does everything behind it.
    constexpr long MARGIN = 512;
    const long half = (long)std::ceil((float)k_ / 2);
    long max_reach = 0;
    bool halt = false;
    for (int j = i + dir; j >= 0 && j < (int)n_ && !halt; j += dir) {
      if (can_cut_) {
        const long pj = vars_[(size_t)j].ref_pos;
        if (dir > 0 ? pj > (long)mid.ref_pos + mid.ref_size - mid.min_size - 1 + max_reach + half + MARGIN
                    : pj + max_gain_ - 1 + max_reach + half + MARGIN < (long)mid.ref_pos)
          break;
      }
      if (!vars_[(size_t)j].is_present || ovl(i, j)) continue;
      if (chains.empty()) {
        if (in_reach(j, 0)) {
          chains.emplace_back(j);
          reach.push_back(gain(j));
          max_reach = std::max<long>(max_reach, reach.back());
        }
        continue;
      }
      bool compatible = false;  // with the tail of at least one chain
      for (size_t c = 0; c < chains.size(); ++c) {
        if (ovl(chains[c].back(), j)) continue;
        compatible = true;
        if (in_reach(j, reach[c])) {
          chains[c].push_back(j);
          reach[c] += gain(j);
          max_reach = std::max<long>(max_reach, reach[c]);
        }
      }
      if (compatible) continue;
      // it clashes with every tail: fork each chain, cut back to the part it is compatible with
      std::vector<Chain> &forks = sc.forks;
      std::vector<int> &fork_reach = sc.fork_reach;
      forks.clear();
      fork_reach.clear();
      for (size_t c = 0; c < chains.size(); ++c) {
        Chain f = chains[c];
        int r = reach[c];
        while (!f.empty() && ovl(f.back(), j)) {
          r -= gain(f.back());
          f.pop_back();
        }
        f.push_back(j);
        if (in_reach(j, r)) {
          forks.push_back(std::move(f));
          fork_reach.push_back(r + gain(j));
        }
      }
      if (forks.empty()) halt = true;  // too far for any chain: nothing further can be added (var_block.hpp:516-519)
      for (size_t c = 0; c < forks.size(); ++c) {
        chains.push_back(std::move(forks[c]));
        reach.push_back(fork_reach[c]);
        max_reach = std::max<long>(max_reach, fork_reach[c]);
      }
    }
  }

  // var_block.hpp:631-677: left chains (reversed into genomic order) x right chains around the mid variant, in sc.full
  const std::vector<Chain> &full_chains(int i, Scratch &sc) const {
    std::vector<Chain> &out = sc.full;
    out.clear();
    if (n_ == 1) {  // (a block of one: nothing to chain)
      out.emplace_back(i);
      return out;
    }
    side_chains(i, +1, sc.right, sc.reach_r, sc);
    side_chains(i, -1, sc.left, sc.reach_l, sc);
    if (sc.left.empty()) sc.left.emplace_back();
    if (sc.right.empty()) sc.right.emplace_back();
    for (const Chain &l : sc.left)
      for (const Chain &r : sc.right) {
        out.emplace_back();
        Chain &c = out.back();
        c.append_reversed(l);
        c.push_back(i);
        c.append(r.data(), r.size());
      }
    return out;
  }

  // var_block.hpp:734-786: the distinct haplotypes (one allele per chain member) carried by the samples of the
  // central variant, as sc.n_haps rows of chain.size() text-ids in sc.haps.  Unphased patterns contribute every way
  // of picking one of the two alleles at each site (combine_haplotypes, var_block.hpp:709-728).
  // The genotype patterns [h1 ids | h2 ids | phased] come from a merge of the members' sparse genotype lists: one
  // row per sample that deviates from the default at some member, plus one all-reference row standing for every
  // other sample (whatever its phasing flags: with no heterozygous site it yields that one haplotype).
  void haplotypes(const Chain &chain, int central, bool haploid, Scratch &sc) const {
    if (chain.size() <= 8 && (haplotypes_scatter(chain, central, haploid, sc) || haplotypes_small(chain, central, haploid, sc))) return;
    const size_t n = chain.size(), W = 2 * n + 1;
    const uint32_t central_samples = (uint32_t)vars_[(size_t)central].n_samples();
    sc.cursor.assign(n, 0);
    size_t max_rows = 1;  // one row per sample with an entry somewhere, + the all-reference row
    for (size_t m = 0; m < n; ++m) max_rows += vars_[(size_t)chain[m]].gts.size();
    if (sc.pat.size() < max_rows * W) sc.pat.resize(max_rows * W);
    size_t n_rows = 0;
    const bool scattered = scatter_rows(chain, haploid, sc, central_samples, n_rows);
    while (!scattered) {
      uint32_t s = 0xFFFFFFFFu;  // the next sample with an entry at some member
      for (size_t m = 0; m < n; ++m) {
        const Variant &v = vars_[(size_t)chain[m]];
        if (sc.cursor[m] < v.gts.size()) s = std::min(s, v.gts[sc.cursor[m]].sample);
      }
      if (s >= central_samples) break;  // (the reference walks the samples of the central variant)
      uint16_t *p = sc.pat.data() + n_rows++ * W;
      uint16_t ph = 1;
      for (size_t m = 0; m < n; ++m) {
        const Variant &v = vars_[(size_t)chain[m]];
        if (sc.cursor[m] < v.gts.size() && v.gts[sc.cursor[m]].sample == s) {
          const GtEntry &g = v.gts[sc.cursor[m]++];
          p[m] = g.h1;
          p[n + m] = haploid ? g.h1 : g.h2;
          if (!g.phased) ph = 0;
        } else {
          p[m] = p[n + m] = 0;
          if (s < v.n_samples() && !v.default_phased) ph = 0;  // (a variant with fewer samples: reference, phased)
        }
      }
      p[2 * n] = haploid ? 1 : ph;
    }
    if (!scattered) {
      if (n_rows < central_samples) {  // samples at their default everywhere
        uint16_t *p = sc.pat.data() + n_rows++ * W;
        std::fill(p, p + 2 * n, (uint16_t)0);
        p[2 * n] = 1;
      }
      distinct_rows(sc.pat, n_rows, W, sc.order, sc.table);
    }
    sc.cand.clear();
    for (uint32_t row : sc.order) {
      const uint16_t *q = sc.pat.data() + (size_t)row * W;
      if (q[2 * n]) {
        sc.cand.insert(sc.cand.end(), q, q + n);
        if (!haploid) sc.cand.insert(sc.cand.end(), q + n, q + 2 * n);
      } else {
        size_t het[64], n_het = 0;  // sites where the two alleles differ
        for (size_t m = 0; m < n && n_het < 64; ++m)
          if (q[m] != q[n + m]) het[n_het++] = m;
        for (uint64_t mask = 0; mask < (1ull << n_het); ++mask) {
          const size_t o = sc.cand.size();
          sc.cand.insert(sc.cand.end(), q, q + n);
          for (size_t b = 0; b < n_het; ++b)
            if ((mask >> b) & 1) sc.cand[o + het[b]] = q[n + het[b]];
        }
      }
    }
    const size_t n_cand = sc.cand.size() / n;
    distinct_rows(sc.cand, n_cand, n, sc.order, sc.table);
    sc.haps.resize(sc.order.size() * n);
    for (size_t i = 0; i < sc.order.size(); ++i)
      memcpy(sc.haps.data() + i * n, sc.cand.data() + (size_t)sc.order[i] * n, n * sizeof(uint16_t));
    sc.n_haps = sc.order.size();
  }

  // The genotype patterns of haplotypes() without the merge -- same precondition as haplotypes_scatter(): all members
  // default to PHASED reference genotypes, or haploid mode.  The members' sparse lists are scattered one after the
  // other into the rows of the samples they mention (a row is created, zeroed, at a sample's first entry), and a
  // row's hash is built up from its non-zero entries as they arrive, so making the rows distinct afterwards costs a
  // table probe per row instead of a pass over its 2n+1 entries.  A chain of 19 members over a 27,934-sample panel
  // (1,100 rows, 27 distinct): 350 k cycles with the merge, a fifth of that here.  Fills sc.pat (rows of W = 2n+1, in
  // order of first mention), n_rows, sc.order (the distinct rows); false -- nothing usable written -- otherwise.
  bool scatter_rows(const Chain &chain, bool haploid, Scratch &sc, uint32_t central_samples, size_t &n_rows) const {
    const size_t n = chain.size(), W = 2 * n + 1;
    if (!haploid)
      for (size_t m = 0; m < n; ++m)
        if (!vars_[(size_t)chain[m]].default_phased) return false;
    if (sc.row_of.size() < central_samples) sc.row_of.resize(central_samples, 0);
    sc.touched.clear();
    sc.row_hash.clear();
    uint16_t *const pat = sc.pat.data();  // (sized by the caller: one row per entry + 1)
    auto mix = [](uint64_t x) {
      x ^= x >> 30;
      x *= 0xBF58476D1CE4E5B9ull;
      x ^= x >> 27;
      x *= 0x94D049BB133111EBull;
      return x ^ (x >> 31);
    };
    n_rows = 0;
    for (size_t m = 0; m < n; ++m)
      for (const GtEntry &g : vars_[(size_t)chain[m]].gts) {
        if (g.sample >= central_samples) break;  // (the reference walks the samples of the central variant)
        uint32_t r = sc.row_of[g.sample];
        if (!r) {
          r = (uint32_t)++n_rows;
          sc.row_of[g.sample] = r;
          sc.touched.push_back(g.sample);
          uint16_t *fresh = pat + (size_t)(r - 1) * W;
          memset(fresh, 0, 2 * n * sizeof(uint16_t));
          fresh[2 * n] = 1;
          sc.row_hash.push_back(0);
        }
        uint16_t *p = pat + (size_t)(r - 1) * W;
        const uint16_t h2 = haploid ? g.h1 : g.h2;
        p[m] = g.h1;
        p[n + m] = h2;
        if (!haploid && !g.phased) p[2 * n] = 0;
        if (g.h1 | h2) sc.row_hash[r - 1] ^= mix(((uint64_t)(m + 1) << 32) | ((uint64_t)g.h1 << 16) | h2);
      }
    for (uint32_t smp : sc.touched) sc.row_of[smp] = 0;
    if (n_rows < central_samples) {  // samples at their default everywhere
      uint16_t *p = pat + n_rows++ * W;
      memset(p, 0, 2 * n * sizeof(uint16_t));
      p[2 * n] = 1;
      sc.row_hash.push_back(0);
    }
    // one index per distinct row, in order of first occurrence
    std::vector<uint32_t> &order = sc.order, &table = sc.table;
    order.clear();
    size_t cap = 64;
    table.assign(cap, 0xFFFFFFFFu);
    const size_t bytes = W * sizeof(uint16_t);
    auto slot_of = [&](size_t row) {
      const uint64_t h = sc.row_hash[row] ^ (pat[row * W + 2 * n] ? 0 : 0x9E3779B97F4A7C15ull);
      return (size_t)(h ^ (h >> 29)) & (cap - 1);
    };
    for (size_t i = 0; i < n_rows; ++i) {
      size_t slot = slot_of(i);
      bool found = false;
      while (table[slot] != 0xFFFFFFFFu) {
        if (memcmp(pat + (size_t)table[slot] * W, pat + i * W, bytes) == 0) {
          found = true;
          break;
        }
        slot = (slot + 1) & (cap - 1);
      }
      if (found) continue;
      table[slot] = (uint32_t)i;
      order.push_back((uint32_t)i);
      if (order.size() * 2 > cap) {  // grow and re-insert the distinct rows
        cap *= 4;
        table.assign(cap, 0xFFFFFFFFu);
        for (uint32_t j : order) {
          size_t s2 = slot_of(j);
          while (table[s2] != 0xFFFFFFFFu) s2 = (s2 + 1) & (cap - 1);
          table[s2] = j;
        }
      }
    }
    return true;
  }

  // haplotypes_small() without the merge, for chains whose members all default to PHASED reference genotypes (or in
  // haploid mode, where phasing is not looked at): a sample without an entry at a member then simply carries allele 0
  // there, so the members' sparse lists can be scattered into per-sample keys one list after the other -- work
  // proportional to the entries, not to samples x members.  Same result as the merge; false (nothing written) when
  // the chain does not qualify or something does not fit.
  bool haplotypes_scatter(const Chain &chain, int central, bool haploid, Scratch &sc) const {
    const size_t n = chain.size();
    const uint32_t central_samples = (uint32_t)vars_[(size_t)central].n_samples();
    if (!haploid)
      for (size_t m = 0; m < n; ++m)
        if (!vars_[(size_t)chain[m]].default_phased) return false;
    if (sc.flag.size() < central_samples) {
      sc.key1.resize(central_samples, 0);
      sc.key2.resize(central_samples, 0);
      sc.flag.resize(central_samples, 0);
    }
    sc.touched.clear();
    bool fits = true;
    for (size_t m = 0; m < n && fits; ++m)
      for (const GtEntry &g : vars_[(size_t)chain[m]].gts) {
        if (g.sample >= central_samples) break;  // (the reference walks the samples of the central variant)
        if ((g.h1 | g.h2) > 255) {
          fits = false;
          break;
        }
        uint8_t &f = sc.flag[g.sample];
        if (!f) sc.touched.push_back(g.sample);
        f |= g.phased ? 1 : 3;
        sc.key1[g.sample] |= (uint64_t)g.h1 << (8 * m);
        sc.key2[g.sample] |= (uint64_t)(haploid ? g.h1 : g.h2) << (8 * m);
      }
    // the distinct keys: while there are few of them (the rule: 3-4 haplotypes among ~20 candidates) a candidate is
    // compared with the ones kept so far; beyond that everything is kept and made distinct by sort + unique at the end
    constexpr size_t MAX_KEYS = 512, FEW = 12;
    uint64_t keys[MAX_KEYS];
    size_t n_keys = 0;
    bool distinct = true;  // no two of keys[0, n_keys) are equal
    auto add_key = [&](uint64_t key) {  // false: no room
      if (distinct) {
        for (size_t i = 0; i < n_keys; ++i)
          if (keys[i] == key) return true;
        if (n_keys >= FEW) distinct = false;
      }
      if (n_keys >= MAX_KEYS) return false;
      keys[n_keys++] = key;
      return true;
    };
    for (uint32_t smp : sc.touched) {  // (every touched slot is cleared again, whatever happens)
      const uint64_t k1 = sc.key1[smp], k2 = sc.key2[smp];
      const bool ph = (sc.flag[smp] & 2) == 0;
      sc.key1[smp] = sc.key2[smp] = 0;
      sc.flag[smp] = 0;
      if (!fits) continue;
      if (haploid) {
        fits = add_key(k1);
      } else if (ph) {
        fits = add_key(k1) && add_key(k2);
      } else {  // unphased: every way of picking one of the two alleles at each heterozygous site
        uint64_t diff = k1 ^ k2, het_mask[8];
        size_t n_het = 0;
        for (size_t m = 0; m < n; ++m)
          if ((diff >> (8 * m)) & 0xFF) het_mask[n_het++] = 0xFFull << (8 * m);
        for (uint64_t mask = 0; mask < (1ull << n_het) && fits; ++mask) {
          uint64_t key = k1;
          for (size_t b = 0; b < n_het; ++b)
            if ((mask >> b) & 1) key = (key & ~het_mask[b]) | (k2 & het_mask[b]);
          fits = add_key(key);
        }
      }
    }
    if (!fits) return false;
    if (sc.touched.size() < central_samples && !add_key(0)) return false;  // samples at their default everywhere
    if (!distinct) {
      std::sort(keys, keys + n_keys);
      n_keys = (size_t)(std::unique(keys, keys + n_keys) - keys);
    }
    sc.haps.resize(n_keys * n);
    for (size_t i = 0; i < n_keys; ++i)
      for (size_t m = 0; m < n; ++m) sc.haps[i * n + m] = (uint16_t)((keys[i] >> (8 * m)) & 0xFF);
    sc.n_haps = n_keys;
    return true;
  }

  // The same set of haplotypes for the common case -- a chain of at most 8 members whose allele ids all fit a byte:
  // a haplotype is then one u64 (member m in byte m), the candidates of all samples are collected in a small array and
  // made distinct by sort + unique.  (The ORDER of sc.haps does not matter: VarBlock::enumerate sorts the signatures
  // they produce.)  Returns false -- nothing written -- when an id or the number of candidates does not fit.
  bool haplotypes_small(const Chain &chain, int central, bool haploid, Scratch &sc) const {
    const size_t n = chain.size();
    constexpr size_t MAX_KEYS = 512;
    uint64_t keys[MAX_KEYS];
    size_t n_keys = 0, cursor[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_rows = 0;
    const Variant *member[8];
    for (size_t m = 0; m < n; ++m) member[m] = &vars_[(size_t)chain[m]];
    const uint32_t central_samples = (uint32_t)vars_[(size_t)central].n_samples();
    while (true) {
      uint32_t s = 0xFFFFFFFFu;  // the next sample with an entry at some member
      for (size_t m = 0; m < n; ++m)
        if (cursor[m] < member[m]->gts.size()) s = std::min(s, member[m]->gts[cursor[m]].sample);
      if (s >= central_samples) break;
      ++n_rows;
      uint64_t k1 = 0, k2 = 0;
      bool ph = true;
      for (size_t m = 0; m < n; ++m) {
        const Variant &v = *member[m];
        if (cursor[m] < v.gts.size() && v.gts[cursor[m]].sample == s) {
          const GtEntry &g = v.gts[cursor[m]++];
          if ((g.h1 | g.h2) > 255) return false;
          k1 |= (uint64_t)g.h1 << (8 * m);
          k2 |= (uint64_t)(haploid ? g.h1 : g.h2) << (8 * m);
          if (!g.phased) ph = false;
        } else if (s < v.n_samples() && !v.default_phased) {
          ph = false;
        }
      }
      if (haploid) {
        if (n_keys + 1 > MAX_KEYS) return false;
        keys[n_keys++] = k1;
      } else if (ph) {
        if (n_keys + 2 > MAX_KEYS) return false;
        keys[n_keys++] = k1;
        keys[n_keys++] = k2;
      } else {  // unphased: every way of picking one of the two alleles at each heterozygous site
        uint64_t diff = k1 ^ k2, het_mask[8];
        size_t n_het = 0;
        for (size_t m = 0; m < n; ++m)
          if ((diff >> (8 * m)) & 0xFF) het_mask[n_het++] = 0xFFull << (8 * m);
        if (n_keys + ((size_t)1 << n_het) > MAX_KEYS) return false;
        for (uint64_t mask = 0; mask < (1ull << n_het); ++mask) {
          uint64_t key = k1;
          for (size_t b = 0; b < n_het; ++b)
            if ((mask >> b) & 1) key = (key & ~het_mask[b]) | (k2 & het_mask[b]);
          keys[n_keys++] = key;
        }
      }
    }
    if (n_rows < central_samples) {  // samples at their default everywhere: the all-reference haplotype
      if (n_keys + 1 > MAX_KEYS) return false;
      keys[n_keys++] = 0;
    }
    std::sort(keys, keys + n_keys);
    n_keys = (size_t)(std::unique(keys, keys + n_keys) - keys);
    sc.haps.resize(n_keys * n);
    for (size_t i = 0; i < n_keys; ++i)
      for (size_t m = 0; m < n; ++m) sc.haps[i * n + m] = (uint16_t)((keys[i] >> (8 * m)) & 0xFF);
    sc.n_haps = n_keys;
    return true;
  }

  // A block of one variant (the common case away from dense regions): the chain is the variant itself and its
  // haplotypes are simply the alleles somebody carries -- the reference allele if some sample has no entry, h1 (and
  // h2 unless haploid) of every entry; phasing cannot matter with a single site.  Same set as haplotypes() gives.
  bool single_site_haplotypes(const Variant &v, bool haploid, Scratch &sc) const {
    if (v.n_alleles() > 64) return false;
    uint64_t present = v.gts.size() < v.n_samples() ? 1ull : 0ull;
    for (const GtEntry &g : v.gts) {
      if (g.sample >= v.n_samples()) break;
      present |= 1ull << g.h1;
      if (!haploid) present |= 1ull << g.h2;
    }
    sc.haps.clear();
    for (uint64_t m = present; m; m &= m - 1) sc.haps.push_back((uint16_t)__builtin_ctzll(m));
    sc.n_haps = sc.haps.size();
    return true;
  }

  // var_block.hpp:114-216
  void signatures_of(int vi, const std::string &reference, bool haploid, Scratch &sc) const {
    if (n_ == 1 && single_site_haplotypes(vars_[0], haploid, sc)) {
      const int self = 0;
      chain_signatures(&self, 1, vi, reference, sc);
      return;
    }
    for (const Chain &chain : full_chains(vi, sc)) {
      haplotypes(chain, vi, haploid, sc);
      chain_signatures(chain.data(), chain.size(), vi, reference, sc);
    }
  }

  // the signatures of variant vi within one chain, one per haplotype in sc.haps (var_block.hpp:120-216).  The text of
  // a haplotype -- alleles of the members with the reference between them -- is put together with plain copies in
  // sc.kmer, then extended with reference text or cut on either side straight into sc.text.
  void chain_signatures(const int *chain, size_t chain_n, int vi, const std::string &reference, Scratch &sc) const {
    const Variant &v = vars_[(size_t)vi];
    struct Piece {  // a stretch of the reference, clamped like std::string::append(str, pos, len)
      const char *p;
      size_t n;
    };
    auto ref_piece = [&](long pos, long len) -> Piece {
      if (len <= 0 || pos < 0 || pos >= (long)reference.size()) return Piece{nullptr, 0};  // (the reference would throw on pos > size)
      return Piece{reference.data() + pos, std::min<size_t>((size_t)len, reference.size() - (size_t)pos)};
    };
    // reference text between consecutive members of the chain (var_block.hpp:682-702)
    Piece between_inl[16];
    std::vector<Piece> between_heap;
    Piece *between = between_inl;
    if (chain_n > 16) {
      between_heap.resize(chain_n);
      between = between_heap.data();
    }
    size_t mid_slot = 0, room = 0;
    for (size_t m = 0; m < chain_n; ++m) {
      const Variant &cur = vars_[(size_t)chain[m]];
      if (chain[m] == vi) mid_slot = m;
      room += (size_t)std::max(cur.max_size, cur.ref_size);
      if (m == 0) continue;
      const Variant &prev = vars_[(size_t)chain[m - 1]];
      between[m - 1] = ref_piece((long)prev.ref_pos + prev.ref_size, (long)cur.ref_pos - (prev.ref_pos + prev.ref_size));
      room += between[m - 1].n;
    }
    if (sc.kmer.size() < room) sc.kmer.resize(room);
    char *const hap = &sc.kmer[0];
    const Variant &first = vars_[(size_t)chain[0]], &last = vars_[(size_t)chain[chain_n - 1]];
    for (size_t hi = 0; hi < sc.n_haps; ++hi) {
      const uint16_t *h = sc.haps.data() + hi * chain_n;
      const int mid_id = h[mid_slot];
      const std::string &mid_allele = v.allele(mid_id);
      SigRec rec{(uint32_t)mid_id, (uint32_t)sc.text.size(), 0, 0};
      if (chain_n == 1 && (int)mid_allele.size() >= k_) {
        // an allele at least k long: every k-mer inside the allele itself (var_block.hpp:130-144)
        for (size_t p = 0; p + (size_t)k_ <= mid_allele.size(); ++p) {
          sc.text.append(mid_allele, p, (size_t)k_);
          ++rec.n_kmers;
        }
      } else {
        size_t len = 0;
        int mid_pos = 0;
        for (size_t m = 0; m < chain_n; ++m) {
          if (m == mid_slot) mid_pos = (int)len;
          const std::string &al = vars_[(size_t)chain[m]].allele(h[m]);
          memcpy(hap + len, al.data(), al.size());
          len += al.size();
          if (m + 1 < chain_n && between[m].n) {
            memcpy(hap + len, between[m].p, between[m].n);
            len += between[m].n;
          }
        }
        const int first_part = mid_pos + (int)mid_allele.size() / 2;
        const int second_part = (int)len - first_part;
        const int missing_prefix = k_ / 2 - first_part;
        const int missing_suffix = (int)std::ceil((float)k_ / 2) - second_part;
        // extend with reference text / cut, left then right
        size_t from = 0;
        if (missing_prefix >= 0) {
          const Piece ext = ref_piece((long)first.ref_pos - missing_prefix, missing_prefix);
          if (ext.n) sc.text.append(ext.p, ext.n);
        } else {
          from = std::min<size_t>(len, (size_t)(-missing_prefix));
        }
        if (missing_suffix >= 0) {
          sc.text.append(hap + from, len - from);
          const Piece ext = ref_piece((long)last.ref_pos + last.ref_size, missing_suffix);
          if (ext.n) sc.text.append(ext.p, ext.n);
        } else {
          const size_t body = len - from, cut = std::min<size_t>(body, (size_t)(-missing_suffix));
          sc.text.append(hap + from, body - cut);
        }
        rec.n_kmers = 1;
      }
      rec.text_len = (uint32_t)(sc.text.size() - rec.text_off);
      sc.sigs.push_back(rec);
    }
  }

  int k_ = 35;
  const Variant *vars_ = nullptr;
  size_t n_ = 0;
  bool can_cut_ = false;  // records sorted by position (checked for blocks of 64 and more)
  int max_gain_ = 0;      // max ref_size - min_size over the block
};

}  // namespace mh
