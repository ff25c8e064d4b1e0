// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for sdsl-lite 2.1.1's
// <sdsl/bit_vectors.hpp>, which is not installed in this image.  It provides
// exactly the container semantics the reference's bloom_filter.hpp relies on
// (bloom_filter.hpp:28,77-78,84,90,96-97,108-110,122,131-134,142-144):
//   bit_vector(size, 0)  / operator[] read+write / serialize / load
//   rank_support_v<1>(&bv) with operator()(i) = number of ones in [0, i)
//   int_vector<16>(n, 0, 16) with uint16_t elements (truncating store)
// Only the plumbing is restated; the reference's own arithmetic runs on top.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iomanip>  // main.cpp:98 relies on a transitive <iomanip>
#include <iostream>
#include <vector>

namespace sdsl {

class bit_vector {
 public:
  class reference {
   public:
    reference(uint64_t *w, unsigned b) : w_(w), b_(b) {}
    operator bool() const { return (*w_ >> b_) & 1ULL; }
    reference &operator=(bool v) {
      if (v) *w_ |= (1ULL << b_); else *w_ &= ~(1ULL << b_);
      return *this;
    }
    reference &operator=(int v) { return *this = (v != 0); }
    reference &operator=(const reference &o) { return *this = (bool)o; }
   private:
    uint64_t *w_;
    unsigned b_;
  };

  bit_vector(size_t n = 0, int v = 0) : n_(n), w_((n + 63) / 64, v ? ~0ULL : 0ULL) {}
  size_t size() const { return n_; }
  bool operator[](size_t i) const { return (w_[i >> 6] >> (i & 63)) & 1ULL; }
  reference operator[](size_t i) { return reference(&w_[i >> 6], (unsigned)(i & 63)); }
  const uint64_t *data() const { return w_.data(); }
  size_t words() const { return w_.size(); }

  // sdsl on-disk layout: u64 size-in-bits followed by the 64-bit words.
  size_t serialize(std::ostream &out) const {
    uint64_t n = n_;
    out.write(reinterpret_cast<const char *>(&n), 8);
    out.write(reinterpret_cast<const char *>(w_.data()), (std::streamsize)(w_.size() * 8));
    return 8 + w_.size() * 8;
  }
  void load(std::istream &in) {
    uint64_t n = 0;
    in.read(reinterpret_cast<char *>(&n), 8);
    n_ = n;
    w_.assign((n + 63) / 64, 0);
    in.read(reinterpret_cast<char *>(w_.data()), (std::streamsize)(w_.size() * 8));
  }

 private:
  size_t n_;
  std::vector<uint64_t> w_;
};

template <uint8_t pat = 1>
class rank_support_v {
 public:
  rank_support_v(const bit_vector *bv = nullptr) : bv_(bv) {
    if (!bv_) return;
    size_t nw = bv_->words();
    blk_.assign(nw / 8 + 2, 0);
    uint64_t acc = 0;
    const uint64_t *d = bv_->data();
    for (size_t w = 0; w < nw; ++w) {
      if ((w & 7) == 0) blk_[w >> 3] = acc;
      acc += (uint64_t)__builtin_popcountll(d[w]);
    }
    for (size_t b = (nw + 7) / 8; b < blk_.size(); ++b) blk_[b] = acc;
  }
  // number of set bits in positions [0, i)
  size_t operator()(size_t i) const {
    const uint64_t *d = bv_->data();
    size_t w = i >> 6;
    uint64_t r = blk_[w >> 3];
    for (size_t x = (w & ~(size_t)7); x < w; ++x) r += (uint64_t)__builtin_popcountll(d[x]);
    unsigned rem = (unsigned)(i & 63);
    if (rem) r += (uint64_t)__builtin_popcountll(d[w] & ((1ULL << rem) - 1));
    return r;
  }

 private:
  const bit_vector *bv_;
  std::vector<uint64_t> blk_;
};

template <uint8_t W>
class int_vector;

template <>
class int_vector<16> {
 public:
  int_vector(size_t n = 0, uint64_t def = 0, uint8_t /*width*/ = 16) : v_(n, (uint16_t)def) {}
  size_t size() const { return v_.size(); }
  uint16_t &operator[](size_t i) { return v_[i]; }
  const uint16_t &operator[](size_t i) const { return v_[i]; }
  // sdsl on-disk layout for a fixed-width int_vector: u64 size-in-bits + words.
  size_t serialize(std::ostream &out) const {
    uint64_t bits = (uint64_t)v_.size() * 16;
    out.write(reinterpret_cast<const char *>(&bits), 8);
    size_t nbytes = ((bits + 63) / 64) * 8;
    std::vector<char> buf(nbytes, 0);
    if (!v_.empty()) std::memcpy(buf.data(), v_.data(), v_.size() * 2);
    out.write(buf.data(), (std::streamsize)nbytes);
    return 8 + nbytes;
  }
  void load(std::istream &in) {
    uint64_t bits = 0;
    in.read(reinterpret_cast<char *>(&bits), 8);
    size_t nbytes = ((bits + 63) / 64) * 8;
    std::vector<char> buf(nbytes, 0);
    in.read(buf.data(), (std::streamsize)nbytes);
    v_.assign(bits / 16, 0);
    if (!v_.empty()) std::memcpy(v_.data(), buf.data(), v_.size() * 2);
  }

 private:
  std::vector<uint16_t> v_;
};

}  // namespace sdsl
