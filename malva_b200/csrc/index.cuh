// Device-resident index layout of one malva-geno run, and the per-thread accessors over it.
//
// Measured on B200 (profiles/round1_k1_v1.md): every random read that misses L2 moves a whole 128-byte
// line from HBM, whatever the size of the load.  The structures the sample scan probes for one k-mer --
// the alt-allele Bloom filter `bf` (bloom_filter.hpp) and the exact ref-allele map `ref_bf` (kmap.hpp) --
// are therefore interleaved into one array of 128-byte PROBE LINES, addressed by the SAME hash:
//
//   line L  (L = bf bit index >> 8)              128 bytes = 8 x uint4
//     u4[0..1]  bits 256L .. 256L+255 of bf      (bit i of bf = bit (i & 31) of u32 word ((i & 255) >> 5))
//     u4[2..7]  six key slots: canonical packed k-mers of the ref-allele signatures whose
//               XXH3 hash maps to a bf index inside this line; {lo, hi}, hi bits 62..63 are not key bits
//     bit 63 of slot 5's hi word = OVERFLOW flag: more than six keys hashed here, the rest live in the
//               open-addressing overflow table (ovf_keys / ovf_counts)
//
// so one k-mer costs one HBM line for both lookups.  bf keeps the reference's exact bit semantics
// (same hash, same `% size`, same bit); only where the bits are stored differs.
//
//   rank[L]        ones of bf before line L (u32)      -> bf_counts[rank + popc below]  (BF::_brank/_counts)
//   bf_counts[r]   u32 accumulator of the r-th set bit; read back & 0xFFFF (uint16 wrap-around)
//   key_counts[6L + s]  u32 count of key slot s of line L   (KMAP value, 32-bit wrap-around)
//   ctx_words      context_bf as a plain u32 bit array (consulted only on bf hits)
//   occ            occupancy pre-filter: one bit per 2^occ_shift bf indices (<= 64 MB, kept resident in L2);
//                  the scan and the reference pass consult it first and skip empty probe lines
#pragma once
#include <cstdint>

#include "xxh3.cuh"

namespace mg {

constexpr int LINE_U4 = 8;
constexpr int LINE_KEYS = 6;
constexpr uint64_t KEY_HI_MASK = 0x3FFFFFFFFFFFFFFFULL;  // k <= 63: a key uses at most 126 bits
constexpr uint32_t OVF_FLAG_W = 0x80000000u;             // in .w of u4[7]
constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ULL;

struct DevView {  // everything the kernels need, passed by value
  const uint4 *lines;
  uint64_t n_lines;
  const uint32_t *ctx_words;
  const uint32_t *rank;
  uint32_t *bf_counts;
  uint32_t *key_counts;
  const u128 *ovf_keys;
  uint32_t *ovf_counts;
  uint64_t ovf_mask;  // capacity-1
  int ovf_shift;      // 64 - log2(capacity)
  // after mg_finalize_alt the overflow keys are a SORTED array of ovf_n entries (binary search): the index image
  // -- key slots sorted within every line, the six smallest keys of a crowded line in the line, the rest here in
  // ascending order -- is then a function of the key SET alone, not of the order or the races of the inserts,
  // so that replicas built independently are identical and their counter arrays add up element by element
  uint64_t ovf_n;
  int ovf_sorted;
  uint64_t bf_bits;
  uint64_t bf_mask;  // bf_bits-1 when bf_bits is a power of two, else 0
  int k, ref_k;
  // occupancy pre-filter (L2-resident): bit (idx >> occ_shift) is set iff some bf bit or some ref key has its
  // bf index in [idx >> occ_shift << occ_shift, +2^occ_shift).  A clear bit means the probe line holds nothing
  // for this k-mer and the HBM access is skipped altogether.  nullptr = no pre-filter.
  const uint32_t *occ;
  int occ_shift;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint64_t bf_index(const DevView &v, uint64_t h) {
  return v.bf_mask ? (h & v.bf_mask) : (h % v.bf_bits);
}
__device__ __forceinline__ bool occ_test(const DevView &v, uint64_t idx) {
  if (!v.occ) return true;
  uint64_t o = idx >> v.occ_shift;
  return (__ldg(v.occ + (o >> 5)) >> (o & 31)) & 1u;
}
__device__ __forceinline__ void occ_set(const DevView &v, uint32_t *occ_rw, uint64_t idx) {
  if (!occ_rw) return;
  uint64_t o = idx >> v.occ_shift;
  uint32_t m = 1u << (o & 31);
  if (!(occ_rw[o >> 5] & m)) atomicOr(occ_rw + (o >> 5), m);
}
__device__ __forceinline__ bool ctx_test(const DevView &v, uint64_t idx) {
  return (__ldg(v.ctx_words + (idx >> 5)) >> (idx & 31)) & 1u;
}
__device__ __forceinline__ uint32_t line_word(const DevView &v, uint64_t line, int w) {  // w in 0..31
  return __ldg(reinterpret_cast<const uint32_t *>(v.lines) + line * 32 + (uint64_t)w);
}
__device__ __forceinline__ bool bf_test(const DevView &v, uint64_t idx) {
  return (line_word(v, idx >> 8, (int)((idx & 255) >> 5)) >> (idx & 31)) & 1u;
}
// rank of a set bit = ones strictly before idx (sdsl rank_support_v<1> semantics, bloom_filter.hpp:108)
__device__ __forceinline__ uint32_t bf_rank_of(const DevView &v, uint64_t idx) {
  uint64_t line = idx >> 8;
  int w = (int)((idx & 255) >> 5);
  uint32_t r = __ldg(v.rank + line);
  for (int x = 0; x < w; ++x) r += __popc(line_word(v, line, x));
  r += __popc(line_word(v, line, w) & ((1u << (idx & 31)) - 1u));
  return r;
}

__device__ __forceinline__ u128 key_of(uint4 q) {
  u128 r;
  r.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
  r.hi = ((uint64_t)q.z | ((uint64_t)q.w << 32)) & KEY_HI_MASK;
  return r;
}
__device__ __forceinline__ bool key_eq(u128 a, u128 b) { return a.lo == b.lo && a.hi == b.hi; }
__device__ __forceinline__ bool key_empty(u128 a) { return a.lo == ~0ull && a.hi == KEY_HI_MASK; }

// 128-bit compare-and-swap (PTX ISA 8.3+, sm_90+): returns the previous value
__device__ __forceinline__ u128 cas128(u128 *addr, u128 cmp, u128 val) {
  u128 old;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, s, r;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 s, {%4, %5};\n\t"
      "atom.global.cas.b128 r, [%6], c, s;\n\t"
      "mov.b128 {%0, %1}, r;\n\t"
      "}"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
      : "memory");
  return old;
}

__device__ __forceinline__ uint64_t ovf_slot0(const DevView &v, uint64_t h) { return (h * GOLD) >> v.ovf_shift; }

// slot of `canon` in the overflow structure, or -1
__device__ __forceinline__ int64_t ovf_find(const DevView &v, uint64_t h, u128 canon) {
  if (v.ovf_sorted) {
    uint64_t lo = 0, hi = v.ovf_n;  // first entry >= canon
    while (lo < hi) {
      uint64_t mid = (lo + hi) >> 1;
      u128 key = key_of(__ldg(reinterpret_cast<const uint4 *>(v.ovf_keys + mid)));
      if (less128(key, canon))
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo < v.ovf_n && key_eq(key_of(__ldg(reinterpret_cast<const uint4 *>(v.ovf_keys + lo))), canon)) return (int64_t)lo;
    return -1;
  }
  uint64_t slot = ovf_slot0(v, h);
  while (true) {
    u128 key = key_of(__ldg(reinterpret_cast<const uint4 *>(v.ovf_keys + slot)));
    if (key_eq(key, canon)) return (int64_t)slot;
    if (key_empty(key)) return -1;
    slot = (slot + 1) & v.ovf_mask;
  }
}

// Where the count of canonical key `canon` (hash h, bf index idx) lives:
//   >= 0            index into key_counts
//   <= -2           -(2 + slot) in ovf_counts
//   -1              key absent
__device__ __forceinline__ int64_t key_locate(const DevView &v, uint64_t h, uint64_t idx, u128 canon) {
  uint64_t line = idx >> 8;
  const uint4 *p = v.lines + line * LINE_U4 + 2;
  uint4 q[LINE_KEYS];
#pragma unroll
  for (int s = 0; s < LINE_KEYS; ++s) q[s] = __ldg(p + s);
#pragma unroll
  for (int s = 0; s < LINE_KEYS; ++s)
    if (key_eq(key_of(q[s]), canon)) return (int64_t)(line * LINE_KEYS + (uint64_t)s);
  if (!(q[LINE_KEYS - 1].w & OVF_FLAG_W)) return -1;
  int64_t slot = ovf_find(v, h, canon);
  return slot < 0 ? -1 : -(2 + slot);
}
__device__ __forceinline__ uint32_t *count_ptr(const DevView &v, int64_t loc) {
  return loc >= 0 ? v.key_counts + loc : v.ovf_counts + (uint64_t)(-loc - 2);
}

// insert into the six slots of a line; returns 1 = newly inserted, 0 = already present, -1 = line full
// (the caller then sets the overflow flag and spills the key to the overflow table)
__device__ __forceinline__ int line_insert(uint4 *lines_rw, uint64_t line, u128 canon) {
  u128 *slots = reinterpret_cast<u128 *>(lines_rw + line * LINE_U4 + 2);
  const u128 empty = {~0ull, KEY_HI_MASK};
  for (int s = 0; s < LINE_KEYS; ++s) {
    u128 old = cas128(slots + s, empty, canon);
    old.hi &= KEY_HI_MASK;
    if (key_empty(old)) return 1;
    if (key_eq(old, canon)) return 0;
  }
  return -1;
}
__device__ __forceinline__ void line_set_overflow(uint4 *lines_rw, uint64_t line) {
  uint32_t *w = reinterpret_cast<uint32_t *>(lines_rw + line * LINE_U4 + 7) + 3;
  atomicOr(w, OVF_FLAG_W);
}
// insert into the open-addressing overflow table; returns 1 = new, 0 = present
__device__ __forceinline__ int ovf_insert(const DevView &v, u128 *ovf_keys_rw, uint64_t h, u128 canon) {
  const u128 empty = {~0ull, KEY_HI_MASK};
  uint64_t slot = ovf_slot0(v, h);
  while (true) {
    u128 old = cas128(ovf_keys_rw + slot, empty, canon);
    old.hi &= KEY_HI_MASK;
    if (key_empty(old)) return 1;
    if (key_eq(old, canon)) return 0;
    slot = (slot + 1) & v.ovf_mask;
  }
}

#endif  // __CUDACC__

}  // namespace mg
