// K6: canonical k-mer counting on the device -- the step BEFORE the MALVA hot path: what the wrapper
// script obtains from `kmc -k43 -ci2 -cs255` (MALVA:107) and malva-geno then lists with the KMC API
// (main.cpp:482-490).  Semantics restated from what the reference's shipped golden pins: canonical
// k-mers (the lexicographic minimum of a k-mer and its reverse complement), windows that contain a non-ACGT symbol
// are skipped, k-mers seen fewer than `min_count` times are dropped, counts saturate at `counter_max`.
//
//   k_count_kmers   each CTA stages a tile of the read bytes (+ k-1 halo) in shared memory; each thread rolls
//                   CNT_RUN consecutive windows through two 128-bit registers (forward strand shifting left, reverse
//                   complement shifting right) and a run length of valid symbols, and inserts the canonical word of
//                   every valid window into an open-addressing table of 32-byte slots {key 16 B, count 4 B}: one
//                   16-byte read of the home slot, a 128-bit CAS only when the slot is empty, one atomicAdd.
//   k_count_rehash  table growth (load factor <= 0.5)
//   k_count_emit    kept entries (count >= min_count, saturated) appended to dense arrays; a radix sort by key then
//                   gives KMC's listing order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "index.cuh"
#include "xxh3.cuh"

namespace mg {

struct CountSlot {
  u128 key;  // canonical packed k-mer; {~0, ~0} = empty
  uint32_t count;
  uint32_t pad[3];
};
static_assert(sizeof(CountSlot) == 32, "one sector per slot");

constexpr int CNT_THREADS = 256;
constexpr int CNT_RUN = 16;
constexpr int CNT_TILE = CNT_THREADS * CNT_RUN;

__device__ __forceinline__ uint64_t count_hash(u128 x) {
  uint64_t h = (x.lo ^ (x.hi * 0x9E3779B97F4A7C15ULL)) * 0xBF58476D1CE4E5B9ULL;
  h ^= h >> 31;
  h *= 0x94D049BB133111EBULL;
  return h ^ (h >> 29);
}

__device__ __forceinline__ void count_insert(CountSlot *table, uint64_t mask, u128 key, uint32_t add,
                                             unsigned long long *n_distinct) {
  uint64_t slot = count_hash(key) & mask;
  const u128 empty = {~0ull, ~0ull};
  while (true) {
    u128 *kp = &table[slot].key;
    uint4 q = __ldcv(reinterpret_cast<const uint4 *>(kp));
    u128 cur = {(uint64_t)q.x | ((uint64_t)q.y << 32), (uint64_t)q.z | ((uint64_t)q.w << 32)};
    // (a half that reads as all-ones may be one side of a 128-bit store in flight: let the CAS return the truth)
    if (cur.lo == ~0ull || cur.hi == ~0ull) {
      cur = cas128(kp, empty, key);
      if (cur.lo == ~0ull && cur.hi == ~0ull) {
        atomicAdd(n_distinct, 1ull);
        cur = key;
      }
    }
    if (cur.lo == key.lo && cur.hi == key.hi) {
      atomicAdd(&table[slot].count, add);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// seq: read bytes, records separated by any non-ACGT byte (the host puts '\n' between reads).  Window end
// positions [k-1, len).  Only canonical k-mers whose top `part_bits` bits fall in [part_lo, part_hi) are counted
// (prefix-partitioned passes for inputs whose distinct k-mers do not fit the table at once; part_bits = 0: all).
template <int K>
__global__ void __launch_bounds__(CNT_THREADS) k_count_kmers(const uint8_t *__restrict__ seq, uint64_t len, int k_rt,
                                                            CountSlot *table, uint64_t mask, int part_bits,
                                                            uint32_t part_lo, uint32_t part_hi,
                                                            unsigned long long *n_distinct,
                                                            unsigned long long *n_instances) {
  extern __shared__ uint8_t cnt_sm[];
  const int k = K > 0 ? K : k_rt;
  uint64_t p0 = (uint64_t)(k - 1) + (uint64_t)blockIdx.x * CNT_TILE;
  uint64_t p1 = p0 + CNT_TILE < len ? p0 + CNT_TILE : len;
  uint64_t base = p0 - (uint64_t)(k - 1);
  int nbytes = (int)(p1 - base);
  for (int i = threadIdx.x; i < nbytes; i += CNT_THREADS) cnt_sm[i] = seq[base + i];
  __syncthreads();
  uint64_t q0 = p0 + (uint64_t)threadIdx.x * CNT_RUN;
  if (q0 >= p1) return;
  uint64_t q1 = q0 + CNT_RUN < p1 ? q0 + CNT_RUN : p1;
  u128 f = {0, 0}, r = {0, 0};
  int run = 0;  // valid symbols ending at the current position (saturating at k)
  const u128 m = mask128(u128{~0ull, ~0ull}, 2 * k);
  const int top = 2 * (k - 1);  // bit position of the first base
  auto push = [&](uint8_t ch) {
    uint32_t c = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
    if (c > 3u) {
      run = 0;
      f.lo = f.hi = r.lo = r.hi = 0;
      return;
    }
    f.hi = (f.hi << 2) | (f.lo >> 62);
    f.lo = (f.lo << 2) | c;
    f.hi &= m.hi;
    f.lo &= m.lo;
    r.lo = (r.lo >> 2) | (r.hi << 62);
    r.hi >>= 2;
    uint64_t cc = 3u - c;
    if (top >= 64)
      r.hi |= cc << (top - 64);
    else
      r.lo |= cc << (top & 63);
    if (run < k) ++run;
  };
  int o = (int)(q0 - base) - (k - 1);
  for (int j = 0; j < k - 1; ++j) push(cnt_sm[o + j]);
  uint32_t mine = 0;
  for (uint64_t p = q0; p < q1; ++p) {
    push(cnt_sm[(int)(p - base)]);
    if (run < k) continue;
    u128 canon = less128(r, f) ? r : f;  // (ties: the two strands are the same word)
    if (part_bits) {
      uint32_t part = (uint32_t)(shr128(canon, 2 * k - part_bits).lo);
      if (part < part_lo || part >= part_hi) continue;
    }
    count_insert(table, mask, canon, 1u, n_distinct);
    ++mine;
  }
  if (mine) atomicAdd(n_instances, (unsigned long long)mine);
}

__global__ void k_count_clear(CountSlot *table, uint64_t cap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  uint4 *p = reinterpret_cast<uint4 *>(table + i);
  p[0] = make_uint4(~0u, ~0u, ~0u, ~0u);
  p[1] = make_uint4(0, 0, 0, 0);
}

__global__ void k_count_rehash(const CountSlot *old_table, uint64_t old_cap, CountSlot *table, uint64_t mask,
                               unsigned long long *n_distinct) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  CountSlot s = old_table[i];
  if (s.key.lo == ~0ull && s.key.hi == ~0ull) return;
  count_insert(table, mask, s.key, s.count, n_distinct);
}

// pass 0 (keys == nullptr): count the kept entries; pass 1: append them
__global__ void k_count_emit(const CountSlot *table, uint64_t cap, uint32_t min_count, uint32_t counter_max,
                             uint64_t max_count, unsigned long long *n_out, u128 *keys, uint32_t *counts) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  CountSlot s;
  if (i < cap) {
    s = table[i];
    keep = !(s.key.lo == ~0ull && s.key.hi == ~0ull) && s.count >= min_count && (uint64_t)s.count <= max_count;
  }
  // warp-aggregated append
  unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (!ballot) return;
  int lane = threadIdx.x & 31, leader = __ffs(ballot) - 1;
  unsigned long long basepos = 0;
  if (lane == leader) basepos = atomicAdd(n_out, (unsigned long long)__popc(ballot));
  basepos = __shfl_sync(0xffffffffu, basepos, leader);
  if (keep && keys) {
    unsigned long long o = basepos + __popc(ballot & ((1u << lane) - 1u));
    keys[o] = s.key;
    counts[o] = s.count > counter_max ? counter_max : s.count;
  }
}

// 128-bit keys for cub::DeviceRadixSort: most significant word first
struct KmerDecomposer {
  __host__ __device__ ::cuda::std::tuple<uint64_t &, uint64_t &> operator()(u128 &key) const { return {key.hi, key.lo}; }
};

}  // namespace mg
