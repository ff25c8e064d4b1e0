// Hand-written sm_100a kernels of the MALVA hot path (see index.cuh for the data layout).
//   K3  k_add_signatures / k_add_packed / k_add_spill / k_line_popc    index-time inserts + switch_mode
//   K2  k_refpass / k_refpass_short                                    reference rolling pass
//   K1  k_scan                                                         sample k-mer scan
//   K4  k_mark_ref / k_lookup / k_coverage                             coverage read-back
//   K5  k_genotype                                                     likelihoods + posterior arg-max
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "geno.cuh"
#include "index.cuh"
#include "xxh3.cuh"

namespace mg {

template <int K>
__device__ __forceinline__ uint64_t canon_hash_k(u128 x, int k, u128 *canon) {
  if constexpr (K > 0) {
    return canon_hash<K>(x, canon);
  } else {
    return canon_hash_rt(x, k, canon);
  }
}

// scalars: [0] new keys, [1] irregular ref keys, [2] popcount, [3] error flag, [4] spilled keys
// ---------------------------------------------------------------------------
// K3a: index-time inserts (add_kmers_to_bf, main.cpp:122-144)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void insert_ref_key(const DevView &v, uint4 *lines_rw, uint32_t *occ_rw, uint64_t h,
                                               u128 canon, uint64_t i, unsigned long long *scalars,
                                               uint32_t *spill_idx) {
  occ_set(v, occ_rw, bf_index(v, h));
  uint64_t line = bf_index(v, h) >> 8;
  int r = line_insert(lines_rw, line, canon);
  if (r == 1) {
    atomicAdd(&scalars[0], 1ull);
  } else if (r < 0) {  // line full: flag it and leave the key for the overflow pass
    line_set_overflow(lines_rw, line);
    unsigned long long p = atomicAdd(&scalars[4], 1ull);
    spill_idx[p] = (uint32_t)i;
  }
}
__device__ __forceinline__ void set_bf_bit(const DevView &v, uint4 *lines_rw, uint32_t *occ_rw, uint64_t h) {
  uint64_t idx = bf_index(v, h);
  occ_set(v, occ_rw, idx);
  uint32_t *w = reinterpret_cast<uint32_t *>(lines_rw) + (idx >> 8) * 32 + ((idx & 255) >> 5);
  atomicOr(w, 1u << (idx & 31));
}

__global__ void __launch_bounds__(128) k_add_signatures(const uint8_t *__restrict__ pool,
                                                       const uint64_t *__restrict__ off,
                                                       const uint8_t *__restrict__ is_ref, uint64_t n, DevView v,
                                                       uint4 *lines_rw, uint32_t *occ_rw, unsigned long long *scalars,
                                                       uint32_t *irregular_idx, uint32_t *spill_idx) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t b = off[i], e = off[i + 1];
  int len = (int)(e - b);
  if (len > 128) {
    atomicExch(&scalars[3], 1ull);
    return;
  }
  uint8_t s[128];
  for (int j = 0; j < len; ++j) s[j] = pool[b + j];
  u128 x, canon;
  bool regular = pack_ascii(s, len, v.k, &x);
  if (is_ref[i]) {  // ref_bf.add_key
    if (!regular) {  // not k symbols of ACGT: can never match a sample k-mer; kept on the host
      unsigned long long p = atomicAdd(&scalars[1], 1ull);
      irregular_idx[p] = (uint32_t)i;
      return;
    }
    uint64_t h = canon_hash_rt(x, v.k, &canon);
    insert_ref_key(v, lines_rw, occ_rw, h, canon, i, scalars, spill_idx);
  } else {  // bf.add_key
    uint64_t h = regular ? canon_hash_rt(x, v.k, &canon) : hash_ascii(s, len);
    set_bf_bit(v, lines_rw, occ_rw, h);
  }
}

// same inserts for signature k-mers that arrive already packed (exactly k symbols of ACGT)
__global__ void __launch_bounds__(256) k_add_packed(const uint4 *__restrict__ kmers, const uint8_t *__restrict__ is_ref,
                                                   uint64_t n, DevView v, uint4 *lines_rw, uint32_t *occ_rw,
                                                   unsigned long long *scalars, uint32_t *spill_idx) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4 q = kmers[i];
  u128 x, canon;
  x.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
  x.hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
  x = mask128(x, 2 * v.k);
  uint64_t h = canon_hash_rt(x, v.k, &canon);
  if (is_ref[i])
    insert_ref_key(v, lines_rw, occ_rw, h, canon, i, scalars, spill_idx);
  else
    set_bf_bit(v, lines_rw, occ_rw, h);
}

// second pass over the keys whose line was full: insert into the overflow table.
// Keys come either from an ASCII pool (off != nullptr) or from a packed array.
__global__ void __launch_bounds__(128) k_add_spill(const uint32_t *__restrict__ spill_idx, uint64_t n_spill,
                                                  const uint8_t *__restrict__ pool, const uint64_t *__restrict__ off,
                                                  const uint4 *__restrict__ packed, DevView v, u128 *ovf_keys_rw,
                                                  unsigned long long *scalars) {
  uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_spill) return;
  uint64_t i = spill_idx[j];
  u128 x, canon;
  if (off) {
    uint64_t b = off[i];
    int len = (int)(off[i + 1] - b);
    uint8_t s[64];
    for (int t = 0; t < len && t < 64; ++t) s[t] = pool[b + t];
    pack_ascii(s, len, v.k, &x);
  } else {
    uint4 q = packed[i];
    x.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
    x.hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
    x = mask128(x, 2 * v.k);
  }
  uint64_t h = canon_hash_rt(x, v.k, &canon);
  if (ovf_insert(v, ovf_keys_rw, h, canon) == 1) atomicAdd(&scalars[0], 1ull);
}

__global__ void k_fill_keys(u128 *keys, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i].lo = ~0ull;
    keys[i].hi = KEY_HI_MASK;
  }
}
// a fresh probe-line array: filter bits 0, key slots empty
__global__ void k_init_lines(uint4 *lines, uint64_t n_lines) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one uint4 per thread
  if (i >= n_lines * LINE_U4) return;
  uint32_t hi = (uint32_t)(KEY_HI_MASK >> 32);
  lines[i] = (i & 7) < 2 ? make_uint4(0, 0, 0, 0) : make_uint4(~0u, ~0u, ~0u, hi);
}

// re-insert every key of an old overflow table into a larger one (counts carried over)
__global__ void k_rehash(const u128 *old_keys, const uint32_t *old_counts, uint64_t old_cap, DevView v,
                         u128 *ovf_keys_rw) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  u128 key = old_keys[i];
  key.hi &= KEY_HI_MASK;
  if (key_empty(key)) return;
  u128 canon;
  uint64_t h = canon_hash_rt(key, v.k, &canon);  // keys are canonical: canon == key
  const u128 empty = {~0ull, KEY_HI_MASK};
  uint64_t slot = ovf_slot0(v, h);
  while (true) {
    u128 old = cas128(ovf_keys_rw + slot, empty, key);
    old.hi &= KEY_HI_MASK;
    if (key_empty(old)) break;
    slot = (slot + 1) & v.ovf_mask;
  }
  v.ovf_counts[slot] = old_counts[i];
}

// ---------------------------------------------------------------------------
// K3c: canonical index image (run by mg_finalize_alt, before any count exists): the key slots of every line in
// ascending order, empty slots last.  Which slot a key took during the build depended on the order (and the races)
// of the inserts; after this pass it depends on the key set only.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sort_line_keys(uint4 *lines, uint64_t n_lines) {
  uint64_t line = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  uint4 *p = lines + line * LINE_U4 + 2;
  uint4 q[LINE_KEYS];
#pragma unroll
  for (int s = 0; s < LINE_KEYS; ++s) q[s] = p[s];
  if (key_empty(key_of(q[1]))) return;  // zero or one key: already canonical (slots fill from the front)
  const uint32_t flag = q[LINE_KEYS - 1].w & OVF_FLAG_W;
  q[LINE_KEYS - 1].w &= ~OVF_FLAG_W;
#pragma unroll
  for (int a = 1; a < LINE_KEYS; ++a) {  // insertion sort on (hi, lo); an empty slot is the largest value
#pragma unroll
    for (int b = a; b > 0; --b) {
      u128 x = key_of(q[b - 1]), y = key_of(q[b]);
      if (less128(y, x)) {
        uint4 t = q[b - 1];
        q[b - 1] = q[b];
        q[b] = t;
      }
    }
  }
  q[LINE_KEYS - 1].w |= flag;
#pragma unroll
  for (int s = 0; s < LINE_KEYS; ++s) p[s] = q[s];
}
// the six key slots of the listed lines, to / from a dense array (overflow canonicalisation on the host)
__global__ void k_gather_line_keys(const uint4 *lines, const uint64_t *ids, uint64_t n, uint4 *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * LINE_KEYS) return;
  out[i] = lines[ids[i / LINE_KEYS] * LINE_U4 + 2 + (i % LINE_KEYS)];
}
__global__ void k_scatter_line_keys(uint4 *lines, const uint64_t *ids, uint64_t n, const uint4 *in) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * LINE_KEYS) return;
  lines[ids[i / LINE_KEYS] * LINE_U4 + 2 + (i % LINE_KEYS)] = in[i];
}

// ---------------------------------------------------------------------------
// K3b: switch_mode (bloom_filter.hpp:93-98): ones per probe line (then an exclusive scan -> rank)
// also used on a plain bit array (stride_u32 = 8) for context_bf statistics
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_line_popc(const uint32_t *__restrict__ words, uint64_t n_units, int stride_u32,
                                                  uint32_t *__restrict__ unit_count, unsigned long long *total) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  if (b < n_units) {
    const uint4 *p = reinterpret_cast<const uint4 *>(words + b * (uint64_t)stride_u32);
    uint4 q0 = p[0], q1 = p[1];
    c = __popc(q0.x) + __popc(q0.y) + __popc(q0.z) + __popc(q0.w) + __popc(q1.x) + __popc(q1.y) + __popc(q1.z) +
        __popc(q1.w);
    if (unit_count) unit_count[b] = c;
  }
  __shared__ uint32_t red[8];
  uint32_t s = c;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffu, s, o);
    if (threadIdx.x == 0 && s) atomicAdd(total, (unsigned long long)s);
  }
}
// the 256 filter bits of every probe line, gathered into a plain bit array (state download)
__global__ void k_extract_bits(const uint4 *__restrict__ lines, uint64_t n_lines, uint4 *__restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one uint4 (128 bits) per thread
  if (i >= n_lines * 2) return;
  out[i] = lines[(i >> 1) * LINE_U4 + (i & 1)];
}

// ---------------------------------------------------------------------------
// K2: reference rolling pass (main.cpp:385-400)
// Each CTA stages a tile of the contig in shared memory (with a ref_k-1 halo); each thread rolls RP_RUN
// consecutive windows through 2-bit registers.  Windows that contain a non-ACGT symbol take the
// byte-exact ASCII path (the RCN table maps IUPAC symbols to NUL, bloom_filter.hpp:36-50).
// ---------------------------------------------------------------------------
constexpr int RP_THREADS = 256;
constexpr int RP_RUN = 16;
constexpr int RP_TILE = RP_THREADS * RP_RUN;

__device__ __forceinline__ uint32_t base_code(uint8_t c) {  // 0..3, or 4 for anything else
  return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

template <int K, int REFK>
__global__ void __launch_bounds__(RP_THREADS) k_refpass(const uint8_t *__restrict__ seq, uint64_t len, DevView v,
                                                        uint32_t *ctx_words_rw) {
  extern __shared__ uint8_t sm[];
  const int k = K > 0 ? K : v.k, ref_k = REFK > 0 ? REFK : v.ref_k;
  const int d = (ref_k - k) / 2;
  const bool odd = ((ref_k - k) & 1) != 0;
  // window end positions handled by this CTA: [p0, p1)
  uint64_t p0 = (uint64_t)(ref_k - 1) + (uint64_t)blockIdx.x * RP_TILE;
  uint64_t p1 = p0 + RP_TILE < len ? p0 + RP_TILE : len;
  uint64_t base = p0 - (uint64_t)(ref_k - 1);  // first byte staged
  int nbytes = (int)(p1 - base);
  for (int i = threadIdx.x; i < nbytes; i += RP_THREADS) sm[i] = seq[base + i];
  __syncthreads();
  uint64_t q0 = p0 + (uint64_t)threadIdx.x * RP_RUN;
  if (q0 >= p1) return;
  uint64_t q1 = q0 + RP_RUN < p1 ? q0 + RP_RUN : p1;
  // prime the rolling state with the ref_k-1 bases before q0
  u128 x = {0, 0};
  uint64_t bad = 0;  // bit j set <=> base (p - j) is not ACGT
  const u128 m = mask128(u128{~0ull, ~0ull}, 2 * ref_k);
  int o = (int)(q0 - base) - (ref_k - 1);
  for (int j = 0; j < ref_k - 1; ++j) {
    uint32_t c = base_code(sm[o + j]);
    x.hi = (x.hi << 2) | (x.lo >> 62);
    x.lo = (x.lo << 2) | (c & 3u);
    bad = (bad << 1) | (c >> 2);
  }
  const uint64_t m43 = ref_k >= 64 ? ~0ull : ((1ull << ref_k) - 1);
  const uint64_t mk = k >= 64 ? ~0ull : ((1ull << k) - 1);
  for (uint64_t p = q0; p < q1; ++p) {
    int sp = (int)(p - base);
    uint32_t c = base_code(sm[sp]);
    x.hi = (x.hi << 2) | (x.lo >> 62);
    x.lo = (x.lo << 2) | (c & 3u);
    x.hi &= m.hi;
    x.lo &= m.lo;
    bad = (bad << 1) | (c >> 2);
    // k-mer window of the reference at this step.  With t = p-(ref_k-1) slides done:
    //   (ref_k-k) even        : ref[p-d-k+1 .. p-d]
    //   odd, t == 0 (primed)  : ref[d .. d+k-1]                 (ends at p-d-1)
    //   odd, 1 <= t < k       : ref[d+t .. d+k-1] ++ ref[k+d+1 .. k+d+t]   (main.cpp:395-397 skips ref[d+k])
    //   odd, t >= k           : ref[p-d-k+1 .. p-d]
    uint64_t t = p - (uint64_t)(ref_k - 1);
    bool quirk = odd && t >= 1 && t < (uint64_t)k;
    int shift = d + ((odd && t == 0) ? 1 : 0);
    uint64_t h35;
    if (!quirk && ((bad >> shift) & mk) == 0) {
      u128 x35 = mask128(shr128(x, 2 * shift), 2 * k), canon;
      h35 = canon_hash_k<K>(x35, k, &canon);
    } else {
      uint8_t s[64];
      if (!quirk) {
        for (int j = 0; j < k; ++j) s[j] = sm[sp - shift - k + 1 + j];
      } else {
        int n_old = k - (int)t;
        for (int j = 0; j < n_old; ++j) s[j] = seq[(uint64_t)d + t + (uint64_t)j];
        for (int j = 0; j < (int)t; ++j) s[n_old + j] = seq[(uint64_t)(k + d + 1) + (uint64_t)j];
      }
      h35 = hash_ascii(s, k);
    }
    const uint64_t i35 = bf_index(v, h35);
    if (!occ_test(v, i35) || !bf_test(v, i35)) continue;
    uint64_t h43;
    if ((bad & m43) == 0) {
      u128 canon;
      h43 = canon_hash_k<REFK>(x, ref_k, &canon);
    } else {
      uint8_t s[64];
      for (int j = 0; j < ref_k; ++j) s[j] = sm[sp - ref_k + 1 + j];
      h43 = hash_ascii(s, ref_k);
    }
    uint64_t cidx = bf_index(v, h43);
    atomicOr(ctx_words_rw + (cidx >> 5), 1u << (cidx & 31));
  }
}

// contig shorter than ref_k: the reference hashes the (shorter) substr() results once
__global__ void k_refpass_short(const uint8_t *seq, uint64_t len, DevView v, uint32_t *ctx_words_rw) {
  if (threadIdx.x || blockIdx.x) return;
  int d = (v.ref_k - v.k) / 2;
  int kl = (int)len - d < v.k ? (int)len - d : v.k;
  uint64_t h = hash_ascii(seq + d, kl);
  if (!bf_test(v, bf_index(v, h))) return;
  uint64_t hc = hash_ascii(seq, (int)len);
  uint64_t cidx = bf_index(v, hc);
  atomicOr(ctx_words_rw + (cidx >> 5), 1u << (cidx & 31));
}

// ---------------------------------------------------------------------------
// K1: sample k-mer scan (main.cpp:487-500)
//   ref_bf.increment(kmer, c);  if (!context_bf.test_key(context)) bf.increment(kmer, c);
// A warp owns 32 k-mers (lane i hashes k-mer i: one coalesced 16-byte load, canonical form, XXH3).
// The probe lines of the 32 k-mers are then copied into the warp's 4 KB shared-memory tile with
// cp.async (LDGSTS, no register staging): in round r the four 8-lane groups of the warp copy the lines
// of k-mers 4r..4r+3, lane j of a group moving uint4 j, so that one k-mer costs exactly one fully
// coalesced 128-byte request, and all 32 lines are in flight together.  The tile is XOR-swizzled
// (uint4 j of line L sits at column j ^ (L & 7)) so that every lane can then read ITS OWN line with
// conflict-free 16-byte shared loads and do the filter-bit test and the six key compares locally.
// The context filter, the rank directory and the counters are touched only on the ~1-4 % hit paths.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// canon_hash<K> (xxh3.cuh) with the 2-bit -> ASCII expansion done by a 256-entry shared-memory table (4 bases per
// look-up) instead of shift/mask/PRMT sequences: the scan kernel is bound by the ALU pipe (LOP3/SHF/PRMT issue at
// half rate), shared-memory loads go through the otherwise idle LSU pipe.  tab[v] = expand4(v).
template <int K>
__device__ __forceinline__ uint64_t canon_hash_lut(u128 x, u128 *canon, const uint32_t *tab) {
  u128 rc = revcomp(x, K);
  bool fwd = less128(x, rc);
  u128 c = fwd ? x : rc, other = fwd ? rc : x;
  u128 r;
  r.lo = ~other.lo;
  r.hi = ~other.hi;
  r = mask128(r, 2 * K);  // LSB-first image of the canonical k-mer
  constexpr int NW = (K + 7) / 8;
  uint64_t w[NW + 1];
  const uint32_t rw[4] = {(uint32_t)r.lo, (uint32_t)(r.lo >> 32), (uint32_t)r.hi, (uint32_t)(r.hi >> 32)};
#pragma unroll
  for (int j = 0; j < NW; ++j) {  // ASCII bytes 8j .. 8j+7 = bases held by bytes 2j and 2j+1 of r
    const uint32_t word = rw[j >> 1];
    uint32_t lo = tab[__byte_perm(word, 0u, (j & 1) ? 0x4442u : 0x4440u)], hi = 0;
    if (8 * j + 4 < K) hi = tab[__byte_perm(word, 0u, (j & 1) ? 0x4443u : 0x4441u)];
    w[j] = (uint64_t)lo | ((uint64_t)hi << 32);
  }
  w[NW] = 0;
  *canon = c;
  return xxh3_64_words(w, K);
}

constexpr int SCAN_THREADS = 256;
// per warp: a 4 KB tile + 32 row indices; per CTA: the 1 KB expansion table
constexpr int SCAN_SMEM = (SCAN_THREADS / 32) * (32 * 128 + 32 * 4 + 4) + 256 * 4;  // (+ a hit counter per warp)

// Where the sample k-mers come from.  MODE 0: packed {lo,hi} words + u32 counts.  MODE 1: raw records of a
// KMC database suffix file (.kmc_suf): (ref_k - p)/4 suffix bytes (2-bit codes, first base most significant)
// + counter_size little-endian count bytes; the p-symbol prefix of record g is the LUT bucket that contains g
// (lut[j] = records before prefix j; listing order = record order).  Decoding them here halves the PCIe
// bytes per k-mer (10 B instead of 20 B for k = 43) and removes the host-side CKmerAPI::to_string pass.
struct ScanSrc {
  const uint4 *kmers;
  const uint32_t *counts;
  const uint8_t *recs;  // 16-byte aligned, padded by 16 bytes
  const uint64_t *lut;  // n_lut entries + a guard of ~0
  uint64_t first_rec;   // global index of recs[0]
  uint32_t n_lut, prefix_mask;
  int prefix_len, suf_bytes, counter_size;
  uint32_t min_count;
  uint64_t max_count;
  // Deferred filter hits.  A k-mer whose bf bit is set needs a second XXH3 (the 43-mer, for the context filter): rare
  // per k-mer (~0.7 %) but not per warp (one iteration in five), and while one lane runs it 31 idle.  The scan
  // therefore only records the hit -- {context k-mer, bf index, count}, 32 bytes, in the warp's own segment of
  // hit_buf -- and k_scan_hits works all of them off afterwards with full warps.  A full segment falls back to the
  // in-line path.  hit_buf == nullptr: always in line.
  uint4 *hit_buf;
  uint32_t *hit_counts;  // per warp of the scan grid
  uint32_t seg_cap;      // entries per warp segment
};

__device__ __forceinline__ uint32_t lut_bucket(const uint64_t *lut, uint32_t n_lut, uint64_t g) {
  uint32_t lo = 0, hi = n_lut;  // largest j with lut[j] <= g (lut[0] == 0, lut[n_lut] == ~0)
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(lut + mid) <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

template <int K, int REFK, int MODE>
__global__ void __launch_bounds__(SCAN_THREADS, 5) k_scan(ScanSrc src, uint64_t n, DevView v) {
  extern __shared__ uint4 scan_sm[];
  const int k = K > 0 ? K : v.k, ref_k = REFK > 0 ? REFK : v.ref_k;
  const int tail = ref_k - k - (ref_k - k) / 2;  // bases of the context after the k-mer (main.cpp:493)
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  uint4 *tile = scan_sm + (threadIdx.x >> 5) * 256;  // 32 lines x 8 uint4
  const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(tile);
  uint32_t *rows = reinterpret_cast<uint32_t *>(scan_sm + (SCAN_THREADS / 32) * 256) + (threadIdx.x >> 5) * 32;
  uint32_t *tab = reinterpret_cast<uint32_t *>(scan_sm + (SCAN_THREADS / 32) * 256) + (SCAN_THREADS / 32) * 32;
  uint32_t *hitc = tab + 256 + (threadIdx.x >> 5);  // this warp's deferred-hit counter
  if (lane == 0) *hitc = 0;
  if constexpr (K > 0) {
    static_assert(SCAN_THREADS == 256, "one table entry per thread");
    tab[threadIdx.x] = expand4(threadIdx.x);
    __syncthreads();
  } else {
    __syncwarp();
  }
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // 32-bit indices: the host never launches more than 2^31 k-mers at once
  const uint32_t n32 = (uint32_t)n;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t step = ((gridDim.x * blockDim.x) >> 5) * 32;
  for (uint32_t base = warp * 32; base < n32; base += step) {
    const uint32_t i = base + lane;
    bool live = i < n32;
    uint32_t cnt;
    u128 x43, canon;
    if constexpr (MODE == 0) {
      uint4 q = live ? __ldg(src.kmers + i) : make_uint4(0, 0, 0, 0);
      cnt = live ? __ldg(src.counts + i) : 0u;
      x43.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
      x43.hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
    } else {
      // stage the warp's 32 records (contiguous bytes) in its shared-memory tile, then decode one per lane
      const int rec = src.suf_bytes + src.counter_size;
      const uint64_t byte0 = (uint64_t)base * (uint64_t)rec, start = byte0 & ~3ull;
      const int n_words = (int)((byte0 - start) + 32u * (uint32_t)rec + 3u) >> 2;
      __syncwarp();
      uint32_t *stage = reinterpret_cast<uint32_t *>(tile);
      for (int w = lane; w < n_words; w += 32) stage[w] = __ldg(reinterpret_cast<const uint32_t *>(src.recs + start) + w);
      __syncwarp();
      const uint8_t *rb = reinterpret_cast<const uint8_t *>(stage) + (byte0 - start) + (uint32_t)lane * (uint32_t)rec;
      u128 suf = {0, 0};
      for (int j = 0; j < src.suf_bytes; ++j) {
        suf.hi = (suf.hi << 8) | (suf.lo >> 56);
        suf.lo = (suf.lo << 8) | rb[j];
      }
      uint64_t c64 = src.counter_size ? 0 : 1;
      for (int j = 0; j < src.counter_size; ++j) c64 |= (uint64_t)rb[src.suf_bytes + j] << (8 * j);
      __syncwarp();  // the tile is reused for the probe lines below
      // prefix of each record: one LUT search per warp in the common case (a prefix bucket spans many records)
      const uint64_t g = src.first_rec + i;
      const uint64_t g_first = src.first_rec + base;
      const uint64_t g_last = src.first_rec + (base + 31 < n32 ? base + 31 : n32 - 1);
      uint32_t pj = lut_bucket(src.lut, src.n_lut, g_first);
      if (__ldg(src.lut + pj + 1) <= g_last) pj = lut_bucket(src.lut, src.n_lut, live ? g : g_first);
      u128 pre = {(uint64_t)(pj & src.prefix_mask), 0};
      const int sh = 8 * src.suf_bytes;  // the suffix holds 4 * suf_bytes symbols
      if (sh >= 64) {
        pre.hi = pre.lo << (sh - 64);
        pre.lo = 0;
      } else {
        pre.hi = sh ? (pre.lo >> (64 - sh)) : 0;
        pre.lo <<= sh;
      }
      x43.lo = pre.lo | suf.lo;
      x43.hi = pre.hi | suf.hi;
      // CKMCFile::ReadNextKmer skips records whose count is outside [min_count, max_count]
      if (c64 < src.min_count || c64 > src.max_count) live = false;
      cnt = (uint32_t)c64;
    }
    u128 x35 = mask128(shr128(x43, 2 * tail), 2 * k);
    uint64_t h;
    if constexpr (K > 0)
      h = canon_hash_lut<K>(x35, &canon, tab);
    else
      h = canon_hash_rt(x35, k, &canon);
    uint64_t idx = bf_index(v, h);
    uint32_t line = (uint32_t)(idx >> 8);  // n_lines < 2^32 (bf_bits < 2^40)
    uint32_t bit = (uint32_t)(idx & 255);
    // occupancy pre-filter (L2): most probe lines hold nothing for a given k-mer; those are never fetched
    const bool need = live && occ_test(v, idx);
    const uint32_t need_mask = __ballot_sync(0xffffffffu, need);
    const int n_need = __popc(need_mask);
    const int rank = __popc(need_mask & ((1u << lane) - 1u));  // row of this lane's line in the tile
    __syncwarp();  // the previous iteration's reads of the tile and of the row list are done
    if (need) rows[rank] = line;
    __syncwarp();
    {
      // row L = grp, grp + 4, ...: uint4 `sub` of line rows[L] goes to column sub ^ (L & 7) of tile row L.  L & 7
      // alternates between grp and grp + 4, i.e. the column toggles bit 2 from one round to the next.
      uint32_t dst = tile_addr + (uint32_t)(grp * 128 + ((sub ^ grp) * 16));
      int toggle = ((sub ^ grp) & 4) ? -64 : 64;
      const uint32_t *rp = rows + grp;
      const char *src0 = reinterpret_cast<const char *>(v.lines) + sub * 16;
#pragma unroll 1
      for (int L = grp; L < n_need; L += 4) {
        cp_async16(dst, src0 + (uint64_t)(*rp) * 128);
        rp += 4;
        dst += 512 + toggle;
        toggle = -toggle;
      }
    }
    if constexpr (MODE == 0) {
      // while the lines are in flight: the warp's next batch (512 B of k-mers + 128 B of counts) into L2
      if (lane < 5 && base + step < n32) {
        const char *pf = lane < 4 ? reinterpret_cast<const char *>(src.kmers + base + step) + lane * 128
                                  : reinterpret_cast<const char *>(src.counts + base + step);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      }
    }
    cp_async_wait_all();
    __syncwarp();
    if (!need) continue;
    // ---- every needing lane now owns row `rank` of the tile ----
    const uint4 *mine = tile + rank * 8;
    const int sw = rank & 7;
    uint32_t wsel = bit >> 5;  // which of the 8 filter words
    uint32_t fw = reinterpret_cast<const uint32_t *>(mine + ((wsel >> 2) ^ sw))[wsel & 3];
    bool bf_hit = (fw >> (bit & 31u)) & 1u;
    const uint32_t c0 = (uint32_t)canon.lo, c1 = (uint32_t)(canon.lo >> 32), c2 = (uint32_t)canon.hi,
                   c3 = (uint32_t)(canon.hi >> 32);
    // the low word of each key slot first (one 4-byte shared load per slot); the other three words only of the
    // slots where it matches (a ref-key hit, ~3 % of the k-mers, or a 2^-32 coincidence)
    uint32_t low_match = 0;  // bit s: the low word of slot s equals the k-mer's
#pragma unroll
    for (int s = 0; s < LINE_KEYS; ++s)
      low_match |= (uint32_t)(reinterpret_cast<const uint32_t *>(mine + ((2 + s) ^ sw))[0] == c0) << s;
    int slot = -1;
    while (low_match) {  // (some lane of the warp gets here in most iterations: keep it to the matching slot)
      const int s = __ffs(low_match) - 1;
      low_match &= low_match - 1;
      uint4 p = mine[(2 + s) ^ sw];
      if (((p.y ^ c1) | (p.z ^ c2) | ((p.w & (uint32_t)(KEY_HI_MASK >> 32)) ^ c3)) == 0) slot = s;
    }
    const uint32_t last_w = reinterpret_cast<const uint32_t *>(mine + ((2 + LINE_KEYS - 1) ^ sw))[3];
    // ---- ref_bf.increment ----
    if (slot >= 0) {
      atomicAdd(v.key_counts + (uint64_t)line * LINE_KEYS + (uint32_t)slot, cnt);
    } else if (last_w & OVF_FLAG_W) {  // line overflowed at index time: the key may live in the overflow array
      const int64_t os = ovf_find(v, h, canon);
      if (os >= 0) atomicAdd(v.ovf_counts + os, cnt);
    }
    // ---- bf.increment unless the context filter vetoes it ----
    if (bf_hit) {
      const uint32_t pos = src.hit_buf ? atomicAdd(hitc, 1u) : 0xFFFFFFFFu;
      if (pos < src.seg_cap) {  // recorded; k_scan_hits finishes it
        uint4 *e = src.hit_buf + ((uint64_t)warp_id * src.seg_cap + pos) * 2;
        e[0] = make_uint4((uint32_t)x43.lo, (uint32_t)(x43.lo >> 32), (uint32_t)x43.hi, (uint32_t)(x43.hi >> 32));
        e[1] = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), cnt, 0u);
      } else {
        u128 c43;
        uint64_t h43;
        if constexpr (REFK > 0)
          h43 = canon_hash_lut<REFK>(x43, &c43, tab);
        else
          h43 = canon_hash_rt(x43, ref_k, &c43);
        if (!ctx_test(v, bf_index(v, h43))) atomicAdd(v.bf_counts + bf_rank_of(v, idx), cnt);
      }
    }
  }
  if (src.hit_buf) {
    __syncwarp();
    if (lane == 0) src.hit_counts[warp_id] = *hitc < src.seg_cap ? *hitc : src.seg_cap;
  }
}

// second half of the scan for the recorded filter hits: 43-mer hash -> context filter -> rank -> counter
template <int REFK>
__global__ void __launch_bounds__(256) k_scan_hits(const uint4 *__restrict__ hit_buf, const uint32_t *__restrict__ hit_counts,
                                                  uint32_t n_warps, uint32_t seg_cap, DevView v) {
  __shared__ uint32_t tab[256];
  tab[threadIdx.x] = expand4(threadIdx.x);
  __syncthreads();
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t w = (uint32_t)(t / seg_cap), slot = (uint32_t)(t % seg_cap);
  if (w >= n_warps || slot >= __ldg(hit_counts + w)) return;
  const uint4 a = __ldg(hit_buf + t * 2), b = __ldg(hit_buf + t * 2 + 1);
  u128 x43 = {(uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)a.z | ((uint64_t)a.w << 32)}, c43;
  const uint64_t idx = (uint64_t)b.x | ((uint64_t)b.y << 32);
  const uint32_t r = bf_rank_of(v, idx);  // (independent of the hash: these loads overlap with it)
  uint64_t h43;
  if constexpr (REFK > 0)
    h43 = canon_hash_lut<REFK>(x43, &c43, tab);
  else
    h43 = canon_hash_rt(x43, v.ref_k, &c43);
  if (!ctx_test(v, bf_index(v, h43))) atomicAdd(v.bf_counts + r, b.z);
}

// ---------------------------------------------------------------------------
// K4: signature look-ups (BF::get_count / KMAP::get_count) + coverage
// ---------------------------------------------------------------------------
// flags the k-mers of allele slot 0 of every variant (they are looked up in ref_bf, main.cpp:167-170)
__global__ void __launch_bounds__(256) k_mark_ref(const uint64_t *__restrict__ var_allele_off,
                                                 const uint64_t *__restrict__ allele_sig_off,
                                                 const uint64_t *__restrict__ sig_kmer_off, uint64_t n_variants,
                                                 uint8_t *__restrict__ flags) {
  uint64_t vi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= n_variants) return;
  uint64_t a0 = var_allele_off[vi];
  if (var_allele_off[vi + 1] == a0) return;
  for (uint64_t s = allele_sig_off[a0]; s < allele_sig_off[a0 + 1]; ++s)
    for (uint64_t q = sig_kmer_off[s]; q < sig_kmer_off[s + 1]; ++q) flags[q] = 1;
}

// mode 0: get_count  (is_ref selects KMAP/BF, out = int32 count)
// mode 1: test_key on filter/table `which` (0 bf, 1 context_bf, 2 ref_bf; out = 0/1, -1 = irregular KMAP key)
__global__ void __launch_bounds__(128) k_lookup(const uint8_t *__restrict__ pool, const uint64_t *__restrict__ off,
                                               const uint8_t *__restrict__ is_ref, uint64_t n, DevView v, int mode,
                                               int which, int32_t *__restrict__ out, unsigned long long *scalars,
                                               const uint8_t *__restrict__ only_flagged) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (only_flagged && !only_flagged[i]) return;  // second pass after k_lookup_fast: the k-mers it deferred
  uint64_t b = off[i], e = off[i + 1];
  int len = (int)(e - b);
  if (len > 128) {
    atomicExch(&scalars[3], 1ull);
    out[i] = 0;
    return;
  }
  uint8_t s[128];
  for (int j = 0; j < len; ++j) s[j] = pool[b + j];
  bool use_table = mode == 0 ? (is_ref[i] != 0) : (which == 2);
  u128 x, canon;
  if (use_table) {
    if (!pack_ascii(s, len, v.k, &x)) {
      out[i] = (mode == 1) ? -1 : 0;  // irregular keys are resolved on the host (always count 0)
      return;
    }
    uint64_t h = canon_hash_rt(x, v.k, &canon);
    int64_t loc = key_locate(v, h, bf_index(v, h), canon);
    if (mode == 1)
      out[i] = loc != -1;
    else
      out[i] = loc != -1 ? (int32_t)*count_ptr(v, loc) : 0;
    return;
  }
  // a Bloom filter: hash the canonical ASCII bytes of whatever length was given
  bool regular = len >= 1 && len <= 64 && pack_ascii(s, len, len, &x);
  uint64_t h = regular ? canon_hash_rt(x, len, &canon) : hash_ascii(s, len);
  uint64_t idx = bf_index(v, h);
  bool set = (mode == 1 && which == 1) ? ctx_test(v, idx) : bf_test(v, idx);
  if (mode == 1) {
    out[i] = set;
  } else {
    out[i] = (set && v.rank) ? (int32_t)(v.bf_counts[bf_rank_of(v, idx)] & 0xFFFFu) : 0;
  }
}

// Fast path of mode 0 for the compiled k: signature k-mers that are exactly K bytes long are read as aligned
// 32-bit words, packed to 2-bit codes four bases at a time and validated by re-expanding the codes to ASCII
// (pack_words, xxh3.cuh); anything that is not K symbols of ACGT, and the few k-mers at the very end of the
// pool, take the generic byte path above.  ~10x fewer instructions per k-mer than the generic kernel.
template <int K>
__global__ void __launch_bounds__(128) k_lookup_fast(const uint8_t *__restrict__ pool, uint64_t pool_bytes,
                                                    const uint64_t *__restrict__ off,
                                                    const uint8_t *__restrict__ is_ref, uint64_t n, DevView v,
                                                    int32_t *__restrict__ out, uint8_t *__restrict__ slow_flag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int NW = (K + 3) / 4;  // 4-base groups
  const uint64_t b = off[i], e = off[i + 1];
  const uint64_t a0 = b & ~3ull;
  if (e - b != (uint64_t)K || a0 + 4ull * (NW + 1) > pool_bytes) {
    slow_flag[i] = 1;
    return;
  }
  const uint32_t *wp = reinterpret_cast<const uint32_t *>(pool + a0);
  const uint32_t sh = (uint32_t)(b & 3) * 8;
  uint32_t raw[NW + 1], t[NW];
#pragma unroll
  for (int j = 0; j <= NW; ++j) raw[j] = __ldg(wp + j);
#pragma unroll
  for (int j = 0; j < NW; ++j) t[j] = __funnelshift_r(raw[j], raw[j + 1], sh);  // bases 4j..4j+3, first base low
  u128 x;
  const uint32_t bad = pack_words<K>(t, &x);
  if (bad) {
    slow_flag[i] = 1;
    return;
  }
  u128 canon;
  uint64_t h = canon_hash<K>(x, &canon);
  uint64_t idx = bf_index(v, h);
  if (is_ref[i]) {
    int64_t loc = key_locate(v, h, idx, canon);
    out[i] = loc != -1 ? (int32_t)*count_ptr(v, loc) : 0;
  } else {
    out[i] = (bf_test(v, idx) && v.rank) ? (int32_t)(v.bf_counts[bf_rank_of(v, idx)] & 0xFFFFu) : 0;
  }
}

// set_coverages (main.cpp:157-182): per allele slot, max over signatures of the order-dependent integer
// running mean of the non-zero k-mer weights
__global__ void __launch_bounds__(128) k_coverage(const int32_t *__restrict__ w, const uint64_t *__restrict__ sig_kmer_off,
                                                 const uint64_t *__restrict__ allele_sig_off, uint64_t n_alleles,
                                                 uint32_t *__restrict__ cov) {
  uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_alleles) return;
  uint32_t allele_cov = 0;
  for (uint64_t s = allele_sig_off[a]; s < allele_sig_off[a + 1]; ++s) {
    uint32_t curr = 0;
    int n = 0;
    for (uint64_t q = sig_kmer_off[s]; q < sig_kmer_off[s + 1]; ++q) {
      int32_t wi = w[q];
      if (wi > 0) {
        curr = (curr * (uint32_t)n + (uint32_t)wi) / (uint32_t)(n + 1);
        ++n;
      }
    }
    if (curr > allele_cov) allele_cov = curr;
  }
  cov[a] = allele_cov;
}

// ---------------------------------------------------------------------------
// K5: genotype likelihoods + posterior arg-max (var_block.hpp:224-330, 367-394)
// ---------------------------------------------------------------------------
MG_HD int genotype_one(const uint32_t *cov, const float *freq, int n, float err, int max_cov, bool haploid,
                       double *lik, int *status, int *best_gt, int *gq) {
  int ng = 0;
  for (int i = 0; i < n; ++i)
    if ((int)cov[i] > max_cov) lik[ng++] = 0.0;  // one {best,0} per offending allele
  if (ng) {
    *status = 1;
    *best_gt = 0;
    *gq = 0;
    return ng;
  }
  if (n == 1) {
    lik[0] = 1.0;
    *status = 0;
    *best_gt = 0;
    *gq = 100;
    return 1;
  }
  uint32_t tot = 0;
  for (int i = 0; i < n; ++i) tot += cov[i];
  if (tot == 0) {
    lik[0] = 0.0;
    *status = 2;
    *best_gt = 0;
    *gq = 0;
    return 1;
  }
  GenoConsts c = geno_consts(err, n);
  double total = 0.0;
  for (int g1 = 0; g1 < n; ++g1) {
    for (int g2 = g1; g2 < n; ++g2) {
      if (haploid && g2 != g1) break;
      double p = (g1 == g2) ? geno_hom(cov[g1], tot, freq[g1], c)
                            : geno_het(cov[g1], cov[g2], tot, freq[g1], freq[g2], n, c);
      lik[ng++] = p;
      total = f64_add(total, p);
    }
  }
  double best = 0.0;
  int bi = 0;
  for (int i = 0; i < ng; ++i) {
    double q = lik[i] / total;
    if (q > best) {
      best = q;
      bi = i;
    }
  }
  *status = 0;
  *best_gt = bi;
  *gq = (int)round(f64_mul(best, 100.0));
  return ng;
}

__global__ void __launch_bounds__(128) k_genotype(const uint32_t *__restrict__ cov, const float *__restrict__ freq,
                                                 const uint64_t *__restrict__ var_allele_off,
                                                 const uint64_t *__restrict__ lik_off, uint64_t n_variants, float err,
                                                 int max_cov, int haploid, double *__restrict__ lik,
                                                 int32_t *__restrict__ n_gts, int32_t *__restrict__ status,
                                                 int32_t *__restrict__ best_gt, int32_t *__restrict__ gq) {
  uint64_t vi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= n_variants) return;
  uint64_t a0 = var_allele_off[vi];
  int n = (int)(var_allele_off[vi + 1] - a0);
  int st, bg, q;
  int ng = genotype_one(cov + a0, freq + a0, n, err, max_cov, haploid != 0, lik + lik_off[vi], &st, &bg, &q);
  n_gts[vi] = ng;
  status[vi] = st;
  best_gt[vi] = bg;
  gq[vi] = q;
}

// ---------------------------------------------------------------------------
// index image export / import (the index file of `malva-geno index`, main.cpp:406-412, 455-461):
// filters travel as sorted lists of set-bit indices, ref_bf as a list of packed keys
// ---------------------------------------------------------------------------
// one thread per 256-bit unit; offs = exclusive scan of the per-unit popcounts
__global__ void __launch_bounds__(256) k_emit_bits(const uint32_t *__restrict__ words, uint64_t n_units, int stride_u32,
                                                  const uint32_t *__restrict__ offs, uint64_t n_bits,
                                                  uint64_t *__restrict__ out) {
  uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_units) return;
  uint64_t o = offs[u];
  for (int w = 0; w < 8; ++w) {
    uint32_t x = words[u * (uint64_t)stride_u32 + (uint64_t)w];
    while (x) {
      int b = __ffs(x) - 1;
      x &= x - 1;
      uint64_t idx = u * 256 + (uint64_t)w * 32 + (uint64_t)b;
      if (idx < n_bits) out[o++] = idx;
    }
  }
}
__global__ void __launch_bounds__(256) k_set_bits(const uint64_t *__restrict__ idx, uint64_t n, uint64_t n_bits,
                                                 uint32_t *words_rw, int as_lines, DevView v, uint32_t *occ_rw) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t b = idx[i];
  if (b >= n_bits) return;
  if (as_lines) occ_set(v, occ_rw, b);
  uint32_t *w = as_lines ? words_rw + (b >> 8) * 32 + ((b & 255) >> 5) : words_rw + (b >> 5);
  atomicOr(w, 1u << (b & 31));
}
// every non-empty key slot of the probe lines (n_slots = 6 * n_lines) or of the overflow table
__global__ void __launch_bounds__(256) k_emit_keys(const uint4 *__restrict__ lines, uint64_t n_lines,
                                                  const u128 *__restrict__ ovf_keys, uint64_t ovf_cap,
                                                  unsigned long long *counter, u128 *__restrict__ out, uint64_t cap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t n_slots = n_lines * LINE_KEYS;
  u128 key;
  if (i < n_slots) {
    key = key_of(lines[(i / LINE_KEYS) * LINE_U4 + 2 + (i % LINE_KEYS)]);
  } else if (i < n_slots + ovf_cap) {
    key = ovf_keys[i - n_slots];
    key.hi &= KEY_HI_MASK;
  } else {
    return;
  }
  if (key_empty(key)) return;
  unsigned long long o = atomicAdd(counter, 1ull);
  if (o < cap) out[o] = key;
}

// dst[i] += src[i] (u32, wrap-around): the counter reduce of replicated contexts (mg_reduce_counts)
__global__ void __launch_bounds__(256) k_add_u32(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint64_t n) {
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    uint4 a = *reinterpret_cast<const uint4 *>(dst + i), b = *reinterpret_cast<const uint4 *>(src + i);
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    *reinterpret_cast<uint4 *>(dst + i) = a;
  } else {
    for (; i < n; ++i) dst[i] += src[i];
  }
}

// ---------------------------------------------------------------------------
// roofline diagnostics: measured ceilings on this device (bench.py records them)
//   k_diag_random  : independent random reads, `gran` separate 4-byte loads inside one aligned
//                    gran*32-byte unit (1, 2 or 4 sectors)
//   k_diag_lines   : random 128-byte lines, each fetched by 8 lanes x 16 B in ONE coalesced request
//                    (the access pattern of k_scan)
//   k_diag_stream  : streaming 16-byte reads
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag_random(const uint32_t *__restrict__ buf, uint64_t n_units, int gran,
                                                    uint64_t per_thread, uint32_t *sink) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = (t + 1) * GOLD;
  uint32_t acc = 0;
#pragma unroll 4
  for (uint64_t i = 0; i < per_thread; ++i) {
    s ^= s >> 29;
    s *= 0xBF58476D1CE4E5B9ULL;
    s ^= s >> 32;
    uint64_t unit = mulhi64(s, n_units);  // uniform in [0, n_units)
    const uint32_t *p = buf + unit * 8 * (uint64_t)gran;
    acc += __ldg(p);
    if (gran >= 2) acc += __ldg(p + 8);
    if (gran >= 4) acc += __ldg(p + 16) + __ldg(p + 24);
    s += GOLD;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void __launch_bounds__(256) k_diag_lines(const uint4 *__restrict__ buf, uint64_t n_units, int lanes_log2,
                                                   uint64_t per_group, uint32_t *sink) {
  // groups of 2^lanes_log2 lanes (8 / 4 / 2) read one random aligned unit of 128 / 64 / 32 bytes per instruction
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = ((t >> lanes_log2) + 1) * GOLD;  // one random stream per lane group
  const int sub = threadIdx.x & ((1 << lanes_log2) - 1);
  uint32_t acc = 0;
#pragma unroll 8
  for (uint64_t i = 0; i < per_group; ++i) {
    s ^= s >> 29;
    s *= 0xBF58476D1CE4E5B9ULL;
    s ^= s >> 32;
    uint64_t unit = mulhi64(s, n_units);
    uint4 q = __ldg(buf + (unit << lanes_log2) + sub);
    acc += q.x ^ q.y ^ q.z ^ q.w;
    s += GOLD;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void __launch_bounds__(256) k_diag_stream(const uint4 *__restrict__ buf, uint64_t n16, uint32_t *sink) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 q = __ldg(buf + i);
    acc += q.x ^ q.y ^ q.z ^ q.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

}  // namespace mg
