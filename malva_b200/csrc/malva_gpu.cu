// libmalva_gpu.so -- hand-written sm_100a kernels + the C ABI of include/malva_gpu.h.
//
// Device data layout (all resident in HBM for the life of a context):
//   bf_words / ctx_words : the two one-hash Bloom filters as u32 words, bit i of
//                          the filter = bit (i & 31) of word (i >> 5)
//   bf_rank              : ones before each 512-bit block of bf (u32, n_blocks+1)
//   bf_counts            : one u32 accumulator per set bit of bf, indexed by rank
//                          (read back & 0xFFFF == the reference's uint16 wrap-around)
//   tab_keys / tab_counts: open-addressing (linear probing, load <= 0.5) exact
//                          table of canonical packed ref-allele k-mers + u32 counts
// Kernels: k_add_signatures (K3), k_block_popc (+CUB scan) (K3), k_refpass (K2),
//          k_scan (K1), k_lookup / k_coverage / k_genotype (K4, K5).
// There is no CPU fallback anywhere: every entry point fails with MG_ERR_CUDA
// when no device is usable.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/malva_gpu.h"
#include "geno.cuh"
#include "xxh3.cuh"

using mg::u128;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return set_err(MG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

extern "C" const char *mg_last_error(void) { return g_err; }
extern "C" int mg_version(void) { return 100; }
extern "C" int mg_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_err(MG_ERR_CUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return MG_ERR_CUDA;
  }
  return n;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ULL;
constexpr int RANK_SHIFT = 9;  // 512-bit rank blocks = 16 u32 words = 64 B
constexpr uint64_t STAGE_KMERS = 1ull << 22;

struct DevView {  // everything the kernels need, passed by value
  const uint32_t *bf_words;
  const uint32_t *ctx_words;
  const uint32_t *bf_rank;
  uint32_t *bf_counts;
  const u128 *tab_keys;
  uint32_t *tab_counts;
  uint64_t bf_bits;
  uint64_t bf_mask;   // bf_bits-1 when bf_bits is a power of two, else 0
  uint64_t tab_mask;  // capacity-1
  int tab_shift;      // 64 - log2(capacity)
  int k, ref_k;
};

struct mg_ctx {
  int device = 0, k = 0, ref_k = 0, sms = 0;
  uint64_t bf_bits = 0, n_words32 = 0, n_blocks = 0;
  uint32_t *bf_words = nullptr, *ctx_words = nullptr, *bf_rank = nullptr, *bf_counts = nullptr;
  uint64_t bf_ones = 0;
  bool alt_final = false, ctx_final = false;
  u128 *tab_keys = nullptr;
  uint32_t *tab_counts = nullptr;
  int tab_log2 = 0;
  uint64_t tab_n = 0;
  unsigned long long *d_scalars = nullptr;  // [0] new table keys, [1] irregular count, [2] popcount, [3] error flag
  std::unordered_map<std::string, int> irregular_ref;  // ref keys that are not k symbols of ACGT (always count 0)
  cudaStream_t stream[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  void *d_stage_k[2] = {nullptr, nullptr};
  uint32_t *d_stage_c[2] = {nullptr, nullptr};
  int next_stage = 0;
  int scan_ilp = 2, scan_ctas_per_sm = 8;
  uint64_t launches = 0;  // kernels launched by this context (bench.py's gpu_launches)
  cudaEvent_t tj = nullptr;
  cudaEvent_t ge[4] = {nullptr, nullptr, nullptr, nullptr};
  void *geno_scratch = nullptr;  // per-k-mer weights + ref flags of mg_genotype
  uint64_t geno_scratch_bytes = 0;
  cudaEvent_t evs[64] = {};

  DevView view() const {
    DevView v;
    v.bf_words = bf_words;
    v.ctx_words = ctx_words;
    v.bf_rank = bf_rank;
    v.bf_counts = bf_counts;
    v.tab_keys = tab_keys;
    v.tab_counts = tab_counts;
    v.bf_bits = bf_bits;
    v.bf_mask = (bf_bits & (bf_bits - 1)) == 0 ? bf_bits - 1 : 0;
    v.tab_mask = (1ull << tab_log2) - 1;
    v.tab_shift = 64 - tab_log2;
    v.k = k;
    v.ref_k = ref_k;
    return v;
  }
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bf_index(const DevView &v, uint64_t h) {
  return v.bf_mask ? (h & v.bf_mask) : (h % v.bf_bits);
}
__device__ __forceinline__ bool test_bit(const uint32_t *words, uint64_t idx) {
  return (__ldg(words + (idx >> 5)) >> (idx & 31)) & 1u;
}
__device__ __forceinline__ uint64_t tab_slot0(const DevView &v, uint64_t h) { return (h * GOLD) >> v.tab_shift; }

__device__ __forceinline__ u128 ld_key(const u128 *p) {
  uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
  u128 r;
  r.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
  r.hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
  return r;
}
__device__ __forceinline__ bool key_eq(u128 a, u128 b) { return a.lo == b.lo && a.hi == b.hi; }
__device__ __forceinline__ bool key_empty(u128 a) { return (a.lo & a.hi) == ~0ull; }

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

// rank of a set bit = ones strictly before idx (sdsl rank_support_v<1> semantics)
__device__ __forceinline__ uint32_t bf_rank_of(const DevView &v, uint64_t idx) {
  uint64_t blk = idx >> RANK_SHIFT;
  uint32_t r = __ldg(v.bf_rank + blk);
  uint64_t w0 = blk << (RANK_SHIFT - 5), w = idx >> 5;
  for (uint64_t x = w0; x < w; ++x) r += __popc(__ldg(v.bf_words + x));
  r += __popc(__ldg(v.bf_words + w) & ((1u << (idx & 31)) - 1u));
  return r;
}

template <int K>
__device__ __forceinline__ uint64_t canon_hash_k(u128 x, int k, u128 *canon) {
  if constexpr (K > 0) {
    return mg::canon_hash<K>(x, canon);
  } else {
    return mg::canon_hash_rt(x, k, canon);
  }
}

// exact-table lookup; returns slot or ~0
__device__ __forceinline__ uint64_t tab_find(const DevView &v, uint64_t h, u128 canon) {
  uint64_t slot = tab_slot0(v, h);
  while (true) {
    u128 key = ld_key(v.tab_keys + slot);
    if (key_eq(key, canon)) return slot;
    if (key_empty(key)) return ~0ull;
    slot = (slot + 1) & v.tab_mask;
  }
}

// ---------------------------------------------------------------------------
// K3a: index-time inserts (add_kmers_to_bf, main.cpp:122-144)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_add_signatures(const uint8_t *__restrict__ pool,
                                                       const uint64_t *__restrict__ off,
                                                       const uint8_t *__restrict__ is_ref, uint64_t n,
                                                       DevView v, uint32_t *bf_words_rw, u128 *tab_keys_rw,
                                                       unsigned long long *scalars, uint32_t *irregular_idx) {
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
  bool regular = mg::pack_ascii(s, len, v.k, &x);
  if (is_ref[i]) {  // ref_bf.add_key
    if (!regular) {
      unsigned long long p = atomicAdd(&scalars[1], 1ull);
      irregular_idx[p] = (uint32_t)i;
      return;
    }
    uint64_t h = mg::canon_hash_rt(x, v.k, &canon);
    uint64_t slot = tab_slot0(v, h);
    const u128 empty = {~0ull, ~0ull};
    while (true) {
      u128 old = cas128(tab_keys_rw + slot, empty, canon);
      if (key_empty(old)) {
        atomicAdd(&scalars[0], 1ull);
        break;
      }
      if (key_eq(old, canon)) break;  // kmers[ckmer] = 0 on an existing key: counts are still 0 at index time
      slot = (slot + 1) & v.tab_mask;
    }
    v.tab_counts[slot] = 0;
  } else {  // bf.add_key
    uint64_t h = regular ? mg::canon_hash_rt(x, v.k, &canon) : mg::hash_ascii(s, len);
    uint64_t idx = bf_index(v, h);
    atomicOr(bf_words_rw + (idx >> 5), 1u << (idx & 31));
  }
}

// same inserts for signature k-mers that arrive already packed (exactly k symbols of ACGT)
__global__ void __launch_bounds__(256) k_add_packed(const uint4 *__restrict__ kmers, const uint8_t *__restrict__ is_ref,
                                                   uint64_t n, DevView v, uint32_t *bf_words_rw, u128 *tab_keys_rw,
                                                   unsigned long long *scalars) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4 q = kmers[i];
  u128 x, canon;
  x.lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
  x.hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
  x = mg::mask128(x, 2 * v.k);
  uint64_t h = mg::canon_hash_rt(x, v.k, &canon);
  if (is_ref[i]) {
    uint64_t slot = tab_slot0(v, h);
    const u128 empty = {~0ull, ~0ull};
    while (true) {
      u128 old = cas128(tab_keys_rw + slot, empty, canon);
      if (key_empty(old)) {
        atomicAdd(&scalars[0], 1ull);
        break;
      }
      if (key_eq(old, canon)) break;
      slot = (slot + 1) & v.tab_mask;
    }
    v.tab_counts[slot] = 0;
  } else {
    uint64_t idx = bf_index(v, h);
    atomicOr(bf_words_rw + (idx >> 5), 1u << (idx & 31));
  }
}

__global__ void k_fill_keys(u128 *keys, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i].lo = ~0ull;
    keys[i].hi = ~0ull;
  }
}

// re-insert every key of an old table into a larger one (counts carried over)
__global__ void k_rehash(const u128 *old_keys, const uint32_t *old_counts, uint64_t old_cap, DevView v,
                         u128 *tab_keys_rw) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  u128 key = old_keys[i];
  if (key_empty(key)) return;
  // canonical keys re-hash through their ASCII image, exactly like a fresh insert
  u128 canon;
  uint64_t h = mg::canon_hash_rt(key, v.k, &canon);
  uint64_t slot = tab_slot0(v, h);
  const u128 empty = {~0ull, ~0ull};
  while (true) {
    u128 old = cas128(tab_keys_rw + slot, empty, key);
    if (key_empty(old)) break;
    slot = (slot + 1) & v.tab_mask;
  }
  v.tab_counts[slot] = old_counts[i];
}

// ---------------------------------------------------------------------------
// K3b: switch_mode (bloom_filter.hpp:93-98): per-block popcounts, then a scan
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_popc(const uint32_t *__restrict__ words, uint64_t n_blocks,
                                                   uint64_t n_words, uint32_t *__restrict__ blk_count,
                                                   unsigned long long *total) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  if (b < n_blocks) {
    uint64_t w0 = b << (RANK_SHIFT - 5);
    if (w0 + 16 <= n_words) {
      const uint4 *p = reinterpret_cast<const uint4 *>(words + w0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = p[j];
        c += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
      }
    } else {
      for (uint64_t w = w0; w < n_words; ++w) c += __popc(words[w]);
    }
    blk_count[b] = c;
  }
  // block-level reduction -> one 64-bit atomic per CTA
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

// ---------------------------------------------------------------------------
// K2: reference rolling pass (main.cpp:385-400)
// Each CTA stages a tile of the contig in shared memory (with a ref_k-1 halo);
// each thread rolls RUN consecutive windows through 2-bit registers.  Windows
// that contain a non-ACGT symbol take the byte-exact ASCII path.
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
  const u128 m = mg::mask128(u128{~0ull, ~0ull}, 2 * ref_k);
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
      u128 x35 = mg::mask128(mg::shr128(x, 2 * shift), 2 * k), canon;
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
      h35 = mg::hash_ascii(s, k);
    }
    uint64_t idx = bf_index(v, h35);
    if (!test_bit(v.bf_words, idx)) continue;
    uint64_t h43;
    if ((bad & m43) == 0) {
      u128 canon;
      h43 = canon_hash_k<REFK>(x, ref_k, &canon);
    } else {
      uint8_t s[64];
      for (int j = 0; j < ref_k; ++j) s[j] = sm[sp - ref_k + 1 + j];
      h43 = mg::hash_ascii(s, ref_k);
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
  uint64_t h = mg::hash_ascii(seq + d, kl);
  if (!test_bit(v.bf_words, bf_index(v, h))) return;
  uint64_t hc = mg::hash_ascii(seq, (int)len);
  uint64_t cidx = bf_index(v, hc);
  atomicOr(ctx_words_rw + (cidx >> 5), 1u << (cidx & 31));
}

// ---------------------------------------------------------------------------
// K1: sample k-mer scan (main.cpp:487-500)
//   ref_bf.increment(kmer, c);  if (!context_bf.test_key(context)) bf.increment(kmer, c);
// One 16-byte coalesced load per k-mer, two independent random probes (alt
// filter word + table bucket) issued back to back; the context filter, the
// rank directory and the counters are touched only on the ~1% hit path.
// ---------------------------------------------------------------------------
template <int K, int REFK, int ILP>
__global__ void __launch_bounds__(256) k_scan(const uint4 *__restrict__ kmers, const uint32_t *__restrict__ counts,
                                              uint64_t n, DevView v) {
  const int k = K > 0 ? K : v.k, ref_k = REFK > 0 ? REFK : v.ref_k;
  const int d = (ref_k - k) / 2;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * ILP) {
    u128 x43[ILP], canon[ILP];
    uint64_t h[ILP], idx[ILP], slot[ILP];
    uint32_t cnt[ILP], word[ILP];
    u128 key[ILP];
    bool live[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      uint64_t i = i0 + (uint64_t)u * stride;
      live[u] = i < n;
      uint4 q = live[u] ? __ldg(kmers + i) : make_uint4(0, 0, 0, 0);
      cnt[u] = live[u] ? __ldg(counts + i) : 0u;
      x43[u].lo = (uint64_t)q.x | ((uint64_t)q.y << 32);
      x43[u].hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      // kmer = context + (ref_k-k)/2 (main.cpp:493): the k-mer starts d bases in, so ref_k-k-d bases follow it
      u128 x35 = mg::mask128(mg::shr128(x43[u], 2 * (ref_k - k - d)), 2 * k);
      h[u] = canon_hash_k<K>(x35, k, &canon[u]);
      idx[u] = bf_index(v, h[u]);
      slot[u] = tab_slot0(v, h[u]);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      word[u] = __ldg(v.bf_words + (idx[u] >> 5));
      key[u] = ld_key(v.tab_keys + slot[u]);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (!live[u]) continue;
      // exact table: linear probing until the key or an empty slot
      u128 kk = key[u];
      uint64_t s = slot[u];
      while (true) {
        if (key_eq(kk, canon[u])) {
          atomicAdd(v.tab_counts + s, cnt[u]);
          break;
        }
        if (key_empty(kk)) break;
        s = (s + 1) & v.tab_mask;
        kk = ld_key(v.tab_keys + s);
      }
      if ((word[u] >> (idx[u] & 31)) & 1u) {
        u128 c43;
        uint64_t h43 = canon_hash_k<REFK>(x43[u], ref_k, &c43);
        if (!test_bit(v.ctx_words, bf_index(v, h43))) atomicAdd(v.bf_counts + bf_rank_of(v, idx[u]), cnt[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K4: signature look-ups (BF::get_count / KMAP::get_count) + coverage
// ---------------------------------------------------------------------------
// mode 0: get_count  (is_ref selects KMAP/BF, out = int32 count)
// mode 1: test_key on filter/table `which` (out = 0/1)
__global__ void __launch_bounds__(128) k_lookup(const uint8_t *__restrict__ pool, const uint64_t *__restrict__ off,
                                               const uint8_t *__restrict__ is_ref, uint64_t n, DevView v, int mode,
                                               int which, int32_t *__restrict__ out, unsigned long long *scalars) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
    if (!mg::pack_ascii(s, len, v.k, &x)) {
      out[i] = (mode == 1) ? -1 : 0;  // irregular keys are resolved on the host (always count 0)
      return;
    }
    uint64_t h = mg::canon_hash_rt(x, v.k, &canon);
    uint64_t slot = tab_find(v, h, canon);
    if (mode == 1)
      out[i] = slot != ~0ull;
    else
      out[i] = slot != ~0ull ? (int32_t)v.tab_counts[slot] : 0;
    return;
  }
  // a Bloom filter: hash the canonical ASCII bytes of whatever length was given
  bool regular = mg::pack_ascii(s, len, len, &x) && len >= 1 && len <= 64;
  uint64_t h = regular ? mg::canon_hash_rt(x, len, &canon) : mg::hash_ascii(s, len);
  uint64_t idx = bf_index(v, h);
  const uint32_t *words = (mode == 1 && which == 1) ? v.ctx_words : v.bf_words;
  bool set = test_bit(words, idx);
  if (mode == 1) {
    out[i] = set;
  } else {
    out[i] = (set && v.bf_rank) ? (int32_t)(v.bf_counts[bf_rank_of(v, idx)] & 0xFFFFu) : 0;
  }
}

// set_coverages (main.cpp:157-182): per allele slot, max over signatures of the
// order-dependent integer running mean of the non-zero k-mer weights
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
  mg::GenoConsts c = mg::geno_consts(err, n);
  double total = 0.0;
  for (int g1 = 0; g1 < n; ++g1) {
    for (int g2 = g1; g2 < n; ++g2) {
      if (haploid && g2 != g1) break;
      double p = (g1 == g2) ? mg::geno_hom(cov[g1], tot, freq[g1], c)
                            : mg::geno_het(cov[g1], cov[g2], tot, freq[g1], freq[g2], n, c);
      lik[ng++] = p;
      total = mg::f64_add(total, p);
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
  *gq = (int)round(mg::f64_mul(best, 100.0));
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
// roofline diagnostics: measured ceilings for independent random sector reads
// and for a streaming read on this device (bench.py records them next to the
// driver's MEASURED_PEAKS.json)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag_random(const uint32_t *__restrict__ buf, uint64_t n_units, int gran,
                                                    uint64_t per_thread, uint32_t *sink) {
  // every access touches `gran` consecutive 32-byte sectors of one aligned gran*32-byte unit
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = (t + 1) * GOLD;
  uint32_t acc = 0;
#pragma unroll 4
  for (uint64_t i = 0; i < per_thread; ++i) {
    s ^= s >> 29;
    s *= 0xBF58476D1CE4E5B9ULL;
    s ^= s >> 32;
    uint64_t unit = mg::mulhi64(s, n_units);  // uniform in [0, n_units)
    const uint32_t *p = buf + unit * 8 * (uint64_t)gran;
    acc += __ldg(p);
    if (gran >= 2) acc += __ldg(p + 8);
    if (gran >= 4) acc += __ldg(p + 16) + __ldg(p + 24);
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

// ---------------------------------------------------------------------------
// host side of the C ABI
// ---------------------------------------------------------------------------
static int grid_for(uint64_t n, int threads) { return (int)((n + (uint64_t)threads - 1) / (uint64_t)threads); }

static int tab_alloc(mg_ctx *c, int log2cap, u128 **keys, uint32_t **counts) {
  uint64_t cap = 1ull << log2cap;
  CU(cudaMalloc(keys, cap * sizeof(u128)));
  CU(cudaMalloc(counts, cap * sizeof(uint32_t)));
  c->launches++;
  k_fill_keys<<<grid_for(cap, 256), 256, 0, c->stream[0]>>>(*keys, cap);
  CU(cudaGetLastError());
  CU(cudaMemsetAsync(*counts, 0, cap * sizeof(uint32_t), c->stream[0]));
  return MG_OK;
}

// make room for `extra` more keys at load <= 0.5
static int tab_reserve(mg_ctx *c, uint64_t extra) {
  uint64_t need = (c->tab_n + extra) * 2;
  if (need <= (1ull << c->tab_log2)) return MG_OK;
  int nl = c->tab_log2;
  while ((1ull << nl) < need) ++nl;
  u128 *nk = nullptr;
  uint32_t *nc = nullptr;
  int rc = tab_alloc(c, nl, &nk, &nc);
  if (rc) return rc;
  u128 *ok = c->tab_keys;
  uint32_t *oc = c->tab_counts;
  uint64_t ocap = 1ull << c->tab_log2;
  c->tab_keys = nk;
  c->tab_counts = nc;
  c->tab_log2 = nl;
  if (c->tab_n) {
    c->launches++;
    k_rehash<<<grid_for(ocap, 256), 256, 0, c->stream[0]>>>(ok, oc, ocap, c->view(), nk);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream[0]));
  CU(cudaFree(ok));
  CU(cudaFree(oc));
  return MG_OK;
}

extern "C" int mg_create(mg_ctx **out, int device, int k, int ref_k, uint64_t bf_bits) {
  if (!out) return set_err(MG_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (k < 1 || ref_k < k || ref_k > 64)
    return set_err(MG_ERR_ARG, "unsupported k=%d ref_k=%d (need 1 <= k <= ref_k <= 64)", k, ref_k);
  if (bf_bits == 0) return set_err(MG_ERR_ARG, "bf_bits must be > 0 (the reference divides by it)");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_err(MG_ERR_CUDA, "device %d not available (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  mg_ctx *c = new mg_ctx();
  c->device = device;
  c->k = k;
  c->ref_k = ref_k;
  c->bf_bits = bf_bits;
  c->n_words32 = ((bf_bits + 511) / 512) * 16;  // whole rank blocks
  c->n_blocks = c->n_words32 / 16;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sms = prop.multiProcessorCount;
  // tuning knobs (read once per context; defaults are the measured best)
  if (const char *e = getenv("MG_SCAN_ILP")) c->scan_ilp = atoi(e);
  if (const char *e = getenv("MG_SCAN_CTAS_PER_SM")) c->scan_ctas_per_sm = atoi(e) > 0 ? atoi(e) : 8;
  {
    // The probes of this workload are independent random 4..16-byte reads: ask the L2 to fetch single
    // 32-byte sectors from HBM instead of promoting every miss to a wider fetch.
    size_t gran = 32;
    if (const char *e = getenv("MG_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(e);
    if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  for (int i = 0; i < 2; ++i) {
    CU(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
  }
  CU(cudaMalloc(&c->bf_words, c->n_words32 * 4));
  CU(cudaMalloc(&c->ctx_words, c->n_words32 * 4));
  CU(cudaMemsetAsync(c->bf_words, 0, c->n_words32 * 4, c->stream[0]));
  CU(cudaMemsetAsync(c->ctx_words, 0, c->n_words32 * 4, c->stream[0]));
  CU(cudaMalloc(&c->d_scalars, 8 * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->d_scalars, 0, 8 * sizeof(unsigned long long), c->stream[0]));
  c->tab_log2 = 10;
  int rc = tab_alloc(c, c->tab_log2, &c->tab_keys, &c->tab_counts);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream[0]));
  *out = c;
  return MG_OK;
}

extern "C" void mg_destroy(mg_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->bf_words);
  cudaFree(c->ctx_words);
  cudaFree(c->bf_rank);
  cudaFree(c->bf_counts);
  cudaFree(c->tab_keys);
  cudaFree(c->tab_counts);
  cudaFree(c->d_scalars);
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->d_stage_k[i]);
    cudaFree(c->d_stage_c[i]);
    if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  }
  if (c->tj) cudaEventDestroy(c->tj);
  for (int i = 0; i < 64; ++i)
    if (c->evs[i]) cudaEventDestroy(c->evs[i]);
  cudaFree(c->geno_scratch);
  for (int i = 0; i < 4; ++i)
    if (c->ge[i]) cudaEventDestroy(c->ge[i]);
  delete c;
}

static int check_too_long(mg_ctx *c) {
  unsigned long long flag = 0;
  CU(cudaMemcpy(&flag, c->d_scalars + 3, 8, cudaMemcpyDeviceToHost));
  if (flag) {
    CU(cudaMemset(c->d_scalars + 3, 0, 8));
    return set_err(MG_ERR_ARG, "signature k-mer longer than 128 bytes");
  }
  return MG_OK;
}

// upload an ASCII pool + offsets (+ optional flags)
struct DevBatch {
  uint8_t *pool = nullptr;
  uint64_t *off = nullptr;
  uint8_t *flags = nullptr;
  ~DevBatch() {
    cudaFree(pool);
    cudaFree(off);
    cudaFree(flags);
  }
};
static int upload_batch(mg_ctx *c, DevBatch &b, const char *pool, const uint64_t *off, const uint8_t *flags,
                        uint64_t n) {
  uint64_t bytes = off[n];
  CU(cudaMalloc(&b.pool, bytes ? bytes : 1));
  CU(cudaMalloc(&b.off, (n + 1) * 8));
  CU(cudaMemcpyAsync(b.pool, pool, bytes, cudaMemcpyHostToDevice, c->stream[0]));
  CU(cudaMemcpyAsync(b.off, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream[0]));
  if (flags) {
    CU(cudaMalloc(&b.flags, n ? n : 1));
    CU(cudaMemcpyAsync(b.flags, flags, n, cudaMemcpyHostToDevice, c->stream[0]));
  }
  return MG_OK;
}

extern "C" int mg_add_signatures(mg_ctx *c, const char *pool, const uint64_t *off, const uint8_t *is_ref,
                                 uint64_t n) {
  if (!c || !off || !is_ref || (!pool && n && off[n])) return set_err(MG_ERR_ARG, "NULL argument");
  if (c->alt_final) return set_err(MG_ERR_STATE, "mg_add_signatures after mg_finalize_alt");
  if (n == 0) return MG_OK;
  if (n > 0xFFFFFFFFull) return set_err(MG_ERR_ARG, "batch too large (max 2^32-1 k-mers per call)");
  CU(cudaSetDevice(c->device));
  uint64_t n_ref = 0;
  for (uint64_t i = 0; i < n; ++i) n_ref += is_ref[i] != 0;
  int rc = tab_reserve(c, n_ref);
  if (rc) return rc;
  DevBatch b;
  rc = upload_batch(c, b, pool, off, is_ref, n);
  if (rc) return rc;
  uint32_t *d_irr = nullptr;
  CU(cudaMalloc(&d_irr, (n_ref ? n_ref : 1) * 4));
  CU(cudaMemsetAsync(c->d_scalars, 0, 2 * sizeof(unsigned long long), c->stream[0]));
  c->launches++;
  k_add_signatures<<<grid_for(n, 128), 128, 0, c->stream[0]>>>(b.pool, b.off, b.flags, n, c->view(), c->bf_words,
                                                              c->tab_keys, c->d_scalars, d_irr);
  CU(cudaGetLastError());
  unsigned long long sc[2];
  CU(cudaMemcpyAsync(sc, c->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[0]));
  c->tab_n += sc[0];
  if (sc[1]) {  // ref keys that are not k x ACGT: keep them host-side (kmap.hpp:86-112 semantics)
    std::vector<uint32_t> idx(sc[1]);
    CU(cudaMemcpy(idx.data(), d_irr, sc[1] * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i : idx) {
      int len = (int)(off[i + 1] - off[i]);
      uint64_t w[18];
      int cut = mg::canonical_ascii(reinterpret_cast<const uint8_t *>(pool) + off[i], len, w);
      c->irregular_ref[std::string(reinterpret_cast<const char *>(w), (size_t)cut)] = 0;
    }
  }
  cudaFree(d_irr);
  return check_too_long(c);
}

extern "C" int mg_finalize_alt(mg_ctx *c) {
  if (!c) return set_err(MG_ERR_ARG, "NULL ctx");
  if (c->alt_final) return MG_OK;
  CU(cudaSetDevice(c->device));
  uint32_t *blk = nullptr;
  CU(cudaMalloc(&blk, (c->n_blocks + 1) * 4));
  CU(cudaMalloc(&c->bf_rank, (c->n_blocks + 1) * 4));
  CU(cudaMemsetAsync(blk + c->n_blocks, 0, 4, c->stream[0]));
  CU(cudaMemsetAsync(c->d_scalars + 2, 0, 8, c->stream[0]));
  c->launches++;
  k_block_popc<<<grid_for(c->n_blocks, 256), 256, 0, c->stream[0]>>>(c->bf_words, c->n_blocks, c->n_words32, blk,
                                                                    c->d_scalars + 2);
  CU(cudaGetLastError());
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, blk, c->bf_rank, (int64_t)(c->n_blocks + 1), c->stream[0]));
  CU(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, blk, c->bf_rank, (int64_t)(c->n_blocks + 1), c->stream[0]));
  unsigned long long ones = 0;
  CU(cudaMemcpyAsync(&ones, c->d_scalars + 2, 8, cudaMemcpyDeviceToHost, c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[0]));
  cudaFree(tmp);
  cudaFree(blk);
  if (ones > 0xFFFFFFFFull) return set_err(MG_ERR_ARG, "bf has %llu set bits; rank directory is 32-bit", ones);
  c->bf_ones = ones;
  CU(cudaMalloc(&c->bf_counts, (ones ? ones : 1) * 4));
  CU(cudaMemset(c->bf_counts, 0, (ones ? ones : 1) * 4));
  c->alt_final = true;
  return MG_OK;
}

template <int K, int REFK>
static cudaError_t launch_refpass(mg_ctx *c, const uint8_t *d_seq, uint64_t len) {
  uint64_t n_pos = len - (uint64_t)(c->ref_k - 1);
  int grid = (int)((n_pos + RP_TILE - 1) / RP_TILE);
  size_t smem = RP_TILE + 64;
  c->launches++;
  k_refpass<K, REFK><<<grid, RP_THREADS, smem, c->stream[0]>>>(d_seq, len, c->view(), c->ctx_words);
  return cudaGetLastError();
}

extern "C" int mg_scan_reference(mg_ctx *c, const char *seq, uint64_t len) {
  if (!c || (!seq && len)) return set_err(MG_ERR_ARG, "NULL argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "mg_scan_reference before mg_finalize_alt");
  if (c->ctx_final) return set_err(MG_ERR_STATE, "mg_scan_reference after mg_finalize_context");
  CU(cudaSetDevice(c->device));
  int d = (c->ref_k - c->k) / 2;
  if (len < (uint64_t)c->ref_k) {
    // the reference's substr(d, k) throws when d > size(); a shorter contig is hashed once, truncated
    if ((uint64_t)d > len) return set_err(MG_ERR_ARG, "contig shorter than (ref_k-k)/2: the reference aborts here");
    uint8_t *d_seq = nullptr;
    CU(cudaMalloc(&d_seq, len ? len : 1));
    CU(cudaMemcpyAsync(d_seq, seq, len, cudaMemcpyHostToDevice, c->stream[0]));
    c->launches++;
    k_refpass_short<<<1, 32, 0, c->stream[0]>>>(d_seq, len, c->view(), c->ctx_words);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d_seq);
    return MG_OK;
  }
  uint8_t *d_seq = nullptr;
  CU(cudaMalloc(&d_seq, len));
  CU(cudaMemcpyAsync(d_seq, seq, len, cudaMemcpyHostToDevice, c->stream[0]));
  cudaError_t e;
  if (c->k == 35 && c->ref_k == 43)
    e = launch_refpass<35, 43>(c, d_seq, len);
  else
    e = launch_refpass<0, 0>(c, d_seq, len);
  if (e != cudaSuccess) {
    cudaFree(d_seq);
    return set_err(MG_ERR_CUDA, "k_refpass launch -> %s", cudaGetErrorString(e));
  }
  CU(cudaStreamSynchronize(c->stream[0]));
  cudaFree(d_seq);
  return MG_OK;
}

extern "C" int mg_finalize_context(mg_ctx *c) {
  if (!c) return set_err(MG_ERR_ARG, "NULL ctx");
  c->ctx_final = true;  // context_bf is only ever test_key()'d after this point: no rank/counters needed
  return MG_OK;
}

template <int K, int REFK, int ILP>
static cudaError_t launch_scan_ilp(mg_ctx *c, const void *d_lohi, const void *d_counts, uint64_t n, cudaStream_t st) {
  uint64_t want = (n + 256ull * ILP - 1) / (256ull * ILP);
  uint64_t cap = (uint64_t)c->sms * (uint64_t)c->scan_ctas_per_sm;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  c->launches++;
  k_scan<K, REFK, ILP><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(d_lohi),
                                             reinterpret_cast<const uint32_t *>(d_counts), n, c->view());
  return cudaGetLastError();
}
template <int K, int REFK>
static cudaError_t launch_scan(mg_ctx *c, const void *d_lohi, const void *d_counts, uint64_t n, cudaStream_t st) {
  switch (c->scan_ilp) {
    case 1: return launch_scan_ilp<K, REFK, 1>(c, d_lohi, d_counts, n, st);
    case 4: return launch_scan_ilp<K, REFK, 4>(c, d_lohi, d_counts, n, st);
    default: return launch_scan_ilp<K, REFK, 2>(c, d_lohi, d_counts, n, st);
  }
}

static int scan_device(mg_ctx *c, const void *d_lohi, const void *d_counts, uint64_t n, cudaStream_t st) {
  cudaError_t e;
  if (c->k == 35 && c->ref_k == 43)
    e = launch_scan<35, 43>(c, d_lohi, d_counts, n, st);
  else
    e = launch_scan<0, 0>(c, d_lohi, d_counts, n, st);
  if (e != cudaSuccess) return set_err(MG_ERR_CUDA, "k_scan launch -> %s", cudaGetErrorString(e));
  return MG_OK;
}

extern "C" int mg_scan_sample_kmers_device(mg_ctx *c, const void *d_lohi, const void *d_counts, uint64_t n) {
  if (!c || ((!d_lohi || !d_counts) && n)) return set_err(MG_ERR_ARG, "NULL argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "scan before mg_finalize_alt (BF::increment is a no-op in write mode)");
  if (n == 0) return MG_OK;
  CU(cudaSetDevice(c->device));
  return scan_device(c, d_lohi, d_counts, n, c->stream[0]);
}

extern "C" int mg_scan_sample_kmers(mg_ctx *c, const uint64_t *lohi, const uint32_t *counts, uint64_t n) {
  if (!c || ((!lohi || !counts) && n)) return set_err(MG_ERR_ARG, "NULL argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "scan before mg_finalize_alt (BF::increment is a no-op in write mode)");
  CU(cudaSetDevice(c->device));
  for (int i = 0; i < 2; ++i) {
    if (!c->d_stage_k[i]) {
      CU(cudaMalloc(&c->d_stage_k[i], STAGE_KMERS * 16));
      CU(cudaMalloc(&c->d_stage_c[i], STAGE_KMERS * 4));
    }
  }
  // chunks alternate between two (stream, device buffer) pairs: the H2D copy of
  // chunk i+1 overlaps the kernel of chunk i.
  for (uint64_t o = 0; o < n; o += STAGE_KMERS) {
    uint64_t m = n - o < STAGE_KMERS ? n - o : STAGE_KMERS;
    int s = c->next_stage;
    c->next_stage ^= 1;
    CU(cudaMemcpyAsync(c->d_stage_k[s], lohi + 2 * o, m * 16, cudaMemcpyHostToDevice, c->stream[s]));
    CU(cudaMemcpyAsync(c->d_stage_c[s], counts + o, m * 4, cudaMemcpyHostToDevice, c->stream[s]));
    int rc = scan_device(c, c->d_stage_k[s], c->d_stage_c[s], m, c->stream[s]);
    if (rc) return rc;
  }
  return MG_OK;
}

extern "C" int mg_sync(mg_ctx *c) {
  if (!c) return set_err(MG_ERR_ARG, "NULL ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[1]));
  return MG_OK;
}

static int lookup_common(mg_ctx *c, const char *pool, const uint64_t *off, const uint8_t *is_ref, uint64_t n,
                         int mode, int which, int32_t *out_host) {
  CU(cudaSetDevice(c->device));
  int rc = mg_sync(c);
  if (rc) return rc;
  DevBatch b;
  rc = upload_batch(c, b, pool, off, is_ref, n);
  if (rc) return rc;
  int32_t *d_out = nullptr;
  CU(cudaMalloc(&d_out, (n ? n : 1) * 4));
  c->launches++;
  k_lookup<<<grid_for(n, 128), 128, 0, c->stream[0]>>>(b.pool, b.off, b.flags, n, c->view(), mode, which, d_out,
                                                      c->d_scalars);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_host, d_out, n * 4, cudaMemcpyDeviceToHost, c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[0]));
  cudaFree(d_out);
  return check_too_long(c);
}

extern "C" int mg_test_keys(mg_ctx *c, int which, const char *pool, const uint64_t *off, uint64_t n, uint8_t *out) {
  if (!c || !off || !out || which < 0 || which > 2) return set_err(MG_ERR_ARG, "bad argument");
  if (n == 0) return MG_OK;
  std::vector<int32_t> tmp(n);
  int rc = lookup_common(c, pool, off, nullptr, n, 1, which, tmp.data());
  if (rc) return rc;
  for (uint64_t i = 0; i < n; ++i) {
    if (tmp[i] < 0) {  // irregular KMAP key: host-side set
      int len = (int)(off[i + 1] - off[i]);
      uint64_t w[18];
      int cut = mg::canonical_ascii(reinterpret_cast<const uint8_t *>(pool) + off[i], len, w);
      out[i] = c->irregular_ref.count(std::string(reinterpret_cast<const char *>(w), (size_t)cut)) ? 1 : 0;
    } else {
      out[i] = (uint8_t)tmp[i];
    }
  }
  return MG_OK;
}

extern "C" int mg_get_counts(mg_ctx *c, const char *pool, const uint64_t *off, const uint8_t *is_ref, uint64_t n,
                             int32_t *out) {
  if (!c || !off || !is_ref || !out) return set_err(MG_ERR_ARG, "NULL argument");
  if (n == 0) return MG_OK;
  return lookup_common(c, pool, off, is_ref, n, 0, 0, out);
}

// flags the k-mers of allele slot 0 of every variant (they are looked up in ref_bf, main.cpp:167-170)
__global__ void __launch_bounds__(256) k_mark_ref(const uint64_t *__restrict__ var_allele_off,
                                                 const uint64_t *__restrict__ allele_sig_off,
                                                 const uint64_t *__restrict__ sig_kmer_off, uint64_t n_variants,
                                                 uint8_t *__restrict__ flags) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_variants) return;
  uint64_t a0 = var_allele_off[v];
  if (var_allele_off[v + 1] == a0) return;
  for (uint64_t s = allele_sig_off[a0]; s < allele_sig_off[a0 + 1]; ++s)
    for (uint64_t q = sig_kmer_off[s]; q < sig_kmer_off[s + 1]; ++q) flags[q] = 1;
}

// all pointers of in/out are DEVICE pointers here; scratch (k-mer flags + weights) is library-owned
static int genotype_on_device(mg_ctx *c, const mg_variant_batch *in, const mg_genotype_out *out,
                              const mg_batch_dims *dm, float error_rate, int max_coverage, int haploid) {
  cudaStream_t st = c->stream[0];
  uint64_t nv = dm->n_variants, na = dm->n_alleles, nk = dm->n_kmers;
  uint64_t need = (nk ? nk : 1) * 5;
  if (c->geno_scratch_bytes < need) {
    cudaFree(c->geno_scratch);
    c->geno_scratch = nullptr;
    c->geno_scratch_bytes = 0;
    CU(cudaMalloc(&c->geno_scratch, need));
    c->geno_scratch_bytes = need;
  }
  int32_t *d_w = reinterpret_cast<int32_t *>(c->geno_scratch);
  uint8_t *d_flags = reinterpret_cast<uint8_t *>(c->geno_scratch) + (nk ? nk : 1) * 4;
  for (int i = 0; i < 4; ++i)
    if (!c->ge[i]) CU(cudaEventCreate(&c->ge[i]));
  CU(cudaEventRecord(c->ge[0], st));
  if (nk) {
    CU(cudaMemsetAsync(d_flags, 0, nk, st));
    c->launches++;
    k_mark_ref<<<grid_for(nv, 256), 256, 0, st>>>(in->var_allele_off, in->allele_sig_off, in->sig_kmer_off, nv, d_flags);
    CU(cudaGetLastError());
    c->launches++;
    k_lookup<<<grid_for(nk, 128), 128, 0, st>>>(reinterpret_cast<const uint8_t *>(in->pool), in->kmer_off, d_flags, nk,
                                                c->view(), 0, 0, d_w, c->d_scalars);
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(c->ge[1], st));
  c->launches++;
  k_coverage<<<grid_for(na, 128), 128, 0, st>>>(d_w, in->sig_kmer_off, in->allele_sig_off, na, out->cov);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ge[2], st));
  c->launches++;
  k_genotype<<<grid_for(nv, 128), 128, 0, st>>>(out->cov, in->freq, in->var_allele_off, out->lik_off, nv, error_rate,
                                               max_coverage, haploid, out->lik, out->n_gts, out->status, out->best_gt,
                                               out->gq);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ge[3], st));
  return MG_OK;
}

extern "C" int mg_genotype_device(mg_ctx *c, const mg_variant_batch *in, const mg_genotype_out *out,
                                  const mg_batch_dims *dims, float error_rate, int max_coverage, int haploid) {
  if (!c || !in || !out || !dims) return set_err(MG_ERR_ARG, "NULL argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "mg_genotype_device before mg_finalize_alt");
  if (dims->n_variants == 0) return MG_OK;
  if (!in->var_allele_off || !in->allele_sig_off || !in->sig_kmer_off || !in->kmer_off || !in->freq || !out->cov ||
      !out->n_gts || !out->status || !out->best_gt || !out->gq || !out->lik_off || !out->lik)
    return set_err(MG_ERR_ARG, "NULL array in batch");
  CU(cudaSetDevice(c->device));
  return genotype_on_device(c, in, out, dims, error_rate, max_coverage, haploid);
}

extern "C" int mg_genotype(mg_ctx *c, const mg_variant_batch *in, const mg_genotype_out *out, float error_rate,
                           int max_coverage, int haploid) {
  if (!c || !in || !out) return set_err(MG_ERR_ARG, "NULL argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "mg_genotype before mg_finalize_alt");
  uint64_t nv = in->n_variants;
  if (nv == 0) return MG_OK;
  if (!in->var_allele_off || !in->allele_sig_off || !in->sig_kmer_off || !in->kmer_off || !in->freq || !out->cov ||
      !out->n_gts || !out->status || !out->best_gt || !out->gq || !out->lik_off)
    return set_err(MG_ERR_ARG, "NULL array in batch");
  CU(cudaSetDevice(c->device));
  int rc = mg_sync(c);
  if (rc) return rc;
  mg_batch_dims dm;
  dm.n_variants = nv;
  dm.n_alleles = in->var_allele_off[nv];
  dm.n_sigs = in->allele_sig_off[dm.n_alleles];
  dm.n_kmers = in->sig_kmer_off[dm.n_sigs];
  uint64_t na = dm.n_alleles, ns = dm.n_sigs, nk = dm.n_kmers, nl = out->lik_off[nv];
  uint64_t pool_bytes = in->kmer_off[nk];
  cudaStream_t st = c->stream[0];
  // one device arena for the whole batch
  auto al = [](uint64_t x) { return (x + 255) & ~255ull; };
  uint64_t o_vao = 0, o_aso = o_vao + al((nv + 1) * 8), o_sko = o_aso + al((na + 1) * 8),
           o_ko = o_sko + al((ns + 1) * 8), o_lo = o_ko + al((nk + 1) * 8), o_freq = o_lo + al((nv + 1) * 8),
           o_pool = o_freq + al(na * 4), o_cov = o_pool + al(pool_bytes), o_i32 = o_cov + al(na * 4),
           o_lik = o_i32 + al(nv * 16), total = o_lik + al(nl * 8) + 256;
  uint8_t *d = nullptr;
  CU(cudaMalloc(&d, total));
  struct Guard {
    uint8_t *p;
    ~Guard() { cudaFree(p); }
  } guard{d};
  CU(cudaMemcpyAsync(d + o_vao, in->var_allele_off, (nv + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_aso, in->allele_sig_off, (na + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_sko, in->sig_kmer_off, (ns + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_ko, in->kmer_off, (nk + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_lo, out->lik_off, (nv + 1) * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_freq, in->freq, na * 4, cudaMemcpyHostToDevice, st));
  if (pool_bytes) CU(cudaMemcpyAsync(d + o_pool, in->pool, pool_bytes, cudaMemcpyHostToDevice, st));
  mg_variant_batch din;
  din.n_variants = nv;
  din.var_allele_off = reinterpret_cast<uint64_t *>(d + o_vao);
  din.allele_sig_off = reinterpret_cast<uint64_t *>(d + o_aso);
  din.sig_kmer_off = reinterpret_cast<uint64_t *>(d + o_sko);
  din.kmer_off = reinterpret_cast<uint64_t *>(d + o_ko);
  din.pool = reinterpret_cast<const char *>(d + o_pool);
  din.freq = reinterpret_cast<float *>(d + o_freq);
  mg_genotype_out dout;
  int32_t *i32 = reinterpret_cast<int32_t *>(d + o_i32);
  dout.cov = reinterpret_cast<uint32_t *>(d + o_cov);
  dout.n_gts = i32;
  dout.status = i32 + nv;
  dout.best_gt = i32 + 2 * nv;
  dout.gq = i32 + 3 * nv;
  dout.lik_off = reinterpret_cast<uint64_t *>(d + o_lo);
  dout.lik = reinterpret_cast<double *>(d + o_lik);
  rc = genotype_on_device(c, &din, &dout, &dm, error_rate, max_coverage, haploid);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out->cov, dout.cov, na * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out->n_gts, dout.n_gts, nv * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out->status, dout.status, nv * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out->best_gt, dout.best_gt, nv * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out->gq, dout.gq, nv * 4, cudaMemcpyDeviceToHost, st));
  if (out->lik && nl) CU(cudaMemcpyAsync(out->lik, dout.lik, nl * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return check_too_long(c);
}

extern "C" int mg_bf_popcount(mg_ctx *c, int which, uint64_t *ones) {
  if (!c || !ones || which < 0 || which > 1) return set_err(MG_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  int rc = mg_sync(c);
  if (rc) return rc;
  uint32_t *blk = nullptr;
  CU(cudaMalloc(&blk, c->n_blocks * 4));
  CU(cudaMemset(c->d_scalars + 2, 0, 8));
  c->launches++;
  k_block_popc<<<grid_for(c->n_blocks, 256), 256, 0, c->stream[0]>>>(which ? c->ctx_words : c->bf_words, c->n_blocks,
                                                                    c->n_words32, blk, c->d_scalars + 2);
  CU(cudaGetLastError());
  unsigned long long v = 0;
  CU(cudaMemcpyAsync(&v, c->d_scalars + 2, 8, cudaMemcpyDeviceToHost, c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[0]));
  cudaFree(blk);
  *ones = v;
  return MG_OK;
}

extern "C" int mg_bf_download_bits(mg_ctx *c, int which, uint64_t *words, uint64_t n_words) {
  if (!c || !words || which < 0 || which > 1) return set_err(MG_ERR_ARG, "bad argument");
  if (n_words * 2 > c->n_words32) return set_err(MG_ERR_ARG, "n_words exceeds the filter");
  CU(cudaSetDevice(c->device));
  int rc = mg_sync(c);
  if (rc) return rc;
  CU(cudaMemcpy(words, which ? c->ctx_words : c->bf_words, n_words * 8, cudaMemcpyDeviceToHost));
  return MG_OK;
}

extern "C" int mg_bf_download_counts(mg_ctx *c, uint16_t *counts, uint64_t n) {
  if (!c || (!counts && n)) return set_err(MG_ERR_ARG, "bad argument");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "no counters before mg_finalize_alt");
  if (n > c->bf_ones) return set_err(MG_ERR_ARG, "n exceeds popcount");
  CU(cudaSetDevice(c->device));
  int rc = mg_sync(c);
  if (rc) return rc;
  std::vector<uint32_t> tmp(n ? n : 1);
  CU(cudaMemcpy(tmp.data(), c->bf_counts, n * 4, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < n; ++i) counts[i] = (uint16_t)tmp[i];  // uint16 wrap-around of int_vector<16>
  return MG_OK;
}

extern "C" int mg_kmap_size(mg_ctx *c, uint64_t *n) {
  if (!c || !n) return set_err(MG_ERR_ARG, "bad argument");
  *n = c->tab_n + c->irregular_ref.size();
  return MG_OK;
}

extern "C" int mg_counter_buffers(mg_ctx *c, void **d_bf_counts, uint64_t *n_bf, void **d_ref_counts,
                                  uint64_t *n_ref) {
  if (!c) return set_err(MG_ERR_ARG, "NULL ctx");
  if (!c->alt_final) return set_err(MG_ERR_STATE, "no counters before mg_finalize_alt");
  if (d_bf_counts) *d_bf_counts = c->bf_counts;
  if (n_bf) *n_bf = c->bf_ones;
  if (d_ref_counts) *d_ref_counts = c->tab_counts;
  if (n_ref) *n_ref = 1ull << c->tab_log2;
  return MG_OK;
}

extern "C" int mg_add_signatures_packed(mg_ctx *c, const uint64_t *lohi, const uint8_t *is_ref, uint64_t n) {
  if (!c || ((!lohi || !is_ref) && n)) return set_err(MG_ERR_ARG, "NULL argument");
  if (c->alt_final) return set_err(MG_ERR_STATE, "mg_add_signatures_packed after mg_finalize_alt");
  if (n == 0) return MG_OK;
  CU(cudaSetDevice(c->device));
  uint64_t n_ref = 0;
  for (uint64_t i = 0; i < n; ++i) n_ref += is_ref[i] != 0;
  int rc = tab_reserve(c, n_ref);
  if (rc) return rc;
  void *d_k = nullptr;
  uint8_t *d_f = nullptr;
  CU(cudaMalloc(&d_k, n * 16));
  CU(cudaMalloc(&d_f, n));
  CU(cudaMemcpyAsync(d_k, lohi, n * 16, cudaMemcpyHostToDevice, c->stream[0]));
  CU(cudaMemcpyAsync(d_f, is_ref, n, cudaMemcpyHostToDevice, c->stream[0]));
  CU(cudaMemsetAsync(c->d_scalars, 0, 2 * sizeof(unsigned long long), c->stream[0]));
  c->launches++;
  k_add_packed<<<grid_for(n, 256), 256, 0, c->stream[0]>>>(reinterpret_cast<const uint4 *>(d_k), d_f, n, c->view(),
                                                          c->bf_words, c->tab_keys, c->d_scalars);
  CU(cudaGetLastError());
  unsigned long long added = 0;
  CU(cudaMemcpyAsync(&added, c->d_scalars, 8, cudaMemcpyDeviceToHost, c->stream[0]));
  CU(cudaStreamSynchronize(c->stream[0]));
  c->tab_n += added;
  cudaFree(d_k);
  cudaFree(d_f);
  return MG_OK;
}

// ---- timing on the library's own streams (torch.cuda.Event cannot see them) ----
extern "C" int mg_event_record(mg_ctx *c, int idx) {
  if (!c || idx < 0 || idx >= 64) return set_err(MG_ERR_ARG, "bad event index");
  CU(cudaSetDevice(c->device));
  if (!c->tj) CU(cudaEventCreateWithFlags(&c->tj, cudaEventDisableTiming));
  if (!c->evs[idx]) CU(cudaEventCreate(&c->evs[idx]));
  // the event follows everything enqueued so far on both streams, and precedes what comes next
  CU(cudaEventRecord(c->tj, c->stream[1]));
  CU(cudaStreamWaitEvent(c->stream[0], c->tj, 0));
  CU(cudaEventRecord(c->evs[idx], c->stream[0]));
  CU(cudaStreamWaitEvent(c->stream[1], c->evs[idx], 0));
  return MG_OK;
}
extern "C" int mg_event_elapsed_ms(mg_ctx *c, int a, int b, float *ms) {
  if (!c || !ms || a < 0 || b < 0 || a >= 64 || b >= 64 || !c->evs[a] || !c->evs[b])
    return set_err(MG_ERR_ARG, "bad event index");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->evs[b]));
  CU(cudaEventElapsedTime(ms, c->evs[a], c->evs[b]));
  return MG_OK;
}
extern "C" int mg_genotype_kernel_ms(mg_ctx *c, float *ms3) {
  if (!c || !ms3 || !c->ge[3]) return set_err(MG_ERR_ARG, "no mg_genotype call to report");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ge[3]));
  for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(&ms3[i], c->ge[i], c->ge[i + 1]));
  return MG_OK;
}
extern "C" int mg_launch_count(mg_ctx *c, uint64_t *n) {
  if (!c || !n) return set_err(MG_ERR_ARG, "bad argument");
  *n = c->launches;
  return MG_OK;
}

// measured ceilings on `device`: mode 0 / 2 / 3 = independent uniformly random reads of aligned 32 / 64 /
// 128-byte units over `bytes` of HBM, mode 1 = streaming 16-byte reads.  Best of `reps`, in GB/s of useful bytes.
extern "C" int mg_diag_bandwidth(int device, int mode, uint64_t bytes, int reps, double *gbs) {
  if (!gbs || bytes < (1ull << 20) || reps < 1) return set_err(MG_ERR_ARG, "bad argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  {
    size_t gran = 32;  // same L2 fetch policy as the contexts use (see mg_create)
    if (const char *e = getenv("MG_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(e);
    if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  uint32_t *buf = nullptr, *sink = nullptr;
  CU(cudaMalloc(&buf, bytes));
  CU(cudaMalloc(&sink, 4));
  CU(cudaMemset(buf, 0, bytes));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  const int grid = prop.multiProcessorCount * 8;
  const uint64_t per_thread = 256;
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    CU(cudaEventRecord(a));
    const int gran = mode == 0 ? 1 : mode == 2 ? 2 : 4;
    if (mode != 1)
      k_diag_random<<<grid * 4, 256>>>(buf, bytes / (32 * (uint64_t)gran), gran, per_thread, sink);
    else
      k_diag_stream<<<grid, 256>>>(reinterpret_cast<const uint4 *>(buf), bytes / 16, sink);
    CU(cudaGetLastError());
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    double useful = mode != 1 ? (double)grid * 4 * 256 * (double)per_thread * 32.0 * gran : (double)bytes;
    double g = useful / (ms * 1e-3) / 1e9;
    if (r > 0 && g > best) best = g;  // first repetition is the warm-up
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(buf);
  cudaFree(sink);
  *gbs = best;
  return MG_OK;
}

extern "C" int mg_host_alloc(void **p, size_t bytes) {
  if (!p) return set_err(MG_ERR_ARG, "NULL argument");
  CU(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return MG_OK;
}
extern "C" int mg_host_free(void *p) {
  CU(cudaFreeHost(p));
  return MG_OK;
}

// ---------------------------------------------------------------------------
// host-side self tests of the __host__ __device__ helpers (CPU-only test-suite)
// ---------------------------------------------------------------------------
extern "C" uint64_t mg_selftest_hash_packed(uint64_t lo, uint64_t hi, int k, uint64_t *canon_lo, uint64_t *canon_hi) {
  u128 x = {lo, hi}, c;
  uint64_t h = mg::canon_hash_rt(x, k, &c);
  if (canon_lo) *canon_lo = c.lo;
  if (canon_hi) *canon_hi = c.hi;
  return h;
}
extern "C" uint64_t mg_selftest_hash_packed_k35(uint64_t lo, uint64_t hi) {
  u128 x = {lo, hi}, c;
  return mg::canon_hash<35>(x, &c);
}
extern "C" uint64_t mg_selftest_hash_packed_k43(uint64_t lo, uint64_t hi) {
  u128 x = {lo, hi}, c;
  return mg::canon_hash<43>(x, &c);
}
extern "C" uint64_t mg_selftest_hash_ascii(const char *s, int len) {
  return mg::hash_ascii(reinterpret_cast<const uint8_t *>(s), len);
}
extern "C" float mg_selftest_logf(float x) { return mg::glibc_logf(x); }
extern "C" int mg_selftest_genotype(const uint32_t *cov, const float *freq, int n_alleles, float error_rate,
                                    int max_cov, int haploid, double *lik, int *status, int *best_gt, int *gq) {
  return genotype_one(cov, freq, n_alleles, error_rate, max_cov, haploid != 0, lik, status, best_gt, gq);
}
