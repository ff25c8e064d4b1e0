#!/usr/bin/env python
"""bench.py -- throughput of the MALVA genotyping hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload wg|small]

Metric (BASELINE.json): sample k-mers/sec through the Bloom-filter + signature count (headline `value`),
with variants genotyped/sec reported beside it.  One "step" = one pass of the call-side hot path over one
batch of synthetic input: scan B sample 43-mers (K1) and genotype the proportional share V of variants
(K4 + K5), against index structures of whole-genome size (BASELINE config 3: 4 GiB filters, ~1e8 alt bits,
~1e8 ref keys).  `value` uses inputs resident in HBM; `e2e` goes through the host-buffer C-ABI calls with
the H2D / D2H copies inside the timed region.

N > 1 (torchrun, one process per GPU): replicate-and-reduce: every rank holds the full index,
scans its own share of the sample stream, and the two counter arrays are sum-reduced to rank 0 with NCCL
inside the timed region.  Weak scaling: B k-mers per rank per step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K, REF_K = 35, 43
SEED = 20261018 + 3  # seed = 20261018 + cfg

WORKLOADS = {
    # BASELINE.json configs[3]: synthetic whole-genome index structures on one GPU (4 GiB filters)
    "wg": dict(bf_bits=1 << 35, n_alt=100_000_000, n_ref=100_000_000, ref_bases=250_000_000, batch=1 << 27,
               variants=2_850_000, name="synthetic whole-genome 30x (cfg[3]): 2^35-bit filters, 1e8 alt + 1e8 ref "
                                        "signature k-mers; step = 2^27 sample 43-mers + 2.85e6 variants"),
    # small twin for quick checks (not a bench line)
    "small": dict(bf_bits=1 << 30, n_alt=2_000_000, n_ref=2_000_000, ref_bases=5_000_000, batch=1 << 22,
                  variants=90_000, name="small twin (not a bench line)"),
}
HIT_REF, HIT_ALT = 0.03, 0.007  # expected per-k-mer hit rates of ref_bf / bf
ALGO_BYTES_PER_KMER = 84         # 20 B streamed + 2 random 32 B sectors


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic data (torch is used for device memory and random numbers only)
# ------------------------------------------------------------------------------------------------
def rand_kmers(torch, n, k, gen, dev):
    """n random k-mers as int64 [n, 2] = (lo, hi) two's-complement images of the u64 words."""
    lo = (torch.randint(0, 1 << 32, (n,), dtype=torch.int64, generator=gen, device=dev) << 32) | \
        torch.randint(0, 1 << 32, (n,), dtype=torch.int64, generator=gen, device=dev)
    bits_hi = 2 * k - 64
    if bits_hi > 0:
        hi = torch.randint(0, 1 << bits_hi, (n,), dtype=torch.int64, generator=gen, device=dev)
    else:
        hi = torch.zeros(n, dtype=torch.int64, device=dev)
        lo = lo & ((1 << (2 * k)) - 1)
    return torch.stack([lo, hi], dim=1).contiguous()


def embed(torch, sig, gen, dev):
    """35-mer words -> 43-mer words with 4 random flanking bases on each side."""
    n = sig.shape[0]
    lo, hi = sig[:, 0], sig[:, 1]
    fl = torch.randint(0, 256, (n, 2), dtype=torch.int64, generator=gen, device=dev)
    nlo = (lo << 8) | fl[:, 1]
    nhi = (hi << 8) | ((lo >> 56) & 0xFF) | (fl[:, 0] << 14)
    return torch.stack([nlo, nhi], dim=1)


def make_sample_batch(torch, n, alt, ref, gen, dev):
    x = rand_kmers(torch, n, REF_K, gen, dev)
    n_ref, n_alt = int(n * HIT_REF), int(n * HIT_ALT)
    pos = torch.randperm(n, generator=gen, device=dev)[: n_ref + n_alt]
    ri = torch.randint(0, ref.shape[0], (n_ref,), generator=gen, device=dev)
    ai = torch.randint(0, alt.shape[0], (n_alt,), generator=gen, device=dev)
    x[pos[:n_ref]] = embed(torch, ref[ri], gen, dev)
    x[pos[n_ref:]] = embed(torch, alt[ai], gen, dev)
    counts = torch.randint(2, 256, (n,), dtype=torch.int32, generator=gen, device=dev)
    return x.contiguous(), counts


def make_kmc_records(torch, x, counts, p=7):
    """A batch of 43-mer words -> (sorted raw .kmc_suf records uint8 [n*10], prefix LUT int64 [4^p]) as KMC lays them out."""
    lo, hi = x[:, 0], x[:, 1]
    o1 = torch.argsort(lo ^ (-(1 << 63)), stable=True)          # unsigned order of the low word
    o2 = torch.argsort(hi[o1], stable=True)
    order = o1[o2]
    lo, hi, c = lo[order], hi[order], counts[order]
    suf_syms = REF_K - p                                         # 36 symbols = 72 bits = 9 bytes
    prefix = hi >> (2 * suf_syms - 64)
    lut = torch.zeros(4 ** p, dtype=torch.int64, device=x.device)
    lut[1:] = torch.cumsum(torch.bincount(prefix, minlength=4 ** p), 0)[:-1]
    rec = torch.empty((x.shape[0], 10), dtype=torch.uint8, device=x.device)
    rec[:, 0] = (hi & 0xFF).to(torch.uint8)
    for j in range(8):
        rec[:, 1 + j] = ((lo >> (56 - 8 * j)) & 0xFF).to(torch.uint8)
    rec[:, 9] = (c & 0xFF).to(torch.uint8)
    return rec.reshape(-1), lut


def kmers_to_ascii(torch, sig, k):
    """[n,2] int64 words -> uint8 [n,k] ASCII."""
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=sig.device)
    cols = []
    for j in range(k):
        sh = 2 * (k - 1 - j)
        code = ((sig[:, 1] >> (sh - 64)) if sh >= 64 else (sig[:, 0] >> sh)) & 3
        cols.append(lut[code])
    return torch.stack(cols, dim=1).contiguous()


def make_variant_batch(torch, nv, alt, ref, gen, dev):
    """CSR of nv variants: 97% biallelic, 3% with 2-3 ALTs; one signature per allele, 1 k-mer (85%) or 2-3."""
    g = np.random.default_rng(SEED + 7)
    n_all = np.where(g.random(nv) < 0.97, 2, g.integers(3, 5, nv)).astype(np.int64)
    vao = np.zeros(nv + 1, np.uint64)
    vao[1:] = np.cumsum(n_all)
    na = int(vao[-1])
    aso = np.arange(na + 1, dtype=np.uint64)              # one signature per allele slot
    n_k = np.where(g.random(na) < 0.85, 1, g.integers(2, 4, na)).astype(np.int64)
    sko = np.zeros(na + 1, np.uint64)
    sko[1:] = np.cumsum(n_k)
    nk = int(sko[-1])
    is_ref_allele = np.zeros(na, bool)
    is_ref_allele[vao[:-1].astype(np.int64)] = True
    kmer_is_ref = np.repeat(is_ref_allele, n_k)
    t_is_ref = torch.from_numpy(kmer_is_ref).to(dev)
    ri = torch.randint(0, ref.shape[0], (nk,), generator=gen, device=dev)
    ai = torch.randint(0, alt.shape[0], (nk,), generator=gen, device=dev)
    words = torch.where(t_is_ref[:, None], ref[ri], alt[ai])
    miss = torch.rand(nk, generator=gen, device=dev) < 0.3   # k-mers the sample does not support
    words = torch.where(miss[:, None], rand_kmers(torch, nk, K, gen, dev), words)
    pool = kmers_to_ascii(torch, words, K).reshape(-1)
    koff = np.arange(nk + 1, dtype=np.uint64) * K
    af = (g.random(na) * 0.3).astype(np.float32)
    freq = af.copy()
    starts = vao[:-1].astype(np.int64)
    sums = np.add.reduceat(af.astype(np.float64), starts) - af[starts]
    freq[starts] = np.maximum(1.0 - sums, 0).astype(np.float32)
    lik_slots = n_all * (n_all + 1) // 2
    lo = np.zeros(nv + 1, np.uint64)
    lo[1:] = np.cumsum(lik_slots)
    return dict(vao=vao, aso=aso, sko=sko, koff=koff, pool=pool, freq=freq, lik_off=lo, dims=(nv, na, na, nk, nk * K))


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi SM clock + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append((time.time(), l.strip()))

    def mark(self):
        """start of the timed region: earlier samples are dropped"""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, l in self.lines:
            if ts < t_mark:
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class DevArray:
    """A library-owned device buffer seen as a torch tensor (for the NCCL reduce)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own BF + KMAP classes (oracle/_ref), 1 thread
# ------------------------------------------------------------------------------------------------
def cpu_reference_scan(bf_bits, alt_np, ref_np, sample_np, counts_np, max_seconds=25.0):
    """Times the reference's scan loop (main.cpp:487-500) on host cores over a bounded sample.
    Returns (kmers_per_sec, n_done, kind)."""
    from oracle import pyoracle

    u64p, u32p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    if pyoracle.have_ref():
        L, kind = pyoracle.ref(), "reference"
        bf, ctx, km = L.ref_bf_new(bf_bits), L.ref_bf_new(bf_bits), L.ref_kmap_new()
        flags = np.concatenate([np.zeros(len(alt_np), np.uint8), np.ones(len(ref_np), np.uint8)])
        keys = np.ascontiguousarray(np.concatenate([alt_np, ref_np]))
        L.ref_add_packed(bf, km, keys.ctypes.data_as(u64p), flags.ctypes.data_as(u8p), len(flags), K)
        L.ref_bf_switch_mode(bf)
        L.ref_bf_switch_mode(ctx)
        scan = lambda a, c, n: L.ref_scan_packed(bf, ctx, km, a.ctypes.data_as(u64p), c.ctypes.data_as(u32p), n, K, REF_K)
        free = lambda: (L.ref_bf_free(bf), L.ref_bf_free(ctx), L.ref_kmap_free(km))
    else:
        L, kind = pyoracle.oracle(), "port"
        bf, ctx, km = L.mo_bf_new(bf_bits), L.mo_bf_new(bf_bits), L.mo_kmap_new()
        from malva_b200.kmc import packed_to_strings, KMER_DTYPE
        for arr, is_ref in ((alt_np, 0), (ref_np, 1)):
            for s in packed_to_strings(arr.view(KMER_DTYPE).reshape(-1), K):
                (L.mo_kmap_add_key(km, s.encode()) if is_ref else L.mo_bf_add_key(bf, s.encode()))
        L.mo_bf_switch_mode(bf)
        L.mo_bf_switch_mode(ctx)
        scan = lambda a, c, n: L.mo_scan_packed(bf, ctx, km, a.ctypes.data_as(u64p), c.ctypes.data_as(u32p), n, K, REF_K)
        free = lambda: (L.mo_bf_free(bf), L.mo_bf_free(ctx), L.mo_kmap_free(km))
    chunk, done, t_used = 200_000, 0, 0.0
    while done + chunk <= len(counts_np) and t_used < max_seconds:
        a = np.ascontiguousarray(sample_np[done:done + chunk])
        c = np.ascontiguousarray(counts_np[done:done + chunk])
        t0 = time.perf_counter()
        scan(a, c, chunk)
        t_used += time.perf_counter() - t0
        done += chunk
    free()
    return done / t_used, done, kind


def cpu_inputs(torch, wl, gen, dev_cpu_only=False):
    """Bounded CPU-side sample of the workload: 2e6 alt + 2e6 ref signature k-mers, 4e6 sample k-mers."""
    g = np.random.default_rng(SEED)
    def rk(n, k):
        lo = g.integers(0, 1 << 63, n, dtype=np.uint64) * 2 + g.integers(0, 2, n, dtype=np.uint64)
        hi = g.integers(0, 1 << (2 * k - 64), n, dtype=np.uint64)
        return np.stack([lo, hi], axis=1)
    n_sig, n_s = 2_000_000, 4_000_000
    alt, ref = rk(n_sig, K), rk(n_sig, K)
    smp = rk(n_s, REF_K)
    def emb(sig):
        fl = g.integers(0, 256, (len(sig), 2), dtype=np.uint64)
        lo = (sig[:, 0] << np.uint64(8)) | fl[:, 1]
        hi = (sig[:, 1] << np.uint64(8)) | (sig[:, 0] >> np.uint64(56)) | (fl[:, 0] << np.uint64(14))
        return np.stack([lo, hi], axis=1)
    n_r, n_a = int(n_s * HIT_REF), int(n_s * HIT_ALT)
    pos = g.permutation(n_s)[: n_r + n_a]
    smp[pos[:n_r]] = emb(ref[g.integers(0, n_sig, n_r)])
    smp[pos[n_r:]] = emb(alt[g.integers(0, n_sig, n_a)])
    counts = g.integers(2, 256, n_s).astype(np.uint32)
    return alt, ref, smp, counts


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    alt, ref, smp, counts = cpu_inputs(None, wl, None)
    per_step = 400_000
    need = per_step * (args.steps + args.warmup)
    reps = -(-need // len(counts))
    smp, counts = np.tile(smp, (reps, 1)), np.tile(counts, reps)
    from oracle import pyoracle

    u64p, u32p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    if not pyoracle.have_ref():
        rate, done, kind = cpu_reference_scan(wl["bf_bits"], alt, ref, smp, counts, 20.0)
        ms = per_step / rate * 1e3
    else:
        L, kind = pyoracle.ref(), "reference"
        bf, ctx, km = L.ref_bf_new(wl["bf_bits"]), L.ref_bf_new(wl["bf_bits"]), L.ref_kmap_new()
        flags = np.concatenate([np.zeros(len(alt), np.uint8), np.ones(len(ref), np.uint8)])
        keys = np.ascontiguousarray(np.concatenate([alt, ref]))
        L.ref_add_packed(bf, km, keys.ctypes.data_as(u64p), flags.ctypes.data_as(u8p), len(flags), K)
        L.ref_bf_switch_mode(bf)
        L.ref_bf_switch_mode(ctx)
        t_total = 0.0
        for s in range(args.warmup + args.steps):
            a = np.ascontiguousarray(smp[s * per_step:(s + 1) * per_step])
            c = np.ascontiguousarray(counts[s * per_step:(s + 1) * per_step])
            t0 = time.perf_counter()
            L.ref_scan_packed(bf, ctx, km, a.ctypes.data_as(u64p), c.ctypes.data_as(u32p), per_step, K, REF_K)
            if s >= args.warmup:
                t_total += time.perf_counter() - t0
        ms = t_total / args.steps * 1e3
        rate = per_step / (ms * 1e-3)
    sample = (f"{per_step} sample 43-mers per step against 2^{int(np.log2(wl['bf_bits']))}-bit filters, 2e6 alt bits and "
              "2e6 ref keys (the full workload has 1e8 each: fewer keys flatter the CPU); the reference's own BF/KMAP "
              "classes, single-threaded like the reference (no threads in malva-geno, KMC run with -t1)")
    line = {"impl": "reference", "metric": "sample_kmers_per_sec", "value": rate, "unit": "k-mers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["name"], "k": K, "ref_k": REF_K},
            "cpu_baseline": {"value": rate, "unit": "k-mers/s", "cores": 1, "kind": kind, "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": rate, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from malva_b200 import MalvaGpu
    from malva_b200.api import diag_bandwidth
    from malva_b200.kmc import KMER_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the MALVA hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED)           # identical index on every rank (replicate-and-reduce)
    t_setup = time.time()
    g = MalvaGpu(k=K, ref_k=REF_K, bf_bits=wl["bf_bits"], device=local_rank)
    alt = rand_kmers(torch, wl["n_alt"], K, gen, dev)
    ref = rand_kmers(torch, wl["n_ref"], K, gen, dev)
    chunk = 1 << 24
    for arr, flag in ((alt, 0), (ref, 1)):
        for o in range(0, arr.shape[0], chunk):
            h = arr[o:o + chunk].cpu().numpy().view(np.uint64).reshape(-1).view(KMER_DTYPE)
            g.add_signatures_packed(h, np.full(len(h), flag, np.uint8))
    g.finalize_alt()
    # reference rolling pass (K2) over a synthetic contig that carries some alt signatures, timed once
    rb = torch.randint(0, 4, (wl["ref_bases"],), dtype=torch.uint8, generator=gen, device=dev)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    ref_seq = lut[rb.long()] if wl["ref_bases"] <= 50_000_000 else torch.cat(
        [lut[rb[i:i + 50_000_000].long()] for i in range(0, wl["ref_bases"], 50_000_000)])
    del rb
    n_plant = min(200_000, wl["ref_bases"] // 1000)
    plant = kmers_to_ascii(torch, alt[:n_plant], K)
    # one plant per stride, jittered: the planted windows never overlap, so the scatter below is deterministic
    # and every rank builds the same context filter
    stride = (wl["ref_bases"] - 200) // n_plant
    ppos = 100 + torch.arange(n_plant, device=dev) * stride + \
        torch.randint(0, max(1, stride - K), (n_plant,), generator=gen, device=dev)
    ref_seq[(ppos[:, None] + torch.arange(K, device=dev)[None, :]).reshape(-1)] = plant.reshape(-1)
    ref_host = ref_seq.cpu().numpy().tobytes()
    del ref_seq
    g.scan_reference(ref_host)   # cold call (allocation, first touch of the pageable host buffer)
    g.event_record(0)
    g.scan_reference(ref_host)   # idempotent (bits are only ever set): the warm call is the one timed
    g.event_record(1)
    refpass_ms = g.event_elapsed_ms(0, 1)           # includes the H2D copy of the contig from pageable memory
    refpass_kernel_ms = g.refpass_kernel_ms()       # the rolling-pass kernel alone
    del ref_host
    g.finalize_context()
    pop_alt, pop_ctx, n_keys = g.popcount(0), g.popcount(1), g.kmap_size()
    # sample batches, device resident (2 batches rotate so that no step re-reads the previous step's lines)
    # every rank scans its own share of the stream (--verify: the same share, so that the reduced counters must be
    # exactly world x one rank's)
    gen.manual_seed(SEED + 100 + (0 if args.verify else rank))
    B = wl["batch"]
    batches = [make_sample_batch(torch, B, alt, ref, gen, dev) for _ in range(2)]
    vb = make_variant_batch(torch, wl["variants"], alt, ref, gen, dev)
    nv, na, ns, nk = vb["dims"][:4]
    d = {k2: torch.from_numpy(vb[k1]).to(dev) for k1, k2 in (("vao", "var_allele_off"), ("aso", "allele_sig_off"),
                                                              ("sko", "sig_kmer_off"), ("koff", "kmer_off"),
                                                              ("lik_off", "lik_off"), ("freq", "freq"))}
    d["pool"] = vb["pool"]
    nl = int(vb["lik_off"][-1])
    d["cov"] = torch.zeros(na, dtype=torch.int32, device=dev)
    for nme in ("n_gts", "status", "best_gt", "gq"):
        d[nme] = torch.zeros(nv, dtype=torch.int32, device=dev)
    d["lik"] = torch.zeros(nl, dtype=torch.float64, device=dev)
    ptrs = {k2: t.data_ptr() for k2, t in d.items()}
    del alt, ref
    torch.cuda.synchronize()
    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s: bf ones {pop_alt}, context ones {pop_ctx}, ref keys {n_keys}, "
        f"K2 reference pass {wl['ref_bases'] / refpass_kernel_ms / 1e6:.1f} Gbases/s (kernel), "
        f"{wl['ref_bases'] / refpass_ms / 1e6:.1f} Gbases/s incl. H2D")

    # ---- host-side inputs of the e2e leg, all in pinned memory ----
    #  (a) the batch as raw KMC suffix records (+ prefix LUT): what malva-geno call reads from <db>.kmc_suf
    #  (b) the same batch as packed {lo,hi} words + u32 counts (the host-decoded form)
    #  (c) the variant CSR; outputs land in pinned buffers too
    pin = lambda t: t.cpu().pin_memory()
    kmc_rec, kmc_lut = make_kmc_records(torch, batches[0][0], batches[0][1])
    h_rec = pin(kmc_rec)
    kmc_db = dict(lut=kmc_lut.cpu().numpy().astype(np.uint64), lut_prefix_len=7, k=REF_K, counter_size=1, min_count=2,
                  max_count=255)
    del kmc_rec
    hk, hc = pin(batches[0][0]), pin(batches[0][1])
    h_in = {k2: pin(torch.from_numpy(vb[k1])) for k1, k2 in (("vao", "var_allele_off"), ("aso", "allele_sig_off"),
                                                             ("sko", "sig_kmer_off"), ("koff", "kmer_off"),
                                                             ("lik_off", "lik_off"), ("freq", "freq"))}
    h_in["pool"] = pin(vb["pool"])
    h_out = {"cov": torch.zeros(na, dtype=torch.int32).pin_memory(), "lik": torch.zeros(nl, dtype=torch.float64).pin_memory()}
    for nme in ("n_gts", "status", "best_gt", "gq"):
        h_out[nme] = torch.zeros(nv, dtype=torch.int32).pin_memory()
    h_ptrs = {k2: t.data_ptr() for k2, t in {**h_in, **h_out}.items()}
    csr_bytes = sum(t.numel() * t.element_size() for t in h_in.values())

    counters = [torch.as_tensor(DevArray(p, n), device=dev) for p, n in g.counter_buffers() if n] if world > 1 else []

    def step(i):
        kk, cc = batches[i & 1]
        g.scan_sample_kmers_ptr(kk.data_ptr(), cc.data_ptr(), B, device=True)
        g.genotype_device(ptrs, vb["dims"], 0.001, 200, False)

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = g.launch_count()
    sampler.mark()
    g.event_record(2)
    for i in range(args.steps):
        g.event_record(10 + 2 * (i % 8))
        kk, cc = batches[i & 1]
        g.scan_sample_kmers_ptr(kk.data_ptr(), cc.data_ptr(), B, device=True)
        g.event_record(11 + 2 * (i % 8))
        g.genotype_device(ptrs, vb["dims"], 0.001, 200, False)
    before = None
    if world > 1:
        g.sync()  # the library's streams -> torch's stream, then the NCCL sum-reduce of both counter arrays
        if args.verify:
            before = [t.clone() for t in counters]
        for t in counters:
            dist.reduce(t, dst=0)
        torch.cuda.synchronize()
    g.event_record(3)
    region_ms = g.event_elapsed_ms(2, 3)
    clocks = sampler.stop()
    launches = g.launch_count() - launches0
    barrier()
    scan_ms = [g.event_elapsed_ms(10 + 2 * j, 11 + 2 * j) for j in range(min(8, args.steps))]
    geno_ms = g.genotype_kernel_ms()
    tm = torch.tensor([region_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    region_ms = float(tm.item())
    verified = None
    if before is not None and rank == 0:
        # identical replicas (canonical index image) that scanned identical batches: sum over ranks == world x mine
        verified = all(bool(torch.equal(t, b * world)) for t, b in zip(counters, before)) and \
            any(int(b.sum().item()) != 0 for b in before)
        log(f"[verify] reduced counters == {world} x rank 0's own counters: {verified}")
    ms_per_step = region_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C-ABI, H2D and D2H inside the timed region ----
    def e2e_leg(scan):
        steps = 0 if args.no_e2e else max(3, min(args.steps, 6))
        if not steps:
            return None
        scan()
        g.genotype_host(h_ptrs, nv, 0.001, 200, False)
        barrier()
        t0 = time.perf_counter()
        g.event_record(4)
        for _ in range(steps):
            scan()                                            # asynchronous, chunked, double-buffered H2D + K1
            g.genotype_host(h_ptrs, nv, 0.001, 200, False)    # H2D CSR, K4+K5, D2H results; returns when they landed
        g.event_record(5)
        ms = max(g.event_elapsed_ms(4, 5), (time.perf_counter() - t0) * 1e3) / steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g.kmc_open(kmc_db)
    e2e_ms = e2e_leg(lambda: g.scan_kmc_records(h_rec.data_ptr(), 0, B, sync=False))
    e2e_packed_ms = e2e_leg(lambda: g.scan_sample_kmers_ptr(hk.data_ptr(), hc.data_ptr(), B, device=False))
    d2h = 4 * na + 16 * nv + 8 * nl
    h2d = B * 10 + csr_bytes
    h2d_packed = B * 20 + csr_bytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (K1 scan), measured live with CUDA events on its stream ----
    peak, peak_src = measured_peaks()
    k1_ms = float(np.mean(scan_ms))
    achieved = B * ALGO_BYTES_PER_KMER / (k1_ms * 1e-3) / 1e9
    rand_gbs = stream_gbs = line_gbs = None
    if not args.no_diag:
        try:
            rand_gbs = diag_bandwidth(local_rank, 0, 16 << 30, 3)
            line_gbs = diag_bandwidth(local_rank, 4, 16 << 30, 3)
            stream_gbs = diag_bandwidth(local_rank, 1, 16 << 30, 3)
        except Exception as e:  # noqa: BLE001
            log("diag_bandwidth failed:", e)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        alt_c, ref_c, smp_c, cnt_c = cpu_inputs(torch, wl, gen)
        rate, done, kind = cpu_reference_scan(wl["bf_bits"], alt_c, ref_c, smp_c, cnt_c, 20.0)
        cpu = {"value": rate, "unit": "k-mers/s", "cores": 1, "kind": kind,
               "sample": f"{done} sample 43-mers against 2^{int(np.log2(wl['bf_bits']))}-bit filters with 2e6 alt bits + 2e6 "
                         "ref keys (full workload: 1e8 each), the reference's own BF/KMAP classes, 1 thread "
                         "(malva-geno is single-threaded)", "host_cores_available": os.cpu_count()}
    line = {
        "metric": "sample_kmers_per_sec", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "k": K, "ref_k": REF_K, "bf_bits": wl["bf_bits"],
                   "kmers_per_step_per_gpu": B, "variants_per_step_per_gpu": nv,
                   "hit_rates": {"ref_bf": HIT_REF, "bf": HIT_ALT},
                   "l2_policy": "inputs larger than L2: 2.7 GB streamed per step, two batches alternate, probes "
                                "spread over 4 GiB + 4.3 GB structures",
                   "parallelism": f"replicate-and-reduce x{world}" if world > 1 else "1 GPU"},
        "variants_per_sec": world * nv / (sum(geno_ms) * 1e-3),
        "variants_per_sec_note": "K4+K5 kernels only (signature look-ups, coverage, likelihood), device-resident CSR",
        "kernel_ms": {"k1_scan": k1_ms, "k4_lookup": geno_ms[0], "k4_coverage": geno_ms[1], "k5_genotype": geno_ms[2],
                      "k2_reference_pass": refpass_kernel_ms, "k2_reference_pass_incl_h2d": refpass_ms},
        "ref_bases_per_sec": wl["ref_bases"] / (refpass_kernel_ms * 1e-3),
        "ref_bases_per_sec_incl_h2d": wl["ref_bases"] / (refpass_ms * 1e-3),
        "index": {"bf_ones": pop_alt, "context_ones": pop_ctx, "ref_keys": n_keys},
        "e2e": None if e2e_ms is None else {
            "value": world * B / (e2e_ms * 1e-3), "unit": "k-mers/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "variants_per_sec": world * nv / (e2e_ms * 1e-3),
            "h2d_GBps": h2d / (e2e_ms * 1e-3) / 1e9,
            "note": "per step: mg_scan_kmc_records(pinned raw .kmc_suf records, 10 B per 43-mer, decoded on the device) + "
                    "mg_genotype(pinned host CSR) -> pinned host results",
            "packed128": {"value": world * B / (e2e_packed_ms * 1e-3), "ms_per_step": e2e_packed_ms,
                          "h2d_bytes_per_step": h2d_packed,
                          "note": "same step with host-decoded {lo,hi} words + u32 counts (20 B per 43-mer) through "
                                  "mg_scan_sample_kmers"}},
        "gpu_launches": launches,
        **({"verify_reduce_exact": verified} if verified is not None else {}),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_scan<35,43>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER,
                     "measured_random_32B_sector_GBps": rand_gbs, "measured_random_128B_line_GBps": line_gbs,
                     "measured_stream_read_GBps": stream_gbs,
                     "frac_of_random_sector_ceiling": (achieved / rand_gbs) if rand_gbs else None,
                     "lines_per_sec_vs_ceiling": (B / (k1_ms * 1e-3)) / (line_gbs * 1e9 / 128) if line_gbs else None},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wg", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--no-diag", action="store_true", help="skip the bandwidth microbenchmarks (profiling runs)")
    ap.add_argument("--verify", action="store_true",
                    help="N > 1: every rank scans the same batches; checks that the NCCL-reduced counters are exactly "
                         "N x one rank's (replicas are identical, the reduce is exact)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, local_rank, world)


if __name__ == "__main__":
    main()
