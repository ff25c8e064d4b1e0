#!/usr/bin/env python
"""Throughput of K6 (canonical 43-mer counting, mg_count_*) on synthetic reads: a random genome sampled at a given
depth with 150-base reads, 0.5 % substitution errors, both strands.  Host memory in (pinned by torch), KMC-ordered
(k-mer, count) arrays out.   python profiles/count_bench.py [genome_Mbp=50] [depth=20] > gpurun_out/count_bench.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from malva_b200 import KmerCounter  # noqa: E402


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50.0
    depth = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
    g = np.random.default_rng(20261018)
    L, rl = int(mbp * 1e6), 150
    genome = g.integers(0, 4, L, dtype=np.uint8)
    n_reads = int(L * depth / rl)
    out = {"genome_bases": L, "reads": n_reads, "read_len": rl, "k": 43}
    t_gen = time.time()
    chunks, per = [], 2_000_000
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for o in range(0, n_reads, per):
        m = min(per, n_reads - o)
        pos = g.integers(0, L - rl, m)
        r = genome[pos[:, None] + np.arange(rl)[None, :]]
        err = g.random((m, rl)) < 0.005
        r = np.where(err, (r + g.integers(1, 4, (m, rl), dtype=np.uint8)) & 3, r)
        rc = g.random(m) < 0.5
        r[rc] = (3 - r[rc])[:, ::-1]
        txt = np.empty((m, rl + 1), np.uint8)
        txt[:, :rl] = lut[r]
        txt[:, rl] = 10
        chunks.append(txt.tobytes())
    out["generate_s"] = round(time.time() - t_gen, 1)
    total = sum(len(c) for c in chunks)
    c = KmerCounter(43)
    c.add(chunks[0][: 1 << 20])        # warm-up (context, first growth)
    c.reset()
    t0 = time.perf_counter()
    for ch in chunks:
        c.add(ch)
    t_add = time.perf_counter() - t0
    st = c.stats()
    t1 = time.perf_counter()
    keys, counts = c.finish(2, 255)
    t_fin = time.perf_counter() - t1
    out.update({"read_bytes": total, "instances": st["instances"], "distinct": st["distinct"], "kept_ci2": int(len(keys)),
                "table_capacity": st["capacity"], "count_s": round(t_add, 3), "count_kernels_s": round(st["kernel_us"] / 1e6, 4),
                "instances_per_s_kernels_only": st["instances"] / max(st["kernel_us"] / 1e6, 1e-9), "finish_sort_download_s": round(t_fin, 3),
                "bases_per_s": total / t_add, "instances_per_s": st["instances"] / t_add,
                "note": "count_s includes the H2D copies of the reads (pageable host memory) and every table growth + rehash; "
                        "one random 32-byte slot access (+ one atomic) per k-mer instance"})
    assert (np.diff(keys["hi"].astype(np.int64)) >= 0).all()
    c.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
