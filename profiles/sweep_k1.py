#!/usr/bin/env python
"""Tuning sweep for the sample-scan kernel (K1) on the bench workload: L2 fetch granularity x ILP x CTAs/SM.
Run on a GPU box:  python profiles/sweep_k1.py [wg|small] > gpurun_out/sweep_k1.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from malva_b200 import MalvaGpu  # noqa: E402
from malva_b200.api import diag_bandwidth  # noqa: E402
from malva_b200.kmc import KMER_DTYPE  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "wg"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(bench.SEED)
alt = bench.rand_kmers(torch, wl["n_alt"], bench.K, gen, dev)
ref = bench.rand_kmers(torch, wl["n_ref"], bench.K, gen, dev)
B = wl["batch"]
batches = [bench.make_sample_batch(torch, B, alt, ref, gen, dev) for _ in range(2)]
alt_h = alt.cpu().numpy().view(np.uint64).reshape(-1).view(KMER_DTYPE)
ref_h = ref.cpu().numpy().view(np.uint64).reshape(-1).view(KMER_DTYPE)
del alt, ref

for gran in (32, 64, 128):
    os.environ["MG_L2_FETCH_GRANULARITY"] = str(gran)
    for m, name in ((0, "rand32"), (2, "rand64"), (3, "rand128"), (1, "stream")):
        print(f"l2_fetch={gran} {name} {diag_bandwidth(0, m, 8 << 30, 3):.1f} GB/s", flush=True)
    for ilp, ctas in ((1, 6), (1, 12), (2, 4), (2, 8), (4, 2), (4, 4)):
        os.environ["MG_SCAN_ILP"] = str(ilp)
        os.environ["MG_SCAN_CTAS_PER_SM"] = str(ctas)
        g = MalvaGpu(k=bench.K, ref_k=bench.REF_K, bf_bits=wl["bf_bits"])
        chunk = 1 << 24
        for arr, flag in ((alt_h, 0), (ref_h, 1)):
            for o in range(0, len(arr), chunk):
                g.add_signatures_packed(arr[o:o + chunk], np.full(len(arr[o:o + chunk]), flag, np.uint8))
        g.finalize_alt()
        g.finalize_context()
        for i in range(3):
            g.scan_sample_kmers_ptr(batches[i & 1][0].data_ptr(), batches[i & 1][1].data_ptr(), B, device=True)
        g.sync()
        g.event_record(0)
        n = 10
        for i in range(n):
            g.scan_sample_kmers_ptr(batches[i & 1][0].data_ptr(), batches[i & 1][1].data_ptr(), B, device=True)
        g.event_record(1)
        ms = g.event_elapsed_ms(0, 1) / n
        print(f"l2_fetch={gran} ilp={ilp} ctas_per_sm={ctas}: {ms:.3f} ms/scan  {B / ms / 1e6:.2f} G k-mers/s  "
              f"{B * 84 / ms / 1e6:.0f} GB/s algorithmic", flush=True)
        g.close()
