"""N > 1 on CPU (gloo, world size 2): the two replicate-and-reduce schemes of DESIGN.md section 7.

Every rank holds a full replica of the index (here: the CPU oracle's), scans its own share of the sample
stream, and the counter arrays are sum-reduced to rank 0.  The reduced state must equal a single-process
scan of the whole stream: bf counters modulo 2^16 (uint16 wrap-around is a ring homomorphism of the
partial sums), ref_bf counts modulo 2^32."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_util as util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(L, seed):
    rng = random.Random(seed)
    k, ref_k, bits = 35, 43, 1 << 16
    genome = util.make_genome(rng, 8000)
    nested, _ = util.synth_signatures(rng, genome, k, 150)
    ks, fl = util.flatten(nested)
    words, packed, counts = util.synth_sample(rng, genome, nested, k, ref_k, 6000, big_counts=True)
    o = util.OracleRun(L, k, ref_k, bits)
    o.add_signatures(ks, fl)
    o.finalize_alt()
    o.scan_reference(genome)
    o.finalize_context()
    return o, ks, packed, counts


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle

    L = pyoracle.oracle()
    o, ks, packed, counts = _build(L, 77)          # identical index on every rank
    lo, hi = rank * len(packed) // world, (rank + 1) * len(packed) // world
    o.scan_sample_kmers(packed[lo:hi].copy(), counts[lo:hi].copy())   # this rank's share of the stream
    bf = torch.from_numpy(o.bf_counts().astype(np.int32))              # u32 partials on the device path
    ref = torch.from_numpy(o.get_counts(ks, [1] * len(ks)).astype(np.int64))
    dist.reduce(bf, dst=0)
    dist.reduce(ref, dst=0)
    if rank == 0:
        np.savez(out_path, bf=(bf.numpy() & 0xFFFF).astype(np.uint16), ref=(ref.numpy() & 0xFFFFFFFF))
    o.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_replicate_and_reduce_world2(oracle_lib, tmp_path):
    out = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    o, ks, packed, counts = _build(oracle_lib, 77)
    o.scan_sample_kmers(packed, counts)
    assert np.array_equal(got["bf"], o.bf_counts())
    assert np.array_equal(got["ref"], o.get_counts(ks, [1] * len(ks)).astype(np.int64) & 0xFFFFFFFF)
    assert o.bf_counts().sum() > 0
    o.close()


def _worker_lookup(rank, world, port, out_path):
    """the per-batch scheme of bench.py at N > 1: every rank looks the variant batch up in its OWN partial counters
    (BF::get_count for alt k-mers, KMAP::get_count for ref k-mers), the results are summed onto rank 0, which masks
    them the way the counters wrap (u16 / u32) before computing coverages"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle

    L = pyoracle.oracle()
    o, ks, packed, counts = _build(L, 78)
    lo, hi = rank * len(packed) // world, (rank + 1) * len(packed) // world
    for _ in range(3):                                                   # (several passes: sums pass 2^16)
        o.scan_sample_kmers(packed[lo:hi].copy(), counts[lo:hi].copy())
    w_alt = torch.from_numpy(o.get_counts(ks, [0] * len(ks)).astype(np.int64))
    w_ref = torch.from_numpy(o.get_counts(ks, [1] * len(ks)).astype(np.int64))
    dist.reduce(w_alt, dst=0)
    dist.reduce(w_ref, dst=0)
    if rank == 0:
        np.savez(out_path, alt=w_alt.numpy() & 0xFFFF, ref=w_ref.numpy() & 0xFFFFFFFF, alt_raw_max=int(w_alt.max()))
    o.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_replicas_sum_lookup_results_world2(oracle_lib, tmp_path):
    out = str(tmp_path / "weights.npz")
    mp.spawn(_worker_lookup, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    o, ks, packed, counts = _build(oracle_lib, 78)
    for _ in range(3):
        o.scan_sample_kmers(packed, counts)
    assert np.array_equal(got["alt"], o.get_counts(ks, [0] * len(ks)).astype(np.int64) & 0xFFFF)
    assert np.array_equal(got["ref"], o.get_counts(ks, [1] * len(ks)).astype(np.int64) & 0xFFFFFFFF)
    assert int(got["alt_raw_max"]) > 0
    o.close()
