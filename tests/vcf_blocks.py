"""Test-side VCF decoding and var_block grouping (variant.hpp:66-211, main.cpp:309-370, 522-579),
used to drive the GPU path over the reference's bundled example with signatures enumerated by
the reference's own VB::extract_kmers (oracle/_ref hook).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import gzip
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Rec:
    chrom: str
    pos0: int
    vid: str
    ref: str
    alts: List[str]
    qual: str
    freqs: List[float]
    is_present: bool
    has_alts: bool
    gts: List[str] = field(default_factory=list)

    @property
    def ref_size(self):
        return len(self.ref)

    @property
    def min_size(self):
        return min([len(self.ref)] + [len(a) for a in self.alts])


def read_vcf(path: str, freq_key: str = "AF", uniform: bool = False, sample_cols=None):
    """sample_cols: indices of the sample columns to keep, ascending (bcf_hdr_set_samples keeps header order)"""
    op = gzip.open if path.endswith(".gz") else open
    header, recs = [], []
    with op(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("#"):
                header.append(line)
                continue
            c = line.split("\t")
            alts = [a.upper() for a in c[4].split(",") if a != "." and not a.startswith("<")]
            has_alts = len(alts) > 0
            freqs, present = [], True
            if has_alts:
                if uniform:
                    freqs = [float(np.float32(1.0 / (len(alts) + 1)))] * (len(alts) + 1)
                else:
                    af = None
                    for kv in c[7].split(";"):
                        if kv.startswith(freq_key + "="):
                            af = [np.float32(float(x)) for x in kv[len(freq_key) + 1:].split(",")]
                    f = [np.float32(0.0)] + af[:len(alts)]
                    f0 = np.float32(1.0 - sum(float(x) for x in f))   # accumulate(..., 0.0) in double
                    if f0 < 0:
                        f0 = np.float32(0.0)
                    f[0] = f0
                    freqs = [float(x) for x in f]
                present = np.float32(freqs[0]) != np.float32(1.0)
            gts = []
            if has_alts and present and len(c) > 9:
                fmt = c[8].split(":")
                gi = fmt.index("GT")
                cols = c[9:] if sample_cols is None else [c[9 + j] for j in sample_cols]
                gts = [s.split(":")[gi] for s in cols if s != ""]
            recs.append(Rec(c[0], int(c[1]) - 1, c[2], c[3].upper(), alts, c[5], freqs, bool(present), has_alts, gts))
    return header, recs


def near(last: Rec, v: Rec, k: int) -> bool:
    # var_block.hpp:417-423 -- the sum is promoted to float by ceil((float)k/2)
    lhs = np.float32(last.pos0 + last.ref_size - last.min_size - 1) + np.float32(math.ceil(np.float32(k) / 2))
    return bool(lhs >= np.float32(v.pos0))


def blocks(recs: List[Rec], k: int, index_mode: bool):
    """Yield (contig, [Rec]) exactly as index_main (skips !is_present) / call_main group them."""
    cur: List[Rec] = []
    last_name = ""
    for v in recs:
        if last_name == "":
            last_name = v.chrom
        if not v.has_alts or (index_mode and not v.is_present):
            continue
        if not cur:
            cur.append(v)
            continue
        if not near(cur[-1], v, k) or last_name != v.chrom:
            yield last_name, cur
            cur = []
            last_name = v.chrom
        cur.append(v)
    if cur:
        yield last_name, cur


def ref_extract(ref_lib, block: List[Rec], reference: str, k: int, haploid: bool):
    """VB::extract_kmers through the oracle/_ref hook -> nested[v][allele] = [signatures]."""
    lines = []
    for v in block:
        gts = v.gts
        if haploid:
            gts = [g.replace(".", "0").split("|")[0].split("/")[0] for g in gts]
        else:
            gts = [g.replace(".", "0") for g in gts]
        lines.append("\t".join([str(v.pos0), v.ref, ",".join(v.alts), "1" if v.is_present else "0", " ".join(gts)]))
    text = ("\n".join(lines) + "\n").encode()
    cap = 1 << 22
    while True:          # (a dense block of a thousand variants prints tens of MB of signatures)
        buf = C.create_string_buffer(cap)
        n = ref_lib.ref_extract_kmers(text, reference.encode(), k, int(haploid), buf, cap)
        if n >= 0 or cap >= (1 << 30):
            break
        cap *= 4
    assert n >= 0, "ref_extract_kmers output buffer too small"
    nested = [[[] for _ in range(len(v.alts) + 1)] for v in block]
    for l in buf.value.decode().split("\n"):
        if not l:
            continue
        vi, ai, ks = l.split("\t")
        assert int(ai) >= 0
        nested[int(vi)][int(ai)].append(ks.split(","))
    return nested
