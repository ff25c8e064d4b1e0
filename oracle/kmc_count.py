"""TEST INFRASTRUCTURE ONLY -- restatement of what `kmc -k<k> -ci<min> -cs<max>` lists (canonical k-mers, windows with
a non-ACGT symbol skipped, counts below min_count dropped, counters saturated), used as the checker of the GPU k-mer
counter (K6) and to make the haploid golden's database (tests/golden/make_golden.py).  What it restates is pinned by the
reference's shipped golden example/haploid.malva.vcf: with these semantics the reference reproduces it byte for byte,
with -ci1 149 of 418 records differ.  The product (malva_b200/) never imports this module."""
from __future__ import annotations

from collections import Counter
from typing import Iterable, Tuple

import numpy as np

from malva_b200.kmc import _CODE, canonical_int, ints_to_packed


def count_kmers(reads: Iterable[str], k: int, min_count: int = 2, counter_max: int = 255,
                canonical: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Emulate ``kmc -k<k> -ci<min_count> -cs<counter_max>`` on an iterable of reads.

    k-mers containing a non-ACGT symbol are skipped, as KMC does.  Returns the
    sorted packed k-mers and their (capped) u32 counts.
    """
    cnt: Counter = Counter()
    mask = (1 << (2 * k)) - 1
    for r in reads:
        r = r.strip().upper()
        x = 0
        valid = 0
        for ch in r:
            c = _CODE.get(ch)
            if c is None:
                valid = 0
                x = 0
                continue
            x = ((x << 2) | c) & mask
            valid += 1
            if valid >= k:
                cnt[canonical_int(x, k) if canonical else x] += 1
    keys = sorted(v for v, c in cnt.items() if c >= min_count)
    counts = np.array([min(cnt[v], counter_max) for v in keys], dtype=np.uint32)
    return ints_to_packed(keys), counts


def read_fastx(path: str) -> list:
    """Minimal FASTA/FASTQ sequence reader (plain or gz)."""
    import gzip

    op = gzip.open if path.endswith(".gz") else open
    seqs = []
    with op(path, "rt") as fh:
        lines = [l.rstrip("\n") for l in fh]
    i = 0
    while i < len(lines):
        l = lines[i]
        if l.startswith("@"):
            seqs.append(lines[i + 1])
            i += 4
        elif l.startswith(">"):
            j = i + 1
            s = []
            while j < len(lines) and not lines[j].startswith(">"):
                s.append(lines[j])
                j += 1
            seqs.append("".join(s))
            i = j
        else:
            i += 1
    return seqs
