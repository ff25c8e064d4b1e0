"""Seeded synthetic inputs for the end-to-end parity tests: a reference FASTA, a VCF with AF + GT columns and a
KMC database of a donor's 43-mers, small enough that the shim-built reference `malva-geno`
(oracle/_ref) finishes in seconds.  Test infrastructure only.

The VCF is deliberately nasty: dense clusters (most variants have neighbours within k/2), SNVs, insertions,
deletions that span later records, multi-allelic records, records at the same position, symbolic ALTs, an
allele longer than k, AF = 0 records (absent from every sample: skipped by `index`, genotyped 0/0 by `call`),
phased / unphased / missing genotypes, lower-case bases, N runs and IUPAC symbols in the reference.
"""
from __future__ import annotations

import gzip
import os
import random
from dataclasses import dataclass, field
from typing import List

import numpy as np

from malva_b200 import kmc

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


@dataclass
class Rec:
    chrom: str
    pos0: int
    ref: str
    alts: List[str]
    af: List[float]
    gts: List[str] = field(default_factory=list)


def make_reference(rng: random.Random, contigs, n_run_every=60_000):
    refs = {}
    for name, length in contigs:
        s = rng.choices("ACGT", k=length)
        for start in range(n_run_every // 2, length, n_run_every):   # N runs
            for i in range(start, min(length, start + 60)):
                s[i] = "N"
        for _ in range(max(1, length // 20_000)):                    # scattered IUPAC symbols
            s[rng.randrange(length)] = rng.choice("RYMKSW")
        refs[name] = "".join(s)
    return refs


def make_variants(rng: random.Random, refs, mean_gap, n_samples, haploid, multi_frac=0.06, sym_frac=0.01,
                  long_frac=0.004, k=35):
    recs: List[Rec] = []
    for chrom, seq in refs.items():
        pos = rng.randrange(5, 60)
        while pos < len(seq) - 60:
            r = rng.random()
            ref_len = 1
            if r < 0.10:
                ref_len = 1 + min(rng.randrange(1, 8), rng.randrange(1, 30))      # deletion
            ref = seq[pos:pos + ref_len]
            if any(c not in "ACGT" for c in ref):
                pos += rng.randrange(1, 2 * mean_gap)
                continue
            n_alt = 1 if rng.random() > multi_frac else rng.randrange(2, 4)
            alts = []
            while len(alts) < n_alt:
                if ref_len > 1:
                    a = ref[0] if rng.random() < 0.8 else ref[0] + "".join(rng.choices("ACGT", k=rng.randrange(1, 3)))
                elif r < 0.20:                                                     # insertion
                    ln = rng.randrange(1, 6) if rng.random() > long_frac * 10 else rng.randrange(k, k + 12)
                    a = ref + "".join(rng.choices("ACGT", k=ln))
                else:
                    a = rng.choice([c for c in "ACGT" if c != ref])
                if a != ref and (a not in alts or ref_len == 1 and r >= 0.20):
                    alts.append(a)                       # (an SNV may list the same base twice: duplicate-text alleles)
                if ref_len == 1 and r >= 0.20 and len(set(alts)) == 3:
                    break
            if rng.random() < sym_frac:
                alts.insert(rng.randrange(len(alts) + 1), "<CN0>")
            real = [a for a in alts if not a.startswith("<")]
            af = []
            for a in alts:
                u = rng.random()
                af.append(0.0 if u < 0.04 else round(min(0.5, 0.02 / max(u, 1e-3)), 5))
            if sum(af) > 0.95:
                af = [x / 2 for x in af]
            # genotypes of the panel samples, drawn from AF over the KEPT alts (htslib indexes all alts: the reference
            # reads allele i of the record, symbolic ones included, and then indexes its own alts list with it --
            # keep GT indices within the kept alts so that the reference does not run off its vector)
            gts = []
            n_real = len(real)
            for _ in range(n_samples):
                def draw():
                    u = rng.random()
                    acc = 0.0
                    for i in range(n_real):
                        acc += af[i]
                        if u < acc:
                            return i + 1
                    return 0
                if n_real == 0:
                    gts.append("0" if haploid else "0|0")
                    continue
                if haploid:
                    g = draw()
                    gts.append("." if rng.random() < 0.01 else str(g))
                else:
                    a, b = draw(), draw()
                    sep = "|" if rng.random() < 0.85 else "/"
                    u = rng.random()
                    if u < 0.01:
                        gts.append("." + sep + str(b))
                    elif u < 0.015:
                        gts.append("./.")
                    else:
                        gts.append(f"{a}{sep}{b}")
            recs.append(Rec(chrom, pos, ref, alts, af, gts))
            if rng.random() < 0.03:
                continue                                                           # another record at the same position
            pos += max(1, int(rng.expovariate(1.0 / mean_gap)))
    return recs


def write_vcf(path, refs, recs, n_samples, haploid, lower_case_frac, rng, freq_key="AF", with_format=True):
    import io

    if path.endswith(".gz"):   # no timestamp / file name in the gzip header: the bytes depend on the seed only
        raw = open(path, "wb")
        fh_ctx = io.TextIOWrapper(gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0))
    else:
        raw, fh_ctx = None, open(path, "wt")
    with fh_ctx as fh:
        fh.write("##fileformat=VCFv4.2\n")
        for name, seq in refs.items():
            fh.write(f"##contig=<ID={name},length={len(seq)}>\n")
        fh.write(f'##INFO=<ID={freq_key},Number=A,Type=Float,Description="Allele frequency">\n')
        fh.write('##INFO=<ID=NS,Number=1,Type=Integer,Description="Samples">\n')
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write('##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">\n')
        cols = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"
        if with_format:
            cols += "\tFORMAT\t" + "\t".join(f"S{i}" for i in range(n_samples))
        fh.write(cols + "\n")
        for i, r in enumerate(recs):
            ref, alts = r.ref, list(r.alts)
            if rng.random() < lower_case_frac:
                ref = ref.lower()
                alts = [a.lower() if not a.startswith("<") else a for a in alts]
            qual = "." if i % 3 else str(rng.choice([30, 99.5, 1234.25, 7]))
            info = f"NS={n_samples};{freq_key}=" + ",".join(f"{x:g}" for x in r.af)
            line = [r.chrom, str(r.pos0 + 1), f"rs{i}" if i % 5 else ".", ref, ",".join(alts), qual, "PASS", info]
            if with_format:
                if i % 7 == 0:
                    line += ["DP:GT"] + [f"{rng.randrange(1, 60)}:{g}" for g in r.gts]
                else:
                    line += ["GT"] + r.gts
            fh.write("\t".join(line) + "\n")
    if raw is not None:
        raw.close()


def write_fasta(path, refs, width=70, chr_prefix=False):
    with open(path, "w") as fh:
        for name, seq in refs.items():
            fh.write(f">{'chr' if chr_prefix else ''}{name} synthetic contig\n")
            for i in range(0, len(seq), width):
                chunk = seq[i:i + width]
                fh.write((chunk.lower() if (i // width) % 11 == 0 else chunk) + "\n")


def donor_haplotypes(rng: random.Random, refs, recs, haploid):
    """Apply a random non-overlapping subset of the variants to the reference: 1 or 2 haplotype strings per contig."""
    haps = {}
    for chrom, seq in refs.items():
        mine = [r for r in recs if r.chrom == chrom]
        out = []
        for _ in range(1 if haploid else 2):
            parts, last = [], 0
            for r in mine:
                real = [a for a in r.alts if not a.startswith("<")]
                if r.pos0 < last or not real:
                    continue
                u, acc, pick = rng.random(), 0.0, None
                for a, f in zip(real, r.af):
                    acc += max(f, 0.15)          # donor carries more variants than the panel average
                    if u < acc:
                        pick = a
                        break
                if pick is None:
                    continue
                parts.append(seq[last:r.pos0])
                parts.append(pick)
                last = r.pos0 + len(r.ref)
            parts.append(seq[last:])
            out.append("".join(parts))
        haps[chrom] = out
    return haps


def _windows(seq: str, k: int):
    """canonical packed k-mers of every all-ACGT window of seq -> (lo, hi) uint64 arrays."""
    b = np.frombuffer(seq.encode(), dtype=np.uint8)
    code = np.full(len(b), 4, np.uint64)
    for i, ch in enumerate(b"ACGT"):
        code[b == ch] = i
    n = len(b) - k + 1
    if n <= 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    bad = (code == 4).astype(np.int64)
    ok = (np.convolve(bad, np.ones(k, np.int64), "valid") == 0)
    code = code & np.uint64(3)
    lo = np.zeros(n, np.uint64)
    hi = np.zeros(n, np.uint64)
    rlo = np.zeros(n, np.uint64)
    rhi = np.zeros(n, np.uint64)
    for j in range(k):                       # symbol j of the window sits at bit 2*(k-1-j)
        c = code[j:j + n]
        sh = 2 * (k - 1 - j)
        if sh >= 64:
            hi |= c << np.uint64(sh - 64)
        else:
            lo |= c << np.uint64(sh)
        rc = np.uint64(3) - c                # reverse complement: symbol j lands at bit 2*j
        sh = 2 * j
        if sh >= 64:
            rhi |= rc << np.uint64(sh - 64)
        else:
            rlo |= rc << np.uint64(sh)
    take_rc = (rhi < hi) | ((rhi == hi) & (rlo < lo))
    lo = np.where(take_rc, rlo, lo)[ok]
    hi = np.where(take_rc, rhi, hi)[ok]
    return lo, hi


def sample_kmc_db(rng: random.Random, prefix, haps, ref_k, mean_cov, error_frac=0.15, min_count=2, counter_max=255):
    """Count the donor's canonical ref_k-mers like `kmc -ci2 -cs255` would (counts drawn per k-mer occurrence
    instead of simulating reads) and write <prefix>.kmc_pre/.kmc_suf."""
    los, his, cts = [], [], []
    g = np.random.default_rng(rng.randrange(1 << 30))
    for chrom, hs in haps.items():
        for h in hs:
            lo, hi = _windows(h, ref_k)
            los.append(lo)
            his.append(hi)
            cts.append(g.poisson(mean_cov / len(hs), len(lo)))
            # sequencing-error k-mers: one substitution in a true window, low count
            n_err = int(len(lo) * error_frac)
            if n_err:
                idx = g.integers(0, max(1, len(h) - ref_k), n_err)
                for i in idx[:2000]:
                    w = h[i:i + ref_k]
                    if len(w) < ref_k or any(c not in "ACGT" for c in w):
                        continue
                    p = rng.randrange(ref_k)
                    w = w[:p] + rng.choice([c for c in "ACGT" if c != w[p]]) + w[p + 1:]
                    elo, ehi = _windows(w, ref_k)
                    los.append(elo)
                    his.append(ehi)
                    cts.append(np.array([rng.randrange(1, 4)]))
    lo, hi, ct = np.concatenate(los), np.concatenate(his), np.concatenate(cts).astype(np.int64)
    keys = np.zeros(len(lo), dtype=kmc.KMER_DTYPE)
    keys["lo"], keys["hi"] = lo, hi
    order = np.lexsort((lo, hi))
    keys, ct = keys[order], ct[order]
    first = np.ones(len(keys), bool)
    first[1:] = (keys["lo"][1:] != keys["lo"][:-1]) | (keys["hi"][1:] != keys["hi"][:-1])
    starts = np.flatnonzero(first)
    sums = np.add.reduceat(ct, starts) if len(ct) else ct
    keep = sums >= min_count
    uk = keys[starts][keep]
    uc = np.minimum(sums[keep], counter_max).astype(np.uint32)
    kmc.write_kmc_db(prefix, uk, uc, ref_k, min_count=min_count, max_count=counter_max)
    return len(uk)


@dataclass
class Case:
    name: str
    seed: int
    contigs: list
    mean_gap: int = 30
    n_samples: int = 6
    haploid: bool = False
    mean_cov: float = 30.0
    k: int = 35
    ref_k: int = 43
    flags: tuple = ()              # extra CLI flags (both programs); "@SAMPLES" = path of the sample-list file
    sample_subset: tuple = ()      # indices of the panel samples listed in that file (file order is NOT header order)
    freq_key: str = "AF"
    gz: bool = False
    chr_prefix: bool = False
    lower_case_frac: float = 0.02


def build_case(case: Case, outdir: str):
    """Writes ref.fa, vars.vcf[.gz], sample.kmc_{pre,suf} into outdir; returns (fasta, vcf, kmc prefix, n_kmers)."""
    rng = random.Random(case.seed)
    os.makedirs(outdir, exist_ok=True)
    refs = make_reference(rng, case.contigs)
    recs = make_variants(rng, refs, case.mean_gap, case.n_samples, case.haploid, k=case.k)
    fa = os.path.join(outdir, "ref.fa")
    vcf = os.path.join(outdir, "vars.vcf" + (".gz" if case.gz else ""))
    write_fasta(fa, refs, chr_prefix=case.chr_prefix)
    write_vcf(vcf, refs, recs, case.n_samples, case.haploid, case.lower_case_frac, rng, freq_key=case.freq_key)
    if case.sample_subset:
        with open(os.path.join(outdir, "samples.txt"), "w") as fh:
            for i in case.sample_subset:
                fh.write(f"S{i}\n")
    haps = donor_haplotypes(rng, refs, recs, case.haploid)
    prefix = os.path.join(outdir, "sample")
    n = sample_kmc_db(rng, prefix, haps, case.ref_k, case.mean_cov)
    return fa, vcf, prefix, n


# small-scale twins of BASELINE.json's configurations + flag coverage
CASES = [
    Case("cfg2_chr20_like_diploid", 20261018 + 2, [("20", 120_000)], mean_gap=40, n_samples=8, freq_key="EUR_AF",
         flags=("-f", "EUR_AF")),
    Case("cfg3_chr1_like_snv_indel", 20261018 + 3, [("1", 150_000)], mean_gap=41, n_samples=6),
    Case("cfg4_wg_like_multiallelic", 20261018 + 4, [("1", 60_000), ("2", 50_000), ("X", 30_000)], mean_gap=36,
         n_samples=8, gz=True),
    Case("cfg5_dense_high_cov", 20261018 + 5, [("1", 80_000)], mean_gap=12, n_samples=24, mean_cov=60.0,
         flags=("-c", "400")),
    Case("haploid_uniform", 20261018 + 6, [("NC_1", 50_000)], mean_gap=25, n_samples=10, haploid=True,
         flags=("-1", "-u")),
    Case("haploid_af", 20261018 + 7, [("NC_1", 50_000)], mean_gap=25, n_samples=10, haploid=True, flags=("-1",)),
    Case("strip_chr_err01_maxcov", 20261018 + 8, [("7", 40_000), ("8", 30_000)], mean_gap=30, n_samples=5,
         chr_prefix=True, flags=("-p", "-e", "0.01", "-c", "25")),
    Case("k31_r39", 20261018 + 9, [("1", 50_000)], mean_gap=30, n_samples=5, k=31, ref_k=39,
         flags=("-k", "31", "-r", "39")),
    Case("samples_subset_uniform", 20261018 + 10, [("3", 40_000)], mean_gap=28, n_samples=12,
         flags=("-s", "@SAMPLES", "-u"), sample_subset=(7, 2, 9, 3)),
]
