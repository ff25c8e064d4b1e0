"""Vectorised generator of chromosome-scale inputs for the whole-program comparison (tests/bench_cli_e2e.py --fast):
SURVEY 8d's cfg3 shape -- one contig of uniform ACGT (60-column lines) with a 10 kb N run every 50 Mb, sorted SNV / indel records with
exponential gaps (mean 41 bp), 88 % SNVs and 12 % indels (geometric length, mean 3, at most 50), AF ~ min(0.5,
1 / (2 N u)), 32 phased samples drawn from AF, and a donor whose two haplotypes carry the SNVs of a genotype drawn from
AF (indel records are genotyped too; the donor carries their reference allele).  tests/synth.py stays the generator of
the seeded parity cases (Python `random`, 10 s per Mbp); this one makes 250 Mbp in about a minute with numpy.
The donor's 43-mers are counted by `malva-geno count` (K6) from donor.fa.  Test infrastructure only."""
from __future__ import annotations

import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def build(outdir: str, n_bases: int, n_samples: int = 32, mean_gap: int = 41, seed: int = 20261018 + 3, k: int = 35):
    g = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    codes = g.integers(0, 4, n_bases, dtype=np.uint8)
    ref = ACGT[codes]
    for s in range(25_000_000, n_bases, 50_000_000):
        ref[s:s + 10_000] = ord("N")
    fa = os.path.join(outdir, "ref.fa")
    with open(fa, "wb") as fh:
        fh.write(b">1\n")
        whole = (n_bases // 60) * 60            # 60-column lines, like the reference genomes in circulation
        fh.write(np.hstack([ref[:whole].reshape(-1, 60), np.full((whole // 60, 1), 10, np.uint8)]).tobytes())
        if whole < n_bases:
            fh.write(ref[whole:].tobytes())
            fh.write(b"\n")
    # ---- records ----
    n_guess = int(n_bases / mean_gap * 1.05) + 16
    pos = np.cumsum(g.exponential(mean_gap, n_guess).astype(np.int64) + 1) + 100
    pos = pos[pos < n_bases - 200]
    kind = g.random(len(pos))                       # < .88 SNV, < .94 insertion, else deletion
    ilen = np.minimum(g.geometric(1 / 3.0, len(pos)), 50).astype(np.int64)
    dele = kind >= 0.94
    # a deletion must end before the next record starts (no overlapping records in this generator)
    nxt = np.empty_like(pos)
    nxt[:-1], nxt[-1] = pos[1:], n_bases
    ilen = np.where(dele, np.minimum(ilen, np.maximum(nxt - pos - 2, 1)), ilen)
    span = np.where(dele, ilen + 1, 1)
    # drop records whose REF touches an N run
    isn = (ref == ord("N")).astype(np.int32)
    csum = np.concatenate([[0], np.cumsum(isn)])
    ok = (csum[np.minimum(pos + span, n_bases)] - csum[pos]) == 0
    pos, kind, ilen, dele, span = pos[ok], kind[ok], ilen[ok], dele[ok], span[ok]
    nv = len(pos)
    snv = kind < 0.88
    ins = ~snv & ~dele
    alt_code = (codes[pos] + g.integers(1, 4, nv, dtype=np.uint8)) & 3     # a different base
    u = 1.0 - g.random(nv)
    af = np.minimum(0.5, 1.0 / (2 * n_samples * u))
    gt = (g.random((nv, n_samples, 2)) < af[:, None, None])
    pat = np.array(["0|0", "0|1", "1|0", "1|1"])
    gcode = gt[:, :, 0] * 2 + gt[:, :, 1]
    ins_seq = ACGT[g.integers(0, 4, (nv, 50), dtype=np.uint8)]
    vcf = os.path.join(outdir, "vars.vcf")
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##contig=<ID=1,length=%d>\n" % n_bases)
        fh.write('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele frequency">\n')
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"S{i}" for i in range(n_samples)) + "\n")
        refb = ref.tobytes()
        CH = 200_000
        for a in range(0, nv, CH):
            b = min(nv, a + CH)
            rows = pat[gcode[a:b]]
            gts = ["\t".join(r) for r in rows.tolist()]
            out = []
            for j in range(a, b):
                p = int(pos[j])
                if snv[j]:
                    r_, a_ = refb[p:p + 1], bytes([ACGT[alt_code[j]]])
                elif ins[j]:
                    r_, a_ = refb[p:p + 1], refb[p:p + 1] + ins_seq[j, :int(ilen[j])].tobytes()
                else:
                    r_, a_ = refb[p:p + 1 + int(ilen[j])], refb[p:p + 1]
                out.append("1\t%d\t.\t%s\t%s\t.\tPASS\tAF=%.6g\tGT\t%s\n" % (p + 1, r_.decode(), a_.decode(), af[j], gts[j - a]))
            fh.write("".join(out))
    # ---- donor: two haplotypes carrying the SNVs of a genotype drawn from AF ----
    donor = os.path.join(outdir, "donor.fa")
    dg = g.random((nv, 2)) < np.maximum(af, 0.05)[:, None]
    with open(donor, "wb") as fh:
        for h in range(2):
            hap = ref.copy()
            sel = snv & dg[:, h]
            hap[pos[sel]] = ACGT[alt_code[sel]]
            fh.write(b">hap%d\n" % h)
            fh.write(hap.tobytes())
            fh.write(b"\n")
    return fa, vcf, donor, nv
