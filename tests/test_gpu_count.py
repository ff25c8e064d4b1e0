"""K6, canonical k-mer counting on the GPU (the `kmc` step in front of malva-geno), against the Python
restatement of KMC's defaults (oracle/kmc_count.py: count_kmers: canonical k-mers, windows with non-ACGT skipped, -ci2,
-cs255) that reproduces the reference's shipped golden, and end to end: reads -> `malva-geno count` -> `index` ->
`call` must give the VCF the reference ships for its haploid example."""
import gzip
import os
import random
import subprocess

import numpy as np
import pytest

from malva_b200 import KmerCounter, MalvaGpu, kmc
from malva_b200 import build as mbuild
from oracle import kmc_count

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "haploid")


def random_reads(rng, n_reads, genome_len=20000, read_len=(60, 260), n_rate=0.002):
    genome = "".join(rng.choice("ACGT") for _ in range(genome_len))
    reads = []
    for _ in range(n_reads):
        ln = rng.randrange(*read_len)
        p = rng.randrange(0, genome_len - ln)
        r = list(genome[p:p + ln])
        if rng.random() < 0.5:
            r = [{"A": "T", "C": "G", "G": "C", "T": "A"}[c] for c in reversed(r)]
        for i in range(len(r)):
            u = rng.random()
            if u < n_rate:
                r[i] = "N"
            elif u < n_rate + 0.003:
                r[i] = rng.choice("ACGT")
        reads.append("".join(r))
    reads += ["ACGT" * 5, "", "A" * 300, "N" * 50, genome[:42], genome[:43], genome[100:143] + "N" + genome[100:143]]
    return reads


@pytest.mark.parametrize("k", [43, 31, 21, 63])
def test_counts_match_the_kmc_restatement(k):
    rng = random.Random(100 + k)
    reads = random_reads(rng, 1500)
    exp_k, exp_c = kmc_count.count_kmers(reads, k, min_count=2, counter_max=255)
    c = KmerCounter(k)
    try:
        c.add(reads)
        got_k, got_c = c.finish(2, 255)
        assert np.array_equal(got_k, exp_k) and np.array_equal(got_c, exp_c)
        st = c.stats()
        assert st["instances"] == sum(max(0, len(s) - k + 1) for r in reads for s in r.split("N"))
        # other thresholds on the same table: -ci1 with a counter cap of 3, -ci5
        for ci, cs in ((1, 3), (5, 255)):
            e_k, e_c = kmc_count.count_kmers(reads, k, min_count=ci, counter_max=cs)
            g_k, g_c = c.finish(ci, cs)
            assert np.array_equal(g_k, e_k) and np.array_equal(g_c, e_c)
    finally:
        c.close()


def test_many_calls_small_chunks_and_partitioned_passes(monkeypatch):
    """records fed one call at a time, device sub-chunks of 1 KiB (k-1 overlap at every seam), and three
    prefix-partitioned passes concatenated: all give the single-pass result"""
    k = 43
    rng = random.Random(7)
    reads = random_reads(rng, 600, read_len=(100, 3000))
    exp_k, exp_c = kmc_count.count_kmers(reads, k, min_count=2, counter_max=255)
    monkeypatch.setenv("MG_COUNT_CHUNK", "1024")
    c = KmerCounter(k)
    try:
        for r in reads:
            c.add([r])
        got_k, got_c = c.finish(2, 255)
        assert np.array_equal(got_k, exp_k) and np.array_equal(got_c, exp_c)
        parts_k, parts_c = [], []
        for lo, hi in ((0, 70), (70, 71), (71, 256)):
            c.reset()
            c.set_partition(8, lo, hi)
            c.add(reads)
            pk, pc = c.finish(2, 255)
            parts_k.append(pk)
            parts_c.append(pc)
        assert np.array_equal(np.concatenate(parts_k), exp_k) and np.array_equal(np.concatenate(parts_c), exp_c)
    finally:
        c.close()


def test_empty_and_too_short_inputs():
    c = KmerCounter(43)
    try:
        c.add([""])
        c.add(["ACGT" * 10])
        k, n = c.finish(1, 255)
        assert len(k) == 0 and len(n) == 0
    finally:
        c.close()


def test_counted_kmers_go_straight_into_the_scan():
    """mg_scan_counted == writing the database and scanning it"""
    import parity_util as util

    rng = random.Random(11)
    k, ref_k, bits = 35, 43, 1 << 22
    genome = util.make_genome(rng, 20000)
    nested, _ = util.synth_signatures(rng, genome, k, 300)
    ks, fl = util.flatten(nested)
    words, _, _ = util.synth_sample(rng, genome, nested, k, ref_k, 6000)   # 43-mers carrying alt alleles, both strands
    reads = [genome[p:p + 150] for p in range(0, len(genome) - 150, 37)] * 3 + [w for w in words for _ in range(2)]
    a, b = MalvaGpu(k=k, ref_k=ref_k, bf_bits=bits), MalvaGpu(k=k, ref_k=ref_k, bf_bits=bits)
    c = KmerCounter(ref_k)
    try:
        for g in (a, b):
            g.add_signatures(ks, fl)
            g.finalize_alt()
            g.scan_reference(genome)
            g.finalize_context()
        c.add(reads)
        packed, counts = c.finish(2, 255)
        assert len(packed) > 1000
        a.scan_sample_kmers(packed, counts)
        b.scan_counted(c)
        assert np.array_equal(a.bf_counts(), b.bf_counts())
        assert a.bf_counts().sum() > 0
        ref_kmers = [x for x, f in zip(ks, fl) if f]
        assert np.array_equal(a.get_counts(ref_kmers, [1] * len(ref_kmers)), b.get_counts(ref_kmers, [1] * len(ref_kmers)))
    finally:
        a.close(), b.close(), c.close()


def test_cli_count_reproduces_the_database_and_the_shipped_golden(tmp_path):
    """README.md:137 from the reads on: `MALVA -1 -k 35 -r 43 -b 1 -f AF haploid.fa haploid.vcf haploid.fq` with
    `malva-geno count` in the place of `kmc -m4 -k43 -t1 -fm` (MALVA:107) -> example/haploid.malva.vcf"""
    mbuild.build()
    cli = mbuild.CLI
    fq = tmp_path / "haploid.fq"
    fq.write_bytes(gzip.open(os.path.join(GOLD, "haploid.fq.gz")).read())
    for f in ("haploid.fa", "haploid.vcf.gz"):
        os.symlink(os.path.join(GOLD, f), tmp_path / f)
    prefix = str(tmp_path / "haploid.fq_malva43.kmercount")
    r = subprocess.run([cli, "count", "-m4", "-k43", "-t1", "-fm", str(fq), prefix, str(tmp_path / "tmp")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    for ext in (".kmc_pre", ".kmc_suf"):
        assert open(prefix + ext, "rb").read() == open(os.path.join(GOLD, "haploid" + ext), "rb").read(), ext
    # gz input, flags separated from their values: same database
    r = subprocess.run([cli, "count", "-k", "43", "-ci", "2", "-cs", "255", os.path.join(GOLD, "haploid.fq.gz"),
                        str(tmp_path / "again")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert open(str(tmp_path / "again") + ".kmc_suf", "rb").read() == open(prefix + ".kmc_suf", "rb").read()
    # three partitioned passes: same listing (the LUT prefix length differs, the records do not)
    r = subprocess.run([cli, "count", "-k43", "--passes", "3", str(fq), str(tmp_path / "p3")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    d1 = subprocess.run([cli, "kmc-dump", prefix], capture_output=True, check=True).stdout
    d3 = subprocess.run([cli, "kmc-dump", str(tmp_path / "p3")], capture_output=True, check=True).stdout
    assert d1 == d3 and d1.count(b"\n") == 4503
    flags = ["-1", "-k", "35", "-r", "43", "-b", "1", "-f", "AF"]
    args = [str(tmp_path / "haploid.fa"), str(tmp_path / "haploid.vcf.gz"), prefix]
    r = subprocess.run([cli, "index"] + flags + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    r = subprocess.run([cli, "call"] + flags + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout == open(os.path.join(GOLD, "haploid.malva.vcf"), "rb").read()


def test_wrapper_script_runs_the_whole_pipeline(tmp_path):
    """README.md:137 verbatim: MALVA -1 -k 35 -r 43 -b 1 -f AF haploid.fa haploid.vcf haploid.fq > out.vcf"""
    mbuild.build()
    wrapper = os.path.join(os.path.dirname(mbuild.CLI), "MALVA")
    (tmp_path / "haploid.fq").write_bytes(gzip.open(os.path.join(GOLD, "haploid.fq.gz")).read())
    (tmp_path / "haploid.vcf").write_bytes(gzip.open(os.path.join(GOLD, "haploid.vcf.gz")).read())
    os.symlink(os.path.join(GOLD, "haploid.fa"), tmp_path / "haploid.fa")
    cmd = [wrapper, "-1", "-k", "35", "-r", "43", "-b", "1", "-f", "AF", "haploid.fa", "haploid.vcf", "haploid.fq"]
    r = subprocess.run(cmd, capture_output=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout == open(os.path.join(GOLD, "haploid.malva.vcf"), "rb").read()
    assert os.path.exists(tmp_path / "haploid.fq_malva43.kmercount.kmc_suf")
    assert os.path.exists(tmp_path / "haploid.vcf.c43.k35.malvax.zst")
    # a second run reuses the k-mer database; the index is rebuilt (its file name does not carry -s / -1 / -u / -f / -p
    # / -b or the reference, all of which its content depends on): other flags, other -- correct -- answer
    r2 = subprocess.run(cmd, capture_output=True, cwd=tmp_path)
    assert r2.returncode == 0 and r2.stdout == r.stdout and b"Found k-mer database" in r2.stderr
    cmd_u = [wrapper, "-1", "-u", "-k", "35", "-r", "43", "-b", "1", "haploid.fa", "haploid.vcf", "haploid.fq"]
    r3 = subprocess.run(cmd_u, capture_output=True, cwd=tmp_path)
    r4 = subprocess.run([mbuild.CLI, "index", "-1", "-u", "-b", "1", "haploid.fa", "haploid.vcf",
                         "haploid.fq_malva43.kmercount"], capture_output=True, cwd=tmp_path)
    r5 = subprocess.run([mbuild.CLI, "call", "-1", "-u", "-b", "1", "haploid.fa", "haploid.vcf",
                         "haploid.fq_malva43.kmercount"], capture_output=True, cwd=tmp_path)
    assert r3.returncode == 0 and r4.returncode == 0 and r5.returncode == 0 and r3.stdout == r5.stdout


def test_cli_count_reads_fasta_and_fastq_layouts(tmp_path):
    """multi-line FASTA (lower case, blank lines, CRLF), FASTQ with an empty read and '@' as the first quality symbol,
    gz: the database lists what the KMC restatement counts on the parsed sequences"""
    mbuild.build()
    cli = mbuild.CLI
    rng = random.Random(21)
    seqs = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(50, 400))) for _ in range(60)]
    seqs[5] = seqs[5][:100] + "N" + seqs[5][100:]
    seqs += seqs[:40]                                # so that some k-mers reach -ci2
    fa = tmp_path / "reads.fa"
    with open(fa, "w", newline="") as fh:
        fh.write("\r\n")
        for i, s in enumerate(seqs):
            fh.write(f">r{i} some text\r\n")
            body = s.lower() if i % 3 == 0 else s
            for o in range(0, len(body), 61):
                fh.write(body[o:o + 61] + "\r\n")
            if i % 7 == 0:
                fh.write("\r\n")
    fq = tmp_path / "reads.fq.gz"
    with gzip.open(fq, "wt") as fh:
        for i, s in enumerate(seqs + [""]):
            fh.write(f"@r{i}\n{s}\n+\n{'@' * len(s)}\n")
    exp_k, exp_c = kmc_count.count_kmers(seqs, 43, min_count=2, counter_max=255)
    exp = [f"{a}\t{b}" for a, b in zip(kmc.packed_to_strings(exp_k, 43), exp_c.tolist())]
    assert len(exp) > 1000
    for src in (fa, fq):
        out = str(tmp_path / (src.name + ".db"))
        r = subprocess.run([cli, "count", "-k43", str(src), out], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        got = subprocess.run([cli, "kmc-dump", out], capture_output=True, text=True, check=True).stdout.split("\n")
        assert [l for l in got if l] == exp, src.name
