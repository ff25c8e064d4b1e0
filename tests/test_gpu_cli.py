"""End to end through the drop-in CLI: `malva-geno index` + `malva-geno call` (C++ host over the C ABI, kernels
on the GPU) must print byte-for-byte the VCF the reference prints on the same inputs.

Expected outputs: tests/golden/cli/*.expected.vcf.gz (the reference's own main.cpp run by
tests/golden/make_cli_golden.py on the seeded inputs of tests/synth.py) and the haploid example's goldens
(tests/golden/haploid: `haploid.malva.vcf` is the file the reference ships in example/)."""
import gzip
import json
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_cli_golden as mk  # noqa: E402
import synth  # noqa: E402
from malva_b200 import build as mbuild  # noqa: E402

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def cli():
    mbuild.build()
    return mbuild.CLI


def run_ours(cli, flags, fa, vcf, prefix, verbose, extra=()):
    r = subprocess.run([cli, "index"] + list(flags) + list(extra) + [fa, vcf, prefix], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"BF creation complete" in r.stderr and b"Reference BF creation complete" in r.stderr
    r = subprocess.run([cli, "call"] + (["-v"] if verbose else []) + list(flags) + list(extra) + [fa, vcf, prefix],
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"BF weights created" in r.stderr and b"Execution completed" in r.stderr
    return r.stdout


def first_diff(a: bytes, b: bytes):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return f"line {i + 1}:\n  got      {x[:400]!r}\n  expected {y[:400]!r}"
    return f"{len(la)} lines vs {len(lb)} expected"


@pytest.mark.parametrize("case", synth.CASES, ids=[c.name for c in synth.CASES])
def test_cli_vcf_is_byte_identical_to_the_reference(cli, case, tmp_path):
    manifest = json.load(open(os.path.join(GOLD, "cli", "manifest.json")))[case.name]
    fa, vcf, prefix, _ = synth.build_case(case, str(tmp_path))
    assert mk.input_hashes(fa, vcf, prefix) == manifest["inputs"], "synthetic inputs differ from the ones the golden was made from"
    expected = gzip.open(os.path.join(GOLD, "cli", case.name + ".expected.vcf.gz")).read()
    got = run_ours(cli, mk.cli_flags(case, str(tmp_path)), fa, vcf, prefix, verbose=True)
    assert got == expected, first_diff(got, expected)


@pytest.mark.parametrize("verbose,golden", [(False, "haploid.malva.vcf"), (True, "haploid.malva.verbose.vcf")])
def test_cli_haploid_example(cli, verbose, golden, tmp_path):
    """README.md:137: malva-geno -1 -k 35 -r 43 -b 1 -f AF on example/haploid -- against the shipped golden"""
    src = os.path.join(GOLD, "haploid")
    for f in ("haploid.fa", "haploid.vcf.gz", "haploid.kmc_pre", "haploid.kmc_suf"):
        os.symlink(os.path.join(src, f), tmp_path / f)
    got = run_ours(cli, ["-1", "-k", "35", "-r", "43", "-b", "1", "-f", "AF"], str(tmp_path / "haploid.fa"),
                   str(tmp_path / "haploid.vcf.gz"), str(tmp_path / "haploid"), verbose)
    expected = open(os.path.join(src, golden), "rb").read()
    assert got == expected, first_diff(got, expected)


def test_cli_sars_cov2_example_config1(cli, tmp_path):
    """BASELINE config 1 on its real input: example/sars_cov2.vcf.gz (15,154 records incl. multi-allelic x 27,934
    haploid samples -- one var_block spanning the genome) + reference_sarsCov2.fasta + the bundled haploid reads,
    `-1 -k 35 -r 43 -b 1`.  Expected: the verbose VCF the reference's own main.cpp printed on the same files
    (tests/golden/sars/sars.expected.verbose.vcf.gz; 310 s index + 307 s call on one core of the build container,
    GT histogram 13169 x 0:0, 1983 x 0:100, 1:94, 1:100 as in SURVEY 8c)."""
    import time
    src, hap = os.path.join(GOLD, "sars"), os.path.join(GOLD, "haploid")
    for f in ("reference_sarsCov2.fasta", "sars_cov2.vcf.gz"):
        os.symlink(os.path.join(src, f), tmp_path / f)
    for f in ("haploid.kmc_pre", "haploid.kmc_suf"):
        os.symlink(os.path.join(hap, f), tmp_path / f)
    t0 = time.time()
    got = run_ours(cli, ["-1", "-k", "35", "-r", "43", "-b", "1"], str(tmp_path / "reference_sarsCov2.fasta"),
                   str(tmp_path / "sars_cov2.vcf.gz"), str(tmp_path / "haploid"), verbose=True)
    print(f"sars_cov2 index + call: {time.time() - t0:.1f} s (reference: 310 s + 307 s)")
    expected = gzip.open(os.path.join(src, "sars.expected.verbose.vcf.gz")).read()
    assert got == expected, first_diff(got, expected)
    gts = [l.split(b"\t")[-1] for l in got.split(b"\n") if l and not l.startswith(b"#")]
    assert len(gts) == 15154 and gts.count(b"0:0") == 13169 and gts.count(b"0:100") == 1983


def test_cli_threads_and_errors(cli, tmp_path):
    """one host thread gives the same bytes; a missing index / KMC database / INFO key fails like the reference"""
    case = synth.CASES[2]
    fa, vcf, prefix, _ = synth.build_case(case, str(tmp_path))
    expected = gzip.open(os.path.join(GOLD, "cli", case.name + ".expected.vcf.gz")).read()
    got = run_ours(cli, mk.cli_flags(case), fa, vcf, prefix, verbose=True, extra=("--threads", "1"))
    assert got == expected, first_diff(got, expected)
    r = subprocess.run([cli, "call", "-b", "1", fa, vcf, prefix + "_nope"], capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR: cannot open" in r.stderr
    os.remove(vcf + f".c{case.ref_k}.k{case.k}.malvax.zst")
    r = subprocess.run([cli, "call", "-b", "1", fa, vcf, prefix], capture_output=True, text=True)
    assert r.returncode == 1 and "index" in r.stderr
    r = subprocess.run([cli, "index", "-b", "1", "-f", "NOPE_AF", fa, vcf, prefix], capture_output=True, text=True)
    assert r.returncode == 1 and "NOPE_AF" in r.stderr


def _n_gpus():
    from malva_b200 import _lib

    return _lib.load().mg_device_count()


@pytest.mark.parametrize("devices", ["0,0", "0,0,0", "0,1"])
def test_cli_call_over_replicated_contexts(cli, devices, tmp_path):
    """`call --devices a,b,..`: the index replicated per context, KMC chunks dealt round-robin, counters added up with
    mg_reduce_counts -- same bytes as the single-context run (contexts may share a device)"""
    if max(int(d) for d in devices.split(",")) >= _n_gpus():
        pytest.skip("needs more GPUs")
    case = synth.CASES[3]
    fa, vcf, prefix, _ = synth.build_case(case, str(tmp_path))
    expected = gzip.open(os.path.join(GOLD, "cli", case.name + ".expected.vcf.gz")).read()
    flags = mk.cli_flags(case)
    r = subprocess.run([cli, "index"] + flags + [fa, vcf, prefix], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    r = subprocess.run([cli, "call", "-v", "--devices", devices] + flags + [fa, vcf, prefix], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout == expected, first_diff(r.stdout, expected)
