#!/usr/bin/env python
"""Generates tests/golden/cli/<case>.expected.vcf.gz: the output of the REFERENCE's own `malva-geno index` +
`malva-geno call -v` (oracle/_ref/malva-geno-ref: the reference's unmodified main.cpp compiled against the
stand-in library headers of oracle/shim/) on the seeded synthetic inputs of tests/synth.py, plus a manifest with
the SHA-256 of every input file so that the GPU test can tell that it regenerated the same inputs.

    python tests/golden/make_cli_golden.py          (needs oracle/_ref, i.e. the build container)
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "malva-geno-ref")
OUT = os.path.join(HERE, "cli")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def input_hashes(fa, vcf, prefix):
    return {"fasta": sha(fa), "vcf": sha(vcf), "kmc_pre": sha(prefix + ".kmc_pre"), "kmc_suf": sha(prefix + ".kmc_suf")}


def cli_flags(case, workdir=None):
    """flags common to both programs: -k/-r/-b 1 + the case's own ("@SAMPLES" -> <workdir>/samples.txt)"""
    fl, it = [], iter(case.flags)
    for f in it:
        if f in ("-k", "-r"):
            next(it)
            continue
        fl.append(os.path.join(workdir, "samples.txt") if f == "@SAMPLES" else f)
    return ["-k", str(case.k), "-r", str(case.ref_k), "-b", "1"] + fl


def run_reference(case, workdir, verbose=True):
    fa, vcf, prefix, n = synth.build_case(case, workdir)
    flags = cli_flags(case, workdir)
    subprocess.run([REF_BIN, "index"] + flags + [fa, vcf, prefix], check=True, capture_output=True)
    r = subprocess.run([REF_BIN, "call"] + (["-v"] if verbose else []) + flags + [fa, vcf, prefix], check=True,
                       capture_output=True)
    os.remove(vcf + f".c{case.ref_k}.k{case.k}.malvax.zst")
    return r.stdout, input_hashes(fa, vcf, prefix), n


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for case in synth.CASES:
        with tempfile.TemporaryDirectory() as d:
            out, hashes, n = run_reference(case, d)
        with gzip.GzipFile(os.path.join(OUT, case.name + ".expected.vcf.gz"), "wb", mtime=0) as fh:
            fh.write(out)
        manifest[case.name] = {"inputs": hashes, "sample_kmers": n, "records": sum(1 for l in out.split(b"\n") if l and l[:1] != b"#"),
                               "expected_sha256": hashlib.sha256(out).hexdigest()}
        print(case.name, manifest[case.name]["records"], "records")
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
