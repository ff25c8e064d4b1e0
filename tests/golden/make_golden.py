#!/usr/bin/env python
"""Regenerates tests/golden/haploid/ from the reference's bundled example.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py

Inputs  : /root/reference/example/haploid.tar.gz (haploid.fa, haploid.fq, haploid.vcf)
          /root/reference/example/haploid.malva.vcf (the reference's shipped golden)
Outputs : haploid.fa, haploid.vcf.gz, haploid.fq.gz -- inputs of the example, verbatim
          haploid.kmc_pre/.kmc_suf             -- reads counted as `kmc -k43 -ci2 -cs255` does
                                                  (oracle/kmc_count.py: count_kmers)
          haploid.malva.vcf                    -- shipped golden (GT:GQ only)
          haploid.malva.verbose.vcf            -- oracle/_ref `call -v` output: pins COVS + GTS
The shim-built reference must reproduce haploid.malva.vcf byte for byte, else this aborts.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from malva_b200 import kmc  # noqa: E402
from oracle import kmc_count  # noqa: E402

REF = "/root/reference/example"
BIN = os.path.join(ROOT, "oracle", "_ref", "malva-geno-ref")
OUT = os.path.join(HERE, "haploid")


def main():
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        tarfile.open(os.path.join(REF, "haploid.tar.gz")).extractall(tmp, filter="data")
        reads = kmc_count.read_fastx(os.path.join(tmp, "haploid.fq"))
        km, ct = kmc_count.count_kmers(reads, 43, min_count=2, counter_max=255)
        kmc.write_kmc_db(os.path.join(tmp, "haploid"), km, ct, 43)
        flags = ["-k", "35", "-r", "43", "-b", "1", "-f", "AF", "-1"]
        args = ["haploid.fa", "haploid.vcf", "haploid"]
        subprocess.check_call([BIN, "index"] + flags + args, cwd=tmp, stderr=subprocess.DEVNULL)
        plain = subprocess.check_output([BIN, "call"] + flags + args, cwd=tmp, stderr=subprocess.DEVNULL)
        verbose = subprocess.check_output([BIN, "call", "-v"] + flags + args, cwd=tmp, stderr=subprocess.DEVNULL)
        shipped = open(os.path.join(REF, "haploid.malva.vcf"), "rb").read()
        if plain != shipped:
            raise SystemExit("oracle/_ref does not reproduce example/haploid.malva.vcf")
        shutil.copy(os.path.join(tmp, "haploid.fa"), OUT)
        with open(os.path.join(tmp, "haploid.vcf"), "rb") as src, \
                gzip.GzipFile(os.path.join(OUT, "haploid.vcf.gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        with open(os.path.join(tmp, "haploid.fq"), "rb") as src, \
                gzip.GzipFile(os.path.join(OUT, "haploid.fq.gz"), "wb", mtime=0) as dst:
            dst.write(src.read())   # the reads themselves: input of the GPU k-mer counter test
        for ext in (".kmc_pre", ".kmc_suf"):
            shutil.copy(os.path.join(tmp, "haploid" + ext), OUT)
        open(os.path.join(OUT, "haploid.malva.vcf"), "wb").write(shipped)
        open(os.path.join(OUT, "haploid.malva.verbose.vcf"), "wb").write(verbose)
    print("wrote", OUT, f"({len(km)} k-mers)")


if __name__ == "__main__":
    main()
