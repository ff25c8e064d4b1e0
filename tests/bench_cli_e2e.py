#!/usr/bin/env python
"""Whole-program comparison on one machine: the reference `malva-geno` (oracle/_ref/malva-geno-ref: the reference's
own main.cpp, CPU, single-threaded like the original) against this repository's `malva-geno` (C++ host + B200
kernels) on the same synthetic chromosome-arm-sized inputs; outputs must be byte-identical.

    python tests/bench_cli_e2e.py [Mbp=5] [samples=32] [bf_gb=1] [fast] [ours] > gpurun_out/cli_e2e.json

`ours`: only this repository's binary is run (profiling runs: the reference takes 8 minutes at 250 Mbp).

`fast`: inputs from tests/synth_fast.py (numpy, about a minute for the 250 Mbp of BASELINE's cfg3) instead of the seeded
generator of the parity cases (tests/synth.py: Python `random`, 10 s per Mbp); the donor's 43-mers are then counted on
the GPU by `malva-geno count -ci1` (K6) from the two donor haplotypes.

(lives under tests/ because it executes the oracle build of the reference; not collected by pytest)
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import synth  # noqa: E402
from malva_b200 import build as mbuild  # noqa: E402
from malva_b200 import kmc  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "malva-geno-ref")


def fast_write_kmc(prefix, keys, counts, k, p=7):
    """numpy version of kmc.write_kmc_db for millions of records (KMC2 layout, counter_size 1)."""
    import struct
    lo, hi = keys["lo"], keys["hi"]
    suf_syms = k - p
    assert suf_syms % 4 == 0 and 64 < 2 * suf_syms < 128
    sb = suf_syms // 4
    pref = (hi >> np.uint64(2 * suf_syms - 64)).astype(np.int64)
    lut = np.zeros(4 ** p + 1, np.uint64)
    lut[1:] = np.cumsum(np.bincount(pref, minlength=4 ** p)).astype(np.uint64)
    rec = np.empty((len(lo), sb + 1), np.uint8)
    hi_bytes = sb - 8
    for j in range(hi_bytes):
        rec[:, j] = ((hi >> np.uint64(8 * (hi_bytes - 1 - j))) & np.uint64(0xFF)).astype(np.uint8)
    for j in range(8):
        rec[:, hi_bytes + j] = ((lo >> np.uint64(56 - 8 * j)) & np.uint64(0xFF)).astype(np.uint8)
    rec[:, sb] = counts.astype(np.uint8)
    with open(prefix + ".kmc_suf", "wb") as fh:
        fh.write(b"KMCS")
        fh.write(rec.tobytes())
        fh.write(b"KMCS")
    sig_len = 5
    hdr = struct.pack("<7IQB", k, 0, 1, p, sig_len, 2, 255, len(lo), 0)
    hdr += b"\0" * (60 - len(hdr)) + struct.pack("<I", 0x200)
    with open(prefix + ".kmc_pre", "wb") as fh:
        fh.write(b"KMCP")
        fh.write(lut.astype("<u8").tobytes())
        fh.write(b"\0" * ((4 ** sig_len + 1) * 4))
        fh.write(hdr + struct.pack("<I", len(hdr)) + b"KMCP")


def phases(stderr):
    """the reference's own progress lines; the one-per-5000-variants lines are summed into one entry"""
    out, acc, n, worst = {}, 0.0, 0, 0.0
    for m in re.finditer(r"\[malva-geno/([^\]]+)\] Execution Time ([0-9.e+-]+)s", stderr):
        if m.group(1).startswith("Processed "):
            acc, n, worst = acc + float(m.group(2)), n + 1, max(worst, float(m.group(2)))
        else:
            out[m.group(1)] = float(m.group(2))
    if n:
        out["Processed N variants (sum of %d progress lines)" % n] = round(acc, 4)
        out["Processed N variants (longest single line)"] = worst
    return out


def traces(stderr):
    """--trace lines of this repository's binary; the per-batch lines are summed per stage"""
    lines = [l for l in stderr.split("\n") if l.startswith("[trace]")]
    batch = [l for l in lines if " batch: " in l]
    out = [l for l in lines if " batch: " not in l]
    if batch:
        tot = {}
        for l in batch:
            for name, ms in re.findall(r"([a-z+ ]+?) ([0-9.]+) ms", l.split(": ", 1)[1]):
                tot[name.strip(", ")] = round(tot.get(name.strip(", "), 0.0) + float(ms), 1)
        out.append("[%d per-batch trace lines, ms summed per stage] %s" % (len(batch), json.dumps(tot)))
        out += batch[:2]
    return out


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    bf_gb = sys.argv[3] if len(sys.argv) > 3 else "1"          # -b: filter size in GB (reference default: 4)
    fast = len(sys.argv) > 4 and sys.argv[4] == "fast"
    ours_only = "ours" in sys.argv[4:]
    mbuild.build()
    case = synth.Case("cli_e2e", 20261018 + 42, [("1", int(mbp * 1e6))], mean_gap=41, n_samples=n_samples)
    synth_write = kmc.write_kmc_db
    kmc.write_kmc_db = lambda prefix, uk, uc, k, **kw: fast_write_kmc(prefix, uk, uc, k)
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        if fast:
            import synth_fast
            fa, vcf, donor, _ = synth_fast.build(d, int(mbp * 1e6), n_samples)
            prefix = os.path.join(d, "sample")
            p = subprocess.run([mbuild.CLI, "count", "-k43", "-ci1", "-cs255", donor, prefix], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
            n_kmers = int(re.search(r"(\d+) written", p.stderr).group(1))
            os.remove(donor)
        else:
            fa, vcf, prefix, n_kmers = synth.build_case(case, d)
        kmc.write_kmc_db = synth_write
        n_var = sum(1 for l in open(vcf) if not l.startswith("#"))
        res = {"reference_bases": int(mbp * 1e6), "variants": n_var, "panel_samples": n_samples, "sample_kmers": n_kmers,
               "generate_s": round(time.time() - t0, 1), "host_cores": os.cpu_count(), "bf_gb": bf_gb,
               "generator": "tests/synth_fast.py + malva-geno count" if fast else "tests/synth.py"}
        outs = {}
        for name, exe in (("reference_cpu", REF), ("malva_b200", mbuild.CLI)):
            if ours_only and name == "reference_cpu":
                continue
            r = {}
            for sub in ("index", "call"):
                t = time.time()
                extra = ["--trace"] if name == "malva_b200" else []
                p = subprocess.run([exe, sub, "-k", "35", "-r", "43", "-b", bf_gb] + extra + [fa, vcf, prefix], capture_output=True, text=True)
                if extra:
                    r[sub + "_trace"] = traces(p.stderr)
                r[sub + "_wall_s"] = round(time.time() - t, 3)
                assert p.returncode == 0, p.stderr[-2000:]
                r[sub + "_phases_s"] = phases(p.stderr)
                if sub == "call":
                    outs[name] = p.stdout
            os.remove(vcf + ".c43.k35.malvax.zst")
            res[name] = r
        if ours_only:
            print(json.dumps(res, indent=1))
            return
        res["outputs_identical"] = outs["reference_cpu"] == outs["malva_b200"]
        a, b = res["reference_cpu"], res["malva_b200"]
        scan_ref = a["call_phases_s"].get("BF weights created")
        scan_ours = b["call_phases_s"].get("BF weights created")
        res["summary"] = {
            "total_speedup": round((a["index_wall_s"] + a["call_wall_s"]) / (b["index_wall_s"] + b["call_wall_s"]), 1),
            "kmc_scan_s": [scan_ref, scan_ours],
            "reference_pass_s": [a["index_phases_s"].get("Reference BF creation complete"),
                                 b["index_phases_s"].get("Reference BF creation complete")],
        }
        print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
