#!/usr/bin/env python
"""Host VCF pass on CPU only: `malva-geno signatures` (read | decode | group | enumerate, no device, nothing printed)
over tests/synth_fast.py inputs and over BASELINE config 1's panel; per-stage times from --trace, wall and CPU time
per run.  `MALVA_GENERAL_DECODE=1` runs are the reader without its short cuts.

    python profiles/host_pass_bench.py [Mbp=20] [samples=32] > profiles/round2_host_pass_cpu.json
"""
import json
import os
import re
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_fast  # noqa: E402
from malva_b200 import build as mbuild  # noqa: E402


def run(cli, fa, vcf, threads, flags=(), general=False, repeat=3):
    env = dict(os.environ, MALVA_SIGNATURES_QUIET="1")
    env.pop("MALVA_GENERAL_DECODE", None)
    if general:
        env["MALVA_GENERAL_DECODE"] = "1"
    best = None
    for _ in range(repeat):
        r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
        t0 = time.time()
        p = subprocess.run([cli, "signatures", "--trace", "--threads", str(threads), *flags, fa, vcf], env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
        wall = time.time() - t0
        r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
        m = re.findall(r"read ([\d.]+) \+ decode ([\d.]+) \+ group ([\d.]+) ms, enumerate ([\d.]+) ms", p.stderr)[-1]
        rec = {"threads": threads, "general_decode": general, "wall_s": round(wall, 3),
               "cpu_s": round(r1.ru_utime - r0.ru_utime + r1.ru_stime - r0.ru_stime, 3),
               "stage_ms_cumulative": dict(zip(("read", "decode", "group", "enumerate"), map(float, m)))}
        if best is None or rec["wall_s"] < best["wall_s"]:
            best = rec
    return best


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cli = mbuild.build_cli()
    n_cpu = os.cpu_count()
    out = {"host_cpus": n_cpu, "note": "best of 3 runs each; stages overlap (reader thread, decode-ahead), their times are cumulative per stage"}
    with tempfile.TemporaryDirectory() as d:
        fa, vcf, _, n = synth_fast.build(d, int(mbp * 1e6), n_samples=n_samples)
        out["synthetic"] = {"reference_bases": int(mbp * 1e6), "variants": n, "panel_samples": n_samples,
                            "runs": [run(cli, fa, vcf, 1), run(cli, fa, vcf, n_cpu), run(cli, fa, vcf, n_cpu, general=True)]}
        for r in out["synthetic"]["runs"]:
            r["variants_per_s"] = round(n / r["wall_s"])
    sars = os.path.join(ROOT, "tests", "golden", "sars")
    fa, vcf = os.path.join(sars, "reference_sarsCov2.fasta"), os.path.join(sars, "sars_cov2.vcf.gz")
    out["config1_sars_cov2_panel"] = {"records": 15154, "panel_samples": 27934,
                                      "runs": [run(cli, fa, vcf, 1, ["-1"]), run(cli, fa, vcf, n_cpu, ["-1"]),
                                               run(cli, fa, vcf, n_cpu, ["-1"], general=True)]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
