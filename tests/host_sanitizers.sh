#!/bin/bash
# Sanitizer builds of the C++ host (malva_b200/csrc/host) run over the CPU-only sub-commands: AddressSanitizer +
# UBSan, then ThreadSanitizer (reader thread | decode-ahead | parallel stages).  Not part of pytest (each build takes
# about a minute); last run: clean on tests/golden/sars (27,934-sample panel, BGZF) and a three-batch 32-sample file.
#   tests/host_sanitizers.sh <reference.fa> <variants.vcf[.gz]> [flags for `signatures`, e.g. -1]
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
fa=$1; vcf=$2; shift 2
out=$(mktemp -d)
libs=("${root}/malva_b200/libmalva_gpu.so" /usr/lib/x86_64-linux-gnu/libzstd.so.1 -lz -lpthread "-Wl,-rpath,${root}/malva_b200")
want=$("${root}/malva_b200/malva-geno" signatures "$@" "${fa}" "${vcf}" 2>/dev/null | md5sum)
for san in address,undefined thread; do
  g++ -O1 -g -fsanitize=${san} -fno-omit-frame-pointer -std=c++17 -o "${out}/geno" "${root}/malva_b200/csrc/host/malva_geno.cpp" "${libs[@]}"
  got=$(ASAN_OPTIONS=detect_leaks=0 "${out}/geno" signatures --threads 4 "$@" "${fa}" "${vcf}" 2>"${out}/err" | md5sum)
  if [[ "${got}" != "${want}" ]] || grep -q "ERROR: AddressSanitizer\|runtime error\|WARNING: ThreadSanitizer" "${out}/err"; then
    echo "-fsanitize=${san}: FAILED (see ${out}/err)"; exit 1
  fi
  if [[ ${san} != thread ]]; then
    ASAN_OPTIONS=detect_leaks=0 "${out}/geno" container-selftest
    ASAN_OPTIONS=detect_leaks=0 "${out}/geno" format-selftest
  fi
  echo "-fsanitize=${san}: clean, output identical"
done
rm -rf "${out}"
