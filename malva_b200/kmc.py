"""Host-side k-mer utilities: 2-bit packing and a KMC-database reader/writer.

Reference call sites served: the KMC listing loop ``main.cpp:482-490`` (reader); the writer produces the
databases the tests and goldens feed to both programs.

The KMC API itself is third party (KMC >= 2.3, not vendored by the reference);
the on-disk layout implemented here is restated from the published format
description (KMC1 "version 0" / KMC2 "0x200") and is UNPINNED by any reference
test -- see DESIGN.md.

Packed k-mer word (the device format of ``mg_scan_sample_kmers``): A=0 C=1 G=2
T=3, first base in the most significant position, right-aligned in 128 bits and
stored as two little-endian u64 ``(lo, hi)``.  Integer order == lexicographic
order, so the canonical form is ``min(x, revcomp(x))``.
"""
from __future__ import annotations

import os
import struct
from typing import Iterable, Tuple

import numpy as np

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
_SYM = "ACGT"
_M64 = (1 << 64) - 1

KMER_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8")])


def pack_kmer(s: str) -> int:
    """ASCII ACGT string -> integer code (first base most significant)."""
    x = 0
    for ch in s:
        x = (x << 2) | _CODE[ch]
    return x


def unpack_kmer(x: int, k: int) -> str:
    return "".join(_SYM[(x >> (2 * (k - 1 - i))) & 3] for i in range(k))


def revcomp_int(x: int, k: int) -> int:
    r = 0
    for _ in range(k):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


def canonical_int(x: int, k: int) -> int:
    return min(x, revcomp_int(x, k))


def ints_to_packed(vals: Iterable[int]) -> np.ndarray:
    vals = list(vals)
    out = np.zeros(len(vals), dtype=KMER_DTYPE)
    out["lo"] = np.array([v & _M64 for v in vals], dtype=np.uint64)
    out["hi"] = np.array([(v >> 64) & _M64 for v in vals], dtype=np.uint64)
    return out


def packed_to_ints(arr: np.ndarray) -> list:
    lo = arr["lo"].tolist()
    hi = arr["hi"].tolist()
    return [(h << 64) | l for l, h in zip(lo, hi)]


def packed_to_strings(arr: np.ndarray, k: int) -> list:
    return [unpack_kmer(v, k) for v in packed_to_ints(arr)]


def _choose_lut_prefix_len(k: int, n: int) -> int:
    best = None
    for p in range(1, min(k, 13) + 1):   # (same rule as csrc/host/kmc_db.hpp: KmcWriter::choose_prefix_len)
        if (k - p) % 4:
            continue
        if best is None or (4 ** p) <= max(64, n):
            best = p
    if best is None:
        raise ValueError(f"no LUT prefix length with (k-p)%4==0 for k={k}")
    return best


def write_kmc_db(prefix: str, kmers: np.ndarray, counts: np.ndarray, k: int, *,
                 lut_prefix_len: int | None = None, version: int = 0x200, counter_size: int = 1,
                 min_count: int = 2, max_count: int = 255, signature_len: int = 5,
                 both_strands: bool = True) -> None:
    """Write ``<prefix>.kmc_pre`` / ``.kmc_suf`` holding the given sorted records."""
    vals = packed_to_ints(kmers)
    order = sorted(range(len(vals)), key=vals.__getitem__)
    vals = [vals[i] for i in order]
    cts = [int(counts[i]) for i in order]
    n = len(vals)
    p = lut_prefix_len if lut_prefix_len is not None else _choose_lut_prefix_len(k, n)
    if (k - p) % 4:
        raise ValueError("(k - lut_prefix_len) must be a multiple of 4")
    suf_syms = k - p
    suf_bytes = suf_syms // 4
    lut = [0] * (4 ** p + 1)
    suf = bytearray(b"KMCS")
    smask = (1 << (2 * suf_syms)) - 1
    for v, c in zip(vals, cts):
        lut[(v >> (2 * suf_syms)) + 1] += 1
        suf += (v & smask).to_bytes(suf_bytes, "big")
        suf += int(c).to_bytes(counter_size, "little")
    suf += b"KMCS"
    for i in range(1, len(lut)):
        lut[i] += lut[i - 1]
    # lut[i] = number of records whose prefix is < i ; lut[4^p] = n (guard)
    pre = bytearray(b"KMCP")
    pre += struct.pack(f"<{len(lut)}Q", *lut)
    if version == 0x200:
        pre += struct.pack(f"<{4 ** signature_len + 1}I", *([0] * (4 ** signature_len + 1)))
        hdr = struct.pack("<7IQBI", k, 0, counter_size, p, signature_len, min_count,
                          max_count & 0xFFFFFFFF, n, 0 if both_strands else 1, max_count >> 32)
    elif version == 0:
        hdr = struct.pack("<6IQBI", k, 0, counter_size, p, min_count, max_count & 0xFFFFFFFF, n,
                          0 if both_strands else 1, max_count >> 32)
    else:
        raise ValueError("version must be 0 or 0x200")
    hdr += b"\0" * (60 - len(hdr)) + struct.pack("<I", version)
    pre += hdr + struct.pack("<I", len(hdr)) + b"KMCP"
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".kmc_pre", "wb") as fh:
        fh.write(pre)
    with open(prefix + ".kmc_suf", "wb") as fh:
        fh.write(suf)


def write_kmc_db_binned(prefix: str, kmers: np.ndarray, counts: np.ndarray, k: int, bin_of, n_bins: int, *,
                        lut_prefix_len: int, counter_size: int = 1, min_count: int = 2, max_count: int = 255,
                        signature_len: int = 5) -> None:
    """KMC2 layout with SEVERAL bins, as the real tool writes it: one prefix LUT of 4^p entries per bin, the records
    of bin 0 first (sorted), then bin 1, ...; ``bin_of(value) -> bin`` stands in for KMC's minimiser-signature map.
    The listing order of such a database is NOT globally sorted (it is sorted within each bin)."""
    vals = packed_to_ints(kmers)
    cts = [int(c) for c in counts]
    p = lut_prefix_len
    suf_syms = k - p
    suf_bytes = suf_syms // 4
    smask = (1 << (2 * suf_syms)) - 1
    bins = [[] for _ in range(n_bins)]
    for v, c in sorted(zip(vals, cts)):
        bins[bin_of(v)].append((v, c))
    lut, suf, n = [], bytearray(b"KMCS"), 0
    for b in bins:
        per = [0] * (4 ** p)
        for v, _ in b:
            per[v >> (2 * suf_syms)] += 1
        for cnt in per:
            lut.append(n)
            n += cnt
        for v, c in b:
            suf += (v & smask).to_bytes(suf_bytes, "big") + int(c).to_bytes(counter_size, "little")
    lut.append(n)
    suf += b"KMCS"
    pre = bytearray(b"KMCP") + struct.pack(f"<{len(lut)}Q", *lut)
    pre += struct.pack(f"<{4 ** signature_len + 1}I", *([0] * (4 ** signature_len + 1)))
    hdr = struct.pack("<7IQB", k, 0, counter_size, p, signature_len, min_count, max_count & 0xFFFFFFFF, n, 0)
    hdr += b"\0" * (60 - len(hdr)) + struct.pack("<I", 0x200)
    pre += hdr + struct.pack("<I", len(hdr)) + b"KMCP"
    with open(prefix + ".kmc_pre", "wb") as fh:
        fh.write(pre)
    with open(prefix + ".kmc_suf", "wb") as fh:
        fh.write(suf)


def _parse_header(h: bytes, version: int):
    """Header of a .kmc_pre file as the KMC API reads it (CKMCFile::ReadParamsFrom_prefix_file_buf): kmer_length, mode,
    counter_size, lut_prefix_length, [signature_len: 0x200 only], min_count, max_count (low word), total_kmers,
    both_strands (one byte, stored inverted), then -- KMC 3 -- the high word of max_count; the rest is reserved."""
    if version == 0x200:
        k, _mode, csz, p, sig, minc, maxc, total = struct.unpack("<7IQ", h[:36])
        o = 36
    else:
        k, _mode, csz, p, minc, maxc, total = struct.unpack("<6IQ", h[:32])
        sig, o = 0, 32
    both = not (h[o] & 1)
    if len(h) >= o + 5 + 4 + 4 + 4:          # (room before version / header_offset / marker)
        maxc |= struct.unpack("<I", h[o + 1:o + 5])[0] << 32
    sigmap = (4 ** sig + 1) * 4 if version == 0x200 else 0
    return k, csz, p, sig, minc, maxc, total, both, sigmap


def read_kmc_db(prefix: str) -> Tuple[np.ndarray, np.ndarray, int]:
    """List a KMC database: (packed k-mers, u32 counts, k); count filter applied."""
    pre = open(prefix + ".kmc_pre", "rb").read()
    if pre[:4] != b"KMCP" or pre[-4:] != b"KMCP":
        raise ValueError("bad .kmc_pre markers")
    version, hoff = struct.unpack("<II", pre[-12:-4])
    if version not in (0, 0x200):
        raise ValueError(f"unsupported KMC version {version:#x}")
    h = pre[len(pre) - 8 - hoff:]
    k, csz, p, sig, minc, maxc, total, _both, sigmap = _parse_header(h, version)
    lut_bytes = len(pre) - 4 - 8 - hoff - sigmap
    lut = np.frombuffer(pre, dtype="<u8", count=lut_bytes // 8, offset=4)
    single = 4 ** p
    n_lut = (len(lut) // single) * single
    suf = open(prefix + ".kmc_suf", "rb").read()
    if suf[:4] != b"KMCS":
        raise ValueError("bad .kmc_suf marker")
    sb = (k - p) // 4
    rec = sb + csz
    vals, cts = [], []
    pi = 0
    for i in range(total):
        while pi + 1 < n_lut and lut[pi + 1] <= i:
            pi += 1
        o = 4 + i * rec
        c = int.from_bytes(suf[o + sb:o + rec], "little") if csz else 1
        if c < minc or c > maxc:
            continue
        vals.append(((pi % single) << (2 * (k - p))) | int.from_bytes(suf[o:o + sb], "big"))
        cts.append(c)
    return ints_to_packed(vals), np.array(cts, dtype=np.uint32), k


def open_kmc_db(prefix: str) -> dict:
    """Raw view of a KMC database for device-side decoding (``mg_kmc_open`` / ``mg_scan_kmc_records``):
    the prefix LUT, the header fields and the suffix records as a flat uint8 array (no per-record work)."""
    pre = open(prefix + ".kmc_pre", "rb").read()
    if pre[:4] != b"KMCP" or pre[-4:] != b"KMCP":
        raise ValueError("bad .kmc_pre markers")
    version, hoff = struct.unpack("<II", pre[-12:-4])
    if version not in (0, 0x200):
        raise ValueError(f"unsupported KMC version {version:#x}")
    h = pre[len(pre) - 8 - hoff:]
    k, csz, p, sig, minc, maxc, total, _both, sigmap = _parse_header(h, version)
    lut_bytes = len(pre) - 4 - 8 - hoff - sigmap
    lut = np.frombuffer(pre, dtype="<u8", count=lut_bytes // 8, offset=4)
    single = 4 ** p
    n_lut = (len(lut) // single) * single
    rec = (k - p) // 4 + csz
    suf = np.memmap(prefix + ".kmc_suf", dtype=np.uint8, mode="r")
    if bytes(suf[:4]) != b"KMCS":
        raise ValueError("bad .kmc_suf marker")
    records = suf[4:4 + total * rec]
    return dict(k=k, lut=np.ascontiguousarray(lut[:n_lut]), lut_prefix_len=p, counter_size=csz, min_count=minc,
                max_count=maxc, total=total, record_bytes=rec, records=records, both_strands=_both, version=version,
                signature_len=sig)
