"""SHA-256 digests of the hot path's intermediates, byte-compatible with `oracle/_ref/ppp_harness digest`
(oracle/harness.cpp: runDigest), so that full-size results of the CUDA path can be compared with the unmodified
reference without shipping gigabytes of golden data (tests/golden/fullsize/*.json, made by
tests/golden/make_fullsize_digests.py on the CPU container).

Byte layouts (little endian):
  stacks            {u32 strand, u32 contig, u32 key, f32 reads, u32 a10}      in (strand, contig, key) order
  freq              {u32 rounded height, f32 frequency}                        ascending, non-zero entries
  grouped_pre/post  f32 [overlap][bin][bias][local]
  signatures        ppp_pair records (api.PAIR_DTYPE)                          in (contig, position, overlap) order
  loci[s]           ppp_locus records (api.LOCUS_DTYPE)                        in (contig, position) order
  transposon_sums   {u32 contig, strand, start, end; f32 R+, R-, hist[W]}      in (contig, start, end) order
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize")
STACK_DTYPE = np.dtype([("strand", "<u4"), ("contig", "<u4"), ("key", "<u4"), ("reads", "<f4"), ("a10", "<u4")])
PAIR_DTYPE = np.dtype([("contig", "<u4"), ("position", "<u4"), ("overlap", "<u4"), ("height_score_bin", "<u4"), ("collapsed_bin", "<u4"),
                       ("local_bin", "<u4"), ("bias_bin", "<u4"), ("reads_plus", "<f4"), ("reads_minus", "<f4"), ("fdr", "<f4")])
LOCUS_DTYPE = np.dtype([("contig", "<u4"), ("position", "<u4"), ("fdr", "<f4"), ("reads_plus", "<f4"), ("reads_minus", "<f4")])


def load(case: str) -> dict:
    with open(os.path.join(GOLDEN, case + ".json")) as f:
        return json.load(f)


def load_pq(case: str) -> np.ndarray:
    return np.load(os.path.join(GOLDEN, case + "_pq.npy"))


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(memoryview(a).cast("B"))
    return h.hexdigest()


class StackHasher:
    """Streams stacks contig by contig (strand-major, like the reference's map walk)."""

    def __init__(self):
        self.h = hashlib.sha256()
        self.n = 0

    def add_dense(self, strand: int, contig: int, reads: np.ndarray, a10: np.ndarray, first_key: int = 0):
        """reads/a10: one value per key of the contig (0 = no stack)."""
        idx = np.flatnonzero(reads)
        rec = np.empty(len(idx), STACK_DTYPE)
        rec["strand"], rec["contig"], rec["key"], rec["reads"], rec["a10"] = strand, contig, idx + first_key, reads[idx], a10[idx]
        self.h.update(memoryview(rec).cast("B"))
        self.n += len(idx)

    def add_records(self, strand, contig, key, reads, a10):
        rec = np.empty(len(key), STACK_DTYPE)
        rec["strand"], rec["contig"], rec["key"], rec["reads"], rec["a10"] = strand, contig, key, reads, a10
        self.h.update(memoryview(rec).cast("B"))
        self.n += len(key)

    def hexdigest(self):
        return self.h.hexdigest()


def digest_freq(counts: np.ndarray) -> tuple[int, str]:
    """counts[h] = number of stacks with rounded height h (exact integers); the reference's counter is float(min(n, 2^24))."""
    counts = np.asarray(counts)
    hts = np.flatnonzero(counts)
    rec = np.empty(len(hts), np.dtype([("h", "<u4"), ("f", "<f4")]))
    rec["h"] = hts
    rec["f"] = counts[hts] if counts.dtype.kind == "f" else np.minimum(counts[hts], 1 << 24).astype(np.float32)
    return len(hts), sha(rec)


def digest_grouped_counts(grouped: np.ndarray) -> str:
    """grouped: exact integer counts [W][1000][2][2] -> the float counters of pingpongpro.cpp:605."""
    grouped = np.asarray(grouped)
    return sha(grouped.astype("<f4") if grouped.dtype.kind == "f" else np.minimum(grouped, 1 << 24).astype("<f4"))


def digest_signatures(sig: np.ndarray) -> str:
    """sig: PAIR_DTYPE array already in (contig, position, overlap) order."""
    assert sig.dtype == PAIR_DTYPE
    return sha(sig)


def digest_loci(loci: np.ndarray) -> str:
    assert loci.dtype == LOCUS_DTYPE
    return sha(loci)


def digest_transposon_sums(contig, strand, start, end, rp, rm, hist, W: int) -> str:
    n = len(contig)
    dt = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4"), ("rp", "<f4"), ("rm", "<f4"), ("hist", "<f4", (W,))])
    rec = np.empty(n, dt)
    rec["contig"], rec["strand"], rec["start"], rec["end"], rec["rp"], rec["rm"] = contig, strand, start, end, rp, rm
    rec["hist"] = np.asarray(hist)[:, :W]
    return sha(rec)
