"""Test-side access to the parity oracle (tests only -- never imported by pingpongpro_b200/).

* ``Oracle``       ctypes wrapper of oracle/_build/libppp_oracle.so (the C restatement, ppp_oracle.c)
* ``run_harness``  runs oracle/_ref/ppp_harness (the unmodified reference behind an include-harness)
* ``parse_dump``   reads the harness' binary dump into numpy arrays
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
MODES = {"weighted": 0, "discard": 1, "unique": 2}


def harness_path(max_overlap: int = 23) -> str:
    return os.path.join(REF_DIR, "ppp_harness" if max_overlap == 23 else f"ppp_harness_max{max_overlap}")


def ref_cli_path(max_overlap: int = 23) -> str:
    return os.path.join(REF_DIR, "pingpongpro_ref" if max_overlap == 23 else f"pingpongpro_ref_max{max_overlap}")


def have_ref(max_overlap: int = 23) -> bool:
    return os.path.exists(harness_path(max_overlap)) and os.path.exists(ref_cli_path(max_overlap))


def _lib() -> C.CDLL:
    path = os.environ.get("PPP_ORACLE_LIB") or os.path.join(ORACLE_DIR, "_build", "libppp_oracle.so")  # (override: a sanitizer build)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
    L = C.CDLL(path)
    L.ppo_new.restype = C.c_void_p
    L.ppo_new.argtypes = [C.c_int] * 5
    L.ppo_free.argtypes = [C.c_void_p]
    L.ppo_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    for f in ("ppo_build", "ppo_scan", "ppo_score"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.ppo_transposons.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ppo_n_stacks.restype = C.c_size_t
    L.ppo_n_stacks.argtypes = [C.c_void_p]
    L.ppo_get_stacks.argtypes = [C.c_void_p] * 6
    L.ppo_total.restype = C.c_double
    L.ppo_total.argtypes = [C.c_void_p]
    L.ppo_freq_len.restype = C.c_uint32
    L.ppo_freq_len.argtypes = [C.c_void_p]
    L.ppo_get_freq.argtypes = [C.c_void_p, C.c_void_p]
    L.ppo_max_height_score.restype = C.c_float
    L.ppo_max_height_score.argtypes = [C.c_void_p]
    L.ppo_get_grouped_pre.argtypes = [C.c_void_p, C.c_void_p]
    L.ppo_n_bins.restype = C.c_uint32
    L.ppo_n_bins.argtypes = [C.c_void_p]
    L.ppo_get_grouped_post.argtypes = [C.c_void_p, C.c_void_p]
    L.ppo_get_fdr_table.argtypes = [C.c_void_p, C.c_void_p]
    L.ppo_get_bin_map.argtypes = [C.c_void_p, C.c_void_p]
    L.ppo_n_sigs.restype = C.c_size_t
    L.ppo_n_sigs.argtypes = [C.c_void_p]
    L.ppo_get_sigs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


_L = None

TP_DTYPE = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4"), ("p", "<f4"), ("q", "<f4"),
                     ("reads_plus", "<f4"), ("reads_minus", "<f4"), ("hist", "<f4", (64,)), ("z64", "<f8"), ("p64", "<f8")])
IV_DTYPE = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4")])


class Oracle:
    """CPU restatement of the hot path (oracle/ppp_oracle.c). reads_per_contig: list of READ_DTYPE arrays."""

    def __init__(self, reads_per_contig, mode="weighted", min_ov=3, max_ov=23, pp_ov=10):
        global _L
        if _L is None:
            _L = _lib()
        self.L = _L
        self.W = max_ov - min_ov + 1
        self.h = self.L.ppo_new(len(reads_per_contig), min_ov, max_ov, pp_ov, MODES[mode])
        for c, r in enumerate(reads_per_contig):
            r = np.ascontiguousarray(r)
            if len(r):
                self.L.ppo_add(self.h, c, r.ctypes.data, len(r))
        self.L.ppo_build(self.h)
        self._scanned = self._scored = False

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ppo_free(self.h)
            self.h = None

    def stacks(self):
        n = self.L.ppo_n_stacks(self.h)
        strand, contig, key, a10 = (np.empty(n, np.uint32) for _ in range(4))
        reads = np.empty(n, np.float32)
        self.L.ppo_get_stacks(self.h, strand.ctypes.data, contig.ctypes.data, key.ctypes.data, reads.ctypes.data, a10.ctypes.data)
        return dict(strand=strand, contig=contig, key=key, reads=reads, a10=a10)

    def total(self) -> float:
        return float(self.L.ppo_total(self.h))

    def scan(self):
        if not self._scanned:
            self.L.ppo_scan(self.h)
            self._scanned = True
        return self

    def score(self):
        self.scan()
        if not self._scored:
            self.L.ppo_score(self.h)
            self._scored = True
        return self

    def freq(self) -> np.ndarray:
        self.scan()
        out = np.empty(self.L.ppo_freq_len(self.h), np.float32)
        self.L.ppo_get_freq(self.h, out.ctypes.data)
        return out

    def grouped_pre(self) -> np.ndarray:
        self.scan()
        out = np.empty((self.W, 1000, 2, 2), np.float32)
        self.L.ppo_get_grouped_pre(self.h, out.ctypes.data)
        return out

    def n_bins(self) -> int:
        self.score()
        return int(self.L.ppo_n_bins(self.h))

    def grouped_post(self) -> np.ndarray:
        out = np.empty((self.W, self.n_bins(), 2, 2), np.float32)
        self.L.ppo_get_grouped_post(self.h, out.ctypes.data)
        return out

    def fdr_table(self) -> np.ndarray:
        out = np.empty((self.W, self.n_bins(), 2, 2), np.float32)
        self.L.ppo_get_fdr_table(self.h, out.ctypes.data)
        return out

    def bin_map(self) -> np.ndarray:
        self.score()
        out = np.empty(1000, np.uint32)
        self.L.ppo_get_bin_map(self.h, out.ctypes.data)
        return out

    def sigs(self):
        """Signatures in the reference's list order (overlap, contig, position)."""
        self.score()
        n = self.L.ppo_n_sigs(self.h)
        ints = np.empty((n, 7), np.uint32)
        flts = np.empty((n, 3), np.float32)
        self.L.ppo_get_sigs(self.h, ints.ctypes.data, flts.ctypes.data)
        return dict(ov=ints[:, 0], contig=ints[:, 1], pos=ints[:, 2], bin_pre=ints[:, 3], bin=ints[:, 4], local=ints[:, 5], bias=ints[:, 6],
                    rp=flts[:, 0], rm=flts[:, 1], fdr=flts[:, 2])

    def transposons(self, intervals: np.ndarray) -> np.ndarray:
        """intervals: IV_DTYPE array. Returns TP_DTYPE array in (contig, start, end) order."""
        self.score()
        iv = np.ascontiguousarray(intervals, dtype=IV_DTYPE)
        out = np.zeros(len(iv), TP_DTYPE)
        self.L.ppo_transposons(self.h, iv.ctypes.data, len(iv), out.ctypes.data)
        return out


def run_harness(out_path, inputs, mode="weighted", min_len=24, max_len=32, annots=(), predict=0, max_overlap=23):
    cmd = [harness_path(max_overlap), "dump", out_path, "-m", mode, "-l", str(min_len), "-L", str(max_len)]
    for a in annots:
        cmd += ["-t", a]
    if predict:
        cmd += ["-T", str(predict)]
    for i in inputs:
        cmd += ["-i", i]
    subprocess.check_call(cmd)
    return parse_dump(out_path)


def parse_dump(path) -> dict:
    data = open(path, "rb").read()
    off = 0
    out = {}
    W = None
    while off < len(data):
        tag = data[off:off + 4].decode()
        n = int(np.frombuffer(data, "<u8", 1, off + 4)[0])
        off += 12
        if tag == "HDR ":
            h = np.frombuffer(data, "<u4", 3, off)
            out["min_ov"], out["max_ov"], out["pp_ov"] = (int(x) for x in h)
            W = out["max_ov"] - out["min_ov"] + 1
            off += 12
        elif tag == "STAK":
            dt = np.dtype([("strand", "<u4"), ("contig", "<u4"), ("key", "<u4"), ("reads", "<f4"), ("a10", "<u4")])
            out["stacks"] = np.frombuffer(data, dt, n, off)
            off += n * dt.itemsize
        elif tag == "TOTL":
            out["total"] = float(np.frombuffer(data, "<f8", 1, off)[0])
            off += 8
        elif tag == "FREQ":
            dt = np.dtype([("height", "<u4"), ("freq", "<f4")])
            out["freq"] = np.frombuffer(data, dt, n, off)
            off += n * dt.itemsize
        elif tag in ("GRP0", "GRP1"):
            cnt = W * n * 4
            out["grouped_pre" if tag == "GRP0" else "grouped_post"] = np.frombuffer(data, "<f4", cnt, off).reshape(W, n, 2, 2)
            off += cnt * 4
        elif tag == "SIGS":
            dt = np.dtype([("ov", "<u4"), ("contig", "<u4"), ("pos", "<u4"), ("bin_pre", "<u4"), ("bin", "<u4"), ("local", "<u4"), ("bias", "<u4"),
                           ("rp", "<f4"), ("rm", "<f4"), ("fdr", "<f4")])
            out["sigs"] = np.frombuffer(data, dt, n, off)
            off += n * dt.itemsize
        elif tag in ("TRSP", "PRED"):
            dt = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4"), ("p", "<f4"), ("q", "<f4"),
                           ("reads_plus", "<f4"), ("reads_minus", "<f4"), ("hist", "<f4", (W,))])
            out["transposons" if tag == "TRSP" else "predicted"] = np.frombuffer(data, dt, n, off)
            off += n * dt.itemsize
        else:
            raise ValueError(f"unknown section {tag!r}")
    return out
