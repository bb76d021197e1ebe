"""ctypes bindings of the host-side native libraries (no CUDA):

* ``libppp_host.so``  -- SAM/BAM decode + per-record feature extraction (include/ppp_host.h),
  the host half of countReadsInBamFile (pingpongpro.cpp:402-484).
* ``libppp_synth.so`` -- the seeded synthetic read model (tools/synth_core.h).

Both return reads in the packed 8-byte ``ppp_read`` layout (include/ppp_types.h), exposed here as
numpy structured arrays of dtype ``READ_DTYPE``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

READ_DTYPE = np.dtype([("key", "<u4"), ("meta", "<u4")])
MODES = {"weighted": 0, "discard": 1, "unique": 2}

READ_MINUS = 1
READ_A10 = 2
READ_NH_SHIFT = 2


def pack_reads(key, minus, a10, nh) -> np.ndarray:
    """Builds ppp_read records from column arrays."""
    key = np.asarray(key, dtype=np.uint32)
    out = np.empty(key.shape[0], dtype=READ_DTYPE)
    out["key"] = key
    out["meta"] = (np.asarray(nh, dtype=np.uint32) << READ_NH_SHIFT) | (np.asarray(a10, dtype=np.uint32) << 1) | np.asarray(minus, dtype=np.uint32)
    return out


def _load(name: str, builder) -> C.CDLL:
    path = os.path.join(_build.LIB, name)
    if not os.path.exists(path):
        builder()
    return C.CDLL(path)


_host = None
_synth = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        L = _load("libppp_host.so", _build.build_host)
        L.ppp_fe_decode.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
        L.ppp_fe_decode.restype = C.c_int
        L.ppp_fe_decode_mt.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ppp_fe_decode_mt.restype = C.c_int
        L.ppp_fe_free.argtypes = [C.c_void_p]
        L.ppp_fe_n_contigs.argtypes = [C.c_void_p]
        L.ppp_fe_contig_name.argtypes = [C.c_void_p, C.c_int]
        L.ppp_fe_contig_name.restype = C.c_char_p
        L.ppp_fe_contig_length.argtypes = [C.c_void_p, C.c_int]
        L.ppp_fe_contig_length.restype = C.c_uint64
        L.ppp_fe_n_reads.argtypes = [C.c_void_p, C.c_int]
        L.ppp_fe_n_reads.restype = C.c_size_t
        L.ppp_fe_reads.argtypes = [C.c_void_p, C.c_int]
        L.ppp_fe_reads.restype = C.c_void_p
        L.ppp_fe_total_weight.argtypes = [C.c_void_p]
        L.ppp_fe_total_weight.restype = C.c_double
        L.ppp_fe_n_records.argtypes = [C.c_void_p]
        L.ppp_fe_n_records.restype = C.c_uint64
        L.ppp_fe_rejected_guesses.argtypes = [C.c_void_p]
        L.ppp_fe_rejected_guesses.restype = C.c_uint64
        L.ppp_fe_n_kept.argtypes = [C.c_void_p]
        L.ppp_fe_n_kept.restype = C.c_uint64
        L.ppp_annot_read.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]
        L.ppp_annot_read.restype = C.c_int
        L.ppp_annot_free.argtypes = [C.c_void_p]
        L.ppp_annot_count.argtypes = [C.c_void_p]
        L.ppp_annot_count.restype = C.c_size_t
        L.ppp_annot_intervals.argtypes = [C.c_void_p, C.c_void_p]
        L.ppp_annot_identifier.argtypes = [C.c_void_p, C.c_size_t]
        L.ppp_annot_identifier.restype = C.c_char_p
        L.ppp_annot_n_names.argtypes = [C.c_void_p]
        L.ppp_annot_name.argtypes = [C.c_void_p, C.c_int]
        L.ppp_annot_name.restype = C.c_char_p
        _host = L
    return _host


ANNOT_DTYPE = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4")])


def read_annotation(paths, contig_names):
    """Transposon intervals of annotation files (.bed .csv .gff .gtf .tsv) with the reference's parsing rules
    (pingpongpro.cpp:993-1175).  Returns (intervals[ANNOT_DTYPE] in reference order, identifiers, extended name store)."""
    L = host_lib()
    pa = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    na = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    h = C.c_void_p()
    if L.ppp_annot_read(pa, len(paths), na, len(contig_names), C.byref(h)) != 0:
        raise OSError("Failed to open transposon file")
    try:
        n = L.ppp_annot_count(h)
        iv = np.zeros(n, ANNOT_DTYPE)
        if n:
            L.ppp_annot_intervals(h, iv.ctypes.data)
        ids = [L.ppp_annot_identifier(h, i).decode(errors="replace") for i in range(n)]
        names = [L.ppp_annot_name(h, i).decode(errors="replace") for i in range(L.ppp_annot_n_names(h))]
        return iv, ids, names
    finally:
        L.ppp_annot_free(h)


def predict_transposons(contigs, positions, contig_names, rng: int):
    """-T: putative transposons from overlap-10 loci in (contig, position) order (pingpongpro.cpp:1326-1354).
    Returns (intervals[ANNOT_DTYPE], identifiers)."""
    L = host_lib()
    c = np.ascontiguousarray(contigs, dtype=np.uint32)
    p = np.ascontiguousarray(positions, dtype=np.uint32)
    na = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
    h = C.c_void_p()
    L.ppp_annot_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    if L.ppp_annot_predict(c.ctypes.data, p.ctypes.data, len(c), na, len(contig_names), rng, C.byref(h)) != 0:
        raise ValueError("contig id out of range")
    try:
        n = L.ppp_annot_count(h)
        iv = np.zeros(n, ANNOT_DTYPE)
        if n:
            L.ppp_annot_intervals(h, iv.ctypes.data)
        return iv, [L.ppp_annot_identifier(h, i).decode(errors="replace") for i in range(n)]
    finally:
        L.ppp_annot_free(h)


class DecodedFile:
    """Result of decoding one SAM/BAM file: contig table + packed reads per contig (file order)."""

    def __init__(self, names, lengths, reads, total_weight, n_records, n_kept):
        self.names = names
        self.lengths = lengths
        self.reads = reads
        self.total_weight = total_weight
        self.n_records = n_records
        self.n_kept = n_kept
        self.rejected_guesses = 0  # BAM, threads > 1: runs of blocks decoded again because the guessed record boundary was wrong


def decode_file(path: str, min_len: int = 24, max_len: int = 32, mode: str = "weighted", threads: int = 1) -> DecodedFile:
    L = host_lib()
    h = C.c_void_p()
    rc = L.ppp_fe_decode_mt(os.fsencode(path), min_len, max_len, MODES[mode], threads, C.byref(h))
    if rc == 1:
        raise OSError(f"Failed to open input file: {path}")
    if rc != 0:
        raise ValueError("Failed to read record")
    try:
        n = L.ppp_fe_n_contigs(h)
        names = [L.ppp_fe_contig_name(h, c).decode() for c in range(n)]
        lengths = [int(L.ppp_fe_contig_length(h, c)) for c in range(n)]
        reads = []
        for c in range(n):
            k = L.ppp_fe_n_reads(h, c)
            if k:
                buf = (C.c_char * (8 * k)).from_address(L.ppp_fe_reads(h, c))
                reads.append(np.frombuffer(buf, dtype=READ_DTYPE).copy())
            else:
                reads.append(np.empty(0, dtype=READ_DTYPE))
        d = DecodedFile(names, lengths, reads, float(L.ppp_fe_total_weight(h)), int(L.ppp_fe_n_records(h)), int(L.ppp_fe_n_kept(h)))
        d.rejected_guesses = int(L.ppp_fe_rejected_guesses(h))
        return d
    finally:
        L.ppp_fe_free(h)


def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        L = _load("libppp_synth.so", _build.build_synth)
        L.ppp_synth_new.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_int]
        L.ppp_synth_new.restype = C.c_void_p
        L.ppp_synth_free.argtypes = [C.c_void_p]
        L.ppp_synth_n_contigs.argtypes = [C.c_void_p]
        L.ppp_synth_contig_length.argtypes = [C.c_void_p, C.c_int]
        L.ppp_synth_contig_length.restype = C.c_uint64
        L.ppp_synth_contig_name.argtypes = [C.c_void_p, C.c_int]
        L.ppp_synth_contig_name.restype = C.c_char_p
        L.ppp_synth_n_intervals.argtypes = [C.c_void_p]
        L.ppp_synth_n_intervals.restype = C.c_uint64
        L.ppp_synth_intervals.argtypes = [C.c_void_p, C.c_void_p]
        L.ppp_synth_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.ppp_synth_generate.restype = C.c_int
        _synth = L
    return _synth


class SynthModel:
    """Seeded synthetic piRNA-seq data set."""

    def __init__(self, genome: str | None = "dm6", n_reads: int = 1_000_000, seed: int = 1, nh1: bool = False, lengths=None):
        L = synth_lib()
        if genome:
            self._h = L.ppp_synth_new(genome.encode(), 0, None, seed, n_reads, int(nh1))
        else:
            arr = (C.c_uint64 * len(lengths))(*lengths)
            self._h = L.ppp_synth_new(None, len(lengths), arr, seed, n_reads, int(nh1))
        if not self._h:
            raise ValueError(f"unknown genome preset {genome!r}")
        self.n_reads = n_reads
        n = L.ppp_synth_n_contigs(self._h)
        self.names = [L.ppp_synth_contig_name(self._h, c).decode() for c in range(n)]
        self.lengths = [int(L.ppp_synth_contig_length(self._h, c)) for c in range(n)]
        self.genome_size = sum(self.lengths)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            synth_lib().ppp_synth_free(h)

    def intervals(self) -> np.ndarray:
        """Transposon-like intervals as uint32[n, 3] = (contig, start, end-exclusive)."""
        L = synth_lib()
        n = L.ppp_synth_n_intervals(self._h)
        out = np.empty((n, 3), dtype=np.uint32)
        L.ppp_synth_intervals(self._h, out.ctypes.data)
        return out

    def generate(self, i0: int = 0, i1: int | None = None, g_lo: int = 0, g_hi: int | None = None, halo: int = 0, min_len: int = 24,
                 max_len: int = 32, mode: str = "weighted", threads: int = 0, alloc=None) -> list[np.ndarray]:
        """Packed reads per contig (index order) of read indices [i0, i1) whose 5'-end key lies in the
        global coordinate range [g_lo, g_hi) (+ `halo` further positions for minus-strand keys)."""
        L = synth_lib()
        i1 = self.n_reads if i1 is None else i1
        g_hi = self.genome_size + 64 if g_hi is None else g_hi
        n = len(self.lengths)
        counts = (C.c_uint64 * n)()
        args = (self._h, i0, i1, g_lo, g_hi, g_hi + halo, min_len, max_len, MODES[mode], threads)
        L.ppp_synth_generate(*args, counts, None)
        alloc = alloc or (lambda k: np.empty(k, dtype=READ_DTYPE))
        outs = [alloc(int(counts[c])) for c in range(n)]
        ptrs = (C.c_void_p * n)(*[o.ctypes.data if len(o) else None for o in outs])
        L.ppp_synth_generate(*args, counts, ptrs)
        return outs
