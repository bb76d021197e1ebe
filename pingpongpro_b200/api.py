"""ctypes binding of the C ABI in include/ppp_b200.h (libppp_b200.so, hand-written sm_100a CUDA).

This is the host-side mirror of the seams the reference's main() calls (pingpongpro.cpp:1490, :1580, :1583,
:1584, :1588, :1606); method names follow the reference's functions:

    ctx = Context(contig_lengths, mode="weighted")
    ctx.count_reads(contig, reads)         # countReadsInBamFile accumulation, :458-488
    ctx.finish()
    freq, max_h = ctx.map_heights_to_scores()      # :502-509
    grouped = ctx.count_stacks_by_group()          # :523-624
    res = ctx.calculate_fdrs(min_stack_height=0)   # collapseBins :631-686 + calculateFDRs :695-754 + :903
    tps = ctx.find_suppressed_transposons(iv)      # :1195-1313

There is NO CPU fallback: loading fails loudly when the CUDA library is missing, and Context() raises
PppError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build
from .host import MODES, READ_DTYPE

MAX_OVERLAPS = 32
BINS = 1000

LOCUS_DTYPE = np.dtype([("contig", "<u4"), ("position", "<u4"), ("fdr", "<f4"), ("reads_plus", "<f4"), ("reads_minus", "<f4")])
PAIR_DTYPE = np.dtype([("contig", "<u4"), ("position", "<u4"), ("overlap", "<u4"), ("height_score_bin", "<u4"), ("collapsed_bin", "<u4"),
                       ("local_bin", "<u4"), ("bias_bin", "<u4"), ("reads_plus", "<f4"), ("reads_minus", "<f4"), ("fdr", "<f4")])
INTERVAL_DTYPE = np.dtype([("contig", "<u4"), ("strand", "<u4"), ("start", "<u4"), ("end", "<u4")])
TP_RESULT_DTYPE = np.dtype([("histogram", "<f4", (MAX_OVERLAPS,)), ("reads_plus", "<f4"), ("reads_minus", "<f4"), ("z", "<f8"), ("p", "<f8"),
                            ("p_value", "<f4"), ("q_value", "<f4")], align=True)


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_contigs", C.c_uint32), ("contig_lengths", C.POINTER(C.c_uint64)), ("min_overlap", C.c_int32),
                ("max_overlap", C.c_int32), ("pingpong_overlap", C.c_int32), ("multi_hits", C.c_int32), ("range_begin", C.c_uint64),
                ("range_end", C.c_uint64), ("pair_capacity", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("stack_build_ms", C.c_float), ("scan_ms", C.c_float), ("lut_ms", C.c_float), ("bin_ms", C.c_float), ("score_ms", C.c_float),
                ("transposon_ms", C.c_float), ("launches", C.c_uint32), ("exchange_ms", C.c_float), ("pass_ms", C.c_float), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p)


class PppError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ppp error {code}: {message}")
        self.code = code


ERR_ARG, ERR_CUDA, ERR_STATE, ERR_RANGE, ERR_DEGENERATE, ERR_NOMEM = 1, 2, 3, 4, 5, 6

_lib = None


def lib_path() -> str:
    return os.path.join(_build.LIB, "libppp_b200.so")


def lib() -> C.CDLL:
    """Loads libppp_b200.so (building it in-tree with nvcc when absent). Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        _build.build_cuda()
    L = C.CDLL(lib_path())
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.ppp_last_error.restype = C.c_char_p
    L.ppp_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ppp_init.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ppp_warmup.argtypes = [C.c_int32]
    L.ppp_destroy.argtypes = [vp]
    L.ppp_destroy.restype = None
    L.ppp_reset.argtypes = [vp]
    L.ppp_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    L.ppp_set_rank.argtypes = [vp, C.c_int, C.c_int]
    L.ppp_nccl_unique_id.argtypes = [vp]
    L.ppp_nccl_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ppp_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.ppp_pinned_free.argtypes = [vp]
    L.ppp_reads_submit.argtypes = [vp, C.c_int32, vp, C.c_size_t, C.c_uint64]
    L.ppp_reads_submit_device.argtypes = [vp, C.c_int32, vp, C.c_size_t]
    L.ppp_reads_submit_packed.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_size_t]
    L.ppp_stacks_finish.argtypes = [vp]
    L.ppp_reads_foreign.argtypes = [vp, u64p]
    L.ppp_stacks_download.argtypes = [vp, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, vp, vp]
    L.ppp_height_hist.argtypes = [vp, C.POINTER(u64p), C.POINTER(C.c_uint32)]
    L.ppp_scan.argtypes = [vp, vp, u64p]
    L.ppp_score.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.ppp_pairs_download.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.ppp_transposons.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.ppp_timings_get.argtypes = [vp, C.POINTER(Timings)]
    L.ppp_timings_reset.argtypes = [vp]
    L.ppp_layout.argtypes = [vp, u64p, u64p]
    L.ppp_host_bin_thresholds.argtypes = [C.c_float, vp]
    L.ppp_emul_bin_thresholds.argtypes = [C.c_float, vp]
    L.ppp_emul_log10f.argtypes = [vp, C.c_size_t, vp]
    L.ppp_emul_fadd_chain.argtypes = [C.c_float, vp, C.c_size_t, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    L.ppp_pass_info.argtypes = [vp, C.POINTER(C.c_uint32), u64p, u64p, u64p]
    for name in ("ppp_device_count", "ppp_warmup", "ppp_init", "ppp_reset", "ppp_set_allreduce", "ppp_set_rank", "ppp_nccl_unique_id", "ppp_nccl_attach", "ppp_pinned_alloc", "ppp_pinned_free", "ppp_reads_submit",
                 "ppp_reads_submit_device", "ppp_reads_submit_packed", "ppp_stacks_finish", "ppp_reads_foreign", "ppp_stacks_download", "ppp_height_hist", "ppp_scan", "ppp_score",
                 "ppp_pairs_download", "ppp_transposons", "ppp_timings_get", "ppp_timings_reset", "ppp_layout", "ppp_host_bin_thresholds", "ppp_emul_bin_thresholds",
                 "ppp_emul_log10f", "ppp_emul_fadd_chain", "ppp_pass_info"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise PppError(rc, lib().ppp_last_error().decode(errors="replace"))


def bin_thresholds(max_freq: float) -> np.ndarray:
    """The 999 height-score bin thresholds for a maximum frequency (host only; pingpongpro.cpp:551, :591-594)."""
    out = np.empty(999, dtype=np.float32)
    _check(lib().ppp_host_bin_thresholds(C.c_float(max_freq), out.ctypes.data_as(C.c_void_p)))
    return out


def emulated_bin_thresholds(max_freq: float) -> np.ndarray:
    """The same table from the restatement of glibc's log10f that the device evaluates (csrc/log10f_emul.h), run on the host."""
    out = np.empty(999, dtype=np.float32)
    _check(lib().ppp_emul_bin_thresholds(C.c_float(max_freq), out.ctypes.data_as(C.c_void_p)))
    return out


def emulated_log10f(x) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _check(lib().ppp_emul_log10f(x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p)))
    return out


def emulated_fadd_chain(h0: float, weights, piece: int):
    """h0 + w0 + w1 + ... in float32, left to right, through the composed per-binade maps of csrc/fadd_chain.h (host).
    Returns (sum, number of additions that were executed for real)."""
    w = np.ascontiguousarray(weights, dtype=np.float32)
    out, real = C.c_float(0), C.c_uint64(0)
    _check(lib().ppp_emul_fadd_chain(C.c_float(h0), w.ctypes.data_as(C.c_void_p), len(w), piece, C.byref(out), C.byref(real)))
    return np.float32(out.value), int(real.value)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().ppp_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class PinnedBuffer:
    """Page-locked host memory from ppp_pinned_alloc, viewed as a numpy array of ppp_read records."""

    def __init__(self, n_records: int):
        self.ptr = C.c_void_p()
        self.n = n_records
        _check(lib().ppp_pinned_alloc(max(8, 8 * n_records), C.byref(self.ptr)))
        buf = (C.c_char * (8 * n_records)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=READ_DTYPE)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ppp_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedReads:
    """The reads of one contig in the 5-byte upload form of ppp_reads_submit_packed, in page-locked host memory: u32 keys,
    u8 meta (strand, A10, 6-bit index into the NH table of the piece) -- cut into pieces of at most 64 distinct NH values."""

    def __init__(self, reads: np.ndarray):
        self.n = len(reads)
        self._keys = self._meta = None
        self.pieces = []  # (first, count, nh_table[64])
        if not self.n:
            return
        self._keys = _pinned(4 * self.n)
        self._meta = _pinned(self.n)
        keys = np.frombuffer((C.c_char * (4 * self.n)).from_address(self._keys.value), dtype=np.uint32)
        meta = np.frombuffer((C.c_char * self.n).from_address(self._meta.value), dtype=np.uint8)
        keys[:] = reads["key"]
        nh = reads["meta"] >> np.uint32(2)
        first = 0
        while first < self.n:  # pieces of at most 64 distinct NH values (nearly always one piece)
            last = self.n
            table = np.unique(nh[first:last])
            while len(table) > 64:
                last = first + max(1, (last - first) // 2)
                table = np.unique(nh[first:last])
            code = np.searchsorted(table, nh[first:last]).astype(np.uint8)
            meta[first:last] = (reads["meta"][first:last] & np.uint32(3)).astype(np.uint8) | (code << np.uint8(2))
            t = np.ones(64, np.uint32)
            t[: len(table)] = table
            self.pieces.append((first, last - first, t))
            first = last

    def free(self):
        for p in (self._keys, self._meta):
            if p:
                lib().ppp_pinned_free(p)
        self._keys = self._meta = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _pinned(nbytes: int) -> C.c_void_p:
    p = C.c_void_p()
    _check(lib().ppp_pinned_alloc(max(8, nbytes), C.byref(p)))
    return p


class ScoreResult:
    def __init__(self, n_bins, bin_map, cells, fdr_table, loci):
        self.n_bins = n_bins
        self.bin_map = bin_map
        self.cells = cells
        self.fdr_table = fdr_table
        self.loci = loci


class Context:
    """One GPU context = one shard of the genome (ppp_ctx)."""

    def __init__(self, contig_lengths, mode="weighted", min_overlap=3, max_overlap=23, pingpong_overlap=10, device=0, range_begin=0, range_end=0,
                 pair_capacity=0):
        L = lib()
        self._lengths = (C.c_uint64 * len(contig_lengths))(*[int(x) for x in contig_lengths])
        cfg = Config(device, len(contig_lengths), self._lengths, min_overlap, max_overlap, pingpong_overlap, MODES[mode] if isinstance(mode, str) else mode,
                     range_begin, range_end, pair_capacity)
        self._h = C.c_void_p()
        _check(L.ppp_init(C.byref(cfg), C.byref(self._h)))
        self.W = max_overlap - min_overlap + 1
        self.min_overlap, self.max_overlap, self.pingpong_overlap = min_overlap, max_overlap, pingpong_overlap
        self.n_contigs = len(contig_lengths)
        self._hook = None

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().ppp_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- stack store ---------------------------------------------------------------------------------
    def reset(self):
        _check(lib().ppp_reset(self._h))

    def count_reads(self, contig: int, reads, first_record_index: int = 0):
        """reads: READ_DTYPE array (pageable or a PinnedBuffer.array) of ONE contig in file order."""
        if isinstance(reads, PinnedBuffer):
            ptr, n = reads.ptr, reads.n
        else:
            reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
            ptr, n = C.c_void_p(reads.ctypes.data), len(reads)
        _check(lib().ppp_reads_submit(self._h, contig, ptr, n, first_record_index))

    def count_reads_packed(self, contig: int, packed: "PackedReads"):
        """The same from the 5-byte upload form (ppp_reads_submit_packed)."""
        for first, count, table in packed.pieces:
            _check(lib().ppp_reads_submit_packed(self._h, contig, C.c_void_p(packed._keys.value + 4 * first), C.c_void_p(packed._meta.value + first),
                                                 table.ctypes.data_as(C.c_void_p), count))

    def count_reads_raw(self, contig: int, host_ptr: int, n: int):
        _check(lib().ppp_reads_submit(self._h, contig, C.c_void_p(host_ptr), n, 0))

    def count_reads_device(self, contig: int, device_ptr: int, n: int):
        _check(lib().ppp_reads_submit_device(self._h, contig, C.c_void_p(device_ptr), n))

    def finish(self):
        _check(lib().ppp_stacks_finish(self._h))

    def foreign_reads(self) -> int:
        n = C.c_uint64(0)
        _check(lib().ppp_reads_foreign(self._h, C.byref(n)))
        return n.value

    def stacks(self, contig: int, strand: int, begin: int, count: int):
        reads = np.empty(count, np.float32)
        a10 = np.empty(count, np.uint8)
        _check(lib().ppp_stacks_download(self._h, contig, strand, begin, count, reads.ctypes.data, a10.ctypes.data))
        return reads, a10

    # -- scan / score --------------------------------------------------------------------------------
    def map_heights_to_scores(self, want_table=True, want_max_height=True):
        """want_table=False, want_max_height=False only enqueues the kernels (no host synchronisation): see ppp_height_hist."""
        fp = C.POINTER(C.c_uint64)()
        mh = C.c_uint32(0)
        _check(lib().ppp_height_hist(self._h, C.byref(fp) if want_table else None, C.byref(mh) if (want_table or want_max_height) else None))
        freq = np.ctypeslib.as_array(fp, shape=(mh.value + 1,)).copy() if want_table else None
        return freq, mh.value

    def count_stacks_by_group(self, want_table=True, want_count=True):
        g = np.empty((self.W, BINS, 2, 2), np.uint64) if want_table else None
        n = C.c_uint64(0)
        _check(lib().ppp_scan(self._h, g.ctypes.data if want_table else None, C.byref(n) if (want_table or want_count) else None))
        if want_table or want_count:
            self.n_pairs = n.value
        return g

    def pass_info(self) -> dict:
        """Tallest stack, signatures, overlap-10 signatures and stacks of the last pass (waits for it)."""
        mh, a, b, s = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().ppp_pass_info(self._h, C.byref(mh), C.byref(a), C.byref(b), C.byref(s)))
        self.n_pairs = a.value
        return dict(max_height=mh.value, n_pairs=a.value, n_pingpong=b.value, n_stacks=s.value)

    def calculate_fdrs(self, min_stack_height=0, want_tables=True, want_loci=True, copy_loci=True) -> ScoreResult:
        """copy_loci=False returns a view of the library's pinned result buffer (valid until the next call)."""
        nb = C.c_uint32(0)
        bin_map = np.empty(BINS, np.uint16) if want_tables else None
        cells = np.empty(self.W * BINS * 4, np.float32) if want_tables else None
        fdr = np.empty(self.W * BINS * 4, np.float32) if want_tables else None
        lp = C.c_void_p()
        nl = C.c_size_t(0)
        _check(lib().ppp_score(self._h, min_stack_height, C.byref(nb), bin_map.ctypes.data if want_tables else None,
                               cells.ctypes.data if want_tables else None, fdr.ctypes.data if want_tables else None,
                               C.byref(lp) if want_loci else None, C.byref(nl)))
        Cn = nb.value
        loci = None
        if want_loci:
            if nl.value:
                buf = (C.c_char * (LOCUS_DTYPE.itemsize * nl.value)).from_address(lp.value)
                loci = np.frombuffer(buf, dtype=LOCUS_DTYPE)
                if copy_loci:
                    loci = loci.copy()
            else:
                loci = np.empty(0, LOCUS_DTYPE)
        if want_tables:
            cells = cells[: self.W * Cn * 4].reshape(self.W, Cn, 2, 2)
            fdr = fdr[: self.W * Cn * 4].reshape(self.W, Cn, 2, 2)
        res = ScoreResult(Cn, bin_map, cells, fdr, loci)
        res.n_loci = nl.value
        return res

    def signatures(self) -> np.ndarray:
        n = C.c_size_t(0)
        _check(lib().ppp_pairs_download(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, PAIR_DTYPE)
        if n.value:
            _check(lib().ppp_pairs_download(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def find_suppressed_transposons(self, intervals):
        iv = np.ascontiguousarray(intervals, dtype=INTERVAL_DTYPE)
        out = np.zeros(len(iv), TP_RESULT_DTYPE)
        order = np.zeros(len(iv), np.uint32)
        _check(lib().ppp_transposons(self._h, iv.ctypes.data, len(iv), out.ctypes.data, order.ctypes.data))
        return out, order

    # -- misc ------------------------------------------------------------------------------------------
    def timings(self) -> dict:
        t = Timings()
        _check(lib().ppp_timings_get(self._h, C.byref(t)))
        return t.as_dict()

    def timings_reset(self):
        _check(lib().ppp_timings_reset(self._h))

    def layout(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(lib().ppp_layout(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_allreduce(self, fn):
        """fn(device_ptr: int, count: int, dtype: 0=u32/1=u64, op: 0=sum/1=max, stream: int) -> None"""
        def tramp(user, buf, count, dtype, op, stream):
            try:
                fn(buf, count, dtype, op, stream)
                return 0
            except Exception as e:  # pragma: no cover
                print("all-reduce hook failed:", e, flush=True)
                return 1
        self._hook = ALLREDUCE_FN(tramp)
        _check(lib().ppp_set_allreduce(self._h, self._hook, None))

    def set_rank(self, rank: int, world: int):
        """With a custom hook: which of `world` shards this context is (enables the compact exchange of tall stacks)."""
        _check(lib().ppp_set_rank(self._h, rank, world))

    def attach_nccl(self, unique_id: bytes, rank: int, world: int):
        """Built-in collective: creates this rank's NCCL communicator from the id of nccl_unique_id() (collective call)."""
        if len(unique_id) != NCCL_UNIQUE_ID_BYTES:
            raise ValueError("NCCL unique id must be %d bytes" % NCCL_UNIQUE_ID_BYTES)
        buf = C.create_string_buffer(unique_id, NCCL_UNIQUE_ID_BYTES)
        _check(lib().ppp_nccl_attach(self._h, buf, rank, world))


NCCL_UNIQUE_ID_BYTES = 128


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    _check(lib().ppp_nccl_unique_id(buf))
    return buf.raw


def run_pipeline(contig_lengths, reads_per_contig, mode="weighted", min_stack_height=0, intervals=None, **kw):
    """The hot path end to end on one GPU, mirroring pingpongpro.cpp:1477-1606. Returns a dict of results."""
    with Context(contig_lengths, mode=mode, **kw) as ctx:
        for c, r in enumerate(reads_per_contig):
            if len(r):
                ctx.count_reads(c, r)
        ctx.finish()
        freq, max_h = ctx.map_heights_to_scores()
        grouped = ctx.count_stacks_by_group()
        res = ctx.calculate_fdrs(min_stack_height)
        out = dict(freq=freq, max_height=max_h, grouped=grouped, score=res, signatures=ctx.signatures(), timings=ctx.timings())
        if intervals is not None:
            out["transposons"], out["transposon_order"] = ctx.find_suppressed_transposons(intervals)
        return out
