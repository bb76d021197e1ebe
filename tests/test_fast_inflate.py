"""The project's own raw-DEFLATE decoder (host/fast_inflate.cpp, tried before zlib on every BGZF block) against zlib:
identical bytes on everything zlib can produce -- all levels and strategies, stored / fixed / dynamic blocks, long
matches and long codes -- and a clean refusal (never a crash, never wrong bytes passed off as right) on truncated or
corrupted streams."""
import ctypes as C
import zlib

import numpy as np
import pytest

from pingpongpro_b200 import host


def decode(stream: bytes, out_len: int):
    L = host.host_lib()
    L.ppp_fast_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ppp_fast_inflate.restype = C.c_int
    out = (C.c_ubyte * (out_len + 64))()  # slack so that an overrun would be visible rather than fatal
    guard = bytes(range(64))
    C.memmove(C.addressof(out) + out_len, guard, 64)
    ok = L.ppp_fast_inflate(stream, len(stream), out, out_len)
    assert bytes(out[out_len:out_len + 64]) == guard, "wrote past the end of the output"
    return bool(ok), bytes(out[:out_len])


def raw_deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8, flush_every=0) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    if not flush_every:
        return c.compress(data) + c.flush()
    out = b""
    for o in range(0, len(data), flush_every):  # several blocks, byte-aligned by empty stored blocks
        out += c.compress(data[o:o + flush_every]) + c.flush(zlib.Z_FULL_FLUSH)
    return out + c.flush()


def payloads():
    rng = np.random.default_rng(7)
    yield "empty", b""
    yield "one byte", b"x"
    yield "zeros", bytes(65280)
    yield "run of 3", b"abc" * 20000
    yield "random", rng.integers(0, 256, 60000, dtype=np.uint8).tobytes()
    yield "dna", bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 65000).tobytes())
    skew = rng.zipf(1.3, 65000) % 256
    yield "skewed (long codes)", skew.astype(np.uint8).tobytes()
    text = b"".join(b"read%07d\t%d\tchr%d\t%d\t255\t%dM\t*\t0\t0\tACGTACGTAAACGTACGTACGTACG\t*\tNH:i:%d\n" % (i, 16 * (i % 2), i % 7, i * 13, 21 + i % 12, 1 + i % 5)
                    for i in range(700))
    yield "sam text", text
    yield "far matches", (rng.integers(0, 256, 30000, dtype=np.uint8).tobytes() + b"q") * 2
    yield "distance 1..9", b"".join(bytes([65 + k]) * (1 + k) + b"xy" * (40 + k) + b"abcdefghi"[:k + 1] * 30 for k in range(9))


@pytest.mark.parametrize("name,data", list(payloads()), ids=[n for n, _ in payloads()])
def test_equals_zlib_for_all_levels_and_strategies(name, data):
    for level in (0, 1, 4, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
            for flush_every in (0, 7001):
                stream = raw_deflate(data, level, strategy, flush_every=flush_every)
                assert zlib.decompress(stream, -15) == data
                ok, out = decode(stream, len(data))
                assert ok and out == data, (level, strategy, flush_every)
                if len(data):
                    assert not decode(stream, len(data) - 1)[0]   # the size is part of the contract
                assert not decode(stream, len(data) + 1)[0]


def test_small_window_and_memory_levels():
    data = (b"ACGTTGCA" * 9000)[:65280]
    for mem in (1, 4, 9):
        stream = raw_deflate(data, 6, mem=mem)
        ok, out = decode(stream, len(data))
        assert ok and out == data


def test_truncated_and_corrupted_streams_are_refused_or_right():
    rng = np.random.default_rng(11)
    data = b"".join(b"r%05d\t0\tc1\t%d\t255\t25M\t*\t0\t0\tACGTACGTAAACGTACGTACGTACG\t*\n" % (i, i * 7) for i in range(900))
    stream = raw_deflate(data, 6)
    for cut in list(range(0, 40)) + list(range(len(stream) - 40, len(stream))):
        ok, out = decode(stream[:cut], len(data))
        assert not ok
    refused = 0
    for trial in range(3000):
        s = bytearray(stream)
        for _ in range(1 + trial % 3):
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        ok, out = decode(bytes(s), len(data))
        if not ok:
            refused += 1
            continue
        try:  # accepted: then zlib must accept it too, with the same bytes
            assert zlib.decompress(bytes(s), -15) == out
        except zlib.error:
            pytest.fail("accepted a stream that zlib rejects")
    assert refused > 1000  # (a flipped bit often still decodes to the right size; then the bytes must be zlib's)
    for trial in range(300):  # garbage
        junk = rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        ok, out = decode(junk, 1000)
        if ok:
            assert zlib.decompress(junk, -15) == out


def test_bgzf_blocks_of_a_synthetic_bam(tmp_path):
    import os
    import subprocess

    from pingpongpro_b200 import _build

    f = str(tmp_path / "x.bam")
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "50000", "--seed", "3", "--out", f])
    raw, off, n = open(f, "rb").read(), 0, 0
    while off + 18 <= len(raw):
        size = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        body, isize = raw[off + 18:off + size - 8], int.from_bytes(raw[off + size - 4:off + size], "little")
        ok, out = decode(body, isize)
        assert ok and out == zlib.decompress(body, -15)
        off += size
        n += 1
    assert n > 20
