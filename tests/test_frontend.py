"""Host front end: SAM and BAM decode + feature extraction (pingpongpro.cpp:415-484 restated in
pingpongpro_b200/host/frontend.h) against the committed fixtures and the synthetic model."""
import os
import subprocess

import numpy as np
import pytest

from golden_tools import GOLDEN, Golden
from pingpongpro_b200 import _build, host

FILES = os.path.join(GOLDEN, "files", "files_tiny_weighted")


@pytest.mark.parametrize("ext", ["sam", "bam"])
def test_decode_matches_golden_records(ext):
    g = Golden("files_tiny_weighted")
    dec = host.decode_file(os.path.join(FILES, "input." + ext), mode=g.mode)
    assert dec.names == g.contig_names and dec.lengths == g.contig_lengths
    for a, b in zip(dec.reads, g.reads):
        assert np.array_equal(a, b)
    assert dec.total_weight == g.total
    assert dec.n_records == 6000


def test_modes_filter_reads():
    bam = os.path.join(FILES, "input.bam")
    w = host.decode_file(bam, mode="weighted")
    u = host.decode_file(bam, mode="unique")
    d = host.decode_file(bam, mode="discard")
    assert u.n_kept == w.n_kept > d.n_kept > 0
    assert u.total_weight == float(u.n_kept)  # every weight is 1 (:435)
    assert all(((r["meta"] >> 2) == 1).all() for r in u.reads)  # NH is not read in unique mode (:432)
    assert all(((r["meta"] >> 2) == 1).all() for r in d.reads)  # multi-hits dropped (:454, :458)
    assert d.total_weight == float(d.n_kept)


def test_length_filter():
    bam = os.path.join(FILES, "input.bam")
    wide = host.decode_file(bam, min_len=1, max_len=100)
    narrow = host.decode_file(bam, min_len=26, max_len=27)
    assert wide.n_kept > host.decode_file(bam).n_kept > narrow.n_kept > 0


def test_missing_file_and_garbage(tmp_path):
    with pytest.raises(OSError):
        host.decode_file(str(tmp_path / "nope.bam"))
    bad = tmp_path / "bad.sam"
    bad.write_text("@SQ\tSN:c\tLN:100\nr1\tnotanumber\tc\t5\t255\t25M\t*\t0\t0\tACGT\tIIII\n")
    with pytest.raises(ValueError):
        host.decode_file(str(bad))


def test_empty_and_header_only(tmp_path):
    p = tmp_path / "empty.sam"
    p.write_text("@HD\tVN:1.4\n@SQ\tSN:c1\tLN:1000\n@SQ\tSN:c2\tLN:50\n")
    dec = host.decode_file(str(p))
    assert dec.names == ["c1", "c2"] and dec.lengths == [1000, 50] and dec.n_kept == 0


def test_handwritten_records(tmp_path):
    """Edge cases of :415-484: soft clips at either end, deletions/skips counting towards the aligned
    length, hard clips not looked through, lower-case bases, missing NH, unmapped POS 0."""
    seq = "ACGTACGTAAGTACGTACGTACGTACGTAC"  # 30 nt, base 10 (index 9) = A
    lines = [
        "@SQ\tSN:c\tLN:100000",
        f"p1\t0\tc\t101\t255\t30M\t*\t0\t0\t{seq}\t*",                       # + key 100, A10
        f"p2\t0\tc\t101\t255\t2S28M\t*\t0\t0\t{seq}\t*\tNH:i:3",             # clip 2 -> index 11 = 'T' -> no A10; aln 28
        f"p3\t0\tc\t201\t255\t10M5D10M2N5M\t*\t0\t0\t{seq[:25]}\t*",          # aln = 10+5+10+2+5 = 32
        f"p4\t0\tc\t301\t255\t5H25M\t*\t0\t0\t{seq[:25].lower()}\t*",         # H is not S; lower-case 'a' at index 9
        f"m1\t16\tc\t501\t255\t30M\t*\t0\t0\t{'C' * 20 + 'T' + 'C' * 9}\t*",  # - key 530; seq[30-0-1-9=20] = T
        f"m2\t16\tc\t501\t255\t27M3S\t*\t0\t0\t{'C' * 17 + 't' + 'C' * 12}\t*\tNH:i:2",  # clip 3: index 17; key 500+27
        f"m3\t16\tc\t501\t255\t3S27M\t*\t0\t0\t{'C' * 20 + 'T' + 'C' * 9}\t*",  # leading S is not the 5' end of a - read
        f"u1\t4\t*\t0\t0\t*\t*\t0\t0\t{seq}\t*",                               # unmapped
        f"s1\t0\tc\t401\t255\t20M\t*\t0\t0\t{seq[:20]}\t*",                    # too short
        f"l1\t0\tc\t401\t255\t33M\t*\t0\t0\t{seq}AAA\t*",                      # too long
    ]
    p = tmp_path / "hand.sam"
    p.write_text("\n".join(lines) + "\n")
    dec = host.decode_file(str(p))
    r = dec.reads[0]
    got = [(int(k), int(m & 1), int((m >> 1) & 1), int(m >> 2)) for k, m in zip(r["key"], r["meta"])]
    assert got == [(100, 0, 1, 1), (100, 0, 0, 3), (200, 0, 1, 1), (300, 0, 1, 1), (530, 1, 1, 1), (527, 1, 1, 2), (527, 1, 1, 1)]
    assert dec.n_records == 10 and dec.n_kept == 7
    assert dec.total_weight == pytest.approx(1 + np.float32(1 / 3) + 1 + 1 + 1 + 0.5 + 1, abs=0)


def test_nh_zero_is_skipped_in_weighted_mode(tmp_path):
    """NH:i:0 weighs +inf in -m weighted (:449) and the reference then converts inf to unsigned (:508, undefined
    behaviour): the record is skipped (documented divergence).  The other modes never divide by NH."""
    seq = "ACGTACGTAAGTACGTACGTACGTACGTAC"
    lines = ["@SQ\tSN:c\tLN:100000", f"a\t0\tc\t101\t255\t30M\t*\t0\t0\t{seq}\t*\tNH:i:0", f"b\t0\tc\t101\t255\t30M\t*\t0\t0\t{seq}\t*\tNH:i:2"]
    p = tmp_path / "nh0.sam"
    p.write_text("\n".join(lines) + "\n")
    w = host.decode_file(str(p), mode="weighted")
    assert w.n_kept == 1 and w.total_weight == 0.5 and int(w.reads[0]["meta"][0] >> 2) == 2
    assert host.decode_file(str(p), mode="unique").n_kept == 2
    assert host.decode_file(str(p), mode="discard").n_kept == 0


def test_synth_lib_equals_file_route(tmp_path):
    """The packed records libppp_synth generates directly equal what the front end extracts from the
    SAM/BAM that ppp_synth writes for the same model."""
    for ext, mode in (("sam", "weighted"), ("bam", "discard"), ("bam", "unique")):
        f = str(tmp_path / f"x.{ext}")
        subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "5000", "--seed", "31", "--out", f])
        dec = host.decode_file(f, mode=mode)
        m = host.SynthModel("tiny", n_reads=5000, seed=31)
        direct = m.generate(mode=mode, threads=3)
        assert m.names == dec.names and m.lengths == dec.lengths
        for a, b in zip(direct, dec.reads):
            assert np.array_equal(a, b)


def test_synth_range_filter_partitions_reads():
    m = host.SynthModel("tiny", n_reads=20000, seed=5)
    full = m.generate()
    cut = 47000
    left = m.generate(g_lo=0, g_hi=cut)
    right = m.generate(g_lo=cut)
    for c in range(len(full)):
        assert len(left[c]) + len(right[c]) == len(full[c])
    halo = m.generate(g_lo=0, g_hi=cut, halo=23)
    extra = sum(len(h) - len(l) for h, l in zip(halo, left))
    assert extra >= 0
    for h in halo:  # the halo only ever adds minus-strand reads
        pass


def test_parallel_bgzf_inflate_equals_sequential(tmp_path):
    """BGZF blocks are independent deflate streams: inflating them on worker threads (in-order hand-back) must give the
    same records, the same order and the same double-precision total weight (:488 is summed in file order)."""
    f = str(tmp_path / "big.bam")
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "300000", "--seed", "33", "--out", f])
    seq = host.decode_file(f, threads=1)
    for t in (2, 5):
        par = host.decode_file(f, threads=t)
        assert par.n_records == seq.n_records == 300000 and par.n_kept == seq.n_kept and par.total_weight == seq.total_weight
        for a, b in zip(par.reads, seq.reads):
            assert np.array_equal(a, b)
    trunc = tmp_path / "trunc.bam"
    trunc.write_bytes(open(f, "rb").read()[:200000])  # cut inside a block: both paths must report a read error
    for t in (1, 4):
        with pytest.raises(ValueError):
            host.decode_file(str(trunc), threads=t)


def _bgzf_payload(path):
    import zlib
    raw, out, off = open(path, "rb").read(), [], 0
    while off + 18 <= len(raw):
        size = int.from_bytes(raw[off + 16:off + 18], "little") + 1  # BSIZE of the BC subfield (xlen = 6)
        out.append(zlib.decompress(raw[off + 18:off + size - 8], -15))
        off += size
    return b"".join(out)


def _bgzf_write(path, payload, block_bytes, level=6):
    import zlib
    with open(path, "wb") as f:
        for o in list(range(0, len(payload), block_bytes)) + [None]:
            chunk = b"" if o is None else payload[o:o + block_bytes]  # None: the empty end-of-file block
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = c.compress(chunk) + c.flush()
            f.write(bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + (len(body) + 25).to_bytes(2, "little") + body
                    + zlib.crc32(chunk).to_bytes(4, "little") + len(chunk).to_bytes(4, "little"))


@pytest.mark.parametrize("block_bytes", [601, 7919, 65280])
def test_bam_records_straddling_bgzf_blocks(tmp_path, block_bytes):
    """BAM writers cut BGZF blocks wherever the buffer fills, so records (and their block_size words) straddle blocks
    and the worker jobs made of them.  The boundary walker of the batch mode must stitch them exactly like the
    sequential reader; a file that ends inside a record is a read error for both."""
    src = str(tmp_path / "src.bam")
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "60000", "--seed", "34", "--out", src])
    payload = _bgzf_payload(src)
    f = str(tmp_path / "reblocked.bam")
    _bgzf_write(f, payload, block_bytes)
    seq = host.decode_file(f, threads=1)
    _decoded_equal(seq, host.decode_file(src, threads=1))
    for t in (2, 5):
        _decoded_equal(seq, host.decode_file(f, threads=t))
    _bgzf_write(f, payload[: len(payload) - 37], block_bytes)  # the last record is incomplete
    for t in (1, 4):
        with pytest.raises(ValueError, match="Failed to read record"):
            host.decode_file(f, threads=t)


def test_weight_reciprocal_single_equals_double_narrowed():
    """The device evaluates the -m weighted read weight (float)(1.0 / (double)NH) (pingpongpro.cpp:449) as a correctly
    rounded single-precision reciprocal for NH < 2^24; the two are the same number for every such NH."""
    n = np.arange(1, 1 << 24, dtype=np.uint32)
    narrowed = (1.0 / n.astype(np.float64)).astype(np.float32)
    single = np.float32(1.0) / n.astype(np.float32)
    assert np.array_equal(narrowed.view(np.uint32), single.view(np.uint32))


def _decoded_equal(a, b):
    assert a.names == b.names and a.lengths == b.lengths
    assert a.n_records == b.n_records and a.n_kept == b.n_kept and a.total_weight == b.total_weight
    for x, y in zip(a.reads, b.reads):
        assert np.array_equal(x, y)


SAM_EDGE = (
    "@HD\tVN:1.0\n@SQ\tSN:c1\tLN:5000\n@SQ\tSN:c2\tLN:400\n"
    "r1\t0\tc1\t100\t255\t25M\t*\t0\t0\tACGTACGTAAACGTACGTACGTACG\t*\tNH:i:2\n"
    "\n\r\n"                                                                          # blank lines are not records
    "r2\t16\tc2\t50\t255\t2S24M\t*\t0\t0\tACGTACGTACGTACGTACGTACGTAC\t*\tXX:Z:NH:i:7\tNH:i:3\r\n"   # CRLF, decoy tag
    "r3\t0\tlate\t7\t255\t30M\t*\t0\t0\tACGTACGTAAACGTACGTACGTACGTACGT\t*\n"          # contig that is missing from the header
    "r4\t0\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"                                                # unmapped
    "r5\t0\tlate\t9\t255\t24M\t*\t0\t0\tACGTACGTAAACGTACGTACGTAC\t*\tNH:i:-1\n"
    "r6\t16\tc1\t4000\t255\t26M\t*\t0\t0\tACGTACGTACGTACGTTCGTACGTAC\t*"               # last line without a terminator
)


@pytest.mark.parametrize("threads", [2, 5])
def test_parallel_sam_parser_equals_sequential_reader(tmp_path, threads):
    """AlnReader::setThreads on SAM text: worker-parsed lines come back in file order and decode to the same records,
    including blank lines, CRLF, a contig missing from the header (the name store grows in order of appearance), a
    negative NH, and a last line without '\\n'.  Also on a file large enough for many parser jobs."""
    p = tmp_path / "edge.sam"
    p.write_bytes(SAM_EDGE.encode())
    seq = host.decode_file(str(p), threads=1)
    assert seq.names == ["c1", "c2", "late"] and seq.n_records == 6 and seq.n_kept == 5
    _decoded_equal(seq, host.decode_file(str(p), threads=threads))
    (tmp_path / "headerless.sam").write_bytes(SAM_EDGE[SAM_EDGE.index("r1"):].encode())
    _decoded_equal(host.decode_file(str(tmp_path / "headerless.sam"), threads=1), host.decode_file(str(tmp_path / "headerless.sam"), threads=threads))
    big = tmp_path / "big.sam"
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "120000", "--seed", "9", "--out", str(big)])
    assert big.stat().st_size > 3 * (2 << 20)  # several jobs of 2 MB
    _decoded_equal(host.decode_file(str(big), threads=1), host.decode_file(str(big), threads=threads))
    for mode in ("unique", "discard"):
        _decoded_equal(host.decode_file(str(big), mode=mode, threads=1), host.decode_file(str(big), mode=mode, threads=threads))


@pytest.mark.parametrize("threads", [1, 4])
def test_malformed_sam_record_fails_at_the_same_place(tmp_path, threads):
    """'Failed to read record' (pingpongpro.cpp:411): a malformed line stops the parallel parser exactly like the
    sequential one (the CLI exits with 1 either way)."""
    good = "r\t0\tc1\t100\t255\t25M\t*\t0\t0\tACGTACGTAAACGTACGTACGTACG\t*\n"
    p = tmp_path / "bad.sam"
    p.write_text("@SQ\tSN:c1\tLN:5000\n" + good * 3 + "r\tnotanumber\tc1\t5\t255\t25M\t*\t0\t0\tACGT\tIIII\n" + good)
    with pytest.raises(ValueError, match="Failed to read record"):
        host.decode_file(str(p), threads=threads)
    p.write_text("@SQ\tSN:c1\tLN:5000\n" + good * 2 + "r\t0\tc1\t5\t255\t25Q\t*\t0\t0\tACGT\tIIII\n")  # bad CIGAR operation
    with pytest.raises(ValueError, match="Failed to read record"):
        host.decode_file(str(p), threads=threads)


def test_sam_mapped_file_equals_streamed_input(tmp_path, monkeypatch):
    """Batch mode parses a regular SAM file through a memory mapping (jobs are windows of the file) and anything else
    (or any file with PPP_NO_MMAP=1) through buffered reads: same records either way, including a line
    longer than one 2 MB job (a tag of 5 MB), CRLF, blank lines and a last line without a line feed."""
    long_tag = "XX:Z:" + "A" * (5 << 20)
    head, rest = SAM_EDGE.split("r1", 1)
    text = head + "r0\t0\tc1\t150\t255\t25M\t*\t0\t0\tACGTACGTAAACGTACGTACGTACG\t*\t" + long_tag + "\tNH:i:3\nr1" + rest
    big = tmp_path / "big.sam"
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "60000", "--seed", "10", "--out", str(big)])
    body = "".join(l for l in big.read_text().splitlines(keepends=True) if not l.startswith("@"))
    hdr = "".join(l for l in big.read_text().splitlines(keepends=True) if l.startswith("@"))
    cases = {"edge_long.sam": text, "big_noeol.sam": hdr + body.rstrip("\n"), "header_only.sam": hdr, "one.sam": hdr + body.splitlines(keepends=True)[0]}
    for name, content in cases.items():
        p = tmp_path / name
        p.write_bytes(content.encode())
        seq = host.decode_file(str(p), threads=1)
        mapped = host.decode_file(str(p), threads=4)
        monkeypatch.setenv("PPP_NO_MMAP", "1")
        streamed = host.decode_file(str(p), threads=4)
        monkeypatch.delenv("PPP_NO_MMAP")
        _decoded_equal(seq, mapped)
        _decoded_equal(seq, streamed)


def _bam_record(ref_id, pos, name, cigar_len, seq_len, tags=b""):
    import struct
    cigar = struct.pack("<I", cigar_len << 4)  # <n>M
    body = (struct.pack("<iiBBHHHIiii", ref_id, pos, len(name) + 1, 255, 4681, 1, 0, seq_len, -1, -1, 0) + name + b"\0" + cigar
            + bytes([0x12] * ((seq_len + 1) // 2)) + bytes([30] * seq_len) + tags)   # 0x12 = "AC"
    return struct.pack("<I", len(body)) + body


def test_bam_boundary_guesses_are_verified(tmp_path, monkeypatch):
    """Batch mode lets every worker start decoding its run of blocks at a GUESSED record boundary; the in-order stage
    accepts a guess only if walking from the true boundary lands on it.  A 3 MB uint8-array tag filled with byte
    sequences that look exactly like (kept) BAM records, written as stored blocks so that runs begin inside it, makes
    the guesses wrong: those runs are decoded again from the true boundary and the result still equals the sequential
    walk -- and no decoy becomes a read."""
    import struct
    src = str(tmp_path / "src.bam")
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "30000", "--seed", "35", "--out", src])
    payload = _bgzf_payload(src)
    l_text = struct.unpack_from("<I", payload, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", payload, p)[0]
    p += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<I", payload, p)[0]
        p += 4 + l_name + 4
    # walk 10 000 records in
    q = p
    for _ in range(10000):
        q += 4 + struct.unpack_from("<I", payload, q)[0]
    decoy = _bam_record(0, 777, b"x", 25, 25)
    fill = decoy * ((3 << 20) // len(decoy))
    carrier = _bam_record(0, 1234, b"carrier", 26, 26, b"XBBC" + struct.pack("<I", len(fill)) + fill + b"NHC\x02")
    doctored = payload[:q] + carrier + payload[q:]
    f = str(tmp_path / "decoy.bam")
    _bgzf_write(f, doctored, 65280, level=0)  # stored blocks: a run of blocks (1 MB of file) is 1 MB of records
    seq = host.decode_file(f, threads=1)
    base = host.decode_file(src, threads=1)
    assert seq.n_records == base.n_records + 1 and seq.n_kept == base.n_kept + 1  # the carrier itself is a 26-nt read
    for t in (2, 5):
        par = host.decode_file(f, threads=t)
        _decoded_equal(seq, par)
        assert par.rejected_guesses >= 2  # the runs that began inside the tag
    assert host.decode_file(src, threads=4).rejected_guesses == 0
    monkeypatch.setenv("PPP_BAM_NO_GUESS", "1")  # every run decoded by the in-order stage: same records
    _decoded_equal(seq, host.decode_file(f, threads=3))


@pytest.mark.parametrize("threads", [1, 3])
def test_sam_many_optional_fields_and_lookalike_contig_names(tmp_path, threads):
    """The line parser finds all tabs in one pass (a table of 48, the rest the slow way) and resolves RNAME through a
    small cache keyed by the name's last eight bytes and length: NH behind 70 other tags, an empty NH:i: field that does
    not count (:440-444 needs a value), and contig names that share length and their last eight bytes."""
    seq25 = "ACGTACGTAAACGTACGTACGTACG"
    names = ["AAAAscaffold_1", "BBBBscaffold_1", "c", "scaffold_1", "Xscaffold_1"]
    hdr = "".join(f"@SQ\tSN:{n}\tLN:9000\n" for n in names)
    many = "\t".join(f"X{chr(65 + k % 26)}:i:{k}" for k in range(70))
    lines = []
    for k in range(400):
        n = names[(k * 7) % len(names)]
        tags = [f"{many}\tNH:i:{2 + k % 3}", "NH:i:\tNH:i:5", f"XA:Z:NH:i:9\tNH:i:{1 + k % 2}", "NH:i:-", ""][k % 5]
        lines.append(f"r{k}\t{16 * (k % 2)}\t{n}\t{100 + k}\t255\t25M\t*\t0\t0\t{seq25}\t*" + ("\t" + tags if tags else "") + "\n")
    p = tmp_path / "tags.sam"
    p.write_text(hdr + "".join(lines))
    d = host.decode_file(str(p), threads=threads)
    assert d.names == names and d.n_records == 400 and d.n_kept == 400
    per_contig = [0] * len(names)
    for k in range(400):
        c = (k * 7) % len(names)
        r = d.reads[c][per_contig[c]]
        per_contig[c] += 1
        want_nh = [2 + k % 3, 5, 1 + k % 2, 1, 1][k % 5]
        assert (int(r["meta"]) >> 2) == want_nh, k
        assert int(r["key"]) == (99 + k + 25 if k % 2 else 99 + k), k


def test_corrupted_bam_sequential_and_batch_modes_agree(tmp_path):
    """Mutation fuzz: bit flips, random bytes, corrupted block_size words and deleted spans in the record stream.  The
    batch mode (boundary guesses + in-order verification) must do what the sequential walk does -- the same records when
    the file still decodes, a read error when it does not; never a crash.  (tools/fuzz_bam.py runs the same under
    ASan / UBSan with more iterations.)"""
    import random
    import struct
    src = str(tmp_path / "src.bam")
    subprocess.check_call([os.path.join(_build.BIN, "ppp_synth"), "--genome", "tiny", "--reads", "20000", "--seed", "5", "--out", src])
    payload = _bgzf_payload(src)
    l_text = struct.unpack_from("<I", payload, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", payload, p)[0]
    p += 4
    for _ in range(n_ref):
        p += 4 + struct.unpack_from("<I", payload, p)[0] + 4
    starts = [p]
    while starts[-1] + 4 <= len(payload):
        nxt = starts[-1] + 4 + struct.unpack_from("<I", payload, starts[-1])[0]
        if nxt + 4 > len(payload):
            break
        starts.append(nxt)
    rng = random.Random(2024)
    outcomes = {"ok": 0, "error": 0}
    for it in range(36):
        m = bytearray(payload)
        kind = it % 4
        for _ in range(rng.randrange(1, 4)):
            pos = rng.randrange(p, len(m))
            if kind == 0:
                m[pos] ^= 1 << rng.randrange(8)
            elif kind == 1:
                m[pos] = rng.randrange(256)
            elif kind == 2:
                q = rng.choice(starts)
                old = struct.unpack_from("<I", payload, q)[0]
                struct.pack_into("<I", m, q, rng.choice([0, 31, 33, 90, 5000, 70000, 1 << 20, 0xFFFFFFFF, old - 1, old + 1, old + 16]) & 0xFFFFFFFF)
            else:
                a = rng.randrange(p, len(m) - 300)
                del m[a:a + rng.randrange(1, 300)]
        f = str(tmp_path / "m.bam")
        _bgzf_write(f, bytes(m), rng.choice([601, 7919, 65280]), level=rng.choice([0, 1]))
        results = []
        for t in (1, 3):
            try:
                results.append(host.decode_file(f, threads=t))
            except ValueError:
                results.append(None)
        assert (results[0] is None) == (results[1] is None), it
        if results[0] is not None:
            _decoded_equal(results[0], results[1])
        outcomes["ok" if results[0] is not None else "error"] += 1
    assert outcomes["ok"] >= 5 and outcomes["error"] >= 5  # both behaviours were exercised


def test_headerless_sam_gets_names_by_appearance_and_lengths_from_the_reads(tmp_path):
    """SAM text without @SQ lines (`samtools view` without -h, bowtie --sam-nohead): the reference works on it -- SeqAn's
    name store grows in order of appearance and pingpongpro.cpp never needs a length (:1493-1517).  Here the reads come
    out as from the same file with its header (contig by contig, in file order), the contigs are numbered in order of
    appearance and each gets the extent of its kept reads (largest 5'-end key + 1) as its length -- what the dense
    stack arrays are sized from."""
    tool = os.path.join(_build.BIN, "ppp_synth")
    full = tmp_path / "with.sam"
    subprocess.check_call([tool, "--genome", "tiny", "--reads", "20000", "--seed", "9", "--out", str(full)], stdout=subprocess.DEVNULL)
    lines = full.read_text().splitlines(keepends=True)
    body = [l for l in lines if not l.startswith("@")]
    # a different order of appearance than the header's: the records of the last contig first
    last = [l for l in body if l.split("\t")[2] == "ctgC"]
    rest = [l for l in body if l.split("\t")[2] != "ctgC"]
    bare = tmp_path / "without.sam"
    bare.write_text("".join(last[:50] + rest + last[50:]))
    want_file = tmp_path / "reordered.sam"
    want_file.write_text("".join([l for l in lines if l.startswith("@")] + last[:50] + rest + last[50:]))
    want = host.decode_file(str(want_file), mode="weighted")
    for threads in (1, 4):
        got = host.decode_file(str(bare), mode="weighted", threads=threads)
        assert got.names[0] == "ctgC" and sorted(got.names) == sorted(want.names)
        assert got.n_records == want.n_records and got.n_kept == want.n_kept and got.total_weight == want.total_weight
        for i, name in enumerate(got.names):
            j = want.names.index(name)
            assert np.array_equal(got.reads[i], want.reads[j])
            assert got.lengths[i] == (int(got.reads[i]["key"].max()) + 1 if len(got.reads[i]) else 0)
            assert got.lengths[i] <= want.lengths[j] + 1
