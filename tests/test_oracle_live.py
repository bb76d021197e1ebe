"""Differential test of the CPU restatement (oracle/ppp_oracle.c) against the UNMODIFIED reference run live
(oracle/_ref/ppp_harness, built from /root/reference by oracle/Makefile): seeded inputs beyond the committed golden
fixtures -- other seeds, densities, contig layouts, the three multi-hit modes and the widened background ranges -- must
agree bit for bit in every intermediate.  Skipped where the harness does not exist (it is built wherever
/root/reference is mounted)."""
import os
import subprocess

import numpy as np
import pytest

from oracle_tools import IV_DTYPE, Oracle, have_ref, run_harness
from pingpongpro_b200 import _build, host

SYNTH = os.path.join(_build.BIN, "ppp_synth")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


CASES = [
    # contigs, reads, seed, mode, max overlap, extra synth flags
    ("a:50000,b:20000", 3000, 101, "weighted", 23, []),
    ("a:50000,b:20000", 3000, 102, "unique", 23, ["--sorted"]),
    ("a:50000,b:20000", 3000, 103, "discard", 23, []),
    ("x:9000", 25000, 104, "weighted", 23, []),                       # very dense: several collapsed bins
    ("x:9000,y:7000,z:300", 30000, 105, "unique", 30, []),
    ("p:200000,q:100000,r:100000,s:5000,t:64", 20000, 106, "weighted", 26, ["--nh1"]),
    ("p:200000,q:100000", 50000, 107, "weighted", 30, []),
    ("only:1000000", 8000, 108, "discard", 26, ["--sorted"]),
]


@pytest.mark.skipif(not have_ref(23) or not have_ref(26) or not have_ref(30), reason="oracle/_ref was not built here")
@pytest.mark.parametrize("contigs,n_reads,seed,mode,max_ov,flags", CASES, ids=[f"{c[3]}-seed{c[2]}-max{c[4]}" for c in CASES])
def test_port_equals_reference_run_live(tmp_path, contigs, n_reads, seed, mode, max_ov, flags):
    sam, annot = str(tmp_path / "in.sam"), str(tmp_path / "annot.tsv")
    subprocess.check_call([SYNTH, "--contigs", contigs, "--reads", str(n_reads), "--seed", str(seed), "--out", sam, "--annot", annot, *flags])
    ref = run_harness(str(tmp_path / "dump.bin"), [sam], mode=mode, annots=[annot], max_overlap=max_ov)
    dec = host.decode_file(sam, mode=mode)
    o = Oracle(dec.reads, mode=mode, min_ov=3, max_ov=max_ov, pp_ov=10)
    W = max_ov - 3 + 1
    st = o.stacks()
    assert len(st["key"]) == len(ref["stacks"])
    for k in ("strand", "contig", "key", "a10"):
        assert np.array_equal(st[k], ref["stacks"][k]), k
    assert np.array_equal(bits(st["reads"]), bits(ref["stacks"]["reads"]))
    assert o.total() == ref["total"] == dec.total_weight
    freq = np.zeros(int(ref["freq"]["height"].max()) + 1, np.float32)
    freq[ref["freq"]["height"]] = ref["freq"]["freq"]
    assert np.array_equal(bits(o.freq()), bits(freq))
    assert np.array_equal(o.grouped_pre(), ref["grouped_pre"])
    assert o.n_bins() == ref["grouped_post"].shape[1]
    assert np.array_equal(bits(o.grouped_post()), bits(ref["grouped_post"]))
    s = o.sigs()
    assert len(s["pos"]) == len(ref["sigs"])
    for k in ("ov", "contig", "pos", "bin_pre", "bin", "local", "bias"):
        assert np.array_equal(s[k], ref["sigs"][k]), k
    for k in ("rp", "rm", "fdr"):
        assert np.array_equal(bits(s[k]), bits(ref["sigs"][k])), k
    t = ref["transposons"]
    iv = np.zeros(len(t), IV_DTYPE)
    for k in ("contig", "strand", "start", "end"):
        iv[k] = t[k]
    tp = o.transposons(iv)
    for k in ("p", "q", "reads_plus", "reads_minus"):
        assert np.array_equal(bits(tp[k]), bits(t[k])), k
    assert np.array_equal(bits(tp["hist"][:, :W]), bits(t["hist"]))


def _random_sam(rng, n_records, contigs):
    """Adversarial but well-formed records: every CIGAR operation, soft and hard clips at either end, reference skips,
    lengths around the 24..32 filter, mixed-case bases, both strands, flags the reference ignores, NH tags in any
    position among other tags.  (Kept out: SEQ shorter than clip + 10, which the reference reads out of bounds.)"""
    lines = ["@HD\tVN:1.0"] + [f"@SQ\tSN:{n}\tLN:{L}" for n, L in contigs]
    for i in range(n_records):
        name, L = contigs[rng.integers(len(contigs))]
        flag = int(rng.choice([0, 16, 16, 0, 256, 272, 1024, 1040, 4, 20]))
        pos = 0 if flag & 4 and rng.random() < 0.5 else int(rng.integers(1, max(2, L - 40)))
        ops, query = [], 0
        if rng.random() < 0.1:
            ops.append((int(rng.integers(1, 4)), "H"))
        if rng.random() < 0.3:
            k = int(rng.integers(1, 6)); ops.append((k, "S")); query += k
        core = int(rng.integers(20, 36))       # aligned length lands on both sides of the 24..32 filter
        first = True
        while core > 0:
            k = int(min(core, rng.integers(1, 30))) if not first else int(min(core, rng.integers(12, 30)))
            op = "M" if first else str(rng.choice(["M", "M", "=", "X", "D", "N"]))
            first = False
            ops.append((k, op)); core -= k
            if op in "M=X":
                query += k
            if rng.random() < 0.2:
                j = int(rng.integers(1, 4)); ops.append((j, "I")); query += j
            if rng.random() < 0.05:
                ops.append((int(rng.integers(1, 3)), "P"))
        if ops[-1][1] in "IP":
            ops.append((1, "M")); query += 1
        if rng.random() < 0.3:
            k = int(rng.integers(1, 6)); ops.append((k, "S")); query += k
        if rng.random() < 0.1:
            ops.append((int(rng.integers(1, 4)), "H"))
        cigar = "".join(f"{k}{op}" for k, op in ops)
        seq = "".join(rng.choice(list("ACGTacgtNn"), query))
        tags = []
        if rng.random() < 0.5:
            tags.append("AS:i:%d" % rng.integers(0, 60))
        if rng.random() < 0.6:
            tags.append("NH:i:%d" % rng.integers(1, 13))
        if rng.random() < 0.5:
            tags.append("XS:Z:NH:i:9")
        rng.shuffle(tags)
        lines.append("\t".join([f"r{i}", str(flag), name, str(pos), "255", cigar, "*", "0", "0", seq, "*", *tags]))
    return "\n".join(lines) + "\n"


@pytest.mark.skipif(not have_ref(23), reason="oracle/_ref was not built here")
@pytest.mark.parametrize("mode", ["weighted", "unique", "discard"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_record_extraction_equals_the_reference_on_adversarial_records(tmp_path, mode, seed):
    """pingpongpro.cpp:415-484 (CIGAR length, NH weight, strand, 5'-end key, A10 with soft clips) as restated in
    host/frontend.h: the stacks built from OUR extraction must be the stacks the reference builds from the same file."""
    rng = np.random.default_rng(1000 + seed)
    sam = tmp_path / "adv.sam"
    sam.write_text(_random_sam(rng, 6000, [("c1", 3000), ("c2", 900), ("c3", 120)]))
    ref = run_harness(str(tmp_path / "dump.bin"), [str(sam)], mode=mode)
    for threads in (1, 3):
        dec = host.decode_file(str(sam), mode=mode, threads=threads)
        st = Oracle(dec.reads, mode=mode).stacks()
        assert len(st["key"]) == len(ref["stacks"]) and len(st["key"]) > 500
        for k in ("strand", "contig", "key", "a10"):
            assert np.array_equal(st[k], ref["stacks"][k]), k
        assert np.array_equal(bits(st["reads"]), bits(ref["stacks"]["reads"]))
        assert dec.total_weight == ref["total"]


def _sam_to_bam(sam_text: str, block_bytes: int = 9000) -> bytes:
    """Minimal SAM -> BAM encoder for the tests (integer tags typed like samtools does: smallest type that fits)."""
    import struct
    import zlib

    header = [ln for ln in sam_text.splitlines() if ln.startswith("@")]
    refs = [(f.split("\t")[1][3:], int(f.split("\t")[2][3:])) for f in header if f.startswith("@SQ")]
    rid = {n: i for i, (n, _) in enumerate(refs)}
    text = ("\n".join(header) + "\n").encode()
    out = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for n, L in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", L)
    for ln in sam_text.splitlines():
        if ln.startswith("@") or not ln:
            continue
        f = ln.split("\t")
        name, flag, rname, pos, cigar, seq = f[0], int(f[1]), f[2], int(f[3]), f[5], f[9]
        ops, num = [], ""
        for ch in cigar:
            if ch.isdigit():
                num += ch
            else:
                ops.append((int(num) << 4) | "MIDNSHP=X".index(ch)); num = ""
        codes = ["=ACMGRSVTWYHKDBN".index(c.upper()) for c in seq]
        packed = bytes(((codes[i] << 4) | (codes[i + 1] if i + 1 < len(codes) else 0)) for i in range(0, len(codes), 2))
        tags = b""
        for t in f[11:]:
            key, typ, val = t.split(":", 2)
            if typ == "i":
                v = int(val)
                for code, fmt, lo, hi in (("C", "<B", 0, 255), ("c", "<b", -128, 127), ("S", "<H", 0, 65535), ("s", "<h", -32768, 32767), ("I", "<I", 0, 2**32 - 1), ("i", "<i", -2**31, 2**31 - 1)):
                    if lo <= v <= hi:
                        tags += key.encode() + code.encode() + struct.pack(fmt, v)
                        break
            else:
                tags += key.encode() + b"Z" + val.encode() + b"\0"
        body = struct.pack("<iiBBHHHIiii", rid.get(rname, -1), pos - 1, len(name) + 1, 255, 4680, len(ops), flag, len(seq), -1, -1, 0)
        body += name.encode() + b"\0" + b"".join(struct.pack("<I", o) for o in ops) + packed + b"\xff" * len(seq) + tags
        out += struct.pack("<i", len(body)) + body
    bam = b""
    for o in list(range(0, len(out), block_bytes)) + [None]:
        chunk = b"" if o is None else out[o:o + block_bytes]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        z = c.compress(chunk) + c.flush()
        bam += bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + (len(z) + 25).to_bytes(2, "little") + z + zlib.crc32(chunk).to_bytes(4, "little") + len(chunk).to_bytes(4, "little")
    return bam


@pytest.mark.parametrize("mode", ["weighted", "unique"])
def test_sam_and_bam_parsers_agree_on_adversarial_records(tmp_path, mode):
    """Two independent code paths (text fields vs binary fields, letters vs 4-bit codes, NH:i text vs typed binary
    integers) must extract the same reads; blocks of 9000 bytes make records straddle BGZF blocks."""
    rng = np.random.default_rng(77)
    text = _random_sam(rng, 5000, [("c1", 3000), ("c2", 900), ("c3", 120)])
    (tmp_path / "a.sam").write_text(text)
    (tmp_path / "a.bam").write_bytes(_sam_to_bam(text))
    want = host.decode_file(str(tmp_path / "a.sam"), mode=mode)
    assert want.n_kept > 1000
    for threads in (1, 4):
        got = host.decode_file(str(tmp_path / "a.bam"), mode=mode, threads=threads)
        assert got.names == want.names and got.lengths == want.lengths
        assert got.n_records == want.n_records and got.n_kept == want.n_kept and got.total_weight == want.total_weight
        for a, b in zip(got.reads, want.reads):
            assert np.array_equal(a, b)


@pytest.mark.skipif(not have_ref(23), reason="oracle/_ref was not built here")
@pytest.mark.parametrize("rng_nt,seed", [(30, 201), (60, 202), (300, 203), (5000, 204)])
def test_predicted_transposons_equal_the_reference(tmp_path, rng_nt, seed):
    """-T RANGE without -t (pingpongpro.cpp:1324-1357): the clustering of the host library on the oracle's overlap-10
    loci, scored by the oracle, against the reference's own prediction + scoring -- including the cluster that the
    reference never emits at the end of a contig and the > 30 nt rule."""
    sam = str(tmp_path / "in.sam")
    subprocess.check_call([SYNTH, "--contigs", "c1:90000,c2:61000,c3:30000,c4:2000", "--reads", "60000", "--seed", str(seed), "--out", sam])
    ref = run_harness(str(tmp_path / "dump.bin"), [sam], mode="weighted", predict=rng_nt)
    dec = host.decode_file(sam, mode="weighted")
    o = Oracle(dec.reads, mode="weighted")
    s = o.sigs()
    pp = s["ov"] == 10 - 3  # overlap index of the ping-pong overlap
    order = np.lexsort((s["pos"][pp], s["contig"][pp]))
    iv, ids = host.predict_transposons(s["contig"][pp][order], s["pos"][pp][order], dec.names, rng_nt)
    t = ref["predicted"]
    assert len(iv) == len(t)
    if rng_nt <= 300:
        assert len(iv) > 3
    for k in ("contig", "strand", "start", "end"):
        assert np.array_equal(iv[k], t[k]), k
    assert all(i == f"{dec.names[c]}:{a}-{b}" for i, c, a, b in zip(ids, iv["contig"], iv["start"], iv["end"]))
    q = np.zeros(len(iv), IV_DTYPE)
    for k in ("contig", "strand", "start", "end"):
        q[k] = iv[k]
    tp = o.transposons(q)
    for k in ("p", "q", "reads_plus", "reads_minus"):
        assert np.array_equal(bits(tp[k]), bits(t[k])), k
    assert np.array_equal(bits(tp["hist"][:, :21]), bits(t["hist"]))
