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
