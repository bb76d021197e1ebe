"""The digest pipeline itself, on the CPU: the C restatement (oracle/ppp_oracle.c) hashed with tests/fullsize_digest.py
must reproduce the digests that `oracle/_ref/ppp_harness digest` (the unmodified reference) wrote for the small cases
of tests/golden/make_fullsize_digests.py.  The full-size digests in the same directory are then compared with the CUDA
path by tests/test_gpu_fullsize.py."""
import os
import subprocess

import numpy as np
import pytest

import fullsize_digest as fd
from oracle_tools import IV_DTYPE, Oracle
from pingpongpro_b200 import _build, host

CASES = {
    "small_tiny_weighted": dict(genome="tiny", reads=60_000, seed=42, mode="weighted", annotation=True),
    "small_dm6_weighted": dict(genome="dm6", reads=400_000, seed=9, mode="weighted", annotation=True),
    "small_dm6_unique_max30": dict(genome="dm6", reads=400_000, seed=9, mode="unique", max_overlap=30),
}


def synth_annotation(c, names, tmp_path):
    tool = os.path.join(_build.BIN, "ppp_synth")
    path = str(tmp_path / "annotation.tsv")
    subprocess.check_call([tool, "--genome", c["genome"], "--reads", str(c["reads"]), "--seed", str(c["seed"]), "--annot", path], stdout=subprocess.DEVNULL)
    iv, _, _ = host.read_annotation([path], names)
    return iv


@pytest.mark.parametrize("case", sorted(CASES))
def test_port_reproduces_reference_digests(case, tmp_path):
    c = CASES[case]
    want = fd.load(case)
    m = host.SynthModel(c["genome"], n_reads=c["reads"], seed=c["seed"])
    reads = m.generate(mode=c["mode"])
    mo = c.get("max_overlap", 23)
    assert (want["min_overlap"], want["max_overlap"], want["mode"]) == (3, mo, c["mode"])
    o = Oracle(reads, mode=c["mode"], max_ov=mo).score()
    W = mo - 3 + 1

    st = o.stacks()
    order = np.lexsort((st["key"], st["contig"], st["strand"]))
    sh = fd.StackHasher()
    sh.add_records(st["strand"][order], st["contig"][order], st["key"][order], st["reads"][order], st["a10"][order])
    assert (sh.n, sh.hexdigest()) == (want["n_stacks"], want["stacks"])

    assert fd.digest_freq(o.freq()) == (want["n_heights"], want["freq"])
    assert fd.digest_grouped_counts(o.grouped_pre()) == want["grouped_pre"]
    assert o.n_bins() == want["grouped_post_bins"] and fd.digest_grouped_counts(o.grouped_post()) == want["grouped_post"]

    s = o.sigs()
    order = np.lexsort((s["ov"], s["pos"], s["contig"]))
    sig = np.empty(len(order), fd.PAIR_DTYPE)
    sig["contig"], sig["position"], sig["overlap"] = s["contig"][order], s["pos"][order], s["ov"][order] + 3
    sig["height_score_bin"], sig["collapsed_bin"], sig["local_bin"], sig["bias_bin"] = s["bin_pre"][order], s["bin"][order], s["local"][order], s["bias"][order]
    sig["reads_plus"], sig["reads_minus"], sig["fdr"] = s["rp"][order], s["rm"][order], s["fdr"][order]
    assert (len(sig), fd.digest_signatures(sig)) == (want["n_signatures"], want["signatures"])

    pp = sig[sig["overlap"] == 10]
    for smin, w in want["loci"].items():
        keep = pp[(pp["reads_plus"] >= np.float32(int(smin))) & (pp["reads_minus"] >= np.float32(int(smin)))]
        loci = np.empty(len(keep), fd.LOCUS_DTYPE)
        for a, b in (("contig", "contig"), ("position", "position"), ("fdr", "fdr"), ("reads_plus", "reads_plus"), ("reads_minus", "reads_minus")):
            loci[a] = keep[b]
        assert (len(loci), fd.digest_loci(loci)) == (w["n"], w["sha256"]), smin

    if c.get("annotation"):
        iv = synth_annotation(c, m.names, tmp_path)
        tp = o.transposons(np.ascontiguousarray(iv).view(IV_DTYPE))
        assert len(tp) == want["n_transposons"]
        assert fd.digest_transposon_sums(tp["contig"], tp["strand"], tp["start"], tp["end"], tp["reads_plus"], tp["reads_minus"], tp["hist"], W) == want["transposon_sums"]
        pq = fd.load_pq(case)
        assert np.array_equal(tp["p"].view(np.uint32), pq[:, 0].view(np.uint32)) and np.array_equal(tp["q"].view(np.uint32), pq[:, 1].view(np.uint32))
        assert fd.sha(np.stack([tp["p"], tp["q"]], axis=1)) == want["transposon_pq"]
