"""Pins the CPU restatement (oracle/ppp_oracle.c) against outputs of the unmodified reference
(committed fixtures, tests/golden/make_golden.py): every intermediate must agree bit for bit."""
import numpy as np
import pytest

from golden_tools import Golden, case_names
from oracle_tools import IV_DTYPE, Oracle


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module", params=case_names())
def pair(request):
    g = Golden(request.param)
    return g, Oracle(g.reads, mode=g.mode, min_ov=g.min_ov, max_ov=g.max_ov, pp_ov=g.pp_ov)


def test_stacks_bit_exact(pair):
    g, o = pair
    st = o.stacks()
    assert len(st["key"]) == len(g.stacks)
    for k in ("strand", "contig", "key", "a10"):
        assert np.array_equal(st[k], g.stacks[k]), k
    assert np.array_equal(bits(st["reads"]), bits(g.stacks["reads"]))
    assert o.total() == g.total


def test_height_frequencies(pair):
    g, o = pair
    assert np.array_equal(bits(o.freq()), bits(g.dense_freq()))


def test_grouped_counts_and_collapse(pair):
    g, o = pair
    assert np.array_equal(o.grouped_pre(), g.grouped_pre)
    assert o.n_bins() == g.grouped_post.shape[1]
    assert np.array_equal(bits(o.grouped_post()), bits(g.grouped_post))


def test_signatures_and_fdr(pair):
    g, o = pair
    s = o.sigs()
    assert len(s["pos"]) == len(g.sigs)
    for k in ("ov", "contig", "pos", "bin_pre", "bin", "local", "bias"):
        assert np.array_equal(s[k], g.sigs[k]), k
    for k in ("rp", "rm", "fdr"):
        assert np.array_equal(bits(s[k]), bits(g.sigs[k])), k


def test_transposons(pair):
    g, o = pair
    if g.transposons is None:
        pytest.skip("case has no annotation")
    t = g.transposons
    iv = np.zeros(len(t), IV_DTYPE)
    for k in ("contig", "strand", "start", "end"):
        iv[k] = t[k]
    tp = o.transposons(iv)
    for k in ("p", "q", "reads_plus", "reads_minus"):
        assert np.array_equal(bits(tp[k]), bits(t[k])), k
    assert np.array_equal(bits(tp["hist"][:, :g.W]), bits(t["hist"]))
