"""BASELINE.json full sizes (-m gpu): config 2 (50 M reads, dm6-sized genome) and config 4 (500 M reads, human-sized
genome, one GPU) checked through size-independent properties, where the CPU oracle would take minutes:
  * checksum of checksums: sum_h h * freq[h] == reads kept (unique mode: every height is an exact count),
    sum_h freq[h] == number of non-empty stack words;
  * every signature is counted exactly once: sum(grouped) == signatures, sum(grouped[overlap 10]) == loci at -s 0;
  * sortedness: signatures ascend in (contig, position, overlap), loci in (contig, position);
  * idempotence: a second scan + score over the same stacks returns identical tables and loci;
  * monotone -s filter: loci(s) is exactly the subset of loci(0) with both heights >= s;
  * order independence in unique mode: reversing the read order changes nothing;
  * a 2-way range-sharded run reproduces the single-context loci (config 2);
  * config 3 (200 M reads, mouse-sized genome, annotation): per-transposon sums recomputed from the signatures;
  * config 5 (the 50 M-read set, background range 3..30, -s 1..20): same properties, same overlap-10 loci.
On top of the properties every configuration is compared BIT FOR BIT with the unmodified reference through SHA-256 digests
of the reference's own full-size intermediates (check_reference_digests; tests/golden/fullsize/)."""
import os
import subprocess

import numpy as np
import pytest

import fullsize_digest as fd
from pingpongpro_b200 import _build, api, host

pytestmark = pytest.mark.gpu


def check_reference_digests(ctx, m, case, freq, grouped, sc, annotation_dir=None):
    """Bit-exact comparison with the UNMODIFIED reference at full size: SHA-256 digests of the reference's own
    intermediates (tests/golden/fullsize/<case>.json, written by `oracle/_ref/ppp_harness digest` on the CPU container,
    tests/golden/make_fullsize_digests.py) against the same byte streams taken from the CUDA path through the C ABI:
    stacks (:91-105, :458-488), height frequencies (:502-509), grouped counts before (:605) and after the collapse
    (:631-686), every signature with its bins and FDR (:608, :750-753), the reported loci for several -s (:903) and,
    with an annotation, the per-transposon sums (:1221-1240) and p / q values (:1266-1312)."""
    want = fd.load(case)
    W = want["max_overlap"] - want["min_overlap"] + 1
    sh = fd.StackHasher()
    for strand in (0, 1):
        for c, L in enumerate(m.lengths):
            h, a = ctx.stacks(c, strand, 0, L + 64)
            sh.add_dense(strand, c, h, a)
    assert (sh.n, sh.hexdigest()) == (want["n_stacks"], want["stacks"])
    assert fd.digest_freq(freq) == (want["n_heights"], want["freq"])
    assert fd.digest_grouped_counts(grouped) == want["grouped_pre"]
    assert sc.n_bins == want["grouped_post_bins"] and fd.sha(sc.cells) == want["grouped_post"]
    sig = ctx.signatures()
    assert sig.dtype == fd.PAIR_DTYPE
    assert (len(sig), fd.digest_signatures(sig)) == (want["n_signatures"], want["signatures"])
    del sig
    for smin, w in want["loci"].items():
        loci = sc.loci if int(smin) == 0 else ctx.calculate_fdrs(int(smin), want_tables=False).loci
        assert (len(loci), fd.digest_loci(loci)) == (w["n"], w["sha256"]), smin
    if "n_transposons" in want:
        tool = os.path.join(_build.BIN, "ppp_synth")
        path = os.path.join(str(annotation_dir), "annotation.tsv")
        subprocess.check_call([tool, "--genome", want["genome"], "--reads", str(want["reads"]), "--seed", str(want["seed"]), "--annot", path], stdout=subprocess.DEVNULL)
        iv, _, _ = host.read_annotation([path], m.names)
        ctx.calculate_fdrs(0, want_tables=False, want_loci=False)
        tp, order = ctx.find_suppressed_transposons(iv)
        srt = iv[order]
        assert len(tp) == want["n_transposons"]
        assert fd.digest_transposon_sums(srt["contig"], srt["strand"], srt["start"], srt["end"], tp["reads_plus"], tp["reads_minus"], tp["histogram"], W) == want["transposon_sums"]
        pq = fd.load_pq(case)
        # the p-value is an fp64 sum narrowed to float (:1271): identical unless the double lands within ~1e-16 of a rounding tie
        assert np.allclose(tp["p_value"], pq[:, 0], rtol=2e-7, atol=0) and np.allclose(tp["q_value"], pq[:, 1], rtol=4e-7, atol=0)
        same = np.count_nonzero(tp["p_value"].view(np.uint32) == pq[:, 0].view(np.uint32))
        assert same >= len(tp) - 2, (same, len(tp))
        return tp
    return None


def build(ctx, reads, reverse=False):
    ctx.reset()
    for c, r in enumerate(reads):
        if len(r):
            ctx.count_reads(c, r[::-1].copy() if reverse else r)
    ctx.finish()


def pass_over(ctx, s=0):
    freq, max_h = ctx.map_heights_to_scores()
    grouped = ctx.count_stacks_by_group()
    sc = ctx.calculate_fdrs(s)
    return freq, max_h, grouped, sc


def check_properties(ctx, reads, mode, n_kept):
    freq, max_h, grouped, sc = pass_over(ctx)
    assert len(freq) == max_h + 1 and freq[max_h] > 0
    if mode == "unique":
        assert int((np.arange(max_h + 1, dtype=np.uint64) * freq).sum()) == n_kept
    assert int(grouped.sum()) == ctx.n_pairs
    assert int(grouped[10 - ctx.min_overlap].sum()) == len(sc.loci)
    key = sc.loci["contig"].astype(np.uint64) << np.uint64(32) | sc.loci["position"].astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    assert np.all((sc.loci["fdr"] >= 0) & (sc.loci["fdr"] <= 1))
    assert np.all(sc.loci["reads_plus"] > 0) and np.all(sc.loci["reads_minus"] > 0)
    # idempotence
    freq2, max_h2, grouped2, sc2 = pass_over(ctx)
    assert max_h2 == max_h and np.array_equal(freq2, freq) and np.array_equal(grouped2, grouped) and np.array_equal(sc2.loci, sc.loci)
    # -s filter
    for s in (2, 7):
        sub = ctx.calculate_fdrs(s).loci
        want = sc.loci[(sc.loci["reads_plus"] >= s) & (sc.loci["reads_minus"] >= s)]
        assert np.array_equal(sub, want)
    return freq, grouped, sc


@pytest.mark.parametrize("mode", ["unique", "weighted"])
def test_config2_dm6_50m(mode):
    m = host.SynthModel("dm6", n_reads=50_000_000, seed=2)
    reads = m.generate(mode=mode)
    n_kept = sum(len(r) for r in reads)
    with api.Context(m.lengths, mode=mode) as ctx:
        build(ctx, reads)
        freq, grouped, sc = check_properties(ctx, reads, mode, n_kept)
        check_reference_digests(ctx, m, "config2_" + mode, freq, grouped, sc)
        sig = ctx.signatures()
        k = (sig["contig"].astype(np.uint64) << np.uint64(40)) | (sig["position"].astype(np.uint64) << np.uint64(8)) | sig["overlap"].astype(np.uint64)
        assert np.all(k[1:] > k[:-1])
        if mode == "unique":  # order independence of the integer path
            build(ctx, reads, reverse=True)
            freq_r, _, grouped_r, sc_r = pass_over(ctx)
            assert np.array_equal(freq_r, freq) and np.array_equal(grouped_r, grouped) and np.array_equal(sc_r.loci, sc.loci)
    # 2-way range sharding, shards run one after the other; tables merged on the host exactly as the all-reduce would
    total = m.genome_size
    cut = total // 2
    parts = []
    for lo, hi in ((0, cut), (cut, total)):
        with api.Context(m.lengths, mode=mode, range_begin=lo, range_end=hi) as ctx:
            build(ctx, reads)  # every context keeps what it owns (+ halo) and counts the rest as foreign
            f, mh = ctx.map_heights_to_scores()
            parts.append((f, ctx.foreign_reads()))
    merged = np.zeros(max(len(p[0]) for p in parts), np.uint64)
    for f, _ in parts:
        merged[: len(f)] += f
    assert np.array_equal(merged, freq)
    assert sum(p[1] for p in parts) >= n_kept - 64  # each read is foreign to exactly one shard (halo copies excepted)


def test_config5_dm6_50m_background_range_30():
    """configs[4]: the 50 M-read set with the background overlap range widened to 3..30 (28 overlaps) and the -s sweep;
    the same size-independent properties must hold, and the overlap-10 loci must be those of the default range (the
    window only adds background overlaps: a locus exists iff both stacks exist, :580-608)."""
    m = host.SynthModel("dm6", n_reads=50_000_000, seed=2)
    reads = m.generate(mode="unique")
    n_kept = sum(len(r) for r in reads)
    with api.Context(m.lengths, mode="unique") as ctx:
        build(ctx, reads)
        _, _, sc23 = check_properties(ctx, reads, "unique", n_kept)
        pairs23 = ctx.n_pairs
    with api.Context(m.lengths, mode="unique", max_overlap=30) as ctx:
        assert ctx.W == 28
        build(ctx, reads)
        freq30, grouped, sc30 = check_properties(ctx, reads, "unique", n_kept)
        check_reference_digests(ctx, m, "config5_max30", freq30, grouped, sc30)
        assert ctx.n_pairs > pairs23 and grouped.shape[0] == 28
        for col in ("contig", "position", "reads_plus", "reads_minus"):
            assert np.array_equal(sc30.loci[col], sc23.loci[col])
        for s in range(1, 21):  # -s 1..20
            sub = ctx.calculate_fdrs(s).loci
            assert np.array_equal(sub, sc30.loci[(sc30.loci["reads_plus"] >= s) & (sc30.loci["reads_minus"] >= s)])
    with api.Context(m.lengths, mode="unique", max_overlap=26) as ctx:
        build(ctx, reads)
        freq26, _, grouped26, sc26 = pass_over(ctx)
        check_reference_digests(ctx, m, "config5_max26", freq26, grouped26, sc26)


@pytest.mark.parametrize("mode", ["unique", "weighted"])
def test_config4_human_500m_single_gpu(mode):
    m = host.SynthModel("human", n_reads=500_000_000, seed=4)
    reads = m.generate(mode=mode)
    n_kept = sum(len(r) for r in reads)
    with api.Context(m.lengths, mode=mode) as ctx:
        words, owned = ctx.layout()
        assert owned == m.genome_size and words * 8 > 24e9  # 24.8 GB of stack arrays on one B200
        build(ctx, reads)
        del reads
        freq, grouped, sc = check_properties(ctx, None, mode, n_kept)
        check_reference_digests(ctx, m, "config4_" + mode, freq, grouped, sc)


def test_config3_mouse_200m_annotation(tmp_path):
    """configs[2]: 200 M reads on the mouse-sized genome with a transposon annotation.  The per-transposon sums of
    findSuppressedTransposons (pingpongpro.cpp:1221-1240) are recomputed for a sample of the (non-overlapping, sorted)
    intervals straight from the downloaded signatures with the literal single-precision loop."""
    m = host.SynthModel("mouse", n_reads=200_000_000, seed=3)
    reads = m.generate(mode="weighted")
    ivs = m.intervals()
    iv = np.zeros(len(ivs), api.INTERVAL_DTYPE)
    iv["contig"], iv["start"], iv["end"] = ivs[:, 0], ivs[:, 1], ivs[:, 2] - 1
    with api.Context(m.lengths, mode="weighted") as ctx:
        build(ctx, reads)
        del reads
        freq, max_h, grouped, sc = pass_over(ctx)
        assert int(grouped.sum()) == ctx.n_pairs and int(grouped[10 - ctx.min_overlap].sum()) == len(sc.loci)
        check_reference_digests(ctx, m, "config3_weighted", freq, grouped, sc, annotation_dir=tmp_path)
        ctx.calculate_fdrs(0, want_tables=False, want_loci=False)
        tp, order = ctx.find_suppressed_transposons(iv)
        sig = ctx.signatures()
    assert len(tp) == len(iv) and np.all((tp["p_value"] >= 0) & (tp["p_value"] <= 1)) and np.all((tp["q_value"] >= 0) & (tp["q_value"] <= 1))
    srt = iv[order]  # results come back in (contig, start, end) order
    assert np.all((srt["contig"][1:] > srt["contig"][:-1]) | ((srt["contig"][1:] == srt["contig"][:-1]) & (srt["start"][1:] >= srt["start"][:-1])))
    W, pp = 21, 10 - 3
    key = sig["contig"].astype(np.uint64) << np.uint64(32) | sig["position"].astype(np.uint64)
    rng = np.random.default_rng(1)
    checked = 0
    for k in rng.choice(len(srt), size=400, replace=False):
        t = srt[k]
        lo = np.searchsorted(key, (np.uint64(t["contig"]) << np.uint64(32)) | np.uint64(t["start"]), "left")
        hi = np.searchsorted(key, (np.uint64(t["contig"]) << np.uint64(32)) | np.uint64(t["end"]), "right")
        s = sig[lo:hi]
        hist = np.zeros(W, np.float32)
        rp = rm = np.float32(0)
        for o in range(W):
            so = s[s["overlap"] == o + 3]
            h = np.float32(0)
            for a, b, f in zip(so["reads_plus"], so["reads_minus"], so["fdr"]):
                h = np.float32(h + np.float32(np.float32(a + b) * np.float32(np.float32(1) - f)))  # :1227
            hist[o] = h
            if o == pp:
                for a, b in zip(so["reads_plus"], so["reads_minus"]):
                    rp, rm = np.float32(rp + a), np.float32(rm + b)                                    # :1230-1234
        got = tp[k]
        assert np.array_equal(got["histogram"][:W].view(np.uint32), hist.view(np.uint32)), (k, got["histogram"][:W], hist)
        assert np.float32(got["reads_plus"]) == rp and np.float32(got["reads_minus"]) == rm
        checked += len(s) > 0
    assert checked > 100
