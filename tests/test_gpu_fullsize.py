"""BASELINE.json full sizes (-m gpu): config 2 (50 M reads, dm6-sized genome) and config 4 (500 M reads, human-sized
genome, one GPU) checked through size-independent properties, where the CPU oracle would take minutes:
  * checksum of checksums: sum_h h * freq[h] == reads kept (unique mode: every height is an exact count),
    sum_h freq[h] == number of non-empty stack words;
  * every signature is counted exactly once: sum(grouped) == signatures, sum(grouped[overlap 10]) == loci at -s 0;
  * sortedness: signatures ascend in (contig, position, overlap), loci in (contig, position);
  * idempotence: a second scan + score over the same stacks returns identical tables and loci;
  * monotone -s filter: loci(s) is exactly the subset of loci(0) with both heights >= s;
  * order independence in unique mode: reversing the read order changes nothing;
  * a 2-way range-sharded run reproduces the single-context loci (config 2)."""
import numpy as np
import pytest

from pingpongpro_b200 import api, host

pytestmark = pytest.mark.gpu


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


def test_config4_human_500m_single_gpu():
    m = host.SynthModel("human", n_reads=500_000_000, seed=4)
    reads = m.generate(mode="unique")
    n_kept = sum(len(r) for r in reads)
    with api.Context(m.lengths, mode="unique") as ctx:
        words, owned = ctx.layout()
        assert owned == m.genome_size and words * 8 > 24e9  # 24.8 GB of stack arrays on one B200
        build(ctx, reads)
        check_properties(ctx, reads, "unique", n_kept)
