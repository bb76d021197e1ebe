"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against
  * the committed golden fixtures (outputs of the unmodified reference), and
  * the CPU restatement oracle/ppp_oracle.c on seeded synthetic inputs,
bit-exact for stacks / histograms / bins / counts / signatures, and (as it turns out) for the fp32 FDR too.
Tolerance for the fp64 quantities (z, p before narrowing): 1e-9 relative (BASELINE.json north_star)."""
import threading

import numpy as np
import pytest

from golden_tools import Golden, case_names
from oracle_tools import IV_DTYPE, Oracle
from pingpongpro_b200 import api, dist, host

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run_gpu(lengths, reads, mode, min_ov=3, max_ov=23, pp_ov=10, min_stack_height=0, intervals=None, **kw):
    return api.run_pipeline(lengths, reads, mode=mode, min_overlap=min_ov, max_overlap=max_ov, pingpong_overlap=pp_ov,
                            min_stack_height=min_stack_height, intervals=intervals, **kw)


def gpu_stacks(lengths, reads, mode, **kw):
    """All non-empty stacks as the reference's dump lists them: (strand, contig, key, reads, a10)."""
    rows = []
    with api.Context(lengths, mode=mode, **kw) as ctx:
        for c, r in enumerate(reads):
            if len(r):
                ctx.count_reads(c, r)
        ctx.finish()
        for strand in (0, 1):
            for c, L in enumerate(lengths):
                h, a = ctx.stacks(c, strand, 0, L + 64)
                nz = np.flatnonzero(h)
                rows.append((np.full(len(nz), strand, np.uint32), np.full(len(nz), c, np.uint32), nz.astype(np.uint32), h[nz], a[nz].astype(np.uint32)))
    return [np.concatenate([r[i] for r in rows]) for i in range(5)]


def check_against(res, ref_sigs, ref_freq_dense, ref_grouped_pre, ref_grouped_post, W, pp_idx, min_ov):
    # height frequencies: exact counts, the reference's float counter is min(count, 2^24)
    f = np.minimum(res["freq"], 1 << 24).astype(np.float32)
    n = max(len(f), len(ref_freq_dense))
    a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
    a[: len(f)] = f
    b[: len(ref_freq_dense)] = ref_freq_dense
    assert np.array_equal(a, b)
    # grouped counts before collapse
    assert np.array_equal(np.minimum(res["grouped"], 1 << 24).astype(np.float32), ref_grouped_pre)
    sc = res["score"]
    assert sc.n_bins == ref_grouped_post.shape[1]
    assert np.array_equal(bits(sc.cells), bits(ref_grouped_post))
    # signatures: GPU order is (contig, position, overlap); reference lists are (overlap, contig, position)
    s = res["signatures"]
    assert len(s) == len(ref_sigs)
    order = np.lexsort((s["position"], s["contig"], s["overlap"]))
    s = s[order]
    assert np.array_equal(s["overlap"] - min_ov, ref_sigs["ov"])
    for mine, theirs in (("contig", "contig"), ("position", "pos"), ("height_score_bin", "bin_pre"), ("collapsed_bin", "bin"),
                         ("local_bin", "local"), ("bias_bin", "bias")):
        assert np.array_equal(s[mine], ref_sigs[theirs]), mine
    assert np.array_equal(bits(s["reads_plus"]), bits(ref_sigs["rp"]))
    assert np.array_equal(bits(s["reads_minus"]), bits(ref_sigs["rm"]))
    assert np.array_equal(bits(s["fdr"]), bits(ref_sigs["fdr"]))
    # FDR table cells that are referenced by a signature
    fdr_tab = sc.fdr_table[ref_sigs["ov"], ref_sigs["bin"], ref_sigs["bias"], ref_sigs["local"]]
    assert np.array_equal(bits(fdr_tab), bits(ref_sigs["fdr"]))
    # reported loci = overlap-10 signatures in (contig, position) order (min_stack_height = 0)
    sel = ref_sigs[ref_sigs["ov"] == pp_idx]
    loci = sc.loci
    assert len(loci) == len(sel)
    assert np.array_equal(loci["contig"], sel["contig"]) and np.array_equal(loci["position"], sel["pos"])
    assert np.array_equal(bits(loci["fdr"]), bits(sel["fdr"]))
    assert np.array_equal(bits(loci["reads_plus"]), bits(sel["rp"])) and np.array_equal(bits(loci["reads_minus"]), bits(sel["rm"]))


def check_transposons(tp, order, ref, W):
    assert np.array_equal(bits(tp["histogram"][:, :W]), bits(ref["hist"][:, :W]))
    assert np.array_equal(bits(tp["reads_plus"]), bits(ref["reads_plus"]))
    assert np.array_equal(bits(tp["reads_minus"]), bits(ref["reads_minus"]))
    # z and p in fp64, before the narrowing store of :1271, against the oracle's doubles: 1e-9 relative (north_star)
    if "z64" in ref.dtype.names:  # (the oracle; the reference's own dumps hold the narrowed floats only)
        live = ~np.isnan(ref["z64"])
        assert np.array_equal(np.isnan(tp["z"]), ~live) and np.all(tp["p"][~live] == 1.0)
        assert np.allclose(tp["z"][live], ref["z64"][live], rtol=REL_TOL, atol=0)
        assert np.allclose(tp["p"], ref["p64"], rtol=REL_TOL, atol=0)
    # the narrowed float may differ by one ulp when the fp64 sum lies within ~1e-16 of a rounding tie
    assert np.allclose(tp["p_value"], ref["p"], rtol=2e-7, atol=0)
    assert np.allclose(tp["q_value"], ref["q"], rtol=4e-7, atol=0)


@pytest.mark.parametrize("name", case_names())
def test_golden(name):
    g = Golden(name)
    st = gpu_stacks(g.contig_lengths, g.reads, g.mode, min_overlap=g.min_ov, max_overlap=g.max_ov, pingpong_overlap=g.pp_ov)
    assert len(st[0]) == len(g.stacks)
    for col, key in zip(st, ("strand", "contig", "key")):
        assert np.array_equal(col, g.stacks[key]), key
    assert np.array_equal(bits(st[3]), bits(g.stacks["reads"]))
    assert np.array_equal(st[4], g.stacks["a10"])
    iv = None
    if g.transposons is not None:
        iv = np.zeros(len(g.transposons), api.INTERVAL_DTYPE)
        for k in ("contig", "strand", "start", "end"):
            iv[k] = g.transposons[k]
        iv = iv[np.random.RandomState(7).permutation(len(iv))]
    res = run_gpu(g.contig_lengths, g.reads, g.mode, g.min_ov, g.max_ov, g.pp_ov, intervals=iv)
    check_against(res, g.sigs, g.dense_freq(), g.grouped_pre, g.grouped_post, g.W, g.pp_ov - g.min_ov, g.min_ov)
    if iv is not None:
        tp, order = res["transposons"], res["transposon_order"]
        srt = iv[order]
        for k in ("contig", "start", "end", "strand"):
            assert np.array_equal(srt[k], g.transposons[k]), k
        check_transposons(tp, order, g.transposons, g.W)


def oracle_reference(lengths, reads, mode, min_ov=3, max_ov=23, pp_ov=10):
    o = Oracle(reads, mode=mode, min_ov=min_ov, max_ov=max_ov, pp_ov=pp_ov).score()
    s = o.sigs()
    dt = np.dtype([(k, s[k].dtype) for k in s])
    arr = np.zeros(len(s["pos"]), dt)
    for k in s:
        arr[k] = s[k]
    return o, arr


@pytest.mark.parametrize("mode", ["weighted", "unique", "discard"])
@pytest.mark.parametrize("genome,n_reads,seed", [("dm6", 400_000, 3), ("tiny", 150_000, 4)])
def test_seeded_vs_oracle(mode, genome, n_reads, seed):
    m = host.SynthModel(genome, n_reads=n_reads, seed=seed)
    reads = m.generate(mode=mode)
    o, sigs = oracle_reference(m.lengths, reads, mode)
    res = run_gpu(m.lengths, reads, mode)
    check_against(res, sigs, o.freq(), o.grouped_pre(), o.grouped_post(), 21, 7, 3)
    iv3 = m.intervals()
    iv = np.zeros(len(iv3), IV_DTYPE)
    iv["contig"], iv["start"], iv["end"] = iv3[:, 0], iv3[:, 1], iv3[:, 2] - 1
    ref = o.transposons(iv)
    with api.Context(m.lengths, mode=mode) as ctx:
        for c, r in enumerate(reads):
            ctx.count_reads(c, r)
        ctx.finish()
        ctx.map_heights_to_scores(False)
        ctx.count_stacks_by_group(False)
        ctx.calculate_fdrs(0, False, False)
        tp, order = ctx.find_suppressed_transposons(iv)
    check_transposons(tp, order, ref, 21)
    # z and p in fp64 against a numpy restatement of :1244-1269 on the (bit-identical) histograms
    h = ref["hist"][:, :21].astype(np.float32)
    bg = np.delete(h, 7, axis=1)
    mean = np.zeros(len(h), np.float32)
    for k in range(20):
        mean = (mean + bg[:, k]).astype(np.float32)
    mean = (mean / np.float32(20)).astype(np.float32)
    live = ~((mean == 0) & (h[:, 7] == 0))
    sd = np.zeros(len(h), np.float32)
    for k in range(20):
        d = (bg[:, k] - mean).astype(np.float32).astype(np.float64)
        sd = (sd.astype(np.float64) + d * d).astype(np.float32)
    sd = np.sqrt((1.0 / 19) * sd.astype(np.float64)).astype(np.float32)
    sd = np.where(sd.astype(np.float64) <= 1e-10, np.float32(1e-10), sd)
    z = ((h[:, 7] - mean).astype(np.float32) / sd).astype(np.float32).astype(np.float64)
    assert np.allclose(tp["z"][live], z[live], rtol=REL_TOL, atol=0)
    assert np.all(np.isnan(tp["z"][~live])) and np.all(tp["p"][~live] == 1.0)
    assert np.allclose(tp["p"][live].astype(np.float32), ref["p"][live], rtol=2e-7, atol=0)


def test_max_overlap_sweep_and_min_stack_height():
    """BASELINE.json config 5: -s 1..20 and background range up to 30 nt."""
    m = host.SynthModel(None, n_reads=120_000, seed=9, lengths=[60000, 40000])
    reads = m.generate(mode="weighted")
    for min_ov, max_ov in ((3, 23), (3, 26), (3, 30), (1, 32), (0, 31), (8, 12)):
        o, sigs = oracle_reference(m.lengths, reads, "weighted", min_ov, max_ov, 10)
        with api.Context(m.lengths, mode="weighted", min_overlap=min_ov, max_overlap=max_ov) as ctx:
            for c, r in enumerate(reads):
                ctx.count_reads(c, r)
            ctx.finish()
            ctx.map_heights_to_scores()
            g = ctx.count_stacks_by_group()
            assert np.array_equal(np.minimum(g, 1 << 24).astype(np.float32), o.grouped_pre())
            sel = sigs[sigs["ov"] == 10 - min_ov]
            for s in (0, 1, 2, 3, 5, 8, 13, 20):
                sc = ctx.calculate_fdrs(s)
                keep = sel[(sel["rp"] >= np.float32(s)) & (sel["rm"] >= np.float32(s))]
                assert len(sc.loci) == len(keep), (min_ov, max_ov, s)
                assert np.array_equal(sc.loci["position"], keep["pos"]) and np.array_equal(bits(sc.loci["fdr"]), bits(keep["fdr"]))


def test_pinned_double_buffer_and_batching_equal_single_submit():
    """Submitting a contig in many small batches (alternating pinned buffers) equals one submit, also for the
    order-dependent float accumulation of -m weighted."""
    m = host.SynthModel("tiny", n_reads=200_000, seed=10)
    reads = m.generate(mode="weighted")
    ref = gpu_stacks(m.lengths, reads, "weighted")
    bufs = [api.PinnedBuffer(7001), api.PinnedBuffer(7001)]
    rows = []
    with api.Context(m.lengths, mode="weighted") as ctx:
        k = 0
        for c, r in enumerate(reads):
            for o in range(0, len(r), 7001):
                chunk = r[o:o + 7001]
                b = bufs[k & 1]
                b.array[: len(chunk)] = chunk
                ctx.count_reads_raw(c, b.ptr.value, len(chunk))
                k += 1
        ctx.finish()
        for strand in (0, 1):
            for c, L in enumerate(m.lengths):
                h, a = ctx.stacks(c, strand, 0, L + 64)
                nz = np.flatnonzero(h)
                rows.append((h[nz], nz))
    got = np.concatenate([r[0] for r in rows])
    assert np.array_equal(bits(got), bits(ref[3]))


def test_packed_five_byte_upload_equals_eight_byte_records():
    """ppp_reads_submit_packed (u32 key + u8 meta with a 6-bit code into the call's NH table) builds the same stacks as
    ppp_reads_submit, also when a contig has more than 64 distinct NH values and is sent in several pieces."""
    m = host.SynthModel("tiny", n_reads=200_000, seed=10)
    reads = m.generate(mode="weighted")
    rng = np.random.default_rng(2)
    r0 = reads[0].copy()
    nh = rng.integers(1, 300, size=len(r0)).astype(np.uint32)            # ~300 distinct multiplicities in one contig
    r0["meta"] = (r0["meta"] & np.uint32(3)) | (nh << np.uint32(2))
    reads[0] = r0
    for mode in ("weighted", "unique", "discard"):
        ref = gpu_stacks(m.lengths, reads, mode)
        rows = []
        with api.Context(m.lengths, mode=mode) as ctx:
            for c, r in enumerate(reads):
                p = api.PackedReads(r)
                assert (len(p.pieces) > 1) == (c == 0)
                ctx.count_reads_packed(c, p)
                p_keep = p  # (pinned memory stays alive until finish)
                rows.append(p_keep)
            ctx.finish()
            got = []
            for strand in (0, 1):
                for c, L in enumerate(m.lengths):
                    h, a = ctx.stacks(c, strand, 0, L + 64)
                    nz = np.flatnonzero(h)
                    got.append((h[nz], a[nz]))
        assert np.array_equal(bits(np.concatenate([g[0] for g in got])), bits(ref[3]))
        assert np.array_equal(np.concatenate([g[1] for g in got]).astype(np.uint32), ref[4])


def test_weighted_depends_on_order_unique_does_not():
    m = host.SynthModel("tiny", n_reads=100_000, seed=11)
    reads = m.generate(mode="weighted")
    rev = [r[::-1].copy() for r in reads]
    a = gpu_stacks(m.lengths, reads, "unique")
    b = gpu_stacks(m.lengths, rev, "unique")
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # weighted: each order must equal the oracle run in that same order
    for rr in (reads, rev):
        st = Oracle(rr, mode="weighted").stacks()
        g = gpu_stacks(m.lengths, rr, "weighted")
        assert np.array_equal(bits(g[3]), bits(st["reads"]))


def test_empty_and_degenerate_inputs():
    lengths = [5000, 3000]
    empty = [np.empty(0, host.READ_DTYPE)] * 2
    res = run_gpu(lengths, empty, "weighted")
    assert res["max_height"] == 0 and len(res["signatures"]) == 0 and len(res["score"].loci) == 0
    # a single pair of stacks with unique heights: the reference computes 0/0
    r = host.pack_reads([100, 100, 110], [0, 0, 1], [0, 0, 0], [1, 1, 1])
    with pytest.raises(api.PppError) as e:
        run_gpu(lengths, [r, empty[1]], "unique")
    assert e.value.code == api.ERR_DEGENERATE
    # a key beyond the contig end is an error, not a silent drop
    bad = host.pack_reads([5000 + 500], [0], [0], [1])
    with api.Context(lengths, mode="unique") as ctx:
        ctx.count_reads(0, bad)
        with pytest.raises(api.PppError) as e:
            ctx.finish()
        assert e.value.code == api.ERR_RANGE
    # call order is enforced
    with api.Context(lengths) as ctx:
        with pytest.raises(api.PppError):
            ctx.count_stacks_by_group()


def test_key_at_contig_end_and_ragged_contigs():
    """A minus key may equal the contig length (:464); contigs shorter than one tile; maximum window reach."""
    lengths = [37, 2048, 2049, 5]
    rows = [
        (0, host.pack_reads([27, 27, 37, 37, 14, 14], [0, 0, 1, 1, 1, 1], [1, 0, 0, 0, 0, 1], [1] * 6)),   # overlap 10 at the very end, 37 = length
        (1, host.pack_reads([2047 - 23, 2047 - 23, 2047, 2047], [0, 0, 1, 1], [0] * 4, [1] * 4)),            # overlap 23 reaching the last position
        (2, host.pack_reads([2040, 2040, 2043, 2043, 2049, 2049], [0, 0, 1, 1, 1, 1], [0] * 6, [1] * 6)),    # overlaps 3 and 9 across a tile edge
        (3, host.pack_reads([1, 1], [0, 0], [0, 0], [1, 1])),
    ]
    reads = [np.empty(0, host.READ_DTYPE)] * 4
    for c, r in rows:
        reads[c] = r
    o, sigs = oracle_reference(lengths, reads, "unique")
    res = run_gpu(lengths, reads, "unique")
    assert len(sigs) == 4
    check_against(res, sigs, o.freq(), o.grouped_pre(), o.grouped_post(), 21, 7, 3)


def test_saturation_of_float_counters():
    """Float counters stop at 2^24: one stack with more than 2^24 reads in -m unique."""
    n = (1 << 24) + 1000
    lengths = [4096]
    r = np.empty(n + 4, host.READ_DTYPE)
    r["key"][:n] = 100
    r["meta"][:n] = 1 << 2
    r["key"][n:], r["meta"][n:] = [110, 110, 300, 300], [(1 << 2) | 1, (1 << 2) | 1, 1 << 2, 1 << 2]
    with api.Context(lengths, mode="unique") as ctx:
        ctx.count_reads(0, r)
        ctx.finish()
        h, _ = ctx.stacks(0, 0, 100, 1)
        assert h[0] == np.float32(1 << 24)
        freq, max_h = ctx.map_heights_to_scores()
        assert max_h == 1 << 24 and freq[1 << 24] == 1 and freq[2] == 2
        ctx.count_stacks_by_group()
        s = ctx.signatures()
        assert len(s) == 1 and s["reads_plus"][0] == np.float32(1 << 24) and s["overlap"][0] == 10


def _sharded(lengths, reads, mode, n_shards, min_stack_height=0, known_rank=False, pair_capacity=None, intervals=None):
    """n_shards contexts on one GPU, one host thread each; the collective hook is a host-side barrier exchange."""
    total = sum(lengths)
    cuts = [total * k // n_shards for k in range(n_shards + 1)]
    barrier = threading.Barrier(n_shards)
    slots = [None] * n_shards
    out = [None] * n_shards
    errs = []
    import torch

    def work(rank):
        try:
            cap = pair_capacity[rank] if pair_capacity else 0
            with api.Context(lengths, mode=mode, range_begin=cuts[rank], range_end=cuts[rank + 1], pair_capacity=cap) as ctx:
                def hook(ptr, count, dtype, op, stream):
                    torch.cuda.synchronize()
                    t = dist.device_tensor(ptr, count, dtype)
                    slots[rank] = t.cpu().numpy().copy()
                    barrier.wait()
                    acc = slots[0].copy()
                    for k in range(1, n_shards):
                        acc = np.maximum(acc, slots[k]) if op == 1 else acc + slots[k]
                    barrier.wait()
                    t.copy_(torch.from_numpy(acc))
                    torch.cuda.synchronize()
                ctx.set_allreduce(hook)
                if known_rank:
                    ctx.set_rank(rank, n_shards)
                for c, r in enumerate(reads):
                    if len(r):
                        ctx.count_reads(c, r)  # broadcast: every shard keeps what it owns (+ its halo)
                ctx.finish()
                ctx.map_heights_to_scores(False)
                ctx.count_stacks_by_group(False)
                sc = ctx.calculate_fdrs(min_stack_height)
                tp = ctx.find_suppressed_transposons(intervals) if intervals is not None else None  # (collective: every rank calls it)
                out[rank] = (sc, ctx.signatures(), ctx.foreign_reads(), tp)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            barrier.abort()

    th = [threading.Thread(target=work, args=(k,)) for k in range(n_shards)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("mode,n_shards,exchange", [("weighted", 2, "tables"), ("unique", 3, "lists"), ("weighted", 8, "lists"), ("weighted", 4, "overflow"),
                                                    ("weighted", 3, "grow")])
def test_range_sharding_equals_single_context(mode, n_shards, exchange, monkeypatch):
    """shards with a max-overlap halo + all-reduced histograms reproduce the single-GPU result;
    loci concatenated in rank order are globally sorted.  `exchange`: how the tall stacks' heights travel -- reduced
    height tables (rank unknown to the context), one compact list per rank (ppp_set_rank), or lists that overflow
    and fall back to the tables, or lists that overflow and are made longer (the pass repeats on every rank)."""
    m = host.SynthModel("tiny", n_reads=150_000, seed=12)
    reads = m.generate(mode=mode)
    single = run_gpu(m.lengths, reads, mode)
    assert single["max_height"] >= 512  # the exchange under test is that of the heights beyond the shared-memory bins
    if exchange == "overflow":
        monkeypatch.setenv("PPP_TALL_CAP", "3")
    if exchange == "grow":
        monkeypatch.setenv("PPP_TALL_CAP_START", "2")
    parts = _sharded(m.lengths, reads, mode, n_shards, known_rank=exchange != "tables")
    loci = np.concatenate([p[0].loci for p in parts])
    sigs = np.concatenate([p[1] for p in parts])
    assert np.array_equal(loci, single["score"].loci)
    assert np.array_equal(sigs, single["signatures"])
    for p in parts:
        assert p[0].n_bins == single["score"].n_bins
        assert np.array_equal(bits(p[0].fdr_table), bits(single["score"].fdr_table))


def test_shard_with_too_small_signature_store_makes_every_rank_rerun():
    """A rank whose signature store overflows cannot re-run its scan alone (the pass contains collectives): the overflow
    travels with the all-reduced grouped counts and every rank repeats the pass (api.cu)."""
    m = host.SynthModel("tiny", n_reads=150_000, seed=12)
    reads = m.generate(mode="weighted")
    single = run_gpu(m.lengths, reads, "weighted")
    assert len(single["signatures"]) > 3000
    for caps in ([0, 64, 0], [128, 0, 256]):
        parts = _sharded(m.lengths, reads, "weighted", 3, known_rank=True, pair_capacity=caps)
        assert np.array_equal(np.concatenate([p[0].loci for p in parts]), single["score"].loci)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), single["signatures"])
    # one context: the store grows and the pass is repeated
    small = run_gpu(m.lengths, reads, "weighted", pair_capacity=100)
    assert np.array_equal(small["signatures"], single["signatures"]) and np.array_equal(small["score"].loci, single["score"].loci)
    assert np.array_equal(small["grouped"], single["grouped"])


def test_tile_with_a_stack_at_every_position():
    """More than 4096 candidate plus stacks in one 8192-position tile (the block scan of the candidate counts
    must not wrap), and more candidates than one staging round holds; every window is full, so every plus stack has
    a record at each of the 21 overlaps."""
    L = 3 * 8192 + 100
    rng = np.random.default_rng(3)
    pos = np.arange(40, L - 40, dtype=np.uint32)
    key = np.concatenate([pos, pos, pos[::7], pos[::5]])
    minus = np.concatenate([np.zeros(len(pos)), np.ones(len(pos)), np.zeros(len(pos[::7])), np.ones(len(pos[::5]))]).astype(np.uint32)
    a10 = (rng.random(len(key)) < 0.3).astype(np.uint32)
    nh = rng.choice(np.array([1, 1, 1, 2, 3], dtype=np.uint32), size=len(key))
    order = rng.permutation(len(key))
    reads = [host.pack_reads(key[order], minus[order], a10[order], nh[order])]
    for mode in ("unique", "weighted"):
        o, sigs = oracle_reference([L], reads, mode)
        res = run_gpu([L], reads, mode)
        assert len(sigs) > 21 * 3 * 8000
        check_against(res, sigs, o.freq(), o.grouped_pre(), o.grouped_post(), 21, 7, 3)


def test_counters_beyond_2_pow_24_against_the_oracle():
    """The reference counts stacks per height (:508), signatures per cell (:605) and collapsed cells (:658) in FLOATS,
    which stop at 2^24 = 16 777 216; BASELINE.json's human-sized config has ~190 M singleton stacks.  Here: 2 x 17.3 M
    singleton stacks (height frequency 34.6 M), three overlaps (9..11) with 17.3 M signatures each in ONE cell, and a few
    taller stacks so that the collapse adds a small odd count to a saturated cell.  Compared with oracle/ppp_oracle.c."""
    n = (1 << 24) + 500_000
    L = n + 4096
    pos = np.arange(0, n, dtype=np.uint32)
    extra_p = np.repeat(np.arange(n + 100, n + 100 + 41 * 50, 50, dtype=np.uint32), 3)       # 41 plus stacks of height 3 ...
    extra_m = np.repeat(np.arange(n + 110, n + 110 + 41 * 50, 50, dtype=np.uint32), 2)       # ... each 10 nt from a minus stack of height 2
    key = np.concatenate([pos, pos + 9, extra_p, extra_m])
    minus = np.concatenate([np.zeros(n), np.ones(n), np.zeros(len(extra_p)), np.ones(len(extra_m))]).astype(np.uint32)
    reads = [host.pack_reads(key, minus, np.zeros(len(key), np.uint32), np.ones(len(key), np.uint32))]
    o, sigs = oracle_reference([L], reads, "unique", 9, 11, 10)
    res = run_gpu([L], reads, "unique", 9, 11, 10)
    # the saturation must actually have fired, in all three places
    assert res["freq"][1] > (1 << 24) and o.freq()[1] == np.float32(1 << 24)
    assert res["grouped"].max() > (1 << 24) and o.grouped_pre().max() == np.float32(1 << 24)
    assert res["score"].cells.max() > np.float32(1 << 24)  # collapse: 41 (an odd count: a rounding tie) + 2^24 in single precision (:658)
    check_against(res, sigs, o.freq(), o.grouped_pre(), o.grouped_post(), 3, 1, 9)


@pytest.mark.parametrize("n_shards", [2, 5])
def test_transposons_across_shards_in_separate_contexts(n_shards):
    """findSuppressedTransposons (:1195-1313) when every shard is driven on its own (one process per GPU): transposons that
    straddle a range boundary, span several ranges, overlap each other (the iterator history of App. D #4) or lie on a contig
    another rank holds must get the sums, p and q of the single-context run, on every rank."""
    m = host.SynthModel("tiny", n_reads=150_000, seed=21)
    reads = m.generate(mode="weighted")
    iv3 = m.intervals()
    rows = [(int(c), 0, int(a), int(b) - 1) for c, a, b in iv3]
    for c, L in enumerate(m.lengths):
        rows += [(c, 1, 0, L - 1), (c, 0, L // 3, 2 * L // 3), (c, 1, L // 3 + 5, L // 3 + 400), (c, 0, L // 2, L // 2)]  # whole contigs, nested, one position
    rng = np.random.default_rng(5)
    for _ in range(60):
        c = int(rng.integers(0, len(m.lengths)))
        a = int(rng.integers(0, m.lengths[c] - 10))
        rows.append((c, int(rng.integers(0, 2)), a, a + int(rng.integers(1, 3000))))
    iv = np.zeros(len(rows), api.INTERVAL_DTYPE)
    for k, (c, st, a, b) in enumerate(rows):
        iv[k] = (c, st, a, b)
    iv = iv[rng.permutation(len(iv))]
    single = run_gpu(m.lengths, reads, "weighted", intervals=iv)
    want, want_order = single["transposons"], single["transposon_order"]
    assert np.count_nonzero(want["histogram"]) > 100
    parts = _sharded(m.lengths, reads, "weighted", n_shards, known_rank=True, intervals=iv)
    for p in parts:
        got, order = p[3]
        assert np.array_equal(order, want_order)
        assert np.array_equal(bits(got["histogram"]), bits(want["histogram"]))
        assert np.array_equal(bits(got["reads_plus"]), bits(want["reads_plus"])) and np.array_equal(bits(got["reads_minus"]), bits(want["reads_minus"]))
        assert np.array_equal(got["p"].view(np.uint64), want["p"].view(np.uint64)) and np.array_equal(bits(got["q_value"]), bits(want["q_value"]))


def test_builtin_nccl_collective_single_rank():
    """ppp_nccl_attach: the library's own NCCL binding behind the all-reduce hook (a one-rank communicator here; the
    multi-rank use is bench.py --gpus N).  The three reductions then run through ncclAllReduce on the context's
    stream and must leave every result unchanged."""
    m = host.SynthModel("tiny", n_reads=120_000, seed=5)
    reads = m.generate(mode="weighted")
    single = run_gpu(m.lengths, reads, "weighted")
    with api.Context(m.lengths, mode="weighted") as ctx:
        ctx.attach_nccl(api.nccl_unique_id(), 0, 1)
        with pytest.raises(api.PppError):
            ctx.attach_nccl(api.nccl_unique_id(), 0, 1)  # one communicator per context
        for c, r in enumerate(reads):
            if len(r):
                ctx.count_reads(c, r)
        ctx.finish()
        ctx.map_heights_to_scores(False)
        ctx.count_stacks_by_group(False)
        sc = ctx.calculate_fdrs(0)
        assert np.array_equal(sc.loci, single["score"].loci)
        assert np.array_equal(ctx.signatures(), single["signatures"])
        assert np.array_equal(bits(sc.fdr_table), bits(single["score"].fdr_table))


def _fold_f32(weights):
    """reads += w in single precision, in order (pingpongpro.cpp:487)."""
    h = np.float32(0.0)
    for w in weights:
        h = np.float32(h + w)
    return h


@pytest.mark.parametrize("batches", [1, 3, 40])  # 40: batches below 4096 records take the single-launch path
def test_tall_stacks_fold_in_file_order(batches):
    """-m weighted on tall stacks (the warp-cooperative fold and, from 8192 records on, the block-wide fold through
    composed per-binade addition maps): run lengths around the 32-record, 256-record, 8192-record and 16384-record
    boundaries, both strands interleaved at one position, multiplicities beyond the weight table and beyond 2^24,
    and sums continued across submissions.  Expected value: the literal single-precision loop."""
    rng = np.random.default_rng(7)
    nh_pool = np.array([1, 1, 1, 1, 2, 3, 5, 7, 10, 63, 64, 100, (1 << 24) - 1, (1 << 24) + 3, 1 << 29], dtype=np.uint32)
    lengths_of_runs = [31, 32, 33, 34, 255, 256, 257, 511, 512, 513, 777, 5000, 40000, 8191, 8192, 8193, 16389, 120001, 700003]
    common = np.array([1, 1, 1, 2, 3, 4], dtype=np.uint32)  # the last two runs: heights that climb through many binades
    key, minus, a10, nh = [], [], [], []
    for i, n in enumerate(lengths_of_runs):
        pos = 100 + 50 * i
        key.append(np.full(n, pos, np.uint32))
        both = i % 3 != 0                      # every third position holds one strand only
        minus.append(rng.random(n) < 0.5 if both else np.full(n, i % 2 == 0))
        a10.append(rng.random(n) < 0.01)
        nh.append(rng.choice(common if n > 100000 else nh_pool, size=n))
    key, minus, a10, nh = map(np.concatenate, (key, minus, a10, nh))
    order = rng.permutation(len(key))         # positions interleaved in the "file"; the order within a position is what matters
    key, minus, a10, nh = key[order], minus[order], a10[order], nh[order]
    reads = host.pack_reads(key, minus, a10, nh)
    w = (1.0 / nh.astype(np.float64)).astype(np.float32)   # :449
    with api.Context([2000], mode="weighted") as ctx:
        cuts = [len(reads) * k // batches for k in range(batches + 1)]
        for k in range(batches):
            ctx.count_reads(0, reads[cuts[k]:cuts[k + 1]])
        ctx.finish()
        for strand in (0, 1):
            h, a = ctx.stacks(0, strand, 0, 2000)
            for i, n in enumerate(lengths_of_runs):
                pos = 100 + 50 * i
                sel = (key == pos) & (minus == bool(strand))
                want = _fold_f32(w[sel])
                assert bits(np.array([h[pos]]))[0] == bits(np.array([want]))[0], (strand, n, h[pos], want)
                assert bool(a[pos]) == bool(a10[sel].any())
            assert np.count_nonzero(h) == sum(1 for i in range(len(lengths_of_runs)) if ((key == 100 + 50 * i) & (minus == bool(strand))).any())


@pytest.mark.parametrize("mode", ["weighted", "unique"])
def test_reset_leaves_no_trace_of_the_previous_run(mode):
    """ppp_reset of a sparse store clears only the sectors whose occupancy bits are set (compact.cu: k_sparse_clear); a
    reset in the middle of an upload falls back to the memset.  A context that is reused for another read set must give
    what a fresh context gives: stacks, signatures, loci."""
    a = host.SynthModel("tiny", n_reads=120_000, seed=21)
    b = host.SynthModel("tiny", n_reads=90_000, seed=22)
    ra, rb = a.generate(mode=mode), b.generate(mode=mode)
    assert a.lengths == b.lengths
    fresh = {id(r): run_gpu(a.lengths, r, mode) for r in (ra, rb)}
    fresh_stacks = {id(r): gpu_stacks(a.lengths, r, mode) for r in (ra, rb)}

    def run(ctx, reads):
        for c, r in enumerate(reads):
            if len(r):
                ctx.count_reads(c, r)
        ctx.finish()
        rows = []
        for strand in (0, 1):
            for c, L in enumerate(a.lengths):
                h, a10 = ctx.stacks(c, strand, 0, L + 64)
                nz = np.flatnonzero(h)
                rows.append((np.full(len(nz), strand, np.uint32), np.full(len(nz), c, np.uint32), nz.astype(np.uint32), h[nz], a10[nz].astype(np.uint32)))
        stacks = [np.concatenate([r[i] for r in rows]) for i in range(5)]
        ctx.map_heights_to_scores(False)
        ctx.count_stacks_by_group(False)
        return stacks, ctx.calculate_fdrs(0).loci.copy(), ctx.signatures().copy()

    with api.Context(a.lengths, mode=mode) as ctx:
        for reads in (ra, rb, ra):
            stacks, loci, sigs = run(ctx, reads)
            for got, want in zip(stacks, fresh_stacks[id(reads)]):
                assert np.array_equal(bits(got) if got.dtype == np.float32 else got, bits(want) if want.dtype == np.float32 else want)
            assert np.array_equal(loci, fresh[id(reads)]["score"].loci) and np.array_equal(sigs, fresh[id(reads)]["signatures"])
            ctx.reset()                         # finished store: sparse clear
        for c, r in enumerate(rb):              # half an upload, then a reset (dense memset), then another run
            if len(r):
                ctx.count_reads(c, r[: len(r) // 2])
        ctx.reset()
        stacks, loci, sigs = run(ctx, ra)
        assert np.array_equal(loci, fresh[id(ra)]["score"].loci) and np.array_equal(sigs, fresh[id(ra)]["signatures"])
        for got, want in zip(stacks, fresh_stacks[id(ra)]):
            assert np.array_equal(bits(got) if got.dtype == np.float32 else got, bits(want) if want.dtype == np.float32 else want)
