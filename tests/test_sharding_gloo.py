"""Multi-rank host logic on CPU (gloo, world_size 2): position-range ownership, halo routing of minus-strand
reads and the table all-reduce that the CUDA layer requests through its collective hook.
The per-shard height histograms (halo copies excluded, as k_scan does) must all-reduce to the single-process
oracle's histogram."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle_tools import Oracle
from pingpongpro_b200 import dist as pdist, host

MAX_OV = 23


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_histogram(lengths, reads, lo, hi, size):
    """Stack build + height histogram of one shard with numpy (-m unique: heights are counts)."""
    cum = np.concatenate([[0], np.cumsum(lengths)])
    hist = np.zeros(size, np.int64)
    owned = 0
    for c, r in enumerate(reads):
        if not len(r):
            continue
        g = cum[c] + r["key"].astype(np.int64)
        minus = (r["meta"] & 1).astype(bool)
        for strand in (False, True):
            sel = minus == strand
            pos, counts = np.unique(g[sel], return_counts=True)
            mine = pos < hi  # minus stacks at or beyond `hi` are halo copies owned by the right neighbour
            # keys beyond a contig's length stay with the shard that holds the contig's last base
            beyond = pos >= cum[c + 1]
            mine |= beyond & (cum[c + 1] <= hi) & (cum[c + 1] > lo)
            owned += int(counts[mine].sum())
            np.add.at(hist, np.minimum(counts[mine], size - 1), 1)
    return hist, owned


def _worker(rank, world, port, lengths, n_reads, seed, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = sum(lengths)
        cuts = pdist.shard_ranges(total, world)
        m = host.SynthModel(None, n_reads=n_reads, seed=seed, lengths=lengths)
        g_hi = cuts[rank + 1] if rank + 1 < world else total + 64
        own = m.generate(mode="unique", g_lo=cuts[rank], g_hi=g_hi, threads=2)
        with_halo = m.generate(mode="unique", g_lo=cuts[rank], g_hi=g_hi, halo=MAX_OV if rank + 1 < world else 0, threads=2)
        for a, b in zip(own, with_halo):  # the halo only ever adds minus-strand reads within max_overlap of the cut
            extra = len(b) - len(a)
            assert extra >= 0
        n_halo = sum(len(b) - len(a) for a, b in zip(own, with_halo))
        hist, owned = _shard_histogram(lengths, with_halo, cuts[rank], g_hi, 4096)
        counts = np.array([owned, n_halo], np.int64)
        hist = pdist.reduce_tables_cpu(hist, "sum")
        counts = pdist.reduce_tables_cpu(counts, "sum")
        mx = pdist.reduce_tables_cpu(np.array([np.flatnonzero(hist).max()], np.int64), "max")
        if rank == 0:
            np.save(out, np.concatenate([counts, mx, hist]))
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    assert pdist.shard_ranges(10, 3) == [0, 3, 6, 10]
    assert pdist.shard_ranges(137547960, 8)[-1] == 137547960


@pytest.mark.timeout(300)
def test_two_rank_histogram_allreduce_matches_oracle(tmp_path):
    lengths, n_reads, seed = [60000, 35011, 9000], 60000, 21
    out = str(tmp_path / "hist.npy")
    mp.spawn(_worker, args=(2, _free_port(), lengths, n_reads, seed, out), nprocs=2, join=True)
    got = np.load(out)
    owned, n_halo, max_h, hist = int(got[0]), int(got[1]), int(got[2]), got[3:]
    m = host.SynthModel(None, n_reads=n_reads, seed=seed, lengths=lengths)
    reads = m.generate(mode="unique")
    assert owned == sum(len(r) for r in reads)  # every read is owned by exactly one rank
    assert n_halo > 0
    freq = Oracle(reads, mode="unique").freq()
    ref = np.zeros(4096, np.int64)
    ref[: len(freq)] = freq.astype(np.int64)
    assert max_h == len(freq) - 1
    assert np.array_equal(hist, ref)
