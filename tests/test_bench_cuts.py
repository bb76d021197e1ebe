"""bench.py's cost-balanced position ranges (strong scaling): cover the genome, ascend, and put about the same number of
reads on every rank (the scan's work follows the stacks, not the positions: VERDICT r1)."""
import importlib.util
import os

import numpy as np

from pingpongpro_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_balanced_cuts_cover_the_genome_and_balance_the_reads():
    bench = _bench()
    m = host.SynthModel("dm6", n_reads=2_000_000, seed=2)
    for world in (1, 2, 4, 8):
        cuts = bench.balanced_cuts(m, world, "weighted")
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == m.genome_size
        assert all(b > a for a, b in zip(cuts, cuts[1:]))
        if world == 1:
            continue
        per_rank = [sum(len(r) for r in m.generate(mode="weighted", g_lo=cuts[k], g_hi=cuts[k + 1])) for k in range(world)]
        cost = [n + (cuts[k + 1] - cuts[k]) / 24 for k, n in enumerate(per_rank)]
        assert max(cost) / (sum(cost) / world) < 1.08, (world, per_rank)
        equal = [sum(len(r) for r in m.generate(mode="weighted", g_lo=m.genome_size * k // world, g_hi=m.genome_size * (k + 1) // world)) for k in range(world)]
        assert max(per_rank) <= max(equal) * 1.02  # never worse than equal positions
