"""Short driver for ncu: builds the stacks of a synthetic data set once, then runs the scan + score pass a few
times.  python -m pingpongpro_b200.tools.profile_pass --genome dm6 --reads 50000000 --mode weighted --passes 4"""
import argparse

from pingpongpro_b200 import api, host

ap = argparse.ArgumentParser()
ap.add_argument("--genome", default="dm6")
ap.add_argument("--reads", type=int, default=50_000_000)
ap.add_argument("--mode", default="weighted")
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--uploads", type=int, default=1)
a = ap.parse_args()
m = host.SynthModel(a.genome, n_reads=a.reads, seed=2)
reads = m.generate(mode=a.mode)
with api.Context(m.lengths, mode=a.mode) as ctx:
    for _ in range(a.uploads):
        ctx.reset()
        for c, r in enumerate(reads):
            ctx.count_reads(c, r)
        ctx.finish()
    import time
    for _ in range(a.passes):
        t0 = time.perf_counter()
        ctx.map_heights_to_scores(False)
        t1 = time.perf_counter()
        ctx.count_stacks_by_group(False)
        t2 = time.perf_counter()
        res = ctx.calculate_fdrs(0, False, True, False)
        t3 = time.perf_counter()
        print("wall ms: height_hist %.3f scan %.3f score %.3f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
    print(ctx.timings(), ctx.n_pairs, res.n_loci)
