"""python -m pingpongpro_b200.tools.e2e_probe: where the end-to-end step goes: host enqueue, device completion of the stack build, the scan/score pass."""
import time
import numpy as np
from pingpongpro_b200 import api, host
m = host.SynthModel("dm6", n_reads=50_000_000, seed=20140401)
pinned = []
def alloc(k):
    b = api.PinnedBuffer(k); pinned.append(b); return b.array
reads = m.generate(mode="weighted", alloc=alloc)
with api.Context(m.lengths, mode="weighted") as ctx:
    for it in range(8):
        t0 = time.perf_counter()
        ctx.reset()
        ts = []
        for c, r in enumerate(reads):
            ctx.count_reads(c, r); ts.append(time.perf_counter())
        t1 = time.perf_counter()
        ctx.finish()
        t2 = time.perf_counter()
        ctx.map_heights_to_scores(False); ctx.count_stacks_by_group(False); res = ctx.calculate_fdrs(0, False, True, False)
        t3 = time.perf_counter()
        print("enqueue %.2f ms (per contig: %s)  finish +%.2f  pass +%.2f  total %.2f" % (1e3*(t1-t0), " ".join("%.2f" % (1e3*(x-t0)) for x in ts), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t3-t0)))
