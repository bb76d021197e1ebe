"""python -m pingpongpro_b200.tools.fold_probe [N]: one long run (a single stack of N reads on one strand) through the weighted stack builder."""
import sys
import numpy as np
from pingpongpro_b200 import api, host
N = int(sys.argv[1]) if len(sys.argv) > 1 else 480_000
rng = np.random.default_rng(1)
nh = rng.choice(np.array([1, 2, 3, 5, 7, 10], dtype=np.uint32), size=N, p=[0.8, 0.08, 0.05, 0.03, 0.02, 0.02])
reads = host.pack_reads(np.full(N, 5000, np.uint32), np.zeros(N, bool), np.zeros(N, bool), nh)
with api.Context([100000], mode="weighted") as ctx:
    for _ in range(3):
        ctx.reset()
        ctx.count_reads(0, reads)
        ctx.finish()
    print(ctx.timings())
