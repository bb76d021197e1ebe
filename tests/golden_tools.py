"""Loading of the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.mode = str(z["mode"])
        self.min_ov, self.max_ov, self.pp_ov = int(z["min_ov"]), int(z["max_ov"]), int(z["pp_ov"])
        self.W = self.max_ov - self.min_ov + 1
        self.contig_names = [str(x) for x in z["contig_names"]]
        self.contig_lengths = [int(x) for x in z["contig_lengths"]]
        bounds = np.concatenate([[0], np.cumsum(z["reads_per_contig"])]).astype(np.int64)
        reads = z["reads"]
        self.reads = [reads[bounds[i]:bounds[i + 1]] for i in range(len(self.contig_lengths))]
        self.total = float(z["total"])
        self.stacks = z["stacks"]
        self.freq = z["freq"]
        self.sigs = z["sigs"]
        self.grouped_post = z["grouped_post"]
        gp = np.zeros(self.W * 1000 * 4, np.float32)
        gp[z["grouped_pre_idx"]] = z["grouped_pre_val"]
        self.grouped_pre = gp.reshape(self.W, 1000, 2, 2)
        self.transposons = z["transposons"] if "transposons" in z.files else None

    def dense_freq(self) -> np.ndarray:
        out = np.zeros(int(self.freq["height"].max()) + 1, np.float32)
        out[self.freq["height"]] = self.freq["freq"]
        return out
