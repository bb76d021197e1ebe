#!/usr/bin/env python
"""Generates tests/golden/fullsize/*.json: SHA-256 digests of every intermediate of the UNMODIFIED reference
(oracle/_ref/ppp_harness digest: /root/reference/pingpongpro.cpp behind the include-harness) on BASELINE.json's
FULL-SIZE configurations.  Runs on the CPU container (needs /root/reference at build time, ~25 GB of RAM for the
human-sized case, 10-40 minutes per case on one core); the GPU tests (tests/test_gpu_fullsize.py) and bench.py hash the
CUDA path's results with tests/fullsize_digest.py / the same byte layouts and compare.

    python tests/golden/make_fullsize_digests.py [case ...] [--jobs N]

Cases (seed = config number):
    config2_weighted  dm6-sized, 50 M reads, seed 2, -m weighted          (BASELINE.json configs[1])
    config2_unique    same, -m unique
    config3_weighted  mouse-sized, 200 M reads, seed 3, -m weighted, synthetic annotation (configs[2])
    config4_weighted  human-sized, 500 M reads, seed 4, -m weighted       (configs[3]; the bench workload)
    config4_unique    same, -m unique
    config5_max26     dm6-sized 50 M reads, seed 2, -m unique, background overlaps 3..26   (configs[4])
    config5_max30     same, 3..30
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "fullsize")
REF = os.path.join(ROOT, "oracle", "_ref")
SYNTH = os.path.join(ROOT, "pingpongpro_b200", "bin", "ppp_synth")
MIN_HEIGHTS = "0,1,2,5,7,10,20"

CASES = {
    "config2_weighted": dict(genome="dm6", reads=50_000_000, seed=2, mode="weighted"),
    "config2_unique": dict(genome="dm6", reads=50_000_000, seed=2, mode="unique"),
    "config3_weighted": dict(genome="mouse", reads=200_000_000, seed=3, mode="weighted", annotation=True),
    "config4_weighted": dict(genome="human", reads=500_000_000, seed=4, mode="weighted"),
    "config4_unique": dict(genome="human", reads=500_000_000, seed=4, mode="unique"),
    "config5_max26": dict(genome="dm6", reads=50_000_000, seed=2, mode="unique", max_overlap=26),
    "config5_max30": dict(genome="dm6", reads=50_000_000, seed=2, mode="unique", max_overlap=30),
    # small cases: they pin the digest pipeline itself on the CPU (tests/test_fullsize_digest.py: C port vs these)
    "small_tiny_weighted": dict(genome="tiny", reads=60_000, seed=42, mode="weighted", annotation=True),
    "small_dm6_weighted": dict(genome="dm6", reads=400_000, seed=9, mode="weighted", annotation=True),
    "small_dm6_unique_max30": dict(genome="dm6", reads=400_000, seed=9, mode="unique", max_overlap=30),
}


def run_case(name: str) -> None:
    c = CASES[name]
    mo = c.get("max_overlap", 23)
    harness = os.path.join(REF, "ppp_harness" if mo == 23 else f"ppp_harness_max{mo}")
    cmd = [harness, "digest", "--genome", c["genome"], "--reads", str(c["reads"]), "--seed", str(c["seed"]), "-m", c["mode"], "-s", MIN_HEIGHTS]
    tmp = f"/tmp/ppp_fullsize_{name}"
    os.makedirs(tmp, exist_ok=True)
    if c.get("annotation"):
        annot = os.path.join(tmp, "annotation.tsv")
        subprocess.check_call([SYNTH, "--genome", c["genome"], "--reads", str(c["reads"]), "--seed", str(c["seed"]), "--annot", annot], stdout=subprocess.DEVNULL)
        cmd += ["-t", annot, "--pq", os.path.join(tmp, "pq.f32")]
    print("+", " ".join(cmd), flush=True)
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True).stdout
    d = json.loads(out)
    d["case"] = name
    d["command"] = " ".join(os.path.relpath(x, ROOT) if x.startswith(ROOT) else x for x in cmd)
    os.makedirs(OUT, exist_ok=True)
    if c.get("annotation"):
        pq = np.fromfile(os.path.join(tmp, "pq.f32"), dtype="<f4").reshape(-1, 2)
        np.save(os.path.join(OUT, name + "_pq.npy"), pq)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
        f.write("\n")
    print(f"[{name}] done: {d['n_stacks']} stacks, {d['n_signatures']} signatures, {d['seconds']}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    ap.add_argument("--jobs", type=int, default=2)
    a = ap.parse_args()
    with ThreadPoolExecutor(a.jobs) as ex:
        for r in ex.map(run_case, a.cases):
            pass


if __name__ == "__main__":
    sys.exit(main())
