#!/usr/bin/env python
"""bench.py -- the reference's headline metric on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--genome G]

Metric: overlap-scan genome positions / second (BASELINE.json `metric`):
  value  G / (one pass of mapHeightsToScores + countStacksByGroup + collapseBins + calculateFDRs through the
         C ABI), stacks resident in HBM when the timed region starts, the scored loci resident in HBM (and
         their count on the host) when it ends;
  e2e    the same metric through the C ABI from HOST buffers: packed reads in pinned memory -> H2D ->
         stack build -> the pass above -> loci in host memory (reads/s of the same region reported beside it).
A step is one pass over the whole synthetic data set.  N=1 workload: BASELINE.json configs[1] (50 M reads,
dm6-sized genome, default parameters).  N>1: weak scaling -- every rank holds one dm6-sized position range of
an N-times larger genome (range sharding of SURVEY §8(e)); the data path has no collective, the histogram
tables are all-reduced over NCCL.  --scaling strong cuts ONE genome into N position ranges instead (the
human-sized configs[3] study: --genome human --reads 500000000).

--impl reference times the UNMODIFIED reference (oracle/_ref/ppp_harness, its own functions on decoded records
held in memory) on one host core, on a bounded genome-range sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "overlap-scan genome positions/sec"
UNIT = "positions/s"
SEED = 2  # seed = config number


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(args):
    """DRAM bytes per k_scan launch from the committed ncu --set full capture of this workload (profiles/), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_k_scan_traffic.json")) as f:
            t = json.load(f)
        if (t["genome"], t["reads"], t["mode"]) == (args.genome, args.reads, args.mode):
            return t["traffic_bytes_per_launch"]
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def reference_sample(genome_size, n_reads):
    """Bounded sample for the CPU arms: a genome range that holds about 2 M reads of the workload."""
    frac = min(1.0, 2.0e6 / max(1, n_reads))
    return 0, max(1, int(genome_size * frac))


def run_harness(args, steps, warmup, genome_size):
    harness = os.path.join(ROOT, "oracle", "_ref", "ppp_harness")
    if not os.path.exists(harness):
        return None
    lo, hi = reference_sample(genome_size, args.reads)
    cmd = [harness, "time", "--genome", args.genome, "--reads", str(args.reads), "--seed", str(SEED), "--gpos", f"{lo}:{hi}", "-m", args.mode,
           "--steps", str(steps), "--warmup", str(warmup)]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = [json.loads(line) for line in out.splitlines() if line.startswith("{")]
    timed = [r for r in rows if r["timed"]]
    return dict(rows=timed, positions=hi - lo, sample=f"genome range [{lo},{hi}) of the {args.reads}-read {args.genome} workload: "
                f"{timed[0]['records']} decoded records held in memory, {timed[0]['pairs']} signatures")


def reference_arm(args, rank, world, json_out):
    if rank != 0:
        return
    from pingpongpro_b200 import host
    m = host.SynthModel(args.genome, n_reads=args.reads, seed=SEED)
    h = run_harness(args, args.steps, args.warmup, m.genome_size)
    if h is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ppp_harness was not built (needs /root/reference at build time)"}), file=json_out, flush=True)
        return
    full = [r["t_count"] + r["t_heights"] + r["t_group"] + r["t_collapse"] + r["t_fdr"] for r in h["rows"]]
    scan = [r["t_heights"] + r["t_group"] + r["t_collapse"] + r["t_fdr"] for r in h["rows"]]
    t = sum(full) / len(full)
    v = h["positions"] / t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[1]: {args.reads} synthetic reads on the {args.genome}-sized genome, -m {args.mode}, default parameters; "
                               "reference timed on a bounded genome-range sample", "sample": h["sample"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": h["sample"],
                         "scan_and_scoring_only": h["positions"] / (sum(scan) / len(scan)),
                         "stage_seconds": {k: sum(r[k] for r in h["rows"]) / len(h["rows"]) for k in ("t_count", "t_heights", "t_group", "t_collapse", "t_fdr")},
                         "reads_per_s": h["rows"][0]["records"] / t},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=json_out, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads per GPU-sized genome (configs[1]: 50 M)")
    ap.add_argument("--genome", default="dm6")
    ap.add_argument("--mode", default="weighted")
    ap.add_argument("--min-stack-height", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one genome + read set per GPU (default); strong = one genome cut into N position ranges")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "torch"], help="N>1: the library's NCCL binding, or the torch.distributed hook")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world, json_out)
        return

    import numpy as np
    import torch

    from pingpongpro_b200 import api, dist as pdist, host

    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    tdist = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if tdist is not None:
            tdist.barrier()

    def max_over_ranks(x):
        if tdist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if tdist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload.  weak (default): rank r owns copy r of the genome (contigs [r*nc, (r+1)*nc)) with its own seeded
    # read set.  strong: ONE genome and read set, cut into `world` equal position ranges; every rank
    # generates the reads of its range plus the minus-strand halo.
    strong = args.scaling == "strong" and world > 1
    model = host.SynthModel(args.genome, n_reads=args.reads, seed=SEED + (0 if strong else rank))
    nc = len(model.lengths)
    G = model.genome_size
    if strong:
        lengths = list(model.lengths)
        cuts = pdist.shard_ranges(G, world)
        r_lo, r_hi = cuts[rank], cuts[rank + 1]
        contig0 = 0
        work_positions = G           # whole-job positions per step
    else:
        lengths = model.lengths * world
        r_lo, r_hi = rank * G, (rank + 1) * G
        contig0 = rank * nc
        work_positions = world * G
    t_gen = time.perf_counter()
    pinned = []

    def alloc(k):
        b = api.PinnedBuffer(k)
        pinned.append(b)
        return b.array

    if strong:  # (+64: keys at or just past a contig's end belong to the shard holding its last base)
        reads = model.generate(mode=args.mode, alloc=alloc, g_lo=r_lo, g_hi=r_hi, halo=23 + 64)
    else:
        reads = model.generate(mode=args.mode, alloc=alloc)
    t_gen = time.perf_counter() - t_gen
    n_reads = sum(len(r) for r in reads)

    ctx = api.Context(lengths, mode=args.mode, device=local_rank, range_begin=r_lo, range_end=r_hi)
    if world > 1:
        if args.collective == "torch":
            ctx.set_allreduce(pdist.torch_allreduce_hook())
        else:
            pdist.attach_nccl(ctx)
    words, owned = ctx.layout()

    # contigs are independent, so the caller is free to choose their order: largest first, so that the copy -> sort ->
    # fold pipeline drains on the small ones
    submit_order = sorted(range(len(pinned)), key=lambda c: -pinned[c].n)

    def upload():
        ctx.reset()
        for c in submit_order:
            if pinned[c].n:
                ctx.count_reads(contig0 + c, pinned[c])
        ctx.finish()

    max_height = [0]

    def scan_and_score(loci_to_host):
        # the three seams of the reference's main(), asking for nothing before the last one: one stream of kernels, one
        # host synchronisation (ppp_height_hist)
        ctx.map_heights_to_scores(want_table=False, want_max_height=False)
        ctx.count_stacks_by_group(want_table=False, want_count=False)
        # loci_to_host: the compacted loci are also copied into the library's pinned host buffer (20 B per locus)
        return ctx.calculate_fdrs(args.min_stack_height, want_tables=False, want_loci=loci_to_host, copy_loci=False)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_clock0 = time.perf_counter()

    # ---- value: stacks resident in HBM (1.1 GB per strand pair >> 126 MB L2: no L2 flush needed); the step ends
    # with the compacted, scored loci resident in HBM and their count on the host (the device->host copy of the
    # loci themselves belongs to the end-to-end figure below)
    upload()
    for _ in range(args.warmup):
        res = scan_and_score(False)
    ctx.timings_reset()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = scan_and_score(False)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    tm = ctx.timings()  # device time of the LAST step's kernels (CUDA events on the context's stream)
    launches_per_step = tm["launches"] / args.steps
    info = ctx.pass_info()
    max_height[0] = info["max_height"]
    n_pairs, n_loci = info["n_pairs"], res.n_loci
    value = work_positions / (dt / args.steps)

    # ---- e2e: host buffers -> H2D -> stack build -> scan + score -> loci on the host
    for _ in range(args.warmup):
        upload()
        scan_and_score(True)
    ctx.timings_reset()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload()
        res = scan_and_score(True)
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    tm_e2e = ctx.timings()
    t_clock1 = time.perf_counter()
    e2e_value = work_positions / (dt_e2e / args.steps)
    total_reads = sum_over_ranks(float(n_reads))
    total_pairs = sum_over_ranks(float(n_pairs))
    total_loci = sum_over_ranks(float(n_loci))

    if rank != 0:
        if tdist is not None:
            tdist.destroy_process_group()
        return

    clocks = sampler.stop(t_clock0, t_clock1)
    peak, peak_src = measured_peak_hbm()
    # dominant kernel: the fused streaming pass k_scan.  Algorithmic bytes per launch: every word of
    # both strand arrays once (8 B per padded position) + 16 B per signature record written.
    scan_bytes = 8.0 * words + 16.0 * n_pairs
    achieved = scan_bytes / (tm["scan_ms"] * 1e-3) / 1e9
    survey_bytes = 16.0 * words + 16.0 * n_pairs  # SURVEY §8(d) counts K2 and K3 as two passes of 8 B/position
    which = {("dm6", 50_000_000): "configs[1]", ("mouse", 200_000_000): "configs[2]", ("human", 500_000_000): "configs[3]"}.get((args.genome, args.reads), "(custom size)")
    per = "in total" if strong else "per GPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json {which}: {args.reads} synthetic piRNA-seq reads {per} on a {args.genome}-sized genome ({G} positions {per}), "
                               f"-m {args.mode}, overlaps 3..23, default parameters", "positions_per_gpu": owned, "reads_per_gpu_after_filters": n_reads,
                   "signatures": total_pairs, "loci": total_loci, "max_stack_height": max_height[0],
                   "sharding": ("one genome cut into equal position ranges, one per rank" if strong else "one genome per rank (a position range of the concatenation)") if world > 1 else "none",
                   "l2": "inputs (8 B x positions = %.2f GB) are far larger than the 126 MB L2; no flush between steps" % (8e-9 * words),
                   "synthetic_generation_s": t_gen},
        "roofline": {"bound": "hbm", "kernel": "k_scan (fused height histogram + overlap scan)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(args), "peak_source": peak_src, "algorithmic_bytes_per_launch": scan_bytes,
                     "kernel_ms": tm["scan_ms"], "achieved_two_pass_accounting_gbs": survey_bytes / (tm["scan_ms"] * 1e-3) / 1e9,
                     "device_ms_per_step": {k: tm[k] for k in ("scan_ms", "lut_ms", "bin_ms", "score_ms")}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * total_reads), "d2h_bytes_per_step": int(20 * total_loci + 64 * world),
                "ms_per_step": 1e3 * dt_e2e / args.steps, "reads_per_s": total_reads / (dt_e2e / args.steps),
                "stack_build_device_ms_per_step": tm_e2e["stack_build_ms"] / args.steps},
        "gpu_launches": int(round(launches_per_step * args.steps)),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            h = run_harness(args, 1, 0, G)
            if h:
                r = h["rows"][0]
                t_scan = r["t_heights"] + r["t_group"] + r["t_collapse"] + r["t_fdr"]
                line["cpu_baseline"] = {"value": h["positions"] / t_scan, "unit": UNIT, "cores": 1, "kind": "reference", "sample": h["sample"],
                                        "e2e_value": h["positions"] / (t_scan + r["t_count"]), "reads_per_s": r["records"] / (t_scan + r["t_count"]),
                                        "stage_seconds": {k: r[k] for k in ("t_count", "t_heights", "t_group", "t_collapse", "t_fdr")}}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line), file=json_out, flush=True)
    if tdist is not None:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
