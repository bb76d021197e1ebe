#!/bin/bash
# N-GPU bench with a long timed region (host jitter averages out).  Usage: gpu_bench_steps.sh N STEPS
n=$1; k=$2
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n --steps $k --warmup 5 > $out/bench_n${n}_k$k.json 2> $out/bench_n${n}_k$k.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_n${n}_k$k.json"))
print("N=%d steps %d ms/step %.4f value %.3g e2e ms %.2f" % (d["n_gpus"], d["steps"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["roofline"]["device_ms_per_step"], "verified", d["config"]["verified"]["identical"])
PY
