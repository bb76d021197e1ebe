#!/bin/bash
# N-GPU bench for the record, then a short traced run (PPP_TRACE=1: host timestamps of every stage of every step).  Usage: gpu_n4_trace.sh N
n=${1:-4}
out=gpurun_out; mkdir -p $out
nproc > $out/nproc_n$n.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_n$n.json"))
print("N=%d ms/step %.4f value %.3g e2e ms %.2f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["roofline"]["device_ms_per_step"], d["roofline"].get("kernel_ms_mean_over_ranks"), "verified", d["config"]["verified"]["identical"])
PY
PPP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 4 --warmup 3 > $out/bench_n${n}_traced.json 2> $out/bench_n${n}_traced.err; echo "traced rc=$?"
grep -c ppp_trace $out/bench_n${n}_traced.err
