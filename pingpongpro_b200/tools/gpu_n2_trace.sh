#!/bin/bash
# Two-GPU bench with the exchange / pass timers, then a short traced run (PPP_TRACE=1: where the host waits).
out=gpurun_out; mkdir -p $out
if [ "$1" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard or range" 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_n2.json"))
print("N=%d ms/step %.4f value %.3g e2e ms %.2f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["roofline"]["device_ms_per_step"], d["roofline"].get("kernel_ms_mean_over_ranks"), "verified", d["config"]["verified"]["identical"])
PY
PPP_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2_traced.json 2> $out/bench_n2_traced.err; echo "traced rc=$?"
grep "ppp_trace" $out/bench_n2_traced.err | grep -n "grow" | head -4
grep "ppp_trace" $out/bench_n2_traced.err | tail -12
