#!/bin/bash
# Strong-scaling run of bench.py on N GPUs of one box (torchrun, one rank per GPU), plus the multi-rank NCCL parity test.
# Usage (gpurun --gpus N): bash pingpongpro_b200/tools/gpu_scale.sh N [extra bench args]
n=$1; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $out/gpus_n$n.txt
timeout 900 python -m pytest tests/test_zz_nccl_ranks.py -m gpu -x -q -s > $out/nccl_ranks_n$n.log 2>&1; echo "nccl test rc=$?"; grep -E "identical|passed|failed|skipped" $out/nccl_ranks_n$n.log | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n "$@" > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_n$n.json"))
print("N=%d ms/step %.4f value %.3g e2e ms %.2f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["roofline"]["device_ms_per_step"], "heaviest rank reads", d["config"]["reads_on_heaviest_rank"], "verified", d["config"]["verified"])
PY
grep "ppp_trace" $out/bench_n$n.err | tail -12
