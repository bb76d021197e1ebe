#!/bin/bash
# Round-2 GPU check: parity tests, smoke, short bench runs.  Usage (on the GPU box): bash pingpongpro_b200/tools/gpu_check.sh [quick|full]
# Everything is wrapped in `timeout` so that a hung kernel cannot hold the box.
mode=${1:-quick}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.csv 2>&1
nproc > $out/nproc.txt; free -g >> $out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
if [ "$mode" = quick ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/pytest_parity.log
else
  timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --genome dm6 --reads 50000000 --no-cpu-baseline > $out/bench_dm6.json 2> $out/bench_dm6.err; echo "bench dm6 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --genome human --reads 500000000 --no-cpu-baseline > $out/bench_human.json 2> $out/bench_human.err; echo "bench human rc=$?"
tail -3 $out/smoke.log; tail -5 $out/pytest_*.log; python - <<'PY'
import json
for f in ("bench_dm6", "bench_human"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["roofline"]["device_ms_per_step"], "build ms", round(d["e2e"]["stack_build_device_ms_per_step"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
