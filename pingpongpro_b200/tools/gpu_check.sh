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
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -3 $out/smoke.log; tail -5 $out/pytest_*.log; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    for name, b in (("human", d), ("dm6", d.get("secondary"))):
        if not b: continue
        print(name, "ms/step", round(b["ms_per_step"], 4), "value %.3g" % b["value"], "e2e ms", round(b["e2e"]["ms_per_step"], 3), b["roofline"]["device_ms_per_step"],
              "frac", round(b["roofline"]["frac"], 3), "build ms", round(b["e2e"]["stack_build_device_ms_per_step"], 3), "verified", (b["config"]["verified"] or {}).get("identical"))
    print("cpu_baseline", {k: d.get("cpu_baseline", {}).get(k) for k in ("value", "e2e_value")}, "launches", d["gpu_launches"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench.err").read()[-2000:])
PY
