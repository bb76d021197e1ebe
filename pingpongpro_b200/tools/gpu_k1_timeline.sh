#!/bin/bash
# Device timeline of the stack builder (copies, sorts, folds per batch) during the last upload of a short bench run.
# Usage: gpu_k1_timeline.sh TAG [bench args...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
PPP_K1_TIMELINE=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary "$@" > $out/k1_timeline_$tag.json 2> $out/k1_timeline_$tag.err; echo "rc=$?"
python - <<PY
lines = [l for l in open("$out/k1_timeline_$tag.err") if l.startswith("[ppp_k1]")]
# the last block: from the last line whose start is 0.000
starts = [i for i, l in enumerate(lines) if " 0.000 .." in l]
blk = lines[starts[-1]:] if starts else lines
open("$out/k1_timeline_$tag.txt", "w").writelines(blk)
print(len(blk), "spans"); print("".join(blk[:12])); print("..."); print("".join(blk[-8:]))
PY
python -c "
import json; d=json.load(open('$out/k1_timeline_$tag.json')); print('e2e ms', d['e2e']['ms_per_step'])"
