#!/bin/bash
# ncu capture of one kernel of the bench step.  Usage: bash pingpongpro_b200/tools/gpu_profile.sh KERNEL_REGEX TAG [bench args...]
k=$1; tag=$2; shift 2
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 --launch-count 1 -f -o $out/$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > $out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 $out/ncu_$tag.log; ls -la $out/$tag.ncu-rep
