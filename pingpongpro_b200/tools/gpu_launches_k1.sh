#!/bin/bash
# ncu launch list of the FIRST upload of a bench run (the stack builder's kernels, batch by batch).  Usage: gpu_launches_k1.sh TAG COUNT [bench args...]
tag=$1; cnt=$2; shift; shift
out=gpurun_out
mkdir -p $out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c $cnt --csv --log-file $out/launches_k1_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary "$@" > $out/ncu_launches_k1_$tag.log 2>&1
echo "rc=$?"; wc -l $out/launches_k1_$tag.csv
