#!/bin/bash
# ncu launch list (gpu__time_duration per launch, no clock control) of a short bench run.  Usage: gpu_launches.sh TAG [bench args...]
tag=$1; shift
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $out/plain_$tag.json 2> $out/plain_$tag.err || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $out/ncu_launches_$tag.log 2>&1
echo "rc=$?"; tail -2 $out/plain_$tag.json | cut -c1-300; wc -l $out/launches_$tag.csv
