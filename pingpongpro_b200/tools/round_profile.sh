#!/bin/bash
# One consolidated measurement pass on a B200 box (run under gpurun from the repo root):
#   bench lines (N=1, both arms), ncu launch list of bench.py, ncu --set full of k_scan, big-genome bench lines,
#   command-line end-to-end timing against the reference binary.  Everything lands in gpurun_out/.
set -u
export PYTHONPATH=$PWD
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
nproc > $O/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> $O/nproc.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --steps 20 --warmup 5 --mode unique --no-cpu-baseline > $O/bench_n1_unique.json 2> $O/bench_n1_unique.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python -m pingpongpro_b200.tools.profile_pass > $O/plain_pass.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 2 -c 1 -o $O/k_scan_full python -m pingpongpro_b200.tools.profile_pass > $O/ncu_full.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --genome human --reads 500000000 > $O/bench_human.json 2> $O/bench_human.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --genome mouse --reads 200000000 > $O/bench_mouse.json 2> $O/bench_mouse.err
# command line end to end (decode included) against the reference binary
python -m pingpongpro_b200.tools.cli_e2e --reads 50000000 > $O/cli_e2e_50M_bam.json 2> $O/cli_e2e_50M_bam.err
python -m pingpongpro_b200.tools.cli_e2e --reads 1000000 --format sam > $O/cli_e2e_1M_sam.json 2> $O/cli_e2e_1M_sam.err
python -m pingpongpro_b200.tools.cli_e2e --reads 20000000 --format sam > $O/cli_e2e_20M_sam.json 2> $O/cli_e2e_20M_sam.err
# where the end-to-end step goes (host enqueue / device completion / scan + score)
python -m pingpongpro_b200.tools.e2e_probe > $O/e2e_probe.log 2>&1
