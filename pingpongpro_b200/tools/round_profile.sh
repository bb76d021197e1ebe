#!/bin/bash
# One consolidated measurement pass on a B200 box (run under gpurun from the repo root):
#   bench lines (N=1, both arms), ncu launch list of bench.py, ncu --set full of k_scan, big-genome bench lines,
#   command-line end-to-end timing against the reference binary.  Everything lands in gpurun_out/.
set -u
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
# command line end to end (decode included), same 5 M-read BAM for both binaries
T=$(mktemp -d)
pingpongpro_b200/bin/ppp_synth --genome dm6 --reads 5000000 --seed 1 --out $T/r.bam
pingpongpro_b200/bin/ppp_synth --genome dm6 --reads 1000000 --seed 1 --out $T/r1m.sam
s=$(date +%s.%N); pingpongpro_b200/bin/pingpongpro -i $T/r.bam -o $T/mine > /dev/null 2>&1; e=$(date +%s.%N); echo "b200 5M bam $(echo "$e - $s" | bc)" > $O/cli_e2e.txt
s=$(date +%s.%N); oracle/_ref/pingpongpro_ref -i $T/r.bam -o $T/ref > /dev/null 2>&1; e=$(date +%s.%N); echo "reference 5M bam $(echo "$e - $s" | bc)" >> $O/cli_e2e.txt
s=$(date +%s.%N); pingpongpro_b200/bin/pingpongpro -i $T/r1m.sam -o $T/mine1 > /dev/null 2>&1; e=$(date +%s.%N); echo "b200 1M sam $(echo "$e - $s" | bc)" >> $O/cli_e2e.txt
s=$(date +%s.%N); oracle/_ref/pingpongpro_ref -i $T/r1m.sam -o $T/ref1 > /dev/null 2>&1; e=$(date +%s.%N); echo "reference 1M sam $(echo "$e - $s" | bc)" >> $O/cli_e2e.txt
cmp $T/mine/ping-pong_signatures.tsv $T/ref/ping-pong_signatures.tsv && echo "5M outputs identical" >> $O/cli_e2e.txt
cmp $T/mine1/ping-pong_signatures.tsv $T/ref1/ping-pong_signatures.tsv && echo "1M outputs identical" >> $O/cli_e2e.txt
wc -l $T/mine/ping-pong_signatures.tsv >> $O/cli_e2e.txt
rm -rf $T
