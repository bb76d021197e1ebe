#!/bin/bash
# ncu --set full captures of the stack builder's kernels after the round-2 rewrite (scatter through shared memory, block-wide
# fold of huge runs, sparse reset), configs[1].  Run only after the plain command has exited 0.
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --genome dm6 --reads 50000000"
timeout 600 $B > $out/prof_plain_dm6.json 2> $out/prof_plain_dm6.err || { echo plain failed; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_radix_scatter|k_fold_huge|k_sparse_clear|k_compact_write|k_unpack_reads' --launch-count 30 -f -o $out/r2_build2_dm6 $B > $out/ncu_r2_build2_dm6.log 2>&1; echo "build2 dm6 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final_dm6.csv $B > $out/ncu_launches_final_dm6.log 2>&1; echo "launch list rc=$?"
ls -la $out/r2_build2_dm6.ncu-rep; wc -l $out/launches_final_dm6.csv
