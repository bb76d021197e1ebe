#!/bin/bash
# Round-2 ncu evidence: --set full captures of every kernel of the step (run only after the plain command has exited 0).
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
timeout 600 $B --genome dm6 --reads 50000000 > $out/prof_plain_dm6.json 2> $out/prof_plain_dm6.err || { echo plain failed; exit 1; }
# (a) the scan + score pass, configs[1]: third pass of the run
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan|k_bin|k_score|k_collapse|k_fdr_table' --launch-skip 16 --launch-count 8 -f -o $out/r2_pass_dm6 $B --genome dm6 --reads 50000000 > $out/ncu_r2_pass_dm6.log 2>&1; echo "pass dm6 rc=$?"
# (b) the stack builder (weighted: sort + fold; compaction), configs[1]: first upload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sort_keys|k_radix_hist|k_radix_scatter|k_fold_runs|k_fold_long|k_tile_popc|k_compact_write' --launch-count 16 -f -o $out/r2_build_dm6 $B --genome dm6 --reads 50000000 > $out/ncu_r2_build_dm6.log 2>&1; echo "build dm6 rc=$?"
# (c) the integer stack builder (-m unique)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stack_build_int' --launch-count 3 -f -o $out/r2_build_int_dm6 $B --genome dm6 --reads 50000000 --mode unique > $out/ncu_r2_build_int_dm6.log 2>&1; echo "build int rc=$?"
# (d) k_scan on the human-sized config
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan' --launch-skip 2 --launch-count 1 -f -o $out/r2_scan_human $B > $out/ncu_r2_scan_human.log 2>&1; echo "scan human rc=$?"
ls -la $out/r2_*.ncu-rep
