#!/bin/bash
# round 2, second session: ncu captures of the kernels as shipped (each after the plain run of the same command exited 0).
#   gpurun -- 'bash profiles/r2b_capture.sh'; then, here:  ROUND=r2b python profiles/summarize_ncu.py full|list ...
O=gpurun_out/r2b/cap; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
B4="python bench.py --workload c4 --no-cpu --no-e2e --steps 22 --warmup 3"
B5="python bench.py --workload c5 --no-cpu --no-e2e --steps 12 --warmup 3"
$B4 > $O/plain_c4.json 2> $O/plain_c4.err || { echo "plain c4 failed"; exit 1; }
$B5 > $O/plain_c5.json 2> $O/plain_c5.err || { echo "plain c5 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c4.csv $B4 > $O/ncu_list_c4.log 2>&1
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_c5_step -f $B5 > $O/ncu_c5.log 2>&1; tail -1 $O/ncu_c5.log
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_c4_step -f $B4 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
$NCU --kernel-name-base demangled -k 'regex:step_reg_kernel<.*true>' -s 1 -c 1 -o $O/prof_c4_step_remap -f $B4 > $O/ncu_c4_remap.log 2>&1; tail -1 $O/ncu_c4_remap.log
$NCU -k regex:swap_paths_kernel -s 1 -c 1 -o $O/prof_c4_swap_paths -f $B4 > $O/ncu_paths.log 2>&1; tail -1 $O/ncu_paths.log
$NCU -k regex:swap_walk_kernel -s 1 -c 1 -o $O/prof_c4_swap_walk -f $B4 > $O/ncu_walk.log 2>&1; tail -1 $O/ncu_walk.log
