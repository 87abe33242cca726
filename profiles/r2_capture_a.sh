#!/bin/bash
# round 2, capture A (1 GPU): ncu --set full of the shipped kernels that had no or a stale summary in round 1:
# K1 at c5 (640 threads / 96 registers), the c4 step kernel, the two ladder-swap kernels at c4, the c3 step kernel.
# Each capture follows a plain run of the same command that exited 0.
set -x
O=gpurun_out/r2
mkdir -p $O
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e"
for wl in c5 c4 c3; do
  $B --workload $wl > $O/plain_$wl.json 2> $O/plain_$wl.err || exit 1
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_r2_c5_step -f $B --workload c5 > $O/ncu_c5.log 2>&1
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_r2_c4_step -f $B --workload c4 > $O/ncu_c4.log 2>&1
$NCU -k regex:swap_chain_kernel -s 1 -c 1 -o $O/prof_r2_c4_swap_chain -f $B --workload c4 > $O/ncu_c4_chain.log 2>&1
$NCU -k regex:permute_rows_kernel -s 2 -c 2 -o $O/prof_r2_c4_permute -f $B --workload c4 > $O/ncu_c4_perm.log 2>&1
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_r2_c3_step -f $B --workload c3 > $O/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r2_c4.csv $B --workload c4 --steps 12 > $O/ncu_list_c4.log 2>&1
python bench.py --workload c4 > $O/bench_c4_base.json 2> $O/bench_c4_base.err
python bench.py --workload c3 --no-cpu > $O/bench_c3_base.json 2> $O/bench_c3_base.err
ls -la $O
