#!/bin/bash
# round 2, capture B (1 GPU): ncu --set full of the kernels as SHIPPED at the end of the round -- the c3 and c4 step
# kernels (wide shape for 16-lane rows; loop-free flags), the ladder swap's three kernels, and the launch list of c4.
O=gpurun_out/r2/capb; mkdir -p $O
B="python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e"
NCU="ncu --set full --clock-control none --import-source on"
$B --workload c4 > $O/plain_c4.json 2> $O/plain_c4.err || exit 1
$B --workload c3 > $O/plain_c3.json 2> $O/plain_c3.err || exit 1
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_c4_step -f $B --workload c4 > $O/ncu_c4.log 2>&1
$NCU -k regex:step_reg_kernel -s 5 -c 1 -o $O/prof_c3_step -f $B --workload c3 > $O/ncu_c3.log 2>&1
$NCU -k regex:swap_chain_kernel -s 1 -c 1 -o $O/prof_c4_swap_chain -f $B --workload c4 > $O/ncu_chain.log 2>&1
$NCU -k regex:swap_uniforms_kernel -s 1 -c 1 -o $O/prof_c4_swap_uniforms -f $B --workload c4 > $O/ncu_unif.log 2>&1
$NCU -k regex:gather_rows_kernel -s 1 -c 1 -o $O/prof_c4_gather_rows -f $B --workload c4 > $O/ncu_gather.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $O/launches_c4.csv $B --workload c4 > $O/ncu_list_c4.log 2>&1
ls -la $O | head -30
