#!/bin/bash
# round 2, single-GPU bench lines kept under profiles/ and the ncu capture of the one-sided step kernel (single rank)
O=gpurun_out/r2/final; mkdir -p $O
for wl in c5 c4 c3 c2 c1; do
  python bench.py --workload $wl > $O/r2_bench_${wl}_1gpu.json 2> $O/r2_bench_${wl}_1gpu.err; echo "$wl rc=$?"
done
python bench.py --workload c5 --move twalk --steps 50 > $O/r2_bench_c5_twalk_1gpu.json 2>/dev/null; echo "twalk rc=$?"
python bench.py --workload c5 --move hmc --steps 10 > $O/r2_bench_c5_hmc_1gpu.json 2>/dev/null; echo "hmc rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python bench.py --workload c5 --n-walkers 524288 --no-cpu --no-e2e > $O/r2_bench_c5_shard_size_1gpu.json 2>/dev/null
python profiles/r2_peer_single.py > $O/peer_single.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_reg_kernel -s 4 -c 1 -o $O/prof_r2_c5_peer_step -f python profiles/r2_peer_single.py > $O/ncu_peer.log 2>&1; tail -2 $O/ncu_peer.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r2_bench_*1gpu.json")) + ["$O/r2_bench_reference_arm.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "swap", (d.get("swap_stage") or {}).get("ms_per_call"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "no line", ex)
PY
