#!/bin/bash
# round 2, second session: the single-GPU bench lines kept under profiles/ (c5 = the driver's default command, with its
# cpu_baseline leg; the others without it: the CPU port did not change, profiles/r2_bench_*_1gpu.json have those legs)
O=gpurun_out/r2b/final; mkdir -p $O
python bench.py > $O/r2b_bench_c5_1gpu.json 2> $O/r2b_bench_c5_1gpu.err; echo "c5 rc=$?"
for wl in c4 c3 c2 c1; do
  python bench.py --workload $wl --no-cpu > $O/r2b_bench_${wl}_1gpu.json 2> $O/r2b_bench_${wl}_1gpu.err; echo "$wl rc=$?"
done
python bench.py --workload c5 --move twalk --steps 50 --no-cpu > $O/r2b_bench_c5_twalk_1gpu.json 2>/dev/null; echo "twalk rc=$?"
python bench.py --workload c5 --move hmc --steps 10 --no-cpu > $O/r2b_bench_c5_hmc_1gpu.json 2>/dev/null; echo "hmc rc=$?"
for wl in c5 c4 c3; do
  python bench.py --workload $wl --dtype f32 --no-cpu --no-e2e > $O/r2b_bench_${wl}_f32_1gpu.json 2>/dev/null; echo "$wl f32 rc=$?"
done
python bench.py --workload c5 --n-walkers 524288 --no-cpu --no-e2e > $O/r2b_bench_c5_shard_size_1gpu.json 2>/dev/null
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r2b_bench_*1gpu.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "swap", (d.get("swap_stage") or {}).get("ms_per_call"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "no line", ex)
PY
