#!/bin/bash
# round 2 (second session), the multi-GPU evidence kept under profiles/: all parity cases (sharded == single-GPU chain) and the bench
# lines at N GPUs.   usage: bash profiles/r2_final_mg.sh N
N=${1:-2}
O=gpurun_out/r2b/final; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 tests/multi_gpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" > $O/r2b_multi_gpu_check_n$N.log
echo "multi_gpu_check rc=${PIPESTATUS[0]}"; tail -4 $O/r2b_multi_gpu_check_n$N.log
for wl in c5 c4 c3; do
  timeout 400 $TR --master-port 2972${#wl} bench.py --gpus $N --workload $wl > $O/r2b_bench_${wl}_${N}gpu.json 2> $O/r2b_bench_${wl}_${N}gpu.err
  echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2b_bench_${wl}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$wl N=$N ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "parity", (d.get("parity_check") or {}).get("identical"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "swap", (d.get("swap_stage") or {}).get("ms_per_call"))
    for k, v in (d.get("extra") or {}).items():
        print("   ", k, {a: b for a, b in v.items() if a not in ("note", "roofline")}, (v.get("roofline") or {}).get("frac"))
except Exception as ex:
    print("no line", ex)
PY
done
