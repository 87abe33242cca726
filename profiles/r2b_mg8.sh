#!/bin/bash
# round 2, second session: the 8-GPU bench lines of c5 and c4 (each runs its quick parity subset before timing)
N=8; O=gpurun_out/r2b/final; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for wl in c5 c4; do
  timeout 300 $TR --master-port 2973${#wl} bench.py --gpus $N --workload $wl --no-cpu > $O/r2b_bench_${wl}_${N}gpu.json 2> $O/r2b_bench_${wl}_${N}gpu.err
  echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2b_bench_${wl}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$wl N=$N ms/step", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "parity", (d.get("parity_check") or {}), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "swap", (d.get("swap_stage") or {}).get("ms_per_call"))
    for k, v in (d.get("extra") or {}).items():
        print("   ", k, {a: b for a, b in v.items() if a not in ("note", "roofline")}, (v.get("roofline") or {}).get("frac"))
except Exception as ex:
    print("no line", ex)
PY
done
