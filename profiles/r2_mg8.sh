#!/bin/bash
# round 2, 8-GPU probe: c5 with the k-schedule and c4 (rung shards, one-sided swap)
N=${1:-8}; TAG=${2:-a}; K=${3:-4}
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BENCH_DIAG=1 timeout 300 $TR --master-port 29631 bench.py --gpus $N --steps 100 --warmup 10 --global-every $K --no-e2e > $O/mg_c5_n${N}_k${K}_$TAG.json 2> $O/mg_c5_n${N}_k${K}_$TAG.err
echo "c5 rc=$?"; tail -c 400 $O/mg_c5_n${N}_k${K}_$TAG.err
timeout 300 $TR --master-port 29632 bench.py --gpus $N --steps 100 --warmup 10 --workload c4 --no-e2e > $O/mg_c4_n${N}_$TAG.json 2> $O/mg_c4_n${N}_$TAG.err
echo "c4 rc=$?"; tail -c 400 $O/mg_c4_n${N}_$TAG.err
python - <<PY
import json
for f in ("$O/mg_c5_n${N}_k${K}_$TAG.json", "$O/mg_c4_n${N}_$TAG.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d.get("parity_check",{}).get("identical"))
        e = d.get("extra") or {}
        for k, v in e.items():
            if k == "diag": print("diag", json.dumps(v))
            else: print(k, {a: b for a, b in v.items() if a not in ("note", "roofline")}, (v.get("roofline") or {}).get("frac"))
    except Exception as ex:
        print(f, "no line", ex)
PY
