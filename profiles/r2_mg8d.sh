#!/bin/bash
# round 2, 8-GPU diagnostics: clock sampler on/off in the timed region, phases of the one-sided ladder swap
N=${1:-8}; TAG=${2:-d}; K=${3:-8}
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for C in 1 0; do
BENCH_CLOCKS=$C BENCH_DIAG=1 timeout 300 $TR --master-port 2963$C bench.py --gpus $N --steps 100 --warmup 10 --global-every $K --no-e2e --no-verify > $O/mg_c5_n${N}_k${K}_${TAG}_clk$C.json 2> $O/mg_c5_n${N}_k${K}_${TAG}_clk$C.err
echo "c5 clocks=$C rc=$?"; tail -c 300 $O/mg_c5_n${N}_k${K}_${TAG}_clk$C.err
done
SCORUS_SWAP_DIAG=1 BENCH_CLOCKS=0 timeout 300 $TR --master-port 29634 bench.py --gpus $N --steps 100 --warmup 10 --workload c4 --no-e2e --no-verify > $O/mg_c4_n${N}_$TAG.json 2> $O/mg_c4_n${N}_$TAG.err
echo "c4 rc=$?"; grep swap_walkers_peer $O/mg_c4_n${N}_$TAG.err | tail -8
python - <<PY
import json
for f in ("$O/mg_c5_n${N}_k${K}_${TAG}_clk1.json", "$O/mg_c5_n${N}_k${K}_${TAG}_clk0.json", "$O/mg_c4_n${N}_$TAG.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "swap_stage", d.get("swap_stage"))
        e = d.get("extra") or {}
        for k, v in e.items():
            if k == "diag": print("diag", json.dumps(v))
            else: print(k, {a: b for a, b in v.items() if a not in ("note", "roofline")}, (v.get("roofline") or {}).get("frac"))
    except Exception as ex:
        print(f, "no line", ex)
PY
