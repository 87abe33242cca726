#!/bin/bash
# round 2 multi-GPU run: parity cases (sharded == single-GPU chain), then bench lines at N GPUs.
# usage: bash profiles/r2_mg.sh N [tag]
N=${1:-2}; TAG=${2:-a}
O=gpurun_out/r2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tests/multi_gpu_check.py > $O/multi_gpu_check_n${N}_$TAG.log 2>&1
echo "multi_gpu_check rc=$?"; tail -30 $O/multi_gpu_check_n${N}_$TAG.log
for k in 4 1 8; do
  timeout 600 $TR --master-port 2962$k bench.py --gpus $N --steps 100 --warmup 10 --global-every $k > $O/mg_c5_n${N}_k${k}_$TAG.json 2> $O/mg_c5_n${N}_k${k}_$TAG.err
  echo "bench k=$k rc=$?"; tail -c 600 $O/mg_c5_n${N}_k${k}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/mg_c5_n${N}_k${k}_$TAG.json").read().strip().splitlines()[-1])
    print("k=$k ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d.get("parity_check",{}).get("identical"))
    print(json.dumps(d.get("extra"), indent=1)[:1800])
except Exception as e:
    print("no line", e)
PY
done
