"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tests/multi_gpu_check.py

Every rank also runs the whole ensemble on its own GPU through the single-GPU path and checks that its shard of the
sharded run is identical (positions bit for bit, log-probs to 1e-12): sharding must not change the chain.
The cases live in scorus_b200/selfcheck.py (bench.py runs the quick subset before timing)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scorus_b200.selfcheck import sharded_parity_cases  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    cases, ok = sharded_parity_cases(quick=False, log=lambda m: print(m, flush=True))
    dist.barrier()
    dist.destroy_process_group()
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"parity_check": {"world": world, "cases": cases, "identical": bool(ok)}}), flush=True)
        if ok:
            print("multi_gpu_check: all cases identical to the single-GPU chain")
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
