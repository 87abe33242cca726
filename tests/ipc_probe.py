import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import scorus_b200 as sb
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s = sb.Sampler(sb.Gaussian(), 1024, 8, seed=1, device=lr)
s.upload(np.full((1024, 8), float(lr + 1)), None)
he, hl = s.ipc_export()
out = [None] * dist.get_world_size()
dist.all_gather_object(out, (he, hl))
try:
    s.ipc_attach(dist.get_world_size(), dist.get_rank(), [o[0] for o in out], [o[1] for o in out])
    print("rank", dist.get_rank(), "ipc attach OK", flush=True)
except Exception as e:
    print("rank", dist.get_rank(), "ipc attach FAILED:", e, flush=True)
dist.barrier(); dist.destroy_process_group()
