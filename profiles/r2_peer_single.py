#!/usr/bin/env python
"""The one-sided exchange step (scorus_sample_pt_peer, kPeer == 2) on ONE GPU with a world of one rank -- every partner
is local -- so that the kernel can be profiled with ncu (which must not wrap a multi-rank command).  It shows the
kernel's instruction count, occupancy and shared-memory stage; the NVLink side is measured by bench.py (extra)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scorus_b200 as sb  # noqa: E402

n, dim, steps = 1 << 22, 64, int(os.environ.get("STEPS", "6"))
s = sb.Sampler(sb.Rosenbrock(), n, dim, seed=1)
s.init_ensemble(0.9, 1.1)
he, hl = s.ipc_export()
s.ipc_attach(1, 0, [he], [hl])
spec = sb.UpdateFlagSpec.Prob(0.5)
s.sample_pt(2.0, spec, [1.0], 3)
s.sync()
t0 = time.perf_counter()
s.sample_pt_peer(2.0, spec, [1.0], steps)
s.sync()
t1 = time.perf_counter()
s.sample_pt(2.0, spec, [1.0], steps)
s.sync()
t2 = time.perf_counter()
print(f"one-sided step, all partners local: {1e3 * (t1 - t0) / steps:.3f} ms/step; single-GPU step: {1e3 * (t2 - t1) / steps:.3f} ms/step",
      s.stats())
