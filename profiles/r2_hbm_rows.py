#!/usr/bin/env python
"""What HBM delivers for K1's access pattern: whole 512-byte rows of a 2 GiB ensemble read once each in a scrambled order
(scorus_peer_read_bandwidth in a world of one rank reads the context's OWN ensemble), next to the same bytes read
sequentially and to a copy-engine device-to-device copy."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scorus_b200 as sb  # noqa: E402

for dim in (64, 32, 10):
    n = (1 << 22) if dim == 64 else (1 << 20)
    s = sb.Sampler(sb.Gaussian(0.5), n, dim, seed=1)
    s.init_ensemble(0.9, 1.1)
    he, hl = s.ipc_export()
    s.ipc_attach(1, 0, [he], [hl])
    ce, sm, rr = s.peer_read_bandwidth(0, 0, 10)
    print(f"dim {dim:3d} ({8 * dim} B rows, {n * dim * 8 / 2**20:.0f} MiB): copy engine d2d {ce:.0f} GB/s (x2 = read + write traffic), "
          f"SM sequential 16-byte loads {sm:.0f} GB/s, whole rows in a scrambled order {rr:.0f} GB/s")
    s.close()
