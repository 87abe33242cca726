#!/usr/bin/env python
"""Differential check of the ladder swap's table stores: two samplers with the same seed run the c4 cycle (steps,
swap, step ...); one writes the sub-chain table through the shared-memory stage, the other straight to global memory
(SCORUS_SWAP_DIRECT_STORE is read at every swap).  Prints the first phase after which the two states differ."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scorus_b200 as sb  # noqa: E402

n, dim, nb = 1 << 20, 10, 64
beta = np.array([1.0 / (1.0 + 0.1 * k) for k in range(nb)])
beta = np.sort(beta)[::-1].copy()
ufs = sb.UpdateFlagSpec.Prob(0.2)


def make():
    s = sb.Sampler(sb.Rastrigin(10.0), n, dim, seed=2024)
    s.init_ensemble(-0.01, 0.01)
    return s


a, b = make(), make()
fuse = os.environ.get("FUSE", "1")
for cyc in range(12):
    for name, s in (("stage", a), ("direct", b)):
        os.environ["SCORUS_SWAP_DIRECT_STORE"] = "1" if name == "direct" else "0"
        s.sample_pt(2.0, ufs, beta, 9)
        s.swap_walkers(beta)
        if fuse == "1":
            s.sample_pt(2.0, ufs, beta, 1)
    ea, la = a.download()
    eb, lb = b.download()
    bad = np.nonzero(np.any(ea != eb, axis=1) | (la != lb))[0]
    print(f"cycle {cyc}: differing walkers {bad.size}", bad[:8], [(int(w) // (n // nb), int(w) % (n // nb)) for w in bad[:8]], flush=True)
    if bad.size:
        break
