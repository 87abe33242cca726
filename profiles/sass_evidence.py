#!/usr/bin/env python
"""Static instruction counts (memory / synchronisation opcodes, DFMA) of selected kernels of the in-tree objects:
  python profiles/sass_evidence.py   ->   the body of profiles/r2b_sass_evidence.txt (needs cuobjdump, no GPU)"""
import subprocess, re, collections, sys
KEEP = re.compile(r"^(LDG|STG|LDS|STS|SHFL|MUFU|BAR|REDG|ATOMG|ATOMS|RED|UBLKCP|SYNCS|DMMA|CCTL|ACQBULK|UTMACMDFLUSH|DFMA|VOTE)")
def kernels(obj):
    out = subprocess.run(["cuobjdump","-sass",obj],capture_output=True,text=True).stdout
    cur=None; res={}
    for line in out.splitlines():
        m=re.search(r"Function : (\S+)", line)
        if m: cur=m.group(1); res[cur]=collections.Counter(); res[cur]["_n"]=0; continue
        m=re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            res[cur]["_n"]+=1
            op=m.group(1)
            if KEEP.match(op):
                res[cur][op if not op.startswith("DFMA") else "DFMA"]+=1
    return res
want = [("build/obj/api_core.o", ["swap_paths_kernel","swap_walk_kernelILb1","apply_swap_kernel"]),
        ("build/obj/launch_LpRosenbrock.o", ["step_reg_kernelINS_12LpRosenbrockELi32ELi1ELi0ELb0","step_reg_kernelINS_12LpRosenbrockELi32ELi1ELi0ELb1"]),
        ("build/obj/launch_LpRastrigin.o", ["step_reg_kernelINS_11LpRastriginELi2ELi3ELi0ELb0","step_reg_kernelINS_11LpRastriginELi2ELi3ELi0ELb1"]),
        ("build/obj/twalk_LpRosenbrock.o", ["twalk_direct_kernelINS_12LpRosenbrockELi32ELi1"]),
        ("build/obj/hmc_LpCorrGaussian.o", ["hmc_kernelINS_14LpCorrGaussianELi32ELi2"])]
for obj, pats in want:
    ks = kernels(obj)
    for name, c in ks.items():
        if any(p in name for p in pats) and "f32" not in name:
            n = c.pop("_n")
            print(f"{obj.split('/')[-1]} :: {name}")
            print(f"    {n} instructions; " + ", ".join(f"{k} {v}" for k, v in c.most_common()))
