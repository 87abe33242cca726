#!/usr/bin/env python
"""Per-CUDA-source-line totals of an ncu report captured with --import-source on (kernels built with -lineinfo):
  python profiles/ncu_lines.py report.ncu-rep [top N]   ->   share of executed warp-instructions, share of stall samples, file:line"""
import csv, sys, subprocess, collections
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
cur=None; hdr=None; agg=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; iex=hdr.index("Instructions Executed"); ist=hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr and r[0].isdigit():
        try: agg.append((int(r[iex]), int(r[ist]), cur, int(r[0]), r[1].strip()[:95]))
        except: pass
tot=sum(a[0] for a in agg); tots=sum(a[1] for a in agg)
print("total executed", tot, "samples", tots)
byfile=collections.Counter()
for a in agg: byfile[a[2]]+=a[0]
print(byfile.most_common())
for a in sorted(agg,reverse=True)[:topn]:
    print(f"{a[0]/tot*100:5.1f}% {a[1]/max(tots,1)*100:5.1f}%  {a[2]}:{a[3]}  {a[4]}")
