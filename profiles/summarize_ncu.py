#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brings back (gpurun_out/) into the small text summaries kept
under profiles/.  Usage, from the repo root (needs `ncu` on PATH, no GPU):

  python profiles/summarize_ncu.py full  gpurun_out/prof_r1_c5_reg_final.ncu-rep c5 "<command line profiled>"
  python profiles/summarize_ncu.py list  gpurun_out/launches_r1_c5.csv c5 "<command line profiled>"
  python profiles/summarize_ncu.py full  gpurun_out/prof_twalk_c5b.ncu-rep c5 "<command>" twalk   (other kernels: a tag)

`full` writes profiles/<round>_<wl>_step_reg_ncu_summary.txt and profiles/traffic_<wl>.json (the DRAM bytes
bench.py reports as roofline.traffic); `list` writes profiles/<round>_<wl>_launch_list_summary.csv (per-kernel
launch counts, mean time and SHARE of the step; per-launch times under ncu are cold-cache and serialised).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUND = os.environ.get("ROUND", "r2")  # prefix of the files written (profiles are named per round)
ALG_BYTES = {"c1": 48 * 16, "c2": 1616 * 65536, "c3": 528 * (1 << 20), "c4": 176 * (1 << 20), "c5": 1040 * (1 << 22)}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(unit, value):
    return float(value) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]


def full(rep, wl, cmd, tag="step_reg"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + int(os.environ.get("NCU_ROW", "0"))]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = [f"{k:75s} {m[k][1]} {m[k][0]}" for k in KEYS if k in m]
    lines += [f"{h:75s} {m[h][1]}" for h in hdr
              if "issue_stalled" in h and "per_issue_active" in h and float(m[h][1]) > 0.03]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 2:
        h2, data = srows[1], srows[2:]
        isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
        mix = collections.Counter()
        for r in data:
            if len(r) <= max(isrc, iex) or not r[isrc].split():
                continue
            t = r[isrc].split()
            mix[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[iex])
        tot = sum(mix.values())
        lines.append("")
        lines.append(f"SASS: {len(data)} instructions; executed warp-instructions {tot}; top opcodes (share):")
        lines.append("  " + ", ".join(f"{op} {c / tot:.3f}" for op, c in mix.most_common(14)))
    out = os.path.join(ROOT, "profiles", f"{ROUND}_{wl}_{tag}_ncu_summary.txt")
    open(out, "w").write(f"{cmd}\n(B200, {ROUND}, one launch = one step; times under ncu are not bench values)\n\n"
                         + "\n".join(lines) + "\n")
    rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
    json.dump({"kernel": m["Kernel Name"][1], "workload": wl, "source": os.path.relpath(out, ROOT),
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "algorithmic_bytes_per_launch": ALG_BYTES.get(wl),
               "note": "rejected rows are not written back, so traffic can be below the algorithmic figure"},
              open(os.path.join(ROOT, "profiles", f"traffic_{wl}.json" if tag == "step_reg" else f"traffic_{wl}_{tag}.json"), "w"),
              indent=1)
    print(out, f"read {rd / 1e9:.3f} GB write {wr / 1e9:.3f} GB")


def launch_list(path, wl, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        agg[r[ik].split("(")[0][-80:]].append(v / 1e3 if r[iu].startswith("n") else v)
    tot = sum(sum(v) for v in agg.values())
    out = ["kernel,launches,total_us,mean_us,share"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k},{len(v)},{sum(v):.1f},{sum(v) / len(v):.1f},{sum(v) / tot:.4f}")
    dst = os.path.join(ROOT, "profiles", f"{ROUND}_{wl}_launch_list_summary.csv")
    open(dst, "w").write(f"# {cmd}\n# per-launch times are cold-cache and serialised under ncu: read SHARES\n" + "\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    mode, path, wl = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    if mode == "full":
        full(path, wl, cmd, sys.argv[5] if len(sys.argv) > 5 else "step_reg")
    else:
        launch_list(path, wl, cmd)
