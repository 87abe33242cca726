"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys
the driver reads, and the workload table matches BASELINE.json's configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, SCORUS_BENCH_CPU_WALKERS="4096")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "walker_steps_per_sec" and d["unit"] == "walker-steps/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    # the best-effort flat CPU variant is reported beside the reference-structured port
    assert cb["flat_threaded"]["value"] > 0 and cb["flat_threaded"]["unit"] == d["unit"]
    assert abs(sum(cb["stage_share"].values()) - 1.0) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workloads_cover_baseline_configs():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5 and sorted(bench.WORKLOADS) == ["c1", "c2", "c3", "c4", "c5"]
    w = bench.WORKLOADS
    assert (w["c1"][2], w["c1"][3]) == (16, 2)                       # test_emcee: 2-D Gaussian, 16 walkers
    assert (w["c2"][2], w["c2"][3]) == (65536, 100)                  # 100-D correlated Gaussian, 65,536 walkers
    assert (w["c3"][2], w["c3"][3]) == (1 << 20, 32)                 # 32-D mixture, 2^20 walkers
    assert (w["c4"][2], w["c4"][3], w["c4"][4]) == (64 * 16384, 10, 64)  # 64 rungs x 16,384, 10-D Rastrigin
    assert (w["c5"][2], w["c5"][3]) == (1 << 22, 64)                 # 64-D Rosenbrock, 2^22 walkers
    assert bench.b_alg(64) == 1040 and bench.b_alg(10) == 176        # 16 D + 16 bytes per walker-step


def test_cpu_legs_of_the_other_move_families():
    """bench.py --move twalk|hmc time the serial C restatements on a bounded sample as their cpu_baseline."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    for move in ("twalk", "hmc"):
        r = bench.cpu_move_leg("c4", move, target_seconds=0.5)
        assert r["value"] > 0 and r["cores"] == 1 and r["kind"] == "port" and move in r["sample"].replace("t-walk", "twalk")
