"""The driver's reference arm (`python bench.py --impl reference ...`) needs no GPU: run it at a small size and check the
contract of the JSON line (keys, units, identical `config` shape to the GPU arm, the to-tolerance solve)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_at_small_size():
    env = dict(os.environ, FC_BENCH_N="20")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                          "--warmup", "1", "--ref-ranks", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "pcorr_dpcg_iterations_per_second" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert set(d["config"]) == {"workload", "cells", "nnz", "solver", "sor", "l2"} and d["config"]["cells"] == 8000
    tt = d["to_tolerance"]
    assert tt["iters_to_tol"] > 10 and tt["resl"] / tt["res0"] < 1e-8
    assert d["serial"]["cores"] == 1 and d["serial"]["value"] > 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, FC_BENCH_N="20", RANK="3", WORLD_SIZE="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
