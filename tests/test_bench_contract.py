"""bench.py's contract on a box without a GPU: the reference arm prints ONE JSON line with the agreed keys; the own arm refuses to run
(no CPU fallback).  The reference arm times the unmodified reference host path (oracle/_ref) when it is built, else the port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, PANDA_BENCH_REF_BUDGET_S="12")      # a short sample: this test checks the contract, not the number
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bn254_g1_msm_2^24_throughput" and d["unit"] == "Mpts/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if d["cpu_baseline"]["kind"] == "reference":           # the fixed / per-point fit that makes the sample comparable with a 2^24 run
        fit = d["cpu_baseline"]["fit"]
        assert fit["fixed_s_per_call"] >= 0 and fit["per_point_us"] > 0 and fit["extrapolated_2^24_Mpts_per_s_all_cores"] > 0
        assert d["config"]["sample_log_points_per_call"] >= 14


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
