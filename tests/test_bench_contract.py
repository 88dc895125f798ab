"""bench.py contract, CPU side: the reference arm runs anywhere (it is the oracle port on the host cores) and prints ONE JSON
line with the keys the driver reads; the CUDA arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("slides/sec") and d["unit"] == "slides/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and "16 cases x 2 stains" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_no_collective_after_the_non_zero_ranks_left():
    """bench.py's ranks > 0 return once the measurements are done; anything after that point runs on rank 0 alone and must not
    contain a collective (a sustained-state run placed there hung a 2-GPU bench until its timeout)."""
    src = open(os.path.join(REPO, "bench.py")).read()
    marker = "    if rank != 0:\n        if world > 1:\n            dist.destroy_process_group()\n        return\n"
    assert src.count(marker) == 1
    tail = src.split(marker)[1]
    for needle in ("timed(", "dist.barrier", "all_reduce", "all_gather", "run_e2e(", "step(feats"):
        assert needle not in tail, needle
