"""bench.py contract checks that need no GPU: the reference arm (CPU implementation of the same path) prints one JSON
line with the agreed keys, and under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    cp = _run({})
    assert cp.returncode == 0, cp.stderr[-500:]
    line = json.loads(cp.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ndt_point_evals_per_sec" and line["unit"] == "point-evals/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C4" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    cp = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert cp.returncode == 0 and cp.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    cp = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-extras", "--steps", "1"], capture_output=True, text=True,
                        timeout=600, cwd=str(ROOT))
    assert cp.returncode != 0 and "no CPU fallback" in (cp.stderr + cp.stdout)
