"""CPU suite: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (the driver computes
its speed-up ratio from it) and the refusal to time 'our' arm without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(*flags, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    done = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--mode", "noise")
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [ln for ln in done.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly ONE JSON line on stdout"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "decoded images/s" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("cfg5") and line["data"] == "synthetic" and line["dtype"] == "f32"
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "Decoder" in base["sample"] or "torch_port" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="a CUDA device is present: this box would run the real bench")
def test_our_arm_refuses_to_run_without_a_cuda_device():
    done = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert done.returncode != 0 and done.stdout.strip() == ""
    assert "no CPU fallback" in done.stderr
