"""bench.py contract on CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints ONE JSON
line with the keys the driver reads, and the default arm refuses to run without a GPU instead of falling back."""
import json
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["metric"] == "seconds per 49-frame 480x720 50-step CFG denoise"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the CPU arm measures a config-1 step and says that its value is an extrapolation, not the same job
    assert d["extrapolated"] is True and d["same_config"] is False and d["measured_step_seconds_config1"] > 0
    assert "15+30 layers" in d["config"]["workload"] and d["cpu_baseline"]["cores"] <= 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_default_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line may be printed without a GPU"
