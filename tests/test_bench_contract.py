"""CPU: bench.py's reference arm prints ONE JSON line with the contract's keys (a tiny sample, a couple of seconds)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "0.2"], capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and "sample" in cb and cb["value"] == line["value"]
    assert set(line["config"]) >= {"workload", "baseline_config", "instances_per_gpu", "agents", "iters_per_step"}
    assert line["vs_baseline"] is None and line["scaling"] == "weak" and line["dtype"] == "f32" and line["data"] == "synthetic"


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
