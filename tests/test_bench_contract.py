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


def test_every_bench_workload_compiles_and_is_described():
    """bench.py's workload table: every entry builds a scenario that compiles to device tables (on the host, no GPU), names its
    dominant kernel and bound, and the committed ncu-derived traffic file only refers to known workloads and existing captures."""
    import json

    sys.path.insert(0, ROOT)
    import bench

    import multiagent_rlrm_b200 as P

    for name, (desc, base, n, iters, bytes_per, kernel) in bench.WORKLOADS.items():
        sc = bench.scenario(name)
        c = P.compile_scenario(sc)
        assert c.n_agents == len(sc.starts) and n > 0 and iters > 0 and kernel and desc and base, name
        assert name in bench.BOUND, f"{name}: no entry in bench.BOUND"
        assert bytes_per is None or bytes_per > 0
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for name, entry in traffic.items():
        if name.startswith("_"):
            continue
        assert name in bench.WORKLOADS, name
        assert entry["dram_bytes_per_active_step"] > 0 and os.path.exists(os.path.join(ROOT, entry["source"])), name
