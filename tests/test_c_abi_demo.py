"""The C ABI from plain C, without Python or torch in the process (`extern "C"`, plain pointers and sizes).

CPU: include/rlrm_b200.h compiles as C and examples/c_abi_demo.c links against librlrm_b200.so.
GPU: the C program runs BASELINE config 1 end to end (one instance, 75,000 iterations = 2,114 episodes) from a scenario blob and
its Q tables equal the LIVE REFERENCE's (fixture long_cfg1_det_qrm) bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import load_golden

import multiagent_rlrm_b200 as P
from multiagent_rlrm_b200 import _lib
from multiagent_rlrm_b200.tables import dump_blob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_demo(out_dir):
    _lib.load()  # makes sure librlrm_b200.so exists (builds it when missing)
    exe = os.path.join(str(out_dir), "c_abi_demo")
    lib_dir = os.path.join(ROOT, "multiagent-rl-rm_b200")
    cmd = [shutil.which("gcc") or "gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L" + lib_dir, "-lrlrm_b200", "-L" + os.path.join(CUDA, "lib64"),
           "-lcudart", "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return exe


def test_header_is_c_and_demo_links(tmp_path):
    exe = build_demo(tmp_path)
    assert os.path.exists(exe)
    blob = tmp_path / "cfg1.blob"
    dump_blob(P.compile_scenario(P.scenario_config1()), blob)
    assert blob.stat().st_size > 1000


@pytest.mark.gpu
def test_c_program_reproduces_the_reference(tmp_path, cuda_device):
    meta, ref = load_golden("long_cfg1_det_qrm")
    exe = build_demo(tmp_path)
    blob, q_out = tmp_path / "cfg1.blob", tmp_path / "q.bin"
    dump_blob(P.compile_scenario(P.Scenario.from_dict(meta["scenario"])), blob)
    res = subprocess.run([exe, str(blob), "1", str(meta["n_iters"]), str(q_out)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert f"episodes={2 * int(ref['episode_end'].sum())}" in res.stdout and "launches=3" in res.stdout, res.stdout
    q = np.fromfile(q_out, dtype=np.float32).reshape(ref["q_final"].shape)
    assert np.array_equal(q, ref["q_final"])
