"""Sparse Q(lambda) with FEWER than four agents per instance: lanes per agent (RLRM_QLS_LG, read at rlrm_create) against
throughput. BASELINE config 4's scenario with its first A agents; device-timed (CUDA events), tables resident.
usage: python profiles/scripts/r02c_qls_lg_agents.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402

out = {}
for agents in (1, 2, 4):
    for lg in (32, 16, 8, 4):
        if lg > 32 // max(1, 1 << (agents - 1).bit_length()):
            continue
        os.environ["RLRM_QLS_LG"] = str(lg)
        sc = P.scenario_config4()
        sc.starts = sc.starts[:agents]
        n = 262144
        eng = Engine(P.compile_scenario(sc), n, device="cuda:0", qlambda_sparse=True)
        eng.reset()
        for _ in range(3):
            eng.train(64)
        torch.cuda.synchronize()
        a0 = eng.total_active_steps()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.train(64)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rate = (eng.total_active_steps() - a0) / (ms * 1e-3)
        out[f"A{agents}_lg{lg}"] = {"agent_steps_per_s": rate, "ms_per_64_iterations": ms / 10}
        print(f"agents={agents} lanes_per_agent={lg}: {rate:.4e} agent-steps/s, {ms / 10:.3f} ms per 64 iterations", flush=True)
        del eng
        torch.cuda.empty_cache()
print(json.dumps({"qls_lanes_per_agent": out}))
