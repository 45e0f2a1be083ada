"""The 192-byte rule of the TMA block fetch on float64 tables: exp5 / exp6 (9 / 10 states = 288 / 320-byte blocks) and the 12-state
chain (384 bytes) with RLRM_QRMB_TMA forced to 0 and 1. Device-timed, tables resident."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402

out = {}
for wl in ("ow_exp6_qrm", "cfg4_qrm"):
    for tma in ("0", "1"):
        os.environ["RLRM_QRMB_TMA"] = tma
        sc = bench.scenario(wl)
        sc.table_dtype = "f64"
        eng = Engine(P.compile_scenario(sc), 65536, device="cuda:0")
        eng.reset()
        for _ in range(3):
            eng.train(256)
        torch.cuda.synchronize()
        a0 = eng.total_active_steps()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.train(256)
        e1.record()
        torch.cuda.synchronize()
        rate = (eng.total_active_steps() - a0) / (e0.elapsed_time(e1) * 1e-3)
        out[f"{wl}_f64_tma{tma}"] = rate
        print(f"{wl} float64 tma={tma}: {rate:.4e} agent-steps/s", flush=True)
        del eng
        torch.cuda.empty_cache()
print(json.dumps({"tma_block_fetch_float64": out}))
