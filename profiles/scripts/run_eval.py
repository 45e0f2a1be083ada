"""Batched greedy evaluation at BASELINE config 3 size (65,536 instances x 2 agents, 10 episodes each) after a short training run:
the command profiles/scripts/capture_r02b.sh profiles for eval_kernel. Prints one JSON line with the wall time of the call."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402

eng = Engine(P.compile_scenario(P.scenario_config3(True)), 65536)
eng.reset()
eng.train(20000)
torch.cuda.synchronize()
out = []
for _ in range(3):
    t0 = time.perf_counter()
    ev = eng.evaluate(10, 0.99, 21.0)
    out.append(time.perf_counter() - t0)
print(json.dumps({"eval_seconds": out, "episodes": int(ev["episodes"].sum()), "successes": int(ev["successes"].sum()),
                  "roofline": {"active_agent_steps_per_launch": float(ev["len_sum"].sum())}}))
