"""rlrm_iterate at BASELINE config 3 size (65,536 instances x 2 agents): 300 one-iteration launches with the record written into
page-locked host memory and one stream synchronisation each — the command profiles/scripts/capture_r02b.sh profiles for the
one-iteration launch (the generic train_kernel with the reward output). Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.vec import BatchedRMEnvironment  # noqa: E402

env = BatchedRMEnvironment(P.scenario_config3(True), 65536)
env.reset()
for _ in range(50):
    env.iterate()
a0 = env.engine.total_active_steps()
t0 = time.perf_counter()
for _ in range(300):
    env.iterate()
dt = time.perf_counter() - t0
active = env.engine.total_active_steps() - a0
print(json.dumps({"us_per_iteration": dt / 300 * 1e6, "agent_steps_per_s": active / dt,
                  "roofline": {"active_agent_steps_per_launch": active / 300}}))
