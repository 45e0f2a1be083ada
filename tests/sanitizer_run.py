"""Small run through every kernel of csrc/rlrm_b200.cu, meant to be executed under compute-sanitizer (memcheck /
racecheck), the GPU analogue of the race-detection row of SURVEY.md §5:
    python tests/sanitizer_run.py && compute-sanitizer --tool memcheck python tests/sanitizer_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402


def main():
    sc4q = P.scenario_config4()
    sc4q.algo, sc4q.learning_rate, sc4q.q_init = "qrm", 0.1, 2.0
    cases = [("cfg3 qrm fast", P.scenario_config3(True), {}), ("cfg3 ql", P.scenario_config3(False), {}),
             ("cfg2 office", P.scenario_config2(True), {}), ("office chain12 qrm", sc4q, {}),
             ("cfg4 qlambda dense", P.scenario_config4(), {}), ("cfg4 qlambda sparse", P.scenario_config4(), {"qlambda_sparse": True}),
             ("cfg5 shared", P.scenario_config5(True), {}), ("cfg5 tables", P.scenario_config5(False), {})]
    for name, sc, kw in cases:
        n = 37
        eng = Engine(P.compile_scenario(sc), n, **kw)
        eng.reset()
        eng.train(40, trace=True)
        eng.train(25)
        if not kw.get("qlambda_sparse"):
            for _ in range(6):
                eng.iterate_unfused()
        eng.sync_tables(with_traces=True)
        eng.evaluate(1, 0.9, 10.0, max_iters=30)
        mask = torch.zeros(n, dtype=torch.uint8)
        mask[::3] = 1
        eng.reset(mask)
        q, ev, r = eng.rm_step(torch.zeros(8, dtype=torch.uint8), torch.arange(8, dtype=torch.int16))
        torch.cuda.synchronize()
        print("ok", name, eng.launches, "launches")


if __name__ == "__main__":
    main()
