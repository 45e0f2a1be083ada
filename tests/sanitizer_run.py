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
    # round 2: float64 tables, the thread-block-cluster shared learner (forced), per-agent machines under Q(lambda)
    f64 = P.scenario_config3(False)
    f64.table_dtype = "f64"
    g = P.frozen_lake_grid("map1").goals
    pa = P.scenario_config3(False)
    pa.algo, pa.lambd, pa.learning_rate, pa.q_init = "qlambda", 0.8, 0.2, 0.0
    pa.starts, pa.detector_positions = [(5, 0), (0, 0)], sorted(g.values())
    pa.rm_transitions_per_agent = [P.tables.frozen_lake_abc_transitions(), [("p0", g["C"], "p1", 3.0), ("p1", g["A"], "p2", 7.0)]]
    cases += [("cfg3 ql float64", f64, {}), ("cfg5 shared, cluster kernel", P.scenario_config5(True), {"_reserved": 4}),
              ("per-agent machines qlambda", pa, {}), ("per-agent machines qlambda sparse", pa, {"qlambda_sparse": True})]
    for name, sc, kw in cases:
        n = 37
        kw = dict(kw)
        c = P.compile_scenario(sc)
        c.config.reserved = kw.pop("_reserved", 0)
        eng = Engine(c, n, **kw)
        eng.reset()
        eng.train(40, trace=True)
        eng.train(25)
        eng.iterate()                      # rlrm_iterate: record (and reward) into page-locked host memory
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
    # the N = 1 learner path: rlrm_select_action / rlrm_update_list on a page-locked staging block
    ql = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=0.1, state_space_size=50, action_space_size=4, use_qrm=True)
    ql.update(0, 0, 0, 0.0, False, info={"qrm_experience": [(s, s % 4, 1.0, (s + 1) % 50, False, 0, 0, 0, 0, 0.0) for s in range(40)]})
    ql.choose_action(3)
    torch.cuda.synchronize()
    print("ok learners")


if __name__ == "__main__":
    main()
