#!/usr/bin/env python
"""The reference's quick start (README: `frozen_lake_main --map map1 --num-episodes 2000`) on a batch of instances:
FrozenLake map1, 2 agents, built-in A -> B -> C reward machine (or --rm-spec FILE), QLearning(use_qrm=True).

    python examples/train_frozen_lake.py --instances 4096 --iterations 20000 [--slippery] [--rm-spec tests/fixtures/frozenlake_abc.json]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=4096)
    ap.add_argument("--iterations", type=int, default=20000)
    ap.add_argument("--slippery", action="store_true")
    ap.add_argument("--rm-spec", default=None)
    ap.add_argument("--seed", type=int, default=111)
    ap.add_argument("--save", default=None, help="write q_tables.npz in the reference's format")
    args = ap.parse_args()

    sc = P.scenario_config3() if args.slippery else P.scenario_config1()
    sc.seed = args.seed
    rm = None
    if args.rm_spec:
        rm, spec = P.load_reward_machine(args.rm_spec, "frozen_lake", "map1")
        print(f"loaded RM spec {spec.name}: {len(rm.transitions)} transitions, final state {rm.get_final_state()}")
    eng = Engine(P.compile_scenario(sc, rm=rm), args.instances)
    eng.reset()
    eng.reset()  # the reference driver resets once before the episode loop and once per episode
    done = 0
    while done < args.iterations:
        chunk = min(2000, args.iterations - done)
        eng.train(chunk)
        done += chunk
        st = eng.stats_numpy()
        print(f"iter {done:7d}: episodes {int(st['episodes'].sum()):9d}  successes {int(st['successes'].sum()):9d}  "
              f"active agent-steps {eng.total_active_steps():12d}")
    res = P.test_policy_optima_batched(eng, episodi_test=5, optimal_steps=21, gamma=sc.gamma)
    print(f"greedy evaluation: success rate {res['success_rate'].mean():.1f} %, mean length {res['avg_timesteps'][res['avg_timesteps'] > 0].mean():.1f}")
    if args.save:
        print("saved", P.save_q_tables(eng, path=args.save))


if __name__ == "__main__":
    main()
