#!/usr/bin/env python
"""The reference's OfficeWorld run (`office_main --map map1 --experiment <name> --rm-spec FILE --algorithm QL|QRM|QL-lambda`)
on a batch of instances: authored reward-machine spec -> device tables -> fused training -> batched greedy evaluation, and the
product MDP of the first agent through RMEnvironmentWrapper.get_mdp (what the reference feeds its value iteration).

    python examples/train_office_world.py --experiment exp3 --algorithm QRM --instances 8192
    python examples/train_office_world.py --rm-spec tests/fixtures/officeworld_acbd.json --algorithm QL --mdp
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import multiagent_rlrm_b200 as P  # noqa: E402
from multiagent_rlrm_b200.engine import Engine  # noqa: E402
from multiagent_rlrm_b200.maps import office_world_grid  # noqa: E402
from multiagent_rlrm_b200.rmspec import scenario_from_rmspec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rm-spec", default=None, help="authored reward-machine spec (.json / .yaml); default: the built-in --experiment")
    ap.add_argument("--experiment", default="exp4", help="built-in task of config_office.get_experiment_for_map (exp0 .. exp7)")
    ap.add_argument("--map", default="map1")
    ap.add_argument("--algorithm", default="QRM", choices=["QL", "QRM", "QL-lambda"])
    ap.add_argument("--stochastic", action="store_true")
    ap.add_argument("--instances", type=int, default=8192)
    ap.add_argument("--iterations", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--mdp", action="store_true", help="also build the product MDP with get_mdp")
    args = ap.parse_args()

    sc = P.scenario_config2(args.stochastic)
    sc.map_name, sc.seed = args.map, args.seed
    sc.algo = {"QL": "ql", "QRM": "qrm", "QL-lambda": "qlambda"}[args.algorithm]
    if sc.algo == "qlambda":
        sc.lambd, sc.learning_rate, sc.q_init = 0.9, 0.1, 0.0
    if args.rm_spec:
        sc, rm = scenario_from_rmspec(args.rm_spec, sc)
    else:
        exp = P.get_experiment_for_map(args.map, args.experiment)
        sc.rm_transitions = [(s, ev, t, r) for (s, ev), (t, r) in exp["transitions"].items()]
        sc.detector_positions = sorted(exp["positions"])
        rm = sc.reward_machine()
        print(f"built-in task {args.experiment}: {exp['description']}")
    print(f"reward machine: {rm.numbers_state()} states, {len(rm.transitions)} transitions, final {rm.get_final_state()}")
    eng = Engine(P.compile_scenario(sc), args.instances, qlambda_sparse=(sc.algo == "qlambda"))
    eng.reset()
    done = 0
    while done < args.iterations:
        chunk = min(4000, args.iterations - done)
        eng.train(chunk)
        done += chunk
        st = eng.stats_numpy()
        print(f"iter {done:7d}: episodes {int(st['episodes'].sum()):9d}  successes {int(st['successes'].sum()):9d}  "
              f"active agent-steps {eng.total_active_steps():12d}")
    optimal = P.OPTIMAL.get(f"{args.map};{args.experiment}", 30)
    res = P.test_policy_optima_batched(eng, episodi_test=3, optimal_steps=optimal, gamma=sc.gamma)
    print(f"greedy evaluation: success rate {res['success_rate'].mean():.1f} %")

    if args.mdp:
        g = office_world_grid(args.map)
        env = P.MultiAgentOfficeWorld(width=g.width, height=g.height, plants=g.hazards, coffee=g.coffee, letters=g.letters,
                                      walls=g.walls, plants_penalty_value=sc.plants_penalty, wall_penalty_value=sc.wall_penalty,
                                      terminate_on_plants=sc.terminate_on_plants, terminate_hit_walls=sc.terminate_hit_walls)
        env.stochastic, env.all_slip, env.high_prob, env.delay_action = sc.stochastic, sc.all_slip, sc.high_prob, sc.delay_action
        ag = P.AgentRL("a1", env)
        ag.set_initial_position(*sc.starts[0])
        ag.add_state_encoder(P.StateEncoderOfficeWorld(ag))
        ag.add_action_encoder(P.ActionEncoderOfficeWorld(ag))
        ag.set_reward_machine(rm)
        env.add_agent(ag)
        all_P, n_states, n_actions = P.RMEnvironmentWrapper(env, [ag]).get_mdp(args.seed)
        n_term = sum(1 for s in all_P["a1"] if all_P["a1"][s][0][0][3] and all_P["a1"][s][0][0][1] == s)
        print(f"product MDP of a1: {n_states['a1']} states x {n_actions['a1']} actions, {n_term} terminal states")
        # the reference's VI comparison (office_main.py:1117-1140, 1408-1431): solve the model, play the VI policy
        V, policy_vi, _Q = P.value_iteration(all_P["a1"], n_states["a1"], n_actions["a1"], gamma=sc.gamma, theta=1e-4)
        if eng.A == 1 and not sc.stochastic:
            res_vi = P.test_policy_opt_multi_batched(eng, policy_vi[None, :], episodes_test=1, optimal_steps=optimal, gamma=sc.gamma)
            print(f"value iteration: V(start) = {V[(sc.starts[0][1] * g.width + sc.starts[0][0]) * rm.numbers_state()]:.4f}; "
                  f"VI policy success rate {res_vi['success_rate'].mean():.1f} %, {res_vi['avg_timesteps'].mean():.0f} steps "
                  f"(learned policy: {res['success_rate'].mean():.1f} %, {res['avg_timesteps'][res['avg_timesteps'] > 0].mean() if (res['avg_timesteps'] > 0).any() else 0:.0f} steps; optimal {optimal})")


if __name__ == "__main__":
    main()
