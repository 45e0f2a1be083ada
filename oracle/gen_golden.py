"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the LIVE reference (needs /root/reference; run here, not on
the GPU box):   python oracle/gen_golden.py

Each fixture holds the scenario (JSON), the run shape and every per-iteration quantity the reference produced
(oracle/ref_harness.py) with Philox-injected randomness, plus the final Q / trace tables.
"""
import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, _HERE)

import multiagent_rlrm_b200 as P  # noqa: E402
import ref_harness as H  # noqa: E402
from multiagent_rlrm_b200.maps import office_world_grid  # noqa: E402
from multiagent_rlrm_b200.tables import Scenario, frozen_lake_abc_transitions  # noqa: E402


def scenarios():
    S = {}
    S["cfg1_det_qrm"] = (P.scenario_config1(), 4, 600, "f32", 1)
    S["cfg1_det_qrm_f64"] = (P.scenario_config1(), 2, 400, "f64", 1)
    S["cfg3_slip_qrm"] = (P.scenario_config3(True), 4, 800, "f32", 1)
    S["cfg3_slip_ql"] = (P.scenario_config3(False), 4, 800, "f32", 1)
    S["cfg3_slip_ql_f64"] = (P.scenario_config3(False), 2, 500, "f64", 1)

    sc = P.scenario_config3(False)
    sc.delay_action, sc.penalty_amount, sc.learning_rate = True, -5, 0.5
    sc.epsilon_start, sc.epsilon_end, sc.epsilon_decay, sc.seed = 0.5, 0.05, 0.9, 77
    S["fl_delay_penalty_epsdecay_ql"] = (sc, 3, 700, "f32", 1)

    sc = P.scenario_config3(False)
    sc.learning_rate, sc.seed, sc.penalty_amount = None, 5, -1.5
    S["fl_lr_none_ql"] = (sc, 3, 600, "f32", 0)

    sc = P.scenario_config3(True)
    sc.learning_rate, sc.seed = None, 6
    S["fl_lr_none_qrm"] = (sc, 2, 600, "f32", 1)

    sc = P.scenario_config3(False)
    sc.algo, sc.lambd, sc.learning_rate, sc.q_init, sc.epsilon_start, sc.epsilon_end = "qlambda", 0.8, 0.2, 0.0, 0.2, 0.2
    sc.seed = 9
    S["fl_qlambda"] = (sc, 2, 500, "f32", 1)

    sc = P.scenario_config3(True)
    sc.reward_modifier, sc.seed = 0.5, 10
    S["fl_reward_modifier_qrm"] = (sc, 2, 500, "f32", 1)

    sc = P.scenario_config3(False)
    sc.rm_transitions = [("state0", None, "state0", -0.125)] + frozen_lake_abc_transitions()
    sc.detector_positions = [tuple(e) for (_s, e, _t, _r) in frozen_lake_abc_transitions()]
    sc.seed = 11
    S["fl_none_event_step_cost_ql"] = (sc, 2, 500, "f32", 1)

    sc = P.scenario_config5(False)
    sc.random_start_positions, sc.seed, sc.starts = True, 51, sc.starts[:3]
    S["fl_random_starts_3agents_qrm"] = (sc, 3, 700, "f32", 1)

    sc = P.scenario_config3(False)
    sc.random_start_positions, sc.seed, sc.starts = True, 52, sc.starts + [(9, 9), (7, 3), (9, 0), (0, 9)]
    S["fl_random_starts_6agents_ql"] = (sc, 2, 500, "f32", 1)

    gl = P.frozen_lake_grid("map1").goals
    other = [("p0", gl["C"], "p1", 3.0), ("p1", gl["A"], "p2", 7.0)]
    third = [("w0", gl["B"], "w1", 1.0), ("w1", gl["A"], "w2", 1.0), ("w2", gl["C"], "w3", 1.0), ("w3", gl["B"], "w4", 5.0), ("w1", gl["C"], "w0", -1.0)]
    for algo, lr, seed in (("qrm", 1.0, 61), ("ql", 0.3, 62)):
        sc = P.scenario_config3(algo == "qrm")
        sc.algo, sc.learning_rate, sc.seed = algo, lr, seed
        sc.starts = [(5, 0), (0, 0), (9, 9)]
        sc.detector_positions = sorted(gl.values())
        sc.rm_transitions_per_agent = [frozen_lake_abc_transitions(), other, third]
        S[f"fl_per_agent_rms_{algo}"] = (sc, 2, 700, "f32", 1)
    import copy

    # agents with different machines under Q(lambda) and under potential-based shaping (frozen_lake_main.py allows
    # --rm-spec-a1 / --rm-spec-a2 with any learner, :97-98, 228-252)
    sc = copy.deepcopy(S["fl_per_agent_rms_ql"][0])
    sc.algo, sc.lambd, sc.learning_rate, sc.q_init, sc.seed = "qlambda", 0.8, 0.2, 0.0, 63
    S["fl_per_agent_rms_qlambda"] = (sc, 2, 500, "f32", 1)
    S["fl_per_agent_rms_qlambda_f64"] = (sc, 1, 500, "f64", 1)
    sc = copy.deepcopy(S["fl_per_agent_rms_qrm"][0])
    sc.use_rsh, sc.rs_kind, sc.learning_rate, sc.seed = True, "vi", 0.5, 64
    S["fl_per_agent_rms_shaping_qrm"] = (sc, 2, 700, "f32", 1)
    sc = copy.deepcopy(S["fl_per_agent_rms_ql"][0])
    sc.use_rsh, sc.rs_kind, sc.rs_alpha, sc.seed = True, "distance", 7, 65
    S["fl_per_agent_rms_shaping_ql"] = (sc, 2, 600, "f32", 1)

    S["cfg5_fl_4agents_qrm"] = (P.scenario_config5(False), 2, 500, "f32", 1)

    S["cfg2_office_det_ql"] = (P.scenario_config2(False), 2, 1500, "f32", 1)
    S["cfg2_office_slip_ql"] = (P.scenario_config2(True), 2, 1500, "f32", 1)

    g = office_world_grid("map1")
    coffee, letter, office = g.coffee, g.letters[0], g.goals["O"]
    exp3 = [("state0", letter, "state1", 0), ("state0", coffee[0], "state2", 0), ("state0", coffee[1], "state2", 0),
            ("state2", letter, "state3", 0), ("state1", coffee[0], "state3", 0), ("state1", coffee[1], "state3", 0),
            ("state3", office, "state4", 1)]
    det = sorted(set(g.goals.values()) | set(g.coffee) | set(g.letters))
    sc = Scenario(env="office_world", starts=[(2, 7), (5, 4)], rm_transitions=exp3, detector_positions=det,
                  stochastic=True, all_slip=True, high_prob=0.7, wall_penalty=-1, algo="qrm", learning_rate=0.5,
                  gamma=0.9, epsilon_start=0.3, epsilon_end=0.3, epsilon_decay=1.0, q_init=2.0, driver="office_main",
                  seed=21)
    S["ow_allslip_wallpen_exp3_qrm"] = (sc, 2, 1300, "f32", 1)

    sc = Scenario(env="office_world", starts=[(2, 7), (5, 0)], rm_transitions=exp3, detector_positions=det,
                  stochastic=True, high_prob=0.8, wall_penalty=-2, plants_penalty=-100, terminate_on_plants=True,
                  terminate_hit_walls=True, algo="ql", learning_rate=0.25, gamma=0.9, epsilon_start=0.4,
                  epsilon_end=0.1, epsilon_decay=0.95, q_init=1.0, driver="office_main", seed=22)
    S["ow_terminate_plants_walls_ql"] = (sc, 3, 600, "f32", 1)

    sc = Scenario(env="office_world", starts=[(2, 7)], rm_transitions=P.tables.office_acbd_transitions(True),
                  detector_positions=det, stochastic=True, delay_action=True, algo="qrm", learning_rate=1.0,
                  gamma=0.9, epsilon_start=0.1, epsilon_end=0.1, epsilon_decay=1.0, q_init=2.0, driver="office_main",
                  seed=23)
    S["ow_delay_completed_acbd_qrm"] = (sc, 2, 1300, "f32", 1)

    sc = P.scenario_config3(False)
    sc.use_rsh, sc.rs_kind, sc.rs_gamma, sc.seed, sc.learning_rate = True, "vi", 0.9, 31, 0.5
    S["fl_shaping_vi_ql"] = (sc, 2, 500, "f32", 1)

    sc = P.scenario_config3(True)
    sc.use_rsh, sc.rs_kind, sc.rs_alpha, sc.seed, sc.learning_rate = True, "distance", 5, 32, 0.5
    S["fl_shaping_distance_qrm"] = (sc, 2, 500, "f32", 1)

    sc = Scenario(env="office_world", starts=[(2, 7)], rm_transitions=exp3, detector_positions=det, stochastic=True,
                  high_prob=0.8, algo="qrm", learning_rate=0.1, gamma=0.9, epsilon_start=0.1, epsilon_end=0.1,
                  epsilon_decay=1.0, q_init=2.0, driver="office_main", seed=33, use_rsh=True, rs_kind="vi", rs_gamma=0.9)
    S["ow_shaping_vi_exp3_qrm"] = (sc, 2, 1200, "f32", 1)

    for mp, starts, seed in (("map0", [(0, 0), (5, 5)], 41), ("map2", [(2, 7), (6, 3)], 42), ("map3", [(2, 7), (10, 10), (0, 0)], 43),
                             ("map4", [(2, 7), (14, 14)], 44)):
        gm = office_world_grid(mp)
        tr = [("u0", gm.coffee[0], "u1", 0.5), ("u0", gm.coffee[1], "u1", 0.5), ("u1", gm.letters[0], "u2", 0), ("u1", gm.goals["A"], "u0", -1),
              ("u2", gm.goals["O"], "u3", 2)]
        if mp == "map4":  # degenerate on purpose: the LAST inserted transition leads back to u0, so final == initial state and
            tr = tr[:3] + [tr[4], tr[3]]  # every wrapper step reports termination (episode of length 1, reset every iteration)
        detm = sorted(set(gm.goals.values()) | set(gm.coffee) | set(gm.letters))
        sc = Scenario(env="office_world", map_name=mp, starts=starts, rm_transitions=tr, detector_positions=detm, stochastic=True,
                      high_prob=0.75, wall_penalty=-0.5, plants_penalty=-10, algo="qrm", learning_rate=0.5, gamma=0.9,
                      epsilon_start=0.5, epsilon_end=0.5, epsilon_decay=1.0, q_init=1.0, driver="office_main", seed=seed)
        S[f"ow_{mp}_walls_qrm"] = (sc, 2, 700, "f32", 1)

    S["cfg4_office_chain12_qlambda"] = (P.scenario_config4(), 1, 1300, "f32", 1)
    S["cfg4_office_chain12_qlambda_f64"] = (P.scenario_config4(), 1, 300, "f64", 1)

    sc = P.scenario_config4()
    sc.algo, sc.learning_rate, sc.q_init, sc.map_name = "qrm", 0.1, 2.0, "map1"
    sc.starts = sc.starts[:2]
    S["ow_chain12_qrm"] = (sc, 1, 1100, "f32", 1)

    # long runs: the short fixtures above end before any agent completes its machine. These reach the learned regime
    # (all RM states, goal rewards propagating, successful episodes); long_cfg1 is BASELINE configs[0] end to end
    # (deterministic FrozenLake map1, 2 agents, QRM, 2000+ back-to-back episodes of the reference driver loop).
    S["long_cfg1_det_qrm"] = (P.scenario_config1(), 1, 75000, "f32", 1)
    S["long_cfg3_slip_qrm"] = (P.scenario_config3(True), 1, 40000, "f32", 1)
    S["long_cfg2_office_det_ql"] = (P.scenario_config2(False), 1, 30000, "f32", 1)
    g1 = office_world_grid("map1")  # Q(lambda) with successful episodes: coffee (either machine) then the office, 2 agents
    tr = [("u0", g1.coffee[0], "u1", 1.0), ("u0", g1.coffee[1], "u1", 1.0), ("u1", g1.goals["O"], "u2", 5.0)]
    sc = Scenario(env="office_world", starts=[(2, 7), (6, 3)], rm_transitions=tr, detector_positions=sorted(set(g1.coffee) | {g1.goals["O"]}),
                  stochastic=True, high_prob=0.8, algo="qlambda", learning_rate=0.1, lambd=0.9, gamma=0.9, epsilon_start=0.3,
                  epsilon_end=0.3, epsilon_decay=1.0, q_init=0.0, driver="office_main", seed=71)
    S["long_office_coffee_qlambda"] = (sc, 1, 20000, "f32", 1)

    # QLearningLambda(learning_rate=None): lr = 1 / visits, float64 arithmetic on float32 tables (qlearning_lambda.py:44-49, 63)
    sc = P.scenario_config3(False)
    sc.algo, sc.lambd, sc.learning_rate, sc.q_init, sc.epsilon_start, sc.epsilon_end, sc.seed = "qlambda", 0.7, None, 0.5, 0.2, 0.2, 12
    sc.penalty_amount = -2.0
    S["fl_qlambda_lr_none"] = (sc, 2, 900, "f32", 1)
    S["fl_qlambda_lr_none_f64"] = (sc, 1, 500, "f64", 1)
    # float64 companions (the reference's NATIVE table type) of fixtures above: the device's float64 table mode
    # (RLRM_TABLE_F64) must reproduce these bit for bit, 20,000-iteration runs included
    for base, n, t in (("cfg3_slip_qrm", 2, 800), ("cfg2_office_slip_ql", 2, 1500), ("fl_lr_none_ql", 2, 600), ("fl_lr_none_qrm", 2, 600),
                       ("fl_shaping_vi_ql", 2, 500), ("fl_shaping_distance_qrm", 2, 500), ("fl_per_agent_rms_qrm", 2, 700),
                       ("ow_chain12_qrm", 1, 1100), ("fl_qlambda", 2, 500), ("ow_allslip_wallpen_exp3_qrm", 2, 1300),
                       ("long_cfg3_slip_qrm", 1, 20000), ("long_cfg2_office_slip_ql", 1, 20000), ("long_office_coffee_qlambda", 1, 20000)):
        src = base if base in S else base.replace("long_cfg2_office_slip_ql", "cfg2_office_slip_ql")
        S[base + "_f64"] = (S[src][0], n, t, "f64", S[src][4])
    return S


def main(only=None):
    out_dir = os.path.join(os.path.dirname(_HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (sc, n, t, real, pre) in scenarios().items():
        if only and name not in only:
            continue
        ref = H.run_reference(sc.to_dict(), n, t, table_dtype=np.float32 if real == "f32" else np.float64, pre_resets=pre)
        meta = {"scenario": sc.to_dict(), "n_instances": n, "n_iters": t, "real": real, "pre_resets": pre,
                "generator": "oracle/gen_golden.py", "reference": "Alee08/multiagent-rl-rm v0.3.0"}
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), meta=json.dumps(meta), **ref)
        print(f"{name}: N={n} T={t} {real} episodes={int(ref['episode_end'].sum())} "
              f"size={os.path.getsize(os.path.join(out_dir, name + '.npz')) // 1024} KiB")


if __name__ == "__main__" and "--eval" not in sys.argv:
    main(set(sys.argv[1:]) or None)


def gen_eval():
    """Greedy-evaluation fixtures: trained tables come from the training fixtures' q_final."""
    import ref_eval

    import oracle as O

    out_dir = os.path.join(os.path.dirname(_HERE), "tests", "golden")
    # (training fixture, extra oracle training iterations so that some evaluation episodes succeed, episodes, gamma, optimal)
    for name, extra, n_ep, gamma, opt in (("cfg1_det_qrm", 30000, 3, 0.99, 21.0), ("cfg3_slip_qrm", 60000, 4, 0.99, 21.0),
                                          ("cfg2_office_det_ql", 0, 2, 0.9, 30.0), ("ow_allslip_wallpen_exp3_qrm", 40000, 2, 0.9, 29.0)):
        z = np.load(os.path.join(out_dir, name + ".npz"))
        meta = json.loads(str(z["meta"]))
        q_tables = z["q_final"]
        if extra:  # tables are only INPUT data for the evaluation: train them further with the (reference-pinned) oracle
            sc = P.Scenario.from_dict(meta["scenario"])
            o = O.Oracle(P.compile_scenario(sc), meta["n_instances"], "f32")
            for _ in range(meta["pre_resets"] + 1):
                o.reset()
            o.train(0, meta["n_iters"] + extra)
            q_tables = o.q.reshape(z["q_final"].shape).copy()
        res = ref_eval.reference_eval(meta["scenario"], q_tables, n_ep, gamma, opt)
        own = res.pop("reference_function_outputs")
        if own:  # deterministic: restated loop == the reference's own test_policy_optima
            for i, o in enumerate(own):
                succ_rate, _mov, avg_t, _std_t, avg_r, _std_r, avg_arps = o
                for k, ag in enumerate(sorted(succ_rate)):
                    assert abs(succ_rate[ag] - 100.0 * res["successes"][i, k] / n_ep) < 1e-9
                    assert abs(avg_r[ag] - res["return_sum"][i, k] / n_ep) < 1e-9
                    assert abs(avg_arps[ag] - res["arps_sum"][i, k] / n_ep) < 1e-9
            print(f"eval_{name}: restated loop == reference test_policy_optima on {len(own)} instance(s)")
        np.savez_compressed(os.path.join(out_dir, "eval_" + name + ".npz"),
                            meta=json.dumps({"train_fixture": name, "n_episodes": n_ep, "gamma": gamma, "optimal_steps": opt,
                                             "generator": "oracle/gen_golden.py gen_eval"}), q_tables=q_tables, **res)
        print(f"eval_{name}: successes={res['successes'].sum()} of {res['episodes'].sum()}")


if __name__ == "__main__" and "--eval" in sys.argv:
    gen_eval()
