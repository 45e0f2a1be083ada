"""TEST INFRASTRUCTURE — golden fixtures for the product-MDP builder: runs the LIVE reference's
RMEnvironmentWrapper.get_mdp (/root/reference/multiagent_rlrm/multi_agent/wrappers/rm_environment_wrapper.py:185-283)
under the in-repo stubs and stores its output as flat arrays in tests/golden/mdp_<name>.npz.

    python oracle/gen_mdp_golden.py

Per agent k the fixture holds count_k [S,4] (entries of P[s][a]) and prob_k / next_k / reward_k / done_k [S,4,4]
(padded with zeros past count). Also recorded: env.stochastic after the call (the reference switches it off and
leaves it off, :196-197).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(_HERE))

import multiagent_rlrm_b200 as P  # noqa: E402
import ref_harness as H  # noqa: E402
from multiagent_rlrm_b200 import tables as T  # noqa: E402
from multiagent_rlrm_b200.maps import office_world_grid  # noqa: E402


def mdp_scenarios():
    S = {}
    S["office_acbd_det"] = P.scenario_config2(False)
    S["office_acbd_stochastic_flag"] = P.scenario_config2(True)  # get_mdp forces the deterministic action map

    sc = P.scenario_config4()
    sc.starts = sc.starts[:2]
    sc.delay_action, sc.wall_penalty, sc.terminate_hit_walls, sc.terminate_on_plants = True, -1.5, True, True
    sc.reward_modifier = 0.5
    S["office_chain12_delay_wallterm_plantterm"] = sc

    sc = P.scenario_config2(False)
    sc.wall_penalty, sc.plants_penalty = -2, -7
    sc.starts = [(2, 7), (0, 0)]
    g = office_world_grid("map1").goals
    sc.rm_transitions_per_agent = [T.office_acbd_transitions(),
                                   [("s0", g["A"], "s1", 1.0), ("s1", g["B"], "s2", 2.5), ("s2", g["C"], "s3", 3.0)]]
    sc.rm_transitions = sc.rm_transitions_per_agent[0]
    S["office_per_agent_rm_wallpen"] = sc

    S["frozen_lake_cfg1"] = P.scenario_config1()  # the reference yields EMPTY outcome lists here
    sc = P.scenario_config3(True)
    sc.penalty_amount, sc.delay_action = -3, True
    S["frozen_lake_slip_delay_penalty"] = sc
    return S


def flatten(P_agent, n_states):
    count = np.zeros((n_states, 4), dtype=np.int32)
    prob = np.zeros((n_states, 4, 4), dtype=np.float64)
    nxt = np.zeros((n_states, 4, 4), dtype=np.int32)
    rew = np.zeros((n_states, 4, 4), dtype=np.float64)
    done = np.zeros((n_states, 4, 4), dtype=np.uint8)
    for s in range(n_states):
        for a in range(4):
            entries = P_agent[s][a]
            count[s, a] = len(entries)
            for j, (p, sn, r, d) in enumerate(entries):
                prob[s, a, j], nxt[s, a, j], rew[s, a, j], done[s, a, j] = p, sn, r, bool(d)
    return count, prob, nxt, rew, done


def main():
    out_dir = os.path.join(os.path.dirname(_HERE), "tests", "golden")
    for name, sc in mdp_scenarios().items():
        d = sc.to_dict()
        rm_env, env, agents = H.build_reference(d, np.float32)
        all_P, n_states, n_actions = rm_env.get_mdp(123)
        arrays = {}
        for k, ag in enumerate(agents):
            assert n_actions[ag.name] == 4
            c, p, n, r, dn = flatten(all_P[ag.name], n_states[ag.name])
            arrays.update({f"count_{k}": c, f"prob_{k}": p, f"next_{k}": n, f"reward_{k}": r, f"done_{k}": dn})
        if sc.env == "office_world":
            # the reference's own value iteration on this model (mdp_vi.py:9-60), both stopping rules: travels to the GPU box
            # as the tolerance anchor for the device sweeps
            import contextlib
            import io

            from multiagent_rlrm.environments.utils_envs.mdp_vi import value_iteration

            for k, ag in enumerate(agents):
                for tag, rel in (("abs", False), ("rel", True)):
                    with contextlib.redirect_stdout(io.StringIO()):
                        V, pol, Q = value_iteration(all_P[ag.name], n_states[ag.name], 4, gamma=0.9, theta=1e-4, delta_rel=rel)
                    arrays.update({f"vi_{tag}_V_{k}": V, f"vi_{tag}_policy_{k}": np.asarray(pol, dtype=np.int32), f"vi_{tag}_Q_{k}": Q})
        meta = {"scenario": d, "seed": 123, "vi": {"gamma": 0.9, "theta": 1e-4}, "n_states": [int(n_states[a.name]) for a in agents],
                "stochastic_after": getattr(env, "stochastic", None), "generator": "oracle/gen_mdp_golden.py"}
        np.savez_compressed(os.path.join(out_dir, "mdp_" + name + ".npz"), meta=json.dumps(meta), **arrays)
        tot = sum(int(arrays[f"count_{k}"].sum()) for k in range(len(agents)))
        print(f"mdp_{name}: agents={len(agents)} states={meta['n_states']} entries={tot} stochastic_after={meta['stochastic_after']}")


if __name__ == "__main__":
    main()
