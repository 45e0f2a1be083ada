"""Product-MDP builder (RMEnvironmentWrapper.get_mdp, rm_environment_wrapper.py:185-283).

Goldens (tests/golden/mdp_*.npz) are the LIVE reference's get_mdp output (oracle/gen_mdp_golden.py). CPU: the oracle's
restatement against them. GPU: rlrm_mdp against the oracle, and the drop-in wrapper's get_mdp (one launch per agent)
against the reference's dictionaries entry by entry."""
import glob
import json
import os

import numpy as np
import pytest

import multiagent_rlrm_b200 as P
from multiagent_rlrm_b200.wrapper import assemble_mdp

from dropin_builder import build_b200

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "mdp_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"mdp_{name}.npz"))
    return z, json.loads(str(z["meta"]))


def golden_P(z, k):
    count, prob, nxt, rew, done = (z[f"{f}_{k}"] for f in ("count", "prob", "next", "reward", "done"))
    return {s: {a: [(float(prob[s, a, j]), int(nxt[s, a, j]), float(rew[s, a, j]), bool(done[s, a, j]))
                    for j in range(count[s, a])] for a in range(4)} for s in range(count.shape[0])}


def distribution(sc):
    """Sub-action indices [4][n_sub] + per-action probabilities, as get_mdp obtains them from the environment."""
    from multiagent_rlrm_b200.tables import mdp_action_distribution

    sub, probs = mdp_action_distribution(P.Scenario.from_dict(sc))
    return sub, [list(probs)] * 4


def same_P(mine, ref):
    assert set(mine) == set(ref)
    for s in ref:
        for a in range(4):
            assert len(mine[s][a]) == len(ref[s][a]), (s, a)
            for m, r in zip(mine[s][a], ref[s][a]):
                assert m[0] == r[0] and m[1] == r[1] and m[2] == r[2] and m[3] == r[3], (s, a, m, r)


def test_fixture_inventory():
    assert len(NAMES) == 6
    z, meta = load("frozen_lake_cfg1")  # the reference's FrozenLake get_mdp only ever fills in the hole states
    assert int(z["count_0"].sum()) == 11 * 4 * 4  # 11 holes x 4 RM states x 4 actions, one self-loop each
    z, meta = load("office_acbd_stochastic_flag")
    assert meta["scenario"]["stochastic"] and meta["stochastic_after"] is False  # side effect of get_mdp


@pytest.mark.parametrize("name", NAMES)
def test_oracle_mdp_equals_reference(name):
    import oracle as O

    z, meta = load(name)
    sc = P.Scenario.from_dict(meta["scenario"])
    o = O.Oracle(P.compile_scenario(sc), 1, "f32")
    sub, probs = distribution(meta["scenario"])
    fl = sc.env == "frozen_lake"
    for k in range(len(sc.starts)):
        nxt, rew, done, term = o.mdp(k, sub, rm_terminal=not fl)
        assert nxt.shape[0] == meta["n_states"][k]
        same_P(assemble_mdp(nxt, rew, done, term, probs, empty_nonterminal=fl), golden_P(z, k))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_mdp_equals_oracle(name):
    import oracle as O

    z, meta = load(name)
    sc = P.Scenario.from_dict(meta["scenario"])
    c = P.compile_scenario(sc)
    from multiagent_rlrm_b200.engine import Engine

    eng, o = Engine(c, 1), O.Oracle(c, 1, "f32")
    sub, _probs = distribution(meta["scenario"])
    for k in range(len(sc.starts)):
        for rm_terminal in (True, False):
            got, exp = eng.mdp(k, sub, rm_terminal=rm_terminal), o.mdp(k, sub, rm_terminal=rm_terminal)
            for g, e in zip(got, exp):
                assert np.array_equal(g, e)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_dropin_get_mdp_equals_reference(name):
    z, meta = load(name)
    rm_env, env, agents = build_b200(meta["scenario"])
    rm_env.reset(7)
    agents[0].set_position(1, 1)
    all_P, n_states, n_actions = rm_env.get_mdp(meta["seed"])
    from multiagent_rlrm_b200.envs import _A2I

    sub, probs = distribution(meta["scenario"])  # host helper == what the environment object reports
    dists = [env.get_action_distribution(a) for a in agents[0].get_actions()]
    assert [[_A2I[x.name] for x in d[0]] for d in dists] == sub and [list(d[1]) for d in dists] == probs
    assert getattr(env, "stochastic", None) == meta["stochastic_after"]
    for k, ag in enumerate(agents):
        assert n_states[ag.name] == meta["n_states"][k] and n_actions[ag.name] == 4
        same_P(all_P[ag.name], golden_P(z, k))
        assert tuple(ag.get_position()) == tuple(meta["scenario"]["starts"][k])  # ends freshly reset (:282)
        assert ag.get_reward_machine().get_current_state() == ag.get_reward_machine().initial_state


@pytest.mark.gpu
def test_repaired_frozen_lake_mdp_is_a_proper_transition_model():
    """Extension (not reference behaviour): with repaired=True FrozenLake gets the MDP OfficeWorld gets."""
    _z, meta = load("frozen_lake_slip_delay_penalty")
    rm_env, env, agents = build_b200(meta["scenario"])
    all_P, n_states, _ = rm_env.get_mdp(1, repaired=True)
    Pa = all_P[agents[0].name]
    rm = agents[0].get_reward_machine()
    nQ, W = rm.numbers_state(), env.grid_width
    final = rm.get_state_index(rm.get_final_state())
    for s, by_action in Pa.items():
        cell, q = divmod(s, nQ)
        hole = (cell % W, cell // W) in env.holes
        for a, entries in by_action.items():
            assert abs(sum(e[0] for e in entries) - 1.0) < 1e-12
            if hole or q == final:
                assert entries == [(1.0, s, env.penalty_amount if hole else 0, True)]
            else:
                assert len(entries) == 4 and entries[0][1] // nQ == cell  # delay mode: first sub-action is "wait"


@pytest.mark.gpu
def test_wait_action_through_the_env_api():
    """env.wait_action is a legal action of the reference's OfficeWorld step (ma_office.py:299-300)."""
    _z, meta = load("office_acbd_det")
    rm_env, env, agents = build_b200(meta["scenario"])
    rm_env.reset(1)
    before = tuple(agents[0].get_position())
    _obs, rewards, term, trunc, _infos = rm_env.step({agents[0].name: env.wait_action})
    assert tuple(agents[0].get_position()) == before and rewards[agents[0].name] == 0
    assert not term[agents[0].name] and not trunc[agents[0].name] and env.timestep == 1
