"""Greedy evaluation (SURVEY.md §8 f2 = test_policy_optima, evaluation_metrics.py:23-190): oracle vs the reference-generated
fixtures on CPU; CUDA (rlrm_evaluate) vs both on the GPU."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

EVAL_FIXTURES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("eval_") and f.endswith(".npz"))
INT_FIELDS = ("episodes", "successes", "len_sum", "len_sqsum")
F64_FIELDS = ("return_sum", "return_sqsum", "arps_sum")


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    train_meta, _ = load_golden(meta["train_fixture"])
    return meta, train_meta, {k: z[k] for k in z.files if k != "meta"}


def _check(ev, ref, n, a, name):
    for f in INT_FIELDS:
        assert np.array_equal(ev[f].reshape(n, a).astype(np.int64), ref[f]), f"{name}: {f}"
    for f in F64_FIELDS:  # same operations in the same order as the Python loop: bit-identical doubles
        assert np.array_equal(ev[f].reshape(n, a), ref[f]), f"{name}: {f}"


@pytest.mark.parametrize("name", EVAL_FIXTURES)
def test_oracle_evaluation_matches_reference(name):
    import multiagent_rlrm_b200 as P
    import oracle as O

    meta, tm, ref = _load(name)
    c = P.compile_scenario(P.Scenario.from_dict(tm["scenario"]))
    n = ref["q_tables"].shape[0]
    o = O.Oracle(c, n, "f32")
    o.q[...] = ref["q_tables"].reshape(o.q.shape)
    ev = o.evaluate(meta["n_episodes"], meta["gamma"], meta["optimal_steps"])
    _check(ev, ref, n, c.n_agents, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", EVAL_FIXTURES)
def test_cuda_evaluation_matches_reference(name, cuda_device):
    import torch

    import multiagent_rlrm_b200 as P
    from multiagent_rlrm_b200.engine import Engine

    meta, tm, ref = _load(name)
    c = P.compile_scenario(P.Scenario.from_dict(tm["scenario"]))
    n = ref["q_tables"].shape[0]
    eng = Engine(c, n)
    eng.q.copy_(torch.from_numpy(ref["q_tables"].reshape(tuple(eng.q.shape))))
    before = (eng.slot.clone(), eng.epsilon.clone())
    ev = eng.evaluate(meta["n_episodes"], meta["gamma"], meta["optimal_steps"], t0=0)
    _check(ev, ref, n, c.n_agents, name)
    assert torch.equal(eng.slot, before[0]) and torch.equal(eng.epsilon, before[1])  # training state untouched


@pytest.mark.gpu
def test_cuda_evaluation_matches_oracle_after_training(cuda_device):
    """Train on the GPU, evaluate both ways: exercises trained tables with a mix of successes and failures."""
    import multiagent_rlrm_b200 as P
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine
    from multiagent_rlrm_b200.evaluation import test_policy_optima_batched

    sc = P.scenario_config1()
    c = P.compile_scenario(sc)
    n = 512
    eng = Engine(c, n)
    eng.reset()
    eng.train(3500)  # partly trained tables
    o = O.Oracle(c, n, "f32")
    o.q[...] = eng.q.cpu().numpy()
    ev_g = eng.evaluate(3, 0.99, 21.0, t0=0)
    ev_o = o.evaluate(3, 0.99, 21.0, t0=0)
    for f in INT_FIELDS + F64_FIELDS:
        assert np.array_equal(ev_g[f], ev_o[f]), f
    assert ev_g["episodes"].sum() == n * 2 * 3 and (ev_g["successes"] <= ev_g["episodes"]).all()
    res = test_policy_optima_batched(eng, episodi_test=3, optimal_steps=21.0, gamma=0.99)
    assert res["success_rate"].shape == (n, 2) and np.all(res["success_rate"] <= 100.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eval_cfg1_det_qrm", "eval_cfg2_office_det_ql"])
def test_explicit_policy_evaluation_matches_reference(name, cuda_device):
    """test_policy_opt_multi_batched vs the live reference's test_policy_opt_multi (fixture values) on the greedy policy
    of the fixture tables, deterministic dynamics."""
    import torch

    import multiagent_rlrm_b200 as P
    from multiagent_rlrm_b200.engine import Engine
    from multiagent_rlrm_b200.evaluation import extract_policy_from_qtable, test_policy_opt_multi_batched

    meta, tm, ref = _load(name)
    c = P.compile_scenario(P.Scenario.from_dict(tm["scenario"]))
    n, A = ref["q_tables"].shape[0], c.n_agents
    eng = Engine(c, n)
    eng.q.copy_(torch.from_numpy(ref["q_tables"].reshape(tuple(eng.q.shape))))
    policies = extract_policy_from_qtable(eng.q.view(n, A, eng.S, 4))
    res = test_policy_opt_multi_batched(eng, policies, episodes_test=meta["n_episodes"], optimal_steps=meta["optimal_steps"],
                                        gamma=meta["gamma"])
    assert np.array_equal(res["success_rate"], ref["optmulti_success_rate"])
    assert np.array_equal(res["avg_timesteps"], ref["optmulti_avg_timesteps"])
    assert np.allclose(res["avg_reward"], ref["optmulti_avg_reward"], rtol=1e-12, atol=0)
    assert np.allclose(res["avg_arps"], ref["optmulti_avg_arps"], rtol=1e-12, atol=0)
    one = test_policy_opt_multi_batched(eng, policies[0], episodes_test=5, optimal_steps=meta["optimal_steps"], gamma=meta["gamma"],
                                        test_deterministic=True)  # [A, S] policy for every instance; a single episode
    assert (one["episodes"] == 1).all()
    with pytest.raises(ValueError):
        test_policy_opt_multi_batched(eng, policies + 4)


@pytest.mark.gpu
def test_explicit_policy_evaluation_stops_one_step_before_truncation(cuda_device):
    """The reference loop runs while timestep < 1000, i.e. a failing episode is 1000 steps long, not the 1001 the
    environment's own truncation gives test_policy_optima: visible in ARPS = reward / timestep / optimal."""
    import torch

    import multiagent_rlrm_b200 as P
    from multiagent_rlrm_b200.engine import Engine
    from multiagent_rlrm_b200.evaluation import test_policy_opt_multi_batched

    sc = P.scenario_config2(False)
    sc.wall_penalty = -1.0  # bumping into the boundary forever: reward -1 per step, never terminates
    c = P.compile_scenario(sc)
    eng = Engine(c, 2)
    policy = np.full((1, eng.S), 2, dtype=np.int64)  # always "left" from (2, 7): two moves, then the wall
    res = test_policy_opt_multi_batched(eng, policy, episodes_test=1, optimal_steps=1.0, gamma=1.0)
    assert (res["success_rate"] == 0).all()
    assert np.allclose(res["avg_reward"], -998.0) and np.allclose(res["avg_arps"], -998.0 / 1000.0)
