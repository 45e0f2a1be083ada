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
