"""GPU: the round-2 entry points of the C ABI — rlrm_iterate (one driver-loop iteration per launch, record written into
page-locked host memory), the counterfactual outputs of rlrm_step, rlrm_update_list, rlrm_merge_replicas — against the
fused kernel, the call-by-call path and the CPU oracle. Every comparison is bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(compiled, n, **kw):
    from multiagent_rlrm_b200.engine import Engine

    return Engine(compiled, n, device="cuda:0", **kw)


def _scenarios():
    import multiagent_rlrm_b200 as P

    sc4 = P.scenario_config4()
    sc4q = P.scenario_config4()
    sc4q.algo, sc4q.learning_rate, sc4q.q_init = "qrm", 0.1, 2.0
    return {"cfg3_qrm": P.scenario_config3(True), "cfg3_ql": P.scenario_config3(False), "cfg2_slip": P.scenario_config2(True),
            "cfg4_qlambda": sc4, "cfg4_qrm": sc4q, "cfg5": P.scenario_config5(False)}


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("name", ["cfg3_qrm", "cfg3_ql", "cfg2_slip", "cfg4_qlambda", "cfg4_qrm", "cfg5"])
def test_iterate_equals_fused_train_and_call_by_call(name, dtype, cuda_device):
    """rlrm_iterate, iteration by iteration with a host synchronisation in between, == rlrm_train over the same iterations
    (trace rows, tables, slot words, epsilon) and its per-step rewards == the rewards of the call-by-call rlrm_step."""
    import multiagent_rlrm_b200 as P

    sc = _scenarios()[name]
    sc.table_dtype = dtype
    c = P.compile_scenario(sc)
    n, iters = (24, 60) if name == "cfg4_qlambda" else (160, 260)
    a, b, u = _engine(c, n), _engine(c, n), _engine(c, n)
    for e in (a, b, u):
        e.reset()
    trace = a.train(iters, trace=True).cpu().numpy()
    for it in range(iters):
        rec, rew = b.iterate()
        assert not rec.is_cuda and rec.is_pinned() and not rew.is_cuda      # written in place into page-locked host memory
        assert np.array_equal(rec.numpy(), trace[it]), f"record of iteration {it}"
        _actions, urec, _over = u.iterate_unfused()
        assert np.array_equal(rew.numpy(), urec["reward"].cpu().numpy()), f"rewards of iteration {it}"
    assert np.array_equal(a.q.cpu().numpy(), b.q.cpu().numpy()) and np.array_equal(a.slot.cpu().numpy(), b.slot.cpu().numpy())
    assert np.array_equal(a.epsilon.cpu().numpy(), b.epsilon.cpu().numpy())
    assert np.array_equal(a.stats.cpu().numpy(), b.stats.cpu().numpy())
    if a.e is not None:
        assert np.array_equal(a.e.cpu().numpy(), b.e.cpu().numpy())
    f = b.unpack_record(rec)
    assert int(f["cell"].max()) < c.config.width * c.config.height and int(f["action"].max()) < 4


def test_iterate_sparse_qlambda_and_shared_learner(cuda_device):
    import multiagent_rlrm_b200 as P

    c = P.compile_scenario(P.scenario_config4())
    a, b = _engine(c, 16, qlambda_sparse=True), _engine(c, 16, qlambda_sparse=True)
    a.reset(); b.reset()
    trace = a.train(80, trace=True).cpu().numpy()
    for it in range(80):
        rec, _rew = b.iterate()
        assert np.array_equal(rec.numpy(), trace[it])
    a.sync_tables(); b.sync_tables()
    assert np.array_equal(a.q.cpu().numpy(), b.q.cpu().numpy())
    c = P.compile_scenario(P.scenario_config5(True))
    a, b = _engine(c, 500), _engine(c, 500)
    a.reset(); b.reset()
    trace = a.train(50, trace=True).cpu().numpy()
    for it in range(50):
        rec, rew = b.iterate(want_reward=False)
        assert rew is None and np.array_equal(rec.numpy(), trace[it])
    assert np.array_equal(a.q.cpu().numpy(), b.q.cpu().numpy())


@pytest.mark.parametrize("name", ["cfg3_qrm", "cfg4_qrm", "per_agent"])
def test_step_host_record_and_counterfactual_lookups_match_oracle(name, cuda_device):
    """rlrm_step driven from page-locked host memory (Engine.step_host, the one-instance bridge of envs.py) with the QRM
    counterfactual outputs cf_q / cf_r == the oracle's step record and lookups."""
    import multiagent_rlrm_b200 as P
    import oracle as O
    import philox

    if name == "per_agent":
        g = P.frozen_lake_grid("map1").goals
        sc = P.scenario_config3(True)
        sc.starts, sc.detector_positions = [(5, 0), (0, 0), (9, 9)], sorted(g.values())
        sc.rm_transitions_per_agent = [P.tables.frozen_lake_abc_transitions(), [("p0", g["C"], "p1", 3.0), ("p1", g["A"], "p2", 7.0)],
                                       [("w0", g["B"], "w1", 1.0), ("w1", g["A"], "w2", 1.0), ("w2", g["C"], "w3", 1.0),
                                        ("w3", g["B"], "w4", 5.0), ("w1", g["C"], "w0", -1.0)]]
    else:
        sc = _scenarios()[name]
    c = P.compile_scenario(sc)
    n = 3
    eng, o = _engine(c, n, host_control=True, with_stats=False), O.Oracle(c, n, "f32")
    eng.reset(); o.reset()
    eng.sync()
    rng = np.random.default_rng(7)
    for it in range(300):
        acts = rng.integers(0, 4, size=n * c.n_agents).astype(np.uint8)
        d = philox.draws(sc.seed, it, n, c.n_agents).astype(np.uint32).reshape(-1)
        got = eng.step_host(acts, d, with_rm=True, counterfactuals=True)
        want = o.step(acts, t=it, draws=d, counterfactuals=True)
        for k, v in want.items():
            assert np.array_equal(np.asarray(got[k]).view(v.dtype), v), f"{k} at step {it}"
        assert np.array_equal(eng.slot.numpy().view(np.uint64), o.slot)
        over = (want["term"].reshape(n, -1).all(axis=1) | want["trunc"].reshape(n, -1).all(axis=1)).astype(np.uint8)
        if over.any():
            import torch

            eng.reset(torch.from_numpy(over)); o.reset(over)
            eng.sync()


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("lr", [0.1, None])
def test_update_list_equals_reference_update_q(lr, dtype, cuda_device):
    """rlrm_update_list through the drop-in QLearning: a QRM experience list applied in ONE launch == the reference's
    update_q loop (qlearning.py:70-106) evaluated with NumPy on the same table dtype."""
    import multiagent_rlrm_b200 as P

    S, rng = 60, np.random.default_rng(3)
    npdt = np.float32 if dtype == "f32" else np.float64
    L = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=lr, state_space_size=S, action_space_size=4, qtable_init=2,
                    use_qrm=True, table_dtype=dtype)
    q = np.full((S, 4), 2, dtype=npdt)
    visits = np.zeros((S, 4))
    for _ in range(40):
        exps = []
        for _k in range(int(rng.integers(1, 40))):   # more than one 16-entry block per call
            s, sn, a = int(rng.integers(0, S)), int(rng.integers(0, S)), int(rng.integers(0, 4))
            r, done = float(rng.choice([0, 1, -1.5, 10])), bool(rng.integers(0, 2))
            exps.append((s, a, r, sn, done, 0, 0, 0, 0, 0.0))
            cur = q[s, a]
            visits[s, a] += 1
            rate = 1 / visits[s, a] if lr is None else lr
            mx = (not done) * np.max(q[sn])
            q[s, a] = (1 - rate) * cur + rate * (r + 0.9 * mx)
        L.update(0, 0, 0, 0.0, False, info={"qrm_experience": exps})
    assert np.array_equal(np.asarray(L.q_table), q)
    assert np.array_equal(np.asarray(L.visits), visits.astype(np.int32))


def test_merge_replicas_kernel_is_the_rank_ordered_mean(cuda_device):
    import ctypes as C

    import torch

    import multiagent_rlrm_b200 as P

    eng = _engine(P.compile_scenario(P.scenario_config5(True)), 4)
    n = eng.q.numel()
    for world in (1, 2, 3, 8):
        g = (torch.randn((world, n), device="cuda:0") * 7).contiguous()
        out = torch.empty(n, device="cuda:0")
        assert eng.L.rlrm_merge_replicas(eng.h, g.data_ptr(), world, n, out.data_ptr(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
        acc = g[0].clone()
        for r in range(1, world):
            acc = acc + g[r]
        want = (acc.double() / world).float() if world in (1, 2, 8) else None  # power of two: exact in any formulation
        if want is not None:
            assert torch.equal(out, want)
        ref = np.float32(1) * acc.cpu().numpy() / np.float32(world)              # correctly rounded float32 division
        assert np.array_equal(out.cpu().numpy(), ref.astype(np.float32))
