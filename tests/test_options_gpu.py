"""GPU: less-travelled options of the C ABI against the oracle (bit-exact): greedy rollouts (learn = 0), visit counting
with a fixed learning rate, hyper-parameter changes between launches (rlrm_set_learner), a state without statistics, the
shared learner on OfficeWorld / plain QL (generic accumulation path), evaluation with per-agent reward machines."""
import numpy as np
import pytest

import multiagent_rlrm_b200 as P

pytestmark = pytest.mark.gpu


def _pair(sc, n, **kw):
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine

    c = P.compile_scenario(sc)
    eng = Engine(c, n, **kw)
    o = O.Oracle(c, n, "f32", track_visits=kw.get("track_visits", False))
    eng.reset(); o.reset()
    return eng, o


def _same(eng, o):
    assert np.array_equal(eng.slot.cpu().numpy().view(np.uint64), o.slot)
    assert np.array_equal(eng.q.cpu().numpy().reshape(-1), o.q.reshape(-1))
    assert np.array_equal(eng.epsilon.cpu().numpy(), o.epsilon)


@pytest.mark.parametrize("name", ["cfg3_qrm", "cfg3_ql", "cfg2", "chain12_qrm"])
def test_greedy_rollout_without_learning(name, cuda_device):
    sc = {"cfg3_qrm": lambda: P.scenario_config3(True), "cfg3_ql": lambda: P.scenario_config3(False),
          "cfg2": lambda: P.scenario_config2(True)}.get(name)
    if sc is None:
        sc = P.scenario_config4()
        sc.algo, sc.learning_rate, sc.q_init = "qrm", 0.1, 2.0
    else:
        sc = sc()
    eng, o = _pair(sc, 300)
    eng.train(400); o.train(0, 400)
    q_before = eng.q.clone()
    tr_g = eng.train(500, learn=False, trace=True)
    tr_o = o.train(400, 500, learn=False, trace=True)
    assert np.array_equal(tr_g.cpu().numpy().view(np.uint32), tr_o)
    assert bool((eng.q == q_before).all())  # nothing learns
    _same(eng, o)


def test_visit_counts_with_fixed_learning_rate(cuda_device):
    eng, o = _pair(P.scenario_config3(True), 200, track_visits=True)
    eng.train(600); o.train(0, 600)
    _same(eng, o)
    assert np.array_equal(eng.visits.cpu().numpy().view(np.uint32).reshape(-1), o.visits.reshape(-1))
    assert int(o.visits.sum()) > 0


def test_hyper_parameters_can_change_between_launches(cuda_device):
    sc = P.scenario_config3(False)
    eng, o = _pair(sc, 256)
    eng.train(300); o.train(0, 300)
    eng.set_learner(0.5, 0.8)
    o.cfg.learning_rate, o.cfg.gamma = 0.5, 0.8
    eng.train(300); o.train(300, 300)
    _same(eng, o)


def test_state_without_statistics(cuda_device):
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine

    c = P.compile_scenario(P.scenario_config3(True))
    eng = Engine(c, 128, with_stats=False)
    o = O.Oracle(c, 128, "f32")
    eng.reset(); o.reset()
    eng.train(700); o.train(0, 700)
    _same(eng, o)


@pytest.mark.parametrize("name", ["office_qrm", "frozen_ql"])
def test_shared_learner_generic_accumulation(name, cuda_device):
    if name == "office_qrm":  # 12-state machine: tables too large for the shared-memory fast path
        sc = P.scenario_config4()
        sc.algo, sc.learning_rate, sc.q_init, sc.shared_q = "qrm", 0.1, 2.0, True
    else:
        sc = P.scenario_config5(True)
        sc.algo, sc.learning_rate = "ql", 0.1
    eng, o = _pair(sc, 700)
    tr_g = eng.train(90, trace=True)
    tr_o = o.train(0, 90, trace=True)
    assert np.array_equal(tr_g.cpu().numpy().view(np.uint32), tr_o)
    _same(eng, o)
    s = eng.stats_numpy()
    for f in ("episodes", "successes", "active_steps", "return_sum"):
        assert np.array_equal(s[f], o.stats[f]), f


def test_evaluation_with_per_agent_reward_machines(cuda_device):
    g = P.frozen_lake_grid("map1").goals
    sc = P.scenario_config1()
    sc.starts = [(5, 0), (0, 0)]
    sc.detector_positions = sorted(g.values())
    sc.rm_transitions_per_agent = [P.tables.frozen_lake_abc_transitions(), [("p0", g["B"], "p1", 3.0), ("p1", g["A"], "p2", 7.0)]]
    eng, o = _pair(sc, 128)
    eng.train(5000); o.train(0, 5000)
    _same(eng, o)
    ev_g, ev_o = eng.evaluate(2, 0.99, 21.0, t0=0), o.evaluate(2, 0.99, 21.0, t0=0)
    for f in ("episodes", "successes", "len_sum", "return_sum", "arps_sum"):
        assert np.array_equal(ev_g[f], ev_o[f]), f
    assert np.array_equal(eng.agent_table(1).cpu().numpy().reshape(-1), o.q.reshape(128, -1, 4)[:, 400:, :].reshape(-1))


@pytest.mark.parametrize("name", ["cfg3_qrm", "cfg4_sparse", "cfg4_dense", "cfg5_shared", "lr_none"])
def test_checkpoint_resume_is_bit_identical(name, cuda_device, tmp_path):
    """train 300 + save + load into a fresh engine + train 300 == train 600 (state_dict carries every state array and t)."""
    import torch

    from multiagent_rlrm_b200.engine import Engine

    kw = {}
    if name == "cfg3_qrm":
        sc = P.scenario_config3(True)
    elif name in ("cfg4_sparse", "cfg4_dense"):
        sc = P.scenario_config4()
        kw = {"qlambda_sparse": name == "cfg4_sparse"}
    elif name == "cfg5_shared":
        sc = P.scenario_config5(True)
    else:
        sc = P.scenario_config3(False)
        sc.learning_rate = None
    c = P.compile_scenario(sc)
    straight, first = Engine(c, 96, **kw), Engine(c, 96, **kw)
    straight.reset(); first.reset()
    straight.train(600)
    first.train(300)
    path = tmp_path / "engine.pt"
    torch.save(first.state_dict(), path)
    resumed = Engine(c, 96, **kw)
    resumed.load_state_dict(torch.load(path))
    resumed.train(300)
    assert resumed.t == straight.t == 600
    for k in Engine._STATE_TENSORS:
        a, b = getattr(straight, k), getattr(resumed, k)
        assert (a is None) == (b is None), k
        if a is not None and k != "tr_work":
            if k in ("tr_idx", "tr_eq", "tr_pos", "acc_last"):
                continue  # scratch: list storage past tr_len (compared through the materialised tables below); acc_last is
                          # only meaningful while acc_cnt == 1 and holds whichever proposal was stored last otherwise
            assert torch.equal(a, b), k
    ea, eb = straight.sync_tables(with_traces=True), resumed.sync_tables(with_traces=True)
    assert torch.equal(straight.q, resumed.q)
    if ea is not None:
        assert torch.equal(ea, eb)
    with pytest.raises(ValueError):
        Engine(c, 95, **kw).load_state_dict(first.state_dict())


@pytest.mark.parametrize("name,n,kw", [("cfg5_tables", 300000, {}), ("cfg4_sparse", 270000, {"qlambda_sparse": True}),
                                        ("cfg3_f64", 600000, {}), ("fl_per_agent", 400000, {})])
def test_pipelined_train_host_equals_train(name, n, kw, cuda_device):
    """Above 1 Mi slots rlrm_train_host splits the instance range into chunks and pipelines upload / kernel / download over two
    extra streams (rlrm_b200.cu sub_state): every array view, the Philox instance offset and the host buffers must line up so
    that the result is bit-identical to one resident launch — for float32 / float64 tables, sparse trace lists and per-agent
    table sections."""
    import torch

    from multiagent_rlrm_b200.engine import Engine

    free, _total = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~50 GB of free device memory")
    if name == "cfg5_tables":
        sc = P.scenario_config5(False)
    elif name == "cfg4_sparse":
        sc = P.scenario_config4()
    elif name == "cfg3_f64":
        sc = P.scenario_config3(False)
        sc.table_dtype = "f64"
    else:
        g = P.frozen_lake_grid("map1").goals
        sc = P.scenario_config3(True)
        sc.starts, sc.detector_positions = [(5, 0), (0, 0), (9, 9)], sorted(g.values())
        sc.rm_transitions_per_agent = [P.tables.frozen_lake_abc_transitions(), [("p0", g["C"], "p1", 3.0), ("p1", g["A"], "p2", 7.0)],
                                       [("w0", g["B"], "w1", 1.0), ("w1", g["A"], "w2", 1.0), ("w2", g["C"], "w3", 1.0)]]
    c = P.compile_scenario(sc, instance_offset=12345)
    resident, hosted = Engine(c, n, **kw), Engine(c, n, **kw)
    resident.reset(); hosted.reset()
    n_slots = n * c.n_agents
    assert n_slots >= 2 * 512 * 1024  # at least two chunks
    host_slot = torch.empty(n_slots, dtype=torch.int64).pin_memory()
    host_eps = torch.empty(n_slots, dtype=torch.float64).pin_memory()
    host_stats = torch.empty((n_slots, 32), dtype=torch.uint8).pin_memory()
    host_slot.copy_(hosted.slot); host_eps.copy_(hosted.epsilon)
    for chunk in (3, 40):
        resident.train(chunk)
        hosted.train_host(chunk, host_stats, host_slot, host_eps)
    resident.sync_tables(); hosted.sync_tables()
    assert torch.equal(hosted.slot, resident.slot) and torch.equal(hosted.epsilon, resident.epsilon)
    assert torch.equal(hosted.q, resident.q) and torch.equal(hosted.stats, resident.stats) and torch.equal(hosted.ep_return, resident.ep_return)
    assert torch.equal(host_slot, resident.slot.cpu()) and torch.equal(host_eps, resident.epsilon.cpu()) and torch.equal(host_stats, resident.stats.cpu())
    if hosted.sparse:
        assert torch.equal(hosted.tr_len, resident.tr_len) and torch.equal(hosted.tr_work, resident.tr_work)
    del resident, hosted
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["cfg3_qrm", "cfg2_ql", "cfg4_sparse"])
def test_train_host_equals_train(name, cuda_device):
    """rlrm_train_host (the end-to-end entry point bench.py's `e2e` times: host slot / epsilon in, fused iterations, host slot /
    epsilon / statistics out) == rlrm_train on resident state, and the host buffers hold exactly what the device holds."""
    import torch

    from multiagent_rlrm_b200.engine import Engine

    kw = {}
    if name == "cfg3_qrm":
        sc = P.scenario_config3(True)
    elif name == "cfg2_ql":
        sc = P.scenario_config2(True)
    else:
        sc, kw = P.scenario_config4(), {"qlambda_sparse": True}
    c = P.compile_scenario(sc)
    resident, hosted = Engine(c, 200, **kw), Engine(c, 200, **kw)
    resident.reset(); hosted.reset()
    n_slots = 200 * c.n_agents
    host_slot = torch.empty(n_slots, dtype=torch.int64).pin_memory()
    host_eps = torch.empty(n_slots, dtype=torch.float64).pin_memory()
    host_stats = torch.empty((n_slots, 32), dtype=torch.uint8).pin_memory()
    host_slot.copy_(hosted.slot); host_eps.copy_(hosted.epsilon)
    for chunk in (37, 300, 163):
        resident.train(chunk)
        hosted.train_host(chunk, host_stats, host_slot, host_eps)
    assert hosted.t == resident.t == 500
    assert torch.equal(hosted.slot, resident.slot) and torch.equal(hosted.epsilon, resident.epsilon)
    resident.sync_tables(); hosted.sync_tables()
    assert torch.equal(hosted.q, resident.q) and torch.equal(hosted.stats, resident.stats)
    assert torch.equal(host_slot, resident.slot.cpu()) and torch.equal(host_eps, resident.epsilon.cpu())
    assert torch.equal(host_stats, resident.stats.cpu())
    # the host copies are INPUTS too: editing them before the call changes what the device runs on
    host_eps.fill_(1.0)
    hosted.train_host(5, host_stats, host_slot, host_eps)
    assert float(hosted.epsilon.max()) <= 1.0 and float(hosted.epsilon.min()) >= float(sc.epsilon_end)
    assert not torch.equal(hosted.slot, resident.slot)
