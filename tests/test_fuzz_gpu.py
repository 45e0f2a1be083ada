"""GPU: seeded random scenarios, CUDA path vs the oracle, bit-exact on every state array.

The fixed fixtures pin the oracle to the reference; this test spreads the CUDA-vs-oracle comparison over option
COMBINATIONS no fixture has (environment x map x number of agents x machine shape x learner x dynamics flags x
shaping x per-agent machines x random starts x shared learner x visit counts x short step caps that force frequent
truncation and auto-reset). Scenarios are drawn from numpy Generators with fixed seeds, so a failure is reproducible
from its test id."""
import os

import numpy as np
import pytest
import torch

import multiagent_rlrm_b200 as P
from multiagent_rlrm_b200.maps import frozen_lake_grid, office_world_grid

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("RLRM_FUZZ_CASES", "128"))  # raise for a one-off soak run (1024 cases take about a minute on a B200)


def random_machine(rng, cells, tag):
    """A chain of 1..5 rewarded events over distinct cells, plus optional self-loops / shortcuts / resets."""
    k = int(rng.integers(1, 6))
    ev = [cells[i] for i in rng.choice(len(cells), size=k, replace=False)]
    trs = [(f"{tag}{i}", ev[i], f"{tag}{i + 1}", float(rng.choice([0, 1, 2.5, 10, -1]))) for i in range(k)]
    extra = []
    for i in range(k):
        for j in range(k):
            if j != i and rng.random() < 0.25:
                kind = rng.integers(0, 3)
                dst = f"{tag}{i}" if kind == 0 else (f"{tag}0" if kind == 1 else f"{tag}{min(i + 2, k)}")
                extra.append((f"{tag}{i}", ev[j], dst, float(rng.choice([0, -0.5, 1]))))
    # the final state is the target of the LAST inserted transition: keep the chain's last link last
    return trs[:-1] + extra + trs[-1:], ev


def random_scenario(seed):
    rng = np.random.default_rng(1000 + seed)
    env = "frozen_lake" if rng.random() < 0.5 else "office_world"
    if env == "frozen_lake":
        map_name, grid = "map1", frozen_lake_grid("map1")
    else:
        map_name = str(rng.choice(["map0", "map1", "map2", "map3", "map4"]))
        grid = office_world_grid(map_name)
    hazards = set(grid.hazards)
    free = [(x, y) for y in range(grid.height) for x in range(grid.width) if (x, y) not in hazards]
    A = int(rng.integers(1, 5))
    starts = [free[i] for i in rng.choice(len(free), size=A, replace=False)]
    algo = str(rng.choice(["ql", "qrm", "qlambda"]))
    per_agent = A > 1 and rng.random() < 0.25
    if per_agent:
        machines = [random_machine(rng, free, f"m{a}_") for a in range(A)]
        rm, per = machines[0][0], [m[0] for m in machines]
        detector = sorted({p for m in machines for p in m[1]})
    else:
        rm, ev = random_machine(rng, free, "s")
        per = None
        detector = sorted(set(ev) | ({free[int(rng.integers(len(free)))]} if rng.random() < 0.3 else set()))
    sc = P.Scenario(env=env, map_name=map_name, starts=starts, rm_transitions=rm, rm_transitions_per_agent=per,
                    detector_positions=detector, algo=algo, seed=int(rng.integers(1, 1 << 30)))
    sc.driver = "frozen_lake_main" if (env == "frozen_lake") != (rng.random() < 0.15) else "office_main"
    sc.stochastic = bool(rng.random() < 0.7)
    sc.delay_action = bool(sc.stochastic and rng.random() < 0.25)
    sc.all_slip = bool(env == "office_world" and rng.random() < 0.3)
    sc.high_prob = float(rng.choice([0.8, 0.6, 0.95]))
    sc.penalty_amount = float(rng.choice([0, -1, -5.5]))
    sc.plants_penalty = float(rng.choice([-100, -3, 0]))
    sc.wall_penalty = float(rng.choice([0, -0.25, -2]))
    sc.terminate_on_plants = bool(rng.random() < 0.3)
    sc.terminate_hit_walls = bool(rng.random() < 0.2)
    sc.max_steps = int(rng.choice([12, 40, 150, 1000]))
    sc.reward_modifier = float(rng.choice([1, 1, 0.5, 3]))
    sc.gamma = float(rng.choice([0.9, 0.99, 0.5]))
    sc.q_init = float(rng.choice([0.0, 2.0, -1.0]))
    sc.epsilon_start = float(rng.choice([0.01, 0.3, 1.0]))
    sc.epsilon_end = float(rng.choice([0.01, 0.1]))
    sc.epsilon_decay = float(rng.choice([1.0, 0.9, 0.999]))
    if algo == "qlambda":
        sc.learning_rate, sc.lambd = float(rng.choice([0.1, 0.5])), float(rng.choice([0.0, 0.5, 0.9]))
        if rng.random() < 0.25:
            sc.learning_rate = None  # 1 / visits, float64 arithmetic (qlearning_lambda.py:44-49)
    else:
        sc.learning_rate = None if rng.random() < 0.2 else float(rng.choice([1.0, 0.1, 0.37]))
        if rng.random() < 0.25:
            sc.use_rsh, sc.rs_kind = True, str(rng.choice(["vi", "distance"]))
            sc.rs_gamma, sc.rs_alpha = float(rng.choice([0.9, 0.99])), float(rng.choice([100, 7]))
    sc.random_start_positions = bool(env == "frozen_lake" and rng.random() < 0.25)
    sc.shared_q = bool(algo != "qlambda" and sc.learning_rate is not None and not sc.use_rsh and rng.random() < 0.2)
    opts = {"n": int(rng.choice([1, 3, 17, 64, 129])), "iters": int(rng.choice([150, 400, 700])),
            "track_visits": bool(not sc.shared_q and rng.random() < 0.25),
            "sparse": bool(algo == "qlambda" and rng.random() < 0.5), "chunks": int(rng.choice([1, 1, 3]))}
    if not sc.shared_q and rng.random() < 0.3:
        sc.table_dtype = "f64"  # the reference's native float64 tables (RLRM_TABLE_F64)
    return sc, opts


@pytest.mark.parametrize("seed", range(N_CASES))
def test_random_scenario_matches_oracle(seed, cuda_device):
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine

    sc, opts = random_scenario(seed)
    c = P.compile_scenario(sc)
    kw = {"track_visits": True} if opts["track_visits"] else {}
    if opts["sparse"]:
        kw["qlambda_sparse"] = True
    eng = Engine(c, opts["n"], **kw)
    o = O.Oracle(c, opts["n"], sc.table_dtype, track_visits=opts["track_visits"])
    eng.reset(); o.reset()
    done = 0
    for part in np.array_split(np.arange(opts["iters"]), opts["chunks"]):  # launch boundaries must not matter
        eng.train(len(part)); o.train(done, len(part))
        done += len(part)
    info = f"seed {seed}: {sc.env}/{sc.map_name} A={len(sc.starts)} {sc.algo} {opts}"
    assert np.array_equal(eng.slot.cpu().numpy().view(np.uint64), o.slot), info
    assert np.array_equal(eng.epsilon.cpu().numpy(), o.epsilon), info
    e = eng.sync_tables(with_traces=True)  # sparse Q(lambda): writes the cached q values back, returns the dense traces
    assert np.array_equal(eng.q.cpu().numpy().reshape(-1), o.q.reshape(-1)), info
    if sc.algo == "qlambda":
        assert np.array_equal(e.cpu().numpy().reshape(-1), o.e.reshape(-1)), info
    if opts["track_visits"] or sc.learning_rate is None:
        assert np.array_equal(eng.visits.cpu().numpy().view(np.uint32).reshape(-1), o.visits.reshape(-1)), info
    st = eng.stats_numpy()
    for f in ("active_steps", "episodes", "successes", "return_sum", "last_return", "last_length"):
        assert np.array_equal(st[f].reshape(-1), o.stats[f].reshape(-1)), (info, f)
    assert eng.total_active_steps() == o.total_active_steps(), info


@pytest.mark.parametrize("seed", range(0, N_CASES, 4))
def test_random_scenario_call_by_call_equals_fused(seed, cuda_device):
    """select / step / update / reset as four C-ABI calls per iteration == the fused launch, on the same random scenarios."""
    from multiagent_rlrm_b200.engine import Engine

    sc, opts = random_scenario(seed)
    c = P.compile_scenario(sc)
    kw = {"track_visits": True} if opts["track_visits"] else {}
    fused, unfused = Engine(c, opts["n"], **kw), Engine(c, opts["n"], **kw)
    fused.reset(); unfused.reset()
    iters = min(opts["iters"], 160)
    fused.train(iters)
    for _ in range(iters):
        unfused.iterate_unfused()
    info = f"seed {seed}: {sc.env}/{sc.map_name} A={len(sc.starts)} {sc.algo} {opts}"
    assert bool((fused.slot == unfused.slot).all()), info
    assert bool((fused.epsilon == unfused.epsilon).all()), info
    assert bool((fused.q.view(torch.int32) == unfused.q.view(torch.int32)).all()), info
    if sc.algo == "qlambda":
        assert bool((fused.e.view(torch.int32) == unfused.e.view(torch.int32)).all()), info


@pytest.mark.parametrize("seed", range(2, N_CASES, 4))
def test_random_scenario_greedy_evaluation_equals_oracle(seed, cuda_device):
    """rlrm_evaluate (test_policy_optima batched) after some training, on the same random scenarios."""
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine

    sc, opts = random_scenario(seed)
    sc.shared_q = False
    c = P.compile_scenario(sc)
    eng = Engine(c, opts["n"])
    eng.reset()
    eng.train(opts["iters"])
    o = O.Oracle(c, opts["n"], sc.table_dtype)
    o.reset()
    o.q[...] = eng.q.cpu().numpy().reshape(o.q.shape)
    o.slot[...] = eng.slot.cpu().numpy().view(np.uint64)
    o.epsilon[...] = eng.epsilon.cpu().numpy()
    ev_g = eng.evaluate(2, sc.gamma, 10.0, t0=5000)
    ev_o = o.evaluate(2, sc.gamma, 10.0, t0=5000)
    info = f"seed {seed}: {sc.env}/{sc.map_name} A={len(sc.starts)} {sc.algo}"
    for f in ev_o.dtype.names:
        if f != "reserved":
            assert np.array_equal(ev_g[f], ev_o[f]), (info, f)


@pytest.mark.parametrize("seed", range(1, N_CASES, 8))
def test_random_scenario_product_mdp_equals_oracle(seed, cuda_device):
    """rlrm_mdp on the random scenarios: every (state, action, sub-action) outcome equals the oracle's reset + set_state + step."""
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine
    from multiagent_rlrm_b200.tables import mdp_action_distribution

    sc, _opts = random_scenario(seed)
    sc.shared_q = False
    c = P.compile_scenario(sc)
    eng, o = Engine(c, 1), O.Oracle(c, 1, sc.table_dtype)
    sub, _probs = mdp_action_distribution(sc)
    for k in range(len(sc.starts)):
        for rm_terminal in (True, False):
            got, exp = eng.mdp(k, sub, rm_terminal=rm_terminal), o.mdp(k, sub, rm_terminal=rm_terminal)
            for g, e, name in zip(got, exp, ("next_state", "reward", "done", "terminal")):
                assert np.array_equal(g, e), (seed, k, rm_terminal, name)


@pytest.mark.parametrize("n_agents", [1, 2, 3, 5, 6, 8])
@pytest.mark.parametrize("env", ["frozen_lake", "office_world"])
def test_sparse_qlambda_lane_groups(env, n_agents, cuda_device):
    """The sparse Q(lambda) kernel maps agents to lane groups of 32 / 16 / 8 / 4 lanes (1 / 2 / 3-4 / 5-8 agents, idle group
    slots when the agent count is not a power of two): every width against the oracle AND the dense kernel."""
    import oracle as O
    from multiagent_rlrm_b200.engine import Engine

    rng = np.random.default_rng(77 + n_agents)
    grid = frozen_lake_grid("map1") if env == "frozen_lake" else office_world_grid("map1")
    hazards = set(grid.hazards)
    free = [(x, y) for y in range(grid.height) for x in range(grid.width) if (x, y) not in hazards]
    rm, ev = random_machine(rng, free, "s")
    sc = P.Scenario(env=env, map_name="map1", starts=[free[i] for i in rng.choice(len(free), size=n_agents, replace=False)],
                    rm_transitions=rm, detector_positions=sorted(set(ev)), algo="qlambda", learning_rate=0.2, lambd=0.8, gamma=0.9,
                    q_init=0.0, epsilon_start=0.3, epsilon_end=0.3, epsilon_decay=1.0, stochastic=True, max_steps=60,
                    driver="frozen_lake_main" if env == "frozen_lake" else "office_main", seed=4242 + n_agents)
    c = P.compile_scenario(sc)
    n = 37
    sparse, dense, o = Engine(c, n, qlambda_sparse=True), Engine(c, n), O.Oracle(c, n, "f32")
    for x in (sparse, dense, o):
        x.reset()
    for chunk in (130, 1, 269):
        sparse.train(chunk); dense.train(chunk)
    o.train(0, 400)
    e_sparse = sparse.sync_tables(with_traces=True)
    for eng in (sparse, dense):
        assert np.array_equal(eng.slot.cpu().numpy().view(np.uint64), o.slot)
        assert np.array_equal(eng.q.cpu().numpy().reshape(-1), o.q.reshape(-1))
        assert np.array_equal(eng.stats_numpy()["episodes"].reshape(-1), o.stats["episodes"].reshape(-1))
    assert np.array_equal(e_sparse.cpu().numpy().reshape(-1), o.e.reshape(-1))
    assert np.array_equal(dense.e.cpu().numpy().reshape(-1), o.e.reshape(-1))
    assert int(o.stats["episodes"].sum()) > 0
