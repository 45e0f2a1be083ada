"""GPU: the reference-shaped object API (envs.py / wrapper.py / agent.py / learners.py) is a drop-in — the reference's
own driver loop (restated once in oracle/ref_harness.run_reference) runs UNCHANGED over these classes and reproduces,
bit for bit, what the live reference produced (tests/golden). Also the reference's known-answer unit tests for this
path, restated against the CUDA-backed classes."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

DROPIN_CASES = [
    ("cfg1_det_qrm", 160), ("cfg3_slip_ql", 200), ("cfg3_slip_qrm", 160), ("fl_qlambda", 120),
    ("fl_delay_penalty_epsdecay_ql", 200), ("fl_lr_none_ql", 160), ("fl_reward_modifier_qrm", 120),
    ("fl_none_event_step_cost_ql", 120), ("cfg2_office_slip_ql", 200), ("ow_allslip_wallpen_exp3_qrm", 150),
    ("ow_terminate_plants_walls_ql", 250), ("cfg4_office_chain12_qlambda", 60), ("fl_shaping_vi_ql", 150),
    ("fl_shaping_distance_qrm", 150), ("ow_shaping_vi_exp3_qrm", 150), ("ow_map2_walls_qrm", 120), ("ow_map3_walls_qrm", 100),
    ("ow_map4_walls_qrm", 60), ("fl_random_starts_3agents_qrm", 200), ("fl_random_starts_6agents_ql", 120),
    ("fl_per_agent_rms_qrm", 200), ("fl_per_agent_rms_ql", 200), ("fl_per_agent_rms_qlambda", 150),
    ("fl_per_agent_rms_shaping_qrm", 200), ("fl_per_agent_rms_shaping_ql", 200),
    # the reference's NATIVE float64 tables: learners built with table_dtype="f64" reproduce the unmodified reference
    ("cfg1_det_qrm_f64", 160), ("cfg3_slip_ql_f64", 200), ("fl_qlambda_lr_none_f64", 120), ("cfg4_office_chain12_qlambda_f64", 40),
    ("fl_lr_none_qrm_f64", 160), ("cfg2_office_slip_ql_f64", 200), ("fl_shaping_vi_ql_f64", 150),
]


@pytest.mark.parametrize("name,iters", DROPIN_CASES)
def test_reference_driver_loop_over_dropin_classes(name, iters, cuda_device):
    import ref_harness as H
    from dropin_builder import build_b200

    meta, ref = load_golden(name)
    n = min(2, meta["n_instances"])
    f64 = meta["real"] == "f64"
    out = H.run_reference(meta["scenario"], n, iters, table_dtype=np.float64 if f64 else np.float32, pre_resets=meta["pre_resets"],
                          builder=build_b200)
    for k in ("action", "cell", "prev_cell", "q", "prev_q", "event_cell", "env_term", "rm_term", "term", "trunc", "active",
              "fail", "agent_steps", "timestep", "renv", "rq", "reward", "epsilon"):
        assert np.array_equal(out[k], ref[k][:iters, :n]), f"{name}: {k}"
    # Q value written by each update: pins the learner arithmetic step by step (float64 fixtures: all 64 bits)
    cast = np.float64 if f64 else np.float32
    assert np.array_equal(out["q_sa"].astype(cast), ref["q_sa"][:iters, :n].astype(cast)), f"{name}: q_sa"
    assert np.array_equal(out["episode_end"], ref["episode_end"][:iters, :n])


# ---- the reference's known-answer unit tests, restated ------------------------------------------------------------
def test_qlearning_update_greedy(cuda_device):
    """/root/reference/tests/test_qlearning.py:6-27"""
    import multiagent_rlrm_b200 as P

    ql = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=1.0, state_space_size=2, action_space_size=2,
                     qtable_init=1.0)
    ql.update(encoded_state=0, encoded_next_state=1, action=0, reward=1.0, terminated=False, info={})
    assert np.isclose(ql.q_table[0, 0], 1.0 + 0.9 * 1.0)
    assert ql.visits[0, 0] == 1


def test_qlearning_epsilon_decay_and_choice(cuda_device):
    """/root/reference/tests/test_qlearning.py:30-53"""
    import multiagent_rlrm_b200 as P

    ql = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=0.5, state_space_size=1, action_space_size=2,
                     epsilon_start=0.5, epsilon_end=0.1, epsilon_decay=0.5)
    ql.q_table[0] = np.array([0.2, 0.8])
    assert ql.choose_action(encoded_state=0, best=True) == 1
    ql.learn_done_episode()
    assert np.isclose(ql.epsilon, 0.25)
    ql.learn_done_episode()
    assert np.isclose(ql.epsilon, 0.125)
    ql.learn_done_episode()
    assert np.isclose(ql.epsilon, 0.1)
    # epsilon-greedy only ever returns usable actions, and exploits the maximum when epsilon is 0
    ql.epsilon = 0.0
    assert all(ql.choose_action(0) == 1 for _ in range(20))
    ql.epsilon = 1.0
    assert {ql.choose_action(0) for _ in range(64)} == {0, 1}
    # choose_action_greedy (qlearning.py:136-143): never explores, whatever epsilon is; ties are split uniformly
    assert all(ql.choose_action_greedy(0, ql.rng) == 1 for _ in range(20)) and ql.epsilon == 1.0
    ql.q_table[0] = np.array([0.8, 0.8])
    assert {ql.choose_action_greedy(0, ql.rng) for _ in range(64)} == {0, 1}


def test_qlearning_lambda_update_and_traces_decay(cuda_device):
    """/root/reference/tests/test_qlearning_lambda.py:6-27"""
    import multiagent_rlrm_b200 as P

    ql = P.QLearningLambda(gamma=0.9, lambd=0.5, action_selection="greedy", learning_rate=1.0, state_space_size=2,
                           action_space_size=1)
    ql.update(encoded_state=0, encoded_next_state=1, action=0, reward=1.0, terminated=False, next_action=0)
    assert np.isclose(ql.q_table[0, 0], 1.0)
    assert np.isclose(ql.e_table[0, 0], 0.9 * 0.5)


def test_qlearning_lambda_adaptive_step_size(cuda_device):
    """learning_rate=None -> lr = 1 / visits[s, a] (qlearning_lambda.py:44-49): against the same updates done by hand in the
    reference's float64-then-cast arithmetic."""
    import multiagent_rlrm_b200 as P

    ql = P.QLearningLambda(gamma=0.9, lambd=0.5, action_selection="greedy", learning_rate=None, state_space_size=3,
                           action_space_size=2)
    q = np.zeros((3, 2), dtype=np.float32)
    e = np.zeros((3, 2), dtype=np.float32)
    visits = np.zeros((3, 2))
    for (s, sn, a, r, term) in [(0, 1, 0, 1.0, False), (1, 2, 1, 0.3, False), (0, 1, 0, -0.7, False), (1, 0, 1, 2.0, True), (0, 1, 0, 0.1, False)]:
        ql.update(encoded_state=s, encoded_next_state=sn, action=a, reward=r, terminated=term)
        visits[s, a] += 1
        lr = 1 / visits[s, a]                                        # np.float64
        best = 0.0 if term else float(np.max(q[sn]))
        td = np.float32(r + 0.9 * best) - q[s, a]                      # python float is weak: rounded to float32 first
        e[s, a] = 1.0
        q = (q.astype(np.float64) + (lr * np.float64(td)) * e.astype(np.float64)).astype(np.float32)
        e = np.zeros_like(e) if term else e * np.float32(0.9 * 0.5)
        assert np.array_equal(np.asarray(ql.q_table), q) and np.array_equal(np.asarray(ql.e_table), e), (s, a)
    assert np.array_equal(np.asarray(ql.visits), visits)


def test_qlearning_lambda_reset_traces(cuda_device):
    """/root/reference/tests/test_qlearning_lambda.py:30-41"""
    import multiagent_rlrm_b200 as P

    ql = P.QLearningLambda(gamma=0.9, lambd=0.5, action_selection="greedy", learning_rate=0.5, state_space_size=1,
                           action_space_size=1)
    ql.e_table[0, 0] = 0.7
    ql.learn_init_episode()
    assert ql.e_table[0, 0] == 0.0


class _StubAgent:
    """/root/reference/tests/test_ma_frozen_lake.py:18-43"""

    def __init__(self, name, pos):
        self.name = name
        self.position = pos
        self.initial_position = pos
        self.state = {"pos_x": pos[0], "pos_y": pos[1]}
        self.reward_machine = None

    def get_position(self):
        return self.position

    def set_position(self, x, y):
        self.position = (x, y)
        self.state["pos_x"], self.state["pos_y"] = x, y

    def set_initial_position(self, x, y):
        self.initial_position = (x, y)
        self.set_position(x, y)

    def get_state(self):
        return self.state

    def reset(self):
        self.set_position(*self.initial_position)

    def get_learning_algorithm(self):
        return None


def test_frozen_lake_apply_action_bounds(cuda_device):
    """/root/reference/tests/test_ma_frozen_lake.py:46-58"""
    import multiagent_rlrm_b200 as P

    env = P.MultiAgentFrozenLake(width=2, height=2, holes=[])
    ag = _StubAgent("a1", (0, 0))
    env.add_agent(ag)
    env.apply_action(ag, "left"); assert ag.get_position() == (0, 0)
    env.apply_action(ag, "up"); assert ag.get_position() == (0, 0)
    env.apply_action(ag, "right"); assert ag.get_position() == (1, 0)
    env.apply_action(ag, "down"); assert ag.get_position() == (1, 1)


def test_frozen_lake_stochastic_outcome_sets(cuda_device):
    """/root/reference/tests/test_ma_frozen_lake.py:61-72"""
    import multiagent_rlrm_b200 as P

    env = P.MultiAgentFrozenLake(width=3, height=3, holes=[])
    ag = _StubAgent("a1", (1, 1))
    env.add_agent(ag)
    env.frozen_lake_stochastic = True
    env.reset(7)
    assert {str(env.get_stochastic_action(ag, "left")) for _ in range(200)} <= {"left", "up", "down"}
    env.delay_action = True
    assert {str(env.get_stochastic_action(ag, "up")) for _ in range(400)} <= {"wait", "up", "left", "right"}
    # and the device-side step only ever lands on the cells those outcomes allow
    env.delay_action = False
    seen = set()
    for seed in range(60):
        env.reset(seed)
        obs, rew, term, trunc, infos = env.step({"a1": "left"})
        seen.add((obs["a1"]["pos_x"], obs["a1"]["pos_y"]))
    assert seen <= {(0, 1), (1, 0), (1, 2)} and (0, 1) in seen


def test_reset_returns_dicts_keyed_by_agent(cuda_device):
    """/root/reference/tests/test_ma_frozen_lake.py:75-83 ; tests/test_officeworld_environment.py:5-25"""
    import multiagent_rlrm_b200 as P

    env = P.MultiAgentFrozenLake(width=2, height=2, holes=[])
    env.agents.append(_StubAgent("a1", (0, 0)))
    obs, infos = env.reset(1)
    assert set(obs) == {"a1"} and set(infos) == {"a1"} and obs["a1"] == {"pos_x": 0, "pos_y": 0}
    ow = P.MultiAgentOfficeWorld(width=3, height=3, plants=[], coffee=[], letters=[], walls=[], plants_penalty_value=-1,
                                 wall_penalty_value=0, terminate_on_plants=False, terminate_hit_walls=False)
    ow.agents.append(_StubAgent("a1", (1, 1)))
    obs, infos = ow.reset(1)
    assert set(obs) == {"a1"} and set(infos) == {"a1"}
    obs, rew, term, trunc, infos = ow.step({"a1": "up"})
    assert (obs["a1"]["pos_x"], obs["a1"]["pos_y"]) == (1, 2)  # OfficeWorld: up is y+1 (ma_office.py:280-281)


def test_wrapper_merges_env_and_rm_reward(cuda_device):
    """/root/reference/tests/test_rm_environment_wrapper.py:70-87: reward = env 0.5 + RM 1.0, termination on RM final,
    infos prev_q / q. (The env reward comes from a hole penalty here, the reference test uses a dummy env.)"""
    import multiagent_rlrm_b200 as P

    env = P.MultiAgentFrozenLake(width=2, height=1, holes=[(1, 0)])
    env.penalty_amount = 0.5
    ag = P.AgentRL("a1", env)
    ag.set_initial_position(0, 0)
    ag.add_state_encoder(P.StateEncoderFrozenLake(ag))
    ag.add_action_encoder(P.ActionEncoderFrozenLake(ag))
    ag.set_reward_machine(P.RewardMachine({("q0", (1, 0)): ("qf", 1.0)}, P.PositionEventDetector({(1, 0)})))
    ag.set_learning_algorithm(P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=1.0, state_space_size=4,
                                          action_space_size=4, use_qrm=True))
    env.add_agent(ag)
    wrap = P.RMEnvironmentWrapper(env, [ag])
    wrap.reset(0)
    obs, rewards, terms, truncs, infos = wrap.step({"a1": ag.action("right")})
    assert rewards["a1"] == 1.5
    assert terms["a1"] is True and truncs["a1"] is False
    assert infos["a1"]["prev_q"] == "q0" and infos["a1"]["q"] == "qf"
    assert infos["a1"]["env_terminated"] is True and infos["a1"]["rm_terminated"] is True
    # /root/reference/tests/test_rm_environment_wrapper.py:94-107: one experience per state in get_all_states()[:-1]
    assert len(infos["a1"]["qrm_experience"]) == len(ag.get_reward_machine().get_all_states()) - 1
    # reward_modifier scales the RM reward only (:110-130)
    wrap.reset(0)
    wrap.reward_modifier = 0.5
    _, rewards, _, _, infos = wrap.step({"a1": ag.action("right")})
    assert rewards["a1"] == 1.0 and infos["a1"]["RQ"] == 0.5


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_learner_pickle_round_trip_and_table_export(dtype, cuda_device, tmp_path):
    """office_main.py --save/--load pickles the learner; evaluation_metrics.save_q_tables writes q_table_<agent>.npz."""
    import pickle

    import multiagent_rlrm_b200 as P
    from multiagent_rlrm_b200.engine import Engine

    ql = P.QLearningLambda(gamma=0.9, lambd=0.5, action_selection="greedy", learning_rate=0.5, state_space_size=6, action_space_size=4,
                           table_dtype=dtype)
    assert np.asarray(ql.q_table).dtype == (np.float64 if dtype == "f64" else np.float32)
    ql.update(0, 1, 2, 1.0, False)
    ql.update(1, 2, 0, -1.0, False)
    ql.epsilon = 0.123
    clone = pickle.loads(pickle.dumps(ql))
    assert np.array_equal(np.asarray(clone.q_table), np.asarray(ql.q_table)) and np.array_equal(np.asarray(clone.e_table), np.asarray(ql.e_table))
    assert clone.epsilon == 0.123 and clone.lambd == 0.5
    clone.update(2, 3, 1, 0.5, True)
    ql.update(2, 3, 1, 0.5, True)
    assert np.array_equal(np.asarray(clone.q_table), np.asarray(ql.q_table))

    sc = P.scenario_config3()
    sc.table_dtype = dtype
    c = P.compile_scenario(sc)
    eng = Engine(c, 8)
    eng.reset(); eng.train(300)
    path = P.save_q_tables(eng, path=str(tmp_path / "data" / "q_tables.npz"))
    tables = P.load_q_tables(path)
    assert set(tables) == {"a1", "a2"} and tables["a1"].shape == (8, 400, 4)
    other = Engine(c, 8)
    P.load_q_tables_into(other, path)
    assert np.array_equal(other.q.cpu().numpy(), eng.q.cpu().numpy())
    single = Engine(c, 3)
    P.load_q_tables_into(single, {"a1": tables["a1"][0], "a2": tables["a2"][0]})
    assert np.array_equal(single.q.cpu().numpy().reshape(3, 2, 400, 4)[2], eng.q.cpu().numpy().reshape(8, 2, 400, 4)[0])


# ---------------------------------------------------------------------------------------------- look-ahead selection
def _genuine_rng_loop(sc_dict, iters, lookahead, seed=77):
    """The reference driver loop over the drop-in classes with GENUINE numpy generators (no injected words): the only
    configuration in which the learners' look-ahead selection (rlrm_update_list_select) is active."""
    import copy

    from multiagent_rlrm_b200.learners import _TabularBase

    from dropin_builder import build_b200

    old = _TabularBase._LOOKAHEAD
    _TabularBase._LOOKAHEAD = lookahead
    try:
        rm_env, env, agents = build_b200(sc_dict)
        for k, ag in enumerate(agents):
            ag.get_learning_algorithm().rng = np.random.default_rng(1000 * seed + k)
        fl = sc_dict["driver"] == "frozen_lake_main"
        log, it, ep = [], 0, 0
        while it < iters:
            states, _ = rm_env.reset(seed + ep)
            ep += 1
            if not fl:
                states = copy.deepcopy(states)
            while it < iters:
                actions = {ag.name: ag.select_action(rm_env.env.get_state(ag)) for ag in rm_env.agents}
                new_states, rewards, term, trunc, infos = rm_env.step(actions)
                log.append((tuple(actions[a.name].name for a in agents), tuple(rewards[a.name] for a in agents),
                            tuple(tuple(sorted(new_states[a.name].items())) for a in agents)))
                for ag in rm_env.agents:
                    ta = (term[ag.name] or trunc[ag.name]) if fl else term[ag.name]
                    ag.update_policy(state=states[ag.name], action=actions[ag.name], reward=rewards[ag.name],
                                     next_state=new_states[ag.name], terminated=ta, infos=infos[ag.name])
                states = copy.deepcopy(new_states)
                it += 1
                if all(term.values()) or all(trunc.values()):
                    break
        tables = [np.array(ag.get_learning_algorithm().q_table) for ag in agents]
        hits = sum(ag.get_learning_algorithm().lookahead_hits for ag in agents)
        misses = sum(ag.get_learning_algorithm().lookahead_misses for ag in agents)
        tail = []  # where every learner's word stream stands: eight forced explorations (action = f(next words))
        for ag in agents:
            la = ag.get_learning_algorithm()
            la.epsilon = 1.0
            tail.append([la.choose_action(0) for _ in range(8)])
        return log, tables, hits, misses, tail
    finally:
        _TabularBase._LOOKAHEAD = old


@pytest.mark.parametrize("name", ["cfg1", "cfg3_ql", "cfg2_office_slip", "office_qlambda"])
def test_lookahead_selection_changes_nothing(name, cuda_device):
    """update_policy launches the NEXT selection together with the update (rlrm_update_list_select) and select_action then
    reads the answer from page-locked memory. Decisions, rewards, positions, Q tables and the state of every learner's
    generator must be identical with the look-ahead switched off — and the look-ahead must actually be used."""
    import multiagent_rlrm_b200 as P

    if name == "office_qlambda":
        sc = P.scenario_config4()
        sc.starts, sc.max_steps = sc.starts[:2], 80
    else:
        sc = {"cfg1": P.scenario_config1, "cfg3_ql": lambda: P.scenario_config3(False), "cfg2_office_slip": lambda: P.scenario_config2(True)}[name]()
    d = sc.to_dict()
    on = _genuine_rng_loop(d, 700, True)
    off = _genuine_rng_loop(d, 700, False)
    assert on[0] == off[0]
    for a, b in zip(on[1], off[1]):
        assert np.array_equal(a, b)
    assert on[4] == off[4]  # every learner's word stream stands at the same position
    assert off[2] == 0 and on[2] > 0.8 * (on[2] + on[3]), (on[2], on[3])


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_update_list_select_equals_update_list_then_select(dtype, cuda_device):
    """C ABI: rlrm_update_list_select == rlrm_update_list followed by rlrm_select_action (same table, same action), for
    random experience lists (also empty ones), states, epsilons, words, with and without best."""
    import ctypes as C

    import torch

    import multiagent_rlrm_b200 as P
    from multiagent_rlrm_b200 import _abi as abi
    from multiagent_rlrm_b200._lib import check

    rng = np.random.default_rng(5)
    for trial in range(12):
        a = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=0.3, state_space_size=97, action_space_size=4,
                        table_dtype=dtype)
        b = P.QLearning(gamma=0.9, action_selection="greedy", learning_rate=0.3, state_space_size=97, action_space_size=4,
                        table_dtype=dtype)
        init = rng.integers(0, 4, size=(97, 4)).astype(np.float64)  # plenty of ties
        a.q_table[:] = init
        b.q_table[:] = init
        for step in range(20):
            n = int(rng.integers(0, 6))
            exps = [(int(rng.integers(0, 97)), int(rng.integers(0, 4)), float(rng.integers(-2, 3)), int(rng.integers(0, 97)),
                     bool(rng.integers(0, 2))) for _ in range(n)]
            state, best = int(rng.integers(0, 97)), bool(rng.integers(0, 4) == 0)
            eps = float(rng.choice([0.0, 0.3, 1.0]))
            words = [int(x) for x in rng.integers(0, 1 << 32, size=4, dtype=np.uint64)]
            # (1) two calls
            if exps:
                b._device_update_list(exps)
            b.epsilon = eps

            class _W:
                def words(self):
                    return words

            want = b.choose_action(state, best=best, rng=_W())
            # (2) one call
            req = abi.SelectReq.from_buffer(a._stage_np, a._SEL_OFF)
            req.state, req.best, req.epsilon, req.seq = state, int(best), eps, 1000 + step
            for j in range(4):
                req.draws[j] = words[j]
            for j, (s_, a_, r_, sn_, done_) in enumerate(exps):
                e = abi.Experience.from_buffer(a._stage_np, a._EXP_OFF + 24 * j)
                e.s, e.sn, e.action, e.terminated, e.reward = s_, sn_, a_, int(done_), r_
            check(a._th.L.rlrm_update_list_select(a._th.h, C.byref(a._st), 0, len(exps), a._base + a._EXP_OFF, a._base + a._SEL_OFF,
                                                  a._th.stream()))
            torch.cuda.synchronize()
            assert req.done_seq == 1000 + step and req.action == want, (trial, step)
            assert np.array_equal(np.array(a.q_table), np.array(b.q_table))
