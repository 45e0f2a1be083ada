"""Batched greedy policy evaluation and Q-table export.

``test_policy_optima_batched`` is the batched counterpart of test_policy_optima / test_policy_opt_multi
(/root/reference/multiagent_rlrm/environments/utils_envs/evaluation_metrics.py:23-190, 505-697): every environment
instance plays ``episodi_test`` greedy episodes with its own tables on the device (rlrm_evaluate); the statistics the
reference returns per agent are returned per (instance, agent).

``save_q_tables`` / ``load_q_tables`` use the reference's ``data/q_tables.npz`` format (evaluation_metrics.py:193-214):
one array per agent under the key ``q_table_<agent name>``; for a batch the arrays gain a leading instance axis.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch


def test_policy_optima_batched(engine, episodi_test: int = 100, optimal_steps: float = 30, gamma: float = 0.9) -> Dict[str, np.ndarray]:
    ev = engine.evaluate(episodi_test, gamma, optimal_steps).reshape(engine.N, engine.A)
    episodes = np.maximum(ev["episodes"], 1).astype(np.float64)
    succ = ev["successes"].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_len = np.where(succ > 0, ev["len_sum"] / np.maximum(succ, 1), 0.0)
        var_len = np.where(succ > 0, ev["len_sqsum"] / np.maximum(succ, 1) - mean_len ** 2, 0.0)
        mean_ret = ev["return_sum"] / episodes
        var_ret = ev["return_sqsum"] / episodes - mean_ret ** 2
    return {
        "success_rate": 100.0 * succ / episodes,            # success_rate_per_agente
        "avg_timesteps": mean_len,                           # over successful episodes, 0 if none
        "std_timesteps": np.sqrt(np.maximum(var_len, 0.0)),
        "avg_reward": mean_ret,                              # discounted, until success
        "std_reward": np.sqrt(np.maximum(var_ret, 0.0)),
        "avg_arps": ev["arps_sum"] / episodes,
        "episodes": ev["episodes"],
    }


test_policy_optima_batched.__test__ = False  # not a pytest test despite the reference-derived name


def extract_policy_from_qtable(q_table) -> np.ndarray:
    """Greedy action per encoded state (first maximum), evaluation_metrics.py:455-502."""
    q = q_table.detach().cpu().numpy() if torch.is_tensor(q_table) else np.asarray(q_table)
    return np.argmax(q, axis=-1)


def save_q_tables(engine_or_agents, agent_names: Optional[Sequence[str]] = None, path: str = "data/q_tables.npz") -> str:
    """np.savez(path, q_table_<agent>=...) — the reference's format (evaluation_metrics.py:193-214)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    arrays = {}
    if hasattr(engine_or_agents, "q") and hasattr(engine_or_agents, "N"):  # engine.Engine
        eng = engine_or_agents
        names = list(agent_names) if agent_names else [f"a{k + 1}" for k in range(eng.A)]
        for k, name in enumerate(names):  # agent_table() flushes sparse Q(lambda) lists and handles per-agent table sections
            table = eng.agent_table(k).detach().cpu().numpy()
            arrays[f"q_table_{name}"] = table[0] if (table.ndim == 3 and table.shape[0] == 1) else table
    else:  # iterable of AgentRL-like objects
        for ag in engine_or_agents:
            arrays[f"q_table_{ag.name}"] = np.asarray(ag.get_learning_algorithm().q_table)
    np.savez(path, **arrays)
    return path


def load_q_tables(path: str) -> Dict[str, np.ndarray]:
    z = np.load(path)
    return {k[len("q_table_"):]: z[k] for k in z.files if k.startswith("q_table_")}


def load_q_tables_into(engine, path_or_tables, agent_names: Optional[Sequence[str]] = None) -> None:
    """Load `q_table_<agent>` arrays (reference format) into an engine: an (S, 4) array is broadcast to every instance,
    an (N, S, 4) array is taken per instance."""
    tables = load_q_tables(path_or_tables) if isinstance(path_or_tables, (str, os.PathLike)) else dict(path_or_tables)
    names = list(agent_names) if agent_names else [f"a{k + 1}" for k in range(engine.A)]
    q = engine.q.view(engine.A, engine.S, 4) if engine.cfg.shared_q else engine.q.view(engine.N, engine.A, engine.S, 4)
    for k, name in enumerate(names):
        t = torch.as_tensor(np.asarray(tables[name]), dtype=engine.q.dtype, device=engine.q.device)
        if engine.cfg.shared_q:
            q[k].copy_(t.reshape(engine.S, 4))
        else:
            q[:, k].copy_(t if t.dim() == 3 else t.reshape(1, engine.S, 4).expand(engine.N, -1, -1))


def test_policy_opt_multi_batched(engine, policies, episodes_test: int = 100, optimal_steps: float = 30, gamma: float = 0.9,
                                  test_deterministic: Optional[bool] = None) -> Dict[str, np.ndarray]:
    """Batched test_policy_opt_multi (evaluation_metrics.py:505-697): evaluate explicit per-agent policies
    (``policy[s] = action index``, e.g. from extract_policy_from_qtable or a value-iteration solution) instead of the
    learners' argmax. ``policies`` is [A, S] (one policy per agent for every instance) or [N, A, S].

    The policy is played by the same device rollout as test_policy_optima_batched: it is turned into one-hot tables, whose
    greedy action is unique, on a scratch engine compiled from the same scenario. Two details of the reference function are
    kept: its loop stops at ``timestep < 1000`` — one step before the environment's own truncation — so the scratch
    scenario's step cap is max_steps - 1; and ``test_deterministic`` sets ``env.stochastic = not test_deterministic``
    (an attribute only OfficeWorld reads) and plays a single episode when true. Returns the per-(instance, agent)
    statistics of test_policy_optima_batched (the reference's per-episode reward lists are not kept on the device)."""
    import copy

    from .engine import Engine
    from .tables import compile_scenario

    pol = policies.detach().cpu().numpy() if torch.is_tensor(policies) else np.asarray(policies)
    if pol.ndim == 2:
        pol = np.broadcast_to(pol[None], (engine.N,) + pol.shape)
    if pol.shape != (engine.N, engine.A, engine.S):
        raise ValueError(f"policies must be [A, S] or [N, A, S] with A={engine.A}, S={engine.S}; got {pol.shape}")
    if pol.min() < 0 or pol.max() >= 4:
        raise ValueError("policy holds an action index outside 0..3")
    if engine.cfg.per_agent_rm or engine.cfg.shared_q:
        raise NotImplementedError("explicit-policy evaluation is implemented for per-instance tables with one reward machine")
    sc = copy.deepcopy(engine.c.scenario)
    sc.max_steps = max(int(sc.max_steps) - 1, 0)
    sc.algo, sc.learning_rate, sc.use_rsh, sc.random_start_positions = "ql", 1.0, False, sc.random_start_positions
    if test_deterministic is not None:
        if sc.env == "office_world":
            sc.stochastic = not test_deterministic
        if test_deterministic:
            episodes_test = 1
    scratch = Engine(compile_scenario(sc, grid=engine.c.grid, rm=engine.c.rm, instance_offset=int(engine.cfg.instance_offset)), engine.N, device=engine.device, with_stats=False)
    onehot = torch.zeros_like(scratch.q).view(engine.N, engine.A, engine.S, 4)
    onehot.scatter_(3, torch.from_numpy(np.array(pol, dtype=np.int64)).to(engine.device).unsqueeze(-1), 1.0)
    scratch.q.copy_(onehot.view_as(scratch.q))
    scratch.t = engine.t
    return test_policy_optima_batched(scratch, episodes_test, optimal_steps, gamma)


test_policy_opt_multi_batched.__test__ = False
