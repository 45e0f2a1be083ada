"""The reference's API at batch granularity: the same method names and call order as the reference drivers, with
tensors shaped [N, A] (N instances, A agents) instead of dicts keyed by agent name.

    env = BatchedRMEnvironment(scenario, n_envs=65536)
    states, infos = env.reset()
    while ...:
        actions = env.select_action(states)                                   # ag.select_action(state) for every agent
        new_states, rewards, terminated, truncated, infos = env.step(actions)  # rm_env.step(actions)
        env.update_policy(states, actions, rewards, new_states, terminated | truncated, infos)   # ag.update_policy(...)
        states = new_states
        env.reset(mask=env.episode_over(terminated, truncated))               # rm_env.reset() of finished instances

Each call is one kernel of csrc/rlrm_b200.cu through the C ABI (rlrm_select_action / rlrm_step / rlrm_update /
rlrm_reset); `train(n)` is the fused equivalent of n iterations of the loop above (bit-identical, far faster).
Mirrors RMEnvironmentWrapper (rm_environment_wrapper.py:28-107), AgentRL.select_action / update_policy
(agent_rl.py:80-192) and the driver loops (frozen_lake_main.py:336-376, office_main.py:1696-1749).
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Callable, Dict, Optional, Tuple

import torch

from . import _abi as abi
from .engine import Engine
from .tables import Compiled, compile_scenario


class LazyTensors(Mapping):
    """Read-only dict of tensors whose entries are computed on first access. A driver that only needs the step's rewards
    and done flags does not pay a kernel launch for every derived view (pos_x, pos_y, boolean casts, ...)."""

    def __init__(self, makers: Dict[str, Callable[[], object]]):
        self._makers = makers
        self._cache: Dict[str, object] = {}

    def __getitem__(self, key):
        if key not in self._cache:
            self._cache[key] = self._makers[key]()
        return self._cache[key]

    def __iter__(self):
        return iter(self._makers)

    def __len__(self):
        return len(self._makers)


class BatchedRMEnvironment:
    def __init__(self, scenario, n_envs: int, device="cuda:0", instance_offset: int = 0, **engine_kwargs):
        compiled = scenario if isinstance(scenario, Compiled) else compile_scenario(scenario, instance_offset=instance_offset)
        self.compiled = compiled
        self.engine = Engine(compiled, n_envs, device=device, **engine_kwargs)
        self.n_envs, self.n_agents = self.engine.N, self.engine.A
        self.width = int(compiled.config.width)
        self.agents = [f"a{k + 1}" for k in range(self.n_agents)]
        self._rec = None

    # -- observations ------------------------------------------------------------------------------
    def _cells(self) -> torch.Tensor:
        return (self.engine.slot & 0xFFFF).view(self.n_envs, self.n_agents)

    def _obs(self, cells) -> Mapping:
        """Observation dict {"pos_x", "pos_y", "cell"} ([N, A] int64) over a cell-index tensor (or a thunk producing one)."""
        def cell():
            c = cells() if callable(cells) else cells
            return c.to(torch.int64) & 0xFFFF

        obs = LazyTensors({})
        obs._makers.update({"pos_x": lambda: obs["cell"] % self.width, "pos_y": lambda: obs["cell"] // self.width, "cell": cell})
        return obs

    @property
    def rm_state(self) -> torch.Tensor:
        """RewardMachine.current_state of every agent as a state index [N, A]."""
        return ((self.engine.slot >> abi.SLOT_RMSTATE_SHIFT) & 0xFF).view(self.n_envs, self.n_agents)

    @property
    def timestep(self) -> torch.Tensor:
        return ((self.engine.slot >> abi.SLOT_TIME_SHIFT) & 0xFFFF).view(self.n_envs, self.n_agents)[:, 0]

    @property
    def agent_steps(self) -> torch.Tensor:
        return ((self.engine.slot >> abi.SLOT_STEPS_SHIFT) & 0xFFFF).view(self.n_envs, self.n_agents)

    @property
    def active_agents(self) -> torch.Tensor:
        return (((self.engine.slot >> abi.SLOT_FLAGS_SHIFT) & abi.FLAG_ACTIVE) != 0).view(self.n_envs, self.n_agents)

    @property
    def q_table(self) -> torch.Tensor:
        """learner.q_table of every agent: [N, A, S, 4] (shared learner: [A, S, 4]); with per-agent reward machines the
        tables differ in size, so a list of [N, S_a, 4] tensors (one per agent) is returned instead."""
        e = self.engine
        e.sync_tables()
        if e.cfg.per_agent_rm:
            return [e.agent_table(a) for a in range(e.A)]
        return e.q.view(e.A, e.S, 4) if e.cfg.shared_q else e.q.view(e.N, e.A, e.S, 4)

    # -- the reference's calls ---------------------------------------------------------------------
    def reset(self, seed=None, mask: Optional[torch.Tensor] = None) -> Tuple[Dict[str, torch.Tensor], Dict]:
        """rm_env.reset(seed) for all (or the masked) instances. `seed` is accepted for signature compatibility; the
        randomness is the Philox stream keyed by scenario.seed and the global instance id."""
        self._first = None
        self.engine.reset(mask)
        return self._obs(self._cells()), {}

    def select_action(self, states=None, best: bool = False) -> torch.Tensor:
        """ag.select_action(state, best) for every agent -> action indices [N, A] (0 up, 1 down, 2 left, 3 right).
        `states` must be the environment's current observations (what the reference drivers pass)."""
        return self.engine.select_action(best=best)

    def step(self, actions: torch.Tensor):
        e = self.engine
        fl_driver = self.compiled.config.driver == abi.DRIVER_FROZEN_LAKE_MAIN
        slot_before = e.slot.clone() if fl_driver else None  # RLRM_FLAG_FIRST as of before the step (driver_states)
        self._first = None if slot_before is None else (lambda: ((slot_before >> abi.SLOT_FLAGS_SHIFT) & abi.FLAG_FIRST) != 0)
        rec = e.step(actions.reshape(-1))
        e.t += 1
        self._rec = rec
        v = lambda x: x.view(self.n_envs, self.n_agents)  # noqa: E731
        new_states = self._obs(v(rec["cell"]))
        infos = LazyTensors({"prev_s": lambda: self._obs(v(rec["prev_cell"])), "s": lambda: new_states,
                             "Renv": lambda: v(rec["renv"]), "RQ": lambda: v(rec["rq"]), "prev_q": lambda: v(rec["prev_q"]),
                             "q": lambda: v(rec["q"]), "event": lambda: v(rec["event"]),
                             "env_terminated": lambda: v(rec["env_term"]).bool(), "rm_terminated": lambda: v(rec["rm_term"]).bool()})
        return new_states, v(rec["reward"]), v(rec["term"]).view(torch.bool), v(rec["trunc"]).view(torch.bool), infos

    def update_policy(self, states, actions, rewards, next_states, terminated, infos=None):
        """ag.update_policy(state, action, reward, next_state, terminated, infos=...) for every agent. `states` are the
        observations the driver kept from before the step; rewards / next states / RM states come from the step record
        (they are what `infos` carries in the reference)."""
        if self._rec is None:
            raise RuntimeError("update_policy must follow step")
        obs_cell = states["cell"].reshape(-1).to(torch.int16)
        self.engine.update(obs_cell, actions.reshape(-1), terminated.reshape(-1).to(torch.uint8), self._rec)

    def episode_over(self, terminated: torch.Tensor, truncated: torch.Tensor) -> torch.Tensor:
        """The drivers' loop exit: all agents terminated, or all truncated (frozen_lake_main.py:345,375; office_main.py:1748)."""
        return terminated.all(dim=1) | truncated.all(dim=1)

    def driver_states(self, states, new_states):
        """What the reference driver holds in `states` when it calls update_policy: the previous observations — except on
        an episode's first iteration under the FrozenLake driver, where `states` still aliases the live agent.state and
        therefore already shows the NEW position (frozen_lake_main.py:337,359)."""
        if self.compiled.config.driver != abi.DRIVER_FROZEN_LAKE_MAIN or self._first is None:
            return states
        first = self._first().view(self.n_envs, self.n_agents)
        return self._obs(lambda: torch.where(first, new_states["cell"], states["cell"]))

    # -- fused ---------------------------------------------------------------------------------------
    def iterate(self, learn: bool = True, want_reward: bool = True):
        """The whole loop body above — select_action, step, update_policy, reset of the finished instances — as ONE launch
        (rlrm_iterate), for drivers that want the reference's per-iteration view of the run without five calls per
        iteration. Returns (record, rewards): page-locked HOST tensors the kernel wrote in place, int32 / float64 of shape
        [N, A]; `Engine.unpack_record(record)` gives action / executed action / new cell / new RM state / terminated /
        truncated / active. Bit-identical to one iteration of the loop above and of `train(1)`."""
        rec, rew = self.engine.iterate(learn=learn, want_reward=want_reward)
        self._rec = None
        return rec.view(self.n_envs, self.n_agents), None if rew is None else rew.view(self.n_envs, self.n_agents)

    def train(self, n_iters: int, learn: bool = True):
        return self.engine.train(n_iters, learn=learn)

    def evaluate(self, episodes: int, gamma: float, optimal_steps: float = 1.0):
        from .evaluation import test_policy_optima_batched

        return test_policy_optima_batched(self.engine, episodes, optimal_steps, gamma)
