"""Reference-shaped environments (one instance, dict-keyed-by-agent API) computed by the CUDA library.

Mirrors multi_agent/base_environment.py:29-80, environments/frozen_lake/ma_frozen_lake.py:12-350 and
environments/office_world/ma_office.py:20-452. Host objects (agents, dicts) are the visible state, exactly as in the
reference; every transition is computed on the device: before a call the host state is packed into the slot words,
`rlrm_step` runs, and the record is unpacked back into the objects. Behaviour flags are plain attributes set after
construction (as the reference's drivers do) and are compiled into the device tables lazily.

For large batches use engine.Engine / trainer.LockstepTrainer: same kernels, tensors instead of dicts.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import _abi as abi
from .actions import ACTION_ORDER, ActionRL
from .maps import GridSpec
from .reward_machine import PositionEventDetector, RewardMachine
from .tables import Scenario, compile_scenario

_A2I = {n: i for i, n in enumerate(ACTION_ORDER)}
_A2I["wait"] = abi.ACTION_WAIT


class BaseEnvironment:
    def __init__(self, width: int, height: int, device="cuda:0"):
        self._agents: List = []
        self.grid_width = width
        self.grid_height = height
        self.active_agents: Dict[str, bool] = {}
        self.agent_fail: Dict[str, bool] = {}
        self.agent_steps: Dict[str, int] = {}
        self.timestep = 0
        self.device = device
        self._engine = None
        self._engine_key = None
        self._engine_quick = None
        self._first = True
        self._last_rec = None

    @property
    def agents(self):
        return self._agents

    def add_agent(self, agent):
        self._agents.append(agent)

    def get_state(self, agent):
        return agent.get_state().copy()

    # ------------------------------------------------------------------ device plumbing
    def _grid(self) -> GridSpec:
        raise NotImplementedError

    def _scenario_fields(self) -> dict:
        raise NotImplementedError

    def _shared_rm(self):
        rms = [getattr(a, "reward_machine", None) for a in self.agents]
        rms = [r for r in rms if r is not None]
        if not rms:
            return None
        if len(rms) != len(self.agents):
            raise NotImplementedError("either every agent has a reward machine or none has")
        first = list(rms[0].transitions.items())
        if any(list(r.transitions.items()) != first or r.initial_state != rms[0].initial_state for r in rms[1:]):
            return rms  # different machines (frozen_lake_main.py --rm-spec-a1/--rm-spec-a2): per-agent table sections
        return rms[0]

    def _get_engine(self, reward_modifier=1):
        from .engine import Engine

        # cheap fingerprint first: the behaviour flags are plain attributes the caller may change between calls (as the
        # reference's drivers do), so they are looked at on every call; the reward machines are identified by object and size
        fields = self._scenario_fields()
        agents = self._agents
        quick = [reward_modifier, *fields.values()]
        for a in agents:
            rm = getattr(a, "reward_machine", None)
            quick += (tuple(getattr(a, "initial_position", None) or a.position), rm, len(getattr(rm, "transitions", ())))
        if self._engine is not None and quick == self._engine_quick:
            return self._engine
        starts = tuple(tuple(getattr(a, "initial_position", None) or a.position) for a in agents)
        rm = self._shared_rm()
        grid = self._grid()
        rm_list = rm if isinstance(rm, list) else ([] if rm is None else [rm])
        rm_key = tuple((tuple(m.transitions.items()), tuple(sorted(m.detector_positions())), m.initial_state) for m in rm_list) or None
        key = (tuple(sorted(fields.items())), starts, rm_key, reward_modifier,
               grid.width, grid.height, tuple(grid.hazards), tuple(grid.walls))
        if self._engine is None or key != self._engine_key:
            sc = Scenario(env=grid.env, starts=list(starts), algo="ql", learning_rate=1.0, reward_modifier=reward_modifier, **fields)
            if rm is None:  # an environment without reward machines: one-state machine, no transitions, no final state
                rm_c = RewardMachine({}, PositionEventDetector(set()))
                rm_c.current_state = rm_c.initial_state = "__none__"
                rm_c.state_indices = {"__none__": 0}
                rm_c.numbers_state = lambda: 1
                rm_c.get_all_states = lambda: ["__none__"]
            else:
                rm_c = rm
            # one instance: the control state lives in page-locked host memory the kernels access in place (no copies)
            self._engine = Engine(compile_scenario(sc, grid=grid, rm=rm_c), 1, device=self.device, with_stats=False, host_control=True)
            self._engine_key = key
            self._slot_np = self._engine.slot.numpy()
        self._engine_quick = quick
        return self._engine

    def _pack_slots(self, out):
        W, first, t = self.grid_width, (abi.FLAG_FIRST if self._first else 0), int(self.timestep) << abi.SLOT_TIME_SHIFT
        active, fail, steps = self.active_agents, self.agent_fail, self.agent_steps
        for i, a in enumerate(self._agents):
            x, y = a.get_position()
            rm = getattr(a, "reward_machine", None)
            q = rm.get_state_index(rm.get_current_state()) if rm is not None else 0
            name = a.name
            flags = (abi.FLAG_ACTIVE if active.get(name, True) else 0) | (abi.FLAG_FAIL if fail.get(name, False) else 0) | first
            out[i] = ((y * W + x) << abi.SLOT_CELL_SHIFT) | (int(steps.get(name, 0)) << abi.SLOT_STEPS_SHIFT) | t \
                | (q << abi.SLOT_RMSTATE_SHIFT) | (flags << abi.SLOT_FLAGS_SHIFT)

    _WORD_BLOCK = 64  # slip words taken from env.rng per numpy call

    def _slip_words(self, out):
        """One 32-bit slip word per agent from env.rng (`rng.words(i)` hook = trace injection) written into the uint32 [A*4]
        draw block `out` (word 3 of every agent). A genuine numpy generator is read _WORD_BLOCK words at a time — the same
        sequence as per-step requests; a block belongs to the generator object it was drawn from (reset() installs a new one)."""
        n = len(self._agents)
        rng = self.rng if self.rng is not None else np.random.default_rng()
        if hasattr(rng, "words"):
            for i in range(n):
                out[4 * i:4 * i + 4] = [int(v) & 0xFFFFFFFF for v in rng.words(i)]
            return
        buf = self.__dict__.get("_slip_buf")
        if buf is None or buf[0] is not rng or buf[2] + n > len(buf[1]):
            buf = self.__dict__["_slip_buf"] = [rng, rng.integers(0, 1 << 32, size=n * max(1, self._WORD_BLOCK // n), dtype=np.uint64).astype(np.uint32), 0]  # a whole number of steps: no word is dropped
        k = buf[2]
        buf[2] = k + n
        out[3::4] = buf[1][k:k + n]

    def _device_step(self, actions, with_rm: bool, reward_modifier=1, counterfactuals=False):
        """rlrm_step on the packed host state: ONE launch that reads the slot words / actions / slip words from page-locked
        host memory and writes the new slot words and the step record back in place, one stream synchronisation; then the
        host objects are brought up to date. Returns the record as numpy views (valid until the next step)."""
        eng = self._get_engine(reward_modifier)
        acts = []
        stochastic_fl = getattr(self, "frozen_lake_stochastic", False)
        agents = self._agents
        for a in agents:
            act = actions[a.name]
            name = act if isinstance(act, str) else act.name
            if name not in _A2I:
                raise ValueError(f"Invalid action: {name}")
            if name == "wait" and stochastic_fl:
                raise KeyError(name)  # the slippery FrozenLake action map has no "wait" entry (ma_frozen_lake.py:283-296)
            acts.append(_A2I[name])
        slot = self._slot_np
        self._pack_slots(slot)
        draws = None
        if eng.cfg.stochastic:
            self._slip_words((eng._hio or eng._host_io())["v"]["draws"])
            draws = True  # written in place
        out = eng.step_host(acts, draws, with_rm=with_rm, counterfactuals=counterfactuals)
        W = self.grid_width
        timestep = self.timestep + 1
        active, fail, steps = self.active_agents, self.agent_fail, self.agent_steps
        words = slot[:len(agents)].tolist()
        for i, a in enumerate(agents):
            w = words[i]
            cell = w & 0xFFFF
            x, y = cell % W, cell // W
            if (x, y) != tuple(a.get_position()):
                a.set_position(x, y)
            fl = (w >> abi.SLOT_FLAGS_SHIFT) & 0xFF
            name = a.name
            active[name] = bool(fl & abi.FLAG_ACTIVE)
            fail[name] = bool(fl & abi.FLAG_FAIL)
            steps[name] = (w >> abi.SLOT_STEPS_SHIFT) & 0xFFFF
            if with_rm:
                rm = getattr(a, "reward_machine", None)
                if rm is not None:
                    rm.current_state = rm.get_state_from_index((w >> abi.SLOT_RMSTATE_SHIFT) & 0xFF)
            if i == 0:
                timestep = (w >> abi.SLOT_TIME_SHIFT) & 0xFFFF
        self.timestep = timestep
        self._first = False
        self._last_rec = out
        return out

    def _pos(self, cell):
        cell = int(cell)
        return {"pos_x": cell % self.grid_width, "pos_y": cell // self.grid_width}

    # ------------------------------------------------------------------ shared reset pieces
    def _reset_common(self, seed):
        from .learners import QLearning, QLearningLambda

        self.rewards = {a.name: 0 for a in self.agents}
        self.timestep = 0
        self.active_agents = {a.name: True for a in self.agents}
        self.agent_fail = {a.name: False for a in self.agents}
        self.agent_steps = {a.name: 0 for a in self.agents}
        self._first = True
        return QLearning, QLearningLambda

    def set_state(self, agent, state):
        x, y, q_rm = state
        agent.set_position(x, y)
        rm = agent.get_reward_machine()
        rm.current_state = rm.get_state_from_index(q_rm) if isinstance(q_rm, int) else q_rm

    def get_current_state(self, agent):
        x, y = agent.get_position()
        rm = agent.get_reward_machine()
        return (x, y, rm.get_state_index(rm.get_current_state()))

    def get_action_distribution(self, action):
        subactions, probabilities = self.get_action_probability_mapping()[action.name]
        agent = self.agents[0]
        return [self.wait_action if n == "wait" else agent.action(n) for n in subactions], probabilities

    def _next_cell(self, agent, action_name):
        """Host view of the compiled move table (single-agent helper API only)."""
        eng = self._get_engine()
        x, y = agent.get_position()
        return int(eng.c.next_cell[y * self.grid_width + x, _A2I[action_name]])


class MultiAgentFrozenLake(BaseEnvironment):
    metadata = {"name": "multi_agent_frozen_lake"}

    def __init__(self, width, height, holes, device="cuda:0"):
        super().__init__(width, height, device)
        self.map_width, self.map_height = width, height
        self.holes = holes
        self.wait_action = ActionRL("wait", [], [])
        self.possible_actions = list(ACTION_ORDER)
        self.rewards = 0
        self.frozen_lake_stochastic = False
        self.penalty_amount = 0
        self.delay_action = False
        self.epsilon = None
        self.random_start_positions = False
        self.rng = np.random.default_rng()

    def _grid(self):
        return GridSpec("frozen_lake", self.grid_width, self.grid_height, hazards=[tuple(h) for h in self.holes])

    def _scenario_fields(self):
        return dict(stochastic=bool(self.frozen_lake_stochastic), delay_action=bool(self.delay_action),
                    penalty_amount=self.penalty_amount)

    def reset(self, seed=123, options=None):
        QLearning, QLearningLambda = self._reset_common(seed)
        self.rng = np.random.default_rng(seed) if seed is not None else np.random.default_rng()
        if self.random_start_positions:
            starts = self._sample_start_positions(self.rng)
        else:
            starts = [getattr(a, "initial_position", None) or a.position for a in self.agents]
        for agent, start in zip(self.agents, starts):
            agent.set_initial_position(*start)
            agent.reset()
            algo = agent.get_learning_algorithm()
            if isinstance(algo, QLearningLambda):
                algo.reset_e_table()
            if isinstance(algo, QLearning):
                algo.learn_done_episode()
        return {a.name: a.state for a in self.agents}, {a.name: {} for a in self.agents}

    def _sample_start_positions(self, rng):
        free = [(x, y) for x in range(self.grid_width) for y in range(self.grid_height) if (x, y) not in self.holes]
        if len(free) < len(self.agents):
            raise ValueError("Not enough free cells to place all agents.")
        rng.shuffle(free)
        return free[: len(self.agents)]

    def step(self, actions):
        return self._step(actions, with_rm=False)[:5]

    def _step(self, actions, with_rm, reward_modifier=1, counterfactuals=False):
        rec = self._device_step(actions, with_rm, reward_modifier, counterfactuals)
        agents, W = self._agents, self.grid_width
        n = len(agents)
        prev, cell, renv = rec["prev_cell"][:n].tolist(), rec["cell"][:n].tolist(), rec["renv"][:n].tolist()
        env_term, trunc = rec["env_term"][:n].tolist(), rec["trunc"][:n].tolist()
        rewards, infos, terms, truncs, obs = {}, {}, {}, {}, {}
        for i, a in enumerate(agents):
            name = a.name
            r = renv[i]
            r = int(r) if r == int(r) else r  # _num
            p, c = prev[i], cell[i]
            infos[name] = {"prev_s": {"pos_x": p % W, "pos_y": p // W}, "s": {"pos_x": c % W, "pos_y": c // W}, "Renv": r}
            rewards[name] = 0 + r
            terms[name] = bool(env_term[i])
            truncs[name] = bool(trunc[i])
            obs[name] = a.state
        self.rewards = rewards
        return obs, rewards, terms, truncs, infos, rec

    def holes_in_the_ice(self, state, agent_name):
        if (state["pos_x"], state["pos_y"]) in self.holes:
            self.agent_fail[agent_name] = True
            return self.penalty_amount
        return 0

    def check_terminations(self):
        terms = {a.name: False for a in self.agents}
        truncs = {a.name: False for a in self.agents}
        for a in self.agents:
            if self.agent_steps.get(a.name, 0) > 1000 or self.timestep > 1000:
                terms[a.name] = truncs[a.name] = True
            rm = getattr(a, "reward_machine", None)
            if rm and rm.get_current_state() == rm.get_final_state():
                terms[a.name] = True
            if self.agent_fail[a.name]:
                terms[a.name] = True
        return terms, truncs

    def apply_action(self, agent, action_name: str):
        if action_name in _A2I and action_name != "wait":
            cell = self._next_cell(agent, action_name)
            agent.set_position(cell % self.grid_width, cell // self.grid_width)
        else:
            agent.set_position(*agent.get_position())

    def get_stochastic_action(self, agent, intended_action):
        actions, probabilities = self._stochastic_action_probability_mapping()[intended_action]
        rng = self.rng or np.random.default_rng()
        return rng.choice(actions, p=probabilities)

    def get_action_probability_mapping(self):
        if not self.frozen_lake_stochastic:
            return {n: ([n], [1.0]) for n in ("left", "right", "up", "down", "wait")}
        return self._stochastic_action_probability_mapping()

    def _stochastic_action_probability_mapping(self):
        perp = {"left": ("up", "down"), "right": ("up", "down"), "up": ("left", "right"), "down": ("left", "right")}
        if self.delay_action:
            return {a: (["wait", a, p[0], p[1]], [0.6, 0.36, 0.02, 0.02]) for a, p in perp.items()}
        return {a: ([a, p[0], p[1]], [0.8, 0.1, 0.1]) for a, p in perp.items()}

    def is_terminal_state_mdp(self, agent, pos_x, pos_y, rm_state):
        if (pos_x, pos_y) in self.holes:
            return True, self.penalty_amount
        rm = agent.get_reward_machine()
        name = rm.get_state_from_index(rm_state) if isinstance(rm_state, int) else rm_state
        return (True, 0) if name == rm.get_final_state() else (False, 0)


def _num(x):
    """Rewards leave the device as doubles; hand integers back as ints like the reference's Python numbers."""
    x = float(x)
    return int(x) if x == int(x) else x


class MultiAgentOfficeWorld(BaseEnvironment):
    metadata = {"name": "multi_agent_office_world"}

    def __init__(self, width, height, plants, coffee, letters, walls, plants_penalty_value, wall_penalty_value,
                 terminate_on_plants, terminate_hit_walls, all_slip=False, device="cuda:0"):
        super().__init__(width, height, device)
        self.wait_action = ActionRL("wait", [], [])
        self.plants, self.coffee, self.letters, self.walls = plants, coffee, letters, walls
        self.terminate_on_plants = terminate_on_plants
        self.terminate_hit_walls = terminate_hit_walls
        self.possible_actions = list(ACTION_ORDER)
        self.rewards = 0
        self.stochastic = False
        self.high_prob = 0.8
        self.plants_penalty_value = plants_penalty_value
        self.wall_penalty_value = wall_penalty_value
        self.delay_action = False
        self.all_slip = all_slip
        self.map_height, self.map_width = height, width
        self.rng = None

    def _grid(self):
        return GridSpec("office_world", self.grid_width, self.grid_height, hazards=[tuple(p) for p in self.plants],
                        walls=[(tuple(a), tuple(b)) for (a, b) in self.walls], coffee=list(self.coffee), letters=list(self.letters))

    def _scenario_fields(self):
        return dict(stochastic=bool(self.stochastic), delay_action=bool(self.delay_action), all_slip=bool(self.all_slip),
                    high_prob=self.high_prob, plants_penalty=self.plants_penalty_value, wall_penalty=self.wall_penalty_value,
                    terminate_on_plants=bool(self.terminate_on_plants), terminate_hit_walls=bool(self.terminate_hit_walls))

    def reset(self, seed=123, options=None):
        QLearning, QLearningLambda = self._reset_common(seed)
        self.rng = np.random.default_rng(seed=seed)
        for agent in self.agents:
            agent.reset()
            algo = agent.get_learning_algorithm()
            if isinstance(algo, QLearningLambda):
                algo.reset_e_table()
            if isinstance(algo, QLearning):
                algo.learn_done_episode()
        return {a.name: a.state for a in self.agents}, {a.name: {} for a in self.agents}

    def step(self, actions):
        return self._step(actions, with_rm=False)[:5]

    def _step(self, actions, with_rm, reward_modifier=1, counterfactuals=False):
        was_active = dict(self.active_agents)
        rec = self._device_step(actions, with_rm, reward_modifier, counterfactuals)
        agents, W = self._agents, self.grid_width
        n = len(agents)
        prev, cell, renv = rec["prev_cell"][:n].tolist(), rec["cell"][:n].tolist(), rec["renv"][:n].tolist()
        env_term, trunc = rec["env_term"][:n].tolist(), rec["trunc"][:n].tolist()
        rewards, infos, terms, truncs, obs = {}, {}, {}, {}, {}
        for i, a in enumerate(agents):
            name = a.name
            rewards[name] = 0
            if was_active.get(name, True):  # inactive agents are skipped and keep an empty info dict (ma_office.py:143-144)
                r = renv[i]
                r = int(r) if r == int(r) else r  # _num
                p, c = prev[i], cell[i]
                infos[name] = {"prev_s": {"pos_x": p % W, "pos_y": p // W}, "s": {"pos_x": c % W, "pos_y": c // W}, "Renv": r}
                rewards[name] += r
            else:
                infos[name] = {}
            terms[name] = bool(env_term[i])
            truncs[name] = bool(trunc[i])
            obs[name] = a.state
        self.rewards = rewards
        return obs, rewards, terms, truncs, infos, rec

    def plants_in_the_office(self, state, agent_name):
        if (state["pos_x"], state["pos_y"]) in self.plants:
            if self.terminate_on_plants:
                self.agent_fail[agent_name] = True
            return self.plants_penalty_value
        return 0

    def calculate_environment_rewards(self, agent, state, wall_penalty):
        """wall penalty + plant penalty (ma_office.py:222-238); single-agent helper, the batched step computes the same sum"""
        return wall_penalty + self.plants_in_the_office(state, agent.name)

    def check_terminations(self):
        terms = {a.name: bool(self.agent_fail[a.name]) for a in self.agents}
        truncs = {a.name: self.timestep > 1000 for a in self.agents}
        return terms, truncs

    def is_wall_collision(self, agent, action_name):
        if action_name == "wait":
            return False
        if action_name not in ACTION_ORDER:
            raise ValueError(f"Invalid action: {action_name}")
        x, y = agent.get_position()
        return self._next_cell(agent, action_name) == y * self.grid_width + x

    def apply_wall_penalty(self, agent, intended_action_name):
        if self.is_wall_collision(agent, intended_action_name):
            if self.terminate_hit_walls:
                self.agent_fail[agent.name] = True
            return self.wall_penalty_value, self.wait_action.name
        return 0, intended_action_name

    def apply_action(self, agent, action_name: str):
        if action_name in ACTION_ORDER:
            cell = self._next_cell(agent, action_name)
            agent.set_position(cell % self.grid_width, cell // self.grid_width)
        else:
            agent.set_position(*agent.get_position())

    def get_action_probability_mapping(self):
        perp = {"left": ("up", "down"), "right": ("up", "down"), "up": ("left", "right"), "down": ("left", "right")}
        if self.delay_action:
            return {a: (["wait", a, p[0], p[1]], [0.6, 0.36, 0.02, 0.02]) for a, p in perp.items()}
        if not self.stochastic:
            return {n: ([n], [1.0]) for n in ("left", "right", "up", "down", "wait")}
        hp = self.high_prob
        if self.all_slip:
            lp = (1 - hp) / 3
            order = {"left": ["left", "right", "up", "down"], "right": ["right", "left", "up", "down"],
                     "up": ["up", "down", "left", "right"], "down": ["down", "up", "left", "right"]}
            return {a: (o, [hp, lp, lp, lp]) for a, o in order.items()}
        lp = (1 - hp) / 2
        return {a: ([a, p[0], p[1]], [hp, lp, lp]) for a, p in perp.items()}

    def get_stochastic_action(self, agent, intended_action_name):
        actions, probabilities = self.get_action_probability_mapping()[intended_action_name]
        return self.rng.choice(actions, p=probabilities)

    def is_terminal_state_mdp(self, agent, pos_x, pos_y, rm_state):
        if (pos_x, pos_y) in self.plants and self.terminate_on_plants:
            return True, self.plants_penalty_value
        if rm_state == agent.get_reward_machine().get_final_state():
            return True, 0
        return False, 0
