"""Reward Machine: host-side automaton bookkeeping + compilation to dense device tables.

Drop-in for ``multiagent_rlrm.multi_agent.reward_machine.RewardMachine``
(/root/reference/multiagent_rlrm/multi_agent/reward_machine.py:5-195) and the position event detectors
(environments/frozen_lake/detect_event.py:5-33, environments/office_world/detect_event.py:4-27).

The automaton itself (a dict ``{(state, event): (next_state, reward)}``) stays a Python object: it is
configuration, built once. What runs per step — ``step`` / ``get_reward_for_non_current_state`` over a batch —
is a table lookup executed by the CUDA library (csrc/rlrm_b200.cu) on the dense ``delta`` / ``rq`` tables that
:meth:`RewardMachine.compile_tables` emits. The bookkeeping rules the tables depend on are reproduced exactly:

* index map: initial state -> 0, the rest sorted lexicographically          (reward_machine.py:20-39)
* final state: target of the LAST inserted transition                        (reward_machine.py:152-163)
* initial state: source of the FIRST inserted transition                     (reward_machine.py:165-177)
* numbers_state(): distinct states appearing in transitions                  (reward_machine.py:130-138)
* get_all_states(): first-appearance order                                   (reward_machine.py:95-111)
"""
from __future__ import annotations

from typing import Dict, Hashable, Iterable, List, Optional, Tuple

import numpy as np

EVENT_NONE = 255
NO_TRANSITION = 255


class EventDetector:
    """Interface of multiagent_rlrm.multi_agent.event_detector.EventDetector (event_detector.py:4-20)."""

    def detect_event(self, current_state):  # pragma: no cover - interface
        raise NotImplementedError


class PositionEventDetector(EventDetector):
    """Event = the (pos_x, pos_y) tuple itself when it is one of ``positions``, else None."""

    def __init__(self, positions: Iterable[Tuple[int, int]]):
        self.positions = positions

    def detect_event(self, current_state):
        here = (current_state["pos_x"], current_state["pos_y"])
        return here if here in self.positions else None


class RewardMachine:
    def __init__(self, transitions: Dict[Tuple[Hashable, Hashable], Tuple[Hashable, float]], event_detector):
        self.transitions = transitions
        self.initial_state = self._get_start_state()
        self.current_state = self.initial_state
        self.state_indices = self._generate_state_indices()
        self.event_detector = event_detector
        self.potentials = None

    # ------------------------------------------------------------------ bookkeeping
    def _states_in_order(self) -> List[Hashable]:
        """First-appearance order of the states (reward_machine.py:95-111). The reference re-derives it on every call — it is
        the top line of its profile, reached from every encode(); here it is cached and re-derived only when the transition
        dict was replaced or changed size (configuration code builds machines by inserting transitions)."""
        tr = self.transitions
        cached = self.__dict__.get("_order_cache")
        if cached is not None and cached[0] is tr and cached[1] == len(tr):
            return cached[2]
        order: List[Hashable] = []
        seen = set()
        for (src, _ev), (dst, _r) in tr.items():
            for s in (src, dst):
                if s not in seen:
                    seen.add(s)
                    order.append(s)
        self.__dict__["_order_cache"] = (tr, len(tr), order)
        return order

    def _generate_state_indices(self):
        rest = set(self._states_in_order()) - {self.current_state}
        ordered = [self.current_state] + sorted(rest)  # initial state first, the others in sorted order
        return {s: i for i, s in enumerate(ordered)}

    def _get_start_state(self):
        for (src, _ev) in self.transitions:
            return src
        return None

    def get_state_index(self, rm_state):
        return self.state_indices[rm_state]

    def get_state_from_index(self, rm_state_index):
        """Inverse of state_indices (reward_machine.py:41-43 scans the dict on every call; here the inverse is cached and
        rebuilt when state_indices was replaced or changed size)."""
        si = self.state_indices
        cached = self.__dict__.get("_inverse_cache")
        if cached is None or cached[0] is not si or cached[1] != len(si):
            cached = self.__dict__["_inverse_cache"] = (si, len(si), {idx: state for state, idx in reversed(list(si.items()))})
        try:
            return cached[2][rm_state_index]
        except (KeyError, TypeError):
            raise ValueError(f"Index {rm_state_index} not present in RewardMachine.state_indices") from None

    def get_all_states(self):
        return list(self._states_in_order())

    def numbers_state(self):
        return len(self._states_in_order())

    def get_final_state(self):
        last = None
        for last in self.transitions.values():
            pass
        return None if last is None else last[0]

    def get_possible_events(self, state_rm):
        return [ev for (src, ev) in self.transitions if src == state_rm]

    def get_current_state(self):
        return self.current_state

    @property
    def get_transitions(self):
        return self.transitions

    def reset_to_initial_state(self):
        self.current_state = self.initial_state
        return self.initial_state

    # ------------------------------------------------------------------ single-instance stepping (host bookkeeping)
    # These three are pure dict lookups on ONE automaton instance; they exist so that configuration code and the
    # reference's own unit tests (tests/test_reward_machine.py) run unchanged. The batched hot path never calls
    # them: it goes through rlrm_step / rlrm_rm_step on the compiled tables.
    def get_reward(self, event):
        hit = self.transitions.get((self.current_state, event))
        if hit is None:
            return 0
        self.current_state = hit[0]
        return hit[1]

    def step(self, current_state):
        return self.get_reward(self.event_detector.detect_event(current_state))

    def get_reward_for_non_current_state(self, state_rm, event):
        if isinstance(event, list):
            event = tuple(event)
        hit = self.transitions.get((state_rm, event))
        return (None, 0) if hit is None else (hit[0], hit[1])

    # ------------------------------------------------------------------ compilation to device tables
    def detector_positions(self) -> List[Tuple[int, int]]:
        """Sorted positions that can ever be reported as an event."""
        pos = getattr(self.event_detector, "positions", None)
        if pos is None:
            pos = [ev for (_s, ev) in self.transitions if isinstance(ev, tuple) and len(ev) == 2]
        return sorted({(int(p[0]), int(p[1])) for p in pos})

    def compile_tables(self, width: int, height: int, reward_modifier=1):
        """Dense tables consumed by the CUDA library (include/rlrm_b200.h: rlrm_tables_t).

        Returns a dict with
          label [W*H] u8   event id per cell (EVENT_NONE when the position is not a detector position)
          delta [nQ, nEv+1] u8   next-state index (NO_TRANSITION = stay); column nEv is the ``None`` event
          rq    [nQ, nEv+1] f64  transition reward * reward_modifier      (rm_environment_wrapper.py:65-69)
          rcf   [nQ, nEv+1] f64  transition reward (QRM counterfactuals)  (rm_environment_wrapper.py:150-153)
          qrm_states [n] u8      indices of get_all_states()[:-1]         (rm_environment_wrapper.py:144)
          final, n_states, n_events, events (list of positions, id = list index)
        """
        events = [p for p in self.detector_positions() if 0 <= p[0] < width and 0 <= p[1] < height]
        if len(events) > 63:
            raise ValueError("at most 63 event positions are supported")
        n_states = len(self.state_indices)
        if self.numbers_state() != n_states:
            # Reference encoders size the table with numbers_state() but index with state_indices; they only
            # disagree for a machine without transitions, which cannot be stepped anyway.
            raise ValueError("reward machine has states outside its transitions")
        if n_states > 32:
            raise ValueError("at most 32 RM states are supported")
        n_ev = len(events)
        label = np.full(width * height, EVENT_NONE, dtype=np.uint8)
        for k, (x, y) in enumerate(events):
            label[y * width + x] = k
        ev_col = {p: k for k, p in enumerate(events)}
        ev_col[None] = n_ev
        delta = np.full((n_states, n_ev + 1), NO_TRANSITION, dtype=np.uint8)
        rq = np.zeros((n_states, n_ev + 1), dtype=np.float64)
        rcf = np.zeros((n_states, n_ev + 1), dtype=np.float64)
        for (src, ev), (dst, reward) in self.transitions.items():
            if isinstance(ev, list):
                ev = tuple(ev)
            col = ev_col.get(ev) if (ev is None or isinstance(ev, tuple)) else None
            if col is None:
                continue  # an event no PositionEventDetector can emit on this grid
            s = self.state_indices[src]
            delta[s, col] = self.state_indices[dst]
            rq[s, col] = reward * reward_modifier
            rcf[s, col] = reward
        final = self.get_final_state()
        qrm_states = np.array([self.state_indices[s] for s in self.get_all_states()[:-1]], dtype=np.uint8)
        return {
            "label": label,
            "delta": delta,
            "rq": rq,
            "rcf": rcf,
            "qrm_states": qrm_states,
            "final": -1 if final is None else self.state_indices[final],
            "n_states": n_states,
            "n_events": n_ev,
            "events": events,
        }

    # ------------------------------------------------------------------ reward shaping (reward_machine.py:197-345)
    def add_reward_shaping(self, gamma, rs_gamma):
        from .shaping import value_iteration_potentials

        self.gamma = gamma
        self.potentials = value_iteration_potentials(self, rs_gamma)

    def add_distance_reward_shaping(self, gamma, rs_gamma, alpha=100):
        from .shaping import distance_potentials

        self.potentials = distance_potentials(self, alpha)

    def get_distance(self, start_state):
        from .shaping import rm_distance

        return rm_distance(self, start_state)

    def get_delta_u(self):
        """delta_u[u1][u2] = event; every state is a key, also pure targets (reward_machine.py:280-292)."""
        out = {}
        for (u1, event), (u2, _r) in self.transitions.items():
            out.setdefault(u1, {})
            out.setdefault(u2, {})
            out[u1][u2] = event
        return out

    def get_delta_r(self):
        """delta_r[u1][u2] = ConstantRewardFunction(reward) (reward_machine.py:294-306)."""
        out = {}
        for (u1, _event), (u2, reward) in self.transitions.items():
            out.setdefault(u1, {})
            out.setdefault(u2, {})
            out[u1][u2] = ConstantRewardFunction(reward)
        return out

    def value_iteration(self, U, delta_u, delta_r, terminal_u, gamma):
        """In-place value iteration over the RM graph (reward_machine.py:308-345); non-constant rewards count as 0."""
        V = {u: 0 for u in U}
        V[terminal_u] = 0
        err = 1
        while err > 0.0000001:
            err = 0
            for u1 in U:
                if not delta_u[u1]:
                    continue
                best = max((delta_r[u1][u2].get_reward(None) if delta_r[u1][u2].get_type() == "constant" else 0) + gamma * V[u2]
                           for u2 in delta_u[u1])
                err = max(err, abs(best - V[u1]))
                V[u1] = best
        return V


class RewardFunction:
    """reward_machine.py:348-357"""

    def get_reward(self, s_info):
        raise NotImplementedError("To be implemented")

    def get_type(self):
        raise NotImplementedError("To be implemented")


class ConstantRewardFunction(RewardFunction):
    """reward_machine.py:360-373"""

    def __init__(self, c):
        super().__init__()
        self.c = c

    def get_type(self):
        return "constant"

    def get_reward(self, s_info):
        return self.c


def builtin_frozen_lake_rm(goals: Dict[str, Tuple[int, int]], detector: Optional[EventDetector] = None) -> RewardMachine:
    """The A -> B -> C machine of frozen_lake_main.py:254-260 (rewards 10 / 15 / 20)."""
    transitions = {
        ("state0", goals["A"]): ("state1", 10),
        ("state1", goals["B"]): ("state2", 15),
        ("state2", goals["C"]): ("state3", 20),
    }
    return RewardMachine(transitions, detector or PositionEventDetector(set(goals.values())))
