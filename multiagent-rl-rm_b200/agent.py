"""AgentRL — the host-side agent record of the drop-in API (reference: multi_agent/agent_rl.py:11-401, the parts the hot
path touches; message passing is not part of it). Attribute and method names are the reference's because the boundary is
duck-typed — its own tests poke `_name`, `actions_`, `actions_dix()` — but the agent holds no arithmetic: `select_action` and
`update_policy` only encode states and forward to the learner, whose selection / update kernels run on the device."""
from __future__ import annotations


class UPValueError(Exception):
    """Raised for an undefined action name (unified_planning.exceptions.UPValueError in the reference, agent_rl.py:55)."""


def _plain_getter(attr):
    """Accessor method `lambda self: self.<attr>` compiled with the attribute name in its bytecode (no getattr call: the
    accessors are called ~100 times per driver-loop iteration of the one-instance classes)."""
    scope = {}
    exec(f"def get(self):\n    return self.{attr}", scope)
    return scope["get"]


def _plain_setter(attr):
    scope = {}
    exec(f"def put(self, value):\n    self.{attr} = value", scope)
    return scope["put"]


class AgentRL:
    # attribute -> factory of its initial value (agent_rl.py:19-40 lists the same attributes one by one)
    _INITIAL = (("actions_dict", dict), ("learning_algorithm", type(None)), ("message_conditions", type(None)), ("messages", dict),
                ("message_sent", bool), ("position", type(None)), ("state", dict), ("actions_", list), ("initial_position", dict),
                ("initial_state", dict), ("rm_state", type(None)), ("next_rm_state", type(None)), ("encoder", type(None)),
                ("action_encoder", type(None)))

    def __init__(self, name: str, ma_problem, reward_machine=None):
        self._name, self.ma_problem, self.reward_machine = name, ma_problem, reward_machine
        for attr, factory in self._INITIAL:
            setattr(self, attr, factory())

    name = property(_plain_getter("_name"))

    # -- actions -----------------------------------------------------------------------------------------------------
    def action(self, name: str):
        found = next((a for a in self.actions_ if a.name == name), None)
        if found is None:
            raise UPValueError(f"Action of name: {name} is not defined!")
        return found

    def add_action(self, action):
        self.actions_.append(action)

    add_rl_action = add_action
    get_actions = _plain_getter("actions_")

    def actions_dix(self):
        """index -> action; entries are (re)written on every call and never dropped, like the reference's dict."""
        self.actions_dict.update(enumerate(self.actions_))
        return self.actions_dict

    def actions_idx(self, action):
        """index of `action` in actions_dix() (first match, None when absent) — looked up through a reverse map that is rebuilt
        whenever the action list or the index dict changed"""
        acts, known = self.actions_, self.actions_dict
        cache = self.__dict__.get("_idx_cache")
        if cache is None or cache[0] != tuple(acts) or cache[1] != len(known):
            table = {}
            for k, v in self.actions_dix().items():
                try:
                    table.setdefault(v, k)
                except TypeError:  # unhashable action objects: keep the linear scan
                    table = None
                    break
            cache = self.__dict__["_idx_cache"] = (tuple(acts), len(self.actions_dict), None, table)
        if cache[3] is not None:
            try:
                return cache[3].get(action)
            except TypeError:
                pass
        return next((k for k, v in self.actions_dix().items() if v == action), None)

    # -- wiring ------------------------------------------------------------------------------------------------------
    add_state_encoder = _plain_setter("encoder")
    set_reward_machine = _plain_setter("reward_machine")
    get_reward_machine = _plain_getter("reward_machine")
    set_learning_algorithm = _plain_setter("learning_algorithm")
    get_learning_algorithm = _plain_getter("learning_algorithm")

    def add_action_encoder(self, encoder):
        self.action_encoder = encoder
        encoder.build_actions()

    def get_reward(self, event):
        rm = self.reward_machine
        return rm.get_reward(event) if rm else 0

    # -- policy: encode, then hand over to the learner (device) ---------------------------------------------------------
    def _need(self, what, present, doing):
        if not present:
            raise Exception(f"{what} not set. " + doing)

    def select_action(self, state, best=False):
        self._need("Encoder", self.encoder, "Please add an encoder before selecting actions.")
        enc, info = self.encoder.encode(state)
        idx = self.learning_algorithm.choose_action(enc, best, info=info)
        chosen = self.actions_dix()[idx]
        if chosen is None:
            raise ValueError(f"Action index {idx} not found in actions dictionary.")
        return chosen

    def update_policy(self, state, action, reward, next_state, terminated, **kwargs):
        step = kwargs.get("infos", {})
        self._need("Encoder", self.encoder, "Please add an encoder before updating policy.")
        self._need("Reward Machine", self.reward_machine, "Cannot update policy without Reward Machine.")
        rm, q_before, q_after = self.reward_machine, step.get("prev_q", 0), step.get("q", 0)
        (enc_s, at_s), (enc_n, at_n) = self.encoder.encode(state, q_before), self.encoder.encode(next_state, q_after)
        # what QLearning.update reads from `info` (agent_rl.py:158-172): cell indices, RM state INDICES, the two reward parts,
        # the counterfactual list and the machine itself; a missing RM label (0) stays 0
        info = dict(prev_s=at_s["s"], s=at_n["s"],
                    prev_q=rm.get_state_index(q_before) if q_before != 0 else 0,
                    q=rm.get_state_index(q_after) if q_after != 0 else 0,
                    Renv=step.get("Renv", 0), RQ=step.get("RQ", 0),
                    qrm_experience=step.get("qrm_experience", []), reward_machine=step.get("reward_machine", []))
        return self.learning_algorithm.update(enc_s, enc_n, self.actions_idx(action), reward, terminated, info=info)

    # -- position / state ----------------------------------------------------------------------------------------------
    get_position = _plain_getter("position")
    get_state = _plain_getter("state")

    def set_position(self, pos_x, pos_y):
        self.position = (pos_x, pos_y)
        for key, value in (("pos_x", pos_x), ("pos_y", pos_y)):
            self.add_to_state(key, value)

    def set_initial_position(self, pos_x, pos_y):
        self.initial_position = (pos_x, pos_y)
        self.set_position(pos_x, pos_y)

    def add_to_state(self, key, value):
        self.state[key] = self.initial_state[key] = value

    def set_state(self, **kwargs):
        self.state.update(kwargs)

    def reset_messages(self):
        self.messages, self.message_conditions = {}, None

    def reset(self):
        """Back to the initial position and state; the reward machine returns to its initial state (agent_rl.py:372-384)."""
        self.set_position(*self.initial_position)
        self.state = dict(self.initial_state)
        self.reset_messages()
        if self.reward_machine:
            self.reward_machine.reset_to_initial_state()
