"""AgentRL — mirrors multi_agent/agent_rl.py:11-401 (the parts the hot path touches; message passing is not part of
it). The agent is a host-side record (name, position, state dict, RM, encoders, learner); selecting and updating go to
the learner, whose arithmetic runs on the device."""
from __future__ import annotations


class UPValueError(Exception):
    """Raised for an undefined action name (unified_planning.exceptions.UPValueError in the reference, agent_rl.py:55)."""


class AgentRL:
    def __init__(self, name: str, ma_problem, reward_machine=None):
        self._name = name
        self.reward_machine = reward_machine
        self.ma_problem = ma_problem
        self.actions_dict = {}
        self.learning_algorithm = None
        self.message_conditions = None
        self.messages = {}
        self.message_sent = False
        self.position = None
        self.state = {}
        self.actions_ = []
        self.initial_position = {}
        self.initial_state = {}
        self.rm_state = None
        self.next_rm_state = None
        self.encoder = None
        self.action_encoder = None

    @property
    def name(self):
        return self._name

    # -- actions ---------------------------------------------------------------------------------
    def action(self, name: str):
        for a in self.actions_:
            if a.name == name:
                return a
        raise UPValueError(f"Action of name: {name} is not defined!")

    def add_action(self, action):
        self.actions_.append(action)

    add_rl_action = add_action

    def get_actions(self):
        return self.actions_

    def actions_dix(self):
        for idx, act in enumerate(self.actions_):
            self.actions_dict[idx] = act
        return self.actions_dict

    def actions_idx(self, action):
        for key, value in self.actions_dix().items():
            if value == action:
                return key
        return None

    # -- wiring ----------------------------------------------------------------------------------
    def add_state_encoder(self, encoder):
        self.encoder = encoder

    def add_action_encoder(self, encoder):
        self.action_encoder = encoder
        encoder.build_actions()

    def set_reward_machine(self, reward_machine):
        self.reward_machine = reward_machine

    def get_reward_machine(self):
        return self.reward_machine

    def get_reward(self, event):
        return self.reward_machine.get_reward(event) if self.reward_machine else 0

    def set_learning_algorithm(self, algorithm):
        self.learning_algorithm = algorithm

    def get_learning_algorithm(self):
        return self.learning_algorithm

    # -- policy ----------------------------------------------------------------------------------
    def select_action(self, state, best=False):
        if not self.encoder:
            raise Exception("Encoder not set. Please add an encoder before selecting actions.")
        encoded_state, info = self.encoder.encode(state)
        idx = self.get_learning_algorithm().choose_action(encoded_state, best, info=info)
        action = self.actions_dix()[idx]
        if action is None:
            raise ValueError(f"Action index {idx} not found in actions dictionary.")
        return action

    def update_policy(self, state, action, reward, next_state, terminated, **kwargs):
        infos = kwargs.get("infos", {})
        if not self.encoder:
            raise Exception("Encoder not set. Please add an encoder before updating policy.")
        if not self.reward_machine:
            raise Exception("Reward Machine not set. Cannot update policy without Reward Machine.")
        state_rm = infos.get("prev_q", 0)
        next_state_rm = infos.get("q", 0)
        enc_s, cur_info = self.encoder.encode(state, state_rm)
        enc_sn, nxt_info = self.encoder.encode(next_state, next_state_rm)
        rm = self.reward_machine
        info = {
            "prev_s": cur_info["s"],
            "s": nxt_info["s"],
            "prev_q": rm.get_state_index(state_rm) if state_rm != 0 else 0,
            "q": rm.get_state_index(next_state_rm) if next_state_rm != 0 else 0,
            "Renv": infos.get("Renv", 0),
            "RQ": infos.get("RQ", 0),
            "qrm_experience": infos.get("qrm_experience", []),
            "reward_machine": infos.get("reward_machine", []),
        }
        return self.get_learning_algorithm().update(enc_s, enc_sn, self.actions_idx(action), reward, terminated, info=info)

    # -- position / state --------------------------------------------------------------------------
    def set_initial_position(self, pos_x, pos_y):
        self.initial_position = (pos_x, pos_y)
        self.set_position(pos_x, pos_y)

    def set_position(self, pos_x, pos_y):
        self.position = (pos_x, pos_y)
        self.add_to_state("pos_x", pos_x)
        self.add_to_state("pos_y", pos_y)

    def get_position(self):
        return self.position

    def add_to_state(self, key, value):
        self.state[key] = value
        self.initial_state[key] = value

    def reset(self):
        self.set_position(*self.initial_position)
        self.state = self.initial_state.copy()
        self.reset_messages()
        if self.reward_machine:
            self.reward_machine.reset_to_initial_state()

    def reset_messages(self):
        self.messages = {}
        self.message_conditions = None

    def set_state(self, **kwargs):
        for key, value in kwargs.items():
            self.state[key] = value

    def get_state(self):
        return self.state
