"""State encoders: enc = (pos_y * W + pos_x) * nQ + q — mirrors multi_agent/state_encoder.py:4-36,
environments/frozen_lake/state_encoder_frozen_lake.py:7-86, environments/office_world/state_encoder_office.py:7-66 and
utils/utils.py:5-35 (encode_state). Index arithmetic on host scalars (configuration-time / N=1 API glue); the batched
kernels compute the same expression per thread."""
from __future__ import annotations


class StateEncoder:
    def __init__(self, agent):
        self.agent = agent

    def encode(self, state, state_rm=None):
        raise NotImplementedError("This method should be overridden by subclasses.")

    def encode_rm_state(self, state_rm):
        rm = self.agent.get_reward_machine()
        return rm.get_state_index(rm.get_current_state() if state_rm is None else state_rm)


class _GridStateEncoder(StateEncoder):
    def _dims(self):
        p = self.agent.ma_problem
        return p.grid_width, p.grid_height, self.agent.get_reward_machine().numbers_state()

    def encode(self, state, state_rm=None):
        agent = self.agent
        p, rm = agent.ma_problem, agent.get_reward_machine()
        width, n_rm = p.grid_width, rm.numbers_state()
        q = rm.get_state_index(rm.get_current_state() if state_rm is None else state_rm)  # encode_rm_state, inlined
        s = state["pos_y"] * width + state["pos_x"]
        enc = s * n_rm + q
        if enc >= width * p.grid_height * n_rm:
            raise ValueError("Encoded state index exceeds total state space size.")
        return enc, {"s": s, "q": q}

    def decode(self, encoded_state):
        width, _height, n_rm = self._dims()
        q, s = encoded_state % n_rm, encoded_state // n_rm
        label = self.agent.get_reward_machine().get_state_from_index(q)
        return {"pos_x": s % width, "pos_y": s // width}, {"q": label}


class StateEncoderFrozenLake(_GridStateEncoder):
    pass


class StateEncoderOfficeWorld(_GridStateEncoder):
    def decode(self, encoded_state):
        """(position dict, RM state LABEL) — state_encoder_office.py:37-66; FrozenLake's decode returns an info dict instead
        (state_encoder_frozen_lake.py:50-86), which is what makes get_mdp degenerate there (see wrapper.get_mdp)."""
        width, height, n_rm = self._dims()
        total = width * height * n_rm
        if encoded_state < 0 or encoded_state >= total:
            raise ValueError(f"Encoded state {encoded_state} out of range [0..{total - 1}].")
        pos, info = super().decode(encoded_state)
        return pos, info["q"]


def encode_state(agent, state, state_reward_machine):
    rm = agent.get_reward_machine()
    n_rm = rm.numbers_state()
    width, height = agent.ma_problem.grid_width, agent.ma_problem.grid_height
    enc = (state["pos_y"] * width + state["pos_x"]) * n_rm + rm.get_state_index(state_reward_machine)
    if enc >= width * height * n_rm:
        raise ValueError("Encoded state index exceeds total state space size", enc, ">=", width * height * n_rm)
    return enc


def _time_dims(agent, state_reward_machine):
    rm = agent.get_reward_machine()
    p = agent.ma_problem
    return p.grid_width, p.grid_height, p.max_time, rm.numbers_state(), rm.get_state_index(state_reward_machine)


def encode_state_with_time(agent, state, state_reward_machine):
    """(position, state["timestamp"], RM state) -> ((y*W + x) * max_time + timestamp) * nQ + q (utils/utils.py:38-75). Time-augmented
    encodings are not used by the reference drivers (nor by the kernels); kept for API parity with the reference's utilities."""
    width, height, max_time, n_rm, q = _time_dims(agent, state_reward_machine)
    enc = ((state["pos_y"] * width + state["pos_x"]) * max_time + state["timestamp"]) * n_rm + q
    if enc >= width * height * max_time * n_rm:
        raise ValueError("Encoded state index exceeds total state space size.")
    return enc


def encode_state_time(agent, state, state_reward_machine):
    """(position, RM state, state["timestep"]) -> ((y*W + x) * nQ + q) * max_time + timestep (utils/utils.py:78-115)."""
    width, height, max_time, n_rm, q = _time_dims(agent, state_reward_machine)
    enc = ((state["pos_y"] * width + state["pos_x"]) * n_rm + q) * max_time + state["timestep"]
    if enc >= width * height * n_rm * max_time:
        raise ValueError("Encoded state index exceeds total state space size.")
    return enc
