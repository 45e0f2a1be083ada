"""Symbolic actions and action encoders — mirrors multi_agent/action_rl.py:1-28, multi_agent/action_encoder.py:4-20,
environments/frozen_lake/action_encoder_frozen_lake.py:7-17, environments/office_world/action_encoder_office_world.py."""
from __future__ import annotations

ACTION_ORDER = ("up", "down", "left", "right")  # index order fixed by the encoders above


class ActionRL:
    def __init__(self, name, preconditions=None, effects=None):
        self.name = name
        self.preconditions = self._as_list(preconditions)
        self.effects = self._as_list(effects)

    @staticmethod
    def _as_list(x):
        if x is None:
            return []
        return list(x) if isinstance(x, (list, tuple)) else [x]

    def __repr__(self):
        return f"ActionRL({self.name!r})"


class ActionEncoder:
    def __init__(self, agent):
        self.agent = agent

    def build_actions(self):
        raise NotImplementedError

    @property
    def action_names(self):
        return [a.name for a in self.agent.actions_]


class _FourMoves(ActionEncoder):
    def build_actions(self):
        for name in ACTION_ORDER:
            self.agent.add_action(ActionRL(name))


class ActionEncoderFrozenLake(_FourMoves):
    pass


class ActionEncoderOfficeWorld(_FourMoves):
    pass
