__all__ = ["Agent", "MultiAgentProblem"]


class Agent:
    def __init__(self, name, ma_problem):
        self._name = name
        self._ma_problem = ma_problem

    @property
    def name(self):
        return self._name


class MultiAgentProblem:
    def __init__(self, *args, **kwargs):
        self._agents = []

    @property
    def agents(self):
        return self._agents

    def add_agent(self, agent):
        self._agents.append(agent)
