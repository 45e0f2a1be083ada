class Discrete:
    def __init__(self, n=0, *a, **k):
        self.n = n


class MultiDiscrete:
    def __init__(self, nvec=(), *a, **k):
        self.nvec = nvec
