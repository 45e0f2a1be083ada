"""Import stub (test infrastructure): names imported but never used by the reference hot path
(/root/reference/multiagent_rlrm/multi_agent/base_environment.py:2)."""
from . import spaces  # noqa: F401
