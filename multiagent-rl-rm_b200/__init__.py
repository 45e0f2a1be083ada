"""multiagent-rl-rm_b200 — B200-native lockstep hot path of multiagent-rl-rm (env step + RM transition + Q update).

Host side is Python/PyTorch (device memory, streams, torch.distributed); the compute is hand-written CUDA for
sm_100a behind the C ABI in include/rlrm_b200.h (csrc/rlrm_b200.cu). There is no CPU fallback.
"""
from . import _abi as abi  # noqa: F401
from .maps import (  # noqa: F401
    frozen_lake_grid,
    office_world_grid,
    parse_map_emoji,
    parse_office_world,
)
from .reward_machine import PositionEventDetector, RewardMachine  # noqa: F401
from .tables import (  # noqa: F401
    Scenario,
    compile_scenario,
    scenario_config1,
    scenario_config2,
    scenario_config3,
    scenario_config4,
    scenario_config5,
)

__version__ = "0.1.0"
from .actions import ActionEncoderFrozenLake, ActionEncoderOfficeWorld, ActionRL  # noqa: E402,F401
from .agent import AgentRL, UPValueError  # noqa: E402,F401
from .encoders import StateEncoderFrozenLake, StateEncoderOfficeWorld, encode_state  # noqa: E402,F401
from .envs import MultiAgentFrozenLake, MultiAgentOfficeWorld  # noqa: E402,F401
from .learners import QLearning, QLearningLambda  # noqa: E402,F401
from .wrapper import RMEnvironmentWrapper  # noqa: E402,F401
from . import rmspec  # noqa: E402,F401
from .evaluation import (extract_policy_from_qtable, load_q_tables, load_q_tables_into, save_q_tables,  # noqa: E402,F401
                         test_policy_opt_multi_batched, test_policy_optima_batched)
from .rmspec import compile_reward_machine, load_reward_machine, load_rmspec  # noqa: E402,F401
from .vec import BatchedRMEnvironment  # noqa: E402,F401
from .experiments import OPTIMAL, get_experiment_for_map, scenario_for_experiment  # noqa: E402,F401
from .mdp_vi import mdp_to_arrays, value_iteration, value_iteration_arrays  # noqa: E402,F401
