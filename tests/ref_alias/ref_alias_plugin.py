"""pytest plugin (container only): serves THIS repo's classes under the reference's module names, so the reference's own test
files can be run unchanged against the drop-in layer. Only the host-side modules those tests import are
aliased; nothing of the reference package itself is importable while the plugin is active (its directory is not on sys.path)."""
import sys
import types

import multiagent_rlrm_b200 as P
from multiagent_rlrm_b200 import actions, agent, encoders, maps, reward_machine, rmspec
from multiagent_rlrm_b200.experiments import get_experiment_for_map


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behaves as a package for `import a.b.c`
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def _office_config():
    out = {"maps": {}}
    for name, (rows, grid_size, start) in maps.OFFICE_WORLD_MAPS.items():
        out["maps"][name] = {"layout": maps.emoji_from_ascii_rows(rows), "grid_size": grid_size, "agents": [{"name": "a1", "position": start}]}
    return out


def pytest_configure(config):
    if "multiagent_rlrm" in sys.modules:
        raise RuntimeError("the reference package is already imported; the alias plugin must come first")
    _module("multiagent_rlrm")
    _module("multiagent_rlrm.multi_agent")
    _module("multiagent_rlrm.multi_agent.reward_machine", RewardMachine=reward_machine.RewardMachine,
            ConstantRewardFunction=reward_machine.ConstantRewardFunction, RewardFunction=reward_machine.RewardFunction)
    _module("multiagent_rlrm.multi_agent.agent_rl", AgentRL=agent.AgentRL)
    _module("multiagent_rlrm.multi_agent.action_rl", ActionRL=actions.ActionRL)
    _module("multiagent_rlrm.multi_agent.state_encoder", StateEncoder=encoders.StateEncoder)
    _module("multiagent_rlrm.multi_agent.action_encoder", ActionEncoder=actions.ActionEncoder)
    _module("multiagent_rlrm.multi_agent.event_detector", EventDetector=reward_machine.EventDetector)
    _module("multiagent_rlrm.rmgen")
    _module("multiagent_rlrm.rmgen.spec", RMSpec=rmspec.RMSpec, TransitionSpec=rmspec.TransitionSpec)
    _module("multiagent_rlrm.rmgen.io", load_rmspec=rmspec.load_rmspec, compile_reward_machine=rmspec.compile_reward_machine)
    _module("multiagent_rlrm.rmgen.normalize", normalize_rmspec_events=rmspec.normalize_rmspec_events, enforce_env_id=rmspec.enforce_env_id,
            autofix_rmspec_states_for_officeworld=rmspec.autofix_rmspec_states_for_officeworld,
            autofix_terminal_reward_violations_for_officeworld=rmspec.autofix_terminal_reward_violations_for_officeworld)
    _module("multiagent_rlrm.rmgen.completion", complete_missing_transitions=rmspec.complete_missing_transitions)
    _module("multiagent_rlrm.rmgen.validator", ValidationError=rmspec.ValidationError, validate_spec=rmspec.validate_spec,
            validate_schema=rmspec.validate_schema, ensure_deterministic=rmspec.ensure_deterministic, validate_semantics=rmspec.validate_semantics)
    _module("multiagent_rlrm.rmgen.summary", format_rmspec_summary=rmspec.format_rmspec_summary)
    _module("multiagent_rlrm.rmgen.exporter", build_reward_machine=rmspec.build_reward_machine, export_spec_to_file=rmspec.export_spec_to_file,
            PassthroughEventDetector=rmspec.PassthroughEventDetector)
    _module("multiagent_rlrm.utils")
    _module("multiagent_rlrm.utils.utils", parse_office_world=maps.parse_office_world, parse_map_emoji=maps.parse_map_emoji,
            encode_state=encoders.encode_state, encode_state_with_time=encoders.encode_state_with_time,
            encode_state_time=encoders.encode_state_time)
    _module("multiagent_rlrm.environments")
    _module("multiagent_rlrm.environments.frozen_lake")
    _module("multiagent_rlrm.environments.frozen_lake.state_encoder_frozen_lake", StateEncoderFrozenLake=encoders.StateEncoderFrozenLake)
    _module("multiagent_rlrm.environments.frozen_lake.action_encoder_frozen_lake", ActionEncoderFrozenLake=actions.ActionEncoderFrozenLake)
    _module("multiagent_rlrm.environments.frozen_lake.detect_event", PositionEventDetector=reward_machine.PositionEventDetector)
    _module("multiagent_rlrm.environments.frozen_lake.event_context", build_frozenlake_context=rmspec.build_frozenlake_context)
    _module("multiagent_rlrm.environments.office_world")
    _module("multiagent_rlrm.environments.office_world.detect_event", PositionEventDetector=reward_machine.PositionEventDetector)
    _module("multiagent_rlrm.environments.office_world.event_context", build_officeworld_context=rmspec.build_officeworld_context)
    _module("multiagent_rlrm.environments.office_world.config_office", config=_office_config(), get_experiment_for_map=get_experiment_for_map)
    _module("unified_planning")
    _module("unified_planning.exceptions", UPValueError=agent.UPValueError)
