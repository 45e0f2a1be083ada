"""Reward-Machine specification files -> device tables, host side.

The on-disk input format of the hot path: a JSON/YAML RM spec plus a map name is validated, its event names are
normalised against the map ("A", "at( a )", "office" -> "at(A)", "at(A)", "at(O)"), optionally completed with
self-loops, expanded to grid positions and compiled into the ``RewardMachine`` whose ``compile_tables`` feeds the
kernels — without any reference object. Behaviour mirrors, function by function:

  spec / reward coercion       /root/reference/multiagent_rlrm/rmgen/spec.py:6-98
  load_rmspec                  rmgen/io.py:39-77
  validate_spec / semantics    rmgen/validator.py:18-103
  complete_missing_transitions rmgen/completion.py:6-38
  normalisation + guardrails   rmgen/normalize.py:9-353
  map contexts                 environments/office_world/event_context.py:10-74, frozen_lake/event_context.py:10-49
  compile_reward_machine       rmgen/io.py:80-171
  runner pipelines             office_world/office_main.py:446-515, frozen_lake/frozen_lake_main.py:125-183

The LLM providers and the rmgen CLI are out of scope.
"""
from __future__ import annotations

import dataclasses
import json
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple, Union

from .maps import frozen_lake_grid, office_world_grid
from .reward_machine import EventDetector, PositionEventDetector, RewardMachine


class ValidationError(Exception):
    """Raised when an RMSpec is invalid."""


class UnknownEventError(ValidationError):
    """An event of the spec is not one the selected map can produce."""

    def __init__(self, *, event, allowed_events, map_name=None, hint=None):
        self.event, self.allowed_events, self.map_name, self.hint = event, list(allowed_events), map_name, hint
        shown = sorted(self.allowed_events)
        more = ", ..." if len(shown) > 25 else ""
        where = f" (map={map_name})" if map_name else ""
        tip = f" Hint: {hint}" if hint else ""
        super().__init__(f"Unknown event '{event}'{where}. Allowed events: {', '.join(shown[:25])}{more}.{tip}")


# ---------------------------------------------------------------------------------------------- spec
def _reward_value(raw, where) -> float:
    """Numbers, numeric strings, or strings with a leading 'r' ("r0.5")."""
    if isinstance(raw, (int, float)):
        return float(raw)
    if isinstance(raw, str):
        text = raw.strip()
        if text[:1].lower() == "r":
            text = text[1:]
        try:
            return float(text)
        except ValueError as exc:
            raise ValueError(f"Invalid reward value '{raw}' in transition {where}") from exc
    raise ValueError(f"Invalid reward type '{type(raw)}' in transition {where}")


@dataclass
class TransitionSpec:
    from_state: str
    event: str
    to_state: str
    reward: float

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "TransitionSpec":
        return cls(data["from_state"], data["event"], data["to_state"], _reward_value(data["reward"], data))

    def to_dict(self):
        return dataclasses.asdict(self)


@dataclass
class RMSpec:
    name: str
    env_id: str
    version: str
    states: List[str]
    initial_state: str
    terminal_states: List[str]
    event_vocabulary: List[str]
    transitions: List[TransitionSpec]
    notes: Optional[str] = None

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "RMSpec":
        return cls(
            name=data["name"], env_id=data["env_id"].strip().lower(), version=data["version"], states=list(data["states"]),
            initial_state=data["initial_state"], terminal_states=list(data.get("terminal_states", [])),
            event_vocabulary=list(data["event_vocabulary"]),
            transitions=[TransitionSpec.from_dict(t) for t in data.get("transitions", [])], notes=data.get("notes"))

    def to_dict(self):
        d = dataclasses.asdict(self)
        d["transitions"] = [t.to_dict() for t in self.transitions]
        return d

    def as_transition_map(self):
        return {(t.from_state, t.event): (t.to_state, t.reward) for t in self.transitions}


def load_rmspec(path: Union[str, Path]) -> RMSpec:
    src = Path(path)
    if not src.exists():
        raise FileNotFoundError(f"RM spec file not found: {src}")
    if not src.is_file():
        raise ValueError(f"RM spec path is not a file: {src}")
    text = src.read_text(encoding="utf-8")
    try:
        data = json.loads(text)
    except json.JSONDecodeError as json_exc:
        if src.suffix.lower() == ".json":
            raise ValueError(f"Invalid JSON in {src}: {json_exc}") from json_exc
        try:
            import yaml
        except ModuleNotFoundError as exc:
            raise ValueError(f"YAML support requires PyYAML, or provide a JSON spec instead: {src}") from exc
        try:
            data = yaml.safe_load(text)
        except Exception as exc:
            raise ValueError(f"Invalid YAML in {src}: {exc}") from exc
    if not isinstance(data, dict):
        raise ValueError(f"RM spec must be a JSON/YAML object at top-level: {src}")
    try:
        return RMSpec.from_dict(data)
    except KeyError as exc:
        raise ValueError(f"RM spec missing required field {exc!s}: {src}") from exc
    except ValueError as exc:
        raise ValueError(f"RM spec has invalid values: {src}: {exc}") from exc


# ---------------------------------------------------------------------------------------------- validation
def _no_duplicates(items, label):
    seen = set()
    for it in items:
        if it in seen:
            raise ValidationError(f"Duplicate {label}: {it}")
        seen.add(it)


def validate_schema(spec: RMSpec) -> None:
    """Structural checks (rmgen/validator.py:18-57)."""
    if not spec.states:
        raise ValidationError("states must be non-empty")
    _no_duplicates(spec.states, "state")
    if spec.initial_state not in spec.states:
        raise ValidationError(f"initial_state {spec.initial_state} not in states")
    for s in spec.terminal_states:
        if s not in spec.states:
            raise ValidationError(f"terminal_state {s} not in states")
    _no_duplicates(spec.terminal_states, "terminal_state")
    if not spec.event_vocabulary:
        raise ValidationError("event_vocabulary must be non-empty")
    _no_duplicates(spec.event_vocabulary, "event")
    if not spec.transitions:
        raise ValidationError("transitions must be non-empty")
    states, events = set(spec.states), set(spec.event_vocabulary)
    for t in spec.transitions:
        if t.from_state not in states:
            raise ValidationError(f"transition from_state {t.from_state} not in states")
        if t.to_state not in states:
            raise ValidationError(f"transition to_state {t.to_state} not in states")
        if t.event not in events:
            raise ValidationError(f"transition event {t.event} not in vocabulary")
    if not any(spec.initial_state in (t.from_state, t.to_state) for t in spec.transitions):
        raise ValidationError(f"initial_state {spec.initial_state} has no incident transitions")


def ensure_deterministic(spec: RMSpec) -> None:
    """One target per (state, event) (rmgen/validator.py:60-69)."""
    target = {}
    for t in spec.transitions:
        key = (t.from_state, t.event)
        if target.setdefault(key, t.to_state) != t.to_state:
            raise ValidationError(f"Non-deterministic transitions for {key}: {target[key]} vs {t.to_state}")


def validate_spec(spec: RMSpec) -> None:
    """All validations; raises ValidationError on the first failure (rmgen/validator.py:72-77)."""
    validate_schema(spec)
    ensure_deterministic(spec)


def validate_semantics(spec: RMSpec, *, max_positive_reward_transitions=None, terminal_reward_must_be_zero=True) -> None:
    if max_positive_reward_transitions is not None:
        positive = [t for t in spec.transitions if t.reward > 0]
        if len(positive) > max_positive_reward_transitions:
            raise ValidationError(f"Positive-reward transitions exceed limit {max_positive_reward_transitions}: {positive}")
    if terminal_reward_must_be_zero:
        terminal = set(spec.terminal_states)
        bad = [t for t in spec.transitions if t.from_state in terminal and t.reward != 0]
        if bad:
            raise ValidationError(f"Terminal transitions must have reward 0. Offenders: {bad}")


def complete_missing_transitions(spec: RMSpec, default_reward: float = 0.0, terminal_self_loop: bool = True):
    """Full (state x event) cartesian product: every missing pair becomes a self-loop with `default_reward`.
    Appended in spec.states x spec.event_vocabulary order (this fixes which transition is inserted LAST, i.e. the
    compiled machine's final state)."""
    if not spec.states or not spec.event_vocabulary:
        return spec, {"added": 0}
    have = {(t.from_state, t.event) for t in spec.transitions}
    extra = []
    for s in spec.states:
        if s in spec.terminal_states and not terminal_self_loop:
            continue
        for ev in spec.event_vocabulary:
            if (s, ev) not in have:
                extra.append(TransitionSpec(s, ev, s, default_reward))
    spec.transitions.extend(extra)
    return spec, {"added": len(extra)}


# ---------------------------------------------------------------------------------------------- normalisation
def normalize_event_key(raw) -> str:
    return "" if raw is None else "".join(str(raw).split()).lower()


def enforce_env_id(spec: RMSpec, expected_env_id: str, *, reason: str) -> RMSpec:
    expected = str(expected_env_id).strip().lower()
    if (spec.env_id or "").strip().lower() != expected:
        print(f"Warning: overriding env_id from '{spec.env_id if spec.env_id else '<missing>'}' to '{expected}' because {reason}")
        spec.env_id = expected
    return spec


def _canonical(event: str, context: Mapping[str, object]) -> str:
    allowed, cmap = context.get("allowed_events"), context.get("canonical_map")
    if not isinstance(allowed, Sequence) or not allowed:
        raise ValueError("context['allowed_events'] must be a non-empty sequence")
    if not isinstance(cmap, Mapping) or not cmap:
        raise ValueError("context['canonical_map'] must be a non-empty mapping")
    map_name = context.get("map_name")
    key = normalize_event_key(event)
    if key not in cmap:
        env = str(context.get("env_id")).strip().lower() if context.get("env_id") is not None else "<context>"
        raise UnknownEventError(event=event, allowed_events=list(allowed), map_name=str(map_name) if map_name else None,
                                hint=f"Use `rmgen --context {env} --map {map_name or '<map>'}` to autoprompt allowed events.")
    canon = str(cmap[key])
    if canon not in set(allowed):
        raise UnknownEventError(event=event, allowed_events=list(allowed), map_name=str(map_name) if map_name else None)
    return canon


def normalize_rmspec_events(rmspec, context):
    """Canonicalise every vocabulary entry and transition event ('at(X)' form); accepts an RMSpec or a raw dict."""
    is_dict = isinstance(rmspec, dict)
    if not is_dict and not isinstance(rmspec, RMSpec):
        raise TypeError(f"Unsupported RMSpec type: {type(rmspec)}")
    vocab_in = rmspec.get("event_vocabulary", []) if is_dict else rmspec.event_vocabulary
    vocab, seen = [], set()
    for ev in vocab_in:
        c = _canonical(str(ev), context)
        if c not in seen:
            seen.add(c)
            vocab.append(c)
    if is_dict:
        transitions = []
        for t in rmspec.get("transitions", []):
            t2 = dict(t)
            t2["event"] = _canonical(str(t2.get("event", "")), context)
            transitions.append(t2)
        events = [t["event"] for t in transitions]
    else:
        transitions, dedup = [], set()
        for t in rmspec.transitions:
            t2 = TransitionSpec(t.from_state, _canonical(t.event, context), t.to_state, t.reward)
            key = (t2.from_state, t2.event, t2.to_state, t2.reward)
            if key not in dedup:
                dedup.add(key)
                transitions.append(t2)
        events = [t.event for t in transitions]
    for ev in events:  # the vocabulary must cover every transition event
        if ev not in seen:
            seen.add(ev)
            vocab.append(ev)
    if is_dict:
        out = dict(rmspec)
        out["event_vocabulary"], out["transitions"] = vocab, transitions
        return out
    return dataclasses.replace(rmspec, event_vocabulary=vocab, transitions=transitions)


def autofix_rmspec_states_for_officeworld(spec: RMSpec) -> RMSpec:
    """Add states that transitions reference but `states` omits; a missing state that is only ever entered with a
    positive reward and never left is also declared terminal."""
    known = set(spec.states)
    missing = sorted({s for t in spec.transitions for s in (t.from_state, t.to_state)} - known)
    if not missing:
        return spec
    states, terminals = list(spec.states) + missing, list(spec.terminal_states)
    for s in missing:
        leaves = any(t.from_state == s for t in spec.transitions)
        entered_pos = any(t.to_state == s and t.reward > 0 for t in spec.transitions)
        entered_nonpos = any(t.to_state == s and t.reward <= 0 for t in spec.transitions)
        if not leaves and entered_pos and not entered_nonpos and s not in terminals:
            terminals.append(s)
    spec.states, spec.terminal_states = states, terminals
    return spec


def _fresh_state_name(existing) -> str:
    nums = [int(m.group(1)) for m in (re.fullmatch(r"q(\d+)", str(s)) for s in existing) if m]
    if nums and f"q{max(nums) + 1}" not in existing:
        return f"q{max(nums) + 1}"
    if "q_terminal" not in existing:
        return "q_terminal"
    i = 1
    while f"q_terminal_{i}" in existing:
        i += 1
    return f"q_terminal_{i}"


def autofix_terminal_reward_violations_for_officeworld(spec: RMSpec) -> RMSpec:
    """A terminal state with a non-zero outgoing reward stops being terminal; if the offender is a self-loop it is
    redirected into a fresh terminal sink so the reward is collected once."""
    terminal = set(spec.terminal_states)
    if not any(t.from_state in terminal and t.reward != 0 for t in spec.transitions):
        return spec
    names = set(spec.states)
    states, transitions = list(spec.states), []
    demoted, sinks, new_sinks = set(), {}, []
    for t in spec.transitions:
        if t.from_state not in terminal or t.reward == 0:
            transitions.append(t)
            continue
        demoted.add(t.from_state)
        if t.to_state != t.from_state:
            print(f"Warning: removing state '{t.from_state}' from terminal_states because it has a non-zero outgoing reward "
                  f"transition (reward={t.reward})")
            transitions.append(t)
            continue
        sink = sinks.get(t.from_state)
        if sink is None:
            sink = sinks[t.from_state] = _fresh_state_name(names)
            names.add(sink)
            states.append(sink)
            new_sinks.append(sink)
            print("Warning: rewrote a terminal self-loop with non-zero reward "
                  f"({t.from_state} --{t.event}--> {t.to_state}, reward={t.reward}) to transition into a new terminal state '{sink}'")
        transitions.append(TransitionSpec(t.from_state, t.event, sink, t.reward))
    terminals = [s for s in spec.terminal_states if s not in demoted]
    terminals += [s for s in sorted(new_sinks) if s not in terminals]
    spec.states, spec.terminal_states, spec.transitions = states, terminals, transitions
    return spec


# ---------------------------------------------------------------------------------------------- map contexts
def _context(env_id, map_name, symbols, aliases):
    allowed, cmap = set(), {}
    for alias, canon in aliases:
        allowed.add(alias)
        cmap[normalize_event_key(alias)] = canon
    return {"env_id": env_id, "map_name": map_name, "allowed_symbols": symbols, "allowed_events": sorted(allowed),
            "canonical_map": cmap}


def build_officeworld_context(map_name: str):
    from .maps import OFFICE_WORLD_MAPS

    if map_name not in OFFICE_WORLD_MAPS:
        raise ValueError(f"Unknown OfficeWorld map '{map_name}'. Available: {sorted(OFFICE_WORLD_MAPS)}")
    g = office_world_grid(map_name)
    symbols, aliases = set(g.goals), []
    for sym in sorted(g.goals):
        aliases += [(sym, f"at({sym})"), (f"at({sym})", f"at({sym})")]
    if "O" in g.goals:
        symbols.add("office")
        aliases += [("office", "at(O)"), ("at(office)", "at(O)")]
    if g.coffee:
        symbols.add("coffee")
        aliases += [("coffee", "at(coffee)"), ("at(coffee)", "at(coffee)")]
    if g.letters:
        symbols |= {"letter", "email"}
        aliases += [(a, "at(letter)") for a in ("letter", "email", "at(letter)", "at(email)")]
    return _context("officeworld", map_name, symbols, aliases)


def build_frozenlake_context(map_name: str):
    from .maps import FROZEN_LAKE_MAPS

    if map_name not in FROZEN_LAKE_MAPS:
        raise ValueError(f"Unknown FrozenLake map '{map_name}'. Available: {sorted(FROZEN_LAKE_MAPS)}")
    g = frozen_lake_grid(map_name)
    aliases = []
    for sym in sorted(g.goals):
        aliases += [(sym, f"at({sym})"), (f"at({sym})", f"at({sym})")]
    return _context("frozenlake", map_name, set(g.goals), aliases)


def officeworld_event_mapping(map_name: str) -> Dict[str, object]:
    """Spec event -> grid position(s), office_main.py:446-473."""
    g = office_world_grid(map_name)
    m: Dict[str, object] = {}
    for label, pos in g.goals.items():
        m[f"at({label})"] = pos
        m[label] = pos
    if "O" in g.goals:
        m["office"] = m["at(office)"] = g.goals["O"]
    if g.coffee:
        m["coffee"], m["at(coffee)"] = list(g.coffee), list(g.coffee)
    if g.letters:
        for k in ("letter", "email", "at(letter)", "at(email)"):
            m[k] = list(g.letters)
    return m


def frozenlake_event_mapping(map_name: str) -> Dict[str, object]:
    m: Dict[str, object] = {}
    for label, pos in frozen_lake_grid(map_name).goals.items():
        m[f"at({label})"] = pos
        m[label] = pos
    return m


# ---------------------------------------------------------------------------------------------- compilation
class PassthroughEventDetector(EventDetector):
    """Returns state["event"] when it is in the vocabulary (rmgen/exporter.py:11-26)."""

    def __init__(self, allowed_events):
        self.allowed_events = set(allowed_events)

    def detect_event(self, current_state):
        ev = current_state.get("event") if isinstance(current_state, dict) else None
        return ev if ev in self.allowed_events else None


def _transition_map(spec: RMSpec, event_mapping):
    if not event_mapping:
        return spec.as_transition_map()
    out = {}
    for t in spec.transitions:
        if t.event not in event_mapping:
            raise ValueError(f"Unknown event '{t.event}' in RMSpec; missing from event mapping (available: {sorted(event_mapping.keys())})")
        mapped = event_mapping[t.event]
        if mapped is None:
            raise ValueError(f"Event mapping for '{t.event}' is None")
        targets = list(mapped) if isinstance(mapped, (list, set, frozenset)) else [mapped]
        if not targets:
            raise ValueError(f"Event mapping for '{t.event}' is empty")
        for ev in targets:
            try:
                hash(ev)
            except TypeError as exc:
                raise ValueError(f"Mapped event for '{t.event}' is not hashable: {ev!r}") from exc
            key, value = (t.from_state, ev), (t.to_state, t.reward)
            if out.get(key, value) != value:
                raise ValueError(f"Event mapping produced conflicting transitions for {key}: {out[key]} vs {value}")
            out[key] = value
    return out


def compile_reward_machine(spec: RMSpec, *, event_detector=None, event_mapping=None, complete_missing_transitions: bool = False,
                           default_reward: float = 0.0, terminal_self_loop: bool = True, max_positive_reward_transitions=None,
                           terminal_reward_must_be_zero: bool = True) -> RewardMachine:
    if complete_missing_transitions:
        spec, _ = globals()["complete_missing_transitions"](spec, default_reward=default_reward, terminal_self_loop=terminal_self_loop)
    validate_spec(spec)
    validate_semantics(spec, max_positive_reward_transitions=max_positive_reward_transitions,
                       terminal_reward_must_be_zero=terminal_reward_must_be_zero)
    rm = RewardMachine(_transition_map(spec, event_mapping), event_detector or PassthroughEventDetector(spec.event_vocabulary))
    rm.initial_state = rm.current_state = spec.initial_state  # the spec, not the first transition, names the start state
    rm.state_indices = rm._generate_state_indices()
    return rm


def load_reward_machine(path, env: str, map_name: str = "map1", *, complete_missing_transitions: bool = False,
                        default_reward: float = 0.0, terminal_self_loop: bool = True, max_positive_reward_transitions=None,
                        terminal_reward_must_be_zero: bool = True) -> Tuple[RewardMachine, RMSpec]:
    """The runners' `--rm-spec` pipeline: office_main.py:487-515 (env="office_world") / frozen_lake_main.py:133-183."""
    spec = load_rmspec(path)
    if env == "office_world":
        spec = enforce_env_id(spec, "officeworld", reason="--rm-spec is set")
        spec = normalize_rmspec_events(spec, build_officeworld_context(map_name))
        spec = autofix_rmspec_states_for_officeworld(spec)
        if terminal_reward_must_be_zero:
            spec = autofix_terminal_reward_violations_for_officeworld(spec)
        mapping = officeworld_event_mapping(map_name)
        g = office_world_grid(map_name)
        positions = set(g.goals.values()) | set(g.coffee) | set(g.letters)  # the map's position_map (config_office.py:103-113)
    else:
        spec = enforce_env_id(spec, "frozenlake", reason="--rm-spec is set")
        spec = normalize_rmspec_events(spec, build_frozenlake_context(map_name))
        mapping = frozenlake_event_mapping(map_name)
        positions = set(frozen_lake_grid(map_name).goals.values())
    for target in mapping.values():
        positions.update(target if isinstance(target, (list, set, frozenset)) else [target])
    rm = compile_reward_machine(spec, event_detector=PositionEventDetector(positions), event_mapping=mapping,
                                complete_missing_transitions=complete_missing_transitions, default_reward=default_reward,
                                terminal_self_loop=terminal_self_loop, max_positive_reward_transitions=max_positive_reward_transitions,
                                terminal_reward_must_be_zero=terminal_reward_must_be_zero)
    return rm, spec


def export_spec_to_file(spec: RMSpec, output_path) -> None:
    """Write the spec back in the on-disk JSON format load_rmspec reads (rmgen/exporter.py:28-31): the output side of f1."""
    out = Path(output_path)
    out.parent.mkdir(parents=True, exist_ok=True)
    with out.open("w", encoding="utf-8") as f:
        json.dump(spec.to_dict(), f, indent=2, ensure_ascii=True)


def build_reward_machine(spec: RMSpec, event_detector=None):
    """RewardMachine over the spec's own event names (rmgen/exporter.py:34-50): no position mapping, the default detector
    passes the ``event`` field of the state dict through. Initial state and index map follow the spec's initial_state."""
    from .reward_machine import RewardMachine

    rm = RewardMachine(spec.as_transition_map(), event_detector or PassthroughEventDetector(spec.event_vocabulary))
    rm.initial_state = spec.initial_state
    rm.current_state = spec.initial_state
    rm.state_indices = rm._generate_state_indices()
    return rm


def format_rmspec_summary(spec: RMSpec, *, agent_names=None, source=None, max_core_transition_lines=200) -> str:
    """The readable summary the drivers log after loading a spec (rmgen/summary.py:8-67; office_main.py:525,
    frozen_lake_main.py:170): header, agents, sizes, the non-self-loop transitions in sorted order, the rewarded ones."""
    out = ["Reward Machine summary" + (f" (from {source})" if source else "")]
    names = list(agent_names) if agent_names else []
    if len(names) == 1:
        out.append(f"Agent: {names[0]}")
    elif names:
        out.append("Shared by agents: [" + ", ".join(names) + "]")
    out += [f"name: {spec.name}", f"env_id: {spec.env_id}",
            f"states: {len(spec.states)} (initial: {spec.initial_state}, terminal: {spec.terminal_states})",
            f"event_vocabulary ({len(spec.event_vocabulary)}): " + ", ".join(spec.event_vocabulary),
            f"transitions_total: {len(spec.transitions)}"]
    core = sorted((t for t in spec.transitions if t.from_state != t.to_state), key=lambda t: (t.from_state, t.event, t.to_state))
    out += [f"core_transitions_count: {len(core)}", "core_transitions (excluding self-loops):"]
    shown = core if max_core_transition_lines is None else core[:max_core_transition_lines]
    out += [f"- {t.from_state} --{t.event}--> {t.to_state}" for t in shown]
    if max_core_transition_lines is not None and len(core) > max_core_transition_lines:
        out.append("... truncated")
    rewarded = [t for t in spec.transitions if t.reward > 0]
    out.append(f"transitions_with_reward>0: {len(rewarded)}")
    out += [f"- {t.from_state} --{t.event}--> {t.to_state} (reward={t.reward})" for t in rewarded]
    return "\n".join(out)


def scenario_from_rmspec(path, scenario, **kwargs):
    """Fill `scenario.rm_transitions` / `detector_positions` from a spec file, so `compile_scenario(scenario)` emits the
    device tables directly from (spec file, map name). Note: a Scenario's machine starts in the source of its first
    transition; specs whose initial_state differs are returned as a RewardMachine by load_reward_machine instead."""
    rm, spec = load_reward_machine(path, scenario.env, scenario.map_name, **kwargs)
    first_src = next(iter(rm.transitions))[0]
    if first_src != spec.initial_state:
        raise ValueError("spec.initial_state is not the source of the first transition; pass the RewardMachine to "
                         "compile_scenario(scenario, rm=...) instead")
    scenario.rm_transitions = [(s, ev, t, r) for (s, ev), (t, r) in rm.transitions.items()]
    scenario.detector_positions = sorted(rm.detector_positions())
    return scenario, rm
