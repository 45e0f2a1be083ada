"""Built-in OfficeWorld tasks: the reward machines ``office_main --experiment expN`` runs when no --rm-spec is given
(/root/reference/multiagent_rlrm/environments/office_world/config_office.py:292-470, get_experiment_for_map) and the
optimal path lengths its evaluation normalises by (office_main.py:111-135, OPTIMAL).

The tasks are written here as (source, symbol, target, reward) rows over map symbols and expanded against the parsed
map; ``tests/test_host_logic.py`` checks the expansion — transition ORDER included, because the RM's final state and
QRM's counterfactual order depend on it — against the live reference for every map."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from .maps import office_world_grid

# symbol: "A".."O" = goal letters, "letter" = the mail cell, "coffee" = both machines (machine 0 first),
# "coffee0" / "coffee1" = one machine
_TASKS: Dict[str, Tuple[str, List[Tuple[int, str, int, float]]]] = {
    "exp1": ("Coffee to Office", [(0, "coffee", 1, 0), (1, "O", 2, 1)]),
    "exp2": ("E-mail to Office", [(0, "letter", 1, 0), (1, "O", 2, 1)]),
    "exp3": ("Coffee + Email to Office", [(0, "letter", 1, 0), (0, "coffee", 2, 0), (2, "letter", 3, 0), (1, "coffee", 3, 0), (3, "O", 4, 1)]),
    "exp4": ("A-B-C-D", [(0, "A", 1, 0), (1, "B", 2, 0), (2, "C", 3, 0), (3, "D", 4, 1)]),
    "exp5": ("A-B-C-D, then Coffee + Email to Office",
             [(0, "A", 1, 0), (1, "B", 2, 0), (2, "C", 3, 0), (3, "D", 4, 0), (4, "coffee", 5, 0), (4, "letter", 6, 0), (6, "coffee", 7, 0),
              (5, "letter", 7, 0), (7, "O", 8, 1)]),
    "exp6": ("A-B-C-D-E then Coffee + Email to Office",
             [(0, "A", 1, 0), (1, "B", 2, 0), (2, "C", 3, 0), (3, "D", 4, 0), (4, "E", 5, 0), (5, "coffee", 6, 0), (5, "letter", 7, 0),
              (7, "coffee", 8, 0), (6, "letter", 8, 0), (8, "O", 9, 1)]),
    "exp7": ("Coffee to Office", [(0, "coffee0", 1, 0), (0, "coffee1", 2, 0), (1, "O", 3, 1000), (2, "O", 3, 1)]),
    "exp0": ("Letter + Coffe to Office", [(0, "letter", 1, 0), (1, "coffee0", 2, 0), (2, "O", 3, 1)]),
    "exp0_simply": ("Letter", [(0, "letter", 1, 1)]),
}
_POSITION_ORDER = {  # the reference's `positions` sets, as symbol lists (a set: order is immaterial)
    "exp1": ["coffee", "O"], "exp2": ["letter", "O"], "exp3": ["letter", "coffee", "O"], "exp4": ["A", "B", "C", "D"],
    "exp5": ["A", "B", "C", "D", "coffee", "letter", "O"], "exp6": ["A", "B", "C", "D", "E", "coffee", "letter", "O"],
    "exp7": ["coffee", "O"], "exp0": ["letter", "coffee0", "O"], "exp0_simply": ["letter"],
}

OPTIMAL = {"map0;exp0_simply": 18, "map0;exp0": 28,
           "map1;exp1": 15, "map1;exp2": 29, "map1;exp3": 29, "map1;exp4": 30, "map1;exp5": 55, "map1;exp6": 75,
           "map2;exp1": 48, "map2;exp2": 98, "map2;exp3": 106, "map2;exp4": 81, "map2;exp5": 152, "map2;exp6": 212,
           "map3;exp1": 50, "map3;exp2": 54, "map3;exp3": 98, "map3;exp4": 108, "map3;exp5": 206, "map3;exp6": 272,
           "map4;exp7": 1000}


def _cells(grid, symbol, strict=True):
    if symbol == "coffee":
        return [grid.coffee[0], grid.coffee[1]]
    if symbol in ("coffee0", "coffee1"):
        return [grid.coffee[int(symbol[-1])]]
    if symbol == "letter":
        return [grid.letters[0]]
    if strict:
        return [grid.goals[symbol]]
    return [grid.goals.get(symbol)]


def get_experiment_for_map(map_name: str, selected_experiment: str) -> Optional[dict]:
    """{"description", "transitions": {(state, position): (next_state, reward)}, "positions": set} or None for an unknown
    experiment key. ("map4", "exp6") is the paper's good/regular-coffee task, i.e. exp7's machine (config_office.py:463-468).
    Like the reference, building the table needs the symbols of EVERY task on the map: a map without goal "A" raises
    KeyError whatever experiment is asked for, except that "A" / "B" / "E" of exp6 may be absent (then the event is None)."""
    grid = office_world_grid(map_name)
    built = {}
    for key, (description, rows) in _TASKS.items():
        lenient = {"A", "B", "E"} if key == "exp6" else set()
        transitions = {}
        for (src, symbol, dst, reward) in rows:
            for cell in _cells(grid, symbol, strict=symbol not in lenient):
                transitions[(f"state{src}", cell)] = (f"state{dst}", reward)
        positions = set()
        for symbol in _POSITION_ORDER[key]:
            positions.update(_cells(grid, symbol, strict=not (key == "exp6" and symbol == "E")))
        built[key] = {"description": description, "transitions": transitions, "positions": positions}
    if map_name == "map4" and selected_experiment == "exp6":
        paper = dict(built["exp7"])
        paper["description"] = "Coffee to Office with good/regular coffee (paper map4 exp6)"
        return paper
    return built.get(selected_experiment)


def scenario_for_experiment(map_name: str, experiment: str, **scenario_kwargs):
    """A tables.Scenario for office_main's built-in task: transitions in the reference's insertion order, detector positions
    = the experiment's position set."""
    from .tables import Scenario

    exp = get_experiment_for_map(map_name, experiment)
    if exp is None:
        raise KeyError(f"unknown experiment {experiment!r}")
    rows = [(s, ev, t, r) for (s, ev), (t, r) in exp["transitions"].items()]
    kw = dict(env="office_world", map_name=map_name, rm_transitions=rows, detector_positions=sorted(exp["positions"]),
              driver="office_main")
    kw.update(scenario_kwargs)
    return Scenario(**kw)
