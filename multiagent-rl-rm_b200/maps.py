"""Grid-world map data and parsers (host side, table build only).

Mirrors the reference's map utilities:
  * ``parse_map_emoji``      -> /root/reference/multiagent_rlrm/utils/utils.py:169-203
  * ``parse_office_world``   -> utils.py:206-243 (+ ``find_disconnected_pairs`` utils.py:300-364)
  * map layouts              -> environments/frozen_lake/config_frozen_lake.py:10-28,
                                environments/office_world/config_office.py:49-262

Layouts are stored in this repo's own one-character-per-token ASCII form:
  ``.`` floor   ``#`` wall / hole   ``+`` door   ``P`` plant   ``c`` coffee   ``m`` letter   ``A-Z0-9`` goal
Emoji layouts in the reference's format are accepted too (tokens are translated first).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

Pos = Tuple[int, int]

_EMOJI_TO_ASCII = {"🟩": ".", "⛔": "#", "🚪": "+", "🪴": "P", "🥤": "c", "✉️": "m", "✉": "m"}
_ASCII_TO_EMOJI = {".": "🟩", "#": "⛔", "+": "🚪", "P": "🪴", "c": "🥤", "m": "✉️"}

FROZEN_LAKE_MAPS: Dict[str, List[str]] = {
    "map1": [
        "B.........",
        "..........",
        "...##.....",
        "..........",
        "....A.....",
        "..........",
        "..........",
        "#######.##",
        "....C.....",
        "..........",
    ],
}

# name -> (rows, grid_size=(height, width) as the reference writes it, default start position)
OFFICE_WORLD_MAPS: Dict[str, Tuple[List[str], Tuple[int, int], Pos]] = {
    "map0": (
        [
            "c.......Ac",
            "..........",
            "..........",
            "C.........",
            "..........",
            "..........",
            "D........B",
            "..........",
            "..........",
            "O...E....m",
        ],
        (10, 10),
        (0, 0),
    ),
    "map1": (
        [
            "...#...#...#...",
            ".B.+.P.+.P.+.C.",
            "...#c..#...#...",
            "#+###+###+###+#",
            "...#...#...#...",
            ".P.#.O.#.m.#.P.",
            "...#...#...#...",
            "#+###########+#",
            "...#...#..c#...",
            ".A.+.P.+.P.+.D.",
            "E..#...#...#...",
        ],
        (9, 12),
        (2, 7),
    ),
    "map2": (
        [
            "EP.#.P.#.P.#...",
            ".B.+...+...+.D.",
            "...#...#...#...",
            "#+###+###+####+",
            "P..#cP.+...#...",
            ".O.#...#...#.P.",
            "...#c..+...#...",
            "#+#######+####+",
            "...#...#P..#...",
            ".A.#.P.+...#P..",
            "...#.P.#m.P#...",
            "+###+########+#",
            "...+PP.#...#...",
            "...#C..+...+...",
            "P..#...#.P.#.PP",
        ],
        (12, 12),
        (2, 7),
    ),
    "map3": (
        [
            ".PE+..P+PPP+P..+.P.",
            ".AP+...+..P+...#.B.",
            "...+...+...+..P#...",
            "#+#######+####+#++#",
            "...#...+...#...#...",
            ".P.#m..#..c#.PP#.P.",
            "...#PPP+.PP#...#...",
            "#+#######+####+##+#",
            "P..#...#...#.P.#...",
            "...#.P.+.P.+.P.#...",
            "...#.P.#...#...#...",
            "+###+#######++###+#",
            "..P+P.c#...#..P+...",
            ".D.#P..+...+..P#P..",
            "...#...#.P.#...#...",
            "+########+###+###+#",
            "...+...#...#...#...",
            ".O.+.P.+...+...#..C",
            "...#...#...#...+...",
        ],
        (15, 15),
        (2, 7),
    ),
    "map4": (
        [
            ".PE+..P+PPP+P..+.Pc",
            ".A.+...+...+...#B..",
            "...+...+...+..P#P..",
            "#+#######+####+#++#",
            "...#...+...#...#...",
            ".P.#m..#...#.PP#PP.",
            "...#PPP+.PP#...#...",
            "#+#######+####+##+#",
            "...#...#...#.P.#...",
            ".P.#.P.+.P.+.P.#.P.",
            "...#.P.#...#...#...",
            "+###+####+##++###+#",
            "..P+P.c#..P#..P+...",
            ".D.#P..#...+..P#P..",
            "...+...#.P.#...#...",
            "+########+###+###+#",
            "...+...#...#...#...",
            ".O.+.P.+.P.+.P.#.PC",
            "...#...#...#...+...",
        ],
        (15, 15),
        (2, 7),
    ),
}


def _tokens(line: str) -> List[str]:
    """Whitespace-separated tokens of one layout line, translated to the ASCII alphabet."""
    return [_EMOJI_TO_ASCII.get(t, t) for t in line.split()]


def ascii_rows_from_emoji(layout: str) -> List[str]:
    """Translate an emoji layout (whitespace-separated tokens) to ASCII rows; empty lines are dropped."""
    rows = []
    for line in layout.strip().split("\n"):
        toks = _tokens(line)
        if toks:
            rows.append("".join(toks))
    return rows


def emoji_from_ascii_rows(rows: List[str]) -> str:
    """Inverse of :func:`ascii_rows_from_emoji` (used by tests to exercise the emoji front-ends)."""
    return "\n" + "\n".join(" ".join(_ASCII_TO_EMOJI.get(c, c) for c in row) for row in rows) + "\n"


# --------------------------------------------------------------------------------------------------
# FrozenLake
# --------------------------------------------------------------------------------------------------
def parse_frozen_lake_rows(rows: List[str]):
    """(holes, goals, (width, height)) from ASCII rows; (0,0) is top-left, y grows downward."""
    holes: List[Pos] = []
    goals: Dict[str, Pos] = {}
    for y, row in enumerate(rows):
        for x, ch in enumerate(row):
            if ch == "#":
                holes.append((x, y))
            elif ch.isalnum():
                goals[ch] = (x, y)
    return holes, goals, (max(len(r) for r in rows), len(rows))


def parse_map_emoji(map_string: str):
    """Drop-in for utils.parse_map_emoji (utils.py:169-203): characters are cells, blanks are ignored."""
    import textwrap

    rows = []
    for raw in textwrap.dedent(map_string).strip().splitlines():
        rows.append("".join("#" if c == "⛔" else c for c in raw if c != " "))
    # everything that is neither a hole nor alphanumeric is floor; note 'P'/'c'/'m' would be goals here,
    # exactly as in the reference (any letter or digit is a goal symbol).
    holes: List[Pos] = []
    goals: Dict[str, Pos] = {}
    for y, row in enumerate(rows):
        for x, ch in enumerate(row):
            if ch == "#":
                holes.append((x, y))
            elif ch.isdigit() or ch.isalpha():
                goals[ch] = (x, y)
    return holes, goals, (max(len(r) for r in rows), len(rows))


# --------------------------------------------------------------------------------------------------
# OfficeWorld
# --------------------------------------------------------------------------------------------------
def _office_walls(grid: List[List[str]]) -> List[Tuple[Pos, Pos]]:
    """One-way wall pairs between cells separated by a single '#' token (utils.py:300-364).

    Lines made only of '#'/'+' tokens do not count as cell rows/columns; cells are renumbered without them.
    """
    n_rows = len(grid)
    n_cols = len(grid[0]) if n_rows else 0
    solid = ("#", "+")

    def renumber(is_wall_line):
        out, nxt = [], 0
        for flag in is_wall_line:
            if flag:
                out.append(1)  # never used for a real cell; kept identical to the reference's bookkeeping
            else:
                out.append(nxt)
                nxt += 1
        return out

    cy = renumber([all(t in solid for t in grid[y]) for y in range(n_rows)])
    cx = renumber([all(grid[y][x] in solid for y in range(n_rows)) for x in range(n_cols)])

    pairs: List[Tuple[Pos, Pos]] = []
    for y in range(n_rows):
        for x in range(n_cols):
            if grid[y][x] in solid:
                continue
            if x + 2 < n_cols and grid[y][x + 1] == "#" and grid[y][x + 2] != "#":
                pairs.append(((cx[x], cy[y]), (cx[x + 2], cy[y])))
            if y + 2 < n_rows and grid[y + 1][x] == "#" and grid[y + 2][x] != "#":
                # the reference skips one specific cell when a plant sits at raw (4,1) (utils.py:349-351)
                if x == 1 and y == 3 and grid[1][4] == "P":
                    continue
                pairs.append(((cx[x], cy[y]), (cx[x], cy[y + 2])))
    return pairs


def parse_office_rows(rows: List[str]):
    """(coordinates, goals, walls) from ASCII rows, same result layout as utils.parse_office_world."""
    grid = [list(r) for r in rows]
    names = {".": "empty_cell", "P": "plant", "c": "coffee", "m": "letter"}
    coordinates: Dict[str, List[Pos]] = {"plant": [], "coffee": [], "letter": [], "empty_cell": []}
    found: Dict[str, List[Pos]] = {}
    y = 0
    for row in grid:
        cells = [t for t in row if t not in ("#", "+")]
        if not cells:
            continue
        for x, t in enumerate(cells):
            if t in names:
                coordinates[names[t]].append((x, y))
            elif len(t) == 1 and (t.isascii() and (t.isupper() or t.isdigit())):
                found.setdefault(t, []).append((x, y))
        y += 1
    order = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
    goals = {k: found[k][0] for k in order if k in found}
    return coordinates, goals, _office_walls(grid)


def parse_office_world(office_world: str):
    """Drop-in for utils.parse_office_world (utils.py:206-243) on the reference's emoji layout strings."""
    lines = [_tokens(line) for line in office_world.strip().split("\n")]
    # The reference builds the wall grid from every line of the stripped string (utils.py:246-249); the layouts
    # it ships have no blank interior lines, so dropping empties here is equivalent.
    rows = ["".join(t if len(t) == 1 else "?" for t in toks) for toks in lines if toks]
    return parse_office_rows(rows)


@dataclass
class GridSpec:
    """Geometry handed to the table compiler (tables.py)."""

    env: str  # "frozen_lake" | "office_world"
    width: int
    height: int
    hazards: List[Pos] = field(default_factory=list)  # holes / plants
    walls: List[Tuple[Pos, Pos]] = field(default_factory=list)  # directed pairs (office only)
    goals: Dict[str, Pos] = field(default_factory=dict)
    coffee: List[Pos] = field(default_factory=list)
    letters: List[Pos] = field(default_factory=list)
    start: Pos = (0, 0)


def frozen_lake_grid(name: str = "map1") -> GridSpec:
    holes, goals, (w, h) = parse_frozen_lake_rows(FROZEN_LAKE_MAPS[name])
    return GridSpec("frozen_lake", w, h, hazards=holes, goals=goals)


def office_world_grid(name: str = "map1") -> GridSpec:
    rows, (gh, gw), start = OFFICE_WORLD_MAPS[name]
    coords, goals, walls = parse_office_rows(rows)
    # office_main.py:416 makes the pairs bidirectional; grid_size is (height, width) (office_main.py:419-421)
    walls = list(walls) + [(b, a) for (a, b) in walls]
    return GridSpec(
        "office_world",
        gw,
        gh,
        hazards=list(coords["plant"]),
        walls=walls,
        goals=goals,
        coffee=list(coords["coffee"]),
        letters=list(coords["letter"]),
        start=start,
    )
