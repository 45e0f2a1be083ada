"""Host-side table compiler: (grid, dynamics flags, reward machine, learner settings) -> flat device tables.

Everything the kernels index at run time is produced here, once, on the host:

  next_cell[W*H][4]   move table   ma_frozen_lake.py:224-242 ; ma_office.py:269-289 + config_office.py:12-39
  cell_flags[W*H]     hole / plant ma_frozen_lake.py:174-187 ; ma_office.py:204-220
  slip thresholds     numpy Generator.choice(p=...) == one uniform + right-searchsorted on the normalised cdf
                      (ma_frozen_lake.py:257-262, 283-296 ; ma_office.py:327-379)
  label/delta/rq/rcf  RewardMachine.compile_tables (reward_machine.py in this package)

A :class:`Scenario` is plain data (JSON-able through ``to_dict``) so the same description drives the reference
harness (oracle/ref_harness.py), the C oracle and the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _abi as abi
from .maps import GridSpec, frozen_lake_grid, office_world_grid
from .reward_machine import PositionEventDetector, RewardMachine

UP, DOWN, LEFT, RIGHT, WAIT = 0, 1, 2, 3, 4


@dataclass
class Scenario:
    env: str = "frozen_lake"  # "frozen_lake" | "office_world"
    map_name: str = "map1"
    starts: List[Tuple[int, int]] = field(default_factory=lambda: [(5, 0), (0, 0)])
    # reward machine as an ordered transition list [(src, event_pos | None, dst, reward)] (dict order matters)
    rm_transitions: List[Tuple[str, Optional[Tuple[int, int]], str, float]] = field(default_factory=list)
    detector_positions: Optional[List[Tuple[int, int]]] = None  # default: the positions used by transitions
    # agents with different reward machines (frozen_lake_main.py --rm-spec-a1/--rm-spec-a2): one transition list per agent
    rm_transitions_per_agent: Optional[List[List[Tuple[str, Optional[Tuple[int, int]], str, float]]]] = None
    reward_modifier: float = 1
    # dynamics
    stochastic: bool = False
    delay_action: bool = False
    all_slip: bool = False
    high_prob: float = 0.8
    penalty_amount: float = 0  # FrozenLake hole penalty
    plants_penalty: float = -100
    wall_penalty: float = 0
    terminate_on_plants: bool = False
    terminate_hit_walls: bool = False
    max_steps: int = 1000
    random_start_positions: bool = False  # FrozenLake only (ma_frozen_lake.py:63-64)
    # learner
    algo: str = "qrm"  # "ql" | "qrm" | "qlambda"
    learning_rate: Optional[float] = 1.0
    gamma: float = 0.99
    lambd: float = 0.0
    epsilon_start: float = 0.01
    epsilon_end: float = 0.01
    epsilon_decay: float = 0.9995
    q_init: float = 2.0
    driver: str = "frozen_lake_main"  # "frozen_lake_main" | "office_main"
    shared_q: bool = False
    # element type of the learner tables on the device: "f32" (NumPy-on-float32 arithmetic, half the bytes, all specialised
    # kernels) or "f64" (the reference's native float64 tables, qlearning.py:26-29 — bit-identical to the unmodified reference)
    table_dtype: str = "f32"
    seed: int = 1234
    # reward shaping on the RM (QL_RS / QRM_RS, office_main.py:543-573): "vi" = add_reward_shaping, "distance" = add_distance_reward_shaping
    use_rsh: bool = False
    rs_kind: str = "vi"
    rs_gamma: float = 0.9
    rs_alpha: float = 100

    def to_dict(self):
        import dataclasses

        d = dataclasses.asdict(self)
        d["starts"] = [list(p) for p in self.starts]
        d["rm_transitions"] = [[s, None if e is None else list(e), t, r] for (s, e, t, r) in self.rm_transitions]
        if self.detector_positions is not None:
            d["detector_positions"] = [list(p) for p in self.detector_positions]
        if self.rm_transitions_per_agent is not None:
            d["rm_transitions_per_agent"] = [[[s, None if e is None else list(e), t, r] for (s, e, t, r) in tr]
                                             for tr in self.rm_transitions_per_agent]
        return d

    @staticmethod
    def from_dict(d):
        d = dict(d)
        d["starts"] = [tuple(p) for p in d["starts"]]
        d["rm_transitions"] = [(s, None if e is None else tuple(e), t, r) for (s, e, t, r) in d["rm_transitions"]]
        if d.get("detector_positions") is not None:
            d["detector_positions"] = [tuple(p) for p in d["detector_positions"]]
        if d.get("rm_transitions_per_agent") is not None:
            d["rm_transitions_per_agent"] = [[(s, None if e is None else tuple(e), t, r) for (s, e, t, r) in tr]
                                             for tr in d["rm_transitions_per_agent"]]
        return Scenario(**d)

    # -- helpers ---------------------------------------------------------------------------------
    def grid(self) -> GridSpec:
        return frozen_lake_grid(self.map_name) if self.env == "frozen_lake" else office_world_grid(self.map_name)

    def reward_machine(self, agent: Optional[int] = None) -> RewardMachine:
        trs = self.rm_transitions if (agent is None or self.rm_transitions_per_agent is None) else self.rm_transitions_per_agent[agent]
        transitions = {(s, e): (t, r) for (s, e, t, r) in trs}
        if self.detector_positions is not None:
            positions = set(self.detector_positions)
        else:
            positions = {e for (_s, e, _t, _r) in trs if e is not None}
        return RewardMachine(transitions, PositionEventDetector(positions))


# --------------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------------
def build_next_cell(grid: GridSpec) -> np.ndarray:
    W, H = grid.width, grid.height
    nxt = np.zeros((W * H, 4), dtype=np.uint16)
    walls = set(grid.walls)
    for y in range(H):
        for x in range(W):
            here = y * W + x
            if grid.env == "frozen_lake":  # (0,0) top-left, "up" is y-1
                cand = {UP: (x, y - 1), DOWN: (x, y + 1), LEFT: (x - 1, y), RIGHT: (x + 1, y)}
            else:  # OfficeWorld: "up" is y+1 (ma_office.py:280-287)
                cand = {UP: (x, y + 1), DOWN: (x, y - 1), LEFT: (x - 1, y), RIGHT: (x + 1, y)}
            for a, (nx, ny) in cand.items():
                ok = 0 <= nx < W and 0 <= ny < H and ((x, y), (nx, ny)) not in walls
                nxt[here, a] = ny * W + nx if ok else here
    return nxt


def build_cell_flags(grid: GridSpec) -> np.ndarray:
    flags = np.zeros(grid.width * grid.height, dtype=np.uint8)
    for (x, y) in grid.hazards:
        if 0 <= x < grid.width and 0 <= y < grid.height:
            flags[y * grid.width + x] |= 1
    return flags


# --------------------------------------------------------------------------------------------------
# slip model
# --------------------------------------------------------------------------------------------------
def slip_table(env: str, stochastic: bool, delay_action: bool, all_slip: bool, high_prob: float):
    """(outcomes[4][n], probabilities[n]) with action indices; n == 1 when deterministic."""
    if not stochastic:
        return [[a] for a in range(4)], [1.0]
    perp = {UP: (LEFT, RIGHT), DOWN: (LEFT, RIGHT), LEFT: (UP, DOWN), RIGHT: (UP, DOWN)}
    opposite = {UP: DOWN, DOWN: UP, LEFT: RIGHT, RIGHT: LEFT}
    if delay_action:
        return [[WAIT, a, perp[a][0], perp[a][1]] for a in range(4)], [0.6, 0.36, 0.02, 0.02]
    if env == "frozen_lake":
        return [[a, perp[a][0], perp[a][1]] for a in range(4)], [0.8, 0.1, 0.1]
    hp = high_prob
    if all_slip:
        lp = (1 - hp) / 3
        outs = [[a, opposite[a]] + [b for b in (UP, DOWN, LEFT, RIGHT) if b not in (a, opposite[a])] for a in range(4)]
        # reference order: left->[left,right,up,down], right->[right,left,up,down], up->[up,down,left,right], down->[down,up,left,right]
        return outs, [hp, lp, lp, lp]
    lp = (1 - hp) / 2
    return [[a, perp[a][0], perp[a][1]] for a in range(4)], [hp, lp, lp]


def mdp_action_distribution(sc: "Scenario"):
    """(sub-actions[4][n], probabilities[n]) RMEnvironmentWrapper.get_mdp works with (rm_environment_wrapper.py:196-197, 244):
    OfficeWorld's ``stochastic`` flag is forced off first, so only ``delay_action`` keeps a real distribution there
    (ma_office.py:333-347); FrozenLake's map follows ``frozen_lake_stochastic`` (ma_frozen_lake.py:264-296)."""
    if sc.env == "frozen_lake":
        return slip_table(sc.env, bool(sc.stochastic), bool(sc.delay_action), False, 0.8)
    return slip_table(sc.env, bool(sc.delay_action), bool(sc.delay_action), False, sc.high_prob)


def slip_thresholds(probabilities: Sequence[float]) -> List[int]:
    """Integer thresholds T_j = ceil(cdf_j * 2^32): for a 32-bit draw k, u = k / 2^32,
    numpy's ``cdf.searchsorted(u, side='right')`` equals ``#{j : k >= T_j}`` exactly."""
    cdf = np.cumsum(np.array(probabilities, dtype=np.float64))
    cdf /= cdf[-1]
    out = []
    for c in cdf[:-1]:
        fr = Fraction(float(c)) * (1 << 32)
        out.append(int(-((-fr.numerator) // fr.denominator)))  # ceil
    return out


# --------------------------------------------------------------------------------------------------
# compiled scenario
# --------------------------------------------------------------------------------------------------
@dataclass
class Compiled:
    scenario: Scenario
    grid: GridSpec
    rm: RewardMachine
    next_cell: np.ndarray
    cell_flags: np.ndarray
    label: np.ndarray
    delta: np.ndarray
    rq: np.ndarray
    rcf: np.ndarray
    qrm_states: np.ndarray
    start_cell: np.ndarray
    events: List[Tuple[int, int]]
    config: abi.Config
    phi: Optional[np.ndarray] = None
    free_cells: Optional[np.ndarray] = None

    @property
    def n_agents(self):
        return self.config.n_agents

    @property
    def n_rm_states(self):
        return self.config.n_rm_states

    @property
    def state_space(self):
        """S = W*H*nQ of the (uniform) per-agent table; with per-agent reward machines use `agent_rows`."""
        return self.config.width * self.config.height * self.config.n_rm_states

    @property
    def agent_rows(self):
        """[S_a] rows of each agent's table (S_a = W*H*nQ_a)."""
        c = self.config
        cells = c.width * c.height
        if c.per_agent_rm:
            return [cells * c.agent_n_rm_states[a] for a in range(c.n_agents)]
        return [cells * c.n_rm_states] * c.n_agents

    def tables_struct(self) -> abi.Tables:
        t = abi.Tables()
        for name in ("next_cell", "cell_flags", "label", "delta", "rq", "rcf", "qrm_states", "start_cell"):
            arr = np.ascontiguousarray(getattr(self, name))
            setattr(self, name, arr)  # keep alive
            setattr(t, name, arr.ctypes.data if arr.size else None)
        if self.phi is not None:
            self.phi = np.ascontiguousarray(self.phi, dtype=np.float64)
            t.phi = self.phi.ctypes.data
        if self.free_cells is not None:
            self.free_cells = np.ascontiguousarray(self.free_cells, dtype=np.uint16)
            t.free_cells = self.free_cells.ctypes.data
        return t


def dump_blob(compiled: "Compiled", path) -> None:
    """Write the compiled scenario for a non-Python host of the C ABI (examples/c_abi_demo.c): the rlrm_config_t bytes, the
    initial Q value (f64), then every table of rlrm_tables_t in field order as <u64 byte count><bytes> (count 0 = NULL)."""
    import ctypes as C
    import struct

    t = compiled.tables_struct()  # makes every array contiguous and keeps it alive on `compiled`
    with open(path, "wb") as f:
        f.write(bytes(memoryview((C.c_char * C.sizeof(compiled.config)).from_address(C.addressof(compiled.config)))))
        f.write(struct.pack("<d", float(compiled.scenario.q_init)))
        for name in ("next_cell", "cell_flags", "label", "delta", "rq", "rcf", "qrm_states", "start_cell", "free_cells", "phi"):
            arr = getattr(compiled, name, None)
            data = b"" if arr is None or getattr(t, name) is None else np.ascontiguousarray(arr).tobytes()
            f.write(struct.pack("<Q", len(data)))
            f.write(data)


_ALGO = {"ql": abi.ALGO_QL, "qrm": abi.ALGO_QRM, "qlambda": abi.ALGO_QLAMBDA}
_DRIVER = {"frozen_lake_main": abi.DRIVER_FROZEN_LAKE_MAIN, "office_main": abi.DRIVER_OFFICE_MAIN}


def compile_scenario(sc: Scenario, grid: Optional[GridSpec] = None, rm: Optional[RewardMachine] = None,
                     instance_offset: int = 0) -> Compiled:
    grid = grid or sc.grid()
    W, H = grid.width, grid.height
    if W * H > abi.MAX_CELLS:
        raise ValueError(f"grid {W}x{H} exceeds {abi.MAX_CELLS} cells")
    if not (1 <= len(sc.starts) <= abi.MAX_AGENTS):
        raise ValueError("1..8 agents per instance")
    per_agent = rm is None and sc.rm_transitions_per_agent is not None
    rms = None
    if isinstance(rm, (list, tuple)):  # explicit list of per-agent RewardMachine objects
        rms, rm, per_agent = list(rm), rm[0], True
    elif per_agent:
        rms = [sc.reward_machine(a) for a in range(len(sc.starts))]
        rm = rms[0]
    if per_agent and len(rms) != len(sc.starts):
        raise ValueError("one reward machine per agent is required")
    rm = rm or sc.reward_machine()
    t = rm.compile_tables(W, H, sc.reward_modifier)

    cfg = abi.Config()
    cfg.abi_version = abi.ABI_VERSION
    cfg.env_kind = abi.ENV_FROZEN_LAKE if grid.env == "frozen_lake" else abi.ENV_OFFICE_WORLD
    cfg.driver = _DRIVER[sc.driver]
    cfg.algo = _ALGO[sc.algo]
    cfg.width, cfg.height = W, H
    cfg.n_agents = len(sc.starts)
    cfg.n_rm_states = t["n_states"]
    cfg.n_events = t["n_events"]
    cfg.rm_final = t["final"]
    cfg.n_qrm_states = len(t["qrm_states"])
    cfg.max_steps = sc.max_steps
    cfg.stochastic = int(bool(sc.stochastic))
    outs, probs = slip_table(grid.env, sc.stochastic, sc.delay_action, sc.all_slip, sc.high_prob)
    cfg.slip_n = len(probs)
    thr = slip_thresholds(probs)
    for j in range(3):
        cfg.slip_thr[j] = thr[j] if j < len(thr) else (1 << 32)
    for a in range(4):
        for j in range(4):
            cfg.slip_outcome[a][j] = outs[a][j] if j < len(outs[a]) else outs[a][-1]
    cfg.terminate_on_plants = int(bool(sc.terminate_on_plants))
    cfg.terminate_hit_walls = int(bool(sc.terminate_hit_walls))
    cfg.hole_penalty = float(sc.penalty_amount if grid.env == "frozen_lake" else sc.plants_penalty)
    cfg.wall_penalty = float(0 if grid.env == "frozen_lake" else sc.wall_penalty)
    cfg.learning_rate = -1.0 if sc.learning_rate is None else float(sc.learning_rate)
    cfg.gamma = float(sc.gamma)
    cfg.lambd = float(sc.lambd)
    cfg.epsilon_start = float(sc.epsilon_start)
    cfg.epsilon_end = float(sc.epsilon_end)
    cfg.epsilon_decay = float(sc.epsilon_decay)
    # env.reset() decays epsilon only for `isinstance(l_algo, QLearning)` (ma_frozen_lake.py:87-88, ma_office.py:107-108)
    cfg.decay_on_reset = int(sc.algo in ("ql", "qrm"))
    cfg.shared_q = int(bool(sc.shared_q))
    cfg.seed_lo = sc.seed & 0xFFFFFFFF
    cfg.seed_hi = (sc.seed >> 32) & 0xFFFFFFFF
    cfg.instance_offset = instance_offset
    cfg.n_actions = 4
    if sc.table_dtype not in ("f32", "f64"):
        raise ValueError("table_dtype must be 'f32' or 'f64'")
    cfg.table_dtype = abi.TABLE_F64 if sc.table_dtype == "f64" else abi.TABLE_F32
    phi = None
    if sc.use_rsh or getattr(rm, "potentials", None) is not None:
        if getattr(rm, "potentials", None) is None:
            if sc.rs_kind == "distance":
                rm.add_distance_reward_shaping(sc.gamma, sc.rs_gamma, sc.rs_alpha)
            else:
                rm.add_reward_shaping(sc.gamma, sc.rs_gamma)
        # row 0: by label (QRM counterfactuals); row 1: potentials.get(<integer index>, 0), the lookup the plain-QL
        # branch of the reference performs (see include/rlrm_b200.h, rlrm_tables_t.phi)
        phi = np.zeros((2, t["n_states"]), dtype=np.float64)
        for state, idx in rm.state_indices.items():
            phi[0, idx] = rm.potentials.get(state, 0)
            phi[1, idx] = rm.potentials.get(idx, 0)
    cfg.use_rsh = int(bool(sc.use_rsh) and sc.algo in ("ql", "qrm"))
    if per_agent:
        # one table section per agent, rows padded to the largest machine; every agent keeps its own event ids
        parts = [m.compile_tables(W, H, sc.reward_modifier) for m in rms]
        if len({tuple(p["events"]) for p in parts}) != 1:
            # different detectors: give every agent the union as event vocabulary (an event outside its own detector set is
            # simply labelled EVENT_NONE in that agent's label section)
            union = sorted({e for p in parts for e in p["events"]})
            for m in rms:
                m._event_vocabulary = union
            parts = [_compile_with_vocabulary(m, W, H, sc.reward_modifier, union) for m in rms]
        n_ev, nq_max = parts[0]["n_events"], max(p["n_states"] for p in parts)
        A = len(rms)
        label = np.stack([p["label"] for p in parts])
        delta = np.full((A, nq_max, n_ev + 1), abi.NO_TRANSITION, dtype=np.uint8)
        rqa, rcfa = np.zeros((A, nq_max, n_ev + 1)), np.zeros((A, nq_max, n_ev + 1))
        qrm = np.zeros((A, nq_max), dtype=np.uint8)
        for a, p in enumerate(parts):
            n = p["n_states"]
            delta[a, :n], rqa[a, :n], rcfa[a, :n] = p["delta"], p["rq"], p["rcf"]
            qrm[a, : len(p["qrm_states"])] = p["qrm_states"]
            cfg.agent_n_rm_states[a], cfg.agent_rm_final[a], cfg.agent_n_qrm[a] = n, p["final"], len(p["qrm_states"])
        cfg.per_agent_rm, cfg.n_rm_states, cfg.n_events = 1, nq_max, n_ev
        cfg.n_qrm_states, cfg.rm_final = max(len(p["qrm_states"]) for p in parts), parts[0]["final"]
        t = dict(t, label=label, delta=delta, rq=rqa, rcf=rcfa, qrm_states=qrm, events=parts[0]["events"])
        if phi is not None:  # every agent's machine carries its own potentials: one [2][nQmax] section per agent
            phi = np.zeros((A, 2, nq_max), dtype=np.float64)
            for a, m in enumerate(rms):
                if getattr(m, "potentials", None) is None:
                    if sc.rs_kind == "distance":
                        m.add_distance_reward_shaping(sc.gamma, sc.rs_gamma, sc.rs_alpha)
                    else:
                        m.add_reward_shaping(sc.gamma, sc.rs_gamma)
                for state, idx in m.state_indices.items():
                    phi[a, 0, idx] = m.potentials.get(state, 0)
                    phi[a, 1, idx] = m.potentials.get(idx, 0)
    start_cell = np.array([y * W + x for (x, y) in sc.starts], dtype=np.uint16)
    free_cells = None
    if sc.random_start_positions:
        if grid.env != "frozen_lake":
            raise ValueError("random_start_positions exists only in MultiAgentFrozenLake (ma_frozen_lake.py:63)")
        holes = set(grid.hazards)
        free = [(x, y) for x in range(W) for y in range(H) if (x, y) not in holes]  # x-major, as ma_frozen_lake.py:163-168
        if len(free) < len(sc.starts):
            raise ValueError("Not enough free cells to place all agents.")
        free_cells = np.array([y * W + x for (x, y) in free], dtype=np.uint16)
        cfg.random_starts, cfg.n_free_cells = 1, len(free)
    return Compiled(
        scenario=sc, grid=grid, rm=rm,
        next_cell=build_next_cell(grid), cell_flags=build_cell_flags(grid),
        label=t["label"], delta=t["delta"], rq=t["rq"], rcf=t["rcf"], qrm_states=t["qrm_states"],
        start_cell=start_cell, events=t["events"], config=cfg, phi=phi, free_cells=free_cells,
    )


def _compile_with_vocabulary(rm, width, height, reward_modifier, vocabulary):
    """compile_tables with a fixed event vocabulary (ids = positions in `vocabulary`); positions outside the machine's own
    detector set stay EVENT_NONE in its label."""
    own = set(rm.detector_positions())
    t = rm.compile_tables(width, height, reward_modifier)
    remap = {k: vocabulary.index(p) for k, p in enumerate(t["events"])}
    n_ev = len(vocabulary)
    label = np.full(width * height, abi.EVENT_NONE, dtype=np.uint8)
    for k, (x, y) in enumerate(vocabulary):
        if (x, y) in own:
            label[y * width + x] = k
    n = t["n_states"]
    delta = np.full((n, n_ev + 1), abi.NO_TRANSITION, dtype=np.uint8)
    rq, rcf = np.zeros((n, n_ev + 1)), np.zeros((n, n_ev + 1))
    for old, new in list(remap.items()) + [(t["n_events"], n_ev)]:
        delta[:, new], rq[:, new], rcf[:, new] = t["delta"][:, old], t["rq"][:, old], t["rcf"][:, old]
    return dict(t, label=label, delta=delta, rq=rq, rcf=rcf, n_events=n_ev, events=list(vocabulary))


# --------------------------------------------------------------------------------------------------
# the BASELINE.json configurations as scenarios
# --------------------------------------------------------------------------------------------------
def frozen_lake_abc_transitions(map_name="map1"):
    g = frozen_lake_grid(map_name).goals
    return [("state0", g["A"], "state1", 10), ("state1", g["B"], "state2", 15), ("state2", g["C"], "state3", 20)]


def scenario_config1() -> Scenario:
    """frozen_lake_main --map map1: 2 agents, deterministic, built-in A->B->C, QLearning(use_qrm=True)."""
    return Scenario(env="frozen_lake", starts=[(5, 0), (0, 0)], rm_transitions=frozen_lake_abc_transitions(),
                    stochastic=False, algo="qrm", learning_rate=1.0, gamma=0.99, epsilon_start=0.01,
                    epsilon_end=0.01, epsilon_decay=0.9995, q_init=2.0, driver="frozen_lake_main", seed=111)


def scenario_config3(use_qrm=True) -> Scenario:
    """65,536 batched FrozenLake map1 instances x 2 agents, slippery, per-instance Q-tables."""
    sc = scenario_config1()
    sc.stochastic = True
    sc.seed = 1234
    if not use_qrm:
        sc.algo, sc.learning_rate = "ql", 0.1
    return sc


def office_acbd_transitions(complete=False):
    """'A -> C -> B -> D, reward on D' on OfficeWorld map1 (authored fixture)."""
    g = office_world_grid("map1").goals
    chain = [("q0", g["A"], "q1", 0.0), ("q1", g["C"], "q2", 0.0), ("q2", g["B"], "q3", 0.0), ("q3", g["D"], "q4", 1.0)]
    if not complete:
        return chain
    events = [g["A"], g["C"], g["B"], g["D"]]
    have = {(s, e) for (s, e, _t, _r) in chain}
    out = list(chain)
    for s in ("q0", "q1", "q2", "q3", "q4"):
        for e in events:
            if (s, e) not in have:
                out.append((s, e, s, 0.0))
    return out


def scenario_config2(stochastic=False) -> Scenario:
    return Scenario(env="office_world", starts=[(2, 7)], rm_transitions=office_acbd_transitions(),
                    detector_positions=sorted(set(office_world_grid("map1").goals.values())
                                              | set(office_world_grid("map1").coffee)
                                              | set(office_world_grid("map1").letters)),
                    stochastic=stochastic, high_prob=0.8, algo="ql", learning_rate=0.1 if stochastic else 1.0,
                    gamma=0.9, epsilon_start=0.1, epsilon_end=0.1, epsilon_decay=1.0, q_init=2.0,
                    driver="office_main", seed=100)


def office_chain12_transitions():
    """Synthetic 12-state completed chain A,B,C,D,E,coffee,letter,O,A,C,B: 108 transitions."""
    gr = office_world_grid("map1")
    g = gr.goals
    vocab = [("A", [g["A"]]), ("B", [g["B"]]), ("C", [g["C"]]), ("D", [g["D"]]), ("E", [g["E"]]),
             ("coffee", list(gr.coffee)), ("letter", list(gr.letters)), ("O", [g["O"]])]
    byname = dict(vocab)
    chain_events = ["A", "B", "C", "D", "E", "coffee", "letter", "O", "A", "C", "B"]
    out, have = [], set()
    for i, ev in enumerate(chain_events):
        r = 1.0 if i == len(chain_events) - 1 else 0.0
        for pos in byname[ev]:
            out.append((f"q{i}", pos, f"q{i + 1}", r))
            have.add((f"q{i}", pos))
    for i in range(12):
        for _name, poss in vocab:
            for pos in poss:
                if (f"q{i}", pos) not in have:
                    out.append((f"q{i}", pos, f"q{i}", 0.0))
                    have.add((f"q{i}", pos))
    return out


def scenario_config4() -> Scenario:
    return Scenario(env="office_world", starts=[(2, 7), (0, 0), (6, 3), (11, 8)],
                    rm_transitions=office_chain12_transitions(), stochastic=True, high_prob=0.8,
                    algo="qlambda", learning_rate=0.1, gamma=0.9, lambd=0.9, epsilon_start=0.1, epsilon_end=0.1,
                    epsilon_decay=1.0, q_init=0.0, driver="office_main", seed=1234)


def scenario_config5(shared=True) -> Scenario:
    sc = scenario_config3(use_qrm=True)
    sc.starts = [(5, 0), (0, 0), (9, 0), (7, 3)]
    sc.shared_q = shared
    return sc
