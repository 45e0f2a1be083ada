"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/rlrm_oracle.c). Not part of the product path.

``Oracle(compiled, n_instances, real="f32")`` owns numpy state with exactly the layout the CUDA library uses
(include/rlrm_b200.h), so tests compare raw buffers. ``real="f64"`` is the reference's native float64 arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import multiagent_rlrm_b200  # noqa: E402  (struct layouts only)
from multiagent_rlrm_b200 import _abi as abi  # noqa: E402

STATS_DTYPE = np.dtype([("active_steps", "<u8"), ("episodes", "<u4"), ("successes", "<u4"), ("return_sum", "<f8"),
                        ("last_return", "<f4"), ("last_length", "<u4")])
assert STATS_DTYPE.itemsize == 32
EVAL_DTYPE = np.dtype([("cum_gamma", "<f8"), ("disc_return", "<f8"), ("return_sum", "<f8"), ("return_sqsum", "<f8"),
                       ("arps_sum", "<f8"), ("len_sum", "<u8"), ("len_sqsum", "<u8"), ("episodes", "<u4"), ("successes", "<u4"),
                       ("in_success", "<u4"), ("reserved", "<u4")])
assert EVAL_DTYPE.itemsize == 72

_libs = {}


def build(force=False):
    """Compile both oracle flavours with gcc (oracle/Makefile)."""
    targets = [os.path.join(_HERE, f"liboracle_{k}.so") for k in ("f32", "f64")]
    src = os.path.join(_HERE, "rlrm_oracle.c")
    hdr = os.path.join(_ROOT, "include", "rlrm_b200.h")
    stale = force or any((not os.path.exists(t)) or os.path.getmtime(t) < max(os.path.getmtime(src), os.path.getmtime(hdr))
                         for t in targets)
    if stale:
        import fcntl

        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:  # several ranks / pytest workers may get here together
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return targets


def lib(real="f32"):
    if real not in _libs:
        path = os.path.join(_HERE, f"liboracle_{real}.so")
        build()  # no-op unless missing or older than rlrm_oracle.c / the header
        L = C.CDLL(path)
        assert L.oracle_real_size() == (4 if real == "f32" else 8)
        _libs[real] = L
    return _libs[real]


def _p(arr):
    return None if arr is None else arr.ctypes.data


class Oracle:
    def __init__(self, compiled, n_instances, real=None, track_visits=False):
        self.c = compiled
        self.cfg = compiled.config
        if real is None:  # follow the scenario's table type unless the caller asks for a flavour explicitly
            real = "f64" if getattr(compiled.config, "table_dtype", 0) == abi.TABLE_F64 else "f32"
        self.tables = compiled.tables_struct()
        self.N, self.A = int(n_instances), compiled.n_agents
        self.S = compiled.state_space
        self.real = real
        self.L = lib(real)
        dt = np.float32 if real == "f32" else np.float64
        n_tab = self.A if self.cfg.shared_q else self.N * self.A
        tab_shape = (n_tab, self.S, 4)
        if self.cfg.per_agent_rm:  # tables of one instance concatenated in agent order, S_a rows each
            tab_shape = (1 if self.cfg.shared_q else self.N, sum(compiled.agent_rows), 4)
        self.slot = np.zeros(self.N * self.A, dtype=np.uint64)
        self.epsilon = np.full(self.N * self.A, self.cfg.epsilon_start, dtype=np.float64)
        self.q = np.full(tab_shape, compiled.scenario.q_init, dtype=dt)
        self.e = np.zeros(tab_shape, dtype=dt) if self.cfg.algo == abi.ALGO_QLAMBDA else None
        need_visits = track_visits or self.cfg.learning_rate < 0
        self.visits = np.zeros(tab_shape, dtype=np.uint32) if need_visits else None
        self.ep_return = np.zeros(self.N * self.A, dtype=np.float64)
        self.stats = np.zeros(self.N * self.A, dtype=STATS_DTYPE)
        shared = bool(self.cfg.shared_q)
        self.acc_sum = np.zeros(tab_shape, dtype=np.int64) if shared else None
        self.acc_cnt = np.zeros(tab_shape, dtype=np.int32) if shared else None
        self.acc_last = np.zeros(tab_shape, dtype=np.float32) if shared else None
        self.state = abi.State(self.N, _p(self.slot), _p(self.epsilon), _p(self.q), _p(self.e), _p(self.visits),
                               _p(self.ep_return), _p(self.stats), _p(self.acc_sum), _p(self.acc_cnt), _p(self.acc_last))

    # -- C calls -------------------------------------------------------------------------------
    def reset(self, mask=None, t=0):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.oracle_reset_at(C.byref(self.cfg), C.byref(self.tables), C.byref(self.state), C.c_void_p(_p(m)), C.c_uint64(t))

    def select_action(self, t=0, draws=None, best=False):
        out = np.zeros(self.N * self.A, dtype=np.uint8)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.uint32)
        self.L.oracle_select_action(C.byref(self.cfg), C.byref(self.tables), C.byref(self.state), C.c_void_p(_p(d)),
                                    C.c_uint64(t), C.c_int(int(best)), C.c_void_p(_p(out)))
        return out.reshape(self.N, self.A)

    def step(self, actions, t=0, draws=None, with_rm=True, counterfactuals=False):
        acts = np.ascontiguousarray(actions, dtype=np.uint8).reshape(-1)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.uint32)
        rec = {k: np.zeros(self.N * self.A, dtype=np.dtype(v)) for k, v in abi.STEP_OUT_FIELDS.items()}
        cf = [None, None]
        if counterfactuals:  # _get_qrm_experiences' RM lookups on the new position: [N*A][n_qrm_states]
            nq = max(1, int(self.cfg.n_qrm_states))
            rec["cf_q"] = np.zeros(self.N * self.A * nq, dtype=np.uint8)
            rec["cf_r"] = np.zeros(self.N * self.A * nq, dtype=np.float64)
            cf = [_p(rec["cf_q"]), _p(rec["cf_r"])]
        so = abi.StepOut(*[_p(rec[k]) for k in abi.STEP_OUT_FIELDS], *cf)
        self.L.oracle_step(C.byref(self.cfg), C.byref(self.tables), C.byref(self.state), C.c_void_p(_p(acts)),
                           C.c_void_p(_p(d)), C.c_uint64(t), C.c_int(int(with_rm)), C.byref(so))
        return rec

    def rm_step(self, q, cell):
        q = np.ascontiguousarray(q, dtype=np.uint8).copy()
        cell = np.ascontiguousarray(cell, dtype=np.uint16)
        ev = np.zeros(q.size, dtype=np.uint8)
        r = np.zeros(q.size, dtype=np.float64)
        self.L.oracle_rm_step(C.byref(self.cfg), C.byref(self.tables), C.c_int64(q.size), C.c_void_p(_p(q)),
                              C.c_void_p(_p(cell)), C.c_void_p(_p(ev)), C.c_void_p(_p(r)))
        return q, ev, r

    def update(self, obs_cell, actions, term_arg, rec):
        obs = np.ascontiguousarray(obs_cell, dtype=np.uint16).reshape(-1)
        acts = np.ascontiguousarray(actions, dtype=np.uint8).reshape(-1)
        term = np.ascontiguousarray(term_arg, dtype=np.uint8).reshape(-1)
        so = abi.StepOut(*[_p(rec[k]) for k in abi.STEP_OUT_FIELDS])
        self.L.oracle_update(C.byref(self.cfg), C.byref(self.tables), C.byref(self.state), C.c_void_p(_p(obs)),
                             C.c_void_p(_p(acts)), C.c_void_p(_p(term)), C.byref(so))

    def train(self, t0, n_iters, learn=True, trace=False):
        tr = np.zeros((n_iters, self.N * self.A), dtype=np.uint32) if trace else None
        self.L.oracle_train(C.byref(self.cfg), C.byref(self.tables), C.byref(self.state), C.c_uint64(t0),
                            C.c_int32(n_iters), C.c_int32(int(learn)), C.c_void_p(_p(tr)))
        return tr

    def evaluate(self, n_episodes, gamma, optimal_steps, t0=0, max_iters=None):
        """Greedy evaluation on a COPY of the environment state (like copy.deepcopy(env) in the reference)."""
        ev = np.zeros(self.N * self.A, dtype=EVAL_DTYPE)
        ev["cum_gamma"] = 1.0
        slot, eps = self.slot.copy(), self.epsilon.copy()
        st = abi.State(self.N, _p(slot), _p(eps), _p(self.q), None, None, None, None, None, None, None)
        self.L.oracle_reset_at(C.byref(self.cfg), C.byref(self.tables), C.byref(st), None, C.c_uint64(t0))
        eps[:] = self.epsilon
        n_iters = max_iters or n_episodes * (self.cfg.max_steps + 1)
        self.L.oracle_evaluate(C.byref(self.cfg), C.byref(self.tables), C.byref(st), C.c_void_p(ev.ctypes.data), C.c_uint64(t0),
                               C.c_int32(n_iters), C.c_int32(n_episodes), C.c_double(gamma), C.c_double(optimal_steps))
        return ev

    def mdp(self, agent, sub_actions, rm_terminal=True):
        """RMEnvironmentWrapper.get_mdp restated (oracle_mdp): arrays indexed [s, a, j] + terminal kind per state."""
        sub = np.ascontiguousarray(sub_actions, dtype=np.uint8)
        n_sub = sub.shape[1]
        S = self.c.agent_rows[agent] if self.cfg.per_agent_rm else self.S
        nxt = np.zeros((S, 4, n_sub), dtype=np.int32)
        rew = np.zeros((S, 4, n_sub), dtype=np.float64)
        done = np.zeros((S, 4, n_sub), dtype=np.uint8)
        term = np.zeros(S, dtype=np.uint8)
        self.L.oracle_mdp(C.byref(self.cfg), C.byref(self.tables), C.c_int(agent), C.c_int(n_sub), C.c_void_p(_p(sub)),
                          C.c_int(int(rm_terminal)), C.c_void_p(_p(nxt)), C.c_void_p(_p(rew)), C.c_void_p(_p(done)), C.c_void_p(_p(term)))
        return nxt, rew, done, term

    def total_active_steps(self):
        """finished episodes' env.agent_steps (stats) + the running episodes' agent_steps (slot words)"""
        running = (self.slot >> np.uint64(abi.SLOT_STEPS_SHIFT)) & np.uint64(0xFFFF)
        return int(self.stats["active_steps"].sum()) + int(running.sum())

    # -- views ---------------------------------------------------------------------------------
    def unpack(self):
        return unpack_slots(self.slot, self.N, self.A)


def value_iteration(prob, next_state, reward, done, gamma=0.9, theta=1e-3, delta_rel=False):
    """mdp_vi.value_iteration restated (oracle_value_iteration, in-place Gauss-Seidel sweeps) on padded [S, 4, n_out] arrays."""
    prob = np.ascontiguousarray(prob, dtype=np.float64)
    S, _four, n_out = prob.shape
    nxt = np.ascontiguousarray(next_state, dtype=np.int32)
    rew = np.ascontiguousarray(reward, dtype=np.float64)
    dn = np.ascontiguousarray(done, dtype=np.uint8)
    V, Q, pol, old = np.zeros(S), np.zeros((S, 4)), np.zeros(S, dtype=np.int32), np.zeros(S)
    L = lib("f64")
    L.oracle_value_iteration.restype = C.c_int
    sweeps = L.oracle_value_iteration(C.c_int64(S), C.c_int(n_out), C.c_void_p(_p(prob)), C.c_void_p(_p(nxt)), C.c_void_p(_p(rew)),
                                      C.c_void_p(_p(dn)), C.c_double(gamma), C.c_double(theta), C.c_int(int(delta_rel)),
                                      C.c_void_p(_p(V)), C.c_void_p(_p(Q)), C.c_void_p(_p(pol)), C.c_void_p(_p(old)))
    return V, pol, Q, sweeps


def unpack_slots(slot, N, A):
    s = np.asarray(slot, dtype=np.uint64).reshape(N, A)
    return {
        "cell": ((s >> np.uint64(abi.SLOT_CELL_SHIFT)) & np.uint64(0xFFFF)).astype(np.int32),
        "agent_steps": ((s >> np.uint64(abi.SLOT_STEPS_SHIFT)) & np.uint64(0xFFFF)).astype(np.int32),
        "timestep": ((s >> np.uint64(abi.SLOT_TIME_SHIFT)) & np.uint64(0xFFFF)).astype(np.int32),
        "q": ((s >> np.uint64(abi.SLOT_RMSTATE_SHIFT)) & np.uint64(0xFF)).astype(np.int32),
        "flags": ((s >> np.uint64(abi.SLOT_FLAGS_SHIFT)) & np.uint64(0xFF)).astype(np.int32),
    }


def unpack_trace(tr, N, A):
    t = np.asarray(tr, dtype=np.uint32).reshape(-1, N, A)
    return {
        "action": (t & 7).astype(np.int32),
        "executed": ((t >> 3) & 7).astype(np.int32),
        "cell": ((t >> 6) & 1023).astype(np.int32),
        "q": ((t >> 16) & 31).astype(np.int32),
        "term": ((t >> 21) & 1).astype(np.int32),
        "trunc": ((t >> 22) & 1).astype(np.int32),
        "stepped": ((t >> 23) & 1).astype(np.int32),
    }
