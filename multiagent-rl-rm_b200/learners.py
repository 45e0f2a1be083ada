"""Tabular learners with the reference's object API, backed by the CUDA library.

Mirrors learning_algorithms/learning_algorithm.py:6-84, qlearning.py:6-160 and qlearning_lambda.py:5-132.
``q_table`` / ``e_table`` / ``visits`` are torch tensors on the GPU; ``update`` and ``choose_action`` launch the same
device functions the fused kernel uses (rlrm_update / rlrm_select_action on a one-slot handle whose "grid" is just
the encoded state space). There is no CPU fallback: constructing a learner without a CUDA device raises.

Randomness: the reference consumes a numpy PCG64 stream in data-dependent order, which cannot be reproduced by a
batched device generator; here every ``choose_action`` call consumes four 32-bit words (explore test, random action,
tie-break, unused) taken from ``rng`` — ``rng.words()`` if it has that hook (trace injection), else
``rng.integers(0, 2**32, 4)``. Decisions given the words are identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np
import torch

from . import _abi as abi
from ._lib import check, load


_EXP_STRUCT = struct.Struct("<IIBB6xd")   # rlrm_experience_t
_SEL_STRUCT = struct.Struct("<IIdIIIII")  # the input half of rlrm_select_req_t


class DeviceTable:
    """(S, A) view of a learner table living in GPU memory, with numpy-style indexing so reference-style code such as
    ``learner.q_table[s] = np.array([...])`` or ``np.isclose(learner.q_table[s, a], x)`` keeps working.
    ``.tensor`` is the underlying torch view (no copy); ``np.asarray(table)`` copies to the host."""

    def __init__(self, tensor):
        self.tensor = tensor

    shape = property(lambda self: tuple(self.tensor.shape))
    ndim = property(lambda self: self.tensor.dim())
    dtype = property(lambda self: np.dtype(str(self.tensor.dtype).replace("torch.", "")))

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, idx):
        r = self.tensor[idx]
        return r.item() if r.dim() == 0 else DeviceTable(r)

    def __setitem__(self, idx, value):
        if isinstance(value, DeviceTable):
            value = value.tensor
        self.tensor[idx] = torch.as_tensor(np.asarray(value) if not torch.is_tensor(value) else value,
                                           dtype=self.tensor.dtype, device=self.tensor.device)

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def copy(self):
        return np.array(self)

    def fill(self, value):
        self.tensor.fill_(value)

    def max(self):
        return self.tensor.max().item()

    def __repr__(self):
        return f"DeviceTable({np.array(self)!r}, device={self.tensor.device})"


def _factor_state_space(S):
    """(cells, nq) with cells * nq >= S, cells <= MAX_CELLS, nq <= MAX_RM_STATES."""
    nq = max(1, -(-S // abi.MAX_CELLS))
    if nq > abi.MAX_RM_STATES:
        raise ValueError(f"state_space_size {S} exceeds {abi.MAX_CELLS * abi.MAX_RM_STATES}")
    return -(-S // nq), nq


class _TableHandle:
    """One-slot C-ABI handle over a learner's own tables: grid = (cells x 1), nQ = nq, enc = cell*nq + q."""

    def __init__(self, S, algo, learning_rate, gamma, lambd, device, n_actions=4, table_dtype="f32"):
        if not torch.cuda.is_available():
            raise RuntimeError("multiagent-rl-rm_b200 learners need a CUDA device (no CPU fallback)")
        self.L = load()
        self.device = torch.device(device)
        self.cells, self.nq = _factor_state_space(S)
        self.S_pad = self.cells * self.nq
        cfg = abi.Config()
        cfg.abi_version = abi.ABI_VERSION
        cfg.env_kind, cfg.driver, cfg.algo = abi.ENV_FROZEN_LAKE, abi.DRIVER_OFFICE_MAIN, algo
        cfg.width, cfg.height, cfg.n_agents = self.cells, 1, 1
        cfg.n_rm_states, cfg.n_events, cfg.rm_final, cfg.n_qrm_states = self.nq, 0, -1, 0
        cfg.max_steps, cfg.stochastic, cfg.slip_n = 1000, 0, 1
        for j in range(3):
            cfg.slip_thr[j] = 1 << 32
        cfg.learning_rate = -1.0 if learning_rate is None else float(learning_rate)
        cfg.gamma, cfg.lambd = float(gamma), float(lambd)
        cfg.epsilon_start = cfg.epsilon_end = cfg.epsilon_decay = 1.0
        cfg.n_actions = n_actions
        cfg.table_dtype = abi.TABLE_F64 if table_dtype == "f64" else abi.TABLE_F32
        self.cfg = cfg
        n = self.cells
        self._arrs = dict(
            next_cell=np.repeat(np.arange(n, dtype=np.uint16)[:, None], 4, axis=1).copy(),
            cell_flags=np.zeros(n, np.uint8), label=np.full(n, abi.EVENT_NONE, np.uint8),
            delta=np.full((self.nq, 1), abi.NO_TRANSITION, np.uint8), rq=np.zeros((self.nq, 1)), rcf=np.zeros((self.nq, 1)),
            qrm_states=np.zeros(1, np.uint8), start_cell=np.zeros(1, np.uint16))
        t = abi.Tables(*[self._arrs[k].ctypes.data for k in ("next_cell", "cell_flags", "label", "delta", "rq", "rcf",
                                                            "qrm_states", "start_cell")], None, None)
        h = C.c_void_p()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self.L.rlrm_create(C.byref(cfg), C.byref(t), dev, C.byref(h)))
        self.h = h
        self._dev_index = dev

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.rlrm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def stream(self):
        """torch's current stream on this device as a raw cudaStream_t (one C call; the Stream object is not built)."""
        return torch._C._cuda_getCurrentRawStream(self._dev_index)


class BaseLearningAlgorithm:
    def __init__(self, state_space_size, action_space_size, gamma=0.99, seed=2020, max_steps=400, device="cuda:0",
                 table_dtype="f32"):
        """table_dtype (this package's extension): "f32" = float32 tables (what the batched kernels are tuned for; equals the
        reference with its tables cast to float32, bit for bit) or "f64" = the reference's own float64 tables
        (learning_algorithm.py / qlearning.py:26-29), bit-identical to the unmodified reference."""
        if table_dtype not in ("f32", "f64"):
            raise ValueError("table_dtype must be 'f32' or 'f64'")
        self.table_dtype = table_dtype
        if not (1 <= action_space_size <= 4):
            raise ValueError("the CUDA path implements the reference's grid worlds: at most 4 actions (up, down, left, right)")
        self.state_space_size = state_space_size
        self.action_space_size = action_space_size
        self.episode = 0
        self.max_steps = max_steps
        self.verbose = 0
        self.seed = seed
        self.gamma = gamma
        self.rleval = None
        self.user_quit = False
        self.agent_quit = False
        self.rng = np.random.default_rng(seed=self.seed)
        self.tosave = ["rng"]
        self.device = torch.device(device)

    def learn_init(self):
        pass

    def learn_init_episode(self):
        pass

    def learn_done_episode(self):
        pass

    def learn_end(self):
        pass


class _TabularBase(BaseLearningAlgorithm):
    _ALGO = abi.ALGO_QL
    # staging block layout (bytes): 0 slot u64 | 8 epsilon f64 | 16 draws 4 x u32 | 59 selected action (output) u8 |
    # 64..111 one rlrm_select_req_t (the look-ahead selection of rlrm_update_list_select) |
    # 1024.. a ring of _RING blocks of _BLOCK rlrm_experience_t (24 bytes each) for rlrm_update_list
    _RING, _BLOCK, _EXP_OFF, _SEL_OFF = 8, 16, 1024, 64
    _LOOKAHEAD = True  # class-wide switch of the look-ahead selection (tests compare on / off: identical decisions and tables)
    lookahead_hits = 0  # per learner: choose_action calls answered by the look-ahead / launched on their own
    lookahead_misses = 0
    _STAGE_BYTES = 1024 + 8 * 16 * 24

    def _setup(self, q_init, lambd=0.0):
        S = self.state_space_size
        self._th = _TableHandle(S, self._ALGO, self.learning_rate, self.gamma, lambd, self.device, self.action_space_size, self.table_dtype)
        d, Sp = self.device, self._th.S_pad
        dt = torch.float64 if self.table_dtype == "f64" else torch.float32
        self._q = torch.full((Sp, 4), float(q_init), dtype=dt, device=d)
        if self.action_space_size < 4:  # tables are 4 wide on the device; unused actions can never be a maximum
            self._q[:, self.action_space_size:] = float("-inf")
        self._e = torch.zeros((Sp, 4), dtype=dt, device=d) if self._ALGO == abi.ALGO_QLAMBDA else None
        self._visits = torch.zeros((Sp, 4), dtype=torch.int32, device=d)
        # One page-locked HOST staging block per learner holds everything a single call needs (slot word, epsilon, Philox
        # words, the selected action, a ring of experience lists). Page-locked memory is device-accessible under unified
        # addressing, so a call is: pack on the host -> ONE kernel launch that reads / writes the block in place -> (for a
        # selection) one stream synchronisation. No staging copies, no torch ops on the call path.
        self._stage = torch.zeros(self._STAGE_BYTES, dtype=torch.uint8).pin_memory()
        self._stage_np = self._stage.numpy()
        self._stage_u32 = self._stage_np.view(np.uint32)
        self._stage_mv = memoryview(self._stage_np)
        self._base = self._stage.data_ptr()
        self._ring = 0
        self._st = abi.State(1, self._base, self._base + 8, self._q.data_ptr(), None if self._e is None else self._e.data_ptr(),
                             self._visits.data_ptr(), None, None, None, None, None)
        self._st_ref = C.byref(self._st)
        self._hp = (self.learning_rate, self.gamma, lambd)
        # look-ahead selection (see _device_update_list): key of the request in flight, its sequence number, and the four
        # words already taken from self.rng for the NEXT choose_action call (consumed by it whether or not the look-ahead hits)
        self._spec = None
        self._spec_seq = 0
        self._spec_stream = None
        self.__dict__.setdefault("_pending_words", None)
        self._pending_rng = self.rng if self._pending_words is not None else None

    # tables as views of exactly (S, A)
    @property
    def q_table(self):
        self._spec = None  # the caller may write through the view: a selection computed ahead of that write is void
        return DeviceTable(self._q[: self.state_space_size, : self.action_space_size])

    @q_table.setter
    def q_table(self, value):
        self.q_table[:] = value

    @property
    def visits(self):
        return DeviceTable(self._visits[: self.state_space_size, : self.action_space_size])

    def _sync_hyper(self):
        hp = (self.learning_rate, self.gamma, getattr(self, "lambd", 0.0))
        if hp != self._hp:  # the reference lets callers mutate these attributes in place
            check(self._th.L.rlrm_set_learner(self._th.h, -1.0 if hp[0] is None else float(hp[0]), float(hp[1]), float(hp[2])))
            self._hp = hp

    def _split(self, enc):
        enc = int(enc)
        if not (0 <= enc < self.state_space_size):
            raise IndexError(f"encoded state {enc} out of range for state_space_size {self.state_space_size}")
        return enc // self._th.nq, enc % self._th.nq

    def _sync(self):
        check(self._th.L.rlrm_stream_sync(self._th.h, self._th.stream()))

    def _device_update_list(self, experiences, next_enc=None):
        """update_q / the Q(lambda) update for a list of (s, a, r, s', terminated) experiences, applied in order on the device
        by ONE launch per _BLOCK experiences (rlrm_update_list). Asynchronous: the next selection (or any read of the
        tables through torch, which runs on the same stream) is ordered after it.

        next_enc: the encoded state the driver loop will select in next (update_policy's next_state). When given, the LAST
        launch also evaluates that selection on the updated table (rlrm_update_list_select) with the four words the next
        choose_action call would draw from self.rng — taken now, kept as the pending words of that call — so that the call
        finds its answer in page-locked memory instead of launching and synchronising. Any other outcome (another state,
        another epsilon, a caller-supplied rng, a table write in between) falls back to the ordinary launch with the SAME
        words, so the stream of random words and every decision are identical with and without the look-ahead."""
        hp = (self.learning_rate, self.gamma, getattr(self, "lambd", 0.0))
        if hp != self._hp:
            self._sync_hyper()
        S, th, mv, base = self.state_space_size, self._th, self._stage_mv, self._base
        L, st_ref = th.L, self._st_ref
        stream = torch._C._cuda_getCurrentRawStream(th._dev_index)
        self._spec = None
        look = (self._LOOKAHEAD and next_enc is not None and self.action_selection == "greedy" and type(self.rng) is np.random.Generator
                and 0 <= next_enc < S)
        n_exp = len(experiences)
        last = n_exp - self._BLOCK
        pack_exp = _EXP_STRUCT.pack_into
        for lo in range(0, max(n_exp, 1), self._BLOCK):
            chunk = experiences[lo:lo + self._BLOCK] if n_exp > self._BLOCK else experiences
            off = self._EXP_OFF
            if chunk:
                blk = self._ring % self._RING
                if blk == 0 and self._ring:
                    self._sync()  # the ring wrapped: make sure the launches that read these blocks have finished
                off = o = self._EXP_OFF + blk * self._BLOCK * 24
                for e in chunk:  # (s, a, r, s', terminated, ...): the wrapper's ten-field QRM tuples are taken as they are
                    s_, a_, r_, sn_, done_ = int(e[0]), e[1], e[2], int(e[3]), e[4]
                    if not (0 <= s_ < S and 0 <= sn_ < S):
                        raise IndexError(f"encoded state {s_ if not 0 <= s_ < S else sn_} out of range for state_space_size {S}")
                    pack_exp(mv, o, s_, sn_, int(a_), 1 if done_ else 0, float(r_))
                    o += 24
                self._ring += 1
            if look and lo >= last:
                if self._pending_words is None or self._pending_rng is not self.rng:
                    self._pending_words = self._own_words()
                    self._pending_rng = self.rng
                seq = self._spec_seq = (self._spec_seq + 1) & 0xFFFFFFFF
                eps = float(self.epsilon)
                _SEL_STRUCT.pack_into(mv, self._SEL_OFF, int(next_enc), 0, eps, *self._pending_words, seq)
                check(L.rlrm_update_list_select(th.h, st_ref, 0, len(chunk), base + off, base + self._SEL_OFF, stream))
                self._spec, self._spec_stream = (int(next_enc), eps), stream
            elif chunk:
                check(L.rlrm_update_list(th.h, st_ref, 0, len(chunk), base + off, stream))

    def _device_update(self, s, sn, action, reward, terminated):
        self._device_update_list([(s, action, reward, sn, terminated)], next_enc=sn)

    _WORD_BLOCK = 64  # words taken from self.rng per numpy call (16 selections' worth)

    def _own_words(self):
        """The next four 32-bit words of self.rng's stream. They are fetched _WORD_BLOCK at a time — numpy fills an array with
        the same per-element routine as a scalar request, so the sequence of words is that of four-word requests — and handed
        out in order; a block drawn from a generator that has since been replaced is dropped."""
        buf = self.__dict__.get("_word_buf")
        if buf is None or buf[0] is not self.rng or buf[2] >= len(buf[1]):
            buf = self.__dict__["_word_buf"] = [self.rng, self.rng.integers(0, 1 << 32, size=self._WORD_BLOCK, dtype=np.uint64).tolist(), 0]
        k = buf[2]
        buf[2] = k + 4
        return buf[1][k:k + 4]

    @staticmethod
    def _raw_words(rng):
        """Four 32-bit words for one selection: the injected Philox words when the rng exposes them, else from the generator."""
        if hasattr(rng, "words"):
            return rng.words()
        return rng.integers(0, 1 << 32, size=4, dtype=np.uint64)

    def choose_action(self, encoded_state, best=False, rng=None, **kwargs):
        """epsilon-greedy with uniform tie-break (qlearning.py:112-143); best=True -> first argmax."""
        if not best and self.action_selection != "greedy":
            if self.action_selection == "softmax":
                raise NotImplementedError("softmax selection is out of scope (no reference driver uses it; DESIGN.md §8)")
            raise ValueError("Unsupported action selection method")
        own = rng is None or rng is self.rng
        spec, self._spec = self._spec, None
        if spec is not None and own and not best and self._pending_rng is self.rng and spec == (encoded_state, self.epsilon):
            # the look-ahead selection launched with the last update answers this call (same state — hence in range —, epsilon,
            # words, table)
            if self._stage_u32[(self._SEL_OFF + 40) >> 2] != self._spec_seq:
                check(self._th.L.rlrm_stream_sync(self._th.h, self._spec_stream))  # the stream that launch went to
            self._pending_words = None
            self.lookahead_hits += 1
            return int(self._stage_np[self._SEL_OFF + 36])
        cell, q = self._split(encoded_state)
        self.lookahead_misses += 1
        if best:
            words = [0, 0, 0, 0]
        elif own and self._pending_words is not None and self._pending_rng is self.rng:
            words, self._pending_words = self._pending_words, None  # drawn ahead for this very call
        else:
            if own:
                self._pending_words = None  # self.rng was replaced since: words of the old generator are void
            if own and type(self.rng) is np.random.Generator:
                words = self._own_words()
            else:
                words = [int(x) & 0xFFFFFFFF for x in self._raw_words(self.rng if rng is None else rng)]
        struct.pack_into("<QdIIII", self._stage_mv, 0, (cell << abi.SLOT_CELL_SHIFT) | (q << abi.SLOT_RMSTATE_SHIFT), float(self.epsilon), *words)
        th = self._th
        check(th.L.rlrm_select_action(th.h, C.byref(self._st), None if best else self._base + 16, 0, int(bool(best)), self._base + 59,
                                      th.stream()))
        self._sync()
        return int(self._stage_np[59])

    def choose_action_greedy(self, encoded_state, rng):
        """Uniform choice among the greedy actions (qlearning.py:136-143): the device selection with exploration switched off
        for this one call, tie-break word taken from `rng`."""
        eps, self.epsilon = self.epsilon, 0.0
        try:
            return self.choose_action(encoded_state, best=False, rng=rng)
        finally:
            self.epsilon = eps

    # -- pickling: office_main.py:1611-1613, 1922-1925 save / load the whole learner object with pickle ---------------
    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if k not in ("_th", "_q", "_e", "_visits", "_stage", "_stage_np", "_stage_u32", "_stage_mv", "_st", "_st_ref", "_base", "_ring", "_spec", "_spec_seq", "_spec_stream", "_pending_rng")}
        d["device"] = str(self.device)
        d["_tables"] = {"q": self._q.cpu().numpy(), "e": None if self._e is None else self._e.cpu().numpy(),
                        "visits": self._visits.cpu().numpy()}
        return d

    def __setstate__(self, d):
        tables = d.pop("_tables")
        self.__dict__.update(d)
        self.__dict__.setdefault("table_dtype", "f32")  # pickles written before the float64 table mode existed
        self.device = torch.device(self.device)
        self._setup(0.0, getattr(self, "lambd", 0.0))
        self._q.copy_(torch.from_numpy(tables["q"]).to(self._q.dtype))
        self._visits.copy_(torch.from_numpy(tables["visits"]))
        if self._e is not None and tables["e"] is not None:
            self._e.copy_(torch.from_numpy(tables["e"]))

    def learn_done_episode(self):
        """Decay epsilon after each episode (qlearning.py:153-155)."""
        self.epsilon = max(self.epsilon_end, self.epsilon * self.epsilon_decay)

    def reset_epsilon(self):
        self.epsilon = self.epsilon_start


class QLearning(_TabularBase):
    """Tabular Q-learning with optional QRM counterfactual experiences (qlearning.py:6-160)."""

    _ALGO = abi.ALGO_QL

    def __init__(self, gamma, action_selection, learning_rate=None, epsilon_start=1.0, epsilon_end=0.2, epsilon_decay=0.99,
                 qtable_init=1, use_qrm=False, use_rsh=False, **kwargs):
        super().__init__(gamma=gamma, **kwargs)
        self.learning_rate = learning_rate
        self.action_selection = action_selection
        self.epsilon_start, self.epsilon_end, self.epsilon_decay = epsilon_start, epsilon_end, epsilon_decay
        self.epsilon = self.epsilon_start
        self.use_qrm = use_qrm
        self.use_rsh = use_rsh
        self.qtable_init = qtable_init
        self.param_str = f"{learning_rate:0.2f},{gamma:0.2f}" if learning_rate is not None else f"None,{gamma:0.2f}"
        self.tosave += ["q_table", "visits", "epsilon"]
        self._setup(qtable_init)

    def update(self, encoded_state, encoded_next_state, action, reward, terminated, **kwargs):
        info = kwargs.get("info", {}) or {}
        rm = info.get("reward_machine", None)
        shaping = self.use_rsh and rm is not None and getattr(rm, "potentials", None) is not None
        if shaping:  # R' = R + gamma*Phi(q') - Phi(q)  (qlearning.py:51-66)
            pq, nq = info.get("prev_q", None), info.get("q", None)
            if pq is not None and nq is not None:
                reward += self.gamma * rm.potentials.get(nq, 0) - rm.potentials.get(pq, 0)
        if self.use_qrm:  # only the counterfactual list is applied (qlearning.py:82-106): one launch for the whole list
            todo = info.get("qrm_experience", [])
            if not isinstance(todo, (list, tuple)):
                todo = list(todo)
            if shaping:
                shaped = []
                for exp in todo:
                    _s, _a, _r, _sn, _done, _, cur_q, _, nxt_q, _ = exp
                    _r += self.gamma * rm.potentials.get(rm.get_state_from_index(nxt_q), 0) - rm.potentials.get(
                        rm.get_state_from_index(cur_q), 0)
                    shaped.append((_s, _a, _r, _sn, _done))
                todo = shaped
            self._device_update_list(todo, next_enc=encoded_next_state)
        else:
            self._device_update(encoded_state, encoded_next_state, action, reward, terminated)
        return False


class QLearningLambda(_TabularBase):
    """Watkins-style Q(lambda) with replacing traces, dense sweep (qlearning_lambda.py:5-132)."""

    _ALGO = abi.ALGO_QLAMBDA

    def __init__(self, gamma, lambd, action_selection, learning_rate=None, epsilon_start=1.0, epsilon_end=0.2,
                 epsilon_decay=0.99, **kwargs):
        super().__init__(gamma=gamma, **kwargs)
        self.learning_rate = learning_rate  # None -> 1 / visits[s, a], evaluated in float64 (qlearning_lambda.py:44-49)
        self.lambd = lambd
        self.action_selection = action_selection
        self.epsilon_start, self.epsilon_end, self.epsilon_decay = epsilon_start, epsilon_end, epsilon_decay
        self.epsilon = self.epsilon_start
        self.tosave += ["q_table", "visits", "epsilon"]
        self._setup(0.0, lambd)

    @property
    def e_table(self):
        return DeviceTable(self._e[: self.state_space_size, : self.action_space_size])

    def update(self, encoded_state, encoded_next_state, action, reward, terminated, next_action=None, **kwargs):
        self._device_update(encoded_state, encoded_next_state, action, reward, terminated)
        if not terminated and next_action is not None:
            # The "cut traces on an exploratory next action" branch (qlearning_lambda.py:82-84) is unreachable through
            # AgentRL.update_policy (next_action defaults to the argmax); honour it when a caller passes next_action.
            # The kernel applied the greedy-branch decay; an exploratory next action wipes the traces instead.
            row = self.q_table[int(encoded_next_state)]
            if not (row[int(next_action)] == row.max()):
                self._e.zero_()

    def learn_init_episode(self):
        self.reset_e_table()

    def reset_e_table(self):
        self._e.zero_()
