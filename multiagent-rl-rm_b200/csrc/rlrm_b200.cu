// rlrm_b200.cu — hand-written CUDA (sm_100a) for the lockstep hot path of multiagent-rl-rm and its C ABI
// (include/rlrm_b200.h). One thread per (environment instance, agent); the map / label / Reward-Machine tables are
// staged once per block in shared memory; Q rows are 16-byte vector gathers/scatters on per-instance tables in HBM;
// randomness is Philox4x32-10 on (iteration, instance, agent) counters; per-instance episode termination is a warp
// ballot over the instance's lane group. No tensor cores (nothing here is a contraction), no CPU fallback.
//
// Reference semantics restated per function (R/ = /root/reference/multiagent_rlrm/):
//   env.step      R/environments/frozen_lake/ma_frozen_lake.py:96-154,189-215 ; office_world/ma_office.py:122-257
//   RM step       R/multi_agent/reward_machine.py:45-59
//   wrapper       R/multi_agent/wrappers/rm_environment_wrapper.py:43-107,122-183
//   select        R/learning_algorithms/qlearning.py:112-143
//   update        R/learning_algorithms/qlearning.py:70-110 ; qlearning_lambda.py:33-84
//   driver loops  R/environments/frozen_lake/frozen_lake_main.py:336-376 ; office_world/office_main.py:1696-1749
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/rlrm_b200.h"

// ------------------------------------------------------------------------------------------------
// kernel parameter block (passed by value)
// ------------------------------------------------------------------------------------------------
struct KP {
  int env_kind, driver, algo;
  int A, G, g_shift;  // agents, lane-group size (power of two >= A), log2(G)
  int nQ, nEv, rm_final, n_qrm, max_steps, ncell;
  int stochastic, slip_n;
  int slip_cnt;              // thresholds below 2^32 (a threshold of 2^32 can never be reached by a 32-bit draw)
  unsigned slip_thr32[3];
  unsigned long long slip_thr[3];
  unsigned char slip_outcome[16];
  int terminate_on_plants, terminate_hit_walls;
  double hole_penalty, wall_penalty;
  double lr, gamma, eps_end, eps_decay;
  float lr_f, one_minus_lr_f, gamma_f, trace_decay_f;  // (float)lr, (float)(1-lr), (float)gamma, (float)(gamma*lambda)
  int decay_on_reset, shared_q, use_rsh, random_starts, n_free;
  // agents with different reward machines (rlrm_config_t.per_agent_rm): per-agent scalars and table strides
  int per_agent, a_nQ[RLRM_MAX_AGENTS], a_final[RLRM_MAX_AGENTS], a_nqrm[RLRM_MAX_AGENTS];
  long long a_prefix4[RLRM_MAX_AGENTS], sum4;  // float offset of agent a's table inside one instance, floats per instance
  int nd;                                       // nQmax * (nEv + 1): one agent's delta / rq / rcf section
  unsigned seed_lo, seed_hi, instance_offset, n_actions;
  unsigned rk[20];  // Philox round keys: rk[2r] = seed_lo + r*0x9E3779B9, rk[2r+1] = seed_hi + r*0xBB67AE85
  long long S4;  // W*H*nQ*4 floats per table
  // table blob in global memory and section offsets (bytes) inside it / inside the shared-memory copy
  const unsigned char* blob;
  int blob_bytes;
  int off_next, off_flags, off_label, off_delta, off_rq, off_rcf, off_qrm, off_start, off_phi, off_free;
};

struct Tab {
  const unsigned short* next_cell;
  const unsigned char* cell_flags;
  const unsigned char* label;
  const unsigned char* delta;
  const double* rq;
  const double* rcf;
  const unsigned char* qrm_states;
  const unsigned short* start_cell;
  const unsigned short* free_cells;
  const double* phi;
};

struct DState {  // rlrm_state_t by value
  long long N;
  unsigned long long* slot;
  double* epsilon;
  float* q;
  float* e;
  unsigned* visits;
  double* ep_return;
  rlrm_stats_t* stats;
  long long* acc_sum;  // shared learner accumulators (include/rlrm_b200.h "Shared learner"), null otherwise
  int* acc_cnt;
  float* acc_last;
  unsigned short* tr_pos;  // Q(lambda) sparse-exact traces (include/rlrm_b200.h), null otherwise
  unsigned short* tr_idx;
  float* tr_e;
  float* tr_q;
  unsigned* tr_len;
  unsigned long long* tr_work;
  int tr_cap;
};

struct Acc {  // accumulators of one agent's shared table, or nulls
  long long* sum;
  int* cnt;
  float* last;
};

struct DOut {  // rlrm_step_out_t by value
  unsigned short *prev_cell, *cell;
  unsigned char *prev_q, *q, *event, *executed;
  double *renv, *rq, *reward;
  unsigned char *env_term, *rm_term, *term, *trunc;
};

extern __shared__ __align__(16) unsigned char smem_raw[];

__device__ __forceinline__ Tab stage_tables(const KP& p) {
  // cooperative 16-byte copy of the (<= ~45 KB, typically 1-3 KB) table blob into shared memory
  const uint4* src = reinterpret_cast<const uint4*>(p.blob);
  uint4* dst = reinterpret_cast<uint4*>(smem_raw);
  for (int k = threadIdx.x; k < p.blob_bytes / 16; k += blockDim.x) dst[k] = __ldg(src + k);
  __syncthreads();
  Tab t;
  t.next_cell = reinterpret_cast<const unsigned short*>(smem_raw + p.off_next);
  t.cell_flags = smem_raw + p.off_flags;
  t.label = smem_raw + p.off_label;
  t.delta = smem_raw + p.off_delta;
  t.rq = reinterpret_cast<const double*>(smem_raw + p.off_rq);
  t.rcf = reinterpret_cast<const double*>(smem_raw + p.off_rcf);
  t.qrm_states = smem_raw + p.off_qrm;
  t.start_cell = reinterpret_cast<const unsigned short*>(smem_raw + p.off_start);
  t.phi = reinterpret_cast<const double*>(smem_raw + p.off_phi);
  t.free_cells = reinterpret_cast<const unsigned short*>(smem_raw + p.off_free);
  return t;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Random123): counter (t_lo, t_hi, instance, agent), key (seed_lo, seed_hi)
// ------------------------------------------------------------------------------------------------
// The 10 round keys (k0 + r*W0, k1 + r*W1) depend only on the seed: they are precomputed on the host into KP::rk so every
// round is two wide multiplies and two three-input XORs whose key operand comes straight from the constant bank.
#define RLRM_PHILOX(c0, c1, c2, c3, p, w) philox4x32_10_rk(c0, c1, c2, c3, (p).rk, w)
__device__ __forceinline__ void philox4x32_10_rk(unsigned c0, unsigned c1, unsigned c2, unsigned c3, const unsigned (&rk)[20],
                                                 unsigned w[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ rk[2 * r], n2 = (unsigned)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
    c1 = (unsigned)p1;
    c3 = (unsigned)p0;
    c0 = n0;
    c2 = n2;
  }
  w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

// ------------------------------------------------------------------------------------------------
// per-agent pieces
// ------------------------------------------------------------------------------------------------
struct Slot {
  unsigned cell, steps, time, rm, flags;
};
__device__ __forceinline__ Slot unpack_slot(unsigned long long w) {
  Slot s;
  s.cell = (unsigned)(w >> RLRM_SLOT_CELL_SHIFT) & 0xFFFFu;
  s.steps = (unsigned)(w >> RLRM_SLOT_STEPS_SHIFT) & 0xFFFFu;
  s.time = (unsigned)(w >> RLRM_SLOT_TIME_SHIFT) & 0xFFFFu;
  s.rm = (unsigned)(w >> RLRM_SLOT_RMSTATE_SHIFT) & 0xFFu;
  s.flags = (unsigned)(w >> RLRM_SLOT_FLAGS_SHIFT) & 0xFFu;
  return s;
}
__device__ __forceinline__ unsigned long long pack_slot(const Slot& s) {
  return ((unsigned long long)s.cell << RLRM_SLOT_CELL_SHIFT) | ((unsigned long long)s.steps << RLRM_SLOT_STEPS_SHIFT) |
         ((unsigned long long)s.time << RLRM_SLOT_TIME_SHIFT) | ((unsigned long long)s.rm << RLRM_SLOT_RMSTATE_SHIFT) |
         ((unsigned long long)s.flags << RLRM_SLOT_FLAGS_SHIFT);
}

// explore iff w0 / 2^32 < epsilon  <=>  w0 < ceil(epsilon * 2^32)   (w0 integer; the scaling by 2^32 is exact)
__device__ __forceinline__ unsigned long long explore_threshold(double eps) {
  if (!(eps > 0.0)) return 0ull;
  if (eps >= 1.0) return 1ull << 32;
  return (unsigned long long)ceil(eps * 4294967296.0);
}

// QLearning.choose_action / choose_action_greedy (qlearning.py:112-143)
__device__ __forceinline__ int select_action(const float4& row, unsigned long long explore_thr, const unsigned w[4], bool best,
                                             unsigned n_actions) {
  // np.argmax: first maximum
  int va = 0;
  float m = row.x;
  if (row.y > m) { m = row.y; va = 1; }
  if (row.z > m) { m = row.z; va = 2; }
  if (row.w > m) { m = row.w; va = 3; }
  if (best) return va;
  if ((unsigned long long)w[0] < explore_thr) return (int)__umulhi(w[1], n_actions);  // rng.choice(range(A))
  const int e0 = row.x == m, e1 = row.y == m, e2 = row.z == m, e3 = row.w == m;
  const int n = e0 + e1 + e2 + e3;
  if (n == 1) return va;
  const int pick = (int)__umulhi(w[2], (unsigned)n);  // rng.choice(maxs): the pick-th maximal index
  // rank of each maximal index among the maxima
  const int r1 = e0, r2 = e0 + e1, r3 = e0 + e1 + e2;
  int a = va;
  if (e1 && r1 == pick) a = 1;
  if (e2 && r2 == pick) a = 2;
  if (e3 && r3 == pick) a = 3;
  if (e0 && pick == 0) a = 0;
  return a;
}

__device__ __forceinline__ int slip_outcome(const KP& p, int intended, unsigned k) {
  int idx = 0;
#pragma unroll
  for (int j = 0; j < 3; j++) idx += (j < p.slip_cnt) && (k >= p.slip_thr32[j]);
  return p.slip_outcome[intended * 4 + idx];
}

struct Rec {
  unsigned prev_cell, cell, prev_q, q, event, executed;
  bool env_term, rm_term, term, trunc, stepped;
  double renv, rq, reward;
};

// env.step + check_terminations + RewardMachine.step + wrapper merge for ONE agent. Every quantity an agent needs is
// its own (the shared env.timestep is replicated per agent), so lanes never exchange data here.
template <int ENV, int STOCH = -1>  // STOCH: -1 = read p.stochastic at run time, 0/1 = compile-time
__device__ __forceinline__ void agent_step(const KP& p, const Tab& tb, Slot& s, int action, unsigned w3, bool with_rm, Rec& r) {
  const bool stochastic = STOCH < 0 ? (p.stochastic != 0) : (STOCH != 0);
  r.prev_cell = s.cell;
  r.executed = 5;
  r.stepped = false;
  r.renv = 0.0;
  r.rq = 0.0;
  const bool active = (s.flags & RLRM_FLAG_ACTIVE) != 0;
  if (ENV == RLRM_ENV_FROZEN_LAKE) {
    const bool rm_done = p.rm_final >= 0 && (int)s.rm == p.rm_final;  // ma_frozen_lake.py:107-115
    if (active && !rm_done) {
      int ex = action;
      if (stochastic) ex = slip_outcome(p, action, w3);
      if (ex != RLRM_ACTION_WAIT) s.cell = tb.next_cell[s.cell * 4 + ex];
      if (tb.cell_flags[s.cell] & 1) {  // holes_in_the_ice
        s.flags |= RLRM_FLAG_FAIL;
        r.renv = p.hole_penalty;
      }
      s.steps++;
      r.executed = ex;
      r.stepped = true;
    }
  } else {
    if (active) {  // ma_office.py:143-186
      int ex = action;
      double wall_pen = 0.0;
      if (tb.next_cell[s.cell * 4 + action] == s.cell) {  // apply_wall_penalty: blocked -> "wait", no slip draw
        if (p.terminate_hit_walls) s.flags |= RLRM_FLAG_FAIL;
        wall_pen = p.wall_penalty;
        ex = RLRM_ACTION_WAIT;
      }
      if (stochastic && ex != RLRM_ACTION_WAIT) ex = slip_outcome(p, ex, w3);
      if (ex != RLRM_ACTION_WAIT) s.cell = tb.next_cell[s.cell * 4 + ex];
      double plant = 0.0;
      if (tb.cell_flags[s.cell] & 1) {  // plants_in_the_office
        if (p.terminate_on_plants) s.flags |= RLRM_FLAG_FAIL;
        plant = p.hole_penalty;
      }
      r.renv = __dadd_rn(wall_pen, plant);
      s.steps++;
      r.executed = ex;
      r.stepped = true;
    }
  }
  r.cell = s.cell;
  s.time++;
  const bool fail = (s.flags & RLRM_FLAG_FAIL) != 0;
  if (ENV == RLRM_ENV_FROZEN_LAKE) {  // ma_frozen_lake.py:189-215 (RM state as of BEFORE this step's RM update)
    r.trunc = ((int)s.steps > p.max_steps) || ((int)s.time > p.max_steps);
    r.env_term = r.trunc || (p.rm_final >= 0 && (int)s.rm == p.rm_final) || fail;
    if (r.env_term) s.flags &= ~RLRM_FLAG_ACTIVE;
  } else {  // ma_office.py:240-257
    r.trunc = (int)s.time > p.max_steps;
    r.env_term = fail;
    if (r.env_term || r.trunc) s.flags &= ~RLRM_FLAG_ACTIVE;
  }
  // rm_environment_wrapper.py:57-107
  r.prev_q = s.rm;
  r.event = tb.label[s.cell];
  r.rm_term = false;
  if (with_rm) {
    const int col = r.event == RLRM_EVENT_NONE ? p.nEv : (int)r.event;
    const unsigned d = tb.delta[s.rm * (p.nEv + 1) + col];
    if (d != RLRM_NO_TRANSITION) {
      r.rq = tb.rq[s.rm * (p.nEv + 1) + col];
      s.rm = d;
    }
    r.rm_term = p.rm_final >= 0 && (int)s.rm == p.rm_final;
  }
  r.q = s.rm;
  r.reward = __dadd_rn(r.renv, r.rq);
  r.term = r.env_term || r.rm_term;
  s.flags &= ~(RLRM_FLAG_DONE | RLRM_FLAG_TRUNC | RLRM_FLAG_FIRST);
  if (r.term) s.flags |= RLRM_FLAG_DONE;
  if (r.trunc) s.flags |= RLRM_FLAG_TRUNC;
}

__device__ __forceinline__ void set_component(float4& v, unsigned c, float x) {
  if (c == 0) v.x = x;
  else if (c == 1) v.y = x;
  else if (c == 2) v.z = x;
  else v.w = x;
}

__device__ __forceinline__ float get_component(const float4& v, unsigned c) {
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

__device__ __forceinline__ float row_max(const float4& v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }

// update_q (qlearning.py:70-79) in the float32 arithmetic numpy performs on a float32 table: weak Python scalars are
// rounded to float32 first, every operation rounds separately (no FMA contraction).
__device__ __forceinline__ void update_q(const KP& p, float* Q, unsigned* V, unsigned s, int a, double r, unsigned sn, bool terminated,
                                         const Acc& acc) {
  const float cur = Q[s * 4 + a];
  const float4 nrow = *reinterpret_cast<const float4*>(Q + sn * 4);
  const float mf = __fmul_rn(terminated ? 0.0f : 1.0f, row_max(nrow));
  const float inner = __fadd_rn(__double2float_rn(r), __fmul_rn(p.gamma_f, mf));
  float out;
  if (p.lr < 0.0) {  // lr = 1/visits is an np.float64: the outer expression is evaluated in double
    const unsigned v = V[s * 4 + a] + 1;
    V[s * 4 + a] = v;
    const double lr = __ddiv_rn(1.0, (double)v);
    out = __double2float_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, lr), (double)cur), __dmul_rn(lr, (double)inner)));
  } else {
    if (V) V[s * 4 + a] += 1;
    out = __fadd_rn(__fmul_rn(p.one_minus_lr_f, cur), __fmul_rn(p.lr_f, inner));
  }
  if (acc.sum) {  // shared learner: propose; apply_shared_kernel turns the proposals of this iteration into their mean
    atomicAdd(reinterpret_cast<unsigned long long*>(acc.sum + s * 4 + a), (unsigned long long)__float2ll_rn(__fmul_rn(out, 1048576.0f)));
    atomicAdd(acc.cnt + s * 4 + a, 1);
    acc.last[s * 4 + a] = out;
  } else {
    Q[s * 4 + a] = out;
  }
}

// QL / QRM update of one agent (agent_rl.py:117-192 -> qlearning.py:41-110; QRM experiences rm_environment_wrapper.py:122-183)
template <int ALGO>
__device__ __forceinline__ void agent_update(const KP& p, const Tab& tb, float* Q, unsigned* V, unsigned obs_cell, int action,
                                             bool term_arg, const Rec& r, const Acc& acc) {
  if (ALGO == RLRM_ALGO_QRM) {
    const int col = r.event == RLRM_EVENT_NONE ? p.nEv : (int)r.event;
    for (int j = 0; j < p.n_qrm; j++) {
      const unsigned u = tb.qrm_states[j];
      const unsigned d = tb.delta[u * (p.nEv + 1) + col];
      const unsigned un = d == RLRM_NO_TRANSITION ? u : d;
      const double ru = d == RLRM_NO_TRANSITION ? 0.0 : tb.rcf[u * (p.nEv + 1) + col];
      const bool done = r.env_term || (p.rm_final >= 0 && (int)un == p.rm_final);
      double rew = __dadd_rn(r.renv, ru);
      if (p.use_rsh) rew = __dadd_rn(rew, __dsub_rn(__dmul_rn(p.gamma, tb.phi[un]), tb.phi[u]));  // qlearning.py:93-105
      update_q(p, Q, V, r.prev_cell * p.nQ + u, action, rew, r.cell * p.nQ + un, done, acc);
    }
  } else {
    double rew = r.reward;
    if (p.use_rsh) rew = __dadd_rn(rew, __dsub_rn(__dmul_rn(p.gamma, tb.phi[p.nQ + r.q]), tb.phi[p.nQ + r.prev_q]));  // qlearning.py:51-66
    update_q(p, Q, V, obs_cell * p.nQ + r.prev_q, action, rew, r.cell * p.nQ + r.q, term_arg, acc);
  }
}

// _sample_start_positions (ma_frozen_lake.py:156-172): agent a's start cell = entry a of a Fisher-Yates shuffle of the free
// cells driven by Philox words keyed on (T, instance) — see rlrm_config_t.random_starts. Every agent replays steps 0..a.
__device__ __forceinline__ unsigned sample_start(const KP& p, const Tab& tb, long long i, int a, unsigned long long T) {
  unsigned pos[RLRM_MAX_AGENTS], val[RLRM_MAX_AGENTS];
  unsigned w[4] = {0, 0, 0, 0}, out = 0;
  for (int k = 0; k <= a; k++) {
    if ((k & 3) == 0)
      RLRM_PHILOX((unsigned)T, ~(unsigned)(T >> 32), p.instance_offset + (unsigned)i, 0x80000000u | (unsigned)(k >> 2), p, w);
    const unsigned wk = (k & 3) == 0 ? w[0] : ((k & 3) == 1 ? w[1] : ((k & 3) == 2 ? w[2] : w[3]));
    const unsigned j = (unsigned)k + __umulhi(wk, (unsigned)(p.n_free - k));
    unsigned vk = tb.free_cells[k], vj = tb.free_cells[j];
    for (int m = 0; m < k; m++) {
      if (pos[m] == (unsigned)k) vk = val[m];
      if (pos[m] == j) vj = val[m];
    }
    out = vj;
    pos[k] = j;
    val[k] = vk;
  }
  return out;
}

// env.reset for one agent; T = iteration index of the new episode's first step (keys the random start positions)
template <bool RANDOM_STARTS = true>  // false: the caller guarantees cfg.random_starts == 0 (keeps the sampler out of hot kernels)
__device__ __forceinline__ void reset_slot(const KP& p, const Tab& tb, long long i, int a, unsigned long long T, Slot& s, double& eps) {
  s.cell = (RANDOM_STARTS && p.random_starts) ? sample_start(p, tb, i, a, T) : tb.start_cell[a];
  s.steps = 0;
  s.time = 0;
  s.rm = 0;  // the initial RM state has index 0 (reward_machine.py:32-36)
  s.flags = RLRM_FLAG_ACTIVE | RLRM_FLAG_FIRST;
  if (p.decay_on_reset) eps = fmax(p.eps_end, __dmul_rn(eps, p.eps_decay));  // learn_done_episode (qlearning.py:153-155)
}

__device__ __forceinline__ size_t table_base(const KP& p, long long i, int a) {
  if (p.per_agent) return (size_t)((p.shared_q ? 0ll : i * p.sum4) + p.a_prefix4[a]);
  return (size_t)(p.shared_q ? (long long)a : i * p.A + a) * (size_t)p.S4;
}

// Per-agent reward machines: turn the uniform parameter block / table pointers into agent a's own (nQ, final state, number
// of counterfactual states, table sections). Called only from kernels instantiated with PA = true; `p` must be the kernel's
// private copy of the parameter block, so with PA = false nothing here exists and the parameters stay in the constant bank.
__device__ __forceinline__ void agent_view(const KP& p_in, KP& p, Tab& tb, int a) {
  p.nQ = p_in.a_nQ[a];
  p.rm_final = p_in.a_final[a];
  p.n_qrm = p_in.a_nqrm[a];
  p.S4 = (long long)p_in.ncell * p.nQ * 4;
  tb.label += (size_t)a * p_in.ncell;
  tb.delta += (size_t)a * p_in.nd;
  tb.rq += (size_t)a * p_in.nd;
  tb.rcf += (size_t)a * p_in.nd;
  tb.qrm_states += (size_t)a * p_in.nQ;
}
__device__ __forceinline__ Acc make_acc(const KP& p, const DState& st, size_t base) {
  Acc acc = {nullptr, nullptr, nullptr};
  if (p.shared_q && st.acc_sum) {
    acc.sum = st.acc_sum + base;
    acc.cnt = st.acc_cnt + base;
    acc.last = st.acc_last + base;
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------
// unfused kernels (the reference's call-by-call API; also the parity path with injected draws)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reset_kernel(KP p, DState st, const unsigned char* mask, unsigned long long t) {
  Tab tb = stage_tables(p);
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= st.N * p.A) return;
  const long long i = k / p.A;
  const int a = (int)(k - i * p.A);
  if (mask && !mask[i]) return;
  Slot s;
  double eps = st.epsilon[k];
  reset_slot(p, tb, i, a, t, s, eps);
  st.slot[k] = pack_slot(s);
  st.epsilon[k] = eps;
  if (st.ep_return) st.ep_return[k] = 0.0;
}

// Q(lambda): reset_e_table for the masked instances (ma_frozen_lake.py:80-81 ; ma_office.py:101-102)
__global__ void __launch_bounds__(256) clear_traces_kernel(KP p, DState st, const unsigned char* mask) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 of one slot's table
  const long long per = p.S4 / 4;
  if (g >= st.N * p.A * per) return;
  const long long slot = g / per;
  if (mask && !mask[slot / p.A]) return;
  reinterpret_cast<float4*>(st.e)[g] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool PA>
__global__ void __launch_bounds__(256) select_kernel(KP p_in, DState st, const unsigned* draws, unsigned long long t, int best,
                                                    unsigned char* actions_out) {
  KP p = p_in;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= st.N * p.A) return;
  const long long i = k / p.A;
  const int a = (int)(k - i * p.A);
  if (PA) p.nQ = p_in.a_nQ[a];
  const Slot s = unpack_slot(st.slot[k]);
  const float* Q = st.q + table_base(p_in, i, a);
  const float4 row = *reinterpret_cast<const float4*>(Q + (size_t)(s.cell * p.nQ + s.rm) * 4);
  unsigned w[4];
  if (draws) {
    const uint4 d = reinterpret_cast<const uint4*>(draws)[k];
    w[0] = d.x; w[1] = d.y; w[2] = d.z; w[3] = d.w;
  } else {
    RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
  }
  actions_out[k] = (unsigned char)select_action(row, explore_threshold(st.epsilon[k]), w, best != 0, p.n_actions);
}

__device__ __forceinline__ void store_rec(const DOut& o, long long k, const Rec& r) {
  if (o.prev_cell) o.prev_cell[k] = (unsigned short)r.prev_cell;
  if (o.cell) o.cell[k] = (unsigned short)r.cell;
  if (o.prev_q) o.prev_q[k] = (unsigned char)r.prev_q;
  if (o.q) o.q[k] = (unsigned char)r.q;
  if (o.event) o.event[k] = (unsigned char)r.event;
  if (o.executed) o.executed[k] = (unsigned char)r.executed;
  if (o.renv) o.renv[k] = r.renv;
  if (o.rq) o.rq[k] = r.rq;
  if (o.reward) o.reward[k] = r.reward;
  if (o.env_term) o.env_term[k] = r.env_term;
  if (o.rm_term) o.rm_term[k] = r.rm_term;
  if (o.term) o.term[k] = r.term;
  if (o.trunc) o.trunc[k] = r.trunc;
}

template <int ENV, bool PA>
__global__ void __launch_bounds__(256) step_kernel(KP p_in, DState st, const unsigned char* actions, const unsigned* draws,
                                                  unsigned long long t, int with_rm, DOut out) {
  KP p = p_in;
  Tab tb = stage_tables(p_in);
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= st.N * p.A) return;
  const long long i = k / p.A;
  const int a = (int)(k - i * p.A);
  if (PA) agent_view(p_in, p, tb, a);
  Slot s = unpack_slot(st.slot[k]);
  unsigned w3 = 0;
  if (p.stochastic) {
    if (draws) {
      w3 = draws[k * 4 + 3];
    } else {
      unsigned w[4];
      RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
      w3 = w[3];
    }
  }
  Rec r;
  agent_step<ENV>(p, tb, s, actions[k], w3, with_rm != 0, r);
  st.slot[k] = pack_slot(s);
  store_rec(out, k, r);
}

// RewardMachine.step on explicit (state, position) pairs (reward_machine.py:45-59)
__global__ void __launch_bounds__(256) rm_step_kernel(KP p_in, int agent, long long n, unsigned char* q, const unsigned short* cell,
                                                     unsigned char* event_out, double* reward_out) {
  KP p = p_in;
  Tab tb = stage_tables(p_in);
  if (p_in.per_agent) agent_view(p_in, p, tb, agent);  // the reward machine of agent `agent`
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned ev = tb.label[cell[k]];
  const int col = ev == RLRM_EVENT_NONE ? p.nEv : (int)ev;
  const unsigned cur = q[k];
  const unsigned d = tb.delta[cur * (p.nEv + 1) + col];
  double r = 0.0;
  if (d != RLRM_NO_TRANSITION) {
    r = tb.rq[cur * (p.nEv + 1) + col];
    q[k] = (unsigned char)d;
  }
  if (event_out) event_out[k] = (unsigned char)ev;
  if (reward_out) reward_out[k] = r;
}

template <int ALGO, bool PA>
__global__ void __launch_bounds__(256) update_kernel(KP p_in, DState st, const unsigned short* obs_cell, const unsigned char* actions,
                                                    const unsigned char* term_arg, DOut o) {
  KP p = p_in;
  Tab tb = stage_tables(p_in);
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= st.N * p.A) return;
  const long long i = k / p.A;
  const int a = (int)(k - i * p.A);
  if (PA) agent_view(p_in, p, tb, a);
  Rec r;
  r.prev_cell = o.prev_cell[k]; r.cell = o.cell[k]; r.prev_q = o.prev_q[k]; r.q = o.q[k]; r.event = o.event[k];
  r.env_term = o.env_term[k] != 0; r.renv = o.renv[k]; r.reward = o.reward[k];
  const size_t base = table_base(p_in, i, a);
  agent_update<ALGO>(p, tb, st.q + base, st.visits ? st.visits + base : nullptr, obs_cell[k], actions[k], term_arg[k] != 0, r,
                     make_acc(p, st, base));
}

// QLearningLambda.update (qlearning_lambda.py:33-84), dense sweep exactly as written: one block per (instance, agent).
// q += (lr*td) * e ; e = terminated ? 0 : e * (gamma*lambda), with e[s,a] replaced by 1 first.
__device__ __forceinline__ void qlambda_sweep(const KP& p, float* Q, float* E, unsigned s, int a, double reward, unsigned sn,
                                              bool terminated, int tid, int nthreads) {
  // every thread reads the two scalars before anyone writes
  const float4 nrow = *reinterpret_cast<const float4*>(Q + sn * 4);
  const float qsa = Q[s * 4 + a];
  __syncthreads();
  const double best = terminated ? 0.0 : (double)row_max(nrow);
  const float td = __fsub_rn(__double2float_rn(__dadd_rn(reward, __dmul_rn(p.gamma, best))), qsa);
  const float c = __fmul_rn(p.lr_f, td);
  const unsigned hot = s * 4 + a;
  float4* Q4 = reinterpret_cast<float4*>(Q);
  float4* E4 = reinterpret_cast<float4*>(E);
  for (long long j = tid; j < p.S4 / 4; j += nthreads) {
    float4 e = E4[j], q = Q4[j];
    if ((unsigned)j == (hot >> 2)) set_component(e, hot & 3, 1.0f);  // replacing trace
    q.x = __fadd_rn(q.x, __fmul_rn(c, e.x));
    q.y = __fadd_rn(q.y, __fmul_rn(c, e.y));
    q.z = __fadd_rn(q.z, __fmul_rn(c, e.z));
    q.w = __fadd_rn(q.w, __fmul_rn(c, e.w));
    if (terminated) {
      e = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {  // next_action defaults to argmax Q[s'] => greedy => decay (qlearning_lambda.py:71-81)
      e.x = __fmul_rn(e.x, p.trace_decay_f);
      e.y = __fmul_rn(e.y, p.trace_decay_f);
      e.z = __fmul_rn(e.z, p.trace_decay_f);
      e.w = __fmul_rn(e.w, p.trace_decay_f);
    }
    Q4[j] = q;
    E4[j] = e;
  }
}

__global__ void __launch_bounds__(256) update_qlambda_kernel(KP p, DState st, const unsigned short* obs_cell,
                                                            const unsigned char* actions, const unsigned char* term_arg, DOut o) {
  const long long k = blockIdx.x;
  const long long i = k / p.A;
  const int a = (int)(k - i * p.A);
  const size_t base = table_base(p, i, a);
  if (st.visits && threadIdx.x == 0) st.visits[base + (size_t)(obs_cell[k] * p.nQ + o.prev_q[k]) * 4 + actions[k]] += 1;
  qlambda_sweep(p, st.q + base, st.e + base, obs_cell[k] * p.nQ + o.prev_q[k], actions[k], o.reward[k],
                o.cell[k] * p.nQ + o.q[k], term_arg[k] != 0, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------------------------------------
// fused persistent kernel: n_iters lockstep iterations, state in registers
// ------------------------------------------------------------------------------------------------
#define TRAIN_BLOCK 128

template <int ENV, int ALGO, bool PA>
__global__ void __launch_bounds__(TRAIN_BLOCK) train_kernel(KP p_in, DState st, unsigned long long t0, int n_iters, int learn,
                                                           unsigned* trace) {
  KP p = p_in;
  Tab tb = stage_tables(p_in);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = tid >> p.g_shift;
  const int a = (int)(tid & (p.G - 1));
  const bool valid = (i < st.N) && (a < p.A);
  const long long k = i * p.A + a;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned group_mask = (p.G == 32 ? 0xFFFFFFFFu : ((1u << p.G) - 1u)) << (lane & ~(unsigned)(p.G - 1));

  Slot s = {0, 0, 0, 0, 0};
  double eps = 0.0, ep_ret = 0.0;
  float* Q = nullptr;
  unsigned* V = nullptr;
  Acc acc = {nullptr, nullptr, nullptr};
  unsigned long long active_steps = 0;
  unsigned episodes = 0, successes = 0, last_length = 0;
  double return_sum = 0.0;
  float last_return = 0.f;
  if (valid) {
    s = unpack_slot(st.slot[k]);
    eps = st.epsilon[k];
    if (st.ep_return) ep_ret = st.ep_return[k];
    if (st.stats) return_sum = st.stats[k].return_sum;
    if (PA) agent_view(p_in, p, tb, a);
    const size_t base = table_base(p_in, i, a);
    Q = st.q + base;
    V = st.visits ? st.visits + base : nullptr;
    acc = make_acc(p, st, base);
  }
  unsigned long long explore_thr = explore_threshold(eps);
  bool had_episode = false;
  // plain QL with a private table, fixed learning rate and no visit counts: carry the current row across iterations
  const bool carry = (ALGO == RLRM_ALGO_QL) && !V && !acc.sum && p.lr >= 0.0;
  float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned row_idx = 0xFFFFFFFFu;

  for (int it = 0; it < n_iters; it++) {
    const unsigned long long t = t0 + (unsigned long long)it;
    bool term = true, trunc = true;
    Rec r;
    int action = 0;
    if (valid) {
      unsigned w[4];
      RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
      // every agent selects on every iteration, finished ones included (frozen_lake_main.py:350-352)
      const unsigned cur_idx = s.cell * p.nQ + s.rm;
      if (cur_idx != row_idx) {  // plain QL carries the row of its current state in registers (1-entry cache of Q)
        row = *reinterpret_cast<const float4*>(Q + (size_t)cur_idx * 4);
        row_idx = carry ? cur_idx : 0xFFFFFFFFu;
      }
      action = select_action(row, explore_thr, w, learn == 0, p.n_actions);
      const unsigned before = s.cell;
      const bool first = (s.flags & RLRM_FLAG_FIRST) != 0;
      agent_step<ENV>(p, tb, s, action, w[3], true, r);
      if (learn) {
        // FrozenLake driver: on an episode's first iteration `states` still aliases agent.state, so update_policy
        // receives the NEW position as `state` (frozen_lake_main.py:337,359 ; office_main.py:1700 deep-copies)
        const unsigned obs = (p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN && first) ? r.cell : before;
        const bool term_arg = p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN ? (r.term || r.trunc) : r.term;
        if (ALGO == RLRM_ALGO_QL && carry) {
          // update_q (qlearning.py:70-79) against the carried row: normally Q[s] is the carried row and only Q[s'] is loaded
          const unsigned sidx = obs * p.nQ + r.prev_q, snidx = r.cell * p.nQ + r.q;
          float4 nrow = (snidx == row_idx) ? row : *reinterpret_cast<const float4*>(Q + (size_t)snidx * 4);
          const float cur = (sidx == row_idx) ? get_component(row, action)
                                               : ((sidx == snidx) ? get_component(nrow, action) : Q[(size_t)sidx * 4 + action]);
          double rew = r.reward;
          if (p.use_rsh) rew = __dadd_rn(rew, __dsub_rn(__dmul_rn(p.gamma, tb.phi[p.nQ + r.q]), tb.phi[p.nQ + r.prev_q]));
          const float mf = __fmul_rn(term_arg ? 0.0f : 1.0f, row_max(nrow));
          const float inner = __fadd_rn(__double2float_rn(rew), __fmul_rn(p.gamma_f, mf));
          const float out = __fadd_rn(__fmul_rn(p.one_minus_lr_f, cur), __fmul_rn(p.lr_f, inner));
          if (__float_as_uint(out) != __float_as_uint(cur)) Q[(size_t)sidx * 4 + action] = out;
          if (sidx == row_idx) set_component(row, action, out);
          if (snidx != row_idx) {  // the next state's row becomes the carried one
            if (sidx == snidx) set_component(nrow, action, out);
            row = nrow;
            row_idx = snidx;
          }
        } else {
          agent_update<ALGO>(p, tb, Q, V, obs, action, term_arg, r, acc);
        }
      }
      term = r.term;
      trunc = r.trunc;
      ep_ret = __dadd_rn(ep_ret, r.reward);
      if (trace)
        trace[(size_t)it * (size_t)(st.N * p.A) + (size_t)k] = (unsigned)action | (r.executed << 3) | (r.cell << 6) | (r.q << 16) |
                                                             ((unsigned)r.term << 21) | ((unsigned)r.trunc << 22) |
                                                             ((unsigned)r.stepped << 23);
    }
    // episode over when all agents terminated, or all truncated (frozen_lake_main.py:345,375 ; office_main.py:1748)
    const unsigned bt = __ballot_sync(0xFFFFFFFFu, term), bc = __ballot_sync(0xFFFFFFFFu, trunc);
    const bool over = ((bt & group_mask) == group_mask) || ((bc & group_mask) == group_mask);
    if (valid && over) {
      episodes++;
      active_steps += s.steps;  // env.agent_steps[agent] of the finished episode
      successes += (p.rm_final >= 0 && (int)s.rm == p.rm_final) ? 1u : 0u;
      last_return = __double2float_rn(ep_ret);
      return_sum = __dadd_rn(return_sum, ep_ret);
      last_length = s.time;
      had_episode = true;
      ep_ret = 0.0;
      reset_slot(p, tb, i, a, t + 1, s, eps);  // next episode starts with rm_env.reset (frozen_lake_main.py:337)
      explore_thr = explore_threshold(eps);
    }
  }
  if (valid) {
    st.slot[k] = pack_slot(s);
    st.epsilon[k] = eps;
    if (st.ep_return) st.ep_return[k] = ep_ret;
    if (st.stats) {
      rlrm_stats_t z = st.stats[k];
      z.active_steps += active_steps;
      z.episodes += episodes;
      z.successes += successes;
      z.return_sum = return_sum;
      if (had_episode) {
        z.last_return = last_return;
        z.last_length = last_length;
      }
      st.stats[k] = z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fast path: QRM with nQ == 4 (BASELINE configs 1/3/5). The 64-byte cell block Q[cell, 0..3, 0..3] lives in registers
// between iterations: it is fetched with two 256-bit loads only when the agent changes cell, the counterfactual
// updates run on registers, and the new values go back as scalar stores. Requires
// qrm_states == [0, 1, 2] (so a static unroll over rows is the reference's update order), per-instance
// tables, fixed learning rate, no visit counts; anything else takes train_kernel.
// ------------------------------------------------------------------------------------------------
struct __align__(32) F8 {
  float v[8];
};
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ float sel4(float a, float b, float c, float d, unsigned k) {
  const float lo = (k & 1u) ? b : a, hi = (k & 1u) ? d : c;
  return (k & 2u) ? hi : lo;
}
__device__ __forceinline__ void load_block4(const float* Q, unsigned cell, float B[16], float bmax[4]) {
  const F8 lo = ldg256(Q + (size_t)cell * 16), hi = ldg256(Q + (size_t)cell * 16 + 8);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    B[j] = lo.v[j];
    B[8 + j] = hi.v[j];
  }
#pragma unroll
  for (int r = 0; r < 4; r++) bmax[r] = fmaxf(fmaxf(B[4 * r], B[4 * r + 1]), fmaxf(B[4 * r + 2], B[4 * r + 3]));
}

// STOCH / LEARN / TRACE are compile-time copies of p.stochastic / learn / (trace != nullptr): the loop body is issue-bound,
// so runtime flag tests and their constant-bank loads are specialised away. n_qrm is 3 on this path.
template <int ENV, bool STOCH, bool LEARN, bool TRACE>
__global__ void __launch_bounds__(TRAIN_BLOCK, 7) train_qrm4_kernel(KP p, DState st, unsigned long long t0, int n_iters,
                                                                unsigned* trace) {
  Tab tb = stage_tables(p);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = tid >> p.g_shift;
  const int a = (int)(tid & (p.G - 1));
  const bool valid = (i < st.N) && (a < p.A);
  const long long k = i * p.A + a;

  Slot s = {0, 0, 0, 0, 0};
  double eps = 0.0, ep_ret = 0.0, return_sum = 0.0;
  float* Q = st.q;
  unsigned long long active_steps = 0;
  unsigned episodes = 0, successes = 0, last_length = 0;
  float last_return = 0.f;
  float B[16], bmax[4];  // carried cell block Q[cell, rm state, action] and its row maxima
#pragma unroll
  for (int j = 0; j < 16; j++) B[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; j++) bmax[j] = 0.f;
  if (valid) {
    s = unpack_slot(st.slot[k]);
    eps = st.epsilon[k];
    if (st.ep_return) ep_ret = st.ep_return[k];
    if (st.stats) return_sum = st.stats[k].return_sum;
    Q = st.q + table_base(p, i, a);
    load_block4(Q, s.cell, B, bmax);
  }
  unsigned long long explore_thr = explore_threshold(eps);
  bool had_episode = false;

  for (int it = 0; it < n_iters; it++) {
    const unsigned long long t = t0 + (unsigned long long)it;
    bool term = true, trunc = true;
    if (valid) {
      unsigned w[4];
      RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
      float4 row;
      row.x = sel4(B[0], B[4], B[8], B[12], s.rm);
      row.y = sel4(B[1], B[5], B[9], B[13], s.rm);
      row.z = sel4(B[2], B[6], B[10], B[14], s.rm);
      row.w = sel4(B[3], B[7], B[11], B[15], s.rm);
      const int action = select_action(row, explore_thr, w, !LEARN, p.n_actions);
      const unsigned before = s.cell;
      Rec r;
      agent_step<ENV, STOCH>(p, tb, s, action, w[3], true, r);
      const bool moved = r.cell != before;
      // values the updates overwrite, read before the carried block is replaced
      const float cur0 = sel4(B[0], B[1], B[2], B[3], (unsigned)action);
      const float cur1 = sel4(B[4], B[5], B[6], B[7], (unsigned)action);
      const float cur2 = sel4(B[8], B[9], B[10], B[11], (unsigned)action);
      if (moved) load_block4(Q, r.cell, B, bmax);  // the carried block becomes the NEXT cell's block
      if (LEARN) {
        // QRM counterfactual experiences (rm_environment_wrapper.py:122-183) applied by update_q (qlearning.py:70-106),
        // in get_all_states()[:-1] order == row order 0..n_qrm-1 on this path. The next state's row maximum comes from
        // the carried block: a different block when the agent moved, else the live one including earlier updates.
        const int col = r.event == RLRM_EVENT_NONE ? p.nEv : (int)r.event;
        float* dst = Q + (size_t)before * 16 + action;  // infos["prev_s"] is the position before the move
#pragma unroll
        for (int u = 0; u < 3; u++) {
          {
            const unsigned d = tb.delta[u * (p.nEv + 1) + col];
            const unsigned un = d == RLRM_NO_TRANSITION ? (unsigned)u : d;
            const double ru = d == RLRM_NO_TRANSITION ? 0.0 : tb.rcf[u * (p.nEv + 1) + col];
            const bool done = r.env_term || (p.rm_final >= 0 && (int)un == p.rm_final);
            const float mx = sel4(bmax[0], bmax[1], bmax[2], bmax[3], un);
            const float cur = u == 0 ? cur0 : (u == 1 ? cur1 : cur2);
            const float mf = __fmul_rn(done ? 0.0f : 1.0f, mx);
            const float inner = __fadd_rn(__double2float_rn(__dadd_rn(r.renv, ru)), __fmul_rn(p.gamma_f, mf));
            const float nv = __fadd_rn(__fmul_rn(p.one_minus_lr_f, cur), __fmul_rn(p.lr_f, inner));
            if (__float_as_uint(nv) != __float_as_uint(cur)) dst[4 * u] = nv;  // a bit-identical value needs no store
            if (!moved) {  // same cell: the carried block is the one just written
#pragma unroll
              for (int c = 0; c < 4; c++) B[4 * u + c] = (c == action) ? nv : B[4 * u + c];
              bmax[u] = fmaxf(fmaxf(B[4 * u], B[4 * u + 1]), fmaxf(B[4 * u + 2], B[4 * u + 3]));
            }
          }
        }
      }
      term = r.term;
      trunc = r.trunc;
      ep_ret = __dadd_rn(ep_ret, r.reward);
      if (TRACE)
        trace[(size_t)it * (size_t)(st.N * p.A) + (size_t)k] = (unsigned)action | (r.executed << 3) | (r.cell << 6) | (r.q << 16) |
                                                             ((unsigned)r.term << 21) | ((unsigned)r.trunc << 22) |
                                                             ((unsigned)r.stepped << 23);
    }
    // episode over <=> every agent of the instance terminated, or every agent truncated: AND-reduce the two flags
    // (packed in one word) over the instance's lane group with xor shuffles
    unsigned flags2 = (term ? 1u : 0u) | (trunc ? 2u : 0u);
    for (int o = 1; o < p.G; o <<= 1) flags2 &= __shfl_xor_sync(0xFFFFFFFFu, flags2, o);
    const bool over = flags2 != 0u;
    if (valid && over) {
      episodes++;
      active_steps += s.steps;  // env.agent_steps[agent] of the finished episode
      successes += (p.rm_final >= 0 && (int)s.rm == p.rm_final) ? 1u : 0u;
      last_return = __double2float_rn(ep_ret);
      return_sum = __dadd_rn(return_sum, ep_ret);
      last_length = s.time;
      had_episode = true;
      ep_ret = 0.0;
      reset_slot<false>(p, tb, i, a, t + 1, s, eps);
      explore_thr = explore_threshold(eps);
      load_block4(Q, s.cell, B, bmax);
    }
  }
  if (valid) {
    st.slot[k] = pack_slot(s);
    st.epsilon[k] = eps;
    if (st.ep_return) st.ep_return[k] = ep_ret;
    if (st.stats) {
      rlrm_stats_t z = st.stats[k];
      z.active_steps += active_steps;
      z.episodes += episodes;
      z.successes += successes;
      z.return_sum = return_sum;
      if (had_episode) {
        z.last_return = last_return;
        z.last_length = last_length;
      }
      st.stats[k] = z;
    }
  }
}

// Q(lambda) fused: one block per instance, one warp per agent; the dense trace sweep is cooperative over the warp.
template <int ENV>
__global__ void __launch_bounds__(256) train_qlambda_kernel(KP p, DState st, unsigned long long t0, int n_iters, int learn,
                                                           unsigned* trace) {
  Tab tb = stage_tables(p);
  __shared__ int sh_term[RLRM_MAX_AGENTS], sh_trunc[RLRM_MAX_AGENTS];
  const long long i = blockIdx.x;
  const int a = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long k = i * p.A + a;
  Slot s = unpack_slot(st.slot[k]);
  double eps = st.epsilon[k];
  double ep_ret = st.ep_return ? st.ep_return[k] : 0.0;
  const size_t base = table_base(p, i, a);
  float* Q = st.q + base;
  float* E = st.e + base;
  unsigned long long active_steps = 0;
  unsigned episodes = 0, successes = 0, last_length = 0;
  float last_return = 0.f;
  bool had_episode = false;
  rlrm_stats_t z;
  if (st.stats) z = st.stats[k];
  double return_sum_add = st.stats ? z.return_sum : 0.0;
  unsigned long long explore_thr = explore_threshold(eps);

  for (int it = 0; it < n_iters; it++) {
    const unsigned long long t = t0 + (unsigned long long)it;
    unsigned w[4];
    RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
    __syncwarp();
    const float4 row = *reinterpret_cast<const float4*>(Q + (size_t)(s.cell * p.nQ + s.rm) * 4);
    const int action = select_action(row, explore_thr, w, learn == 0, p.n_actions);
    const unsigned before = s.cell;
    const bool first = (s.flags & RLRM_FLAG_FIRST) != 0;
    Rec r;
    agent_step<ENV>(p, tb, s, action, w[3], true, r);  // all 32 lanes compute the same scalars
    if (learn) {
      const unsigned obs = (p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN && first) ? r.cell : before;
      const bool term_arg = p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN ? (r.term || r.trunc) : r.term;
      // warp-cooperative dense sweep (the __syncthreads inside qlambda_sweep is replaced by __syncwarp here)
      const unsigned sidx = obs * p.nQ + r.prev_q, snidx = r.cell * p.nQ + r.q;
      const float4 nrow = *reinterpret_cast<const float4*>(Q + snidx * 4);
      const float qsa = Q[sidx * 4 + action];
      __syncwarp();
      const double best = term_arg ? 0.0 : (double)row_max(nrow);
      const float td = __fsub_rn(__double2float_rn(__dadd_rn(r.reward, __dmul_rn(p.gamma, best))), qsa);
      const float c = __fmul_rn(p.lr_f, td);
      const unsigned hot = sidx * 4 + action;
      float4* Q4 = reinterpret_cast<float4*>(Q);
      float4* E4 = reinterpret_cast<float4*>(E);
      for (long long j = lane; j < p.S4 / 4; j += 32) {
        float4 e = E4[j], q = Q4[j];
        if ((unsigned)j == (hot >> 2)) set_component(e, hot & 3, 1.0f);
        q.x = __fadd_rn(q.x, __fmul_rn(c, e.x));
        q.y = __fadd_rn(q.y, __fmul_rn(c, e.y));
        q.z = __fadd_rn(q.z, __fmul_rn(c, e.z));
        q.w = __fadd_rn(q.w, __fmul_rn(c, e.w));
        if (term_arg) {
          e = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          e.x = __fmul_rn(e.x, p.trace_decay_f);
          e.y = __fmul_rn(e.y, p.trace_decay_f);
          e.z = __fmul_rn(e.z, p.trace_decay_f);
          e.w = __fmul_rn(e.w, p.trace_decay_f);
        }
        Q4[j] = q;
        E4[j] = e;
      }
      __syncwarp();
    }
    ep_ret = __dadd_rn(ep_ret, r.reward);
    if (trace && lane == 0)
      trace[(size_t)it * (size_t)(st.N * p.A) + (size_t)k] = (unsigned)action | (r.executed << 3) | (r.cell << 6) | (r.q << 16) |
                                                           ((unsigned)r.term << 21) | ((unsigned)r.trunc << 22) |
                                                           ((unsigned)r.stepped << 23);
    if (lane == 0) {
      sh_term[a] = r.term;
      sh_trunc[a] = r.trunc;
    }
    __syncthreads();
    bool all_term = true, all_trunc = true;
    for (int b = 0; b < p.A; b++) {
      all_term = all_term && sh_term[b];
      all_trunc = all_trunc && sh_trunc[b];
    }
    __syncthreads();
    if (all_term || all_trunc) {
      episodes++;
      active_steps += s.steps;  // env.agent_steps[agent] of the finished episode
      successes += (p.rm_final >= 0 && (int)s.rm == p.rm_final) ? 1u : 0u;
      last_return = __double2float_rn(ep_ret);
      return_sum_add = __dadd_rn(return_sum_add, ep_ret);
      last_length = s.time;
      had_episode = true;
      ep_ret = 0.0;
      reset_slot(p, tb, i, a, t + 1, s, eps);
      explore_thr = explore_threshold(eps);
      // reset_e_table (ma_office.py:101-102)
      float4* E4 = reinterpret_cast<float4*>(E);
      for (long long j = lane; j < p.S4 / 4; j += 32) E4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
    }
  }
  if (lane == 0) {
    st.slot[k] = pack_slot(s);
    st.epsilon[k] = eps;
    if (st.ep_return) st.ep_return[k] = ep_ret;
    if (st.stats) {
      z.active_steps += active_steps;
      z.episodes += episodes;
      z.successes += successes;
      z.return_sum = return_sum_add;
      if (had_episode) {
        z.last_return = last_return;
        z.last_length = last_length;
      }
      st.stats[k] = z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Q(lambda), sparse-exact traces: one block per instance, one warp per agent. Only entries with a live trace are
// touched; each agent's live entries sit in a list that carries the trace AND the current q value (a write-back
// cache over the table), so one update step reads/writes the list once, coalesced. Bit-identical to the dense sweep of
// QLearningLambda.update (qlearning_lambda.py:33-84) up to the sign of zero: unlisted entries have e == 0 and receive +0.
// ------------------------------------------------------------------------------------------------
struct TraceList {
  unsigned short* pos;  // [S*4]
  unsigned short* idx;  // [cap]
  float* e;
  float* q;
};

// current value of table entry j: the listed copy when the entry has a live trace, else the table
__device__ __forceinline__ float trace_lookup(const float* Q, const TraceList& L, unsigned j) {
  const unsigned pz = L.pos[j];
  return pz ? L.q[pz - 1] : Q[j];
}

// write the listed values back and forget the list (traces wiped: reset_e_table / e_table.fill(0))
__device__ __forceinline__ void trace_flush(float* Q, const TraceList& L, unsigned len, int lane) {
  for (unsigned j = lane; j < len; j += 32) {
    const unsigned id = L.idx[j];
    Q[id] = L.q[j];
    L.pos[id] = 0;
  }
}

template <int ENV>
__global__ void __launch_bounds__(256) train_qlambda_sparse_kernel(KP p, DState st, unsigned long long t0, int n_iters, int learn,
                                                                  unsigned* trace) {
  Tab tb = stage_tables(p);
  __shared__ int sh_term[RLRM_MAX_AGENTS], sh_trunc[RLRM_MAX_AGENTS];
  const long long i = blockIdx.x;
  const int a = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long k = i * p.A + a;
  Slot s = unpack_slot(st.slot[k]);
  double eps = st.epsilon[k];
  double ep_ret = st.ep_return ? st.ep_return[k] : 0.0;
  float* Q = st.q + table_base(p, i, a);
  TraceList L;
  L.pos = st.tr_pos + (size_t)k * (size_t)p.S4;
  L.idx = st.tr_idx + (size_t)k * (size_t)st.tr_cap;
  L.e = st.tr_e + (size_t)k * (size_t)st.tr_cap;
  L.q = st.tr_q + (size_t)k * (size_t)st.tr_cap;
  unsigned len = st.tr_len[k];
  unsigned long long work = 0, active_steps = 0;
  unsigned episodes = 0, successes = 0, last_length = 0;
  float last_return = 0.f;
  bool had_episode = false;
  rlrm_stats_t z;
  if (st.stats) z = st.stats[k];
  double return_sum = st.stats ? z.return_sum : 0.0;
  unsigned long long explore_thr = explore_threshold(eps);

  for (int it = 0; it < n_iters; it++) {
    const unsigned long long t = t0 + (unsigned long long)it;
    unsigned w[4];
    RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
    __syncwarp();
    // Q row of the current state: lanes 0..3 fetch one action value each
    const unsigned rbase = (s.cell * p.nQ + s.rm) * 4;
    const float mine = lane < 4 ? trace_lookup(Q, L, rbase + lane) : 0.f;
    float4 row;
    row.x = __shfl_sync(0xFFFFFFFFu, mine, 0);
    row.y = __shfl_sync(0xFFFFFFFFu, mine, 1);
    row.z = __shfl_sync(0xFFFFFFFFu, mine, 2);
    row.w = __shfl_sync(0xFFFFFFFFu, mine, 3);
    const int action = select_action(row, explore_thr, w, learn == 0, p.n_actions);
    const unsigned before = s.cell;
    const bool first = (s.flags & RLRM_FLAG_FIRST) != 0;
    Rec r;
    agent_step<ENV>(p, tb, s, action, w[3], true, r);  // all lanes compute the same scalars
    if (learn) {
      const unsigned obs = (p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN && first) ? r.cell : before;
      const bool term_arg = p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN ? (r.term || r.trunc) : r.term;
      const unsigned hot = (obs * p.nQ + r.prev_q) * 4 + action, nbase = (r.cell * p.nQ + r.q) * 4;
      // lanes 0..3: next-state row, lane 4: Q[s,a] — five independent lookups in flight
      const float got = lane < 4 ? trace_lookup(Q, L, nbase + lane) : (lane == 4 ? trace_lookup(Q, L, hot) : 0.f);
      const float n0 = __shfl_sync(0xFFFFFFFFu, got, 0), n1 = __shfl_sync(0xFFFFFFFFu, got, 1);
      const float n2 = __shfl_sync(0xFFFFFFFFu, got, 2), n3 = __shfl_sync(0xFFFFFFFFu, got, 3);
      const float qsa = __shfl_sync(0xFFFFFFFFu, got, 4);
      const double best = term_arg ? 0.0 : (double)fmaxf(fmaxf(n0, n1), fmaxf(n2, n3));
      const float td = __fsub_rn(__double2float_rn(__dadd_rn(r.reward, __dmul_rn(p.gamma, best))), qsa);
      const float c = __fmul_rn(p.lr_f, td);
      bool found = false;
      for (unsigned j = lane; j < len; j += 32) {  // one coalesced pass over the live entries
        float e = L.e[j], q = L.q[j];
        if (L.idx[j] == hot) {
          e = 1.0f;  // replacing trace
          found = true;
        }
        q = __fadd_rn(q, __fmul_rn(c, e));
        e = term_arg ? 0.0f : __fmul_rn(e, p.trace_decay_f);
        L.q[j] = q;
        L.e[j] = e;
      }
      work += len;
      if (!__any_sync(0xFFFFFFFFu, found)) {  // first visit since the last wipe: the table value is current
        if (lane == 0) {
          L.idx[len] = (unsigned short)hot;
          L.q[len] = __fadd_rn(Q[hot], __fmul_rn(c, 1.0f));
          L.e[len] = term_arg ? 0.0f : __fmul_rn(1.0f, p.trace_decay_f);
          L.pos[hot] = (unsigned short)(len + 1);
        }
        len++;
      }
      __syncwarp();
      if (term_arg) {  // e_table.fill(0): nothing is live any more
        trace_flush(Q, L, len, lane);
        len = 0;
        __syncwarp();
      }
    }
    ep_ret = __dadd_rn(ep_ret, r.reward);
    if (trace && lane == 0)
      trace[(size_t)it * (size_t)(st.N * p.A) + (size_t)k] = (unsigned)action | (r.executed << 3) | (r.cell << 6) | (r.q << 16) |
                                                           ((unsigned)r.term << 21) | ((unsigned)r.trunc << 22) |
                                                           ((unsigned)r.stepped << 23);
    if (lane == 0) {
      sh_term[a] = r.term;
      sh_trunc[a] = r.trunc;
    }
    __syncthreads();
    bool all_term = true, all_trunc = true;
    for (int b = 0; b < p.A; b++) {
      all_term = all_term && sh_term[b];
      all_trunc = all_trunc && sh_trunc[b];
    }
    __syncthreads();
    if (all_term || all_trunc) {
      episodes++;
      active_steps += s.steps;
      successes += (p.rm_final >= 0 && (int)s.rm == p.rm_final) ? 1u : 0u;
      last_return = __double2float_rn(ep_ret);
      return_sum = __dadd_rn(return_sum, ep_ret);
      last_length = s.time;
      had_episode = true;
      ep_ret = 0.0;
      reset_slot(p, tb, i, a, t + 1, s, eps);
      explore_thr = explore_threshold(eps);
      trace_flush(Q, L, len, lane);  // reset_e_table (ma_office.py:101-102)
      len = 0;
      __syncwarp();
    }
  }
  if (lane == 0) {
    st.slot[k] = pack_slot(s);
    st.epsilon[k] = eps;
    st.tr_len[k] = len;
    if (st.tr_work) st.tr_work[k] += work;
    if (st.ep_return) st.ep_return[k] = ep_ret;
    if (st.stats) {
      z.active_steps += active_steps;
      z.episodes += episodes;
      z.successes += successes;
      z.return_sum = return_sum;
      if (had_episode) {
        z.last_return = last_return;
        z.last_length = last_length;
      }
      st.stats[k] = z;
    }
  }
}

// sparse Q(lambda): listed values -> table (lists stay live); optionally scatter the traces into a dense buffer
__global__ void __launch_bounds__(256) qlambda_materialize_kernel(KP p, DState st, float* e_dense) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= st.N * p.A) return;
  float* Q = st.q + (size_t)warp * (size_t)p.S4;
  const unsigned short* idx = st.tr_idx + (size_t)warp * (size_t)st.tr_cap;
  const float* lq = st.tr_q + (size_t)warp * (size_t)st.tr_cap;
  const float* le = st.tr_e + (size_t)warp * (size_t)st.tr_cap;
  const unsigned len = st.tr_len[warp];
  for (unsigned j = lane; j < len; j += 32) {
    Q[idx[j]] = lq[j];
    if (e_dense) e_dense[(size_t)warp * (size_t)p.S4 + idx[j]] = le[j];
  }
}

// sparse Q(lambda): reset_e_table for the masked instances = flush + forget the lists
__global__ void __launch_bounds__(256) qlambda_sparse_reset_kernel(KP p, DState st, const unsigned char* mask) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= st.N * p.A) return;
  if (mask && !mask[warp / p.A]) return;
  TraceList L;
  L.pos = st.tr_pos + (size_t)warp * (size_t)p.S4;
  L.idx = st.tr_idx + (size_t)warp * (size_t)st.tr_cap;
  L.e = st.tr_e + (size_t)warp * (size_t)st.tr_cap;
  L.q = st.tr_q + (size_t)warp * (size_t)st.tr_cap;
  trace_flush(st.q + (size_t)warp * (size_t)p.S4, L, st.tr_len[warp], lane);
  __syncwarp();
  if (lane == 0) st.tr_len[warp] = 0;
}

// ------------------------------------------------------------------------------------------------
// greedy evaluation (test_policy_optima, evaluation_metrics.py:23-190): the driver loop with best=True and no update
// ------------------------------------------------------------------------------------------------
template <int ENV, bool PA>
__global__ void __launch_bounds__(TRAIN_BLOCK) eval_kernel(KP p_in, DState st, rlrm_eval_t* evs, unsigned long long t0, int n_iters,
                                                          int n_episodes, double gamma, double optimal_steps) {
  KP p = p_in;
  Tab tb = stage_tables(p_in);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = tid >> p.g_shift;
  const int a = (int)(tid & (p.G - 1));
  const bool valid = (i < st.N) && (a < p.A);
  const long long k = i * p.A + a;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned group_mask = (p.G == 32 ? 0xFFFFFFFFu : ((1u << p.G) - 1u)) << (lane & ~(unsigned)(p.G - 1));
  Slot s = {0, 0, 0, 0, 0};
  rlrm_eval_t e;
  memset(&e, 0, sizeof(e));
  const float* Q = st.q;
  if (valid) {
    s = unpack_slot(st.slot[k]);
    e = evs[k];
    if (PA) agent_view(p_in, p, tb, a);
    Q = st.q + table_base(p_in, i, a);
  }
  const unsigned w0[4] = {0, 0, 0, 0};
  for (int it = 0; it < n_iters; it++) {
    const unsigned long long t = t0 + (unsigned long long)it;
    bool term = true, trunc = true;
    const bool running = valid && (int)e.episodes < n_episodes;  // all agents of an instance finish episodes together
    if (running) {
      const float4 row = *reinterpret_cast<const float4*>(Q + (size_t)(s.cell * p.nQ + s.rm) * 4);
      const int action = select_action(row, 0ull, w0, true, p.n_actions);
      unsigned w3 = 0;
      if (p.stochastic) {
        unsigned w[4];
        RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
        w3 = w[3];
      }
      Rec r;
      agent_step<ENV>(p, tb, s, action, w3, true, r);
      if (!e.in_success) {
        e.disc_return = __dadd_rn(e.disc_return, __dmul_rn(e.cum_gamma, r.reward));
        if (r.term && p.rm_final >= 0 && (int)r.q == p.rm_final) {
          e.successes++;
          e.in_success = 1;
        }
      }
      e.cum_gamma = __dmul_rn(e.cum_gamma, gamma);
      term = r.term;
      trunc = r.trunc;
    }
    const unsigned bt = __ballot_sync(0xFFFFFFFFu, term), bc = __ballot_sync(0xFFFFFFFFu, trunc);
    const bool over = ((bt & group_mask) == group_mask) || ((bc & group_mask) == group_mask);
    if (running && over) {
      const unsigned long long len = s.time;
      e.episodes++;
      e.return_sum = __dadd_rn(e.return_sum, e.disc_return);
      e.return_sqsum = __dadd_rn(e.return_sqsum, __dmul_rn(e.disc_return, e.disc_return));
      if (e.in_success) {
        e.len_sum += len;
        e.len_sqsum += len * len;
      }
      if (len > 0) e.arps_sum = __dadd_rn(e.arps_sum, __ddiv_rn(__ddiv_rn(e.disc_return, (double)len), optimal_steps));
      e.cum_gamma = 1.0;
      e.disc_return = 0.0;
      e.in_success = 0;
      double eps_unused = 0.0;
      KP q = p;
      q.decay_on_reset = 0;  // evaluation runs on a copy of the env: the training epsilon is not touched
      reset_slot(q, tb, i, a, t + 1, s, eps_unused);
    }
  }
  if (valid) {
    st.slot[k] = pack_slot(s);
    evs[k] = e;
  }
}

// ------------------------------------------------------------------------------------------------
// shared learner, fast path: ONE lockstep iteration over all instances by persistent blocks. The per-agent shared
// tables (A*S*4 floats, 25.6 KB for config 5) and the proposal accumulators live in SHARED memory: Q reads are LDS,
// proposals are shared-memory atomics, and each block flushes its non-empty accumulators to the global ones once.
// Integer sums make the result independent of the block/thread order (include/rlrm_b200.h "Shared learner").
// Slot state is streamed from HBM (8 B in + 8 B out per slot); statistics are only touched when an episode ends.
// ------------------------------------------------------------------------------------------------
#define SHARED_BLOCK 1024

template <int ENV, int ALGO>
__global__ void __launch_bounds__(SHARED_BLOCK, 1) shared_propose_kernel(KP p, DState st, unsigned long long t, int learn,
                                                                        unsigned* trace) {
  Tab tb = stage_tables(p);
  const int n_ent = p.A * (int)p.S4;
  float* Qs = reinterpret_cast<float*>(smem_raw + p.blob_bytes);
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(Qs + n_ent);
  int* s_cnt = reinterpret_cast<int*>(s_sum + n_ent);
  float* s_last = reinterpret_cast<float*>(s_cnt + n_ent);
  for (int j = threadIdx.x; j < n_ent / 4; j += blockDim.x)
    reinterpret_cast<float4*>(Qs)[j] = __ldg(reinterpret_cast<const float4*>(st.q) + j);
  for (int j = threadIdx.x; j < n_ent; j += blockDim.x) {
    s_sum[j] = 0ull;
    s_cnt[j] = 0;
  }
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u;
  const unsigned group_mask = (p.G == 32 ? 0xFFFFFFFFu : ((1u << p.G) - 1u)) << (lane & ~(unsigned)(p.G - 1));
  const long long total = st.N << p.g_shift;
  for (long long b0 = (long long)blockIdx.x * blockDim.x; b0 < total; b0 += (long long)gridDim.x * blockDim.x) {
    const long long tid = b0 + threadIdx.x;
    const long long i = tid >> p.g_shift;
    const int a = (int)(tid & (p.G - 1));
    const bool valid = (i < st.N) && (a < p.A);
    const long long k = i * p.A + a;
    bool term = true, trunc = true;
    Slot s = {0, 0, 0, 0, 0};
    double eps = 0.0;
    Rec r;
    r.reward = 0.0;
    if (valid) {
      s = unpack_slot(st.slot[k]);
      eps = st.epsilon[k];
      unsigned w[4];
      RLRM_PHILOX((unsigned)t, (unsigned)(t >> 32), p.instance_offset + (unsigned)i, (unsigned)a, p, w);
      float* Q = Qs + (size_t)a * (size_t)p.S4;
      const float4 row = *reinterpret_cast<const float4*>(Q + (size_t)(s.cell * p.nQ + s.rm) * 4);
      const int action = select_action(row, explore_threshold(eps), w, learn == 0, p.n_actions);
      const unsigned before = s.cell;
      const bool first = (s.flags & RLRM_FLAG_FIRST) != 0;
      agent_step<ENV>(p, tb, s, action, w[3], true, r);
      if (learn) {
        const unsigned obs = (p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN && first) ? r.cell : before;
        const bool term_arg = p.driver == RLRM_DRIVER_FROZEN_LAKE_MAIN ? (r.term || r.trunc) : r.term;
        Acc acc = {reinterpret_cast<long long*>(s_sum) + (size_t)a * (size_t)p.S4, s_cnt + (size_t)a * (size_t)p.S4,
                   s_last + (size_t)a * (size_t)p.S4};
        agent_update<ALGO>(p, tb, Q, nullptr, obs, action, term_arg, r, acc);
      }
      term = r.term;
      trunc = r.trunc;
      if (r.reward != 0.0 && st.ep_return) st.ep_return[k] = __dadd_rn(st.ep_return[k], r.reward);
      if (trace)
        trace[k] = (unsigned)action | (r.executed << 3) | (r.cell << 6) | (r.q << 16) | ((unsigned)r.term << 21) |
                   ((unsigned)r.trunc << 22) | ((unsigned)r.stepped << 23);
    }
    const unsigned bt = __ballot_sync(0xFFFFFFFFu, term), bc = __ballot_sync(0xFFFFFFFFu, trunc);
    const bool over = ((bt & group_mask) == group_mask) || ((bc & group_mask) == group_mask);
    if (valid) {
      if (over) {
        const double ret = st.ep_return ? st.ep_return[k] : 0.0;
        if (st.stats) {
          rlrm_stats_t z = st.stats[k];
          z.episodes++;
          z.active_steps += s.steps;
          z.successes += (p.rm_final >= 0 && (int)s.rm == p.rm_final) ? 1u : 0u;
          z.last_return = __double2float_rn(ret);
          z.return_sum = __dadd_rn(z.return_sum, ret);
          z.last_length = s.time;
          st.stats[k] = z;
        }
        if (st.ep_return) st.ep_return[k] = 0.0;
        reset_slot(p, tb, i, a, t + 1, s, eps);
        st.epsilon[k] = eps;
      }
      st.slot[k] = pack_slot(s);
    }
  }
  __syncthreads();
  if (learn) {
    for (int j = threadIdx.x; j < n_ent; j += blockDim.x) {
      const int c = s_cnt[j];
      if (c) {
        atomicAdd(st.acc_cnt + j, c);
        atomicAdd(reinterpret_cast<unsigned long long*>(st.acc_sum) + j, s_sum[j]);
        st.acc_last[j] = s_last[j];  // only read back when the GLOBAL count is 1, i.e. exactly one block wrote it
      }
    }
  }
}

// shared learner: every touched entry becomes the mean of this iteration's proposals; accumulators are cleared
__global__ void __launch_bounds__(256) apply_shared_kernel(KP p, DState st) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (p.per_agent ? p.sum4 : (long long)p.A * p.S4)) return;
  const int c = st.acc_cnt[j];
  if (c == 0) return;
  if (c == 1) st.q[j] = st.acc_last[j];
  else st.q[j] = __double2float_rn(__dmul_rn(__ddiv_rn((double)st.acc_sum[j], (double)c), 9.5367431640625e-07));
  st.acc_cnt[j] = 0;
  st.acc_sum[j] = 0;
}

// ------------------------------------------------------------------------------------------------
// host side: handle + C ABI
// ------------------------------------------------------------------------------------------------
struct rlrm_handle {
  KP kp;
  rlrm_config_t cfg;
  int device;
  unsigned char* d_blob;
  int smem_bytes;
  long long launches;
  int qrm4_fast;  // train_qrm4_kernel is applicable (see its header comment)
  int shared_fast;       // shared_propose_kernel is applicable (tables + accumulators fit in shared memory)
  int shared_smem_bytes;
  int num_sms;
};

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return fail(RLRM_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
  } while (0)

static int align16(int x) { return (x + 15) & ~15; }

extern "C" int rlrm_abi_version(void) { return RLRM_ABI_VERSION; }
extern "C" const char* rlrm_last_error(void) { return g_err; }
extern "C" int rlrm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static void fill_learner(KP& kp, double lr, double gamma, double lambd) {
  kp.lr = lr;
  kp.gamma = gamma;
  kp.lr_f = (float)lr;
  kp.one_minus_lr_f = (float)(1.0 - lr);
  kp.gamma_f = (float)gamma;
  kp.trace_decay_f = (float)(gamma * lambd);
}

extern "C" int rlrm_create(const rlrm_config_t* cfg, const rlrm_tables_t* tb, int device, rlrm_handle_t** out) {
  if (!cfg || !tb || !out) return fail(RLRM_ERR_ARG, "null argument");
  if (cfg->abi_version != RLRM_ABI_VERSION) return fail(RLRM_ERR_ARG, "abi_version mismatch");
  const int ncell = cfg->width * cfg->height;
  if (ncell <= 0 || ncell > RLRM_MAX_CELLS) return fail(RLRM_ERR_ARG, "width*height out of range");
  if (cfg->n_agents < 1 || cfg->n_agents > RLRM_MAX_AGENTS) return fail(RLRM_ERR_ARG, "n_agents out of range");
  if (cfg->n_rm_states < 1 || cfg->n_rm_states > RLRM_MAX_RM_STATES) return fail(RLRM_ERR_ARG, "n_rm_states out of range");
  if (cfg->n_events < 0 || cfg->n_events > RLRM_MAX_EVENTS) return fail(RLRM_ERR_ARG, "n_events out of range");
  if (cfg->n_qrm_states < 0 || cfg->n_qrm_states > RLRM_MAX_RM_STATES) return fail(RLRM_ERR_ARG, "n_qrm_states out of range");
  if (cfg->slip_n < 1 || cfg->slip_n > 4) return fail(RLRM_ERR_ARG, "slip_n out of range");
  if (cfg->n_actions < 1 || cfg->n_actions > RLRM_N_ACTIONS) return fail(RLRM_ERR_ARG, "n_actions out of range");
  if (cfg->random_starts && (!tb || !tb->free_cells || cfg->n_free_cells < cfg->n_agents || cfg->n_free_cells > ncell))
    return fail(RLRM_ERR_ARG, "random_starts needs tables.free_cells with n_agents <= n_free_cells <= width*height");
  if (cfg->algo < 0 || cfg->algo > RLRM_ALGO_QLAMBDA) return fail(RLRM_ERR_ARG, "unknown algo");
  if (cfg->env_kind != RLRM_ENV_FROZEN_LAKE && cfg->env_kind != RLRM_ENV_OFFICE_WORLD) return fail(RLRM_ERR_ARG, "unknown env_kind");
  if (cfg->algo == RLRM_ALGO_QLAMBDA && cfg->learning_rate < 0)
    return fail(RLRM_ERR_UNSUPPORTED, "Q(lambda) with learning_rate=None is not supported");
  if (cfg->algo == RLRM_ALGO_QLAMBDA && cfg->shared_q) return fail(RLRM_ERR_UNSUPPORTED, "Q(lambda) with a shared table is not supported");
  if (!tb->next_cell || !tb->cell_flags || !tb->label || !tb->delta || !tb->rq || !tb->rcf || !tb->start_cell)
    return fail(RLRM_ERR_ARG, "null table");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(RLRM_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(RLRM_ERR_ARG, "device out of range");
  CUDA_TRY(cudaSetDevice(device));

  rlrm_handle* h = new (std::nothrow) rlrm_handle();
  if (!h) return fail(RLRM_ERR_ARG, "out of host memory");
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->device = device;
  KP& kp = h->kp;
  kp.env_kind = cfg->env_kind; kp.driver = cfg->driver; kp.algo = cfg->algo;
  kp.A = cfg->n_agents;
  kp.G = 1; kp.g_shift = 0;
  while (kp.G < kp.A) { kp.G <<= 1; kp.g_shift++; }
  kp.nQ = cfg->n_rm_states; kp.nEv = cfg->n_events; kp.rm_final = cfg->rm_final; kp.n_qrm = cfg->n_qrm_states;
  kp.max_steps = cfg->max_steps; kp.ncell = ncell;
  kp.stochastic = cfg->stochastic; kp.slip_n = cfg->slip_n;
  for (int j = 0; j < 3; j++) kp.slip_thr[j] = cfg->slip_thr[j];
  kp.slip_cnt = 0;
  for (int j = 0; j + 1 < cfg->slip_n && j < 3; j++) {
    if (cfg->slip_thr[j] > 0xFFFFFFFFull) break;  // cumulative thresholds are non-decreasing
    kp.slip_thr32[j] = (unsigned)cfg->slip_thr[j];
    kp.slip_cnt = j + 1;
  }
  for (int a = 0; a < 4; a++)
    for (int j = 0; j < 4; j++) kp.slip_outcome[a * 4 + j] = cfg->slip_outcome[a][j];
  kp.terminate_on_plants = cfg->terminate_on_plants; kp.terminate_hit_walls = cfg->terminate_hit_walls;
  kp.hole_penalty = cfg->hole_penalty; kp.wall_penalty = cfg->wall_penalty;
  kp.eps_end = cfg->epsilon_end; kp.eps_decay = cfg->epsilon_decay;
  fill_learner(kp, cfg->learning_rate, cfg->gamma, cfg->lambd);
  kp.decay_on_reset = cfg->decay_on_reset; kp.shared_q = cfg->shared_q;
  kp.use_rsh = (cfg->use_rsh && tb->phi) ? 1 : 0;
  kp.per_agent = cfg->per_agent_rm ? 1 : 0;
  kp.nd = kp.nQ * (kp.nEv + 1);
  kp.sum4 = 0;
  for (int a = 0; a < kp.A; a++) {
    kp.a_nQ[a] = kp.per_agent ? cfg->agent_n_rm_states[a] : kp.nQ;
    kp.a_final[a] = kp.per_agent ? cfg->agent_rm_final[a] : kp.rm_final;
    kp.a_nqrm[a] = kp.per_agent ? cfg->agent_n_qrm[a] : kp.n_qrm;
    kp.a_prefix4[a] = kp.sum4;
    kp.sum4 += (long long)ncell * kp.a_nQ[a] * 4;
    if (kp.a_nQ[a] < 1 || kp.a_nQ[a] > kp.nQ || kp.a_nqrm[a] < 0 || kp.a_nqrm[a] > kp.nQ)
      return fail(RLRM_ERR_ARG, "per-agent reward machine sizes out of range");
  }
  if (kp.per_agent && (cfg->algo == RLRM_ALGO_QLAMBDA || kp.use_rsh))
    return fail(RLRM_ERR_UNSUPPORTED, "per-agent reward machines are supported for QL / QRM without shaping");
  kp.random_starts = cfg->random_starts ? 1 : 0;
  kp.n_free = cfg->n_free_cells;
  kp.seed_lo = cfg->seed_lo; kp.seed_hi = cfg->seed_hi; kp.instance_offset = cfg->instance_offset;
  kp.n_actions = (unsigned)cfg->n_actions;
  for (int r = 0; r < 10; r++) {
    kp.rk[2 * r] = cfg->seed_lo + (unsigned)r * 0x9E3779B9u;
    kp.rk[2 * r + 1] = cfg->seed_hi + (unsigned)r * 0xBB67AE85u;
  }
  kp.S4 = (long long)ncell * kp.nQ * 4;

  // pack the tables into one 16-byte aligned blob
  const int sec = kp.per_agent ? kp.A : 1;  // machine-dependent tables have one section per agent
  const int nd = kp.nQ * (kp.nEv + 1) * sec;
  int off = 0;
  kp.off_phi = off; off = align16(off + 2 * RLRM_MAX_RM_STATES * 8);
  kp.off_rq = off; off = align16(off + nd * 8);
  kp.off_rcf = off; off = align16(off + nd * 8);
  kp.off_next = off; off = align16(off + ncell * 4 * 2);
  kp.off_start = off; off = align16(off + RLRM_MAX_AGENTS * 2);
  kp.off_flags = off; off = align16(off + ncell);
  kp.off_label = off; off = align16(off + ncell * sec);
  kp.off_delta = off; off = align16(off + nd);
  kp.off_qrm = off; off = align16(off + RLRM_MAX_RM_STATES * sec);
  kp.off_free = off; off = align16(off + (kp.random_starts ? kp.n_free : 0) * 2);
  kp.blob_bytes = off;
  unsigned char* host = new (std::nothrow) unsigned char[off];
  if (!host) { delete h; return fail(RLRM_ERR_ARG, "out of host memory"); }
  memset(host, 0, off);
  if (tb->phi) memcpy(host + kp.off_phi, tb->phi, (size_t)kp.nQ * 2 * 8);
  memcpy(host + kp.off_rq, tb->rq, (size_t)nd * 8);
  memcpy(host + kp.off_rcf, tb->rcf, (size_t)nd * 8);
  memcpy(host + kp.off_next, tb->next_cell, (size_t)ncell * 8);
  memcpy(host + kp.off_start, tb->start_cell, (size_t)kp.A * 2);
  memcpy(host + kp.off_flags, tb->cell_flags, ncell);
  memcpy(host + kp.off_label, tb->label, (size_t)ncell * sec);
  memcpy(host + kp.off_delta, tb->delta, nd);
  if (kp.n_qrm > 0 && tb->qrm_states) memcpy(host + kp.off_qrm, tb->qrm_states, kp.per_agent ? (size_t)kp.nQ * sec : (size_t)kp.n_qrm);
  if (kp.random_starts) memcpy(host + kp.off_free, tb->free_cells, (size_t)kp.n_free * 2);
  cudaError_t e = cudaMalloc(&h->d_blob, off);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_blob, host, off, cudaMemcpyHostToDevice);
  delete[] host;
  if (e != cudaSuccess) { delete h; return fail(RLRM_ERR_CUDA, "table upload: %s", cudaGetErrorString(e)); }
  kp.blob = h->d_blob;
  h->smem_bytes = off;
  h->qrm4_fast = (kp.algo == RLRM_ALGO_QRM && kp.nQ == 4 && kp.n_qrm == 3 && !kp.shared_q && !kp.use_rsh && !kp.random_starts && !kp.per_agent && cfg->learning_rate >= 0.0 &&
                  !(cfg->reserved & 1));
  for (int j = 0; j < kp.n_qrm; j++)
    if (!tb->qrm_states || tb->qrm_states[j] != j) h->qrm4_fast = 0;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  {
    const long long n_ent = (long long)kp.A * kp.S4;
    const long long need = (long long)off + n_ent * 20;  // Q 4 B + sum 8 B + count 4 B + last 4 B per entry
    h->shared_smem_bytes = (int)need;
    h->shared_fast = (kp.shared_q && !kp.per_agent && kp.algo != RLRM_ALGO_QLAMBDA && cfg->learning_rate >= 0.0 && need <= 200 * 1024 &&
                      !(cfg->reserved & 1));
    if (h->shared_fast) {
      cudaError_t e1 = cudaSuccess;
      if (kp.env_kind == RLRM_ENV_FROZEN_LAKE) {
        if (kp.algo == RLRM_ALGO_QRM) e1 = cudaFuncSetAttribute(shared_propose_kernel<RLRM_ENV_FROZEN_LAKE, RLRM_ALGO_QRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        else e1 = cudaFuncSetAttribute(shared_propose_kernel<RLRM_ENV_FROZEN_LAKE, RLRM_ALGO_QL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
      } else {
        if (kp.algo == RLRM_ALGO_QRM) e1 = cudaFuncSetAttribute(shared_propose_kernel<RLRM_ENV_OFFICE_WORLD, RLRM_ALGO_QRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        else e1 = cudaFuncSetAttribute(shared_propose_kernel<RLRM_ENV_OFFICE_WORLD, RLRM_ALGO_QL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
      }
      if (e1 != cudaSuccess) h->shared_fast = 0;
    }
  }
  *out = h;
  return RLRM_OK;
}

extern "C" int rlrm_destroy(rlrm_handle_t* h) {
  if (!h) return RLRM_OK;
  cudaSetDevice(h->device);
  if (h->d_blob) cudaFree(h->d_blob);
  delete h;
  return RLRM_OK;
}

extern "C" int rlrm_set_learner(rlrm_handle_t* h, double learning_rate, double gamma, double lambd) {
  if (!h) return fail(RLRM_ERR_ARG, "null handle");
  if (h->cfg.algo == RLRM_ALGO_QLAMBDA && learning_rate < 0) return fail(RLRM_ERR_UNSUPPORTED, "Q(lambda) needs a fixed learning rate");
  h->cfg.learning_rate = learning_rate; h->cfg.gamma = gamma; h->cfg.lambd = lambd;
  fill_learner(h->kp, learning_rate, gamma, lambd);
  if (learning_rate < 0.0) h->qrm4_fast = 0;
  return RLRM_OK;
}

extern "C" int64_t rlrm_launch_count(const rlrm_handle_t* h) { return h ? (int64_t)h->launches : 0; }

static DState dstate(const rlrm_state_t* st) {
  DState d;
  d.N = st->n_instances; d.slot = (unsigned long long*)st->slot; d.epsilon = st->epsilon; d.q = st->q; d.e = st->e;
  d.visits = st->visits; d.ep_return = st->ep_return; d.stats = st->stats;
  d.acc_sum = (long long*)st->acc_sum; d.acc_cnt = st->acc_cnt; d.acc_last = st->acc_last;
  d.tr_pos = st->tr_pos; d.tr_idx = st->tr_idx; d.tr_e = st->tr_e; d.tr_q = st->tr_q; d.tr_len = st->tr_len;
  d.tr_work = (unsigned long long*)st->tr_work; d.tr_cap = st->tr_cap;
  return d;
}
static DOut dout(const rlrm_step_out_t* o) {
  DOut d;
  memset(&d, 0, sizeof(d));
  if (o) {
    d.prev_cell = o->prev_cell; d.cell = o->cell; d.prev_q = o->prev_q; d.q = o->q; d.event = o->event; d.executed = o->executed;
    d.renv = o->renv; d.rq = o->rq; d.reward = o->reward; d.env_term = o->env_term; d.rm_term = o->rm_term; d.term = o->term;
    d.trunc = o->trunc;
  }
  return d;
}

static int check_state(const rlrm_handle_t* h, const rlrm_state_t* st, bool need_q) {
  if (!h || !st) return fail(RLRM_ERR_ARG, "null handle/state");
  if (st->n_instances <= 0) return fail(RLRM_ERR_ARG, "n_instances must be positive");
  if (!st->slot || !st->epsilon) return fail(RLRM_ERR_ARG, "state.slot / state.epsilon are required");
  if (need_q && !st->q) return fail(RLRM_ERR_ARG, "state.q is required");
  if (need_q && h->cfg.algo == RLRM_ALGO_QLAMBDA && !st->e) {
    if (!st->tr_pos || !st->tr_idx || !st->tr_e || !st->tr_q || !st->tr_len)
      return fail(RLRM_ERR_ARG, "Q(lambda) needs state.e (dense traces) or state.tr_* (sparse traces)");
    if (st->tr_cap < h->cfg.max_steps + 1) return fail(RLRM_ERR_ARG, "tr_cap must be >= max_steps + 1");
    if (h->kp.S4 > 65535) return fail(RLRM_ERR_UNSUPPORTED, "sparse traces need S*4 <= 65535");
  }
  if (need_q && h->cfg.learning_rate < 0 && !st->visits) return fail(RLRM_ERR_ARG, "learning_rate=None needs state.visits");
  if (need_q && h->cfg.shared_q && h->cfg.learning_rate < 0) return fail(RLRM_ERR_UNSUPPORTED, "shared table with learning_rate=None");
  if (need_q && h->cfg.shared_q && (!st->acc_sum || !st->acc_cnt || !st->acc_last))
    return fail(RLRM_ERR_ARG, "shared_q needs state.acc_sum / acc_cnt / acc_last");
  return RLRM_OK;
}

#define LAUNCH_CHECK(h)                                                                     \
  do {                                                                                      \
    (h)->launches++;                                                                        \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) return fail(RLRM_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
  } while (0)

static unsigned blocks_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

extern "C" int rlrm_reset_at(rlrm_handle_t* h, const rlrm_state_t* st, const uint8_t* mask, uint64_t t, void* stream);
extern "C" int rlrm_reset(rlrm_handle_t* h, const rlrm_state_t* st, const uint8_t* mask, void* stream) {
  return rlrm_reset_at(h, st, mask, 0, stream);
}

extern "C" int rlrm_reset_at(rlrm_handle_t* h, const rlrm_state_t* st, const uint8_t* mask, uint64_t t, void* stream) {
  int rc = check_state(h, st, false);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = st->n_instances * h->kp.A;
  reset_kernel<<<blocks_for(n, 256), 256, h->smem_bytes, s>>>(h->kp, dstate(st), mask, t);
  LAUNCH_CHECK(h);
  if (h->cfg.algo == RLRM_ALGO_QLAMBDA && st->e) {
    clear_traces_kernel<<<blocks_for(n * (h->kp.S4 / 4), 256), 256, 0, s>>>(h->kp, dstate(st), mask);
    LAUNCH_CHECK(h);
  } else if (h->cfg.algo == RLRM_ALGO_QLAMBDA && st->tr_pos && st->q) {
    qlambda_sparse_reset_kernel<<<blocks_for(n * 32, 256), 256, 0, s>>>(h->kp, dstate(st), mask);
    LAUNCH_CHECK(h);
  }
  return RLRM_OK;
}

extern "C" int rlrm_select_action(rlrm_handle_t* h, const rlrm_state_t* st, const uint32_t* draws, uint64_t t, int best,
                                  uint8_t* actions_out, void* stream) {
  int rc = check_state(h, st, true);
  if (rc) return rc;
  if (!actions_out) return fail(RLRM_ERR_ARG, "actions_out is null");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = st->n_instances * h->kp.A;
  if (h->kp.per_agent) select_kernel<true><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(h->kp, dstate(st), draws, t, best, actions_out);
  else select_kernel<false><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(h->kp, dstate(st), draws, t, best, actions_out);
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_step(rlrm_handle_t* h, const rlrm_state_t* st, const uint8_t* actions, const uint32_t* draws, uint64_t t,
                         int with_rm, const rlrm_step_out_t* out, void* stream) {
  int rc = check_state(h, st, false);
  if (rc) return rc;
  if (!actions) return fail(RLRM_ERR_ARG, "actions is null");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = st->n_instances * h->kp.A;
  cudaStream_t s = (cudaStream_t)stream;
#define RLRM_STEP(ENV, PA) step_kernel<ENV, PA><<<blocks_for(n, 256), 256, h->smem_bytes, s>>>(h->kp, dstate(st), actions, draws, t, with_rm, dout(out))
  if (h->kp.env_kind == RLRM_ENV_FROZEN_LAKE) {
    if (h->kp.per_agent) RLRM_STEP(RLRM_ENV_FROZEN_LAKE, true); else RLRM_STEP(RLRM_ENV_FROZEN_LAKE, false);
  } else {
    if (h->kp.per_agent) RLRM_STEP(RLRM_ENV_OFFICE_WORLD, true); else RLRM_STEP(RLRM_ENV_OFFICE_WORLD, false);
  }
#undef RLRM_STEP
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_rm_step_agent(rlrm_handle_t* h, int agent, int64_t n_slots, uint8_t* q, const uint16_t* cell, uint8_t* event_out,
                                  double* reward_out, void* stream);
extern "C" int rlrm_rm_step(rlrm_handle_t* h, int64_t n_slots, uint8_t* q, const uint16_t* cell, uint8_t* event_out,
                            double* reward_out, void* stream) {
  return rlrm_rm_step_agent(h, 0, n_slots, q, cell, event_out, reward_out, stream);
}

extern "C" int rlrm_rm_step_agent(rlrm_handle_t* h, int agent, int64_t n_slots, uint8_t* q, const uint16_t* cell, uint8_t* event_out,
                                  double* reward_out, void* stream) {
  if (!h || !q || !cell) return fail(RLRM_ERR_ARG, "null argument");
  if (agent < 0 || agent >= h->kp.A) return fail(RLRM_ERR_ARG, "agent out of range");
  if (n_slots <= 0) return RLRM_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  rm_step_kernel<<<blocks_for(n_slots, 256), 256, h->smem_bytes, (cudaStream_t)stream>>>(h->kp, agent, n_slots, q, cell, event_out, reward_out);
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_update(rlrm_handle_t* h, const rlrm_state_t* st, const uint16_t* obs_cell, const uint8_t* actions,
                           const uint8_t* term_arg, const rlrm_step_out_t* out, void* stream) {
  int rc = check_state(h, st, true);
  if (rc) return rc;
  if (!obs_cell || !actions || !term_arg || !out) return fail(RLRM_ERR_ARG, "null argument");
  if (!out->prev_cell || !out->cell || !out->prev_q || !out->q || !out->event || !out->env_term || !out->renv || !out->reward)
    return fail(RLRM_ERR_ARG, "step record is missing fields the update needs");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = st->n_instances * h->kp.A;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->kp.algo == RLRM_ALGO_QLAMBDA && !st->e)
    return fail(RLRM_ERR_UNSUPPORTED, "rlrm_update on Q(lambda) needs dense traces (state.e); sparse traces are for rlrm_train");
  if (h->kp.algo == RLRM_ALGO_QLAMBDA)
    update_qlambda_kernel<<<(unsigned)n, 256, 0, s>>>(h->kp, dstate(st), obs_cell, actions, term_arg, dout(out));
#define RLRM_UPD(ALGO, PA) update_kernel<ALGO, PA><<<blocks_for(n, 256), 256, h->smem_bytes, s>>>(h->kp, dstate(st), obs_cell, actions, term_arg, dout(out))
  else if (h->kp.algo == RLRM_ALGO_QRM) {
    if (h->kp.per_agent) RLRM_UPD(RLRM_ALGO_QRM, true); else RLRM_UPD(RLRM_ALGO_QRM, false);
  } else {
    if (h->kp.per_agent) RLRM_UPD(RLRM_ALGO_QL, true); else RLRM_UPD(RLRM_ALGO_QL, false);
  }
#undef RLRM_UPD
  LAUNCH_CHECK(h);
  if (h->kp.shared_q) {
    apply_shared_kernel<<<blocks_for(h->kp.per_agent ? h->kp.sum4 : (long long)h->kp.A * h->kp.S4, 256), 256, 0, s>>>(h->kp, dstate(st));
    LAUNCH_CHECK(h);
  }
  return RLRM_OK;
}

template <int ENV>
static void launch_train(rlrm_handle_t* h, const rlrm_state_t* st, uint64_t t0, int n_iters, int learn, uint32_t* trace, cudaStream_t s) {
  const KP& kp = h->kp;
  if (kp.algo == RLRM_ALGO_QLAMBDA && !st->e) {
    train_qlambda_sparse_kernel<ENV><<<(unsigned)st->n_instances, kp.A * 32, h->smem_bytes, s>>>(kp, dstate(st), t0, n_iters, learn, trace);
  } else if (kp.algo == RLRM_ALGO_QLAMBDA) {
    train_qlambda_kernel<ENV><<<(unsigned)st->n_instances, kp.A * 32, h->smem_bytes, s>>>(kp, dstate(st), t0, n_iters, learn, trace);
  } else {
    const long long threads = st->n_instances * kp.G;
    const unsigned grid = blocks_for(threads, TRAIN_BLOCK);
#define RLRM_TRAIN(ALGO, PA) train_kernel<ENV, ALGO, PA><<<grid, TRAIN_BLOCK, h->smem_bytes, s>>>(kp, dstate(st), t0, n_iters, learn, trace)
    if (kp.algo == RLRM_ALGO_QRM && h->qrm4_fast && !st->visits) {
      const DState d = dstate(st);
#define RLRM_QRM4(ST, LE, TR) train_qrm4_kernel<ENV, ST, LE, TR><<<grid, TRAIN_BLOCK, h->smem_bytes, s>>>(kp, d, t0, n_iters, trace)
      const int key = (kp.stochastic ? 4 : 0) | (learn ? 2 : 0) | (trace ? 1 : 0);
      switch (key) {
        case 0: RLRM_QRM4(false, false, false); break;
        case 1: RLRM_QRM4(false, false, true); break;
        case 2: RLRM_QRM4(false, true, false); break;
        case 3: RLRM_QRM4(false, true, true); break;
        case 4: RLRM_QRM4(true, false, false); break;
        case 5: RLRM_QRM4(true, false, true); break;
        case 6: RLRM_QRM4(true, true, false); break;
        default: RLRM_QRM4(true, true, true); break;
      }
#undef RLRM_QRM4
    }
    else if (kp.algo == RLRM_ALGO_QRM) {
      if (kp.per_agent) RLRM_TRAIN(RLRM_ALGO_QRM, true); else RLRM_TRAIN(RLRM_ALGO_QRM, false);
    } else {
      if (kp.per_agent) RLRM_TRAIN(RLRM_ALGO_QL, true); else RLRM_TRAIN(RLRM_ALGO_QL, false);
    }
#undef RLRM_TRAIN
  }
}

extern "C" int rlrm_train(rlrm_handle_t* h, const rlrm_state_t* st, uint64_t t0, int32_t n_iters, int32_t learn, uint32_t* trace,
                          void* stream) {
  int rc = check_state(h, st, true);
  if (rc) return rc;
  if (n_iters < 0) return fail(RLRM_ERR_ARG, "n_iters < 0");
  if (n_iters == 0) return RLRM_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (h->kp.shared_q) {
    // synchronous iterations: one propose launch over all instances + one apply launch per lockstep iteration
    const size_t stride = (size_t)st->n_instances * h->kp.A;
    for (int it = 0; it < n_iters; it++) {
      uint32_t* tr = trace ? trace + (size_t)it * stride : nullptr;
      if (h->shared_fast && !st->visits) {
        const long long threads = st->n_instances * h->kp.G;
        long long want = (threads + SHARED_BLOCK - 1) / SHARED_BLOCK;
        const unsigned grid = (unsigned)(want < h->num_sms ? want : h->num_sms);
        const DState d = dstate(st);
        if (h->kp.env_kind == RLRM_ENV_FROZEN_LAKE) {
          if (h->kp.algo == RLRM_ALGO_QRM) shared_propose_kernel<RLRM_ENV_FROZEN_LAKE, RLRM_ALGO_QRM><<<grid, SHARED_BLOCK, h->shared_smem_bytes, s>>>(h->kp, d, t0 + it, learn, tr);
          else shared_propose_kernel<RLRM_ENV_FROZEN_LAKE, RLRM_ALGO_QL><<<grid, SHARED_BLOCK, h->shared_smem_bytes, s>>>(h->kp, d, t0 + it, learn, tr);
        } else {
          if (h->kp.algo == RLRM_ALGO_QRM) shared_propose_kernel<RLRM_ENV_OFFICE_WORLD, RLRM_ALGO_QRM><<<grid, SHARED_BLOCK, h->shared_smem_bytes, s>>>(h->kp, d, t0 + it, learn, tr);
          else shared_propose_kernel<RLRM_ENV_OFFICE_WORLD, RLRM_ALGO_QL><<<grid, SHARED_BLOCK, h->shared_smem_bytes, s>>>(h->kp, d, t0 + it, learn, tr);
        }
      } else if (h->kp.env_kind == RLRM_ENV_FROZEN_LAKE) launch_train<RLRM_ENV_FROZEN_LAKE>(h, st, t0 + it, 1, learn, tr, s);
      else launch_train<RLRM_ENV_OFFICE_WORLD>(h, st, t0 + it, 1, learn, tr, s);
      LAUNCH_CHECK(h);
      if (learn) {
        apply_shared_kernel<<<blocks_for(h->kp.per_agent ? h->kp.sum4 : (long long)h->kp.A * h->kp.S4, 256), 256, 0, s>>>(h->kp, dstate(st));
        LAUNCH_CHECK(h);
      }
    }
    return RLRM_OK;
  }
  if (h->kp.env_kind == RLRM_ENV_FROZEN_LAKE) launch_train<RLRM_ENV_FROZEN_LAKE>(h, st, t0, n_iters, learn, trace, s);
  else launch_train<RLRM_ENV_OFFICE_WORLD>(h, st, t0, n_iters, learn, trace, s);
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_evaluate(rlrm_handle_t* h, const rlrm_state_t* st, rlrm_eval_t* ev, uint64_t t0, int32_t n_iters,
                             int32_t n_episodes, double gamma, double optimal_steps, void* stream) {
  int rc = check_state(h, st, false);
  if (rc) return rc;
  if (!st->q || !ev) return fail(RLRM_ERR_ARG, "state.q and ev are required");
  if (n_iters <= 0 || n_episodes <= 0) return RLRM_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = blocks_for(st->n_instances * h->kp.G, TRAIN_BLOCK);
  cudaStream_t s = (cudaStream_t)stream;
#define RLRM_EVAL(ENV, PA) eval_kernel<ENV, PA><<<grid, TRAIN_BLOCK, h->smem_bytes, s>>>(h->kp, dstate(st), ev, t0, n_iters, n_episodes, gamma, optimal_steps)
  if (h->kp.env_kind == RLRM_ENV_FROZEN_LAKE) {
    if (h->kp.per_agent) RLRM_EVAL(RLRM_ENV_FROZEN_LAKE, true); else RLRM_EVAL(RLRM_ENV_FROZEN_LAKE, false);
  } else {
    if (h->kp.per_agent) RLRM_EVAL(RLRM_ENV_OFFICE_WORLD, true); else RLRM_EVAL(RLRM_ENV_OFFICE_WORLD, false);
  }
#undef RLRM_EVAL
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_qlambda_materialize(rlrm_handle_t* h, const rlrm_state_t* st, float* e_dense, void* stream) {
  if (!h || !st) return fail(RLRM_ERR_ARG, "null handle/state");
  if (!st->q || !st->tr_idx || !st->tr_q || !st->tr_e || !st->tr_len) return fail(RLRM_ERR_ARG, "no sparse trace state");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = st->n_instances * h->kp.A;
  qlambda_materialize_kernel<<<blocks_for(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(h->kp, dstate(st), e_dense);
  LAUNCH_CHECK(h);
  return RLRM_OK;
}

extern "C" int rlrm_train_host(rlrm_handle_t* h, const rlrm_state_t* st, uint64_t t0, int32_t n_iters, int32_t learn,
                               uint64_t* host_slot, double* host_epsilon, rlrm_stats_t* host_stats, void* stream) {
  int rc = check_state(h, st, true);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)st->n_instances * h->kp.A;
  if (host_slot) CUDA_TRY(cudaMemcpyAsync(st->slot, host_slot, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  if (host_epsilon) CUDA_TRY(cudaMemcpyAsync(st->epsilon, host_epsilon, n * sizeof(double), cudaMemcpyHostToDevice, s));
  rc = rlrm_train(h, st, t0, n_iters, learn, nullptr, stream);
  if (rc) return rc;
  if (host_stats) {
    if (!st->stats) return fail(RLRM_ERR_ARG, "host_stats requested but state.stats is null");
    CUDA_TRY(cudaMemcpyAsync(host_stats, st->stats, n * sizeof(rlrm_stats_t), cudaMemcpyDeviceToHost, s));
  }
  if (host_slot) CUDA_TRY(cudaMemcpyAsync(host_slot, st->slot, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  if (host_epsilon) CUDA_TRY(cudaMemcpyAsync(host_epsilon, st->epsilon, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return RLRM_OK;
}
