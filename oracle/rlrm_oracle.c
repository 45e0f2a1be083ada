/*
 * TEST INFRASTRUCTURE — CPU oracle for the lockstep hot path. NOT part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference algorithm (Alee08/multiagent-rl-rm v0.3.0, R/ = multiagent_rlrm/):
 *   env.step      R/environments/frozen_lake/ma_frozen_lake.py:96-154, 189-215, 224-262
 *                 R/environments/office_world/ma_office.py:122-202, 204-257, 269-325, 368-379
 *   env.reset     ma_frozen_lake.py:43-94 ; ma_office.py:77-120
 *   RM step       R/multi_agent/reward_machine.py:45-59
 *   wrapper       R/multi_agent/wrappers/rm_environment_wrapper.py:43-107, 122-183
 *   select        R/multi_agent/agent_rl.py:80-106 ; R/learning_algorithms/qlearning.py:112-143
 *   update        agent_rl.py:117-192 ; qlearning.py:41-110 ; qlearning_lambda.py:33-84
 *   driver loops  R/environments/frozen_lake/frozen_lake_main.py:336-376 ; office_world/office_main.py:1696-1749
 *
 * PARITY PINNED: validated against (a) the reference's own known-answer unit tests and (b) traces of
 * the live reference driven with injected Philox draws (oracle/ref_harness.py -> tests/golden/ *.npz).
 *
 * Table arithmetic type: -DORACLE_REAL=float restates what numpy computes when q_table/e_table are float32
 * (NEP-50 weak Python scalars, no FMA); -DORACLE_REAL=double restates the reference's native float64 tables.
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rlrm_b200.h"

#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
typedef ORACLE_REAL real;

#define SLOT_GET(w, sh, bits) ((uint32_t)(((w) >> (sh)) & ((1ull << (bits)) - 1)))

typedef struct {
  uint32_t cell, steps, time, rm, flags;
} slot_t;

static slot_t unpack(uint64_t w) {
  slot_t s;
  s.cell = SLOT_GET(w, RLRM_SLOT_CELL_SHIFT, 16);
  s.steps = SLOT_GET(w, RLRM_SLOT_STEPS_SHIFT, 16);
  s.time = SLOT_GET(w, RLRM_SLOT_TIME_SHIFT, 16);
  s.rm = SLOT_GET(w, RLRM_SLOT_RMSTATE_SHIFT, 8);
  s.flags = SLOT_GET(w, RLRM_SLOT_FLAGS_SHIFT, 8);
  return s;
}
static uint64_t pack(slot_t s) {
  return ((uint64_t)s.cell << RLRM_SLOT_CELL_SHIFT) | ((uint64_t)s.steps << RLRM_SLOT_STEPS_SHIFT) |
         ((uint64_t)s.time << RLRM_SLOT_TIME_SHIFT) | ((uint64_t)s.rm << RLRM_SLOT_RMSTATE_SHIFT) |
         ((uint64_t)s.flags << RLRM_SLOT_FLAGS_SHIFT);
}

int oracle_real_size(void) { return (int)sizeof(real); }

/* ---- Philox4x32-10 (Random123) ------------------------------------------------------------------ */
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void get_draws(const rlrm_config_t* cfg, const uint32_t* draws, uint64_t t, int64_t i, int a, uint32_t w[4]) {
  if (draws) {
    memcpy(w, draws + ((size_t)i * cfg->n_agents + a) * 4, 16);
  } else {
    oracle_philox((uint32_t)t, (uint32_t)(t >> 32), cfg->instance_offset + (uint32_t)i, (uint32_t)a, cfg->seed_lo, cfg->seed_hi, w);
  }
}

/* ---- per-agent reward machines (rlrm_config_t.per_agent_rm): agent a's scalars and table sections ---------- */
typedef struct {
  int nQ, final, n_qrm;
  const uint8_t *label, *delta, *qrm_states;
  const double *rq, *rcf, *phi;
} arm_t;

static arm_t agent_rm(const rlrm_config_t* cfg, const rlrm_tables_t* tb, int a) {
  arm_t v;
  int nd = cfg->n_rm_states * (cfg->n_events + 1), nc = cfg->width * cfg->height;
  if (cfg->per_agent_rm) {
    v.nQ = cfg->agent_n_rm_states[a]; v.final = cfg->agent_rm_final[a]; v.n_qrm = cfg->agent_n_qrm[a];
    v.label = tb->label + (size_t)a * nc; v.delta = tb->delta + (size_t)a * nd; v.rq = tb->rq + (size_t)a * nd;
    v.rcf = tb->rcf + (size_t)a * nd; v.qrm_states = tb->qrm_states ? tb->qrm_states + (size_t)a * cfg->n_rm_states : NULL;
    v.phi = tb->phi ? tb->phi + (size_t)a * 2 * cfg->n_rm_states : NULL;
  } else {
    v.nQ = cfg->n_rm_states; v.final = cfg->rm_final; v.n_qrm = cfg->n_qrm_states;
    v.label = tb->label; v.delta = tb->delta; v.rq = tb->rq; v.rcf = tb->rcf; v.qrm_states = tb->qrm_states; v.phi = tb->phi;
  }
  return v;
}

static size_t table_rows(const rlrm_config_t* cfg, int a) { /* S_a = W*H*nQ_a */
  return (size_t)cfg->width * cfg->height * (cfg->per_agent_rm ? cfg->agent_n_rm_states[a] : cfg->n_rm_states);
}

static size_t table_base(const rlrm_config_t* cfg, int64_t i, int a) {
  if (cfg->per_agent_rm) {
    size_t sum = 0, pre = 0;
    for (int b = 0; b < cfg->n_agents; b++) { if (b < a) pre += table_rows(cfg, b); sum += table_rows(cfg, b); }
    return (cfg->shared_q ? pre : (size_t)i * sum + pre) * 4;
  }
  size_t S = (size_t)cfg->width * cfg->height * cfg->n_rm_states;
  return (cfg->shared_q ? (size_t)a : (size_t)i * cfg->n_agents + a) * S * 4;
}

/* ---- reset: ma_frozen_lake.py:43-94 ; ma_office.py:77-120 ; agent_rl.py:372-384 ------------------- */
/* _sample_start_positions (ma_frozen_lake.py:156-172): distinct free cells; the reference shuffles the whole free-cell list
 * and keeps the first A entries, here the first A steps of a Fisher-Yates shuffle with injected Philox words (see
 * rlrm_config_t.random_starts) */
static void sample_starts(const rlrm_config_t* cfg, const rlrm_tables_t* tb, int64_t i, uint64_t T, uint32_t* out) {
  int F = cfg->n_free_cells, A = cfg->n_agents;
  uint32_t pos[RLRM_MAX_AGENTS], val[RLRM_MAX_AGENTS];
  int n_over = 0;
  uint32_t w[4] = {0, 0, 0, 0};
  for (int k = 0; k < A; k++) {
    if ((k & 3) == 0)
      oracle_philox((uint32_t)T, ~(uint32_t)(T >> 32), cfg->instance_offset + (uint32_t)i, 0x80000000u | (uint32_t)(k >> 2),
                    cfg->seed_lo, cfg->seed_hi, w);
    uint32_t j = (uint32_t)k + (uint32_t)(((uint64_t)w[k & 3] * (uint32_t)(F - k)) >> 32);
    uint32_t vk = tb->free_cells[k], vj = tb->free_cells[j];
    for (int m = 0; m < n_over; m++) { if (pos[m] == (uint32_t)k) vk = val[m]; if (pos[m] == j) vj = val[m]; }
    out[k] = vj;          /* list[k] <- list[j] */
    pos[n_over] = j; val[n_over] = vk; n_over++; /* list[j] <- list[k] (later entries override earlier ones) */
  }
}

static void reset_instance_at(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, int64_t i, uint64_t T) {
  uint32_t starts[RLRM_MAX_AGENTS];
  if (cfg->random_starts) sample_starts(cfg, tb, i, T, starts);
  for (int a = 0; a < cfg->n_agents; a++) {
    size_t k = (size_t)i * cfg->n_agents + a;
    slot_t s = {cfg->random_starts ? starts[a] : tb->start_cell[a], 0, 0, 0 /* RM initial state has index 0: reward_machine.py:32-36 */,
                RLRM_FLAG_ACTIVE | RLRM_FLAG_FIRST};
    st->slot[k] = pack(s);
    if (cfg->algo == RLRM_ALGO_QLAMBDA && st->e && !cfg->shared_q) /* reset_e_table: ma_frozen_lake.py:80-81 */
      memset((real*)st->e + table_base(cfg, i, a), 0, table_rows(cfg, a) * 4 * sizeof(real));
    if (cfg->decay_on_reset) { /* learn_done_episode: qlearning.py:153-155 */
      double e = st->epsilon[k] * cfg->epsilon_decay;
      st->epsilon[k] = cfg->epsilon_end > e ? cfg->epsilon_end : e;
    }
    if (st->ep_return) st->ep_return[k] = 0.0;
  }
}

int oracle_reset_at(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, const uint8_t* mask, uint64_t t) {
  for (int64_t i = 0; i < st->n_instances; i++)
    if (!mask || mask[i]) reset_instance_at(cfg, tb, st, i, t);
  return 0;
}

int oracle_reset(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, const uint8_t* mask) {
  return oracle_reset_at(cfg, tb, st, mask, 0);
}

/* ---- select: agent_rl.py:80-106 ; qlearning.py:112-143 -------------------------------------------- */
static int select_one(const rlrm_config_t* cfg, const real* row, double eps, const uint32_t w[4], int best) {
  int va = 0; /* np.argmax: first maximum */
  for (int j = 1; j < 4; j++) if (row[j] > row[va]) va = j;
  if (best) return va;
  double u = (double)w[0] / 4294967296.0; /* rng.uniform(0,1) < epsilon */
  if (u < eps) return (int)(((uint64_t)w[1] * (uint32_t)cfg->n_actions) >> 32); /* rng.choice(range(A)) */
  int maxs[4], n = 0;
  for (int j = 0; j < 4; j++) if (row[j] == row[va]) maxs[n++] = j;
  return maxs[((uint64_t)w[2] * (uint32_t)n) >> 32]; /* rng.choice(maxs) */
}

int oracle_select_action(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, const uint32_t* draws,
                         uint64_t t, int best, uint8_t* actions_out) {
  for (int64_t i = 0; i < st->n_instances; i++)
    for (int a = 0; a < cfg->n_agents; a++) {
      size_t k = (size_t)i * cfg->n_agents + a;
      slot_t s = unpack(st->slot[k]);
      const real* row = (const real*)st->q + table_base(cfg, i, a) + ((size_t)s.cell * agent_rm(cfg, tb, a).nQ + s.rm) * 4;
      uint32_t w[4];
      get_draws(cfg, draws, t, i, a, w);
      actions_out[k] = (uint8_t)select_one(cfg, row, st->epsilon[k], w, best);
    }
  return 0;
}

/* ---- slip: ma_frozen_lake.py:244-262 ; ma_office.py:368-379 --------------------------------------- */
static int slip(const rlrm_config_t* cfg, int intended, uint32_t k) {
  int idx = 0;
  for (int j = 0; j + 1 < cfg->slip_n; j++) idx += ((uint64_t)k >= cfg->slip_thr[j]);
  return cfg->slip_outcome[intended][idx];
}

typedef struct {
  uint32_t prev_cell, cell, prev_q, q, event, executed, env_term, rm_term, term, trunc, stepped;
  double renv, rq, reward;
} rec_t;

/* ---- one wrapper step of one instance ------------------------------------------------------------- */
static void step_instance(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, int64_t i,
                          const uint8_t* act_i /* [A] actions of this instance */, const uint32_t* draws, uint64_t t,
                          int with_rm, rec_t* rec) {
  int A = cfg->n_agents, nEv = cfg->n_events;
  slot_t s[RLRM_MAX_AGENTS];
  for (int a = 0; a < A; a++) s[a] = unpack(st->slot[(size_t)i * A + a]);
  uint32_t time = s[0].time;

  for (int a = 0; a < A; a++) { /* agents in registration order */
    rec_t* r = &rec[a];
    memset(r, 0, sizeof(*r));
    r->prev_cell = s[a].cell;
    r->executed = 5;
    int active = (s[a].flags & RLRM_FLAG_ACTIVE) != 0;
    uint32_t w[4];
    if (cfg->env_kind == RLRM_ENV_FROZEN_LAKE) {
      int fin = agent_rm(cfg, tb, a).final;
      int rm_done = fin >= 0 && (int)s[a].rm == fin; /* ma_frozen_lake.py:107-115 */
      if (active && !rm_done) {
        int act = act_i[a], ex = act;
        if (cfg->stochastic && act != RLRM_ACTION_WAIT) { get_draws(cfg, draws, t, i, a, w); ex = slip(cfg, act, w[3]); }
        if (ex != RLRM_ACTION_WAIT) s[a].cell = tb->next_cell[s[a].cell * 4 + ex]; /* :121-126, 224-242 */
        if (tb->cell_flags[s[a].cell] & 1) { s[a].flags |= RLRM_FLAG_FAIL; r->renv = cfg->hole_penalty; } /* :174-187 */
        s[a].steps++;
        r->executed = ex; r->stepped = 1;
      }
    } else {
      if (active) { /* ma_office.py:143-186 (no RM-final test: a finished agent keeps moving) */
        int act = act_i[a], ex = act;
        double wall_pen = 0.0;
        /* is_wall_collision -> "wait", no slip draw (:311-325); a "wait" action itself never collides (:299-300) */
        if (act != RLRM_ACTION_WAIT && tb->next_cell[s[a].cell * 4 + act] == s[a].cell) {
          if (cfg->terminate_hit_walls) s[a].flags |= RLRM_FLAG_FAIL;
          wall_pen = cfg->wall_penalty; ex = RLRM_ACTION_WAIT;
        }
        if (cfg->stochastic && ex != RLRM_ACTION_WAIT) { get_draws(cfg, draws, t, i, a, w); ex = slip(cfg, ex, w[3]); }
        if (ex != RLRM_ACTION_WAIT) s[a].cell = tb->next_cell[s[a].cell * 4 + ex]; /* apply_action re-checks (:269-289) */
        double plant = 0.0;
        if (tb->cell_flags[s[a].cell] & 1) { /* plants_in_the_office (:204-220) */
          if (cfg->terminate_on_plants) s[a].flags |= RLRM_FLAG_FAIL;
          plant = cfg->hole_penalty;
        }
        r->renv = wall_pen + plant; /* calculate_environment_rewards (:222-238) */
        s[a].steps++;
        r->executed = ex; r->stepped = 1;
      }
    }
    r->cell = s[a].cell;
  }
  time++;
  for (int a = 0; a < A; a++) { /* check_terminations: ma_frozen_lake.py:189-215 ; ma_office.py:240-257 */
    rec_t* r = &rec[a];
    int fail = (s[a].flags & RLRM_FLAG_FAIL) != 0;
    if (cfg->env_kind == RLRM_ENV_FROZEN_LAKE) {
      r->trunc = ((int)s[a].steps > cfg->max_steps) || ((int)time > cfg->max_steps);
      int fin = agent_rm(cfg, tb, a).final;
      r->env_term = r->trunc || (fin >= 0 && (int)s[a].rm == fin) || fail;
      if (r->env_term) s[a].flags &= ~RLRM_FLAG_ACTIVE;
    } else {
      r->trunc = (int)time > cfg->max_steps;
      r->env_term = fail;
      if (r->env_term || r->trunc) s[a].flags &= ~RLRM_FLAG_ACTIVE;
    }
    s[a].time = time;
  }
  for (int a = 0; a < A; a++) { /* rm_environment_wrapper.py:57-107 */
    rec_t* r = &rec[a];
    r->prev_q = s[a].rm;
    arm_t v = agent_rm(cfg, tb, a);
    r->event = v.label[s[a].cell];
    r->q = s[a].rm;
    if (with_rm) {
      int col = r->event == RLRM_EVENT_NONE ? nEv : (int)r->event;
      uint8_t d = v.delta[s[a].rm * (nEv + 1) + col]; /* reward_machine.py:45-59 */
      if (d != RLRM_NO_TRANSITION) { r->rq = v.rq[s[a].rm * (nEv + 1) + col]; s[a].rm = d; }
      r->q = s[a].rm;
      r->rm_term = v.final >= 0 && (int)s[a].rm == v.final;
    }
    r->reward = r->renv + r->rq;
    r->term = r->env_term || r->rm_term;
    s[a].flags &= ~(RLRM_FLAG_DONE | RLRM_FLAG_TRUNC);
    if (r->term) s[a].flags |= RLRM_FLAG_DONE;
    if (r->trunc) s[a].flags |= RLRM_FLAG_TRUNC;
    st->slot[(size_t)i * A + a] = pack(s[a]); /* RLRM_FLAG_FIRST is cleared by the caller after the update */
  }
}

static void store_rec(const rlrm_step_out_t* o, size_t k, const rec_t* r) {
  if (!o) return;
  if (o->prev_cell) o->prev_cell[k] = (uint16_t)r->prev_cell;
  if (o->cell) o->cell[k] = (uint16_t)r->cell;
  if (o->prev_q) o->prev_q[k] = (uint8_t)r->prev_q;
  if (o->q) o->q[k] = (uint8_t)r->q;
  if (o->event) o->event[k] = (uint8_t)r->event;
  if (o->executed) o->executed[k] = (uint8_t)r->executed;
  if (o->renv) o->renv[k] = r->renv;
  if (o->rq) o->rq[k] = r->rq;
  if (o->reward) o->reward[k] = r->reward;
  if (o->env_term) o->env_term[k] = (uint8_t)r->env_term;
  if (o->rm_term) o->rm_term[k] = (uint8_t)r->rm_term;
  if (o->term) o->term[k] = (uint8_t)r->term;
  if (o->trunc) o->trunc[k] = (uint8_t)r->trunc;
}

static void clear_first(const rlrm_config_t* cfg, const rlrm_state_t* st, int64_t i) {
  for (int a = 0; a < cfg->n_agents; a++)
    st->slot[(size_t)i * cfg->n_agents + a] &= ~((uint64_t)RLRM_FLAG_FIRST << RLRM_SLOT_FLAGS_SHIFT);
}

int oracle_step(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, const uint8_t* actions,
                const uint32_t* draws, uint64_t t, int with_rm, const rlrm_step_out_t* out) {
  rec_t rec[RLRM_MAX_AGENTS];
  for (int64_t i = 0; i < st->n_instances; i++) {
    step_instance(cfg, tb, st, i, actions + (size_t)i * cfg->n_agents, draws, t, with_rm, rec);
    clear_first(cfg, st, i);
    for (int a = 0; a < cfg->n_agents; a++) {
      size_t k = (size_t)i * cfg->n_agents + a;
      store_rec(out, k, &rec[a]);
      if (out && (out->cf_q || out->cf_r)) { /* _get_qrm_experiences' RM lookups on the new position (rm_environment_wrapper.py:144-153) */
        arm_t v = agent_rm(cfg, tb, a);
        int col = rec[a].event == RLRM_EVENT_NONE ? cfg->n_events : (int)rec[a].event;
        for (int j = 0; j < v.n_qrm; j++) {
          int u = v.qrm_states[j];
          int d = v.delta[u * (cfg->n_events + 1) + col];
          if (out->cf_q) out->cf_q[k * cfg->n_qrm_states + j] = (uint8_t)(d == RLRM_NO_TRANSITION ? u : d);
          if (out->cf_r) out->cf_r[k * cfg->n_qrm_states + j] = d == RLRM_NO_TRANSITION ? 0.0 : v.rcf[u * (cfg->n_events + 1) + col];
        }
      }
    }
  }
  return 0;
}

/* ---- RMEnvironmentWrapper.get_mdp (rm_environment_wrapper.py:185-283) ----------------------------------------------
 * Follows the reference's procedure literally: for every encoded state of `agent`, skip terminal states
 * (is_terminal_state_mdp: ma_frozen_lake.py:321-335, ma_office.py:411-432); otherwise for every nominal action and
 * every sub-action: reset the instance (wrapper.reset), overwrite the agent's position / RM state (env.set_state),
 * step once with env.stochastic = False, the other agents doing anything (their moves cannot touch this agent), and
 * record (next encoded state, reward, terminated or truncated). */
int oracle_mdp(const rlrm_config_t* cfg_in, const rlrm_tables_t* tb, int agent, int n_sub, const uint8_t* sub_actions,
               int rm_terminal, int32_t* next_state, double* reward, uint8_t* done, uint8_t* terminal) {
  rlrm_config_t cfg = *cfg_in;
  cfg.stochastic = 0;   /* rm_environment_wrapper.py:196-197 */
  cfg.random_starts = 0; /* set_state overrides this agent's start; the others' positions are irrelevant */
  int A = cfg.n_agents;
  arm_t v = agent_rm(&cfg, tb, agent);
  int nQ = v.nQ, ncell = cfg.width * cfg.height;
  uint64_t slot[RLRM_MAX_AGENTS];
  double eps[RLRM_MAX_AGENTS], ret[RLRM_MAX_AGENTS];
  rlrm_state_t st;
  memset(&st, 0, sizeof(st));
  st.n_instances = 1; st.slot = slot; st.epsilon = eps; st.ep_return = ret;
  rec_t rec[RLRM_MAX_AGENTS];
  uint8_t acts[RLRM_MAX_AGENTS];
  for (int enc = 0; enc < ncell * nQ; enc++) {
    int cell = enc / nQ, q = enc % nQ;
    int kind = 0; double tr = 0.0;
    if ((tb->cell_flags[cell] & 1) && (cfg.env_kind == RLRM_ENV_FROZEN_LAKE || cfg.terminate_on_plants)) { kind = 1; tr = cfg.hole_penalty; }
    else if (rm_terminal && v.final >= 0 && q == v.final) kind = 2;
    terminal[enc] = (uint8_t)kind;
    for (int a = 0; a < 4; a++)
      for (int j = 0; j < n_sub; j++) {
        size_t k = ((size_t)enc * 4 + a) * n_sub + j;
        if (kind) { next_state[k] = enc; reward[k] = tr; done[k] = 1; continue; }
        for (int b = 0; b < A; b++) { eps[b] = cfg.epsilon_start; slot[b] = 0; }
        reset_instance_at(&cfg, tb, &st, 0, 0);
        slot_t s = unpack(slot[agent]);
        s.cell = (uint32_t)cell; s.rm = (uint32_t)q;
        slot[agent] = pack(s);
        memset(acts, 0, sizeof(acts));
        acts[agent] = sub_actions[a * n_sub + j];
        step_instance(&cfg, tb, &st, 0, acts, NULL, 0, 1, rec);
        next_state[k] = (int32_t)(rec[agent].cell * nQ + rec[agent].q);
        reward[k] = rec[agent].reward;
        done[k] = (uint8_t)(rec[agent].term || rec[agent].trunc);
      }
  }
  return 0;
}

/* ---- mdp_vi.value_iteration (environments/utils_envs/mdp_vi.py:9-60), literally: V = 0; repeat { old_V = V.copy(); for s in
 * order: q[a] = sum over P[s][a] of prob * (reward + gamma * V[s_next] * (not done)) -- V is updated IN PLACE while sweeping;
 * delta = max |max_q - old_V[s]| (relative variant: / |max_q| when > 1e-12) } until delta < theta; policy = argmax(Q, axis=1).
 * P is given as padded arrays [S][4][n_out], prob 0 = unused slot. Returns the number of sweeps. */
int oracle_value_iteration(int64_t S, int n_out, const double* prob, const int32_t* next_state, const double* reward,
                           const uint8_t* done, double gamma, double theta, int delta_rel, double* V, double* Q, int32_t* policy,
                           double* old_v) {
  for (int64_t s = 0; s < S; s++) V[s] = 0.0;
  for (int64_t k = 0; k < S * 4; k++) Q[k] = 0.0;
  int sweeps = 0;
  for (;;) {
    double delta = 0.0;
    memcpy(old_v, V, (size_t)S * sizeof(double));
    for (int64_t s = 0; s < S; s++) {
      double max_q = 0.0;
      for (int a = 0; a < 4; a++) {
        double q = 0.0;
        size_t base = ((size_t)s * 4 + a) * (size_t)n_out;
        for (int j = 0; j < n_out; j++) {
          if (prob[base + j] == 0.0) continue;
          double fut = gamma * V[next_state[base + j]];
          fut = done[base + j] ? fut * 0.0 : fut * 1.0; /* gamma * V[s_next] * (not done) */
          q += prob[base + j] * (reward[base + j] + fut);
        }
        Q[s * 4 + a] = q;
        if (a == 0 || q > max_q) max_q = q;
      }
      double diff = fabs(max_q - old_v[s]);
      double cdelta = delta_rel ? diff / (fabs(max_q) > 1e-12 ? fabs(max_q) : 1.0) : diff;
      if (cdelta > delta) delta = cdelta;
      V[s] = max_q;
    }
    sweeps++;
    if (delta < theta) break;
  }
  for (int64_t s = 0; s < S; s++) {
    int arg = 0;
    for (int a = 1; a < 4; a++) if (Q[s * 4 + a] > Q[s * 4 + arg]) arg = a;
    policy[s] = arg;
  }
  return sweeps;
}

int oracle_rm_step(const rlrm_config_t* cfg, const rlrm_tables_t* tb, int64_t n_slots, uint8_t* q, const uint16_t* cell,
                   uint8_t* event_out, double* reward_out) {
  int nEv = cfg->n_events;
  for (int64_t k = 0; k < n_slots; k++) {
    uint8_t ev = tb->label[cell[k]];
    int col = ev == RLRM_EVENT_NONE ? nEv : ev;
    uint8_t d = tb->delta[q[k] * (nEv + 1) + col];
    double r = 0.0;
    if (d != RLRM_NO_TRANSITION) { r = tb->rq[q[k] * (nEv + 1) + col]; q[k] = d; }
    if (event_out) event_out[k] = ev;
    if (reward_out) reward_out[k] = r;
  }
  return 0;
}

/* ---- update_q: qlearning.py:70-79 ------------------------------------------------------------------ */
static real row_max(const real* row) { /* np.max */
  real m = row[0];
  for (int j = 1; j < 4; j++) if (row[j] > m) m = row[j];
  return m;
}

typedef struct { int64_t* sum; int32_t* cnt; float* last; } acc_t; /* shared learner: per-agent-table accumulators */

static void update_q(const rlrm_config_t* cfg, real* Q, uint32_t* visits, size_t s, int a, double r, size_t sn, int terminated,
                     const acc_t* acc) {
  real cur = Q[s * 4 + a];
  if (visits) visits[s * 4 + a] += 1;
  real mf = (real)(terminated ? 0 : 1) * row_max(Q + sn * 4); /* (not terminated) * np.max(q_table[sn]) */
  real inner = (real)r + (real)cfg->gamma * mf;               /* weak Python scalars adopt the table dtype */
  if (cfg->learning_rate < 0) { /* lr = 1 / visits is an np.float64 (strong): outer expression is evaluated in double */
    double lr = 1.0 / (double)visits[s * 4 + a];
    Q[s * 4 + a] = (real)((1.0 - lr) * (double)cur + lr * (double)inner);
  } else {
    double lr = cfg->learning_rate;
    real nv = (real)(1.0 - lr) * cur + (real)lr * inner;
    if (acc) { /* shared learner: propose instead of writing (include/rlrm_b200.h, "Shared learner") */
      acc->sum[s * 4 + a] += llrintf((float)nv * 1048576.0f);
      acc->cnt[s * 4 + a] += 1;
      acc->last[s * 4 + a] = (float)nv;
    } else {
      Q[s * 4 + a] = nv;
    }
  }
}

/* ---- QLearningLambda.update: qlearning_lambda.py:33-84 (dense sweep, as written) ------------------- */
static void update_qlambda(const rlrm_config_t* cfg, real* Q, real* E, uint32_t* visits, size_t S, size_t s, int a,
                           double reward, size_t sn, int terminated) {
  if (visits) visits[s * 4 + a] += 1;
  double best = terminated ? 0.0 : (double)row_max(Q + sn * 4);
  real td = (real)(reward + cfg->gamma * best) - Q[s * 4 + a]; /* double sum, rounded at the subtraction */
  E[s * 4 + a] = (real)1;
  if (cfg->learning_rate < 0) {
    /* lr = 1 / visits[s, a] is an np.float64 (qlearning_lambda.py:44-49): lr * td_error and (lr * td_error) * e_table are
     * float64, and `q_table +=` adds in float64 before casting back to the table's dtype */
    double c = (1.0 / (double)visits[s * 4 + a]) * (double)td;
    for (size_t j = 0; j < S * 4; j++) Q[j] = (real)((double)Q[j] + c * (double)E[j]);
  } else {
    real c = (real)cfg->learning_rate * td;
    for (size_t j = 0; j < S * 4; j++) Q[j] = Q[j] + c * E[j];
  }
  if (terminated) {
    memset(E, 0, S * 4 * sizeof(real));
  } else { /* next_action defaults to argmax Q[s'] => greedy => decay (:71-81) */
    real d = (real)(cfg->gamma * cfg->lambd);
    for (size_t j = 0; j < S * 4; j++) E[j] = E[j] * d;
  }
}

static void update_slot(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, int64_t i, int a,
                        uint32_t obs_cell, int action, int term_arg, const rec_t* r) {
  arm_t v = agent_rm(cfg, tb, a);
  int nQ = v.nQ, nEv = cfg->n_events, pstride = cfg->n_rm_states; /* phi rows are nQmax apart */
  size_t S = (size_t)cfg->width * cfg->height * nQ;
  size_t base = table_base(cfg, i, a);
  real* Q = (real*)st->q + base;
  uint32_t* V = st->visits ? st->visits + base : NULL;
  acc_t acc = {NULL, NULL, NULL};
  const acc_t* accp = NULL;
  if (cfg->shared_q && st->acc_sum) {
    acc.sum = st->acc_sum + base; acc.cnt = st->acc_cnt + base; acc.last = st->acc_last + base;
    accp = &acc;
  }
  if (cfg->algo == RLRM_ALGO_QRM) { /* rm_environment_wrapper.py:122-183 -> qlearning.py:82-106 */
    int col = r->event == RLRM_EVENT_NONE ? nEv : (int)r->event;
    for (int j = 0; j < v.n_qrm; j++) {
      int u = v.qrm_states[j];
      uint8_t d = v.delta[u * (nEv + 1) + col];
      int un = d == RLRM_NO_TRANSITION ? u : d;
      double ru = d == RLRM_NO_TRANSITION ? 0.0 : v.rcf[u * (nEv + 1) + col];
      int done = r->env_term || (v.final >= 0 && un == v.final);
      double rew = r->renv + ru;
      if (cfg->use_rsh && v.phi) rew += cfg->gamma * v.phi[un] - v.phi[u]; /* qlearning.py:93-105 */
      update_q(cfg, Q, V, (size_t)r->prev_cell * nQ + u, action, rew, (size_t)r->cell * nQ + un, done, accp);
    }
  } else {
    size_t s = (size_t)obs_cell * nQ + r->prev_q, sn = (size_t)r->cell * nQ + r->q; /* agent_rl.py:154-155 */
    if (cfg->algo == RLRM_ALGO_QL) {
      double rew = r->reward;
      if (cfg->use_rsh && v.phi) rew += cfg->gamma * v.phi[pstride + r->q] - v.phi[pstride + r->prev_q]; /* qlearning.py:51-66 */
      update_q(cfg, Q, V, s, action, rew, sn, term_arg, accp);
    }
    else update_qlambda(cfg, Q, (real*)st->e + base, V, S, s, action, r->reward, sn, term_arg);
  }
}

int oracle_update(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, const uint16_t* obs_cell,
                  const uint8_t* actions, const uint8_t* term_arg, const rlrm_step_out_t* o) {
  for (int64_t i = 0; i < st->n_instances; i++)
    for (int a = 0; a < cfg->n_agents; a++) {
      size_t k = (size_t)i * cfg->n_agents + a;
      rec_t r;
      memset(&r, 0, sizeof(r));
      r.prev_cell = o->prev_cell[k]; r.cell = o->cell[k]; r.prev_q = o->prev_q[k]; r.q = o->q[k];
      r.event = o->event[k]; r.env_term = o->env_term[k]; r.renv = o->renv[k]; r.reward = o->reward[k];
      update_slot(cfg, tb, st, i, a, obs_cell[k], actions[k], term_arg[k], &r);
    }
  return 0;
}

/* ---- driver loop: frozen_lake_main.py:345-376 ; office_main.py:1709-1749 --------------------------- */
static void train_iteration(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, int64_t i, uint64_t t,
                            int32_t it, int32_t learn, uint32_t* trace) {
  int A = cfg->n_agents;
  rec_t rec[RLRM_MAX_AGENTS];
  uint8_t act[RLRM_MAX_AGENTS];
  uint32_t before[RLRM_MAX_AGENTS];
  int first = 0;
  for (int a = 0; a < A; a++) { /* every agent selects, finished ones included (:350-352) */
    size_t k = (size_t)i * A + a;
    slot_t s = unpack(st->slot[k]);
    before[a] = s.cell;
    first = (s.flags & RLRM_FLAG_FIRST) != 0;
    const real* row = (const real*)st->q + table_base(cfg, i, a) + ((size_t)s.cell * agent_rm(cfg, tb, a).nQ + s.rm) * 4;
    uint32_t w[4];
    get_draws(cfg, NULL, t, i, a, w);
    act[a] = (uint8_t)select_one(cfg, row, st->epsilon[k], w, !learn); /* learn == 0: greedy evaluation, best=True */
  }
  step_instance(cfg, tb, st, i, act, NULL, t, 1, rec);
  int all_term = 1, all_trunc = 1;
  for (int a = 0; a < A; a++) {
    size_t k = (size_t)i * A + a;
    /* FrozenLake driver: on the first iteration `states` still aliases agent.state (frozen_lake_main.py:337 vs
     * office_main.py:1700 which deep-copies) so update_policy sees the NEW position as `state`. */
    uint32_t obs = (cfg->driver == RLRM_DRIVER_FROZEN_LAKE_MAIN && first) ? rec[a].cell : before[a];
    int term_arg = cfg->driver == RLRM_DRIVER_FROZEN_LAKE_MAIN ? (rec[a].term || rec[a].trunc) : rec[a].term;
    if (learn) update_slot(cfg, tb, st, i, a, obs, act[a], term_arg, &rec[a]);
    all_term &= (int)rec[a].term; all_trunc &= (int)rec[a].trunc;
    if (st->ep_return) st->ep_return[k] += rec[a].reward;
    if (trace)
      trace[(size_t)it * st->n_instances * A + k] = (uint32_t)act[a] | (rec[a].executed << 3) | (rec[a].cell << 6) |
          (rec[a].q << 16) | (rec[a].term << 21) | (rec[a].trunc << 22) | (rec[a].stepped << 23);
  }
  clear_first(cfg, st, i);
  if (all_term || all_trunc) { /* episode over -> next episode starts with rm_env.reset (:337 / :1699) */
    for (int a = 0; a < A; a++) {
      size_t k = (size_t)i * A + a;
      if (st->stats) {
        slot_t s = unpack(st->slot[k]);
        rlrm_stats_t* z = &st->stats[k];
        z->episodes++;
        z->active_steps += s.steps; /* env.agent_steps[agent] of the finished episode */
        { int fin = agent_rm(cfg, tb, a).final; z->successes += (fin >= 0 && (int)s.rm == fin); }
        double ret = st->ep_return ? st->ep_return[k] : 0.0;
        z->last_return = (float)ret;
        z->return_sum += ret;
        z->last_length = s.time;
      }
    }
    reset_instance_at(cfg, tb, st, i, t + 1);
  }
}

/* shared learner: every touched entry becomes the mean of this iteration's proposals (include/rlrm_b200.h) */
static void apply_shared(const rlrm_config_t* cfg, const rlrm_state_t* st) {
  size_t n = 0;
  for (int a = 0; a < cfg->n_agents; a++) n += table_rows(cfg, a) * 4;
  real* Q = (real*)st->q;
  for (size_t j = 0; j < n; j++) {
    int32_t c = st->acc_cnt[j];
    if (c == 1) Q[j] = (real)st->acc_last[j];
    else if (c > 1) Q[j] = (real)(float)((double)st->acc_sum[j] / (double)c * 9.5367431640625e-07);
    st->acc_cnt[j] = 0; st->acc_sum[j] = 0;
  }
}

/* ---- driver loop: frozen_lake_main.py:345-376 ; office_main.py:1709-1749 --------------------------- */
int oracle_train(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, uint64_t t0, int32_t n_iters,
                 int32_t learn, uint32_t* trace) {
  if (cfg->shared_q) { /* synchronous iterations over all instances */
    for (int32_t it = 0; it < n_iters; it++) {
      for (int64_t i = 0; i < st->n_instances; i++) train_iteration(cfg, tb, st, i, t0 + (uint64_t)it, it, learn, trace);
      if (learn && st->acc_sum) apply_shared(cfg, st);
    }
  } else { /* instances are independent */
    for (int64_t i = 0; i < st->n_instances; i++)
      for (int32_t it = 0; it < n_iters; it++) train_iteration(cfg, tb, st, i, t0 + (uint64_t)it, it, learn, trace);
  }
  return 0;
}

/* ---- greedy evaluation: evaluation_metrics.py:23-190 (test_policy_optima) --------------------------- */
int oracle_evaluate(const rlrm_config_t* cfg, const rlrm_tables_t* tb, const rlrm_state_t* st, rlrm_eval_t* ev, uint64_t t0,
                    int32_t n_iters, int32_t n_episodes, double gamma, double optimal_steps) {
  int A = cfg->n_agents;
  rlrm_config_t c2 = *cfg;
  c2.decay_on_reset = 0; /* evaluation runs on a deep copy of the env: the training epsilon is untouched */
  rec_t rec[RLRM_MAX_AGENTS];
  uint8_t act[RLRM_MAX_AGENTS];
  for (int64_t i = 0; i < st->n_instances; i++) {
    for (int32_t it = 0; it < n_iters; it++) {
      if ((int32_t)ev[(size_t)i * A].episodes >= n_episodes) break;
      uint64_t t = t0 + (uint64_t)it;
      for (int a = 0; a < A; a++) {
        size_t k = (size_t)i * A + a;
        slot_t s = unpack(st->slot[k]);
        const real* row = (const real*)st->q + table_base(cfg, i, a) + ((size_t)s.cell * agent_rm(cfg, tb, a).nQ + s.rm) * 4;
        uint32_t w[4] = {0, 0, 0, 0};
        act[a] = (uint8_t)select_one(cfg, row, 0.0, w, 1); /* best=True: np.argmax, no randomness (:84-87) */
      }
      step_instance(&c2, tb, st, i, act, NULL, t, 1, rec);
      clear_first(cfg, st, i);
      int all_term = 1, all_trunc = 1;
      for (int a = 0; a < A; a++) {
        size_t k = (size_t)i * A + a;
        rlrm_eval_t* e = &ev[k];
        if (!e->in_success) { /* :96-110 */
          e->disc_return += e->cum_gamma * rec[a].reward;
          { int fin = agent_rm(cfg, tb, a).final; if (rec[a].term && fin >= 0 && (int)rec[a].q == fin) { e->successes++; e->in_success = 1; } }
        }
        e->cum_gamma *= gamma;
        all_term &= (int)rec[a].term; all_trunc &= (int)rec[a].trunc;
      }
      if (all_term || all_trunc) { /* the `timestep > 1000` exit (:117) coincides with truncation */
        for (int a = 0; a < A; a++) {
          size_t k = (size_t)i * A + a;
          rlrm_eval_t* e = &ev[k];
          uint64_t len = unpack(st->slot[k]).time;
          e->episodes++;
          e->return_sum += e->disc_return;
          e->return_sqsum += e->disc_return * e->disc_return;
          if (e->in_success) { e->len_sum += len; e->len_sqsum += len * len; }
          if (len > 0) e->arps_sum += (e->disc_return / (double)len) / optimal_steps; /* :131-139 */
          e->cum_gamma = 1.0; e->disc_return = 0.0; e->in_success = 0;
        }
        reset_instance_at(&c2, tb, st, i, t + 1);
      }
    }
  }
  return 0;
}
