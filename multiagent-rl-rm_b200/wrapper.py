"""RMEnvironmentWrapper — mirrors multi_agent/wrappers/rm_environment_wrapper.py:4-183 (reset, step, QRM experiences).
`step` is ONE device call (env.step + RewardMachine.step + merge, rlrm_step with_rm=1); the QRM counterfactual
transitions reported in infos["qrm_experience"] come from rlrm_rm_step on the same tables. get_mdp
(rm_environment_wrapper.py:185-283) is ONE rlrm_mdp launch per agent instead of reset + set_state + step per
(state, action, sub-action)."""
from __future__ import annotations

import torch

from .encoders import _GridStateEncoder
from .envs import _A2I, _num


def assemble_mdp(next_state, reward, done, terminal, probabilities, empty_nonterminal=False):
    """Arrays of rlrm_mdp ([S,4,n_sub] + terminal [S]) -> the reference's P[s][a] = [(prob, s', reward, done), ...]
    (rm_environment_wrapper.py:218-276): terminal states carry one self-loop entry with probability 1.0.
    probabilities: per nominal action, the list matching that action's sub-actions."""
    S = next_state.shape[0]
    P = {}
    for s in range(S):
        if terminal[s]:
            P[s] = {a: [(1.0, s, _num(reward[s, a, 0]), True)] for a in range(4)}
        elif empty_nonterminal:
            P[s] = {a: [] for a in range(4)}
        else:
            P[s] = {a: [(probabilities[a][j], int(next_state[s, a, j]), _num(reward[s, a, j]), bool(done[s, a, j]))
                        for j in range(len(probabilities[a]))] for a in range(4)}
    return P


class RMEnvironmentWrapper:
    def __init__(self, env, agents):
        self.env = env
        self.agents = agents
        self.reward_modifier = 1

    def reset(self, seed):
        observations, infos = self.env.reset(seed)
        for agent in self.agents:
            agent.get_reward_machine().reset_to_initial_state()
        return observations, infos

    def check_terminations(self):
        return {a.name: a.get_reward_machine().get_current_state() == a.get_reward_machine().get_final_state()
                for a in self.agents}

    def step(self, actions):
        agents = self.agents
        rms = [a.get_reward_machine() for a in agents]
        prev_q = [rm.get_current_state() for rm in rms]
        qrm_agents = [getattr(a.get_learning_algorithm(), "use_qrm", False) for a in agents]
        any_qrm = any(qrm_agents)
        # the same launch also evaluates the hypothetical RM transitions on the new position (cf_q / cf_r of rlrm_step)
        observations, rewards, env_term, env_trunc, infos, rec = self.env._step(actions, with_rm=True,
                                                                                reward_modifier=self.reward_modifier,
                                                                                counterfactuals=any_qrm)
        n = len(agents)
        terminations = {}
        self._cf_cache = {}
        if any_qrm:
            stride = max(1, int(self.env._engine.cfg.n_qrm_states))
            cf_q, cf_r = rec["cf_q"][:n * stride].tolist(), rec["cf_r"][:n * stride].tolist()
            for i, agent in enumerate(agents):
                k = rms[i].numbers_state() - 1
                if qrm_agents[i] and k > 0:
                    self._cf_cache[agent.name] = (cf_q[i * stride:i * stride + k], cf_r[i * stride:i * stride + k])
        rq, rm_term, term = rec["rq"][:n].tolist(), rec["rm_term"][:n].tolist(), rec["term"][:n].tolist()
        for i, agent in enumerate(agents):
            rm, name = rms[i], agent.name
            info = infos[name]
            current_state = info.get("prev_s", observations[name])
            reward_rm = rq[i]
            reward_rm = int(reward_rm) if reward_rm == int(reward_rm) else reward_rm  # _num
            info["RQ"] = reward_rm
            info["prev_q"] = prev_q[i]
            q_now = rm.get_current_state()
            info["q"] = q_now
            info["reward_machine"] = rm
            if qrm_agents[i]:
                info["qrm_experience"] = self._get_qrm_experiences(agent, current_state, observations[name], actions[name],
                                                                   rewards[name], q_now, env_term[name])
            rewards[name] += reward_rm
            info["env_terminated"] = env_term[name]
            info["rm_terminated"] = bool(rm_term[i])
            terminations[name] = bool(term[i])
        return observations, rewards, terminations, env_trunc, infos

    def get_mdp(self, seed, repaired=False):
        """Transition model of every agent's product MDP: (all_P, all_num_states, all_num_actions), same structure and
        values as the reference (rm_environment_wrapper.py:185-283), including its side effects: env.stochastic is
        switched off and stays off (:196-197), and the wrapper ends freshly reset (:282).

        FrozenLake: the reference's StateEncoderFrozenLake.decode hands back an info DICT where get_mdp expects the RM
        state (state_encoder_frozen_lake.py:80-86), so every step raises inside its try/except (:259-264) and P keeps
        EMPTY outcome lists for all non-hole states, with the RM's final state never terminal. That is reproduced as is;
        ``repaired=True`` (an extension, not reference behaviour) builds the MDP the way OfficeWorld's is built."""
        if hasattr(self.env, "stochastic"):
            self.env.stochastic = False
        width = getattr(self.env, "map_width", None) or self.env.grid_width
        height = getattr(self.env, "map_height", None) or self.env.grid_height
        all_P, all_num_states, all_num_actions = {}, {}, {}
        for idx, agent in enumerate(self.env.agents):
            rm = agent.get_reward_machine()
            num_states = width * height * rm.numbers_state()
            actions_list = agent.get_actions()
            if len(actions_list) != 4:
                raise NotImplementedError("get_mdp supports the 4-action grid agents")
            _pos, rm_state0 = agent.encoder.decode(0)
            degenerate = isinstance(rm_state0, dict) and not repaired
            dists = [self.env.get_action_distribution(a) for a in actions_list]
            n_sub = max(len(d[0]) for d in dists)
            sub = [[_A2I[getattr(x, "name", x)] for x in d[0]] + [_A2I["wait"]] * (n_sub - len(d[0])) for d in dists]
            eng = self.env._get_engine(self.reward_modifier)
            nxt, rew, done, term = eng.mdp(idx, sub, rm_terminal=not degenerate)
            if nxt.shape[0] != num_states:
                raise RuntimeError("state-space size mismatch between the encoder and the compiled tables")
            all_P[agent.name] = assemble_mdp(nxt, rew, done, term, [list(d[1]) for d in dists], empty_nonterminal=degenerate)
            all_num_states[agent.name] = num_states
            all_num_actions[agent.name] = len(actions_list)
        self.reset(seed)
        return all_P, all_num_states, all_num_actions

    def _counterfactual_lookups(self, observations):
        """Hypothetical RM transitions (q_u, new position) -> (q', reward) for every agent whose learner uses QRM, evaluated with
        ONE device call when all agents share a machine (per-agent machines: one call per agent, rlrm_rm_step_agent)."""
        wanted = [(i, a) for i, a in enumerate(self.agents) if a.name in observations
                  and getattr(a.get_learning_algorithm(), "use_qrm", False) and a.get_reward_machine().get_all_states()[:-1]]
        if not wanted:
            return {}
        eng = self.env._get_engine(1)  # counterfactual rewards ignore reward_modifier (rm_environment_wrapper.py:150-153)
        W = self.env.grid_width
        per_agent = bool(eng.cfg.per_agent_rm)
        out, q_all, cell_all, spans = {}, [], [], []
        for i, a in wanted:
            rm = a.get_reward_machine()
            states = rm.get_all_states()[:-1]
            obs = observations[a.name]
            q_in = [rm.get_state_index(s) for s in states]
            cells = [obs["pos_y"] * W + obs["pos_x"]] * len(states)
            if per_agent:
                q_out, _ev, r = eng.rm_step(torch.tensor(q_in, dtype=torch.uint8), torch.tensor(cells, dtype=torch.int16), agent=i)
                out[a.name] = (q_out.cpu().numpy(), r.cpu().numpy())
            else:
                spans.append((a.name, len(q_all), len(states)))
                q_all += q_in
                cell_all += cells
        if spans:
            q_out, _ev, r = eng.rm_step(torch.tensor(q_all, dtype=torch.uint8), torch.tensor(cell_all, dtype=torch.int16), agent=0)
            q_out, r = q_out.cpu().numpy(), r.cpu().numpy()
            for name, lo, n in spans:
                out[name] = (q_out[lo:lo + n], r[lo:lo + n])
        return out

    def _get_qrm_experiences(self, agent, current_state, next_state, action, env_reward, next_rm_state, env_termination):
        """Counterfactual transitions for every RM state in get_all_states()[:-1] (rm_environment_wrapper.py:122-183);
        the hypothetical RM transitions are evaluated on the device (rlrm_rm_step)."""
        rm = agent.get_reward_machine()
        states = rm.get_all_states()[:-1]
        if not states:
            return []
        cached = getattr(self, "_cf_cache", {}).get(agent.name)
        if cached is None:  # called outside step(): evaluate just this agent
            cached = self._counterfactual_lookups({agent.name: next_state}).get(agent.name)
        q_out, r = cached
        final = rm.get_final_state()
        a_idx = agent.actions_idx(action)
        out = []
        enc = agent.encoder
        if type(enc).encode is _GridStateEncoder.encode:  # the stock grid encoders: enc = (y*W + x)*nQ + index(q), inlined
            W, _H, n_rm = enc._dims()
            s_cur = current_state["pos_y"] * W + current_state["pos_x"]
            s_nxt = next_state["pos_y"] * W + next_state["pos_x"]
            limit = W * _H * n_rm
            index = rm.state_indices
            from_index = rm.get_state_from_index
            for k, s in enumerate(states):
                qn = int(q_out[k])
                nxt = from_index(qn)
                qs = index[s]
                enc_s, enc_n = s_cur * n_rm + qs, s_nxt * n_rm + qn
                if enc_s >= limit or enc_n >= limit:
                    raise ValueError("Encoded state index exceeds total state space size.")
                ru = float(r[k])
                ru = int(ru) if ru == int(ru) else ru  # _num
                out.append((enc_s, a_idx, env_reward + ru, enc_n, env_termination or nxt == final, s_cur, qs, s_nxt, qn, ru))
            return out
        for k, s in enumerate(states):
            nxt = rm.get_state_from_index(int(q_out[k]))
            enc_s, info_s = enc.encode(current_state, s)
            enc_n, info_n = enc.encode(next_state, nxt)
            ru = _num(r[k])
            out.append((enc_s, a_idx, env_reward + ru, enc_n, env_termination or nxt == final, info_s["s"], info_s["q"],
                        info_n["s"], info_n["q"], ru))
        return out
