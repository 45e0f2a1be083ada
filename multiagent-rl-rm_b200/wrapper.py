"""RMEnvironmentWrapper — mirrors multi_agent/wrappers/rm_environment_wrapper.py:4-183 (reset, step, QRM experiences).
`step` is ONE device call (env.step + RewardMachine.step + merge, rlrm_step with_rm=1); the QRM counterfactual
transitions reported in infos["qrm_experience"] come from rlrm_rm_step on the same tables. get_mdp (VI tooling) is
out of scope."""
from __future__ import annotations

import torch

from .envs import _num


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
        prev_q = {a.name: a.get_reward_machine().get_current_state() for a in self.agents}
        observations, rewards, env_term, env_trunc, infos, rec = self.env._step(actions, with_rm=True,
                                                                                reward_modifier=self.reward_modifier)
        terminations = {}
        for i, agent in enumerate(self.agents):
            rm = agent.get_reward_machine()
            info = infos[agent.name]
            current_state = info.get("prev_s", observations[agent.name])
            reward_rm = _num(rec["rq"][i])
            info["RQ"] = reward_rm
            info["prev_q"] = prev_q[agent.name]
            info["q"] = rm.get_current_state()
            info["reward_machine"] = rm
            if getattr(agent.get_learning_algorithm(), "use_qrm", False):
                info["qrm_experience"] = self._get_qrm_experiences(agent, current_state, observations[agent.name],
                                                                   actions[agent.name], rewards[agent.name],
                                                                   rm.get_current_state(), env_term[agent.name])
            rewards[agent.name] += reward_rm
            info["env_terminated"] = env_term[agent.name]
            info["rm_terminated"] = bool(rec["rm_term"][i])
            terminations[agent.name] = bool(rec["term"][i])
        return observations, rewards, terminations, env_trunc, infos

    def _get_qrm_experiences(self, agent, current_state, next_state, action, env_reward, next_rm_state, env_termination):
        """Counterfactual transitions for every RM state in get_all_states()[:-1] (rm_environment_wrapper.py:122-183);
        the hypothetical RM transitions are evaluated on the device (rlrm_rm_step)."""
        rm = agent.get_reward_machine()
        states = rm.get_all_states()[:-1]
        if not states:
            return []
        eng = self.env._get_engine(1)  # counterfactual rewards ignore reward_modifier (rm_environment_wrapper.py:150-153)
        W = self.env.grid_width
        cell = next_state["pos_y"] * W + next_state["pos_x"]
        q_in = torch.tensor([rm.get_state_index(s) for s in states], dtype=torch.uint8)
        q_out, _ev, r = eng.rm_step(q_in, torch.full((len(states),), cell, dtype=torch.int16), agent=self.agents.index(agent))
        q_out, r = q_out.cpu().numpy(), r.cpu().numpy()
        final = rm.get_final_state()
        a_idx = agent.actions_idx(action)
        out = []
        for k, s in enumerate(states):
            nxt = rm.get_state_from_index(int(q_out[k]))
            enc_s, info_s = agent.encoder.encode(current_state, s)
            enc_n, info_n = agent.encoder.encode(next_state, nxt)
            ru = _num(r[k])
            out.append((enc_s, a_idx, env_reward + ru, enc_n, env_termination or nxt == final, info_s["s"], info_s["q"],
                        info_n["s"], info_n["q"], ru))
        return out
