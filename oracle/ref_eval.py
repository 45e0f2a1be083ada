"""TEST INFRASTRUCTURE — greedy-evaluation reference: a restatement of test_policy_optima
(/root/reference/multiagent_rlrm/environments/utils_envs/evaluation_metrics.py:23-190) with injected slip draws, plus a
self-check against the reference's OWN function on deterministic dynamics (where no randomness is consumed)."""
from __future__ import annotations

import numpy as np

import ref_harness as H


def reference_eval(sc: dict, q_tables, n_episodes: int, gamma: float, optimal_steps: float, t0: int = 0):
    """q_tables: float32 [N, A, S, 4]. Returns dict of per-(instance, agent) aggregates, same fields as rlrm_eval_t."""
    n, A = q_tables.shape[0], q_tables.shape[1]
    src = H.DrawSource(sc["seed"], n, A)
    out = {k: np.zeros((n, A)) for k in ("return_sum", "return_sqsum", "arps_sum")}
    out.update({k: np.zeros((n, A), dtype=np.int64) for k in ("len_sum", "len_sqsum", "episodes", "successes")})
    own = []
    for i in range(n):
        rm_env, env, agents = H.build_reference(sc, np.float32)
        for k, ag in enumerate(agents):
            ag.get_learning_algorithm().q_table = q_tables[i, k].astype(np.float32).copy()
        index_of = {ag.name: k for k, ag in enumerate(agents)}
        orig = env.get_stochastic_action

        def gsa(agent, intended, _orig=orig, _env=env, _idx=index_of):
            _env.rng.agent_index = _idx[agent.name]
            return _orig(agent, intended)

        env.get_stochastic_action = gsa
        t = t0
        for ep in range(n_episodes):
            states, infos = rm_env.reset(10000 + ep)
            env.rng = H.EnvRNG(src, i)
            done = {ag.name: False for ag in agents}
            success = {ag.name: False for ag in agents}
            ep_rew = {ag.name: 0 for ag in agents}
            cum_gamma, timestep = 1.0, 0
            while not all(done.values()):
                src.t = t
                actions = {ag.name: ag.select_action(rm_env.env.get_state(ag), best=True) for ag in agents}
                new_states, rewards, done, trunc, infos = rm_env.step(actions)
                for ag in agents:
                    if not success[ag.name]:
                        ep_rew[ag.name] += cum_gamma * rewards[ag.name]
                        rm = ag.get_reward_machine()
                        if done[ag.name] and rm.get_current_state() == rm.get_final_state():
                            out["successes"][i, index_of[ag.name]] += 1
                            success[ag.name] = True
                cum_gamma *= gamma
                timestep += 1
                t += 1
                if all(done.values()) or all(trunc.values()) or timestep > 1000:
                    break
            for k, ag in enumerate(agents):
                out["episodes"][i, k] += 1
                out["return_sum"][i, k] += ep_rew[ag.name]
                out["return_sqsum"][i, k] += ep_rew[ag.name] * ep_rew[ag.name]
                if success[ag.name]:
                    out["len_sum"][i, k] += timestep
                    out["len_sqsum"][i, k] += timestep * timestep
                if timestep > 0:
                    out["arps_sum"][i, k] += (ep_rew[ag.name] / timestep) / optimal_steps
        if not sc["stochastic"]:
            # the reference's own function on the same objects (deterministic: no draws are consumed)
            from multiagent_rlrm.environments.utils_envs.evaluation_metrics import test_policy_optima

            own.append(test_policy_optima(rm_env, episodi_test=n_episodes, window_size=1, optimal_steps=optimal_steps, gamma=gamma))
            # explicit-policy evaluation (test_policy_opt_multi, evaluation_metrics.py:505-697) of the greedy policy of the same tables
            from multiagent_rlrm.environments.utils_envs.evaluation_metrics import test_policy_opt_multi

            policy = {ag.name: np.argmax(ag.get_learning_algorithm().q_table, axis=1) for ag in agents}
            res = test_policy_opt_multi(rm_env, policy, episodes_test=n_episodes, window_size=1, optimal_steps=optimal_steps, gamma=gamma)
            for key, val in zip(("success_rate", "avg_timesteps", "avg_reward", "avg_arps"), (res[0], res[2], res[4], res[6])):
                out.setdefault("optmulti_" + key, np.zeros((n, A)))[i] = [float(val[ag.name]) for ag in agents]
    out["reference_function_outputs"] = own
    return out
