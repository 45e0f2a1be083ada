"""Times BASELINE configs[0] (frozen_lake_main --map map1: 2 agents, built-in RM, QRM learner) through the reference's own
driver loop on (a) this repo's reference-shaped N = 1 classes (device-backed) and (b) the live Python reference staged in
oracle/_ref, same host, same process, own PCG64 randomness.   python profiles/scripts/time_dropin_n1.py [seconds] [--profile]"""
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import multiagent_rlrm_b200 as P  # noqa: E402


def loop(rm_env, env, agents, fl, seed, seconds, max_iters=10**9):
    active = iters = episodes = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds and iters < max_iters:
        states, _ = rm_env.reset(seed)
        if not fl:
            states = copy.deepcopy(states)
        while True:
            actions = {ag.name: ag.select_action(rm_env.env.get_state(ag)) for ag in rm_env.agents}
            new_states, rewards, term, trunc, infos = rm_env.step(actions)
            for ag in rm_env.agents:
                ta = (term[ag.name] or trunc[ag.name]) if fl else term[ag.name]
                ag.update_policy(state=states[ag.name], action=actions[ag.name], reward=rewards[ag.name],
                                 next_state=new_states[ag.name], terminated=ta, infos=infos[ag.name])
            states = copy.deepcopy(new_states)
            iters += 1
            if all(term.values()) or all(trunc.values()):
                break
        active += sum(env.agent_steps.values())
        episodes += 1
    dt = time.perf_counter() - t0
    return {"episodes": episodes, "iterations": iters, "active_agent_steps": active, "seconds": dt,
            "active_agent_steps_per_s": active / dt, "us_per_iteration": dt / max(iters, 1) * 1e6}


def breakdown(rm_env, env, agents, fl, seed, n_iters=4000):
    """Where one driver-loop iteration goes: wall time inside select_action / rm_env.step / update_policy / the rest."""
    acc = {"select_action": 0.0, "step": 0.0, "update_policy": 0.0}
    pc = time.perf_counter
    iters, t_all = 0, pc()
    while iters < n_iters:
        states, _ = rm_env.reset(seed)
        if not fl:
            states = copy.deepcopy(states)
        while True:
            t0 = pc()
            actions = {ag.name: ag.select_action(rm_env.env.get_state(ag)) for ag in rm_env.agents}
            t1 = pc()
            new_states, rewards, term, trunc, infos = rm_env.step(actions)
            t2 = pc()
            for ag in rm_env.agents:
                ta = (term[ag.name] or trunc[ag.name]) if fl else term[ag.name]
                ag.update_policy(state=states[ag.name], action=actions[ag.name], reward=rewards[ag.name],
                                 next_state=new_states[ag.name], terminated=ta, infos=infos[ag.name])
            t3 = pc()
            acc["select_action"] += t1 - t0
            acc["step"] += t2 - t1
            acc["update_policy"] += t3 - t2
            states = copy.deepcopy(new_states)
            iters += 1
            if all(term.values()) or all(trunc.values()):
                break
    total = pc() - t_all
    out = {k: v / iters * 1e6 for k, v in acc.items()}
    out["other (reset, deepcopy, loop)"] = (total - sum(acc.values())) / iters * 1e6
    out["total_us_per_iteration"] = total / iters * 1e6
    return out


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 5.0
    out = {}
    for name, sc in (("cfg1", P.scenario_config1()), ("cfg3_slip_ql", P.scenario_config3(False)), ("cfg2_office_slip_ql", P.scenario_config2(True))):
        d = sc.to_dict()
        fl = d["driver"] == "frozen_lake_main"
        from dropin_builder import build_b200

        rm_env, env, agents = build_b200(d)
        loop(rm_env, env, agents, fl, d["seed"], 1.0)  # warm
        res = {"b200_dropin": loop(rm_env, env, agents, fl, d["seed"], seconds)}
        if "--breakdown" in sys.argv:
            res["b200_dropin_breakdown_us"] = breakdown(rm_env, env, agents, fl, d["seed"])
        if "--profile" in sys.argv and name == "cfg1":
            import cProfile
            import pstats

            pr = cProfile.Profile()
            pr.enable()
            loop(rm_env, env, agents, fl, d["seed"], 1e9, max_iters=3000)
            pr.disable()
            pstats.Stats(pr).sort_stats("tottime").print_stats(28)
        try:
            import numpy as np
            import ref_harness as H

            if H.reference_available():
                r_env, r_e, r_agents = H.build_reference(d, np.float64)
                loop(r_env, r_e, r_agents, fl, d["seed"], 0.5)
                res["python_reference"] = loop(r_env, r_e, r_agents, fl, d["seed"], seconds)
                if "--breakdown" in sys.argv:
                    res["python_reference_breakdown_us"] = breakdown(r_env, r_e, r_agents, fl, d["seed"])
        except Exception as exc:  # the reference is optional here
            res["python_reference"] = {"unavailable": repr(exc)}
        out[name] = res
        print(name, json.dumps(res), flush=True)
    print(json.dumps({"n1_dropin_vs_reference": out}))


if __name__ == "__main__":
    main()
