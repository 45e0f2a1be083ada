"""TEST INFRASTRUCTURE — times the LIVE Python reference (/root/reference under oracle/ref_shim) with its own PCG64
randomness, restating only the driver loops (frozen_lake_main.py:336-376 / office_main.py:1696-1749). Runs in the
build container (the reference cannot travel to the GPU box); output: profiles/r01_reference_python_container.json.
    python oracle/time_reference_python.py [seconds_per_case]
"""
import copy
import json
import multiprocessing as mp
import os
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, _HERE)

import numpy as np  # noqa: E402

import multiagent_rlrm_b200 as P  # noqa: E402
import ref_harness as H  # noqa: E402


def run_case(args):
    name, sc, seconds, max_episodes, seed_mode = args
    rm_env, env, agents = H.build_reference(sc, np.float64)  # the reference's native float64 tables
    fl = sc["driver"] == "frozen_lake_main"
    rm_env.reset(sc["seed"])
    active = slots = iters = episodes = 0
    t0 = time.perf_counter()
    while (time.perf_counter() - t0 < seconds) and episodes < max_episodes:
        states, _ = rm_env.reset(sc["seed"] if seed_mode == "fixed" else sc["seed"] * 1000 + episodes)
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
        slots = iters * len(agents)
        episodes += 1
    dt = time.perf_counter() - t0
    return {"case": name, "episodes": episodes, "iterations": iters, "active_agent_steps": active, "seconds": dt,
            "active_agent_steps_per_s": active / dt, "slot_steps_per_s": slots / dt}


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    procs = os.cpu_count() or 1
    cases = {
        "config1 FrozenLake map1 det, 2 agents, QRM lr=1 (2000 episodes max)": (P.scenario_config1().to_dict(), 2000, "fixed"),
        "config2 OfficeWorld map1, A->C->B->D spec, QL, deterministic": (P.scenario_config2(False).to_dict(), 10**9, "episode"),
        "config3 FrozenLake slippery, 2 agents, QRM": (P.scenario_config3(True).to_dict(), 10**9, "fixed"),
        "config4 OfficeWorld 12-state RM, 4 agents, Q(lambda) dense": (P.scenario_config4().to_dict(), 10**9, "episode"),
        "config5 FrozenLake slippery, 4 agents, QRM": (P.scenario_config5(False).to_dict(), 10**9, "fixed"),
    }
    out = {"host": "build container (NOT the B200 host)", "cores": procs, "python": sys.version.split()[0], "numpy": np.__version__,
           "what": "live reference classes, own PCG64 streams, float64 tables", "single_process": [], "all_cores": []}
    for name, (sc, max_ep, mode) in cases.items():
        out["single_process"].append(run_case((name, sc, seconds, max_ep, mode)))
        print(out["single_process"][-1], flush=True)
    with mp.get_context("fork").Pool(procs) as pool:
        for name, (sc, max_ep, mode) in list(cases.items())[2:]:
            res = pool.map(run_case, [(name, dict(sc, seed=sc["seed"] + k), seconds, max_ep, mode) for k in range(procs)])
            agg = {"case": name, "processes": procs, "active_agent_steps_per_s": sum(r["active_agent_steps_per_s"] for r in res),
                   "slot_steps_per_s": sum(r["slot_steps_per_s"] for r in res)}
            out["all_cores"].append(agg)
            print(agg, flush=True)
    path = os.path.join(os.path.dirname(_HERE), "profiles", "r01_reference_python_container.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
