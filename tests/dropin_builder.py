"""Builds THIS repo's reference-shaped objects for a scenario dict — the counterpart of
oracle/ref_harness.build_reference, so the restated reference driver loop can run unchanged over both."""
import multiagent_rlrm_b200 as P
from multiagent_rlrm_b200.maps import frozen_lake_grid, office_world_grid


def build_b200(sc, table_dtype=None):
    if sc["env"] == "frozen_lake":
        g = frozen_lake_grid(sc["map_name"])
        env = P.MultiAgentFrozenLake(width=g.width, height=g.height, holes=g.hazards)
        env.frozen_lake_stochastic = bool(sc["stochastic"])
        env.penalty_amount = sc["penalty_amount"]
        env.delay_action = bool(sc["delay_action"])
        env.random_start_positions = bool(sc.get("random_start_positions", False))
        SEnc, AEnc = P.StateEncoderFrozenLake, P.ActionEncoderFrozenLake
    else:
        g = office_world_grid(sc["map_name"])
        env = P.MultiAgentOfficeWorld(width=g.width, height=g.height, plants=g.hazards, coffee=g.coffee, letters=g.letters,
                                      walls=g.walls, plants_penalty_value=sc["plants_penalty"],
                                      wall_penalty_value=sc["wall_penalty"], terminate_on_plants=bool(sc["terminate_on_plants"]),
                                      terminate_hit_walls=bool(sc["terminate_hit_walls"]))
        env.all_slip = bool(sc["all_slip"])
        env.stochastic = bool(sc["stochastic"])
        env.high_prob = sc["high_prob"]
        env.delay_action = bool(sc["delay_action"])
        SEnc, AEnc = P.StateEncoderOfficeWorld, P.ActionEncoderOfficeWorld
    def machine(k):
        trs = sc["rm_transitions"] if not sc.get("rm_transitions_per_agent") else sc["rm_transitions_per_agent"][k]
        tmap = {}
        for (s, e, t, r) in trs:
            tmap[(s, None if e is None else tuple(e))] = (t, r)
        if sc.get("detector_positions") is not None:
            pos = {tuple(p) for p in sc["detector_positions"]}
        else:
            pos = {ev for (_s, ev) in tmap if ev is not None}
        return tmap, pos

    agents = []
    for k, (x, y) in enumerate(sc["starts"]):
        transitions, positions = machine(k)
        ag = P.AgentRL(f"a{k + 1}", env)
        ag.set_initial_position(x, y)
        ag.add_state_encoder(SEnc(ag))
        ag.add_action_encoder(AEnc(ag))
        rm = P.RewardMachine(dict(transitions), P.PositionEventDetector(set(positions)))
        ag.set_reward_machine(rm)
        env.add_agent(ag)
        n_states = env.grid_width * env.grid_height * rm.numbers_state()
        import numpy as np

        common = dict(table_dtype="f64" if table_dtype is np.float64 else "f32",
                      state_space_size=n_states, action_space_size=4, learning_rate=sc["learning_rate"], gamma=sc["gamma"],
                      action_selection="greedy", epsilon_start=sc["epsilon_start"], epsilon_end=sc["epsilon_end"],
                      epsilon_decay=sc["epsilon_decay"])
        if sc["algo"] == "qlambda":
            learner = P.QLearningLambda(lambd=sc["lambd"], **common)
            learner.q_table.fill(sc["q_init"])
        else:
            learner = P.QLearning(qtable_init=sc["q_init"], use_qrm=(sc["algo"] == "qrm"), **common)
        if sc.get("use_rsh") and sc["algo"] != "qlambda":
            learner.use_rsh = True
            if sc.get("rs_kind", "vi") == "distance":
                rm.add_distance_reward_shaping(sc["gamma"], sc["rs_gamma"], sc["rs_alpha"])
            else:
                rm.add_reward_shaping(sc["gamma"], sc["rs_gamma"])
        ag.set_learning_algorithm(learner)
        agents.append(ag)
    rm_env = P.RMEnvironmentWrapper(env, agents)
    rm_env.reward_modifier = sc["reward_modifier"]
    return rm_env, env, agents
