"""Potential-based reward shaping on the Reward Machine graph, host side.

Mirrors RewardMachine.add_reward_shaping / value_iteration (reward_machine.py:197-216, 301-345) and
add_distance_reward_shaping / get_distance (reward_machine.py:218-278). The potentials become the ``phi[nQ]`` device
table; the kernels add ``gamma*Phi(q') - Phi(q)`` to the reward exactly as qlearning.py:51-66, 93-105 do."""
from __future__ import annotations

from collections import deque


def value_iteration_potentials(rm, rs_gamma, tol=1e-7):
    """Phi(u) = -V(u), V from in-place value iteration over the RM graph; between two states the LAST inserted
    transition's reward counts (delta_r[u1][u2] is overwritten, reward_machine.py:284-299)."""
    states = list(rm.state_indices.keys())
    succ = {u: {} for u in states}
    for (u1, _ev), (u2, reward) in rm.transitions.items():
        succ.setdefault(u1, {})[u2] = reward
        succ.setdefault(u2, {})
    V = {u: 0 for u in states}
    err = 1
    while err > tol:
        err = 0
        for u1 in states:
            if not succ[u1]:
                continue
            best = max(r + rs_gamma * V[u2] for u2, r in succ[u1].items())
            err = max(err, abs(best - V[u1]))
            V[u1] = best
    return {u: -v for u, v in V.items()}


def rm_distance(rm, start_state):
    """Minimum number of RM transitions from start_state to the final state, 999999 if unreachable."""
    final = rm.get_final_state()
    if start_state == final:
        return 0
    seen, queue = {start_state}, deque([(start_state, 0)])
    while queue:
        cur, d = queue.popleft()
        if cur == final:
            return d
        for (u, _ev), (v, _r) in rm.transitions.items():
            if u == cur and v not in seen:
                seen.add(v)
                queue.append((v, d + 1))
    return 999999


def distance_potentials(rm, alpha):
    final = rm.get_final_state()
    return {u: (0 if u == final else -alpha * rm_distance(rm, u)) for u in rm.get_all_states()}
