"""Value iteration on the product MDP — the consumer of RMEnvironmentWrapper.get_mdp in the reference's "VI comparison"
(office_main.py:1117-1140: get_mdp -> value_iteration -> policy -> test_policy_opt_multi), mirrors
/root/reference/multiagent_rlrm/environments/utils_envs/mdp_vi.py:9-60.

The sweeps run on the device (rlrm_value_iteration): one thread per state, all states of a sweep in parallel from the
previous value function (Jacobi), where the reference updates V in place state by state (Gauss-Seidel). Both stop on
the same rule and converge to the same fixed point, so the value functions agree within 2*theta*gamma/(1-gamma) — not bit
for bit; tests/test_mdp.py states and checks that tolerance against the reference's own function."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from ._lib import check, load


def mdp_to_arrays(P, num_states: int, num_actions: int = 4):
    """P[s][a] = [(prob, s_next, reward, done), ...] -> padded arrays prob / next_state / reward / done [S, 4, n_out]."""
    if num_actions != 4:
        raise NotImplementedError("the grid-world product MDPs have 4 actions")
    n_out = max(1, max((len(P[s][a]) for s in range(num_states) for a in range(4)), default=1))
    prob = np.zeros((num_states, 4, n_out), dtype=np.float64)
    nxt = np.zeros((num_states, 4, n_out), dtype=np.int32)
    rew = np.zeros((num_states, 4, n_out), dtype=np.float64)
    done = np.zeros((num_states, 4, n_out), dtype=np.uint8)
    for s in range(num_states):
        for a in range(4):
            for j, (p, sn, r, d) in enumerate(P[s][a]):
                prob[s, a, j], nxt[s, a, j], rew[s, a, j], done[s, a, j] = p, sn, r, bool(d)
    return prob, nxt, rew, done


def value_iteration_arrays(prob, next_state, reward, done, gamma=0.9, theta=1e-3, delta_rel=False, device="cuda:0",
                           max_sweeps=1_000_000) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """rlrm_value_iteration on padded outcome arrays (numpy or tensors). Returns device tensors V [S], policy [S], Q [S, 4]
    and the number of sweeps."""
    if not torch.cuda.is_available():
        raise RuntimeError("multiagent-rl-rm_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device(device)
    t = lambda x, dt: torch.as_tensor(x).to(device=dev, dtype=dt).contiguous()  # noqa: E731
    prob, nxt, rew, dn = t(prob, torch.float64), t(next_state, torch.int32), t(reward, torch.float64), t(done, torch.uint8)
    S, four, n_out = prob.shape
    if four != 4 or nxt.shape != prob.shape or rew.shape != prob.shape or dn.shape != prob.shape:
        raise ValueError("prob / next_state / reward / done must all be [S, 4, n_out]")
    if int(nxt.min()) < 0 or int(nxt.max()) >= S:
        raise ValueError("next_state holds a state index outside 0..S-1")
    V = torch.empty(S, dtype=torch.float64, device=dev)
    Q = torch.empty((S, 4), dtype=torch.float64, device=dev)
    policy = torch.empty(S, dtype=torch.int32, device=dev)
    work = torch.empty(S + 1, dtype=torch.float64, device=dev)
    sweeps = C.c_int32(0)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    check(load().rlrm_value_iteration(index, S, n_out, prob.data_ptr(), nxt.data_ptr(), rew.data_ptr(), dn.data_ptr(), float(gamma),
                                      float(theta), int(bool(delta_rel)), int(max_sweeps), V.data_ptr(), Q.data_ptr(), policy.data_ptr(),
                                      work.data_ptr(), C.byref(sweeps), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return V, policy, Q, int(sweeps.value)


def value_iteration(P, num_states, num_actions, gamma=0.9, theta=1e-3, delta_rel=False, device="cuda:0"):
    """Drop-in for mdp_vi.value_iteration: (V, policy, Q) as numpy arrays, policy[s] = first argmax of Q[s]."""
    V, policy, Q, _sweeps = value_iteration_arrays(*mdp_to_arrays(P, num_states, num_actions), gamma=gamma, theta=theta,
                                                   delta_rel=delta_rel, device=device)
    return V.cpu().numpy(), policy.cpu().numpy().astype(np.int64), Q.cpu().numpy()
