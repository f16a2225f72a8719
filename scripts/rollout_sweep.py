"""Kernel time of pgb_rollout vs horizon H and particle count N (separates per-scenario setup from per-step cost)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pgopt_b200 as pg  # noqa: E402
from pgopt_b200._lib import lib  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.default_rng(0)
basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
n_x, n_phi = 4, basis.n_phi
A = rng.normal(0, 0.05, (K, n_x, n_phi))
B = rng.normal(0, 0.1, (K, n_x, n_x))
Q = B @ np.transpose(B, (0, 2, 1)) + 0.05 * np.eye(n_x)
um = rng.normal(0, 1, (K, 1))
Cm = np.zeros((1, n_x)); Cm[0, 0] = 1.0
for H, N in [(1, 30), (11, 30), (41, 30), (41, 2), (81, 30)]:
    w = rng.uniform(size=(K, N)); w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    u_init = np.sin(np.arange(H))[None, :]
    pg.scenario_rollout(samples, basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=1)
    tms, kms = C.c_double(0), C.c_double(0)
    lib().pgb_rollout_last_ms(C.byref(tms), C.byref(kms))
    print("K=%d H=%d N=%d kernel %.3f ms  (%.2f us per scenario-CTA-slot at 888 resident)" % (K, H, N, kms.value, kms.value * 1e3 / K * 888))
