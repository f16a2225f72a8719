"""Device time of the scenario-screening kernel (pgb_scenario_constraint_max) at configs[4]'s shape and larger, against
the measured HBM peak.  Prints one JSON line.  Usage: python scripts/bench_screening.py [K] [H]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import pgopt_b200 as pg
from pgopt_b200 import _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 41
rng = np.random.default_rng(0)
Y = rng.normal(size=(K, H, 1)) * 3
lo = np.full((1, H), -np.inf); lo[0, H // 2:H // 2 + 6] = 2.0
hi = np.full((1, H), np.inf)
kms, hms, wall = [], [], []
for it in range(8):
    t = time.perf_counter()
    h, order = pg.scenario_constraint_max(Y, lo, hi)
    wall.append(time.perf_counter() - t)
    a, b = C.c_double(0), C.c_double(0)
    _lib.lib().pgb_scenario_constraint_last_ms(C.byref(a), C.byref(b))
    hms.append(a.value); kms.append(b.value)
t = time.perf_counter()
ref = np.argsort(-(2.0 - Y[:, H // 2:H // 2 + 6, 0]).max(axis=1), kind="stable") + 1
t_np = time.perf_counter() - t
assert np.array_equal(order, ref)
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except Exception:
    pass
k_ms = float(np.median(kms[3:]))
print(json.dumps({"workload": "scenario screening K=%d H=%d n_y=1" % (K, H), "kernel_ms": k_ms,
                  "algorithmic_bytes": K * H * 8 + K * 8, "hbm_GBps_algorithmic": (K * H * 8 + K * 8) / k_ms / 1e6,
                  "h2d_ms": float(np.median(hms[3:])), "host_call_ms": float(np.median(wall[3:])) * 1e3,
                  "numpy_vectorised_ms": t_np * 1e3, "measured_peaks": peak}))
