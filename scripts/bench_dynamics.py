"""Stacked-scenario dynamics + Jacobians (pgb_dynamics_dev) at the BASELINE.json configs[4] shape: kernel time (CUDA
events inside the call) and host call, K = 10^5 scenarios, n_x = 4, n_phi = 625.  python scripts/bench_dynamics.py [K] [all: every kernel variant]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pgopt_b200 as pg
from pgopt_b200._lib import lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = bench._c5_problem(K)
H, n_x, n_phi = p["H"], p["n_x"], p["n_phi"]
g = pg.LinearMeasurement(p["Cm"])
with pg.DeviceSamples.from_samples(p["batch"], p["basis"]) as ds:
    sc = ds.rollout(g, 0.1, H, u_init=p["u_init"], seed=1).scenarios(want=("X",))
    x = np.ascontiguousarray(sc["X"][:, 5].reshape(-1))
    res = {}
    variants = [("", -1)] if len(sys.argv) <= 2 else [("strided_", 0), ("bulk_w2_", 2), ("bulk_w4_", 1), ("bulk_w8_", 3), ("bulk_auto_", -1)]
    for vname, force in variants:
      lib().pgb_test_dynamics_kernel(force)
      for name, jac in ((vname + "with_jacobians", True), (vname + "state_only", False)):
          ts, ks = [], []
          for i in range(5):
              t0 = time.perf_counter()
              ds.dynamics(x, p["u_init"][:, 5], t=5, phi_variant=1, jacobians=jac)
              ts.append(time.perf_counter() - t0)
              kd = C.c_double(0)
              lib().pgb_dynamics_last_ms(ds._s, C.byref(kd))
              ks.append(kd.value)
          kms = min(ks[1:])
          bytes_alg = K * 8 * (n_x * n_phi + 2 * n_x + (n_x * (n_x + 1) if jac else 0))
          res[name] = {"kernel_ms": kms, "host_call_ms": 1e3 * min(ts[1:]), "hbm_GBps_algorithmic": bytes_alg / (kms * 1e-3) / 1e9,
                       "hbm_frac_of_measured_peak": bytes_alg / (kms * 1e-3) / 1e9 / bench.measured_peaks()[0],
                       "scenario_steps_per_s": K / (kms * 1e-3)}
print(json.dumps({"workload": "pgb_dynamics_dev, K=%d scenarios, n_x=4, n_phi=625 (configs[4] shape)" % K, **res}))
