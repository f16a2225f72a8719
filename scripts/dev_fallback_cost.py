"""cost of the serial exact fallback of the sweep kernel: N = 2^20, T = 24, every scan forced to fall back vs none"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import pgopt_b200 as pg
from pgopt_b200._lib import lib
N, T = 1 << 20, 24
prob = bench.build_problem("generic", T, 0)
eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=1)
eng.set_data(prob["u"], prob["y"]); eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0)); eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
eng.set_rng_philox(1234, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
res = {}
for force in (0, 1, 0):
    lib().pgb_test_force_fallback(force)
    ms = bench.time_filter(eng, prob, 2, 1)
    st = eng.status(0, clear=True)
    res["force=%d" % force] = {"ms_per_sweep": ms, "status": list(st)}
lib().pgb_test_force_fallback(0)
print(json.dumps(res))
