"""Development aid (experiment build only): per-section cycle split of the v2 sweep kernel.
   python scripts/dev_phase.py [model] [N] [T]   with PGOPT_B200_LIB pointing at a -DPGB_EXPERIMENT build"""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from pgopt_b200._lib import lib, check
model = sys.argv[1] if len(sys.argv) > 1 else "generic"
dtype = sys.argv[4] if len(sys.argv) > 4 else "f64"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
import pgopt_b200 as pg
prob = bench.build_problem(model, T, seed=0)
eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=1, precision=dtype)
eng.set_data(prob["u"], prob["y"])
eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
eng.set_rng_philox(1234, 0)
for _ in range(2):
    eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2)
    eng.sweep_filter()
eng.synchronize()
cyc = (C.c_double * 16)()
check(lib().pgb_experiment_phase_cycles(eng._h, cyc), eng._h)
names = {1: "PA", 2: "wait B1", 13: "PB collective", 3: "PB rest", 4: "wait B2", 14: "PC collective", 15: "PC scan-local", 5: "PC shortcut+rest", 6: "wait B3", 12: "PD resolve", 11: "PD prefixes+counts", 0: "PD scatter", 7: "PD emit+max", 8: "(PD end)", 9: "wait B0", 10: "rare"}
tot = sum(cyc)
for i in (1, 2, 13, 3, 4, 14, 15, 5, 6, 12, 11, 0, 7, 8, 9, 10):
    print("%-22s %8.1f cyc/step  %5.1f %%" % (names[i], cyc[i] / (T - 1), 100 * cyc[i] / tot))
print("total %.1f cycles/step = %.1f us at 1.965 GHz" % (tot / (T - 1), tot / (T - 1) / 1965.0))
