import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, bench, pgopt_b200 as pg
from pgopt_b200._lib import lib, check
model = sys.argv[1] if len(sys.argv) > 1 else "known"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
prob = bench.build_problem(model, T, 0)
eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T)
eng.set_data(prob["u"], prob["y"]); eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0)); eng.set_rng_philox(1234, 0)
for i in range(3):
    eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2)
    check(lib().pgb_timer_start(eng._h)); eng.sweep_filter(); t = C.c_double(0); check(lib().pgb_timer_stop(eng._h, C.byref(t)))
cyc = (C.c_double * 16)(); n = C.c_int32(0)
check(lib().pgb_get_phase_cycles(eng._h, cyc, C.byref(n)))
names = ["init", "A prep", "wait B1", "B scan-local main", "B normalise side", "wait B2", "C resolve", "C prefixes+expand",
         "C scan-local side", "wait B3", "D side resolve+index", "D heavy+weights", "wait B4", "fixup test", "-", "-"]
tot = sum(cyc)
print("model", model, "N", N, "T", T, "sweep ms", t.value, "us/step", 1e3 * t.value / (T - 1), "CTAs", n.value, "status", eng.status())
for i, nm in enumerate(names):
    if cyc[i]: print("  %-24s %10.0f cyc/step  %5.1f%%" % (nm, cyc[i] / (T - 1), 100 * cyc[i] / tot))
