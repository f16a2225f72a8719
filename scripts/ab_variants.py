import sys, os, ctypes as C, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
# run phase_profile-like timing for each lib variant and mode in a fresh process each
code = r'''
import sys, os, ctypes as C
sys.path.insert(0, ".")
import numpy as np, bench, pgopt_b200 as pg
from pgopt_b200._lib import lib, check
model, N, T, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
prob = bench.build_problem(model, T, 0)
eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T)
eng.set_mode(mode)
eng.set_data(prob["u"], prob["y"]); eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0)); eng.set_rng_philox(1234, 0)
best = 1e9
for i in range(3):
    eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2)
    check(lib().pgb_timer_start(eng._h)); eng.sweep_filter(); t = C.c_double(0); check(lib().pgb_timer_stop(eng._h, C.byref(t)))
    best = min(best, t.value)
print("%.1f us/step" % (1e3 * best / (T - 1)))
'''
model = sys.argv[1] if len(sys.argv) > 1 else "known"
libs = sorted(f for f in os.listdir("pgopt_b200/lib") if f.startswith("libv_"))
for l in libs:
    for mode in [int(m) for m in os.environ.get("AB_MODES", "1").split(",")]:
        env = dict(os.environ, PGOPT_B200_LIB=os.path.abspath(os.path.join("pgopt_b200/lib", l)))
        r = subprocess.run([sys.executable, "-c", code, model, "1048576", "80", str(mode)], env=env, capture_output=True, text=True)
        print(l, "mode", mode, r.stdout.strip(), r.stderr.strip()[-200:])
