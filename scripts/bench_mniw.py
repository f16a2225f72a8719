"""k_mniw (MNIW_sample, particle_Gibbs.jl:56-81) device time per draw inside PG sweeps of the reference's generic example
(n_phi = 125, N = 30) and of the known-basis one (n_phi = 5), from the library's per-class event profile; one chain and 64
batched chains.  python scripts/bench_mniw.py"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import pgopt_b200 as pg
from pgopt_b200._lib import lib, check

res = {}
for model in ("generic", "known"):
    for nc in (1, 64):
        T, N, sweeps = 400, 30, 10
        prob = bench.build_problem(model, T, 0)
        eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=nc)
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(nc):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        eng.set_rng_philox(7, 0)
        eng.sweep(2)
        eng.synchronize()
        check(lib().pgb_profile(eng._h, 1), eng._h)
        eng.sweep(sweeps)
        n = len(bench.PROF_NAMES)
        launches, pms = (C.c_int64 * n)(), (C.c_double * n)()
        check(lib().pgb_get_profile(eng._h, launches, pms), eng._h)
        check(lib().pgb_profile(eng._h, 0), eng._h)
        prof = {bench.PROF_NAMES[i]: round(pms[i] / max(launches[i], 1), 4) for i in range(n) if launches[i]}
        res["%s_nphi%d_chains%d" % (model, prob["basis"].n_phi, nc)] = {"ms_per_launch": prof, "status": list(eng.status(0))}
        eng.close()
print(json.dumps({"workload": "k_mniw inside PG sweeps, N=30, T=400", **res}))
