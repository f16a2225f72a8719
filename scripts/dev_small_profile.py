"""one batched N = 30 sweep for ncu (k_sweep_small, known basis, 128 chains, T = 300)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import pgopt_b200 as pg
model = sys.argv[1] if len(sys.argv) > 1 else "known"
prob = bench.build_problem(model, 300, 0)
eng = pg.ParticleGibbsEngine(prob["basis"], 1, 30, 300, n_chains=128)
eng.set_data(prob["u"], prob["y"]); eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0)); eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
for c in range(128):
    eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
eng.set_rng_philox(7, 0)
eng.sweep(3); eng.synchronize(); eng.close()
