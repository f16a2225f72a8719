"""BASELINE.json configs[0] / configs[1] — the reference's own examples (N = 30 particles, T = 2000; known basis and the
5x5x5 Hilbert-GP basis) as PG sweeps/s (filter + trace + statistics + MNIW): one chain, and `nb` independent chains
batched in one handle (the way a GPU is filled at this particle count), next to the C oracle on one host core."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bench  # noqa: E402
import pgopt_b200 as pg  # noqa: E402


def run(model, n_chains, sweeps, N=30, T=2000):
    prob = bench.build_problem(model, T, 0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=n_chains)
    try:
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(n_chains):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        eng.set_rng_philox(7, 0)
        eng.sweep(2)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.sweep(sweeps)
        eng.synchronize()
        dt = time.perf_counter() - t0
        st = list(eng.status(0))
    finally:
        eng.close()
    return n_chains * sweeps / dt, st


def oracle(model, sweeps=2, N=30, T=2000):
    import _oracle as O
    prob = bench.build_problem(model, T, 0)
    om = O.known_model() if model == "known" else O.hilbert_model(2, 1, prob["basis"].count, prob["basis"].stride, prob["basis"].L)
    t0 = time.perf_counter()
    O.particle_gibbs(om, N, T, prob["u"], prob["y"], [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), prob["A"], prob["Q"],
                     prob["V"], 100 * np.eye(2), 10.0, sweeps, 0, 0, seed=7, chain=0)
    return sweeps / (time.perf_counter() - t0)


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for model in ("known", "generic"):
        one, st1 = run(model, 1, 20)
        many, stn = run(model, nb, 10)
        res = {"config": "N=30, T=2000, %s basis (BASELINE.json configs[%d])" % (model, 0 if model == "known" else 1),
               "pg_sweeps_per_s_one_chain": one, "pg_sweeps_per_s_%d_chains_batched" % nb: many, "status": [st1, stn]}
        try:
            res["pg_sweeps_per_s_oracle_one_core"] = oracle(model)
        except Exception as e:
            res["oracle"] = "unavailable: %r" % e
        print(json.dumps(res))
