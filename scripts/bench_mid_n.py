"""Mid-size particle counts (32 < N <= 1024): the one-CTA-per-chain kernel (mode 3) against the tile persistent kernel
(mode 1), PG sweeps/s at T = 500 for one chain and for 64 batched chains, both bases.  One JSON line per (model, N)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import pgopt_b200 as pg  # noqa: E402


def run(model, N, T, n_chains, sweeps, mode):
    prob = bench.build_problem(model, T, 0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=n_chains, x_history="store")
    try:
        eng.set_mode(mode)
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(n_chains):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        eng.set_rng_philox(7, 0)
        eng.sweep(1)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.sweep(sweeps)
        eng.synchronize()
        dt = time.perf_counter() - t0
        assert list(eng.status(0))[0] == 0
        return n_chains * sweeps / dt
    finally:
        eng.close()


if __name__ == "__main__":
    T = 500
    for model in ("known", "generic"):
        for N in (64, 128, 256, 512, 1024):
            r = {"model": model, "N": N, "T": T}
            for nc in (1, 64):
                r["cta_kernel_%d_chains" % nc] = run(model, N, T, nc, 4, 3)
                r["tile_kernel_%d_chains" % nc] = run(model, N, T, nc, 4, 1)
            r["us_per_step_one_chain_cta"] = 1e6 / (r["cta_kernel_1_chains"] * (T - 1))
            r["us_per_step_one_chain_tile"] = 1e6 / (r["tile_kernel_1_chains"] * (T - 1))
            print(json.dumps(r), flush=True)
