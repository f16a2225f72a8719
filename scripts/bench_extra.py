"""Secondary measurements (not the bench.py headline): scenario rollout (BASELINE.json configs[4] shape) and
chain-batched PG sweeps (configs[3] shape).  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import pgopt_b200 as pg  # noqa: E402


def rollout(K=100000, H=41, N=30, cpu_K=300):
    rng = np.random.default_rng(0)
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])  # n_x = 4, n_phi = 625
    n_x, n_phi = 4, basis.n_phi
    A = rng.normal(0, 0.05, (K, n_x, n_phi))
    B = rng.normal(0, 0.1, (K, n_x, n_x))
    Q = B @ np.transpose(B, (0, 2, 1)) + 0.05 * np.eye(n_x)
    w = rng.uniform(size=(K, N)); w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, 1))
    Cm = np.zeros((1, n_x)); Cm[0, 0] = 1.0
    u_init = np.sin(np.arange(H))[None, :]
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    pg.scenario_rollout(samples[:64], basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=1)  # warm-up
    t0 = time.perf_counter()
    out = pg.scenario_rollout(samples, basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=1)
    dt = time.perf_counter() - t0
    import ctypes as C
    from pgopt_b200._lib import lib
    tms, kms = C.c_double(0), C.c_double(0)
    lib().pgb_rollout_last_ms(C.byref(tms), C.byref(kms))
    batch = pg.PGSampleBatch(A, Q, w, np.transpose(x, (0, 2, 1)), um)  # stacked arrays: no per-sample Python packing
    t0 = time.perf_counter()
    out_b = pg.scenario_rollout(batch, basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=1)
    dt_batch = time.perf_counter() - t0
    a_bytes = K * n_x * n_phi * 8
    res = {"metric": "scenario rollout (host buffers in/out, incl. packing + H2D of A)", "K": K, "H": H, "n_x": n_x,
           "n_phi": n_phi, "seconds": dt, "scenario_steps_per_s": K * H / dt, "seconds_with_PGSampleBatch": dt_batch,
           "batch_equals_list": bool(np.array_equal(out_b["raw"]["X"], out["raw"]["X"])), "kernel_ms": kms.value,
           "alloc_h2d_ms": tms.value, "kernel_scenario_steps_per_s": K * H / (kms.value * 1e-3),
           "kernel_GBps_algorithmic": (a_bytes + K * (N + n_x + (2 * H + 2) * n_x + 2 * H) * 8) / (kms.value * 1e-3) / 1e9,
           "kernel_fp64_tflops": 2.0 * K * (H + 1) * n_x * n_phi / (kms.value * 1e-3) / 1e12}
    try:
        import _oracle as O
        om = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L)
        t0 = time.perf_counter()
        ref = O.rollout(om, cpu_K, H, N, A[:cpu_K], Q[:cpu_K], w[:cpu_K], x[:cpu_K], um[:cpu_K], Cm, [[0.1]], u_init, seed=1)
        dtc = time.perf_counter() - t0
        res["cpu_oracle_scenario_steps_per_s"] = cpu_K * H / dtc
        res["bit_identical_to_oracle_on_first_%d" % cpu_K] = bool(np.array_equal(out["raw"]["X"][:cpu_K], ref["X"]))
    except Exception as e:  # oracle not built
        res["cpu_oracle"] = "unavailable: %s" % e
    print(json.dumps(res))


def chains(n_chains=8, N=2 ** 16, T=2000, sweeps=3):
    import bench
    prob = bench.build_problem("known", T, 0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=n_chains)
    eng.set_data(prob["u"], prob["y"]); eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
    eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0)); eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
    for c in range(n_chains):
        eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
    eng.set_rng_philox(7, 0)
    eng.sweep(1); eng.synchronize()
    t0 = time.perf_counter()
    eng.sweep(sweeps); eng.synchronize()
    dt = time.perf_counter() - t0
    st = eng.status(0)
    print(json.dumps({"metric": "chain-batched PG sweeps (filter + trace + statistics + MNIW)", "n_chains": n_chains, "N": N,
                      "T": T, "sweeps_per_s_all_chains": n_chains * sweeps / dt,
                      "particle_steps_per_s": n_chains * sweeps * N * (T - 1) / dt, "status": st}))
    eng.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("rollout", "all"):
        rollout(K=int(sys.argv[2]) if len(sys.argv) > 2 else 100000)
    if what in ("chains", "all"):
        chains()
