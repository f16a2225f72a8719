"""End-to-end run of the reference's example Julia/examples/PG_OCP_known_basis_functions_Altro.jl (paper Sec. V-B,
BASELINE.json configs[0]) up to the point where the OCP solver takes over: data generation (:58-98), priors (:36-50),
`particle_Gibbs(u_training, y_training, K, K_b, k_d, N, phi, ...)` (:110) and `test_prediction(PG_samples, phi, g, R,
10, u_test, y_test)` (:115) — both on the GPU through the Python mirror of the reference API.  Prints one JSON line:
wall times, PG sweeps/s, the mean RMSE of the forward simulations against the test output (the number the
reference prints, particle_Gibbs.jl:311-312) next to the standard deviation of the test output.

    python scripts/example_end_to_end.py [K] [K_b] [k_d] [basis]      (defaults 100, 1000, 50, known: K_total = 6050)

basis = generic runs the same data through the setup of examples/PG_OCP_generic_basis_functions_Altro.jl (paper
Sec. V-C, BASELINE.json configs[1]): the reduced-rank GP basis with 5 x 5 x 5 functions on L = [10, 20, 20] (:34-48) and
the squared-exponential prior V with l = 2 pi, s_f = 100^2 / (8 pi^4) (:111-115).
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pgopt_b200 as pg  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    K_b = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    k_d = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    N, T, T_test, n_x, n_phi = 30, 2000, 500, 2, 5
    R = 0.1
    rng = np.random.default_rng(82)  # the example seeds Julia's RNG with 82 (:14); streams differ, the setup does not

    def f_true(x, u):  # :69
        return np.array([0.8 * x[0] - 0.5 * x[1] + 0.1 * np.cos(3 * x[0]) * x[1],
                         0.4 * x[0] + 0.5 * x[1] + (1 + 0.3 * np.sin(2 * x[1])) * u])

    Q_true = np.array([[0.03, -0.004], [-0.004, 0.01]])  # :70
    Lq = np.linalg.cholesky(Q_true)
    T_all = T + T_test
    u = np.concatenate([rng.normal(0, 3, T), 3 * np.sin(2 * np.pi / T_test * np.arange(T_test))])[None, :]  # :77-80
    x = np.empty((n_x, T_all + 1))
    y = np.empty((1, T_all))
    x[:, 0] = np.array([2.0, 2.0]) + rng.normal(size=2)  # :84
    for t in range(T_all):  # :86-89; MvNormal(zeros(n_y), R::Real) has standard deviation R (Distributions.jl)
        x[:, t + 1] = f_true(x[:, t], u[0, t]) + Lq @ rng.normal(size=2)
        y[0, t] = x[0, t] + R * rng.normal()
    u_tr, y_tr, u_te, y_te = u[:, :T], y[:, :T], u[:, T:], y[:, T:]

    which = sys.argv[4] if len(sys.argv) > 4 else "known"
    g = pg.LinearMeasurement([[1.0, 0.0]])
    if which == "generic":
        basis = pg.HilbertGPBasis([5, 5, 5], [10.0, 20.0, 20.0])
        n_phi = basis.n_phi
        V = pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3), 100 ** 2 / (8 * np.pi ** 4))
    else:
        basis, V = pg.KnownBasisVB(), 10 * np.ones(n_phi)
    K_total = K_b + 1 + (K - 1) * (k_d + 1)  # particle_Gibbs.jl:116
    t0 = time.perf_counter()
    samples = pg.particle_Gibbs(u_tr, y_tr, K, K_b, k_d, N, basis, 100 * np.eye(n_x), 10.0, 100 * np.eye(n_x),
                                V, np.zeros((n_x, n_phi)), pg.MvNormal([2.0, 2.0], 1.0), g, R, seed=82)
    t_pg = time.perf_counter() - t0
    t0 = time.perf_counter()
    tp = pg.test_prediction(samples, basis, g, R, 10, u_te, y_te, seed=83)
    t_tp = time.perf_counter() - t0
    print(json.dumps({
        "example": "PG_OCP_%s_basis_functions (paper Sec. V-%s): N=30, T=2000, n_phi=%d, K=%d, K_b=%d, k_d=%d -> K_total=%d sweeps"
                   % (which, "C" if which == "generic" else "B", n_phi, K, K_b, k_d, K_total),
        "particle_gibbs_seconds": t_pg, "pg_sweeps_per_s": K_total / t_pg,
        "test_prediction_seconds": t_tp, "test_prediction_simulations": K * 10, "T_test": T_test,
        "mean_rmse": tp["mean_rmse"], "std_of_y_test": float(np.std(y_te))}))


if __name__ == "__main__":
    main()
