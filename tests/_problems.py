"""Synthetic problems shaped like the reference examples (test + bench infrastructure)."""
import numpy as np

Q_TRUE = np.array([[0.03, -0.004], [-0.004, 0.01]])


def f_true(x, u):
    # Julia/examples/PG_OCP_known_basis_functions_Altro.jl:69
    return np.array([0.8 * x[0] - 0.5 * x[1] + 0.1 * np.cos(3 * x[0]) * x[1],
                     0.4 * x[0] + 0.5 * x[1] + (1 + 0.3 * np.sin(2 * x[1])) * u])


def make_data(T, seed=0):
    """u ~ N(0, 3^2), x0 ~ N([2,2], I), y = x1 + e  (known example :77-98; e has std R = 0.1 there)."""
    rng = np.random.default_rng(seed)
    Lq = np.linalg.cholesky(Q_TRUE)
    u = rng.normal(0, 3, (1, T))
    x = np.zeros((2, T + 1))
    x[:, 0] = rng.normal(2, 1, 2)
    y = np.zeros((1, T))
    for t in range(T):
        x[:, t + 1] = f_true(x[:, t], u[0, t]) + Lq @ rng.normal(size=2)
        y[0, t] = x[0, t] + 0.1 * rng.normal()
    return u, y, x


def plausible_A_known():
    # f_true expressed in the scaled known basis (known example :34,:69)
    return np.array([[8.0, -5.0, 0.0, 10.0, 0.0], [4.0, 5.0, 1.0, 0.0, 3.0]])


def injected_streams(N, T, n_x, n_phi, seed, df=None):
    rng = np.random.default_rng(seed)
    s = dict(Z0=rng.normal(size=(N - 1, n_x)), Ures=rng.uniform(size=T), Z=rng.normal(size=(T, N, n_x)),
             Uanc=rng.uniform(size=T), Ustar=float(rng.uniform()))
    B = np.tril(rng.normal(size=(n_x, n_x)))
    dfv = df if df is not None else 100.0
    for i in range(n_x):
        B[i, i] = np.sqrt(rng.chisquare(dfv - i))
    s["bartlett"] = B
    s["Xmn"] = rng.normal(size=(n_x, n_phi))
    return s
