"""An INDEPENDENT whole-sweep check of the oracle.

The oracle and the kernels share pgb_math.h, so GPU == oracle on exp / sin / log is same-code agreement.  This file
transcribes particle_Gibbs.jl:154-230 a second time in NumPy — platform libm (np.exp / np.sin / np.cos), NumPy's own
matrix products and sums, a plain Python loop for `q = q + W[n]` (:23-32) — shares nothing with oracle/pg_oracle.c or
pgb_math.h, and is fed the same injected streams.  Bar: every ancestor / trajectory index equal, floats within 1e-10
relative (the bar north_star sets against the real Julia run).

Also here: the loader for REFERENCE-HELD golden vectors.  `julia/make_fixtures.jl`, run by anyone with Julia 1.10 and the
reference checked out, writes tests/golden/ref_*.npz from the unmodified reference fed pre-drawn streams; when such a
file is present the oracle (and, on a GPU box, the CUDA path) is compared with it — the route from "parity unpinned"
to pinned.  No such file can be produced in this image (no Julia), so that test skips itself, saying so.
"""
import numpy as np
import pytest

import _oracle as O
import _problems as P


# ----------------------------------------------------------------------------------------------------------------
# literal NumPy transcription of the reference
# ----------------------------------------------------------------------------------------------------------------
def np_systematic_resampling(W, N, rnd):
    """particle_Gibbs.jl:17-34, `rnd` standing for rand() at :21.  Returns 1-based indices."""
    W = W / np.sum(W)
    u = 1 / N * rnd
    idx = np.zeros(N, dtype=np.int64)
    q = 0.0
    n = 0
    for i in range(N):
        while q < u:
            n = n + 1
            q = q + W[n - 1]
        idx[i] = n
        u = u + 1 / N
    return idx


def np_phi_known(x, u):
    """examples/PG_OCP_known_basis_functions_Altro.jl:34, x (2, n), u (1, n)"""
    return np.vstack([0.1 * x[0], 0.1 * x[1], u[0], 0.01 * np.cos(3 * x[0]) * x[1], 0.1 * np.sin(2 * x[1]) * u[0]])


def np_phi_hilbert(n_phi_dims, L):
    """phi_sampling of examples/PG_OCP_generic_basis_functions_Altro.jl:52-99 as a closure over the index table"""
    n_phi_dims = np.asarray(n_phi_dims)
    n_z, n_phi = len(n_phi_dims), int(np.prod(n_phi_dims))
    j_vec = np.zeros((n_phi, n_z))
    variants = np.concatenate([[1], np.cumprod(n_phi_dims[:-1])])
    for i in range(n_phi):
        rem = i
        for j in range(n_z - 1, -1, -1):
            j_vec[i, j] = rem // variants[j] + 1
            rem = rem % variants[j]
    j_vec = j_vec[:, ::-1]                                   # :82
    L = np.asarray(L, dtype=float)
    L_sqrt_inv = 1 / np.sqrt(L)                              # :85
    pi_j_over_2L = np.pi * j_vec / (2 * L)                   # :86

    def phi(x, u):
        z = np.vstack([u, x])                                # :91
        out = np.ones((n_phi, x.shape[1]))
        for k in range(n_z):                                 # :93-96
            out = out * (L_sqrt_inv[k] * np.sin(pi_j_over_2L[:, [k]] * (z[[k]] + L[k])))
        return out
    return phi


def np_sweep(phi, A, Q, C, R, u_tr, y_tr, N, x_prim, first, x0_mean, x0_chol, Z0, Ures, Z, Uanc, Ustar):
    """One iteration of the k loop, particle_Gibbs.jl:154-225, with the random draws replaced by the injected streams
    in the reference's order of consumption (rand(mvn) = mean + chol(Q).L * z)."""
    n_x, T = x_prim.shape
    n_y = y_tr.shape[0]
    Lq = np.linalg.cholesky(Q)
    f = lambda x, u: A @ phi(x, u)                           # :156
    w = np.zeros((T, N))
    x_pf = np.zeros((n_x, N, T))
    a = np.zeros((T, N), dtype=np.int64)
    x_pf[:, N - 1, :] = x_prim                               # :160
    for n in range(N - 1):                                   # :163-165
        x_pf[:, n, 0] = x0_mean + x0_chol @ Z0[n]
    c0 = -(n_x * np.log(2 * np.pi) + np.log(np.linalg.det(Q))) / 2
    for t in range(T):
        if t >= 1:
            ut = np.repeat(u_tr[:, [t - 1]], N, axis=1)
            if not first:
                idx = np_systematic_resampling(w[t - 1], N - 1, Ures[t])                               # :172
                a[t, :N - 1] = idx
                x_pf[:, :N - 1, t] = f(x_pf[:, idx - 1, t - 1], ut[:, :N - 1]) + Lq @ Z[t, :N - 1].T      # :175
                d = Lq_solve(Lq, f(x_pf[:, :, t - 1], ut) - x_pf[:, [N - 1], t])                      # :178-179
                waN = w[t - 1] * np.exp(c0 - 0.5 * np.sum(d * d, axis=0))
                waN = waN / np.sum(waN)                                                               # :180
                a[t, N - 1] = np_systematic_resampling(waN, 1, Uanc[t])[0]                             # :181
            else:
                idx = np_systematic_resampling(w[t - 1], N, Ures[t])                                   # :185
                a[t] = idx
                x_pf[:, :, t] = f(x_pf[:, idx - 1, t - 1], ut) + Lq @ Z[t].T                           # :188
        g = C @ x_pf[:, :, t]
        if n_y == 1:
            log_w = -(g[0] - y_tr[0, t]) ** 2 / 2 / R                                                  # :194
        else:
            r = y_tr[:, [t]] - g
            log_w = -np.sum(r * np.linalg.solve(R, r), axis=0) / 2                                     # :196
        w[t] = np.exp(log_w - np.max(log_w))                                                           # :198
        w[t] = w[t] / np.sum(w[t])                                                                     # :199
    star = np_systematic_resampling(w[T - 1], 1, Ustar)[0]                                             # :213
    star0 = star
    xp = np.zeros((n_x, T))
    xp[:, T - 1] = x_pf[:, star - 1, T - 1]
    for t in range(T - 2, -1, -1):                                                                     # :215-218
        star = a[t + 1, star - 1]
        xp[:, t] = x_pf[:, star - 1, t]
    zeta = xp[:, 1:]                                                                                   # :221
    z = phi(xp[:, :T - 1], u_tr[:, :T - 1])                                                            # :222
    return dict(a=a, x=x_pf, w=w, star=star0, x_prim=xp, Phi=zeta @ zeta.T, Psi=zeta @ z.T, Sigma=z @ z.T)


def Lq_solve(L, B):
    import scipy.linalg
    return scipy.linalg.solve_triangular(L, B, lower=True)


CASES = [("known", 30, 200, False), ("known", 30, 200, True), ("hilbert", 30, 200, False), ("hilbert", 64, 40, True),
         ("known", 257, 30, False)]


@pytest.mark.parametrize("model,N,T,first", CASES)
def test_oracle_sweep_equals_independent_numpy_transcription(model, N, T, first):
    if model == "known":
        om, phi, A, n_phi = O.known_model(), np_phi_known, P.plausible_A_known(), 5
    else:
        dims, L = [5, 5, 5], [10.0, 20.0, 20.0]
        nz = 3
        count = [dims[nz - 1 - k] for k in range(nz)]
        stride = [int(np.prod(dims[:nz - 1 - k])) for k in range(nz)]
        om, phi, n_phi = O.hilbert_model(2, 1, count, stride, L), np_phi_hilbert(dims, L), 125
        A = np.random.default_rng(9).normal(0, 2.0, (2, 125))
    u, y, _ = P.make_data(T, 41)
    s = P.injected_streams(N, T, 2, n_phi, 43)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    Cm, R = np.array([[1.0, 0.0]]), 0.1
    xp = np.zeros((2, T))
    if not first:
        xp = O.sweep(om, N, T, u, y, Cm, R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, True, st, 1)["x_prim"]
    ref = O.sweep(om, N, T, u, y, Cm, R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, first, st, 1 if first else 2)
    ind = np_sweep(phi, A, P.Q_TRUE, Cm, R, u, y, N, xp.copy(), first, np.array([2.0, 2.0]), np.eye(2), s["Z0"], s["Ures"],
                   s["Z"], s["Uanc"], s["Ustar"])
    assert ref["status"] == 0
    assert np.array_equal(ref["a"][1:], ind["a"][1:]), "%d ancestor indices differ" % np.sum(ref["a"][1:] != ind["a"][1:])
    assert ref["star"] == ind["star"]
    np.testing.assert_allclose(ref["x"], np.transpose(ind["x"], (2, 1, 0)), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ref["w"], ind["w"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(ref["x_prim"], ind["x_prim"], rtol=1e-10, atol=1e-12)
    for k in ("Phi", "Psi", "Sigma"):
        scale = np.abs(ind[k]).max()
        np.testing.assert_allclose(ref[k], ind[k], rtol=1e-10, atol=1e-12 * scale)


# ----------------------------------------------------------------------------------------------------------------
# reference-held golden vectors (written by julia/make_fixtures.jl; absent in this image)
# ----------------------------------------------------------------------------------------------------------------
import _reffix as RF  # noqa: E402


@pytest.mark.skipif(not RF.fixtures(), reason=RF.SKIP_REASON)
@pytest.mark.parametrize("path", RF.fixtures() or ["-"])
def test_oracle_against_reference_held_vectors(path):
    """The oracle fed the streams the unmodified Julia reference consumed: the final particles x_pf[:,:,T] of every
    sweep (a function of every resampling index) and the traced trajectory must agree within 1e-10 relative — which
    they only can when every ancestor index is the same."""
    g = RF.load(path)
    om, _ = RF.oracle_model(g)
    N, T, K = int(RF.scalar(g, "N")), int(RF.scalar(g, "T")), int(RF.scalar(g, "K"))
    xp = np.zeros((int(RF.scalar(g, "n_x")), T))
    for k in range(1, K + 1):
        s, A, Q = RF.sweep_inputs(g, k)
        st = O.make_streams(philox=False, **s)
        r = O.sweep(om, N, T, g["u"], g["y"], g["C"], RF.scalar(g, "R"), np.ravel(g["x0_mean"]), g["x0_chol"], A, Q, xp,
                    k == 1, st, k)
        assert r["status"] == 0
        np.testing.assert_allclose(r["w_last"], np.ravel(g["w_m1_%d" % k]), rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(r["x_last"].T, g["x_m1_%d" % k], rtol=1e-10, atol=1e-12)
        xp = r["x_prim"]
    np.testing.assert_allclose(xp, g["x_prim_final"], rtol=1e-10, atol=1e-12)
