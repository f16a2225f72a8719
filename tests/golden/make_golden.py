"""Generates the fixtures in tests/golden/ from the CPU oracle (oracle/pg_oracle.c).

PARITY UNPINNED: the reference (Julia) ships no tests / golden vectors and cannot be executed in this image, so these
vectors are NOT outputs of the reference.  They freeze the oracle's defined semantics (summation order, elementary
functions, Philox stream addressing) on small seeded inputs so that (a) an accidental change of the oracle is caught on
the CPU (tests/test_golden.py, -m "not gpu") and (b) the CUDA path is checked against committed bytes on the GPU box
without trusting a freshly built oracle there (tests/test_golden.py, -m gpu).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import _oracle as O  # noqa: E402
import _problems as P  # noqa: E402

HG_L = [10.0, 20.0, 20.0]


def hilbert2():
    """n_x = 2, n_u = 1, 5x5x5 basis in the layout pgopt_b200.HilbertGPBasis produces (count/stride/L per dimension)."""
    from pgopt_b200.api import HilbertGPBasis
    hb = HilbertGPBasis([5, 5, 5], HG_L)
    return O.hilbert_model(2, 1, hb.count, hb.stride, hb.L)


def hilbert_prior_V():
    # NumPy restatement of Julia/examples/PG_OCP_generic_basis_functions_Altro.jl:29-60 lives in pgopt_b200.api; the
    # fixture stores V so the test does not depend on it
    from pgopt_b200.api import hilbert_gp_prior_V
    return hilbert_gp_prior_V([5, 5, 5], HG_L, 2 * np.pi * np.ones(3), 100 ** 2 / (8 * np.pi ** 4))


def resampling_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for N, n_draw, kind in [(7, 7, "uniform"), (30, 30, "exp"), (30, 1, "exp"), (1000, 1000, "peaked"), (4097, 4097, "sparse"),
                            (5000, 4999, "lognormal")]:
        if kind == "uniform":
            W = np.full(N, 1.0 / N)
        elif kind == "exp":
            W = rng.exponential(size=N)
        elif kind == "peaked":
            W = rng.exponential(size=N) * 1e-6
            W[rng.integers(N)] = 1.0
        elif kind == "sparse":
            W = np.where(rng.uniform(size=N) < 0.02, rng.exponential(size=N), 0.0)
        else:
            W = np.exp(3 * rng.normal(size=N))
        W = W / O.pairwise_sum(W)
        rnd = float(rng.uniform())
        idx, st = O.systematic_resampling(W, n_draw, rnd)
        assert st == 0
        cases.append((W, n_draw, rnd, idx))
    out = {"n_cases": np.array(len(cases))}
    for i, (W, n_draw, rnd, idx) in enumerate(cases):
        out["W%d" % i], out["n_draw%d" % i], out["rnd%d" % i], out["idx%d" % i] = W, np.array(n_draw), np.array(rnd), idx
    return out


def sweep_case(model_name, N, T, seed):
    if model_name == "known":
        om, A = O.known_model(), P.plausible_A_known()
    else:
        om = hilbert2()
        A = None
    u, y, xtrue = P.make_data(T, seed)
    if A is None:
        rng = np.random.default_rng(seed + 1)
        A = rng.normal(0, 0.02, (2, om.n_phi))
    Q = P.Q_TRUE if model_name == "known" else 0.3 * np.eye(2)
    Cm, R, x0m, x0c = np.array([[1.0, 0.0]]), 0.1, np.array([2.0, 2.0]), np.eye(2)
    st1 = O.make_streams(philox=True, seed=5, chain=1)
    first = O.sweep(om, N, T, u, y, Cm, R, x0m, x0c, A, Q, np.zeros((2, T)), True, st1, 1, hist=False)
    st2 = O.make_streams(philox=True, seed=5, chain=1)
    cond = O.sweep(om, N, T, u, y, Cm, R, x0m, x0c, A, Q, first["x_prim"], False, st2, 2, hist=True)
    assert first["status"] == 0 and cond["status"] == 0
    return dict(N=np.array(N), T=np.array(T), u=u, y=y, A=A, Q=Q, x_prim_in=first["x_prim"], a=cond["a"], x=cond["x"],
                w_last=cond["w_last"], star=np.array(cond["star"]), x_prim_out=cond["x_prim"], Phi=cond["Phi"],
                Psi=cond["Psi"], Sigma=cond["Sigma"])


def pg_case(model_name, N, T, K, K_b, k_d):
    if model_name == "known":
        om, V = O.known_model(), 10 * np.ones(5)
    else:
        om, V = hilbert2(), hilbert_prior_V()
    u, y, _ = P.make_data(T, 13)
    ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, om.n_phi)),
                           100 * np.eye(2), V, 100 * np.eye(2), 10.0, K, K_b, k_d, seed=99, chain=2)
    assert ref["status"] == 0
    return dict(N=np.array(N), T=np.array(T), K=np.array(K), K_b=np.array(K_b), k_d=np.array(k_d), u=u, y=y, V=V,
                A=ref["A"], Q=ref["Q"], w=ref["w"], x=ref["x"], x_prim=ref["x_prim"])


def mniw_case():
    rng = np.random.default_rng(31)
    n_x, n_phi, Tm1 = 2, 5, 399
    zeta, z = rng.normal(size=(n_x, Tm1)), rng.normal(size=(n_phi, Tm1))
    Phi, Psi, Sig = zeta @ zeta.T, zeta @ z.T, z @ z.T
    Sig = 0.5 * (Sig + Sig.T)
    V, Lam, ell = 10 * np.ones(n_phi), 100 * np.eye(n_x), 10.0
    s = P.injected_streams(4, 4, n_x, n_phi, 3, df=ell + Tm1 * n_x)
    A1, Q1, rc1 = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1,
                                O.make_streams(philox=False, bartlett=s["bartlett"], Xmn=s["Xmn"]), 1)
    A2, Q2, rc2 = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, O.make_streams(philox=True, seed=77, chain=3), 5)
    assert rc1 == 0 and rc2 == 0
    return dict(Phi=Phi, Psi=Psi, Sigma=Sig, V=V, Lam=Lam, ell=np.array(ell), Tm1=np.array(Tm1), bartlett=s["bartlett"],
                Xmn=s["Xmn"], A_injected=A1, Q_injected=Q1, A_philox=A2, Q_philox=Q2)


def rollout_case():
    rng = np.random.default_rng(41)
    from pgopt_b200.api import HilbertGPBasis
    count, L = [1, 3, 3, 3, 3], [4.0, 5.0, 6.0, 7.0, 8.0]  # constructor arguments (reference order)
    hb = HilbertGPBasis(count, L)
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    K, H, N, n_x, n_phi = 9, 41, 30, 4, om.n_phi
    A = rng.normal(0, 0.3, (K, n_x, n_phi))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(n_x))(rng.normal(0, 0.2, (n_x, n_x))) for _ in range(K)])
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, 1))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    u_init = rng.normal(0, 1, (1, H))
    ref = O.rollout(om, K, H, N, A, Q, w, x, um, Cm, [[0.1]], u_init, seed=17)
    assert ref["status"] == 0
    return dict(count=np.array(count), L=np.array(L), A=A, Q=Q, w=w, x=x, um=um, Cm=Cm, u_init=u_init, x0=ref["x0"],
                v=ref["v"], e=ref["e"], X=ref["X"], Y=ref["Y"])


def prediction_case():
    """test_prediction (particle_Gibbs.jl:253-314) on the models of the rollout fixture: K = 9 models x k_n = 2."""
    g = rollout_case()
    from pgopt_b200.api import HilbertGPBasis
    hb = HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    rng = np.random.default_rng(43)
    K, N, k_n, T_test = g["A"].shape[0], g["w"].shape[1], 2, 12
    u_test, y_test = rng.normal(0, 1, (1, T_test)), rng.normal(0, 1, (1, T_test))
    ref = O.test_prediction(om, K, k_n, T_test, N, g["A"], g["Q"], g["w"], g["x"], g["um"], g["Cm"], [[0.1]], u_test, y_test,
                            seed=19)
    assert ref["status"] == 0
    return dict(k_n=np.array(k_n), u_test=u_test, y_test=y_test, x_sim=ref["x_sim"], y_sim=ref["y_sim"],
                mean_rmse=np.array(ref["mean_rmse"]))


def dynamics_case():
    """One stacked dynamics step + Jacobians (optimal_control_Altro.jl:79-107, :51) on the rollout fixture's samples, with
    the rollout's own process noise: state t -> t+1 for both phi roundings, and a dense-V MNIW draw."""
    g = dict(np.load(os.path.join(HERE, "rollout.npz")))
    from pgopt_b200.api import HilbertGPBasis
    hb = HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    t = 4
    out = {"t": np.array(t)}
    for variant in (0, 1):
        xn, jx, ju = O.dynamics(om, g["A"], g["X"][:, t], g["u_init"][:, t], g["v"], t, variant)
        out["x_next_%d" % variant], out["dfdx_%d" % variant], out["dfdu_%d" % variant] = xn, jx, ju
    assert np.array_equal(out["x_next_0"], g["X"][:, t + 1])  # the step the rollout iterates
    # MNIW_sample with a dense prior covariance V (particle_Gibbs.jl:56, :64)
    rng = np.random.default_rng(77)
    n_x, n_phi, Tm1 = 2, 9, 120
    zeta, z = rng.normal(size=(n_x, Tm1)), rng.normal(size=(n_phi, Tm1))
    B = rng.normal(size=(n_phi, n_phi))
    V = B @ B.T + 0.5 * np.eye(n_phi)
    V = 0.5 * (V + V.T)
    Sig = z @ z.T
    Sig = 0.5 * (Sig + Sig.T)
    st = O.make_streams(philox=True, seed=13, chain=2)
    A, Q, rc = O.mniw_sample(n_x, n_phi, zeta @ zeta.T, zeta @ z.T, Sig, V, 100 * np.eye(n_x), 10.0, Tm1, st, 4)
    assert rc == 0
    out.update(mn_Phi=zeta @ zeta.T, mn_Psi=zeta @ z.T, mn_Sigma=Sig, mn_V=V, mn_Tm1=np.array(Tm1), mn_A=A, mn_Q=Q)
    return out


def main():
    save = lambda name, d: np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    save("resampling", resampling_cases())
    save("sweep_known", sweep_case("known", 30, 40, 3))
    save("sweep_hilbert", sweep_case("hilbert", 48, 16, 4))
    save("pg_known", pg_case("known", 30, 60, 3, 2, 1))
    save("pg_hilbert", pg_case("hilbert", 30, 30, 2, 1, 0))
    save("mniw", mniw_case())
    save("rollout", rollout_case())
    save("prediction", prediction_case())
    save("dynamics", dynamics_case())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
