"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).

PARITY UNPINNED: the fixtures are oracle outputs, not reference (Julia) outputs — the reference has no tests or golden
vectors and cannot run in this image.  CPU tests check that the oracle still reproduces them bit for bit; GPU tests
check the CUDA path (through the C ABI) against the committed bytes, without calling the oracle."""
import os

import numpy as np
import pytest

import _oracle as O
import pgopt_b200 as pg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CM, R, X0M = np.array([[1.0, 0.0]]), 0.1, np.array([2.0, 2.0])


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def same(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (name, a.shape, b.shape)
    if a.dtype.kind == "f":
        bad = (a.view(np.uint64) != b.view(np.uint64)) & ~((a == 0) & (b == 0))
    else:
        bad = a != b
    assert not bad.any(), "%s: %d of %d entries differ" % (name, bad.sum(), bad.size)


def models(name):
    if name == "known":
        return pg.KnownBasisVB(), O.known_model()
    hb = pg.HilbertGPBasis([5, 5, 5], [10, 20, 20])
    return hb, O.hilbert_model(2, 1, hb.count, hb.stride, hb.L)


# ------------------------------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_resampling_fixture():
    g = load("resampling")
    for i in range(int(g["n_cases"])):
        idx, st = O.systematic_resampling(g["W%d" % i], int(g["n_draw%d" % i]), float(g["rnd%d" % i]))
        assert st == 0
        same("idx%d" % i, idx, g["idx%d" % i])


@pytest.mark.parametrize("model", ["known", "hilbert"])
def test_oracle_reproduces_sweep_fixture(model):
    g = load("sweep_" + model)
    _, om = models(model)
    st = O.make_streams(philox=True, seed=5, chain=1)
    r = O.sweep(om, int(g["N"]), int(g["T"]), g["u"], g["y"], CM, R, X0M, np.eye(2), g["A"], g["Q"], g["x_prim_in"], False, st, 2)
    assert r["status"] == 0 and r["star"] == int(g["star"])
    for k, kk in (("a", "a"), ("x", "x"), ("w_last", "w_last"), ("x_prim", "x_prim_out"), ("Phi", "Phi"), ("Psi", "Psi"),
                  ("Sigma", "Sigma")):
        same(k, r[k][1:] if k == "a" else r[k], g[kk][1:] if k == "a" else g[kk])


@pytest.mark.parametrize("model", ["known", "hilbert"])
def test_oracle_reproduces_pg_fixture(model):
    g = load("pg_" + model)
    _, om = models(model)
    r = O.particle_gibbs(om, int(g["N"]), int(g["T"]), g["u"], g["y"], CM, R, X0M, np.eye(2), np.zeros((2, om.n_phi)),
                         100 * np.eye(2), g["V"], 100 * np.eye(2), 10.0, int(g["K"]), int(g["K_b"]), int(g["k_d"]), seed=99, chain=2)
    assert r["status"] == 0
    for k in ("A", "Q", "w", "x", "x_prim"):
        same(k, r[k], g[k])


def test_oracle_reproduces_mniw_and_rollout_fixtures():
    g = load("mniw")
    A, Q, rc = O.mniw_sample(2, 5, g["Phi"], g["Psi"], g["Sigma"], g["V"], g["Lam"], float(g["ell"]), int(g["Tm1"]),
                             O.make_streams(philox=False, bartlett=g["bartlett"], Xmn=g["Xmn"]), 1)
    same("A", A, g["A_injected"]); same("Q", Q, g["Q_injected"])
    A, Q, rc = O.mniw_sample(2, 5, g["Phi"], g["Psi"], g["Sigma"], g["V"], g["Lam"], float(g["ell"]), int(g["Tm1"]),
                             O.make_streams(philox=True, seed=77, chain=3), 5)
    same("A philox", A, g["A_philox"]); same("Q philox", Q, g["Q_philox"])
    g = load("rollout")
    hb = pg.HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    K, H = g["A"].shape[0], g["u_init"].shape[1]
    r = O.rollout(om, K, H, g["w"].shape[1], g["A"], g["Q"], g["w"], g["x"], g["um"], g["Cm"], [[0.1]], g["u_init"], seed=17)
    for k in ("x0", "v", "e", "X", "Y"):
        same(k, r[k], g[k])


def _prediction_inputs():
    g, gp = load("rollout"), load("prediction")
    hb = pg.HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    return g, gp, hb


def test_oracle_reproduces_prediction_fixture():
    g, gp, hb = _prediction_inputs()
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    K, T_test = g["A"].shape[0], gp["y_test"].shape[1]
    r = O.test_prediction(om, K, int(gp["k_n"]), T_test, g["w"].shape[1], g["A"], g["Q"], g["w"], g["x"], g["um"], g["Cm"],
                          [[0.1]], gp["u_test"], gp["y_test"], seed=19)
    same("x_sim", r["x_sim"], gp["x_sim"]); same("y_sim", r["y_sim"], gp["y_sim"])
    assert r["mean_rmse"] == float(gp["mean_rmse"])


def _dynamics_inputs():
    g, gd = load("rollout"), load("dynamics")
    hb = pg.HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    return g, gd, hb


def test_oracle_reproduces_dynamics_fixture():
    """stacked dynamics + Jacobians (optimal_control_Altro.jl:79-107, :51) and MNIW_sample with a dense V"""
    g, gd, hb = _dynamics_inputs()
    om = O.hilbert_model(4, 1, hb.count, hb.stride, hb.L)
    t = int(gd["t"])
    for variant in (0, 1):
        xn, jx, ju = O.dynamics(om, g["A"], g["X"][:, t], g["u_init"][:, t], g["v"], t, variant)
        same("x_next", xn, gd["x_next_%d" % variant]); same("dfdx", jx, gd["dfdx_%d" % variant])
        same("dfdu", ju, gd["dfdu_%d" % variant])
    same("x_next == X[t+1]", gd["x_next_0"], g["X"][:, t + 1])
    st = O.make_streams(philox=True, seed=13, chain=2)
    n_x, n_phi = gd["mn_Psi"].shape
    A, Q, rc = O.mniw_sample(n_x, n_phi, gd["mn_Phi"], gd["mn_Psi"], gd["mn_Sigma"], gd["mn_V"], 100 * np.eye(n_x), 10.0,
                             int(gd["mn_Tm1"]), st, 4)
    assert rc == 0
    same("A (dense V)", A, gd["mn_A"]); same("Q (dense V)", Q, gd["mn_Q"])


def test_oracle_dynamics_matches_numpy():
    """orc_dynamics against a NumPy statement of x_n[k] = A_k phi(x[k], u) + v[:, t+1, k] with
    phi_j(z) = prod_k L_k^-1/2 sin(pi j_k (z_k + L_k) / (2 L_k)) (examples/PG_OCP_generic_basis_functions_Altro.jl:204-217) and
    of its analytic Jacobians (what RobotDynamics.@autodiff derives, :51): values to 1e-12, and the Jacobians also against
    central differences of the NumPy function."""
    g, gd, hb = _dynamics_inputs()
    t = int(gd["t"])
    count, L, nz, n_x, n_u = hb.count, hb.L, len(hb.count), 4, 1
    K = g["A"].shape[0]
    digits = np.array(np.unravel_index(np.arange(hb.n_phi), count)).T      # last augmented dimension fastest

    def phi_and_grad(z):
        arg = [np.pi * np.arange(1, count[k] + 1) * (z[k] + L[k]) / (2 * L[k]) for k in range(nz)]
        tk = [np.sin(arg[k]) / np.sqrt(L[k]) for k in range(nz)]
        dk = [np.cos(arg[k]) * np.pi * np.arange(1, count[k] + 1) / (2 * L[k]) / np.sqrt(L[k]) for k in range(nz)]
        fac = np.stack([tk[k][digits[:, k]] for k in range(nz)], axis=1)
        phi = fac.prod(axis=1)
        grad = np.stack([np.prod(np.delete(fac, k, axis=1), axis=1) * dk[k][digits[:, k]] for k in range(nz)], axis=1)
        return phi, grad

    for k in range(K):
        x, u = g["X"][k, t], g["u_init"][:, t]
        z = np.concatenate([u, x])
        phi, grad = phi_and_grad(z)
        np.testing.assert_allclose(gd["x_next_0"][k], g["A"][k] @ phi + g["v"][k, t], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(gd["dfdu_0"][k], g["A"][k] @ grad[:, :n_u], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(gd["dfdx_0"][k], g["A"][k] @ grad[:, n_u:], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(gd["x_next_1"][k], gd["x_next_0"][k], rtol=1e-12, atol=1e-13)   # phi_opt rounding only
        eps = 1e-6
        for c in range(nz):
            zp, zm = z.copy(), z.copy()
            zp[c] += eps
            zm[c] -= eps
            fd = g["A"][k] @ (phi_and_grad(zp)[0] - phi_and_grad(zm)[0]) / (2 * eps)
            np.testing.assert_allclose((gd["dfdu_0"][k][:, c] if c < n_u else gd["dfdx_0"][k][:, c - n_u]), fd, rtol=0, atol=5e-8)


# ------------------------------------------------------------------------------------------ GPU: CUDA path vs fixtures
@pytest.mark.gpu
def test_gpu_reproduces_dynamics_fixture():
    g, gd, hb = _dynamics_inputs()
    K, t = g["A"].shape[0], int(gd["t"])
    batch = pg.PGSampleBatch(g["A"], g["Q"], g["w"], np.transpose(g["x"], (0, 2, 1)), g["um"])
    with pg.DeviceSamples.from_samples(batch, hb) as ds:
        sc = ds.rollout(pg.LinearMeasurement(g["Cm"]), 0.1, g["u_init"].shape[1], u_init=g["u_init"], seed=17).scenarios(want=("X", "v"))
        same("rollout X", sc["X"], g["X"])
        for variant in (0, 1):
            xn, jx, ju = ds.dynamics(g["X"][:, t].reshape(-1), g["u_init"][:, t], t=t, phi_variant=variant)
            same("x_next", xn.reshape(K, 4), gd["x_next_%d" % variant]); same("dfdx", jx, gd["dfdx_%d" % variant])
            same("dfdu", ju, gd["dfdu_%d" % variant])
    n_x, n_phi = gd["mn_Psi"].shape
    A, Q = pg.MNIW_sample(gd["mn_Phi"], gd["mn_Psi"], gd["mn_Sigma"], gd["mn_V"], 100 * np.eye(n_x), 10.0, int(gd["mn_Tm1"]),
                          seed=13, chain=2, sweep_index=4)
    same("A (dense V)", A, gd["mn_A"]); same("Q (dense V)", Q, gd["mn_Q"])



@pytest.mark.gpu
def test_gpu_reproduces_prediction_fixture():
    g, gp, hb = _prediction_inputs()
    K = g["A"].shape[0]
    samples = [pg.PG_sample(g["A"][k], g["Q"][k], g["w"][k], g["x"][k].T, g["um"][k]) for k in range(K)]
    out = pg.test_prediction(samples, hb, pg.LinearMeasurement(g["Cm"]), 0.1, int(gp["k_n"]), gp["u_test"], gp["y_test"],
                             seed=19)
    same("x_sim", out["raw"]["x_sim"], gp["x_sim"]); same("y_sim", out["raw"]["y_sim"], gp["y_sim"])
    assert out["mean_rmse"] == float(gp["mean_rmse"])


@pytest.mark.gpu
def test_gpu_reproduces_resampling_fixture():
    g = load("resampling")
    for i in range(int(g["n_cases"])):
        idx = pg.systematic_resampling(g["W%d" % i], int(g["n_draw%d" % i]), rnd=float(g["rnd%d" % i]))
        same("idx%d" % i, idx, g["idx%d" % i])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("model", ["known", "hilbert"])
def test_gpu_reproduces_sweep_fixture(model, mode):
    g = load("sweep_" + model)
    basis, _ = models(model)
    N, T = int(g["N"]), int(g["T"])
    with pg.ParticleGibbsEngine(basis, 1, N, T, keep_w_history=True) as e:
        e.set_mode(mode)
        e.set_data(g["u"], g["y"])
        e.set_measurement(pg.LinearMeasurement(CM), R)
        e.set_init_dist(pg.MvNormal(X0M, 1.0))
        e.set_state(g["A"], g["Q"], g["x_prim_in"], 2)
        e.set_rng_philox(5, 1)
        e.sweep_filter()
        a, x, _ = e.get_history()
        Phi, Psi, Sig, star = e.get_stats()
        _, _, xprim, _ = e.get_state()
        smp = e.get_sample()
        stat, nfb, _ = e.status()
    assert stat == 0 and nfb == 0 and star == int(g["star"])
    same("a", a[1:], g["a"][1:]); same("x", x, g["x"]); same("w_last", smp.w_m1, g["w_last"])
    same("x_prim", xprim, g["x_prim_out"]); same("Phi", Phi, g["Phi"]); same("Psi", Psi, g["Psi"]); same("Sigma", Sig, g["Sigma"])


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["known", "hilbert"])
def test_gpu_reproduces_pg_fixture(model):
    g = load("pg_" + model)
    basis, _ = models(model)
    T, K = int(g["T"]), int(g["K"])
    smp = pg.particle_Gibbs(g["u"], g["y"], K, int(g["K_b"]), int(g["k_d"]), int(g["N"]), basis, 100 * np.eye(2), 10.0,
                            100 * np.eye(2), g["V"], np.zeros((2, basis.n_phi)), pg.MvNormal(X0M, 1.0),
                            pg.LinearMeasurement(CM), R, x_prim=np.zeros((2, T)), seed=99, chain=2)
    assert len(smp) == K
    for k in range(K):
        same("A[%d]" % k, smp[k].A, g["A"][k]); same("Q[%d]" % k, smp[k].Q, g["Q"][k])
        same("w[%d]" % k, smp[k].w_m1, g["w"][k]); same("x[%d]" % k, smp[k].x_m1.T, g["x"][k])


@pytest.mark.gpu
def test_gpu_reproduces_mniw_and_rollout_fixtures():
    g = load("mniw")
    A, Q = pg.MNIW_sample(g["Phi"], g["Psi"], g["Sigma"], g["V"], g["Lam"], float(g["ell"]), int(g["Tm1"]),
                          bartlett=g["bartlett"], Xmn=g["Xmn"])
    same("A", A, g["A_injected"]); same("Q", Q, g["Q_injected"])
    A, Q = pg.MNIW_sample(g["Phi"], g["Psi"], g["Sigma"], g["V"], g["Lam"], float(g["ell"]), int(g["Tm1"]), seed=77, chain=3,
                          sweep_index=5)
    same("A philox", A, g["A_philox"]); same("Q philox", Q, g["Q_philox"])
    g = load("rollout")
    basis = pg.HilbertGPBasis([int(c) for c in g["count"]], [float(v) for v in g["L"]])
    K, H = g["A"].shape[0], g["u_init"].shape[1]
    samples = [pg.PG_sample(g["A"][k], g["Q"][k], g["w"][k], g["x"][k].T, g["um"][k]) for k in range(K)]
    out = pg.scenario_rollout(samples, basis, pg.LinearMeasurement(g["Cm"]), 0.1, H, u_init=g["u_init"], seed=17)
    for k in ("x0", "v", "e", "X", "Y"):
        same(k, out["raw"][k], g[k])
