"""GPU parity of the supplied-scenario rollout (the x_vec_0 / v_vec / e_vec keyword arguments of solve_PG_OCP_Ipopt,
Julia/src/optimal_control_Ipopt.jl:37,48,67,77 and the pre-solve re-entry :88-125) and of the device-resident pipeline
sampler -> DeviceSamples -> rollout -> screening (the reference reads PG_samples[k].A in place, :47-64): every output
bit-identical to the oracle / to the host-buffer path."""
import numpy as np
import pytest

import _oracle as O
import _problems as P
import pgopt_b200 as pg
from pgopt_b200 import _lib
from test_gpu_parity import _assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["warp", "cta"])
def rollout_kernel(request):
    _lib.lib().pgb_test_rollout_kernel(1 if request.param == "warp" else 0)
    yield request.param
    _lib.lib().pgb_test_rollout_kernel(-1)


def _scenario_problem(seed, n_x=4, dims=(1, 5, 5, 5, 5), N=30, K=13, H=21, n_y=1):
    rng = np.random.default_rng(seed)
    L = [4.0 + k for k in range(len(dims))]
    basis = pg.HilbertGPBasis(list(dims), L)
    om = O.hilbert_model(n_x, len(dims) - n_x, basis.count, basis.stride, basis.L, n_y=n_y)
    n_u = len(dims) - n_x
    A = rng.normal(0, 0.2, (K, n_x, basis.n_phi))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(n_x))(rng.normal(0, 0.2, (n_x, n_x))) for _ in range(K)])
    w = rng.uniform(size=(K, N)) ** 2
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, n_u))
    Cm = rng.normal(0, 1, (n_y, n_x))
    u_init = rng.normal(0, 1, (n_u, H))
    return dict(basis=basis, om=om, A=A, Q=Q, w=w, x=x, um=um, Cm=Cm, u_init=u_init, K=K, H=H, N=N, n_x=n_x, n_y=n_y,
                n_u=n_u, rng=rng)


@pytest.mark.parametrize("which", [("x0", "v", "e"), ("v",), ("x0",), ("e",), ("v", "e")])
def test_rollout_supplied_scenarios_bit_exact(which, rollout_kernel):
    """A supplied x_vec_0 / v_vec / e_vec replaces the draw and nothing else (optimal_control_Ipopt.jl:48,67,77)."""
    p = _scenario_problem(3)
    K, H, n_x, n_y = p["K"], p["H"], p["n_x"], p["n_y"]
    rng = p["rng"]
    sup = dict(x0=rng.normal(0, 1, (K, n_x)), v=rng.normal(0, 0.1, (K, H, n_x)), e=rng.normal(0, 0.1, (K, H, n_y)))
    kw = {k + "_in": (sup[k] if k in which else None) for k in ("x0", "v", "e")}
    ref = O.rollout_supplied(p["om"], K, H, p["N"], p["A"], p["Q"], p["w"], p["x"], p["um"], p["Cm"], [[0.1]],
                             p["u_init"], seed=9, k_offset=2, **kw)
    batch = pg.PGSampleBatch(p["A"], p["Q"], p["w"], np.transpose(p["x"], (0, 2, 1)), p["um"])
    out = pg.scenario_rollout(
        batch, p["basis"], pg.LinearMeasurement(p["Cm"]), 0.1, H, u_init=p["u_init"], seed=9, k_offset=2,
        x_vec_0=sup["x0"].reshape(-1) if "x0" in which else None,
        v_vec=np.transpose(sup["v"], (2, 1, 0)) if "v" in which else None,
        e_vec=np.transpose(sup["e"], (2, 1, 0)) if "e" in which else None)
    assert ref["status"] == 0
    for name in ("x0", "v", "e", "X", "Y"):
        _assert_same(name, out["raw"][name], ref[name])
    for name in which:  # supplied arrays come back unchanged
        _assert_same("echo " + name, out["raw"][name], sup[name])


def test_pre_solve_reentry_same_noise_new_input(rollout_kernel):
    """optimal_control_Ipopt.jl:88-125: the scenarios of the first entry, a new u_init, the rollout again — through the
    host arrays and through the arrays kept on the device (reuse), both equal to the oracle's second entry."""
    p = _scenario_problem(4, K=9, H=17)
    K, H = p["K"], p["H"]
    batch = pg.PGSampleBatch(p["A"], p["Q"], p["w"], np.transpose(p["x"], (0, 2, 1)), p["um"])
    g = pg.LinearMeasurement(p["Cm"])
    first = pg.scenario_rollout(batch, p["basis"], g, 0.1, H, u_init=p["u_init"], seed=5)
    u2 = p["rng"].normal(0, 1, p["u_init"].shape)
    ref1 = O.rollout(p["om"], K, H, p["N"], p["A"], p["Q"], p["w"], p["x"], p["um"], p["Cm"], [[0.1]], p["u_init"], seed=5)
    ref2 = O.rollout_supplied(p["om"], K, H, p["N"], p["A"], p["Q"], p["w"], p["x"], p["um"], p["Cm"], [[0.1]], u2, seed=5,
                              x0_in=ref1["x0"], v_in=ref1["v"], e_in=ref1["e"])
    second = pg.scenario_rollout(batch, p["basis"], g, 0.1, H, u_init=u2, seed=5, x_vec_0=first["x_vec_0"],
                                 v_vec=first["v_vec"], e_vec=first["e_vec"])
    for name in ("x0", "v", "e", "X", "Y"):
        _assert_same("host re-entry " + name, second["raw"][name], ref2[name])
    with pg.DeviceSamples.from_samples(batch, p["basis"], n_y=p["n_y"]) as ds:
        a = ds.rollout(g, 0.1, H, u_init=p["u_init"], seed=5).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("device first " + name, a[name], ref1[name])
        b = ds.rollout(g, 0.1, H, u_init=u2, seed=5, reuse=("x_vec_0", "v_vec", "e_vec")).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("device re-entry " + name, b[name], ref2[name])
        # host-supplied arrays into the resident set
        c = ds.rollout(g, 0.1, H, u_init=u2, seed=5, x_vec_0=first["x_vec_0"], v_vec=first["v_vec"],
                       e_vec=first["e_vec"]).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("device supplied " + name, c[name], ref2[name])


@pytest.mark.parametrize("n_y", [1, 2])
def test_device_pipeline_rollout_and_screening(n_y, rollout_kernel):
    """rollout + constraint screening on device-resident samples == oracle (and == the host-buffer entry points)."""
    p = _scenario_problem(6, n_x=2, dims=(5, 5, 5), N=100, K=41, H=30, n_y=n_y)
    K, H = p["K"], p["H"]
    Br = p["rng"].normal(0, 0.2, (n_y, n_y))
    R = Br @ Br.T + 0.05 * np.eye(n_y) if n_y > 1 else 0.1
    Le = np.linalg.cholesky(R) if n_y > 1 else [[0.1]]
    ref = O.rollout(p["om"], K, H, p["N"], p["A"], p["Q"], p["w"], p["x"], p["um"], p["Cm"], Le, p["u_init"], seed=8, k_offset=3)
    y_min = np.full((n_y, H), -np.inf)
    y_max = np.full((n_y, H), np.inf)
    y_min[0, 3:20] = -0.5
    y_max[n_y - 1, 10:] = 0.7
    sref = O.scenario_constraint_max(ref["Y"], y_min, y_max)
    batch = pg.PGSampleBatch(p["A"], p["Q"], p["w"], np.transpose(p["x"], (0, 2, 1)), p["um"])
    with pg.DeviceSamples.from_samples(batch, p["basis"], n_y=n_y) as ds:
        back = ds.to_batch()
        for name in ("A", "Q", "w_m1", "x_m1", "u_m1"):
            _assert_same("round trip " + name, getattr(back, name), getattr(batch, name))
        out = ds.rollout(pg.LinearMeasurement(p["Cm"]), R, H, u_init=p["u_init"], seed=8, k_offset=3).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same(name, out[name], ref[name])
        h, order = ds.constraint_max(y_min, y_max)
        _assert_same("h_max", h, sref["h_max"])
        assert np.array_equal(order, sref["order"])
        ms = ds.last_ms()
        assert ms["rollout_kernel_ms"] > 0 and ms["screening_kernel_ms"] > 0
        # x_T drawn where the sample lives: the same star the rollout draws, and the rollout that uses it
        xT = ds.draw_xT(seed=8, k_offset=3)
        for k in range(K):
            _assert_same("x_T[%d]" % k, xT[k], O.draw_xT(p["w"][k], p["x"][k], 8, 3 + k))
        out2 = ds.rollout(pg.LinearMeasurement(p["Cm"]), R, H, u_init=p["u_init"], seed=8, k_offset=3).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("with x_T " + name, out2[name], ref[name])


@pytest.mark.parametrize("model,N,T,K,K_b,k_d,nc", [("known", 30, 80, 3, 2, 1, 2), ("hilbert", 200, 30, 2, 1, 0, 1),
                                                     ("known", 3000, 12, 2, 0, 1, 3)])
def test_particle_gibbs_device_resident_samples(model, N, T, K, K_b, k_d, nc):
    """pgb_particle_gibbs_dev leaves exactly the samples pgb_particle_gibbs returns (and the oracle computes), and the
    rollout on them equals the rollout on the host copies."""
    from test_gpu_parity import _models
    basis, om, _ = _models()[model]
    n_phi = basis.n_phi
    u, y, _ = P.make_data(T, 51)
    V = 10 * np.ones(n_phi) if model == "known" else pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3),
                                                                           100 ** 2 / (8 * np.pi ** 4))
    A0, Q0 = np.zeros((2, n_phi)), 100 * np.eye(2)

    def engine():
        e = pg.ParticleGibbsEngine(basis, 1, N, T, n_chains=nc)
        e.set_data(u, y)
        e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        e.set_prior(V, 100 * np.eye(2), 10.0)
        for c in range(nc):
            e.set_state(A0, Q0, None, 1, chain=c)
        e.set_rng_philox(13, 4)
        return e

    with engine() as e:
        A_s, Q_s, w_s, x_s, u_m1 = e.particle_gibbs(K, K_b, k_d)
    with engine() as e:
        ds = e.particle_gibbs_dev(K, K_b, k_d)
    try:
        b = ds.to_batch()
        assert len(b) == nc * K
        for c in range(nc):
            for k in range(K):
                s = c * K + k
                _assert_same("A", b.A[s], A_s[c, k].reshape((2, n_phi), order="F"))
                _assert_same("Q", b.Q[s], Q_s[c, k].reshape((2, 2), order="F"))
                _assert_same("w", b.w_m1[s], w_s[c, k])
                _assert_same("x", b.x_m1[s], x_s[c, k].reshape((2, N), order="F"))
                _assert_same("u", b.u_m1[s], u_m1)
        ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A0, Q0, V, 100 * np.eye(2), 10.0,
                               K, K_b, k_d, seed=13, chain=4)
        for k in range(K):
            _assert_same("oracle A", b.A[k], ref["A"][k])
            _assert_same("oracle w", b.w_m1[k], ref["w"][k])
        H = 12
        u_init = np.cos(np.arange(H))[None, :]
        g = pg.LinearMeasurement([[1.0, 0.0]])
        host = pg.scenario_rollout(b, basis, g, 0.1, H, u_init=u_init, seed=2)["raw"]
        dev = ds.rollout(g, 0.1, H, u_init=u_init, seed=2).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same(name, dev[name], host[name])
    finally:
        ds.close()


# ------------------------------------------------------------------------------------------------------------------
# serialisation / checkpoint
# ------------------------------------------------------------------------------------------------------------------
def test_samples_pack_unpack_round_trip():
    p = _scenario_problem(7, n_x=2, dims=(5, 5, 5), N=50, K=9, H=11)
    batch = pg.PGSampleBatch(p["A"], p["Q"], p["w"], np.transpose(p["x"], (0, 2, 1)), p["um"])
    g = pg.LinearMeasurement(p["Cm"])
    with pg.DeviceSamples.from_samples(batch, p["basis"]) as ds:
        ref = ds.rollout(g, 0.1, p["H"], u_init=p["u_init"], seed=3).scenarios()
        full = ds.pack(with_particles=True)
        xT = ds.draw_xT(seed=3)
        slim = ds.pack(with_particles=False)
        rec = ds.records()
    assert bytes(full[:7]) == b"PGB200S" and int(np.frombuffer(full[8:12].tobytes(), dtype=np.uint32)[0]) == 1
    R = 2 * 125 + 4 + 2 + 1
    assert slim.size == 64 + 9 * R * 8 and full.size == 64 + 9 * (R + 50 + 100) * 8
    for k in range(9):  # record = [A column-major | Q | x_T | u_m1]
        _assert_same("A", rec[k, :250], p["A"][k].ravel(order="F"))
        _assert_same("Q", rec[k, 250:254], p["Q"][k].ravel(order="F"))
        _assert_same("x_T", rec[k, 254:256], xT[k])
        _assert_same("u_m1", rec[k, 256:], p["um"][k])
    # full layout -> a fresh set -> identical rollout (star re-drawn from w_m1 / x_m1)
    with pg.DeviceSamples(p["basis"], 1, 50, 9) as d2:
        d2.unpack(full)
        b2 = d2.to_batch()
        _assert_same("w", b2.w_m1, batch.w_m1)
        _assert_same("x", b2.x_m1, batch.x_m1)
        out = d2.rollout(g, 0.1, p["H"], u_init=p["u_init"], seed=3).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("full " + name, out[name], ref[name])
    # records only (what the NCCL gather moves): the rollout runs from x_T, no particle arrays needed
    with pg.DeviceSamples(p["basis"], 1, 50, 12) as d3:
        d3.unpack(slim)
        out = d3.rollout(g, 0.1, p["H"], u_init=p["u_init"], seed=3).scenarios()
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("slim " + name, out[name][:9], ref[name])
        bad = slim.copy()
        bad[0] = ord("X")
        with pytest.raises(pg.PgoptError):
            d3.unpack(bad)
        with pytest.raises(pg.PgoptError):
            d3.unpack(slim[:100])
    # records packed BEFORE x_T was drawn and without particle blocks cannot be rolled out: refused, not silently zero
    with pg.DeviceSamples.from_samples(batch, p["basis"]) as d4:
        bare = d4.pack(with_particles=False)
    with pg.DeviceSamples(p["basis"], 1, 50, 9) as d5:
        d5.unpack(bare)
        with pytest.raises(pg.PgoptError):
            d5.rollout(g, 0.1, p["H"], u_init=p["u_init"], seed=3)
        with pytest.raises(pg.PgoptError):
            d5.draw_xT(seed=3)
        x0 = np.zeros(9 * 2)
        d5.rollout(g, 0.1, p["H"], u_init=p["u_init"], seed=3, x_vec_0=x0)  # a supplied x_vec_0 needs neither


def test_checkpoint_resume_is_bit_identical():
    """3 + 2 sweeps with a checkpoint / restore into a NEW handle in between == 5 sweeps straight (warm start of
    particle_Gibbs.jl:114,126-128 made exact)."""
    from test_gpu_parity import _models
    basis, om, _ = _models()["known"]
    N, T, nc = 500, 40, 2
    u, y, _ = P.make_data(T, 61)

    def engine():
        e = pg.ParticleGibbsEngine(basis, 1, N, T, n_chains=nc)
        e.set_data(u, y)
        e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        e.set_prior(10 * np.ones(5), 100 * np.eye(2), 10.0)
        return e

    with engine() as e:
        for c in range(nc):
            e.set_state(np.zeros((2, 5)), 100 * np.eye(2), None, 1, chain=c)
        e.set_rng_philox(5, 7)
        e.sweep(5)
        ref = [e.get_state(c) for c in range(nc)]
    with engine() as e:
        for c in range(nc):
            e.set_state(np.zeros((2, 5)), 100 * np.eye(2), None, 1, chain=c)
        e.set_rng_philox(5, 7)
        e.sweep(3)
        ck = e.checkpoint()
    assert bytes(ck[:7]) == b"PGB200C"
    with engine() as e:
        e.restore(ck)
        e.sweep(2)
        got = [e.get_state(c) for c in range(nc)]
    for c in range(nc):
        for a, b, name in zip(got[c], ref[c], ("A", "Q", "x_prim", "sweep_index")):
            if name == "sweep_index":
                assert a == b == 6
            else:
                _assert_same("chain %d %s" % (c, name), a, b)


# ------------------------------------------------------------------------------------------------------------------
# stacked-scenario dynamics + Jacobians (optimal_control_Altro.jl:79-107, @autodiff :51)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["auto", "strided", "bulk4", "bulk2"])
def dynamics_kernel(request):
    """k_dynamics (strided loads) and k_dynamics_bulk (cp.async.bulk staging; as many warps per CTA as shared memory holds
    rows of A, or capped at 4 / 2) on the same shapes"""
    _lib.lib().pgb_test_dynamics_kernel({"auto": -1, "strided": 0, "bulk4": 1, "bulk2": 2}[request.param])
    yield request.param
    _lib.lib().pgb_test_dynamics_kernel(-1)


@pytest.mark.parametrize("model,variant", [("known", 0), ("hilbert4", 0), ("hilbert4", 1), ("hilbert2", 0),
                                           ("hilbert4_ragged", 0), ("hilbert2_5", 1), ("hilbert4_tail1", 1)])
def test_stacked_dynamics_and_jacobians(model, variant, dynamics_kernel):
    rng = np.random.default_rng(91)
    if model == "known":
        basis, om, n_x = pg.KnownBasisVB(), O.known_model(), 2
    elif model == "hilbert4":
        basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
        om, n_x = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L), 4
    elif model == "hilbert4_ragged":  # n_phi = 360: c = 3, the last lanes' chunks are partial / empty
        basis = pg.HilbertGPBasis([2, 3, 5, 4, 3], [4.0, 5.0, 6.0, 7.0, 8.0])
        om, n_x = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L), 4
    elif model == "hilbert4_tail1":   # n_phi = 625 with five functions in the fastest dimension (runs = its rows)
        basis = pg.HilbertGPBasis([5, 5, 5, 5, 1], [4.0, 5.0, 6.0, 7.0, 8.0])
        om, n_x = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L), 4
    elif model == "hilbert2_5":      # n_x = 2 with three inputs
        basis = pg.HilbertGPBasis([2, 1, 3, 4, 5], [4.0, 5.0, 6.0, 7.0, 8.0], n_u=3)
        om, n_x = O.hilbert_model(2, 3, basis.count, basis.stride, basis.L), 2
    else:
        basis = pg.HilbertGPBasis([3, 4, 2], [6.0, 9.0, 11.0])
        om, n_x = O.hilbert_model(2, 1, basis.count, basis.stride, basis.L), 2
    n_u = len(basis.count) - n_x if model != "known" else 1
    K, H, N = 77, 7, 12
    A = rng.normal(0, 0.3, (K, n_x, basis.n_phi))
    Bq = rng.normal(0, 0.1, (K, n_x, n_x))
    Q = Bq @ np.transpose(Bq, (0, 2, 1)) + 0.05 * np.eye(n_x)
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, n_u))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    u_init = rng.normal(0, 1, (n_u, H))
    batch = pg.PGSampleBatch(A, Q, w, np.transpose(x, (0, 2, 1)), um)
    with pg.DeviceSamples.from_samples(batch, basis) as ds:
        sc = ds.rollout(pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=4).scenarios()
        t = 3
        xn, jx, ju = ds.dynamics(sc["X"][:, t].reshape(-1), u_init[:, t], t=t, phi_variant=variant)
        rx, rjx, rju = O.dynamics(om, A, sc["X"][:, t], u_init[:, t], sc["v"], t, variant)
        _assert_same("x_next", xn.reshape(K, n_x), rx)
        _assert_same("df/dx", jx, rjx)
        _assert_same("df/du", ju, rju)
        if variant == 0:  # the same function the rollout iterates: X[t+1], bit for bit
            _assert_same("x_next == X[t+1]", xn.reshape(K, n_x), sc["X"][:, t + 1])
        # without the noise term, and against central differences of the kernel itself
        x0 = sc["X"][:, t].copy()
        f0, _, _ = ds.dynamics(x0.reshape(-1), u_init[:, t], t=None, phi_variant=variant, jacobians=False)
        _assert_same("no noise", f0.reshape(K, n_x), O.dynamics(om, A, x0, u_init[:, t], None, 0, variant)[0])
        eps = 1e-6
        for c in range(n_x):
            xp, xm = x0.copy(), x0.copy()
            xp[:, c] += eps
            xm[:, c] -= eps
            fd = (ds.dynamics(xp.reshape(-1), u_init[:, t], t=None, phi_variant=variant, jacobians=False)[0]
                  - ds.dynamics(xm.reshape(-1), u_init[:, t], t=None, phi_variant=variant, jacobians=False)[0]) / (2 * eps)
            np.testing.assert_allclose(jx[:, :, c], fd.reshape(K, n_x), rtol=0, atol=2e-8)
        for c in range(n_u):
            du = np.zeros(n_u)
            du[c] = eps
            fu = (ds.dynamics(x0.reshape(-1), u_init[:, t] + du, t=None, phi_variant=variant, jacobians=False)[0]
                  - ds.dynamics(x0.reshape(-1), u_init[:, t] - du, t=None, phi_variant=variant, jacobians=False)[0]) / (2 * eps)
            np.testing.assert_allclose(ju[:, :, c], fu.reshape(K, n_x), rtol=0, atol=2e-8)


@pytest.mark.parametrize("jac", [True, False])
def test_stacked_dynamics_many_scenarios(dynamics_kernel, jac):
    """K large enough that every warp of the persistent grid walks several scenarios (both copy stages, both mbarrier
    phases of k_dynamics_bulk), the BASELINE.json configs[4] shape (n_x = 4, n_phi = 625), a ragged last shard."""
    rng = np.random.default_rng(5)
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
    om, n_x, N = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L), 4, 3
    K, H = 148 * 11 * 3 + 37, 2
    A = rng.normal(0, 0.3, (K, n_x, basis.n_phi))
    Q = np.broadcast_to(0.05 * np.eye(n_x), (K, n_x, n_x)).copy()
    w = np.full((K, N), 1.0 / N)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, 1))
    u_init = rng.normal(0, 1, (1, H))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    batch = pg.PGSampleBatch(A, Q, w, np.transpose(x, (0, 2, 1)), um)
    xv = rng.normal(0, 2, (K, n_x))
    with pg.DeviceSamples.from_samples(batch, basis) as ds:
        sc = ds.rollout(pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=4).scenarios(want=("v",))
        for variant in (0, 1):
            xn, jx, ju = ds.dynamics(xv.reshape(-1), u_init[:, 1], t=1, phi_variant=variant, jacobians=jac)
            rx, rjx, rju = O.dynamics(om, A, xv, u_init[:, 1], sc["v"], 1, variant)
            _assert_same("x_next", xn.reshape(K, n_x), rx)
            if jac:
                _assert_same("df/dx", jx, rjx)
                _assert_same("df/du", ju, rju)
