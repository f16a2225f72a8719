"""The F32 precision mode (pgb_config.precision = PGB_PRECISION_F32: "FP32 particle state" of the north star) against
the oracle's F32 mode (oracle/pg_oracle.c: orc_sweep_f32) — every ancestor index, every state, weight and statistic bit
for bit, with injected streams and with Philox streams, stored and replayed histories, one chain and batched chains,
up to the benchmarked N = 2^20 geometry; the device build of pgb_mathf.h against the host build; and the stated
tolerance against the FP64 mode."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import _problems as P
import pgopt_b200 as pg
from test_gpu_parity import _assert_same, _models

pytestmark = pytest.mark.gpu


def _engine(basis, N, T, u, y, n_chains=1, keep_w=True, x_history="auto", R=0.1, n_y=1, Cm=None):
    e = pg.ParticleGibbsEngine(basis, n_y, N, T, n_chains=n_chains, keep_w_history=keep_w, x_history=x_history, precision="f32")
    e.set_data(u, y)
    e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]] if Cm is None else Cm), R)
    e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
    return e


@pytest.mark.parametrize("model", ["known", "hilbert"])
@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("N,T", [(30, 120), (1500, 25), (9000, 12)])
def test_f32_sweep_bit_exact_injected_streams(model, first, N, T):
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 11)
    s = P.injected_streams(N, T, 2, basis.n_phi, 21)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    args = (om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE)
    if not first:
        xp = O.sweep(*args, xp, True, st, 1, f32=True)["x_prim"]
    k = 1 if first else 2
    ref = O.sweep(*args, xp, first, st, k, f32=True)
    with _engine(basis, N, T, u, y) as e:
        e.set_state(A, P.Q_TRUE, xp, k)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        Phi, Psi, Sig, star = e.get_stats()
        _, _, xprim, _ = e.get_state()
        smp = e.get_sample()
        stat, nfb, _ = e.status()
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x_pf", x, ref["x"])
    _assert_same("w", w, ref["w"])
    assert star == ref["star"]
    _assert_same("x_prim", xprim, ref["x_prim"])
    _assert_same("Phi", Phi, ref["Phi"])
    _assert_same("Psi", Psi, ref["Psi"])
    _assert_same("Sigma", Sig, ref["Sigma"])
    _assert_same("w_m1", smp.w_m1, ref["w_last"])
    _assert_same("x_m1", smp.x_m1.T, ref["x_last"])
    assert stat == ref["status"] == 0 and nfb == 0
    assert np.array_equal(x, x.astype(np.float32).astype(np.float64))  # the state really is FP32


@pytest.mark.parametrize("model,N,T,x_history", [("known", 5000, 20, "store"), ("known", 5000, 20, "replay"),
                                                 ("hilbert", 3000, 16, "replay"), ("hilbert", 700, 30, "store")])
def test_f32_sweeps_philox_store_and_replay(model, N, T, x_history):
    """two sweeps (k = 1 bootstrap, k = 2 conditional) with Philox streams; the replayed lineage equals the stored one"""
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 12)
    st = O.make_streams(philox=True, seed=77, chain=6)
    args = (om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE)
    r1 = O.sweep(*args, np.zeros((2, T)), True, st, 1, f32=True)
    r2 = O.sweep(*args, r1["x_prim"], False, st, 2, f32=True)
    with _engine(basis, N, T, u, y, x_history=x_history) as e:
        e.set_rng_philox(77, 6)
        for k, (xp, ref) in enumerate(((np.zeros((2, T)), r1), (r1["x_prim"], r2)), start=1):
            e.set_state(A, P.Q_TRUE, xp, k)
            e.sweep_filter()
            a, _, w = e.get_history(want_x=False)
            Phi, Psi, Sig, star = e.get_stats()
            _, _, xprim, _ = e.get_state()
            smp = e.get_sample()
            _assert_same("ancestors k=%d" % k, a[1:], ref["a"][1:])
            _assert_same("w", w, ref["w"])
            assert star == ref["star"]
            _assert_same("x_prim", xprim, ref["x_prim"])
            _assert_same("Sigma", Sig, ref["Sigma"])
            _assert_same("Psi", Psi, ref["Psi"])
            _assert_same("w_m1", smp.w_m1, ref["w_last"])
            _assert_same("x_m1", smp.x_m1.T, ref["x_last"])
        assert e.status()[0] == 0


def test_f32_two_outputs_general_basis_batched_chains():
    """n_y = 2 (Cholesky-solve log-weight in FP32), general Hilbert basis (BASIS 2 path), 3 batched chains"""
    rng = np.random.default_rng(4)
    basis = pg.HilbertGPBasis([3, 4, 2], [6.0, 9.0, 11.0])
    om = O.hilbert_model(2, 1, basis.count, basis.stride, basis.L, n_y=2)
    N, T, nc = 1200, 14, 3
    u, y1, _ = P.make_data(T, 13)
    y = np.vstack([y1, 0.5 * y1 + 0.1 * rng.normal(size=y1.shape)])
    Cm = np.array([[1.0, 0.0], [0.3, 1.0]])
    R = np.array([[0.1, 0.02], [0.02, 0.2]])
    A = rng.normal(0, 1.0, (2, basis.n_phi))
    refs = []
    for c in range(nc):
        st = O.make_streams(philox=True, seed=5, chain=10 + c)
        r1 = O.sweep(om, N, T, u, y, Cm, R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, np.zeros((2, T)), True, st, 1, f32=True)
        refs.append((r1, O.sweep(om, N, T, u, y, Cm, R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, r1["x_prim"], False, st, 2, f32=True)))
    with _engine(basis, N, T, u, y, n_chains=nc, R=R, n_y=2, Cm=Cm) as e:
        e.set_rng_philox(5, 10)
        for c in range(nc):
            e.set_state(A, P.Q_TRUE, refs[c][0]["x_prim"], 2, chain=c)
        e.sweep_filter()
        for c in range(nc):
            a, x, w = e.get_history(chain=c)
            _assert_same("ancestors chain %d" % c, a[1:], refs[c][1]["a"][1:])
            _assert_same("x", x, refs[c][1]["x"])
            _assert_same("w", w, refs[c][1]["w"])
            _assert_same("x_prim", e.get_state(chain=c)[2], refs[c][1]["x_prim"])


@pytest.mark.parametrize("model,T", [("hilbert", 5), ("known", 6)])
def test_f32_headline_geometry(model, T):
    """N = 2^20, replay mode, Philox streams: the benchmarked F32 configuration against the oracle"""
    basis, om, A = _models()[model]
    N = 2 ** 20
    u, y, _ = P.make_data(T, 31)
    st = O.make_streams(philox=True, seed=2024, chain=5)
    args = (om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE)
    r1 = O.sweep(*args, np.zeros((2, T)), True, st, 1, f32=True)
    r2 = O.sweep(*args, r1["x_prim"], False, st, 2, f32=True)
    with _engine(basis, N, T, u, y, x_history="replay") as e:
        e.set_rng_philox(2024, 5)
        for k, (xp, ref) in enumerate(((np.zeros((2, T)), r1), (r1["x_prim"], r2)), start=1):
            e.set_state(A, P.Q_TRUE, xp, k)
            e.sweep_filter()
            a, _, w = e.get_history(want_x=False)
            _assert_same("ancestors k=%d" % k, a[1:], ref["a"][1:])
            _assert_same("w", w, ref["w"])
            assert e.get_stats()[3] == ref["star"]
            _assert_same("x_prim", e.get_state()[2], ref["x_prim"])
            smp = e.get_sample()
            _assert_same("w_m1", smp.w_m1, ref["w_last"])
            _assert_same("x_m1", smp.x_m1.T, ref["x_last"])
        stat, nfb, _ = e.status()
        assert stat == 0 and nfb == 0


def test_f32_device_math_is_bit_identical_to_host():
    """pgb_mathf.h compiled for sm_100a == compiled by gcc (the oracle): exp, log, sin, cos, wide-range exp, normals"""
    from pgopt_b200._lib import lib, check
    rng = np.random.default_rng(8)
    cases = {30: (0, rng.uniform(-90, 89, 200000)), 31: (1, (rng.integers(0, 2 ** 23, 200000) + 0.5) * 2.0 ** -23),
             32: (2, np.concatenate([rng.uniform(-20, 20, 100000), rng.uniform(-4e4, 4e4, 100000)])),
             33: (3, np.concatenate([rng.uniform(-20, 20, 100000), rng.uniform(-4e4, 4e4, 100000)])),
             34: (4, rng.uniform(-710, 100, 200000))}
    for fn, (ofn, x) in cases.items():
        x = np.ascontiguousarray(x.astype(np.float32).astype(np.float64))
        y = np.empty_like(x)
        check(lib().pgb_test_math(C.c_int32(0), C.c_int32(fn), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int64(x.size)))
        xf = x.astype(np.float32)
        if ofn == 4:  # the wide exponential is an FP64 value: compare through its exact decomposition
            ref = np.empty_like(xf)
            O.lib().orc_vmathf(C.c_int(4), xf.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), C.c_int64(xf.size))
            ok = np.isfinite(ref) & (ref != 0) & (xf > -87)
            assert np.array_equal(y[ok].astype(np.float32), ref[ok])
            assert (y[x < -700] == 0).all() and np.all(y[(x > -699) & (x < -88)] > 0)
        else:
            ref = np.empty_like(xf)
            O.lib().orc_vmathf(C.c_int(ofn), xf.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), C.c_int64(xf.size))
            assert np.array_equal(y.astype(np.float32).view(np.uint32), ref.view(np.uint32)), fn
    n = 50000
    x = np.array([42.0])
    y = np.empty(4 * n)
    check(lib().pgb_test_math(C.c_int32(0), C.c_int32(35), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int64(n)))
    z = np.empty((n, 4), dtype=np.float32)
    O.lib().orc_vnormal_pairs_f32(C.c_uint64(42), C.c_uint32(2), C.c_uint32(0), C.c_uint32(1), C.c_uint32(5), C.c_int64(n),
                                  C.c_int(4), z.ctypes.data_as(C.c_void_p))
    assert np.array_equal(y.reshape(n, 4).astype(np.float32).view(np.uint32), z.view(np.uint32))


def test_f32_ancestor_density_below_flt_min():
    """a step where the transition density of EVERY particle is below FLT_MIN (reference trajectory far from the model's
    prediction): the wide-range exponential keeps the ancestor weights finite, as the oracle does"""
    basis, om, A = _models()["known"]
    N, T = 2000, 8
    u, y, _ = P.make_data(T, 15)
    st = O.make_streams(philox=True, seed=3, chain=1)
    args = (om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE)
    xp = O.sweep(*args, np.zeros((2, T)), True, st, 1, f32=True)["x_prim"]
    xp[:, 4] += 1.6  # |L^-1 (F - x'_4)|^2 / 2 of the order of 170: exp underflows FP32 (e^-87), not the wide-range version
    ref = O.sweep(*args, xp, False, st, 2, f32=True)
    assert ref["status"] == 0
    with _engine(basis, N, T, u, y) as e:
        e.set_rng_philox(3, 1)
        e.set_state(A, P.Q_TRUE, xp, 2)
        e.sweep_filter()
        a, x, w = e.get_history()
        assert e.status()[0] == 0
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x", x, ref["x"])
    _assert_same("w", w, ref["w"])


def test_f32_particle_gibbs_whole_sampler_runs_and_is_sane():
    """the whole sampler in F32 (MNIW stays FP64): finite samples, and the learned model predicts the training data about
    as well as the FP64 sampler's (the two are different realisations — indices differ after a few steps — so this is a
    coarse statement; the bit-level statement is the oracle parity above)"""
    basis, om, _ = _models()["known"]
    N, T, K = 200, 150, 12
    u, y, _ = P.make_data(T, 14)
    V = 10 * np.ones(5)
    rmse = {}
    for prec in ("f32", "f64"):
        smp = pg.particle_Gibbs(u, y, K, 20, 1, N, basis, 100 * np.eye(2), 10.0, 100 * np.eye(2), V, np.zeros((2, 5)),
                                pg.MvNormal([2.0, 2.0], 1.0), pg.LinearMeasurement([[1.0, 0.0]]), 0.1, seed=3, chain=0,
                                precision=prec)
        assert all(np.isfinite(s.A).all() and np.isfinite(s.Q).all() for s in smp)
        out = pg.test_prediction(smp, basis, pg.LinearMeasurement([[1.0, 0.0]]), 0.1, 5, u, y, seed=1)
        rmse[prec] = out["mean_rmse"]
    assert np.isfinite(rmse["f32"]) and rmse["f32"] < 2.0 * rmse["f64"] + 0.5


def test_f32_rejects_other_kernels_and_modes():
    basis, om, A = _models()["known"]
    u, y, _ = P.make_data(10, 1)
    with _engine(basis, 30, 10, u, y) as e:  # N = 30 would pick the warp kernel in FP64; F32 always runs the sweep kernel
        e.set_state(A, P.Q_TRUE, None, 1)
        e.set_rng_philox(1, 0)
        e.sweep_filter()
        assert e.status()[0] == 0
