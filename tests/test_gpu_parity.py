"""GPU parity tests: the CUDA path (through the C ABI of libpgopt_b200.so) against the CPU oracle on the
same seeded inputs.  Bar: bit-exact for indices; floats bit-exact too wherever the same defined order
runs on both sides (everything except the stated-tolerance cases)."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import _problems as P
import pgopt_b200 as pg
from pgopt_b200 import _lib

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------ math
@pytest.mark.parametrize("fn,name,lo,hi", [(0, "exp", -750, 710), (1, "log", 1e-300, 1e300), (2, "sin", -1e5, 1e5),
                                           (3, "cos", -1e5, 1e5)])
def test_device_math_is_bit_identical_to_host(fn, name, lo, hi):
    rng = np.random.default_rng(fn)
    x = rng.uniform(lo, hi, 200000) if fn != 1 else 10.0 ** rng.uniform(-300, 300, 200000)
    x = np.concatenate([x, rng.uniform(-3, 3, 100000) if fn != 1 else rng.uniform(0, 2, 100000)])
    y = np.empty_like(x)
    _lib.check(_lib.lib().pgb_test_math(0, fn, _p(x), _p(y), C.c_int64(x.size)))
    assert np.array_equal(y.view(np.uint64), O.vfun(name, x).view(np.uint64))


def test_device_philox_normals_bit_identical():
    n = 100000
    y = np.empty(2 * n)
    x = np.array([1234.0])
    _lib.check(_lib.lib().pgb_test_math(0, 4, _p(x), _p(y), C.c_int64(n)))
    assert np.array_equal(y.view(np.uint64), O.normal_pairs(1234, 2, 0, 1, 5, n).view(np.uint64))


@pytest.mark.parametrize("fn,name", [(10, "exp"), (11, "log"), (12, "sin"), (13, "cos")])
def test_vector_math_is_bit_identical_to_scalar(fn, name):
    """pgb_vmath.cuh (4-wide, branch-free, used in the persistent kernel) == pgb_math.h on the host."""
    rng = np.random.default_rng(fn)
    if fn == 10:
        x = np.concatenate([rng.uniform(-750, 712, 300001), rng.uniform(-40, 1, 200000),
                            [0.0, -0.0, 709.78, 709.79, 710.5, -745.1, -745.2, -746.0, -746.1, -1e300, 1e300, np.nan, np.inf, -np.inf]])
    elif fn == 11:
        x = np.concatenate([rng.uniform(0, 1, 300000), 10.0 ** rng.uniform(-300, 300, 100003), [2.0 ** -54, 1.0, 2.0 ** -1022]])
    else:
        x = np.concatenate([rng.uniform(-7, 7, 300001), rng.uniform(-1e5, 1e5, 200000), [0.0, -0.0, 1e14, 2e15, np.inf, np.nan]])
    y = np.empty_like(x)
    _lib.check(_lib.lib().pgb_test_math(0, fn, _p(x), _p(y), C.c_int64(x.size)))
    ref = O.vfun(name, x)
    same = (y.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(y) & np.isnan(ref))
    assert same.all(), "%d differ, first x=%r: %r vs %r" % ((~same).sum(), x[~same][0], y[~same][0], ref[~same][0])


def test_invariant_division_is_correctly_rounded():
    """div_inv (reciprocal + two FMA corrections, used for the shared-divisor quotients w/S, lw/R, ...) must equal the
    IEEE division bit for bit: 2^32 self-generated operand pairs on the device (random and adversarial mantissas,
    numerators next to exact multiples) plus explicit special values through both paths."""
    n = 1 << 20
    x = np.arange(n, dtype=np.float64) + 17.0
    y = np.empty(n)
    _lib.check(_lib.lib().pgb_test_math(0, 21, _p(x), _p(y), C.c_int64(n)))
    assert y.sum() == 0, "%d mismatching quotients" % int(y.sum())
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.uniform(0, 1, 50000), 10.0 ** rng.uniform(-320, 300, 50000), [0.0, -0.0, 5e-324, 1e-310, np.inf, np.nan, 1.0]])
    d = np.concatenate([rng.uniform(0.5, 2000, 50000), 10.0 ** rng.uniform(-320, 300, 50000), [3.0, 3.0, 3.0, 7.0, 2.0, 2.0, 0.0]])
    xy = np.ascontiguousarray(np.stack([a, d], axis=1).ravel())
    ok = np.empty(a.size)
    _lib.check(_lib.lib().pgb_test_math(0, 20, _p(xy), _p(ok), C.c_int64(a.size)))
    assert ok.all()


def test_vector_philox_normals_bit_identical():
    n = 100003
    y = np.empty(2 * n)
    x = np.array([987654321.0])
    _lib.check(_lib.lib().pgb_test_math(0, 14, _p(x), _p(y), C.c_int64(n)))
    assert np.array_equal(y.view(np.uint64), O.normal_pairs(987654321, 2, 0, 1, 5, n).view(np.uint64))


# ------------------------------------------------------------------------------------------ resampling
def _weights(kind, N, rng):
    if kind == 0:
        return rng.uniform(0, 1, N)
    if kind == 1:
        return np.exp(-0.5 * rng.normal(0, 3, N) ** 2 / 0.1)
    if kind == 2:
        w = np.zeros(N)
        w[rng.integers(0, N)] = 1.0
        w[N - 1] += 1e-300
        return w
    if kind == 3:
        return rng.uniform(0, 1, N) * (rng.uniform(0, 1, N) < 0.05) + 1e-320
    if kind == 4:
        return np.exp(rng.normal(0, 30, N))
    if kind == 5:
        return np.ones(N)
    w = np.exp(-0.5 * (rng.normal(0, 0.2, N)) ** 2 / 0.1)
    return w / w.sum()


def test_resampling_hand_trace():
    assert pg.systematic_resampling([.1, .2, .3, .4], 4, rnd=0.5).tolist() == [2, 3, 4, 4]
    assert pg.systematic_resampling([1, 1, 1, 1], 3, rnd=0.999).tolist() == [2, 3, 4]
    idx, st = pg.systematic_resampling([1.0, 1.0], 2, rnd=0.0, return_status=True)
    assert idx.tolist() == [0, 1] and st & _lib.STATUS_INDEX_ZERO


@pytest.mark.parametrize("N", [2, 3, 5, 29, 30, 31, 1000, 1024, 1025, 4097, 65536, 2 ** 18 + 3])
def test_resampling_bit_exact(N):
    rng = np.random.default_rng(N)
    for kind in range(7):
        w = _weights(kind, N, rng)
        for nd in sorted({1, max(1, N - 1), N, min(2 * N, 1000)}):
            rnd = rng.uniform()
            a, st = O.systematic_resampling(w, nd, rnd)
            b, stg = pg.systematic_resampling(w, nd, rnd=rnd, return_status=True)
            assert np.array_equal(a, b), "N=%d kind=%d nd=%d: %d indices differ" % (N, kind, nd, np.sum(a != b))
            assert (stg & 3) == st
            assert not (stg & 0x100), "serial fallback ran on a sane input (N=%d kind=%d nd=%d)" % (N, kind, nd)


@pytest.mark.parametrize("N", [2 ** 20, 2 ** 22])
def test_resampling_bit_exact_large(N):
    rng = np.random.default_rng(7)
    w = np.exp(-0.5 * rng.normal(0, 1, N) ** 2 / 0.3)
    a, st = O.systematic_resampling(w, N - 1, 0.37)
    b, stg = pg.systematic_resampling(w, N - 1, rnd=0.37, return_status=True)
    assert st == 0 and np.array_equal(a, b) and not (stg & 0x100)
    # size-independent properties
    assert np.all(np.diff(b) >= 0) and b.min() >= 1 and b.max() <= N
    counts = np.bincount(b - 1, minlength=N)
    assert np.all(np.abs(counts - (N - 1) * w / w.sum()) <= 1.0 + 1e-6)


def test_resampling_forced_fallback_is_still_exact():
    rng = np.random.default_rng(3)
    _lib.lib().pgb_test_force_fallback(1)
    try:
        for N in (30, 5000, 70000):
            w = _weights(1, N, rng)
            rnd = rng.uniform()
            a, _ = O.systematic_resampling(w, N - 1, rnd)
            b, stg = pg.systematic_resampling(w, N - 1, rnd=rnd, return_status=True)
            assert np.array_equal(a, b) and (stg & 0x100)
    finally:
        _lib.lib().pgb_test_force_fallback(0)


def test_resampling_overrun_status():
    # weights whose sequential sum stays below the last grid point: the reference raises BoundsError
    w = np.array([0.1] * 10)
    rnd = 1.0 - 2.0 ** -53
    a, st = O.systematic_resampling(w, 10, rnd)
    b, stg = pg.systematic_resampling(w, 10, rnd=rnd, return_status=True)
    assert np.array_equal(a, b) and (stg & 1) == (st & 1)
    if st & 1:
        with pytest.raises(IndexError):
            pg.systematic_resampling(w, 10, rnd=rnd)


# ------------------------------------------------------------------------------------------ sweep
def _models():
    hb = pg.HilbertGPBasis([5, 5, 5], [10, 20, 20])
    return {"known": (pg.KnownBasisVB(), O.known_model(), P.plausible_A_known()),
            "hilbert": (hb, O.hilbert_model(2, 1, hb.count, hb.stride, hb.L),
                        np.random.default_rng(9).normal(0, 2.0, (2, 125)))}


def _engine(basis, N, T, u, y, keep_w=True, n_chains=1, persistent=True, R=0.1):
    e = pg.ParticleGibbsEngine(basis, 1, N, T, n_chains=n_chains, keep_w_history=keep_w)
    e.set_mode(persistent)
    e.set_data(u, y)
    e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), R)
    e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
    return e


def _assert_same(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        bad = a.view(np.uint64) != b.view(np.uint64)
        bad &= ~((a == 0) & (b == 0))
    else:
        bad = a != b
    assert not bad.any(), "%s: %d of %d entries differ (first at %s: %r vs %r)" % (
        name, bad.sum(), bad.size, np.argwhere(bad)[0], a[tuple(np.argwhere(bad)[0])], b[tuple(np.argwhere(bad)[0])])


@pytest.mark.parametrize("persistent", [4, 1, 0])
@pytest.mark.parametrize("model", ["known", "hilbert"])
@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("N,T", [(30, 120), (1500, 25), (9000, 12)])
def test_sweep_bit_exact_injected_streams(model, first, N, T, persistent):
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 11)
    s = P.injected_streams(N, T, 2, basis.n_phi, 21)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    if not first:  # a plausible conditioned trajectory: the output of a first sweep
        xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, True, st, 1)["x_prim"]
    k = 1 if first else 2
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, first, st, k)
    with _engine(basis, N, T, u, y, persistent=persistent) as e:
        e.set_state(A, P.Q_TRUE, xp, k)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        Phi, Psi, Sig, star = e.get_stats()
        _, _, xprim, _ = e.get_state()
        smp = e.get_sample()
        stat, nfb, nscan = e.status()
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
    # T main scans (+1 star draw counted in them); the ancestor draw only needs the exact scan when the certified
    # approximate shortcut is undecided (persistent mode), always in the multi-kernel mode
    assert (T <= nscan <= 2 * T - 1) if (persistent and not first) else nscan == (T if first else 2 * T - 1)


@pytest.mark.parametrize("model", ["known", "hilbert"])
@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("N,T", [(30, 200), (32, 40), (17, 60), (2, 30), (5, 25)])
def test_small_n_kernel_bit_exact_injected_streams(model, first, N, T):
    """The warp-per-chain kernel (mode 3; the automatic choice for N <= 32, the reference's examples use N = 30)
    against the oracle: every ancestor index, particle, weight and statistic bit for bit."""
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 13)
    s = P.injected_streams(N, T, 2, basis.n_phi, 23)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    if not first:
        xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, True, st, 1)["x_prim"]
    k = 1 if first else 2
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, first, st, k)
    for mode in (3, True):
        with _engine(basis, N, T, u, y, persistent=mode) as e:
            e.set_state(A, P.Q_TRUE, xp, k)
            e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
            e.sweep_filter()
            a, x, w = e.get_history()
            Phi, Psi, Sig, star = e.get_stats()
            _, _, xprim, _ = e.get_state()
            smp = e.get_sample()
            stat, nfb, nscan = e.status()
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
        assert stat == ref["status"] == 0 and nfb == 0 and nscan == 0  # no exact-scan machinery on this path


@pytest.mark.parametrize("model", ["known", "hilbert"])
@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("N,T", [(33, 40), (64, 30), (100, 60), (257, 25), (1000, 12), (1024, 10)])
def test_mid_n_cta_kernel_bit_exact_injected_streams(model, first, N, T):
    """The one-CTA-per-chain kernel (mode 3 for 32 < N <= 1024; thread = particle, block tree sums, sequential prefix
    walked once into shared memory, binary search) against the oracle, bit for bit."""
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 14)
    s = P.injected_streams(N, T, 2, basis.n_phi, 25)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    if not first:
        xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, True, st, 1)["x_prim"]
    k = 1 if first else 2
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, first, st, k)
    with _engine(basis, N, T, u, y, persistent=3) as e:
        e.set_state(A, P.Q_TRUE, xp, k)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        Phi, Psi, Sig, star = e.get_stats()
        _, _, xprim, _ = e.get_state()
        smp = e.get_sample()
        stat, nfb, nscan = e.status()
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x_pf", x, ref["x"])
    _assert_same("w", w, ref["w"])
    assert star == ref["star"]
    _assert_same("x_prim", xprim, ref["x_prim"])
    _assert_same("Sigma", Sig, ref["Sigma"])
    _assert_same("w_m1", smp.w_m1, ref["w_last"])
    _assert_same("x_m1", smp.x_m1.T, ref["x_last"])
    assert stat == ref["status"] == 0 and nfb == 0 and nscan == 0


def test_small_n_kernel_matches_tile_kernel_on_whole_sampler():
    """particle_Gibbs at the reference's N = 30 through the automatic (warp-per-chain) path and through the tile
    persistent kernel: identical samples; both equal the oracle (Philox streams)."""
    basis, om, _ = _models()["hilbert"]
    N, T, K, K_b, k_d = 30, 80, 3, 2, 1
    u, y, _ = P.make_data(T, 3)
    V = pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3), 100 ** 2 / (8 * np.pi ** 4))
    ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, 125)),
                           100 * np.eye(2), V, 100 * np.eye(2), 10.0, K, K_b, k_d, seed=5, chain=4)
    outs = []
    for mode in (True, 1):
        e = pg.ParticleGibbsEngine(basis, 1, N, T)
        try:
            e.set_mode(mode)
            e.set_data(u, y)
            e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
            e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
            e.set_prior(V, 100 * np.eye(2), 10.0)
            e.set_state(np.zeros((2, 125)), 100 * np.eye(2), np.zeros((2, T)), 1)
            e.set_rng_philox(5, 4)
            outs.append(e.particle_gibbs(K, K_b, k_d))
        finally:
            e.close()
    for name, i in (("A", 0), ("Q", 1), ("w", 2), ("x", 3)):
        _assert_same(name + " auto vs tile", outs[0][i], outs[1][i])
    for kk in range(K):
        _assert_same("A[%d]" % kk, outs[0][0][0, kk].reshape((2, 125), order="F"), ref["A"][kk])
        _assert_same("w[%d]" % kk, outs[0][2][0, kk], ref["w"][kk])


@pytest.mark.parametrize("persistent", [4, 1, 0])
def test_sweep_with_degenerate_weights(persistent):
    """A very sharp likelihood makes a handful of particles own almost all offspring (the cooperative
    heavy-particle expansion of the persistent kernel); results must not change."""
    basis, om, A = _models()["known"]
    N, T = 6000, 15
    u, y, _ = P.make_data(T, 15)
    s = P.injected_streams(N, T, 2, 5, 23)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    R = 1e-7
    xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, np.zeros((2, T)), True, st, 1)["x_prim"]
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], R, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, False, st, 2)
    assert max(np.bincount(ref["a"][t, :N - 1]).max() for t in range(1, T)) > 1000
    with _engine(basis, N, T, u, y, persistent=persistent, R=R) as e:
        e.set_state(A, P.Q_TRUE, xp, 2)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        stat, nfb, nscan = e.status()
    assert (stat & 3) == (ref["status"] & 3)
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x_pf", x, ref["x"])
    _assert_same("w", w, ref["w"])


@pytest.mark.parametrize("persistent", [4, 1, 0])
def test_sweep_with_underflowing_ancestor_weights(persistent):
    """An implausible conditioned trajectory makes every ancestor weight underflow: waN ./ sum(waN) is NaN
    (particle_Gibbs.jl:180).  The reference carries on silently (index 1); so do we, bit for bit, and the
    NONFINITE status bit is raised."""
    basis, om, A = _models()["known"]
    N, T = 30, 10
    u, y, _ = P.make_data(T, 11)
    s = P.injected_streams(N, T, 2, 5, 21)
    xp = np.random.default_rng(4).normal(2, 1, (2, T))
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, False, st, 2)
    with _engine(basis, N, T, u, y, persistent=persistent) as e:
        e.set_state(A, P.Q_TRUE, xp, 2)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        stat, nfb, nscan = e.status()
    assert ref["status"] & 4 and stat == ref["status"]
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x_pf", x, ref["x"])
    _assert_same("w", w, ref["w"])


@pytest.mark.parametrize("model", ["known", "hilbert"])
@pytest.mark.parametrize("N,T", [(30, 100), (3000, 20)])
def test_ancestor_draw_shortcut_equals_exact_scan(model, N, T):
    """Persistent mode decides the ancestor index of the conditioned trajectory from an approximate scan with a
    rigorous error bound; forcing the exact scan instead (test hook 2) must give the same history."""
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 16)
    s = P.injected_streams(N, T, 2, basis.n_phi, 24)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, np.zeros((2, T)), True, st, 1)["x_prim"]
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, False, st, 2)
    hist = []
    for force in (0, 2):
        _lib.lib().pgb_test_force_fallback(force)
        try:
            with _engine(basis, N, T, u, y, persistent=1) as e:
                e.set_state(A, P.Q_TRUE, xp, 2)
                e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
                e.sweep_filter()
                a, x, w = e.get_history()
                stat, nfb, nscan = e.status()
        finally:
            _lib.lib().pgb_test_force_fallback(0)
        _assert_same("ancestors (force=%d)" % force, a[1:], ref["a"][1:])
        assert nfb == 0 and stat == 0
        hist.append(nscan)
    assert hist[1] == 2 * T - 1 and hist[0] < hist[1]  # the shortcut really skipped exact scans


@pytest.mark.parametrize("persistent", [4, 1, 0])
@pytest.mark.parametrize("N", [300, 5000])
def test_sweep_forced_fallback_is_still_exact(persistent, N):
    basis, om, A = _models()["known"]
    T = 20
    u, y, _ = P.make_data(T, 12)
    s = P.injected_streams(N, T, 2, 5, 22)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, np.zeros((2, T)), True, st, 1)["x_prim"]
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, False, st, 2)
    _lib.lib().pgb_test_force_fallback(1)
    try:
        with _engine(basis, N, T, u, y, persistent=persistent) as e:
            e.set_state(A, P.Q_TRUE, xp, 2)
            e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
            e.sweep_filter()
            a, x, w = e.get_history()
            stat, nfb, nscan = e.status()
    finally:
        _lib.lib().pgb_test_force_fallback(0)
    assert nfb == nscan and nfb > 0 and stat == ref["status"]
    _assert_same("ancestors", a[1:], ref["a"][1:])
    _assert_same("x_pf", x, ref["x"])
    _assert_same("w", w, ref["w"])


@pytest.mark.parametrize("model", ["known", "hilbert"])
def test_mniw_bit_exact(model):
    basis, om, A = _models()[model]
    rng = np.random.default_rng(31)
    n_x, n_phi, Tm1 = 2, basis.n_phi, 399
    zeta, z = rng.normal(size=(n_x, Tm1)), rng.normal(size=(n_phi, Tm1))
    Phi, Psi, Sig = zeta @ zeta.T, zeta @ z.T, z @ z.T
    Sig = 0.5 * (Sig + Sig.T)
    V = 10 * np.ones(n_phi) if model == "known" else pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3),
                                                                              100 ** 2 / (8 * np.pi ** 4))
    Lam, ell = 100 * np.eye(n_x), 10.0
    s = P.injected_streams(4, 4, n_x, n_phi, 3, df=ell + Tm1 * n_x)
    st = O.make_streams(philox=False, bartlett=s["bartlett"], Xmn=s["Xmn"])
    Ar, Qr, rc = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, st, 1)
    Ag, Qg = pg.MNIW_sample(Phi, Psi, Sig, V, Lam, ell, Tm1, bartlett=s["bartlett"], Xmn=s["Xmn"])
    assert rc == 0
    _assert_same("Q", Qg, Qr)
    _assert_same("A", Ag, Ar)
    # Philox-driven draw (Marsaglia-Tsang chi-square on the device) must match too
    stp = O.make_streams(philox=True, seed=77, chain=3)
    Ar, Qr, rc = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, stp, 5)
    Ag, Qg = pg.MNIW_sample(Phi, Psi, Sig, V, Lam, ell, Tm1, seed=77, chain=3, sweep_index=5)
    _assert_same("Q philox", Qg, Qr)
    _assert_same("A philox", Ag, Ar)


@pytest.mark.parametrize("n_x,n_phi", [(2, 135), (2, 136), (3, 200), (4, 625)])
def test_mniw_bit_exact_sizes(n_x, n_phi):
    """MNIW_sample at the edge of the shared-memory path (n_phi <= 135), just beyond it, and at the configs[4] basis size
    (n_phi = 5^4 = 625: global-memory path) — bit-identical to the oracle either way."""
    rng = np.random.default_rng(100 + n_phi)
    Tm1 = n_phi + 50
    zeta, z = rng.normal(size=(n_x, Tm1)), rng.normal(size=(n_phi, Tm1))
    Phi, Psi, Sig = zeta @ zeta.T, zeta @ z.T, z @ z.T
    Sig = 0.5 * (Sig + Sig.T)
    V = rng.uniform(0.5, 20.0, n_phi)
    Lam, ell = 100 * np.eye(n_x), 10.0
    stp = O.make_streams(philox=True, seed=5, chain=2)
    Ar, Qr, rc = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, stp, 3)
    assert rc == 0
    Ag, Qg = pg.MNIW_sample(Phi, Psi, Sig, V, Lam, ell, Tm1, seed=5, chain=2, sweep_index=3)
    _assert_same("Q", Qg, Qr)
    _assert_same("A", Ag, Ar)


def _dense_V(rng, n_phi, scale=1.0):
    B = rng.normal(size=(n_phi, n_phi)) / np.sqrt(n_phi)
    V = scale * (B @ B.T + 0.5 * np.eye(n_phi))
    return 0.5 * (V + V.T)


@pytest.mark.parametrize("n_x,n_phi", [(2, 5), (2, 125), (3, 200)])
def test_mniw_dense_V_bit_exact(n_x, n_phi):
    """MNIW_sample with a dense prior covariance V (particle_Gibbs.jl:56 takes any V; :64): pgb_mniw_sample_dense ==
    orc_mniw_sample_dense with injected and with Philox draws, on the shared-memory and the global-memory path."""
    rng = np.random.default_rng(200 + n_phi)
    Tm1 = n_phi + 60
    zeta, z = rng.normal(size=(n_x, Tm1)), rng.normal(size=(n_phi, Tm1))
    Phi, Psi, Sig = zeta @ zeta.T, zeta @ z.T, z @ z.T
    Sig = 0.5 * (Sig + Sig.T)
    V = _dense_V(rng, n_phi, 4.0)
    Lam, ell = 100 * np.eye(n_x), 10.0
    s = P.injected_streams(4, 4, n_x, n_phi, 3, df=ell + Tm1 * n_x)
    st = O.make_streams(philox=False, bartlett=s["bartlett"], Xmn=s["Xmn"])
    Ar, Qr, rc = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, st, 1)
    assert rc == 0
    Ag, Qg = pg.MNIW_sample(Phi, Psi, Sig, V, Lam, ell, Tm1, bartlett=s["bartlett"], Xmn=s["Xmn"])
    _assert_same("Q", Qg, Qr)
    _assert_same("A", Ag, Ar)
    stp = O.make_streams(philox=True, seed=6, chain=1)
    Ar, Qr, rc = O.mniw_sample(n_x, n_phi, Phi, Psi, Sig, V, Lam, ell, Tm1, stp, 2)
    Ag, Qg = pg.MNIW_sample(Phi, Psi, Sig, V, Lam, ell, Tm1, seed=6, chain=1, sweep_index=2)
    _assert_same("Q philox", Qg, Qr)
    _assert_same("A philox", Ag, Ar)
    # a diagonal matrix takes the diagonal path: the same bits as passing the diagonal
    Vd = rng.uniform(0.5, 20.0, n_phi)
    A1, Q1 = pg.MNIW_sample(Phi, Psi, Sig, Vd, Lam, ell, Tm1, seed=6, chain=1, sweep_index=2)
    A2, Q2 = pg.MNIW_sample(Phi, Psi, Sig, np.diag(Vd), Lam, ell, Tm1, seed=6, chain=1, sweep_index=2)
    _assert_same("diag == vector", A2, A1)
    with pytest.raises(pg.PgoptError):
        pg.MNIW_sample(Phi, Psi, Sig, V - 10 * np.eye(n_phi), Lam, ell, Tm1, seed=6)


@pytest.mark.parametrize("model,N", [("known", 30), ("hilbert", 300), ("known", 2500)])
def test_particle_gibbs_dense_V_bit_exact(model, N):
    """Whole sampler with a dense V (pgb_set_prior_dense) on the warp, CTA and sweep kernels: every kept sample identical
    to the oracle's."""
    basis, om, _ = _models()[model]
    T, K, K_b, k_d = 30, 3, 2, 1
    u, y, _ = P.make_data(T, 17)
    n_phi = basis.n_phi
    V = _dense_V(np.random.default_rng(3), n_phi, 10.0)
    ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, n_phi)),
                           100 * np.eye(2), V, 100 * np.eye(2), 10.0, K, K_b, k_d, seed=41, chain=1)
    smp = pg.particle_Gibbs(u, y, K, K_b, k_d, N, basis, 100 * np.eye(2), 10.0, 100 * np.eye(2), V, np.zeros((2, n_phi)),
                            pg.MvNormal([2.0, 2.0], 1.0), pg.LinearMeasurement([[1.0, 0.0]]), 0.1, seed=41, chain=1)
    assert ref["status"] == 0 and len(smp) == K
    for k in range(K):
        _assert_same("A[%d]" % k, smp[k].A, ref["A"][k])
        _assert_same("Q[%d]" % k, smp[k].Q, ref["Q"][k])
        _assert_same("w_m1[%d]" % k, smp[k].w_m1, ref["w"][k])


@pytest.mark.parametrize("model,N,T,K,K_b,k_d", [("known", 30, 150, 4, 6, 2), ("hilbert", 30, 60, 3, 3, 1),
                                                 ("known", 2500, 30, 3, 2, 1), ("hilbert", 300, 25, 2, 2, 1),
                                                 ("known", 1024, 20, 2, 1, 1)])
def test_particle_gibbs_bit_exact_philox(model, N, T, K, K_b, k_d):
    """Whole sampler (particle_Gibbs.jl:114-237) with counter-based streams: every kept sample identical."""
    basis, om, _ = _models()[model]
    u, y, _ = P.make_data(T, 13)
    n_phi = basis.n_phi
    V = 10 * np.ones(n_phi) if model == "known" else pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3),
                                                                            100 ** 2 / (8 * np.pi ** 4))
    ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, n_phi)),
                           100 * np.eye(2), V, 100 * np.eye(2), 10.0, K, K_b, k_d, seed=99, chain=2)
    xp = np.zeros((2, T))
    smp = pg.particle_Gibbs(u, y, K, K_b, k_d, N, basis, 100 * np.eye(2), 10.0, 100 * np.eye(2), V, np.zeros((2, n_phi)),
                            pg.MvNormal([2.0, 2.0], 1.0), pg.LinearMeasurement([[1.0, 0.0]]), 0.1, x_prim=xp, seed=99, chain=2)
    assert ref["status"] == 0 and len(smp) == K
    for k in range(K):
        _assert_same("A[%d]" % k, smp[k].A, ref["A"][k])
        _assert_same("Q[%d]" % k, smp[k].Q, ref["Q"][k])
        _assert_same("w_m1[%d]" % k, smp[k].w_m1, ref["w"][k])
        _assert_same("x_m1[%d]" % k, smp[k].x_m1.T, ref["x"][k])
        assert smp[k].u_m1[0] == u[0, -1]
    _assert_same("x_prim (in-place update)", xp, ref["x_prim"])


def test_batched_chains_match_single_chains():
    """n_chains batched on one handle == the same chains run one at a time (and == the oracle)."""
    basis, om, _ = _models()["known"]
    N, T, nc = 700, 25, 3
    u, y, _ = P.make_data(T, 14)
    V = 10 * np.ones(5)
    with _engine(basis, N, T, u, y, keep_w=False, n_chains=nc) as e:
        e.set_prior(V, 100 * np.eye(2), 10.0)
        for c in range(nc):
            e.set_state(np.zeros((2, 5)), 100 * np.eye(2), None, 1, chain=c)
        e.set_rng_philox(5, 10)
        A_s, Q_s, w_s, x_s, _ = e.particle_gibbs(2, 2, 0)
    for c in range(nc):
        ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, 5)),
                               100 * np.eye(2), V, 100 * np.eye(2), 10.0, 2, 2, 0, seed=5, chain=10 + c)
        for k in range(2):
            _assert_same("A", A_s[c, k].reshape((2, 5), order="F"), ref["A"][k])
            _assert_same("w", w_s[c, k], ref["w"][k])


# ------------------------------------------------------------------------------------------ rollout
@pytest.fixture(params=["warp", "cta"])
def rollout_kernel(request):
    """pgb_rollout picks the warp-per-scenario kernel whenever A_k fits 32 lanes (n_phi <= 640); the test hook
    pgb_test_rollout_kernel forces the CTA-per-scenario kernel (production: n_phi > 640) on the same shapes."""
    _lib.lib().pgb_test_rollout_kernel(1 if request.param == "warp" else 0)
    yield request.param
    _lib.lib().pgb_test_rollout_kernel(-1)


@pytest.mark.parametrize("model", ["known", "hilbert4"])
def test_rollout_bit_exact(model, rollout_kernel):
    rng = np.random.default_rng(41)
    if model == "known":
        basis, om = pg.KnownBasisVB(), O.known_model()
        n_x = 2
    else:
        basis = pg.HilbertGPBasis([1, 3, 3, 3, 3], [4.0, 5.0, 6.0, 7.0, 8.0])
        om = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L)
        n_x = 4
    K, H, N, n_phi = 37, 41, 30, basis.n_phi
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
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    out = pg.scenario_rollout(samples, basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=17)
    assert ref["status"] == 0
    for name in ("x0", "v", "e", "X", "Y"):
        _assert_same(name, out["raw"][name], ref[name])
    # reference layouts (optimal_control_Ipopt.jl:48,68,78,115,117)
    assert out["x_vec_0"].shape == (n_x * K,) and out["v_vec"].shape == (n_x, H, K) and out["e_vec"].shape == (1, H, K)
    assert out["X_init"].shape == (n_x * K, H + 1) and out["Y_init"].shape == (K, H)
    np.testing.assert_array_equal(out["X_init"][n_x * 3:n_x * 4, 5], ref["X"][3, 5])
    np.testing.assert_array_equal(out["v_vec"][:, 7, 2], ref["v"][2, 7])


@pytest.mark.parametrize("n_x,n_y,dims,N,K,H", [
    (4, 1, [1, 5, 5, 5, 5], 30, 9, 41),      # configs[4] shape: n_phi = 625, runs aligned with the last dimension
    (4, 2, [2, 3, 5, 3, 4], 33, 7, 17),      # n_phi = 360: 3 columns per run, last dimension 4 (unaligned runs)
    (4, 1, [1, 5, 5, 5, 4], 1, 5, 9),        # n_phi = 500: ragged last runs, a single particle
    (3, 1, [2, 5, 5, 5], 2500, 6, 12),       # n_x = 3; w_m1 longer than a warp chunk (tree + chunked walk)
    (2, 2, [5, 5, 5], 100, 11, 70),          # configs[1] basis; horizon longer than a warp
    (1, 1, [3, 7], 64, 8, 5),                # n_x = 1
    (2, 1, [1, 8, 8], 4097, 4, 6),           # N just above a power of two
    (2, 1, [5, 5, 5], 30, 1, 8),             # a single scenario (three of the CTA's four warps idle)
    (4, 1, [1, 5, 5, 5, 5], 2, 2, 3),        # two scenarios, two particles
])
def test_rollout_bit_exact_shapes(n_x, n_y, dims, N, K, H, rollout_kernel):
    rng = np.random.default_rng(47 + n_x + N)
    L = [4.0 + k for k in range(len(dims))]
    basis = pg.HilbertGPBasis(dims, L)
    om = O.hilbert_model(n_x, len(dims) - n_x, basis.count, basis.stride, basis.L, n_y=n_y)
    n_u, n_phi = len(dims) - n_x, basis.n_phi
    A = rng.normal(0, 0.3, (K, n_x, n_phi))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(n_x))(rng.normal(0, 0.2, (n_x, n_x))) for _ in range(K)])
    w = rng.uniform(size=(K, N)) ** 3
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, n_u))
    Cm = rng.normal(0, 1, (n_y, n_x))
    Br = rng.normal(0, 0.2, (n_y, n_y))
    R = Br @ Br.T + 0.05 * np.eye(n_y)
    u_init = rng.normal(0, 1, (n_u, H))
    ref = O.rollout(om, K, H, N, A, Q, w, x, um, Cm, np.linalg.cholesky(R), u_init, seed=23, k_offset=5)
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    out = pg.scenario_rollout(samples, basis, pg.LinearMeasurement(Cm), R, H, u_init=u_init, seed=23, k_offset=5)
    assert ref["status"] == 0
    for name in ("x0", "v", "e", "X", "Y"):
        _assert_same(name, out["raw"][name], ref[name])
    batch = pg.PGSampleBatch(A, Q, w, np.transpose(x, (0, 2, 1)), um)  # stacked arrays instead of the sample list
    out2 = pg.scenario_rollout(batch, basis, pg.LinearMeasurement(Cm), R, H, u_init=u_init, seed=23, k_offset=5)
    for name in ("x0", "v", "e", "X", "Y"):
        _assert_same(name + " (batch)", out2["raw"][name], ref[name])


def test_rollout_many_scenarios_persistent_grid(rollout_kernel):
    """More scenarios than resident warps: the warp kernel walks them with a grid stride; every one must still be the
    oracle's (checked on a strided subset) and the two kernels must agree on all of them."""
    rng = np.random.default_rng(5)
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
    om = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L)
    K, H, N, n_x = 3001, 11, 30, 4
    A = rng.normal(0, 0.05, (K, n_x, basis.n_phi))
    Bq = rng.normal(0, 0.1, (K, n_x, n_x))
    Q = Bq @ np.transpose(Bq, (0, 2, 1)) + 0.05 * np.eye(n_x)
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, 1))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    u_init = np.sin(np.arange(H))[None, :]
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    out = pg.scenario_rollout(samples, basis, pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=1)["raw"]
    sel = np.arange(0, K, 97)
    for k in sel:
        ref = O.rollout(om, 1, H, N, A[k:k + 1], Q[k:k + 1], w[k:k + 1], x[k:k + 1], um[k:k + 1], Cm, [[0.1]], u_init,
                        seed=1, k_offset=int(k))
        for name in ("x0", "v", "e", "X", "Y"):
            _assert_same("%s[%d]" % (name, k), out[name][k:k + 1], ref[name])
    import hashlib
    digest = hashlib.sha256(b"".join(out[n].tobytes() for n in ("x0", "v", "e", "X", "Y"))).hexdigest()
    seen = _ROLLOUT_DIGESTS.setdefault("many", digest)
    assert seen == digest, "warp-per-scenario and CTA-per-scenario kernels disagree"


_ROLLOUT_DIGESTS = {}


@pytest.mark.parametrize("model,K,k_n,T_test,N", [("known", 5, 3, 40, 30), ("hilbert", 3, 4, 300, 64),
                                                    ("hilbert4", 4, 2, 20, 30)])
def test_prediction_bit_exact(model, K, k_n, T_test, N, rollout_kernel):
    """pgopt_b200.test_prediction (particle_Gibbs.jl:253-314) against the oracle: simulated trajectories and outputs
    bit for bit, mean_rmse (device tree reduction) bit for bit."""
    rng = np.random.default_rng(61)
    if model == "known":
        basis, om, n_x = pg.KnownBasisVB(), O.known_model(), 2
    elif model == "hilbert":
        basis = pg.HilbertGPBasis([5, 5, 5], [10.0, 20.0, 20.0])
        om, n_x = O.hilbert_model(2, 1, basis.count, basis.stride, basis.L), 2
    else:
        basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
        om, n_x = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L), 4
    A = rng.normal(0, 0.2, (K, n_x, basis.n_phi))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(n_x))(rng.normal(0, 0.2, (n_x, n_x))) for _ in range(K)])
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, n_x))
    um = rng.normal(0, 1, (K, 1))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    u_test = rng.normal(0, 1, (1, T_test))
    y_test = rng.normal(0, 1, (1, T_test))
    ref = O.test_prediction(om, K, k_n, T_test, N, A, Q, w, x, um, Cm, [[0.1]], u_test, y_test, seed=29)
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    out = pg.test_prediction(samples, basis, pg.LinearMeasurement(Cm), 0.1, k_n, u_test, y_test, seed=29)
    assert ref["status"] == 0
    _assert_same("x_sim", out["raw"]["x_sim"], ref["x_sim"])
    _assert_same("y_sim", out["raw"]["y_sim"], ref["y_sim"])
    assert out["mean_rmse"] == ref["mean_rmse"]
    assert out["y_test_sim"].shape == (1, T_test, K * k_n) and out["x_test_sim"].shape == (n_x, T_test + 1, K * k_n)
    np.testing.assert_array_equal(out["y_test_sim"][:, 7, K + 1], ref["y_sim"][K + 1, 7])


# ------------------------------------------------------------------------------------------ other state / output dimensions
def _general_problem(n_x, n_y, T, seed):
    """A stable synthetic system with n_x states, 1 input and n_y outputs on a small Hilbert basis."""
    rng = np.random.default_rng(seed)
    dims = [2] + [3] * n_x if n_x < 4 else [1, 3, 3, 2, 2]
    L = [6.0] + [8.0 + d for d in range(n_x)]
    basis = pg.HilbertGPBasis(dims, L)
    om = O.hilbert_model(n_x, 1, basis.count, basis.stride, basis.L, n_y=n_y)
    A = rng.normal(0, 1.5, (n_x, basis.n_phi))
    Bq = rng.normal(0, 0.2, (n_x, n_x))
    Q = Bq @ Bq.T + 0.05 * np.eye(n_x)
    Cm = rng.normal(0, 1, (n_y, n_x))
    Br = rng.normal(0, 0.1, (n_y, n_y))
    R = Br @ Br.T + 0.05 * np.eye(n_y) if n_y > 1 else 0.1
    u = rng.normal(0, 1, (1, T))
    y = rng.normal(0, 1, (n_y, T))
    return basis, om, A, Q, Cm, R, u, y


@pytest.mark.parametrize("persistent", [1, 2, 0])
@pytest.mark.parametrize("n_x,n_y,N,T", [(1, 1, 700, 14), (3, 1, 1300, 12), (4, 1, 40, 30), (4, 2, 2100, 9), (2, 2, 1025, 10)])
def test_sweep_bit_exact_other_dimensions(n_x, n_y, N, T, persistent):
    """The sweep kernels are templated on n_x (1..4) and take n_y x n_y measurement noise: all instantiations, both the
    first (bootstrap) and a conditional sweep, against the oracle."""
    basis, om, A, Q, Cm, R, u, y = _general_problem(n_x, n_y, T, 100 + 10 * n_x + n_y)
    x0m, x0c = np.zeros(n_x), np.eye(n_x)
    s = P.injected_streams(N, T, n_x, basis.n_phi, 31)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    ref1 = O.sweep(om, N, T, u, y, Cm, R, x0m, x0c, A, Q, np.zeros((n_x, T)), True, st, 1)
    ref2 = O.sweep(om, N, T, u, y, Cm, R, x0m, x0c, A, Q, ref1["x_prim"], False, st, 2)
    with pg.ParticleGibbsEngine(basis, n_y, N, T, keep_w_history=True) as e:
        e.set_mode(persistent)
        e.set_data(u, y)
        e.set_measurement(pg.LinearMeasurement(Cm), R)
        e.set_init_dist(pg.MvNormal(x0m, 1.0))
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        for k, ref, xp in ((1, ref1, np.zeros((n_x, T))), (2, ref2, ref1["x_prim"])):
            e.set_state(A, Q, xp, k)
            e.sweep_filter()
            a, x, w = e.get_history()
            Phi, Psi, Sig, star = e.get_stats()
            _, _, xprim, _ = e.get_state()
            stat, nfb, _ = e.status(clear=True)
            _assert_same("ancestors k=%d" % k, a[1:, :N - 1], ref["a"][1:, :N - 1])
            if k > 1:
                _assert_same("ancestor of the conditioned particle", a[1:, N - 1], ref["a"][1:, N - 1])
            _assert_same("x_pf k=%d" % k, x, ref["x"])
            _assert_same("w k=%d" % k, w, ref["w"])
            assert star == ref["star"]
            _assert_same("x_prim", xprim, ref["x_prim"])
            _assert_same("Phi", Phi, ref["Phi"])
            _assert_same("Psi", Psi, ref["Psi"])
            _assert_same("Sigma", Sig, ref["Sigma"])
            assert stat == ref["status"] and nfb == 0


# ------------------------------------------------------------------------------------------ replay mode (no x history)
@pytest.mark.parametrize("persistent", [1, 2, 0])
@pytest.mark.parametrize("model,N,T,philox", [("known", 30, 150, True), ("known", 2500, 40, False), ("hilbert", 700, 30, True),
                                               ("hilbert", 33, 60, False)])
def test_replay_mode_reproduces_the_traced_trajectory(model, N, T, philox, persistent):
    """x_history="replay" keeps two rows of x and recomputes the ancestor lineage of the new conditioned trajectory
    (particle_Gibbs.jl:213-218): x_prim, the MNIW statistics, the kept sample and all ancestors must be those of the
    oracle, over a first sweep and two conditional sweeps (the lineage passes through the pinned slot N sometimes)."""
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 17)
    s = P.injected_streams(N, T, 2, basis.n_phi, 27)
    kw = dict(philox=True, seed=321, chain=4) if philox else dict(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    with pg.ParticleGibbsEngine(basis, 1, N, T, x_history="replay") as e:
        e.set_mode(persistent)
        e.set_data(u, y)
        e.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        e.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        if philox:
            e.set_rng_philox(321, 4)
        else:
            e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        for k in (1, 2, 3):
            ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, k == 1, O.make_streams(**kw), k)
            e.set_state(A, P.Q_TRUE, xp, k)
            e.sweep_filter()
            a, _, _ = e.get_history(want_w=False, want_x=False)
            Phi, Psi, Sig, star = e.get_stats()
            _, _, xprim, _ = e.get_state()
            smp = e.get_sample()
            stat, nfb, _ = e.status(clear=True)
            assert star == ref["star"] and stat == ref["status"] == 0
            _assert_same("ancestors k=%d" % k, a[1:, :N - 1], ref["a"][1:, :N - 1])
            _assert_same("x_prim k=%d" % k, xprim, ref["x_prim"])
            _assert_same("Phi", Phi, ref["Phi"])
            _assert_same("Psi", Psi, ref["Psi"])
            _assert_same("Sigma", Sig, ref["Sigma"])
            _assert_same("w_m1", smp.w_m1, ref["w_last"])
            _assert_same("x_m1", smp.x_m1.T, ref["x_last"])
            xp = ref["x_prim"]
        with pytest.raises(pg.PgoptError):
            e.get_history(want_w=False)  # the x history does not exist in this mode


def test_replay_mode_whole_sampler_matches_store_mode():
    """particle_Gibbs end to end (filter + trace + statistics + MNIW, K kept samples) in replay mode == oracle."""
    basis, om, _ = _models()["hilbert"]
    N, T, K = 64, 40, 3
    u, y, _ = P.make_data(T, 13)
    V = pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3), 100 ** 2 / (8 * np.pi ** 4))
    ref = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), np.zeros((2, 125)), 100 * np.eye(2), V,
                           100 * np.eye(2), 10.0, K, 2, 1, seed=5, chain=0)
    smp = pg.particle_Gibbs(u, y, K, 2, 1, N, basis, 100 * np.eye(2), 10.0, 100 * np.eye(2), V, np.zeros((2, 125)),
                            pg.MvNormal([2.0, 2.0], 1.0), pg.LinearMeasurement([[1.0, 0.0]]), 0.1, x_prim=np.zeros((2, T)),
                            seed=5, chain=0, x_history="replay")
    for k in range(K):
        _assert_same("A[%d]" % k, smp[k].A, ref["A"][k])
        _assert_same("Q[%d]" % k, smp[k].Q, ref["Q"][k])
        _assert_same("x_m1[%d]" % k, smp[k].x_m1.T, ref["x"][k])


def test_rollout_shards_reproduce_the_unsharded_rollout():
    """Scenario sharding (pgopt_b200.multi.rollout_sharded): shard [lo, hi) with k_offset = lo gives rows lo..hi of the
    full rollout, and both equal the oracle."""
    rng = np.random.default_rng(43)
    basis, om = pg.KnownBasisVB(), O.known_model()
    K, H, N = 23, 12, 30
    A = rng.normal(0, 0.3, (K, 2, 5))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(2))(rng.normal(0, 0.2, (2, 2))) for _ in range(K)])
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, N, 2))
    um = rng.normal(0, 1, (K, 1))
    u_init = rng.normal(0, 1, (1, H))
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(K)]
    g = pg.LinearMeasurement([[1.0, 0.0]])
    full = pg.scenario_rollout(samples, basis, g, 0.1, H, u_init=u_init, seed=3)["raw"]
    ref = O.rollout(om, K, H, N, A, Q, w, x, um, [[1.0, 0.0]], [[0.1]], u_init, seed=3)
    from pgopt_b200 import multi
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = multi.shard_range(K, r, world)
            parts.append(pg.scenario_rollout(samples[lo:hi], basis, g, 0.1, H, u_init=u_init, seed=3, k_offset=lo)["raw"])
        for name in ("x0", "v", "e", "X", "Y"):
            cat = np.concatenate([p[name] for p in parts], axis=0)
            _assert_same("%s world=%d" % (name, world), cat, full[name])
            _assert_same("%s vs oracle" % name, cat, ref[name])


# ---- scenario screening of the greedy support search (optimal_control_Ipopt.jl:334-338) ------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("K,H,n_y,ties", [(1, 7, 1, False), (200, 41, 1, False), (333, 41, 2, True), (5000, 100, 3, True),
                                          (3, 1, 1, False)])
def test_scenario_constraint_max_bit_exact(K, H, n_y, ties):
    from test_oracle_kat import scenario_screening_case
    y, y_min, y_max = scenario_screening_case(K, H, n_y, seed=K + H, ties=ties)
    if H == 1:
        y_max[0, 0] = 0.5
    ref = O.scenario_constraint_max(np.transpose(y, (2, 1, 0)), y_min, y_max)
    h, order = pg.scenario_constraint_max(y, y_min, y_max)  # reference layout (n_y, H, K)
    assert np.array_equal(h, ref["h_max"]) and np.array_equal(order, ref["order"])
    h2, order2 = pg.scenario_constraint_max(np.transpose(y, (2, 1, 0)), y_min, y_max)  # rollout layout (K, H, n_y)
    assert np.array_equal(h2, h) and np.array_equal(order2, order)


@pytest.mark.gpu
def test_scenario_constraint_max_nan_and_errors():
    from test_oracle_kat import scenario_screening_case
    y, y_min, y_max = scenario_screening_case(64, 41, 1, seed=9)
    y[0, 22, 5] = np.nan  # inside the bounded window: maximum() propagates NaN, sortperm(rev=true) puts it first
    y[0, 0, 6] = np.nan   # outside every finite bound: not evaluated
    ref = O.scenario_constraint_max(np.transpose(y, (2, 1, 0)), y_min, y_max)
    h, order = pg.scenario_constraint_max(y, y_min, y_max)
    assert np.isnan(h[5]) and not np.isnan(h[6]) and order[0] == 6
    assert np.array_equal(h, ref["h_max"], equal_nan=True) and np.array_equal(order, ref["order"])
    with pytest.raises(pg.PgoptError):
        pg.scenario_constraint_max(y, np.full((1, 41), -np.inf), np.full((1, 41), np.inf))


@pytest.mark.gpu
def test_scenario_constraint_max_full_size_properties():
    """configs[4] size (K = 1e5, H = 41): sortedness, permutation, and equality with a vectorised NumPy maximum."""
    K, H = 100000, 41
    rng = np.random.default_rng(3)
    Y = rng.normal(size=(K, H, 1)) * 3
    y_min = np.full((1, H), -np.inf); y_min[0, 20:26] = 2.0
    y_max = np.full((1, H), np.inf); y_max[0, 30:] = 6.0
    h, order = pg.scenario_constraint_max(Y, y_min, y_max)
    expect = np.maximum((2.0 - Y[:, 20:26, 0]).max(axis=1), (Y[:, 30:, 0] - 6.0).max(axis=1))
    assert np.array_equal(h, expect)
    assert np.array_equal(np.sort(order), np.arange(1, K + 1)) and np.all(np.diff(h[order - 1]) <= 0)


# ------------------------------------------------------------------------------------------ reference-held vectors
import _reffix as RF  # noqa: E402


@pytest.mark.skipif(not RF.fixtures(), reason=RF.SKIP_REASON)
@pytest.mark.parametrize("path", RF.fixtures() or ["-"])
def test_gpu_against_reference_held_vectors(path):
    """The CUDA path fed the streams the unmodified Julia reference consumed (julia/make_fixtures.jl): final particles
    and weights of every sweep and the traced trajectory within 1e-10 relative of the real Julia run."""
    g = RF.load(path)
    _, hil = RF.oracle_model(g)
    basis = pg.KnownBasisVB() if hil is None else pg.HilbertGPBasis(hil[0], hil[1])
    N, T, K = int(RF.scalar(g, "N")), int(RF.scalar(g, "T")), int(RF.scalar(g, "K"))
    xp = np.zeros((basis.n_x, T))
    with pg.ParticleGibbsEngine(basis, 1, N, T) as e:
        e.set_data(g["u"], g["y"])
        e.set_measurement(pg.LinearMeasurement(g["C"]), RF.scalar(g, "R"))
        e.set_init_dist(pg.MvNormal(np.ravel(g["x0_mean"]), g["x0_chol"] @ g["x0_chol"].T))
        for k in range(1, K + 1):
            s, A, Q = RF.sweep_inputs(g, k)
            e.set_state(A, Q, xp, k)
            e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
            e.sweep_filter()
            smp = e.get_sample()
            np.testing.assert_allclose(smp.w_m1, np.ravel(g["w_m1_%d" % k]), rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(smp.x_m1, g["x_m1_%d" % k], rtol=1e-10, atol=1e-12)
            xp = e.get_state()[2]
    np.testing.assert_allclose(xp, g["x_prim_final"], rtol=1e-10, atol=1e-12)
