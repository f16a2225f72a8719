"""pgb_mathf.h (the bit-defined FP32 functions of the F32 precision mode) on the host: accuracy against float64
references — the "stated FP32 tolerance" of the F32 mode is FP32 round-off — and sanity of the FP32 Box-Muller stream.
The F32 mode of the oracle sweep itself: it must track the FP64 sweep within that tolerance where no index flips."""
import ctypes as C

import numpy as np

import _oracle as O
import _problems as P


def _f(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    O.lib().orc_vmathf(C.c_int(fn), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int64(x.size))
    return y


def _ulp(y, ref):
    u = np.spacing(np.abs(ref.astype(np.float32))).astype(np.float64)
    return float(np.max(np.abs(y.astype(np.float64) - ref) / u))


def test_expf_accuracy_and_edges():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-87, 88, 400000), rng.uniform(-30, 0, 400000), rng.uniform(-1, 1, 100000)]).astype(np.float32)
    assert _ulp(_f(0, x), np.exp(x.astype(np.float64))) < 1.1
    y = _f(0, np.array([0.0, -87.0, -87.5, -1e30, 88.4, 88.6, np.inf, -np.inf, np.nan], dtype=np.float32))
    assert y[0] == 1.0 and y[1] > 0 and y[2] == 0 and y[3] == 0 and np.isfinite(y[4]) and np.isinf(y[5]) and np.isinf(y[6])
    assert y[7] == 0 and np.isnan(y[8])


def test_logf_accuracy_on_the_box_muller_grid():
    rng = np.random.default_rng(2)
    x = ((rng.integers(0, 2 ** 23, 500000) + 0.5) * 2.0 ** -23).astype(np.float32)
    x = np.concatenate([x, np.array([2.0 ** -24, 1 - 2.0 ** -24, 0.5, 0.70710678, 0.7071068], dtype=np.float32)])
    assert _ulp(_f(1, x), np.log(x.astype(np.float64))) < 1.0


def test_sincosf_accuracy():
    rng = np.random.default_rng(3)
    x = rng.uniform(-20, 20, 500000).astype(np.float32)  # the basis-function arguments: pi*j/(2L)*(z + L) <= 5 pi
    assert _ulp(_f(2, x), np.sin(x.astype(np.float64))) < 1.7
    assert _ulp(_f(3, x), np.cos(x.astype(np.float64))) < 1.7
    for hi in (1000.0, 32768.0, 1e7):  # absolute error stays at FP32 round-off, also beyond the Cody-Waite range
        x = rng.uniform(-hi, hi, 300000).astype(np.float32)
        assert np.max(np.abs(_f(2, x) - np.sin(x.astype(np.float64)))) < 1.6 * 2.0 ** -24 * 1.05
        assert np.max(np.abs(_f(3, x) - np.cos(x.astype(np.float64)))) < 1.6 * 2.0 ** -24 * 1.05


def test_normals_f32_moments_and_counter_semantics():
    n, nx = 300000, 4
    z = np.empty((n, nx), dtype=np.float32)
    O.lib().orc_vnormal_pairs_f32(C.c_uint64(42), C.c_uint32(2), C.c_uint32(0), C.c_uint32(1), C.c_uint32(5), C.c_int64(n),
                                  C.c_int(nx), z.ctypes.data_as(C.c_void_p))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.05
    assert np.max(np.abs(z)) < 5.8  # the 2^-23 uniform grid bounds the tails: sqrt(-2 log 2^-24) = 5.77
    c = np.corrcoef(z.T)
    assert np.max(np.abs(c - np.eye(nx))) < 0.01
    z2 = np.empty((50, 2), dtype=np.float32)
    O.lib().orc_vnormal_pairs_f32(C.c_uint64(42), C.c_uint32(2), C.c_uint32(0), C.c_uint32(1), C.c_uint32(5), C.c_int64(50),
                                  C.c_int(2), z2.ctypes.data_as(C.c_void_p))
    assert np.array_equal(z2, z[:50, :2])  # dims (0, 1) come from words (0, 1) whatever n_x is


def test_oracle_f32_sweep_tracks_f64_with_injected_streams():
    """Same injected streams, both precisions: the F32 sweep stores FP32 values, and until the first ancestor index
    differs (a discontinuity, not an error) states agree to FP32 round-off accumulated over the steps."""
    N, T = 40, 30
    u, y, _ = P.make_data(T, 5)
    s = P.injected_streams(N, T, 2, 5, 9)
    st = O.make_streams(philox=False, **{k: s[k] for k in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    A = P.plausible_A_known()
    args = (O.known_model(), N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, np.zeros((2, T)), True, st, 1)
    r64 = O.sweep(*args)
    r32 = O.sweep(*args, f32=True)
    assert r32["status"] == 0
    assert np.array_equal(r32["x"], r32["x"].astype(np.float32).astype(np.float64))  # FP32 values in FP64 storage
    same = np.all(r32["a"] == r64["a"], axis=1)
    t_same = int(np.argmin(same)) if not same.all() else T
    assert t_same >= 3
    np.testing.assert_allclose(r32["x"][:t_same], r64["x"][:t_same], rtol=0, atol=2e-4)
    np.testing.assert_allclose(r32["w"][:t_same], r64["w"][:t_same], rtol=0, atol=2e-4)
    assert abs(r32["w"].sum(1) - 1).max() < 1e-12  # the normalisation is FP64
