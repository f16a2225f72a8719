"""ctypes loader for the CPU parity oracle (oracle/libpg_oracle.so).

Test infrastructure only: nothing under pgopt_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORC_MAX_Z = 8
BASIS_KNOWN_VB, BASIS_HILBERT_GP = 0, 1


class OrcModel(C.Structure):
    _fields_ = [("basis", C.c_int32), ("n_x", C.c_int32), ("n_u", C.c_int32), ("n_y", C.c_int32),
                ("n_phi", C.c_int32), ("count", C.c_int32 * ORC_MAX_Z), ("stride", C.c_int32 * ORC_MAX_Z),
                ("L", C.c_double * ORC_MAX_Z)]


class OrcStreams(C.Structure):
    _fields_ = [("philox", C.c_int32), ("seed", C.c_uint64), ("chain", C.c_uint32),
                ("Z0", C.c_void_p), ("Ures", C.c_void_p), ("Z", C.c_void_p), ("Uanc", C.c_void_p),
                ("Ustar", C.c_double), ("bartlett", C.c_void_p), ("Xmn", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "libpg_oracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("pg_oracle.c", "pg_oracle.h")] + [
            os.path.join(ROOT, "pgopt_b200", "csrc", "pgb_math.h"), os.path.join(ROOT, "pgopt_b200", "csrc", "pgb_mathf.h")]
        if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        _lib = C.CDLL(so)
        _lib.orc_pairwise_sum.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def known_model():
    m = OrcModel()
    m.basis, m.n_x, m.n_u, m.n_y, m.n_phi = BASIS_KNOWN_VB, 2, 1, 1, 5
    return m


def hilbert_model(n_x, n_u, count, stride, L, n_y=1):
    m = OrcModel()
    m.basis, m.n_x, m.n_u, m.n_y = BASIS_HILBERT_GP, n_x, n_u, n_y
    n_phi = 1
    for k, (c, s, l) in enumerate(zip(count, stride, L)):
        m.count[k], m.stride[k], m.L[k] = int(c), int(s), float(l)
        n_phi *= int(c)
    m.n_phi = n_phi
    return m


def vfun(name, x):
    x = f64(x)
    y = np.empty_like(x)
    getattr(lib(), "orc_v" + name)(_p(x), _p(y), C.c_int64(x.size))
    return y


def normal_pairs(seed, purpose, chain, sweep, t, npairs):
    z = np.empty(2 * npairs)
    lib().orc_vnormal_pairs(C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(chain), C.c_uint32(sweep),
                            C.c_uint32(t), C.c_int64(npairs), _p(z))
    return z


def pairwise_sum(v):
    v = f64(v)
    return lib().orc_pairwise_sum(_p(v), C.c_int64(v.size))


def systematic_resampling(W, n_draw, rnd):
    W = f64(W)
    idx = np.zeros(n_draw, dtype=np.int64)
    st = lib().orc_systematic_resampling(_p(W), C.c_int64(W.size), C.c_int64(n_draw), C.c_double(rnd), _p(idx))
    return idx, st


def phi(model, x, u):
    x, u = f64(x), f64(u)
    out = np.empty(model.n_phi)
    lib().orc_phi(C.byref(model), _p(x), _p(u), _p(out))
    return out


def make_streams(philox=True, seed=0, chain=0, Z0=None, Ures=None, Z=None, Uanc=None, Ustar=0.0,
                 bartlett=None, Xmn=None):
    s = OrcStreams()
    s.philox, s.seed, s.chain = int(philox), seed, chain
    keep = []
    for name, arr in (("Z0", Z0), ("Ures", Ures), ("Z", Z), ("Uanc", Uanc), ("bartlett", bartlett), ("Xmn", Xmn)):
        if arr is not None:
            arr = f64(np.asarray(arr, dtype=np.float64).ravel(order="F")) if name in ("bartlett", "Xmn") else f64(arr)
            keep.append(arr)
            setattr(s, name, arr.ctypes.data)
    s.Ustar = float(Ustar)
    s._keep = keep
    return s


def sweep(model, N, T, u, y, Cm, R, x0_mean, x0_chol, A, Q, x_prim, first, streams, sweep_index, hist=True, f32=False):
    """f32: the F32 precision mode (orc_sweep_f32).  Returns dict with a (T,N) 1-based, x (T,N,n_x), w (T,N), w_last, x_last (N,n_x), Phi, Psi, Sigma,
    star, x_prim (n_x,T) Fortran-order arrays where the reference uses matrices."""
    n_x, n_phi = model.n_x, model.n_phi
    u, y, Cm, R = f64(np.asfortranarray(u).ravel(order="F")), f64(np.asfortranarray(y).ravel(order="F")), \
        f64(np.asfortranarray(Cm).ravel(order="F")), f64(np.atleast_2d(R).ravel(order="F"))
    A = f64(np.asfortranarray(A).ravel(order="F"))
    Q = f64(np.asfortranarray(Q).ravel(order="F"))
    xp = f64(np.asfortranarray(x_prim).ravel(order="F")).copy()
    x0m, x0c = f64(x0_mean), f64(np.asfortranarray(x0_chol).ravel(order="F"))
    a = np.zeros((T, N), dtype=np.int64) if hist else None
    xh = np.zeros((T, N, n_x)) if hist else None
    wh = np.zeros((T, N)) if hist else None
    w_last, x_last = np.zeros(N), np.zeros((N, n_x))
    Phi, Psi, Sig = np.zeros(n_x * n_x), np.zeros(n_x * n_phi), np.zeros(n_phi * n_phi)
    star = C.c_int64(0)
    fn = lib().orc_sweep_f32 if f32 else lib().orc_sweep
    st = fn(C.byref(model), C.c_int64(N), C.c_int64(T), _p(u), _p(y), _p(Cm), _p(R), _p(x0m), _p(x0c),
            _p(A), _p(Q), _p(xp), C.c_int(int(first)), C.byref(streams), C.c_uint32(sweep_index),
            _p(a), _p(xh), _p(wh), _p(w_last), _p(x_last), _p(Phi), _p(Psi), _p(Sig), C.byref(star))
    return dict(status=st, a=a, x=xh, w=wh, w_last=w_last, x_last=x_last,
                Phi=Phi.reshape((n_x, n_x), order="F"), Psi=Psi.reshape((n_x, n_phi), order="F"),
                Sigma=Sig.reshape((n_phi, n_phi), order="F"), star=star.value,
                x_prim=xp.reshape((n_x, T), order="F"))


def mniw_sample(n_x, n_phi, Phi, Psi, Sigma, V_diag, Lambda_Q, ell_Q, Tm1, streams, sweep_index):
    """V_diag: the diagonal (n_phi,) of V, or a dense symmetric positive definite V (n_phi, n_phi)"""
    A, Q = np.zeros(n_x * n_phi), np.zeros(n_x * n_x)
    args = [f64(np.asfortranarray(v).ravel(order="F")) for v in (Phi, Psi, Sigma, V_diag, Lambda_Q)]
    fn = lib().orc_mniw_sample_dense if np.ndim(V_diag) == 2 else lib().orc_mniw_sample
    st = fn(C.c_int(n_x), C.c_int(n_phi), *[_p(v) for v in args], C.c_double(ell_Q),
            C.c_int64(Tm1), C.byref(streams), C.c_uint32(sweep_index), _p(A), _p(Q))
    return A.reshape((n_x, n_phi), order="F"), Q.reshape((n_x, n_x), order="F"), st


def spd_inverse(V):
    """orc_spd_inverse: I / V of a dense symmetric positive definite V (the oracle's and the library's definition)"""
    V = np.asarray(V, dtype=np.float64)
    n = V.shape[0]
    out = np.zeros(n * n)
    rc = lib().orc_spd_inverse(_p(f64(np.asfortranarray(V).ravel(order="F"))), C.c_int(n), _p(out))
    return out.reshape((n, n), order="F"), rc


def particle_gibbs(model, N, T, u, y, Cm, R, x0_mean, x0_chol, A_init, Q_init, V_diag, Lambda_Q, ell_Q,
                   K, K_b, k_d, seed, chain, x_prim=None, keep_particles=True):
    n_x, n_phi = model.n_x, model.n_phi
    fl = lambda v: f64(np.asfortranarray(v).ravel(order="F"))
    xp = np.zeros(n_x * T) if x_prim is None else fl(x_prim).copy()
    A_s, Q_s = np.zeros((K, n_x * n_phi)), np.zeros((K, n_x * n_x))
    w_s = np.zeros((K, N)) if keep_particles else None
    x_s = np.zeros((K, N, n_x)) if keep_particles else None
    fn = lib().orc_particle_gibbs_dense if np.ndim(V_diag) == 2 else lib().orc_particle_gibbs
    st = fn(C.byref(model), C.c_int64(N), C.c_int64(T), _p(fl(u)), _p(fl(y)), _p(fl(Cm)),
            _p(fl(np.atleast_2d(R))), _p(f64(x0_mean)), _p(fl(x0_chol)), _p(fl(A_init)),
            _p(fl(Q_init)), _p(fl(V_diag)), _p(fl(Lambda_Q)), C.c_double(ell_Q),
            C.c_int64(K), C.c_int64(K_b), C.c_int64(k_d), C.c_uint64(seed), C.c_uint32(chain),
            _p(xp), _p(A_s), _p(Q_s), _p(w_s), _p(x_s))
    return dict(status=st, A=A_s.reshape((K, n_phi, n_x)).transpose(0, 2, 1), Q=Q_s.reshape((K, n_x, n_x)).transpose(0, 2, 1),
                w=w_s, x=x_s, x_prim=xp.reshape((n_x, T), order="F"))


def rollout(model, K, H, N, A, Q, w_m1, x_m1, u_m1, Cm, Le, u_init, seed, k_offset=0):
    """A (K,n_x,n_phi), Q (K,n_x,n_x), w_m1 (K,N), x_m1 (K,N,n_x), u_m1 (K,n_u), u_init (n_u,H)."""
    n_x, n_y = model.n_x, model.n_y
    Af = f64(np.transpose(A, (0, 2, 1)))  # per scenario column-major
    Qf = f64(np.transpose(Q, (0, 2, 1)))
    fl = lambda v: f64(np.asfortranarray(v).ravel(order="F"))
    x0, v, e = np.zeros((K, n_x)), np.zeros((K, H, n_x)), np.zeros((K, H, n_y))
    X, Y = np.zeros((K, H + 1, n_x)), np.zeros((K, H, n_y))
    st = lib().orc_rollout(C.byref(model), C.c_int64(K), C.c_int64(H), C.c_int64(N), _p(Af), _p(Qf), _p(f64(w_m1)),
                           _p(f64(x_m1)), _p(f64(u_m1)), _p(fl(Cm)), _p(fl(np.atleast_2d(Le))), _p(fl(u_init)),
                           C.c_uint64(seed), C.c_uint32(k_offset), _p(x0), _p(v), _p(e), _p(X), _p(Y))
    return dict(status=st, x0=x0, v=v, e=e, X=X, Y=Y)


def rollout_supplied(model, K, H, N, A, Q, w_m1, x_m1, u_m1, Cm, Le, u_init, seed, k_offset=0, x0_in=None, v_in=None,
                     e_in=None):
    """orc_rollout_supplied: x0_in (K,n_x) / v_in (K,H,n_x) / e_in (K,H,n_y) replace the draws (None = draw)."""
    n_x, n_y = model.n_x, model.n_y
    Af = f64(np.transpose(A, (0, 2, 1)))
    Qf = f64(np.transpose(Q, (0, 2, 1)))
    fl = lambda v: f64(np.asfortranarray(v).ravel(order="F"))
    opt = lambda a: None if a is None else f64(a)
    x0i, vi, ei = opt(x0_in), opt(v_in), opt(e_in)
    x0, v, e = np.zeros((K, n_x)), np.zeros((K, H, n_x)), np.zeros((K, H, n_y))
    X, Y = np.zeros((K, H + 1, n_x)), np.zeros((K, H, n_y))
    st = lib().orc_rollout_supplied(C.byref(model), C.c_int64(K), C.c_int64(H), C.c_int64(N), _p(Af), _p(Qf),
                                    _p(f64(w_m1)), _p(f64(x_m1)), _p(f64(u_m1)), _p(fl(Cm)), _p(fl(np.atleast_2d(Le))),
                                    _p(fl(u_init)), C.c_uint64(seed), C.c_uint32(k_offset),
                                    None if x0i is None else _p(x0i), None if vi is None else _p(vi),
                                    None if ei is None else _p(ei), _p(x0), _p(v), _p(e), _p(X), _p(Y))
    return dict(status=st, x0=x0, v=v, e=e, X=X, Y=Y)


def draw_xT(w_m1, x_m1, seed, cid):
    """x_m1 (N, n_x); returns x_m1[star] for star ~ systematic_resampling(w_m1, 1) with scenario stream cid"""
    x_m1 = f64(x_m1)
    N, n_x = x_m1.shape
    out = np.zeros(n_x)
    st = lib().orc_draw_xT(C.c_int64(N), C.c_int(n_x), _p(f64(w_m1)), _p(x_m1), C.c_uint64(seed), C.c_uint32(cid), _p(out))
    assert st == 0
    return out


def dynamics(model, A, x_vec, u, v=None, t=0, variant=0):
    """orc_dynamics: A (K,n_x,n_phi), x_vec (K,n_x), u (n_u,), v (K,H,n_x) or None -> x_next (K,n_x), dfdx (K,n_x,n_x),
    dfdu (K,n_x,n_u) (matrices as NumPy row-major views of the column-major blocks: [k, row, col])"""
    K, n_x, n_phi = A.shape
    n_u = model.n_u
    Af = f64(np.transpose(A, (0, 2, 1)))
    xn, jx, ju = np.zeros((K, n_x)), np.zeros((K, n_x, n_x)), np.zeros((K, n_u, n_x))
    vv = None if v is None else f64(v)
    H = 0 if v is None else v.shape[1]
    st = lib().orc_dynamics(C.byref(model), C.c_int64(K), _p(Af), _p(f64(x_vec)), _p(f64(u)), None if vv is None else _p(vv),
                            C.c_int64(H), C.c_int64(t), C.c_int(variant), _p(xn), _p(jx), _p(ju))
    assert st == 0
    return xn, np.transpose(jx, (0, 2, 1)), np.transpose(ju, (0, 2, 1))


def test_prediction(model, K, k_n, T_test, N, A, Q, w_m1, x_m1, u_m1, Cm, Le, u_test, y_test, seed):
    """orc_test_prediction: x_sim (K*k_n, T_test+1, n_x), y_sim (K*k_n, T_test, n_y), mean_rmse."""
    n_x, n_y = model.n_x, model.n_y
    Af = f64(np.transpose(A, (0, 2, 1)))
    Qf = f64(np.transpose(Q, (0, 2, 1)))
    fl = lambda v: f64(np.asfortranarray(v).ravel(order="F"))
    S = K * k_n
    xs, ys = np.zeros((S, T_test + 1, n_x)), np.zeros((S, T_test, n_y))
    rmse = C.c_double(0.0)
    st = lib().orc_test_prediction(C.byref(model), C.c_int64(K), C.c_int64(k_n), C.c_int64(T_test), C.c_int64(N), _p(Af),
                                   _p(Qf), _p(f64(w_m1)), _p(f64(x_m1)), _p(f64(u_m1)), _p(fl(Cm)),
                                   _p(fl(np.atleast_2d(Le))), _p(fl(np.atleast_2d(u_test))), _p(fl(np.atleast_2d(y_test))),
                                   C.c_uint64(seed), _p(xs), _p(ys), C.byref(rmse))
    return dict(status=st, x_sim=xs, y_sim=ys, mean_rmse=rmse.value)



def scenario_constraint_max(Y, y_min, y_max):
    """orc_scenario_constraint_max: Y (K, H, n_y) scenario-major; y_min / y_max (n_y, H).  Returns h_max (K,), order (K,) 1-based."""
    Y = f64(Y)
    K, H, n_y = Y.shape
    fl = lambda v: f64(np.asfortranarray(np.atleast_2d(v)).ravel(order="F"))
    h = np.zeros(K)
    order = np.zeros(K, dtype=np.int64)
    st = lib().orc_scenario_constraint_max(C.c_int64(K), C.c_int64(H), C.c_int32(n_y), _p(Y), _p(fl(y_min)), _p(fl(y_max)),
                                           _p(h), _p(order))
    return dict(status=st, h_max=h, order=order)
