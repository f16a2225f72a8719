"""Host-side mirror of the reference's Julia API for the hot path, over libpgopt_b200.so.

Same names, argument order and meaning as Julia/src/particle_Gibbs.jl and
Julia/src/optimal_control_Ipopt.jl; arrays are NumPy, column-major ("F") where the reference passes
matrices, indices 1-based like Julia's.  Julia closures cannot cross a C ABI, so `phi` and `g` are
descriptor objects (KnownBasisVB / HilbertGPBasis, LinearMeasurement); passing an arbitrary callable
raises TypeError — there is no CPU fallback path.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PgbConfig, PgoptError, check, lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    """contiguous float64 vector holding `a` in column-major order"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


# ----------------------------------------------------------------------------------------------
# descriptors
# ----------------------------------------------------------------------------------------------
@dataclass
class PG_sample:
    """Julia/src/PGopt.jl:23-29"""
    A: np.ndarray      # (n_x, n_phi)
    Q: np.ndarray      # (n_x, n_x)
    w_m1: np.ndarray   # (N,)
    x_m1: np.ndarray   # (n_x, N)
    u_m1: np.ndarray   # (n_u,)


def _prior_V(V):
    """(array for the C ABI, dense?): the diagonal of V when V is a vector or a diagonal matrix, else the column-major
    dense matrix (symmetric: either triangle order gives the same bytes)."""
    V = np.asarray(V, dtype=np.float64)
    if V.ndim == 2:
        if V.shape[0] != V.shape[1]:
            raise ValueError("V must be square")
        if np.count_nonzero(V - np.diag(np.diag(V))) == 0:
            return np.diag(V).copy(), False
        if not np.array_equal(V, V.T):
            raise ValueError("V must be symmetric")
        return np.ascontiguousarray(V), True
    return V, False


class PGSampleBatch:
    """K posterior samples as stacked arrays — the same content as a Vector{PG_sample} (PGopt.jl:23-29), without one
    Python object per sample: A (K, n_x, n_phi), Q (K, n_x, n_x), w_m1 (K, N), x_m1 (K, n_x, N), u_m1 (K, n_u).
    `scenario_rollout` / `test_prediction` take it in place of the list and skip the per-sample packing (2 s of Python
    at K = 10^5).  Indexing gives the PG_sample view of one entry."""

    def __init__(self, A, Q, w_m1, x_m1, u_m1):
        self.A, self.Q = np.asarray(A, dtype=np.float64), np.asarray(Q, dtype=np.float64)
        self.w_m1, self.x_m1 = np.asarray(w_m1, dtype=np.float64), np.asarray(x_m1, dtype=np.float64)
        self.u_m1 = np.asarray(u_m1, dtype=np.float64).reshape(self.A.shape[0], -1)
        K = self.A.shape[0]
        if not (self.Q.shape[0] == self.w_m1.shape[0] == self.x_m1.shape[0] == K):
            raise ValueError("all arrays must have K entries along the first axis")

    def __len__(self):
        return self.A.shape[0]

    def __getitem__(self, k):
        return PG_sample(self.A[k], self.Q[k], self.w_m1[k], self.x_m1[k], self.u_m1[k])

    @classmethod
    def from_samples(cls, samples):
        return cls(np.stack([s.A for s in samples]), np.stack([s.Q for s in samples]),
                   np.stack([np.ravel(s.w_m1) for s in samples]), np.stack([s.x_m1 for s in samples]),
                   np.stack([np.ravel(s.u_m1) for s in samples]))


def _pack_samples(PG_samples):
    """(A, Q, w, x, um, N) in the C ABI's layout: per scenario column-major A (n_x x n_phi), Q, x_m1 (n_x x N)."""
    if isinstance(PG_samples, PGSampleBatch):
        b = PG_samples
        K = len(b)
        A = np.ascontiguousarray(np.transpose(b.A, (0, 2, 1))).reshape(K, -1)
        Q = np.ascontiguousarray(np.transpose(b.Q, (0, 2, 1))).reshape(K, -1)
        x = np.ascontiguousarray(np.transpose(b.x_m1, (0, 2, 1))).reshape(K, -1)
        return A, Q, np.ascontiguousarray(b.w_m1), x, np.ascontiguousarray(b.u_m1), b.w_m1.shape[1]
    A = np.ascontiguousarray(np.stack([_f(s.A) for s in PG_samples]))
    Q = np.ascontiguousarray(np.stack([_f(s.Q) for s in PG_samples]))
    w = np.ascontiguousarray(np.stack([np.asarray(s.w_m1, dtype=np.float64).ravel() for s in PG_samples]))
    x = np.ascontiguousarray(np.stack([_f(s.x_m1) for s in PG_samples]))
    um = np.ascontiguousarray(np.stack([np.asarray(s.u_m1, dtype=np.float64).ravel() for s in PG_samples]))
    return A, Q, w, x, um, PG_samples[0].w_m1.size


class KnownBasisVB:
    """phi of Julia/examples/PG_OCP_known_basis_functions_Altro.jl:34 (n_x = 2, n_u = 1, n_phi = 5)."""
    n_x, n_u, n_phi = 2, 1, 5

    def fill(self, cfg):
        cfg.basis = _lib.BASIS_KNOWN_VB


class HilbertGPBasis:
    """phi_sampling of Julia/examples/PG_OCP_generic_basis_functions_Altro.jl:52-99.

    n_phi_dims = [n_phi_u..., n_phi_x...] and L = [L_u..., L_x...] exactly as in the example.  The
    reference enumerates the index vectors with slot 1 fastest (:59-72) and then reverses them
    (:82), so augmented dimension k (z = [u; x]) uses slot n_z+1-k: its count is
    n_phi_dims[n_z+1-k] and its stride prod(n_phi_dims[1:n_z-k]) — the last state varies fastest.
    """

    def __init__(self, n_phi_dims, L, n_u=1):
        self.n_phi_dims = [int(v) for v in np.ravel(n_phi_dims)]
        self.L = [float(v) for v in np.ravel(L)]
        if len(self.n_phi_dims) != len(self.L):
            raise ValueError("n_phi_dims and L must have one entry per augmented dimension")
        self.n_z = len(self.L)
        self.n_u = int(n_u)
        self.n_x = self.n_z - self.n_u
        self.n_phi = int(np.prod(self.n_phi_dims))
        nz = self.n_z
        self.count = [self.n_phi_dims[nz - 1 - k] for k in range(nz)]
        self.stride = [int(np.prod(self.n_phi_dims[:nz - 1 - k])) for k in range(nz)]

    def fill(self, cfg):
        cfg.basis = _lib.BASIS_HILBERT_GP
        for k in range(self.n_z):
            cfg.hg_count[k], cfg.hg_stride[k], cfg.hg_L[k] = self.count[k], self.stride[k], self.L[k]


def hilbert_gp_prior_V(n_phi_dims, L, l, sf):
    """Diagonal of the prior V from the SE spectral density, generic example :59-80,:111-115.
    Replicates the reference: lambda is computed *before* the index reversal (:78 vs :82)."""
    n_phi_dims = [int(v) for v in np.ravel(n_phi_dims)]
    L = np.asarray(L, dtype=np.float64).ravel()
    l = np.asarray(l, dtype=np.float64).ravel()
    nz = len(L)
    n_phi = int(np.prod(n_phi_dims))
    variants = np.concatenate([[1], np.cumprod(n_phi_dims[:-1])]).astype(np.int64)
    V = np.empty(n_phi)
    for i in range(n_phi):
        rem = i
        sub = np.zeros(nz)
        for j in range(nz - 1, -1, -1):
            sub[j] = rem // variants[j] + 1
            rem = rem % variants[j]
        lam = (np.pi * sub / (2 * L)) ** 2
        V[i] = sf * ((2 * np.pi) ** (nz / 2)) * np.prod(l) * np.exp(-0.5 * np.sum((l ** 2) * lam))
    return V


class LinearMeasurement:
    """g(x, u) = C*x (both reference examples use [1 0]*x, known example :54)."""

    def __init__(self, C_):
        self.C = np.atleast_2d(np.asarray(C_, dtype=np.float64))


class MvNormal:
    """x_init_dist = MvNormal(mean, cov) (known example :48-50); cov scalar, vector (diag) or matrix."""

    def __init__(self, mean, cov):
        self.mean = np.asarray(mean, dtype=np.float64).ravel()
        n = self.mean.size
        cov = np.asarray(cov, dtype=np.float64)
        if cov.ndim == 0:
            cov = float(cov) * np.eye(n)
        elif cov.ndim == 1:
            cov = np.diag(cov)
        self.cov = cov
        self.chol = np.linalg.cholesky(cov)


def _basis(phi):
    if isinstance(phi, (KnownBasisVB, HilbertGPBasis)):
        return phi
    raise TypeError("phi must be a KnownBasisVB or HilbertGPBasis descriptor: arbitrary closures cannot cross the "
                    "C ABI and pgopt_b200 has no CPU fallback")


def _meas(g):
    if isinstance(g, LinearMeasurement):
        return g
    if isinstance(g, np.ndarray):
        return LinearMeasurement(g)
    raise TypeError("g must be a LinearMeasurement (g(x,u) = C*x) or the matrix C")


PRECISION = {"f64": 0, "f32": 1, None: 0}


def _make_cfg(phi, n_y, N, T, n_chains=1, device=0, keep_w_history=False, x_history="auto", precision="f64"):
    b = _basis(phi)
    cfg = PgbConfig()
    cfg.device, cfg.n_x, cfg.n_u, cfg.n_y, cfg.n_phi = device, b.n_x, b.n_u, n_y, b.n_phi
    b.fill(cfg)
    cfg.N, cfg.T, cfg.n_chains, cfg.keep_w_history = int(N), int(T), int(n_chains), int(keep_w_history)
    cfg.x_history = _lib.X_HISTORY[x_history]
    cfg.precision = PRECISION[precision]
    return cfg


# ----------------------------------------------------------------------------------------------
# engine (one handle = one device, one stream, n_chains batched chains)
# ----------------------------------------------------------------------------------------------
class ParticleGibbsEngine:
    def __init__(self, phi, n_y, N, T, n_chains=1, device=0, keep_w_history=False, x_history="auto", precision="f64"):
        """x_history: "store" keeps x_pf[n_x,N,T] on the device (get_history needs it), "replay" keeps two steps and
        recomputes the traced trajectory (bit-identical, N*T no longer bounded by memory), "auto": store up to 1 GiB.
        precision: "f64" (the reference's arithmetic) or "f32" — FP32 particle states / basis functions / noise /
        log-weights with bit-defined FP32 functions, FP64 weight normalisation and cumulative sums (include/pgopt_b200.h:
        PGB_PRECISION_F32); indices stay bit-exact against the oracle's F32 mode."""
        self.cfg = _make_cfg(phi, n_y, N, T, n_chains, device, keep_w_history, x_history, precision)
        self._phi = phi
        self.n_x, self.n_u, self.n_y, self.n_phi = self.cfg.n_x, self.cfg.n_u, n_y, self.cfg.n_phi
        self.N, self.T, self.n_chains = int(N), int(T), int(n_chains)
        self._h = C.c_void_p()
        check(lib().pgb_create(C.byref(self._h), C.byref(self.cfg)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().pgb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        check(rc, self._h)

    def set_mode(self, persistent=True):
        """True / -1 (library default): automatic — the warp-per-chain kernel for N <= 32 (the reference's examples use
        N = 30), the one-CTA-per-chain kernel for N <= 1024, otherwise the tile persistent kernel; 1: tile persistent
        kernel, one cooperative launch per sweep; 2: warp-tile persistent kernel (alternative formulation, slower);
        0 / False: multi-kernel path (13 launches per step; per-kernel profiling and cross-check); 3: warp / one-CTA
        kernels (fall back to 1 when N > 1024).  All of them are bit-identical."""
        mode = -1 if persistent is True else int(persistent)
        self._ck(lib().pgb_set_mode(self._h, C.c_int32(mode)))

    def set_data(self, u, y):
        u, y = _f(np.atleast_2d(u)), _f(np.atleast_2d(y))
        assert u.size == self.n_u * self.T and y.size == self.n_y * self.T
        self._ck(lib().pgb_set_data(self._h, _p(u), _p(y)))

    def set_measurement(self, g, R):
        g = _meas(g)
        Cm, Rm = _f(g.C), _f(np.atleast_2d(np.asarray(R, dtype=np.float64)))
        assert Cm.size == self.n_y * self.n_x and Rm.size == self.n_y * self.n_y
        self._ck(lib().pgb_set_measurement(self._h, _p(Cm), _p(Rm)))

    def set_init_dist(self, dist):
        self._ck(lib().pgb_set_init_dist(self._h, _p(_f(dist.mean)), _p(_f(dist.chol))))

    def set_prior(self, V, Lambda_Q, ell_Q):
        """V: the diagonal (n_phi,) or the matrix (n_phi, n_phi) of the MN prior's left covariance (particle_Gibbs.jl:56,
        :64).  A diagonal matrix takes the diagonal path (1 / V[i], as both reference examples), any other symmetric
        positive definite V the dense one (I / V by Cholesky, once)."""
        V, dense = _prior_V(V)
        LQ = np.asarray(Lambda_Q, dtype=np.float64)
        fn = lib().pgb_set_prior_dense if dense else lib().pgb_set_prior
        self._ck(fn(self._h, _p(_f(V)), _p(_f(LQ)), C.c_double(float(ell_Q))))

    def set_state(self, A, Q, x_prim=None, sweep_index=1, chain=0):
        xp = None if x_prim is None else _f(x_prim)
        self._ck(lib().pgb_set_state(self._h, C.c_int32(chain), _p(_f(A)), _p(_f(Q)), _p(xp), C.c_int64(sweep_index)))

    def get_state(self, chain=0):
        A, Q, xp = np.empty(self.n_x * self.n_phi), np.empty(self.n_x * self.n_x), np.empty(self.n_x * self.T)
        k = C.c_int64(0)
        self._ck(lib().pgb_get_state(self._h, C.c_int32(chain), _p(A), _p(Q), _p(xp), C.byref(k)))
        return (A.reshape((self.n_x, self.n_phi), order="F"), Q.reshape((self.n_x, self.n_x), order="F"),
                xp.reshape((self.n_x, self.T), order="F"), k.value)

    def set_rng_philox(self, seed, chain_base=0):
        self._ck(lib().pgb_set_rng_philox(self._h, C.c_uint64(seed), C.c_uint32(chain_base)))

    def set_rng_injected(self, Z0, Ures, Z, Uanc, Ustar, bartlett=None, Xmn=None, chain=0):
        """Z0 (N-1, n_x), Ures (T,), Z (T, N, n_x), Uanc (T,), Ustar, bartlett (n_x,n_x) lower, Xmn (n_x,n_phi)."""
        Z0c = np.ascontiguousarray(Z0, dtype=np.float64)
        Zc = np.ascontiguousarray(Z, dtype=np.float64)
        assert Z0c.size == (self.N - 1) * self.n_x and Zc.size == self.T * self.N * self.n_x
        b = None if bartlett is None else _f(bartlett)
        X = None if Xmn is None else _f(Xmn)
        self._ck(lib().pgb_set_rng_injected(self._h, C.c_int32(chain), _p(Z0c), _p(_f(Ures)), _p(Zc), _p(_f(Uanc)),
                                            C.c_double(float(Ustar)), _p(b), _p(X)))

    def sweep(self, n_sweeps=1):
        self._ck(lib().pgb_sweep(self._h, C.c_int64(n_sweeps)))

    def sweep_filter(self):
        self._ck(lib().pgb_sweep_filter(self._h))

    def sweep_mniw(self):
        self._ck(lib().pgb_sweep_mniw(self._h))

    def synchronize(self):
        self._ck(lib().pgb_synchronize(self._h))

    def get_sample(self, chain=0):
        A, Q = np.empty(self.n_x * self.n_phi), np.empty(self.n_x * self.n_x)
        w, x, u = np.empty(self.N), np.empty(self.n_x * self.N), np.empty(self.n_u)
        self._ck(lib().pgb_get_sample(self._h, C.c_int32(chain), _p(A), _p(Q), _p(w), _p(x), _p(u)))
        return PG_sample(A.reshape((self.n_x, self.n_phi), order="F"), Q.reshape((self.n_x, self.n_x), order="F"), w,
                         x.reshape((self.n_x, self.N), order="F"), u)

    def get_history(self, chain=0, want_w=True, want_x=True):
        a = np.zeros((self.T, self.N), dtype=np.int64)
        x = np.zeros((self.T, self.N, self.n_x)) if want_x else None
        w = np.zeros((self.T, self.N)) if want_w else None
        self._ck(lib().pgb_get_history(self._h, C.c_int32(chain), _p(a), _p(x), _p(w)))
        return a, x, w

    def get_stats(self, chain=0):
        Phi, Psi, Sig = np.empty(self.n_x ** 2), np.empty(self.n_x * self.n_phi), np.empty(self.n_phi ** 2)
        star = C.c_int64(0)
        self._ck(lib().pgb_get_stats(self._h, C.c_int32(chain), _p(Phi), _p(Psi), _p(Sig), C.byref(star)))
        return (Phi.reshape((self.n_x, self.n_x), order="F"), Psi.reshape((self.n_x, self.n_phi), order="F"),
                Sig.reshape((self.n_phi, self.n_phi), order="F"), star.value)

    def status(self, chain=0, clear=False):
        st, nf, ns = C.c_int32(0), C.c_int64(0), C.c_int64(0)
        self._ck(lib().pgb_get_status(self._h, C.c_int32(chain), C.byref(st), C.byref(nf), C.byref(ns), C.c_int32(int(clear))))
        return st.value, nf.value, ns.value

    def particle_gibbs(self, K, K_b, k_d, keep_particles=True):
        nc, nx, np_, N = self.n_chains, self.n_x, self.n_phi, self.N
        A_s, Q_s = np.empty((nc, K, nx * np_)), np.empty((nc, K, nx * nx))
        w_s = np.empty((nc, K, N)) if keep_particles else None
        x_s = np.empty((nc, K, nx * N)) if keep_particles else None
        u_m1 = np.empty(self.n_u)
        self._ck(lib().pgb_particle_gibbs(self._h, C.c_int64(K), C.c_int64(K_b), C.c_int64(k_d), _p(A_s), _p(Q_s),
                                          _p(w_s), _p(x_s), _p(u_m1)))
        return A_s, Q_s, w_s, x_s, u_m1

    def particle_gibbs_packed(self, K, K_b, k_d, roll_seed, k_offset):
        """the sampler, then x_T = x_m1[:, star] drawn on the device (scenario stream ids k_offset + chain*K + k):
        (n_chains, K, R) packed records [A | Q | x_T | u_m1]; the particle arrays never reach the host"""
        ds = self.particle_gibbs_dev(K, K_b, k_d)
        try:
            ds.draw_xT(roll_seed, k_offset)
            return ds.records().reshape(self.n_chains, K, -1)
        finally:
            ds.close()

    def checkpoint(self):
        """everything needed to continue the chains bit-identically on a handle of the same configuration
        (per chain A, Q, x_prim; sweep index; Philox seed / chain_base) as a uint8 array"""
        n = lib().pgb_checkpoint_bytes(self._h)
        buf = np.empty(n, dtype=np.uint8)
        self._ck(lib().pgb_checkpoint_save(self._h, _p(buf), C.c_int64(n)))
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8))
        self._ck(lib().pgb_checkpoint_load(self._h, _p(buf), C.c_int64(buf.size)))

    def particle_gibbs_dev(self, K, K_b, k_d):
        """the sampler with the kept samples left on the device: DeviceSamples with slot chain*K + k"""
        out = DeviceSamples(self._phi, self.n_y, self.N, self.n_chains * K, self.cfg.device)
        try:
            self._ck(lib().pgb_particle_gibbs_dev(self._h, C.c_int64(K), C.c_int64(K_b), C.c_int64(k_d), out._s))
        except Exception:
            out.close()
            raise
        return out


class MultiGPUEngine:
    """pgb_multi_*: independent chains / independent scenarios over several GPUs of one node driven from ONE process
    (what the Julia shim uses: `device = [0, 1, ...]`).  n_chains is the TOTAL over all devices; results are independent
    of the device count (global Philox stream ids)."""

    def __init__(self, phi, n_y, N, T, n_chains, devices, x_history="auto"):
        self.cfg = _make_cfg(phi, n_y, N, T, n_chains, 0, False, x_history)
        self.basis = _basis(phi)
        self.n_x, self.n_u, self.n_y, self.n_phi = self.cfg.n_x, self.cfg.n_u, int(n_y), self.cfg.n_phi
        self.N, self.T, self.n_chains = int(N), int(T), int(n_chains)
        self.devices = [int(d) for d in devices]
        self.K = None
        self._n_scen = None
        self._m = C.c_void_p()
        dv = (C.c_int32 * len(self.devices))(*self.devices)
        check(lib().pgb_multi_create(C.byref(self._m), C.byref(self.cfg), dv, C.c_int32(len(self.devices))))

    def _ck(self, rc):
        if rc != 0:
            msg = lib().pgb_multi_last_error(self._m)
            raise PgoptError(rc, msg.decode() if msg else "unknown")

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            lib().pgb_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def record_doubles(self):
        return int(lib().pgb_packed_record_doubles(C.c_int32(self.n_x), C.c_int32(self.n_phi), C.c_int32(self.n_u)))

    def set_data(self, u, y):
        self._ck(lib().pgb_multi_set_data(self._m, _p(_f(np.atleast_2d(u))), _p(_f(np.atleast_2d(y)))))

    def set_measurement(self, g, R):
        g = _meas(g)
        self._ck(lib().pgb_multi_set_measurement(self._m, _p(_f(g.C)), _p(_f(np.atleast_2d(np.asarray(R, dtype=np.float64))))))

    def set_init_dist(self, dist):
        self._ck(lib().pgb_multi_set_init_dist(self._m, _p(_f(dist.mean)), _p(_f(dist.chol))))

    def set_prior(self, V, Lambda_Q, ell_Q):
        V, dense = _prior_V(V)
        fn = lib().pgb_multi_set_prior_dense if dense else lib().pgb_multi_set_prior
        self._ck(fn(self._m, _p(_f(V)), _p(_f(np.asarray(Lambda_Q, dtype=np.float64))), C.c_double(float(ell_Q))))

    def set_state(self, A, Q, x_prim=None, sweep_index=1, chain=-1):
        xp = None if x_prim is None else _f(x_prim)
        self._ck(lib().pgb_multi_set_state(self._m, C.c_int32(chain), _p(_f(A)), _p(_f(Q)), _p(xp), C.c_int64(sweep_index)))

    def set_rng_philox(self, seed, chain_base=0):
        self._ck(lib().pgb_multi_set_rng_philox(self._m, C.c_uint64(seed), C.c_uint32(chain_base)))

    def particle_gibbs(self, K, K_b, k_d, roll_seed=0):
        """(n_chains, K, R) packed records [A | Q | x_T | u_m1], gathered over NCCL to every device and read from the first"""
        out = np.empty((self.n_chains, K, self.record_doubles))
        self._ck(lib().pgb_multi_particle_gibbs(self._m, C.c_int64(K), C.c_int64(K_b), C.c_int64(k_d), C.c_uint64(roll_seed), _p(out)))
        self.K, self._n_scen = int(K), self.n_chains * int(K)
        return out

    def packed_on(self, which_device):
        out = np.empty((self._n_scen, self.record_doubles))
        self._ck(lib().pgb_multi_get_packed(self._m, C.c_int32(which_device), _p(out)))
        return out

    def get_sample(self, chain, k):
        A, Q = np.empty(self.n_x * self.n_phi), np.empty(self.n_x * self.n_x)
        w, x, u = np.empty(self.N), np.empty(self.n_x * self.N), np.empty(self.n_u)
        self._ck(lib().pgb_multi_get_sample(self._m, C.c_int32(chain), C.c_int64(k), _p(A), _p(Q), _p(w), _p(x), _p(u)))
        return PG_sample(A.reshape((self.n_x, self.n_phi), order="F"), Q.reshape((self.n_x, self.n_x), order="F"), w,
                         x.reshape((self.n_x, self.N), order="F"), u)

    def set_samples(self, PG_samples, roll_seed=0):
        A, Q, w, x, um, N = _pack_samples(PG_samples)
        assert N == self.N
        self._ck(lib().pgb_multi_set_samples(self._m, C.c_int64(len(PG_samples)), _p(A), _p(Q), _p(w), _p(x), _p(um), C.c_uint64(roll_seed)))
        self._n_scen = len(PG_samples)

    def rollout(self, g, R, H, u_init=None, seed=0, want=("x0", "X", "Y"), out=None):
        gm = _meas(g)
        Le = _noise_factor(R)
        if u_init is None:
            u_init = np.zeros((self.n_u, H))
        Kt, nx, ny = self._n_scen, self.n_x, self.n_y
        shapes = dict(x0=(Kt, nx), X=(Kt, H + 1, nx), Y=(Kt, H, ny))
        bufs = {k: (out[k] if out and k in out else np.empty(shapes[k])) for k in want}
        self._ck(lib().pgb_multi_rollout(self._m, C.c_int64(H), _p(_f(gm.C)), _p(_f(Le)), _p(_f(np.atleast_2d(u_init))), C.c_uint64(seed),
                                         _p(bufs.get("x0")), _p(bufs.get("X")), _p(bufs.get("Y"))))
        return bufs

    def dynamics(self, x_vec, u, t=None, phi_variant=0, jacobians=True):
        """discrete_dynamics + Jacobians of the stacked scenario model (optimal_control_Altro.jl:79-107, :51) on the sharded
        resident samples: every device evaluates its own scenarios (see DeviceSamples.dynamics for the layouts).
        Returns (x_next (K*n_x,), dfdx (K, n_x, n_x), dfdu (K, n_x, n_u), slowest device's kernel ms)."""
        Kt, nx, nu = self._n_scen, self.n_x, self.n_u
        xv = np.ascontiguousarray(np.asarray(x_vec, dtype=np.float64).reshape(-1))
        assert xv.size == Kt * nx
        xn = np.empty(Kt * nx)
        jx = np.empty((Kt, nx, nx)) if jacobians else None
        ju = np.empty((Kt, nu, nx)) if jacobians else None
        ms = C.c_double(0)
        self._ck(lib().pgb_multi_dynamics(self._m, _p(xv), _p(_f(np.atleast_1d(u))), C.c_int64(-1 if t is None else int(t)),
                                          C.c_int32(int(phi_variant)), _p(xn), _p(jx), _p(ju), C.byref(ms)))
        if not jacobians:
            return xn, None, None, ms.value
        return xn, np.transpose(jx, (0, 2, 1)), np.transpose(ju, (0, 2, 1)), ms.value

    def scenarios_on(self, which_device, H):
        """the all-gathered x_vec_0 / X / Y as device `which_device` holds them after rollout()"""
        Kt, nx, ny = self._n_scen, self.n_x, self.n_y
        bufs = dict(x0=np.empty((Kt, nx)), X=np.empty((Kt, H + 1, nx)), Y=np.empty((Kt, H, ny)))
        self._ck(lib().pgb_multi_get_scenarios(self._m, C.c_int32(which_device), _p(bufs["x0"]), _p(bufs["X"]), _p(bufs["Y"])))
        return bufs

    def last_ms(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self._ck(lib().pgb_multi_last_ms(self._m, C.byref(a), C.byref(b), C.byref(c)))
        return dict(sample_gather_ms=a.value, rollout_kernel_ms_max=b.value, scenario_gather_ms=c.value)


# ----------------------------------------------------------------------------------------------
# the reference's exported functions
# ----------------------------------------------------------------------------------------------
def systematic_resampling(W, N, rnd=None, rng=None, device=0, return_status=False):
    """systematic_resampling(W, N) — Julia/src/particle_Gibbs.jl:17-34.  `rnd` is the rand() of :21
    (drawn from `rng` / NumPy when omitted).  Returns 1-based Int64 indices."""
    W = np.ascontiguousarray(W, dtype=np.float64).ravel()
    if rnd is None:
        rnd = (rng or np.random.default_rng()).random()
    idx = np.empty(int(N), dtype=np.int64)
    st = C.c_int32(0)
    check(lib().pgb_systematic_resampling(C.c_int32(device), _p(W), C.c_int64(W.size), C.c_int64(int(N)),
                                          C.c_double(float(rnd)), _p(idx), C.byref(st)))
    if st.value & _lib.STATUS_RESAMPLE_OVERRUN and not return_status:
        raise IndexError("BoundsError: attempt to access %d-element Vector at index [%d] (weights sum below the "
                         "sampling grid, Julia/src/particle_Gibbs.jl:28)" % (W.size, W.size + 1))
    return (idx, st.value) if return_status else idx


def MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T, bartlett=None, Xmn=None, seed=0, chain=0, sweep_index=1,
                device=0):
    """MNIW_sample(Phi, Psi, Sigma, V, Lambda_Q, ell_Q, T) — Julia/src/particle_Gibbs.jl:56-81.
    The Bartlett factor / matrix-normal normals may be injected (parity) or come from Philox."""
    Phi, Psi, Sigma = np.atleast_2d(Phi), np.atleast_2d(Psi), np.atleast_2d(Sigma)
    n_x, n_phi = Psi.shape
    V, dense = _prior_V(V)
    A, Q = np.empty(n_x * n_phi), np.empty(n_x * n_x)
    b = None if bartlett is None else _f(bartlett)
    X = None if Xmn is None else _f(Xmn)
    fn = lib().pgb_mniw_sample_dense if dense else lib().pgb_mniw_sample
    check(fn(C.c_int32(device), C.c_int32(n_x), C.c_int32(n_phi), _p(_f(Phi)), _p(_f(Psi)),
             _p(_f(Sigma)), _p(_f(V)), _p(_f(np.asarray(Lambda_Q, dtype=np.float64))),
             C.c_double(float(ell_Q)), C.c_int64(int(T)), _p(b), _p(X), C.c_uint64(seed),
             C.c_uint32(chain), C.c_uint32(sweep_index), _p(A), _p(Q)))
    return A.reshape((n_x, n_phi), order="F"), Q.reshape((n_x, n_x), order="F")


def particle_Gibbs(u_training, y_training, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist, g, R,
                   x_prim=None, seed=0, chain=0, device=0, verbose=False, x_history="auto", precision="f64"):
    """particle_Gibbs(u, y, K, K_b, k_d, N, phi, Lambda_Q, ell_Q, Q_init, V, A_init, x_init_dist, g, R; x_prim)
    — Julia/src/particle_Gibbs.jl:114-237.  Returns a list of K PG_sample.  `x_prim`, when given, is
    updated in place like the reference does (:214-217)."""
    u = np.atleast_2d(np.asarray(u_training, dtype=np.float64))
    y = np.atleast_2d(np.asarray(y_training, dtype=np.float64))
    A_init = np.atleast_2d(np.asarray(A_init, dtype=np.float64))
    n_y, T = y.shape
    eng = ParticleGibbsEngine(phi, n_y, N, T, 1, device, x_history=x_history, precision=precision)
    try:
        eng.set_data(u, y)
        eng.set_measurement(g, R)
        eng.set_init_dist(x_init_dist)
        eng.set_prior(V, Lambda_Q, ell_Q)
        eng.set_state(A_init, np.asarray(Q_init, dtype=np.float64) * np.eye(A_init.shape[0])
                      if np.ndim(Q_init) == 0 else Q_init, x_prim, 1)
        eng.set_rng_philox(seed, chain)
        if verbose:
            print("### Started learning algorithm")
        A_s, Q_s, w_s, x_s, u_m1 = eng.particle_gibbs(K, K_b, k_d)
        if x_prim is not None:
            x_prim[...] = eng.get_state()[2]
        n_x, n_phi = eng.n_x, eng.n_phi
        out = []
        for k in range(K):
            out.append(PG_sample(A_s[0, k].reshape((n_x, n_phi), order="F"), Q_s[0, k].reshape((n_x, n_x), order="F"),
                                 w_s[0, k].copy(), x_s[0, k].reshape((n_x, N), order="F"), u_m1.copy()))
        if verbose:
            print("### Learning complete")
        return out
    finally:
        eng.close()


def _noise_factor(R):
    if np.ndim(R) == 0:
        return np.array([[float(R)]])  # MvNormal(zeros(n_y), R::Real) reads R as a std (reference quirk, :79)
    return np.linalg.cholesky(np.atleast_2d(np.asarray(R, dtype=np.float64)))


def _supplied(x_vec_0, v_vec, e_vec, K, H, n_x, n_y):
    """the reference's keyword arrays (x_vec_0 (n_x*K,), v_vec (n_x,H,K), e_vec (n_y,H,K)) in the C ABI's
    scenario-major layouts, or None"""
    x0 = None if x_vec_0 is None else np.ascontiguousarray(np.asarray(x_vec_0, dtype=np.float64).reshape(K, n_x))
    v = None if v_vec is None else np.ascontiguousarray(np.transpose(np.asarray(v_vec, dtype=np.float64).reshape(n_x, H, K), (2, 1, 0)))
    e = None if e_vec is None else np.ascontiguousarray(np.transpose(np.asarray(e_vec, dtype=np.float64).reshape(n_y, H, K), (2, 1, 0)))
    return x0, v, e


def _scenario_dict(x0, v, e, X, Y, K, H, n_x, n_y):
    return dict(x_vec_0=x0.reshape(-1), v_vec=np.transpose(v, (2, 1, 0)), e_vec=np.transpose(e, (2, 1, 0)),
                X_init=np.transpose(X, (0, 2, 1)).reshape(K * n_x, H + 1),
                Y_init=np.transpose(Y, (0, 2, 1)).reshape(K * n_y, H), raw=dict(x0=x0, v=v, e=e, X=X, Y=Y))


def scenario_rollout(PG_samples, phi, g, R, H, u_init=None, seed=0, k_offset=0, device=0, x_vec_0=None, v_vec=None,
                     e_vec=None):
    """Scenario sampling and forward rollout that solve_PG_OCP_Ipopt performs before building the NLP
    (Julia/src/optimal_control_Ipopt.jl:47-83,114-125).  Returns the arrays in the reference's layouts so
    they can be passed back through its `x_vec_0`, `v_vec`, `e_vec`, `u_init` keyword arguments:
    x_vec_0 (n_x*K,), v_vec (n_x,H,K), e_vec (n_y,H,K), X_init (n_x*K,H+1), Y_init (n_y*K,H).
    `x_vec_0` / `v_vec` / `e_vec`, when given, are used instead of the draws exactly as the reference's keyword
    arguments are (:48, :67, :77) — the second entry of the pre-solve step (:88-112)."""
    b = _basis(phi)
    gm = _meas(g)
    K = len(PG_samples)
    n_x, n_u, n_phi = b.n_x, b.n_u, b.n_phi
    n_y = gm.C.shape[0]
    A, Q, w, x, um, N = _pack_samples(PG_samples)
    Le = _noise_factor(R)
    if u_init is None:
        u_init = np.zeros((n_u, H))  # :110
    cfg = _make_cfg(b, n_y, N, 2, 1, device)
    x0i, vi, ei = _supplied(x_vec_0, v_vec, e_vec, K, H, n_x, n_y)
    x0, v, e = np.empty((K, n_x)), np.empty((K, H, n_x)), np.empty((K, H, n_y))
    X, Y = np.empty((K, H + 1, n_x)), np.empty((K, H, n_y))
    check(lib().pgb_rollout_supplied(C.byref(cfg), C.c_int64(K), C.c_int64(H), _p(A), _p(Q), _p(w), _p(x), _p(um),
                                     _p(_f(gm.C)), _p(_f(Le)), _p(_f(np.atleast_2d(u_init))), C.c_uint64(seed),
                                     C.c_uint32(k_offset), _p(x0i), _p(vi), _p(ei), _p(x0), _p(v), _p(e), _p(X), _p(Y)))
    return _scenario_dict(x0, v, e, X, Y, K, H, n_x, n_y)


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib().pgb_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """NumPy array over page-locked host memory (pgb_host_alloc): D2H / H2D of the scenario arrays run at PCIe speed
    instead of the pageable-copy rate (a pageable cudaMemcpy of the configs[4] inputs was 261 ms of a 0.74 s call)"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib().pgb_host_alloc(C.c_size_t(max(n, 16)))
    if not ptr:
        raise MemoryError("pgb_host_alloc(%d) failed" % n)
    owner = _PinnedOwner(ptr)
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    arr = arr.view(_PinnedArray)
    arr._owner = owner
    return arr


class _PinnedArray(np.ndarray):
    """ndarray that keeps its page-locked allocation alive"""
    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)



class DeviceSamples:
    """K posterior samples resident on one GPU (pgb_samples): what `Vector{PG_sample}` is to the reference
    (PGopt.jl:23-29), except that sampler -> rollout -> screening read it in place on the device, as
    optimal_control_Ipopt.jl:47-64 reads PG_samples[k].A in place on the host."""

    def __init__(self, phi, n_y, N, K, device=0):
        self.basis = _basis(phi)
        self.cfg = _make_cfg(self.basis, n_y, N, 2, 1, device)
        self.K, self.N, self.n_y = int(K), int(N), int(n_y)
        self.n_x, self.n_u, self.n_phi = self.basis.n_x, self.basis.n_u, self.basis.n_phi
        self.H = None
        self._s = C.c_void_p()
        check(lib().pgb_samples_create(C.byref(self._s), C.byref(self.cfg), C.c_int64(self.K)))

    def _ck(self, rc):
        if rc != 0:
            msg = lib().pgb_samples_last_error(self._s)
            raise PgoptError(rc, msg.decode() if msg else "unknown")

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            lib().pgb_samples_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self):
        return self.K

    @classmethod
    def from_samples(cls, PG_samples, phi, n_y=1, device=0):
        A, Q, w, x, um, N = _pack_samples(PG_samples)
        s = cls(phi, n_y, N, len(PG_samples), device)
        s._ck(lib().pgb_samples_set(s._s, C.c_int64(0), C.c_int64(s.K), _p(A), _p(Q), _p(w), _p(x), _p(um)))
        return s

    def to_batch(self):
        K, nx, npb, N, nu = self.K, self.n_x, self.n_phi, self.N, self.n_u
        A, Q = np.empty((K, nx * npb)), np.empty((K, nx * nx))
        w, x, um = np.empty((K, N)), np.empty((K, nx * N)), np.empty((K, nu))
        self._ck(lib().pgb_samples_get(self._s, C.c_int64(0), C.c_int64(K), _p(A), _p(Q), _p(w), _p(x), _p(um)))
        return PGSampleBatch(np.transpose(A.reshape(K, npb, nx), (0, 2, 1)), np.transpose(Q.reshape(K, nx, nx), (0, 2, 1)),
                             w, np.transpose(x.reshape(K, N, nx), (0, 2, 1)), um)

    def store(self, k, engine, chain=0):
        """slot k <- the current PG_sample of `chain` of `engine`, device to device (particle_Gibbs.jl:204-208)"""
        self._ck(lib().pgb_samples_store(self._s, C.c_int64(k), engine._h, C.c_int32(chain)))

    def draw_xT(self, seed=0, k_offset=0):
        """x_m1[:, star_k] with star_k ~ systematic_resampling(w_m1, 1) (optimal_control_Ipopt.jl:58-59), drawn on the
        device that holds the sample; (K, n_x)"""
        xT = np.empty((self.K, self.n_x))
        self._ck(lib().pgb_samples_draw_xT(self._s, C.c_uint64(seed), C.c_uint32(k_offset), _p(xT)))
        return xT

    def rollout(self, g, R, H, u_init=None, seed=0, k_offset=0, x_vec_0=None, v_vec=None, e_vec=None, reuse=()):
        """scenario_rollout on the resident samples; results stay on the device (fetch with scenarios()).  `reuse`: any
        of "x_vec_0", "v_vec", "e_vec" — keep that array of the previous rollout of this set instead of drawing (the
        pre-solve re-entry with a new u_init, optimal_control_Ipopt.jl:88-125)."""
        gm = _meas(g)
        Le = _noise_factor(R)
        if u_init is None:
            u_init = np.zeros((self.n_u, H))
        x0i, vi, ei = _supplied(x_vec_0, v_vec, e_vec, self.K, H, self.n_x, self.n_y)
        bits = sum(b for name, b in (("x_vec_0", 1), ("v_vec", 2), ("e_vec", 4)) if name in reuse)
        self._ck(lib().pgb_rollout_dev(self._s, C.c_int64(H), _p(_f(gm.C)), _p(_f(Le)), _p(_f(np.atleast_2d(u_init))),
                                       C.c_uint64(seed), C.c_uint32(k_offset), C.c_int32(bits), _p(x0i), _p(vi), _p(ei)))
        self.H = int(H)
        return self

    def scenarios(self, want=("x0", "v", "e", "X", "Y"), out=None):
        """the reference-layout dict of scenario_rollout for the last rollout (D2H of the requested arrays only)"""
        K, H, nx, ny = self.K, self.H, self.n_x, self.n_y
        shapes = dict(x0=(K, nx), v=(K, H, nx), e=(K, H, ny), X=(K, H + 1, nx), Y=(K, H, ny))
        bufs = {k: (out[k] if out and k in out else np.empty(shapes[k])) for k in want}
        self._ck(lib().pgb_scenarios_get(self._s, *[_p(bufs.get(k)) for k in ("x0", "v", "e", "X", "Y")]))
        return bufs

    def constraint_max(self, y_min, y_max):
        """scenario_constraint_max on the Y the last rollout left on the device"""
        y_min = np.atleast_2d(np.asarray(y_min, dtype=np.float64))
        y_max = np.atleast_2d(np.asarray(y_max, dtype=np.float64))
        if y_min.shape != (self.n_y, self.H) or y_max.shape != (self.n_y, self.H):
            raise ValueError("y_min and y_max must both be (n_y, H)")
        h, order = np.empty(self.K), np.empty(self.K, dtype=np.int64)
        self._ck(lib().pgb_scenario_constraint_max_dev(self._s, _p(_f(y_min)), _p(_f(y_max)), _p(h), _p(order)))
        return h, order

    def pack(self, with_particles=True):
        """the versioned packed layout (include/pgopt_b200.h: header + [A | Q | x_T | u_m1] records [+ particle blocks])
        as a bytes-like uint8 array — the on-disk / on-wire form of a Vector{PG_sample}"""
        n = lib().pgb_samples_pack_bytes(self._s, C.c_int32(int(with_particles)))
        buf = np.empty(n, dtype=np.uint8)
        self._ck(lib().pgb_samples_pack(self._s, C.c_int32(int(with_particles)), _p(buf), C.c_int64(n)))
        return buf

    def unpack(self, buf):
        buf = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8))
        self._ck(lib().pgb_samples_unpack(self._s, _p(buf), C.c_int64(buf.size)))
        return self

    def records(self):
        """(K, R) array of packed records [A | Q | x_T | u_m1] (x_T zeros until draw_xT)"""
        buf = self.pack(with_particles=False)
        R = int(lib().pgb_packed_record_doubles(C.c_int32(self.n_x), C.c_int32(self.n_phi), C.c_int32(self.n_u)))
        return buf[64:].view(np.float64).reshape(-1, R).copy()

    def dynamics(self, x_vec, u, t=None, phi_variant=0, jacobians=True):
        """discrete_dynamics of the stacked scenario model (Julia/src/optimal_control_Altro.jl:79-107) and its Jacobians
        (RobotDynamics.@autodiff, :51): x_vec (K*n_x,) stacked as in the reference, u (n_u,), t = 0-based time index whose
        process noise v_vec[:, t+1, k] (from the last rollout of this set) is added, None = no noise.  Returns
        (x_vec_next (K*n_x,), dfdx (K, n_x, n_x), dfdu (K, n_x, n_u)); the full Jacobian w.r.t. x_vec is block diagonal."""
        K, nx, nu = self.K, self.n_x, self.n_u
        xv = np.ascontiguousarray(np.asarray(x_vec, dtype=np.float64).reshape(-1))
        assert xv.size == K * nx
        xn = np.empty(K * nx)
        jx = np.empty((K, nx, nx)) if jacobians else None
        ju = np.empty((K, nu, nx)) if jacobians else None
        self._ck(lib().pgb_dynamics_dev(self._s, _p(xv), _p(_f(np.atleast_1d(u))), C.c_int64(-1 if t is None else int(t)),
                                        C.c_int32(int(phi_variant)), _p(xn), _p(jx), _p(ju)))
        if not jacobians:
            return xn, None, None
        return xn, np.transpose(jx, (0, 2, 1)), np.transpose(ju, (0, 2, 1))

    def last_ms(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self._ck(lib().pgb_samples_last_ms(self._s, C.byref(a), C.byref(b), C.byref(c)))
        return dict(rollout_kernel_ms=a.value, screening_kernel_ms=b.value, upload_ms=c.value)


def test_prediction(PG_samples, phi, g, R, k_n, u_test, y_test, seed=0, device=0):
    """Simulate the PG samples forward over the test input and compare with the test output
    (Julia/src/particle_Gibbs.jl:253-314): every model k_n times, on the scenario-rollout kernel.
    Returns dict(x_test_sim (n_x, T_test+1, K*k_n), y_test_sim (n_y, T_test, K*k_n), mean_rmse) with the
    reference's layouts after its reshape (:301-302); it does not plot (plot_predictions is out of scope).
    The reference leaves x_test_sim[:, T_test+1, :] undefined (:278); here it holds one more transition."""
    b = _basis(phi)
    gm = _meas(g)
    K = len(PG_samples)
    n_x, n_u = b.n_x, b.n_u
    n_y = gm.C.shape[0]
    A, Q, w, x, um, N = _pack_samples(PG_samples)
    u_test = np.atleast_2d(np.asarray(u_test, dtype=np.float64))
    y_test = np.atleast_2d(np.asarray(y_test, dtype=np.float64))
    T_test = y_test.shape[1]
    Le = _noise_factor(R)  # MvNormal(zeros(n_y), R::Real) reads R as a std (reference quirk, :261)
    cfg = _make_cfg(b, n_y, N, 2, 1, device)
    S = K * int(k_n)
    xs, ys = np.empty((S, T_test + 1, n_x)), np.empty((S, T_test, n_y))
    rmse = C.c_double(0.0)
    check(lib().pgb_test_prediction(C.byref(cfg), C.c_int64(K), C.c_int64(int(k_n)), C.c_int64(T_test), _p(A), _p(Q),
                                    _p(w), _p(x), _p(um), _p(_f(gm.C)), _p(_f(Le)), _p(_f(u_test[:, :T_test])),
                                    _p(_f(y_test)), C.c_uint64(seed), _p(xs), _p(ys), C.byref(rmse)))
    return dict(x_test_sim=np.transpose(xs, (2, 1, 0)), y_test_sim=np.transpose(ys, (2, 1, 0)),
                mean_rmse=rmse.value, raw=dict(x_sim=xs, y_sim=ys))


def scenario_constraint_max(y, y_min, y_max, device=0):
    """The scenario ordering of the greedy support-sub-sample search (Julia/src/optimal_control_Ipopt.jl:334-338,
    optimal_control_Altro.jl:496-500) for the examples' box output constraints
    (examples/PG_OCP_known_basis_functions_Ipopt.jl:198-214): `y` is y_opt (n_y, H, K) — or the scenario-major
    (K, H, n_y) array `scenario_rollout(...)["raw"]["Y"]` when `y_min` has shape (n_y, H) and y.shape[0] != n_y.
    Returns (h_scenario_max (K,), scenarios_sorted (K,) 1-based = sortperm(h_scenario_max; rev=true))."""
    y_min = np.atleast_2d(np.asarray(y_min, dtype=np.float64))
    y_max = np.atleast_2d(np.asarray(y_max, dtype=np.float64))
    n_y, H = y_min.shape
    if y_max.shape != (n_y, H):
        raise ValueError("y_min and y_max must both be (n_y, H)")
    y = np.asarray(y, dtype=np.float64)
    if y.ndim != 3:
        raise ValueError("y must be (n_y, H, K) or (K, H, n_y)")
    if y.shape[:2] == (n_y, H):
        Y = np.ascontiguousarray(np.transpose(y, (2, 1, 0)))
    elif y.shape[1:] == (H, n_y):
        Y = np.ascontiguousarray(y)
    else:
        raise ValueError("y %s does not match bounds (n_y=%d, H=%d)" % (y.shape, n_y, H))
    K = Y.shape[0]
    h = np.empty(K)
    order = np.empty(K, dtype=np.int64)
    check(lib().pgb_scenario_constraint_max(C.c_int32(device), C.c_int64(K), C.c_int64(H), C.c_int32(n_y), _p(Y),
                                            _p(_f(y_min)), _p(_f(y_max)), _p(h), _p(order)))
    return h, order
