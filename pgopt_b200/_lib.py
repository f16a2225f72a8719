"""ctypes binding of libpgopt_b200.so (include/pgopt_b200.h).  Fails loudly when the CUDA library is
missing: there is no CPU path in this package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGOPT_B200_LIB", os.path.join(HERE, "lib", "libpgopt_b200.so"))
PGB_MAX_Z = 8

PGB_OK, PGB_ERR_INVALID, PGB_ERR_CUDA, PGB_ERR_NOT_POSDEF, PGB_ERR_STATE = 0, 1, 2, 3, 4
BASIS_KNOWN_VB, BASIS_HILBERT_GP = 0, 1
STATUS_RESAMPLE_OVERRUN, STATUS_INDEX_ZERO, STATUS_NONFINITE, STATUS_NOT_POSDEF = 1, 2, 4, 16


class PgbConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_x", C.c_int32), ("n_u", C.c_int32), ("n_y", C.c_int32),
                ("n_phi", C.c_int32), ("basis", C.c_int32), ("hg_count", C.c_int32 * PGB_MAX_Z),
                ("hg_stride", C.c_int32 * PGB_MAX_Z), ("hg_L", C.c_double * PGB_MAX_Z), ("N", C.c_int64),
                ("T", C.c_int64), ("n_chains", C.c_int32), ("keep_w_history", C.c_int32), ("x_history", C.c_int32),
                ("precision", C.c_int32)]


X_HISTORY = {"auto": 0, "store": 1, "replay": 2}


class PgoptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pgopt_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None
EXPORTS = ["pgb_version", "pgb_last_error", "pgb_create", "pgb_destroy", "pgb_set_data", "pgb_set_measurement",
           "pgb_set_init_dist", "pgb_set_prior", "pgb_set_prior_dense", "pgb_set_state", "pgb_get_state", "pgb_set_rng_philox",
           "pgb_set_rng_injected", "pgb_sweep", "pgb_sweep_filter", "pgb_sweep_mniw", "pgb_particle_gibbs",
           "pgb_get_sample", "pgb_get_history", "pgb_get_stats", "pgb_get_status", "pgb_synchronize",
           "pgb_systematic_resampling", "pgb_resampling_last_ms", "pgb_mniw_sample", "pgb_mniw_sample_dense", "pgb_rollout", "pgb_test_prediction", "pgb_scenario_constraint_max", "pgb_scenario_constraint_last_ms", "pgb_rollout_last_ms", "pgb_test_force_fallback", "pgb_test_math", "pgb_test_fp64_peak", "pgb_test_rollout_kernel", "pgb_test_dynamics_kernel",
           "pgb_rollout_supplied", "pgb_samples_create", "pgb_samples_destroy", "pgb_samples_last_error", "pgb_samples_set",
           "pgb_samples_get", "pgb_samples_store", "pgb_particle_gibbs_dev", "pgb_samples_draw_xT", "pgb_rollout_dev",
           "pgb_scenarios_get", "pgb_scenario_constraint_max_dev", "pgb_samples_last_ms", "pgb_dynamics_dev", "pgb_dynamics_last_ms", "pgb_host_alloc", "pgb_host_free",
           "pgb_multi_create", "pgb_multi_destroy", "pgb_multi_last_error", "pgb_multi_n_devices", "pgb_multi_set_data",
           "pgb_multi_set_measurement", "pgb_multi_set_init_dist", "pgb_multi_set_prior", "pgb_multi_set_prior_dense", "pgb_multi_set_state",
           "pgb_multi_set_rng_philox", "pgb_packed_record_doubles", "pgb_multi_particle_gibbs", "pgb_multi_get_packed",
           "pgb_multi_get_sample", "pgb_multi_set_samples", "pgb_multi_rollout", "pgb_multi_get_scenarios", "pgb_multi_dynamics", "pgb_multi_last_ms",
           "pgb_samples_pack_bytes", "pgb_samples_pack", "pgb_samples_unpack", "pgb_checkpoint_bytes",
           "pgb_checkpoint_save", "pgb_checkpoint_load",
           "pgb_set_mode", "pgb_timer_start", "pgb_timer_stop", "pgb_profile", "pgb_get_profile", "pgb_get_launch_count"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpgopt_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` or `python pgopt_b200/build.py`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pgb_version.restype = C.c_char_p
        L.pgb_last_error.restype = C.c_char_p
        L.pgb_last_error.argtypes = [C.c_void_p]
        L.pgb_destroy.restype = None
        L.pgb_destroy.argtypes = [C.c_void_p]
        L.pgb_samples_last_error.restype = C.c_char_p
        L.pgb_samples_last_error.argtypes = [C.c_void_p]
        L.pgb_samples_destroy.restype = None
        L.pgb_samples_destroy.argtypes = [C.c_void_p]
        L.pgb_host_alloc.restype = C.c_void_p
        L.pgb_host_alloc.argtypes = [C.c_size_t]
        L.pgb_multi_last_error.restype = C.c_char_p
        L.pgb_multi_last_error.argtypes = [C.c_void_p]
        L.pgb_multi_destroy.restype = None
        L.pgb_multi_destroy.argtypes = [C.c_void_p]
        for f in (L.pgb_packed_record_doubles, L.pgb_samples_pack_bytes, L.pgb_checkpoint_bytes):
            f.restype = C.c_int64
        L.pgb_host_free.restype = None
        L.pgb_host_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = lib().pgb_last_error(handle)
        raise PgoptError(rc, msg.decode() if msg else "unknown")
