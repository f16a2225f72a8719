"""The C-ABI library loads, exports every symbol include/pgopt_b200.h declares, validates arguments on the
host and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pgopt_b200
from pgopt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "pgopt_b200.h")).read()
    declared = set(re.findall(r"\b(pgb_[A-Za-z_0-9]+)\s*\(", hdr))
    declared.discard("pgb_last_error")  # matched below as well; keep explicit
    declared.add("pgb_last_error")
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libpgopt_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == declared
    assert b"sm_100a" in L.pgb_version()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_closures_are_rejected():
    with pytest.raises(TypeError):
        pgopt_b200.api._basis(lambda x, u: x)
    with pytest.raises(TypeError):
        pgopt_b200.api._meas(lambda x, u: x)


def test_config_validation_on_host():
    cfg = _lib.PgbConfig()
    cfg.device, cfg.n_x, cfg.n_u, cfg.n_y, cfg.n_phi, cfg.basis = 0, 2, 1, 1, 7, _lib.BASIS_KNOWN_VB
    cfg.N, cfg.T, cfg.n_chains = 30, 100, 1
    h = C.c_void_p()
    assert _lib.lib().pgb_create(C.byref(h), C.byref(cfg)) == _lib.PGB_ERR_INVALID
    assert b"KNOWN_VB" in _lib.lib().pgb_last_error(None)
    b = pgopt_b200.HilbertGPBasis([5, 5, 5], [10, 20, 20])
    assert b.count == [5, 5, 5] and b.stride == [25, 5, 1] and b.n_phi == 125 and b.n_x == 2
    b = pgopt_b200.HilbertGPBasis([1, 5, 5, 5, 5], [4, 5, 6, 7, 8])
    assert b.count == [5, 5, 5, 5, 1] and b.stride == [125, 25, 5, 1, 1] and b.n_phi == 625 and b.n_x == 4


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(pgopt_b200.PgoptError) as ei:
        pgopt_b200.systematic_resampling(np.ones(4), 4, rnd=0.5)
    assert ei.value.code == _lib.PGB_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(pgopt_b200.PgoptError) as ei:
        pgopt_b200.ParticleGibbsEngine(pgopt_b200.KnownBasisVB(), 1, 30, 100)
    assert ei.value.code == _lib.PGB_ERR_CUDA


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_new_entry_points():
    """the device-resident, multi-GPU and supplied-scenario entry points also refuse to compute without a device, and
    validate what they can on the host"""
    basis = pgopt_b200.KnownBasisVB()
    with pytest.raises(pgopt_b200.PgoptError) as ei:
        pgopt_b200.DeviceSamples(basis, 1, 30, 4)
    assert ei.value.code == _lib.PGB_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(pgopt_b200.PgoptError) as ei:
        pgopt_b200.MultiGPUEngine(basis, 1, 30, 100, 2, [0])
    assert ei.value.code == _lib.PGB_ERR_CUDA
    with pytest.raises(pgopt_b200.PgoptError) as ei:  # n_chains is the TOTAL and must cover the devices
        pgopt_b200.MultiGPUEngine(basis, 1, 30, 100, 1, [0, 1])
    assert ei.value.code == _lib.PGB_ERR_INVALID
    rng = np.random.default_rng(0)
    smp = [pgopt_b200.PG_sample(rng.normal(size=(2, 5)), np.eye(2), np.ones(3) / 3, rng.normal(size=(2, 3)), np.zeros(1))]
    with pytest.raises(pgopt_b200.PgoptError) as ei:
        pgopt_b200.scenario_rollout(smp, basis, pgopt_b200.LinearMeasurement([[1.0, 0.0]]), 0.1, 4, x_vec_0=np.zeros(2))
    assert ei.value.code == _lib.PGB_ERR_CUDA
    cfg = _lib.PgbConfig()
    cfg.device, cfg.n_x, cfg.n_u, cfg.n_y, cfg.n_phi, cfg.basis = 0, 2, 1, 1, 5, _lib.BASIS_KNOWN_VB
    cfg.N, cfg.T, cfg.n_chains, cfg.precision = 30, 100, 1, 7
    h = C.c_void_p()
    assert _lib.lib().pgb_create(C.byref(h), C.byref(cfg)) == _lib.PGB_ERR_INVALID
    assert b"precision" in _lib.lib().pgb_last_error(None)
    assert _lib.lib().pgb_packed_record_doubles(C.c_int32(2), C.c_int32(125), C.c_int32(1)) == 2 * 125 + 4 + 2 + 1
    assert _lib.lib().pgb_samples_pack_bytes(None, C.c_int32(1)) == 0 and _lib.lib().pgb_checkpoint_bytes(None) == 0


def test_sample_batch_packs_like_the_sample_list():
    """PGSampleBatch (stacked arrays) must hand the C ABI exactly the bytes the Vector{PG_sample} mirror does."""
    from pgopt_b200.api import _pack_samples
    rng = np.random.default_rng(0)
    K, nx, nphi, N = 7, 3, 11, 5
    A, Q = rng.normal(size=(K, nx, nphi)), rng.normal(size=(K, nx, nx))
    w, x, um = rng.uniform(size=(K, N)), rng.normal(size=(K, nx, N)), rng.normal(size=(K, 2))
    lst = [pgopt_b200.PG_sample(A[k], Q[k], w[k], x[k], um[k]) for k in range(K)]
    ref = _pack_samples(lst)
    for other in (_pack_samples(pgopt_b200.PGSampleBatch(A, Q, w, x, um)),
                  _pack_samples(pgopt_b200.PGSampleBatch.from_samples(lst))):
        for a, b in zip(ref[:5], other[:5]):
            assert np.array_equal(np.ravel(a), np.ravel(b))
        assert ref[5] == other[5]
    b = pgopt_b200.PGSampleBatch(A, Q, w, x, um)
    assert len(b) == K and np.array_equal(b[3].A, A[3]) and np.array_equal(b[3].x_m1, x[3])

