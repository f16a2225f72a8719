"""Loader for reference-held golden vectors written by julia/make_fixtures.jl (tests/golden/ref_<case>/*.npy, or the
same arrays zipped as tests/golden/ref_<case>.npz).  None exist in this repository: the build image has no Julia, so
the reference cannot run here (parity UNPINNED); the tests that use this module skip themselves and say so."""
import glob
import os

import numpy as np

import _oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SKIP_REASON = ("no tests/golden/ref_* fixtures: the reference (Julia) cannot run in this image; run julia/make_fixtures.jl "
               "with Julia 1.10 + the PGopt checkout to produce them — until then parity stays UNPINNED")


def fixtures():
    return sorted(p for p in glob.glob(os.path.join(GOLDEN, "ref_*")) if os.path.isdir(p) or p.endswith(".npz"))


def load(path):
    if os.path.isdir(path):
        return {os.path.basename(f)[:-4]: np.load(f) for f in glob.glob(os.path.join(path, "*.npy"))}
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def scalar(g, k):
    return float(np.ravel(g[k])[0])


def oracle_model(g):
    if int(scalar(g, "basis_is_known")):
        return O.known_model(), None
    dims, L = [int(v) for v in np.ravel(g["n_phi_dims"])], [float(v) for v in np.ravel(g["L"])]
    nz = len(dims)
    count = [dims[nz - 1 - k] for k in range(nz)]
    stride = [int(np.prod(dims[:nz - 1 - k])) for k in range(nz)]
    return O.hilbert_model(int(scalar(g, "n_x")), int(scalar(g, "n_u")), count, stride, L), (dims, L)


def sweep_inputs(g, k):
    """streams and (A, Q) of sweep k (1-based) in the oracle's / the C ABI's injected layout"""
    s = dict(Z0=np.ascontiguousarray(g["Z0_%d" % k]), Ures=np.ravel(g["Ures_%d" % k]),
             Z=np.ascontiguousarray(g["Z_%d" % k]), Uanc=np.ravel(g["Uanc_%d" % k]), Ustar=scalar(g, "Ustar_%d" % k))
    return s, np.asarray(g["A_%d" % k]), np.asarray(g["Q_%d" % k])
