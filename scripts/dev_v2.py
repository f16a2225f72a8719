"""Development aid: one sweep through a chosen kernel mode against the oracle, reporting WHERE the first difference is
(step, quantity, slot) instead of a bare assertion.  python scripts/dev_v2.py [model N T first mode]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O
import _problems as P
import pgopt_b200 as pg
from test_gpu_parity import _models, _engine

def run(model="known", N=30, T=12, first=0, mode=4):
    basis, om, A = _models()[model]
    u, y, _ = P.make_data(T, 11)
    s = P.injected_streams(N, T, 2, basis.n_phi, 21)
    st = O.make_streams(philox=False, **{kk: s[kk] for kk in ("Z0", "Ures", "Z", "Uanc", "Ustar")})
    xp = np.zeros((2, T))
    if not first:
        xp = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, True, st, 1)["x_prim"]
    k = 1 if first else 2
    ref = O.sweep(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A, P.Q_TRUE, xp, first, st, k)
    with _engine(basis, N, T, u, y, persistent=mode) as e:
        e.set_state(A, P.Q_TRUE, xp, k)
        e.set_rng_injected(s["Z0"], s["Ures"], s["Z"], s["Uanc"], s["Ustar"])
        e.sweep_filter()
        a, x, w = e.get_history()
        stat, nfb, nscan = e.status()
        Phi, Psi, Sig, star = e.get_stats()
    print("model %s N %d T %d first %d mode %d: status %d nfb %d nscan %d star %d/%d" % (model, N, T, first, mode, stat, nfb, nscan, star, ref["star"]))
    ok = True
    for t in range(T):
        bx = np.argwhere(x[t].view(np.uint64) != ref["x"][t].view(np.uint64))
        bw = np.argwhere(w[t].view(np.uint64) != ref["w"][t].view(np.uint64))
        ba = np.argwhere(a[t] != ref["a"][t]) if t >= 1 else []
        if len(bx) or len(bw) or len(ba):
            ok = False
            print(" t=%d: x differs at %d slots (first %s), w at %d (first %s), a at %d (first %s)" % (
                t, len(bx), bx[0] if len(bx) else "-", len(bw), bw[0] if len(bw) else "-", len(ba), ba[0] if len(ba) else "-"))
            if len(bw):
                i = bw[0][0]
                print("   w[%d] got %r ref %r ; sum got %r ref %r" % (i, w[t, i], ref["w"][t, i], w[t].sum(), ref["w"][t].sum()))
            if len(bx):
                i = bx[0][0]
                print("   x[%d] got %r ref %r  a got %d ref %d" % (i, x[t, i], ref["x"][t, i], a[t, i], ref["a"][t, i]))
            break
    print(" OK" if ok else " MISMATCH")
    return ok

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    else:
        for args in [("known", 30, 12, 0, 4), ("known", 30, 12, 1, 4), ("known", 1500, 8, 0, 4), ("hilbert", 1500, 8, 0, 4),
                     ("known", 9000, 6, 0, 4), ("known", 200000, 5, 0, 4), ("hilbert", 300000, 4, 1, 4)]:
            try:
                run(*args)
            except Exception as ex:
                print("EXC", args, repr(ex))
