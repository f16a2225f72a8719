"""pgb_math.h on the host: accuracy vs mpmath and the no-contraction self test."""
import mpmath as mp
import numpy as np

import _oracle as O

mp.mp.prec = 120


def _max_ulp(y, x, f):
    worst = 0.0
    for xi, yi in zip(x, y):
        t = f(mp.mpf(float(xi)))
        if t == 0:
            continue
        worst = max(worst, abs(float((mp.mpf(float(yi)) - t) / mp.mpf(float(np.spacing(abs(float(t))))))))
    return worst


def test_selftest_no_contraction():
    assert O.lib().orc_selftest() == 0


def test_exp_accuracy():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-745, 709, 3000), rng.uniform(-50, 0, 3000), rng.uniform(-1, 1, 1000)])
    assert _max_ulp(O.vfun("exp", x), x, mp.exp) < 1.0
    y = O.vfun("exp", np.array([0.0, -746.5, 710.0, -745.13]))
    assert y[0] == 1.0 and y[1] == 0.0 and np.isinf(y[2]) and y[3] == 5e-324


def test_log_accuracy():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(0, 1, 3000), 10.0 ** rng.uniform(-300, 300, 1000), rng.uniform(0.5, 2, 2000)])
    assert _max_ulp(O.vfun("log", x), x, mp.log) < 1.0


def test_sincos_accuracy():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-20, 20, 3000), rng.uniform(-1e5, 1e5, 3000), rng.uniform(-1, 1, 1000)])
    assert _max_ulp(O.vfun("sin", x), x, mp.sin) < 2.0
    assert _max_ulp(O.vfun("cos", x), x, mp.cos) < 2.0


def test_normals_moments():
    z = O.normal_pairs(42, 2, 0, 1, 5, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.05
    # counter-based: same arguments, same values; different chain, different values
    assert np.array_equal(z[:100], O.normal_pairs(42, 2, 0, 1, 5, 50))
    assert not np.array_equal(z[:100], O.normal_pairs(42, 2, 1, 1, 5, 50))
