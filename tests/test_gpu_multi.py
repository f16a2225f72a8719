"""pgb_multi_* (several GPUs of one node driven from one process, in-process ncclCommInitAll): chain-sharded sampler with
the NCCL gather of the packed [A | Q | x_T | u_m1] records and the scenario-sharded rollout with the NCCL gather of
x_vec_0 / X / Y — bit-identical to the single-device path and to the oracle, whatever the device count.  Runs on however
many GPUs are visible (1 device exercises the same code without the collective; gpurun --gpus 2 / 8 the real thing)."""
import numpy as np
import pytest

import _oracle as O
import _problems as P
import pgopt_b200 as pg
from test_gpu_parity import _assert_same, _models

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0]]
    if n >= 2:
        lists.append([0, 1])
    if n >= 3:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("n_chains,N,T,K,K_b,k_d", [(5, 30, 50, 2, 1, 1), (8, 2000, 10, 2, 0, 0)])
def test_multi_chains_gather_equals_single_device(n_chains, N, T, K, K_b, k_d):
    basis, om, _ = _models()["known"]
    u, y, _ = P.make_data(T, 71)
    V = 10 * np.ones(5)
    A0, Q0 = np.zeros((2, 5)), 100 * np.eye(2)
    H = 9
    u_init = np.sin(np.arange(H))[None, :]
    g = pg.LinearMeasurement([[1.0, 0.0]])
    # oracle: chain c with stream id 3 + c; x_T with scenario stream id c*K + k; rollout from the full samples
    ref_rec = np.empty((n_chains, K, 10 + 4 + 2 + 1))
    refs = []
    for c in range(n_chains):
        r = O.particle_gibbs(om, N, T, u, y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), A0, Q0, V, 100 * np.eye(2), 10.0, K, K_b,
                             k_d, seed=21, chain=3 + c)
        refs.append(r)
        for k in range(K):
            xT = O.draw_xT(r["w"][k], r["x"][k], 8, c * K + k)
            ref_rec[c, k] = np.concatenate([r["A"][k].ravel(order="F"), r["Q"][k].ravel(order="F"), xT, u[:, -1]])
    Aall = np.concatenate([r["A"] for r in refs])
    Qall = np.concatenate([r["Q"] for r in refs])
    wall = np.concatenate([r["w"] for r in refs])
    xall = np.concatenate([r["x"] for r in refs])
    um = np.tile(u[:, -1], (n_chains * K, 1))
    roll = O.rollout(om, n_chains * K, H, N, Aall, Qall, wall, xall, um, [[1.0, 0.0]], [[0.1]], u_init, seed=8)
    for devs in _device_lists():
        if len(devs) > n_chains:
            continue
        with pg.MultiGPUEngine(basis, 1, N, T, n_chains, devs) as m:
            m.set_data(u, y)
            m.set_measurement(g, 0.1)
            m.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
            m.set_prior(V, 100 * np.eye(2), 10.0)
            m.set_state(A0, Q0, None, 1)
            m.set_rng_philox(21, 3)
            rec = m.particle_gibbs(K, K_b, k_d, roll_seed=8)
            _assert_same("records on %s" % devs, rec, ref_rec)
            for d in range(len(devs)):  # every device holds all the records after the all-gather
                _assert_same("device %d copy" % d, m.packed_on(d).reshape(rec.shape), ref_rec)
            s = m.get_sample(n_chains - 1, K - 1)
            _assert_same("w_m1 stays with its chain", s.w_m1, refs[-1]["w"][K - 1])
            out = m.rollout(g, 0.1, H, u_init=u_init, seed=8)
            for name in ("x0", "X", "Y"):
                _assert_same("%s on %s" % (name, devs), out[name], roll[name])
            ms = m.last_ms()
            assert ms["rollout_kernel_ms_max"] > 0


def test_multi_scenario_sharded_rollout():
    """configs[4] shape, scaled: host samples dealt to the devices, rolled out where they live, outputs all-gathered."""
    rng = np.random.default_rng(81)
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
    om = O.hilbert_model(4, 1, basis.count, basis.stride, basis.L)
    K, H, N, n_x = 203, 13, 30, 4
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
    ref = O.rollout(om, K, H, N, A, Q, w, x, um, Cm, [[0.1]], u_init, seed=6)
    batch = pg.PGSampleBatch(A, Q, w, np.transpose(x, (0, 2, 1)), um)
    for devs in _device_lists():
        with pg.MultiGPUEngine(basis, 1, N, 2, len(devs), devs) as m:
            m.set_samples(batch, roll_seed=6)
            out = m.rollout(pg.LinearMeasurement(Cm), 0.1, H, u_init=u_init, seed=6)
            for name in ("x0", "X", "Y"):
                _assert_same("%s on %s" % (name, devs), out[name], ref[name])
            for d in range(len(devs)):  # ... and every device holds all of them after the all-gather
                got = m.scenarios_on(d, H)
                for name in ("x0", "X", "Y"):
                    _assert_same("%s gathered on device %d of %s" % (name, d, devs), got[name], ref[name])
            # the stacked dynamics step + Jacobians on the sharded samples (optimal_control_Altro.jl:79-107): every device
            # its own scenarios, with the process noise of the rollout above; == the oracle and == the rollout's next state
            t = 4
            xn, jx, ju, ms = m.dynamics(ref["X"][:, t].reshape(-1), u_init[:, t], t=t, phi_variant=0)
            rx, rjx, rju = O.dynamics(om, A, ref["X"][:, t], u_init[:, t], ref["v"], t, 0)
            _assert_same("x_next on %s" % devs, xn.reshape(K, n_x), rx)
            _assert_same("df/dx on %s" % devs, jx, rjx)
            _assert_same("df/du on %s" % devs, ju, rju)
            _assert_same("x_next == X[t+1] on %s" % devs, xn.reshape(K, n_x), ref["X"][:, t + 1])
            assert ms > 0
