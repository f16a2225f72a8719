"""N > 1 host logic on the CPU: world_size-2 gloo.  Each rank produces the samples of its own chains (here with
the CPU oracle standing in for the GPU engine), packs (A, Q, x_T) and all-gathers; the result must equal the
single-process run of all chains, i.e. sharding must not change any sample."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as O
import _problems as P
from pgopt_b200 import multi

N_CHAINS, K, N, T = 5, 2, 30, 40


class OracleEngine:
    """duck-typed stand-in for ParticleGibbsEngine.particle_gibbs on the CPU"""
    n_x, n_phi = 2, 5

    def __init__(self, n_local, chain_base):
        self.n_local, self.chain_base = n_local, chain_base
        self.u, self.y, _ = P.make_data(T, 3)

    def set_rng_philox(self, seed, chain_base):
        self.seed, self.chain_base = seed, chain_base

    def particle_gibbs(self, K_, K_b, k_d):
        A_s, Q_s = np.empty((self.n_local, K_, 10)), np.empty((self.n_local, K_, 4))
        w_s, x_s = np.empty((self.n_local, K_, N)), np.empty((self.n_local, K_, 2 * N))
        for i in range(self.n_local):
            r = O.particle_gibbs(O.known_model(), N, T, self.u, self.y, [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2),
                                 np.zeros((2, 5)), 100 * np.eye(2), 10 * np.ones(5), 100 * np.eye(2), 10.0, K_, K_b, k_d,
                                 seed=self.seed, chain=self.chain_base + i)
            A_s[i] = r["A"].transpose(0, 2, 1).reshape(K_, -1)
            Q_s[i] = r["Q"].transpose(0, 2, 1).reshape(K_, -1)
            w_s[i] = r["w"]
            x_s[i] = r["x"].reshape(K_, -1)  # (N, n_x) C-order == (n_x, N) F-order
        return A_s, Q_s, w_s, x_s, self.u[:, -1]

    def particle_gibbs_packed(self, K_, K_b, k_d, roll_seed, k_offset):
        """what ParticleGibbsEngine.particle_gibbs_packed computes on the device: records [A | Q | x_T | u_m1] with the
        star drawn from the scenario's Philox stream (optimal_control_Ipopt.jl:58-59)"""
        A_s, Q_s, w_s, x_s, u_m1 = self.particle_gibbs(K_, K_b, k_d)
        rec = np.empty((self.n_local, K_, 10 + 4 + 2 + 1))
        for i in range(self.n_local):
            for k in range(K_):
                xT = O.draw_xT(w_s[i, k], x_s[i, k].reshape(N, 2), roll_seed, k_offset + i * K_ + k)
                rec[i, k] = np.concatenate([A_s[i, k], Q_s[i, k], xT, u_m1])
        return rec

    def close(self):
        pass


def _worker(rank, world, port, out_dir, n_chains):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = multi.run_chains(OracleEngine, n_chains, K, 3, 1, seed=9, dist=dist, roll_seed=4, record_doubles=17)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), got)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (1, 5, 8, 64, 100000):
        for world in (1, 2, 3, 8):
            blocks = [multi.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_chains", [N_CHAINS, 1])
def test_two_ranks_gather_equals_single_process(tmp_path, n_chains):
    """n_chains = 1: the second rank owns no chain, creates no engine and still takes part in the gather"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_chains), nprocs=2, join=True)
    ref = multi.run_chains(OracleEngine, n_chains, K, 3, 1, seed=9, dist=None, roll_seed=4)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == (n_chains, K, 10 + 4 + 2 + 1)
        np.testing.assert_array_equal(got, ref)
    A, Q, xT, um = multi.unpack_samples(ref, 2, 5)
    assert A.shape == (n_chains, K, 10) and Q.shape == (n_chains, K, 4) and xT.shape == (n_chains, K, 2) and um.shape == (n_chains, K, 1)


# ------------------------------------------------------------------------------------------ scenario sharding
R_K, R_H, R_N = 7, 9, 12


def _rollout_inputs():
    rng = np.random.default_rng(5)
    A = rng.normal(0, 0.3, (R_K, 2, 5))
    Q = np.stack([(lambda B: B @ B.T + 0.05 * np.eye(2))(rng.normal(0, 0.2, (2, 2))) for _ in range(R_K)])
    w = rng.uniform(size=(R_K, R_N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (R_K, R_N, 2))
    um = rng.normal(0, 1, (R_K, 1))
    return A, Q, w, x, um, rng.normal(0, 1, (1, R_H))


def _oracle_rollout(lo, hi):
    A, Q, w, x, um, u_init = _rollout_inputs()
    r = O.rollout(O.known_model(), hi - lo, R_H, R_N, A[lo:hi], Q[lo:hi], w[lo:hi], x[lo:hi], um[lo:hi], [[1.0, 0.0]], [[0.1]],
                  u_init, seed=17, k_offset=lo)
    assert r["status"] == 0
    return {k: r[k] for k in ("x0", "v", "e", "X", "Y")}


def _rollout_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = multi.rollout_sharded(_oracle_rollout, R_K, dist=dist)
        np.savez(os.path.join(out_dir, "roll%d.npz" % rank), **got)
    finally:
        dist.destroy_process_group()


def test_two_ranks_scenario_rollout_equals_single_process(tmp_path):
    """Scenarios sharded over two ranks (stream ids offset by the shard start) reproduce the unsharded rollout."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rollout_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = _oracle_rollout(0, R_K)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "roll%d.npz" % r))
        for k in ref:
            assert got[k].shape == ref[k].shape
            np.testing.assert_array_equal(got[k], ref[k])
