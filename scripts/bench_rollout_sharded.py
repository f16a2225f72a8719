"""Scenario rollout of BASELINE.json configs[4] (K = 10^5 scenarios x H = 41, n_x = 4, n_phi = 625) sharded over the
ranks of one node: `python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_rollout_sharded.py [K]`
(or plain `python` for one GPU).  Strong scaling: K is fixed, rank r rolls the scenarios of `shard_range(K, r, world)`
with k_offset = lo (global stream ids), no collective on the data path; the per-scenario outputs (x0, X, Y: 1.7 KB per
scenario) are all-gathered over NCCL at the end.  Device time = max over ranks of the rollout kernel's CUDA-event
time (pgb_rollout_last_ms); the gather and the host call are reported beside it.  Rank 0 prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pgopt_b200 as pg  # noqa: E402
from pgopt_b200 import multi  # noqa: E402
from pgopt_b200._lib import lib  # noqa: E402


def main():
    import torch
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    H, N = 41, 30
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = multi.shard_range(K, rank, world)
    Kl = hi - lo
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])
    n_x, n_phi = 4, basis.n_phi
    rng = np.random.default_rng([7, rank])  # synthetic posterior samples of this shard
    A = rng.normal(0, 0.05, (Kl, n_x, n_phi))
    B = rng.normal(0, 0.1, (Kl, n_x, n_x))
    Q = B @ np.transpose(B, (0, 2, 1)) + 0.05 * np.eye(n_x)
    w = rng.uniform(size=(Kl, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (Kl, N, n_x))
    um = rng.normal(0, 1, (Kl, 1))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    u_init = np.sin(np.arange(H))[None, :]
    samples = [pg.PG_sample(A[k], Q[k], w[k], x[k].T, um[k]) for k in range(Kl)]
    g = pg.LinearMeasurement(Cm)
    pg.scenario_rollout(samples[:64], basis, g, 0.1, H, u_init=u_init, seed=1, k_offset=lo, device=local)  # warm-up
    kms_all, host_all = [], []
    out = None
    for _ in range(3):
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pg.scenario_rollout(samples, basis, g, 0.1, H, u_init=u_init, seed=1, k_offset=lo, device=local)["raw"]
        host_all.append(time.perf_counter() - t0)
        tms, kms = C.c_double(0), C.c_double(0)
        lib().pgb_rollout_last_ms(C.byref(tms), C.byref(kms))
        kms_all.append(kms.value)
    kms, host = min(kms_all), min(host_all)
    # gather of the per-scenario outputs (no collective inside the rollout itself)
    t0 = time.perf_counter()
    gathered = multi.rollout_sharded(lambda a, b: {n: out[n] for n in ("x0", "X", "Y")}, K, dist, torch.device("cuda", local))
    tg = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([kms, host, tg], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        kms, host, tg = (float(v) for v in tt.tolist())
    if rank == 0:
        assert gathered["X"].shape == (K, H + 1, n_x) and np.isfinite(gathered["X"]).all()
        print(json.dumps({
            "metric": "scenario-steps/s (rollout kernel, max over ranks)", "n_gpus": world, "K": K, "H": H, "n_x": n_x,
            "n_phi": n_phi, "scaling": "strong", "kernel_ms_max_over_ranks": kms,
            "value": K * H / (kms * 1e-3), "unit": "scenario-steps/s",
            "host_call_s_max_over_ranks_incl_packing_h2d_d2h": host, "nccl_gather_s": tg,
            "gathered_bytes": int(sum(gathered[n].nbytes for n in gathered)),
            "note": "synthetic A_k ~ N(0, 0.05^2), Q_k random SPD; rank r owns scenarios shard_range(K, r, world), streams keyed "
                    "by the global scenario id"}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
