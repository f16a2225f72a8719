"""BASELINE.json configs[3]: 64 independent PG chains x N = 2^16 particles (T = 2000, known basis) sharded over the
ranks of one node, with the NCCL gather of the packed posterior samples (A, Q, x_T) at the end —
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_chains_sharded.py [K_kept]`.
K (kept samples per chain) is bounded (default 4 sweeps instead of configs[3]'s 1000) so that the run ends in
seconds; every sweep is a full filter + trace + statistics + MNIW pass.  Strong scaling: the 64 chains are fixed,
rank r runs the chains of shard_range(64, r, world), batched in one handle (chain = CTA group of the persistent kernel).
Time = max over ranks of the wall time of the sampler call (device-synchronised on both sides); rank 0 prints JSON."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402  (problem generator)
import pgopt_b200 as pg  # noqa: E402
from pgopt_b200 import multi  # noqa: E402


def main():
    import torch
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_chains, N, T = 64, 2 ** 16, 2000
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prob = bench.build_problem("known", T, 0)

    def factory(n_local, chain_base):
        eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=n_local, device=local)
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(n_local):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        return eng

    lo, hi = multi.shard_range(n_chains, rank, world)
    # warm-up: one sweep of a single chain (module load, first-launch costs)
    e0 = factory(1, lo)
    e0.set_rng_philox(7, lo)
    e0.sweep(1)
    e0.synchronize()
    e0.close()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pack = multi.run_chains(factory, n_chains, K, 0, 0, 7, dist, torch.device("cuda", local))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    if rank == 0:
        assert pack.shape[0] == n_chains and pack.shape[1] == K and np.isfinite(pack).all()
        print(json.dumps({
            "metric": "PG sweeps/s (all chains; filter + trace + statistics + MNIW + sample store), incl. engine set-up, "
                      "D2H of the kept samples and the NCCL gather of (A, Q, x_T)",
            "n_gpus": world, "n_chains": n_chains, "chains_per_gpu": hi - lo, "N": N, "T": T, "K_kept_per_chain": K,
            "scaling": "strong", "seconds_max_over_ranks": dt, "value": n_chains * K / dt, "unit": "PG sweeps/s",
            "particle_steps_per_s": n_chains * K * N * (T - 1) / dt, "gathered_pack_shape": list(pack.shape)}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
