"""Multi-GPU plumbing: independent PG chains (and independent scenarios) sharded over ranks, one process per GPU.

The within-chain filter does not shard, so there is no collective on the data path: every rank runs
its own chains with their own Philox stream ids and only the kept samples — packed (A, Q, x_T) per sample, about
2 KB — are gathered at the end with one all-gather (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def unpack_samples(pack, n_x, n_phi, n_u=1):
    """packed records [A | Q | x_T | u_m1] (include/pgopt_b200.h: pgb_packed_record_doubles) -> (A, Q, x_T, u_m1)"""
    a, q = n_x * n_phi, n_x * n_x
    return pack[..., :a], pack[..., a:a + q], pack[..., a + q:a + q + n_x], pack[..., a + q + n_x:a + q + n_x + n_u]


def gather_packed(local_pack, n_total, dist=None, device=None):
    """All-gather the per-rank packs (uneven shards are padded to the largest) -> (n_total, K, P) on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_pack
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    nmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    K, P = local_pack.shape[1], local_pack.shape[2]
    buf = torch.zeros((nmax, K, P), dtype=torch.float64, device=device)
    buf[:local_pack.shape[0]] = torch.from_numpy(local_pack).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def rollout_sharded(rollout_fn, K, dist=None, device=None):
    """Scenario rollout sharded over the process group (BASELINE.json configs[4]): rank r rolls the scenarios
    [lo, hi) of `shard_range(K, r, world)` with `rollout_fn(lo, hi)` — which must key its random streams by the global
    scenario id (`k_offset = lo`) — and the per-scenario outputs are all-gathered, so that every rank ends with exactly
    the arrays of a single-process rollout.  rollout_fn returns a dict of arrays whose first axis is the scenario."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = shard_range(K, rank, world)
    local = rollout_fn(lo, hi)
    if world == 1:
        return local
    out = {}
    for name in sorted(local):
        a = np.ascontiguousarray(local[name], dtype=np.float64)
        flat = a.reshape(a.shape[0], 1, -1)
        out[name] = gather_packed(flat, K, dist, device).reshape((K,) + a.shape[1:])
    return out


def run_chains(engine_factory, n_chains, K, K_b, k_d, seed, dist=None, device=None, roll_seed=0, record_doubles=None):
    """Run `n_chains` independent chains sharded over the process group (one process per GPU) and gather the packed
    [A | Q | x_T | u_m1] records.  engine_factory(n_local, chain_base) -> a configured ParticleGibbsEngine (data, prior,
    state set).  x_T = x_m1[:, star] is drawn on the rank that owns the chain, on the device, from the scenario's own
    Philox stream (roll_seed, id = global sample index) — the same star pgb_rollout would draw — so the N-sized particle
    arrays never travel.  A rank that owns no chain creates no engine and contributes an empty block."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = shard_range(n_chains, rank, world)
    if hi > lo:
        eng = engine_factory(hi - lo, lo)
        try:
            eng.set_rng_philox(seed, lo)
            local = np.ascontiguousarray(eng.particle_gibbs_packed(K, K_b, k_d, roll_seed, lo * K), dtype=np.float64)
        finally:
            eng.close()
    else:
        if record_doubles is None:
            raise ValueError("a rank without chains needs record_doubles to size its empty block")
        local = np.zeros((0, K, record_doubles))
    return gather_packed(local, n_chains, dist, device)
