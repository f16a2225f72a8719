#!/usr/bin/env python
"""bench.py — CPF-AS particle-steps/s (N = 2^20, T = 2000) and PG sweeps/s on B200.

A "step" is one conditional particle-Gibbs sweep (k > 1 branch of Julia/src/particle_Gibbs.jl:168-200:
T-1 resample / propagate / ancestor-sample / weight steps over N particles, then trace + statistics)
on synthetic data shaped like BASELINE.json configs[2]: the reduced-rank-GP ("generic") basis model,
n_x = 2, n_phi = 125, single GPU, counter-based (Philox) random streams.  With --gpus N > 1 every rank
runs its own independent chain of the same size (weak scaling, no collective on the data path); the
end-to-end leg gathers the packed (A, Q, x_T) samples over NCCL.

  value : particle-steps/s of the filter phase, inputs resident in HBM, CUDA-event timed
  e2e   : the same metric through the public API with HOST buffers (pinned) — u, y, A, Q, x_prim go
          H2D and the PG_sample (A, Q, w_m1, x_m1) comes back D2H inside the timed region, and the
          sweep includes the MNIW draw
  --impl reference : the CPU oracle (restatement of the reference; Julia is not installable here) on
          a bounded sample of the same workload, one host thread like the reference's sampler.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_PARTICLE_STEP = lambda n_x, f32=False: (8 * n_x + 12) if f32 else (16 * n_x + 20)  # x r+w, w r+w, Int32 ancestor
FLOPS_PER_PARTICLE_STEP = {"generic": 1250.0, "known": 100.0}
# dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel (k_sweep_v2) per particle-step, from the committed
# `ncu --set full` captures at N = 2^20, T = 24 (profiles/r02_v2_sweep_kernel_N2p20.md)
NCU_DRAM_BYTES_PER_PARTICLE_STEP = {"generic": (3.514112e6 + 891.111936e6) / (23 * 2 ** 20),
                                    "known": (1.25184e6 + 798.978304e6) / (23 * 2 ** 20)}


# ------------------------------------------------------------------------------------------- workload
def f_true(x, u):
    return np.array([0.8 * x[0] - 0.5 * x[1] + 0.1 * np.cos(3 * x[0]) * x[1],
                     0.4 * x[0] + 0.5 * x[1] + (1 + 0.3 * np.sin(2 * x[1])) * u])


def make_data(T, seed):
    rng = np.random.default_rng(seed)
    Lq = np.linalg.cholesky(np.array([[0.03, -0.004], [-0.004, 0.01]]))
    u = rng.normal(0, 3, (1, T))
    x = np.zeros((2, T + 1))
    x[:, 0] = rng.normal(2, 1, 2)
    y = np.zeros((1, T))
    for t in range(T):
        x[:, t + 1] = f_true(x[:, t], u[0, t]) + Lq @ rng.normal(size=2)
        y[0, t] = x[0, t] + 0.1 * rng.normal()
    return u, y, x


def hilbert_phi_np(count, L, x, u):
    """vectorised NumPy statement of the reduced-rank GP basis (only used to build a plausible A)"""
    z = np.vstack([u, x])
    nz = len(L)
    phi = np.ones((1, z.shape[1]))
    for k in range(nz):
        j = np.arange(1, count[k] + 1)[:, None]
        tk = (1 / np.sqrt(L[k])) * np.sin((np.pi * j / (2 * L[k])) * (z[k][None, :] + L[k]))
        phi = (phi[:, None, :] * tk[None, :, :]).reshape(-1, z.shape[1])
    return phi


def build_problem(model, T, seed=0):
    import pgopt_b200 as pg
    u, y, x = make_data(T, seed)
    if model == "known":
        basis = pg.KnownBasisVB()
        A = np.array([[8.0, -5.0, 0.0, 10.0, 0.0], [4.0, 5.0, 1.0, 0.0, 3.0]])
        V = 10 * np.ones(5)
    else:
        basis = pg.HilbertGPBasis([5, 5, 5], [10.0, 20.0, 20.0])
        Phi = hilbert_phi_np(basis.count, basis.L, x[:, :T], u)
        V = pg.hilbert_gp_prior_V([5, 5, 5], [10, 20, 20], 2 * np.pi * np.ones(3), 100 ** 2 / (8 * np.pi ** 4))
        # what the sampler itself would produce given the true trajectory: the MNIW posterior mean of A
        # (particle_Gibbs.jl:64-67) and the residual covariance as Q
        A = np.linalg.solve(Phi @ Phi.T + np.diag(1 / V), Phi @ x[:, 1:T + 1].T).T
        return dict(basis=basis, u=u, y=y, A=A, Q=np.cov(x[:, 1:T + 1] - A @ Phi), V=V, x_prim=x[:, :T].copy(),
                    n_phi=basis.n_phi)
    Q = np.array([[0.05, 0.0], [0.0, 0.03]])
    return dict(basis=basis, u=u, y=y, A=A, Q=Q, V=V, x_prim=x[:, :T].copy(), n_phi=basis.n_phi)


# ------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


PROF_NAMES = ["k_prep", "k_norm", "k_scan_local", "k_resolve", "k_sweep_v2 (whole T loop)", "k_seq_exact",
              "k_expand_fallback", "k_weights", "k_build_utabs", "tail(trace+stats)", "k_mniw"]
PHASE_NAMES = ["init", "A prep (w, F, waN)", "wait B1", "B scan-local main", "B normalise side", "wait B2", "C resolve",
               "C exact prefixes + expansion", "C ancestor shortcut", "wait B3", "D ancestor index", "D heavy + weights",
               "wait B4", "fix-up test"]


def dist_setup(n_gpus):
    if n_gpus <= 1 or "RANK" not in os.environ:
        return None, 0, 1, 0
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, rank, world, local


# ------------------------------------------------------------------------------------------- GPU arm
def time_filter(eng, prob, steps, warmup, sweep_index=2):
    """device-resident timing of pgb_sweep_filter; returns list of ms"""
    from pgopt_b200._lib import lib, check
    ms = []
    for i in range(warmup + steps):
        eng.set_state(prob["A"], prob["Q"], prob["x_prim"], sweep_index)   # untimed: same state every step
        check(lib().pgb_timer_start(eng._h), eng._h)
        eng.sweep_filter()
        t = C.c_double(0)
        check(lib().pgb_timer_stop(eng._h, C.byref(t)), eng._h)
        if i >= warmup:
            ms.append(t.value)
    return ms


def run_gpu(args):
    import torch
    import pgopt_b200 as pg
    from pgopt_b200._lib import lib, check
    dist, rank, world, local = dist_setup(args.gpus)
    N, T = args.particles, args.T
    prob = build_problem(args.model, T, seed=0)
    n_x, n_u, n_y, n_phi = 2, 1, 1, prob["n_phi"]
    f32 = args.dtype == "f32"
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=1, device=local, precision=args.dtype)
    eng.set_data(prob["u"], prob["y"])
    eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
    eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
    eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
    eng.set_rng_philox(1234, rank + args.chain_base)

    def barrier():
        eng.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident timing (value)
    n0 = C.c_int64(0)
    for _ in range(args.warmup):
        eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2)
        eng.sweep_filter()
    barrier()
    check(lib().pgb_get_launch_count(eng._h, C.byref(n0)), eng._h)
    clk = ClockSampler(local)
    clk.start()
    ms = time_filter(eng, prob, args.steps, 0)
    barrier()
    clocks = clk.stop()
    n1 = C.c_int64(0)
    check(lib().pgb_get_launch_count(eng._h, C.byref(n1)), eng._h)
    t_total = sum(ms) / 1e3
    per_rank = None
    if dist:
        tt = torch.tensor([t_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = float(tt.item())
        # every rank's own step times and clocks (a single throttled GPU sets the max-over-ranks time: it must be visible)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "ms_per_step": [round(v, 2) for v in ms], "sm_mhz": clocks["sm_mhz"],
                                          "reasons": clocks["reasons"]})
        clocks = dict(clocks)
        clocks["reasons"] = sorted({r for p_ in per_rank for r in p_["reasons"]})
        mhz = [p_["sm_mhz"] for p_ in per_rank if p_["sm_mhz"]]
        clocks["sm_mhz_min_over_ranks"] = min(mhz) if mhz else None
    psteps = N * (T - 1) * args.steps
    value = world * psteps / t_total
    st, nfb, nscan = eng.status(0, clear=True)

    # ---- per-kernel profile (separate pass: event pairs around every launch perturb the step time)
    check(lib().pgb_profile(eng._h, 1), eng._h)
    eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2)
    eng.sweep_filter()
    launches = (C.c_int64 * len(PROF_NAMES))()
    pms = (C.c_double * len(PROF_NAMES))()
    check(lib().pgb_get_profile(eng._h, launches, pms), eng._h)
    check(lib().pgb_profile(eng._h, 0), eng._h)
    prof = {PROF_NAMES[i]: {"launches": int(launches[i]), "ms": float(pms[i])} for i in range(len(PROF_NAMES))
            if launches[i]}
    tot_prof = sum(v["ms"] for v in prof.values()) or 1.0
    if dist:  # which kernels every rank actually ran (a shape that does not fit the sweep kernel falls back silently)
        kr = [None] * world
        dist.all_gather_object(kr, {k: [v["launches"], round(v["ms"], 2)] for k, v in prof.items()})
        for p_, k_ in zip(per_rank, kr):
            p_["kernels"] = k_
    dom = max(prof, key=lambda k: prof[k]["ms"])
    fp64_peak = C.c_double(0)
    check(lib().pgb_test_fp64_peak(local, C.byref(fp64_peak)))

    # ---- end to end through the public API with pinned host buffers
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        return t, t.numpy()
    keep = [pinned(prob["u"]), pinned(prob["y"]), pinned(np.asfortranarray(prob["A"]).ravel(order="F")),
            pinned(np.asfortranarray(prob["Q"]).ravel(order="F")), pinned(np.asfortranarray(prob["x_prim"]).ravel(order="F"))]
    hu, hy, hA, hQ, hxp = [k[1] for k in keep]
    outA, outQ = pinned(np.zeros(n_x * n_phi)), pinned(np.zeros(n_x * n_x))
    outw, outx = pinned(np.zeros(N)), pinned(np.zeros(n_x * N))
    outu = np.zeros(n_u)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    h2d = (hu.nbytes + hy.nbytes + hA.nbytes + hQ.nbytes + hxp.nbytes)
    d2h = outA[1].nbytes + outQ[1].nbytes + outw[1].nbytes + outx[1].nbytes

    def e2e_step():
        check(lib().pgb_set_data(eng._h, p(hu), p(hy)), eng._h)
        check(lib().pgb_set_state(eng._h, 0, p(hA), p(hQ), p(hxp), C.c_int64(2)), eng._h)
        check(lib().pgb_sweep(eng._h, C.c_int64(1)), eng._h)
        check(lib().pgb_get_sample(eng._h, 0, p(outA[1]), p(outQ[1]), p(outw[1]), p(outx[1]), p(outu)), eng._h)
        if dist:  # gather of the packed (A, Q, x_T) sample of every chain
            pack = torch.from_numpy(np.concatenate([outA[1], outQ[1], outx[1][:n_x]])).cuda()
            allp = [torch.empty_like(pack) for _ in range(world)]
            dist.all_gather(allp, pack)
            _ = allp[0].cpu()

    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    te = time.perf_counter() - t0
    if dist:
        tt = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    e2e_value = world * psteps / te
    sweeps_per_s = world * args.steps / te

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(args.model, args.cpu_particles, args.cpu_T, f32=f32)

    if rank == 0:
        peak, peak_src = measured_peaks()
        bps = BYTES_PER_PARTICLE_STEP(n_x, f32)
        t_per_sweep = t_total / args.steps
        achieved = bps * N * (T - 1) / t_per_sweep / 1e9
        out = {
            "metric": "CPF-AS particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_per_sweep, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "CPF-AS conditional sweep (k>1), %s-basis model n_x=2 n_phi=%d, N=%d particles, T=%d, "
                                   "Philox streams, ancestor history stored in HBM, x history %s; BASELINE.json configs[2]"
                                   % (args.model, n_phi, N, T, "stored" if N * T * n_x * 8 <= 2 ** 30 else
                                      "replayed (two rows kept, traced lineage recomputed)"),
                       "N": N, "T": T, "n_phi": n_phi, "chains_per_gpu": 1,
                       "l2": "inputs of a step (x_{t-1}, lw, F: 40 MB at N = 2^20) are re-read from L2 every step; the ancestor "
                             "history (%d GB) is written once, larger than L2" % (N * T * 4 >> 30)},
            "pg_sweeps_per_s": sweeps_per_s,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "includes": "set_data+set_state H2D, filter+trace+stats+MNIW, get_sample D2H"},
            "gpu_launches": int(n1.value - n0.value),
            "clocks": clocks,
            "per_rank": per_rank,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_PARTICLE_STEP[args.model] * N * (T - 1) / 1e9 if (N == 2 ** 20 and not f32) else None,
                         "traffic_unit": "GB per launch (ncu DRAM bytes per particle-step of the FP64 kernel at T=24, x this launch's "
                                         "particle-steps); below the algorithmic bytes: x_{t-1} / F / lw are re-read out of L2, the per-CTA "
                                         "weight arrays never leave shared memory", "peak_source": peak_src,
                         "kernel": "k_sweep_v2 (chunk-stationary): one cooperative launch = one whole conditional sweep (T-1 steps x 4 grid "
                                   "barriers); %d launches per sweep in total" % ((n1.value - n0.value) // max(1, args.steps)),
                         "algorithmic_bytes_per_particle_step": bps,
                         "dominant_kernel": dom, "dominant_share": prof[dom]["ms"] / tot_prof,
                         "fp64": {"algorithmic_tflops_achieved": FLOPS_PER_PARTICLE_STEP[args.model] * N * (T - 1) / t_per_sweep / 1e12,
                                  "dfma_peak_tflops_measured": fp64_peak.value,
                                  "flops_per_particle_step": FLOPS_PER_PARTICLE_STEP[args.model]},
                         "note": "measured (ncu, profiles/r02_v2_sweep_kernel_N2p20.md): instruction-bound — 2 600 (generic) / 1 600 (known) "
                                 "thread instructions per particle-step (bit-defined exp/log/sin/cos, Philox, exact divisions, "
                                 "exact-scan integer algebra), issue slots 55 % busy, DRAM ~10 % busy"},
            "kernel_profile_ms_per_sweep": prof,
            "resampling": {"status_bits": st, "exact_scans": nscan, "serial_fallbacks": nfb},
        }
        if cpu:
            out["cpu_baseline"] = cpu
    eng.close()
    eng = None
    if dist:  # the side measurements below run after every rank has released its GPU: no collective is pending
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank == 0:
        if not args.no_secondary:
            if world == 1:
                try:
                    if args.model == "generic":
                        out["roofline"]["known"] = known_basis_roofline(N, T, max(2, args.steps), f32)
                    out["roofline"]["resampling_only"] = resampling_only()
                except Exception as e:
                    out["roofline"]["known"] = {"error": repr(e)}
                try:
                    out["secondary"] = secondary_measurements(fp64_peak.value)
                except Exception as e:  # never let the side measurements take the headline line down
                    out["secondary"] = {"error": repr(e)}
            else:
                time.sleep(2.0)  # the other ranks are exiting
                out["secondary"] = {}
            try:
                out["secondary"].update(sharded_secondary(world))
            except Exception as e:
                out["secondary"]["sharded"] = {"error": repr(e)}
        print(json.dumps(out))


def secondary_measurements(fp64_peak_tflops=None):
    """The other SURVEY.md section-8 rows, bounded: scenario rollout at the BASELINE.json configs[4] shape
    (K scaled to 20000) and chain-batched PG sweeps at the configs[3] shape (8 chains per GPU, one sweep)."""
    import pgopt_b200 as pg
    from pgopt_b200._lib import lib
    res = {}
    K, H = 20000, 41
    p = _c5_problem(K)
    n_x, n_phi, Np = p["n_x"], p["n_phi"], p["N"]
    g = pg.LinearMeasurement(p["Cm"])
    # device-resident pipeline (pgb_samples_*): the samples are uploaded once (in the real pipeline pgb_particle_gibbs_dev leaves
    # them there), the timed host call is rollout + D2H of x_vec_0 / X / Y into pinned arrays; then screening and one
    # stacked-dynamics + Jacobians step on the same resident set
    out = dict(x0=pg.pinned_empty((K, n_x)), X=pg.pinned_empty((K, H + 1, n_x)), Y=pg.pinned_empty((K, H, 1)))
    with pg.DeviceSamples.from_samples(p["batch"], p["basis"]) as ds:
        calls = []
        for i in range(4):
            t0 = time.perf_counter()
            ds.rollout(g, 0.1, H, u_init=p["u_init"], seed=1).scenarios(want=("x0", "X", "Y"), out=out)
            calls.append(time.perf_counter() - t0)
        ms = ds.last_ms()
        y_min, y_max = np.full((1, H), -np.inf), np.full((1, H), np.inf)
        y_min[0, 20:26] = 2.0
        t0 = time.perf_counter()
        ds.constraint_max(y_min, y_max)
        t_screen = time.perf_counter() - t0
        ms2 = ds.last_ms()
        kd, kd0 = C.c_double(0), C.c_double(0)
        for rep in range(2):  # the second call is the measured one (first: lazy allocation of the output buffers)
            t0 = time.perf_counter()
            ds.dynamics(out["X"][:, 5].reshape(-1), p["u_init"][:, 5], t=5, phi_variant=1)
            t_dyn = time.perf_counter() - t0
            lib().pgb_dynamics_last_ms(ds._s, C.byref(kd))
            ds.dynamics(out["X"][:, 5].reshape(-1), p["u_init"][:, 5], t=5, phi_variant=1, jacobians=False)
            lib().pgb_dynamics_last_ms(ds._s, C.byref(kd0))
    kms = ms["rollout_kernel_ms"]
    bytes_alg = K * 8 * (n_x * n_phi + n_x * n_x + Np + n_x + 1 + (2 * H + 2) * n_x + 2 * H)
    res["rollout"] = {"workload": "K=%d scenarios x H=%d, n_x=4, n_phi=625 (configs[4] shape, K scaled), device-resident samples" % (K, H),
                      "kernel_ms": kms, "scenario_steps_per_s_kernel": K * H / (kms * 1e-3),
                      "hbm_GBps_algorithmic": bytes_alg / (kms * 1e-3) / 1e9,
                      "host_call_ms_rollout_plus_d2h_pinned": 1e3 * min(calls[1:]),
                      "screening_kernel_ms": ms2["screening_kernel_ms"], "screening_host_call_ms": 1e3 * t_screen,
                      "dynamics_jacobians_kernel_ms": kd.value, "dynamics_host_call_ms": 1e3 * t_dyn,
                      "dynamics_hbm_GBps": K * 8 * (n_x * n_phi + 2 * n_x + n_x * (n_x + 1)) / (kd.value * 1e-3) / 1e9 if kd.value else None,
                      "dynamics_state_only_kernel_ms": kd0.value,
                      "dynamics_state_only_hbm_GBps": K * 8 * (n_x * n_phi + 2 * n_x) / (kd0.value * 1e-3) / 1e9 if kd0.value else None,
                      "dynamics_note": "k_dynamics_bulk: A_k staged by cp.async.bulk (HBM-bound kernel; K = 2*10^4 here is a short launch, "
                                       "profiles/r02b_dynamics_bulk_K1e5.json has K = 10^5: 83 % / 99 % of the measured HBM peak)"}
    # FP64 side of the roofline (2 n_x n_phi + ~20 sines + ~780 multiplications ~ 6.6 kflop
    # per scenario-step; AI ~ 12 flop/B, so the FP64 pipe binds before HBM does)
    flops = 6600.0 * K * (H + 1)
    res["rollout"]["fp64_tflops_algorithmic"] = flops / (kms * 1e-3) / 1e12
    if fp64_peak_tflops:
        res["rollout"]["fp64_frac_of_measured_dfma_peak"] = res["rollout"]["fp64_tflops_algorithmic"] / fp64_peak_tflops
    peak_hbm, _ = measured_peaks()
    res["rollout"]["hbm_frac"] = res["rollout"]["hbm_GBps_algorithmic"] / peak_hbm
    nc, Nc, Tc = 8, 2 ** 16, 2000
    prob = build_problem("known", Tc, 0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, Nc, Tc, n_chains=nc)
    try:
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(nc):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        eng.set_rng_philox(7, 0)
        eng.sweep(1)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.sweep(2)
        eng.synchronize()
        dt = time.perf_counter() - t0
        res["chain_batch"] = {"workload": "%d chains x N=%d, T=%d on one GPU (configs[3] shape), known basis, "
                                          "filter+trace+statistics+MNIW" % (nc, Nc, Tc),
                              "pg_sweeps_per_s_all_chains": nc * 2 / dt, "particle_steps_per_s": nc * 2 * Nc * (Tc - 1) / dt,
                              "status": list(eng.status(0))}
    finally:
        eng.close()
    # the reference's own examples (BASELINE.json configs[0], configs[1]): N = 30 particles, T = 2000, on the
    # warp-per-chain kernel; PG sweeps/s (filter + trace + statistics + MNIW) for one chain and for 256 batched chains
    ex = {}
    for model in ("known", "generic"):
        ex[model] = {"one_chain": _small_config(model, 1, 10), "256_chains_batched": _small_config(model, 256, 5)}
    res["reference_examples_N30_T2000_pg_sweeps_per_s"] = ex
    return res


def resampling_only():
    """K3 alone: systematic_resampling (particle_Gibbs.jl:17-34) of N weights into N indices —
    tile sums, exact sequential-order scan, index search — device time from CUDA events inside the call.  Algorithmic
    bytes: 8 read (W) + 4 written (Int32 index on the device) per particle."""
    from pgopt_b200._lib import lib, check
    peak, _ = measured_peaks()
    res = {}
    rng = np.random.default_rng(3)
    for lg in (20, 22):
        n = 1 << lg
        W = rng.uniform(size=n) ** 2
        idx = np.empty(n, dtype=np.int64)
        st, ms, best = C.c_int32(0), C.c_double(0), None
        for rep in range(4):
            check(lib().pgb_systematic_resampling(C.c_int32(0), W.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(n),
                                                  C.c_double(0.37), idx.ctypes.data_as(C.c_void_p), C.byref(st)))
            check(lib().pgb_resampling_last_ms(C.byref(ms)))
            if rep:
                best = ms.value if best is None else min(best, ms.value)
        gbs = 12.0 * n / (best * 1e-3) / 1e9
        res["N=2^%d" % lg] = {"kernel_ms": best, "GBps_algorithmic_12B_per_particle": gbs, "hbm_frac": gbs / peak,
                              "status": int(st.value), "sorted": bool(np.all(np.diff(idx) >= 0))}
    return res


def known_basis_roofline(N, T, steps, f32):
    """the HBM-bound variant of the headline (known basis, n_phi = 5): same kernel, same N and T, device-timed"""
    import pgopt_b200 as pg
    prob = build_problem("known", T, seed=0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=1, precision="f32" if f32 else "f64")
    try:
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        eng.set_rng_philox(1234, 0)
        ms = time_filter(eng, prob, steps, 2)
    finally:
        eng.close()
    peak, _ = measured_peaks()
    bps = BYTES_PER_PARTICLE_STEP(2, f32)
    t = float(np.mean(ms)) * 1e-3
    return {"value": N * (T - 1) / t, "unit": "particle-steps/s", "ms_per_step": 1e3 * t, "achieved": bps * N * (T - 1) / t / 1e9,
            "frac": bps * N * (T - 1) / t / 1e9 / peak, "algorithmic_bytes_per_particle_step": bps,
            "workload": "known basis (n_phi = 5), N=%d, T=%d, %s" % (N, T, "f32" if f32 else "f64")}


def sharded_secondary(n_gpus):
    """BASELINE.json configs[3] / configs[4] through pgb_multi_* on the node's n_gpus devices, each as its own bounded
    `bench.py --workload c4|c5` child process (one process, in-process NCCL; no torch.distributed), so that the driver's
    scaling run carries the two sharded workloads.  Returns {"c4": line, "c5": line}; a failure or a time-out is
    reported in place of the line and never touches the headline."""
    import subprocess
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
                                                             "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE",
                                                             "TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT",
                                                             "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_USE_AGENT_STORE")}
    res = {}
    for wl, extra in (("c5", ["--steps", "3", "--warmup", "1"]), ("c4", ["--kept", "4"])):
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", wl, "--gpus", str(n_gpus)] + extra
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            res[wl] = json.loads(lines[-1]) if lines else {"error": "rc=%d %s" % (r.returncode, r.stderr[-300:])}
        except Exception as e:
            res[wl] = {"error": repr(e)}
    return res


def _small_config(model, n_chains, sweeps, N=30, T=2000):
    import pgopt_b200 as pg
    prob = build_problem(model, T, 0)
    eng = pg.ParticleGibbsEngine(prob["basis"], 1, N, T, n_chains=n_chains)
    try:
        eng.set_data(prob["u"], prob["y"])
        eng.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        eng.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        eng.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        for c in range(n_chains):
            eng.set_state(prob["A"], prob["Q"], prob["x_prim"], 2, chain=c)
        eng.set_rng_philox(7, 0)
        eng.sweep(2)
        eng.synchronize()
        t0 = time.perf_counter()
        eng.sweep(sweeps)
        eng.synchronize()
        return n_chains * sweeps / (time.perf_counter() - t0)
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------- sharded workloads
def _c5_problem(K, H=41, N=30, seed=7):
    import pgopt_b200 as pg
    basis = pg.HilbertGPBasis([1, 5, 5, 5, 5], [4.0, 5.0, 6.0, 7.0, 8.0])  # n_x = 4, n_u = 1, n_phi = 5^4 = 625
    n_x, n_phi = 4, basis.n_phi
    rng = np.random.default_rng(seed)
    A = rng.normal(0, 0.05, (K, n_x, n_phi))
    B = rng.normal(0, 0.1, (K, n_x, n_x))
    Q = B @ np.transpose(B, (0, 2, 1)) + 0.05 * np.eye(n_x)
    w = rng.uniform(size=(K, N))
    w /= w.sum(1, keepdims=True)
    x = rng.normal(0, 1, (K, n_x, N))
    um = rng.normal(0, 1, (K, 1))
    Cm = np.zeros((1, n_x))
    Cm[0, 0] = 1.0
    return dict(basis=basis, batch=pg.PGSampleBatch(A, Q, w, x, um), Cm=Cm, u_init=np.sin(np.arange(H))[None, :], H=H, N=N,
                n_x=n_x, n_phi=n_phi)


def run_c5(args):
    """BASELINE.json configs[4]: K = 10^5 posterior samples x H = 41, n_x = 4, n_phi = 625, scenario-sharded over the
    GPUs of the node through pgb_multi_* (ONE process, in-process NCCL): the samples are resident where they live
    (uploaded once, untimed — in the pipeline they come from the chains that ran there), the timed call is the whole
    pgb_multi_rollout: rollout kernels on every device + NCCL all-gather of x_vec_0 / X / Y + D2H into pinned host
    arrays.  Strong scaling (K fixed)."""
    import pgopt_b200 as pg
    if int(os.environ.get("RANK", 0)) != 0:
        return
    K, n_dev = args.scenarios, args.gpus
    p = _c5_problem(K)
    H, n_x = p["H"], p["n_x"]
    g = pg.LinearMeasurement(p["Cm"])
    out = dict(x0=pg.pinned_empty((K, n_x)), X=pg.pinned_empty((K, H + 1, n_x)), Y=pg.pinned_empty((K, H, 1)))
    with pg.MultiGPUEngine(p["basis"], 1, p["N"], 2, n_dev, list(range(n_dev))) as m:
        t0 = time.perf_counter()
        m.set_samples(p["batch"], roll_seed=1)
        t_up = time.perf_counter() - t0
        calls = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            m.rollout(g, 0.1, H, u_init=p["u_init"], seed=1, out=out)
            if i >= args.warmup:
                calls.append(time.perf_counter() - t0)
        ms = m.last_ms()
        # one stacked-dynamics step + Jacobians on the same sharded samples (pgb_multi_dynamics: no collective)
        dyn = {}
        try:
            xs = np.ascontiguousarray(out["X"][:, 5].reshape(-1))
            for rep in range(3):
                t0 = time.perf_counter()
                _, _, _, kms = m.dynamics(xs, p["u_init"][:, 5], t=5, phi_variant=1)
                t_dyn = time.perf_counter() - t0
            dyn = {"kernel_ms_max_over_devices": kms, "host_call_ms": 1e3 * t_dyn,
                   "hbm_GBps_per_gpu": K / n_dev * 8 * (n_x * p["n_phi"] + 2 * n_x + n_x * (n_x + 1)) / (kms * 1e-3) / 1e9}
        except Exception as e:
            dyn = {"error": repr(e)}
    call = float(np.mean(calls))
    bytes_alg = K * 8 * (n_x * p["n_phi"] + n_x * n_x + n_x + 1 + (2 * H + 2) * n_x + 2 * H)
    peak, _ = measured_peaks()
    print(json.dumps({
        "workload": "c5", "metric": "scenario-steps/s, whole host call (kernels on all GPUs + NCCL gather + D2H)",
        "value": K * H / call, "unit": "scenario-steps/s", "n_gpus": n_dev, "scaling": "strong", "steps": args.steps,
        "warmup": args.warmup, "host_call_ms": 1e3 * call, "rollout_kernel_ms_max_over_devices": ms["rollout_kernel_ms_max"],
        "nccl_gather_ms": ms["scenario_gather_ms"], "kernel_scenario_steps_per_s": K * H / (1e-3 * ms["rollout_kernel_ms_max"]),
        "d2h_bytes_per_call": int(sum(v.nbytes for v in out.values())), "sample_upload_s_untimed": t_up,
        "roofline": {"bound": "fp64", "flops_per_scenario_step": 6600.0,
                     "achieved_tflops": 6600.0 * K * (H + 1) / (1e-3 * ms["rollout_kernel_ms_max"]) / 1e12 / n_dev,
                     "hbm_GBps_per_gpu": bytes_alg / n_dev / (1e-3 * ms["rollout_kernel_ms_max"]) / 1e9, "hbm_peak_GBps": peak},
        "config": {"workload": "scenario rollout K=%d x H=%d, n_x=4, n_phi=625 (BASELINE.json configs[4]), sharded over %d GPU(s) "
                               "through pgb_multi_rollout" % (K, H, n_dev)},
        "dynamics_jacobians": dyn,
        "finite": bool(np.isfinite(out["X"]).all())}))


def run_c4(args):
    """BASELINE.json configs[3]: 64 independent PG chains x N = 2^16 (T = 2000, known basis), chain-sharded over the GPUs
    of the node through pgb_multi_* (one process, one host thread per device), K kept samples per chain (K_b = k_d = 0,
    so K sweeps per chain), then x_T drawn on the owning GPU and ONE NCCL all-gather of the packed [A | Q | x_T | u_m1]
    records.  Strong scaling (64 chains fixed).  `value` counts the whole sampler call incl. the gather and the D2H of
    the records; handle creation (4 GB of ancestor history per 8 chains) is reported beside it."""
    import pgopt_b200 as pg
    if int(os.environ.get("RANK", 0)) != 0:
        return
    n_chains, N, T, K, n_dev = args.chains, args.chain_particles, args.T, args.kept, args.gpus
    prob = build_problem("known", T, 0)
    t0 = time.perf_counter()
    m = pg.MultiGPUEngine(prob["basis"], 1, N, T, n_chains, list(range(n_dev)))
    t_create = time.perf_counter() - t0
    try:
        m.set_data(prob["u"], prob["y"])
        m.set_measurement(pg.LinearMeasurement([[1.0, 0.0]]), 0.1)
        m.set_init_dist(pg.MvNormal([2.0, 2.0], 1.0))
        m.set_prior(prob["V"], 100 * np.eye(2), 10.0)
        times = []
        for i in range(2):  # first pass = warm-up (module load, first cooperative launch), second timed
            m.set_state(prob["A"], prob["Q"], prob["x_prim"], 1)
            m.set_rng_philox(7, 0)
            t0 = time.perf_counter()
            rec = m.particle_gibbs(K if i else 1, 0, 0, roll_seed=3)
            times.append(time.perf_counter() - t0)
        ms = m.last_ms()
    finally:
        m.close()
    dt = times[1]
    print(json.dumps({
        "workload": "c4", "metric": "PG sweeps/s over all chains (filter + trace + statistics + MNIW + device-side sample store), "
                                     "incl. the x_T draw, the NCCL gather of the packed records and their D2H",
        "value": n_chains * K / dt, "unit": "PG sweeps/s", "n_gpus": n_dev, "scaling": "strong", "n_chains": n_chains,
        "N": N, "T": T, "K_kept_per_chain": K, "seconds": dt, "particle_steps_per_s": n_chains * K * N * (T - 1) / dt,
        "nccl_gather_ms": ms["sample_gather_ms"], "handle_creation_s": t_create, "warmup_call_s": times[0],
        "records_shape": list(rec.shape), "finite": bool(np.isfinite(rec).all()),
        "config": {"workload": "%d chains x N=%d, T=%d, known basis, K=%d kept samples per chain (BASELINE.json configs[3] with K "
                               "bounded), sharded over %d GPU(s) through pgb_multi_particle_gibbs" % (n_chains, N, T, K, n_dev)}}))


# ------------------------------------------------------------------------------------------- CPU arm
_SINGLE_CORE = {}


def cpu_sample(model, Ns, Ts, threads=None, f32=False):
    """Oracle (C restatement of particle_Gibbs.jl) timed on a bounded sample of the workload.  The reference sampler
    is single-threaded (particle_Gibbs.jl:154-230); the host cores are used the only way the path shards — one
    independent chain per thread (ctypes releases the GIL during the C call) — and the aggregate is reported."""
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    threads = threads or max(1, min(os.cpu_count() or 1, 256))
    prob = build_problem(model, Ts, seed=0)
    b = prob["basis"]

    def one(chain):
        om = O.known_model() if model == "known" else O.hilbert_model(2, 1, b.count, b.stride, b.L)
        st = O.make_streams(philox=True, seed=1234, chain=chain)
        O.sweep(om, Ns, Ts, prob["u"], prob["y"], [[1.0, 0.0]], 0.1, [2.0, 2.0], np.eye(2), prob["A"], prob["Q"],
                prob["x_prim"], False, st, 2, hist=False, f32=f32)

    O.lib()
    key = (model, Ns, Ts, f32)
    if key not in _SINGLE_CORE:  # the like-for-like figure (the reference sampler is single-threaded): measured once
        t0 = time.perf_counter()
        one(0)
        _SINGLE_CORE[key] = time.perf_counter() - t0
    t1 = _SINGLE_CORE[key]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(threads)))
    dt = time.perf_counter() - t0
    return {"value": threads * Ns * (Ts - 1) / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "single_core_value": Ns * (Ts - 1) / t1,
            "sample": "%d independent conditional sweeps of the same model at the workload's N=%d with T reduced to %d, one per "
                      "host thread (%.1f s wall; one sweep alone on one thread: %.1f s); the reference sampler itself is "
                      "single-threaded (particle_Gibbs.jl:154-230), so single_core_value is the like-for-like figure and "
                      "value the best the box's cores can do with independent chains" % (threads, Ns, Ts, dt, t1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        # warm-up passes (page faults of the history arrays, thread pool) run 2 time steps; the timed ones the stated sample
        r = cpu_sample(args.model, args.cpu_particles, args.cpu_T if i >= args.warmup else 3, f32=(args.dtype == "f32"))
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    sec = vals[-1]["cores"] * args.cpu_particles * (args.cpu_T - 1) / v
    print(json.dumps({
        "impl": "reference", "metric": "CPF-AS particle-steps/sec", "value": v, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "CPF-AS conditional sweep (k>1), %s-basis model, N=%d particles, T=%d (bounded sample of the "
                               "N=%d, T=%d workload: same N, fewer time steps; the per-particle-step cost does not depend on T)"
                               % (args.model, args.cpu_particles, args.cpu_T, args.particles, args.T),
                   "N": args.cpu_particles, "T": args.cpu_T, "same_N_as_gpu_arm": args.cpu_particles == args.particles},
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": vals[-1]["cores"], "kind": "port",
                         "single_core_value": vals[-1]["single_core_value"], "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Julia/MATLAB are not installable in this image: the reference arm is the C oracle restating "
                "Julia/src/particle_Gibbs.jl (oracle/pg_oracle.c)"}))


# single-GPU side benchmarks kept as scripts, reachable from the one entry point
SIDE_WORKLOADS = {"small": "bench_small_configs.py", "mid": "bench_mid_n.py", "dynamics": "bench_dynamics.py",
                  "mniw": "bench_mniw.py", "screening": "bench_screening.py"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="generic", choices=["generic", "known"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f64: the reference's arithmetic (default, the headline); f32: the F32 precision mode "
                         "(FP32 particle state / basis / noise / log-weights, FP64 normalisation and cumulative sums)")
    ap.add_argument("--particles", type=int, default=2 ** 20)
    ap.add_argument("--T", type=int, default=2000)
    ap.add_argument("--cpu-particles", type=int, default=2 ** 20,
                    help="particles of the CPU sample: the workload's own N = 2^20 by default (the oracle's rate per particle-step is "
                         "flat in N: README, profiles/r02_cpu_oracle_rate_vs_N.json)")
    ap.add_argument("--cpu-T", type=int, default=6, help="time steps of the CPU sample (reduced from T = 2000: stated in the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chain-base", type=int, default=0, help="Philox stream id of rank 0's chain (rank r uses chain_base + r)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bounded rollout / chain-batch side measurements")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "c4", "c5"] + sorted(SIDE_WORKLOADS),
                    help="sweep: the headline CPF-AS sweep (default; the driver's contract).  c4 / c5: BASELINE.json configs[3] / [4] "
                         "sharded over --gpus devices from ONE process through pgb_multi_* (run with plain python).  "
                         "small / mid / dynamics / mniw / screening: the single-GPU side benchmarks (scripts/bench_*.py)")
    ap.add_argument("--scenarios", type=int, default=100000)
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--chain-particles", type=int, default=2 ** 16)
    ap.add_argument("--kept", type=int, default=8)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload in SIDE_WORKLOADS:
        import runpy
        sys.argv = [SIDE_WORKLOADS[args.workload]]
        runpy.run_path(os.path.join(ROOT, "scripts", SIDE_WORKLOADS[args.workload]), run_name="__main__")
    elif args.workload == "c5":
        run_c5(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
