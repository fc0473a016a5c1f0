#!/usr/bin/env python
"""Benchmark of the B200-native SymBoltz hot path (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): ΛCDM (lmax = 10, nx = 4 => 82 unknowns per mode),
Planck18-like synthetic parameter set, CMB TT/EE/TE C_l at the 129 multipoles l in {2,3,5,10,20:20:2500}, solving every one of the
≈2020 fine k-modes (step π/τ0 on [1e-2, 2e3] H0/c) directly, 300 line-of-sight times.
A "step" = one pass of the hot path over one cosmology: perturbation solve of all modes (dense output at the 300 times)
-> source functions -> line-of-sight integration -> C_l.   metric = k-modes/s (whole job, all GPUs).

  python bench.py [--gpus N --steps K --warmup W]        our arm (N > 1: launched with torch.distributed.run, one rank per GPU)
  python bench.py --impl reference [...]                 CPU arm: the oracle port of the reference path on the host cores
                                                         (the reference itself is Julia, which this image does not have)
Weak scaling: every rank processes its own synthetic cosmology per step; no data-path collective.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
NCU_DRAM_BYTES_PER_LAUNCH = 18081792 + 345558528  # ncu --set full capture of the statically scheduled sb_integrate_kernel launch at the bench size (profiles/integrate_r1.md, r1o)
FP64_PEAK_TFLOPS = 33.84  # measured on this pool's B200 with scripts/fp64_peak.cu (profiles/fp64_peak_r1.txt); MEASURED_PEAKS.json has no FP64 figure


def synthetic_pars(sb, M, rank, step=0):
    """Deterministic synthetic parameter set around Planck18 (one per rank: data-parallel sweep of cosmologies)."""
    p = sb.parameters_Planck18(M)
    rng = np.random.default_rng(1000 + rank)
    f = 1 + 0.02 * (rng.random(3) - 0.5) if rank > 0 else np.ones(3)
    p["Omega_c"] *= f[0]
    p["Omega_b"] *= f[1]
    p["ns"] *= f[2]
    return p


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.splitlines()[0].split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def step_flops(prob, stats, nsave):
    """Algorithmic FP64 flops of the integrator launch from its own counters (per-function counts from the generator)."""
    N = prob.N
    nacc, nrej, nf, nsolve = (float(stats[:, i].sum()) for i in range(4))
    natt = nacc + nrej
    F_f, F_lu, F_solve = prob.flops["f"], prob.flops["lu"], prob.flops["solve"]
    comb = 2 * N * 49      # 21 a-terms + 28 C-terms per component
    wood = 8 * N           # two hub dot products + rank-2 update per solve
    return nf * F_f + nacc * 2 * F_f + natt * (F_lu + comb + 6 * N + 4 * 4 * N) + nsolve * (F_solve + wood) + nsave * 8 * N


def oracle_sample(pars_oracle, bg_knots, ks_sample, taus, ls, nthreads=0):
    """CPU port of the same per-step work on a bounded sample of the k-modes: perturbation solve + sources (+ LOS of the sample)."""
    from oracle import sbref
    obg = sbref.Background.from_knots(pars_oracle, *bg_knots)
    t0 = time.time()
    sol = sbref.solvept(obg, ks_sample, saveat=taus, nthreads=nthreads)
    S = sbref.sources(obg, ks_sample, taus, sol["usave"])
    return time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--schedule", default="static", choices=["static", "queue"], help="integrator work distribution: static preemptive schedule from a learnt cost model, or the atomic queue")
    ap.add_argument("--cpu-sample", type=int, default=0, help="number of k-modes in the CPU sample (0 = auto)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
    config = {"workload": "LCDM lmax=10 nx=4 (82 unknowns/mode): CMB TT/EE/TE C_l, 129 l <= 2500, all ~2020 fine k-modes solved directly (step pi/tau0 on [1e-2,2e3] H0/c), 300 LOS times, Rodas5P reltol=abstol=1e-5",
              "cosmologies_per_step_per_gpu": 1, "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"one cosmology per GPU x {max(world, args.gpus)}"}

    if args.impl == "reference":
        # CPU arm.  The reference is Julia (not installable here: no julia binary, no network); the arm runs the oracle port of the
        # same path (oracle/sbref.cpp, OpenMP over modes like Threads.@spawn per mode, src/solve.jl:566) on all host cores.
        if rank != 0:
            return
        from oracle import sbref
        import symboltz.jl_b200 as sb
        M = sb.ΛCDM(lmax=10)
        pars = synthetic_pars(sb, M, 0)
        prob = sb.CosmologyProblem(M, pars)
        bg = sb.solvebg(prob)
        ks_fine, taus = sb.cmb_grids(bg)
        nsamp = args.cpu_sample or 128
        idx = np.linspace(0, len(ks_fine) - 1, nsamp).round().astype(int)  # spans the whole k-range, so cost per mode is representative
        op = sbref.planck18(lmax=10, Omega_c=pars["Omega_c"], Omega_b=pars["Omega_b"], ns=pars["ns"])
        knots = (bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0)
        cores = os.cpu_count()
        for _ in range(args.warmup):
            oracle_sample(op, knots, ks_fine[idx[:8]], taus, ls)
        times = [oracle_sample(op, knots, ks_fine[idx], taus, ls) for _ in range(args.steps)]
        t = float(np.mean(times))
        v = nsamp / t
        print(json.dumps({"impl": "reference", "metric": "k-modes/s (LCDM perturbations -> C_l TT/EE/TE, l<=2500)", "value": v, "unit": "k-modes/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "k-modes/s", "cores": cores, "kind": "port", "sample": f"{nsamp} of {len(ks_fine)} k-modes evenly spaced over the k-range, per step; perturbation solve + sources on the host (OpenMP, all cores)"},
                          "e2e": {"value": v, "unit": "k-modes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "reference is Julia (unavailable in this image); this arm times the CPU oracle port of the same path"}))
        return

    import torch
    import symboltz.jl_b200 as sb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path is GPU-only)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    M = sb.ΛCDM(lmax=10)
    pars = synthetic_pars(sb, M, rank)
    prob = sb.CosmologyProblem(M, pars)
    bg = sb.solvebg(prob)  # host background (outside the timed region, "precomputed on the host exactly as the reference does")
    jl = sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.001)
    plan = sb.CMBPlan(prob, bg, jl, modes=("TT", "EE", "TE"), direct=True)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    plan.upload()
    for i in range(args.warmup):
        plan.run()
        if i == 0 and args.schedule == "static":
            # plan-time artefact, like an FFT plan: the first (queue-scheduled) solve yields the step count of every mode; a smooth
            # 48-knot fit of attempts(k) -- not the per-mode counts -- feeds the static preemptive schedule of all later launches
            plan.learn_schedule()
        plan.run_e2e()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)  # let the first nvidia-smi query get in flight; all samples are taken while the timed steps run
    # ---- device-resident timing: CUDA events per step on the launching stream, L2 flushed between steps
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(args.steps)]
    barrier()
    for a, b, c, d in ev:
        flush.fill_(1.0)
        a.record()
        plan.solve()
        b.record()
        plan.sources()
        d.record()
        plan.los_cl()
        c.record()
    barrier()
    t_step = np.array([a.elapsed_time(c) for a, b, c, d in ev]) * 1e-3
    t_kernel = np.array([a.elapsed_time(b) for a, b, c, d in ev]) * 1e-3
    t_src = np.array([b.elapsed_time(d) for a, b, c, d in ev]) * 1e-3
    t_los = np.array([d.elapsed_time(c) for a, b, c, d in ev]) * 1e-3
    stats = plan.d_stats.cpu().numpy()
    ok = bool((plan.d_ret.cpu().numpy() == 0).all())
    # ---- end-to-end timing through the public plan API: pinned host -> device, all kernels, C_l back on the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0)
        Cl = plan.run_e2e()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / args.steps
    # subtract nothing: the flush is inside (it costs ~0.1 ms of a ~100 ms step)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=10)
    total = torch.tensor([t_step.sum(), t_e2e * args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    T_res, T_e2e = float(total[0]), float(total[1])
    nmodes = plan.nk
    value = world * nmodes * args.steps / T_res
    e2e = world * nmodes * args.steps / T_e2e
    if rank == 0:
        flops = step_flops(prob, stats, plan.nk * plan.nt)
        src_bytes = float(plan.nk * plan.nt * (prob.N + 2) * 8)
        hbm_peak, hbm_src = 6452.8, "fallback: SURVEY 8d figure"
        try:
            mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            cand = [v for k, v in mp_.items() if isinstance(v, (int, float)) and "hbm" in k.lower() and "sustain" in k.lower()] or \
                   [v for k, v in mp_.items() if isinstance(v, (int, float)) and "hbm" in k.lower()]
            if cand:
                hbm_peak, hbm_src = float(cand[0]), "MEASURED_PEAKS.json"
        except Exception:
            pass
        achieved = flops / float(t_kernel.mean()) / 1e12
        out = {"metric": "k-modes/s (LCDM perturbations -> C_l TT/EE/TE, l<=2500)", "value": value, "unit": "k-modes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * T_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": dict(config, modes_per_step_per_gpu=nmodes, success=ok,
                              schedule=("static preemptive lists over %d resident warps, cost model attempts(k) learnt from the first warm-up solve" % plan.nlists) if plan.d_items is not None else "atomic queue, descending k"),
               "cl_wall_time_ms": 1e3 * T_e2e / args.steps,
               "e2e": {"value": e2e, "unit": "k-modes/s", "h2d_bytes_per_step": plan.h2d_bytes, "d2h_bytes_per_step": plan.d2h_bytes},
               "gpu_launches": plan.launches_resident * args.steps,
               "roofline": {"kernel": "sb_integrate_kernel", "bound": "fp64", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this launch (profiles/integrate_r1.md, r1o): 18.1 MB read + 345.6 MB written = the saved states usave[2019][300][82] f64 (397 MB algorithmic)",
                            "peak_source": "measured DFMA peak, scripts/fp64_peak.cu (FP64 kernel: neither HBM nor tensor bound; MEASURED_PEAKS.json has no FP64 entry)",
                            "kernel_ms": 1e3 * float(t_kernel.mean()), "algorithmic_flops_per_launch": flops,
                            "attempted_steps_per_launch": float(stats[:, 0].sum() + stats[:, 1].sum())},
               # the S(tau,k) stage (north star: achieved HBM GB/s of the source traffic): sb_srcbg_kernel + sb_source_kernel read the saved
               # states usave[nk][nt][N] once and write S[nk][2][nt]; algorithmic bytes per step over the event time of the two launches
               "roofline_sources": {"kernel": "sb_srcbg_kernel + sb_source_kernel", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
                                    "algorithmic_bytes_per_step": src_bytes, "kernel_ms": 1e3 * float(t_src.mean()), "achieved": src_bytes / float(t_src.mean()) / 1e9,
                                    "frac": src_bytes / float(t_src.mean()) / 1e9 / hbm_peak},
               "los_cl_ms": 1e3 * float(t_los.mean()),
               "clocks": sampler.summary()}
        # CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
        if world == 1:
            try:
                from oracle import sbref
                nsamp = args.cpu_sample or 64
                idx = np.linspace(0, len(plan.ks_fine) - 1, nsamp).round().astype(int)
                op = sbref.planck18(lmax=10, Omega_c=pars["Omega_c"], Omega_b=pars["Omega_b"], ns=pars["ns"])
                tc = oracle_sample(op, (bg.t, bg.y, bg.dy, bg.tau0, bg.kappa0), plan.ks_fine[idx], plan.taus, ls)
                out["cpu_baseline"] = {"value": nsamp / tc, "unit": "k-modes/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"{nsamp} of {len(plan.ks_fine)} k-modes evenly spaced over the k-range: perturbation solve + sources with the CPU oracle (OpenMP over modes; zero-skipping LU in a fill-reducing order, compressed Jacobian probing)"}
            except Exception as e:  # the oracle is optional at bench time
                out["cpu_baseline"] = {"value": None, "unit": "k-modes/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
