#!/usr/bin/env python
"""Benchmark of the B200-native SymBoltz hot path (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): ΛCDM (lmax = 10, nx = 4 => 82 unknowns per mode),
Planck18-like synthetic parameter set, CMB TT/EE/TE C_l at the 129 multipoles l in {2,3,5,10,20:20:2500}, solving every one of the
≈2020 fine k-modes (step π/τ0 on [1e-2, 2e3] H0/c) directly, 300 line-of-sight times.
A "step" = one pass of the hot path over one cosmology per GPU: perturbation solve of all modes with the source functions
S(τ,k) formed at the 300 save times -> line-of-sight integration -> C_l [-> NCCL all-gather of the C_l of all ranks when N > 1].
metric = k-modes/s (whole job, all GPUs).

  python bench.py [--gpus N --steps K --warmup W]        our arm (N > 1: launched with torch.distributed.run, one rank per GPU)
  python bench.py --impl reference [...]                 CPU arm: the oracle port of the reference path on the host cores
                                                         (the reference itself is Julia, which this image does not have)
Weak scaling: every rank processes its own synthetic cosmology per step (cosmologies are the independent units); the only
exchange of that path is the gather of the results, which is inside the timed region.  The north-star strong-scaling paths are timed
beside it and reported as extra keys of the same line: `strong_single_cosmology` (ONE cosmology, modes strided over the N ranks,
NCCL all-reduce of the sources and of the partial C_l sums) and `config4_sweep` (BASELINE configs[3]: 4096 w0waCDM cosmologies ×
256 modes sharded over the N ranks, NCCL gather of P(k)); `config1_pk` and `cl_default_chebyshev` are the reference's own two
headline workloads (paper: 0.3 s and 3.1 s on a laptop) on one GPU.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    # The CPU arm uses every host core whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1, which made the
    # round-1 reference arm 30x slower at N > 1).  Must happen before numpy / the OpenMP runtime are loaded.
    _nc = str(len(os.sched_getaffinity(0)))
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = _nc

import numpy as np  # noqa: E402

LS = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
MODES = ("TT", "EE", "TE")
METRIC = "k-modes/s (LCDM perturbations -> C_l TT/EE/TE, l<=2500)"


def synthetic_overrides(rank):
    """Deterministic synthetic parameter set around Planck18 (one per rank: data-parallel sweep of cosmologies).  Plain numbers, so
    that the CPU arm does not need the product package."""
    h = 0.6736
    rng = np.random.default_rng(1000 + rank)
    f = 1 + 0.02 * (rng.random(3) - 0.5) if rank > 0 else np.ones(3)
    return dict(Omega_c=0.1200 / h**2 * f[0], Omega_b=0.0224 / h**2 * f[1], ns=0.965 * f[2])


def make_config(ngpu):
    return {"workload": "LCDM lmax=10 nx=4 (82 unknowns/mode): CMB TT/EE/TE C_l, 129 l <= 2500, all ~2020 fine k-modes solved directly (step pi/tau0 on [1e-2,2e3] H0/c), 300 LOS times, Rodas5P reltol=abstol=1e-5",
            "cosmologies_per_step_per_gpu": 1, "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"one cosmology per GPU x {ngpu}; C_l of all ranks gathered (NCCL) inside the step"}


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


def step_flops(prob, stats, nsave, src_flops=0):
    """Algorithmic FP64 flops of the integrator launch from its own counters (per-function counts from the generator)."""
    N = prob.N
    nacc, nrej, nf, nsolve = (float(stats[:, i].sum()) for i in range(4))
    natt = nacc + nrej
    F_f, F_lu, F_solve = prob.flops["f"], prob.flops["lu"], prob.flops["solve"]
    comb = 2 * N * 49      # 21 a-terms + 28 C-terms per component
    wood = 8 * N           # two hub dot products + rank-2 update per solve
    return nf * F_f + nacc * 2 * F_f + natt * (F_lu + comb + 6 * N + 4 * 4 * N) + nsolve * (F_solve + wood) + nsave * (8 * N + src_flops)


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
class ReferenceArm:
    """The reference's CPU path restated (oracle/sbref.cpp + sbref.py): background solve, Rodas5P + sparse LU per mode (OpenMP over modes
    like Threads.@spawn per mode, src/solve.jl:566), sources, line-of-sight integration and C_l -- nothing of the product is imported."""

    def __init__(self, over, nsamp):
        from oracle import sbref
        self.sbref = sbref
        self.cores = len(os.sched_getaffinity(0))
        self.obg = sbref.Background(sbref.planck18(lmax=10, **over))
        self.ks_fine, self.taus = sbref.cmb_grids(self.obg)
        self.nsamp = nsamp
        self.idx = np.linspace(0, len(self.ks_fine) - 1, nsamp).round().astype(int)  # spans the whole k-range, so cost per mode is representative
        self.ojl = sbref.SphericalBesselCache(LS, xcut=2e3 * self.obg.tau0 * 1.001)  # the j_l cache is an input of spectrum_cmb in the reference too (built once)

    def step(self, idx=None):
        idx = self.idx if idx is None else idx
        t0 = time.perf_counter()
        self.sbref.spectrum_cmb(list(MODES), self.obg, self.ojl, ks=self.ks_fine[idx], nthreads=self.cores)
        return time.perf_counter() - t0

    def describe(self):
        return (f"{self.nsamp} of {len(self.ks_fine)} k-modes evenly spaced over the k-range per step: perturbation solve with dense output at 300 times, "
                f"sources, line-of-sight integration at 129 l and C_l of the sample, CPU oracle port on {self.cores} OpenMP threads")


def run_reference(args):
    """`--impl reference`: the reference is Julia (not installable here: no julia binary, no network); this arm times the CPU oracle port
    of the same path on all host cores, on a bounded sample of the same workload per step.  Rank 0 only under torchrun."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    ngpu = max(int(os.environ.get("WORLD_SIZE", 1)), args.gpus)
    arm = ReferenceArm(synthetic_overrides(0), args.cpu_sample or 128)
    for _ in range(args.warmup):
        arm.step(arm.idx[::8])
    times = [arm.step() for _ in range(args.steps)]
    t = float(np.mean(times))
    v = arm.nsamp / t
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "k-modes/s", "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(ngpu),
                      "cpu_baseline": {"value": v, "unit": "k-modes/s", "cores": arm.cores, "kind": "port", "sample": arm.describe()},
                      "e2e": {"value": v, "unit": "k-modes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "omp_threads": arm.cores,
                      "note": "reference is Julia (unavailable in this image); this arm times the CPU oracle port of the same path (solve + sources + LOS + C_l) and uses no GPU"}))


# ------------------------------------------------------------------------------------------------ extras (our arm)
def config4_thetas(n):
    """Latin hypercube, numpy default_rng(0)."""
    lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3])
    hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
    rng = np.random.default_rng(0)
    u = (rng.permuted(np.tile(np.arange(n), (7, 1)), axis=1).T + rng.random((n, 7))) / n
    th = lo + (hi - lo) * u
    th[:, 1] /= th[:, 0] ** 2
    th[:, 2] /= th[:, 0] ** 2
    return ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"], th


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--schedule", default="static", choices=["static", "queue"], help="integrator work distribution: static preemptive schedule from a cost model learnt on ANOTHER cosmology, or the atomic queue")
    ap.add_argument("--cpu-sample", type=int, default=0, help="number of k-modes in the CPU sample (0 = auto)")
    ap.add_argument("--config4", type=int, default=4096, help="cosmologies of the config-4 sweep extra (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="headline only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    config = make_config(max(world, args.gpus))

    import ctypes as C
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

    def problem(r):
        pars = sb.parameters_Planck18(M)
        pars.update(synthetic_overrides(r))
        return sb.CosmologyProblem(M, pars)

    prob = problem(rank)
    bg = sb.solvebg(prob)  # host background (outside the timed region, "precomputed on the host exactly as the reference does")
    jl = sb.SphericalBesselCache(LS, xcut=2e3 * bg.tau0 * 1.02)
    plan = sb.CMBPlan(prob, bg, jl, modes=MODES, direct=True)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    gathered = torch.empty((world,) + tuple(plan.d_Cl.shape), dtype=torch.float64, device="cuda") if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_cl():
        if dist is not None:
            dist.all_gather_into_tensor(gathered, plan.d_Cl)  # the sweep's one exchange: every rank ends with the C_l of all cosmologies

    # plan-time artefact, like an FFT plan: attempts(k) learnt from the step counters of ANOTHER cosmology (a 48-knot fit, not per-mode
    # counts) feeds the static preemptive schedule of the timed cosmology
    model = None
    if args.schedule == "static":
        prob_c = problem(rank + 7919)
        bg_c = sb.solvebg(prob_c)
        ks_c, _ = sb.cmb_grids(bg_c)
        st = sb.solvept(prob_c, bg_c, ks_c).stats
        model = sb.ModeCostModel(ks_c, st[:, 0] + st[:, 1])
    plan.upload()
    for i in range(args.warmup):
        plan.run()
        if i == 0 and model is not None:
            plan.learn_schedule(model)
        plan.run_e2e()
        gather_cl()
    peak = C.c_double(0.0)
    if sb.api.los_lib().sbl_dfma_peak(C.c_int(1 << 14), C.c_int(3), C.byref(peak), C.c_void_p(torch.cuda.current_stream().cuda_stream)) != 0 or not peak.value > 0:
        raise SystemExit("bench.py: FP64 peak measurement failed")
    fp64_peak = float(peak.value)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)  # let the first nvidia-smi query get in flight; all samples are taken while the timed steps run
    # ---- device-resident timing: CUDA events per step on the launching stream, L2 flushed between steps
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    barrier()
    for a, b, c in ev:
        flush.fill_(1.0)
        a.record()
        plan.solve()
        b.record()
        plan.los_cl()
        gather_cl()
        c.record()
    barrier()
    t_step = np.array([a.elapsed_time(c) for a, b, c in ev]) * 1e-3
    t_kernel = np.array([a.elapsed_time(b) for a, b, c in ev]) * 1e-3
    t_los = np.array([b.elapsed_time(c) for a, b, c in ev]) * 1e-3
    stats = plan.d_stats.cpu().numpy()
    ok = bool((plan.d_ret.cpu().numpy() == 0).all())
    # ---- end-to-end timing through the public plan API: pinned host -> device, all kernels, C_l back on the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0)
        plan.upload()
        plan.run()
        gather_cl()
        Cl = plan.download()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / args.steps
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=10)
    total = torch.tensor([t_step.sum(), t_e2e * args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    T_res, T_e2e = float(total[0]), float(total[1])
    nmodes = plan.nk

    # ---- the other work distribution, for the record (3 steps)
    other = {}
    try:
        saved = plan.d_items
        if saved is not None:
            plan.d_items = None
        else:
            plan.learn_schedule(model) if model is not None else plan.learn_schedule()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan.run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        other = {"schedule": "atomic queue, descending k" if saved is not None else "static", "ms_per_step": e0.elapsed_time(e1) / 3}
        plan.d_items = saved
    except Exception as e:
        other = {"error": repr(e)}

    extras = {}
    if not args.no_extras:
        extras = run_extras(sb, torch, dist, args, rank, world, problem, prob, bg, jl, model)

    if rank == 0:
        value = world * nmodes * args.steps / T_res
        e2e = world * nmodes * args.steps / T_e2e
        src_flops = plan.src_flops if plan.fused else 0
        flops = step_flops(prob, stats, plan.nk * plan.nt, src_flops)
        hbm_peak, hbm_src = 6452.8, "fallback: SURVEY 8d figure"
        try:
            mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            if isinstance(mp_.get("hbm_gbs"), (int, float)):
                hbm_peak, hbm_src = float(mp_["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        traffic, traffic_src = None, "no capture on file"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "integrate_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        achieved = flops / float(t_kernel.mean()) / 1e12
        s_bytes = float(plan.nk * plan.nt * 2 * 8)
        out = {"metric": METRIC, "value": value, "unit": "k-modes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * T_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config, "modes_per_step_per_gpu": nmodes, "success": ok,
               "schedule": ("static preemptive lists over %d resident warps, cost model attempts(k) learnt on a DIFFERENT synthetic cosmology" % plan.nlists) if args.schedule == "static" else "atomic queue, descending k",
               "other_schedule": other,
               "cl_wall_time_ms": 1e3 * T_e2e / args.steps,
               "e2e": {"value": e2e, "unit": "k-modes/s", "h2d_bytes_per_step": plan.h2d_bytes, "d2h_bytes_per_step": plan.d2h_bytes},
               "gpu_launches": (plan.launches_resident + (1 if world > 1 else 0)) * args.steps,
               "roofline": {"kernel": "sb_integrate_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic,
                            "traffic_source": traffic_src,
                            "peak_source": "DFMA peak measured in this run (sbl_dfma_peak: 8 independent FMA chains per thread, 8 x 256 threads per SM, best of 3) -- FP64 kernel: neither HBM nor tensor bound; MEASURED_PEAKS.json has no FP64 entry",
                            "kernel_ms": 1e3 * float(t_kernel.mean()), "algorithmic_flops_per_launch": flops,
                            "attempted_steps_per_launch": float(stats[:, 0].sum() + stats[:, 1].sum()),
                            "us_per_attempt_per_warp": 1e6 * float(t_kernel.mean()) * plan.nlists / float(stats[:, 0].sum() + stats[:, 1].sum()) if plan.d_items is not None else None},
               # S(tau,k): formed inside the integrator at the save times and written once (north star "achieved HBM GB/s for the S(k,tau) traffic"):
               # the only HBM bytes of the stage are the stores of S[nk][2][nt]; they are spread over the whole integrator launch, so the achieved
               # bandwidth is a tiny fraction of the peak by construction -- the stage is no longer a separate HBM-bound kernel
               "roofline_sources": {"kernel": "sb_integrate_kernel (fused source evaluation at the save times)" if plan.fused else "sb_srcbg_kernel + sb_source_kernel", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
                                    "algorithmic_bytes_per_step": s_bytes, "kernel_ms": 1e3 * float(t_kernel.mean()), "achieved": s_bytes / float(t_kernel.mean()) / 1e9,
                                    "frac": s_bytes / float(t_kernel.mean()) / 1e9 / hbm_peak},
               "los_cl_ms": 1e3 * float(t_los.mean()),
               "clocks": sampler.summary()}
        out.update(extras)
        # CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
        if world == 1:
            try:
                arm = ReferenceArm(synthetic_overrides(0), args.cpu_sample or 64)
                arm.step(arm.idx[::8])
                tc = arm.step()
                out["cpu_baseline"] = {"value": arm.nsamp / tc, "unit": "k-modes/s", "cores": arm.cores, "kind": "port", "sample": arm.describe()}
            except Exception as e:  # the oracle is optional at bench time
                out["cpu_baseline"] = {"value": None, "unit": "k-modes/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(sb, torch, dist, args, rank, world, problem, prob, bg, jl, model=None):
    """Workloads timed beside the headline (same process, after it): each guarded, a failure is reported as text, never fatal."""
    import warnings
    ex = {}

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxtime(t):
        if dist is None:
            return t
        x = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x[0])

    # (1) BASELINE configs[0]: P(k) at 100 log-spaced k (the paper's 0.3 s workload), through the public call, host in -> host out
    try:
        ks = sb.loggrid(1e-4, 1.0, length=100) / sb.k0
        sb.spectrum_matter(prob, ks, bgsol=bg)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            P, sol = sb.spectrum_matter(prob, ks, bgsol=bg, return_solution=True)
            ts.append(time.perf_counter() - t0)
        st = sol.stats
        ex["config1_pk"] = {"workload": "P(k, z=0), 100 log-spaced k in 1e-4..1 h/Mpc, spectrum_matter(prob, ks) from host arrays to host P(k), background precomputed", "ms": 1e3 * float(np.median(ts)),
                            "k_modes_per_s": 100 / float(np.median(ts)), "max_attempts_of_a_mode": int((st[:, 0] + st[:, 1]).max()), "success": bool(sol.success)}
    except Exception as e:
        ex["config1_pk"] = {"error": repr(e)}
    # (1b) the same workload with the reference's lower-accuracy integrators (ptalg(prob; accuracy = 1 / 0), src/solve.jl:333-337), each with its P(k) error against a
    # converged solve beside the time: what the accuracy knob buys on this engine (nothing: Rodas5P on the split mapping is the fastest AND the most accurate of the three)
    try:
        Pt = sb.spectrum_matter(prob, ks, bgsol=bg, reltol=1e-10, abstol=1e-10)
        alt = {}
        for alg, tol in (("Rodas5P", 1e-5), ("KenCarp4", 1e-5), ("TRBDF2", 1e-5)):  # (the accuracy knob changes the algorithm, not the default tolerances)
            sb.spectrum_matter(prob, ks, bgsol=bg, alg=alg, reltol=tol, abstol=tol)
            torch.cuda.synchronize()
            tt = []
            for _ in range(3):
                t0 = time.perf_counter()
                Pa = sb.spectrum_matter(prob, ks, bgsol=bg, alg=alg, reltol=tol, abstol=tol)
                tt.append(time.perf_counter() - t0)
            alt[f"{alg}, reltol = abstol = {tol:g}"] = {"ms": 1e3 * float(np.median(tt)), "max_rel_error_of_Pk_vs_converged": float(np.abs(Pa / Pt - 1).max())}
        ex["config1_pk_integrators"] = alt
    except Exception as e:
        ex["config1_pk_integrators"] = {"error": repr(e)}
    # (2) the reference's default C_l path: 61 Chebyshev nodes + barycentric interpolation (the paper's 3.1 s workload)
    try:
        planc = sb.CMBPlan(prob, bg, jl, modes=MODES, direct=False)
        planc.run_e2e()
        planc.run_e2e()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            planc.run_e2e()
            ts.append(time.perf_counter() - t0)
        st = planc.d_stats.cpu().numpy()
        ex["cl_default_chebyshev"] = {"workload": "C_l TT/EE/TE at 129 l, 61 Chebyshev k-nodes + barycentric interpolation to the fine grid (reference default, angular.jl:267-273), host knots in -> host C_l out",
                                      "ms": 1e3 * float(np.median(ts)), "max_attempts_of_a_mode": int((st[:, 0] + st[:, 1]).max())}
        del planc
    except Exception as e:
        ex["cl_default_chebyshev"] = {"error": repr(e)}
    # (3) north-star item 4, single cosmology: modes strided over the ranks, all-reduce of S and of the partial C_l sums
    if dist is not None:
        try:
            prob0 = problem(0)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                bg0 = sb.solvebg(prob0)
                jl0 = jl if rank == 0 else sb.SphericalBesselCache(LS, xcut=2e3 * bg0.tau0 * 1.02)
                # work distribution inside a rank (chosen by solvept): with no more modes than the split mapping holds, one CTA per mode; above that the
                # atomic queue over the resident warps (with ~1000 modes per rank every mode has a warp of its own: a static schedule has nothing to
                # balance and measured 47.0 instead of 44.8 ms at N = 2, profiles/scaling_r2.md)
                nk_rank = -(-len(sb.cmb_grids(bg0)[0]) // world)
                sched = {}
                sb.spectrum_cmb(list(MODES), prob0, jl0, bgsol=bg0, direct=True, **sched)
                ts = []
                for _ in range(3):
                    sync()
                    t0 = time.perf_counter()
                    sb.spectrum_cmb(list(MODES), prob0, jl0, bgsol=bg0, direct=True, **sched)
                    torch.cuda.synchronize()
                    ts.append(maxtime(time.perf_counter() - t0))
            ex["strong_single_cosmology"] = {"workload": "the headline cosmology's C_l with its ~2020 modes strided over the ranks: NCCL all-reduce (disjoint supports = all-gather) of S[nk][2][300] (9.7 MB), LOS on contiguous fine-k slices, NCCL all-reduce of the partial C_l sums [3][129]",
                                             "ms": 1e3 * float(np.median(ts)), "ranks": world, "modes_per_rank": nk_rank, "work_distribution": ("one CTA per mode (split kernel)" if sb.split_pays(prob0, sb.cmb_grids(bg0)[0][rank::world]) else "atomic queue, one warp per mode"),
                                             "limiter": "latency of the slowest mode (its sequential Rosenbrock attempts), not the collectives"}
        except Exception as e:
            ex["strong_single_cosmology"] = {"error": repr(e)}
    # (4) BASELINE configs[3]: w0waCDM sweep sharded by cosmology, NCCL gather of P(k)
    if args.config4 > 0:
        try:
            Mw = sb.w0waCDM(lmax=10)
            probw = sb.CosmologyProblem(Mw, sb.parameters_Planck18(Mw))
            names, th = config4_thetas(args.config4)
            ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sb.spectrum_matter_sweep(probw, names, th[:2 * world], ks)
                sync()
                t0 = time.perf_counter()
                P, info = sb.spectrum_matter_sweep(probw, names, th, ks, chunk=32, return_info=True)
                torch.cuda.synchronize()
                t = maxtime(time.perf_counter() - t0)
            ex["config4_sweep"] = {"workload": f"{args.config4} w0waCDM cosmologies (Latin hypercube, seed 0) x 256 k-modes, cosmologies strided over the ranks, host background solves on {max(1, (os.cpu_count() or 1) // world)} threads per rank, one NCCL all-reduce gather of P(k) [{args.config4}][256] f64",
                                   "wall_s": t, "k_modes_per_s": args.config4 * 256 / t, "ranks": world, "scaling": "strong", "background_failures": info["background_failures"], "mode_failures": info["mode_failures"],
                                   "finite_rows": int(np.isfinite(P).all(axis=1).sum())}
        except Exception as e:
            ex["config4_sweep"] = {"error": repr(e)}
    return ex


if __name__ == "__main__":
    main()
