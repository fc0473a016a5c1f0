"""Host-side mirror of the reference's public interface for the hot path (Julia names kept; reference file:line
relative to /root/reference):

    ΛCDM / w0waCDM, parameters_Planck18        src/models/cosmologies.jl:27-106, 211-215; src/parameters.jl:1-21
    CosmologyProblem, parameter_updater        src/solve.jl:129-236, 272-307
    solve, solvebg, solvept, issuccess         src/solve.jl:380-402, 427-435, 543-569, 604-606
    spectrum_primordial, spectrum_matter       src/observables/fourier.jl:14-30, 74-101
    source_grid, source_kinterp, ChebyshevInterpolator   src/observables/fourier.jl:232-291, 419-467, 524-547
    SphericalBesselCache, los_integrate, spectrum_cmb    src/observables/angular.jl:9-48, 109-185, 198-223, 260-359
    lingrid, loggrid, cosgrid, chebgrid        src/utils.jl:269-292

The Julia toolchain is not available in this image, so the host side is Python over the same C ABI a Julia `ccall`
shim would bind (include/symboltz_b200.h, julia/SymBoltzB200.jl).  PyTorch is used only for device
memory, streams and torch.distributed plumbing.  All perturbation / line-of-sight work runs in the CUDA libraries built
by build.py; if they (or a GPU) are missing the calls raise -- there is no CPU fallback.
"""
import ctypes as C
import math
import time
import warnings

import numpy as np
import torch

from . import build as _build

k0 = 1.0 / 2997.92458  # h/Mpc in units of H0/c (src/constants.jl:19)

# ---------------------------------------------------------------------------------------------- constants (src/constants.jl)
_c, _h, _kB, _GN = 299792458.0, 6.62607015e-34, 1.380649e-23, 6.67430e-11
_hbar = _h / (2 * math.pi)
_Mpc, _eV = 3.0856775814913673e22, 1.602176634e-19
_H100 = 100 * 1e3 / _Mpc
_amu = 1.66053906660e-27
_mH, _mHe = 1.008 * _amu, 4.0026022 * _amu


# ---------------------------------------------------------------------------------------------- grids (src/utils.jl:269-292)
def lingrid(a, b, step=None, length=None):
    if not a < b:
        raise ValueError("Right grid interval must be higher than left")
    if (step is None) == (length is None):
        raise ValueError("Must provide exactly one of step or length")
    if step is not None:
        length = int(math.ceil((b - a) / step)) + 1
    return np.linspace(a, b, length)


def loggrid(a, b, **kw):
    x = np.exp(lingrid(math.log(a), math.log(b), **kw))
    x[0], x[-1] = a, b
    return x


def cospi(x):
    """cos(πx) with the argument reduced BEFORE the multiplication by π, as Julia's `cospi` (the reference's cosgrid uses it,
    src/utils.jl:285): exact at the half-integers, so the grid ends land on a and b bit for bit.  Quadrant reduction to |r| ≤ 1/4, then
    cos/sin of π·r."""
    x = np.abs(np.asarray(x, dtype=np.float64))
    r = x - 2.0 * np.floor(x / 2.0)                       # [0, 2), exact
    r = np.where(r > 1.0, 2.0 - r, r)                     # cos is even about 1: [0, 1]
    sign = np.where(r > 0.5, -1.0, 1.0)
    r = np.where(r > 0.5, 1.0 - r, r)                     # cos(π(1 − r)) = −cos(πr): [0, 1/2]
    return sign * np.where(r <= 0.25, np.cos(np.pi * r), np.sin(np.pi * (0.5 - r)))


def cosgrid(a, b, step=None, length=None):
    return a + (b - a) * (1 - cospi(lingrid(0.0, 0.5, step=None if step is None else step / math.pi, length=length)))


def chebpoints(order, a, b):
    i = np.arange(order + 1)
    return a + (b - a) * (1 + np.cos(np.pi * i / order)) / 2


def chebgrid(a, b, order):
    return chebpoints(order, a, b)[::-1].copy()


# ---------------------------------------------------------------------------------------------- models & parameters
class ModelSpec:
    def __init__(self, name, lmax, nx, w0wa):
        self.name, self.lmax, self.nx, self.w0wa = name, int(lmax), int(nx), bool(w0wa)

    def __repr__(self):
        return f"{self.name}(lmax={self.lmax}, nx={self.nx})"


def ΛCDM(lmax=10, nx=4, **_):
    return ModelSpec("ΛCDM", lmax, nx, False)


LCDM = ΛCDM


def w0waCDM(lmax=10, nx=4, **_):
    return ModelSpec("w0waCDM", lmax, nx, True)


def parameters_Planck18(M):
    h = 0.6736
    p = dict(h=h, T0=2.7255, Omega_c=0.1200 / h**2, Omega_b=0.0224 / h**2, YHe=0.2454, Neff=2.99, m_eV=0.02, N=3.0,
             ln_As1e10=math.log(2.099e-9 * 1e10), ns=0.965)
    if M.w0wa:
        p.update(w0=-0.9, wa=0.1, cs2=1.0)
    return p


_quad_cache = {}


def momentum_quadrature(N, L=100.0):
    """Cached wrapper of _momentum_quadrature (the rule depends only on N)."""
    if (N, L) not in _quad_cache:
        _quad_cache[(N, L)] = _momentum_quadrature(N, L)
    xs, Ws = _quad_cache[(N, L)]
    return xs.copy(), Ws.copy()


def _momentum_quadrature(N, L=100.0):
    """Gauss nodes/weights for ∫dx x² f0(x) g(x), f0 = 1/(eˣ+1), in u = 1/(1+x/L) (src/models/neutrinos.jl:55-60).
    Lanczos tridiagonalisation of the discretised measure + Golub-Welsch."""
    t, w = np.polynomial.legendre.leggauss(1500)
    u, wu = 0.5 * (t + 1), 0.5 * w
    x = L * (1 - u) / u
    with np.errstate(over="ignore"):
        mu = wu * (L / u**2) * x**2 / (np.exp(x) + 1)
    # Lanczos with starting vector sqrt(mu)
    q = np.sqrt(mu)
    beta0 = np.sum(mu)
    q /= np.linalg.norm(q)
    qprev = np.zeros_like(q)
    al, be = np.zeros(N), np.zeros(N)
    for j in range(N):
        v = u * q
        al[j] = q @ v
        v -= al[j] * q + (be[j - 1] if j else 0.0) * qprev
        for _ in range(2):  # full reorthogonalisation is unnecessary for N ≲ 32; one Gram-Schmidt refinement
            v -= (q @ v) * q
        be[j] = np.linalg.norm(v)
        qprev, q = q, v / be[j]
    J = np.diag(al) + np.diag(be[: N - 1], 1) + np.diag(be[: N - 1], -1)
    ev, V = np.linalg.eigh(J)
    Ws = beta0 * V[0] ** 2
    xs = L * (1 - ev) / ev
    o = np.argsort(xs)
    return xs[o], Ws[o]


def _cptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _h2d(a, dev):
    """Asynchronous upload through pinned memory (a pageable copy would first synchronise the stream, i.e. wait for kernels that are
    themselves waiting for SM slots held by a running persistent launch)."""
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().to(dev, non_blocking=True)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("symboltz.jl_b200: no CUDA device available -- the perturbation/LOS path is GPU-only (no CPU fallback)")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_libs = {}


def _load(path):
    if path not in _libs:
        _libs[path] = C.CDLL(path)
    return _libs[path]


def los_lib():
    lib = _load(_build.build_los())
    return lib


def comm_lib():
    return _load(_build.build_comm())


class Communicator:
    """NCCL communicator owned by the library (libsbc.so, `sbc_*`): what a host without torch.distributed (Julia over ccall) uses for the
    exchange steps of the sharded paths.  Rank 0 creates `Communicator.unique_id()` and hands the 128 bytes to the other
    ranks by any host-side means; every rank then constructs `Communicator(rank, world, id)` with its GPU current.  Accepted wherever the
    host API takes `group=` (spectrum_cmb, spectrum_matter_sweep)."""

    def __init__(self, rank, world, unique_id):
        _require_cuda()
        self.rank, self.world = int(rank), int(world)
        self._h = C.c_void_p()
        rc = comm_lib().sbc_comm_init(C.c_char_p(bytes(unique_id)), C.c_int(self.rank), C.c_int(self.world), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"sbc_comm_init failed with code {rc}")

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(int(comm_lib().sbc_unique_id_bytes()))
        rc = comm_lib().sbc_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"sbc_unique_id failed with code {rc}")
        return buf.raw

    def allreduce_sum(self, t):
        """In-place sum over ranks of a contiguous float64 device tensor (asynchronous on torch's current stream)."""
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("allreduce_sum: contiguous float64 device tensor expected")
        rc = comm_lib().sbc_allreduce_sum(self._h, _cptr(t), C.c_longlong(t.numel()), _stream())
        if rc != 0:
            raise RuntimeError(f"sbc_allreduce_sum failed with code {rc}")
        return t

    def close(self):
        if self._h:
            comm_lib().sbc_comm_destroy(self._h)
            self._h = C.c_void_p()


def _ranks(group):
    """(world, rank) of `group`: a library Communicator, a torch.distributed group, or the default group if one is initialised."""
    if isinstance(group, Communicator):
        return group.world, group.rank
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _allreduce(t, group):
    if isinstance(group, Communicator):
        return group.allreduce_sum(t)
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t


class CosmologyProblem:
    """Numerical problem for model M with parameters `pars` (reference CosmologyProblem, src/solve.jl:129-236).
    Building it generates + compiles the per-model CUDA engine (cached)."""

    def __init__(self, M, pars, ivspan=(1e-6, 100.0)):
        self.M, self.pars, self.ivspan = M, dict(pars), tuple(ivspan)
        so, self.info = _build.build_model(M.lmax, M.nx, M.w0wa)
        self.lib = _load(so)
        self.lib.sbm_solvebg.restype = C.c_int
        self.lib.sbm_key.restype = C.c_char_p
        inf = (C.c_int * 16)()
        self.lib.sbm_info(inf)
        self.N, self.npar, self.NBETA = inf[0], inf[1], inf[2]
        self.iP_kappa0, self.iP_tau0 = inf[13], inf[14]
        self.flops = dict(f=inf[8], lu=inf[9], solve=inf[10])
        self.xs, self.Ws = momentum_quadrature(M.nx)
        self.P = self._pvector()

    # dependent parameters (src/models/photons.jl:44-46, cosmologies.jl:78-84,221-227, neutrinos.jl:81-85, baryons.jl:151, inflation.jl:3-7)
    def _pvector(self):
        p = self.pars
        h, T0 = p["h"], p["T0"]
        H0 = _H100 * h
        Og = math.pi**2 / 15 * (_kB * T0) ** 4 / (_hbar**3 * _c**5) * 8 * math.pi * _GN / (3 * H0**2)
        Tnu = (4 / 11) ** (1 / 3) * T0
        Onu = p["Neff"] * 7 / 8 * (4 / 11) ** (4 / 3) * Og
        y0 = p["m_eV"] * _eV / (_kB * Tnu)
        Irho0 = float(np.sum(self.Ws * np.sqrt(self.xs**2 + y0**2)))
        Oh = p["N"] * 8 * math.pi / 3 * 2 / (2 * math.pi**2) * (_kB * Tnu) ** 4 / (_hbar * _c) ** 3 * Irho0 / ((H0 * _c) ** 2 / _GN)
        OL = 1 - (Og + Onu + p["Omega_c"] + p["Omega_b"] + Oh)
        fHe = p["YHe"] / (_mHe / _mH * (1 - p["YHe"]))
        self.derived = dict(Omega_g=Og, Omega_nu=Onu, Omega_h=Oh, Omega_L=OL, fHe=fHe, y0=y0, Irho0=Irho0,
                            As=math.exp(p["ln_As1e10"]) / 1e10, kpivot=0.05 / _Mpc / (_H100 / _c) / h)
        base = [h, p["Omega_c"], p["Omega_b"], Og, Onu, 3 / (8 * math.pi) * Oh / Irho0, OL, T0, p["YHe"], fHe, y0,
                p.get("w0", -1.0), p.get("wa", 0.0), p.get("cs2", 1.0), float("nan"), float("nan")]
        dl = -self.xs / (1 + np.exp(-self.xs))
        P = np.array(base + list(self.xs) + list(self.Ws) + list(dl), dtype=np.float64)
        assert len(P) == self.npar, (len(P), self.npar)
        return P

    def __repr__(self):
        return f"Cosmology problem for model {self.M}: background 5 unknowns; perturbations {self.N} unknowns, {self.info['nnz_full']} nonzeros in W"


def parameter_updater(prob, idxs):
    """θ ↦ new CosmologyProblem with parameters `idxs` replaced (reference src/solve.jl:272-307). Reuses the compiled engine."""
    idxs = list(idxs)

    def updater(theta):
        if isinstance(theta, dict):
            theta = [theta[n] for n in idxs]
        pars = dict(prob.pars)
        pars.update({n: float(v) for n, v in zip(idxs, theta)})
        return CosmologyProblem(prob.M, pars, prob.ivspan)

    return updater


RETCODES = {0: "Success", 1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable", 4: "ScheduleTimeout"}


class BackgroundSolution:
    """Background solve result: knots of the cubic Hermite spline of (a, _κ, XH⁺, XHe⁺, ΔT) (src/utils.jl:118-127)."""

    def __init__(self, prob, t, y, dy, info):
        self.prob, self.t, self.y, self.dy = prob, t, y, dy
        self.tau0, self.kappa0, self.taurec = float(info[0]), float(info[1]), float(info[2])
        self.retcode, self.naccept, self.nreject = int(info[3]), int(info[4]), int(info[5])
        self.dtevent = float(info[6]) if len(info) > 6 else 0.0  # length of the solver step in which a crossed 1 (lockstep lanes re-take it)
        self.P = prob.P.copy()
        self.P[prob.iP_kappa0], self.P[prob.iP_tau0] = self.kappa0, self.tau0  # callback semantics, src/solve.jl:183-189
        self._dev = None

    @property
    def success(self):
        return self.retcode == 0

    def device(self, msub=16):
        """Upload spline knots and parameters, build the β-table on the GPU (once).
        msub: sub-intervals per background knot interval (see sb_table_kernel)."""
        if self._dev is None or self._dev["msub"] != msub:
            _require_cuda()
            dev = torch.device("cuda")
            d = dict(P=torch.from_numpy(self.P).to(dev), t=torch.from_numpy(self.t).to(dev), y=torch.from_numpy(self.y).to(dev), dy=torch.from_numpy(self.dy).to(dev))
            nb = len(self.t)
            nnode = (nb - 1) * msub + 1
            d["tab"] = torch.empty((nnode, 2, self.prob.NBETA), dtype=torch.float64, device=dev)
            nlut = 4096
            s0 = math.log(self.t[0])
            dsl = (math.log(self.t[-1]) - s0) / nlut
            lut = np.clip(np.searchsorted(self.t, np.exp(s0 + dsl * np.arange(nlut)), side="right") - 1, 0, nb - 2).astype(np.int32)
            d.update(msub=msub, nlut=nlut, s0=s0, dsl=dsl, lut=torch.from_numpy(lut).to(dev))
            rc = self.prob.lib.sbm_build_table(_cptr(d["P"]), C.c_int(nb), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(msub), _cptr(d["tab"]), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_build_table failed with code {rc}")
            self._dev = d
        return self._dev

    def spline(self, tau):
        """Evaluate the Hermite spline (and its derivative) on the host."""
        yo, ypo = np.zeros(5), np.zeros(5)
        self.prob.lib.sbm_debug_spline(C.c_int(len(self.t)), _cptr(self.t), _cptr(self.y), _cptr(self.dy), C.c_double(tau), _cptr(yo), _cptr(ypo))
        return yo, ypo


def solvebg(prob, reltol=1e-7, abstol=1e-7):
    """Host background solve (reference solvebg, src/solve.jl:427-435; Rodas5P, "today" callback src/solve.jl:158-202)."""
    cap = 20000
    t, y, dy, info = np.zeros(cap), np.zeros((cap, 5)), np.zeros((cap, 5)), np.zeros(8)
    nb = prob.lib.sbm_solvebg(_cptr(prob.P), C.c_double(prob.ivspan[0]), C.c_double(prob.ivspan[1]), C.c_double(reltol), C.c_double(abstol), C.c_int(cap), _cptr(t), _cptr(y), _cptr(dy), _cptr(info))
    if nb <= 0:
        raise RuntimeError("background solve produced no knots")
    sol = BackgroundSolution(prob, t[:nb].copy(), y[:nb].copy(), dy[:nb].copy(), info)
    if not sol.success:
        warnings.warn(f"Background solution failed with return code {RETCODES.get(sol.retcode)}.\nCheck the parameters and precision settings!")
    return sol


def solvebg_lock(prob, primal):
    """Background of `prob` in LOCKSTEP with the finished solve `primal` of a neighbouring cosmology (`sbm_solvebg_lock`): the same
    step sequence (knots) and event step, no error control -- the result is then a smooth function of the parameters, which is what the
    parameter lanes of `sensitivity_matter` / `sensitivity_cmb` difference (the reference pushes ForwardDiff duals through the same
    Rodas5P steps, src/solve.jl:278-284)."""
    nfix = len(primal.t)
    cap = nfix + 64
    t, y, dy, info = np.zeros(cap), np.zeros((cap, 5)), np.zeros((cap, 5)), np.zeros(8)
    prob.lib.sbm_solvebg_lock.restype = C.c_int
    nb = prob.lib.sbm_solvebg_lock(_cptr(prob.P), C.c_double(prob.ivspan[0]), C.c_double(prob.ivspan[1]), C.c_int(nfix), _cptr(np.ascontiguousarray(primal.t)), C.c_double(primal.dtevent),
                                   C.c_int(cap), _cptr(t), _cptr(y), _cptr(dy), _cptr(info))
    if nb != nfix or info[7] != 0 or info[3] != 0:
        raise RuntimeError(f"lockstep background lost the primal's step sequence ({nb} knots for {nfix}, retcode {RETCODES.get(int(info[3]))}): the parameter step is too large")
    return BackgroundSolution(prob, t[:nb].copy(), y[:nb].copy(), dy[:nb].copy(), info)


def solvebg_batch(probs, reltol=1e-7, abstol=1e-7, cap=4096, warn=True):
    """Background solves of many cosmologies sharing one model structure in ONE kernel launch, one thread per cosmology
    (`sbm_solvebg_batch`; SURVEY §8f rank 1: in a parameter sweep the reference calls `solvebg` once per θ on the host,
    src/solve.jl:427-435 via docs/src/forecasting.md:56-59).  The device runs the same __host__ __device__ solver as `solvebg`
    (Rodas5P, "today" callback, τrec); results agree with the host solve to the solver tolerance, not bit for bit (the device
    exp/pow/log differ from glibc's in the last place, which moves the adaptive step sequence).
    Returns a list of BackgroundSolution (knots downloaded in one copy)."""
    _require_cuda()
    probs = list(probs)
    if not probs:
        return []
    p0 = probs[0]
    if any((p.M.lmax, p.M.nx, p.M.w0wa) != (p0.M.lmax, p0.M.nx, p0.M.w0wa) or p.ivspan != p0.ivspan for p in probs):
        raise ValueError("solvebg_batch: all problems must share the model structure and ivspan")
    n, npar = len(probs), p0.npar
    dev = torch.device("cuda")
    hP = torch.from_numpy(np.stack([p.P for p in probs])).pin_memory()
    dP = hP.to(dev, non_blocking=True)
    dt = torch.empty((n, cap), dtype=torch.float64, device=dev)
    dy = torch.empty((n, cap, 5), dtype=torch.float64, device=dev)
    ddy = torch.empty((n, cap, 5), dtype=torch.float64, device=dev)
    dinfo = torch.empty((n, 8), dtype=torch.float64, device=dev)
    dnb = torch.empty(n, dtype=torch.int32, device=dev)
    p0.lib.sbm_solvebg_batch.restype = C.c_int
    rc = p0.lib.sbm_solvebg_batch(C.c_int(n), _cptr(dP), C.c_double(p0.ivspan[0]), C.c_double(p0.ivspan[1]), C.c_double(reltol), C.c_double(abstol), C.c_int(cap),
                                  _cptr(dt), _cptr(dy), _cptr(ddy), _cptr(dinfo), _cptr(dnb), _stream())
    if rc != 0:
        raise RuntimeError(f"sbm_solvebg_batch failed with code {rc}")
    nb = dnb.cpu().numpy()
    info = dinfo.cpu().numpy()
    m = max(int(nb.max()), 2)
    t, y, d = dt[:, :m].cpu().numpy(), dy[:, :m].cpu().numpy(), ddy[:, :m].cpu().numpy()  # one strided copy each, only the used part
    sols = []
    for i, p in enumerate(probs):
        k = int(nb[i])
        if k <= 0:  # no knots, or more than `cap`: a failed solution (retcode MaxIters) like the host path reports, not an exception for the whole batch
            inf = info[i].copy()
            inf[3] = inf[3] if inf[3] != 0 else 1
            k = 2
            sol = BackgroundSolution(p, np.array([p.ivspan[0], p.ivspan[1]]), np.full((2, 5), np.nan), np.full((2, 5), np.nan), inf)
        else:
            sol = BackgroundSolution(p, t[i, :k].copy(), y[i, :k].copy(), d[i, :k].copy(), info[i])
        if warn and not sol.success:
            warnings.warn(f"Background solution {i} failed with return code {RETCODES.get(sol.retcode)}.\nCheck the parameters and precision settings!")
        sols.append(sol)
    return sols


class PerturbationSolution:
    def __init__(self, prob, bgsol, ks, tini, saveat, uend, usave, retcode, stats, dks):
        self.prob, self.bg, self.ks, self.tini, self.saveat = prob, bgsol, ks, tini, saveat
        self.d_uend, self.d_usave, self.d_retcode, self.d_stats, self.d_ks = uend, usave, retcode, stats, dks
        self.d_S = None  # fused source functions [nk][nS][nsave] (solvept(..., sources=...))
        self._rc = None

    @property
    def retcode(self):
        if self._rc is None:
            self._rc = self.d_retcode.cpu().numpy()
        return self._rc

    @property
    def stats(self):
        return self.d_stats.cpu().numpy()

    @property
    def uend(self):
        return self.d_uend.cpu().numpy()

    @property
    def usave(self):
        return None if self.d_usave is None else self.d_usave.cpu().numpy()

    @property
    def success(self):
        return bool((self.retcode == 0).all())


def build_schedule(cost, nlists, min_piece=24, attempts=None):
    """Static preemptive schedule of independent modes over `nlists` resident warps (McNaughton's wrap-around rule).

    The reference spawns one dynamic task per mode (src/solve.jl:566); the kernel's default is the same thing, an atomic queue
    in descending-k order.  With fewer than ~2 modes per resident warp that is far from balanced (two of the `nlists + 1`
    largest modes must share a warp), so here the modes are laid end to end in units of estimated attempted steps
    (`cost[i]`, any positive estimate) and cut into `nlists` equal chunks of length T = max(Σcost / nlists, max cost).
    A mode that straddles a cut runs its FIRST attempts (the part after the cut) as the first item of the next list, parks,
    and is finished as the LAST item of the previous list, so the two pieces never overlap in time when the estimate is exact.
    Pieces shorter than `min_piece` attempts are not split off.  `attempts` (optional): the modes' attempted steps when `cost` is in
    other units (e.g. measured time per mode): the quota of a split-off piece is an attempt count, so the cut position is converted
    with the mode's own attempts per unit of cost.  Returns (items[nitems, 3] int32 = (mode, quota, cont), ibeg[nlists + 1] int32, T)."""
    cost = np.asarray(cost, dtype=np.float64)
    apc = np.ones(len(cost)) if attempts is None else np.maximum(np.asarray(attempts, dtype=np.float64), 1.0) / np.maximum(cost, 1e-300)  # attempts per unit of cost
    cost = np.maximum(cost, 1.0 / apc)
    nk = len(cost)
    order = np.argsort(-cost, kind="stable")
    T = max(cost.sum() / nlists, cost.max()) * (1 + 1e-12)
    lists = [[] for _ in range(nlists)]
    tails = [None] * nlists  # continuation that ends list w
    pos = 0.0
    for m in order:
        c = cost[m]
        w = min(int(pos / T), nlists - 1)
        cut = (w + 1) * T
        end = pos + c
        if end <= cut or w == nlists - 1:
            lists[w].append((m, 0, 0))
        else:
            p1, p2 = (cut - pos) * apc[m], (end - cut) * apc[m]  # before / after the cut, in attempts
            if p2 < min_piece:
                lists[w].append((m, 0, 0))
            elif p1 < min_piece:
                lists[w + 1].insert(0, (m, 0, 0))
            else:
                lists[w + 1].insert(0, (m, int(round(p2)), 0))
                tails[w] = (m, 0, 1)
        pos = end
    items, ibeg = [], [0]
    for w in range(nlists):
        items.extend(lists[w])
        if tails[w] is not None:
            items.append(tails[w])
        ibeg.append(len(items))
    items = np.asarray(items, dtype=np.int32).reshape(-1, 3)
    assert sorted(items[items[:, 2] == 0, 0].tolist()) == list(range(nk))
    return items, np.asarray(ibeg, dtype=np.int32), T


class ModeCostModel:
    """Attempted Rosenbrock steps per mode as a smooth function of k, learnt from the statistics of a finished solve
    (piecewise linear in k through `nknots` quantile knots; the curve barely depends on the cosmology, scripts/cost_model.py)."""

    def __init__(self, ks, attempts, nknots=48):
        ks, attempts = np.asarray(ks, dtype=np.float64), np.asarray(attempts, dtype=np.float64)
        o = np.argsort(ks)
        ks, attempts = ks[o], attempts[o]
        edges = np.unique(np.linspace(0, len(ks), min(nknots, len(ks)) + 1).astype(int))
        self.kn = np.array([ks[0]] + [ks[a:b].mean() for a, b in zip(edges[:-1], edges[1:])] + [ks[-1]])
        self.cn = np.array([attempts[0]] + [attempts[a:b].mean() for a, b in zip(edges[:-1], edges[1:])] + [attempts[-1]])
        self.kn, i = np.unique(self.kn, return_index=True)
        self.cn = self.cn[i]

    def __call__(self, ks):
        return np.interp(np.asarray(ks, dtype=np.float64), self.kn, self.cn)


def resident_warps(prob, batch=False):
    """Warps the integrator keeps resident on this GPU (= lists of a static schedule); `batch`: of the batched instantiation."""
    return int(prob.lib.sbm_resident_warps_batch() if batch else prob.lib.sbm_resident_warps())


class SbmSrc(C.Structure):
    """sbm_src_t (include/symboltz_b200.h): request for the fused source evaluation of the *_src solves."""
    _fields_ = [("dsrcbg", C.c_void_p), ("dS", C.c_void_p), ("nS", C.c_int), ("scale_k", C.c_int), ("taurec", C.c_double)]


def _src_request(sources, taurec_default=0.0):
    nS = int(sources.get("nS", 2))
    if nS not in (2, 3):
        raise ValueError("sources: nS must be 2 (ST, SE) or 3 (+ Sψ)")
    return nS, 1 if sources.get("scale_k", True) else 0, float(sources.get("taurec", taurec_default))


def source_background(prob, d, nb, dsave):
    """Per-save-time background table of the source evaluation (`sbm_srcbg`): κ̇, κ̈, κ⃛, e^{−κ}, τ0 − τ, β_m and their flow derivatives."""
    out = torch.empty((len(dsave), int(prob.lib.sbm_srcbg_stride())), dtype=torch.float64, device=dsave.device)
    rc = prob.lib.sbm_srcbg(_cptr(d["P"]), C.c_int(nb), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(len(dsave)), _cptr(dsave), _cptr(out), _stream())
    if rc != 0:
        raise RuntimeError(f"sbm_srcbg failed with code {rc}")
    return out


def ptalg(prob=None, accuracy=2):
    """Perturbation integrator by accuracy level (reference ptalg(prob; accuracy), src/solve.jl:326-341): 0 -> "TRBDF2", 1 -> "KenCarp4", otherwise "Rodas5P";
    pass the result as `alg=` to solvept / solve(ptopts=...) / spectrum_matter / spectrum_cmb(ptopts=...)."""
    return "TRBDF2" if accuracy == 0 else ("KenCarp4" if accuracy == 1 else "Rodas5P")


# Which mapping for a launch?  The split kernel's CTAs are persistent (atomic queue, descending k): with more modes than CTAs fit, a second wave of (cheaper) modes
# follows on the CTAs that finish first.  At ≈8.5 us per attempt against ≈12.5 us for a lone warp that pays as long as the work per CTA stays below 1.5 x the slowest mode:
# decided from a crude estimate of the attempts per mode, a(k) = 250 + 27 k^0.6 (k in H0/c; within 25 % of the step counters of the LCDM models at the default tolerance).
# Measured (scripts/split_waves.py): C_l grid 253 / 404 / 505 / 673 modes: split 23.1 / 25.0 / 32.6 / 38.0 ms, warp per mode 32.0 / 32.1 / 32.2 / 36.2 ms; P(k) log grids of
# 300 ... 800 modes: 27.5 ... 27.9 ms against 38.1 ... 39.7 ms.
SPLIT_WAVES = 4  # hard limit on the launch size, in units of split_capacity


def split_pays(prob, ks):
    cap = split_capacity(prob)
    ks = np.asarray(ks, dtype=np.float64)
    if cap <= 0 or len(ks) == 0 or len(ks) > SPLIT_WAVES * cap:
        return False
    if len(ks) <= cap:
        return True
    a = 250.0 + 27.0 * np.maximum(np.nan_to_num(ks, nan=0.0), 0.0) ** 0.6
    t_split = 8.5 * max(a.max(), 1.15 * a.sum() / cap)
    t_warp = 12.5 * max(a.max(), a.sum() / max(1, resident_warps(prob)))
    return bool(t_split < t_warp)


def split_capacity(prob):
    """Modes that run concurrently under the split mapping (one CTA of SB_R warps per mode, `sbm_solvept_split`); 0 if the model has none."""
    if not hasattr(prob, "_split_cap"):
        prob._split_cap = int(prob.lib.sbm_split_capacity()) if hasattr(prob.lib, "sbm_split_capacity") else 0
    return prob._split_cap


def solvept(prob, bgsol, ks, ptivini=-math.inf, reltol=1e-5, abstol=1e-5, saveat=None, maxiters=100000, msub=16, nctas=0, warn=True, sync=True, trace=0, cost=None, sources=None,
            keep_states=True, split=None, cost_attempts=None, alg="Rodas5P"):
    """Perturbation solve over independent k-modes on the GPU (reference solvept, src/solve.jl:543-569).
    ks in H0/c.  ptivini: number or callable k -> τini (clamped to the background span, src/solve.jl:527).
    cost: optional per-mode estimate of attempted steps (array or vectorised callable ks -> cost, e.g. a ModeCostModel): run under the static preemptive
    schedule of `build_schedule` instead of the atomic queue (same results, better balance for few modes per warp).
    sources: dict(nS = 2 | 3, scale_k = True, taurec = bgsol.taurec) -- evaluate the CMB source functions at the `saveat` times INSIDE the
    integrator (the reference's output_func, src/observables/fourier.jl:272-278) into sol.d_S[nk][nS][nsave]; with keep_states = False the saved
    states never leave the SM (sol.d_usave is None).
    split: None (default) = choose the mapping by the size of the launch (`split_pays`): with no more modes than `split_capacity(prob)` (296 on a B200 for the nx = 4
    models), or somewhat more when most of them are cheap, every mode gets a CTA of SB_R warps (`sbm_solvept_split`: the row-parallel phases of an attempt are spread over the
    warps; ≈40 % lower latency, bit-identical results), otherwise one warp per mode; True / False force one or the other.
    alg: "Rodas5P" (the reference's default, `ptalg(prob; accuracy = 2)`), "KenCarp4" (`accuracy = 1`: the ESDIRK half of ARK4(3)6L[2]SA, five factorisations
    and six solves per step) or "TRBDF2" (`accuracy = 0`: second order, two factorisations and three solves per step), src/solve.jl:333-337; published schemes
    with every stage solved exactly (the system is linear) -- OrdinaryDiffEq.jl's own step selection cannot be pinned here; one warp per mode, queue only."""
    _require_cuda()
    ks = np.ascontiguousarray(np.atleast_1d(ks), dtype=np.float64)
    nk = len(ks)
    f = ptivini if callable(ptivini) else (lambda k: ptivini)
    with np.errstate(divide="ignore", invalid="ignore"):
        tini = np.array([min(max(f(k), bgsol.t[0]), bgsol.t[-1]) if k == k else bgsol.t[0] for k in ks], dtype=np.float64)
    order = np.argsort(-np.nan_to_num(ks, nan=0.0), kind="stable").astype(np.int32)  # most expensive (largest k) first
    d = bgsol.device(msub)
    dev = d["P"].device
    host = torch.from_numpy(np.concatenate([ks, tini])).pin_memory()
    dkt = host.to(dev, non_blocking=True)
    dks, dtini = dkt[:nk], dkt[nk:]
    dorder = torch.from_numpy(order).to(dev)
    N = prob.N
    uend = torch.empty((nk, N), dtype=torch.float64, device=dev)
    retcode = torch.empty(nk, dtype=torch.int32, device=dev)
    stats = torch.empty((nk, 4), dtype=torch.int64, device=dev)
    queue = torch.zeros(1, dtype=torch.int32, device=dev)
    src, dS, keep = None, None, None
    if saveat is not None:
        saveat = np.ascontiguousarray(saveat, dtype=np.float64)
        dsave = torch.from_numpy(saveat).to(dev)
        ns = len(saveat)
        if sources is not None and ns > 0:
            nS, sk, trec = _src_request(sources, bgsol.taurec)
            keep = source_background(prob, d, len(bgsol.t), dsave)
            dS = torch.empty((nk, nS, ns), dtype=torch.float64, device=dev)
            src = SbmSrc(keep.data_ptr(), dS.data_ptr(), nS, sk, trec)
        usave = torch.empty((nk, ns, N), dtype=torch.float64, device=dev) if (keep_states or src is None) else None
    else:
        dsave, usave, ns = None, None, 0
    srcp = C.byref(src) if src is not None else None
    dtrace = torch.zeros((trace, 3), dtype=torch.float64, device=dev) if trace else None  # debug: (t, dt, EEst) of mode 0
    if alg not in ("Rodas5P", "TRBDF2", "KenCarp4"):
        raise ValueError(f"unknown perturbation integrator {alg!r}: Rodas5P, KenCarp4 or TRBDF2 (the reference's ptalg(prob; accuracy = 2 / 1 / 0))")
    if alg != "Rodas5P":
        if cost is not None or trace or nctas or split is True:
            raise ValueError(f"{alg} runs one warp per mode from the atomic queue: no cost=, trace=, nctas= or split=True")
        rc = (prob.lib.sbm_solvept_trbdf2 if alg == "TRBDF2" else prob.lib.sbm_solvept_kencarp4)(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(d["msub"]), C.c_int(d["nlut"]), C.c_double(d["s0"]), C.c_double(d["dsl"]), _cptr(d["lut"]), _cptr(d["tab"]),
                                         C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dorder), C.c_double(bgsol.tau0), C.c_int(ns), _cptr(dsave), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                         _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), _stream(), srcp)
    elif cost is not None and nk > 0:
        cvec = np.asarray(cost(ks) if callable(cost) else cost, dtype=np.float64) * np.ones(nk)
        wpc = int(prob.lib.sbm_warps_per_cta())
        nlists = max(wpc, min(resident_warps(prob), nk) // wpc * wpc)
        items, ibeg, _ = build_schedule(np.nan_to_num(cvec, nan=1.0), nlists, attempts=cost_attempts)
        ditems, dibeg = torch.from_numpy(items).to(dev), torch.from_numpy(ibeg).to(dev)
        dcont = torch.empty(nk * int(prob.lib.sbm_cont_stride()), dtype=torch.float64, device=dev)
        dflags = torch.zeros(nk, dtype=torch.int32, device=dev)
        rc = prob.lib.sbm_solvept_sched_src(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(d["msub"]), C.c_int(d["nlut"]), C.c_double(d["s0"]), C.c_double(d["dsl"]), _cptr(d["lut"]), _cptr(d["tab"]),
                                            C.c_int(nk), _cptr(dks), _cptr(dtini), C.c_double(bgsol.tau0), C.c_int(ns), _cptr(dsave), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                            _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), _cptr(ditems), _cptr(dibeg), C.c_int(nlists), _cptr(dcont), _cptr(dflags), _stream(), srcp)
    elif (split is True or (split is None and trace == 0 and nctas == 0 and split_pays(prob, ks))) and split_capacity(prob) > 0:
        rc = prob.lib.sbm_solvept_split(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(d["msub"]), C.c_int(d["nlut"]), C.c_double(d["s0"]), C.c_double(d["dsl"]), _cptr(d["lut"]), _cptr(d["tab"]),
                                        C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dorder), C.c_double(bgsol.tau0), C.c_int(ns), _cptr(dsave), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                        _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), _stream(), srcp)
    elif src is not None:
        rc = prob.lib.sbm_solvept_src(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(d["msub"]), C.c_int(d["nlut"]), C.c_double(d["s0"]), C.c_double(d["dsl"]), _cptr(d["lut"]), _cptr(d["tab"]),
                                      C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dorder), C.c_double(bgsol.tau0), C.c_int(ns), _cptr(dsave), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                      _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), C.c_int(nctas), _stream(), srcp)
    else:
        rc = prob.lib.sbm_solvept(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(d["msub"]), C.c_int(d["nlut"]), C.c_double(d["s0"]), C.c_double(d["dsl"]), _cptr(d["lut"]), _cptr(d["tab"]),
                                  C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dorder), C.c_double(bgsol.tau0), C.c_int(ns), _cptr(dsave), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                  _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), C.c_int(nctas), _stream(), _cptr(dtrace), C.c_int(trace))
    if rc < 0:
        raise RuntimeError(f"sbm_solvept failed with code {rc}")
    sol = PerturbationSolution(prob, bgsol, ks, tini, saveat, uend, usave, retcode, stats, dks)
    sol.d_S, sol._keep = dS, keep
    sol.grid = rc
    sol.trace = dtrace.cpu().numpy() if trace else None
    if sync and warn:
        for i in np.nonzero(sol.retcode != 0)[0]:  # warn, don't throw (src/solve.jl:557-560)
            warnings.warn(f"Perturbation (mode k = {ks[i]}) solution failed with return code {RETCODES.get(int(sol.retcode[i]))}.\nCheck the parameters and precision settings!")
    return sol


# sbm_cosmo_t (include/symboltz_b200.h): natural C layout, 128 bytes
COSMO_DTYPE = np.dtype({"names": ["P", "nb", "t", "y", "dy", "tb_nb", "msub", "nlut", "s0", "inv_dsl", "tb_t", "lut", "tab", "tend", "saveat", "srcbg", "taurec"],
                        "formats": ["u8", "i4", "u8", "u8", "u8", "i4", "i4", "i4", "f8", "f8", "u8", "u8", "u8", "f8", "u8", "u8", "f8"],
                        "offsets": [0, 8, 16, 24, 32, 40, 44, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120], "itemsize": 128})


def cosmo_record(bgsol, msub=16, dsave=None):
    """The sbm_cosmo_t of a background solution (uploads knots and builds its β-table on first use)."""
    d = bgsol.device(msub)
    r = np.zeros((), dtype=COSMO_DTYPE)
    nb = len(bgsol.t)
    r["P"], r["nb"], r["t"], r["y"], r["dy"] = d["P"].data_ptr(), nb, d["t"].data_ptr(), d["y"].data_ptr(), d["dy"].data_ptr()
    r["tb_nb"], r["msub"], r["nlut"], r["s0"], r["inv_dsl"] = nb, d["msub"], d["nlut"], d["s0"], 1.0 / d["dsl"]
    r["tb_t"], r["lut"], r["tab"], r["tend"] = d["t"].data_ptr(), d["lut"].data_ptr(), d["tab"].data_ptr(), bgsol.tau0
    r["saveat"] = 0 if dsave is None else dsave.data_ptr()
    r["srcbg"], r["taurec"] = 0, bgsol.taurec
    return r


class CosmoArena:
    """Reusable staging for a batch of background cosmologies: one pinned host buffer and ONE host->device copy for all knots and
    parameters, one for the interval look-ups, one β-table buffer -- instead of five small uploads and a 9 MB allocation per
    cosmology.  `load` returns the sbm_cosmo_t records; the per-cosmology device views stay valid until the next `load`."""

    NLUT = 4096

    def __init__(self):
        self.h_f = self.d_f = self.h_i = self.d_i = self.d_tab = None
        self.views = []

    @staticmethod
    def _grow(t, n, **kw):
        return t if t is not None and t.numel() >= n else torch.empty(int(n * 1.25) + 64, **kw)

    def load(self, bgsols, msub=16):
        prob = bgsols[0].prob
        dev = torch.device("cuda")
        npar, NB = prob.npar, prob.NBETA
        nbs = np.array([len(b.t) for b in bgsols])
        fo = np.concatenate([[0], np.cumsum(npar + 11 * nbs)])          # doubles: P | t | y | dy per cosmology
        to = np.concatenate([[0], np.cumsum(((nbs - 1) * msub + 1) * 2 * NB)])
        nc = len(bgsols)
        self.h_f = self._grow(self.h_f, fo[-1], dtype=torch.float64, pin_memory=True)
        self.d_f = self._grow(self.d_f, fo[-1], dtype=torch.float64, device=dev)
        self.h_i = self._grow(self.h_i, nc * self.NLUT, dtype=torch.int32, pin_memory=True)
        self.d_i = self._grow(self.d_i, nc * self.NLUT, dtype=torch.int32, device=dev)
        self.d_tab = self._grow(self.d_tab, to[-1], dtype=torch.float64, device=dev)
        hf, hi = self.h_f.numpy(), self.h_i.numpy()
        recs = np.zeros(nc, dtype=COSMO_DTYPE)
        self.views = []
        q = np.arange(self.NLUT)
        for i, b in enumerate(bgsols):
            nb, o = int(nbs[i]), int(fo[i])
            hf[o:o + npar] = b.P
            hf[o + npar:o + npar + nb] = b.t
            hf[o + npar + nb:o + npar + 6 * nb] = b.y.ravel()
            hf[o + npar + 6 * nb:o + npar + 11 * nb] = b.dy.ravel()
            s0 = math.log(b.t[0])
            dsl = (math.log(b.t[-1]) - s0) / self.NLUT
            hi[i * self.NLUT:(i + 1) * self.NLUT] = np.clip(np.searchsorted(b.t, np.exp(s0 + dsl * q), side="right") - 1, 0, nb - 2)
            dP, dt, dy, ddy = (self.d_f[o:o + npar], self.d_f[o + npar:o + npar + nb], self.d_f[o + npar + nb:o + npar + 6 * nb], self.d_f[o + npar + 6 * nb:o + npar + 11 * nb])
            tab = self.d_tab[int(to[i]):int(to[i + 1])]
            lut = self.d_i[i * self.NLUT:(i + 1) * self.NLUT]
            self.views.append(dict(P=dP, t=dt, y=dy, dy=ddy, tab=tab, lut=lut, nb=nb))
            r = recs[i]
            r["P"], r["nb"], r["t"], r["y"], r["dy"] = dP.data_ptr(), nb, dt.data_ptr(), dy.data_ptr(), ddy.data_ptr()
            r["tb_nb"], r["msub"], r["nlut"], r["s0"], r["inv_dsl"] = nb, msub, self.NLUT, s0, 1.0 / dsl
            r["tb_t"], r["lut"], r["tab"], r["tend"] = dt.data_ptr(), lut.data_ptr(), tab.data_ptr(), b.tau0
            r["srcbg"], r["taurec"] = 0, b.taurec
        self.d_f[:fo[-1]].copy_(self.h_f[:fo[-1]], non_blocking=True)
        self.d_i[:nc * self.NLUT].copy_(self.h_i[:nc * self.NLUT], non_blocking=True)
        for v in self.views:
            rc = prob.lib.sbm_build_table(_cptr(v["P"]), C.c_int(v["nb"]), _cptr(v["t"]), _cptr(v["y"]), _cptr(v["dy"]), C.c_int(msub), _cptr(v["tab"]), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_build_table failed with code {rc}")
        return recs


class BatchSolution:
    """Result of `solvept_batch`: flat device arrays over all (cosmology, mode) pairs + per-cosmology views."""

    def __init__(self, sols, offsets, d_uend, d_usave, d_retcode, d_stats, keep):
        self.sols, self.offsets, self.d_uend, self.d_usave, self.d_retcode, self.d_stats, self._keep = sols, offsets, d_uend, d_usave, d_retcode, d_stats, keep

    def __len__(self):
        return len(self.sols)

    def __getitem__(self, i):
        return self.sols[i]

    @property
    def success(self):
        return bool((self.d_retcode == 0).all().item())


def solvept_batch(bgsols, ks, ptivini=-math.inf, reltol=1e-5, abstol=1e-5, saveat=None, maxiters=100000, msub=16, cost=None, arena=None, max_lists=None, sources=None,
                  keep_states=True):
    """Perturbation solve of several cosmologies (same model structure) in ONE integrator launch over all (cosmology, mode)
    pairs, ordered by descending k across cosmologies (SURVEY §8b batched variant; the reference runs one `solvept` per
    cosmology, docs/src/forecasting.md:56-59).  bgsols: BackgroundSolutions; ks: one array for all, or one array per cosmology;
    saveat: None, or one array of save times per cosmology (equal lengths).  Per-mode results are bit-identical to `solvept` on that
    cosmology.  cost: optional vectorised callable ks -> estimated attempts, switches to the static preemptive schedule; every list of
    a static schedule must be resident from the start of the launch, so callers that keep several scheduled launches in flight
    (the sweep's stream slots) pass max_lists = resident warps / launches in flight.
    sources / keep_states: as in `solvept` (fused source evaluation at each cosmology's own save times; taurec of each cosmology)."""
    _require_cuda()
    nc = len(bgsols)
    prob = bgsols[0].prob
    if any(b.prob.N != prob.N or b.prob.lib._name != prob.lib._name for b in bgsols):
        raise ValueError("solvept_batch: all cosmologies must share one model structure")
    ks_list = [np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64) for k in (ks if isinstance(ks, (list, tuple)) else [ks] * nc)]
    if len(ks_list) != nc:
        raise ValueError("solvept_batch: need one k-array per cosmology")
    f = ptivini if callable(ptivini) else (lambda k: ptivini)
    offsets = np.concatenate([[0], np.cumsum([len(k) for k in ks_list])]).astype(np.int64)
    nk = int(offsets[-1])
    kall = np.concatenate(ks_list)
    with np.errstate(divide="ignore", invalid="ignore"):
        tini = np.concatenate([np.array([min(max(f(k), b.t[0]), b.t[-1]) if k == k else b.t[0] for k in kk], dtype=np.float64) for b, kk in zip(bgsols, ks_list)])
    cosmo_of = np.repeat(np.arange(nc, dtype=np.int32), [len(k) for k in ks_list])
    order = np.argsort(-np.nan_to_num(kall, nan=0.0), kind="stable").astype(np.int32)
    dev = torch.device("cuda")
    ns = 0
    dsaves = [None] * nc
    if saveat is not None:
        saveat = [np.ascontiguousarray(sv, dtype=np.float64) for sv in saveat]
        ns = len(saveat[0])
        if len(saveat) != nc or any(len(sv) != ns for sv in saveat):
            raise ValueError("solvept_batch: need one saveat array per cosmology, all of the same length")
        dsv = _h2d(np.stack(saveat), dev)
        dsaves = [dsv[i] for i in range(nc)]
    arena = arena if arena is not None else CosmoArena()
    recs = arena.load(bgsols, msub)  # uploads knots, builds the β-tables (same kernel and node layout as BackgroundSolution.device)
    src, dS, dsb = None, None, None
    if sources is not None and ns > 0:
        nS, sk, _ = _src_request(sources)
        stride = int(prob.lib.sbm_srcbg_stride())
        dsb = torch.empty((nc, ns, stride), dtype=torch.float64, device=dev)
        for i, v in enumerate(arena.views):
            rc = prob.lib.sbm_srcbg(_cptr(v["P"]), C.c_int(v["nb"]), _cptr(v["t"]), _cptr(v["y"]), _cptr(v["dy"]), C.c_int(ns), _cptr(dsaves[i]), _cptr(dsb[i]), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_srcbg failed with code {rc}")
            recs[i]["srcbg"] = dsb[i].data_ptr()
        dS = torch.empty((int(offsets[-1]), nS, ns), dtype=torch.float64, device=dev)
        src = SbmSrc(None, dS.data_ptr(), nS, sk, 0.0)
    for i in range(nc):
        recs[i]["saveat"] = 0 if dsaves[i] is None else dsaves[i].data_ptr()
    dcos = _h2d(np.frombuffer(recs.tobytes(), dtype=np.uint8).reshape(nc, COSMO_DTYPE.itemsize).copy(), dev)
    dkt = _h2d(np.concatenate([kall, tini]), dev)
    dks, dtini = dkt[:nk], dkt[nk:]
    dcof, dorder = _h2d(cosmo_of, dev), _h2d(order, dev)
    N = prob.N
    uend = torch.empty((nk, N), dtype=torch.float64, device=dev)
    usave = torch.empty((nk, ns, N), dtype=torch.float64, device=dev) if (ns and (keep_states or src is None)) else None
    retcode = torch.empty(nk, dtype=torch.int32, device=dev)
    stats = torch.empty((nk, 4), dtype=torch.int64, device=dev)
    queue = torch.zeros(1, dtype=torch.int32, device=dev)
    ditems = dibeg = dcont = dflags = None
    nlists = 0
    if cost is not None and nk > 0:
        wpc = int(prob.lib.sbm_warps_per_cta())
        nres = resident_warps(prob, batch=True)
        nlists = max(wpc, min(nres if max_lists is None else min(nres, int(max_lists)), nk) // wpc * wpc)
        items, ibeg, _ = build_schedule(np.nan_to_num(np.asarray(cost(kall) if callable(cost) else cost, dtype=np.float64) * np.ones(nk), nan=1.0), nlists)
        ditems, dibeg = _h2d(items, dev), _h2d(ibeg, dev)
        dcont = torch.empty(nk * int(prob.lib.sbm_cont_stride()), dtype=torch.float64, device=dev)
        dflags = torch.zeros(nk, dtype=torch.int32, device=dev)
    rc = prob.lib.sbm_solvept_batch_src(C.c_int(nc), _cptr(dcos), C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dcof), _cptr(dorder), C.c_int(ns), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                        _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), _cptr(ditems), _cptr(dibeg), C.c_int(nlists), _cptr(dcont), _cptr(dflags), _stream(),
                                        C.byref(src) if src is not None else None)
    if rc < 0:
        raise RuntimeError(f"sbm_solvept_batch failed with code {rc}")
    sols = []
    for i, b in enumerate(bgsols):
        a, e = int(offsets[i]), int(offsets[i + 1])
        sols.append(PerturbationSolution(b.prob, b, ks_list[i], tini[a:e], None if saveat is None else saveat[i], uend[a:e], None if usave is None else usave[a:e], retcode[a:e], stats[a:e], dks[a:e]))
        sols[-1].d_S = None if dS is None else dS[a:e]
    return BatchSolution(sols, offsets, uend, usave, retcode, stats, keep=(arena, dcos, dkt, dcof, dorder, queue, ditems, dibeg, dcont, dflags, dsaves, dsb, dS))


def solvept_lanes(bgsols, ks, invdelta, ptivini=-math.inf, reltol=1e-5, abstol=1e-5, saveat=None, maxiters=100000, msub=16, sources=None, keep_states=False):
    """Perturbation solve of G = len(bgsols) ≤ 8 neighbouring cosmologies (lane 0: the primal; lane j: one parameter moved by δ_j =
    1/invdelta[j]) for the SAME wavenumbers `ks` in lockstep: one CTA of G warps per mode, ONE step controller whose error norm covers
    the primal and the partials (u^j − u^0)/δ_j (`sbm_solvept_lanes`; BASELINE config 5 "dual-number lanes in the batched solve";
    reference: Dual parameters through solvept, test/runtests.jl:363-422).  saveat: one array of save times per lane (equal lengths;
    save decisions follow the primal's).  Returns one PerturbationSolution per lane (contiguous per-lane arrays)."""
    _require_cuda()
    G = len(bgsols)
    prob = bgsols[0].prob
    if not 2 <= G <= 8:
        raise ValueError("solvept_lanes: 2 to 8 lanes")
    if any(b.prob.N != prob.N or b.prob.lib._name != prob.lib._name for b in bgsols):
        raise ValueError("solvept_lanes: all lanes must share one model structure")
    ks = np.ascontiguousarray(np.atleast_1d(ks), dtype=np.float64)
    nk = len(ks)
    f = ptivini if callable(ptivini) else (lambda k: ptivini)
    b0 = bgsols[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        tini = np.array([min(max(f(k), b0.t[0]), b0.t[-1]) if k == k else b0.t[0] for k in ks], dtype=np.float64)
    dev = torch.device("cuda")
    ns, dsaves = 0, [None] * G
    if saveat is not None:
        saveat = [np.ascontiguousarray(sv, dtype=np.float64) for sv in saveat]
        ns = len(saveat[0])
        if len(saveat) != G or any(len(sv) != ns for sv in saveat):
            raise ValueError("solvept_lanes: need one saveat array per lane, all of the same length")
        dsv = _h2d(np.stack(saveat), dev)
        dsaves = [dsv[i] for i in range(G)]
    arena = CosmoArena()
    recs = arena.load(bgsols, msub)
    src, dS, dsb = None, None, None
    if sources is not None and ns > 0:
        nS, sk, _ = _src_request(sources)
        dsb = torch.empty((G, ns, int(prob.lib.sbm_srcbg_stride())), dtype=torch.float64, device=dev)
        for i, v in enumerate(arena.views):
            rc = prob.lib.sbm_srcbg(_cptr(v["P"]), C.c_int(v["nb"]), _cptr(v["t"]), _cptr(v["y"]), _cptr(v["dy"]), C.c_int(ns), _cptr(dsaves[i]), _cptr(dsb[i]), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_srcbg failed with code {rc}")
            recs[i]["srcbg"] = dsb[i].data_ptr()
        dS = torch.empty((nk * G, nS, ns), dtype=torch.float64, device=dev)
        src = SbmSrc(None, dS.data_ptr(), nS, sk, 0.0)
    for i in range(G):
        recs[i]["saveat"] = 0 if dsaves[i] is None else dsaves[i].data_ptr()
    dcos = _h2d(np.frombuffer(recs.tobytes(), dtype=np.uint8).reshape(G, COSMO_DTYPE.itemsize).copy(), dev)
    dkt = _h2d(np.concatenate([np.repeat(ks, G), np.repeat(tini, G)]), dev)
    dks, dtini = dkt[:nk * G], dkt[nk * G:]
    dcof = _h2d(np.tile(np.arange(G, dtype=np.int32), nk), dev)
    dorder = _h2d(np.argsort(-np.nan_to_num(ks, nan=0.0), kind="stable").astype(np.int32), dev)
    N = prob.N
    uend = torch.empty((nk * G, N), dtype=torch.float64, device=dev)
    usave = torch.empty((nk * G, ns, N), dtype=torch.float64, device=dev) if (ns and (keep_states or src is None)) else None
    retcode = torch.empty(nk * G, dtype=torch.int32, device=dev)
    stats = torch.empty((nk * G, 4), dtype=torch.int64, device=dev)
    queue = torch.zeros(1, dtype=torch.int32, device=dev)
    invd = np.ascontiguousarray(np.concatenate([[0.0], np.asarray(invdelta, dtype=np.float64)[1:G]]) if len(invdelta) == G else np.concatenate([[0.0], np.asarray(invdelta, dtype=np.float64)]))
    if len(invd) != G:
        raise ValueError("solvept_lanes: invdelta needs one entry per lane (the primal's is ignored) or one per non-primal lane")
    prob.lib.sbm_solvept_lanes.restype = C.c_int
    rc = prob.lib.sbm_solvept_lanes(C.c_int(G), _cptr(dcos), C.c_int(nk), _cptr(dks), _cptr(dtini), _cptr(dcof), _cptr(dorder), C.c_int(ns), C.c_double(reltol), C.c_double(abstol), C.c_int(maxiters),
                                    _cptr(usave), _cptr(uend), _cptr(retcode), _cptr(stats), _cptr(queue), _cptr(invd), C.c_double(min(b.tau0 for b in bgsols)), _stream(),
                                    C.byref(src) if src is not None else None)
    if rc < 0:
        raise RuntimeError(f"sbm_solvept_lanes failed with code {rc}")
    sols = []
    for j, b in enumerate(bgsols):  # [mode][lane] -> contiguous per-lane arrays
        sol = PerturbationSolution(b.prob, b, ks, tini, None if saveat is None else saveat[j], uend.view(nk, G, N)[:, j].contiguous(),
                                   None if usave is None else usave.view(nk, G, ns, N)[:, j].contiguous(), retcode.view(nk, G)[:, j].contiguous(), stats.view(nk, G, 4)[:, j].contiguous(),
                                   dks.view(nk, G)[:, j].contiguous())
        sol.d_S = None if dS is None else dS.view(nk, G, dS.shape[1], ns)[:, j].contiguous()
        sols.append(sol)
    sols[0]._keep = (arena, dcos, dkt, dcof, dorder, queue, dsaves, dsb)
    return sols


class CosmologySolution:
    """Result of `solve(prob, ks)` (reference CosmologySolution, src/solve.jl:343-378).  Calling it, `sol(vars, τs, ks)`, evaluates perturbation
    unknowns at arbitrary times and wavenumbers (reference src/solve.jl:720-776): dense output in time for the solved modes, linear
    interpolation in ktransform(k) (default ln k) between the two neighbouring solved modes."""

    def __init__(self, prob, bg, ks, pts, ptivini=-math.inf, ptopts=None):
        self.prob, self.bg, self.ks, self.pts = prob, bg, ks, pts
        self._ptivini, self._ptopts, self._dense = ptivini, dict(ptopts or {}), {}

    def _states_at(self, taus):
        """Dense output of every solved mode at `taus`: one launch with saveat (the solve is deterministic, so this is the interpolant of
        the step sequence `solve` took); cached per time grid."""
        key = taus.tobytes()
        if key not in self._dense:
            opts = {k: v for k, v in self._ptopts.items() if k not in ("saveat", "sources", "keep_states")}
            self._dense = {key: solvept(self.prob, self.bg, self.ks, self._ptivini, saveat=taus, warn=False, **opts).usave}
        return self._dense[key]

    def __call__(self, vars, taus, ks, ktransform=np.log):
        if self.ks is None or len(self.ks) == 0:
            raise RuntimeError("No perturbations solved for. Pass ks to solve().")
        if (np.diff(self.ks) < 0).any():
            raise RuntimeError("Solution wavenumbers are not sorted in ascending order")
        scal = (isinstance(vars, str), np.ndim(taus) == 0, np.ndim(ks) == 0)
        names = self.prob.info["unames"]
        observed = {"ST": 0, "SE": 1, "Spsi": 2, "Sψ": 2}  # the CMB source functions (reference M.ST, M.SE, M.Sψ, src/models/cosmologies.jl:99-105): formed at the query times
        idx = []
        for v in ([vars] if scal[0] else list(vars)):
            if v in observed:
                idx.append(-1 - observed[v])
            elif v in names:
                idx.append(names.index(v))
            else:
                raise KeyError(f"{v!r} is neither a perturbation unknown of {self.prob.M} (available: {names[:8]} ...) nor one of the observed source functions {sorted(observed)}")
        taus = np.ascontiguousarray(np.atleast_1d(taus), dtype=np.float64)
        kq = np.atleast_1d(np.asarray(ks, dtype=np.float64))
        kmin, kmax = self.ks[0], self.ks[-1]
        if kq.min() < kmin:
            raise ValueError(f"Requested wavenumber k = {kq.min()} is below the minimum solved wavenumber {kmin}")
        if kq.max() > kmax:
            raise ValueError(f"Requested wavenumber k = {kq.max()} is above the maximum solved wavenumber {kmax}")
        U = self._states_at(taus)
        if any(i < 0 for i in idx):                   # observed source functions at (τ, k_solved): the integrator's fused evaluation, unscaled (ST, SE, Sψ)
            key = b"S" + taus.tobytes()
            if key not in self._dense:
                opts = {k: v for k, v in self._ptopts.items() if k not in ("saveat", "sources", "keep_states")}
                sg = source_grid(self.prob, taus, self.ks, self.bg, scale_k=False, lensing=True, ptivini=self._ptivini, warn=False, **opts)
                self._dense[key] = sg.dS.cpu().numpy().transpose(0, 2, 1)  # [nk][nτ][3]
            U = np.concatenate([U, self._dense[key][:, :, ::-1]], axis=2)   # (index −1 − s addresses source s)
        U = U[:, :, idx]                              # [nk_solved][nτ][nvars]
        out = np.empty((len(idx), len(taus), len(kq)))
        for ik, k in enumerate(kq):                   # neighboring_modes_indices, src/solve.jl:706-717
            if k == kmin:
                i1 = i2 = 0
            elif k == kmax:
                i1 = i2 = len(self.ks) - 1
            else:
                i2 = int(np.searchsorted(self.ks, k, side="left"))
                i1 = i2 - 1
            v = U[i1]
            if i1 != i2:
                w = (ktransform(k) - ktransform(self.ks[i1])) / (ktransform(self.ks[i2]) - ktransform(self.ks[i1]))
                v = v + (U[i2] - U[i1]) * w
            out[:, :, ik] = v.T
        out = out[0] if scal[0] else out
        if scal[1]:
            out = out[..., 0, :]
        if scal[2]:
            out = out[..., 0]
        return out


def solve(prob, ks=None, bgopts=None, ptopts=None, ptivini=-math.inf, **kw):
    """reference solve(prob, ks), src/solve.jl:380-402"""
    bg = solvebg(prob, **(bgopts or {}))
    if ks is None or len(np.atleast_1d(ks)) == 0:
        return CosmologySolution(prob, bg, None, None)
    pts = solvept(prob, bg, ks, ptivini, **(ptopts or {}), **kw)
    return CosmologySolution(prob, bg, np.atleast_1d(np.asarray(ks, dtype=np.float64)), pts, ptivini, dict(ptopts or {}, **{k: v for k, v in kw.items() if k in ("reltol", "abstol", "maxiters", "msub")}))


def issuccess(sol):
    return sol.bg.success and (sol.pts is None or sol.pts.success)


def spectrum_primordial(k, prob_or_sol):
    """P0(k) = 2π² As k⁻³ (k/kp)^(ns−1) (src/observables/fourier.jl:14-23)."""
    prob = prob_or_sol.prob if hasattr(prob_or_sol, "prob") else prob_or_sol
    k = np.asarray(k, dtype=np.float64)
    return 2 * math.pi**2 * prob.derived["As"] / k**3 * (k / prob.derived["kpivot"]) ** (prob.pars["ns"] - 1)


def spectrum_matter(prob, ks, kτini=1e-2, τinimax=1e-4, bgsol=None, return_solution=False, sourceopts=None, coarse_length=9, **kw):
    """Total matter P(k, τ0) in (c/H0)³ for ks in H0/c (reference spectrum_matter(prob, k), src/observables/fourier.jl:79-101).
    A 2-tuple `ks = (kmin, kmax)` selects the adaptive method (src/observables/fourier.jl:112-128): log-spaced coarse grid of
    `coarse_length` points refined with `sourceopts` (default atol = 4.0, rtol = 4e-3 on Δm); returns (ks, P)."""
    if isinstance(ks, tuple) and len(ks) == 2:
        kmin, kmax = float(ks[0]), float(ks[1])
        k0s = np.exp(np.linspace(math.log(kmin), math.log(kmax), coarse_length))
        k0s[0], k0s[-1] = kmin, kmax  # exp(log(k)) ≠ k with floats (fourier.jl:116-117)
        opts = dict(atol=4.0, rtol=4e-3) if sourceopts is None else dict(sourceopts)
        kk, D = source_grid_adaptive(prob, None, k0s, bgsol=bgsol, sources="matter", ktransform=(math.log, math.exp), **opts, **kw)
        return kk, spectrum_primordial(kk, prob) * D[:, 0, 0] ** 2
    ks = np.ascontiguousarray(np.atleast_1d(ks), dtype=np.float64)
    bg = bgsol if bgsol is not None else solvebg(prob)
    sol = solvept(prob, bg, ks, ptivini=lambda k: min(kτini / k, τinimax) if k > 0 else τinimax, **kw)
    d = bg.device()
    dm = torch.empty(len(ks), dtype=torch.float64, device=sol.d_uend.device)
    rc = prob.lib.sbm_delta_m(_cptr(d["P"]), C.c_int(len(bg.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_double(bg.tau0), C.c_int(len(ks)), _cptr(sol.d_ks), _cptr(sol.d_uend), _cptr(dm), _stream())
    if rc != 0:
        raise RuntimeError(f"sbm_delta_m failed with code {rc}")
    P = spectrum_primordial(ks, prob) * dm.cpu().numpy() ** 2
    return (P, sol) if return_solution else P


def refine_grid(evaluate, ks, atol=0.0, rtol=None, ktransform=None, nmaxks=1024, sort=True, skip_last_tau=True):
    """Adaptive bisection of a k-grid (the refinement rule of source_grid_adaptive, src/observables/fourier.jl:312-407):
    every interval (k1, k3) gets its midpoint k2 = f⁻¹((f(k1)+f(k3))/2) evaluated; if at any τ (the last one excluded when there
    are several: χ = 0) `isapprox(S(k2), (S(k1)+S(k3))/2; atol, rtol)` fails (2-norm over the source vector), both halves are
    refined in turn.  The reference spawns one task per midpoint; here each *level* of the bisection tree is one batched call
    `evaluate(ks_new) -> [nk][nτ][nS]` (one GPU launch over all midpoints of the level).  The decision for an interval depends
    only on its three points, so the final grid is the same tree.  Raises if more than `nmaxks` points are needed."""
    f, finv = ktransform if ktransform is not None else ((lambda x: x), (lambda x: x))
    ks = [float(k) for k in ks]
    if len(ks) < 2:
        raise ValueError("Initial k-grid must have at least 2 values")
    for kk in (min(ks), max(ks)):
        if not math.isclose(kk, finv(f(kk)), rel_tol=1e-8):
            raise ValueError("ktransform is not a tuple of inverse functions")
    if rtol is None:
        rtol = math.sqrt(np.finfo(np.float64).eps) if atol <= 0 else 0.0  # Base.isapprox default
    S = [v for v in np.asarray(evaluate(np.array(ks)), dtype=np.float64)]
    nt = S[0].shape[0]
    its = slice(0, max(nt - 1, 1)) if skip_last_tau else slice(0, nt)
    queue = [(i - 1, i) for i in range(1, len(ks))]
    while queue:
        mids = [finv((f(ks[i1]) + f(ks[i3])) / 2) for (i1, i3) in queue]
        if len(ks) + len(mids) > nmaxks:
            raise RuntimeError(f"Source function refinement needs more than {nmaxks} k-refinements. Reduce refinement criteria.")
        vals = np.asarray(evaluate(np.array(mids)), dtype=np.float64)
        nxt = []
        for (i1, i3), k2, v in zip(queue, mids, vals):
            i2 = len(ks)
            ks.append(k2)
            S.append(v)
            lin = (S[i1][its] + S[i3][its]) / 2
            err = np.linalg.norm(v[its] - lin, axis=-1)
            tol = np.maximum(atol, rtol * np.maximum(np.linalg.norm(v[its], axis=-1), np.linalg.norm(lin, axis=-1)))
            if (~(err <= tol)).any():
                nxt += [(i2, i3), (i1, i2)]
        queue = nxt
    ks, S = np.array(ks), np.stack(S)
    if sort:
        o = np.argsort(ks, kind="stable")
        ks, S = ks[o], S[o]
    return ks, S


def source_grid_adaptive(prob, taus, ks, bgsol=None, sources="cmb", atol=0.0, rtol=None, ktransform=None, sort=True, nmaxks=1024, **ptopts):
    """Adaptively refined source grid (reference source_grid_adaptive, src/observables/fourier.jl:312-416).  `sources`: "cmb" =
    (ST, SE) at the times `taus`, or "matter" = Δm at τ0 (`taus` must be None: final time only).  All modes start at the first
    background time (ptivini = −∞, src/solve.jl:533).  Returns (ks, S[nk][nτ][nS])."""
    bg = bgsol if bgsol is not None else solvebg(prob)
    if sources == "matter":
        if taus is not None:
            raise ValueError("matter sources are evaluated at the final time only")

        def evaluate(kk):
            sol = solvept(prob, bg, kk, **ptopts)
            d = bg.device()
            dm = torch.empty(len(kk), dtype=torch.float64, device=sol.d_uend.device)
            rc = prob.lib.sbm_delta_m(_cptr(d["P"]), C.c_int(len(bg.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_double(bg.tau0), C.c_int(len(kk)), _cptr(sol.d_ks), _cptr(sol.d_uend), _cptr(dm), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_delta_m failed with code {rc}")
            return dm.cpu().numpy()[:, None, None]
    elif sources == "cmb":
        if taus is None:
            raise ValueError("CMB sources need the times `taus`")

        def evaluate(kk):
            return source_grid(prob, taus, kk, bg, scale_k=False, **ptopts).dS.cpu().numpy().transpose(0, 2, 1)
    else:
        raise ValueError(f"unknown sources {sources!r}")
    return refine_grid(evaluate, ks, atol=atol, rtol=rtol, ktransform=ktransform, nmaxks=nmaxks, sort=sort)


# ---------------------------------------------------------------------------------------------- sources and k-interpolation
class ChebyshevInterpolator:
    """Chebyshev nodes (stored descending) + barycentric weights (src/observables/fourier.jl:419-459)."""

    def __init__(self, xmin, xmax, order, f=None, finv=None):
        """f / finv: monotone domain transform y = f(x) and its inverse (e.g. fk_tanh for the lensing grid, angular.jl:225-226)."""
        if not xmax > xmin:
            raise ValueError(f"Interval {(xmin, xmax)} is not sorted")
        self.f = f if f is not None else (lambda x: x)
        if f is None:
            self.xs = chebpoints(order, xmin, xmax)
            self.ys = self.xs
        else:
            if finv is None:
                raise ValueError("a transform f needs its inverse finv")
            self.ys = chebpoints(order, float(f(xmin)), float(f(xmax)))
            self.xs = np.asarray(finv(self.ys), dtype=np.float64)
            if not (np.isclose(self.xs[-1], xmin) and np.isclose(self.xs[0], xmax)):
                raise ValueError("f(x) and f⁻¹(x) are not inverses")
        self.xs[-1], self.xs[0] = xmin, xmax
        n = order
        self.ws = np.array([1.0 if j % 2 == 0 else -1.0 for j in range(n + 1)])
        self.ws[0] /= 2
        self.ws[-1] /= 2

    def minimum(self):
        return self.xs[-1]

    def maximum(self):
        return self.xs[0]

    def matrix(self, x_fine):
        """Barycentric interpolation matrix B[nfine, ncoarse] (formula of fourier.jl:524-535; exact hit returns the node value)."""
        x_fine = np.asarray(self.f(np.asarray(x_fine, dtype=np.float64)), dtype=np.float64)
        D = x_fine[:, None] - self.ys[None, :]
        hit = D == 0
        with np.errstate(divide="ignore", invalid="ignore"):
            T = self.ws[None, :] / D
            B = T / T.sum(axis=1, keepdims=True)
        rows = hit.any(axis=1)
        B[rows] = hit[rows].astype(np.float64)
        return B


def _natural_spline_matrix(y, yq):
    """B[nq, n] with (natural cubic spline through (y, f))(yq) = B f -- the spline is linear in the data (DataInterpolations.CubicSpline)."""
    y, yq = np.asarray(y, dtype=np.float64), np.asarray(yq, dtype=np.float64)
    n = len(y)
    h = np.diff(y)
    Mmap = np.zeros((n, n))  # second derivatives as a linear map of f; natural ends: M_0 = M_{n-1} = 0
    if n > 2:
        A = np.zeros((n - 2, n - 2))
        R = np.zeros((n - 2, n))
        for i in range(1, n - 1):
            A[i - 1, i - 1] = 2 * (h[i - 1] + h[i])
            if i > 1:
                A[i - 1, i - 2] = h[i - 1]
            if i < n - 2:
                A[i - 1, i] = h[i]
            R[i - 1, i - 1], R[i - 1, i], R[i - 1, i + 1] = 6 / h[i - 1], -6 / h[i - 1] - 6 / h[i], 6 / h[i]
        Mmap[1:-1] = np.linalg.solve(A, R)
    i = np.clip(np.searchsorted(y, yq, side="right") - 1, 0, n - 2)
    t1, t0, hi = y[i + 1] - yq, yq - y[i], h[i]
    B = (t1**3 / (6 * hi) - hi * t1 / 6)[:, None] * Mmap[i] + (t0**3 / (6 * hi) - hi * t0 / 6)[:, None] * Mmap[i + 1]
    rows = np.arange(len(yq))
    B[rows, i] += t1 / hi
    B[rows, i + 1] += t0 / hi
    return B


class CubicSplineInterpolator:
    """Natural cubic spline in y = f(x) through ascending nodes (reference src/observables/fourier.jl:199-231)."""

    def __init__(self, xs, xmax=None, n=None, f=None):
        if xmax is not None:  # CubicSplineInterpolator(xmin, xmax, n): n + 1 equispaced nodes
            xs = np.linspace(float(xs), float(xmax), int(n) + 1)
        self.xs = np.array(xs, dtype=np.float64)
        if (np.diff(self.xs) <= 0).any():
            raise ValueError("Input points must be sorted in ascending order")
        self.f = f if f is not None else (lambda x: x)
        self.ys = np.asarray(self.f(self.xs), dtype=np.float64)

    def minimum(self):
        return self.xs[0]

    def maximum(self):
        return self.xs[-1]

    def matrix(self, x_fine):
        return _natural_spline_matrix(self.ys, np.asarray(self.f(np.asarray(x_fine, dtype=np.float64)), dtype=np.float64))


class EquispacedInterpolator:
    """Barycentric interpolation on order + 1 equispaced nodes, weights (−1)^j binom(n, j) (reference src/observables/fourier.jl:549-571)."""

    def __init__(self, xmin, xmax, order):
        if not xmax > xmin:
            raise ValueError(f"Interval {(xmin, xmax)} is not sorted")
        self.xs = lingrid(xmin, xmax, length=order + 1)
        self.ys = self.xs
        self.ws = np.array([math.comb(order, j) * (1.0 if j % 2 == 0 else -1.0) for j in range(order + 1)])
        self.f = lambda x: x

    def minimum(self):
        return self.xs[0]

    def maximum(self):
        return self.xs[-1]

    matrix = ChebyshevInterpolator.matrix


class PiecewiseChebyshevInterpolator:
    """Chebyshev sub-grids on consecutive intervals sharing their end points (reference src/observables/fourier.jl:477-521); `xs` holds all
    unique nodes in descending order, `iranges[j]` the slice of `xs` that belongs to sub-grid j (ascending sub-grid order)."""

    def __init__(self, xbreaks, orders, f=None, finv=None):
        N = len(orders)
        if len(xbreaks) != N + 1:
            raise ValueError(f"Need {N + 1} x-breaks for {N} intervals, got {len(xbreaks)}")
        fs = list(f) if isinstance(f, (tuple, list)) else [f] * N
        finvs = list(finv) if isinstance(finv, (tuple, list)) else [finv] * N
        if len(fs) != N or len(finvs) != N:
            raise ValueError(f"Need {N} f and f⁻¹")
        self.subgrids = [ChebyshevInterpolator(xbreaks[j], xbreaks[j + 1], orders[j], f=fs[j], finv=finvs[j]) for j in range(N)]
        xs = list(self.subgrids[-1].xs)
        for j in range(N - 2, -1, -1):
            xs += list(self.subgrids[j].xs[1:])
        self.xs = np.array(xs)
        self.iranges, i = [None] * N, 0
        for j in range(N - 1, -1, -1):
            n = len(self.subgrids[j].xs)
            self.iranges[j] = slice(i, i + n)
            i += n - 1

    def minimum(self):
        return self.xs[-1]

    def maximum(self):
        return self.xs[0]

    def matrix(self, x_fine):
        x_fine = np.asarray(x_fine, dtype=np.float64)
        B = np.zeros((len(x_fine), len(self.xs)))
        for g, r in zip(self.subgrids, self.iranges):  # a point on a shared break is written twice with the same value (the node's row)
            m = (x_fine >= g.minimum()) & (x_fine <= g.maximum())
            if m.any():
                blk = np.zeros((int(m.sum()), len(self.xs)))
                blk[:, r] = g.matrix(x_fine[m])
                B[m] = blk
        return B


def fk_tanh(k, k0=2000.0):
    return np.tanh(np.asarray(k) / k0)


def fk_tanh_inv(y, k0=2000.0):
    return k0 * np.arctanh(np.asarray(y))


class SourceGrid:
    """Device-resident sources S[ik][iS][iτ] (iS: 0 = ST-like, 1 = SE-like) on (ks, τs)."""

    def __init__(self, dS, ks, taus, sol):
        self.dS, self.ks, self.taus, self.sol = dS, ks, taus, sol

    def julia_layout(self):
        """numpy array Ss[iτ, ik, iS] — the reference's Matrix{SVector{2}}(nτ, nk) (src/observables/fourier.jl:270-277)."""
        return np.ascontiguousarray(self.dS.cpu().numpy().transpose(2, 0, 1))


def source_grid(prob, taus, ks, bgsol, scale_k=True, lensing=False, fused=True, keep_states=False, **ptopts):
    """Solve the modes `ks` and evaluate the CMB sources (k·ST, k²·SE[, Sψ if lensing]) at `taus`
    (reference source_grid(prob, Ss, τs, ks, bgsol), src/observables/fourier.jl:267-281 with Ss of angular.jl:293).
    fused (default): the sources are formed inside the integrator from the dense output, as the reference's output_func does, and only
    S[nk][nS][nτ] reaches HBM (keep_states = True additionally returns the states in .sol.d_usave).  fused = False: the states
    usave[nk][nτ][N] are written and a second kernel (`sbm_sources`) evaluates the sources from them -- the same algebra (results
    agree to rounding), 40× the traffic."""
    taus = np.ascontiguousarray(taus, dtype=np.float64)
    if taus.min() < bgsol.t[0] or taus.max() > bgsol.t[-1]:
        raise ValueError("input τs and computed background solution have different timespans")
    if fused:
        sol = solvept(prob, bgsol, ks, saveat=taus, sources=dict(nS=3 if lensing else 2, scale_k=scale_k, taurec=bgsol.taurec), keep_states=keep_states, **ptopts)
        return SourceGrid(sol.d_S, sol.ks, taus, sol)
    sol = solvept(prob, bgsol, ks, saveat=taus, **ptopts)
    d = bgsol.device()
    dev = sol.d_uend.device
    nk, nt = len(sol.ks), len(taus)
    nS = 3 if lensing else 2
    dS = torch.empty((nk, nS, nt), dtype=torch.float64, device=dev)
    stride = prob.lib.sbm_srcbg_stride()
    scratch = torch.empty(nt * stride, dtype=torch.float64, device=dev)
    dtaus = torch.from_numpy(taus).to(dev)
    rc = prob.lib.sbm_sources(_cptr(d["P"]), C.c_int(len(bgsol.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_int(nt), _cptr(dtaus), _cptr(scratch), C.c_int(nk), _cptr(sol.d_ks),
                              _cptr(sol.d_usave), _cptr(dS), C.c_int(1 if scale_k else 0), C.c_int(nS), C.c_double(bgsol.taurec), _stream())
    if rc != 0:
        raise RuntimeError(f"sbm_sources failed with code {rc}")
    return SourceGrid(dS, sol.ks, taus, sol)


def source_kinterp(Sgrid, kinterp, ks_fine):
    """Barycentric interpolation coarse → fine k on the GPU (reference source_kinterp, src/observables/fourier.jl:232-247)."""
    _require_cuda()
    dev = Sgrid.dS.device
    B = torch.from_numpy(kinterp.matrix(ks_fine)).to(dev)
    nk, nc, n2t = len(ks_fine), len(kinterp.xs), Sgrid.dS.shape[1] * Sgrid.dS.shape[2]
    out = torch.empty((nk,) + tuple(Sgrid.dS.shape[1:]), dtype=torch.float64, device=dev)
    rc = los_lib().sbl_kinterp(C.c_int(nk), C.c_int(nc), _cptr(B), _cptr(Sgrid.dS), C.c_int(n2t), _cptr(out), _stream())
    if rc != 0:
        raise RuntimeError(f"sbl_kinterp failed with code {rc}")
    return SourceGrid(out, np.asarray(ks_fine), Sgrid.taus, Sgrid.sol)


# ---------------------------------------------------------------------------------------------- Bessel cache, LOS, C_l
class SphericalBesselCache:
    """Uniform-x table of j_l and j_l′, built on the GPU (reference src/observables/angular.jl:9-48).
    The x-grid is exactly the reference's (`range(0, xmax, length = trunc(xmax/dx))`, xmax = 20·l_end, dx = 2π/15), but only
    the points with x ≤ xcut are tabulated (the reference asserts x ≤ table end; values beyond k_max·τ0 are never read)."""

    def __init__(self, ls, xmax=None, dx=2 * math.pi / 15, xcut=None):
        _require_cuda()
        self.l = np.asarray(ls)
        if not np.issubdtype(self.l.dtype, np.integer):
            raise TypeError("integer multipoles only")
        if not (np.diff(self.l) > 0).all():
            raise ValueError("ls must be sorted and unique")
        xmax = 20.0 * float(self.l[-1]) if xmax is None else float(xmax)
        n = int(xmax / dx)
        self.step = xmax / (n - 1)
        self.invdx, self.dx, self.xmax = 1.0 / self.step, dx, xmax
        self.nfull = n + 1
        self.nx = self.nfull if (xcut is None or xcut >= xmax) else min(self.nfull, int(math.ceil(xcut / self.step)) + 2)
        dev = torch.device("cuda")
        self.d_l = torch.from_numpy(self.l.astype(np.int32)).to(dev)
        self.y = torch.empty((self.nx, len(self.l)), dtype=torch.float64, device=dev)
        self.dy = torch.empty_like(self.y)
        rc = los_lib().sbl_bessel_table(C.c_int(len(self.l)), _cptr(self.d_l), C.c_int(self.nx), C.c_double(self.step), _cptr(self.y), _cptr(self.dy), _stream())
        if rc != 0:
            raise RuntimeError(f"sbl_bessel_table failed with code {rc}")
        if self.nx == self.nfull:  # reference pads with a duplicate of the last point
            self.y[-1] = self.y[-2]
            self.dy[-1] = self.dy[-2]

    @property
    def xend(self):
        return (self.nx - 2) * self.step

    def __call__(self, il, x):
        """Hermite evaluation on the host (for tests); il = cache index."""
        x = np.asarray(x, dtype=np.float64)
        if (x < 0).any() or (x > self.xend).any():
            raise IndexError("x outside the cached range")
        w = x * self.invdx
        i = np.trunc(w).astype(np.int64)
        w = w - i
        wm1 = w - 1
        y, dy = self.y.cpu().numpy(), self.dy.cpu().numpy()
        return (1 + 2 * w) * wm1 * wm1 * y[i, il] + w * w * (3 - 2 * w) * y[i + 1, il] + w * wm1 * (wm1 * dy[i, il] + w * dy[i + 1, il]) * self.dx


def _trapz_weights(taus):
    w = np.empty_like(taus)
    w[0] = 0.5 * (taus[1] - taus[0])
    w[1:-1] = 0.5 * (taus[2:] - taus[:-2])
    w[-1] = 0.5 * (taus[-1] - taus[-2])
    return w


def los_integrate(Sgrid, jl, ks_fine=None, kinterp=None, k_range=None, theta=None, l_limber=None):
    """Θ_l(k) = Σ_τ w_τ S(τ,k) j_l(k(τ0−τ)) with Θ_T/k and Θ_E √((l+2)!/(l−2)!)/k² rescaling
    (reference los_integrate + rescale, src/observables/angular.jl:109-152, 301-306).
    Sgrid holds (k·ST, k²·SE[, Sψ]).  With `kinterp`, Sgrid is on the coarse nodes and the interpolation to ks_fine is fused.
    For a third (lensing) source the Limber approximation is used for l ≥ l_limber (angular.jl:155-178, 307-309; default: never).
    Returns device tensor Theta[nS][nl][nk_fine]."""
    _require_cuda()
    taus = Sgrid.taus
    ks_fine = Sgrid.ks if ks_fine is None else np.ascontiguousarray(ks_fine, dtype=np.float64)
    if not taus[1] > taus[0]:
        raise ValueError("τs must be sorted in ascending order")
    if len(ks_fine) > 1 and not ks_fine[1] > ks_fine[0]:
        raise ValueError("ks must be sorted in ascending order")
    if jl.xend < ks_fine[-1] * (taus[-1] - taus[0]):
        raise ValueError("jl.x[end] < kmax*τmax")
    dev = Sgrid.dS.device
    nk, nt, nl = len(ks_fine), len(taus), len(jl.l)
    chi = torch.from_numpy(taus[-1] - taus).to(dev)
    wt = torch.from_numpy(_trapz_weights(taus)).to(dev)
    dks = torch.from_numpy(ks_fine).to(dev)
    if kinterp is not None:
        Bw = torch.from_numpy(kinterp.matrix(ks_fine)).to(dev)
        nc = len(kinterp.xs)
    else:
        Bw, nc = None, nk
    nS = Sgrid.dS.shape[1]
    if theta is None:
        theta = torch.zeros((nS, nl, nk), dtype=torch.float64, device=dev)
    k_lo, k_hi = (0, nk) if k_range is None else k_range
    rc = los_lib().sbl_los(C.c_int(k_hi - k_lo), C.c_int(k_lo), C.c_int(nk), _cptr(dks), C.c_int(nc), _cptr(Bw), _cptr(Sgrid.dS), C.c_int(nS), C.c_int(nt), _cptr(chi), _cptr(wt), C.c_int(nl), _cptr(jl.d_l),
                           _cptr(jl.y), _cptr(jl.dy), C.c_double(jl.invdx), C.c_double(jl.dx), C.c_int(jl.nx), _cptr(theta), C.c_int(2**31 - 1 if l_limber is None else int(l_limber)), _stream())
    if rc != 0:
        raise RuntimeError(f"sbl_los failed with code {rc}")
    return theta


def natural_spline_weights(x):
    """w with ∫ natural-cubic-spline(x, f) dx = w·f (DataInterpolations CubicSpline + integral, angular.jl:212-213).
    The spline integral is linear in f: trapezoid − (1/24) Σ h_i³ (M_i + M_{i+1}), M = T⁻¹ R f."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    h = np.diff(x)
    w = np.zeros(n)
    w[:-1] += h / 2
    w[1:] += h / 2
    if n > 2:
        m = n - 2  # interior second derivatives M_1..M_{n-2}
        c = (h[:-1] ** 3 + h[1:] ** 3) / 24.0
        # solve T z = c with T tridiagonal (sub h[1:m], diag 2(h[i-1]+h[i]), sup h[1:m]) -- Thomas algorithm
        dg = 2 * (h[:-1] + h[1:])
        sub = h[1:-1].copy()
        cp, dp = np.zeros(m), np.zeros(m)
        cp[0] = (sub[0] / dg[0]) if m > 1 else 0.0
        dp[0] = c[0] / dg[0]
        for i in range(1, m):
            den = dg[i] - sub[i - 1] * cp[i - 1]
            cp[i] = (sub[i] / den) if i < m - 1 else 0.0
            dp[i] = (c[i] - sub[i - 1] * dp[i - 1]) / den
        z = np.zeros(m)
        z[-1] = dp[-1]
        for i in range(m - 2, -1, -1):
            z[i] = dp[i] - cp[i] * z[i + 1]
        # r_i = 6[(f_{i+1}-f_i)/h_i − (f_i−f_{i-1})/h_{i-1}]  (i = 1..n-2)
        w[0:m] -= 6 * z / h[:-1]
        w[1:m + 1] += 6 * z * (1 / h[:-1] + 1 / h[1:])
        w[2:m + 2] -= 6 * z / h[1:]
    return w


_MODE_IDX = {"T": 0, "E": 1, "ψ": 2, "P": 2}  # P: ASCII alias of ψ (lensing potential)


def spectrum_cmb_from_theta(theta, modes, P0s, ls, ks, normalization="Cl", k_mask=None):
    """C_l^{AB} = (2/π)∫dk k² P0 Θ^A Θ^B via the natural cubic spline through (0,0)+ks (reference src/observables/angular.jl:198-223).
    theta: device tensor [2][nl][nk].  k_mask selects the k owned by this rank (partial sums for multi-GPU). Returns device tensor [nmodes][nl]."""
    ks = np.asarray(ks, dtype=np.float64)
    w = natural_spline_weights(np.concatenate([[0.0], ks]))[1:]
    ck = w * (2 / math.pi) * ks**2 * P0s
    if k_mask is not None:
        ck = ck * k_mask
    dev = theta.device
    dck = torch.from_numpy(ck).to(dev)
    mA = torch.tensor([_MODE_IDX[m[0]] for m in modes], dtype=torch.int32, device=dev)
    mB = torch.tensor([_MODE_IDX[m[1]] for m in modes], dtype=torch.int32, device=dev)
    nl, nk = theta.shape[1], theta.shape[2]
    Cl = torch.empty((len(modes), nl), dtype=torch.float64, device=dev)
    rc = los_lib().sbl_cl(C.c_int(nl), C.c_int(nk), C.c_int(0), C.c_int(nk), _cptr(dck), _cptr(theta), C.c_int(len(modes)), _cptr(mA), _cptr(mB), _cptr(Cl), _stream())
    if rc != 0:
        raise RuntimeError(f"sbl_cl failed with code {rc}")
    if normalization == "Dl":
        lsd = torch.from_numpy(np.asarray(ls, dtype=np.float64)).to(dev)
        Cl = Cl * (lsd * (lsd + 1) / (2 * math.pi))[None, :]
    elif normalization != "Cl":
        raise ValueError(f"Normalization {normalization} is not Cl or Dl")
    return Cl


def cmb_grids(bgsol, kmin=1e-2, kmax=2e3, dkt0=math.pi, ntau=300, taucut=1e-2):
    """k and τ grids of spectrum_cmb (reference src/observables/angular.jl:275-290)."""
    ks_fine = lingrid(kmin, kmax, step=dkt0 / bgsol.tau0)
    ts = bgsol.t[bgsol.t >= taucut]
    taus = ts[0] + (ts[-1] - ts[0]) * cosgrid(0.0, 1.0, length=ntau)
    taus[-1] = ts[-1]
    return ks_fine, taus


def spline_ls(spectra_coarse, ls_coarse, ls_fine):
    """coarse-l → all-l natural cubic spline (reference src/observables/angular.jl:348-359). spectra_coarse: [nl, nmodes]."""
    x = np.asarray(ls_coarse, dtype=np.float64)
    xf = np.asarray(ls_fine, dtype=np.float64)
    n = len(x)
    h = np.diff(x)
    out = np.zeros((len(xf), spectra_coarse.shape[1]))
    for j in range(spectra_coarse.shape[1]):
        f = spectra_coarse[:, j]
        M = np.zeros(n)
        if n > 2:
            A = np.zeros((n - 2, n - 2))
            r = 6 * ((f[2:] - f[1:-1]) / h[1:] - (f[1:-1] - f[:-2]) / h[:-1])
            idx = np.arange(n - 2)
            A[idx, idx] = 2 * (h[:-1] + h[1:])
            A[idx[:-1], idx[:-1] + 1] = h[1:-1]
            A[idx[1:], idx[1:] - 1] = h[1:-1]
            M[1:-1] = np.linalg.solve(A, r)
        i = np.clip(np.searchsorted(x, xf, side="right") - 1, 0, n - 2)
        t1, t0 = x[i + 1] - xf, xf - x[i]
        out[:, j] = (M[i] * t1**3 + M[i + 1] * t0**3) / (6 * h[i]) + (f[i] / h[i] - M[i] * h[i] / 6) * t1 + (f[i + 1] / h[i] - M[i + 1] * h[i] / 6) * t0
    return out


def spectrum_cmb(modes, prob, jl, ls=None, normalization="Cl", kinterp=None, direct=False, dkt0=math.pi, ntau=300, taucut=1e-2, l_limber=10,
                 bgsol=None, ptopts=None, group=None, return_all=False):
    """Angular spectra C_l^{AB}, A, B ∈ {T, E, ψ} at jl.l (optionally splined to `ls`)
    (reference spectrum_cmb(modes, prob, jl[, ls]), src/observables/angular.jl:260-359).  With a lensing mode (ψ, alias P) the default
    k-grid is the tanh-stretched order-130 Chebyshev grid to k = 1e4 and the ψ line-of-sight integral uses Limber for l ≥ l_limber.
    direct=True solves every fine k instead of interpolating from the Chebyshev nodes.
    With torch.distributed initialised (group or default), or with `group` a library `Communicator` (NCCL behind the C ABI), the modes are
    sharded over ranks (strided), the sources all-gathered, the line of sight run on contiguous fine-k slices and the partial C_l all-reduced."""
    modes = [modes] if isinstance(modes, str) else list(modes)
    for m in modes:
        if len(m) != 2 or m[0] not in _MODE_IDX or m[1] not in _MODE_IDX:
            raise ValueError(f"Unknown CMB power spectrum mode {m}")
    world, rank = _ranks(group)
    bg = bgsol if bgsol is not None else solvebg(prob)
    lensing = any(c in ("ψ", "P") for m in modes for c in m)
    if kinterp is None:  # angular.jl:267-273
        kinterp = ChebyshevInterpolator(1e-2, 1e4, 130, f=fk_tanh, finv=fk_tanh_inv) if lensing else ChebyshevInterpolator(1e-2, 2e3, 60)
    ks_fine, taus = cmb_grids(bg, kinterp.minimum(), kinterp.maximum(), dkt0, ntau, taucut)
    ptopts = dict(ptopts or {})
    nkf = len(ks_fine)
    ks_solve = ks_fine if direct else kinterp.xs
    # shard the ODE solves over ranks, strided in k so that cost (∝ k) balances; no exchange during the solve
    mine = np.arange(rank, len(ks_solve), world)
    S = source_grid(prob, taus, ks_solve[mine], bg, lensing=lensing, **ptopts)
    if world > 1:
        full = torch.zeros((len(ks_solve), S.dS.shape[1], len(taus)), dtype=torch.float64, device=S.dS.device)
        S.dS[:, :, -1] = 0  # χ = 0 row (Inf/NaN in SE, Sψ) is dropped by the LOS kernel anyway; keep the reduction finite
        full[torch.from_numpy(mine).to(S.dS.device)] = S.dS
        _allreduce(full, group)  # gather of the sources: supports are disjoint, so a sum is an all-gather
        S = SourceGrid(full, ks_solve, taus, S.sol)
    lo, hi = (nkf * rank) // world, (nkf * (rank + 1)) // world  # contiguous fine-k slice for LOS + partial C_l
    theta = los_integrate(S, jl, ks_fine=ks_fine, kinterp=None if direct else kinterp, k_range=(lo, hi), l_limber=l_limber if lensing else None)
    mask = np.zeros(nkf)
    mask[lo:hi] = 1.0
    P0s = spectrum_primordial(ks_fine, prob)
    Cl = spectrum_cmb_from_theta(theta, modes, P0s, jl.l, ks_fine, normalization, k_mask=None if world == 1 else mask)
    if world > 1:
        _allreduce(Cl, group)  # partial k-sums → full C_l (north star: NCCL all-reduce of the partial C_l sums)
    out = Cl.cpu().numpy().T.copy()  # [nl, nmodes] like the reference's spectra[il, imode]
    if ls is not None:
        if (min(ls), max(ls)) != (jl.l[0], jl.l[-1]):
            raise ValueError("jl.l and ls have different extrema")
        out = spline_ls(out, jl.l, ls)
    if return_all:
        return out, dict(ks_fine=ks_fine, taus=taus, S=S, theta=theta, bg=bg)
    return out


class CMBPlan:
    """Preallocated, allocation-free execution plan for the C_l hot path of one cosmology:
    (H2D knots) -> β-table -> perturbation solve (saveat) -> sources -> [k-interp +] LOS -> C_l -> (D2H C_l).
    Used by bench.py for the device-resident number (`run`) and the end-to-end number (`run_e2e`).  `stage(bg)` re-targets the plan
    at another cosmology with the same grid sizes (knots, fine k, times) -- it restages every cosmology-dependent array, not only
    the knots; `upload()` must run once before the first `run()`.  Every launch goes to torch's current stream."""

    def __init__(self, prob, bg, jl, modes=("TT", "EE", "TE"), direct=True, kinterp=None, dkt0=math.pi, ntau=300, taucut=1e-2,
                 reltol=1e-5, abstol=1e-5, maxiters=100000, msub=16, normalization="Cl", fused=True, split=True):
        _require_cuda()
        self.prob, self.jl, self.modes, self.direct, self.fused, self.split = prob, jl, list(modes), direct, bool(fused), bool(split)
        self.reltol, self.abstol, self.maxiters, self.msub, self.normalization = reltol, abstol, maxiters, msub, normalization
        self.kinterp = kinterp if kinterp is not None else ChebyshevInterpolator(1e-2, 2e3, 60)
        dev = torch.device("cuda")
        self.dev = dev
        self._grid_opts = (dkt0, ntau, taucut)
        ks_fine, taus = cmb_grids(bg, self.kinterp.minimum(), self.kinterp.maximum(), dkt0, ntau, taucut)
        nkf, nt, N, nl = len(ks_fine), len(taus), prob.N, len(jl.l)
        nk = nkf if direct else len(self.kinterp.xs)
        self.nk, self.nkf, self.nt, self.nl = nk, nkf, nt, nl
        self.nb = len(bg.t)
        # pinned host staging for the per-cosmology inputs: knots (t, y, dy), parameters, and the grids that move with the cosmology
        # (fine k: step π/τ0; saved times and χ = τ0 − τ; trapezoid weights; C_l k-weights incl. P0; start times; interval look-up)
        self.nlut = 4096
        self._seg = dict(t=self.nb, y=5 * self.nb, dy=5 * self.nb, P=prob.npar, ks=nk, tini=nk, ksf=nkf, taus=nt, chi=nt, wt=nt, ck=nkf)
        self._off, o = {}, 0
        for name, n in self._seg.items():
            self._off[name] = (o, o + n)
            o += n
        self.h_in = torch.empty(o, dtype=torch.float64).pin_memory()
        self.d_in = torch.empty_like(self.h_in, device=dev)
        self.h_lut = torch.empty(self.nlut, dtype=torch.int32).pin_memory()
        self.d_lut = torch.empty(self.nlut, dtype=torch.int32, device=dev)
        dv = lambda name: self.d_in[self._off[name][0]:self._off[name][1]]
        self.d_ks, self.d_tini, self.d_ksf, self.d_taus, self.d_chi, self.d_wt, self.d_ck = dv("ks"), dv("tini"), dv("ksf"), dv("taus"), dv("chi"), dv("wt"), dv("ck")
        self.d_Bw = None if direct else torch.empty((nkf, nk), dtype=torch.float64, device=dev)
        self.stage(bg)
        order = np.argsort(-self.ks_solve, kind="stable").astype(np.int32)  # ascending grids of fixed length: the order does not depend on the cosmology
        f64 = dict(dtype=torch.float64, device=dev)
        self.d_order = torch.from_numpy(order).to(dev)
        self.d_usave = None if self.fused else torch.empty((nk, nt, N), **f64)  # fused: the saved states never reach HBM
        self.d_uend = torch.empty((nk, N), **f64)
        self.d_ret = torch.empty(nk, dtype=torch.int32, device=dev)
        self.d_stats = torch.empty((nk, 4), dtype=torch.int64, device=dev)
        self.d_queue = torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_S = torch.empty((nk, 2, nt), **f64)
        self.d_srcbg = torch.empty(nt * prob.lib.sbm_srcbg_stride(), **f64)
        self.d_theta = torch.zeros((2, nl, nkf), **f64)
        self.d_mA = torch.tensor([_MODE_IDX[m[0]] for m in self.modes], dtype=torch.int32, device=dev)
        self.d_mB = torch.tensor([_MODE_IDX[m[1]] for m in self.modes], dtype=torch.int32, device=dev)
        self.d_Cl = torch.empty((len(self.modes), nl), **f64)
        self.h_Cl = torch.empty((len(self.modes), nl), dtype=torch.float64).pin_memory()
        nnode = (self.nb - 1) * msub + 1
        self.d_tab = torch.empty((nnode, 2, prob.NBETA), **f64)
        self.h2d_bytes = self.h_in.numel() * 8 + self.h_lut.numel() * 4
        self.d2h_bytes = self.h_Cl.numel() * 8
        # algorithmic flops of one source evaluation (SURVEY §8d F_S): u̇ = J u (2 flop per stored Jacobian entry and factor), hub terms, Ψ̇, Π̈ rows, closing algebra
        self.src_flops = 3 * prob.info["nnz_full"] + 120
        self._src = SbmSrc(self.d_srcbg.data_ptr(), self.d_S.data_ptr(), 2, 1, 0.0)
        self.launches_resident, self.launches_e2e = (3, 5) if self.fused else (5, 6)
        self.cost_model, self.d_items = None, None

    def learn_schedule(self, model=None, save_cost=None):
        """Switch the perturbation launch from the atomic queue to the static preemptive schedule (`build_schedule`), with the
        per-mode cost taken from `model` (a ModeCostModel) or learnt from the step counters of this plan's last solve
        (blocking read of 64 KB).  Returns the model so that other plans / later cosmologies can reuse it.
        save_cost: cost of one save point (dense output + fused source evaluation) in units of a Rosenbrock attempt; the same for every
        mode, so it matters for lists that hold many short modes."""
        save_cost = (0.2 if self.fused else 0.05) if save_cost is None else save_cost  # measured: scripts/ab_fused.py (profiles/integrate_r2.md)
        lib = self.prob.lib
        if model is None:
            st = self.d_stats.cpu().numpy()
            model = ModeCostModel(self.ks_solve, st[:, 0] + st[:, 1])
        self.cost_model = model
        wpc = int(lib.sbm_warps_per_cta())
        self.nlists = max(wpc, min(int(lib.sbm_resident_warps()), self.nk) // wpc * wpc)
        items, ibeg, self.sched_T = build_schedule(model(self.ks_solve) + save_cost * self.nt, self.nlists)
        self.d_items, self.d_ibeg = torch.from_numpy(items).to(self.dev), torch.from_numpy(ibeg).to(self.dev)
        self.d_cont = torch.empty(self.nk * int(lib.sbm_cont_stride()), dtype=torch.float64, device=self.dev)
        self.d_flags = torch.zeros(self.nk, dtype=torch.int32, device=self.dev)
        return model

    def stage(self, bg, prob=None):
        """Stage a cosmology (a background solution with the plan's number of knots and grid sizes) into the pinned buffers:
        knots and parameters AND everything derived from them -- the fine k-grid (step π/τ0), the saved times, χ = τ0 − τ, the
        trapezoid weights, the C_l k-weights (P0 depends on As, ns, h), the start times and the interval look-up of the β-table.
        `upload()` then moves all of it in two copies.  Raises if the grid sizes differ from the plan's (build a new plan)."""
        prob = prob if prob is not None else bg.prob
        if len(bg.t) != self.nb:
            raise ValueError("plan was built for a different number of background knots")
        dkt0, ntau, taucut = self._grid_opts
        ks_fine, taus = cmb_grids(bg, self.kinterp.minimum(), self.kinterp.maximum(), dkt0, ntau, taucut)
        if len(ks_fine) != self.nkf or len(taus) != self.nt:
            raise ValueError(f"plan was built for {self.nkf} fine wavenumbers and {self.nt} times; this cosmology needs {len(ks_fine)} and {len(taus)}")
        if self.jl.xend < ks_fine[-1] * (taus[-1] - taus[0]):
            raise ValueError("jl.x[end] < kmax*τmax")  # reference assertion (src/observables/angular.jl:110-116); the kernel would clamp silently
        self.ks_fine, self.taus = ks_fine, taus
        self.ks_solve = ks_fine if self.direct else self.kinterp.xs
        w = natural_spline_weights(np.concatenate([[0.0], ks_fine]))[1:]
        vals = dict(t=bg.t, y=bg.y.ravel(), dy=bg.dy.ravel(), P=bg.P, ks=self.ks_solve, tini=np.full(self.nk, bg.t[0]), ksf=ks_fine, taus=taus, chi=taus[-1] - taus,
                    wt=_trapz_weights(taus), ck=w * (2 / math.pi) * ks_fine**2 * spectrum_primordial(ks_fine, prob))
        h = self.h_in.numpy()
        for name, v in vals.items():
            a, b = self._off[name]
            h[a:b] = v
        self.tau0, self.s0, self.dsl = bg.tau0, math.log(bg.t[0]), (math.log(bg.t[-1]) - math.log(bg.t[0])) / self.nlut
        self.h_lut.numpy()[:] = np.clip(np.searchsorted(bg.t, np.exp(self.s0 + self.dsl * np.arange(self.nlut)), side="right") - 1, 0, self.nb - 2).astype(np.int32)
        if self.d_Bw is not None:  # barycentric weights coarse -> fine k (the fine grid moves with τ0)
            self.d_Bw.copy_(torch.from_numpy(self.kinterp.matrix(ks_fine)))
        self._bg = bg

    def _views(self):
        v = lambda name: self.d_in[self._off[name][0]:self._off[name][1]]
        return v("P"), v("t"), v("y"), v("dy")

    def upload(self):
        """H2D of the staged inputs + β-table build."""
        self.d_in.copy_(self.h_in, non_blocking=True)
        self.d_lut.copy_(self.h_lut, non_blocking=True)
        P, t, y, dy = self._views()
        rc = self.prob.lib.sbm_build_table(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.msub), _cptr(self.d_tab), _stream())
        if rc != 0:
            raise RuntimeError(f"sbm_build_table failed with code {rc}")
        if self.fused:  # per-save-time background of the source evaluation (the unfused path builds it inside sbm_sources)
            rc = self.prob.lib.sbm_srcbg(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.nt), _cptr(self.d_taus), _cptr(self.d_srcbg), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_srcbg failed with code {rc}")

    def solve(self):
        """Perturbation solve of all modes up to the sources S(τ,k) on the device."""
        self._integrate()
        if not self.fused:
            self.sources()

    def _integrate(self):
        P, t, y, dy = self._views()
        lib = self.prob.lib
        srcp = C.byref(self._src) if self.fused else None
        if self.d_items is not None:
            rc = lib.sbm_solvept_sched_src(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.msub), C.c_int(self.nlut), C.c_double(self.s0), C.c_double(self.dsl), _cptr(self.d_lut), _cptr(self.d_tab),
                                           C.c_int(self.nk), _cptr(self.d_ks), _cptr(self.d_tini), C.c_double(self.tau0), C.c_int(self.nt), _cptr(self.d_taus), C.c_double(self.reltol), C.c_double(self.abstol),
                                           C.c_int(self.maxiters), _cptr(self.d_usave), _cptr(self.d_uend), _cptr(self.d_ret), _cptr(self.d_stats), _cptr(self.d_queue), _cptr(self.d_items), _cptr(self.d_ibeg),
                                           C.c_int(self.nlists), _cptr(self.d_cont), _cptr(self.d_flags), _stream(), srcp)
            if rc < 0:
                raise RuntimeError(f"sbm_solvept_sched_src failed with code {rc}")
            return
        if self.split and split_pays(self.prob, self.ks_solve):  # few modes (the default 61-node path): one CTA of SB_R warps per mode
            rc = lib.sbm_solvept_split(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.msub), C.c_int(self.nlut), C.c_double(self.s0), C.c_double(self.dsl), _cptr(self.d_lut), _cptr(self.d_tab),
                                       C.c_int(self.nk), _cptr(self.d_ks), _cptr(self.d_tini), _cptr(self.d_order), C.c_double(self.tau0), C.c_int(self.nt), _cptr(self.d_taus), C.c_double(self.reltol), C.c_double(self.abstol),
                                       C.c_int(self.maxiters), _cptr(self.d_usave), _cptr(self.d_uend), _cptr(self.d_ret), _cptr(self.d_stats), _cptr(self.d_queue), _stream(), srcp)
            if rc < 0:
                raise RuntimeError(f"sbm_solvept_split failed with code {rc}")
            return
        rc = lib.sbm_solvept_src(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.msub), C.c_int(self.nlut), C.c_double(self.s0), C.c_double(self.dsl), _cptr(self.d_lut), _cptr(self.d_tab),
                                 C.c_int(self.nk), _cptr(self.d_ks), _cptr(self.d_tini), _cptr(self.d_order), C.c_double(self.tau0), C.c_int(self.nt), _cptr(self.d_taus), C.c_double(self.reltol), C.c_double(self.abstol),
                                 C.c_int(self.maxiters), _cptr(self.d_usave), _cptr(self.d_uend), _cptr(self.d_ret), _cptr(self.d_stats), _cptr(self.d_queue), C.c_int(0), _stream(), srcp)
        if rc < 0:
            raise RuntimeError(f"sbm_solvept_src failed with code {rc}")

    def sources(self):
        P, t, y, dy = self._views()
        rc = self.prob.lib.sbm_sources(_cptr(P), C.c_int(self.nb), _cptr(t), _cptr(y), _cptr(dy), C.c_int(self.nt), _cptr(self.d_taus), _cptr(self.d_srcbg), C.c_int(self.nk), _cptr(self.d_ks),
                                       _cptr(self.d_usave), _cptr(self.d_S), C.c_int(1), C.c_int(2), C.c_double(0.0), _stream())
        if rc != 0:
            raise RuntimeError(f"sbm_sources failed with code {rc}")

    def los_cl(self):
        jl, L = self.jl, los_lib()
        rc = L.sbl_los(C.c_int(self.nkf), C.c_int(0), C.c_int(self.nkf), _cptr(self.d_ksf), C.c_int(self.nk), _cptr(self.d_Bw), _cptr(self.d_S), C.c_int(2), C.c_int(self.nt), _cptr(self.d_chi), _cptr(self.d_wt),
                       C.c_int(self.nl), _cptr(jl.d_l), _cptr(jl.y), _cptr(jl.dy), C.c_double(jl.invdx), C.c_double(jl.dx), C.c_int(jl.nx), _cptr(self.d_theta), C.c_int(2**31 - 1), _stream())
        if rc != 0:
            raise RuntimeError(f"sbl_los failed with code {rc}")
        rc = L.sbl_cl(C.c_int(self.nl), C.c_int(self.nkf), C.c_int(0), C.c_int(self.nkf), _cptr(self.d_ck), _cptr(self.d_theta), C.c_int(len(self.modes)), _cptr(self.d_mA), _cptr(self.d_mB), _cptr(self.d_Cl), _stream())
        if rc != 0:
            raise RuntimeError(f"sbl_cl failed with code {rc}")

    def run(self):
        """Device-resident pass: solve (-> sources) -> LOS -> C_l (inputs and β-table already in HBM)."""
        self.solve()
        self.los_cl()

    def download(self):
        self.h_Cl.copy_(self.d_Cl, non_blocking=False)
        out = self.h_Cl.numpy().T.copy()
        if self.normalization == "Dl":
            ls = np.asarray(self.jl.l, dtype=np.float64)
            out = out * (ls * (ls + 1) / (2 * math.pi))[:, None]
        return out

    def run_e2e(self):
        """End-to-end pass from host buffers: H2D (pinned) -> table -> solve -> sources -> LOS -> C_l -> D2H."""
        self.upload()
        self.run()
        return self.download()


def shard_rows(n, rank, world):
    """Rows (cosmologies) owned by `rank`: strided, so that a sorted or structured parameter list is spread evenly (batch
    configs shard by cosmology, no exchange until the end)."""
    return np.arange(rank, n, world)


def gather_rows(local, mine, n, group=None):
    """All ranks obtain the full [n, ...] array from the rows each of them owns.  The supports are disjoint, so a sum all-reduce is an
    exact gather (x + 0 = x; NaN rows of failed cosmologies stay NaN).  NCCL needs device tensors, gloo takes host tensors."""
    import torch.distributed as dist
    local = np.asarray(local, dtype=np.float64)
    full = torch.zeros((n,) + local.shape[1:], dtype=torch.float64)
    full[torch.from_numpy(np.asarray(mine, dtype=np.int64))] = torch.from_numpy(local)
    if isinstance(group, Communicator) or dist.get_backend(group) == "nccl":
        full = full.cuda()
    _allreduce(full, group)
    return full.cpu().numpy()


def spectrum_matter_sweep(prob, names, thetas, ks, chunk=32, nthreads=None, kτini=1e-2, τinimax=1e-4, reltol=1e-5, abstol=1e-5, return_info=False, cost=None, msub=16, nslots=2,
                          background="host", group=None):
    """P(k) for a batch of cosmologies θ ↦ parameter_updater(prob, names)(θ) (BASELINE config 4: emulator / MCMC sweeps;
    the reference runs a serial outer loop of `spectrum_matter(probgen(θ), ks)`, docs/src/forecasting.md:56-59).
    Host background solves run on a thread pool (the ctypes calls release the GIL).  The perturbation solves of `chunk` cosmologies
    go into ONE integrator launch over all their (cosmology, mode) pairs (`solvept_batch`), so the resident warps are kept busy by
    a single descending-k queue instead of one short launch per cosmology; consecutive chunks alternate between two CUDA streams
    (uploads and table builds of chunk c+1 overlap the solve of chunk c, whose tail is filled by the next launch).
    background = "device": all background solves run first in one `solvebg_batch` launch (one thread per cosmology) instead of the
    host thread pool -- for boxes with few host cores per GPU; P(k) then agrees with the host-background result to the background
    tolerance (≈1e-6), not bit for bit.
    With torch.distributed initialised (`group` or the default group, one process per GPU) the cosmologies are sharded over the ranks
    (`shard_rows`), every rank sweeps its share with no exchange, and one all-reduce gathers P(k) (`gather_rows`; BASELINE config 4:
    "4096 cosmologies × 256 k-modes sharded across 8 GPUs"); the failure counts of `info` are summed over ranks.
    thetas: [ncosmo, len(names)].  Returns P[ncosmo, nk] (NaN rows where the background failed); bit-identical to single calls."""
    import concurrent.futures as cf
    import os
    import torch.distributed as dist
    _require_cuda()
    thetas = np.asarray(thetas, dtype=np.float64)
    thetas = thetas.reshape(0, len(names)) if thetas.size == 0 else np.atleast_2d(thetas)
    ks = np.ascontiguousarray(ks, dtype=np.float64)
    if _ranks(group)[0] > 1:
        world, rank = _ranks(group)
        mine = shard_rows(len(thetas), rank, world)
        kw = dict(chunk=chunk, nthreads=nthreads or max(1, (os.cpu_count() or 1) // world), kτini=kτini, τinimax=τinimax, reltol=reltol, abstol=abstol, cost=cost, msub=msub,
                  nslots=nslots, background=background)
        if len(mine):
            Pm, info = _spectrum_matter_sweep_local(prob, names, thetas[mine], ks, return_info=True, **kw)
        else:
            Pm, info = np.zeros((0, len(ks))), dict(background_failures=0, mode_failures=0, launches=0)
        out = gather_rows(Pm, mine, len(thetas), group)
        if return_info:
            cnt = gather_rows(np.array([[info["background_failures"], info["mode_failures"], info["launches"]]], dtype=np.float64), [rank], world, group).sum(axis=0)
            return out, dict(background_failures=int(cnt[0]), mode_failures=int(cnt[1]), launches=int(cnt[2]), ranks=world)
        return out
    return _spectrum_matter_sweep_local(prob, names, thetas, ks, chunk=chunk, nthreads=nthreads, kτini=kτini, τinimax=τinimax, reltol=reltol, abstol=abstol, return_info=return_info,
                                        cost=cost, msub=msub, nslots=nslots, background=background)


def _spectrum_matter_sweep_local(prob, names, thetas, ks, chunk=32, nthreads=None, kτini=1e-2, τinimax=1e-4, reltol=1e-5, abstol=1e-5, return_info=False, cost=None, msub=16, nslots=2,
                                 background="host"):
    """One rank's share of `spectrum_matter_sweep`."""
    import concurrent.futures as cf
    import os
    upd = parameter_updater(prob, names)
    nthreads = nthreads or os.cpu_count()
    n, nk = len(thetas), len(ks)

    def host(theta):
        p = upd(theta)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return p, solvebg(p)

    streams = [torch.cuda.Stream() for _ in range(nslots)]
    # up to `nslots` launches are in flight at once; a statically scheduled launch needs ALL its lists resident from the start
    # (continuation items wait for a flag set by the first item of another list), so each gets an equal share of the resident warps
    max_lists = None if cost is None else max(1, resident_warps(prob, batch=True) // nslots)
    out = np.full((n, nk), np.nan)
    info = dict(background_failures=0, mode_failures=0, launches=0)
    timeline, tstart = [], time.perf_counter()
    f = lambda k: min(kτini / k, τinimax) if k > 0 else τinimax
    inflight = [None] * nslots
    arenas = [CosmoArena() for _ in range(nslots)]

    def finish(slot):
        job = inflight[slot]
        if job is None:
            return
        idx, probs, bgs, batch, dm, h_dm, h_rc, ev = job
        t_ = time.perf_counter()
        ev.synchronize()
        timeline.append(("wait", slot, t_ - tstart, time.perf_counter() - tstart))
        dmh = h_dm.numpy().reshape(len(idx), nk)
        for j, i in enumerate(idx):
            out[i] = spectrum_primordial(ks, probs[j]) * dmh[j] ** 2
        info["mode_failures"] += int((h_rc.numpy() != 0).sum())  # read from the pinned copy: a kernel on another stream would queue behind the persistent CTAs
        inflight[slot] = None

    if background not in ("host", "device"):
        raise ValueError("background must be 'host' or 'device'")
    with cf.ThreadPoolExecutor(nthreads) as pool:
        if background == "device":
            dprobs = [upd(t) for t in thetas]
            results = iter(zip(dprobs, solvebg_batch(dprobs, warn=False)))
        else:
            results = pool.map(host, thetas)
        for c, c0 in enumerate(range(0, n, chunk)):
            slot = c % nslots
            finish(slot)
            t_ = time.perf_counter()
            group = [next(results) for _ in range(min(chunk, n - c0))]
            timeline.append(("gather", slot, t_ - tstart, time.perf_counter() - tstart))
            t_ = time.perf_counter()
            idx = [c0 + j for j, (p, bg) in enumerate(group) if bg.success]
            info["background_failures"] += len(group) - len(idx)
            if not idx:
                continue
            probs, bgs = [group[i - c0][0] for i in idx], [group[i - c0][1] for i in idx]
            with torch.cuda.stream(streams[slot]):
                batch = solvept_batch(bgs, ks, ptivini=f, reltol=reltol, abstol=abstol, msub=msub, cost=cost, arena=arenas[slot], max_lists=max_lists)
                dm = torch.empty(len(idx) * nk, dtype=torch.float64, device=batch.d_uend.device)
                for j, (p, bg, sol) in enumerate(zip(probs, bgs, batch.sols)):
                    d = arenas[slot].views[j]
                    rc = p.lib.sbm_delta_m(_cptr(d["P"]), C.c_int(len(bg.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_double(bg.tau0), C.c_int(nk), _cptr(sol.d_ks), _cptr(sol.d_uend),
                                           _cptr(dm[j * nk:(j + 1) * nk]), _stream())
                    if rc != 0:
                        raise RuntimeError(f"sbm_delta_m failed with code {rc}")
                h_dm = torch.empty(len(idx) * nk, dtype=torch.float64, pin_memory=True)
                h_dm.copy_(dm, non_blocking=True)
                h_rc = torch.empty(len(idx) * nk, dtype=torch.int32, pin_memory=True)
                h_rc.copy_(batch.d_retcode, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            info["launches"] += 1
            inflight[slot] = (idx, probs, bgs, batch, dm, h_dm, h_rc, ev)
            timeline.append(("launch", slot, t_ - tstart, time.perf_counter() - tstart))
    for slot in range(nslots):  # finish() is a no-op for an empty slot
        finish(slot)
    if return_info == "timeline":
        info["timeline"] = timeline
    if return_info:
        return out, info
    return out


def spectrum_cmb_batch(modes, probs, jl, normalization="Cl", kinterp=None, direct=False, dkt0=math.pi, ntau=300, taucut=1e-2, bgsols=None, ptopts=None, nthreads=None,
                       return_info=False):
    """C_l^{AB} (A, B ∈ {T, E}) of several cosmologies sharing one model structure: the perturbation solves of ALL cosmologies go
    into one integrator launch (`solvept_batch`, each cosmology with its own τ-grid), followed per cosmology by the source, line-of-
    sight and C_l kernels of `spectrum_cmb` (reference: a serial loop over `spectrum_cmb(modes, probgen(θ), jl)`,
    docs/src/forecasting.md:56-59).  Per cosmology the result is bit-identical to `spectrum_cmb(modes, prob, jl, ...)`.
    Returns [ncosmo, nl, nmodes] (NaN where the background failed)."""
    import concurrent.futures as cf
    import os
    _require_cuda()
    modes = [modes] if isinstance(modes, str) else list(modes)
    for m in modes:
        if len(m) != 2 or m[0] not in "TE" or m[1] not in "TE":
            raise ValueError(f"spectrum_cmb_batch handles the T and E modes only, got {m}")
    if bgsols is None:
        def host(p):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return solvebg(p)
        with cf.ThreadPoolExecutor(nthreads or os.cpu_count()) as pool:
            bgsols = list(pool.map(host, probs))
    kinterp = kinterp if kinterp is not None else ChebyshevInterpolator(1e-2, 2e3, 60)
    good = [i for i, b in enumerate(bgsols) if b.success]
    out = np.full((len(probs), len(jl.l), len(modes)), np.nan)
    if not good:
        return (out, dict(background_failures=len(probs), mode_failures=0)) if return_info else out
    grids = [cmb_grids(bgsols[i], kinterp.minimum(), kinterp.maximum(), dkt0, ntau, taucut) for i in good]
    ks_list = [g[0] if direct else kinterp.xs for g in grids]
    arena = CosmoArena()
    # the sources are formed inside the integrator (each cosmology with its own save times and background table): only S reaches HBM, so
    # direct = True batches are bounded by nk·2·nτ·8 bytes per cosmology (9.7 MB), not by the saved states (397 MB)
    batch = solvept_batch([bgsols[i] for i in good], ks_list, saveat=[g[1] for g in grids], arena=arena, sources=dict(nS=2, scale_k=True), keep_states=False, **dict(ptopts or {}))
    cls = []
    for j, i in enumerate(good):
        prob, sol = probs[i], batch.sols[j]
        ks_fine, taus = grids[j]
        theta = los_integrate(SourceGrid(sol.d_S, sol.ks, taus, sol), jl, ks_fine=ks_fine, kinterp=None if direct else kinterp)
        cls.append(spectrum_cmb_from_theta(theta, modes, spectrum_primordial(ks_fine, prob), jl.l, ks_fine, normalization))
    allcl = torch.stack(cls).cpu().numpy()  # [ngood, nmodes, nl]
    for j, i in enumerate(good):
        out[i] = allcl[j].T
    if return_info:
        return out, dict(background_failures=len(probs) - len(good), mode_failures=int((batch.d_retcode != 0).sum().item()))
    return out


def _central_log_points(theta0, relstep):
    """2p points of a central difference in ln θ (reference test: FiniteDiff central, relstep 1e-3 on log-parameters, runtests.jl:375,401).
    Step per parameter h_j = relstep·max(|ln θ_j|, 1)... FiniteDiff uses relstep·|x| with an absolute floor; parameters may be negative
    (w0), so the log is taken of |θ| and the sign restored."""
    theta0 = np.asarray(theta0, dtype=np.float64)
    x = np.log(np.abs(theta0))
    h = relstep * np.maximum(np.abs(x), 1.0)
    pts = []
    for j in range(len(x)):
        for sgn in (+1, -1):
            y = x.copy()
            y[j] += sgn * h[j]
            pts.append(np.sign(theta0) * np.exp(y))
    return np.array(pts), h


_PRIMORDIAL = ("ln_As1e10", "ns")  # enter only through P0(k): their columns are closed-form, no lane needed


def _lane_problems(prob, names, delta, central=False):
    """The lane cosmologies of a sensitivity: lane 0 = `prob`, then per non-primordial parameter one lane moved by +δ in ln|θ_j| (sign
    kept) and, with `central`, one moved by −δ; parameters that enter only through the primordial spectrum need no lane.
    Returns (lane parameter names, problems, signed steps per lane)."""
    lane_names = [n for n in names if n not in _PRIMORDIAL]
    per = 2 if central else 1
    if per * len(lane_names) > 7:
        raise ValueError("at most 7 parameter lanes per call (8 warps per CTA): %d parameters × %d" % (len(lane_names), per))
    upd = parameter_updater(prob, lane_names)
    th0 = np.array([prob.pars[n] for n in lane_names], dtype=np.float64)
    probs, steps = [prob], [0.0]
    for j in range(len(lane_names)):
        for sgn in ((+1.0, -1.0) if central else (+1.0,)):
            th = th0.copy()
            th[j] *= math.exp(sgn * delta)
            probs.append(upd(th))
            steps.append(sgn * delta)
    return lane_names, probs, steps


def _lane_quotients(vals, steps, central):
    """Difference quotients per parameter from per-lane values (lane order of `_lane_problems`)."""
    out = []
    j = 1
    while j < len(vals):
        if central:
            out.append((vals[j] - vals[j + 1]) / (steps[j] - steps[j + 1]))
            j += 2
        else:
            out.append((vals[j] - vals[0]) / steps[j])
            j += 1
    return out


def _lane_backgrounds(probs, bgsol=None):
    bg0 = bgsol if bgsol is not None else solvebg(probs[0])
    if not bg0.success:
        raise RuntimeError("sensitivity: the primal background solve failed")
    return [solvebg_lock(p, bg0) for p in probs]  # lane 0 too: the primal re-solved on its own steps shares the lanes' arithmetic path (rounding cancels in the quotients)


def sensitivity_background(prob, names, delta=1e-6, bgsol=None):
    """dτ0/dθ_j and dκ0/dθ_j (derivatives with respect to θ itself, as the reference's "Background differentiation test",
    test/runtests.jl:480-491, which pushes ForwardDiff duals through `solvebg`): lockstep background lanes (`solvebg_lock`), one per
    parameter, on the primal's step sequence -- the quotient is the derivative of the discrete solution map including the moving
    event time.  Returns dict(tau0 = [p], kappa0 = [p])."""
    bg0 = bgsol if bgsol is not None else solvebg(prob)
    bg0 = solvebg_lock(prob, bg0)  # the primal re-solved on its own steps: the same arithmetic path as the lanes, so that rounding cancels in the quotient
    upd = parameter_updater(prob, list(names))
    th0 = np.array([prob.pars[n] for n in names], dtype=np.float64)
    dt0, dk0 = np.zeros(len(names)), np.zeros(len(names))
    for j in range(len(names)):
        h = delta * max(abs(th0[j]), 1e-300)
        th = th0.copy()
        th[j] += h
        lane = solvebg_lock(upd(th), bg0)
        dt0[j], dk0[j] = (lane.tau0 - bg0.tau0) / h, (lane.kappa0 - bg0.kappa0) / h
    return dict(tau0=dt0, kappa0=dk0)


def sensitivity_matter(prob, names, ks, method="lanes", delta=None, central=False, norm_partials=True, relstep=1e-3, kτini=1e-2, τinimax=1e-4, bgsol=None, return_info=False, **kw):
    """∂ln P(k)/∂ln θ_j, [nk, p] (BASELINE config 5's quantity; the reference obtains it with ForwardDiff duals through the whole solve
    and tests it against a central finite difference, runtests.jl:363-376).
    method = "lanes" (default): the primal and one cosmology per non-primordial parameter (moved by `delta` in ln θ) are solved IN
    LOCKSTEP -- backgrounds on the primal's step sequence (`solvebg_lock`), perturbations by CTAs of 1 + p warps sharing one step
    controller with the partials in its error norm (`solvept_lanes`) -- so that the difference quotient is the derivative of the discrete
    solution map, as forward-mode AD gives it, at the cost of (1 + p) solves running side by side; ln_As1e10 and ns columns are closed-form.
    delta: step in ln θ (default 1e-5 one-sided: truncation O(δ), rounding noise O(1e-11/δ) -- lockstep removes the step-selection noise
    of independent solves, ≈1e-4 relative, not the rounding); central = True: ±δ lanes (1 + 2p warps per mode, default δ = 1e-3, truncation
    O(δ²)); norm_partials: the shared controller's error norm covers the partials, as OrdinaryDiffEq's norm of Duals does (≈1.6× the
    primal's steps); False: the lanes take exactly the primal's steps.
    method = "fd": round-1 path, 2p independent solves, central difference with relative step `relstep` (noisy: adaptive step
    sequences differ between the two sides)."""
    ks = np.ascontiguousarray(np.atleast_1d(ks), dtype=np.float64)
    if method == "fd":
        th0 = np.array([prob.pars[n] for n in names], dtype=np.float64)
        pts, h = _central_log_points(th0, relstep)
        P = spectrum_matter_sweep(prob, names, pts, ks, chunk=len(pts), kτini=kτini, τinimax=τinimax, **kw)
        L = np.log(P)
        return np.stack([(L[2 * j] - L[2 * j + 1]) / (2 * h[j]) for j in range(len(names))], axis=1)
    if method != "lanes":
        raise ValueError("method must be 'lanes' or 'fd'")
    delta = (1e-3 if central else 1e-5) if delta is None else float(delta)
    lane_names, probs, steps = _lane_problems(prob, names, delta, central)
    J = np.zeros((len(ks), len(names)))
    info = dict(lanes=len(probs), delta=delta, central=central)
    if len(probs) > 1:
        bgs = _lane_backgrounds(probs, bgsol)
        sols = solvept_lanes(bgs, ks, [0.0] + [(1.0 / st if norm_partials else 0.0) for st in steps[1:]], ptivini=lambda k: min(kτini / k, τinimax) if k > 0 else τinimax, **kw)
        lnP = []
        for p, b, sol in zip(probs, bgs, sols):
            d = b.device()
            dm = torch.empty(len(ks), dtype=torch.float64, device=sol.d_uend.device)
            rc = p.lib.sbm_delta_m(_cptr(d["P"]), C.c_int(len(b.t)), _cptr(d["t"]), _cptr(d["y"]), _cptr(d["dy"]), C.c_double(b.tau0), C.c_int(len(ks)), _cptr(sol.d_ks), _cptr(sol.d_uend), _cptr(dm), _stream())
            if rc != 0:
                raise RuntimeError(f"sbm_delta_m failed with code {rc}")
            lnP.append(np.log(spectrum_primordial(ks, p) * dm.cpu().numpy() ** 2))
        info.update(success=all(s_.success for s_ in sols), attempts=int((sols[0].stats[:, 0] + sols[0].stats[:, 1]).sum()))
        for n, q in zip(lane_names, _lane_quotients(lnP, steps, central)):
            J[:, names.index(n)] = q
    for n in names:
        if n == "ln_As1e10":
            J[:, names.index(n)] = prob.pars["ln_As1e10"]                                   # P ∝ exp(x): ∂ln P/∂ln x = x
        elif n == "ns":
            J[:, names.index(n)] = prob.pars["ns"] * np.log(ks / prob.derived["kpivot"])    # P ∝ (k/kp)^(ns−1)
    return (J, info) if return_info else J


def sensitivity_cmb(mode, prob, names, jl, method="lanes", delta=None, central=False, norm_partials=True, relstep=1e-3, normalization="Dl", kinterp=None, dkt0=math.pi, ntau=300,
                    taucut=1e-2, bgsol=None, return_info=False, **kw):
    """∂ln C_l^{mode}/∂ln θ_j, [nl, p] (BASELINE config 5; reference ForwardDiff.jacobian of log D_l, runtests.jl:391-406).
    method = "lanes": as in `sensitivity_matter`; every lane runs sources -> line of sight -> C_l with its OWN τ0 (χ = τ0 − τ, save times)
    on the primal's k-quadrature grid and Chebyshev nodes (grids are index sets, not functions of θ -- as with duals, where the grids
    are built from values); ln_As1e10 and ns columns come from the primal's Θ_l(k) with the weights ∂P0/∂θ.  delta: default 1e-4 one-sided
    (C_l sums ≈6e5 source values per multipole, its rounding noise is ≈1e-7 relative, so a smaller step only amplifies it), 1e-3 with
    central = True.  method = "fd": round-1 path (independent solves: step-selection noise ≈1e-4 relative in C_l, i.e. ≈0.1 in the quotient)."""
    if method == "fd":
        th0 = np.array([prob.pars[n] for n in names], dtype=np.float64)
        pts, h = _central_log_points(th0, relstep)
        upd = parameter_updater(prob, names)
        Cl = spectrum_cmb_batch([mode], [upd(t) for t in pts], jl, normalization=normalization, **kw)[:, :, 0]
        L = np.log(np.abs(Cl))
        return np.stack([(L[2 * j] - L[2 * j + 1]) / (2 * h[j]) for j in range(len(names))], axis=1)
    if method != "lanes":
        raise ValueError("method must be 'lanes' or 'fd'")
    if len(mode) != 2 or mode[0] not in "TE" or mode[1] not in "TE":
        raise ValueError("sensitivity_cmb handles the T and E modes")
    delta = (1e-3 if central else 1e-4) if delta is None else float(delta)
    lane_names, probs, steps = _lane_problems(prob, names, delta, central)
    bgs = _lane_backgrounds(probs, bgsol)
    kinterp = kinterp if kinterp is not None else ChebyshevInterpolator(1e-2, 2e3, 60)
    ks_fine, taus0 = cmb_grids(bgs[0], kinterp.minimum(), kinterp.maximum(), dkt0, ntau, taucut)
    i0 = int(np.searchsorted(bgs[0].t, taucut, side="left"))  # first knot ≥ τcut: the same knot in every lane (lockstep backgrounds)
    cg = cosgrid(0.0, 1.0, length=ntau)
    saves = []
    for b in bgs:
        tj = b.t[i0] + (b.t[-1] - b.t[i0]) * cg
        tj[-1] = b.t[-1]
        saves.append(tj)
    assert np.array_equal(saves[0], taus0)
    if len(bgs) > 1:
        sols = solvept_lanes(bgs, kinterp.xs, [0.0] + [(1.0 / st if norm_partials else 0.0) for st in steps[1:]], saveat=saves, sources=dict(nS=2, scale_k=True), **kw)
    else:
        sols = [solvept(prob, bgs[0], kinterp.xs, saveat=saves[0], sources=dict(nS=2, scale_k=True), keep_states=False, **kw)]
    lnC, theta0, Cl0 = [], None, None
    for p, b, sol, tj in zip(probs, bgs, sols, saves):
        theta = los_integrate(SourceGrid(sol.d_S, kinterp.xs, tj, sol), jl, ks_fine=ks_fine, kinterp=kinterp)
        Cl = spectrum_cmb_from_theta(theta, [mode], spectrum_primordial(ks_fine, p), jl.l, ks_fine, normalization).cpu().numpy()[0]
        if theta0 is None:
            theta0, Cl0 = theta, Cl
        lnC.append(np.log(np.abs(Cl)))
    J = np.zeros((len(jl.l), len(names)))
    for n, q in zip(lane_names, _lane_quotients(lnC, steps, central)):
        J[:, names.index(n)] = q
    P0 = spectrum_primordial(ks_fine, prob)
    for n in names:
        if n == "ln_As1e10":
            J[:, names.index(n)] = prob.pars["ln_As1e10"]
        elif n == "ns":  # ∂C_l/∂ns = Σ_k c_k ln(k/kp) Θ^A Θ^B
            dC = spectrum_cmb_from_theta(theta0, [mode], P0 * np.log(ks_fine / prob.derived["kpivot"]), jl.l, ks_fine, normalization).cpu().numpy()[0]
            J[:, names.index(n)] = prob.pars["ns"] * dC / Cl0  # d ln|C| = dC / C
    info = dict(lanes=len(probs), delta=delta, central=central, success=all(s_.success for s_ in sols), attempts=int((sols[0].stats[:, 0] + sols[0].stats[:, 1]).sum()))
    return (J, info) if return_info else J
