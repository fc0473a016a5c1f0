# SymBoltzB200.jl -- the Julia package extension that routes SymBoltz's hot path to the B200 libraries through `ccall`.
#
# STATUS: written against the C ABI of include/symboltz_b200.h, which symboltz.jl_b200/api.py binds with ctypes and the test-suite
# exercises symbol by symbol; THIS FILE HAS NEVER BEEN PARSED OR RUN -- the build image has no Julia toolchain (DESIGN.md §1).  Treat
# it as the binding a maintainer would start from, not as tested code.
#
# Dispatch: the existing call sites keep their signatures; the extension adds methods specialised on a marker algorithm type,
#     solve(prob, ks; ptopts = (alg = B200Rodas5P(),))            src/solve.jl:380-402
#     solvept(ptprob, bgsol, ks, ptivini; alg = B200Rodas5P(), …)  src/solve.jl:543-569
#     source_grid(prob, S, τs, ks, bgsol; ptopts = (alg = B200Rodas5P(),))   src/observables/fourier.jl:267-281
# Streams: CUDA.jl allocates and copies on its task-local stream, so that stream's handle is passed to every library call and the
# results are read after `CUDA.synchronize()` (the round-1 sketch passed C_NULL = the legacy default stream, which does not order
# against CUDA.jl's copies).
module SymBoltzB200

using SymBoltz, CUDA   # CUDA.jl is used only as a device allocator (CuArray / CuPtr); all kernels live in the shared libraries

struct B200Rodas5P end  # marker algorithm types: dispatch targets for `solvept(...; alg = ...)` / `ptalg(prob; accuracy)` (src/solve.jl:326-341)
struct B200KenCarp4 end # accuracy = 1
struct B200TRBDF2 end   # accuracy = 0
const B200Alg = Union{B200Rodas5P, B200KenCarp4, B200TRBDF2}

const RETCODE = Dict(0 => :Success, 1 => :MaxIters, 2 => :DtLessThanMin, 3 => :Unstable, 4 => :ScheduleTimeout)

modelkey(lmax, nx, w0wa) = "l$(lmax)_x$(nx)_$(w0wa ? "w0wa" : "lcdm")"
libsbm(lmax, nx, w0wa) = joinpath(@__DIR__, "..", "symboltz.jl_b200", "_build", modelkey(lmax, nx, w0wa), "libsbm_$(modelkey(lmax, nx, w0wa)).so")
const libsbl = joinpath(@__DIR__, "..", "symboltz.jl_b200", "_build", "libsbl.so")
const libsbc = joinpath(@__DIR__, "..", "symboltz.jl_b200", "_build", "libsbc.so")

"Model constants from the library itself (N, NPAR, NBETA, …, index of κ0 and τ0 in P): nothing about the model is hard-coded here."
function model_info(lib)
    out = zeros(Cint, 16)
    ccall((:sbm_info, lib), Cint, (Ptr{Cint},), out)
    return (; N = Int(out[1]), npar = Int(out[2]), nbeta = Int(out[3]), lmax = Int(out[5]), nx = Int(out[6]), w0wa = out[7] != 0, iκ0 = Int(out[14]) + 1, iτ0 = Int(out[15]) + 1)
end

stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))   # the stream CUDA.jl's copies of this task run on

"Parameter vector in the layout documented in include/symboltz_b200.h (dependent parameters as in src/models/*.jl); nx momentum nodes."
function parameter_vector(prob::SymBoltz.CosmologyProblem, bgsol, info)
    M = prob.M
    ps = bgsol.ps
    x, W = SymBoltz.momentum_quadrature(x -> 1 / (exp(x) + 1), info.nx)
    dl = @. -x / (1 + exp(-x))
    Ch = 3 / (8π) * ps[M.h.Ω₀] / ps[M.h.Iρ₀]
    w0, wa, cs2 = info.w0wa ? (ps[M.X.w0], ps[M.X.wa], ps[M.X.cₛ²]) : (-1.0, 0.0, 1.0)
    P = Float64[ps[M.g.h], ps[M.c.Ω₀], ps[M.b.Ω₀], ps[M.γ.Ω₀], ps[M.ν.Ω₀], Ch, info.w0wa ? ps[M.X.Ω₀] : ps[M.Λ.Ω₀], ps[M.γ.T₀], ps[M.b.YHe], ps[M.b.fHe], ps[M.h.y₀],
                w0, wa, cs2, ps[M.b.κ0], ps[M.τ0], x..., W..., dl...]
    length(P) == info.npar || error("parameter vector has $(length(P)) entries, the library expects $(info.npar)")
    return P
end

"Interval look-up of the β-table: lut[q] = knot interval containing exp(s0 + q·dsl) (what api.py's BackgroundSolution.device builds)."
function interval_lut(ts; nlut = 4096)
    s0 = log(ts[begin]); dsl = (log(ts[end]) - s0) / nlut
    lut = Int32.(clamp.(searchsortedlast.(Ref(ts), exp.(s0 .+ dsl .* (0:nlut-1))) .- 1, 0, length(ts) - 2))
    return lut, s0, dsl
end

struct SbmSrc               # sbm_src_t
    dsrcbg::CuPtr{Float64}
    dS::CuPtr{Float64}
    nS::Cint
    scale_k::Cint
    taurec::Cdouble
end

"""
`solvept` on the staged (device-pointer) level: upload knots -> `sbm_build_table` -> `sbm_solvept_src` (sources formed inside the
integrator at `saveat` when `sources = true`, the reference's output_func of source_grid, src/observables/fourier.jl:272-278).
"""
function SymBoltz.solvept(ptprob, bgsol, ks::AbstractArray, ptivini, alg::B200Alg; lib, reltol = 1e-5, abstol = 1e-5, saveat = Float64[], maxiters = 100_000, msub = 16,
                          sources = false, scale_k = true, τrec = 0.0, P)
    info = model_info(lib); N = info.N
    ts = bgsol.t; nb = length(ts); nk = length(ks); ns = length(saveat)
    y = reduce(hcat, bgsol(ts, Val{0}).u); dy = reduce(hcat, bgsol(ts, Val{1}).u)      # 5 × nb column-major == [nb][5]; src/utils.jl:118-127
    tini = clamp.(ptivini.(ks), ts[begin], ts[end])                                    # src/solve.jl:527
    order = Int32.(sortperm(ks; rev = true) .- 1)
    lut, s0, dsl = interval_lut(ts)
    dP, dt, dy_, ddy = CuArray(P), CuArray(ts), CuArray(vec(y)), CuArray(vec(dy))
    dtab = CUDA.zeros(Float64, ((nb - 1) * msub + 1) * 2 * info.nbeta)
    dks, dtini, dorder, dlut, dsave = CuArray(Float64.(ks)), CuArray(tini), CuArray(order), CuArray(lut), CuArray(Float64.(saveat))
    duend = CUDA.zeros(Float64, nk * N); dret = CUDA.zeros(Int32, nk); dstats = CUDA.zeros(Int64, 4nk); dqueue = CUDA.zeros(Int32, 1)
    stride = ccall((:sbm_srcbg_stride, lib), Cint, ())
    dsb = CUDA.zeros(Float64, max(1, ns * stride)); dS = CUDA.zeros(Float64, max(1, nk * 2 * ns))
    st = stream()
    GC.@preserve dP dt dy_ ddy dtab dks dtini dorder dlut dsave duend dret dstats dqueue dsb dS begin
        rc = ccall((:sbm_build_table, lib), Cint, (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, CuPtr{Float64}, Ptr{Cvoid}), dP, nb, dt, dy_, ddy, msub, dtab, st)
        rc == 0 || error("sbm_build_table failed ($rc)")
        src = Ref(SbmSrc(pointer(dsb), pointer(dS), 2, scale_k ? 1 : 0, τrec))
        if sources && ns > 0
            rc = ccall((:sbm_srcbg, lib), Cint, (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), dP, nb, dt, dy_, ddy, ns, dsave, dsb, st)
            rc == 0 || error("sbm_srcbg failed ($rc)")
        end
        srcp = (sources && ns > 0) ? src : C_NULL
        # few modes (config 1, the default 61-node C_l path): one CTA per mode, same results bit for bit, about half the latency
        if !(alg isa B200Rodas5P)  # the ESDIRK integrators: one warp per mode, same argument list as the split entry point
            sym = alg isa B200KenCarp4 ? :sbm_solvept_kencarp4 : :sbm_solvept_trbdf2
            rc = ccall((sym, lib), Cint,
                       (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble, CuPtr{Int32}, CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64},
                        CuPtr{Int32}, Cdouble, Cint, CuPtr{Float64}, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, Ptr{Cvoid}, Ptr{SbmSrc}),
                       dP, nb, dt, dy_, ddy, msub, length(lut), s0, dsl, dlut, dtab, nk, dks, dtini, dorder, ts[end], ns, dsave, reltol, abstol, maxiters,
                       CU_NULL, duend, dret, dstats, dqueue, st, srcp)
        elseif 0 < nk <= ccall((:sbm_split_capacity, lib), Cint, ())
            rc = ccall((:sbm_solvept_split, lib), Cint,
                       (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble, CuPtr{Int32}, CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64},
                        CuPtr{Int32}, Cdouble, Cint, CuPtr{Float64}, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, Ptr{Cvoid}, Ptr{SbmSrc}),
                       dP, nb, dt, dy_, ddy, msub, length(lut), s0, dsl, dlut, dtab, nk, dks, dtini, dorder, ts[end], ns, dsave, reltol, abstol, maxiters,
                       CU_NULL, duend, dret, dstats, dqueue, st, srcp)
        else
            rc = ccall((:sbm_solvept_src, lib), Cint,
                       (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble, CuPtr{Int32}, CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64},
                        CuPtr{Int32}, Cdouble, Cint, CuPtr{Float64}, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, Cint, Ptr{Cvoid}, Ptr{SbmSrc}),
                       dP, nb, dt, dy_, ddy, msub, length(lut), s0, dsl, dlut, dtab, nk, dks, dtini, dorder, ts[end], ns, dsave, reltol, abstol, maxiters,
                       CU_NULL, duend, dret, dstats, dqueue, 0, st, srcp)
        end
        rc >= 0 || error("sbm_solvept failed ($rc)")
        CUDA.synchronize()
    end
    ret = Array(dret)
    for (i, r) in enumerate(ret)   # warn, don't throw (src/solve.jl:557-560)
        r != 0 && @warn "Perturbation (mode k = $(ks[i])) solution failed with return code $(get(RETCODE, Int(r), r)).\nCheck the parameters and precision settings!"
    end
    S = (sources && ns > 0) ? permutedims(reshape(Array(dS), ns, 2, nk), (1, 3, 2)) : nothing    # C [nk][nS][nτ] -> Ss[iτ, ik, iS] (fourier.jl:270-277)
    return (; uend = reshape(Array(duend), N, nk), S, retcode = ret, stats = reshape(Array(dstats), 4, nk))
end

"Write one sbm_cosmo_t (128 bytes, natural C layout, include/symboltz_b200.h) into `recs` at cosmology `c` (1-based)."
function write_cosmo_record!(recs::Vector{UInt8}, c, pP, nb, pt, py, pd, msub, nlut, s0, inv_dsl, plut, ptab, tend; psave = CU_NULL, psrcbg = CU_NULL, τrec = 0.0)
    io = IOBuffer(); u(p) = UInt64(UInt(p))
    write(io, u(pP)); write(io, Int32(nb)); write(io, Int32(0)); write(io, u(pt)); write(io, u(py)); write(io, u(pd))          # P, nb (+pad), t, y, dy          0..39
    write(io, Int32(nb)); write(io, Int32(msub)); write(io, Int32(nlut)); write(io, Int32(0)); write(io, Float64(s0)); write(io, Float64(inv_dsl))   # 40..71
    write(io, u(pt)); write(io, u(plut)); write(io, u(ptab)); write(io, Float64(tend)); write(io, u(psave)); write(io, u(psrcbg)); write(io, Float64(τrec))  # 72..127
    b = take!(io); length(b) == 128 || error("sbm_cosmo_t must be 128 bytes")
    recs[(c - 1) * 128 .+ (1:128)] .= b
    return recs
end

"""
Parameter sweep (docs/src/forecasting.md:56-59: `for θ in θs; spectrum_matter(probgen(θ), ks); end`) as three library calls:
all backgrounds in one kernel (`sbm_solvebg_batch`), one β-table per cosmology (`sbm_build_table`), ONE integrator launch over all
(cosmology, mode) pairs (`sbm_solvept_batch`).  `Ps` is the npar × n matrix of parameter vectors (κ0/τ0 are filled in on the device).
"""
function sweep_b200(lib, Ps::Matrix{Float64}, ks; cap = 4096, msub = 16, reltol = 1e-5, abstol = 1e-5, maxiters = 100_000)
    info = model_info(lib); N = info.N
    n = size(Ps, 2); nk = length(ks); st = stream()
    dP = CuArray(Ps); dt = CUDA.zeros(Float64, cap, n); dy = CUDA.zeros(Float64, 5, cap, n); ddy = CUDA.zeros(Float64, 5, cap, n)
    dinfo = CUDA.zeros(Float64, 8, n); dnb = CUDA.zeros(Int32, n)
    rc = ccall((:sbm_solvebg_batch, lib), Cint, (Cint, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, Ptr{Cvoid}),
               n, dP, 1e-6, 100.0, 1e-7, 1e-7, cap, dt, dy, ddy, dinfo, dnb, st)
    rc == 0 || error("sbm_solvebg_batch failed ($rc)")
    CUDA.synchronize()
    nb = Array(dnb); binfo = Array(dinfo)                      # τ0 = binfo[1, c], κ0 = binfo[2, c], retcode = binfo[4, c]
    recs = Vector{UInt8}(undef, 128n); tabs = CuArray{Float64}[]; luts = CuArray{Int32}[]
    for c in 1:n
        ts = Array(view(dt, 1:nb[c], c)); lut, s0, dsl = interval_lut(ts)
        push!(luts, CuArray(lut)); push!(tabs, CUDA.zeros(Float64, ((nb[c] - 1) * msub + 1) * 2 * info.nbeta))
        pP, pt, py, pd = pointer(dP, 1 + (c - 1) * size(Ps, 1)), pointer(dt, 1 + (c - 1) * cap), pointer(dy, 1 + (c - 1) * 5cap), pointer(ddy, 1 + (c - 1) * 5cap)
        ccall((:sbm_build_table, lib), Cint, (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, CuPtr{Float64}, Ptr{Cvoid}), pP, nb[c], pt, py, pd, msub, tabs[end], st)
        write_cosmo_record!(recs, c, pP, nb[c], pt, py, pd, msub, length(lut), s0, 1 / dsl, pointer(luts[end]), pointer(tabs[end]), binfo[1, c])
    end
    dcos = CuArray(recs); kall = repeat(Float64.(ks), n); dks = CuArray(kall)
    dtini = CuArray(repeat(clamp.(min.(1e-2 ./ ks, 1e-4), 1e-6, Inf), n)); dcof = CuArray(Int32.(repeat(0:n-1; inner = nk))); dorder = CuArray(Int32.(sortperm(kall; rev = true) .- 1))
    duend = CUDA.zeros(Float64, N, nk * n); dret = CUDA.zeros(Int32, nk * n); dstats = CUDA.zeros(Int64, 4, nk * n); dqueue = CUDA.zeros(Int32, 1)
    rc = ccall((:sbm_solvept_batch, lib), Cint, (Cint, CuPtr{UInt8}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int32}, Cint, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64},
                CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, Cint, CuPtr{Float64}, CuPtr{Int32}, Ptr{Cvoid}),
               n, dcos, nk * n, dks, dtini, dcof, dorder, 0, reltol, abstol, maxiters, CU_NULL, duend, dret, dstats, dqueue, CU_NULL, CU_NULL, 0, CU_NULL, CU_NULL, st)
    rc >= 0 || error("sbm_solvept_batch failed ($rc)")
    CUDA.synchronize()
    return reshape(Array(duend), N, nk, n), reshape(Array(dret), nk, n)    # Δm and P(k) follow with sbm_delta_m per cosmology
end

"""
Multi-GPU exchange with the NCCL communicator owned by the library (libsbc.so): one Julia process (or task with its own device) per GPU.
Rank 0 calls `unique_id()` and ships the 128 bytes to the others (file, socket, MPI); then every rank calls `comm_init`.
Sharded C_l of one cosmology: rank r solves the modes `r+1:world:nk` with `solvept(...; sources = true)`, scatters its rows into a
zero-initialised full `S[nτ, 2, nk]`, `allreduce_sum!(comm, dS)`, runs `sbl_los` / `sbl_cl` on the fine-k slice
`sbc_slice_begin(nkf, r, world) : sbc_slice_end(...)`, and `allreduce_sum!(comm, dCl)` (src/solve.jl:566 fan-out across GPUs).
"""
function unique_id()
    id = zeros(UInt8, ccall((:sbc_unique_id_bytes, libsbc), Cint, ()))
    rc = ccall((:sbc_unique_id, libsbc), Cint, (Ptr{UInt8},), id); rc == 0 || error("sbc_unique_id failed ($rc)")
    return id
end
function comm_init(id::Vector{UInt8}, rank, world)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:sbc_comm_init, libsbc), Cint, (Ptr{UInt8}, Cint, Cint, Ptr{Ptr{Cvoid}}), id, rank, world, h); rc == 0 || error("sbc_comm_init failed ($rc)")
    return h[]
end
function allreduce_sum!(comm, d::CuArray{Float64})
    rc = ccall((:sbc_allreduce_sum, libsbc), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Clonglong, Ptr{Cvoid}), comm, d, length(d), stream()); rc == 0 || error("sbc_allreduce_sum failed ($rc)")
    return d
end

end # module
