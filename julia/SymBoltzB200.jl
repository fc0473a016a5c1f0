# SymBoltzB200.jl -- sketch of the Julia package extension that routes SymBoltz's hot path to the B200 libraries through `ccall`.
# NOT EXECUTED in this repository's CI: the build image has no Julia toolchain (see DESIGN.md §1).  It is written against the
# same C ABI (include/symboltz_b200.h) that symboltz.jl_b200/api.py binds with ctypes, and mirrors that file function by function.
module SymBoltzB200

using SymBoltz, CUDA   # CUDA.jl is used only as a device allocator (CuArray / CuPtr); all kernels live in the shared libraries

struct B200Rodas5P end  # marker algorithm type: `solve(prob, ks; ptopts = (alg = B200Rodas5P(),))`

const RETCODE = Dict(0 => :Success, 1 => :MaxIters, 2 => :DtLessThanMin, 3 => :Unstable)

libsbm(lmax, nx, w0wa) = joinpath(@__DIR__, "..", "symboltz.jl_b200", "_build", "l$(lmax)_x$(nx)_$(w0wa ? "w0wa" : "lcdm")", "libsbm_l$(lmax)_x$(nx)_$(w0wa ? "w0wa" : "lcdm").so")
const libsbl = joinpath(@__DIR__, "..", "symboltz.jl_b200", "_build", "libsbl.so")

"Parameter vector in the layout documented in include/symboltz_b200.h (dependent parameters as in src/models/*.jl)."
function parameter_vector(prob::SymBoltz.CosmologyProblem, bgsol)
    M = prob.M
    ps = bgsol.ps
    x, W = SymBoltz.momentum_quadrature(x -> 1 / (exp(x) + 1), 4)
    dl = @. -x / (1 + exp(-x))
    Ch = 3 / (8π) * ps[M.h.Ω₀] / ps[M.h.Iρ₀]
    return Float64[ps[M.g.h], ps[M.c.Ω₀], ps[M.b.Ω₀], ps[M.γ.Ω₀], ps[M.ν.Ω₀], Ch, ps[M.Λ.Ω₀], ps[M.γ.T₀], ps[M.b.YHe], ps[M.b.fHe], ps[M.h.y₀],
                   -1.0, 0.0, 1.0, ps[M.b.κ0], ps[M.τ0], x..., W..., dl...]
end

"Drop-in for `solvept(ptprob, bgsol, ks, ptivini; ...)` (src/solve.jl:543-569)."
function solvept_b200(lib, P, bgsol, ks, ptivini; reltol = 1e-5, abstol = 1e-5, saveat = Float64[], maxiters = 100_000, msub = 16, nbeta)
    ts = bgsol.t; nb = length(ts)
    y = reduce(hcat, bgsol(ts, Val{0}).u); dy = reduce(hcat, bgsol(ts, Val{1}).u)      # src/utils.jl:118-127
    tini = clamp.(ptivini.(ks), ts[begin], ts[end])                                    # src/solve.jl:527
    order = Int32.(sortperm(ks; rev = true) .- 1)
    nlut = 4096; s0 = log(ts[begin]); dsl = (log(ts[end]) - s0) / nlut
    lut = Int32.(clamp.(searchsortedlast.(Ref(ts), exp.(s0 .+ dsl .* (0:nlut-1))) .- 1, 0, nb - 2))
    dP, dt, dy_, ddy = CuArray(P), CuArray(ts), CuArray(vec(y)), CuArray(vec(dy))
    dtab = CUDA.zeros(Float64, ((nb - 1) * msub + 1) * 2 * nbeta)
    dks, dtini, dorder, dlut, dsave = CuArray(Float64.(ks)), CuArray(tini), CuArray(order), CuArray(lut), CuArray(Float64.(saveat))
    N = 82; nk = length(ks); ns = length(saveat)
    dusave = CUDA.zeros(Float64, max(1, nk * ns * N)); duend = CUDA.zeros(Float64, nk * N)
    dret = CUDA.zeros(Int32, nk); dstats = CUDA.zeros(Int64, 4nk); dqueue = CUDA.zeros(Int32, 1)
    GC.@preserve dP dt dy_ ddy dtab dks dtini dorder dlut dsave dusave duend dret dstats dqueue begin
        rc = ccall((:sbm_build_table, lib), Cint, (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, CuPtr{Float64}, Ptr{Cvoid}),
                   dP, nb, dt, dy_, ddy, msub, dtab, C_NULL)
        rc == 0 || error("sbm_build_table failed ($rc)")
        rc = ccall((:sbm_solvept, lib), Cint,
                   (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, Cint, Cdouble, Cdouble, CuPtr{Int32}, CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64},
                    CuPtr{Int32}, Cdouble, Cint, CuPtr{Float64}, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, Cint, Ptr{Cvoid}, Ptr{Cdouble}, Cint),
                   dP, nb, dt, dy_, ddy, msub, nlut, s0, dsl, dlut, dtab, nk, dks, dtini, dorder, ts[end], ns, dsave, reltol, abstol, maxiters,
                   ns > 0 ? pointer(dusave) : CU_NULL, duend, dret, dstats, dqueue, 0, C_NULL, C_NULL, 0)
        rc >= 0 || error("sbm_solvept failed ($rc)")
    end
    ret = Array(dret)
    for (i, r) in enumerate(ret)   # warn, don't throw (src/solve.jl:557-560)
        r != 0 && @warn "Perturbation (mode k = $(ks[i])) solution failed with return code $(RETCODE[r]).\nCheck the parameters and precision settings!"
    end
    return (; uend = reshape(Array(duend), N, nk), usave = ns > 0 ? reshape(Array(dusave), N, ns, nk) : nothing, retcode = ret, stats = reshape(Array(dstats), 4, nk))
end

"""
Parameter sweep (docs/src/forecasting.md:56-59: `for θ in θs; spectrum_matter(probgen(θ), ks); end`) as three library calls:
all backgrounds in one kernel (`sbm_solvebg_batch`, one thread per cosmology), one β-table per cosmology (`sbm_build_table`), and
ONE integrator launch over all (cosmology, mode) pairs (`sbm_solvept_batch`).  `Ps` is the npar × n matrix of parameter vectors
(`parameter_vector` without κ0/τ0: the device fills them in).  Mirrors `spectrum_matter_sweep(..., background = "device")` of api.py.
"""
function sweep_b200(lib, Ps::Matrix{Float64}, ks; cap = 4096, msub = 16, nbeta, reltol = 1e-5, abstol = 1e-5, maxiters = 100_000, N = 84)
    n = size(Ps, 2); nk = length(ks)
    dP = CuArray(Ps); dt = CUDA.zeros(Float64, cap, n); dy = CUDA.zeros(Float64, 5, cap, n); ddy = CUDA.zeros(Float64, 5, cap, n)
    dinfo = CUDA.zeros(Float64, 8, n); dnb = CUDA.zeros(Int32, n)
    rc = ccall((:sbm_solvebg_batch, lib), Cint, (Cint, CuPtr{Float64}, Cdouble, Cdouble, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, Ptr{Cvoid}),
               n, dP, 1e-6, 100.0, 1e-7, 1e-7, cap, dt, dy, ddy, dinfo, dnb, C_NULL)
    rc == 0 || error("sbm_solvebg_batch failed ($rc)")
    nb = Array(dnb); info = Array(dinfo)                       # τ0 = info[1, c], κ0 = info[2, c], retcode = info[4, c]
    # one sbm_cosmo_t (112 bytes, layout in the header) per cosmology: pointers into the batch arrays, β-table, interval look-up
    recs = Vector{UInt8}(undef, 112n); tabs = CuArray{Float64}[]; luts = CuArray{Int32}[]
    for c in 1:n
        ts = Array(view(dt, 1:nb[c], c)); nlut = 4096; s0 = log(ts[begin]); dsl = (log(ts[end]) - s0) / nlut
        push!(luts, CuArray(Int32.(clamp.(searchsortedlast.(Ref(ts), exp.(s0 .+ dsl .* (0:nlut-1))) .- 1, 0, nb[c] - 2))))
        push!(tabs, CUDA.zeros(Float64, ((nb[c] - 1) * msub + 1) * 2 * nbeta))
        pP, pt, py, pd = pointer(dP, 1 + (c - 1) * size(Ps, 1)), pointer(dt, 1 + (c - 1) * cap), pointer(dy, 1 + (c - 1) * 5cap), pointer(ddy, 1 + (c - 1) * 5cap)
        ccall((:sbm_build_table, lib), Cint, (CuPtr{Float64}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cint, CuPtr{Float64}, Ptr{Cvoid}), pP, nb[c], pt, py, pd, msub, tabs[end], C_NULL)
        write_cosmo_record!(recs, c, pP, nb[c], pt, py, pd, msub, nlut, s0, 1 / dsl, pointer(luts[end]), pointer(tabs[end]), info[1, c])   # field by field, as CosmoArena.load does
    end
    dcos = CuArray(recs); kall = repeat(Float64.(ks), n); dks = CuArray(kall)
    dtini = CuArray(repeat(clamp.(min.(1e-2 ./ ks, 1e-4), 1e-6, Inf), n)); dcof = CuArray(Int32.(repeat(0:n-1; inner = nk))); dorder = CuArray(Int32.(sortperm(kall; rev = true) .- 1))
    duend = CUDA.zeros(Float64, N, nk * n); dret = CUDA.zeros(Int32, nk * n); dstats = CUDA.zeros(Int64, 4, nk * n); dqueue = CUDA.zeros(Int32, 1)
    rc = ccall((:sbm_solvept_batch, lib), Cint, (Cint, CuPtr{UInt8}, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, CuPtr{Int32}, Cint, Cdouble, Cdouble, Cint, CuPtr{Float64}, CuPtr{Float64},
                CuPtr{Int32}, CuPtr{Int64}, CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, Cint, CuPtr{Float64}, CuPtr{Int32}, Ptr{Cvoid}),
               n, dcos, nk * n, dks, dtini, dcof, dorder, 0, reltol, abstol, maxiters, CU_NULL, duend, dret, dstats, dqueue, CU_NULL, CU_NULL, 0, CU_NULL, CU_NULL, C_NULL)
    rc >= 0 || error("sbm_solvept_batch failed ($rc)")
    return reshape(Array(duend), N, nk, n), reshape(Array(dret), nk, n)    # Δm and P(k) follow with sbm_delta_m per cosmology
end

end # module
