# BoltCUDA.jl -- the reference-side binding of libbolt_cuda.so (include/bolt_cuda.h).
#
# EXPERIMENTAL: Julia is not installed in the build image, so this file has never been executed there.  What IS checked
# automatically: every `ccall` below is parsed by tests/test_julia_shim.py and its argument tuple is compared, type by type,
# with the prototype in include/bolt_cuda.h; the struct layouts are asserted against the header as well.  The Python ctypes
# driver (bolt.jl_b200/capi.py + api.py) makes the same calls in the same order and is what the GPU tests drive.
#
# The module adds METHODS to Bolt's own exported functions (src/Bolt.jl:8-9), dispatched on two new types, so that no
# method of Bolt is overwritten and user scripts change in one place only -- the integrator tag:
#
#     using Bolt, BoltCUDA
#     dev  = BoltCUDA.Device()                                        # instead of BasicNewtonian()
#     sf   = source_grid(𝕡, bg, ih, ks, dev)                         # src/spectra.jl:6   -> DeviceSourceGrid
#     sf_P = source_grid_P(𝕡, bg, ih, ks, dev)                       # src/spectra.jl:25  (served from the same solve)
#     Cᵀᵀ  = cltt(ℓs, 𝕡, bg, ih, sf)                                 # src/spectra.jl:147 -> ONE bolt_project call
#     C₂   = cltt(2, 𝕡, bg, ih, sf)                                  # src/spectra.jl:132 (scalar ℓ = 1-element batch)
#     Cᵀᵀ  = cltt(ℓs, sf, quadratic_k(0.01bg.H₀, 1000bg.H₀, 5000), 𝕡, bg)   # src/spectra.jl:84 (explicit dense k grid)
#     sol  = boltsolve(Hierarchy(dev, 𝕡, bg, ih, k); reltol=1e-9)    # src/perturbations.jl:25 -> sol(x)
#     U    = boltsolve_rsa(Hierarchy(dev, 𝕡, bg, ih, k))             # src/perturbations.jl:86 -> Matrix(n, n_x)
#     pL   = plin(ks, 𝕡, bg, ih)                                      # vector of k: one batched call; plin(k, 𝕡, bg, ih, dev) scalar
module BoltCUDA

using Bolt
import Bolt: source_grid, source_grid_P, cltt, clte, clee, plin, boltsolve, boltsolve_rsa,
             AbstractCosmoParams, Hierarchy, quadratic_k
using ForwardDiff

const lib = get(ENV, "BOLT_CUDA_LIB", joinpath(@__DIR__, "..", "csrc", "libbolt_cuda.so"))
const ABI_VERSION = 3

"""Integrator tag that routes a call to the GPU (dispatch replaces `BasicNewtonian()`)."""
struct Device <: Bolt.PerturbationIntegrator
    ordinal::Int
end
Device() = Device(0)

# ---- C structs (must match include/bolt_cuda.h) -------------------------------------------------------------
struct CosmoDesc
    abi_version::Int32; nd::Int32; n_x::Int32; nq::Int32
    x0::Float64; dx::Float64
    scalars::Ptr{Float64}; quad_pts::Ptr{Float64}; quad_wts::Ptr{Float64}; tables::Ptr{Float64}
end
struct Opts
    l_gamma::Int32; l_nu::Int32; l_mnu::Int32; mode::Int32
    reltol::Float64; abstol::Float64; fixed_dt::Float64
    max_steps::Int64; ix_first::Int32; reserved::Int32
end
adaptive(ℓᵧ, ℓ_ν, ℓ_mν, reltol, abstol) = Opts(ℓᵧ, ℓ_ν, ℓ_mν, 0, reltol, abstol, 0.0, 0, 0, 0)

check(ctx, rc) = rc == 0 || error("libbolt_cuda error $rc: ", unsafe_string(ccall((:bolt_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))

# one context per Julia thread (the reference calls plin/cltt from many threads: ThreadPools, spectra.jl:10,149)
const ctxs = Dict{Tuple{Int,Int},Ptr{Cvoid}}()
const ctxs_lock = ReentrantLock()
function context(dev::Int)
    lock(ctxs_lock) do
        get!(ctxs, (Threads.threadid(), dev)) do
            r = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:bolt_init, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), dev, r)
            rc == 0 || error("bolt_init failed ($rc): no usable CUDA device; there is no CPU fallback")
            ccall((:bolt_abi_version, lib), Cint, ()) == ABI_VERSION || error("libbolt_cuda ABI version mismatch")
            r[]
        end
    end
end

# A Vector{Dual{Tag,Float64,N}} is bit-identical to a (1+N) x len Float64 matrix, value first.
flat(v::AbstractVector{Float64}) = (v, 1)
flat(v::AbstractVector{ForwardDiff.Dual{T,Float64,N}}) where {T,N} = (reinterpret(Float64, v), 1 + N)
coefs(itp) = vec(collect(itp.itp.coefs))            # spline(f, x_grid) = scale(interpolate(...)), src/util.jl:11
unflat(::Type{Float64}, a, nd) = vec(a)
unflat(::Type{D}, a, nd) where {D<:ForwardDiff.Dual} = collect(reinterpret(D, vec(a)))
plain(v) = Float64.(ForwardDiff.value.(v))

"""Upload what Background and IonizationHistory computed on the host (src/background.jl:104-128, recfast.jl:674-726),
run `f(ctx, cosmo, nd)` and free the device copy whatever happens."""
function with_cosmo(f, dev::Device, 𝕡::AbstractCosmoParams{T}, bg, ih) where T
    ctx = context(dev.ordinal)
    sc = T[𝕡.h, 𝕡.Ω_r, 𝕡.Ω_b, 𝕡.Ω_c, 𝕡.A, 𝕡.n, 𝕡.Y_p, 𝕡.N_ν, 𝕡.Σm_ν, bg.H₀, bg.η₀, bg.ρ_crit, bg.Ω_Λ]
    tabs = vcat((coefs(t) for t in (bg.ℋ, bg.ℋ′, bg.ℋ′′, bg.η, bg.ρ₀ℳ, ih.τ, ih.τ′, ih.τ′′, ih.g̃, ih.g̃′, ih.g̃′′, ih.csb²))...)
    scf, nd = flat(sc); tbf, _ = flat(tabs)
    qp = plain(bg.quad_pts); qw = plain(bg.quad_wts)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve scf tbf qp qw begin
        d = CosmoDesc(ABI_VERSION, nd, length(bg.x_grid), length(qp), first(bg.x_grid), step(bg.x_grid),
                      pointer(scf), pointer(qp), pointer(qw), pointer(tbf))
        check(ctx, ccall((:bolt_cosmo_upload, lib), Cint, (Ptr{Cvoid}, Ref{CosmoDesc}, Ref{Ptr{Cvoid}}), ctx, d, out))
    end
    try
        return f(ctx, out[], nd)
    finally
        ccall((:bolt_cosmo_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, out[])
    end
end

warn_status(status, where) = any(s -> s != 0 && s != 4, status) &&
    @warn "$where: some k-modes did not finish cleanly (1 max steps, 2 step underflow, 3 non-finite); the reference never checks retcode" count(s -> s != 0 && s != 4, status)

"""bolt_solve for a vector of k.  Returns (S_T, S_P, u_hist, status, nsteps); any of the three arrays may be skipped."""
function solve(𝕡::AbstractCosmoParams{T}, bg, ih, ks, dev::Device, o::Opts; sources=true, history=false) where T
    with_cosmo(dev, 𝕡, bg, ih) do ctx, c, nd
        k = plain(ks); nk = length(k); nx = length(bg.x_grid)
        n = Int(ccall((:bolt_state_dim, lib), Cint, (Cint, Cint, Cint, Cint), o.l_gamma, o.l_nu, o.l_mnu, length(bg.quad_pts)))
        S_T = sources ? zeros(Float64, nd * nx * nk) : Float64[]
        S_P = sources ? zeros(Float64, nd * nx * nk) : Float64[]
        hist = history ? zeros(Float64, nd * n * nx * nk) : Float64[]
        status = zeros(Int32, nk); nsteps = zeros(Int64, nk); nreject = zeros(Int64, nk)
        p(a) = isempty(a) ? Ptr{Float64}(C_NULL) : pointer(a)
        GC.@preserve k S_T S_P hist status nsteps nreject check(ctx, ccall((:bolt_solve, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
            ctx, c, k, nk, o, p(S_T), p(S_P), p(hist), C_NULL, status, nsteps, nreject))
        warn_status(status, "bolt_solve")
        (S_T, S_P, hist, status, nsteps, nd, n)
    end
end

# ---- boltsolve / boltsolve_rsa (src/perturbations.jl:25-33, 86-111) ------------------------------------------
"""What `boltsolve` returns on the device: callable like the reference's ODESolution, `sol(x) -> Vector(n)`.  The solution is
held on bg.x_grid (the device's Hermite dense output sampled at every grid point: exact there); between grid points a
four-point cubic through the neighbouring rows is used."""
struct DeviceSolution{T}
    t::Vector{Float64}
    u::Matrix{T}          # (n, n_x)
    retcode::Int
    nsteps::Int
end
function (sol::DeviceSolution)(x)
    nx = length(sol.t); dx = (sol.t[end] - sol.t[1]) / (nx - 1)
    t = (x - sol.t[1]) / dx
    i = clamp(floor(Int, t), 0, nx - 2)
    t == i && return sol.u[:, i + 1]
    j = clamp(i - 1, 0, nx - 4); s = t - j
    w = (-(s - 1) * (s - 2) * (s - 3) / 6, s * (s - 2) * (s - 3) / 2, -s * (s - 1) * (s - 3) / 2, s * (s - 1) * (s - 2) / 6)
    sum(w[m] * sol.u[:, j + m] for m in 1:4)
end

function boltsolve(h::Hierarchy{T,Device}, ode_alg=nothing; reltol=1e-6, abstol=1e-6) where T
    _, _, hist, status, nsteps, nd, n = solve(h.par, h.bg, h.ih, [h.k], h.integrator,
                                               adaptive(h.ℓᵧ, h.ℓ_ν, h.ℓ_mν, reltol, abstol); sources=false, history=true)
    u = reshape(unflat(T, hist, nd), n, length(h.bg.x_grid))       # C [n_x][n][nd] = Julia (n, n_x) of T
    DeviceSolution{T}(collect(Float64, h.bg.x_grid), u, Int(status[1]), Int(nsteps[1]))
end

function boltsolve_rsa(h::Hierarchy{T,Device}, ode_alg=nothing; reltol=1e-6, abstol=1e-6) where T
    sol = boltsolve(h; reltol=reltol, abstol=abstol)
    x_grid = h.bg.x_grid
    results = copy(sol.u)
    xrsa_hor = findfirst(>(240), @. h.k * h.bg.η)                                      # src/perturbations.jl:97-100
    xrsa_od = findfirst(>(100), @. -h.ih.τ′ * h.bg.ℋ / h.bg.η)
    xrsa_hor = isnothing(xrsa_hor) ? length(x_grid) : xrsa_hor
    xrsa_od = isnothing(xrsa_od) ? length(x_grid) : xrsa_od
    switch = x_grid[max(xrsa_hor, xrsa_od)]
    hb = Hierarchy(Bolt.BasicNewtonian(), h.par, h.bg, h.ih, h.k, h.ℓᵧ, h.ℓ_ν, h.ℓ_mν, h.nq)
    for i in findall(>(switch), x_grid)
        Bolt.rsa_perts!(view(results, :, i), hb, x_grid[i])                             # the reference's own post-processing
    end
    results
end

# ---- source grids (src/spectra.jl:6-42) ---------------------------------------------------------------------
"""A source grid that lives on the host as the reference's `LinearInterpolation((x_grid, k_grid), grid, Line())` (so every CPU
function of Bolt keeps working on it: `sf(x, k)`) and remembers the raw matrix for the device projection."""
struct DeviceSourceGrid{T,I}
    itp::I
    k_grid::Vector{Float64}
    grid::Matrix{T}       # (n_x, n_k)
    ih::Any               # the ionization history of the grid's cosmology (the explicit-k-grid methods do not receive it)
    dev::Device
end
(s::DeviceSourceGrid)(x, k) = s.itp(x, k)

function grids(𝕡::AbstractCosmoParams{T}, bg, ih, k_grid, dev::Device; ℓᵧ=8, reltol=1e-11) where T
    S_T, S_P, _, _, _, nd, _ = solve(𝕡, bg, ih, k_grid, dev, adaptive(ℓᵧ, 8, 10, reltol, 1e-6))   # Hierarchy defaults, perturbations.jl:20-21
    nx, nk = length(bg.x_grid), length(k_grid)
    mk(S) = begin
        g = reshape(unflat(T, S, nd), nx, nk)
        itp = Bolt.LinearInterpolation((bg.x_grid, k_grid), g, extrapolation_bc = Bolt.Line())
        DeviceSourceGrid{T,typeof(itp)}(itp, plain(k_grid), g, ih, dev)
    end
    mk(S_T), mk(S_P)
end

# ONE solve per k feeds both sources: the sibling grid of the last solve is kept, keyed on the identity of every input (the
# objects are referenced by the entry, so an id can never be recycled while the entry lives) and on the solver options.
const sibling = Ref{Any}(nothing)
const sibling_lock = ReentrantLock()
same(key, args...) = key !== nothing && length(key) == length(args) && all(a === b for (a, b) in zip(key, args))
function paired(which::Int, 𝕡, bg, ih, k_grid, dev; ℓᵧ=8, reltol=1e-11)
    lock(sibling_lock) do
        e = sibling[]
        if e !== nothing && e.which == which && same(e.key, 𝕡, bg, ih, k_grid, ℓᵧ, reltol)
            sibling[] = nothing
            return e.grid
        end
        sT, sP = grids(𝕡, bg, ih, k_grid, dev; ℓᵧ=ℓᵧ, reltol=reltol)
        sibling[] = (which = 3 - which, key = (𝕡, bg, ih, k_grid, ℓᵧ, reltol), grid = which == 1 ? sP : sT)
        which == 1 ? sT : sP
    end
end
source_grid(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device; kw...) = paired(1, 𝕡, bg, ih, k_grid, dev; kw...)
source_grid_P(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device; kw...) = paired(2, 𝕡, bg, ih, k_grid, dev; kw...)

# ---- cltt / clte / clee (src/spectra.jl:84-160) ---------------------------------------------------------------
"""The dense integration grid as (kmin, kmax, n) if `kgrid` is `quadratic_k(kmin, kmax, n)` (src/spectra.jl:60-63), else nothing."""
function quadratic_spec(kgrid)
    n = length(kgrid); n < 2 && return nothing
    kmax = Float64(ForwardDiff.value(kgrid[end])); k1 = Float64(ForwardDiff.value(kgrid[1]))
    kmin = (k1 * n^2 - kmax) / (n^2 - 1)
    ok = all(i -> isapprox(ForwardDiff.value(kgrid[i]), kmin + (kmax - kmin) * (i / n)^2; rtol = 1e-12), 1:n)
    ok ? (kmin, kmax, n) : nothing
end

"""TT, TE, EE for a vector of multipoles in ONE bolt_project call (instead of qmap over ℓ, src/spectra.jl:147-160)."""
function project(ℓ⃗, 𝕡::AbstractCosmoParams{T}, bg, sf, sf_P, spec) where T
    ref = sf === nothing ? sf_P : sf
    ells = Int32.(collect(ℓ⃗)); perm = sortperm(ells); ells_sorted = unique(ells[perm])      # the library wants strictly increasing multipoles
    kmin, kmax, nkd = spec
    tt, te, ee = with_cosmo(ref.dev, 𝕡, bg, ref.ih) do ctx, c, nd
        k = ref.k_grid; nk = length(k); nℓ = length(ells_sorted)
        g(s) = s === nothing ? Ptr{Float64}(C_NULL) : pointer(flat(vec(s.grid))[1])
        tt = zeros(Float64, nd * nℓ); te = similar(tt); ee = similar(tt)
        ix_start = findfirst(bg.x_grid .> -8) - 1                     # src/spectra.jl:86, 0-based across the ABI
        GC.@preserve sf sf_P k ells_sorted tt te ee check(ctx, ccall((:bolt_project, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Int32}, Cint, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            ctx, c, g(sf), g(sf_P), k, nk, ells_sorted, nℓ, kmin, kmax, nkd, ix_start, tt, te, ee))
        unflat(T, tt, nd), unflat(T, te, nd), unflat(T, ee, nd)
    end
    at = [searchsortedfirst(ells_sorted, ℓ) for ℓ in ells]             # back to the caller's order (duplicates allowed)
    tt[at], te[at], ee[at]
end
default_dense(bg) = (0.01 * Float64(ForwardDiff.value(bg.H₀)), 1000 * Float64(ForwardDiff.value(bg.H₀)), 5000)   # src/spectra.jl:133

# (ℓ⃗, par, bg, ih, sf[, sf_P]): src/spectra.jl:147-160
cltt(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf::DeviceSourceGrid) = project(ℓ⃗, 𝕡, bg, sf, nothing, default_dense(bg))[1]
clte(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf::DeviceSourceGrid, sf_P::DeviceSourceGrid) = project(ℓ⃗, 𝕡, bg, sf, sf_P, default_dense(bg))[2]
clee(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf_P::DeviceSourceGrid) = project(ℓ⃗, 𝕡, bg, nothing, sf_P, default_dense(bg))[3]
# (ℓ::Int, par, bg, ih, sf[, sf_P]): src/spectra.jl:132-145 -- a 1-element batch
cltt(ℓ::Int, 𝕡::AbstractCosmoParams, bg, ih, sf::DeviceSourceGrid) = cltt([ℓ], 𝕡, bg, ih, sf)[1]
clte(ℓ::Int, 𝕡::AbstractCosmoParams, bg, ih, sf::DeviceSourceGrid, sf_P::DeviceSourceGrid) = clte([ℓ], 𝕡, bg, ih, sf, sf_P)[1]
clee(ℓ::Int, 𝕡::AbstractCosmoParams, bg, ih, sf_P::DeviceSourceGrid) = clee([ℓ], 𝕡, bg, ih, sf_P)[1]
# (ℓ, s_itp, kgrid, par, bg): src/spectra.jl:84-130, explicit dense k grid.  The device integrates on quadratic_k grids (every
# grid the reference builds); any other grid falls back to the reference's own CPU loop on the host copy of the sources.
function explicit(which, ℓ, sf, sf_P, kgrid, 𝕡, bg)
    spec = quadratic_spec(kgrid)
    if spec === nothing
        @warn "dense k grid is not a quadratic_k grid: integrating on the CPU (reference path)" maxlog = 1
        f(l) = which == 1 ? invoke(cltt, Tuple{Any,Any,Any,AbstractCosmoParams,Any}, l, sf.itp, kgrid, 𝕡, bg) :
               which == 2 ? invoke(clte, Tuple{Any,Any,Any,Any,AbstractCosmoParams,Any}, l, sf.itp, sf_P.itp, kgrid, 𝕡, bg) :
                            invoke(clee, Tuple{Any,Any,Any,AbstractCosmoParams,Any}, l, sf_P.itp, kgrid, 𝕡, bg)
        return ℓ isa Integer ? f(ℓ) : map(f, ℓ)
    end
    r = project(ℓ isa Integer ? [ℓ] : ℓ, 𝕡, bg, sf, sf_P, spec)[which]
    ℓ isa Integer ? r[1] : r
end
cltt(ℓ, s::DeviceSourceGrid, kgrid, 𝕡::AbstractCosmoParams, bg) = explicit(1, ℓ, s, nothing, kgrid, 𝕡, bg)
clte(ℓ, s::DeviceSourceGrid, sP::DeviceSourceGrid, kgrid, 𝕡::AbstractCosmoParams, bg) = explicit(2, ℓ, s, sP, kgrid, 𝕡, bg)
clee(ℓ, sP::DeviceSourceGrid, kgrid, 𝕡::AbstractCosmoParams, bg) = explicit(3, ℓ, nothing, sP, kgrid, 𝕡, bg)

# ---- plin (src/spectra.jl:163-198) -----------------------------------------------------------------------------
"""plin for a vector of k: one batched bolt_plin call.  x = 0 runs the device epilogue; x ≠ 0 solves on the device, takes
perturb(x) from the history and evaluates the reference's own epilogue on it."""
function plin(ks::AbstractVector, 𝕡::AbstractCosmoParams{T}, bg, ih, n_q=15, ℓᵧ=50, ℓ_ν=50, ℓ_mν=20, x=0, reltol=1e-5; dev=Device()) where T
    n_q == length(bg.quad_pts) || error("n_q must match the background's quadrature")
    o = adaptive(ℓᵧ, ℓ_ν, ℓ_mν, reltol, 1e-6)
    if x != 0
        _, _, hist, status, nsteps, nd, n = solve(𝕡, bg, ih, ks, dev, o; sources=false, history=true)
        nx = length(bg.x_grid); u = reshape(unflat(T, hist, nd), n, nx, length(ks))
        return [plin_epilogue(DeviceSolution{T}(collect(Float64, bg.x_grid), u[:, :, i], Int(status[i]), Int(nsteps[i]))(x),
                              ks[i], 𝕡, bg, x, n_q, ℓᵧ, ℓ_ν, ℓ_mν) for i in eachindex(ks)]
    end
    with_cosmo(dev, 𝕡, bg, ih) do ctx, c, nd
        k = plain(ks); nk = length(k)
        pk = zeros(Float64, nd * nk); status = zeros(Int32, nk); nsteps = zeros(Int64, nk)
        GC.@preserve k pk status nsteps check(ctx, ccall((:bolt_plin, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}), ctx, c, k, nk, o, pk, status, nsteps))
        warn_status(status, "bolt_plin")
        unflat(T, pk, nd)
    end
end
"""Scalar k on the device (the reference's own scalar method, src/spectra.jl:163, stays the CPU path)."""
plin(k::Real, 𝕡::AbstractCosmoParams, bg, ih, dev::Device, args...) = plin([k], 𝕡, bg, ih, args...; dev=dev)[1]

# src/spectra.jl:170-197 on a state vector `results = perturb(x)`
function plin_epilogue(results, k, 𝕡, bg, x, n_q, ℓᵧ, ℓ_ν, ℓ_mν)
    o = 2(ℓᵧ + 1) + (ℓ_ν + 1)
    ℳρ, _ = Bolt.ρ_σ(results[o+1:o+n_q], results[o+2n_q+1:o+3n_q], bg, exp(x), 𝕡) ./ bg.ρ₀ℳ(x)
    ℳθ = k * Bolt.θ(results[o+n_q+1:o+2n_q], bg, exp(x), 𝕡) ./ bg.ρ₀ℳ(x)
    s = o + (ℓ_mν + 1) * n_q
    δcN, δbN, vcN, vbN = results[s+2], results[s+4], results[s+3], results[s+5]
    vmνN = -ℳθ / k
    Tγ = (15 / π^2 * bg.ρ_crit * 𝕡.Ω_r)^(1 / 4)
    νfac = (90 * 1.2020569 / (11 * π^4)) * (𝕡.Ω_r * 𝕡.h^2 / Tγ) * ((𝕡.N_ν / 3)^(3 / 4))
    Ω_ν = 𝕡.Σm_ν * νfac / 𝕡.h^2
    Ωm = 𝕡.Ω_c + 𝕡.Ω_b + Ω_ν
    δc = δcN - 3bg.ℋ(x) * vcN / k; δb = δbN - 3bg.ℋ(x) * vbN / k
    δmν = ℳρ - 3bg.ℋ(x) * vmνN / k
    δm = (𝕡.Ω_c * δc + 𝕡.Ω_b * δb + Ω_ν * δmν) / Ωm
    (2π^2 / k^3) * δm^2 * 𝕡.A * (k / 0.05)^(𝕡.n - 1)
end

# ---- batches and multi-GPU ----------------------------------------------------------------------------------------
"""TT/TE/EE for a batch of parameter sets (Float64 only): one bolt_spectra_batch call -- all hierarchy solves in ONE launch.
`batch` is a vector of (𝕡, bg, ih); every cosmology gets `quadratic_k(0.1H₀, 1000H₀, nk)` as in examples/basic_usage.jl."""
function spectra_batch(ℓ⃗, batch::AbstractVector; nk=2000, ℓᵧ=8, reltol=1e-11, dev=Device())
    ctx = context(dev.ordinal); ncos = length(batch); nℓ = length(ℓ⃗)
    H₀ = [Float64(bg.H₀) for (_, bg, _) in batch]
    k = reduce(hcat, [collect(quadratic_k(0.1h0, 1000h0, nk)) for h0 in H₀])          # [nk, ncos] = C [ncos][nk]
    kd_min = 0.01 .* H₀; kd_max = 1000 .* H₀
    bg1 = batch[1][2]; ix_start = findfirst(bg1.x_grid .> -8) - 1
    tt = zeros(Float64, nℓ, ncos); te = similar(tt); ee = similar(tt)
    status = zeros(Int32, nk, ncos); nsteps = zeros(Int64, nk, ncos)
    o = adaptive(ℓᵧ, 8, 10, reltol, 1e-6)
    ells = Int32.(collect(ℓ⃗))
    nested(i, cs) = i > ncos ? begin
        GC.@preserve cs k ells kd_min kd_max tt te ee status nsteps check(ctx, ccall((:bolt_spectra_batch, lib), Cint,
            (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Int32}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Cint,
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
            ctx, cs, ncos, k, nk, o, ells, nℓ, kd_min, kd_max, 5000, ix_start, tt, te, ee, status, nsteps))
    end : with_cosmo(dev, batch[i]...) do _, c, _; nested(i + 1, push!(copy(cs), c)); end      # every upload freed on the way out
    nested(1, Ptr{Cvoid}[])
    warn_status(status, "bolt_spectra_batch")
    tt, te, ee
end

"""Multi-GPU (one Julia process per GPU, e.g. under MPI.jl): rank 0 calls `comm_unique_id`, the host broadcasts the 128 bytes
(`MPI.Bcast!`), every rank calls `comm_init`; `spectra_sharded` is then bolt_spectra with the k-modes and multipoles of ONE
cosmology sharded over the ranks (one ncclAllGather of the sources, one ncclAllReduce of C_ℓ inside the library)."""
function comm_unique_id(dev::Device=Device())
    id = zeros(UInt8, 128); ctx = context(dev.ordinal)
    GC.@preserve id check(ctx, ccall((:bolt_comm_unique_id, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, id))
    id
end
function comm_init(rank::Integer, nranks::Integer, id::Vector{UInt8}, dev::Device=Device())
    ctx = context(dev.ordinal)
    GC.@preserve id check(ctx, ccall((:bolt_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), ctx, rank, nranks, id))
end
comm_free(dev::Device=Device()) = ccall((:bolt_comm_free, lib), Cint, (Ptr{Cvoid},), context(dev.ordinal))

function spectra_sharded(ℓ⃗, 𝕡::AbstractCosmoParams{T}, bg, ih, k_grid; ℓᵧ=8, reltol=1e-11, dev=Device()) where T
    with_cosmo(dev, 𝕡, bg, ih) do ctx, c, nd
        k = plain(k_grid); nk = length(k); ells = Int32.(collect(ℓ⃗)); nℓ = length(ells)
        tt = zeros(Float64, nd * nℓ); te = similar(tt); ee = similar(tt)
        status = zeros(Int32, nk); nsteps = zeros(Int64, nk); nreject = zeros(Int64, nk)
        kmin, kmax, nkd = default_dense(bg); ix_start = findfirst(bg.x_grid .> -8) - 1
        o = adaptive(ℓᵧ, 8, 10, reltol, 1e-6)
        GC.@preserve k ells tt te ee status nsteps nreject check(ctx, ccall((:bolt_spectra_sharded, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Int32}, Cint, Cdouble, Cdouble, Cint, Cint,
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
            ctx, c, k, nk, o, ells, nℓ, kmin, kmax, nkd, ix_start, tt, te, ee, status, nsteps, nreject))
        warn_status(status, "bolt_spectra_sharded")
        unflat(T, tt, nd), unflat(T, te, nd), unflat(T, ee, nd)
    end
end

"""`plin` of ONE cosmology over the ranks of the communicator (comm_init): K1 and the P(k) epilogue on each rank's shard, one
ncclAllGather; same result on every rank as `plin(ks, 𝕡, bg, ih, …)` (x = 0)."""
function plin_sharded(ks::AbstractVector, 𝕡::AbstractCosmoParams{T}, bg, ih, n_q=15, ℓᵧ=50, ℓ_ν=50, ℓ_mν=20, reltol=1e-5; dev=Device()) where T
    n_q == length(bg.quad_pts) || error("n_q must match the background's quadrature")
    o = adaptive(ℓᵧ, ℓ_ν, ℓ_mν, reltol, 1e-6)
    with_cosmo(dev, 𝕡, bg, ih) do ctx, c, nd
        k = plain(ks); nk = length(k)
        pk = zeros(Float64, nd * nk); status = zeros(Int32, nk); nsteps = zeros(Int64, nk)
        GC.@preserve k pk status nsteps check(ctx, ccall((:bolt_plin_sharded, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}), ctx, c, k, nk, o, pk, status, nsteps))
        warn_status(status, "bolt_plin_sharded")
        unflat(T, pk, nd)
    end
end

"""FFTLog on the device (src/util.jl:33-108): `plan_fftlog(r, μ, q, k₀r₀; kropt)` followed by `mul!` (inverse = false) or `ldiv!`.
Returns (y::Vector{ComplexF64}, k::Vector{Float64}).  N must be a power of two ≤ 4096."""
function fftlog(r::AbstractVector, a::AbstractVector, μ, q, k₀r₀=1.0; kropt=true, inverse=false, dev=Device())
    ctx = context(dev.ordinal); N = length(r)
    rr = Float64.(r); are = Float64.(real.(a)); aim = Float64.(imag.(a))
    y = zeros(Float64, 2, N); k = zeros(Float64, N); used = zeros(Float64, 1)
    GC.@preserve rr are aim y k used check(ctx, ccall((:bolt_fftlog, lib), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Cint, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx, rr, N, μ, q, k₀r₀, kropt ? 1 : 0, inverse ? 1 : 0, are, aim, y, k, used))
    complex.(y[1, :], y[2, :]), k
end

# ---- batched input tables (src/background.jl:104-128, src/ionization/recfast.jl:446-536,674-726), SURVEY 8f row n1 ------------------
"""`hostgen_batch(pars)`: the 12 spline-coefficient tables and 13 scalars of every cosmology in `pars` computed on the device in
one call (what `Background` + `RECFAST` + `IonizationHistory` compute on the host one at a time).  Returns
(tables[n_x+2, 12, ncos], scalars[13, ncos], status[ncos]) in the layout `bolt_cosmo_upload` reads (nd = 1)."""
function hostgen_batch(pars::AbstractVector{<:AbstractCosmoParams}; x0=-20.0, dx=0.01, n_x=2001, nq=15, dev=Device())
    ncos = length(pars)
    P = zeros(Float64, 9, ncos)
    for (i, 𝕡) in enumerate(pars)
        P[:, i] .= (𝕡.h, 𝕡.Ω_r, 𝕡.Ω_b, 𝕡.Ω_c, 𝕡.A, 𝕡.n, 𝕡.Y_p, 𝕡.N_ν, 𝕡.Σm_ν)
    end
    pts, wts = Bolt.gausslegendre(nq)
    pts = Float64.(pts); wts = Float64.(wts)
    tabs = zeros(Float64, n_x + 2, 12, ncos); sc = zeros(Float64, 13, ncos); st = zeros(Int32, ncos)
    rc = GC.@preserve P pts wts tabs sc st ccall((:bolt_hostgen_batch, lib), Cint,
        (Cint, Ptr{Float64}, Cint, Cdouble, Cdouble, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        dev.ordinal, P, ncos, x0, dx, n_x, pts, wts, nq, tabs, sc, st)
    rc == 0 || error("bolt_hostgen_batch: " * unsafe_string(ccall((:bolt_hostgen_last_error, lib), Cstring, ())))
    tabs, sc, st
end

# ---- Bessel-moment tables and the Filon rule (src/bessel/*.jl), SURVEY 8f row n2 ------------------------------------------------
moments_check(rc, what) = rc == 0 || error(what * ": " * unsafe_string(ccall((:bolt_moments_last_error, lib), Cstring, ())))

"""Device counterpart of `Bolt.MomentTable` (src/bessel/interpolator.jl:19-38): callable, returns the `order` moments at x."""
mutable struct DeviceMomentTable
    handle::Ptr{Cvoid}
    ν::Int
    order::Int
    kη_min::Float64
    kη_max::Float64
end
Bolt.getnu(t::DeviceMomentTable) = t.ν
Bolt.getorder(t::DeviceMomentTable) = t.order

"""`sph_bessel_interpolator(dev, ν, order, kη_min, kη_max, N; weniger_cut=50)` (src/bessel/interpolator.jl:67-80)."""
function Bolt.sph_bessel_interpolator(dev::Device, ν::Int, order, kη_min, kη_max, N::Int; weniger_cut=50)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    moments_check(ccall((:bolt_moment_table_create, lib), Cint, (Cint, Cint, Cint, Cdouble, Cdouble, Cint, Cdouble, Ref{Ptr{Cvoid}}),
        dev.ordinal, ν, order, kη_min, kη_max, N, weniger_cut, h), "bolt_moment_table_create")
    t = DeviceMomentTable(h[], ν, order, kη_min, kη_max)
    finalizer(x -> ccall((:bolt_moment_table_free, lib), Cvoid, (Ptr{Cvoid},), x.handle), t)
    t
end

function (t::DeviceMomentTable)(x::AbstractVector)
    xx = Float64.(x); out = zeros(Float64, t.order, length(xx))
    GC.@preserve xx out moments_check(ccall((:bolt_moment_table_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}),
        t.handle, xx, length(xx), out), "bolt_moment_table_eval")
    out
end
(t::DeviceMomentTable)(x::Real) = t([x])[:, 1]

"""Direct moments ∫₀ˣ t^p j_ν(t) dt on the device: method 0 small-argument evaluator (the role of `sph_j_moment_weniger_₁F₂`),
1 Lommel asymptotic form (`sph_j_moment_asymp`), 2 Maclaurin series (`sph_j_moment_maclaurin_₁F₂`); src/bessel/moments.jl:56-83."""
function sph_j_moments(x::AbstractVector, ν::Int, powers::AbstractVector; method=0, dev=Device())
    xx = Float64.(x); pp = Float64.(powers); out = zeros(Float64, length(pp), length(xx))
    GC.@preserve xx pp out moments_check(ccall((:bolt_sph_j_moments, lib), Cint,
        (Cint, Cint, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}),
        dev.ordinal, ν, length(pp), pp, method, xx, length(xx), out), "bolt_sph_j_moments")
    out
end

"""`integrate_sph_bessel_filon(f, f′, f″, k, a, b, itp)` for vectors of independent pieces (src/bessel/integrator.jl:7-20)."""
function Bolt.integrate_sph_bessel_filon(f::AbstractVector, f′::AbstractVector, f″::AbstractVector, k::AbstractVector,
                                         a::AbstractVector, b::AbstractVector, itp::DeviceMomentTable)
    n = length(f); v = map(x -> Float64.(x), (f, f′, f″, k, a, b)); out = zeros(Float64, n)
    GC.@preserve v out moments_check(ccall((:bolt_filon_pieces, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        itp.handle, n, v[1], v[2], v[3], v[4], v[5], v[6], out), "bolt_filon_pieces")
    out
end
Bolt.integrate_sph_bessel_filon(f::Real, f′::Real, f″::Real, k::Real, a::Real, b::Real, itp::DeviceMomentTable) =
    Bolt.integrate_sph_bessel_filon([f], [f′], [f″], [k], [a], [b], itp)[1]

"""The loop form (src/bessel/integrator.jl:25-38) batched: for every k the sum of the pieces between consecutive `nodes`, with the
quadratic's f, f′, f″ given at the nodes as (n_nodes, n_k) matrices.  One device block per k; each node's moments are evaluated once."""
function filon_chain(nodes::AbstractVector, f::AbstractMatrix, f′::AbstractMatrix, f″::AbstractMatrix, k::AbstractVector, itp::DeviceMomentTable)
    nn = length(nodes); nk = length(k)
    size(f) == (nn, nk) || error("f must be (n_nodes, n_k)")
    xs = Float64.(nodes); kk = Float64.(k); F = Float64.(f); F1 = Float64.(f′); F2 = Float64.(f″); out = zeros(Float64, nk)
    GC.@preserve xs kk F F1 F2 out moments_check(ccall((:bolt_filon_chain, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cfloat}),
        itp.handle, nk, nn, xs, F, F1, F2, kk, out, C_NULL), "bolt_filon_chain")
    out
end

end # module
