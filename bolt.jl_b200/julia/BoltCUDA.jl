# BoltCUDA.jl -- the reference-side binding of libbolt_cuda.so (include/bolt_cuda.h).
#
# Julia is not installed in the build image, so this file has never been executed there; it is kept short and
# is mirrored 1:1 by the Python ctypes driver (bolt.jl_b200/capi.py + api.py), which IS tested.  It adds methods
# to Bolt's own exported functions (src/Bolt.jl:8-9) so user scripts keep working unchanged:
#
#     using Bolt, BoltCUDA
#     sf   = source_grid(𝕡, bg, ih, ks, BoltCUDA.Device())           # src/spectra.jl:6
#     sf_P = source_grid_P(𝕡, bg, ih, ks, BoltCUDA.Device())         # src/spectra.jl:25 (served from the same solve)
#     Cᵀᵀ  = cltt(ℓs, 𝕡, bg, ih, sf)                                  # src/spectra.jl:147 -> ONE bolt_project call
#     pL   = plin(ks, 𝕡, bg, ih)                                      # vector method; scalar method = 1-element batch
module BoltCUDA

using Bolt
import Bolt: source_grid, source_grid_P, cltt, clte, clee, plin, boltsolve, AbstractCosmoParams
using ForwardDiff

const lib = get(ENV, "BOLT_CUDA_LIB", joinpath(@__DIR__, "..", "csrc", "libbolt_cuda.so"))

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

check(ctx, rc) = rc == 0 || error("libbolt_cuda error $rc: ", unsafe_string(ccall((:bolt_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))

# one context per Julia thread (the reference calls plin/cltt from many threads: ThreadPools, spectra.jl:10,149)
const ctxs = Dict{Tuple{Int,Int},Ptr{Cvoid}}()
function context(dev::Int)
    get!(ctxs, (Threads.threadid(), dev)) do
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bolt_init, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), dev, r)
        rc == 0 || error("bolt_init failed ($rc): no usable CUDA device; there is no CPU fallback")
        r[]
    end
end

# A Vector{Dual{Tag,Float64,N}} is bit-identical to a (1+N) x len Float64 matrix, value first.
flat(v::AbstractVector{Float64}) = (v, 1)
flat(v::AbstractVector{ForwardDiff.Dual{T,Float64,N}}) where {T,N} = (reinterpret(Float64, v), 1 + N)
coefs(itp) = vec(collect(itp.itp.coefs))            # spline(f, x_grid) = scale(interpolate(...)), src/util.jl:11

"""Upload what Background and IonizationHistory computed on the host (src/background.jl:104-128, recfast.jl:674-726)."""
function upload(ctx, 𝕡::AbstractCosmoParams{T}, bg, ih) where T
    sc = T[𝕡.h, 𝕡.Ω_r, 𝕡.Ω_b, 𝕡.Ω_c, 𝕡.A, 𝕡.n, 𝕡.Y_p, 𝕡.N_ν, 𝕡.Σm_ν, bg.H₀, bg.η₀, bg.ρ_crit, bg.Ω_Λ]
    tabs = vcat((coefs(t) for t in (bg.ℋ, bg.ℋ′, bg.ℋ′′, bg.η, bg.ρ₀ℳ, ih.τ, ih.τ′, ih.τ′′, ih.g̃, ih.g̃′, ih.g̃′′, ih.csb²))...)
    scf, nd = flat(sc); tbf, _ = flat(tabs)
    qp = Float64.(ForwardDiff.value.(bg.quad_pts)); qw = Float64.(ForwardDiff.value.(bg.quad_wts))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve scf tbf qp qw begin
        d = CosmoDesc(1, nd, length(bg.x_grid), length(qp), first(bg.x_grid), step(bg.x_grid),
                      pointer(scf), pointer(qp), pointer(qw), pointer(tbf))
        check(ctx, ccall((:bolt_cosmo_upload, lib), Cint, (Ptr{Cvoid}, Ref{CosmoDesc}, Ref{Ptr{Cvoid}}), ctx, d, out))
    end
    out[], nd
end

unflat(::Type{Float64}, a, nd) = vec(a)
unflat(::Type{D}, a, nd) where {D<:ForwardDiff.Dual} = collect(reinterpret(D, vec(a)))

function grids(𝕡::AbstractCosmoParams{T}, bg, ih, k_grid, dev::Device; ℓᵧ=8, reltol=1e-11) where T
    ctx = context(dev.ordinal); c, nd = upload(ctx, 𝕡, bg, ih)
    k = Float64.(ForwardDiff.value.(k_grid)); nk = length(k); nx = length(bg.x_grid)
    S_T = zeros(Float64, nd * nx * nk); S_P = similar(S_T); status = zeros(Int32, nk); nsteps = zeros(Int64, nk)
    o = Opts(ℓᵧ, 8, 10, 0, reltol, 1e-6, 0.0, 0, 0, 0)           # Hierarchy defaults, src/perturbations.jl:20-21,25
    GC.@preserve k S_T S_P status nsteps check(ctx, ccall((:bolt_solve, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
        ctx, c, k, nk, o, S_T, S_P, C_NULL, C_NULL, status, nsteps, C_NULL))
    ccall((:bolt_cosmo_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, c)
    any(!=(0), status) && @warn "bolt_solve: some k-modes did not finish cleanly" status   # the reference never checks retcode
    mk(S) = Bolt.LinearInterpolation((bg.x_grid, k_grid), reshape(unflat(T, S, nd), nx, nk), extrapolation_bc = Bolt.Line())
    mk(S_T), mk(S_P)
end

const pair = Ref{Any}(nothing)      # the sibling source grid of the last solve
function source_grid(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device; kw...)
    sT, sP = grids(𝕡, bg, ih, k_grid, dev; kw...); pair[] = (objectid(k_grid), sP); sT
end
function source_grid_P(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device; kw...)
    p = pair[]
    p !== nothing && p[1] == objectid(k_grid) && (pair[] = nothing; return p[2])
    grids(𝕡, bg, ih, k_grid, dev; kw...)[2]
end

"""cltt / clte / clee for a vector of multipoles: ONE bolt_project call instead of qmap over ℓ (src/spectra.jl:147-160)."""
function project(ℓ⃗, 𝕡::AbstractCosmoParams{T}, bg, ih, sf, sf_P; dev=Device()) where T
    ctx = context(dev.ordinal); c, nd = upload(ctx, 𝕡, bg, ih)
    ref = sf === nothing ? sf_P : sf
    k = Float64.(ForwardDiff.value.(ref.itp.knots[2])); nk = length(k); nℓ = length(ℓ⃗)
    g(s) = s === nothing ? Ptr{Float64}(C_NULL) : pointer(flat(vec(s.itp.coefs))[1])
    tt = zeros(Float64, nd * nℓ); te = similar(tt); ee = similar(tt)
    ix_start = findfirst(bg.x_grid .> -8) - 1                     # src/spectra.jl:86, 0-based across the ABI
    H₀ = ForwardDiff.value(bg.H₀)
    GC.@preserve sf sf_P k tt te ee check(ctx, ccall((:bolt_project, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Int32}, Cint, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx, c, g(sf), g(sf_P), k, nk, Int32.(collect(ℓ⃗)), nℓ, 0.01H₀, 1000H₀, 5000, ix_start, tt, te, ee))   # src/spectra.jl:133
    ccall((:bolt_cosmo_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, c)
    unflat(T, tt, nd), unflat(T, te, nd), unflat(T, ee, nd)
end
cltt(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf) = project(ℓ⃗, 𝕡, bg, ih, sf, nothing)[1]
clte(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf, sf_P) = project(ℓ⃗, 𝕡, bg, ih, sf, sf_P)[2]
clee(ℓ⃗::AbstractVector, 𝕡::AbstractCosmoParams, bg, ih, sf_P) = project(ℓ⃗, 𝕡, bg, ih, nothing, sf_P)[3]

"""plin for a vector of k: one batched bolt_plin call (src/spectra.jl:163-198; x = 0)."""
function plin(ks::AbstractVector, 𝕡::AbstractCosmoParams{T}, bg, ih, n_q=15, ℓᵧ=50, ℓ_ν=50, ℓ_mν=20, x=0, reltol=1e-5; dev=Device()) where T
    x == 0 || error("the device evaluates plin at x = 0")
    ctx = context(dev.ordinal); c, nd = upload(ctx, 𝕡, bg, ih)
    k = Float64.(ForwardDiff.value.(ks)); nk = length(k)
    pk = zeros(Float64, nd * nk); status = zeros(Int32, nk); nsteps = zeros(Int64, nk)
    o = Opts(ℓᵧ, ℓ_ν, ℓ_mν, 0, reltol, 1e-6, 0.0, 0, 0, 0)
    GC.@preserve k pk status nsteps check(ctx, ccall((:bolt_plin, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}), ctx, c, k, nk, o, pk, status, nsteps))
    ccall((:bolt_cosmo_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, c)
    unflat(T, pk, nd)
end

"""TT/TE/EE for a batch of parameter sets (Float64 only): one bolt_spectra_batch call -- all hierarchy solves in ONE launch.
`batch` is a vector of (𝕡, bg, ih); every cosmology gets `quadratic_k(0.1H₀, 1000H₀, nk)` as in examples/basic_usage.jl."""
function spectra_batch(ℓ⃗, batch::AbstractVector; nk=2000, ℓᵧ=8, reltol=1e-11, dev=Device())
    ctx = context(dev.ordinal); ncos = length(batch); nℓ = length(ℓ⃗)
    cs = [upload(ctx, 𝕡, bg, ih)[1] for (𝕡, bg, ih) in batch]
    H₀ = [Float64(bg.H₀) for (_, bg, _) in batch]
    k = reduce(hcat, [collect(quadratic_k(0.1h0, 1000h0, nk)) for h0 in H₀])          # [nk, ncos] = C [ncos][nk]
    kd_min = 0.01 .* H₀; kd_max = 1000 .* H₀
    bg1 = batch[1][2]; ix_start = findfirst(bg1.x_grid .> -8) - 1
    tt = zeros(Float64, nℓ, ncos); te = similar(tt); ee = similar(tt)
    status = zeros(Int32, nk, ncos); nsteps = zeros(Int64, nk, ncos)
    o = Opts(ℓᵧ, 8, 10, 0, reltol, 1e-6, 0.0, 0, 0, 0)
    GC.@preserve cs k tt te ee status nsteps check(ctx, ccall((:bolt_spectra_batch, lib), Cint,
        (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Float64}, Cint, Ref{Opts}, Ptr{Int32}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Cint,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
        ctx, cs, ncos, k, nk, o, Int32.(collect(ℓ⃗)), nℓ, kd_min, kd_max, 5000, ix_start, tt, te, ee, status, nsteps))
    foreach(c -> ccall((:bolt_cosmo_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, c), cs)
    any(status .∉ Ref((0, 4))) && @warn "some k-modes did not finish" count(status .∉ Ref((0, 4)))
    tt, te, ee
end

end # module
