# AugCUDA.jl — the Julia glue a maintainer of AugmentedGPLikelihoods.jl would add so that the reference's
# own verbs dispatch to libaugcuda.so for device-resident data.
#
# STATUS: written to the C ABI of include/augcuda.h, NOT executed: there is no Julia in the build image
# (INTEGRATION.md).  Everything the tests and bench.py exercise goes through the same ABI from Python
# (augmentedgplikelihoods.jl_b200/api.py), function for function.  What CAN be checked without Julia is checked
# statically by tests/test_abi_cpu.py: every ccall names a declared symbol and its argument-type tuple agrees,
# type class by type class, with the C prototype in include/augcuda.h and with the ctypes table.
#
# Design: Julia multiple dispatch is the reference's plugin mechanism (src/AugmentedGPLikelihoods.jl:18-30),
# so the "plugin" is a set of METHODS of the reference's generic functions, specialised on a device array
# type.  qΩ stays a MeasureTheory.For over a TupleVector and Ω stays a TupleVector — only their field
# arrays are device vectors — so `only(qΩ.inds).c`, `Ω.ω`, `length(Ω)` keep working.
module AugCUDA

using AugmentedGPLikelihoods
using AugmentedGPLikelihoods: AbstractLikelihood, BijectiveSimplexLink, LogisticSoftMaxLink,
    ScaledLogistic, InvScaledLogistic, LaplaceLikelihood, StudentTLikelihood
using GPLikelihoods: BernoulliLikelihood, PoissonLikelihood, NegativeBinomialLikelihood, NBParamFailure,
    HeteroscedasticGaussianLikelihood, CategoricalLikelihood, LogisticLink
using Distributions: Normal, mean, var
using MeasureTheory: For
using TupleVectors: TupleVector
using Random: AbstractRNG

const AGPL = AugmentedGPLikelihoods
const lib = "libaugcuda"

# ---------------------------------------------------------------- C ABI mirrors (include/augcuda.h)
struct AugLik               # typedef struct aug_lik
    kind::Int32
    nlatent::Int32
    r_is_int::Int32
    flags::Int32            # AUG_LIK_FAITHFUL_QUIRKS = 1 (include/augcuda.h)
    p::NTuple{4,Float64}
    logtheta::Ptr{Float64}  # host pointer
end

const BERNOULLI, NEGBIN, POISSON, LAPLACE, STUDENTT, HETERO, CAT_BIJ, CAT = Int32.(0:7)
const S_ELT, S_KL, S_EAUGLL, S_LOGTILT, S_LOGPRIOR, S_AUGLL = 1:6   # 1-based slots of the scalar block

mutable struct Context
    h::Ptr{Cvoid}
    device::Int
end

function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:aug_strerror, lib), Cstring, (Int32,), rc))
    # mirrors the reference's error(...) / ArgumentError behaviour (SURVEY §5)
    rc == -3 ? error(msg) : throw(ErrorException("libaugcuda: $msg (rc=$rc)"))
end

function Context(device::Integer=0; stream::Ptr{Cvoid}=C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:aug_ctx_create, lib), Int32, (Ref{Ptr{Cvoid}}, Int32, Ptr{Cvoid}), h, device, stream))
    ctx = Context(h[], device)
    finalizer(c -> ccall((:aug_ctx_destroy, lib), Int32, (Ptr{Cvoid},), c.h), ctx)
    return ctx
end

const CTX = Ref{Context}()
ctx() = isassigned(CTX) ? CTX[] : (CTX[] = Context())

# ---------------------------------------------------------------- device containers
"Device vector owned by Julia (allocated with aug_malloc) or borrowed (e.g. from CUDA.jl: pass its pointer)."
mutable struct AugDeviceVector{T} <: AbstractVector{T}
    ptr::Ptr{T}
    len::Int
    owner::Bool
    parent::Any      # views keep the owning vector reachable (its finalizer frees the allocation)
end
AugDeviceVector{T}(ptr::Ptr{T}, len::Integer, owner::Bool) where {T} = AugDeviceVector{T}(ptr, len, owner, nothing)
"non-owning view of `n` elements of `v` starting at element offset `off`; keeps `v` alive"
subvector(v::AugDeviceVector{T}, off::Integer, n::Integer) where {T} =
    AugDeviceVector{T}(v.ptr + off * sizeof(T), n, false, v)
"one vector per latent out of a latent-major [nl][n] block (the tuple the reference's verbs return)"
split_latents(v::AugDeviceVector, n::Integer, nl::Integer) = ntuple(j -> subvector(v, (j - 1) * n, n), nl)
split_latents(v::Vector, n::Integer, nl::Integer) = ntuple(j -> view(v, (j - 1) * n + 1:j * n), nl)
Base.size(v::AugDeviceVector) = (v.len,)
Base.getindex(::AugDeviceVector, ::Int) = error("scalar indexing of a device vector: copy it with Array(v)")
Base.pointer(v::AugDeviceVector) = v.ptr

function AugDeviceVector{T}(::UndefInitializer, n::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:aug_malloc, lib), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Csize_t), ctx().h, p, n * sizeof(T)))
    v = AugDeviceVector{T}(Ptr{T}(p[]), n, true, nothing)
    finalizer(x -> x.owner && ccall((:aug_free, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, x.ptr), v)
    return v
end
function AugDeviceVector(a::Vector{T}) where {T}
    v = AugDeviceVector{T}(undef, length(a))
    check(ccall((:aug_memcpy_h2d, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                ctx().h, v.ptr, a, sizeof(a)))
    return v
end
function Base.Array(v::AugDeviceVector{T}) where {T}
    a = Vector{T}(undef, v.len)
    check(ccall((:aug_memcpy_d2h, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                ctx().h, a, v.ptr, sizeof(a)))
    return a
end

"qf::AbstractVector{<:Normal} on the device: struct-of-arrays of mean and VARIANCE (utils.jl:1-7 reads only these)."
struct DeviceNormals <: AbstractVector{Normal{Float64}}
    μ::AugDeviceVector{Float64}
    σ²::AugDeviceVector{Float64}
    ld::Int                       # leading dimension for the 2-latent heteroscedastic layout, else 0
end
DeviceNormals(μ::AugDeviceVector{Float64}, σ²::AugDeviceVector{Float64}) = DeviceNormals(μ, σ², 0)
Base.size(q::DeviceNormals) = (q.ld == 0 ? q.μ.len : q.ld,)

"rng argument: counter-based Philox4x32-10 stream (seed, offset); every sampling verb consumes one tick."
mutable struct AugPhilox <: AbstractRNG
    seed::UInt64
    offset::UInt64
end
const GLOBAL_PHILOX = AugPhilox(0x243f6a8885a308d3, 0)
seed!(r::AugPhilox) = check(ccall((:aug_ctx_seed, lib), Int32, (Ptr{Cvoid}, UInt64, UInt64), ctx().h, r.seed, r.offset))
function sync_offset!(r::AugPhilox)
    o = Ref{UInt64}(0)
    check(ccall((:aug_ctx_get_offset, lib), Int32, (Ptr{Cvoid}, Ref{UInt64}), ctx().h, o))
    r.offset = o[]
end

# ---------------------------------------------------------------- likelihood descriptors
desc(::BernoulliLikelihood{<:LogisticLink}) = AugLik(BERNOULLI, 1, 0, 0, (0.0, 0.0, 0.0, 0.0), C_NULL)
function desc(l::NegativeBinomialLikelihood{<:NBParamFailure})
    r = l.params.failures
    AugLik(NEGBIN, 1, r isa Integer ? 1 : 0, 0, (Float64(r), 0.0, 0.0, 0.0), C_NULL)
end
desc(l::PoissonLikelihood{<:ScaledLogistic}) = AugLik(POISSON, 1, 0, 0, (Float64(l.invlink.λ), 0.0, 0.0, 0.0), C_NULL)
desc(l::LaplaceLikelihood) = AugLik(LAPLACE, 1, 0, 0, (Float64(l.β), 0.0, 0.0, 0.0), C_NULL)
desc(l::StudentTLikelihood) = AugLik(STUDENTT, 1, 0, 0, (Float64(l.ν), Float64(l.σ), 0.0, 0.0), C_NULL)
desc(l::HeteroscedasticGaussianLikelihood{<:InvScaledLogistic}) =
    AugLik(HETERO, 2, 0, 0, (Float64(l.invlink.λ), 0.0, 0.0, 0.0), C_NULL)
# Categorical: the logθ vector must stay alive during the call (GC.@preserve in `withdesc`)
logθ(l::CategoricalLikelihood{<:BijectiveSimplexLink{<:LogisticSoftMaxLink}}) = l.invlink.link.logθ
logθ(l::CategoricalLikelihood{<:LogisticSoftMaxLink}) = l.invlink.logθ
function withdesc(f, l::CategoricalLikelihood)
    θ = convert(Vector{Float64}, logθ(l))
    kind = l.invlink isa BijectiveSimplexLink ? CAT_BIJ : CAT
    GC.@preserve θ f(Ref(AugLik(kind, AGPL.nlatent(l), 0, 0, (0.0, 0.0, 0.0, 0.0), pointer(θ))))
end
withdesc(f, l::AbstractLikelihood) = f(Ref(desc(l)))

const DV = AugDeviceVector
state(qΩ::For) = only(qΩ.inds)                                   # the SoA of variational parameters
s0(φ) = haskey(φ, :c) ? φ.c : haskey(φ, :μ) ? φ.μ : φ.β          # bernoulli.jl:7-11 ... studentt.jl:39-44
s1(φ) = haskey(φ, :λ) ? φ.λ : haskey(φ, :p) ? φ.p : nothing
s2(φ) = haskey(φ, :ψ) ? φ.ψ : haskey(φ, :y) ? φ.y : nothing
ptr(::Nothing) = C_NULL
ptr(v::DV) = Ptr{Cvoid}(v.ptr)

# ---------------------------------------------------------------- variational verbs
# aux_posterior!(qΩ, lik, y, qf)                       -> aug_aux_posterior           (a5)
function AGPL.aux_posterior!(qΩ::For, lik::AbstractLikelihood, y::DV, qf::DeviceNormals)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_aux_posterior, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx().h, d, length(qΩ), y.ptr, qf.μ.ptr, qf.σ².ptr, qf.ld, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ))))
    end
    return qΩ                                                    # bernoulli.jl:24
end

# expected_auglik_potential_and_precision(lik, qΩ, y[, qf]) -> aug_expected_potential_precision   (a7)
function AGPL.expected_auglik_potential_and_precision(lik::AbstractLikelihood, qΩ::For, y::DV,
                                                      qf::Union{Nothing,DeviceNormals}=nothing)
    n, nl = length(qΩ), AGPL.nlatent(lik)
    β, γ = DV{Float64}(undef, n * nl), DV{Float64}(undef, n * nl)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_expected_potential_precision, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, y.ptr, qf === nothing ? C_NULL : qf.μ.ptr, qf === nothing ? 0 : qf.ld,
                    ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ)), β.ptr, γ.ptr, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end
AGPL.expected_auglik_potential(lik::AbstractLikelihood, qΩ::For, y::DV, qf=nothing) =
    first(AGPL.expected_auglik_potential_and_precision(lik, qΩ, y, qf))
AGPL.expected_auglik_precision(lik::AbstractLikelihood, qΩ::For, y::DV, qf=nothing) =
    last(AGPL.expected_auglik_potential_and_precision(lik, qΩ, y, qf))

# expected_logtilt / aux_kldivergence / expected_aug_loglik -> aug_expected_elbo_terms (a9, a11, a13)
function elbo_terms(lik, qΩ::For, y::DV, qf::DeviceNormals)
    sc = DV{Float64}(undef, 8)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_expected_elbo_terms, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                    ctx().h, d, length(qΩ), y.ptr, qf.μ.ptr, qf.σ².ptr, qf.ld, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ)), sc.ptr))
    end
    return Array(sc)          # multi-GPU callers run aug_allreduce_scalars on `sc` first (INTEGRATION.md)
end
AGPL.expected_logtilt(lik::AbstractLikelihood, qΩ::For, y::DV, qf::DeviceNormals) = elbo_terms(lik, qΩ, y, qf)[S_ELT]
AGPL.expected_aug_loglik(lik::AbstractLikelihood, qΩ::For, y::DV, qf::DeviceNormals) = elbo_terms(lik, qΩ, y, qf)[S_EAUGLL]
# aux_kldivergence(lik, qΩ, y) has no qf in the reference; the KL kernels only read the state and y, so a
# zero qf of the right length is passed (the heteroscedastic prior needs the real qf: use elbo_terms).

# The fused call of a CAVI iteration (examples/bernoulli/script.jl:29-39): aux_posterior!(qΩ, lik, y, qf) +
# expected_auglik_potential_and_precision + the expected_logtilt / aux_kldivergence sums in ONE pass -> aug_cavi_step
function cavi_step!(qΩ::For, lik::AbstractLikelihood, y::DV, qf::DeviceNormals)
    n, nl = length(qΩ), AGPL.nlatent(lik)
    β, γ = DV{Float64}(undef, n * nl), DV{Float64}(undef, n * nl)
    sc = DV{Float64}(undef, 8)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_cavi_step, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                    ctx().h, d, n, y.ptr, qf.μ.ptr, qf.σ².ptr, qf.ld, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ)), β.ptr, γ.ptr, n,
                    sc.ptr))
    end
    return qΩ, split_latents(β, n, nl), split_latents(γ, n, nl), Array(sc)      # sc[S_ELT], sc[S_KL], sc[S_EAUGLL]
end

# ---------------------------------------------------------------- sampling verbs
# aux_sample!(rng, Ω, lik, y, f)                      -> aug_aux_sample                 (a14-a19)
function AGPL.aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::AbstractLikelihood, y::DV, f::DV; i0::Integer=0)
    seed!(rng)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    withdesc(lik) do d
        check(ccall((:aug_aux_sample, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, length(y) ÷ max(1, AGPL.nlatent(lik) * (lik isa CategoricalLikelihood)), i0,
                    y.ptr, f.ptr, lik isa HeteroscedasticGaussianLikelihood ? length(f) ÷ 2 : 0, Ω.ω.ptr, ptr(nv)))
    end
    sync_offset!(rng)
    return Ω                                                     # generic.jl:11
end
AGPL.aux_sample!(Ω::TupleVector, lik::AbstractLikelihood, y::DV, f::DV) = AGPL.aux_sample!(GLOBAL_PHILOX, Ω, lik, y, f)

# init_aux_variables(rng, lik, n)                     -> aug_init_aux_variables          (a4)
function AGPL.init_aux_variables(rng::AugPhilox, lik::AbstractLikelihood, n::Int; i0::Integer=0)
    seed!(rng)
    m = lik isa CategoricalLikelihood ? n * AGPL.nlatent(lik) : n
    ω = DV{Float64}(undef, m)
    needs_n = lik isa Union{PoissonLikelihood,HeteroscedasticGaussianLikelihood,CategoricalLikelihood}
    nv = needs_n ? DV{Int64}(undef, m) : nothing
    withdesc(lik) do d
        check(ccall((:aug_init_aux_variables, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Float64}, Ptr{Int64}), ctx().h, d, n, i0, ω.ptr, ptr(nv)))
    end
    sync_offset!(rng)
    return needs_n ? TupleVector((; ω, n=nv)) : TupleVector((; ω))
end

# auglik_potential_and_precision(lik, Ω, y[, f])      -> aug_potential_precision         (a20)
function AGPL.auglik_potential_and_precision(lik::AbstractLikelihood, Ω::TupleVector, y::DV, f::Union{Nothing,DV}=nothing)
    nl = AGPL.nlatent(lik)
    n = length(Ω.ω) ÷ (lik isa CategoricalLikelihood ? nl : 1)
    β, γ = DV{Float64}(undef, n * nl), DV{Float64}(undef, n * nl)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    withdesc(lik) do d
        check(ccall((:aug_potential_precision, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64},
                     Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, y.ptr, f === nothing ? C_NULL : f.ptr, f === nothing ? 0 : length(f) ÷ 2,
                    Ω.ω.ptr, ptr(nv), β.ptr, γ.ptr, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end

# logtilt / aug_loglik                                -> aug_sampled_loglik_terms         (a21-a24)
function sampled_terms(lik, Ω::TupleVector, y::DV, f::DV, with_prior::Bool)
    sc = DV{Float64}(undef, 8)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = lik isa CategoricalLikelihood ? length(Ω.ω) ÷ AGPL.nlatent(lik) : length(Ω.ω)
    withdesc(lik) do d
        check(ccall((:aug_sampled_loglik_terms, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Int32,
                     Ptr{Float64}),
                    ctx().h, d, n, y.ptr, f.ptr, lik isa HeteroscedasticGaussianLikelihood ? n : 0, Ω.ω.ptr, ptr(nv),
                    with_prior, sc.ptr))
    end
    return Array(sc)
end
AGPL.logtilt(lik::AbstractLikelihood, Ω::TupleVector, y::DV, f::DV) = sampled_terms(lik, Ω, y, f, false)[S_LOGTILT]
AGPL.aug_loglik(lik::AbstractLikelihood, Ω::TupleVector, y::DV, f::DV) = sampled_terms(lik, Ω, y, f, true)[S_AUGLL]

# ---------------------------------------------------------------- host `Vector` methods (the reference's own argument types)
# src/generic.jl:1-88 dispatches every verb on plain host vectors.  These methods keep those signatures — y::Vector,
# qf::AbstractVector{<:Normal}, f::Vector, Ω / qΩ with Vector fields — and bind to the *_host entry points: the
# library stages chunks through the GPU (H2D, kernel, D2H overlapped) and the results come back in host vectors.
# The only host-side work is the AoS → SoA split of qf (mean.(qf), var.(qf)): the ABI takes struct-of-arrays.
const HV = Vector
hptr(::Nothing) = C_NULL
hptr(v::Vector) = Ptr{Cvoid}(pointer(v))
moments(qf::AbstractVector{<:Normal}) = (convert(Vector{Float64}, mean.(qf)), convert(Vector{Float64}, var.(qf)))
# two-latent heteroscedastic qfg = (qf, qg) (heteroscedasticgaussian.jl:38): latent-major [2][n]
moments(qfg::Tuple) = (vcat((mean.(q) for q in qfg)...), vcat((var.(q) for q in qfg)...))
ldof(lik, n) = lik isa HeteroscedasticGaussianLikelihood ? n : 0

function AGPL.aux_posterior!(qΩ::For, lik::AbstractLikelihood, y::HV, qf)
    φ = state(qΩ); μ, σ² = moments(qf); n = length(qΩ)
    withdesc(lik) do d
        GC.@preserve y μ σ² φ check(ccall((:aug_aux_posterior_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx().h, d, n, hptr(y), μ, σ², ldof(lik, n), hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ))))
    end
    return qΩ
end

function AGPL.expected_auglik_potential_and_precision(lik::AbstractLikelihood, qΩ::For, y::HV, qf=nothing)
    n, nl = length(qΩ), AGPL.nlatent(lik)
    β, γ = Vector{Float64}(undef, n * nl), Vector{Float64}(undef, n * nl)
    φ = state(qΩ)
    μ = qf === nothing ? nothing : first(moments(qf))
    withdesc(lik) do d
        GC.@preserve y μ φ β γ check(ccall((:aug_expected_potential_precision_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, hptr(y), μ === nothing ? C_NULL : pointer(μ), ldof(lik, n),
                    hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ)), β, γ, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end

function elbo_terms(lik, qΩ::For, y::HV, qf)
    sc = zeros(Float64, 8); φ = state(qΩ); μ, σ² = moments(qf); n = length(qΩ)
    withdesc(lik) do d
        GC.@preserve y μ σ² φ check(ccall((:aug_expected_elbo_terms_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                    ctx().h, d, n, hptr(y), μ, σ², ldof(lik, n), hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ)), sc))
    end
    return sc
end
AGPL.expected_logtilt(lik::AbstractLikelihood, qΩ::For, y::HV, qf) = elbo_terms(lik, qΩ, y, qf)[S_ELT]
AGPL.expected_aug_loglik(lik::AbstractLikelihood, qΩ::For, y::HV, qf) = elbo_terms(lik, qΩ, y, qf)[S_EAUGLL]

"fused CAVI iteration on host vectors; `want_state = false` / `want_β = false` skip those outputs (and their D2H bytes)"
function cavi_step!(qΩ::For, lik::AbstractLikelihood, y::HV, qf; want_state::Bool=true, want_β::Bool=true)
    n, nl = length(qΩ), AGPL.nlatent(lik)
    β = want_β ? Vector{Float64}(undef, n * nl) : nothing
    γ = Vector{Float64}(undef, n * nl)
    sc = zeros(Float64, 8); φ = state(qΩ); μ, σ² = moments(qf)
    st = want_state ? (hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ))) : (C_NULL, C_NULL, C_NULL)
    withdesc(lik) do d
        GC.@preserve y μ σ² φ β γ check(ccall((:aug_cavi_step_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                    ctx().h, d, n, hptr(y), μ, σ², ldof(lik, n), st[1], st[2], st[3],
                    β === nothing ? C_NULL : pointer(β), γ, n, sc))
    end
    return qΩ, (β === nothing ? nothing : split_latents(β, n, nl)), split_latents(γ, n, nl), sc
end

function AGPL.aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::AbstractLikelihood, y::HV, f::HV; i0::Integer=0)
    seed!(rng)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = lik isa CategoricalLikelihood ? length(Ω.ω) ÷ AGPL.nlatent(lik) : length(Ω.ω)
    withdesc(lik) do d
        GC.@preserve y f Ω check(ccall((:aug_aux_sample_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, n, i0, hptr(y), f, ldof(lik, n), Ω.ω, nv === nothing ? C_NULL : pointer(nv)))
    end
    sync_offset!(rng)
    return Ω
end

function init_aux_variables_host(rng::AugPhilox, lik::AbstractLikelihood, n::Int; i0::Integer=0)
    seed!(rng)
    m = lik isa CategoricalLikelihood ? n * AGPL.nlatent(lik) : n
    ω = Vector{Float64}(undef, m)
    needs_n = lik isa Union{PoissonLikelihood,HeteroscedasticGaussianLikelihood,CategoricalLikelihood}
    nv = needs_n ? Vector{Int64}(undef, m) : nothing
    withdesc(lik) do d
        GC.@preserve ω nv check(ccall((:aug_init_aux_variables_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, n, i0, ω, nv === nothing ? C_NULL : pointer(nv)))
    end
    sync_offset!(rng)
    return needs_n ? TupleVector((; ω, n=nv)) : TupleVector((; ω))
end

function AGPL.auglik_potential_and_precision(lik::AbstractLikelihood, Ω::TupleVector, y::HV, f::Union{Nothing,HV}=nothing)
    nl = AGPL.nlatent(lik)
    n = length(Ω.ω) ÷ (lik isa CategoricalLikelihood ? nl : 1)
    β, γ = Vector{Float64}(undef, n * nl), Vector{Float64}(undef, n * nl)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    withdesc(lik) do d
        GC.@preserve y f Ω β γ check(ccall((:aug_potential_precision_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64},
                     Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, hptr(y), f === nothing ? C_NULL : pointer(f), ldof(lik, n),
                    Ω.ω, nv === nothing ? C_NULL : pointer(nv), β, γ, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end

function sampled_terms(lik, Ω::TupleVector, y::HV, f::HV, with_prior::Bool)
    sc = zeros(Float64, 8)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = lik isa CategoricalLikelihood ? length(Ω.ω) ÷ AGPL.nlatent(lik) : length(Ω.ω)
    withdesc(lik) do d
        GC.@preserve y f Ω check(ccall((:aug_sampled_loglik_terms_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Int32,
                     Ptr{Float64}),
                    ctx().h, d, n, hptr(y), f, ldof(lik, n), Ω.ω, nv === nothing ? C_NULL : pointer(nv),
                    with_prior, sc))
    end
    return sc
end
AGPL.logtilt(lik::AbstractLikelihood, Ω::TupleVector, y::HV, f::HV) = sampled_terms(lik, Ω, y, f, false)[S_LOGTILT]
AGPL.aug_loglik(lik::AbstractLikelihood, Ω::TupleVector, y::HV, f::HV) = sampled_terms(lik, Ω, y, f, true)[S_AUGLL]

# ---- multi-GPU (one Julia process per GPU, e.g. Distributed / MPI.jl) ------------------------------------------
# Peer-memory mailbox: the all-reduce of the 64-byte scalar block runs INSIDE the reducing kernels (include/augcuda.h).
# `allgather(bytes)` is the caller's transport (MPI.Allgather, a Distributed remotecall, ...): it must return the
# concatenation of every rank's 64-byte handle in rank order, and act as a barrier.
function init_p2p!(nranks::Integer, rank::Integer, allgather::Function; fused::Bool=true)
    h = Vector{UInt8}(undef, 64)
    check(ccall((:aug_comm_p2p_export, lib), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Ptr{Cvoid}}), ctx().h, h, C_NULL))
    all = allgather(h)::Vector{UInt8}
    length(all) == 64 * nranks || error("allgather must return nranks × 64 bytes")
    check(ccall((:aug_comm_p2p_attach, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx().h, nranks, rank, all))
    allgather(h)                       # barrier: nobody exchanges before everybody has mapped everybody
    fused && check(ccall((:aug_comm_set_fused, lib), Int32, (Ptr{Cvoid}, Int32), ctx().h, 1))
    return nothing
end
"split-phase exchange: the reducing kernels only publish, the gather rides in the next aux_sample! launch or in `flush!()`"
set_deferred!(on::Bool=true) = check(ccall((:aug_comm_set_deferred, lib), Int32, (Ptr{Cvoid}, Int32), ctx().h, on))
flush!() = check(ccall((:aug_comm_flush, lib), Int32, (Ptr{Cvoid},), ctx().h))

# ---- callers on either side of the path (SURVEY §8(f) rows 3-4) --------------------------------------------------
# opt_lik of examples/heteroscedasticgaussian/script.jl:41-51 for device arguments
function opt_lik(lik::HeteroscedasticGaussianLikelihood, qfg::DeviceNormals, y::DV; n_total::Integer=length(y))
    out = AugDeviceVector{Float64}(undef, 1)
    n = length(y)
    check(ccall((:aug_hetero_lambda_stats, lib), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                ctx().h, n, y.ptr, qfg.μ.ptr, qfg.σ².ptr, n, out.ptr))
    s = Array(out)[1]
    return HeteroscedasticGaussianLikelihood(InvScaledLogistic(max(n_total / (2s), lik.invlink.λ)))
end

# (l::LogisticSoftMaxLink)(f) row-wise on a device matrix stored class-fastest (categorical.jl:32-35)
function logisticsoftmax_rows(lik::CategoricalLikelihood, f::DV, n::Integer)
    K = lik.invlink isa BijectiveSimplexLink ? nlatent(lik) + 1 : nlatent(lik)
    out = AugDeviceVector{Float64}(undef, n * K)
    withdesc(lik) do d
        check(ccall((:aug_logisticsoftmax, lib), Int32, (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Float64}, Ptr{Float64}),
                    ctx().h, d, n, f.ptr, out.ptr))
    end
    return out
end

# ---- the sparse-GP steps either side of the path (SURVEY §8(f) rows 1-2) ---------------------------------------
# κ = K_Z \ K_ZX is an M×N device matrix stored column-major (what `AugDeviceVector` of length M*N holds), B = K_Z − S.
# marginals(post_u(x)) of examples/bernoulli/script.jl:32-33 for a device κ:
function sparse_marginals(κ::DV, M::Integer, m::DV, B::DV, kdiag::DV)
    N = length(κ) ÷ M
    μ = AugDeviceVector{Float64}(undef, N); σ² = AugDeviceVector{Float64}(undef, N)
    check(ccall((:aug_sparse_marginals, lib), Int32,
                (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx().h, N, M, κ.ptr, m.ptr, B.ptr, kdiag.ptr, μ.ptr, σ².ptr))
    return DeviceNormals(μ, σ²)
end
# docs/src/index.md:156-160: (K_Z⁻¹ + κ Diagonal(γ) κᵀ, κβ + K_Z⁻¹μ₀) as one [M*M + M] device vector
function sparse_precision_potential(κ::DV, M::Integer, γ::DV, β::DV; P0=nothing, r0=nothing)
    N = length(κ) ÷ M
    Pr = AugDeviceVector{Float64}(undef, M * M + M)
    check(ccall((:aug_sparse_precision_potential, lib), Int32,
                (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx().h, N, M, κ.ptr, γ.ptr, β.ptr, ptr(P0), ptr(r0), Pr.ptr))
    return Pr
end
# one CAVI iteration of cavi!(…) (examples/bernoulli/script.jl:29-39, sparse update) in ONE pass over κ:
# marginals → aux_posterior!(qΩ, lik, y, ·) → E[β], E[γ] (+ ELBO sums) → P, rhs.  Returns (Pr, scalars).
function sparse_cavi_sweep!(qΩ, lik::AbstractLikelihood, y::DV, κ::DV, M::Integer, m::DV, B::DV, kdiag::DV;
                            P0=nothing, r0=nothing)
    N = length(y)
    Pr = AugDeviceVector{Float64}(undef, M * M + M)
    sc = AugDeviceVector{Float64}(undef, 8)
    φ = state(qΩ)
    s = (ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ)))   # C_NULL where the likelihood has no such field
    withdesc(lik) do d
        check(ccall((:aug_sparse_cavi_sweep, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    ctx().h, d, N, M, y.ptr, κ.ptr, m.ptr, B.ptr, kdiag.ptr, C_NULL, C_NULL, s[1], s[2], s[3],
                    C_NULL, C_NULL, ptr(P0), ptr(r0), Pr.ptr, sc.ptr))
    end
    return Pr, Array(sc)
end

end # module
