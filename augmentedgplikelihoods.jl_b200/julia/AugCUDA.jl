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
#
# Dispatch: the reference specialises most verbs on the LIKELIHOOD type with loosely typed arrays, e.g.
# `aux_posterior!(qΩ, ::BernoulliLikelihood{<:LogisticLink}, ::AbstractVector, qf::AbstractVector{<:Normal})`
# (src/likelihoods/bernoulli.jl:17-22).  A method `(qΩ::For, lik::AbstractLikelihood, y::AugDeviceVector, …)` would be
# AMBIGUOUS with it (each is more specific in one argument) and Julia would throw at the call.  So the file has two
# layers: implementations `dev_*` / `host_*` (one ccall each, any likelihood `withdesc` knows), and a generated dispatch
# layer at the end that adds, for every likelihood type the reference dispatches on (REFERENCE_LIKELIHOODS, same aliases
# as the reference), methods whose likelihood argument is EXACTLY the reference's and whose other arguments are
# subtypes of the reference's — strictly more specific, never ambiguous.
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
s0(φ) = hasproperty(φ, :c) ? φ.c : hasproperty(φ, :μ) ? φ.μ : φ.β          # bernoulli.jl:7-11 ... studentt.jl:39-44
s1(φ) = hasproperty(φ, :λ) ? φ.λ : hasproperty(φ, :p) ? φ.p : nothing
s2(φ) = hasproperty(φ, :ψ) ? φ.ψ : hasproperty(φ, :y) ? φ.y : nothing
iscat(lik) = lik isa CategoricalLikelihood
"number of observations behind a field array: the Categorical fields are flat [n][nl] device vectors (class fastest)"
nobs(lik, v::AbstractVector) = iscat(lik) ? length(v) ÷ AGPL.nlatent(lik) : length(v)
nobs(lik, qΩ::For) = nobs(lik, s0(state(qΩ)))
ptr(::Nothing) = C_NULL
ptr(v::DV) = Ptr{Cvoid}(v.ptr)

# ---------------------------------------------------------------- variational verbs (implementations)
# init_aux_posterior(T, lik, n)                        -> aug_init_aux_posterior      (a3)
# The For closure (row of the SoA -> the law of that observation's auxiliary variables) is the reference's own
# (bernoulli.jl:7-11 ... categorical.jl:59-70), taken from a zero-length CPU instance; only the field arrays differ.
function dev_init_aux_posterior(lik::AbstractLikelihood, n::Integer)
    q0 = AGPL.init_aux_posterior(Float64, lik, 0)
    names = propertynames(only(q0.inds))                        # (:c,), (:y, :c, :λ), (:c, :λ, :ψ), (:y, :c, :p), ...
    m = iscat(lik) ? n * AGPL.nlatent(lik) : n
    arrays = map(names) do k
        k === :y ? (iscat(lik) ? DV{Bool}(undef, m) : DV{Int64}(undef, m)) : DV{Float64}(undef, m)
    end
    φ = TupleVector(NamedTuple{names}(arrays))
    withdesc(lik) do d
        check(ccall((:aug_init_aux_posterior, lib), Int32, (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx().h, d, n, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ))))
    end
    return For(q0.f, φ)
end

# aux_posterior!(qΩ, lik, y, qf)                       -> aug_aux_posterior           (a5)
function dev_aux_posterior!(qΩ::For, lik::AbstractLikelihood, y::DV, qf::DeviceNormals)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_aux_posterior, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx().h, d, nobs(lik, qΩ), y.ptr, qf.μ.ptr, qf.σ².ptr, qf.ld, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ))))
    end
    return qΩ                                                    # bernoulli.jl:24
end

# expected_auglik_potential_and_precision(lik, qΩ, y[, qf]) -> aug_expected_potential_precision   (a7)
function dev_expected_potential_precision(lik::AbstractLikelihood, qΩ::For, y::DV, qf::Union{Nothing,DeviceNormals})
    n, nl = nobs(lik, qΩ), AGPL.nlatent(lik)
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

# expected_logtilt / aux_kldivergence / expected_aug_loglik -> aug_expected_elbo_terms (a9, a11, a13)
function dev_elbo_terms(lik::AbstractLikelihood, qΩ::For, y::DV, qf::DeviceNormals)
    sc = DV{Float64}(undef, 8)
    φ = state(qΩ)
    withdesc(lik) do d
        check(ccall((:aug_expected_elbo_terms, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                    ctx().h, d, nobs(lik, qΩ), y.ptr, qf.μ.ptr, qf.σ².ptr, qf.ld, ptr(s0(φ)), ptr(s1(φ)), ptr(s2(φ)), sc.ptr))
    end
    return Array(sc)          # multi-GPU callers run aug_allreduce_scalars on `sc` first (INTEGRATION.md)
end
# aux_kldivergence(lik, qΩ, y) has no qf in the reference (generic.jl:56-58); the KL kernels only read the state and
# y, so a zero qf of the right length is passed.  (The heteroscedastic likelihood has no prior / tilt split: its KL is
# part of expected_aug_loglik, heteroscedasticgaussian.jl:129-143, and needs the real qf.)
function dev_kldivergence(lik::AbstractLikelihood, qΩ::For, y::DV)
    z = AugDeviceVector(zeros(Float64, length(s0(state(qΩ)))))
    return dev_elbo_terms(lik, qΩ, y, DeviceNormals(z, z))[S_KL]
end

# The fused call of a CAVI iteration (examples/bernoulli/script.jl:29-39): aux_posterior!(qΩ, lik, y, qf) +
# expected_auglik_potential_and_precision + the expected_logtilt / aux_kldivergence sums in ONE pass -> aug_cavi_step
# (AugCUDA's own function: the reference has no fused verb, so there is nothing to be ambiguous with)
function cavi_step!(qΩ::For, lik::AbstractLikelihood, y::DV, qf::DeviceNormals)
    n, nl = nobs(lik, qΩ), AGPL.nlatent(lik)
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

# ---------------------------------------------------------------- sampling verbs (implementations)
needs_n(lik) = lik isa Union{PoissonLikelihood,HeteroscedasticGaussianLikelihood,CategoricalLikelihood}
"uninitialised device Ω with the reference's field names (ω, and n where the law has an integer part)"
function dev_alloc_aux_variables(lik::AbstractLikelihood, n::Integer)
    m = iscat(lik) ? n * AGPL.nlatent(lik) : n
    ω = DV{Float64}(undef, m)
    return needs_n(lik) ? TupleVector((; ω, n=DV{Int64}(undef, m))) : TupleVector((; ω))
end

# aux_sample!(rng, Ω, lik, y, f)                      -> aug_aux_sample                 (a14-a19)
function dev_aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::AbstractLikelihood, y::DV, f::DV; i0::Integer=0)
    seed!(rng)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = nobs(lik, Ω.ω)
    withdesc(lik) do d
        check(ccall((:aug_aux_sample, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, n, i0, y.ptr, f.ptr, ldof(lik, n), Ω.ω.ptr, ptr(nv)))
    end
    sync_offset!(rng)
    return Ω                                                     # generic.jl:11
end

# init_aux_variables(rng, lik, n)                     -> aug_init_aux_variables          (a4)
function dev_init_aux_variables(rng::AugPhilox, lik::AbstractLikelihood, n::Integer; i0::Integer=0)
    seed!(rng)
    Ω = dev_alloc_aux_variables(lik, n)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    withdesc(lik) do d
        check(ccall((:aug_init_aux_variables, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Float64}, Ptr{Int64}), ctx().h, d, n, i0, Ω.ω.ptr, ptr(nv)))
    end
    sync_offset!(rng)
    return Ω
end

# auglik_potential_and_precision(lik, Ω, y[, f])      -> aug_potential_precision         (a20)
function dev_potential_precision(lik::AbstractLikelihood, Ω::TupleVector, y::DV, f::Union{Nothing,DV})
    nl = AGPL.nlatent(lik)
    n = nobs(lik, Ω.ω)
    β, γ = DV{Float64}(undef, n * nl), DV{Float64}(undef, n * nl)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    withdesc(lik) do d
        check(ccall((:aug_potential_precision, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64},
                     Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, y.ptr, f === nothing ? C_NULL : f.ptr, f === nothing ? 0 : ldof(lik, n),
                    Ω.ω.ptr, ptr(nv), β.ptr, γ.ptr, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end

# logtilt / aug_loglik                                -> aug_sampled_loglik_terms         (a21-a24)
function dev_sampled_terms(lik::AbstractLikelihood, Ω::TupleVector, y::DV, f::DV, with_prior::Bool)
    sc = DV{Float64}(undef, 8)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = nobs(lik, Ω.ω)
    withdesc(lik) do d
        check(ccall((:aug_sampled_loglik_terms, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Int32,
                     Ptr{Float64}),
                    ctx().h, d, n, y.ptr, f.ptr, ldof(lik, n), Ω.ω.ptr, ptr(nv), with_prior, sc.ptr))
    end
    return Array(sc)
end

# ---------------------------------------------------------------- the two auxiliary-variable laws as handles (a10, a14)
# aux_prior(lik, y) and aux_full_conditional(lik, y, f) build one distribution object per observation in the reference
# (bernoulli.jl:13-15, 51-53, generic.jl:26-30).  The device side never materialises those: a handle carries the
# arguments, `logdensity_def(prior, Ω)` (generic.jl:49) and `tvrand(rng, cond)` (TestUtils.jl:108-110) run on the device.
struct DeviceAuxPrior{L<:AbstractLikelihood,Y}
    lik::L
    y::AugDeviceVector{Y}
end
struct DeviceAuxFullConditional{L<:AbstractLikelihood,Y}
    lik::L
    y::AugDeviceVector{Y}
    f::AugDeviceVector{Float64}
end
Base.length(p::DeviceAuxPrior) = nobs(p.lik, p.y)
Base.length(p::DeviceAuxFullConditional) = nobs(p.lik, p.y)
# the prior does not depend on f: slot S_LOGPRIOR of the sampled terms at f = 0
function AGPL.logdensity_def(p::DeviceAuxPrior, Ω::TupleVector)
    f0 = AugDeviceVector(zeros(Float64, length(Ω.ω) * (p.lik isa HeteroscedasticGaussianLikelihood ? 2 : 1)))
    return dev_sampled_terms(p.lik, Ω, p.y, f0, true)[S_LOGPRIOR]
end
AGPL.SpecialDistributions.tvrand(rng::AugPhilox, d::DeviceAuxFullConditional; i0::Integer=0) =
    dev_aux_sample!(rng, dev_alloc_aux_variables(d.lik, length(d)), d.lik, d.y, d.f; i0=i0)

# ---------------------------------------------------------------- host `Vector` implementations (the reference's own argument types)
# src/generic.jl:1-88 dispatches every verb on plain host vectors.  These bind to the *_host entry points: the library
# stages chunks through the GPU (H2D, kernel, D2H overlapped) and the results come back in host vectors.  The only
# host-side work is the AoS → SoA split of qf (mean.(qf), var.(qf)): the ABI takes struct-of-arrays.  Covered: the six
# likelihoods whose reference containers are plain `Vector`s.  The Categorical containers of the reference are
# ArraysOfArrays nested views (over a BitMatrix for y, categorical.jl:59-70): bind those through the device methods.
hptr(::Nothing) = C_NULL
hptr(v::Vector) = Ptr{Cvoid}(pointer(v))
moments(qf::AbstractVector{<:Normal}) = (convert(Vector{Float64}, mean.(qf)), convert(Vector{Float64}, var.(qf)))
# two-latent heteroscedastic qfg[i] = (qf_i, qg_i) (heteroscedasticgaussian.jl:34-46) -> latent-major [2][n]
moments(qfg::AbstractVector{<:AbstractVector{<:Normal}}) =
    (convert(Vector{Float64}, vcat(mean.(first.(qfg)), mean.(last.(qfg)))),
     convert(Vector{Float64}, vcat(var.(first.(qfg)), var.(last.(qfg)))))
# f for the sampling side: a plain vector, or fg[i] = (f_i, g_i) (heteroscedasticgaussian.jl:28-32) -> latent-major
latents(f::Vector{Float64}) = f
latents(fg::AbstractVector) = convert(Vector{Float64}, vcat(first.(fg), last.(fg)))
ldof(lik, n) = lik isa HeteroscedasticGaussianLikelihood ? n : 0

function host_aux_posterior!(qΩ::For, lik::AbstractLikelihood, y::Vector, qf)
    φ = state(qΩ); μ, σ² = moments(qf); n = length(qΩ)
    withdesc(lik) do d
        GC.@preserve y μ σ² φ check(ccall((:aug_aux_posterior_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx().h, d, n, hptr(y), μ, σ², ldof(lik, n), hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ))))
    end
    return qΩ
end

function host_expected_potential_precision(lik::AbstractLikelihood, qΩ::For, y::Vector, qf)
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

function host_elbo_terms(lik::AbstractLikelihood, qΩ::For, y::Vector, qf)
    sc = zeros(Float64, 8); φ = state(qΩ); μ, σ² = moments(qf); n = length(qΩ)
    withdesc(lik) do d
        GC.@preserve y μ σ² φ check(ccall((:aug_expected_elbo_terms_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                    ctx().h, d, n, hptr(y), μ, σ², ldof(lik, n), hptr(s0(φ)), hptr(s1(φ)), hptr(s2(φ)), sc))
    end
    return sc
end
host_kldivergence(lik::AbstractLikelihood, qΩ::For, y::Vector) =
    host_elbo_terms(lik, qΩ, y, fill(Normal(0.0, 0.0), length(qΩ)))[S_KL]

"fused CAVI iteration on host vectors; `want_state = false` / `want_β = false` skip those outputs (and their D2H bytes)"
function cavi_step!(qΩ::For, lik::AbstractLikelihood, y::Vector, qf; want_state::Bool=true, want_β::Bool=true)
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

function host_aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::AbstractLikelihood, y::Vector, fs::AbstractVector; i0::Integer=0)
    seed!(rng)
    f = latents(fs)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = length(Ω.ω)
    withdesc(lik) do d
        GC.@preserve y f Ω check(ccall((:aug_aux_sample_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, n, i0, hptr(y), f, ldof(lik, n), Ω.ω, nv === nothing ? C_NULL : pointer(nv)))
    end
    sync_offset!(rng)
    return Ω
end

"init_aux_variables on the device into HOST vectors (the counter-based stream instead of the reference's CPU rng)"
function init_aux_variables_host(rng::AugPhilox, lik::AbstractLikelihood, n::Int; i0::Integer=0)
    seed!(rng)
    m = iscat(lik) ? n * AGPL.nlatent(lik) : n
    ω = Vector{Float64}(undef, m)
    nv = needs_n(lik) ? Vector{Int64}(undef, m) : nothing
    withdesc(lik) do d
        GC.@preserve ω nv check(ccall((:aug_init_aux_variables_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Int64, Ptr{Float64}, Ptr{Int64}),
                    ctx().h, d, n, i0, ω, nv === nothing ? C_NULL : pointer(nv)))
    end
    sync_offset!(rng)
    return nv === nothing ? TupleVector((; ω)) : TupleVector((; ω, n=nv))
end

function host_potential_precision(lik::AbstractLikelihood, Ω::TupleVector, y::Vector, fs)
    nl = AGPL.nlatent(lik)
    n = length(Ω.ω)
    β, γ = Vector{Float64}(undef, n * nl), Vector{Float64}(undef, n * nl)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    f = fs === nothing ? nothing : latents(fs)
    withdesc(lik) do d
        GC.@preserve y f Ω β γ check(ccall((:aug_potential_precision_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64},
                     Ptr{Float64}, Ptr{Float64}, Int64),
                    ctx().h, d, n, hptr(y), f === nothing ? C_NULL : pointer(f), ldof(lik, n),
                    Ω.ω, nv === nothing ? C_NULL : pointer(nv), β, γ, n))
    end
    return split_latents(β, n, nl), split_latents(γ, n, nl)
end

function host_sampled_terms(lik::AbstractLikelihood, Ω::TupleVector, y::Vector, fs::AbstractVector, with_prior::Bool)
    sc = zeros(Float64, 8)
    f = latents(fs)
    nv = hasproperty(Ω, :n) ? Ω.n : nothing
    n = length(Ω.ω)
    withdesc(lik) do d
        GC.@preserve y f Ω check(ccall((:aug_sampled_loglik_terms_host, lib), Int32,
                    (Ptr{Cvoid}, Ref{AugLik}, Int64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}, Int32,
                     Ptr{Float64}),
                    ctx().h, d, n, hptr(y), f, ldof(lik, n), Ω.ω, nv === nothing ? C_NULL : pointer(nv),
                    with_prior, sc))
    end
    return sc
end

# ---------------------------------------------------------------- the dispatch layer: methods of the reference's verbs
# (likelihood type as the reference dispatches on it, element type of y as the C ABI takes it, container of qf / f as
#  the reference's own methods take it)
const NormalVec = AbstractVector{<:Normal}
const NormalVecVec = AbstractVector{<:AbstractVector{<:Normal}}
const REFERENCE_LIKELIHOODS = (
    (BernoulliLikelihood{<:LogisticLink}, Bool, NormalVec),               # bernoulli.jl:3-70    (Julia Bool = 1 byte)
    (AGPL.NegBinomialLikelihood, Int64, NormalVec),                       # negativebinomial.jl:1
    (AGPL.AugPoisson, Int64, NormalVec),                                  # poisson.jl:7
    (LaplaceLikelihood, Float64, NormalVec),                              # laplace.jl:13-17
    (StudentTLikelihood, Float64, NormalVec),                             # studentt.jl:14-21
    (AGPL.AugHeteroGaussian, Float64, NormalVecVec),                      # heteroscedasticgaussian.jl:7, :34-39
    (AGPL.BijectiveLogisticSoftMaxLikelihood, Bool, nothing),             # categorical.jl:38-41 (device containers only)
    (AGPL.LogisticSoftMaxLikelihood, Bool, nothing),
)

for (L, Y, QF) in REFERENCE_LIKELIHOODS
    # ---- device containers: every argument is a subtype of the reference's, the likelihood type is the reference's
    @eval begin
        AGPL.init_aux_posterior(::Type{AugDeviceVector}, lik::$L, n::Int) = dev_init_aux_posterior(lik, n)
        AGPL.aux_posterior!(qΩ::For, lik::$L, y::DV{$Y}, qf::DeviceNormals) = dev_aux_posterior!(qΩ, lik, y, qf)
        AGPL.aux_posterior(lik::$L, y::DV{$Y}, qf::DeviceNormals) =
            dev_aux_posterior!(dev_init_aux_posterior(lik, nobs(lik, y)), lik, y, qf)          # generic.jl:22-24
        AGPL.expected_auglik_potential_and_precision(lik::$L, qΩ::For, y::DV{$Y}) =
            dev_expected_potential_precision(lik, qΩ, y, nothing)
        AGPL.expected_auglik_potential_and_precision(lik::$L, qΩ::For, y::DV{$Y}, qf::Union{Nothing,DeviceNormals}) =
            dev_expected_potential_precision(lik, qΩ, y, qf)
        AGPL.expected_auglik_potential(lik::$L, qΩ::For, y::DV{$Y}) = first(dev_expected_potential_precision(lik, qΩ, y, nothing))
        AGPL.expected_auglik_potential(lik::$L, qΩ::For, y::DV{$Y}, qf::Union{Nothing,DeviceNormals}) =
            first(dev_expected_potential_precision(lik, qΩ, y, qf))
        AGPL.expected_auglik_precision(lik::$L, qΩ::For, y::DV{$Y}) = last(dev_expected_potential_precision(lik, qΩ, y, nothing))
        AGPL.expected_auglik_precision(lik::$L, qΩ::For, y::DV{$Y}, qf::Union{Nothing,DeviceNormals}) =
            last(dev_expected_potential_precision(lik, qΩ, y, qf))
        AGPL.expected_logtilt(lik::$L, qΩ::For, y::DV{$Y}, qf::DeviceNormals) = dev_elbo_terms(lik, qΩ, y, qf)[S_ELT]
        AGPL.expected_aug_loglik(lik::$L, qΩ::For, y::DV{$Y}, qf::DeviceNormals) = dev_elbo_terms(lik, qΩ, y, qf)[S_EAUGLL]
        AGPL.init_aux_variables(rng::AugPhilox, lik::$L, n::Int; i0::Integer=0) = dev_init_aux_variables(rng, lik, n; i0=i0)
        AGPL.aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::$L, y::DV{$Y}, f::DV{Float64}; i0::Integer=0) =
            dev_aux_sample!(rng, Ω, lik, y, f; i0=i0)
        AGPL.aux_sample!(Ω::TupleVector, lik::$L, y::DV{$Y}, f::DV{Float64}) = dev_aux_sample!(GLOBAL_PHILOX, Ω, lik, y, f)
        AGPL.aux_sample(rng::AugPhilox, lik::$L, y::DV{$Y}, f::DV{Float64}; i0::Integer=0) =
            dev_aux_sample!(rng, dev_alloc_aux_variables(lik, nobs(lik, y)), lik, y, f; i0=i0)   # generic.jl:18-20
        AGPL.aux_sample(lik::$L, y::DV{$Y}, f::DV{Float64}) = AGPL.aux_sample(GLOBAL_PHILOX, lik, y, f)
        AGPL.auglik_potential_and_precision(lik::$L, Ω::TupleVector, y::DV{$Y}) = dev_potential_precision(lik, Ω, y, nothing)
        AGPL.auglik_potential_and_precision(lik::$L, Ω::TupleVector, y::DV{$Y}, f::Union{Nothing,DV{Float64}}) =
            dev_potential_precision(lik, Ω, y, f)
        AGPL.auglik_potential(lik::$L, Ω::TupleVector, y::DV{$Y}) = first(dev_potential_precision(lik, Ω, y, nothing))
        AGPL.auglik_potential(lik::$L, Ω::TupleVector, y::DV{$Y}, f::Union{Nothing,DV{Float64}}) =
            first(dev_potential_precision(lik, Ω, y, f))
        AGPL.auglik_precision(lik::$L, Ω::TupleVector, y::DV{$Y}) = last(dev_potential_precision(lik, Ω, y, nothing))
        AGPL.auglik_precision(lik::$L, Ω::TupleVector, y::DV{$Y}, f::Union{Nothing,DV{Float64}}) =
            last(dev_potential_precision(lik, Ω, y, f))
        AGPL.aux_full_conditional(lik::$L, y::DV{$Y}, f::DV{Float64}) = DeviceAuxFullConditional(lik, y, f)
        AGPL.logtilt(lik::$L, Ω::TupleVector, y::DV{$Y}, f::DV{Float64}) = dev_sampled_terms(lik, Ω, y, f, false)[S_LOGTILT]
        AGPL.aug_loglik(lik::$L, Ω::TupleVector, y::DV{$Y}, f::DV{Float64}) = dev_sampled_terms(lik, Ω, y, f, true)[S_AUGLL]
    end
    if L !== AGPL.AugHeteroGaussian        # no prior / tilt split (see dev_kldivergence)
        @eval AGPL.aux_kldivergence(lik::$L, qΩ::For, y::DV{$Y}) = dev_kldivergence(lik, qΩ, y)
        @eval AGPL.aux_prior(lik::$L, y::DV{$Y}) = DeviceAuxPrior(lik, y)
    end
    QF === nothing && continue
    # ---- host Vectors.  y::Vector{Y} and qf::QF are subtypes of (or equal to) what the reference's methods take, qΩ::For /
    # Ω::TupleVector are narrower than its untyped arguments: with AugCUDA loaded these calls run on the GPU.  The
    # sampling verbs go to the GPU when the rng is an AugPhilox (the reference's CPU rngs keep the CPU path).
    @eval begin
        AGPL.aux_posterior!(qΩ::For, lik::$L, y::Vector{$Y}, qf::$QF) = host_aux_posterior!(qΩ, lik, y, qf)
        AGPL.expected_auglik_potential_and_precision(lik::$L, qΩ::For, y::Vector{$Y}) =
            host_expected_potential_precision(lik, qΩ, y, nothing)
        AGPL.expected_auglik_potential_and_precision(lik::$L, qΩ::For, y::Vector{$Y}, ::Nothing) =
            host_expected_potential_precision(lik, qΩ, y, nothing)
        AGPL.expected_auglik_potential_and_precision(lik::$L, qΩ::For, y::Vector{$Y}, qf::$QF) =
            host_expected_potential_precision(lik, qΩ, y, qf)
        AGPL.expected_auglik_potential(lik::$L, qΩ::For, y::Vector{$Y}) = first(host_expected_potential_precision(lik, qΩ, y, nothing))
        AGPL.expected_auglik_potential(lik::$L, qΩ::For, y::Vector{$Y}, qf::$QF) =
            first(host_expected_potential_precision(lik, qΩ, y, qf))
        AGPL.expected_auglik_precision(lik::$L, qΩ::For, y::Vector{$Y}) = last(host_expected_potential_precision(lik, qΩ, y, nothing))
        AGPL.expected_auglik_precision(lik::$L, qΩ::For, y::Vector{$Y}, qf::$QF) =
            last(host_expected_potential_precision(lik, qΩ, y, qf))
        AGPL.expected_logtilt(lik::$L, qΩ::For, y::Vector{$Y}, qf::$QF) = host_elbo_terms(lik, qΩ, y, qf)[S_ELT]
        AGPL.expected_aug_loglik(lik::$L, qΩ::For, y::Vector{$Y}, qf::$QF) = host_elbo_terms(lik, qΩ, y, qf)[S_EAUGLL]
        AGPL.aux_sample!(rng::AugPhilox, Ω::TupleVector, lik::$L, y::Vector{$Y}, f::AbstractVector; i0::Integer=0) =
            host_aux_sample!(rng, Ω, lik, y, f; i0=i0)
        AGPL.auglik_potential_and_precision(lik::$L, Ω::TupleVector, y::Vector{$Y}) = host_potential_precision(lik, Ω, y, nothing)
        AGPL.auglik_potential_and_precision(lik::$L, Ω::TupleVector, y::Vector{$Y}, ::Nothing) =
            host_potential_precision(lik, Ω, y, nothing)
        AGPL.logtilt(lik::$L, Ω::TupleVector, y::Vector{$Y}, f::Vector{Float64}) = host_sampled_terms(lik, Ω, y, f, false)[S_LOGTILT]
        AGPL.aug_loglik(lik::$L, Ω::TupleVector, y::Vector{$Y}, f::Vector{Float64}) = host_sampled_terms(lik, Ω, y, f, true)[S_AUGLL]
    end
    if L !== AGPL.AugHeteroGaussian
        @eval AGPL.aux_kldivergence(lik::$L, qΩ::For, y::Vector{$Y}) = host_kldivergence(lik, qΩ, y)
    else
        # the sampled-side containers of the reference for two latents: fg[i] = (f_i, g_i) (heteroscedasticgaussian.jl:48-66)
        @eval begin
            AGPL.auglik_potential_and_precision(lik::$L, Ω::TupleVector, y::Vector{$Y}, fg::AbstractVector{<:AbstractVector{<:Real}}) =
                host_potential_precision(lik, Ω, y, fg)
            AGPL.auglik_potential(lik::$L, Ω::TupleVector, y::Vector{$Y}, fg::AbstractVector{<:AbstractVector{<:Real}}) =
                first(host_potential_precision(lik, Ω, y, fg))
            AGPL.auglik_precision(lik::$L, Ω::TupleVector, y::Vector{$Y}, fg::AbstractVector{<:AbstractVector{<:Real}}) =
                last(host_potential_precision(lik, Ω, y, fg))
        end
    end
end

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
    K = lik.invlink isa BijectiveSimplexLink ? AGPL.nlatent(lik) + 1 : AGPL.nlatent(lik)
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
