# make_golden_julia.jl — run the REFERENCE (AugmentedGPLikelihoods.jl itself) on the inputs of golden_cavi.json and
# write what it returns to golden_julia.json.
#
# Why: the build image has no Julia, so the oracle (oracle/aug_oracle.cpp) is pinned on the reference's own
# known-answer tests and on an independent mpmath restatement (make_golden.py), but NOT on Julia output for the values
# the reference's test-suite never checks (expected_logtilt, aux_kldivergence, expected_aug_loglik, everything
# heteroscedastic / categorical: src/TestUtils.jl:193-204 only asserts `isa Real`).  This script closes that gap on any
# machine that has Julia:
#
#     julia --project=/path/to/AugmentedGPLikelihoods.jl -e 'using Pkg; Pkg.add("JSON")'
#     julia --project=/path/to/AugmentedGPLikelihoods.jl tests/golden/make_golden_julia.jl
#
# and tests/test_julia_golden.py then holds the oracle (CPU) and libaugcuda (GPU) to 1e-12 of the file it wrote.
# Every call is wrapped: where the reference throws (e.g. the heteroscedastic expected_aug_loglik calls
# `var(first(qg))` on a Normal, heteroscedasticgaussian.jl:140) the error text is recorded instead of a value.
using AugmentedGPLikelihoods
using AugmentedGPLikelihoods: nlatent, aux_posterior, expected_auglik_potential_and_precision, expected_logtilt,
    aux_kldivergence, expected_aug_loglik, aux_sample, auglik_potential_and_precision, logtilt, aug_loglik, aux_prior
using GPLikelihoods, Distributions, ArraysOfArrays, TupleVectors, Random, JSON

const HERE = @__DIR__
cases = JSON.parsefile(joinpath(HERE, "golden_cavi.json"))

function make_lik(c)
    k, p = c["kind"], c["params"]
    k == 0 && return BernoulliLikelihood()
    k == 1 && return NegativeBinomialLikelihood(NBParamFailure(get(c, "r_is_int", 0) == 1 ? Int(p[1]) : Float64(p[1])), LogisticLink())
    k == 2 && return PoissonLikelihood(ScaledLogistic(Float64(p[1])))
    k == 3 && return LaplaceLikelihood(Float64(p[1]))
    k == 4 && return StudentTLikelihood(Float64(p[1]), Float64(p[2]))
    k == 5 && return HeteroscedasticGaussianLikelihood(InvScaledLogistic(Float64(p[1])))
    lt = Float64.(c["logtheta"])
    k == 6 && return CategoricalLikelihood(BijectiveSimplexLink(LogisticSoftMaxLink(lt)))
    return CategoricalLikelihood(LogisticSoftMaxLink(lt))
end

tovec(x) = x isa AbstractVector{<:AbstractVector} ? [Float64.(collect(v)) for v in x] : Float64.(collect(x))
attempt(f) = try
    (f(), nothing)
catch e
    (nothing, sprint(showerror, e))
end

out = Dict{String,Any}()
for (name, c) in cases
    lik = make_lik(c)
    k, nl = c["kind"], get(c, "nlatent", c["kind"] == 5 ? 2 : 1)
    r = Dict{String,Any}()
    # ---- inputs in the shapes the reference's methods take
    if k in (6, 7)                                   # [n][nl] row lists -> vectors over observations
        n = length(c["y"])
        y = nestedview(hcat([Bool.(row .!= 0) for row in c["y"]]...))
        qf = [Normal.(Float64.(c["mu"][i]), sqrt.(Float64.(c["var"][i]))) for i in 1:n]
        f = [Float64.(c["mu"][i]) for i in 1:n]
    elseif k == 5                                    # per-observation pairs qfg[i] = [qf_i, qg_i], fg[i] = [f_i, g_i]: the verbs
        y = Float64.(c["y"])                         # broadcast first / last over them (heteroscedasticgaussian.jl:34-45, 48-66;
        n = length(y)                                # examples/heteroscedasticgaussian/script.jl:59-60 passes invert(posts_fs))
        qf = [[Normal(Float64(c["mu"][j][i]), sqrt(Float64(c["var"][j][i]))) for j in 1:2] for i in 1:n]
        f = [[Float64(c["mu"][j][i]) for j in 1:2] for i in 1:n]
    else
        y = k == 0 ? Bool.(c["y"] .!= 0) : (k in (1, 2) ? Int.(c["y"]) : Float64.(c["y"]))
        qf = Normal.(Float64.(c["mu"]), sqrt.(Float64.(c["var"])))
        f = Float64.(c["mu"])
    end
    # ---- variational side
    qΩ, err = attempt(() -> aux_posterior(lik, y, qf))
    r["err_aux_posterior"] = err
    if qΩ !== nothing
        φ = only(qΩ.inds)
        for (slot, field) in (("s0", (:c, :μ, :β)), ("s1", (:λ, :p)), ("s2", (:ψ,)))
            for fld in field
                hasproperty(φ, fld) && (r[slot] = tovec(getproperty(φ, fld)))
            end
        end
        bg, err = attempt(() -> expected_auglik_potential_and_precision(lik, qΩ, y, qf))
        r["err_expected_potential_precision"] = err
        bg !== nothing && (r["beta"] = [tovec(b) for b in bg[1]]; r["gamma"] = [tovec(g) for g in bg[2]])
        for (key, fn) in (("elt", () -> expected_logtilt(lik, qΩ, y, qf)),
                          ("kl", () -> aux_kldivergence(lik, qΩ, y)),
                          ("eall", () -> expected_aug_loglik(lik, qΩ, y, qf)))
            v, err = attempt(fn)
            r[key] = v
            r["err_" * key] = err
        end
    end
    # ---- sampling side: Ω drawn by the reference, then the deterministic verbs on it
    Ω, err = attempt(() -> aux_sample(MersenneTwister(1), lik, y, f))
    r["err_aux_sample"] = err
    if Ω !== nothing
        r["omega"] = tovec(Ω.ω)
        hasproperty(Ω, :n) && (r["n"] = tovec(Ω.n))
        bg, err = attempt(() -> auglik_potential_and_precision(lik, Ω, y, f))
        r["err_potential_precision"] = err
        bg !== nothing && (r["s_beta"] = [tovec(b) for b in bg[1]]; r["s_gamma"] = [tovec(g) for g in bg[2]])
        for (key, fn) in (("logtilt", () -> logtilt(lik, Ω, y, f)), ("aug_loglik", () -> aug_loglik(lik, Ω, y, f)))
            v, err = attempt(fn)
            r[key] = v
            r["err_" * key] = err
        end
    end
    out[name] = r
end
open(joinpath(HERE, "golden_julia.json"), "w") do io
    JSON.print(io, out, 1)
end
println("wrote ", joinpath(HERE, "golden_julia.json"), " (", length(out), " cases)")
