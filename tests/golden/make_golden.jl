# Parity pin kit: run with the reference itself to pin this repository's oracle and CUDA path to AutoGP.jl.
#
#     julia --project=/path/to/AutoGP.jl tests/golden/make_golden.jl [output directory]
#
# (AutoGP.jl @ 2ad372d, v0.1.19; any environment in which `using AutoGP, Gen` works.)  Writes
# `julia_golden.json` next to this file (or into the given directory).  tests/test_julia_golden.py picks the file up
# when it exists and compares BOTH the CPU oracle and the CUDA path against it; without the file those tests are
# skipped with a message that says so.  Nothing in this repository can produce these values: Julia is not installed
# in the build image, and the LML arithmetic lives in Gen.jl / Distributions.jl / PDMats.jl / OpenBLAS.
#
# Contents (same inputs as tests/golden/make_golden.py, which holds the oracle's own values):
#   gram    : GP.eval_cov(kernel, ts) for the 6 base kernels of test/test_GP.jl:24-33 and the 108 composites of
#             :54-56, on every 7th point of the test's grid (15 points)                  [114][15][15], row = i
#   gram_noise / gram_scalar : compute_cov_matrix_vectorized / compute_cov_matrix with noise 0.1 for the same kernels
#   lml     : Gen.logpdf(Gen.mvnormal, xs, zeros(n), compute_cov_matrix_vectorized(kernel, 0.1, ts)) for the same
#   hmc_lml : the three benchmark (kernel, noise) pairs of test/experiment_hmc.jl:180-184 on 40 points of its grid,
#             noise + Model.JITTER as src/Model.jl:134 adds it
#   hmc_grad: Gen.logpdf_grad(Gen.mvnormal, ...) contracted with dK/dnoise = I, i.e. dLML/dnoise, for the same
#   predictive: Distributions.MvNormal(kernel, noise, ts, xs, ts_pred) mean and covariance (src/GP.jl:731-758) for the
#             Plus benchmark, and the identity logpdf(dist, xs_test) = LML(all) - LML(obs) of experiment_hmc.jl:111-132
using AutoGP
using AutoGP: GP, Model, Transforms
import Gen
import LinearAlgebra
using Printf

outdir = length(ARGS) >= 1 ? ARGS[1] : @__DIR__

base_kernels() = [
    GP.WhiteNoise(1), GP.Constant(0.5), GP.Linear(0.1, 1.3, 0.7), GP.SquaredExponential(0.47, 0.13),
    GP.GammaExponential(0.42, 0.58, 3.2), GP.Periodic(0.96, 0.21, 1.1)]

function fixture_kernels()
    base = base_kernels()
    out = Any[b for b in base]
    for b1 in base, b2 in base
        push!(out, GP.Plus(b1, b2))
        push!(out, GP.Times(b1, b2))
        push!(out, GP.ChangePoint(b1, b2, 0.5, 0.95))
    end
    return out
end

ds_raw = collect(range(start=-10, stop=10, length=100))
transformation = Transforms.LinearTransform(ds_raw, 0, 1)
ds = Transforms.apply(transformation, ds_raw)
sub = ds[1:7:end]
fixture_xs(ts) = sin.(3.0 .* ts) .+ 0.2 .* cos.(11.0 .* ts)

lml(kernel, noise, ts, xs) = Gen.logpdf(Gen.mvnormal, xs, zeros(length(ts)), GP.compute_cov_matrix_vectorized(kernel, noise, ts))

# JSON by hand (no JSON.jl dependency): shortest round-trip digits of every Float64
num(x::Real) = isfinite(x) ? repr(Float64(x)) : (isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity"))
vec_json(v) = "[" * join((num(x) for x in v), ",") * "]"
mat_json(M) = "[" * join((vec_json(M[i, :]) for i in 1:size(M, 1)), ",") * "]"

kernels = fixture_kernels()
xs = fixture_xs(sub)
io = IOBuffer()
print(io, "{\n")
print(io, "\"autogp_version\": \"", string(pkgversion(AutoGP)), "\",\n")
print(io, "\"julia_version\": \"", string(VERSION), "\",\n")
print(io, "\"ts\": ", vec_json(sub), ",\n")
print(io, "\"xs\": ", vec_json(xs), ",\n")
print(io, "\"noise\": 0.1,\n")
print(io, "\"gram\": [", join((mat_json(GP.eval_cov(k, sub)) for k in kernels), ","), "],\n")
print(io, "\"gram_noise\": [", join((mat_json(GP.compute_cov_matrix_vectorized(k, 0.1, sub)) for k in kernels), ","), "],\n")
print(io, "\"gram_scalar\": [", join((mat_json(GP.compute_cov_matrix(k, 0.1, sub)) for k in kernels), ","), "],\n")
print(io, "\"lml\": ", vec_json([lml(k, 0.1, sub, xs) for k in kernels]), ",\n")

# experiment_hmc.jl:180-184 benchmarks on 40 points of its grid (ts in [0, 10], as the experiment uses them)
benchmarks = [
    (GP.SquaredExponential(2), 0.01),
    (GP.Plus(GP.Linear(.5), GP.Periodic(2, 1)), 0.05),
    (GP.ChangePoint(GP.Linear(.5), GP.Linear(1.5), 1, .001), 0.001)]
ts_h = collect(range(0, 10, length=1000))[1:25:end]
xs_h = fixture_xs(ts_h ./ 10.0)
print(io, "\"hmc_ts\": ", vec_json(ts_h), ",\n")
print(io, "\"hmc_xs\": ", vec_json(xs_h), ",\n")
print(io, "\"hmc_lml\": ", vec_json([lml(k, nz + Model.JITTER, ts_h, xs_h) for (k, nz) in benchmarks]), ",\n")
hmc_grad = Float64[]
for (k, nz) in benchmarks
    K = GP.compute_cov_matrix_vectorized(k, nz + Model.JITTER, ts_h)
    (_, _, dK) = Gen.logpdf_grad(Gen.mvnormal, xs_h, zeros(length(ts_h)), K)
    push!(hmc_grad, LinearAlgebra.tr(dK))   # dLML/dnoise = sum_ii dLML/dK_ii
end
print(io, "\"hmc_grad_noise\": ", vec_json(hmc_grad), ",\n")

# predictive distribution and the Bayes identity on the Plus benchmark: 30 observed, 10 held-out points
(k2, nz2) = benchmarks[2]
noise2 = nz2 + Model.JITTER
ts_obs, xs_obs, ts_test, xs_test = ts_h[1:30], xs_h[1:30], ts_h[31:40], xs_h[31:40]
dist = GP.Distributions.MvNormal(k2, noise2, ts_obs, xs_obs, ts_test)
print(io, "\"pred_mean\": ", vec_json(GP.Distributions.mean(dist)), ",\n")
print(io, "\"pred_cov\": ", mat_json(Matrix(GP.Distributions.cov(dist))), ",\n")
print(io, "\"pred_logpdf\": ", num(GP.Distributions.logpdf(dist, xs_test)), ",\n")
print(io, "\"pred_logpdf_bayes\": ", num(lml(k2, noise2, ts_h, xs_h) - lml(k2, noise2, ts_obs, xs_obs)), "\n")
print(io, "}\n")
path = joinpath(outdir, "julia_golden.json")
open(path, "w") do f
    write(f, String(take!(io)))
end
println("wrote ", path, ": ", length(kernels), " kernels")
