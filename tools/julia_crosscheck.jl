# julia_crosscheck.jl — pins the oracle (and through it the GPU path) to the REAL reference.
#
# Runs the UNMODIFIED SelfConcordantSmoothOptimization.jl on the eight problems of tests/cases.py through the package's
# documented closure path (README.md:153-187: grad_fx / hess_fx / out_fn / jac_yx / grad_fy / hess_fy supplied by the
# user, exactly the derivatives ForwardDiff would produce) and writes Solution.x / obj / fval / epochs as hex floats:
#
#     python tools/dump_case_inputs.py /tmp/scs_cases
#     julia --project=<checkout of the reference package> tools/julia_crosscheck.jl /tmp/scs_cases tests/golden/julia
#     python -m pytest tests/test_julia_crosscheck.py        # diffs tests/golden/julia/*.json against tests/golden/*.json
#
# No Julia binary exists in the build image (and no network for the package's dependencies), so this script has not been
# executed there; tests/test_julia_crosscheck.py skips until its output is present.
#
#     julia ... tools/julia_crosscheck.jl --bench <ggn|n|lqn> <logistic|ls> <l1|gl|indbox> <rows> <cols> <steps>
# times iterate! on random data of the benchmark's shape (bench.py --impl reference uses it when julia is on PATH) and
# prints "SECONDS_PER_ITER <s>".
using SelfConcordantSmoothOptimization
using LinearAlgebra, Printf, Random

const SCS = SelfConcordantSmoothOptimization

hexf(v) = v === nothing ? "null" : "\"" * @sprintf("%a", Float64(v)) * "\""
hexlist(v) = "[" * join((hexf(e) for e in v), ", ") * "]"

readvec(path) = collect(reinterpret(Float64, read(path)))

# ---- the README / test closures (README.md:113,135-139,212-214,233-239; test/test_algs.jl:9-11) ----------------------
sigm(z) = 1 ./ (1 .+ exp.(-z))
function logistic_closures(c::Float64, consistent::Bool)
    yc(y) = consistent ? (y .+ 1) ./ 2 : y
    f(A, y, x) = c * sum(log.(1 .+ exp.(-y .* (A * x))))
    f(y, yhat) = -c * sum(yc(y) .* log.(yhat) .+ (1 .- yc(y)) .* log.(1 .- yhat))
    out_fn(A, x) = sigm(A * x)
    function grad_fx(A, y, x)
        e = exp.(-y .* (A * x))
        return A' * (c .* (-y) .* (e ./ (1 .+ e)))
    end
    function hess_fx(A, y, x)
        e = exp.(-y .* (A * x))
        return A' * (Diagonal(c .* (y .* y) .* (e ./ ((1 .+ e) .* (1 .+ e)))) * A)
    end
    function jac_yx(A, y, yhat, x)
        e2 = exp.(-(A * x))
        return Matrix(Diagonal((yhat ./ (1 .+ e2)) .* e2) * A)
    end
    grad_fy(A, y, yhat) = Vector(-c .* (yc(y) ./ yhat .- (1 .- yc(y)) ./ (1 .- yhat)))
    hess_fy(A, y, yhat) = Diagonal(Vector(c .* (yc(y) ./ (yhat .* yhat) .+ (1 .- yc(y)) ./ ((1 .- yhat) .* (1 .- yhat)))))
    return (f=f, out_fn=out_fn, grad_fx=grad_fx, hess_fx=hess_fx, jac_yx=jac_yx, grad_fy=grad_fy, hess_fy=hess_fy)
end
function ls_closures(den::Float64)
    f(A, y, x) = 0.5 * sum((A * x .- y) .^ 2) / den
    f(y, yhat) = 0.5 * sum((yhat .- y) .^ 2) / den
    out_fn(A, x) = A * x
    grad_fx(A, y, x) = A' * ((A * x .- y) ./ den)
    hess_fx(A, y, x) = (A' * A) ./ den
    jac_yx(A, y, yhat, x) = Matrix(A)
    grad_fy(A, y, yhat) = Vector((yhat .- y) ./ den)
    hess_fy(A, y, yhat) = Diagonal(fill(1 / den, length(y)))
    return (f=f, out_fn=out_fn, grad_fx=grad_fx, hess_fx=hess_fx, jac_yx=jac_yx, grad_fy=grad_fy, hess_fy=hess_fy)
end

# the same table as tests/cases.py::build
function build(name, A, y, x0)
    n, m = size(A)
    ggn(c) = (out_fn=c.out_fn, jac_yx=c.jac_yx, grad_fy=c.grad_fy, hess_fy=c.hess_fy)
    if name == "c1_readme_logreg_n"
        c = logistic_closures(1 / m, false)
        return ProxNSCORE(), Problem(A, y, x0, c.f, 1e-1; grad_fx=c.grad_fx, hess_fx=c.hess_fx), "l1",
               PHuberSmootherL1L2(1.0), (max_epoch=100, x_tol=1e-6, f_tol=1e-6)
    elseif name == "c2_logreg_ggn_l1"
        c = logistic_closures(1 / n, true)
        return ProxGGNSCORE(), Problem(A, y, x0, c.f, 1e-2; ggn(c)...), "l1", PHuberSmootherL1L2(1.0), (max_epoch=12, α=1)
    elseif name == "c2b_logreg_ggn_literal"
        c = logistic_closures(1 / n, false)
        return ProxGGNSCORE(), Problem(A, y, x0, c.f, 1e-3; ggn(c)...), "l1", PHuberSmootherL1L2(1.0), (max_epoch=6, α=1)
    elseif name == "c3_logreg_lqn_l1"
        c = logistic_closures(1 / n, false)
        return ProxLQNSCORE(m=10), Problem(A, y, x0, c.f, 1e-2; grad_fx=c.grad_fx), "l1", PHuberSmootherL1L2(1.0),
               (max_epoch=30, α=1)
    elseif name == "c3b_logreg_lqn_l2_bb"
        c = logistic_closures(1 / n, false)
        return ProxLQNSCORE(ss_type=2, m=5), Problem(A, y, x0, c.f, 1e-3; grad_fx=c.grad_fx), "l2",
               PHuberSmootherL1L2(0.5), (max_epoch=12,)
    elseif name == "c4_ls_ggn_gl"
        gsz = 64; ng = m ÷ gsz
        ind = vcat([(g * gsz + 1) for g in 0:ng-1]', [((g + 1) * gsz) for g in 0:ng-1]', ones(Int, ng)')
        P = SCS.get_P(m, collect(1:m), Matrix{Int}(ind))
        c = ls_closures(Float64(n))
        model = Problem(A, y, x0, c.f, [1e-8, 1e-2]; P=P, ggn(c)...)
        return ProxGGNSCORE(), model, "gl", PHuberSmootherGL(1e-2, model), (max_epoch=10, α=1)
    elseif name == "c5_ls_n_indbox"
        c = ls_closures(Float64(n))
        return ProxNSCORE(), Problem(A, y, x0, c.f, 1e-4; C_set=(-0.5, 0.5), grad_fx=c.grad_fx, hess_fx=c.hess_fx),
               "indbox", PHuberSmootherIndBox(-0.5, 0.5, 0.6), (max_epoch=15, α=0.8)
    elseif name == "c5b_ls_lqn_logexp"
        c = ls_closures(Float64(n))
        return ProxLQNSCORE(m=10), Problem(A, y, x0, c.f, 1.0; C_set=(-0.5, Inf), grad_fx=c.grad_fx), "indbox",
               LogExpSmootherIndBox(-0.5, Inf, 10.0), (max_epoch=15, α=1)
    end
    error("unknown case $name")
end

function crosscheck(indir, outdir)
    mkpath(outdir)
    for line in eachline(joinpath(indir, "cases.tsv"))
        name, ns, ms = split(line, '\t')
        n, m = parse(Int, ns), parse(Int, ms)
        A = reshape(readvec(joinpath(indir, name * "_A.bin")), n, m)
        y = readvec(joinpath(indir, name * "_y.bin"))
        x0 = readvec(joinpath(indir, name * "_x0.bin"))
        method, model, reg, hμ, kw = build(String(name), A, y, x0)
        sol = iterate!(method, model, reg, hμ; verbose=0, kw...)
        open(joinpath(outdir, String(name) * ".json"), "w") do io
            println(io, "{\"case\": \"$name\", \"source\": \"SelfConcordantSmoothOptimization.jl (unmodified), closure path, julia $(VERSION)\",")
            println(io, " \"epochs\": $(sol.epochs),")
            println(io, " \"x\": ", hexlist(sol.x), ",")
            println(io, " \"obj\": ", hexlist(sol.obj), ",")
            println(io, " \"fval\": ", hexlist(sol.fval), ",")
            println(io, " \"pri_res_norm\": ", hexlist(sol.pri_res_norm), ",")
            println(io, " \"support\": [", join(findall(!iszero, sol.x) .- 1, ", "), "]}")
        end
        println(name, ": epochs ", sol.epochs, ", history ", length(sol.obj))
    end
end

function bench(method_s, loss_s, reg_s, n, m, steps)
    Random.seed!(1234)
    A = randn(n, m) ./ sqrt(m)
    xt = [rand() < 0.05 ? 3 * randn() : 0.0 for _ in 1:m]
    x0 = randn(m)
    if loss_s == "logistic"
        y = [rand() < 1 / (1 + exp(-z)) ? 1.0 : -1.0 for z in A * xt]
        c = logistic_closures(1 / n, true)
    else
        y = A * xt .+ 0.1 .* randn(n)
        c = ls_closures(Float64(n))
    end
    kw = Dict{Symbol,Any}()
    λ, α, hμ_of = 1e-3, 1.0, (model -> PHuberSmootherL1L2(1.0))
    if reg_s == "gl"
        ng = m ÷ 64
        ind = vcat([(g * 64 + 1) for g in 0:ng-1]', [((g + 1) * 64) for g in 0:ng-1]', ones(Int, ng)')
        kw[:P] = SCS.get_P(m, collect(1:m), Matrix{Int}(ind)); λ = [1e-8, 1e-2]; hμ_of = (model -> PHuberSmootherGL(1e-2, model))
    elseif reg_s == "indbox"
        kw[:C_set] = (-0.5, 0.5); λ = 1e-4; α = 0.8; hμ_of = (model -> PHuberSmootherIndBox(-0.5, 0.5, 0.6))
    end
    if method_s == "ggn"
        method = ProxGGNSCORE(); merge!(kw, Dict(:out_fn => c.out_fn, :jac_yx => c.jac_yx, :grad_fy => c.grad_fy, :hess_fy => c.hess_fy))
    elseif method_s == "n"
        method = ProxNSCORE(); merge!(kw, Dict(:grad_fx => c.grad_fx, :hess_fx => c.hess_fx))
    else
        method = ProxLQNSCORE(m=10); kw[:grad_fx] = c.grad_fx
    end
    model = Problem(A, y, x0, c.f, λ; kw...)
    hμ = hμ_of(model)
    iterate!(method, model, reg_s, hμ; verbose=0, max_epoch=1, α=α, x_tol=0.0, f_tol=0.0)   # compile
    model = Problem(A, y, x0, c.f, λ; kw...)
    t = @elapsed iterate!(method, model, reg_s, hμ; verbose=0, max_epoch=steps, α=α, x_tol=0.0, f_tol=0.0)
    println("SECONDS_PER_ITER ", t / steps)
end

if length(ARGS) >= 1 && ARGS[1] == "--bench"
    bench(ARGS[2], ARGS[3], ARGS[4], parse(Int, ARGS[5]), parse(Int, ARGS[6]), parse(Int, ARGS[7]))
else
    crosscheck(length(ARGS) >= 1 ? ARGS[1] : "/tmp/scs_cases", length(ARGS) >= 2 ? ARGS[2] : "tests/golden/julia")
end
