# SCSB200.jl — Julia shim that routes the per-iteration hot path of SelfConcordantSmoothOptimization.jl
# (ProxNSCORE / ProxGGNSCORE / ProxLQNSCORE) to libscs_b200.so through `ccall`.
#
# Drop-in: `Problem(A, y, x0, f, λ; ...)`, `iterate!(method, problem, reg_name, hμ; ...)` — the SAME function, this module
# adds a more specific method for `model::Problem` — and the `Solution` fields keep their reference meaning (histories,
# `fvaltest` from `Atest/ytest`, `metricvals`, `times`, verbose printing through the reference's own `show_stat!`).
# The only change a user makes is to pass one of the built-in loss objects
# (`LogisticLoss`, `LeastSquaresLoss`, `QuadFormLoss`; all `<: Function`, so `Problem`'s `f::Function` signature,
# src/problems.jl:65, accepts them).  An arbitrary closure `f` — which only ForwardDiff could differentiate — is
# rejected with an error instead of silently running on the CPU.
#
# NOTE: no Julia binary exists in the build image, so this file has not been executed there.  It is kept 1:1 with
# the Python ctypes mirror (scs_b200/api.py), which is what the test-suite drives; both bind the same C symbols
# (include/scs_b200.h).
module SCSB200

using SelfConcordantSmoothOptimization
import SelfConcordantSmoothOptimization: iterate!, ProximalMethod, ProxNSCORE, ProxGGNSCORE, ProxLQNSCORE,
    Problem, Solution, PHuberSmootherL1L2, PHuberSmootherIndBox, PHuberSmootherGL, ExponentialSmootherIndBox,
    LogExpSmootherIndBox, OsBaSmootherL1L2, OsBaSmootherGL, bounds_sanity_check
using LinearAlgebra, Dates, Random, SparseArrays

export LogisticLoss, LeastSquaresLoss, QuadFormLoss, GPUContext, gpu_iterate!, iterate!

const LIB = get(ENV, "SCS_B200_LIB", joinpath(@__DIR__, "..", "libscs_b200.so"))

# ---- built-in losses: callable like the README closures, so the CPU reference can run the same object ----------
struct LogisticLoss <: Function      # README.md:113,135-139 / test/test_algs.jl:9-11
    scale::Float64
    consistent_labels::Bool          # false: cross-entropy sees y as given (README feeds ±1); true: (y+1)/2
end
LogisticLoss(scale::Real) = LogisticLoss(Float64(scale), false)
(L::LogisticLoss)(A, y, x) = L.scale * sum(log.(1 .+ exp.(-y .* (A * x))))
function (L::LogisticLoss)(y, ŷ)
    yc = L.consistent_labels ? (y .+ 1) ./ 2 : y
    return -L.scale * sum(yc .* log.(ŷ) .+ (1 .- yc) .* log.(1 .- ŷ))
end
struct LeastSquaresLoss <: Function  # README.md:212-214,233-235
    denom::Float64
end
(L::LeastSquaresLoss)(A, y, x) = 0.5 * sum((A * x .- y) .^ 2) / L.denom
(L::LeastSquaresLoss)(y, ŷ) = 0.5 * sum((ŷ .- y) .^ 2) / L.denom
struct QuadFormLoss <: Function      # test/test_algs.jl:90
end
(L::QuadFormLoss)(A, y, x) = 1 / 2 * (x' * (A * x)) + (y' * x)

loss_code(L::LogisticLoss) = (Cint(0), L.scale, Cint(L.consistent_labels ? 1 : 0))
loss_code(L::LeastSquaresLoss) = (Cint(1), L.denom, Cint(0))
loss_code(L::QuadFormLoss) = (Cint(2), 0.0, Cint(0))
loss_code(f) = Base.error("scs_b200: arbitrary user f (ForwardDiff-only path) is not supported on the GPU; " *
                          "pass LogisticLoss / LeastSquaresLoss / QuadFormLoss")

# ---- status handling: rethrow with Base.error so the user-visible failure class matches the reference -----------
lasterr() = unsafe_string(ccall((:scs_last_error, LIB), Cstring, ()))
check(code::Cint) = code == 0 ? nothing : Base.error("scs_b200 [$code]: " * lasterr())

mutable struct GPUContext
    h::Ptr{Cvoid}
    function GPUContext(device::Integer=0; rank::Integer=0, world::Integer=1, unique_id::Vector{UInt8}=UInt8[])
        out = Ref{Ptr{Cvoid}}(C_NULL)
        idp = world > 1 ? pointer(unique_id) : Ptr{UInt8}(C_NULL)
        GC.@preserve unique_id check(ccall((:scs_ctx_create, LIB), Cint, (Cint, Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
                                           device, rank, world, idp, out))
        c = new(out[])
        finalizer(x -> ccall((:scs_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:scs_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end

mutable struct GPUProblem
    h::Ptr{Cvoid}
    m::Int
    ctx::GPUContext
end

# Upload model.A / model.y once (replaces the per-iteration `Matrix(As')` copy of iterate.jl:206-207).
function GPUProblem(ctx::GPUContext, model)
    code, param, lmode = loss_code(model.f)
    any(x -> x !== nothing, (model.grad_fx, model.hess_fx, model.jac_yx, model.grad_fy, model.hess_fy)) &&
        Base.error("scs_b200: user derivative closures cannot run on the GPU")
    (model.out_fn !== nothing) && @info "out_fn is ignored: the built-in loss carries its own model output function"
    y = Vector{Float64}(vec(model.y))
    n, m = size(model.A)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    if model.A isa SparseMatrixCSC   # README.md:105 builds A with sprandn: only the stored entries cross PCIe
        S = SparseMatrixCSC{Float64,Int64}(model.A)
        colptr, rowval, nzval = S.colptr, S.rowval, S.nzval
        GC.@preserve colptr rowval nzval y check(ccall((:scs_problem_create_csc, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Cint, Float64, Cint, Cint,
             Ref{Ptr{Cvoid}}), ctx.h, colptr, rowval, nzval, 1, n, m, y, code, param, lmode, 0, out))  # storage 0 = auto
    else
        A = Matrix{Float64}(model.A)
        GC.@preserve A y check(ccall((:scs_problem_create, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Cint, Float64, Cint, Ref{Ptr{Cvoid}}),
            ctx.h, A, n, m, stride(A, 2), y, code, param, lmode, out))
    end
    p = GPUProblem(out[], m, ctx)
    finalizer(x -> ccall((:scs_problem_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    return p
end

const REG = Dict("l1" => 0, "l2" => 1, "indbox" => 2, "gl" => 3)
smoother_code(::PHuberSmootherL1L2) = 0
smoother_code(::PHuberSmootherIndBox) = 1
smoother_code(::PHuberSmootherGL) = 2
smoother_code(::ExponentialSmootherIndBox) = 3
smoother_code(::LogExpSmootherIndBox) = 4
smoother_code(::OsBaSmootherL1L2) = 5
smoother_code(::OsBaSmootherGL) = 6
method_code(::ProxNSCORE) = 0
method_code(::ProxGGNSCORE) = 1
method_code(::ProxLQNSCORE) = 2

function bounds_of(C_set)   # prox-operators.jl:36-45
    if SelfConcordantSmoothOptimization.is_interval_set(C_set)
        return isa(C_set, Tuple) ? ([minimum.(C_set)...], [maximum.(C_set)...]) : ([minimum(C_set)], [maximum(C_set)])
    end
    return (collect(Float64, C_set[1]), collect(Float64, C_set[2]))
end

# The IndBox smoothers close over their bounds; the shim needs them again for the device descriptor, so the caller
# passes them through `smoother_bounds` (defaults to model.C_set, which is what README / tests use).
function configure!(p::GPUProblem, method, model, reg_name::String, hμ; smoother_bounds=nothing)
    haskey(REG, reg_name) || Base.error("reg_name not valid.")
    λ = model.λ
    lam1, lam2 = reg_name == "gl" ? (Float64(λ[1]), Float64(λ[2])) : (Float64(length(λ) > 1 ? λ[1] : λ), 0.0)
    reg_name == "gl" && length(λ) != 2 &&
        Base.error("Please provide a Tuple or Vector with exactly two entries for λ, e.g. [λ1, λ2]")
    ind = reg_name == "gl" ? Matrix{Int64}(model.P.ind) : zeros(Int64, 3, 0)
    perm = reg_name == "gl" ? Vector{Int64}(model.P.G) : Int64[]
    lb, ub = reg_name == "indbox" ? bounds_of(model.C_set) : (Float64[], Float64[])
    GC.@preserve ind perm lb ub check(ccall((:scs_set_regularizer, LIB), Cint,
        (Ptr{Cvoid}, Cint, Float64, Float64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
        p.h, REG[reg_name], lam1, lam2, isempty(ind) ? C_NULL : pointer(ind), size(ind, 2),
        isempty(perm) ? C_NULL : pointer(perm), isempty(lb) ? C_NULL : pointer(lb), length(lb),
        isempty(ub) ? C_NULL : pointer(ub), length(ub)))
    slb, sub = smoother_bounds === nothing ? (lb, ub) : (collect(Float64, smoother_bounds[1]), collect(Float64, smoother_bounds[2]))
    GC.@preserve slb sub check(ccall((:scs_set_smoother, LIB), Cint,
        (Ptr{Cvoid}, Cint, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
        p.h, smoother_code(hμ), Float64(hμ.μ), isempty(slb) ? C_NULL : pointer(slb), length(slb),
        isempty(sub) ? C_NULL : pointer(sub), length(sub)))
    mem = method isa ProxLQNSCORE ? method.m : 10
    check(ccall((:scs_set_method, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint),
                p.h, method_code(method), method.ss_type, method.use_prox ? 1 : 0, mem))
    check(ccall((:scs_set_L, LIB), Cint, (Ptr{Cvoid}, Cint, Float64), p.h, model.L === nothing ? 0 : 1,
                model.L === nothing ? 0.0 : Float64(model.L)))
end

# model.f(model.A, model.y, x) and get_reg(model, x, reg_name)   — iterate.jl:189-190
function gpu_objective(p::GPUProblem, x::Vector{Float64})
    f = Ref{Float64}(0.0); r = Ref{Float64}(0.0)
    GC.@preserve x check(ccall((:scs_objective, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ref{Float64}), p.h, x, f, r))
    return f[], r[]
end

# step!(method, model, reg_name, hμ, As, x, x_prev, ys, Cmat, iter)   — iterate.jl:233
function gpu_step(p::GPUProblem, x::Vector{Float64}, x_prev::Vector{Float64}, iter::Integer; return_dx=false)
    x_new = Vector{Float64}(undef, p.m)
    dx = return_dx ? Vector{Float64}(undef, p.m) : Float64[]
    pri = Ref{Float64}(0.0)
    GC.@preserve x x_prev x_new dx check(ccall((:scs_step, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
        p.h, x, x_prev, iter, x_new, return_dx ? pointer(dx) : C_NULL, pri))
    return return_dx ? (x_new, dx, pri[]) : (x_new, pri[])
end

# rows [lo, hi) (0-based, half-open) of the resident shard take part in the following passes: one mini-batch
set_active_rows!(p::GPUProblem, lo::Integer, hi::Integer) =
    check(ccall((:scs_set_active_rows, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), p.h, lo, hi))

# The data loader of optim_loop! (iterate.jl:122-145, utils.jl:14-25): row order after the one-time shuffle and the
# batch offsets.  MLUtils.DataLoader(batchsize, shuffle, partial=true) gives ceil(n/b) consecutive batches.  max_iter and
# iend are fixed BEFORE slice_samples sets batch_size = 1 (:124-127 vs :136-138): without a batch_size max_iter stays 1, so
# the loader subset 1:iend (:145) is ONE entry — for slice_samples that is the first row only.
function batch_plan(n::Integer; batch_size=nothing, slice_samples=false, shuffle_batch=true, local_max_iter=nothing)
    batch_size !== nothing && slice_samples && (slice_samples = false)
    max_iter = batch_size !== nothing ? cld(n, batch_size) : 1
    iend = (local_max_iter !== nothing && Int(floor(local_max_iter)) > 0) ? min(Int(floor(local_max_iter)), max_iter) : max_iter
    slice_samples && ((batch_size, shuffle_batch) = (1, false))
    batch_size === nothing && ((batch_size, shuffle_batch) = (n, false))
    order = shuffle_batch ? Random.shuffle(1:n) : nothing
    offsets = [min(i * batch_size, n) for i in 0:iend]
    return order, offsets
end

const BuiltinLoss = Union{LogisticLoss,LeastSquaresLoss,QuadFormLoss}

# ---- the drop-in entry point -------------------------------------------------------------------------------------
# `iterate!(method, problem, reg_name, hμ; ...)` ITSELF (iterate.jl:56-76): this method is more specific than the
# reference's (`model::Problem` vs `model::SCMOModel`), so after `using .SCSB200` every `iterate!` on a `Problem` lands here.
# A built-in loss goes to the GPU.  Anything else (an arbitrary closure f, which only ForwardDiff can differentiate) is
# REJECTED — never run on the CPU silently; `backend=:cpu` asks for the reference's own CPU path explicitly.
function iterate!(method::ProximalMethod, model::Problem, reg_name, hμ; backend::Symbol=:gpu, ctx=nothing,
                  smoother_bounds=nothing, kwargs...)
    if backend == :cpu
        return invoke(iterate!, Tuple{ProximalMethod,SelfConcordantSmoothOptimization.SCMOModel,Any,Any}, method, model,
                      reg_name, hμ; kwargs...)
    end
    model.f isa BuiltinLoss ||
        Base.error("scs_b200: arbitrary user f (ForwardDiff-only path) is not supported on the GPU; pass LogisticLoss / " *
                   "LeastSquaresLoss / QuadFormLoss, or call iterate!(...; backend=:cpu) for the reference CPU path")
    return gpu_iterate!(method, model, reg_name, hμ; ctx=(ctx === nothing ? GPUContext(0) : ctx),
                        smoother_bounds=smoother_bounds, kwargs...)
end

# optim_loop! (iterate.jl:100-266) with the two hot call sites routed to the GPU.  Everything that is host bookkeeping
# upstream stays the reference's own code: Options, show_stat! (verbose printing, ftest, metrics: utils.jl:50-104) and
# update_stat! (histories: utils.jl:106-113) are called as they are.  Mini-batches: the rows are uploaded once in the
# loader's (shuffled) order, so every batch is a contiguous row range of the resident matrix; the objective is always
# taken over all rows (:189).
function gpu_iterate!(method::ProximalMethod, model, reg_name, hμ; ctx::GPUContext=GPUContext(0),
                      metrics=nothing, α=nothing, batch_size=nothing, slice_samples=false, shuffle_batch=true,
                      max_epoch=1000, comm_rounds=100, local_max_iter=nothing, x_tol=1e-10, f_tol=1e-10, verbose=1,
                      smoother_bounds=nothing)
    SCS = SelfConcordantSmoothOptimization
    opt = SCS.Options(metrics=metrics, α=α, batch_size=batch_size, slice_samples=slice_samples,
                      shuffle_batch=shuffle_batch, max_epoch=(local_max_iter !== nothing ? 1 : max_epoch),
                      comm_rounds=comm_rounds, local_max_iter=local_max_iter, x_tol=x_tol, f_tol=f_tol, verbose=verbose)
    max_epoch = opt.max_epoch                                    # iterate.jl:58-70
    implemented_algs = []
    SCS.set_name!(method, implemented_algs)                      # :112
    α !== nothing && (model.L = 1 / α)                           # :113-115
    if method.name in implemented_algs && method.ss_type == 1 && model.L === nothing && verbose > 0
        @info "Neither L nor α is set for the problem... Now fixing α = 0.5..."   # :116-120
    end
    n = size(model.A, 1)
    if batch_size !== nothing && slice_samples
        @info "Cannot use both batch_size and slice_samples=true...\nNow setting slice_samples=false..."  # :128-131
    end
    batched = batch_size !== nothing || slice_samples
    order, offsets = batch_plan(n; batch_size=batch_size, slice_samples=slice_samples, shuffle_batch=shuffle_batch,
                                local_max_iter=local_max_iter)
    opt.max_iter = batch_size !== nothing ? cld(n, batch_size) : 1
    dev_model = model
    if order !== nothing   # one-time shuffle: upload the rows in loader order (the caller's model is left untouched)
        dev_model = deepcopy(model); dev_model.A = model.A[order, :]; dev_model.y = model.y[order]
    end
    windows = batched ? [(offsets[i], offsets[i+1]) for i in 1:length(offsets)-1] : [(0, n)]
    iend = length(windows)
    p = GPUProblem(ctx, dev_model)
    configure!(p, method, model, reg_name, hμ; smoother_bounds=smoother_bounds)
    objective(v) = (batched && set_active_rows!(p, 0, n); gpu_objective(p, v))
    # held-out data (iterate.jl:169-176): a second resident shard; ftest(x) = model.f(Atest, ytest, x)
    test_model = all(x -> x !== nothing, (model.Atest, model.ytest))
    if xor(model.Atest === nothing, model.ytest === nothing)
        @info "Both input (Atest) and target (ytest) data are required for testing the model, but only one of these has been provided.\nWill skip testing..."
    end
    ptest = nothing
    if test_model
        tm = deepcopy(model); tm.A = model.Atest; tm.y = model.ytest
        ptest = GPUProblem(ctx, tm)
        configure!(ptest, method, model, reg_name, hμ; smoother_bounds=smoother_bounds)
    end
    ftest = test_model ? (v -> gpu_objective(ptest, v)[1]) : (v -> nothing)
    fvals, pri_res_norms, fvaltests, objs, rel_errors, f_rel_errors, times = [], [], [], [], [], [], []
    metric_vals = Dict()
    if metrics !== nothing
        for name in keys(metrics); metric_vals[name] = []; end
    end
    epochs = 0
    x_star = model.x
    fs, rs = objective(x_star)
    obj_star = fs + rs                                           # :179
    x = model.x0; x_prev = deepcopy(x)
    pri_res_norm = nothing
    l_split = "="^30 * "\n"
    check(ccall((:scs_method_init, LIB), Cint, (Ptr{Cvoid},), p.h))   # init!(method, x) :183
    rel_err(v) = reg_name == "gl" ? SCS.mean_square_error(x_star, v) : max(norm(v - x_star) / max.(norm(x_star), 1), x_tol)
    frel(o) = max((norm(o - obj_star)) / norm(obj_star), f_tol)
    t0 = now()
    for epoch_t in 1:max_epoch
        Δtime = (now() - t0).value / 1000
        fval, reg = objective(x); obj = fval + reg               # :189-190
        rel_error = rel_err(x)
        f_rel_error = frel(obj)
        SCS.show_stat!(opt, model, method, x, test_model, ftest, fvaltests, epoch_t - 1, obj, fval, pri_res_norm, rel_error, Δtime, metric_vals, l_split)
        SCS.update_stat!(objs, obj, fvals, fval, pri_res_norms, pri_res_norm, rel_errors, rel_error, f_rel_errors, f_rel_error, times, Δtime)
        for (i, (lo, hi)) in enumerate(windows)                  # :204
            if verbose > 2
                (i in [1, iend]) || (i % 100 == 0) ? print("\n[$i/$iend]") : print("#")
            end
            if epoch_t == max_epoch && i == iend                 # :219-231
                Δtime = (now() - t0).value / 1000
                fval, reg = objective(x); obj = fval + reg
                rel_error = rel_err(x)
                SCS.show_stat!(opt, model, method, x, test_model, ftest, fvaltests, epoch_t, obj, fval, pri_res_norm, rel_error, Δtime, metric_vals, l_split; is_max_epoch=true)
                f_rel_error = frel(obj)
                SCS.update_stat!(objs, obj, fvals, fval, pri_res_norms, pri_res_norm, rel_errors, rel_error, f_rel_errors, f_rel_error, times, Δtime)
            end
            batched && set_active_rows!(p, lo, hi)
            x_new, pri_res_norm = gpu_step(p, x, x_prev, epoch_t)   # :233
            if norm(x_new - x) < x_tol * max.(norm(x), 1) || f_rel_error ≤ f_tol || pri_res_norm < x_tol
                if epoch_t != max_epoch                          # :235-247
                    Δtime = (now() - t0).value / 1000
                    fval, reg = objective(x_new); obj = fval + reg
                    rel_error = rel_err(x_new)
                    SCS.show_stat!(opt, model, method, x_new, test_model, ftest, fvaltests, epoch_t, obj, fval, pri_res_norm, rel_error, Δtime, metric_vals, l_split; is_terminate_epoch=true)
                    f_rel_error = frel(obj)
                    SCS.update_stat!(objs, obj, fvals, fval, pri_res_norms, pri_res_norm, rel_errors, rel_error, f_rel_errors, f_rel_error, times, Δtime)
                end
                x_prev = deepcopy(x); x = x_new; epochs += 1
                break
            end
            x_prev = deepcopy(x); x = x_new
        end
        if norm(x - x_prev) < x_tol * max.(norm(x_prev), 1) || f_rel_error ≤ f_tol || pri_res_norm < x_tol   # :257
            break
        end
        epochs += 1
        verbose > 2 && print("\n" * l_split)
    end
    return Solution(x, objs, fvals, pri_res_norms, fvaltests, rel_errors, f_rel_errors, metric_vals, times, epochs, model)
end

end # module
