/*
 * scs_b200.h — C ABI of the B200-native hot path for the proximal SCORE solvers
 * (ProxNSCORE / ProxGGNSCORE / ProxLQNSCORE) of SelfConcordantSmoothOptimization.jl v0.1.8.
 *
 * This is the drop-in boundary: the Julia shim (julia/SCSB200.jl) `ccall`s exactly these symbols;
 * the Python ctypes harness (scs_b200/_capi.py) binds the same ones.  Plain pointers and sizes only.
 * All floating point on the wire is fp64, all sizes/indices int64.  Every entry point returns an
 * scs_status; the message of the last failure on the calling thread is scs_last_error().
 * The library never falls back to the CPU: anything it cannot run on the GPU is SCS_UNSUPPORTED.
 *
 * Reference interface each entry replaces (paths relative to the reference repository):
 *   scs_problem_create      Problem(A, y, x0, f, λ; …)                     src/problems.jl:61-81
 *                           + the per-iteration copy As = Matrix(As')       src/algorithms/iterate.jl:206-207 (eliminated)
 *   scs_set_regularizer     reg_name + model.λ / model.C_set / model.P      src/regularizers/regularizers.jl:4-31,
 *                                                                           src/prox/prox-operators.jl:68-80,
 *                                                                           src/utils/prox-reg-utils.jl:27-62
 *   scs_set_smoother        hμ = PHuberSmootherL1L2(μ) | …IndBox | …GL | …  src/regularizers/phuber-smooth.jl:27,59-64,137-148,
 *                                                                           exponential-smooth.jl:28-34, log-exp-smooth.jl:28-34,
 *                                                                           ostrovskii-bach-smooth.jl:27,59-70
 *   scs_set_method          ProxNSCORE/ProxGGNSCORE/ProxLQNSCORE fields     src/algorithms/prox-N-SCORE.jl:6-22,
 *                                                                           prox-GGN-SCORE.jl:6-22, prox-L-BFGS-SCORE.jl:6-30
 *   scs_set_L               model.L = 1/α                                   src/algorithms/iterate.jl:113-115
 *   scs_method_init         init!(method, x)                                iterate.jl:183, prox-L-BFGS-SCORE.jl:31-36
 *   scs_objective           model.f(model.A, model.y, x), get_reg(model,x,reg_name)   iterate.jl:168,189-190
 *   scs_step                step!(method, model, reg_name, hμ, As, x, x_prev, ys, Cmat, iter; return_dx)
 *                                                                           iterate.jl:52-54,233; prox-*-SCORE.jl step!
 *   scs_solve               optim_loop!(method, model, reg_name, hμ; opt)   iterate.jl:100-266
 *   scs_set_test_problem    model.Atest / model.ytest, ftest(x), Solution.fvaltest          iterate.jl:169-176, utils.jl:55-57
 *   scs_set_batches /       get_data_loader / get_loader_subset / the inner `for (i, sample) in enumerate(data)`
 *   scs_set_active_rows                                                     iterate.jl:139-145,204-207, utils.jl:14-25
 *   scs_loss_eval           f / gradient(f,x) / out_fn pieces               prox-N-SCORE.jl:49-69, prox-GGN-SCORE.jl:44-56
 *   scs_gram                hessian(f,x) | Jt*Q*Jt'                         prox-N-SCORE.jl:63, prox-GGN-SCORE.jl:129
 *   scs_linear_solve        (H + λHr) \ ∇q | qr(JQJ) \ Je                   prox-N-SCORE.jl:70, prox-GGN-SCORE.jl:131
 *   scs_smoother_eval       hμ.grad(Cmat,x), hμ.hess(Cmat,x)                prox-*-SCORE.jl (N:40-42, GGN:41-43, LQN:76-78)
 *   scs_prox                prox_step(invoke_prox(model,reg,x,h,λ,α))       prox-operators.jl:8-80
 *   scs_reg_value           get_reg(model, x, reg_name)                     regularizers.jl:4-31
 */
#ifndef SCS_B200_H
#define SCS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCS_B200_VERSION 100 /* 0.1.0 */

typedef struct scs_ctx scs_ctx;         /* one per (process, GPU); owns stream, workspaces, communicator */
typedef struct scs_problem scs_problem; /* device-resident row shard of (A, y) + solver state */

typedef enum {
  SCS_OK = 0,
  SCS_INVALID_ARG = 1,  /* Base.error-class failures of the reference: bad reg_name, λ not length 2 for gl, bad ss_type, μ<=0, bound-length mismatch */
  SCS_UNSUPPORTED = 2,  /* arbitrary user f / ForwardDiff-only path, ProblemGeneric, mini-batch, ss_type=2 for N/GGN after iter 1 */
  SCS_NOT_SPD = 3,      /* reserved: factorisation failed and the pivoted fallback was disabled */
  SCS_CUDA_ERROR = 4,
  SCS_NCCL_ERROR = 5,
  SCS_OOM = 6,
  SCS_STATE_ERROR = 7   /* call order violated (e.g. scs_step before scs_set_method) */
} scs_status;

typedef enum { SCS_LOSS_LOGISTIC = 0, SCS_LOSS_LEASTSQUARES = 1, SCS_LOSS_QUADFORM = 2 } scs_loss_kind;
/* LOGISTIC:      f(A,y,x) = p*sum(log(1+exp(-y.*(A*x))));  f(y,ŷ) = -p*sum(y log ŷ + (1-y) log(1-ŷ)); out_fn = σ(Ax)  (p = loss_param = scale)
 * LEASTSQUARES:  f(A,y,x) = 0.5*sum((A*x-y).^2)/p;        f(y,ŷ) = 0.5*sum((ŷ-y).^2)/p;              out_fn = Ax     (p = denominator)
 * QUADFORM:      f(A,y,x) = 1/2*(x'*(A*x)) + y'*x  (A square, single GPU, N / LQN only)                               */
typedef enum { SCS_LABELS_LITERAL = 0, SCS_LABELS_CONSISTENT = 1 } scs_label_mode;
typedef enum { SCS_REG_L1 = 0, SCS_REG_L2 = 1, SCS_REG_INDBOX = 2, SCS_REG_GL = 3 } scs_reg_kind;
typedef enum {
  SCS_SMOOTH_PHUBER_L1L2 = 0, SCS_SMOOTH_PHUBER_INDBOX = 1, SCS_SMOOTH_PHUBER_GL = 2,
  SCS_SMOOTH_EXP_INDBOX = 3, SCS_SMOOTH_LOGEXP_INDBOX = 4, SCS_SMOOTH_OSBA_L1L2 = 5, SCS_SMOOTH_OSBA_GL = 6
} scs_smoother_kind;
typedef enum { SCS_METHOD_N = 0, SCS_METHOD_GGN = 1, SCS_METHOD_LQN = 2 } scs_method_kind;
typedef enum { SCS_WEIGHTS_NEWTON = 0, SCS_WEIGHTS_GGN = 1 } scs_weight_kind;

int scs_version(void);
const char* scs_last_error(void);

/* ---- context / communicator ------------------------------------------------------------ */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host (torch.distributed, MPI.jl …). */
int scs_comm_unique_id(void* id128);
/* world == 1: id128 may be NULL and no communicator is made.  One context per process per GPU. */
int scs_ctx_create(int device, int rank, int world, const void* id128, scs_ctx** out);
int scs_ctx_destroy(scs_ctx* ctx);
int scs_ctx_sync(scs_ctx* ctx);
/* the CUDA stream every kernel of this context is launched on (a cudaStream_t, as an integer) */
int scs_ctx_stream(scs_ctx* ctx, uint64_t* stream_out);

/* ---- problem --------------------------------------------------------------------------- */
/* A: this rank's rows, column-major, leading dimension lda >= n_local (Julia Matrix{Float64}); y: n_local
 * doubles.  Host pointers; copied once to the device, never retained.  n_total = sum of n_local over ranks. */
int scs_problem_create(scs_ctx* ctx, const double* A_colmajor, int64_t n_local, int64_t m, int64_t lda,
                       const double* y, int loss_kind, double loss_param, int label_mode, scs_problem** out);
/* The same from a compressed-sparse-column shard (Julia SparseMatrixCSC{Float64,Int64}: colptr (m+1), rowval, nzval;
 * index_base 1 for Julia, 0 for scipy).  Only the stored entries cross PCIe.  storage: 1 = expand on the device into the
 * dense resident layout (dense kernels: HBM-roofline passes, tensor-core Gram); 2 = keep the shard sparse (CSR + CSC
 * copies; k_sp_forward / k_sp_adjoint / k_sp_gram, ~12 bytes per stored entry and pass; no int8 Gram, no
 * scs_problem_read_rows); 0 = sparse when fewer than 4 % of the entries are stored.
 * scs_problem_is_sparse reports which one was taken.  scs_get_gram_path returns 3 for the sparse Gram. */
int scs_problem_create_csc(scs_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                           int64_t index_base, int64_t n_local, int64_t m, const double* y, int loss_kind,
                           double loss_param, int label_mode, int storage, scs_problem** out);
int scs_problem_is_sparse(scs_problem* p, int* sparse, int64_t* nnz);
/* Drops the explicit zeros of a dense resident shard and keeps it in the sparse layout instead (CSR + CSC copies built on
 * the device, the dense matrix is freed).  For benchmark-sized synthetic shards (scs_problem_create_synthetic with
 * density < 1, the README's sprandn(n, m, 0.01) look-alike): call right after creation, before the first pass. */
int scs_problem_sparsify(scs_problem* p);
/* Synthetic shard generated on the device (bench / full-size invariants): rows [row0,row0+n_local) of the
 * n_total x m problem of oracle/synth.py (Philox-4x32-10, seed).  task: 0 = logistic labels, 1 = LS targets. */
int scs_problem_create_synthetic(scs_ctx* ctx, int64_t n_total, int64_t row0, int64_t n_local, int64_t m,
                                 int loss_kind, double loss_param, int label_mode, uint64_t seed, double density,
                                 scs_problem** out);
int scs_problem_destroy(scs_problem* p);
/* copy rows [row0,row0+nrows) of the resident shard (and y) back to the host, column-major ld = nrows */
int scs_problem_read_rows(scs_problem* p, int64_t row0, int64_t nrows, double* A_out, double* y_out);

/* reg: lam1 = λ (l1/l2/indbox) or λ1 (gl); lam2 = λ2 (gl).  gl: ind3xG = 3 x ngroups int64, column-major
 * (1-based start, end, integer weight) and perm = P.G (1-based, length m) or NULL for identity.
 * indbox: lb/ub of length nlb/nub in {1, m} = model.C_set. */
int scs_set_regularizer(scs_problem* p, int reg_kind, double lam1, double lam2, const int64_t* ind3xG,
                        int64_t ngroups, const int64_t* perm, const double* lb, int64_t nlb, const double* ub,
                        int64_t nub);
/* smoother hμ; lb/ub are the smoother's own bounds (IndBox kinds), ±Inf allowed (mapped to ±1e32). */
int scs_set_smoother(scs_problem* p, int smoother_kind, double mu, const double* lb, int64_t nlb,
                     const double* ub, int64_t nub);
int scs_set_method(scs_problem* p, int method_kind, int ss_type, int use_prox, int lbfgs_m);
int scs_set_L(scs_problem* p, int has_L, double L);
/* Gram kernel selection: 0 = auto (tcgen05 int8 emulated-fp64 path when the shard is large — m >= 512, >= 32768 active
 * rows — and its residue planes fit in HBM, DMMA otherwise), 1 = always DMMA.8x8x4, 2 = tcgen05 int8 whenever possible.
 * The int8 path serves weights of either sign (scs_get_gram_signed); only non-finite weights fall back to DMMA, which
 * propagates them.  scs_get_gram_path reports what the last Gram used (1 = DMMA, 2 = tcgen05 int8, 3 = sparse). */
int scs_set_gram_mode(scs_problem* p, int mode);
int scs_get_gram_path(scs_problem* p, int* path);
/* Fixed-point precision of the emulated-fp64 Gram.  The columns of diag(sqrt w) A are scaled to a common 2-norm T before
 * rounding to integers (then every entry of the integer Gram is below T^2, which is what the CRT range P/2 has to hold);
 * an entry of the Gram is off by ~0.4/T * sqrt(w_max / mean w) of the diagonal scale (standard deviation; the largest
 * entry error is a few times that), whatever the number of rows.  `bits` = log2 of the T asked
 * for (24..58, default 46: ~3e-15 rms, the rounding noise of an fp64 DGEMM over ~1000 rows; call before the first Gram).  The
 * library takes the shortest prefix of its 15 moduli that holds T (12 for the default), then uses all of that prefix's
 * range; single entries are capped at 2^50 (exact fp64 integers).  scs_get_gram_info reports the moduli count and
 * floor(log2 T) of the T actually used (0, 0 before the int8 path has run). */
int scs_set_gram_bits(scs_problem* p, int bits);
int scs_get_gram_info(scs_problem* p, int* nmod, int* bits);
/* Weights of both signs (the README's +-1-label cross-entropy pair, README.md:135-139 / test/test_algs.jl:9-11, makes
 * Q_ii negative for part of the rows): the fixed-point image is built from sqrt(|w|) and the rows of the MINORITY sign
 * are additionally compacted into a second set of planes, G = X'X - 2 Xc'Xc (or -X'X + 2 Xc'Xc when most rows are
 * negative) — one int8 SYRK over all rows plus one over the compacted rows, combined exactly inside the CRT.
 * Reports how the last emulated Gram ran: *compact_rows = rows in the compacted planes (-1: the weights had one sign),
 * *minority_negative = 1 if those were the negative rows. */
int scs_get_gram_signed(scs_problem* p, int64_t* compact_rows, int* minority_negative);
/* Mini-batches (optim_loop!'s data loader, iterate.jl:139-145,204-207; utils.jl:18-25).  The host lays the rows of a
 * shard out batch after batch (after its one-time shuffle, if any); a batch is then a contiguous local row range.
 * scs_set_active_rows restricts every following pass (scs_step, scs_loss_eval, scs_gram, and scs_objective — the
 * host resets it to [0, n_local) for the objective, which the reference always takes over the whole data) to rows
 * [row_lo, row_hi) of the shard; any bounds are accepted, the kernels sweep the 128-row-aligned superset and mask.
 * scs_set_batches gives scs_solve the table of nbatch+1 local offsets it steps through every epoch (nbatch = 0:
 * full batch).  With several ranks every rank passes its own offsets for the same nbatch batches. */
int scs_set_active_rows(scs_problem* p, int64_t row_lo, int64_t row_hi);
int scs_set_batches(scs_problem* p, int64_t nbatch, const int64_t* offsets);
/* Held-out data (Problem(...; Atest, ytest), src/problems.jl:28-29): `test` is a second problem created from this
 * rank's rows of (Atest, ytest) with the same loss.  scs_solve then records ftest(x) = model.f(Atest, ytest, x) next to
 * every history entry (iterate.jl:169-176, utils.jl:55-57); scs_get_test_history copies them out (n = entries).  The
 * host-driven loop simply calls scs_loss_eval on the test problem.  test = NULL detaches. */
int scs_set_test_problem(scs_problem* p, scs_problem* test);
int scs_get_test_history(scs_problem* p, double* out, int64_t cap, int64_t* n);
/* Streaming-pass selection for "objective + gradient at the same x": 0 = auto (the single-pass cluster kernel
 * k_fused_grad when m <= 4096, else two passes), 1 = always two passes (k_forward + k_adjoint), 2 = fused, and
 * SCS_UNSUPPORTED if the shape has no fused kernel.  scs_get_stream_path reports what the last gradient used
 * (1 = two passes, 2 = fused). */
int scs_set_stream_mode(scs_problem* p, int mode);
int scs_get_stream_path(scs_problem* p, int* path);
int scs_method_init(scs_problem* p);

/* ---- the two call sites of optim_loop! ------------------------------------------------- */
int scs_objective(scs_problem* p, const double* x, double* fval, double* reg);
/* dx may be NULL (return_dx=false).  x, x_prev, x_new, dx: host buffers of m doubles. */
int scs_step(scs_problem* p, const double* x, const double* x_prev, int64_t iter, double* x_new, double* dx,
             double* pri_res_norm);

/* ---- whole loop on the device (x never leaves HBM between iterations) ------------------- */
/* Histories have capacity max_epoch+2 (iterate.jl pushes once per epoch plus one closing entry).
 * pri_res_norm[0] is NaN where the reference stores `nothing`. */
int scs_solve(scs_problem* p, const double* x0, const double* x_star, int64_t max_epoch, double x_tol,
              double f_tol, double* x_out, double* obj, double* fval, double* pri_res_norm, double* rel,
              double* objrel, int64_t* n_hist, int64_t* epochs);

/* ---- component entry points (parity tests, partial adoption) ---------------------------- */
/* weight_kind selects (r,w): NEWTON = d f/dz, d²f/dz² of f(A,y,x); GGN = J'res / J'QJ weights of the
 * (out_fn, f(y,ŷ)) pair.  Any output pointer may be NULL.  grad is the allreduced A'r (length m);
 * z, r, w are this rank's n_local rows. */
int scs_loss_eval(scs_problem* p, const double* x, int weight_kind, double* fval, double* grad, double* z,
                  double* r, double* w);
/* G = A'diag(w)A (allreduced), full symmetric m x m column-major. */
int scs_gram(scs_problem* p, const double* x, int weight_kind, double* G);
/* Solve M d = b for symmetric M (m x m, column-major, lower triangle read).  Cholesky on the device; if a
 * pivot is not positive the pivoted-LU fallback runs (also on the device) and *used_fallback = 1. */
int scs_linear_solve(scs_ctx* ctx, const double* M, const double* b, int64_t m, double* d, int* used_fallback);
int scs_smoother_eval(scs_problem* p, const double* x, double* gr, double* hr);
int scs_prox(scs_problem* p, const double* u, const double* hr, double ss, double* out);
int scs_reg_value(scs_problem* p, const double* x, double* out);

/* ---- instrumentation ------------------------------------------------------------------- */
/* Kernel launches issued by this context since creation / last reset, and device milliseconds accumulated per
 * stage while profiling is on.  stage ids: 0 forward, 1 adjoint, 2 gram, 3 solve, 4 vector, 5 allreduce,
 * 6 fused forward+adjoint (single pass), 7 gram finalize (split sum + mirror / CRT), 8 residue planes of the
 * emulated-fp64 Gram, 9 reserved.  ms / calls: arrays of SCS_NUM_STAGES entries. */
#define SCS_NUM_STAGES 10
int scs_get_counters(scs_ctx* ctx, int64_t* launches, int reset);
int scs_set_profiling(scs_ctx* ctx, int enable);
int scs_get_stage_ms(scs_ctx* ctx, double* ms, int64_t* calls, int reset);
/* Measurement aid: the int8 tensor-pipe rate of this device with the library's own 128x256x32 kind::i8 UMMA on operands
 * resident in shared memory (no memory traffic) — the denominator bench.py uses for k_i8syrk's roofline.  burst = best of
 * five ~20 ms launches, sustained = back-to-back launches for `seconds`.  TOP/s. */
int scs_measure_i8_peak(scs_ctx* ctx, double seconds, double* tops_burst, double* tops_sustained);
/* Tuning aid: TOP/s of k_i8syrk's main loop (4-stage TMA -> UMMA ring, one CTA per SM, no cluster / epilogue / lock-step)
 * on an operand that stays in L2.  tma_mode 0: operands resident in shared memory, 1: the A slab (16 KB per stage) through
 * TMA, 2: A and B slabs (48 KB per stage).  Separates the SM-level pipeline limit from DRAM / multicast effects. */
int scs_i8_pipe_probe(scs_ctx* ctx, int tma_mode, double* tops);

#ifdef __cplusplus
}
#endif
#endif /* SCS_B200_H */
