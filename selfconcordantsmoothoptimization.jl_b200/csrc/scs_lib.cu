// scs_b200 — C-ABI host layer over the sm_100a kernels (see include/scs_b200.h for the contract and the
// reference file:line each entry replaces).  One context per (process, GPU); all work of a context is issued on
// its own stream; row-sharded problems reduce [g ‖ loss] and the Gram with NCCL fp64 sum all-reduces.
#include "../../include/scs_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_gram.cuh"
#include "kernels_i8gram.cuh"
#include "kernels_fused.cuh"
#include "kernels_solve.cuh"
#include "kernels_sparse.cuh"
#include "kernels_stream.cuh"
#include "kernels_vec.cuh"
#include "nccl_dyn.hpp"
#include "synth.cuh"

using namespace scs;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(e_ == cudaErrorMemoryAllocation ? SCS_OOM : SCS_CUDA_ERROR,                     \
                  std::string(#expr) + " failed: " + cudaGetErrorString(e_));                     \
  } while (0)
#define SCS_TRY(expr)          \
  do {                         \
    int s_ = (expr);           \
    if (s_ != SCS_OK) return s_; \
  } while (0)

static NcclApi g_nccl;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum {
  ST_FWD = 0, ST_ADJ = 1, ST_GRAM = 2, ST_SOLVE = 3, ST_VEC = 4, ST_COMM = 5, ST_FUSED = 6, ST_GRAMFIN = 7, ST_RESID = 8,
  ST_N = SCS_NUM_STAGES
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct scs_ctx {
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // trailing updates of the look-ahead Cholesky (run_solve)
  cudaEvent_t ev_trsm[2] = {nullptr, nullptr}, ev_upd[2] = {nullptr, nullptr};
  cudaEvent_t ev_slab[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // slab-pipelined Gram all-reduce
  int solve_mode = 0;  // 0 = look-ahead sequence, 1 = k_panel / k_syrk_update sequence (SCS_SOLVE_LEGACY=1)
  int solve_pair = -1;  // look-ahead sequence with the trailing update applied two panels at a time: -1 = by size
                        // (m > kPairMinM), 0 / 1 forced (SCS_SOLVE_PAIR)
  bool solve_attr_set = false;
  int p2p_capacity = -1;
  size_t sp_gram_smem = 0;
  NcclComm comm = nullptr;
  int num_sms = 148;
  int64_t launches = 0;
  bool profiling = false;
  double stage_ms[ST_N] = {0};
  int64_t stage_calls[ST_N] = {0};
  struct Pending {
    int stage;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> ev_pool;
  EncodeTiledFn encode = nullptr;
  double* d_flag = nullptr;  // one device double: cross-rank agreement on a status (agree_ok)
};

struct StageTimer {  // records an event pair around a stage when profiling is on
  scs_ctx* c;
  int stage;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  StageTimer(scs_ctx* c_, int s, cudaStream_t stream = nullptr) : c(c_), stage(s), st(stream ? stream : c_->stream) {
    if (!c->profiling) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->ev_pool.empty()) {
        e = c->ev_pool.back();
        c->ev_pool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, st);
  }
  ~StageTimer() {
    if (!c->profiling) return;
    cudaEventRecord(b, st);
    c->pending.push_back({stage, a, b});
  }
};
static void harvest_timers(scs_ctx* c) {  // call after a stream synchronize
  for (auto& p : c->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      c->stage_ms[p.stage] += ms;
      c->stage_calls[p.stage] += 1;
    }
    c->ev_pool.push_back(p.a);
    c->ev_pool.push_back(p.b);
  }
  c->pending.clear();
}

#define LAUNCH(ctx, kern, grid, block, smem, ...)                                   \
  do {                                                                              \
    kern<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);                        \
    (ctx)->launches += 1;                                                           \
    cudaError_t le_ = cudaGetLastError();                                           \
    if (le_ != cudaSuccess)                                                         \
      return fail(SCS_CUDA_ERROR, std::string(#kern) + " launch: " + cudaGetErrorString(le_)); \
  } while (0)

static int ctx_sync(scs_ctx* c) {
  CU_TRY(cudaStreamSynchronize(c->stream));
  harvest_timers(c);
  return SCS_OK;
}

static int allreduce(scs_ctx* c, double* buf, size_t count, cudaStream_t stream = nullptr) {
  if (c->world <= 1) return SCS_OK;
  if (!stream) stream = c->stream;
  StageTimer t(c, ST_COMM, stream);
  int r = g_nccl.AllReduce(buf, buf, count, NcclApi::kFloat64, NcclApi::kSum, c->comm, stream);
  if (r != 0) return fail(SCS_NCCL_ERROR, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  return SCS_OK;
}
// A status that can differ from rank to rank (out of memory, an unsupported shape) must be agreed on BEFORE the next
// collective: a rank that returned early would leave its peers blocked inside NCCL.  *all_ok = every rank passed.
static int agree_ok(scs_ctx* c, bool local_ok, bool* all_ok) {
  *all_ok = local_ok;
  if (c->world <= 1) return SCS_OK;
  const double v = local_ok ? 0.0 : 1.0;
  double tot = 0.0;
  CU_TRY(cudaMemcpyAsync(c->d_flag, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SCS_TRY(allreduce(c, c->d_flag, 1));
  CU_TRY(cudaMemcpyAsync(&tot, c->d_flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  *all_ok = tot == 0.0;
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// problem
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  double* p = nullptr;
  size_t n = 0;
};

struct scs_problem {
  scs_ctx* ctx = nullptr;
  int64_t n = 0, m = 0, ldd = 0, mp = 0;  // mp = m rounded up to 16 (padded vector length)
  // active row window (mini-batches, iterate.jl:204-207): rows [win_lo, win_hi) of the shard take part in the passes;
  // kernels sweep the 128-row-aligned superset [alo, ahi) and mask the rest.  Default: the whole shard.
  int64_t win_lo = 0, win_hi = 0, alo = 0, ahi = 0;
  int64_t win_rows_global = 0;  // rows of the window summed over ranks
  std::vector<int64_t> batch_off;  // local offsets of the mini-batches scs_solve iterates over (empty: full batch)
  std::vector<int64_t> batch_rows_global;  // rows of each batch summed over ranks; last entry: the whole problem
  double *dA = nullptr, *dy = nullptr, *dz = nullptr, *dr = nullptr, *dw = nullptr;
  LossParams loss{};
  // regulariser / smoother
  bool has_reg = false, has_sm = false, has_method = false;
  RegDesc reg{};
  SmoothDesc sm{};
  double Mh = 0, nu = 0;
  double *d_rlb = nullptr, *d_rub = nullptr, *d_slb = nullptr, *d_sub = nullptr, *d_cdiag = nullptr;
  int64_t *d_ind = nullptr, *d_perm = nullptr;
  // method
  int method = -1, ss_type = 1, use_prox = 1, lbfgs_m = 10;
  bool has_L = false;
  double L = 0;
  // m-vectors (each mp doubles, zero padded)
  double *vx[3] = {nullptr, nullptr, nullptr};  // rotating x buffers
  double *d_gl = nullptr;                       // [g (m) ‖ loss sum (1)] all-reduce buffer, mp+16
  double *d_gr = nullptr, *d_hr = nullptr, *d_rhs = nullptr, *d_sol = nullptr, *d_d = nullptr, *d_dx = nullptr,
         *d_delta = nullptr, *d_gq = nullptr, *d_gqprev = nullptr, *d_gamma = nullptr, *d_q = nullptr,
         *d_t1 = nullptr, *d_t2 = nullptr, *d_xstar = nullptr, *d_trial = nullptr, *d_gnewton = nullptr;
  double* d_scal = nullptr;
  double* h_scal = nullptr;  // pinned
  // stream workspaces
  int64_t fwd_blocks = 0, adj_blocks = 0;
  double *d_losspart = nullptr, *d_adjpart = nullptr;
  // gram / solve
  double *d_G = nullptr, *d_Gsave = nullptr, *d_partial = nullptr, *d_Linv = nullptr;
  int* d_info = nullptr;
  GramPlan plan{};
  int ldp = 0;
  CUtensorMap amap{};
  bool gram_ready = false;
  // emulated-fp64 Gram on tcgen05 int8 (kernels_i8gram.cuh)
  int gram_mode = 0;   // 0 auto, 1 DMMA, 2 tcgen05 int8 when eligible (w >= 0), else DMMA
  int last_gram_path = 0;  // 1 DMMA, 2 int8
  bool i8_ready = false, i8_failed = false, i8_planes_valid = false;
  int8_t *d_planes = nullptr, *d_i8partial = nullptr;
  // signed weights: compacted planes of the minority-sign rows (i8_setup_signed), per-block statistics
  int8_t* d_cplanes = nullptr;
  double* d_Gpack = nullptr;  // several ranks: packed upper triangle of the Gram, what the all-reduce moves
  // own all-reduce over NVLink peer memory (k_p2p_*): d_Gpack is then an IPC-shared allocation [count | flag slots]
  bool p2p_tried = false, p2p_ok = false;
  P2PPeers p2p{};
  void* p2p_opened[kP2PMaxWorld] = {nullptr};
  unsigned long long p2p_epoch = 0;
  int64_t ldc = 0, i8_ccount = 0;
  int64_t* d_negbase = nullptr;
  double* d_wpart = nullptr;
  bool i8_signed = false, i8_minor_neg = true;
  int i8_pchunks_cap = 0;
  CUtensorMap cmap{}, cmap_b{};
  bool labels_unit = false;  // every label lies in [-1, 1] (checked at upload): consistent-mode GGN weights are >= 0
  double *d_colmax = nullptr, *d_wstat = nullptr, *d_colscale = nullptr;
  double *d_colinv = nullptr, *d_colnorm2 = nullptr;
  double i8_T = 0.0;  // column 2-norm target of the fixed-point image (run_gram_i8)
  int2* d_i8tiles = nullptr;
  unsigned long long* d_i8progress = nullptr;
  int64_t ldx = 0;
  int i8_b = 0, i8_clusters = 0, i8_bits = 46, i8_nmod = 0;
  I8Plan i8plan{};
  CUtensorMap xmap{}, xmap_b{};
  // single-pass fused gradient (kernels_fused.cuh)
  int stream_mode = 0;       // 0 auto, 1 two passes (k_forward + k_adjoint), 2 fused whenever the shape is supported
  int last_stream_path = 0;  // 1 two passes, 2 fused
  bool fu_ready = false, fu_failed = false;
  int fu_cluster = 1, fu_clusters = 0, fu_last_ncl = 0;
  int64_t fu_col0 = 0;  // m > 4096: the cluster kernel covers columns [fu_col0, m)
  CUtensorMap fumap{};
  double *d_fupart = nullptr, *d_fuloss = nullptr;
  double* d_u = nullptr;  // row vector of the GGN wide branch / A d of the line search (ldd doubles, allocated on first use)
  double *d_wide = nullptr, *d_widepart = nullptr, *d_wcnt = nullptr;  // replicated batch of the GGN wide branch
  size_t wide_cap = 0, widepart_cap = 0;
  double* d_lspart = nullptr;  // line search: per-block trial sums | 8 sums | f(x), <∇q,d>
  int64_t lspart_cap = 0;
  // sparse shard (kernels_sparse.cuh): CSR + CSC copies, no dense A
  bool sparse = false;
  int64_t nnz = 0;
  int64_t *d_rowptr = nullptr, *d_colptr = nullptr;
  int *d_colidx = nullptr, *d_rowidx = nullptr;
  double *d_vals = nullptr, *d_cvals = nullptr, *d_spscale = nullptr;
  // held-out data (model.Atest / model.ytest): a second resident shard whose loss is recorded next to every history
  // entry of scs_solve (ftest, iterate.jl:169-176; utils.jl:55-57)
  scs_problem* test = nullptr;
  std::vector<double> test_hist;
  // l-bfgs
  double *d_S = nullptr, *d_Y = nullptr;
  int64_t* d_state = nullptr;
  int lbfgs_cap = 0;
  // caches, keyed by x-vector ids
  uint64_t next_id = 1;
  uint64_t fwd_id = 0;
  int fwd_wk = -1;
  uint64_t grad_id = 0;  // d_gl[0..m) = allreduced A'r for this id (and fwd_wk)
  uint64_t gq_id = 0;    // d_gq = ∇q at this id
  uint64_t gqprev_id = 0;
  bool loss_reduced = false;  // d_gl[m] already all-reduced for fwd_id
  struct Shadow {
    std::vector<double> x;
    uint64_t id = 0;
  };
  Shadow shadow[4];
  int shadow_next = 0;
  int last_used_fallback = 0;
};

struct XRef {
  double* d;
  uint64_t id;
};

static int dalloc(double** p, size_t n) {
  CU_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(double)));
  CU_TRY(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(double)));
  // the memset runs on the legacy default stream, which the contexts' non-blocking streams do not wait for: finish it
  // before anything is queued on them (set-up path only)
  CU_TRY(cudaStreamSynchronize(0));
  return SCS_OK;
}
static void dfree(void* p) {
  if (p) cudaFree(p);
}

static int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// stream passes
// ------------------------------------------------------------------------------------------------
static int run_forward(scs_problem* p, const double* dx, int wk) {
  scs_ctx* c = p->ctx;
  StageTimer t(c, ST_FWD);
  LossParams lp = p->loss;
  lp.weight_kind = wk;
  if (p->sparse) {  // all rows, the mini-batch window is a mask
    const int64_t blocks = (p->n + kSpFwdRows - 1) / kSpFwdRows;
    LAUNCH(c, k_sp_forward, (unsigned)blocks, kSpFwdThreads, 0, (const int64_t*)p->d_rowptr, (const int*)p->d_colidx,
           (const double*)p->d_vals, p->n, p->win_lo, p->win_hi, dx, (const double*)p->dy, lp, p->dz, p->dr, p->dw,
           p->d_losspart);
    LAUNCH(c, k_sum_partials, 1, kVecThreads, 0, p->d_losspart, blocks, p->d_gl + p->m);
    return SCS_OK;
  }
  const int64_t nproc = p->ahi - p->alo;
  const int64_t blocks = (nproc + kFwdRows - 1) / kFwdRows;
  if (blocks > 0)
    LAUNCH(c, k_forward<8>, (unsigned)blocks, kFwdThreads, 0, p->dA + p->alo, p->ldd, nproc, p->win_lo - p->alo,
           p->win_hi - p->alo, (int)p->m, dx, p->dy + p->alo, lp, p->dz + p->alo, p->dr + p->alo, p->dw + p->alo,
           p->d_losspart);
  LAUNCH(c, k_sum_partials, 1, kVecThreads, 0, p->d_losspart, blocks, p->d_gl + p->m);
  return SCS_OK;
}
static int run_adjoint(scs_problem* p, const double* dr, double* dout) {
  scs_ctx* c = p->ctx;
  StageTimer t(c, ST_ADJ);
  if (p->sparse) {
    LAUNCH(c, k_sp_adjoint, (unsigned)((p->m + 7) / 8), 256, 0, (const int64_t*)p->d_colptr, (const int*)p->d_rowidx,
           (const double*)p->d_cvals, (int)p->m, dr, dout);
    return SCS_OK;
  }
  const int64_t nproc = p->ahi - p->alo;
  const int64_t blocks = (nproc + 64 * 8 - 1) / (64 * 8);
  if (blocks > 0)
    LAUNCH(c, k_adjoint<8>, (unsigned)blocks, kAdjThreads, 0, p->dA + p->alo, p->ldd, nproc, (int)p->m, dr + p->alo,
           p->d_adjpart);
  LAUNCH(c, k_colsum, (unsigned)((p->m + 31) / 32), 256, 0, p->d_adjpart, blocks, (int)p->m, dout);
  return SCS_OK;
}

// z_out = A v on the rows of the active window (no loss pieces): the line search's A d
static int run_matvec(scs_problem* p, const double* dv, double* zout) {
  scs_ctx* c = p->ctx;
  StageTimer t(c, ST_FWD);
  LossParams lp = p->loss;
  lp.kind = 2;  // z only
  lp.weight_kind = 0;
  if (p->sparse) {
    const int64_t blocks = (p->n + kSpFwdRows - 1) / kSpFwdRows;
    LAUNCH(c, k_sp_forward, (unsigned)blocks, kSpFwdThreads, 0, (const int64_t*)p->d_rowptr, (const int*)p->d_colidx,
           (const double*)p->d_vals, p->n, p->win_lo, p->win_hi, dv, (const double*)p->dy, lp, zout, (double*)nullptr,
           (double*)nullptr, p->d_losspart);
    return SCS_OK;
  }
  const int64_t nproc = p->ahi - p->alo;
  const int64_t blocks = (nproc + kFwdRows - 1) / kFwdRows;
  if (blocks > 0)
    LAUNCH(c, k_forward<8>, (unsigned)blocks, kFwdThreads, 0, p->dA + p->alo, p->ldd, nproc, p->win_lo - p->alo,
           p->win_hi - p->alo, (int)p->m, dv, p->dy + p->alo, lp, zout + p->alo, (double*)nullptr, (double*)nullptr,
           p->d_losspart);
  return SCS_OK;
}

static int allreduce(scs_ctx* c, double* buf, size_t count, cudaStream_t stream);
// Select the rows that take part in the following passes.  Everything cached for the previous window is dropped.
static int set_window(scs_problem* p, int64_t lo, int64_t hi, int64_t rows_global = -1) {
  if (lo < 0 || hi < lo || hi > p->n) return fail(SCS_INVALID_ARG, "row window outside the shard");
  // Unchanged window: nothing to recompute — but with several ranks and no global row count from the caller the
  // all-reduce below must still be entered by EVERY rank, whatever its local window looks like (a rank whose slice of
  // two consecutive batches is empty at the same offset would otherwise skip it and hang the others).
  const bool same = lo == p->win_lo && hi == p->win_hi && p->win_rows_global > 0;
  if (same && (rows_global >= 0 || p->ctx->world == 1)) return SCS_OK;
  if (!same) {
    p->win_lo = lo;
    p->win_hi = hi;
    p->alo = lo / 128 * 128;
    p->ahi = hi > lo ? std::min(round_up(hi, 128), p->ldd) : p->alo;
    p->fwd_id = 0;
    p->grad_id = 0;
    p->gq_id = 0;
    p->gqprev_id = 0;
    p->loss_reduced = false;
    p->i8_planes_valid = false;
  }
  p->win_rows_global = rows_global >= 0 ? rows_global : hi - lo;
  if (p->ctx->world > 1 && rows_global < 0) {  // the GGN wide-branch test needs the global batch size
    const double v = (double)(hi - lo);
    CU_TRY(cudaMemcpyAsync(p->d_scal + SC_ALLOC - 1, &v, sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    SCS_TRY(allreduce(p->ctx, p->d_scal + SC_ALLOC - 1, 1, nullptr));
    double tot = 0;
    CU_TRY(cudaMemcpyAsync(&tot, p->d_scal + SC_ALLOC - 1, sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
    p->win_rows_global = (int64_t)(tot + 0.5);
  }
  return SCS_OK;
}

// ---- single-pass fused gradient ---------------------------------------------------------------------------------
static void fused_config(scs_problem* p, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int nclusters) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3((unsigned)(nclusters * p->fu_cluster));
  cfg->blockDim = dim3(kFuThreads);
  cfg->dynamicSmemBytes = kFuSmemBytes;
  cfg->stream = p->ctx->stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p->fu_cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

// The shape decides the cluster geometry: CTA rank c owns columns [256c, 256c + 256); at most 16 CTAs per cluster.
// m > 4096: the cluster covers the last 4096 columns [col0, m), the columns before col0 go through k_forward (z) and
// k_adjoint (g) — 1.5 reads of A per objective + gradient instead of 2 (kernels_fused.cuh).
static bool fused_shape(const scs_problem* p, int* cluster, int64_t* col0) {
  if (p->loss.kind == SCS_LOSS_QUADFORM || p->sparse) return false;
  if (p->ldd >= (int64_t)1 << 31) return false;  // TMA coordinates are 32-bit
  const int64_t cap = (int64_t)kFuMaxCluster * kFuCols;
  *col0 = p->m > cap ? p->m - cap : 0;
  *cluster = (int)((p->m - *col0 + kFuCols - 1) / kFuCols);
  return true;
}

static int fused_setup(scs_problem* p) {
  if (p->fu_ready) return SCS_OK;
  scs_ctx* c = p->ctx;
  if (!fused_shape(p, &p->fu_cluster, &p->fu_col0)) {
    p->fu_failed = true;
    return fail(SCS_UNSUPPORTED, "fused gradient pass: shape not supported (quadform loss or sparse shard)");
  }
  if (!c->encode) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)p->ldd, (cuuint64_t)(p->m - p->fu_col0)};
  cuuint64_t gstride[1] = {(cuuint64_t)p->ldd * 8};
  cuuint32_t box[2] = {(cuuint32_t)kFuStageRows, (cuuint32_t)kFuBoxCols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(&p->fumap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, p->dA + p->fu_col0 * p->ldd, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled (fused pass) failed: " + std::to_string((int)r));
  CU_TRY(cudaFuncSetAttribute(k_fused_grad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes));
  CU_TRY(cudaFuncSetAttribute(k_fused_grad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes));
  if (p->fu_cluster > 8) {
    CU_TRY(cudaFuncSetAttribute(k_fused_grad<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU_TRY(cudaFuncSetAttribute(k_fused_grad<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fused_config(p, &cfg, attr, c->num_sms / p->fu_cluster);
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, k_fused_grad<false>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    ncl = 0;
  }
  if (ncl < 1) {
    p->fu_failed = true;
    return fail(SCS_UNSUPPORTED, "fused gradient pass: no cluster of " + std::to_string(p->fu_cluster) + " CTAs can be resident");
  }
  p->fu_clusters = ncl;
  SCS_TRY(dalloc(&p->d_fupart, (size_t)ncl * p->m));
  SCS_TRY(dalloc(&p->d_fuloss, (size_t)ncl * p->fu_cluster));
  if (p->fu_col0 > 0 && !p->d_u) SCS_TRY(dalloc(&p->d_u, p->ldd));
  p->fu_ready = true;
  return SCS_OK;
}

static bool fused_wanted(scs_problem* p) {
  if (p->stream_mode == 1 || p->fu_failed) return false;
  int cl;
  int64_t c0;
  if (p->stream_mode == 0 && !p->fu_ready && !fused_shape(p, &cl, &c0)) return false;
  return true;
}

// one read of A: z, r, w, the local loss sum in d_gl[m] and the local A'r in d_gl[0..m)
// reduce = false: the per-cluster partial gradients / per-CTA loss sums are left for the caller's kernel to fold
// (k_lqn_update on one GPU); p->fu_last_ncl = clusters launched.
static int run_fused(scs_problem* p, const double* dx, int wk, bool reduce = true) {
  scs_ctx* c = p->ctx;
  SCS_TRY(fused_setup(p));
  StageTimer t(c, ST_FUSED);
  LossParams lp = p->loss;
  lp.weight_kind = wk;
  const int64_t npanels = (p->ahi - p->alo) / kFuRows;
  const int ncl = (int)std::max<int64_t>(1, std::min<int64_t>(p->fu_clusters, (npanels + 1) / 2));
  const int64_t col0 = p->fu_col0;
  const int mf = (int)(p->m - col0);  // columns the cluster kernel covers
  const double* zin = nullptr;
  const int64_t nproc = p->ahi - p->alo;
  if (col0 > 0) {  // z1 = A[:, :col0] x[:col0] for the rows of the window (rows outside it come out as 0)
    LossParams lz = p->loss;
    lz.kind = 2;
    lz.weight_kind = 0;
    const int64_t blocks = (nproc + kFwdRows - 1) / kFwdRows;
    if (blocks > 0)
      LAUNCH(c, k_forward<8>, (unsigned)blocks, kFwdThreads, 0, p->dA + p->alo, p->ldd, nproc, p->win_lo - p->alo,
             p->win_hi - p->alo, (int)col0, dx, p->dy + p->alo, lz, p->d_u + p->alo, (double*)nullptr, (double*)nullptr,
             p->d_losspart);
    zin = p->d_u;
  }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fused_config(p, &cfg, attr, ncl);
  cudaError_t le;
  if (getenv("SCS_FUSED_PROF")) {  // tuning aid: per-CTA wait-cycle counters printed to stderr (synchronises)
    long long* dprof = nullptr;
    const size_t np = (size_t)cfg.gridDim.x * 8;
    CU_TRY(cudaMalloc((void**)&dprof, np * sizeof(long long)));
    CU_TRY(cudaMemsetAsync(dprof, 0, np * sizeof(long long), c->stream));
    le = cudaLaunchKernelEx(&cfg, k_fused_grad<true>, p->fumap, dx + col0, (const double*)p->dy, lp, p->alo, p->win_lo,
                            p->win_hi, npanels, mf, p->dz, p->dr, p->dw, p->d_fuloss, p->d_fupart, dprof,
                            atoi(getenv("SCS_FUSED_PROF")) == 2 ? 1 : 0, zin);
    std::vector<long long> h(np);
    cudaMemcpyAsync(h.data(), dprof, np * sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(dprof);
    double s[8] = {0};
    for (size_t i = 0; i < np; ++i) s[i % 8] += (double)h[i];
    const double nb = (double)cfg.gridDim.x;
    const double panels = (double)npanels / ncl;
    fprintf(stderr,
            "[k_fused_grad prof] clusters %d x %d, panels/cluster %.0f; mean cycles per CTA: total %.0f (%.0f per panel) | "
            "warp0: wait full %.0f, wait rbar %.0f, busy A %.0f | producer wait empty %.0f | push wait pbar %.0f | loss: "
            "wait zbar %.0f, busy %.0f\n",
            ncl, p->fu_cluster, panels, s[0] / nb, s[0] / nb / panels, s[1] / nb, s[2] / nb, s[7] / nb, s[3] / nb,
            s[4] / nb, s[5] / nb, s[6] / nb);
  } else {
    le = cudaLaunchKernelEx(&cfg, k_fused_grad<false>, p->fumap, dx + col0, (const double*)p->dy, lp, p->alo, p->win_lo,
                            p->win_hi, npanels, mf, p->dz, p->dr, p->dw, p->d_fuloss, p->d_fupart,
                            (long long*)nullptr, 0, zin);
  }
  c->launches += 1;
  if (le != cudaSuccess) return fail(SCS_CUDA_ERROR, std::string("k_fused_grad launch: ") + cudaGetErrorString(le));
  p->fu_last_ncl = ncl;
  if (col0 > 0) {  // g[:col0] = A[:, :col0]' r
    const int64_t blocks = (nproc + 64 * 8 - 1) / (64 * 8);
    if (blocks > 0)
      LAUNCH(c, k_adjoint<8>, (unsigned)blocks, kAdjThreads, 0, p->dA + p->alo, p->ldd, nproc, (int)col0,
             (const double*)(p->dr + p->alo), p->d_adjpart);
    LAUNCH(c, k_colsum, (unsigned)((col0 + 31) / 32), 256, 0, p->d_adjpart, blocks, (int)col0, p->d_gl);
    reduce = true;  // the caller's kernel only knows the single-launch layout
  }
  if (reduce) {
    LAUNCH(c, k_colsum, (unsigned)((mf + 31) / 32), 256, 0, p->d_fupart, (int64_t)ncl, mf, p->d_gl + col0);
    LAUNCH(c, k_sum_partials, 1, kVecThreads, 0, p->d_fuloss, (int64_t)ncl * p->fu_cluster, p->d_gl + p->m);
  }
  p->last_stream_path = 2;
  return SCS_OK;
}

// forward pass at x (cached by id and weight kind); leaves z, r, w on the device and the local loss sum in d_gl[m]
static int ensure_forward(scs_problem* p, XRef x, int wk) {
  if (p->fwd_id == x.id && p->fwd_wk == wk) return SCS_OK;
  SCS_TRY(run_forward(p, x.d, wk));
  p->fwd_id = x.id;
  p->fwd_wk = wk;
  p->grad_id = 0;
  p->loss_reduced = false;
  return SCS_OK;
}
// loss sum all-reduced (objective only)
static int ensure_loss(scs_problem* p, XRef x, int wk) {
  SCS_TRY(ensure_forward(p, x, wk));
  if (!p->loss_reduced) {
    if (p->loss.kind == SCS_LOSS_QUADFORM) {
      StageTimer t(p->ctx, ST_VEC);
      LAUNCH(p->ctx, k_quadform_value, 1, kVecThreads, 0, x.d, p->dz, p->dy, (int)p->m, p->d_gl + p->m);
    } else {
      SCS_TRY(allreduce(p->ctx, p->d_gl + p->m, 1));
    }
    p->loss_reduced = true;
  }
  return SCS_OK;
}
// gradient of f at x into d_gl[0..m) (all-reduced together with the loss sum)
static int ensure_grad(scs_problem* p, XRef x, int wk) {
  if (!(p->fwd_id == x.id && p->fwd_wk == wk) && fused_wanted(p)) {
    int rc = run_fused(p, x.d, wk);
    if (rc == SCS_OK) {
      p->fwd_id = x.id;
      p->fwd_wk = wk;
      SCS_TRY(allreduce(p->ctx, p->d_gl, p->m + 1));
      p->loss_reduced = true;
      p->grad_id = x.id;
      return SCS_OK;
    }
    if (!(rc == SCS_UNSUPPORTED && p->stream_mode == 0)) return rc;  // auto mode: fall through to the two-pass kernels
  }
  SCS_TRY(ensure_forward(p, x, wk));
  if (p->grad_id == x.id) return SCS_OK;
  p->last_stream_path = 1;
  scs_ctx* c = p->ctx;
  if (p->loss.kind == SCS_LOSS_QUADFORM) {
    // g = 0.5*(A x + A' x) + y ; the adjoint pass takes r := x (zero padded to ldd)
    CU_TRY(cudaMemsetAsync(p->dr, 0, p->ldd * sizeof(double), c->stream));
    CU_TRY(cudaMemcpyAsync(p->dr, x.d, p->m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    SCS_TRY(run_adjoint(p, p->dr, p->d_t1));
    StageTimer t(c, ST_VEC);
    LAUNCH(c, k_quadform_grad, (unsigned)((p->m + 255) / 256), 256, 0, p->dz, p->d_t1, p->dy, (int)p->m, p->d_gl);
    if (!p->loss_reduced) {
      LAUNCH(c, k_quadform_value, 1, kVecThreads, 0, x.d, p->dz, p->dy, (int)p->m, p->d_gl + p->m);
      p->loss_reduced = true;
    }
  } else {
    SCS_TRY(run_adjoint(p, p->dr, p->d_gl));
    if (!p->loss_reduced) {
      SCS_TRY(allreduce(c, p->d_gl, p->m + 1));
      p->loss_reduced = true;
    } else {
      SCS_TRY(allreduce(c, p->d_gl, p->m));
    }
  }
  p->grad_id = x.id;
  return SCS_OK;
}

static double fval_from_sum(const scs_problem* p, double S) {
  switch (p->loss.kind) {
    case SCS_LOSS_LOGISTIC:
      return p->loss.p * S;
    case SCS_LOSS_LEASTSQUARES:
      return 0.5 * S / p->loss.p;
    default:
      return S;
  }
}

// ------------------------------------------------------------------------------------------------
// gram + solve
// ------------------------------------------------------------------------------------------------
static int gram_setup(scs_problem* p) {
  if (p->gram_ready) return SCS_OK;
  scs_ctx* c = p->ctx;
  const int64_t m = p->m;
  SCS_TRY(dalloc(&p->d_G, (size_t)m * m));
  SCS_TRY(dalloc(&p->d_Gsave, round_up(m, 16)));  // the diagonal (run_solve, sym = true)
  const int nblk = (int)((m + kNB - 1) / kNB);
  SCS_TRY(dalloc(&p->d_Linv, (size_t)nblk * kNB * kNB + round_up(m, 16) + 16 + nblk));  // 1/L_jj | inverted diagonal blocks | barrier
  CU_TRY(cudaMalloc((void**)&p->d_info, sizeof(int)));
  if (p->loss.kind != SCS_LOSS_QUADFORM && !p->sparse) {
    GramPlan& pl = p->plan;
    pl.nt = (int)((m + kGT - 1) / kGT);
    pl.ntiles = pl.nt * (pl.nt + 1) / 2;
    pl.kt = p->ldd / kGBK;
    p->ldp = (int)round_up(m, 2);
    // pick the K split that best fills whole waves of num_sms CTAs; each unit keeps >= 8 stages of work and
    // the partial buffers stay under 4 GiB
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const double per_split = (double)m * p->ldp * 8.0;
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= 32; ++s) {
      if (s > 1 && pl.kt / s < 8) break;
      if (s > 1 && per_split * s > std::min<double>(4.0 * (1ull << 30), 0.25 * (double)free_b)) break;
      const double units = (double)pl.ntiles * s;
      const double waves = std::ceil(units / c->num_sms);
      const double eff = units / (waves * c->num_sms);
      if (eff > best_eff + 0.004) {
        best_eff = eff;
        best = s;
      }
    }
    pl.splits = best;
    pl.units = (int64_t)pl.ntiles * pl.splits;
    SCS_TRY(dalloc(&p->d_partial, (size_t)pl.splits * m * p->ldp));
    if (!c->encode) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[2] = {(cuuint64_t)p->ldd, (cuuint64_t)m};
    cuuint64_t gstride[1] = {(cuuint64_t)p->ldd * 8};
    cuuint32_t box[2] = {(cuuint32_t)kGBK, (cuuint32_t)kGT};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = c->encode(&p->amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, p->dA, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    CU_TRY(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemBytes));
  }
  p->gram_ready = true;
  return SCS_OK;
}


// ---- emulated-fp64 Gram (tcgen05 int8 + CRT) ----------------------------------------------------------------
static int i8_setup(scs_problem* p) {
  if (p->i8_ready) return SCS_OK;
  scs_ctx* c = p->ctx;
  const int64_t m = p->m;
  p->ldx = round_up(p->ldd, kI8BK);
  // column statistics first: they decide the moduli count
  if (!p->d_colmax) SCS_TRY(dalloc(&p->d_colmax, m));
  if (!p->d_colnorm2) SCS_TRY(dalloc(&p->d_colnorm2, m));
  {
    StageTimer t(c, ST_FWD);
    LAUNCH(c, k_colabsmax, (unsigned)m, 256, 0, p->dA, p->ldd, p->n, (int)m, p->d_colmax, p->d_colnorm2);
  }
  std::vector<double> hmax(m), hn2(m);
  CU_TRY(cudaMemcpyAsync(hmax.data(), p->d_colmax, m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(hn2.data(), p->d_colnorm2, m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  // r_j = max_i |A_ij| / ||A_j||_2 in (0, 1]: the largest entry of column j of X is T r_j
  double rmax = 0.0;
  for (int64_t j = 0; j < m; ++j)
    if (hn2[j] > 0.0 && std::isfinite(hn2[j])) rmax = std::max(rmax, hmax[j] / std::sqrt(hn2[j]));
  if (rmax == 0.0) rmax = 1.0;  // A = 0
  // T_k: largest column norm the prefix of k moduli can hold: (T (1 + 1e-6) + sqrt(rows)/2)^2 < P_k / 2  (the 1e-6 covers
  // the rounding of the norms and scales, sqrt(rows)/2 the rint of every entry).  Entries must stay below 2^50 (exact
  // fp64 integer arithmetic in k_residues): T <= 2^50 / rmax.  Shortest prefix with T_k >= 2^bits, then all of that
  // prefix's range is used.
  const double T_cap = std::ldexp(1.0, 50) / rmax;
  const double T_req = std::min(std::ldexp(1.0, p->i8_bits), T_cap);
  auto T_of = [&](int k) {
    return (std::exp2((h_log2P[k - kNModMin] - 1.0) / 2.0) - 0.5 * std::sqrt((double)p->ldx)) * (1.0 - 1e-6);
  };
  int nmod = kNMod;
  for (int k = kNModMin; k <= kNMod; ++k)
    if (T_of(k) >= T_req) {
      nmod = k;
      break;
    }
  p->i8_nmod = nmod;
  p->i8_T = std::min(T_of(nmod), T_cap);
  p->i8_b = (int)std::floor(std::log2(p->i8_T));  // log2 of the column norm actually used
  const size_t plane_bytes = (size_t)nmod * p->ldx * m;
  I8Plan& pl = p->i8plan;
  pl.m = (int)m;
  pl.nmod = nmod;
  pl.kb_lo = 0;
  pl.kblocks = p->ldx / kI8BK;
  pl.chunk_kblocks = kI8ChunkRows / kI8BK;
  pl.nchunks = (int)((pl.kblocks + pl.chunk_kblocks - 1) / pl.chunk_kblocks);
  pl.ldp = (int)round_up(m, 16);
  // tile groups (g, bj): rows [g*C*128, (g+1)*C*128) x columns [bj*256, +256) that touch the lower triangle
  std::vector<int2> tiles;
  const int nbi = (int)((m + kI8BM - 1) / kI8BM), nbj = (int)((m + kI8BN - 1) / kI8BN);
  const int ngrp = (nbi + kI8Cluster - 1) / kI8Cluster;
  for (int g = 0; g < ngrp; ++g)
    for (int bj = 0; bj < nbj; ++bj)
      if ((int64_t)bj * kI8BN <= (int64_t)(g + 1) * kI8Cluster * kI8BM - 1) tiles.push_back(make_int2(g, bj));
  pl.ntiles = (int)tiles.size();
  pl.units = (int64_t)pl.nmod * pl.nchunks * pl.ntiles;
  const size_t partial_bytes = (size_t)nmod * pl.nchunks * m * pl.ldp;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  bool fits = plane_bytes + partial_bytes + (2ull << 30) <= free_b, all_fit = true;
  SCS_TRY(agree_ok(c, fits, &all_fit));  // every rank takes the same Gram path: the all-reduce of G follows
  if (!all_fit) {
    p->i8_failed = true;
    char msg[256];
    snprintf(msg, sizeof(msg), "not enough HBM for the int8 residue planes (%s: %.1f GB planes + %.1f GB partials, %.1f GB free)",
             fits ? "on another rank" : "this rank", plane_bytes / 1e9, partial_bytes / 1e9, free_b / 1e9);
    return fail(SCS_OOM, msg);
  }
  CU_TRY(cudaMalloc((void**)&p->d_planes, plane_bytes));
  CU_TRY(cudaMemsetAsync(p->d_planes, 0, plane_bytes, c->stream));
  CU_TRY(cudaMalloc((void**)&p->d_i8partial, partial_bytes));
  CU_TRY(cudaMemsetAsync(p->d_i8partial, 0, partial_bytes, c->stream));
  SCS_TRY(dalloc(&p->d_wstat, WS_COUNT));
  {
    const size_t nblk = (size_t)((p->ldd + kWsRows - 1) / kWsRows);
    SCS_TRY(dalloc(&p->d_wpart, 4 * nblk));
    CU_TRY(cudaMalloc((void**)&p->d_negbase, (nblk + 1) * sizeof(int64_t)));
  }
  p->i8_pchunks_cap = pl.nchunks;
  SCS_TRY(dalloc(&p->d_colscale, m));
  SCS_TRY(dalloc(&p->d_colinv, m));
  CU_TRY(cudaMalloc((void**)&p->d_i8tiles, tiles.size() * sizeof(int2)));
  CU_TRY(cudaMalloc((void**)&p->d_i8progress, sizeof(unsigned long long)));
  CU_TRY(cudaMemcpyAsync(p->d_i8tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  if (!c->encode) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[3] = {(cuuint64_t)p->ldx, (cuuint64_t)m, (cuuint64_t)nmod};
  cuuint64_t gstride[2] = {(cuuint64_t)p->ldx, (cuuint64_t)p->ldx * (cuuint64_t)m};
  cuuint32_t box[3] = {(cuuint32_t)kI8BK, 128u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode(&p->xmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, p->d_planes, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled (int8 planes) failed: " + std::to_string((int)r));
  cuuint32_t boxb[3] = {(cuuint32_t)kI8BK, (cuuint32_t)kI8BPart, 1u};
  r = c->encode(&p->xmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, p->d_planes, gdim, gstride, boxb, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled (int8 B slab) failed: " + std::to_string((int)r));
  CU_TRY(cudaFuncSetAttribute(k_i8syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes));
  CU_TRY(cudaFuncSetAttribute(k_i8syrk2, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8Smem2Bytes));
  p->i8_ready = true;
  return SCS_OK;
}

// Compacted planes for the minority-sign rows of a signed Gram (allocated on first use: at most half of the rows, +50 %
// plane memory) and a partial-residue buffer with room for their chunks.
static int i8_setup_signed(scs_problem* p) {
  if (p->d_cplanes) return SCS_OK;
  scs_ctx* c = p->ctx;
  const int64_t m = p->m;
  p->ldc = round_up(p->ldx / 2 + kI8BK, kI8BK);
  const I8Plan& pl = p->i8plan;
  const int cchunks = (int)((p->ldc / kI8BK + pl.chunk_kblocks - 1) / pl.chunk_kblocks);
  const size_t cbytes = (size_t)pl.nmod * p->ldc * m;
  const size_t partial_bytes = (size_t)pl.nmod * (pl.nchunks + cchunks) * m * pl.ldp;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t have = (size_t)pl.nmod * pl.nchunks * m * pl.ldp;  // the partial buffer is re-allocated
  if (cbytes + partial_bytes + (2ull << 30) > free_b + have)
    return fail(SCS_OOM, "not enough HBM for the compacted planes of a signed int8 Gram");
  CU_TRY(cudaStreamSynchronize(c->stream));
  dfree(p->d_i8partial);
  p->d_i8partial = nullptr;
  CU_TRY(cudaMalloc((void**)&p->d_i8partial, partial_bytes));
  CU_TRY(cudaMemsetAsync(p->d_i8partial, 0, partial_bytes, c->stream));
  CU_TRY(cudaMalloc((void**)&p->d_cplanes, cbytes));
  p->i8_pchunks_cap = pl.nchunks + cchunks;
  cuuint64_t gdim[3] = {(cuuint64_t)p->ldc, (cuuint64_t)m, (cuuint64_t)pl.nmod};
  cuuint64_t gstride[2] = {(cuuint64_t)p->ldc, (cuuint64_t)p->ldc * (cuuint64_t)m};
  cuuint32_t box[3] = {(cuuint32_t)kI8BK, 128u, 1u};
  cuuint32_t boxb[3] = {(cuuint32_t)kI8BK, (cuuint32_t)kI8BPart, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode(&p->cmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, p->d_cplanes, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = c->encode(&p->cmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, p->d_cplanes, gdim, gstride, boxb, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled (compacted planes) failed: " + std::to_string((int)r));
  return SCS_OK;
}

// Are the Gram weights of kind `wk` non-negative by construction?  Then nothing about them has to come back to the
// host before the kernels are queued.  Newton weights p*y^2*e/(1+e)^2, least squares 1/p, and the GGN weights of the
// consistent cross-entropy (yc = (y+1)/2 in [0,1]) are; the README's literal +-1-label pair is not.
static bool weights_nonneg_apriori(const scs_problem* p, int wk) {
  if (!(p->loss.p > 0.0)) return false;
  if (p->loss.kind == SCS_LOSS_LEASTSQUARES) return true;
  if (p->loss.kind != SCS_LOSS_LOGISTIC) return false;
  if (wk == SCS_WEIGHTS_NEWTON) return true;
  return p->loss.label_mode == SCS_LABELS_CONSISTENT && p->labels_unit;
}

static int i8_launch_syrk(scs_problem* p, const CUtensorMap& amap, const CUtensorMap& bmap, const I8Plan& pl) {
  scs_ctx* c = p->ctx;
  // SCS_I8_2CTA=1 selects the cta_group::2 SYRK (kernels_i8gram.cuh).  Measured equal to the 1-CTA kernel on this part
  // (C2: 104.0 vs 104.9 ms, both power-capped at 1.56 GHz; 400k x 2048: 11.1 vs 10.6 ms), so the default stays 1-CTA.
  static const bool two_cta = []() {
    const char* e = getenv("SCS_I8_2CTA");
    return e && e[0] == '1';
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(c->num_sms / kI8Cluster * kI8Cluster));
  cfg.blockDim = dim3(kI8Threads);
  cfg.dynamicSmemBytes = two_cta ? kI8Smem2Bytes : kI8SmemBytes;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kI8Cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: exactly as many clusters as can be co-resident (GPC shapes strand some SMs for clusters of 4)
  if (p->i8_clusters == 0) {
    int nc = 0;
    cudaError_t oe = two_cta ? cudaOccupancyMaxActiveClusters(&nc, k_i8syrk2, &cfg)
                             : cudaOccupancyMaxActiveClusters(&nc, k_i8syrk, &cfg);
    if (oe != cudaSuccess || nc < 1) nc = c->num_sms / kI8Cluster / 2;
    p->i8_clusters = nc;
  }
  const int ncl = (int)std::min<int64_t>(p->i8_clusters, pl.units);
  cfg.gridDim = dim3((unsigned)(ncl * kI8Cluster));
  CU_TRY(cudaMemsetAsync(p->d_i8progress, 0, sizeof(unsigned long long), c->stream));
  static const bool nolock = getenv("SCS_I8_NOLOCK") && atoi(getenv("SCS_I8_NOLOCK"));
  static const int slack = getenv("SCS_I8_SLACK") ? atoi(getenv("SCS_I8_SLACK")) : 0;
  I8Plan plv = pl;
  plv.lock_slack = slack;
  unsigned long long* prog = nolock ? nullptr : p->d_i8progress;
  cudaError_t le = two_cta ? cudaLaunchKernelEx(&cfg, k_i8syrk2, amap, bmap, plv, (const int2*)p->d_i8tiles,
                                                p->d_i8partial, prog)
                           : cudaLaunchKernelEx(&cfg, k_i8syrk, amap, bmap, plv, (const int2*)p->d_i8tiles,
                                                p->d_i8partial, prog);
  c->launches += 1;
  if (le != cudaSuccess) return fail(SCS_CUDA_ERROR, std::string("k_i8syrk launch: ") + cudaGetErrorString(le));
  return SCS_OK;
}

// Maps every rank's packed-Gram buffer into every process (cudaIpc) so that the exchange step can be our own kernels
// over NVLink instead of an NCCL call.  One-time, collective; if any rank cannot do it (no peer access, IPC not
// permitted in this container, SCS_P2P=0) every rank stays on NCCL.
static int p2p_setup(scs_problem* p, size_t count) {
  scs_ctx* c = p->ctx;
  if (p->p2p_tried) return SCS_OK;
  p->p2p_tried = true;
  const int W = c->world;
  static const bool disabled = getenv("SCS_P2P") && atoi(getenv("SCS_P2P")) == 0;
  bool ok = !disabled && W <= kP2PMaxWorld;
  double* base = nullptr;
  const size_t bytes = count * sizeof(double) + 4096;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    ok = cudaMalloc((void**)&base, bytes) == cudaSuccess && cudaMemset(base, 0, bytes) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine, base) == cudaSuccess;
    cudaGetLastError();
  }
  // exchange the 64-byte handles: one value per byte in a zero-filled [W x 64] table of doubles, summed by the all-reduce
  // (every slot has exactly one non-zero contributor, so the sum is exact)
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  std::vector<double> tab((size_t)W * (hb + 1), 0.0);
  if (ok) {
    const unsigned char* hbytes = reinterpret_cast<const unsigned char*>(&mine);
    for (size_t i = 0; i < hb; ++i) tab[(size_t)c->rank * (hb + 1) + i] = (double)hbytes[i];
    tab[(size_t)c->rank * (hb + 1) + hb] = 1.0;  // "I have a handle"
  }
  double* dtab = nullptr;
  CU_TRY(cudaMalloc((void**)&dtab, tab.size() * sizeof(double)));
  CU_TRY(cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int rc = allreduce(c, dtab, tab.size());
  if (rc == SCS_OK) {
    cudaMemcpyAsync(tab.data(), dtab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  cudaFree(dtab);
  SCS_TRY(rc);
  for (int r = 0; r < W && ok; ++r) ok = tab[(size_t)r * (hb + 1) + hb] == 1.0;
  if (ok) {
    for (int r = 0; r < W && ok; ++r) {
      if (r == c->rank) {
        p->p2p.base[r] = base;
      } else {
        cudaIpcMemHandle_t h;
        unsigned char* hbytes = reinterpret_cast<unsigned char*>(&h);
        for (size_t i = 0; i < hb; ++i) hbytes[i] = (unsigned char)(tab[(size_t)r * (hb + 1) + i] + 0.5);
        void* ptr = nullptr;
        ok = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        cudaGetLastError();
        p->p2p_opened[r] = ok ? ptr : nullptr;
        p->p2p.base[r] = (double*)ptr;
      }
      if (ok) p->p2p.flags[r] = reinterpret_cast<unsigned long long*>(p->p2p.base[r] + count);
    }
  }
  bool all_ok = false;
  SCS_TRY(agree_ok(c, ok, &all_ok));
  if (!all_ok) {
    for (int r = 0; r < W; ++r)
      if (p->p2p_opened[r]) {
        cudaIpcCloseMemHandle(p->p2p_opened[r]);
        p->p2p_opened[r] = nullptr;
      }
    if (base) cudaFree(base);
    if (c->rank == 0 && !disabled)
      fprintf(stderr, "[scs_b200] peer-memory all-reduce unavailable (cudaIpc / peer access): the Gram exchange uses NCCL\n");
    return SCS_OK;
  }
  dfree(p->d_Gpack);
  p->d_Gpack = base;
  p->p2p_ok = true;
  return SCS_OK;
}

// returns *done = 0 when this call cannot be served (non-finite weights, or no memory for the compacted planes of a
// signed Gram): the caller then runs the DMMA kernel
static int run_gram_i8(scs_problem* p, int* done) {
  scs_ctx* c = p->ctx;
  *done = 0;
  SCS_TRY(i8_setup(p));
  const int m = (int)p->m;
  // least squares: w = 1/denominator on every row, so the planes of the whole shard never change
  const bool const_w = p->loss.kind == SCS_LOSS_LEASTSQUARES && p->win_lo == 0 && p->win_hi == p->n;
  const int64_t nproc = p->ahi - p->alo;
  if (nproc <= 0) return SCS_OK;  // empty window: the DMMA path writes the zero Gram
  I8Plan pl = p->i8plan;  // restricted to the active row window
  pl.kb_lo = p->alo / kI8BK;
  pl.kblocks = pl.kb_lo + (nproc + kI8BK - 1) / kI8BK;
  pl.nchunks = (int)((pl.kblocks - pl.kb_lo + pl.chunk_kblocks - 1) / pl.chunk_kblocks);
  pl.units = (int64_t)pl.nmod * pl.nchunks * pl.ntiles;
  const bool nonneg = weights_nonneg_apriori(p, p->fwd_wk);
  if (!(const_w && p->i8_planes_valid)) {
    StageTimer t(c, ST_RESID);
    const int nblk = (int)((nproc + kWsRows - 1) / kWsRows);
    LAUNCH(c, k_wstat_part, nblk, 256, 0, (const double*)(p->dw + p->alo), nproc, p->d_wpart);
    LAUNCH(c, k_wstat_fin, 1, kVecThreads, 0, (const double*)p->d_wpart, nblk, p->d_wstat, p->d_negbase);
    p->i8_signed = false;
    p->i8_minor_neg = true;
    p->i8_ccount = 0;
    if (!nonneg) {  // the sign pattern decides the launch plan: one small read-back (the a-priori case has none)
      double st[WS_COUNT];
      CU_TRY(cudaMemcpyAsync(st, p->d_wstat, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
      CU_TRY(cudaStreamSynchronize(c->stream));
      // NaN / Inf weights: one GPU hands the call to the fp64 kernel, which propagates them; with several ranks every
      // rank must stay on the same exchange path, so the int8 path goes on and k_crt poisons the result (WS_BAD)
      const bool bad = st[WS_BAD] != 0.0;
      if (bad && c->world == 1) return SCS_OK;
      const int64_t nneg = bad ? 0 : (int64_t)(st[WS_NNEG] + 0.5);
      int rc = nneg > 0 ? i8_setup_signed(p) : SCS_OK;
      if (rc != SCS_OK && rc != SCS_OOM) return rc;
      bool all_ok = rc == SCS_OK;
      if (c->world > 1) SCS_TRY(agree_ok(c, rc == SCS_OK, &all_ok));  // every rank takes the same Gram path
      if (!all_ok) return SCS_OK;  // no room for the compacted planes somewhere: all ranks fall back to DMMA together
      if (nneg > 0) {
        p->i8_signed = true;
        p->i8_minor_neg = nneg <= nproc - nneg;
        p->i8_ccount = p->i8_minor_neg ? nneg : nproc - nneg;
      }
    }
    LAUNCH(c, k_colscale, (m + 255) / 256, 256, 0, (const double*)p->d_colnorm2, (const double*)p->d_wstat, m, p->i8_T,
           p->d_colinv, p->d_colscale);
    const unsigned gx = (unsigned)nblk;
    const dim3 rgrid(gx, (unsigned)std::min(m, 64));
    switch (p->i8_nmod) {
#define SCS_RES_CASE(K)                                                                                                \
  case K:                                                                                                              \
    if (p->i8_signed)                                                                                                  \
      LAUNCH(c, (k_residues<K, true>), rgrid, 256, 0, p->dA + p->alo, p->ldd, nproc, m, p->dw + p->alo, p->d_colscale, \
             p->d_planes + p->alo, p->ldx, (const int64_t*)p->d_negbase, p->i8_minor_neg ? 1 : 0, p->d_cplanes,       \
             p->ldc);                                                                                                  \
    else                                                                                                               \
      LAUNCH(c, (k_residues<K, false>), rgrid, 256, 0, p->dA + p->alo, p->ldd, nproc, m, p->dw + p->alo,              \
             p->d_colscale, p->d_planes + p->alo, p->ldx, (const int64_t*)nullptr, 0, (int8_t*)nullptr, (int64_t)0);  \
    break;
      SCS_RES_CASE(10)
      SCS_RES_CASE(11)
      SCS_RES_CASE(12)
      SCS_RES_CASE(13)
      SCS_RES_CASE(14)
      SCS_RES_CASE(15)
#undef SCS_RES_CASE
      default:
        return fail(SCS_STATE_ERROR, "int8 Gram: bad moduli count");
    }
    if (p->i8_signed && p->i8_ccount % kI8BK)
      LAUNCH(c, k_cpad, dim3(m, p->i8_nmod), 128, 0, p->d_cplanes, p->ldc, m, p->i8_ccount);
    p->i8_planes_valid = true;
  }
  // chunk slots of the partial-residue buffer: the main SYRK first, then the compacted rows
  I8Plan cpl = pl;
  cpl.kb_lo = 0;
  cpl.kblocks = (p->i8_ccount + kI8BK - 1) / kI8BK;
  cpl.nchunks = p->i8_signed ? (int)((cpl.kblocks + cpl.chunk_kblocks - 1) / cpl.chunk_kblocks) : 0;
  cpl.units = (int64_t)cpl.nmod * cpl.nchunks * cpl.ntiles;
  pl.pchunks = cpl.pchunks = pl.nchunks + cpl.nchunks;
  pl.pchunk0 = 0;
  cpl.pchunk0 = pl.nchunks;
  pl.main_chunks = cpl.main_chunks = pl.nchunks;
  pl.sign_main = cpl.sign_main = (p->i8_signed && !p->i8_minor_neg) ? -1 : 1;
  pl.sign_extra = cpl.sign_extra = p->i8_minor_neg ? -2 : 2;
  if (pl.pchunks > p->i8_pchunks_cap) return fail(SCS_STATE_ERROR, "int8 Gram: partial-residue buffer too small");
  {
    StageTimer t(c, ST_GRAM);
    // SCS_I8_OVERLAP_PROBE=1 (measurement only): a second k_residues launch with the same arguments (it rewrites the
    // planes with identical bytes) runs on the second stream WHILE the SYRK runs: the `gram` stage then reports the time
    // of the two together, i.e. what a chunk-pipelined residue pass could hide at best.
    static const bool overlap_probe = getenv("SCS_I8_OVERLAP_PROBE") && atoi(getenv("SCS_I8_OVERLAP_PROBE"));
    const bool probe = overlap_probe && !p->i8_signed && p->i8_nmod == 12;
    if (probe) {
      CU_TRY(cudaEventRecord(c->ev_trsm[0], c->stream));
      CU_TRY(cudaStreamWaitEvent(c->stream2, c->ev_trsm[0], 0));
      const int nblk = (int)((nproc + kWsRows - 1) / kWsRows);
      k_residues<12, false><<<dim3((unsigned)nblk, (unsigned)std::min(m, 64)), 256, 0, c->stream2>>>(
          p->dA + p->alo, p->ldd, nproc, m, p->dw + p->alo, p->d_colscale, p->d_planes + p->alo, p->ldx,
          (const int64_t*)nullptr, 0, (int8_t*)nullptr, (int64_t)0);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaEventRecord(c->ev_upd[0], c->stream2));
    }
    SCS_TRY(i8_launch_syrk(p, p->xmap, p->xmap_b, pl));
    if (cpl.nchunks > 0) SCS_TRY(i8_launch_syrk(p, p->cmap, p->cmap_b, cpl));
    if (probe) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_upd[0], 0));
  }
  if (c->world <= 1) {
    StageTimer t(c, ST_GRAMFIN);
    LAUNCH(c, k_crt, dim3((m + 255) / 256, (m + 3) / 4), 256, 0, p->d_i8partial, pl, (const double*)p->d_colinv,
           (const double*)p->d_wstat, nonneg ? 1 : 0, p->d_G, 0, m, (double*)nullptr);
    *done = 1;
    return SCS_OK;
  }
  // Several ranks: the exchange step.  Only the packed upper triangle travels (half the bytes of the square), and it
  // travels slab by slab: the CRT of slab s+1 runs on the main stream while NCCL reduces slab s on the second stream.
  // Slab boundaries split the packed triangle into equal byte counts (rows ~ sqrt).
  const size_t npack = (size_t)m * (m + 1) / 2;
  SCS_TRY(p2p_setup(p, npack));
  if (p->p2p_ok) {
    // the exchange step as our own kernels over NVLink peer memory (kernels_i8gram.cuh: k_p2p_*)
    {
      StageTimer t(c, ST_GRAMFIN);
      LAUNCH(c, k_crt, dim3((m + 255) / 256, (m + 3) / 4), 256, 0, p->d_i8partial, pl, (const double*)p->d_colinv,
             (const double*)p->d_wstat, nonneg ? 1 : 0, p->d_G, 0, m, p->d_Gpack);
    }
    {
      StageTimer t(c, ST_COMM);
      const int W = c->world;
      const unsigned gx = (unsigned)std::min<size_t>(4 * (size_t)c->num_sms, (npack / W / 2 + 255) / 256 + 1);
      LAUNCH(c, k_p2p_barrier, 1, 32, 0, p->p2p, c->rank, W, ++p->p2p_epoch);
      LAUNCH(c, k_p2p_reduce_scatter, gx, 256, 0, p->p2p, c->rank, W, npack);
      LAUNCH(c, k_p2p_barrier, 1, 32, 0, p->p2p, c->rank, W, ++p->p2p_epoch);
      LAUNCH(c, k_p2p_allgather, dim3(std::max(1u, gx / (unsigned)(W - 1)), (unsigned)(W - 1)), 256, 0, p->p2p, c->rank, W,
             npack);
      LAUNCH(c, k_p2p_barrier, 1, 32, 0, p->p2p, c->rank, W, ++p->p2p_epoch);
    }
    {
      StageTimer t(c, ST_GRAMFIN);
      const int t32 = (m + 31) / 32;
      LAUNCH(c, k_unpack_upper, t32 * (t32 + 1) / 2, 256, 0, (const double*)p->d_Gpack, m, p->d_G);
    }
    *done = 2;
    return SCS_OK;
  }
  if (!p->d_Gpack) SCS_TRY(dalloc(&p->d_Gpack, npack));
  static const int slab_env = getenv("SCS_GRAM_SLABS") ? atoi(getenv("SCS_GRAM_SLABS")) : 0;  // tuning switch (1..4)
  const int nslab = m >= 2048 ? (slab_env >= 1 && slab_env <= 4 ? slab_env : 4) : 1;
  int j0 = 0;
  for (int sidx = 0; sidx < nslab; ++sidx) {
    int j1 = sidx == nslab - 1 ? m : (int)std::floor(m * std::sqrt((sidx + 1.0) / nslab));
    j1 = std::min(m, std::max(j0 + 4, j1 / 4 * 4));
    if (sidx == nslab - 1) j1 = m;
    {
      StageTimer t(c, ST_GRAMFIN);
      LAUNCH(c, k_crt, dim3((j1 + 255) / 256, (j1 - j0 + 3) / 4), 256, 0, p->d_i8partial, pl, (const double*)p->d_colinv,
             (const double*)p->d_wstat, nonneg ? 1 : 0, p->d_G, j0, j1, p->d_Gpack);
    }
    CU_TRY(cudaEventRecord(c->ev_slab[sidx], c->stream));
    CU_TRY(cudaStreamWaitEvent(c->stream2, c->ev_slab[sidx], 0));
    const size_t off0 = (size_t)j0 * (j0 + 1) / 2, off1 = (size_t)j1 * (j1 + 1) / 2;
    SCS_TRY(allreduce(c, p->d_Gpack + off0, off1 - off0, c->stream2));
    j0 = j1;
    if (j0 >= m) break;
  }
  CU_TRY(cudaEventRecord(c->ev_slab[4], c->stream2));
  CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_slab[4], 0));
  {
    StageTimer t(c, ST_GRAMFIN);
    const int t32 = (m + 31) / 32;
    LAUNCH(c, k_unpack_upper, t32 * (t32 + 1) / 2, 256, 0, (const double*)p->d_Gpack, m, p->d_G);
  }
  *done = 2;  // all-reduced already
  return SCS_OK;
}

// G = A' diag(w) A (all-reduced, both triangles) into d_G, using the weights currently in d_w
static int run_gram(scs_problem* p, XRef x) {
  scs_ctx* c = p->ctx;
  SCS_TRY(gram_setup(p));
  const int m = (int)p->m;
  if (p->loss.kind == SCS_LOSS_QUADFORM) {
    StageTimer t(c, ST_GRAM);
    LAUNCH(c, k_quadform_hess, dim3((m + 255) / 256, m), 256, 0, p->dA, p->ldd, m, p->d_G);
    return SCS_OK;
  }
  if (p->sparse) {
    p->last_gram_path = 3;
    {
      StageTimer t(c, ST_GRAM);
      const size_t smem = (size_t)m * sizeof(long long);
      size_t& attr_smem = c->sp_gram_smem;  // function attributes are per device: cached in the context
      if (smem > attr_smem) {
        if (smem > 220 * 1024) return fail(SCS_UNSUPPORTED, "sparse Gram: m > 28160 is not supported");
        CU_TRY(cudaFuncSetAttribute(k_sp_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
      }
      if (!p->d_spscale) SCS_TRY(dalloc(&p->d_spscale, m));
      LAUNCH(c, k_sp_colscale, (unsigned)((m + 7) / 8), 256, 0, (const int64_t*)p->d_colptr, (const int*)p->d_rowidx,
             (const double*)p->d_cvals, (const double*)p->dw, m, p->d_spscale);
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / std::max<size_t>(smem + 64, 1)));
      const int grid = std::min(m, c->num_sms * per_sm);
      LAUNCH(c, k_sp_gram, grid, 256, smem, (const int64_t*)p->d_colptr, (const int*)p->d_rowidx, (const double*)p->d_cvals,
             (const int64_t*)p->d_rowptr, (const int*)p->d_colidx, (const double*)p->d_vals, (const double*)p->dw,
             (const double*)p->d_spscale, m, p->d_G);
    }
    {
      StageTimer t(c, ST_GRAMFIN);  // k_sp_gram wrote the lower triangle of every column: mirror it
      LAUNCH(c, k_symmetrize, dim3((m + 255) / 256, m), 256, 0, p->d_G, (int64_t)m, m);
    }
    SCS_TRY(allreduce(c, p->d_G, (size_t)m * m));
    return SCS_OK;
  }
  bool want_i8 = p->gram_mode == 2 || (p->gram_mode == 0 && p->m >= 512 && p->win_hi - p->win_lo >= 32768);
  if (want_i8 && !p->i8_failed) {
    int done = 0;
    int rc = run_gram_i8(p, &done);
    if (rc != SCS_OK && !(rc == SCS_OOM && p->gram_mode == 0)) return rc;
    if (rc == SCS_OOM)  // auto mode: say so once — the DMMA kernel is ~4x slower, a silent downgrade would look like a regression
      fprintf(stderr, "[scs_b200] rank %d: %s; the Gram falls back to the native fp64 (DMMA) kernel for this problem\n",
              c->rank, g_err.c_str());
    if (done) {
      p->last_gram_path = 2;
      if (done == 1) SCS_TRY(allreduce(c, p->d_G, (size_t)m * m));
      return SCS_OK;
    }
  }
  p->last_gram_path = 1;
  {
    StageTimer t(c, ST_GRAM);
    GramPlan pl = p->plan;
    pl.kt = (p->ahi - p->alo) / kGBK;  // k-tiles of the active row window (the K split was planned for the whole shard)
    const int grid = (int)std::min<int64_t>(c->num_sms, pl.units);
    LAUNCH(c, k_gram, grid, kGThreads, kGSmemBytes, p->amap, p->dw, m, p->ldp, pl, p->alo / kGBK, p->d_partial);
  }
  {
    StageTimer t(c, ST_GRAMFIN);
    const int t32 = (m + 31) / 32;
    LAUNCH(c, k_gram_finalize, t32 * (t32 + 1) / 2, 256, 0, p->d_partial, p->plan.splits, m, p->ldp, p->d_G);
  }
  SCS_TRY(allreduce(c, p->d_G, (size_t)m * m));
  return SCS_OK;
}

// Above this order the look-ahead Cholesky applies the trailing update two panels at a time (chol_enqueue): the
// m x m matrix (8 m^2 bytes) no longer fits the 126 MB L2 and each pass over the trailing part is DRAM traffic.
constexpr int kPairMinM = 4096;

// Enqueues the factorisation part of run_solve (save, look-ahead or legacy Cholesky with the forward solve folded in) on
// c->stream / c->stream2.  Capturable: no host synchronisation, the second stream forks from and re-joins the first.
static int chol_enqueue(scs_ctx* c, double* M, double* Msave, double* Linv, int* d_info, double* b, double* tmp,
                        double* dsol, int m, bool sym, long long* d_prof) {
  if (sym)
    LAUNCH(c, k_save_diag, (m + 255) / 256, 256, 0, (const double*)M, (int64_t)m, m, Msave);
  else
    CU_TRY(cudaMemcpyAsync(Msave, M, (size_t)m * m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  // the LU fallback needs the untouched right-hand side: keep it in dsol until the factorisation has succeeded
  CU_TRY(cudaMemcpyAsync(dsol, b, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CU_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), c->stream));
  const int nblk = (m + kNB - 1) / kNB;
  double* rdiag = Linv;  // first m doubles of the workspace hold 1/L_jj
  if (c->solve_mode == 0) {
    // look-ahead sequence (kernels_solve.cuh): diag -> trsm on the main stream, the trailing update of step k on the
    // second stream while diag(k+1) runs; trsm(k+1) waits for it.
    // Paired variant (large m, where the trailing matrix no longer fits the L2 and every pass over it is DRAM
    // traffic): after an even step only the next block column receives panel k (few tiles: trsm(k+1) needs them);
    // after the odd step k+1 the whole trailing matrix receives panels k and k+1 in ONE read-modify-write, and the
    // diagonal block at the pair boundary applies both pending panels to its own tile.
    const bool pair = c->solve_pair < 0 ? m > kPairMinM : c->solve_pair != 0;
    bool upd_pending = false;
    int upd_idx = 0;
    for (int k = 0; k < nblk; ++k) {
      const int k0 = k * kNB;
      const int nb = std::min(kNB, m - k0);
      const int rem = m - k0 - nb;
      const int rb = (rem + kNB - 1) / kNB;
      const int nprev = k == 0 ? 0 : ((pair && !(k & 1)) ? 2 : 1);
      LAUNCH(c, k_chol_diag, 1, kCholDiagThreads, kCholDiagSmem, M, (int64_t)m, m, k0, rdiag, d_info, (const double*)b,
             tmp, (k == 1 ? d_prof : (long long*)nullptr), nprev);
      if (rem <= 0) break;
      if (upd_pending) {
        CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_upd[upd_idx], 0));
        upd_pending = false;
      }
      LAUNCH(c, k_chol_trsm<false>, rb, kTrsmThreads, 0, M, (int64_t)m, m, k0, (const double*)rdiag, b,
             (const double*)tmp, (double*)nullptr, (k == 1 && d_prof ? d_prof + 11 : (long long*)nullptr));
      const bool colonly = pair && !(k & 1);
      // the first diagonal tile is k_chol_diag(k+1)'s
      const int ntiles = colonly ? rb - 1 : rb * (rb + 1) / 2 - 1;
      if (ntiles > 0) {
        CU_TRY(cudaEventRecord(c->ev_trsm[k & 1], c->stream));
        CU_TRY(cudaStreamWaitEvent(c->stream2, c->ev_trsm[k & 1], 0));
        k_syrk_update<<<ntiles, 128, kTileSmem, c->stream2>>>(M, (int64_t)m, m, k0, ntiles, b, (const double*)tmp, 1,
                                                              (pair && (k & 1)) ? 2 : 1, colonly ? 1 : 0);
        c->launches += 1;
        CU_TRY(cudaGetLastError());
        upd_idx = k & 1;
        CU_TRY(cudaEventRecord(c->ev_upd[upd_idx], c->stream2));
        upd_pending = true;
      }
    }
    if (upd_pending) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_upd[upd_idx], 0));  // join (a later trsm normally has)
  } else {
    // b rides along as an extra row of the factorisation: tmp receives y = L^-1 b block by block
    for (int k = 0; k < nblk; ++k) {
      const int k0 = k * kNB;
      const int nb = std::min(kNB, m - k0);
      const int rem = m - k0 - nb;
      const int rb = (rem + kNB - 1) / kNB;
      LAUNCH(c, k_panel, 2 + rb, 256, 0, M, (int64_t)m, m, k0, rdiag, d_info, b, tmp, rb);
      if (rem > 0) {
        const int ntiles = rb * (rb + 1) / 2;
        LAUNCH(c, k_syrk_update, ntiles + (rem + 127) / 128, 128, kTileSmem, M, (int64_t)m, m, k0, ntiles, b,
               (const double*)tmp, 0, 1, 0);
      }
    }
  }
  return SCS_OK;
}

// Solve M d = b on the device.  M (ld = m) is destroyed.  For the LU fallback the original must survive: sym = true
// (M holds both triangles of a symmetric matrix): only the diagonal is saved (Msave: m doubles), the Cholesky never
// writes the upper triangle; sym = false (only the lower triangle is trusted): Msave (m x m) receives a copy first.
// b is destroyed; result in dsol.  tmp: m doubles.
// (Replaying the ~190 launches + event operations of the look-ahead sequence as one captured CUDA graph was measured
// slower than issuing them: 3.45 vs 3.12 ms at m = 4096, 17.2 vs 15.4 ms at m = 8192 — the cross-branch graph edges cost
// more than the stream-event waits; the sequence is therefore launched directly.)
static int run_solve(scs_ctx* c, double* M, double* Msave, double* Linv, int* d_info, double* b, double* tmp,
                     double* dsol, int m, int* used_fallback, bool sym) {
  StageTimer t(c, ST_SOLVE);
  const int nblk = (m + kNB - 1) / kNB;
  {
    bool& attr_set = c->solve_attr_set;  // function attributes are per device: cached in the context
    if (!attr_set) {
      CU_TRY(cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
      // same shared-memory carve-out for every kernel of the sequence: no SM reconfiguration between launches
      CU_TRY(cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CU_TRY(cudaFuncSetAttribute(k_panel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CU_TRY(cudaFuncSetAttribute(k_bwd_all, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CU_TRY(cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholDiagSmem));
      CU_TRY(cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CU_TRY(cudaFuncSetAttribute(k_chol_trsm<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CU_TRY(cudaFuncSetAttribute(k_chol_trsm<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      attr_set = true;
    }
  }
  double* rdiag = Linv;  // first m doubles of the workspace hold 1/L_jj
  double* Wt = Linv + round_up(m, 16);  // (L_kk^-1)' per diagonal block
  long long* d_prof = nullptr;  // SCS_CHOL_PROF=1: clock64 stamps of the second k_chol_diag launch go to stderr
  static const bool chol_prof = getenv("SCS_CHOL_PROF") && atoi(getenv("SCS_CHOL_PROF"));
  if (chol_prof && c->solve_mode == 0) {
    CU_TRY(cudaMalloc((void**)&d_prof, 16 * sizeof(long long)));
    CU_TRY(cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), c->stream));
  }
  SCS_TRY(chol_enqueue(c, M, Msave, Linv, d_info, b, tmp, dsol, m, sym, d_prof));
  int info = 0;
  CU_TRY(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  if (d_prof) {
    long long st[16];
    CU_TRY(cudaMemcpy(st, d_prof, sizeof(st), cudaMemcpyDeviceToHost));
    fprintf(stderr, "k_chol_diag stamps (cycles from entry):");
    for (int i = 1; i <= 10; ++i) fprintf(stderr, " %lld", st[i] ? st[i] - st[0] : -1LL);
    fprintf(stderr, "; k_chol_trsm: %lld %lld %lld\n", st[12] - st[11], st[13] - st[11], st[14] - st[11]);
    cudaFree(d_prof);
  }
  if (info == 0) {
    // tmp now holds y
    // backward substitution: invert the diagonal blocks (all at once), then one persistent kernel for the sweep
    unsigned long long* bar = (unsigned long long*)(Wt + (size_t)nblk * kNB * kNB);
    if (c->solve_mode == 0)
      LAUNCH(c, k_chol_trsm<true>, nblk, kTrsmThreads, 0, M, (int64_t)m, m, 0, (const double*)rdiag, (double*)nullptr,
             (const double*)nullptr, Wt, (long long*)nullptr);
    else
      LAUNCH(c, k_invdiag, nblk, 64, 0, (const double*)M, (int64_t)m, m, Wt);
    int* flags = (int*)(bar + 2);
    const double* Mc = M;
    int64_t ldm = m;
    int mm = m;
    const double* Wc = Wt;
    int& p2p_capacity = c->p2p_capacity;  // co-resident CTAs of k_bwd_p2p on this device
    if (p2p_capacity < 0) {
      int per_sm = 0;
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bwd_p2p, 256, 0));
      p2p_capacity = per_sm * c->num_sms;
    }
    if (c->solve_mode == 0 && nblk <= p2p_capacity) {
      CU_TRY(cudaMemsetAsync(flags, 0, (size_t)nblk * sizeof(int), c->stream));
      const double* yc = tmp;
      void* args[] = {(void*)&Mc, (void*)&ldm, (void*)&mm, (void*)&Wc, (void*)&yc, (void*)&dsol, (void*)&flags};
      cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_bwd_p2p, dim3(nblk), dim3(256), args, 0, c->stream);
      c->launches += 1;
      if (le != cudaSuccess) return fail(SCS_CUDA_ERROR, std::string("k_bwd_p2p launch: ") + cudaGetErrorString(le));
    } else {
      CU_TRY(cudaMemsetAsync(bar, 0, sizeof(unsigned long long), c->stream));
      int grid = std::min(c->num_sms, nblk);
      void* args[] = {(void*)&Mc, (void*)&ldm, (void*)&mm, (void*)&Wc, (void*)&tmp, (void*)&dsol, (void*)&bar};
      cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_bwd_all, dim3(grid), dim3(256), args, 0, c->stream);
      c->launches += 1;
      if (le != cudaSuccess) return fail(SCS_CUDA_ERROR, std::string("k_bwd_all launch: ") + cudaGetErrorString(le));
    }
    *used_fallback = 0;
    return SCS_OK;
  }
  // not positive definite: partial-pivoting LU on the saved copy (symmetrised), still on the device
  *used_fallback = 1;
  CU_TRY(cudaMemcpyAsync(b, dsol, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (sym) {
    LAUNCH(c, k_restore_lower, dim3((m + 255) / 256, m), 256, 0, M, (int64_t)m, m, (const double*)Msave);
    Msave = M;  // factor in place
  } else {
    LAUNCH(c, k_symmetrize, dim3((m + 255) / 256, m), 256, 0, Msave, (int64_t)m, m);
  }
  // restore b (the Cholesky path has not touched b yet, but keep the contract explicit)
  CU_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), c->stream));
  for (int k = 0; k < m; ++k) {
    LAUNCH(c, k_lu_pivot, 1, kVecThreads, 0, Msave, (int64_t)m, m, k, b, d_info);
    const int rem = m - k - 1;
    if (rem > 0) LAUNCH(c, k_lu_update, dim3((rem + 255) / 256, (rem + 15) / 16), 256, 0, Msave, (int64_t)m, m, k, b);
  }
  LAUNCH(c, k_lu_backsolve, 1, kVecThreads, 0, Msave, (int64_t)m, m, b, dsol);
  return SCS_OK;
}

// ProxGGNSCORE underdetermined branch (rows + 1 <= m, prox-GGN-SCORE.jl:124-127) on the rows of the active window;
// leaves gr, Hr, the damping scalars (k_pre) and d (before negation) in d_gr, d_hr, d_scal, d_sol.  See kernels_solve.cuh.
// The n x n system couples every pair of rows of the batch.  One rank with a dense shard works on its rows in place;
// several ranks (or a sparse shard) first REPLICATE the batch — it is small by definition, rows < m — into a dense buffer
// (k_wide_gather + one sum all-reduce) and then every rank runs the same kernels on the same bits: the direction is
// bitwise identical everywhere, like the rest of the replicated x-side state.
static int run_ggn_wide(scs_problem* p, XRef x, double lam) {
  scs_ctx* c = p->ctx;
  const int m = (int)p->m;
  const int nbl = (int)(p->win_hi - p->win_lo);  // this rank's rows of the batch (may be 0)
  const int nb = (int)p->win_rows_global;        // the whole batch
  if (nb < 1) return fail(SCS_INVALID_ARG, "empty batch");
  const bool repl = c->world > 1 || p->sparse;
  SCS_TRY(gram_setup(p));
  if (!p->d_u) SCS_TRY(dalloc(&p->d_u, p->ldd));
  {
    StageTimer t(c, ST_VEC);
    LAUNCH(c, k_pre, 1, kVecThreads, 0, p->sm, lam, x.d, (const double*)nullptr, m, p->d_gr, p->d_hr, p->d_rhs, p->d_scal);
    LAUNCH(c, k_wide_v, (m + 255) / 256, 256, 0, p->d_gr, p->d_hr, m, p->d_t1);
  }
  // pieces s, res, q of the (out_fn, f(y,ŷ)) pair for the window's rows -> dz, dr, dw;  u = A (gr ./ Hr) -> d_u
  SCS_TRY(run_forward(p, x.d, 2));
  p->fwd_id = 0;  // not a pass the caches know about (no loss sum, z replaced by s)
  p->grad_id = 0;
  p->loss_reduced = false;
  SCS_TRY(run_matvec(p, p->d_t1, p->d_u));
  const double *Ab = p->dA ? p->dA + p->win_lo : nullptr, *vs = p->dz + p->win_lo, *vres = p->dr + p->win_lo,
               *vq = p->dw + p->win_lo;
  double* vu = p->d_u + p->win_lo;
  int64_t ldb = p->ldd;
  if (repl) {
    // where this rank's rows sit inside the batch: exclusive prefix of the per-rank row counts
    int64_t off = 0;
    if (c->world > 1) {
      std::vector<double> cnt(c->world, 0.0);
      cnt[c->rank] = (double)nbl;
      if (!p->d_wcnt) SCS_TRY(dalloc(&p->d_wcnt, c->world));
      CU_TRY(cudaMemcpyAsync(p->d_wcnt, cnt.data(), c->world * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      SCS_TRY(allreduce(c, p->d_wcnt, c->world));
      CU_TRY(cudaMemcpyAsync(cnt.data(), p->d_wcnt, c->world * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CU_TRY(cudaStreamSynchronize(c->stream));
      int64_t tot = 0;
      for (int r = 0; r < c->world; ++r) {
        if (r < c->rank) off += (int64_t)(cnt[r] + 0.5);
        tot += (int64_t)(cnt[r] + 0.5);
      }
      if (tot != nb) return fail(SCS_STATE_ERROR, "wide GGN branch: the ranks disagree on the batch size");
    }
    const int64_t ldw = round_up(nb, 16);
    const size_t need = (size_t)ldw * (m + 4);
    if (p->wide_cap < need) {
      dfree(p->d_wide);
      p->d_wide = nullptr;
      SCS_TRY(dalloc(&p->d_wide, need));
      p->wide_cap = need;
    }
    const int64_t blocks = (ldw + 64 * 8 - 1) / (64 * 8);
    if (p->widepart_cap < (size_t)blocks * m) {
      dfree(p->d_widepart);
      p->d_widepart = nullptr;
      SCS_TRY(dalloc(&p->d_widepart, (size_t)blocks * m));
      p->widepart_cap = (size_t)blocks * m;
    }
    StageTimer t(c, ST_FWD);
    CU_TRY(cudaMemsetAsync(p->d_wide, 0, need * sizeof(double), c->stream));
    if (nbl > 0) {
      LAUNCH(c, k_wide_gather, dim3((nbl + 255) / 256, m + 4), 256, 0, (const double*)p->dA, p->ldd, p->win_lo, nbl, m,
             (const double*)p->dz, (const double*)p->dr, (const double*)p->dw, (const double*)p->d_u, p->d_wide, ldw, off);
      if (p->sparse)
        LAUNCH(c, k_wide_gather_csr, (nbl + 255) / 256, 256, 0, (const int64_t*)p->d_rowptr, (const int*)p->d_colidx,
               (const double*)p->d_vals, p->win_lo, nbl, p->d_wide, ldw, off);
    }
    SCS_TRY(allreduce(c, p->d_wide, need));
    Ab = p->d_wide;
    ldb = ldw;
    vs = p->d_wide + (size_t)m * ldw;
    vres = vs + ldw;
    vq = vres + ldw;
    vu = p->d_wide + (size_t)(m + 3) * ldw;
  }
  StageTimer t(c, ST_SOLVE);
  const int nt = (nb + 63) / 64;
  LAUNCH(c, k_rowgram, nt * (nt + 1) / 2, 256, 0, Ab, ldb, nb, m, p->d_hr, p->d_G, nb);
  LAUNCH(c, k_wide_system, dim3((nb + 255) / 256, nb), 256, 0, p->d_G, nb, nb, vs, vres, vq, (const double*)vu, lam, p->d_q);
  // general (non-symmetric) system: partial-pivoting LU on the device
  CU_TRY(cudaMemsetAsync(p->d_info, 0, sizeof(int), c->stream));
  for (int k = 0; k < nb; ++k) {
    LAUNCH(c, k_lu_pivot, 1, kVecThreads, 0, p->d_G, (int64_t)nb, nb, k, p->d_q, p->d_info);
    const int rem = nb - k - 1;
    if (rem > 0) LAUNCH(c, k_lu_update, dim3((rem + 255) / 256, (rem + 15) / 16), 256, 0, p->d_G, (int64_t)nb, nb, k, p->d_q);
  }
  LAUNCH(c, k_lu_backsolve, 1, kVecThreads, 0, p->d_G, (int64_t)nb, nb, p->d_q, p->d_t2);
  LAUNCH(c, k_wide_scale, (nb + 255) / 256, 256, 0, vs, p->d_t2, nb, vu);
  if (repl) {  // A'(s ∘ B') from the replicated batch: the same bits on every rank
    const int64_t blocks = (ldb + 64 * 8 - 1) / (64 * 8);
    LAUNCH(c, k_adjoint<8>, (unsigned)blocks, kAdjThreads, 0, Ab, ldb, ldb, m, (const double*)vu, p->d_widepart);
    LAUNCH(c, k_colsum, (unsigned)((m + 31) / 32), 256, 0, p->d_widepart, blocks, m, p->d_t1);
  } else {
    SCS_TRY(run_adjoint(p, p->d_u, p->d_t1));
  }
  LAUNCH(c, k_wide_dir, (m + 255) / 256, 256, 0, p->d_t1, p->d_gr, p->d_hr, lam, m, p->d_sol);
  p->last_used_fallback = 1;
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// exported: misc
// ------------------------------------------------------------------------------------------------
extern "C" int scs_version(void) { return SCS_B200_VERSION; }
extern "C" const char* scs_last_error(void) { return g_err.c_str(); }

extern "C" int scs_comm_unique_id(void* id128) {
  if (!id128) return fail(SCS_INVALID_ARG, "id128 is NULL");
  const char* why = "";
  if (!g_nccl.load(&why)) return fail(SCS_NCCL_ERROR, why);
  NcclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(SCS_NCCL_ERROR, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  memcpy(id128, &id, 128);
  return SCS_OK;
}

extern "C" int scs_ctx_create(int device, int rank, int world, const void* id128, scs_ctx** out) {
  if (!out) return fail(SCS_INVALID_ARG, "out is NULL");
  if (world < 1 || rank < 0 || rank >= world) return fail(SCS_INVALID_ARG, "bad rank/world");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SCS_CUDA_ERROR, std::string("no CUDA device available (this library has no CPU path): ") +
                                    cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(SCS_INVALID_ARG, "device index out of range");
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SCS_UNSUPPORTED, "scs_b200 kernels are built for sm_100a only; found compute capability " +
                                     std::to_string(prop.major) + "." + std::to_string(prop.minor));
  scs_ctx* c = new scs_ctx();
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->num_sms = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU_TRY(cudaEventCreateWithFlags(&c->ev_trsm[i], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_upd[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 5; ++i) CU_TRY(cudaEventCreateWithFlags(&c->ev_slab[i], cudaEventDisableTiming));
  if (const char* e = getenv("SCS_SOLVE_LEGACY")) c->solve_mode = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SCS_SOLVE_PAIR")) c->solve_pair = atoi(e) ? 1 : 0;
  CU_TRY(cudaMalloc((void**)&c->d_flag, sizeof(double)));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    c->encode = (EncodeTiledFn)fn;
  if (world > 1) {
    if (!id128) {
      delete c;
      return fail(SCS_INVALID_ARG, "world > 1 needs a unique id");
    }
    const char* why = "";
    if (!g_nccl.load(&why)) {
      delete c;
      return fail(SCS_NCCL_ERROR, why);
    }
    NcclUniqueId id;
    memcpy(&id, id128, 128);
    int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
      delete c;
      return fail(SCS_NCCL_ERROR, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
  }
  *out = c;
  return SCS_OK;
}

extern "C" int scs_ctx_destroy(scs_ctx* c) {
  if (!c) return SCS_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  harvest_timers(c);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->d_flag) cudaFree(c->d_flag);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_trsm[i]) cudaEventDestroy(c->ev_trsm[i]);
    if (c->ev_upd[i]) cudaEventDestroy(c->ev_upd[i]);
  }
  for (int i = 0; i < 5; ++i)
    if (c->ev_slab[i]) cudaEventDestroy(c->ev_slab[i]);
  delete c;
  return SCS_OK;
}
extern "C" int scs_ctx_sync(scs_ctx* c) {
  if (!c) return fail(SCS_INVALID_ARG, "ctx is NULL");
  return ctx_sync(c);
}
extern "C" int scs_ctx_stream(scs_ctx* c, uint64_t* s) {
  if (!c || !s) return fail(SCS_INVALID_ARG, "NULL argument");
  *s = (uint64_t)(uintptr_t)c->stream;
  return SCS_OK;
}
extern "C" int scs_get_counters(scs_ctx* c, int64_t* launches, int reset) {
  if (!c) return fail(SCS_INVALID_ARG, "ctx is NULL");
  if (launches) *launches = c->launches;
  if (reset) c->launches = 0;
  return SCS_OK;
}
extern "C" int scs_set_profiling(scs_ctx* c, int enable) {
  if (!c) return fail(SCS_INVALID_ARG, "ctx is NULL");
  c->profiling = enable != 0;
  return SCS_OK;
}
extern "C" int scs_get_stage_ms(scs_ctx* c, double* ms8, int64_t* calls8, int reset) {
  if (!c) return fail(SCS_INVALID_ARG, "ctx is NULL");
  SCS_TRY(ctx_sync(c));
  for (int i = 0; i < ST_N; ++i) {
    if (ms8) ms8[i] = c->stage_ms[i];
    if (calls8) calls8[i] = c->stage_calls[i];
    if (reset) {
      c->stage_ms[i] = 0;
      c->stage_calls[i] = 0;
    }
  }
  return SCS_OK;
}

// int8 tensor-pipe peak of this device, measured with the library's own UMMA (k_i8peak): burst = best of 5 launches of
// ~20 ms, sustained = back-to-back launches for `seconds`.  TOP/s = 2 * 128 * 256 * 32 ops per UMMA.
extern "C" int scs_measure_i8_peak(scs_ctx* c, double seconds, double* tops_burst, double* tops_sustained) {
  if (!c) return fail(SCS_INVALID_ARG, "ctx is NULL");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaFuncSetAttribute(k_i8peak, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8PeakSmem));
  uint32_t* sink = nullptr;
  CU_TRY(cudaMalloc((void**)&sink, sizeof(uint32_t)));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const double ops_per_group = 64.0 * (kI8BK / 32) * 2.0 * kI8BM * kI8BN * 32.0 * c->num_sms;
  const long long groups = 700;  // ~20 ms per launch at ~2.5 POP/s
  auto launch = [&]() {
    k_i8peak<<<c->num_sms, 128, kI8PeakSmem, c->stream>>>(groups, sink);
    c->launches += 1;
  };
  launch();  // warm-up
  CU_TRY(cudaStreamSynchronize(c->stream));
  CU_TRY(cudaGetLastError());
  double best = 0.0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0, c->stream);
    launch();
    cudaEventRecord(e1, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::max(best, ops_per_group * groups / (ms * 1e-3) / 1e12);
  }
  const int reps = std::max(3, (int)(seconds * best * 1e12 / (ops_per_group * groups)));
  cudaEventRecord(e0, c->stream);
  for (int r = 0; r < reps; ++r) launch();
  cudaEventRecord(e1, c->stream);
  CU_TRY(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (tops_burst) *tops_burst = best;
  if (tops_sustained) *tops_sustained = ops_per_group * groups * reps / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  CU_TRY(cudaGetLastError());
  return SCS_OK;
}

// Tuning aid: TOP/s of k_i8syrk's main loop (4-stage TMA -> UMMA ring, one CTA per SM) on an L2-resident operand.
// tma_mode 0: operands stay in shared memory (only the per-stage commit / slot hand-shake), 1: the A slab (16 KB per
// stage) through TMA, 2: A and B (48 KB per stage).  Best of 5 launches of ~15 ms.
extern "C" int scs_i8_pipe_probe(scs_ctx* c, int tma_mode, double* tops) {
  if (!c || !tops) return fail(SCS_INVALID_ARG, "NULL argument");
  if (tma_mode < 0 || tma_mode > 2) return fail(SCS_INVALID_ARG, "tma_mode must be 0, 1 or 2");
  if (!c->encode) return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  CU_TRY(cudaSetDevice(c->device));
  const int kspan = 512;  // 512 k-blocks x 128 bytes x 384 rows = 24 MB: stays in L2
  const size_t ldk = (size_t)kspan * kI8BK, bytes = ldk * 384;
  uint8_t* buf = nullptr;
  uint32_t* sink = nullptr;
  CU_TRY(cudaMalloc((void**)&buf, bytes));
  CU_TRY(cudaMalloc((void**)&sink, sizeof(uint32_t)));
  k_fill_random_bytes<<<1024, 256, 0, c->stream>>>(reinterpret_cast<uint32_t*>(buf), bytes / 4);  // realistic operand bits
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)ldk, 384u, 1u};
  cuuint64_t gstride[2] = {(cuuint64_t)ldk, (cuuint64_t)ldk * 384u};
  cuuint32_t box[3] = {(cuuint32_t)kI8BK, 128u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cudaFree(buf);
    cudaFree(sink);
    return fail(SCS_CUDA_ERROR, "cuTensorMapEncodeTiled (pipe probe) failed");
  }
  const int smem = kI8Stages * kI8StageBytes + 1024 + 256;
  CU_TRY(cudaFuncSetAttribute(k_i8pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const long long iters = 40000;  // x 4 UMMAs
  const double ops = (double)iters * (kI8BK / 32) * 2.0 * kI8BM * kI8BN * 32.0 * c->num_sms;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, c->stream);
    k_i8pipe<<<c->num_sms, 128, smem, c->stream>>>(map, tma_mode, iters, kspan, sink);
    c->launches += 1;
    cudaEventRecord(e1, c->stream);
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
  }
  *tops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// exported: problem lifetime
// ------------------------------------------------------------------------------------------------
static int problem_alloc(scs_ctx* ctx, int64_t n_local, int64_t m, int loss_kind, double loss_param, int label_mode,
                         scs_problem** out, bool dense = true) {
  if (!ctx || !out) return fail(SCS_INVALID_ARG, "NULL argument");
  if (n_local < 1 || m < 1) return fail(SCS_INVALID_ARG, "n_local and m must be positive");
  if (m > 2000000000LL) return fail(SCS_INVALID_ARG, "m too large");
  if (loss_kind < 0 || loss_kind > 2) return fail(SCS_INVALID_ARG, "unknown loss_kind");
  if (label_mode < 0 || label_mode > 1) return fail(SCS_INVALID_ARG, "unknown label_mode");
  if (loss_kind == SCS_LOSS_QUADFORM) {
    if (n_local != m) return fail(SCS_INVALID_ARG, "quadform loss needs a square A");
    if (ctx->world != 1) return fail(SCS_UNSUPPORTED, "quadform loss is single-GPU only");
  }
  CU_TRY(cudaSetDevice(ctx->device));
  scs_problem* p = new scs_problem();
  p->ctx = ctx;
  p->n = n_local;
  p->m = m;
  p->ldd = round_up(n_local, 16);
  p->mp = round_up(m, 16);
  p->loss.kind = loss_kind;
  p->loss.label_mode = label_mode;
  p->loss.weight_kind = 0;
  p->loss.p = loss_param;
  *out = p;
  if (dense) SCS_TRY(dalloc(&p->dA, (size_t)p->ldd * m));
  p->sparse = !dense;
  SCS_TRY(dalloc(&p->dy, p->ldd));
  SCS_TRY(dalloc(&p->dz, p->ldd));
  SCS_TRY(dalloc(&p->dr, p->ldd));
  SCS_TRY(dalloc(&p->dw, p->ldd));
  for (int i = 0; i < 3; ++i) SCS_TRY(dalloc(&p->vx[i], p->mp));
  SCS_TRY(dalloc(&p->d_gl, p->mp + 16));
  double** vecs[] = {&p->d_gr,    &p->d_hr,     &p->d_rhs,   &p->d_sol, &p->d_d,  &p->d_dx, &p->d_delta,
                     &p->d_gq,    &p->d_gqprev, &p->d_gamma, &p->d_q,   &p->d_t1, &p->d_t2, &p->d_xstar,
                     &p->d_trial, &p->d_gnewton};
  for (auto v : vecs) SCS_TRY(dalloc(v, p->mp));
  SCS_TRY(dalloc(&p->d_scal, SC_ALLOC));
  CU_TRY(cudaMallocHost((void**)&p->h_scal, (SC_COUNT + 8) * sizeof(double)));
  p->fwd_blocks = dense ? (p->ldd + kFwdRows - 1) / kFwdRows : (n_local + kSpFwdRows - 1) / kSpFwdRows;
  p->win_lo = 0;
  p->win_hi = n_local;
  p->alo = 0;
  p->ahi = p->ldd;
  p->win_rows_global = 0;  // resolved (all-reduced) by the first set_window
  p->adj_blocks = (p->ldd + 64 * 8 - 1) / (64 * 8);
  SCS_TRY(dalloc(&p->d_losspart, p->fwd_blocks));
  if (dense) SCS_TRY(dalloc(&p->d_adjpart, (size_t)p->adj_blocks * m));
  SCS_TRY(set_window(p, 0, n_local));  // whole shard; with several ranks this all-reduces the global row count
  return SCS_OK;
}

extern "C" int scs_problem_destroy(scs_problem* p) {
  if (!p) return SCS_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < kP2PMaxWorld; ++r)
    if (p->p2p_opened[r]) cudaIpcCloseMemHandle(p->p2p_opened[r]);
  void* bufs[] = {p->dA,      p->dy,      p->dz,     p->dr,      p->dw,       p->vx[0],   p->vx[1],    p->vx[2],
                  p->d_gl,    p->d_gr,    p->d_hr,   p->d_rhs,   p->d_sol,    p->d_d,     p->d_dx,     p->d_delta,
                  p->d_gq,    p->d_gqprev, p->d_gamma, p->d_q,   p->d_t1,     p->d_t2,    p->d_xstar,  p->d_trial,
                  p->d_gnewton, p->d_scal, p->d_losspart, p->d_adjpart, p->d_G, p->d_Gsave, p->d_partial, p->d_Linv,
                  p->d_info,  p->d_S,     p->d_Y,    p->d_state, p->d_rlb,    p->d_rub,   p->d_slb,    p->d_sub,
                  p->d_cdiag, p->d_ind,   p->d_perm,  p->d_planes, p->d_i8partial, p->d_colmax, p->d_wstat,
                  p->d_colscale, p->d_colinv, p->d_colnorm2, p->d_i8tiles, p->d_i8progress, p->d_cplanes, p->d_negbase, p->d_wpart, p->d_lspart, p->d_spscale, p->d_Gpack, p->d_wide, p->d_widepart, p->d_wcnt, p->d_fupart, p->d_fuloss, p->d_u, p->d_rowptr,
                  p->d_colptr, p->d_colidx, p->d_rowidx, p->d_vals, p->d_cvals};
  for (void* b : bufs) dfree(b);
  if (p->h_scal) cudaFreeHost(p->h_scal);
  delete p;
  return SCS_OK;
}

static bool labels_in_unit_range(const double* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    if (!(y[i] >= -1.0 && y[i] <= 1.0)) return false;
  return true;
}

extern "C" int scs_problem_create(scs_ctx* ctx, const double* A, int64_t n_local, int64_t m, int64_t lda,
                                  const double* y, int loss_kind, double loss_param, int label_mode,
                                  scs_problem** out) {
  if (!A || !y) return fail(SCS_INVALID_ARG, "A or y is NULL");
  if (lda < n_local) return fail(SCS_INVALID_ARG, "lda < n_local");
  int s = problem_alloc(ctx, n_local, m, loss_kind, loss_param, label_mode, out);
  if (s != SCS_OK) {
    if (out && *out) {
      scs_problem_destroy(*out);
      *out = nullptr;
    }
    return s;
  }
  scs_problem* p = *out;
  p->labels_unit = labels_in_unit_range(y, n_local);
  cudaError_t e = cudaMemcpy2DAsync(p->dA, p->ldd * sizeof(double), A, lda * sizeof(double),
                                    n_local * sizeof(double), m, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(p->dy, y, n_local * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    scs_problem_destroy(p);
    *out = nullptr;
    return fail(SCS_CUDA_ERROR, std::string("upload of A/y failed: ") + cudaGetErrorString(e));
  }
  return SCS_OK;
}

extern "C" int scs_problem_create_csc(scs_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                                      int64_t index_base, int64_t n_local, int64_t m, const double* y, int loss_kind,
                                      double loss_param, int label_mode, int storage, scs_problem** out) {
  if (!colptr || !y) return fail(SCS_INVALID_ARG, "colptr or y is NULL");
  if (index_base != 0 && index_base != 1) return fail(SCS_INVALID_ARG, "index_base must be 0 or 1");
  if (storage < 0 || storage > 2) return fail(SCS_INVALID_ARG, "storage must be 0 (auto), 1 (dense) or 2 (sparse)");
  if (m < 1 || n_local < 1) return fail(SCS_INVALID_ARG, "n_local and m must be positive");
  const int64_t nnz = colptr[m] - colptr[0];
  if (colptr[0] != index_base || nnz < 0) return fail(SCS_INVALID_ARG, "malformed colptr");
  for (int64_t j = 0; j < m; ++j)
    if (colptr[j + 1] < colptr[j]) return fail(SCS_INVALID_ARG, "colptr must be non-decreasing");
  if (nnz > 0 && (!rowval || !nzval)) return fail(SCS_INVALID_ARG, "rowval or nzval is NULL");
  for (int64_t p = 0; p < nnz; ++p) {
    const int64_t i = rowval[p] - index_base;
    if (i < 0 || i >= n_local) return fail(SCS_INVALID_ARG, "rowval entry outside 1..n_local");
  }
  // auto: the sparse kernels move 12 bytes per stored entry and gather x / r at random, the dense ones stream 8 bytes
  // per element at the HBM roofline and have the tensor-core Gram: sparse pays off below a few percent density
  const bool sparse = storage == 2 || (storage == 0 && (double)nnz < 0.04 * (double)n_local * (double)m && m <= 28160 &&
                                       n_local < (1LL << 31) && loss_kind != SCS_LOSS_QUADFORM);
  if (sparse && (m > 28160 || n_local >= (1LL << 31) || loss_kind == SCS_LOSS_QUADFORM))
    return fail(SCS_UNSUPPORTED, "sparse storage needs m <= 28160, n_local < 2^31 and a row-separable loss");
  int s = problem_alloc(ctx, n_local, m, loss_kind, loss_param, label_mode, out, !sparse);
  if (s != SCS_OK) {
    if (out && *out) {
      scs_problem_destroy(*out);
      *out = nullptr;
    }
    return s;
  }
  scs_problem* p = *out;
  auto bail = [&](cudaError_t e, const char* what) {
    scs_problem_destroy(p);
    *out = nullptr;
    return fail(e == cudaErrorMemoryAllocation ? SCS_OOM : SCS_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
  };
  p->labels_unit = labels_in_unit_range(y, n_local);
  cudaError_t e = cudaMemcpyAsync(p->dy, y, n_local * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return bail(e, "upload of y failed");
  // zero-based CSC on the host (32-bit row indices), then CSR by a counting pass (columns ascend inside every row)
  std::vector<int64_t> cp(m + 1);
  for (int64_t j = 0; j <= m; ++j) cp[j] = colptr[j] - index_base;
  if (!sparse) {
    int64_t *d_cp = nullptr, *d_rv = nullptr;
    double* d_nz = nullptr;
    int* d_bad = nullptr;
    e = cudaMalloc((void**)&d_cp, (m + 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_rv, std::max<int64_t>(nnz, 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_nz, std::max<int64_t>(nnz, 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_bad, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cp, colptr, (m + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_rv, rowval, nnz * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(d_nz, nzval, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      k_scatter_csc<<<(unsigned)std::min<int64_t>(m, 4096), 256, 0, ctx->stream>>>(d_cp, d_rv, d_nz, index_base, n_local,
                                                                                (int)m, p->ldd, p->dA, d_bad);
      ctx->launches += 1;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dfree(d_cp), dfree(d_rv), dfree(d_nz), dfree(d_bad);
    if (e != cudaSuccess) return bail(e, "CSC upload failed");
    return SCS_OK;
  }
  p->nnz = nnz;
  std::vector<int> ri(std::max<int64_t>(nnz, 1)), ci(std::max<int64_t>(nnz, 1));
  std::vector<int64_t> rp(n_local + 1, 0);
  std::vector<double> rv(std::max<int64_t>(nnz, 1));
  for (int64_t q = 0; q < nnz; ++q) {
    ri[q] = (int)(rowval[q] - index_base);
    rp[ri[q] + 1] += 1;
  }
  for (int64_t i = 0; i < n_local; ++i) rp[i + 1] += rp[i];
  {
    std::vector<int64_t> fill(rp.begin(), rp.end() - 1);
    for (int64_t j = 0; j < m; ++j)
      for (int64_t q = cp[j]; q < cp[j + 1]; ++q) {
        const int64_t dst = fill[ri[q]]++;
        ci[dst] = (int)j;
        rv[dst] = nzval[q];
      }
  }
  const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
  e = cudaMalloc((void**)&p->d_colptr, (m + 1) * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_rowptr, (n_local + 1) * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_rowidx, nz * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_colidx, nz * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_cvals, nz * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_vals, nz * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_colptr, cp.data(), (m + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rowptr, rp.data(), (n_local + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && nnz > 0) {
    e = cudaMemcpyAsync(p->d_rowidx, ri.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_colidx, ci.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_cvals, nzval, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_vals, rv.data(), nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return bail(e, "sparse upload failed");
  return SCS_OK;
}
extern "C" int scs_problem_is_sparse(scs_problem* p, int* sparse, int64_t* nnz) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (sparse) *sparse = p->sparse ? 1 : 0;
  if (nnz) *nnz = p->nnz;
  return SCS_OK;
}

extern "C" int scs_problem_create_synthetic(scs_ctx* ctx, int64_t n_total, int64_t row0, int64_t n_local, int64_t m,
                                            int loss_kind, double loss_param, int label_mode, uint64_t seed,
                                            double density, scs_problem** out) {
  if (loss_kind == SCS_LOSS_QUADFORM) return fail(SCS_INVALID_ARG, "no synthetic generator for quadform");
  if (row0 < 0 || row0 + n_local > n_total) return fail(SCS_INVALID_ARG, "row range outside n_total");
  int s = problem_alloc(ctx, n_local, m, loss_kind, loss_param, label_mode, out);
  if (s != SCS_OK) {
    if (out && *out) {
      scs_problem_destroy(*out);
      *out = nullptr;
    }
    return s;
  }
  scs_problem* p = *out;
  const unsigned gy = (unsigned)std::min<int64_t>(m, 1024);
  LAUNCH(ctx, k_synth_A, dim3((unsigned)((n_local + 255) / 256), gy), 256, 0, p->dA, p->ldd, n_local, (int)m, row0,
         n_total, seed, density, 1.0 / std::sqrt((double)m));
  LAUNCH(ctx, k_synth_xtrue, (unsigned)((m + 255) / 256), 256, 0, p->d_t1, (int)m, seed + 1, 0.05, 3.0);
  // z = A x_true (a forward pass with the z-only loss kind), then labels / targets
  LossParams lp = p->loss;
  lp.kind = 2;
  LAUNCH(ctx, k_forward<8>, (unsigned)p->fwd_blocks, kFwdThreads, 0, p->dA, p->ldd, p->ldd, (int64_t)0, p->n, (int)m,
         p->d_t1, p->dy, lp, p->dz, (double*)nullptr, (double*)nullptr, p->d_losspart);
  LAUNCH(ctx, k_synth_y, (unsigned)((n_local + 255) / 256), 256, 0, p->dy, p->dz, n_local, row0, seed + 2,
         loss_kind == SCS_LOSS_LOGISTIC ? 0 : 1, 0.1);
  CU_TRY(cudaMemsetAsync(p->dz, 0, p->ldd * sizeof(double), ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  p->labels_unit = loss_kind == SCS_LOSS_LOGISTIC;  // k_synth_y draws +-1 labels
  return SCS_OK;
}

// exclusive scan of counts[0..n) in place, counts[n] = total (device)
static int device_scan(scs_ctx* c, int64_t* counts, int64_t n) {
  const int64_t nb = (n + kScanBlock - 1) / kScanBlock;
  int64_t* btot = nullptr;
  CU_TRY(cudaMalloc((void**)&btot, std::max<int64_t>(nb, 1) * sizeof(int64_t)));
  LAUNCH(c, k_scan_blocks, (unsigned)nb, 256, 0, counts, n, btot);
  LAUNCH(c, k_scan_top, 1, 1, 0, btot, nb, counts + n);
  LAUNCH(c, k_scan_add, (unsigned)nb, 256, 0, counts, n, (const int64_t*)btot);
  CU_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(btot);
  return SCS_OK;
}

// Turn a dense resident shard into the sparse layout (CSR + CSC copies) on the device and free the dense matrix.  For
// benchmark-sized synthetic shards (scs_problem_create_synthetic with density < 1): must be called before the first pass.
extern "C" int scs_problem_sparsify(scs_problem* p) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (p->sparse) return SCS_OK;
  if (p->loss.kind == SCS_LOSS_QUADFORM || p->m > 28160 || p->n >= (1LL << 31))
    return fail(SCS_UNSUPPORTED, "sparse storage needs m <= 28160, n_local < 2^31 and a row-separable loss");
  if (p->gram_ready || p->i8_ready || p->fu_ready || p->fwd_id != 0)
    return fail(SCS_STATE_ERROR, "scs_problem_sparsify must be called right after the problem was created");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  const int64_t n = p->n;
  const int m = (int)p->m;
  CU_TRY(cudaMalloc((void**)&p->d_rowptr, (n + 1) * sizeof(int64_t)));
  CU_TRY(cudaMalloc((void**)&p->d_colptr, (m + 1) * sizeof(int64_t)));
  LAUNCH(c, k_nnz_rows, (unsigned)((n + 255) / 256), 256, 0, (const double*)p->dA, p->ldd, n, m, p->d_rowptr);
  LAUNCH(c, k_nnz_cols, (unsigned)m, 256, 0, (const double*)p->dA, p->ldd, n, m, p->d_colptr);
  SCS_TRY(device_scan(c, p->d_rowptr, n));
  SCS_TRY(device_scan(c, p->d_colptr, m));
  int64_t nnz = 0, nnz2 = 0;
  CU_TRY(cudaMemcpy(&nnz, p->d_rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(&nnz2, p->d_colptr + m, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (nnz != nnz2) return fail(SCS_STATE_ERROR, "sparsify: row and column counts disagree");
  const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
  CU_TRY(cudaMalloc((void**)&p->d_colidx, nz * sizeof(int)));
  CU_TRY(cudaMalloc((void**)&p->d_rowidx, nz * sizeof(int)));
  CU_TRY(cudaMalloc((void**)&p->d_vals, nz * sizeof(double)));
  CU_TRY(cudaMalloc((void**)&p->d_cvals, nz * sizeof(double)));
  LAUNCH(c, k_fill_csr, (unsigned)((n + 255) / 256), 256, 0, (const double*)p->dA, p->ldd, n, m,
         (const int64_t*)p->d_rowptr, p->d_colidx, p->d_vals);
  LAUNCH(c, k_fill_csc, (unsigned)m, 256, 0, (const double*)p->dA, p->ldd, n, m, (const int64_t*)p->d_colptr, p->d_rowidx,
         p->d_cvals);
  CU_TRY(cudaStreamSynchronize(c->stream));
  dfree(p->dA);
  p->dA = nullptr;
  dfree(p->d_adjpart);
  p->d_adjpart = nullptr;
  p->sparse = true;
  p->nnz = nnz;
  p->fwd_blocks = (n + kSpFwdRows - 1) / kSpFwdRows;
  dfree(p->d_losspart);
  p->d_losspart = nullptr;
  SCS_TRY(dalloc(&p->d_losspart, p->fwd_blocks));
  return SCS_OK;
}

extern "C" int scs_problem_read_rows(scs_problem* p, int64_t row0, int64_t nrows, double* A_out, double* y_out) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (row0 < 0 || nrows < 0 || row0 + nrows > p->n) return fail(SCS_INVALID_ARG, "row range outside the shard");
  if (p->sparse) return fail(SCS_UNSUPPORTED, "scs_problem_read_rows: the shard is resident in sparse form");
  CU_TRY(cudaSetDevice(p->ctx->device));
  CU_TRY(cudaStreamSynchronize(p->ctx->stream));
  if (A_out && nrows > 0)
    CU_TRY(cudaMemcpy2D(A_out, nrows * sizeof(double), p->dA + row0, p->ldd * sizeof(double),
                        nrows * sizeof(double), p->m, cudaMemcpyDeviceToHost));
  if (y_out && nrows > 0) CU_TRY(cudaMemcpy(y_out, p->dy + row0, nrows * sizeof(double), cudaMemcpyDeviceToHost));
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// exported: configuration
// ------------------------------------------------------------------------------------------------
static int upload_bounds(double** dst, const double* src, int64_t n, bool sanity) {
  std::vector<double> h(src, src + n);
  if (sanity)  // bounds_sanity_check: ±Inf -> ±1e32 (prox-reg-utils.jl:156-157)
    for (auto& v : h) {
      if (v == -std::numeric_limits<double>::infinity()) v = -1e32;
      if (v == std::numeric_limits<double>::infinity()) v = 1e32;
    }
  dfree(*dst);
  *dst = nullptr;
  CU_TRY(cudaMalloc((void**)dst, n * sizeof(double)));
  CU_TRY(cudaMemcpy(*dst, h.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  return SCS_OK;
}

extern "C" int scs_set_regularizer(scs_problem* p, int reg_kind, double lam1, double lam2, const int64_t* ind3xG,
                                   int64_t ngroups, const int64_t* perm, const double* lb, int64_t nlb,
                                   const double* ub, int64_t nub) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (reg_kind < 0 || reg_kind > 3) return fail(SCS_INVALID_ARG, "reg_name not valid.");
  CU_TRY(cudaSetDevice(p->ctx->device));
  RegDesc rd{};
  rd.kind = reg_kind;
  rd.lam1 = lam1;
  rd.lam2 = lam2;
  if (reg_kind == SCS_REG_INDBOX) {
    if (!lb || !ub) return fail(SCS_INVALID_ARG, "indbox needs C_set bounds");
    if (!((nlb == 1 || nlb == p->m) && (nub == 1 || nub == p->m)))
      return fail(SCS_INVALID_ARG, "Lengths of the bounds do not match that of the variable.");
    SCS_TRY(upload_bounds(&p->d_rlb, lb, nlb, false));
    SCS_TRY(upload_bounds(&p->d_rub, ub, nub, false));
    rd.lb = p->d_rlb;
    rd.ub = p->d_rub;
    rd.nlb = (int)nlb;
    rd.nub = (int)nub;
  }
  if (reg_kind == SCS_REG_GL) {
    if (!ind3xG || ngroups < 1) return fail(SCS_INVALID_ARG, "gl needs the group index table (model.P)");
    // groups must tile 1..m contiguously (then Cmat == diag(weights), prox-reg-utils.jl:121-142)
    std::vector<double> cd(p->m, 0.0);
    int64_t expect = 1;
    for (int64_t g = 0; g < ngroups; ++g) {
      const int64_t gs = ind3xG[3 * g], ge = ind3xG[3 * g + 1], gw = ind3xG[3 * g + 2];
      if (gs != expect || ge < gs || ge > p->m)
        return fail(SCS_UNSUPPORTED, "gl groups must be contiguous, non-overlapping and cover 1..m in order");
      for (int64_t k = gs; k <= ge; ++k) cd[k - 1] = (double)gw;
      expect = ge + 1;
    }
    if (expect != p->m + 1) return fail(SCS_UNSUPPORTED, "gl groups must cover all m variables");
    dfree(p->d_ind);
    p->d_ind = nullptr;
    CU_TRY(cudaMalloc((void**)&p->d_ind, 3 * ngroups * sizeof(int64_t)));
    CU_TRY(cudaMemcpy(p->d_ind, ind3xG, 3 * ngroups * sizeof(int64_t), cudaMemcpyHostToDevice));
    dfree(p->d_perm);
    p->d_perm = nullptr;
    if (perm) {
      for (int64_t k = 0; k < p->m; ++k)
        if (perm[k] < 1 || perm[k] > p->m) return fail(SCS_INVALID_ARG, "perm entries must be in 1..m");
      CU_TRY(cudaMalloc((void**)&p->d_perm, p->m * sizeof(int64_t)));
      CU_TRY(cudaMemcpy(p->d_perm, perm, p->m * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    dfree(p->d_cdiag);
    p->d_cdiag = nullptr;
    CU_TRY(cudaMalloc((void**)&p->d_cdiag, p->m * sizeof(double)));
    CU_TRY(cudaMemcpy(p->d_cdiag, cd.data(), p->m * sizeof(double), cudaMemcpyHostToDevice));
    rd.ind = p->d_ind;
    rd.ngroups = (int)ngroups;
    rd.perm = p->d_perm;
    p->sm.cdiag = p->d_cdiag;
  }
  p->reg = rd;
  p->has_reg = true;
  return SCS_OK;
}

extern "C" int scs_set_smoother(scs_problem* p, int kind, double mu, const double* lb, int64_t nlb, const double* ub,
                                int64_t nub) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (kind < 0 || kind > 6) return fail(SCS_INVALID_ARG, "unknown smoother kind");
  CU_TRY(cudaSetDevice(p->ctx->device));
  SmoothDesc sd{};
  sd.kind = kind;
  sd.mu = mu;
  sd.cdiag = p->d_cdiag;
  const bool box = kind == SCS_SMOOTH_PHUBER_INDBOX || kind == SCS_SMOOTH_EXP_INDBOX || kind == SCS_SMOOTH_LOGEXP_INDBOX;
  if (box) {
    if (!lb || !ub) return fail(SCS_INVALID_ARG, "IndBox smoothers need lb and ub");
    if (!((nlb == 1 && nub == 1) || (nlb == p->m && nub == p->m)))
      return fail(SCS_INVALID_ARG, "Lengths of the bounds do not match that of the variable.");
    SCS_TRY(upload_bounds(&p->d_slb, lb, nlb, true));
    SCS_TRY(upload_bounds(&p->d_sub, ub, nub, true));
    sd.lb = p->d_slb;
    sd.ub = p->d_sub;
    sd.nlb = (int)nlb;
    sd.nub = (int)nub;
  }
  switch (kind) {  // Mh, ν constants: phuber-smooth.jl:3-4, exponential-smooth.jl:25-26, log-exp-smooth.jl:25-26,
                   // ostrovskii-bach-smooth.jl:3-4
    case SCS_SMOOTH_PHUBER_L1L2:
    case SCS_SMOOTH_PHUBER_INDBOX:
    case SCS_SMOOTH_PHUBER_GL:
      p->Mh = 2.0;
      p->nu = 2.6;
      break;
    case SCS_SMOOTH_EXP_INDBOX:
    case SCS_SMOOTH_LOGEXP_INDBOX:
      p->Mh = 1.0;
      p->nu = 2.0;
      break;
    default:
      p->Mh = 2.0 * std::sqrt(2.0);
      p->nu = 3.0;
  }
  p->sm = sd;
  p->has_sm = true;
  return SCS_OK;
}

extern "C" int scs_set_method(scs_problem* p, int method_kind, int ss_type, int use_prox, int lbfgs_m) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (method_kind < 0 || method_kind > 2) return fail(SCS_INVALID_ARG, "unknown method kind");
  CU_TRY(cudaSetDevice(p->ctx->device));
  if (method_kind == SCS_METHOD_GGN && p->loss.kind == SCS_LOSS_QUADFORM)
    return fail(SCS_UNSUPPORTED, "ProxGGNSCORE needs a model output function (out_fn); the quadform loss has none");
  p->method = method_kind;
  p->ss_type = ss_type;  // validated in scs_step, like the reference (error raised inside step!)
  p->use_prox = use_prox != 0;
  if (method_kind == SCS_METHOD_LQN) {
    if (lbfgs_m < 1 || lbfgs_m > 64) return fail(SCS_INVALID_ARG, "L-BFGS memory must be in 1..64");
    if (lbfgs_m != p->lbfgs_cap) {
      dfree(p->d_S);
      dfree(p->d_Y);
      dfree(p->d_state);
      p->d_S = p->d_Y = nullptr;
      p->d_state = nullptr;
      SCS_TRY(dalloc(&p->d_S, (size_t)lbfgs_m * p->m));
      SCS_TRY(dalloc(&p->d_Y, (size_t)lbfgs_m * p->m));
      CU_TRY(cudaMalloc((void**)&p->d_state, 2 * sizeof(int64_t)));
      CU_TRY(cudaMemset(p->d_state, 0, 2 * sizeof(int64_t)));
      p->lbfgs_cap = lbfgs_m;
    }
    p->lbfgs_m = lbfgs_m;
  }
  p->has_method = true;
  return SCS_OK;
}

extern "C" int scs_set_L(scs_problem* p, int has_L, double L) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  p->has_L = has_L != 0;
  p->L = L;
  return SCS_OK;
}

extern "C" int scs_set_gram_mode(scs_problem* p, int mode) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (mode < 0 || mode > 2) return fail(SCS_INVALID_ARG, "gram mode must be 0 (auto), 1 (DMMA) or 2 (tcgen05 int8)");
  p->gram_mode = mode;
  return SCS_OK;
}
extern "C" int scs_set_active_rows(scs_problem* p, int64_t row_lo, int64_t row_hi) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  CU_TRY(cudaSetDevice(p->ctx->device));
  return set_window(p, row_lo, row_hi);
}
extern "C" int scs_set_batches(scs_problem* p, int64_t nbatch, const int64_t* offsets) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (nbatch < 0 || (nbatch > 0 && !offsets)) return fail(SCS_INVALID_ARG, "bad batch table");
  CU_TRY(cudaSetDevice(p->ctx->device));
  p->batch_off.clear();
  p->batch_rows_global.clear();
  if (nbatch == 0) return set_window(p, 0, p->n);
  if (offsets[0] != 0 || offsets[nbatch] > p->n) return fail(SCS_INVALID_ARG, "batch offsets must start at 0 and stay inside the shard");
  for (int64_t i = 0; i < nbatch; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(SCS_INVALID_ARG, "batch offsets must be non-decreasing");
  p->batch_off.assign(offsets, offsets + nbatch + 1);
  // global batch sizes (and the global row count) in one all-reduce
  std::vector<double> cnt(nbatch + 1);
  for (int64_t i = 0; i < nbatch; ++i) cnt[i] = (double)(offsets[i + 1] - offsets[i]);
  cnt[nbatch] = (double)p->n;
  if (p->ctx->world > 1) {
    double* d = nullptr;
    CU_TRY(cudaMalloc((void**)&d, cnt.size() * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(d, cnt.data(), cnt.size() * sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    int rc = allreduce(p->ctx, d, cnt.size());
    if (rc == SCS_OK) {
      cudaMemcpyAsync(cnt.data(), d, cnt.size() * sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream);
      cudaStreamSynchronize(p->ctx->stream);
    }
    cudaFree(d);
    SCS_TRY(rc);
  }
  for (double v : cnt) p->batch_rows_global.push_back((int64_t)(v + 0.5));
  return SCS_OK;
}
extern "C" int scs_set_test_problem(scs_problem* p, scs_problem* test) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (test && (test->m != p->m || test->ctx != p->ctx))
    return fail(SCS_INVALID_ARG, "the test problem must live in the same context and have the same number of columns");
  p->test = test;
  return SCS_OK;
}
extern "C" int scs_get_test_history(scs_problem* p, double* out, int64_t cap, int64_t* n) {
  if (!p || !n) return fail(SCS_INVALID_ARG, "NULL argument");
  *n = (int64_t)p->test_hist.size();
  if (out)
    for (int64_t i = 0; i < std::min<int64_t>(cap, *n); ++i) out[i] = p->test_hist[i];
  return SCS_OK;
}
extern "C" int scs_set_stream_mode(scs_problem* p, int mode) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (mode < 0 || mode > 2) return fail(SCS_INVALID_ARG, "stream mode must be 0 (auto), 1 (two passes) or 2 (fused)");
  p->stream_mode = mode;
  return SCS_OK;
}
extern "C" int scs_get_stream_path(scs_problem* p, int* path) {
  if (!p || !path) return fail(SCS_INVALID_ARG, "NULL argument");
  *path = p->last_stream_path;
  return SCS_OK;
}
extern "C" int scs_set_gram_bits(scs_problem* p, int bits) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (bits < 24 || bits > 58) return fail(SCS_INVALID_ARG, "gram bits must be in 24..58");
  if (p->i8_ready) return fail(SCS_STATE_ERROR, "scs_set_gram_bits must be called before the first Gram");
  p->i8_bits = bits;
  return SCS_OK;
}
extern "C" int scs_get_gram_info(scs_problem* p, int* nmod, int* bits) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (nmod) *nmod = p->i8_nmod;
  if (bits) *bits = p->i8_b;
  return SCS_OK;
}
extern "C" int scs_get_gram_signed(scs_problem* p, int64_t* compact_rows, int* minority_negative) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (compact_rows) *compact_rows = p->i8_signed ? p->i8_ccount : -1;
  if (minority_negative) *minority_negative = p->i8_minor_neg ? 1 : 0;
  return SCS_OK;
}
extern "C" int scs_get_gram_path(scs_problem* p, int* path) {
  if (!p || !path) return fail(SCS_INVALID_ARG, "NULL argument");
  *path = p->last_gram_path;
  return SCS_OK;
}

extern "C" int scs_method_init(scs_problem* p) {  // init!(method, x): prox-L-BFGS-SCORE.jl:31-36
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (!p->has_method) return fail(SCS_STATE_ERROR, "scs_set_method has not been called");
  CU_TRY(cudaSetDevice(p->ctx->device));
  if (p->method == SCS_METHOD_LQN) {
    CU_TRY(cudaMemsetAsync(p->d_state, 0, 2 * sizeof(int64_t), p->ctx->stream));
    const double one = 1.0;
    CU_TRY(cudaMemcpyAsync(p->d_scal + SC_H0, &one, sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    CU_TRY(cudaStreamSynchronize(p->ctx->stream));
  }
  p->gq_id = 0;
  p->gqprev_id = 0;
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// x interning: host vectors get ids so that passes are reused when the same iterate comes back
// ------------------------------------------------------------------------------------------------
static int intern_x(scs_problem* p, const double* hx, double* dbuf, XRef* out) {
  const size_t bytes = p->m * sizeof(double);
  uint64_t id = 0;
  for (auto& s : p->shadow)
    if (s.id != 0 && s.x.size() == (size_t)p->m && memcmp(s.x.data(), hx, bytes) == 0) {
      id = s.id;
      break;
    }
  if (id == 0) {
    auto& s = p->shadow[p->shadow_next];
    p->shadow_next = (p->shadow_next + 1) % 4;
    s.x.assign(hx, hx + p->m);
    s.id = id = p->next_id++;
  }
  CU_TRY(cudaMemcpyAsync(dbuf, hx, bytes, cudaMemcpyHostToDevice, p->ctx->stream));
  out->d = dbuf;
  out->id = id;
  return SCS_OK;
}
static void remember_x(scs_problem* p, const double* hx, uint64_t id) {
  auto& s = p->shadow[p->shadow_next];
  p->shadow_next = (p->shadow_next + 1) % 4;
  s.x.assign(hx, hx + p->m);
  s.id = id;
}

static int get_Mg(const scs_problem* p, double* Mg) {  // smoothing.jl:12-26
  if (p->Mh < 0) return fail(SCS_INVALID_ARG, "Mh must be nonnegative.");
  if (!(p->sm.mu > 0)) return fail(SCS_INVALID_ARG, "μ must be positive.");
  if (p->nu > 0 && p->nu <= 3)
    *Mg = std::pow((double)p->m, (3 - p->nu) / 2) * std::pow(p->sm.mu, p->nu / 2 - 2) * p->Mh;
  else if (p->nu > 3)
    *Mg = std::pow(p->sm.mu, 4 - 3 * p->nu / 2) * p->Mh;
  else
    return fail(SCS_INVALID_ARG, "ν must be positive.");
  return SCS_OK;
}

static int check_ready(const scs_problem* p) {
  if (!p) return fail(SCS_INVALID_ARG, "problem is NULL");
  if (!p->has_reg) return fail(SCS_STATE_ERROR, "scs_set_regularizer has not been called");
  return SCS_OK;
}

// objective pieces at x: loss sum (all-reduced) in d_gl[m], reg value in d_scal[SC_REGX]
static int objective_device(scs_problem* p, XRef x) {
  const int wk = p->method == SCS_METHOD_GGN ? SCS_WEIGHTS_GGN : SCS_WEIGHTS_NEWTON;
  // reuse whatever forward pass is cached for this x: the loss value does not depend on the weight kind
  if (p->fwd_id == x.id)
    SCS_TRY(ensure_loss(p, x, p->fwd_wk));
  else if (p->has_method && p->has_sm && fused_wanted(p))
    SCS_TRY(ensure_grad(p, x, wk));  // step! at the same x follows (iterate.jl:189 then :233): one read of A serves both
  else
    SCS_TRY(ensure_loss(p, x, wk));
  StageTimer t(p->ctx, ST_VEC);
  LAUNCH(p->ctx, k_reg, 1, kVecThreads, 0, p->reg, x.d, (int)p->m, p->d_scal + SC_REGX);
  return SCS_OK;
}

extern "C" int scs_objective(scs_problem* p, const double* x, double* fval, double* reg) {
  SCS_TRY(check_ready(p));
  if (!x) return fail(SCS_INVALID_ARG, "x is NULL");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  XRef xr;
  SCS_TRY(intern_x(p, x, p->vx[0], &xr));
  SCS_TRY(objective_device(p, xr));
  CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_gl + p->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(p->h_scal + 1, p->d_scal + SC_REGX, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  if (fval) *fval = fval_from_sum(p, p->h_scal[0]);
  if (reg) *reg = p->h_scal[1];
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
// ∇q(v) = grad_f(v) + λ·hμ.grad(v) into dout (invalidates the forward cache unless v is the cached x)
static int compute_gq(scs_problem* p, XRef v, double lam, double* dout) {
  SCS_TRY(ensure_grad(p, v, SCS_WEIGHTS_NEWTON));
  StageTimer t(p->ctx, ST_VEC);
  LAUNCH(p->ctx, k_pre, 1, kVecThreads, 0, p->sm, lam, v.d, p->d_gl, (int)p->m, p->d_t1, p->d_t2, dout, p->d_scal + SC_SCRATCH);
  return SCS_OK;
}

// Armijo backtracking of utils.jl:27-35 with f = nonsmooth objective, grad = smoothed ∇q.  f(x) and <∇q,d> are
// evaluated once (the reference recomputes the same values every trial).  The host loop below is only used for the
// quadform loss (not a sum over rows); everything else goes through linesearch_device.
static int linesearch_hostloop(scs_problem* p, XRef x, const double* d_dir, double dsign, const double* d_gqx,
                               double* ss_out) {
  scs_ctx* c = p->ctx;
  SCS_TRY(ensure_loss(p, x, p->fwd_id == x.id ? p->fwd_wk : SCS_WEIGHTS_NEWTON));
  LAUNCH(c, k_reg, 1, kVecThreads, 0, p->reg, x.d, (int)p->m, p->d_scal + SC_REGX);
  LAUNCH(c, k_dot, 1, kVecThreads, 0, d_gqx, d_dir, (int)p->m, p->d_scal, (int)SC_GD);
  CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_gl + p->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(p->h_scal + 1, p->d_scal + SC_REGX, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(p->h_scal + 2, p->d_scal + SC_GD, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  const double fx = fval_from_sum(p, p->h_scal[0]) + p->h_scal[1];
  const double gd = dsign * p->h_scal[2];
  double alpha = 1.0;
  for (int it = 0; it < 4000; ++it) {
    LAUNCH(c, k_axpy_out, (unsigned)((p->m + 255) / 256), 256, 0, x.d, dsign * alpha, d_dir, (int)p->m, p->d_trial);
    XRef tr{p->d_trial, p->next_id++};
    SCS_TRY(ensure_loss(p, tr, SCS_WEIGHTS_NEWTON));
    LAUNCH(c, k_reg, 1, kVecThreads, 0, p->reg, p->d_trial, (int)p->m, p->d_scal + SC_REGX);
    CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_gl + p->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(p->h_scal + 1, p->d_scal + SC_REGX, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    const double ft = fval_from_sum(p, p->h_scal[0]) + p->h_scal[1];
    if (!(ft > fx + 1e-4 * alpha * gd)) break;
    alpha = 0.5 * alpha;
  }
  const double a = alpha;
  CU_TRY(cudaMemcpyAsync(p->d_scal + SC_SS, &a, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  *ss_out = alpha;
  return SCS_OK;
}

// The same search with the trials on the device: ONE pass over A gives zd = A d; with z = A x from the pass at x every
// trial point is z + α zd, so a batch of 8 step sizes (α, α/2, ..., α/128) costs one sweep over two row vectors
// (k_ls_terms), one 8-double all-reduce when the rows are sharded, and one single-CTA kernel that adds get_reg(x + α d)
// and applies the Armijo test in order (k_ls_pick).  The accepted α stays on the device (scal[SC_SS], read by k_tail);
// the host looks at the "found" flag once per batch instead of synchronising on every trial.
static int linesearch_device(scs_problem* p, XRef x, const double* d_dir, double dsign, const double* d_gqx,
                             double* ss_out) {
  scs_ctx* c = p->ctx;
  if (p->loss.kind == SCS_LOSS_QUADFORM) return linesearch_hostloop(p, x, d_dir, dsign, d_gqx, ss_out);
  const int m = (int)p->m;
  // z = A x and the loss sum at x: whatever pass produced them is cached by id
  SCS_TRY(ensure_loss(p, x, p->fwd_id == x.id ? p->fwd_wk : SCS_WEIGHTS_NEWTON));
  if (!p->d_u) SCS_TRY(dalloc(&p->d_u, p->ldd));
  const int64_t r0 = p->sparse ? 0 : p->alo, nproc = p->sparse ? p->n : p->ahi - p->alo;
  const int64_t nblk = std::max<int64_t>(1, (nproc + kLsRowsPerBlock - 1) / kLsRowsPerBlock);
  if (!p->d_lspart || p->lspart_cap < nblk) {
    dfree(p->d_lspart);
    p->d_lspart = nullptr;
    SCS_TRY(dalloc(&p->d_lspart, (size_t)nblk * kLsTrials + kLsTrials + LS_COUNT));
    p->lspart_cap = nblk;
  }
  double* sums = p->d_lspart + (size_t)p->lspart_cap * kLsTrials;
  double* ls = sums + kLsTrials;
  SCS_TRY(run_matvec(p, d_dir, p->d_u));
  CU_TRY(cudaMemsetAsync(p->d_scal + SC_LSFOUND, 0, sizeof(double), c->stream));
  double alpha0 = 1.0;
  for (int batch = 0; batch < 140; ++batch) {  // 140 * 8 halvings: past the point where α underflows to 0 (accepts)
    StageTimer t(c, ST_VEC);
    LAUNCH(c, k_ls_terms, (unsigned)nblk, 256, 0, (const double*)(p->dz + r0), (const double*)(p->d_u + r0),
           (const double*)(p->dy + r0), nproc, p->win_lo - r0, p->win_hi - r0, p->loss, dsign, alpha0,
           (const double*)p->d_scal, p->d_lspart);
    LAUNCH(c, k_ls_sum, 1, kVecThreads, 0, (const double*)p->d_lspart, nblk, (const double*)p->d_scal, sums);
    SCS_TRY(allreduce(c, sums, kLsTrials));  // every rank queues the same sequence (the flag is replicated)
    LAUNCH(c, k_ls_pick, 1, kVecThreads, 0, p->reg, p->loss.kind, p->loss.p, (const double*)x.d, d_dir, dsign, d_gqx, m,
           alpha0, batch == 0 ? 1 : 0, (const double*)(p->d_gl + p->m), (const double*)sums, p->d_trial, ls, p->d_scal);
    CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_scal + SC_SS, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (p->h_scal[1] != 0.0) {
      *ss_out = p->h_scal[0];
      return SCS_OK;
    }
    alpha0 = std::ldexp(alpha0, -kLsTrials);
    if (alpha0 == 0.0) break;
  }
  // α underflowed: f(x + 0 d) > f(x) is false, the reference accepts α = 0
  const double zero = 0.0;
  CU_TRY(cudaMemcpyAsync(p->d_scal + SC_SS, &zero, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  *ss_out = 0.0;
  return SCS_OK;
}

// One step! on device vectors.  x, x_prev: inputs; xnew: output buffer.  Scalars land in h_scal after the sync
// the caller performs.
static int step_device(scs_problem* p, XRef x, XRef xprev, int64_t iter, double* xnew, uint64_t xnew_id,
                       const double* d_xstar) {
  scs_ctx* c = p->ctx;
  const int m = (int)p->m;
  const double lam = p->reg.lam1;  // model.λ or model.λ[1]
  double Mg = 0;
  SCS_TRY(get_Mg(p, &Mg));
  double ss = 0.5;
  const bool type1 = p->ss_type == 1;
  if (type1) ss = p->has_L ? std::min(1.0 / p->L, 1.0) : 0.5;

  if (p->method == SCS_METHOD_N || p->method == SCS_METHOD_GGN) {
    const bool ggn = p->method == SCS_METHOD_GGN;
    const int wk = ggn ? SCS_WEIGHTS_GGN : SCS_WEIGHTS_NEWTON;
    const bool wide = ggn && p->win_rows_global + 1 <= p->m;  // prox-GGN-SCORE.jl:124
    if (!type1 && p->ss_type != 2 && p->ss_type != 3) return fail(SCS_INVALID_ARG, "Please, choose ss_type in [1, 2, 3].");
    if (p->ss_type == 2) {
      // prox-N-SCORE.jl:81-83 / prox-GGN-SCORE.jl:78-80 reference an undefined ∇f: the reference throws after iter 1
      if (iter == 1)
        ss = 1.0;
      else
        return fail(SCS_UNSUPPORTED,
                    "ss_type=2 is broken in the reference for ProxNSCORE/ProxGGNSCORE (UndefVarError: ∇f); rejected");
    }
    if (wide) {
      SCS_TRY(run_ggn_wide(p, x, lam));
    } else {
      SCS_TRY(ensure_grad(p, x, wk));
      SCS_TRY(run_gram(p, x));
      {
        StageTimer t(c, ST_VEC);
        LAUNCH(c, k_pre, 1, kVecThreads, 0, p->sm, lam, x.d, p->d_gl, m, p->d_gr, p->d_hr, p->d_rhs, p->d_scal);
        LAUNCH(c, k_add_diag, (m + 255) / 256, 256, 0, p->d_G, m, lam, p->d_hr);
        CU_TRY(cudaMemcpyAsync(p->d_q, p->d_rhs, m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      }
      SCS_TRY(run_solve(c, p->d_G, p->d_Gsave, p->d_Linv, p->d_info, p->d_q, p->d_t1, p->d_sol, m,
                        &p->last_used_fallback, true));
    }
    if (p->ss_type == 3) {
      const double* gqx = p->d_rhs;  // N: ∇q = grad_f + λgr
      if (ggn) {                     // GGN's rhs is J'res + λgr; the line search wants the gradient of f(A,y,x)
        // (compute_gq's k_pre recomputes gr / Hr / η² at the same x into scratch: d_gr, d_hr, SC_ETASQ stay valid)
        SCS_TRY(compute_gq(p, x, lam, p->d_gnewton));
        gqx = p->d_gnewton;
      }
      SCS_TRY(linesearch_device(p, x, p->d_sol, -1.0, gqx, &ss));
    }
    StageTimer t(c, ST_VEC);
    LAUNCH(c, k_tail, 1, kVecThreads, 0, p->reg, p->use_prox, ss, Mg, -1.0, x.d, p->d_sol, p->d_hr, d_xstar, m, xnew,
           p->d_dx, (double*)nullptr, p->d_scal, (const double*)nullptr);
    return SCS_OK;
  }

  // ---- ProxLQNSCORE: two persistent single-CTA kernels around the gradient pass --------------------------------
  //   k_lqn_head   : gr, Hr, ∇q (or the one carried over), two-loop recursion, step size, damping, prox, norms, get_reg
  //   k_fused_grad : f(x⁺), ∇f(x⁺) in one read of A
  //   k_lqn_update : (partial sums of the pass,) ∇q(x⁺), γ, curvature guard, memory push, H0
  if (!(type1 || p->ss_type == 2 || p->ss_type == 3 || !p->has_L))
    return fail(SCS_INVALID_ARG, "Please, choose ss_type in [1, 2, 3].");
  const bool carried = p->gq_id == x.id;  // ∇q(x) is the ∇q_new of the previous step (same x, same bits)
  if (!carried) SCS_TRY(ensure_grad(p, x, SCS_WEIGHTS_NEWTON));
  int mode = 0;  // 0: ss known here, 1: BB on the device, 2: direction only (line search follows)
  if (!type1) {
    if (p->ss_type == 2 || !p->has_L) {  // prox-L-BFGS-SCORE.jl:112-119 (ss_type 3 without L lands here too)
      if (iter == 1) {
        ss = 1.0;
      } else {
        mode = 1;
        if (p->gqprev_id != xprev.id) {
          // ∇q(x_prev) is not the one kept from the previous step: recompute it.  That evicts the cached pass at x, so
          // a ∇q(x) that has not been formed yet is formed first.
          if (!carried) {
            StageTimer t(c, ST_VEC);
            LAUNCH(c, k_pre, 1, kVecThreads, 0, p->sm, lam, x.d, p->d_gl, m, p->d_gr, p->d_hr, p->d_gq, p->d_scal);
            p->gq_id = x.id;
          }
          SCS_TRY(compute_gq(p, xprev, lam, p->d_gqprev));
          p->gqprev_id = xprev.id;
        }
      }
    } else {
      mode = 2;
    }
  }
  {
    StageTimer t(c, ST_VEC);
    const bool have_gq = p->gq_id == x.id;
    LbfgsMem mem{p->d_S, p->d_Y, p->d_state, p->lbfgs_cap};
    LAUNCH(c, k_lqn_head, 1, kVecThreads, 0, p->sm, p->reg, mem, lam, Mg, ss, mode, p->use_prox ? 1 : 0, iter,
           (const double*)x.d, (const double*)xprev.d, have_gq ? (const double*)nullptr : (const double*)p->d_gl, m,
           p->d_gr, p->d_hr, p->d_gq, (const double*)p->d_gqprev, p->d_q, p->d_d, d_xstar, xnew, p->d_dx, p->d_delta,
           p->d_scal);
    p->gq_id = x.id;
  }
  if (mode == 2) {
    SCS_TRY(linesearch_device(p, x, p->d_d, 1.0, p->d_gq, &ss));
    StageTimer t(c, ST_VEC);
    LAUNCH(c, k_tail, 1, kVecThreads, 0, p->reg, p->use_prox, ss, Mg, 1.0, x.d, p->d_d, p->d_hr, d_xstar, m, xnew,
           p->d_dx, p->d_delta, p->d_scal, (const double*)(p->d_scal + SC_SS));
  }
  // second gradient at x⁺ (prox-L-BFGS-SCORE.jl:148-150); it is next iteration's ∇q, and its loss value is next
  // epoch's objective, so neither is recomputed
  XRef xn{xnew, xnew_id};
  const bool fold = c->world == 1 && fused_wanted(p) && !(p->fwd_id == xn.id) && p->m <= (int64_t)kFuMaxCluster * kFuCols;
  int folded = 0;
  if (fold) {
    int rc = run_fused(p, xn.d, SCS_WEIGHTS_NEWTON, /*reduce=*/false);
    if (rc == SCS_OK) {
      p->fwd_id = xn.id;
      p->fwd_wk = SCS_WEIGHTS_NEWTON;
      p->loss_reduced = true;
      p->grad_id = xn.id;
      folded = 1;
    } else if (!(rc == SCS_UNSUPPORTED && p->stream_mode == 0)) {
      return rc;
    }
  }
  if (!folded) SCS_TRY(ensure_grad(p, xn, SCS_WEIGHTS_NEWTON));
  {
    StageTimer t(c, ST_VEC);
    LbfgsMem mem{p->d_S, p->d_Y, p->d_state, p->lbfgs_cap};
    LAUNCH(c, k_lqn_update, 1, kVecThreads, 0, p->sm, lam, mem, (const double*)xnew, p->d_gl,
           folded ? (const double*)p->d_fupart : (const double*)nullptr, (int64_t)p->fu_last_ncl,
           (const double*)p->d_fuloss, (int64_t)p->fu_last_ncl * p->fu_cluster, (const double*)p->d_delta, m, p->d_gq,
           p->d_gqprev, p->d_gamma, p->d_t1, p->d_t2, p->d_scal);
    p->gqprev_id = x.id;
    p->gq_id = xnew_id;
  }
  return SCS_OK;
}

extern "C" int scs_step(scs_problem* p, const double* x, const double* x_prev, int64_t iter, double* x_new, double* dx,
                        double* pri_res_norm) {
  SCS_TRY(check_ready(p));
  if (!p->has_sm) return fail(SCS_STATE_ERROR, "scs_set_smoother has not been called");
  if (!p->has_method) return fail(SCS_STATE_ERROR, "scs_set_method has not been called");
  if (!x || !x_new) return fail(SCS_INVALID_ARG, "x or x_new is NULL");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  XRef xr, xp;
  SCS_TRY(intern_x(p, x, p->vx[0], &xr));
  if (x_prev)
    SCS_TRY(intern_x(p, x_prev, p->vx[1], &xp));
  else
    xp = xr;
  const uint64_t nid = p->next_id++;
  SCS_TRY(step_device(p, xr, xp, iter, p->vx[2], nid, nullptr));
  CU_TRY(cudaMemcpyAsync(x_new, p->vx[2], p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (dx) CU_TRY(cudaMemcpyAsync(dx, p->d_dx, p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  remember_x(p, x_new, nid);
  if (pri_res_norm) *pri_res_norm = std::sqrt(p->h_scal[SC_PRI2]);
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// whole loop on the device: optim_loop! (iterate.jl:100-266), full batch
// ------------------------------------------------------------------------------------------------
extern "C" int scs_solve(scs_problem* p, const double* x0, const double* x_star, int64_t max_epoch, double x_tol,
                         double f_tol, double* x_out, double* obj, double* fval, double* pri_res_norm, double* rel,
                         double* objrel, int64_t* n_hist, int64_t* epochs_out) {
  SCS_TRY(check_ready(p));
  if (!p->has_sm) return fail(SCS_STATE_ERROR, "scs_set_smoother has not been called");
  if (!p->has_method) return fail(SCS_STATE_ERROR, "scs_set_method has not been called");
  if (!x0 || !x_out || !n_hist || !epochs_out) return fail(SCS_INVALID_ARG, "NULL argument");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  const int m = (int)p->m;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<double> zeros(m, 0.0);
  const double* xs = x_star ? x_star : zeros.data();
  CU_TRY(cudaMemcpyAsync(p->d_xstar, xs, m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double nxstar = 0;
  for (int i = 0; i < m; ++i) nxstar += xs[i] * xs[i];
  nxstar = std::sqrt(nxstar);
  const bool gl = p->reg.kind == SCS_REG_GL;
  int64_t nh = 0, epochs = 0;
  auto push = [&](double o, double f, double pr, double re, double fr) {
    if (obj) obj[nh] = o;
    if (fval) fval[nh] = f;
    if (pri_res_norm) pri_res_norm[nh] = pr;
    if (rel) rel[nh] = re;
    if (objrel) objrel[nh] = fr;
    ++nh;
  };
  // fetch objective pieces of x (loss sum, reg, ‖x−x*‖², ‖x‖²) -> host
  auto objective_at = [&](XRef v, double* f_out, double* reg_out, double* err2, double* nx2) -> int {
    SCS_TRY(objective_device(p, v));
    LAUNCH(c, k_bb, 1, kVecThreads, 0, v.d, p->d_xstar, v.d, p->d_xstar, m, p->d_scal);  // SC_GG = ‖v−x*‖²
    LAUNCH(c, k_dot, 1, kVecThreads, 0, v.d, v.d, m, p->d_scal, (int)SC_NX2);
    CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(p->h_scal + SC_COUNT, p->d_gl + p->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SCS_TRY(ctx_sync(c));
    *f_out = fval_from_sum(p, p->h_scal[SC_COUNT]);
    *reg_out = p->h_scal[SC_REGX];
    *err2 = p->h_scal[SC_GG];
    *nx2 = p->h_scal[SC_NX2];
    if (p->test) {  // ftest(x) = model.f(Atest, ytest, x)
      scs_problem* q = p->test;
      CU_TRY(cudaMemcpyAsync(q->vx[0], v.d, m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      XRef tv{q->vx[0], q->next_id++};
      SCS_TRY(ensure_loss(q, tv, SCS_WEIGHTS_NEWTON));
      CU_TRY(cudaMemcpyAsync(q->h_scal, q->d_gl + q->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      SCS_TRY(ctx_sync(c));
      p->test_hist.push_back(fval_from_sum(q, q->h_scal[0]));
    }
    return SCS_OK;
  };
  p->test_hist.clear();
  if (p->test && (p->test->m != p->m || p->test->ctx != p->ctx))
    return fail(SCS_INVALID_ARG, "the test problem must live in the same context and have the same number of columns");
  auto rel_err = [&](double err2) {  // iterate.jl:192-197
    if (gl) return err2 / (double)m;
    return std::max(std::sqrt(err2) / std::max(nxstar, 1.0), x_tol);
  };
  double obj_star;
  {
    XRef s{p->d_xstar, p->next_id++};
    double f, r, e2, n2;
    SCS_TRY(objective_at(s, &f, &r, &e2, &n2));
    obj_star = f + r;  // iterate.jl:179
    p->test_hist.clear();  // obj_star is not a history entry
  }
  auto frel = [&](double o) {  // iterate.jl:200 (NaN-propagating max)
    const double v = std::fabs(o - obj_star) / std::fabs(obj_star);
    return (v != v) ? v : std::max(v, f_tol);
  };
  int cur = 0, prv = 1, nxt = 2;
  CU_TRY(cudaMemcpyAsync(p->vx[cur], x0, m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(p->vx[prv], x0, m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  uint64_t id_cur = p->next_id++, id_prv = id_cur;
  SCS_TRY(scs_method_init(p));
  // mini-batches (iterate.jl:139-145,204): the objective always sees the whole shard, step! one batch at a time
  const bool batched = !p->batch_off.empty();
  const int64_t iend = batched ? (int64_t)p->batch_off.size() - 1 : 1;
  double pri = nan, f_rel_error = 0;
  for (int64_t epoch_t = 1; epoch_t <= max_epoch; ++epoch_t) {
    double f, r, e2, nx2;
    if (batched) SCS_TRY(set_window(p, 0, p->n, p->batch_rows_global.back()));
    {
      XRef xc{p->vx[cur], id_cur};
      SCS_TRY(objective_at(xc, &f, &r, &e2, &nx2));
    }
    double o = f + r;
    f_rel_error = frel(o);
    push(o, f, pri, rel_err(e2), f_rel_error);
    double diff = 0, nx = 0;
    for (int64_t i = 1; i <= iend; ++i) {
      XRef xc{p->vx[cur], id_cur}, xp{p->vx[prv], id_prv};
      if (epoch_t == max_epoch && i == iend) {  // iterate.jl:219-231: the current x once more
        if (batched && i > 1) {
          SCS_TRY(set_window(p, 0, p->n, p->batch_rows_global.back()));
          SCS_TRY(objective_at(xc, &f, &r, &e2, &nx2));
          o = f + r;
        } else if (p->test && !p->test_hist.empty()) {
          p->test_hist.push_back(p->test_hist.back());  // same x as the entry just recorded
        }
        f_rel_error = frel(o);
        push(o, f, pri, rel_err(e2), f_rel_error);
      }
      if (batched) SCS_TRY(set_window(p, p->batch_off[i - 1], p->batch_off[i], p->batch_rows_global[i - 1]));
      const uint64_t id_new = p->next_id++;
      SCS_TRY(step_device(p, xc, xp, epoch_t, p->vx[nxt], id_new, p->d_xstar));
      CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      SCS_TRY(ctx_sync(c));
      pri = std::sqrt(p->h_scal[SC_PRI2]);
      diff = std::sqrt(p->h_scal[SC_DIFF2]);
      nx = std::sqrt(p->h_scal[SC_NX2]);  // ‖x‖ of the step's input (k_pre)
      const bool stop = diff < x_tol * std::max(nx, 1.0) || f_rel_error <= f_tol || pri < x_tol;  // :234
      XRef xn{p->vx[nxt], id_new};
      if (stop && epoch_t != max_epoch) {  // :235-247
        double f2, r2, e22, n22;
        if (batched) SCS_TRY(set_window(p, 0, p->n, p->batch_rows_global.back()));
        SCS_TRY(objective_at(xn, &f2, &r2, &e22, &n22));
        const double o2 = f2 + r2;
        f_rel_error = frel(o2);
        push(o2, f2, pri, rel_err(e22), f_rel_error);
      }
      // x_prev = x; x = x_new
      const int old_prv = prv;
      prv = cur;
      id_prv = id_cur;
      cur = nxt;
      id_cur = id_new;
      nxt = old_prv;
      if (stop) {
        epochs += 1;
        break;
      }
    }
    // :257 — ‖x − x_prev‖ is the ‖x⁺ − x‖ of the last step taken
    if (diff < x_tol * std::max(nx, 1.0) || f_rel_error <= f_tol || pri < x_tol) break;
    epochs += 1;
  }
  if (batched) SCS_TRY(set_window(p, 0, p->n, p->batch_rows_global.back()));
  CU_TRY(cudaMemcpyAsync(x_out, p->vx[cur], m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  *n_hist = nh;
  *epochs_out = epochs;
  return SCS_OK;
}

// ------------------------------------------------------------------------------------------------
// exported: component entry points
// ------------------------------------------------------------------------------------------------
extern "C" int scs_loss_eval(scs_problem* p, const double* x, int weight_kind, double* fval, double* grad, double* z,
                             double* r, double* w) {
  if (!p || !x) return fail(SCS_INVALID_ARG, "NULL argument");
  if (weight_kind < 0 || weight_kind > 1) return fail(SCS_INVALID_ARG, "unknown weight_kind");
  if (weight_kind == SCS_WEIGHTS_GGN && p->loss.kind == SCS_LOSS_QUADFORM)
    return fail(SCS_UNSUPPORTED, "quadform loss has no out_fn / GGN weights");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  XRef xr;
  SCS_TRY(intern_x(p, x, p->vx[0], &xr));
  if (grad)
    SCS_TRY(ensure_grad(p, xr, weight_kind));
  else
    SCS_TRY(ensure_loss(p, xr, weight_kind));
  CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_gl + p->m, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (grad) CU_TRY(cudaMemcpyAsync(grad, p->d_gl, p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (z) CU_TRY(cudaMemcpyAsync(z, p->dz, p->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (r) CU_TRY(cudaMemcpyAsync(r, p->dr, p->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (w) CU_TRY(cudaMemcpyAsync(w, p->dw, p->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  if (fval) *fval = fval_from_sum(p, p->h_scal[0]);
  return SCS_OK;
}

extern "C" int scs_gram(scs_problem* p, const double* x, int weight_kind, double* G) {
  if (!p || !x || !G) return fail(SCS_INVALID_ARG, "NULL argument");
  if (weight_kind < 0 || weight_kind > 1) return fail(SCS_INVALID_ARG, "unknown weight_kind");
  if (weight_kind == SCS_WEIGHTS_GGN && p->loss.kind == SCS_LOSS_QUADFORM)
    return fail(SCS_UNSUPPORTED, "quadform loss has no out_fn / GGN weights");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  XRef xr;
  SCS_TRY(intern_x(p, x, p->vx[0], &xr));
  SCS_TRY(ensure_forward(p, xr, weight_kind));
  SCS_TRY(run_gram(p, xr));
  CU_TRY(cudaMemcpyAsync(G, p->d_G, (size_t)p->m * p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  return SCS_OK;
}

extern "C" int scs_linear_solve(scs_ctx* c, const double* M, const double* b, int64_t m64, double* d,
                                int* used_fallback) {
  if (!c || !M || !b || !d) return fail(SCS_INVALID_ARG, "NULL argument");
  if (m64 < 1 || m64 > 65536) return fail(SCS_INVALID_ARG, "m out of range");
  CU_TRY(cudaSetDevice(c->device));
  const int m = (int)m64;
  double *dM = nullptr, *dS = nullptr, *dL = nullptr, *db = nullptr, *dt = nullptr, *dd = nullptr;
  int* di = nullptr;
  const int nblk = (m + kNB - 1) / kNB;
  int rc = SCS_OK;
  auto cleanup = [&]() {
    dfree(dM), dfree(dS), dfree(dL), dfree(db), dfree(dt), dfree(dd), dfree(di);
  };
  if ((rc = dalloc(&dM, (size_t)m * m)) || (rc = dalloc(&dS, (size_t)m * m)) ||
      (rc = dalloc(&dL, (size_t)nblk * kNB * kNB + round_up(m, 16) + 16 + nblk)) || (rc = dalloc(&db, m)) || (rc = dalloc(&dt, m)) ||
      (rc = dalloc(&dd, m))) {
    cleanup();
    return rc;
  }
  if (cudaMalloc((void**)&di, sizeof(int)) != cudaSuccess) {
    cleanup();
    return fail(SCS_OOM, "cudaMalloc failed");
  }
  cudaMemcpyAsync(dM, M, (size_t)m * m * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(db, b, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  int fb = 0;
  rc = run_solve(c, dM, dS, dL, di, db, dt, dd, m, &fb, false);
  if (rc == SCS_OK) {
    cudaMemcpyAsync(d, dd, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    rc = ctx_sync(c);
  }
  if (used_fallback) *used_fallback = fb;
  cleanup();
  return rc;
}

extern "C" int scs_smoother_eval(scs_problem* p, const double* x, double* gr, double* hr) {
  if (!p || !x) return fail(SCS_INVALID_ARG, "NULL argument");
  if (!p->has_sm) return fail(SCS_STATE_ERROR, "scs_set_smoother has not been called");
  if ((p->sm.kind == SCS_SMOOTH_PHUBER_GL || p->sm.kind == SCS_SMOOTH_OSBA_GL) && !p->sm.cdiag)
    return fail(SCS_STATE_ERROR, "GL smoothers need scs_set_regularizer(gl) first (they read model.P)");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(p->d_trial, x, p->m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_smoother, 1, kVecThreads, 0, p->sm, p->d_trial, (int)p->m, p->d_t1, p->d_t2);
  if (gr) CU_TRY(cudaMemcpyAsync(gr, p->d_t1, p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (hr) CU_TRY(cudaMemcpyAsync(hr, p->d_t2, p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return ctx_sync(c);
}

extern "C" int scs_prox(scs_problem* p, const double* u, const double* hr, double ss, double* out) {
  SCS_TRY(check_ready(p));
  if (!u || !hr || !out) return fail(SCS_INVALID_ARG, "NULL argument");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(p->d_trial, u, p->m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaMemcpyAsync(p->d_t2, hr, p->m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_prox, 1, kVecThreads, 0, p->reg, ss, p->d_t2, p->d_trial, (int)p->m);
  CU_TRY(cudaMemcpyAsync(out, p->d_trial, p->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return ctx_sync(c);
}

extern "C" int scs_reg_value(scs_problem* p, const double* x, double* out) {
  SCS_TRY(check_ready(p));
  if (!x || !out) return fail(SCS_INVALID_ARG, "NULL argument");
  scs_ctx* c = p->ctx;
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(p->d_trial, x, p->m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_reg, 1, kVecThreads, 0, p->reg, p->d_trial, (int)p->m, p->d_scal + SC_REGX);
  CU_TRY(cudaMemcpyAsync(p->h_scal, p->d_scal + SC_REGX, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SCS_TRY(ctx_sync(c));
  *out = p->h_scal[0];
  return SCS_OK;
}
