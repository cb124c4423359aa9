// K1 — streaming passes over the device-resident row shard of A (column-major, fp64, leading dim ldd).
//
//   k_forward : z = A x fused with the per-row loss pieces (loss term, adjoint weight r, Gram weight w)
//               replaces  model.f(A,y,x) (iterate.jl:168,189), out_fn / jac_yx / grad_fy / hess_fy
//               (prox-GGN-SCORE.jl:44-56) and the z-dependent half of gradient/hessian(f,x)
//               (prox-N-SCORE.jl:56-64).
//   k_adjoint : g = A' r     replaces Jt*residual (prox-GGN-SCORE.jl:130) and the A' half of gradient(f,x).
//
// Both are HBM-bound: 8*n*m algorithmic bytes per pass.  Data layout: thread t of a CTA owns the two
// adjacent rows 2t,2t+1 of its row block, so every warp-level load instruction covers 512 contiguous
// bytes of one column and each CTA-level column access 4 KB; loads are 128-bit, L1-bypassing.
// Reductions use fixed trees (no atomics): results are bit-reproducible run to run.
#pragma once
#include "common.cuh"

namespace scs {

struct LossParams {
  int kind;         // scs_loss_kind
  int label_mode;   // scs_label_mode
  int weight_kind;  // scs_weight_kind
  double p;         // scale (logistic) or denominator (least squares)
};

// Per-row loss pieces.  Formulas follow oracle/scs_oracle.py (which cites README.md:113,135-139,212-214,
// 233-239 and test/test_algs.jl:9-11); evaluated literally so overflow behaviour matches too.
SCS_DEVINL void loss_row(const LossParams& lp, double z, double y, double& term, double& r, double& w) {
  if (lp.kind == 0) {  // logistic
    const double e = exp(-y * z);
    term = log(1.0 + e);
    if (lp.weight_kind == 0) {  // d/dz and d2/dz2 of p*log(1+exp(-y z))
      r = lp.p * (-y) * (e / (1.0 + e));
      w = lp.p * (y * y) * (e / ((1.0 + e) * (1.0 + e)));
    } else {  // GGN: yhat = 1/(1+exp(-z)), cross-entropy f(y,yhat)
      const double yc = lp.label_mode == 0 ? y : (y + 1.0) / 2.0;
      const double e2 = exp(-z);
      const double yhat = 1.0 / (1.0 + e2);
      const double s = (yhat / (1.0 + e2)) * e2;
      const double om = 1.0 - yhat;
      const double res = -lp.p * (yc / yhat - (1.0 - yc) / om);
      const double q = lp.p * (yc / (yhat * yhat) + (1.0 - yc) / (om * om));
      if (lp.weight_kind == 2) {  // wide branch (prox-GGN-SCORE.jl:124-127) wants the pieces, not the products
        r = res;
        w = q;
        term = s;  // NB: the caller of kind 2 reads the Jacobian scale from `term`
      } else {
        r = s * res;
        w = (s * s) * q;
      }
    }
  } else if (lp.kind == 1) {  // least squares: 0.5*sum((z-y)^2)/p
    const double d = z - y;
    term = lp.weight_kind == 2 ? 1.0 : d * d;  // kind 2: Jacobian scale (out_fn = A x)
    r = d / lp.p;
    w = 1.0 / lp.p;
  } else {  // quadform: the pass only produces z
    term = 0.0;
    r = 0.0;
    w = 0.0;
  }
}

constexpr int kFwdThreads = 256;
constexpr int kFwdRows = 2 * kFwdThreads;  // rows per CTA
constexpr int kXChunk = 2048;              // columns of x staged in shared memory at a time (16 KB)

// z = A x, then loss pieces.  Grid: ceil(nproc / 512) CTAs of 256 threads.
// Rows [0, nproc) of the (row-window) base pointers are processed (nproc even); rows outside [row_lo, row_hi) are
// padding or belong to other mini-batches: their z, r, w are written as zeros and they add nothing to the loss.
// loss_part[blockIdx.x] = sum of this CTA's loss terms.
template <int UNR>
__global__ void __launch_bounds__(kFwdThreads)
k_forward(const double* __restrict__ A, int64_t ldd, int64_t nproc, int64_t row_lo, int64_t row_hi, int m,
          const double* __restrict__ x, const double* __restrict__ y, LossParams lp, double* __restrict__ z_out,
          double* __restrict__ r_out, double* __restrict__ w_out, double* __restrict__ loss_part) {
  __shared__ double xs[kXChunk];
  __shared__ double red[32];
  const int64_t i0 = ((int64_t)blockIdx.x * kFwdThreads + threadIdx.x) * 2;
  const bool active = i0 < nproc;
  const double* Ap = A + (active ? i0 : 0);
  // four independent accumulators per row: shorter dependency chains, tighter rounding than one chain
  double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
  for (int jc = 0; jc < m; jc += kXChunk) {
    const int cn = min(kXChunk, m - jc);
    __syncthreads();
    for (int t = threadIdx.x; t < cn; t += kFwdThreads) xs[t] = x[jc + t];
    __syncthreads();
    if (active) {
      const double* p = Ap + (int64_t)jc * ldd;
      int j = 0;
      for (; j + UNR <= cn; j += UNR) {
        double2 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = ldg_stream2(p + (int64_t)(j + u) * ldd);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const double xv = xs[j + u];
          a0[u & 3] = fma(v[u].x, xv, a0[u & 3]);
          a1[u & 3] = fma(v[u].y, xv, a1[u & 3]);
        }
      }
      for (; j < cn; ++j) {
        const double2 v = ldg_stream2(p + (int64_t)j * ldd);
        const double xv = xs[j];
        a0[0] = fma(v.x, xv, a0[0]);
        a1[0] = fma(v.y, xv, a1[0]);
      }
    }
  }
  double part = 0.0;
  if (active) {
    double z0 = (a0[0] + a0[1]) + (a0[2] + a0[3]);
    double z1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
    const double2 yy = *reinterpret_cast<const double2*>(y + i0);
    double t0, r0, w0, t1, r1, w1;
    loss_row(lp, z0, yy.x, t0, r0, w0);
    loss_row(lp, z1, yy.y, t1, r1, w1);
    if (i0 < row_lo || i0 >= row_hi) t0 = r0 = w0 = z0 = 0.0;
    if (i0 + 1 < row_lo || i0 + 1 >= row_hi) t1 = r1 = w1 = z1 = 0.0;
    part = t0 + t1;
    if (lp.weight_kind == 2) {  // GGN parts: z_out carries the Jacobian scale s, r_out = residual, w_out = Q_ii
      z0 = t0;
      z1 = t1;
      part = 0.0;
    }
    if (z_out) *reinterpret_cast<double2*>(z_out + i0) = make_double2(z0, z1);
    if (r_out) *reinterpret_cast<double2*>(r_out + i0) = make_double2(r0, r1);
    if (w_out) *reinterpret_cast<double2*>(w_out + i0) = make_double2(w0, w1);
  }
  const double tot = block_sum<kFwdThreads>(part, red);
  if (threadIdx.x == 0) loss_part[blockIdx.x] = tot;
}

// Deterministic sum of `cnt` partials into out[0] (single CTA).
__global__ void __launch_bounds__(kVecThreads) k_sum_partials(const double* __restrict__ part, int64_t cnt,
                                                              double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += kVecThreads) s += part[i];
  s = block_sum<kVecThreads>(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// g = A' r over rows [0, nproc) of the base pointers (ldd = column stride).  CTA b owns the 64*K rows [b*64K, (b+1)*64K); warp w sweeps columns w, w+8, ...; lane l keeps
// r for rows base + 2l + 64k (+1) in registers.  part[b*m + j] = this row block's contribution to g_j.
constexpr int kAdjThreads = 256;
constexpr int kAdjWarps = kAdjThreads / 32;

template <int K>
__global__ void __launch_bounds__(kAdjThreads)
k_adjoint(const double* __restrict__ A, int64_t ldd, int64_t nproc, int m, const double* __restrict__ r,
          double* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * (64 * K) + 2 * lane;
  // rows are valid while < nproc (even; padding rows hold zeros in A, rows outside the active window zeros in r)
  int nk = 0;
  if (base < nproc) {
    const int64_t q = (nproc - base + 63) / 64;
    nk = q < (int64_t)K ? (int)q : K;
  }
  double2 rr[K];
#pragma unroll
  for (int k = 0; k < K; ++k)
    rr[k] = k < nk ? *reinterpret_cast<const double2*>(r + base + 64 * k) : make_double2(0.0, 0.0);
  const bool full = __all_sync(0xffffffffu, nk == K);
  double* out = part + (int64_t)blockIdx.x * m;
  const double* Ab = A + (base < nproc ? base : 0);
  if (full) {
    int j = warp;
    for (; j + kAdjWarps < m; j += 2 * kAdjWarps) {  // two columns in flight per warp
      const double* p0 = Ab + (int64_t)j * ldd;
      const double* p1 = p0 + (int64_t)kAdjWarps * ldd;
      double2 v0[K], v1[K];
#pragma unroll
      for (int k = 0; k < K; ++k) v0[k] = ldg_stream2(p0 + 64 * k);
#pragma unroll
      for (int k = 0; k < K; ++k) v1[k] = ldg_stream2(p1 + 64 * k);
      double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        s0 = fma(v0[k].x, rr[k].x, s0);
        s1 = fma(v0[k].y, rr[k].y, s1);
        t0 = fma(v1[k].x, rr[k].x, t0);
        t1 = fma(v1[k].y, rr[k].y, t1);
      }
      const double a = warp_sum(s0 + s1), b = warp_sum(t0 + t1);
      if (lane == 0) {
        out[j] = a;
        out[j + kAdjWarps] = b;
      }
    }
    for (; j < m; j += kAdjWarps) {
      const double* p0 = Ab + (int64_t)j * ldd;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double2 v = ldg_stream2(p0 + 64 * k);
        s0 = fma(v.x, rr[k].x, s0);
        s1 = fma(v.y, rr[k].y, s1);
      }
      const double a = warp_sum(s0 + s1);
      if (lane == 0) out[j] = a;
    }
  } else {  // ragged last row block
    for (int j = warp; j < m; j += kAdjWarps) {
      const double* p0 = Ab + (int64_t)j * ldd;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (k < nk) {
          const double2 v = ldg_stream2(p0 + 64 * k);
          s0 = fma(v.x, rr[k].x, s0);
          s1 = fma(v.y, rr[k].y, s1);
        }
      }
      const double a = warp_sum(s0 + s1);
      if (lane == 0) out[j] = a;
    }
  }
}

// out[j] = sum_b part[b*m + j] in fixed order.  Block = 32 columns x 8 slices.
__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ part, int64_t nblk, int m,
                                                double* __restrict__ out) {
  __shared__ double sh[8][33];
  const int c = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  double acc = 0.0;
  if (j < m)
    for (int64_t b = s; b < nblk; b += 8) acc += part[b * m + j];
  sh[s][c] = acc;
  __syncthreads();
  if (s == 0 && j < m) {
    double t = sh[0][c];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += sh[q][c];
    out[j] = t;
  }
}

// ---- quadform helpers (test/test_algs.jl:90: f = 1/2 x'(Ax) + y'x, A square m x m) ----------------
// scal[0] = 0.5 * x.z + y.x
__global__ void __launch_bounds__(kVecThreads) k_quadform_value(const double* __restrict__ x,
                                                                const double* __restrict__ z,
                                                                const double* __restrict__ y, int m,
                                                                double* __restrict__ out) {
  __shared__ double red[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    a += x[i] * z[i];
    b += y[i] * x[i];
  }
  a = block_sum<kVecThreads>(a, red);
  b = block_sum<kVecThreads>(b, red);
  if (threadIdx.x == 0) out[0] = 0.5 * a + b;
}
// g = 0.5*(z + At x) + y
__global__ void k_quadform_grad(const double* __restrict__ z, const double* __restrict__ atx,
                                const double* __restrict__ y, int m, double* __restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) g[i] = 0.5 * (z[i] + atx[i]) + y[i];
}
// H = 0.5*(A + A')  (full m x m, column-major ld = m) from A with leading dimension ldd
__global__ void k_quadform_hess(const double* __restrict__ A, int64_t ldd, int m, double* __restrict__ H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < m) H[(int64_t)j * m + i] = 0.5 * (A[(int64_t)j * ldd + i] + A[(int64_t)i * ldd + j]);
}

}  // namespace scs
