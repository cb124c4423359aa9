// K3 — on-device solve of the m x m Newton / Gauss-Newton system.
// Replaces  (H + λHr) \ ∇q  (prox-N-SCORE.jl:70, dense LU in the reference) and  qr(JQJ) \ Je
// (prox-GGN-SCORE.jl:131).  The system matrix is symmetric; it is SPD whenever the Gram weights are
// non-negative, so the fast path is a blocked Cholesky (NB = 64, right-looking, inverted diagonal blocks so
// that the panel solve and both triangular solves are GEMM/GEMV-shaped).  If a pivot is not positive (possible
// with the README's +-1-label cross-entropy) the flag is raised and the host runs the
// partial-pivoting LU below on the saved copy — still on the device, never on the CPU.
//
// Matrices are column-major with leading dimension ld; only the lower triangle is read by the Cholesky.
#pragma once
#include "common.cuh"

namespace scs {

constexpr int kNB = 64;

// ---- blocked Cholesky, NB = 64 --------------------------------------------------------------------------------
// Step k = one k_panel launch + one k_syrk_update launch.
//
// k_panel: every CTA factors the 64x64 diagonal block itself (redundant but parallel, so the factor never has to
// travel), CTA 0 stores it, CTA b >= 1 then solves X L11' = A21 for its 64 rows by column substitution.  Thread
// layout: tid = 4*row + slice; a thread keeps the entries (row, 4k+slice), k = 0..15, of its row in registers, so
// the rank-1 updates are 16 independent FMAs, values move inside a row with quad shuffles and between rows through a
// double-buffered 64-entry column in shared memory (one barrier per column).
constexpr int kLS = kNB + 1;
SCS_DEVINL double quad_bcast(double v, int src_slice) {
  return __shfl_sync(0xffffffffu, v, (threadIdx.x & 28) | src_slice);
}
__global__ void __launch_bounds__(256) k_panel(double* __restrict__ M, int64_t ld, int m, int k0,
                                               double* __restrict__ rdiag_g, int* __restrict__ info) {
  __shared__ double Ls[kNB * kLS];   // factored diagonal block, Ls[r][c]
  __shared__ double colraw[2][kNB];  // column j before scaling
  __shared__ double rds[kNB];        // 1 / L_jj
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = 4 * k + q;
    a[k] = (r < nb && c < nb && c <= r) ? M[(int64_t)(k0 + c) * ld + k0 + r] : ((r == c) ? 1.0 : 0.0);
  }
#pragma unroll
  for (int j = 0; j < kNB; ++j) {
    const int kq = j & 3, kk = j >> 2;
    if (q == kq) colraw[j & 1][r] = a[kk];
    __syncthreads();
    const double piv = colraw[j & 1][j];
    if (blockIdx.x == 0 && tid == 0 && j < nb && !(piv > 0.0) && info[0] == 0) info[0] = k0 + j + 1;
    const double rd = rsqrt(piv);
    const double lij = (r > j) ? colraw[j & 1][r] * rd : 0.0;
    if (q == kq) {
      if (r > j) a[kk] = lij;
      if (r == j) {
        a[kk] = piv * rd;
        rds[j] = rd;
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = 4 * k + q;
      if (4 * k + 3 > j) {  // compile-time prune: no column of this register group is right of j otherwise
        if (c > j && c <= r) a[k] -= lij * (colraw[j & 1][c] * rd);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = 4 * k + q;
    Ls[r * kLS + c] = (c <= r) ? a[k] : 0.0;
    if (blockIdx.x == 0 && r < nb && c < nb && c <= r) M[(int64_t)(k0 + c) * ld + k0 + r] = a[k];
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < nb) rdiag_g[k0 + tid] = rds[tid];
    return;
  }
  // ---- rows below: X L11' = A21  ->  x_:c = (a_:c - sum_{p<c} x_:p L[c][p]) / L[c][c], right-looking over c
  const int row = k0 + nb + (blockIdx.x - 1) * kNB + r;
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = 4 * k + q;
    x[k] = (row < m && c < nb) ? M[(int64_t)(k0 + c) * ld + row] : 0.0;
  }
#pragma unroll
  for (int c = 0; c < kNB; ++c) {
    const int kq = c & 3, kk = c >> 2;
    const double xc = quad_bcast(x[kk] * rds[c], kq);  // only slice kq's value is used
    if (q == kq) x[kk] = xc;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int cp = 4 * k + q;
      if (4 * k + 3 > c) {
        if (cp > c) x[k] -= xc * Ls[cp * kLS + c];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = 4 * k + q;
    if (row < m && c < nb) M[(int64_t)(k0 + c) * ld + row] = x[k];
  }
}

// 64x64x64 tile product on the FP64 tensor pipe: acc += As^T-tile * Bs-tile with both operands staged in shared
// memory as [p][row] (row stride 68 doubles: conflict-free 64-bit fragment loads).  4 warps, 32x32 each.
constexpr int kTS = 68;
constexpr int kTileSmem = 2 * kNB * kTS * 8;
SCS_DEVINL void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
SCS_DEVINL void tile_mma_64(const double* As, const double* Bs, int wm, int wn, int g, int t, double (&acc)[4][4][2]) {
#pragma unroll 4
  for (int p0 = 0; p0 < kNB; p0 += 4) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = As[(p0 + t) * kTS + wm + 8 * i + g];
      b[i] = Bs[(p0 + t) * kTS + wn + 8 * i + g];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}
// dst[p][i] = M[row0 + i, k0 + p] for i < 64, p < nb (zero elsewhere): 128 threads, 128-bit loads along the rows,
// eight loads in flight per thread before the first shared-memory store.
SCS_DEVINL void load_tile_T(double* dst, const double* __restrict__ M, int64_t ld, int m, int k0, int row0, int nb,
                            int tid) {
  const bool vec_ok = ((row0 | (int)(ld & 1)) & 1) == 0 && row0 + kNB <= m;
  const int i = (tid & 31) * 2, pb = tid >> 5;  // this thread: rows i, i+1 of columns pb, pb+4, ...
  if (vec_ok) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int p = pb + 4 * (8 * h + u);
        v[u] = p < nb ? *reinterpret_cast<const double2*>(M + (int64_t)(k0 + p) * ld + row0 + i) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) *reinterpret_cast<double2*>(dst + (pb + 4 * (8 * h + u)) * kTS + i) = v[u];
    }
  } else {
    for (int u = 0; u < 16; ++u) {
      const int p = pb + 4 * u;
      double2 v = make_double2(0.0, 0.0);
      if (p < nb) {
        const double* src = M + (int64_t)(k0 + p) * ld + row0 + i;
        if (row0 + i < m) v.x = src[0];
        if (row0 + i + 1 < m) v.y = src[1];
      }
      *reinterpret_cast<double2*>(dst + p * kTS + i) = v;
    }
  }
}

// Trailing update A22 -= L21 * L21'  (lower 64x64 tiles only).  grid = (#tiles in the lower triangle), 128 threads.
__global__ void __launch_bounds__(128) k_syrk_update(double* __restrict__ M, int64_t ld, int m, int k0) {
  extern __shared__ double tile_sh[];
  double* As = tile_sh;
  double* Bs = tile_sh + kNB * kTS;
  const int nb = min(kNB, m - k0);
  const int base = k0 + nb;
  int ti, tj;
  {
    const int tt = blockIdx.x;
    int r = (int)((sqrt(8.0 * (double)tt + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= tt) ++r;
    while (r * (r + 1) / 2 > tt) --r;
    ti = r;
    tj = tt - r * (r + 1) / 2;
  }
  const int r0 = base + ti * kNB, c0 = base + tj * kNB;
  const int tid = threadIdx.x;
  load_tile_T(As, M, ld, m, k0, r0, nb, tid);
  if (ti == tj) {
    Bs = As;  // diagonal tile: both operands are the same panel rows
  } else {
    load_tile_T(Bs, M, ld, m, k0, c0, nb, tid);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  double acc[4][4][2] = {};
  tile_mma_64(As, Bs, wm, wn, g, t, acc);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = r0 + wm + 8 * i + g, col = c0 + wn + 8 * j + 2 * t + h;
        if (row < m && col < m && row >= col) M[(int64_t)col * ld + row] -= acc[i][j][h];
      }
}

// Loads the factored diagonal block k into shared memory Ls[r][c] (lower part, zero above), 256 threads.
SCS_DEVINL void load_diag_block(double* Ls, const double* __restrict__ M, int64_t ld, int k0, int nb, int tid) {
  const int r = tid & 63, cb = tid >> 6;
  double v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = cb + 4 * u;
    v[u] = (r < nb && c < nb && c <= r) ? M[(int64_t)(k0 + c) * ld + k0 + r] : 0.0;
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) Ls[r * kLS + cb + 4 * u] = v[u];
}

// Forward substitution step for block k (L y = b): every CTA solves L_kk y_k = b_k itself (one warp, 64 sequential
// columns), CTA 0 stores y_k, and CTA c updates its 256 rows below: b_i -= sum_c L[i, k0+c] y_c.
__global__ void __launch_bounds__(256) k_fwd_step(const double* __restrict__ M, int64_t ld, int m, int k0,
                                                  const double* __restrict__ rdiag_g, double* __restrict__ b,
                                                  double* __restrict__ y) {
  __shared__ double Ls[kNB * kLS];
  __shared__ double bk[kNB], yk[kNB], rd[kNB];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x;
  load_diag_block(Ls, M, ld, k0, nb, tid);
  if (tid < kNB) {
    bk[tid] = tid < nb ? b[k0 + tid] : 0.0;
    rd[tid] = tid < nb ? rdiag_g[k0 + tid] : 0.0;
  }
  // prefetch this thread's row of the panel while warp 0 runs the substitution
  const int row = k0 + nb + blockIdx.x * 256 + tid;
  __syncthreads();
  if (tid < 32) {
    for (int c = 0; c < nb; ++c) {
      const double yc = bk[c] * rd[c];
      if (tid == 0) yk[c] = yc;
      const int r0 = tid, r1 = tid + 32;
      if (r0 > c) bk[r0] -= Ls[r0 * kLS + c] * yc;
      if (r1 > c) bk[r1] -= Ls[r1 * kLS + c] * yc;
      __syncwarp();
    }
  }
  __syncthreads();
  if (row < m) {
    const double* rp = M + (int64_t)k0 * ld + row;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int c = 0;
    for (; c + 16 <= nb; c += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = rp[(int64_t)(c + u) * ld];
#pragma unroll
      for (int u = 0; u < 16; ++u) s[u & 3] = fma(v[u], yk[c + u], s[u & 3]);
    }
    for (; c < nb; ++c) s[0] = fma(rp[(int64_t)c * ld], yk[c], s[0]);
    b[row] -= (s[0] + s[1]) + (s[2] + s[3]);
  }
  if (blockIdx.x == 0 && tid < nb) y[k0 + tid] = yk[tid];
}

// Backward substitution step for block k (L' d = y), right-looking: solve L_kk' d_k = y_k, then for every column
// j < k0: y_j -= sum_{r in block k} L[r, j] d_r.
__global__ void __launch_bounds__(256) k_bwd_step(const double* __restrict__ M, int64_t ld, int m, int k0,
                                                  const double* __restrict__ rdiag_g, double* __restrict__ y,
                                                  double* __restrict__ d) {
  __shared__ double Ls[kNB * kLS];
  __shared__ double yk[kNB], dk[kNB], rd[kNB];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x;
  load_diag_block(Ls, M, ld, k0, nb, tid);
  if (tid < kNB) {
    yk[tid] = tid < nb ? y[k0 + tid] : 0.0;
    rd[tid] = tid < nb ? rdiag_g[k0 + tid] : 0.0;
    dk[tid] = 0.0;
  }
  __syncthreads();
  if (tid < 32) {
    for (int c = nb - 1; c >= 0; --c) {
      const double dc = yk[c] * rd[c];
      if (tid == 0) dk[c] = dc;
      const int r0 = tid, r1 = tid + 32;
      if (r0 < c) yk[r0] -= Ls[c * kLS + r0] * dc;  // (L')[r][c] = L[c][r]
      if (r1 < c) yk[r1] -= Ls[c * kLS + r1] * dc;
      __syncwarp();
    }
  }
  __syncthreads();
  const int col = blockIdx.x * 256 + tid;
  if (col < k0) {
    const double* colp = M + (int64_t)col * ld + k0;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int r = 0;
    for (; r + 16 <= nb; r += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = colp[r + u];
#pragma unroll
      for (int u = 0; u < 16; ++u) s[u & 3] = fma(v[u], dk[r + u], s[u & 3]);
    }
    for (; r < nb; ++r) s[0] = fma(colp[r], dk[r], s[0]);
    y[col] -= (s[0] + s[1]) + (s[2] + s[3]);
  }
  if (blockIdx.x == 0 && tid < nb) d[k0 + tid] = dk[tid];
}

// ---- pivoted LU fallback (unblocked, right-looking) -------------------------------------------
// Fill the upper triangle from the lower one.
__global__ void k_symmetrize(double* __restrict__ M, int64_t ld, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < m && i > j) M[(int64_t)i * ld + j] = M[(int64_t)j * ld + i];
}
// Column k: find the first row of max |M[i,k]|, i >= k; swap rows k,p in all columns and in b; scale column.
__global__ void __launch_bounds__(kVecThreads) k_lu_pivot(double* __restrict__ M, int64_t ld, int m, int k,
                                                          double* __restrict__ b, int* __restrict__ info) {
  __shared__ double vmax[32];
  __shared__ int imax[32];
  __shared__ int piv;
  double best = -1.0;
  int bi = k;
  for (int i = k + threadIdx.x; i < m; i += kVecThreads) {
    const double v = fabs(M[(int64_t)k * ld + i]);
    if (v > best) {
      best = v;
      bi = i;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    vmax[wid] = best;
    imax[wid] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bv = vmax[0];
    int bb = imax[0];
    for (int q = 1; q < kVecThreads / 32; ++q)
      if (vmax[q] > bv || (vmax[q] == bv && imax[q] < bb)) {
        bv = vmax[q];
        bb = imax[q];
      }
    piv = bb;
    if (!(bv > 0.0) && info[0] == 0) info[0] = k + 1;  // singular
  }
  __syncthreads();
  const int p = piv;
  if (p != k) {
    for (int j = threadIdx.x; j < m; j += kVecThreads) {
      const double t = M[(int64_t)j * ld + k];
      M[(int64_t)j * ld + k] = M[(int64_t)j * ld + p];
      M[(int64_t)j * ld + p] = t;
    }
    if (threadIdx.x == 0) {
      const double t = b[k];
      b[k] = b[p];
      b[p] = t;
    }
  }
  __syncthreads();
  const double d = M[(int64_t)k * ld + k];
  for (int i = k + 1 + threadIdx.x; i < m; i += kVecThreads) M[(int64_t)k * ld + i] /= d;
}
// Rank-1 update of the trailing block and of b:  M[i,j] -= M[i,k]*M[k,j];  b[i] -= M[i,k]*b[k]  (i,j > k)
__global__ void __launch_bounds__(256) k_lu_update(double* __restrict__ M, int64_t ld, int m, int k,
                                                   double* __restrict__ b) {
  const int i = k + 1 + blockIdx.x * 256 + threadIdx.x;
  const int j0 = k + 1 + blockIdx.y * 16;
  if (i >= m) return;
  const double l = M[(int64_t)k * ld + i];
  for (int j = j0; j < min(j0 + 16, m); ++j) M[(int64_t)j * ld + i] -= l * M[(int64_t)j * ld + k];
  if (blockIdx.y == 0) b[i] -= l * b[k];
}
// Back substitution U d = b (single CTA, column-oriented).
__global__ void __launch_bounds__(kVecThreads) k_lu_backsolve(const double* __restrict__ M, int64_t ld, int m,
                                                              double* __restrict__ b, double* __restrict__ d) {
  __shared__ double xk;
  for (int k = m - 1; k >= 0; --k) {
    if (threadIdx.x == 0) {
      xk = b[k] / M[(int64_t)k * ld + k];
      d[k] = xk;
    }
    __syncthreads();
    const double v = xk;
    for (int i = threadIdx.x; i < k; i += kVecThreads) b[i] -= M[(int64_t)k * ld + i] * v;
    __syncthreads();
  }
}

}  // namespace scs
