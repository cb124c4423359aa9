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

// Factor the diagonal block k (size nb <= 64) in place, and write inv(L_kk) (lower) to Linv (64x64, column-major
// ld 64).  One CTA of 256 threads.  info[0] = 1-based column of the first non-positive pivot (0 = ok).
__global__ void __launch_bounds__(256) k_potf2(double* __restrict__ M, int64_t ld, int m, int k0,
                                               double* __restrict__ Linv, int* __restrict__ info) {
  __shared__ double L[kNB][kNB + 1];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x;
  for (int e = tid; e < kNB * kNB; e += 256) {
    const int i = e % kNB, j = e / kNB;
    L[i][j] = (i < nb && j < nb && i >= j) ? M[(int64_t)(k0 + j) * ld + k0 + i] : 0.0;
    Linv[e] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double piv = L[j][j];
    if (tid == 0 && !(piv > 0.0) && info[0] == 0) info[0] = k0 + j + 1;
    __syncthreads();
    const double d = sqrt(piv);
    if (tid > j && tid < nb) L[tid][j] = L[tid][j] / d;
    if (tid == j) L[j][j] = d;
    __syncthreads();
    // trailing update of the block: L[i][c] -= L[i][j]*L[c][j] for j < c <= i < nb
    const int rem = nb - j - 1;
    for (int e = tid; e < rem * rem; e += 256) {
      const int i = j + 1 + e % rem, c = j + 1 + e / rem;
      if (i >= c) L[i][c] -= L[i][j] * L[c][j];
    }
    __syncthreads();
  }
  // inverse of the lower-triangular block, one column per thread (forward substitution); the column lives in
  // global memory (thread-private, so plain loads see the thread's own earlier stores)
  if (tid < nb) {
    const int c = tid;
    double* col = Linv + c * kNB;
    for (int i = c; i < nb; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int p = c; p < i; ++p) s -= L[i][p] * col[p];
      col[i] = s / L[i][i];
    }
  }
  for (int e = tid; e < kNB * kNB; e += 256) {
    const int i = e % kNB, j = e / kNB;
    if (i < nb && j < nb && i >= j) M[(int64_t)(k0 + j) * ld + k0 + i] = L[i][j];
  }
}

// 64x64 micro-kernel shared by the panel solve and the trailing update: acc[i][j] += sum_p As[tr+i][p]*Bs[tc+j][p]
// over a 32-wide slice of p staged in shared memory.
constexpr int kKC = 32;
SCS_DEVINL void mm_slice(const double (*As)[kKC + 1], const double (*Bs)[kKC + 1], int tr, int tc, int pn,
                         double (&acc)[4][4]) {
  for (int p = 0; p < pn; ++p) {
    double a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a[q] = As[tr + q][p];
      b[q] = Bs[tc + q][p];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// Panel: X = A21 * inv(L11)'  for rows below the diagonal block.  Each CTA handles 64 rows; 256 threads,
// each thread a 4x4 micro-tile of the 64x64 output.
__global__ void __launch_bounds__(256) k_trsm_panel(double* __restrict__ M, int64_t ld, int m, int k0,
                                                    const double* __restrict__ Linv) {
  __shared__ double As[kNB][kKC + 1];  // As[r][p] = A21[r][p0+p]
  __shared__ double Bs[kNB][kKC + 1];  // Bs[c][p] = Linv(c, p0+p)
  const int nb = min(kNB, m - k0);
  const int r0 = k0 + nb + blockIdx.x * kNB;
  const int tid = threadIdx.x;
  const int tr = (tid % 16) * 4, tc = (tid / 16) * 4;
  double acc[4][4] = {};
  for (int p0 = 0; p0 < nb; p0 += kKC) {
    __syncthreads();
    for (int e = tid; e < kNB * kKC; e += 256) {
      const int i = e % kNB, j = e / kNB;
      const int row = r0 + i, p = p0 + j;
      As[i][j] = (row < m && p < nb) ? M[(int64_t)(k0 + p) * ld + row] : 0.0;
      Bs[i][j] = p < kNB ? Linv[p * kNB + i] : 0.0;
    }
    __syncthreads();
    mm_slice(As, Bs, tr, tc, min(kKC, nb - p0), acc);
  }
  __syncthreads();  // every thread has finished reading the panel rows this CTA is about to overwrite
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = r0 + tr + i, col = tc + j;
      if (row < m && col < nb) M[(int64_t)(k0 + col) * ld + row] = acc[i][j];
    }
}

// Trailing update A22 -= L21 * L21'  (lower 64x64 tiles only).  grid = (#tiles in the lower triangle).
__global__ void __launch_bounds__(256) k_syrk_update(double* __restrict__ M, int64_t ld, int m, int k0) {
  __shared__ double As[kNB][kKC + 1];
  __shared__ double Bs[kNB][kKC + 1];
  const int nb = min(kNB, m - k0);
  const int base = k0 + nb;
  int ti, tj;
  {
    const int t = blockIdx.x;
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    while (r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - r * (r + 1) / 2;
  }
  const int r0 = base + ti * kNB, c0 = base + tj * kNB;
  const int tid = threadIdx.x;
  const int tr = (tid % 16) * 4, tc = (tid / 16) * 4;
  double acc[4][4] = {};
  for (int p0 = 0; p0 < nb; p0 += kKC) {
    __syncthreads();
    for (int e = tid; e < kNB * kKC; e += 256) {
      const int i = e % kNB, j = e / kNB;
      const int p = p0 + j;
      As[i][j] = (r0 + i < m && p < nb) ? M[(int64_t)(k0 + p) * ld + r0 + i] : 0.0;
      Bs[i][j] = (c0 + i < m && p < nb) ? M[(int64_t)(k0 + p) * ld + c0 + i] : 0.0;
    }
    __syncthreads();
    mm_slice(As, Bs, tr, tc, min(kKC, nb - p0), acc);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = r0 + tr + i, col = c0 + tc + j;
      if (row < m && col < m && row >= col) M[(int64_t)col * ld + row] -= acc[i][j];
    }
}

// Forward substitution step for block k (L y = b): every CTA recomputes y_k = inv(L_kk) b_k, CTA 0 stores it
// into y, and CTA c updates its 256 rows below: b_i -= sum_c L[i, k0+c] y_c.  (y is a separate vector so no CTA
// reads a slot another CTA overwrites.)
__global__ void __launch_bounds__(256) k_fwd_step(const double* __restrict__ M, int64_t ld, int m, int k0,
                                                   const double* __restrict__ Linv_k, double* __restrict__ b,
                                                   double* __restrict__ y) {
  __shared__ double yk[kNB];
  __shared__ double bk[kNB];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x;
  if (tid < kNB) bk[tid] = tid < nb ? b[k0 + tid] : 0.0;
  __syncthreads();
  if (tid < kNB) {
    double s = 0.0;
    for (int p = 0; p <= tid && p < nb; ++p) s = fma(Linv_k[p * kNB + tid], bk[p], s);
    yk[tid] = s;
  }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * 256 + tid;
  if (row < m) {
    double s = 0.0;
    for (int c = 0; c < nb; ++c) s = fma(M[(int64_t)(k0 + c) * ld + row], yk[c], s);
    b[row] -= s;
  }
  if (blockIdx.x == 0 && tid < nb) y[k0 + tid] = yk[tid];
}

// Backward substitution step for block k (L' d = y), right-looking: d_k = inv(L_kk)' y_k, then for every
// column j < k0: y_j -= sum_{r in block k} L[r, j] d_r.
__global__ void __launch_bounds__(256) k_bwd_step(const double* __restrict__ M, int64_t ld, int m, int k0,
                                                  const double* __restrict__ Linv_k, double* __restrict__ y,
                                                  double* __restrict__ d) {
  __shared__ double dk[kNB];
  __shared__ double yk[kNB];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x;
  if (tid < kNB) yk[tid] = tid < nb ? y[k0 + tid] : 0.0;
  __syncthreads();
  if (tid < kNB) {
    double s = 0.0;
    for (int p = tid; p < nb; ++p) s = fma(Linv_k[tid * kNB + p], yk[p], s);  // (Linv')[tid][p] = Linv[p][tid]
    dk[tid] = s;
  }
  __syncthreads();
  const int col = blockIdx.x * 256 + tid;
  if (col < k0) {
    double s = 0.0;
    const double* colp = M + (int64_t)col * ld + k0;
    for (int r = 0; r < nb; ++r) s = fma(colp[r], dk[r], s);
    y[col] -= s;
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
