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
// Grid: CTA 0 stores the factored diagonal block; CTAs 1..rb substitute 64 rows of the panel each; the last CTA
// (index 1 + rb) carries the right-hand side as one more row: y_k = L_kk^-1 b_k (forward substitution is folded into
// the factorisation; the matching update of b below the block is done by k_syrk_update's GEMV CTAs).
// The register window a[0..15] always starts at the current column group (it is rotated after every four columns), so
// the column loop is a compact runtime loop with static register indices.
__global__ void __launch_bounds__(256) k_panel(double* __restrict__ M, int64_t ld, int m, int k0,
                                               double* __restrict__ rdiag_g, int* __restrict__ info,
                                               double* __restrict__ bvec, double* __restrict__ yvec, int rb) {
  __shared__ double Ls[kNB * kLS];   // factored diagonal block, Ls[r][c]
  __shared__ double colraw[2][kNB];  // column j before scaling
  __shared__ double rds[kNB];        // 1 / L_jj
  __shared__ double bk[kNB];
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = 4 * k + q;
    a[k] = (r < nb && c < nb && c <= r) ? M[(int64_t)(k0 + c) * ld + k0 + r] : ((r == c) ? 1.0 : 0.0);
  }
  // Entries above the diagonal (c > r) are carried along but never consumed: the per-element triangle test is not
  // needed, only the uniform "window still inside the block" bound (k < 16 - kk).
#pragma unroll 1
  for (int kk = 0; kk < 16; ++kk) {
    const int kleft = 16 - kk;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const int j = 4 * kk + kq;
      if (q == kq) colraw[kq & 1][r] = a[0];
      __syncthreads();
      const double* cr = colraw[kq & 1];
      const double piv = cr[j];
      if (blockIdx.x == 0 && tid == 0 && j < nb && !(piv > 0.0) && info[0] == 0) info[0] = k0 + j + 1;
      const double rd = rsqrt(piv);
      const double lij = (r > j) ? cr[r] * rd : 0.0;
      if (q == kq) {
        if (r > j) a[0] = lij;
        if (r == j) {
          a[0] = piv * rd;
          rds[j] = rd;
        }
      }
      const double lr = lij * rd;  // L[r][j] / L[j][j]: the update uses the unscaled column
      const double* crq = cr + 4 * kk + q;
      if (q > kq) a[0] = fma(-lr, crq[0], a[0]);
#pragma unroll
      for (int k = 1; k < 16; ++k)
        if (k < kleft) a[k] = fma(-lr, crq[4 * k], a[k]);
    }
    // column group kk is final: park it in shared memory and rotate the window
    Ls[r * kLS + 4 * kk + q] = (4 * kk + q <= r) ? a[0] : 0.0;
#pragma unroll
    for (int k = 0; k < 15; ++k) a[k] = a[k + 1];
    a[15] = 0.0;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < kNB * kNB; e += 256) {
      const int rr = e & 63, c = e >> 6;
      if (rr < nb && c < nb && c <= rr) M[(int64_t)(k0 + c) * ld + k0 + rr] = Ls[rr * kLS + c];
    }
    if (tid < nb) rdiag_g[k0 + tid] = rds[tid];
    return;
  }
  if ((int)blockIdx.x == 1 + rb) {  // right-hand-side row
    if (tid < kNB) bk[tid] = tid < nb ? bvec[k0 + tid] : 0.0;
    __syncthreads();
    if (tid < 32) {
      for (int c = 0; c < nb; ++c) {
        const double yc = bk[c] * rds[c];
        if (tid == 0) yvec[k0 + c] = yc;
        const int r0 = tid, r1 = tid + 32;
        if (r0 > c) bk[r0] -= Ls[r0 * kLS + c] * yc;
        if (r1 > c) bk[r1] -= Ls[r1 * kLS + c] * yc;
        __syncwarp();
      }
    }
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
#pragma unroll 1
  for (int kc = 0; kc < 16; ++kc) {
    const int kleft = 16 - kc;
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      const int c = 4 * kc + cq;
      const double xc = quad_bcast(x[0] * rds[c], cq);  // slice cq owns column c
      if (q == cq) x[0] = xc;
      const double* lc = Ls + (4 * kc + q) * kLS + c;  // L[cp][c] for cp = 4*(kc+k)+q
      if (q > cq) x[0] = fma(-xc, lc[0], x[0]);
#pragma unroll
      for (int k = 1; k < 16; ++k)
        if (k < kleft) x[k] = fma(-xc, lc[4 * k * kLS], x[k]);
    }
    const int cst = 4 * kc + q;
    if (row < m && cst < nb) M[(int64_t)(k0 + cst) * ld + row] = x[0];
#pragma unroll
    for (int k = 0; k < 15; ++k) x[k] = x[k + 1];
    x[15] = 0.0;
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

// Trailing update A22 -= L21 * L21'  (lower 64x64 tiles only) on CTAs [0, ntiles); CTAs >= ntiles update the
// right-hand side below the block: b_i -= sum_c L[i, k0+c] y_c (128 rows each).  128 threads.
// tile0: first triangular tile index handled by CTA 0 (the look-ahead sequence passes 1: the diagonal tile right below
// the panel belongs to k_chol_diag).
// kpan: 64-column panels applied in one read-modify-write of the tile, ending with panel k0 (2 = panels k0 - 64 and k0:
// the paired sequence, which halves the passes over the trailing matrix).  colonly: the tiles are those of the first
// block column only (tile index = block row).
__global__ void __launch_bounds__(128, 3) k_syrk_update(double* __restrict__ M, int64_t ld, int m, int k0, int ntiles,
                                                     double* __restrict__ bvec, const double* __restrict__ yvec,
                                                     int tile0, int kpan, int colonly) {
  extern __shared__ double tile_sh[];
  double* As = tile_sh;
  double* Bs = tile_sh + kNB * kTS;
  const int nb = min(kNB, m - k0);
  const int base = k0 + nb;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= ntiles) {
    if (tid < kNB) As[tid] = tid < nb ? yvec[k0 + tid] : 0.0;
    __syncthreads();
    const int row = base + ((int)blockIdx.x - ntiles) * 128 + tid;
    if (row < m) {
      const double* rp = M + (int64_t)k0 * ld + row;
      double sacc[4] = {0.0, 0.0, 0.0, 0.0};
      int c = 0;
      for (; c + 16 <= nb; c += 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = rp[(int64_t)(c + u) * ld];
#pragma unroll
        for (int u = 0; u < 16; ++u) sacc[u & 3] = fma(v[u], As[c + u], sacc[u & 3]);
      }
      for (; c < nb; ++c) sacc[0] = fma(rp[(int64_t)c * ld], As[c], sacc[0]);
      bvec[row] -= (sacc[0] + sacc[1]) + (sacc[2] + sacc[3]);
    }
    return;
  }
  int ti, tj;
  if (colonly) {
    ti = blockIdx.x + tile0;
    tj = 0;
  } else {
    const int tt = blockIdx.x + tile0;
    int r = (int)((sqrt(8.0 * (double)tt + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= tt) ++r;
    while (r * (r + 1) / 2 > tt) --r;
    ti = r;
    tj = tt - r * (r + 1) / 2;
  }
  const int r0 = base + ti * kNB, c0 = base + tj * kNB;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  double acc[4][4][2] = {};
  if (ti == tj) Bs = As;  // diagonal tile: both operands are the same panel rows
#pragma unroll 1
  for (int h = kpan - 1; h >= 0; --h) {  // earlier panels are full (only the last block of the matrix can be short)
    const int kc = k0 - h * kNB, kn = h ? kNB : nb;
    if (h != kpan - 1) __syncthreads();
    load_tile_T(As, M, ld, m, kc, r0, kn, tid);
    if (ti != tj) load_tile_T(Bs, M, ld, m, kc, c0, kn, tid);
    __syncthreads();
    tile_mma_64(As, Bs, wm, wn, g, t, acc);
  }
  // read-modify-write in four batches of eight independent loads: the batch latency hides behind the other resident
  // CTAs (keeping all 32 old values in registers across the MMAs cost a third CTA per SM: 230 -> ~170 registers)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double old[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = r0 + wm + 8 * i + g, col = c0 + wn + 8 * j + 2 * t + h;
        old[j][h] = (row < m && col < m && row >= col) ? __ldcg(M + (int64_t)col * ld + row) : 0.0;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = r0 + wm + 8 * i + g, col = c0 + wn + 8 * j + 2 * t + h;
        if (row < m && col < m && row >= col) M[(int64_t)col * ld + row] = old[j][h] - acc[i][j][h];
      }
  }
}

// ---- look-ahead sequence (default): k_chol_diag -> k_chol_trsm on the main stream, k_syrk_update on a second one ------
// The per-step chain of the right-looking factorisation is  factor L_kk -> L21 = A21 L_kk^-T -> update of the next
// diagonal tile -> factor L_(k+1)(k+1).  Everything else (the rest of the trailing update) is off that chain, so it runs
// on a second stream while the next diagonal block is being factored:
//   k_chol_diag(k)  : ONE CTA.  D = A_kk - L_(k,k-1) L_(k,k-1)' (the previous panel's update of this tile, DMMA), then
//                     a two-level factorisation of the 64x64 tile: four 16-column panels; a panel is factored by ONE
//                     warp entirely in registers (lane l owns rows l and l+32 of the panel, pivots and multipliers move
//                     by warp shuffles: no barrier and no shared-memory round trip on the per-column chain), the rank-16
//                     update of the rest of the tile is 8x8x4 DMMAs spread over all warps.  The right-hand side is a
//                     65th row of the tile, so y_k = L_kk^-1 b_k falls out of the same sweep.
//   k_chol_trsm(k)  : X L_kk' = A21 by substitution, one thread per row with the 64 entries of the row in registers
//                     (L_kk broadcast from shared memory), and b_i -= L21[i,:] y_k for the rows below.
//   k_syrk_update(k): A22 -= L21 L21' for all lower tiles except the first diagonal one (tile0 = 1).
constexpr int kCholDiagThreads = 256;
constexpr int kCholDiagSmem = (kNB * kLS + 2 * kNB * kTS + 4 * kNB) * 8;
#define SCS_STAMP(i)                                     \
  do {                                                   \
    if (prof != nullptr && tid == 0) prof[i] = clock64(); \
  } while (0)
__global__ void __launch_bounds__(kCholDiagThreads)
k_chol_diag(double* __restrict__ M, int64_t ld, int m, int k0, double* __restrict__ rdiag_g, int* __restrict__ info,
            const double* __restrict__ bvec, double* __restrict__ yvec, long long* __restrict__ prof, int nprev) {
  // nprev: previous panels whose update of this tile is still pending (0 for the first block, 1 in the plain
  // look-ahead sequence, 2 at the pair boundaries of the paired sequence)
  extern __shared__ double dsh[];
  double* D = dsh;                 // the tile, D[r][c] (row stride kLS); the factor on exit
  double* As = D + kNB * kLS;      // phase 0: As[p][i] = L[k0+i, k0-64+p]; phase 1: Ps[k][r] = panel column k, row r
  double* As2 = As + kNB * kTS;    // phase 0, nprev = 2: As2[p][i] = L[k0+i, k0-128+p]
  double* rds = As2 + kNB * kTS;   // 1 / L_jj
  double* bk = rds + kNB;          // right-hand side of this block, updated panel by panel
  double* ys = bk + kNB;           // y_k
  const int nb = min(kNB, m - k0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  SCS_STAMP(0);
  // ---- phase 0: this tile with the previous panel's update applied (identity padding past nb).  One round trip:
  // every thread has its 16 entries of the tile and its 16 entries of the panel rows in flight before the first store.
  {
    const int i = tid & 63, pq = tid >> 6;
    double vd[16], va[16], va2[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = pq + 4 * u;
      vd[u] = (i < nb && c <= i) ? M[(int64_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0);
      va[u] = (nprev >= 1 && i < nb) ? M[(int64_t)(k0 - kNB + c) * ld + k0 + i] : 0.0;
      va2[u] = (nprev >= 2 && i < nb) ? M[(int64_t)(k0 - 2 * kNB + c) * ld + k0 + i] : 0.0;
    }
    if (tid >= 192) bk[tid - 192] = (tid - 192) < nb ? bvec[k0 + tid - 192] : 0.0;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = pq + 4 * u;
      D[i * kLS + c] = vd[u];
      As[c * kTS + i] = va[u];
      if (nprev >= 2) As2[c * kTS + i] = va2[u];
    }
  }
  __syncthreads();
  SCS_STAMP(1);
  if (nprev >= 1) {  // D -= L(k,k-1) L(k,k-1)' (and the panel before): 4 warps x (32x32), 8 LDS per 16 DMMAs
    if (tid < 128) {
      const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
      double acc[4][4][2] = {};
      if (nprev >= 2) tile_mma_64(As2, As2, wm, wn, g, t, acc);
      tile_mma_64(As, As, wm, wn, g, t, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = wm + 8 * i + g, col = wn + 8 * j + 2 * t + h;
            if (col <= row) D[row * kLS + col] -= acc[i][j][h];  // rows past nb: As is zero there
          }
    }
    __syncthreads();
  }
  SCS_STAMP(2);
  // ---- phase 1: four 16-column panels
#pragma unroll 1
  for (int c0 = 0; c0 < kNB; c0 += 16) {
    if (warp == 0) {
      const int rA = c0 + lane, rB = c0 + lane + 32;
      double a[16], b2[16], e[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        a[k] = rA < kNB ? D[rA * kLS + c0 + k] : 0.0;
        b2[k] = rB < kNB ? D[rB * kLS + c0 + k] : 0.0;
        e[k] = bk[c0 + k];
      }
      double piv = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (lane == 0 && c0 + j < nb && !(piv > 0.0) && info[0] == 0) info[0] = k0 + c0 + j + 1;
        const double rd = rsqrt(piv);
        const double la = lane > j ? a[j] * rd : (lane == j ? piv * rd : 0.0);
        const double lb = b2[j] * rd, le = e[j] * rd;
        a[j] = la;
        b2[j] = lb;
        e[j] = le;
        if (lane == j) rds[c0 + j] = rd;
        if (j + 1 < 16) {
          // the next pivot first, from lane j+1's own multiplier (no shuffle on the pivot chain); the general update
          // below produces the same bits for that entry
          piv = __shfl_sync(0xffffffffu, fma(-la, la, a[j + 1]), j + 1);
        }
#pragma unroll
        for (int c = j + 1; c < 16; ++c) {
          const double mlt = __shfl_sync(0xffffffffu, la, c);  // L[c0+c][c0+j]
          a[c] = fma(-la, mlt, a[c]);
          b2[c] = fma(-lb, mlt, b2[c]);
          e[c] = fma(-le, mlt, e[c]);
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (rA < kNB) {
          D[rA * kLS + c0 + k] = (c0 + k <= rA) ? a[k] : 0.0;
          if (lane >= 16) As[k * kTS + rA] = a[k];
        }
        if (rB < kNB) {
          D[rB * kLS + c0 + k] = b2[k];
          As[k * kTS + rB] = b2[k];
        }
        if (lane == 0) ys[c0 + k] = e[k];
      }
    }
    __syncthreads();
    SCS_STAMP(3 + (c0 >> 3));
    const int T0 = c0 + 16;
    if (T0 < kNB) {
      const int nt8 = (kNB - T0) >> 3, ntile = nt8 * (nt8 + 1) / 2;
      for (int tix = warp; tix < ntile; tix += kCholDiagThreads / 32) {
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= tix) ++ti;
        const int tj = tix - ti * (ti + 1) / 2;
        const int r0 = T0 + 8 * ti, cc0 = T0 + 8 * tj;
        double u0 = 0.0, u1 = 0.0;
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx)
          dmma_m8n8k4(u0, u1, As[(4 * sidx + t) * kTS + r0 + g], As[(4 * sidx + t) * kTS + cc0 + g]);
        D[(r0 + g) * kLS + cc0 + 2 * t] -= u0;
        D[(r0 + g) * kLS + cc0 + 2 * t + 1] -= u1;
      }
      if (tid >= 192 && T0 + (tid - 192) < kNB) {  // the right-hand-side row
        const int c = T0 + (tid - 192);
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) sacc = fma(ys[c0 + k], As[k * kTS + c], sacc);
        bk[c] -= sacc;
      }
      __syncthreads();
      SCS_STAMP(4 + (c0 >> 3));
    }
  }
  // ---- phase 2: store L_kk, 1/L_jj and y_k
  {
    const int i = tid & 63, pq = tid >> 6;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = pq + 4 * u;
      if (i < nb && c <= i) M[(int64_t)(k0 + c) * ld + k0 + i] = D[i * kLS + c];
    }
  }
  if (tid < nb) {
    rdiag_g[k0 + tid] = rds[tid];
    yvec[k0 + tid] = ys[tid];
  }
  SCS_STAMP(10);
}

// X L_kk' = A21 for the 64 rows [k0 + 64 + 64*blockIdx.x, ...) by substitution, right-looking over the columns, and the
// update of the right-hand side below the block.  Two threads per row (the first four warps; all eight stage L_kk):
// thread h of a row keeps the 32 entries of its row with column parity h in registers; column c is finished by its
// owner (h = c & 1), handed to the partner lane with one shuffle, and both subtract it from their later columns with
// L_kk broadcast from shared memory.  L_kk is stored parity-split (Lsm[c][h][k] = L[2k+h][c]) so that a 128-bit load
// brings two entries of one parity: the kernel is bound by these broadcast LDS.128 (11 cycles each, measured), and the
// split halves their number per warp (one thread per row with 64 registers: 9.5k cycles; the 4-threads-per-row
// shuffle layout of k_panel: 15.5k).  Block k is full (there are rows below it).
// kIdentity: the rows are those of the identity instead, for every diagonal block at once (k0 = 64*blockIdx.x), and
// X = L_kk^-T is stored as Wt[blockIdx.x][r][c] — the layout k_bwd_all / k_bwd_p2p read.
constexpr int kTrsmThreads = 256;
constexpr int kLH = kNB / 2 + 2;  // doubles per (column, parity) slice: 32 entries + padding (even: 16-byte aligned slices,
                                  // and the two parities of a column land in different banks)
template <bool kIdentity>
__global__ void __launch_bounds__(kTrsmThreads) k_chol_trsm(double* __restrict__ M, int64_t ld, int m, int k0_,
                                                            const double* __restrict__ rdiag_g,
                                                            double* __restrict__ bvec, const double* __restrict__ yvec,
                                                            double* __restrict__ Wt_all, long long* __restrict__ prof) {
  __shared__ __align__(16) double Lsm[kNB * 2 * kLH];  // Lsm[(2c + h) * kLH + k] = L[k0 + 2k + h][k0 + c]
  __shared__ double rds[kNB], ys[kNB];
  const int tid = threadIdx.x;
  const int k0 = kIdentity ? (int)blockIdx.x * kNB : k0_;
  const int nb = min(kNB, m - k0);
  const int r = tid >> 1, h = tid & 1;  // row within the block and column parity (threads 0..127)
  const int row = k0 + kNB + (int)blockIdx.x * kNB + r;
  if (blockIdx.x != 0) prof = nullptr;
  SCS_STAMP(0);
  double x[kNB / 2];
  {
    const int i = tid & 63, pq = tid >> 6;
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = pq + 4 * u;
      v[u] = (i < nb && c < nb && i >= c) ? M[(int64_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0);
    }
    if (tid < 2 * kNB) {
#pragma unroll
      for (int k = 0; k < kNB / 2; ++k) {
        const int c = 2 * k + h;
        if (kIdentity)
          x[k] = c == r ? 1.0 : 0.0;
        else
          x[k] = row < m ? M[(int64_t)(k0 + c) * ld + row] : 0.0;
      }
    }
    if (tid < kNB) {
      rds[tid] = tid < nb ? rdiag_g[k0 + tid] : 1.0;
      ys[tid] = (!kIdentity && tid < nb) ? yvec[k0 + tid] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) Lsm[(2 * (pq + 4 * u) + (i & 1)) * kLH + (i >> 1)] = v[u];
  }
  __syncthreads();
  SCS_STAMP(1);
  if (tid >= 2 * kNB) return;
  const double* lh = Lsm + h * kLH;  // this thread's parity slices: lh[2 c kLH + k] = L[2k + h][c]
  const int lane = tid & 31;
  double sacc = 0.0;
#pragma unroll
  for (int c = 0; c < kNB; ++c) {
    const int ko = c >> 1, ho = c & 1;
    const double xc = __shfl_sync(0xffffffffu, x[ko] * rds[c], (lane & ~1) | ho);  // the owner's finished entry
    if (h == ho) {
      x[ko] = xc;
      sacc = fma(xc, ys[c], sacc);
    }
    const double* lc = lh + 2 * c * kLH;
    if (ho == 0) {  // even column: the odd-parity thread still holds column c + 1 at index ko
      if (h == 1) x[ko] = fma(-xc, lc[ko], x[ko]);
    }
#pragma unroll
    for (int k = ko + 1; k < kNB / 2; ++k) x[k] = fma(-xc, lc[k], x[k]);
  }
  SCS_STAMP(2);
  if (kIdentity) {
    double* Wt = Wt_all + (int64_t)blockIdx.x * kNB * kNB;
#pragma unroll
    for (int k = 0; k < kNB / 2; ++k) Wt[r * kNB + 2 * k + h] = x[k];
  } else {
    sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
    if (row < m) {
#pragma unroll
      for (int k = 0; k < kNB / 2; ++k) M[(int64_t)(k0 + 2 * k + h) * ld + row] = x[k];
      if (h == 0) bvec[row] -= sacc;
    }
  }
  SCS_STAMP(3);
}

// ---- backward substitution  L' d = y  in two launches ----------------------------------------------------------
// k_invdiag: Wt_k = (L_kk^-1)' for every 64x64 diagonal block at once (one CTA per block; thread j builds column j of
// L_kk^-1 by forward substitution, all 64 columns in parallel).  Stored as Wt[k][c][r] = (L_kk^-1)[r][c].
__global__ void __launch_bounds__(64) k_invdiag(const double* __restrict__ M, int64_t ld, int m,
                                                double* __restrict__ Wt) {
  __shared__ double Ws[kNB * kLS];  // Ws[p][j] = (L^-1)[p][j]
  const int k0 = blockIdx.x * kNB, nb = min(kNB, m - k0), j = threadIdx.x;
  const double* Lb = M + (int64_t)k0 * ld + k0;  // L[i][p] = Lb[p*ld + i] (a broadcast read: same address in every lane)
  for (int i = 0; i < kNB; ++i) {  // row i of L^-1; thread j owns column j (zero above the diagonal)
    double acc = 0.0;
    if (i < nb) {
      for (int p = 0; p < i; ++p) acc = fma(__ldg(Lb + (int64_t)p * ld + i), Ws[p * kLS + j], acc);  // Ws[p][j] = 0 for p < j
      Ws[i * kLS + j] = ((i == j ? 1.0 : 0.0) - acc) / __ldg(Lb + (int64_t)i * ld + i);
    } else {
      Ws[i * kLS + j] = i == j ? 1.0 : 0.0;  // ragged last block: identity padding
    }
  }
  __syncthreads();
  double* out = Wt + (int64_t)blockIdx.x * kNB * kNB;
  for (int e = j; e < kNB * kNB; e += 64) {
    const int r = e & 63, c = e >> 6;
    out[c * kNB + r] = Ws[r * kLS + c];
  }
}

// k_bwd_all: one persistent kernel, gridDim.x <= number of SMs (all CTAs co-resident; cooperative launch).  For block
// k = nblk-1 .. 0: every CTA waits until all updates of y_k have landed (grid barrier = one global counter), forms
// d_k = Wt_k y_k itself (64x64 mat-vec, redundant but parallel), then applies  y_j -= L[block k, j]' d_k  to its own
// columns j < k0 (64 columns per CTA and sweep, 4 threads per column).
__global__ void __launch_bounds__(256) k_bwd_all(const double* __restrict__ M, int64_t ld, int m,
                                                 const double* __restrict__ Wt, double* __restrict__ y,
                                                 double* __restrict__ d, unsigned long long* __restrict__ bar) {
  __shared__ double yk[kNB], dk[kNB];
  const int nblk = (m + kNB - 1) / kNB;
  const int tid = threadIdx.x, c = tid >> 2, q = tid & 3;
  unsigned long long want = 0;
  for (int k = nblk - 1; k >= 0; --k) {
    const int k0 = k * kNB, nb = min(kNB, m - k0);
    if (tid == 0 && want > 0) {  // every CTA has finished the previous block's updates
      const long long t0 = clock64();
      while (true) {
        unsigned long long seen;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory");
        if (seen >= want) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __syncthreads();
    if (tid < kNB) yk[tid] = tid < nb ? __ldcg(y + k0 + tid) : 0.0;
    __syncthreads();
    {  // d_k[c] = sum_r (L_kk^-1)[r][c] y_k[r] = sum_r Wt[c][r] y_k[r]
      const double* w = Wt + (int64_t)k * kNB * kNB + c * kNB + 16 * q;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc = fma(w[i], yk[16 * q + i], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (q == 0) dk[c] = c < nb ? acc : 0.0;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) d[k0 + tid] = dk[tid];
    for (int col0 = blockIdx.x * kNB; col0 < k0; col0 += gridDim.x * kNB) {
      const int col = col0 + c;  // col < k0 <= m always (k0 is a multiple of 64)
      const double* lp = M + (int64_t)col * ld + k0 + 16 * q;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (16 * q + i < nb) acc = fma(lp[i], dk[16 * q + i], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (q == 0) y[col] -= acc;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(bar, 1ULL);
    }
    want += gridDim.x;
  }
}

// k_bwd_p2p: the same sweep with point-to-point dependencies instead of grid barriers.  CTA c owns column block c: it
// keeps y_c in registers, subtracts L[block j, block c]' d_j for j = nblk-1 .. c+1 as soon as CTA j has published d_j
// (one flag per block, release / acquire), then forms d_c = (L_cc^-1)' y_c and publishes it.  The only serial chain is
// flag -> 64 doubles of d_j -> two 64x64 mat-vecs -> flag; the L blocks and Wt_c are prefetched before the wait.
// All CTAs must be co-resident (cooperative launch, gridDim.x = nblk).
__global__ void __launch_bounds__(256) k_bwd_p2p(const double* __restrict__ M, int64_t ld, int m,
                                                 const double* __restrict__ Wt, const double* __restrict__ y,
                                                 double* __restrict__ d, int* __restrict__ flags) {
  __shared__ double dj[kNB], yk[kNB];
  const int nblk = (m + kNB - 1) / kNB;
  const int cblk = blockIdx.x, k0c = cblk * kNB, nbc = min(kNB, m - k0c);
  const int tid = threadIdx.x, col = tid >> 2, q = tid & 3;
  double w[16];
  {
    const double* wp = Wt + (int64_t)cblk * kNB * kNB + col * kNB + 16 * q;
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = wp[i];
  }
  double yv = (q == 0 && col < nbc) ? y[k0c + col] : 0.0;
  for (int j = nblk - 1; j > cblk; --j) {
    const int k0j = j * kNB, nbj = min(kNB, m - k0j);
    double lp[16];
    {
      const double* src = M + (int64_t)(k0c + col) * ld + k0j + 16 * q;  // col < nbc = 64 here (cblk is not the last block)
#pragma unroll
      for (int i = 0; i < 16; ++i) lp[i] = (16 * q + i < nbj) ? src[i] : 0.0;
    }
    if (tid == 0) {
      const long long t0 = clock64();
      while (true) {
        int seen;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + j) : "memory");
        if (seen != 0) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __syncthreads();
    if (tid < kNB) dj[tid] = tid < nbj ? __ldcg(d + k0j + tid) : 0.0;
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = fma(lp[i], dj[16 * q + i], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0) yv -= acc;
  }
  if (q == 0) yk[col] = yv;
  __syncthreads();
  {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = fma(w[i], yk[16 * q + i], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0 && col < nbc) d[k0c + col] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + cblk), "r"(1) : "memory");
  }
}

// ---- ProxGGNSCORE underdetermined branch (n+1 <= m, prox-GGN-SCORE.jl:124-127) ----------------------------------
// With N = n+1, Q = diag(q, 0), Jt = [J' | λ gr], J = diag(s) A the reference solves (I + Q Jt' H⁻¹ Jt) B = [res; 1]
// and returns d = H⁻¹ Jt B.  The last row of Q is zero, so B_N = 1 and the first n unknowns satisfy
//     (I + diag(q s) T diag(s)) B' = res − λ (q s) ∘ u,   T = A H⁻¹ A' (n x n),  u = A H⁻¹ gr,
// after which d = H⁻¹ (A' (s ∘ B') + λ gr).  T is a Gram over the COLUMNS of A (K = m), the transpose of the tall
// branch's; n < m <= 8192 here, so a plain tiled fp64 kernel is enough.
// T[i,k] = sum_j A[i,j] A[k,j] hinv[j] for the lower 64x64 tiles (i >= k), mirrored.  A: rows [0,n) of the window.
__global__ void __launch_bounds__(256)
k_rowgram(const double* __restrict__ A, int64_t ldd, int n, int m, const double* __restrict__ hr, double* __restrict__ T,
          int ldt) {
  __shared__ double Ai[16][65], Ak[16][65];
  int ti, tk;
  {
    const int tt = blockIdx.x;
    int r = (int)((sqrt(8.0 * (double)tt + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= tt) ++r;
    while (r * (r + 1) / 2 > tt) --r;
    ti = r;
    tk = tt - r * (r + 1) / 2;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  double acc[4][4] = {};
  for (int j0 = 0; j0 < m; j0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int jj = e >> 6, rr = e & 63, j = j0 + jj;
      const int ri = ti * 64 + rr, rk = tk * 64 + rr;
      const double hinv = j < m ? 1.0 / hr[j] : 0.0;
      Ai[jj][rr] = (j < m && ri < n) ? A[(int64_t)j * ldd + ri] * hinv : 0.0;
      Ak[jj][rr] = (j < m && rk < n) ? A[(int64_t)j * ldd + rk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = Ai[jj][ty + 16 * u];
        b[u] = Ak[jj][tx + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = ti * 64 + ty + 16 * u, k = tk * 64 + tx + 16 * v;
      if (i < n && k < n) {
        T[(int64_t)k * ldt + i] = acc[u][v];
        if (ti != tk) T[(int64_t)i * ldt + k] = acc[u][v];
      }
    }
}
// M = I + diag(q s) T diag(s) in place (column-major, ld = ldt);  rhs = res − λ (q s) ∘ u
__global__ void k_wide_system(double* __restrict__ T, int ldt, int n, const double* __restrict__ s,
                              const double* __restrict__ res, const double* __restrict__ q,
                              const double* __restrict__ u, double lam, double* __restrict__ rhs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (i >= n) return;
  const double c = q[i] * s[i];
  T[(int64_t)k * ldt + i] = (i == k ? 1.0 : 0.0) + (c * T[(int64_t)k * ldt + i]) * s[k];
  if (k == 0) rhs[i] = res[i] - lam * (c * u[i]);
}
// t = s ∘ B'
__global__ void k_wide_scale(const double* __restrict__ s, const double* __restrict__ b, int n, double* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = s[i] * b[i];
}
// d = (g + λ gr) / Hr
__global__ void k_wide_dir(const double* __restrict__ g, const double* __restrict__ gr, const double* __restrict__ hr,
                           double lam, int m, double* __restrict__ d) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) d[j] = (g[j] + lam * gr[j]) / hr[j];
}
// v = gr / Hr
__global__ void k_wide_v(const double* __restrict__ gr, const double* __restrict__ hr, int m, double* __restrict__ v) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) v[j] = gr[j] / hr[j];
}

// Replicated batch of the wide branch when its rows are spread over several ranks or held in sparse form: every rank
// writes its rows at their global position into a zero-initialised column-major buffer W[ldw x (m + 4)] — the m columns of
// A followed by the row vectors s, res, q, u — and a sum all-reduce makes the batch identical everywhere.
__global__ void k_wide_gather(const double* __restrict__ A, int64_t ldd, int64_t row0, int nrows, int m,
                              const double* __restrict__ s, const double* __restrict__ res,
                              const double* __restrict__ q, const double* __restrict__ u, double* __restrict__ W,
                              int64_t ldw, int64_t off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= nrows) return;
  double v;
  if (j < m)
    v = A ? A[(int64_t)j * ldd + row0 + i] : 0.0;
  else
    v = (j == m ? s : j == m + 1 ? res : j == m + 2 ? q : u)[row0 + i];
  if (j >= m || A) W[(int64_t)j * ldw + off + i] = v;
}
// the A part of the same buffer from the CSR copy of a sparse shard (one thread per row; rows are short)
__global__ void k_wide_gather_csr(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx,
                                  const double* __restrict__ vals, int64_t row0, int nrows, double* __restrict__ W,
                                  int64_t ldw, int64_t off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (int64_t p = rowptr[row0 + i]; p < rowptr[row0 + i + 1]; ++p) W[(int64_t)colidx[p] * ldw + off + i] = vals[p];
}

// ---- pivoted LU fallback (unblocked, right-looking) -------------------------------------------
// The Cholesky sequence writes the lower triangle only.  When the caller's matrix is symmetric in storage (every Gram
// this library forms is), the original is therefore still there: upper triangle + a saved diagonal.  No m x m copy.
__global__ void k_save_diag(const double* __restrict__ M, int64_t ld, int m, double* __restrict__ diag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) diag[j] = M[(int64_t)j * ld + j];
}
__global__ void k_restore_lower(double* __restrict__ M, int64_t ld, int m, const double* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < m && i > j) M[(int64_t)j * ld + i] = M[(int64_t)i * ld + j];
  if (i == j && i < m) M[(int64_t)j * ld + j] = diag[j];
}
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
