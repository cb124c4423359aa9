// K1s / K2s — streaming passes and Gram for a SPARSE shard (the README builds A with sprandn(n, m, 0.01),
// README.md:105; ggn_score_step types J as possibly SparseMatrixCSC, prox-GGN-SCORE.jl:114).
// The shard is resident twice: CSR (row pointers, 32-bit column indices, values) for the row-oriented forward pass and
// CSC for the column-oriented adjoint and Gram.  Per stored entry a pass moves 12 bytes instead of 8 bytes per dense
// element: at 1 % density a pass is ~66x less traffic than the dense kernels.
//
//   k_sp_forward : z_i = sum_p val_p x[col_p] per row (8 lanes per row), then the same loss_row as the dense pass.
//   k_sp_adjoint : g_j = sum_p val_p r[row_p] per column (one warp per column, fixed shuffle tree).
//   k_sp_gram    : column k of G = A' diag(w) A:  sum over the stored rows i of column k of (w_i a_ik) * (row i of A),
//                  accumulated by the 8 warps of a CTA in a shared-memory vector of m 64-bit FIXED-POINT integers with
//                  integer atomics (order-independent => deterministic although the warps interleave freely).  The
//                  lower triangle is then mirrored.
// All reductions use fixed trees or exact integer addition: results are bit-reproducible.
#pragma once
#include "common.cuh"
#include "kernels_stream.cuh"

namespace scs {

constexpr int kSpFwdThreads = 256;
constexpr int kSpFwdRows = kSpFwdThreads / 8;  // rows per CTA

__global__ void __launch_bounds__(kSpFwdThreads)
k_sp_forward(const int64_t* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals, int64_t n,
             int64_t row_lo, int64_t row_hi, const double* __restrict__ x, const double* __restrict__ y, LossParams lp,
             double* __restrict__ z_out, double* __restrict__ r_out, double* __restrict__ w_out,
             double* __restrict__ loss_part) {
  __shared__ double red[32];
  const int sub = threadIdx.x & 7;
  const int64_t i = (int64_t)blockIdx.x * kSpFwdRows + (threadIdx.x >> 3);
  double part = 0.0;
  const bool valid = i < n;
  double acc = 0.0;
  if (valid) {
    const int64_t p1 = rowptr[i + 1];
    for (int64_t p = rowptr[i] + sub; p < p1; p += 8) acc = fma(vals[p], x[colidx[p]], acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);  // unconditional: every lane of the warp takes part
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (valid && sub == 0) {
    double z = acc, t, r, w;
    loss_row(lp, z, y[i], t, r, w);
    if (i < row_lo || i >= row_hi) t = r = w = z = 0.0;  // rows of another mini-batch
    if (lp.weight_kind == 2) {  // GGN parts (see k_forward)
      z = t;
      t = 0.0;
    }
    part = t;
    if (z_out) z_out[i] = z;
    if (r_out) r_out[i] = r;
    if (w_out) w_out[i] = w;
  }
  const double tot = block_sum<kSpFwdThreads>(part, red);
  if (threadIdx.x == 0) loss_part[blockIdx.x] = tot;
}

// g_j = sum over the stored entries of column j of val * r[row].  One warp per column; grid = ceil(m / 8) CTAs of 256.
__global__ void __launch_bounds__(256)
k_sp_adjoint(const int64_t* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ cvals, int m,
             const double* __restrict__ r, double* __restrict__ g) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= m) return;
  double a0 = 0.0, a1 = 0.0;
  const int64_t p1 = colptr[j + 1];
  int64_t p = colptr[j] + lane;
  for (; p + 32 < p1; p += 64) {
    a0 = fma(cvals[p], r[rowidx[p]], a0);
    a1 = fma(cvals[p + 32], r[rowidx[p + 32]], a1);
  }
  if (p < p1) a0 = fma(cvals[p], r[rowidx[p]], a0);
  const double s = warp_sum(a0 + a1);
  if (lane == 0) g[j] = s;
}

// Sparse Gram, column by column (Gustavson with a dense accumulator), all 8 warps of a CTA on the same column:
//   G[j, k] = sum over the stored rows i of column k of (w_i a_ik) a_ij,   j >= k  (the lower triangle; mirrored afterwards)
// Warp q takes rows q, q + 8, ... of the column, its lanes the entries of a row.  The warps meet in ONE shared accumulator,
// so the additions are made ORDER-INDEPENDENT instead of ordered: every addend is converted to 64-bit fixed point and added
// with an integer shared-memory atomic — integer addition is associative, the result is bit-reproducible whatever the
// interleaving.  Scaling: with cs_j = 1 / sqrt(sum_i |w_i| a_ij^2) (k_sp_colscale) every scaled entry satisfies
// |G_jk cs_j cs_k| <= 1 (Cauchy-Schwarz on the absolute values, which also bounds every partial sum), so addends are taken
// as rint(x 2^61): at most 0.5 * 2^-61 off each, <= nnz_col * 2^-62 ~ 1e-15 of the diagonal scale in the worst case.
// A non-finite addend (NaN / Inf weight) poisons the column like it would in floating point.
__global__ void __launch_bounds__(256)
k_sp_colscale(const int64_t* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ cvals,
              const double* __restrict__ w, int m, double* __restrict__ cs) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= m) return;
  double a = 0.0;
  for (int64_t p = colptr[j] + lane; p < colptr[j + 1]; p += 32) a = fma(fabs(w[rowidx[p]]), cvals[p] * cvals[p], a);
  a = warp_sum(a);
  if (lane == 0) cs[j] = a > 0.0 ? 1.0 / sqrt(a) : 0.0;  // NaN sums compare false: the column scale 0 makes cinv = 0 ...
}
constexpr double kSpFix = 2305843009213693952.0;  // 2^61
#ifndef SCS_SP_ROWS
#define SCS_SP_ROWS 2
#endif
constexpr int kSpRows = SCS_SP_ROWS;  // stored rows a warp keeps in flight
__global__ void __launch_bounds__(256)
k_sp_gram(const int64_t* __restrict__ colptr, const int* __restrict__ rowidx, const double* __restrict__ cvals,
          const int64_t* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
          const double* __restrict__ w, const double* __restrict__ cs, int m, double* __restrict__ G) {
  // 64-bit accumulators kept as two 32-bit halves: shared memory has native 32-bit atomic adds (ATOMS.ADD) but a 64-bit
  // add compiles to a compare-and-swap loop; the carry out of the low half is recovered from the value ATOMS.ADD returns
  extern __shared__ unsigned int sp_acc[];  // [0, m): low halves, [m, 2m): high halves
  __shared__ int s_bad;
  unsigned int* acc_lo = sp_acc;
  unsigned int* acc_hi = sp_acc + m;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  for (int k = blockIdx.x; k < m; k += gridDim.x) {
    for (int j = k + threadIdx.x; j < m; j += 256) acc_lo[j] = acc_hi[j] = 0u;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const double sk = cs[k] * kSpFix;
    const int64_t p1 = colptr[k + 1];
    // a warp takes 32 consecutive stored rows of the column at a time: their row index, weight and row extent are fetched
    // lane-parallel (one round trip for 32 rows instead of three dependent ones per row), then walked with shuffles
    for (int64_t p0 = colptr[k] + wq * 32; p0 < p1; p0 += 8 * 32) {
      const int64_t p = p0 + lane;
      double c = 0.0;
      int64_t q0 = 0, q1 = 0;
      if (p < p1) {
        const int i = rowidx[p];
        const double wi = w[i];
        c = wi * cvals[p] * sk;
        if (!(fabs(wi) < 1.0e300)) s_bad = 1;  // NaN / Inf weight (benign race: every writer stores 1)
        q0 = rowptr[i];
        q1 = rowptr[i + 1];
      }
      const int nrow = (int)(p1 - p0 < 32 ? p1 - p0 : 32);
      auto add = [&](int j, double v) {
        const unsigned long long u = (unsigned long long)__double2ll_rn(v);
        const unsigned int lo = (unsigned int)u;
        const unsigned int old = atomicAdd(&acc_lo[j], lo);
        atomicAdd(&acc_hi[j], (unsigned int)(u >> 32) + ((unsigned int)(old + lo) < lo ? 1u : 0u));
      };
      // kSpRows rows per round, up to 64 entries of each in flight before the first is consumed (a row walk is otherwise
      // a chain of dependent round trips: entries -> column scale -> atomic, twice for a row of 33..64 entries)
      for (int r = 0; r < nrow; r += kSpRows) {
        double cr[kSpRows];
        int64_t b0[kSpRows], b1[kSpRows];
#pragma unroll
        for (int h = 0; h < kSpRows; ++h) {
          const int rr = r + h < nrow ? r + h : r;
          cr[h] = r + h < nrow ? __shfl_sync(0xffffffffu, c, rr) : 0.0;
          b0[h] = __shfl_sync(0xffffffffu, q0, rr);
          b1[h] = cr[h] != 0.0 ? __shfl_sync(0xffffffffu, q1, rr) : b0[h];  // w = 0 (other mini-batch): empty walk
        }
        int jj[2 * kSpRows];
        double vv[2 * kSpRows];
#pragma unroll
        for (int h = 0; h < kSpRows; ++h)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int64_t q = b0[h] + lane + 32 * u;
            const bool on = q < b1[h];
            jj[2 * h + u] = on ? colidx[q] : -1;
            vv[2 * h + u] = on ? vals[q] : 0.0;
          }
        double sc[2 * kSpRows];
#pragma unroll
        for (int e = 0; e < 2 * kSpRows; ++e) sc[e] = jj[e] >= k ? cs[jj[e]] : 0.0;
#pragma unroll
        for (int e = 0; e < 2 * kSpRows; ++e)
          if (jj[e] >= k) add(jj[e], cr[e >> 1] * (vv[e] * sc[e]));
#pragma unroll
        for (int h = 0; h < kSpRows; ++h)  // rows with more than 64 stored entries
          for (int64_t q = b0[h] + lane + 64; q < b1[h]; q += 32) {
            const int j = colidx[q];
            if (j >= k) add(j, cr[h] * (vals[q] * cs[j]));
          }
      }
    }
    __syncthreads();
    double* out = G + (int64_t)k * m;
    const bool bad = s_bad != 0;
    const double ck = cs[k];
    for (int j = k + threadIdx.x; j < m; j += 256) {
      const double den = cs[j] * ck * kSpFix;
      const long long acc = (long long)(((unsigned long long)acc_hi[j] << 32) | acc_lo[j]);
      double g = den > 0.0 ? (double)acc / den : 0.0;
      if (bad) g = __longlong_as_double(0x7ff8000000000000LL);
      out[j] = g;
    }
    __syncthreads();
  }
}

// ---- dense -> sparse conversion on the device (benchmark-sized synthetic shards: the generator fills the dense layout,
// explicit zeros are then dropped; nothing crosses PCIe) ---------------------------------------------------------------
// counts[i] = stored entries of row i (thread per row; adjacent threads read adjacent rows of a column: coalesced)
__global__ void __launch_bounds__(256) k_nnz_rows(const double* __restrict__ A, int64_t ldd, int64_t n, int m,
                                                  int64_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int64_t c = 0;
  for (int j = 0; j < m; ++j) c += A[(int64_t)j * ldd + i] != 0.0 ? 1 : 0;
  counts[i] = c;
}
// counts[j] = stored entries of column j (CTA per column)
__global__ void __launch_bounds__(256) k_nnz_cols(const double* __restrict__ A, int64_t ldd, int64_t n, int m,
                                                  int64_t* __restrict__ counts) {
  __shared__ double red[32];
  const double* col = A + (int64_t)blockIdx.x * ldd;
  double c = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) c += col[i] != 0.0 ? 1.0 : 0.0;
  c = block_sum<256>(c, red);
  if (threadIdx.x == 0) counts[blockIdx.x] = (int64_t)c;
}
// In-place exclusive scan of `data[0..n)` (data[n] receives the total) in three launches: per-block scans of 2048 entries,
// a single-CTA scan of the block totals, and the add-back.
constexpr int kScanBlock = 2048;
__global__ void __launch_bounds__(256) k_scan_blocks(int64_t* __restrict__ data, int64_t n, int64_t* __restrict__ btot) {
  __shared__ int64_t wsum[8];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * 8;
  int64_t v[8], t = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] = base + q < n ? data[base + q] : 0;
    t += v[q];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t inc = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  int64_t off = inc - t;
  for (int q = 0; q < wid; ++q) off += wsum[q];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (base + q < n) data[base + q] = off;
    off += v[q];
  }
  if (threadIdx.x == 255) btot[blockIdx.x] = off;
}
__global__ void __launch_bounds__(1) k_scan_top(int64_t* __restrict__ btot, int64_t nb, int64_t* __restrict__ total) {
  int64_t run = 0;
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t v = btot[b];
    btot[b] = run;
    run += v;
  }
  total[0] = run;
}
__global__ void __launch_bounds__(256) k_scan_add(int64_t* __restrict__ data, int64_t n, const int64_t* __restrict__ btot) {
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + threadIdx.x * 8;
  const int64_t off = btot[blockIdx.x];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (base + q < n) data[base + q] += off;
}
// CSR fill: thread per row, columns ascending
__global__ void __launch_bounds__(256) k_fill_csr(const double* __restrict__ A, int64_t ldd, int64_t n, int m,
                                                  const int64_t* __restrict__ rowptr, int* __restrict__ colidx,
                                                  double* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int64_t p = rowptr[i];
  for (int j = 0; j < m; ++j) {
    const double v = A[(int64_t)j * ldd + i];
    if (v != 0.0) {
      colidx[p] = j;
      vals[p] = v;
      ++p;
    }
  }
}
// CSC fill: CTA per column, rows ascending (ballot compaction, 256 rows per round)
__global__ void __launch_bounds__(256) k_fill_csc(const double* __restrict__ A, int64_t ldd, int64_t n, int m,
                                                  const int64_t* __restrict__ colptr, int* __restrict__ rowidx,
                                                  double* __restrict__ cvals) {
  __shared__ int wcnt[8];
  const double* col = A + (int64_t)blockIdx.x * ldd;
  int64_t p = colptr[blockIdx.x];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t i0 = 0; i0 < n; i0 += 256) {
    const int64_t i = i0 + threadIdx.x;
    const double v = i < n ? col[i] : 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, v != 0.0);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1u)), tot = 0;
    for (int q = 0; q < 8; ++q) {
      if (q < wid) before += wcnt[q];
      tot += wcnt[q];
    }
    if (v != 0.0) {
      rowidx[p + before] = (int)i;
      cvals[p + before] = v;
    }
    p += tot;
    __syncthreads();
  }
}

}  // namespace scs
