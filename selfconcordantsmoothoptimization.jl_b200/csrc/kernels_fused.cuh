// K1f — single-pass fused gradient:  z = A x,  per-row loss pieces (term, r, w),  g = A' r  with ONE read of A.
// Replaces the k_forward + k_adjoint pair (two reads of A) wherever the objective and the gradient are both needed
// at the same x: every iteration of ProxLQNSCORE (prox-L-BFGS-SCORE.jl:101,150), and the f(x) + Jt*residual /
// gradient(f,x) pair of ProxGGNSCORE / ProxNSCORE (iterate.jl:189, prox-GGN-SCORE.jl:44-56,130, prox-N-SCORE.jl:49-69).
//
// HBM-bound: 8*n*m algorithmic bytes per logical gradient — half of the two-pass path.
//
// A row of A spans all m columns (16-32 KB), so a row panel cannot stay in one SM between its two uses.  A thread-
// block CLUSTER owns the panel instead: CTA rank c takes the column slice [256c, 256c+256).  A TMA stage is 32 rows
// of the slice (two boxes of 32 rows x 128 columns, 256 contiguous bytes per column: the TMA engine serves about one
// box row per 8 cycles per SM whatever its length, so 128-byte rows cap the stream at ~4.5 TB/s, 256-byte rows at
// ~6.5 TB/s — tools/micro/tma_stream_bench.cu).  A stage is consumed as two 16-row panels: each thread copies its 16
// values of a panel into REGISTERS and the stage is freed as soon as both halves are out; the partial dot
// products of the slices are exchanged over distributed shared memory, and LAG panels later every thread applies
// the row weights r to the values it kept, accumulating its share of g in registers over the whole kernel.
//
//   warps 0-7  column warps; warp w covers columns [32w, 32w+32) of the slice, lane = (row pair rp, column sub-group):
//              the thread owns rows 2rp, 2rp+1 of 8 columns.  phase A(t): LDS.128 x 8 (a quarter-warp reads 128
//              contiguous bytes of one column: conflict-free without swizzling) -> registers; 16 FMAs into
//              the two row sums, two shuffle steps over the 4 sub-groups -> zpart[slot][warp][row]; arrive pbar[slot].
//              phase B(t-LAG): wait rbar[slot]; g_j(thread) += a_ij r_i for its two rows: 16 FMAs, no reduction — the
//              8 row-pair lanes of a column are folded once, after the last panel.
//   warp 8     TMA producer (one lane): two boxes per stage, mbarrier complete_tx.
//   warp 9     push warp: sums the 8 warp partials (fixed order) and st.async's the 16 row sums to the panel's OWNER
//              CTA (rank t mod cluster size); the store itself signals complete_tx on the owner's zbar[slot].
//   warp 10    loss warp, for the panels this CTA owns: z_i = sum over ranks in rank order, r_i first (it is on the
//              critical path) and st.async of r to rs[slot] of every CTA (complete_tx on their rbar[slot]); then the
//              loss term, w_i and the z / r / w rows to global memory.
// More than 16 x 256 columns do not fit one cluster: the kernel then covers the LAST 4096 columns (the tensor map, x and
// gpart are offset by the host) and receives the other columns' contribution to z through z_in (a k_forward pass over
// them; their share of g is a k_adjoint pass afterwards): 1.5 reads of A per objective + gradient instead of 2.
// The kernel works on the row window [row_base, row_base + 16*npanels) of the shard (row_base a multiple of 32); rows
// outside [win_lo, win_hi) get r = w = z = 0 (padding, or rows of another mini-batch).
// Rotating the owner spreads the exp/log work over the cluster.  No atomics anywhere: per-cluster partial g and
// per-CTA loss sums are reduced by k_colsum / k_sum_partials in fixed order => bit-reproducible for a given grid.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels_gram.cuh"    // mbarrier / TMA helpers
#include "kernels_i8gram.cuh"  // cluster_ctarank, cluster_sync_all
#include "kernels_stream.cuh"  // LossParams, loss_row

namespace scs {

constexpr int kFuRows = 16;       // rows per exchange panel
constexpr int kFuStageRows = 32;  // rows per TMA stage = two panels: 256 contiguous bytes per column and request
constexpr int kFuCols = 256;      // columns per CTA
constexpr int kFuBoxCols = 128;   // columns per TMA box (two boxes per stage)
constexpr int kFuColWarps = 8;
constexpr int kFuThreads = (kFuColWarps + 4) * 32;  // + producer, push, loss, (idle) = one helper warpgroup
constexpr int kFuStageBytes = kFuStageRows * kFuCols * 8;  // 64 KB
constexpr int kFuStages = 3;
constexpr int kFuLag = 3;                // phase B trails phase A by this many panels (kept in registers)
constexpr int kFuBufs = kFuLag + 1;
constexpr int kFuSlots = 8;              // exchange ring depth (>= 2*LAG + 2)
constexpr int kFuMaxCluster = 16;
constexpr int kFuSmemBytes = kFuStages * kFuStageBytes + kFuSlots * kFuColWarps * 128 + kFuSlots * kFuMaxCluster * 128 +
                             kFuSlots * 128 + (2 * kFuStages + 3 * kFuSlots) * 8 + 1024;

SCS_DEVINL uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
SCS_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t rank) {  // same offset in the shared memory of CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// remote store whose completion is counted (8 bytes) by the receiver's mbarrier
SCS_DEVINL void st_async_f64(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v),
               "r"(rbar)
               : "memory");
}
SCS_DEVINL double2 lds_v2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
// mbarrier / TMA primitives on 32-bit shared-memory addresses (computed once per thread, outside the panel loop).
// Waits are plain try_wait polls (acquire.cta): the data they guard lives in this CTA's shared memory — a
// cluster-scope acquire would make ptxas emit an L1 invalidate (CCTL.IVALL) per wait.  They are bounded: a protocol
// bug traps (the launch fails with an error) instead of hanging the device.
SCS_DEVINL bool fu_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
SCS_DEVINL void fu_wait(uint32_t bar, uint32_t parity) {
  if (fu_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!fu_try(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
SCS_DEVINL void fu_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SCS_DEVINL void fu_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SCS_DEVINL void fu_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
SCS_DEVINL double lds_f64u(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
SCS_DEVINL void sts_v2(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
SCS_DEVINL void sts_f64(uint32_t addr, double a) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory");
}

// r only (same expressions, same rounding as loss_row): the part of the loss evaluation the cluster waits for
SCS_DEVINL double loss_r_only(const LossParams& lp, double z, double y) {
  if (lp.kind == 0) {
    if (lp.weight_kind == 0) {
      const double e = exp(-y * z);
      return lp.p * (-y) * (e / (1.0 + e));
    }
    const double yc = lp.label_mode == 0 ? y : (y + 1.0) / 2.0;
    const double e2 = exp(-z);
    const double yhat = 1.0 / (1.0 + e2);
    const double s = (yhat / (1.0 + e2)) * e2;
    const double om = 1.0 - yhat;
    const double res = -lp.p * (yc / yhat - (1.0 - yc) / om);
    return s * res;
  }
  if (lp.kind == 1) return (z - y) / lp.p;
  return 0.0;
}

// PROF: per-CTA wait-cycle counters (tuning only): prof[blockIdx.x*8 + {0 total, 1 warp0 wait full, 2 warp0 wait rbar,
// 3 producer wait empty, 4 push wait pbar, 5 loss wait zbar, 6 loss busy, 7 warp0 busy in phase A}]
template <bool PROF>
__global__ void __launch_bounds__(kFuThreads, 1)
k_fused_grad(const __grid_constant__ CUtensorMap amap, const double* __restrict__ x, const double* __restrict__ y,
             LossParams lp, int64_t row_base, int64_t win_lo, int64_t win_hi, int64_t npanels, int m,
             double* __restrict__ z_out, double* __restrict__ r_out,
             double* __restrict__ w_out, double* __restrict__ loss_part /* [gridDim.x] */,
             double* __restrict__ gpart /* [clusters][m] */, long long* __restrict__ prof, int dbg_mode,
             const double* __restrict__ z_in /* optional: contribution of columns this launch does not cover (m > 4096) */) {
  const long long k_t0 = PROF ? clock64() : 0;
  long long c_a = 0, c_b = 0, c_c = 0;
  constexpr int S = kFuStages, LAG = kFuLag, NB = kFuBufs;
  extern __shared__ uint8_t smem_raw[];
  // shared-memory map (32-bit addresses; identical in every CTA, so mapa reaches the same object in a peer)
  const uint32_t sm_tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;                   // [S][32 KB]
  const uint32_t sm_zpart = sm_tiles + (uint32_t)(S * kFuStageBytes);                // [slots][8 warps][16]
  const uint32_t sm_zx = sm_zpart + (uint32_t)(kFuSlots * kFuColWarps * 128);        // [slots][16 ranks][16] (owner)
  const uint32_t sm_rs = sm_zx + (uint32_t)(kFuSlots * kFuMaxCluster * 128);         // [slots][16]
  const uint32_t sm_full = sm_rs + (uint32_t)(kFuSlots * 128);
  const uint32_t sm_empty = sm_full + 8u * S;
  const uint32_t sm_pbar = sm_empty + 8u * S;
  const uint32_t sm_zbar = sm_pbar + 8u * kFuSlots;
  const uint32_t sm_rbar = sm_zbar + 8u * kFuSlots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const int64_t cid = blockIdx.x / csize, ncl = gridDim.x / csize;
  // This cluster's panels: 32-row blocks b = cid, cid + ncl, ... (two adjacent 16-row panels each, so consecutive
  // boxes fetch 256 contiguous bytes of every column); the very last block may hold a single panel.
  const int64_t nblocks = (npanels + 1) >> 1;
  int cnt = 0;
  if (nblocks > cid) {
    const int64_t nb = (nblocks - 1 - cid) / ncl + 1;
    const bool owns_last = ((nblocks - 1 - cid) % ncl) == 0;
    cnt = (int)(2 * nb - ((owns_last && (npanels & 1)) ? 1 : 0));
  }
  auto panel_of = [&](int t) -> int64_t { return 2 * (cid + (int64_t)(t >> 1) * ncl) + (t & 1); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_full + 8u * s), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_empty + 8u * s), "r"(kFuColWarps) : "memory");
    }
    for (int q = 0; q < kFuSlots; ++q) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_pbar + 8u * q), "r"(kFuColWarps) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_zbar + 8u * q), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_rbar + 8u * q), "r"(1) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before a peer stores into them

  if (warp < kFuColWarps) {
    // ===== column warps (two warpgroups): most of the register file =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // lane = (sub, rp): row pair rp = lane & 7 (rows 2rp, 2rp+1 of the panel), column sub-group sub = lane >> 3;
    // the thread owns columns c_k = 32*warp + 4k + sub (k = 0..7) of the CTA's slice, two rows of each
    const int rp = lane & 7, sub = lane >> 3;
    int jk[8];
    double xk[8], g[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      jk[k] = (int)crank * kFuCols + warp * 32 + 4 * k + sub;
      xk[k] = jk[k] < m ? x[jk[k]] : 0.0;
      g[k][0] = g[k][1] = 0.0;
    }
    // stage layout (no swizzle): box b = column / 128, then [column % 128][32 rows]: 256 bytes per column; a
    // quarter-warp (8 row pairs of one column, one 16-row half) reads 128 contiguous bytes: conflict-free
    const uint32_t t_mine = sm_tiles + (uint32_t)((warp >> 2) * (kFuStageBytes / 2) + (32 * (warp & 3) + sub) * 256 + rp * 16);
    const uint32_t rs_mine = sm_rs + (uint32_t)(rp * 16);
    const uint32_t zp_mine = sm_zpart + (uint32_t)(warp * 128 + lane * 16);
    double buf[NB][16];
    int sa = 0;          // stage of the next phase A
    uint32_t pha = 0;    // its full-barrier parity

    auto phase_a = [&](int t, double (&a)[16]) {
      const uint32_t slot = (uint32_t)t & (kFuSlots - 1);
      const long long w0 = PROF ? clock64() : 0;
      if (!(t & 1)) fu_wait(sm_full + 8u * sa, pha);  // the odd panel is the second half of the same stage
      const long long w1 = PROF ? clock64() : 0;
      if (PROF) c_a += w1 - w0;
      const uint32_t ta = t_mine + (uint32_t)(sa * kFuStageBytes) + (uint32_t)((t & 1) * 128);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double2 v = lds_v2(ta + (uint32_t)(k * 1024));
        a[2 * k] = v.x;
        a[2 * k + 1] = v.y;
      }
      __syncwarp();
      if (lane == 0) {
        if (t & 1) fu_arrive(sm_empty + 8u * sa);  // both halves live in registers now: the stage can be refilled
        if (warp == 0) fu_expect_tx(sm_rbar + 8u * slot, kFuRows * 8);  // r of this panel will come from its owner
      }
      if (t & 1) {
        if (++sa == S) {
          sa = 0;
          pha ^= 1u;
        }
      }
      double z0 = a[0] * xk[0], z1 = a[1] * xk[0], z2 = a[2] * xk[1], z3 = a[3] * xk[1];
#pragma unroll
      for (int k = 2; k < 8; k += 2) {
        z0 = fma(a[2 * k], xk[k], z0);
        z1 = fma(a[2 * k + 1], xk[k], z1);
        z2 = fma(a[2 * k + 2], xk[k + 1], z2);
        z3 = fma(a[2 * k + 3], xk[k + 1], z3);
      }
      z0 += z2;
      z1 += z3;
      z0 += __shfl_xor_sync(0xffffffffu, z0, 8);
      z1 += __shfl_xor_sync(0xffffffffu, z1, 8);
      z0 += __shfl_xor_sync(0xffffffffu, z0, 16);
      z1 += __shfl_xor_sync(0xffffffffu, z1, 16);
      if (lane < 8) sts_v2(zp_mine + slot * (uint32_t)(kFuColWarps * 128), z0, z1);
      __syncwarp();
      if (lane == 0) fu_arrive(sm_pbar + 8u * slot);
      if (PROF) c_c += clock64() - w1;
    };
    auto phase_b = [&](int t, const double (&a)[16]) {
      const uint32_t slot = (uint32_t)t & (kFuSlots - 1);
      const long long w0 = PROF ? clock64() : 0;
      fu_wait(sm_rbar + 8u * slot, ((uint32_t)t >> 3) & 1u);
      if (PROF) c_b += clock64() - w0;
      const double2 r = lds_v2(rs_mine + slot * 128u);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[k][0] = fma(a[2 * k], r.x, g[k][0]);
        g[k][1] = fma(a[2 * k + 1], r.y, g[k][1]);
      }
    };

    if (PROF && dbg_mode == 1) {  // TMA-only experiment: drain the ring, no math, no exchange
      for (int u = 0; u < (cnt + 1) / 2; ++u) {
        fu_wait(sm_full + 8u * sa, pha);
        __syncwarp();
        if (lane == 0) fu_arrive(sm_empty + 8u * sa);
        if (++sa == S) {
          sa = 0;
          pha ^= 1u;
        }
      }
      cnt = 0;
    }
    for (int t0 = 0; t0 < cnt + LAG && cnt > 0; t0 += NB) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int t = t0 + i;
        if (t < cnt) phase_a(t, buf[i]);
        if (t >= LAG && t - LAG < cnt) phase_b(t - LAG, buf[(i + NB - LAG) % NB]);
      }
    }
    // each thread summed its two rows of every panel: fold the 8 row-pair lanes of a column once, at the very end
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double v = g[k][0] + g[k][1];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (rp == 0 && jk[k] < m) gpart[cid * m + jk[k]] = v;
    }
    if (PROF && threadIdx.x == 0) {
      prof[blockIdx.x * 8 + 1] = c_a;
      prof[blockIdx.x * 8 + 2] = c_b;
      prof[blockIdx.x * 8 + 7] = c_c;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == kFuColWarps) {
      // ===== TMA producer =====
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 1;  // empty-barrier parity: the first pass over the ring finds every stage free
        const int col0 = (int)crank * kFuCols;
        for (int u = 0; u < (cnt + 1) / 2; ++u) {  // stage u = 32-row block cid + u*ncl (rows past ldd are zero-filled)
          const long long w0 = PROF ? clock64() : 0;
          fu_wait(sm_empty + 8u * s, ph);
          if (PROF) c_a += clock64() - w0;
          fu_expect_tx(sm_full + 8u * s, (uint32_t)kFuStageBytes);
          const int row0 = (int)(row_base + (cid + (int64_t)u * ncl) * kFuStageRows);
          fu_tma_2d(sm_tiles + (uint32_t)(s * kFuStageBytes), &amap, sm_full + 8u * s, row0, col0);
          fu_tma_2d(sm_tiles + (uint32_t)(s * kFuStageBytes + kFuStageBytes / 2), &amap, sm_full + 8u * s, row0,
                    col0 + kFuBoxCols);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (PROF) prof[blockIdx.x * 8 + 3] = c_a;
      }
    } else if (warp == kFuColWarps + 1) {
      // ===== push warp: warp partials -> row sums of this CTA's column slice -> the panel's owner =====
      if (PROF && dbg_mode == 1) cnt = 0;
      uint32_t owner = 0;
      const uint32_t zx_mine = sm_zx + (crank * 16u + (uint32_t)(lane & 15)) * 8u;
      for (int t = 0; t < cnt; ++t) {
        const uint32_t slot = (uint32_t)t & (kFuSlots - 1);
        const long long w0 = PROF ? clock64() : 0;
        fu_wait(sm_pbar + 8u * slot, ((uint32_t)t >> 3) & 1u);
        if (PROF) c_a += clock64() - w0;
        if (lane < kFuRows) {
          const uint32_t zp = sm_zpart + slot * (uint32_t)(kFuColWarps * 128) + (uint32_t)(lane * 8);
          double v = lds_f64u(zp);
#pragma unroll
          for (int w = 1; w < kFuColWarps; ++w) v += lds_f64u(zp + (uint32_t)(w * 128));
          st_async_f64(mapa_u32(zx_mine + slot * (uint32_t)(kFuMaxCluster * 128), owner), v,
                       mapa_u32(sm_zbar + 8u * slot, owner));
        }
        if (++owner == csize) owner = 0;
      }
      if (PROF && lane == 0) prof[blockIdx.x * 8 + 4] = c_a;
    } else if (warp == kFuColWarps + 2) {
      // ===== loss warp: the panels this CTA owns =====
      if (PROF && dbg_mode == 1) cnt = 0;
      double lacc = 0.0;
      uint32_t used = 0;  // bit q = parity of zbar[q]'s next phase on this CTA
      for (int t = (int)crank; t < cnt; t += (int)csize) {
        const uint32_t slot = (uint32_t)t & (kFuSlots - 1);
        const int64_t row = row_base + panel_of(t) * kFuRows + (lane & 15);
        const double yv = y[row];  // row < row_base + 16*npanels <= ldd (y is zero padded up to ldd)
        if (lane == 0) fu_expect_tx(sm_zbar + 8u * slot, csize * (uint32_t)(kFuRows * 8));
        const long long w0 = PROF ? clock64() : 0;
        fu_wait(sm_zbar + 8u * slot, (used >> slot) & 1u);
        const long long w1 = PROF ? clock64() : 0;
        if (PROF) c_a += w1 - w0;
        used ^= 1u << slot;
        if (lane < kFuRows) {
          const uint32_t zr = sm_zx + slot * (uint32_t)(kFuMaxCluster * 128) + (uint32_t)(lane * 8);
          double z = lds_f64u(zr);
          for (uint32_t pr = 1; pr < csize; ++pr) z += lds_f64u(zr + pr * 128u);
          if (z_in) z += z_in[row];
          const bool pad = row < win_lo || row >= win_hi;  // padding rows, or rows of another mini-batch
          const double r = pad ? 0.0 : loss_r_only(lp, z, yv);
          const uint32_t dst = sm_rs + slot * 128u + (uint32_t)(lane * 8), bar = sm_rbar + 8u * slot;
          for (uint32_t pr = 0; pr < csize; ++pr) st_async_f64(mapa_u32(dst, pr), r, mapa_u32(bar, pr));
          // off the critical path: loss term, Gram weight, row outputs
          double term, r2, w;
          loss_row(lp, z, yv, term, r2, w);
          if (pad) {
            term = 0.0;
            w = 0.0;
            z = 0.0;
          }
          if (z_out) z_out[row] = z;
          if (r_out) r_out[row] = r;
          if (w_out) w_out[row] = w;
          lacc += term;
        }
        if (PROF) c_b += clock64() - w1;
      }
      lacc = warp_sum(lacc);
      if (lane == 0) loss_part[blockIdx.x] = lacc;
      if (PROF && lane == 0) {
        prof[blockIdx.x * 8 + 5] = c_a;
        prof[blockIdx.x * 8 + 6] = c_b;
      }
    }
  }
  __syncthreads();
  cluster_sync_all();  // no CTA exits while a peer may still store into its shared memory
  if (PROF && threadIdx.x == 0) prof[blockIdx.x * 8 + 0] = clock64() - k_t0;
}

}  // namespace scs
