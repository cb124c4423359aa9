// K2 — Gauss-Newton / Newton Gram  G = A' diag(w) A  (lower triangle, fp64).
// Replaces  JQJ = Jt*Q*Jt'  (prox-GGN-SCORE.jl:121-123,129) and  hessian(f,x)  (prox-N-SCORE.jl:63).
//
// Tensor-bound: n*m*(m+1) algorithmic flops.  fp64 has no tcgen05 kind; the FP64 tensor path sm_100a
// executes is DMMA.8x8x4 (mma.sync.m8n8k4.f64 — the wider PTX shapes lower to it), so that is what the
// main loop issues.  Structure (one persistent CTA per SM, 9 warps):
//   warp 8   : TMA producer.  Per 16-row stage: two cp.async.bulk.tensor.2d boxes (16 rows x 128 columns of
//              A, SWIZZLE_128B) + one cp.async.bulk of 16 weights, completing on an mbarrier ("full").
//   warps 0-7: consumers, 4(M) x 2(N) warp grid, 32 x 64 warp tile, 64 fp64 accumulators per lane.
//              Fragments come straight from the swizzled tile; the k index inside a stage is permuted
//              (k_phys = 2*ks + 8*(t>>1) + (t&1)) so that every half-warp 64-bit LDS is conflict-free.
//              diag(w) is applied to the A-side fragment in registers (w may be negative: never sqrt(w)).
//              A warp releases the stage with one mbarrier arrive ("empty").
// Work units = (128x128 lower-triangle tile, K split).  Units are ordered split-major so the CTAs running
// concurrently stream the same row range of A and share column panels through L2.  Each unit writes its tile
// to the split's private partial buffer (no atomics => deterministic); k_gram_finalize sums the splits in
// fixed order and mirrors the result.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace scs {

constexpr int kGT = 128;      // tile edge (columns of A)
constexpr int kGBK = 16;      // rows of A per pipeline stage (one 128-byte swizzle row)
constexpr int kGStages = 6;
constexpr int kGConsumerWarps = 8;
constexpr int kGThreads = (kGConsumerWarps + 1) * 32;
constexpr int kGTileBytes = kGT * kGBK * 8;  // 16 KB
constexpr int kGSmemBytes = kGStages * (2 * kGTileBytes + kGBK * 8) + 2 * kGStages * 8 + 1024;

SCS_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SCS_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SCS_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SCS_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SCS_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
SCS_DEVINL void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SCS_DEVINL void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
SCS_DEVINL double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
SCS_DEVINL void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GramPlan {
  int nt;        // tiles per edge = ceil(m/128)
  int ntiles;    // nt*(nt+1)/2
  int splits;    // K splits
  int64_t kt;    // total k-tiles = ldd/16
  int64_t units; // ntiles*splits
};

// tile index (row-major over the lower triangle) -> (bi >= bj)
SCS_DEVINL void tri_decode(int t, int& bi, int& bj) {
  int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  bi = r;
  bj = t - r * (r + 1) / 2;
}

__global__ void __launch_bounds__(kGThreads, 1)
k_gram(const __grid_constant__ CUtensorMap amap, const double* __restrict__ w, int m, int ldp, GramPlan plan,
       int64_t kt0 /* first 16-row k-tile of the active row window */,
       double* __restrict__ partial /* [splits][m*ldp], element (jc,kc) at jc*ldp+kc */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* tilesA = smem;
  uint8_t* tilesB = smem + kGStages * kGTileBytes;
  double* wt = (double*)(smem + 2 * kGStages * kGTileBytes);
  uint64_t* full = (uint64_t*)(wt + kGStages * kGBK);
  uint64_t* empty = full + kGStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();

  if (warp == kGConsumerWarps) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = blockIdx.x; u < plan.units; u += gridDim.x) {
        const int split = (int)(u / plan.ntiles);
        int bi, bj;
        tri_decode((int)(u % plan.ntiles), bi, bj);
        const int64_t k0 = plan.kt * split / plan.splits, k1 = plan.kt * (split + 1) / plan.splits;
        const bool diag = bi == bj;
        const uint32_t bytes = (diag ? 1 : 2) * kGTileBytes + kGBK * 8;
        for (int64_t kt = k0; kt < k1; ++kt) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], bytes);
          tma_load_2d(tilesA + stage * kGTileBytes, &amap, &full[stage], (int)((kt0 + kt) * kGBK), bi * kGT);
          if (!diag) tma_load_2d(tilesB + stage * kGTileBytes, &amap, &full[stage], (int)((kt0 + kt) * kGBK), bj * kGT);
          bulk_load_1d(wt + stage * kGBK, w + (kt0 + kt) * kGBK, kGBK * 8, &full[stage]);
          if (++stage == kGStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 32;   // warp tile origin (M) inside the CTA tile
  const int wn = (warp >> 2) * 64;  // (N)
  // per-lane constant parts of the swizzled fragment address: column c = base + 8*i + g  =>  c & 7 == g, so the
  // 16-byte chunk index of k_phys = 2*ks + 8*(t>>1) + (t&1) inside a 128-byte column row is (ks + 4*(t>>1)) ^ g
  const uint32_t sA = smem_u32(tilesA) + (uint32_t)((wm + g) * 128 + ((t & 1) << 3));
  const uint32_t sB = smem_u32(tilesB) + (uint32_t)((wn + g) * 128 + ((t & 1) << 3));
  const uint32_t sW = smem_u32(wt) + (uint32_t)((8 * (t >> 1) + (t & 1)) * 8);
  const uint32_t cx = (uint32_t)((4 * (t >> 1)) ^ g);  // chunk index for ks = 0; ks only touches bits 0-1
  int stage = 0;
  uint32_t phase = 0;
  for (int64_t u = blockIdx.x; u < plan.units; u += gridDim.x) {
    const int split = (int)(u / plan.ntiles);
    int bi, bj;
    tri_decode((int)(u % plan.ntiles), bi, bj);
    const int64_t k0 = plan.kt * split / plan.splits, k1 = plan.kt * (split + 1) / plan.splits;
    const bool diag = bi == bj;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int64_t kt = k0; kt < k1; ++kt) {
      mbar_wait(&full[stage], phase);
      const uint32_t ta = sA + (uint32_t)(stage * kGTileBytes);
      const uint32_t tb = diag ? (sA - (uint32_t)(wm * 128) + (uint32_t)(wn * 128) + (uint32_t)(stage * kGTileBytes))
                               : sB + (uint32_t)(stage * kGTileBytes);
      const uint32_t tw = sW + (uint32_t)(stage * kGBK * 8);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = (cx ^ (uint32_t)ks) << 4;
        const double wv = lds_f64(tw + ks * 16);
        double af[4], bf[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = lds_f64(ta + i * 1024 + off) * wv;
#pragma unroll
        for (int j = 0; j < 8; ++j) bf[j] = lds_f64(tb + j * 1024 + off);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kGStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    // epilogue: C[mrow][ncol] -> partial[split][(bi*128+mrow)*ldp + bj*128+ncol]
    double* P = partial + (int64_t)split * m * ldp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jc = bi * kGT + wm + 8 * i + g;
      if (jc < m) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kc = bj * kGT + wn + 8 * j + 2 * t;
          double* dst = P + (int64_t)jc * ldp + kc;
          if (kc + 1 < m) {
            *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
          } else if (kc < m) {
            dst[0] = acc[i][j][0];
          }
        }
      }
    }
  }
}

// G (m x m column-major, ld = m, both triangles) = sum over splits of the lower-triangle partials.
// 32x32 tiles; block (32,8).  Only tiles with tile-row >= tile-col are launched (grid.x enumerates them).
__global__ void __launch_bounds__(256)
k_gram_finalize(const double* __restrict__ partial, int splits, int m, int ldp, double* __restrict__ G) {
  __shared__ double tile[32][33];
  int tr, tc;
  tri_decode(blockIdx.x, tr, tc);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t pstride = (int64_t)m * ldp;
  for (int rr = ty; rr < 32; rr += 8) {
    const int jc = tr * 32 + rr, kc = tc * 32 + tx;
    double s = 0.0;
    if (jc < m && kc < m && jc >= kc) {
      const double* p = partial + (int64_t)jc * ldp + kc;
      for (int q = 0; q < splits; ++q) s += p[q * pstride];
      G[(int64_t)jc * m + kc] = s;  // upper triangle entry (row kc, col jc), coalesced along kc
    }
    tile[rr][tx] = s;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int kc = tc * 32 + cc, jc = tr * 32 + tx;
    if (jc < m && kc < m && jc >= kc) G[(int64_t)kc * m + jc] = tile[tx][cc];  // lower entry (row jc, col kc)
  }
}

// G[j,j] += lam*hr[j]   (λ.*Diagonal(Hr_diag), prox-GGN-SCORE.jl:129 / prox-N-SCORE.jl:43,70)
__global__ void k_add_diag(double* __restrict__ G, int m, double lam, const double* __restrict__ hr) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) G[(int64_t)j * m + j] += lam * hr[j];
}

}  // namespace scs
