// K2' — emulated-FP64 Gram on the 5th-generation tensor cores (tcgen05 int8, exact int32 accumulation in TMEM).
//
// fp64 has no tcgen05.mma kind, so  G = A' diag(w) A  (prox-GGN-SCORE.jl:129, prox-N-SCORE.jl:63) is evaluated
// through exact integer arithmetic (Ozaki scheme II / CRT):
//
//   1. C = diag(sqrt(w)) A is brought to fixed point per column:  X_ij = rint(C_ij * s_j),  s_j = T / (sqrt(wmax) ||A_j||_2),
//      so ||X_j||_2 <= T for every column and |sum_i X_ij X_ik| <= T^2 (Cauchy-Schwarz).  Needs w >= 0 (Newton weights,
//      consistent-label GGN weights, least squares); otherwise the caller stays on the DMMA kernel.
//   2. k_residues writes, for each of the first nmod (10..15) pairwise-coprime moduli p_l <= 256, the symmetric residue plane
//      x_l = X mod p_l as int8 (column-major, K = rows contiguous).
//   3. k_i8syrk computes S_l = x_l' x_l mod p_l on lower-triangle 128x256 tiles with tcgen05.mma.kind::i8
//      (TMA SWIZZLE_128B operand tiles -> 4-stage mbarrier ring -> UMMA 128x256x32, int32 accumulators in TMEM,
//      double-buffered), one K chunk (<= 65536 rows: |acc| <= 2^30) at a time; the epilogue warps pull the
//      accumulator with tcgen05.ld, reduce mod p_l and store int8 partial residues.
//   4. k_crt sums the chunk residues, reconstructs R = sum_i X_ij X_ik exactly by CRT in 128-bit integers
//      (P = prod p_l > 2 T^2) and writes G_jk = R / (s_j s_k) to both triangles.
//
// The only rounding is the fixed-point quantisation of C and the final conversion to fp64; the integer Gram itself is
// exact and bit-reproducible.  The quantisation error of an entry is ~0.4 / T of the diagonal scale (standard deviation, for
// w = const; the maximum over all entries is a few times that) whatever n is (12 moduli: T = 2^46.9 -> 3e-15; 13: 2e-16,
// i.e. fp64 rounding level).  The host picks the shortest moduli prefix whose
// T leaves every column at least the requested bits below its largest entry (default 38: 12 moduli for Gaussian-like
// columns at n = 1e6), then uses all of that prefix's range.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels_gram.cuh"  // mbarrier / TMA helpers, tri_decode

namespace scs {

#include "i8_moduli.inc"

constexpr int kI8BM = 128;       // tile rows   (UMMA M)
constexpr int kI8BN = 256;       // tile cols   (UMMA N)
constexpr int kI8BK = 128;       // K bytes per stage = one 128-byte swizzle row = 4 UMMA k-steps of 32
constexpr int kI8Stages = 4;
constexpr int kI8ABytes = kI8BM * kI8BK;  // 16 KB
constexpr int kI8BBytes = kI8BN * kI8BK;  // 32 KB
constexpr int kI8StageBytes = kI8ABytes + kI8BBytes;
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024 + 256;
constexpr int kI8Threads = 192;           // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int kI8ChunkRows = 65536;       // 128*128*65536 = 2^30 < 2^31
constexpr uint32_t kI8TmemCols = 512;     // two 256-column int32 accumulators
#ifndef SCS_I8_CLUSTER
#define SCS_I8_CLUSTER 4
#endif
#ifndef SCS_I8_SEGKB
#define SCS_I8_SEGKB 64
#endif
constexpr int kI8Cluster = SCS_I8_CLUSTER;             // CTAs per cluster: 4 vertically adjacent tiles, B slab multicast (64 rows each)
constexpr int kI8BPart = kI8BN / kI8Cluster;  // rows of the B slab each CTA fetches and multicasts
constexpr int kI8SegKb = SCS_I8_SEGKB;              // producers re-align every 64 k-blocks (8192 rows): keeps all clusters inside
                                          // one L2-sized window of the plane so operand panels are shared, not re-read

struct I8Plan {
  int m;
  int nmod;
  int nchunks;
  int ntiles;             // tile groups: kI8Cluster vertically adjacent 128x256 tiles that share one B slab
  int64_t kb_lo;          // first 128-row k-block of the active row window
  int64_t kblocks;        // one past its last k-block (<= ldx / 128)
  int64_t chunk_kblocks;  // kI8ChunkRows / 128
  int64_t units;          // nmod * nchunks * ntiles
  int ldp;                // bytes per row of a partial-residue matrix
  // partial-residue buffer: [nmod][pchunks][m][ldp]; this launch writes chunk slots pchunk0 .. pchunk0 + nchunks - 1.
  // (signed weights: the compacted minority-sign rows are a second launch into the slots after the main ones)
  int pchunks, pchunk0;
  int main_chunks;        // k_crt: slots [0, main_chunks) carry weight sign_main, the rest sign_extra
  int sign_main, sign_extra;
  int lock_slack;         // lock-step: a producer may start segment t once all have finished issuing segment t - 1 - lock_slack
};

// ---- column / row statistics -------------------------------------------------------------------------------
// colmax[j] = max_i |A_ij|, colnorm2[j] = sum_i A_ij^2  (one pass over A, once per problem; fixed reduction tree)
__global__ void __launch_bounds__(256) k_colabsmax(const double* __restrict__ A, int64_t ldd, int64_t n, int m,
                                                   double* __restrict__ colmax, double* __restrict__ colnorm2) {
  __shared__ double red[32], red2[32];
  const int j = blockIdx.x;
  const double* col = A + (int64_t)j * ldd;
  double v = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x * 2; i < n; i += 512) {
    const double2 a = ldg_stream2(col + i);
    const double ay = i + 1 < n ? a.y : 0.0;
    v = fmax(v, fmax(fabs(a.x), fabs(ay)));
    s2 = fma(a.x, a.x, s2);
    s2 = fma(ay, ay, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = v;
    red2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) {
      v = fmax(v, red[q]);
      s2 += red2[q];
    }
    colmax[j] = v;
    colnorm2[j] = s2;
  }
}
// ---- weight statistics, entirely on the device (no host round trip per Gram) -----------------------------------
// The rows are cut into blocks of kWsRows = 2048 (= one k_residues CTA: 256 threads x 8 rows).
//   k_wstat_part : part[4b + {0: max |w|, 1: # rows with w < 0, 2: 1 if a weight is not finite}] for block b
//   k_wstat_fin  : wstat[0] = max |w|, wstat[1] = # negative rows, wstat[2] = non-finite flag;
//                  negbase[b] = # negative rows in blocks < b (exclusive scan: where block b's compacted rows start)
constexpr int kWsRows = 2048;
enum { WS_MAXABS = 0, WS_NNEG = 1, WS_BAD = 2, WS_COUNT = 4 };
__global__ void __launch_bounds__(256) k_wstat_part(const double* __restrict__ w, int64_t nproc,
                                                    double* __restrict__ part) {
  __shared__ double smax[8], sneg[8], sbad[8];
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  double mx = 0.0, neg = 0.0, bad = 0.0;
  if (i0 < nproc) {  // nproc is a multiple of 16
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const double2 v = *reinterpret_cast<const double2*>(w + i0 + q);
      mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
      neg += (v.x < 0.0 ? 1.0 : 0.0) + (v.y < 0.0 ? 1.0 : 0.0);
      if (!(fabs(v.x) < 1.0e300) || !(fabs(v.y) < 1.0e300)) bad = 1.0;  // NaN or Inf (fmax drops NaNs)
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    neg += __shfl_xor_sync(0xffffffffu, neg, o);
    bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = mx;
    sneg[threadIdx.x >> 5] = neg;
    sbad[threadIdx.x >> 5] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) {
      mx = fmax(mx, smax[q]);
      neg += sneg[q];
      bad = fmax(bad, sbad[q]);
    }
    part[4 * blockIdx.x + 0] = mx;
    part[4 * blockIdx.x + 1] = neg;
    part[4 * blockIdx.x + 2] = bad;
  }
}
__global__ void __launch_bounds__(kVecThreads) k_wstat_fin(const double* __restrict__ part, int nblk,
                                                           double* __restrict__ wstat, int64_t* __restrict__ negbase) {
  __shared__ double smax[32], sbad[32];
  __shared__ long long ssum[32];
  const int per = (nblk + kVecThreads - 1) / kVecThreads;  // consecutive blocks per thread
  const int b0 = threadIdx.x * per, b1 = min(nblk, b0 + per);
  double mx = 0.0, bad = 0.0;
  long long cnt = 0;
  for (int b = b0; b < b1; ++b) {
    mx = fmax(mx, part[4 * b]);
    cnt += (long long)part[4 * b + 1];
    bad = fmax(bad, part[4 * b + 2]);
  }
  // exclusive scan of cnt over the threads
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  if (lane == 31) ssum[wid] = inc;
  if (lane == 0) {
    smax[wid] = mx;
    sbad[wid] = bad;
  }
  __syncthreads();
  long long base = 0, total = 0;
  for (int q = 0; q < kVecThreads / 32; ++q) {
    if (q < wid) base += ssum[q];
    total += ssum[q];
  }
  long long run = base + inc - cnt;
  for (int b = b0; b < b1; ++b) {
    negbase[b] = run;
    run += (long long)part[4 * b + 1];
  }
  if (threadIdx.x == 0) {
    for (int q = 1; q < kVecThreads / 32; ++q) {
      mx = fmax(mx, smax[q]);
      bad = fmax(bad, sbad[q]);
    }
    wstat[WS_MAXABS] = mx;
    wstat[WS_NNEG] = (double)total;
    wstat[WS_BAD] = bad;
  }
}
// Norm-equalised fixed point: scale_j = T / (sqrt(wmax) * ||A_j||_2), so that every column of X = rint(C scale) has
// ||X_j||_2 <= T (C = diag(sqrt |w|) A, |w| <= wmax = stat[WS_MAXABS]) and, by Cauchy-Schwarz, every entry of X'X is below T^2 < P/2: the CRT
// range is spent on precision, not on the worst case n * max^2.  inv_j = 1 / scale_j undoes it after the CRT.
__global__ void k_colscale(const double* __restrict__ colnorm2, const double* __restrict__ stat, int m, double T,
                           double* __restrict__ inv, double* __restrict__ scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double bound = sqrt(fmax(stat[0], 0.0)) * sqrt(colnorm2[j]);
  const bool ok = bound > 0.0 && bound < 1e300;
  scale[j] = ok ? T / bound : 0.0;  // all-zero column (or no weight): X = 0
  inv[j] = ok ? bound / T : 0.0;
}

// ---- residue planes ------------------------------------------------------------------------------------------
// planes[l][i + j*ldx] = (rint(sqrt(w_i) * A_ij * scale_j)) mod p_l  (symmetric, int8).  Thread = 8 rows x 1 column.
// All roundings / conversions use the 1.5*2^52 magic constant (DADD/DFMA only, no XU-pipe conversions):
//   X  = (v + M) - M                      v rounded to the nearest integer, |X| <= 2^50
//   q  = fma(X, 1/p, M) - M               nearest integer to X/p (off by at most one, only next to a half-way point)
//   r  = fma(-q, p, X)                    exact, |r| <= p/2 (+1 for p = 256 at a tie) -> its low byte is a valid residue
//   lo32(r + M)                           two's-complement bits of r
constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52
// SIGNED (weights of both signs, e.g. the README's +-1-label cross-entropy pair): X is built from sqrt(|w|), so
//   G = sum_i w_i a_i a_i' = X'X - 2 * Xc'Xc   (Xc = the rows of X whose weight is negative),  or, when most rows are
// negative,  G = -X'X + 2 * Xc'Xc  with Xc = the non-negative rows: the MINORITY sign is compacted.  Besides its plane
// entries every thread writes its minority rows, byte by byte, to the compacted planes at
//   cbase(block) + (# minority rows of the block before this thread) + ...   (consecutive threads -> consecutive bytes).
template <int NMOD, bool SIGNED>
__global__ void __launch_bounds__(256)
k_residues(const double* __restrict__ A, int64_t ldd, int64_t nproc, int m, const double* __restrict__ w,
           const double* __restrict__ scale, int8_t* __restrict__ planes, int64_t ldx,
           const int64_t* __restrict__ negbase, int minor_neg, int8_t* __restrict__ cplanes, int64_t ldc) {
  __shared__ double s_p[NMOD], s_ip[NMOD];
  __shared__ int s_wsum[8];
  if (threadIdx.x < NMOD) {
    s_p[threadIdx.x] = (double)c_mod_p[threadIdx.x];
    s_ip[threadIdx.x] = 1.0 / (double)c_mod_p[threadIdx.x];
  }
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  const bool live = i0 < nproc;  // nproc is a multiple of 8; rows outside the active window have w = 0
  double sw[8];
  uint32_t cmask = 0;  // SIGNED: which of this thread's rows go to the compacted planes
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const double2 ww = live ? *reinterpret_cast<const double2*>(w + i0 + q) : make_double2(0.0, 0.0);
    sw[q] = sqrt(SIGNED ? fabs(ww.x) : ww.x);
    sw[q + 1] = sqrt(SIGNED ? fabs(ww.y) : ww.y);
    if (SIGNED && live) {
      cmask |= ((ww.x < 0.0) == (minor_neg != 0) ? 1u : 0u) << q;
      cmask |= ((ww.y < 0.0) == (minor_neg != 0) ? 1u : 0u) << (q + 1);
    }
  }
  int64_t cpos = 0;  // first compacted row of this thread
  if (SIGNED) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int mine = __popc(cmask);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    int before = inc - mine;
    for (int q = 0; q < wid; ++q) before += s_wsum[q];
    const int64_t nb = negbase[blockIdx.x];  // negative rows in earlier blocks
    cpos = (minor_neg ? nb : (int64_t)blockIdx.x * kWsRows - nb) + before;
  } else {
    __syncthreads();
  }
  if (!live) return;
  const int64_t plane_stride = ldx * (int64_t)m;
  const int64_t cplane_stride = ldc * (int64_t)m;
  for (int j = blockIdx.y; j < m; j += gridDim.y) {
    const double sc = scale[j];
    const double* col = A + (int64_t)j * ldd + i0;
    double X[8];
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const double2 a = ldg_stream2(col + q);
      X[q] = ((sw[q] * a.x) * sc) + kMagic;  // kept biased: XM = X + M (exact, |X| <= 2^50)
      X[q + 1] = ((sw[q + 1] * a.y) * sc) + kMagic;
    }
    int8_t* dst = planes + (int64_t)j * ldx + i0;
    int8_t* cdst = SIGNED ? cplanes + (int64_t)j * ldc + cpos : nullptr;
#pragma unroll
    for (int l = 0; l < NMOD; ++l) {
      const double p = s_p[l], ip = s_ip[l];
      uint32_t b[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double x = X[q] - kMagic;
        const double qd = fma(x, ip, kMagic) - kMagic;
        b[q] = (uint32_t)__double2loint(fma(-qd, p, X[q]));  // = M + r exactly; the low word holds r
      }
      // gather the low bytes of eight words into two
      const uint32_t lo = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
      const uint32_t hi = __byte_perm(__byte_perm(b[4], b[5], 0x0040), __byte_perm(b[6], b[7], 0x0040), 0x5410);
      *reinterpret_cast<uint2*>(dst + (int64_t)l * plane_stride) = make_uint2(lo, hi);
      if (SIGNED && cmask) {
        int8_t* cd = cdst + (int64_t)l * cplane_stride;
        int k = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if ((cmask >> q) & 1u) cd[k++] = (int8_t)(b[q] & 0xffu);
      }
    }
  }
}
// zero the tail of the compacted planes up to the next 128-row k-block: rows [cnt, round_up(cnt, 128))
__global__ void __launch_bounds__(128) k_cpad(int8_t* __restrict__ cplanes, int64_t ldc, int m, int64_t cnt) {
  const int64_t pos = cnt + threadIdx.x;
  const int64_t end = (cnt + 127) / 128 * 128;
  if (pos < end) cplanes[((int64_t)blockIdx.y * m + blockIdx.x) * ldc + pos] = 0;
}

// ---- tcgen05 helpers -------------------------------------------------------------------------------------------
SCS_DEVINL void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
SCS_DEVINL void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
SCS_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SCS_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
SCS_DEVINL void tc_commit_mc(uint64_t* bar, uint16_t mask) {  // arrive on the same barrier in every CTA of `mask`
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
SCS_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SCS_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SCS_DEVINL void tc_commit(uint64_t* bar) {  // arrives on the mbarrier when all previously issued MMAs have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, N = 256, K = 32
SCS_DEVINL void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart (cute::UMMA::SmemDescriptor)
SCS_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) /* LBO (unused for swizzled K-major) */ |
         (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* descriptor version (sm_100) */ |
         (2ull << 61) /* SWIZZLE_128B */;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32, A = B = INT8, both K-major, N = 256, M = 128
constexpr uint32_t kI8Idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kI8BN >> 3) << 17) | ((uint32_t)(kI8BM >> 4) << 24);
// the same with N = 128: tiles on the diagonal whose right half lies above it
constexpr uint32_t kI8IdescHalf = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((kI8BN / 2) >> 3) << 17) | ((uint32_t)(kI8BM >> 4) << 24);
// Columns of the 128 x 256 tile of CTA `crank` of tile group (g, bj) that reach the lower triangle (col <= last row of the
// tile): 0 = the whole tile lies above the diagonal (the CTA only relays its share of the B slab), <= 128 = N = 128 UMMAs.
SCS_DEVINL int i8_live_cols(int2 tile, int crank) {
  const int r1 = (tile.x * kI8Cluster + crank) * kI8BM + kI8BM;  // one past the last row
  const int c0 = tile.y * kI8BN;
  return max(0, min(kI8BN, r1 - c0));
}

SCS_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// unit -> (modulus l, chunk c, tile t): all tiles of one (l, c) are adjacent so concurrently running CTAs stream the
// same rows of the same plane and share operand panels through L2
SCS_DEVINL void i8_unit(const I8Plan& pl, int64_t u, int& l, int& c, int& t) {
  t = (int)(u % pl.ntiles);
  const int64_t lc = u / pl.ntiles;
  c = (int)(lc % pl.nchunks);
  l = (int)(lc / pl.nchunks);
}

// Launched as clusters of kI8Cluster CTAs.  CTA rank r of a cluster owns tile (tiles[t].x * kI8Cluster + r, tiles[t].y);
// it fetches its own A slab and rows [r*kI8BPart, (r+1)*kI8BPart) of the shared B slab, multicast to the whole cluster.
__global__ void __launch_bounds__(kI8Threads, 1)
k_i8syrk(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap bmap, I8Plan pl,
         const int2* __restrict__ tiles, int8_t* __restrict__ partial /* [nmod][nchunks][m][ldp] */,
         unsigned long long* __restrict__ progress /* zeroed before the launch */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + kI8Stages * kI8StageBytes);
  uint64_t* empty = full + kI8Stages;
  uint64_t* acc_full = empty + kI8Stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int64_t cid = blockIdx.x / kI8Cluster, ncl = gridDim.x / kI8Cluster;
  constexpr uint16_t kMask = (uint16_t)((1u << kI8Cluster) - 1u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kI8Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kI8Cluster);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kI8TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before any peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    // All CTAs are co-resident (the grid is sized by cudaOccupancyMaxActiveClusters), so the producers can keep in
    // step through a global counter: before issuing segment t every producer has finished issuing segment t-1.
    // The wait has a bounded spin: if the assumption were ever violated the kernel degrades to free-running.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int64_t steps = (pl.units + ncl - 1) / ncl;
      const int segs = (int)((pl.chunk_kblocks + kI8SegKb - 1) / kI8SegKb);
      const unsigned long long ncta = gridDim.x;
      bool in_step = true;
      for (int64_t st = 0; st < steps; ++st) {
        const int64_t u = cid + st * ncl;
        const bool valid = u < pl.units;
        int l = 0, c = 0, t = 0;
        if (valid) i8_unit(pl, u, l, c, t);
        const int2 tile = valid ? tiles[t] : make_int2(0, 0);
        const bool live = i8_live_cols(tile, crank) > 0;
        const int64_t kb0 = pl.kb_lo + (int64_t)c * pl.chunk_kblocks;
        const int64_t kb1 = kb0 + pl.chunk_kblocks < pl.kblocks ? kb0 + pl.chunk_kblocks : pl.kblocks;
        for (int sg = 0; sg < segs; ++sg) {
          const unsigned long long point = (unsigned long long)(st * segs + sg);
          if (point > 0 && progress != nullptr) {  // progress == nullptr: free-running (tuning aid SCS_I8_NOLOCK=1)
            atomicAdd(progress, 1ULL);
            if (in_step) {
              const unsigned long long want = point > (unsigned long long)pl.lock_slack ? (point - pl.lock_slack) * ncta : 0ULL;
              int spins = 0;
              while (true) {
                unsigned long long seen;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(progress) : "memory");
                if (seen >= want) break;
                if (++spins > 40000) {  // ~4 ms: give up on lock-step, keep computing
                  in_step = false;
                  break;
                }
                __nanosleep(100);
              }
            }
          }
          if (!valid) continue;
          const int64_t s0 = kb0 + (int64_t)sg * kI8SegKb;
          const int64_t s1 = s0 + kI8SegKb < kb1 ? s0 + kI8SegKb : kb1;
          for (int64_t kb = s0; kb < s1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);  // all CTAs of the cluster have consumed this slot
            mbar_expect_tx(&full[stage], live ? kI8StageBytes : kI8BBytes);
            uint8_t* sa = smem + stage * kI8StageBytes;
            if (live) tma_load_3d(sa, &xmap, &full[stage], (int)(kb * kI8BK), (tile.x * kI8Cluster + crank) * kI8BM, l);
            tma_load_3d_mc(sa + kI8ABytes + crank * (kI8BPart * kI8BK), &bmap, &full[stage], (int)(kb * kI8BK),
                           tile.y * kI8BN + crank * kI8BPart, l, kMask);
            if (++stage == kI8Stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int64_t u = cid; u < pl.units; u += ncl) {
        int l, c, t;
        i8_unit(pl, u, l, c, t);
        const int ncols = i8_live_cols(tiles[t], crank);
        const uint32_t idesc = ncols > kI8BN / 2 ? kI8Idesc : kI8IdescHalf;
        const int64_t kb0 = pl.kb_lo + (int64_t)c * pl.chunk_kblocks;
        const int64_t kb1 = kb0 + pl.chunk_kblocks < pl.kblocks ? kb0 + pl.chunk_kblocks : pl.kblocks;
        mbar_wait(&acc_empty[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * kI8BN);
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kI8StageBytes);
          const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kI8ABytes);
          if (ncols > 0) {  // a tile above the diagonal issues nothing: the commit below releases the slot at once
#pragma unroll
            for (int k = 0; k < kI8BK / 32; ++k)
              umma_i8(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit_mc(&empty[stage], kMask);  // tells every producer of the cluster that this CTA is done with the slot
          if (++stage == kI8Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&acc_full[as]);  // accumulator complete
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> mod p -> int8 partial residues =====
    const int quarter = warp & 3;  // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int64_t u = cid; u < pl.units; u += ncl) {
      int l, c, t;
      i8_unit(pl, u, l, c, t);
      const int2 tile = tiles[t];
      const int p = c_mod_p[l];
      const double pd = (double)p, ip = 1.0 / pd;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const int jc = (tile.x * kI8Cluster + crank) * kI8BM + row_in_tile;
      int8_t* prow = partial + (((int64_t)l * pl.pchunks + pl.pchunk0 + c) * pl.m + jc) * pl.ldp + (int64_t)tile.y * kI8BN;
      const uint32_t taddr = tmem_base + (uint32_t)(as * kI8BN) + ((uint32_t)(quarter * 32) << 16);
      const int ncc = (i8_live_cols(tile, crank) + 31) / 32;  // column chunks that reach the lower triangle (k_crt reads kc <= jc)
#pragma unroll 1
      for (int cc = 0; cc < ncc; ++cc) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(cc * 32), v);
        uint32_t packed[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t word = 0;
#pragma unroll
          for (int bq = 0; bq < 4; ++bq) {
            const double dv = (double)(int)v[4 * q + bq];
            int r = (int)fma(-rint(dv * ip), pd, dv);
            if (2 * r >= p) r -= p;
            if (2 * r < -p) r += p;
            word |= (uint32_t)(r & 0xff) << (8 * bq);
          }
          packed[q] = word;
        }
        const int kc0 = tile.y * kI8BN + cc * 32;
        if (jc < pl.m && kc0 < pl.ldp) {
          uint4* d4 = reinterpret_cast<uint4*>(prow + cc * 32);
          d4[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          if (kc0 + 16 < pl.ldp) d4[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still multicast into its shared memory / barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
  }
}

// ---- 2-CTA variant (opt-in: SCS_I8_2CTA=1; parity-tested, measured equal to k_i8syrk under the 1 kW power cap) ---
// Motivation: in k_i8syrk the shared-memory port is busier than the tensor pipe: at full rate a 128x256x32 UMMA (135 cycles)
// reads 12 KB of operands and its share of the TMA fill is another 12 KB — 178 B/cycle against the SM's 128 B/cycle,
// which caps the IMMA pipe at 72 % (ncu: 71.7 % active).  In cta_group::2 mode the two CTAs of a pair (ranks 2p, 2p+1
// of the cluster) issue one 256x256x32 UMMA: each CTA supplies its own 128 rows of A and HALF of the B slab
// (128 columns of G), so fill and reads per SM drop to 8 + 8 KB per UMMA (118 B/cycle).
//   * cluster of 4 = two pairs stacked vertically (512 rows x 256 columns of G, as before);
//   * B half q (= rank & 1) is needed by ranks q and q+2: each of them fetches 64 of its 128 rows and multicasts to both;
//   * only the pair leader (even rank) issues tcgen05.mma.cta_group::2; the peer's MMA warp relays "my stage is full"
//     to the leader's pfull barrier (remote mbarrier arrive); tcgen05.commit.cta_group::2 multicasts the stage release to
//     all four CTAs (both leaders must be done before a slot is refilled) and the accumulator hand-off to the pair;
//   * epilogue warps of both CTAs drain their own 128 TMEM lanes and arrive on the LEADER's acc_empty barrier.
// bounded wait: a protocol error in the pair hand-shakes traps (launch error) instead of hanging the device
SCS_DEVINL void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (ok) return;
    if (t0 == 0)
      t0 = clock64();
    else if (clock64() - t0 > 6000000000LL)
      __trap();
  }
}
constexpr int kI8Stages2 = 6;
constexpr int kI8BHalf = kI8BN / 2;                       // B rows (columns of G) held by one CTA of a pair
constexpr int kI8Stage2Bytes = kI8ABytes + kI8BHalf * kI8BK;  // 16 + 16 KB
constexpr int kI8Smem2Bytes = kI8Stages2 * kI8Stage2Bytes + 1024 + 512;
constexpr uint32_t kI8Idesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kI8BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

SCS_DEVINL void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
SCS_DEVINL void tc_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
SCS_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {  // arrive on the same barrier in CTA `rank`
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// Control-only signal (the data it announces was written by TMA into this CTA's own shared memory and is consumed by
// this SM's tensor core): relaxed, so that no cluster-scope fence (ERRBAR + CCTL.IVALL) is paid per pipeline stage.
SCS_DEVINL void mbar_arrive_remote_relaxed(uint32_t raddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

__global__ void __launch_bounds__(kI8Threads, 1)
k_i8syrk2(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap bmap, I8Plan pl,
          const int2* __restrict__ tiles, int8_t* __restrict__ partial /* [nmod][nchunks][m][ldp] */,
          unsigned long long* __restrict__ progress /* zeroed before the launch */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + kI8Stages2 * kI8Stage2Bytes);
  uint64_t* empty = full + kI8Stages2;
  uint64_t* pfull = empty + kI8Stages2;  // leader only: the peer's stage is full
  uint64_t* acc_full = pfull + kI8Stages2;
  uint64_t* acc_empty = acc_full + 2;    // leader only: both CTAs' epilogues have drained the accumulator
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int q = crank & 1;           // position inside the pair = which half of the B slab this CTA holds
  const int leader_rank = crank & ~1;
  const bool is_leader = q == 0;
  const int64_t cid = blockIdx.x / kI8Cluster, ncl = gridDim.x / kI8Cluster;
  constexpr uint16_t kMaskAll = (uint16_t)((1u << kI8Cluster) - 1u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kI8Stages2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kI8Cluster / 2);  // one tcgen05.commit arrival from each pair leader of the cluster
      mbar_init(&pfull[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);  // 4 epilogue warps of each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation for the pair: one warp of each CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kI8TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (lock-step with the other clusters, as in k_i8syrk) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int64_t steps = (pl.units + ncl - 1) / ncl;
      const int segs = (int)((pl.chunk_kblocks + kI8SegKb - 1) / kI8SegKb);
      const unsigned long long ncta = gridDim.x;
      bool in_step = true;
      const int h = crank >> 1;  // which 64 of the 128 rows of its B half this CTA fetches
      // the CTAs that hold half q: ranks q and q + 2 (cluster of 4 = two pairs), or just this one (cluster of 2 = one pair)
      const uint16_t bmask = (uint16_t)(((1u << q) | (1u << (q + 2))) & kMaskAll);
      for (int64_t st = 0; st < steps; ++st) {
        const int64_t u = cid + st * ncl;
        const bool valid = u < pl.units;
        int l = 0, c = 0, t = 0;
        if (valid) i8_unit(pl, u, l, c, t);
        const int2 tile = valid ? tiles[t] : make_int2(0, 0);
        const int64_t kb0 = pl.kb_lo + (int64_t)c * pl.chunk_kblocks;
        const int64_t kb1 = kb0 + pl.chunk_kblocks < pl.kblocks ? kb0 + pl.chunk_kblocks : pl.kblocks;
        for (int sg = 0; sg < segs; ++sg) {
          const unsigned long long point = (unsigned long long)(st * segs + sg);
          if (point > 0 && progress != nullptr) {  // progress == nullptr: free-running (tuning aid SCS_I8_NOLOCK=1)
            atomicAdd(progress, 1ULL);
            if (in_step) {
              const unsigned long long want = point > (unsigned long long)pl.lock_slack ? (point - pl.lock_slack) * ncta : 0ULL;
              int spins = 0;
              while (true) {
                unsigned long long seen;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(progress) : "memory");
                if (seen >= want) break;
                if (++spins > 40000) {
                  in_step = false;
                  break;
                }
                __nanosleep(100);
              }
            }
          }
          if (!valid) continue;
          const int64_t s0 = kb0 + (int64_t)sg * kI8SegKb;
          const int64_t s1 = s0 + kI8SegKb < kb1 ? s0 + kI8SegKb : kb1;
          for (int64_t kb = s0; kb < s1; ++kb) {
            mbar_wait_bounded(&empty[stage], phase ^ 1);  // both pairs of the cluster have consumed this slot
            mbar_expect_tx(&full[stage], kI8Stage2Bytes);
            uint8_t* sa = smem + stage * kI8Stage2Bytes;
            tma_load_3d(sa, &xmap, &full[stage], (int)(kb * kI8BK), (tile.x * kI8Cluster + crank) * kI8BM, l);
            tma_load_3d_mc(sa + kI8ABytes + h * (kI8BPart * kI8BK), &bmap, &full[stage], (int)(kb * kI8BK),
                           tile.y * kI8BN + q * kI8BHalf + h * kI8BPart, l, bmask);
            if (++stage == kI8Stages2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (is_leader) {
        // ===== MMA issuer for the pair (one thread) =====
        int as = 0;
        uint32_t aphase = 0;
        const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
        for (int64_t u = cid; u < pl.units; u += ncl) {
          int l, c, t;
          i8_unit(pl, u, l, c, t);
          const int64_t kb0 = pl.kb_lo + (int64_t)c * pl.chunk_kblocks;
          const int64_t kb1 = kb0 + pl.chunk_kblocks < pl.kblocks ? kb0 + pl.chunk_kblocks : pl.kblocks;
          mbar_wait_bounded(&acc_empty[as], aphase ^ 1);  // both epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(as * kI8BN);
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            mbar_wait_bounded(&full[stage], phase);
            mbar_wait_bounded(&pfull[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kI8Stage2Bytes);
            const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kI8ABytes);
#pragma unroll
            for (int k = 0; k < kI8BK / 32; ++k)
              umma2_i8(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kI8Idesc2, (kb > kb0 || k > 0) ? 1u : 0u);
            tc_commit2_mc(&empty[stage], kMaskAll);
            if (++stage == kI8Stages2) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit2_mc(&acc_full[as], pair_mask);  // accumulator complete: wake the epilogues of both CTAs
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
      } else {
        // ===== peer: tell the leader when this CTA's half of the stage has landed =====
        uint32_t pf_remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pf_remote) : "r"(smem_u32(pfull)), "r"((uint32_t)leader_rank));
        for (int64_t u = cid; u < pl.units; u += ncl) {
          int l, c, t;
          i8_unit(pl, u, l, c, t);
          const int64_t kb0 = pl.kb_lo + (int64_t)c * pl.chunk_kblocks;
          const int64_t kb1 = kb0 + pl.chunk_kblocks < pl.kblocks ? kb0 + pl.chunk_kblocks : pl.kblocks;
          for (int64_t kb = kb0; kb < kb1; ++kb) {
            mbar_wait_bounded(&full[stage], phase);
            mbar_arrive_remote_relaxed(pf_remote + 8u * (uint32_t)stage);
            if (++stage == kI8Stages2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> mod p -> int8 partial residues =====
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int64_t u = cid; u < pl.units; u += ncl) {
      int l, c, t;
      i8_unit(pl, u, l, c, t);
      const int2 tile = tiles[t];
      const int p = c_mod_p[l];
      const double pd = (double)p, ip = 1.0 / pd;
      mbar_wait_bounded(&acc_full[as], aphase);
      tc_fence_after();
      const int jc = (tile.x * kI8Cluster + crank) * kI8BM + row_in_tile;
      int8_t* prow = partial + (((int64_t)l * pl.pchunks + pl.pchunk0 + c) * pl.m + jc) * pl.ldp + (int64_t)tile.y * kI8BN;
      const uint32_t taddr = tmem_base + (uint32_t)(as * kI8BN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < kI8BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(cc * 32), v);
        uint32_t packed[8];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          uint32_t word = 0;
#pragma unroll
          for (int bq = 0; bq < 4; ++bq) {
            const double dv = (double)(int)v[4 * qq + bq];
            int r = (int)fma(-rint(dv * ip), pd, dv);
            if (2 * r >= p) r -= p;
            if (2 * r < -p) r += p;
            word |= (uint32_t)(r & 0xff) << (8 * bq);
          }
          packed[qq] = word;
        }
        const int kc0 = tile.y * kI8BN + cc * 32;
        if (jc < pl.m && kc0 < pl.ldp) {
          uint4* d4 = reinterpret_cast<uint4*>(prow + cc * 32);
          d4[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          if (kc0 + 16 < pl.ldp) d4[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (is_leader)
          mbar_arrive(&acc_empty[as]);
        else
          mbar_arrive_remote(&acc_empty[as], (uint32_t)leader_rank);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
  }
}

// ---- all-reduce of the packed Gram over NVLink peer memory (one node, one process per GPU) ---------------------------
// Every rank's packed triangle lives in an allocation the other ranks have mapped (cudaIpc); the exchange step is three
// small kernels of our own instead of an NCCL call:
//   k_p2p_barrier        everybody's triangle is complete (stream order locally, flags across ranks)
//   k_p2p_reduce_scatter rank r sums slice r of all ranks' triangles — loads straight from peer HBM over NVLink, always in
//                        rank order 0..W-1 — and writes the result into its own copy
//   k_p2p_barrier, k_p2p_allgather (every rank copies the other W-1 reduced slices), k_p2p_barrier
// Each entry is reduced exactly once, by one rank, in a fixed order: all ranks end with bitwise the same matrix (the
// replicated solve depends on that) and the result does not depend on timing.  Flags: receiver-owned, one 64-bit slot
// per sender, monotonically increasing epochs; system-scope release / acquire; bounded spins (a dead peer traps the
// kernel instead of hanging the device).
constexpr int kP2PMaxWorld = 16;
struct P2PPeers {
  double* base[kP2PMaxWorld];               // rank r's buffer as mapped in this process (own buffer for r == rank)
  unsigned long long* flags[kP2PMaxWorld];  // rank r's flag slots (receiver-owned: flags[r][sender])
};
__global__ void __launch_bounds__(32) k_p2p_barrier(P2PPeers pp, int rank, int world, unsigned long long epoch) {
  const int r = threadIdx.x;
  __threadfence_system();
  if (r < world && r != rank)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flags[r] + rank), "l"(epoch) : "memory");
  if (r < world && r != rank) {
    const long long t0 = clock64();
    while (true) {
      unsigned long long seen;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(pp.flags[rank] + r) : "memory");
      if (seen >= epoch) break;
      if (clock64() - t0 > 120000000000LL) __trap();  // ~60 s: a peer is gone
    }
  }
  __syncwarp();
  __threadfence_system();
}
// this rank's slice of [0, count): [count * rank / world, count * (rank + 1) / world) rounded to even offsets
SCS_DEVINL void p2p_slice(size_t count, int r, int world, size_t& lo, size_t& hi) {
  lo = (count * (size_t)r / (size_t)world) & ~(size_t)1;
  hi = r == world - 1 ? count : (count * (size_t)(r + 1) / (size_t)world) & ~(size_t)1;
}
__global__ void __launch_bounds__(256) k_p2p_reduce_scatter(P2PPeers pp, int rank, int world, size_t count) {
  size_t lo, hi;
  p2p_slice(count, rank, world, lo, hi);
  const size_t npair = (hi - lo) / 2;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < npair; i += (size_t)gridDim.x * 256) {
    const size_t e = lo + 2 * i;
    double2 v[kP2PMaxWorld];  // all peers' loads in flight before the first add (a remote load is a few microseconds)
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const double2*>(pp.base[r] + e));
    double2 acc = v[0];
#pragma unroll
    for (int r = 1; r < kP2PMaxWorld; ++r)
      if (r < world) {  // rank order: the same sum on every run
        acc.x += v[r].x;
        acc.y += v[r].y;
      }
    *reinterpret_cast<double2*>(pp.base[rank] + e) = acc;
  }
  if (((hi - lo) & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail (last slice only)
    const size_t e = hi - 1;
    double acc = __ldcg(pp.base[0] + e);
    for (int r = 1; r < world; ++r) acc += __ldcg(pp.base[r] + e);
    pp.base[rank][e] = acc;
  }
}
__global__ void __launch_bounds__(256) k_p2p_allgather(P2PPeers pp, int rank, int world, size_t count) {
  // blockIdx.y = which peer's slice (skipping this rank's own)
  const int r = (int)blockIdx.y >= rank ? (int)blockIdx.y + 1 : (int)blockIdx.y;
  size_t lo, hi;
  p2p_slice(count, r, world, lo, hi);
  const size_t npair = (hi - lo) / 2;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < npair; i += (size_t)gridDim.x * 256) {
    const size_t e = lo + 2 * i;
    *reinterpret_cast<double2*>(pp.base[rank] + e) = __ldcg(reinterpret_cast<const double2*>(pp.base[r] + e));
  }
  if (((hi - lo) & 1) && blockIdx.x == 0 && threadIdx.x == 0) pp.base[rank][hi - 1] = __ldcg(pp.base[r] + hi - 1);
}

// ---- int8 tensor-pipe peak (measurement aid for bench.py: the denominator of k_i8syrk's roofline) -----------------
// One CTA per SM issues the SAME 128x256x32 kind::i8 UMMA k_i8syrk issues, back to back, on operands that never leave
// shared memory: no TMA, no epilogue, no global traffic — what the tensor pipe sustains at the clock the power cap allows
// with realistic (pseudo-random, non-zero) operand bits.  Commits are double-buffered so the pipe never drains.
constexpr int kI8PeakSmem = kI8StageBytes + 1024 + 64;
__global__ void __launch_bounds__(128, 1) k_i8peak(long long groups /* of 64 x 4 UMMAs */, uint32_t* __restrict__ sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = (uint64_t*)(smem + kI8StageBytes);
  uint32_t* tmem_slot = (uint32_t*)(bar + 2);
  for (int i = threadIdx.x; i < kI8StageBytes / 4; i += 128)
    reinterpret_cast<uint32_t*>(smem)[i] = ((uint32_t)i * 2654435761u + blockIdx.x * 40503u) ^ ((uint32_t)i << 13);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // the generic-proxy fill above is read by the tensor core
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kI8TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kI8ABytes);
    uint32_t ph[2] = {0, 0};
    for (long long g = 0; g < groups; ++g) {
#pragma unroll 4
      for (int q = 0; q < 64; ++q) {
        const uint32_t tacc = tmem_base + (uint32_t)((q & 1) * kI8BN);
#pragma unroll
        for (int k = 0; k < kI8BK / 32; ++k) umma_i8(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kI8Idesc, 1u);
      }
      tc_commit(&bar[g & 1]);
      if (g > 0) {  // wait for the PREVIOUS group: at most two groups (512 UMMAs) are ever queued
        mbar_wait(&bar[(g - 1) & 1], ph[(g - 1) & 1]);
        ph[(g - 1) & 1] ^= 1u;
      }
    }
    if (groups > 0) mbar_wait(&bar[(groups - 1) & 1], ph[(groups - 1) & 1]);
    tc_fence_after();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base, v);
    if (sink && v[threadIdx.x & 31] == 0x7fffffffu) sink[0] = v[0];  // keeps the accumulator observable
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
}

// pseudo-random bytes (the toggle rate of the operand bits decides the tensor-pipe power, hence the clock under the cap)
__global__ void k_fill_random_bytes(uint32_t* __restrict__ p, size_t nwords) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i * 2654435761u + 0x9e3779b9u;
    v ^= v >> 15;
    v *= 0x85ebca6bu;
    v ^= v >> 13;
    p[i] = v;
  }
}

// ---- pipeline probe (tuning aid): k_i8syrk's main loop on an L2-resident operand ------------------------------------
// The same 4-stage TMA -> mbarrier -> UMMA ring as k_i8syrk (one CTA per SM, no cluster, no multicast, no epilogue, no
// lock-step), reading a 24 MB matrix that stays in L2.  tma_mode: 0 = no TMA at all (operands filled once; only the
// per-stage commit / slot hand-shake remains), 1 = the A slab (16 KB per stage) comes through TMA, 2 = A and B (48 KB per
// stage, what an unclustered CTA would move).  Separates what the SM-level pipeline can sustain from DRAM / L2-miss /
// multicast / lock-step effects.
__global__ void __launch_bounds__(128, 1)
k_i8pipe(const __grid_constant__ CUtensorMap map /* dims {K bytes, 384 rows, 1}, box {128, 128, 1} */, int tma_mode,
         long long iters, int kspan /* k-blocks in the matrix */, uint32_t* __restrict__ sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + kI8Stages * kI8StageBytes);
  uint64_t* empty = full + kI8Stages;
  uint32_t* tmem_slot = (uint32_t*)(empty + kI8Stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kI8Stages * kI8StageBytes / 4; i += 128)
    reinterpret_cast<uint32_t*>(smem)[i] = ((uint32_t)i * 2654435761u + blockIdx.x * 40503u) ^ ((uint32_t)i << 13);
  if (threadIdx.x == 0) {
    for (int q = 0; q < kI8Stages; ++q) {
      mbar_init(&full[q], 1);
      mbar_init(&empty[q], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kI8TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0 && tma_mode > 0) {
    int stage = 0;
    uint32_t phase = 0;
    int kb = (int)((blockIdx.x * 37u) % (unsigned)kspan);
    for (long long it = 0; it < iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sa = smem + stage * kI8StageBytes;
      mbar_expect_tx(&full[stage], tma_mode == 2 ? kI8StageBytes : kI8ABytes);
      tma_load_3d(sa, &map, &full[stage], kb * kI8BK, 0, 0);
      if (tma_mode == 2) {
        tma_load_3d(sa + kI8ABytes, &map, &full[stage], kb * kI8BK, 128, 0);
        tma_load_3d(sa + kI8ABytes + kI8ABytes, &map, &full[stage], kb * kI8BK, 256, 0);
      }
      if (++kb == kspan) kb = 0;
      if (++stage == kI8Stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (long long it = 0; it < iters; ++it) {
      if (tma_mode > 0) {
        mbar_wait(&full[stage], phase);
      } else if (it >= kI8Stages) {
        mbar_wait(&empty[stage], phase ^ 1);  // the slot's previous group has retired: at most 4 groups are queued
      }
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kI8StageBytes);
      const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kI8ABytes);
      const uint32_t tacc = tmem_base + (uint32_t)((it & 1) * kI8BN);
#pragma unroll
      for (int k = 0; k < kI8BK / 32; ++k) umma_i8(tacc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kI8Idesc, 1u);
      tc_commit(&empty[stage]);
      if (++stage == kI8Stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    // drain: the last group's commit
    const int last = (int)((iters - 1) % kI8Stages);
    const uint32_t lph = (uint32_t)(((iters - 1) / kI8Stages) & 1);
    if (iters > 0) mbar_wait(&empty[last], lph);
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 1) {
    uint32_t v[32];
    tmem_ld32(tmem_base, v);
    if (sink && v[lane] == 0x7fffffffu) sink[0] = v[0];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
}

// ---- CRT reconstruction -----------------------------------------------------------------------------------------
// For each lower-triangle (jc >= kc): R = CRT({sum_c partial[l][c][jc][kc] mod p_l}) in (-P/2, P/2), then
// G[jc,kc] = G[kc,jc] = R * inv_jc * inv_kc.  One thread reconstructs 4 consecutive kc (32-bit loads of the
// int8 partial residues).
__global__ void __launch_bounds__(256)
k_crt(const int8_t* __restrict__ partial, I8Plan pl, const double* __restrict__ inv, const double* __restrict__ wstat,
      int expect_nonneg, double* __restrict__ G, int jc_lo, int jc_hi, double* __restrict__ Gpack) {
  // rows [jc_lo, jc_hi) of the lower triangle (several ranks: the Gram is reconstructed slab by slab so that the
  // all-reduce of one slab overlaps the CRT of the next).  Gpack != null: write the packed upper triangle
  // Gpack[jc (jc+1)/2 + kc] = G[kc, jc] (kc <= jc) instead of the two full triangles — half the all-reduce volume.
  const int kc0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int jc = jc_lo + blockIdx.y * 4 + (threadIdx.x >> 6);
  if (jc >= jc_hi || kc0 >= pl.m || kc0 > jc) return;
  // a non-finite weight poisons every entry of the fp64 Gram; so does (defensively) a negative weight on a path that
  // was planned for non-negative ones — never a silently wrong matrix
  const bool poisoned = wstat[WS_BAD] != 0.0 || (expect_nonneg && wstat[WS_NNEG] != 0.0);
  const int64_t cstride = (int64_t)pl.m * pl.ldp;
  unsigned __int128 acc[4] = {0, 0, 0, 0};
  double frac[4] = {0.0, 0.0, 0.0, 0.0};
  const int var = pl.nmod - kNModMin;  // which moduli prefix the planes were built with
#pragma unroll 1
  for (int l = 0; l < pl.nmod; ++l) {
    const int p = c_mod_p[l], ql = c_mod_q[var][l];
    const int8_t* src = partial + ((int64_t)l * pl.pchunks * pl.m + jc) * pl.ldp + kc0;
    int s[4] = {0, 0, 0, 0};
    for (int c = 0; c < pl.pchunks; ++c) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(src + c * cstride);
      const int sg = c < pl.main_chunks ? pl.sign_main : pl.sign_extra;
      s[0] += sg * (int)(int8_t)(v & 0xff);
      s[1] += sg * (int)(int8_t)((v >> 8) & 0xff);
      s[2] += sg * (int)(int8_t)((v >> 16) & 0xff);
      s[3] += sg * (int)(int8_t)(v >> 24);
    }
    const unsigned __int128 Ml = ((unsigned __int128)c_mod_Mhi[var][l] << 64) | c_mod_Mlo[var][l];
    const double ipd = 1.0 / (double)p;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int r = s[e] % p;
      if (r < 0) r += p;
      const unsigned tl = (unsigned)((r * ql) % p);
      acc[e] += Ml * tl;
      frac[e] += (double)tl * ipd;
    }
  }
  const unsigned __int128 P = ((unsigned __int128)c_P_hi[var] << 64) | c_P_lo[var];
  const double ij = inv[jc];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int kc = kc0 + e;
    if (kc >= pl.m || kc > jc) continue;
    long long k = (long long)floor(frac[e]);
    if (k < 0) k = 0;
    __int128 r = (__int128)acc[e] - (__int128)(P * (unsigned __int128)k);
    while (r < 0) r += (__int128)P;
    while (r >= (__int128)P) r -= (__int128)P;
    if (r > (__int128)(P >> 1)) r -= (__int128)P;
    const bool neg = r < 0;
    const unsigned __int128 mag = neg ? (unsigned __int128)(-r) : (unsigned __int128)r;
    double d = ldexp((double)(unsigned long long)(mag >> 64), 64) + (double)(unsigned long long)mag;
    if (neg) d = -d;
    double g = (d * ij) * inv[kc];
    if (poisoned) g = __longlong_as_double(0x7ff8000000000000LL);
    if (Gpack) {
      Gpack[(int64_t)jc * (jc + 1) / 2 + kc] = g;
    } else {
      G[(int64_t)kc * pl.m + jc] = g;
      G[(int64_t)jc * pl.m + kc] = g;
    }
  }
}

// G (m x m, both triangles) from the packed upper triangle P[jc (jc+1)/2 + kc] (kc <= jc).  32 x 32 tiles, grid.x enumerates
// the tiles with tile-row (kc) <= tile-column (jc); reads and both writes are coalesced (transposed half through smem).
__global__ void __launch_bounds__(256) k_unpack_upper(const double* __restrict__ P, int m, double* __restrict__ G) {
  __shared__ double tile[32][33];
  int tj, tk;  // jc tile >= kc tile
  tri_decode(blockIdx.x, tj, tk);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int cc = ty; cc < 32; cc += 8) {
    const int jc = tj * 32 + cc, kc = tk * 32 + tx;
    double v = 0.0;
    if (jc < m && kc <= jc) {
      v = P[(int64_t)jc * (jc + 1) / 2 + kc];
      G[(int64_t)jc * m + kc] = v;  // (row kc, column jc)
    }
    tile[cc][tx] = v;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int kc = tk * 32 + rr, jc = tj * 32 + tx;
    if (jc < m && kc < jc) G[(int64_t)kc * m + jc] = tile[tx][rr];  // (row jc, column kc)
  }
}

}  // namespace scs
