// Micro-benchmark (tuning aid, not product code): HBM read bandwidth of TMA box streams over a column-major fp64
// matrix as a function of the box shape (rows x cols), the number of CTAs and the ring depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench tma_stream_bench.cu -lcuda
//   ./tma_stream_bench n m
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!try_wait(bar, parity)) if (clock64() - t0 > 4000000000LL) __trap();
}

__global__ void __launch_bounds__(64, 1)
k_stream(const __grid_constant__ CUtensorMap map, int R, int Cb, int CS, int64_t npanels, int stages, int box_bytes, int active_ctas) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (su32(raw) + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)stages * box_bytes;
  if ((int)blockIdx.x >= active_ctas) return;
  const int slice = blockIdx.x % CS, group = blockIdx.x / CS, ngroups = active_ctas / CS;
  if (group >= ngroups) return;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * s), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8u * (stages + s)), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 1;
    for (int64_t p = group; p < npanels; p += ngroups) {
      wait(bars + 8u * (stages + s), ph);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * s), "r"(box_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(base + (uint32_t)s * box_bytes), "l"(&map), "r"(bars + 8u * s), "r"((int)(p * R)), "r"(slice * Cb) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int64_t p = group; p < npanels; p += ngroups) {
      wait(bars + 8u * s, ph);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * (stages + s)) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 2000000, m = argc > 2 ? atoll(argv[2]) : 2048;
  double* A;
  CK(cudaMalloc(&A, (size_t)n * m * 8));
  CK(cudaMemset(A, 0, (size_t)n * m * 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int R, Cb, stages, ctas, promo; };
  std::vector<Cfg> cfgs;
  for (int ctas : {120, 144})
    for (int R : {16, 32, 64, 128, 256})
      cfgs.push_back({R, 4096 / R, 6, ctas, 2});
  cfgs.push_back({16, 256, 6, 144, 0});
  cfgs.push_back({16, 256, 6, 144, 1});
  cfgs.push_back({16, 256, 3, 144, 2});
  cfgs.push_back({32, 256, 3, 144, 2});
  cfgs.push_back({64, 128, 3, 144, 2});
  for (auto c : cfgs) {
    const int CS = (int)(m / c.Cb);
    if (CS < 1 || c.ctas / CS < 1) continue;
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    cuuint64_t gstr[1] = {(cuuint64_t)n * 8};
    cuuint32_t box[2] = {(cuuint32_t)c.R, (cuuint32_t)c.Cb};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapL2promotion promo = c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.R == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed R=%d Cb=%d: %d\n", c.R, c.Cb, (int)r); continue; }
    const int box_bytes = c.R * c.Cb * 8;
    const size_t smem = (size_t)c.stages * box_bytes + 16 * c.stages + 2048;
    const int64_t npanels = n / c.R;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_stream<<<148, 64, smem>>>(map, c.R, c.Cb, CS, npanels, c.stages, box_bytes, c.ctas);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double bytes = (double)npanels * c.R * (double)(CS * c.Cb) * 8.0;
    printf("box %3d rows x %3d cols (%3d B contiguous), stages %d x %d KB, %3d CTAs, L2promo %d: %.3f ms  %.0f GB/s\n", c.R, c.Cb, c.R * 8,
           c.stages, box_bytes / 1024, c.ctas, c.promo, best, bytes / best / 1e6);
  }
  return 0;
}
