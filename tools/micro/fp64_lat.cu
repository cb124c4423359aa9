// Micro-benchmark (tuning aid, not product code): dependent-issue latencies of the fp64 instructions that sit on the
// per-column chain of the on-device Cholesky (DFMA, DMUL, rsqrt(), 64-bit SHFL, broadcast LDS.128), one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cuda_runtime.h>
#include <cstdio>
#define N 512
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ __align__(16) double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed * 1e-9 * i;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001, acc = 0;
  long long t0, t1;
  // dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  acc += x;
  // dependent DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  acc += x;
  // dependent rsqrt
  x = 2.0 + seed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  acc += x;
  // dependent 64-bit shuffle
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  acc += x;
  // independent DFMAs: 8 chains (issue rate)
  double c0 = x, c1 = x + 1, c2 = x + 2, c3 = x + 3, c4 = x + 4, c5 = x + 5, c6 = x + 6, c7 = x + 7;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    c0 = fma(c0, y, 1e-9); c1 = fma(c1, y, 1e-9); c2 = fma(c2, y, 1e-9); c3 = fma(c3, y, 1e-9);
    c4 = fma(c4, y, 1e-9); c5 = fma(c5, y, 1e-9); c6 = fma(c6, y, 1e-9); c7 = fma(c7, y, 1e-9);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  acc += c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
  // broadcast LDS.128 feeding independent DFMAs (the trsm inner loop shape)
  double d[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    const double2* p = reinterpret_cast<const double2*>(sm + ((i * 8) & 1023));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 v = p[u];
      d[2 * u] = fma(-y, v.x, d[2 * u]);
      d[2 * u + 1] = fma(-y, v.y, d[2 * u + 1]);
    }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  for (int u = 0; u < 8; ++u) acc += d[u];
  // dependent DFMA -> FSEL pair (select) chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = (threadIdx.x > (i & 31)) ? fma(x, y, 1e-9) : x * 0.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  acc += x;
  out[threadIdx.x] = acc;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
  for (int warps = 1; warps <= 4; warps *= 4) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
    printf("warps/CTA %d: DFMA dep %.1f cyc, DMUL dep %.1f, rsqrt()+DADD dep %.1f, SHFL64 dep %.1f, DFMA 8-way ILP %.2f cyc/inst, LDS.128+2DFMA %.2f cyc/DFMA, DFMA/DMUL select %.1f\n",
           warps, h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (8.0 * N), h[5] / (8.0 * N), h[6] / (double)N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
