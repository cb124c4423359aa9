// Shared device helpers for the scs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SCS_DEVINL __device__ __forceinline__

namespace scs {

constexpr int kVecThreads = 1024;  // single-CTA vector kernels

// Butterfly sum: every lane ends with the same value, fixed order => deterministic.
SCS_DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for NT threads (NT multiple of 32, <= 1024). sh: >= 32 doubles of shared memory.
// All threads receive the total.  Fixed tree => deterministic.
template <int NT>
SCS_DEVINL double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double t = (lane < NT / 32) ? sh[lane] : 0.0;
  return warp_sum(t);
}

// 128-bit streaming load of two doubles: read-only path, do not allocate in L1 (data is touched once).
SCS_DEVINL double2 ldg_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

SCS_DEVINL double julia_sign(double u) {  // Base.sign: sign(±0)=±0, sign(NaN)=NaN
  return u > 0.0 ? 1.0 : (u < 0.0 ? -1.0 : u);
}

SCS_DEVINL double fmax_nan(double a, double b) {  // Julia max(): NaN if either is NaN
  return (a != a || b != b) ? (a + b) : (a > b ? a : b);
}
SCS_DEVINL double fmin_nan(double a, double b) {
  return (a != a || b != b) ? (a + b) : (a < b ? a : b);
}

}  // namespace scs
