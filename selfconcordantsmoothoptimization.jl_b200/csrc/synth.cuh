// Device twin of oracle/synth.py: Philox-4x32-10 counter-based generator, bit-identical to the numpy
// restatement (tests/test_gpu_synth.py).  Used only to materialise benchmark-sized inputs directly in HBM
// (SURVEY.md §8(d)); the parity tests upload host data through scs_problem_create instead.
#pragma once
#include "common.cuh"

namespace scs {

struct Philox4 {
  uint32_t v[4];
};

SCS_DEVINL Philox4 philox4x32_10(uint64_t idx, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// unit-variance Irwin-Hall(8): exact integer sum of eight uint16 words, one fp64 multiply
SCS_DEVINL double ih8_normal(uint64_t idx, uint32_t stream, uint64_t seed) {
  const Philox4 o = philox4x32_10(idx, stream, seed);
  uint32_t s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += (o.v[q] & 0xFFFFu) + (o.v[q] >> 16);
  return ((double)s - 262140.0) * 0x1.3988e1412ed76p-16;
}
SCS_DEVINL double uniform53_01(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

// A[i + j*ldd] for local rows i < n_local (global row row0+i), counter = (row0+i) + j*n_total
__global__ void k_synth_A(double* __restrict__ A, int64_t ldd, int64_t n_local, int m, int64_t row0,
                          int64_t n_total, uint64_t seed, double density, double inv_sqrt_m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  for (int j = blockIdx.y; j < m; j += gridDim.y) {
    const uint64_t idx = (uint64_t)(row0 + i) + (uint64_t)j * (uint64_t)n_total;
    double v = ih8_normal(idx, 0u, seed) * inv_sqrt_m;
    if (density < 1.0) {
      const Philox4 o = philox4x32_10(idx, 0u, seed + 7);
      if (!(uniform53_01(o.v[2], o.v[3]) < density)) v = 0.0;
    }
    A[(int64_t)j * ldd + i] = v;
  }
}
// x_true: 5% nonzeros ~ 3*ih8
__global__ void k_synth_xtrue(double* __restrict__ x, int m, uint64_t seed, double frac, double sigma) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const Philox4 o = philox4x32_10((uint64_t)j, 1u, seed);
  const bool keep = uniform53_01(o.v[0], o.v[1]) < frac;
  x[j] = keep ? sigma * ih8_normal((uint64_t)j, 2u, seed) : 0.0;
}
// task 0: y = +1 if u < 1/(1+exp(-z)) else -1;  task 1: y = z + noise*ih8
__global__ void k_synth_y(double* __restrict__ y, const double* __restrict__ z, int64_t n_local, int64_t row0,
                          uint64_t seed, int task, double noise) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const uint64_t idx = (uint64_t)(row0 + i);
  if (task == 0) {
    const Philox4 o = philox4x32_10(idx, 3u, seed);
    y[i] = uniform53_01(o.v[0], o.v[1]) < 1.0 / (1.0 + exp(-z[i])) ? 1.0 : -1.0;
  } else {
    y[i] = z[i] + noise * ih8_normal(idx, 3u, seed);
  }
}

// Expands a CSC shard (Julia SparseMatrixCSC: colptr / rowval / nzval, index base 0 or 1) into the zero-initialised
// dense column-major resident layout.  One CTA per column (grid-stride), threads over that column's stored entries.
__global__ void k_scatter_csc(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                              const double* __restrict__ nzval, int64_t base, int64_t n, int m, int64_t ldd,
                              double* __restrict__ A, int* __restrict__ bad) {
  for (int j = blockIdx.x; j < m; j += gridDim.x) {
    const int64_t p0 = colptr[j] - base, p1 = colptr[j + 1] - base;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      const int64_t i = rowval[p] - base;
      if (i < 0 || i >= n)
        atomicExch(bad, 1);
      else
        A[(int64_t)j * ldd + i] = nzval[p];
    }
  }
}

}  // namespace scs
