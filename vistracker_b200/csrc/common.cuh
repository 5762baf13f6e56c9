// Shared helpers for the vistracker_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace vt {

// ---- error plumbing: every extern "C" entry returns 0 or a negative code; the message is kept per thread
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define VT_CHECK_ARG(cond, ...) do { if (!(cond)) { vt::set_error(__VA_ARGS__); return -1; } } while (0)
#define VT_CHECK_LAUNCH(what) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return vt::cuda_fail(e__, what); } while (0)

constexpr int kNumSMs = 148;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- per-channel GroupNorm statistics: double stats[img][ld][2] = (sum, sum of squares) per channel.
// Blocks reduce in fp32 over <= a few hundred values per channel, then add into the fp64 slot, so
// E[x^2]-E[x]^2 is formed in double.

// Sum of v over the 32 lanes for 32 "columns" at once: on return lane l holds sum_lanes v[l].
// Recursive halving, 31 shuffles (instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32]) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      // lanes with bit `half` set keep columns [half, 2*half), the others keep [0, half)
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  // after the last step lane l holds column bitreverse-free index: column index == lane (bits consumed MSB first)
  return v[0];
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// fp32 -> (hi, lo) fp16 pair with x ~= hi + lo * 2^-11.  |x| is saturated to the fp16 range and the
// saturation is counted in *overflow so the host can fail loudly.
constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo, int& sat) {
  if (fabsf(x) > 65504.0f) { sat = 1; x = copysignf(65504.0f, x); }
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}

}  // namespace vt
