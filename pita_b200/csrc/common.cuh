// Shared helpers for the pita_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pita_b200.h"

namespace pita {

void set_error(const char *fmt, ...);

#define PITA_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      ::pita::set_error(__VA_ARGS__);   \
      return (code);                    \
    }                                   \
  } while (0)

#define PITA_CHECK_LAUNCH(what)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      ::pita::set_error("%s: %s", (what), cudaGetErrorString(e__));          \
      return PITA_ECUDA;                                                     \
    }                                                                        \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float fast_rcp(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// logistic sigmoid; |rel err| ~ 3e-7 (ex2.approx + rcp.approx)
__device__ __forceinline__ float sigmoidf_fast(float z) { return fast_rcp(1.0f + __expf(-z)); }

// SiLU value and derivative from one sigmoid:  silu = z*s,  silu' = s*(1 + z*(1-s))
__device__ __forceinline__ void silu_both(float z, float &val, float &der) {
  const float s = sigmoidf_fast(z);
  val = z * s;
  der = s * fmaf(z, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float silu_val(float z) { return z * sigmoidf_fast(z); }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pita
