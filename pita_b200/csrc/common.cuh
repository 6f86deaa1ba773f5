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

__device__ __forceinline__ float fast_ex2(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// logistic sigmoid; |rel err| ~ 3e-7 (ex2.approx + rcp.approx; flush-to-zero forms: no range fix-up code)
__device__ __forceinline__ float sigmoidf_fast(float z) { return fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * z)); }

// explicit shared-space 128-bit accesses (generic pointers into dynamic shared memory otherwise compile to LD.E/ST.E)
__device__ __forceinline__ float4 lds4(const float *p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p)))
               : "memory");
  return v;
}
// Schedulable 128-bit shared-memory load for tables that are CONSTANT once published (layer vectors, class tables): a plain
// asm (not volatile, no memory clobber) whose only input is the address, so the compiler may issue it ahead of the arithmetic
// that consumes the previous one — the volatile lds4 above keeps program order and exposes one shared-memory latency per load.
// Ordering against the barrier that published the table comes from the POINTER: pass it through opq() after that barrier (the
// stage helpers do so on entry), which also keeps the loads from being hoisted out of the surrounding loops.
__device__ __forceinline__ float4 lds4c(const float *p) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))));
  return v;
}
__device__ __forceinline__ const float *opq(const float *p) {
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ void sts4(float *p, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// 16-byte asynchronous global -> shared copies (LDGSTS): issue, commit as one group, wait for every group of this thread
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float lds1(const float *p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
  return v;
}

// SiLU value and derivative from one sigmoid:  silu = z*s,  silu' = s*(1 + z*(1-s))
__device__ __forceinline__ void silu_both(float z, float &val, float &der) {
  const float s = sigmoidf_fast(z);
  val = z * s;
  der = s * fmaf(z, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float silu_val(float z) { return z * sigmoidf_fast(z); }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pita
