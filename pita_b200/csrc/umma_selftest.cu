// Self-test of the tcgen05 plumbing in umma.cuh: D[128x32] = A[128x32] * B[32x32]^T with operands written to
// swizzled shared memory by the CUDA cores, one tile, one CTA.  split=1 uses the 3xTF32 error-compensated form.
#include "common.cuh"
#include "umma.cuh"

namespace pita {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                            float *__restrict__ D, int split) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-B alignment of the dynamic shared window
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float *sA = reinterpret_cast<float *>(base);             // 128 x 32 fp32 = 16 KB
  float *sAlo = sA + 128 * 32;                             // 16 KB
  float *sB = sAlo + 128 * 32;                             // 32 x 32 fp32 = 4 KB
  float *sBlo = sB + 32 * 32;                              // 4 KB
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sBlo + 32 * 32);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) umma::tmem_alloc<32>(tmem_slot);
  if (tid == 0) { umma::mbar_init(mbar, 1); umma::fence_mbar_init(); }
  {
    float v[32], hi[32], lo[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { v[k] = A[tid * 32 + k]; umma::split_tf32(v[k], hi[k], lo[k]); }
    umma::store_row_sw128(sA, tid, split ? hi : v);
    umma::store_row_sw128(sAlo, tid, lo);
    if (tid < 32) {
#pragma unroll
      for (int k = 0; k < 32; ++k) { v[k] = Bm[tid * 32 + k]; umma::split_tf32(v[k], hi[k], lo[k]); }
      umma::store_row_sw128(sB, tid, split ? hi : v);
      umma::store_row_sw128(sBlo, tid, lo);
    }
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
    const uint64_t dA = umma::make_desc_sw128_kmajor(umma::smem_u32(sA)), dAl = umma::make_desc_sw128_kmajor(umma::smem_u32(sAlo));
    const uint64_t dB = umma::make_desc_sw128_kmajor(umma::smem_u32(sB)), dBl = umma::make_desc_sw128_kmajor(umma::smem_u32(sBlo));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_base, dA + 2 * k, dB + 2 * k, idesc, k > 0);  // +32 B along K per step
    if (split) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_base, dAl + 2 * k, dB + 2 * k, idesc, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_base, dA + 2 * k, dBl + 2 * k, idesc, 1);
    }
    umma::commit(mbar);
  }
  umma::mbar_wait(mbar, 0);
  umma::fence_after_thread_sync();
  float acc[32];
  umma::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
#pragma unroll
  for (int c = 0; c < 32; ++c) D[tid * 32 + c] = acc[c];
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<32>(tmem_base);
}

}  // namespace pita

extern "C" int pita_umma_selftest(const float *A, const float *B, float *D, int split, void *stream) {
  using namespace pita;
  PITA_REQUIRE(A && B && D, PITA_EINVAL, "umma_selftest: null pointer");
  const size_t bytes = 1024 + (2 * 128 * 32 + 2 * 32 * 32) * sizeof(float) + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("umma_selftest: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  umma_selftest_kernel<<<1, 128, bytes, static_cast<cudaStream_t>(stream)>>>(A, B, D, split);
  PITA_CHECK_LAUNCH("umma_selftest_kernel");
  return PITA_OK;
}
