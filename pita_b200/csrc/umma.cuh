// Minimal sm_100a tensor-core plumbing used by the EGNN tangent kernel: tcgen05.mma (kind::tf32, operands in
// shared memory, fp32 accumulator in TMEM), TMEM allocation, mbarrier completion, tcgen05.ld epilogue loads.
// Operand tiles are K-major rows of exactly 128 bytes (32 x fp32) in the 128-byte swizzle, written by the
// CUDA cores (one thread per row), so no TMA is involved; `fence.proxy.async` orders those generic-proxy
// stores before the tensor core's async-proxy reads.
#pragma once
#include <stdint.h>

namespace pita {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128_kmajor(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)0 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // layout type: SWIZZLE_128B
  return d;
}

// ---- instruction descriptor: TF32 x TF32 -> F32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) /* c = F32 */ | (2u << 7) /* a = TF32 */ | (2u << 10) /* b = TF32 */ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand in TMEM (row r of the 128 x 32 tile = lane r, K element k = 32-bit column k; 8 columns per K = 8 step)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued tcgen05.mma of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t *mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 32 columns of fp32: thread `lane` of the warp receives row (lane_base + lane), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// thread `lane` of the warp writes 32 fp32 values to row (lane_base + lane), columns [col, col+32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
      "tcgen05.wait::st.sync.aligned;" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

// ---- writing one K-major operand row (32 fp32 = 128 B) into a SWIZZLE_128B tile whose base is 1024-B aligned.
// 16-byte chunk j of row r lands at chunk (j ^ (r & 7)).
__device__ __forceinline__ void store_row_sw128(float *tile, int row, const float (&v)[32]) {
  const uint32_t base = smem_u32(tile) + (uint32_t)row * 128u;
  const uint32_t x = (uint32_t)(row & 7);
#pragma unroll
  for (uint32_t j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((j ^ x) << 4)), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                 : "memory");
}

// Split an fp32 value into two TF32-representable parts, hi + lo ~= a (for 3xTF32).  Both parts are rounded to
// nearest (add half an ulp of the 10-bit mantissa, then clear the 13 low bits) instead of being left to the tensor
// core's truncation: |a - hi - lo| <= 2^-24 |a|, which keeps the three-product sum at plain-fp32 accuracy
// (measured: truncating splits are ~8x less accurate on the LJ-55 network).
__device__ __forceinline__ float round_tf32(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
  hi = round_tf32(a);
  lo = round_tf32(a - hi);
}

// Tangent rows (forward-mode derivative products, which only feed the divergence): hi rounded to nearest, lo = a - hi left
// to the tensor core's truncation.  |a - hi - trunc(lo)| <= 2^-21 |a| with a sign that follows the (random) sign of lo,
// i.e. unbiased; 3 instead of 5 integer/float instructions per element.  Primal rows keep split_tf32.
__device__ __forceinline__ void split_tf32_tangent(float a, float &hi, float &lo) {
  hi = round_tf32(a);
  lo = a - hi;
}

}  // namespace umma
}  // namespace pita
