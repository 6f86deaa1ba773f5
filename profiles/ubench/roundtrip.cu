// Micro-benchmark of ONE "round trip" of the row engine (pita_b200/csrc/rowgemm.cuh): every thread of a 128-row team
// hands its 32-float operand row to the tensor core, one elected thread issues the [128x32]x[32x32] tcgen05.mma
// (kind::tf32; x3 for 3xTF32), the team waits on the mbarrier and reads its accumulator row back with tcgen05.ld.
// Variants:  SS = A operand in 128B-swizzled shared memory (what the kernels do today: STS + fence.proxy.async)
//            TS = A operand in TMEM (tcgen05.st of the row into the thread's own lane, no shared-memory tile, no proxy fence)
// Reports cycles per phase (hand-over, sync+issue+wait, read-back) for 1 and 2 teams per CTA, and checks TS == SS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../pita_b200/csrc -o roundtrip.bin roundtrip.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "rowgemm.cuh"

using namespace pita;
using namespace pita::rg;

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool SPLIT, bool TS, int NTEAM, int WORK>
__global__ void __launch_bounds__(NTEAM * 128, 1)
rt_kernel(const float *__restrict__ w, int iters, float *__restrict__ out, long long *__restrict__ cyc) {
  extern __shared__ __align__(16) float sm_raw[];
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  float *wsm = reinterpret_cast<float *>(base);                                   // one weight slot (hi [+ lo])
  uint8_t *a_base = base + Bytes<SPLIT>::kW;                                      // NTEAM A tiles
  uint64_t *mbars = reinterpret_cast<uint64_t *>(a_base + NTEAM * Bytes<SPLIT>::kA);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbars + NTEAM);
  const int tid = threadIdx.x, team = tid >> 7, tt = tid & 127, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int k = 0; k < NTEAM; ++k) umma::mbar_init(mbars + k, 1);
    umma::fence_mbar_init();
  }
  WeightSrc src;
  src.p[0] = w;
  src.count = 1;
  load_weight_tiles<SPLIT>(wsm, src, tid, NTEAM * 128);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  Team<SPLIT> T;
  const uint32_t tmem_base = uniform32(*tmem_slot), team_u = uniform32((uint32_t)team);
  T.a_hi = reinterpret_cast<float *>(a_base + (size_t)team * Bytes<SPLIT>::kA);
  T.a_addr = uniform32(umma::smem_u32(a_base)) + team_u * (uint32_t)Bytes<SPLIT>::kA;
  T.w_addr = uniform32(umma::smem_u32(wsm));
  T.mbar_addr = uniform32(umma::smem_u32(mbars)) + team_u * 8u;
  T.phase = 0;
  T.tmem_col = tmem_base + team_u * (uint32_t)(512 / NTEAM);
  T.tmem = T.tmem_col + (((uint32_t)((warp & 3) * 32)) << 16);
  T.bar_id = 1 + team;
  T.tt = tt;
  T.issuer = uniform32((uint32_t)(warp & 3)) == 0u;
  constexpr int sAcc = 0, sAhi = 1, sAlo = 2;

  float row[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) row[k] = 0.01f * (float)((tt * 7 + k * 3) % 41) - 0.2f;
  long long c_store = 0, c_rt = 0, c_ld = 0, c_work = 0;
  for (int it = 0; it < iters; ++it) {
    const long long t0 = clock64();
    if (TS) {
      if (SPLIT) {
        float h[32], l[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) umma::split_tf32(row[k], h[k], l[k]);
        T.st(sAhi, h);
        T.st(sAlo, l);
      } else {
        T.st(sAhi, row);
      }
    } else {
      T.store_row(row);
    }
    const long long t1 = clock64();
    if (TS) {
      umma::fence_before_thread_sync();
      T.sync();
      if (T.issuer) {
        if (elect_one()) {
          umma::fence_after_thread_sync();
          constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
          const uint32_t d = T.tmem_col + 32u * sAcc, ah = T.tmem_col + 32u * sAhi, al = T.tmem_col + 32u * sAlo;
          const uint64_t dB = umma::make_desc_sw128_kmajor(T.w_addr), dBl = umma::make_desc_sw128_kmajor(T.w_addr + 4096u);
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_ts(d, al + 8u * k, dB + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_ts(d, ah + 8u * k, dBl + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, k > 0 ? 1u : 0u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(T.mbar_addr) : "memory");
        }
        __syncwarp();
      }
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "RTB_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
          "@P1 bra RTB_DONE;\n\t"
          "bra RTB_WAIT;\n\t"
          "RTB_DONE:\n\t}\n" ::"r"(T.mbar_addr),
          "r"(T.phase)
          : "memory");
      T.phase ^= 1u;
      umma::fence_after_thread_sync();
    } else {
      T.round_trip([&] { T.mma(sAcc, 0, false); });
    }
    const long long t2 = clock64();
    T.ld(sAcc, row);
    const long long t3 = clock64();
    // element-wise stage stand-in: WORK dependent FMAs per channel
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float v = row[k] * 0.05f + 0.01f * (float)(k & 3);
#pragma unroll
      for (int r = 0; r < WORK; ++r) v = fmaf(v, 0.999f, 0.001f);
      row[k] = v;
    }
    const long long t4 = clock64();
    c_store += t1 - t0; c_rt += t2 - t1; c_ld += t3 - t2; c_work += t4 - t3;
  }
  if (blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 32; ++k) out[(team * 128 + tt) * 32 + k] = row[k];
    if (tt == 0) { cyc[team * 4 + 0] = c_store; cyc[team * 4 + 1] = c_rt; cyc[team * 4 + 2] = c_ld; cyc[team * 4 + 3] = c_work; }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem_base);
}

template <bool SPLIT, bool TS, int NTEAM, int WORK>
static void run(const char *name, const float *dw, float *dout, long long *dcyc, float *hout, int iters) {
  const size_t bytes = 1024 + Bytes<SPLIT>::kW + NTEAM * Bytes<SPLIT>::kA + 64;
  cudaFuncSetAttribute(rt_kernel<SPLIT, TS, NTEAM, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rt_kernel<SPLIT, TS, NTEAM, WORK><<<148, NTEAM * 128, bytes>>>(dw, 10, dout, dcyc);
  cudaEventRecord(e0);
  rt_kernel<SPLIT, TS, NTEAM, WORK><<<148, NTEAM * 128, bytes>>>(dw, iters, dout, dcyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c[8] = {0};
  cudaMemcpy(c, dcyc, sizeof(long long) * 4 * NTEAM, cudaMemcpyDeviceToHost);
  cudaMemcpy(hout, dout, sizeof(float) * NTEAM * 128 * 32, cudaMemcpyDeviceToHost);
  printf("%-34s teams %d work %2d  %7.1f ns/iter | cycles/iter: hand-over %6.0f  sync+mma+wait %6.0f  tcgen05.ld %5.0f  elementwise %6.0f  (%s)\n",
         name, NTEAM, WORK, 1e6 * ms / iters, (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters, (double)c[3] / iters,
         cudaGetErrorString(err));
}

int main() {
  float hw[32 * 32];
  for (int i = 0; i < 1024; ++i) hw[i] = 0.03f * (float)((i * 13) % 23) - 0.3f;
  float *dw, *dout;
  long long *dcyc;
  cudaMalloc(&dw, sizeof(hw)); cudaMalloc(&dout, sizeof(float) * 2 * 128 * 32); cudaMalloc(&dcyc, sizeof(long long) * 8);
  cudaMemcpy(dw, hw, sizeof(hw), cudaMemcpyHostToDevice);
  static float ref[2 * 128 * 32], got[2 * 128 * 32];
  const int iters = 2000;
  // correctness: 3 iterations, TS must reproduce SS
  for (int split = 0; split < 2; ++split) {
    if (split) { run<true, false, 1, 0>("check SS 3xTF32", dw, dout, dcyc, ref, 3); run<true, true, 1, 0>("check TS 3xTF32", dw, dout, dcyc, got, 3); }
    else { run<false, false, 1, 0>("check SS TF32", dw, dout, dcyc, ref, 3); run<false, true, 1, 0>("check TS TF32", dw, dout, dcyc, got, 3); }
    double md = 0, mr = 0;
    for (int i = 0; i < 128 * 32; ++i) { md = fmax(md, fabs((double)ref[i] - got[i])); mr = fmax(mr, fabs((double)ref[i])); }
    printf("TS vs SS (%s): max |diff| %.3e  (max |ref| %.3e)\n", split ? "3xTF32" : "TF32", md, mr);
  }
  run<false, false, 1, 0>("SS TF32", dw, dout, dcyc, got, iters);
  run<false, true, 1, 0>("TS TF32", dw, dout, dcyc, got, iters);
  run<true, false, 1, 0>("SS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, true, 1, 0>("TS 3xTF32", dw, dout, dcyc, got, iters);
  run<false, false, 2, 0>("SS TF32", dw, dout, dcyc, got, iters);
  run<false, true, 2, 0>("TS TF32", dw, dout, dcyc, got, iters);
  run<true, false, 2, 0>("SS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, true, 2, 0>("TS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, false, 2, 8>("SS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, true, 2, 8>("TS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, false, 2, 24>("SS 3xTF32", dw, dout, dcyc, got, iters);
  run<true, true, 2, 24>("TS 3xTF32", dw, dout, dcyc, got, iters);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
