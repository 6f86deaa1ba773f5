// Micro-benchmark of the sm_100a FP32 issue rates that bound the Lennard-Jones kernel (profiles/README.md).
// Reports warp-instructions / clk / SM for scalar vs packed (f32x2) FP32 ops and MUFU.RCP, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o fp32_pipes.bin fp32_pipes.cu && ./fp32_pipes.bin
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float rcpa(float a) { float d; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

constexpr int ILP = 8;
enum { K_FFMA, K_FFMA2, K_FADD, K_FADD2, K_FMUL, K_FMUL2, K_FFMA_FADD, K_FFMA2_FADD2, K_RCP, K_FFMA8_RCP1, K_FFMA2x9_RCP2, K_FFMA2_BCAST, K_COUNT };
static const char *names[K_COUNT] = {"FFMA", "FFMA2", "FADD", "FADD2", "FMUL", "FMUL2", "FFMA+FADD 1:1", "FFMA2+FADD2 1:1", "MUFU.RCP",
                                     "FFMA x8 + RCP x1", "FFMA2 x9 + RCP x2", "FFMA2 scalar-bcast operand"};
static const int instr_per_iter[K_COUNT] = {ILP, ILP, ILP, ILP, ILP, ILP, 2 * ILP, 2 * ILP, ILP, 9 * ILP, 11 * ILP, ILP};

template <int K>
__global__ void __launch_bounds__(256) bench(const float *in, float *out, long long *cyc, int iters) {
  float a[ILP], b = in[threadIdx.x & 31], c = in[32 + (threadIdx.x & 31)];
  u64 A[ILP], Bp = pk(b, c), Cp = pk(c, b);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = in[64 + i] + threadIdx.x; A[i] = pk(a[i], a[i] + 1.f); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (K == K_FFMA) a[i] = ffma(a[i], b, c);
      if (K == K_FFMA2) A[i] = ffma2(A[i], Bp, Cp);
      if (K == K_FADD) a[i] = fadd(a[i], b);
      if (K == K_FADD2) A[i] = fadd2(A[i], Bp);
      if (K == K_FMUL) a[i] = fmul(a[i], b);
      if (K == K_FMUL2) A[i] = fmul2(A[i], Bp);
      if (K == K_FFMA_FADD) { a[i] = ffma(a[i], b, c); a[i] = fadd(a[i], c); }
      if (K == K_FFMA2_FADD2) { A[i] = ffma2(A[i], Bp, Cp); A[i] = fadd2(A[i], Cp); }
      if (K == K_RCP) a[i] = rcpa(a[i]);
      if (K == K_FFMA8_RCP1) {
#pragma unroll
        for (int r = 0; r < 8; ++r) a[i] = ffma(a[i], b, c);
        a[i] = rcpa(a[i]);
      }
      if (K == K_FFMA2x9_RCP2) {
#pragma unroll
        for (int r = 0; r < 9; ++r) A[i] = ffma2(A[i], Bp, Cp);
        float lo, hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i]));
        A[i] = pk(rcpa(lo), rcpa(hi));
      }
      if (K == K_FFMA2_BCAST) A[i] = ffma2(A[i], pk(b, b), Cp);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i])); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K>
void run(const float *in, float *out, long long *cyc, int blocks_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096, blocks = sms * blocks_per_sm;
  bench<K><<<blocks, 256>>>(in, out, cyc, iters);  // warm
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<K><<<blocks, 256>>>(in, out, cyc, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; long long mx = 0;
  for (int i = 0; i < blocks; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= blocks;
  double winstr_per_sm = (double)blocks_per_sm * 8 * iters * instr_per_iter[K];
  printf("%-28s blocks/SM %d  warp-instr/clk/SM %.3f (avg-cycles)  %.3f (max-cycles)  time %.3f ms  -> %.2f Ginstr/s/SM\n", names[K], blocks_per_sm,
         winstr_per_sm / avg, winstr_per_sm / mx, ms, winstr_per_sm / (ms * 1e6));
  delete[] h;
}

int main() {
  float *in, *out; long long *cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 8 * 256 * 4 * 2); cudaMalloc(&cyc, 148 * 8 * 8 * 2);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 1e-3f * i;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  for (int bps = 2; bps <= 4; bps += 2) {
    run<K_FFMA>(in, out, cyc, bps); run<K_FFMA2>(in, out, cyc, bps); run<K_FADD>(in, out, cyc, bps); run<K_FADD2>(in, out, cyc, bps);
    run<K_FMUL>(in, out, cyc, bps); run<K_FMUL2>(in, out, cyc, bps); run<K_FFMA_FADD>(in, out, cyc, bps); run<K_FFMA2_FADD2>(in, out, cyc, bps);
    run<K_RCP>(in, out, cyc, bps); run<K_FFMA8_RCP1>(in, out, cyc, bps); run<K_FFMA2x9_RCP2>(in, out, cyc, bps); run<K_FFMA2_BCAST>(in, out, cyc, bps);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
