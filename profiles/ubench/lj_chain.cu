// Micro-benchmark: the Lennard-Jones pair chain of csrc/lj.cu (18 packed FP32 instructions + 2 MUFU.RCP per two pairs) as a
// function of (warps per SM, chains in flight per warp).  Answers: what FMA-pipe occupancy can this instruction mix reach at
// the 8-10 warps per SM the register budget of the real kernel allows, and how much interleaving does it take?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o lj_chain.bin lj_chain.cu && ./lj_chain.bin
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(u64 a, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rcpa(float a) { float d; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

// R own atoms per thread; STAGED = evaluate the R chains stage by stage in the source (vs. one after the other)
template <int R, bool STAGED, int UNR>
__global__ void __launch_bounds__(256) chain(const float *in, float *out, long long *cyc, int iters) {
  float ax[R], ay[R], az[R];
  u64 fx[R], fy[R], fz[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ax[r] = in[r] + threadIdx.x * 1e-3f; ay[r] = in[8 + r]; az[r] = in[16 + r];
    fx[r] = fy[r] = fz[r] = pk(0.f, 0.f);
  }
  u64 bx = pk(in[24], in[25]), by = pk(in[26], in[27]), bz = pk(in[28], in[29]);
  const u64 dlt = pk(in[30] * 1e-3f, in[31] * 1e-3f), eps2 = pk(1e-6f, 1e-6f);
  u64 e6 = pk(0.f, 0.f), en3 = pk(0.f, 0.f), rx = pk(0.f, 0.f), ry = rx, rz = rx;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll UNR
  for (int it = 0; it < iters; ++it) {
    bx = add2(bx, dlt);   // (1 extra packed instruction per step: the "streamed atom")
    if (!STAGED) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const u64 dx = sub2(pk(ax[r], ax[r]), bx), dy = sub2(pk(ay[r], ay[r]), by), dz = sub2(pk(az[r], az[r]), bz);
        u64 s = fma2(dx, dx, eps2); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
        float lo, hi; unpk(s, lo, hi);
        const u64 ninv = pk(rcpa(-lo), rcpa(-hi));
        const u64 inv2 = mul2(ninv, ninv), ni3 = mul2(inv2, ninv);
        e6 = fma2(ni3, ni3, e6); en3 = add2(en3, ni3);
        const u64 w = mul2(inv2, inv2), nfs = fma2(w, ni3, w);
        fx[r] = fma2(nfs, dx, fx[r]); fy[r] = fma2(nfs, dy, fy[r]); fz[r] = fma2(nfs, dz, fz[r]);
        rx = fma2(nfs, dx, rx); ry = fma2(nfs, dy, ry); rz = fma2(nfs, dz, rz);
      }
    } else {
      u64 dx[R], dy[R], dz[R], q[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        dx[r] = sub2(pk(ax[r], ax[r]), bx); dy[r] = sub2(pk(ay[r], ay[r]), by); dz[r] = sub2(pk(az[r], az[r]), bz);
        u64 s = fma2(dx[r], dx[r], eps2); s = fma2(dy[r], dy[r], s); q[r] = fma2(dz[r], dz[r], s);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) { float lo, hi; unpk(q[r], lo, hi); q[r] = pk(rcpa(-lo), rcpa(-hi)); }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const u64 ninv = q[r], inv2 = mul2(ninv, ninv), ni3 = mul2(inv2, ninv);
        e6 = fma2(ni3, ni3, e6); en3 = add2(en3, ni3);
        const u64 w = mul2(inv2, inv2); q[r] = fma2(w, ni3, w);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        fx[r] = fma2(q[r], dx[r], fx[r]); fy[r] = fma2(q[r], dy[r], fy[r]); fz[r] = fma2(q[r], dz[r], fz[r]);
        rx = fma2(q[r], dx[r], rx); ry = fma2(q[r], dy[r], ry); rz = fma2(q[r], dz[r], rz);
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f, lo, hi;
#pragma unroll
  for (int r = 0; r < R; ++r) { unpk(fx[r], lo, hi); s += lo + hi; unpk(fy[r], lo, hi); s += lo + hi; unpk(fz[r], lo, hi); s += lo + hi; }
  unpk(e6, lo, hi); s += lo + hi; unpk(en3, lo, hi); s += lo + hi; unpk(rx, lo, hi); s += lo + hi; unpk(ry, lo, hi); s += lo + hi;
  unpk(rz, lo, hi); s += lo + hi;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, bool STAGED, int UNR = 1>
void run(const float *in, float *out, long long *cyc, int warps_per_block, int blocks_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048, blocks = sms * blocks_per_sm;
  chain<R, STAGED, UNR><<<blocks, 32 * warps_per_block>>>(in, out, cyc, iters);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain<R, STAGED, UNR><<<blocks, 32 * warps_per_block>>>(in, out, cyc, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = new long long[blocks]; cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < blocks; ++i) if (h[i] > mx) mx = h[i];
  // FMA-pipe cycles needed per SM sub-partition: 18 packed instr x 2 cycles per chain evaluation (+2 for the step's extra add)
  const double warps_sm = (double)warps_per_block * blocks_per_sm;
  const double pipe_cycles = warps_sm / 4.0 * iters * (R * 18 * 2 + 2);
  printf("R=%d unroll %2d (%5d instr straight-line) %-7s warps/SM %4.0f (%d x %d)  cycles %lld  FMA-pipe occupancy %.3f   (%.3f ms)\n", R, UNR, UNR * (R * 20 + 1), STAGED ? "staged" : "serial", warps_sm,
         warps_per_block, blocks_per_sm, mx, pipe_cycles / mx, ms);
  delete[] h;
}

int main() {
  float *in, *out; long long *cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 16 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 0.37f * (i % 7) + 1e-3f * i;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const int cfgs[][2] = {{4, 1}, {5, 1}, {4, 2}, {5, 2}, {4, 3}, {8, 2}, {8, 4}};
  for (auto &c : cfgs) {
    run<1, false>(in, out, cyc, c[0], c[1]);
    run<2, false>(in, out, cyc, c[0], c[1]);
    run<4, false>(in, out, cyc, c[0], c[1]);
    run<4, true>(in, out, cyc, c[0], c[1]);
    run<8, false>(in, out, cyc, c[0], c[1]);
    run<8, true>(in, out, cyc, c[0], c[1]);
  }
  // the same chains as straight-line code of growing size (the real kernels are fully unrolled: 1 700 - 4 300 instructions)
  const int cfg2[][2] = {{4, 2}, {5, 2}, {8, 2}};
  for (auto &c : cfg2) {
    run<4, true, 1>(in, out, cyc, c[0], c[1]);
    run<4, true, 8>(in, out, cyc, c[0], c[1]);
    run<4, true, 32>(in, out, cyc, c[0], c[1]);
    run<4, true, 128>(in, out, cyc, c[0], c[1]);
    run<4, false, 32>(in, out, cyc, c[0], c[1]);
    run<4, false, 128>(in, out, cyc, c[0], c[1]);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
