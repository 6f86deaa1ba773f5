// Fused Euler-Maruyama / Feynman-Kac step, chunk-quantile clamp, and the post-processing row kernels.
// All HBM-bound: one warp per particle row, every [B,3n] tensor is touched exactly once.
// Replaces sdes.py:168-251 (drift assembly, diffusion), sde_integration.py:278-282,347-349 (update and
// gating), data_utils.py:4-26 (remove_mean), sdes.py:230 (quantile clamp), sde_integration.py:28-45,353-470.
#include "common.cuh"

namespace pita {

// ---- Philox4x32-10 + Box-Muller (in-kernel noise; parity mode passes a materialised tensor instead)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  n0 = r * c; n1 = r * s;
}

struct SdeP {
  float g2, gamma, dgamma_dt, dh_dt, dt, sqrt_dt, noise_scale;
  int debias, freeze_x, remove_mean;
  uint64_t seed, offset;
};

// ITEMS = ceil(3n/32): each lane owns elements lane, lane+32, ...
template <int ITEMS>
__global__ void __launch_bounds__(256)
sde_fk_step_kernel(const float *x, const float *__restrict__ gradU, const float *__restrict__ score,
                   const float *__restrict__ noise, const float *__restrict__ div, const float *__restrict__ dE_dh,
                   const float *__restrict__ energy, int64_t B, int n, SdeP p, float *x_out, float *__restrict__ a_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int D = 3 * n;
  const int64_t base = row * D;
  float v[ITEMS];
  float cross = 0.f;
  float nz[ITEMS];
  if (noise == nullptr) {
    constexpr int CALLS = (ITEMS + 3) / 4;
#pragma unroll
    for (int q = 0; q < CALLS; ++q) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(lane * CALLS + q), (uint32_t)p.offset),
                                    make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
      float z[4];
      box_muller(r.x, r.y, z[0], z[1]);
      box_muller(r.z, r.w, z[2], z[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (4 * q + k < ITEMS) nz[4 * q + k] = z[k];
    }
  }
  const float half_g2 = __fmul_rn(p.g2, 0.5f);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    v[k] = 0.f;
    if (c < D) {
      const float xv = x[base + c];
      if (p.freeze_x) { v[k] = xv; continue; }
      const float sc = __ldg(score + base + c);
      const float z = noise ? __ldg(noise + base + c) : nz[k];
      float dX;
      if (p.debias) {
        const float gu = __ldg(gradU + base + c);
        const float bt = __fmul_rn(sc, half_g2);                             // b_t = s*g^2/2      (sdes.py:168)
        dX = __fadd_rn(__fmul_rn(__fmul_rn(p.gamma, -gu), half_g2), __fmul_rn(p.gamma, bt));  // (:172-174)
        cross = __fadd_rn(cross, __fmul_rn(-gu, bt));                        // <-grad U, b_t>     (:220)
      } else {
        dX = __fmul_rn(p.gamma, __fmul_rn(sc, p.g2));                        // f_not_debiased     (:120-122)
      }
      const float dx = __fadd_rn(__fmul_rn(dX, p.dt), __fmul_rn(__fmul_rn(p.noise_scale, z), p.sqrt_dt));  // sde_integration.py:347
      v[k] = __fadd_rn(xv, dx);
    }
  }
  if (p.remove_mean) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int c = lane + 32 * k;
      if (c < D) { acc0 += comp == 0 ? v[k] : 0.f; acc1 += comp == 1 ? v[k] : 0.f; acc2 += comp == 2 ? v[k] : 0.f; }
      comp = (comp + 2) % 3;
    }
    const float inv = 1.0f / (float)n;
    const float m0 = warp_sum(acc0) * inv, m1 = warp_sum(acc1) * inv, m2 = warp_sum(acc2) * inv;
    comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      v[k] -= comp == 0 ? m0 : (comp == 1 ? m1 : m2);
      comp = (comp + 2) % 3;
    }
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    if (c < D) x_out[base + c] = v[k];
  }
  if (a_raw != nullptr) {
    cross = warp_sum(cross);
    if (lane == 0) {
      float r = 0.f;
      if (p.debias) {
        const float div_b = __fmul_rn(__ldg(div + row), half_g2);           // sdes.py:203
        const float du_dt = __fmul_rn(__ldg(dE_dh + row), p.dh_dt);          // chain rule of :218
        r = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(p.gamma, p.gamma), cross), __fmul_rn(p.gamma, div_b)),
                                __fmul_rn(p.gamma, du_dt)),
                      __fmul_rn(p.dgamma_dt, __ldg(energy + row)));          // :222-227
      }
      a_raw[row] = r;
    }
  }
}

// ---- per-chunk quantile clamp + log-weight accumulation -----------------------------------------
constexpr int kQThreads = 256;
constexpr int kQMax = 8192;

__global__ void __launch_bounds__(kQThreads)
fk_quantile_kernel(const float *__restrict__ a_raw, const float *a, int64_t B, int chunk, float q, float dt, int zero_a,
                   float *a_out, float *__restrict__ drift_A_out) {
  extern __shared__ float s_v[];
  __shared__ float s_q;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int m = (int)min((int64_t)chunk, B - lo);
  int P = 1;
  while (P < m) P <<= 1;
  for (int i = threadIdx.x; i < P; i += kQThreads) s_v[i] = i < m ? a_raw[lo + i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += kQThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float va = s_v[i], vb = s_v[ixj];
          const bool up = (i & k) == 0;
          if ((va > vb) == up) { s_v[i] = vb; s_v[ixj] = va; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    // torch.quantile(v, q), 'linear' interpolation (ATen quantile_compute): rank = q*(m-1) in fp32,
    // lerp(sorted[floor], sorted[ceil], frac) with torch.lerp's two-branch formula.
    const float rank = __fmul_rn(q, (float)(m - 1));
    const float fl = floorf(rank);
    const int ilo = (int)fl, ihi = (int)ceilf(rank);
    const float w = __fsub_rn(rank, fl);
    const float va = s_v[ilo], vb = s_v[ihi];
    const float diff = __fsub_rn(vb, va);
    s_q = w < 0.5f ? __fadd_rn(va, __fmul_rn(w, diff)) : __fsub_rn(vb, __fmul_rn(diff, __fsub_rn(1.0f, w)));
  }
  __syncthreads();
  const float qv = s_q;
  for (int i = threadIdx.x; i < m; i += kQThreads) {
    const float d = fminf(a_raw[lo + i], qv);  // torch.clamp(drift_A, max=quantile)  (sdes.py:230)
    if (drift_A_out) drift_A_out[lo + i] = d;
    float o;
    if (a == nullptr) o = d;
    else o = zero_a ? 0.f : __fadd_rn(a[lo + i], __fmul_rn(d, dt));  // sde_integration.py:349, :278-282
    a_out[lo + i] = o;
  }
}

// ---- post-processing row kernels ------------------------------------------------------------------
// mode 0: descent   x_out = x + f*dt (+ noise*sqrt(2dt))                      (sde_integration.py:353-360)
// mode 1: propose   x_out = x + 0.5*dt*f + sqrt(dt)*noise, aux = log q(x'|x)  (:28-38)
template <int ITEMS>
__global__ void __launch_bounds__(256)
row_update_kernel(int mode, const float *x, const float *__restrict__ f, const float *__restrict__ noise, int64_t B, int n,
                  float dt, float sdt, int remove_mean, float *x_out, float *__restrict__ aux) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int D = 3 * n;
  const int64_t base = row * D;
  float v[ITEMS];
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    v[k] = 0.f;
    if (c < D) {
      const float xv = x[base + c], fv = __ldg(f + base + c);
      if (mode == 0) {
        v[k] = __fadd_rn(xv, __fmul_rn(fv, dt));
        if (noise) v[k] = __fadd_rn(v[k], __fmul_rn(__ldg(noise + base + c), sdt));
      } else {
        const float mean = __fadd_rn(xv, __fmul_rn(__fmul_rn(0.5f, dt), fv));
        v[k] = __fadd_rn(mean, __fmul_rn(sdt, __ldg(noise + base + c)));
        const float dlt = __fsub_rn(v[k], mean);
        q = __fadd_rn(q, __fmul_rn(dlt, dlt));
      }
    }
  }
  if (remove_mean) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int c = lane + 32 * k;
      if (c < D) { acc0 += comp == 0 ? v[k] : 0.f; acc1 += comp == 1 ? v[k] : 0.f; acc2 += comp == 2 ? v[k] : 0.f; }
      comp = (comp + 2) % 3;
    }
    const float inv = 1.0f / (float)n;
    const float m0 = warp_sum(acc0) * inv, m1 = warp_sum(acc1) * inv, m2 = warp_sum(acc2) * inv;
    comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { v[k] -= comp == 0 ? m0 : (comp == 1 ? m1 : m2); comp = (comp + 2) % 3; }
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    if (c < D) x_out[base + c] = v[k];
  }
  if (mode == 1) {
    q = warp_sum(q);
    if (lane == 0) aux[row] = -q / (2.0f * dt);
  }
}

template <int ITEMS>
__global__ void __launch_bounds__(256)
mala_accept_kernel(float *x, float *logp, const float *__restrict__ x_prop, const float *__restrict__ logp_prop,
                   const float *__restrict__ f_prop, const float *__restrict__ log_q_fwd, const float *__restrict__ uniform,
                   int64_t B, int n, float dt, int remove_mean, float *__restrict__ accepted) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int D = 3 * n;
  const int64_t base = row * D;
  float xo[ITEMS], xp[ITEMS];
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    xo[k] = xp[k] = 0.f;
    if (c < D) {
      xo[k] = x[base + c];
      xp[k] = __ldg(x_prop + base + c);
      const float bmean = __fadd_rn(xp[k], __fmul_rn(__fmul_rn(0.5f, dt), __ldg(f_prop + base + c)));  // :42
      const float dlt = __fsub_rn(xo[k], bmean);
      q = __fadd_rn(q, __fmul_rn(dlt, dlt));
    }
  }
  q = warp_sum(q);
  const float log_q_bwd = -q / (2.0f * dt);
  const float lp0 = logp[row], lp1 = __ldg(logp_prop + row);
  const float ratio = (lp1 - lp0) + (log_q_bwd - __ldg(log_q_fwd + row));  // :379-381
  const bool acc = logf(__ldg(uniform + row)) < ratio;                      // :383
  float v[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) v[k] = acc ? xp[k] : xo[k];
  if (remove_mean) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int c = lane + 32 * k;
      if (c < D) { acc0 += comp == 0 ? v[k] : 0.f; acc1 += comp == 1 ? v[k] : 0.f; acc2 += comp == 2 ? v[k] : 0.f; }
      comp = (comp + 2) % 3;
    }
    const float inv = 1.0f / (float)n;
    const float m0 = warp_sum(acc0) * inv, m1 = warp_sum(acc1) * inv, m2 = warp_sum(acc2) * inv;
    comp = lane % 3;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { v[k] -= comp == 0 ? m0 : (comp == 1 ? m1 : m2); comp = (comp + 2) % 3; }
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int c = lane + 32 * k;
    if (c < D) x[base + c] = v[k];
  }
  if (lane == 0) {
    logp[row] = acc ? lp1 : lp0;
    accepted[row] = acc ? 1.0f : 0.0f;
  }
}

static inline int items_for(int n) { return (3 * n + 31) / 32; }

}  // namespace pita

using namespace pita;

#define PITA_DISPATCH_ITEMS(items, CALL)            \
  switch (items) {                                  \
    case 1: { constexpr int IT = 1; CALL; } break;  \
    case 2: { constexpr int IT = 2; CALL; } break;  \
    case 3: { constexpr int IT = 3; CALL; } break;  \
    case 4: { constexpr int IT = 4; CALL; } break;  \
    case 5: { constexpr int IT = 5; CALL; } break;  \
    case 6: { constexpr int IT = 6; CALL; } break;  \
    case 7: { constexpr int IT = 7; CALL; } break;  \
    case 8: { constexpr int IT = 8; CALL; } break;  \
    default: set_error("row kernels support 3n <= 256 (n=%d)", n); return PITA_EUNSUP; \
  }

extern "C" int pita_sde_fk_step(const float *x, const float *gradU, const float *score, const float *noise,
                                const float *div, const float *dE_dh, const float *energy, int64_t B, int n,
                                const pita_sde_params *ph, float *x_out, float *a_raw, void *stream) {
  PITA_REQUIRE(x && x_out && ph, PITA_EINVAL, "sde_fk_step: null pointer");
  PITA_REQUIRE(ph->freeze_x || score, PITA_EINVAL, "sde_fk_step: score is required unless freeze_x");
  PITA_REQUIRE(!ph->debias || ph->freeze_x || gradU, PITA_EINVAL, "sde_fk_step: gradU is required when debias=1");
  PITA_REQUIRE(!ph->debias || !a_raw || (div && dE_dh && energy), PITA_EINVAL,
               "sde_fk_step: div, dE_dh and energy are required for the FK weight (debias=1, a_raw != NULL)");
  PITA_REQUIRE(n > 0 && B >= 0, PITA_EINVAL, "sde_fk_step: bad sizes");
  if (B == 0) return PITA_OK;
  SdeP p{ph->g2, ph->gamma, ph->dgamma_dt, ph->dh_dt, ph->dt, ph->sqrt_dt, ph->noise_scale,
         ph->debias, ph->freeze_x, ph->remove_mean, ph->seed, ph->offset};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((B + 7) / 8);
  PITA_DISPATCH_ITEMS(items_for(n), (sde_fk_step_kernel<IT><<<grid, 256, 0, st>>>(x, gradU, score, noise, div, dE_dh, energy, B, n, p, x_out, a_raw)));
  PITA_CHECK_LAUNCH("sde_fk_step_kernel");
  return PITA_OK;
}

extern "C" int pita_fk_quantile_accumulate(const float *a_raw, const float *a, int64_t B, int chunk, float q, float dt,
                                           int zero_a, float *a_out, float *drift_A_out, void *stream) {
  PITA_REQUIRE(a_raw && a_out, PITA_EINVAL, "fk_quantile: null pointer");
  PITA_REQUIRE(chunk >= 1 && chunk <= kQMax, PITA_EINVAL, "fk_quantile: chunk must be in [1,%d]", kQMax);
  PITA_REQUIRE(q >= 0.f && q <= 1.f, PITA_EINVAL, "fk_quantile: q must be in [0,1]");
  if (B <= 0) return B == 0 ? PITA_OK : PITA_EINVAL;
  int P = 1;
  while (P < chunk) P <<= 1;
  const int64_t blocks = (B + chunk - 1) / chunk;
  fk_quantile_kernel<<<(unsigned)blocks, kQThreads, P * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      a_raw, a, B, chunk, q, dt, zero_a, a_out, drift_A_out);
  PITA_CHECK_LAUNCH("fk_quantile_kernel");
  return PITA_OK;
}

extern "C" int pita_descent_step(const float *x, const float *force, const float *noise, int64_t B, int n, float dt,
                                 int remove_mean, float *x_out, void *stream) {
  PITA_REQUIRE(x && force && x_out, PITA_EINVAL, "descent_step: null pointer");
  PITA_REQUIRE(n > 0 && B >= 0, PITA_EINVAL, "descent_step: bad sizes");
  if (B == 0) return PITA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((B + 7) / 8);
  const float sdt = (float)sqrt(2.0 * (double)dt);
  PITA_DISPATCH_ITEMS(items_for(n), (row_update_kernel<IT><<<grid, 256, 0, st>>>(0, x, force, noise, B, n, dt, sdt, remove_mean, x_out, nullptr)));
  PITA_CHECK_LAUNCH("row_update_kernel(descent)");
  return PITA_OK;
}

extern "C" int pita_mala_propose(const float *x, const float *force, const float *noise, int64_t B, int n, float dt,
                                 float *x_prop, float *log_q_fwd, void *stream) {
  PITA_REQUIRE(x && force && noise && x_prop && log_q_fwd, PITA_EINVAL, "mala_propose: null pointer");
  PITA_REQUIRE(n > 0 && B >= 0 && dt > 0.f, PITA_EINVAL, "mala_propose: bad sizes");
  if (B == 0) return PITA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((B + 7) / 8);
  const float sdt = sqrtf(dt);  // torch.sqrt(torch.tensor(dt))
  PITA_DISPATCH_ITEMS(items_for(n), (row_update_kernel<IT><<<grid, 256, 0, st>>>(1, x, force, noise, B, n, dt, sdt, 0, x_prop, log_q_fwd)));
  PITA_CHECK_LAUNCH("row_update_kernel(propose)");
  return PITA_OK;
}

extern "C" int pita_mala_accept(float *x, float *logp, const float *x_prop, const float *logp_prop, const float *force_prop,
                                const float *log_q_fwd, const float *uniform, int64_t B, int n, float dt, int remove_mean,
                                float *accepted, void *stream) {
  PITA_REQUIRE(x && logp && x_prop && logp_prop && force_prop && log_q_fwd && uniform && accepted, PITA_EINVAL, "mala_accept: null pointer");
  PITA_REQUIRE(n > 0 && B >= 0 && dt > 0.f, PITA_EINVAL, "mala_accept: bad sizes");
  if (B == 0) return PITA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((B + 7) / 8);
  PITA_DISPATCH_ITEMS(items_for(n), (mala_accept_kernel<IT><<<grid, 256, 0, st>>>(x, logp, x_prop, logp_prop, force_prop, log_q_fwd, uniform, B, n, dt, remove_mean, accepted)));
  PITA_CHECK_LAUNCH("mala_accept_kernel");
  return PITA_OK;
}
