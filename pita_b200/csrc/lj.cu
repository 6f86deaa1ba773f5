// Lennard-Jones target energy + analytic force, fused (replaces lennardjones_energy.py:121-155,213-227).
//
// One thread per (configuration, atom); a CTA stages CPB whole configurations in shared memory as
// padded float4 (coalesced float4 global loads), each thread walks the n-1 partners of its atom
// (ordered pairs: no atomics, deterministic), per-configuration energy / centre of mass are reduced
// with warp shuffles, forces leave through shared memory as coalesced float4 stores.
//
// Algorithmic work per configuration: n(n-1)/2 * 31 + 15n FLOP, 24n + 4 bytes.
#include "common.cuh"

namespace pita {

constexpr float kBgflowEps = 1e-6f;  // distances_from_vectors(eps=1e-6), bgflow/utils/geometry.py

template <int NA, int CPB>
struct LJCfg {
  static constexpr int kItems = NA * CPB;
  static constexpr int kThreads = (kItems + 31) / 32 * 32;
  static constexpr int kWarps = kThreads / 32;
  static_assert((CPB * NA * 3) % 4 == 0, "block tile must be a whole number of float4");
};

template <int NA, int CPB>
__global__ void __launch_bounds__(LJCfg<NA, CPB>::kThreads)
lj_energy_force_kernel(const float *__restrict__ x, int64_t B, float inv_T, float energy_factor, float osc,
                       float *__restrict__ logp, float *__restrict__ force) {
  using C = LJCfg<NA, CPB>;
  constexpr int D = 3 * NA;
  __shared__ float4 s_pos[C::kItems];
  __shared__ float s_val[C::kItems];
  __shared__ float4 s_com[CPB];
  __shared__ __align__(16) float s_f[C::kItems * 3];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t cfg0 = (int64_t)blockIdx.x * CPB;
  const int ncfg = (int)min((int64_t)CPB, B - cfg0);
  const int nflt = ncfg * D;
  const float *__restrict__ src = x + cfg0 * D;

  // ---- stage coordinates: coalesced float4 loads, scattered into the padded float4 layout
  {
    float *sp = reinterpret_cast<float *>(s_pos);
    const int nvec = nflt >> 2;
    for (int v = tid; v < nvec; v += C::kThreads) {
      const float4 q = __ldg(reinterpret_cast<const float4 *>(src) + v);
      const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = 4 * v + u;
        const int atom = f / 3, k = f - 3 * atom;
        sp[atom * 4 + k] = e[u];
      }
    }
    for (int f = (nvec << 2) + tid; f < nflt; f += C::kThreads) {
      const int atom = f / 3, k = f - 3 * atom;
      sp[atom * 4 + k] = __ldg(src + f);
    }
  }
  __syncthreads();

  // ---- centre of mass per configuration (one warp per configuration, shuffle reduction)
  for (int c = warp; c < ncfg; c += C::kWarps) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int a = lane; a < NA; a += 32) {
      const float4 p = s_pos[c * NA + a];
      sx += p.x; sy += p.y; sz += p.z;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) s_com[c] = make_float4(sx * (1.0f / NA), sy * (1.0f / NA), sz * (1.0f / NA), 0.f);
  }
  __syncthreads();

  // ---- pair loop: thread = (configuration c, atom i), all partners j != i
  const int c = tid / NA;
  const int i = tid - c * NA;
  const bool active = tid < C::kItems && c < ncfg;
  if (active) {
    const float4 *__restrict__ cfg = s_pos + c * NA;
    const float4 pi = cfg[i];
    float e6 = 0.f, e3 = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll 6
    for (int jj = 0; jj < NA - 1; ++jj) {
      const int j = jj + (jj >= i ? 1 : 0);
      const float4 pj = cfg[j];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const float s = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, kBgflowEps)));
      const float inv = fast_rcp(s);
      const float i3 = inv * inv * inv;
      const float i6 = i3 * i3;
      e6 += i6;
      e3 += i3;
      const float fs = (i6 - i3) * inv;
      fx = fmaf(fs, dx, fx);
      fy = fmaf(fs, dy, fy);
      fz = fmaf(fs, dz, fz);
    }
    const float4 com = s_com[c];
    const float cx = pi.x - com.x, cy = pi.y - com.y, cz = pi.z - com.z;
    s_val[tid] = energy_factor * (e6 - 2.0f * e3) + osc * 0.5f * (cx * cx + cy * cy + cz * cz);
    // d logp / d x_i = -(1/T) [ ef * (-24) * sum_j fs*d  + osc*(x_i - com) ]
    const float k24 = 24.0f * energy_factor * inv_T;
    const float ko = osc * inv_T;
    s_f[tid * 3 + 0] = fmaf(k24, fx, -ko * cx);
    s_f[tid * 3 + 1] = fmaf(k24, fy, -ko * cy);
    s_f[tid * 3 + 2] = fmaf(k24, fz, -ko * cz);
  }
  __syncthreads();

  // ---- per-configuration energy (warp per configuration) and coalesced force store
  for (int cc = warp; cc < ncfg; cc += C::kWarps) {
    float v = 0.f;
    for (int a = lane; a < NA; a += 32) v += s_val[cc * NA + a];
    v = warp_sum(v);
    if (lane == 0) logp[cfg0 + cc] = -v * inv_T;
  }
  if (force != nullptr) {
    float *__restrict__ dst = force + cfg0 * D;
    const int nvec = nflt >> 2;
    for (int v = tid; v < nvec; v += C::kThreads)
      reinterpret_cast<float4 *>(dst)[v] = reinterpret_cast<const float4 *>(s_f)[v];
    for (int f = (nvec << 2) + tid; f < nflt; f += C::kThreads) dst[f] = s_f[f];
  }
}

template <int NA, int CPB>
static int launch_lj(const float *x, int64_t B, float T, float ef, float osc, float *logp, float *force,
                     cudaStream_t st) {
  using C = LJCfg<NA, CPB>;
  const int64_t blocks = (B + CPB - 1) / CPB;
  PITA_REQUIRE(blocks < (1ll << 31), PITA_EINVAL, "lj: batch too large");
  lj_energy_force_kernel<NA, CPB><<<(unsigned)blocks, C::kThreads, 0, st>>>(x, B, 1.0f / T, ef, osc, logp, force);
  PITA_CHECK_LAUNCH("lj_energy_force_kernel");
  return PITA_OK;
}

}  // namespace pita

extern "C" int pita_lj_energy_force(const float *x, int64_t B, int n, float temperature, float energy_factor,
                                    float oscillator_scale, float *logp, float *force, void *stream) {
  using namespace pita;
  PITA_REQUIRE(B >= 0, PITA_EINVAL, "lj: negative batch");
  if (B == 0) return PITA_OK;  // empty batch: nothing to read or write (pointers may be NULL)
  PITA_REQUIRE(x && logp, PITA_EINVAL, "lj: null pointer");
  PITA_REQUIRE(aligned16(x) && (force == nullptr || aligned16(force)), PITA_EINVAL, "lj: pointers must be 16-byte aligned");
  PITA_REQUIRE(temperature > 0.f, PITA_EINVAL, "lj: temperature must be positive");
  if (B == 0) return PITA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (n) {
    case 13: return launch_lj<13, 32>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
    case 55: return launch_lj<55, 8>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
    default:
      set_error("lj: n_particles=%d unsupported (reference raises NotImplementedError for n not in {13,55}, "
                "lennardjones_energy.py:177-178)", n);
      return PITA_EUNSUP;
  }
}
