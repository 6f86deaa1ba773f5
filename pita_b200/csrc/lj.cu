// Lennard-Jones target energy + analytic force, fused (replaces lennardjones_energy.py:121-155,213-227).
//
// One thread per (configuration, atom); a CTA stages CPB whole configurations in shared memory as
// padded float4 (coalesced float4 global loads), each thread walks the n-1 partners of its atom
// (ordered pairs: no atomics, deterministic), per-configuration energy / centre of mass are reduced
// with warp shuffles, forces leave through shared memory as coalesced float4 stores.
//
// Algorithmic work per configuration: n(n-1)/2 * 31 + 15n FLOP, 24n + 4 bytes.
#include <stdlib.h>

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


// =============================================================================================================
// "Paired" kernel (default): every UNORDERED pair is evaluated once and its force applied to both atoms
// (Newton's third law), in packed fp32 (FFMA2 / FADD2 / FMUL2: two pairs per instruction, the only way sm_100a
// reaches its FP32 peak -- profiles/ubench).  18 packed instructions + 2 MUFU.RCP per two pairs
// => 31 algorithmic FLOP in 9 FMA-pipe issue slots (scalar ordered-pair kernel above: 32 slots).
//
// Mapping: lane = configuration (32 configurations per warp-set), warp = atom group g (G groups of S = NA/G atoms).
// Thread (c, g) keeps its S atoms and their force accumulators in registers and streams atoms b = gS + t,
// t = 1 .. S-1+K (K = (NA-1)/2, circulant half-neighbourhood: atom a pairs with a+1 .. a+K mod NA, which covers
// every unordered pair exactly once for odd NA) from shared memory two at a time; the pair (a = gS + r, b) is
// evaluated iff 1 <= t - r <= K (static after unrolling; a half-valid packed pair is neutralised by adding 1e30
// to that half's squared distance).  Reactions on b leave through shared memory "slots" q = (t-1)/S: for a fixed
// slot every atom has exactly one writer, so there are no atomics and the summation order is fixed (deterministic).
// =============================================================================================================
struct f2 {
  unsigned long long v;
};
__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}

template <int NA, int G, int CW>
struct LJPairCfg {
  static_assert(NA % 2 == 1 && NA % G == 0, "circulant pairing needs an odd atom count split into equal groups");
  static constexpr int S = NA / G;            // atoms owned by a thread
  static constexpr int K = (NA - 1) / 2;      // partners per atom
  static constexpr int T = S - 1 + K;         // streamed atoms per thread
  static constexpr int QMAX = T / S;          // streamed atom t belongs to group g + t/S; its reaction goes to slot t/S
  static constexpr int RLAST = T - QMAX * S + 1;  // atoms per group touched in the last (partial) slot
  static constexpr int CPB = 32 * CW;         // configurations per CTA
  static constexpr int kThreads = 32 * G * CW;
  static constexpr int kPosBytes = CPB * NA * 12;
  // slot q (1 <= q <= QMAX, q % G != 0) holds one reaction vector per (configuration, atom); the last slot only RLAST atoms
  // per group.  Slots with q % G == 0 target the thread's own atoms and stay in registers.
  __host__ __device__ static constexpr int slot_pitch(int q) { return q == QMAX ? 3 * G * RLAST : 3 * NA; }
  __host__ __device__ static constexpr int slot_base(int q) {   // in floats, from the start of the slot area
    int o = 0;
    for (int i = 1; i < q; ++i) o += (i % G == 0) ? 0 : CPB * slot_pitch(i);
    return o;
  }
  static constexpr int kSlotFloats = slot_base(QMAX + 1);
  static constexpr int kSmemBytes = kPosBytes + kSlotFloats * 4 + G * CPB * 16;   // one [G][CPB] float4 area, used twice
};

template <int NA, int G, int CW, bool FORCE, int MINB>
__global__ void __launch_bounds__(LJPairCfg<NA, G, CW>::kThreads, MINB)
lj_pairs_kernel(const float *__restrict__ x, int64_t B, float inv_T, float energy_factor, float osc,
                float *__restrict__ logp, float *__restrict__ force) {
  using C = LJPairCfg<NA, G, CW>;
  constexpr int S = C::S, K = C::K, T = C::T, D = 3 * NA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_pos = reinterpret_cast<float *>(smem_raw);                     // [CPB][3*NA], the global layout
  float *s_react = reinterpret_cast<float *>(smem_raw + C::kPosBytes);    // reaction slots
  float4 *s_part = reinterpret_cast<float4 *>(s_react + C::kSlotFloats);  // [G][CPB]: (sum x, sum y, sum z), later .w = energy

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = warp % G, cset = warp / G;
  const int c = cset * 32 + lane;
  const int64_t cfg0 = (int64_t)blockIdx.x * C::CPB;
  const int ncfg = (int)min((int64_t)C::CPB, B - cfg0);
  const int nflt = ncfg * D;
  const float *__restrict__ src = x + cfg0 * D;

  // ---- stage: straight float4 copy; a configuration's row pitch (3*NA floats) is odd, so lane = configuration
  //      makes every scalar LDS / STS below bank-conflict free
  if (ncfg == C::CPB && (C::CPB * D) % 4 == 0) {
    // full tile: every load is issued before the first store (the plain loop below waits for each load in turn: with a dozen
    // iterations that latency was most of an LJ-13 CTA's life)
    constexpr int NV = C::CPB * D / 4, PER = (NV + C::kThreads - 1) / C::kThreads;
    float4 buf[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (tid + k * C::kThreads < NV) buf[k] = __ldg(reinterpret_cast<const float4 *>(src) + tid + k * C::kThreads);
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (tid + k * C::kThreads < NV) reinterpret_cast<float4 *>(s_pos)[tid + k * C::kThreads] = buf[k];
  } else {
    const int nvec = nflt >> 2;
    for (int v = tid; v < nvec; v += C::kThreads)
      reinterpret_cast<float4 *>(s_pos)[v] = __ldg(reinterpret_cast<const float4 *>(src) + v);
    for (int f = (nvec << 2) + tid; f < nflt; f += C::kThreads) s_pos[f] = __ldg(src + f);
  }
  __syncthreads();

  const bool active = c < ncfg;
  const float *__restrict__ cfg = s_pos + c * D;
  float ax[S], ay[S], az[S];
  f2 fx[S], fy[S], fz[S];   // packed partial sums of P = -fs*d over the pairs where the atom is on the "a" side
  float e_pairs = 0.f;
  if (active) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int r = 0; r < S; ++r) {
      ax[r] = cfg[3 * (g * S + r) + 0]; ay[r] = cfg[3 * (g * S + r) + 1]; az[r] = cfg[3 * (g * S + r) + 2];
      sx += ax[r]; sy += ay[r]; sz += az[r];
      fx[r] = pk2(0.f, 0.f); fy[r] = pk2(0.f, 0.f); fz[r] = pk2(0.f, 0.f);
    }
    s_part[g * C::CPB + c] = make_float4(sx, sy, sz, 0.f);

    const f2 eps2 = pk2(kBgflowEps, kBgflowEps);
    f2 e6 = pk2(0.f, 0.f), en3 = pk2(0.f, 0.f);
    // The streamed pair of step s+1 is loaded BEFORE the reaction stores of step s in program order: loads cannot be moved above
    // stores that may alias them, so without this every step drains its dependent FMA -> RCP -> FMA chains before the next one's
    // operands even leave shared memory (r1d ncu: stall_wait 1.7 per issue, FMA pipe 62 %).
    auto load_pair = [&](int t0, f2 &px, f2 &py, f2 &pz) {
      const int t1 = t0 + 1;
      int b0 = g * S + t0; b0 -= (b0 >= NA) ? NA : 0;
      int b1 = g * S + (t1 <= T ? t1 : t0); b1 -= (b1 >= NA) ? NA : 0;
      // scalar loads straight into the halves of the packed operands (a vector load would need re-packing moves)
      px = pk2(lds1(cfg + 3 * b0 + 0), lds1(cfg + 3 * b1 + 0));
      py = pk2(lds1(cfg + 3 * b0 + 1), lds1(cfg + 3 * b1 + 1));
      pz = pk2(lds1(cfg + 3 * b0 + 2), lds1(cfg + 3 * b1 + 2));
    };
    f2 nbx, nby, nbz;
    load_pair(1, nbx, nby, nbz);
#pragma unroll
    for (int t0 = 1; t0 <= T; t0 += 2) {
      const int t1 = t0 + 1;
      const bool has1 = t1 <= T;
      const f2 bx = nbx, by = nby, bz = nbz;
      if (t0 + 2 <= T) load_pair(t0 + 2, nbx, nby, nbz);
      f2 rx = pk2(0.f, 0.f), ry = pk2(0.f, 0.f), rz = pk2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const bool v0 = (t0 - r >= 1) && (t0 - r <= K);
        const bool v1 = has1 && (t1 - r >= 1) && (t1 - r <= K);
        if (v0 || v1) {
          const f2 dx = sub2(pk2(ax[r], ax[r]), bx), dy = sub2(pk2(ay[r], ay[r]), by), dz = sub2(pk2(az[r], az[r]), bz);
          f2 s = fma2(dx, dx, eps2);
          s = fma2(dy, dy, s);
          s = fma2(dz, dz, s);
          if (!v0) s = add2(s, pk2(1e30f, 0.f));
          if (!v1) s = add2(s, pk2(0.f, 1e30f));
          float s_lo, s_hi;
          unpk2(s, s_lo, s_hi);
          const f2 ninv = pk2(fast_rcp(-s_lo), fast_rcp(-s_hi));   // -1/s
          const f2 inv2 = mul2(ninv, ninv);                         //  1/s^2
          const f2 ni3 = mul2(inv2, ninv);                          // -1/s^3
          e6 = fma2(ni3, ni3, e6);                                  // sum r^-12
          en3 = add2(en3, ni3);                                     // -sum r^-6
          if (FORCE) {
            const f2 w = mul2(inv2, inv2);                          //  1/s^4
            const f2 nfs = fma2(w, ni3, w);                         // -(s^-6 - s^-3)/s
            fx[r] = fma2(nfs, dx, fx[r]); fy[r] = fma2(nfs, dy, fy[r]); fz[r] = fma2(nfs, dz, fz[r]);
            rx = fma2(nfs, dx, rx); ry = fma2(nfs, dy, ry); rz = fma2(nfs, dz, rz);
          }
        }
      }
      if (FORCE) {
        // reaction on the streamed atoms: +fs*d(b) = the same product P, but the owner's accumulator carries -P
        float xl, xh, yl, yh, zl, zh;
        unpk2(rx, xl, xh); unpk2(ry, yl, yh); unpk2(rz, zl, zh);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int t = half ? t1 : t0;
          if (half && !has1) continue;
          const int q = t / S, r_t = t % S;
          const float vx = half ? xh : xl, vy = half ? yh : yl, vz = half ? zh : zl;
          if (q % G == 0) {                      // one of this thread's own atoms: registers
            float lo, hi;
            unpk2(fx[r_t], lo, hi); fx[r_t] = pk2(lo - vx, hi);
            unpk2(fy[r_t], lo, hi); fy[r_t] = pk2(lo - vy, hi);
            unpk2(fz[r_t], lo, hi); fz[r_t] = pk2(lo - vz, hi);
          } else {
            int gt = g + q % G; gt -= (gt >= G) ? G : 0;
            float *d = s_react + C::slot_base(q) + c * C::slot_pitch(q) +
                       (q == C::QMAX ? 3 * (gt * C::RLAST + r_t) : 3 * (gt * S + r_t));
            d[0] = vx; d[1] = vy; d[2] = vz;
          }
        }
      }
    }
    float a_lo, a_hi, b_lo, b_hi;
    unpk2(e6, a_lo, a_hi);
    unpk2(en3, b_lo, b_hi);
    e_pairs = (a_lo + a_hi) + 2.0f * (b_lo + b_hi);  // sum over this thread's unordered pairs of r^-12 - 2 r^-6
  }
  __syncthreads();

  // ---- centre of mass, harmonic term, reaction gather
  float fo[FORCE ? 3 * S : 1];
  float e_total = 0.f;
  if (active) {
    float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
    for (int gg = 0; gg < G; ++gg) {
      const float4 q = s_part[gg * C::CPB + c];
      cx += q.x; cy += q.y; cz += q.z;
    }
    cx *= (1.0f / NA); cy *= (1.0f / NA); cz *= (1.0f / NA);
    const float k24 = 24.0f * energy_factor * inv_T, ko = osc * inv_T;
    float e_osc = 0.f;
#pragma unroll
    for (int r = 0; r < S; ++r) {
      const float ux = ax[r] - cx, uy = ay[r] - cy, uz = az[r] - cz;
      e_osc = fmaf(ux, ux, fmaf(uy, uy, fmaf(uz, uz, e_osc)));
      if (FORCE) {
        float lo, hi, px, py, pz;
        unpk2(fx[r], lo, hi); px = -(lo + hi);
        unpk2(fy[r], lo, hi); py = -(lo + hi);
        unpk2(fz[r], lo, hi); pz = -(lo + hi);
#pragma unroll
        for (int q = 1; q <= C::QMAX; ++q) {
          if (q % G != 0 && q * S + r <= T) {    // thread g - q streamed this atom at step t = q*S + r
            const float *sp = s_react + C::slot_base(q) + c * C::slot_pitch(q) +
                              (q == C::QMAX ? 3 * (g * C::RLAST + r) : 3 * (g * S + r));
            px += sp[0]; py += sp[1]; pz += sp[2];
          }
        }
        fo[3 * r + 0] = fmaf(k24, px, -ko * ux);
        fo[3 * r + 1] = fmaf(k24, py, -ko * uy);
        fo[3 * r + 2] = fmaf(k24, pz, -ko * uz);
      }
    }
    // reference energy sums ORDERED pairs (= 2 x unordered), lennardjones_energy.py:121-140
    e_total = energy_factor * 2.0f * e_pairs + osc * 0.5f * e_osc;
  }
  __syncthreads();   // every group has read the centre-of-mass partials: the area is reused for the energy partials
  if (active) s_part[g * C::CPB + c].w = e_total;
  __syncthreads();
  if (g == 0 && active) {
    float v = 0.f;
#pragma unroll
    for (int gg = 0; gg < G; ++gg) v += s_part[gg * C::CPB + c].w;
    logp[cfg0 + c] = -v * inv_T;
  }
  if (FORCE) {
    float *s_out = s_pos;  // coordinates are dead (registers hold them): reuse as the coalescing stage
    if (active) {
#pragma unroll
      for (int r = 0; r < S; ++r) {
        float *o = s_out + c * D + 3 * (g * S + r);
        o[0] = fo[3 * r + 0]; o[1] = fo[3 * r + 1]; o[2] = fo[3 * r + 2];
      }
    }
    __syncthreads();
    float *__restrict__ dst = force + cfg0 * D;
    const int nvec = nflt >> 2;
    for (int v = tid; v < nvec; v += C::kThreads)
      reinterpret_cast<float4 *>(dst)[v] = reinterpret_cast<const float4 *>(s_out)[v];
    for (int f = (nvec << 2) + tid; f < nflt; f += C::kThreads) dst[f] = s_out[f];
  }
}

// =============================================================================================================
// "Blocked" paired kernel (default for LJ-55).  Same arithmetic per unordered pair as lj_pairs_kernel, different work split:
//   * G = 8 atom groups of S = 7 (55 atoms + one ghost parked at 1e15, whose pair terms flush to exactly 0),
//     warp = group, lane = configuration: 8 warps per CTA and two CTAs per SM = 16 warps, four on every warp scheduler.
//     Why: the five-warp CTAs of lj_pairs_kernel<55,5,..> load the four schedulers 3:3:2:2 at best (10/12 = 83 %, the
//     measured ceiling of the pair chain at 5 x 2 warps, profiles/r2o_ubench_lj_chain.txt) and ptxas emits each pair's
//     FMA -> MUFU.RCP -> FMUL -> FMA chain back to back, so with 2.5 warps per scheduler the FMA pipe idles 38 % of the
//     time (profiles/r1d_ncu_lj55.txt).  The same chain at 16 warps per SM reaches 96 % in the micro-benchmark without
//     any help from the instruction scheduler.
//   * the unordered pairs are tiled by group: thread g evaluates its own triangle (21 pairs), the three rectangles
//     (g, g+1), (g, g+2), (g, g+3) in full (49 pairs each) and its share of (g, g+4): groups 0-3 take their seven atoms
//     against atoms 0-3 of group g+4, groups 4-7 their atoms 4-6 against all seven atoms of group g-4.
//   * reactions on streamed atoms leave through shared memory, one slot per group offset (1, 2, 3, 4): every slot entry
//     has one writer and the gather order is fixed -- deterministic, no atomics.  The position rows of a tile double as the
//     coalescing stage of its force rows.
//   * NOT persistent: a two-CTA-per-SM persistent variant with cp.async double buffering and two reaction slots measured
//     8 % slower (profiles/r2s_lj_ab_persistent_variant.txt) -- co-resident persistent CTAs run their barrier / gather / copy-out phases in
//     lock step, while independently scheduled CTAs drift apart and fill each other's bubbles.
// =============================================================================================================
constexpr float kGhost = 1e15f;   // a padding atom this far away contributes exactly 0 to every sum (1/s^2 underflows)

template <int NA>
struct LJBlockCfg {
  static constexpr int G = 8, S = 7, NP = G * S;       // padded atom count
  static_assert(NA == 55, "the blocked kernel is laid out for 55 atoms (56 = 8 x 7 with one ghost)");
  static constexpr int CPB = 32, kThreads = 32 * G;
  static constexpr int kPitch = 3 * NA;                // the global layout; odd pitch: lane = configuration is bank-conflict free
  static constexpr int kPosFloats = CPB * kPitch + 4;  // (+ the three floats a ghost load of the last row touches)
  static constexpr int kSlots = 4;
  static constexpr int kReactFloats = kSlots * NP * 3 * CPB;   // [slot][atom][xyz][lane]
  static constexpr int kSmemBytes = (kPosFloats + kReactFloats) * 4 + G * CPB * 16 + G * CPB * 4;
};

// One packed step: two streamed atoms (lo / hi halves of bx, by, bz) against the own atoms r in [LO0, HI0] (lo half) and
// [LO1, HI1] (hi half; HI1 < LO1 = no second atom).  Accumulates the owners' force sums and the pair energies and returns
// the reactions on the two streamed atoms in (rx, ry, rz).
template <int S, int LO0, int HI0, int LO1, int HI1, bool FORCE>
__device__ __forceinline__ void lj_block_step(const float (&ax)[S], const float (&ay)[S], const float (&az)[S], f2 (&fx)[S],
                                              f2 (&fy)[S], f2 (&fz)[S], f2 bx, f2 by, f2 bz, f2 &e6, f2 &en3, f2 &rx, f2 &ry,
                                              f2 &rz) {
  const f2 eps2 = pk2(kBgflowEps, kBgflowEps);
  rx = pk2(0.f, 0.f); ry = rx; rz = rx;
#pragma unroll
  for (int r = 0; r < S; ++r) {
    const bool v0 = r >= LO0 && r <= HI0, v1 = r >= LO1 && r <= HI1;
    if (v0 || v1) {
      const f2 dx = sub2(pk2(ax[r], ax[r]), bx), dy = sub2(pk2(ay[r], ay[r]), by), dz = sub2(pk2(az[r], az[r]), bz);
      f2 s = fma2(dx, dx, eps2);
      s = fma2(dy, dy, s);
      s = fma2(dz, dz, s);
      if (!v0) s = add2(s, pk2(1e30f, 0.f));
      if (!v1) s = add2(s, pk2(0.f, 1e30f));
      float s_lo, s_hi;
      unpk2(s, s_lo, s_hi);
      const f2 ninv = pk2(fast_rcp(-s_lo), fast_rcp(-s_hi));   // -1/s
      const f2 inv2 = mul2(ninv, ninv);                         //  1/s^2
      const f2 ni3 = mul2(inv2, ninv);                          // -1/s^3
      e6 = fma2(ni3, ni3, e6);                                  // sum r^-12
      en3 = add2(en3, ni3);                                     // -sum r^-6
      if (FORCE) {
        const f2 w = mul2(inv2, inv2);                          //  1/s^4
        const f2 nfs = fma2(w, ni3, w);                         // -(s^-6 - s^-3)/s
        fx[r] = fma2(nfs, dx, fx[r]); fy[r] = fma2(nfs, dy, fy[r]); fz[r] = fma2(nfs, dz, fz[r]);
        rx = fma2(nfs, dx, rx); ry = fma2(nfs, dy, ry); rz = fma2(nfs, dz, rz);
      }
    }
  }
}

#pragma nv_diag_suppress 549   // own-atom registers are set and used only by active lanes
template <int NA, bool FORCE>
__global__ void __launch_bounds__(LJBlockCfg<NA>::kThreads, 2)
lj_blocked_kernel(const float *__restrict__ x, int64_t B, float inv_T, float energy_factor, float osc,
                  float *__restrict__ logp, float *__restrict__ force) {
  using C = LJBlockCfg<NA>;
  constexpr int S = C::S, G = C::G, NP = C::NP, D = 3 * NA, P = C::kPitch;
  constexpr int NV = C::CPB * D / 4, PER = (NV + C::kThreads - 1) / C::kThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_pos = reinterpret_cast<float *>(smem_raw);          // [CPB][3*NA], the global layout
  float *s_react = s_pos + C::kPosFloats;                      // [slot][atom][xyz][lane]
  float4 *s_part = reinterpret_cast<float4 *>(s_react + C::kReactFloats);   // [G][CPB] centre-of-mass partial sums
  float *s_en = reinterpret_cast<float *>(s_part + G * C::CPB);               // [G][CPB] energy partial sums

  const int tid = threadIdx.x, c = tid & 31, g = tid >> 5;
  const int a0 = g * S;
  {
    const int tile = blockIdx.x;
    const int ncfg = (int)min((int64_t)C::CPB, B - (int64_t)tile * C::CPB);
    const int nflt = ncfg * D;
    const float *__restrict__ src = x + (int64_t)tile * C::CPB * D;
    if (ncfg == C::CPB) {   // full tile: every load is issued before the first store (a plain loop waits for each in turn)
      float4 buf[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (tid + k * C::kThreads < NV) buf[k] = __ldg(reinterpret_cast<const float4 *>(src) + tid + k * C::kThreads);
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (tid + k * C::kThreads < NV) reinterpret_cast<float4 *>(s_pos)[tid + k * C::kThreads] = buf[k];
    } else {
      for (int f = tid; f < nflt; f += C::kThreads) s_pos[f] = __ldg(src + f);
    }
    __syncthreads();

    const bool active = c < ncfg;
    const float *__restrict__ cfg = s_pos + c * P;
    float ax[S], ay[S], az[S];
    f2 fx[S], fy[S], fz[S];
    f2 e6 = pk2(0.f, 0.f), en3 = pk2(0.f, 0.f), rx, ry, rz;
    float xl, xh, yl, yh, zl, zh;
    f2 bx, by, bz, nx, ny, nz;
#define PITA_UNPACK_R() do { unpk2(rx, xl, xh); unpk2(ry, yl, yh); unpk2(rz, zl, zh); } while (0)
    // streamed atom q of the rectangles: group g + 1 + q/7 (mod 8), atom q % 7 of it
    auto rect_atom = [&](int q) { int gb = g + 1 + q / S; gb -= (gb >= G) ? G : 0; return gb * S + q % S; };
    auto half_atom = [&](int u) { int gb = g + 4; gb -= (gb >= G) ? G : 0; return gb * S + u; };
    // atom index NA is the ghost: its coordinates are replaced after the load (warp-uniform select)
    auto load2 = [&](int b0, int b1, f2 &px, f2 &py, f2 &pz) {
      float x0 = lds1(cfg + 3 * b0 + 0), y0 = lds1(cfg + 3 * b0 + 1), z0 = lds1(cfg + 3 * b0 + 2);
      float x1 = lds1(cfg + 3 * b1 + 0), y1 = lds1(cfg + 3 * b1 + 1), z1 = lds1(cfg + 3 * b1 + 2);
      if (b0 == NA) { x0 = kGhost; y0 = kGhost; z0 = kGhost; }
      if (b1 == NA) { x1 = kGhost; y1 = kGhost; z1 = kGhost; }
      px = pk2(x0, x1); py = pk2(y0, y1); pz = pk2(z0, z1);
    };
    // Reactions: slot = group offset - 1; every slot entry has exactly one writer.
    auto put = [&](int slot, int atom, float vx, float vy, float vz) {
      float *d = s_react + ((slot * NP + atom) * 3) * C::CPB + c;
      d[0] = vx; d[C::CPB] = vy; d[2 * C::CPB] = vz;
    };
    auto own_react = [&](int r, float vx, float vy, float vz) {   // the owner's accumulator carries -P
      float lo, hi;
      unpk2(fx[r], lo, hi); fx[r] = pk2(lo - vx, hi);
      unpk2(fy[r], lo, hi); fy[r] = pk2(lo - vy, hi);
      unpk2(fz[r], lo, hi); fz[r] = pk2(lo - vz, hi);
    };

    if (active) {
      float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
      for (int r = 0; r < S; ++r) {
        ax[r] = cfg[3 * (a0 + r) + 0]; ay[r] = cfg[3 * (a0 + r) + 1]; az[r] = cfg[3 * (a0 + r) + 2];
        const bool real = (r < S - 1) || (g < G - 1);
        if (real) { sx += ax[r]; sy += ay[r]; sz += az[r]; }
        else { ax[r] = kGhost; ay[r] = kGhost; az[r] = kGhost; }   // (the load read the next row / the pad: any value)
        fx[r] = pk2(0.f, 0.f); fy[r] = pk2(0.f, 0.f); fz[r] = pk2(0.f, 0.f);
      }
      s_part[g * C::CPB + c] = make_float4(sx, sy, sz, 0.f);

      // ---- own triangle: streamed atom u against own atoms r < u, two at a time (operands straight from registers)
      lj_block_step<S, 0, 0, 0, 1, FORCE>(ax, ay, az, fx, fy, fz, pk2(ax[1], ax[2]), pk2(ay[1], ay[2]), pk2(az[1], az[2]), e6, en3, rx, ry, rz);
      if (FORCE) { PITA_UNPACK_R(); own_react(1, xl, yl, zl); own_react(2, xh, yh, zh); }
      lj_block_step<S, 0, 2, 0, 3, FORCE>(ax, ay, az, fx, fy, fz, pk2(ax[3], ax[4]), pk2(ay[3], ay[4]), pk2(az[3], az[4]), e6, en3, rx, ry, rz);
      if (FORCE) { PITA_UNPACK_R(); own_react(3, xl, yl, zl); own_react(4, xh, yh, zh); }
      lj_block_step<S, 0, 4, 0, 5, FORCE>(ax, ay, az, fx, fy, fz, pk2(ax[5], ax[6]), pk2(ay[5], ay[6]), pk2(az[5], az[6]), e6, en3, rx, ry, rz);
      if (FORCE) { PITA_UNPACK_R(); own_react(5, xl, yl, zl); own_react(6, xh, yh, zh); }

      // ---- rectangles (g, g+1..3): streamed atoms 0..19 (atom 20 rides with the half rectangle); the next pair is loaded
      //      before this pair's reaction stores (loads cannot be hoisted above stores that may alias them)
      load2(rect_atom(0), rect_atom(1), nx, ny, nz);
#pragma unroll
      for (int q = 0; q < 20; q += 2) {
        bx = nx; by = ny; bz = nz;
        if (q + 2 < 20) load2(rect_atom(q + 2), rect_atom(q + 3), nx, ny, nz);
        lj_block_step<S, 0, S - 1, 0, S - 1, FORCE>(ax, ay, az, fx, fy, fz, bx, by, bz, e6, en3, rx, ry, rz);
        if (FORCE) {
          PITA_UNPACK_R();
          put(q / S, rect_atom(q), xl, yl, zl);
          put((q + 1) / S, rect_atom(q + 1), xh, yh, zh);
        }
      }
      // ---- the last rectangle atom and this thread's share of the (g, g+4) rectangle
      if (g < G / 2) {   // own atoms 0..6 against atoms 0..3 of group g+4
        load2(rect_atom(20), half_atom(0), bx, by, bz);
        load2(half_atom(1), half_atom(2), nx, ny, nz);
        lj_block_step<S, 0, S - 1, 0, S - 1, FORCE>(ax, ay, az, fx, fy, fz, bx, by, bz, e6, en3, rx, ry, rz);
        load2(half_atom(3), half_atom(3), bx, by, bz);
        if (FORCE) { PITA_UNPACK_R(); put(2, rect_atom(20), xl, yl, zl); put(3, half_atom(0), xh, yh, zh); }
        lj_block_step<S, 0, S - 1, 0, S - 1, FORCE>(ax, ay, az, fx, fy, fz, nx, ny, nz, e6, en3, rx, ry, rz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(1), xl, yl, zl); put(3, half_atom(2), xh, yh, zh); }
        lj_block_step<S, 0, S - 1, 1, 0, FORCE>(ax, ay, az, fx, fy, fz, bx, by, bz, e6, en3, rx, ry, rz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(3), xl, yl, zl); }
      } else {           // own atoms 4..6 against all seven atoms of group g-4
        load2(half_atom(0), half_atom(1), bx, by, bz);
        load2(half_atom(2), half_atom(3), nx, ny, nz);
        lj_block_step<S, 4, 6, 4, 6, FORCE>(ax, ay, az, fx, fy, fz, bx, by, bz, e6, en3, rx, ry, rz);
        load2(half_atom(4), half_atom(5), bx, by, bz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(0), xl, yl, zl); put(3, half_atom(1), xh, yh, zh); }
        lj_block_step<S, 4, 6, 4, 6, FORCE>(ax, ay, az, fx, fy, fz, nx, ny, nz, e6, en3, rx, ry, rz);
        load2(half_atom(6), rect_atom(20), nx, ny, nz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(2), xl, yl, zl); put(3, half_atom(3), xh, yh, zh); }
        lj_block_step<S, 4, 6, 4, 6, FORCE>(ax, ay, az, fx, fy, fz, bx, by, bz, e6, en3, rx, ry, rz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(4), xl, yl, zl); put(3, half_atom(5), xh, yh, zh); }
        lj_block_step<S, 4, 6, 0, S - 1, FORCE>(ax, ay, az, fx, fy, fz, nx, ny, nz, e6, en3, rx, ry, rz);
        if (FORCE) { PITA_UNPACK_R(); put(3, half_atom(6), xl, yl, zl); put(2, rect_atom(20), xh, yh, zh); }
      }
    }
#undef PITA_UNPACK_R
    __syncthreads();

    // Every streamed position was read before this barrier, so the position rows of this tile are free: they become the
    // coalescing stage of the force rows.
    if (active) {
      float a_lo, a_hi, b_lo, b_hi;
      unpk2(e6, a_lo, a_hi);
      unpk2(en3, b_lo, b_hi);
      const float e_pairs = (a_lo + a_hi) + 2.0f * (b_lo + b_hi);  // this thread's unordered pairs: sum r^-12 - 2 r^-6
      float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) {
        const float4 q = s_part[gg * C::CPB + c];
        cx += q.x; cy += q.y; cz += q.z;
      }
      cx *= (1.0f / NA); cy *= (1.0f / NA); cz *= (1.0f / NA);
      const float k24 = 24.0f * energy_factor * inv_T, ko = osc * inv_T;
      float e_osc = 0.f;
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const bool real = (r < S - 1) || (g < G - 1);
        const float ux = ax[r] - cx, uy = ay[r] - cy, uz = az[r] - cz;
        if (real) e_osc = fmaf(ux, ux, fmaf(uy, uy, fmaf(uz, uz, e_osc)));
        if (FORCE && real) {
          float lo, hi, px, py, pz;
          unpk2(fx[r], lo, hi); px = -(lo + hi);
          unpk2(fy[r], lo, hi); py = -(lo + hi);
          unpk2(fz[r], lo, hi); pz = -(lo + hi);
          const float *sp = s_react + ((a0 + r) * 3) * C::CPB + c;
#pragma unroll
          for (int sl = 0; sl < C::kSlots; ++sl) {
            // slots 0..2: written by threads g-1, g-2, g-3 for every atom; slot 3: by thread g+4 for every atom of groups
            // 0-3, by thread g-4 for atoms 0..3 of groups 4-7
            if (sl < 3 || r < 4 || g < G / 2) {
              const float *q = sp + sl * NP * 3 * C::CPB;
              px += q[0]; py += q[C::CPB]; pz += q[2 * C::CPB];
            }
          }
          float *o = s_pos + c * P + 3 * (a0 + r);
          o[0] = fmaf(k24, px, -ko * ux); o[1] = fmaf(k24, py, -ko * uy); o[2] = fmaf(k24, pz, -ko * uz);
        }
      }
      s_en[g * C::CPB + c] = energy_factor * 2.0f * e_pairs + osc * 0.5f * e_osc;
    }
    __syncthreads();
    if (g == 0 && active) {
      float v = 0.f;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) v += s_en[gg * C::CPB + c];
      logp[(int64_t)tile * C::CPB + c] = -v * inv_T;
    }
    if (FORCE) {
      float *__restrict__ dst = force + (int64_t)tile * C::CPB * D;
      const int nvec = nflt >> 2;
      for (int v = tid; v < nvec; v += C::kThreads)
        reinterpret_cast<float4 *>(dst)[v] = reinterpret_cast<const float4 *>(s_pos)[v];
      for (int f = (nvec << 2) + tid; f < nflt; f += C::kThreads) dst[f] = s_pos[f];
    }
  }
}

#pragma nv_diag_default 549

template <int NA>
static int launch_lj_blocked(const float *x, int64_t B, float T, float ef, float osc, float *logp, float *force,
                             cudaStream_t st) {
  using C = LJBlockCfg<NA>;
  const int64_t blocks = (B + C::CPB - 1) / C::CPB;
  PITA_REQUIRE(blocks < (1ll << 31), PITA_EINVAL, "lj: batch too large");
  cudaFuncSetAttribute(lj_blocked_kernel<NA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  cudaFuncSetAttribute(lj_blocked_kernel<NA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (force != nullptr)
    lj_blocked_kernel<NA, true><<<(unsigned)blocks, C::kThreads, C::kSmemBytes, st>>>(x, B, 1.0f / T, ef, osc, logp, force);
  else
    lj_blocked_kernel<NA, false><<<(unsigned)blocks, C::kThreads, C::kSmemBytes, st>>>(x, B, 1.0f / T, ef, osc, logp, force);
  PITA_CHECK_LAUNCH("lj_blocked_kernel");
  return PITA_OK;
}

template <int NA, int G, int CW, int MINB>
static int launch_lj_pairs(const float *x, int64_t B, float T, float ef, float osc, float *logp, float *force,
                           cudaStream_t st) {
  using C = LJPairCfg<NA, G, CW>;
  const int64_t blocks = (B + C::CPB - 1) / C::CPB;
  PITA_REQUIRE(blocks < (1ll << 31), PITA_EINVAL, "lj: batch too large");
  // per launch: the attribute is per device, and one process may drive several
  cudaFuncSetAttribute(lj_pairs_kernel<NA, G, CW, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  cudaFuncSetAttribute(lj_pairs_kernel<NA, G, CW, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (force != nullptr)
    lj_pairs_kernel<NA, G, CW, true, MINB><<<(unsigned)blocks, C::kThreads, C::kSmemBytes, st>>>(x, B, 1.0f / T, ef, osc, logp, force);
  else
    lj_pairs_kernel<NA, G, CW, false, MINB><<<(unsigned)blocks, C::kThreads, C::kSmemBytes, st>>>(x, B, 1.0f / T, ef, osc, logp, force);
  PITA_CHECK_LAUNCH("lj_pairs_kernel");
  return PITA_OK;
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
  // PITA_LJ_KERNEL=ordered selects the scalar ordered-pair kernel (kept for A/B measurements; same results to rounding)
  static const bool ordered = [] { const char *e = getenv("PITA_LJ_KERNEL"); return e && e[0] == 'o'; }();
  // (A three-CTA-per-SM build of the LJ-55 kernel -- 128 registers, 15 warps -- measured 6 % SLOWER than this two-CTA
  //  one: profiles/r1g_bench_lj55_minb3.jsonl; the kernel is bound by its dependent FMA/MUFU chains, not by occupancy.)
  switch (n) {
    case 13:
      if (ordered) return launch_lj<13, 32>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
      return launch_lj_pairs<13, 1, 4, 3>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
    case 55:
      if (ordered) return launch_lj<55, 8>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
      {  // PITA_LJ_CFG selects alternative mappings of the paired kernel (A/B measurements, profiles/r2j_*)
        static const int cfg = [] { const char *e = getenv("PITA_LJ_CFG"); return e ? atoi(e) : 0; }();
        if (cfg == 1) return launch_lj_pairs<55, 11, 1, 3>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
        if (cfg == 2) return launch_lj_pairs<55, 11, 1, 4>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
        if (cfg == 3) return launch_lj_pairs<55, 11, 1, 2>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
        if (cfg == 4) return launch_lj_pairs<55, 5, 2, 1>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
        if (cfg == 8) return launch_lj_pairs<55, 5, 1, 2>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
      }
      return launch_lj_blocked<55>(x, B, temperature, energy_factor, oscillator_scale, logp, force, st);
    default:
      set_error("lj: n_particles=%d unsupported (reference raises NotImplementedError for n not in {13,55}, "
                "lennardjones_energy.py:177-178)", n);
      return PITA_EUNSUP;
  }
}
