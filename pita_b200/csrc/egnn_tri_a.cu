// Phase A of the round-2 score / divergence engine (see egnn_tri.cuh): primal forward of the 3-layer EGNN on the row
// engine (thread = (particle, node) row, every 32x32 product a 3xTF32 tcgen05.mma with the operand row handed over through
// the row's own TMEM lane), extended by
//   layer 0: the forward-mode tangent of every edge (i, k) w.r.t. the coordinates of its sender k   -> omega_ik, M_ik, PA, PB
//   layer 2: the reverse-mode cotangent of the trace through every edge (k, j) of output node k      -> gamma, w, X
// written as per-pair tables into the per-particle workspace, and the part of the trace that needs no middle-layer edge
// ("direct").  Algebra and names: oracle/egnn_bilinear.py::pair_tables / phase_a_tables.
// Reference being differentiated: egnn_temp_conditioned.py:56-93,265-356; score_net.py:13-43; utils.py:30-51.
#include "egnn_tri.cuh"

namespace pita {
namespace tri {

using rg::Geo;
using rg::kNumVec;
using rg::kRows;
using rg::Team;
using namespace rg;  // Vec enum, stage helpers

constexpr int kWS = 9;  // weight-tile slots (hi + lo, 8 KB each)
// TMEM column slots (32 columns each) of one team: operand row (hi, lo), two accumulators, three per-row vector
// accumulators (S_a in layer 0, Gamma_a in layer 2) and one per-row keeper (f1*c01 / f3^0 / P^1 / P^2).
enum Slot { sAh = 0, sAl = 1, sD0 = 2, sD1 = 3, sS0 = 4, sK = 7 };

struct WSrc {
  const float *p[kWS];
  int count;
};

template <int NP, int NTEAM>
struct SmemA {
  static constexpr int PB = kRows / NP;
  static constexpr size_t oW = 0;
  static constexpr size_t oF = oW + (size_t)kWS * 8192;
  static constexpr int fVec = 0;                        // [3][kNumVec][32]
  static constexpr int fEmb = fVec + 3 * kNumVec * 32;  // [3][32]
  static constexpr int fCls = fEmb + 96;                // [6][32]  A0 e0, A0 e1, A0 eb + b1, B0 e0, B0 e1, B0 eb
  static constexpr int fCls3 = fCls + 192;              // [3][32]  W3h0 e0, W3h0 e1, W3h0 eb
  static constexpr int fMisc = fCls3 + 96;              // mbarriers + tmem slot
  static constexpr int fTeam = fMisc + 2 * NTEAM + 8;
  static constexpr int tQ = 0;                          // [128][32] swizzled sender rows: Q^1, later Q^2
  static constexpr int tF3 = tQ + kRows * 32;           // [128][32] swizzled: h^1 (own row, transient), later f3^1 (gathered)
  static constexpr int tX = tF3 + kRows * 32;           // [4][128] float4: y, x^1, x^2, x^3
  static constexpr int tRed = tX + 4 * kRows * 4;       // [128]
  static constexpr int tMean = tRed + kRows;            // [128] float4
  static constexpr int kTeamFloats = tMean + kRows * 4;
  static constexpr size_t kBytes = 1024 + oF + (size_t)(fTeam + NTEAM * kTeamFloats) * 4;
};

__device__ __forceinline__ void load_weights9(float *wsm, const WSrc &src, int tid, int nthreads) {
  __syncthreads();  // every team is done with the MMAs that read the old tiles
  for (int item = tid; item < src.count * 32; item += nthreads) {
    const int m = item >> 5, row = item & 31;
    const float4 *g = reinterpret_cast<const float4 *>(src.p[m] + row * 32);
    float h[32], l[32];
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 q = __ldg(g + k4);
      umma::split_tf32(q.x, h[4 * k4], l[4 * k4]);
      umma::split_tf32(q.y, h[4 * k4 + 1], l[4 * k4 + 1]);
      umma::split_tf32(q.z, h[4 * k4 + 2], l[4 * k4 + 2]);
      umma::split_tf32(q.w, h[4 * k4 + 3], l[4 * k4 + 3]);
    }
    float *hi = wsm + m * 2048;
    umma::store_row_sw128(hi, row, h);
    umma::store_row_sw128(hi + 1024, row, l);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
}

// operand row (hi + lo split, round to nearest) into the row's TMEM lane
__device__ __forceinline__ void put(const Team<true> &T, const float (&v)[32]) {
  float h[32], l[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) umma::split_tf32(v[k], h[k], l[k]);
  T.st(sAh, h);
  T.st(sAl, l);
}
// D[dslot] (+)= row x W[wslot]^T as 3xTF32 (lo*hi, hi*lo, hi*hi; the small products first)
__device__ __forceinline__ void mma3(const Team<true> &T, int dslot, int wslot, bool accumulate) {
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
  const uint32_t d = T.tmem_col + 32u * dslot, ah = T.tmem_col + 32u * sAh, al = T.tmem_col + 32u * sAl;
  const uint32_t wa = T.w_addr + (uint32_t)wslot * 8192u;
  const uint64_t dB = umma::make_desc_sw128_kmajor(wa), dBl = umma::make_desc_sw128_kmajor(wa + 4096u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, al + 8u * k, dB + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dBl + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, 1u);
}

__device__ __forceinline__ float dot32r(const float (&a)[32], const float (&b)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    s0 = fmaf(a[k], b[k], s0); s1 = fmaf(a[k + 1], b[k + 1], s1);
    s2 = fmaf(a[k + 2], b[k + 2], s2); s3 = fmaf(a[k + 3], b[k + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dot32s(const float (&a)[32], const float *svec) {  // svec: constant table in shared memory
  svec = opq(svec);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = lds4c(svec + 4 * k4);
    s0 = fmaf(a[4 * k4], q.x, s0); s1 = fmaf(a[4 * k4 + 1], q.y, s1);
    s2 = fmaf(a[4 * k4 + 2], q.z, s2); s3 = fmaf(a[4 * k4 + 3], q.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// stage 3 variant: returns tanh(u) and leaves  rowc = wc2 * silu'(zc)  in acc (the operand of the Wc1^T product)
__device__ __forceinline__ float stage3_v(float (&acc)[32], const float *vec) {
  vec = opq(vec);
  float u = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 b = lds4c(vec + vBC1 * 32 + 4 * k4);
    const float4 w = lds4c(vec + vWC2 * 32 + 4 * k4);
    const float bb[4] = {b.x, b.y, b.z, b.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      float a, f;
      silu_both(acc[k] + bb[e], a, f);
      u = fmaf(ww[e], a, u);
      acc[k] = ww[e] * f;
    }
  }
  return tanhf(u);
}

// 3x3 helpers ([b][a] row-major)
struct M3 {
  float m[9];
};
__device__ __forceinline__ void store9(float *dst, const M3 &a) {  // dst 4-byte aligned only
#pragma unroll
  for (int q = 0; q < 9; ++q) dst[q] = a.m[q];
}

template <int NP, int NTEAM>
__global__ void __launch_bounds__(NTEAM * 128, 1)
tri_phase_a_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                   const float *__restrict__ beta_in, int64_t b0, int64_t nb, float *__restrict__ score,
                   float *__restrict__ ws, int want_div) {
  extern __shared__ __align__(16) float sm_raw[];
  using S = SmemA<NP, NTEAM>;
  using W = WS<NP>;
  constexpr int PB = S::PB;
  constexpr int R = PB * NP;
  constexpr int L = 3;
  const float rng = kCoordsRange / (float)L;

  // ---- carve shared memory, TMEM, barriers
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, team = tid >> 7, tt = tid & 127, warp = tid >> 5;
  float *wsm = reinterpret_cast<float *>(base + S::oW);
  float *fl = reinterpret_cast<float *>(base + S::oF);
  float *sVec = fl + S::fVec, *sEmb = fl + S::fEmb, *sCls = fl + S::fCls, *sCls3 = fl + S::fCls3;
  uint64_t *mbars = reinterpret_cast<uint64_t *>(fl + S::fMisc);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fl + S::fMisc + 2 * NTEAM);
  float *tm = fl + S::fTeam + team * S::kTeamFloats;
  float *sQ = tm + S::tQ, *sF3 = tm + S::tF3, *sRed = tm + S::tRed;
  float4 *sX = reinterpret_cast<float4 *>(tm + S::tX);
  float4 *sMean = reinterpret_cast<float4 *>(tm + S::tMean);
  const int p = tt / NP, i = tt - p * NP;
  const bool row_ok = tt < R;
  const int pp = row_ok ? p : 0, ic = row_ok ? i : 0;

  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int k = 0; k < NTEAM; ++k) umma::mbar_init(mbars + k, 1);
    umma::fence_mbar_init();
  }
  for (int k = tid; k < 3 * kNumVec * 32; k += NTEAM * 128) {
    const int l = k / (kNumVec * 32), r = k % (kNumVec * 32);
    sVec[k] = __ldg(wpack + pk::kHeader + l * pk::kLayer + pk::c1 + r);
  }
  for (int k = tid; k < 96; k += NTEAM * 128) sEmb[k] = __ldg(wpack + k);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = uniform32(*tmem_slot);
  Team<true> T;
  {
    const uint32_t team_u = uniform32((uint32_t)team);
    T.a_hi = nullptr;
    T.a_addr = 0;
    T.w_addr = uniform32(umma::smem_u32(wsm));
    T.mbar_addr = uniform32(umma::smem_u32(mbars)) + team_u * 8u;
    T.phase = 0;
    T.tmem_col = tmem_base + team_u * (uint32_t)(512 / NTEAM);
    T.tmem = T.tmem_col + (((uint32_t)((warp & 3) * 32)) << 16);
    T.bar_id = 1 + team;
    T.tt = tt;
    T.issuer = uniform32((uint32_t)(warp & 3)) == 0u;
  }
  const float *W0 = wpack + pk::kHeader, *W1 = W0 + pk::kLayer, *W2l = W1 + pk::kLayer;
  const float *vec0 = sVec, *vec1 = sVec + kNumVec * 32, *vec2 = sVec + 2 * kNumVec * 32;

  // layer-0 class tables (h^0 takes three values: node features (t,t), (t,beta), (beta,beta); egnn_temp_conditioned.py:63-78)
  for (int item = tid; item < 9 * 32; item += NTEAM * 128) {
    const int v = item >> 5, ch = item & 31;
    const float *Mx = W0 + (v < 3 ? pk::A_b : (v < 6 ? pk::B_b : pk::W3h_b)) + ch * 32;
    const float *e = sEmb + (v % 3) * 32;
    float acc = (v == 2) ? __ldg(W0 + pk::b1 + ch) : 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(__ldg(Mx + k), e[k], acc);
    (v < 6 ? sCls + v * 32 : sCls3 + (v - 6) * 32)[ch] = acc;
  }
  __syncthreads();

  WSrc set1, set2, set3;
  set1.p[0] = W0 + pk::W2_b; set1.p[1] = W0 + pk::Wc1_b; set1.p[2] = W0 + pk::Wc1_f; set1.p[3] = W0 + pk::W3a_b;
  set1.p[4] = W0 + pk::W4_b; set1.p[5] = W1 + pk::A_b; set1.p[6] = W1 + pk::B_b;
  set1.count = 7;
  set2.p[0] = W1 + pk::W2_b; set2.p[1] = W1 + pk::Wc1_b; set2.p[2] = W1 + pk::W3a_b; set2.p[3] = W1 + pk::W3h_b;
  set2.p[4] = W1 + pk::W4_b; set2.p[5] = W2l + pk::A_b; set2.p[6] = W2l + pk::B_b;
  set2.count = 7;
  set3.p[0] = W2l + pk::W2_b; set3.p[1] = W2l + pk::Wc1_b; set3.p[2] = W2l + pk::Wc1_f; set3.p[3] = W2l + pk::W2_f;
  set3.p[4] = W2l + pk::A_f; set3.p[5] = W2l + pk::B_f; set3.p[6] = W1 + pk::W4_f; set3.p[7] = W1 + pk::W3a_f;
  set3.p[8] = W1 + pk::W3h_f;
  set3.count = 9;

  auto sender = [&](int u) { int j = ic + 1 + u; return j >= NP ? j - NP : j; };
  auto tsync = [&]() {
    umma::fence_before_thread_sync();
    T.sync();
    umma::fence_after_thread_sync();
  };

  const int64_t ntile = (nb + (int64_t)NTEAM * PB - 1) / ((int64_t)NTEAM * PB);
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t lp0 = (tile * NTEAM + team) * PB;
    const bool team_active = lp0 < nb;
    const int64_t lp = lp0 + p;
    const bool ok = row_ok && lp < nb;
    const int64_t lpc = ok ? lp : (nb - 1);
    const int64_t part = b0 + lpc;
    float *wsp = ws + (size_t)lpc * (size_t)W::kFloats;  // this row's particle
    const float h = __ldg(ht + part), beta = __ldg(beta_in + part);
    const float c_in = rsqrtf(1.0f + h), c_s = 1.0f / (1.0f + h), c_out = sqrtf(h) * c_in, c_noise = 0.125f * logf(h);
    const float *src = x + part * 3 * NP + 3 * ic;
    const float xr0 = __ldg(src), xr1 = __ldg(src + 1), xr2 = __ldg(src + 2);
    const float4 yi = make_float4(c_in * xr0, c_in * xr1, c_in * xr2, 0.f);
    sX[tt] = yi;
    float f0i, f1i;
    node_feats<NP>(ic, c_noise, beta, f0i, f1i);
    float row[32], va[32], vb[32], agg[32];
    M3 Dx1;  // own-direction coordinate tangent of x^1 (3x3, [b][a])
#pragma unroll
    for (int q = 0; q < 9; ++q) Dx1.m[q] = (q % 4 == 0) ? 1.0f : 0.0f;

    // =========================================================================================== layer 0, sweep 1
    load_weights9(wsm, set1, tid, NTEAM * 128);
    if (team_active) {
      if (want_div && ok) *reinterpret_cast<float4 *>(wsp + W::oY + 4 * i) = yi;
      tsync();  // sX[0] published
      float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) { agg[k] = 0.f; row[k] = 0.f; }
      if (want_div) { T.st(sS0, row); T.st(sS0 + 1, row); T.st(sS0 + 2, row); }
#pragma unroll 1
      for (int u = 0; u < NP - 1; ++u) {
        const int j = sender(u), rj = pp * NP + j;
        const float4 yj = sX[rj];
        const Geo g = rg::edge_geo4(yi, yj, yi, yj);
        float f0j, f1j;
        node_feats<NP>(j, c_noise, beta, f0j, f1j);
        z1_layer0(row, sCls, f0i, f1i, f0j, f1j);
        stage1<true, false>(row, va, nullptr, 0, vec0, g.r2, g.ea);  // row = silu(z1), va = f1
        put(T, row);
        if (want_div) {  // keep f1 * (c1 + d1), the operand of the wvec product
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 cc = lds4(vec0 + vC1 * 32 + 4 * k4), dd = lds4(vec0 + vD1 * 32 + 4 * k4);
            va[4 * k4] *= (cc.x + dd.x); va[4 * k4 + 1] *= (cc.y + dd.y);
            va[4 * k4 + 2] *= (cc.z + dd.z); va[4 * k4 + 3] *= (cc.w + dd.w);
          }
          T.st(sK, va);
        }
        T.round_trip_ts([&] { mma3(T, sD0, 0, false); });
        T.ld(sD0, row);
        const float att = stage2<true>(row, va, vb, vec0);  // row = m*att, va = m, vb = f2
#pragma unroll
        for (int k = 0; k < 32; ++k) agg[k] += row[k];
        put(T, row);
        T.round_trip_ts([&] { mma3(T, sD1, 1, false); });
        float th;
        if (!want_div) {
          T.ld(sD1, row);
          th = stage3<false>(row, row, vec0);
        } else {
          // wvec = T (c1 + d1)
          T.ld(sK, row);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 0, false); });
          T.ld(sD0, row);
          tangent_mid(row, va, vb, att, vec0);  // row = wvec_ij  (m, f2 dead from here)
          // coordinate branch: th, v0 = Wc1^T (wc2 * silu'(zc)), sigma = <v0, wvec>
          T.ld(sD1, va);
          th = stage3_v(va, vec0);
          put(T, va);
          T.round_trip_ts([&] { mma3(T, sD0, 2, false); });
          T.ld(sD0, va);
          const float sigma = dot32r(va, row);
          // S_a += cf_a(j)|_i wvec,  cf(j)|_i = 2 (y_i - y_j)
          {
            const float c3[3] = {2.0f * g.d0, 2.0f * g.d1, 2.0f * g.d2};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              T.ld(sS0 + a, va);
#pragma unroll
              for (int k = 0; k < 32; ++k) va[k] = fmaf(c3[a], row[k], va[k]);
              T.st(sS0 + a, va);
            }
          }
          // yv = W3a wvec  (omega_ij = W4 (f3^0_i * yv) once f3^0 is known: sweep 2)
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 3, false); });
          T.ld(sD0, row);
          if (ok) rg::store_vec_global(wsp + W::oOM + ((size_t)i * NP + j) * 32, row);
          // M_ij (tangent node j) and this row's own Dx1
          const float phi = rng * th, cphi = rng * (1.0f - th * th), k2 = g.inv * g.inv / g.nrm, cs = cphi * sigma;
          const float d3[3] = {g.d0, g.d1, g.d2};
          M3 Mij;
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float Nba = ((a == b) ? g.inv : 0.0f) - k2 * d3[b] * d3[a];
              const float t2 = cs * d3[b] * g.inv * 2.0f * d3[a];  // cphi sigma dhat_b (2 d_a)
              Mij.m[b * 3 + a] = -phi * Nba - t2;                   // cf_a(i)|_j = -2 d_a
              Dx1.m[b * 3 + a] += phi * Nba + t2;
            }
          if (ok) {
            store9(wsp + W::oTR + ((size_t)i * NP + j) * kTR + trM, Mij);
            store9(wsp + W::oTS + ((size_t)j * NP + i) * kTS + tsM, Mij);
          }
        }
        const float f = g.inv * th * rng;
        dx0 = fmaf(g.d0, f, dx0); dx1 = fmaf(g.d1, f, dx1); dx2 = fmaf(g.d2, f, dx2);
      }
      const float4 x1 = make_float4(yi.x + dx0, yi.y + dx1, yi.z + dx2, 0.f);
      sX[kRows + tt] = x1;
      if (want_div && ok) *reinterpret_cast<float4 *>(wsp + W::oX1 + 4 * i) = x1;
      // ---- node update: z3 = W3h h0 + W3a agg + b3,  h1 = h0 + W4 silu(z3) + b4
      put(T, agg);
      T.round_trip_ts([&] { mma3(T, sD0, 3, false); });
      T.ld(sD0, row);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 a0 = lds4(sCls3 + 4 * k4), a1 = lds4(sCls3 + 32 + 4 * k4), ab = lds4(sCls3 + 64 + 4 * k4);
        const float4 b3 = lds4(vec0 + vB3 * 32 + 4 * k4);
        row[4 * k4] += fmaf(f0i, a0.x, fmaf(f1i, a1.x, ab.x)) + b3.x;
        row[4 * k4 + 1] += fmaf(f0i, a0.y, fmaf(f1i, a1.y, ab.y)) + b3.y;
        row[4 * k4 + 2] += fmaf(f0i, a0.z, fmaf(f1i, a1.z, ab.z)) + b3.z;
        row[4 * k4 + 3] += fmaf(f0i, a0.w, fmaf(f1i, a1.w, ab.w)) + b3.w;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) { float a; silu_both(row[k], a, va[k]); row[k] = a; }  // va = f3^0
      T.st(sK, va);
      put(T, row);
      T.round_trip_ts([&] { mma3(T, sD0, 4, false); });
      T.ld(sD0, row);
      embed<NP>(vb, sEmb, ic, c_noise, beta);
      add_vec(row, vec0 + vB4 * 32);
#pragma unroll
      for (int k = 0; k < 32; ++k) vb[k] += row[k];  // vb = h^1
      rg::qrow_store(sF3, tt, vb);                   // own row, read back below
      if (want_div) {
        // ---- Omega_i[a] = W4 (f3^0 * W3a S_a);  A1 Omega, B1 Omega
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
          T.ld(sS0 + a, row);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 3, false); });
          T.ld(sD0, row);
          T.ld(sK, va);
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] *= va[k];
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 4, false); });
          T.ld(sD0, row);
          if (ok) rg::store_vec_global(wsp + W::oOmg + ((size_t)i * 3 + a) * 32, row);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 5, false); mma3(T, sD1, 6, false); });
          T.ld(sD0, row);
          if (ok) rg::store_vec_global(wsp + W::oAOm + ((size_t)i * 3 + a) * 32, row);
          T.ld(sD1, row);
          if (ok) rg::store_vec_global(wsp + W::oBOm + ((size_t)i * 3 + a) * 32, row);
        }
        // =========================================================================================== layer 0, sweep 2
        T.ld(sK, va);  // f3^0
        // the yv row of the next edge is fetched while this edge's two products are in flight (vb is free here: h^1 is
        // re-read from shared memory after the sweep)
        if (ok) rg::load_vec_global(wsp + W::oOM + ((size_t)i * NP + sender(0)) * 32, vb);
#pragma unroll 1
        for (int u = 0; u < NP - 1; ++u) {
          const int j = sender(u);
          float *om = wsp + W::oOM + ((size_t)i * NP + j) * 32;
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] = ok ? vb[k] * va[k] : 0.f;
          if (ok && u + 1 < NP - 1) rg::load_vec_global(wsp + W::oOM + ((size_t)i * NP + sender(u + 1)) * 32, vb);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 4, false); });
          T.ld(sD0, row);  // omega_ij
          if (ok) rg::store_vec_global(om, row);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 5, false); mma3(T, sD1, 6, false); });
          T.ld(sD0, row);
          if (ok) rg::store_vec_global(wsp + W::oTR + ((size_t)i * NP + j) * kTR + trPA, row);
          T.ld(sD1, row);
          if (ok) rg::store_vec_global(wsp + W::oTS + ((size_t)j * NP + i) * kTS + tsPB, row);
        }
      }
      // ---- layer-1 node products  P^1 = A1 h1 + b1 (TMEM sK), Q^1 = B1 h1 (shared rows)
      rg::qrow_load(sF3, tt, vb);
      put(T, vb);
      T.round_trip_ts([&] { mma3(T, sD0, 5, false); mma3(T, sD1, 6, false); });
      T.ld(sD0, row);
      add_vec(row, vec1 + vB1 * 32);
      T.st(sK, row);
      if (want_div && ok) rg::store_vec_global(wsp + W::oP1 + (size_t)i * 32, row);
      T.ld(sD1, row);
      rg::qrow_store(sQ, tt, row);
      if (want_div && ok) rg::store_vec_global(wsp + W::oQ1 + (size_t)i * 32, row);
    }

    // =========================================================================================== layer 1 (primal only)
    load_weights9(wsm, set2, tid, NTEAM * 128);
    if (team_active) {
      tsync();  // Q^1 rows and x^1 visible
      const float4 xi = sX[kRows + tt];
      float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) agg[k] = 0.f;
#pragma unroll 1
      for (int u = 0; u < NP - 1; ++u) {
        const int rj = pp * NP + sender(u);
        const Geo g = rg::edge_geo4(xi, sX[kRows + rj], yi, sX[rj]);
        T.ld(sK, row);
        stage1<false>(row, va, sQ, rj, vec1, g.r2, g.ea);
        put(T, row);
        T.round_trip_ts([&] { mma3(T, sD0, 0, false); });
        T.ld(sD0, row);
        stage2<false>(row, va, va, vec1);
#pragma unroll
        for (int k = 0; k < 32; ++k) agg[k] += row[k];
        put(T, row);
        T.round_trip_ts([&] { mma3(T, sD1, 1, false); });
        T.ld(sD1, row);
        const float th = stage3<false>(row, va, vec1);
        const float f = g.inv * th * rng;
        dx0 = fmaf(g.d0, f, dx0); dx1 = fmaf(g.d1, f, dx1); dx2 = fmaf(g.d2, f, dx2);
      }
      sX[2 * kRows + tt] = make_float4(xi.x + dx0, xi.y + dx1, xi.z + dx2, 0.f);
      // node update
      put(T, agg);
      T.round_trip_ts([&] { mma3(T, sD0, 2, false); });
      put(T, vb);  // h^1
      T.round_trip_ts([&] { mma3(T, sD0, 3, true); });
      T.ld(sD0, row);
      add_vec(row, vec1 + vB3 * 32);
#pragma unroll
      for (int k = 0; k < 32; ++k) { float a; silu_both(row[k], a, va[k]); row[k] = a; }  // va = f3^1
      tsync();                         // every row is done gathering nothing from sF3 yet; own h^1 row no longer needed
      rg::qrow_store(sF3, tt, va);     // f3^1, gathered by the layer-2 pass
      put(T, row);
      T.round_trip_ts([&] { mma3(T, sD0, 4, false); });
      T.ld(sD0, row);
      add_vec(row, vec1 + vB4 * 32);
#pragma unroll
      for (int k = 0; k < 32; ++k) vb[k] += row[k];  // h^2
      // layer-2 node products
      put(T, vb);
      T.round_trip_ts([&] { mma3(T, sD0, 5, false); mma3(T, sD1, 6, false); });
      T.ld(sD0, row);
      add_vec(row, vec2 + vB1 * 32);
      T.st(sK, row);  // P^2
      T.ld(sD1, row);
      tsync();  // all rows done with the Q^1 gathers
      rg::qrow_store(sQ, tt, row);
    }

    // =========================================================================================== layer 2 (+ cotangents)
    load_weights9(wsm, set3, tid, NTEAM * 128);
    float direct = 0.f;
    if (team_active) {
      tsync();  // Q^2, f3^1 rows and x^2 visible
      const float4 xi = sX[2 * kRows + tt];
      float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
      M3 Gx;
#pragma unroll
      for (int q = 0; q < 9; ++q) Gx.m[q] = (q % 4 == 0) ? 1.0f : 0.0f;
      if (want_div) {
#pragma unroll
        for (int k = 0; k < 32; ++k) row[k] = 0.f;
        T.st(sS0, row); T.st(sS0 + 1, row); T.st(sS0 + 2, row);
      }
#pragma unroll 1
      for (int u = 0; u < NP - 1; ++u) {
        const int j = sender(u), rj = pp * NP + j;
        const float4 yj = sX[rj];
        const Geo g = rg::edge_geo4(xi, sX[2 * kRows + rj], yi, yj);
        if (want_div && ok) {  // lines read at the end of this edge (written by row j in layer 0): start them moving now
          const float *trj = wsp + W::oTR + ((size_t)j * NP + i) * kTR + trM;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(trj));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(trj + 8));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(wsp + W::oOM + ((size_t)j * NP + i) * 32));
        }
        T.ld(sK, row);
        stage1<true>(row, agg, sQ, rj, vec2, g.r2, g.ea);  // agg = f1 (the aggregate is dead code in the last layer)
        put(T, row);
        T.round_trip_ts([&] { mma3(T, sD0, 0, false); });
        T.ld(sD0, row);
        const float att = stage2<true>(row, va, vb, vec2);  // va = m, vb = f2
        put(T, row);
        T.round_trip_ts([&] { mma3(T, sD1, 1, false); });
        T.ld(sD1, row);
        float th;
        if (!want_div) {
          th = stage3<false>(row, row, vec2);
        } else {
          th = stage3_v(row, vec2);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 2, false); });
          T.ld(sD0, row);  // v
          // vt = T^T v = f1 * W2^T ( f2 * (att v + att (1 - att) <m, v> wa) )
          const float mv = dot32r(va, row) * att * (1.0f - att);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 w4 = lds4(vec2 + vWA * 32 + 4 * k4);
            row[4 * k4] = vb[4 * k4] * fmaf(att, row[4 * k4], mv * w4.x);
            row[4 * k4 + 1] = vb[4 * k4 + 1] * fmaf(att, row[4 * k4 + 1], mv * w4.y);
            row[4 * k4 + 2] = vb[4 * k4 + 2] * fmaf(att, row[4 * k4 + 2], mv * w4.z);
            row[4 * k4 + 3] = vb[4 * k4 + 3] * fmaf(att, row[4 * k4 + 3], mv * w4.w);
          }
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 3, false); });
          T.ld(sD0, row);
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] *= agg[k];  // vt
          const float rho = dot32s(row, vec2 + vC1 * 32), delta = dot32s(row, vec2 + vD1 * 32);
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 4, false); mma3(T, sD1, 5, false); });  // alpha = A2^T vt, beta = B2^T vt
          const float phi = rng * th, cphi = rng * (1.0f - th * th), k2 = g.inv * g.inv / g.nrm;
          const float d3[3] = {g.d0, g.d1, g.d2};
          const float wv[3] = {cphi * g.d0 * g.inv, cphi * g.d1 * g.inv, cphi * g.d2 * g.inv};
          const float cfk[3] = {2.0f * (yi.x - yj.x), 2.0f * (yi.y - yj.y), 2.0f * (yi.z - yj.z)};  // cf(j)|_k
          const float alpha_j = wv[0] * cfk[0] + wv[1] * cfk[1] + wv[2] * cfk[2];
          T.ld(sD0, row);  // alpha_kj
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            T.ld(sS0 + a, va);
#pragma unroll
            for (int k = 0; k < 32; ++k) va[k] = fmaf(wv[a], row[k], va[k]);
            T.st(sS0 + a, va);
          }
          T.ld(sD1, vb);  // beta_kj
          put(T, vb);
          T.round_trip_ts([&] { mma3(T, sD0, 6, false); });  // W4_1^T beta
          T.ld(sD0, row);
          rg::qrow_load(sF3, rj, va);  // f3^1 of the sender
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] *= va[k];
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 7, false); mma3(T, sD1, 8, false); });  // gamma = W3a^T pb, W3h^T pb
          T.ld(sD1, row);
#pragma unroll
          for (int k = 0; k < 32; ++k) vb[k] += row[k];  // beta'
          T.ld(sD0, row);                                  // gamma_kj
          M3 X;
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float Nba = ((a == b) ? g.inv : 0.0f) - k2 * d3[b] * d3[a];
              X.m[b * 3 + a] = phi * Nba + 2.0f * rho * d3[b] * wv[a];
              Gx.m[b * 3 + a] += X.m[b * 3 + a];
            }
          direct = fmaf(delta, alpha_j, direct);
          if (ok) {
            // pair term  -<X_kj, M_jk> + alpha_j <beta'_kj, omega_jk>   (M_jk, omega_jk: written by row j in layer 0)
            float *trj = wsp + W::oTR + ((size_t)j * NP + i) * kTR;  // TR[j][k = i]
            float xm = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) xm = fmaf(X.m[q], trj[trM + q], xm);
            rg::load_vec_global(wsp + W::oOM + ((size_t)j * NP + i) * 32, va);
            direct += alpha_j * dot32r(vb, va) - xm;
            rg::store_vec_global(trj + trGam, row);
            trj[trW] = wv[0]; trj[trW + 1] = wv[1]; trj[trW + 2] = wv[2];
            trj[trAlpha] = alpha_j;
#pragma unroll
            for (int q = 0; q < 9; ++q) trj[trGX + q] = -X.m[q];
          }
        }
        const float f = g.inv * th * rng;
        dx0 = fmaf(g.d0, f, dx0); dx1 = fmaf(g.d1, f, dx1); dx2 = fmaf(g.d2, f, dx2);
      }
      sX[3 * kRows + tt] = make_float4(xi.x + dx0, xi.y + dx1, xi.z + dx2, 0.f);
      if (want_div) {
        // ---- output node's own cotangents: Gamma'_k[a], GAgg_k[a];  direct += <Gamma'_k[a], Omega_k[a]> + <Gx, Dx1>
        rg::qrow_load(sF3, tt, agg);  // own f3^1
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
          T.ld(sS0 + a, vb);
          put(T, vb);
          T.round_trip_ts([&] { mma3(T, sD0, 6, false); });
          T.ld(sD0, row);
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] *= agg[k];
          put(T, row);
          T.round_trip_ts([&] { mma3(T, sD0, 7, false); mma3(T, sD1, 8, false); });
          T.ld(sD1, row);
#pragma unroll
          for (int k = 0; k < 32; ++k) vb[k] += row[k];  // Gamma'
          T.ld(sD0, row);
          if (ok) {
            rg::store_vec_global(wsp + W::oGAgg + ((size_t)i * 3 + a) * 32, row);
            rg::load_vec_global(wsp + W::oOmg + ((size_t)i * 3 + a) * 32, va);
            direct += dot32r(vb, va);
          }
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) direct = fmaf(Gx.m[q], Dx1.m[q], direct);
        if (ok) {
          // diagonal table entries: the full-rank tangent / cotangent of node i itself
          float *trd = wsp + W::oTR + ((size_t)i * NP + i) * kTR, *tsd = wsp + W::oTS + ((size_t)i * NP + i) * kTS;
#pragma unroll
          for (int k4 = 0; k4 < 16; ++k4) *reinterpret_cast<float4 *>(trd + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
          trd[trW] = 0.f; trd[trW + 1] = 0.f; trd[trW + 2] = 0.f; trd[trAlpha] = 0.f;
          store9(trd + trGX, Gx);
          store9(trd + trM, Dx1);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) *reinterpret_cast<float4 *>(tsd + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
          store9(tsd + tsM, Dx1);
          wsp[W::oDirect + i] = direct;
        }
      }
      tsync();  // x^3 published
      // ---- score = ((c_s - 1) x + c_out * remove_mean(x^3 - y)) / h        (score_net.py:21-43)
      const float4 xl = sX[3 * kRows + tt];
      const float4 v = make_float4(xl.x - yi.x, xl.y - yi.y, xl.z - yi.z, 0.f);
      sMean[tt] = v;
      T.sync();
      float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll 1
      for (int k = 0; k < NP; ++k) {
        const float4 q = sMean[pp * NP + k];
        m0 += q.x; m1 += q.y; m2 += q.z;
      }
      m0 /= NP; m1 /= NP; m2 /= NP;
      if (ok) {
        float *dst = score + part * 3 * NP + 3 * i;
        dst[0] = ((c_s * xr0 + c_out * (v.x - m0)) - xr0) / h;
        dst[1] = ((c_s * xr1 + c_out * (v.y - m1)) - xr1) / h;
        dst[2] = ((c_s * xr2 + c_out * (v.z - m2)) - xr2) / h;
      }
      T.sync();
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem_base);
}

template <int NP, int NTEAM>
static int launch_a(const float *w, const float *ht, const float *x, const float *beta, int64_t b0, int64_t nb, float *score,
                    float *ws, int want_div, cudaStream_t s) {
  using S = SmemA<NP, NTEAM>;
  static_assert(S::kBytes <= 227 * 1024, "phase A shared memory plan exceeds 227 KB");
  auto k = tri_phase_a_kernel<NP, NTEAM>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
  if (e != cudaSuccess) { set_error("tri_phase_a_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  const int64_t ntile = (nb + (int64_t)NTEAM * S::PB - 1) / ((int64_t)NTEAM * S::PB);
  const unsigned grid = (unsigned)(ntile < kNumSMs ? ntile : kNumSMs);
  k<<<grid, NTEAM * 128, S::kBytes, s>>>(w, ht, x, beta, b0, nb, score, ws, want_div);
  PITA_CHECK_LAUNCH("tri_phase_a_kernel");
  return PITA_OK;
}

int launch_tri_phase_a(int n, const float *w, const float *ht, const float *x, const float *beta, int64_t b0, int64_t nb,
                       float *score, float *ws, int want_div, cudaStream_t s) {
  if (n == 13) return launch_a<13, 2>(w, ht, x, beta, b0, nb, score, ws, want_div, s);
  return launch_a<55, 2>(w, ht, x, beta, b0, nb, score, ws, want_div, s);
}

int tri_particles_per_cta_a(int n) { return 2 * (kRows / n); }

int64_t workspace_floats_per_particle(int n) { return n == 13 ? WS<13>::kFloats : WS<55>::kFloats; }

template <int NP>
static int64_t layout(int64_t *out, int max_out) {
  using W = WS<NP>;
  const int64_t v[] = {W::kFloats, W::oTS, W::oTR, W::oOM, W::oY, W::oX1, W::oP1, W::oQ1, W::oOmg, W::oAOm, W::oBOm, W::oGAgg,
                       W::oDirect, W::oPartB, kTS, kTR};
  const int cnt = (int)(sizeof(v) / sizeof(v[0]));
  for (int k = 0; k < cnt && k < max_out; ++k) out[k] = v[k];
  return cnt;
}
int64_t workspace_layout(int n, int64_t *out, int max_out) { return n == 13 ? layout<13>(out, max_out) : layout<55>(out, max_out); }

}  // namespace tri
}  // namespace pita
