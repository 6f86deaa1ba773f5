// EGNN denoiser kernels (fp32 SIMT): primal forward, hand-derived reverse pass (energy net: E, grad_x E,
// dE/dh) and hand-derived forward-mode tangent pass (score net: score and exact divergence).
// Reference algebra: egnn_temp_conditioned.py:56-93,172-194,265-356; energy_net.py:14-62; score_net.py:13-43;
// utils.py:43-51.  The derivative algebra is stated autograd-free in oracle/egnn_analytic.py and checked
// there against the reference's autograd path.
//
// Mapping: one CTA per particle (one n-atom configuration), one warp per receiver node (round-robin),
// lane = hidden channel (H = 32).  A 32x32 linear layer is `out[lane] = sum_k W[lane][k] * in[k]` with the
// weight row held in 32 registers and the input vector broadcast from shared memory as float4; every
// per-edge activation stays on-chip, nothing but x, h(t), beta is read from HBM and only the [3n] outputs
// and per-particle scalars are written.
#pragma once
#include "common.cuh"

namespace pita {

constexpr int H = 32;
constexpr float kCoordsRange = 15.0f;  // EGNN(coords_range=15), egnn_temp_conditioned.py:133,143
constexpr float kNormEps = 1e-8f;      // coord2radial, :353

// ---- packed weight layout (floats).  *_f: [k][c] = W[c][k] (forward, lane c reads its row coalesced),
//      *_b: [k][c] = W[k][c] (torch layout; the transposed product needed by the reverse pass).
namespace pk {
constexpr int kHeader = 96;  // embW0[32] embW1[32] embB[32]
constexpr int A_f = 0, B_f = 1024, A_b = 2048, B_b = 3072, W2_f = 4096, W2_b = 5120, Wc1_f = 6144, Wc1_b = 7168,
              W3h_f = 8192, W3h_b = 9216, W3a_f = 10240, W3a_b = 11264, W4_f = 12288, W4_b = 13312, c1 = 14336,
              d1 = 14368, b1 = 14400, b2 = 14432, wa = 14464, ba = 14496, bc1 = 14528, wc2 = 14560, b3 = 14592,
              b4 = 14624, kLayer = 14656;
}  // namespace pk

__device__ __forceinline__ void load_row(float (&w)[H], const float *__restrict__ src, int lane) {
#pragma unroll
  for (int k = 0; k < H; ++k) w[k] = __ldg(src + k * H + lane);
}

// sum_k w[k] * v[k], v warp-uniform shared-memory vector (16-byte aligned)
__device__ __forceinline__ float dot32(const float (&w)[H], const float *v) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < H / 4; ++k4) {
    const float4 q = *reinterpret_cast<const float4 *>(v + 4 * k4);
    a0 = fmaf(w[4 * k4 + 0], q.x, a0);
    a1 = fmaf(w[4 * k4 + 1], q.y, a1);
    a0 = fmaf(w[4 * k4 + 2], q.z, a0);
    a1 = fmaf(w[4 * k4 + 3], q.w, a1);
  }
  return a0 + a1;
}

// per-lane scalars of the edge/coord/attention MLPs of one layer
struct EdgeScal {
  float c1, d1, b2, wa, ba, bc1, wc2;
};
__device__ __forceinline__ EdgeScal load_edge_scal(const float *__restrict__ Wl, int lane) {
  EdgeScal s;
  s.c1 = __ldg(Wl + pk::c1 + lane); s.d1 = __ldg(Wl + pk::d1 + lane); s.b2 = __ldg(Wl + pk::b2 + lane);
  s.wa = __ldg(Wl + pk::wa + lane); s.ba = __ldg(Wl + pk::ba); s.bc1 = __ldg(Wl + pk::bc1 + lane);
  s.wc2 = __ldg(Wl + pk::wc2 + lane);
  return s;
}

// geometry of one edge (warp-uniform)
struct EdgeGeo {
  float d[3];  // x_i - x_j
  float r2, nrm, inv, ea;
};
__device__ __forceinline__ EdgeGeo edge_geo(const float4 xi, const float4 xj, const float4 x0i, const float4 x0j) {
  EdgeGeo g;
  g.d[0] = xi.x - xj.x; g.d[1] = xi.y - xj.y; g.d[2] = xi.z - xj.z;
  g.r2 = g.d[0] * g.d[0] + g.d[1] * g.d[1] + g.d[2] * g.d[2];
  g.nrm = sqrtf(g.r2 + kNormEps);
  g.inv = 1.0f / (g.nrm + 1.0f);
  const float e0 = x0i.x - x0j.x, e1 = x0i.y - x0j.y, e2 = x0i.z - x0j.z;
  g.ea = e0 * e0 + e1 * e1 + e2 * e2;
  return g;
}

// per-lane primal quantities of one edge kept for the derivative passes
struct EdgeP {
  float f1, f2, fc;  // silu'(z1), silu'(z2), silu'(zc)
  float m, s, ms;    // m (pre-attention), attention gate, gated message
  float th, phi;     // tanh(u), phi = tanh(u) * range   (uniform)
};

// per-warp staging area: primal [2][H] then tangents [2][TTMAX][H]
template <int TTMAX>
struct Stage {
  static constexpr int kFloats = 2 * H + 2 * TTMAX * H;
  float *pa, *pb, *ta, *tb;
  __device__ __forceinline__ Stage(float *base) : pa(base), pb(base + H), ta(base + 2 * H), tb(base + 2 * H + TTMAX * H) {}
};

// Tangent bundle of one edge for TT directions.  dpq: d(p_i)+d(q_j) per lane; Dd, dea: uniform.
template <int TT>
struct EdgeT {
  float dpq[TT > 0 ? TT : 1];
  float Dd[TT > 0 ? TT : 1][3];
  float dea[TT > 0 ? TT : 1];
};

// Evaluates one edge: primal (edge MLP, attention, coord MLP) and TT tangents, sharing two __syncwarp()s.
// Outputs: primal EdgeP; tangent d(ms) per lane in dms[t] and d(trans) (uniform, 3 comps) in dtr[t].
template <int TT, int TTMAX>
__device__ __forceinline__ EdgeP edge_eval(const float (&w2)[H], const float (&wc1)[H], const EdgeScal &sc, float rng,
                                           float p_plus_q, const EdgeGeo &g, const Stage<TTMAX> &st, int lane,
                                           const EdgeT<TT> &tin, float (&dms)[TT > 0 ? TT : 1],
                                           float (&dtr)[TT > 0 ? TT : 1][3]) {
  EdgeP e;
  // ---- phase 1: first edge linear + SiLU
  const float z1 = p_plus_q + sc.c1 * g.r2 + sc.d1 * g.ea;
  float a1;
  silu_both(z1, a1, e.f1);
  st.pa[lane] = a1;
  float dr2h[TT > 0 ? TT : 1];  // <d, Dd[t]>
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    dr2h[t] = g.d[0] * tin.Dd[t][0] + g.d[1] * tin.Dd[t][1] + g.d[2] * tin.Dd[t][2];
    const float dz1 = tin.dpq[t] + sc.c1 * (2.0f * dr2h[t]) + sc.d1 * tin.dea[t];
    st.ta[t * H + lane] = e.f1 * dz1;
  }
  __syncwarp();
  // ---- phase 2: second edge linear + SiLU, attention gate
  const float z2 = sc.b2 + dot32(w2, st.pa);
  silu_both(z2, e.m, e.f2);
  e.s = sigmoidf_fast(warp_sum(sc.wa * e.m) + sc.ba);
  e.ms = e.m * e.s;
  st.pb[lane] = e.ms;
  const float s1s = e.s * (1.0f - e.s);
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const float dm = e.f2 * dot32(w2, st.ta + t * H);
    const float ds = s1s * warp_sum(sc.wa * dm);
    dms[t] = dm * e.s + e.m * ds;
    st.tb[t * H + lane] = dms[t];
  }
  __syncwarp();
  // ---- phase 3: coordinate MLP
  const float zc = sc.bc1 + dot32(wc1, st.pb);
  float ac;
  silu_both(zc, ac, e.fc);
  const float u = warp_sum(sc.wc2 * ac);
  e.th = tanhf(u);
  e.phi = e.th * rng;
  const float dphi_du = rng * (1.0f - e.th * e.th);
  const float wfc = sc.wc2 * e.fc;
  const float k2 = g.inv * g.inv / g.nrm;
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const float du = warp_sum(wfc * dot32(wc1, st.tb + t * H));
    const float dphi = dphi_du * du;
    const float c = dr2h[t] * k2;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const float ddhat = tin.Dd[t][b] * g.inv - g.d[b] * c;
      dtr[t][b] = ddhat * e.phi + g.d[b] * g.inv * dphi;
    }
  }
  return e;
}

// ------------------------------------------------------------------------------------------------
// Shared-memory plan
// ------------------------------------------------------------------------------------------------
template <int NP, int NW, int L>
struct Plan {
  static constexpr int kThreads = NW * 32;
  static constexpr int kNPW = (NP + NW - 1) / NW;  // receiver nodes per warp
  // primal state (floats).  Everything up to kPersist must survive the forward pass (it is re-read by the
  // derivative passes); sH / sAgg are forward-only scratch and may be overlaid afterwards.
  static constexpr int oX = 0;                           // [L+1][NP][4]
  static constexpr int oQ = oX + (L + 1) * NP * 4;       // [L][NP][H]
  static constexpr int oP = oQ + L * NP * H;             // [L][NP][H]
  static constexpr int oZ3 = oP + L * NP * H;            // [L-1][NP][H]
  static constexpr int oRed = oZ3 + (L - 1) * NP * H;    // [64] scratch for CTA reductions
  static constexpr int kPersist = oRed + 64;
  static constexpr int oH = kPersist;                    // [NP][H]   current node features
  static constexpr int oAgg = oH + NP * H;               // [NP][H]
  static constexpr int kPrimal = oAgg + NP * H;
};

// Primal forward for the CTA's particle.  On return sX[0..L], sQ, sP, sZ3 hold the per-layer state.
template <int NP, int NW, int L>
__device__ void primal_forward(float *sm, const float *__restrict__ wpack, float tcond, float beta, float *stage_base) {
  using P = Plan<NP, NW, L>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sX = reinterpret_cast<float4 *>(sm + P::oX);
  float *sH = sm + P::oH, *sQ = sm + P::oQ, *sP = sm + P::oP, *sZ3 = sm + P::oZ3, *sAgg = sm + P::oAgg;
  Stage<1> st(stage_base + warp * Stage<1>::kFloats);
  const float rng = kCoordsRange / (float)L;

  // node embedding with the reference's cat/reshape feature layout (egnn_temp_conditioned.py:63-78):
  // node k sees (f[2k], f[2k+1]) of f = [t]*n ++ [beta]*n
  {
    const float e0 = __ldg(wpack + lane), e1 = __ldg(wpack + 32 + lane), eb = __ldg(wpack + 64 + lane);
    for (int i = warp; i < NP; i += NW) {
      const float f0 = (2 * i < NP) ? tcond : beta;
      const float f1 = (2 * i + 1 < NP) ? tcond : beta;
      sH[i * H + lane] = fmaf(e0, f0, fmaf(e1, f1, eb));
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const float *__restrict__ Wl = wpack + pk::kHeader + l * pk::kLayer;
    {  // p_i = A h_i + b1,  q_i = B h_i
      float wA[H], wB[H];
      load_row(wA, Wl + pk::A_f, lane);
      load_row(wB, Wl + pk::B_f, lane);
      const float b1 = __ldg(Wl + pk::b1 + lane);
      for (int i = warp; i < NP; i += NW) {
        sP[(l * NP + i) * H + lane] = b1 + dot32(wA, sH + i * H);
        sQ[(l * NP + i) * H + lane] = dot32(wB, sH + i * H);
      }
    }
    __syncthreads();
    {  // edges: receiver i, senders j
      float w2[H], wc1[H];
      load_row(w2, Wl + pk::W2_f, lane);
      load_row(wc1, Wl + pk::Wc1_f, lane);
      const EdgeScal sc = load_edge_scal(Wl, lane);
      EdgeT<0> tin;
      float d0[1], d1[1][3];
      for (int i = warp; i < NP; i += NW) {
        const float4 xi = sX[l * NP + i], x0i = sX[i];
        const float pi = sP[(l * NP + i) * H + lane];
        float agg = 0.f, dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < NP; ++j) {
          if (j == i) continue;
          const EdgeGeo g = edge_geo(xi, sX[l * NP + j], x0i, sX[j]);
          const EdgeP e = edge_eval<0, 1>(w2, wc1, sc, rng, pi + sQ[(l * NP + j) * H + lane], g, st, lane, tin, d0, d1);
          agg += e.ms;
          const float f = g.inv * e.phi;
          dx0 = fmaf(g.d[0], f, dx0); dx1 = fmaf(g.d[1], f, dx1); dx2 = fmaf(g.d[2], f, dx2);
        }
        if (lane == 0) sX[(l + 1) * NP + i] = make_float4(xi.x + dx0, xi.y + dx1, xi.z + dx2, 0.f);
        sAgg[i * H + lane] = agg;
      }
    }
    __syncwarp();
    if (l < L - 1) {  // node update (dead code for the output in the last layer)
      float wa_[H], wb_[H];
      load_row(wa_, Wl + pk::W3h_f, lane);
      load_row(wb_, Wl + pk::W3a_f, lane);
      const float b3 = __ldg(Wl + pk::b3 + lane), b4 = __ldg(Wl + pk::b4 + lane);
      for (int i = warp; i < NP; i += NW) {
        const float z3 = b3 + dot32(wa_, sH + i * H) + dot32(wb_, sAgg + i * H);
        sZ3[(l * NP + i) * H + lane] = z3;
        __syncwarp();                       // every lane has read row i of sAgg (racecheck: profiles/r2p_racecheck_*.txt)
        sAgg[i * H + lane] = silu_val(z3);  // reuse as the input of the second node linear
      }
      __syncwarp();
      load_row(wa_, Wl + pk::W4_f, lane);
      for (int i = warp; i < NP; i += NW) sH[i * H + lane] += b4 + dot32(wa_, sAgg + i * H);
    }
    __syncthreads();
  }
}

// mean over nodes of sX[layer] (all threads must call); returns float4 mean
template <int NP>
__device__ float4 node_mean(const float4 *x, float *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = lane; i < NP; i += 32) { a += x[i].x; b += x[i].y; c += x[i].z; }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { red[0] = a / NP; red[1] = b / NP; red[2] = c / NP; }
  }
  __syncthreads();
  const float4 m = make_float4(red[0], red[1], red[2], 0.f);
  __syncthreads();
  return m;
}

}  // namespace pita
