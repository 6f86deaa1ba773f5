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
  // primal state (floats)
  static constexpr int oX = 0;                           // [L+1][NP][4]
  static constexpr int oH = oX + (L + 1) * NP * 4;       // [NP][H]   current node features
  static constexpr int oQ = oH + NP * H;                 // [L][NP][H]
  static constexpr int oP = oQ + L * NP * H;             // [L][NP][H]
  static constexpr int oZ3 = oP + L * NP * H;            // [L-1][NP][H]
  static constexpr int oAgg = oZ3 + (L - 1) * NP * H;    // [NP][H]
  static constexpr int oRed = oAgg + NP * H;             // [64] scratch for CTA reductions
  static constexpr int kPrimal = oRed + 64;
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

// ================================================================================================
// Kernel 1: plain forward   vel = EGNN_dynamics(tcond, y, beta)
// ================================================================================================
template <int NP, int NW, int L>
__global__ void __launch_bounds__(NW * 32)
egnn_forward_kernel(const float *__restrict__ wpack, const float *__restrict__ tcond, const float *__restrict__ y,
                    const float *__restrict__ beta, int64_t B, float *__restrict__ vel) {
  using P = Plan<NP, NW, L>;
  extern __shared__ __align__(16) float sm[];
  float *stage = sm + P::kPrimal;
  float4 *sX = reinterpret_cast<float4 *>(sm + P::oX);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int i = threadIdx.x; i < NP; i += P::kThreads)
      sX[i] = make_float4(y[b * 3 * NP + 3 * i], y[b * 3 * NP + 3 * i + 1], y[b * 3 * NP + 3 * i + 2], 0.f);
    __syncthreads();
    primal_forward<NP, NW, L>(sm, wpack, __ldg(tcond + b), __ldg(beta + b), stage);
    // vel = remove_mean(x_L - x_0)
    float4 *sV = reinterpret_cast<float4 *>(sm + P::oAgg);
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      const float4 a = sX[L * NP + i], c = sX[i];
      sV[i] = make_float4(a.x - c.x, a.y - c.y, a.z - c.z, 0.f);
    }
    __syncthreads();
    const float4 mean = node_mean<NP>(sV, sm + P::oRed);
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      vel[b * 3 * NP + 3 * i + 0] = sV[i].x - mean.x;
      vel[b * 3 * NP + 3 * i + 1] = sV[i].y - mean.y;
      vel[b * 3 * NP + 3 * i + 2] = sV[i].z - mean.z;
    }
    __syncthreads();
  }
}

// ================================================================================================
// Kernel 2: energy net — E, grad_x E, dE/dh by a hand-written reverse pass
// ================================================================================================
template <int NP, int NW, int L>
struct EPlan {
  using P = Plan<NP, NW, L>;
  static constexpr int oGH = P::kPrimal;            // [NP][H]  cotangent of node features
  static constexpr int oGAgg = oGH + NP * H;        // [NP][H]
  static constexpr int oGP = oGAgg + NP * H;        // [NP][H]
  static constexpr int oGQw = oGP + NP * H;         // [NW][NP][H] per-warp private scatter targets (deterministic)
  static constexpr int oGX = oGQw + NW * NP * H;    // [NP][4]   cotangent of coordinates (layer input side)
  static constexpr int oGXw = oGX + NP * 4;         // [NW][NP][4] per-warp private, coordinate path
  static constexpr int oGX0w = oGXw + NW * NP * 4;  // [NW][NP][4] per-warp private, edge_attr path (input coords)
  static constexpr int oStage = oGX0w + NW * NP * 4;
  static constexpr int kFloats = oStage + NW * Stage<1>::kFloats;
};

template <int NP, int NW, int L>
__global__ void __launch_bounds__(NW * 32)
egnn_energy_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                   const float *__restrict__ beta, int64_t B, float *__restrict__ energy, float *__restrict__ grad_x,
                   float *__restrict__ dE_dh) {
  using P = Plan<NP, NW, L>;
  using E = EPlan<NP, NW, L>;
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sX = reinterpret_cast<float4 *>(sm + P::oX);
  float *sQ = sm + P::oQ, *sP = sm + P::oP, *sZ3 = sm + P::oZ3, *sRed = sm + P::oRed;
  float *sGH = sm + E::oGH, *sGAgg = sm + E::oGAgg, *sGP = sm + E::oGP, *sGQw = sm + E::oGQw;
  float4 *sGX = reinterpret_cast<float4 *>(sm + E::oGX);
  float *sGXw = sm + E::oGXw, *sGX0w = sm + E::oGX0w;
  float *stage = sm + E::oStage;
  Stage<1> st(stage + warp * Stage<1>::kFloats);
  const float rng = kCoordsRange / (float)L;
  const bool want_grad = (grad_x != nullptr) || (dE_dh != nullptr);

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float h = __ldg(ht + b);
    const float c_in = rsqrtf(1.0f + h);
    const float c_noise = 0.125f * logf(h);
    for (int i = threadIdx.x; i < NP; i += P::kThreads)
      sX[i] = make_float4(c_in * x[b * 3 * NP + 3 * i], c_in * x[b * 3 * NP + 3 * i + 1], c_in * x[b * 3 * NP + 3 * i + 2], 0.f);
    __syncthreads();
    primal_forward<NP, NW, L>(sm, wpack, c_noise, __ldg(beta + b), stage);

    // vel, U = <vel, y>, cotangent w = y - mean(y)
    const float4 ymean = node_mean<NP>(sX, sRed);
    float4 *sV = reinterpret_cast<float4 *>(sm + P::oAgg);  // vel before mean removal
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      const float4 a = sX[L * NP + i], c = sX[i];
      sV[i] = make_float4(a.x - c.x, a.y - c.y, a.z - c.z, 0.f);
    }
    __syncthreads();
    const float4 vmean = node_mean<NP>(sV, sRed);
    // U and |x|^2 (x = y / c_in) by warp 0
    if (warp == 0) {
      float u = 0.f, y2 = 0.f;
      for (int i = lane; i < NP; i += 32) {
        const float4 yv = sX[i], v = sV[i];
        u += (v.x - vmean.x) * yv.x + (v.y - vmean.y) * yv.y + (v.z - vmean.z) * yv.z;
        y2 += yv.x * yv.x + yv.y * yv.y + yv.z * yv.z;
      }
      u = warp_sum(u); y2 = warp_sum(y2);
      if (lane == 0) { sRed[4] = u; sRed[5] = y2; }
    }
    __syncthreads();
    const float U = sRed[4];
    const float x2 = sRed[5] * (1.0f + h);  // |x|^2 = |y|^2 / c_in^2
    const float rs_h = rsqrtf(h);
    if (threadIdx.x == 0) energy[b] = x2 / (2.0f * (1.0f + h)) - rs_h * U;  // energy_net.py:37-39
    if (!want_grad) { __syncthreads(); continue; }

    // ---- reverse pass.  cotangent on x_L is w_i = y_i - mean(y)
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      const float4 yv = sX[i];
      sGX[i] = make_float4(yv.x - ymean.x, yv.y - ymean.y, yv.z - ymean.z, 0.f);
    }
    for (int k = threadIdx.x; k < NP * H; k += P::kThreads) { sGH[k] = 0.f; sGAgg[k] = 0.f; }
    for (int k = threadIdx.x; k < NW * NP * 4; k += P::kThreads) sGX0w[k] = 0.f;
    __syncthreads();

#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      const float *__restrict__ Wl = wpack + pk::kHeader + l * pk::kLayer;
      // zero the per-warp private scatter buffers of this layer
      for (int k = threadIdx.x; k < NW * NP * H; k += P::kThreads) sGQw[k] = 0.f;
      for (int k = threadIdx.x; k < NW * NP * 4; k += P::kThreads) sGXw[k] = 0.f;
      if (l < L - 1) {  // node-MLP reverse: gz3 = f3 * (W4^T gh'), gh += W3h^T gz3, gagg = W3a^T gz3
        float wr[H];
        load_row(wr, Wl + pk::W4_b, lane);
        float *stg = st.pa;
        for (int i = warp; i < NP; i += NW) {
          float v, f3;
          silu_both(sZ3[(l * NP + i) * H + lane], v, f3);
          const float gz3 = f3 * dot32(wr, sGH + i * H);
          __syncwarp();
          sGAgg[i * H + lane] = gz3;  // staged in place: only this warp touches row i here
        }
        __syncwarp();
        float wr2[H];
        load_row(wr, Wl + pk::W3h_b, lane);
        load_row(wr2, Wl + pk::W3a_b, lane);
        for (int i = warp; i < NP; i += NW) {
          const float add_h = dot32(wr, sGAgg + i * H);
          const float ga = dot32(wr2, sGAgg + i * H);
          __syncwarp();
          sGH[i * H + lane] += add_h;
          sGAgg[i * H + lane] = ga;
        }
        (void)stg;
      }
      __syncthreads();
      {  // edge reverse for receiver i (owner warp), scatter to sender j through this warp's private rows
        float w2[H], wc1[H], w2b[H], wc1b[H];
        load_row(w2, Wl + pk::W2_f, lane);
        load_row(wc1, Wl + pk::Wc1_f, lane);
        load_row(w2b, Wl + pk::W2_b, lane);
        load_row(wc1b, Wl + pk::Wc1_b, lane);
        const EdgeScal sc = load_edge_scal(Wl, lane);
        EdgeT<0> tin;
        float d0[1], d1[1][3];
        float *gqw = sGQw + warp * NP * H;
        float *gxw = sGXw + warp * NP * 4;
        float *gx0w = sGX0w + warp * NP * 4;
        for (int i = warp; i < NP; i += NW) {
          const float4 xi = sX[l * NP + i], x0i = sX[i];
          const float pi = sP[(l * NP + i) * H + lane];
          const float4 gxo = sGX[i];  // cotangent of x'_i
          const float gagg = sGAgg[i * H + lane];
          float gp = 0.f, gxi0 = 0.f, gxi1 = 0.f, gxi2 = 0.f, g0i0 = 0.f, g0i1 = 0.f, g0i2 = 0.f;
#pragma unroll 1
          for (int j = 0; j < NP; ++j) {
            if (j == i) continue;
            const float4 x0j = sX[j];
            const EdgeGeo g = edge_geo(xi, sX[l * NP + j], x0i, x0j);
            const EdgeP e = edge_eval<0, 1>(w2, wc1, sc, rng, pi + sQ[(l * NP + j) * H + lane], g, st, lane, tin, d0, d1);
            // coordinate branch
            const float gphi = (gxo.x * g.d[0] + gxo.y * g.d[1] + gxo.z * g.d[2]) * g.inv;
            const float gu = gphi * rng * (1.0f - e.th * e.th);
            const float gzc = gu * sc.wc2 * e.fc;
            __syncwarp();
            st.pa[lane] = gzc;
            __syncwarp();
            const float gms = gagg + dot32(wc1b, st.pa);
            const float gs = warp_sum(gms * e.m);
            const float gm = gms * e.s + sc.wa * (gs * e.s * (1.0f - e.s));
            const float gz2 = gm * e.f2;
            st.pb[lane] = gz2;
            __syncwarp();
            const float gz1 = dot32(w2b, st.pb) * e.f1;
            gp += gz1;
            gqw[j * H + lane] += gz1;
            const float gr2 = warp_sum(sc.c1 * gz1);
            const float gea = warp_sum(sc.d1 * gz1);
            // d/d(delta): through dhat = delta/(nrm+1) (times phi) and through r2
            const float gd_dot = (gxo.x * g.d[0] + gxo.y * g.d[1] + gxo.z * g.d[2]) * e.phi;
            const float k2 = gd_dot * g.inv * g.inv / g.nrm;
            const float a0 = gxo.x * e.phi * g.inv - g.d[0] * k2 + 2.0f * g.d[0] * gr2;
            const float a1 = gxo.y * e.phi * g.inv - g.d[1] * k2 + 2.0f * g.d[1] * gr2;
            const float a2 = gxo.z * e.phi * g.inv - g.d[2] * k2 + 2.0f * g.d[2] * gr2;
            gxi0 += a0; gxi1 += a1; gxi2 += a2;
            const float e0 = 2.0f * (x0i.x - x0j.x) * gea, e1 = 2.0f * (x0i.y - x0j.y) * gea, e2 = 2.0f * (x0i.z - x0j.z) * gea;
            g0i0 += e0; g0i1 += e1; g0i2 += e2;
            if (lane == 0) {
              gxw[j * 4 + 0] -= a0; gxw[j * 4 + 1] -= a1; gxw[j * 4 + 2] -= a2;
              gx0w[j * 4 + 0] -= e0; gx0w[j * 4 + 1] -= e1; gx0w[j * 4 + 2] -= e2;
            }
          }
          sGP[i * H + lane] = gp;
          if (lane == 0) {
            gxw[i * 4 + 0] += gxi0; gxw[i * 4 + 1] += gxi1; gxw[i * 4 + 2] += gxi2;
            gx0w[i * 4 + 0] += g0i0; gx0w[i * 4 + 1] += g0i1; gx0w[i * 4 + 2] += g0i2;
          }
          __syncwarp();
        }
      }
      __syncthreads();
      {  // combine: gh_i += A^T gp_i + B^T gq_i ; gx_i += sum_w gxw[w][i]   (fixed order -> deterministic)
        float wr[H], wr2[H];
        load_row(wr, Wl + pk::A_b, lane);
        load_row(wr2, Wl + pk::B_b, lane);
        for (int i = warp; i < NP; i += NW) {
          float gq = 0.f;
#pragma unroll 1
          for (int w = 0; w < NW; ++w) gq += sGQw[(w * NP + i) * H + lane];
          __syncwarp();
          sGAgg[i * H + lane] = gq;  // stage (gagg is dead after the edge pass)
          __syncwarp();
          sGH[i * H + lane] += dot32(wr, sGP + i * H) + dot32(wr2, sGAgg + i * H);
          if (lane < 3) {
            float acc = 0.f;
#pragma unroll 1
            for (int w = 0; w < NW; ++w) acc += sGXw[(w * NP + i) * 4 + lane];
            reinterpret_cast<float *>(sGX)[i * 4 + lane] += acc;
          }
        }
      }
      __syncthreads();
    }

    // ---- assemble outputs
    // dU/dy_i = (vel_i) - w_i + gx_i + gx0_i ;  dU/dtcond = sum_i <gh0_i, d h0_i/d tcond>
    float part_dt = 0.f;
    {
      const float e0 = __ldg(wpack + lane), e1 = __ldg(wpack + 32 + lane);
      for (int i = warp; i < NP; i += NW) {
        const float dh0 = ((2 * i < NP) ? e0 : 0.f) + ((2 * i + 1 < NP) ? e1 : 0.f);
        part_dt += sGH[i * H + lane] * dh0;
      }
      part_dt = warp_sum(part_dt);
    }
    // per-thread over nodes: dU/dy and <dU/dy, x>
    float part_dot = 0.f;
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      float g0[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int w = 0; w < NW; ++w) {
        g0[0] += sGX0w[(w * NP + i) * 4 + 0]; g0[1] += sGX0w[(w * NP + i) * 4 + 1]; g0[2] += sGX0w[(w * NP + i) * 4 + 2];
      }
      const float4 yv = sX[i], v = sV[i], gx = sGX[i];
      const float wv[3] = {yv.x - ymean.x, yv.y - ymean.y, yv.z - ymean.z};
      const float vv[3] = {v.x - vmean.x, v.y - vmean.y, v.z - vmean.z};
      const float gxx[3] = {gx.x, gx.y, gx.z};
      const float yy[3] = {yv.x, yv.y, yv.z};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float dUdy = vv[k] - wv[k] + gxx[k] + g0[k];
        const float xk = yy[k] / c_in;
        part_dot += dUdy * xk;
        if (grad_x) grad_x[b * 3 * NP + 3 * i + k] = xk / (1.0f + h) - rs_h * c_in * dUdy;  // energy_net.py:37-39,61
      }
    }
    part_dot = warp_sum(part_dot);
    __syncthreads();
    if (lane == 0) { sRed[8 + warp] = part_dt; sRed[8 + NW + warp] = part_dot; }
    __syncthreads();
    if (threadIdx.x == 0 && dE_dh) {
      float dUdc = 0.f, dot = 0.f;
      for (int w = 0; w < NW; ++w) { dUdc += sRed[8 + w]; dot += sRed[8 + NW + w]; }
      const float op = 1.0f + h;
      const float dU_dh = dUdc / (8.0f * h) + dot * (-0.5f) * rsqrtf(op) / op;
      dE_dh[b] = -x2 / (2.0f * op * op) + 0.5f * rs_h / h * U - rs_h * dU_dh;
    }
    __syncthreads();
  }
}

// ================================================================================================
// Kernel 3: score net — score and exact divergence by forward-mode tangents
//   TN tangent nodes (T = 3*TN directions) per pass; layer 0 touches only edges incident to a tangent
//   node, the middle layer is dense, the last layer only receivers that are tangent nodes.
// ================================================================================================
template <int NP, int NW, int L, int TN>
struct DPlan {
  using P = Plan<NP, NW, L>;
  static constexpr int T = 3 * TN;
  static constexpr int oDQ = P::kPrimal;            // [NP][T][H]   B^1 dh^1  (read by every receiver in the dense layer)
  static constexpr int oDXa = oDQ + NP * T * H;     // [NP][T][4]   d x^1
  static constexpr int oDHQ = oDXa + NP * T * 4;    // [NP][T][H]   dh^1 (owner-private) then B^2 dh^2 (published)
  static constexpr int oDXb = oDHQ + NP * T * H;    // [NP][T][4]   d x^2
  static constexpr int oDP2 = oDXb + NP * T * 4;    // [TN][3][H]   A^2 dh^2 of the tangent nodes (own directions)
  static constexpr int oStage = oDP2 + TN * 3 * H;
  static constexpr int kFloats = oStage + NW * Stage<T>::kFloats;
};

template <int NP, int NW, int L, int TN>
__global__ void __launch_bounds__(NW * 32)
egnn_score_div_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                      const float *__restrict__ beta, int64_t B, float *__restrict__ score, float *__restrict__ divergence) {
  static_assert(L == 3, "tangent pass is written for first/dense/last = 3 layers");
  using P = Plan<NP, NW, L>;
  using Dp = DPlan<NP, NW, L, TN>;
  constexpr int T = Dp::T;
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sX = reinterpret_cast<float4 *>(sm + P::oX);
  float *sQ = sm + P::oQ, *sP = sm + P::oP, *sZ3 = sm + P::oZ3, *sRed = sm + P::oRed;
  float *sDQ = sm + Dp::oDQ, *sDHQ = sm + Dp::oDHQ, *sDP2 = sm + Dp::oDP2;
  float4 *sDXa = reinterpret_cast<float4 *>(sm + Dp::oDXa);
  float4 *sDXb = reinterpret_cast<float4 *>(sm + Dp::oDXb);
  float *stage = sm + Dp::oStage;
  Stage<T> st(stage + warp * Stage<T>::kFloats);
  const float rng = kCoordsRange / (float)L;
  const float *__restrict__ W0 = wpack + pk::kHeader;
  const float *__restrict__ W1 = W0 + pk::kLayer;
  const float *__restrict__ W2l = W1 + pk::kLayer;

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float h = __ldg(ht + b);
    const float c_in = rsqrtf(1.0f + h);
    const float c_s = 1.0f / (1.0f + h);
    const float c_out = sqrtf(h) * c_in;
    const float c_noise = 0.125f * logf(h);
    for (int i = threadIdx.x; i < NP; i += P::kThreads)
      sX[i] = make_float4(c_in * x[b * 3 * NP + 3 * i], c_in * x[b * 3 * NP + 3 * i + 1], c_in * x[b * 3 * NP + 3 * i + 2], 0.f);
    __syncthreads();
    primal_forward<NP, NW, L>(sm, wpack, c_noise, __ldg(beta + b), stage);

    // ---- score = ((c_s - 1) x + c_out * vel) / h        (score_net.py:21-43)
    {
      float4 *sV = reinterpret_cast<float4 *>(sm + P::oAgg);
      for (int i = threadIdx.x; i < NP; i += P::kThreads) {
        const float4 a = sX[L * NP + i], c = sX[i];
        sV[i] = make_float4(a.x - c.x, a.y - c.y, a.z - c.z, 0.f);
      }
      __syncthreads();
      const float4 vmean = node_mean<NP>(sV, sRed);
      for (int i = threadIdx.x; i < NP; i += P::kThreads) {
        const float vv[3] = {sV[i].x - vmean.x, sV[i].y - vmean.y, sV[i].z - vmean.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float xv = x[b * 3 * NP + 3 * i + k];
          const float den = c_s * xv + c_out * vv[k];
          score[b * 3 * NP + 3 * i + k] = (den - xv) / h;
        }
      }
      __syncthreads();
    }
    if (divergence == nullptr) continue;

    // ---- tangent passes.  trace accumulates  sum_k sum_a d x_L[k,a] / d y[k,a]
    float trace = 0.f;  // per-warp partial (uniform across lanes)
#pragma unroll 1
    for (int k0 = 0; k0 < NP; k0 += TN) {
      // =========================== phase A: layer 0 (sparse) + node update + layer-1 pre-products
      {
        float w2[H], wc1[H];
        load_row(w2, W0 + pk::W2_f, lane);
        load_row(wc1, W0 + pk::Wc1_f, lane);
        const EdgeScal sc = load_edge_scal(W0, lane);
        for (int i = warp; i < NP; i += NW) {
          float dagg[T];
          float dxi[T][3];
#pragma unroll
          for (int t = 0; t < T; ++t) { dagg[t] = 0.f; dxi[t][0] = dxi[t][1] = dxi[t][2] = 0.f; }
          const float4 xi = sX[i];
          const float pi = sP[i * H + lane];
          const int ii = i - k0;  // index of i among the tangent nodes, if 0 <= ii < TN
          // (a) i is a tangent node: its own directions act on every edge (i, j)
          // (b) every tangent node j != i acts on edge (i, j) with the opposite sign
#pragma unroll 1
          for (int j = 0; j < NP; ++j) {
            if (j == i) continue;
            const int jj = j - k0;
            const bool own = (ii >= 0 && ii < TN), oth = (jj >= 0 && jj < TN && j < NP);
            if (!own && !oth) continue;
            const EdgeGeo g = edge_geo(xi, sX[j], xi, sX[j]);
            const float pq = pi + sQ[j * H + lane];
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
              if (which == 0 && !own) continue;
              if (which == 1 && !oth) continue;
              const float sgn = which == 0 ? 1.0f : -1.0f;
              const int slot = which == 0 ? ii : jj;
              EdgeT<3> tin;
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                tin.dpq[a] = 0.f;
                tin.Dd[a][0] = a == 0 ? sgn : 0.f; tin.Dd[a][1] = a == 1 ? sgn : 0.f; tin.Dd[a][2] = a == 2 ? sgn : 0.f;
                tin.dea[a] = 2.0f * sgn * g.d[a];  // layer 0: edge_attr == radial
              }
              float dms[3], dtr[3][3];
              edge_eval<3, T>(w2, wc1, sc, rng, pq, g, st, lane, tin, dms, dtr);
#pragma unroll
              for (int t = 0; t < T; ++t) {
                if (t / 3 == slot) {
                  dagg[t] += dms[t % 3];
                  dxi[t][0] += dtr[t % 3][0]; dxi[t][1] += dtr[t % 3][1]; dxi[t][2] += dtr[t % 3][2];
                }
              }
            }
          }
          // identity path of the coordinates: d x^0_i / d y_{k,a}
#pragma unroll
          for (int t = 0; t < T; ++t)
            if (t / 3 == ii) dxi[t][t % 3] += 1.0f;
          if (lane < T) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < T; ++t)
              if (t == lane) o = make_float4(dxi[t][0], dxi[t][1], dxi[t][2], 0.f);
            sDXa[i * T + lane] = o;
          }
          // stash dagg for the node update below (owner-private rows of sDHQ)
#pragma unroll
          for (int t = 0; t < T; ++t) sDHQ[(i * T + t) * H + lane] = dagg[t];
        }
        __syncwarp();
      }
      {  // node update of layer 0 on the tangents: dh^1 = W4 (f3 * (W3a dagg))   (dh^0 = 0)
        float wr[H];
        load_row(wr, W0 + pk::W3a_f, lane);
        for (int i = warp; i < NP; i += NW) {
          float v, f3;
          silu_both(sZ3[i * H + lane], v, f3);
          float tmp[T];
#pragma unroll
          for (int t = 0; t < T; ++t) tmp[t] = f3 * dot32(wr, sDHQ + (i * T + t) * H);
          __syncwarp();
#pragma unroll
          for (int t = 0; t < T; ++t) sDHQ[(i * T + t) * H + lane] = tmp[t];
        }
        __syncwarp();
        load_row(wr, W0 + pk::W4_f, lane);
        for (int i = warp; i < NP; i += NW) {
          float tmp[T];
#pragma unroll
          for (int t = 0; t < T; ++t) tmp[t] = dot32(wr, sDHQ + (i * T + t) * H);
          __syncwarp();
#pragma unroll
          for (int t = 0; t < T; ++t) sDHQ[(i * T + t) * H + lane] = tmp[t];  // dh^1_i
        }
        __syncwarp();
        load_row(wr, W1 + pk::B_f, lane);
        for (int i = warp; i < NP; i += NW) {
#pragma unroll
          for (int t = 0; t < T; ++t) sDQ[(i * T + t) * H + lane] = dot32(wr, sDHQ + (i * T + t) * H);  // B^1 dh^1_i
        }
      }
      __syncthreads();
      // =========================== phase B: layer 1 (dense) + node update + layer-2 pre-products
      {
        float w2[H], wc1[H];
        load_row(w2, W1 + pk::W2_f, lane);
        load_row(wc1, W1 + pk::Wc1_f, lane);
        const EdgeScal sc = load_edge_scal(W1, lane);
        for (int i = warp; i < NP; i += NW) {
          const int ii = i - k0;
          float dp[T];
          {
            float wr[H];
            load_row(wr, W1 + pk::A_f, lane);
#pragma unroll
            for (int t = 0; t < T; ++t) dp[t] = dot32(wr, sDHQ + (i * T + t) * H);
          }
          float dagg[T], dxo[T][3], dxi[T][3];
#pragma unroll
          for (int t = 0; t < T; ++t) {
            dagg[t] = 0.f;
            const float4 q = sDXa[i * T + t];
            dxi[t][0] = q.x; dxi[t][1] = q.y; dxi[t][2] = q.z;
            dxo[t][0] = q.x; dxo[t][1] = q.y; dxo[t][2] = q.z;  // identity path x^2 = x^1 + ...
          }
          const float4 xi = sX[NP + i], x0i = sX[i];
          const float pi = sP[(NP + i) * H + lane];
#pragma unroll 1
          for (int j = 0; j < NP; ++j) {
            if (j == i) continue;
            const int jj = j - k0;
            const float4 x0j = sX[j];
            const EdgeGeo g = edge_geo(xi, sX[NP + j], x0i, x0j);
            EdgeT<T> tin;
            const float e0[3] = {x0i.x - x0j.x, x0i.y - x0j.y, x0i.z - x0j.z};
#pragma unroll
            for (int t = 0; t < T; ++t) {
              tin.dpq[t] = dp[t] + sDQ[(j * T + t) * H + lane];
              const float4 q = sDXa[j * T + t];
              tin.Dd[t][0] = dxi[t][0] - q.x; tin.Dd[t][1] = dxi[t][1] - q.y; tin.Dd[t][2] = dxi[t][2] - q.z;
              // d edge_attr: only directions of tangent node i (+) or j (-)
              const float sgn = (t / 3 == ii) ? 1.0f : ((t / 3 == jj) ? -1.0f : 0.0f);
              tin.dea[t] = 2.0f * sgn * e0[t % 3];
            }
            float dms[T], dtr[T][3];
            edge_eval<T, T>(w2, wc1, sc, rng, pi + sQ[(NP + j) * H + lane], g, st, lane, tin, dms, dtr);
#pragma unroll
            for (int t = 0; t < T; ++t) {
              dagg[t] += dms[t];
              dxo[t][0] += dtr[t][0]; dxo[t][1] += dtr[t][1]; dxo[t][2] += dtr[t][2];
            }
          }
          if (lane < T) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < T; ++t)
              if (t == lane) o = make_float4(dxo[t][0], dxo[t][1], dxo[t][2], 0.f);
            sDXb[i * T + lane] = o;
          }
          // node update: dz3 = W3h dh^1 + W3a dagg ; dh^2 = dh^1 + W4 (f3 * dz3)
          float dz3[T];
          {
            float wr[H];
            load_row(wr, W1 + pk::W3h_f, lane);
#pragma unroll
            for (int t = 0; t < T; ++t) dz3[t] = dot32(wr, sDHQ + (i * T + t) * H);
            __syncwarp();
#pragma unroll
            for (int t = 0; t < T; ++t) st.ta[t * H + lane] = dagg[t];
            __syncwarp();
            load_row(wr, W1 + pk::W3a_f, lane);
            float v, f3;
            silu_both(sZ3[(NP + i) * H + lane], v, f3);
#pragma unroll
            for (int t = 0; t < T; ++t) dz3[t] = f3 * (dz3[t] + dot32(wr, st.ta + t * H));
            __syncwarp();
#pragma unroll
            for (int t = 0; t < T; ++t) st.ta[t * H + lane] = dz3[t];
            __syncwarp();
            load_row(wr, W1 + pk::W4_f, lane);
            float dh2[T];
#pragma unroll
            for (int t = 0; t < T; ++t) dh2[t] = sDHQ[(i * T + t) * H + lane] + dot32(wr, st.ta + t * H);
            __syncwarp();
#pragma unroll
            for (int t = 0; t < T; ++t) st.ta[t * H + lane] = dh2[t];
            __syncwarp();
            // publish B^2 dh^2_i (all directions) and, for a tangent node, A^2 dh^2_i of its own directions
            load_row(wr, W2l + pk::B_f, lane);
#pragma unroll
            for (int t = 0; t < T; ++t) sDHQ[(i * T + t) * H + lane] = dot32(wr, st.ta + t * H);
            if (ii >= 0 && ii < TN) {
              load_row(wr, W2l + pk::A_f, lane);
#pragma unroll
              for (int t = 0; t < T; ++t)
                if (t / 3 == ii) sDP2[(ii * 3 + t % 3) * H + lane] = dot32(wr, st.ta + t * H);
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
      // =========================== phase C: last layer, receivers = tangent nodes, own directions only.
      // Edge (k, j) is evaluated by the owner warp of the SENDER j so the work spreads over the CTA.
      {
        float w2[H], wc1[H];
        load_row(w2, W2l + pk::W2_f, lane);
        load_row(wc1, W2l + pk::Wc1_f, lane);
        const EdgeScal sc = load_edge_scal(W2l, lane);
        for (int j = warp; j < NP; j += NW) {
#pragma unroll 1
          for (int kk = 0; kk < TN; ++kk) {
            const int k = k0 + kk;
            if (k >= NP || k == j) continue;
            const float4 xk = sX[2 * NP + k], xj = sX[2 * NP + j], x0k = sX[k], x0j = sX[j];
            const EdgeGeo g = edge_geo(xk, xj, x0k, x0j);
            EdgeT<3> tin;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const int t = kk * 3 + a;
              tin.dpq[a] = sDP2[(kk * 3 + a) * H + lane] + sDHQ[(j * T + t) * H + lane];
              const float4 qk = sDXb[k * T + t], qj = sDXb[j * T + t];
              tin.Dd[a][0] = qk.x - qj.x; tin.Dd[a][1] = qk.y - qj.y; tin.Dd[a][2] = qk.z - qj.z;
            }
            tin.dea[0] = 2.0f * (x0k.x - x0j.x); tin.dea[1] = 2.0f * (x0k.y - x0j.y); tin.dea[2] = 2.0f * (x0k.z - x0j.z);
            float dms[3], dtr[3][3];
            edge_eval<3, T>(w2, wc1, sc, rng, sP[(2 * NP + k) * H + lane] + sQ[(2 * NP + j) * H + lane], g, st, lane, tin, dms, dtr);
            trace += dtr[0][0] + dtr[1][1] + dtr[2][2];
          }
          // identity path x^3_k = x^2_k + ... for a tangent node owned here
          const int jj = j - k0;
          if (jj >= 0 && jj < TN) {
            const float4 q0 = sDXb[j * T + jj * 3 + 0], q1 = sDXb[j * T + jj * 3 + 1], q2 = sDXb[j * T + jj * 3 + 2];
            trace += q0.x + q1.y + q2.z;
          }
        }
      }
      __syncthreads();
    }
    // ---- reduce the per-warp traces (fixed order) and finish:  div = ((c_s-1) D + c_out c_in (tr - D)) / h
    if (lane == 0) sRed[8 + warp] = trace;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tr = 0.f;
      for (int w = 0; w < NW; ++w) tr += sRed[8 + w];
      const float Dn = (float)(3 * NP);
      divergence[b] = ((c_s - 1.0f) * Dn + c_out * c_in * (tr - Dn)) / h;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes, const char *name) {
  if (bytes > 227 * 1024) { set_error("%s needs %zu bytes of shared memory (> 227 KB)", name, bytes); return PITA_EUNSUP; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e)); return PITA_ECUDA; }
  return PITA_OK;
}

static inline unsigned grid_for(int64_t B, int ctas_per_sm) {
  const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  return (unsigned)(B < cap ? B : cap);
}

template <int NP, int NW>
static int launch_forward(const float *w, const float *tc, const float *y, const float *beta, int64_t B, float *vel, cudaStream_t s) {
  using P = Plan<NP, NW, 3>;
  const size_t bytes = (P::kPrimal + NW * Stage<1>::kFloats) * sizeof(float);
  auto k = egnn_forward_kernel<NP, NW, 3>;
  int rc = set_smem(k, bytes, "egnn_forward_kernel");
  if (rc) return rc;
  k<<<grid_for(B, 4), NW * 32, bytes, s>>>(w, tc, y, beta, B, vel);
  PITA_CHECK_LAUNCH("egnn_forward_kernel");
  return PITA_OK;
}
template <int NP, int NW>
static int launch_energy(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *e, float *g, float *dh, cudaStream_t s) {
  using E = EPlan<NP, NW, 3>;
  const size_t bytes = E::kFloats * sizeof(float);
  auto k = egnn_energy_kernel<NP, NW, 3>;
  int rc = set_smem(k, bytes, "egnn_energy_kernel");
  if (rc) return rc;
  k<<<grid_for(B, 4), NW * 32, bytes, s>>>(w, ht, x, beta, B, e, g, dh);
  PITA_CHECK_LAUNCH("egnn_energy_kernel");
  return PITA_OK;
}
template <int NP, int NW, int TN>
static int launch_score(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *sc, float *dv, cudaStream_t s) {
  using D = DPlan<NP, NW, 3, TN>;
  const size_t bytes = D::kFloats * sizeof(float);
  auto k = egnn_score_div_kernel<NP, NW, 3, TN>;
  int rc = set_smem(k, bytes, "egnn_score_div_kernel");
  if (rc) return rc;
  k<<<grid_for(B, 4), NW * 32, bytes, s>>>(w, ht, x, beta, B, sc, dv);
  PITA_CHECK_LAUNCH("egnn_score_div_kernel");
  return PITA_OK;
}

static int check_common(const char *what, const float *w, int hidden, int layers, int n, const void *a, const void *b_, const void *c, int64_t B) {
  PITA_REQUIRE(B >= 0, PITA_EINVAL, "%s: negative batch", what);
  PITA_REQUIRE(B == 0 || (w && a && b_ && c), PITA_EINVAL, "%s: null pointer", what);
  if (hidden != 32 || layers != 3) { set_error("%s: only hidden_nf=32, n_layers=3 (configs/model/net/egnn_temp.yaml) is built; got %d/%d", what, hidden, layers); return PITA_EUNSUP; }
  if (n != 13 && n != 55) { set_error("%s: n_particles=%d unsupported (13 or 55)", what, n); return PITA_EUNSUP; }
  return PITA_OK;
}

}  // namespace pita

using namespace pita;

extern "C" int64_t pita_egnn_pack_floats(int hidden, int layers) {
  if (hidden != 32 || layers < 1) return -1;
  return pk::kHeader + (int64_t)layers * pk::kLayer;
}

extern "C" int pita_egnn_forward(const float *wpack, int hidden, int layers, int n, const float *tcond, const float *y,
                                 const float *beta, int64_t B, float *vel, void *stream) {
  int rc = check_common("egnn_forward", wpack, hidden, layers, n, tcond, y, beta, B);
  if (rc) return rc;
  if (B == 0) return PITA_OK;
  PITA_REQUIRE(vel, PITA_EINVAL, "egnn_forward: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return n == 13 ? launch_forward<13, 13>(wpack, tcond, y, beta, B, vel, s) : launch_forward<55, 11>(wpack, tcond, y, beta, B, vel, s);
}

extern "C" int pita_egnn_energy(const float *wpack, int hidden, int layers, int n, const float *ht, const float *x,
                                const float *beta, int64_t B, float *energy, float *grad_x, float *dE_dh, void *stream) {
  int rc = check_common("egnn_energy", wpack, hidden, layers, n, ht, x, beta, B);
  if (rc) return rc;
  if (B == 0) return PITA_OK;
  PITA_REQUIRE(energy, PITA_EINVAL, "egnn_energy: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return n == 13 ? launch_energy<13, 13>(wpack, ht, x, beta, B, energy, grad_x, dE_dh, s)
                 : launch_energy<55, 11>(wpack, ht, x, beta, B, energy, grad_x, dE_dh, s);
}

extern "C" int pita_egnn_score_div(const float *wpack, int hidden, int layers, int n, const float *ht, const float *x,
                                   const float *beta, int64_t B, float *score, float *div, void *stream) {
  int rc = check_common("egnn_score_div", wpack, hidden, layers, n, ht, x, beta, B);
  if (rc) return rc;
  if (B == 0) return PITA_OK;
  PITA_REQUIRE(score, PITA_EINVAL, "egnn_score_div: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return n == 13 ? launch_score<13, 13, 2>(wpack, ht, x, beta, B, score, div, s)
                 : launch_score<55, 11, 2>(wpack, ht, x, beta, B, score, div, s);
}
