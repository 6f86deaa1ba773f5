// Exact Laplacian of the model energy, tr(Hess_x E) -- the reference's `compute_laplacian_exact` (utils.py:68-77:
// vmap(hessian) then the trace), used by VEReverseSDE.f when there is no score net (sdes.py:150-153, 204-216):
//     b = -grad U g^2 / 2,   div b = -laplacian(U) g^2 / 2.
//
// Method: second-order Taylor-mode forward propagation.  For every coordinate direction e the network is evaluated on
// y + eps*e with every quantity carried as a truncated series  v + a*eps + b*eps^2  ("jet"); linear layers act on the
// three coefficients, a nonlinearity g maps (v, a, b) to (g(v), g'(v) a, g'(v) b + g''(v) a^2 / 2), products follow the
// Leibniz rule.  The eps^2 coefficient of u(y) = <vel(y), y> is half the second directional derivative, so
//     laplacian_y u = sum over the 3n directions of 2 b_u,      laplacian_x E = (1 - c_s) 3n / h - c_out c_in / h * laplacian_y u
// (E = (1-c_s)/(2h) |x|^2 - c_out/(c_in h) u(c_in x), energy_net.py:14-48).  No reverse pass and no [D, D] Hessian.
//
// Mapping: one CTA per particle, warp = receiver node (round-robin), lane = hidden channel, the same fp32 SIMT layout
// as egnn_common.cuh (weight rows in registers, broadcast operands from shared memory).  This branch is not on the
// default configuration's path (the reference ships a score net); it is built exact and simple rather than fast:
// 3n full jet forwards per particle.  Algebra checked against oracle/pita_oracle.py::exact_laplacian (fp64 autograd)
// and the reference's own output (tests/golden/fk_n13_laplacian.npz).
#include "egnn_common.cuh"

namespace pita {
namespace lap {

struct J {
  float v, a, b;
};
__device__ __forceinline__ J jmake(float v, float a, float b) { J r; r.v = v; r.a = a; r.b = b; return r; }
__device__ __forceinline__ J jadd(J x, J y) { return jmake(x.v + y.v, x.a + y.a, x.b + y.b); }
__device__ __forceinline__ J jsub(J x, J y) { return jmake(x.v - y.v, x.a - y.a, x.b - y.b); }
__device__ __forceinline__ J jmul(J x, J y) {
  return jmake(x.v * y.v, fmaf(x.v, y.a, x.a * y.v), fmaf(x.v, y.b, fmaf(x.a, y.a, x.b * y.v)));
}
__device__ __forceinline__ J jscale(float c, J x) { return jmake(c * x.v, c * x.a, c * x.b); }
__device__ __forceinline__ J jfma(float c, J x, J y) { return jmake(fmaf(c, x.v, y.v), fmaf(c, x.a, y.a), fmaf(c, x.b, y.b)); }
// g(x) from g, g', g'' at x.v
__device__ __forceinline__ J jcomp(J x, float g0, float g1, float g2) {
  return jmake(g0, g1 * x.a, fmaf(g1, x.b, 0.5f * g2 * x.a * x.a));
}
__device__ __forceinline__ J jsilu(J z) {
  const float s = 1.0f / (1.0f + __expf(-z.v)), t = s * (1.0f - s);
  return jcomp(z, z.v * s, s + z.v * t, t * (2.0f + z.v * (1.0f - 2.0f * s)));
}
__device__ __forceinline__ J jsigmoid(J z) {
  const float s = 1.0f / (1.0f + __expf(-z.v)), t = s * (1.0f - s);
  return jcomp(z, s, t, t * (1.0f - 2.0f * s));
}
__device__ __forceinline__ J jtanh(J z) {
  const float t = tanhf(z.v), u = 1.0f - t * t;
  return jcomp(z, t, u, -2.0f * t * u);
}
__device__ __forceinline__ J jsqrt(J z) {
  const float r = sqrtf(z.v), ir = 1.0f / r;
  return jcomp(z, r, 0.5f * ir, -0.25f * ir * ir * ir);
}
__device__ __forceinline__ J jrcp(J z) {
  const float i = 1.0f / z.v;
  return jcomp(z, i, -i * i, 2.0f * i * i * i);
}
__device__ __forceinline__ J jwarp_sum(J x) { return jmake(warp_sum(x.v), warp_sum(x.a), warp_sum(x.b)); }

// shared-memory jet arrays are stored as three planes [3][count]
template <int NP, int NW>
struct LPlan {
  static constexpr int kThreads = NW * 32;
  static constexpr int kNH = NP * H;
  static constexpr int oX = 0;                 // [2 buffers][3 planes][NP][4]
  static constexpr int oH = oX + 2 * 3 * NP * 4;   // [3][NP][H]
  static constexpr int oP = oH + 3 * kNH;
  static constexpr int oQ = oP + 3 * kNH;
  static constexpr int oAgg = oQ + 3 * kNH;
  static constexpr int oY = oAgg + 3 * kNH;    // [NP][4] input coordinates y = c_in x
  static constexpr int oStage = oY + NP * 4;   // per warp [2][3][H]
  static constexpr int kFloats = oStage + NW * 6 * H;
  static constexpr int kBytes = kFloats * 4;
};

__device__ __forceinline__ J dot3(const float (&w)[H], const float *planes, int stride) {
  return jmake(dot32(w, planes), dot32(w, planes + stride), dot32(w, planes + 2 * stride));
}

template <int NP, int NW, int L>
__global__ void __launch_bounds__(NW * 32, 1)
egnn_laplacian_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                      const float *__restrict__ beta, int64_t B, float *__restrict__ lap_out) {
  using P = LPlan<NP, NW>;
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *sXb = reinterpret_cast<float4 *>(sm + P::oX);
  float *sH = sm + P::oH, *sP = sm + P::oP, *sQ = sm + P::oQ, *sAgg = sm + P::oAgg;
  float4 *sY = reinterpret_cast<float4 *>(sm + P::oY);
  float *st = sm + P::oStage + warp * 6 * H;   // [pa: 3][H], [pb: 3][H]
  constexpr int NH = P::kNH;
  const float rng = kCoordsRange / (float)L;

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float h = __ldg(ht + b), bet = __ldg(beta + b);
    const float c_in = rsqrtf(1.0f + h), tcond = 0.125f * __logf(h);
    __syncthreads();
    for (int i = threadIdx.x; i < NP; i += P::kThreads) {
      const float *xp = x + b * (3 * NP) + 3 * i;
      sY[i] = make_float4(c_in * __ldg(xp), c_in * __ldg(xp + 1), c_in * __ldg(xp + 2), 0.f);
    }
    __syncthreads();
    float lap_acc = 0.f;   // warp 0 accumulates sum over directions of the eps^2 coefficient of u

#pragma unroll 1
    for (int dir = 0; dir < 3 * NP; ++dir) {
      const int k = dir / 3, c = dir - 3 * k;
      // ---- layer-0 state: x = y + eps e, h = embedding (no eps dependence)
      int cur = 0;
      for (int i = threadIdx.x; i < NP; i += P::kThreads) {
        sXb[(0 * 3 + 0) * NP + i] = sY[i];
        sXb[(0 * 3 + 1) * NP + i] = make_float4(i == k && c == 0 ? 1.f : 0.f, i == k && c == 1 ? 1.f : 0.f, i == k && c == 2 ? 1.f : 0.f, 0.f);
        sXb[(0 * 3 + 2) * NP + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      {
        const float e0 = __ldg(wpack + lane), e1 = __ldg(wpack + 32 + lane), eb = __ldg(wpack + 64 + lane);
        for (int i = warp; i < NP; i += NW) {
          const float f0 = (2 * i < NP) ? tcond : bet;       // the reference's cat/reshape feature layout
          const float f1 = (2 * i + 1 < NP) ? tcond : bet;   //   (egnn_temp_conditioned.py:63-78)
          sH[i * H + lane] = fmaf(e0, f0, fmaf(e1, f1, eb));
          sH[NH + i * H + lane] = 0.f;
          sH[2 * NH + i * H + lane] = 0.f;
        }
      }
      __syncthreads();
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const float *__restrict__ Wl = wpack + pk::kHeader + l * pk::kLayer;
        const float4 *sX = sXb + cur * 3 * NP;
        float4 *sXn = sXb + (cur ^ 1) * 3 * NP;
        {  // p_i = A h_i + b1, q_i = B h_i on the three coefficients
          float wA[H], wB[H];
          load_row(wA, Wl + pk::A_f, lane);
          load_row(wB, Wl + pk::B_f, lane);
          const float b1 = __ldg(Wl + pk::b1 + lane);
          for (int i = warp; i < NP; i += NW) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {
              sP[p * NH + i * H + lane] = (p == 0 ? b1 : 0.f) + dot32(wA, sH + p * NH + i * H);
              sQ[p * NH + i * H + lane] = dot32(wB, sH + p * NH + i * H);
            }
          }
        }
        __syncthreads();
        {
          float w2[H], wc1[H];
          load_row(w2, Wl + pk::W2_f, lane);
          load_row(wc1, Wl + pk::Wc1_f, lane);
          const EdgeScal sc = load_edge_scal(Wl, lane);
          for (int i = warp; i < NP; i += NW) {
            const J pi = jmake(sP[i * H + lane], sP[NH + i * H + lane], sP[2 * NH + i * H + lane]);
            const float4 xiv = sX[i], xia = sX[NP + i], xib = sX[2 * NP + i], yi = sY[i];
            J agg = jmake(0.f, 0.f, 0.f), dx0 = agg, dx1 = agg, dx2 = agg;
#pragma unroll 1
            for (int j = 0; j < NP; ++j) {
              if (j == i) continue;
              const float4 xjv = sX[j], xja = sX[NP + j], xjb = sX[2 * NP + j], yj = sY[j];
              const J d0 = jmake(xiv.x - xjv.x, xia.x - xja.x, xib.x - xjb.x);
              const J d1 = jmake(xiv.y - xjv.y, xia.y - xja.y, xib.y - xjb.y);
              const J d2 = jmake(xiv.z - xjv.z, xia.z - xja.z, xib.z - xjb.z);
              const J r2 = jadd(jmul(d0, d0), jadd(jmul(d1, d1), jmul(d2, d2)));
              // edge_attr = |y_i - y_j|^2 of the INPUT coordinates (egnn_temp_conditioned.py:79), with y_k,c -> y_k,c + eps
              const float e0 = yi.x - yj.x, e1 = yi.y - yj.y, e2 = yi.z - yj.z;
              const float sgn = (i == k ? 1.f : 0.f) - (j == k ? 1.f : 0.f);
              const float ec = c == 0 ? e0 : (c == 1 ? e1 : e2);
              const J ea = jmake(e0 * e0 + e1 * e1 + e2 * e2, 2.0f * sgn * ec, sgn * sgn);
              // ---- edge MLP
              J z1 = jadd(pi, jmake(sQ[j * H + lane], sQ[NH + j * H + lane], sQ[2 * NH + j * H + lane]));
              z1 = jfma(sc.c1, r2, jfma(sc.d1, ea, z1));
              const J m1 = jsilu(z1);
              st[lane] = m1.v; st[H + lane] = m1.a; st[2 * H + lane] = m1.b;
              __syncwarp();
              J z2 = dot3(w2, st, H);
              z2.v += sc.b2;
              const J m = jsilu(z2);
              J q = jwarp_sum(jscale(sc.wa, m));
              q.v += sc.ba;
              const J ms = jmul(m, jsigmoid(q));
              agg = jadd(agg, ms);
              st[3 * H + lane] = ms.v; st[4 * H + lane] = ms.a; st[5 * H + lane] = ms.b;
              __syncwarp();
              // ---- coordinate MLP
              J zc = dot3(wc1, st + 3 * H, H);
              zc.v += sc.bc1;
              const J u = jwarp_sum(jscale(sc.wc2, jsilu(zc)));
              const J phi = jscale(rng, jtanh(u));
              J nrm = jsqrt(jmake(r2.v + kNormEps, r2.a, r2.b));
              nrm.v += 1.0f;
              const J gsc = jmul(jrcp(nrm), phi);
              dx0 = jadd(dx0, jmul(d0, gsc)); dx1 = jadd(dx1, jmul(d1, gsc)); dx2 = jadd(dx2, jmul(d2, gsc));
            }
            if (lane == 0) {
              sXn[i] = make_float4(xiv.x + dx0.v, xiv.y + dx1.v, xiv.z + dx2.v, 0.f);
              sXn[NP + i] = make_float4(xia.x + dx0.a, xia.y + dx1.a, xia.z + dx2.a, 0.f);
              sXn[2 * NP + i] = make_float4(xib.x + dx0.b, xib.y + dx1.b, xib.z + dx2.b, 0.f);
            }
            sAgg[i * H + lane] = agg.v; sAgg[NH + i * H + lane] = agg.a; sAgg[2 * NH + i * H + lane] = agg.b;
          }
        }
        __syncthreads();
        if (l < L - 1) {  // node update (dead for the output in the last layer)
          float wa_[H], wb_[H];
          load_row(wa_, Wl + pk::W3h_f, lane);
          load_row(wb_, Wl + pk::W3a_f, lane);
          const float b3 = __ldg(Wl + pk::b3 + lane), b4 = __ldg(Wl + pk::b4 + lane);
          for (int i = warp; i < NP; i += NW) {
            J z3;
            z3.v = b3 + dot32(wa_, sH + i * H) + dot32(wb_, sAgg + i * H);
            z3.a = dot32(wa_, sH + NH + i * H) + dot32(wb_, sAgg + NH + i * H);
            z3.b = dot32(wa_, sH + 2 * NH + i * H) + dot32(wb_, sAgg + 2 * NH + i * H);
            const J a3 = jsilu(z3);
            __syncwarp();   // every lane has read row i of sAgg: reuse it as the input of the second linear
            sAgg[i * H + lane] = a3.v; sAgg[NH + i * H + lane] = a3.a; sAgg[2 * NH + i * H + lane] = a3.b;
          }
          __syncwarp();
          load_row(wa_, Wl + pk::W4_f, lane);
          for (int i = warp; i < NP; i += NW) {
            sH[i * H + lane] += b4 + dot32(wa_, sAgg + i * H);
            sH[NH + i * H + lane] += dot32(wa_, sAgg + NH + i * H);
            sH[2 * NH + i * H + lane] += dot32(wa_, sAgg + 2 * NH + i * H);
          }
        }
        __syncthreads();
        cur ^= 1;
      }
      // ---- eps^2 coefficient of u = sum_i <v_i, y_i>, v = (x_L - y) - mean(x_L - y), along direction (k, c):
      //      u_b = sum_i <w_b,i, y_i> - <mean(w_b), sum_i y_i> + (w_a,k,c - mean_j w_a,j,c),  w_a = x_L,a - e,  w_b = x_L,b
      if (warp == 0) {
        const float4 *sX = sXb + cur * 3 * NP;
        float s_wy = 0.f, s_wb0 = 0.f, s_wb1 = 0.f, s_wb2 = 0.f, s_y0 = 0.f, s_y1 = 0.f, s_y2 = 0.f, s_wa = 0.f;
        for (int i = lane; i < NP; i += 32) {
          const float4 wb = sX[2 * NP + i], wa = sX[NP + i], y = sY[i];
          s_wy += wb.x * y.x + wb.y * y.y + wb.z * y.z;
          s_wb0 += wb.x; s_wb1 += wb.y; s_wb2 += wb.z;
          s_y0 += y.x; s_y1 += y.y; s_y2 += y.z;
          s_wa += (c == 0 ? wa.x : (c == 1 ? wa.y : wa.z));
        }
        s_wy = warp_sum(s_wy); s_wb0 = warp_sum(s_wb0); s_wb1 = warp_sum(s_wb1); s_wb2 = warp_sum(s_wb2);
        s_y0 = warp_sum(s_y0); s_y1 = warp_sum(s_y1); s_y2 = warp_sum(s_y2); s_wa = warp_sum(s_wa);
        const float4 wak = sX[NP + k];
        const float wakc = (c == 0 ? wak.x : (c == 1 ? wak.y : wak.z)) - 1.0f;
        const float mean_wa = (s_wa - 1.0f) * (1.0f / NP);
        const float ub = s_wy - (s_wb0 * s_y0 + s_wb1 * s_y1 + s_wb2 * s_y2) * (1.0f / NP) + (wakc - mean_wa);
        lap_acc += 2.0f * ub;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float c_s = 1.0f / (1.0f + h);
      const float c_out = sqrtf(h) * c_in;
      lap_out[b] = (1.0f - c_s) * (3.0f * NP) / h - c_out * c_in / h * lap_acc;
    }
  }
}

template <int NP, int NW>
static int launch(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *lap, cudaStream_t s) {
  using P = LPlan<NP, NW>;
  auto kern = egnn_laplacian_kernel<NP, NW, 3>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t per_sm = (NP == 13) ? 2 : 1;
  const int64_t grid = B < sms * per_sm ? B : sms * per_sm;
  kern<<<(unsigned)grid, P::kThreads, P::kBytes, s>>>(w, ht, x, beta, B, lap);
  PITA_CHECK_LAUNCH("egnn_laplacian_kernel");
  return PITA_OK;
}

int launch_laplacian(int n, const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *lap, cudaStream_t s) {
  if (n == 13) return launch<13, 13>(w, ht, x, beta, B, lap, s);
  return launch<55, 11>(w, ht, x, beta, B, lap, s);
}

}  // namespace lap
}  // namespace pita
