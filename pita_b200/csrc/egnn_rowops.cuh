// Element-wise stages of one E_GCL edge on the row engine (rowgemm.cuh), shared by egnn_rows.cu (forward, energy, v1 score /
// divergence) and egnn_tri.cu (round-2 score / divergence).  Reference: egnn_temp_conditioned.py:232-318.
#pragma once
#include "rowgemm.cuh"

namespace pita {
namespace rg {

// ---- element-wise stages of one edge ----------------------------------------------------------------
// stage 1: z1 = P_i + Q_j + c1 r2 + d1 ea  ->  row = silu(z1), f1 = silu'(z1)
template <bool KEEP, bool HAS_Q = true>
__device__ __forceinline__ void stage1(float (&row)[32], float (&f1)[32], const float *qbase, int qrow, const float *vec, float r2,
                                       float ea) {
  vec = opq(vec);
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = HAS_Q ? qrow_ld4(qbase, qrow, k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 c = lds4c(vec + vC1 * 32 + 4 * k4);
    const float4 d = lds4c(vec + vD1 * 32 + 4 * k4);
    const float qq[4] = {q.x, q.y, q.z, q.w}, cc[4] = {c.x, c.y, c.z, c.w}, dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      const float z = row[k] + qq[e] + cc[e] * r2 + dd[e] * ea;
      float a, f;
      silu_both(z, a, f);
      row[k] = a;
      if (KEEP) f1[k] = f;
    }
  }
}

// stage 2: z2 = acc + b2 -> m = silu(z2), f2; attention gate; row = m * att.  Returns att.
template <bool KEEP>
__device__ __forceinline__ float stage2(float (&row)[32], float (&m)[32], float (&f2)[32], const float *vec) {
  vec = opq(vec);
  float dot = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 b = lds4c(vec + vB2 * 32 + 4 * k4);
    const float4 w = lds4c(vec + vWA * 32 + 4 * k4);
    const float bb[4] = {b.x, b.y, b.z, b.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      float a, f;
      silu_both(row[k] + bb[e], a, f);
      row[k] = a;
      if (KEEP) { m[k] = a; f2[k] = f; }
      dot = fmaf(ww[e], a, dot);
    }
  }
  const float att = sigmoidf_fast(dot + lds1(vec + vBA * 32));
#pragma unroll
  for (int k = 0; k < 32; ++k) row[k] *= att;
  return att;
}

// stage 3: zc = acc + bc1 -> u = <wc2, silu(zc)>, th = tanh(u); fc = silu'(zc)
template <bool KEEP>
__device__ __forceinline__ float stage3(const float (&acc)[32], float (&fc)[32], const float *vec) {
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
      if (KEEP) fc[k] = f;
      u = fmaf(ww[e], a, u);
    }
  }
  return tanhf(u);
}

__device__ __forceinline__ void add_vec(float (&row)[32], const float *vec32) {
  vec32 = opq(vec32);
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 b = lds4c(vec32 + 4 * k4);
    row[4 * k4] += b.x; row[4 * k4 + 1] += b.y; row[4 * k4 + 2] += b.z; row[4 * k4 + 3] += b.w;
  }
}

// Node embedding h^0 of the thread's node (egnn_temp_conditioned.py:63-78: node k sees (f[2k], f[2k+1]) of
// f = [t]*n ++ [beta]*n).
template <int NP>
__device__ __forceinline__ void embed(float (&h)[32], const float *sEmb, int i, float tcond, float beta) {
  const float f0 = (2 * i < NP) ? tcond : beta;
  const float f1 = (2 * i + 1 < NP) ? tcond : beta;
#pragma unroll
  for (int k = 0; k < 32; ++k) h[k] = fmaf(sEmb[k], f0, fmaf(sEmb[32 + k], f1, sEmb[64 + k]));
}

// acc = W2 dz1  ->  d(ms) in place:  dm = f2*acc, ds = att(1-att) <wa, dm>, dms = dm*att + m*ds
__device__ __forceinline__ void tangent_mid(float (&row)[32], const float (&m)[32], const float (&f2)[32], float att, const float *vec) {
  vec = opq(vec);
  float dsd = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 w = lds4c(vec + vWA * 32 + 4 * k4);
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      row[k] *= f2[k];
      dsd = fmaf(ww[e], row[k], dsd);
    }
  }
  const float ds = att * (1.0f - att) * dsd;
#pragma unroll
  for (int k = 0; k < 32; ++k) row[k] = fmaf(m[k], ds, row[k] * att);
}

// du = < wc2 * fc, Wc1 dms >
__device__ __forceinline__ float tangent_du(const float (&acc)[32], const float (&fc)[32], const float *vec) {
  vec = opq(vec);
  float du = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 w = lds4c(vec + vWC2 * 32 + 4 * k4);
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) du = fmaf(ww[e] * fc[4 * k4 + e], acc[4 * k4 + e], du);
  }
  return du;
}

// row = f1 * (row + c1 dr2 + d1 dea)
__device__ __forceinline__ void tangent_in(float (&row)[32], const float (&f1)[32], const float *vec, float dr2, float dea) {
  vec = opq(vec);
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 c = lds4c(vec + vC1 * 32 + 4 * k4);
    const float4 d = lds4c(vec + vD1 * 32 + 4 * k4);
    const float cc[4] = {c.x, c.y, c.z, c.w}, dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * k4 + e;
      row[k] = f1[k] * (row[k] + cc[e] * dr2 + dd[e] * dea);
    }
  }
}

// z1 (without the geometric terms) of a layer-0 edge from the class tables:  P0_i + Q0_j
__device__ __forceinline__ void z1_layer0(float (&row)[32], const float *sCls, float f0i, float f1i, float f0j, float f1j) {
  sCls = opq(sCls);
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 a0 = lds4c(sCls + 4 * k4), a1 = lds4c(sCls + 32 + 4 * k4), ab = lds4c(sCls + 64 + 4 * k4);
    const float4 b0 = lds4c(sCls + 96 + 4 * k4), b1 = lds4c(sCls + 128 + 4 * k4), bb = lds4c(sCls + 160 + 4 * k4);
    row[4 * k4 + 0] = fmaf(f0i, a0.x, fmaf(f1i, a1.x, ab.x)) + fmaf(f0j, b0.x, fmaf(f1j, b1.x, bb.x));
    row[4 * k4 + 1] = fmaf(f0i, a0.y, fmaf(f1i, a1.y, ab.y)) + fmaf(f0j, b0.y, fmaf(f1j, b1.y, bb.y));
    row[4 * k4 + 2] = fmaf(f0i, a0.z, fmaf(f1i, a1.z, ab.z)) + fmaf(f0j, b0.z, fmaf(f1j, b1.z, bb.z));
    row[4 * k4 + 3] = fmaf(f0i, a0.w, fmaf(f1i, a1.w, ab.w)) + fmaf(f0j, b0.w, fmaf(f1j, b1.w, bb.w));
  }
}

template <int NP>
__device__ __forceinline__ void node_feats(int node, float tcond, float beta, float &f0, float &f1) {
  f0 = (2 * node < NP) ? tcond : beta;
  f1 = (2 * node + 1 < NP) ? tcond : beta;
}

__device__ __forceinline__ float comp(const float4 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

}  // namespace rg
}  // namespace pita
