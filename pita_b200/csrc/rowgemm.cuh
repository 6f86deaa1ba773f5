// Row-batched tensor-core engine for the EGNN kernels (sm_100a, tcgen05 + TMEM).
//
// Every dense product of the E_GCL block (edge MLP, coordinate MLP, node MLP, and their tangent /
// cotangent versions) is a [rows x 32] x [32 x 32] GEMM whose rows are independent.  A "team" of 128
// threads owns one 128-row tile:  thread = row.  Row r of a team stands for (particle p, node i) with
// r = p * n + i, so one tile carries PB = floor(128 / n) whole particles (9 for LJ-13, 2 for LJ-55), and the
// fully-connected graph is walked as n-1 "sender slots" u:  in slot u node i receives from j = (i+1+u) mod n —
// a cyclic shift, hence a bijection i -> j inside every particle.  Per slot each thread
//   builds its 32-channel operand row in registers  ->  stores it K-major into the team's 128B-swizzled A tile
//   ->  one elected thread issues tcgen05.mma (kind::tf32, weights as the K-major B operand, accumulator in TMEM)
//   ->  every thread reads its accumulator row back with tcgen05.ld and applies the element-wise stage.
// Receiver-side state (P_i, accumulators, tangents) therefore never leaves the thread / its TMEM lane, sender-side
// state is gathered from padded shared-memory rows, and nothing but x, h(t), beta is read from HBM.
// SPLIT = true evaluates every product as 3xTF32 (hi/lo split of both operands: fp32-accurate); SPLIT = false is
// plain TF32 (the precision torch uses for these matmuls under set_float32_matmul_precision("high"),
// energytemp_module.py:39).
#pragma once
#include "egnn_common.cuh"
#include "umma.cuh"

namespace pita {
namespace rg {

constexpr int kRows = 128;  // rows (threads) per team
// Tangent operand rows (TS path).  false (default): the row is rounded to TF32 once and multiplied with the EXACTLY split
// weights (hi + lo): 8 MMAs per product.  true: hi + lo parts of the row as well (full 3xTF32, 12 MMAs).  The error of plain
// TF32 on this network comes from rounding the WEIGHTS (the same rounded matrix hits every edge and direction: a coherent
// bias, 3e-2 on the LJ-55 divergence), not from rounding the rows (independent per row, averages out over the 3n x (n-1)
// terms of the trace): measured divergence error 1.2e-5 (LJ-55) / 3.7e-6 (LJ-13) with hi-only rows vs 5.4e-6 / 8e-7 with
// hi + lo rows (profiles/r1t_err_by_mode.jsonl), 6 % faster.  Primal rows (store_row) always carry hi + lo.
constexpr bool kTangentLo = false;
constexpr int kWSlots = 5;  // weight-tile slots shared by the CTA's teams
constexpr int kNumVec = 10; // per-layer 32-float vectors: c1 d1 b1 b2 wa ba bc1 wc2 b3 b4
enum Vec { vC1 = 0, vD1 = 1, vB1 = 2, vB2 = 3, vWA = 4, vBA = 5, vBC1 = 6, vWC2 = 7, vB3 = 8, vB4 = 9 };

template <bool SPLIT>
struct Bytes {
  static constexpr int kA = (SPLIT ? 2 : 1) * 128 * 128;  // A tile: 128 rows x 128 B (hi [+ lo])
  static constexpr int kW = (SPLIT ? 2 : 1) * 32 * 128;   // W tile:  32 rows x 128 B (hi [+ lo])
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- weight tiles -----------------------------------------------------------------------------
// A weight matrix lives in the packed buffer as 32 rows of 32 contiguous floats; row = N index of the B operand.
struct WeightSrc {
  const float *p[kWSlots];
  int count;
};

// All threads of the CTA call this between two __syncthreads() (the caller provides them).
template <bool SPLIT>
__device__ __forceinline__ void load_weight_tiles(float *wsm, const WeightSrc &src, int tid, int nthreads) {
  for (int item = tid; item < src.count * 32; item += nthreads) {
    const int m = item >> 5, row = item & 31;
    const float4 *g = reinterpret_cast<const float4 *>(src.p[m] + row * 32);
    float v[32];
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 q = __ldg(g + k4);
      v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
    }
    float *hi = wsm + m * (Bytes<SPLIT>::kW / 4);
    if (SPLIT) {
      float h[32], l[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) umma::split_tf32(v[k], h[k], l[k]);
      umma::store_row_sw128(hi, row, h);
      umma::store_row_sw128(hi + 1024, row, l);
    } else {
      umma::store_row_sw128(hi, row, v);
    }
  }
  umma::fence_proxy_async_smem();
}

// ---- one team ---------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t uniform32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

template <bool SPLIT>
struct Team {
  float *a_hi;        // A tile (hi); lo follows at +4096 floats when SPLIT
  uint32_t a_addr;    // shared-space byte address of the A tile       (warp-uniform)
  uint32_t w_addr;    // shared-space byte address of weight slot 0     (warp-uniform)
  uint32_t mbar_addr; // shared-space byte address of the team mbarrier (warp-uniform)
  uint32_t tmem_col;  // TMEM address (lane 0) of column 0 of this team (warp-uniform)
  uint32_t phase;
  uint32_t tmem;      // tmem_col with the lane offset of this warp folded in
  int bar_id;         // named barrier of the team
  int tt;             // thread index in the team == tile row
  bool issuer;        // this warp issues the team's MMAs (warp 0 of the team)

  __device__ __forceinline__ void sync() const { named_sync(bar_id, kRows); }

  __device__ __forceinline__ void store_row(const float (&v)[32]) const {
    if (SPLIT) {
      float h[32], l[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) umma::split_tf32(v[k], h[k], l[k]);
      umma::store_row_sw128(a_hi, tt, h);
      umma::store_row_sw128(a_hi + 4096, tt, l);
    } else {
      umma::store_row_sw128(a_hi, tt, v);
    }
  }

  // D[128 x 32] (TMEM columns [32*dslot, 32*dslot+32)) (+)= A tile x W[wslot]^T.  Called by the elected lane of the
  // issuing warp; every operand is derived from warp-uniform values so that the descriptors stay in uniform registers.
  __device__ __forceinline__ void mma(int dslot, int wslot, bool accumulate) const {
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
    const uint32_t d = tmem_col + 32u * dslot;
    const uint32_t wa = w_addr + (uint32_t)wslot * (uint32_t)Bytes<SPLIT>::kW;
    const uint64_t dA = umma::make_desc_sw128_kmajor(a_addr);
    const uint64_t dB = umma::make_desc_sw128_kmajor(wa);
    if (SPLIT) {
      // 3xTF32: the two correction products first, so that their partial sums are rounded (the tensor core truncates
      // the fp32 accumulator) at their own small magnitude and only the four hi*hi steps round at full magnitude.
      const uint64_t dAl = umma::make_desc_sw128_kmajor(a_addr + 16384u);
      const uint64_t dBl = umma::make_desc_sw128_kmajor(wa + 4096u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(d, dAl + 2 * k, dB + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(d, dA + 2 * k, dBl + 2 * k, idesc, 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(d, dA + 2 * k, dB + 2 * k, idesc, 1u);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(d, dA + 2 * k, dB + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
  }

  // Round trip: every thread has stored its operand row; `issue` (one elected lane of the team's first warp) issues
  // the MMAs; on return the accumulators are complete and visible to tcgen05.ld of every thread of the team.
  template <class F>
  __device__ __forceinline__ void round_trip(F issue) {
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    sync();
    if (issuer) {
      if (elect_one()) {
        umma::fence_after_thread_sync();
        issue();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
      }
      __syncwarp();
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "RT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra RT_DONE;\n\t"
        "bra RT_WAIT;\n\t"
        "RT_DONE:\n\t}\n" ::"r"(mbar_addr),
        "r"(phase)
        : "memory");
    phase ^= 1u;
    umma::fence_after_thread_sync();
  }

  // ---- TS form: the A operand row goes into the thread's own TMEM lane (two 32-column slots: hi, lo) instead of the
  // shared-memory tile.  No STS, no generic->async proxy fence; measured 541 vs 990 cycles per 3xTF32 round trip
  // (profiles/ubench/roundtrip.cu, bit-identical results).  The slots must not be a live accumulator.
  __device__ __forceinline__ void store_row_tmem(int slot_hi, int slot_lo, const float (&v)[32]) const {
    if (SPLIT && kTangentLo) {
      float h[32], l[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) umma::split_tf32_tangent(v[k], h[k], l[k]);  // (only tangent rows take the TS path)
      umma::tmem_st_32x32(tmem + 32u * slot_hi, h);
      umma::tmem_st_32x32(tmem + 32u * slot_lo, l);
    } else if (SPLIT) {
      float h[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) h[k] = umma::round_tf32(v[k]);
      umma::tmem_st_32x32(tmem + 32u * slot_hi, h);
    } else {
      umma::tmem_st_32x32(tmem + 32u * slot_hi, v);
    }
  }
  __device__ __forceinline__ void mma_ts(int dslot, int aslot_hi, int aslot_lo, int wslot, bool accumulate) const {
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
    const uint32_t d = tmem_col + 32u * dslot, ah = tmem_col + 32u * aslot_hi, al = tmem_col + 32u * aslot_lo;
    const uint32_t wa = w_addr + (uint32_t)wslot * (uint32_t)Bytes<SPLIT>::kW;
    const uint64_t dB = umma::make_desc_sw128_kmajor(wa);
    if (SPLIT) {
      const uint64_t dBl = umma::make_desc_sw128_kmajor(wa + 4096u);
      if (kTangentLo) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, al + 8u * k, dB + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dBl + 2 * k, idesc, (kTangentLo || accumulate || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, 1u);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
  }
  template <class F>
  __device__ __forceinline__ void round_trip_ts(F issue) {
    umma::fence_before_thread_sync();
    sync();
    if (issuer) {
      if (elect_one()) {
        umma::fence_after_thread_sync();
        issue();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
      }
      __syncwarp();
    }
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "RTS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra RTS_DONE;\n\t"
        "bra RTS_WAIT;\n\t"
        "RTS_DONE:\n\t}\n" ::"r"(mbar_addr),
        "r"(phase)
        : "memory");
    phase ^= 1u;
    umma::fence_after_thread_sync();
  }

  __device__ __forceinline__ void ld(int slot, float (&v)[32]) const { umma::tmem_ld_32x32(tmem + 32u * slot, v); }
  __device__ __forceinline__ void st(int slot, const float (&v)[32]) const { umma::tmem_st_32x32(tmem + 32u * slot, v); }
};

// ---- gatherable row vectors: [128][32] floats, 16-byte chunk c of row r stored at chunk (c ^ (r & 7)) so that a warp
// reading 32 different rows (the sender gather) touches every bank once per quarter-warp, without padding.
__device__ __forceinline__ float4 qrow_ld4(const float *base, int row, int k4) {
  return lds4(base + row * 32 + ((k4 ^ (row & 7)) << 2));
}
__device__ __forceinline__ void qrow_store(float *base, int row, const float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4)
    sts4(base + row * 32 + ((k4 ^ (row & 7)) << 2), make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]));
}
__device__ __forceinline__ void qrow_load(const float *base, int row, float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = qrow_ld4(base, row, k4);
    v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
  }
}

// ---- row-vector helpers (thread-private 32-float vectors) ----------------------------------------
__device__ __forceinline__ void load_vec_smem(const float *p, float (&v)[32]) {  // p 16-byte aligned
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = lds4(p + 4 * k4);
    v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
  }
}
__device__ __forceinline__ void store_vec_smem(float *p, const float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) sts4(p + 4 * k4, make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]));
}
__device__ __forceinline__ void load_vec_global(const float *p, float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = *reinterpret_cast<const float4 *>(p + 4 * k4);
    v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
  }
}
__device__ __forceinline__ void store_vec_global(float *p, const float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4)
    *reinterpret_cast<float4 *>(p + 4 * k4) = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
}

// Coalesced own-row vectors in global memory: vector element block k4 of row r lives at base[(k4 * 128 + r) * 4 .. +4),
// so one warp-wide 128-bit access covers 512 contiguous bytes (4 lines) instead of 32 different lines.  L2-only
// (ld/st.global.cg): every value is read back by the thread that wrote it, there is nothing for L1 to share.
__device__ __forceinline__ void load_vec_global_co(const float *base, int r, float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 q = __ldcg(reinterpret_cast<const float4 *>(base + (k4 * kRows + r) * 4));
    v[4 * k4] = q.x; v[4 * k4 + 1] = q.y; v[4 * k4 + 2] = q.z; v[4 * k4 + 3] = q.w;
  }
}
__device__ __forceinline__ void store_vec_global_co(float *base, int r, const float (&v)[32]) {
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4)
    __stcg(reinterpret_cast<float4 *>(base + (k4 * kRows + r) * 4), make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]));
}

// geometry of one edge from the thread's own and the sender's coordinates (layer input and network input)
struct Geo {
  float d0, d1, d2;  // x_i - x_j
  float r2, nrm, inv, ea;
};
__device__ __forceinline__ Geo edge_geo4(const float4 xi, const float4 xj, const float4 x0i, const float4 x0j) {
  Geo g;
  g.d0 = xi.x - xj.x; g.d1 = xi.y - xj.y; g.d2 = xi.z - xj.z;
  g.r2 = g.d0 * g.d0 + g.d1 * g.d1 + g.d2 * g.d2;
  g.nrm = sqrtf(g.r2 + kNormEps);
  g.inv = 1.0f / (g.nrm + 1.0f);
  const float e0 = x0i.x - x0j.x, e1 = x0i.y - x0j.y, e2 = x0i.z - x0j.z;
  g.ea = e0 * e0 + e1 * e1 + e2 * e2;
  return g;
}

}  // namespace rg
}  // namespace pita
