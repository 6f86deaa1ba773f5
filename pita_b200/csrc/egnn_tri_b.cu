// Phase B of the round-2 score / divergence engine (see egnn_tri.cuh): the bilinear pairing on the middle layer's edges.
//
// Thread = one middle-layer edge (i, j) (receiver i, sender j); a team of 128 threads = the edges of RPT receivers; the
// thread keeps its edge's k-independent quantities in registers (f1 = silu'(z1), f2 = silu'(z2), vt = T^T v, gate, geometry:
// recomputed per tile with four 3xTF32 products, no edge cache in memory) and loops over the tangent/output node k:
//     u    = f1 * (alpha_i PA_ik + alpha_j PB_jk + beta c1)        -> own TMEM lane (operand row, TF32)
//     D    = W2 u                                                  -> ONE tcgen05.mma group per (edge, k), accumulator in TMEM
//     term = s <f2*D, gamma_ki> + s(1-s) <m, gamma_ki> <f2*D, wa>  (+ the coordinate-branch terms: dot products with vt)
// <m, gamma_ki> for all k comes from one extra MMA per tile (N = 128 table columns).  The pair tables arrive in k-chunks by
// cp.async.bulk (TMA bulk copy) into a two-stage ring filled by a producer warp; a second helper warp issues every MMA, so
// the compute warps never wait for each other, only for data.  Two products are in flight per team (two operand / two
// accumulator slots).  Items per row: n generic (k = 0..n-1; rows whose receiver is k idle), one "S" item (k = j: the
// full-rank part of the sender's tangent) and three "R" items (k = i, one per direction).
// Algebra and names: oracle/egnn_bilinear.py::phase_b_items.  Reference being differentiated: egnn_temp_conditioned.py:265-356.
#include "egnn_tri.cuh"

namespace pita {
namespace tri {

using namespace rg;

constexpr int kComputeThreads = 256;
// One helper warpgroup (warps 0..3: MMA issuer, copy producer, two idle warps) + 8 compute warps (2 teams = warpgroups 1 and 2:
// the scheduler favours the higher warp ids, so the compute warps win the issue slot over a polling helper).  Registers are
// allocated per warpgroup on sm_100: the helper group hands its share to the compute groups (setmaxnreg).
constexpr int kThreadsB = 384;
constexpr int kRegsCompute = 224, kRegsHelper = 56;

template <int NP>
struct CfgB;
template <>
struct CfgB<55> {
  static constexpr int RS = 64, RPT = 2, KC = 3, NST = 3;  // k per chunk, ring stages
};
template <>
struct CfgB<13> {
  static constexpr int RS = 16, RPT = 8, KC = 5, NST = 3;
};

template <int NP>
struct SmemB {
  using C = CfgB<NP>;
  static constexpr int NSLOT = 2 * C::RPT;                       // receiver slots per CTA tile
  static constexpr int NG = (NP + NSLOT - 1) / NSLOT;            // CTA tiles per particle
  static constexpr int NCH = (NP + C::KC - 1) / C::KC;           // k-chunks per tile
  static constexpr int kTSChunk = C::KC * NP * kTS;              // floats
  static constexpr int kTRChunk = NSLOT * C::KC * kTR;           // floats
  static constexpr int kChunk = kTSChunk + kTRChunk;             // floats per ring stage
  static constexpr size_t oW = 0;                                // 4 weight tiles (hi + lo): W2, Wc1, Wc1^T, W2^T of layer 1
  static constexpr size_t oG = oW + 4 * 8192;                    // [2 teams] 128 x 128 B staging of the gamma rows (B operand)
  static constexpr size_t oCh = oG + 2 * 16384;                  // [NST stages][kChunk]
  static constexpr size_t oF = oCh + C::NST * (size_t)kChunk * 4;
  static constexpr int fVec = 0;                                 // [kNumVec][32] layer-1 vectors
  static constexpr int fY = fVec + kNumVec * 32;                 // [2 teams][NP] float4
  static constexpr int fRed = fY + 2 * NP * 4;                   // [2 teams][4]
  static constexpr int fBar = fRed + 8;                          // mbarriers (8 B each): 16 of them, then the tmem slot
  static constexpr int kFloats = fBar + 2 * 16 + 4;
  static constexpr size_t kBytes = 1024 + oF + (size_t)kFloats * 4;
};

// ---- mbarrier / bulk-copy plumbing
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "MW_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra MW_DONE;\n\t"
      "bra MW_WAIT;\n\t"
      "MW_DONE:\n\t}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// wait of the helper warps (MMA issuer, copy producer): with a suspend-time hint, so that the polling lane sleeps in hardware
// instead of taking issue slots from the compute warps of its scheduler (ncu r2c: 129 M polls per launch without it)
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "MS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra MS_DONE;\n\t"
      "bra MS_WAIT;\n\t"
      "MS_DONE:\n\t}\n" ::"r"(addr),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}

// TMEM columns of one team (256): two operand / accumulator stages and the <m, gamma> table
constexpr uint32_t cA0 = 0, cD0 = 32, cA1 = 64, cD1 = 96, cG = 128;

struct TeamB {
  uint32_t tmem_col;  // column 0 of the team (lane 0)
  uint32_t tmem;      // with this warp's lane offset
  uint32_t ops_addr;  // ops_ready[2]  (count 128)
  uint32_t dr_addr;   // d_ready[2]    (count 1, tcgen05.commit)
  uint32_t q;         // running request counter
  int bar_id;
  __device__ __forceinline__ void st(uint32_t col, const float (&v)[32]) const { umma::tmem_st_32x32(tmem + col, v); }
  __device__ __forceinline__ void ld(uint32_t col, float (&v)[32]) const { umma::tmem_ld_32x32(tmem + col, v); }
  // hand the operand row(s) written so far to the MMA warp; returns the request id
  __device__ __forceinline__ uint32_t post() {
    umma::fence_before_thread_sync();
    mbar_arrive(ops_addr + 8u * (q & 1u));
    return q++;
  }
  __device__ __forceinline__ void wait(uint32_t rq) const {
    mbar_wait_addr(dr_addr + 8u * (rq & 1u), (rq >> 1) & 1u);
    umma::fence_after_thread_sync();
  }
  __device__ __forceinline__ void sync() const { named_sync(bar_id, kRows); }
};

// ---- MMA issue helpers (one elected lane of the MMA warp)
__device__ __forceinline__ void issue_full(uint32_t tcol, uint32_t w_addr, int wslot) {  // D0 = (A0 hi, A1 lo) x W^T, 3xTF32
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
  const uint32_t d = tcol + cD0, ah = tcol + cA0, al = tcol + cA1;
  const uint32_t wa = w_addr + (uint32_t)wslot * 8192u;
  const uint64_t dB = umma::make_desc_sw128_kmajor(wa), dBl = umma::make_desc_sw128_kmajor(wa + 4096u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, al + 8u * k, dB + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dBl + 2 * k, idesc, 1u);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, ah + 8u * k, dB + 2 * k, idesc, 1u);
}
__device__ __forceinline__ void issue_item(uint32_t tcol, uint32_t w_addr, uint32_t stage, bool with_lo) {
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
  const uint32_t d = tcol + (stage ? cD1 : cD0), a = tcol + (stage ? cA1 : cA0);
  const uint64_t dB = umma::make_desc_sw128_kmajor(w_addr), dBl = umma::make_desc_sw128_kmajor(w_addr + 4096u);
  if (with_lo) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, a + 8u * k, dBl + 2 * k, idesc, k > 0 ? 1u : 0u);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(d, a + 8u * k, dB + 2 * k, idesc, (with_lo || k > 0) ? 1u : 0u);
}
__device__ __forceinline__ void issue_gtable(uint32_t tcol, uint32_t g_addr) {  // G[128 x 128] = m (in cD1) x gamma rows^T
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, 128);
  const uint64_t dB = umma::make_desc_sw128_kmajor(g_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ts(tcol + cG, tcol + cD1 + 8u * k, dB + 2 * k, idesc, k > 0 ? 1u : 0u);
}
__device__ __forceinline__ void commit_to(uint32_t mbar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}

__device__ __forceinline__ float4 ldg4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
// Shared-memory loads of the hot loops: NOT the `asm volatile` lds4 of common.cuh — volatile asm statements keep their
// program order, which serialises every load behind the arithmetic that consumes the previous one (33 exposed LDS
// latencies per item: the r2c/r2d profiles).  A plain asm with the address as its only input can be scheduled freely; the
// `tok` operand ties it to the barrier wait that made the data visible, so it cannot be hoisted above that wait.
__device__ __forceinline__ float4 lds4s(const float *p, uint32_t tok) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+0];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p)) + tok));
  return v;
}

// packed fp32 pairs (FFMA2 / FMUL2: two channels per issue slot — the element-wise work around every product is what bounds
// this kernel, not the tensor pipe)
struct pf2 {
  unsigned long long v;
};
__device__ __forceinline__ pf2 pk(float lo, float hi) {
  pf2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(pf2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ pf2 fma2(pf2 a, pf2 b, pf2 c) {
  pf2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ pf2 mul2(pf2 a, pf2 b) {
  pf2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ float hsum(pf2 a) {
  float lo, hi;
  unpk(a, lo, hi);
  return lo + hi;
}

template <int NP>
__global__ void __launch_bounds__(kThreadsB, 1)
tri_phase_b_kernel(const float *__restrict__ wpack, float *__restrict__ ws, int64_t nb) {
  extern __shared__ __align__(16) float sm_raw[];
  using S = SmemB<NP>;
  using C = CfgB<NP>;
  using W = WS<NP>;
  constexpr int RS = C::RS, RPT = C::RPT, KC = C::KC, NST = C::NST, NSLOT = S::NSLOT, NG = S::NG, NCH = S::NCH;
  static_assert(NST <= 4, "barrier slots");
  constexpr int NIT = NP + 4;
  constexpr int L = 3;
  const float rng = kCoordsRange / (float)L;

  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float *wsm = reinterpret_cast<float *>(base + S::oW);
  float *gst = reinterpret_cast<float *>(base + S::oG);
  float *chb = reinterpret_cast<float *>(base + S::oCh);
  float *fl = reinterpret_cast<float *>(base + S::oF);
  float *sVec = fl + S::fVec;
  uint64_t *bars = reinterpret_cast<uint64_t *>(fl + S::fBar);
  // barrier indices: 0..3 ops_ready[team][stage], 4..7 d_ready[team][stage], 8..8+NST-1 full[stage], 12..12+NST-1 empty[stage]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fl + S::fBar + 32);
  const uint32_t bar0 = umma::smem_u32(bars);

  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int k = 0; k < 4; ++k) umma::mbar_init(bars + k, 128);
    for (int k = 4; k < 8; ++k) umma::mbar_init(bars + k, 1);
    for (int k = 8; k < 8 + NST; ++k) umma::mbar_init(bars + k, 1);
    for (int k = 12; k < 12 + NST; ++k) umma::mbar_init(bars + k, kComputeThreads);
    umma::fence_mbar_init();
  }
  const float *W1 = wpack + pk::kHeader + pk::kLayer;
  for (int k = tid; k < kNumVec * 32; k += kThreadsB) sVec[k] = __ldg(W1 + pk::c1 + k);
  {  // weight tiles, hi + lo
    const float *srcs[4] = {W1 + pk::W2_b, W1 + pk::Wc1_b, W1 + pk::Wc1_f, W1 + pk::W2_f};
    for (int item = tid; item < 4 * 32; item += kThreadsB) {
      const int m = item >> 5, row = item & 31;
      const float4 *g = reinterpret_cast<const float4 *>(srcs[m] + row * 32);
      float h[32], l[32];
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 q = __ldg(g + k4);
        umma::split_tf32(q.x, h[4 * k4], l[4 * k4]);
        umma::split_tf32(q.y, h[4 * k4 + 1], l[4 * k4 + 1]);
        umma::split_tf32(q.z, h[4 * k4 + 2], l[4 * k4 + 2]);
        umma::split_tf32(q.w, h[4 * k4 + 3], l[4 * k4 + 3]);
      }
      umma::store_row_sw128(wsm + m * 2048, row, h);
      umma::store_row_sw128(wsm + m * 2048 + 1024, row, l);
    }
    umma::fence_proxy_async_smem();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = uniform32(*tmem_slot);
  const uint32_t w_addr = uniform32(umma::smem_u32(wsm));
  const int64_t ntile = nb * NG;
  uint32_t tok0 = 0;  // opaque zero born after the setup barrier: orders the schedulable loads of the layer vectors (lds4s)
  asm volatile("" : "+r"(tok0)::"memory");

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsHelper));
  if (warp == 0) {
    // =================================================================================== MMA issuer (one elected lane)
    if (elect_one()) {
      uint32_t rq = 0;  // both teams post the same request sequence
      const uint32_t g_addr = umma::smem_u32(gst);
      for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
#pragma unroll 1
        for (int step = 0; step < 5 + NIT; ++step, ++rq) {
#pragma unroll 1
          for (uint32_t team = 0; team < 2; ++team) {
            mbar_wait_sleepy(bar0 + 8u * (team * 2u + (rq & 1u)), (rq >> 1) & 1u);
            umma::fence_after_thread_sync();
            const uint32_t tc = tmem_base + team * 256u;
            if (step < 4) issue_full(tc, w_addr, step);            // prologue products: W2, Wc1, Wc1^T, W2^T
            else if (step == 4) issue_gtable(tc, g_addr + team * 16384u);
            else issue_item(tc, w_addr, rq & 1u, step - 5 >= NP);  // generic: TF32; S / R items: split weights
            commit_to(bar0 + 8u * (4u + team * 2u + (rq & 1u)));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================================================================== copy producer (one elected lane)
    if (elect_one()) {
      uint32_t gc = 0;
      for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int64_t lp = tile / NG;
        const int grp = (int)(tile - lp * NG);
        const float *wsp = ws + (size_t)lp * (size_t)W::kFloats;
        for (int ch = 0; ch < NCH; ++ch, ++gc) {
          const uint32_t b = gc % NST;
          mbar_wait_sleepy(bar0 + 8u * (12 + b), ((gc / NST) & 1u) ^ 1u);
          const int k0 = ch * KC, kc = (NP - k0 < KC) ? (NP - k0) : KC;
          int nrecv = NP - grp * NSLOT;
          if (nrecv > NSLOT) nrecv = NSLOT;
          const uint32_t full = bar0 + 8u * (8 + b);
          mbar_expect_tx(full, (uint32_t)(kc * NP * kTS * 4 + nrecv * kc * kTR * 4));
          const uint32_t dst = umma::smem_u32(chb + (size_t)b * S::kChunk);
          for (int q = 0; q < kc; ++q)  // one bulk copy per k: several smaller copies in flight instead of one long one
            bulk_g2s(dst + (uint32_t)(q * NP * kTS) * 4u, wsp + W::oTS + (size_t)(k0 + q) * NP * kTS, (uint32_t)(NP * kTS * 4), full);
          for (int s = 0; s < nrecv; ++s)
            bulk_g2s(dst + (uint32_t)(S::kTSChunk + s * KC * kTR) * 4u,
                     wsp + W::oTR + ((size_t)(grp * NSLOT + s) * NP + k0) * kTR, (uint32_t)(kc * kTR * 4), full);
        }
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    // =================================================================================== compute teams
    const int team = (warp >> 2) - 1, r = tid & 127;  // warps 4..7: team 0, warps 8..11: team 1
    TeamB T;
    T.tmem_col = tmem_base + (uint32_t)team * 256u;
    T.tmem = T.tmem_col + (((uint32_t)((warp & 3) * 32)) << 16);
    T.ops_addr = bar0 + 8u * (uint32_t)(team * 2);
    T.dr_addr = bar0 + 8u * (uint32_t)(4 + team * 2);
    T.q = 0;
    T.bar_id = 1 + team;
    float4 *sY = reinterpret_cast<float4 *>(fl + S::fY) + team * NP;
    float *sRed = fl + S::fRed + team * 4;
    float *gteam = gst + team * 4096;
    const int slot = r / RS, jj = r - slot * RS;
    const int cslot = team * RPT + slot;  // receiver slot within the CTA tile
    uint32_t gc = 0;

    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
      const int64_t lp = tile / NG;
      const int grp = (int)(tile - lp * NG);
      const float *wsp = ws + (size_t)lp * (size_t)W::kFloats;
      const int i_raw = grp * NSLOT + cslot;
      const bool valid = (i_raw < NP) && (jj < NP - 1);
      const int i = (i_raw < NP) ? i_raw : (NP - 1);
      int j = (jj < NP - 1) ? jj : 0;
      j += (j >= i) ? 1 : 0;
      // ---- node data
      for (int k = r; k < NP; k += kRows) sY[k] = ldg4(wsp + W::oY + 4 * k);
      T.sync();
      const float4 yi = ldg4(wsp + W::oY + 4 * i), yj = ldg4(wsp + W::oY + 4 * j);
      const float4 xi = ldg4(wsp + W::oX1 + 4 * i), xj = ldg4(wsp + W::oX1 + 4 * j);
      const Geo g = edge_geo4(xi, xj, yi, yj);
      float f1[32], f2[32], vt[32], row[32];
      float att, th;
      float g1R[3], g1S;
      // ---- prologue: primal quantities of the edge (four 3xTF32 products) ------------------------------------
      {
        float hsplit[32], lsplit[32];
        auto put2 = [&](const float(&v)[32]) {
#pragma unroll
          for (int k = 0; k < 32; ++k) umma::split_tf32(v[k], hsplit[k], lsplit[k]);
          T.st(cA0, hsplit);
          T.st(cA1, lsplit);
        };
        // z1 = P1_i + Q1_j + c1 r2 + d1 ea
        {
          const float *Pp = wsp + W::oP1 + (size_t)i * 32, *Qp = wsp + W::oQ1 + (size_t)j * 32;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 a = ldg4(Pp + 4 * k4), b = ldg4(Qp + 4 * k4);
            row[4 * k4] = a.x + b.x; row[4 * k4 + 1] = a.y + b.y; row[4 * k4 + 2] = a.z + b.z; row[4 * k4 + 3] = a.w + b.w;
          }
        }
        stage1<true, false>(row, f1, nullptr, 0, sVec, g.r2, g.ea);
        put2(row);
        uint32_t rq = T.post();
        T.wait(rq);
        T.ld(cD0, row);
        att = stage2<true>(row, vt, f2, sVec);  // row = m*att, vt = m (for now), f2
        T.st(cD1, vt);                          // park m: operand of the <m, gamma> table, re-read for <m, v>
        // exact <m, cot> for the four full-precision items
        {
          const float *tr = wsp + W::oTR + ((size_t)i * NP + j) * kTR + trGam;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const float *ga = wsp + W::oGAgg + (size_t)i * 96;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 t4 = ldg4(tr + 4 * k4), r0 = ldg4(ga + 4 * k4), r1 = ldg4(ga + 32 + 4 * k4), r2 = ldg4(ga + 64 + 4 * k4);
            const float mm[4] = {vt[4 * k4], vt[4 * k4 + 1], vt[4 * k4 + 2], vt[4 * k4 + 3]};
            a0 += mm[0] * t4.x + mm[1] * t4.y + mm[2] * t4.z + mm[3] * t4.w;
            a1 += mm[0] * r0.x + mm[1] * r0.y + mm[2] * r0.z + mm[3] * r0.w;
            a2 += mm[0] * r1.x + mm[1] * r1.y + mm[2] * r1.z + mm[3] * r1.w;
            a3 += mm[0] * r2.x + mm[1] * r2.y + mm[2] * r2.z + mm[3] * r2.w;
          }
          g1S = a0; g1R[0] = a1; g1R[1] = a2; g1R[2] = a3;
        }
        put2(row);
        rq = T.post();
        T.wait(rq);
        T.ld(cD0, row);
        {  // th and rowc = wc2 * silu'(zc)
          float u = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 b = lds4(sVec + vBC1 * 32 + 4 * k4), w4 = lds4(sVec + vWC2 * 32 + 4 * k4);
            const float bb[4] = {b.x, b.y, b.z, b.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a, f;
              silu_both(row[4 * k4 + e] + bb[e], a, f);
              u = fmaf(ww[e], a, u);
              row[4 * k4 + e] = ww[e] * f;
            }
          }
          th = tanhf(u);
        }
        put2(row);
        rq = T.post();
        T.wait(rq);
        T.ld(cD0, row);  // v
        T.ld(cD1, vt);   // m
        float mv = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) mv = fmaf(vt[k], row[k], mv);
        mv *= att * (1.0f - att);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 w4 = lds4(sVec + vWA * 32 + 4 * k4);
          row[4 * k4] = f2[4 * k4] * fmaf(att, row[4 * k4], mv * w4.x);
          row[4 * k4 + 1] = f2[4 * k4 + 1] * fmaf(att, row[4 * k4 + 1], mv * w4.y);
          row[4 * k4 + 2] = f2[4 * k4 + 2] * fmaf(att, row[4 * k4 + 2], mv * w4.z);
          row[4 * k4 + 3] = f2[4 * k4 + 3] * fmaf(att, row[4 * k4 + 3], mv * w4.w);
        }
        put2(row);
        rq = T.post();
        T.wait(rq);
        T.ld(cD0, vt);
#pragma unroll
        for (int k = 0; k < 32; ++k) vt[k] *= f1[k];
      }
      float duc = 0.f, dvd1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 c4 = lds4(sVec + vC1 * 32 + 4 * k4), d4 = lds4(sVec + vD1 * 32 + 4 * k4);
        duc += vt[4 * k4] * c4.x + vt[4 * k4 + 1] * c4.y + vt[4 * k4 + 2] * c4.z + vt[4 * k4 + 3] * c4.w;
        dvd1 += vt[4 * k4] * d4.x + vt[4 * k4 + 1] * d4.y + vt[4 * k4 + 2] * d4.z + vt[4 * k4 + 3] * d4.w;
      }
      // ---- <m, gamma_ki> table: stage the gamma rows of the team's receivers (row = slot * GS + k) and run one MMA
      {
        constexpr int GS = kRows / RPT;
        const int gs_slot = r / GS, gk = r - gs_slot * GS;
        const int gi = grp * NSLOT + team * RPT + gs_slot;
        float v[32];
        if (gi < NP && gk < NP) {
          const float *tr = wsp + W::oTR + ((size_t)gi * NP + gk) * kTR + trGam;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 t4 = ldg4(tr + 4 * k4);
            v[4 * k4] = t4.x; v[4 * k4 + 1] = t4.y; v[4 * k4 + 2] = t4.z; v[4 * k4 + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0.f;
        }
        umma::store_row_sw128(gteam, r, v);
        umma::fence_proxy_async_smem();
        const uint32_t rq = T.post();
        T.wait(rq);
      }
      // ---- per-edge scalars
      const float sgate = att, sgate2 = att * (1.0f - att);
      const float phi = rng * th, cphi = rng * (1.0f - th * th), ivn = g.inv / g.nrm;
      const float dl[3] = {g.d0, g.d1, g.d2};
      const float dh[3] = {g.d0 * g.inv, g.d1 * g.inv, g.d2 * g.inv};
      float acc = 0.f;
      constexpr int GS = kRows / RPT;
      const uint32_t gcol = cG + (uint32_t)(slot * GS);

      // ---- items: iteration `it` builds item it and finishes item it-1 ----------------------------------------
      uint32_t rq_prev = 0;
      const float *cot_prev = nullptr;  // shared-memory gamma row of the pending generic item
      bool mask_prev = false;
      int k_prev = 0;
#pragma unroll 1
      for (int it = 0; it <= NIT; ++it) {
        uint32_t rq_cur = 0;
        const float *cot_cur = nullptr;
        bool mask_cur = false;
        if (it < NP) {
          // ======================================================================= generic item, k = it
          const int k = it, ch = k / KC, kk = k - ch * KC;
          const uint32_t b = (gc + (uint32_t)ch) % NST;
          if (kk == 0) mbar_wait_addr(bar0 + 8u * (8 + b), ((gc + (uint32_t)ch) / NST) & 1u);
          uint32_t tok = 0;  // born after the chunk's full-barrier wait
          asm volatile("" : "+r"(tok)::"memory");
          const float *cb = chb + (size_t)b * S::kChunk;
          const float *tse = cb + ((size_t)kk * NP + j) * kTS;
          const float *tre = cb + S::kTSChunk + ((size_t)cslot * KC + kk) * kTR;
          // small tables
          const float4 q0 = lds4s(tre + trW, tok);        // w0 w1 w2 alpha_i
          const float4 q1 = lds4s(tre + trGX, tok);       // GX 0..3
          const float4 q2 = lds4s(tre + trGX + 4, tok);   // GX 4..7
          const float4 q3 = lds4s(tre + trGX + 8, tok);   // GX 8, M 0..2
          const float4 q4 = lds4s(tre + trGX + 12, tok);  // M 3..6
          const float4 q5 = lds4s(tre + trGX + 16, tok);  // M 7..8, pad
          const float4 s0 = lds4s(tse + tsM, tok), s1 = lds4s(tse + tsM + 4, tok), s2 = lds4s(tse + tsM + 8, tok);
          const float4 yk = lds4s(reinterpret_cast<const float *>(sY + k), tok0);
          const float GX[9] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x};
          const float Mi[9] = {q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w, q5.x, q5.y};
          const float Mj[9] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x};
          const float cfi[3] = {-2.0f * (yi.x - yk.x), -2.0f * (yi.y - yk.y), -2.0f * (yi.z - yk.z)};
          const float cfj[3] = {-2.0f * (yj.x - yk.x), -2.0f * (yj.y - yk.y), -2.0f * (yj.z - yk.z)};
          float xis[3], gv[3], fro = 0.f;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            xis[a] = GX[a] * dh[0] + GX[3 + a] * dh[1] + GX[6 + a] * dh[2];
            gv[a] = 2.0f * ((Mi[a] - Mj[a]) * dl[0] + (Mi[3 + a] - Mj[3 + a]) * dl[1] + (Mi[6 + a] - Mj[6 + a]) * dl[2]);
          }
#pragma unroll
          for (int q = 0; q < 9; ++q) fro = fmaf(GX[q], Mi[q] - Mj[q], fro);
          const float al_i = q0.w;
          const float al_j = q0.x * cfj[0] + q0.y * cfj[1] + q0.z * cfj[2];
          const float bet = q0.x * gv[0] + q0.y * gv[1] + q0.z * gv[2];
          const float xg = xis[0] * gv[0] + xis[1] * gv[1] + xis[2] * gv[2];
          float b1, b2;
          {
            const pf2 AI = pk(al_i, al_i), AJ = pk(al_j, al_j), BT = pk(bet, bet);
            pf2 b1p[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, b2p[2] = {pk(0.f, 0.f), pk(0.f, 0.f)};  // two chains each: ILP at 2 warps / scheduler
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 pa = lds4s(tre + trPA + 4 * k4, tok), pb = lds4s(tse + tsPB + 4 * k4, tok), c4 = lds4s(sVec + vC1 * 32 + 4 * k4, tok0);
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = 4 * k4 + 2 * e;
                const pf2 pa2 = e ? pk(pa.z, pa.w) : pk(pa.x, pa.y), pb2 = e ? pk(pb.z, pb.w) : pk(pb.x, pb.y);
                const pf2 c2 = e ? pk(c4.z, c4.w) : pk(c4.x, c4.y);
                const pf2 vt2 = pk(vt[c], vt[c + 1]);
                const pf2 u2 = mul2(pk(f1[c], f1[c + 1]), fma2(AI, pa2, fma2(AJ, pb2, mul2(BT, c2))));
                unpk(u2, row[c], row[c + 1]);
                b1p[e] = fma2(vt2, pa2, b1p[e]);
                b2p[e] = fma2(vt2, pb2, b2p[e]);
              }
            }
            b1 = hsum(b1p[0]) + hsum(b1p[1]);
            b2 = hsum(b2p[0]) + hsum(b2p[1]);
          }
          const float tB = cphi * ((xis[0] * cfi[0] + xis[1] * cfi[1] + xis[2] * cfi[2]) * b1 +
                                   (xis[0] * cfj[0] + xis[1] * cfj[1] + xis[2] * cfj[2]) * b2 + xg * duc);
          const float e2 = phi * (g.inv * fro - 0.5f * ivn * xg);
          mask_cur = valid && (k != i);
          acc += mask_cur ? (tB + e2) : 0.f;
          T.st((T.q & 1u) ? cA1 : cA0, row);
          rq_cur = T.post();
          cot_cur = tre + trGam;
        } else if (it == NP) {
          // ======================================================================= S item (k = j)
          const float *tr = wsp + W::oTR + ((size_t)i * NP + j) * kTR;
          const float4 q0 = ldg4(tr + trW);
          const float4 q1 = ldg4(tr + trGX), q2 = ldg4(tr + trGX + 4);
          const float gx8 = tr[trGX + 8];
          const float GX[9] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, gx8};
          float xis[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) xis[a] = GX[a] * dh[0] + GX[3 + a] * dh[1] + GX[6 + a] * dh[2];
          const float *bo = wsp + W::oBOm + (size_t)j * 96;
          float bs = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 o0 = ldg4(bo + 4 * k4), o1 = ldg4(bo + 32 + 4 * k4), o2 = ldg4(bo + 64 + 4 * k4);
            const float4 d4 = lds4(sVec + vD1 * 32 + 4 * k4);
            const float o0_[4] = {o0.x, o0.y, o0.z, o0.w}, o1_[4] = {o1.x, o1.y, o1.z, o1.w}, o2_[4] = {o2.x, o2.y, o2.z, o2.w};
            const float d_[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * k4 + e;
              const float uin = fmaf(q0.x, o0_[e], fmaf(q0.y, o1_[e], fmaf(q0.z, o2_[e], q0.w * d_[e])));
              row[c] = umma::round_tf32(f1[c] * uin);
              bs = fmaf(vt[c], fmaf(xis[0], o0_[e], fmaf(xis[1], o1_[e], xis[2] * o2_[e])), bs);
            }
          }
          const float xcf = -2.0f * (xis[0] * (yi.x - yj.x) + xis[1] * (yi.y - yj.y) + xis[2] * (yi.z - yj.z));
          acc += valid ? cphi * (bs + xcf * dvd1) : 0.f;
          mask_cur = valid;
          T.st((T.q & 1u) ? cA1 : cA0, row);
          rq_cur = T.post();
        } else if (it < NIT) {
          // ======================================================================= R item (k = i), direction a
          const int a = it - NP - 1;
          const float *trd = wsp + W::oTR + ((size_t)i * NP + i) * kTR;
          const float *tsr = wsp + W::oTS + ((size_t)i * NP + j) * kTS;  // TS[k = i][j]
          float GX[9], dM[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            GX[q] = trd[trGX + q];
            dM[q] = trd[trM + q] - tsr[tsM + q];
          }
          float xis[3], gv[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            xis[c] = GX[c] * dh[0] + GX[3 + c] * dh[1] + GX[6 + c] * dh[2];
            gv[c] = 2.0f * (dM[c] * dl[0] + dM[3 + c] * dl[1] + dM[6 + c] * dl[2]);
          }
          const float cfa = 2.0f * (a == 0 ? (yi.x - yj.x) : (a == 1 ? (yi.y - yj.y) : (yi.z - yj.z)));
          const float ga = (a == 0) ? gv[0] : ((a == 1) ? gv[1] : gv[2]);
          const float xa = (a == 0) ? xis[0] : ((a == 1) ? xis[1] : xis[2]);
          const float *ao = wsp + W::oAOm + ((size_t)i * 3 + a) * 32;
          float vtan = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 o = ldg4(ao + 4 * k4), pb = ldg4(tsr + tsPB + 4 * k4);
            const float4 c4 = lds4(sVec + vC1 * 32 + 4 * k4), d4 = lds4(sVec + vD1 * 32 + 4 * k4);
            const float o_[4] = {o.x, o.y, o.z, o.w}, pb_[4] = {pb.x, pb.y, pb.z, pb.w}, c_[4] = {c4.x, c4.y, c4.z, c4.w},
                        d_[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * k4 + e;
              const float tan = fmaf(cfa, pb_[e] + d_[e], fmaf(ga, c_[e], o_[e]));
              vtan = fmaf(vt[c], tan, vtan);
              row[c] = umma::round_tf32(f1[c] * tan);
            }
          }
          float add = cphi * xa * vtan;
          if (a == 0) {
            float fro = 0.f;
#pragma unroll
            for (int q = 0; q < 9; ++q) fro = fmaf(GX[q], dM[q], fro);
            add += phi * (g.inv * fro - 0.5f * ivn * (xis[0] * gv[0] + xis[1] * gv[1] + xis[2] * gv[2]));
          }
          acc += valid ? add : 0.f;
          mask_cur = valid;
          T.st((T.q & 1u) ? cA1 : cA0, row);
          rq_cur = T.post();
        }
        // ------------------------------------------------------------------------- finish the previous item
        if (it > 0) {
          const int ip = it - 1;
          T.wait(rq_prev);
          uint32_t tokf = 0;
          asm volatile("" : "+r"(tokf)::"memory");
          float dg = 0.f, dw = 0.f, g1;
          T.ld((rq_prev & 1u) ? cD1 : cD0, row);
          if (ip < NP) {
            // <m, gamma_ki> from the table (the slots of one warp differ when RS < 32)
            if (RS >= 32) {
              g1 = tmem_ld1(T.tmem + gcol + (uint32_t)k_prev);
            } else {
              g1 = 0.f;
              const int s_lo = ((warp & 3) * 32) / RS;
#pragma unroll
              for (int ss = 0; ss < 32 / RS; ++ss) {
                const float v1 = tmem_ld1(T.tmem + cG + (uint32_t)((s_lo + ss) * GS + k_prev));
                if (slot == s_lo + ss) g1 = v1;
              }
            }
            {
              pf2 dgp[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, dwp[2] = {pk(0.f, 0.f), pk(0.f, 0.f)};
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 c4 = lds4s(cot_prev + 4 * k4, tokf), w4 = lds4s(sVec + vWA * 32 + 4 * k4, tok0);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int c = 4 * k4 + 2 * e;
                  const pf2 t2 = mul2(pk(f2[c], f2[c + 1]), pk(row[c], row[c + 1]));
                  dgp[e] = fma2(t2, e ? pk(c4.z, c4.w) : pk(c4.x, c4.y), dgp[e]);
                  dwp[e] = fma2(t2, e ? pk(w4.z, w4.w) : pk(w4.x, w4.y), dwp[e]);
                }
              }
              dg = hsum(dgp[0]) + hsum(dgp[1]);
              dw = hsum(dwp[0]) + hsum(dwp[1]);
            }
            // release the chunk once its last item is finished
            const int chp = ip / KC;
            if (ip == NP - 1 || ip - chp * KC == KC - 1) {
              asm volatile("" ::"f"(dg), "f"(dw) : "memory");  // the chunk's last reads are complete before it is handed back
              mbar_arrive(bar0 + 8u * (12 + ((gc + (uint32_t)chp) % NST)));
            }
          } else {
            const float *cot = (ip == NP) ? (wsp + W::oTR + ((size_t)i * NP + j) * kTR + trGam)
                                          : (wsp + W::oGAgg + ((size_t)i * 3 + (ip - NP - 1)) * 32);
            g1 = (ip == NP) ? g1S : (ip == NP + 1 ? g1R[0] : (ip == NP + 2 ? g1R[1] : g1R[2]));
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 c4 = ldg4(cot + 4 * k4), w4 = lds4(sVec + vWA * 32 + 4 * k4);
              const float c_[4] = {c4.x, c4.y, c4.z, c4.w}, w_[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float t = f2[4 * k4 + e] * row[4 * k4 + e];
                dg = fmaf(t, c_[e], dg);
                dw = fmaf(t, w_[e], dw);
              }
            }
          }
          acc += mask_prev ? (sgate * dg + sgate2 * g1 * dw) : 0.f;
        }
        rq_prev = rq_cur;
        cot_prev = cot_cur;
        mask_prev = mask_cur;
        k_prev = it;
      }
      gc += NCH;
      // ---- team partial of the trace
      acc = warp_sum(acc);
      if (lane == 0) sRed[warp & 3] = acc;
      T.sync();
      if (r == 0) ws[(size_t)lp * (size_t)W::kFloats + W::oPartB + grp * 2 + team] = (sRed[0] + sRed[1]) + (sRed[2] + sRed[3]);
      T.sync();
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem_base);
}

// div[p] = ((c_s - 1) D + c_out c_in (tr - D)) / h,  tr = sum of the direct part and the phase-B partials
template <int NP>
__global__ void tri_finalize_kernel(const float *__restrict__ ws, const float *__restrict__ ht, int64_t b0, int64_t nb,
                                    float *__restrict__ divergence) {
  using W = WS<NP>;
  constexpr int NPART = SmemB<NP>::NG * 2;
  const int64_t lp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= nb) return;
  const float *wsp = ws + (size_t)lp * (size_t)W::kFloats;
  float tr = 0.f;
  for (int k = 0; k < NP; ++k) tr += wsp[W::oDirect + k];
  for (int k = 0; k < NPART; ++k) tr += wsp[W::oPartB + k];
  const float h = ht[b0 + lp];
  const float c_in = rsqrtf(1.0f + h), c_s = 1.0f / (1.0f + h), c_out = sqrtf(h) * c_in;
  const float Dn = (float)(3 * NP);
  divergence[b0 + lp] = ((c_s - 1.0f) * Dn + c_out * c_in * (tr - Dn)) / h;
}

template <int NP>
static int launch_b(const float *w, const float *ht, int64_t b0, int64_t nb, float *ws, float *divergence, cudaStream_t s) {
  using S = SmemB<NP>;
  static_assert(S::kBytes <= 227 * 1024, "phase B shared memory plan exceeds 227 KB");
  static_assert(S::NG * 2 <= 32, "phase B partial slots");
  auto k = tri_phase_b_kernel<NP>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
  if (e != cudaSuccess) { set_error("tri_phase_b_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  const int64_t ntile = nb * S::NG;
  const unsigned grid = (unsigned)(ntile < kNumSMs ? ntile : kNumSMs);
  k<<<grid, kThreadsB, S::kBytes, s>>>(w, ws, nb);
  PITA_CHECK_LAUNCH("tri_phase_b_kernel");
  tri_finalize_kernel<NP><<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ws, ht, b0, nb, divergence);
  PITA_CHECK_LAUNCH("tri_finalize_kernel");
  return PITA_OK;
}

int launch_tri_phase_b(int n, const float *w, const float *ht, int64_t b0, int64_t nb, float *ws, float *divergence,
                       cudaStream_t s) {
  if (n == 13) return launch_b<13>(w, ht, b0, nb, ws, divergence, s);
  return launch_b<55>(w, ht, b0, nb, ws, divergence, s);
}

}  // namespace tri
}  // namespace pita
