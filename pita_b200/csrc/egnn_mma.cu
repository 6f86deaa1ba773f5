// Score net: score + exact divergence, with the dense (middle-layer) tangent contraction on the sm_100a
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).  Replaces vmap(jacrev(score_net)) + trace
// (models/components/utils.py:30-51, called at sdes.py:193-203) — >95 % of the reference's step time.
//
// Per particle (one CTA) the tangent bundle d/dy[k,a] is pushed through the 3 E_GCL layers in passes of
// kTN tangent nodes (kT = 6 directions).  Layer 0 (edges touching a tangent node) and layer 2 (receivers
// that are tangent nodes) are sparse and stay on the fp32 SIMT path of egnn.cu.  Layer 1 is dense:
// for every edge (i,j) and direction t,
//     dms = D2 (W2 (D1 dz1)) ... ,   dzc = Wc1 dms
// i.e. two [rows x 32] x [32 x 32] GEMMs whose rows are (edge, direction) pairs.  A team of 4 warps owns a
// 128-row tile = 16 edge slots x 8 rows (6 directions, 1 spare, 1 PRIMAL row: the edge's own activations ride
// through the same GEMMs, so no SIMT mat-vec is left in this layer).  thread = row: each thread builds its
// operand row in registers, stores it K-major into 128B-swizzled shared memory, one elected thread issues
// the MMAs, and every thread reads its accumulator row back from TMEM (tcgen05.ld 32x32b) for the
// element-wise stage.  mode 1 = 3xTF32 (hi/lo split of both operands, fp32-accurate), mode 2 = plain TF32.
#include "egnn_common.cuh"
#include "umma.cuh"

namespace pita {

constexpr int kTN = 2;       // tangent nodes per pass
constexpr int kT = 3 * kTN;  // directions per pass
constexpr int kR = 8;        // rows per edge slot: kT directions, one spare (zero) row, one primal row
constexpr int kSlots = 16;   // edge slots per 128-row tile
constexpr int kDQS = 36;     // padded row stride (floats) of per-node tangent rows: conflict-free float4 row reads
constexpr int kPrimalRow = kR - 1;

template <int NP, int NTEAM, int G>
struct MPlan {
  using P = Plan<NP, NTEAM * 4, 3>;
  static constexpr int kThreads = NTEAM * 128;
  static constexpr int SR = kSlots / G;  // sender slots per receiver per tile
  static constexpr int kTilesPerGroup = (NP - 1 + SR - 1) / SR;
  static constexpr int kGroups = (NP + G - 1) / G;
  // ---- float-indexed regions (tangent state overlays the forward-only scratch of the primal plan)
  static constexpr int oDQ = P::kPersist;                  // [NP][kT][kDQS]   B^1 dh^1
  static constexpr int oDXa = oDQ + NP * kT * kDQS;        // [NP][kT][4]      d x^1
  static constexpr int oDXb = oDXa + NP * kT * 4;          // [NP][kT][4]      d x^2
  static constexpr int oDP2 = oDXb + NP * kT * 4;          // [kTN][3][H]      A^2 dh^2 of tangent nodes
  static constexpr int oMisc = oDP2 + kTN * 3 * H;         // mbarriers (2 per team) + tmem slot
  static constexpr int oTeam = oMisc + 32;
  static constexpr int tDP = 0;                            // [G][kR][kDQS]    A^1 dh^1 of the group's receivers
  static constexpr int tV = tDP + G * kR * kDQS;           // 4 x [kSlots][H]  per-edge primal vectors
  static constexpr int kTeamFloats = tV + 4 * kSlots * H;
  static constexpr int oEnd0 = oTeam + NTEAM * kTeamFloats;
  static constexpr int oEnd = oEnd0 > P::kPrimal ? oEnd0 : P::kPrimal;
  // ---- 1024-byte aligned operand tiles: W2 hi/lo, Wc1 hi/lo (4 KB each), then per team A hi / A lo (16 KB each)
  static constexpr size_t kTileBase = ((size_t)oEnd * 4 + 1023) / 1024 * 1024;
  static constexpr size_t kWBytes = 4 * 4096;
  static constexpr size_t kBytes = 1024 + kTileBase + kWBytes + (size_t)NTEAM * 2 * 16384;
  static constexpr int kXFloats = 2 * NP * kT * H;         // global scratch per CTA: dh^1 -> B^2 dh^2, and A^1 dh^1
};

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }

__device__ __forceinline__ float group8_sum(float v) {  // sum over the 8 lanes of an aligned 8-lane group
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <bool SPLIT>
__device__ __forceinline__ void store_operand_row(float *hi_tile, float *lo_tile, int row, const float (&v)[32]) {
  if (SPLIT) {
    float h[32], l[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) umma::split_tf32(v[k], h[k], l[k]);
    umma::store_row_sw128(hi_tile, row, h);
    umma::store_row_sw128(lo_tile, row, l);
  } else {
    umma::store_row_sw128(hi_tile, row, v);
  }
}

template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const float *a_hi, const float *a_lo, const float *b_hi,
                                           const float *b_lo, uint64_t *mbar) {
  constexpr uint32_t idesc = umma::make_idesc_tf32(128, 32);
  const uint64_t dA = umma::make_desc_sw128_kmajor(umma::smem_u32(a_hi)), dB = umma::make_desc_sw128_kmajor(umma::smem_u32(b_hi));
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_d, dA + 2 * k, dB + 2 * k, idesc, k > 0);
  if (SPLIT) {
    const uint64_t dAl = umma::make_desc_sw128_kmajor(umma::smem_u32(a_lo)), dBl = umma::make_desc_sw128_kmajor(umma::smem_u32(b_lo));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_d, dAl + 2 * k, dB + 2 * k, idesc, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::mma_tf32_ss(tmem_d, dA + 2 * k, dBl + 2 * k, idesc, 1);
  }
  umma::commit(mbar);
}

template <int NP, int NTEAM, int G, bool SPLIT>
__global__ void __launch_bounds__(NTEAM * 128, 1)
egnn_score_div_mma_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                          const float *__restrict__ beta, int64_t B, float *__restrict__ score, float *__restrict__ divergence,
                          float *__restrict__ scratch) {
  constexpr int L = 3;
  constexpr int NW = NTEAM * 4;
  using P = Plan<NP, NW, L>;
  using M = MPlan<NP, NTEAM, G>;
  constexpr int SR = M::SR;
  extern __shared__ __align__(16) float sm_raw[];
  // align the whole window to 1024 B so that tile offsets computed from the plan are 1024-aligned
  float *sm = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = warp >> 2, tt = tid & 127;  // team-local thread = tile row
  float4 *sX = reinterpret_cast<float4 *>(sm + P::oX);
  float *sQ = sm + P::oQ, *sP = sm + P::oP, *sZ3 = sm + P::oZ3, *sRed = sm + P::oRed;
  float *sDQ = sm + M::oDQ, *sDP2 = sm + M::oDP2;
  float4 *sDXa = reinterpret_cast<float4 *>(sm + M::oDXa);
  float4 *sDXb = reinterpret_cast<float4 *>(sm + M::oDXb);
  uint64_t *mbars = reinterpret_cast<uint64_t *>(sm + M::oMisc);  // [NTEAM][2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + M::oMisc + 4 * NTEAM);
  float *tm = sm + M::oTeam + team * M::kTeamFloats;
  float *sDP = tm + M::tDP;
  float *V0 = tm + M::tV, *V1 = V0 + kSlots * H, *V2 = V1 + kSlots * H, *V3 = V2 + kSlots * H;
  uint8_t *tiles = reinterpret_cast<uint8_t *>(sm) + M::kTileBase;
  float *sW2hi = reinterpret_cast<float *>(tiles), *sW2lo = sW2hi + 1024, *sWc1hi = sW2lo + 1024, *sWc1lo = sWc1hi + 1024;
  float *sAhi = reinterpret_cast<float *>(tiles + M::kWBytes + (size_t)team * 32768), *sAlo = sAhi + 4096;
  // SIMT staging (phases A / C, group epilogues, primal forward) lives in the team's A tiles while no MMA is in flight
  float *simt_stage = sAhi + 128 * kDQS + (warp & 3) * Stage<kT>::kFloats;  // after the [128][kDQS] reduction area
  Stage<kT> st(simt_stage);
  float *fwd_stage = reinterpret_cast<float *>(tiles + M::kWBytes);  // NW * Stage<1> floats, only used by primal_forward
  float *gX = scratch + (size_t)blockIdx.x * M::kXFloats;  // [NP][kT][H] dh^1 -> B^2 dh^2
  float *gXP = gX + NP * kT * H;                           // [NP][kT][H] A^1 dh^1
  const float rng = kCoordsRange / (float)L;
  const float *__restrict__ W0 = wpack + pk::kHeader;
  const float *__restrict__ W1 = W0 + pk::kLayer;
  const float *__restrict__ W2l = W1 + pk::kLayer;

  // ---- one-time setup: TMEM, mbarriers, layer-1 weight operand tiles (B operands: [N=out][K=in], K-major)
  if (warp == 0) umma::tmem_alloc<NTEAM * 64>(tmem_slot);
  if (tid == 0) {
    for (int k = 0; k < 2 * NTEAM; ++k) umma::mbar_init(mbars + k, 1);
    umma::fence_mbar_init();
  }
  if (tid < 64) {
    const float *src = (tid < 32 ? W1 + pk::W2_b : W1 + pk::Wc1_b) + (tid & 31) * H;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __ldg(src + k);
    if (tid < 32) store_operand_row<SPLIT>(sW2hi, sW2lo, tid, v);
    else store_operand_row<SPLIT>(sWc1hi, sWc1lo, tid - 32, v);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmD1 = tmem_base + team * 64, tmD2 = tmD1 + 32;
  const uint32_t tm_lane = ((uint32_t)((warp & 3) * 32)) << 16;
  uint32_t ph1 = 0, ph2 = 0;  // mbarrier phases (same value in every thread of the team)

  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float h = __ldg(ht + b);
    const float c_in = rsqrtf(1.0f + h);
    const float c_s = 1.0f / (1.0f + h);
    const float c_out = sqrtf(h) * c_in;
    const float c_noise = 0.125f * logf(h);
    for (int i = tid; i < NP; i += M::kThreads)
      sX[i] = make_float4(c_in * x[b * 3 * NP + 3 * i], c_in * x[b * 3 * NP + 3 * i + 1], c_in * x[b * 3 * NP + 3 * i + 2], 0.f);
    __syncthreads();
    primal_forward<NP, NW, L>(sm, wpack, c_noise, __ldg(beta + b), fwd_stage);
    {
      float4 *sV = reinterpret_cast<float4 *>(sm + P::oAgg);
      for (int i = tid; i < NP; i += M::kThreads) {
        const float4 a = sX[L * NP + i], c = sX[i];
        sV[i] = make_float4(a.x - c.x, a.y - c.y, a.z - c.z, 0.f);
      }
      __syncthreads();
      const float4 vmean = node_mean<NP>(sV, sRed);
      for (int i = tid; i < NP; i += M::kThreads) {
        const float vv[3] = {sV[i].x - vmean.x, sV[i].y - vmean.y, sV[i].z - vmean.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float xv = x[b * 3 * NP + 3 * i + k];
          score[b * 3 * NP + 3 * i + k] = ((c_s * xv + c_out * vv[k]) - xv) / h;  // score_net.py:33-35
        }
      }
      __syncthreads();
    }
    if (divergence == nullptr) continue;

    float trace = 0.f;
#pragma unroll 1
    for (int k0 = 0; k0 < NP; k0 += kTN) {
      // =========================== phase A (SIMT): layer 0 on edges touching a tangent node
      {
        float w2[H], wc1[H];
        load_row(w2, W0 + pk::W2_f, lane);
        load_row(wc1, W0 + pk::Wc1_f, lane);
        const EdgeScal sc = load_edge_scal(W0, lane);
        for (int i = warp; i < NP; i += NW) {
          float dagg[kT], dxi[kT][3];
#pragma unroll
          for (int t = 0; t < kT; ++t) { dagg[t] = 0.f; dxi[t][0] = dxi[t][1] = dxi[t][2] = 0.f; }
          const float4 xi = sX[i];
          const float pi = sP[i * H + lane];
          const int ii = i - k0;
          const bool own = (ii >= 0 && ii < kTN);
#pragma unroll 1
          for (int j = 0; j < NP; ++j) {
            if (j == i) continue;
            const int jj = j - k0;
            const bool oth = (jj >= 0 && jj < kTN);
            if (!own && !oth) continue;
            const EdgeGeo g = edge_geo(xi, sX[j], xi, sX[j]);
            const float pq = pi + sQ[j * H + lane];
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
              if (which == 0 ? !own : !oth) continue;
              const float sgn = which == 0 ? 1.0f : -1.0f;
              const int slot = which == 0 ? ii : jj;
              EdgeT<3> tin;
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                tin.dpq[a] = 0.f;
                tin.Dd[a][0] = a == 0 ? sgn : 0.f; tin.Dd[a][1] = a == 1 ? sgn : 0.f; tin.Dd[a][2] = a == 2 ? sgn : 0.f;
                tin.dea[a] = 2.0f * sgn * g.d[a];
              }
              float dms[3], dtr[3][3];
              edge_eval<3, kT>(w2, wc1, sc, rng, pq, g, st, lane, tin, dms, dtr);
#pragma unroll
              for (int t = 0; t < kT; ++t)
                if (t / 3 == slot) {
                  dagg[t] += dms[t % 3];
                  dxi[t][0] += dtr[t % 3][0]; dxi[t][1] += dtr[t % 3][1]; dxi[t][2] += dtr[t % 3][2];
                }
            }
          }
#pragma unroll
          for (int t = 0; t < kT; ++t)
            if (t / 3 == ii) dxi[t][t % 3] += 1.0f;
          if (lane < kT) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < kT; ++t)
              if (t == lane) o = make_float4(dxi[t][0], dxi[t][1], dxi[t][2], 0.f);
            sDXa[i * kT + lane] = o;
          }
#pragma unroll
          for (int t = 0; t < kT; ++t) gX[(i * kT + t) * H + lane] = dagg[t];
        }
        __syncwarp();
      }
      {  // node update on the tangents (dh^0 = 0): dh^1 = W4 (f3 * (W3a dagg)); then A^1 dh^1 and B^1 dh^1
        float wr[H];
        load_row(wr, W0 + pk::W3a_f, lane);
        for (int i = warp; i < NP; i += NW) {
          float v, f3, tmp[kT];
          silu_both(sZ3[i * H + lane], v, f3);
#pragma unroll
          for (int t = 0; t < kT; ++t) tmp[t] = f3 * dot32(wr, gX + (i * kT + t) * H);
          __syncwarp();
#pragma unroll
          for (int t = 0; t < kT; ++t) gX[(i * kT + t) * H + lane] = tmp[t];
        }
        __syncwarp();
        load_row(wr, W0 + pk::W4_f, lane);
        for (int i = warp; i < NP; i += NW) {
          float tmp[kT];
#pragma unroll
          for (int t = 0; t < kT; ++t) tmp[t] = dot32(wr, gX + (i * kT + t) * H);
          __syncwarp();
#pragma unroll
          for (int t = 0; t < kT; ++t) gX[(i * kT + t) * H + lane] = tmp[t];  // dh^1_i
        }
        __syncwarp();
        load_row(wr, W1 + pk::B_f, lane);
        for (int i = warp; i < NP; i += NW) {
#pragma unroll
          for (int t = 0; t < kT; ++t) sDQ[(i * kT + t) * kDQS + lane] = dot32(wr, gX + (i * kT + t) * H);
        }
        load_row(wr, W1 + pk::A_f, lane);
        for (int i = warp; i < NP; i += NW) {
#pragma unroll
          for (int t = 0; t < kT; ++t) gXP[(i * kT + t) * H + lane] = dot32(wr, gX + (i * kT + t) * H);
        }
      }
      __syncthreads();

      // =========================== phase B: layer 1 (dense) on the tensor cores, thread = row
      {
        const int slot = tt >> 3, r = tt & 7;
        const int gsel = slot / SR, us = slot % SR;  // receiver within the group, sender slot within the tile
        // the spare rows of both operand tiles must be zero (the tiles double as SIMT scratch between groups)
        for (int grp = team; grp < M::kGroups; grp += NTEAM) {
          const int i = grp * G + gsel;
          const bool ivalid = i < NP;
          const int ic = ivalid ? i : NP - 1;
          // group prologue: A^1 dh^1 rows of the G receivers -> team smem (padded rows)
          for (int e = tt; e < G * kT * (H / 4); e += 128) {
            const int row = e / (H / 4), c4 = e % (H / 4);
            const int gi = grp * G + row / kT;
            const float4 v = gi < NP ? *reinterpret_cast<const float4 *>(gXP + ((size_t)gi * kT + row % kT) * H + 4 * c4)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(sDP + ((row / kT) * kR + row % kT) * kDQS + 4 * c4) = v;
          }
          if (r == kT) {  // spare row: zero once per group
            float z[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) z[k] = 0.f;
            umma::store_row_sw128(sAhi, tt, z);
            if (SPLIT) umma::store_row_sw128(sAlo, tt, z);
          }
          team_sync(team);
          const float4 xi = sX[NP + ic], x0i = sX[ic];
          float aggacc[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) aggacc[k] = 0.f;
          float dxacc0 = 0.f, dxacc1 = 0.f, dxacc2 = 0.f;
          float4 dxi = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < kT) dxi = sDXa[ic * kT + r];
          const int tnode = k0 + r / 3;  // tangent node of this row's direction (r < kT)
#pragma unroll 1
          for (int tile = 0; tile < M::kTilesPerGroup; ++tile) {
            const int u = tile * SR + us;
            const bool valid = ivalid && (u < NP - 1);
            const int uc = u < NP - 1 ? u : NP - 2;
            const int j = uc + (uc >= ic ? 1 : 0);
            const float4 xj = sX[NP + j], x0j = sX[j];
            const EdgeGeo g = edge_geo(xi, xj, x0i, x0j);
            // ---------- stage 1: first edge linear.  The slot's 8 threads each own 4 primal channels.
            {
              const int c0 = 4 * r;
              const float4 p4 = *reinterpret_cast<const float4 *>(sP + (NP + ic) * H + c0);
              const float4 q4 = *reinterpret_cast<const float4 *>(sQ + (NP + j) * H + c0);
              const float4 c14 = __ldg(reinterpret_cast<const float4 *>(W1 + pk::c1 + c0));
              const float4 d14 = __ldg(reinterpret_cast<const float4 *>(W1 + pk::d1 + c0));
              float a[4], f[4];
              silu_both(p4.x + q4.x + c14.x * g.r2 + d14.x * g.ea, a[0], f[0]);
              silu_both(p4.y + q4.y + c14.y * g.r2 + d14.y * g.ea, a[1], f[1]);
              silu_both(p4.z + q4.z + c14.z * g.r2 + d14.z * g.ea, a[2], f[2]);
              silu_both(p4.w + q4.w + c14.w * g.r2 + d14.w * g.ea, a[3], f[3]);
              *reinterpret_cast<float4 *>(V0 + slot * H + c0) = make_float4(a[0], a[1], a[2], a[3]);  // a1
              *reinterpret_cast<float4 *>(V1 + slot * H + c0) = make_float4(f[0], f[1], f[2], f[3]);  // f1
            }
            __syncwarp();
            float Dd0 = 0.f, Dd1 = 0.f, Dd2 = 0.f, dotD = 0.f;
            if (r < kT) {
              const float4 dxj = sDXa[j * kT + r];
              Dd0 = dxi.x - dxj.x; Dd1 = dxi.y - dxj.y; Dd2 = dxi.z - dxj.z;
              dotD = g.d[0] * Dd0 + g.d[1] * Dd1 + g.d[2] * Dd2;
              const float e0a = (r % 3 == 0) ? (x0i.x - x0j.x) : ((r % 3 == 1) ? (x0i.y - x0j.y) : (x0i.z - x0j.z));
              const float sgn = (tnode == ic) ? 1.0f : ((tnode == j) ? -1.0f : 0.0f);
              const float dr2 = 2.0f * dotD, dea = 2.0f * sgn * e0a;
              float row[32];
              const float *dp = sDP + (gsel * kR + r) * kDQS;
              const float *dq = sDQ + (j * kT + r) * kDQS;
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 a = *reinterpret_cast<const float4 *>(dp + 4 * k4);
                const float4 bq = *reinterpret_cast<const float4 *>(dq + 4 * k4);
                const float4 f = *reinterpret_cast<const float4 *>(V1 + slot * H + 4 * k4);
                const float4 cc = __ldg(reinterpret_cast<const float4 *>(W1 + pk::c1 + 4 * k4));
                const float4 dd = __ldg(reinterpret_cast<const float4 *>(W1 + pk::d1 + 4 * k4));
                row[4 * k4 + 0] = f.x * (a.x + bq.x + cc.x * dr2 + dd.x * dea);
                row[4 * k4 + 1] = f.y * (a.y + bq.y + cc.y * dr2 + dd.y * dea);
                row[4 * k4 + 2] = f.z * (a.z + bq.z + cc.z * dr2 + dd.z * dea);
                row[4 * k4 + 3] = f.w * (a.w + bq.w + cc.w * dr2 + dd.w * dea);
              }
              store_operand_row<SPLIT>(sAhi, sAlo, tt, row);
            } else if (r == kPrimalRow) {
              float row[32];
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 a = *reinterpret_cast<const float4 *>(V0 + slot * H + 4 * k4);
                row[4 * k4] = a.x; row[4 * k4 + 1] = a.y; row[4 * k4 + 2] = a.z; row[4 * k4 + 3] = a.w;
              }
              store_operand_row<SPLIT>(sAhi, sAlo, tt, row);
            }
            umma::fence_proxy_async_smem();
            umma::fence_before_thread_sync();
            team_sync(team);
            if (tt == 0) {
              umma::fence_after_thread_sync();
              issue_gemm<SPLIT>(tmD1, sAhi, sAlo, sW2hi, sW2lo, mbars + 2 * team);
            }
            umma::mbar_wait(mbars + 2 * team, ph1);
            ph1 ^= 1;
            umma::fence_after_thread_sync();
            float acc[32];
            umma::tmem_ld_32x32(tmD1 + tm_lane, acc);
            // ---------- stage 2: second edge linear output -> SiLU, attention gate, gated message
            if (r == kPrimalRow) {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 b2 = __ldg(reinterpret_cast<const float4 *>(W1 + pk::b2 + 4 * k4));
                *reinterpret_cast<float4 *>(V0 + slot * H + 4 * k4) =
                    make_float4(acc[4 * k4] + b2.x, acc[4 * k4 + 1] + b2.y, acc[4 * k4 + 2] + b2.z, acc[4 * k4 + 3] + b2.w);
              }
            }
            __syncwarp();
            float s_att;
            {
              const int c0 = 4 * r;
              const float4 z = *reinterpret_cast<const float4 *>(V0 + slot * H + c0);
              const float4 wa = __ldg(reinterpret_cast<const float4 *>(W1 + pk::wa + c0));
              float m[4], f[4];
              silu_both(z.x, m[0], f[0]); silu_both(z.y, m[1], f[1]); silu_both(z.z, m[2], f[2]); silu_both(z.w, m[3], f[3]);
              const float part = wa.x * m[0] + wa.y * m[1] + wa.z * m[2] + wa.w * m[3];
              s_att = sigmoidf_fast(group8_sum(part) + __ldg(W1 + pk::ba));
              __syncwarp();  // every lane has read its z before V0 is overwritten with ms
              *reinterpret_cast<float4 *>(V0 + slot * H + c0) = make_float4(m[0] * s_att, m[1] * s_att, m[2] * s_att, m[3] * s_att);  // ms
              *reinterpret_cast<float4 *>(V1 + slot * H + c0) = make_float4(m[0], m[1], m[2], m[3]);                                      // m
              *reinterpret_cast<float4 *>(V2 + slot * H + c0) = make_float4(f[0] * s_att, f[1] * s_att, f[2] * s_att, f[3] * s_att);  // f2*s
              *reinterpret_cast<float4 *>(V3 + slot * H + c0) = make_float4(wa.x * f[0], wa.y * f[1], wa.z * f[2], wa.w * f[3]);      // wa*f2
            }
            __syncwarp();
            if (r < kT) {
              float dsd = 0.f;
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 w = *reinterpret_cast<const float4 *>(V3 + slot * H + 4 * k4);
                dsd += w.x * acc[4 * k4] + w.y * acc[4 * k4 + 1] + w.z * acc[4 * k4 + 2] + w.w * acc[4 * k4 + 3];
              }
              const float ds = s_att * (1.0f - s_att) * dsd;
              const float vmask = valid ? 1.0f : 0.0f;
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 fs = *reinterpret_cast<const float4 *>(V2 + slot * H + 4 * k4);
                const float4 m = *reinterpret_cast<const float4 *>(V1 + slot * H + 4 * k4);
                acc[4 * k4 + 0] = fmaf(acc[4 * k4 + 0], fs.x, m.x * ds);
                acc[4 * k4 + 1] = fmaf(acc[4 * k4 + 1], fs.y, m.y * ds);
                acc[4 * k4 + 2] = fmaf(acc[4 * k4 + 2], fs.z, m.z * ds);
                acc[4 * k4 + 3] = fmaf(acc[4 * k4 + 3], fs.w, m.w * ds);
              }
#pragma unroll
              for (int k = 0; k < 32; ++k) aggacc[k] = fmaf(acc[k], vmask, aggacc[k]);
              store_operand_row<SPLIT>(sAhi, sAlo, tt, acc);  // dms row
            } else if (r == kPrimalRow) {
              float row[32];
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 a = *reinterpret_cast<const float4 *>(V0 + slot * H + 4 * k4);
                row[4 * k4] = a.x; row[4 * k4 + 1] = a.y; row[4 * k4 + 2] = a.z; row[4 * k4 + 3] = a.w;
              }
              store_operand_row<SPLIT>(sAhi, sAlo, tt, row);  // ms row
            }
            umma::fence_proxy_async_smem();
            umma::fence_before_thread_sync();
            team_sync(team);
            if (tt == 0) {
              umma::fence_after_thread_sync();
              issue_gemm<SPLIT>(tmD2, sAhi, sAlo, sWc1hi, sWc1lo, mbars + 2 * team + 1);
            }
            umma::mbar_wait(mbars + 2 * team + 1, ph2);
            ph2 ^= 1;
            umma::fence_after_thread_sync();
            umma::tmem_ld_32x32(tmD2 + tm_lane, acc);
            // ---------- stage 3: coordinate MLP head
            if (r == kPrimalRow) {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 bc = __ldg(reinterpret_cast<const float4 *>(W1 + pk::bc1 + 4 * k4));
                *reinterpret_cast<float4 *>(V0 + slot * H + 4 * k4) =
                    make_float4(acc[4 * k4] + bc.x, acc[4 * k4 + 1] + bc.y, acc[4 * k4 + 2] + bc.z, acc[4 * k4 + 3] + bc.w);
              }
            }
            __syncwarp();
            float phi, dphi_du;
            {
              const int c0 = 4 * r;
              const float4 z = *reinterpret_cast<const float4 *>(V0 + slot * H + c0);
              const float4 wc = __ldg(reinterpret_cast<const float4 *>(W1 + pk::wc2 + c0));
              float a[4], f[4];
              silu_both(z.x, a[0], f[0]); silu_both(z.y, a[1], f[1]); silu_both(z.z, a[2], f[2]); silu_both(z.w, a[3], f[3]);
              const float u = group8_sum(wc.x * a[0] + wc.y * a[1] + wc.z * a[2] + wc.w * a[3]);
              const float th = tanhf(u);
              phi = th * rng;
              dphi_du = rng * (1.0f - th * th);
              *reinterpret_cast<float4 *>(V2 + slot * H + c0) = make_float4(wc.x * f[0], wc.y * f[1], wc.z * f[2], wc.w * f[3]);  // wc2*fc
            }
            __syncwarp();
            if (r < kT) {
              float du = 0.f;
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 w = *reinterpret_cast<const float4 *>(V2 + slot * H + 4 * k4);
                du += w.x * acc[4 * k4] + w.y * acc[4 * k4 + 1] + w.z * acc[4 * k4 + 2] + w.w * acc[4 * k4 + 3];
              }
              const float dphi = dphi_du * du;
              const float c = dotD * g.inv * g.inv / g.nrm;
              const float vmask = valid ? 1.0f : 0.0f;
              dxacc0 += vmask * ((Dd0 * g.inv - g.d[0] * c) * phi + g.d[0] * g.inv * dphi);
              dxacc1 += vmask * ((Dd1 * g.inv - g.d[1] * c) * phi + g.d[1] * g.inv * dphi);
              dxacc2 += vmask * ((Dd2 * g.inv - g.d[2] * c) * phi + g.d[2] * g.inv * dphi);
            }
            __syncwarp();  // V0..V3 are rewritten by the next tile's stage 1
          }
          // ---------- group epilogue: reduce over the SR sender slots (fixed order), node update, publish
          float *red = sAhi;  // [128][kDQS]: per-row partial aggregates (+ dx in columns 32..34)
          team_sync(team);    // the last GEMM has been consumed by every thread of the team
          if (r < kT) {
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4)
              *reinterpret_cast<float4 *>(red + tt * kDQS + 4 * k4) =
                  make_float4(aggacc[4 * k4], aggacc[4 * k4 + 1], aggacc[4 * k4 + 2], aggacc[4 * k4 + 3]);
            *reinterpret_cast<float4 *>(red + tt * kDQS + 32) = make_float4(dxacc0, dxacc1, dxacc2, 0.f);
          }
          team_sync(team);
          // SIMT, lane = channel: the team's 4 warps share the G*kT (receiver, direction) rows
          for (int row = (warp & 3); row < G * kT; row += 4) {
            const int gi = row / kT, t = row % kT;
            const int node = grp * G + gi;
            if (node >= NP) continue;
            float dagg = 0.f, dx = 0.f;
#pragma unroll 1
            for (int sidx = 0; sidx < SR; ++sidx) {
              const float *src = red + (((gi * SR + sidx) * kR) + t) * kDQS;
              dagg += src[lane];
              if (lane < 3) dx += src[32 + lane];
            }
            // d x^2 = d x^1 + sum_j d trans
            if (lane < 3) reinterpret_cast<float *>(sDXb)[(node * kT + t) * 4 + lane] = reinterpret_cast<const float *>(sDXa)[(node * kT + t) * 4 + lane] + dx;
            if (lane == 3) reinterpret_cast<float *>(sDXb)[(node * kT + t) * 4 + 3] = 0.f;
            // node update: dz3 = W3h dh^1 + W3a dagg ; dh^2 = dh^1 + W4 (f3 * dz3)
            float wr[H];
            float *stg = st.pa;
            const float dh1 = gX[(node * kT + t) * H + lane];
            __syncwarp();
            stg[lane] = dh1;
            stg[H + lane] = dagg;
            __syncwarp();
            load_row(wr, W1 + pk::W3h_f, lane);
            float dz3 = dot32(wr, stg);
            load_row(wr, W1 + pk::W3a_f, lane);
            dz3 += dot32(wr, stg + H);
            float v, f3;
            silu_both(sZ3[(NP + node) * H + lane], v, f3);
            __syncwarp();
            stg[lane] = f3 * dz3;
            __syncwarp();
            load_row(wr, W1 + pk::W4_f, lane);
            const float dh2 = dh1 + dot32(wr, stg);
            __syncwarp();
            stg[lane] = dh2;
            __syncwarp();
            load_row(wr, W2l + pk::B_f, lane);
            gX[(node * kT + t) * H + lane] = dot32(wr, stg);  // B^2 dh^2 (overwrites dh^1 of this row)
            const int ii = node - k0;
            if (ii >= 0 && ii < kTN && t / 3 == ii) {
              load_row(wr, W2l + pk::A_f, lane);
              sDP2[(ii * 3 + t % 3) * H + lane] = dot32(wr, stg);
            }
          }
          team_sync(team);  // red / staging are free again before the next group's operand rows
        }
      }
      __syncthreads();
      // =========================== phase C (SIMT): last layer, receivers = tangent nodes, own directions only
      {
        float w2[H], wc1[H];
        load_row(w2, W2l + pk::W2_f, lane);
        load_row(wc1, W2l + pk::Wc1_f, lane);
        const EdgeScal sc = load_edge_scal(W2l, lane);
        for (int j = warp; j < NP; j += NW) {
#pragma unroll 1
          for (int kk = 0; kk < kTN; ++kk) {
            const int k = k0 + kk;
            if (k >= NP || k == j) continue;
            const float4 xk = sX[2 * NP + k], xj = sX[2 * NP + j], x0k = sX[k], x0j = sX[j];
            const EdgeGeo g = edge_geo(xk, xj, x0k, x0j);
            EdgeT<3> tin;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const int t = kk * 3 + a;
              tin.dpq[a] = sDP2[(kk * 3 + a) * H + lane] + gX[(j * kT + t) * H + lane];
              const float4 qk = sDXb[k * kT + t], qj = sDXb[j * kT + t];
              tin.Dd[a][0] = qk.x - qj.x; tin.Dd[a][1] = qk.y - qj.y; tin.Dd[a][2] = qk.z - qj.z;
            }
            tin.dea[0] = 2.0f * (x0k.x - x0j.x); tin.dea[1] = 2.0f * (x0k.y - x0j.y); tin.dea[2] = 2.0f * (x0k.z - x0j.z);
            float dms[3], dtr[3][3];
            edge_eval<3, kT>(w2, wc1, sc, rng, sP[(2 * NP + k) * H + lane] + sQ[(2 * NP + j) * H + lane], g, st, lane, tin, dms, dtr);
            trace += dtr[0][0] + dtr[1][1] + dtr[2][2];
          }
          const int jj = j - k0;
          if (jj >= 0 && jj < kTN) {
            const float4 q0 = sDXb[j * kT + jj * 3 + 0], q1 = sDXb[j * kT + jj * 3 + 1], q2 = sDXb[j * kT + jj * 3 + 2];
            trace += q0.x + q1.y + q2.z;
          }
        }
      }
      __syncthreads();
    }
    if (lane == 0) sRed[8 + warp] = trace;
    __syncthreads();
    if (tid == 0) {
      float tr = 0.f;
      for (int w = 0; w < NW; ++w) tr += sRed[8 + w];
      const float Dn = (float)(3 * NP);
      divergence[b] = ((c_s - 1.0f) * Dn + c_out * c_in * (tr - Dn)) / h;
    }
    __syncthreads();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<NTEAM * 64>(tmem_base);
}

template <int NP, int NTEAM, int G, bool SPLIT>
static int launch_mma(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *sc, float *dv,
                      float *scratch, int64_t scratch_bytes, cudaStream_t s) {
  using M = MPlan<NP, NTEAM, G>;
  const unsigned grid = (unsigned)(B < kNumSMs ? B : kNumSMs);
  PITA_REQUIRE(dv == nullptr || (scratch != nullptr && scratch_bytes >= (int64_t)kNumSMs * M::kXFloats * 4), PITA_EINVAL,
               "egnn_score_div: workspace too small (need %lld bytes)", (long long)kNumSMs * M::kXFloats * 4);
  auto k = egnn_score_div_mma_kernel<NP, NTEAM, G, SPLIT>;
  PITA_REQUIRE(M::kBytes <= 227 * 1024, PITA_EUNSUP, "egnn_score_div_mma_kernel needs %zu bytes of shared memory", M::kBytes);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M::kBytes);
  if (e != cudaSuccess) { set_error("egnn_score_div_mma_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  k<<<grid, NTEAM * 128, M::kBytes, s>>>(w, ht, x, beta, B, sc, dv, scratch);
  PITA_CHECK_LAUNCH("egnn_score_div_mma_kernel");
  return PITA_OK;
}

int64_t score_div_mma_workspace_bytes(int n) {
  if (n == 13) return (int64_t)kNumSMs * MPlan<13, 2, 4>::kXFloats * 4;
  if (n == 55) return (int64_t)kNumSMs * MPlan<55, 2, 2>::kXFloats * 4;
  return -1;
}

int launch_score_div_mma(int n, bool split, const float *w, const float *ht, const float *x, const float *beta, int64_t B,
                         float *sc, float *dv, float *scratch, int64_t scratch_bytes, cudaStream_t s) {
  if (n == 13)
    return split ? launch_mma<13, 2, 4, true>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s)
                 : launch_mma<13, 2, 4, false>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s);
  return split ? launch_mma<55, 2, 2, true>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s)
               : launch_mma<55, 2, 2, false>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s);
}

}  // namespace pita
