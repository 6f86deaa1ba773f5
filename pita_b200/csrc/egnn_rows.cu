// EGNN denoiser on the row-batched tcgen05 engine (rowgemm.cuh): plain forward (EGNN_dynamics.forward,
// egnn_temp_conditioned.py:56-93) and score + exact divergence (score_net.py:13-43, utils.py:30-51).
// See rowgemm.cuh for the thread = (particle, node) row mapping; the derivative algebra follows
// oracle/egnn_analytic.py (checked there against autograd on the CPU).
#include "egnn_rowops.cuh"

namespace pita {
namespace rg {

// TMEM column slots (32 columns each) of one team
enum Slot { sAcc0 = 0, sAccC = 1, sT0 = 2, sT1 = 3, sT2 = 4, sP = 5, sPA = 6, sH = 7, kSlots = 8 };
// energy (reverse) kernel: the tangent accumulators sT1/sT2/sPA are free and hold per-row state of the reverse pass
enum RevSlot { sGAgg = sT0, sF30 = sT1, sF31 = sT2, sH1 = sPA, sGH = sH };
// weight slots
enum WSlot { wA = 0, wB = 1, wW2 = 2, wWc1 = 3, wW3a = 4 };

template <int NP>
struct Shape {
  static constexpr int PB = kRows / NP;  // particles per team tile
  static constexpr int R = PB * NP;      // active rows
};

// ---- shared-memory plan -------------------------------------------------------------------------
// MODE: 0 = plain forward, 1 = score + divergence (tangent regions), 2 = energy (reverse-pass scatter regions)
constexpr int kModeFwd = 0, kModeTan = 1, kModeRev = 2;
template <int NP, int NTEAM, bool SPLIT, int MODE>
struct Smem {
  static constexpr bool TANGENT = (MODE == kModeTan);
  static constexpr int PB = Shape<NP>::PB;
  // byte offsets from a 1024-aligned base
  static constexpr size_t oW = 0;
  static constexpr size_t oA = oW + (size_t)kWSlots * Bytes<SPLIT>::kW;
  static constexpr size_t oF = oA + (size_t)NTEAM * Bytes<SPLIT>::kA;  // float regions start here
  // CTA-wide float regions (float index relative to oF)
  static constexpr int fVec = 0;                            // [3][kNumVec][32]
  static constexpr int fEmb = fVec + 3 * kNumVec * 32;      // [3][32] embedding e0 e1 eb
  static constexpr int fCls = fEmb + 96;                    // layer-0 class tables: P0c[3][32], Q0c[3][32]
  static constexpr int fMisc = fCls + 192;                  // mbarriers [NTEAM] (8 B each) + tmem slot
  static constexpr int fTeam = fMisc + 2 * NTEAM + 8;
  // per-team float regions
  static constexpr int tQa = 0;                             // [128][32] swizzled: Q rows of layer 1 (persist in tangent passes)
  static constexpr int tQb = tQa + kRows * 32;              // [128][32] swizzled: Q rows of layers 0 / 2; sender tangent base vectors
  static constexpr int tX = tQb + kRows * 32;               // [3 or 4][128] float4 (x^3 aliases tDX in the tangent kernel)
  static constexpr int tDX = tX + (TANGENT ? 3 : 4) * kRows * 4;  // [128][3] float4   sender-side coordinate tangents
  static constexpr int tX3 = TANGENT ? tDX : tX + 3 * kRows * 4;
  static constexpr int tCoef = tDX + (TANGENT ? kRows * 12 : 0);  // [128] float4  sender-side coefficients
  static constexpr int tOwnA = tCoef + (TANGENT ? kRows * 4 : 0); // [PB][3][32]   A1 dh1 of the tangent node (own dirs)
  static constexpr int tKdP2 = tOwnA;                             // [PB][3][32]   A2 dh2 of the tangent node (layer-2 stage only)
  static constexpr int tOwnB = tOwnA + (TANGENT ? PB * 96 : 0);   // [PB][3][32]   B1 dh1 of the tangent node
  static constexpr int tKP2 = tOwnB + (TANGENT ? PB * 96 : 0);    // [PB][32]      P2 of the tangent node
  static constexpr int tRed = tKP2 + (TANGENT ? PB * 32 : 0);     // [128]         per-particle reductions
  static constexpr int tGX = tRed + kRows;                        // MODE 2: [128] float4 scatter target, coordinate cotangents
  static constexpr int tGX0 = tGX + (MODE == kModeRev ? kRows * 4 : 0);   // [128] float4 scatter target, edge_attr path
  static constexpr int tMean = tGX0 + (MODE == kModeRev ? kRows * 4 : 0); // [128] float4 scratch of particle_mean
  static constexpr int kTeamFloats = ((tMean + (MODE == kModeRev ? kRows * 4 : 0) + 3) / 4) * 4;
  static constexpr size_t kBytes = 1024 + oF + (size_t)(fTeam + NTEAM * kTeamFloats) * 4;
};


// ---- common context --------------------------------------------------------------------------------
template <int NP, int NTEAM, bool SPLIT, int MODE>
struct Ctx {
  using S = Smem<NP, NTEAM, SPLIT, MODE>;
  static constexpr int PB = Shape<NP>::PB;
  static constexpr int R = Shape<NP>::R;
  Team<SPLIT> T;
  float *wsm, *sVec, *sEmb, *sCls, *tm;  // CTA regions and this team's float region
  float4 *sX, *sX3;
  float *sQa, *sQb;
  int tid, team, p, i;
  bool row_ok;  // tt < R

  __device__ __forceinline__ const float *vec(int l) const { return sVec + l * kNumVec * 32; }
  __device__ __forceinline__ int sender(int u) const { int j = i + 1 + u; return j >= NP ? j - NP : j; }
};

// One-time CTA setup: carve shared memory, TMEM, mbarriers, per-layer vectors.
template <int NP, int NTEAM, bool SPLIT, int MODE>
__device__ __forceinline__ void setup(Ctx<NP, NTEAM, SPLIT, MODE> &c, float *sm_raw, const float *__restrict__ wpack,
                                      uint32_t &tmem_base_out) {
  using S = Smem<NP, NTEAM, SPLIT, MODE>;
  uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  c.tid = threadIdx.x;
  c.team = c.tid >> 7;
  const int tt = c.tid & 127, warp = c.tid >> 5;
  c.wsm = reinterpret_cast<float *>(base + S::oW);
  float *fl = reinterpret_cast<float *>(base + S::oF);
  c.sVec = fl + S::fVec;
  c.sEmb = fl + S::fEmb;
  c.sCls = fl + S::fCls;
  uint64_t *mbars = reinterpret_cast<uint64_t *>(fl + S::fMisc);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fl + S::fMisc + 2 * NTEAM);
  c.tm = fl + S::fTeam + c.team * S::kTeamFloats;
  c.sQa = c.tm + S::tQa;
  c.sQb = c.tm + S::tQb;
  c.sX = reinterpret_cast<float4 *>(c.tm + S::tX);
  c.sX3 = reinterpret_cast<float4 *>(c.tm + S::tX3);
  c.p = tt / NP;
  c.i = tt - c.p * NP;
  c.row_ok = tt < Shape<NP>::R;

  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  if (c.tid == 0) {
    for (int k = 0; k < NTEAM; ++k) umma::mbar_init(mbars + k, 1);
    umma::fence_mbar_init();
  }
  for (int k = c.tid; k < 3 * kNumVec * 32; k += NTEAM * 128) {
    const int l = k / (kNumVec * 32), r = k % (kNumVec * 32);
    c.sVec[k] = __ldg(wpack + pk::kHeader + l * pk::kLayer + pk::c1 + r);
  }
  for (int k = c.tid; k < 96; k += NTEAM * 128) c.sEmb[k] = __ldg(wpack + k);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = uniform32(*tmem_slot);
  tmem_base_out = tmem_base;
  const uint32_t team_u = uniform32((uint32_t)c.team);
  c.T.a_hi = reinterpret_cast<float *>(base + S::oA + (size_t)c.team * Bytes<SPLIT>::kA);
  c.T.a_addr = uniform32(umma::smem_u32(base + S::oA)) + team_u * (uint32_t)Bytes<SPLIT>::kA;
  c.T.w_addr = uniform32(umma::smem_u32(c.wsm));
  c.T.mbar_addr = uniform32(umma::smem_u32(mbars)) + team_u * 8u;
  c.T.phase = 0;
  c.T.tmem_col = tmem_base + team_u * (uint32_t)(512 / NTEAM);
  c.T.tmem = c.T.tmem_col + (((uint32_t)((warp & 3) * 32)) << 16);
  c.T.bar_id = 1 + c.team;
  c.T.tt = tt;
  c.T.issuer = uniform32((uint32_t)(warp & 3)) == 0u;
}

template <int NP, int NTEAM, bool SPLIT, int MODE>
__device__ __forceinline__ void load_weights(Ctx<NP, NTEAM, SPLIT, MODE> &c, const WeightSrc &src) {
  __syncthreads();  // every team is done with the MMAs that read the old tiles
  load_weight_tiles<SPLIT>(c.wsm, src, c.tid, NTEAM * 128);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
}

__device__ __forceinline__ WeightSrc layer_edge_set(const float *Wl) {
  WeightSrc s;
  s.p[wA] = Wl + pk::A_b; s.p[wB] = Wl + pk::B_b; s.p[wW2] = Wl + pk::W2_b; s.p[wWc1] = Wl + pk::Wc1_b; s.p[wW3a] = Wl + pk::W3a_b;
  s.count = 5;
  return s;
}


// Primal forward of the three layers for the team's PB particles.
//   in : sX[0] = network input coordinates (published, team-synced by the caller), tcond/beta of the thread's particle
//   out: sX[1..3]; TMEM sP = P^1 (+b1) if KEEP_L1; sQa = Q^1.  When `scratch` != nullptr the per-row vectors the
//        tangent passes need later are written there: f3^0, f3^1 (silu'(z3)), P^2 (+b1), Q^2   (own-row layout).
template <int NP, int NTEAM, bool SPLIT, int MODE>
__device__ __forceinline__ void primal_forward_rows(Ctx<NP, NTEAM, SPLIT, MODE> &c, const float *__restrict__ wpack,
                                                    float tcond, float beta, bool team_active, float *scratch_row) {
  constexpr bool TANGENT = (MODE == kModeTan);
  constexpr int L = 3;
  const float rng = kCoordsRange / (float)L;
  Team<SPLIT> &T = c.T;
  const int tt = T.tt;
  float row[32];
  float dummy[32];
  if (team_active) {
    embed<NP>(row, c.sEmb, c.i, tcond, beta);
    T.st(sH, row);
  }
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const float *Wl = wpack + pk::kHeader + l * pk::kLayer;
    const float *vec = c.vec(l);
    load_weights(c, layer_edge_set(Wl));
    float *sQ = (l == 1) ? c.sQa : c.sQb;
    if (team_active) {
      // ---- node products P = A h + b1, Q = B h
      T.ld(sH, row);
      T.store_row(row);
      T.round_trip([&] { T.mma(sP, wA, false); T.mma(sAcc0, wB, false); });
      T.ld(sAcc0, row);
      qrow_store(sQ, tt, row);
      T.ld(sP, row);
      add_vec(row, vec + vB1 * 32);
      T.st(sP, row);
      if (TANGENT && scratch_row) {
        if (l == 1) store_vec_global(scratch_row + 4 * kRows * 32, row);  // P^1
        if (l == 2) {
          store_vec_global(scratch_row + 2 * kRows * 32, row);  // P^2
          float q[32];
          qrow_load(sQ, tt, q);
          store_vec_global(scratch_row + 3 * kRows * 32, q);    // Q^2
        }
      }
      umma::fence_before_thread_sync();
      T.sync();  // Q rows and TMEM stores visible before the gathers / next MMAs
      umma::fence_after_thread_sync();
      // ---- edges
      const float4 xi = c.sX[l * kRows + tt], x0i = c.sX[tt];
      float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
      // agg_i = sum_j m*_ij is summed in registers (round-to-nearest fp32) and goes through W3a ONCE after the slots:
      // accumulating W3a m*_ij inside TMEM over n-1 slots would round the growing sum 12 x (n-1) times with the tensor
      // core's truncating fp32 add (measured: 3.6e-4 score error at n = 55) and costs 12 more MMAs per edge.
      float agg[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) agg[k] = 0.f;
#pragma unroll 1
      for (int u = 0; u < NP - 1; ++u) {
        const int rj = c.p * NP + c.sender(u);
        const Geo g = edge_geo4(xi, c.sX[l * kRows + rj], x0i, c.sX[rj]);
        T.ld(sP, row);
        stage1<false>(row, dummy, sQ, rj, vec, g.r2, g.ea);
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, wW2, false); });
        T.ld(sAcc0, row);
        stage2<false>(row, dummy, dummy, vec);
        T.store_row(row);
        if (l < L - 1) {
#pragma unroll
          for (int k = 0; k < 32; ++k) agg[k] += row[k];
        }
        T.round_trip([&] { T.mma(sAccC, wWc1, false); });
        T.ld(sAccC, row);
        const float th = stage3<false>(row, dummy, vec);
        const float f = g.inv * th * rng;
        dx0 = fmaf(g.d0, f, dx0); dx1 = fmaf(g.d1, f, dx1); dx2 = fmaf(g.d2, f, dx2);
      }
      (l == L - 1 ? c.sX3 : c.sX + (l + 1) * kRows)[tt] = make_float4(xi.x + dx0, xi.y + dx1, xi.z + dx2, 0.f);
      if (l < L - 1) {
        T.store_row(agg);
        T.round_trip([&] { T.mma(sT0, wW3a, false); });
      }
    }
    if (l < L - 1) {
      // ---- node update h += W4 silu(W3h h + W3a agg + b3) + b4
      WeightSrc s2;
      s2.p[0] = Wl + pk::W3h_b;
      s2.p[1] = Wl + pk::W4_b;
      s2.count = 2;
      load_weights(c, s2);
      if (team_active) {
        T.ld(sH, row);
        T.store_row(row);
        T.round_trip([&] { T.mma(sT0, 0, true); });
        T.ld(sT0, row);
        add_vec(row, vec + vB3 * 32);
        if (TANGENT && scratch_row) {
          float f3[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) { float a; silu_both(row[k], a, f3[k]); row[k] = a; }
          store_vec_global(scratch_row + l * kRows * 32, f3);  // f3^l
        } else if (MODE == kModeRev) {  // the reverse pass keeps f3^l in its own TMEM lane
          float f3[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) { float a; silu_both(row[k], a, f3[k]); row[k] = a; }
          T.st(sF30 + l, f3);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] = silu_val(row[k]);
        }
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, 1, false); });
        float hh[32];
        T.ld(sH, hh);
        T.ld(sAcc0, row);
        add_vec(row, vec + vB4 * 32);
#pragma unroll
        for (int k = 0; k < 32; ++k) hh[k] += row[k];
        T.st(sH, hh);
        if (MODE == kModeRev && l == 0) T.st(sH1, hh);  // h^1 is needed again to rebuild P^1, Q^1
      }
    }
  }
  if (team_active) {
    umma::fence_before_thread_sync();
    T.sync();
    umma::fence_after_thread_sync();
  }
}

// Per-particle mean over the NP rows of a particle of a float4 held by each row (team-wide helper).
// `buf` is a [128] float4 scratch in shared memory; all threads of the team call.
template <int NP, bool SPLIT>
__device__ __forceinline__ float4 particle_mean(const Team<SPLIT> &T, float4 *buf, float4 v, int p) {
  T.sync();
  buf[T.tt] = v;
  T.sync();
  float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll 1
  for (int k = 0; k < NP; ++k) {
    const float4 q = buf[p * NP + k];
    a += q.x; b += q.y; cc += q.z;
  }
  return make_float4(a / NP, b / NP, cc / NP, 0.f);
}

// sum over the NP rows of the thread's particle of a per-row scalar (all threads of the team call)
template <int NP, bool SPLIT>
__device__ __forceinline__ float particle_sum(const Team<SPLIT> &T, float *buf, float v, int p) {
  T.sync();
  buf[T.tt] = v;
  T.sync();
  float a = 0.f;
#pragma unroll 1
  for (int k = 0; k < NP; ++k) a += buf[p * NP + k];
  return a;
}

// ================================================================================================
// Kernel: plain forward   vel = EGNN_dynamics(tcond, y, beta)
// ================================================================================================
template <int NP, int NTEAM, bool SPLIT>
__global__ void __launch_bounds__(NTEAM * 128, 1)
egnn_forward_rows_kernel(const float *__restrict__ wpack, const float *__restrict__ tcond, const float *__restrict__ y,
                         const float *__restrict__ beta, int64_t B, float *__restrict__ vel) {
  extern __shared__ __align__(16) float sm_raw[];
  using C = Ctx<NP, NTEAM, SPLIT, kModeFwd>;
  C c;
  uint32_t tmem_base;
  setup(c, sm_raw, wpack, tmem_base);
  constexpr int PB = C::PB;
  const int64_t nbatch = (B + (int64_t)NTEAM * PB - 1) / ((int64_t)NTEAM * PB);
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int64_t p0 = (batch * NTEAM + c.team) * PB;
    const bool team_active = p0 < B;
    const int64_t part = p0 + c.p;
    const bool ok = c.row_ok && part < B;
    const int64_t pc = ok ? part : (B - 1);
    const int ic = c.row_ok ? c.i : 0;
    const float tc = __ldg(tcond + pc), be = __ldg(beta + pc);
    const float *src = y + pc * 3 * NP + 3 * ic;
    const float4 y0 = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
    c.sX[c.T.tt] = y0;
    primal_forward_rows(c, wpack, tc, be, team_active, nullptr);
    if (team_active) {
      const float4 xl = c.sX3[c.T.tt];
      const float4 v = make_float4(xl.x - y0.x, xl.y - y0.y, xl.z - y0.z, 0.f);
      const float4 mean = particle_mean<NP, SPLIT>(c.T, reinterpret_cast<float4 *>(c.sQb), v, c.row_ok ? c.p : 0);
      if (ok) {
        float *dst = vel + part * 3 * NP + 3 * c.i;
        dst[0] = v.x - mean.x; dst[1] = v.y - mean.y; dst[2] = v.z - mean.z;
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if ((c.tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

template <int NP, int NTEAM, bool SPLIT>
static int launch_forward_rows(const float *w, const float *tc, const float *y, const float *beta, int64_t B, float *vel,
                               cudaStream_t s) {
  using S = Smem<NP, NTEAM, SPLIT, kModeFwd>;
  auto k = egnn_forward_rows_kernel<NP, NTEAM, SPLIT>;
  static_assert(S::kBytes <= 227 * 1024, "shared memory plan exceeds 227 KB");
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
  if (e != cudaSuccess) { set_error("egnn_forward_rows_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  const int64_t nbatch = (B + (int64_t)NTEAM * Shape<NP>::PB - 1) / ((int64_t)NTEAM * Shape<NP>::PB);
  const unsigned grid = (unsigned)(nbatch < kNumSMs ? nbatch : kNumSMs);
  k<<<grid, NTEAM * 128, S::kBytes, s>>>(w, tc, y, beta, B, vel);
  PITA_CHECK_LAUNCH("egnn_forward_rows_kernel");
  return PITA_OK;
}


// ================================================================================================
// Kernel: energy net — E, grad_x E, dE/dh (energy_net.py:14-62) by a hand-derived reverse pass on the row engine.
// Algebra: oracle/egnn_analytic.py::u_theta_backward / energy_terms.
//   forward : primal_forward_rows keeps x^0..x^3 (shared), f3^0, f3^1, h^1 (TMEM lanes), P^2 (TMEM) and Q^2 (shared).
//   reverse : per layer l = 2,1,0 and sender slot u the edge (i, j) is re-evaluated (2 MMAs), its cotangent goes
//             back through Wc1^T and W2^T (2 MMAs); receiver-side sums stay in the thread, sender-side sums
//             (B^T path, coordinates) are scattered to the sender's shared-memory row — a slot is a bijection
//             i -> j inside every particle, so the read-modify-write needs no atomics and is deterministic.
// ================================================================================================
template <int NP, int NTEAM, bool SPLIT>
__global__ void __launch_bounds__(NTEAM * 128, 1)
egnn_energy_rows_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                        const float *__restrict__ beta_in, int64_t B, float *__restrict__ energy,
                        float *__restrict__ grad_x, float *__restrict__ dE_dh) {
  extern __shared__ __align__(16) float sm_raw[];
  using C = Ctx<NP, NTEAM, SPLIT, kModeRev>;
  using S = Smem<NP, NTEAM, SPLIT, kModeRev>;
  constexpr int PB = C::PB;
  constexpr int L = 3;
  const float rng = kCoordsRange / (float)L;
  C c;
  uint32_t tmem_base;
  setup(c, sm_raw, wpack, tmem_base);
  Team<SPLIT> &T = c.T;
  const int tt = T.tt;
  float *sGQ = c.sQa;  // free once the forward is done (Q^1 is rebuilt from h^1)
  float4 *sGX = reinterpret_cast<float4 *>(c.tm + S::tGX);
  float4 *sGX0 = reinterpret_cast<float4 *>(c.tm + S::tGX0);
  float4 *sMean = reinterpret_cast<float4 *>(c.tm + S::tMean);
  float *sRed = c.tm + S::tRed;
  const bool want_grad = (grad_x != nullptr) || (dE_dh != nullptr);

  const int64_t nbatch = (B + (int64_t)NTEAM * PB - 1) / ((int64_t)NTEAM * PB);
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int64_t p0 = (batch * NTEAM + c.team) * PB;
    const bool team_active = p0 < B;
    const int64_t part = p0 + c.p;
    const bool ok = c.row_ok && part < B;
    const int64_t pc = ok ? part : (B - 1);
    const int ic = c.row_ok ? c.i : 0;
    const int pp = c.row_ok ? c.p : 0;
    const float h = __ldg(ht + pc), beta = __ldg(beta_in + pc);
    const float c_in = rsqrtf(1.0f + h), c_noise = 0.125f * logf(h), rs_h = rsqrtf(h);
    const float *src = x + pc * 3 * NP + 3 * ic;
    const float xr[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    const float4 yi = make_float4(c_in * xr[0], c_in * xr[1], c_in * xr[2], 0.f);
    c.sX[tt] = yi;
    primal_forward_rows(c, wpack, c_noise, beta, team_active, nullptr);

    // ---- U = <remove_mean(x^3 - y), y>,  E = |x|^2 / (2 (1 + h)) - U / sqrt(h)        (energy_net.py:29-39)
    float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), vmean = vi, ymean = vi;
    float U = 0.f, x2 = 0.f;
    if (team_active) {
      const float4 xl = c.sX3[tt];
      vi = make_float4(xl.x - yi.x, xl.y - yi.y, xl.z - yi.z, 0.f);
      vmean = particle_mean<NP, SPLIT>(T, sMean, vi, pp);
      ymean = particle_mean<NP, SPLIT>(T, sMean, yi, pp);
      const float ui = c.row_ok ? (vi.x - vmean.x) * yi.x + (vi.y - vmean.y) * yi.y + (vi.z - vmean.z) * yi.z : 0.f;
      const float y2i = c.row_ok ? yi.x * yi.x + yi.y * yi.y + yi.z * yi.z : 0.f;
      U = particle_sum<NP, SPLIT>(T, sRed, ui, pp);
      x2 = particle_sum<NP, SPLIT>(T, sRed, y2i, pp) * (1.0f + h);  // |x|^2 = |y|^2 / c_in^2
      if (ok && c.i == 0) energy[part] = x2 / (2.0f * (1.0f + h)) - rs_h * U;
    }
    if (!want_grad) continue;

    // ---- reverse pass.  Cotangent on x^3 is w_i = y_i - mean(y); on h^3 it is 0.
    float gx[3] = {yi.x - ymean.x, yi.y - ymean.y, yi.z - ymean.z};
    float g0[3] = {0.f, 0.f, 0.f};  // own part of the cotangent that reaches y through edge_attr
    float row[32], f1[32], m[32], f2[32], fc[32], gp[32];
    if (team_active) {
#pragma unroll
      for (int k = 0; k < 32; ++k) row[k] = 0.f;
      T.st(sGH, row);
      T.st(sGAgg, row);
      qrow_store(sGQ, tt, row);
      sGX[tt] = make_float4(0.f, 0.f, 0.f, 0.f);
      sGX0[tt] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      const float *Wl = wpack + pk::kHeader + l * pk::kLayer;
      const float *vec = c.vec(l);
      if (l < L - 1) {
        // ---- node-MLP reverse: gz3 = f3 * (W4^T gh);  gh += W3h^T gz3;  gagg = W3a^T gz3;  then rebuild P^l, Q^l
        WeightSrc sN;
        sN.p[0] = Wl + pk::W4_f; sN.p[1] = Wl + pk::W3h_f; sN.p[2] = Wl + pk::W3a_f; sN.p[3] = Wl + pk::A_b; sN.p[4] = Wl + pk::B_b;
        sN.count = 5;
        load_weights(c, sN);
        if (team_active) {
          T.ld(sGH, row);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 0, false); });
          T.ld(sAcc0, row);
          T.ld(sF30 + l, f1);
#pragma unroll
          for (int k = 0; k < 32; ++k) row[k] *= f1[k];
          T.store_row(row);
          T.round_trip([&] { T.mma(sGH, 1, true); T.mma(sGAgg, 2, false); });
          if (l == 1) T.ld(sH1, row);
          else embed<NP>(row, c.sEmb, c.i, c_noise, beta);
          T.store_row(row);
          T.round_trip([&] { T.mma(sP, 3, false); T.mma(sAcc0, 4, false); });
          T.ld(sAcc0, row);
          qrow_store(c.sQb, tt, row);
          T.ld(sP, row);
          add_vec(row, vec + vB1 * 32);
          T.st(sP, row);
        }
      }
      WeightSrc sE;
      sE.p[0] = Wl + pk::W2_b; sE.p[1] = Wl + pk::Wc1_b; sE.p[2] = Wl + pk::Wc1_f; sE.p[3] = Wl + pk::W2_f; sE.p[4] = Wl + pk::A_f;
      sE.count = 5;
      load_weights(c, sE);  // (its barriers also publish the Q rows / zeroed scatter rows to the team)
      if (team_active) {
        const float4 xi = c.sX[l * kRows + tt];
        float gxi[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 32; ++k) gp[k] = 0.f;
#pragma unroll 1
        for (int u = 0; u < NP - 1; ++u) {
          const int rj = pp * NP + c.sender(u);
          const float4 yj = c.sX[rj];
          const Geo g = edge_geo4(xi, c.sX[l * kRows + rj], yi, yj);
          T.ld(sP, row);
          stage1<true>(row, f1, c.sQb, rj, vec, g.r2, g.ea);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 0, false); });
          T.ld(sAcc0, row);
          const float att = stage2<true>(row, m, f2, vec);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAccC, 1, false); });
          T.ld(sAccC, row);
          const float th = stage3<true>(row, fc, vec);
          const float phi = rng * th;
          // coordinate branch cotangent:  gu = <gx, dhat> * range * (1 - th^2);  gzc = gu * wc2 * fc
          const float gdot = gx[0] * g.d0 + gx[1] * g.d1 + gx[2] * g.d2;
          const float gu = gdot * g.inv * rng * (1.0f - th * th);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 w = lds4(vec + vWC2 * 32 + 4 * k4);
            row[4 * k4] = gu * w.x * fc[4 * k4]; row[4 * k4 + 1] = gu * w.y * fc[4 * k4 + 1];
            row[4 * k4 + 2] = gu * w.z * fc[4 * k4 + 2]; row[4 * k4 + 3] = gu * w.w * fc[4 * k4 + 3];
          }
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 2, false); });
          T.ld(sAcc0, row);
          T.ld(sGAgg, fc);  // (fc is dead: reuse its registers for gagg)
          float gs = 0.f;
#pragma unroll
          for (int k = 0; k < 32; ++k) { row[k] += fc[k]; gs = fmaf(row[k], m[k], gs); }  // gms, gs = <gms, m>
          const float ga = gs * att * (1.0f - att);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 w = lds4(vec + vWA * 32 + 4 * k4);
            row[4 * k4] = fmaf(row[4 * k4], att, w.x * ga) * f2[4 * k4];
            row[4 * k4 + 1] = fmaf(row[4 * k4 + 1], att, w.y * ga) * f2[4 * k4 + 1];
            row[4 * k4 + 2] = fmaf(row[4 * k4 + 2], att, w.z * ga) * f2[4 * k4 + 2];
            row[4 * k4 + 3] = fmaf(row[4 * k4 + 3], att, w.w * ga) * f2[4 * k4 + 3];
          }
          T.store_row(row);  // gz2
          T.round_trip([&] { T.mma(sAcc0, 3, false); });
          T.ld(sAcc0, row);
          float gr2 = 0.f, gea = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 cc = lds4(vec + vC1 * 32 + 4 * k4), dd = lds4(vec + vD1 * 32 + 4 * k4);
            const float c4[4] = {cc.x, cc.y, cc.z, cc.w}, d4[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = 4 * k4 + e;
              row[k] *= f1[k];  // gz1
              gp[k] += row[k];
              gr2 = fmaf(c4[e], row[k], gr2);
              gea = fmaf(d4[e], row[k], gea);
            }
          }
          // sender-side sums: gq_j += gz1   (rows beyond the last whole particle of the tile must not scatter)
#pragma unroll
          for (int k4 = 0; k4 < 8 && c.row_ok; ++k4) {
            float *q = sGQ + rj * 32 + ((k4 ^ (rj & 7)) << 2);
            const float4 o = lds4(q);
            sts4(q, make_float4(o.x + row[4 * k4], o.y + row[4 * k4 + 1], o.z + row[4 * k4 + 2], o.w + row[4 * k4 + 3]));
          }
          // geometry: d/d(delta) through dhat = delta / (nrm + 1) (times phi) and through r2; edge_attr path on y
          const float k2 = gdot * phi * g.inv * g.inv / g.nrm;
          const float a0 = gx[0] * phi * g.inv - g.d0 * k2 + 2.0f * g.d0 * gr2;
          const float a1 = gx[1] * phi * g.inv - g.d1 * k2 + 2.0f * g.d1 * gr2;
          const float a2 = gx[2] * phi * g.inv - g.d2 * k2 + 2.0f * g.d2 * gr2;
          gxi[0] += a0; gxi[1] += a1; gxi[2] += a2;
          const float e0 = 2.0f * (yi.x - yj.x) * gea, e1 = 2.0f * (yi.y - yj.y) * gea, e2 = 2.0f * (yi.z - yj.z) * gea;
          g0[0] += e0; g0[1] += e1; g0[2] += e2;
          if (c.row_ok) {
            const float4 oa = sGX[rj], ob = sGX0[rj];
            sGX[rj] = make_float4(oa.x - a0, oa.y - a1, oa.z - a2, 0.f);
            sGX0[rj] = make_float4(ob.x - e0, ob.y - e1, ob.z - e2, 0.f);
          }
        }
        // ---- combine: gh += A^T gp (+ B^T gq below);  gx += own + scattered parts
        T.store_row(gp);
        T.round_trip([&] { T.mma(sGH, 4, true); });  // (its team barrier orders the last scatter before the reads below)
        qrow_load(sGQ, tt, row);
        const float4 sc = sGX[tt];
        gx[0] += gxi[0] + sc.x; gx[1] += gxi[1] + sc.y; gx[2] += gxi[2] + sc.z;
#pragma unroll
        for (int k = 0; k < 32; ++k) f1[k] = 0.f;
        qrow_store(sGQ, tt, f1);
        sGX[tt] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      WeightSrc sB;
      sB.p[0] = Wl + pk::B_f;
      sB.count = 1;
      load_weights(c, sB);
      if (team_active) {
        T.store_row(row);  // gq
        T.round_trip([&] { T.mma(sGH, 0, true); });
      }
    }

    // ---- outputs: dU/dy_i = vel_i - w_i + gx_i + g0_i;  dU/dtcond = sum_i <gh^0_i, d h^0_i / d tcond>
    if (team_active) {
      const float4 s0 = sGX0[tt];
      const float dUdy[3] = {(vi.x - vmean.x) - (yi.x - ymean.x) + gx[0] + g0[0] + s0.x,
                             (vi.y - vmean.y) - (yi.y - ymean.y) + gx[1] + g0[1] + s0.y,
                             (vi.z - vmean.z) - (yi.z - ymean.z) + gx[2] + g0[2] + s0.z};
      float dot_i = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dot_i = fmaf(dUdy[k], xr[k], dot_i);
        if (ok && grad_x) grad_x[part * 3 * NP + 3 * c.i + k] = xr[k] / (1.0f + h) - rs_h * c_in * dUdy[k];
      }
      float dt_i = 0.f;
      if (dE_dh != nullptr) {
        T.ld(sGH, row);
        const float t0 = (2 * ic < NP) ? 1.0f : 0.0f, t1 = (2 * ic + 1 < NP) ? 1.0f : 0.0f;
#pragma unroll
        for (int k = 0; k < 32; ++k) dt_i = fmaf(row[k], t0 * c.sEmb[k] + t1 * c.sEmb[32 + k], dt_i);
        const float dUdc = particle_sum<NP, SPLIT>(T, sRed, c.row_ok ? dt_i : 0.f, pp);
        const float dot = particle_sum<NP, SPLIT>(T, sRed, c.row_ok ? dot_i : 0.f, pp);
        if (ok && c.i == 0) {
          const float op = 1.0f + h;
          const float dU_dh = dUdc / (8.0f * h) + dot * (-0.5f) * rsqrtf(op) / op;
          dE_dh[part] = -x2 / (2.0f * op * op) + 0.5f * rs_h / h * U - rs_h * dU_dh;
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if ((c.tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

template <int NP, int NTEAM, bool SPLIT>
static int launch_energy_rows(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *e, float *g,
                              float *dh, cudaStream_t s) {
  using S = Smem<NP, NTEAM, SPLIT, kModeRev>;
  auto k = egnn_energy_rows_kernel<NP, NTEAM, SPLIT>;
  static_assert(S::kBytes <= 227 * 1024, "shared memory plan exceeds 227 KB");
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
  if (err != cudaSuccess) { set_error("egnn_energy_rows_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return PITA_ECUDA; }
  const int64_t nbatch = (B + (int64_t)NTEAM * Shape<NP>::PB - 1) / ((int64_t)NTEAM * Shape<NP>::PB);
  const unsigned grid = (unsigned)(nbatch < kNumSMs ? nbatch : kNumSMs);
  k<<<grid, NTEAM * 128, S::kBytes, s>>>(w, ht, x, beta, B, e, g, dh);
  PITA_CHECK_LAUNCH("egnn_energy_rows_kernel");
  return PITA_OK;
}


// ================================================================================================
// Kernel: score + exact divergence (forward-mode tangents on the row engine)
//
// trace = sum_k sum_a d x^3[k,a] / d y[k,a] is accumulated in NP passes, pass k carrying the three directions
// (k, a) of every particle of the tile at once.  Structure used (oracle/egnn_analytic.py):
//   layer 0: h^0 does not depend on y, so for i != k the tangent of node i is rank one,
//            dh^1_i[(k,a)] = cf_a * omega_ik,  cf_a = -2 (y_i - y_k)_a,  omega_ik = W4 (f3_i * W3a wvec_ik),
//            wvec_ik = d ms_ik / d r2 (one extra row through W2); for i == k it is the sum over i's 12 edges,
//            pre-computed for every node in a pre-phase (own-direction vectors, kept in the scratch buffer);
//   layer 1: dense — every receiver row walks its n-1 sender slots, 3 tangent rows per edge through W2 and
//            Wc1, the aggregate through W3a accumulated over slots inside TMEM;
//   layer 2: only d x^3_k / d y_k is needed: edge (k, j) is evaluated by the SENDER's thread j.
// ================================================================================================
enum Scr { qF30 = 0, qF31 = 1, qP2 = 2, qQ2 = 3, qP1 = 4, qOmega = 5, qDh1o = 6, qPiAo = 9, qPiBo = 12, qDxo = 15, kScrVecs = 16 };
// Layer-1 edge cache: the primal quantities of edge (i, sender slot u) -- silu'(z1), m, silu'(z2), silu'(zc), the attention
// gate and tanh(u) -- do not depend on the tangent node k, so pass k = 0 writes them to the team's scratch (own-row
// layout: every thread only ever reads back what it wrote itself) and passes k >= 1 load them instead of redoing two
// MMA round trips and 96 SiLU evaluations per edge.  Floats per slot: 4 vectors x [128][32] + one float4 per row.
constexpr int kEdgeVecs = 4;
constexpr int kEdgeFloats = kEdgeVecs * kRows * 32 + kRows * 4;
template <int NP>
__host__ __device__ constexpr int64_t team_scratch_floats() { return (int64_t)kScrVecs * kRows * 32 + (int64_t)(NP - 1) * kEdgeFloats; }



template <int NP, int NTEAM, bool SPLIT>
__global__ void __launch_bounds__(NTEAM * 128, 1)
egnn_score_div_rows_kernel(const float *__restrict__ wpack, const float *__restrict__ ht, const float *__restrict__ x,
                           const float *__restrict__ beta_in, int64_t B, float *__restrict__ score,
                           float *__restrict__ divergence, float *__restrict__ scratch) {
  extern __shared__ __align__(16) float sm_raw[];
  using C = Ctx<NP, NTEAM, SPLIT, kModeTan>;
  using S = Smem<NP, NTEAM, SPLIT, kModeTan>;
  constexpr int PB = C::PB;
  constexpr int L = 3;
  const float rng = kCoordsRange / (float)L;
  C c;
  uint32_t tmem_base;
  setup(c, sm_raw, wpack, tmem_base);
  Team<SPLIT> &T = c.T;
  const int tt = T.tt;
  const float *W0 = wpack + pk::kHeader, *W1 = W0 + pk::kLayer, *W2l = W1 + pk::kLayer;
  float4 *sDX = reinterpret_cast<float4 *>(c.tm + S::tDX);
  float4 *sCoef = reinterpret_cast<float4 *>(c.tm + S::tCoef);
  float *sOwnA = c.tm + S::tOwnA, *sOwnB = c.tm + S::tOwnB, *sKP2 = c.tm + S::tKP2, *sKdP2 = c.tm + S::tKdP2;
  float *team_scr = scratch ? scratch + ((size_t)blockIdx.x * NTEAM + c.team) * (size_t)team_scratch_floats<NP>() : nullptr;
  float *scr = team_scr ? team_scr + (size_t)tt * 32 : nullptr;
  auto SCR = [&](int v) { return scr + (size_t)v * kRows * 32; };
  float *escr = team_scr ? team_scr + (size_t)kScrVecs * kRows * 32 : nullptr;
  asm volatile("" : "+l"(escr));  // opaque: keep the pointer in a register / stack slot instead of re-deriving it from blockIdx
  auto ESCR = [&](int u, int v) { return escr + (size_t)u * kEdgeFloats + (size_t)v * kRows * 32; };  // coalesced block
  auto ESCAL = [&](int u) { return reinterpret_cast<float4 *>(escr + (size_t)u * kEdgeFloats + kEdgeVecs * kRows * 32) + tt; };

  // layer-0 class tables:  A0 e0, A0 e1, A0 eb + b1, B0 e0, B0 e1, B0 eb
  if (c.tid < 192) {
    const int v = c.tid >> 5, ch = c.tid & 31;
    const float *M = W0 + (v < 3 ? pk::A_b : pk::B_b) + ch * 32;
    const float *e = c.sEmb + (v % 3) * 32;
    float acc = (v == 2) ? __ldg(W0 + pk::b1 + ch) : 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(__ldg(M + k), e[k], acc);
    c.sCls[v * 32 + ch] = acc;
  }
  __syncthreads();

  const int64_t nbatch = (B + (int64_t)NTEAM * PB - 1) / ((int64_t)NTEAM * PB);
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int64_t p0 = (batch * NTEAM + c.team) * PB;
    const bool team_active = p0 < B;
    const int64_t part = p0 + c.p;
    const bool ok = c.row_ok && part < B;
    const int64_t pc = ok ? part : (B - 1);
    const int ic = c.row_ok ? c.i : 0;
    const int pp = c.row_ok ? c.p : 0;
    const float h = __ldg(ht + pc), beta = __ldg(beta_in + pc);
    const float c_in = rsqrtf(1.0f + h), c_s = 1.0f / (1.0f + h), c_out = sqrtf(h) * c_in, c_noise = 0.125f * logf(h);
    const float *src = x + pc * 3 * NP + 3 * ic;
    const float xr0 = __ldg(src), xr1 = __ldg(src + 1), xr2 = __ldg(src + 2);
    const float4 yi = make_float4(c_in * xr0, c_in * xr1, c_in * xr2, 0.f);
    c.sX[tt] = yi;
    primal_forward_rows(c, wpack, c_noise, beta, team_active, divergence ? scr : nullptr);

    // ---- score = ((c_s - 1) x + c_out * remove_mean(x^3 - y)) / h        (score_net.py:21-43)
    if (team_active) {
      const float4 xl = c.sX3[tt];
      const float4 v = make_float4(xl.x - yi.x, xl.y - yi.y, xl.z - yi.z, 0.f);
      const float4 mean = particle_mean<NP, SPLIT>(T, reinterpret_cast<float4 *>(c.sQb), v, pp);
      if (ok) {
        float *dst = score + part * 3 * NP + 3 * c.i;
        dst[0] = ((c_s * xr0 + c_out * (v.x - mean.x)) - xr0) / h;
        dst[1] = ((c_s * xr1 + c_out * (v.y - mean.y)) - xr1) / h;
        dst[2] = ((c_s * xr2 + c_out * (v.z - mean.z)) - xr2) / h;
      }
    }
    if (divergence == nullptr) continue;

    float f0i, f1i;
    node_feats<NP>(ic, c_noise, beta, f0i, f1i);
    const float *vec0 = c.vec(0), *vec1 = c.vec(1), *vec2 = c.vec(2);
    float row[32], f1[32], m[32], f2[32], fc[32];

    // =========================== pre-phase: own-direction layer-0 tangents of every node
    WeightSrc setA;
    setA.p[0] = W0 + pk::W2_b; setA.p[1] = W0 + pk::Wc1_b; setA.p[2] = W0 + pk::W3a_b; setA.p[3] = W0 + pk::W4_b;
    setA.p[4] = W1 + pk::A_b;
    setA.count = 5;
    WeightSrc setB;  // B1, then the layer-1 edge set and W3h1
    setB.p[0] = W1 + pk::B_b; setB.p[1] = W1 + pk::W2_b; setB.p[2] = W1 + pk::Wc1_b; setB.p[3] = W1 + pk::W3a_b;
    setB.p[4] = W1 + pk::W3h_b;
    setB.count = 5;
    WeightSrc setC;  // W4_1, layer-2 node products and edge set
    setC.p[0] = W1 + pk::W4_b; setC.p[1] = W2l + pk::A_b; setC.p[2] = W2l + pk::B_b; setC.p[3] = W2l + pk::W2_b;
    setC.p[4] = W2l + pk::Wc1_b;
    setC.count = 5;
    load_weights(c, setA);
    if (team_active) {
      float dxo[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) dxo[a][b] = (a == b) ? 1.0f : 0.0f;
#pragma unroll 1
      for (int u = 0; u < NP - 1; ++u) {
        const int j = c.sender(u), rj = pp * NP + j;
        const float4 yj = c.sX[rj];
        const Geo g = edge_geo4(yi, yj, yi, yj);
        float f0j, f1j;
        node_feats<NP>(j, c_noise, beta, f0j, f1j);
        z1_layer0(row, c.sCls, f0i, f1i, f0j, f1j);
        stage1<true, false>(row, f1, nullptr, 0, vec0, g.r2, g.ea);
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, 0, false); });
        T.ld(sAcc0, row);
        const float att = stage2<true>(row, m, f2, vec0);
        T.store_row(row);
        T.round_trip([&] { T.mma(sAccC, 1, false); });
        T.ld(sAccC, row);
        const float th = stage3<true>(row, fc, vec0);
        const float phi = rng * th, dphi_du = rng * (1.0f - th * th);
        // base row: d z1 / d r2 (edge_attr == radial in layer 0)
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 cc = lds4(vec0 + vC1 * 32 + 4 * k4), dd = lds4(vec0 + vD1 * 32 + 4 * k4);
          row[4 * k4] = f1[4 * k4] * (cc.x + dd.x); row[4 * k4 + 1] = f1[4 * k4 + 1] * (cc.y + dd.y);
          row[4 * k4 + 2] = f1[4 * k4 + 2] * (cc.z + dd.z); row[4 * k4 + 3] = f1[4 * k4 + 3] * (cc.w + dd.w);
        }
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, 0, false); });
        T.ld(sAcc0, row);
        tangent_mid(row, m, f2, att, vec0);  // wvec
        T.store_row(row);
        const float dd3[3] = {g.d0, g.d1, g.d2};
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // S_a += 2 d_a wvec   (TMEM-resident accumulators)
          float tmp[32];
          if (u == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) tmp[k] = 2.0f * dd3[a] * row[k];
          } else {
            T.ld(sT0 + a, tmp);
#pragma unroll
            for (int k = 0; k < 32; ++k) tmp[k] = fmaf(2.0f * dd3[a], row[k], tmp[k]);
          }
          T.st(sT0 + a, tmp);
        }
        T.round_trip([&] { T.mma(sAccC, 1, false); });
        T.ld(sAccC, row);
        const float du_base = tangent_du(row, fc, vec0);
        const float k2 = g.inv * g.inv / g.nrm;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float dphi = dphi_du * du_base * 2.0f * dd3[a];
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const float ddh = ((a == b) ? g.inv : 0.0f) - dd3[b] * dd3[a] * k2;
            dxo[a][b] += ddh * phi + dd3[b] * g.inv * dphi;
          }
        }
      }
      // finish: dh1o[a] = W4 (f3^0 * W3a S_a);  piAo = A1 dh1o, piBo = B1 dh1o
      load_vec_global(SCR(qF30), f1);  // f1 <- f3^0
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        T.ld(sT0 + a, row);
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, 2, false); });
        T.ld(sAcc0, row);
#pragma unroll
        for (int k = 0; k < 32; ++k) row[k] *= f1[k];
        T.store_row(row);
        T.round_trip([&] { T.mma(sAcc0, 3, false); });
        T.ld(sAcc0, row);
        store_vec_global(SCR(qDh1o + a), row);
        T.store_row(row);
        T.round_trip([&] { T.mma(sAccC, 4, false); });
        T.ld(sAccC, row);
        store_vec_global(SCR(qPiAo + a), row);
      }
      {
        float *d = SCR(qDxo);
        *reinterpret_cast<float4 *>(d) = make_float4(dxo[0][0], dxo[0][1], dxo[0][2], 0.f);
        *reinterpret_cast<float4 *>(d + 4) = make_float4(dxo[1][0], dxo[1][1], dxo[1][2], 0.f);
        *reinterpret_cast<float4 *>(d + 8) = make_float4(dxo[2][0], dxo[2][1], dxo[2][2], 0.f);
      }
      // P^1 back into its TMEM slot for the passes
      load_vec_global(SCR(qP1), row);
      T.st(sP, row);
    }
    {  // piBo[a] = B1 dh1o[a]
      WeightSrc sB1;
      sB1.p[0] = W1 + pk::B_b;
      sB1.count = 1;
      load_weights(c, sB1);
      if (team_active) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          load_vec_global(SCR(qDh1o + a), row);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 0, false); });
          T.ld(sAcc0, row);
          store_vec_global(SCR(qPiBo + a), row);
        }
      }
    }

    // =========================== layer-1 edge cache (see kEdgeFloats): f1, m, f2, v = Wc1^T (wc2 * fc), att, tanh(u)
    {
      WeightSrc sL1;
      sL1.p[0] = W1 + pk::W2_b; sL1.p[1] = W1 + pk::Wc1_b; sL1.p[2] = W1 + pk::Wc1_f;
      sL1.count = 3;
      load_weights(c, sL1);
      if (team_active) {
        const float4 xi = c.sX[kRows + tt];
#pragma unroll 1
        for (int u = 0; u < NP - 1; ++u) {
          const int rj = pp * NP + c.sender(u);
          const Geo g = edge_geo4(xi, c.sX[kRows + rj], yi, c.sX[rj]);
          T.ld(sP, row);
          stage1<true>(row, f1, c.sQa, rj, vec1, g.r2, g.ea);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 0, false); });
          T.ld(sAcc0, row);
          const float att = stage2<true>(row, m, f2, vec1);
          T.store_row(row);
          T.round_trip([&] { T.mma(sAccC, 1, false); });
          T.ld(sAccC, row);
          const float th = stage3<true>(row, fc, vec1);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 w = lds4(vec1 + vWC2 * 32 + 4 * k4);
            row[4 * k4] = w.x * fc[4 * k4]; row[4 * k4 + 1] = w.y * fc[4 * k4 + 1];
            row[4 * k4 + 2] = w.z * fc[4 * k4 + 2]; row[4 * k4 + 3] = w.w * fc[4 * k4 + 3];
          }
          T.store_row(row);
          T.round_trip([&] { T.mma(sAcc0, 2, false); });
          T.ld(sAcc0, row);  // v_ij
          store_vec_global_co(ESCR(u, 0), tt, f1);
          store_vec_global_co(ESCR(u, 1), tt, m);
          store_vec_global_co(ESCR(u, 2), tt, f2);
          store_vec_global_co(ESCR(u, 3), tt, row);
          __stcg(ESCAL(u), make_float4(att, th, 0.f, 0.f));
        }
      }
    }

    // =========================== passes over the tangent node k
    float trace = 0.f;
#pragma unroll 1
    for (int k = 0; k < NP; ++k) {
      const bool is_k = c.row_ok && (c.i == k);
      const int rk = pp * NP + k;
      float cf[3] = {0.f, 0.f, 0.f};
      float dxi[3][3];
      load_weights(c, setA);
      // ---------------- layer 0: edge (i, k)
      if (team_active) {
        const float4 yk = c.sX[rk];
        const Geo g = edge_geo4(yi, yk, yi, yk);
        float f0k, f1k;
        node_feats<NP>(k, c_noise, beta, f0k, f1k);
        z1_layer0(row, c.sCls, f0i, f1i, f0k, f1k);
        stage1<true, false>(row, f1, nullptr, 0, vec0, g.r2, g.ea);
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 0, false); });
        T.ld(sAcc0, row);
        const float att = stage2<true>(row, m, f2, vec0);
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAccC, sH, sP, 1, false); });
        T.ld(sAccC, row);
        const float th = stage3<true>(row, fc, vec0);
        const float phi = rng * th, dphi_du = rng * (1.0f - th * th);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 cc = lds4(vec0 + vC1 * 32 + 4 * k4), dd = lds4(vec0 + vD1 * 32 + 4 * k4);
          row[4 * k4] = f1[4 * k4] * (cc.x + dd.x); row[4 * k4 + 1] = f1[4 * k4 + 1] * (cc.y + dd.y);
          row[4 * k4 + 2] = f1[4 * k4 + 2] * (cc.z + dd.z); row[4 * k4 + 3] = f1[4 * k4 + 3] * (cc.w + dd.w);
        }
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 0, false); });
        T.ld(sAcc0, row);
        tangent_mid(row, m, f2, att, vec0);  // wvec_ik
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAccC, sH, sP, 1, false); T.mma_ts(sAcc0, sH, sP, 2, false); });
        T.ld(sAccC, row);
        const float du_base = tangent_du(row, fc, vec0);
        T.ld(sAcc0, row);
        load_vec_global(SCR(qF30), f1);
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) row[kk] *= f1[kk];
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 3, false); });
        T.ld(sAcc0, row);  // omega_ik
        store_vec_global(SCR(qOmega), row);
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sPA, sH, sP, 4, false); });  // piA_ik = A1 omega -> its TMEM slot (piB follows once B1 is loaded)
        const float dd3[3] = {g.d0, g.d1, g.d2};
        if (!is_k) {
          const float k2 = g.inv * g.inv / g.nrm;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            cf[a] = -2.0f * dd3[a];
            const float dphi = dphi_du * du_base * cf[a];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const float ddh = ((a == b) ? -g.inv : 0.0f) + dd3[b] * dd3[a] * k2;
              dxi[a][b] = ddh * phi + dd3[b] * g.inv * dphi;
            }
          }
        } else {
          const float *d = SCR(qDxo);
          const float4 q0 = *reinterpret_cast<const float4 *>(d), q1 = *reinterpret_cast<const float4 *>(d + 4),
                       q2 = *reinterpret_cast<const float4 *>(d + 8);
          dxi[0][0] = q0.x; dxi[0][1] = q0.y; dxi[0][2] = q0.z;
          dxi[1][0] = q1.x; dxi[1][1] = q1.y; dxi[1][2] = q1.z;
          dxi[2][0] = q2.x; dxi[2][1] = q2.y; dxi[2][2] = q2.z;
          // publish the tangent node's own-direction vectors and its P^2
#pragma unroll 1
          for (int a = 0; a < 3; ++a) {
            load_vec_global(SCR(qPiAo + a), row);
            store_vec_smem(sOwnA + (pp * 3 + a) * 32, row);
            load_vec_global(SCR(qPiBo + a), row);
            store_vec_smem(sOwnB + (pp * 3 + a) * 32, row);
          }
          load_vec_global(SCR(qP2), row);
          store_vec_smem(sKP2 + pp * 32, row);
        }
        sCoef[tt] = make_float4(cf[0], cf[1], cf[2], 0.f);
#pragma unroll
        for (int a = 0; a < 3; ++a) sDX[tt * 3 + a] = make_float4(dxi[a][0], dxi[a][1], dxi[a][2], 0.f);
      }
      // ---------------- layer 1 (dense)
      load_weights(c, setB);
      if (team_active) {  // piB_ik = B1 omega (the A tile still holds the omega rows)
        T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 0, false); });
        T.ld(sAcc0, row);
        if (is_k) {
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) row[kk] = 0.f;
        }
        qrow_store(c.sQb, tt, row);
        umma::fence_before_thread_sync();
        T.sync();
        umma::fence_after_thread_sync();
      }
      float dx2[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) dx2[a][b] = dxi[a][b];
      if (team_active) {
        const float4 xi = c.sX[kRows + tt];
#pragma unroll 1
        for (int u = 0; u < NP - 1; ++u) {
          const int j = c.sender(u), rj = pp * NP + j;
          const bool jk = (j == k);
          const float4 yj = c.sX[rj];
          const Geo g = edge_geo4(xi, c.sX[kRows + rj], yi, yj);
          // primal quantities of this edge from the layer-1 edge cache (filled once per tile, before the passes)
          {  // the cache of all resident teams (237 MB at n = 13) does not stay in L2: pull the next slot in while this one
             // is being processed (528 lines of 128 B per slot, 4-5 per thread; wraps to slot 0 for the next pass)
            const int un = (u + 1 < NP - 1) ? u + 1 : 0;
            const char *nb = reinterpret_cast<const char *>(escr + (size_t)un * kEdgeFloats);
#pragma unroll
            for (int i = 0; i < (kEdgeFloats * 4 / 128 + kRows - 1) / kRows; ++i) {
              const int line = tt + i * kRows;
              if (line < kEdgeFloats * 4 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (size_t)line * 128));
            }
          }
          load_vec_global_co(ESCR(u, 0), tt, f1);
          load_vec_global_co(ESCR(u, 1), tt, m);
          load_vec_global_co(ESCR(u, 2), tt, f2);
          load_vec_global_co(ESCR(u, 3), tt, fc);  // fc <- v_ij = Wc1^T (wc2 * silu'(zc))
          const float4 sc4 = __ldcg(ESCAL(u));
          const float att = sc4.x, th = sc4.y;
          const float phi = rng * th, dphi_du = rng * (1.0f - th * th);
          const float k2 = g.inv * g.inv / g.nrm;
          const float4 cfj = sCoef[rj];
          const float e03[3] = {yi.x - yj.x, yi.y - yj.y, yi.z - yj.z};
          const float sgn = (is_k ? 1.0f : 0.0f) - (jk ? 1.0f : 0.0f);
          const bool accz = u > 0;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float4 dxj = sDX[rj * 3 + a];
            const float D0 = dxi[a][0] - dxj.x, D1 = dxi[a][1] - dxj.y, D2 = dxi[a][2] - dxj.z;
            const float dotD = g.d0 * D0 + g.d1 * D1 + g.d2 * D2;
            // tangent input  dP_i + dQ_j
            T.ld(sPA, row);
            const float cA = cf[a], cB = comp(cfj, a);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) row[kk] = is_k ? 0.f : cA * row[kk];
            if (is_k) add_vec(row, sOwnA + (pp * 3 + a) * 32);
            {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 q = qrow_ld4(c.sQb, rj, k4);
                row[4 * k4] = fmaf(cB, q.x, row[4 * k4]); row[4 * k4 + 1] = fmaf(cB, q.y, row[4 * k4 + 1]);
                row[4 * k4 + 2] = fmaf(cB, q.z, row[4 * k4 + 2]); row[4 * k4 + 3] = fmaf(cB, q.w, row[4 * k4 + 3]);
              }
            }
            if (jk) add_vec(row, sOwnB + (pp * 3 + a) * 32);
            tangent_in(row, f1, vec1, 2.0f * dotD, 2.0f * sgn * e03[a]);
            // TS form: operand row handed over through the row's own TMEM lane (slots sH / sP are free during the passes)
            T.store_row_tmem(sH, sP, row);
            T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 1, false); });
            T.ld(sAcc0, row);
            tangent_mid(row, m, f2, att, vec1);  // row = d(m*_ij)
            // du = <wc2 * silu'(zc), Wc1 dms> = <v_ij, dms>: a dot product with the cached vector, no MMA
            float du = 0.f;
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) du = fmaf(fc[kk], row[kk], du);
            const float dphi = dphi_du * du;
            // d agg_i[a] += dms: summed in fp32 (round-to-nearest) in the row's own TMEM lane; W3a is applied once per
            // direction after the slots (linearity) instead of one accumulating MMA per edge
            {
              float tmp[32];
              if (accz) {
                T.ld(sT0 + a, tmp);
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) tmp[kk] += row[kk];
                T.st(sT0 + a, tmp);
              } else {
                T.st(sT0 + a, row);
              }
            }
            const float cg = dotD * k2;
            dx2[a][0] += (D0 * g.inv - g.d0 * cg) * phi + g.d0 * g.inv * dphi;
            dx2[a][1] += (D1 * g.inv - g.d1 * cg) * phi + g.d1 * g.inv * dphi;
            dx2[a][2] += (D2 * g.inv - g.d2 * cg) * phi + g.d2 * g.inv * dphi;
          }
        }
        // dz3[a] = W3a1 dagg[a] + W3h1 dh^1[a]   (W3a1 / W3h1 sit in slots 3 / 4 of this set): one round trip per direction,
        // the aggregate handed over through TMEM (TS form), dh^1 through the shared-memory tile.  The direction loops of
        // the per-k sections are NOT unrolled (direction-indexed values come from shared memory): this code runs once per
        // pass and its unrolled size (166 KB) made instruction fetch 37 % of its stalls (profiles/README.md).
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
          const float cfa = comp(sCoef[tt], a);
          T.ld(sT0 + a, row);
          T.store_row_tmem(sH, sP, row);
          if (is_k) {
            load_vec_global(SCR(qDh1o + a), row);
          } else {
            load_vec_global(SCR(qOmega), row);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) row[kk] *= cfa;
          }
          T.store_row(row);
          T.round_trip([&] { T.mma_ts(sT0 + a, sH, sP, 3, false); T.mma(sT0 + a, 4, true); });
        }
      }
      // ---------------- node update of layer 1 on the tangents, then layer 2 on edge (k, i) by the sender's thread
      load_weights(c, setC);
      if (team_active) {
#pragma unroll
        for (int a = 0; a < 3; ++a) sDX[tt * 3 + a] = make_float4(dx2[a][0], dx2[a][1], dx2[a][2], 0.f);
        umma::fence_before_thread_sync();
        T.sync();
        umma::fence_after_thread_sync();
        const float4 yk = c.sX[rk];
        const Geo g = edge_geo4(c.sX[2 * kRows + rk], c.sX[2 * kRows + tt], yk, yi);  // d = x_k - x_i (receiver k)
        load_vec_smem(sKP2 + pp * 32, row);
        load_vec_global(SCR(qQ2), f1);
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) row[kk] += f1[kk];
        stage1<true, false>(row, f1, nullptr, 0, vec2, g.r2, g.ea);
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 3, false); });
        T.ld(sAcc0, row);
        const float att = stage2<true>(row, m, f2, vec2);
        T.store_row_tmem(sH, sP, row);
        T.round_trip_ts([&] { T.mma_ts(sAccC, sH, sP, 4, false); });
        T.ld(sAccC, row);
        const float th = stage3<true>(row, fc, vec2);
        const float phi = rng * th, dphi_du = rng * (1.0f - th * th);
        const float k2 = g.inv * g.inv / g.nrm;
        const float4 dd4 = make_float4(g.d0, g.d1, g.d2, 0.f);
        const float4 e04 = make_float4(yk.x - yi.x, yk.y - yi.y, yk.z - yi.z, 0.f);
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
          float tmp[32];
          const float cfa = comp(sCoef[tt], a), dda = comp(dd4, a), e0a = comp(e04, a);
          const float4 dxo = sDX[tt * 3 + a];  // this row's own d x^2 [a] (published above)
          T.ld(sT0 + a, row);  // dz3 = W3a dagg (summed over the slots) + W3h dh1
          load_vec_global(SCR(qF31), tmp);
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) row[kk] *= tmp[kk];
          T.store_row_tmem(sH, sP, row);
          T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 0, false); });
          T.ld(sAcc0, row);
          if (is_k) {
            load_vec_global(SCR(qDh1o + a), tmp);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) row[kk] += tmp[kk];
          } else {
            load_vec_global(SCR(qOmega), tmp);
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) row[kk] = fmaf(cfa, tmp[kk], row[kk]);
          }
          T.store_row_tmem(sH, sP, row);  // dh^2[a]
          T.round_trip_ts([&] { T.mma_ts(sAccC, sH, sP, 1, false); T.mma_ts(sAcc0, sH, sP, 2, false); });
          T.ld(sAccC, tmp);  // (tcgen05.ld is warp-collective: never inside a divergent branch)
          if (is_k) store_vec_smem(sKdP2 + (pp * 3 + a) * 32, tmp);
          umma::fence_before_thread_sync();
          T.sync();
          umma::fence_after_thread_sync();
          T.ld(sAcc0, row);  // dQ2_i[a]
          add_vec(row, sKdP2 + (pp * 3 + a) * 32);
          const float4 dxk = sDX[rk * 3 + a];
          const float D0 = dxk.x - dxo.x, D1 = dxk.y - dxo.y, D2 = dxk.z - dxo.z;
          const float dotD = g.d0 * D0 + g.d1 * D1 + g.d2 * D2;
          tangent_in(row, f1, vec2, 2.0f * dotD, 2.0f * e0a);
          T.store_row_tmem(sH, sP, row);
          T.round_trip_ts([&] { T.mma_ts(sAcc0, sH, sP, 3, false); });
          T.ld(sAcc0, row);
          tangent_mid(row, m, f2, att, vec2);
          T.store_row_tmem(sH, sP, row);
          T.round_trip_ts([&] { T.mma_ts(sAccC, sH, sP, 4, false); });
          T.ld(sAccC, row);
          const float dphi = dphi_du * tangent_du(row, fc, vec2);
          const float Da = (a == 0) ? D0 : ((a == 1) ? D1 : D2);
          if (is_k) trace += comp(dxo, a);
          else trace += (Da * g.inv - dda * dotD * k2) * phi + dda * g.inv * dphi;
        }
      }
    }
    // ---- div = ((c_s - 1) D + c_out c_in (tr - D)) / h
    if (team_active) {
      const float tr = particle_sum<NP, SPLIT>(T, c.tm + S::tRed, c.row_ok ? trace : 0.f, pp);
      if (ok && c.i == 0) {
        const float Dn = (float)(3 * NP);
        divergence[part] = ((c_s - 1.0f) * Dn + c_out * c_in * (tr - Dn)) / h;
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if ((c.tid >> 5) == 0) umma::tmem_dealloc<512>(tmem_base);
}

template <int NP, int NTEAM, bool SPLIT>
static int launch_score_div_rows(const float *w, const float *ht, const float *x, const float *beta, int64_t B, float *sc,
                                 float *dv, float *scratch, int64_t scratch_bytes, cudaStream_t s) {
  using S = Smem<NP, NTEAM, SPLIT, kModeTan>;
  auto k = egnn_score_div_rows_kernel<NP, NTEAM, SPLIT>;
  static_assert(S::kBytes <= 227 * 1024, "shared memory plan exceeds 227 KB");
  const int64_t need = (int64_t)kNumSMs * NTEAM * team_scratch_floats<NP>() * 4;
  PITA_REQUIRE(dv == nullptr || (scratch != nullptr && scratch_bytes >= need), PITA_EINVAL,
               "egnn_score_div: workspace too small (need %lld bytes)", (long long)need);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
  if (e != cudaSuccess) { set_error("egnn_score_div_rows_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PITA_ECUDA; }
  const int64_t nbatch = (B + (int64_t)NTEAM * Shape<NP>::PB - 1) / ((int64_t)NTEAM * Shape<NP>::PB);
  const unsigned grid = (unsigned)(nbatch < kNumSMs ? nbatch : kNumSMs);
  k<<<grid, NTEAM * 128, S::kBytes, s>>>(w, ht, x, beta, B, sc, dv, dv ? scratch : nullptr);
  PITA_CHECK_LAUNCH("egnn_score_div_rows_kernel");
  return PITA_OK;
}

}  // namespace rg

int launch_forward_rows(int n, bool split, const float *w, const float *tc, const float *y, const float *beta, int64_t B,
                        float *vel, cudaStream_t s) {
  if (n == 13)
    return split ? rg::launch_forward_rows<13, 2, true>(w, tc, y, beta, B, vel, s)
                 : rg::launch_forward_rows<13, 2, false>(w, tc, y, beta, B, vel, s);
  return split ? rg::launch_forward_rows<55, 2, true>(w, tc, y, beta, B, vel, s)
               : rg::launch_forward_rows<55, 2, false>(w, tc, y, beta, B, vel, s);
}

int launch_energy_rows(int n, bool split, const float *w, const float *ht, const float *x, const float *beta, int64_t B,
                       float *e, float *g, float *dh, cudaStream_t s) {
  if (n == 13)
    return split ? rg::launch_energy_rows<13, 2, true>(w, ht, x, beta, B, e, g, dh, s)
                 : rg::launch_energy_rows<13, 2, false>(w, ht, x, beta, B, e, g, dh, s);
  return split ? rg::launch_energy_rows<55, 2, true>(w, ht, x, beta, B, e, g, dh, s)
               : rg::launch_energy_rows<55, 2, false>(w, ht, x, beta, B, e, g, dh, s);
}

int64_t score_div_rows_workspace_bytes(int n) {
  return (int64_t)kNumSMs * 2 * (n == 13 ? rg::team_scratch_floats<13>() : rg::team_scratch_floats<55>()) * 4;
}

int launch_score_div_rows(int n, bool split, const float *w, const float *ht, const float *x, const float *beta, int64_t B,
                          float *sc, float *dv, float *scratch, int64_t scratch_bytes, cudaStream_t s) {
  if (n == 13)
    return split ? rg::launch_score_div_rows<13, 2, true>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s)
                 : rg::launch_score_div_rows<13, 2, false>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s);
  return split ? rg::launch_score_div_rows<55, 2, true>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s)
               : rg::launch_score_div_rows<55, 2, false>(w, ht, x, beta, B, sc, dv, scratch, scratch_bytes, s);
}

}  // namespace pita
