// Systematic resampling (replaces sample_cat_sys, models/components/utils.py:111-120, and the gather
// x_next[choice], sde_integration.py:293).  All HBM-bound, fp32 weights, fp64 prefix sums.
//
//   softmax_clip : two passes over the logits (online max/sum partials, then normalise + clip)
//   systematic   : fp64 tile sums -> single-CTA scan of tile sums -> bins = fp32(inclusive fp64 scan)
//                  -> one binary search per offspring slot (np.digitize(right=True) semantics)
//   gather       : one warp per destination row, source rows may live in NVLink peer memory
#include "common.cuh"

namespace pita {

constexpr int kSmThreads = 256;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 weights per CTA
constexpr int kMaxPartials = kNumSMs * 8;

struct ResampleWs {
  float gmax;
  float gsum;
  unsigned long long changes;
  unsigned long long pad;
};
// workspace layout: [ResampleWs (32 B)] [float2 partials[kMaxPartials]] [double tile_sums[ntiles]] [float bins[N]]
__host__ __device__ inline int64_t ws_tiles(int64_t N) { return (N + kScanTile - 1) / kScanTile; }
__host__ __device__ inline int64_t ws_off_partials() { return 32; }
__host__ __device__ inline int64_t ws_off_tiles() { return ws_off_partials() + (int64_t)kMaxPartials * 8; }
__host__ __device__ inline int64_t ws_off_bins(int64_t N) { return ws_off_tiles() + ((ws_tiles(N) * 8 + 15) / 16) * 16; }
__host__ __device__ inline int64_t ws_bytes(int64_t N) { return ws_off_bins(N) + ((N * 4 + 15) / 16) * 16; }

__device__ __forceinline__ void online_merge(float &m, float &s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = s * (m == -INFINITY ? 0.f : expf(m - mn)) + s2 * (m2 == -INFINITY ? 0.f : expf(m2 - mn));
  m = mn;
}

__global__ void __launch_bounds__(kSmThreads)
softmax_partials_kernel(const float *__restrict__ a, int64_t N, float2 *__restrict__ partials) {
  float m = -INFINITY, s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kSmThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kSmThreads) {
    const float v = __ldg(a + i);
    if (v > m) { s = s * expf(m - v) + 1.0f; m = v; }
    else s += expf(v - m);
  }
  __shared__ float sm[kSmThreads / 32], ss[kSmThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kSmThreads / 32; ++w) online_merge(m, s, sm[w], ss[w]);
    partials[blockIdx.x] = make_float2(m, s);
  }
}

__global__ void softmax_finalize_kernel(const float2 *__restrict__ partials, int nparts, ResampleWs *ws) {
  // single warp: deterministic merge of the per-CTA partials
  float m = -INFINITY, s = 0.f;
  for (int p = threadIdx.x; p < nparts; p += 32) online_merge(m, s, partials[p].x, partials[p].y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if (threadIdx.x == 0) { ws->gmax = m; ws->gsum = s; }
}

__global__ void __launch_bounds__(kSmThreads)
softmax_clip_kernel(const float *__restrict__ a, int64_t N, const ResampleWs *__restrict__ ws, float *__restrict__ w) {
  const float gmax = ws->gmax, gsum = ws->gsum;
  for (int64_t i = (int64_t)blockIdx.x * kSmThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kSmThreads) {
    const float p = expf(__ldg(a + i) - gmax) / gsum;
    w[i] = fminf(fmaxf(p, 1e-6f), 1.0f);  // torch.clip(softmax, 1e-6, 1.0), utils.py:114
  }
}

// ---- fp64 scan -------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double *sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < kScanThreads / 32 ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const float *__restrict__ w, int64_t N, double *__restrict__ tile_sums) {
  __shared__ double sh[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < N) acc += (double)__ldg(w + base + k);
  const double tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// exclusive scan of tile sums in place, by one CTA (sequential chunks per thread + CTA scan)
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(double *__restrict__ tile_sums, int64_t ntiles, ResampleWs *ws) {
  __shared__ double sh[1024];
  const int64_t per = (ntiles + 1023) / 1024;
  const int64_t lo = (int64_t)threadIdx.x * per, hi = min(lo + per, ntiles);
  double acc = 0.0;
  for (int64_t i = lo; i < hi; ++i) acc += tile_sums[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {  // 1024 sequential adds: negligible, and keeps the order fixed
    double run = 0.0;
    for (int t = 0; t < 1024; ++t) { const double v = sh[t]; sh[t] = run; run += v; }
    ws->changes = 0ull;
  }
  __syncthreads();
  double run = sh[threadIdx.x];
  for (int64_t i = lo; i < hi; ++i) { const double v = tile_sums[i]; tile_sums[i] = run; run += v; }
}

__global__ void __launch_bounds__(kScanThreads)
scan_write_bins_kernel(const float *__restrict__ w, int64_t N, const double *__restrict__ tile_offsets, float *__restrict__ bins) {
  __shared__ double sh_w[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  double v[kScanItems];
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < N) ? (double)__ldg(w + base + k) : 0.0;
    acc += v[k];
  }
  // exclusive scan of per-thread totals across the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double incl = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh_w[warp] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int q = 0; q < warp; ++q) woff += sh_w[q];
  double run = tile_offsets[blockIdx.x] + woff + (incl - acc);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    run += v[k];
    if (base + k < N) bins[base + k] = (float)run;  // torch.cumsum(fp32) on CPU: fp64 accumulate, fp32 store
  }
}

// ---- search ------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t systematic_id(const float *__restrict__ bins, int64_t N, double u0, float invN, int64_t slot) {
  // u = (u0 + fl32(fl32(1/N) * fl32(i))) % 1.0   (utils.py:113; the grid is evaluated in float32)
  double u = u0 + (double)__fmul_rn(invN, (float)slot);
  if (u >= 1.0) u -= 1.0;
  int64_t lo = 0, hi = N;  // first k with bins[k] >= u  == np.digitize(u, bins, right=True)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((double)__ldg(bins + mid) < u) lo = mid + 1; else hi = mid;
  }
  return lo == N ? N - 1 : lo;  // utils.py:119
}

__global__ void __launch_bounds__(256)
systematic_search_kernel(const float *__restrict__ bins, int64_t N, double u0, int64_t slot_lo, int64_t slot_hi,
                         int64_t *__restrict__ ids, ResampleWs *ws) {
  __shared__ int64_t s_ids[257];
  __shared__ int s_cnt[8];
  const float invN = (float)(1.0 / (double)N);
  const int64_t slot = slot_lo + (int64_t)blockIdx.x * 256 + threadIdx.x;
  int64_t id = -1;
  if (slot < slot_hi) {
    id = systematic_id(bins, N, u0, invN, slot);
    ids[slot - slot_lo] = id;
  }
  s_ids[threadIdx.x + 1] = id;
  if (threadIdx.x == 0) {
    const int64_t first = slot_lo + (int64_t)blockIdx.x * 256;
    s_ids[0] = first < slot_hi ? systematic_id(bins, N, u0, invN, first == 0 ? N - 1 : first - 1) : -1;
  }
  __syncthreads();
  int ch = (slot < slot_hi && s_ids[threadIdx.x] != id) ? 1 : 0;
  ch = __reduce_add_sync(0xffffffffu, ch);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = ch;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int q = 0; q < 8; ++q) tot += s_cnt[q];
    if (tot) atomicAdd(&ws->changes, (unsigned long long)tot);
  }
}

__global__ void copy_changes_kernel(const ResampleWs *ws, int64_t *out) { *out = (int64_t)ws->changes; }

// ---- gather / mean removal ---------------------------------------------------------------------
constexpr int kMaxRanks = 16;
struct SrcPtrs { const float *p[kMaxRanks]; };

__global__ void __launch_bounds__(256)
gather_rows_kernel(SrcPtrs src, int64_t rows_per_rank, const int64_t *__restrict__ ids, int64_t n_out, int D,
                   float *__restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_out) return;
  const int64_t id = __ldg(ids + row);
  const int64_t r = id / rows_per_rank;
  const float *__restrict__ s = src.p[r] + (id - r * rows_per_rank) * D;
  float *__restrict__ d = dst + row * D;
  for (int c = lane; c < D; c += 32) d[c] = s[c];
}

__global__ void __launch_bounds__(256)
remove_mean_kernel(const float *__restrict__ x, int64_t B, int n, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B) return;
  const int D = 3 * n;
  const float *__restrict__ s = x + row * D;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int c = lane; c < D; c += 32) {
    const float v = s[c];
    const int k = c % 3;
    acc[0] += k == 0 ? v : 0.f; acc[1] += k == 1 ? v : 0.f; acc[2] += k == 2 ? v : 0.f;
  }
  const float inv = 1.0f / (float)n;
  const float m0 = warp_sum(acc[0]) * inv, m1 = warp_sum(acc[1]) * inv, m2 = warp_sum(acc[2]) * inv;
  float *__restrict__ d = out + row * D;
  for (int c = lane; c < D; c += 32) {
    const int k = c % 3;
    d[c] = s[c] - (k == 0 ? m0 : (k == 1 ? m1 : m2));
  }
}

}  // namespace pita

using namespace pita;

extern "C" int64_t pita_resample_workspace_bytes(int64_t N) { return N < 0 ? 0 : ws_bytes(N); }

extern "C" int pita_softmax_clip(const float *logits, int64_t N, float *w, void *workspace, void *stream) {
  PITA_REQUIRE(logits && w && workspace, PITA_EINVAL, "softmax_clip: null pointer");
  PITA_REQUIRE(N > 0, PITA_EINVAL, "softmax_clip: N must be positive");
  PITA_REQUIRE(aligned16(workspace), PITA_EINVAL, "softmax_clip: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char *wsb = static_cast<char *>(workspace);
  ResampleWs *ws = reinterpret_cast<ResampleWs *>(wsb);
  float2 *partials = reinterpret_cast<float2 *>(wsb + ws_off_partials());
  const int nparts = (int)min((int64_t)kMaxPartials, (N + kSmThreads - 1) / kSmThreads);
  softmax_partials_kernel<<<nparts, kSmThreads, 0, st>>>(logits, N, partials);
  softmax_finalize_kernel<<<1, 32, 0, st>>>(partials, nparts, ws);
  softmax_clip_kernel<<<nparts, kSmThreads, 0, st>>>(logits, N, ws, w);
  PITA_CHECK_LAUNCH("softmax_clip");
  return PITA_OK;
}

extern "C" int pita_resample_systematic(const float *w, int64_t N, double u0, int64_t slot_lo, int64_t slot_hi,
                                        int64_t *ids_out, int64_t *changes_out, void *workspace, void *stream) {
  PITA_REQUIRE(w && ids_out && workspace, PITA_EINVAL, "resample: null pointer");
  PITA_REQUIRE(N > 0 && slot_lo >= 0 && slot_hi <= N && slot_lo <= slot_hi, PITA_EINVAL, "resample: bad slot range");
  PITA_REQUIRE(N < (1ll << 24), PITA_EINVAL, "resample: N >= 2^24 not representable on the reference's float32 offset grid");
  PITA_REQUIRE(u0 >= 0.0 && u0 < 1.0, PITA_EINVAL, "resample: u0 must be in [0,1)");
  PITA_REQUIRE(aligned16(workspace), PITA_EINVAL, "resample: workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char *wsb = static_cast<char *>(workspace);
  ResampleWs *ws = reinterpret_cast<ResampleWs *>(wsb);
  double *tiles = reinterpret_cast<double *>(wsb + ws_off_tiles());
  float *bins = reinterpret_cast<float *>(wsb + ws_off_bins(N));
  const int64_t ntiles = ws_tiles(N);
  scan_tile_sums_kernel<<<(unsigned)ntiles, kScanThreads, 0, st>>>(w, N, tiles);
  scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(tiles, ntiles, ws);
  scan_write_bins_kernel<<<(unsigned)ntiles, kScanThreads, 0, st>>>(w, N, tiles, bins);
  const int64_t nslots = slot_hi - slot_lo;
  if (nslots > 0)
    systematic_search_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(bins, N, u0, slot_lo, slot_hi, ids_out, ws);
  if (changes_out) copy_changes_kernel<<<1, 1, 0, st>>>(ws, changes_out);
  PITA_CHECK_LAUNCH("resample_systematic");
  return PITA_OK;
}

extern "C" int pita_gather_rows(const float *const *src_ranks_host, int n_ranks, int64_t rows_per_rank, const int64_t *ids,
                                int64_t n_out, int row_floats, float *dst, void *stream) {
  PITA_REQUIRE(src_ranks_host && ids && dst, PITA_EINVAL, "gather: null pointer");
  PITA_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxRanks, PITA_EINVAL, "gather: n_ranks must be in [1,%d]", kMaxRanks);
  PITA_REQUIRE(rows_per_rank > 0 && row_floats > 0 && n_out >= 0, PITA_EINVAL, "gather: bad sizes");
  if (n_out == 0) return PITA_OK;
  SrcPtrs sp;
  for (int r = 0; r < kMaxRanks; ++r) sp.p[r] = r < n_ranks ? src_ranks_host[r] : nullptr;
  for (int r = 0; r < n_ranks; ++r) PITA_REQUIRE(sp.p[r], PITA_EINVAL, "gather: null source pointer for rank %d", r);
  gather_rows_kernel<<<(unsigned)((n_out + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(sp, rows_per_rank, ids, n_out, row_floats, dst);
  PITA_CHECK_LAUNCH("gather_rows_kernel");
  return PITA_OK;
}

extern "C" int pita_remove_mean(const float *x, int64_t B, int n, float *x_out, void *stream) {
  PITA_REQUIRE(x && x_out, PITA_EINVAL, "remove_mean: null pointer");
  PITA_REQUIRE(n > 0 && B >= 0, PITA_EINVAL, "remove_mean: bad sizes");
  if (B == 0) return PITA_OK;
  remove_mean_kernel<<<(unsigned)((B + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, n, x_out);
  PITA_CHECK_LAUNCH("remove_mean_kernel");
  return PITA_OK;
}
