// Round-2 score + exact divergence of the 3-layer EGNN ("triangle" engine): shared definitions of the two phases.
//
// tr(d x_L / d y) is evaluated in the bilinear form of oracle/egnn_bilinear.py (checked there against vmap(jacrev),
// the reference's compute_divergence_exact, utils.py:43-51):
//   phase A  (egnn_tri_a.cu, thread = (particle, node) row): primal forward (-> score), forward-mode tangent of layer 0 and
//            reverse-mode cotangent of layer 2 as per-PAIR tables in a per-particle workspace, plus the "direct" part;
//   phase B  (egnn_tri_b.cu, thread = middle-layer edge (i, j), loop over the tangent node k): ONE dense 32x32 product per
//            (edge, k) on the tensor core, everything around it a dot product with a table row.
// Work per particle: n^3 products instead of the 3 n^2 (n-1) of the forward-mode kernel (egnn_rows.cu), and no layer-1
// edge cache: the only HBM/L2 traffic are the pair tables (2 MB per LJ-55 particle, written once, read through L2).
#pragma once
#include "egnn_rowops.cuh"

namespace pita {
namespace tri {

// ---- per-particle workspace (float offsets).  TS = sender-side pair table, streamed by phase B in k-chunks;
//      TR = receiver-side pair table, resident per tile; OM = omega_ik scratch of phase A.
constexpr int kTS = 44;  // [0,32) PB_jk = B1 omega_jk | [32,41) M'_jk (3x3, [b][a]) | pad
constexpr int kTR = 88;  // [0,32) PA_ik | [32,64) gamma_ki | [64,67) w(ki) | [67] alpha_i | [68,77) GXs | [77,86) M'_ik | pad
constexpr int trPA = 0, trGam = 32, trW = 64, trAlpha = 67, trGX = 68, trM = 77;
constexpr int tsPB = 0, tsM = 32;

template <int NP>
struct WS {
  static constexpr int64_t oTS = 0;                              // [k][j][kTS]
  static constexpr int64_t oTR = oTS + (int64_t)NP * NP * kTS;   // [i][k][kTR]
  static constexpr int64_t oOM = oTR + (int64_t)NP * NP * kTR;   // [i][k][32]
  static constexpr int64_t oY = oOM + (int64_t)NP * NP * 32;     // [n] float4  network input coordinates y
  static constexpr int64_t oX1 = oY + NP * 4;                    // [n] float4  x^1
  static constexpr int64_t oP1 = oX1 + NP * 4;                   // [n][32]     A1 h1 + b1
  static constexpr int64_t oQ1 = oP1 + NP * 32;                  // [n][32]     B1 h1
  static constexpr int64_t oOmg = oQ1 + NP * 32;                 // [n][3][32]  Omega_i[a] (own-direction tangent of h^1)
  static constexpr int64_t oAOm = oOmg + NP * 96;                // [n][3][32]  A1 Omega
  static constexpr int64_t oBOm = oAOm + NP * 96;                // [n][3][32]  B1 Omega
  static constexpr int64_t oGAgg = oBOm + NP * 96;               // [n][3][32]  cotangent on agg_i for output node i
  static constexpr int64_t oDirect = oGAgg + NP * 96;            // [n]         direct part of the trace, per node
  static constexpr int64_t oPartB = oDirect + ((NP + 3) / 4) * 4;  // [32]      phase-B partial sums (one per CTA tile)
  static constexpr int64_t kFloats = ((oPartB + 32 + 31) / 32) * 32;
};

int64_t workspace_floats_per_particle(int n);

}  // namespace tri
}  // namespace pita
