/*
 * pita_b200 — C ABI of the B200-native annealed-sampling hot path of taraak/pita.
 *
 * The reference has NO native interface for this path: the path sits behind
 * Python objects.  Each entry point below therefore replaces the body of one reference Python
 * function; the Python classes in pita_b200/ keep the reference's signatures and call these.
 * Reference paths are relative to /root/reference/pita/src/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch) owns all
 *     buffers, the library never allocates, frees or retains pointers past return;
 *   - all tensors are contiguous row-major fp32, particles-major: x[B][3n] with xyz interleaved;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it;
 *   - return 0 on success, a negative PITA_E* code otherwise (pita_last_error() has the text);
 *     the Python layer turns non-zero into RuntimeError, the reference's convention for this path
 *     being Python exceptions;
 *   - re-entrant per stream; no global mutable state besides the last-error string (thread local).
 */
#ifndef PITA_B200_H
#define PITA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PITA_OK 0
#define PITA_EINVAL -1   /* bad shape / null pointer / misaligned pointer            */
#define PITA_EUNSUP -2   /* unsupported n_particles / hidden size / layer count      */
#define PITA_ECUDA -3    /* CUDA launch error                                         */

#define PITA_ABI_VERSION 1

int pita_abi_version(void);
const char *pita_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Lennard-Jones target: log-prob (= -E/T) and force (= grad_x log-prob).
 * Replaces LennardJonesPotential._energy/_log_prob (energies/lennardjones_energy.py:121-155) and the
 * autograd force of LennardJonesEnergy.__call__ (:213-227).  E = energy_factor * sum_{i!=j}(r^-12 -
 * 2 r^-6) + oscillator_scale * 0.5 * sum_i |x_i - mean|^2,  r = sqrt(|x_i-x_j|^2 + 1e-6) (bgflow eps).
 * force may be NULL (energy only).  n in {13, 55}.
 * ------------------------------------------------------------------------------------------------ */
int pita_lj_energy_force(const float *x, int64_t B, int n, float temperature, float energy_factor,
                         float oscillator_scale, float *logp, float *force, void *stream);

/* ------------------------------------------------------------------------------------------------
 * EGNN denoiser (EGNN_dynamics, models/components/egnn_temp_conditioned.py:56-93,172-194,321-356) with
 * the egnn_temp.yaml configuration: hidden 32, SiLU, recurrent, tanh, attention, agg=sum, time and
 * temperature conditioning.  Weights are passed as ONE packed fp32 device buffer whose layout is
 * produced by pita_egnn_pack_floats / pita_b200.egnn_temp_conditioned.pack_state_dict (host side).
 *
 * The same three entry points (forward, energy, score_div) also serve the alanine-dipeptide denoiser
 * EGNN_dynamics_AD2_cat (models/components/egnn_dynamics_ad2_cat.py:11-218 over models/components/egnn.py:108-184):
 * (hidden, layers, n) = (64, 5, 22) with condition_beta, node features one_hot(atom type, 21 columns) ++ t ++ beta;
 * weight buffer from pita_b200.egnn_dynamics_ad2_cat.pack_state_dict_ad2 (pita_egnn_pack_floats(64, 5) floats).  That
 * network runs on the fp32 CUDA cores in every `mode` and needs no workspace.  Any other (hidden, layers, n) returns
 * PITA_EUNSUP.
 * ------------------------------------------------------------------------------------------------ */
/* number of floats in a packed weight buffer for `layers` E_GCL blocks of width `hidden` */
int64_t pita_egnn_pack_floats(int hidden, int layers);

/* vel[B][3n] = EGNN_dynamics.forward(tcond[B], y[B][3n], beta[B])   (egnn_temp_conditioned.py:56-93) */
int pita_egnn_forward(const float *wpack, int hidden, int layers, int n, const float *tcond, const float *y,
                      const float *beta, int64_t B, float *vel, void *stream);

/* EnergyNet.forward_energy / .forward / d/dh (models/components/energy_net.py:14-62, pin=False) in one
 * pass: energy[B], grad_x[B][3n] (NULL to skip), dE_dh[B] (NULL to skip).  dU/dt of sdes.py:218 is
 * dE_dh * dh/dt(t).  ht[B] = h(t) per particle, beta[B]. */
int pita_egnn_energy(const float *wpack, int hidden, int layers, int n, const float *ht, const float *x,
                     const float *beta, int64_t B, float *energy, float *grad_x, float *dE_dh, void *stream);

/* compute_laplacian_exact of EnergyNet.forward_energy (models/components/utils.py:68-77, called from sdes.py:204-216
 * when the SDE has no score net): laplacian[B] = tr(Hess_x E)(ht[B], x[B][3n], beta[B]), pin=False, without the
 * precondition_beta factor (the host wrapper applies it).  Exact (second-order forward Taylor mode, fp32); built for
 * hidden=32, layers=3, n in {13, 55}; returns PITA_EUNSUP otherwise. */
int pita_egnn_energy_laplacian(const float *wpack, int hidden, int layers, int n, const float *ht, const float *x,
                               const float *beta, int64_t B, float *laplacian, void *stream);

/* ScoreNet.forward (models/components/score_net.py:13-43) and the exact divergence
 * tr(d score / d x) of compute_divergence_exact (models/components/utils.py:43-51):
 * score[B][3n], div[B] (NULL to skip the divergence).
 * mode selects how the dense tangent contraction of the divergence is evaluated:
 *   PITA_DIV_FP32   fp32 CUDA cores (reference-accurate, slowest)
 *   PITA_DIV_3XTF32 tcgen05 tensor cores, error-compensated TF32 (the round-1 default): weights and primal operand rows split
 *                   hi + lo (3xTF32), tangent operand rows rounded once against the exactly split weights; measured
 *                   divergence error <= 1.2e-5 relative, score error as fp32
 *   PITA_DIV_TF32   tcgen05 tensor cores, plain TF32 for the divergence (looser, stated bound 5e-2 relative on the
 *                   divergence: measured 5e-4 on LJ-13, 3e-2 on LJ-55); the score stays fp32-accurate (3xTF32)
 * The tensor-core modes need `workspace` (device, >= pita_egnn_score_div_workspace_bytes(n, mode) bytes: 255 MB for
 * n = 13, 1.08 GB for n = 55 — per-team scratch and the layer-1 edge cache of the divergence passes; independent of B;
 * its contents are meaningless between calls, so one buffer per stream can be shared by every call).
 *   PITA_DIV_BILINEAR (round 2, the Python default) tcgen05 tensor cores, the trace in its bilinear form: forward-mode
 *                   tangent through layer 0, reverse-mode cotangent through layer 2, ONE dense product per (middle-layer
 *                   edge, tangent node) instead of three (oracle/egnn_bilinear.py); primal products 3xTF32, the n^3
 *                   products plain TF32 (they carry < 1e-3 of the trace; measured divergence error < 1e-5 relative).
 *                   workspace: 148 x (particles per phase-A CTA) x 2.1 MB (n = 55) / 118 KB (n = 13), 128-byte aligned;
 *                   smaller workspaces are accepted (>= one particle) and lower the batch per launch pair. */
#define PITA_DIV_FP32 0
#define PITA_DIV_3XTF32 1
#define PITA_DIV_TF32 2
#define PITA_DIV_BILINEAR 3
int64_t pita_egnn_score_div_workspace_bytes(int n, int mode);
/* Introspection for the tests: float offsets of the per-particle workspace of PITA_DIV_BILINEAR (kFloats, oTS, oTR, oOM, oY,
 * oX1, oP1, oQ1, oOmg, oAOm, oBOm, oGAgg, oDirect, oPartB, kTS, kTR; csrc/egnn_tri.cuh).  Returns the number of entries. */
int64_t pita_egnn_tri_workspace_layout(int n, int64_t *out, int max_out);
int pita_egnn_score_div(const float *wpack, int hidden, int layers, int n, const float *ht, const float *x,
                        const float *beta, int64_t B, float *score, float *div, int mode, void *workspace,
                        int64_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Euler-Maruyama + Feynman-Kac step.  Replaces, for one step over the rank-local particles:
 *   VEReverseSDE.f's drift assembly (models/components/sdes.py:168-227),
 *   VEReverseSDE.diffusion (:245-251),  euler_maruyama_step's update (sde_integration.py:347-349),
 *   the start/end gating (:278-282) and remove_mean (utils/data_utils.py:4-26; sde_integration.py:148).
 *     drift_X = gamma*(-gradU)*g2/2 + gamma*(score*g2/2)          (debias)   | gamma*score*g2 (no debias)
 *     x_out   = remove_mean(x + drift_X*dt + (noise_scale*noise)*sqrt_dt)      unless freeze_x
 *     a_raw   = gamma^2*<-gradU, score*g2/2> + gamma*(div*g2/2) + gamma*(dE_dh*dh_dt) + dgamma*energy
 * a_raw[B] is then clamped per chunk by pita_fk_quantile_accumulate.  noise==NULL draws N(0,1) in-kernel
 * (Philox4x32-10, seed/offset) instead of reading a materialised tensor.  gradU/div/dE_dh/energy may be
 * NULL when debias==0.  x_out may alias x.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  float g2;          /* g(t)^2                         */
  float gamma;       /* annealing factor gamma(t)       */
  float dgamma_dt;   /* d gamma / dt                    */
  float dh_dt;       /* dh/dt(t)  (chain rule for dU/dt)*/
  float dt;          /* time_range / num_steps          */
  float sqrt_dt;     /* float(np.sqrt(dt))              */
  float noise_scale; /* diffusion_scale * g(t)          */
  int debias;        /* 1: FK-debiased drift; 0: f_not_debiased (sdes.py:117-128) */
  int freeze_x;      /* 1: step < start_resampling_step -> x_out = remove_mean(x) */
  int remove_mean;   /* should_mean_free                */
  uint64_t seed;     /* Philox key   (noise == NULL)    */
  uint64_t offset;   /* Philox counter offset, e.g. step index */
} pita_sde_params;

int pita_sde_fk_step(const float *x, const float *gradU, const float *score, const float *noise,
                     const float *div, const float *dE_dh, const float *energy, int64_t B, int n,
                     const pita_sde_params *params_host, float *x_out, float *a_raw, void *stream);

/* Per-chunk 0.9-quantile clamp of the FK drift and log-weight accumulation:
 *   drift_A = clamp(a_raw, max=torch.quantile(a_raw[chunk], q))   (sdes.py:230, one quantile per
 *   inference_batch_size chunk, sde_integration.py:312-343);   a_out = a + drift_A*dt, or 0 when zero_a
 *   (step < start or step >= end, sde_integration.py:278-282).  drift_A_out may be NULL. chunk <= 8192.
 * With a == NULL the clamped values themselves are written to a_out (used by the end resample, :179). */
int pita_fk_quantile_accumulate(const float *a_raw, const float *a, int64_t B, int chunk, float q, float dt,
                                int zero_a, float *a_out, float *drift_A_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Systematic resampling.  Replaces sample_cat_sys (models/components/utils.py:111-120) and the row
 * gather x_next[choice] (sde_integration.py:293).
 * ------------------------------------------------------------------------------------------------ */
/* bytes of scratch needed by the two calls below for N particles */
int64_t pita_resample_workspace_bytes(int64_t N);

/* w[N] = clip(softmax(logits), 1e-6, 1)  — not renormalised (utils.py:114) */
int pita_softmax_clip(const float *logits, int64_t N, float *w, void *workspace, void *stream);

/* ids[slot_lo..slot_hi) of the N-particle systematic resample with offset u0 (fp64):
 *   bins = fp32(cumsum in fp64 of w);  u_i = (u0 + fl32(fl32(1/N)*i)) mod 1;  ids_i = #{bins < u_i}, N -> N-1.
 * ids_out has slot_hi-slot_lo entries (int64).  changes_out (int64[1], may be NULL) receives the number
 * of slots in the range whose id differs from the previous slot's (cyclically) — summed over ranks this
 * is len(np.unique(choice)) (sde_integration.py:295) unless it is 0, which means 1. */
int pita_resample_systematic(const float *w, int64_t N, double u0, int64_t slot_lo, int64_t slot_hi,
                             int64_t *ids_out, int64_t *changes_out, void *workspace, void *stream);

/* dst[i][:] = src_r[local][:] where global id = ids[i], owner r = id / rows_per_rank, local = id % rows_per_rank
 * and src_r = src_ranks_host[r] (device pointers, possibly NVLink peer memory of other GPUs).  With
 * n_ranks == 1 this is the plain gather x[choice].  src_ranks_host is a HOST array of n_ranks pointers. */
int pita_gather_rows(const float *const *src_ranks_host, int n_ranks, int64_t rows_per_rank, const int64_t *ids,
                     int64_t n_out, int row_floats, float *dst, void *stream);

/* x_out = remove_mean(x) (utils/data_utils.py:4-26); may alias */
int pita_remove_mean(const float *x, int64_t B, int n, float *x_out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Post-processing on the target: negative_time_descent step and MALA accept/reject
 * (sde_integration.py:28-45, 353-470) built on pita_lj_energy_force.
 * ------------------------------------------------------------------------------------------------ */
/* x_out = remove_mean(x + force*dt [+ noise*sqrt(2 dt)])  (sde_integration.py:353-360) */
int pita_descent_step(const float *x, const float *force, const float *noise, int64_t B, int n, float dt,
                      int remove_mean, float *x_out, void *stream);

/* MALA: proposal x_prop = x + 0.5*dt*force + sqrt(dt)*noise and log q(x_prop|x) (mala_proposal :28-38) */
int pita_mala_propose(const float *x, const float *force, const float *noise, int64_t B, int n, float dt,
                      float *x_prop, float *log_q_fwd, void *stream);

/* MALA accept/reject (:40-45, 379-396): log q(x|x_prop) from force_prop, ratio, accept with log(uniform)
 * < ratio; writes x, logp in place, accepted[B] (0/1 as float), optional mean removal. */
int pita_mala_accept(float *x, float *logp, const float *x_prop, const float *logp_prop, const float *force_prop,
                     const float *log_q_fwd, const float *uniform, int64_t B, int n, float dt, int remove_mean,
                     float *accepted, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Self-test of the sm_100a tensor-core plumbing (tcgen05.mma kind::tf32, TMEM, mbarrier) used by the EGNN
 * tangent kernel: D[128][32] = A[128][32] * B[32][32]^T on one CTA.  split=1: 3xTF32 (fp32-accurate).
 * No reference counterpart; exists so the tests can pin descriptor / swizzle conventions on hardware.
 * ------------------------------------------------------------------------------------------------ */
int pita_umma_selftest(const float *A, const float *B, float *D, int split, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PITA_B200_H */
