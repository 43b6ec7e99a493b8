/*
 * gnca.h -- C ABI of libgnca.so: the B200 (sm_100a) graph-NCA step / rollout kernels.
 *
 * This is the drop-in boundary for the hot path of Psylocibe23/Graph_Neural_Cellular_Automata.  The
 * reference has no FFI layer (it is pure PyTorch); what these entry points replace are the bodies of
 *
 *   FixedSobelPerception.forward      src/modules/perception.py:21-26            -> gnca_perception_fwd/_bwd
 *   NeuralCA._alive_mask              src/modules/nca.py:55-62                   -> gnca_alive_mask
 *   GraphAugmentation.forward         src/modules/graph_augmentation.py:104-169  -> gnca_graph_fwd/_bwd
 *   NeuralCA.forward                  src/modules/nca.py:64-105                  -> gnca_step_fwd/_bwd (GNCA_F_GRAPH clear)
 *   NeuralCAGraph.forward             src/modules/ncagraph.py:106-168            -> gnca_step_fwd/_bwd (GNCA_F_GRAPH set)
 *   the rollout loop                  src/training/train_graph_augmented_nca.py:302-324 -> gnca_rollout_fwd/_bwd
 *   loss_premult_rgba (+ its grad)    src/training/train_graph_augmented_nca.py:52-61,337-339 -> gnca_loss_premult_rgba
 *   grad normalise + Adam             src/training/train_graph_augmented_nca.py:370-375 -> gnca_normalize_adam
 *   multiplicative damage             src/utils/damage.py:16-98                  -> gnca_apply_mask
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer; tensors are fp32,
 *     contiguous NCHW exactly as the reference's modules take them;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, >0 = a cudaError_t, <0 = one of GNCA_ERR_*.  Nothing throws;
 *   - no call ever falls back to the host: an unsupported configuration returns GNCA_ERR_UNSUPPORTED.
 *
 * Parameters travel as ONE flat fp32 device buffer in the canonical order of gnca_param_layout()
 * (same orientation as the reference's state_dict tensors).  gnca_pack_weights() derives the
 * kernel-side buffer (adds transposed copies for the forward MLP); gradients come back in the canonical
 * layout.
 */
#ifndef GNCA_H_
#define GNCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNCA_VERSION 104

#define GNCA_ERR_ARG (-1)         /* null pointer / bad size */
#define GNCA_ERR_UNSUPPORTED (-2) /* shape or flag combination without a kernel */
#define GNCA_ERR_WORKSPACE (-3)   /* workspace too small */

#define GNCA_F_GRAPH 1u          /* NeuralCAGraph (else classic NeuralCA) */
#define GNCA_F_TORUS 2u          /* graph shift = torch.roll (graph_zero_padded_shift=False) */
#define GNCA_F_HIDDEN_ONLY 4u    /* ncagraph.py:98-101 */
#define GNCA_F_ALIVE_TO_ALIVE 8u /* graph_augmentation.py:116-117,130-133 */
#define GNCA_F_GROUPNORM 16u     /* use_groupnorm */

#define GNCA_MAX_K 64 /* chosen offsets per step */

typedef struct gnca_model {
  int32_t C;       /* n_channels: 4, 8, 16 or 32 */
  int32_t hidden;  /* update_hidden, multiple of 4, <= 256 */
  int32_t d_model; /* graph_d_model (used by the zero-padded-shift attention only), <= 64 */
  uint32_t flags;  /* GNCA_F_* */
  float update_gain;
  float alpha_thr;       /* model.alpha_thr  (pre / post alive) */
  float graph_alpha_thr; /* model.graph.alpha_thr (sender alive) */
  float gn_eps;          /* 1e-3 in the reference */
} gnca_model;

/* offsets (in floats) of each tensor inside the canonical flat parameter / gradient buffer; -1 = absent */
typedef struct gnca_layout {
  int64_t w1, b1, w2;   /* update_net.0.weight [hidden,3C], update_net.0.bias [hidden], update_net.2.weight [C,hidden] */
  int64_t gamma, beta;  /* norm.weight, norm.bias [C] (present even when GroupNorm is off) */
  int64_t wm, bm;       /* graph.msg_proj [C,C],[C] */
  int64_t wq, bq;       /* graph.query_proj [d,C],[d] */
  int64_t wk, bk;       /* graph.key_proj [d,C],[d] */
  int64_t scaling;      /* graph.scaling [1] */
  int64_t total;        /* floats in the canonical buffer */
  int64_t packed_total; /* floats in the packed (kernel-side) buffer */
} gnca_layout;

/* Per-step schedule of a rollout; every array lives on the DEVICE. */
typedef struct gnca_schedule {
  int32_t T;                  /* steps */
  int32_t k;                  /* offsets per step (0 for classic) */
  const float* fire_rate;     /* [T]   fire_rate of step t (>= 1 -> no mask) */
  const float* message_gain;  /* [T]   model.message_gain at step t (0 -> message skipped) */
  const int8_t* offsets;      /* [T][k][2] (dy,dx) drawn by random.sample at step t */
  const int32_t* steps;       /* [B] per-sample step counts (sample b advances while steps[b] > t) or NULL */
  const float* fire_u;        /* [T][B][H][W] uniforms (the reference's torch.rand draws) or NULL */
  uint64_t philox_seed;       /* used when fire_u == NULL: in-kernel Philox4x32-10 uniforms */
  uint64_t philox_offset;
  const float* damage;        /* multiplicative mask applied to x BEFORE step damage_step, or NULL; layout below */
  int32_t damage_step;
  int32_t max_offset;         /* max(|dy|,|dx|) over all offsets (halo depth of the resident kernel); 0 = unknown */
  int32_t damage_layout;      /* GNCA_DMG_DENSE [B][C][H][W] | GNCA_DMG_PLANE [B][H][W] on every channel |
                                 GNCA_DMG_PLANE_ALPHA [B][H][W] on the alpha channel only (salt & pepper) */
  int32_t reserved_;
} gnca_schedule;

#define GNCA_DMG_DENSE 0
#define GNCA_DMG_PLANE 1
#define GNCA_DMG_PLANE_ALPHA 2

/* Damage DESCRIPTOR (src/utils/damage.py:16-98, policy :101-138): one kind + one size for the whole batch, per-sample
 * geometry.  gnca_damage_plane evaluates it in closed form into a per-cell plane [B][H][W] (square / circle / stripes /
 * gaussian from the positions; alpha_drop / saltpepper from the caller's uniforms), which the rollout kernels multiply
 * in at `damage_step` (schedule.damage with layout GNCA_DMG_PLANE / GNCA_DMG_PLANE_ALPHA) -- no [B][C][H][W] mask. */
#define GNCA_DK_SQUARE 1      /* cutout_square_  :16-24   pos[b] = top-left (y, x), side `size`                         */
#define GNCA_DK_CIRCLE 2      /* cutout_circle_  :27-37   pos[b] = centre, radius `size`                                */
#define GNCA_DK_STRIPE_H 3    /* stripe_wipe_    :40-51   rows  [pos[0][0], +size) of EVERY sample                      */
#define GNCA_DK_STRIPE_V 4    /*                           cols  [pos[0][0], +size)                                      */
#define GNCA_DK_GAUSSIAN 5    /* gaussian_hole_  :87-98   clamp(1 - exp(-r2 / (2 (size*softness)^2)), 0, 1)             */
#define GNCA_DK_ALPHA_DROP 6  /* alpha_dropout_  :54-66   0 where rand < p and alpha > alpha_thr (hard: all channels)   */
#define GNCA_DK_SALTPEPPER 7  /* salt_pepper_alpha_ :69-73  0 where rand < p, alpha channel only                        */
typedef struct gnca_damage {
  int32_t kind;
  int32_t size;
  float softness;             /* gaussian */
  float p;                    /* alpha_drop / saltpepper */
  float alpha_thr;            /* alpha_drop */
  int32_t reserved_;
  const int64_t* pos;         /* device [B][2] (y, x) -- what torch.randint drew, never read back by the host */
  const float* rand;          /* device [B][H][W] uniforms (torch.rand_like(alpha)) for kinds 6, 7 */
} gnca_damage;
/* plane[b][y][x] of the descriptor; `state` ([B][C][H][W]) is read for alpha_drop only.  *layout_out = the schedule
 * layout the plane must be applied with. */
int gnca_damage_plane(const gnca_damage* d, int B, int C, int H, int W, const float* state_dev, float* plane_dev,
                      int32_t* layout_out, void* stream);
/* state *= plane (layout GNCA_DMG_PLANE or GNCA_DMG_PLANE_ALPHA), in place */
int gnca_apply_plane(int B, int C, int H, int W, float* state_dev, const float* plane_dev, int32_t layout, void* stream);

int gnca_version(void);
const char* gnca_error_string(int code);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
unsigned long long gnca_launch_count(void);
/* Optional per-kernel timing (CUDA events on the launching stream around each launch of the kernel family):
 * ids 0 k_update, 1 k_apply, 2 resident forward rollout, 3 k_bwd_mlp, 6 resident backward.  read() waits for the
 * recorded events, returns the summed duration and launch count since the last read, and clears them. */
int gnca_profile_enable(int on);
int gnca_profile_read(int kernel_id, double* total_ms, unsigned long long* launches);

int gnca_param_layout(const gnca_model* m, gnca_layout* out);
/* canonical flat params -> packed kernel-side buffer (packed_total floats) */
int gnca_pack_weights(const gnca_model* m, const float* params_dev, float* packed_dev, void* stream);

/* ------------------------------------------------------------------ single operators ------- */
/* perception.py:21-26 : x [B,C,H,W] -> y [B,3C,H,W] = [identity | sobel_x | sobel_y], zero halo */
int gnca_perception_fwd(int B, int C, int H, int W, const float* x_dev, float* y_dev, void* stream);
/* transpose: gy [B,3C,H,W] -> gx [B,C,H,W] */
int gnca_perception_bwd(int B, int C, int H, int W, const float* gy_dev, float* gx_dev, void* stream);
/* nca.py:55-62 : (maxpool3x3(x[:,3]) > thr) as float [B,1,H,W] */
int gnca_alive_mask(int B, int C, int H, int W, const float* x_dev, float thr, float* mask_dev, void* stream);

/* ------------------------------------------------------------------ one CA step ------------- */
size_t gnca_step_workspace_bytes(const gnca_model* m, int B, int H, int W);

/*
 * x_out = step(x_in).  fire_u_dev: [B,H,W] uniforms or NULL (required iff fire_rate < 1).
 * offsets_host: k (dy,dx) pairs, HOST memory (the module's random.sample result); k may be 0.
 * u_dev [B,C,H,W] and stats_dev [B,2] receive the masked pre-norm update and (mean, rstd): scratch for
 * inference, saved tensors for gnca_step_bwd.  attn_dev: optional [B,H,W] normalised attention map.
 */
int gnca_step_fwd(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                  const float* x_in_dev, float* x_out_dev,
                  const float* fire_u_dev, float fire_rate,
                  const int32_t* offsets_host, int k, float message_gain,
                  float* u_dev, float* stats_dev, float* attn_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * Backward of gnca_step_fwd.  gx_dev = dL/dx_in (overwritten); gparams_dev (canonical layout) is
 * ACCUMULATED into (+=) so a rollout can sum over steps; zero it first for a single step.
 */
int gnca_step_bwd(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                  const float* x_in_dev, const float* fire_u_dev, float fire_rate,
                  const int32_t* offsets_host, int k, float message_gain,
                  const float* u_dev, const float* stats_dev,
                  const float* gout_dev, float* gx_dev, float* gparams_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);

/* GraphAugmentation.forward alone: m [B,C,H,W] (+ optional attn [B,H,W]); bwd accumulates into gparams */
int gnca_graph_fwd(const gnca_model* m, const float* packed_dev, int B, int H, int W, const float* x_dev,
                   const int32_t* offsets_host, int k, float* msg_dev, float* attn_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);
int gnca_graph_bwd(const gnca_model* m, const float* packed_dev, int B, int H, int W, const float* x_dev,
                   const int32_t* offsets_host, int k, const float* gmsg_dev, float* gx_dev,
                   float* gparams_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ rollout ----------------- */
size_t gnca_rollout_workspace_bytes(const gnca_model* m, int B, int H, int W, int T);

/*
 * T steps in one call.  x_hist_dev: NULL for inference, else [T+1][B][C][H][W] receives x_0..x_T
 * (what the backward recomputes from; x_T is also written to xT_dev).  stats_hist_dev: [T][B][2] or NULL.
 * u_hist_dev: optional [T][B][C][H][W]; when given (together with stats_hist_dev) the masked pre-norm update of
 * the ACTIVE cells of every step is kept, and gnca_rollout_bwd skips recomputing the forward MLP.
 * `impl`: 0 = auto, 1 = streaming per-step kernels (any shape), 2 = cluster-resident kernels
 * (state lives in shared memory across all T steps; needs the sample to fit, torus or classic: the
 * replicated-state kernel when the whole sample fits one CTA, else the banded one), 3 = banded kernel only.
 */
int gnca_rollout_fwd(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                     const gnca_schedule* sched, const float* x0_dev, float* xT_dev,
                     float* x_hist_dev, float* stats_hist_dev, float* u_hist_dev,
                     void* workspace_dev, size_t workspace_bytes, int impl, void* stream);

/* BPTT: given gT = dL/dx_T and the x_hist of the forward, produce g0 = dL/dx_0 and ACCUMULATE gparams. */
int gnca_rollout_bwd(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                     const gnca_schedule* sched, const float* x_hist_dev, const float* stats_hist_dev,
                     const float* u_hist_dev,
                     const float* gT_dev, float* g0_dev, float* gparams_dev,
                     void* workspace_dev, size_t workspace_bytes, int impl, void* stream);

/*
 * Resident BPTT (16-channel / 128-hidden models, classic or torus graph, grids up to 2048 cells with W % 4 == 0).
 * Instead of the dense x_t / u_t history the forward keeps one compact record per ACTIVE cell and step (perception,
 * pre-norm update, gathered sender state ...) plus three bitmaps and the GroupNorm statistics per step, in the opaque
 * buffer `bptt_dev` (gnca_bptt_bytes; 0 = configuration not supported -> use gnca_rollout_fwd/_bwd).  The backward
 * is one cluster-resident kernel for the propagation of dL/dx plus one batched pass over all records for the weight
 * gradients.  x_hist_dev may be NULL (it is not needed by the backward).  gparams_dev is ACCUMULATED into.
 */
size_t gnca_bptt_bytes(const gnca_model* m, int B, int H, int W, int T);
int gnca_rollout_fwd_bptt(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                          const gnca_schedule* sched, const float* x0_dev, float* xT_dev, float* x_hist_dev,
                          void* bptt_dev, size_t bptt_bytes, void* workspace_dev, size_t workspace_bytes,
                          void* stream);
int gnca_rollout_bwd_bptt(const gnca_model* m, const float* packed_dev, int B, int H, int W,
                          const gnca_schedule* sched, void* bptt_dev, size_t bptt_bytes, const float* gT_dev,
                          float* g0_dev, float* gparams_dev, void* workspace_dev, size_t workspace_bytes,
                          void* stream);

/*
 * HOST function (no device work): T consecutive `random.sample(range(n), k)` draws of CPython's `random` module on
 * its own MT19937 state (`mt[624]`, `*mti` = random.getstate()[1][:624], [624]); state advanced in place so the caller
 * can `random.setstate` it back -- the python RNG stream stays exactly what T reference `forward` calls consume
 * (graph_augmentation.py:120-121).  out_idx: [T][k] indices into the offset table.
 */
int gnca_host_sample_indices(uint32_t* mt, int32_t* mti, int n, int k, int T, int32_t* out_idx);
/*
 * HOST function: the same T draws taken from `n_words` raw MT19937 outputs (the caller pulled them with
 * random.getrandbits(32 * n_words), first output = word 0); writes the chosen (dy, dx) pairs of `table` [n][2] to
 * out_off [T][k][2] and the number of words consumed to *used -- the caller restores the saved state and skips exactly
 * `used` outputs, so the python stream again equals T reference forward calls.  GNCA_ERR_UNSUPPORTED: block too short.
 */
int gnca_host_sample_offsets_words(const uint32_t* words, int n_words, const int8_t* table, int n, int k, int T,
                                   int8_t* out_off, int32_t* used);

/* ------------------------------------------------------------------ training glue ----------- */
/* per_sample[b] = mean_{4HW}([rgb*a, a] - target)^2 ; gx (optional) = d(scale * sum_b per_sample[b])/dx, [B,C,H,W] */
int gnca_loss_premult_rgba(int B, int C, int H, int W, const float* x_dev, const float* target_dev /*[4,H,W]*/,
                           float* per_sample_dev, float* gx_dev, float scale, void* stream);
/*
 * For each of n_seg parameter tensors (seg_off[i]..seg_off[i+1] in the flat buffers, host array of n_seg+1):
 * g /= (||g||_2 + 1e-8) unless seg_has_grad[i]==0 (skipped like a None grad), then torch.optim.Adam
 * with coupled L2 weight decay.  step is 1-based.
 */
int gnca_normalize_adam(float* params_dev, float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                        const int64_t* seg_off_host, const int32_t* seg_has_grad_host, int n_seg,
                        int normalize, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int64_t step, void* stream);
/* x *= mask (damage.py as multiplicative masks); mask [B,C,H,W] */
int gnca_apply_mask(int64_t n, float* x_dev, const float* mask_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNCA_H_ */
