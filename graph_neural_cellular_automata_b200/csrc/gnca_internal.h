// Internal (non-ABI) declarations shared between the translation units of libgnca.so.
#pragma once
#include <stddef.h>
#include "gnca_common.cuh"

namespace gnca {

struct FwdWorkspace {
  double* partials;   // [B][nchunks][2]
  float* rowsum;      // [B][C][H]
  float* attn_w;      // [B][MAX_K]
  float* absmean;     // [B][H][W]
  void* tc2;          // tile masks / lists / partials of k_update_tc2 (update_tc2_workspace_bytes)
  uint32_t* actbits;  // [B][ceil(HW/32)] active bits of the step (k_compact -> k_apply)
  uint32_t* alivebits;   // [B][ceil(HW/32)] sender-alive bits of the step (k_compact -> k_update_tc)
  size_t bytes;
};

FwdWorkspace carve_fwd_workspace(void* base, const gnca_model& m, int B, int H, int W);
void fill_step_args(StepArgs& a, const gnca_model& m, int B, int H, int W);
int set_host_offsets(StepArgs& a, const int32_t* offsets_host, int k);
int dispatch_step_fwd(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const FwdWorkspace& ws,
                      float* attn_out, cudaStream_t st);
// recompute only the masked pre-norm update u (+ statistics) of a step: the forward half the backward needs
int dispatch_step_recompute(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a,
                            const FwdWorkspace& ws, cudaStream_t st);

int run_attn_prepass(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const FwdWorkspace& ws,
                     cudaStream_t st);
int run_attn_map(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const FwdWorkspace& ws,
                 float* attn_out, cudaStream_t st);
// zero-padded-shift attention backward (gnca_graph.cu): softmax / logits / pooled Q,K -> gparams (wq,bq,wk,bk,
// scaling) and the per-(b,c,row) additive term `grow` of dL/dx.  gw_part: [B][nparts][k] partial sums of
// dL/dw_i; rowsum / a.attn_w from run_attn_prepass on the same x.
struct AttnBwdScratch {
  float* gw_part;   // [B][nparts][GNCA_MAX_K]
  float* grow;      // [B][C][H]
  float* pw;        // [B][2*d*C + 2*d + 1]
};
size_t attn_bwd_scratch_bytes(const gnca_model& m, int B, int H, int nparts);
AttnBwdScratch carve_attn_bwd(void* base, const gnca_model& m, int B, int H, int nparts);
int run_attn_bwd(const gnca_model& m, const Packed& P, const float* packed, const StepArgs& a, const float* rowsum,
                 const AttnBwdScratch& sc, int nparts, float* gparams, cudaStream_t st, bool defer_reduce);
int run_attn_param_reduce(const gnca_model& m, int B, const AttnBwdScratch& sc, float* gparams, cudaStream_t st);

size_t graph_workspace_bytes(const gnca_model& m, int B, int H, int W);
size_t bwd_workspace_bytes(const gnca_model& m, int B, int H, int W);
// Backward of one step.  `a` describes the step (x_in, u, schedule...).  Weight-gradient partials live in the
// workspace: zero_partials clears them first, reduce_partials folds them into gparams at the end (a rollout
// clears at its first step and reduces after its last).
int run_step_bwd(const gnca_model& m, const Packed& P, const float* packed, const StepArgs& a, const float* stats,
                 const float* gout, float* gx, float* gparams, const FwdWorkspace& fws, void* bwd_ws, bool zero_partials,
                 bool reduce_partials, cudaStream_t st);

// tensor-core k_update (gnca_update_tc.cu): the balanced large-problem path for C in {16, 32}, hidden 128.
// glist / prefix: the global active list of k_compact / k_scan.  Sets a.npart (3 partial slots per chunk).
constexpr int kTcChunk = 1024;
bool update_tc_supported(const gnca_model& m, const StepArgs& a);
int launch_update_tc(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const uint16_t* glist,
                     const int* prefix, cudaStream_t st);

// dense-tile, TMA-staged tensor-core k_update (gnca_update_tc2.cu): H % 8 == 0, W % 16 == 0.  Finishes the GroupNorm
// statistics itself: a.stats_ready points at (mean, rstd) per sample afterwards (written to a.stats when that is set).
size_t update_tc2_workspace_bytes(int B, int H, int W);
bool update_tc2_supported(const gnca_model& m, const StepArgs& a);
int launch_update_tc2(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, void* ws, cudaStream_t st);

// cluster-resident forward rollout (gnca_resident.cu); GNCA_ERR_UNSUPPORTED when the configuration has no
// resident kernel (zero-padded graph shift, sample too large for the cluster's shared memory)
int run_resident_fwd(const gnca_model& m, const Packed& P, const float* packed, int B, int H, int W,
                     const gnca_schedule& sched, const float* x0, float* xT, float* hist, float* stats_hist,
                     float* u_hist, float* ping, float* pong, float* alpha_tmp, cudaStream_t st);

// replicated-state cluster rollout (gnca_rep.cu): every CTA of a sample's cluster holds the whole sample; first
// choice of the resident path.  scratch: >= B*C*H*W floats of device memory (overflow of the in-smem u buffer).
int run_rep_fwd(const gnca_model& m, const Packed& P, const float* packed, int B, int H, int W,
                const gnca_schedule& sched, const float* x0, float* xT, float* hist, float* stats_hist,
                float* u_hist, float* rec, uint32_t* masks, float* scratch, cudaStream_t st);
// resident BPTT on the records of run_rep_fwd (gnca_rep_bwd.cu)
size_t rep_bptt_bytes(const gnca_model& m, int B, int H, int W, int T);     // 0: configuration not supported
void rep_bptt_carve(void* base, int B, int H, int W, int T, float** rec, uint32_t** masks, float** stats,
                    float** hgh = nullptr);
size_t rep_bwd_workspace_bytes(const gnca_model& m, int B, int H, int W);
bool rep_bwd_supported(const gnca_model& m, int B, int H, int W, int k);    // a launch configuration of k_rep_bwd exists
int run_rep_bwd(const gnca_model& m, const Packed& P, const float* packed, int B, int H, int W,
                const gnca_schedule& sched, void* bptt, const float* gT, float* g0, float* gparams, void* workspace,
                cudaStream_t st);

}  // namespace gnca
