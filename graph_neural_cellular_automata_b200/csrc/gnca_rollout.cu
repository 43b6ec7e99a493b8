// Rollout entry points (the loop of train_graph_augmented_nca.py:302-324 as ONE call).
//   impl 1 (streaming): per-step kernels of gnca_fwd.cu / gnca_bwd.cu driven by a device-resident schedule --
//                       no host synchronisation inside the loop, so the whole call can sit in a CUDA graph.
//   impl 2 (resident):  cluster kernel of gnca_resident.cu (state stays in shared memory across steps).
// BPTT stores only x_t (x_hist); u, the hidden layer, masks and the message are recomputed in the backward.
#include <cstdlib>
#include "gnca_common.cuh"
#include "gnca_internal.h"

namespace gnca {

__global__ void k_mul_inplace(size_t n, float* __restrict__ x, DamageView d, int C, int HW) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / HW;
    x[i] *= d.at((int)(bc / C), (int)(bc % C), (int)(i - bc * HW), C, HW);
  }
}

struct RolloutWorkspace {
  float* ping;     // [B][C][H][W]
  float* pong;
  float* u;        // [B][C][H][W]
  float* stats;    // [B][2]
  float* g_a;      // gradient ping-pong
  float* g_b;
  float* alpha_tmp;  // [B][H][W] (resident kernel)
  char* step_ws;   // forward + backward per-step workspace
  size_t step_bytes;
  size_t bytes;
};

static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

static RolloutWorkspace carve_rollout(void* base, const gnca_model& m, int B, int H, int W) {
  RolloutWorkspace r;
  const size_t N = (size_t)B * m.C * H * W * sizeof(float);
  char* p = reinterpret_cast<char*>(base);
  size_t o = 0;
  r.ping = reinterpret_cast<float*>(p + o); o = al(o + N);
  r.pong = reinterpret_cast<float*>(p + o); o = al(o + N);
  r.u = reinterpret_cast<float*>(p + o); o = al(o + N);
  r.g_a = reinterpret_cast<float*>(p + o); o = al(o + N);
  r.g_b = reinterpret_cast<float*>(p + o); o = al(o + N);
  r.stats = reinterpret_cast<float*>(p + o); o = al(o + (size_t)B * 2 * sizeof(float));
  r.alpha_tmp = reinterpret_cast<float*>(p + o); o = al(o + (size_t)B * H * W * sizeof(float));
  r.step_ws = p + o;
  r.step_bytes = carve_fwd_workspace(nullptr, m, B, H, W).bytes + bwd_workspace_bytes(m, B, H, W);
  o = al(o + r.step_bytes);
  r.bytes = o;
  return r;
}

static void schedule_to_args(StepArgs& a, const gnca_schedule& s, int t, int B, int H, int W) {
  a.t = t;
  a.k = s.k;
  a.fire_rate_dev = s.fire_rate;
  a.message_gain_dev = s.message_gain;
  a.offsets_dev = s.offsets;
  a.steps = s.steps;
  a.fire_u = s.fire_u ? s.fire_u + (size_t)t * B * H * W : nullptr;
  a.philox_seed = s.philox_seed;
  a.philox_offset = s.philox_offset;
}

}  // namespace gnca

using namespace gnca;

extern "C" {

size_t gnca_rollout_workspace_bytes(const gnca_model* m, int B, int H, int W, int T) {
  (void)T;
  if (!m || B <= 0 || H <= 0 || W <= 0) return 0;
  const size_t a = carve_rollout(nullptr, *m, B, H, W).bytes;
  const size_t b = (m->C == 16) ? rep_bwd_workspace_bytes(*m, B, H, W) : 0;
  return a > b ? a : b;
}

int gnca_rollout_fwd(const gnca_model* m, const float* packed_dev, int B, int H, int W, const gnca_schedule* sched,
                     const float* x0_dev, float* xT_dev, float* x_hist_dev, float* stats_hist_dev, float* u_hist_dev,
                     void* workspace_dev,
                     size_t workspace_bytes, int impl, void* stream) {
  if (!m || !packed_dev || !sched || !x0_dev || !xT_dev || !workspace_dev) return GNCA_ERR_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || sched->T < 0) return GNCA_ERR_ARG;
  if (!model_supported(*m)) return GNCA_ERR_UNSUPPORTED;
  const bool graph = (m->flags & GNCA_F_GRAPH) != 0;
  if (sched->T > 0 && (!sched->fire_rate || (graph && !sched->message_gain))) return GNCA_ERR_ARG;
  if (graph && sched->k > 0 && !sched->offsets) return GNCA_ERR_ARG;
  if (sched->k < 0 || sched->k > GNCA_MAX_K) return GNCA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  RolloutWorkspace r = carve_rollout(workspace_dev, *m, B, H, W);
  if (r.bytes > workspace_bytes) return GNCA_ERR_WORKSPACE;
  const size_t N = (size_t)B * m->C * H * W;
  const Packed P = make_packed(m->C, m->hidden, m->d_model, graph);
  FwdWorkspace fws = carve_fwd_workspace(r.step_ws, *m, B, H, W);
  const int T = sched->T;
  if (impl != 1 && T > 0) {
    // auto: small samples (the launch-latency-bound regime) go to the cluster-resident kernel
    const bool want = impl == 2 || impl == 3 || (size_t)H * W <= 16384;
    if (want) {
      const char* kind = getenv("GNCA_RESIDENT_KIND");      // development: "band" forces the banded kernel
      int rc = GNCA_ERR_UNSUPPORTED;
      if (impl != 3 && !(kind && kind[0] == 'b'))
        rc = run_rep_fwd(*m, P, packed_dev, B, H, W, *sched, x0_dev, xT_dev, x_hist_dev, stats_hist_dev, u_hist_dev,
                         nullptr, nullptr, r.ping, st);
      if (rc == GNCA_ERR_UNSUPPORTED)
        rc = run_resident_fwd(*m, P, packed_dev, B, H, W, *sched, x0_dev, xT_dev, x_hist_dev, stats_hist_dev,
                              u_hist_dev, r.ping, r.pong, r.alpha_tmp, st);
      if (rc != GNCA_ERR_UNSUPPORTED || impl == 2 || impl == 3) return rc;
    }
  }

  // where x_t lives: x_hist slices when a history is requested, else ping/pong (x_T straight into xT_dev)
  auto x_at = [&](int t) -> float* {
    if (x_hist_dev) return x_hist_dev + (size_t)t * N;
    if (t == T) return xT_dev;
    return (t & 1) ? r.pong : r.ping;
  };
  GNCA_CHECK_CUDA(cudaMemcpyAsync(x_at(0), x0_dev, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int t = 0; t < T; ++t) {
    if (sched->damage && t == sched->damage_step) {
      k_mul_inplace<<<(int)((N + 1023) / 1024 < 2368 ? (N + 1023) / 1024 : 2368), 256, 0, st>>>(N, x_at(t), DamageView{sched->damage, sched->damage_layout}, m->C, H * W);
      GNCA_LAUNCH_CHECK();
    }
    StepArgs a;
    fill_step_args(a, *m, B, H, W);
    schedule_to_args(a, *sched, t, B, H, W);
    if (!graph) a.k = 0;
    a.x_in = x_at(t);
    a.x_out = x_at(t + 1);
    a.u = u_hist_dev ? u_hist_dev + (size_t)t * N : r.u;
    a.stats = stats_hist_dev ? stats_hist_dev + (size_t)t * B * 2 : r.stats;
    int rc = dispatch_step_fwd(*m, P, packed_dev, a, fws, nullptr, st);
    if (rc) return rc;
  }
  if (x_hist_dev || T == 0)
    GNCA_CHECK_CUDA(cudaMemcpyAsync(xT_dev, x_at(T), N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int gnca_rollout_bwd(const gnca_model* m, const float* packed_dev, int B, int H, int W, const gnca_schedule* sched,
                     const float* x_hist_dev, const float* stats_hist_dev, const float* u_hist_dev,
                     const float* gT_dev, float* g0_dev,
                     float* gparams_dev, void* workspace_dev, size_t workspace_bytes, int impl, void* stream) {
  if (!m || !packed_dev || !sched || !x_hist_dev || !gT_dev || !g0_dev || !gparams_dev || !workspace_dev)
    return GNCA_ERR_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || sched->T < 0) return GNCA_ERR_ARG;
  if (!model_supported(*m)) return GNCA_ERR_UNSUPPORTED;
  (void)impl;             // the backward currently always takes the streaming kernels (x_hist layout is shared)
  const bool saved = u_hist_dev != nullptr && stats_hist_dev != nullptr;   // else u and the statistics are recomputed
  const bool graph = (m->flags & GNCA_F_GRAPH) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  RolloutWorkspace r = carve_rollout(workspace_dev, *m, B, H, W);
  if (r.bytes > workspace_bytes) return GNCA_ERR_WORKSPACE;
  const size_t N = (size_t)B * m->C * H * W;
  const Packed P = make_packed(m->C, m->hidden, m->d_model, graph);
  FwdWorkspace fws = carve_fwd_workspace(r.step_ws, *m, B, H, W);
  char* bws = r.step_ws + fws.bytes;
  const int T = sched->T;
  if (T == 0) {
    GNCA_CHECK_CUDA(cudaMemcpyAsync(g0_dev, gT_dev, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const float* gcur = gT_dev;
  for (int t = T - 1; t >= 0; --t) {
    StepArgs a;
    fill_step_args(a, *m, B, H, W);
    schedule_to_args(a, *sched, t, B, H, W);
    if (!graph) a.k = 0;
    a.x_in = x_hist_dev + (size_t)t * N;
    a.x_out = nullptr;
    const float* stats_t = r.stats;
    int rc = 0;
    if (saved) {
      a.u = const_cast<float*>(u_hist_dev) + (size_t)t * N;
      stats_t = stats_hist_dev + (size_t)t * B * 2;
    } else {
      a.u = r.u;
      a.stats = r.stats;
      rc = dispatch_step_recompute(*m, P, packed_dev, a, fws, st);
      if (rc) return rc;
    }
    float* gnext = (t == 0) ? g0_dev : (((T - 1 - t) & 1) ? r.g_b : r.g_a);
    rc = run_step_bwd(*m, P, packed_dev, a, stats_t, gcur, gnext, gparams_dev, fws, bws, t == T - 1, t == 0, st);
    if (rc) return rc;
    if (sched->damage && t == sched->damage_step) {
      k_mul_inplace<<<(int)((N + 1023) / 1024 < 2368 ? (N + 1023) / 1024 : 2368), 256, 0, st>>>(N, gnext, DamageView{sched->damage, sched->damage_layout}, m->C, H * W);
      GNCA_LAUNCH_CHECK();
    }
    gcur = gnext;
  }
  return 0;
}


/* ---- resident BPTT: records instead of the dense x_t / u_t history (gnca_rep.cu / gnca_rep_bwd.cu) ---- */
size_t gnca_bptt_bytes(const gnca_model* m, int B, int H, int W, int T) {
  if (!m || B <= 0 || H <= 0 || W <= 0 || T < 0) return 0;
  if ((m->flags & GNCA_F_GRAPH) && !(m->flags & GNCA_F_TORUS)) return 0;
  return rep_bptt_bytes(*m, B, H, W, T);
}

int gnca_rollout_fwd_bptt(const gnca_model* m, const float* packed_dev, int B, int H, int W, const gnca_schedule* sched,
                          const float* x0_dev, float* xT_dev, float* x_hist_dev, void* bptt_dev, size_t bptt_bytes,
                          void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!m || !packed_dev || !sched || !x0_dev || !xT_dev || !bptt_dev || !workspace_dev) return GNCA_ERR_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || sched->T < 0) return GNCA_ERR_ARG;
  if (!model_supported(*m)) return GNCA_ERR_UNSUPPORTED;
  const size_t need = gnca_bptt_bytes(m, B, H, W, sched->T);
  if (need == 0) return GNCA_ERR_UNSUPPORTED;
  if (need > bptt_bytes) return GNCA_ERR_WORKSPACE;
  RolloutWorkspace r = carve_rollout(workspace_dev, *m, B, H, W);
  if (r.bytes > workspace_bytes) return GNCA_ERR_WORKSPACE;
  const bool graph = (m->flags & GNCA_F_GRAPH) != 0;
  // the records are only worth writing if the resident backward can consume them for this batch / shape
  if (!rep_bwd_supported(*m, B, H, W, graph ? sched->k : 0)) return GNCA_ERR_UNSUPPORTED;
  const Packed P = make_packed(m->C, m->hidden, m->d_model, graph);
  float *rec, *stats;
  uint32_t* masks;
  rep_bptt_carve(bptt_dev, B, H, W, sched->T, &rec, &masks, &stats);
  if (sched->T == 0) {
    cudaMemcpyAsync(xT_dev, x0_dev, (size_t)B * m->C * H * W * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (x_hist_dev)
      cudaMemcpyAsync(x_hist_dev, x0_dev, (size_t)B * m->C * H * W * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return 0;
  }
  return run_rep_fwd(*m, P, packed_dev, B, H, W, *sched, x0_dev, xT_dev, x_hist_dev, stats, nullptr, rec, masks, r.ping,
                     (cudaStream_t)stream);
}

int gnca_rollout_bwd_bptt(const gnca_model* m, const float* packed_dev, int B, int H, int W, const gnca_schedule* sched,
                          void* bptt_dev, size_t bptt_bytes, const float* gT_dev, float* g0_dev, float* gparams_dev,
                          void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!m || !packed_dev || !sched || !bptt_dev || !gT_dev || !g0_dev || !gparams_dev || !workspace_dev) return GNCA_ERR_ARG;
  if (B <= 0 || H <= 0 || W <= 0 || sched->T < 0) return GNCA_ERR_ARG;
  if (!model_supported(*m)) return GNCA_ERR_UNSUPPORTED;
  const size_t need = gnca_bptt_bytes(m, B, H, W, sched->T);
  if (need == 0) return GNCA_ERR_UNSUPPORTED;
  if (need > bptt_bytes) return GNCA_ERR_WORKSPACE;
  if (rep_bwd_workspace_bytes(*m, B, H, W) > workspace_bytes) return GNCA_ERR_WORKSPACE;
  const bool graph = (m->flags & GNCA_F_GRAPH) != 0;
  const Packed P = make_packed(m->C, m->hidden, m->d_model, graph);
  if (sched->T == 0) {
    cudaMemcpyAsync(g0_dev, gT_dev, (size_t)B * m->C * H * W * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return 0;
  }
  return run_rep_bwd(*m, P, packed_dev, B, H, W, *sched, bptt_dev, gT_dev, g0_dev, gparams_dev, workspace_dev,
                     (cudaStream_t)stream);
}

/* ---- host-side schedule construction: T x random.sample(range(n), k) on CPython's MT19937 state ----
 * CPython: random.sample on a short population is a partial Fisher-Yates over a copy of the population driven by
 * _randbelow(m) = { k = m.bit_length(); do r = getrandbits(k) while (r >= m); }, getrandbits(k <= 32) =
 * genrand_uint32() >> (32 - k)   (Lib/random.py, Modules/_randommodule.c).  `mt` is the 624-word state, *mti the
 * position (the 625th element of random.getstate()[1]); both are advanced in place. */
int gnca_host_sample_indices(uint32_t* mt, int32_t* mti, int n, int k, int T, int32_t* out_idx) {
  if (!mt || !mti || !out_idx || n <= 0 || k < 0 || k > n || n > 4096 || T < 0) return GNCA_ERR_ARG;
  int pos = *mti;
  auto next_u32 = [&]() -> uint32_t {
    if (pos >= 624) {
      const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX = 0x9908b0dfu;
      int kk;
      for (kk = 0; kk < 624 - 397; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
      }
      for (; kk < 623; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
      }
      const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
      pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  };
  int32_t pool[4096];
  for (int t = 0; t < T; ++t) {
    for (int i = 0; i < n; ++i) pool[i] = i;
    for (int i = 0; i < k; ++i) {
      const uint32_t m = (uint32_t)(n - i);
      int bits = 0;
      for (uint32_t v = m; v; v >>= 1) ++bits;
      uint32_t r;
      do { r = next_u32() >> (32 - bits); } while (r >= m);
      out_idx[t * k + i] = pool[r];
      pool[r] = pool[n - i - 1];
    }
  }
  *mti = pos;
  return 0;
}

// The same draws from a block of raw MT19937 outputs the caller pulled with random.getrandbits(32 * n_words): no
// state conversion on the python side (getstate -> numpy -> setstate costs more than the sampling).  Writes the
// chosen offsets straight from the table; *used = words consumed (the caller restores the state and skips exactly
// that many outputs).  Returns GNCA_ERR_UNSUPPORTED when the block is too short (the caller retries with more).
int gnca_host_sample_offsets_words(const uint32_t* words, int n_words, const int8_t* table /*[n][2]*/, int n, int k,
                                   int T, int8_t* out_off /*[T][k][2]*/, int32_t* used) {
  if (!words || !table || !out_off || !used || n <= 0 || k < 0 || k > n || n > 4096 || T < 0) return GNCA_ERR_ARG;
  int pos = 0;
  int32_t pool[4096];
  for (int t = 0; t < T; ++t) {
    for (int i = 0; i < n; ++i) pool[i] = i;
    for (int i = 0; i < k; ++i) {
      const uint32_t m = (uint32_t)(n - i);
      int bits = 0;
      for (uint32_t v = m; v; v >>= 1) ++bits;
      uint32_t r;
      do {
        if (pos >= n_words) return GNCA_ERR_UNSUPPORTED;
        r = words[pos++] >> (32 - bits);
      } while (r >= m);
      const int src = pool[r];
      out_off[((size_t)t * k + i) * 2] = table[2 * src];
      out_off[((size_t)t * k + i) * 2 + 1] = table[2 * src + 1];
      pool[r] = pool[n - i - 1];
    }
  }
  *used = pos;
  return 0;
}
}  // extern "C"
