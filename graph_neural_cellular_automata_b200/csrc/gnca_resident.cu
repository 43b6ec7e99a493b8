// Cluster-resident forward rollout: ONE launch runs all T steps of a batch.
//
// Each sample is owned by a thread-block cluster of NC CTAs (NC = 16/8/4/2/1 chosen so that B*NC fills the 148
// SMs).  Weights live in shared memory for the whole rollout; the per-step pipeline of the streaming path
// (alive/fire test -> compaction of active cells -> perception -> MLP (+graph message) -> GroupNorm statistics ->
// bounded update -> post-alive gate) runs inside the kernel with three hardware cluster barriers per step instead
// of two kernel launches, and the per-sample GroupNorm reduction goes through distributed shared memory.
// Cells are dealt to the CTAs of a cluster in 32-cell segments round-robin, so the (spatially clustered) active
// cells are balanced over the CTAs.  The state itself is exchanged through an L2-resident global buffer
// (x_t / x_{t+1} ping-pong, or the x_hist slices when a history is requested for BPTT).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include "gnca_common.cuh"
#include "gnca_internal.h"

namespace cg = cooperative_groups;

namespace gnca {

constexpr int kRThreads = 256;
constexpr int kRWarps = kRThreads / 32;
constexpr int kMB = 32;          // active cells per MLP batch
constexpr int kMBP = kMB + 4;    // padded row stride of the staged [feature][cell] tiles

struct ResidentArgs {
  StepArgs s;                 // model scalars + schedule pointers (t / fire_u are set per step in-kernel)
  int T, NC, maxown, nseg;
  const float* fire_u_base;   // [T][B][H][W] or null
  const float* x0;
  float* xT;
  float* hist;                // [T+1][B][C][HW] or null
  float* ping;                // [B][C][HW]
  float* pong;
  float* alpha_tmp;           // [B][HW]  pre-gate updated alpha
  float* stats_hist;          // [T][B][2] or null
  const float* damage;        // [B][C][HW] or null
  int damage_step;
};

// coherent (L2) loads of state written by other CTAs during the kernel
__device__ __forceinline__ float ldc(const float* p) { return __ldcg(p); }

__device__ __forceinline__ bool alive_at_c(const float* alpha, int y, int x, int H, int W, float thr) {
  float m = -INFINITY;
#pragma unroll
  for (int i = -1; i <= 1; ++i) {
    const int yy = y + i;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
      const int xx = x + j;
      if (xx < 0 || xx >= W) continue;
      m = fmaxf(m, ldc(alpha + yy * W + xx));
    }
  }
  return m > thr;
}

template <int C>
__global__ void __launch_bounds__(kRThreads, 1) k_resident_fwd(ResidentArgs R, Packed P, int hid,
                                                               const float* __restrict__ packed) {
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int C3 = 3 * C;
  StepArgs a = R.s;
  const int NC = R.NC;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / NC;
  const int H = a.H, W = a.W, HW = H * W;
  const bool graph = (a.flags & GNCA_F_GRAPH) != 0;
  const bool gn = (a.flags & GNCA_F_GROUPNORM) != 0;
  const bool torus = (a.flags & GNCA_F_TORUS) != 0;
  const bool a2a = (a.flags & GNCA_F_ALIVE_TO_ALIVE) != 0;
  const int c_lo = ((a.flags & GNCA_F_HIDDEN_ONLY) && C >= 4) ? 4 : 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int maxown = R.maxown;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sW1T = reinterpret_cast<float*>(smem_raw);
  float* sb1 = sW1T + C3 * hid;
  float* sW2T = sb1 + pad4(hid);
  float* sWmT = sW2T + hid * C;
  float* sbm = sWmT + (graph ? C * C : 0);
  float* Yt = sbm + (graph ? C : 0);          // [3C][kMBP]
  float* XSt = Yt + C3 * kMBP;                // [C][kMBP]
  float* AS = XSt + C * kMBP;                 // [kMBP]
  float* MSGt = AS + kMBP;                    // [C][kMBP]   gain * tanh(agg) on gated channels, else 0
  float* Ht = MSGt + C * kMBP;                // [hid][kMBP]
  float* RED = Ht + hid * kMBP;               // [kRWarps][C][kMB] layer-2 partials
  float* Ut = RED + kRWarps * C * kMB;        // [C][maxown]  masked pre-norm update of this CTA's active cells
  int* s_q = reinterpret_cast<int*>(Ut + (size_t)C * maxown);   // [kMB][GNCA_MAX_K] sender cell or -1
  int* s_actlist = s_q + kMB * GNCA_MAX_K;    // [maxown] local index of active cells (compacted)
  int* s_slot = s_actlist + maxown;           // [maxown] slot of a local cell in the active list or -1
  __shared__ double s_part[2];                // this CTA's (sum u, sum u^2); read by peers through DSMEM
  __shared__ double s_wred[kRWarps][2];
  __shared__ float s_sc[C], s_bi[C], s_idle[C], s_stat[2];
  __shared__ int s_wcount[kRWarps], s_wbase[kRWarps + 1];

  block_copy(sW1T, packed + P.w1t, C3 * hid);
  block_copy(sb1, packed + P.b1, pad4(hid));
  block_copy(sW2T, packed + P.w2t, hid * C);
  if (graph) { block_copy(sWmT, packed + P.wmt, C * C); block_copy(sbm, packed + P.bm, C); }

  // cells owned by this CTA: 32-cell segments dealt round-robin over the cluster
  const int nseg = R.nseg;
  const int my_nseg = (nseg - rank + NC - 1) / NC;      // segments rank, rank+NC, ...
  const int nown = my_nseg * 32;                          // local index space (tail cells may be >= HW)
  auto global_cell = [&](int local) { return ((local >> 5) * NC + rank) * 32 + (local & 31); };

  const size_t sample_off = (size_t)b * C * HW;
  auto x_ptr = [&](int t) -> float* {
    if (R.hist) return R.hist + (size_t)t * a.B * C * HW + sample_off;
    return ((t & 1) ? R.pong : R.ping) + sample_off;
  };
  // x_0: copy own cells (each CTA its own; a cluster barrier publishes them)
  {
    float* x0d = x_ptr(0);
    const float* x0s = R.x0 + sample_off;
    for (int i = tid; i < nown * C; i += kRThreads) {
      const int c = i / nown, l = i - c * nown;
      const int cell = global_cell(l);
      if (cell < HW) x0d[(size_t)c * HW + cell] = x0s[(size_t)c * HW + cell];
    }
  }
  __syncthreads();
  cluster.sync();

  const int my_steps = a.steps ? min(a.steps[b], R.T) : R.T;
  float* alpha_tmp = R.alpha_tmp + (size_t)b * HW;

  for (int t = 0; t < R.T; ++t) {
    float* xc = x_ptr(t);
    float* xn = x_ptr(t + 1);
    if (R.damage && t == R.damage_step) {      // multiplicative damage applied to x_t in place
      const float* D = R.damage + sample_off;
      for (int i = tid; i < nown * C; i += kRThreads) {
        const int c = i / nown, l = i - c * nown;
        const int cell = global_cell(l);
        if (cell < HW) xc[(size_t)c * HW + cell] = ldc(xc + (size_t)c * HW + cell) * D[(size_t)c * HW + cell];
      }
      __syncthreads();
      cluster.sync();
    }
    if (t >= my_steps) {                        // frozen sample: state passes through (whole cluster agrees)
      for (int i = tid; i < nown * C; i += kRThreads) {
        const int c = i / nown, l = i - c * nown;
        const int cell = global_cell(l);
        if (cell < HW) xn[(size_t)c * HW + cell] = ldc(xc + (size_t)c * HW + cell);
      }
      continue;
    }
    a.t = t;
    a.fire_u = R.fire_u_base ? R.fire_u_base + (size_t)t * a.B * HW : nullptr;
    const float fr = step_fire_rate(a);
    const float gain_m = graph ? step_message_gain(a) : 0.f;
    const bool msg_on = graph && gain_m != 0.f && a.k > 0;
    const float* alpha = xc + 3 * HW;

    // ---- P1: alive & fire on own cells, deterministic compaction --------------------------------------
    {
      int cnt = 0;
      // each warp handles local cells [warp*per, (warp+1)*per) in 32-wide strips
      const int per = ((nown + kRWarps * 32 - 1) / (kRWarps * 32)) * 32;
      const int lo = warp * per;
      for (int base = lo; base < lo + per; base += 32) {
        const int l = base + lane;
        bool act = false;
        if (l < nown) {
          const int cell = global_cell(l);
          if (cell < HW) {
            const int y = cell / W, x = cell - y * W;
            act = alive_at_c(alpha, y, x, H, W, a.alpha_thr) && fires(a, fr, b, cell);
          }
          s_slot[l] = act ? 1 : -1;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, act));
      }
      if (lane == 0) s_wcount[warp] = cnt;
      __syncthreads();
      if (tid == 0) {
        int s = 0;
        for (int w = 0; w < kRWarps; ++w) { s_wbase[w] = s; s += s_wcount[w]; }
        s_wbase[kRWarps] = s;
      }
      __syncthreads();
      int pos = s_wbase[warp];
      for (int base = lo; base < lo + per; base += 32) {
        const int l = base + lane;
        const bool act = l < nown && s_slot[l] > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (act) {
          const int slot = pos + __popc(bal & ((1u << lane) - 1u));
          s_actlist[slot] = l;
          s_slot[l] = slot;
        }
        pos += __popc(bal);
      }
      __syncthreads();
    }
    const int nact = s_wbase[kRWarps];

    // ---- P2: perception + MLP + message on the active cells, in batches of kMB --------------------------
    float ps1 = 0.f, ps2 = 0.f;
    for (int base = 0; base < nact; base += kMB) {
      const int nb = min(kMB, nact - base);
      // 2a: sender table (cell, offset) and perception (cell, channel)
      if (msg_on) {
        for (int i = tid; i < nb * a.k; i += kRThreads) {
          const int cl = i / a.k, oi = i - cl * a.k;
          const int cell = global_cell(s_actlist[base + cl]);
          const int y = cell / W, x = cell - y * W;
          int dy, dx, qy, qx, q = -1;
          step_offset(a, oi, dy, dx);
          if (sender_of(y, x, dy, dx, H, W, torus, qy, qx) &&
              (!a2a || alive_at_c(alpha, qy, qx, H, W, a.graph_alpha_thr)))
            q = qy * W + qx;
          s_q[cl * GNCA_MAX_K + oi] = q;
        }
      }
      for (int i = tid; i < kMB * C; i += kRThreads) {
        const int cl = i % kMB, c = i / kMB;
        float vid = 0.f, vsx = 0.f, vsy = 0.f;
        if (cl < nb) {
          const int cell = global_cell(s_actlist[base + cl]);
          const int y = cell / W, x = cell - y * W;
          const bool up = y > 0, dn = y < H - 1, lf = x > 0, rt = x < W - 1;
          const float* p = xc + (size_t)c * HW + cell;
          const float a00 = (up && lf) ? ldc(p - W - 1) : 0.f, a01 = up ? ldc(p - W) : 0.f,
                      a02 = (up && rt) ? ldc(p - W + 1) : 0.f;
          const float a10 = lf ? ldc(p - 1) : 0.f, a12 = rt ? ldc(p + 1) : 0.f;
          const float a20 = (dn && lf) ? ldc(p + W - 1) : 0.f, a21 = dn ? ldc(p + W) : 0.f,
                      a22 = (dn && rt) ? ldc(p + W + 1) : 0.f;
          vid = ldc(p);
          vsx = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
          vsy = (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
        }
        Yt[c * kMBP + cl] = vid;
        Yt[(C + c) * kMBP + cl] = vsx;
        Yt[(2 * C + c) * kMBP + cl] = vsy;
      }
      __syncthreads();
      // 2b: gather the (alive) senders' state: xs = sum_i w_i x(q_i), as = sum_i w_i
      if (msg_on) {
        const float wuni = 1.0f / (float)a.k;
        for (int i = tid; i < kMB * (C + 1); i += kRThreads) {
          const int cl = i % kMB, c = i / kMB;
          float v = 0.f;
          if (cl < nb) {
            for (int oi = 0; oi < a.k; ++oi) {
              const int q = s_q[cl * GNCA_MAX_K + oi];
              if (q < 0) continue;
              const float w = a.attn_w ? a.attn_w[(size_t)b * a.k + oi] : wuni;
              v = (c < C) ? fmaf(w, ldc(xc + (size_t)c * HW + q), v) : v + w;
            }
          }
          if (c < C) XSt[c * kMBP + cl] = v; else AS[cl] = v;
        }
      }
      // 2c: layer 1, warp = group of 4 cells, lane = group of 4 hidden units (loop if hid > 128)
      {
        const int ngroups = (nb + 3) >> 2;
        for (int cgp = warp; cgp < ngroups; cgp += kRWarps) {
          for (int j = lane * 4; j < hid; j += 128) {
            float acc[4][4];
            const float4 bb = *reinterpret_cast<const float4*>(sb1 + j);
#pragma unroll
            for (int m = 0; m < 4; ++m) { acc[m][0] = bb.x; acc[m][1] = bb.y; acc[m][2] = bb.z; acc[m][3] = bb.w; }
#pragma unroll 8
            for (int k = 0; k < C3; ++k) {
              const float4 yv = *reinterpret_cast<const float4*>(Yt + k * kMBP + 4 * cgp);
              const float4 w = *reinterpret_cast<const float4*>(sW1T + k * hid + j);
              const float ym[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                acc[m][0] = fmaf(ym[m], w.x, acc[m][0]); acc[m][1] = fmaf(ym[m], w.y, acc[m][1]);
                acc[m][2] = fmaf(ym[m], w.z, acc[m][2]); acc[m][3] = fmaf(ym[m], w.w, acc[m][3]);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float4 hv;
              hv.x = fmaxf(acc[0][jj], 0.f); hv.y = fmaxf(acc[1][jj], 0.f);
              hv.z = fmaxf(acc[2][jj], 0.f); hv.w = fmaxf(acc[3][jj], 0.f);
              *reinterpret_cast<float4*>(Ht + (j + jj) * kMBP + 4 * cgp) = hv;
            }
          }
        }
      }
      __syncthreads();
      // message projection: MSGt[c][cl] = gain * tanh(bm[c]*as + sum_ci Wm[c][ci] xs[ci]) on gated channels
      for (int i = tid; i < kMB * C; i += kRThreads) {
        const int cl = i % kMB, c = i / kMB;
        float v = 0.f;
        if (msg_on && cl < nb && c >= c_lo) {
          float agg = sbm[c] * AS[cl];
#pragma unroll 4
          for (int ci = 0; ci < C; ++ci) agg = fmaf(sWmT[ci * C + c], XSt[ci * kMBP + cl], agg);
          v = tanhf(agg) * gain_m;
        }
        MSGt[c * kMBP + cl] = v;
      }
      // 2d: layer 2 with the hidden dimension split over the warps: lane -> (4 cells x 4 channels) tile
      {
        constexpr int CQ = C / 4;               // channel quads
        const int ntile = (kMB / 4) * CQ;       // tiles of 4 cells x 4 channels
        const int jper = (hid + kRWarps - 1) / kRWarps;
        const int j0 = warp * jper, j1 = min(hid, j0 + jper);
        for (int tile = lane; tile < ntile; tile += 32) {
          const int cgp = tile % (kMB / 4), cq = tile / (kMB / 4);
          float acc[4][4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[m][cc] = 0.f;
          if (4 * cgp < nb) {
            for (int j = j0; j < j1; ++j) {
              const float4 hv = *reinterpret_cast<const float4*>(Ht + j * kMBP + 4 * cgp);
              const float4 w = *reinterpret_cast<const float4*>(sW2T + j * C + 4 * cq);
              const float hm[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                acc[m][0] = fmaf(hm[m], w.x, acc[m][0]); acc[m][1] = fmaf(hm[m], w.y, acc[m][1]);
                acc[m][2] = fmaf(hm[m], w.z, acc[m][2]); acc[m][3] = fmaf(hm[m], w.w, acc[m][3]);
              }
            }
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<float4*>(RED + ((size_t)warp * C + 4 * cq + cc) * kMB + 4 * cgp) =
                make_float4(acc[0][cc], acc[1][cc], acc[2][cc], acc[3][cc]);
        }
      }
      __syncthreads();
      for (int i = tid; i < kMB * C; i += kRThreads) {
        const int cl = i % kMB, c = i / kMB;
        if (cl < nb) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kRWarps; ++w) v += RED[((size_t)w * C + c) * kMB + cl];
          v += MSGt[c * kMBP + cl];
          Ut[(size_t)c * maxown + base + cl] = v;
          ps1 += v;
          ps2 = fmaf(v, v, ps2);
        }
      }
      __syncthreads();
    }

    // ---- GroupNorm(1,C) statistics over the whole sample: block partial -> DSMEM all-gather ----------------
    float mu = 0.f, rstd = 1.f;
    if (gn) {
      const double d1 = warp_sum((double)ps1), d2 = warp_sum((double)ps2);
      if (lane == 0) { s_wred[warp][0] = d1; s_wred[warp][1] = d2; }
      __syncthreads();
      if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < kRWarps; ++w) { t1 += s_wred[w][0]; t2 += s_wred[w][1]; }
        s_part[0] = t1; s_part[1] = t2;
      }
      cluster.sync();                                                      // barrier #1
      if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < NC; ++r) {
          const double* rp = cluster.map_shared_rank(s_part, r);
          t1 += rp[0]; t2 += rp[1];
        }
        const double n = (double)C * (double)HW;
        const double m = t1 / n;
        double var = t2 / n - m * m;
        if (var < 0.0) var = 0.0;
        s_stat[0] = (float)m;
        s_stat[1] = (float)(1.0 / sqrt(var + (double)a.gn_eps));
        if (rank == 0 && R.stats_hist) {
          R.stats_hist[((size_t)t * a.B + b) * 2] = s_stat[0];
          R.stats_hist[((size_t)t * a.B + b) * 2 + 1] = s_stat[1];
        }
      }
      __syncthreads();
      mu = s_stat[0]; rstd = s_stat[1];
    }
    if (tid < C) {
      float sc = 1.f, bi = 0.f;
      if (gn) { sc = rstd * packed[P.gamma + tid]; bi = packed[P.beta + tid] - mu * sc; }
      s_sc[tid] = sc; s_bi[tid] = bi;
      s_idle[tid] = tanhf(bi) * a.update_gain;
    }
    __syncthreads();

    // ---- P3: bounded update of own cells -> x_{t+1} (alpha goes to alpha_tmp, pre-gate) ------------------------
    for (int i = tid; i < nown * C; i += kRThreads) {
      const int c = i / nown, l = i - c * nown;
      const int cell = global_cell(l);
      if (cell >= HW) continue;
      const int slot = s_slot[l];
      const float xin = ldc(xc + (size_t)c * HW + cell);
      const float d = slot >= 0 ? tanhf(fmaf(Ut[(size_t)c * maxown + slot], s_sc[c], s_bi[c])) * a.update_gain : s_idle[c];
      if (c == 3) alpha_tmp[cell] = xin + d;
      else xn[(size_t)c * HW + cell] = xin + d;
    }
    __syncthreads();
    cluster.sync();                                                        // barrier #2
    // ---- P4: post-alive gate on the updated alpha ----------------------------------------------------------------
    for (int l = tid; l < nown; l += kRThreads) {
      const int cell = global_cell(l);
      if (cell >= HW) continue;
      const int y = cell / W, x = cell - y * W;
      const bool post = alive_at_c(alpha_tmp, y, x, H, W, a.alpha_thr);
      xn[(size_t)3 * HW + cell] = post ? ldc(alpha_tmp + cell) : 0.f;
    }
    __syncthreads();
    cluster.sync();                                                        // barrier #3
  }

  // x_T -> xT (own cells)
  {
    const float* xl = x_ptr(R.T);
    float* xo = R.xT + sample_off;
    for (int i = tid; i < nown * C; i += kRThreads) {
      const int c = i / nown, l = i - c * nown;
      const int cell = global_cell(l);
      if (cell < HW) xo[(size_t)c * HW + cell] = ldc(xl + (size_t)c * HW + cell);
    }
  }
}

// ------------------------------------------------------------------------------------------------
static size_t resident_smem_bytes(int C, int hid, bool graph, int maxown) {
  size_t f = (size_t)3 * C * hid + pad4(hid) + (size_t)hid * C;
  if (graph) f += C * C + C;
  f += (size_t)3 * C * kMBP + (size_t)C * kMBP + kMBP + (size_t)C * kMBP + (size_t)hid * kMBP;
  f += (size_t)kRWarps * C * kMB;
  f += (size_t)C * maxown;
  size_t bytes = f * sizeof(float);
  bytes += (size_t)kMB * GNCA_MAX_K * sizeof(int) + 2 * (size_t)maxown * sizeof(int);
  return bytes;
}

template <int C>
static int launch_resident(const gnca_model& m, const Packed& P, const float* packed, ResidentArgs& R, int B,
                           cudaStream_t st) {
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const int HW = R.s.H * R.s.W;
  const int nseg = (HW + 31) / 32;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GNCA_CHECK_CUDA(cudaFuncSetAttribute(k_resident_fwd<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int cands[5] = {16, 8, 4, 2, 1};
  const char* env_nc = getenv("GNCA_RESIDENT_NC");      // development override of the cluster size
  const bool debug = getenv("GNCA_DEBUG") != nullptr;
  for (int ci = 0; ci < 5; ++ci) {
    const int NC = cands[ci];
    if (env_nc && atoi(env_nc) != NC) continue;
    if (NC > nseg) continue;
    if ((long long)B * NC > sms && NC > 1) continue;       // keep the whole batch co-resident when possible
    const int maxown = ((nseg + NC - 1) / NC) * 32;
    const size_t smem = resident_smem_bytes(C, m.hidden, graph, maxown);
    if (smem > 226 * 1024) continue;
    GNCA_CHECK_CUDA(cudaFuncSetAttribute(k_resident_fwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(B * NC);
    cfg.blockDim = dim3(kRThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_resident_fwd<C>, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      continue;
    }
    if (debug)
      fprintf(stderr, "[gnca] resident fwd: B=%d NC=%d grid=%d smem=%zu maxActiveClusters=%d\n", B, NC, B * NC, smem,
              nclusters);
    R.NC = NC; R.maxown = maxown; R.nseg = nseg;
    prof_begin(PROF_RESIDENT_FWD, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_resident_fwd<C>, R, P, m.hidden, packed);
    prof_end(PROF_RESIDENT_FWD, st);
    if (e != cudaSuccess) return (int)e;
    GNCA_LAUNCH_CHECK();
    return 0;
  }
  return GNCA_ERR_UNSUPPORTED;
}

// entry used by gnca_rollout_fwd (impl 2 / auto)
int run_resident_fwd(const gnca_model& m, const Packed& P, const float* packed, int B, int H, int W,
                     const gnca_schedule& sched, const float* x0, float* xT, float* hist, float* stats_hist,
                     float* ping, float* pong, float* alpha_tmp, cudaStream_t st) {
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  if (graph && !(m.flags & GNCA_F_TORUS) && sched.k > 0) return GNCA_ERR_UNSUPPORTED;   // zero-pad: streaming path
  ResidentArgs R{};
  fill_step_args(R.s, m, B, H, W);
  R.s.k = graph ? sched.k : 0;
  R.s.fire_rate_dev = sched.fire_rate;
  R.s.message_gain_dev = sched.message_gain;
  R.s.offsets_dev = sched.offsets;
  R.s.steps = sched.steps;
  R.s.philox_seed = sched.philox_seed;
  R.s.philox_offset = sched.philox_offset;
  R.fire_u_base = sched.fire_u;
  R.T = sched.T;
  R.x0 = x0; R.xT = xT; R.hist = hist; R.stats_hist = stats_hist;
  R.damage = sched.damage; R.damage_step = sched.damage_step;
  R.ping = ping; R.pong = pong; R.alpha_tmp = alpha_tmp;
  switch (m.C) {
    case 4: return launch_resident<4>(m, P, packed, R, B, st);
    case 8: return launch_resident<8>(m, P, packed, R, B, st);
    case 16: return launch_resident<16>(m, P, packed, R, B, st);
    case 32: return launch_resident<32>(m, P, packed, R, B, st);
  }
  return GNCA_ERR_UNSUPPORTED;
}

}  // namespace gnca
