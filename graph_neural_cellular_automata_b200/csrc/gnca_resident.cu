// Cluster-resident forward rollout: ONE launch runs all T steps of a batch, state in shared memory.
//
// A sample is owned by a thread-block cluster of NC CTAs (NC in {8,4,2,1}: largest that keeps the batch
// co-resident and divides H into bands of at least HALO rows).  CTA r holds, in shared memory and for the whole
// rollout, the band of rows [r*own, (r+1)*own) of every channel plus HALO = radius+1 rows of each neighbouring
// band (torus-wrapped, so the mid-range message reads its senders in place).  HBM is touched at the rollout
// boundaries (x_0 in, x_T out) and, only when BPTT history is requested, by one store of x_t per step.
//
// Per step (reference semantics: ncagraph.py:106-168):
//   P1  alive & fire test on own cells, deterministic compaction of the active cells
//   P2  perception + MLP (+ graph message) on active cells in staged batches (FFMA register tiles, weights in smem)
//   --  per-sample GroupNorm statistics: block partial -> DSMEM all-gather           [cluster barrier 1]
//   P3  bounded update in place; idle update of the halo copies; pre-gate alpha rows pushed to the neighbours
//                                                                                    [cluster barrier 2]
//   P4  post-alive gate; gated alpha rows and the active cells' channels pushed into the neighbours' halos
//                                                                                    [cluster barrier 3]
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "gnca_common.cuh"
#include "gnca_internal.h"

namespace cg = cooperative_groups;

namespace gnca {

constexpr int kRThreads = 512;
constexpr int kRWarps = kRThreads / 32;
constexpr int kRSlices = 8;      // hidden-dimension slices of layer 2 (partials reduced through smem)

struct ResidentArgs {
  StepArgs s;                 // model scalars + schedule pointers (t / fire_u are set per step in-kernel)
  int T, NC, own_rows, halo, MB, plane_stride;
  double inv_n;               // 1 / (C*H*W)
  const float* fire_u_base;   // [T][B][H][W] or null
  const float* x0;
  float* xT;
  float* hist;                // [T+1][B][C][HW] or null
  float* stats_hist;          // [T][B][2] or null
  const float* damage;        // [B][C][HW] or null
  int damage_step;
  unsigned long long* dbg;    // optional phase-cycle counters (GNCA_PHASE_TIMING=<cta index>), else null
  int dbg_cta;
  int dbg_skip;               // development: bitmask of phases to skip (timing experiments only; results wrong)
};

// phase counters accumulate in shared memory (a global RMW per mark would sit on the critical path)
#define GNCA_PHASE_MARK(idx)                                                        \
  do {                                                                              \
    if (R.dbg && tid == 0) {                                                        \
      const long long _n = clock64();                                               \
      s_dbg[idx] += (unsigned long long)(_n - t_prev);                              \
      t_prev = _n;                                                                  \
    }                                                                               \
  } while (0)

// cluster barrier with cluster-scope release/acquire (cooperative_groups' cluster.sync() fences at GPU scope)
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Out-of-line helpers: the per-step code of the resident kernel has to stay small enough for the instruction
// caches (every phase is executed once per step by few warps, so cold-code fetch latency is on the critical path).
__device__ __noinline__ float tanh_ool(float v) { return tanhf(v); }

__device__ __noinline__ double warp_sum_ool(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// alive(maxpool3x3 > thr) around local row lr / column x of an alpha-like plane with row stride W; rows outside
// the image (by GLOBAL row gy of the centre) count as -inf exactly like F.max_pool2d (nca.py:61).
__device__ __noinline__ bool alive_local_ool(const float* plane, int lr, int x, int gy, int H, int W, float thr) {
  float m = -INFINITY;
#pragma unroll
  for (int i = -1; i <= 1; ++i) {
    const int gyy = gy + i;
    if (gyy < 0 || gyy >= H) continue;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
      const int xx = x + j;
      if (xx < 0 || xx >= W) continue;
      m = fmaxf(m, plane[(lr + i) * W + xx]);
    }
  }
  return m > thr;
}

template <int C>
__global__ void __launch_bounds__(kRThreads, 1) k_resident_fwd(ResidentArgs R, Packed P, int hid,
                                                               const float* __restrict__ packed) {
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int C3 = 3 * C;
  constexpr int CQ = C / 4;                                     // channel quads (power of two)
  constexpr int LCQ = (CQ == 1) ? 0 : (CQ == 2) ? 1 : (CQ == 4) ? 2 : 3;
  StepArgs a = R.s;
  const int NC = R.NC;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / NC;
  const int H = a.H, W = a.W, HW = H * W;
  const bool graph = (a.flags & GNCA_F_GRAPH) != 0;
  const bool gn = (a.flags & GNCA_F_GROUPNORM) != 0;
  const bool a2a = (a.flags & GNCA_F_ALIVE_TO_ALIVE) != 0;
  const int c_lo = ((a.flags & GNCA_F_HIDDEN_ONLY) && C >= 4) ? 4 : 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int own = R.own_rows, HALO = R.halo, RS = own + 2 * HALO;
  const int r0 = rank * own;                 // first global row of the band
  const int nown = own * W;                  // own cells
  const int PL = R.plane_stride;             // padded plane stride (== 1 mod 32: bank-conflict free across channels)
  const int MB = R.MB, MBP = MB + 4;
  const int JG = hid >> 7;                   // groups of 128 hidden units (hid % 128 == 0)
  const int kk = a.k > 0 ? a.k : 1;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sW1T = reinterpret_cast<float*>(smem_raw);     // [3C][hid], hidden index permuted (see below)
  float* sb1 = sW1T + C3 * hid;                         // [hid] same permutation
  float* sW2T = sb1 + hid;                              // [hid][C]
  float* sWmT = sW2T + hid * C;
  float* sbm = sWmT + (graph ? C * C : 0);
  float* sX = sbm + (graph ? C : 0);                    // [C][RS][W]   resident state: band + halos
  float* sAt = sX + pad4(C * PL);                       // [own+2][W]   pre-gate updated alpha (1-row halo)
  float* Yt = sAt + pad4((own + 2) * W);                // [3C][MBP]
  float* XSt = Yt + C3 * MBP;                           // [C][MBP]
  float* AS = XSt + C * MBP;                            // [MBP]
  float* MSGt = AS + MBP;                               // [C][MBP]
  float* Ht = MSGt + C * MBP;                           // [hid][MBP]
  float* RED = Ht + hid * MBP;                          // [kRSlices][C][MB]
  float* Ut = RED + kRSlices * C * MB;                  // [C][nown] masked pre-norm update of active cells
  int* s_q = reinterpret_cast<int*>(Ut + (size_t)C * nown);   // [MB][k] sender (lr*W+x) or -1
  int* s_actlist = s_q + MB * kk;                       // [nown] own-cell index of active cells (compacted)
  int* s_slot = s_actlist + nown;                       // [nown] slot in the active list or -1
  unsigned char* s_oy = reinterpret_cast<unsigned char*>(s_slot + nown);   // [nown] own row of an own cell
  unsigned char* s_ox = s_oy + nown;                                        // [nown] column
  float* s_fr = reinterpret_cast<float*>(s_ox + ((nown + 15) & ~15));      // [T] fire rate per step
  float* s_gain = s_fr + R.T;                                               // [T] message gain per step
  signed char* s_off = reinterpret_cast<signed char*>(s_gain + R.T);        // [T][k][2] offsets
  __shared__ double s_parts[8][2];           // (sum u, sum u^2) of every CTA of the cluster, pushed by its owner
  __shared__ double s_wred[kRWarps][2];
  __shared__ float s_sc[C], s_bi[C], s_idle[C], s_gam[C], s_bet[C], s_stat[2];
  __shared__ int s_wcount[kRWarps], s_wbase[kRWarps + 1];
  __shared__ unsigned long long s_dbg[16];
  if (tid < 16) s_dbg[tid] = 0;

  // ---- weights -> smem.  Hidden units are permuted so that lane l of a warp owns units {g*128 + l + 32*jj}:
  //      its 4 weights of one input k are one float4, and the warp's stores of h are bank-conflict free.
#pragma unroll 1
  for (int i = tid; i < C3 * hid; i += kRThreads) {
    const int k = i / hid, jp = i - k * hid;                   // jp = g*128 + l*4 + jj
    const int g = jp >> 7, l = (jp & 127) >> 2, jj = jp & 3;
    sW1T[i] = packed[P.w1t + k * hid + (g * 128 + l + 32 * jj)];
  }
#pragma unroll 1
  for (int jp = tid; jp < hid; jp += kRThreads) {
    const int g = jp >> 7, l = (jp & 127) >> 2, jj = jp & 3;
    sb1[jp] = packed[P.b1 + g * 128 + l + 32 * jj];
  }
  block_copy(sW2T, packed + P.w2t, hid * C);
  if (graph) { block_copy(sWmT, packed + P.wmt, C * C); block_copy(sbm, packed + P.bm, C); }
  if (tid < C) { s_gam[tid] = packed[P.gamma + tid]; s_bet[tid] = packed[P.beta + tid]; }
  // the whole schedule comes on chip once: no global-memory latency inside the step loop
#pragma unroll 1
  for (int i = tid; i < R.T; i += kRThreads) {
    s_fr[i] = a.fire_rate_dev[i];
    s_gain[i] = graph ? a.message_gain_dev[i] : 0.f;
  }
#pragma unroll 1
  for (int i = tid; i < R.T * a.k * 2; i += kRThreads) s_off[i] = a.offsets_dev[i];
#pragma unroll 1
  for (int oc = tid; oc < nown; oc += kRThreads) {
    const int orow = oc / W;
    s_oy[oc] = (unsigned char)orow;
    s_ox[oc] = (unsigned char)(oc - orow * W);
  }

  const size_t sample_off = (size_t)b * C * HW;
  auto wrap_row = [&](int lr) { int gr = (r0 - HALO + lr) % H; return gr < 0 ? gr + H : gr; };
  // ---- x_0 -> smem (band + wrapped halos) ---------------------------------------------------------------
#pragma unroll 1
  for (int row = warp; row < C * RS; row += kRWarps) {
    const int c = row / RS, lr = row - c * RS;
    const float* src = R.x0 + sample_off + (size_t)c * HW + wrap_row(lr) * W;
#pragma unroll 1
    for (int x = lane; x < W; x += 32) sX[(size_t)c * PL + lr * W + x] = src[x];
  }
  float* sA = sX + (size_t)3 * PL;                      // alpha plane
  // neighbour CTAs (torus order) and their views of our pushes
  const int prev = (rank + NC - 1) % NC, next = (rank + 1) % NC;
  float* pX = cluster.map_shared_rank(sX, prev);
  float* nX = cluster.map_shared_rank(sX, next);
  float* pAt = cluster.map_shared_rank(sAt, prev);
  float* nAt = cluster.map_shared_rank(sAt, next);
  __syncthreads();
  cluster_barrier();

  const int my_steps = a.steps ? min(a.steps[b], R.T) : R.T;
  long long t_prev = clock64();

  auto store_own = [&](float* dst) {   // own rows of every channel -> global [C][H][W] slice of this sample
#pragma unroll 1
    for (int c = 0; c < C; ++c)
#pragma unroll 1
      for (int oc = tid; oc < nown; oc += kRThreads)
        dst[(size_t)c * HW + r0 * W + oc] = sX[(size_t)c * PL + HALO * W + oc];
  };
  auto alive_local = [&](const float* plane, int lr, int x, int gy, float thr) {
    return alive_local_ool(plane, lr, x, gy, H, W, thr);
  };

  for (int t = 0; t < R.T; ++t) {
    if (R.damage && t == R.damage_step) {      // multiplicative damage on every copy we hold (own + halos)
      const float* D = R.damage + sample_off;
#pragma unroll 1
      for (int row = warp; row < C * RS; row += kRWarps) {
        const int c = row / RS, lr = row - c * RS;
        const float* src = D + (size_t)c * HW + wrap_row(lr) * W;
#pragma unroll 1
        for (int x = lane; x < W; x += 32) sX[(size_t)c * PL + lr * W + x] *= src[x];
      }
      __syncthreads();
    }
    if (R.hist) store_own(R.hist + (size_t)t * a.B * C * HW + sample_off);
    if (t >= my_steps) continue;               // frozen sample (whole cluster agrees): state passes through
    a.t = t;
    a.fire_u = R.fire_u_base ? R.fire_u_base + (size_t)t * a.B * HW : nullptr;
    const float fr = s_fr[t];
    const float gain_m = s_gain[t];
    const bool msg_on = graph && gain_m != 0.f && a.k > 0;

    // ---- P1: alive & fire on own cells, deterministic compaction ------------------------------------------
    {
      int cnt = 0;
      const int per = ((nown + kRThreads - 1) / kRThreads) * 32;     // own cells per warp (multiple of 32)
      const int lo = warp * per;
      for (int base = lo; base < lo + per; base += 32) {
        const int oc = base + lane;
        bool act = false;
        if (oc < nown) {
          const int orow = s_oy[oc], x = s_ox[oc];
          act = alive_local(sA, HALO + orow, x, r0 + orow, a.alpha_thr) && fires(a, fr, b, (r0 + orow) * W + x);
          s_slot[oc] = act ? 1 : -1;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, act));
      }
      if (lane == 0) s_wcount[warp] = cnt;
      __syncthreads();
      if (warp == 0) {                        // exclusive scan of the per-warp counts
        const int v = lane < kRWarps ? s_wcount[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < kRWarps; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        if (lane < kRWarps) s_wbase[lane] = incl - v;
        if (lane == kRWarps - 1) s_wbase[kRWarps] = incl;
      }
      __syncthreads();
      int pos = s_wbase[warp];
      for (int base = lo; base < lo + per; base += 32) {
        const int oc = base + lane;
        const bool act = oc < nown && s_slot[oc] > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (act) {
          const int slot = pos + __popc(bal & ((1u << lane) - 1u));
          s_actlist[slot] = oc;
          s_slot[oc] = slot;
        }
        pos += __popc(bal);
      }
      __syncthreads();
    }
    const int nact = s_wbase[kRWarps];
    GNCA_PHASE_MARK(0);
    if (R.dbg && tid == 0) s_dbg[8] += nact;

    // ---- P2: perception + MLP + message on the active cells, in batches of MB ----------------------------
    float ps1 = 0.f, ps2 = 0.f;
    for (int base = 0; base < ((R.dbg_skip & 1) ? 0 : nact); base += MB) {
      const int nb = min(MB, nact - base);
      const int G = nb > 32 ? 8 : 4;                    // cells per warp tile in layer 1
      const int nbp = ((nb + G - 1) / G) * G;           // staged cells (zero padded to the tile)
      // 2a/2b: one half-warp per cell: perception of its C channels, sender table, gathered sender state
      {
        const int hw = tid >> 4, l16 = tid & 15;
        for (int cl = hw; cl < nbp; cl += kRThreads / 16) {      // nbp is even: both halves of a warp iterate alike
          const bool valid = cl < nb;
          int orow = 0, x = 0;
          if (valid) { const int oc = s_actlist[base + cl]; orow = s_oy[oc]; x = s_ox[oc]; }
          const int gy = r0 + orow;
          if (msg_on) {
            if (valid) {
              for (int oi = l16; oi < a.k; oi += 16) {
                const int dy = s_off[(t * a.k + oi) * 2], dx = s_off[(t * a.k + oi) * 2 + 1];
                int gq = (gy - dy) % H; gq = gq < 0 ? gq + H : gq;       // sender's global row (torus)
                int qx = (x - dx) % W; qx = qx < 0 ? qx + W : qx;
                const int lq = HALO + orow - dy;                          // its local row (halo holds the wrap)
                int q = lq * W + qx;
                if (a2a && !alive_local(sA, lq, qx, gq, a.graph_alpha_thr)) q = -1;
                s_q[cl * kk + oi] = q;
              }
            }
            __syncwarp();
          }
          const bool up = gy > 0, dn = gy < H - 1, lf = x > 0, rt = x < W - 1;
          for (int c = l16; c < C; c += 16) {
            float vid = 0.f, vsx = 0.f, vsy = 0.f, xs = 0.f;
            if (valid) {
              const float* p = sX + (size_t)c * PL + (HALO + orow) * W + x;
              const float a00 = (up && lf) ? p[-W - 1] : 0.f, a01 = up ? p[-W] : 0.f, a02 = (up && rt) ? p[-W + 1] : 0.f;
              const float a10 = lf ? p[-1] : 0.f, a12 = rt ? p[1] : 0.f;
              const float a20 = (dn && lf) ? p[W - 1] : 0.f, a21 = dn ? p[W] : 0.f, a22 = (dn && rt) ? p[W + 1] : 0.f;
              vid = p[0];
              vsx = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
              vsy = (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
              if (msg_on) {
                const float wuni = 1.0f / (float)a.k;
                for (int oi = 0; oi < a.k; ++oi) {
                  const int q = s_q[cl * kk + oi];
                  if (q >= 0) xs = fmaf(wuni, sX[(size_t)c * PL + q], xs);
                }
              }
            }
            Yt[c * MBP + cl] = vid;
            Yt[(C + c) * MBP + cl] = vsx;
            Yt[(2 * C + c) * MBP + cl] = vsy;
            XSt[c * MBP + cl] = xs;
          }
          if (l16 == 0) {
            float as = 0.f;
            if (valid && msg_on) {
              const float wuni = 1.0f / (float)a.k;
              for (int oi = 0; oi < a.k; ++oi) if (s_q[cl * kk + oi] >= 0) as += wuni;
            }
            AS[cl] = as;
          }
        }
      }
      __syncthreads();
      GNCA_PHASE_MARK(10);
      // message projection: MSGt[c][cl] = gain * tanh(bm[c]*as + sum_ci Wm[c][ci] xs[ci]) on gated channels
      for (int i = tid; i < nb * CQ; i += kRThreads) {
        const int cq = i & (CQ - 1), cl = i >> LCQ;
        float agg[4] = {0.f, 0.f, 0.f, 0.f};
        if (msg_on && 4 * cq + 3 >= c_lo) {
          const float as = AS[cl];
          const float4 bmv = *reinterpret_cast<const float4*>(sbm + 4 * cq);
          agg[0] = bmv.x * as; agg[1] = bmv.y * as; agg[2] = bmv.z * as; agg[3] = bmv.w * as;
#pragma unroll 4
          for (int ci = 0; ci < C; ++ci) {
            const float xv = XSt[ci * MBP + cl];
            const float4 w = *reinterpret_cast<const float4*>(sWmT + ci * C + 4 * cq);
            agg[0] = fmaf(w.x, xv, agg[0]); agg[1] = fmaf(w.y, xv, agg[1]);
            agg[2] = fmaf(w.z, xv, agg[2]); agg[3] = fmaf(w.w, xv, agg[3]);
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) agg[cc] = (4 * cq + cc >= c_lo) ? tanh_ool(agg[cc]) * gain_m : 0.f;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) MSGt[(4 * cq + cc) * MBP + cl] = agg[cc];
      }
      // 2c: layer 1.  warp = tile of G cells, lane = 4 (permuted) hidden units per 128-group
      auto layer1 = [&](auto gtag) {
        constexpr int GG = decltype(gtag)::value;
        const int ngroups = nbp / GG;
        for (int item = warp; item < ngroups * JG; item += kRWarps) {
          const int cgp = item % ngroups, g = item / ngroups;
          const int jp = g * 128 + lane * 4;            // permuted index of this lane's 4 units
          float acc[GG][4];
          const float4 bb = *reinterpret_cast<const float4*>(sb1 + jp);
#pragma unroll
          for (int m = 0; m < GG; ++m) { acc[m][0] = bb.x; acc[m][1] = bb.y; acc[m][2] = bb.z; acc[m][3] = bb.w; }
#pragma unroll 4
          for (int k = 0; k < C3; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(sW1T + k * hid + jp);
            float ym[GG];
#pragma unroll
            for (int m4 = 0; m4 < GG / 4; ++m4) {
              const float4 yv = *reinterpret_cast<const float4*>(Yt + k * MBP + GG * cgp + 4 * m4);
              ym[4 * m4] = yv.x; ym[4 * m4 + 1] = yv.y; ym[4 * m4 + 2] = yv.z; ym[4 * m4 + 3] = yv.w;
            }
#pragma unroll
            for (int m = 0; m < GG; ++m) {
              acc[m][0] = fmaf(ym[m], w.x, acc[m][0]); acc[m][1] = fmaf(ym[m], w.y, acc[m][1]);
              acc[m][2] = fmaf(ym[m], w.z, acc[m][2]); acc[m][3] = fmaf(ym[m], w.w, acc[m][3]);
            }
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = g * 128 + lane + 32 * jj;     // true hidden index (row of Ht / W2T)
#pragma unroll
            for (int m4 = 0; m4 < GG / 4; ++m4) {
              float4 hv;
              hv.x = fmaxf(acc[4 * m4][jj], 0.f); hv.y = fmaxf(acc[4 * m4 + 1][jj], 0.f);
              hv.z = fmaxf(acc[4 * m4 + 2][jj], 0.f); hv.w = fmaxf(acc[4 * m4 + 3][jj], 0.f);
              *reinterpret_cast<float4*>(Ht + j * MBP + GG * cgp + 4 * m4) = hv;
            }
          }
        }
      };
      if (G == 8) layer1(std::integral_constant<int, 8>{}); else layer1(std::integral_constant<int, 4>{});
      __syncthreads();
      GNCA_PHASE_MARK(11);
      // 2d: layer 2: the hidden dimension is split in kRSlices slices (warp % kRSlices); the warps of one slice
      //     share the (4 cells x 4 channels) tiles
      {
        const int ncg = (nb + 3) >> 2;
        const int ntile = ncg * CQ;
        const int ks = warp % kRSlices, tw = warp / kRSlices;
        const int jper = hid / kRSlices;
        const int j0 = ks * jper, j1 = j0 + jper;
        for (int tile = tw * 32 + lane; tile < ntile; tile += 32 * (kRWarps / kRSlices)) {
          const int cq = tile & (CQ - 1), cgp = tile >> LCQ;
          float acc[4][4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[m][cc] = 0.f;
#pragma unroll 4
          for (int j = j0; j < j1; ++j) {
            const float4 hv = *reinterpret_cast<const float4*>(Ht + j * MBP + 4 * cgp);
            const float4 w = *reinterpret_cast<const float4*>(sW2T + j * C + 4 * cq);
            const float hm[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              acc[m][0] = fmaf(hm[m], w.x, acc[m][0]); acc[m][1] = fmaf(hm[m], w.y, acc[m][1]);
              acc[m][2] = fmaf(hm[m], w.z, acc[m][2]); acc[m][3] = fmaf(hm[m], w.w, acc[m][3]);
            }
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<float4*>(RED + ((size_t)ks * C + 4 * cq + cc) * MB + 4 * cgp) =
                make_float4(acc[0][cc], acc[1][cc], acc[2][cc], acc[3][cc]);
        }
      }
      __syncthreads();
      GNCA_PHASE_MARK(12);
      {
        const int nstrip = (nb + 31) >> 5;
        for (int item = warp; item < nstrip * C; item += kRWarps) {
          const int c = item & (C - 1), cl = (item / C) * 32 + lane;
          if (cl < nb) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kRSlices; ++w) v += RED[((size_t)w * C + c) * MB + cl];
            v += MSGt[c * MBP + cl];
            Ut[(size_t)c * nown + base + cl] = v;
            ps1 += v;
            ps2 = fmaf(v, v, ps2);
          }
        }
      }
      __syncthreads();
    }
    GNCA_PHASE_MARK(1);

    // ---- GroupNorm(1,C) statistics over the whole sample: block partial -> DSMEM all-gather ----------------
    float mu = 0.f, rstd = 1.f;
    if (gn && !(R.dbg_skip & 2)) {
      const float f1 = warp_sum(ps1), f2 = warp_sum(ps2);
      if (lane == 0) { s_wred[warp][0] = (double)f1; s_wred[warp][1] = (double)f2; }
      __syncthreads();
      if (warp == 0) {
        double t1 = lane < kRWarps ? s_wred[lane][0] : 0.0, t2 = lane < kRWarps ? s_wred[lane][1] : 0.0;
        t1 = warp_sum_ool(t1); t2 = warp_sum_ool(t2);
        if (lane < NC) {                       // push our partial into slot `rank` of every CTA (incl. ourselves)
          double* dst = cluster.map_shared_rank(&s_parts[0][0], lane);
          dst[rank * 2] = t1; dst[rank * 2 + 1] = t2;
        }
      }
      GNCA_PHASE_MARK(2);
      cluster_barrier();                                                      // barrier 1
      GNCA_PHASE_MARK(3);
      if (warp == 0) {
        double t1 = 0.0, t2 = 0.0;
        if (lane < NC) { t1 = s_parts[lane][0]; t2 = s_parts[lane][1]; }
        // same butterfly in every CTA of the cluster => bit-identical statistics cluster-wide
        GNCA_PHASE_MARK(13);
        const double a1 = warp_sum_ool(t1), a2 = warp_sum_ool(t2);
        GNCA_PHASE_MARK(14);
        if (lane == 0) {
          const double invn = R.inv_n;
          const double m = a1 * invn;
          double var = fma(a2, invn, -m * m);
          if (var < 0.0) var = 0.0;
          s_stat[0] = (float)m;
          s_stat[1] = 1.0f / sqrtf((float)var + a.gn_eps);
          if (rank == 0 && R.stats_hist) {
            R.stats_hist[((size_t)t * a.B + b) * 2] = s_stat[0];
            R.stats_hist[((size_t)t * a.B + b) * 2 + 1] = s_stat[1];
          }
        }
        __syncwarp();
        GNCA_PHASE_MARK(15);
        if (lane < C) {                       // C <= 32: the same warp finishes the per-channel affine
          const float m_ = s_stat[0], r_ = s_stat[1];
          const float sc = r_ * s_gam[lane];
          const float bi = s_bet[lane] - m_ * sc;
          s_sc[lane] = sc; s_bi[lane] = bi;
          s_idle[lane] = tanh_ool(bi) * a.update_gain;
        }
      }
    } else if (tid < C) {
      s_sc[tid] = 1.f; s_bi[tid] = 0.f; s_idle[tid] = 0.f;
    }
    __syncthreads();
    (void)mu; (void)rstd;
    GNCA_PHASE_MARK(4);

    // ---- P3: bounded update in place (own cells); idle update of the halo copies; pre-gate alpha rows -------
    for (int item = warp; item < ((R.dbg_skip & 4) ? 0 : ((nown + 31) >> 5) * CQ); item += kRWarps) {
      const int cq = item & (CQ - 1), oc = (item >> LCQ) * 32 + lane;
      if (oc >= nown) continue;
      const int slot = s_slot[oc];
      float* px = sX + (size_t)(4 * cq) * PL + HALO * W + oc;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * cq + cc;
        const float d = slot >= 0 ? tanh_ool(fmaf(Ut[(size_t)c * nown + slot], s_sc[c], s_bi[c])) * a.update_gain : s_idle[c];
        const float v = px[(size_t)cc * PL] + d;
        if (c == 3) {
          sAt[W + oc] = v;                                   // rows 1..own of sAt
          const int orow = s_oy[oc];
          if (orow == 0) pAt[(own + 1) * W + oc] = v;        // our top row = prev's bottom halo row
          if (orow == own - 1) nAt[oc - (own - 1) * W] = v;  // our bottom row = next's top halo row
        } else {
          px[(size_t)cc * PL] = v;
        }
      }
    }
    for (int item = warp; item < ((2 * HALO * W + 31) >> 5) * CQ; item += kRWarps) {   // halo cells, channels != 3
      const int cq = item & (CQ - 1), hc = (item >> LCQ) * 32 + lane;
      if (hc >= 2 * HALO * W) continue;
      const int off = hc < HALO * W ? hc : own * W + hc;                    // top halo | bottom halo
      float* px = sX + (size_t)(4 * cq) * PL + off;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        if (4 * cq + cc != 3) px[(size_t)cc * PL] += s_idle[4 * cq + cc];
    }
    __syncthreads();
    GNCA_PHASE_MARK(5);
    if (!(R.dbg_skip & 16)) cluster_barrier();                                // barrier 2
    GNCA_PHASE_MARK(6);
    // ---- P4: post-alive gate on own cells; push gated alpha rows + active cells' channels to the halos -------
    for (int oc = tid; oc < ((R.dbg_skip & 8) ? 0 : nown); oc += kRThreads) {
      const int orow = s_oy[oc], x = s_ox[oc];
      const bool post = alive_local(sAt, 1 + orow, x, r0 + orow, a.alpha_thr);
      const float v = post ? sAt[W + oc] : 0.f;
      sA[HALO * W + oc] = v;
      if (orow < HALO) pX[(size_t)3 * PL + (HALO + own) * W + oc] = v;                 // prev's bottom halo
      if (orow >= own - HALO) nX[(size_t)3 * PL + oc - (own - HALO) * W] = v;           // next's top halo
    }
    for (int item = warp; item < ((nact + 31) >> 5) * CQ; item += kRWarps) {
      const int cq = item & (CQ - 1), sl = (item >> LCQ) * 32 + lane;
      if (sl >= nact) continue;
      const int oc = s_actlist[sl];
      const int orow = s_oy[oc];
      const bool to_prev = orow < HALO, to_next = orow >= own - HALO;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * cq + cc;
        if (c == 3) continue;
        const float v = sX[(size_t)c * PL + HALO * W + oc];
        if (to_prev) pX[(size_t)c * PL + (HALO + own) * W + oc] = v;
        if (to_next) nX[(size_t)c * PL + oc - (own - HALO) * W] = v;
      }
    }
    __syncthreads();
    GNCA_PHASE_MARK(7);
    if (!(R.dbg_skip & 16)) cluster_barrier();                                // barrier 3
    GNCA_PHASE_MARK(9);
  }

  if (R.hist) store_own(R.hist + (size_t)R.T * a.B * C * HW + sample_off);
  store_own(R.xT + sample_off);
  if (R.dbg && blockIdx.x == R.dbg_cta && tid < 16) R.dbg[tid] = s_dbg[tid];
  cluster_barrier();     // nobody exits while a neighbour may still address its shared memory
}

// ------------------------------------------------------------------------------------------------
static int plane_stride_of(int own, int halo, int W) {
  const int pl = (own + 2 * halo) * W;
  return pl + ((33 - pl % 32) % 32);        // == 1 (mod 32)
}

static size_t resident_smem_bytes(int C, int hid, bool graph, int own, int halo, int W, int MB, int k, int T) {
  const int MBP = MB + 4;
  size_t f = (size_t)3 * C * hid + hid + (size_t)hid * C;
  if (graph) f += C * C + C;
  f += (size_t)pad4(C * plane_stride_of(own, halo, W));
  f += (size_t)pad4((own + 2) * W);
  f += (size_t)3 * C * MBP + (size_t)C * MBP + MBP + (size_t)C * MBP + (size_t)hid * MBP;
  f += (size_t)kRSlices * C * MB;
  f += (size_t)C * own * W;
  size_t bytes = f * sizeof(float);
  bytes += (size_t)MB * (k > 0 ? k : 1) * sizeof(int) + 2 * (size_t)own * W * sizeof(int) + 2 * (size_t)own * W;
  bytes += 16 + (size_t)T * 8 + (size_t)T * k * 2;        // schedule: fire rate, gain, offsets
  return bytes + 64;
}

template <int C>
static int launch_resident(const gnca_model& m, const Packed& P, const float* packed, ResidentArgs& R, int B, int radius,
                           cudaStream_t st) {
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const int H = R.s.H, W = R.s.W;
  if (m.hidden % 128 != 0 || C < 4 || W > 255) return GNCA_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int halo = (graph && R.s.k > 0 ? radius : 0) + 1;
  const char* env_nc = getenv("GNCA_RESIDENT_NC");      // development override of the cluster size
  const bool debug = getenv("GNCA_DEBUG") != nullptr;
  const int cands[4] = {8, 4, 2, 1};
  for (int ci = 0; ci < 4; ++ci) {
    const int NC = cands[ci];
    if (env_nc && atoi(env_nc) != NC) continue;
    if (H % NC != 0 || H / NC < halo || H / NC > 255) continue;
    if ((long long)B * NC > sms && NC > 1 && !env_nc) continue;     // keep the whole batch co-resident when possible
    const int own = H / NC;
    int MB = 64;
    size_t smem = resident_smem_bytes(C, m.hidden, graph, own, halo, W, MB, R.s.k, R.T);
    if (smem > 226 * 1024) { MB = 32; smem = resident_smem_bytes(C, m.hidden, graph, own, halo, W, MB, R.s.k, R.T); }
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
      fprintf(stderr, "[gnca] resident fwd: B=%d NC=%d own_rows=%d halo=%d MB=%d smem=%zu maxActiveClusters=%d\n", B, NC,
              own, halo, MB, smem, nclusters);
    R.NC = NC; R.own_rows = own; R.halo = halo; R.MB = MB; R.plane_stride = plane_stride_of(own, halo, W);
    R.inv_n = 1.0 / ((double)C * (double)H * (double)W);
    if (getenv("GNCA_SKIP")) R.dbg_skip = atoi(getenv("GNCA_SKIP"));
    static unsigned long long* dbg_buf = nullptr;
    if (getenv("GNCA_PHASE_TIMING")) {
      if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
      cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), st);
      R.dbg = dbg_buf;
      R.dbg_cta = atoi(getenv("GNCA_PHASE_TIMING"));
    }
    prof_begin(PROF_RESIDENT_FWD, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_resident_fwd<C>, R, P, m.hidden, packed);
    prof_end(PROF_RESIDENT_FWD, st);
    if (e != cudaSuccess) return (int)e;
    GNCA_LAUNCH_CHECK();
    if (R.dbg) {   // development only: synchronous read-back of the phase counters of CTA 0
      unsigned long long h[16];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
      const char* names[16] = {"P1 alive+compact", "P2 tail(reduce)", "stats reduce", "barrier1", "stats finish",
                               "P3 update", "barrier2", "P4 gate+push", "(sum nact)", "barrier3", "P2 stage",
                               "P2 msg+layer1", "P2 layer2", "st:dsmem-ld", "st:warpsum", "st:math"};
      fprintf(stderr, "[gnca phase cycles, CTA%d, T=%d]", R.dbg_cta, R.T);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " %s=%llu", names[i], h[i]);
      fprintf(stderr, "\n");
    }
    return 0;
  }
  return GNCA_ERR_UNSUPPORTED;
}

// entry used by gnca_rollout_fwd (impl 2 / auto)
int run_resident_fwd(const gnca_model& m, const Packed& P, const float* packed, int B, int H, int W,
                     const gnca_schedule& sched, const float* x0, float* xT, float* hist, float* stats_hist,
                     float* ping, float* pong, float* alpha_tmp, cudaStream_t st) {
  (void)ping; (void)pong; (void)alpha_tmp;
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  if (graph && !(m.flags & GNCA_F_TORUS) && sched.k > 0) return GNCA_ERR_UNSUPPORTED;   // zero-pad: streaming path
  if (graph && sched.k > 0 && sched.max_offset <= 0) return GNCA_ERR_UNSUPPORTED;       // halo depth unknown
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
  const int radius = sched.max_offset;       // largest |dy| / |dx| in the schedule: sets the halo depth
  switch (m.C) {
    case 4: return launch_resident<4>(m, P, packed, R, B, radius, st);
    case 8: return launch_resident<8>(m, P, packed, R, B, radius, st);
    case 16: return launch_resident<16>(m, P, packed, R, B, radius, st);
    case 32: return launch_resident<32>(m, P, packed, R, B, radius, st);
  }
  return GNCA_ERR_UNSUPPORTED;
}

}  // namespace gnca
