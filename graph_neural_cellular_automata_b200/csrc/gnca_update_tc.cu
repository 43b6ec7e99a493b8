// k_update_tc: the per-cell update MLP of the streaming step on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as the balanced k_update of gnca_fwd.cu (masked pre-norm update u of the ACTIVE cells + GroupNorm
// partial sums; nca.py:75-87, ncagraph.py:128-150), for C in {16, 32} and hidden = 128.  What changes is where the
// 2*(3C*128 + 128*C) flops per cell run:
//   * a persistent CTA per SM, three warpgroups; a warpgroup owns a TILE of 128 active cells, thread t <-> cell t;
//   * the thread computes its cell's perception vector y (3C values) and graph message in registers, splits y into
//     tf32 (hi, lo) and writes both with tcgen05.st into TENSOR MEMORY lane t -- the activations never touch shared
//     memory;
//   * layer 1: D1[128 cells x 128] = y W1^T as 3 x (3C/8) tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8, A from
//     TMEM, B = the (hi, lo) split of W1 resident in shared memory in the canonical K-major layout):
//     hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM -- the large hi*hi products and the two small cross terms in
//     SEPARATE accumulators added in the epilogue: the tensor core truncates when it adds into an accumulator, and 36
//     truncations at the magnitude of the result cost 3x the error of 12 (measured: scratch/umma_probe.cu);
//   * epilogue 1 (tcgen05.ld, the thread reads ITS cell's 128 hidden units): + b1, ReLU, split, back to TMEM;
//   * layer 2: D2[128 x C] = h W2^T the same way (N = C, K = 128); epilogue 2 adds the message and stores u.
// Accuracy of the split (scratch/umma_probe.cu on a B200): max |err| 1.7e-6 against fp64 for outputs of magnitude ~1
// (fp32 FFMA chain: 2.3e-7) -- inside the 1e-5 single-step bar of the north star, tests/test_gpu_scale.py.
// One set of TMEM columns per CTA: the three warpgroups take turns on the tensor-core chain (named-barrier token,
// ~8 k cycles per tile) while the other two run their load-latency-bound perception / sender gathers.
#include "gnca_common.cuh"
#include "gnca_internal.h"
#include "gnca_tc.cuh"
#include <cstdio>

namespace gnca {

using namespace tc;

constexpr int kTcWG = 3;                    // warpgroups per CTA
constexpr int kTcThreads = kTcWG * 128;
constexpr int kTcRange = 1152;              // active cells (ranks) per unit = 3 rounds of 384; >= kTcChunk
constexpr int kTcHid = 128;
constexpr int kBarTurn = 1;                 // named barriers 1..3: the tensor-chain token of warpgroup g
constexpr int kTcNB = 4;                    // channels per batch of perception loads (36 in flight per thread)
constexpr int kBarWg = 4;                   // named barriers 4..6: warpgroup-local sync


// ---- latency-tolerant gathers of the tile kernel ------------------------------------------------------------------------
// The generic `perceive` / `gather_senders` of gnca_common.cuh compile to one basic block per channel / sender (the edge
// predicates become branches), so a warp has at most 9 loads in flight and pays one L2 round trip per channel: 30 k
// cycles per 128-cell tile each (profiles/r02_k_update_tc2.md).  The two functions below request NB channels x 9 taps
// (perception) or two senders' alive windows + all their channels (message) before the first use, without branches.

// perception of one cell, same arithmetic as perceive<C> (perception.py:21-26): the taps beyond a grid edge are read
// from a clamped (valid) address and replaced by the zero halo with a select
template <int C, int NB>
__device__ __forceinline__ void perceive_batched(const float* __restrict__ xs, int y, int x, int H, int W,
                                                 float (&out)[3 * C]) {
  static_assert(C % NB == 0, "channel batches");
  const int HW = H * W;
  const bool up = y > 0, dn = y < H - 1, lf = x > 0, rt = x < W - 1;
  const int oU = up ? -W : 0, oD = dn ? W : 0, oL = lf ? -1 : 0, oR = rt ? 1 : 0;
  const float* __restrict__ p = xs + y * W + x;
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += NB) {
    float v[NB][9];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const float* pc = p + (size_t)(c0 + n) * HW;
      v[n][0] = __ldg(pc + (oU + oL)); v[n][1] = __ldg(pc + oU); v[n][2] = __ldg(pc + (oU + oR));
      v[n][3] = __ldg(pc + oL);        v[n][4] = __ldg(pc);      v[n][5] = __ldg(pc + oR);
      v[n][6] = __ldg(pc + (oD + oL)); v[n][7] = __ldg(pc + oD); v[n][8] = __ldg(pc + (oD + oR));
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const int c = c0 + n;
      const float a00 = (up && lf) ? v[n][0] : 0.f, a01 = up ? v[n][1] : 0.f, a02 = (up && rt) ? v[n][2] : 0.f;
      const float a10 = lf ? v[n][3] : 0.f, a12 = rt ? v[n][5] : 0.f;
      const float a20 = (dn && lf) ? v[n][6] : 0.f, a21 = dn ? v[n][7] : 0.f, a22 = (dn && rt) ? v[n][8] : 0.f;
      out[c] = v[n][4];
      out[C + c] = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
      out[2 * C + c] = (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
    }
  }
}

// The same with 8-byte loads for even W: the three taps of a row lie inside two aligned pairs [q0, q0+1], [q1, q1+1]
// (q0 = even(x - 1) raised to 0, q1 = q0 + 2 lowered to W - 2: at the grid edge the clamped pair still holds the in-range
// taps at the positions the selects read, the out-of-range tap is replaced by the zero halo anyway): 6 requests per
// channel instead of 9 -- the gathers of this kernel are bound by the L1 tag stage (~3 lines per request), not by bytes.
template <int C, int NB>
__device__ __forceinline__ void perceive_pairs(const float* __restrict__ xs, int y, int x, int H, int W,
                                               float (&out)[3 * C]) {
  static_assert(C % NB == 0, "channel batches");
  const int HW = H * W;
  const bool up = y > 0, dn = y < H - 1, lf = x > 0, rt = x < W - 1;
  const int xa = (x - 1) & ~1;
  const int q0 = xa < 0 ? 0 : xa, q1 = xa + 2 > W - 2 ? W - 2 : xa + 2;
  const bool hi = ((x - 1) - xa) != 0;            // taps are (v1, v2, v3) instead of (v0, v1, v2)
  const int rU = (up ? y - 1 : y) * W, rM = y * W, rD = (dn ? y + 1 : y) * W;
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += NB) {
    float2 v[NB][3][2];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const float* pc = xs + (size_t)(c0 + n) * HW;
      v[n][0][0] = __ldg(reinterpret_cast<const float2*>(pc + rU + q0)); v[n][0][1] = __ldg(reinterpret_cast<const float2*>(pc + rU + q1));
      v[n][1][0] = __ldg(reinterpret_cast<const float2*>(pc + rM + q0)); v[n][1][1] = __ldg(reinterpret_cast<const float2*>(pc + rM + q1));
      v[n][2][0] = __ldg(reinterpret_cast<const float2*>(pc + rD + q0)); v[n][2][1] = __ldg(reinterpret_cast<const float2*>(pc + rD + q1));
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const int c = c0 + n;
      float t[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        t[r][0] = hi ? v[n][r][0].y : v[n][r][0].x;
        t[r][1] = hi ? v[n][r][1].x : v[n][r][0].y;
        t[r][2] = hi ? v[n][r][1].y : v[n][r][1].x;
      }
      const float a00 = (up && lf) ? t[0][0] : 0.f, a01 = up ? t[0][1] : 0.f, a02 = (up && rt) ? t[0][2] : 0.f;
      const float a10 = lf ? t[1][0] : 0.f, a12 = rt ? t[1][2] : 0.f;
      const float a20 = (dn && lf) ? t[2][0] : 0.f, a21 = dn ? t[2][1] : 0.f, a22 = (dn && rt) ? t[2][2] : 0.f;
      out[c] = t[1][1];
      out[C + c] = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
      out[2 * C + c] = (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
    }
  }
}

// xs / as of gather_senders<C> (graph_augmentation.py:104-158), two senders per round, straight-line: the offsets
// (reduced modulo the grid) and weights of the step wait in shared memory, all channels of both senders are requested
// first, then their 3x3 alive windows (edge taps clamped into the window: a duplicate never changes a max), and only then
// is anything consumed.  A sender that is out of range or not alive contributes with weight 0 (fmaf(0, finite, acc) ==
// acc: the bits of the generic function's `continue`).
template <int C>
__device__ __forceinline__ void gather_senders_pairs(const StepArgs& a, const float* __restrict__ xs_base, int b, int y, int x,
                                                     const int* __restrict__ s_dy, const int* __restrict__ s_dx,
                                                     const float* __restrict__ s_wt, float (&xs)[C], float& as) {
  const int H = a.H, W = a.W, HW = H * W, k = a.k;
  const bool torus = (a.flags & GNCA_F_TORUS) != 0, A2A = (a.flags & GNCA_F_ALIVE_TO_ALIVE) != 0;
  const float* __restrict__ alpha = xs_base + 3 * HW;
  const uint32_t* __restrict__ abits = a.alivebits ? a.alivebits + (size_t)b * ((HW + 31) >> 5) : nullptr;
#pragma unroll
  for (int c = 0; c < C; ++c) xs[c] = 0.f;
  as = 0.f;
#pragma unroll 1
  for (int i = 0; i < k; i += 2) {
    float w[2], v[2][C], al[2][9] = {};
    int qy[2], qx[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int ii = min(i + s, k - 1);
      const int dy = s_dy[ii], dx = s_dx[ii];
      int yy = y - dy, xx = torus ? x - dx : x;
      bool ok = i + s < k;
      if (torus) {
        yy = yy < 0 ? yy + H : (yy >= H ? yy - H : yy);
        xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
      } else {
        ok = ok && yy >= 0 && yy < H;
        yy = ok ? yy : y;
      }
      qy[s] = yy; qx[s] = xx;
      w[s] = ok ? s_wt[ii] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float* qp = xs_base + (qy[s] * W + qx[s]);
#pragma unroll
      for (int c = 0; c < C; ++c) v[s][c] = __ldg(qp + (size_t)c * HW);
    }
    uint32_t abit[2] = {0u, 0u};
    if (A2A) {
      if (a.alivebits) {            // k_compact left the sender-alive bit of every cell of x_in: one word instead of nine taps
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int q = qy[s] * W + qx[s];
          abit[s] = (__ldg(abits + (q >> 5)) >> (q & 31)) & 1u;
        }
      } else {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int ru = (qy[s] > 0 ? qy[s] - 1 : qy[s]) * W, rc = qy[s] * W, rd = (qy[s] < H - 1 ? qy[s] + 1 : qy[s]) * W;
          const int xl = qx[s] > 0 ? qx[s] - 1 : qx[s], xc = qx[s], xr = qx[s] < W - 1 ? qx[s] + 1 : qx[s];
          al[s][0] = __ldg(alpha + ru + xl); al[s][1] = __ldg(alpha + ru + xc); al[s][2] = __ldg(alpha + ru + xr);
          al[s][3] = __ldg(alpha + rc + xl); al[s][4] = __ldg(alpha + rc + xc); al[s][5] = __ldg(alpha + rc + xr);
          al[s][6] = __ldg(alpha + rd + xl); al[s][7] = __ldg(alpha + rd + xc); al[s][8] = __ldg(alpha + rd + xr);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float ww = w[s];
      if (A2A) {
        const float m = fmaxf(fmaxf(fmaxf(fmaxf(al[s][0], al[s][1]), fmaxf(al[s][2], al[s][3])),
                                    fmaxf(fmaxf(al[s][4], al[s][5]), fmaxf(al[s][6], al[s][7]))), al[s][8]);
        ww = (a.alivebits ? abit[s] != 0u : m > a.graph_alpha_thr) ? ww : 0.f;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) xs[c] = fmaf(ww, v[s][c], xs[c]);
      as += ww;
    }
  }
}

template <int C>
struct TcCols {                             // TMEM column map of the CTA's single tile pipeline
  static constexpr int K1 = 3 * C;
  // D1 holds the hi*hi products of layer 1 and then (in place) the hi part of h; Hl holds the two small cross terms
  // (hi*lo + lo*hi) of layer 1 and then (in place) the lo part of h; D2 / D2s the same split for layer 2.
  static constexpr int Yh = 0, Yl = K1, D1 = 2 * K1, Hl = 2 * K1 + kTcHid, D2 = 2 * K1 + 2 * kTcHid, D2s = D2 + C;
  static_assert(D2s + C <= 512, "TMEM columns");
};

template <int C>
static size_t tc_smem_bytes(int nchunks, bool graph) {
  size_t f = (size_t)2 * kTcHid * 3 * C + (size_t)2 * C * kTcHid + kTcHid + (graph ? C * C + C : 0);
  return f * sizeof(float) + kTcWG * 4 * 2 * sizeof(double) + kTcWG * sizeof(uint64_t) + 2 * sizeof(uint32_t) +
         (size_t)kTcWG * (nchunks + 1) * sizeof(int) + (graph ? (size_t)kTcWG * (C * 128 + 3 * GNCA_MAX_K) * sizeof(float) : 0) + 128;
}

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1) k_update_tc(StepArgs a, Packed P, const float* __restrict__ packed,
                                                             const uint16_t* __restrict__ glist,
                                                             const int* __restrict__ prefix) {
  constexpr int K1 = 3 * C;
  using TC = TcCols<C>;
  const int H = a.H, W = a.W, HW = H * W, nchunks = a.nchunks;
  const bool graph = (a.flags & GNCA_F_GRAPH) != 0;
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = tid >> 5, lane = tid & 31;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  float* sW1 = reinterpret_cast<float*>(smem_raw);              // [2][128 x 3C] canonical (hi, lo)
  float* sW2 = sW1 + 2 * kTcHid * K1;                           // [2][C x 128]
  float* sb1 = sW2 + 2 * C * kTcHid;
  float* sWmT = sb1 + kTcHid;
  float* sbm = sWmT + (graph ? C * C : 0);
  double* sred = reinterpret_cast<double*>(sbm + (graph ? C : 0));     // [3][4][2]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sred + kTcWG * 4 * 2);  // [3]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + kTcWG);
  int* s_pf = reinterpret_cast<int*>(tslot + 2) + g * (nchunks + 1);   // this warpgroup's prefix table
  // the tile's message waits here ([c][t], conflict-free) while perception and the tensor chain need the registers
  float* s_msg0 = reinterpret_cast<float*>(reinterpret_cast<int*>(tslot + 2) + kTcWG * (nchunks + 1));
  float* s_msg = s_msg0 + (size_t)g * C * 128 + t;
  // this warpgroup's copy of the step's sender offsets (reduced modulo the grid) and of the sample's attention weights
  int* s_dy = reinterpret_cast<int*>(s_msg0 + (size_t)kTcWG * C * 128) + g * 3 * GNCA_MAX_K;
  int* s_dx = s_dy + GNCA_MAX_K;
  float* s_wt = reinterpret_cast<float*>(s_dx + GNCA_MAX_K);

  block_copy(sW1, packed + P.w1c, 2 * kTcHid * K1);
  block_copy(sW2, packed + P.w2c, 2 * C * kTcHid);
  block_copy(sb1, packed + P.b1, kTcHid);
  if (graph) { block_copy(sWmT, packed + P.wmt, C * C); block_copy(sbm, packed + P.bm, C); }
  if (warp == 0) tmem_alloc(tslot, 512);
  if (tid == 32) {
    for (int i = 0; i < kTcWG; ++i) mbar_init(smem_u32(mbar + i), 1);
    mbar_init_fence();
  }
  fence_proxy_async();          // the weights were written through the generic proxy; the MMAs read them through the async one
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *tslot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tYh = tb + TC::Yh, tYl = tb + TC::Yl, tD1 = tb + TC::D1, tHl = tb + TC::Hl, tD2 = tb + TC::D2, tD2s = tb + TC::D2s;
  constexpr uint32_t idesc1 = idesc_tf32(128, kTcHid), idesc2 = idesc_tf32(128, C);
  const uint32_t w1h = smem_u32(sW1), w1l = smem_u32(sW1 + kTcHid * K1);
  const uint32_t w2h = smem_u32(sW2), w2l = smem_u32(sW2 + C * kTcHid);
  const uint32_t my_bar = smem_u32(mbar + g);
  uint32_t parity = 0;

  const float gain_m = graph ? step_message_gain(a) : 0.f;
  const bool do_msg = graph && gain_m != 0.f && a.k > 0;
  const int c_lo = ((a.flags & GNCA_F_HIDDEN_ONLY) && C >= 4) ? 4 : 0;
  const bool pairs_ok = !(W & 1) && W >= 4 && !(reinterpret_cast<uintptr_t>(a.x_in) & 7);   // 8-byte aligned tap pairs

  if (g == kTcWG - 1) bar_arrive(kBarTurn + 0, 256);            // the first turn belongs to warpgroup 0

#ifdef GNCA_PHASE_COUNTERS          /* development: where a tile's cycles go (thread 0 of every warpgroup of CTA 0 prints) */
  long long ph_t = clock64(), ph[6] = {0, 0, 0, 0, 0, 0};
  int ph_tiles = 0;
#define TC_MARK(i) do { const long long n_ = clock64(); ph[i] += n_ - ph_t; ph_t = n_; } while (0)
#else
#define TC_MARK(i) do { } while (0)
#endif
  const int n_units = nchunks * a.B;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int j = unit / a.B, b = unit - j * a.B;               // j-major: the non-empty units of all samples come first
    if (!sample_active(a, b)) continue;
    const int* pf = prefix + (size_t)b * (nchunks + 1);
    const int total = __ldg(pf + nchunks);
    const int r_lo = j * kTcRange;
    double* part = a.partials + ((size_t)b * a.npart + (size_t)j * kTcWG + g) * 2;
    if (r_lo >= total) {
      if (t == 0) { part[0] = 0.0; part[1] = 0.0; }
      continue;
    }
    const int r_hi = min(r_lo + kTcRange, total);
    for (int i = t; i <= nchunks; i += 128) s_pf[i] = __ldg(pf + i);
    if (do_msg && t < a.k) {
      int dy, dx;
      step_offset(a, t, dy, dx);
      const bool torus = (a.flags & GNCA_F_TORUS) != 0;
      s_dy[t] = torus ? dy % H : dy; s_dx[t] = torus ? dx % W : 0;
      s_wt[t] = a.attn_w ? __ldg(a.attn_w + (size_t)b * a.k + t) : 1.0f / (float)a.k;
    }
    bar_sync(kBarWg + g, 128);
    const float* xs_base = a.x_in + (size_t)b * C * HW;
    float s1 = 0.f, s2 = 0.f;
    // rank -> cell: binary search in the chunk prefix table, then the chunk's list; done one tile ahead (the list entry
    // arrives during the tensor chain)
    auto locate = [&](int rank) -> int {
      if (rank >= r_hi) return -1;
      int lo = 0, hi = nchunks;                                 // s_pf[lo] <= rank < s_pf[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_pf[mid] <= rank) lo = mid; else hi = mid; }
      return lo * kTcChunk + (int)__ldg(glist + (size_t)b * HW + (size_t)lo * kTcChunk + (rank - s_pf[lo]));
    };
    int cell_next = locate(r_lo + g * 128 + t);
    for (int base = r_lo; base < r_hi; base += kTcThreads) {
      const int cell = cell_next;
      const int cy = cell >= 0 ? cell / W : 0, cx = cell >= 0 ? cell - cy * W : 0;
      TC_MARK(0);
      // ---- graph message (CUDA cores; graph_augmentation.py:104-169 by linearity, ncagraph.py:94-104,141) ----------
      if (do_msg) {
        if (cell >= 0) {
          float xsnd[C], as;
          gather_senders_pairs<C>(a, xs_base, b, cy, cx, s_dy, s_dx, s_wt, xsnd, as);
          // agg = Wm xs + bm as (msg_proj by linearity), message = gain * tanh(agg) on the updated channels
          // (ncagraph.py:94-104,141): a rolled loop over quads of output channels keeps the kernel's code inside the
          // instruction cache
#pragma unroll 1
          for (int c4 = 0; c4 < C; c4 += 4) {
            const float4 bq = *reinterpret_cast<const float4*>(sbm + c4);
            float g0 = bq.x * as, g1 = bq.y * as, g2 = bq.z * as, g3 = bq.w * as;
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
              const float4 wq = *reinterpret_cast<const float4*>(sWmT + ci * C + c4);
              g0 = fmaf(wq.x, xsnd[ci], g0); g1 = fmaf(wq.y, xsnd[ci], g1);
              g2 = fmaf(wq.z, xsnd[ci], g2); g3 = fmaf(wq.w, xsnd[ci], g3);
            }
            s_msg[(c4 + 0) * 128] = c4 + 0 >= c_lo ? tanhf(g0) * gain_m : 0.f;
            s_msg[(c4 + 1) * 128] = c4 + 1 >= c_lo ? tanhf(g1) * gain_m : 0.f;
            s_msg[(c4 + 2) * 128] = c4 + 2 >= c_lo ? tanhf(g2) * gain_m : 0.f;
            s_msg[(c4 + 3) * 128] = c4 + 3 >= c_lo ? tanhf(g3) * gain_m : 0.f;
          }
        }
      }
      TC_MARK(1);
      // ---- perception (perception.py:21-26) --------------------------------------------------------------------------
      float yv[K1];
      // an idle lane reads the window of cell 0 and drops it
      if (pairs_ok) perceive_pairs<C, kTcNB>(xs_base, cy, cx, H, W, yv);
      else perceive_batched<C, kTcNB>(xs_base, cy, cx, H, W, yv);
      if (cell < 0) {
#pragma unroll
        for (int k = 0; k < K1; ++k) yv[k] = 0.f;
      }
      // ================= tensor-core chain: one warpgroup at a time ===================================================
      TC_MARK(2);
      cell_next = locate(base + kTcThreads + g * 128 + t);
      bar_sync(kBarTurn + g, 256);
      fence_after();
      TC_MARK(3);
#pragma unroll
      for (int c0 = 0; c0 < K1; c0 += 16) {
        uint32_t vh[16], vl[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) split_tf32_fast(yv[c0 + q], vh[q], vl[q]);
        tmem_st16(tYh + lane_base + c0, vh);
        tmem_st16(tYl + lane_base + c0, vl);
      }
      wait_st();
      fence_before();
      bar_sync(kBarWg + g, 128);
      if (t == 0) {                                             // layer 1: D1 = y W1^T  (update_net.0, ncagraph.py:61)
        fence_after();
        constexpr uint32_t sbo = (K1 / 4) * 128;
#pragma unroll 1
        for (int ks = 0; ks < K1 / 8; ++ks) {
          const uint64_t bh = smem_desc(w1h + ks * 256, 128, sbo), bl = smem_desc(w1l + ks * 256, 128, sbo);
          mma_ts(tD1, tYh + ks * 8, bh, idesc1, ks > 0);
          mma_ts(tHl, tYh + ks * 8, bl, idesc1, ks > 0);
          mma_ts(tHl, tYl + ks * 8, bh, idesc1, 1);
        }
        mma_commit(my_bar);
      }
      mbar_wait(my_bar, parity); parity ^= 1;
      fence_after();
#pragma unroll 2
      for (int c0 = 0; c0 < kTcHid; c0 += 16) {                 // epilogue 1: + b1, ReLU (update_net.1), split, back to TMEM
        uint32_t v[16], vl[16];
        tmem_ld16(tD1 + lane_base + c0, v);
        tmem_ld16(tHl + lane_base + c0, vl);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float h = fmaxf((__uint_as_float(v[q]) + __uint_as_float(vl[q])) + sb1[c0 + q], 0.f);
          split_tf32_fast(h, v[q], vl[q]);
        }
        tmem_st16(tD1 + lane_base + c0, v);
        tmem_st16(tHl + lane_base + c0, vl);
      }
      wait_st();
      fence_before();
      bar_sync(kBarWg + g, 128);
      if (t == 0) {                                             // layer 2: D2 = h W2^T  (update_net.2, no bias)
        fence_after();
        constexpr uint32_t sbo = (kTcHid / 4) * 128;
#pragma unroll 1
        for (int ks = 0; ks < kTcHid / 8; ++ks) {
          const uint64_t bh = smem_desc(w2h + ks * 256, 128, sbo), bl = smem_desc(w2l + ks * 256, 128, sbo);
          mma_ts(tD2, tD1 + ks * 8, bh, idesc2, ks > 0);
          mma_ts(tD2s, tD1 + ks * 8, bl, idesc2, ks > 0);
          mma_ts(tD2s, tHl + ks * 8, bh, idesc2, 1);
        }
        mma_commit(my_bar);
      }
      mbar_wait(my_bar, parity); parity ^= 1;
      fence_after();
      float dxv[C];
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 16) {
        uint32_t v[16], vs[16];
        tmem_ld16(tD2 + lane_base + c0, v);
        tmem_ld16(tD2s + lane_base + c0, vs);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) dxv[c0 + q] = __uint_as_float(v[q]) + __uint_as_float(vs[q]);
      }
      fence_before();
      bar_arrive(kBarTurn + (g + 1) % kTcWG, 256);              // pass the token
      TC_MARK(4);
      // ================================================================================================================
      if (cell >= 0) {
        float* up = a.u + (size_t)b * C * HW + cell;
        if (do_msg) {
#pragma unroll
          for (int c = 0; c < C; ++c) dxv[c] += s_msg[c * 128];
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float v = dxv[c];
          up[(size_t)c * HW] = v;
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
      }
      TC_MARK(5);
#ifdef GNCA_PHASE_COUNTERS
      ++ph_tiles;
#endif
    }
    // deterministic per-warpgroup partial of (sum u, sum u^2): fixed shuffle tree, warps added in order
    const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
    if (lane == 0) { sred[(g * 4 + (warp & 3)) * 2] = d1; sred[(g * 4 + (warp & 3)) * 2 + 1] = d2; }
    bar_sync(kBarWg + g, 128);
    if (t == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int w = 0; w < 4; ++w) { t1 += sred[(g * 4 + w) * 2]; t2 += sred[(g * 4 + w) * 2 + 1]; }
      part[0] = t1; part[1] = t2;
    }
  }
#ifdef GNCA_PHASE_COUNTERS
  if (blockIdx.x == 0 && t == 0 && a.t == 2 && ph_tiles > 0)
    printf("[k_update_tc phases, CTA 0 warpgroup %d, %d tiles] per tile: lookup %lld  message %lld  perception %lld  wait-turn %lld  "
           "tensor chain %lld  store %lld cycles\n", g, ph_tiles, ph[0] / ph_tiles, ph[1] / ph_tiles, ph[2] / ph_tiles,
           ph[3] / ph_tiles, ph[4] / ph_tiles, ph[5] / ph_tiles);
#endif
#undef TC_MARK
  if (g == 0) bar_sync(kBarTurn + 0, 256);                      // absorb the last token
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tb, 512);
}

template <int C>
static int launch_tc(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const uint16_t* glist,
                     const int* prefix, cudaStream_t st) {
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const size_t smem = tc_smem_bytes<C>(a.nchunks, graph);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    GNCA_CHECK_CUDA(cudaGetDevice(&dev));
    GNCA_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  GNCA_CHECK_CUDA(cudaFuncSetAttribute(k_update_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int units = a.nchunks * a.B;
  k_update_tc<C><<<units < n_sm ? units : n_sm, kTcThreads, smem, st>>>(a, P, packed, glist, prefix);
  return 0;
}

bool update_tc_supported(const gnca_model& m, const StepArgs& a) {
  if (!(m.C == 16 || m.C == 32) || m.hidden != kTcHid) return false;
  if (a.chunk != kTcChunk) return false;
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const size_t smem = m.C == 16 ? tc_smem_bytes<16>(a.nchunks, graph) : tc_smem_bytes<32>(a.nchunks, graph);
  return smem <= 227 * 1024;
}

int launch_update_tc(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, const uint16_t* glist,
                     const int* prefix, cudaStream_t st) {
  a.npart = kTcWG * a.nchunks;
  if (m.C == 16) return launch_tc<16>(m, P, packed, a, glist, prefix, st);
  return launch_tc<32>(m, P, packed, a, glist, prefix, st);
}

}  // namespace gnca
