// k_update_tc2: the streaming step's update on tensor cores over DENSE 8x16 tiles staged by TMA (sm_100a).
//
// k_update_tc (gnca_update_tc.cu) compacts the active cells of a sample and lets every thread gather its cell's 3x3x C
// neighbourhood and its k senders straight from global memory: ncu (profiles/r02_k_update_tc_v1_summary.md) shows it
// latency-bound -- long_scoreboard 4.2 stalls per issue, L2 at 9 % of its throughput, tensor pipe 9 % -- because a
// thread keeps only a handful of scalar loads in flight.  Here the unit of work is a dense TILE of 8 x 16 cells:
//   * k_tilemask   one warp per tile: alive & fire of its 128 cells -> a 128-bit active mask, the sender-alive byte plane
//                  (graph_augmentation.py:116-117) and the count of active cells;
//   * k_tilescan   deterministic compaction of the tiles that hold at least one active cell (+ prefix over the samples);
//   * k_update_tc2 persistent CTA per SM, three warpgroups, thread t <-> cell t of the tile.  The tile's perception
//                  footprint [C][8+2][4+16+4] comes in with ONE bulk tensor copy (cp.async.bulk.tensor.3d, zero fill
//                  outside the grid = the perception's zero halo, perception.py:16) signalled on an mbarrier; the copy of
//                  the warpgroup's NEXT tile is issued as soon as the current one has been read, so it flies during the
//                  sender gather and the tensor-core chain.  MLP: the 3xTF32 tcgen05 chain of k_update_tc (activations
//                  in TMEM, split accumulators).  The MMAs also run over the inactive cells of a tile (their rows are
//                  discarded): tensor work is cheap, loads in flight are what the kernel was short of.
// Same outputs as k_update / k_update_tc: masked pre-norm update u of the active cells, GroupNorm partial sums (one
// float2 per tile, reduced in fixed order by k_tilestats -> (mean, rstd) per sample).
#include <cuda.h>
#include "gnca_common.cuh"
#include "gnca_internal.h"
#include "gnca_tc.cuh"
#include <cstdio>

namespace gnca {

using namespace tc;

constexpr int kT2WG = 3;
constexpr int kT2Threads = kT2WG * 128;
constexpr int kT2Hid = 128;
constexpr int kTileH = 8, kTileW = 16;           // 128 cells: thread t <-> (t >> 4, t & 15)
constexpr int kBoxH = kTileH + 2, kBoxW = 24;    // rows: halo of 1.  Columns: the box must START on a 16-byte boundary of the row
constexpr int kBoxX0 = 4;                        // (a misaligned innermost coordinate is an illegal instruction: scratch/
                                                 // tma_probe3.cu), so it starts 4 cells left of the tile and is 4 + 16 + 4 wide
constexpr int kBarTurn2 = 1, kBarWg2 = 4;

// ------------------------------------------------------------------------------------------------
// tile masks
// ------------------------------------------------------------------------------------------------
// one warp per tile; lane l owns cells t = l + 32 j (j < 4): rows 2j + (l >> 4), column l & 15
__global__ void __launch_bounds__(256) k_tilemask(StepArgs a, int C, int tiles_x, int ntiles, uint32_t* __restrict__ tmask,
                                                  int* __restrict__ tcnt, uint8_t* __restrict__ salive,
                                                  float2* __restrict__ tpart) {
  const int b = blockIdx.y, tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= ntiles) return;
  const int H = a.H, W = a.W, HW = H * W;
  const size_t ti = (size_t)b * ntiles + tile;
  if (!sample_active(a, b)) {
    if (lane == 0) { tcnt[ti] = 0; tpart[ti] = make_float2(0.f, 0.f); }
    return;
  }
  const float fr = step_fire_rate(a);
  const float* alpha = a.x_in + (size_t)b * C * HW + 3 * HW;
  const int ty0 = (tile / tiles_x) * kTileH, tx0 = (tile % tiles_x) * kTileW;
  const bool same_thr = a.graph_alpha_thr == a.alpha_thr;
  const bool want_s = (a.flags & GNCA_F_GRAPH) && (a.flags & GNCA_F_ALIVE_TO_ALIVE);
  int n = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = ty0 + 2 * j + (lane >> 4), x = tx0 + (lane & 15), cell = y * W + x;
    const bool alive = alive_at(alpha, y, x, H, W, a.alpha_thr);
    if (want_s) salive[(size_t)b * HW + cell] = same_thr ? alive : alive_at(alpha, y, x, H, W, a.graph_alpha_thr);
    const bool act = alive && fires(a, fr, b, cell);
    const uint32_t w = __ballot_sync(0xffffffffu, act);
    if (lane == 0) tmask[ti * 4 + j] = w;
    n += __popc(w);
  }
  if (lane == 0) {
    tcnt[ti] = n;
    if (n == 0) tpart[ti] = make_float2(0.f, 0.f);
  }
}

// one block: per-sample compaction of the tiles with active cells (order = tile index), then the prefix over samples
__global__ void __launch_bounds__(1024) k_tilescan(int B, int ntiles, const int* __restrict__ tcnt, int* __restrict__ tlist,
                                                   int* __restrict__ tprefix) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += 32) {
    int carry = 0;
    for (int i0 = 0; i0 < ntiles; i0 += 32) {
      const int i = i0 + lane;
      const bool f = i < ntiles && tcnt[(size_t)b * ntiles + i] > 0;
      const uint32_t w = __ballot_sync(0xffffffffu, f);
      if (f) tlist[(size_t)b * ntiles + carry + __popc(w & ((1u << lane) - 1u))] = i;
      carry += __popc(w);
    }
    if (lane == 0) tprefix[b + 1] = carry;            // counts first; prefix below
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    tprefix[0] = 0;
    for (int b = 0; b < B; ++b) { s += tprefix[b + 1]; tprefix[b + 1] = s; }
  }
}

// (mean, rstd) of every sample from the per-tile partials, fixed order (one warp per sample)
__global__ void k_tilestats(StepArgs a, int C, int ntiles, const float2* __restrict__ tpart, float* __restrict__ stats) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float mu = 0.f, rstd = 1.f;
  if (a.flags & GNCA_F_GROUPNORM) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < ntiles; i += 32) {
      const float2 p = tpart[(size_t)b * ntiles + i];
      s1 += (double)p.x; s2 += (double)p.y;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const double n = (double)C * (double)a.H * (double)a.W;
    const double m = s1 / n;
    double var = s2 / n - m * m;
    if (var < 0.0) var = 0.0;
    mu = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)a.gn_eps));
  }
  if (lane == 0) { stats[b * 2] = mu; stats[b * 2 + 1] = rstd; }
}

// ------------------------------------------------------------------------------------------------
// TMA helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, int x, int y, int z, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

template <int C>
struct Tc2Cols {
  static constexpr int K1 = 3 * C;
  static constexpr int Yh = 0, Yl = K1, D1 = 2 * K1, Hl = 2 * K1 + kT2Hid, D2 = 2 * K1 + 2 * kT2Hid, D2s = D2 + C;
  static_assert(D2s + C <= 512, "TMEM columns");
};

template <int C>
static size_t tc2_smem_bytes(int B, bool graph) {
  size_t f = (size_t)2 * kT2Hid * 3 * C + (size_t)2 * C * kT2Hid + kT2Hid + (graph ? C * C + C : 0);
  return 128 + f * sizeof(float) + 128 + (size_t)kT2WG * C * kBoxH * kBoxW * sizeof(float) + kT2WG * 4 * 2 * sizeof(float) +
         2 * kT2WG * sizeof(uint64_t) + 16 + (size_t)(B + 1) * sizeof(int);
}

template <int C>
__global__ void __launch_bounds__(kT2Threads, 1) k_update_tc2(StepArgs a, Packed P, const float* __restrict__ packed,
                                                              const __grid_constant__ CUtensorMap tmap,
                                                              const uint32_t* __restrict__ tmask, const int* __restrict__ tlist,
                                                              const int* __restrict__ tprefix,
                                                              const uint8_t* __restrict__ salive, float2* __restrict__ tpart,
                                                              int ntiles, int tiles_x) {
  constexpr int K1 = 3 * C;
  constexpr int kBox = C * kBoxH * kBoxW;            // floats of one staged tile
  using TC = Tc2Cols<C>;
  const int H = a.H, W = a.W, HW = H * W, B = a.B;
  const bool graph = (a.flags & GNCA_F_GRAPH) != 0;
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = tid >> 5, lane = tid & 31;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  float* sW1 = reinterpret_cast<float*>(base);                  // [2][128 x 3C] canonical (hi, lo)
  float* sW2 = sW1 + 2 * kT2Hid * K1;                           // [2][C x 128]
  float* sb1 = sW2 + 2 * C * kT2Hid;
  float* sWmT = sb1 + kT2Hid;
  float* sbm = sWmT + (graph ? C * C : 0);
  float* endw = sbm + (graph ? C : 0);
  float* sXall = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(endw) + 127) & ~(uintptr_t)127);   // [3][C][10][20]
  float* sred = sXall + kT2WG * kBox;                           // [3][4][2]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sred + kT2WG * 4 * 2);   // [3] MMA completion, [3] TMA completion
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 2 * kT2WG);
  int* s_tpf = reinterpret_cast<int*>(tslot + 4);               // [B + 1] prefix of the tile counts over the samples
  float* sX = sXall + g * kBox;

  block_copy(sW1, packed + P.w1c, 2 * kT2Hid * K1);
  block_copy(sW2, packed + P.w2c, 2 * C * kT2Hid);
  block_copy(sb1, packed + P.b1, kT2Hid);
  if (graph) { block_copy(sWmT, packed + P.wmt, C * C); block_copy(sbm, packed + P.bm, C); }
  for (int i = tid; i <= B; i += kT2Threads) s_tpf[i] = tprefix[i];
  if (warp == 0) tmem_alloc(tslot, 512);
  if (tid == 32) {
    for (int i = 0; i < 2 * kT2WG; ++i) mbar_init(smem_u32(mbar + i), 1);
    mbar_init_fence();
  }
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = *tslot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tYh = tb + TC::Yh, tYl = tb + TC::Yl, tD1 = tb + TC::D1, tHl = tb + TC::Hl, tD2 = tb + TC::D2, tD2s = tb + TC::D2s;
  constexpr uint32_t idesc1 = idesc_tf32(128, kT2Hid), idesc2 = idesc_tf32(128, C);
  const uint32_t w1h = smem_u32(sW1), w1l = smem_u32(sW1 + kT2Hid * K1);
  const uint32_t w2h = smem_u32(sW2), w2l = smem_u32(sW2 + C * kT2Hid);
  const uint32_t mma_bar = smem_u32(mbar + g), tma_bar = smem_u32(mbar + kT2WG + g), sx_addr = smem_u32(sX);
  uint32_t mma_par = 0, tma_par = 0;

  const float gain_m = graph ? step_message_gain(a) : 0.f;
  const bool do_msg = graph && gain_m != 0.f && a.k > 0;
  const bool a2a = (a.flags & GNCA_F_ALIVE_TO_ALIVE) != 0, torus = (a.flags & GNCA_F_TORUS) != 0;
  const int c_lo = ((a.flags & GNCA_F_HIDDEN_ONLY) && C >= 4) ? 4 : 0;
  const int total = s_tpf[B];
  const int ty = t >> 4, tx = t & 15;

  // global tile index -> (sample, tile)
  auto locate = [&](int gi, int& b, int& tile) {
    int lo = 0, hi = B;                                          // s_tpf[lo] <= gi < s_tpf[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_tpf[mid] <= gi) lo = mid; else hi = mid; }
    b = lo;
    tile = __ldg(tlist + (size_t)lo * ntiles + (gi - s_tpf[lo]));
  };
  auto issue_tma = [&](int b, int tile) {                        // one thread: stage the tile's perception footprint
    const int ty0 = (tile / tiles_x) * kTileH, tx0 = (tile % tiles_x) * kTileW;
    mbar_expect_tx(tma_bar, (uint32_t)(kBox * sizeof(float)));
    tma_load_3d(sx_addr, &tmap, tx0 - kBoxX0, ty0 - 1, b * C, tma_bar);
  };

#ifdef GNCA_PHASE_COUNTERS
  long long ph_t = clock64(), ph[7] = {0, 0, 0, 0, 0, 0, 0};
  int ph_tiles = 0;
#define T2_MARK(i) do { const long long n_ = clock64(); ph[i] += n_ - ph_t; ph_t = n_; } while (0)
#else
#define T2_MARK(i) do { } while (0)
#endif
  if (g == kT2WG - 1) bar_arrive(kBarTurn2 + 0, 256);           // the first turn on the tensor chain belongs to warpgroup 0
  {
    const int g0 = blockIdx.x + g * gridDim.x;
    if (t == 0 && g0 < total) { int b0, tile0; locate(g0, b0, tile0); issue_tma(b0, tile0); }
  }

  for (int r = 0; blockIdx.x + 3 * r * (int)gridDim.x < total; ++r) {
    const int gi = blockIdx.x + (3 * r + g) * gridDim.x;
    const bool have = gi < total;
    int b = 0, tile = 0;
    if (have) locate(gi, b, tile);
    const int ty0 = (tile / tiles_x) * kTileH, tx0 = (tile % tiles_x) * kTileW;
    const int cy = ty0 + ty, cx = tx0 + tx, cell = cy * W + cx;
    const bool act = have && ((__ldg(tmask + ((size_t)b * ntiles + tile) * 4 + (t >> 5)) >> (t & 31)) & 1u);
    T2_MARK(0);
    // ---- graph message of the ACTIVE cells (graph_augmentation.py:104-169 by linearity, ncagraph.py:94-104,141); it needs
    //      nothing from the staged tile, so it runs first: msg[C] + y[3C] then fit the 168-register budget without spills ----
    float msg[C];
#pragma unroll
    for (int c = 0; c < C; ++c) msg[c] = 0.f;
    if (do_msg && act) {
      const float* xs_base = a.x_in + (size_t)b * C * HW;
      float xsnd[C], as = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) xsnd[c] = 0.f;
      const float wuni = 1.0f / (float)a.k;
      for (int i = 0; i < a.k; ++i) {
        int dy, dx, qy, qx;
        step_offset(a, i, dy, dx);
        if (!sender_of(cy, cx, dy, dx, H, W, torus, qy, qx)) continue;
        const int q = qy * W + qx;
        if (a2a && !__ldg(salive + (size_t)b * HW + q)) continue;
        const float w = a.attn_w ? __ldg(a.attn_w + (size_t)b * a.k + i) : wuni;
        const float* qp = xs_base + q;
#pragma unroll
        for (int c = 0; c < C; ++c) xsnd[c] = fmaf(w, __ldg(qp + (size_t)c * HW), xsnd[c]);
        as += w;
      }
      float agg[C];
      msg_project<C>(xsnd, as, sWmT, sbm, agg);
#pragma unroll
      for (int c = 0; c < C; ++c) msg[c] = c >= c_lo ? tanhf(agg[c]) * gain_m : 0.f;
    }
    T2_MARK(1);
    // ---- perception from the staged tile (perception.py:21-26) ------------------------------------------------------
    float yv[K1];
    if (have) {
      mbar_wait(tma_bar, tma_par); tma_par ^= 1;
      T2_MARK(2);
      const float* p0 = sX + ty * kBoxW + tx + (kBoxX0 - 1);       // p0[0] = cell (cy - 1, cx - 1)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* p = p0 + c * (kBoxH * kBoxW);
        const float a00 = p[0], a01 = p[1], a02 = p[2];
        const float a10 = p[kBoxW], a11 = p[kBoxW + 1], a12 = p[kBoxW + 2];
        const float a20 = p[2 * kBoxW], a21 = p[2 * kBoxW + 1], a22 = p[2 * kBoxW + 2];
        yv[c] = a11;
        yv[C + c] = (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
        yv[2 * C + c] = (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
      }
    } else {
#pragma unroll
      for (int k = 0; k < K1; ++k) yv[k] = 0.f;
    }
    bar_sync(kBarWg2 + g, 128);                                  // everybody has read the tile: the buffer may be refilled
    {
      const int gn = gi + 3 * (int)gridDim.x;
      if (t == 0 && gn < total) { int b2, tile2; locate(gn, b2, tile2); issue_tma(b2, tile2); }
    }
    T2_MARK(3);
    // ================= tensor-core chain: one warpgroup at a time (as k_update_tc) ===================================
    bar_sync(kBarTurn2 + g, 256);
    fence_after();
    T2_MARK(4);
#pragma unroll
    for (int c0 = 0; c0 < K1; c0 += 16) {
      uint32_t vh[16], vl[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) split_tf32(yv[c0 + q], vh[q], vl[q]);
      tmem_st16(tYh + lane_base + c0, vh);
      tmem_st16(tYl + lane_base + c0, vl);
    }
    wait_st();
    fence_before();
    bar_sync(kBarWg2 + g, 128);
    if (t == 0) {                                               // layer 1 (update_net.0): hi*hi -> D1, cross terms -> Hl
      fence_after();
      constexpr uint32_t sbo = (K1 / 4) * 128;
#pragma unroll 1
      for (int ks = 0; ks < K1 / 8; ++ks) {
        const uint64_t bh = smem_desc(w1h + ks * 256, 128, sbo), bl = smem_desc(w1l + ks * 256, 128, sbo);
        mma_ts(tD1, tYh + ks * 8, bh, idesc1, ks > 0);
        mma_ts(tHl, tYh + ks * 8, bl, idesc1, ks > 0);
        mma_ts(tHl, tYl + ks * 8, bh, idesc1, 1);
      }
      mma_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_par); mma_par ^= 1;
    fence_after();
#pragma unroll
    for (int c0 = 0; c0 < kT2Hid; c0 += 16) {                   // epilogue 1: + b1, ReLU, split, back to TMEM
      uint32_t v[16], vl[16];
      tmem_ld16(tD1 + lane_base + c0, v);
      tmem_ld16(tHl + lane_base + c0, vl);
      wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float h = fmaxf((__uint_as_float(v[q]) + __uint_as_float(vl[q])) + sb1[c0 + q], 0.f);
        split_tf32(h, v[q], vl[q]);
      }
      tmem_st16(tD1 + lane_base + c0, v);
      tmem_st16(tHl + lane_base + c0, vl);
    }
    wait_st();
    fence_before();
    bar_sync(kBarWg2 + g, 128);
    if (t == 0) {                                               // layer 2 (update_net.2, no bias)
      fence_after();
      constexpr uint32_t sbo = (kT2Hid / 4) * 128;
#pragma unroll 1
      for (int ks = 0; ks < kT2Hid / 8; ++ks) {
        const uint64_t bh = smem_desc(w2h + ks * 256, 128, sbo), bl = smem_desc(w2l + ks * 256, 128, sbo);
        mma_ts(tD2, tD1 + ks * 8, bh, idesc2, ks > 0);
        mma_ts(tD2s, tD1 + ks * 8, bl, idesc2, ks > 0);
        mma_ts(tD2s, tHl + ks * 8, bh, idesc2, 1);
      }
      mma_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_par); mma_par ^= 1;
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
    bar_arrive(kBarTurn2 + (g + 1) % kT2WG, 256);               // pass the token
    T2_MARK(5);
    // =================================================================================================================
    float s1 = 0.f, s2 = 0.f;
    if (act) {
      float* up = a.u + (size_t)b * C * HW + cell;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v = dxv[c] + msg[c];
        up[(size_t)c * HW] = v;
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
    // deterministic per-tile partial of (sum u, sum u^2)
    const float f1 = warp_sum(s1), f2 = warp_sum(s2);
    if (lane == 0) { sred[(g * 4 + (warp & 3)) * 2] = f1; sred[(g * 4 + (warp & 3)) * 2 + 1] = f2; }
    bar_sync(kBarWg2 + g, 128);
    if (t == 0 && have) {
      float t1 = 0.f, t2 = 0.f;
      for (int w = 0; w < 4; ++w) { t1 += sred[(g * 4 + w) * 2]; t2 += sred[(g * 4 + w) * 2 + 1]; }
      tpart[(size_t)b * ntiles + tile] = make_float2(t1, t2);
    }
    T2_MARK(6);
#ifdef GNCA_PHASE_COUNTERS
    ++ph_tiles;
#endif
  }
#ifdef GNCA_PHASE_COUNTERS
  if (blockIdx.x == 0 && t == 0 && a.t == 2 && ph_tiles > 0)
    printf("[k_update_tc2 phases, CTA 0 warpgroup %d, %d tiles] per tile: locate %lld  message %lld  wait-tma %lld  perception+prefetch %lld  "
           "wait-turn %lld  tensor chain %lld  store+reduce %lld cycles\n", g, ph_tiles, ph[0] / ph_tiles, ph[1] / ph_tiles, ph[2] / ph_tiles,
           ph[3] / ph_tiles, ph[4] / ph_tiles, ph[5] / ph_tiles, ph[6] / ph_tiles);
#endif
#undef T2_MARK
  if (g == 0) bar_sync(kBarTurn2 + 0, 256);                     // absorb the last token
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tb, 512);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

size_t update_tc2_workspace_bytes(int B, int H, int W) {
  const size_t ntiles = (size_t)((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  return al(B * ntiles * 16) + 2 * al(B * ntiles * 4) + al((size_t)(B + 1) * 4) + al((size_t)B * H * W) + al(B * ntiles * 8) +
         al((size_t)B * 2 * 4);
}

bool update_tc2_supported(const gnca_model& m, const StepArgs& a) {
  if (!(m.C == 16 || m.C == 32) || m.hidden != kT2Hid) return false;
  if (a.H % kTileH || a.W % kTileW || a.B > 4096) return false;
  if (!encode_tiled_fn()) return false;
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const size_t smem = m.C == 16 ? tc2_smem_bytes<16>(a.B, graph) : tc2_smem_bytes<32>(a.B, graph);
  return smem <= 227 * 1024;
}

template <int C>
static int launch_tc2(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, void* ws, float* stats_out,
                      cudaStream_t st) {
  const bool graph = (m.flags & GNCA_F_GRAPH) != 0;
  const int B = a.B, H = a.H, W = a.W;
  const int tiles_x = W / kTileW, ntiles = tiles_x * (H / kTileH);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  char* p = reinterpret_cast<char*>(ws);
  uint32_t* tmask = reinterpret_cast<uint32_t*>(p); p += al((size_t)B * ntiles * 16);
  int* tcnt = reinterpret_cast<int*>(p); p += al((size_t)B * ntiles * 4);
  int* tlist = reinterpret_cast<int*>(p); p += al((size_t)B * ntiles * 4);
  int* tprefix = reinterpret_cast<int*>(p); p += al((size_t)(B + 1) * 4);
  uint8_t* salive = reinterpret_cast<uint8_t*>(p); p += al((size_t)B * H * W);
  float2* tpart = reinterpret_cast<float2*>(p); p += al((size_t)B * ntiles * 8);
  float* stats_tmp = reinterpret_cast<float*>(p);

  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {kBoxW, kBoxH, (cuuint32_t)C};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = encode_tiled_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.x_in), gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return GNCA_ERR_UNSUPPORTED;

  k_tilemask<<<dim3((ntiles + 7) / 8, B), 256, 0, st>>>(a, C, tiles_x, ntiles, tmask, tcnt, salive, tpart);
  k_tilescan<<<1, 1024, 0, st>>>(B, ntiles, tcnt, tlist, tprefix);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    GNCA_CHECK_CUDA(cudaGetDevice(&dev));
    GNCA_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = tc2_smem_bytes<C>(B, graph);
  GNCA_CHECK_CUDA(cudaFuncSetAttribute(k_update_tc2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_update_tc2<C><<<n_sm, kT2Threads, smem, st>>>(a, P, packed, tmap, tmask, tlist, tprefix, salive, tpart, ntiles, tiles_x);
  float* stats = stats_out ? stats_out : stats_tmp;
  k_tilestats<<<B, 32, 0, st>>>(a, C, ntiles, tpart, stats);
  g_launches += 3;
  a.stats_ready = stats;
  return 0;
}

int launch_update_tc2(const gnca_model& m, const Packed& P, const float* packed, StepArgs& a, void* ws, cudaStream_t st) {
  if (m.C == 16) return launch_tc2<16>(m, P, packed, a, ws, a.stats, st);
  return launch_tc2<32>(m, P, packed, a, ws, a.stats, st);
}

}  // namespace gnca
