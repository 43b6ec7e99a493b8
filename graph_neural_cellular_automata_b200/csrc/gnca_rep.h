// Shared between the replicated-state forward (gnca_rep.cu) and the resident backward (gnca_rep_bwd.cu):
// layout of what the forward keeps for BPTT instead of the dense x_t / u_t history.
#pragma once
#include <stdint.h>

namespace gnca {

// One record per ACTIVE cell (fire & pre-alive) of a step, in cell order, at
//   rec[((t*B + b)*HW + slot) * kRecStride],  slot = rank of the cell among the active cells of (t, b).
// Forward writes  y[48] (perception: identity | sobel_x | sobel_y), u[16] (masked pre-norm update),
// xs[16] (gathered sender state sum_i w x(q_i) A(q_i)), th[16] = tanh(Wm xs + bm as) and as.
// The backward overwrites u with gd = dL/du and th with gm = dL/d(Wm xs + bm as); the batched weight-gradient
// kernel then needs nothing but the records.
constexpr int kRecY = 0, kRecU = 48, kRecXs = 64, kRecTh = 80, kRecAs = 96, kRecStride = 100;
// The backward additionally keeps, per record, the hidden layer h[128] and its gradient gh[128] (unit order permuted:
// index 4*l + jj <-> unit l + 32*jj) at hgh[((t*B + b)*HW + slot) * kHghStride]: the weight-gradient pass then is a pure
// stream of outer products.
constexpr int kHghStride = 256;
// bitmaps of a step: [0] sender-alive (graph_alpha_thr), [1] active = fire & pre-alive, [2] post-alive
constexpr int kMaskWords = 64;
// blocks of the batched weight-gradient kernel (two per SM)
constexpr int kMaxWgradBlocks = 296;

}  // namespace gnca
