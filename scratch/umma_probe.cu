// tcgen05 probe for the update-MLP of the NCA step (sm_100a): is a 3xTF32 UMMA with the activations in TENSOR MEMORY
// (A operand from TMEM, thread t <-> cell t <-> TMEM lane t) accurate and fast enough to replace the FFMA MLP?
//   T1  SS  D[128x128] = A[128x96] * B[128x96]^T, one TF32 pass, operands in the no-swizzle K-major canonical layout
//   T2  TS  same product with A written to TMEM by tcgen05.st (32x32b: lane = row, column = k)
//   T3  TS  the whole MLP of a 128-cell tile: layer 1 as 3 TF32 passes (hi*hi + hi*lo + lo*hi), bias + ReLU + split in
//           registers, H back to TMEM, layer 2 (N = C) the same way; compared with an fp64 host reference; then the
//           same tile repeated `reps` times on every SM for a cycles-per-tile figure.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/umma_probe scratch/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (r, k) of an [R x K] fp32 operand in the no-swizzle K-major canonical layout, in floats:
// core matrix = 8 rows x 16 bytes, 128 contiguous bytes; core matrices adjacent along K (LBO = 128 B);
// 8-row groups strided by SBO = (K/4) * 128 B
__host__ __device__ inline int canon_idx(int r, int k, int K) { return (r >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
  return d;                        // base offset 0, layout type 0 = no swizzle
}
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ uint32_t tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - __uint_as_float(hi));
}

constexpr int M = 128, K1 = 96, N1 = 128;

// ---- T1 / T2: one TF32 pass -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_single(const float* __restrict__ A /*[128][96] row-major*/,
                                                const float* __restrict__ Bc /*canonical [128 x 96]*/, float* __restrict__ D, int ts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + M * K1;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N1 * K1; i += 128) sB[i] = Bc[i];
  if (!ts) for (int i = tid; i < M * K1; i += 128) { int r = i / K1, k = i % K1; sA[canon_idx(r, k, K1)] = A[i]; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t tA = tb + 256, tD = tb;          // D: columns [0,128); A: columns [256, 352)
  if (ts) {
    for (int c0 = 0; c0 < K1; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(A[tid * K1 + c0 + j]);
      tmem_st16(tA + lane_base + c0, v);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(M, N1);
    const uint32_t sbo = (K1 / 4) * 128;
    for (int ks = 0; ks < K1 / 8; ++ks) {
      const uint64_t bd = make_desc(smem_u32(sB) + ks * 256, 128, sbo);
      if (ts) mma_ts(tD, tA + ks * 8, bd, idesc, ks > 0);
      else mma_ss(tD, make_desc(smem_u32(sA) + ks * 256, 128, sbo), bd, idesc, ks > 0);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N1; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tD + lane_base + c0, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * N1 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

// ---- T3: the MLP of a 128-cell tile, 3xTF32, activations in TMEM -----------------------------------------------------
// weights: W1 hi/lo canonical [128 x 96], W2 hi/lo canonical [C x 128]
template <int C>
__global__ void __launch_bounds__(128) k_mlp(const float* __restrict__ Y /*[128][96]*/, const float* __restrict__ W1c /*[2][128*96]*/,
                                             const float* __restrict__ b1, const float* __restrict__ W2c /*[2][C*128]*/,
                                             float* __restrict__ out /*[grid][128][C]*/, int reps, long long* __restrict__ cycles, int split_acc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sW1 = reinterpret_cast<float*>(smem);       // hi, lo
  float* sW2 = sW1 + 2 * N1 * K1;                    // hi, lo
  float* sb1 = sW2 + 2 * C * N1;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * N1 * K1; i += 128) sW1[i] = W1c[i];
  for (int i = tid; i < 2 * C * N1; i += 128) sW2[i] = W2c[i];
  for (int i = tid; i < N1; i += 128) sb1[i] = b1[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // columns: Y hi [0,96) lo [96,192); D1 -> H hi [192,320); cross terms -> H lo [320,448); D2 [448,448+C); D2 cross [448+C,448+2C)
  const uint32_t tYh = tb, tYl = tb + 96, tD1 = tb + 192, tHl = tb + 320, tD2 = tb + 448, tD2s = tb + 448 + C;
  const uint32_t idesc1 = make_idesc_tf32(M, N1), idesc2 = make_idesc_tf32(M, C);
  uint32_t parity = 0;
  float yreg[K1];
#pragma unroll
  for (int k = 0; k < K1; ++k) yreg[k] = Y[tid * K1 + k];
  float dx[C];
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    // ---- Y -> TMEM (hi, lo) ----
#pragma unroll
    for (int c0 = 0; c0 < K1; c0 += 16) {
      uint32_t vh[16], vl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_tf32(yreg[c0 + j], vh[j], vl[j]);
      tmem_st16(tYh + lane_base + c0, vh);
      tmem_st16(tYl + lane_base + c0, vl);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sbo = (K1 / 4) * 128;
      const uint32_t w1h = smem_u32(sW1), w1l = smem_u32(sW1 + N1 * K1);
      for (int ks = 0; ks < K1 / 8; ++ks) {
        mma_ts(tD1, tYh + ks * 8, make_desc(w1h + ks * 256, 128, sbo), idesc1, ks > 0);
        if (split_acc) {
          mma_ts(tHl, tYh + ks * 8, make_desc(w1l + ks * 256, 128, sbo), idesc1, ks > 0);
          mma_ts(tHl, tYl + ks * 8, make_desc(w1h + ks * 256, 128, sbo), idesc1, 1);
        } else {
          mma_ts(tD1, tYh + ks * 8, make_desc(w1l + ks * 256, 128, sbo), idesc1, 1);
          mma_ts(tD1, tYl + ks * 8, make_desc(w1h + ks * 256, 128, sbo), idesc1, 1);
        }
      }
      mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), parity); parity ^= 1;
    tc_fence_after();
    // ---- epilogue 1: bias + ReLU + split, H back to TMEM (hi in place of D1) ----
#pragma unroll
    for (int c0 = 0; c0 < N1; c0 += 16) {
      uint32_t v[16], vl[16];
      tmem_ld16(tD1 + lane_base + c0, v);
      if (split_acc) tmem_ld16(tHl + lane_base + c0, vl);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = fmaxf((__uint_as_float(v[j]) + (split_acc ? __uint_as_float(vl[j]) : 0.f)) + sb1[c0 + j], 0.f);
        split_tf32(h, v[j], vl[j]);
      }
      tmem_st16(tD1 + lane_base + c0, v);
      tmem_st16(tHl + lane_base + c0, vl);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sbo = (N1 / 4) * 128;
      const uint32_t w2h = smem_u32(sW2), w2l = smem_u32(sW2 + C * N1);
      for (int ks = 0; ks < N1 / 8; ++ks) {
        mma_ts(tD2, tD1 + ks * 8, make_desc(w2h + ks * 256, 128, sbo), idesc2, ks > 0);
        if (split_acc) {
          mma_ts(tD2s, tD1 + ks * 8, make_desc(w2l + ks * 256, 128, sbo), idesc2, ks > 0);
          mma_ts(tD2s, tHl + ks * 8, make_desc(w2h + ks * 256, 128, sbo), idesc2, 1);
        } else {
          mma_ts(tD2, tD1 + ks * 8, make_desc(w2l + ks * 256, 128, sbo), idesc2, 1);
          mma_ts(tD2, tHl + ks * 8, make_desc(w2h + ks * 256, 128, sbo), idesc2, 1);
        }
      }
      mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), parity); parity ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16) {
      uint32_t v[16], vs[16];
      tmem_ld16(tD2 + lane_base + c0, v);
      if (split_acc) tmem_ld16(tD2s + lane_base + c0, vs);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) dx[c0 + j] = __uint_as_float(v[j]) + (split_acc ? __uint_as_float(vs[j]) : 0.f);
    }
    // keep the loop honest: the next tile's input depends on this tile's output (by a negligible amount)
    yreg[0] += dx[0] * 1e-30f;
  }
  const long long t1 = clock64();
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
#pragma unroll
  for (int c = 0; c < C; ++c) out[((size_t)blockIdx.x * M + tid) * C + c] = dx[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }
static float tf32_rna_host(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

template <int C>
static void run_mlp(const std::vector<float>& Y, const std::vector<float>& W1, const std::vector<float>& b1, int nsm, int split_acc) {
  std::vector<float> W2((size_t)C * N1);
  srand(7 + C);
  for (auto& v : W2) v = ((rand() / (float)RAND_MAX) - 0.5f) * 0.2f;
  std::vector<float> W1c(2 * N1 * K1), W2c(2 * C * N1);
  for (int n = 0; n < N1; ++n) for (int k = 0; k < K1; ++k) {
    const float w = W1[n * K1 + k], hi = tf32_rna_host(w), lo = tf32_rna_host(w - hi);
    W1c[canon_idx(n, k, K1)] = hi; W1c[N1 * K1 + canon_idx(n, k, K1)] = lo;
  }
  for (int c = 0; c < C; ++c) for (int k = 0; k < N1; ++k) {
    const float w = W2[c * N1 + k], hi = tf32_rna_host(w), lo = tf32_rna_host(w - hi);
    W2c[canon_idx(c, k, N1)] = hi; W2c[C * N1 + canon_idx(c, k, N1)] = lo;
  }
  float *dY, *dW1c, *db1, *dW2c, *dout; long long* dcyc;
  CK(cudaMalloc(&dY, Y.size() * 4)); CK(cudaMalloc(&dW1c, W1c.size() * 4)); CK(cudaMalloc(&db1, b1.size() * 4));
  CK(cudaMalloc(&dW2c, W2c.size() * 4)); CK(cudaMalloc(&dout, (size_t)nsm * M * C * 4)); CK(cudaMalloc(&dcyc, nsm * 8));
  CK(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW1c, W1c.data(), W1c.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW2c, W2c.data(), W2c.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(2 * N1 * K1 + 2 * C * N1 + N1) * 4;
  CK(cudaFuncSetAttribute(k_mlp<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mlp<C><<<1, 128, smem>>>(dY, dW1c, db1, dW2c, dout, 1, nullptr, split_acc);
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)M * C);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double emax = 0, e32max = 0, refmax = 0;
  for (int r = 0; r < M; ++r) {
    double h[N1]; float hf[N1];
    for (int n = 0; n < N1; ++n) {
      double s = b1[n]; float sf = b1[n];
      for (int k = 0; k < K1; ++k) { s += (double)W1[n * K1 + k] * Y[r * K1 + k]; sf = fmaf(W1[n * K1 + k], Y[r * K1 + k], sf); }
      h[n] = s > 0 ? s : 0; hf[n] = sf > 0 ? sf : 0;
    }
    for (int c = 0; c < C; ++c) {
      double s = 0; float sf = 0;
      for (int n = 0; n < N1; ++n) { s += (double)W2[c * N1 + n] * h[n]; sf = fmaf(W2[c * N1 + n], hf[n], sf); }
      emax = fmax(emax, fabs(out[r * C + c] - s)); e32max = fmax(e32max, fabs((double)sf - s)); refmax = fmax(refmax, fabs(s));
    }
  }
  printf("T3 C=%d split_acc=%d  3xTF32 MLP: max|err| vs fp64 = %.3e (fp32 FFMA chain: %.3e), max|ref| = %.3e -> rel %.3e\n", C, split_acc, emax, e32max, refmax, emax / refmax);
  // throughput: every SM, reps tiles
  const int reps = 2000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_mlp<C><<<nsm, 128, smem>>>(dY, dW1c, db1, dW2c, dout, 50, dcyc, split_acc);
  CK(cudaEventRecord(e0));
  k_mlp<C><<<nsm, 128, smem>>>(dY, dW1c, db1, dW2c, dout, reps, dcyc, split_acc);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(nsm); CK(cudaMemcpy(cyc.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  const double cells = (double)nsm * reps * M;
  const double flop = cells * 2.0 * (K1 * N1 + N1 * C);
  printf("T3 C=%d  %d SMs x %d tiles (1 warpgroup per SM, serial): %.3f ms, %.1f cycles/tile, %.3e cells/s, %.1f TFLOP/s fp32-equivalent\n",
         C, nsm, reps, ms, (double)cyc[0] / reps, cells / (ms * 1e-3), flop / (ms * 1e-3) / 1e12);
  cudaFree(dY); cudaFree(dW1c); cudaFree(db1); cudaFree(dW2c); cudaFree(dout); cudaFree(dcyc);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  std::vector<float> A(M * K1), B(N1 * K1), b1(N1);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : B) v = ((rand() / (float)RAND_MAX) - 0.5f) * 0.3f;
  for (auto& v : b1) v = ((rand() / (float)RAND_MAX) - 0.5f) * 0.1f;
  std::vector<float> Bc(N1 * K1);
  for (int n = 0; n < N1; ++n) for (int k = 0; k < K1; ++k) Bc[canon_idx(n, k, K1)] = B[n * K1 + k];
  float *dA, *dBc, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBc, Bc.size() * 4)); CK(cudaMalloc(&dD, M * N1 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBc, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(M * K1 + N1 * K1) * 4;
  CK(cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int ts = 0; ts < 2; ++ts) {
    CK(cudaMemset(dD, 0, M * N1 * 4));
    k_single<<<1, 128, smem>>>(dA, dBc, dD, ts);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N1);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double e_full = 0, e_trunc = 0, e_rna = 0, rmax = 0;
    for (int r = 0; r < M; ++r) for (int n = 0; n < N1; ++n) {
      double s = 0, st = 0, sr = 0;
      for (int k = 0; k < K1; ++k) {
        s += (double)A[r * K1 + k] * B[n * K1 + k];
        st += (double)tf32_trunc(A[r * K1 + k]) * tf32_trunc(B[n * K1 + k]);
        sr += (double)tf32_rna_host(A[r * K1 + k]) * tf32_rna_host(B[n * K1 + k]);
      }
      const double d = D[r * N1 + n];
      e_full = fmax(e_full, fabs(d - s)); e_trunc = fmax(e_trunc, fabs(d - st)); e_rna = fmax(e_rna, fabs(d - sr)); rmax = fmax(rmax, fabs(s));
    }
    printf("T%d %s  max|D - exact| = %.3e   max|D - tf32(trunc) ref| = %.3e   max|D - tf32(rna) ref| = %.3e   (max|ref| %.3e)\n",
           ts + 1, ts ? "TS" : "SS", e_full, e_trunc, e_rna, rmax);
  }
  for (int sa = 0; sa < 2; ++sa) {
    run_mlp<32>(A, B, b1, prop.multiProcessorCount, sa);
    run_mlp<16>(A, B, b1, prop.multiProcessorCount, sa);
  }
  printf("done\n");
  return 0;
}
