// DSMEM all-gather microbenchmark: every CTA of an 8-cluster pushes N cells x 64 B to its 7 peers and waits for the
// 7 x N x 64 B that the peers push to it (mbarrier transaction count) -- the exchange of k_rep_fwd's step S3.
//   mode 0: st.async.f32,    lane = (cell, channel): 32 lanes x 4 B per instruction    (what the kernel did)
//   mode 1: st.async.v4.f32, lane = (cell, channel quad): 32 lanes x 16 B = 8 cells per instruction
//   mode 2: st.async.v2.f32, lane = (cell, channel pair): 32 lanes x 8 B = 4 cells per instruction
//   mode 3: cp.async.bulk shared::cta -> shared::cluster, one 64 B line per (cell, peer), issued by lane = peer
// Reports cycles per round (issue of the pushes + wait for the incoming bytes), CTA 0, median-ish mean over rounds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
constexpr int NC = 8, kCells = 1600;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok = 0;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory"); } while (!ok);
}
__global__ void __launch_bounds__(512, 1) k(int mode, int n_my, int rounds, int warps_used, long long* out) {
  extern __shared__ __align__(128) float sX[];       // [kCells][16]
  __shared__ __align__(8) unsigned long long mbar;
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank(), tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kCells * 16; i += 512) sX[i] = (float)i;
  const uint32_t mb = smem_u32(&mbar);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  cl.sync();
  long long acc = 0, acc_issue = 0;
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(7 * n_my * 64)) : "memory");
    const int base = rank * n_my;                       // my cells: [base, base + n_my)
    if (mode == 0) {
      for (int s = warp * 2 + (lane >> 4); s < n_my; s += warps_used * 2) {
        if (warp >= warps_used) break;
        const uint32_t la = smem_u32(&sX[(base + s) * 16 + (lane & 15)]);
        const float v = (float)(r + s);
#pragma unroll
        for (int p = 1; p < NC; ++p) {
          const int peer = (rank + p) & (NC - 1);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(mapa(la, peer)), "f"(v), "r"(mapa(mb, peer)) : "memory");
        }
      }
    } else if (mode == 1) {
      for (int s = warp * 8 + (lane >> 2); s < n_my; s += warps_used * 8) {
        if (warp >= warps_used) break;
        const uint32_t la = smem_u32(&sX[(base + s) * 16 + 4 * (lane & 3)]);
        const float v = (float)(r + s);
#pragma unroll
        for (int p = 1; p < NC; ++p) {
          const int peer = (rank + p) & (NC - 1);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %1, %1, %1}, [%2];" ::"r"(mapa(la, peer)), "f"(v), "r"(mapa(mb, peer)) : "memory");
        }
      }
    } else if (mode == 2) {
      for (int s = warp * 4 + (lane >> 3); s < n_my; s += warps_used * 4) {
        if (warp >= warps_used) break;
        const uint32_t la = smem_u32(&sX[(base + s) * 16 + 2 * (lane & 7)]);
        const float v = (float)(r + s);
#pragma unroll
        for (int p = 1; p < NC; ++p) {
          const int peer = (rank + p) & (NC - 1);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %1}, [%2];" ::"r"(mapa(la, peer)), "f"(v), "r"(mapa(mb, peer)) : "memory");
        }
      }
    } else {
      // local lines are already in smem (generic-proxy writes happened before the __syncthreads): make them visible to the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int s = warp * 4 + (lane >> 3); s < n_my; s += warps_used * 4) {
        if (warp >= warps_used) break;
        const int p = lane & 7;
        if (p >= 1) {
          const int peer = (rank + p) & (NC - 1);
          const uint32_t la = smem_u32(&sX[(base + s) * 16]);
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(mapa(la, peer)), "r"(la), "r"(mapa(mb, peer)) : "memory");
        }
      }
    }
    const long long t1 = clock64();
    mbar_wait(mb, (uint32_t)(r & 1));
    __syncthreads();
    const long long t2 = clock64();
    if (r >= 2) { acc += t2 - t0; acc_issue += t1 - t0; }
  }
  if (blockIdx.x == 0 && tid == 0) { out[0] = acc / (rounds - 2); out[1] = acc_issue / (rounds - 2); }
  cl.sync();
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kCells * 64);
  const char* names[4] = {"st.async.f32 (4 B/lane)", "st.async.v4.f32 (16 B/lane)", "st.async.v2.f32 (8 B/lane)", "cp.async.bulk 64 B/line"};
  for (int nclusters : {1, 8}) for (int n_my : {8, 20, 32, 64, 136}) for (int mode = 0; mode < 4; ++mode) for (int wu : {16, 4}) {
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(NC * nclusters); q.blockDim = dim3(512); q.dynamicSmemBytes = kCells * 64;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = NC; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&q, k, mode, n_my, 50, wu, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("clusters %d n_my %3d warps %2d  %-28s: %6lld cycles per round (issue %5lld)  %5.1f B/clk in  [%s]\n", nclusters, n_my, wu, names[mode], h[0], h[1],
           h[0] ? 7.0 * n_my * 64 / (double)h[0] : 0.0, cudaGetErrorString(e));
  }
  return 0;
}
