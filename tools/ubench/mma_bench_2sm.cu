// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair, K = 16, SS mode)
// vs N and issue pattern -- the 2-SM counterpart of mma_bench.cu.  One cluster of two CTAs; the leader's
// thread 0 issues `iters` MMAs back to back on zeroed resident smem operands, then commits (multicast).
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o mma_bench_2sm mma_bench_2sm.cu && ./mma_bench_2sm
#include <cstdio>
#include <cuda_runtime.h>
#include "../../frame_interpolation_b200/csrc/film_tc_ptx.cuh"
using namespace film::tc;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_bench(int mode, int n1, int n2, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc_2sm(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (leader && threadIdx.x == 0) {
    auto idesc_of = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); };
    const uint32_t id1 = idesc_of(n1), id2 = idesc_of(n2);
    const uint32_t a0 = base, b0 = base + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // mode 0: same shape n1, operands walk through 4 k-steps and 4 A tiles
      // mode 1: alternate n1 / n2 (the fused 2-instruction pattern), same accumulator
      // mode 2: same shape, alternate between two accumulators
      const uint32_t koff = (uint32_t)(i & 3) * 32;
      const uint32_t aoff = (uint32_t)((i >> 2) & 3) * 16384;
      const uint64_t ad = make_desc(a0 + aoff + koff), bd = make_desc(b0 + koff);
      const uint32_t id = (mode == 1 && (i & 1)) ? id2 : id1;
      const uint32_t d = (mode == 2 && (i & 1)) ? tm + 256 : tm;
      umma_2sm(d, ad, bd, id, i > 1 ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit_2sm_mc(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  if (!leader && threadIdx.x == 0) mbar_wait(smem_u32(&bar), 0);  // the multicast commit also lands here
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_2sm(tm, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  struct { int mode, n1, n2; const char* name; } cases[] = {
      {0, 32, 0, "N=32"}, {0, 64, 0, "N=64"}, {0, 128, 0, "N=128"}, {0, 256, 0, "N=256"},
      {1, 64, 32, "alternate N=64/N=32  (BN=32 fused)"}, {1, 128, 64, "alternate N=128/N=64 (BN=64 fused)"},
      {1, 256, 128, "alternate N=256/N=128"}, {2, 64, 0, "N=64 two accumulators"}, {2, 128, 0, "N=128 two accumulators"}};
  for (auto& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      k_bench<<<2, 128, 200 * 1024>>>(c.mode, c.n1, c.n2, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=256 (cta_group::2) %-36s issue %.1f cyc/MMA   complete %.1f cyc/MMA\n", c.name, (double)h[0] / iters, (double)h[1] / iters);
  }
  return 0;
}
