// What does a tcgen05.commit cost the issuing thread, and how long until its mbarrier arrival is visible?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/commit_bench scripts/exp/commit_bench.cu && /tmp/commit_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// mode 0: commit + wait for its arrival (round trip); 1: commits to 8 barriers round robin, waiting 7 behind (issue cost when pipelined);
// 2: plain mbarrier.arrive + wait (no tensor unit); 3: two threads ping-pong through two mbarriers with plain arrives (wake-up latency)
__global__ void bench(int mode, int iters, long long* out) {
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tslot;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const uint32_t b0 = smem_u32(bars);
  long long t0 = 0, t1 = 0;
  if (mode == 3) {
    if (threadIdx.x == 0 || threadIdx.x == 32) {
      const int me = threadIdx.x == 0 ? 0 : 1;
      t0 = clock64();
      uint32_t ph = 0;
      for (int i = 0; i < iters; i++) {
        if (me == 0) { arrive(b0); while (!try_wait(b0 + 8, ph)) {} }
        else { while (!try_wait(b0, ph)) {} arrive(b0 + 8); }
        ph ^= 1;
      }
      t1 = clock64();
    }
  } else if (threadIdx.x == 0) {
    t0 = clock64();
    if (mode == 0) {
      uint32_t ph = 0;
      for (int i = 0; i < iters; i++) { commit(b0); while (!try_wait(b0, ph)) {} ph ^= 1; }
    } else if (mode == 1) {
      uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < iters; i++) {
        const int s = i & 7;
        if (i >= 8) { while (!try_wait(b0 + s * 8, ph[s])) {} ph[s] ^= 1; }
        commit(b0 + s * 8);
      }
    } else if (mode == 2) {
      uint32_t ph = 0;
      for (int i = 0; i < iters; i++) { arrive(b0); while (!try_wait(b0, ph)) {} ph ^= 1; }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tslot) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  const char* names[4] = {"commit + wait (round trip)", "commit, 8 barriers round robin (pipelined issue)", "mbarrier.arrive + wait", "two-thread ping-pong via mbarriers (plain arrives)"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; mode++) {
      const int iters = 2000;
      bench<<<grid, 64>>>(mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("grid %3d  %-52s %8.1f cycles per iteration (%s)\n", grid, names[mode], (double)h[0] / iters, cudaGetErrorString(e));
    }
  return 0;
}
