// micro-benchmark: how fast can 86 CTAs each write a 128 x BN fp32 tile of a row-major matrix with different store patterns?
#include <cstdio>
#include <cuda_runtime.h>
template <int PATTERN>
__global__ void k(float* D, long long ldd, int tiles_n, int BN) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  if (PATTERN == 0) {          // 8 warps: quarter q rows, half h chunks; per instr 4 rows x 128 B
    int q = warp & 3, h = warp >> 2;
    for (int ch = h; ch < BN / 32; ch += 2)
      for (int i = 0; i < 8; i++) {
        int r = tm * 128 + q * 32 + 4 * i + (lane >> 3), c = tn * BN + ch * 32 + (lane & 7) * 4;
        *reinterpret_cast<float4*>(D + (size_t)r * ldd + c) = v;
      }
  } else if (PATTERN == 1) {   // each warp instruction writes 512 contiguous bytes of one row (needs BN >= 128)
    for (int r = warp; r < 128; r += 8)
      for (int c = lane * 4; c < BN; c += 128)
        *reinterpret_cast<float4*>(D + (size_t)(tm * 128 + r) * ldd + tn * BN + c) = v;
  } else {                     // one thread per row, 128 B per thread per chunk (the old pattern)
    int r = tm * 128 + (threadIdx.x & 127);
    for (int ch = threadIdx.x >> 7; ch < BN / 32; ch += 2)
      for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(D + (size_t)r * ldd + tn * BN + ch * 32 + j * 4) = v;
  }
}
int main() {
  const int M = 256, N = 5408, BN = 128;
  float* D; cudaMalloc(&D, (size_t)M * N * 4);
  int tiles_n = (N + BN - 1) / BN;  // last tile partially OOB: allocate extra
  cudaFree(D); cudaMalloc(&D, (size_t)M * (tiles_n * BN) * 4);
  long long ldd = tiles_n * BN;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pat = 0; pat < 3; pat++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      for (int it = 0; it < 20; it++) {
        if (pat == 0) k<0><<<2 * tiles_n, 256>>>(D, ldd, tiles_n, BN);
        if (pat == 1) k<1><<<2 * tiles_n, 256>>>(D, ldd, tiles_n, BN);
        if (pat == 2) k<2><<<2 * tiles_n, 256>>>(D, ldd, tiles_n, BN);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("pattern %d: %.2f us per launch (%d CTAs, %.1f MB)\n", pat, ms / 20 * 1e3, 2 * tiles_n, M * ldd * 4 / 1e6);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
