// Micro-benchmark: cost of the boundary between two consecutively launched CUDA graphs (same stream), with and without a
// cudaStreamWaitEvent on another stream's event in between (the train loop's slot wait).
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void link(const float* __restrict__ in, float* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] + 1.f;
}
__global__ void tiny(float* p) { p[0] += 1.f; }
int build(cudaStream_t s, float* a, float* b, int n, int links, cudaGraphExec_t* ge) {
  cudaGraph_t g;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  for (int l = 0; l < links; l++) link<<<148, 256, 0, s>>>((l & 1) ? b : a, (l & 1) ? a : b, n);
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(ge, g, 0));
  return 0;
}
int main() {
  const int n = 1 << 16, reps = 400;
  float *a, *b, *c;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, 64));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(c, 0, 64));
  cudaStream_t s, s2; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e0, e1, ev, evf; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); cudaEventCreateWithFlags(&evf, cudaEventDisableTiming);
  cudaGraphExec_t g10a, g10b, g20;
  if (build(s, a, b, n, 10, &g10a) || build(s, a, b, n, 10, &g10b) || build(s, a, b, n, 20, &g20)) return 1;
  float ms;
  for (int mode = 0; mode < 4; mode++) {
    for (int w = 0; w < 2; w++) {
      if (w == 1) CK(cudaEventRecord(e0, s));
      for (int i = 0; i < reps; i++) {
        if (mode == 0) { CK(cudaGraphLaunch(g20, s)); }
        else if (mode == 1) { CK(cudaGraphLaunch(g10a, s)); CK(cudaGraphLaunch(g10b, s)); }
        else if (mode == 2) {                      // side-stream kernel forked from here, waited for before the next graph (prefetch)
          CK(cudaGraphLaunch(g10a, s));
          CK(cudaEventRecord(evf, s)); CK(cudaStreamWaitEvent(s2, evf, 0)); tiny<<<1, 1, 0, s2>>>(c); CK(cudaEventRecord(ev, s2));
          CK(cudaStreamWaitEvent(s, ev, 0));
          CK(cudaGraphLaunch(g10b, s));
        } else {                                   // like the train loop: the side work was forked ONE graph earlier
          CK(cudaEventRecord(evf, s)); CK(cudaStreamWaitEvent(s2, evf, 0)); tiny<<<1, 1, 0, s2>>>(c); cudaEvent_t evn = ev; CK(cudaEventRecord(evn, s2));
          CK(cudaGraphLaunch(g10a, s));
          CK(cudaStreamWaitEvent(s, evn, 0));
          CK(cudaGraphLaunch(g10b, s));
        }
      }
      if (w == 1) { CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); cudaEventElapsedTime(&ms, e0, e1); }
      CK(cudaDeviceSynchronize());
    }
    const char* names[] = {"one 20-kernel graph", "two 10-kernel graphs", "two graphs + fork/join of a side-stream kernel between them",
                           "two graphs, side kernel forked one graph earlier (train loop)"};
    printf("%-66s %.2f us per 20 kernels\n", names[mode], ms * 1e3 / reps);
  }
  return 0;
}
