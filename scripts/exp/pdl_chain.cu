// Micro-benchmark: cost per link of a chain of short dependent kernels inside a CUDA graph, with plain edges vs programmatic
// dependent launch (griddepcontrol).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <bool PDL, bool EARLY>
__global__ void link(const float* __restrict__ in, float* __restrict__ out, int n) {
  if (PDL) {
    if (EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] + 1.f;
}

template <bool PDL, bool EARLY>
int run(const char* label, int grid, int n, int links) {
  float *a, *b;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4));
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  for (int l = 0; l < links; l++) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at; memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization; at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = PDL ? 1 : 0;
    const float* in = (l & 1) ? b : a; float* out = (l & 1) ? a : b;
    CK(cudaLaunchKernelEx(&cfg, link<PDL, EARLY>, in, out, n));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 20; i++) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  const int reps = 200;
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < reps; i++) CK(cudaGraphLaunch(ge, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  float h; CK(cudaMemcpy(&h, (links & 1) ? b : a, 4, cudaMemcpyDeviceToHost));
  printf("%-34s grid %4d n %8d: %.2f us per link   (check %.0f)\n", label, grid, n, ms * 1e3 / reps / links, h);
  cudaFree(a); cudaFree(b);
  return 0;
}

int main() {
  const int links = 20;
  for (int n : {1 << 10, 1 << 20, 1 << 22}) {
    for (int grid : {32, 148, 592}) {
      if (run<false, false>("plain edges", grid, n, links)) return 1;
      if (run<true, false>("PDL wait->launch_dependents", grid, n, links)) return 1;
      if (run<true, true>("PDL launch_dependents->wait", grid, n, links)) return 1;
    }
  }
  return 0;
}
