// Floor of a dependent kernel chain inside a CUDA graph: empty kernels, with/without cluster dims and PDL.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_empty(float* p) { if (p == nullptr && threadIdx.x == 1000) p[0] = 1; }
__global__ void k_pdl(float* p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p == nullptr && threadIdx.x == 1000) p[0] = 1;
}
static float run(int mode, int grid, int cluster) {
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < 200; ++i) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid / cluster, cluster); cfg.blockDim = dim3(128); cfg.stream = s;
    cudaLaunchAttribute at[2]; int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = cluster; at[na].val.clusterDim.z = 1; na++;
    if (mode == 1) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    cfg.attrs = at; cfg.numAttrs = na;
    if (mode == 1) cudaLaunchKernelEx(&cfg, k_pdl, (float*)nullptr); else cudaLaunchKernelEx(&cfg, k_empty, (float*)nullptr);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1000.f / (10 * 200);
}
int main() {
  for (int grid : {128, 512, 1024})
    for (int cl : {1, 2, 8})
      printf("grid=%4d cluster=%d  plain %.2f us/kernel   PDL %.2f us/kernel\n", grid, cl, run(0, grid, cl), run(1, grid, cl));
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
