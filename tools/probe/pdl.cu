// Probe: cost of a kernel->kernel boundary inside a CUDA graph, with and without programmatic
// dependent launch (griddepcontrol.wait / launch_dependents at the top of every kernel).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl pdl.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p, int work, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
  extern __shared__ float sm[];
  float v = p[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  sm[threadIdx.x] = v;
  p[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}
static float run(int n, int work, int pdl, size_t smem, int grid) {
  float* p; cudaMalloc(&p, grid * 256 * 4); cudaMemset(p, 0, grid * 256 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, p, work, pdl);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -1; }
  }
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture: %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(a, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  e = cudaGetLastError(); if (e != cudaSuccess) printf("err: %s\n", cudaGetErrorString(e));
  cudaFree(p);
  return ms / 10 / n * 1000.f;   // us per kernel
}
int main() {
  for (int work : {0, 2000, 20000})
    for (size_t smem : {(size_t)1024, (size_t)200 * 1024})
      for (int grid : {148, 148 * 4}) {
        if (smem > 100000 && grid > 148) continue;
        float t0 = run(600, work, 0, smem, grid), t1 = run(600, work, 1, smem, grid);
        printf("work %6d smem %6zu grid %4d : plain %.2f us/kernel, pdl %.2f us/kernel, saved %.2f us\n", work, smem, grid, t0, t1, t0 - t1);
      }
  return 0;
}
