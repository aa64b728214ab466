// Probe: do all SMs run the same SM-local work in the same number of cycles / nanoseconds?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(long long* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.000001f;
  long long c0 = clock64();
  unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  for (int i = 0; i < iters; ++i) a = fmaf(a, b, 1e-7f);
  long long c1 = clock64();
  unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) { out[blockIdx.x * 4] = smid; out[blockIdx.x * 4 + 1] = c1 - c0; out[blockIdx.x * 4 + 2] = (long long)(g1 - g0); out[blockIdx.x * 4 + 3] = (long long)a; }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  long long h[148 * 4];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 3; ++rep) {
    probe<<<148, 32, 200 * 1024>>>(d, 200000);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long cmin = 1ll << 60, cmax = 0, gmin = 1ll << 60, gmax = 0;
  for (int i = 0; i < 148; ++i) { if (h[i*4+1] < cmin) cmin = h[i*4+1]; if (h[i*4+1] > cmax) cmax = h[i*4+1]; if (h[i*4+2] < gmin) gmin = h[i*4+2]; if (h[i*4+2] > gmax) gmax = h[i*4+2]; }
  printf("cycles min %lld max %lld ; ns min %lld max %lld\n", cmin, cmax, gmin, gmax);
  for (int i = 0; i < 148; ++i) printf("%lld:%lld:%lld ", h[i*4], h[i*4+1] / 1000, h[i*4+2] / 1000);
  printf("\n");
  return 0;
}
