// Probe: cost of an mbarrier hand-off between warps of one CTA (sm_100a).
//  mode 0: warp A arrives on b0, warp B waits b0 then arrives on b1, A waits b1 (ping-pong, 1 token)
//  mode 1: same, but NW extra warps also wait on b0 every round (like 8 producer warps on `empty`)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void binit(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void barrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(b)) : "memory"); }
template <int TRY>
__device__ __forceinline__ void bwait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) {
    if (TRY)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(su32(b)), "r"(par) : "memory");
  }
}
template <int TRY, int LANE0>
__device__ __forceinline__ void wwait(uint64_t* b, uint32_t par) {
  if (LANE0) { if ((threadIdx.x & 31) == 0) bwait<TRY>(b, par); __syncwarp(); }
  else bwait<TRY>(b, par);
}
template <int TRY, int LANE0>
__global__ void pp(long long* out, int rounds, int nwait) {
  __shared__ uint64_t b0, b1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { binit(&b0, 1); binit(&b1, 1 + nwait); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    for (int r = 0; r < rounds; ++r) {
      if (lane == 0) barrive(&b0);
      __syncwarp();
      wwait<TRY, LANE0>(&b1, r & 1);
    }
  } else if (warp <= 1 + nwait) {    // warp 1 = responder; warps 2.. = extra waiters that also arrive on b1
    for (int r = 0; r < rounds; ++r) {
      wwait<TRY, LANE0>(&b0, r & 1);
      if (lane == 0) barrive(&b1);
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  long long h[148];
  const int R = 2000;
#define RUN(TRY, L0, NW, name) do { pp<TRY, L0><<<148, 32 * (2 + NW)>>>(d, R, NW); cudaDeviceSynchronize(); pp<TRY, L0><<<148, 32 * (2 + NW)>>>(d, R, NW); cudaDeviceSynchronize(); \
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); long long mn = 1ll << 60, mx = 0; for (int i = 0; i < 148; ++i) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; } \
  printf("%-40s round trip cycles: min %.0f max %.0f\n", name, (double)mn / R, (double)mx / R); } while (0)
  RUN(1, 1, 0, "try_wait lane0, 1 waiter");
  RUN(0, 1, 0, "test_wait lane0, 1 waiter");
  RUN(1, 0, 0, "try_wait all lanes, 1 waiter");
  RUN(1, 1, 7, "try_wait lane0, 8 waiters");
  RUN(0, 1, 7, "test_wait lane0, 8 waiters");
  RUN(1, 0, 7, "try_wait all lanes, 8 waiters");
  RUN(1, 1, 11, "try_wait lane0, 12 waiters");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
