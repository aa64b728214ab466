// Probe: issue/execution rate of tcgen05.mma (kind::f16, cta_group::1, SS) as a function of the operand
// layout and shape -- what bounds the narrow-layer conv kernels (M = 128 positions, N = Cout = 16..48,
// K = 16 per instruction, A rows of only Cin*2 = 32 bytes).
//   per config: one warp issues G groups of 9 accumulating MMAs (the 9 taps of a 3x3 conv) into
//   alternating TMEM accumulators, commits, waits; cycles per MMA = (t1 - t0) / (9 G).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
struct Cfg { int layout_a, layout_b, a_mn, b_mn, M, N, pitch_a, pitch_b, a_step, b_step, nacc; int rr; int sbo_a, lbo_a; };   // rr: round-robin accumulators per MMA; sbo_a/lbo_a: explicit A strides (un-swizzled)
// layout: 0 none (core matrices), 6 SW32, 4 SW64, 2 SW128
__global__ void probe(Cfg c, int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    // descriptors
    auto mk = [&](int layout, int mn, int pitch, uint32_t addr) -> uint64_t {
      uint64_t sbo, lbo;
      if (layout == 0) { sbo = 128; lbo = 128 * 16; }            // K-major core matrices: 16 M-groups then next K chunk
      else if (!mn) { sbo = 8 * pitch; lbo = 16; }
      else { sbo = 8 * pitch; lbo = 64 * pitch; }                   // MN-major: next 8 K rows / next MN group
      return ((uint64_t)layout << 61) | (1ull << 46) | ((sbo >> 4) << 32) | ((lbo >> 4) << 16) | (uint64_t)((addr >> 4) & 0x3fff);
    };
    const uint32_t a0 = su32(smem), b0 = su32(smem + 96 * 1024);
    const uint32_t idesc = (1u << 4) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                           ((uint32_t)(c.M >> 4) << 24);
    uint64_t ad = mk(c.layout_a, c.a_mn, c.pitch_a, a0);
    const uint64_t bd = mk(c.layout_b, c.b_mn, c.pitch_b, b0);
    if (c.sbo_a) ad = (1ull << 46) | ((uint64_t)(c.sbo_a >> 4) << 32) | ((uint64_t)(c.lbo_a >> 4) << 16) | (uint64_t)((a0 >> 4) & 0x3fff);
    long long t0 = clock64();
    if (!c.rr) {
      for (int g = 0; g < groups; ++g) {
        const uint32_t d = tmem + (uint32_t)((g % c.nacc) * c.N);
#pragma unroll
        for (int t = 0; t < 9; ++t)
          umma(d, ad + (uint64_t)((t * c.a_step) >> 4), bd + (uint64_t)((t * c.b_step) >> 4), idesc, t ? 1u : 0u);
      }
    } else {
      // consecutive MMAs write DIFFERENT accumulators (rr of them), each still receives 9 taps per group
      for (int g = 0; g < groups; g += c.rr) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
          for (int a = 0; a < c.rr; ++a)
            umma(tmem + (uint32_t)(a * c.N), ad + (uint64_t)((t * c.a_step + a * 8 * c.pitch_a) >> 4),
                 bd + (uint64_t)((t * c.b_step) >> 4), idesc, t ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(su32(&bar)), "r"(0) : "memory");
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  long long h[148 * 2];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int G = 128;
  auto run = [&](const char* name, Cfg c) {
    for (int rep = 0; rep < 2; ++rep) { probe<<<148, 64, 170 * 1024>>>(c, G, d); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%-58s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); } }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mi = 1e30, me = 1e30;
    for (int i = 0; i < 148; ++i) { if (h[2 * i] < mi) mi = h[2 * i]; if (h[2 * i + 1] < me) me = h[2 * i + 1]; }
    printf("%-58s issue %.1f  complete %.1f cycles/MMA   (%.0f MACs/cycle/SM)\n", name, mi / (9.0 * G), me / (9.0 * G),
           (double)c.M * c.N * 16 * 9 * G / me);
  };
  for (int rr : {1, 2, 4, 8})
    for (int N : {16, 48}) {
      char nm[96]; snprintf(nm, sizeof nm, "K-major SW32 M=128 N=%-3d round-robin over %d accumulators", N, rr);
      run(nm, Cfg{6, 6, 0, 0, 128, N, 32, 32, 256, 0, 4, rr});
    }
  // un-swizzled "planar" A (8 pixels x 16 B core matrices, halo rows 640 B apart, chunk planes 11584 B apart):
  // tap = start address shifted by t pixels (16 B): aligned (t*128) vs unaligned (t*16) starts
  for (int N : {16, 48}) {
    char nm[96];
    snprintf(nm, sizeof nm, "planar A (no swizzle) M=128 N=%-3d starts t*128 B (aligned)", N);
    run(nm, Cfg{0, 6, 0, 0, 128, N, 16, 32, 128, 0, 4, 0, 640, 11584});
    snprintf(nm, sizeof nm, "planar A (no swizzle) M=128 N=%-3d starts t*16 B (pixel shifts)", N);
    run(nm, Cfg{0, 6, 0, 0, 128, N, 16, 32, 16, 0, 4, 0, 640, 11584});
    snprintf(nm, sizeof nm, "planar A (no swizzle) M=128 N=%-3d starts t*16 B, planes 11648 B", N);
    run(nm, Cfg{0, 6, 0, 0, 128, N, 16, 32, 16, 0, 4, 0, 640, 11648});
  }
  // K-major A (conv forward): rows of pitch bytes, taps 32 B apart in the descriptor start (a_step)
  const int lay[4] = {6, 4, 2, 0}; const int pit[4] = {32, 64, 128, 16}; const char* ln[4] = {"SW32 (Cin=16)", "SW64 (Cin=32)", "SW128 (Cin=64)", "no swizzle"};
  for (int li = 0; li < 1; ++li)
    for (int N : {16, 48, 128}) {
      char nm[96]; snprintf(nm, sizeof nm, "K-major A %-15s M=128 N=%-3d", ln[li], N);
      run(nm, Cfg{lay[li], lay[li], 0, 0, 128, N, pit[li], pit[li], 8 * pit[li], 0, 512 / N < 4 ? 512 / N : 4});
    }
  for (int N : {16, 48, 96}) {
    char nm[96]; snprintf(nm, sizeof nm, "K-major A SW32 M=64 N=%-3d", N);
    run(nm, Cfg{6, 6, 0, 0, 64, N, 32, 32, 256, 0, 4});
  }
  // MN-major operands (filter gradient): A = 8 groups of 16 channels (LBO apart), K rows of 32 B
  for (int li = 0; li < 1; ++li)
    for (int N : {16}) {
      for (int M : {128, 64}) {
        char nm[96]; snprintf(nm, sizeof nm, "MN-major A,B %-15s M=%-3d N=%-3d", ln[li], M, N);
        run(nm, Cfg{lay[li], lay[li], 1, 1, M, N, pit[li], pit[li], 16 * pit[li], 16 * pit[li], 4});
      }
    }
  return 0;
}
