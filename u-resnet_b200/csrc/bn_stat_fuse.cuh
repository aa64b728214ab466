// Per-warp running BN statistics for conv epilogues (gconv_halo.cu, gconv_tc.cu).
#pragma once
#include "common.cuh"

namespace {

// Fused BN statistics (SURVEY 8a K5 'stats fused into the conv epilogue').  A warp holds 32
// positions x 16 channels of fp32 accumulator values.  Per channel the sum and the sum of squares
// over the 32 positions are formed with a halving butterfly (16 shuffles per quantity: after the
// xor-16/8/4/2 steps each lane holds one channel, the xor-1 step completes it), turned into the
// chunk's (mean, M2) and merged into the lane's running (n, mean, M2) with Chan's update -- so no
// E[x^2]-E[x]^2 cancellation over more than 128 values (one tile plane) ever happens in fp32.  Lane L owns channel
// c0 + (L >> 1).  Invalid positions (tile overhang) contribute zeros and are not counted.
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 16;
    const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 8;
    const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 4;
    const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool hi = lane & 2;
    const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
struct RunStat { float n, mean, m2; };
__device__ __forceinline__ void stat_merge(RunStat& r, float (&a)[16], float (&b)[16], unsigned nv, int lane) {
  if (nv == 0) return;                                   // warp-uniform
  const float s1 = butterfly16(a, lane), s2 = butterfly16(b, lane);
  const float cn = (float)nv, cm = s1 / cn;
  const float cm2 = fmaxf(s2 - s1 * cm, 0.f);
  const float nn = r.n + cn, delta = cm - r.mean;
  r.mean += delta * (cn / nn);
  r.m2 += cm2 + delta * delta * (r.n * cn / nn);
  r.n = nn;
}

}  // namespace
