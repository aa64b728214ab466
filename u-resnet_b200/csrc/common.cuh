// Shared helpers for the uresnet_b200 kernels (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <type_traits>

#include "../../include/uresnet_b200.h"

typedef __nv_bfloat16 bf16;
typedef __half f16;

// ---- error plumbing (thread-local message, int status across the C ABI) ----
void urn_set_error(const char* fmt, ...);
#define URN_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      urn_set_error(__VA_ARGS__);            \
      return 1;                              \
    }                                        \
  } while (0)
#define URN_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      urn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)
extern long long g_urn_launches;   // kernels launched by this library (host-side count)
#define URN_LAUNCH_CHECK()          \
  do {                              \
    ++g_urn_launches;               \
    URN_CUDA(cudaGetLastError());   \
  } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// CTAs of a persistent kernel: one per SM.  URN_MAX_CTAS (tests) caps it so that small shapes
// exercise the many-tiles-per-CTA paths (ring wrap-around, look-ahead waits, partial merges).
static inline int persistent_ctas() {
  if (const char* e = getenv("URN_MAX_CTAS")) {
    const int v = atoi(e);
    if (v > 0) return v < num_sms() ? v : num_sms();
  }
  return num_sms();
}

// ---- typed element access: storage T in {float, bf16}, math in fp32 ----
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float ld1(const f16* p) { return __half2float(*p); }
__device__ __forceinline__ void st1(f16* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float4 ld4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 ld4(const f16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 fa = __half22float2(*reinterpret_cast<__half2*>(&u.x));
  float2 fb = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(f16* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y);
  __half2 b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void st4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 8 x 16-bit in one 16-byte load
__device__ __forceinline__ void load8(const f16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t r[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t r[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(r[i] << 16);
    v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 values -> 16 bytes of 16-bit storage (32 bytes of fp32)
__device__ __forceinline__ void store8(f16* p, const float* v) {
  uint4 u;
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  uint4 u;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- storage dtype dispatch: f(T*) is called with a null pointer of the storage type ----
template <typename F>
static inline int dispatch_dtype(int dt, F&& f) {
  switch (dt) {
    case URN_F32: return f((float*)nullptr);
    case URN_BF16: return f((bf16*)nullptr);
    case URN_F16: return f((f16*)nullptr);
  }
  urn_set_error("bad dtype %d", dt);
  return 1;
}
#define URN_T(p) typename std::remove_pointer<decltype(p)>::type
static inline int dtype_size(int dt) { return dt == URN_F32 ? 4 : 2; }

// ---- generalised conv launch description shared by the FFMA and tcgen05 paths ----
// out[n, o*os + oo, :] = sum_t in[n, o*is + off_t, :] . Wt      (o over a sub-grid)
struct GTap {
  int od, oh, ow;  // input offset of this tap (relative to o*is)
  int widx;        // tap index into the filter (kd,kh,kw linearised)
};
struct GConv {
  int n;
  int id, ih, iw, cin;   // input tensor dims
  int od, oh, ow, cout;  // FULL output tensor dims
  int gd, gh, gw;        // sub-grid extents (number of o's per dim)
  int is, os;            // input sampling stride / output placement stride
  int ood, ooh, oow;     // output placement offset
  int ntaps;
  GTap taps[27];
  // filter: element (tap, kin, kout) at w[tap*wtap + kin*wldi + kout*wldo]
  int64_t wtap, wldi, wldo;
  int64_t ldin, ldout;
  int accumulate;
};

// Internal entry points (implemented in gconv_ffma.cu / gconv_tc.cu)
int gconv_ffma_launch(const GConv& g, int in_dtype, const void* in, const float* w, void* out,
                      int out_dtype, cudaStream_t st);
int wgrad_ffma_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                      const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);
int64_t wgrad_ffma_workspace(const urn_conv_desc& d);
