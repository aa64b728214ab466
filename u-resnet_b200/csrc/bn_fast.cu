// Fast paths of the batch-norm kernels for the shapes the network
// actually has: power-of-two channel counts >= 8, pitches that are multiples of 8 elements and
// 16-byte aligned views.  Every thread owns ONE group of 8 channels for its whole life (the grid
// stride is a multiple of the groups per row), so the per-channel constants live in registers,
// all index arithmetic is shifts, every global access is a 16-byte vector, and U independent
// rows are in flight per thread.  Reductions keep the fixed-order, double-precision second
// stage of bn.cu (deterministic).
//
// Backward byte diet (16 B/elem instead of 18, 10 instead of 14):
//   * BN without ReLU (resnet_conv1, /root/reference/lib/resnet_module.py:43-51): the stored
//     forward output IS xhat + beta, so xhat = out - beta and the fp32 raw tensor is not read;
//   * BN -> ReLU without residual (conv0, deconv, conv1): the mask out > 0 is recomputed from
//     raw with the forward's exact expression fma(x - mean, rstd, beta), so `out` is not read.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "bn_fast.cuh"

namespace {

constexpr int BF_T = 256;

__device__ __forceinline__ void load_params8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 consecutive elements kept PACKED while their loads are in flight (4 registers per 16 bytes):
// U rows x several tensors can be outstanding per thread without the unpacked fp32 copies
// inflating the register footprint (reads need ~100 KB in flight per SM to saturate HBM3e; the
// first version unpacked at the load and topped out at 3.3 TB/s of read traffic).
template <typename T> struct Raw8 { uint4 a; };
template <> struct Raw8<float> { uint4 a, b; };
template <typename T> __device__ __forceinline__ Raw8<T> ldraw8(const T* p) {
  Raw8<T> r; r.a = *reinterpret_cast<const uint4*>(p); return r;
}
template <> __device__ __forceinline__ Raw8<float> ldraw8<float>(const float* p) {
  Raw8<float> r; r.a = *reinterpret_cast<const uint4*>(p); r.b = *reinterpret_cast<const uint4*>(p + 4); return r;
}
template <typename T> __device__ __forceinline__ Raw8<T> zero8() { Raw8<T> r; r.a = make_uint4(0, 0, 0, 0); return r; }
template <> __device__ __forceinline__ Raw8<float> zero8<float>() {
  Raw8<float> r; r.a = make_uint4(0, 0, 0, 0); r.b = r.a; return r;
}
__device__ __forceinline__ void unpack8(const Raw8<f16>& r, float* v) {
  const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack8(const Raw8<bf16>& r, float* v) {
  const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float* v) {
  v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
  v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
}

// Sum K per-thread vectors of 8 channels over all threads of the block that own the same
// channel group; thread ch < c then writes partial[(block*K + k)*c + ch] as double.
template <int K>
__device__ __forceinline__ void block_group_reduce(float (&acc)[K][8], int cg, int c, double* __restrict__ partial) {
  __shared__ float red[BF_T * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid & (cg - 1);
  int nl, slot;
  if (cg < 32) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[k][j];
        for (int o = 16; o >= cg; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[k][j] = v;
      }
    nl = BF_T / 32;
    slot = lane < cg ? warp * cg + g : -1;
  } else {
    nl = BF_T / cg;
    slot = tid;                        // == (tid / cg) * cg + g
  }
  for (int k = 0; k < K; ++k) {
    __syncthreads();
    if (slot >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[slot * 8 + j] = acc[k][j];
    }
    __syncthreads();
    for (int ch = tid; ch < c; ch += BF_T) {
      double s = 0.0;
      for (int l = 0; l < nl; ++l) s += (double)red[(l * cg + (ch >> 3)) * 8 + (ch & 7)];
      partial[((long long)blockIdx.x * K + k) * c + ch] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------
// forward statistics (shifted sums, see bn.cu)
// ---------------------------------------------------------------------------------------
template <typename TW>      // TW: storage type of the raw conv output (float, or f16 in the 16-bit network modes)
__global__ void __launch_bounds__(BF_T)
bn_stats_fast_kernel(const TW* __restrict__ raw, long long ld, long long rows, int c, int lg,
                     double* __restrict__ partial) {
  constexpr int U = 4;
  const int cg = c >> 3;
  const long long tid = (long long)blockIdx.x * BF_T + threadIdx.x;
  const int ch = (int)(tid & (cg - 1)) * 8;
  const long long rstride = ((long long)gridDim.x * BF_T) >> lg;
  float K[8];
  load8(raw + ch, K);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  for (long long r = tid >> lg; r < rows; r += rstride * U) {
    float x[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      if (rr < rows) load8(raw + rr * ld + ch, x[u]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = K[j];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = x[u][j] - K[j];
        acc[0][j] += v;
        acc[1][j] = fmaf(v, v, acc[1][j]);
      }
  }
  block_group_reduce<2>(acc, cg, c, partial);
}

// ---------------------------------------------------------------------------------------
// forward apply: out = act(BN(raw) [+ res | + BN(res_raw)])
// ---------------------------------------------------------------------------------------
template <typename TW, typename TR, typename TO, int RES, bool RELU, bool SHADOW>
__global__ void __launch_bounds__(BF_T)
bn_apply_fast_kernel(const TW* __restrict__ raw, long long ld, long long rows, int c, int lg,
                     const float* __restrict__ stats, const float* __restrict__ beta,
                     const TR* __restrict__ res, long long ldres, const TW* __restrict__ res_raw,
                     long long ldres_raw, const float* __restrict__ res_stats,
                     const float* __restrict__ res_beta, TO* __restrict__ out, long long ldout,
                     bf16* __restrict__ out2) {
  constexpr int U = RES == 2 ? 2 : 4;
  const int cg = c >> 3;
  const long long tid = (long long)blockIdx.x * BF_T + threadIdx.x;
  const int ch = (int)(tid & (cg - 1)) * 8;
  const long long rstride = ((long long)gridDim.x * BF_T) >> lg;
  float mu[8], rs[8], be[8], mu2[8], rs2[8], be2[8];
  load_params8(stats + ch, mu); load_params8(stats + c + ch, rs); load_params8(beta + ch, be);
  if (RES == 2) { load_params8(res_stats + ch, mu2); load_params8(res_stats + c + ch, rs2); load_params8(res_beta + ch, be2); }
  for (long long r = tid >> lg; r < rows; r += rstride * U) {
    float x[U][8], x2[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      if (rr < rows) {
        load8(raw + rr * ld + ch, x[u]);
        if (RES == 2) load8(res_raw + rr * ldres_raw + ch, x2[u]);
        if (RES == 1) load8(res + rr * ldres + ch, x2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      if (rr >= rows) continue;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(x[u][j] - mu[j], rs[j], be[j]);
        if (RES == 2) v += fmaf(x2[u][j] - mu2[j], rs2[j], be2[j]);
        if (RES == 1) v += x2[u][j];
        y[j] = RELU ? fmaxf(v, 0.f) : v;
      }
      store8(out + rr * ldout + ch, y);
      if (SHADOW) store8(out2 + rr * ldout + ch, y);
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward pass 1: per-channel sums of g, g*xhat (and g*xhat2)
//   XMODE 0: xhat from raw; 1: xhat = out - beta (no ReLU); 2: xhat and the ReLU mask from raw
// ---------------------------------------------------------------------------------------
template <typename TW, typename TG, typename TA, bool RELU, bool DUAL, int XMODE, int U>
__global__ void __launch_bounds__(BF_T, DUAL ? 2 : 3)
bn_bwd_reduce_fast_kernel(const TG* __restrict__ dout, long long lddout, const TA* __restrict__ outp,
                          long long ldo, long long rows, int c, int lg, const TW* __restrict__ raw,
                          long long ldraw, const float* __restrict__ stats, const float* __restrict__ beta,
                          const TW* __restrict__ raw2, long long ldraw2,
                          const float* __restrict__ stats2, double* __restrict__ partial) {
  constexpr bool NEED_OUT = (RELU && XMODE != 2) || XMODE == 1;
  constexpr bool NEED_RAW = XMODE != 1;
  const int cg = c >> 3;
  const long long tid = (long long)blockIdx.x * BF_T + threadIdx.x;
  const int ch = (int)(tid & (cg - 1)) * 8;
  const long long rstride = ((long long)gridDim.x * BF_T) >> lg;
  float mu[8], rs[8], be[8], mu2[8], rs2[8];
  load_params8(stats + ch, mu); load_params8(stats + c + ch, rs);
  if (XMODE != 0) load_params8(beta + ch, be);
  if (DUAL) { load_params8(stats2 + ch, mu2); load_params8(stats2 + c + ch, rs2); }
  float acc[3][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  for (long long r = tid >> lg; r < rows; r += rstride * U) {
    Raw8<TG> gb[U];
    Raw8<TA> ob[NEED_OUT ? U : 1];
    Raw8<TW> xb[NEED_RAW ? U : 1], x2b[DUAL ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      const bool ok = rr < rows;
      gb[u] = ok ? ldraw8(dout + rr * lddout + ch) : zero8<TG>();
      if (NEED_OUT) ob[u] = ok ? ldraw8(outp + rr * ldo + ch) : zero8<TA>();
      if (NEED_RAW) xb[u] = ok ? ldraw8(raw + rr * ldraw + ch) : zero8<TW>();
      if (DUAL) x2b[u] = ok ? ldraw8(raw2 + rr * ldraw2 + ch) : zero8<TW>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float g[8], o[8], x[8], x2[8];
      unpack8(gb[u], g);
      if (NEED_OUT) unpack8(ob[u], o);
      if (NEED_RAW) unpack8(xb[u], x);
      if (DUAL) unpack8(x2b[u], x2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gv = g[j], xh;
        if (XMODE == 1) xh = o[j] - be[j];
        else xh = (x[j] - mu[j]) * rs[j];
        if (RELU) {
          if (XMODE == 2) { if (!(fmaf(x[j] - mu[j], rs[j], be[j]) > 0.f)) gv = 0.f; }
          else if (!(o[j] > 0.f)) gv = 0.f;
        }
        acc[0][j] += gv;
        acc[1][j] = fmaf(gv, xh, acc[1][j]);
        if (DUAL) acc[2][j] = fmaf(gv, (x2[j] - mu2[j]) * rs2[j], acc[2][j]);
      }
    }
  }
  block_group_reduce<3>(acc, cg, c, partial);
}

// ---------------------------------------------------------------------------------------
// backward pass 2: d(raw) = rstd*(g - mean(g) - xhat*mean(g*xhat))  (+ dual BN, + residual fan-out)
//   DRES 0: none, 1: write g, 2: accumulate g
// ---------------------------------------------------------------------------------------
template <typename TW, typename TG, typename TA, bool RELU, bool DUAL, int DRES, int XMODE>
__global__ void __launch_bounds__(BF_T)
bn_bwd_apply_fast_kernel(const TG* __restrict__ dout, long long lddout, const TA* __restrict__ outp,
                         long long ldo, long long rows, int c, int lg, const TW* __restrict__ raw,
                         long long ldraw, const float* __restrict__ stats, const float* __restrict__ beta,
                         TG* __restrict__ draw, long long lddraw, const TW* __restrict__ raw2,
                         long long ldraw2, const float* __restrict__ stats2, TG* __restrict__ draw2,
                         long long lddraw2, TG* __restrict__ dres, long long lddres,
                         const float* __restrict__ sums) {
  constexpr int U = DUAL ? 2 : 4;
  constexpr bool NEED_OUT = (RELU && XMODE != 2) || XMODE == 1;
  constexpr bool NEED_RAW = XMODE != 1;
  const int cg = c >> 3;
  const long long tid = (long long)blockIdx.x * BF_T + threadIdx.x;
  const int ch = (int)(tid & (cg - 1)) * 8;
  const long long rstride = ((long long)gridDim.x * BF_T) >> lg;
  const float inv = 1.0f / (float)rows;
  float mu[8], rs[8], be[8], mg[8], mgx[8], mu2[8], rs2[8], mgx2[8];
  load_params8(stats + ch, mu); load_params8(stats + c + ch, rs);
  load_params8(sums + ch, mg); load_params8(sums + c + ch, mgx);
  if (XMODE != 0) load_params8(beta + ch, be);
  if (DUAL) { load_params8(stats2 + ch, mu2); load_params8(stats2 + c + ch, rs2); load_params8(sums + 2 * c + ch, mgx2); }
#pragma unroll
  for (int j = 0; j < 8; ++j) { mg[j] *= inv; mgx[j] *= inv; if (DUAL) mgx2[j] *= inv; }
  for (long long r = tid >> lg; r < rows; r += rstride * U) {
    Raw8<TG> gb[U], drb[DRES == 2 ? U : 1];
    Raw8<TA> ob[NEED_OUT ? U : 1];
    Raw8<TW> xb[NEED_RAW ? U : 1], x2b[DUAL ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      if (rr < rows) {
        gb[u] = ldraw8(dout + rr * lddout + ch);
        if (NEED_OUT) ob[u] = ldraw8(outp + rr * ldo + ch);
        if (NEED_RAW) xb[u] = ldraw8(raw + rr * ldraw + ch);
        if (DUAL) x2b[u] = ldraw8(raw2 + rr * ldraw2 + ch);
        if (DRES == 2) drb[u] = ldraw8(dres + rr * lddres + ch);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u * rstride;
      if (rr >= rows) continue;
      float g[8], o[8], x[8], x2[8], dr[8], o1[8], o2[8], gg[8];
      unpack8(gb[u], g);
      if (NEED_OUT) unpack8(ob[u], o);
      if (NEED_RAW) unpack8(xb[u], x);
      if (DUAL) unpack8(x2b[u], x2);
      if (DRES == 2) unpack8(drb[u], dr);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gv = g[j], xh;
        if (XMODE == 1) xh = o[j] - be[j];
        else xh = (x[j] - mu[j]) * rs[j];
        if (RELU) {
          if (XMODE == 2) { if (!(fmaf(x[j] - mu[j], rs[j], be[j]) > 0.f)) gv = 0.f; }
          else if (!(o[j] > 0.f)) gv = 0.f;
        }
        o1[j] = rs[j] * (gv - mg[j] - xh * mgx[j]);
        if (DUAL) o2[j] = rs2[j] * (gv - mg[j] - (x2[j] - mu2[j]) * rs2[j] * mgx2[j]);
        gg[j] = DRES == 2 ? gv + dr[j] : gv;
      }
      store8(draw + rr * lddraw + ch, o1);
      if (DUAL) store8(draw2 + rr * lddraw2 + ch, o2);
      if (DRES) store8(dres + rr * lddres + ch, gg);
    }
  }
}

inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// grid: a multiple of the SM count; grid*256 is always a multiple of c/8 <= 64
inline int fast_grid(long long rows, int c, int per_sm) {
  const long long items = rows * (c >> 3);
  long long b = (items + BF_T - 1) / BF_T;
  const long long cap = (long long)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

bool bn_fast_shape(int c, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  if (c < 8 || c > 512 || (c & (c - 1))) return false;
  for (long long l : lds)
    if (l % 8) return false;
  for (const void* p : ptrs)
    if (p && !al16(p)) return false;
  return true;
}

int bn_stats_fast_launch(const void* raw, int raw_dtype, long long ld, long long rows, int c, double* partial,
                         int* nblocks, cudaStream_t st) {
  const int nb = std::min(fast_grid(rows, c, 4), 148 * 4);      // partial buffer: BN_MAX_BLOCKS entries
  if (raw_dtype == URN_F16) bn_stats_fast_kernel<f16><<<nb, BF_T, 0, st>>>((const f16*)raw, ld, rows, c, ilog2(c >> 3), partial);
  else bn_stats_fast_kernel<float><<<nb, BF_T, 0, st>>>((const float*)raw, ld, rows, c, ilog2(c >> 3), partial);
  URN_LAUNCH_CHECK();
  *nblocks = nb;
  return 0;
}

template <typename TW, typename TR, typename TO>
static int bn_apply_fast_t(const TW* raw, long long ld, long long rows, int c, const float* stats,
                           const float* beta, const void* res, long long ldres, const TW* res_raw,
                           long long ldres_raw, const float* res_stats, const float* res_beta, int relu,
                           void* out, long long ldout, void* out2, cudaStream_t st) {
  const int nb = fast_grid(rows, c, 8), lg = ilog2(c >> 3);
  const int resk = res_raw ? 2 : (res ? 1 : 0);
#define BF_APPLY(RESV, RELUV, SHV)                                                                     \
  bn_apply_fast_kernel<TW, TR, TO, RESV, RELUV, SHV><<<nb, BF_T, 0, st>>>(raw, ld, rows, c, lg, stats, beta, \
      (const TR*)res, ldres, res_raw, ldres_raw, res_stats, res_beta, (TO*)out, ldout, (bf16*)out2)
#define BF_APPLY_SH(RESV, RELUV) do { if (out2) BF_APPLY(RESV, RELUV, true); else BF_APPLY(RESV, RELUV, false); } while (0)
  if (resk == 0) { if (relu) BF_APPLY_SH(0, true); else BF_APPLY_SH(0, false); }
  else if (resk == 1) { if (relu) BF_APPLY_SH(1, true); else BF_APPLY_SH(1, false); }
  else { if (relu) BF_APPLY_SH(2, true); else BF_APPLY_SH(2, false); }
#undef BF_APPLY_SH
#undef BF_APPLY
  URN_LAUNCH_CHECK();
  return 0;
}

int bn_apply_fast_launch(const void* raw, int raw_dtype, long long ld, long long rows, int c, const float* stats,
                         const float* beta, const void* res, long long ldres, int res_dtype,
                         const void* res_raw, long long ldres_raw, const float* res_stats,
                         const float* res_beta, int relu, void* out, long long ldout, int out_dtype,
                         void* out2, cudaStream_t st) {
  if (!res) res_dtype = out_dtype;
  if (raw_dtype == URN_F16) {       // fp16 raw only exists next to fp16 activations
    URN_CHECK(res_dtype == URN_F16 && out_dtype == URN_F16, "bn_apply: f16 raw needs f16 activations");
    return bn_apply_fast_t<f16, f16, f16>((const f16*)raw, ld, rows, c, stats, beta, res, ldres, (const f16*)res_raw,
                                          ldres_raw, res_stats, res_beta, relu, out, ldout, out2, st);
  }
  return dispatch_dtype(res_dtype, [&](auto pr) {
    using TR = URN_T(pr);
    return dispatch_dtype(out_dtype, [&](auto po) {
      using TO = URN_T(po);
      return bn_apply_fast_t<float, TR, TO>((const float*)raw, ld, rows, c, stats, beta, res, ldres, (const float*)res_raw,
                                            ldres_raw, res_stats, res_beta, relu, out, ldout, out2, st);
    });
  });
}

template <typename TW, typename TG, typename TA>
static int bn_bwd_fast_t(const void* dout, long long lddout, const void* out, long long ldo, int relu,
                         long long rows, int c, const TW* raw, long long ldraw, const float* stats,
                         const float* beta, void* draw, long long lddraw, const TW* raw2, long long ldraw2,
                         const float* stats2, void* draw2, long long lddraw2, void* dres, long long lddres,
                         int dres_acc, double* partial, const float* sums, int phase, int* nblocks,
                         cudaStream_t st) {
  const int lg = ilog2(c >> 3);
  const bool dual = raw2 != nullptr;
  // xhat source: plain BN (no ReLU, no second BN, no residual fan-out) -> from `out`;
  // BN -> ReLU without residual -> mask from raw; everything else -> classic
  int xmode = 0;
  if (beta && !dual && !dres) xmode = relu ? 2 : 1;
  if (phase == 0) {
    static const int psm = getenv("URN_BN_PSM") ? atoi(getenv("URN_BN_PSM")) : (dual ? 2 : 3);
    const int nb = std::min(fast_grid(rows, c, psm), 148 * 4);
    *nblocks = nb;
    static const int uenv = getenv("URN_BN_U") ? atoi(getenv("URN_BN_U")) : 4;
#define BF_RED(RELUV, DUALV, XM)                                                                        \
  do {                                                                                                  \
    if (uenv == 8) bn_bwd_reduce_fast_kernel<TW, TG, TA, RELUV, DUALV, XM, 8><<<nb, BF_T, 0, st>>>((const TG*)dout, lddout, \
      (const TA*)out, ldo, rows, c, lg, raw, ldraw, stats, beta, raw2, ldraw2, stats2, partial);        \
    else if (uenv == 2) bn_bwd_reduce_fast_kernel<TW, TG, TA, RELUV, DUALV, XM, 2><<<nb, BF_T, 0, st>>>((const TG*)dout, lddout, \
      (const TA*)out, ldo, rows, c, lg, raw, ldraw, stats, beta, raw2, ldraw2, stats2, partial);        \
    else bn_bwd_reduce_fast_kernel<TW, TG, TA, RELUV, DUALV, XM, 4><<<nb, BF_T, 0, st>>>((const TG*)dout, lddout, \
      (const TA*)out, ldo, rows, c, lg, raw, ldraw, stats, beta, raw2, ldraw2, stats2, partial);        \
  } while (0)
    if (xmode == 1) BF_RED(false, false, 1);
    else if (xmode == 2) BF_RED(true, false, 2);
    else if (dual) { if (relu) BF_RED(true, true, 0); else BF_RED(false, true, 0); }
    else { if (relu) BF_RED(true, false, 0); else BF_RED(false, false, 0); }
#undef BF_RED
    URN_LAUNCH_CHECK();
    return 0;
  }
  const int nb = fast_grid(rows, c, 8);
  const int dr = dres ? (dres_acc ? 2 : 1) : 0;
#define BF_APP(RELUV, DUALV, DRESV, XM)                                                                  \
  bn_bwd_apply_fast_kernel<TW, TG, TA, RELUV, DUALV, DRESV, XM><<<nb, BF_T, 0, st>>>((const TG*)dout, lddout, \
      (const TA*)out, ldo, rows, c, lg, raw, ldraw, stats, beta, (TG*)draw, lddraw, raw2, ldraw2, stats2,  \
      (TG*)draw2, lddraw2, (TG*)dres, lddres, sums)
#define BF_APP_DR(RELUV, DUALV)                                         \
  do {                                                                  \
    if (dr == 0) BF_APP(RELUV, DUALV, 0, 0);                            \
    else if (dr == 1) BF_APP(RELUV, DUALV, 1, 0);                       \
    else BF_APP(RELUV, DUALV, 2, 0);                                    \
  } while (0)
  if (xmode == 1) BF_APP(false, false, 0, 1);
  else if (xmode == 2) BF_APP(true, false, 0, 2);
  else if (dual) { if (relu) BF_APP_DR(true, true); else BF_APP_DR(false, true); }
  else { if (relu) BF_APP_DR(true, false); else BF_APP_DR(false, false); }
#undef BF_APP_DR
#undef BF_APP
  URN_LAUNCH_CHECK();
  return 0;
}

int bn_bwd_fast_launch(int phase, const void* dout, long long lddout, const void* out, long long ldo, int dtype,
                       int out_dtype, int relu, long long rows, int c, const void* raw, int raw_dtype, long long ldraw,
                       const float* stats, const float* beta, void* draw, long long lddraw, const void* raw2,
                       long long ldraw2, const float* stats2, void* draw2, long long lddraw2, void* dres,
                       long long lddres, int dres_acc, double* partial, const float* sums, int* nblocks,
                       cudaStream_t st) {
#define URN_BBF(WD, GD, AD, TWT, TGT, TAT)                                                               \
  if (raw_dtype == WD && dtype == GD && out_dtype == AD)                                                 \
    return bn_bwd_fast_t<TWT, TGT, TAT>(dout, lddout, out, ldo, relu, rows, c, (const TWT*)raw, ldraw, stats, beta, draw, \
                                   lddraw, (const TWT*)raw2, ldraw2, stats2, draw2, lddraw2, dres, lddres, dres_acc, \
                                   partial, sums, phase, nblocks, st);
  URN_BBF(URN_F32, URN_F32, URN_F32, float, float, float)
  URN_BBF(URN_F32, URN_BF16, URN_BF16, float, bf16, bf16)
  URN_BBF(URN_F32, URN_BF16, URN_F16, float, bf16, f16)
  URN_BBF(URN_F16, URN_BF16, URN_F16, f16, bf16, f16)
#undef URN_BBF
  urn_set_error("bn_bwd_fast: unsupported dtypes (raw %d, grad %d, act %d)", raw_dtype, dtype, out_dtype);
  return 1;
}
