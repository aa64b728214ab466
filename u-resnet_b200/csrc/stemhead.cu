// Dedicated CUDA-core kernels for the two layers whose channel counts do not fit a tensor-core
// tile: the stem conv0 (Cin = 1 -> F, /root/reference/lib/uresnet.py:37-45) and the head conv2
// (F -> num_class <= 4, uresnet.py:113-121).  Both are 3x3 / 3x3x3, stride 1, TF 'same'.
// They are bandwidth-shaped: the stem reads 1 value and writes F per pixel, the head reads F and
// writes 3.  On the generic FFMA implicit-GEMM kernel (tiles of 16 channels) they cost 22 ms of
// a 73 ms 2D step and 215 ms of a 3D step; here every thread owns whole pixels, the taps are
// unrolled, filters live in shared memory, and HBM access is vectorised and coalesced.
//
//   stem forward      y[p][co]  = sum_t x[p+t] w[t][co]              4 pixels x CO per thread
//   stem filter grad  dw[t][co] = sum_q x[q] dy[q-t][co]            x-centric (warps whose 32
//                                                                   pixels are all zero skip)
//   head forward      y[p][k]   = sum_t sum_ci x[p+t][ci] w[t][ci][k]
//   head input grad   dx[q][ci] = sum_t sum_k dy[q-t][k] w[t][ci][k]
//   head filter grad  dw[t][ci][k] = sum_p x[p+t][ci] dy[p][k]      one (kd,kh) tap row per
//                                                                   blockIdx.y, 3*CI*NC accumulators
// Filter gradients use per-block fp32 partials and a fixed-order second stage (deterministic).
#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"

namespace {

constexpr int SH_T = 128;   // threads per block
constexpr unsigned FULL = 0xffffffffu;

struct Geo {
  int n, D, H, W, kd, pd;
  long long rows;
};

__device__ __forceinline__ void decode(long long p, const Geo& g, int& nn, int& d, int& h, int& w) {
  w = (int)(p % g.W); p /= g.W;
  h = (int)(p % g.H); p /= g.H;
  d = (int)(p % g.D);
  nn = (int)(p / g.D);
}

// ------------------------------------------------------------------------------------
// stem forward
// ------------------------------------------------------------------------------------
template <typename TI, int KD, int CO>
__global__ void __launch_bounds__(SH_T)
stem_fwd_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ w,
                float* __restrict__ y, long long ldy, const Geo g) {
  constexpr int NT = KD * 9, Q = CO / 4, NV = CO;     // NV float4 per lane (4 pixels x Q)
  __shared__ __align__(16) float wsm[NT * CO];
  __shared__ float4 stage[SH_T / 32][32 * NV];
  for (int i = threadIdx.x; i < NT * CO; i += SH_T) wsm[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long W4 = g.W / 4, total = g.rows / 4;
  const long long total_pad = (total + 31) / 32 * 32;
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
  for (long long grp = (long long)blockIdx.x * SH_T + threadIdx.x; grp < total_pad;
       grp += (long long)gridDim.x * SH_T) {
    const bool valid = grp < total;
    float acc[4][CO];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < CO; ++c) acc[i][c] = 0.f;
    if (valid) {
      long long r = grp;
      const int w0 = (int)(r % W4) * 4; r /= W4;
      const int h = (int)(r % g.H); r /= g.H;
      const int d = (int)(r % g.D);
      const int nn = (int)(r / g.D);
#pragma unroll
      for (int kd = 0; kd < KD; ++kd) {
        const int zz = d + kd - g.pd;
        if (zz < 0 || zz >= g.D) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int yy = h + kh - 1;
          if (yy < 0 || yy >= g.H) continue;
          const TI* row = x + (((long long)nn * g.D + zz) * g.H + yy) * g.W * ldx;
          float xv[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int ww = w0 - 1 + i;
            xv[i] = (ww >= 0 && ww < g.W) ? ld1(row + ww * ldx) : 0.f;
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int t = (kd * 3 + kh) * 3 + kw;
#pragma unroll
            for (int c4 = 0; c4 < Q; ++c4) {
              const float4 wv = w4[t * Q + c4];
#pragma unroll
              for (int px = 0; px < 4; ++px) {
                const float a = xv[px + kw];
                acc[px][c4 * 4 + 0] = fmaf(a, wv.x, acc[px][c4 * 4 + 0]);
                acc[px][c4 * 4 + 1] = fmaf(a, wv.y, acc[px][c4 * 4 + 1]);
                acc[px][c4 * 4 + 2] = fmaf(a, wv.z, acc[px][c4 * 4 + 2]);
                acc[px][c4 * 4 + 3] = fmaf(a, wv.w, acc[px][c4 * 4 + 3]);
              }
            }
          }
        }
      }
    }
    // transpose through shared memory so that every store instruction writes 512 contiguous bytes
    float4* sg = stage[warp];
    __syncwarp();
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int c4 = 0; c4 < Q; ++c4) {
        const int col = px * Q + c4;
        sg[lane * NV + (col ^ (lane & (NV - 1)))] =
            make_float4(acc[px][c4 * 4], acc[px][c4 * 4 + 1], acc[px][c4 * 4 + 2], acc[px][c4 * 4 + 3]);
      }
    __syncwarp();
    const long long gbase = grp - lane;                    // first group of this warp
    float4* out4 = reinterpret_cast<float4*>(y + gbase * 4 * ldy);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = k * 32 + lane, row = c / NV, col = c % NV;
      if (gbase + row < total) out4[c] = sg[row * NV + (col ^ (row & (NV - 1)))];
    }
  }
}

// ------------------------------------------------------------------------------------
// stem filter gradient (x-centric)
// ------------------------------------------------------------------------------------
template <typename TI, typename TG, int KD, int CO>
__global__ void __launch_bounds__(SH_T)
stem_wgrad_kernel(const TI* __restrict__ x, long long ldx, const TG* __restrict__ dy, long long ldy,
                  float* __restrict__ partial, const Geo g) {
  constexpr int NT = KD * 9, Q = CO / 4;
  __shared__ float red[SH_T / 32][NT * CO];
  float acc[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total = g.rows * Q;
  const long long total_pad = (total + 31) / 32 * 32;
  // UW independent x loads are in flight per thread before any of them is tested: with one load
  // per iteration the loop was a chain of DRAM latencies (1.5 ms for a 192^3 batch of 4)
  constexpr int UW = 4;
  const long long stride = (long long)gridDim.x * SH_T;
  for (long long base = (long long)blockIdx.x * SH_T + threadIdx.x; base < total_pad; base += stride * UW) {
    float xs[UW];
#pragma unroll
    for (int u = 0; u < UW; ++u) {
      const long long idx = base + u * stride;
      const long long q = idx / Q;
      xs[u] = (idx < total_pad && q < g.rows) ? ld1(x + q * ldx) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UW; ++u) {
      const long long idx = base + u * stride;
      if (idx >= total_pad) break;                       // warp-uniform (total_pad and stride are multiples of 32)
      const float xv = xs[u];
      if (!__any_sync(FULL, xv != 0.f)) continue;
      if (xv == 0.f) continue;
      const long long q = idx / Q;
      const int quad = (int)(idx % Q);
      int nn, d, h, w;
      decode(q, g, nn, d, h, w);
#pragma unroll
      for (int kd = 0; kd < KD; ++kd) {
        const int pz = d - (kd - g.pd);
        if (pz < 0 || pz >= g.D) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int py = h - (kh - 1);
          if (py < 0 || py >= g.H) continue;
          const TG* row = dy + (((long long)nn * g.D + pz) * g.H + py) * g.W * ldy + quad * 4;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int pw = w - (kw - 1);
            if (pw < 0 || pw >= g.W) continue;
            const float4 gv = ld4(row + pw * ldy);
            const int t = (kd * 3 + kh) * 3 + kw;
            acc[t][0] = fmaf(xv, gv.x, acc[t][0]);
            acc[t][1] = fmaf(xv, gv.y, acc[t][1]);
            acc[t][2] = fmaf(xv, gv.z, acc[t][2]);
            acc[t][3] = fmaf(xv, gv.w, acc[t][3]);
          }
        }
      }
    }
  }
  // lanes with equal (lane % Q) hold the same channel quad
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[t][j];
#pragma unroll
      for (int o = 16; o >= Q; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      acc[t][j] = v;
    }
  if (lane < Q) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][t * CO + lane * 4 + j] = acc[t][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * CO; i += SH_T) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < SH_T / 32; ++wv) s += red[wv][i];
    partial[(long long)blockIdx.x * (NT * CO) + i] = s;
  }
}

__global__ void sh_partial_reduce_kernel(const float* __restrict__ partial, int nw, int nblocks,
                                         float* __restrict__ dw, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)partial[(long long)b * nw + i];
  dw[i] = accumulate ? dw[i] + (float)s : (float)s;
}

// ------------------------------------------------------------------------------------
// head forward
// ------------------------------------------------------------------------------------
template <typename T, int N>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 f = ld4(p + i);
    v[i] = f.x; v[i + 1] = f.y; v[i + 2] = f.z; v[i + 3] = f.w;
  }
}
template <typename TI, int KD, int CI, int NC>
__global__ void __launch_bounds__(SH_T)
head_fwd_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ w,
                float* __restrict__ y, long long ldy, const Geo g) {
  constexpr int NT = KD * 9;
  __shared__ __align__(16) float wsm[NT * CI * 4];       // [t][ci][4] (NC padded to 4)
  for (int i = threadIdx.x; i < NT * CI * 4; i += SH_T) {
    const int k = i & 3, tc = i >> 2;
    wsm[i] = k < NC ? w[tc * NC + k] : 0.f;
  }
  __syncthreads();
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
  for (long long p = (long long)blockIdx.x * SH_T + threadIdx.x; p < g.rows;
       p += (long long)gridDim.x * SH_T) {
    int nn, d, h, w0;
    decode(p, g, nn, d, h, w0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kd = 0; kd < KD; ++kd) {
      const int zz = d + kd - g.pd;
      if (zz < 0 || zz >= g.D) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = h + kh - 1;
        if (yy < 0 || yy >= g.H) continue;
        const TI* row = x + (((long long)nn * g.D + zz) * g.H + yy) * g.W * ldx;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ww = w0 + kw - 1;
          if (ww < 0 || ww >= g.W) continue;
          const int t = (kd * 3 + kh) * 3 + kw;
          float xv[CI];
#pragma unroll
          for (int c = 0; c < CI; c += 8) load8(row + ww * ldx + c, xv + c);
#pragma unroll
          for (int c = 0; c < CI; ++c) {
            const float4 wv = w4[t * CI + c];
            acc[0] = fmaf(xv[c], wv.x, acc[0]);
            acc[1] = fmaf(xv[c], wv.y, acc[1]);
            if (NC > 2) acc[2] = fmaf(xv[c], wv.z, acc[2]);
            if (NC > 3) acc[3] = fmaf(xv[c], wv.w, acc[3]);
          }
        }
      }
    }
    float* o = y + p * ldy;
#pragma unroll
    for (int k = 0; k < NC; ++k) o[k] = acc[k];
  }
}

// ------------------------------------------------------------------------------------
// head input gradient
// ------------------------------------------------------------------------------------
template <typename TG, int KD, int CI, int NC>
__global__ void __launch_bounds__(SH_T)
head_dgrad_kernel(const TG* __restrict__ dy, long long ldy, const float* __restrict__ w,
                  TG* __restrict__ dx, long long ldx, int accumulate, const Geo g) {
  constexpr int NT = KD * 9;
  __shared__ __align__(16) float wsm[NT * NC * CI];      // [t][k][ci]
  for (int i = threadIdx.x; i < NT * NC * CI; i += SH_T) {
    const int ci = i % CI, k = (i / CI) % NC, t = i / (CI * NC);
    wsm[i] = w[(t * CI + ci) * NC + k];
  }
  __syncthreads();
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
  for (long long q = (long long)blockIdx.x * SH_T + threadIdx.x; q < g.rows;
       q += (long long)gridDim.x * SH_T) {
    int nn, d, h, w0;
    decode(q, g, nn, d, h, w0);
    float acc[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) acc[c] = 0.f;
#pragma unroll
    for (int kd = 0; kd < KD; ++kd) {
      const int pz = d - (kd - g.pd);
      if (pz < 0 || pz >= g.D) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int py = h - (kh - 1);
        if (py < 0 || py >= g.H) continue;
        const TG* row = dy + (((long long)nn * g.D + pz) * g.H + py) * g.W * ldy;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int pw = w0 - (kw - 1);
          if (pw < 0 || pw >= g.W) continue;
          const int t = (kd * 3 + kh) * 3 + kw;
          float gv[NC];
#pragma unroll
          for (int k = 0; k < NC; ++k) gv[k] = ld1(row + pw * ldy + k);
#pragma unroll
          for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int c4 = 0; c4 < CI / 4; ++c4) {
              const float4 wv = w4[(t * NC + k) * (CI / 4) + c4];
              acc[c4 * 4 + 0] = fmaf(gv[k], wv.x, acc[c4 * 4 + 0]);
              acc[c4 * 4 + 1] = fmaf(gv[k], wv.y, acc[c4 * 4 + 1]);
              acc[c4 * 4 + 2] = fmaf(gv[k], wv.z, acc[c4 * 4 + 2]);
              acc[c4 * 4 + 3] = fmaf(gv[k], wv.w, acc[c4 * 4 + 3]);
            }
        }
      }
    }
    TG* o = dx + q * ldx;
#pragma unroll
    for (int c = 0; c < CI; c += 4) {
      float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
      if (accumulate) {
        const float4 old = ld4(o + c);
        v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
      }
      st4(o + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------
// head filter gradient: blockIdx.y = (kd, kh) tap row; thread = output pixel p
// ------------------------------------------------------------------------------------
template <typename TI, typename TG, int KD, int CI, int NC>
__global__ void __launch_bounds__(SH_T)
head_wgrad_kernel(const TI* __restrict__ x, long long ldx, const TG* __restrict__ dy, long long ldy,
                  float* __restrict__ partial, const Geo g) {
  constexpr int NA = 3 * CI * NC;
  __shared__ float red[SH_T / 32][NA];
  const int kd = blockIdx.y / 3, kh = blockIdx.y % 3;
  float acc[3][CI][NC];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < CI; ++c)
#pragma unroll
      for (int k = 0; k < NC; ++k) acc[a][c][k] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long p = (long long)blockIdx.x * SH_T + threadIdx.x; p < g.rows;
       p += (long long)gridDim.x * SH_T) {
    int nn, d, h, w0;
    decode(p, g, nn, d, h, w0);
    const int zz = d + kd - g.pd, yy = h + kh - 1;
    if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H) continue;
    float gv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) gv[k] = ld1(dy + p * ldy + k);
    const TI* row = x + (((long long)nn * g.D + zz) * g.H + yy) * g.W * ldx;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ww = w0 + kw - 1;
      if (ww < 0 || ww >= g.W) continue;
      float xv[CI];
#pragma unroll
      for (int c = 0; c < CI; c += 8) load8(row + ww * ldx + c, xv + c);
#pragma unroll
      for (int c = 0; c < CI; ++c)
#pragma unroll
        for (int k = 0; k < NC; ++k) acc[kw][c][k] = fmaf(xv[c], gv[k], acc[kw][c][k]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < CI; ++c)
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float v = warp_sum(acc[a][c][k]);
        if (lane == 0) red[warp][(a * CI + c) * NC + k] = v;
      }
  __syncthreads();
  // partial layout: [blockIdx.x][tap][ci][k]
  const int tap0 = (kd * 3 + kh) * 3;
  float* dst = partial + ((long long)blockIdx.x * (KD * 9) + tap0) * (CI * NC);
  for (int i = threadIdx.x; i < NA; i += SH_T) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < SH_T / 32; ++wv) s += red[wv][i];
    dst[i] = s;
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
bool base_ok(const urn_conv_desc& d) {
  return d.s == 1 && d.kh == 3 && d.kw == 3 && (d.kd == 1 || d.kd == 3) && d.xd == d.yd && d.xh == d.yh &&
         d.xw == d.yw && d.pad_h == 1 && d.pad_w == 1 && d.pad_d == (d.kd == 3 ? 1 : 0) &&
         (long long)d.n * d.xd * d.xh * d.xw < (1ll << 40);
}
Geo geo(const urn_conv_desc& d) {
  Geo g;
  g.n = d.n; g.D = d.xd; g.H = d.xh; g.W = d.xw; g.kd = d.kd; g.pd = d.pad_d;
  g.rows = (long long)d.n * d.xd * d.xh * d.xw;
  return g;
}
int grid_for(long long items, int per_sm) {
  const long long want = (items + SH_T - 1) / SH_T;
  return (int)std::max<long long>(1, std::min<long long>(want, (long long)num_sms() * per_sm));
}
int wgrad_blocks(long long rows) { return grid_for(rows, 8); }

}  // namespace

bool stem_supported(const urn_conv_desc& d, int64_t ldx, int64_t ldy) {
  return base_ok(d) && d.cx == 1 && (d.cy == 8 || d.cy == 16) && d.xw % 4 == 0 && ldy == d.cy && ldx >= 1;
}
bool head_supported(const urn_conv_desc& d, int64_t ldx, int64_t ldy) {
  return base_ok(d) && (d.cx == 8 || d.cx == 16) && d.cy >= 2 && d.cy <= 4 && ldx % 8 == 0 && ldy >= d.cy;
}

int64_t stemhead_wgrad_workspace(const urn_conv_desc& d) {
  if (!base_ok(d)) return 0;
  const long long rows = (long long)d.n * d.xd * d.xh * d.xw;
  if (d.cx == 1 && (d.cy == 8 || d.cy == 16))
    return (int64_t)wgrad_blocks(rows * (d.cy / 4)) * d.kd * 9 * d.cy * 4;
  if ((d.cx == 8 || d.cx == 16) && d.cy >= 2 && d.cy <= 4)
    return (int64_t)wgrad_blocks(rows) * d.kd * 9 * d.cx * d.cy * 4;
  return 0;
}

#define SH_KD(KDV, ...) do { if (d.kd == 1) { constexpr int KDV = 1; __VA_ARGS__; } else { constexpr int KDV = 3; __VA_ARGS__; } } while (0)

int stem_fwd_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, const float* w,
                    float* y, int64_t ldy, cudaStream_t st) {
  URN_CHECK(stem_supported(d, ldx, ldy), "stem_fwd: unsupported geometry");
  const Geo g = geo(d);
  const int grid = grid_for(g.rows / 4, 8);
  return dispatch_dtype(x_dtype, [&](auto pt) {
    using TI = URN_T(pt);
    SH_KD(KD, {
      if (d.cy == 16) stem_fwd_kernel<TI, KD, 16><<<grid, SH_T, 0, st>>>((const TI*)x, ldx, w, y, ldy, g);
      else stem_fwd_kernel<TI, KD, 8><<<grid, SH_T, 0, st>>>((const TI*)x, ldx, w, y, ldy, g);
    });
    URN_LAUNCH_CHECK();
    return 0;
  });
}

int stem_wgrad_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int g_dtype,
                      const void* dy, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  URN_CHECK(stem_supported(d, ldx, d.cy) && ldy % 4 == 0, "stem_wgrad: unsupported geometry");
  const Geo g = geo(d);
  const int nb = wgrad_blocks(g.rows * (d.cy / 4));
  const int nw = d.kd * 9 * d.cy;
  float* partial = (float*)ws;
  int rc = dispatch_dtype(x_dtype, [&](auto pt) {
    using TI = URN_T(pt);
    return dispatch_dtype(g_dtype, [&](auto pg) {
      using TG = URN_T(pg);
      SH_KD(KD, {
        if (d.cy == 16) stem_wgrad_kernel<TI, TG, KD, 16><<<nb, SH_T, 0, st>>>((const TI*)x, ldx, (const TG*)dy, ldy, partial, g);
        else stem_wgrad_kernel<TI, TG, KD, 8><<<nb, SH_T, 0, st>>>((const TI*)x, ldx, (const TG*)dy, ldy, partial, g);
      });
      URN_LAUNCH_CHECK();
      return 0;
    });
  });
  if (rc) return rc;
  sh_partial_reduce_kernel<<<(nw + 127) / 128, 128, 0, st>>>(partial, nw, nb, dw, accumulate);
  URN_LAUNCH_CHECK();
  return 0;
}

#define SH_HEAD(CIV, NCV, ...)                                                                     \
  do {                                                                                             \
    if (d.cx == 16 && d.cy == 3) { constexpr int CIV = 16, NCV = 3; __VA_ARGS__; }                 \
    else if (d.cx == 16 && d.cy == 2) { constexpr int CIV = 16, NCV = 2; __VA_ARGS__; }            \
    else if (d.cx == 16 && d.cy == 4) { constexpr int CIV = 16, NCV = 4; __VA_ARGS__; }            \
    else if (d.cx == 8 && d.cy == 3) { constexpr int CIV = 8, NCV = 3; __VA_ARGS__; }              \
    else if (d.cx == 8 && d.cy == 2) { constexpr int CIV = 8, NCV = 2; __VA_ARGS__; }              \
    else { constexpr int CIV = 8, NCV = 4; __VA_ARGS__; }                                          \
  } while (0)

int head_fwd_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, const float* w,
                    float* y, int64_t ldy, cudaStream_t st) {
  URN_CHECK(head_supported(d, ldx, ldy), "head_fwd: unsupported geometry");
  const Geo g = geo(d);
  const int grid = grid_for(g.rows, 8);
  return dispatch_dtype(x_dtype, [&](auto pt) {
    using TI = URN_T(pt);
    SH_KD(KD, SH_HEAD(CI, NC, head_fwd_kernel<TI, KD, CI, NC><<<grid, SH_T, 0, st>>>((const TI*)x, ldx, w, y, ldy, g)));
    URN_LAUNCH_CHECK();
    return 0;
  });
}

int head_dgrad_launch(const urn_conv_desc& d, int g_dtype, const void* dy, int64_t ldy, const float* w,
                      void* dx, int64_t ldx, int accumulate, cudaStream_t st) {
  URN_CHECK(head_supported(d, ldx, ldy) && ldx % 4 == 0, "head_dgrad: unsupported geometry");
  const Geo g = geo(d);
  const int grid = grid_for(g.rows, 8);
  return dispatch_dtype(g_dtype, [&](auto pg) {
    using TG = URN_T(pg);
    SH_KD(KD, SH_HEAD(CI, NC, head_dgrad_kernel<TG, KD, CI, NC><<<grid, SH_T, 0, st>>>((const TG*)dy, ldy, w, (TG*)dx, ldx, accumulate, g)));
    URN_LAUNCH_CHECK();
    return 0;
  });
}

int head_wgrad_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int g_dtype,
                      const void* dy, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  URN_CHECK(head_supported(d, ldx, ldy), "head_wgrad: unsupported geometry");
  const Geo g = geo(d);
  const int nb = wgrad_blocks(g.rows);
  const int nw = d.kd * 9 * d.cx * d.cy;
  float* partial = (float*)ws;
  const dim3 grid((unsigned)nb, (unsigned)(d.kd * 3));
  int rc = dispatch_dtype(x_dtype, [&](auto pt) {
    using TI = URN_T(pt);
    return dispatch_dtype(g_dtype, [&](auto pg) {
      using TG = URN_T(pg);
      SH_KD(KD, SH_HEAD(CI, NC, head_wgrad_kernel<TI, TG, KD, CI, NC><<<grid, SH_T, 0, st>>>((const TI*)x, ldx, (const TG*)dy, ldy, partial, g)));
      URN_LAUNCH_CHECK();
      return 0;
    });
  });
  if (rc) return rc;
  sh_partial_reduce_kernel<<<(nw + 127) / 128, 128, 0, st>>>(partial, nw, nb, dw, accumulate);
  URN_LAUNCH_CHECK();
  return 0;
}
