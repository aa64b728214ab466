// Batch-norm statistics / apply / backward.
// Bandwidth-bound kernels: vectorised (4 channels per thread) coalesced row-major
// access over [rows, c] channel-last tensors, fp32 math, per-block partials reduced
// in a fixed order (deterministic) in double.
#include "common.cuh"
#include "bn_fast.cuh"
#include "gconv_tc.cuh"

namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_BLOCKS = 148 * 4;

struct ColMap {
  int vec;     // 4 or 1
  int cg;      // column groups = c / vec
  int lanes;   // row lanes per block
};
__host__ __device__ inline ColMap col_map(int c, bool aligned) {
  ColMap m;
  m.vec = aligned ? 4 : 1;
  m.cg = c / m.vec;
  m.lanes = BN_THREADS / m.cg;
  if (m.lanes < 1) m.lanes = 1;
  return m;
}

// Reduce K per-thread partial vectors (vec channels each) over the row lanes of the
// block and write them as doubles: partial[block][k][c].
template <int K>
__device__ __forceinline__ void block_col_reduce(const float (&acc)[K][4], const ColMap& m, int c,
                                                 double* __restrict__ partial) {
  __shared__ float red[BN_THREADS * 4];
  const int tid = threadIdx.x;
  const int lane = tid / m.cg, g = tid % m.cg;
  const bool active = lane < m.lanes && tid < m.lanes * m.cg;
  for (int k = 0; k < K; ++k) {
    __syncthreads();
    if (active)
      for (int v = 0; v < m.vec; ++v) red[lane * c + g * m.vec + v] = acc[k][v];
    __syncthreads();
    for (int ch = tid; ch < c; ch += BN_THREADS) {
      double s = 0.0;
      for (int l = 0; l < m.lanes; ++l) s += (double)red[l * c + ch];
      partial[((long long)blockIdx.x * K + k) * c + ch] = s;
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float* __restrict__ raw, long long ld, long long rows, int c, bool aligned,
                double* __restrict__ partial) {
  const ColMap m = col_map(c, aligned);
  const int tid = threadIdx.x;
  const int lane = tid / m.cg, g = tid % m.cg;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  if (lane < m.lanes) {
    // Shifted sums: accumulate (x-K), (x-K)^2 with K = the channel's value in row 0
    // (a background pixel for LArTPC images) so that var = E[(x-K)^2] - E[x-K]^2 does not
    // cancel catastrophically on nearly-constant channels.
    float K[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v = 0; v < m.vec; ++v) K[v] = raw[g * m.vec + v];
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * m.lanes + lane; r < rows;
         r += (long long)gridDim.x * m.lanes) {
      const float* p = raw + r * ld + g * m.vec;
      if (m.vec == 4) {
        float4 v = ld4(p);
        v.x -= K[0]; v.y -= K[1]; v.z -= K[2]; v.w -= K[3];
        acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
        acc[1][0] = fmaf(v.x, v.x, acc[1][0]); acc[1][1] = fmaf(v.y, v.y, acc[1][1]);
        acc[1][2] = fmaf(v.z, v.z, acc[1][2]); acc[1][3] = fmaf(v.w, v.w, acc[1][3]);
      } else {
        float v = *p - K[0];
        acc[0][0] += v;
        acc[1][0] = fmaf(v, v, acc[1][0]);
      }
    }
  }
  block_col_reduce<2>(acc, m, c, partial);
}

// Deterministic parallel second stage: one block per channel, thread t adds partials
// t, t+128, ... in double, then a fixed-shape shared-memory tree.
template <int K>
__device__ __forceinline__ void channel_partial_sum(const double* __restrict__ partial, int nblocks, int c,
                                                    int ch, double (&out)[K]) {
  __shared__ double sh[K][128];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 128)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += partial[((long long)b * K + k) * c + ch];
#pragma unroll
  for (int k = 0; k < K; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int k = 0; k < K; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = sh[k][0];
}

template <typename TW>
__global__ void __launch_bounds__(128)
bn_stats_finalize_kernel(const double* __restrict__ partial, int nblocks, int c, long long rows,
                         float eps, const TW* __restrict__ raw, float* __restrict__ stats) {
  const int ch = blockIdx.x;
  double v[2];
  channel_partial_sum<2>(partial, nblocks, c, ch, v);
  if (threadIdx.x == 0) {
    const double dm = v[0] / (double)rows;                 // mean of (x - K)
    double var = v[1] / (double)rows - dm * dm;
    if (var < 0.0) var = 0.0;
    stats[ch] = (float)((double)ld1(raw + ch) + dm);
    stats[c + ch] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// Merge of (n, mean, M2) partials produced by the conv epilogues (gconv_halo.cu): Chan's parallel
// update in double, thread t takes parts t, t+128, ... then a fixed-shape tree (deterministic).
__device__ __forceinline__ void chan_merge(double& n, double& mean, double& m2, double nb, double mb, double m2b) {
  if (nb <= 0.0) return;
  const double nn = n + nb, delta = mb - mean;
  mean += delta * (nb / nn);
  m2 += m2b + delta * delta * (n * nb / nn);
  n = nn;
}
__global__ void __launch_bounds__(128)
bn_stats_merge_kernel(const float* __restrict__ parts, int nparts, int c_stride, float eps,
                      float* __restrict__ stats, int c) {
  __shared__ double sh[3][128];
  const int ch = blockIdx.x;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int pi = threadIdx.x; pi < nparts; pi += 128) {
    const float* q = parts + ((long long)pi * c_stride + ch) * 3;
    chan_merge(n, mean, m2, (double)q[0], (double)q[1], (double)q[2]);
  }
  sh[0][threadIdx.x] = n; sh[1][threadIdx.x] = mean; sh[2][threadIdx.x] = m2;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double a = sh[0][threadIdx.x], b = sh[1][threadIdx.x], c2 = sh[2][threadIdx.x];
      chan_merge(a, b, c2, sh[0][threadIdx.x + o], sh[1][threadIdx.x + o], sh[2][threadIdx.x + o]);
      sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c2;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double var = sh[0][0] > 0.0 ? sh[2][0] / sh[0][0] : 0.0;
    stats[ch] = (float)sh[1][0];
    stats[c + ch] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename TR, typename TO>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ raw, long long ld, long long rows, int c, int vec,
                const float* __restrict__ stats, const float* __restrict__ beta,
                const TR* __restrict__ res, long long ldres, const float* __restrict__ res_raw,
                long long ldres_raw, const float* __restrict__ res_stats,
                const float* __restrict__ res_beta, int relu, TO* __restrict__ out, long long ldout,
                bf16* __restrict__ out2) {
  const int cg = c / vec;
  const long long total = rows * cg;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int ch = (int)(i % cg) * vec;
    if (vec == 4) {
      float4 x = ld4(raw + r * ld + ch);
      float4 mu = ld4(stats + ch), rs = ld4(stats + c + ch), be = ld4(beta + ch);
      float4 y;
      y.x = (x.x - mu.x) * rs.x + be.x; y.y = (x.y - mu.y) * rs.y + be.y;
      y.z = (x.z - mu.z) * rs.z + be.z; y.w = (x.w - mu.w) * rs.w + be.w;
      if (res_raw) {
        float4 x2 = ld4(res_raw + r * ldres_raw + ch);
        float4 mu2 = ld4(res_stats + ch), rs2 = ld4(res_stats + c + ch), be2 = ld4(res_beta + ch);
        y.x += (x2.x - mu2.x) * rs2.x + be2.x; y.y += (x2.y - mu2.y) * rs2.y + be2.y;
        y.z += (x2.z - mu2.z) * rs2.z + be2.z; y.w += (x2.w - mu2.w) * rs2.w + be2.w;
      } else if (res) {
        float4 x2 = ld4(res + r * ldres + ch);
        y.x += x2.x; y.y += x2.y; y.z += x2.z; y.w += x2.w;
      }
      if (relu) {
        y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
      }
      st4(out + r * ldout + ch, y);
      if (out2) st4(out2 + r * ldout + ch, y);
    } else {
      float y = (raw[r * ld + ch] - stats[ch]) * stats[c + ch] + beta[ch];
      if (res_raw) y += (res_raw[r * ldres_raw + ch] - res_stats[ch]) * res_stats[c + ch] + res_beta[ch];
      else if (res) y += ld1(res + r * ldres + ch);
      if (relu) y = fmaxf(y, 0.f);
      st1(out + r * ldout + ch, y);
      if (out2) st1(out2 + r * ldout + ch, y);
    }
  }
}

// ---- backward: pass 1, per-channel sums of g, g*xhat (and g*xhat2) ----
template <typename T, typename TA>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(const T* __restrict__ dout, long long lddout, const TA* __restrict__ outp,
                     long long ldo, int relu, long long rows, int c, bool aligned,
                     const float* __restrict__ raw, long long ldraw, const float* __restrict__ stats,
                     const float* __restrict__ raw2, long long ldraw2,
                     const float* __restrict__ stats2, double* __restrict__ partial) {
  const ColMap m = col_map(c, aligned);
  const int tid = threadIdx.x;
  const int lane = tid / m.cg, g = tid % m.cg;
  float acc[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  if (lane < m.lanes) {
    const int ch = g * m.vec;
    float mu[4], rs[4], mu2[4] = {0, 0, 0, 0}, rs2[4] = {0, 0, 0, 0};
    for (int v = 0; v < m.vec; ++v) {
      mu[v] = stats[ch + v];
      rs[v] = stats[c + ch + v];
      if (raw2) { mu2[v] = stats2[ch + v]; rs2[v] = stats2[c + ch + v]; }
    }
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * m.lanes + lane; r < rows;
         r += (long long)gridDim.x * m.lanes) {
      float gv[4], xv[4], x2v[4] = {0, 0, 0, 0};
      if (m.vec == 4) {
        float4 d4 = ld4(dout + r * lddout + ch);
        gv[0] = d4.x; gv[1] = d4.y; gv[2] = d4.z; gv[3] = d4.w;
        if (relu) {
          float4 o4 = ld4(outp + r * ldo + ch);
          if (!(o4.x > 0.f)) gv[0] = 0.f;
          if (!(o4.y > 0.f)) gv[1] = 0.f;
          if (!(o4.z > 0.f)) gv[2] = 0.f;
          if (!(o4.w > 0.f)) gv[3] = 0.f;
        }
        float4 x4 = ld4(raw + r * ldraw + ch);
        xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
        if (raw2) {
          float4 y4 = ld4(raw2 + r * ldraw2 + ch);
          x2v[0] = y4.x; x2v[1] = y4.y; x2v[2] = y4.z; x2v[3] = y4.w;
        }
      } else {
        gv[0] = ld1(dout + r * lddout + ch);
        if (relu && !(ld1(outp + r * ldo + ch) > 0.f)) gv[0] = 0.f;
        xv[0] = raw[r * ldraw + ch];
        if (raw2) x2v[0] = raw2[r * ldraw2 + ch];
      }
      for (int v = 0; v < m.vec; ++v) {
        acc[0][v] += gv[v];
        acc[1][v] = fmaf(gv[v], (xv[v] - mu[v]) * rs[v], acc[1][v]);
        if (raw2) acc[2][v] = fmaf(gv[v], (x2v[v] - mu2[v]) * rs2[v], acc[2][v]);
      }
    }
  }
  block_col_reduce<3>(acc, m, c, partial);
}

// sums[0..c) = sum g ; [c..2c) = sum g*xhat ; [2c..3c) = sum g*xhat2 ; also emits dbeta(2)
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const double* __restrict__ partial, int nblocks, int c, float* __restrict__ sums,
                       float* __restrict__ dbeta, float* __restrict__ dbeta2, int acc_beta) {
  const int ch = blockIdx.x;
  double v[3];
  channel_partial_sum<3>(partial, nblocks, c, ch, v);
  if (threadIdx.x == 0) {
    sums[ch] = (float)v[0];
    sums[c + ch] = (float)v[1];
    sums[2 * c + ch] = (float)v[2];
    if (dbeta) dbeta[ch] = acc_beta ? dbeta[ch] + (float)v[0] : (float)v[0];
    if (dbeta2) dbeta2[ch] = acc_beta ? dbeta2[ch] + (float)v[0] : (float)v[0];
  }
}

// ---- backward: pass 2, elementwise ----
template <typename T, typename TA>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dout, long long lddout, const TA* __restrict__ outp,
                    long long ldo, int relu, long long rows, int c, int vec,
                    const float* __restrict__ raw, long long ldraw, const float* __restrict__ stats,
                    T* __restrict__ draw, long long lddraw, const float* __restrict__ raw2,
                    long long ldraw2, const float* __restrict__ stats2, T* __restrict__ draw2,
                    long long lddraw2, T* __restrict__ dres, long long lddres, int dres_acc,
                    const float* __restrict__ sums) {
  const int cg = c / vec;
  const long long total = rows * cg;
  const float inv = 1.0f / (float)rows;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int ch = (int)(i % cg) * vec;
    float gv[4], xv[4], x2v[4] = {0, 0, 0, 0};
    if (vec == 4) {
      float4 d4 = ld4(dout + r * lddout + ch);
      gv[0] = d4.x; gv[1] = d4.y; gv[2] = d4.z; gv[3] = d4.w;
      if (relu) {
        float4 o4 = ld4(outp + r * ldo + ch);
        if (!(o4.x > 0.f)) gv[0] = 0.f;
        if (!(o4.y > 0.f)) gv[1] = 0.f;
        if (!(o4.z > 0.f)) gv[2] = 0.f;
        if (!(o4.w > 0.f)) gv[3] = 0.f;
      }
      float4 x4 = ld4(raw + r * ldraw + ch);
      xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
      if (raw2) {
        float4 y4 = ld4(raw2 + r * ldraw2 + ch);
        x2v[0] = y4.x; x2v[1] = y4.y; x2v[2] = y4.z; x2v[3] = y4.w;
      }
    } else {
      gv[0] = ld1(dout + r * lddout + ch);
      if (relu && !(ld1(outp + r * ldo + ch) > 0.f)) gv[0] = 0.f;
      xv[0] = raw[r * ldraw + ch];
      if (raw2) x2v[0] = raw2[r * ldraw2 + ch];
    }
    float o1[4], o2[4];
    for (int v = 0; v < vec; ++v) {
      const float mg = sums[ch + v] * inv;
      const float rs = stats[c + ch + v];
      const float xh = (xv[v] - stats[ch + v]) * rs;
      o1[v] = rs * (gv[v] - mg - xh * (sums[c + ch + v] * inv));
      if (raw2) {
        const float rs2 = stats2[c + ch + v];
        const float xh2 = (x2v[v] - stats2[ch + v]) * rs2;
        o2[v] = rs2 * (gv[v] - mg - xh2 * (sums[2 * c + ch + v] * inv));
      }
    }
    if (vec == 4) {
      st4(draw + r * lddraw + ch, make_float4(o1[0], o1[1], o1[2], o1[3]));
      if (raw2) st4(draw2 + r * lddraw2 + ch, make_float4(o2[0], o2[1], o2[2], o2[3]));
      if (dres) {
        T* p = dres + r * lddres + ch;
        float4 gg = make_float4(gv[0], gv[1], gv[2], gv[3]);
        if (dres_acc) {
          float4 o = ld4(p);
          gg.x += o.x; gg.y += o.y; gg.z += o.z; gg.w += o.w;
        }
        st4(p, gg);
      }
    } else {
      st1(draw + r * lddraw + ch, o1[0]);
      if (raw2) st1(draw2 + r * lddraw2 + ch, o2[0]);
      if (dres) {
        T* p = dres + r * lddres + ch;
        st1(p, gv[0] + (dres_acc ? ld1(p) : 0.f));
      }
    }
  }
}

inline int stat_blocks(long long rows, const ColMap& m) {
  long long b = (rows + m.lanes - 1) / m.lanes;
  if (b > BN_MAX_BLOCKS) b = BN_MAX_BLOCKS;
  if (b < 1) b = 1;
  return (int)b;
}
inline unsigned ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148ll * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}
inline bool aligned4(int c, std::initializer_list<long long> lds) {
  if (c % 4) return false;
  for (long long l : lds)
    if (l % 4) return false;
  return true;
}

}  // namespace

int bn_stats_merge_launch(const float* parts, int nparts, int c_stride, int c, long long rows, float eps,
                          float* stats, cudaStream_t st) {
  (void)rows;
  bn_stats_merge_kernel<<<c, 128, 0, st>>>(parts, nparts, c_stride, eps, stats, c);
  URN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t urn_bn_workspace(int c) {
  // partials [BN_MAX_BLOCKS][3][c] doubles + sums [3][c] floats (256 B aligned)
  int64_t b = (int64_t)BN_MAX_BLOCKS * 3 * c * 8;
  b = (b + 255) / 256 * 256;
  return b + (int64_t)3 * c * 4 + 256;
}

int bn_stats_any(const void* raw_, int raw_dtype, int64_t ld, int64_t rows, int c, float eps, float* stats,
                 void* ws, cudaStream_t st) {
  URN_CHECK(c >= 1 && c <= 1024, "urn_bn_stats: c=%d out of range", c);
  URN_CHECK(rows > 0, "urn_bn_stats: empty input");
  URN_CHECK(raw_dtype == URN_F32 || raw_dtype == URN_F16, "urn_bn_stats: raw dtype %d", raw_dtype);
  if (bn_fast_shape(c, {ld}, {raw_})) {
    int nbf = 0;
    if (bn_stats_fast_launch(raw_, raw_dtype, ld, rows, c, (double*)ws, &nbf, st)) return 1;
    if (raw_dtype == URN_F16)
      bn_stats_finalize_kernel<f16><<<c, 128, 0, st>>>((const double*)ws, nbf, c, rows, eps, (const f16*)raw_, stats);
    else
      bn_stats_finalize_kernel<float><<<c, 128, 0, st>>>((const double*)ws, nbf, c, rows, eps, (const float*)raw_, stats);
    URN_LAUNCH_CHECK();
    return 0;
  }
  URN_CHECK(raw_dtype == URN_F32, "urn_bn_stats: f16 raw needs power-of-two channels >= 8 and 8-aligned pitches (c=%d)", c);
  const float* raw = (const float*)raw_;
  bool al = aligned4(c, {ld}) && ((uintptr_t)raw % 16 == 0);
  ColMap m = col_map(c, al);
  URN_CHECK(m.cg <= BN_THREADS, "urn_bn_stats: c=%d needs 4-aligned pitch", c);
  int nb = stat_blocks(rows, m);
  double* partial = (double*)ws;
  bn_stats_kernel<<<nb, BN_THREADS, 0, st>>>(raw, ld, rows, c, al, partial);
  URN_LAUNCH_CHECK();
  bn_stats_finalize_kernel<float><<<c, 128, 0, st>>>(partial, nb, c, rows, eps, raw, stats);
  URN_LAUNCH_CHECK();
  return 0;
}

extern "C" int urn_bn_stats(const float* raw, int64_t ld, int64_t rows, int c, float eps,
                            float* stats, void* ws, void* stream) {
  return bn_stats_any(raw, URN_F32, ld, rows, c, eps, stats, ws, (cudaStream_t)stream);
}
extern "C" int urn_bn_stats_t(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, float eps,
                              float* stats, void* ws, void* stream) {
  return bn_stats_any(raw, raw_dtype, ld, rows, c, eps, stats, ws, (cudaStream_t)stream);
}

template <typename TR, typename TO>
static int bn_apply_t(const float* raw, int64_t ld, int64_t rows, int c, const float* stats,
                      const float* beta, const void* res, int64_t ldres, const float* res_raw,
                      int64_t ldres_raw, const float* res_stats, const float* res_beta, int relu,
                      void* out, int64_t ldout, void* out2, cudaStream_t st) {
  bool al = aligned4(c, {ld, ldout, res ? ldres : 0, res_raw ? ldres_raw : 0});
  int vec = al ? 4 : 1;
  long long total = rows * (c / vec);
  bn_apply_kernel<TR, TO><<<ew_blocks(total), 256, 0, st>>>(
      raw, ld, rows, c, vec, stats, beta, (const TR*)res, ldres, res_raw, ldres_raw, res_stats,
      res_beta, relu, (TO*)out, ldout, (bf16*)out2);
  URN_LAUNCH_CHECK();
  return 0;
}

int bn_apply_any(const void* raw_, int raw_dtype, int64_t ld, int64_t rows, int c, const float* stats,
                 const float* beta, const void* res, int64_t ldres, int res_dtype, const void* res_raw_,
                 int64_t ldres_raw, const float* res_stats, const float* res_beta, int relu, void* out,
                 int64_t ldout, int out_dtype, void* out2, cudaStream_t st) {
  URN_CHECK(rows > 0 && c >= 1, "urn_bn_apply: empty input");
  URN_CHECK(raw_dtype == URN_F32 || raw_dtype == URN_F16, "urn_bn_apply: raw dtype %d", raw_dtype);
  if (bn_fast_shape(c, {ld, ldout, res ? ldres : 0, res_raw_ ? ldres_raw : 0}, {raw_, out, out2, res, res_raw_}))
    return bn_apply_fast_launch(raw_, raw_dtype, ld, rows, c, stats, beta, res, ldres, res_dtype, res_raw_, ldres_raw,
                                res_stats, res_beta, relu, out, ldout, out_dtype, out2, st);
  URN_CHECK(raw_dtype == URN_F32, "urn_bn_apply: f16 raw needs power-of-two channels >= 8 and 8-aligned pitches (c=%d)", c);
  const float* raw = (const float*)raw_;
  const float* res_raw = (const float*)res_raw_;
  if (!res) res_dtype = out_dtype;
  return dispatch_dtype(res_dtype, [&](auto pr) {
    using TR = URN_T(pr);
    return dispatch_dtype(out_dtype, [&](auto po) {
      using TO = URN_T(po);
      return bn_apply_t<TR, TO>(raw, ld, rows, c, stats, beta, res, ldres, res_raw, ldres_raw,
                                res_stats, res_beta, relu, out, ldout, out2, st);
    });
  });
}

extern "C" int urn_bn_apply(const float* raw, int64_t ld, int64_t rows, int c, const float* stats,
                            const float* beta, const void* res, int64_t ldres, int res_dtype,
                            const float* res_raw, int64_t ldres_raw, const float* res_stats,
                            const float* res_beta, int relu, void* out, int64_t ldout,
                            int out_dtype, void* out2, void* stream) {
  return bn_apply_any(raw, URN_F32, ld, rows, c, stats, beta, res, ldres, res_dtype, res_raw, ldres_raw, res_stats,
                      res_beta, relu, out, ldout, out_dtype, out2, (cudaStream_t)stream);
}
extern "C" int urn_bn_apply_t(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, const float* stats,
                              const float* beta, const void* res, int64_t ldres, int res_dtype,
                              const void* res_raw, int64_t ldres_raw, const float* res_stats,
                              const float* res_beta, int relu, void* out, int64_t ldout,
                              int out_dtype, void* out2, void* stream) {
  return bn_apply_any(raw, raw_dtype, ld, rows, c, stats, beta, res, ldres, res_dtype, res_raw, ldres_raw, res_stats,
                      res_beta, relu, out, ldout, out_dtype, out2, (cudaStream_t)stream);
}

template <typename T, typename TA>
static int bn_bwd_t(const void* dout, int64_t lddout, const void* out, int64_t ldo, int relu,
                    int64_t rows, int c, const float* raw, int64_t ldraw, const float* stats,
                    float* dbeta, void* draw, int64_t lddraw, const float* raw2, int64_t ldraw2,
                    const float* stats2, float* dbeta2, void* draw2, int64_t lddraw2, void* dres,
                    int64_t lddres, int dres_acc, int acc_beta, void* ws, cudaStream_t st) {
  bool al = aligned4(c, {lddout, relu ? ldo : 0, ldraw, lddraw, raw2 ? ldraw2 : 0,
                         raw2 ? lddraw2 : 0, dres ? lddres : 0});
  ColMap m = col_map(c, al);
  URN_CHECK(m.cg <= BN_THREADS, "urn_bn_bwd: c=%d needs 4-aligned pitches", c);
  int nb = stat_blocks(rows, m);
  double* partial = (double*)ws;
  int64_t pb = ((int64_t)BN_MAX_BLOCKS * 3 * c * 8 + 255) / 256 * 256;
  float* sums = (float*)((char*)ws + pb);
  bn_bwd_reduce_kernel<T, TA><<<nb, BN_THREADS, 0, st>>>((const T*)dout, lddout, (const TA*)out, ldo,
                                                     relu, rows, c, al, raw, ldraw, stats, raw2,
                                                     ldraw2, stats2, partial);
  URN_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<c, 128, 0, st>>>(partial, nb, c, sums, dbeta,
                                                          raw2 ? dbeta2 : nullptr, acc_beta);
  URN_LAUNCH_CHECK();
  int vec = al ? 4 : 1;
  long long total = rows * (c / vec);
  bn_bwd_apply_kernel<T, TA><<<ew_blocks(total), 256, 0, st>>>(
      (const T*)dout, lddout, (const TA*)out, ldo, relu, rows, c, vec, raw, ldraw, stats, (T*)draw,
      lddraw, raw2, ldraw2, stats2, (T*)draw2, lddraw2, (T*)dres, lddres, dres_acc, sums);
  URN_LAUNCH_CHECK();
  return 0;
}

extern "C" int urn_bn_bwd(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
                          int out_dtype, int relu, int64_t rows, int c, const float* raw,
                          int64_t ldraw, const float* stats, float* dbeta, void* draw,
                          int64_t lddraw, const float* raw2, int64_t ldraw2, const float* stats2,
                          float* dbeta2, void* draw2, int64_t lddraw2, void* dres, int64_t lddres,
                          int dres_accumulate, int acc_beta, void* ws, void* stream) {
  return bn_bwd_full(dout, lddout, out, ldo, dtype, out_dtype, relu, rows, c, raw, URN_F32, ldraw, stats, nullptr, dbeta,
                     draw, lddraw, raw2, ldraw2, stats2, dbeta2, draw2, lddraw2, dres, lddres, dres_accumulate,
                     acc_beta, ws, (cudaStream_t)stream);
}

extern "C" int urn_bn_bwd_beta(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
                               int out_dtype, int relu, int64_t rows, int c, const float* raw,
                               int64_t ldraw, const float* stats, const float* beta, float* dbeta,
                               void* draw, int64_t lddraw, const float* raw2, int64_t ldraw2,
                               const float* stats2, float* dbeta2, void* draw2, int64_t lddraw2,
                               void* dres, int64_t lddres, int dres_accumulate, int acc_beta, void* ws,
                               void* stream) {
  return bn_bwd_full(dout, lddout, out, ldo, dtype, out_dtype, relu, rows, c, raw, URN_F32, ldraw, stats, beta, dbeta,
                     draw, lddraw, raw2, ldraw2, stats2, dbeta2, draw2, lddraw2, dres, lddres, dres_accumulate,
                     acc_beta, ws, (cudaStream_t)stream);
}
extern "C" int urn_bn_bwd_t(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
                            int out_dtype, int relu, int64_t rows, int c, const void* raw, int raw_dtype,
                            int64_t ldraw, const float* stats, const float* beta, float* dbeta,
                            void* draw, int64_t lddraw, const void* raw2, int64_t ldraw2,
                            const float* stats2, float* dbeta2, void* draw2, int64_t lddraw2,
                            void* dres, int64_t lddres, int dres_accumulate, int acc_beta, void* ws,
                            void* stream) {
  return bn_bwd_full(dout, lddout, out, ldo, dtype, out_dtype, relu, rows, c, raw, raw_dtype, ldraw, stats, beta, dbeta,
                     draw, lddraw, raw2, ldraw2, stats2, dbeta2, draw2, lddraw2, dres, lddres, dres_accumulate,
                     acc_beta, ws, (cudaStream_t)stream);
}

int bn_bwd_full(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype, int out_dtype,
                int relu, int64_t rows, int c, const void* raw_, int raw_dtype, int64_t ldraw, const float* stats,
                const float* beta, float* dbeta, void* draw, int64_t lddraw, const void* raw2_,
                int64_t ldraw2, const float* stats2, float* dbeta2, void* draw2, int64_t lddraw2, void* dres,
                int64_t lddres, int dres_accumulate, int acc_beta, void* ws, cudaStream_t st) {
  URN_CHECK(rows > 0 && c >= 1 && c <= 1024, "urn_bn_bwd: bad shape");
  URN_CHECK(raw_dtype == URN_F32 || raw_dtype == URN_F16, "urn_bn_bwd: raw dtype %d", raw_dtype);
  const float* raw = (const float*)raw_;      // generic kernels below: fp32 raw only (checked)
  const float* raw2 = (const float*)raw2_;
  URN_CHECK(!relu || out, "urn_bn_bwd: relu needs the forward output");
  const bool pair_ok = (dtype == URN_F32 && out_dtype == URN_F32) || (dtype == URN_BF16 && out_dtype != URN_F32);
  if (pair_ok && out &&
      bn_fast_shape(c, {lddout, ldo, ldraw, lddraw, raw2 ? ldraw2 : 0, raw2 ? lddraw2 : 0, dres ? lddres : 0},
                    {dout, out, raw, draw, raw2, draw2, dres})) {
    double* partial = (double*)ws;
    const int64_t pb = ((int64_t)BN_MAX_BLOCKS * 3 * c * 8 + 255) / 256 * 256;
    float* sums = (float*)((char*)ws + pb);
    int nb = 0;
    if (bn_bwd_fast_launch(0, dout, lddout, out, ldo, dtype, out_dtype, relu, rows, c, raw_, raw_dtype, ldraw, stats, beta,
                           draw, lddraw, raw2_, ldraw2, stats2, draw2, lddraw2, dres, lddres, dres_accumulate,
                           partial, sums, &nb, st))
      return 1;
    bn_bwd_finalize_kernel<<<c, 128, 0, st>>>(partial, nb, c, sums, dbeta, raw2 ? dbeta2 : nullptr, acc_beta);
    URN_LAUNCH_CHECK();
    return bn_bwd_fast_launch(1, dout, lddout, out, ldo, dtype, out_dtype, relu, rows, c, raw_, raw_dtype, ldraw, stats, beta,
                              draw, lddraw, raw2_, ldraw2, stats2, draw2, lddraw2, dres, lddres, dres_accumulate,
                              partial, sums, &nb, st);
  }
  URN_CHECK(raw_dtype == URN_F32, "urn_bn_bwd: f16 raw needs the fast-path shape (power-of-two channels >= 8, 8-aligned pitches, 16-bit gradients)");
#define URN_BB(GD, AD, TG, TAA)                                                                   \
  if (dtype == GD && out_dtype == AD)                                                             \
    return bn_bwd_t<TG, TAA>(dout, lddout, out, ldo, relu, rows, c, raw, ldraw, stats, dbeta, draw, \
                             lddraw, raw2, ldraw2, stats2, dbeta2, draw2, lddraw2, dres, lddres,  \
                             dres_accumulate, acc_beta, ws, st);
  URN_BB(URN_F32, URN_F32, float, float)
  URN_BB(URN_BF16, URN_BF16, bf16, bf16)
  URN_BB(URN_BF16, URN_F16, bf16, f16)
  URN_BB(URN_BF16, URN_F32, bf16, float)
#undef URN_BB
  urn_set_error("urn_bn_bwd: unsupported dtype pair (grad %d, act %d)", dtype, out_dtype);
  return 1;
}
