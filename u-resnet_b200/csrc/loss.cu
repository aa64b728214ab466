// Per-pixel class-weighted softmax cross-entropy, its gradient, softmax and the two
// argmax accuracies in ONE pass over the logits (SURVEY.md 8a K9, K10, K11;
// /root/reference/lib/ssnet.py:57-71).  Bandwidth-bound: thread per pixel,
// coalesced reads of label/weight/data, warp-shuffle + block reduction, per-block
// partials summed in a fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int CE_THREADS = 256;
constexpr int CE_MAX_BLOCKS = 148 * 8;
constexpr int CE_MAX_CLASSES = 32;

template <typename TD>
__global__ void __launch_bounds__(CE_THREADS)
ce_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ label,
          const float* __restrict__ weight, const float* __restrict__ data, long long total, int c,
          float inv_batch, TD* __restrict__ dlogits, long long lddl, float* __restrict__ softmax,
          int* __restrict__ err_flag, double* __restrict__ partial) {
  double loss = 0.0;
  unsigned long long n_all = 0, n_cnz = 0, n_nz = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    float z[CE_MAX_CLASSES];
    const float* lp = logits + p * ld;
    float mx = -INFINITY;
    int arg = 0;
    for (int k = 0; k < c; ++k) {
      z[k] = lp[k];
      if (z[k] > mx) { mx = z[k]; arg = k; }   // strict '>' keeps the first max on ties
    }
    float se = 0.f;
    for (int k = 0; k < c; ++k) { z[k] = expf(z[k] - mx); se += z[k]; }
    const float inv = 1.0f / se;
    const int lab = (int)label[p];               // float -> int truncation (ssnet.py:40)
    const float w = weight ? weight[p] : 1.0f;
    const bool ok = lab >= 0 && lab < c;
    if (!ok) atomicOr(err_flag, 1);
    const float zl = ok ? lp[lab] : 0.f;
    loss += ok ? (double)((mx + logf(se) - zl) * w) : (double)NAN;
    if (arg == lab) ++n_all;
    if (data[p] > 0.f) { ++n_nz; if (arg == lab) ++n_cnz; }
    if (softmax)
      for (int k = 0; k < c; ++k) softmax[p * c + k] = z[k] * inv;
    if (dlogits) {
      const float sc = w * inv_batch;
      for (int k = 0; k < c; ++k)
        st1(dlogits + p * lddl + k, (z[k] * inv - (k == lab ? 1.f : 0.f)) * sc);
    }
  }
  // block reduction
  __shared__ double s_loss[CE_THREADS / 32];
  __shared__ unsigned long long s_cnt[3][CE_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    n_all += __shfl_xor_sync(0xffffffffu, n_all, o);
    n_cnz += __shfl_xor_sync(0xffffffffu, n_cnz, o);
    n_nz += __shfl_xor_sync(0xffffffffu, n_nz, o);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_loss[wid] = loss; s_cnt[0][wid] = n_all; s_cnt[1][wid] = n_cnz; s_cnt[2][wid] = n_nz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    unsigned long long a = 0, b = 0, d = 0;
    for (int i = 0; i < CE_THREADS / 32; ++i) { l += s_loss[i]; a += s_cnt[0][i]; b += s_cnt[1][i]; d += s_cnt[2][i]; }
    double* q = partial + (long long)blockIdx.x * 4;
    q[0] = l; q[1] = (double)a; q[2] = (double)b; q[3] = (double)d;
  }
}

__global__ void __launch_bounds__(128)
ce_finalize_kernel(const double* __restrict__ partial, int nblocks, double inv_batch,
                   double* __restrict__ result) {
  // fixed-shape (deterministic) parallel sum of the per-block partials
  __shared__ double sh[4][128];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += 128)
    for (int k = 0; k < 4; ++k) acc[k] += partial[b * 4 + k];
  for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    result[0] = sh[0][0] * inv_batch; result[1] = sh[1][0]; result[2] = sh[2][0]; result[3] = sh[3][0];
  }
}

template <typename TD>
__global__ void softmax_only_kernel(const float* __restrict__ logits, long long ld, long long total,
                                    int c, float* __restrict__ softmax) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    float z[CE_MAX_CLASSES];
    const float* lp = logits + p * ld;
    float mx = -INFINITY;
    for (int k = 0; k < c; ++k) { z[k] = lp[k]; mx = fmaxf(mx, z[k]); }
    float se = 0.f;
    for (int k = 0; k < c; ++k) { z[k] = expf(z[k] - mx); se += z[k]; }
    const float inv = 1.0f / se;
    for (int k = 0; k < c; ++k) softmax[p * c + k] = z[k] * inv;
  }
}

// Shower/track decision of the analysis step (lib/ssnet_trainval.py:269-287):
//   decision = (1*(p1 > p2) + 2*(p2 >= p1)) * (data > threshold),  p = softmax(logits)
// computed from the same softmax arithmetic as softmax_only_kernel, so that it equals the host
// expression applied to the softmax this library returns, value for value (ties included).
__global__ void decision_kernel(const float* __restrict__ logits, long long ld, long long total, int c,
                                const float* __restrict__ data, long long data_ld, float threshold,
                                float* __restrict__ decision) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    float z[CE_MAX_CLASSES];
    const float* lp = logits + p * ld;
    float mx = -INFINITY;
    for (int k = 0; k < c; ++k) { z[k] = lp[k]; mx = fmaxf(mx, z[k]); }
    float se = 0.f;
    for (int k = 0; k < c; ++k) { z[k] = expf(z[k] - mx); se += z[k]; }
    const float inv = 1.0f / se;
    const float shower = z[1] * inv, track = z[2] * inv;
    const float r = (shower > track ? 1.f : 0.f) + (track >= shower ? 2.f : 0.f);
    decision[p] = data[p * data_ld] > threshold ? r : 0.f;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2,
                            float eps) {
  const long long n4 = n / 4;
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = ld4(p + i * 4), gg = ld4(g + i * 4), mm = ld4(m + i * 4), vv = ld4(v + i * 4);
    // TF ApplyAdam form: m += (g-m)*(1-b1); v += (g*g-v)*(1-b2); var -= m*alpha/(sqrt(v)+eps)
#define URN_ADAM1(c)                                   \
    mm.c += (gg.c - mm.c) * omb1;                       \
    vv.c += (gg.c * gg.c - vv.c) * omb2;                \
    pp.c -= (mm.c * lr_t) / (sqrtf(vv.c) + eps);
    URN_ADAM1(x) URN_ADAM1(y) URN_ADAM1(z) URN_ADAM1(w)
#undef URN_ADAM1
    st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float mm = m[i] + (g[i] - m[i]) * omb1;
      float vv = v[i] + (g[i] * g[i] - v[i]) * omb2;
      m[i] = mm; v[i] = vv;
      p[i] -= (mm * lr_t) / (sqrtf(vv) + eps);
    }
}

}  // namespace

extern "C" int64_t urn_loss_workspace(int64_t) { return 256 + (int64_t)CE_MAX_BLOCKS * 4 * 8; }

extern "C" int urn_ce_fwd_bwd(const float* logits, int64_t ld, const float* label,
                              const float* weight, const float* data, int64_t batch, int64_t pixels,
                              int c, void* dlogits, int64_t lddl, int dl_dtype, float* softmax,
                              double* result, void* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  URN_CHECK(c >= 1 && c <= CE_MAX_CLASSES, "urn_ce_fwd_bwd: num_class=%d unsupported (max %d)", c,
            CE_MAX_CLASSES);
  URN_CHECK(batch > 0 && pixels > 0, "urn_ce_fwd_bwd: empty input");
  const long long total = batch * pixels;
  int nb = (int)((total + CE_THREADS - 1) / CE_THREADS);
  if (nb > CE_MAX_BLOCKS) nb = CE_MAX_BLOCKS;
  int* flag = (int*)ws;
  double* partial = (double*)((char*)ws + 256);
  const float invb = 1.0f / (float)batch;
  URN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (dispatch_dtype(dl_dtype, [&](auto pd) {
        using TD = URN_T(pd);
        ce_kernel<TD><<<nb, CE_THREADS, 0, st>>>(logits, ld, label, weight, data, total, c, invb,
                                                 (TD*)dlogits, lddl, softmax, flag, partial);
        return 0;
      }))
    return 1;
  URN_LAUNCH_CHECK();
  ce_finalize_kernel<<<1, 128, 0, st>>>(partial, nb, 1.0 / (double)batch, result);
  URN_LAUNCH_CHECK();
  return 0;
}

int urn_softmax_only(const float* logits, int64_t ld, int64_t total, int c, float* softmax,
                     cudaStream_t st) {
  URN_CHECK(c >= 1 && c <= CE_MAX_CLASSES, "softmax: num_class=%d unsupported", c);
  int nb = (int)((total + 255) / 256);
  if (nb > CE_MAX_BLOCKS) nb = CE_MAX_BLOCKS;
  softmax_only_kernel<float><<<nb, 256, 0, st>>>(logits, ld, total, c, softmax);
  URN_LAUNCH_CHECK();
  return 0;
}

extern "C" int urn_ssnet_decision(const float* logits, int64_t ld, int64_t total, int num_class,
                                  const float* data, int64_t data_ld, float threshold,
                                  float* decision, void* stream) {
  URN_CHECK(logits && data && decision, "urn_ssnet_decision: null argument");
  URN_CHECK(num_class >= 3 && num_class <= CE_MAX_CLASSES,
            "urn_ssnet_decision: needs shower (1) and track (2) classes, num_class=%d", num_class);
  URN_CHECK(ld >= num_class && data_ld >= 1 && total >= 0, "urn_ssnet_decision: bad strides");
  if (total == 0) return 0;
  int nb = (int)((total + 255) / 256);
  if (nb > CE_MAX_BLOCKS) nb = CE_MAX_BLOCKS;
  decision_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(logits, ld, total, num_class, data, data_ld,
                                                         threshold, decision);
  URN_LAUNCH_CHECK();
  return 0;
}

extern "C" int urn_adam_step(float* param, const float* grad, float* m, float* v, int64_t count,
                             float lr, float beta1, float beta2, float eps, int64_t t,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  URN_CHECK(t >= 1, "urn_adam_step: t must be >= 1");
  if (count == 0) return 0;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) /
                      (1.0 - pow((double)beta1, (double)t));
  long long nb = (count / 4 + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  adam_kernel<<<(unsigned)nb, 256, 0, st>>>(param, grad, m, v, count, (float)lr_t, beta1, beta2, eps);
  URN_LAUNCH_CHECK();
  return 0;
}
