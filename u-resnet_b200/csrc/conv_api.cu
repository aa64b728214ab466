// C-ABI conv entry points: geometry -> generalised-conv launches (FFMA or tcgen05).
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"

static thread_local char g_err[1024] = "";
long long g_urn_launches = 0;
extern "C" int64_t urn_launch_count(void) { return g_urn_launches; }
void urn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* urn_last_error(void) { return g_err; }
extern "C" const char* urn_build_info(void) {
  return "uresnet_b200 sm_100a (nvcc " URN_STR(__CUDACC_VER_MAJOR__) "." URN_STR(__CUDACC_VER_MINOR__) ")";
}

static int same_pad_before(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

extern "C" int urn_conv_desc_same(urn_conv_desc* d, int nd, int n, int xd, int xh, int xw, int cx,
                                  int cy, int k, int s) {
  URN_CHECK(d, "urn_conv_desc_same: null descriptor");
  URN_CHECK(nd == 2 || nd == 3, "urn_conv_desc_same: nd must be 2 or 3");
  URN_CHECK(n > 0 && xd > 0 && xh > 0 && xw > 0 && cx > 0 && cy > 0, "urn_conv_desc_same: bad dims");
  URN_CHECK((k == 1 || k == 3) && (s == 1 || s == 2), "urn_conv_desc_same: k in {1,3}, s in {1,2}");
  URN_CHECK(nd == 3 || xd == 1, "urn_conv_desc_same: 2D needs xd == 1");
  d->n = n; d->xd = xd; d->xh = xh; d->xw = xw; d->cx = cx; d->cy = cy;
  d->kd = nd == 3 ? k : 1; d->kh = k; d->kw = k; d->s = s;
  d->yd = nd == 3 ? (xd + s - 1) / s : 1;
  d->yh = (xh + s - 1) / s;
  d->yw = (xw + s - 1) / s;
  d->pad_d = nd == 3 ? same_pad_before(xd, k, s) : 0;
  d->pad_h = same_pad_before(xh, k, s);
  d->pad_w = same_pad_before(xw, k, s);
  return 0;
}

static int check_desc(const urn_conv_desc* d) {
  URN_CHECK(d, "null conv descriptor");
  URN_CHECK(d->n > 0 && d->cx > 0 && d->cy > 0 && d->xd > 0 && d->xh > 0 && d->xw > 0 &&
                d->yd > 0 && d->yh > 0 && d->yw > 0, "conv descriptor: non-positive dims");
  URN_CHECK(d->kd * d->kh * d->kw <= 27 && d->kd >= 1 && d->kh >= 1 && d->kw >= 1,
            "conv descriptor: at most 27 taps");
  URN_CHECK(d->s == 1 || d->s == 2, "conv descriptor: stride must be 1 or 2");
  return 0;
}

// Y[o] = sum_t X[o*s + t - pad] . W[t]           (in = X, out = Y)
void gconv_gather(const urn_conv_desc& d, int64_t ldx, int64_t ldy, GConv* g) {
  g->n = d.n;
  g->id = d.xd; g->ih = d.xh; g->iw = d.xw; g->cin = d.cx;
  g->od = d.yd; g->oh = d.yh; g->ow = d.yw; g->cout = d.cy;
  g->gd = d.yd; g->gh = d.yh; g->gw = d.yw;
  g->is = d.s; g->os = 1; g->ood = g->ooh = g->oow = 0;
  g->ntaps = 0;
  for (int td = 0; td < d.kd; ++td)
    for (int th = 0; th < d.kh; ++th)
      for (int tw = 0; tw < d.kw; ++tw) {
        GTap& t = g->taps[g->ntaps++];
        t.od = td - d.pad_d; t.oh = th - d.pad_h; t.ow = tw - d.pad_w;
        t.widx = (td * d.kh + th) * d.kw + tw;
      }
  g->wtap = (int64_t)d.cx * d.cy; g->wldi = d.cy; g->wldo = 1;
  g->ldin = ldx; g->ldout = ldy; g->accumulate = 0;
}

// X[x] = sum_{o,t : o*s + t - pad = x} Y[o] . W[t]^T   (in = Y, out = X), one launch per
// output parity class (pd,ph,pw) in [0,s)^3.  Returns false if the class is empty.
bool gconv_scatter_class(const urn_conv_desc& d, int64_t ldy, int64_t ldx, int pd, int ph, int pw,
                         int accumulate, GConv* g) {
  const int s = d.s;
  if (pd >= d.xd || ph >= d.xh || pw >= d.xw) return false;
  g->n = d.n;
  g->id = d.yd; g->ih = d.yh; g->iw = d.yw; g->cin = d.cy;
  g->od = d.xd; g->oh = d.xh; g->ow = d.xw; g->cout = d.cx;
  g->gd = (d.xd - pd + s - 1) / s; g->gh = (d.xh - ph + s - 1) / s; g->gw = (d.xw - pw + s - 1) / s;
  g->is = 1; g->os = s; g->ood = pd; g->ooh = ph; g->oow = pw;
  g->ntaps = 0;
  for (int td = 0; td < d.kd; ++td) {
    int ed = pd + d.pad_d - td;
    if (((ed % s) + s) % s) continue;
    for (int th = 0; th < d.kh; ++th) {
      int eh = ph + d.pad_h - th;
      if (((eh % s) + s) % s) continue;
      for (int tw = 0; tw < d.kw; ++tw) {
        int ew = pw + d.pad_w - tw;
        if (((ew % s) + s) % s) continue;
        GTap& t = g->taps[g->ntaps++];
        // exact division (remainder is zero), valid for negative numerators too
        t.od = ed / s; t.oh = eh / s; t.ow = ew / s;
        t.widx = (td * d.kh + th) * d.kw + tw;
      }
    }
  }
  g->wtap = (int64_t)d.cx * d.cy; g->wldi = 1; g->wldo = d.cy;
  g->ldin = ldy; g->ldout = ldx; g->accumulate = accumulate;
  return true;
}

extern "C" int urn_conv_fwd(const urn_conv_desc* d, int dtype, int impl, const void* x, int64_t ldx,
                            const float* w, void* y, int64_t ldy, int y_dtype, void* stream) {
  if (check_desc(d)) return 1;
  URN_CHECK(ldx >= d->cx && ldy >= d->cy, "urn_conv_fwd: pitch smaller than channel count");
  URN_CHECK(impl != URN_IMPL_TC, "urn_conv_fwd: the tcgen05 path runs through the net handle "
            "(packed bf16 filters); per-kernel entry is urn_tc_conv");
  if (impl == URN_IMPL_AUTO && y_dtype == URN_F32) {     // dedicated stem / head kernels (stemhead.cu)
    if (stem_supported(*d, ldx, ldy)) return stem_fwd_launch(*d, dtype, x, ldx, w, (float*)y, ldy, (cudaStream_t)stream);
    if (head_supported(*d, ldx, ldy)) return head_fwd_launch(*d, dtype, x, ldx, w, (float*)y, ldy, (cudaStream_t)stream);
  }
  GConv g;
  gconv_gather(*d, ldx, ldy, &g);
  return gconv_ffma_launch(g, dtype, x, w, y, y_dtype, (cudaStream_t)stream);
}

extern "C" int urn_conv_dgrad(const urn_conv_desc* d, int dtype, int impl, const void* dy,
                              int64_t ldy, const float* w, void* dx, int64_t ldx, int x_dtype,
                              int accumulate, void* stream) {
  if (check_desc(d)) return 1;
  URN_CHECK(ldx >= d->cx && ldy >= d->cy, "urn_conv_dgrad: pitch smaller than channel count");
  URN_CHECK(impl != URN_IMPL_TC, "urn_conv_dgrad: tcgen05 path runs through the net handle");
  if (impl == URN_IMPL_AUTO && x_dtype == dtype && head_supported(*d, ldx, ldy) && ldx % 4 == 0)
    return head_dgrad_launch(*d, dtype, dy, ldy, w, dx, ldx, accumulate, (cudaStream_t)stream);
  for (int pd = 0; pd < d->s; ++pd)
    for (int ph = 0; ph < d->s; ++ph)
      for (int pw = 0; pw < d->s; ++pw) {
        GConv g;
        if (!gconv_scatter_class(*d, ldy, ldx, pd, ph, pw, accumulate, &g)) continue;
        if (g.ntaps == 0 && accumulate) continue;          // a class no tap reaches (1x1 stride 2): nothing to add
        if (gconv_ffma_launch(g, dtype, dy, w, dx, x_dtype, (cudaStream_t)stream)) return 1;
      }
  return 0;
}

extern "C" int64_t urn_conv_wgrad_workspace(const urn_conv_desc* d) {
  if (check_desc(d)) return -1;
  return wgrad_any_workspace(*d);
}

extern "C" int urn_conv_wgrad(const urn_conv_desc* d, int dtype, int impl, const void* x, int64_t ldx,
                              const void* y, int64_t ldy, int y_dtype, float* dw, int accumulate,
                              void* ws, void* stream) {
  if (check_desc(d)) return 1;
  URN_CHECK(ldx >= d->cx && ldy >= d->cy, "urn_conv_wgrad: pitch smaller than channel count");
  if (impl == URN_IMPL_TC) {
    URN_CHECK(wgrad_tc_supported(*d, dtype, y_dtype, ldx, ldy), "urn_conv_wgrad: shape/dtype not supported by the tcgen05 path");
    return wgrad_tc_auto(*d, dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, (cudaStream_t)stream);
  }
  if (impl == URN_IMPL_AUTO) {
    if (stem_supported(*d, ldx, d->cy) && ldy % 4 == 0)
      return stem_wgrad_launch(*d, dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, (cudaStream_t)stream);
    if (head_supported(*d, ldx, ldy))
      return head_wgrad_launch(*d, dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, (cudaStream_t)stream);
  }
  return wgrad_ffma_launch(*d, dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, (cudaStream_t)stream);
}

// ---------------- tcgen05 per-kernel entry points ----------------
namespace {
template <typename T>
__global__ void pack_one_kernel(const float* __restrict__ w, int taps, int cx, int cy, int transpose,
                                T* __restrict__ dst) {
  const long long n = (long long)taps * cx * cy;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i % cy);
    const int x = (int)((i / cy) % cx);
    const long long t = i / ((long long)cx * cy);
    st1(dst + (transpose ? (t * cy + y) * cx + x : i), w[i]);
  }
}
int pack_filter(const urn_conv_desc& d, int dtype, const float* w, void* wpack, bool gather, cudaStream_t st) {
  const int taps = d.kd * d.kh * d.kw;
  const long long n = (long long)taps * d.cx * d.cy;
  const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148 * 4);
  if (dispatch_dtype(dtype, [&](auto pt) {
        using T = URN_T(pt);
        pack_one_kernel<T><<<nb, 256, 0, st>>>(w, taps, d.cx, d.cy, gather ? 1 : 0, (T*)wpack);
        return 0;
      }))
    return 1;
  URN_LAUNCH_CHECK();
  return 0;
}
}  // namespace

extern "C" int urn_conv_fwd_tc(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx,
                               const float* w, void* wpack, void* y, int64_t ldy, int y_dtype,
                               int accumulate, void* stream) {
  if (check_desc(d)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  GConv g;
  gconv_gather(*d, ldx, ldy, &g);
  g.accumulate = accumulate;
  URN_CHECK(gconv_tc_supported(g, dtype, y_dtype), "urn_conv_fwd_tc: shape/dtype not supported by the tcgen05 path");
  if (pack_filter(*d, dtype, w, wpack, true, st)) return 1;
  return gconv_tc_auto(g, dtype, x, wpack, y, y_dtype, st);
}

extern "C" int urn_conv_fwd_tc_stats_t(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx,
                                       const float* w, void* wpack, void* y, int y_dtype, int64_t ldy, float eps,
                                       float* stats, void* ws, void* stream) {
  if (check_desc(d)) return 1;
  URN_CHECK(y_dtype == URN_F32 || y_dtype == URN_F16, "urn_conv_fwd_tc_stats: raw output dtype %d", y_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  GConv g;
  gconv_gather(*d, ldx, ldy, &g);
  const bool halo = gconv_halo_supported(g, dtype, y_dtype);
  URN_CHECK(halo || gconv_tc_supported(g, dtype, y_dtype), "urn_conv_fwd_tc_stats: shape/dtype not taken by the tensor-core kernels");
  if (pack_filter(*d, dtype, w, wpack, true, st)) return 1;
  int nparts = 0;
  if (halo) {
    if (gconv_halo_launch_stats(g, dtype, x, wpack, y, y_dtype, 0, (float*)ws, &nparts, st)) return 1;
  } else if (gconv_tc_launch_stats(g, dtype, x, wpack, y, y_dtype, (float*)ws, &nparts, st)) return 1;
  URN_CHECK(nparts > 0, "urn_conv_fwd_tc_stats: statistics were not fused for this launch");
  const long long rows = (long long)d->n * d->yd * d->yh * d->yw;
  return bn_stats_merge_launch((const float*)ws, nparts, d->cy, d->cy, rows, eps, stats, st);
}
extern "C" int urn_conv_fwd_tc_stats(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx,
                                     const float* w, void* wpack, float* y, int64_t ldy, float eps,
                                     float* stats, void* ws, void* stream) {
  return urn_conv_fwd_tc_stats_t(d, dtype, x, ldx, w, wpack, y, URN_F32, ldy, eps, stats, ws, stream);
}

extern "C" int urn_conv_dgrad_tc(const urn_conv_desc* d, int dtype, const void* dy, int64_t ldy,
                                 const float* w, void* wpack, void* dx, int64_t ldx, int x_dtype,
                                 int accumulate, void* stream) {
  if (check_desc(d)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (pack_filter(*d, dtype, w, wpack, false, st)) return 1;
  if (d->s == 2 && d->kh == 3) {          // all parity classes in one launch when the halo kernel takes them
    const int rc = gconv_scatter_tc(*d, ldy, ldx, accumulate, dtype, dy, wpack, dx, x_dtype, st);
    if (rc >= 0) return rc;
  }
  for (int pd = 0; pd < d->s; ++pd)
    for (int ph = 0; ph < d->s; ++ph)
      for (int pw = 0; pw < d->s; ++pw) {
        GConv g;
        if (!gconv_scatter_class(*d, ldy, ldx, pd, ph, pw, accumulate, &g)) continue;
        if (g.ntaps == 0) {   // parity class without taps (1x1 stride-2): zeros, CUDA-core kernel
          if (accumulate) continue;                         // ... or nothing at all when adding
          if (gconv_ffma_launch(g, dtype, dy, w, dx, x_dtype, st)) return 1;
          continue;
        }
        URN_CHECK(gconv_tc_supported(g, dtype, x_dtype), "urn_conv_dgrad_tc: shape/dtype not supported by the tcgen05 path");
        if (gconv_tc_auto(g, dtype, dy, wpack, dx, x_dtype, st)) return 1;
      }
  return 0;
}
