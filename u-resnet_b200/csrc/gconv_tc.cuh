// tcgen05 implicit-GEMM convolution (declarations).  See gconv_tc.cu.
#pragma once
#include "common.cuh"
#define URN_STR_(x) #x
#define URN_STR(x) URN_STR_(x)

void gconv_gather(const urn_conv_desc& d, int64_t ldx, int64_t ldy, GConv* g);
bool gconv_scatter_class(const urn_conv_desc& d, int64_t ldy, int64_t ldx, int pd, int ph, int pw,
                         int accumulate, GConv* g);
int urn_softmax_only(const float* logits, int64_t ld, int64_t total, int c, float* softmax,
                     cudaStream_t st);

// tcgen05 path (gconv_tc.cu).  wpack: 16-bit filter, same dtype as `in`, laid out [tap][cout][cin].
bool gconv_tc_supported(const GConv& g, int in_dtype, int out_dtype);
int gconv_tc_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                    int out_dtype, cudaStream_t st);
int gconv_tc_launch_stats(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                          int out_dtype, float* stats_part, int* nparts, cudaStream_t st);
bool wgrad_tc_supported(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy);
int64_t wgrad_tc_workspace(const urn_conv_desc& d);
int wgrad_tc_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                    const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);

// Halo-tile variant for narrow stride-1 3x3(x3) layers (gconv_halo.cu); same wpack layout.
bool gconv_halo_supported(const GConv& g, int in_dtype, int out_dtype);
int gconv_halo_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                      int out_dtype, cudaStream_t st);
// Zero-padded filter variant: g.cout is the padded width, only cout_store channels are written.
bool gconv_halo_supported_padded(const GConv& g, int in_dtype, int out_dtype, int cout_store);
int gconv_halo_launch_padded(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                             int out_dtype, int cout_store, cudaStream_t st);
// Same, with the BN statistics of the fp32 output fused into the epilogue: stats_part receives
// *nparts (n, mean, M2) float triples per channel ([part][cout][3]); *nparts = 0 if not fused.
int gconv_halo_launch_stats(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                            int out_dtype, int cout_store, float* stats_part, int* nparts, cudaStream_t st);
// Chan merge of the partials -> stats (mean | rstd), deterministic (bn.cu).
int bn_stats_merge_launch(const float* parts, int nparts, int c_stride, int c, long long rows, float eps,
                          float* stats, cudaStream_t st);
// All parity classes of a stride-2 scatter (conv input gradient / conv-transpose forward) in ONE
// launch: the small side is staged once, every class has its own TMEM accumulators and the
// interleaved outputs of a tile are written by the same CTA.
bool gconv_halo_classes_supported(const GConv* gs, int ncls, int in_dtype, int out_dtype);
int gconv_halo_classes_launch(const GConv* gs, int ncls, int in_dtype, const void* in, const void* wpack, void* out,
                              int out_dtype, float* stats_part, int* nparts, cudaStream_t st);
// Scatter form of a conv relation: one fused launch if possible, else one launch per parity class.
// Returns -1 when the tensor-core path does not take the shape at all (caller falls back).
static inline int gconv_scatter_tc(const urn_conv_desc& d, int64_t ldy, int64_t ldx, int accumulate, int in_dtype,
                                   const void* in, const void* wpack, void* out, int out_dtype, cudaStream_t st,
                                   float* stats_part = nullptr, int* nparts = nullptr);
// Picks the best tensor-core kernel for the shape.
static inline int gconv_tc_auto(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                                int out_dtype, cudaStream_t st) {
  if (gconv_halo_supported(g, in_dtype, out_dtype)) return gconv_halo_launch(g, in_dtype, in, wpack, out, out_dtype, st);
  return gconv_tc_launch(g, in_dtype, in, wpack, out, out_dtype, st);
}

// Dedicated CUDA-core kernels for the stem (Cin = 1) and the head (Cout <= 4) -- stemhead.cu.
bool stem_supported(const urn_conv_desc& d, int64_t ldx, int64_t ldy);
bool head_supported(const urn_conv_desc& d, int64_t ldx, int64_t ldy);
int64_t stemhead_wgrad_workspace(const urn_conv_desc& d);
int stem_fwd_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, const float* w,
                    float* y, int64_t ldy, cudaStream_t st);
int stem_wgrad_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int g_dtype,
                      const void* dy, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);
int head_fwd_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, const float* w,
                    float* y, int64_t ldy, cudaStream_t st);
int head_dgrad_launch(const urn_conv_desc& d, int g_dtype, const void* dy, int64_t ldy, const float* w,
                      void* dx, int64_t ldx, int accumulate, cudaStream_t st);
int head_wgrad_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int g_dtype,
                      const void* dy, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);

// Halo-tile filter gradient for narrow stride-1 3x3(x3) layers (wgrad_halo.cu).
bool wgrad_halo_supported(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy);
int64_t wgrad_halo_workspace(const urn_conv_desc& d);
int wgrad_halo_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                      const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);
// Best tensor-core filter-gradient kernel for the shape.
static inline int wgrad_tc_auto(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                                const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  if (wgrad_halo_supported(d, x_dtype, y_dtype, ldx, ldy))
    return wgrad_halo_launch(d, x_dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, st);
  return wgrad_tc_launch(d, x_dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, st);
}
static inline int64_t wgrad_any_workspace(const urn_conv_desc& d) {
  int64_t a = wgrad_ffma_workspace(d), b = wgrad_tc_workspace(d), c = wgrad_halo_workspace(d), e = stemhead_wgrad_workspace(d);
  a = a > b ? a : b; c = c > e ? c : e;
  return a > c ? a : c;
}

static inline int gconv_scatter_tc(const urn_conv_desc& d, int64_t ldy, int64_t ldx, int accumulate, int in_dtype,
                                   const void* in, const void* wpack, void* out, int out_dtype, cudaStream_t st,
                                   float* stats_part, int* nparts) {
  GConv gs[8];
  int n = 0;
  for (int pd = 0; pd < d.s; ++pd)
    for (int ph = 0; ph < d.s; ++ph)
      for (int pw = 0; pw < d.s; ++pw)
        if (gconv_scatter_class(d, ldy, ldx, pd, ph, pw, accumulate, &gs[n])) ++n;
  if ((n == 4 || n == 8) && gconv_halo_classes_supported(gs, n, in_dtype, out_dtype))
    return gconv_halo_classes_launch(gs, n, in_dtype, in, wpack, out, out_dtype, stats_part, nparts, st);
  return -1;
}
