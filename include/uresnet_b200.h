/* uresnet_b200.h -- C ABI of the B200-native U-ResNet hot path.
 *
 * The reference (DeepLearnPhysics/u-resnet) has no FFI of its own: the seam is the
 * set of TensorFlow-1.x ops its graph invokes and the Python
 * class API of lib/ssnet.py.  Each entry point below names the reference call site
 * (file:line under /root/reference) whose arithmetic it replaces.  Python binds this
 * header with ctypes (u-resnet_b200/_capi.py); torch tensors are used only as buffers.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; urn_last_error() gives
 *    the thread-local message.  Nothing throws or exits across the boundary.
 *  - all pointers are DEVICE pointers owned by the caller unless noted "host".
 *  - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no hidden
 *    synchronisation) unless noted.
 *  - activations are channel-last (NHWC / NDHWC); 2D is handled as 3D with depth 1.
 *    A tensor argument is (pointer, ld) where ld = element pitch between consecutive
 *    pixels (>= channels) so that channel slices of a concat buffer are first class.
 *  - storage dtypes: URN_F32 (float), URN_BF16, URN_F16; math is always fp32-accumulate.
 *    Network modes (urn_net_cfg.dtype): URN_F32 = everything float on the CUDA cores (the
 *    1e-3 parity mode); URN_F16 = "mix16": activations / filters fp16, gradients bf16
 *    (tcgen05 tensor cores where the layer shape allows); URN_BF16 = all-bf16 storage.
 */
#ifndef URESNET_B200_H
#define URESNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { URN_F32 = 0, URN_BF16 = 1, URN_F16 = 2 };
enum { URN_IMPL_AUTO = 0, URN_IMPL_FFMA = 1, URN_IMPL_TC = 2 };

/* Geometry of one conv-type layer, stated as the relation between a "big side"
 * tensor X [n, xd, xh, xw, cx] and a "small side" tensor Y [n, yd, yh, yw, cy]:
 *
 *      Y[o, :] = sum_t  X[o*s + t - pad, :] . W[t]        W[t] is [cx, cy]
 *
 * which is the TF filter layout for BOTH slim.conv (X = input, cx = Cin; filter
 * [k..,Cin,Cout]) and slim.conv_transpose (X = OUTPUT, cx = Cout; filter
 * [k..,Cout,Cin]) -- SURVEY.md 3.4.  pad[] is TF 'same' pad_before per dim. */
typedef struct urn_conv_desc {
  int32_t n;
  int32_t xd, xh, xw, cx;
  int32_t yd, yh, yw, cy;
  int32_t kd, kh, kw;   /* kernel extent per dim (kd = 1 for 2D) */
  int32_t s;            /* stride (same in every dim)            */
  int32_t pad_d, pad_h, pad_w;
} urn_conv_desc;

/* Fill a descriptor with TF 'same' geometry: x-side spatial dims (xd = 1 for 2D),
 * kernel k, stride s.  Replaces the padding arithmetic inside slim.conv2d/3d
 * (lib/uresnet.py:37-45, lib/resnet_module.py:25-66). */
int urn_conv_desc_same(urn_conv_desc* d, int nd, int n, int xd, int xh, int xw, int cx,
                       int cy, int k, int s);

const char* urn_last_error(void);
/* Library / build information: returns a static string (arch, build flags). */
const char* urn_build_info(void);

/* ---------------- per-kernel entry points ---------------- */

/* K1/K2/K3  slim.conv2d / conv3d forward (lib/uresnet.py:37-45,103-121,
 * lib/resnet_module.py:25-66):  y = conv_same(x, w).  x is `dtype`, w is f32 in TF
 * layout [k..,Cin,Cout], y is f32 (raw pre-BN output).  Also the input-gradient of
 * slim.conv*_transpose (Conv*BackpropInput used as forward, lib/uresnet.py:68-88). */
int urn_conv_fwd(const urn_conv_desc* d, int dtype, int impl, const void* x, int64_t ldx,
                 const float* w, void* y, int64_t ldy, int y_dtype, void* stream);

/* Conv*BackpropInput: dx = conv_same^T(dy, w); dx and dy are `dtype`.
 * accumulate != 0 adds into dx.  Also slim.conv*_transpose FORWARD
 * (lib/uresnet.py:68-88) with x := deconv output, y := deconv input; then the
 * output x may be f32 (x_dtype). */
int urn_conv_dgrad(const urn_conv_desc* d, int dtype, int impl, const void* dy, int64_t ldy,
                   const float* w, void* dx, int64_t ldx, int x_dtype, int accumulate,
                   void* stream);

/* The same two operations on the tcgen05 tensor-core path (16-bit operands, fp32 accumulate;
 * needs cx % 16 == 0 and cy % 16 == 0).  The f32 TF-layout filter is first packed to the
 * operand dtype into `wpack` (caller scratch, >= 2*taps*cx*cy bytes); inside a network handle
 * the packing happens once per step instead.  y_dtype/x_dtype: URN_F32 or URN_BF16; y_dtype may
 * also be URN_F16 (the mix16 raw output; not with accumulate). */
int urn_conv_fwd_tc(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx, const float* w,
                    void* wpack, void* y, int64_t ldy, int y_dtype, int accumulate, void* stream);
int urn_conv_dgrad_tc(const urn_conv_desc* d, int dtype, const void* dy, int64_t ldy, const float* w,
                      void* wpack, void* dx, int64_t ldx, int x_dtype, int accumulate, void* stream);
/* urn_conv_fwd_tc (f32 output) with the slim.batch_norm statistics of y fused into the conv
 * epilogue (K5 without a separate pass over the raw tensor): stats = float[2*cy] (mean | rstd) as
 * urn_bn_stats would produce them.  ws: urn_bn_workspace(cy).  Fails (non-zero) when the layer
 * shape is not one whose kernel fuses the statistics (halo-tile kernel: Cin <= 64; tap-box kernel:
 * Cout <= 256).  urn_conv_fwd_tc_stats_t: y stored as y_dtype (URN_F32 or URN_F16); the statistics
 * are those of the fp32 accumulators either way. */
int urn_conv_fwd_tc_stats(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx, const float* w,
                          void* wpack, float* y, int64_t ldy, float eps, float* stats, void* ws,
                          void* stream);
int urn_conv_fwd_tc_stats_t(const urn_conv_desc* d, int dtype, const void* x, int64_t ldx, const float* w,
                            void* wpack, void* y, int y_dtype, int64_t ldy, float eps, float* stats, void* ws,
                            void* stream);

/* Conv*BackpropFilter: dw[t][cx][cy] (+)= sum_o x[o*s+t-pad][cx] * y[o][cy].
 * ws must hold urn_conv_wgrad_workspace(d) bytes. */
int64_t urn_conv_wgrad_workspace(const urn_conv_desc* d);
int urn_conv_wgrad(const urn_conv_desc* d, int dtype, int impl, const void* x, int64_t ldx,
                   const void* y, int64_t ldy, int y_dtype, float* dw, int accumulate, void* ws,
                   void* stream);

/* K5  slim.batch_norm forward statistics (normalizer_fn of every conv call site):
 * per-channel mean and rsqrt(biased var + eps) over all `rows` pixels of raw f32
 * [rows, c] (pitch ld).  stats = float[2*c] (mean | rstd).  ws: urn_bn_workspace(c). */
int64_t urn_bn_workspace(int c);
int urn_bn_stats(const float* raw, int64_t ld, int64_t rows, int c, float eps, float* stats,
                 void* ws, void* stream);

/* K5/K6  out = act( (raw-mean)*rstd + beta  [+ residual] ).
 * residual: res (dtype, pitch ldres) added as is (identity shortcut,
 * lib/resnet_module.py:22-23,68) or, if res_raw != NULL, BN(res_raw; res_stats,
 * res_beta) (conv shortcut, lib/resnet_module.py:25-33,68).  relu != 0 applies
 * tf.nn.relu.  out is out_dtype with pitch ldout.  out2 (may be NULL) receives the same values
 * as bf16 with the same pitch: the copy the tcgen05 filter-gradient kernel multiplies with the
 * bf16 gradients (kind::f16 needs both operands in one 16-bit format). */
int urn_bn_apply(const float* raw, int64_t ld, int64_t rows, int c, const float* stats,
                 const float* beta, const void* res, int64_t ldres, int res_dtype,
                 const float* res_raw, int64_t ldres_raw, const float* res_stats,
                 const float* res_beta, int relu, void* out, int64_t ldout, int out_dtype,
                 void* out2, void* stream);

/* K8  batch-norm (+ReLU, +residual fan-out) backward.
 * g = dout * (relu ? out>0 : 1);  dbeta = sum g;  draw = rstd*(g - mean(g) - xhat*mean(g*xhat)).
 * If raw2 != NULL the same g also flows through a second BN (conv shortcut) giving
 * dbeta2/draw2.  If dres != NULL, g itself is written (or accumulated) there
 * (identity shortcut).  dbeta(2) are accumulated (+=) when acc_beta != 0.
 * dtype = storage of dout/draw/draw2/dres; out_dtype = storage of the forward output. */
int urn_bn_bwd(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
               int out_dtype, int relu, int64_t rows, int c,
               const float* raw, int64_t ldraw, const float* stats, float* dbeta, void* draw,
               int64_t lddraw,
               const float* raw2, int64_t ldraw2, const float* stats2, float* dbeta2,
               void* draw2, int64_t lddraw2,
               void* dres, int64_t lddres, int dres_accumulate, int acc_beta,
               void* ws, void* stream);

/* Same as urn_bn_bwd, told the layer's beta (f32[c]).  With beta the fast path needs fewer
 * bytes: for BN without ReLU the forward output is xhat + beta, so xhat = out - beta and raw is
 * not read; for BN -> ReLU without residual the mask is recomputed from raw as
 * fma(raw - mean, rstd, beta) > 0 and `out` is not read.  beta == NULL behaves as urn_bn_bwd. */
int urn_bn_bwd_beta(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
                    int out_dtype, int relu, int64_t rows, int c,
                    const float* raw, int64_t ldraw, const float* stats, const float* beta,
                    float* dbeta, void* draw, int64_t lddraw,
                    const float* raw2, int64_t ldraw2, const float* stats2, float* dbeta2,
                    void* draw2, int64_t lddraw2,
                    void* dres, int64_t lddres, int dres_accumulate, int acc_beta,
                    void* ws, void* stream);

/* The three batch-norm entry points with an explicit storage type of the raw conv output(s):
 * raw_dtype = URN_F32 (the functions above) or URN_F16.  In the mix16 network mode the conv
 * epilogues store raw as fp16 next to the fp16 activations (statistics are still taken from the
 * fp32 accumulators): 2 instead of 4 bytes per element in the conv write, the apply read and
 * both backward reads.  URN_F16 needs power-of-two channel counts >= 8, pitches that are multiples
 * of 8 elements, fp16 activations and bf16 gradients.  urn_bn_bwd_t takes beta like
 * urn_bn_bwd_beta (may be NULL); raw2 / res_raw share raw_dtype. */
int urn_bn_stats_t(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, float eps,
                   float* stats, void* ws, void* stream);
int urn_bn_apply_t(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, const float* stats,
                   const float* beta, const void* res, int64_t ldres, int res_dtype,
                   const void* res_raw, int64_t ldres_raw, const float* res_stats,
                   const float* res_beta, int relu, void* out, int64_t ldout, int out_dtype,
                   void* out2, void* stream);
int urn_bn_bwd_t(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype,
                 int out_dtype, int relu, int64_t rows, int c,
                 const void* raw, int raw_dtype, int64_t ldraw, const float* stats, const float* beta,
                 float* dbeta, void* draw, int64_t lddraw,
                 const void* raw2, int64_t ldraw2, const float* stats2, float* dbeta2,
                 void* draw2, int64_t lddraw2,
                 void* dres, int64_t lddres, int dres_accumulate, int acc_beta,
                 void* ws, void* stream);

/* K9/K10/K11  tf.nn.sparse_softmax_cross_entropy_with_logits + weight + per-image
 * sum + batch mean, argmax accuracies, softmax (lib/ssnet.py:57-71).
 * logits f32 [batch*pixels, c]; label/weight/data f32 [batch*pixels] (weight may be
 * NULL); dlogits (dl_dtype, may be NULL) = (softmax - onehot)*w/batch; softmax f32 may be
 * NULL.  result (device, double[4]) = loss, n_correct_all, n_correct_nonzero, n_nonzero.
 * An out-of-range label makes result[0] NaN and raises the device error flag that
 * urn_check_labels() reports.  ws: urn_loss_workspace(batch*pixels). */
int64_t urn_loss_workspace(int64_t total_pixels);
int urn_ce_fwd_bwd(const float* logits, int64_t ld, const float* label, const float* weight,
                   const float* data, int64_t batch, int64_t pixels, int c, void* dlogits,
                   int64_t lddl, int dl_dtype, float* softmax, double* result, void* ws,
                   void* stream);

/* Shower/track decision of the analysis step (lib/ssnet_trainval.py:269-287, the numpy
 * expression applied per entry before the larcv write):
 *   decision[p] = (1*(p1 > p2) + 2*(p2 >= p1)) * (data[p*data_ld] > threshold),
 * p = softmax(logits[p*ld .. +num_class)) evaluated exactly as urn_ce_fwd_bwd's softmax output,
 * so the result equals the host expression on that softmax value for value.  num_class >= 3. */
int urn_ssnet_decision(const float* logits, int64_t ld, int64_t total, int num_class,
                       const float* data, int64_t data_ld, float threshold, float* decision,
                       void* stream);

/* TF-1.x AdamOptimizer apply (lib/ssnet.py:72-79), epsilon outside the bias
 * correction, on flat f32 buffers; t is the 1-based step. */
int urn_adam_step(float* param, const float* grad, float* m, float* v, int64_t count, float lr,
                  float beta1, float beta2, float eps, int64_t t, void* stream);

/* ---------------- whole-network handle ---------------- */

typedef struct urn_net_cfg {
  int32_t nd;              /* 2 or 3                                   */
  int32_t spatial[3];      /* H,W  or  X,Y,Z  (unused trailing = 1)    */
  int32_t in_channels;     /* dims[-1] of lib/ssnet.py:10 (1)          */
  int32_t batch;           /* images per call                          */
  int32_t base_filters;    /* BASE_NUM_FILTERS   lib/config.py:9       */
  int32_t num_strides;     /* uresnet(num_strides=5) lib/uresnet.py:15 */
  int32_t num_class;       /* NUM_CLASS          lib/config.py:8       */
  int32_t dtype;           /* URN_F32 | URN_F16 (mix16) | URN_BF16     */
  int32_t impl;            /* URN_IMPL_*                               */
  int32_t training;        /* keep activations for urn_backward        */
} urn_net_cfg;

typedef struct urn_net urn_net;

int urn_create(const urn_net_cfg* cfg, urn_net** out);
void urn_destroy(urn_net* h);

/* Parameters: 2 per layer in tf.trainable_variables() order (weights, beta).
 * name is the TF variable name ("UResNet/conv0/weights"); shape has `rank` entries;
 * offset is the element offset into the flat f32 parameter/gradient buffers. */
int urn_param_count(const urn_net* h);
int urn_param_info(const urn_net* h, int i, const char** name, int* rank, int64_t* shape,
                   int64_t* offset, int64_t* count);
int64_t urn_param_total(const urn_net* h);          /* elements of the flat buffer   */
int urn_layer_count(const urn_net* h);
int64_t urn_workspace_bytes(const urn_net* h);
int64_t urn_logits_offset_bytes(const urn_net* h);   /* f32 logits inside the workspace */

/* uresnet._build (lib/uresnet.py:22-123) + lib/ssnet.py:34 reshape: data is host-layout
 * f32 [batch, spatial.., in_channels] on the device.  Leaves f32 logits
 * [batch*pixels, num_class] in the workspace (urn_logits_offset_bytes). */
int urn_forward(urn_net* h, const float* params, const float* data, void* ws, void* stream);

/* Loss/metrics on the logits left by urn_forward; if want_grad, seeds backward. */
int urn_loss_metrics(urn_net* h, const float* label, const float* weight, const float* data,
                     float* softmax, double* result, int want_grad, void* ws, void* stream);

/* urn_ssnet_decision on the logits left by urn_forward; data as given to urn_forward (channel 0
 * is thresholded); decision f32 [batch*pixels]. */
int urn_decision(urn_net* h, const float* data, float threshold, float* decision, void* ws,
                 void* stream);

/* tf.gradients of the loss w.r.t. all parameters, accumulated (+=) into grads when
 * accumulate != 0 (lib/ssnet.py:76-78).  If ready_events != NULL it is a host array
 * of urn_layer_count() cudaEvent_t recorded on `stream` as soon as that layer's
 * weight+beta gradients are final (layers finish in reverse order) -- the hook the
 * data-parallel bucketed allreduce overlaps on. */
int urn_backward(urn_net* h, const float* params, float* grads, int accumulate, void* ws,
                 void* const* ready_events, void* stream);

/* Host-side check of the device flag the last urn_loss_metrics call raised for a label outside
 * [0, num_class) (the same call makes the loss NaN): returns nonzero and sets urn_last_error()
 * if it is set.  Reads 4 bytes from the workspace and synchronises the stream. */
int urn_check_labels(urn_net* h, void* ws, void* stream);

/* Per-layer timing/debug access: pointer + geometry of a layer's tensors in ws. */
int urn_layer_info(const urn_net* h, int layer, const char** scope, urn_conv_desc* d,
                   int* is_transpose, int64_t* raw_off_bytes, int64_t* out_off_bytes,
                   int64_t* out_ld, int64_t* stats_off_bytes);

/* ---------------- plan introspection + partial execution (parity hooks) ----------------
 * The reference exposes every intermediate of its graph by name (any tf.Tensor can be fetched or
 * fed through sess.run's feed_dict, lib/ssnet.py:91-139).  These entry points give the same
 * access to the static launch schedule: where each graph node / layer tensor lives in the
 * workspace, what each schedule step reads and writes, and execution of a sub-range of steps, so
 * that a test can feed one step the reference's (oracle's) tensors and fetch what it produces
 * (tests/test_parity_gpu.py: teacher-forced per-layer parity).  Not used by the product path. */
typedef struct urn_tensor_ref {
  int64_t off_bytes;   /* byte offset of element (row 0, channel 0) in the workspace; -1 = absent */
  int64_t ld;          /* element pitch between consecutive rows (pixels)                         */
  int64_t rows;        /* batch * spatial positions                                              */
  int32_t c;           /* channels                                                               */
  int32_t dtype;       /* URN_F32 | URN_BF16 | URN_F16                                           */
} urn_tensor_ref;

/* One step of the forward / backward schedule.
 * forward:  kind 0 = conv (+BN statistics): reads node_a, writes raw/stats of `layer`;
 *           kind 1 = BN apply (+ReLU): reads raw/stats of `layer`, writes node_a;
 *           kind 2 = residual merge relu(BN(layer) + shortcut): shortcut = BN(layer2) if
 *                    layer2 >= 0 else node_b (identity); writes node_a.
 * backward: kind 0 = BN(+ReLU) backward of `layer`: reads the gradient of node_a, writes
 *                    d(raw) + d(beta) of `layer` (and of the conv shortcut `layer2` if >= 0) and,
 *                    if node_b >= 0, the masked gradient into the gradient of node_b (identity
 *                    shortcut; added when `accumulate`);
 *           kind 1 = filter gradient of `layer` (reads its input node node_a and d(raw));
 *           kind 2 = input gradient of `layer`: d(raw) -> gradient of node_a (added when
 *                    `accumulate`). */
typedef struct urn_step_info {
  int32_t kind, layer, layer2, node_a, node_b, relu, accumulate;
} urn_step_info;

int urn_plan_counts(const urn_net* h, int* n_nodes, int* n_fwd_steps, int* n_bwd_steps);
/* Graph node `node`: name = producing scope under UResNet/ ("conv0", "resnet_module3/module1"
 * [the module output], "resnet_module3/module1/resnet_conv1" [BN output inside the module],
 * "deconv2", "concat2", "data"); value = stored activation; shadow = its bf16 copy (mix16) or
 * absent; grad = its gradient buffer in the backward arena (valid only while live). */
int urn_node_info(const urn_net* h, int node, const char** name, urn_tensor_ref* value,
                  urn_tensor_ref* shadow, urn_tensor_ref* grad);
/* Layer tensors: raw fp32 conv output, d(raw) (pitch = padded width for the tensor-core head),
 * byte offset of the statistics (float[2*c]: mean | rstd), element offsets of the filter / beta
 * in the flat parameter + gradient buffers, input node, node written by its own BN apply (-1 if
 * it is merged into a module output). */
int urn_layer_tensors(const urn_net* h, int layer, urn_tensor_ref* raw, urn_tensor_ref* draw,
                      int64_t* stats_off_bytes, int64_t* w_off, int64_t* b_off, int* in_node,
                      int* out_node);
int urn_fwd_step_info(const urn_net* h, int step, urn_step_info* info);
int urn_bwd_step_info(const urn_net* h, int step, urn_step_info* info);
/* Steps [first, last) of urn_forward / urn_backward.  urn_forward_range runs the input
 * conversion and the per-step filter packing first when data != NULL (pass NULL to keep what a
 * previous urn_forward left). */
int urn_forward_range(urn_net* h, const float* params, const float* data, void* ws, int first,
                      int last, void* stream);
int urn_backward_range(urn_net* h, const float* params, float* grads, int accumulate, void* ws,
                       int first, int last, void* stream);

/* ---------------- measurement hooks ---------------- */
/* Number of kernels this library has launched in this process (host-side counter). */
int64_t urn_launch_count(void);

/* Per-layer, per-phase device timing with CUDA events recorded on the launching stream.
 * Phases: 0 conv forward, 1 BN statistics, 2 BN apply(+ReLU/+residual), 3 BN backward,
 * 4 filter gradient, 5 input gradient, 6 loss/metrics + input staging, 7 unused.
 * urn_profile_collect synchronises, ADDS the elapsed milliseconds and launch-group counts
 * into ms[layer*8+phase] / counts[layer*8+phase] (arrays of urn_layer_count()*8) and
 * clears the recorded events. */
int urn_profile_enable(urn_net* h, int on);
int urn_profile_collect(urn_net* h, double* ms, int64_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* URESNET_B200_H */
