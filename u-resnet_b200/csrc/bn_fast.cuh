// Fast-path batch-norm launches (bn_fast.cu) and the internal backward entry that knows beta.
#pragma once
#include <initializer_list>

#include "common.cuh"

bool bn_fast_shape(int c, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs);
// raw_dtype: storage of the raw conv output(s): URN_F32, or URN_F16 next to fp16 activations (mix16)
int bn_stats_fast_launch(const void* raw, int raw_dtype, long long ld, long long rows, int c, double* partial,
                         int* nblocks, cudaStream_t st);
int bn_apply_fast_launch(const void* raw, int raw_dtype, long long ld, long long rows, int c, const float* stats,
                         const float* beta, const void* res, long long ldres, int res_dtype,
                         const void* res_raw, long long ldres_raw, const float* res_stats,
                         const float* res_beta, int relu, void* out, long long ldout, int out_dtype,
                         void* out2, cudaStream_t st);
// phase 0: per-block partial sums (double [nblocks][3][c]); phase 1: elementwise pass using sums
int bn_bwd_fast_launch(int phase, const void* dout, long long lddout, const void* out, long long ldo, int dtype,
                       int out_dtype, int relu, long long rows, int c, const void* raw, int raw_dtype, long long ldraw,
                       const float* stats, const float* beta, void* draw, long long lddraw, const void* raw2,
                       long long ldraw2, const float* stats2, void* draw2, long long lddraw2, void* dres,
                       long long lddres, int dres_acc, double* partial, const float* sums, int* nblocks,
                       cudaStream_t st);
// urn_bn_bwd with the layer's beta (may be NULL): lets the fast path take xhat from `out`
// (BN without ReLU) or the ReLU mask from raw (BN -> ReLU) instead of reading both tensors.
int bn_bwd_full(const void* dout, int64_t lddout, const void* out, int64_t ldo, int dtype, int out_dtype,
                int relu, int64_t rows, int c, const void* raw, int raw_dtype, int64_t ldraw, const float* stats,
                const float* beta, float* dbeta, void* draw, int64_t lddraw, const void* raw2,
                int64_t ldraw2, const float* stats2, float* dbeta2, void* draw2, int64_t lddraw2, void* dres,
                int64_t lddres, int dres_accumulate, int acc_beta, void* ws, cudaStream_t st);
// statistics / apply with an explicit raw storage dtype (the C-ABI urn_bn_stats / urn_bn_apply are
// the URN_F32 cases); URN_F16 raw needs the fast-path shape (power-of-two channels >= 8, 8-aligned pitches)
int bn_stats_any(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, float eps, float* stats,
                 void* ws, cudaStream_t st);
int bn_apply_any(const void* raw, int raw_dtype, int64_t ld, int64_t rows, int c, const float* stats,
                 const float* beta, const void* res, int64_t ldres, int res_dtype, const void* res_raw,
                 int64_t ldres_raw, const float* res_stats, const float* res_beta, int relu, void* out,
                 int64_t ldout, int out_dtype, void* out2, cudaStream_t st);
