// tcgen05 implicit-GEMM generalised convolution for sm_100a.
//
//   out[n, o*os+oo, :] = sum_t in[n, o*is+off_t, :] . W_t          (o over a sub-grid)
//
// GEMM view per CTA tile: M = 128 output positions (a TD x TH x TW box of the sub-grid),
// N = an output-channel tile (<= 256), K = taps x Cin.  No im2col buffer exists anywhere:
// for every (tap, Cin-chunk) the TMA engine loads the tap-shifted (and, for stride-2 convs,
// stride-sampled via elementStrides) activation box straight from the NHWC/NDHWC tensor into
// a 128B/64B/32B-swizzled K-major shared-memory tile -- out-of-range coordinates are
// zero-filled by the hardware, which is exactly TF 'same' padding (including the asymmetric
// (0,1) pad of stride-2 convs) -- together with the matching [Cout_tile x Cin-chunk] slice of
// the pre-packed 16-bit filter.  One elected thread issues tcgen05.mma (kind::f16, fp32
// accumulate) into a double-buffered TMEM accumulator; four epilogue warps drain it with
// tcgen05.ld and write fp32 raw outputs (or accumulate into bf16 gradients) while the next
// tile's MMAs run.  Persistent CTAs, warp-specialised: warp0 = TMA producer, warp1 = MMA
// issuer, warp2 = TMEM allocator, warps4-7 = epilogue.
#include <cuda.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr uint32_t kSpinLimit = 1u << 26;   // bounded waits: trap instead of hanging the GPU

struct TcParams {
  int n, gd, gh, gw;            // output sub-grid extents
  int td, th, tw;               // tile box (TD*TH*TW == 128)
  int ntd, nth, ntw;            // tiles per dim
  long long ntiles;             // n * ntd * nth * ntw
  int is, os, ood, ooh, oow;    // input sampling stride / output placement
  int od, oh, ow;               // full output dims
  long long ldout;
  int cout, n_tile, n_tiles;    // output channels, N per work item, number of N tiles
  int kc, nchunks;              // Cin chunk (16/32/64 elements) and chunks per tap
  int ntaps;
  int tap_od[27], tap_oh[27], tap_ow[27], tap_w[27];
  int accumulate, out_dtype;
  int stages, a_bytes, b_bytes; // pipeline depth, smem bytes per stage of A and B
  int tx_bytes;                 // bytes the two TMA loads of a stage deliver
  uint32_t idesc;               // tcgen05 instruction descriptor
  uint64_t desc_hi;             // smem descriptor without the start address (layout, SBO, LBO)
  int tmem_cols;                // allocated TMEM columns (2 accumulators)
};

// ------------------------------ PTX wrappers ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();   // a protocol bug must not hang the device
  }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------ the kernel ------------------------------
__global__ void __launch_bounds__(kThreads, 1)
gconv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ TcParams p, void* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A][stages x B] (1024-aligned) | barriers | tmem base
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * p.b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nwork = p.ntiles * p.n_tiles;
  const int ksteps = p.ntaps * p.nchunks;          // pipeline stages consumed per work item

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int nt = (int)(w % p.n_tiles);
        long long t = w / p.n_tiles;
        const int wt = (int)(t % p.ntw); t /= p.ntw;
        const int ht = (int)(t % p.nth); t /= p.nth;
        const int dt = (int)(t % p.ntd);
        const int nn = (int)(t / p.ntd);
        const int w0 = wt * p.tw * p.is, h0 = ht * p.th * p.is, d0 = dt * p.td * p.is;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int ch = 0; ch < p.nchunks; ++ch, ++it) {
            const int s = it % p.stages;
            mbar_wait(&empty[s], ((it / p.stages) & 1) ^ 1);
            mbar_expect_tx(&full[s], (uint32_t)p.tx_bytes);
            tma_load_5d(smem_a + (size_t)s * p.a_bytes, &map_a, &full[s], ch * p.kc, w0 + p.tap_ow[tap],
                        h0 + p.tap_oh[tap], d0 + p.tap_od[tap], nn);
            tma_load_3d(smem_b + (size_t)s * p.b_bytes, &map_b, &full[s], ch * p.kc, nt * p.n_tile,
                        p.tap_w[tap]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, wi = 0;
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++wi) {
        const uint32_t acc = wi & 1;
        mbar_wait(&tempty[acc], ((wi >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.n_tile;
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % p.stages;
          mbar_wait(&full[s], (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + (size_t)s * p.a_bytes);
          const uint32_t b_addr = smem_u32(smem_b + (size_t)s * p.b_bytes);
          for (int k = 0; k < p.kc / 16; ++k) {
            const uint64_t adesc = p.desc_hi | (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFF);
            const uint64_t bdesc = p.desc_hi | (uint64_t)(((b_addr + k * 32) >> 4) & 0x3FFF);
            umma_f16(d_tmem, adesc, bdesc, p.idesc, (ks | k) ? 1u : 0u);
          }
          umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
        }
        umma_commit(&tfull[acc]);            // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;              // accumulator row = position inside the tile box
    const int lw = r % p.tw, lh = (r / p.tw) % p.th, ld_ = r / (p.tw * p.th);
    uint32_t wi = 0;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++wi) {
      const uint32_t acc = wi & 1;
      const int nt = (int)(w % p.n_tiles);
      long long t = w / p.n_tiles;
      const int wt = (int)(t % p.ntw); t /= p.ntw;
      const int ht = (int)(t % p.nth); t /= p.nth;
      const int dt = (int)(t % p.ntd);
      const int nn = (int)(t / p.ntd);
      const int gw_ = wt * p.tw + lw, gh_ = ht * p.th + lh, gd_ = dt * p.td + ld_;
      const bool valid = gw_ < p.gw && gh_ < p.gh && gd_ < p.gd;
      const long long opix = (((long long)nn * p.od + (gd_ * p.os + p.ood)) * p.oh + (gh_ * p.os + p.ooh)) * p.ow +
                             (gw_ * p.os + p.oow);
      mbar_wait(&tfull[acc], (wi >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.n_tile;
      const int c_base = nt * p.n_tile;
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (c0 + 16 >= p.n_tile) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        const int c = c_base + c0;
        if (!valid || c >= p.cout) continue;
        if (p.out_dtype == URN_F32) {
          float* o = reinterpret_cast<float*>(out) + opix * p.ldout + c;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 x = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.accumulate) {
              float4 y = *reinterpret_cast<float4*>(o + j);
              x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            }
            *reinterpret_cast<float4*>(o + j) = x;
          }
        } else {
          bf16* o = reinterpret_cast<bf16*>(out) + opix * p.ldout + c;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 x = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.accumulate) {
              float4 y = ld4(o + j);
              x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            }
            st4(o + j, x);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------ host side ------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

}  // namespace

bool gconv_tc_supported(const GConv& g, int in_dtype, int out_dtype) {
  if (in_dtype != URN_BF16 && in_dtype != URN_F16) return false;
  if (out_dtype != URN_F32 && out_dtype != URN_BF16) return false;
  if (g.cin % 16 || g.cout % 16) return false;
  if (g.ldin % 8 || g.ldout % 4) return false;
  if (g.ntaps < 1) return false;
  return true;
}

// wpack: 16-bit filter (same dtype as `in`) laid out [tap][cout][cin]
int gconv_tc_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                    int out_dtype, cudaStream_t st) {
  URN_CHECK(gconv_tc_supported(g, in_dtype, out_dtype), "gconv_tc: unsupported shape/dtype");
  EncodeTiledFn enc = get_encode();
  URN_CHECK(enc, "gconv_tc: cuTensorMapEncodeTiled not available");
  long long M = (long long)g.n * g.gd * g.gh * g.gw;
  if (M == 0) return 0;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n = g.n; p.gd = g.gd; p.gh = g.gh; p.gw = g.gw;
  // tile box: powers of two with product 128, W first
  p.tw = std::min(pow2_ceil(g.gw), 128);
  if (p.tw > 16 && g.gh > 1) p.tw = 16;                       // keep boxes 2-D for halo locality
  if (g.is * p.tw > 256) p.tw = 256 / g.is;
  p.th = std::min(pow2_ceil(g.gh), 128 / p.tw);
  p.td = 128 / (p.tw * p.th);
  if (p.td > pow2_ceil(g.gd)) {                                // tiny grids: widen W/H again
    p.td = pow2_ceil(g.gd);
    p.th = std::min(pow2_ceil(g.gh), 128 / (p.tw * p.td));
    p.tw = 128 / (p.th * p.td);
  }
  URN_CHECK(p.tw * p.th * p.td == 128 && p.tw * g.is <= 256 && p.th * g.is <= 256 && p.td * g.is <= 256,
            "gconv_tc: cannot tile the output grid (%d,%d,%d)", g.gd, g.gh, g.gw);
  p.ntw = (g.gw + p.tw - 1) / p.tw; p.nth = (g.gh + p.th - 1) / p.th; p.ntd = (g.gd + p.td - 1) / p.td;
  p.ntiles = (long long)g.n * p.ntd * p.nth * p.ntw;
  p.is = g.is; p.os = g.os; p.ood = g.ood; p.ooh = g.ooh; p.oow = g.oow;
  p.od = g.od; p.oh = g.oh; p.ow = g.ow; p.ldout = g.ldout;
  p.cout = g.cout;
  p.n_tile = g.cout <= 256 ? g.cout : 256;
  if (g.cout > 256 && g.cout % 256) p.n_tile = 128;
  p.n_tiles = (g.cout + p.n_tile - 1) / p.n_tile;
  p.kc = g.cin % 64 == 0 ? 64 : (g.cin % 32 == 0 ? 32 : 16);
  p.nchunks = g.cin / p.kc;
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) {
    p.tap_od[t] = g.taps[t].od; p.tap_oh[t] = g.taps[t].oh; p.tap_ow[t] = g.taps[t].ow; p.tap_w[t] = g.taps[t].widx;
  }
  p.accumulate = g.accumulate; p.out_dtype = out_dtype;
  p.a_bytes = 128 * p.kc * 2;
  p.b_bytes = ((p.n_tile * p.kc * 2 + 1023) / 1024) * 1024;
  p.tx_bytes = p.a_bytes + p.n_tile * p.kc * 2;
  int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = std::max(2, std::min(kMaxStages, (160 * 1024) / stage_bytes));
  const int fmt = in_dtype == URN_BF16 ? 1 : 0;
  p.idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(p.n_tile >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  const uint64_t layout = p.kc == 64 ? 2 : (p.kc == 32 ? 4 : 6);          // SWIZZLE_128B / 64B / 32B
  const uint64_t sbo = (uint64_t)(8 * p.kc * 2) >> 4;                      // 8 rows of kc elements
  p.desc_hi = (layout << 61) | (1ull << 46) | (sbo << 32) | (1ull << 16);
  p.tmem_cols = std::max(32, pow2_ceil(2 * p.n_tile));
  URN_CHECK(p.tmem_cols <= 512, "gconv_tc: accumulator does not fit TMEM");

  const CUtensorMapDataType dt = in_dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[5] = {(cuuint64_t)g.cin, (cuuint64_t)g.iw, (cuuint64_t)g.ih, (cuuint64_t)g.id, (cuuint64_t)g.n};
    cuuint64_t strides[4] = {(cuuint64_t)g.ldin * 2, (cuuint64_t)g.ldin * 2 * g.iw,
                             (cuuint64_t)g.ldin * 2 * g.iw * g.ih, (cuuint64_t)g.ldin * 2 * g.iw * g.ih * g.id};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, (cuuint32_t)(p.tw * g.is), (cuuint32_t)(p.th * g.is), (cuuint32_t)(p.td * g.is), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)g.is, (cuuint32_t)g.is, (cuuint32_t)g.is, 1};
    CUresult r = enc(&map_a, dt, 5, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    const int T = 27;  // taps are addressed by their filter index; the packed filter has k^nd <= 27 of them
    cuuint64_t dims[3] = {(cuuint64_t)g.cin, (cuuint64_t)g.cout, (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)g.cin * 2, (cuuint64_t)g.cin * 2 * g.cout};
    cuuint32_t box[3] = {(cuuint32_t)p.kc, (cuuint32_t)p.n_tile, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_b, dt, 3, const_cast<void*>(wpack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)p.stages * stage_bytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    URN_CUDA(cudaFuncSetAttribute(gconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  long long nwork = p.ntiles * p.n_tiles;
  int grid = (int)std::min<long long>(nwork, num_sms());
  gconv_tc_kernel<<<grid, kThreads, smem, st>>>(map_a, map_b, p, out);
  URN_LAUNCH_CHECK();
  return 0;
}
