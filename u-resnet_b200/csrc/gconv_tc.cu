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
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"
#include "bn_stat_fuse.cuh"

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
  float* stats_part;            // fused BN statistics partials [(cta*4 + warp)][cout][3], or NULL
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

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// All lanes wait: a lane-0 poll + __syncwarp costs 3-6x more per hand-off (tools/probe/pingpong.cu).
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// descriptors passed as 32-bit halves (the high halves are loop invariants)
__device__ __forceinline__ void umma_f16_2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
struct Ring {
  int idx; uint32_t phase; int n;
  __device__ __forceinline__ void init(int n_) { idx = 0; phase = 0; n = n_; }
  __device__ __forceinline__ void next() { if (++idx == n) { idx = 0; phase ^= 1u; } }
};

// NOTE on the role loops below: the producer and MMA warps run their loops with ALL lanes and
// warp-uniform values (address arithmetic stays in the uniform datapath); only the TMA / tcgen05
// instructions are issued by one elected lane.  A `if (lane == 0)` structure around the whole
// loop costs ~4x more cycles per pipeline stage (measured: 3370 -> 830 cycles per 9 MMAs).

// ------------------------------ the kernel ------------------------------
// STATS: BN statistics of the fp32 output fused into the epilogue (one N tile per layer only).
template <bool STATS>
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
    Ring st; st.init(p.stages);
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      const unsigned wu = (unsigned)w;
      const unsigned q1 = wu / (unsigned)p.n_tiles; const int nt = (int)(wu - q1 * (unsigned)p.n_tiles);
      const unsigned q2 = q1 / (unsigned)p.ntw;     const int wt = (int)(q1 - q2 * (unsigned)p.ntw);
      const unsigned q3 = q2 / (unsigned)p.nth;     const int ht = (int)(q2 - q3 * (unsigned)p.nth);
      const unsigned q4 = q3 / (unsigned)p.ntd;     const int dt = (int)(q3 - q4 * (unsigned)p.ntd);
      const int nn = (int)q4;
      const int w0 = wt * p.tw * p.is, h0 = ht * p.th * p.is, d0 = dt * p.td * p.is;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int cw = w0 + p.tap_ow[tap], chh = h0 + p.tap_oh[tap], cd = d0 + p.tap_od[tap], tw_ = p.tap_w[tap];
        for (int ch = 0; ch < p.nchunks; ++ch) {
          mbar_wait(&empty[st.idx], st.phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full[st.idx], (uint32_t)p.tx_bytes);
            tma_load_5d(smem_a + (size_t)st.idx * p.a_bytes, &map_a, &full[st.idx], ch * p.kc, cw, chh, cd, nn);
            tma_load_3d(smem_b + (size_t)st.idx * p.b_bytes, &map_b, &full[st.idx], ch * p.kc, nt * p.n_tile, tw_);
          }
          __syncwarp();
          st.next();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    Ring st, ac; st.init(p.stages); ac.init(2);
    const uint32_t hi = (uint32_t)(p.desc_hi >> 32), lo_base = (uint32_t)p.desc_hi;
    const int kper = p.kc / 16;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      mbar_wait(&tempty[ac.idx], ac.phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)ac.idx * (uint32_t)p.n_tile;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full[st.idx], st.phase);
        tc_fence_after();
        const uint32_t a_lo = lo_base | (smem_u32(smem_a + (size_t)st.idx * p.a_bytes) >> 4);
        const uint32_t b_lo = lo_base | (smem_u32(smem_b + (size_t)st.idx * p.b_bytes) >> 4);
        if (elect_one()) {
          for (int k = 0; k < kper; ++k)
            umma_f16_2(d_tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, p.idesc, (ks | k) ? 1u : 0u);
          umma_commit(&empty[st.idx]);          // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        st.next();
      }
      if (elect_one()) umma_commit(&tfull[ac.idx]);   // accumulator complete -> epilogue
      __syncwarp();
      ac.next();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;              // accumulator row = position inside the tile box
    const int lw = r % p.tw, lh = (r / p.tw) % p.th, ld_ = r / (p.tw * p.th);
    uint32_t wi = 0;
    RunStat rs[STATS ? 16 : 1];
    if (STATS) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { rs[i].n = 0.f; rs[i].mean = 0.f; rs[i].m2 = 0.f; }
    }
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
      mbar_wait_lane0(&tfull[acc], (wi >> 1) & 1);
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
        if (STATS) {
          float a[16], b[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[j] = valid ? v[j] : 0.f; b[j] = a[j] * a[j]; }
          const unsigned nv = __popc(__ballot_sync(0xffffffffu, valid));
#pragma unroll
          for (int ci = 0; ci < 16; ++ci)
            if (ci * 16 == c0) stat_merge(rs[ci], a, b, nv, lane);
        }
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
        } else if (p.out_dtype == URN_F16) {      // fp16 raw output (mix16); the statistics above saw the fp32 values
          f16* o = reinterpret_cast<f16*>(out) + opix * p.ldout + c;
#pragma unroll
          for (int j = 0; j < 16; j += 4) st4(o + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
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
    if (STATS && !(lane & 1)) {
      // lane L owns channel ci*16 + (L >> 1) of every 16-channel chunk; partial[(cta*4 + q)][cout][3]
      float* dst = p.stats_part + ((long long)(blockIdx.x * 4 + q) * p.cout) * 3;
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const int ch = ci * 16 + (lane >> 1);
        if (ch < p.cout) { dst[ch * 3 + 0] = rs[ci].n; dst[ch * 3 + 1] = rs[ci].mean; dst[ch * 3 + 2] = rs[ci].m2; }
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
  if (out_dtype != URN_F32 && out_dtype != URN_BF16 && out_dtype != URN_F16) return false;
  if (out_dtype == URN_F16 && g.accumulate) return false;        // fp16 = freshly written raw output
  if (g.cin % 16 || g.cout % 16) return false;
  if (g.ldin % 8 || g.ldout % 4) return false;
  if (g.ntaps < 1) return false;
  return true;
}

// wpack: 16-bit filter (same dtype as `in`) laid out [tap][cout][cin]
int gconv_tc_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                    int out_dtype, cudaStream_t st) {
  return gconv_tc_launch_stats(g, in_dtype, in, wpack, out, out_dtype, nullptr, nullptr, st);
}

int gconv_tc_launch_stats(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                          int out_dtype, float* stats_part, int* nparts, cudaStream_t st) {
  if (nparts) *nparts = 0;
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
  if (const char* e = getenv("URN_TC_STAGES")) p.stages = std::max(1, std::min(p.stages, atoi(e)));
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
    URN_CUDA(cudaFuncSetAttribute(gconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    URN_CUDA(cudaFuncSetAttribute(gconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  long long nwork = p.ntiles * p.n_tiles;
  URN_CHECK(nwork < (1ll << 31), "gconv_tc: too many tiles");
  int grid = (int)std::min<long long>(nwork, persistent_ctas());
  // fused statistics: a single launch writing the whole fp32 tensor once, one N tile of <= 256 channels
  const bool stats = stats_part && nparts && out_dtype != URN_BF16 && !g.accumulate && g.os == 1 && g.gd == g.od &&
                     g.gh == g.oh && g.gw == g.ow && p.n_tiles == 1 && p.n_tile <= 256 && !getenv("URN_NO_FUSED_STATS");
  if (stats) {
    p.stats_part = stats_part;
    *nparts = grid * 4;
    gconv_tc_kernel<true><<<grid, kThreads, smem, st>>>(map_a, map_b, p, out);
  } else
  gconv_tc_kernel<false><<<grid, kThreads, smem, st>>>(map_a, map_b, p, out);
  URN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// Filter gradient on tcgen05:  dW[t][cx][cy] = sum_pos X[pos*s + off_t][cx] * Y[pos][cy]
//
// GEMM view: M = (tap, cx) -- 128/AW groups of AW channels, each group one tap-shifted TMA box
// of X; N = cy; K = output positions (a PD x PH x PW box of 64 positions per pipeline stage).
// Both operands are MN-major (the channel axis is the contiguous one in NHWC), which the UMMA
// shared-memory descriptor expresses directly: the TMA boxes [positions][channels] land as
// swizzled K-rows and consecutive channel groups sit LBO bytes apart.  The positions are split
// over CTAs (split-K); every CTA writes an fp32 partial [taps][cx][cy] that a second kernel
// reduces in a fixed order (deterministic).  kind::f16 wants ONE 16-bit format for both
// operands (a mixed fp16 x bf16 instruction descriptor raises an illegal-instruction fault on
// B200), so in mix16 mode the activations enter through their bf16 shadow copy.
// =====================================================================================
namespace {

constexpr int kWgP = 64;   // positions per pipeline stage

struct WgParams {
  int n, yd, yh, yw;          // small-side (Y) grid
  int pd, ph, pw;             // position box per stage (pd*ph*pw == kWgP)
  int nbd, nbh, nbw;          // boxes per dim
  long long nkb;              // K blocks in total
  int s;                      // stride (X sampling)
  int pad_d, pad_h, pad_w, kd, kh, kw;
  int cx, cy;
  int aw, bw;                 // channel-group widths of A (X) and B (Y)
  int gx;                     // cx / aw
  int groups;                 // taps * gx   (M space in groups)
  int gpt;                    // groups per M tile (128 / aw)
  int m_tiles, n_tile, n_tiles, ksplit;
  long long kb_per_split;
  int stages, a_bytes, b_bytes;
  uint32_t idesc;
  uint64_t adesc_hi, bdesc_hi;
  int tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                const __grid_constant__ WgParams p, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
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
  const long long nwork = (long long)p.m_tiles * p.n_tiles * p.ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_y);
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

  // work item -> (m tile, n tile, k split); k split varies fastest so that the CTAs running
  // concurrently stream disjoint positions
  auto decode = [&](long long w, int& mt, int& nt, long long& kb0, long long& kb1) {
    const int ks = (int)(w % p.ksplit);
    long long t = w / p.ksplit;
    nt = (int)(t % p.n_tiles);
    mt = (int)(t / p.n_tiles);
    kb0 = (long long)ks * p.kb_per_split;
    kb1 = kb0 + p.kb_per_split;
    if (kb1 > p.nkb) kb1 = p.nkb;
  };

  if (warp == 0) {
    Ring st; st.init(p.stages);
    const int nb_groups = p.n_tile / p.bw;
    const int a_gbytes = kWgP * p.aw * 2, b_gbytes = kWgP * p.bw * 2;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      int mt, nt; long long kb0, kb1;
      decode(w, mt, nt, kb0, kb1);
      const int g0 = mt * p.gpt;
      const int ng = min(p.gpt, p.groups - g0);
      const uint32_t tx = (uint32_t)(ng * a_gbytes + nb_groups * b_gbytes);
      // per-group box origins of this M tile (fixed for the whole K loop, kept in registers)
      int gcx[8], gow[8], goh[8], god[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int gi = min(g0 + g, p.groups - 1);
        const int tap = gi / p.gx;
        gcx[g] = (gi - tap * p.gx) * p.aw;
        gow[g] = tap % p.kw - p.pad_w;
        goh[g] = (tap / p.kw) % p.kh - p.pad_h;
        god[g] = tap / (p.kw * p.kh) - p.pad_d;
      }
      // position-box iterator (incremental: no divisions inside the K loop)
      long long t = kb0;
      int bw_ = (int)(t % p.nbw); t /= p.nbw;
      int bh_ = (int)(t % p.nbh); t /= p.nbh;
      int bd_ = (int)(t % p.nbd);
      int nn = (int)(t / p.nbd);
      for (long long kb = kb0; kb < kb1; ++kb) {
        const int w0 = bw_ * p.pw, h0 = bh_ * p.ph, d0 = bd_ * p.pd;
        mbar_wait(&empty[st.idx], st.phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[st.idx], tx);
          uint8_t* a_dst = smem_a + (size_t)st.idx * p.a_bytes;
          const int xw = w0 * p.s, xh = h0 * p.s, xd = d0 * p.s;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g < ng)
              tma_load_5d(a_dst + (size_t)g * a_gbytes, &map_x, &full[st.idx], gcx[g], xw + gow[g], xh + goh[g], xd + god[g], nn);
          uint8_t* b_dst = smem_b + (size_t)st.idx * p.b_bytes;
          for (int g = 0; g < nb_groups; ++g)
            tma_load_5d(b_dst + (size_t)g * b_gbytes, &map_y, &full[st.idx], nt * p.n_tile + g * p.bw, w0, h0, d0, nn);
        }
        __syncwarp();
        st.next();
        if (++bw_ == p.nbw) { bw_ = 0; if (++bh_ == p.nbh) { bh_ = 0; if (++bd_ == p.nbd) { bd_ = 0; ++nn; } } }
      }
    }
  } else if (warp == 1) {
    Ring st, ac; st.init(p.stages); ac.init(2);
    const uint32_t a_hi = (uint32_t)(p.adesc_hi >> 32), a_lob = (uint32_t)p.adesc_hi;
    const uint32_t b_hi = (uint32_t)(p.bdesc_hi >> 32), b_lob = (uint32_t)p.bdesc_hi;
    const uint32_t a_kstep = (uint32_t)(16 * p.aw * 2) >> 4, b_kstep = (uint32_t)(16 * p.bw * 2) >> 4;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      int mt, nt; long long kb0, kb1;
      decode(w, mt, nt, kb0, kb1);
      mbar_wait(&tempty[ac.idx], ac.phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)ac.idx * (uint32_t)p.n_tile;
      uint32_t first = 1;
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[st.idx], st.phase);
        tc_fence_after();
        const uint32_t a_lo = a_lob | (smem_u32(smem_a + (size_t)st.idx * p.a_bytes) >> 4);
        const uint32_t b_lo = b_lob | (smem_u32(smem_b + (size_t)st.idx * p.b_bytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kWgP / 16; ++k)   // 16 positions further = 16 K-rows of (group width * 2) bytes
            umma_f16_2(d_tmem, a_lo + k * a_kstep, a_hi, b_lo + k * b_kstep, b_hi, p.idesc, (first && k == 0) ? 0u : 1u);
          umma_commit(&empty[st.idx]);
        }
        __syncwarp();
        first = 0;
        st.next();
      }
      if (elect_one()) umma_commit(&tfull[ac.idx]);
      __syncwarp();
      ac.next();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // accumulator row = (group, channel in group)
    uint32_t wi = 0;
    const long long nw = (long long)p.kd * p.kh * p.kw * p.cx * p.cy;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++wi) {
      int mt, nt; long long kb0, kb1;
      decode(w, mt, nt, kb0, kb1);
      const int ks = (int)(w % p.ksplit);
      const uint32_t acc = wi & 1;
      const int gi = mt * p.gpt + r / p.aw;
      const bool valid = gi < p.groups;
      const int tap = valid ? gi / p.gx : 0;
      const int cx = valid ? (gi % p.gx) * p.aw + r % p.aw : 0;
      float* dst = partial + (long long)ks * nw + ((long long)tap * p.cx + cx) * p.cy + nt * p.n_tile;
      mbar_wait_lane0(&tfull[acc], (wi >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.n_tile;
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (c0 + 16 >= p.n_tile) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (!valid) continue;
        if (kb1 <= kb0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;   // empty split: nothing was accumulated
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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

int make_map_5d(CUtensorMap* map, int dtype, const void* base, int c, long long ld, int w, int h, int d, int n,
                int box_c, int box_w, int box_h, int box_d, int stride) {
  EncodeTiledFn enc = get_encode();
  URN_CHECK(enc, "cuTensorMapEncodeTiled not available");
  const CUtensorMapDataType dt = dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int bytes = box_c * 2;
  const CUtensorMapSwizzle sw = bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h, (cuuint64_t)ld * 2 * w * h * d};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), (cuuint32_t)(box_d * stride), 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(map, dt, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  URN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return 0;
}

__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, long long nw, int nsplit,
                                       float* __restrict__ dw, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += partial[(long long)k * nw + i];
  dw[i] = accumulate ? dw[i] + s : s;
}

int wgrad_tc_plan(const urn_conv_desc& d, WgParams* pp) {
  WgParams& p = *pp;
  memset(&p, 0, sizeof(p));
  if (d.cx % 16 || d.cy % 16) return 1;       // stem / head: CUDA-core kernel
  p.n = d.n; p.yd = d.yd; p.yh = d.yh; p.yw = d.yw;
  p.pw = std::min(pow2_ceil(d.yw), 16);
  p.ph = std::min(pow2_ceil(d.yh), kWgP / p.pw);
  p.pd = kWgP / (p.pw * p.ph);
  if (p.pd > pow2_ceil(d.yd)) {
    p.pd = pow2_ceil(d.yd);
    p.ph = std::min(pow2_ceil(d.yh), kWgP / (p.pw * p.pd));
    p.pw = kWgP / (p.ph * p.pd);
  }
  if (p.pw * p.ph * p.pd != kWgP || p.pw * d.s > 256 || p.ph * d.s > 256 || p.pd * d.s > 256) return 1;
  p.nbw = (d.yw + p.pw - 1) / p.pw; p.nbh = (d.yh + p.ph - 1) / p.ph; p.nbd = (d.yd + p.pd - 1) / p.pd;
  p.nkb = (long long)d.n * p.nbd * p.nbh * p.nbw;
  p.s = d.s; p.pad_d = d.pad_d; p.pad_h = d.pad_h; p.pad_w = d.pad_w; p.kd = d.kd; p.kh = d.kh; p.kw = d.kw;
  p.cx = d.cx; p.cy = d.cy;
  p.aw = d.cx % 64 == 0 ? 64 : (d.cx % 32 == 0 ? 32 : 16);
  p.bw = d.cy % 64 == 0 ? 64 : (d.cy % 32 == 0 ? 32 : 16);
  p.gx = d.cx / p.aw;
  p.groups = d.kd * d.kh * d.kw * p.gx;
  p.gpt = 128 / p.aw;
  p.m_tiles = (p.groups + p.gpt - 1) / p.gpt;
  p.n_tile = d.cy <= 256 ? d.cy : (d.cy % 256 == 0 ? 256 : 128);
  if (d.cy % p.n_tile) return 1;
  p.n_tiles = d.cy / p.n_tile;
  const long long tiles = (long long)p.m_tiles * p.n_tiles;
  const long long nw = (long long)d.kd * d.kh * d.kw * d.cx * d.cy;
  long long ks = std::max<long long>(1, (2ll * 148 + tiles - 1) / tiles);
  ks = std::min<long long>(ks, std::max<long long>(1, p.nkb / 4));
  ks = std::min<long long>(ks, std::max<long long>(1, (64ll << 20) / nw));      // partials <= 256 MB
  p.kb_per_split = (p.nkb + ks - 1) / ks;
  p.ksplit = (int)((p.nkb + p.kb_per_split - 1) / p.kb_per_split);
  p.a_bytes = 128 * kWgP * 2;
  p.b_bytes = ((p.n_tile * kWgP * 2 + 1023) / 1024) * 1024;
  p.stages = std::max(2, std::min(kMaxStages, (160 * 1024) / (p.a_bytes + p.b_bytes)));
  p.tmem_cols = std::max(32, pow2_ceil(2 * p.n_tile));
  return 0;
}

}  // namespace

bool wgrad_tc_supported(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy) {
  if ((x_dtype != URN_BF16 && x_dtype != URN_F16) || x_dtype != y_dtype) return false;
  if (d.cx % 16 || d.cy % 16 || ldx % 8 || ldy % 8) return false;
  WgParams p;
  return wgrad_tc_plan(d, &p) == 0 && p.tmem_cols <= 512;
}

int64_t wgrad_tc_workspace(const urn_conv_desc& d) {
  WgParams p;
  if (wgrad_tc_plan(d, &p)) return 0;
  return (int64_t)p.ksplit * d.kd * d.kh * d.kw * d.cx * d.cy * 4;
}

int wgrad_tc_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                    const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  WgParams p;
  URN_CHECK(wgrad_tc_plan(d, &p) == 0, "wgrad_tc: unsupported geometry");
  const int fa = x_dtype == URN_BF16 ? 1 : 0, fb = y_dtype == URN_BF16 ? 1 : 0;
  p.idesc = (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | (1u << 15) | (1u << 16) |
            ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto desc_hi = [](int gw_elems) {
    const uint64_t bytes = (uint64_t)gw_elems * 2;
    const uint64_t layout = bytes == 128 ? 2 : (bytes == 64 ? 4 : 6);
    const uint64_t sbo = (8 * bytes) >> 4;                 // next 8 positions (K)
    const uint64_t lbo = ((uint64_t)kWgP * bytes) >> 4;   // next channel group (MN)
    return (layout << 61) | (1ull << 46) | (sbo << 32) | (lbo << 16);
  };
  p.adesc_hi = desc_hi(p.aw);
  p.bdesc_hi = desc_hi(p.bw);
  CUtensorMap map_x, map_y;
  if (make_map_5d(&map_x, x_dtype, x, d.cx, ldx, d.xw, d.xh, d.xd, d.n, p.aw, p.pw, p.ph, p.pd, d.s)) return 1;
  if (make_map_5d(&map_y, y_dtype, y, d.cy, ldy, d.yw, d.yh, d.yd, d.n, p.bw, p.pw, p.ph, p.pd, 1)) return 1;
  const size_t smem = 1024 + (size_t)p.stages * (p.a_bytes + p.b_bytes) + 256;
  static bool attr_set = false;
  if (!attr_set) {
    URN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long nwork = (long long)p.m_tiles * p.n_tiles * p.ksplit;
  const int grid = (int)std::min<long long>(nwork, persistent_ctas());
  wgrad_tc_kernel<<<grid, kThreads, smem, st>>>(map_x, map_y, p, (float*)ws);
  URN_LAUNCH_CHECK();
  const long long nw = (long long)d.kd * d.kh * d.kw * d.cx * d.cy;
  wgrad_tc_reduce_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>((const float*)ws, nw, p.ksplit, dw, accumulate);
  URN_LAUNCH_CHECK();
  return 0;
}
