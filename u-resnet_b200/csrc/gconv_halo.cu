// Halo-tile tcgen05 convolution for the narrow layers (Cin in {16,32,64}, Cout <= 64,
// stride 1, 3x3 / 3x3x3 taps, forward and input-gradient).
//
// Why: the tap-per-TMA-box kernel (gconv_tc.cu) issues 9 (2D) or 27 (3D) tap-shifted boxes of
// 128 rows per tile and is bound by the TMA box-row rate when the rows are only 32-64 B.  Here
// each input plane of a tile is staged ONCE, by four producer warps with 16-byte cp.async
// (zero-fill for TF 'same' padding), into a swizzled K-major shared-memory layout whose rows
// are ordered (hy, kw, w): for every halo row hy the three horizontally shifted 8-pixel groups
// kw = 0,1,2 sit next to each other.  The operand of tap (kh,kw) is then just a descriptor
// view -- start (kh*3+kw)*8 rows, 8-row groups 24 rows apart -- so all 9 in-plane taps (and in
// 3D the 3 depth taps, which accumulate into different TMEM accumulators of a TD-plane tile)
// reuse the same staged plane.  Filters for all taps stay resident in shared memory.
//   warps 0-7 : cp.async producers          warp 8 : TMEM alloc, filter TMA, MMA issue
//   warps 9-12: epilogue (tcgen05.ld -> fp32 raw output / bf16 gradient, accumulate flag)
// Every CTA owns a contiguous range of tiles; ring indices and tile coordinates advance
// incrementally (no divisions in the role loops).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"
#include "halo_common.cuh"

namespace {

constexpr int kHThreads = 416;       // 13 warps: 8 producers, 1 MMA, 4 epilogue
constexpr int kHMaxStages = 8;
constexpr int kMaxAcc = 8;           // TMEM accumulator ring (hides the MMA<->epilogue hand-off latency)
constexpr int kMaxPF = 4;            // cp.async groups kept in flight per producer thread (<= stages-1)
constexpr int kTH = 16, kTW = 8;     // 128 positions per plane tile
constexpr int kHaloRows = (kTH + 2) * 3 * kTW;   // 432 rows per staged plane

struct HaloParams {
  int n, D, H, W;
  int cin, cout;
  long long ldin, ldout;
  int nkd;                   // 1 (2D) or 3 (3D)
  int TD;                    // output planes per tile
  int ntw, nth, ntd;
  long long ntiles;
  int stages, plane_bytes, w_tap_bytes, ntaps_w;
  int tapw[27];              // packed-filter tap index for offset (od+1,oh+1,ow+1); -1 = absent
  int accumulate, out_dtype, in_dtype;
  uint32_t idesc;
  uint64_t adesc_hi, bdesc_hi;
  int tmem_cols;
  int nacc;                  // accumulator buffers in the TMEM ring
  int dbg;                   // URN_HALO_DBG bitmask (perf experiments): 1 no stores, 2 no loads, 4 no MMA
};

// Tile iterator over a contiguous range of tiles: coordinates advance incrementally (the role
// loops run in very few threads, where an integer division costs hundreds of cycles).
struct TileIter {
  int nn, dt, ht, wt;
  __device__ __forceinline__ void init(const HaloParams& p, long long t) {
    wt = (int)(t % p.ntw); t /= p.ntw;
    ht = (int)(t % p.nth); t /= p.nth;
    dt = (int)(t % p.ntd);
    nn = (int)(t / p.ntd);
  }
  __device__ __forceinline__ void next(const HaloParams& p) {
    if (++wt == p.ntw) { wt = 0; if (++ht == p.nth) { ht = 0; if (++dt == p.ntd) { dt = 0; ++nn; } } }
  }
};
constexpr int kNumProducerWarps = 8;
constexpr int kProducerThreads = kNumProducerWarps * 32;
constexpr int kMmaWarp = kNumProducerWarps;            // warp 8
constexpr int kEpiWarp0 = kNumProducerWarps + 1;       // warps 9..12
constexpr int kMaxChunks = (kHaloRows * 8 + kProducerThreads - 1) / kProducerThreads;   // cin = 64

template <int NCH>   // 16-byte chunks staged per producer thread and plane: 4 / 7 / 14 for Cin 16 / 32 / 64
__global__ void __launch_bounds__(kHThreads, 1)
gconv_halo_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ HaloParams p,
                  const void* __restrict__ in, void* __restrict__ out, long long* __restrict__ tl) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                             // resident filters
  // optional timeline (URN_HALO_TIMELINE): CTA 0 records clock64() at role checkpoints of its first 64 tiles
#define TL(role, k, pt) do { if (tl && blockIdx.x == 0 && (k) < 64) tl[((role) * 64 + (k)) * 8 + (pt)] = clock64(); } while (0)
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  uint8_t* smem_a = smem + w_bytes;                                   // staged planes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * p.plane_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kHMaxStages;
  uint64_t* tfull = bars + 2 * kHMaxStages;
  uint64_t* tempty = tfull + kMaxAcc;
  uint64_t* wfull = tempty + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = p.cin * 2;                  // bytes per staged row
  const int cpr = pitch / 16;                   // 16-byte chunks per row
  const int od_lo = p.nkd == 3 ? -1 : 0, od_hi = p.nkd == 3 ? 1 : 0;
  // contiguous tile range of this CTA
  const long long t_begin = p.ntiles * (long long)blockIdx.x / gridDim.x;
  const long long t_end = p.ntiles * (long long)(blockIdx.x + 1) / gridDim.x;
  const int my_tiles = (int)(t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { h_mbar_init(&full[s], kNumProducerWarps); h_mbar_init(&empty[s], 1); }
    for (int a = 0; a < p.nacc; ++a) { h_mbar_init(&tfull[a], 1); h_mbar_init(&tempty[a], 4); }
    h_mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  h_fence_before();
  __syncthreads();
  h_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kNumProducerWarps) {
    // ===================== producers: stage one input plane per pipeline slot =====================
    // Per-thread chunk descriptors are tile-invariant: dst offset (16 B units), halo row, halo
    // column and chunk-in-row, packed into one register each.
    const int t = threadIdx.x;
    const int total = kHaloRows * cpr;
    const int cpr_log2 = cpr == 8 ? 3 : (cpr == 4 ? 2 : 1);
    const int sw_shift = cpr == 8 ? 0 : (cpr == 4 ? 1 : 2);
    const int sw_mask = cpr == 8 ? 7 : (cpr == 4 ? 3 : 1);
    uint32_t desc[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int idx = t + j * kProducerThreads;
      const int row = idx >> cpr_log2, c = idx & (cpr - 1);
      const int hy = (row * 2731) >> 16;                       // row / 24, exact for row < 432
      const int rem = row - hy * 24;
      const int dx = rem - 7 * (rem >> 3);                     // kw + w
      const int sw = (row >> sw_shift) & sw_mask;              // K-major SW128 / SW64 / SW32 chunk swizzle
      const uint32_t dst16 = (uint32_t)(row * (pitch >> 4) + (c ^ sw));
      desc[j] = idx < total ? (dst16 | ((uint32_t)hy << 12) | ((uint32_t)dx << 17) | ((uint32_t)c << 21) | (1u << 31)) : 0u;
    }
    const char* src_base = reinterpret_cast<const char*>(in);
    const int pix_bytes = (int)(p.ldin * 2);        // one plane is < 2 GB: 32-bit offsets inside it
    const int row_bytes = pix_bytes * p.W;
    TileIter ti; ti.init(p, t_begin);
    Ring st, ar; st.init(p.stages); ar.init(p.stages);
    uint32_t issued = 0;
    const int pf = max(1, min(kMaxPF, p.stages / 2));
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      const int ybase = ti.ht * kTH - 1, xbase = ti.wt * kTW - 1;
      const int zhi = min(d0 + p.TD, p.D) - 1 + od_hi;
      for (int z = d0 + od_lo; z <= zhi; ++z) {
        if (z < 0 || z >= p.D) continue;                 // a plane of zeros contributes nothing
        h_mbar_wait_warp(&empty[st.idx], st.phase ^ 1);
        if (t == 0) TL(0, k, 0);
        const uint32_t dst_base = h_smem_u32(smem_a + (size_t)st.idx * p.plane_bytes);
        const char* plane_ptr = src_base + ((long long)ti.nn * p.D + z) * p.H * (long long)row_bytes;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const uint32_t dsc = desc[j];
          if (!(dsc >> 31)) continue;
          const int y = ybase + (int)((dsc >> 12) & 31), x = xbase + (int)((dsc >> 17) & 15);
          const bool ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
          const char* src = ok ? plane_ptr + (y * row_bytes + x * pix_bytes + (int)((dsc >> 21) & 7) * 16) : src_base;
          if (!(p.dbg & 2)) cp_async16_zfill(dst_base + ((dsc & 0xFFFu) << 4), src, ok ? 16u : 0u);
        }
        // One barrier arrival per WARP (256 per-thread arrivals on one mbarrier word serialise
        // for ~2 us per stage): each thread waits until its copies of the stage issued kPF stages
        // ago have landed, the warp converges, lane 0 arrives.  The consumer issues the
        // generic->async proxy fence after its barrier wait.
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (t == 0) TL(0, k, 1);
        if (issued >= (uint32_t)pf) {
          if (pf >= 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
          else if (pf == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
          else if (pf == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
          else asm volatile("cp.async.wait_group 1;" ::: "memory");
          __syncwarp();
          if (lane == 0) h_mbar_arrive(&full[ar.idx]);
          ar.next();
        }
        if (t == 0) TL(0, k, 2);
        ++issued;
        st.next();
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (uint32_t j = (issued > (uint32_t)pf ? issued - pf : 0); j < issued; ++j) { h_mbar_arrive(&full[ar.idx]); ar.next(); }
  } else if (warp == kMmaWarp) {
    // ===================== filter load + MMA issue =====================
    // The WHOLE warp runs this loop with warp-uniform values (so the address arithmetic lives in
    // the uniform datapath and no lane-0 branch serialises it); only the tcgen05 instructions
    // themselves are issued by one elected lane.
    if (h_elect_one()) {
      h_mbar_expect_tx(wfull, (uint32_t)(p.ntaps_w * p.w_tap_bytes));
      for (int tp = 0; tp < p.ntaps_w; ++tp) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(h_smem_u32(smem_w + (size_t)tp * p.w_tap_bytes)), "l"(reinterpret_cast<uint64_t>(&map_w)),
            "r"(h_smem_u32(wfull)), "r"(0), "r"(0), "r"(tp)
            : "memory");
      }
    }
    __syncwarp();
    // filter descriptor low words of the 27 (od,kh,kw) taps (all present in this kernel)
    const uint32_t w_lo0 = h_smem_u32(smem_w) >> 4;
    const uint32_t b_hi = (uint32_t)(p.bdesc_hi >> 32), a_hi = (uint32_t)(p.adesc_hi >> 32);
    const uint32_t b_lo_base = (uint32_t)p.bdesc_hi, a_lo_base = (uint32_t)p.adesc_hi;
    uint32_t boff[27];
#pragma unroll
    for (int i = 0; i < 27; ++i)
      boff[i] = b_lo_base | (w_lo0 + (uint32_t)((p.tapw[i] < 0 ? 0 : p.tapw[i]) * (p.w_tap_bytes >> 4)));
    const uint32_t a_step = (uint32_t)(8 * pitch) >> 4;        // one (kh,kw) view further
    const int ksteps = p.cin >> 4;
    const bool no_mma = (p.dbg & 4) != 0;
    h_mbar_wait(wfull, 0);     // all lanes poll: keeps the warp converged and its values in the uniform datapath
    TileIter ti; ti.init(p, t_begin);
    Ring st, ac; st.init(p.stages); ac.init(p.nacc);
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      if (lane == 0) TL(1, k, 0);
      h_mbar_wait(&tempty[ac.idx], ac.phase ^ 1);
      h_fence_after();
      if (lane == 0) TL(1, k, 1);
      uint32_t touched = 0;                              // bit dd: accumulator of plane dd written
      const int nd_out = min(d0 + p.TD, p.D) - d0;       // output planes of this tile
      const uint32_t d_tmem0 = tmem_base + (uint32_t)(ac.idx * p.TD * p.cout);
      const int zhi = d0 + nd_out - 1 + od_hi;
      for (int z = d0 + od_lo; z <= zhi; ++z) {
        if (z < 0 || z >= p.D) continue;
        h_mbar_wait(&full[st.idx], st.phase);
        if (lane == 0) TL(1, k, 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> UMMA (async proxy)
        h_fence_after();
        if (lane == 0) TL(1, k, 3);
        const uint32_t a_lo0 = a_lo_base | (h_smem_u32(smem_a + (size_t)st.idx * p.plane_bytes) >> 4);
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          if (p.nkd == 1 && oi != 1) continue;
          const int dd = z - (oi - 1) - d0;              // od = oi - 1
          if (dd < 0 || dd >= nd_out) continue;
          const uint32_t d_tmem = d_tmem0 + (uint32_t)(dd * p.cout);
          const uint32_t accf = (touched >> dd) & 1u;
          touched |= 1u << dd;
          if (!no_mma && h_elect_one()) {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const uint32_t a_lo = a_lo0 + (uint32_t)j * a_step, b_lo = boff[oi * 9 + j];
              h_umma2(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, j == 0 ? accf : 1u);
              if (ksteps > 1) h_umma2(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, p.idesc, 1u);
              if (ksteps > 2) {
                h_umma2(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, p.idesc, 1u);
                h_umma2(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, p.idesc, 1u);
              }
            }
          }
          __syncwarp();
        }
        if (lane == 0) TL(1, k, 4);
        if (h_elect_one()) h_commit(&empty[st.idx]);
        __syncwarp();
        st.next();
      }
      if (h_elect_one()) h_commit(&tfull[ac.idx]);
      __syncwarp();
      if (lane == 0) TL(1, k, 5);
      ac.next();
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int lh = r >> 3, lw = r & 7;
    TileIter ti; ti.init(p, t_begin);
    Ring ac; ac.init(p.nacc);
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      const int y = ti.ht * kTH + lh, x = ti.wt * kTW + lw;
      const bool valid_hw = y < p.H && x < p.W;
      const int dmax = min(d0 + p.TD, p.D);
      if (warp == kEpiWarp0 && lane == 0) TL(2, k, 0);
      h_mbar_wait_warp(&tfull[ac.idx], ac.phase);
      h_fence_after();
      if (warp == kEpiWarp0 && lane == 0) TL(2, k, 1);
      for (int dd = 0; dd < p.TD; ++dd) {
        const int d = d0 + dd;
        if (d >= dmax) continue;                       // warp-uniform: accumulator never written
        const long long opix = (((long long)ti.nn * p.D + d) * p.H + y) * p.W + x;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ac.idx * p.TD + dd) * p.cout);
        for (int c0 = 0; c0 < p.cout; c0 += 16) {
          float v[16];
          h_tmem_ld16(taddr + c0, v);
          if (!valid_hw || (p.dbg & 1)) continue;
          if (p.out_dtype == URN_F32) {
            float* o = reinterpret_cast<float*>(out) + opix * p.ldout + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 xx = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.accumulate) {
                float4 yy = *reinterpret_cast<float4*>(o + j);
                xx.x += yy.x; xx.y += yy.y; xx.z += yy.z; xx.w += yy.w;
              }
              *reinterpret_cast<float4*>(o + j) = xx;
            }
          } else {
            bf16* o = reinterpret_cast<bf16*>(out) + opix * p.ldout + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 xx = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.accumulate) {
                float4 yy = ld4(o + j);
                xx.x += yy.x; xx.y += yy.y; xx.z += yy.z; xx.w += yy.w;
              }
              st4(o + j, xx);
            }
          }
        }
      }
      if (warp == kEpiWarp0 && lane == 0) TL(2, k, 2);
      h_fence_before();
      __syncwarp();
      if (lane == 0) h_mbar_arrive(&tempty[ac.idx]);
      if (warp == kEpiWarp0 && lane == 0) TL(2, k, 3);
      ac.next();
    }
  }

  h_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    h_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

bool halo_plan(const GConv& g, int in_dtype, int out_dtype, HaloParams* pp) {
  HaloParams& p = *pp;
  memset(&p, 0, sizeof(p));
  if (getenv("URN_NO_HALO")) return false;
  if (in_dtype != URN_BF16 && in_dtype != URN_F16) return false;
  if (out_dtype != URN_F32 && out_dtype != URN_BF16) return false;
  if (g.is != 1 || g.os != 1 || g.ood || g.ooh || g.oow) return false;
  if (g.id != g.od || g.ih != g.oh || g.iw != g.ow || g.gd != g.od || g.gh != g.oh || g.gw != g.ow) return false;
  if (g.cin != 16 && g.cin != 32) return false;   // at Cin=64 the tap-box kernel (128 B TMA rows) is as fast
  if (g.cout % 16 || g.cout > 64) return false;
  if (g.ldin % 8 || g.ldout % 4) return false;
  if (g.ntaps != 9 && g.ntaps != 27) return false;
  for (int i = 0; i < 27; ++i) p.tapw[i] = -1;
  bool has_d = false;
  int maxw = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    const GTap& tp = g.taps[t];
    if (tp.od < -1 || tp.od > 1 || tp.oh < -1 || tp.oh > 1 || tp.ow < -1 || tp.ow > 1) return false;
    if (tp.od != 0) has_d = true;
    int& slot = p.tapw[(tp.od + 1) * 9 + (tp.oh + 1) * 3 + (tp.ow + 1)];
    if (slot >= 0) return false;
    slot = tp.widx;
    maxw = std::max(maxw, tp.widx);
  }
  if (has_d != (g.ntaps == 27)) return false;
  p.nkd = has_d ? 3 : 1;
  p.n = g.n; p.D = g.od; p.H = g.oh; p.W = g.ow;
  p.cin = g.cin; p.cout = g.cout; p.ldin = g.ldin; p.ldout = g.ldout;
  p.TD = has_d ? (g.cout <= 32 ? 4 : 2) : 1;
  if (has_d && p.D < p.TD) p.TD = std::max(1, p.D);
  p.ntw = (p.W + kTW - 1) / kTW; p.nth = (p.H + kTH - 1) / kTH; p.ntd = (p.D + p.TD - 1) / p.TD;
  p.ntiles = (long long)p.n * p.ntd * p.nth * p.ntw;
  p.plane_bytes = kHaloRows * g.cin * 2;
  p.w_tap_bytes = g.cout * g.cin * 2;
  p.ntaps_w = maxw + 1;
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  const int budget = 198 * 1024 - w_bytes;
  p.stages = std::min(kHMaxStages, budget / p.plane_bytes);
  if (p.stages < 2) return false;
  if (p.ntiles >= (1ll << 31)) return false;
  p.accumulate = g.accumulate; p.out_dtype = out_dtype; p.in_dtype = in_dtype;
  const int fmt = in_dtype == URN_BF16 ? 1 : 0;
  p.idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(g.cout >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  const uint64_t pitch = (uint64_t)g.cin * 2;
  const uint64_t layout = pitch == 128 ? 2 : (pitch == 64 ? 4 : 6);
  p.adesc_hi = (layout << 61) | (1ull << 46) | (((24 * pitch) >> 4) << 32) | (1ull << 16);   // 8-row groups 24 rows apart
  p.bdesc_hi = (layout << 61) | (1ull << 46) | (((8 * pitch) >> 4) << 32) | (1ull << 16);
  p.nacc = std::max(2, std::min(kMaxAcc, 512 / (p.TD * g.cout)));
  if (const char* e = getenv("URN_HALO_NACC")) p.nacc = std::max(2, std::min(p.nacc, atoi(e)));
  p.tmem_cols = std::max(32, pow2c(p.nacc * p.TD * g.cout));
  if (p.tmem_cols > 512) return false;
  if (const char* e = getenv("URN_HALO_DBG")) p.dbg = atoi(e);
  return true;
}

}  // namespace

bool gconv_halo_supported(const GConv& g, int in_dtype, int out_dtype) {
  HaloParams p;
  return halo_plan(g, in_dtype, out_dtype, &p);
}

// wpack: 16-bit filter (dtype of `in`) laid out [tap][cout][cin]
int gconv_halo_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                      int out_dtype, cudaStream_t st) {
  HaloParams p;
  URN_CHECK(halo_plan(g, in_dtype, out_dtype, &p), "gconv_halo: unsupported shape/dtype");
  if (p.ntiles == 0) return 0;
  EncodeTiledFnH enc = get_encode_h();
  URN_CHECK(enc, "gconv_halo: cuTensorMapEncodeTiled not available");
  CUtensorMap map_w;
  {
    const CUtensorMapDataType dt = in_dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw = g.cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (g.cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[3] = {(cuuint64_t)g.cin, (cuuint64_t)g.cout, (cuuint64_t)p.ntaps_w};
    cuuint64_t strides[2] = {(cuuint64_t)g.cin * 2, (cuuint64_t)g.cin * 2 * g.cout};
    cuuint32_t box[3] = {(cuuint32_t)g.cin, (cuuint32_t)g.cout, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_w, dt, 3, const_cast<void*>(wpack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_halo: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  const size_t smem = 1024 + (size_t)w_bytes + (size_t)p.stages * p.plane_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    URN_CUDA(cudaFuncSetAttribute(gconv_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    URN_CUDA(cudaFuncSetAttribute(gconv_halo_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    URN_CUDA(cudaFuncSetAttribute(gconv_halo_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = (int)std::min<long long>(p.ntiles, num_sms());
  long long* tl = nullptr;
  if (const char* e = getenv("URN_HALO_TIMELINE")) tl = (long long*)strtoull(e, nullptr, 0);
  if (g.cin == 16) gconv_halo_kernel<4><<<grid, kHThreads, smem, st>>>(map_w, p, in, out, tl);
  else if (g.cin == 32) gconv_halo_kernel<7><<<grid, kHThreads, smem, st>>>(map_w, p, in, out, tl);
  else gconv_halo_kernel<14><<<grid, kHThreads, smem, st>>>(map_w, p, in, out, tl);
  URN_LAUNCH_CHECK();
  return 0;
}
