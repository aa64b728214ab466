// Halo-tile tcgen05 filter gradient for the narrow layers (cx, cy in {16, 32, 64}): 3x3 / 3x3x3
// and 1x1 taps, X sampled with stride 1 or 2 (template S2):
//     dW[t][cx][cy] = sum_p X[p*s + t - pad][cx] * dY[p][cy].
//
// Why: the tap-box kernel (wgrad_tc_kernel) issues one TMA box per (tap, channel group) and
// K block, i.e. it re-reads every activation 9 / 27 times through TMA rows of only 32-64 bytes
// and is bound by the TMA box-row rate (0.85 ms per 2D C=16 layer, 7-26 ms in 3D).  Here each
// X plane of a 16 x (8*KS) super-tile is staged ONCE (with its halo, 16-byte cp.async, zero fill =
// TF 'same' padding) into the same (halo row, kw, w)-ordered swizzled layout the forward halo
// kernel uses; the in-plane taps are then descriptor views 8*KS rows apart, which is exactly the
// MN-group stride (LBO) of an MN-major UMMA operand: ONE tcgen05.mma with M = 128 covers 128/cx
// taps (the last M tile of a cx = 16 layer holds one live tap and is issued with M = 64).
//   GEMM view   M = (tap, cx)   N = cy   K = 128 positions of a sub-tile (8 MMAs of K = 16)
//   A = X views (MN-major, SBO = s*RW rows = next output row, LBO = 8*KS rows = next tap)
//   B = dY tile (MN-major, rows = positions)
// In 3D a CTA walks a column of planes: X plane z meets the dY planes of its depth taps.  With
// stride 1 those are z+1, z, z-1: they sit in consecutive slots of the dY ring (slots 0, 1 are
// mirrored behind the ring so that the window never wraps) and are issued as ONE MMA with
// N = 3*cy whose N groups are the ring slots (LBO = slot stride), so the X views are read once
// for all three depth taps.  Accumulators stay in TMEM for the CTA's whole range of tiles; at the
// end every CTA writes one fp32 partial [tap][cx][cy] and a fixed-order second stage sums the
// <= 148 partials (deterministic).
//   warps 0-7 : cp.async producers (X plane + dY tile per step, register-resident offset tables)
//   warp 8    : TMEM alloc, MMA issue        warps 9-12: final epilogue (tcgen05.ld -> fp32 partial)
//
// Round 2.  (i) Images-as-rows plan (wh_plan_rows): a 2D batch of >= 12 images is relabelled as a volume (planes
// := image rows, tile rows := images), so the kh taps become depth taps concatenated along N.  Its operands are
// staged by TMA, ONE box per X plane and per dY plane: rows (image, pixel), positions ordered (image, w), and
// the kw taps are M groups ONE pixel apart (LBO = pixel pitch) -- a UMMA view may start on any 16-byte boundary
// of a swizzled tile.  (ii) Rings: `pf` cp.async groups in flight need pf + 2 stages (see wh_plan).  (iii) 3D
// layers whose 27 x cx x cy accumulators exceed TMEM run one pass per depth tap (wh_split_desc); the depth
// chunks of small grids are balanced over the CTAs.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"
#include "halo_common.cuh"

namespace {

constexpr int kWThreads = 416;
constexpr int kWProdWarps = 8, kWProdThreads = kWProdWarps * 32;
constexpr int kWMmaWarp = kWProdWarps, kWEpiWarp0 = kWProdWarps + 1;
constexpr int kWTH = 16, kWTW = 8;
constexpr int kWNCH = 14;                              // 16-byte X chunks per producer thread and plane
constexpr int kWYCH = 4;                               // 16-byte dY chunks per producer thread and tile
constexpr int kWMaxStages = 8;
constexpr int kYMaxStages = 8;                         // dY ring: 3 planes live + the producers' look-ahead

struct WHParams {
  int n, D, H, W;            // Y (small side) grid: the positions the GEMM's K dimension runs over
  int XD, XH, XW;            // X (big side) dims
  int s;                     // X sampling stride (1 / 2)
  int nkh, nkw, ntp;         // in-plane tap box, taps per plane
  int od0, oh0, ow0;         // first tap offset per dim (= -pad)
  int hr;                    // staged X halo rows = (kWTH-1)*s + nkh
  int cx, cy;
  long long ldx, ldy;
  int nkd, LA;               // depth taps (1 / 3); dY look-ahead in planes (0 / 1)
  int DC, ndc;               // planes per depth chunk, number of chunks
  int ks, rw;                // sub-tiles (16x8 positions) per super-tile along w; staged rows per halo row = 24*ks
  uint32_t idesc_last;       // instruction descriptor of the last M tile (M = 64 when only one 16-channel tap is live)
  uint32_t idesc_n[2][3];    // [last M tile][N blocks - 1]: N = 1,2,3 x cy (3D: depth taps concatenated along N)
  int ymirror;               // dY ring slots 0,1 are mirrored at physical slots ystages, ystages+1 (no window wrap)
  int ntw, nth;
  long long nunits;          // (n, depth chunk, tile row, tile column)
  int stages, ystages, pf;   // X ring, dY ring (power of two), cp.async groups in flight
  int xplane_bytes, yplane_bytes;
  int tpm, nM;               // taps per MMA (128 / cx), MMAs per depth tap
  uint32_t idesc;
  uint64_t adesc_hi, bdesc_hi;
  int tmem_cols;
  int ntaps;
  // strides in bytes: next tile row / next plane / next image of X and of dY (images-as-rows plan: the
  // "row" is the next IMAGE and the "plane" the next image row)
  int xrow_b, yrow_b;
  long long xplane_b, ximg_b, yplane_b2, yimg_b;
  // TMA staging (images-as-rows plan): ONE box of hr tile rows x xb pixels per X plane, the kw taps are
  // views shifted by whole pixels (LBO = one pixel); one box of 16 x 8*ks positions per dY plane
  int tma, xb;
};

struct UnitIter {
  int nn, dc, ht, wt;
  __device__ __forceinline__ void init(const WHParams& p, long long u) {
    wt = (int)(u % p.ntw); u /= p.ntw;
    ht = (int)(u % p.nth); u /= p.nth;
    dc = (int)(u % p.ndc);
    nn = (int)(u / p.ndc);
  }
  __device__ __forceinline__ void next(const WHParams& p) {
    if (++wt == p.ntw) { wt = 0; if (++ht == p.nth) { ht = 0; if (++dc == p.ndc) { dc = 0; ++nn; } } }
  }
};

template <bool S2>    // S2: X sampled with stride 2 (compile-time, so that the stride-1 kernel keeps its tight loops)
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_halo_kernel(const __grid_constant__ WHParams p, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_y, const void* __restrict__ x, const void* __restrict__ dy,
                  float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_x = smem;
  const int x_bytes = ((p.stages * p.xplane_bytes + 1023) / 1024) * 1024;
  uint8_t* smem_y = smem + x_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_y + (size_t)(p.ystages + p.ymirror) * p.yplane_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWMaxStages;
  uint64_t* yempty = bars + 2 * kWMaxStages;
  uint64_t* tfull = yempty + kYMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  uint32_t* touched_slot = tmem_slot + 1;

  constexpr int PS = S2 ? 2 : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u_begin = p.nunits * (long long)blockIdx.x / gridDim.x;
  const long long u_end = p.nunits * (long long)(blockIdx.x + 1) / gridDim.x;
  const int my_units = (int)(u_end - u_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { h_mbar_init(&full[s], p.tma ? 1 : kWProdWarps); h_mbar_init(&empty[s], 1); }
    for (int s = 0; s < p.ystages; ++s) h_mbar_init(&yempty[s], 1);
    h_mbar_init(tfull, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  h_fence_before();
  __syncthreads();
  h_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWProdWarps && p.tma) {
    // ===================== TMA producer (stride 1, images-as-rows): one elected thread =====================
    // Out-of-range coordinates are filled with zeros by the hardware: the 'same' padding of X and the
    // positions of a ragged tile that must not contribute (dY = 0).
    if (warp == 0) {
      UnitIter ui; ui.init(p, u_begin);
      Ring st, yr; st.init(p.stages); yr.init(p.ystages);
      for (int k = 0; k < my_units; ++k, ui.next(p)) {
        const int d0 = ui.dc * p.DC, d1 = min(d0 + p.DC, p.D);
        const int y0 = ui.ht * kWTH, x0 = ui.wt * kWTW * p.ks;
        for (int z = d0 + p.od0; z <= d1 - 1 + p.od0 + p.nkd - 1; ++z) {
          const bool xvalid = z >= 0 && z < p.XD;
          const int yp = z - p.od0;
          const bool yvalid = yp >= d0 && yp < d1;
          h_mbar_wait_warp(&empty[st.idx], st.phase ^ 1);
          if (yvalid) h_mbar_wait_warp(&yempty[yr.idx], yr.phase ^ 1);
          const int ncopy = (p.ymirror && yr.idx < p.ymirror) ? 2 : 1;
          if (h_elect_one()) {
            const uint32_t bar = h_smem_u32(&full[st.idx]);
            h_mbar_expect_tx(&full[st.idx], (xvalid ? (uint32_t)p.xplane_bytes : 0u) +
                                                (yvalid ? (uint32_t)(p.yplane_bytes * ncopy) : 0u));
            if (xvalid)
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                  " [%0], [%1, {%3, %4, %5, %6}], [%2];"
                  ::"r"(h_smem_u32(smem_x + (size_t)st.idx * p.xplane_bytes)), "l"(reinterpret_cast<uint64_t>(&map_x)), "r"(bar),
                  "r"(0), "r"(x0 + p.ow0), "r"(z), "r"(y0 + p.oh0)
                  : "memory");
            if (yvalid)
              for (int cpy = 0; cpy < ncopy; ++cpy)
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
                    " [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(h_smem_u32(smem_y + (size_t)(yr.idx + cpy * p.ystages) * p.yplane_bytes)),
                    "l"(reinterpret_cast<uint64_t>(&map_y)), "r"(bar), "r"(0), "r"(x0), "r"(yp), "r"(y0)
                    : "memory");
          }
          __syncwarp();
          if (yvalid) yr.next();
          st.next();
        }
      }
    }
  } else if (warp < kWProdWarps) {
    // ===================== producers =====================
    // Register-resident chunk tables (see gconv_halo.cu): byte offset from the tile's halo origin,
    // swizzled slot in the stage, halo row / column for the border test; surplus slots replicate
    // slot 0; interior tiles skip the border test.
    const int t = threadIdx.x;
    const int pitch = p.cx * 2, cpr = pitch / 16;
    const int total = p.hr * p.rw * cpr;
    const int cpr_log2 = cpr == 8 ? 3 : (cpr == 4 ? 2 : 1);
    const int sw_shift = cpr == 8 ? 0 : (cpr == 4 ? 1 : 2);
    const int sw_mask = cpr == 8 ? 7 : (cpr == 4 ? 3 : 1);
    const int run = 8 * p.ks;
    const char* x_base = reinterpret_cast<const char*>(x);
    const char* y_base = reinterpret_cast<const char*>(dy);
    const int xpix = (int)(p.ldx * 2), xrow = p.xrow_b;
    const int ypix = (int)(p.ldy * 2), yrow = p.yrow_b;
    const int H = p.H, W = p.W, XH = p.XH, XW = p.XW;
    int soff[kWNCH], hdv[kWNCH];          // hdv = halo row | X column << 8 (border test of edge tiles)
    uint32_t doff[kWNCH];
#pragma unroll
    for (int j = 0; j < kWNCH; ++j) {
      int idx = t + j * kWProdThreads;
      if (idx >= total) idx = t;
      const int row = idx >> cpr_log2, c = idx & (cpr - 1);
      const int hy = row / p.rw;
      const int rem = row - hy * p.rw;
      const int kw = rem / run;
      const int dx = kw + (rem - kw * run) * PS;              // X column relative to the tile origin
      const int sw = (row >> sw_shift) & sw_mask;
      doff[j] = (uint32_t)(row * (pitch >> 4) + (c ^ sw)) << 4;
      soff[j] = hy * xrow + dx * xpix + c * 16;
      hdv[j] = hy | (dx << 8);
    }
    // dY tile: rows ordered (sub-tile, h, w8), cy*2 bytes each
    const int ypitch = p.cy * 2, ycpr = ypitch / 16;
    const int ycpr_log2 = ycpr == 8 ? 3 : (ycpr == 4 ? 2 : 1);
    const int ysw_shift = ycpr == 8 ? 0 : (ycpr == 4 ? 1 : 2);
    const int ysw_mask = ycpr == 8 ? 7 : (ycpr == 4 ? 3 : 1);
    const int ytotal = p.ks * 128 * ycpr;
    int ysoff[kWYCH], yhx[kWYCH];
    uint32_t ydoff[kWYCH];
#pragma unroll
    for (int j = 0; j < kWYCH; ++j) {
      int idx = t + j * kWProdThreads;
      if (idx >= ytotal) idx = t;
      const int row = idx >> ycpr_log2, c = idx & (ycpr - 1);
      const int sj = row >> 7, h = (row >> 3) & 15, w = row & 7;
      const int sw = (row >> ysw_shift) & ysw_mask;
      ydoff[j] = (uint32_t)(row * ycpr + (c ^ sw)) << 4;
      yhx[j] = h | ((sj * 8 + w) << 8);
      ysoff[j] = h * yrow + (sj * 8 + w) * ypix + c * 16;
    }
    UnitIter ui; ui.init(p, u_begin);
    Ring st, ar, yr; st.init(p.stages); ar.init(p.stages); yr.init(p.ystages);
    uint32_t issued = 0;
    const int pf = p.pf;
    for (int k = 0; k < my_units; ++k, ui.next(p)) {
      const int d0 = ui.dc * p.DC, d1 = min(d0 + p.DC, p.D);
      const int y0 = ui.ht * kWTH, x0 = ui.wt * kWTW * p.ks;
      const int xyb = y0 * PS + p.oh0, xxb = x0 * PS + p.ow0;          // X origin of the staged halo
      const bool interior = xyb >= 0 && xyb + p.hr <= XH && xxb >= 0 && xxb + (run - 1) * PS + p.nkw <= XW &&
                            y0 + kWTH <= H && x0 + run <= W;
      const long long xorg = (long long)xyb * xrow + (long long)xxb * xpix;
      const long long yorg = (long long)y0 * yrow + (long long)x0 * ypix;
      for (int z = d0 * PS + p.od0; z <= (d1 - 1) * PS + p.od0 + p.nkd - 1; ++z) {
        const bool xvalid = z >= 0 && z < p.XD;
        // the dY plane first needed at this step (depth tap 0 of X plane z)
        const int tnew = z - p.od0;
        const int yp = S2 ? tnew >> 1 : tnew;
        const bool yvalid = !(S2 && (tnew & 1)) && yp >= d0 && yp < d1;
        h_mbar_wait_warp(&empty[st.idx], st.phase ^ 1);
        if (yvalid) h_mbar_wait_warp(&yempty[yr.idx], yr.phase ^ 1);
        if (xvalid) {
          const uint32_t dst_base = h_smem_u32(smem_x + (size_t)st.idx * p.xplane_bytes);
          const char* origin = x_base + (long long)ui.nn * p.ximg_b + (long long)z * p.xplane_b + xorg;
          if (interior) {
#pragma unroll
            for (int j = 0; j < kWNCH; ++j) cp_async16_zfill(dst_base + doff[j], origin + soff[j], 16u);
          } else {
#pragma unroll
            for (int j = 0; j < kWNCH; ++j) {
              const bool ok = (unsigned)(xyb + (hdv[j] & 255)) < (unsigned)XH && (unsigned)(xxb + (hdv[j] >> 8)) < (unsigned)XW;
              cp_async16_zfill(dst_base + doff[j], ok ? origin + soff[j] : x_base, ok ? 16u : 0u);
            }
          }
        }
        if (yvalid) {
          const uint32_t dst_base = h_smem_u32(smem_y + (size_t)yr.idx * p.yplane_bytes);
          const char* origin = y_base + (long long)ui.nn * p.yimg_b + (long long)yp * p.yplane_b2 + yorg;
          // ring slots 0 and 1 are also written to their mirrors behind the ring, so that the
          // three-plane window of the MMA warp never wraps
          const int ncopy = (p.ymirror && yr.idx < p.ymirror) ? 2 : 1;
          for (int cpy = 0; cpy < ncopy; ++cpy) {
            const uint32_t dstb = dst_base + (uint32_t)(cpy * p.ystages * p.yplane_bytes);
            if (interior) {
#pragma unroll
              for (int j = 0; j < kWYCH; ++j) cp_async16_zfill(dstb + ydoff[j], origin + ysoff[j], 16u);
            } else {
#pragma unroll
              for (int j = 0; j < kWYCH; ++j) {
                const bool ok = y0 + (yhx[j] & 255) < H && x0 + (yhx[j] >> 8) < W;
                cp_async16_zfill(dstb + ydoff[j], ok ? origin + ysoff[j] : y_base, ok ? 16u : 0u);
              }
            }
          }
          yr.next();
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (issued >= (uint32_t)pf) {
          if (pf >= 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
          else if (pf == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
          else if (pf == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
          else if (pf == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> UMMA (async proxy)
          __syncwarp();
          if (lane == 0) h_mbar_arrive(&full[ar.idx]);
          ar.next();
        }
        ++issued;
        st.next();
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (uint32_t j = (issued > (uint32_t)pf ? issued - pf : 0); j < issued; ++j) { h_mbar_arrive(&full[ar.idx]); ar.next(); }
  } else if (warp == kWMmaWarp) {
    // ===================== MMA issue (warp-uniform loop, one elected lane issues) =====================
    const uint32_t a_hi = (uint32_t)(p.adesc_hi >> 32), b_hi = (uint32_t)(p.bdesc_hi >> 32);
    const uint32_t a_lo_base = (uint32_t)p.adesc_hi, b_lo_base = (uint32_t)p.bdesc_hi;
    const uint32_t pitch = (uint32_t)p.cx * 2;
    // cp.async layout: taps are 8*KS rows apart; 16 positions = 2 output rows = 2*s halo rows; the next sub-tile
    // is 8 pixels further in the run.  TMA layout (rows [h][xb pixels], positions ordered (h, w)): taps are one
    // pixel apart; the K step kk of a 128-position block covers pixels 16*(kk & hm) .. +15 of row kk >> hs.
    const uint32_t a_mstep = ((uint32_t)(p.tma ? p.tpm : p.tpm * 8 * p.ks) * pitch) >> 4;   // next M tile: tpm taps further
    const uint32_t a_row = ((uint32_t)(p.tma ? p.xb : 2 * PS * p.rw) * pitch) >> 4;
    const uint32_t a_h16 = (16u * pitch) >> 4;
    const int hs = (p.tma && p.ks == 4) ? 1 : 0;
    const uint32_t hm = (p.tma && p.ks == 4) ? 1u : 0u;
    const uint32_t a_sub = p.tma ? (uint32_t)(128 / (8 * p.ks)) * a_row : (8u * pitch) >> 4;
#define WH_AOFF(kk) ((uint32_t)((kk) >> hs) * a_row + ((uint32_t)(kk) & hm) * a_h16)
    const uint32_t b_kstep = (16u * (uint32_t)p.cy * 2u) >> 4;
    const uint32_t b_sub = (128u * (uint32_t)p.cy * 2u) >> 4;
    const uint32_t y_lo0 = h_smem_u32(smem_y) >> 4, y_stage = (uint32_t)p.yplane_bytes >> 4;
    UnitIter ui; ui.init(p, u_begin);
    Ring st; st.init(p.stages);
    uint32_t yseq0 = 0, touched = 0;
    const uint32_t ymask = (uint32_t)p.ystages - 1u;
    for (int k = 0; k < my_units; ++k, ui.next(p)) {
      const int d0 = ui.dc * p.DC, d1 = min(d0 + p.DC, p.D);
      // sub-tiles inside the image (TMA layout: 128-position blocks of whole rows, ragged columns hold dY = 0)
      const int nsub = p.tma ? p.ks : min(p.ks, (p.W - ui.wt * kWTW * p.ks + kWTW - 1) / kWTW);
      for (int z = d0 * PS + p.od0; z <= (d1 - 1) * PS + p.od0 + p.nkd - 1; ++z) {
        const bool xvalid = z >= 0 && z < p.XD;
        h_mbar_wait(&full[st.idx], st.phase);
        h_fence_after();                   // (proxy fence: issued by the producers before they arrive)
        if (xvalid) {
          const uint32_t a_lo0 = a_lo_base | (h_smem_u32(smem_x + (size_t)st.idx * p.xplane_bytes) >> 4);
          // dY planes paired with this X plane: z-1 .. z+1 (depth taps kd = 2, 1, 0), clipped to the
          // chunk; consecutive planes sit in consecutive ring slots, so the whole window is ONE B
          // operand of N = (#planes)*cy (LBO = slot stride) and the three depth taps share one
          // read of the X views.  Accumulator columns of M tile mi: [(mi*3 + g)*cy], g = plane-(z-1).
          // window of dY planes: (z - od0 - kdi) / s for the depth taps kdi = nkd-1 .. 0
          int pa, pb, wlo;
          if (!S2) {
            wlo = z - p.od0 - (p.nkd - 1);
            pa = max(d0, wlo); pb = min(d1 - 1, z - p.od0);
          } else {
            // stride 2: taps of one parity only; handled one plane at a time (window of length 1 per tap)
            wlo = 0; pa = 1; pb = 0;
            for (int kdi = 0; kdi < p.nkd; ++kdi) {
              const int tt = z - p.od0 - kdi;
              if (tt < 0 || (tt & 1)) continue;
              const int pl = tt >> 1;
              if (pl < d0 || pl >= d1) continue;
              const uint32_t slot = (yseq0 + (uint32_t)(pl - d0)) & ymask;
              const uint32_t b_lo0 = b_lo_base | (y_lo0 + slot * y_stage);
              const int gs = p.nkd - 1 - kdi;
#pragma unroll 1
              for (int mi = 0; mi < p.nM; ++mi) {
                const uint32_t bit1 = 1u << (mi * p.nkd + gs);
                const uint32_t af = (touched & bit1) ? 1u : 0u;
                touched |= bit1;
                const uint32_t d_tmem = tmem_base + (uint32_t)((mi * p.nkd + gs) * p.cy);
                const uint32_t a_lo = a_lo0 + (uint32_t)mi * a_mstep;
                const int last = mi == p.nM - 1 ? 1 : 0;
                if (h_elect_one()) {
#pragma unroll 1
                  for (int sj = 0; sj < nsub; ++sj)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                      h_umma2(d_tmem, a_lo + (uint32_t)sj * a_sub + WH_AOFF(kk), a_hi,
                              b_lo0 + (uint32_t)sj * b_sub + (uint32_t)kk * b_kstep, b_hi, p.idesc_n[last][0],
                              (sj | kk) == 0 ? af : 1u);
                }
                __syncwarp();
              }
            }
          }
          if (pa <= pb) {
            const int g0 = pa - wlo;
            const uint32_t s0 = (yseq0 + (uint32_t)(pa - d0)) & ymask;            // ring slot of plane pa
            int nseg = 1, len0 = pb - pa + 1;
            if (!p.ymirror && (int)s0 + len0 > p.ystages) { len0 = p.ystages - (int)s0; nseg = 2; }
#pragma unroll 1
            for (int sg = 0; sg < nseg; ++sg) {
              const int gs = sg == 0 ? g0 : g0 + len0;
              const int len = sg == 0 ? len0 : (pb - pa + 1) - len0;
              const uint32_t slot = sg == 0 ? s0 : 0u;
              const uint32_t b_lo0 = b_lo_base | (y_lo0 + slot * y_stage);
#pragma unroll 1
              for (int mi = 0; mi < p.nM; ++mi) {
                const uint32_t d_tmem = tmem_base + (uint32_t)((mi * p.nkd + gs) * p.cy);
                // first write of these accumulator columns?
                uint32_t bits = ((1u << len) - 1u) << (mi * p.nkd + gs);
                const uint32_t accf = (touched & bits) == bits ? 1u : 0u;
                const bool mixed = (touched & bits) != 0u && !accf;
                const uint32_t a_lo = a_lo0 + (uint32_t)mi * a_mstep;
                const int last = mi == p.nM - 1 ? 1 : 0;
                if (mixed) {
                  // some columns new, some not (chunk edges): one MMA per plane
                  for (int gi = 0; gi < len; ++gi) {
                    const uint32_t bit1 = 1u << (mi * p.nkd + gs + gi);
                    const uint32_t af = (touched & bit1) ? 1u : 0u;
                    if (h_elect_one()) {
#pragma unroll 1
                      for (int sj = 0; sj < nsub; ++sj)
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                          h_umma2(d_tmem + (uint32_t)(gi * p.cy), a_lo + (uint32_t)sj * a_sub + WH_AOFF(kk), a_hi,
                                  b_lo0 + (uint32_t)gi * y_stage + (uint32_t)sj * b_sub + (uint32_t)kk * b_kstep, b_hi,
                                  p.idesc_n[last][0], (sj | kk) == 0 ? af : 1u);
                    }
                    __syncwarp();
                  }
                } else {
                  if (h_elect_one()) {
#pragma unroll 1
                    for (int sj = 0; sj < nsub; ++sj)
#pragma unroll
                      for (int kk = 0; kk < 8; ++kk)
                        h_umma2(d_tmem, a_lo + (uint32_t)sj * a_sub + WH_AOFF(kk), a_hi,
                                b_lo0 + (uint32_t)sj * b_sub + (uint32_t)kk * b_kstep, b_hi, p.idesc_n[last][len - 1],
                                (sj | kk) == 0 ? accf : 1u);
                  }
                  __syncwarp();
                }
                touched |= bits;
              }
            }
          }
        }
        // last use of a dY plane is its depth tap nkd-1: plane (z - od0 - (nkd-1)) / s
        const int trel = z - p.od0 - (p.nkd - 1);
        const int rel = S2 ? trel >> 1 : trel;
        const bool relv = trel >= 0 && !(S2 && (trel & 1)) && rel >= d0 && rel < d1;
        if (h_elect_one()) {
          h_commit(&empty[st.idx]);
          if (relv) h_commit(&yempty[(yseq0 + (uint32_t)(rel - d0)) & ymask]);
        }
        __syncwarp();
        st.next();
      }
      yseq0 += (uint32_t)(d1 - d0);
    }
    if (lane == 0) {
      *touched_slot = touched;
      h_mbar_arrive(tfull);
    }
    __syncwarp();
    if (h_elect_one()) h_commit(tfull);
    __syncwarp();
  } else {
    // ===================== final epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    h_mbar_wait_warp(tfull, 0);
    h_fence_after();
    const uint32_t touched = *reinterpret_cast<volatile uint32_t*>(touched_slot);
    const int tl = r / p.cx, ci = r - tl * p.cx;
    for (int oi = 0; oi < p.nkd; ++oi)
      for (int mi = 0; mi < p.nM; ++mi) {
        const int ai = mi * p.nkd + (p.nkd - 1 - oi);          // column block g = plane - (z-1) = 2 - oi
        const int tap_local = mi * p.tpm + tl;
        const bool valid = tap_local < p.ntp;
        const bool have = (touched >> ai) & 1u;
        float* dst = partial + (((long long)blockIdx.x * p.ntaps + (oi * p.ntp + tap_local)) * p.cx + ci) * p.cy;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ai * p.cy);
        for (int c0 = 0; c0 < p.cy; c0 += 16) {
          float v[16];
          h_tmem_ld16(taddr + c0, v);
          if (!valid) continue;
          if (!have) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
  }

  h_fence_before();
  __syncthreads();
  if (warp == kWMmaWarp) {
    h_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// fixed-order sum of the per-CTA partials
__global__ void wgrad_halo_reduce_kernel(const float* __restrict__ partial, int nw, int nsplit,
                                         float* __restrict__ dw, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 4 <= nsplit; k += 4) {
    s0 += partial[(long long)(k + 0) * nw + i];
    s1 += partial[(long long)(k + 1) * nw + i];
    s2 += partial[(long long)(k + 2) * nw + i];
    s3 += partial[(long long)(k + 3) * nw + i];
  }
  for (; k < nsplit; ++k) s0 += partial[(long long)k * nw + i];
  const float s = (s0 + s1) + (s2 + s3);
  dw[i] = accumulate ? dw[i] + s : s;
}

bool wh_plan(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy, WHParams* pp, bool rows = false) {
  WHParams& p = *pp;
  memset(&p, 0, sizeof(p));
  if (getenv("URN_NO_HALO_WGRAD")) return false;
  if ((x_dtype != URN_BF16 && x_dtype != URN_F16) || x_dtype != y_dtype) return false;
  if (d.s != 1 && d.s != 2) return false;
  if (rows) {       // images-as-rows relabelling of a 2D 3x3 layer: 1x3 in-plane taps, 3 "depth" taps (wh_plan_rows)
    if (!(d.kd == 3 && d.kh == 1 && d.kw == 3 && d.s == 1)) return false;
  } else if (!((d.kh == 3 && d.kw == 3) || (d.kh == 1 && d.kw == 1)) || (d.kd != 1 && d.kd != d.kh)) return false;
  if (getenv("URN_WHALO_ONLY3X3S1") && !(d.s == 1 && d.kh == 3)) return false;
  if (d.kh == 1 && d.s == 2) return false;      // only every other staged row would be used: the tap-box kernel is faster
  if (d.yd != (d.xd + d.s - 1) / d.s || d.yh != (d.xh + d.s - 1) / d.s || d.yw != (d.xw + d.s - 1) / d.s) return false;
  if (d.cx != 16 && d.cx != 32 && d.cx != 64) return false;
  if (d.cy != 16 && d.cy != 32 && d.cy != 64) return false;
  if (ldx % 8 || ldy % 8) return false;
  p.n = d.n; p.D = d.yd; p.H = d.yh; p.W = d.yw; p.XD = d.xd; p.XH = d.xh; p.XW = d.xw;
  p.s = d.s; p.cx = d.cx; p.cy = d.cy; p.ldx = ldx; p.ldy = ldy;
  if ((long long)p.XH * p.XW * ldx * 2 >= (1ll << 31) || (long long)p.H * p.W * ldy * 2 >= (1ll << 31)) return false;
  p.nkd = d.kd; p.nkh = d.kh; p.nkw = d.kw; p.ntp = d.kh * d.kw;
  p.od0 = -d.pad_d; p.oh0 = -d.pad_h; p.ow0 = -d.pad_w;
  p.LA = 0;
  p.ntaps = p.ntp * p.nkd;
  p.hr = (kWTH - 1) * d.s + d.kh;
  p.ndc = (p.D + 47) / 48;
  p.DC = (p.D + p.ndc - 1) / p.ndc;
  p.ndc = (p.D + p.DC - 1) / p.DC;
  p.xrow_b = (int)(ldx * 2 * p.XW); p.xplane_b = (long long)p.xrow_b * p.XH; p.ximg_b = p.xplane_b * p.XD;
  p.yrow_b = (int)(ldy * 2 * p.W); p.yplane_b2 = (long long)p.yrow_b * p.H; p.yimg_b = p.yplane_b2 * p.D;
  p.tpm = 128 / p.cx;
  p.nM = (p.ntp + p.tpm - 1) / p.tpm;
  const int cols = p.nkd * p.nM * p.cy;
  if (cols > 512) return false;
  p.tmem_cols = std::max(32, pow2c(cols));
  // super-tile: KS sub-tiles share one staged X plane (<= 56 KB) and one barrier round trip
  const int cpr = p.cx * 2 / 16;
  // TMA staging for the images-as-rows plan (URN_WHALO_TMA=0 keeps the cp.async producers): the X plane is
  // staged ONCE (no kw-shifted copies: 20 KB instead of 48 KB for cx = 16), which leaves room for a deeper ring
  bool want_tma = rows && !(getenv("URN_WHALO_TMA") && atoi(getenv("URN_WHALO_TMA")) == 0) && get_encode_h() != nullptr;
  for (int attempt = 0; attempt < 2; ++attempt) {
    p.ks = 4;
    if (const char* e = getenv("URN_WHALO_KS")) p.ks = std::max(1, std::min(p.ks, atoi(e)));
    while (p.ks > 1 && (p.W + kWTW - 1) / kWTW <= p.ks / 2) p.ks /= 2;
    while (p.ks > 1 && p.ks * 128 * (p.cy * 2 / 16) > kWYCH * kWProdThreads) p.ks /= 2;
    if (!want_tma)
      while (p.ks > 1 && p.hr * p.nkw * 8 * p.ks * cpr > kWNCH * kWProdThreads) p.ks /= 2;
    if (want_tma && p.ks < 2) { want_tma = false; continue; }
    break;
  }
  p.tma = want_tma ? 1 : 0;
  p.rw = p.nkw * 8 * p.ks;
  p.xb = 8 * p.ks + 8;                       // pixels per staged row (TMA): the run, its halo, padded to whole swizzle atoms
  p.ntw = (p.W + kWTW * p.ks - 1) / (kWTW * p.ks); p.nth = (p.H + kWTH - 1) / kWTH;
  if (!rows && p.D >= 16 && !getenv("URN_WHALO_NO_BALANCE")) {
    // small grids: fewer units than two waves of CTAs -- cut the depth into more chunks when that shortens
    // the busiest CTA's walk (each chunk re-stages nkd - 1 halo planes)
    const long long per_chunk = (long long)p.n * p.nth * p.ntw;
    const int ctas = persistent_ctas();
    if (per_chunk * p.ndc < 2ll * ctas) {
      int best = p.ndc; long long best_cost = -1;
      for (int ndc = p.ndc; ndc <= std::max(p.ndc, p.D / 8); ++ndc) {
        const int dc = (p.D + ndc - 1) / ndc;
        const long long units = per_chunk * ((p.D + dc - 1) / dc);
        const long long cost = ((units + ctas - 1) / ctas) * (dc + p.nkd - 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ndc; }
      }
      p.DC = (p.D + best - 1) / best;
      p.ndc = (p.D + p.DC - 1) / p.DC;
    }
  }
  p.nunits = (long long)p.n * p.ndc * p.nth * p.ntw;
  if (p.nunits >= (1ll << 31)) return false;
  p.xplane_bytes = p.tma ? p.hr * p.xb * p.cx * 2 : p.hr * p.rw * p.cx * 2;
  p.yplane_bytes = p.ks * 128 * p.cy * 2;
  if (!p.tma) {
    if (p.hr * p.rw * cpr > kWNCH * kWProdThreads) return false;
    // the garbage taps of the last M tile read up to (nM*tpm - ntp) tap strides past the plane: keep
    // that inside the dY region that follows the X stages
    if ((p.nM * p.tpm - p.ntp) * 8 * p.ks * p.cx * 2 > 2 * p.yplane_bytes + 8192) return false;
  }
  // Producers run `pf` cp.async groups ahead of the full-barrier arrival; the dY slot of plane z+1
  // frees two steps after its previous tenant's last use, so the dY ring needs >= pf + 3 slots
  // (and the X ring >= pf + 2, see below) or producers and MMA warp would wait on each other.
  const int budget = 222 * 1024;
  p.pf = -1;
  p.ymirror = (p.nkd == 3 && p.s == 1) ? 2 : 0;
  if (p.tma) {
    // no cp.async groups: the look-ahead is the ring depth.  dY ring of 8 (3 planes live + 5 ahead) when at
    // least 3 X stages still fit, else 4.
    for (int ys : {8, 4}) {
      if (const char* e = getenv("URN_WHALO_YS")) { if (atoi(e) != ys) continue; }
      const int st = std::min(kWMaxStages, (budget - (ys + p.ymirror) * p.yplane_bytes) / p.xplane_bytes);
      if (st >= (ys == 8 ? 3 : 2)) { p.pf = 1; p.ystages = ys; p.stages = st; break; }
    }
  } else {
    // The arrival of plane i is posted after the loads of plane i+pf have been issued, which needs the slot
    // of plane i+pf-stages: with stages == pf+1 that is plane i-1, i.e. every MMA batch would wait for its
    // predecessor to finish PLUS the issue time of a plane (measured: 3D 16->16 5130 cycles per plane against
    // 3712 of MMA issue).  Require stages >= pf+2; a 2-stage ring runs with pf = 0 (wait for the plane's own
    // loads, arrive, then wait for the next free slot).
    p.pf = -1;
    int pf_max = 4;
    if (const char* e = getenv("URN_WHALO_PF")) pf_max = std::max(0, std::min(4, atoi(e)));
    for (int pf = pf_max; pf >= 0; --pf) {
      const int ys = pow2c(pf + (p.nkd == 3 ? 3 : 1));
      const int st = std::min(kWMaxStages, (budget - (ys + p.ymirror) * p.yplane_bytes) / p.xplane_bytes);
      if (st >= pf + 2) { p.pf = pf; p.ystages = ys; p.stages = st; break; }
    }
  }
  if (p.pf < 0) return false;
  const int fmt = x_dtype == URN_BF16 ? 1 : 0;
  const uint32_t idesc0 = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16) |
                          ((uint32_t)(p.cy >> 3) << 17);
  p.idesc = idesc0 | ((uint32_t)(128 >> 4) << 24);
  // cx = 16: the last M tile holds one live tap (accumulator rows 0..15 = TMEM lanes 0..15 in the
  // M = 64 layout too): a half-height MMA reads half the shared memory.
  const bool m64 = p.cx == 16 && (p.ntp - (p.nM - 1) * p.tpm) == 1 && !getenv("URN_WHALO_NO_M64");
  p.idesc_last = m64 ? (idesc0 | ((uint32_t)(64 >> 4) << 24)) : p.idesc;
  {
    const uint32_t idm = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16);
    const uint32_t m_last = m64 ? 64u : 128u;
    for (int nb = 0; nb < 3; ++nb) {
      const uint32_t nfield = (uint32_t)(((nb + 1) * p.cy) >> 3) << 17;
      p.idesc_n[0][nb] = idm | nfield | ((128u >> 4) << 24);
      p.idesc_n[1][nb] = idm | nfield | ((m_last >> 4) << 24);
    }
  }
  auto layout_of = [](uint64_t bytes) -> uint64_t { return bytes == 128 ? 2 : (bytes == 64 ? 4 : 6); };
  const uint64_t xp = (uint64_t)p.cx * 2, yp = (uint64_t)p.cy * 2;
  // MN-major: SBO = next 8 positions (K), LBO = next channel group (MN)
  p.adesc_hi = (layout_of(xp) << 61) | (1ull << 46) | ((((uint64_t)(p.s * p.rw) * xp) >> 4) << 32) | ((((uint64_t)(8 * p.ks) * xp) >> 4) << 16);
  // TMA layout: the next 8 positions are the next 8 pixels of the row, the next tap (M group) is ONE pixel further
  if (p.tma) p.adesc_hi = (layout_of(xp) << 61) | (1ull << 46) | (((8 * xp) >> 4) << 32) | ((xp >> 4) << 16);
  // B: N groups of cy channels = consecutive dY ring slots (LBO = slot stride)
  p.bdesc_hi = (layout_of(yp) << 61) | (1ull << 46) | (((8 * yp) >> 4) << 32) | ((((uint64_t)p.yplane_bytes) >> 4) << 16);
  if (((uint64_t)p.yplane_bytes >> 4) >= (1u << 14)) return false;
  return true;
}

// Images-as-rows plan for the 2D 3x3 stride-1 layers of a batch of >= 12 images (default on,
// URN_WHALO_ROWS=0 disables it; same idea as halo_plan_rows in gconv_halo.cu): planes := image rows,
// tile rows := the IMAGES of the batch, columns := image columns.  The kh taps become the depth taps,
// whose three dY "planes" (image rows y-1, y, y+1) are concatenated along N: per 16 positions ONE MMA
// with M = (kw, cx), N = 3*cy instead of two with M = (8 taps | 1 tap, cx), N = cy -- every
// tcgen05.mma occupies the tensor pipe >= ~58 cycles whatever its shape (tools/probe/mmarate.cu).
// The chunking of the "depth" (image rows) is chosen so that the units spread evenly over the CTAs.
bool wh_plan_rows(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy, WHParams* pp) {
  if (const char* e = getenv("URN_WHALO_ROWS")) { if (atoi(e) == 0) return false; }
  if (d.n < 12 || d.xd != 1 || d.yd != 1 || d.kd != 1 || d.kh != 3 || d.kw != 3 || d.s != 1) return false;
  if (d.xh != d.yh || d.xw != d.yw || d.xh < 4) return false;
  const long long ximg = ldx * 2 * (long long)d.xw * d.xh, yimg = ldy * 2 * (long long)d.yw * d.yh;
  if (ximg * 17 >= (1ll << 31) || yimg * 17 >= (1ll << 31)) return false;      // 32-bit offsets across the images of a tile
  urn_conv_desc v = d;
  v.n = 1;
  v.xd = d.xh; v.xh = d.n; v.xw = d.xw;
  v.yd = d.yh; v.yh = d.n; v.yw = d.yw;
  v.kd = 3; v.kh = 1; v.kw = 3;
  v.pad_d = d.pad_h; v.pad_h = 0; v.pad_w = d.pad_w;
  if (!wh_plan(v, x_dtype, y_dtype, ldx, ldy, pp, true)) return false;
  WHParams& p = *pp;
  // depth chunks: few enough to keep the 2-plane halo per chunk small, balanced over the CTAs
  {
    const long long per_chunk = (long long)p.nth * p.ntw;
    const int ctas = persistent_ctas();
    int best = p.ndc; double best_cost = 1e30;
    for (int ndc = std::max(1, (p.D + 95) / 96); ndc <= std::max(1, p.D / 8); ++ndc) {
      const int dc = (p.D + ndc - 1) / ndc;
      const long long units = per_chunk * ((p.D + dc - 1) / dc);
      const double cost = (double)((units + ctas - 1) / ctas) * (dc + 2);      // planes staged by the busiest CTA
      if (cost < best_cost - 1e-9) { best_cost = cost; best = ndc; }
    }
    p.DC = (p.D + best - 1) / best;
    p.ndc = (p.D + p.DC - 1) / p.DC;
    p.nunits = (long long)p.n * p.ndc * p.nth * p.ntw;
  }
  p.xplane_b = ldx * 2 * d.xw;              // next image row
  p.xrow_b = (int)ximg;                     // next image
  p.ximg_b = 0;
  p.yplane_b2 = ldy * 2 * d.yw;
  p.yrow_b = (int)yimg;
  p.yimg_b = 0;
  return true;
}

// 3D stride-1 3x3x3 layers whose 27 x cx x cy fp32 accumulators exceed the TMEM (cx = cy = 64): ONE PASS PER DEPTH
// TAP -- each pass is the nkd = 1 plan with the X planes shifted by the tap (pad_d - kd), so X and dY are streamed
// three times out of L2 instead of 27 times through the tap-shifted TMA boxes of wgrad_tc_kernel
// (URN_WHALO_SPLIT=0 disables it).
static bool wh_split_desc(const urn_conv_desc& d, int kd_sel, urn_conv_desc* o) {
  if (d.kd != 3 || d.kh != 3 || d.kw != 3 || d.s != 1 || d.xd < 2) return false;
  if (const char* e = getenv("URN_WHALO_SPLIT")) { if (atoi(e) == 0) return false; }
  *o = d;
  o->kd = 1;
  o->pad_d = d.pad_d - kd_sel;
  return true;
}

}  // namespace

bool wgrad_halo_supported(const urn_conv_desc& d, int x_dtype, int y_dtype, int64_t ldx, int64_t ldy) {
  WHParams p;
  if (wh_plan_rows(d, x_dtype, y_dtype, ldx, ldy, &p) || wh_plan(d, x_dtype, y_dtype, ldx, ldy, &p)) return true;
  urn_conv_desc sd;
  return wh_split_desc(d, 0, &sd) && wh_plan(sd, x_dtype, y_dtype, ldx, ldy, &p);
}

int64_t wgrad_halo_workspace(const urn_conv_desc& d) {
  WHParams p;
  if (wh_plan(d, URN_BF16, URN_BF16, 64, 64, &p)) return (int64_t)num_sms() * p.ntaps * d.cx * d.cy * 4;
  urn_conv_desc sd;
  if (wh_split_desc(d, 0, &sd) && wh_plan(sd, URN_BF16, URN_BF16, 64, 64, &p)) return (int64_t)num_sms() * p.ntaps * d.cx * d.cy * 4;
  return 0;
}

namespace {
int wh_launch_planned(const urn_conv_desc& d, WHParams& p, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                      const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st);
}

int wgrad_halo_launch(const urn_conv_desc& d, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                      const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  WHParams p;
  if (wh_plan_rows(d, x_dtype, y_dtype, ldx, ldy, &p) || wh_plan(d, x_dtype, y_dtype, ldx, ldy, &p))
    return wh_launch_planned(d, p, x_dtype, x, ldx, y_dtype, y, ldy, dw, accumulate, ws, st);
  for (int kd = 0; kd < 3; ++kd) {
    urn_conv_desc sd;
    URN_CHECK(wh_split_desc(d, kd, &sd) && wh_plan(sd, x_dtype, y_dtype, ldx, ldy, &p), "wgrad_halo: unsupported shape/dtype");
    if (wh_launch_planned(sd, p, x_dtype, x, ldx, y_dtype, y, ldy, dw + (long long)kd * d.kh * d.kw * d.cx * d.cy, accumulate, ws, st))
      return 1;
  }
  return 0;
}

namespace {
int wh_launch_planned(const urn_conv_desc& d, WHParams& p, int x_dtype, const void* x, int64_t ldx, int y_dtype,
                      const void* y, int64_t ldy, float* dw, int accumulate, void* ws, cudaStream_t st) {
  const int x_bytes = ((p.stages * p.xplane_bytes + 1023) / 1024) * 1024;
  const size_t smem = 1024 + (size_t)x_bytes + (size_t)(p.ystages + p.ymirror) * p.yplane_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    URN_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    URN_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = (int)std::min<long long>(p.nunits, persistent_ctas());
  float* partial = (float*)ws;
  CUtensorMap map_x, map_y;
  memset(&map_x, 0, sizeof(map_x)); memset(&map_y, 0, sizeof(map_y));
  if (p.tma) {
    // X and dY as 4-D tensors (c, w, plane, tile row) with the plan's byte strides (images-as-rows: plane =
    // image row, tile row = image)
    EncodeTiledFnH enc = get_encode_h();
    const CUtensorMapDataType dt = x_dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    auto swz = [](int c) { return c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); };
    cuuint32_t estr[4] = {1, 1, 1, 1};
    {
      cuuint64_t dims[4] = {(cuuint64_t)p.cx, (cuuint64_t)p.XW, (cuuint64_t)p.XD, (cuuint64_t)p.XH};
      cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)p.xplane_b, (cuuint64_t)p.xrow_b};
      cuuint32_t box[4] = {(cuuint32_t)p.cx, (cuuint32_t)p.xb, 1u, (cuuint32_t)p.hr};
      CUresult r = enc(&map_x, dt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.cx),
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      URN_CHECK(r == CUDA_SUCCESS, "wgrad_halo: cuTensorMapEncodeTiled(X) failed: %d", (int)r);
    }
    {
      cuuint64_t dims[4] = {(cuuint64_t)p.cy, (cuuint64_t)p.W, (cuuint64_t)p.D, (cuuint64_t)p.H};
      cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)p.yplane_b2, (cuuint64_t)p.yrow_b};
      cuuint32_t box[4] = {(cuuint32_t)p.cy, (cuuint32_t)(8 * p.ks), 1u, (cuuint32_t)kWTH};
      CUresult r = enc(&map_y, dt, 4, const_cast<void*>(y), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.cy),
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      URN_CHECK(r == CUDA_SUCCESS, "wgrad_halo: cuTensorMapEncodeTiled(dY) failed: %d", (int)r);
    }
  }
  if (p.s == 2) wgrad_halo_kernel<true><<<grid, kWThreads, smem, st>>>(p, map_x, map_y, x, y, partial);
  else wgrad_halo_kernel<false><<<grid, kWThreads, smem, st>>>(p, map_x, map_y, x, y, partial);
  URN_LAUNCH_CHECK();
  const int nw = p.ntaps * d.cx * d.cy;
  wgrad_halo_reduce_kernel<<<(nw + 127) / 128, 128, 0, st>>>(partial, nw, grid, dw, accumulate);
  URN_LAUNCH_CHECK();
  return 0;
}
}  // namespace
