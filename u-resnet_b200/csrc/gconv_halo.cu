// Halo-tile tcgen05 convolution for the narrow layers (Cin in {16,32,64}, Cout <= 64): any
// generalised-conv launch whose tap offsets form a box -- 3x3(x3) stride 1 (forward, input
// gradient), stride-2 forward (input sampled with stride 2), 1x1 shortcuts, and the parity
// classes of stride-2 input gradients / conv-transposes (1-8 taps, output placed with stride 2).
//
// Why: the tap-per-TMA-box kernel (gconv_tc.cu) issues 9 (2D) or 27 (3D) tap-shifted boxes of
// 128 rows per tile and is bound by the TMA box-row rate when the rows are only 32-64 B.  Here
// each input plane of a tile is staged ONCE, by eight producer warps with 16-byte cp.async
// (zero-fill for TF 'same' padding), into a swizzled K-major shared-memory layout whose rows
// are ordered (hy, kw, w): for every halo row hy the three horizontally shifted pixel runs
// kw = 0,1,2 sit next to each other.  The operand of tap (kh,kw) for the 16x8 sub-tile j is then
// just a descriptor view -- start kh*RW + kw*8*KS + j*8 rows, 8-row groups RW rows apart -- so all
// 9 in-plane taps (and in 3D the 3 depth taps, which accumulate into different TMEM accumulators
// of a TD-plane tile) reuse the same staged plane.  Filters for all taps stay resident in smem.
// A tile is a SUPER-tile of KS = 4/2/1 sub-tiles (16 x 32/16/8 positions, ~55 KB per staged
// plane): the per-tile costs that bound the first version of this kernel (mbarrier round trips,
// proxy fences, commits, ~1300 cycles per 128 positions, measured with URN_HALO_TIMELINE) are
// paid once per 512 positions and the halo overhead drops from 1.41x to 1.20x.
//   warps 0-7 : cp.async producers          warp 8 : TMEM alloc, filter TMA, MMA issue
//   warps 9-12: epilogue (tcgen05.ld -> fp32 raw output / bf16 gradient, accumulate flag)
// Every CTA owns a contiguous range of tiles; ring indices and tile coordinates advance
// incrementally (no divisions in the role loops).
//
// Round 2 -- stride-1 inputs (everything but the stride-2 forward convs) are staged by TMA instead:
// ONE tiled box [halo rows][8*KS + 8 pixels][Cin] per plane (hardware zero fill outside the tensor =
// 'same' padding), rows ordered (hy, w).  The kw tap is the SAME rows viewed one pixel further: the
// swizzle of a UMMA operand is a function of the shared-memory address, so a descriptor may start
// on any 16-byte boundary of a swizzled tile.  Stages shrink from 49-55 KB to 20-23 KB, the ring
// deepens to 8, the L2 -> shared-memory traffic drops to a third, and with one issuing thread a CTA
// needs 6 warps: launch classes that gain from it run TWO CTAs per SM (halo_plan).  The
// images-as-rows plan (halo_plan_rows) relabels a 2D batch as a volume so that the kh taps are
// concatenated along N like the depth taps of a real volume.  Accumulating bf16 outputs leave
// through the TMA reduce unit (cp.reduce.async.bulk.tensor .add from swizzled staging buffers).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gconv_tc.cuh"
#include "halo_common.cuh"

namespace {

constexpr int kHThreads = 416;       // 13 warps: 8 producers, 1 MMA, 4 epilogue
constexpr int kHMaxStages = 16;         // barrier slots; the ring depth in use is max_stages() (URN_HALO_MAXST)
inline int max_stages() { static const int v = getenv("URN_HALO_MAXST") ? std::max(2, std::min(kHMaxStages, atoi(getenv("URN_HALO_MAXST")))) : 8; return v; }
constexpr int kMaxAcc = 8;           // TMEM accumulator ring (hides the MMA<->epilogue hand-off latency)
constexpr int kMaxPF = 4;            // cp.async groups kept in flight per producer thread (<= stages-1)
constexpr int kTH = 16, kTW = 8;     // 128 positions per sub-tile
constexpr int kNCH = 14;             // 16-byte chunks per producer thread and plane (<= 55 KB planes)

struct HaloParams {
  int n, D, H, W;            // INPUT tensor dims
  int OD, OH, OW;            // full OUTPUT tensor dims
  int gd, gh, gw;            // output sub-grid of this launch (outputs per dim)
  int is, os;                // input sampling stride, output placement stride
  int ood, ooh, oow;         // output placement offset (parity class)
  int nkd, nkh, nkw;         // tap box extents
  int od0, oh0, ow0;         // smallest tap offset per dim
  int hr;                    // staged halo rows per plane = (kTH-1)*is + nkh
  int cin, cout;
  int cout_store;            // channels actually written (< cout when the filter was zero-padded)
  long long ldin, ldout;
  int TD;                    // output planes per tile
  int ks, rw;                // sub-tiles per tile (along w); staged rows per halo row = nkw * 8 * ks
  int ntw, nth, ntd;
  long long ntiles;
  int stages, plane_bytes, w_tap_bytes, ntaps_w;
  int tapw[27];              // packed-filter tap index of box position (kd_i*9 + kh_i*3 + kw_i)
  int accumulate, out_dtype, in_dtype;
  uint32_t idesc;
  uint64_t adesc_hi, bdesc_hi;
  int tmem_cols;
  int nacc;                  // accumulator buffers in the TMEM ring
  int ncls;                  // output parity classes computed by this launch (1, or 4 / 8 for stride-2 scatter)
  int coff[8][3];            // output placement offset (d,h,w) per class
  int ctapw[8][27];          // per class: packed-filter tap of box position, -1 = the class has no such tap
  int concat;                // 3x3x3 stride 1: the 3 depth taps of a staged plane are ONE MMA with N = 3*cout
  int zero_bytes;            // shared-memory zero block (accumulator clearing MMAs), 0 if !concat
  uint32_t idesc_n[3];       // instruction descriptors for N = 1,2,3 x cout
  uint32_t idesc_zero;       // N = TD*cout
  int in_row_b;              // input strides in bytes: next row (y), next plane (z), next image (n)
  long long in_plane_b, in_img_b;
  long long o_row, o_plane, o_img;   // output strides in pixels
  float* stats_part;         // fused BN statistics: per (CTA, epilogue warp, channel) (n, mean, M2) partials, or NULL
  // TMA staging (stride-1 inputs): the halo plane is nkw tiled TMA boxes [hr rows][run pixels][cin] -- one per
  // kw shift, hardware zero fill outside the tensor = TF 'same' padding -- issued by ONE thread and completed on
  // the stage's mbarrier by byte count; staged rows are then ordered (kw, hy, w) instead of (hy, kw, w)
  int tma;                   // 0 cp.async producers, 1 one TMA box per kw shift, 2 ONE box per plane (kw taps = views shifted by whole pixels)
  int xb;                    // tma == 2: staged pixels per halo row
  int dual;                  // TMA staging with 6 warps per CTA and TWO CTAs per SM (half-size rings each): two independent tile pipelines hide each other's hand-off latencies
  int tma_slot[3];           // coordinate slot (2..4) of the tile row origin / plane / image in the tensor map (dims sorted by stride)
  // Accumulating bf16 outputs through the TMA reduce unit (stride-1 outputs, one class): every epilogue warp packs
  // its 32 positions x cout channels into shared memory and ONE thread issues cp.reduce.async.bulk.tensor .add --
  // the addition happens in L2, the SM never reads the old gradient.
  int red, red_bytes;        // on/off; shared memory of the 4 warps x 2 staging buffers
  int red_slot[3];           // coordinate slot (2..4) of tile row / plane / image in the output tensor map
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

// Does any depth tap of the tile's output planes [d0, d0+nd_out) read input plane z?
__device__ __forceinline__ bool plane_used(const HaloParams& p, int z, int d0, int nd_out) {
  if (p.is == 1) return true;
  for (int kdi = 0; kdi < p.nkd; ++kdi) {
    const int tt = z - p.od0 - kdi;
    if (tt < 0 || (tt & 1)) continue;
    const int dd = (tt >> 1) - d0;
    if (dd >= 0 && dd < nd_out) return true;
  }
  return false;
}

// OUTK: 0 fp32 raw output, 1 bf16 gradient, 2 fp16 raw output (mix16: the statistics still come from
// the fp32 accumulators, only the stored copy is rounded).  MODE: 0 vector store, 1 vector accumulate,
// 2 partial scalar store (zero-padded filter).  Template parameters instead of runtime switches keep
// each instantiation's code small.
template <int OUTK, int MODE, bool FULL3, bool STATS>   // FULL3: 3x3 in-plane tap box (unrolled issue loop); STATS: fused BN statistics
__global__ void __launch_bounds__(kHThreads, 1)
gconv_halo_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_o, const __grid_constant__ HaloParams p,
                  const void* __restrict__ in, void* __restrict__ out, long long* __restrict__ tl) {
  constexpr bool OUT_F32 = OUTK == 0;
  using T16 = typename std::conditional<OUTK == 2, f16, bf16>::type;     // 16-bit output element
  static_assert(!(OUTK == 2 && MODE == 1), "fp16 raw outputs are never accumulated");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                             // resident filters
  // optional timeline (URN_HALO_TIMELINE): CTA 0 records clock64() at role checkpoints of its first 64 tiles
#define TL(role, k, pt) do { if (tl && blockIdx.x == 0 && (k) < 64) tl[((role) * 64 + (k)) * 8 + (pt)] = clock64(); } while (0)
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  uint8_t* smem_zero = smem + w_bytes;                                // zeros (B operand of the accumulator-clearing MMAs)
  uint8_t* smem_a = smem_zero + p.zero_bytes;                         // staged planes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * p.plane_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kHMaxStages;
  uint64_t* tfull = bars + 2 * kHMaxStages;
  uint64_t* tempty = tfull + kMaxAcc;
  uint64_t* wfull = tempty + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  uint8_t* smem_red = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars) + 512 + 1023) & ~uintptr_t(1023));

  const long long tl_t0 = tl ? clock64() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // roles: 13 warps (0-7 producers, 8 MMA, 9-12 epilogue), or -- dual mode, 6 warps -- 0 TMA, 1 MMA, 2-5 epilogue
  const int w_mma = p.dual ? 1 : kMmaWarp;
  const int n_prod = p.dual ? 1 : kNumProducerWarps;
  const int w_epi0 = w_mma + 1;
  const int pitch = p.cin * 2;                  // bytes per staged row
  const int cpr = pitch / 16;                   // 16-byte chunks per row
  // contiguous tile range of this CTA
  const long long t_begin = p.ntiles * (long long)blockIdx.x / gridDim.x;
  const long long t_end = p.ntiles * (long long)(blockIdx.x + 1) / gridDim.x;
  const int my_tiles = (int)(t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { h_mbar_init(&full[s], p.tma ? 1 : kNumProducerWarps); h_mbar_init(&empty[s], 1); }
    for (int a = 0; a < p.nacc; ++a) { h_mbar_init(&tfull[a], 1); h_mbar_init(&tempty[a], 4); }
    h_mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == w_mma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(h_smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  h_fence_before();
  __syncthreads();
  h_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < n_prod && p.tma) {
    // ===================== TMA producer: nkw box loads per input plane, one elected thread =====================
    if (warp == 0) {
      TileIter ti; ti.init(p, t_begin);
      Ring st; st.init(p.stages);
      const int run = 8 * p.ks;
      const uint32_t box_bytes = p.tma == 2 ? (uint32_t)p.plane_bytes : (uint32_t)(p.hr * run * pitch);
      for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
        const int d0 = ti.dt * p.TD;
        const int ybase = ti.ht * kTH + p.oh0, xbase = ti.wt * kTW * p.ks + p.ow0;
        const int nd_out = min(d0 + p.TD, p.gd) - d0;
        const int zlo = d0 + p.od0, zhi = d0 + nd_out - 1 + p.od0 + p.nkd - 1;
        for (int z = zlo; z <= zhi; ++z) {
          if (z < 0 || z >= p.D) continue;                 // a plane of zeros contributes nothing
          h_mbar_wait_warp(&empty[st.idx], st.phase ^ 1);
          if (lane == 0) TL(0, k, 0);
          int crd[3];
          crd[p.tma_slot[0] - 2] = ybase; crd[p.tma_slot[1] - 2] = z; crd[p.tma_slot[2] - 2] = ti.nn;
          if (h_elect_one()) {
            const int nbox = p.tma == 2 ? 1 : p.nkw;
            h_mbar_expect_tx(&full[st.idx], box_bytes * (uint32_t)nbox);
            const uint32_t dst = h_smem_u32(smem_a + (size_t)st.idx * p.plane_bytes);
            for (int kw = 0; kw < nbox; ++kw)
              asm volatile(
                  "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
                  " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                  ::"r"(dst + (uint32_t)kw * box_bytes), "l"(reinterpret_cast<uint64_t>(&map_a)), "r"(h_smem_u32(&full[st.idx])),
                  "r"(0), "r"(xbase + kw), "r"(crd[0]), "r"(crd[1]), "r"(crd[2])
                  : "memory");
          }
          __syncwarp();
          if (lane == 0) TL(0, k, 2);
          st.next();
        }
      }
    }
  } else if (warp < n_prod) {
    // ===================== producers: stage one input plane per pipeline slot =====================
    // Per-thread chunk tables are tile-invariant and live in registers: byte offset of the chunk
    // from the tile's halo origin in global memory, byte offset of its swizzled slot in the
    // stage, and (halo row, halo column) for the border test.  Slots beyond the plane's chunk
    // count replicate slot 0 (the same 16 bytes are copied twice), so the copy loop has no
    // validity branch; interior tiles skip the border test altogether.  (The first version
    // decoded a packed descriptor per chunk behind a `continue`: ~25 dependent instructions and
    // a divergence region per 16 bytes, 150 cycles per chunk -- it bounded the whole kernel.)
    const int t = threadIdx.x;
    const int total = p.hr * p.rw * cpr;
    const int cpr_log2 = cpr == 8 ? 3 : (cpr == 4 ? 2 : 1);
    const int sw_shift = cpr == 8 ? 0 : (cpr == 4 ? 1 : 2);
    const int sw_mask = cpr == 8 ? 7 : (cpr == 4 ? 3 : 1);
    const int run = 8 * p.ks;                                  // pixels per (hy, kw) run
    const char* src_base = reinterpret_cast<const char*>(in);
    const int pix_bytes = (int)(p.ldin * 2);        // one plane is < 2 GB: 32-bit offsets inside it
    const int row_bytes = p.in_row_b;
    const int H = p.H, W = p.W;
    int soff[kNCH], hyv[kNCH], dxv[kNCH];
    uint32_t doff[kNCH];
#pragma unroll
    for (int j = 0; j < kNCH; ++j) {
      int idx = t + j * kProducerThreads;
      if (idx >= total) idx = t;                               // replicate slot 0
      const int row = idx >> cpr_log2, c = idx & (cpr - 1);
      const int hy = row / p.rw;
      const int rem = row - hy * p.rw;
      const int kw = rem / run;
      const int dx = kw + (rem - kw * run) * p.is;             // input column of this row, relative to the tile origin
      const int sw = (row >> sw_shift) & sw_mask;              // K-major SW128 / SW64 / SW32 chunk swizzle
      doff[j] = (uint32_t)(row * (pitch >> 4) + (c ^ sw)) << 4;
      soff[j] = hy * row_bytes + dx * pix_bytes + c * 16;
      hyv[j] = hy; dxv[j] = dx;
    }
    TileIter ti; ti.init(p, t_begin);
    Ring st, ar; st.init(p.stages); ar.init(p.stages);
    uint32_t issued = 0;
    const int pf = max(1, min(kMaxPF, p.stages / 2));
    const bool no_load = (p.dbg & 2) != 0;
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      const int ybase = ti.ht * kTH * p.is + p.oh0, xbase = ti.wt * kTW * p.ks * p.is + p.ow0;
      const bool interior = ybase >= 0 && ybase + p.hr <= H && xbase >= 0 && xbase + (run - 1) * p.is + p.nkw <= W;
      const long long org_off = (long long)ybase * row_bytes + (long long)xbase * pix_bytes;
      const int nd_out = min(d0 + p.TD, p.gd) - d0;
      const int zlo = d0 * p.is + p.od0, zhi = (d0 + nd_out - 1) * p.is + p.od0 + p.nkd - 1;
      for (int z = zlo; z <= zhi; ++z) {
        if (z < 0 || z >= p.D) continue;                 // a plane of zeros contributes nothing
        if (!plane_used(p, z, d0, nd_out)) continue;     // stride 2: no tap of this tile reads the plane
        h_mbar_wait_warp(&empty[st.idx], st.phase ^ 1);
        if (t == 0) TL(0, k, 0);
        const uint32_t dst_base = h_smem_u32(smem_a + (size_t)st.idx * p.plane_bytes);
        const char* origin = src_base + (long long)ti.nn * p.in_img_b + (long long)z * p.in_plane_b + org_off;
        if (no_load) {
        } else if (interior) {
#pragma unroll
          for (int j = 0; j < kNCH; ++j) cp_async16_zfill(dst_base + doff[j], origin + soff[j], 16u);
        } else {
#pragma unroll
          for (int j = 0; j < kNCH; ++j) {
            const bool ok = (unsigned)(ybase + hyv[j]) < (unsigned)H && (unsigned)(xbase + dxv[j]) < (unsigned)W;
            cp_async16_zfill(dst_base + doff[j], ok ? origin + soff[j] : src_base, ok ? 16u : 0u);
          }
        }
        // One barrier arrival per WARP (256 per-thread arrivals on one mbarrier word serialise
        // for ~2 us per stage): each thread waits until its copies of the stage issued kPF stages
        // ago have landed and fences them towards the async proxy, the warp converges,
        // lane 0 arrives.
        if (!(p.dbg & 4096)) asm volatile("cp.async.commit_group;" ::: "memory");
        if (t == 0) TL(0, k, 1);
        if (issued >= (uint32_t)pf) {
          if (p.dbg & 4096) {}
          else if (pf >= 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
          else if (pf == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
          else if (pf == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
          else asm volatile("cp.async.wait_group 1;" ::: "memory");
          // writer-side generic->async proxy fence: the copies that just landed are read by UMMA
          if (!(p.dbg & 16384)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    if (!(p.dbg & 16384)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (uint32_t j = (issued > (uint32_t)pf ? issued - pf : 0); j < issued; ++j) { h_mbar_arrive(&full[ar.idx]); ar.next(); }
  } else if (warp == w_mma) {
    // ===================== filter load + MMA issue =====================
    // The WHOLE warp runs this loop with warp-uniform values (so the address arithmetic lives in
    // the uniform datapath and no lane-0 branch serialises it); only the tcgen05 instructions
    // themselves are issued by one elected lane.
    // Filter slots: tap widx goes to slot widx, or -- concat mode -- the three depth taps of each
    // (kh,kw) are placed next to each other in DESCENDING kd order: they form one [3*cout][cin]
    // B tile whose N blocks line up with the ascending output planes z-1, z, z+1 of input plane z.
    for (int i = lane * 16; i < p.zero_bytes; i += 32 * 16) *reinterpret_cast<uint4*>(smem_zero + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros -> read by UMMA (async proxy)
    if (p.ncls > 1) {
      if (h_elect_one()) {
        h_mbar_expect_tx(wfull, (uint32_t)(p.ntaps_w * p.w_tap_bytes));
        for (int tp = 0; tp < p.ntaps_w; ++tp)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
              " [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(h_smem_u32(smem_w + (size_t)tp * p.w_tap_bytes)), "l"(reinterpret_cast<uint64_t>(&map_w)),
              "r"(h_smem_u32(wfull)), "r"(0), "r"(0), "r"(tp)
              : "memory");
      }
    } else if (h_elect_one()) {
      h_mbar_expect_tx(wfull, (uint32_t)(p.nkd * p.nkh * p.nkw * p.w_tap_bytes));
      for (int kd = 0; kd < p.nkd; ++kd)
        for (int kh = 0; kh < p.nkh; ++kh)
          for (int kw = 0; kw < p.nkw; ++kw) {
            const int tp = p.tapw[kd * 9 + kh * 3 + kw];
            const int slot = p.concat ? ((kh * 3 + kw) * 3 + (2 - kd)) : tp;
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(h_smem_u32(smem_w + (size_t)slot * p.w_tap_bytes)), "l"(reinterpret_cast<uint64_t>(&map_w)),
                "r"(h_smem_u32(wfull)), "r"(0), "r"(0), "r"(tp)
                : "memory");
          }
    }
    __syncwarp();
    // filter descriptor low words of the 27 (od,kh,kw) taps (all present in this kernel)
    const uint32_t w_lo0 = h_smem_u32(smem_w) >> 4;
    const uint32_t b_hi = (uint32_t)(p.bdesc_hi >> 32), a_hi = (uint32_t)(p.adesc_hi >> 32);
    const uint32_t b_lo_base = (uint32_t)p.bdesc_hi, a_lo_base = (uint32_t)p.adesc_hi;
    const uint32_t b_base = b_lo_base | w_lo0;
    const uint32_t w_tap16 = (uint32_t)p.w_tap_bytes >> 4;
    // tap view (kh,kw) of sub-tile j starts at staged row kh*RW + kw*8*KS + j*8
    const uint32_t row16 = (uint32_t)pitch >> 4;                  // 16-byte units per staged row
    // descriptor start offsets (16-byte units) of the next kw tap / next kh tap (one halo row) / next sub-tile.
    // cp.async staging orders the rows (hy, kw, w); TMA staging (kw, hy, w): one box per kw shift
    // (tma == 2: rows (hy, w) staged once, the kw tap is a view shifted by one pixel)
    const uint32_t a_kw = p.tma == 2 ? row16 : (p.tma ? (uint32_t)(p.hr * 8 * p.ks) * row16 : (uint32_t)(8 * p.ks) * row16);
    const uint32_t a_kh = p.tma == 2 ? (uint32_t)p.xb * row16 : (p.tma ? (uint32_t)(8 * p.ks) * row16 : (uint32_t)p.rw * row16);
    const uint32_t a_sub = 8u * row16;
    const int ksteps = p.cin >> 4;                                 // K = 16 per MMA
    const bool no_mma = (p.dbg & 4) != 0;
    h_mbar_wait_warp(wfull, 0);
    TileIter ti; ti.init(p, t_begin);
    Ring st, ac; st.init(p.stages); ac.init(p.nacc);
    const uint32_t acc_cols = (uint32_t)(p.ncls * p.TD * p.ks * p.cout);
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      if (lane == 0) TL(1, k, 0);
      h_mbar_wait_warp(&tempty[ac.idx], ac.phase ^ 1);
      h_fence_after();
      if (lane == 0) TL(1, k, 1);
      uint32_t touched = 0;                              // bit dd: accumulators of plane dd written
      const int nd_out = min(d0 + p.TD, p.gd) - d0;      // output planes of this tile
      const int nsub = min(p.ks, (p.gw - ti.wt * kTW * p.ks + kTW - 1) / kTW);   // sub-tiles inside the sub-grid
      const uint32_t d_tmem0 = tmem_base + (uint32_t)ac.idx * acc_cols;
      const int zlo = d0 * p.is + p.od0, zhi = (d0 + nd_out - 1) * p.is + p.od0 + p.nkd - 1;
      for (int z = zlo; z <= zhi; ++z) {
        if (z < 0 || z >= p.D) continue;
        if (!plane_used(p, z, d0, nd_out)) continue;
        h_mbar_wait_warp(&full[st.idx], st.phase);
        if (lane == 0) TL(1, k, 2);
        // (the cp.async -> UMMA proxy fence is issued by the producers before they arrive; a
        // consumer-side fence here stalls the issue warp for up to ~900 cycles per tile)
        if (p.dbg & 2048) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        h_fence_after();
        if (lane == 0) TL(1, k, 3);
        const uint32_t a_lo0 = a_lo_base | (h_smem_u32(smem_a + (size_t)st.idx * p.plane_bytes) >> 4);
        if (p.ncls > 1) {
          // Parity classes of a stride-2 scatter (conv input gradient / conv-transpose): the staged
          // plane of the SMALL side feeds every class; class c owns the accumulators
          // [(c*TD + dd)*ks + sj] and the subset of the 2x2(x2) tap box its parity allows.
#pragma unroll 1
          for (int kdi = 0; kdi < p.nkd; ++kdi) {
            const int dd = z - p.od0 - kdi - d0;
            if (dd < 0 || dd >= nd_out) continue;
#pragma unroll 1
            for (int c = 0; c < p.ncls; ++c) {
              int any = 0;
              for (int j = 0; j < 9; ++j) any |= (p.ctapw[c][kdi * 9 + j] >= 0);
              if (!any) continue;
              const uint32_t bit = (uint32_t)(c * p.TD + dd);
              const uint32_t accf = (touched >> bit) & 1u;
              touched |= 1u << bit;
              if (!no_mma && h_elect_one()) {
#pragma unroll 1
                for (int sj = 0; sj < nsub; ++sj) {
                  const uint32_t d_tmem = d_tmem0 + (uint32_t)(((c * p.TD + dd) * p.ks + sj) * p.cout);
                  const uint32_t a_sj = a_lo0 + (uint32_t)sj * a_sub;
                  uint32_t acc = accf;
#pragma unroll 1
                  for (int kh = 0; kh < p.nkh; ++kh)
#pragma unroll 1
                    for (int kw = 0; kw < p.nkw; ++kw) {
                      const int tp = p.ctapw[c][kdi * 9 + kh * 3 + kw];
                      if (tp < 0) continue;
                      const uint32_t a_lo = a_sj + (uint32_t)kh * a_kh + (uint32_t)kw * a_kw;
                      const uint32_t b_lo = b_base + (uint32_t)tp * w_tap16;
                      h_umma2(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, acc);
                      acc = 1u;
                      if (ksteps > 1) h_umma2(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, p.idesc, 1u);
                      if (ksteps > 2) {
                        h_umma2(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, p.idesc, 1u);
                        h_umma2(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, p.idesc, 1u);
                      }
                    }
                }
              }
              __syncwarp();
            }
          }
        } else if (p.concat) {
          // output planes fed by input plane z: z-1-d0 .. z+1-d0 (depth taps kd = 2, 1, 0), clipped to the tile
          const int dlo = max(0, z - 1 - d0), dhi = min(nd_out - 1, z + 1 - d0);
          const int skip = dlo - (z - 1 - d0);
          const uint32_t idn = p.idesc_n[dhi - dlo];
          if (!no_mma && dhi >= dlo && h_elect_one()) {
#pragma unroll 1
            for (int sj = 0; sj < nsub; ++sj) {
              const uint32_t a_sj = a_lo0 + (uint32_t)sj * a_sub;
              if (!touched)      // first plane of the tile: clear all TD accumulators of this sub-tile (D = A * 0)
                h_umma2(d_tmem0 + (uint32_t)(sj * p.TD * p.cout), a_sj, a_hi, b_lo_base | (h_smem_u32(smem_zero) >> 4), b_hi,
                        p.idesc_zero, 0u);
              const uint32_t d_tmem = d_tmem0 + (uint32_t)((sj * p.TD + dlo) * p.cout);
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                if (!FULL3 && kh >= p.nkh) break;      // images-as-rows plan of a 2D layer: nkh = 1
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const uint32_t a_lo = a_sj + (uint32_t)kh * a_kh + (uint32_t)kw * a_kw;
                  const uint32_t b_lo = b_base + (uint32_t)((kh * 3 + kw) * 3 + skip) * w_tap16;
                  h_umma2(d_tmem, a_lo, a_hi, b_lo, b_hi, idn, 1u);
                  if (ksteps > 1) h_umma2(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, idn, 1u);
                  if (ksteps > 2) {
                    h_umma2(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, idn, 1u);
                    h_umma2(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, idn, 1u);
                  }
                }
              }
            }
          }
          __syncwarp();
          touched = 1u;
        } else
#pragma unroll 1
        for (int kdi = 0; kdi < p.nkd; ++kdi) {
          const int tt = z - p.od0 - kdi;                // = dd_global * is
          if (tt < 0 || (p.is == 2 && (tt & 1))) continue;
          const int dd = (p.is == 2 ? tt >> 1 : tt) - d0;
          if (dd < 0 || dd >= nd_out) continue;
          const uint32_t accf = (touched >> dd) & 1u;
          touched |= 1u << dd;
          if (!no_mma && h_elect_one()) {
#pragma unroll 1
            for (int sj = 0; sj < nsub; ++sj) {
              const uint32_t d_tmem = d_tmem0 + (uint32_t)((dd * p.ks + sj) * p.cout);
              const uint32_t a_sj = a_lo0 + (uint32_t)sj * a_sub;
              uint32_t acc = accf;
              const int nkh = FULL3 ? 3 : p.nkh, nkw = FULL3 ? 3 : p.nkw;
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                if (!FULL3 && kh >= nkh) break;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  if (!FULL3 && kw >= nkw) break;
                  const uint32_t a_lo = a_sj + (uint32_t)kh * a_kh + (uint32_t)kw * a_kw;
                  const uint32_t b_lo = b_base + (uint32_t)p.tapw[kdi * 9 + kh * 3 + kw] * w_tap16;
                  h_umma2(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, acc);
                  acc = 1u;
                  if (ksteps > 1) h_umma2(d_tmem, a_lo + 2, a_hi, b_lo + 2, b_hi, p.idesc, 1u);
                  if (ksteps > 2) {
                    h_umma2(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, p.idesc, 1u);
                    h_umma2(d_tmem, a_lo + 6, a_hi, b_lo + 6, b_hi, p.idesc, 1u);
                  }
                }
              }
            }
          }
          __syncwarp();
        }
        if (lane == 0) TL(1, k, 4);
        if (h_elect_one()) { if (p.dbg & 1024) h_mbar_arrive(&empty[st.idx]); else h_commit(&empty[st.idx]); }
        __syncwarp();
        st.next();
      }
      if (h_elect_one()) { if (p.dbg & 1024) h_mbar_arrive(&tfull[ac.idx]); else h_commit(&tfull[ac.idx]); }
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
    const uint32_t acc_cols = (uint32_t)(p.ncls * p.TD * p.ks * p.cout);
    constexpr bool do_stats = STATS;
    RunStat rs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { rs[i].n = 0.f; rs[i].mean = 0.f; rs[i].m2 = 0.f; }
    uint32_t red_buf = 0;
    for (int k = 0; k < my_tiles; ++k, ti.next(p)) {
      const int d0 = ti.dt * p.TD;
      const int y = ti.ht * kTH + lh, x0 = ti.wt * kTW * p.ks + lw;
      const int nsub = min(p.ks, (p.gw - ti.wt * kTW * p.ks + kTW - 1) / kTW);
      const int dmax = min(d0 + p.TD, p.gd);
      if (warp == w_epi0 && lane == 0) TL(2, k, 0);
      // accumulate mode (bf16 gradients): the old values of a group of sub-tiles are fetched BEFORE
      // the accumulators are waited for, so the read-modify-write does not serialise behind the MMAs
      uint4 oldv[4][2];
      bool old_ready = false;
      auto fetch_old = [&](long long opix0_, int c0_) {
#pragma unroll
        for (int sj = 0; sj < 4; ++sj) {
          if (sj < nsub && y < p.gh && x0 + sj * kTW < p.gw) {
            const bf16* o = reinterpret_cast<const bf16*>(out) + (opix0_ + sj * kTW * p.os) * p.ldout + c0_;
            oldv[sj][0] = *reinterpret_cast<const uint4*>(o);
            oldv[sj][1] = *reinterpret_cast<const uint4*>(o + 8);
          }
        }
      };
      if (!OUT_F32 && MODE == 1 && !p.red) {
        const long long opix_first = (long long)ti.nn * p.o_img + (long long)(d0 * p.os + p.coff[0][0]) * p.o_plane +
                                     (long long)(y * p.os + p.coff[0][1]) * p.o_row + (x0 * p.os + p.coff[0][2]);
        fetch_old(opix_first, 0);
        old_ready = true;
      }
      h_mbar_wait_warp(&tfull[ac.idx], ac.phase);
      h_fence_after();
      if (warp == w_epi0 && lane == 0) TL(2, k, 1);
      if (!OUT_F32 && MODE == 1 && p.red) {
        // accumulate through the TMA reduce unit: per (plane, sub-tile) the warp's 32 x cout block goes to one of
        // its two staging buffers (swizzled like the tensor map, conflict-free 16-byte stores), then one box
        // [cout][8 w][4 rows] is added to the gradient in L2
        const int rpitch = p.cout * 2;
        const int rsw = rpitch == 128 ? (lane & 7) : (rpitch == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
        uint8_t* rbase = smem_red + (size_t)q * 2 * 32 * rpitch;
        for (int dd = 0; dd < p.TD; ++dd) {
          const int d = d0 + dd;
          if (d >= dmax) continue;
          const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ac.idx * acc_cols +
                                  (uint32_t)((p.concat ? dd : dd * p.ks) * p.cout);
          const uint32_t sj_cols = (uint32_t)((p.concat ? p.TD : 1) * p.cout);
          for (int sj = 0; sj < nsub; ++sj) {
            uint8_t* rb = rbase + (size_t)(red_buf & 1) * 32 * rpitch;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer's previous box has been read
            __syncwarp();
            for (int c0 = 0; c0 < p.cout; c0 += 16) {
              float v[16];
              h_tmem_ld16(taddr0 + (uint32_t)sj * sj_cols + (uint32_t)c0, v);
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&b2);
              }
              const int ch = c0 >> 3;                                  // 16-byte chunk of this position's row
              *reinterpret_cast<uint4*>(rb + lane * rpitch + ((ch ^ rsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(rb + lane * rpitch + (((ch + 1) ^ rsw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && !(p.dbg & 1)) {
              int crd[3];
              crd[p.red_slot[0] - 2] = ti.ht * kTH + 4 * q; crd[p.red_slot[1] - 2] = d; crd[p.red_slot[2] - 2] = ti.nn;
              asm volatile(
                  "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&map_o)), "r"(h_smem_u32(rb)), "r"(0), "r"((ti.wt * p.ks + sj) * kTW),
                  "r"(crd[0]), "r"(crd[1]), "r"(crd[2])
                  : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++red_buf;
          }
        }
      } else
      for (int cd = 0; cd < p.ncls * p.TD; ++cd) {
        const int cls = cd / p.TD, dd = cd - cls * p.TD;
        const int d = d0 + dd;
        if (d >= dmax) continue;                       // warp-uniform: accumulators never written
        const long long opix0 = (long long)ti.nn * p.o_img + (long long)(d * p.os + p.coff[cls][0]) * p.o_plane +
                                (long long)(y * p.os + p.coff[cls][1]) * p.o_row + (x0 * p.os + p.coff[cls][2]);
        // accumulator columns: [class][dd][sj][cout], or [sj][dd][cout] in concat mode
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ac.idx * acc_cols +
                                (uint32_t)((p.concat ? dd : cd * p.ks) * p.cout);
        const uint32_t sj_cols = (uint32_t)((p.concat ? p.TD : 1) * p.cout);
        for (int c0 = 0; c0 < p.cout_store; c0 += 16) {
          // TMEM loads of a group of sub-tiles in flight, one wait, then statistics + stores.  With
          // fused statistics the group is 2 sub-tiles (the running sums need the registers).
          constexpr int G = STATS ? 2 : 4;
          float sa[16], sb[16];
          unsigned nv = 0;
          if (STATS) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
          }
#pragma unroll
          for (int s0 = 0; s0 < 4; s0 += G) {
            if (s0 >= nsub) break;
            float v[G][16];
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (s0 + g < nsub) h_tmem_ld16_nowait(taddr0 + (uint32_t)(s0 + g) * sj_cols + (uint32_t)c0, v[g]);
            if (!OUT_F32 && MODE == 1) {
              if (!old_ready) fetch_old(opix0, c0);
              old_ready = false;
            }
            h_tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const int sj = s0 + g;
              if (sj >= nsub) continue;
              const int x = x0 + sj * kTW;
              const bool valid = y < p.gh && x < p.gw;
              if (STATS) {
                nv += __popc(__ballot_sync(0xffffffffu, valid));
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { sa[j] += v[g][j]; sb[j] = fmaf(v[g][j], v[g][j], sb[j]); }
                }
              }
              if (!valid || (p.dbg & 1)) continue;
              const long long opix = opix0 + sj * kTW * p.os;
              if (MODE == 2) {                           // zero-padded filter: only cout_store channels exist
                if (OUT_F32) {
                  float* o = reinterpret_cast<float*>(out) + opix * p.ldout + c0;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.cout_store) o[j] = p.accumulate ? o[j] + v[g][j] : v[g][j];
                } else {
                  T16* o = reinterpret_cast<T16*>(out) + opix * p.ldout + c0;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.cout_store) st1(o + j, p.accumulate ? ld1(o + j) + v[g][j] : v[g][j]);
                }
              } else if (OUT_F32) {
                float* o = reinterpret_cast<float*>(out) + opix * p.ldout + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  float4 xx = make_float4(v[g][j], v[g][j + 1], v[g][j + 2], v[g][j + 3]);
                  if (MODE == 1) {
                    float4 yy = *reinterpret_cast<float4*>(o + j);
                    xx.x += yy.x; xx.y += yy.y; xx.z += yy.z; xx.w += yy.w;
                  }
                  *reinterpret_cast<float4*>(o + j) = xx;
                }
              } else {
                T16* o = reinterpret_cast<T16*>(out) + opix * p.ldout + c0;
                if (MODE == 1) {
                  const uint32_t w[8] = {oldv[sj][0].x, oldv[sj][0].y, oldv[sj][0].z, oldv[sj][0].w,
                                         oldv[sj][1].x, oldv[sj][1].y, oldv[sj][1].z, oldv[sj][1].w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    v[g][2 * j] += __uint_as_float(w[j] << 16);
                    v[g][2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
                  }
                }
                store8(o, v[g]); store8(o + 8, v[g] + 8);
              }
            }
          }
          if (STATS) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
              if (ci * 16 == c0) stat_merge(rs[ci], sa, sb, nv, lane);
          }
        }
      }
      if (warp == w_epi0 && lane == 0) TL(2, k, 2);
      h_fence_before();
      __syncwarp();
      if (lane == 0) h_mbar_arrive(&tempty[ac.idx]);
      if (warp == w_epi0 && lane == 0) TL(2, k, 3);
      ac.next();
    }
    if (!OUT_F32 && MODE == 1 && p.red && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (do_stats && !(lane & 1)) {
      // partial[(cta*4 + q)][channel][3]
      float* dst = p.stats_part + ((long long)(blockIdx.x * 4 + q) * p.cout) * 3;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int ch = ci * 16 + (lane >> 1);
        if (ci * 16 < p.cout_store) { dst[ch * 3 + 0] = rs[ci].n; dst[ch * 3 + 1] = rs[ci].mean; dst[ch * 3 + 2] = rs[ci].m2; }
      }
    }
  }

  if (tl && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[3 * 64 * 8 + blockIdx.x] = ((clock64() - tl_t0) << 8) | (long long)(smid & 255);
  }
  h_fence_before();
  __syncthreads();
  if (warp == w_mma) {
    h_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

bool halo_plan_impl(const GConv& g, int in_dtype, int out_dtype, int cout_store, HaloParams* pp, int ncls, bool dual);
// Dual mode: stride-1 (TMA-staged) launches with 6 warps per CTA, half the shared memory and TMEM columns, TWO
// CTAs per SM -- two independent tile pipelines hide each other's hand-off latencies.  Measured per layer
// (B200, mix16): a clear win for the parity-class launches of stride-2 input gradients / conv-transposes (2D
// 16->32 @512: 0.519 -> 0.275 ms, deconv4 0.289 -> 0.169 ms; 3D 1.64 -> 1.22 ms, 1.32 -> 0.95 ms), the 3D
// Cout = 16 full-resolution layers (0.82 -> 0.72 ms) and the zero-padded head (3D 1.16 -> 0.85 ms); a loss for
// the 3D Cout = 32 layers (0.23 -> 0.32 ms: their accumulators no longer fit a half TMEM with the 512-position
// tile) and for accumulating epilogues; neutral in 2D images-as-rows launches.  Hence the rule below
// (URN_HALO_DUAL=0: never, =2: wherever the shape fits).  Falls back to the single-CTA plan when it does not fit.
bool halo_plan(const GConv& g, int in_dtype, int out_dtype, int cout_store, HaloParams* pp, int ncls = 1) {
  static const int mode = getenv("URN_HALO_DUAL") ? atoi(getenv("URN_HALO_DUAL")) : 1;
  bool want = false;
  if (mode == 2) want = true;
  else if (mode == 1) {
    const bool is3d = g.id > 1 && g.od > 1;
    want = ncls > 1 || (is3d && g.cout == 16 && !g.accumulate && g.ntaps == 27) || (cout_store > 0 && cout_store < g.cout);
  }
  if (want && g.is == 1 && halo_plan_impl(g, in_dtype, out_dtype, cout_store, pp, ncls, true) && pp->tma) return true;
  return halo_plan_impl(g, in_dtype, out_dtype, cout_store, pp, ncls, false);
}
bool halo_plan_impl(const GConv& g, int in_dtype, int out_dtype, int cout_store, HaloParams* pp, int ncls, bool dual) {
  HaloParams& p = *pp;
  memset(&p, 0, sizeof(p));
  p.dual = dual ? 1 : 0;
  const int tmem_limit = dual ? 256 : 512;              // TMEM columns and dynamic shared memory available to one CTA
  const int smem_limit = dual ? 108 * 1024 : 224 * 1024;
  if (getenv("URN_NO_HALO")) return false;
  if (in_dtype != URN_BF16 && in_dtype != URN_F16) return false;
  if (out_dtype != URN_F32 && out_dtype != URN_BF16 && out_dtype != URN_F16) return false;
  if (out_dtype == URN_F16 && (g.accumulate || (cout_store > 0 && (cout_store & 15)))) return false;   // fp16 = freshly written raw output
  if ((g.is != 1 && g.is != 2) || (g.os != 1 && g.os != 2) || (g.is == 2 && g.os == 2)) return false;
  if (ncls != 1 && ncls != 4 && ncls != 8) return false;
  if (g.cin != 16 && g.cin != 32 && g.cin != 64) return false;
  if (g.cout % 16 || g.cout > 64) return false;
  if (cout_store <= 0) cout_store = g.cout;
  if (cout_store > g.cout) return false;
  if (g.ldin % 8 || g.ldout < cout_store) return false;
  if (cout_store % 16 == 0 && g.ldout % (out_dtype == URN_F32 ? 4 : 8)) return false;   // 16-byte vector stores
  if (g.ntaps < 1 || g.ntaps > 27 || g.gd < 1 || g.gh < 1 || g.gw < 1) return false;
  // the tap offsets must form a full box [o0, o0+n) per dim
  int lo[3] = {1 << 20, 1 << 20, 1 << 20}, hi[3] = {-(1 << 20), -(1 << 20), -(1 << 20)};
  for (int t = 0; t < g.ntaps; ++t) {
    const int o[3] = {g.taps[t].od, g.taps[t].oh, g.taps[t].ow};
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], o[k]); hi[k] = std::max(hi[k], o[k]); }
  }
  p.nkd = hi[0] - lo[0] + 1; p.nkh = hi[1] - lo[1] + 1; p.nkw = hi[2] - lo[2] + 1;
  if (p.nkd > 3 || p.nkh > 3 || p.nkw > 3 || p.nkd * p.nkh * p.nkw != g.ntaps) return false;
  p.od0 = lo[0]; p.oh0 = lo[1]; p.ow0 = lo[2];
  for (int i = 0; i < 27; ++i) p.tapw[i] = -1;
  int maxw = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    int& slot = p.tapw[(g.taps[t].od - lo[0]) * 9 + (g.taps[t].oh - lo[1]) * 3 + (g.taps[t].ow - lo[2])];
    if (slot >= 0) return false;
    slot = g.taps[t].widx;
    maxw = std::max(maxw, g.taps[t].widx);
  }
  p.n = g.n; p.D = g.id; p.H = g.ih; p.W = g.iw;
  p.OD = g.od; p.OH = g.oh; p.OW = g.ow;
  p.gd = g.gd; p.gh = g.gh; p.gw = g.gw;
  p.is = g.is; p.os = g.os; p.ood = g.ood; p.ooh = g.ooh; p.oow = g.oow;
  // every output needs one tap that is always inside the input (so that each TMEM accumulator of
  // a tile is written at least once): offset 0 in the box and the last sample inside the input
  const int gs[3] = {g.gd, g.gh, g.gw}, ins[3] = {g.id, g.ih, g.iw};
  for (int k = 0; k < 3; ++k)
    if (lo[k] > 0 || hi[k] < 0 || (gs[k] - 1) * g.is > ins[k] - 1) return false;
  p.cin = g.cin; p.cout = g.cout; p.cout_store = cout_store; p.ldin = g.ldin; p.ldout = g.ldout;
  if ((long long)p.H * p.W * g.ldin * 2 >= (1ll << 31)) return false;     // 32-bit offsets inside a plane
  p.in_row_b = (int)(g.ldin * 2 * p.W);
  p.in_plane_b = (long long)p.in_row_b * p.H;
  p.in_img_b = p.in_plane_b * p.D;
  p.o_row = p.OW; p.o_plane = (long long)p.OH * p.OW; p.o_img = p.o_plane * p.OD;
  p.hr = (kTH - 1) * g.is + p.nkh;
  const int cpr = g.cin * 2 / 16;
  // super-tile width: the widest whose staged plane fits kNCH chunks per producer thread
  p.ks = 4;
  if (const char* e = getenv("URN_HALO_KS")) p.ks = std::max(1, std::min(p.ks, atoi(e)));
  while (p.ks > 1 && (p.gw + kTW - 1) / kTW <= p.ks / 2) p.ks /= 2;
  // (the single-box TMA staging has no per-thread chunk tables: its limit is the shared memory, checked below)
  const int tmode = getenv("URN_HALO_TMA") ? std::max(0, std::min(2, atoi(getenv("URN_HALO_TMA")))) : 2;
  const bool one_box = g.is == 1 && tmode == 2 && get_encode_h() != nullptr && !getenv("URN_HALO_KS_CPASYNC");
  if (one_box) {
    // as wide as two stages allow beside the filter taps and the zero block of the concat mode (losing the
    // N-concatenation costs far more than a narrow tile), but never narrower than the cp.async rule's choice
    int ks_cp = p.ks;
    while (ks_cp > 1 && p.hr * p.nkw * 8 * ks_cp * cpr > kNCH * 256) ks_cp /= 2;
    const int zero_est = (g.gd > 1 || p.nkd > 1) ? ((4 * g.cout * g.cin * 2 + 1023) / 1024) * 1024 : 0;
    const int room = smem_limit - (((maxw + 1) * g.cout * g.cin * 2 + 1023) / 1024) * 1024 - zero_est - 2048;
    while (p.ks > ks_cp && 2 * p.hr * (8 * p.ks + 8) * g.cin * 2 > room) p.ks /= 2;
  } else
    while (p.ks > 1 && p.hr * p.nkw * 8 * p.ks * cpr > kNCH * 256) p.ks /= 2;
  const bool has_d = g.gd > 1 || p.nkd > 1;
  p.TD = has_d ? 4 : 1;
  p.ncls = ncls;
  if (dual)        // keep the depth tile (halo planes per output plane), narrow the super-tile first
    while (p.ks > 1 && ncls * p.TD * p.ks * g.cout * 2 > tmem_limit) p.ks /= 2;
  while (p.TD > 1 && ncls * p.TD * p.ks * g.cout * 2 > tmem_limit) p.TD /= 2;
  while (p.ks > 1 && ncls * p.TD * p.ks * g.cout * 2 > tmem_limit) p.ks /= 2;
  if (ncls * p.TD * p.ks * g.cout * 2 > tmem_limit) return false;
  p.rw = p.nkw * 8 * p.ks;
  if (!one_box && p.hr * p.rw * cpr > kNCH * 256) return false;
  if (p.gd < p.TD) p.TD = std::max(1, p.gd);
  p.ntw = (p.gw + kTW * p.ks - 1) / (kTW * p.ks); p.nth = (p.gh + kTH - 1) / kTH; p.ntd = (p.gd + p.TD - 1) / p.TD;
  p.ntiles = (long long)p.n * p.ntd * p.nth * p.ntw;
  // TMA staging for stride-1 inputs (URN_HALO_TMA=0 keeps the cp.async producers, =1 one box per kw shift with
  // the rows ordered (kw, hy, w)); default 2: ONE box of hr x xb pixels per plane, the kw taps are descriptor
  // views shifted by whole pixels (the swizzle is a function of the shared-memory address, so a view may start
  // anywhere on a 16-byte boundary) -- a third of the L2 -> shared-memory traffic and 2.4x smaller stages
  const uint64_t pitch = (uint64_t)g.cin * 2;
  p.tma = g.is == 1 ? tmode : 0;
  if (p.tma == 1 && ((uint64_t)(p.hr * 8 * p.ks) * pitch) % (pitch == 128 ? 1024 : (pitch == 64 ? 512 : 256))) p.tma = 0;
  if (p.tma && !get_encode_h()) p.tma = 0;
  if (one_box && p.tma != 2) return false;
  p.xb = 8 * p.ks + (p.nkw > 1 ? 8 : 0);
  p.plane_bytes = p.tma == 2 ? p.hr * p.xb * g.cin * 2 : p.hr * p.rw * g.cin * 2;
  p.w_tap_bytes = g.cout * g.cin * 2;
  p.ntaps_w = maxw + 1;
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  p.red = (out_dtype == URN_BF16 && g.accumulate && g.os == 1 && ncls == 1 && !dual && (cout_store == 0 || cout_store == g.cout) &&
           (g.cout == 16 || g.cout == 32 || g.cout == 64) && g.ldout % 8 == 0 && get_encode_h() != nullptr &&
           !(getenv("URN_HALO_RED") && atoi(getenv("URN_HALO_RED")) == 0)) ? 1 : 0;
  p.red_bytes = p.red ? 4 * 2 * 32 * g.cout * 2 + 1024 : 0;
  const int budget = smem_limit - w_bytes - p.red_bytes;
  p.stages = std::min(max_stages(), budget / p.plane_bytes);
  if (p.stages < 2) return false;
  if (p.ntiles >= (1ll << 31)) return false;
  p.accumulate = g.accumulate; p.out_dtype = out_dtype; p.in_dtype = in_dtype;
  const int fmt = in_dtype == URN_BF16 ? 1 : 0;
  p.idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(g.cout >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  const uint64_t layout = pitch == 128 ? 2 : (pitch == 64 ? 4 : 6);
  // 8-row groups (consecutive output rows of the sub-tile) are is*RW staged rows apart
  p.adesc_hi = (layout << 61) | (1ull << 46) | ((((uint64_t)(g.is * p.rw) * pitch) >> 4) << 32) | (1ull << 16);
  // TMA staging: 8-row groups (consecutive output rows of a sub-tile) are one staged halo row apart (8*ks or xb
  // pixels); every box starts on a multiple of the swizzle pattern (256 / 512 / 1024 B for SW32 / 64 / 128)
  if (p.tma) p.adesc_hi = (layout << 61) | (1ull << 46) | ((((uint64_t)(p.tma == 2 ? p.xb : 8 * p.ks) * pitch) >> 4) << 32) | (1ull << 16);
  if (dual && !p.tma) return false;
  p.bdesc_hi = (layout << 61) | (1ull << 46) | (((8 * pitch) >> 4) << 32) | (1ull << 16);
  // concat mode: 3x3x3 stride-1 convs; the zero block is one B tile of N = TD*cout rows
  p.coff[0][0] = g.ood; p.coff[0][1] = g.ooh; p.coff[0][2] = g.oow;
  for (int i = 0; i < 27; ++i) p.ctapw[0][i] = p.tapw[i];
  p.concat = (ncls == 1 && p.nkd == 3 && (p.nkh == 3 || p.nkh == 1) && p.nkw == 3 && g.is == 1 && g.os == 1 && 3 * g.cout <= 256 && p.TD >= 2 &&
              !getenv("URN_HALO_NO_CONCAT")) ? 1 : 0;
  if (p.concat) {
    p.zero_bytes = ((p.TD * g.cout * (int)pitch + 1023) / 1024) * 1024;
    const uint32_t id0 = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(128 >> 4) << 24);
    for (int i = 0; i < 3; ++i) p.idesc_n[i] = id0 | ((uint32_t)(((i + 1) * g.cout) >> 3) << 17);
    p.idesc_zero = id0 | ((uint32_t)((p.TD * g.cout) >> 3) << 17);
    p.stages = std::min(max_stages(), (budget - p.zero_bytes) / p.plane_bytes);
    if (p.stages < 2) { p.concat = 0; p.zero_bytes = 0; p.stages = std::min(max_stages(), budget / p.plane_bytes); }
  }
  p.nacc = std::max(2, std::min(kMaxAcc, tmem_limit / (ncls * p.TD * p.ks * g.cout)));
  if (const char* e = getenv("URN_HALO_NACC")) p.nacc = std::max(2, std::min(p.nacc, atoi(e)));
  p.tmem_cols = std::max(32, pow2c(p.nacc * ncls * p.TD * p.ks * g.cout));
  if (p.tmem_cols > tmem_limit) return false;
  if (const char* e = getenv("URN_HALO_DBG")) p.dbg = atoi(e);
  // Cin = 64 with 1-4 taps: the tap-box kernel (128-byte TMA rows, few boxes) is as fast or faster
  if (g.cin == 64 && !(p.nkh == 3 && p.nkw == 3) && !getenv("URN_HALO_ALL")) return false;
  if (getenv("URN_HALO_ONLY3X3") && !(g.is == 1 && g.os == 1 && p.nkh == 3 && p.nkw == 3)) return false;
  return true;
}

// Images-as-rows plan (default for 2D stride-1 3x3 layers of a batch of >= 12 images; URN_HALO_ROWS=0
// disables it): the layer is run through the 3D concat path by relabelling the axes -- planes := image
// rows, tile rows := the IMAGES of the batch, columns := image columns.  The kh taps become the depth
// taps, which the concat mode fuses along N: per 128 outputs 3 MMAs of N = 3*cout instead of 9 of
// N = cout.  Every tcgen05.mma occupies the tensor pipe >= ~58 cycles whatever its N <= 96 and M
// (tools/probe/mmarate.cu), so a third of the instructions is a third of the pipe time.  Only the
// strides differ from a real volume: the next "row" is the next image, the next "plane" the next
// image row.  Measured per layer at 512^2 / 256^2, B = 32 (fp16 raw outputs): 16->16 forward 0.233 ->
// 0.207 ms, input gradient 0.218 -> 0.192 ms, 32->16 forward 0.427 -> 0.346 ms; slower only for
// ACCUMULATING input gradients of >= 32 channels (0.172 -> 0.190 ms), which keep the 16x8 tiles.
bool halo_plan_rows(const GConv& g, int in_dtype, int out_dtype, int cout_store, HaloParams* pp) {
  if (const char* e = getenv("URN_HALO_ROWS")) { if (atoi(e) == 0) return false; }
  if (g.n < 12 || (g.accumulate && g.cout >= 32)) return false;
  if (g.id != 1 || g.od != 1 || g.gd != 1 || g.is != 1 || g.os != 1 || g.ntaps != 9) return false;
  if (g.ood || g.ooh || g.oow || g.gh != g.oh || g.gw != g.ow || g.ih != g.oh || g.iw != g.ow) return false;
  if (cout_store > 0 && cout_store != g.cout && getenv("URN_HALO_ROWS_NO_HEAD")) return false;
  if (g.ih < 4 || (long long)g.n * g.ih * g.iw * g.ldin * 2 >= (1ll << 31)) return false;   // 32-bit offsets across the images of a plane
  GConv v = g;
  v.n = 1;
  v.id = g.ih; v.ih = g.n; v.iw = g.iw;
  v.od = g.oh; v.oh = g.n; v.ow = g.ow;
  v.gd = g.gh; v.gh = g.n; v.gw = g.gw;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.taps[t].od != 0) return false;
    v.taps[t].od = g.taps[t].oh; v.taps[t].oh = 0; v.taps[t].ow = g.taps[t].ow;
  }
  if (!halo_plan(v, in_dtype, out_dtype, cout_store, pp, 1)) return false;
  HaloParams& p = *pp;
  if (!p.concat) return false;
  p.in_plane_b = (long long)g.ldin * 2 * g.iw;              // next image row
  const long long img_b = p.in_plane_b * g.ih;              // next image
  if (img_b * 17 >= (1ll << 31)) return false;
  p.in_row_b = (int)img_b;
  p.in_img_b = 0;
  p.o_plane = g.ow; p.o_row = (long long)g.oh * g.ow; p.o_img = 0;
  return true;
}

// All parity classes of a stride-2 scatter in one launch: gs[c] are the per-class gather launches
// (gconv_scatter_class).  The union of their taps is a 2x2(x2) box of offsets {-1, 0}.
bool halo_plan_classes(const GConv* gs, int ncls, int in_dtype, int out_dtype, HaloParams* pp) {
  if (getenv("URN_HALO_NO_CLASSES") || (ncls != 4 && ncls != 8)) return false;
  const GConv& g0 = gs[0];
  GConv u = g0;                              // union launch: geometry of class 0, taps = the full box
  int lo[3] = {1 << 20, 1 << 20, 1 << 20}, hi[3] = {-(1 << 20), -(1 << 20), -(1 << 20)};
  int maxw = 0;
  for (int c = 0; c < ncls; ++c) {
    const GConv& g = gs[c];
    if (g.is != 1 || g.os != 2 || g.ntaps < 1) return false;
    if (g.n != g0.n || g.id != g0.id || g.ih != g0.ih || g.iw != g0.iw || g.cin != g0.cin || g.cout != g0.cout ||
        g.od != g0.od || g.oh != g0.oh || g.ow != g0.ow || g.gd != g0.gd || g.gh != g0.gh || g.gw != g0.gw ||
        g.ldin != g0.ldin || g.ldout != g0.ldout || g.accumulate != g0.accumulate)
      return false;
    for (int t = 0; t < g.ntaps; ++t) {
      const int o[3] = {g.taps[t].od, g.taps[t].oh, g.taps[t].ow};
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], o[k]); hi[k] = std::max(hi[k], o[k]); }
      maxw = std::max(maxw, g.taps[t].widx);
    }
  }
  for (int k = 0; k < 3; ++k)
    if (hi[k] - lo[k] > 1) return false;
  u.ntaps = 0;
  for (int a = lo[0]; a <= hi[0]; ++a)
    for (int b = lo[1]; b <= hi[1]; ++b)
      for (int c2 = lo[2]; c2 <= hi[2]; ++c2) {
        GTap& t = u.taps[u.ntaps];
        t.od = a; t.oh = b; t.ow = c2; t.widx = u.ntaps;
        ++u.ntaps;
      }
  if (!halo_plan(u, in_dtype, out_dtype, 0, pp, ncls)) return false;
  HaloParams& p = *pp;
  if (p.concat) return false;
  p.ntaps_w = maxw + 1;
  p.w_tap_bytes = g0.cout * g0.cin * 2;
  // smem budget again with the real number of resident taps
  {
    const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
    p.stages = std::min(max_stages(), ((p.dual ? 108 : 224) * 1024 - w_bytes) / p.plane_bytes);
    if (p.stages < 2) return false;
  }
  for (int c = 0; c < ncls; ++c) {
    const GConv& g = gs[c];
    p.coff[c][0] = g.ood; p.coff[c][1] = g.ooh; p.coff[c][2] = g.oow;
    for (int i = 0; i < 27; ++i) p.ctapw[c][i] = -1;
    bool has0 = false;
    for (int t = 0; t < g.ntaps; ++t) {
      const GTap& tp = g.taps[t];
      p.ctapw[c][(tp.od - lo[0]) * 9 + (tp.oh - lo[1]) * 3 + (tp.ow - lo[2])] = tp.widx;
      if (tp.od == 0 && tp.oh == 0 && tp.ow == 0) has0 = true;
    }
    if (!has0) return false;               // the always-inside tap that guarantees a written accumulator
  }
  return true;
}

}  // namespace

bool gconv_halo_classes_supported(const GConv* gs, int ncls, int in_dtype, int out_dtype) {
  HaloParams p;
  return halo_plan_classes(gs, ncls, in_dtype, out_dtype, &p);
}

static int halo_launch_params(HaloParams& p, int cin, int cout, int in_dtype, int out_dtype, const void* in,
                              const void* wpack, void* out, cudaStream_t st);

int gconv_halo_classes_launch(const GConv* gs, int ncls, int in_dtype, const void* in, const void* wpack, void* out,
                              int out_dtype, float* stats_part, int* nparts, cudaStream_t st) {
  HaloParams p;
  URN_CHECK(halo_plan_classes(gs, ncls, in_dtype, out_dtype, &p), "gconv_halo: unsupported class set");
  if (nparts) *nparts = 0;
  if (p.ntiles == 0) return 0;
  // the classes together write every output element exactly once: statistics can be fused
  if (stats_part && nparts && out_dtype != URN_BF16 && !gs[0].accumulate && !getenv("URN_NO_FUSED_STATS") &&
      (long long)gs[0].gd * 2 == gs[0].od + (gs[0].od == 1) && gs[0].gh * 2 == gs[0].oh && gs[0].gw * 2 == gs[0].ow) {
    p.stats_part = stats_part;
    *nparts = (int)std::min<long long>(p.ntiles, (long long)persistent_ctas() * (p.dual ? 2 : 1)) * 4;
  }
  return halo_launch_params(p, gs[0].cin, gs[0].cout, in_dtype, out_dtype, in, wpack, out, st);
}

bool gconv_halo_supported(const GConv& g, int in_dtype, int out_dtype) {
  HaloParams p;
  return halo_plan(g, in_dtype, out_dtype, 0, &p);
}
bool gconv_halo_supported_padded(const GConv& g, int in_dtype, int out_dtype, int cout_store) {
  HaloParams p;
  return halo_plan(g, in_dtype, out_dtype, cout_store, &p);
}

// wpack: 16-bit filter (dtype of `in`) laid out [tap][cout][cin].  cout_store (0 = all): only the
// first cout_store output channels are written (the filter was zero-padded to a multiple of 16).
int gconv_halo_launch_stats(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                            int out_dtype, int cout_store, float* stats_part, int* nparts, cudaStream_t st) {
  HaloParams p;
  if (!halo_plan_rows(g, in_dtype, out_dtype, cout_store, &p))
    URN_CHECK(halo_plan(g, in_dtype, out_dtype, cout_store, &p), "gconv_halo: unsupported shape/dtype");
  if (nparts) *nparts = 0;
  if (p.ntiles == 0) return 0;
  const int grid_ = (int)std::min<long long>(p.ntiles, (long long)persistent_ctas() * (p.dual ? 2 : 1));
  // statistics only when this single launch produces the whole raw tensor, freshly written
  if (stats_part && nparts && out_dtype != URN_BF16 && !g.accumulate && g.os == 1 && g.gd == g.od && g.gh == g.oh &&
      g.gw == g.ow && !getenv("URN_NO_FUSED_STATS")) {
    p.stats_part = stats_part;
    *nparts = grid_ * 4;
  }
  return halo_launch_params(p, g.cin, g.cout, in_dtype, out_dtype, in, wpack, out, st);
}

static int halo_launch_params(HaloParams& p, int cin, int cout, int in_dtype, int out_dtype, const void* in,
                              const void* wpack, void* out, cudaStream_t st) {
  EncodeTiledFnH enc = get_encode_h();
  URN_CHECK(enc, "gconv_halo: cuTensorMapEncodeTiled not available");
  CUtensorMap map_w;
  {
    const CUtensorMapDataType dt = in_dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw = cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)p.ntaps_w};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * 2 * cout};
    cuuint32_t box[3] = {(cuuint32_t)cin, (cuuint32_t)cout, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_w, dt, 3, const_cast<void*>(wpack), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_halo: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  CUtensorMap map_a = map_w;             // (unused unless p.tma)
  if (p.tma) {
    // input as a 5-D tensor (c, w, row, plane, image) with the plan's byte strides (the images-as-rows plan
    // relabels rows := images, planes := image rows); box = one kw-shifted run of 8*ks pixels x hr halo rows
    const CUtensorMapDataType dt = in_dtype == URN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapSwizzle sw = cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const cuuint64_t img_stride = p.in_img_b > 0 ? (cuuint64_t)p.in_img_b : (cuuint64_t)p.in_plane_b * (cuuint64_t)p.D;
    // outer dims (tile rows, planes, images) in ascending stride order (the images-as-rows plan has row stride >
    // plane stride); a dim of extent 1 goes last
    struct OD { cuuint64_t size, stride; cuuint32_t box; int what; };
    OD od[3] = {{(cuuint64_t)p.H, (cuuint64_t)p.in_row_b, (cuuint32_t)p.hr, 0}, {(cuuint64_t)p.D, (cuuint64_t)p.in_plane_b, 1u, 1},
                {(cuuint64_t)p.n, img_stride, 1u, 2}};
    std::stable_sort(od, od + 3, [](const OD& a, const OD& b) { return (a.size > 1 ? a.stride : ~0ull) < (b.size > 1 ? b.stride : ~0ull); });
    for (int i = 0; i < 3; ++i) p.tma_slot[od[i].what] = 2 + i;
    // strides of extent-1 dims only have to be legal: keep them ascending
    for (int i = 1; i < 3; ++i) if (od[i].size <= 1 && od[i].stride < od[i - 1].stride * od[i - 1].size) od[i].stride = od[i - 1].stride * od[i - 1].size;
    cuuint64_t dims[5] = {(cuuint64_t)cin, (cuuint64_t)p.W, od[0].size, od[1].size, od[2].size};
    cuuint64_t strides[4] = {(cuuint64_t)p.ldin * 2, od[0].stride, od[1].stride, od[2].stride};
    cuuint32_t box[5] = {(cuuint32_t)cin, (cuuint32_t)(p.tma == 2 ? p.xb : 8 * p.ks), od[0].box, od[1].box, od[2].box};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map_a, dt, 5, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_halo: cuTensorMapEncodeTiled(A) failed: %d (dims %d %d %d %d %d)", (int)r, cin, p.W, p.H, p.D, p.n);
  }
  const int w_bytes = ((p.ntaps_w * p.w_tap_bytes + 1023) / 1024) * 1024;
  CUtensorMap map_o = map_w;             // (unused unless p.red)
  if (p.red) {
    // output gradient as a 5-D bf16 tensor (c, w, row, plane, image) in the plan's strides; box = 8 pixels x 4 tile rows
    const cuuint64_t pix = (cuuint64_t)p.ldout * 2;
    const cuuint64_t img_stride = p.o_img > 0 ? (cuuint64_t)p.o_img * pix : (cuuint64_t)p.o_plane * pix * (cuuint64_t)p.OD;
    struct OD { cuuint64_t size, stride; cuuint32_t box; int what; };
    OD od[3] = {{(cuuint64_t)p.OH, (cuuint64_t)p.o_row * pix, 4u, 0}, {(cuuint64_t)p.OD, (cuuint64_t)p.o_plane * pix, 1u, 1},
                {(cuuint64_t)p.n, img_stride, 1u, 2}};
    std::stable_sort(od, od + 3, [](const OD& a, const OD& b) { return (a.size > 1 ? a.stride : ~0ull) < (b.size > 1 ? b.stride : ~0ull); });
    for (int i = 0; i < 3; ++i) p.red_slot[od[i].what] = 2 + i;
    for (int i = 1; i < 3; ++i) if (od[i].size <= 1 && od[i].stride < od[i - 1].stride * od[i - 1].size) od[i].stride = od[i - 1].stride * od[i - 1].size;
    const CUtensorMapSwizzle sw = p.cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.cout == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[5] = {(cuuint64_t)p.cout, (cuuint64_t)p.OW, od[0].size, od[1].size, od[2].size};
    cuuint64_t strides[4] = {pix, od[0].stride, od[1].stride, od[2].stride};
    cuuint32_t box[5] = {(cuuint32_t)p.cout, (cuuint32_t)kTW, od[0].box, od[1].box, od[2].box};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    URN_CHECK(r == CUDA_SUCCESS, "gconv_halo: cuTensorMapEncodeTiled(out) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)w_bytes + (size_t)p.zero_bytes + (size_t)p.stages * p.plane_bytes + 1024 + (size_t)p.red_bytes;
  const int grid = (int)std::min<long long>(p.ntiles, (long long)persistent_ctas() * (p.dual ? 2 : 1));
  const int nthreads = p.dual ? 192 : kHThreads;
  long long* tl = nullptr;
  if (const char* e = getenv("URN_HALO_TIMELINE")) tl = (long long*)strtoull(e, nullptr, 0);
  const int mode = (p.cout_store & 15) ? 2 : (p.accumulate ? 1 : 0);
  const bool full3 = p.nkh == 3 && p.nkw == 3;
  const bool stats = p.stats_part != nullptr;
#define HALO_GO(OK, MODE, F3, ST)                                                                             \
  do {                                                                                                        \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      URN_CUDA(cudaFuncSetAttribute(gconv_halo_kernel<OK, MODE, F3, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    227 * 1024));                                                             \
      attr_set = true;                                                                                        \
    }                                                                                                         \
    gconv_halo_kernel<OK, MODE, F3, ST><<<grid, nthreads, smem, st>>>(map_w, map_a, map_o, p, in, out, tl);          \
  } while (0)
  URN_CHECK(!stats || (out_dtype != URN_BF16 && mode != 1), "gconv_halo: statistics need a freshly written raw output");
  if (out_dtype == URN_F32) {
    if (mode == 0) {
      if (stats) { if (full3) HALO_GO(0, 0, true, true); else HALO_GO(0, 0, false, true); }
      else { if (full3) HALO_GO(0, 0, true, false); else HALO_GO(0, 0, false, false); }
    } else if (mode == 1) {
      if (full3) HALO_GO(0, 1, true, false); else HALO_GO(0, 1, false, false);
    } else {
      if (stats) { if (full3) HALO_GO(0, 2, true, true); else HALO_GO(0, 2, false, true); }
      else { if (full3) HALO_GO(0, 2, true, false); else HALO_GO(0, 2, false, false); }
    }
  } else if (out_dtype == URN_F16) {
    URN_CHECK(mode == 0, "gconv_halo: fp16 outputs are plain stores");
    if (stats) { if (full3) HALO_GO(2, 0, true, true); else HALO_GO(2, 0, false, true); }
    else { if (full3) HALO_GO(2, 0, true, false); else HALO_GO(2, 0, false, false); }
  } else {
    if (mode == 0) { if (full3) HALO_GO(1, 0, true, false); else HALO_GO(1, 0, false, false); }
    else if (mode == 1) { if (full3) HALO_GO(1, 1, true, false); else HALO_GO(1, 1, false, false); }
    else { if (full3) HALO_GO(1, 2, true, false); else HALO_GO(1, 2, false, false); }
  }
#undef HALO_GO
  URN_LAUNCH_CHECK();
  return 0;
}

int gconv_halo_launch_padded(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                             int out_dtype, int cout_store, cudaStream_t st) {
  return gconv_halo_launch_stats(g, in_dtype, in, wpack, out, out_dtype, cout_store, nullptr, nullptr, st);
}
int gconv_halo_launch(const GConv& g, int in_dtype, const void* in, const void* wpack, void* out,
                      int out_dtype, cudaStream_t st) {
  return gconv_halo_launch_padded(g, in_dtype, in, wpack, out, out_dtype, 0, st);
}
