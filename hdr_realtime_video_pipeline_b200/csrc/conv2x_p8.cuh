// Two chained 3x3 convolutions in one kernel: conv A (3x3, 32 outputs, ReLU, optional in-kernel SFT) feeds conv B
// (3x3 on those 32 channels) through a ring of rows in shared memory, so the 32-channel intermediate never reaches HBM.
//
//   conv_first -> SFT_layer1 -> HR_conv1          (HDRUNet3T1_arch.py:166-168)
//   HR_conv2 -> conv_last (+ img)                 (:198-205)
//   ResBlock_with_SFT: conv1 -> sft2 -> conv2 + x (arch_util.py:87-95)
//
// A CTA owns a strip of 126 output pixels and a band of output rows.  Conv A is evaluated on a 128-pixel tile that
// starts one pixel to the left (mid pixel e <-> image x = x0 - 1 + e) and on one extra row above and below the band;
// mid pixels / rows outside the image are forced to zero (they are conv B's zero padding).  Conv B then reads the mid
// ring exactly like an input ring: out pixel m <-> x = x0 + m reads mid entries m .. m+2; its tile rows 126 and 127 are
// discarded.  The two convs have their own MMA-issuing warps and their own epilogue warpgroups, so with one CTA per SM
// the tensor pipe still sees two independent instruction streams:
//
//   warp 0      TMA producer: input rows (+ the SFT stage-0 row of the same mid row, in the same ring slot)
//   warp 1      MMA issuer of conv A        warp 2   MMA issuer of conv B        warp 12  MMA issuer of the SFT stage-1 GEMM
//   warps 3-6   epilogue A: TMEM -> ReLU, SFT -> fp16 -> mid ring    warps 7-10  epilogue B: TMEM -> global
//   warp 11     TMA producer of conv B's residual rows (read from shared memory by epilogue B: no DRAM latency per row)
//
// Three issuing warps: one thread issues a tcgen05.mma every ~76 cycles at best and every mbarrier wait / commit /
// elect costs it another 50-110 cycles even when already satisfied (scripts/sync_probe.py), while an N = 96 MMA
// occupies the pipe for ~56 cycles: a single issuer cannot keep the pipe busy (measured: 1.3-2.2x slower).
//
// Row folding.  A tcgen05.mma with M = 128, K = 16 occupies the tensor pipe for max(~45, N/2) cycles, so a 32-output
// conv issued tap by tap (N = 32) runs the pipe at a third of its rate and every input row is multiplied three times.
// Here the three vertical taps are folded into N: input row q is multiplied ONCE by [W(dy=2) | W(dy=1) | W(dy=0)]
// (N = 96) and the three 32-column results accumulate into the TMEM blocks of output rows q-2, q-1 and q, neighbours in
// a ring of R accumulator blocks (a window that crosses the end of the ring is issued as two MMAs).  The block of the
// newest row is initialised by the bias step (accumulate off).  A row costs 1 + 6 MMAs instead of 19, every input row
// is consumed and released once, and the R - 3 spare blocks hide the commit -> epilogue -> release latency.
#pragma once
#include "conv_p8.cuh"

namespace hdrtv {

constexpr int kC2Threads = 32 * 13;
constexpr int kC2Strip = 126;          // output pixels per strip
constexpr int kC2MidRing = 5;          // mid rows in flight (each is consumed once by conv B)
constexpr int kC2MidSlot = 4 * kPlaneBytes;
constexpr int kC2InRing = 5;           // input (+ stage-0) rows in flight, each consumed once by conv A: covers the HBM latency
constexpr int kC2ResRing = 4;          // residual rows in flight
constexpr int kC2ResPlane = kTileM * 16;
constexpr int kC2OnesOff = 1024;       // barrier table below, constant "ones" operand (bias steps) here
constexpr int kC2Header = kC2OnesOff + kPlaneBytes;
static_assert(kC2Header % 128 == 0, "operand alignment");

struct Conv2xParams {
  // conv A input
  const uint4* in;
  long in_row_entries;
  uint32_t copy_src0, copy_src_stride;       // source entry of channel-chunk plane c: src0 + c * stride (before the row/x offset)
  int H, W, band;                            // image size (conv A, mid and conv B all share it); band unused (1-D grid)
  int strips;                                // 126-pixel strips per row; CTA b owns items [b*T/G, (b+1)*T/G) of the strip-major (strip, row) list
  const uint4* wpkA; int wA_bytes;           // conv A packed weights, row-folded (N = 96 per step) + bias step (N = 32)
  const uint4* wpk2; int w2_bytes;           // SFT stage-1 weights (SFTG)
  const uint4* s0; long s0_row_entries; uint32_t s0_src0, s0_wp;
  const uint4* wpkB; int wB_bytes;           // conv B packed weights, row-folded (N = 3 NB per step) + bias step (N = NB)
  int has_res, has_res2, has_raw;
  P8 res, res2, raw, out;                    // res: natural layout (streamed through shared memory)
  __half* planar; long planar_plane; int planar_W;
  // STORE_PLANAR only, optional: the RGB48 feeder pack fused into this epilogue (gui_pipeline_worker_feeders.py:193-249:
  // fp32 clamp * 65535 + 0.5, truncate, RGB order, HWC interleave) so that the one-call frame path needs no pack launch
  // INT8 layouts (Type B: f16 MMAs on de-quantised values, the reference's eager INT8 semantics): qmid = input quantiser of
  // conv B, applied to the mid rows; outq = uint8 copy of the output through `out_q` for an INT8 (kind::i8) consumer
  ActQuant qmid, out_q;
  int has_outq, skip_out;                    // skip_out: the fp16 `out` tensor has no reader
  P8 outq;                                   // uint8 tensor: 16 channels per 16-byte entry (chunks = C / 16)
  uint16_t* rgb48;                           // uint16 (H, W, 3), device memory
  const uint16_t* rgb48_lut;                 // optional transfer code table indexed by the half bit pattern of the clamped value
  unsigned long long* rgb48_cksum;           // optional descriptor checksum accumulator (kernels_io.cuh: cks_term)
  int* err;
  uint32_t opaque_zero;      // 0 at run time, unknown to ptxas: data-dependent slot release (ptx.cuh, mbar_arrive_after)
  long long* trace;          // only read when compiled with HDRTV_CHAIN_TRACE: clock64 stamps of CTA 0 [row<64][role<8][8]
};

// fp32 x8 -> fp16 x8 (round to nearest), optionally with the ReLU fused into the conversion
template <int ACT>
__device__ __forceinline__ uint4 pack8_act(const float* f) {
  uint4 u;
  uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (ACT == ACT_RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(f[2 * i + 1]), "f"(f[2 * i]));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(f[2 * i + 1]), "f"(f[2 * i]));
  }
  return u;
}
__device__ __forceinline__ void hadd2x4(uint4& a, const uint4& b) {
  __half2* x = reinterpret_cast<__half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hadd2(x[i], y[i]);
}

#ifdef HDRTV_CHAIN_TRACE
#define C2X_STAMP(role, row, k) do { if (p.trace && blockIdx.x == 0 && (row) < 64 && lane == 0) p.trace[(((row) * 8) + (role)) * 8 + (k)] = clock64(); } while (0)
#else
#define C2X_STAMP(role, row, k) do { } while (0)
#endif

// ACTB: activation of conv B (ACT_NONE / ACT_RELU).  Conv B's epilogue follows the reference's fp16 graph op by op: the
// conv result is rounded to fp16 (with the ReLU fused into the conversion), then the residuals are added in fp16.
template <int KINDA, int KCHA, bool SFTGA, int NB, int MODEB, int ACTB>
__global__ void __launch_bounds__(kC2Threads, 1) conv2x_p8_kernel(const __grid_constant__ Conv2xParams p) {
  static_assert(KINDA == IN_NAT3x3 || KINDA == IN_NAT3x3_C8, "conv A: 3x3 stride 1");
  static_assert(MODEB == STORE_P8 || MODEB == STORE_PLANAR, "conv B store mode");
  static_assert(ACTB == ACT_NONE || ACTB == ACT_RELU, "conv B activation");
  constexpr int NA = 32;
  constexpr int SPDA = kind_spd(KINDA, KCHA), NCOPY = kind_copies(KINDA, KCHA);
  constexpr int SPDB = kind_spd(IN_NAT3x3, 4);
  // accumulator ring blocks: 3 accumulating + spare ones being drained.  With the SFT GEMM's 128 columns TMEM holds 12
  // blocks of 32: the cheap 3-channel conv A gets by with one spare block, conv B gets the rest.
  constexpr int RA = !SFTGA ? 8 : (KINDA == IN_NAT3x3_C8 ? 4 : 6), RB = !SFTGA ? 8 : (KINDA == IN_NAT3x3_C8 ? 8 : 6);
  constexpr int CHR = MODEB == STORE_PLANAR ? 1 : NB / 8;  // residual planes per row
  constexpr int RES_SLOT = CHR * kC2ResPlane;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t colA = 0, colB = colA + RA * NA, colS = colB + RB * NB;
  static_assert(colS + (SFTGA ? 128 : 0) <= kTmemCols, "TMEM budget");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  auto in_full = [&](int i) { return bar(i); };
  auto in_empty = [&](int i) { return bar(8 + i); };
  auto mid_full = [&](int i) { return bar(32 + i); };
  auto mid_empty = [&](int i) { return bar(40 + i); };
  auto a_tfull = [&](int i) { return bar(48 + i); };
  auto a_tempty = [&](int i) { return bar(56 + i); };
  auto b_tfull = [&](int i) { return bar(64 + i); };
  auto b_tempty = [&](int i) { return bar(72 + i); };
  auto st_full = [&](int i) { return bar(80 + i); };       // scale|shift accumulator stage written
  auto st_empty = [&](int i) { return bar(82 + i); };      // ... and read back by epilogue A
  auto res_full = [&](int i) { return bar(84 + i); };
  auto res_empty = [&](int i) { return bar(92 + i); };
  const uint32_t wfull_bar = bar(100);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 102);
  static_assert(kC2InRing <= 8 && kC2MidRing <= 8 && kC2ResRing <= 8 && RA <= 8 && RB <= 8, "barrier table layout");
  uint8_t* ones = smem + kC2OnesOff;
  uint8_t* wsmA = smem + kC2Header;
  uint8_t* wsm2 = wsmA + ((p.wA_bytes + 127) & ~127);
  uint8_t* wsmB = wsm2 + (SFTGA ? ((p.w2_bytes + 127) & ~127) : 0);
  uint8_t* ring = wsmB + ((p.wB_bytes + 127) & ~127);
  constexpr int IN_SLOT = (NCOPY + (SFTGA ? 4 : 0)) * kPlaneBytes;      // input planes, then the four stage-0 planes
  uint8_t* mring = ring + kC2InRing * IN_SLOT;
  uint8_t* rring = mring + kC2MidRing * kC2MidSlot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // This CTA's contiguous range of (strip, output row) items, walked as segments that stay inside one strip.  Ring
  // slots, accumulator blocks and barrier phases simply continue from one segment to the next.
  struct Seg { int x0, oy0, n_out, jv0, jv1, n_mid_valid, n_in; };
  long w_lo, w_hi;
  {
    const long total = static_cast<long>(p.strips) * p.H;
    w_lo = total * blockIdx.x / gridDim.x;
    w_hi = total * (blockIdx.x + 1) / gridDim.x;
  }
  auto next_seg = [&](long& lo, Seg& sg) -> bool {
    if (lo >= w_hi) return false;
    const int strip = static_cast<int>(lo / p.H);
    sg.x0 = strip * kC2Strip;
    sg.oy0 = static_cast<int>(lo - static_cast<long>(strip) * p.H);
    sg.n_out = static_cast<int>(min(static_cast<long>(p.H - sg.oy0), w_hi - lo));
    lo += sg.n_out;
    // mid rows j = 0 .. n_out+1 <-> image rows oy0-1+j; the valid ones (inside the image) are jv0 .. jv1
    sg.jv0 = (sg.oy0 == 0) ? 1 : 0;
    sg.jv1 = (sg.oy0 + sg.n_out >= p.H) ? sg.n_out : sg.n_out + 1;
    sg.n_mid_valid = sg.jv1 - sg.jv0 + 1;
    sg.n_in = sg.n_mid_valid + 2;                     // input image rows oy0-2+jv0 .. : all inside [-1, H]
    return true;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kC2InRing; ++i) { mbar_init(in_full(i), 1); mbar_init(in_empty(i), SFTGA ? 2 : 1); }
    for (int i = 0; i < kC2MidRing; ++i) { mbar_init(mid_full(i), 4); mbar_init(mid_empty(i), 1); }
    for (int i = 0; i < kC2ResRing; ++i) { mbar_init(res_full(i), 1); mbar_init(res_empty(i), 4); }
    for (int i = 0; i < RA; ++i) { mbar_init(a_tfull(i), 1); mbar_init(a_tempty(i), 4); }
    for (int i = 0; i < RB; ++i) { mbar_init(b_tfull(i), 1); mbar_init(b_tempty(i), 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(st_full(i), 1); mbar_init(st_empty(i), 4); }
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kPlaneEntries) {
    reinterpret_cast<uint4*>(ones)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
  auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
  const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, p.wA_bytes + p.wB_bytes + (SFTGA ? p.w2_bytes : 0));
      bulk_g2s(smem_u32(wsmA), p.wpkA, p.wA_bytes, wfull_bar);
      bulk_g2s(smem_u32(wsmB), p.wpkB, p.wB_bytes, wfull_bar);
      if constexpr (SFTGA) bulk_g2s(smem_u32(wsm2), p.wpk2, p.w2_bytes, wfull_bar);
      grid_dep_wait();
      uint32_t slot = 0, ph = 1;
      const long sstride = p.copy_src_stride;
      long lo = w_lo;
      Seg sg;
      while (next_seg(lo, sg)) {
        const int x0 = sg.x0, oy0 = sg.oy0, jv0 = sg.jv0;
        // Row slots start at image pixel x0 - 2, i.e. tensor entry x0 - 1.  For the first strip that is one entry
        // before the plane: skip it (the pixel only feeds mid pixel x = -1, which is forced to zero).
        const uint32_t lead = (x0 == 0) ? 1u : 0u;
        const uint32_t row_bytes = kPlaneBytes - 16 * lead;
        // first input row: image row (oy0 - 1 + jv0) - 1; tensor row index = image row + 1
        const uint4* src = p.in + static_cast<long>(oy0 - 1 + jv0) * p.in_row_entries + static_cast<long>(p.copy_src0) +
                           (x0 - 1 + static_cast<int>(lead));
        // stage-0 row of mid row q (image row oy0 - 1 + jv0 + q) travels with input row q: mid pixel e <-> entry x0 + e
        const uint4* ssrc = SFTGA ? p.s0 + static_cast<long>(oy0 - 1 + jv0 + 1) * p.s0_row_entries + static_cast<long>(p.s0_src0) + x0
                                  : nullptr;
        for (int q = 0; q < sg.n_in; ++q) {
          C2X_STAMP(0, q, 0);
          mbar_wait(in_empty(slot), ph, p.err, 21);
          C2X_STAMP(0, q, 1);
          const bool with_s = SFTGA && q < sg.n_mid_valid;
          mbar_expect_tx(in_full(slot), NCOPY * row_bytes + (with_s ? 4 * kPlaneBytes : 0));
          const uint32_t dst = smem_u32(ring) + slot * IN_SLOT + 16 * lead;
#pragma unroll
          for (int c = 0; c < NCOPY; ++c) {
            unsigned long long a;
            asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(static_cast<uint32_t>(c * sstride)), "l"(src));
            bulk_g2s(dst + c * kPlaneBytes, reinterpret_cast<const void*>(a), row_bytes, in_full(slot));
          }
          src += p.in_row_entries;
          if constexpr (SFTGA) {
            if (with_s) {
              const uint32_t sdst = smem_u32(ring) + slot * IN_SLOT + NCOPY * kPlaneBytes;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                unsigned long long a;
                asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(c * p.s0_wp), "l"(ssrc));
                bulk_g2s(sdst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, in_full(slot));
              }
              ssrc += p.s0_row_entries;
            }
          }
          if (++slot == kC2InRing) { slot = 0; ph ^= 1; }
        }
      }
      grid_dep_launch();
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------------ TMA producer, residual rows of conv B
    if (lane == 0 && p.has_res) {
      grid_dep_wait();
      uint32_t slot = 0, ph = 1;
      const uint32_t wp = static_cast<uint32_t>(p.res.Wp);
      const long row_entries = p.res.row_entries();
      long lo = w_lo;
      Seg sg;
      while (next_seg(lo, sg)) {
        // out pixel m <-> tensor entry x0 + 1 + m; stay inside the plane on the last strip
        const uint32_t bytes = 16u * static_cast<uint32_t>(min(kTileM, p.res.Wp - (sg.x0 + 1)));
        const uint4* src = reinterpret_cast<const uint4*>(p.res.base) + static_cast<long>(sg.oy0 + 1) * row_entries + (sg.x0 + 1);
        for (int t = 0; t < sg.n_out; ++t) {
          mbar_wait(res_empty(slot), ph, p.err, 36);
          mbar_expect_tx(res_full(slot), CHR * bytes);
          const uint32_t dst = smem_u32(rring) + slot * RES_SLOT;
#pragma unroll
          for (int c = 0; c < CHR; ++c) {
            unsigned long long a;
            asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(c * wp), "l"(src));
            bulk_g2s(dst + c * kC2ResPlane, reinterpret_cast<const void*>(a), bytes, res_full(slot));
          }
          src += row_entries;
          if (++slot == kC2ResRing) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const bool lead = elect_one();         // the issuing lane of this warp, elected once
    // ------------------------------------------------------------------ MMA issuer, conv A (valid mid rows only)
    mbar_wait(wfull_bar, 0, p.err, 23);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(3 * NA) << 16, b_step = 3 * NA * 2;
    constexpr uint32_t a_lbo = (kind_a_lbo(KINDA) >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsmA) >> 4) | b_lbo;
    const uint64_t bias_desc = mkdesc(((smem_u32(wsmA) + SPDA * 3 * NA * 32) >> 4) | (static_cast<uint32_t>(NA) << 16));
    constexpr uint32_t slot16 = IN_SLOT >> 4;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    int slot = 0, ph = 0;
    int g0 = 0;                                        // valid mid rows of earlier segments: ring position / phase origin
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      const int nmv = sg.n_mid_valid;
      for (int q = 0; q < sg.n_in; ++q) {              // input row q feeds mid rows q-2 (dy 2), q-1 (dy 1), q (dy 0)
        const int lo_r = max(q - 2, 0), hi_r = min(q, nmv - 1);
        const int pos_new = (g0 + q) % RA;
        C2X_STAMP(1, g0 + q, 0);
        if (q < nmv) mbar_wait(a_tempty(pos_new), (((g0 + q) / RA) & 1) ^ 1, p.err, 24);   // block of mid row q drained
        C2X_STAMP(1, g0 + q, 1);
        mbar_wait(in_full(slot), ph, p.err, 25);
        tc_fence_after();
        C2X_STAMP(1, g0 + q, 2);
        const int pos_lo = (g0 + lo_r) % RA, nwin = hi_r - lo_r + 1;
        const int n1 = min(nwin, RA - pos_lo), n2 = nwin - n1;      // the window may cross the end of the ring
        const uint32_t d1 = tmem_base + colA + static_cast<uint32_t>(pos_lo) * NA, d2 = tmem_base + colA;
        const uint32_t idesc1 = make_idesc_f16_m128(static_cast<uint32_t>(NA * n1));
        const uint32_t idesc2 = make_idesc_f16_m128(static_cast<uint32_t>(NA * max(n2, 1)));
        const uint32_t b_lo1 = b_lo0 + static_cast<uint32_t>(NA * (2 - (q - lo_r))), b_lo2 = b_lo1 + static_cast<uint32_t>(NA * n1);
        const uint32_t a16 = ring16 + slot * slot16;
        if (lead) {
          if (q < nmv)                                   // bias step: initialises the accumulator of the newest row
            tc_mma_f16(tmem_base + colA + static_cast<uint32_t>(pos_new) * NA, ones_desc, bias_desc, make_idesc_f16_m128(NA), 0u);
          static_for<0, SPDA>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr uint32_t a_off16 = kind_a_off(KINDA, KCHA, i) >> 4;
            const uint64_t ad = mkdesc((a16 + a_off16) | a_lbo);
            tc_mma_f16(d1, ad, mkdesc(b_lo1 + i * b_step), idesc1, 1u);
            if (n2) tc_mma_f16(d2, ad, mkdesc(b_lo2 + i * b_step), idesc2, 1u);
          });
          tc_commit(in_empty(slot));                   // every input row is read exactly once
          if (q >= 2) tc_commit(a_tfull((g0 + q - 2) % RA));          // mid row q-2 is complete
        }
        __syncwarp();
        C2X_STAMP(1, g0 + q, 3);
        if (++slot == kC2InRing) { slot = 0; ph ^= 1; }
      }
      g0 += nmv;
    }
  } else if (warp == 12) {
    const bool lead = elect_one();         // the issuing lane of this warp, elected once
    // ------------------------------------------------------------------ MMA issuer, SFT stage-1 GEMM (scale | shift)
    if constexpr (SFTGA) {
      mbar_wait(wfull_bar, 0, p.err, 23);
      constexpr uint32_t idesc64 = make_idesc_f16_m128(64);
      const uint32_t sb16 = (smem_u32(wsm2) >> 4) | (64u << 16);
      const uint32_t sring16 = (smem_u32(ring) + NCOPY * kPlaneBytes) >> 4;
      int slot = 0, ph = 0, g = 0;
      long lo = w_lo;
      Seg sg;
      while (next_seg(lo, sg)) {
        for (int q = 0; q < sg.n_in; ++q) {            // the stage-0 row of mid row q shares the ring slot of input row q
          if (q < sg.n_mid_valid) {
            mbar_wait(st_empty(g & 1), ((g >> 1) & 1) ^ 1, p.err, 33);
            mbar_wait(in_full(slot), ph, p.err, 26);
            tc_fence_after();
            if (lead) {
              const uint32_t sa16 = (sring16 + slot * (IN_SLOT >> 4)) | ((kPlaneBytes >> 4) << 16);
              const uint32_t s_tmem = tmem_base + colS + (g & 1) * 64;
              tc_mma_f16(s_tmem, mkdesc(sa16), mkdesc(sb16), idesc64, 0u);
              tc_mma_f16(s_tmem, mkdesc(sa16 + ((2 * kPlaneBytes) >> 4)), mkdesc(sb16 + 128), idesc64, 1u);
              tc_mma_f16(s_tmem, ones_desc, mkdesc(sb16 + 256), idesc64, 1u);
              tc_commit(st_full(g & 1));
              tc_commit(in_empty(slot));
            }
            __syncwarp();
            ++g;
          } else {
            if (lane == 0) mbar_arrive(in_empty(slot));    // no stage-0 row in this slot
            __syncwarp();
          }
          if (++slot == kC2InRing) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    const bool lead = elect_one();         // the issuing lane of this warp, elected once
    // ------------------------------------------------------------------ MMA issuer, conv B (reads the mid ring)
    mbar_wait(wfull_bar, 0, p.err, 27);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(3 * NB) << 16, b_step = 3 * NB * 2;
    constexpr uint32_t a_lbo = (kPlaneBytes >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsmB) >> 4) | b_lbo;
    const uint64_t bias_desc = mkdesc(((smem_u32(wsmB) + SPDB * 3 * NB * 32) >> 4) | (static_cast<uint32_t>(NB) << 16));
    const uint32_t mring16 = smem_u32(mring) >> 4;
    int slot = 0, ph = 0;
    int g0 = 0;                                        // output rows of earlier segments
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      const int n_out = sg.n_out;
      for (int j = 0; j < n_out + 2; ++j) {            // mid row j feeds output rows j-2 (dy 2), j-1 (dy 1), j (dy 0)
        const int lo_t = max(j - 2, 0), hi_t = min(j, n_out - 1);
        const int pos_new = (g0 + j) % RB;
        C2X_STAMP(2, g0 + j, 0);
        if (j < n_out) mbar_wait(b_tempty(pos_new), (((g0 + j) / RB) & 1) ^ 1, p.err, 28);
        C2X_STAMP(2, g0 + j, 1);
        mbar_wait(mid_full(slot), ph, p.err, 29);
        tc_fence_after();
        C2X_STAMP(2, g0 + j, 2);
        const int pos_lo = (g0 + lo_t) % RB, nwin = hi_t - lo_t + 1;
        const int n1 = min(nwin, RB - pos_lo), n2 = nwin - n1;
        const uint32_t d1 = tmem_base + colB + static_cast<uint32_t>(pos_lo) * NB, d2 = tmem_base + colB;
        const uint32_t idesc1 = make_idesc_f16_m128(static_cast<uint32_t>(NB * n1));
        const uint32_t idesc2 = make_idesc_f16_m128(static_cast<uint32_t>(NB * max(n2, 1)));
        const uint32_t b_lo1 = b_lo0 + static_cast<uint32_t>(NB * (2 - (j - lo_t))), b_lo2 = b_lo1 + static_cast<uint32_t>(NB * n1);
        const uint32_t a16 = mring16 + slot * (kC2MidSlot >> 4);
        if (lead) {
          if (j < n_out)
            tc_mma_f16(tmem_base + colB + static_cast<uint32_t>(pos_new) * NB, ones_desc, bias_desc, make_idesc_f16_m128(NB), 0u);
          static_for<0, SPDB>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr uint32_t a_off16 = kind_a_off(IN_NAT3x3, 4, i) >> 4;
            const uint64_t ad = mkdesc((a16 + a_off16) | a_lbo);
            tc_mma_f16(d1, ad, mkdesc(b_lo1 + i * b_step), idesc1, 1u);
            if (n2) tc_mma_f16(d2, ad, mkdesc(b_lo2 + i * b_step), idesc2, 1u);
          });
          tc_commit(mid_empty(slot));
          if (j >= 2) tc_commit(b_tfull((g0 + j - 2) % RB));
        }
        __syncwarp();
        C2X_STAMP(2, g0 + j, 3);
        if (++slot == kC2MidRing) { slot = 0; ph ^= 1; }
      }
      g0 += n_out;
    }
  } else if (warp < 7) {
    // ------------------------------------------------------------------ epilogue A: accumulator -> mid ring row
    const int lg = warp & 3;
    const int e = lg * 32 + lane;                      // mid pixel index = TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    int g = 0;                                         // valid mid rows seen (all segments)
    int slot = 0, ph = 1;                              // mid ring: wait for "empty" with the producer-side parity
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      const int x = sg.x0 - 1 + e;
      const bool inside_x = x >= 0 && x < p.W;
      for (int j = 0; j <= sg.n_out + 1; ++j) {
        const bool valid = j >= sg.jv0 && j <= sg.jv1;
        uint4 h[4];
        if (valid) {
          const int pos = g % RA, stage = g & 1;
          float sv[SFTGA ? 32 : 1], tv[SFTGA ? 32 : 1];
          if (warp == 3) C2X_STAMP(3, g, 0);
          if constexpr (SFTGA) {                       // scale|shift first: they were issued two rows ago
            mbar_wait(st_full(stage), (g >> 1) & 1, p.err, 37);
            if (warp == 3) C2X_STAMP(3, g, 1);
            tc_fence_after();
            tmem_ld32_async(tlane + colS + stage * 64, reinterpret_cast<uint32_t*>(sv));
            tmem_ld32_async(tlane + colS + stage * 64 + 32, reinterpret_cast<uint32_t*>(tv));
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(st_empty(stage));
          }
          if (warp == 3) C2X_STAMP(3, g, 2);
          mbar_wait(a_tfull(pos), (g / RA) & 1, p.err, 30);
          tc_fence_after();
          if (warp == 3) C2X_STAMP(3, g, 3);
          float v[32];
          tmem_ld32_async(tlane + colA + pos * NA, reinterpret_cast<uint32_t*>(v));
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_tempty(pos));
          if (warp == 3) C2X_STAMP(3, g, 4);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              a[k] = fmaxf(v[c * 8 + k], 0.f);                                                  // ReLU
              if constexpr (SFTGA) a[k] = fmaf(a[k], sv[c * 8 + k], tv[c * 8 + k]);           // x*(scale+1)+shift, +1 in the bias step
            }
            h[c] = inside_x ? pack8(a) : make_uint4(0, 0, 0, 0);
            if (p.qmid.mode && inside_x) h[c] = fq_entry(h[c], p.qmid);                         // conv B is a W8A8 layer (INT8 layouts)
          }
          ++g;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) h[c] = make_uint4(0, 0, 0, 0);                            // row outside the image
        }
        if (warp == 3 && valid) C2X_STAMP(3, g - 1, 5);
        mbar_wait(mid_empty(slot), ph, p.err, 31);
        if (warp == 3 && valid) C2X_STAMP(3, g - 1, 6);
        uint4* dst = reinterpret_cast<uint4*>(mring + slot * kC2MidSlot) + e;
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c * kPlaneEntries] = h[c];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(mid_full(slot));
        if (warp == 3 && valid) C2X_STAMP(3, g - 1, 7);
        if (++slot == kC2MidRing) { slot = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue B: accumulator -> global
    grid_dep_wait();
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    int g = 0;                                         // output rows seen (all segments)
    int rslot = 0, rph = 0;                            // residual ring
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      const int x = sg.x0 + m;
      const int oy0 = sg.oy0;
      const bool xin = m < kC2Strip && x < p.W;
      ColRef out, res2, raw, outq;
      out.init(p.out, x);
      if (p.has_outq) outq.init(p.outq, x);
      if (p.has_res2) res2.init(p.res2, x);
      if (p.has_raw) raw.init(p.raw, x);
      for (int t = 0; t < sg.n_out; ++t, ++g) {
        const int pos = g % RB, oy = oy0 + t;
        constexpr int CH = MODEB == STORE_PLANAR ? 1 : NB / 8;
        uint4 r4[CH], q4[CH];
        if (MODEB == STORE_P8 && p.has_res2 && xin) {
#pragma unroll
          for (int c = 0; c < CH; ++c) q4[c] = __ldcg(res2.at(oy, c));
        }
        if (warp == 7) C2X_STAMP(4, g, 0);
        if (p.has_res) {
          mbar_wait(res_full(rslot), rph, p.err, 38);
          const uint4* rs = reinterpret_cast<const uint4*>(rring + rslot * RES_SLOT) + m;
#pragma unroll
          for (int c = 0; c < CH; ++c) r4[c] = rs[c * kTileM];
          // Release the slot only once the loads have RETURNED: the barrier address below depends on the loaded
          // registers (a warp-wide ld.shared completes for all lanes at once).  With a plain arrive the slot could be
          // refilled by the residual producer's TMA while the ld.shared was still queued behind other shared-memory
          // traffic: rare rows then carried the residual of row t + 4 (found by scripts/stress_determinism.py when
          // other kernels share the SM).
          uint32_t dep = 0;
#pragma unroll
          for (int c = 0; c < CH; ++c) dep |= r4[c].x | r4[c].w;
          __syncwarp();
          if (lane == 0) mbar_arrive_after(res_empty(rslot), dep, p.opaque_zero);
          if (++rslot == kC2ResRing) { rslot = 0; rph ^= 1; }
        } else {
#pragma unroll
          for (int c = 0; c < CH; ++c) r4[c] = make_uint4(0, 0, 0, 0);
        }
        if (warp == 7) C2X_STAMP(4, g, 1);
        mbar_wait(b_tfull(pos), (g / RB) & 1, p.err, 32);
        tc_fence_after();
        if (warp == 7) C2X_STAMP(4, g, 2);
        float v[MODEB == STORE_PLANAR ? 8 : NB];
        tmem_ld_cols<(MODEB == STORE_PLANAR ? 8 : NB)>(tlane + colB + pos * NB, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_tempty(pos));
        if (warp == 7) C2X_STAMP(4, g, 3);
        if constexpr (MODEB == STORE_PLANAR) {
          unsigned long long cks = 0;
          if (xin) {
            uint4 o = pack8_act<ACTB>(v);
            hadd2x4(o, r4[0]);
            const __half* oh = reinterpret_cast<const __half*>(&o);
#pragma unroll
            for (int k = 0; k < 3; ++k) p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = oh[k];
            if (p.rgb48) {                                   // fused feeder pack: same arithmetic as pack_rgb48_kernel
              const long e0 = (static_cast<long>(oy) * p.planar_W + x) * 3;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                float f = fminf(fmaxf(__half2float(oh[k]), 0.f), 1.f);
                uint32_t code;
                if (p.rgb48_lut) code = p.rgb48_lut[__half_as_ushort(__float2half_rn(f))];
                else code = static_cast<uint32_t>(__fadd_rn(__fmul_rn(f, 65535.0f), 0.5f));
                p.rgb48[e0 + k] = static_cast<uint16_t>(code);
                cks += static_cast<unsigned long long>(code) * static_cast<unsigned long long>(static_cast<uint32_t>((e0 + k) % 65521) + 1u);
              }
            }
            if (p.has_raw) {
              o.y &= 0x0000FFFFu; o.z = 0u; o.w = 0u;                                            // channels 3..7 stay zero
              *raw.at(oy, 0) = o;
            }
          }
          if (p.rgb48_cksum) {                               // warp-uniform branch
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) cks += __shfl_xor_sync(0xffffffffu, cks, off);
            if (lane == 0 && cks) atomicAdd(p.rgb48_cksum, cks);
          }
        } else {
          if (xin) {
            uint4 o[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              o[c] = pack8_act<ACTB>(v + c * 8);
              hadd2x4(o[c], r4[c]);
            }
            if (p.has_res2) {
#pragma unroll
              for (int c = 0; c < CH; ++c) hadd2x4(o[c], q4[c]);
            }
            if (p.has_raw) {
#pragma unroll
              for (int c = 0; c < CH; ++c) *raw.at(oy, c) = o[c];
            }
            if (!p.skip_out) {
#pragma unroll
              for (int c = 0; c < CH; ++c) *out.at(oy, c) = o[c];
            }
            if constexpr (CH >= 2) {
              if (p.has_outq) {                    // uint8 codes of the (fp16) output for an INT8 consumer
#pragma unroll
                for (int c = 0; c < CH; c += 2) {
                  float f0[8], f1[8];
                  unpack8(o[c], f0);
                  unpack8(o[c + 1], f1);
                  *outq.at(oy, c >> 1) = pack16_u8h(f0, f1, p.out_q);
                }
              }
            }
          }
        }
        if (warp == 7) C2X_STAMP(4, g, 4);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int KINDA, int KCHA, bool SFTGA, int NB, int MODEB>
inline size_t conv2x_smem_bytes(const Conv2xParams& p) {
  constexpr int CHR = MODEB == STORE_PLANAR ? 1 : NB / 8;
  return kC2Header + ((p.wA_bytes + 127) & ~127) + (SFTGA ? ((p.w2_bytes + 127) & ~127) : 0) + ((p.wB_bytes + 127) & ~127) +
         static_cast<size_t>(kC2InRing) * (kind_copies(KINDA, KCHA) + (SFTGA ? 4 : 0)) * kPlaneBytes +
         static_cast<size_t>(kC2MidRing) * kC2MidSlot + static_cast<size_t>(kC2ResRing) * CHR * kC2ResPlane;
}

}  // namespace hdrtv
