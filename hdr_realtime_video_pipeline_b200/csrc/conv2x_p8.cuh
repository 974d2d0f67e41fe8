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
//   warp 0      TMA producer: input rows (and the SFT stage-0 rows of conv A's output rows)
//   warp 1      MMA issuer of conv A (+ the SFT stage-1 GEMM)       warp 2      MMA issuer of conv B
//   warps 3-6   epilogue A: TMEM -> ReLU, SFT -> fp16 -> mid ring    warps 7-10  epilogue B: TMEM -> global
#pragma once
#include "conv_p8.cuh"

namespace hdrtv {

constexpr int kC2Threads = 32 * 11;
constexpr int kC2Strip = 126;          // output pixels per strip
constexpr int kC2MidRing = 6;          // mid rows in flight (3 in use by conv B + 3 so that conv A runs ahead)
constexpr int kC2MidSlot = 4 * kPlaneBytes;
constexpr int kC2InRing = 6;           // input rows in flight: 3 in use + prefetch (HBM latency ~ 2 row periods)
constexpr int kC2SRing = 6;            // SFT stage-0 rows in flight

struct Conv2xParams {
  // conv A input
  const uint4* in;
  long in_row_entries;
  uint32_t copy_src0, copy_src_stride;       // source entry of channel-chunk plane c: src0 + c * stride (before the row/x offset)
  int H, W, band;                            // image size (conv A, mid and conv B all share it); band unused (1-D grid)
  int strips;                                // 126-pixel strips per row; CTA b owns items [b*T/G, (b+1)*T/G) of the strip-major (strip, row) list
  const uint4* wpkA; int wA_bytes;           // conv A packed weights (tap steps + bias step), N = 32
  const uint4* wpk2; int w2_bytes;           // SFT stage-1 weights (SFTG)
  const uint4* s0; long s0_row_entries; uint32_t s0_src0, s0_wp;
  const uint4* wpkB; int wB_bytes;           // conv B packed weights, N = NB
  float slopeB;                              // conv B activation as max(v, slope*v)
  int has_res, has_res2, has_raw;
  P8 res, res2, raw, out;
  __half* planar; long planar_plane; int planar_W;
  int* err;
};

template <int KINDA, int KCHA, bool SFTGA, int NB, int MODEB>
__global__ void __launch_bounds__(kC2Threads, 1) conv2x_p8_kernel(const __grid_constant__ Conv2xParams p) {
  static_assert(KINDA == IN_NAT3x3 || KINDA == IN_NAT3x3_C8, "conv A: 3x3 stride 1");
  static_assert(MODEB == STORE_P8 || MODEB == STORE_PLANAR, "conv B store mode");
  constexpr int NA = 32;
  constexpr int SPDA = kind_spd(KINDA, KCHA), NCOPY = kind_copies(KINDA, KCHA);
  constexpr int SPDB = kind_spd(IN_NAT3x3, 4);
  constexpr uint32_t kTmemCols = 256;                 // A 2 x 32 | B 2 x 32 | scale|shift 2 x 64
  constexpr uint32_t colA = 0, colB = 64, colS = 128;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  auto in_full = [&](int i) { return bar(i); };
  auto in_empty = [&](int i) { return bar(8 + i); };
  auto s_full = [&](int i) { return bar(16 + i); };
  auto s_empty = [&](int i) { return bar(24 + i); };
  auto mid_full = [&](int i) { return bar(32 + i); };
  auto mid_empty = [&](int i) { return bar(40 + i); };
  auto a_tfull = [&](int i) { return bar(48 + i); };
  auto a_tempty = [&](int i) { return bar(50 + i); };
  auto b_tfull = [&](int i) { return bar(52 + i); };
  auto b_tempty = [&](int i) { return bar(54 + i); };
  const uint32_t wfull_bar = bar(56);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 57);
  static_assert(kC2InRing <= 8 && kC2SRing <= 8 && kC2MidRing <= 8, "barrier table layout");
  uint8_t* ones = smem + 512;
  uint8_t* wsmA = smem + kSmemHeader;
  uint8_t* wsm2 = wsmA + ((p.wA_bytes + 127) & ~127);
  uint8_t* wsmB = wsm2 + (SFTGA ? ((p.w2_bytes + 127) & ~127) : 0);
  uint8_t* ring = wsmB + ((p.wB_bytes + 127) & ~127);
  uint8_t* sring = ring + kC2InRing * (NCOPY * kPlaneBytes);
  uint8_t* mring = sring + (SFTGA ? kC2SRing * kSSlotBytes : 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // This CTA's contiguous range of (strip, output row) items, walked as segments that stay inside one strip.  Ring
  // slots, stage parities and barrier phases simply continue from one segment to the next.
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
    for (int i = 0; i < kC2InRing; ++i) { mbar_init(in_full(i), 1); mbar_init(in_empty(i), 1); }
    for (int i = 0; i < kC2SRing; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 1); }
    for (int i = 0; i < kC2MidRing; ++i) { mbar_init(mid_full(i), 4); mbar_init(mid_empty(i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_tfull(i), 1); mbar_init(a_tempty(i), 4);
      mbar_init(b_tfull(i), 1); mbar_init(b_tempty(i), 4);
    }
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
      uint32_t slot = 0, ph = 1, sslot = 0, sph = 1;
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
        const uint4* ssrc = SFTGA ? p.s0 + static_cast<long>(oy0 - 1 + jv0 + 1) * p.s0_row_entries + static_cast<long>(p.s0_src0) + x0
                                  : nullptr;
        int ts = 0;
        for (int q = 0; q < sg.n_in; ++q) {
          mbar_wait(in_empty(slot), ph, p.err, 21);
          mbar_expect_tx(in_full(slot), NCOPY * row_bytes);
          const uint32_t dst = smem_u32(ring) + slot * (NCOPY * kPlaneBytes) + 16 * lead;
#pragma unroll
          for (int c = 0; c < NCOPY; ++c) {
            unsigned long long a;
            asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(static_cast<uint32_t>(c * sstride)), "l"(src));
            bulk_g2s(dst + c * kPlaneBytes, reinterpret_cast<const void*>(a), row_bytes, in_full(slot));
          }
          src += p.in_row_entries;
          if (++slot == kC2InRing) { slot = 0; ph ^= 1; }
          if constexpr (SFTGA) {
            while (ts < sg.n_mid_valid && ts + 2 <= q) {   // stage-0 row of the mid row whose last input row was just requested
              mbar_wait(s_empty(sslot), sph, p.err, 22);
              mbar_expect_tx(s_full(sslot), kSSlotBytes);
              const uint32_t sdst = smem_u32(sring) + sslot * kSSlotBytes;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                unsigned long long a;
                asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(c * p.s0_wp), "l"(ssrc));
                bulk_g2s(sdst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, s_full(sslot));
              }
              ssrc += p.s0_row_entries;
              ++ts;
              if (++sslot == kC2SRing) { sslot = 0; sph ^= 1; }
            }
          }
        }
      }
      grid_dep_launch();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer, conv A (valid mid rows only)
    mbar_wait(wfull_bar, 0, p.err, 23);
    constexpr uint32_t idesc = make_idesc_f16_m128(NA);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(NA) << 16, b_step = NA * 2;
    constexpr uint32_t a_lbo = (kind_a_lbo(KINDA) >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsmA) >> 4) | b_lbo;
    constexpr uint32_t slot16 = (NCOPY * kPlaneBytes) >> 4;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    int base_slot = 0, base_ph = 0, sslot = 0, sph = 0;
    int rg = 0;                                        // valid mid rows issued so far (all segments): TMEM stage parity
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      int waited = -1;
      for (int r = 0; r < sg.n_mid_valid; ++r, ++rg) {
        const int stage = rg & 1;
        mbar_wait(a_tempty(stage), ((rg >> 1) & 1) ^ 1, p.err, 24);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + colA + stage * NA;
        int slot = base_slot, ph = base_ph;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int q = r + dy;
          if (q > waited) {
            mbar_wait(in_full(slot), ph, p.err, 25);
            waited = q;
            tc_fence_after();
          }
          const uint32_t a16 = ring16 + slot * slot16;
          if (elect_one()) {
            static_for<0, SPDA>([&](auto ic) {
              constexpr int i = decltype(ic)::value;
              constexpr uint32_t a_off16 = kind_a_off(KINDA, KCHA, i) >> 4;
              tc_mma_f16(d_tmem, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (dy * SPDA + i) * b_step), idesc, (dy | i) ? 1u : 0u);
            });
            // input row r is not needed by later mid rows; the last mid row of a segment also frees the two rows below it
            if (dy == 0 || r == sg.n_mid_valid - 1) tc_commit(in_empty(slot));
          }
          __syncwarp();
          if (++slot == kC2InRing) { slot = 0; ph ^= 1; }
        }
        if constexpr (SFTGA) {
          mbar_wait(s_full(sslot), sph, p.err, 26);
          tc_fence_after();
        }
        if (elect_one()) {
          if constexpr (SFTGA) {
            constexpr uint32_t idesc64 = make_idesc_f16_m128(64);
            // mid pixel e <-> stage-0 entry x0 + e: the slot starts at entry x0, operand offset 0
            const uint32_t sa = ((smem_u32(sring) + sslot * kSSlotBytes) >> 4) | ((kPlaneBytes >> 4) << 16);
            const uint32_t sb = (smem_u32(wsm2) >> 4) | (64u << 16);
            const uint32_t s_tmem = tmem_base + colS + stage * 64;
            tc_mma_f16(s_tmem, mkdesc(sa), mkdesc(sb), idesc64, 0u);
            tc_mma_f16(s_tmem, mkdesc(sa + ((2 * kPlaneBytes) >> 4)), mkdesc(sb + 128), idesc64, 1u);
            tc_mma_f16(s_tmem, ones_desc, mkdesc(sb + 256), idesc64, 1u);
            tc_commit(s_empty(sslot));
          }
          tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + (3 * SPDA) * b_step), idesc, 1u);      // conv A bias
          tc_commit(a_tfull(stage));
        }
        __syncwarp();
        if constexpr (SFTGA) { if (++sslot == kC2SRing) { sslot = 0; sph ^= 1; } }
        if (++base_slot == kC2InRing) { base_slot = 0; base_ph ^= 1; }
      }
      // the segment used n_in = n_mid_valid + 2 ring slots: step over the two trailing ones
      for (int k = 0; k < 2; ++k)
        if (++base_slot == kC2InRing) { base_slot = 0; base_ph ^= 1; }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer, conv B (reads the mid ring)
    mbar_wait(wfull_bar, 0, p.err, 27);
    constexpr uint32_t idesc = make_idesc_f16_m128(NB);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(NB) << 16, b_step = NB * 2;
    constexpr uint32_t a_lbo = (kPlaneBytes >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsmB) >> 4) | b_lbo;
    const uint32_t mring16 = smem_u32(mring) >> 4;
    int base_slot = 0, base_ph = 0;
    int tg = 0;                                        // output rows issued so far (all segments): TMEM stage parity
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      int waited = -1;
      for (int t = 0; t < sg.n_out; ++t, ++tg) {
        const int stage = tg & 1;
        mbar_wait(b_tempty(stage), ((tg >> 1) & 1) ^ 1, p.err, 28);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + colB + stage * 32;
        int slot = base_slot, ph = base_ph;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int j = t + dy;                          // mid row
          if (j > waited) {
            mbar_wait(mid_full(slot), ph, p.err, 29);
            waited = j;
            tc_fence_after();
          }
          const uint32_t a16 = mring16 + slot * (kC2MidSlot >> 4);
          if (elect_one()) {
            static_for<0, SPDB>([&](auto ic) {
              constexpr int i = decltype(ic)::value;
              constexpr uint32_t a_off16 = kind_a_off(IN_NAT3x3, 4, i) >> 4;
              tc_mma_f16(d_tmem, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (dy * SPDB + i) * b_step), idesc, (dy | i) ? 1u : 0u);
            });
            // mid row t is not needed by later output rows; the last output row of a segment frees the two rows below it
            if (dy == 0 || t == sg.n_out - 1) tc_commit(mid_empty(slot));
            if (dy == 2) {
              tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + (3 * SPDB) * b_step), idesc, 1u);      // conv B bias
              tc_commit(b_tfull(stage));
            }
          }
          __syncwarp();
          if (++slot == kC2MidRing) { slot = 0; ph ^= 1; }
        }
        if (++base_slot == kC2MidRing) { base_slot = 0; base_ph ^= 1; }
      }
      for (int k = 0; k < 2; ++k)                      // the segment used n_out + 2 mid slots
        if (++base_slot == kC2MidRing) { base_slot = 0; base_ph ^= 1; }
    }
  } else if (warp < 7) {
    // ------------------------------------------------------------------ epilogue A: accumulator -> mid ring row
    const int lg = warp & 3;
    const int e = lg * 32 + lane;                      // mid pixel index = TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    int r = 0;                                         // valid mid rows seen (all segments)
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
          const int stage = r & 1;
          mbar_wait(a_tfull(stage), (r >> 1) & 1, p.err, 30);
          tc_fence_after();
          float v[32];
          float sv[SFTGA ? 32 : 1], tv[SFTGA ? 32 : 1];
          tmem_ld32_async(tlane + colA + stage * NA, reinterpret_cast<uint32_t*>(v));
          if constexpr (SFTGA) {
            tmem_ld32_async(tlane + colS + stage * 64, reinterpret_cast<uint32_t*>(sv));
            tmem_ld32_async(tlane + colS + stage * 64 + 32, reinterpret_cast<uint32_t*>(tv));
          }
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_tempty(stage));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float val = fmaxf(v[c * 8 + k], 0.f);                                             // ReLU
              if constexpr (SFTGA) val = fmaf(val, sv[c * 8 + k], val) + tv[c * 8 + k];       // x*(scale+1)+shift
              a[k] = inside_x ? val : 0.f;
            }
            h[c] = pack8(a);
          }
          ++r;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) h[c] = make_uint4(0, 0, 0, 0);                            // row outside the image
        }
        mbar_wait(mid_empty(slot), ph, p.err, 31);
        uint4* dst = reinterpret_cast<uint4*>(mring + slot * kC2MidSlot) + e;
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c * kPlaneEntries] = h[c];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(mid_full(slot));
        if (++slot == kC2MidRing) { slot = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue B: accumulator -> global
    grid_dep_wait();
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const float slope = p.slopeB;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    int tg = 0;                                        // output rows seen (all segments)
    long lo = w_lo;
    Seg sg;
    while (next_seg(lo, sg)) {
      const int x = sg.x0 + m;
      const int oy0 = sg.oy0;
      const bool xin = m < kC2Strip && x < p.W;
      ColRef out, res, res2, raw;
      out.init(p.out, x);
      if (p.has_res) res.init(p.res, x);
      if (p.has_res2) res2.init(p.res2, x);
      if (p.has_raw) raw.init(p.raw, x);
      for (int t = 0; t < sg.n_out; ++t, ++tg) {
        const int stage = tg & 1, oy = oy0 + t;
        if constexpr (MODEB == STORE_PLANAR) {
          uint4 r4 = make_uint4(0, 0, 0, 0);
          if (xin && p.has_res) r4 = *res.at(oy, 0);
          mbar_wait(b_tfull(stage), (tg >> 1) & 1, p.err, 32);
          tc_fence_after();
          float v[8];
          tmem_ld_cols<8>(tlane + colB + stage * 32, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_tempty(stage));
          if (xin) {
            float val[8], rr[8];
            unpack8(r4, rr);
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = (k < 3) ? fmaxf(v[k], slope * v[k]) + rr[k] : 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k)
              p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = __float2half_rn(val[k]);
            if (p.has_raw) *raw.at(oy, 0) = pack8(val);
          }
        } else {
          constexpr int CH = NB / 8;
          uint4 r4[CH], q4[CH];
          if (xin) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              if (p.has_res) r4[c] = *res.at(oy, c);
              if (p.has_res2) q4[c] = *res2.at(oy, c);
            }
          }
          mbar_wait(b_tfull(stage), (tg >> 1) & 1, p.err, 32);
          tc_fence_after();
          float v[NB];
          tmem_ld_cols<NB>(tlane + colB + stage * 32, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_tempty(stage));
          if (xin) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              float val[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) val[k] = fmaxf(v[c * 8 + k], slope * v[c * 8 + k]);
              if (p.has_res) {
                float rr[8];
                unpack8(r4[c], rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] += rr[k];
              }
              if (p.has_res2) {
                float rr[8];
                unpack8(q4[c], rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] += rr[k];
              }
              if (p.has_raw) *raw.at(oy, c) = pack8(val);
              *out.at(oy, c) = pack8(val);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int KINDA, int KCHA, bool SFTGA>
inline size_t conv2x_smem_bytes(const Conv2xParams& p) {
  return kSmemHeader + ((p.wA_bytes + 127) & ~127) + (SFTGA ? ((p.w2_bytes + 127) & ~127) : 0) + ((p.wB_bytes + 127) & ~127) +
         static_cast<size_t>(kC2InRing) * kind_copies(KINDA, KCHA) * kPlaneBytes + (SFTGA ? kC2SRing * kSSlotBytes : 0) +
         static_cast<size_t>(kC2MidRing) * kC2MidSlot;
}

}  // namespace hdrtv
