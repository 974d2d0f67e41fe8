// Fused per-pixel layer chains on tcgen05: a conv (3x3 / 1x1 from the TMA row ring) followed by 1x1 layers whose
// 64-channel intermediates never leave the SM.
//
//   ProgAGCM : 1x1 3->64 ReLU, 1x1 64->64 ReLU, 1x1 64->3                   (Condition_arch.py:571-583, GFM folded)
//   ProgCond : 3x3 3->64, 1x1, 1x1 [store cond], 1x1, 1x1, 1x1 ->16 [store cond1], LeakyReLU(0.1)
//                                                                          (HDRUNet3T1_arch.py:41-46, 160-161)
//   ProgCondSft : ProgCond + the stage-0 convs of the full-resolution SFT layers on cond1 (arch_util.py:63-72)
//
// Measured on B200 (scripts/tensor_probe.py, scripts/chain_trace.py, profiles/): one tcgen05.mma of K = 16 occupies
// the SM's tensor pipe for >= ~45 cycles whatever N <= 64 is, the issuing thread cannot run ahead of the pipe, an
// mbarrier hand-off between warps costs 200-1000 cycles, and every dynamically indexed parameter load in the issuing
// warp is a long-scoreboard stall with the pipe idle.  A chain step (5 MMAs, ~240 pipe cycles) is therefore latency-
// bound.  The kernel (a) is specialised per chain "program" so that every layer loop is unrolled and nothing is looked
// up at run time, and (b) hides the issue -> MMA -> commit -> TMEM load -> activation -> fp16 operand tile -> next issue
// round trip with SIX independent row slots per CTA.  A slot is one warpgroup (4 warps = the 4 TMEM lane quadrants)
// that owns a TMEM accumulator and an operand tile and runs its rows strictly in sequence; one of its warps also
// issues the slot's MMAs, so the only cross-warp hand-offs are the tcgen05.commit -> mbarrier wake-up and a 128-thread
// named barrier.  All slots share one resident copy of the weights and one TMA row ring (warp 0).  One CTA per SM;
// the (strip, row) space is cut into one contiguous range per CTA so that all 148 SMs carry the same number of rows.
#pragma once
#include <type_traits>

#include "conv_p8.cuh"

namespace hdrtv {

constexpr int kMaxChain = 9;
constexpr int kChainGroups = 5;
constexpr int kChainMaxRing = 16;                // >= kChainGroups + ks - 1 rows are in use at once; the rest is prefetch
constexpr int kChainThreads = 32 * (1 + kChainGroups * 4);
constexpr int kTileBytes = 8 * kPlaneBytes;      // one 64-channel operand tile in shared memory (probes only)
constexpr int kSlotCols = 96;                    // TMEM columns per row slot: 64 accumulator + 32 operand (64 ch fp16)

// ---- chain programs (compile-time layer tables) -------------------------------------------------------------------
// store: 0 none, 1 P8, 2 planar fp16 (3 channels) + single-chunk P8.   act: 0 none, 1 ReLU, 2 LeakyReLU(0.1).
struct ProgAGCM {
  static constexpr int L = 3, KS = 1, IN_PLANES = 1;
  static constexpr int N[L] = {64, 64, 16};
  static constexpr int STEPS[L] = {1, 4, 4};          // K/16 MMAs per layer (layer 0: the input side's tap steps)
  static constexpr int APLANE[L] = {0, 0, 0};
  static constexpr int WRITE[L] = {1, 1, 0};
  static constexpr int STORE[L] = {0, 0, 2};
  static constexpr int ACT[L] = {1, 1, 0};
  // layer-0 operand (one 8-channel plane per ring slot): byte offset / K-half distance of each tap step
  static constexpr int A0_OFF[1] = {16};
  static constexpr int A0_LBO[1] = {16};
};
struct ProgCond {
  static constexpr int L = 6, KS = 3, IN_PLANES = 1;
  static constexpr int N[L] = {64, 64, 64, 64, 64, 16};
  static constexpr int STEPS[L] = {6, 4, 4, 4, 4, 4};
  static constexpr int APLANE[L] = {0, 0, 0, 0, 0, 0};
  static constexpr int WRITE[L] = {1, 1, 1, 1, 1, 0};
  static constexpr int STORE[L] = {0, 0, 1, 0, 0, 1};
  static constexpr int ACT[L] = {2, 2, 2, 2, 2, 0};
  static constexpr int A0_OFF[2] = {0, 32};           // per input row: taps (dx 0, dx 1) paired, then dx 2 (+ zero half)
  static constexpr int A0_LBO[2] = {16, 16};
};
// ProgCond + the stage-0 convs of the two full-resolution SFT layers (16 -> 2 x 32, LeakyReLU) as a seventh step on
// cond1: S0STORE = 1 stores all 64 channels to outs[6]; 3 stores chunks 0-3 to outs[6] and chunks 4-7 to outs2 (the
// parity-split home of the layer an up-conv applies).                                   (arch_util.py:63-72)
// C1STORE = 1 also stores cond1 (only a debugging output once its sole consumer, the SFT stage 0, is the next step).
template <int S0STORE, int C1STORE = 0>
struct ProgCondSft {
  static constexpr int L = 7, KS = 3, IN_PLANES = 1;
  static constexpr int N[L] = {64, 64, 64, 64, 64, 16, 64};
  static constexpr int STEPS[L] = {6, 4, 4, 4, 4, 4, 1};
  static constexpr int APLANE[L] = {0, 0, 0, 0, 0, 0, 0};
  static constexpr int WRITE[L] = {1, 1, 1, 1, 1, 1, 0};
  static constexpr int STORE[L] = {0, 0, 1, 0, 0, C1STORE, S0STORE};
  static constexpr int ACT[L] = {2, 2, 2, 2, 2, 0, 2};
  static constexpr int A0_OFF[2] = {0, 32};
  static constexpr int A0_LBO[2] = {16, 16};
};
// Tails of the condition pyramid, one launch per level: the 1x1 convs that follow the stride-2 conv of CondNet2 / CondNet3
// and stage 0 of the level's four SFT layers (16 -> 4 x 32 channels, LeakyReLU), split into two 64-channel steps that both
// read the 16-channel condition map left in the slot's operand columns.  Layer 0 is a 1x1 conv on the 64-channel row in
// the ring (eight channel-chunk planes per slot).  The last step stores chunks 0-3 next to the first half and chunks
// 4-7 to the parity-split home of the layer an up-conv applies (STORE 3).        (HDRUNet3T1_arch.py:40-62, arch_util.py:63-72)
struct ProgTail2 {            // CondNet2.2 (64->64, LeakyReLU) -> CondNet2.4 (64->16) -> stage 0 a -> stage 0 b
  static constexpr int L = 4, KS = 1, IN_PLANES = 8;
  static constexpr int N[L] = {64, 16, 64, 64};
  static constexpr int STEPS[L] = {4, 4, 1, 1};
  static constexpr int APLANE[L] = {0, 0, 0, 0};
  static constexpr int WRITE[L] = {1, 1, 0, 0};
  static constexpr int STORE[L] = {0, 0, 1, 3};
  static constexpr int ACT[L] = {2, 0, 2, 2};
  static constexpr int A0_OFF[4] = {16, 16 + 2 * kPlaneBytes, 16 + 4 * kPlaneBytes, 16 + 6 * kPlaneBytes};
  static constexpr int A0_LBO[4] = {kPlaneBytes, kPlaneBytes, kPlaneBytes, kPlaneBytes};
};
struct ProgTail3 {            // CondNet3.4 (64->16) -> stage 0 a -> stage 0 b
  static constexpr int L = 3, KS = 1, IN_PLANES = 8;
  static constexpr int N[L] = {16, 64, 64};
  static constexpr int STEPS[L] = {4, 1, 1};
  static constexpr int APLANE[L] = {0, 0, 0};
  static constexpr int WRITE[L] = {1, 0, 0};
  static constexpr int STORE[L] = {0, 1, 3};
  static constexpr int ACT[L] = {0, 2, 2};
  static constexpr int A0_OFF[4] = {16, 16 + 2 * kPlaneBytes, 16 + 4 * kPlaneBytes, 16 + 6 * kPlaneBytes};
  static constexpr int A0_LBO[4] = {kPlaneBytes, kPlaneBytes, kPlaneBytes, kPlaneBytes};
};
struct ProgTail4 {            // level 3: stage 0 of its eight SFT layers (16 -> 8 x 32) as four 64-channel steps.  Layer 0 is an
  static constexpr int L = 5, KS = 1, IN_PLANES = 2;   // identity 1x1 that moves the 16-channel map from the ring into the operand columns
  static constexpr int N[L] = {16, 64, 64, 64, 64};
  static constexpr int STEPS[L] = {1, 1, 1, 1, 1};
  static constexpr int APLANE[L] = {0, 0, 0, 0, 0};
  static constexpr int WRITE[L] = {1, 0, 0, 0, 0};
  static constexpr int STORE[L] = {0, 1, 1, 1, 1};
  static constexpr int ACT[L] = {0, 2, 2, 2, 2};
  static constexpr int A0_OFF[1] = {16};
  static constexpr int A0_LBO[1] = {kPlaneBytes};
};
template <class P>
__host__ __device__ constexpr int prog_w_off(int l) {   // byte offset of layer l's packed weights (steps + bias step)
  int off = 0;
  for (int i = 0; i < l; ++i) off += (P::STEPS[i] + 1) * P::N[i] * 32;
  return off;
}

struct ChainParams {
  ConvParams base;           // input side (ring geometry), weights pointer / total bytes, Ho/Wo, planar output
  int active_slots;          // row slots in use (<= kChainGroups; tuning / experiments)
  int strips;                // 128-pixel strips per row; the grid is 1-D, each CTA takes a contiguous (strip, row) range
  P8 outs[kMaxChain];        // per layer, where STORE != 0
  P8 outs2;                  // second home of a split store (STORE == 3)
  ActQuant opq[kMaxChain];   // INT8 layouts: layer l+1 is a W8A8 layer -> its input quantiser is applied to layer l's operand
                             // (only for layers whose output is not also stored: the stored tensor would be the raw one)
  // INT8 layouts: uint8 copies of the stored 64-channel layer (STORE == 1; `cond`) through the input quantisers of its W8A8
  // consumers that run on kind::i8 (CondNet3.0, CondNet4.0): 16 channels per 16-byte entry
  ActQuant q8[2];
  P8 outq8[2];
  long long* trace;          // only read when compiled with HDRTV_CHAIN_TRACE: clock64 stamps of CTA 0 [step<64][slot<8][8]
};

__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}
// Parked wait for the epilogue side (long suspend hint: fewer polling instructions competing with working warps).
// Bounded by a poll COUNT (2^27 polls of >= ~30 ns each, up to 2 us when parked: between ~4 s and a few minutes): a loaded
// box only makes polls slower, so it cannot trip the bound early, and the loop carries no 64-bit timer state - this wait sits
// in the hottest loop of the chain kernels, which run at their register cap (the %globaltimer variant lives in the out-of-line
// mbar_wait_slow).
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, int* err_word, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 2000)) {
    if (++spins > (1u << 27)) {
      if (err_word) atomicExch(err_word, code);
      __threadfence_system();
      __trap();
    }
  }
}

// Contiguous range of (strip, row) work items of this CTA, walked as segments that stay inside one strip.
struct ChainWalk {
  long lo, hi;
  int Ho;
  __device__ __forceinline__ ChainWalk(int strips, int Ho_) : Ho(Ho_) {
    const long total = static_cast<long>(strips) * Ho_;
    lo = total * blockIdx.x / gridDim.x;
    hi = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ __forceinline__ bool next(int& strip, int& r0, int& n) {
    if (lo >= hi) return false;
    strip = static_cast<int>(lo / Ho);
    r0 = static_cast<int>(lo - static_cast<long>(strip) * Ho);
    n = static_cast<int>(min(static_cast<long>(Ho - r0), hi - lo));
    lo += n;
    return true;
  }
};

#ifdef HDRTV_CHAIN_TRACE
#define CHAIN_STAMP(k) do { if (tr) cp.trace[(e * 8 + g) * 8 + (k)] = clock64(); } while (0)
#else
#define CHAIN_STAMP(k) do { } while (0)
#endif

// QOP: INT8 layouts - cp.opq[l] (the input quantiser of a W8A8 layer l+1) is applied to layer l's operand.  Compile-time, so
// that the FP16 instances carry none of it.
template <class Prog, bool QOP = false>
__global__ void __launch_bounds__(kChainThreads, 1) chain_p8_kernel(const __grid_constant__ ChainParams cp) {
  const ConvParams& p = cp.base;
  constexpr int G = kChainGroups;
  constexpr int L = Prog::L, KS = Prog::KS;
  constexpr int SPD = Prog::STEPS[0] / KS;            // layer-0 tap steps per input row
  constexpr uint32_t kTmemCols = 512;                 // G * kSlotCols = 480 rounded up to a power of two
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kChainMaxRing + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (2 * kChainMaxRing + i); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kChainMaxRing + G);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kChainMaxRing + G + 1));
  uint8_t* ones = smem + 512;
  uint8_t* wsm = smem + kSmemHeader;
  uint8_t* ring = wsm + ((p.w_bytes + 127) & ~127);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), KS);                    // one arrival per (output row, dy) use of the input row
    }
    for (int i = 0; i < G; ++i) mbar_init(tfull_bar(i), 1);
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kPlaneEntries) {
    reinterpret_cast<uint4*>(ones)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ChainWalk walk(cp.strips, p.Ho);
  int strip, r0, n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (p.weights_dynamic) grid_dep_wait();
      mbar_expect_tx(wfull_bar, p.w_bytes);
      bulk_g2s(smem_u32(wsm), p.wpk, p.w_bytes, wfull_bar);
      grid_dep_wait();          // static weights are on their way; activations need the previous kernel finished
      const uint32_t row_tx = p.n_copies * p.copy_bytes;
      int slot = 0, ph = 1;
      while (walk.next(strip, r0, n)) {
        const uint4* src = p.in + (static_cast<long>(r0) + p.row_bias) * p.in_row_entries + strip * kTileM;
        for (int q = 0; q < n - 1 + KS; ++q) {
          mbar_wait(empty_bar(slot), ph, p.err, 11);
          mbar_expect_tx(full_bar(slot), row_tx);
          const uint32_t dst = smem_u32(ring) + slot * p.slot_bytes;
          for (int c = 0; c < p.n_copies; ++c)
            bulk_g2s(dst + p.copies[c].dst_off, src + p.copies[c].src_off, p.copy_bytes, full_bar(slot));
          src += p.in_row_entries;
          if (++slot == p.ring) { slot = 0; ph ^= 1; }
        }
      }
      grid_dep_launch();        // all input requested: the next kernel's CTAs may start their prologue
    }
  } else {
    // ------------------------------------------------------------------ row slot g: one warpgroup
    grid_dep_wait();            // the epilogue overwrites buffers the previous kernel may still read
    const int g = (warp - 1) >> 2;
    const bool issuer = ((warp - 1) & 3) == (g & 3);  // this warp also issues the slot's MMAs; spread over the 4 SMSPs
    const bool lead = elect_one();                     // the issuing lane of an issuer warp, elected once
    const int lg = warp & 3;                          // TMEM lane quadrant this warp may read
    const int m = lg * 32 + lane;                     // pixel inside the strip = TMEM lane
    const uint32_t d_tmem = tmem_base + g * kSlotCols;          // accumulator: 64 columns
    const uint32_t a_tmem = d_tmem + 64;                        // next layer's A operand: 32 columns (2 channels each)
    const uint32_t tlane = d_tmem + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t alane = a_tmem + (static_cast<uint32_t>(lg * 32) << 16);
    // issuer state
    const uint32_t ring16 = smem_u32(ring) >> 4, w16 = smem_u32(wsm) >> 4;
    const uint32_t slot16 = p.slot_bytes >> 4;
    const int ring_n = p.ring;
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
    auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);
    if (issuer) mbar_wait(wfull_bar, 0, p.err, 12);

    int e = 0;                 // steps this slot has run (parity of its tfull barrier)
    int R = 0, Q = 0;          // output rows / input rows of the segments before the current one
    while (walk.next(strip, r0, n)) {
      const int x = strip * kTileM + m;
      const bool xin = x < p.Wo;
      const int GA = cp.active_slots;
      for (int t = (g - R % GA + GA) % GA; t < n && g < GA; t += GA) {
        const int oy = r0 + t;
        static_for<0, L>([&](auto lc) {
          constexpr int l = decltype(lc)::value;
          constexpr int N = Prog::N[l];
          constexpr int NSTEPS = Prog::STEPS[l];
          constexpr int kStore = Prog::STORE[l], kWrite = Prog::WRITE[l], kAct = Prog::ACT[l], kAPlane = Prog::APLANE[l];
#ifdef HDRTV_CHAIN_TRACE
          const bool tr = cp.trace != nullptr && blockIdx.x == 0 && e < 64 && lane == 0 && issuer;
#endif
          CHAIN_STAMP(0);
          if (issuer) {
            tc_fence_after();                         // the slot's previous epilogue (group barrier) drained TMEM
            constexpr uint32_t idesc = make_idesc_f16_m128(N);
            constexpr uint32_t b_lbo = static_cast<uint32_t>(N) << 16;        // (N*16 bytes) >> 4, in the LBO field
            constexpr uint32_t b_step = static_cast<uint32_t>(N) * 2;         // (N*32 bytes) >> 4
            const uint32_t b_lo0 = (w16 + (prog_w_off<Prog>(l) >> 4)) | b_lbo;
            if constexpr (l == 0) {
#pragma unroll
              for (int dy = 0; dy < KS; ++dy) {
                const int q = Q + t + dy;
                const int slot = q % ring_n, ph = (q / ring_n) & 1;
                mbar_wait(full_bar(slot), ph, p.err, 14);
                tc_fence_after();
                const uint32_t a16 = ring16 + slot * slot16;
                if (lead) {
                  static_for<0, SPD>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    constexpr uint32_t a_off16 = Prog::A0_OFF[i] >> 4, a_lbo16 = Prog::A0_LBO[i] >> 4;
                    tc_mma_f16(d_tmem, mkdesc((a16 + a_off16) | (a_lbo16 << 16)), mkdesc(b_lo0 + (dy * SPD + i) * b_step), idesc,
                               (dy | i) ? 1u : 0u);
                  });
                  // release: this (row t, dy) use of the input row is finished once the MMAs above complete.  The
                  // first and last rows of a segment have fewer users than KS; their issuer supplies the missing arrivals.
                  const int uses = 1 + (t == 0 ? KS - 1 - dy : 0) + (t == n - 1 ? dy : 0);
                  for (int u = 0; u < uses; ++u) tc_commit(empty_bar(slot));
                }
                __syncwarp();
              }
              if (lead) {
                tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + NSTEPS * b_step), idesc, 1u);      // + bias
                tc_commit(tfull_bar(g));
              }
            } else {
              if (lead) {
                // A = the previous layer's activations, written to TMEM by the epilogue (8 columns per K = 16 step)
#pragma unroll
                for (int i = 0; i < NSTEPS; ++i)
                  tc_mma_f16_ts(d_tmem, a_tmem + kAPlane * 4 + i * 8, mkdesc(b_lo0 + i * b_step), idesc, i ? 1u : 0u);
                tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + NSTEPS * b_step), idesc, 1u);      // + bias
                tc_commit(tfull_bar(g));
              }
            }
            __syncwarp();
          }
          CHAIN_STAMP(1);
          mbar_wait_parked(tfull_bar(g), e & 1, p.err, 15);
          tc_fence_after();
          CHAIN_STAMP(2);
          auto act = [](float v) {
            if constexpr (kAct == 1) return fmaxf(v, 0.f);
            else if constexpr (kAct == 2) return fmaxf(v, 0.1f * v);
            else return v;
          };
          if constexpr (N == 64) {
            float v[64];
            tmem_ld_cols<64>(tlane, v);
            tc_fence_before();
            CHAIN_STAMP(3);
            uint4 h[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float a[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] = act(v[c * 8 + k]);
              h[c] = pack8(a);
              if constexpr (QOP && kWrite != 0 && kStore == 0) {
                if (cp.opq[l].mode) h[c] = fq_entry(h[c], cp.opq[l]);
              }
            }
            CHAIN_STAMP(4);
            if constexpr (kWrite != 0) {
              tmem_st32(alane, reinterpret_cast<const uint32_t*>(h));
              CHAIN_STAMP(5);
              tc_wait_st();
            }
            CHAIN_STAMP(6);
            if constexpr (kStore == 1) {
              if (xin) {
                ColRef o;
                o.init(cp.outs[l], x);
#pragma unroll
                for (int c = 0; c < 8; ++c) *o.at(oy, c) = h[c];
                if constexpr (QOP) {
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    if (cp.q8[i].mode) {
                      ColRef oq;
                      oq.init(cp.outq8[i], x);
#pragma unroll
                      for (int c = 0; c < 8; c += 2) {
                        float f0[8], f1[8];
                        unpack8(h[c], f0);
                        unpack8(h[c + 1], f1);
                        *oq.at(oy, c >> 1) = pack16_u8h(f0, f1, cp.q8[i]);
                      }
                    }
                  }
                }
              }
            } else if constexpr (kStore == 3) {
              if (xin) {
                ColRef o, o2;
                o.init(cp.outs[l], x);
                o2.init(cp.outs2, x);
#pragma unroll
                for (int c = 0; c < 4; ++c) *o.at(oy, c) = h[c];
#pragma unroll
                for (int c = 0; c < 4; ++c) *o2.at(oy, c) = h[4 + c];
              }
            }
          } else {   // N == 16
            float v[16];
            tmem_ld_cols<16>(tlane, v);
            tc_fence_before();
            uint4 h[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float a[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] = (kStore == 2 && c * 8 + k >= 3) ? 0.f : act(v[c * 8 + k]);
              h[c] = pack8(a);
            }
            if constexpr (kWrite != 0) {
              tmem_st8(alane, reinterpret_cast<const uint32_t*>(h));
              tc_wait_st();
            }
            if constexpr (kStore != 0) {
              if (xin) {
                ColRef o;
                o.init(cp.outs[l], x);
                if constexpr (kStore == 1) {
                  *o.at(oy, 0) = h[0];
                  *o.at(oy, 1) = h[1];
                } else {
                  const __half* hv = reinterpret_cast<const __half*>(&h[0]);
#pragma unroll
                  for (int k = 0; k < 3; ++k) p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = hv[k];
                  *o.at(oy, 0) = h[0];
                }
              }
            }
          }
          CHAIN_STAMP(7);
          tc_fence_before();
          group_barrier(1 + g);                       // operand written + accumulator drained by all four warps of the slot
          ++e;
        });
      }
      R += n;
      Q += n - 1 + KS;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

inline size_t chain_smem_bytes(const ChainParams& cp) {
  return kSmemHeader + ((cp.base.w_bytes + 127) & ~127) + static_cast<size_t>(cp.base.ring) * cp.base.slot_bytes;
}

}  // namespace hdrtv
