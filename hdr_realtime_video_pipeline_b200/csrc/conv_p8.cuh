// Row-streaming implicit-GEMM convolution on tcgen05 / TMEM for the P8 activation layout.
//
// One CTA owns a strip of 128 output pixels (the UMMA M dimension) and walks down a band of output rows.
// Every input row of the strip is fetched ONCE by 1-D bulk TMA copies (one per channel-chunk plane) into a
// ring of shared-memory row slots; the 3x3 taps are then nothing but byte shifts of the K-major, un-swizzled
// UMMA operand descriptor over those resident rows (dx = +16 B, dy = next ring slot).  The accumulator of one
// output row (128 pixels x N output channels, fp32) lives in TMEM, double-buffered so that the epilogue warps
// drain row t while the single MMA-issuing thread already runs row t+1.
//
// The bias is added by the tensor core too: one extra K=16 step multiplies a constant [1,1,0,..] operand with
// [bias_hi, bias_lo, 0,..] (fp16 hi/lo split, ~22-bit bias), so the epilogue has no per-channel loads.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..9 = epilogue:
// two warps per TMEM lane quadrant, each draining one half of the N columns.  Residual / SFT operands of a row
// are requested BEFORE the wait on the accumulator so their HBM latency hides behind the MMAs.
//
// Replaces the cuDNN conv2d calls of the reference's eager path (Condition_arch.py:571-583,
// HDRUNet3T1_arch.py:160-205, arch_util.py:68-95) with bias / activation / residual / SFT / PixelShuffle fused.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace hdrtv {

constexpr int kTileM = 128;
constexpr int kPlaneEntries = 136;                 // 128 + halo + pairing slack, multiple of 8
constexpr int kPlaneBytes = kPlaneEntries * 16;    // 2176
constexpr int kMaxSteps = 40;
constexpr int kMaxCopies = 16;
constexpr int kMaxRing = 8;
constexpr int kConvThreads = 320;
constexpr int kSmemHeader = 512 + kPlaneBytes + 128;   // barriers, step-descriptor table, constant "ones" operand (bias step)

enum StoreMode : int { STORE_P8 = 0, STORE_PS = 1, STORE_PLANAR = 2 };

// Input side of a convolution = how the ring slot of one input row is laid out and which K = 16 operand windows
// ("tap steps") one output row reads from it.  Compile-time per kernel instance: the single MMA-issuing warp must not
// look anything up while the tensor pipe waits (a dynamically indexed parameter load is a long-scoreboard stall).
enum InKind : int { IN_NAT3x3 = 0, IN_NAT1x1 = 1, IN_NAT3x3_C8 = 2, IN_NAT1x1_C8 = 3, IN_PAR3x3S2 = 4, IN_PAR1x1 = 5 };
__host__ __device__ constexpr int kind_ks(int k) { return (k == IN_NAT3x3 || k == IN_NAT3x3_C8 || k == IN_PAR3x3S2) ? 3 : 1; }
__host__ __device__ constexpr int kind_stride(int k) { return k == IN_PAR3x3S2 ? 2 : 1; }
__host__ __device__ constexpr int kind_copies(int k, int kch) { return k == IN_PAR3x3S2 ? 2 * kch : kch; }
// tap steps per input row (dy)
__host__ __device__ constexpr int kind_spd(int k, int kch) {
  return (k == IN_NAT3x3 || k == IN_PAR3x3S2) ? 3 * kch / 2 : (k == IN_NAT3x3_C8 ? 2 : (k == IN_NAT1x1_C8 ? 1 : kch / 2));
}
// byte offset of tap step i's A operand inside the slot, and distance between its two 8-channel K halves
__host__ __device__ constexpr uint32_t kind_a_off(int k, int kch, int i) {
  constexpr uint32_t PB = 2176;   // kPlaneBytes
  if (k == IN_NAT3x3) return static_cast<uint32_t>(2 * (i % (kch / 2))) * PB + static_cast<uint32_t>(i / (kch / 2)) * 16;
  if (k == IN_NAT1x1 || k == IN_PAR1x1) return static_cast<uint32_t>(2 * i) * PB + 16;
  if (k == IN_NAT3x3_C8) return i == 0 ? 0 : 32;
  if (k == IN_NAT1x1_C8) return 16;
  // IN_PAR3x3S2: dx = i / (kch/2); input x = 2*ox + dx - 1 -> parity plane (dx == 1 ? even : odd), shift (dx == 0 ? 0 : 16)
  const int dx = i / (kch / 2), c = i % (kch / 2);
  return static_cast<uint32_t>((2 * c) * 2 + (dx == 1 ? 0 : 1)) * PB + (dx == 0 ? 0u : 16u);
}
__host__ __device__ constexpr uint32_t kind_a_lbo(int k) {
  return (k == IN_NAT3x3_C8 || k == IN_NAT1x1_C8) ? 16u : (k == IN_PAR3x3S2 ? 2u * 2176u : 2176u);
}

struct ConvStep {
  uint16_t row;      // input row of this output row's window (dy)
  uint16_t release;  // 1: the ring slot of `row` is dead after this step
  uint32_t a_off;    // byte offset of the A operand inside the slot (plane + horizontal tap shift)
  uint32_t a_lbo;    // byte distance between the two 8-channel K halves
};
struct ConvCopy {
  uint32_t src_off;  // entries, relative to (row start + x0)
  uint32_t dst_off;  // bytes inside the slot
};

struct ConvParams {
  const uint4* in;
  long in_row_entries;
  long in_z_entries;   // added to the source per blockIdx.z (parity plane of a parity-split input read by a 1x1)
  int xmul;            // output pixel x = xmul * (strip pixel) + blockIdx.z
  int row_bias, stride, ks;
  int n_copies, copy_bytes;
  ConvCopy copies[kMaxCopies];
  uint32_t copy_src0, copy_src_stride, copy_par_off;   // source entry of copy c: src0 + (c / npar) * stride + (c % npar) * par_off
  int slot_bytes, ring;
  int n_steps;         // tap steps; the weight buffer holds n_steps + 1 (the last one is the bias step)
  ConvStep steps[kMaxSteps];
  const uint4* wpk;
  int w_bytes;
  // in-kernel SFT generator (SFTG instances): per output row, scale/shift = stage-1 1x1 conv (32 -> 64, block diagonal)
  // of the 32-channel stage-0 map `s0` (arch_util.py:63-72), accumulated in TMEM next to the conv accumulator
  const uint4* wpk2;        // packed stage-1 weights: 2 tap steps + bias step, N = 64
  int w2_bytes;
  const uint4* s0;          // stage-0 map (P8, natural layout, output resolution)
  long s0_row_entries;      // entries per row of that tensor
  uint32_t s0_src0;         // first chunk plane of this SFT layer: j0 * Wp
  uint32_t s0_wp;
  int Ho, Wo, band;
  float slope;         // activation as max(v, slope*v): 1 = none, 0 = ReLU, 0.1 = LeakyReLU(0.1)
  int has_res, has_res2, has_sft, has_raw;
  P8 res, res2, sft, out, raw;
  int out_split;            // > 0: output chunks >= out_split go to out2 (chunk index - out_split) instead of out
  P8 out2;
  // zsplit > 1: `zsplit` convolutions that read the SAME input (different weights / outputs of identical geometry) share
  // one launch; CTA blockIdx.x serves variant blockIdx.x % zsplit of strip blockIdx.x / zsplit, so the variants of a
  // strip run side by side and the input rows are fetched from HBM once (the other reads hit L2).
  int weights_dynamic;      // the packed weights are written by the previous kernel in the stream (AGCM fold)
  int zsplit;
  const uint4* wpk_z[3];
  P8 out_zp[3];             // per-variant output tensors (layouts may differ)
  __half* planar;
  long planar_plane;
  int planar_W;
  int* err;
  // ---- INT8 layouts (W8A8Conv2d, hdrtvnet_torch.py:296-364) -------------------------------------------------------
  // I8 instances: `in` is a uint8 tensor (16 channels per 16-byte entry: the P8 geometry of C/2 fp16 channels) holding the
  // codes q of the layer's input quantiser, the weights are int8, the accumulator is S32 (tcgen05.mma.kind::i8) and the
  // epilogue de-quantises:  conv(x^, w)[n] + b[n] = acc[n] * (x_scale * w_scale[n]) + beta[cls][n]   with
  // beta[cls][n] = b[n] + x_zero * w_scale[n] * sum over the taps that are INSIDE the image for border class cls of
  // sum_c w_int8[n, c, tap]  (zero padding is applied to x^, not to q: a padded tap contributes nothing, not x_zero).
  const float* i8_alpha;      // [N]
  const float* i8_beta;       // [16][N], cls = cy * 4 + cx, bit 0: first tap outside, bit 1: last tap outside
  int i8_H, i8_W;             // input size
  int* i8_acc_dump;           // test build: raw S32 accumulators, planar [N][Ho][Wo]
  // Producer side of a W8A8 consumer: the values stored to `out` pass through that consumer's input quantiser, either as
  // fp16 x^ (the consumer runs f16 MMAs on de-quantised values, like the reference's eager INT8 path) or, out_u8 = 1, as
  // the uint8 codes themselves (`out` is then a uint8 tensor for an I8 consumer).  `raw` / `out2` stay unquantised.
  ActQuant out_q;
  int out_u8;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// Per-thread view of a P8 tensor at a fixed pixel column: entry(y, j) = (y+1)*row_entries + j*Wp + xoff.
struct ColRef {
  uint4* base;
  long row_entries;
  long xoff;
  int Wp;
  __device__ __forceinline__ void init(const P8& t, int x) {
    base = reinterpret_cast<uint4*>(t.base);
    row_entries = t.row_entries();
    Wp = t.Wp;
    xoff = t.parity ? static_cast<long>(x & 1) * (t.Wp >> 1) + ((x >> 1) + 1) : static_cast<long>(x + 1);
  }
  __device__ __forceinline__ uint4* at(int y, int j) const { return base + (static_cast<long>(y + 1) * row_entries + static_cast<long>(j) * Wp + xoff); }
};

template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

constexpr int kSRing = 4;                         // stage-0 rows in flight (SFTG)
constexpr int kSSlotBytes = 4 * kPlaneBytes;      // one 32-channel stage-0 row
constexpr int kSRingPS = 2;                       // PixelShuffle consumers: 4 sub-pixel rows (2 fine rows x 2 column
constexpr int kSSlotBytesPS = 16 * kPlaneBytes;   // parities) of 32 channels per coarse output row

// FOLD (stride-2 3x3 only): the vertical taps are folded into N as in conv2x_p8.cuh.  Output row m reads input rows
// 2m (dy 0), 2m+1 (dy 1), 2m+2 (dy 2), so an even input row 2m is multiplied once by [W(dy=2) | W(dy=0)] into the
// neighbouring accumulator blocks of output rows m-1 and m, an odd one by W(dy=1) into block m: 2/3 of the MMAs, every
// input row consumed (and its ring slot released) once, and a ring of kFoldR accumulator blocks instead of two stages.
constexpr int kFoldR = 8;

// eight uint8 codes of fp16-representable values -> two 32-bit words
__device__ __forceinline__ uint2 quant8_u8h(const float* x, const ActQuant& q) {
  float u[8];
  quant_round8(x, q, u);
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = __float_as_uint(u[k]);
  uint2 r;
  r.x = __byte_perm(__byte_perm(w[0], w[1], 0x0040), __byte_perm(w[2], w[3], 0x0040), 0x5410);
  r.y = __byte_perm(__byte_perm(w[4], w[5], 0x0040), __byte_perm(w[6], w[7], 0x0040), 0x5410);
  return r;
}
// 16 uint8 codes (two 8-channel groups of one pixel, fp16-representable values) -> one 16-byte entry
__device__ __forceinline__ uint4 pack16_u8h(const float* a, const float* b, const ActQuant& q) {
  const uint2 lo = quant8_u8h(a, q), hi = quant8_u8h(b, q);
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
}
// fp16 entry -> the fake-quantised fp16 entry x^ = q * scale + zero (two roundings, like the reference), cast back to fp16
__device__ __forceinline__ uint4 fq_entry(const uint4& h, const ActQuant& q) {
  float x[8], u[8];
  unpack8(h, x);
  quant_round8(x, q, u);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = __fadd_rn(__fmul_rn(__fsub_rn(u[k], kRintMagic), q.scale), q.zero);
  return pack8(x);
}
// the same for fp32 epilogue values: rounded to fp16 first (the tensor a W8A8 layer quantises is its producer's fp16 output)
__device__ __forceinline__ uint4 pack16_u8(const float* a, const float* b, const ActQuant& q) {
  float fa[8], fb[8];
  unpack8(pack8(a), fa);
  unpack8(pack8(b), fb);
  return pack16_u8h(fa, fb, q);
}
__device__ __forceinline__ uint4 pack8_fq(const float* f, const ActQuant& q) { return fq_entry(pack8(f), q); }
// border class of an output coordinate o of a 3-tap window with the given stride over an input of `size`
__device__ __forceinline__ int tap_class(int o, int stride, int size) {
  return ((o * stride - 1 < 0) ? 1 : 0) | ((o * stride + 1 >= size) ? 2 : 0);
}

template <int KIND, int KCH, int N, int MODE, bool AUX, bool SFTG = false, bool FOLD = false, bool I8 = false>
__global__ void __launch_bounds__(kConvThreads, (MODE == STORE_PS) ? 1 : 2) conv_p8_kernel(const __grid_constant__ ConvParams p) {
  static_assert(!I8 || (!FOLD && kind_ks(KIND) == 3 && MODE != STORE_PLANAR), "INT8 instances: plain 3x3 convs");
  static_assert(!SFTG || (AUX && ((MODE == STORE_P8 && N == 32) || (MODE == STORE_PS && N == 128))),
                "in-kernel SFT generator: 32-channel outputs only");
  static_assert(!FOLD || (KIND == IN_PAR3x3S2 && !SFTG && MODE == STORE_P8 && N <= 64), "row folding: plain stride-2 3x3 convs");
  constexpr bool PSG = SFTG && MODE == STORE_PS;
  constexpr bool OQ = I8;               // output quantisers (out_q / out_u8) exist only in the INT8-layout instances
  // SFTG: 2 x 32 conv + 2 x 64 scale|shift columns; PixelShuffle: 2 x 128 conv + 4 sub-pixels x 64 (single-buffered)
  constexpr uint32_t kTmemCols = FOLD ? (kFoldR * N < 32 ? 32 : kFoldR * N) : SFTG ? (PSG ? 512 : 256) : ((2 * N < 32) ? 32 : 2 * N);
  constexpr int SRING = PSG ? kSRingPS : kSRing;
  constexpr int SSLOT = PSG ? kSSlotBytesPS : kSSlotBytes;
  constexpr int KS = kind_ks(KIND), STRIDE = kind_stride(KIND), SPD = kind_spd(KIND, KCH), NCOPY = kind_copies(KIND, KCH);
  constexpr int NPAR = KIND == IN_PAR3x3S2 ? 2 : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kMaxRing + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * ((FOLD ? 32 : 2 * kMaxRing) + i); };
  auto tempty_bar = [&](int i) { return bar0 + 8u * ((FOLD ? 40 : 2 * kMaxRing + 2) + i); };
  static_assert(2 * kMaxRing + 8 + 2 * kSRing <= 32 && 40 + kFoldR <= 64, "barrier table layout");
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxRing + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxRing + 5));
  auto sfull_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + 8 + i); };
  auto sempty_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + 8 + kSRing + i); };   // kSRing >= kSRingPS
  uint8_t* ones = smem + 512;
  uint8_t* wsm = smem + kSmemHeader;
  uint8_t* wsm2 = wsm + ((p.w_bytes + 127) & ~127);
  uint8_t* ring = wsm2 + (SFTG ? ((p.w2_bytes + 127) & ~127) : 0);
  uint8_t* sring = ring + p.ring * (kind_copies(KIND, KCH) * kPlaneBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int zs = p.zsplit > 1 ? p.zsplit : 1;
  const int zsel = zs > 1 ? static_cast<int>(blockIdx.x % zs) : 0;
  const int x0 = static_cast<int>(blockIdx.x / zs) * kTileM;
  const int oy0 = blockIdx.y * p.band;
  const int nrows_out = min(p.band, p.Ho - oy0);
  const int nrows_in = (nrows_out - 1) * STRIDE + KS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < (FOLD ? kFoldR : 2); ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 8);
    }
    mbar_init(wfull_bar, 1);
    if constexpr (SFTG) {
      for (int i = 0; i < SRING; ++i) {
        mbar_init(sfull_bar(i), 1);
        mbar_init(sempty_bar(i), 1);
      }
    }
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kPlaneEntries) {
    // constant A operand of the bias step: every row = [1, 1, 0, 0, 0, 0, 0, 0]
    reinterpret_cast<uint4*>(ones)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (p.weights_dynamic) grid_dep_wait();
      mbar_expect_tx(wfull_bar, p.w_bytes + (SFTG ? p.w2_bytes : 0));
      bulk_g2s(smem_u32(wsm), zs > 1 ? p.wpk_z[zsel] : p.wpk, p.w_bytes, wfull_bar);
      if constexpr (SFTG) bulk_g2s(smem_u32(wsm2), p.wpk2, p.w2_bytes, wfull_bar);
      grid_dep_wait();          // static weights are on their way; activations need the previous kernel finished
      constexpr uint32_t row_tx = NCOPY * kPlaneBytes;
      const uint32_t ring_n = p.ring, slot_bytes = NCOPY * kPlaneBytes;
      // The per-copy source address is formed with an explicit mad.wide: for `pointer + 32-bit offset` feeding
      // cp.async.bulk, ptxas 12.9 emitted a 32-bit uniform ULEA with a zeroed high word (illegal global address).
      const uint32_t sstride = p.copy_src_stride, spar = p.copy_par_off;
      const long row_entries = p.in_row_entries;
      uint32_t slot = 0, ph = 1;
      const uint4* src = p.in + (static_cast<long>(oy0) * STRIDE + p.row_bias) * row_entries + x0 + blockIdx.z * p.in_z_entries +
                         static_cast<long>(p.copy_src0);
      int ts = 0;
      uint32_t sslot = 0, sph = 1;
      // first stage-0 row of the band: output row oy0 (fine row 2*oy0 for PixelShuffle); +1 = the tensor's top pad row
      const uint4* ssrc = SFTG ? p.s0 + (static_cast<long>(oy0) * (PSG ? 2 : 1) + 1) * p.s0_row_entries +
                                     static_cast<long>(p.s0_src0) + x0
                               : nullptr;
      for (int q = 0; q < nrows_in; ++q) {
        mbar_wait(empty_bar(slot), ph, p.err, 1);
        mbar_expect_tx(full_bar(slot), row_tx);
        const uint32_t dst = smem_u32(ring) + slot * slot_bytes;
#pragma unroll
        for (int c = 0; c < NCOPY; ++c) {
          unsigned long long a;
          asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"((c / NPAR) * sstride + (c % NPAR) * spar), "l"(src));
          bulk_g2s(dst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, full_bar(slot));
        }
        src += row_entries;
        if (++slot == ring_n) { slot = 0; ph ^= 1; }
        if constexpr (SFTG) {
          // stage-0 row of every output row whose last input row has just been requested
          while (ts < nrows_out && ts * STRIDE + KS - 1 <= q) {
            mbar_wait(sempty_bar(sslot), sph, p.err, 6);
            mbar_expect_tx(sfull_bar(sslot), SSLOT);
            const uint32_t sdst = smem_u32(sring) + sslot * SSLOT;
            if constexpr (PSG) {
              // sub-pixel (i, j) of coarse pixel x: fine row 2*oy + i, column parity plane j, entry x (parity layout)
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const int sub = c >> 2, ch = c & 3;
                unsigned long long a;
                asm volatile("mad.wide.u32 %0, %1, 16, %2;"
                             : "=l"(a)
                             : "r"(ch * p.s0_wp + (sub & 1) * (p.s0_wp >> 1)), "l"(ssrc + (sub >> 1) * p.s0_row_entries));
                bulk_g2s(sdst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, sfull_bar(sslot));
              }
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                unsigned long long a;
                asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(c * p.s0_wp), "l"(ssrc));
                bulk_g2s(sdst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, sfull_bar(sslot));
              }
            }
            ssrc += (PSG ? 2 : 1) * p.s0_row_entries;
            ++ts;
            if (++sslot == SRING) { sslot = 0; sph ^= 1; }
          }
        }
      }
      // all of this CTA's input has been requested: let the next kernel's CTAs start their prologue.  (Triggering
      // earlier lets them pile onto whichever SMs drain first and unbalances the single-wave grids.)
      grid_dep_launch();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs this (warp-uniform) loop so that descriptor arithmetic stays in uniform registers; one
    // elected lane issues.  Tap-step offsets are compile-time constants: the issuing warp looks nothing up while
    // the tensor pipe waits (an M=128 x N<=64 x K=16 MMA costs the pipe only ~45 cycles).
    mbar_wait(wfull_bar, 0, p.err, 2);
    const bool lead = elect_one();         // the issuing lane, elected once (an elect.sync per row is ~50 cycles of idle pipe)
    constexpr uint32_t idesc = make_idesc_f16_m128(N);
    constexpr uint32_t idesc_i8 = make_idesc_i8_m128(N);
    (void)idesc_i8;
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
    auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(N) << 16;   // (N*16 bytes) >> 4 in the LBO field
    constexpr uint32_t b_step = static_cast<uint32_t>(N) * 2;    // (N*32 bytes) >> 4
    constexpr uint32_t a_lbo = (kind_a_lbo(KIND) >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsm) >> 4) | b_lbo;
    constexpr uint32_t slot16 = (NCOPY * kPlaneBytes) >> 4;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const int ring_n = p.ring;
    if constexpr (FOLD) {
      constexpr int R = kFoldR;
      constexpr uint32_t bE_lbo = static_cast<uint32_t>(2 * N) << 16, bE_step = 2 * N * 2;      // even rows: [dy=2 | dy=0]
      constexpr uint32_t bO_lbo = static_cast<uint32_t>(N) << 16, bO_step = N * 2;              // odd rows: dy=1
      const uint32_t bE0 = (smem_u32(wsm) >> 4) | bE_lbo;
      const uint32_t bO0 = ((smem_u32(wsm) + SPD * 2 * N * 32) >> 4) | bO_lbo;
      const uint64_t bias_desc = mkdesc(((smem_u32(wsm) + SPD * 3 * N * 32) >> 4) | (static_cast<uint32_t>(N) << 16));
      int slot = 0, ph = 0;
      for (int q = 0; q < nrows_in; ++q) {
        const int m = q >> 1;
        const bool even = (q & 1) == 0;
        const bool has_new = even && m < nrows_out;          // first contribution to output row m
        const int lo = even ? (m >= 1 ? m - 1 : m) : m, hi = even ? (m < nrows_out ? m : m - 1) : m;
        if (has_new) mbar_wait(tempty_bar(m % R), ((m / R) & 1) ^ 1, p.err, 3);
        mbar_wait(full_bar(slot), ph, p.err, 4);
        tc_fence_after();
        const int pos_lo = lo % R, nwin = hi - lo + 1;
        const int n1 = min(nwin, R - pos_lo), n2 = nwin - n1;      // the window may cross the end of the ring
        const uint32_t d1 = tmem_base + static_cast<uint32_t>(pos_lo) * N, d2 = tmem_base;
        const uint32_t idesc1 = make_idesc_f16_m128(static_cast<uint32_t>(N * n1));
        const uint32_t b1 = even ? bE0 + static_cast<uint32_t>(lo == m ? N : 0) : bO0, b2 = b1 + static_cast<uint32_t>(N * n1);
        const uint32_t bstep = even ? bE_step : bO_step;
        const uint32_t a16 = ring16 + slot * slot16;
        if (lead) {
          if (has_new)                                         // bias step: initialises the accumulator of the newest row
            tc_mma_f16(tmem_base + static_cast<uint32_t>(m % R) * N, ones_desc, bias_desc, idesc, 0u);
          static_for<0, SPD>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr uint32_t a_off16 = kind_a_off(KIND, KCH, i) >> 4;
            const uint64_t ad = mkdesc((a16 + a_off16) | a_lbo);
            tc_mma_f16(d1, ad, mkdesc(b1 + i * bstep), idesc1, 1u);
            if (n2) tc_mma_f16(d2, ad, mkdesc(b2 + i * bstep), idesc, 1u);
          });
          tc_commit(empty_bar(slot));                          // every input row is read exactly once
          if (even && m >= 1) tc_commit(tfull_bar((m - 1) % R));       // output row m-1 is complete
        }
        __syncwarp();
        if (++slot == ring_n) { slot = 0; ph ^= 1; }
      }
    } else {
    int waited = -1;
    int base_slot = 0, base_ph = 0;              // ring slot / phase of input row t*stride
    int sslot = 0, sph = 0;                      // stage-0 ring (SFTG)
    for (int t = 0; t < nrows_out; ++t) {
      const int stage = t & 1;
      mbar_wait(tempty_bar(stage), ((t >> 1) & 1) ^ 1, p.err, 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + stage * N;
      int slot = base_slot, ph = base_ph;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        const int q = t * STRIDE + dy;
        if (q > waited) {
          mbar_wait(full_bar(slot), ph, p.err, 4);
          waited = q;
          tc_fence_after();
        }
        const uint32_t a16 = ring16 + slot * slot16;
        if (lead) {
          static_for<0, SPD>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr uint32_t a_off16 = kind_a_off(KIND, KCH, i) >> 4;
            if constexpr (I8)
              tc_mma_i8(d_tmem, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (dy * SPD + i) * b_step), idesc_i8, (dy | i) ? 1u : 0u);
            else
              tc_mma_f16(d_tmem, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (dy * SPD + i) * b_step), idesc, (dy | i) ? 1u : 0u);
          });
          if (dy < STRIDE) tc_commit(empty_bar(slot));      // this input row is not needed by later output rows
          if (!SFTG && dy == KS - 1) {
            if constexpr (!I8) tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + (KS * SPD) * b_step), idesc, 1u);      // + bias
            tc_commit(tfull_bar(stage));
          }
        }
        __syncwarp();
        if (++slot == ring_n) { slot = 0; ph ^= 1; }
      }
      if constexpr (SFTG) {
        // scale|shift of this output row: [128 px x 32 stage-0 channels] x [32 -> 64 block-diagonal] (+ bias step);
        // PixelShuffle: one such GEMM per sub-pixel, into a single-buffered accumulator -> wait for the previous
        // row's epilogue (it arrives on tempty of the other stage after its last TMEM read)
        mbar_wait(sfull_bar(sslot), sph, p.err, 7);
        if (PSG && t > 0) mbar_wait(tempty_bar((t - 1) & 1), ((t - 1) >> 1) & 1, p.err, 8);
        tc_fence_after();
        if (lead) {
          constexpr uint32_t idesc64 = make_idesc_f16_m128(64);
          const uint32_t sb = (smem_u32(wsm2) >> 4) | (64u << 16);
#pragma unroll
          for (int sub = 0; sub < (PSG ? 4 : 1); ++sub) {
            const uint32_t sa = ((smem_u32(sring) + sslot * SSLOT + sub * kSSlotBytes + 16) >> 4) | ((kPlaneBytes >> 4) << 16);
            const uint32_t s_tmem = PSG ? tmem_base + 2 * N + sub * 64 : tmem_base + 2 * N + stage * 64;
            tc_mma_f16(s_tmem, mkdesc(sa), mkdesc(sb), idesc64, 0u);
            tc_mma_f16(s_tmem, mkdesc(sa + ((2 * kPlaneBytes) >> 4)), mkdesc(sb + 128), idesc64, 1u);
            tc_mma_f16(s_tmem, ones_desc, mkdesc(sb + 256), idesc64, 1u);
          }
          tc_commit(sempty_bar(sslot));
          if constexpr (!I8) tc_mma_f16(d_tmem, ones_desc, mkdesc(b_lo0 + (KS * SPD) * b_step), idesc, 1u);      // conv bias
          tc_commit(tfull_bar(stage));
        }
        __syncwarp();
        if (++sslot == SRING) { sslot = 0; sph ^= 1; }
      }
      base_slot += STRIDE;
      if (base_slot >= ring_n) { base_slot -= ring_n; base_ph ^= 1; }
    }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quadrant
    grid_dep_wait();            // reads skip/SFT operands and overwrites buffers the previous kernel may still use
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int x = p.xmul * (x0 + lg * 32 + lane) + blockIdx.z;
    const bool xin = x < p.Wo;
    const float slope = p.slope;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);

    if constexpr (MODE == STORE_PLANAR) {
      ColRef res, raw;
      if (p.has_res) res.init(p.res, x);
      if (p.has_raw) raw.init(p.raw, x);
      for (int t = 0; t < nrows_out; ++t) {
        const int stage = t & 1, oy = oy0 + t;
        uint4 r4 = make_uint4(0, 0, 0, 0);
        if (half == 0 && xin && p.has_res) r4 = __ldcg(res.at(oy, 0));
        mbar_wait(tfull_bar(stage), (t >> 1) & 1, p.err, 5);
        tc_fence_after();
        float v[8];
        if (half == 0) tmem_ld_cols<8>(tlane + stage * N, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(stage));
        if (half == 0 && xin) {
          float val[8], r[8];
          unpack8(r4, r);
#pragma unroll
          for (int k = 0; k < 8; ++k) val[k] = (k < 3) ? fmaxf(v[k], slope * v[k]) + r[k] : 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = __float2half_rn(val[k]);
          if (p.has_raw) *raw.at(oy, 0) = pack8(val);
        }
      }
    } else if constexpr (MODE == STORE_PS) {
      // N = 128 conv channels -> 32 channels at (2*oy+i, 2*x+j); conv channel n = 4*c + 2*i + j.
      // This warp: conv channels [64*half, 64*half+64) = output chunks 2*half, 2*half+1 of all four sub-pixels.
      ColRef res[2], sft[2], raw[2], out[2];     // index = sub-pixel column parity j
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        out[j].init(p.out, 2 * x + j);
        if (p.has_res) res[j].init(p.res, 2 * x + j);
        if (p.has_sft) sft[j].init(p.sft, 2 * x + j);
        if (p.has_raw) raw[j].init(p.raw, 2 * x + j);
      }
      for (int t = 0; t < nrows_out; ++t) {
        const int stage = t & 1, oy = oy0 + t;
        // request the row's skip-connection operands (4 sub-pixels x 2 chunks) before waiting for the accumulator:
        // with one CTA per SM nothing else hides their HBM latency
        uint4 r4[4][2];
        bool ok[4];
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const int Y = 2 * oy + (sub >> 1), j = sub & 1;
          ok[sub] = xin && Y < p.out.H && 2 * x + j < p.out.W;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            r4[sub][c] = make_uint4(0, 0, 0, 0);
            if (p.has_res && ok[sub]) r4[sub][c] = __ldcg(res[j].at(Y, 2 * half + c));
          }
        }
        mbar_wait(tfull_bar(stage), (t >> 1) & 1, p.err, 5);
        tc_fence_after();
        float v[64];
        tmem_ld_cols<64>(tlane + stage * N + half * 64, v);
        if constexpr (I8) {        // de-quantise: S32 accumulator -> conv + bias (border class of this pixel)
          const int cls = tap_class(oy, STRIDE, p.i8_H) * 4 + tap_class(x, STRIDE, p.i8_W);
          const float4* al = reinterpret_cast<const float4*>(p.i8_alpha + half * 64);
          const float4* be = reinterpret_cast<const float4*>(p.i8_beta + cls * N + half * 64);
#pragma unroll
          for (int k4 = 0; k4 < 16; ++k4) {
            const float4 a4 = __ldg(al + k4), b4 = __ldg(be + k4);
            if (p.i8_acc_dump && xin) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                p.i8_acc_dump[(static_cast<long>(half * 64 + 4 * k4 + e) * p.Ho + oy) * p.Wo + x] = __float_as_int(v[4 * k4 + e]);
            }
            v[4 * k4 + 0] = fmaf(__int2float_rn(__float_as_int(v[4 * k4 + 0])), a4.x, b4.x);
            v[4 * k4 + 1] = fmaf(__int2float_rn(__float_as_int(v[4 * k4 + 1])), a4.y, b4.y);
            v[4 * k4 + 2] = fmaf(__int2float_rn(__float_as_int(v[4 * k4 + 2])), a4.z, b4.z);
            v[4 * k4 + 3] = fmaf(__int2float_rn(__float_as_int(v[4 * k4 + 3])), a4.w, b4.w);
          }
        }
        if constexpr (!SFTG) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(stage));
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = fmaxf(v[k], slope * v[k]);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const int Y = 2 * oy + (sub >> 1), j = sub & 1;
          float sv[SFTG ? 16 : 1], tv[SFTG ? 16 : 1];
          if constexpr (SFTG) {     // scale / shift of this sub-pixel, channels [16*half, 16*half + 16)
            tmem_ld16_async(tlane + 2 * N + sub * 64 + 16 * half, reinterpret_cast<uint32_t*>(sv));
            tmem_ld16_async(tlane + 2 * N + sub * 64 + 32 + 16 * half, reinterpret_cast<uint32_t*>(tv));
            tc_wait_ld();
            if (sub == 3) {         // last TMEM read of this row: release the conv stage and the scale|shift columns
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(stage));
            }
          }
          if (ok[sub]) {
            uint4 s4[2], t4[2];
            if (!SFTG && p.has_sft) {
#pragma unroll
              for (int c = 0; c < 2; ++c) { s4[c] = __ldcg(sft[j].at(Y, 2 * half + c)); t4[c] = __ldcg(sft[j].at(Y, 2 * half + c + 4)); }
            }
            float fin[OQ ? 2 : 1][8];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int ch = 2 * half + c;
              float* val = fin[OQ ? c : 0];
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) val[cc] = v[32 * c + 4 * cc + sub];
              if (p.has_res) {
                float r[8];
                unpack8(r4[sub][c], r);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] += r[k];
              }
              if (p.has_raw) *raw[j].at(Y, ch) = pack8(val);
              if constexpr (SFTG) {
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], sv[c * 8 + k], tv[c * 8 + k]);   // sv = scale + 1 (bias step)
              } else if (p.has_sft) {
                float s[8], tt[8];
                unpack8(s4[c], s);
                unpack8(t4[c], tt);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], s[k], val[k]) + tt[k];
              }
              if constexpr (OQ) {
                if (!p.out_u8) *out[j].at(Y, ch) = p.out_q.mode ? pack8_fq(val, p.out_q) : pack8(val);
              } else {
                *out[j].at(Y, ch) = pack8(val);
              }
            }
            if constexpr (OQ) {
              if (p.out_u8) *out[j].at(Y, half) = pack16_u8(fin[0], fin[1], p.out_q);   // 16 channels = one uint8 entry
            }
          }
        }
      }
    } else {
      constexpr int COLS = N / 2;          // columns drained by this warp
      constexpr int CH = COLS / 8;         // 8-channel chunks per thread
      const int j0 = half * CH;
      ColRef out, out2, res, res2, sft, raw;
      if (zs > 1) out.init(p.out_zp[zsel], x);
      else out.init(p.out, x);
      const ActQuant oq = !OQ ? ActQuant{} : p.out_q;
      const bool ou8 = OQ && p.out_u8 != 0;
      if (p.out_split > 0) out2.init(p.out2, x);
      if constexpr (AUX) {
        if (p.has_res) res.init(p.res, x);
        if (p.has_res2) res2.init(p.res2, x);
        if (p.has_sft) sft.init(p.sft, x);
        if (p.has_raw) raw.init(p.raw, x);
      }
      for (int t = 0; t < nrows_out; ++t) {
        const int stage = FOLD ? t % kFoldR : t & 1, oy = oy0 + t;
        const uint32_t tpar = FOLD ? (t / kFoldR) & 1 : (t >> 1) & 1;
        uint4 r4[AUX ? CH : 1], q4[AUX ? CH : 1], s4[AUX ? CH : 1], t4[AUX ? CH : 1];
        if constexpr (AUX) {
          if (xin) {       // request the row's auxiliary operands before waiting for the accumulator
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              if (p.has_res) r4[c] = __ldcg(res.at(oy, j0 + c));
              if (p.has_res2) q4[c] = __ldcg(res2.at(oy, j0 + c));
              if (p.has_sft) { s4[c] = __ldcg(sft.at(oy, j0 + c)); t4[c] = __ldcg(sft.at(oy, j0 + c + N / 8)); }
            }
          }
        }
        mbar_wait(tfull_bar(stage), tpar, p.err, 5);
        tc_fence_after();
        float v[COLS];
        float sv[SFTG ? COLS : 1], tv[SFTG ? COLS : 1];
        tmem_ld_cols<COLS>(tlane + stage * N + half * COLS, v);
        if constexpr (SFTG) {
          tmem_ld_cols<COLS>(tlane + 2 * N + stage * 64 + half * COLS, sv);
          tmem_ld_cols<COLS>(tlane + 2 * N + stage * 64 + 32 + half * COLS, tv);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(stage));
        if constexpr (I8) {        // de-quantise: S32 accumulator -> conv + bias (border class of this pixel)
          const int cls = tap_class(oy, STRIDE, p.i8_H) * 4 + tap_class(x, STRIDE, p.i8_W);
          const float* al = p.i8_alpha + half * COLS;
          const float* be = p.i8_beta + cls * N + half * COLS;
#pragma unroll
          for (int k = 0; k < COLS; ++k) {
            if (p.i8_acc_dump && xin)
              p.i8_acc_dump[(static_cast<long>(half * COLS + k) * p.Ho + oy) * p.Wo + x] = __float_as_int(v[k]);
            v[k] = fmaf(__int2float_rn(__float_as_int(v[k])), __ldg(al + k), __ldg(be + k));
          }
        }
        if (xin) {
          float fin[OQ ? CH : 1][8];
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            float* val = fin[OQ ? c : 0];
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = fmaxf(v[c * 8 + k], slope * v[c * 8 + k]);
            if constexpr (AUX) {
              if (p.has_res) {
                float r[8];
                unpack8(r4[c], r);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] += r[k];
              }
              if (p.has_res2) {
                float r[8];
                unpack8(q4[c], r);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] += r[k];
              }
              if (p.has_raw) *raw.at(oy, j0 + c) = pack8(val);
              if constexpr (SFTG) {
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], sv[c * 8 + k], tv[c * 8 + k]);   // sv = scale + 1 (bias step)
              } else if (p.has_sft) {
                float s[8], tt[8];
                unpack8(s4[c], s);
                unpack8(t4[c], tt);
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], s[k], val[k]) + tt[k];   // x*(scale+1)+shift
              }
            }
            if (p.out_split > 0 && j0 + c >= p.out_split) *out2.at(oy, j0 + c - p.out_split) = pack8(val);
            else if constexpr (OQ) {
              if (!ou8) *out.at(oy, j0 + c) = oq.mode ? pack8_fq(val, oq) : pack8(val);
            } else {
              *out.at(oy, j0 + c) = pack8(val);
            }
          }
          if constexpr (OQ && CH >= 2) {
            if (ou8) {               // 16 channels = one uint8 entry; chunk index in units of 16 channels
#pragma unroll
              for (int c = 0; c < CH; c += 2) *out.at(oy, (j0 + c) >> 1) = pack16_u8(fin[c], fin[c + 1], oq);
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

inline size_t conv_sftg_bytes(const ConvParams& p, bool ps) {
  if (!p.wpk2) return 0;
  return ((p.w2_bytes + 127) & ~127) + (ps ? static_cast<size_t>(kSRingPS) * kSSlotBytesPS : static_cast<size_t>(kSRing) * kSSlotBytes);
}
inline size_t conv_smem_bytes(const ConvParams& p, bool ps = false) {
  const size_t sftg = conv_sftg_bytes(p, ps);
  return kSmemHeader + ((p.w_bytes + 127) & ~127) + static_cast<size_t>(p.ring) * p.slot_bytes + sftg;
}

}  // namespace hdrtv
