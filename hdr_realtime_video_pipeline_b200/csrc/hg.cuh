// HG stage (highlight generation, the third HDRTVNet++ stage): Hallucination_Generator + HG_Composite
// (reference: src/models/hdrtvnet_modules/Hallucination_arch.py:53-137, HG_Composite_arch.py:77-107).
//
// The U-Net's convolutions have 64 ... 2048 output channels and up to 1024 input channels: unlike the 32/64-channel LE
// convs (row streaming with resident weights, conv_p8.cuh) neither a weight set (up to 18.9 MB) nor an input row
// (139 KB at 512 channels) fits in shared memory, so this is a K-STREAMED implicit GEMM:
//
//   gconv_kernel<KIND, NT, EPI>   persistent CTAs (one per SM), static round-robin over tiles of
//                                 4 output rows x 128 pixels x NT output channels (NT = 128: TMEM 4 x 128 columns = all of it).
//   K loop in groups of 16 input channels (3x3; 64 for 1x1): one pipeline stage = the A rows of the group
//   ((4 + 2) rows x 2 channel-chunk planes, 1-D bulk TMA, the P8 halo layout supplies the zero padding) + the
//   group's weights for all taps (9 x [16 x NT] K-major blocks, one bulk copy).  Per stage 36 MMAs of
//   M = 128, N = 128, K = 16 (64 tensor-pipe cycles each at the dense rate: N >= 128, DESIGN fact 1 / 13) stand against
//   63 KB of L2 -> SM traffic = 27 B/clk/SM, under the measured ~42 B/clk/SM L2 throughput cap: tensor-bound.
//   Accumulators roll: the MMA warp goes row by row (r = 0..3) inside every stage, so after a tile's last stage the
//   epilogue of row r overlaps the next tile's first stages on the rows already drained (per-row tfull / tempty barriers).
//   The bias is one more MMA step (constant [1,1,0..] A operand against [b_hi, b_lo] weights), appended to the last
//   stage's weight block.
//   Epilogues: ReLU -> P8 | ReLU + MaxPool2d(2) (+ optional un-pooled store: the skip connection) | PixelShuffle(2) + ReLU.
//   1x1 convs read the concatenation of two tensors (torch.cat((up, skip), 1)) as consecutive K groups of two sources.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..9 = epilogue
// (two per TMEM lane quadrant, each draining half of the NT columns).
#pragma once
#include "common.cuh"
#include "conv_p8.cuh"
#include "ptx.cuh"

namespace hdrtv {

enum GKind : int { G_3x3 = 0, G_3x3_C8 = 1, G_1x1 = 2 };
// GE_POOL_DOT / GE_PS_DOT: the layer's full-resolution output feeds only conv10 (1x1, 128 -> 3): instead of storing it
// (128 B/px written and read back), the epilogue multiplies the half-rounded activations with conv10's weights and stores
// the per-warp partial sums (3 floats per pixel and slice); hg_tail_dot_kernel adds the slices in a fixed order.
enum GEpi : int { GE_P8 = 0, GE_POOL = 1, GE_PS = 2, GE_POOL_DOT = 3, GE_PS_DOT = 4 };
constexpr int kGRows = 4;                       // output rows per tile (accumulators in TMEM)
constexpr int kHgNoMask = 0x7f7f7f7f;           // gate words after cudaMemset(0x7f): no masked pixel seen
constexpr int kHgCell = 64;                     // gate cell (full-resolution pixels)
constexpr int kHgCellShift = 6;
constexpr int kHgCellReach = 3;                 // floor(186 / 64) + 1: cells a dependency cone of 186 px can reach
constexpr int kGThreads = 320;

__host__ __device__ constexpr int g_planes(int k) { return k == G_3x3 ? 2 : (k == G_3x3_C8 ? 1 : 8); }      // per K group
__host__ __device__ constexpr int g_rows(int k, int rb = kGRows) { return k == G_1x1 ? rb : rb + 2; }               // A rows per stage
__host__ __device__ constexpr int g_steps(int k) { return k == G_3x3 ? 9 : (k == G_3x3_C8 ? 6 : 4); }       // MMAs per row and group
// C8: one K group per tile, 14-27 KB stages: prefetch 4 tiles deep.  (3x3 with two-row tiles: 57 KB stages, four would not fit)
__host__ __device__ constexpr int g_stages(int k, int rb = kGRows) { return k == G_3x3 ? 3 : (k == G_3x3_C8 ? 4 : 2); }
__host__ __device__ constexpr int g_group_channels(int k) { return k == G_3x3 ? 16 : (k == G_3x3_C8 ? 8 : 64); }
// A operand of step i for output row r: input-row slot and byte offset inside the stage's A block, K-half distance
__host__ __device__ constexpr int g_step_dy(int k, int i) { return k == G_3x3 ? i / 3 : (k == G_3x3_C8 ? i / 2 : 0); }
__host__ __device__ constexpr uint32_t g_step_off(int k, int i) {
  if (k == G_3x3) return static_cast<uint32_t>(i % 3) * 16u;
  if (k == G_3x3_C8) return (i % 2) ? 32u : 0u;                           // taps dx = 0,1 share one K = 16 step, dx = 2 + zero weights
  return static_cast<uint32_t>(2 * i) * kPlaneBytes + 16u;                // 1x1: planes 2i, 2i+1; +16 skips the halo entry
}
__host__ __device__ constexpr uint32_t g_step_lbo(int k) { return k == G_3x3_C8 ? 16u : static_cast<uint32_t>(kPlaneBytes); }
__host__ __device__ constexpr uint32_t g_a_bytes(int k, int rb = kGRows) { return static_cast<uint32_t>(g_rows(k, rb) * g_planes(k)) * kPlaneBytes; }
__host__ __device__ constexpr uint32_t g_b_bytes(int k, int nt) { return static_cast<uint32_t>(g_steps(k) + 1) * nt * 32u; }
__host__ __device__ constexpr uint32_t g_stage_bytes(int k, int nt, int rb = kGRows) { return g_a_bytes(k, rb) + g_b_bytes(k, nt); }
constexpr int kGHeader = 256 + kPlaneBytes + 128;                          // barriers, TMEM slot, constant "ones" operand
__host__ __device__ constexpr size_t g_smem_bytes(int k, int nt, int rb = kGRows) { return kGHeader + static_cast<size_t>(g_stages(k, rb)) * g_stage_bytes(k, nt, rb) + 128; }
static_assert(g_smem_bytes(G_3x3, 128, 4) <= 227 * 1024 && g_smem_bytes(G_3x3, 128, 2) <= 227 * 1024 && g_smem_bytes(G_1x1, 128, 4) <= 227 * 1024,
              "pipeline stages exceed shared memory");

struct GConvParams {
  const uint4* in0;          // source 0 (P8, natural layout)
  long in0_row_entries;
  uint32_t in0_wp;
  int in0_groups;            // K groups read from source 0; the remaining (kgroups - in0_groups) come from source 1
  const uint4* in1;
  long in1_row_entries;
  uint32_t in1_wp;
  int kgroups;
  const uint4* wpk;          // [ntile][kgroup][steps] blocks of NT*32 bytes, + one bias block after the last group of a tile
  long w_tile_bytes;
  int ntiles, strips, rowblocks, tiles;
  int H, W;                  // conv output size = input size (stride 1, same padding)
  int relu;
  P8 out;                    // GE_P8: H x W; GE_POOL: H/2 x W/2; GE_PS: 2H x 2W
  P8 out_full;               // GE_POOL: optional store of the un-pooled rows (skip connection)
  int has_full;
  const float* dot_w;        // *_DOT: conv10 weights of this layer's channels, [3][channels] (half-rounded values)
  float* dot_out;            // *_DOT: partial sums, [slices][3][dot_H][dot_W] planar fp32; slice = ntile * 2 + half
  int dot_H, dot_W;
  // Highlight gate (optional).  gate[0] = kHgNoMask while no pixel of the frame is inside the highlight mask: the stage's
  // output is then the base image itself (mask * hg + img with mask = 0) and the launch returns at once.  Otherwise
  // `gate_active` is a map of kHgCell x kHgCell full-resolution cells: 1 = within kHgCellReach cells (186 pixels: the reach of the
  // U-Net's dependency cone: five 2x poolings and their 3x3 convs) of a cell with a masked pixel.  Only tiles that touch
  // an active cell are computed: nothing else can reach a masked output, and unmasked outputs do not use the U-Net.
  const int* gate;
  const uint8_t* gate_active;
  int gate_cw, gate_ch;      // cells per row / column
  int lvl;                   // resolution level of this conv's pixels (0 = full resolution, 5 = 1/32)
  int* err;
};

// 8 half-rounded channel values of one pixel x conv10 weights of channels [ch0, ch0 + 8) -> 3 partial sums
__device__ __forceinline__ void dot3_acc(const float* val, const float* __restrict__ w, int channels, int ch0, float* acc) {
  float f[8];
  unpack8(pack8(val), f);                     // the tensor conv10 reads is a half tensor
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + k * channels + ch0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + k * channels + ch0 + 4));
    acc[k] = fmaf(f[0], w0.x, acc[k]); acc[k] = fmaf(f[1], w0.y, acc[k]); acc[k] = fmaf(f[2], w0.z, acc[k]); acc[k] = fmaf(f[3], w0.w, acc[k]);
    acc[k] = fmaf(f[4], w1.x, acc[k]); acc[k] = fmaf(f[5], w1.y, acc[k]); acc[k] = fmaf(f[6], w1.z, acc[k]); acc[k] = fmaf(f[7], w1.w, acc[k]);
  }
}

// RB = output rows per tile.  4: the accumulators fill TMEM (4 x NT columns) and roll row by row between consecutive tiles -
// least L2 traffic per MMA, right for deep K.  2: two tiles' accumulators alternate in TMEM (2 x 2 x NT columns), so a tile's
// epilogue overlaps the whole next tile and no stage is a row-by-row hand-over - right for the layers with 1-4 K groups per
// tile, where that hand-over stage is a quarter (or all) of the tile.
template <int KIND, int NT, int EPI, int RB = kGRows>
__global__ void __launch_bounds__(kGThreads, 1) gconv_kernel(const __grid_constant__ GConvParams p) {
  static_assert(NT == 64 || NT == 128, "N tile");
  static_assert(RB == 4 || (RB == 2 && KIND != G_1x1), "rows per tile");
  static_assert((EPI != GE_PS && EPI != GE_PS_DOT) || NT == 128, "PixelShuffle epilogue: 128 conv channels = 32 output channels per tile");
  constexpr int PL = g_planes(KIND), ROWS = g_rows(KIND, RB), NSTEPS = g_steps(KIND), S = g_stages(KIND, RB);
  constexpr uint32_t A_BYTES = g_a_bytes(KIND, RB), STAGE = g_stage_bytes(KIND, NT, RB), BLK = NT * 32u;
  if (p.gate != nullptr && *reinterpret_cast<const volatile int*>(p.gate) == kHgNoMask) return;      // uniform: before any barrier / TMEM
  // does the tile (x0 .. x0+127, y0 .. y0+RB-1 at this conv's level) touch an active gate cell?
  auto tile_on = [&](int x0, int y0) -> bool {
    if (p.gate == nullptr) return true;
    const int cx0 = (x0 << p.lvl) >> kHgCellShift, cx1 = min((((x0 + kTileM) << p.lvl) - 1) >> kHgCellShift, p.gate_cw - 1);
    const int cy0 = (y0 << p.lvl) >> kHgCellShift, cy1 = min((((y0 + RB) << p.lvl) - 1) >> kHgCellShift, p.gate_ch - 1);
    for (int cy = cy0; cy <= cy1; ++cy)
      for (int cx = cx0; cx <= cx1; ++cx)
        if (__ldg(p.gate_active + cy * p.gate_cw + cx)) return true;      // <= 2 KB map, written before this launch: L1-resident
    return false;
  };
  constexpr uint32_t kTmemCols = 4 * NT;                 // RB = 4: one tile of four rows; RB = 2: two tiles of two rows
  // accumulator block / barrier index of row r of this CTA's `it`-th tile, and the parity its barriers are in
  auto acc_idx = [](int it, int r) { return RB == 4 ? r : ((it & 1) * 2 + r); };
  auto acc_par = [](int it) { return static_cast<uint32_t>(RB == 4 ? (it & 1) : ((it >> 1) & 1)); };
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  const uint32_t bar0 = smem_u32(smem);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (4 + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (8 + i); };
  auto tempty_bar = [&](int i) { return bar0 + 8u * (12 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 16);
  uint8_t* ones = smem + 256;
  uint8_t* stage0 = smem + kGHeader;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); }
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kPlaneEntries) {
    reinterpret_cast<uint4*>(ones)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0u, 0u, 0u);      // [1, 1, 0, 0, 0, 0, 0, 0]
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int KG = p.kgroups;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t st = 0, ph = 1;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int nt = t % p.ntiles, rest = t / p.ntiles;
        const int x0 = (rest % p.strips) * kTileM, y0 = (rest / p.strips) * RB;
        if (!tile_on(x0, y0)) continue;
        const uint4* wt = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.wpk) + static_cast<long>(nt) * p.w_tile_bytes);
        for (int kg = 0; kg < KG; ++kg) {
          mbar_wait(empty_bar(st), ph, p.err, 21);
          const bool last = kg == KG - 1;
          const uint32_t b_bytes = (NSTEPS + (last ? 1 : 0)) * BLK;
          mbar_expect_tx(full_bar(st), A_BYTES + b_bytes);
          const uint32_t dst = smem_u32(stage0) + st * STAGE;
          bulk_g2s(dst + A_BYTES, wt + static_cast<long>(kg) * (NSTEPS * BLK / 16), b_bytes, full_bar(st));
          const bool s0 = kg < p.in0_groups;
          const uint4* base = s0 ? p.in0 : p.in1;
          const long row_entries = s0 ? p.in0_row_entries : p.in1_row_entries;
          const uint32_t wp = s0 ? p.in0_wp : p.in1_wp;
          const int g = s0 ? kg : kg - p.in0_groups;
          // first P8 row of the stage: 3x3 -> input row y0 - 1 = P8 row y0; 1x1 -> input row y0 = P8 row y0 + 1
          const uint4* src = base + static_cast<long>(y0 + (KIND == G_1x1 ? 1 : 0)) * row_entries + static_cast<long>(g) * PL * wp + x0;
#pragma unroll 1
          for (int q = 0; q < ROWS; ++q) {
#pragma unroll
            for (int pl = 0; pl < PL; ++pl) {
              unsigned long long a;
              asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(pl * wp), "l"(src));
              bulk_g2s(dst + (q * PL + pl) * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, full_bar(st));
            }
            src += row_entries;
          }
          if (++st == S) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    constexpr uint32_t idesc = make_idesc_f16_m128(NT);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
    auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(NT) << 16;   // (NT*16 bytes) >> 4 in the LBO field
    constexpr uint32_t b_step = static_cast<uint32_t>(NT) * 2;    // (NT*32 bytes) >> 4
    constexpr uint32_t a_lbo = (g_step_lbo(KIND) >> 4) << 16;
    uint32_t st = 0, ph = 0;
    int it = 0;
    const bool lead = elect_one();        // the issuing lane, elected once: the loop below stays warp-uniform around it
    // one accumulator row of one stage: 9 (6, 4) MMAs, + the bias step and the hand-over to the epilogue on the last stage
    auto issue_row = [&](int it_, int r, uint32_t a16, uint32_t b16, bool first, bool last) {
      const uint32_t d_tmem = tmem_base + acc_idx(it_, r) * NT;
      static_for<0, NSTEPS>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int dy = g_step_dy(KIND, i);
        constexpr uint32_t off16 = g_step_off(KIND, i) >> 4;
        const uint32_t arow = a16 + static_cast<uint32_t>((r + dy) * PL) * (kPlaneBytes >> 4);
        tc_mma_f16(d_tmem, mkdesc((arow + off16) | a_lbo), mkdesc(b16 + i * b_step), idesc, (first && i == 0) ? 0u : 1u);
      });
      if (last) {
        tc_mma_f16(d_tmem, ones_desc, mkdesc(b16 + NSTEPS * b_step), idesc, 1u);      // + bias
        tc_commit(tfull_bar(acc_idx(it_, r)));
      }
    };
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      {
        const int rest = t / p.ntiles;
        if (!tile_on((rest % p.strips) * kTileM, (rest / p.strips) * RB)) continue;      // same decision in all three roles
      }
      for (int kg = 0; kg < KG; ++kg) {
        const uint32_t a16 = (smem_u32(stage0) + st * STAGE) >> 4;
        const uint32_t b16 = ((smem_u32(stage0) + st * STAGE + A_BYTES) >> 4) | b_lbo;
        const bool first = kg == 0, last = kg == KG - 1;
        mbar_wait(full_bar(st), ph, p.err, 22);
        tc_fence_after();
        if (first) {
          // the previous tile's epilogue must have drained accumulator r: row by row, so that the first rows of this tile
          // already run while the last rows of the previous one are still being read
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            mbar_wait(tempty_bar(acc_idx(it, r)), acc_par(it) ^ 1, p.err, 23);
            tc_fence_after();
            if (lead) issue_row(it, r, a16, b16, true, last);
            __syncwarp();
          }
        } else {
          // steady state: all 36 MMAs of the stage in one straight run of the issuing lane (nothing between them but
          // descriptor arithmetic: the pipe's queue is shallow, every detour of the issuer is idle pipe time)
          if (lead) {
#pragma unroll
            for (int r = 0; r < RB; ++r) issue_row(it, r, a16, b16, false, last);
          }
          __syncwarp();
        }
        if (lead) tc_commit(empty_bar(st));
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1; }
      }
      ++it;                                   // tiles this CTA has processed (barrier phases), not tiles it has looked at
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quadrant
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = NT / 2;          // columns drained by this warp
    constexpr int CH = COLS / 8;          // 8-channel chunks per thread
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + half * COLS;
    const bool relu = p.relu != 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      const int nt = t % p.ntiles, rest = t / p.ntiles;
      const int x = (rest % p.strips) * kTileM + lg * 32 + lane, y0 = (rest / p.strips) * RB;
      if (!tile_on((rest % p.strips) * kTileM, y0)) continue;
      const uint32_t tpar = acc_par(it);
      const int ab = acc_idx(it, 0);                   // first accumulator block / barrier of this tile
      if constexpr (EPI == GE_P8) {
        ColRef out;
        out.init(p.out, x);
        const int j0 = nt * (NT / 8) + half * CH;
        // rows are drained in pairs before anything is stored: an accumulator row goes back to the MMA warp as soon as it
        // is in registers, not after the previous row's convert / address / store chain
#pragma unroll 1
        for (int rp = 0; rp < RB / 2; ++rp) {
          float va[COLS], vb[COLS];
          mbar_wait(tfull_bar(ab + 2 * rp), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<COLS>(tlane + (ab + 2 * rp) * NT, va);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp));
          mbar_wait(tfull_bar(ab + 2 * rp + 1), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<COLS>(tlane + (ab + 2 * rp + 1) * NT, vb);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp + 1));
          const int y = y0 + 2 * rp;
          if (x < p.W) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              float val[8];
              if (y < p.H) {
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = relu ? fmaxf(va[c * 8 + k], 0.f) : va[c * 8 + k];
                *out.at(y, j0 + c) = pack8(val);
              }
              if (y + 1 < p.H) {
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = relu ? fmaxf(vb[c * 8 + k], 0.f) : vb[c * 8 + k];
                *out.at(y + 1, j0 + c) = pack8(val);
              }
            }
          }
        }
      } else if constexpr (EPI == GE_POOL || EPI == GE_POOL_DOT) {
        ColRef out, full;
        out.init(p.out, x >> 1);
        if (EPI == GE_POOL && p.has_full) full.init(p.out_full, x);
        const int j0 = nt * (NT / 8) + half * CH;
#pragma unroll 1
        for (int rp = 0; rp < RB / 2; ++rp) {
          float a[COLS], b[COLS];
          mbar_wait(tfull_bar(ab + 2 * rp), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<COLS>(tlane + (ab + 2 * rp) * NT, a);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp));
          mbar_wait(tfull_bar(ab + 2 * rp + 1), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<COLS>(tlane + (ab + 2 * rp + 1) * NT, b);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp + 1));
          const int y = y0 + 2 * rp;
          const bool in = x < p.W && y < p.H;        // H, W even: row y + 1 and column x ^ 1 are inside with (y, x)
#pragma unroll
          for (int k = 0; k < COLS; ++k) {
            if (relu) { a[k] = fmaxf(a[k], 0.f); b[k] = fmaxf(b[k], 0.f); }
          }
          if constexpr (EPI == GE_POOL_DOT) {
            if (in) {
              float da[3] = {0.f, 0.f, 0.f}, db[3] = {0.f, 0.f, 0.f};
#pragma unroll
              for (int c = 0; c < CH; ++c) {
                dot3_acc(a + 8 * c, p.dot_w, p.ntiles * NT, (j0 + c) * 8, da);
                dot3_acc(b + 8 * c, p.dot_w, p.ntiles * NT, (j0 + c) * 8, db);
              }
              const long plane = static_cast<long>(p.dot_H) * p.dot_W;
              float* o = p.dot_out + static_cast<long>(nt * 2 + half) * 3 * plane + static_cast<long>(y) * p.dot_W + x;
#pragma unroll
              for (int k = 0; k < 3; ++k) { o[k * plane] = da[k]; o[k * plane + p.dot_W] = db[k]; }
            }
          } else if (p.has_full && in) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              *full.at(y, j0 + c) = pack8(a + 8 * c);
              *full.at(y + 1, j0 + c) = pack8(b + 8 * c);
            }
          }
#pragma unroll
          for (int k = 0; k < COLS; ++k) {
            float m = fmaxf(a[k], b[k]);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            a[k] = m;
          }
          if (in && (lane & 1) == 0) {
#pragma unroll
            for (int c = 0; c < CH; ++c) *out.at(y >> 1, j0 + c) = pack8(a + 8 * c);
          }
        }
      } else {
        // PixelShuffle(2) + ReLU: conv channel n = 4*c + 2*i + j -> output channel c at (2y + i, 2x + j).
        // This warp: conv channels [nt*128 + 64*half, +64) = output chunks nt*4 + 2*half, +1 of all four sub-pixels.
        ColRef out[2];
        if constexpr (EPI == GE_PS) {
          out[0].init(p.out, 2 * x);
          out[1].init(p.out, 2 * x + 1);
        }
        const int j0 = nt * 4 + half * 2;
#pragma unroll 1
        for (int rp = 0; rp < RB / 2; ++rp) {
          float vv[2][64];                             // two rows drained before either is shuffled out (see GE_P8)
          mbar_wait(tfull_bar(ab + 2 * rp), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<64>(tlane + (ab + 2 * rp) * NT, vv[0]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp));
          mbar_wait(tfull_bar(ab + 2 * rp + 1), tpar, p.err, 24);
          tc_fence_after();
          tmem_ld_cols<64>(tlane + (ab + 2 * rp + 1) * NT, vv[1]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab + 2 * rp + 1));
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float* v = vv[rr];
            const int y = y0 + 2 * rp + rr;
            if (x < p.W && y < p.H) {
#pragma unroll
              for (int sub = 0; sub < 4; ++sub) {
                float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  float val[8];
#pragma unroll
                  for (int cc = 0; cc < 8; ++cc) {
                    const float u = v[32 * c + 4 * cc + sub];
                    val[cc] = relu ? fmaxf(u, 0.f) : u;
                  }
                  if constexpr (EPI == GE_PS) *out[sub & 1].at(2 * y + (sub >> 1), j0 + c) = pack8(val);
                  else dot3_acc(val, p.dot_w, p.ntiles * 32, (j0 + c) * 8, d);
                }
                if constexpr (EPI == GE_PS_DOT) {
                  const long plane = static_cast<long>(p.dot_H) * p.dot_W;
                  float* o = p.dot_out + static_cast<long>(nt * 2 + half) * 3 * plane + static_cast<long>(2 * y + (sub >> 1)) * p.dot_W + 2 * x + (sub & 1);
#pragma unroll
                  for (int k = 0; k < 3; ++k) o[k * plane] = d[k];
                }
              }
            }
          }
        }
      }
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Stage-in: the base model's output (planar, model dtype) -> single-chunk P8 image of the padded size (Hp, Wp multiples
// of 32), reflect padding on the right / bottom (F.pad(mode="reflect"), HG_Composite_arch.py:94-101: index n + i reads
// n - 2 - i).  Channels 3..7 of an entry are zero.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hg_mask_half(float m);
template <typename T>
__global__ void __launch_bounds__(128) hg_stage_in_kernel(const T* __restrict__ src, P8 dst, int H, int W, int Hp, int Wp, int* gate,
                                                          uint8_t* seen, int cw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  bool masked = false;
  if (x < Wp && y < Hp) {
    const int sx = x < W ? x : 2 * W - 2 - x, sy = y < H ? y : 2 * H - 2 - y;
    const long plane = static_cast<long>(H) * W, o = static_cast<long>(sy) * W + sx;
    float v[8] = {static_cast<float>(src[o]), static_cast<float>(src[plane + o]), static_cast<float>(src[2 * plane + o]), 0.f, 0.f, 0.f, 0.f, 0.f};
    reinterpret_cast<uint4*>(dst.base)[dst.entry(y, 0, x)] = pack8(v);
    masked = x < W && y < H && hg_mask_half(fmaxf(v[0], fmaxf(v[1], v[2]))) != 0.f;
  }
  // gate: which cells hold a masked pixel (un-padded area: only those outputs exist); gate[0] leaves kHgNoMask with the first one
  if (gate != nullptr && masked) {
    seen[(y >> kHgCellShift) * cw + (x >> kHgCellShift)] = 1;
    *gate = 0;
  }
}
// active cell = within kHgCellReach cells of a cell that holds a masked pixel (one small block per frame)
__global__ void hg_gate_dilate_kernel(const uint8_t* __restrict__ seen, uint8_t* __restrict__ active, int cw, int ch) {
  for (int i = threadIdx.x; i < cw * ch; i += blockDim.x) {
    const int cx = i % cw, cy = i / cw;
    uint8_t a = 0;
    for (int dy = -kHgCellReach; dy <= kHgCellReach; ++dy)
      for (int dx = -kHgCellReach; dx <= kHgCellReach; ++dx) {
        const int yy = cy + dy, xx = cx + dx;
        if (yy >= 0 && yy < ch && xx >= 0 && xx < cw && seen[yy * cw + xx]) a = 1;
      }
    active[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tail: conv10 (1x1, cat(Up_conv5 output, conv1_out) 128 -> 3), conv_last (1x1, cat(conv10_out, img) 6 -> 3), the mask
// blend  out = mask * hg + img  (Hallucination_arch.py:131-137) and the crop to H x W, one thread per pixel.
// FP16 model semantics: conv10 / conv_last outputs are half tensors (rounded here), the mask is a FLOAT tensor, so
// mask * out + img promotes to fp32 - the reference's HG output is float32 in both precisions.
// Mask (HG_Composite_arch.py:77-84): ((max_c(img) - r) / (1 - r)).clamp(0, 1) > 0.1 evaluated in the model dtype.
// ---------------------------------------------------------------------------------------------------------------------
struct HgTail {
  float w10[3][128];
  float b10[3];
  float wl[3][6];
  float bl[3];
};
__device__ __forceinline__ float hg_mask_half(float m) {            // m = max_c(img), a half value
  const __half d = __float2half_rn(m - 0.75f);                      // half tensor - python scalar: computed in fp32, rounded to half
  __half q = __float2half_rn(__half2float(d) / 0.25f);
  float qf = fminf(fmaxf(__half2float(q), 0.f), 1.f);
  return qf > __half2float(__float2half_rn(0.1f)) ? 1.f : 0.f;      // the scalar 0.1 is cast to the tensor's dtype
}
__device__ __forceinline__ float hg_mask_f32(float m) {
  float q = __fdiv_rn(__fsub_rn(m, 0.75f), 0.25f);
  q = fminf(fmaxf(q, 0.f), 1.f);
  return q > 0.1f ? 1.f : 0.f;
}
__global__ void __launch_bounds__(128) hg_tail_kernel(P8 up, P8 skip, P8 img, const HgTail* __restrict__ tw, float* __restrict__ out, int H, int W) {
  __shared__ HgTail s;
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(HgTail) / 4); i += blockDim.x) reinterpret_cast<float*>(&s)[i] = reinterpret_cast<const float*>(tw)[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  float acc[3] = {s.b10[0], s.b10[1], s.b10[2]};
#pragma unroll
  for (int src = 0; src < 2; ++src) {
    const P8& t = src ? skip : up;
    const uint4* base = reinterpret_cast<const uint4*>(t.base);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
      unpack8(__ldcg(base + t.entry(y, j, x)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ci = src * 64 + j * 8 + k;
        acc[0] = fmaf(f[k], s.w10[0][ci], acc[0]);
        acc[1] = fmaf(f[k], s.w10[1][ci], acc[1]);
        acc[2] = fmaf(f[k], s.w10[2][ci], acc[2]);
      }
    }
  }
  float im[8];
  unpack8(__ldcg(reinterpret_cast<const uint4*>(img.base) + img.entry(y, 0, x)), im);
  float c10[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) c10[k] = __half2float(__float2half_rn(acc[k]));
  const float mask = hg_mask_half(fmaxf(im[0], fmaxf(im[1], im[2])));
  const long plane = static_cast<long>(H) * W, o = static_cast<long>(y) * W + x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = s.bl[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) v = fmaf(c10[i], s.wl[k][i], v);
#pragma unroll
    for (int i = 0; i < 3; ++i) v = fmaf(im[i], s.wl[k][3 + i], v);
    v = __half2float(__float2half_rn(v));
    out[k * plane + o] = __fadd_rn(__fmul_rn(mask, v), im[k]);
  }
}

// The same tail when conv10's two halves were reduced by the producers' epilogues (GE_POOL_DOT / GE_PS_DOT): `part` holds
// `slices` partial sums per pixel and output channel, [slices][3][Hp][Wp]; they are added in slice order (bit-reproducible).
__global__ void __launch_bounds__(128) hg_tail_dot_kernel(const float* __restrict__ part, int slices, int Hp, int Wp, P8 img,
                                                          const HgTail* __restrict__ tw, float* __restrict__ out, int H, int W,
                                                          const int* gate) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  if (gate != nullptr && *gate == kHgNoMask) { // no highlight anywhere: mask * hg + img = img (the U-Net did not run)
    float im0[8];
    unpack8(__ldcg(reinterpret_cast<const uint4*>(img.base) + img.entry(y, 0, x)), im0);
    const long pl = static_cast<long>(H) * W, oo = static_cast<long>(y) * W + x;
    out[oo] = im0[0]; out[pl + oo] = im0[1]; out[2 * pl + oo] = im0[2];
    return;
  }
  const long pp = static_cast<long>(Hp) * Wp, po = static_cast<long>(y) * Wp + x;
  float acc[3] = {__ldg(&tw->b10[0]), __ldg(&tw->b10[1]), __ldg(&tw->b10[2])};
  for (int s = 0; s < slices; ++s) {
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] += __ldcg(part + (static_cast<long>(s) * 3 + k) * pp + po);
  }
  float im[8];
  unpack8(__ldcg(reinterpret_cast<const uint4*>(img.base) + img.entry(y, 0, x)), im);
  float c10[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) c10[k] = __half2float(__float2half_rn(acc[k]));
  const float mask = hg_mask_half(fmaxf(im[0], fmaxf(im[1], im[2])));
  const long plane = static_cast<long>(H) * W, o = static_cast<long>(y) * W + x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = __ldg(&tw->bl[k]);
#pragma unroll
    for (int i = 0; i < 3; ++i) v = fmaf(c10[i], __ldg(&tw->wl[k][i]), v);
#pragma unroll
    for (int i = 0; i < 3; ++i) v = fmaf(im[i], __ldg(&tw->wl[k][3 + i]), v);
    v = __half2float(__float2half_rn(v));
    // outside the mask the partial sums may come from tiles the gate skipped (stale, possibly not even finite): not used
    out[k * plane + o] = __fadd_rn(__fmul_rn(mask, mask != 0.f ? v : 0.f), im[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// FP32 parity path (precision = "fp32"): planar fp32 CUDA-core kernels, <= 1e-4 against the reference's FP32 output.
// ---------------------------------------------------------------------------------------------------------------------
struct HgConvF32 {
  const float* in0; int C0;       // cat((in0, in1), 1): channels [0, C0) from in0, [C0, C0 + C1) from in1
  const float* in1; int C1;
  const float* w;                 // [Cout][C0 + C1][ks][ks] (BatchNorm folded)
  const float* b;
  float* out;                     // [Cout][H][W]; ps: [Cout/4][2H][2W]
  int Cout, H, W, ks, relu, ps;
};
constexpr int kHgF32Chunk = 32;   // input channels staged per pass
template <int COB, int PXT>
__global__ void __launch_bounds__(64) hg_conv_f32_kernel(const HgConvF32 p) {
  extern __shared__ __align__(16) float wsm[];   // [chunk*taps][COB]
  const int co0 = blockIdx.z * COB, taps = p.ks * p.ks, Cin = p.C0 + p.C1, pad = p.ks / 2;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * PXT, oy = blockIdx.y;
  const long plane = static_cast<long>(p.H) * p.W;
  float acc[PXT][COB];
#pragma unroll
  for (int q = 0; q < PXT; ++q)
#pragma unroll
    for (int c = 0; c < COB; ++c) acc[q][c] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += kHgF32Chunk) {
    const int cn = min(kHgF32Chunk, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * taps * COB; i += blockDim.x) {
      const int k = i / COB, c = i % COB;
      wsm[i] = (co0 + c < p.Cout) ? p.w[(static_cast<long>(co0 + c) * Cin + c0) * taps + k] : 0.f;
    }
    __syncthreads();
    if (ox0 < p.W) {
      for (int ci = 0; ci < cn; ++ci) {
        const int cg = c0 + ci;
        const float* ip = cg < p.C0 ? p.in0 + cg * plane : p.in1 + (cg - p.C0) * plane;
        for (int ky = 0; ky < p.ks; ++ky) {
          const int iy = oy + ky - pad;
          if (iy < 0 || iy >= p.H) continue;
          const float* row = ip + static_cast<long>(iy) * p.W;
          float v[PXT + 2];
#pragma unroll
          for (int j = 0; j < PXT + 2; ++j) {
            const int ix = ox0 - pad + j;
            v[j] = (j < PXT - 1 + p.ks && ix >= 0 && ix < p.W) ? __ldg(row + ix) : 0.f;
          }
          const float* wrow = wsm + (ci * p.ks + ky) * p.ks * COB;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (kx < p.ks) {
              const float4* wp = reinterpret_cast<const float4*>(wrow + kx * COB);
#pragma unroll
              for (int c = 0; c < COB / 4; ++c) {
                const float4 w4 = wp[c];
#pragma unroll
                for (int q = 0; q < PXT; ++q) {
                  const float xv = v[q + kx];
                  acc[q][4 * c + 0] = fmaf(xv, w4.x, acc[q][4 * c + 0]);
                  acc[q][4 * c + 1] = fmaf(xv, w4.y, acc[q][4 * c + 1]);
                  acc[q][4 * c + 2] = fmaf(xv, w4.z, acc[q][4 * c + 2]);
                  acc[q][4 * c + 3] = fmaf(xv, w4.w, acc[q][4 * c + 3]);
                }
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PXT; ++q) {
    const int ox = ox0 + q;
    if (ox >= p.W) break;
#pragma unroll
    for (int c = 0; c < COB; ++c) {
      const int co = co0 + c;
      if (co >= p.Cout) break;
      float v = acc[q][c] + __ldg(p.b + co);
      if (p.relu) v = fmaxf(v, 0.f);
      long o;
      if (p.ps) o = (static_cast<long>(co >> 2) * (2 * p.H) + 2 * oy + ((co & 3) >> 1)) * (2 * p.W) + 2 * ox + (co & 1);
      else o = static_cast<long>(co) * plane + static_cast<long>(oy) * p.W + ox;
      p.out[o] = v;
    }
  }
}
__global__ void hg_maxpool_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long n = static_cast<long>(C) * Ho * Wo, i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = static_cast<int>(i % Wo), y = static_cast<int>((i / Wo) % Ho);
  const long c = i / (static_cast<long>(Wo) * Ho);
  const float* s = in + (c * H + 2 * y) * W + 2 * x;
  out[i] = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[W], s[W + 1]));
}
__global__ void hg_reflect_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Hp, int Wp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
  if (x >= Wp) return;
  const int sx = x < W ? x : 2 * W - 2 - x, sy = y < H ? y : 2 * H - 2 - y;
  dst[(static_cast<long>(c) * Hp + y) * Wp + x] = src[(static_cast<long>(c) * H + sy) * W + sx];
}
// out = mask * hg + img, cropped to H x W (hg, img at the padded size)
__global__ void hg_blend_f32_kernel(const float* __restrict__ hg, const float* __restrict__ img, float* __restrict__ out, int H, int W, int Hp, int Wp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const long pp = static_cast<long>(Hp) * Wp, o = static_cast<long>(y) * Wp + x, plane = static_cast<long>(H) * W;
  const float i0 = img[o], i1 = img[pp + o], i2 = img[2 * pp + o];
  const float mask = hg_mask_f32(fmaxf(i0, fmaxf(i1, i2)));
  const long q = static_cast<long>(y) * W + x;
  out[q] = __fadd_rn(__fmul_rn(mask, hg[o]), i0);
  out[plane + q] = __fadd_rn(__fmul_rn(mask, hg[pp + o]), i1);
  out[2 * plane + q] = __fadd_rn(__fmul_rn(mask, hg[2 * pp + o]), i2);
}

}  // namespace hdrtv
