// Fused per-pixel layer chains on tcgen05: conv (3x3 / 1x1 from the row ring) followed by up to five 1x1 layers whose
// 64-channel intermediates never leave the SM.
//
//   AGCM MLP        : 1x1 3->64 ReLU, 1x1 64->64 ReLU, 1x1 64->3            (Condition_arch.py:571-583, GFM folded)
//   LE cond pyramid : 3x3 3->64, 1x1, 1x1 [store cond], 1x1, 1x1, 1x1 ->16 [store cond1], LeakyReLU(0.1)
//                                                                          (HDRUNet3T1_arch.py:41-46, 160-161)
//
// Per output row (128 pixels = UMMA M) the layers run strictly in sequence:
//   MMA(l) -> TMEM accumulator -> epilogue warps (activation, fp16) -> shared-memory operand tile -> MMA(l+1) ...
// Two rows are in flight per CTA ("ping-pong"): epilogue warps 2..5 own even rows, warps 6..9 odd rows, each with its
// own TMEM accumulator and operand tile, while the single MMA-issuing thread alternates between them, so the tensor
// pipe works on one row while the other row's activations are being converted.
#pragma once
#include "conv_p8.cuh"

namespace hdrtv {

constexpr int kMaxChain = 6;
constexpr int kTileBytes = 8 * kPlaneBytes;      // one 64-channel operand tile (8 channel-chunk planes)

struct ChainLayer {
  int n_steps;   // K/16 MMAs (layer 0: the input side's tap steps)
  int N;         // 64, or 16 for a last layer
  int w_off;     // byte offset of this layer's packed weights (tap steps + bias step) in the chain weight buffer
  int store;     // 0 none, 1 P8, 2 planar fp16 (3 channels) + P8 single chunk
  float slope;
  P8 out;
};

struct ChainParams {
  ConvParams base;           // input side (ring geometry, layer-0 steps), weights pointer / total bytes, Ho/Wo/band
  int n_layers;
  ChainLayer layers[kMaxChain];
};

__device__ __forceinline__ uint32_t act_half2(uint32_t h2, float slope) {
  __half2 v = *reinterpret_cast<__half2*>(&h2);
  const __half2 s = __float2half2_rn(slope);
  v = __hmax2(v, __hmul2(v, s));
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kConvThreads, 2) chain_p8_kernel(const __grid_constant__ ChainParams cp) {
  const ConvParams& p = cp.base;
  constexpr uint32_t kTmemCols = 128;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kMaxRing + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + i); };
  auto afull_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + 2 + i); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxRing + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxRing + 5));
  uint8_t* ones = smem + 512;
  uint8_t* wsm = smem + kSmemHeader;
  uint8_t* tiles = wsm + ((p.w_bytes + 127) & ~127);
  uint8_t* ring = tiles + 2 * kTileBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * kTileM;
  const int oy0 = blockIdx.y * p.band;
  const int nrows_out = min(p.band, p.Ho - oy0);
  const int nrows_in = (nrows_out - 1) * p.stride + p.ks;
  const int L = cp.n_layers;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(afull_bar(i), 4);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (same as conv_p8_kernel)
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, p.w_bytes);
      bulk_g2s(smem_u32(wsm), p.wpk, p.w_bytes, wfull_bar);
      const uint32_t row_tx = p.n_copies * p.copy_bytes;
      int slot = 0, ph = 1;
      const uint4* src = p.in + (static_cast<long>(oy0) * p.stride + p.row_bias) * p.in_row_entries + x0;
      for (int q = 0; q < nrows_in; ++q) {
        mbar_wait(empty_bar(slot), ph, p.err, 11);
        mbar_expect_tx(full_bar(slot), row_tx);
        const uint32_t dst = smem_u32(ring) + slot * p.slot_bytes;
        for (int c = 0; c < p.n_copies; ++c)
          bulk_g2s(dst + p.copies[c].dst_off, src + p.copies[c].src_off, p.copy_bytes, full_bar(slot));
        src += p.in_row_entries;
        if (++slot == p.ring) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer, alternating between two rows
    uint2* dsc = reinterpret_cast<uint2*>(smem + 8 * (2 * kMaxRing + 6));   // layer-0 {a_lo, b_lo}
    const uint32_t ring_base = smem_u32(ring), w_base = smem_u32(wsm), tile_base = smem_u32(tiles);
    const int N0 = cp.layers[0].N;
    for (int s = lane; s < p.n_steps; s += 32) {
      const ConvStep st = p.steps[s];
      dsc[s] = make_uint2(((st.a_off >> 4) & 0x3FFF) | (((st.a_lbo >> 4) & 0x3FFF) << 16),
                          (((w_base + cp.layers[0].w_off + s * (N0 * 32)) >> 4) & 0x3FFF) | ((((N0 * 16) >> 4) & 0x3FFF) << 16));
    }
    __syncwarp();
    if (lane == 0) {
      mbar_wait(wfull_bar, 0, p.err, 12);
      constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
      auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
      const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);
      const int spd = p.n_steps / p.ks;
      const uint32_t slot16 = p.slot_bytes >> 4, ring16 = ring_base >> 4;
      int waited = -1;
      int base_slot = 0, base_ph = 0;
      for (int tp = 0; tp < nrows_out; tp += 2) {
        for (int l = 0; l < L; ++l) {
          const ChainLayer& ly = cp.layers[l];
          const uint32_t idesc = make_idesc_f16_m128(ly.N);
          const uint32_t b_lbo = (((ly.N * 16) >> 4) & 0x3FFF) << 16;
          for (int g = 0; g < 2; ++g) {
            const int t = tp + g;
            if (t >= nrows_out) break;
            const int e = (t >> 1) * L + l;             // event index of this (row, layer) in slot g
            if (e > 0) {                                // previous event's epilogue done: TMEM free, tile written
              mbar_wait(afull_bar(g), (e - 1) & 1, p.err, 13);
              tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + g * 64;
            uint32_t acc = 0;
            if (l == 0) {
              int slot = base_slot, ph = base_ph;
              for (int dy = 0; dy < p.ks; ++dy) {
                const int q = t * p.stride + dy;
                if (q > waited) {
                  mbar_wait(full_bar(slot), ph, p.err, 14);
                  waited = q;
                  tc_fence_after();
                }
                const uint32_t a16 = ring16 + slot * slot16;
                const uint2* d = dsc + dy * spd;
#pragma unroll 2
                for (int i = 0; i < spd; ++i) {
                  const uint2 lo = d[i];
                  tc_mma_f16(d_tmem, mkdesc(a16 + lo.x), mkdesc(lo.y), idesc, acc);
                  acc = 1;
                }
                if (dy < p.stride) tc_commit(empty_bar(slot));
                if (++slot == p.ring) { slot = 0; ph ^= 1; }
              }
              base_slot += p.stride;
              if (base_slot >= p.ring) { base_slot -= p.ring; base_ph ^= 1; }
            } else {
              const uint32_t a_lo0 = ((tile_base + g * kTileBytes) >> 4) | (((kPlaneBytes >> 4) & 0x3FFF) << 16);
              const uint32_t b_lo0 = ((w_base + ly.w_off) >> 4) | b_lbo;
#pragma unroll 4
              for (int i = 0; i < ly.n_steps; ++i) {
                tc_mma_f16(d_tmem, mkdesc(a_lo0 + i * ((2 * kPlaneBytes) >> 4)), mkdesc(b_lo0 + i * ((ly.N * 32) >> 4)), idesc, acc);
                acc = 1;
              }
            }
            const uint32_t bias_lo = ((w_base + ly.w_off + ly.n_steps * (ly.N * 32)) >> 4) | b_lbo;
            tc_mma_f16(d_tmem, ones_desc, mkdesc(bias_lo), idesc, 1u);
            tc_commit(tfull_bar(g));
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group g owns rows t = g, g+2, ...
    const int g = (warp - 2) >> 2;
    const int lg = warp & 3;
    const int m = lg * 32 + lane;                    // pixel inside the strip = TMEM lane
    const int x = x0 + m;
    const bool xin = x < p.Wo;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + g * 64;
    uint4* tile = reinterpret_cast<uint4*>(tiles + g * kTileBytes) + m;
    ColRef outs[kMaxChain];
    for (int l = 0; l < L; ++l)
      if (cp.layers[l].store) outs[l].init(cp.layers[l].out, x);
    int e = 0;
    for (int t = g; t < nrows_out; t += 2) {
      const int oy = oy0 + t;
      for (int l = 0; l < L; ++l, ++e) {
        const ChainLayer& ly = cp.layers[l];
        mbar_wait(tfull_bar(g), e & 1, p.err, 15);
        tc_fence_after();
        if (ly.N == 64) {
          float v[64];
          tmem_ld_cols<64>(tlane, v);
          tc_fence_before();
          uint4 h[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              __half2 hh = __floats2half2_rn(v[c * 8 + 2 * k], v[c * 8 + 2 * k + 1]);
              w[k] = act_half2(*reinterpret_cast<uint32_t*>(&hh), ly.slope);
            }
            h[c] = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if (l + 1 < L) {
#pragma unroll
            for (int c = 0; c < 8; ++c) tile[c * kPlaneEntries] = h[c];
            fence_proxy_async_smem();
          }
          if (ly.store == 1 && xin) {
#pragma unroll
            for (int c = 0; c < 8; ++c) *outs[l].at(oy, c) = h[c];
          }
        } else {   // N == 16
          float v[16];
          tmem_ld_cols<16>(tlane, v);
          tc_fence_before();
          if (xin) {
            if (ly.store == 1) {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                float a[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = fmaxf(v[c * 8 + k], ly.slope * v[c * 8 + k]);
                *outs[l].at(oy, c) = pack8(a);
              }
            } else if (ly.store == 2) {
              float a[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) a[k] = k < 3 ? fmaxf(v[k], ly.slope * v[k]) : 0.f;
#pragma unroll
              for (int k = 0; k < 3; ++k)
                p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = __float2half_rn(a[k]);
              *outs[l].at(oy, 0) = pack8(a);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(afull_bar(g));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

inline size_t chain_smem_bytes(const ChainParams& cp) {
  return kSmemHeader + ((cp.base.w_bytes + 127) & ~127) + 2 * kTileBytes + static_cast<size_t>(cp.base.ring) * cp.base.slot_bytes;
}

}  // namespace hdrtv
