// Row-streaming implicit-GEMM convolution on tcgen05 / TMEM for the P8 activation layout.
//
// One CTA owns a strip of 128 output pixels (the UMMA M dimension) and walks down a band of output rows.
// Every input row of the strip is fetched ONCE by 1-D bulk TMA copies (one per channel-chunk plane) into a
// ring of shared-memory row slots; the 3x3 taps are then nothing but byte shifts of the K-major, un-swizzled
// UMMA operand descriptor over those resident rows (dx = +16 B, dy = next ring slot).  The accumulator of one
// output row (128 pixels x N output channels, fp32) lives in TMEM, double-buffered so that the epilogue warps
// drain row t while the single MMA-issuing thread already runs row t+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
//
// Replaces the cuDNN conv2d calls of the reference's eager path (Condition_arch.py:571-583,
// HDRUNet3T1_arch.py:160-205, arch_util.py:68-95) with bias / activation / residual / SFT / PixelShuffle fused
// into the epilogue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace hdrtv {

constexpr int kTileM = 128;
constexpr int kPlaneEntries = 136;                 // 128 + halo + pairing slack, multiple of 8
constexpr int kPlaneBytes = kPlaneEntries * 16;    // 2176
constexpr int kMaxSteps = 40;
constexpr int kMaxCopies = 16;
constexpr int kMaxRing = 8;
constexpr int kConvThreads = 192;

enum StoreMode : int { STORE_P8 = 0, STORE_PS = 1, STORE_PLANAR = 2 };

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
  int slot_bytes, ring;
  int n_steps;
  ConvStep steps[kMaxSteps];
  const uint4* wpk;
  int w_bytes;
  const float* bias;
  int Ho, Wo, band;
  int act;
  int has_res, has_res2, has_sft, has_raw;
  P8 res, res2, sft, out, raw;
  __half* planar;
  long planar_plane;
  int planar_W;
  int* err;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU) return v >= 0.f ? v : 0.1f * v;
  return v;
}

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

// Finish one 8-channel chunk of one output pixel: residual add, raw store, SFT modulation, store.
__device__ __forceinline__ void finish_chunk(const ConvParams& p, float* val, int y, int j, int x, int sft_chunks) {
  if (p.has_res) {
    float r[8];
    unpack8(reinterpret_cast<const uint4*>(p.res.base)[p.res.entry(y, j, x)], r);
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] += r[k];
  }
  if (p.has_res2) {
    float r[8];
    unpack8(reinterpret_cast<const uint4*>(p.res2.base)[p.res2.entry(y, j, x)], r);
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] += r[k];
  }
  if (p.has_raw) reinterpret_cast<uint4*>(p.raw.base)[p.raw.entry(y, j, x)] = pack8(val);
  if (p.has_sft) {
    float s[8], t[8];
    unpack8(reinterpret_cast<const uint4*>(p.sft.base)[p.sft.entry(y, j, x)], s);
    unpack8(reinterpret_cast<const uint4*>(p.sft.base)[p.sft.entry(y, j + sft_chunks, x)], t);
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] = fmaf(val[k], s[k], val[k]) + t[k];   // x*(scale+1)+shift
  }
  reinterpret_cast<uint4*>(p.out.base)[p.out.entry(y, j, x)] = pack8(val);
}

template <int N, int MODE>
__global__ void __launch_bounds__(kConvThreads) conv_p8_kernel(const __grid_constant__ ConvParams p) {
  constexpr uint32_t kTmemCols = (2 * N < 32) ? 32 : 2 * N;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (kMaxRing + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + i); };
  auto tempty_bar = [&](int i) { return bar0 + 8u * (2 * kMaxRing + 2 + i); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxRing + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxRing + 5));
  uint8_t* wsm = smem + 256;
  uint8_t* ring = wsm + ((p.w_bytes + 127) & ~127);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * kTileM;
  const int oy0 = blockIdx.y * p.band;
  const int nrows_out = min(p.band, p.Ho - oy0);
  const int nrows_in = (nrows_out - 1) * p.stride + p.ks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
    }
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, p.w_bytes);
      bulk_g2s(smem_u32(wsm), p.wpk, p.w_bytes, wfull_bar);
      const uint32_t row_tx = p.n_copies * p.copy_bytes;
      for (int q = 0; q < nrows_in; ++q) {
        const int slot = q % p.ring;
        mbar_wait(empty_bar(slot), ((q / p.ring) & 1) ^ 1, p.err, 1);
        mbar_expect_tx(full_bar(slot), row_tx);
        const uint4* src = p.in + (static_cast<long>(oy0) * p.stride + q + p.row_bias) * p.in_row_entries + x0 +
                           blockIdx.z * p.in_z_entries;
        const uint32_t dst = smem_u32(ring) + slot * p.slot_bytes;
        for (int c = 0; c < p.n_copies; ++c)
          bulk_g2s(dst + p.copies[c].dst_off, src + p.copies[c].src_off, p.copy_bytes, full_bar(slot));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(wfull_bar, 0, p.err, 2);
      constexpr uint32_t idesc = make_idesc_f16_m128(N);
      const uint32_t ring_base = smem_u32(ring), w_base = smem_u32(wsm);
      int waited = -1;
      for (int t = 0; t < nrows_out; ++t) {
        const int stage = t & 1;
        mbar_wait(tempty_bar(stage), ((t >> 1) & 1) ^ 1, p.err, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + stage * N;
        for (int s = 0; s < p.n_steps; ++s) {
          const ConvStep st = p.steps[s];
          const int q = t * p.stride + st.row;
          if (q > waited) {
            for (int r = waited + 1; r <= q; ++r) mbar_wait(full_bar(r % p.ring), (r / p.ring) & 1, p.err, 4);
            waited = q;
            tc_fence_after();
          }
          const uint32_t a_addr = ring_base + (q % p.ring) * p.slot_bytes + st.a_off;
          const uint64_t adesc = make_smem_desc(a_addr, st.a_lbo, 128);
          const uint64_t bdesc = make_smem_desc(w_base + s * (N * 32), N * 16, 128);
          tc_mma_f16(d_tmem, adesc, bdesc, idesc, s > 0 ? 1u : 0u);
          if (st.release) tc_commit(empty_bar(q % p.ring));
        }
        tc_commit(tfull_bar(stage));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int lg = warp & 3;
    const int x = p.xmul * (x0 + lg * 32 + lane) + blockIdx.z;
    const bool xin = x < p.Wo;
    for (int t = 0; t < nrows_out; ++t) {
      const int stage = t & 1;
      const int oy = oy0 + t;
      mbar_wait(tfull_bar(stage), (t >> 1) & 1, p.err, 5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + stage * N;

      if constexpr (MODE == STORE_PLANAR) {
        float v[16];
        tmem_ld16(taddr, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(stage));
        if (xin) {
          float val[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) val[k] = (k < 3) ? apply_act(v[k] + __ldg(p.bias + k), p.act) : 0.f;
          if (p.has_res) {
            float r[8];
            unpack8(reinterpret_cast<const uint4*>(p.res.base)[p.res.entry(oy, 0, x)], r);
#pragma unroll
            for (int k = 0; k < 3; ++k) val[k] += r[k];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k)
            p.planar[k * p.planar_plane + static_cast<long>(oy) * p.planar_W + x] = __float2half_rn(val[k]);
          if (p.has_raw) reinterpret_cast<uint4*>(p.raw.base)[p.raw.entry(oy, 0, x)] = pack8(val);
        }
      } else if constexpr (MODE == STORE_PS) {
        // N = 128 conv channels -> 32 channels at (2*oy+i, 2*x+j); conv channel n = 4*c + 2*i + j.
#pragma unroll 1
        for (int pass = 0; pass < N / 32; ++pass) {
          float v[32];
          tmem_ld32(taddr + pass * 32, v);
          if (pass == N / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(stage));
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = apply_act(v[k] + __ldg(p.bias + pass * 32 + k), p.act);
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) {
            const int Y = 2 * oy + (sub >> 1), X = 2 * x + (sub & 1);
            if (xin && Y < p.out.H && X < p.out.W) {
              float val[8];
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) val[cc] = v[4 * cc + sub];
              finish_chunk(p, val, Y, pass, X, N / 32);
            }
          }
        }
      } else {
        constexpr int kCols = (N < 32) ? N : 32;
#pragma unroll 1
        for (int pass = 0; pass < N / kCols; ++pass) {
          float v[kCols];
          if constexpr (kCols == 16) tmem_ld16(taddr + pass * kCols, v);
          else tmem_ld32(taddr + pass * kCols, v);
          if (pass == N / kCols - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(stage));
          }
          if (xin) {
#pragma unroll
            for (int ch = 0; ch < kCols / 8; ++ch) {
              const int j = pass * (kCols / 8) + ch;
              float val[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) val[k] = apply_act(v[ch * 8 + k] + __ldg(p.bias + j * 8 + k), p.act);
              finish_chunk(p, val, oy, j, x, N / 8);
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

inline size_t conv_smem_bytes(const ConvParams& p) {
  return 256 + ((p.w_bytes + 127) & ~127) + static_cast<size_t>(p.ring) * p.slot_bytes;
}

}  // namespace hdrtv
