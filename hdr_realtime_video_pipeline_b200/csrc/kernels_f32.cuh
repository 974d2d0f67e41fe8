// FP32 CUDA-core kernels: the precision="fp32" product path (BASELINE config 1, <= 1e-4 vs the reference's
// FP32 output) and the AGCM condition classifier (always FP32, both precisions).
// Activations are planar NCHW fp32.  Reference: Condition_arch.py:8-35, 559-585; HDRUNet3T1_arch.py:152-206.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace hdrtv {

// ------------------------------------------------------------------------------------------------
// Direct convolution, planar fp32.  One thread = one output pixel x COB output channels.
// Weights for the block's COB channels are staged in shared memory as [cin][tap][COB].
// Epilogue: +bias, activation, optional residual add, optional PixelShuffle(2) scatter with crop.
// ------------------------------------------------------------------------------------------------
struct ConvF32 {
  const float* in;   // [Cin][H][W]
  const float* w;    // [Cout][Cin][ks][ks]
  const float* b;    // [Cout]
  float* out;        // [Cout][Ho][Wo]  (PixelShuffle: [Cout/4][outH][outW])
  const float* res;  // optional, same shape as out
  int Cin, Cout, H, W, Ho, Wo, ks, stride, act;
  float slope;
  int ps;            // 1: PixelShuffle(2) + crop to (outH,outW)
  int outH, outW;
  ActQuant q;        // fake-quantisation of the input (INT8 layouts)
};

// Register tile: PXT consecutive output pixels x COB output channels per thread.  Per (ci, ky) the thread loads the input
// row segment its PXT pixels need once (stride * (PXT - 1) + ks values) and per tap COB weights from shared memory as
// 128-bit broadcasts, so PXT * COB FMAs stand against COB / 4 shared loads: the loop is FMA-bound.  For the INT8 layouts the
// caller passes a pre-quantised tensor and p.q.mode = 0 whenever it can.  The accumulation order per output (ci, ky, kx)
// does not depend on the tile, and a tap outside the image adds an exact zero, so results are identical to a one-pixel,
// skip-the-tap kernel.
template <int COB, int PXT>
__global__ void __launch_bounds__(64) conv_f32_kernel(const ConvF32 p) {
  extern __shared__ __align__(16) float wsm[];  // [Cin*ks*ks][COB]
  const int co0 = blockIdx.z * COB;
  const int taps = p.ks * p.ks;
  const int kk = p.Cin * taps;
  for (int i = threadIdx.x; i < kk * COB; i += blockDim.x) {
    const int k = i / COB, c = i % COB;
    wsm[i] = (co0 + c < p.Cout) ? p.w[static_cast<long>(co0 + c) * kk + k] : 0.f;
  }
  __syncthreads();
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * PXT;
  const int oy = blockIdx.y;
  if (ox0 >= p.Wo) return;
  float acc[PXT][COB];
#pragma unroll
  for (int q = 0; q < PXT; ++q)
#pragma unroll
    for (int c = 0; c < COB; ++c) acc[q][c] = 0.f;
  const int pad = p.ks / 2;
  constexpr int SEG = 2 * (PXT - 1) + 3;          // widest segment: stride 2, 3 taps
  const int seg = p.stride * (PXT - 1) + p.ks;
  const int ix0 = ox0 * p.stride - pad;
  for (int ci = 0; ci < p.Cin; ++ci) {
    const float* ip = p.in + static_cast<long>(ci) * p.H * p.W;
    for (int ky = 0; ky < p.ks; ++ky) {
      const int iy = oy * p.stride + ky - pad;
      if (iy < 0 || iy >= p.H) continue;           // the whole row of taps is padding: adds nothing
      const float* row = ip + static_cast<long>(iy) * p.W;
      float v[SEG];
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        const int ix = ix0 + j;
        v[j] = (j < seg && ix >= 0 && ix < p.W) ? fake_quant(__ldg(row + ix), p.q) : 0.f;
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
              const float x = p.stride == 2 ? v[2 * q + kx] : v[q + kx];
              acc[q][4 * c + 0] = fmaf(x, w4.x, acc[q][4 * c + 0]);
              acc[q][4 * c + 1] = fmaf(x, w4.y, acc[q][4 * c + 1]);
              acc[q][4 * c + 2] = fmaf(x, w4.z, acc[q][4 * c + 2]);
              acc[q][4 * c + 3] = fmaf(x, w4.w, acc[q][4 * c + 3]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PXT; ++q) {
    const int ox = ox0 + q;
    if (ox >= p.Wo) break;
#pragma unroll
    for (int c = 0; c < COB; ++c) {
      const int co = co0 + c;
      if (co >= p.Cout) break;
      float v = acc[q][c] + __ldg(p.b + co);
      if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ACT_LRELU) v = v >= 0.f ? v : v * p.slope;
      long o;
      if (p.ps) {
        const int Y = 2 * oy + ((co & 3) >> 1), X = 2 * ox + (co & 1);
        if (Y >= p.outH || X >= p.outW) continue;
        o = (static_cast<long>(co >> 2) * p.outH + Y) * p.outW + X;
      } else {
        o = (static_cast<long>(co) * p.Ho + oy) * p.Wo + ox;
      }
      if (p.res) v += __ldg(p.res + o);
      p.out[o] = v;
    }
  }
}
// INT8 layouts on the FP32 path: a layer's input passes through its quantiser once, here, instead of once per tap and
// output-channel block inside the convolution
__global__ void fake_quant_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long n, ActQuant q) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fake_quant(x[i], q);
}

// y = x * (scale + 1) + shift   (arch_util.py:72)
__global__ void sft_mod_f32_kernel(const float* x, const float* scale, const float* shift, float* y, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * (scale[i] + 1.f) + shift[i];
}
__global__ void add_f32_kernel(const float* a, const float* b, float* y, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}
__global__ void half_to_f32_kernel(const __half* a, float* y, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __half2float(a[i]);
}
__global__ void f32_to_half_kernel(const float* a, __half* y, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(a[i]);
}

// ------------------------------------------------------------------------------------------------
// AGCM condition classifier (Color_Condition, Condition_arch.py:19-35).
// One kernel per block level: y = LeakyReLU_0.2( AvgPool3/2/1( conv1x1( IN_prev(x) ) ) ), and the per-channel
// sum / sum-of-squares of y are accumulated in FP64 for the InstanceNorm that the NEXT level folds into its
// input read (IN is affine per channel once mean/var are known).  count_include_pad=True => always /9, and the
// zero padding applies to the conv OUTPUT (bias included), so border windows simply have fewer terms.
// ------------------------------------------------------------------------------------------------
struct ClsLevel {
  const void* in;          // level 0: planar cond [3][H][W] (fp16 or fp32); deeper levels: pixel-major fp32 [H][W][Cin]
  int in_planar;
  int in_is_half;
  const double* in_stats;  // [Cin][2] sum, sumsq of the previous level's output (nullptr: no IN before this level)
  const float* gamma;      // IN affine of the previous level
  const float* beta;
  const float* w;          // [Cin][Cout]  (transposed copy made at weight-load time)
  const float* b;          // [Cout]
  float* out;              // pixel-major [Ho][Wo][Cout]
  double* out_stats;       // [Cout][2]: written (not accumulated) by the last block of the level
  double* partials;        // [blocks][2 * Cout] per-block sums, then [groups][2 * Cout] per-group sums
  unsigned int* counter;   // [groups] blocks of a group that have published, then [1] groups that have published; self-resetting
  int Cin, Cout, H, W, Ho, Wo;
  int pix;                 // pooled pixels per block
  ActQuant q;              // fake-quantisation of this level's conv input (after the IN affine)
  ActQuant stat_q;         // applied to y before it enters out_stats' SUM (the input quantiser of the conv that follows
                           // the last level, whose global mean is all that is used); sum^2 then is unused
};

// One level of the AGCM condition classifier (Condition_arch.py:8-35): IN-affine of the previous level, AvgPool(3,2,1)
// and the 1x1 conv commute (all linear), so a block of 256 threads computes for `pix` pooled pixels
//   1. S[px][ci] = sum over the valid 3x3 window of IN(x)[ci]           (coalesced: channels are the fast index)
//   2. y[px][co] = LeakyReLU_0.2((W[co,:] . S[px,:] + nwin * b[co]) / 9) with the level's weights staged in shared memory
//   3. per-channel sum / sum^2 in FP64: registers -> fixed-order shared-memory tree per block.
// The work is tiny (<= 17 MMAC per level); the code is built for latency: ~256 blocks per level, nine independent loads
// per window, per-block partial sums added to the level's global FP64 totals (one atomic pair per channel per block).
// Measured and dropped: all five levels + head in ONE launch on a 16-CTA cluster (hardware cluster barriers, partial sums
// exchanged through distributed shared memory): 172 us at 1080p against 92 us for the six launches - sixteen SMs cannot
// hide the load latency of the two large levels that 148 SMs hide.
struct ClsSmem {
  float* Wsm;    // [Cin][Cout]
  float* S;      // [PIX][Cin]
  float* na;     // [Cin] IN scale
  float* nb;     // [Cin] IN shift
  int* nwin;     // [PIX]
};
__device__ __forceinline__ ClsSmem cls_smem(const ClsLevel& p, float* sm) {
  ClsSmem m;
  m.Wsm = sm;
  m.S = m.Wsm + p.Cin * p.Cout;
  m.na = m.S + p.pix * p.Cin;
  m.nb = m.na + p.Cin;
  m.nwin = reinterpret_cast<int*>(m.nb + p.Cin);
  return m;
}
template <int NT>
__device__ __forceinline__ void cls_stage_weights(const ClsLevel& p, const ClsSmem& m) {
  for (int i = threadIdx.x; i < p.Cin * p.Cout; i += NT) m.Wsm[i] = __ldg(p.w + i);
}
// IN affine of the previous level from its per-channel totals [Cin][2] (sum, sum of squares); stats == nullptr: identity
template <int NT>
__device__ __forceinline__ void cls_norm_coeffs(const ClsLevel& p, const ClsSmem& m, const double* stats) {
  const double cnt = static_cast<double>(p.H) * p.W;
  for (int ci = threadIdx.x; ci < p.Cin; ci += NT) {
    if (stats) {
      const double mean = stats[2 * ci] / cnt;
      double var = stats[2 * ci + 1] / cnt - mean * mean;
      var = var < 0 ? 0 : var;
      const double rstd = 1.0 / sqrt(var + 1e-5);
      m.na[ci] = static_cast<float>(rstd * p.gamma[ci]);
      m.nb[ci] = static_cast<float>(p.beta[ci] - mean * rstd * p.gamma[ci]);
    } else {
      m.na[ci] = 1.f;
      m.nb[ci] = 0.f;
    }
  }
}
// one tile of p.pix pooled pixels starting at pix0; adds this thread's share of the tile to (s1, s2) of channel tid % Cout.
// Ends with the tile's shared-memory buffers free for reuse.  The nine taps of a window are loaded unconditionally from
// clamped coordinates (nine independent loads in flight; a tap outside the map contributes an exact 0), in the same
// summation order as a conditional loop.
template <int NT>
__device__ __forceinline__ void cls_tile(const ClsLevel& p, const ClsSmem& m, int pix0, double& s1, double& s2) {
  const int Cin = p.Cin, Cout = p.Cout, PIX = p.pix, tid = threadIdx.x;
  const int npix = p.Ho * p.Wo;
  for (int item = tid; item < PIX * Cin; item += NT) {
    int px, ci;
    if (p.in_planar) { px = item % PIX; ci = item / PIX; }      // planar input: pixels are the fast index
    else { ci = item % Cin; px = item / Cin; }
    const int idx = pix0 + px;
    float acc = 0.f;
    int n = 0;
    if (idx < npix) {
      const int oy = idx / p.Wo, ox = idx % p.Wo;
      const float a = m.na[ci], b = m.nb[ci];
      float v[9];
      bool ok[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky - 1;
        const int iyc = min(max(iy, 0), p.H - 1);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = 2 * ox + kx - 1;
          const int ixc = min(max(ix, 0), p.W - 1);
          ok[ky * 3 + kx] = iy == iyc && ix == ixc;
          if (p.in_planar) {
            const long o = (static_cast<long>(ci) * p.H + iyc) * p.W + ixc;
            v[ky * 3 + kx] = p.in_is_half ? __half2float(reinterpret_cast<const __half*>(p.in)[o]) : reinterpret_cast<const float*>(p.in)[o];
          } else {
            v[ky * 3 + kx] = reinterpret_cast<const float*>(p.in)[(static_cast<long>(iyc) * p.W + ixc) * Cin + ci];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (ok[k]) {
          acc += fake_quant(fmaf(v[k], a, b), p.q);
          ++n;
        }
      }
    }
    m.S[px * Cin + ci] = acc;
    if (ci == 0) m.nwin[px] = n;
  }
  __syncthreads();
  const int co = tid % Cout;                        // fixed per thread: Cout divides NT
  const float bias = __ldg(p.b + co);
  for (int item = tid; item < PIX * Cout; item += NT) {
    const int px = item / Cout;
    const int idx = pix0 + px;
    if (idx >= npix) break;
    const float* sp = m.S + px * Cin;
    float acc0 = 0.f, acc1 = 0.f;
    int ci = 0;
    for (; ci + 1 < Cin; ci += 2) {
      acc0 = fmaf(sp[ci], m.Wsm[ci * Cout + co], acc0);
      acc1 = fmaf(sp[ci + 1], m.Wsm[(ci + 1) * Cout + co], acc1);
    }
    if (ci < Cin) acc0 = fmaf(sp[ci], m.Wsm[ci * Cout + co], acc0);
    float y = (acc0 + acc1 + m.nwin[px] * bias) / 9.0f;
    y = y >= 0.f ? y : 0.2f * y;
    p.out[static_cast<long>(idx) * Cout + co] = y;
    s1 += fake_quant(y, p.stat_q);
    s2 += static_cast<double>(y) * y;
  }
  __syncthreads();
}
// per-channel totals of the block: thread t < 2 * Cout returns, for channel t >> 1 and statistic t & 1, the sum over the
// NT / Cout threads that own the channel, in fixed order (shared memory, no atomics); other threads return 0
template <int NT>
__device__ __forceinline__ double cls_block_reduce(int Cout, double s1, double s2, double* red /* [2 * NT] */) {
  const int tid = threadIdx.x;
  red[tid] = s1;
  red[NT + tid] = s2;
  __syncthreads();
  double tot = 0.0;
  if (tid < 2 * Cout) {
    const int ch = tid >> 1;
    const double* r = red + (tid & 1) * NT;
    for (int k = ch; k < NT; k += Cout) tot += r[k];
  }
  __syncthreads();
  return tot;
}

constexpr unsigned kClsGroup = 32;     // blocks per first-stage group of the level-total reduction
__global__ void __launch_bounds__(256) cls_level_kernel(const ClsLevel p) {
  extern __shared__ float sm[];
  __shared__ double red[512];
  __shared__ bool last_block;
  const ClsSmem m = cls_smem(p, sm);
  // Programmatic dependent launch: the next level may start its prologue (static weights -> shared memory) while this
  // level drains; everything below the wait reads what the previous launch wrote (input map, its statistics).
  grid_dep_launch();
  cls_stage_weights<256>(p, m);
  grid_dep_wait();
  cls_norm_coeffs<256>(p, m, p.in_stats);
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  cls_tile<256>(p, m, blockIdx.x * p.pix, s1, s2);
  const double tot = cls_block_reduce<256>(p.Cout, s1, s2, red);
  // Level totals without atomics on the data (threadfence reduction, two stages): every block publishes its partial sums;
  // the block that arrives last in its group of kClsGroup blocks adds the group's partials in block order and publishes the
  // group sum; the group that arrives last adds the group sums in group order.  FP64 atomicAdd in arrival order made the
  // totals - and through the InstanceNorm's E[x^2] - mean^2 the whole condition vector - differ in the last bit from run to
  // run on some frames (one fp32 ulp of `fea`: invisible in the FP16 output, amplified by bucket flips in the INT8 layouts).
  const int tid = threadIdx.x, nq = 2 * p.Cout;                // nq <= 256
  const unsigned nblk = gridDim.x, ngrp = (nblk + kClsGroup - 1) / kClsGroup;
  const unsigned grp = blockIdx.x / kClsGroup, b0 = grp * kClsGroup, gsize = min(kClsGroup, nblk - b0);
  double* gpart = p.partials + static_cast<long>(nblk) * nq;
  if (tid < nq) p.partials[static_cast<long>(blockIdx.x) * nq + tid] = tot;
  __threadfence();
  __syncthreads();
  if (tid == 0) last_block = atomicAdd(p.counter + grp, 1u) == gsize - 1;
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  if (tid < nq) {
    double a = 0.0;
#pragma unroll 8
    for (unsigned b = 0; b < gsize; ++b) a += __ldcg(p.partials + static_cast<long>(b0 + b) * nq + tid);
    gpart[static_cast<long>(grp) * nq + tid] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    p.counter[grp] = 0u;                                        // ready for the next frame
    last_block = atomicAdd(p.counter + ngrp, 1u) == ngrp - 1;
  }
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  if (tid < nq) {
    double a = 0.0;
#pragma unroll 8
    for (unsigned g = 0; g < ngrp; ++g) a += __ldcg(gpart + static_cast<long>(g) * nq + tid);
    p.out_stats[tid] = a;                                       // [Cout][2]: index = 2 * channel + statistic
  }
  if (tid == 0) p.counter[ngrp] = 0u;
}

}  // namespace hdrtv
