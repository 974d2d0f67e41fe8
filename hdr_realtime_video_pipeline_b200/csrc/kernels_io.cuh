// Bandwidth-bound ends of the path: uint8 BGR -> normalised planar RGB (+ antialiased 1/4 condition image),
// and the clamp / quantise / interleave packs (RGB48 for the mpv/ffmpeg feeders, BGR24 for postprocess()).
// Reference: hdrtvnet_torch.py:2239-2296 (preprocess), :2352-2368 (postprocess),
//            gui_pipeline_worker_feeders.py:193-249 (_tensor_to_rgb48_bytes).
#pragma once
#include "common.cuh"

namespace hdrtv {

__device__ __forceinline__ float norm_u8(uint32_t b, bool half_round) {
  // float(u8) * fp32(1/255): a multiply, not a divide (hdrtvnet_torch.py:2259); fp16 rounds the product once.
  float f = __fmul_rn(static_cast<float>(b), 0.003921568859368563f);
  return half_round ? __half2float(__float2half_rn(f)) : f;
}

// ---- normalise: 16 pixels per thread, 3 x 128-bit loads, planar 128/256-bit stores (needs W % 16 == 0) ----
template <typename T>
__global__ void __launch_bounds__(256) normalize_vec16_kernel(const uint8_t* __restrict__ bgr, T* __restrict__ x, int H,
                                                              int W) {
  const long g = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;  // group of 16 pixels
  const long ngroups = static_cast<long>(H) * W / 16;
  if (g >= ngroups) return;
  const uint4* src = reinterpret_cast<const uint4*>(bgr) + g * 3;
  alignas(16) uint4 q[3];
  q[0] = __ldg(src);
  q[1] = __ldg(src + 1);
  q[2] = __ldg(src + 2);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(q);
  const long plane = static_cast<long>(H) * W;
  constexpr bool kHalf = sizeof(T) == 2;
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // output channel c = R,G,B  <- input byte 2-c of each BGR triple
    alignas(16) T v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float f = __fmul_rn(static_cast<float>(bytes[3 * i + (2 - c)]), 0.003921568859368563f);
      if constexpr (kHalf) v[i] = __float2half_rn(f);
      else v[i] = f;
    }
    uint4* dst = reinterpret_cast<uint4*>(x + c * plane + g * 16);
    const uint4* vs = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T)) * 16 / 16; ++i) dst[i] = vs[i];
  }
}
template <typename T>
__global__ void normalize_scalar_kernel(const uint8_t* __restrict__ bgr, T* __restrict__ x, int H, int W) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long plane = static_cast<long>(H) * W;
  if (i >= plane) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float f = __fmul_rn(static_cast<float>(bgr[3 * i + (2 - c)]), 0.003921568859368563f);
    if constexpr (sizeof(T) == 2) x[c * plane + i] = __float2half_rn(f);
    else x[c * plane + i] = f;
  }
}

// ---- normalise + tensor-core staging in one pass (FP16 path): besides the planar (1,3,H,W) tensor the API returns, the
// same threads write the image as the single-chunk P8 tensor [R,G,B,0,0,0,0,0] the AGCM chain reads (one 16-byte entry
// per pixel).  The entries go through shared memory so that consecutive lanes store consecutive entries (128-bit,
// coalesced); a block covers 4096 consecutive pixels of the frame (needs W % 16 == 0).
__global__ void __launch_bounds__(256) normalize_p8_kernel(const uint8_t* __restrict__ bgr, __half* __restrict__ x, P8 xp8,
                                                           int H, int W) {
  __shared__ uint2 px[4096];                                   // (R,G | B,0) halves of one pixel
  const long g = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;  // group of 16 pixels
  const long plane = static_cast<long>(H) * W;
  const long ngroups = plane / 16;
  if (g < ngroups) {
    const uint4* src = reinterpret_cast<const uint4*>(bgr) + g * 3;
    alignas(16) uint4 q[3];
    q[0] = __ldg(src);
    q[1] = __ldg(src + 1);
    q[2] = __ldg(src + 2);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(q);
    alignas(16) __half v[3][16];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 16; ++i) v[c][i] = __float2half_rn(__fmul_rn(static_cast<float>(bytes[3 * i + (2 - c)]), 0.003921568859368563f));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint4* dst = reinterpret_cast<uint4*>(x + c * plane + g * 16);
      dst[0] = reinterpret_cast<const uint4*>(v[c])[0];
      dst[1] = reinterpret_cast<const uint4*>(v[c])[1];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint2 e;
      e.x = static_cast<uint32_t>(__half_as_ushort(v[0][i])) | (static_cast<uint32_t>(__half_as_ushort(v[1][i])) << 16);
      e.y = static_cast<uint32_t>(__half_as_ushort(v[2][i]));
      px[threadIdx.x * 16 + i] = e;
    }
  }
  __syncthreads();
  const long p0 = static_cast<long>(blockIdx.x) * 4096;
  uint4* base = reinterpret_cast<uint4*>(xp8.base);
#pragma unroll 4
  for (int k = threadIdx.x; k < 4096; k += 256) {
    const long pidx = p0 + k;
    if (pidx >= plane) break;
    const int y = static_cast<int>(pidx / W), xx = static_cast<int>(pidx - static_cast<long>(y) * W);
    const uint2 e = px[k];
    base[xp8.entry(y, 0, xx)] = make_uint4(e.x, e.y, 0u, 0u);
  }
}

// ---- condition image: antialiased bicubic x0.25 (separable 16-tap Keys cubic, a = -0.5, normalised taps) ----
// Tap tables are built on the host per resolution: start index + 16 fp32 weights (zero beyond `count`).
// One block = 8x32 output pixels; the block's u8 window is normalised once into shared memory, a horizontal
// pass writes row sums, a vertical pass finishes.  fp32 accumulation, one rounding to the output dtype.
struct CondTaps {
  const int* xstart;
  const float* xw;  // [Wc][16]
  const int* ystart;
  const float* yw;  // [Hc][16]
};
constexpr int kCondTW = 32, kCondTH = 8;
template <typename T>
__global__ void __launch_bounds__(256) cond_aa_kernel(const uint8_t* __restrict__ bgr, T* __restrict__ cond, int H, int W,
                                                      int Hc, int Wc, CondTaps tp, int mode /*0 aa, 1 zero, 2 bilinear*/) {
  // window: rows [ys0, ys0+RH), cols [xs0, xs0+RW)
  constexpr int RW = kCondTW * 4 + 16, RH = kCondTH * 4 + 16;
  __shared__ float hs[3][RH][kCondTW + 1];
  const int ox0 = blockIdx.x * kCondTW, oy0 = blockIdx.y * kCondTH;
  const int tx = threadIdx.x % kCondTW, ty = threadIdx.x / kCondTW;
  const int ox = ox0 + tx, oy = oy0 + ty;
  const long cplane = static_cast<long>(Hc) * Wc;
  if (mode == 1) {
    if (ox < Wc && oy < Hc)
      for (int c = 0; c < 3; ++c) {
        if constexpr (sizeof(T) == 2) cond[c * cplane + static_cast<long>(oy) * Wc + ox] = __float2half_rn(0.f);
        else cond[c * cplane + static_cast<long>(oy) * Wc + ox] = 0.f;
      }
    return;
  }
  if (mode == 2) {
    // fast_condition_resize (hdrtvnet_torch.py:2268-2275): bilinear x0.25, align_corners=False -> source coordinate
    // 4i + 1.5, i.e. the mean of pixels 4i+1 and 4i+2 in both directions (fp32 accumulate, one rounding)
    if (ox < Wc && oy < Hc) {
      for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int dy = 1; dy <= 2; ++dy) {
          const int iy = min(4 * oy + dy, H - 1);
          float rowv = 0.f;
#pragma unroll
          for (int dx = 1; dx <= 2; ++dx) {
            const int ix = min(4 * ox + dx, W - 1);
            rowv += 0.5f * norm_u8(bgr[(static_cast<long>(iy) * W + ix) * 3 + (2 - c)], sizeof(T) == 2);
          }
          acc += 0.5f * rowv;
        }
        if constexpr (sizeof(T) == 2) cond[c * cplane + static_cast<long>(oy) * Wc + ox] = __float2half_rn(acc);
        else cond[c * cplane + static_cast<long>(oy) * Wc + ox] = acc;
      }
    }
    return;
  }
  const int ys0 = tp.ystart[min(oy0, Hc - 1)];
  // Stage the block's u8 window (RH rows x RW pixels, BGR interleaved) in shared memory with coalesced 32-bit loads:
  // the horizontal pass below reads every byte 4 times (16 taps / stride 4), and byte loads straight from global
  // memory ran this kernel at 229 GB/s (profiles/r1_launches_1080p.md).
  __shared__ uint32_t win[RH][(RW * 3 + 3) / 4 + 2];
  const int xs0 = tp.xstart[min(ox0, Wc - 1)];
  const long row_bytes = static_cast<long>(W) * 3;
  const int b0 = xs0 * 3;
  const int b0a = b0 & ~3;                                    // word-aligned start inside a row (rows are W*3 bytes, W % 4 == 0
  const int skew = b0 - b0a;                                  //  is not required: alignment is per absolute byte address below)
  constexpr int WORDS = (RW * 3 + 3) / 4 + 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(bgr) & 3) == 0) && ((row_bytes & 3) == 0);
  for (int r = ty; r < RH; r += kCondTH) {
    const int iy = ys0 + r;
    if (iy >= H) continue;
    const uint8_t* row = bgr + static_cast<long>(iy) * row_bytes;
    for (int wd = tx; wd < WORDS; wd += kCondTW) {
      const long off = static_cast<long>(b0a) + 4 * wd;
      uint32_t v = 0;
      if (aligned && off + 4 <= row_bytes) v = __ldg(reinterpret_cast<const uint32_t*>(row + off));
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (off + k < row_bytes) v |= static_cast<uint32_t>(row[off + k]) << (8 * k);
      }
      win[r][wd] = v;
    }
  }
  __syncthreads();
  // horizontal pass: thread (row r, out col tx) for r = ty, ty+8, ...
  const int oxc = min(ox, Wc - 1);
  const int xs = tp.xstart[oxc];
  float wx[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) wx[k] = tp.xw[oxc * 16 + k];
  for (int r = ty; r < RH; r += kCondTH) {
    const int iy = ys0 + r;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (iy < H) {
      const uint8_t* wrow = reinterpret_cast<const uint8_t*>(win[r]) + skew;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int ix = min(xs + k, W - 1) - xs0;  // weights beyond `count` are zero
        const uint8_t* px = wrow + 3 * ix;
        a0 = fmaf(wx[k], norm_u8(px[2], sizeof(T) == 2), a0);
        a1 = fmaf(wx[k], norm_u8(px[1], sizeof(T) == 2), a1);
        a2 = fmaf(wx[k], norm_u8(px[0], sizeof(T) == 2), a2);
      }
    }
    hs[0][r][tx] = a0;
    hs[1][r][tx] = a1;
    hs[2][r][tx] = a2;
  }
  __syncthreads();
  if (ox >= Wc || oy >= Hc) return;
  const int r0 = tp.ystart[oy] - ys0;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float wy = tp.yw[oy * 16 + k];
    const int r = min(r0 + k, RH - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = fmaf(wy, hs[c][r][tx], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if constexpr (sizeof(T) == 2) cond[c * cplane + static_cast<long>(oy) * Wc + ox] = __float2half_rn(acc[c]);
    else cond[c * cplane + static_cast<long>(oy) * Wc + ox] = acc[c];
  }
  (void)RW;
}

// ---- RGB48 pack: FP32 math, clamp(0,1) -> *65535 (rounded) -> +0.5 (rounded) -> truncate; RGB order, HWC ----
__device__ __forceinline__ uint32_t q16(float f) {
  f = fminf(fmaxf(f, 0.f), 1.f);
  f = __fmul_rn(f, 65535.0f);
  f = __fadd_rn(f, 0.5f);
  return static_cast<uint32_t>(f);  // truncating convert (value in [0.5, 65535.5])
}
template <typename T>
__device__ __forceinline__ float ldf(const T* p) {
  if constexpr (sizeof(T) == 2) return __half2float(*p);
  else return *p;
}
// 8 pixels per thread: planar loads, 3 x 128-bit interleaved stores (needs W*H % 8 == 0).  `lut` (optional,
// fp16 inputs only) maps the half bit pattern of the clamped value to a code (PQ transfer option).
// `checksum` (optional): the frame descriptor's order-sensitive 64-bit checksum sum_i code[i] * ((i mod 65521) + 1) over the
// flattened (H,W,3) array (sharding.frame_checksum), accumulated with integer atomics (exact, order-independent).
__device__ __forceinline__ unsigned long long cks_term(uint32_t code, long elem) {
  return static_cast<unsigned long long>(code) * static_cast<unsigned long long>(static_cast<uint32_t>(elem % 65521) + 1u);
}
__device__ __forceinline__ void cks_block_add(unsigned long long v, unsigned long long* checksum) {
  __shared__ unsigned long long part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += part[i];
    if (t) atomicAdd(checksum, t);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) pack_rgb48_kernel(const T* __restrict__ src, uint16_t* __restrict__ dst, long npix,
                                                         const uint16_t* __restrict__ lut, int vec_ok,
                                                         unsigned long long* __restrict__ checksum) {
  const long g = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i0 = g * 8;
  unsigned long long cks = 0;
  if (i0 >= npix) {
    if (checksum) cks_block_add(0, checksum);
    return;
  }
  if (vec_ok && i0 + 8 <= npix) {
    alignas(16) uint16_t o[24];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      alignas(16) T v[8];
      if constexpr (sizeof(T) == 2) {
        *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(src + c * npix + i0));
      } else {
        reinterpret_cast<uint4*>(v)[0] = __ldg(reinterpret_cast<const uint4*>(src + c * npix + i0));
        reinterpret_cast<uint4*>(v)[1] = __ldg(reinterpret_cast<const uint4*>(src + c * npix + i0) + 1);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (lut != nullptr && sizeof(T) == 2) {
          __half h = __float2half_rn(fminf(fmaxf(ldf(&v[k]), 0.f), 1.f));
          o[3 * k + c] = lut[__half_as_ushort(h)];
        } else {
          o[3 * k + c] = static_cast<uint16_t>(q16(ldf(&v[k])));
        }
      }
    }
    uint4* d = reinterpret_cast<uint4*>(dst + i0 * 3);
    const uint4* os = reinterpret_cast<const uint4*>(o);
    d[0] = os[0];
    d[1] = os[1];
    d[2] = os[2];
    if (checksum) {
#pragma unroll
      for (int k = 0; k < 24; ++k) cks += cks_term(o[k], i0 * 3 + k);
    }
  } else {
    const long i1 = i0 + 8 < npix ? i0 + 8 : npix;
    for (long i = i0; i < i1; ++i)
      for (int c = 0; c < 3; ++c) {
        uint16_t code;
        if (lut != nullptr && sizeof(T) == 2) {
          __half h = __float2half_rn(fminf(fmaxf(ldf(src + c * npix + i), 0.f), 1.f));
          code = lut[__half_as_ushort(h)];
        } else {
          code = static_cast<uint16_t>(q16(ldf(src + c * npix + i)));
        }
        dst[i * 3 + c] = code;
        cks += cks_term(code, i * 3 + c);
      }
  }
  if (checksum) cks_block_add(cks, checksum);
}

// ---- BGR24 pack: arithmetic IN THE TENSOR'S DTYPE (half: each of mul/add rounds to half), truncate, RGB->BGR ----
template <typename T>
__device__ __forceinline__ uint32_t q8(T v) {
  if constexpr (sizeof(T) == 2) {
    __half h = __hmin(__hmax(v, __float2half_rn(0.f)), __float2half_rn(1.f));
    h = __hmul(h, __float2half_rn(255.f));
    h = __hadd(h, __float2half_rn(0.5f));
    return static_cast<uint32_t>(__half2float(h));
  } else {
    float f = fminf(fmaxf(v, 0.f), 1.f);
    f = __fmul_rn(f, 255.f);
    f = __fadd_rn(f, 0.5f);
    return static_cast<uint32_t>(f);
  }
}
// 4 pixels per thread -> 12 bytes = 3 x 32-bit stores (needs npix % 4 == 0 for the vector body).
template <typename T>
__global__ void __launch_bounds__(256) pack_bgr24_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, long npix,
                                                         int vec_ok) {
  const long g = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i0 = g * 4;
  if (i0 >= npix) return;
  if (vec_ok && i0 + 4 <= npix) {
    alignas(4) uint8_t o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[3 * k + (2 - c)] = static_cast<uint8_t>(q8<T>(src[c * npix + i0 + k]));
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + i0 * 3);
    const uint32_t* os = reinterpret_cast<const uint32_t*>(o);
    d[0] = os[0];
    d[1] = os[1];
    d[2] = os[2];
  } else {
    const long i1 = i0 + 4 < npix ? i0 + 4 : npix;
    for (long i = i0; i < i1; ++i)
      for (int c = 0; c < 3; ++c) dst[i * 3 + (2 - c)] = static_cast<uint8_t>(q8<T>(src[c * npix + i]));
  }
}

// ---- planar fp16 (3,H,W) -> P8 single chunk [R,G,B,0,0,0,0,0] ----
__global__ void planar_to_p8_kernel(const __half* __restrict__ src, P8 dst, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const long plane = static_cast<long>(H) * W, o = static_cast<long>(y) * W + x;
  alignas(16) __half v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __float2half_rn(0.f);
  v[0] = src[o];
  v[1] = src[plane + o];
  v[2] = src[2 * plane + o];
  reinterpret_cast<uint4*>(dst.base)[dst.entry(y, 0, x)] = *reinterpret_cast<const uint4*>(v);
}

// ---------------------------------------------------------------------------------------------------------------------
// Letterbox (gui_scaling.py:228-244 `_letterbox_bgr`): aspect-preserving cv2.resize (INTER_AREA when shrinking,
// INTER_CUBIC when enlarging) centred on a black canvas, uint8 HxWx3 in and out, one thread per canvas pixel.
// The arithmetic is OpenCV's own (modules/imgproc/src/resize.cpp, restated in oracle/hdrtvnet_oracle.py and pinned
// against cv2 there), operation by operation so that the bytes are identical:
//   LB_AREA_INT  integer factors: block sum; 2x2 -> (s + 2) >> 2, else rint(float(s) * float(1 / area))
//   LB_AREA      general: fp32 table weights, horizontal then vertical accumulation in table order, every product
//                rounded before it is added (__fmul_rn / __fadd_rn: no FMA contraction)
//   LB_CUBIC     a = -0.75 weights as shorts (x 2048); horizontal pass exact in int32 over replicate-clamped taps;
//                vertical pass in fp32 (S3*b3, S2*b2 + ., S1*b1 + ., S0*b0 + .) with b = short * 2^-22, rint; the last
//                (new_w * 3) mod 8 bytes of a row use the fixed-point form (the scalar tail of OpenCV's vector loop)
// ---------------------------------------------------------------------------------------------------------------------
enum { LB_COPY = 0, LB_AREA_INT = 1, LB_AREA = 2, LB_CUBIC = 3 };
struct Letterbox {
  const uint8_t* src;
  uint8_t* dst;
  int H, W, out_H, out_W, new_H, new_W, y0, x0, mode;
  int fy, fx;                       // LB_AREA_INT: integer factors
  const int* xofs; const int* yofs; // LB_AREA: first source index per destination index; LB_CUBIC: 4 clamped tap indices
  const int* xcnt; const int* ycnt; // LB_AREA: taps per destination index
  const float* xw; const float* yw; // LB_AREA: [dst][K] weights; LB_CUBIC: yw = [dst][4] short weights * 2^-22
  const int* xi; const int* yi;     // LB_CUBIC: [dst][4] short weights (as int)
  int KX, KY;
};
__device__ __forceinline__ uint8_t lb_sat(float v) { return static_cast<uint8_t>(min(max(__float2int_rn(v), 0), 255)); }
__global__ void __launch_bounds__(256) letterbox_kernel(const Letterbox p) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
  if (X >= p.out_W) return;
  uint8_t* o = p.dst + (static_cast<long>(Y) * p.out_W + X) * 3;
  const int dx = X - p.x0, dy = Y - p.y0;
  if (dx < 0 || dx >= p.new_W || dy < 0 || dy >= p.new_H) { o[0] = 0; o[1] = 0; o[2] = 0; return; }
  const long rs = static_cast<long>(p.W) * 3;
  if (p.mode == LB_COPY) {
    const uint8_t* s = p.src + dy * rs + dx * 3;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
  } else if (p.mode == LB_AREA_INT) {
    int s0 = 0, s1 = 0, s2 = 0;
    for (int ky = 0; ky < p.fy; ++ky) {
      const uint8_t* s = p.src + (static_cast<long>(dy) * p.fy + ky) * rs + static_cast<long>(dx) * p.fx * 3;
      for (int kx = 0; kx < p.fx; ++kx) { s0 += s[3 * kx]; s1 += s[3 * kx + 1]; s2 += s[3 * kx + 2]; }
    }
    if (p.fx == 2 && p.fy == 2) { o[0] = (s0 + 2) >> 2; o[1] = (s1 + 2) >> 2; o[2] = (s2 + 2) >> 2; }
    else {
      const float sc = 1.f / static_cast<float>(p.fx * p.fy);
      o[0] = lb_sat(__fmul_rn(static_cast<float>(s0), sc));
      o[1] = lb_sat(__fmul_rn(static_cast<float>(s1), sc));
      o[2] = lb_sat(__fmul_rn(static_cast<float>(s2), sc));
    }
  } else if (p.mode == LB_AREA) {
    const int sx0 = p.xofs[dx], nx = p.xcnt[dx], sy0 = p.yofs[dy], ny = p.ycnt[dy];
    const float* xw = p.xw + static_cast<long>(dx) * p.KX;
    const float* yw = p.yw + static_cast<long>(dy) * p.KY;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int ky = 0; ky < ny; ++ky) {
      const uint8_t* s = p.src + (sy0 + ky) * rs + static_cast<long>(sx0) * 3;
      float b0 = 0.f, b1 = 0.f, b2 = 0.f;
      for (int kx = 0; kx < nx; ++kx) {
        const float w = xw[kx];
        b0 = __fadd_rn(b0, __fmul_rn(static_cast<float>(s[3 * kx]), w));
        b1 = __fadd_rn(b1, __fmul_rn(static_cast<float>(s[3 * kx + 1]), w));
        b2 = __fadd_rn(b2, __fmul_rn(static_cast<float>(s[3 * kx + 2]), w));
      }
      const float b = yw[ky];
      a0 = ky ? __fadd_rn(a0, __fmul_rn(b0, b)) : __fmul_rn(b0, b);
      a1 = ky ? __fadd_rn(a1, __fmul_rn(b1, b)) : __fmul_rn(b1, b);
      a2 = ky ? __fadd_rn(a2, __fmul_rn(b2, b)) : __fmul_rn(b2, b);
    }
    o[0] = lb_sat(a0); o[1] = lb_sat(a1); o[2] = lb_sat(a2);
  } else {
    int hr[4][3];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const uint8_t* s = p.src + p.yofs[dy * 4 + ky] * rs;
      int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const uint8_t* q = s + static_cast<long>(p.xofs[dx * 4 + kx]) * 3;
        const int a = p.xi[dx * 4 + kx];
        h0 += q[0] * a; h1 += q[1] * a; h2 += q[2] * a;
      }
      hr[ky][0] = h0; hr[ky][1] = h1; hr[ky][2] = h2;
    }
    const int flat = p.new_W * 3, tail0 = flat - (flat & 7);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (dx * 3 + c >= tail0) {        // scalar tail of the vector loop: fixed point
        long f = 0;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) f += static_cast<long>(hr[ky][c]) * p.yi[dy * 4 + ky];
        o[c] = static_cast<uint8_t>(min(max(static_cast<int>((f + (1 << 21)) >> 22), 0), 255));
      } else {
        float acc = __fmul_rn(static_cast<float>(hr[3][c]), p.yw[dy * 4 + 3]);
        acc = __fadd_rn(__fmul_rn(static_cast<float>(hr[2][c]), p.yw[dy * 4 + 2]), acc);
        acc = __fadd_rn(__fmul_rn(static_cast<float>(hr[1][c]), p.yw[dy * 4 + 1]), acc);
        acc = __fadd_rn(__fmul_rn(static_cast<float>(hr[0][c]), p.yw[dy * 4 + 0]), acc);
        o[c] = lb_sat(acc);
      }
    }
  }
}

// debug: P8 -> planar fp32 (C,H,W)
__global__ void p8_to_planar_f32_kernel(P8 src, int j0, int C, float* dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= src.W) return;
  for (int c = 0; c < C; ++c) {
    const __half* e = src.base + src.entry(y, j0 + c / 8, x) * 8;
    dst[(static_cast<long>(c) * src.H + y) * src.W + x] = __half2float(e[c % 8]);
  }
}

}  // namespace hdrtv
