// Shared types for the HDRTVNet++ B200 engine.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace hdrtv {

// ---------------------------------------------------------------------------------------------
// "P8" activation layout (fp16): rows of channel-chunk planes.
//   entry = 8 consecutive channels of one pixel = 16 bytes (one UMMA core-matrix row)
//   element (y, chunk j, x):
//     natural : entry index ((y+1)*chunks + j)*Wp + (x+1)
//     parity  : entry index ((y+1)*chunks + j)*Wp + (x&1)*(Wp/2) + ((x>>1)+1)      (x = -1 -> odd plane, entry 0)
//   One zero row above and below, zero entries left/right of every plane (written once at allocation, never
//   touched by any epilogue) provide the convolution zero padding, so the 1-D bulk TMA row loads need no
//   predication.  Parity-split planes are what a stride-2 consumer needs: each of its three horizontal taps is
//   then a contiguous run of 16-byte rows, i.e. a legal K-major UMMA operand addressed by a byte shift.
// ---------------------------------------------------------------------------------------------
struct P8 {
  __half* base = nullptr;
  int chunks = 0;  // channel-chunk planes per row (C/8)
  int Wp = 0;      // entries per plane
  int parity = 0;
  int H = 0, W = 0;
  __host__ __device__ long row_entries() const { return static_cast<long>(chunks) * Wp; }
  __host__ __device__ long entries() const { return row_entries() * (H + 2); }
  __host__ __device__ long entry(int y, int j, int x) const {
    long e = (static_cast<long>(y + 1) * chunks + j) * Wp;
    return parity ? e + (x & 1) * (Wp >> 1) + ((x >> 1) + 1) : e + (x + 1);
  }
};

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

// Static activation fake-quantisation of a layer input (reference W8A8Conv2d / W8A8Linear.forward,
// hdrtvnet_torch.py:350-364): mode 2 (asymmetric) q = clamp(round((x - zero) / scale), 0, 255), x^ = q*scale + zero;
// mode 1 (symmetric) q = clamp(round(x / scale), -128, 127), x^ = q*scale; torch.round = round-half-even = rintf.
// A division, not a multiplication by the reciprocal: the bucket boundaries must be the reference's.
struct ActQuant {
  float scale = 1.f, zero = 0.f;
  int mode = 0;
  float inv = 1.f;          // 1 / scale (host): fast path of quant_code
};
// mode 2: the uint8 code as a float in [0, 255] = clamp(rint((x - zero) / scale), 0, 255) bit for bit, without an IEEE division
// and without the quarter-rate FRND / F2I conversions per element:
//   * the quotient is formed with the host-computed reciprocal and clamped first (the clamp commutes with rint: its bounds are
//     integers).  With d = fl(x - zero) common to both forms, fl(d / scale) and fl(d * fl(1 / scale)) differ by at most
//     3 * 2^-24 * |quotient| <= 4.6e-5 inside the clamp range;
//   * rint is the magic-number addition (1.5 * 2^23: round-to-nearest-even happens in the FADD, the code is the low mantissa
//     byte of the sum);
//   * only when the clamped quotient lands within 1e-4 of a bucket boundary (k + 0.5) is the true division evaluated - out of
//     line, so that the division's slow path is not inlined dozens of times into the epilogues.
constexpr float kRintMagic = 12582912.f;
constexpr float kQuantNear = 0.4999f;
__device__ __noinline__ float quant_code_exact(float d, float scale) { return fminf(fmaxf(rintf(__fdiv_rn(d, scale)), 0.f), 255.f); }
// u = clamped quotient + magic (low mantissa byte = the code unless the return value, the distance from the rounded value,
// exceeds kQuantNear)
__device__ __forceinline__ float quant_round(float x, const ActQuant& q, float& u) {
  const float t = __fmul_rn(__fsub_rn(x, q.zero), q.inv);
  const float tc = fminf(fmaxf(t, 0.f), 255.f);
  u = __fadd_rn(tc, kRintMagic);
  return fabsf(__fsub_rn(tc, __fsub_rn(u, kRintMagic)));
}
__device__ __forceinline__ float quant_code(float x, const ActQuant& q) {
  float u;
  const float e = quant_round(x, q, u);
  float v = __fsub_rn(u, kRintMagic);
  if (e > kQuantNear) v = quant_code_exact(__fsub_rn(x, q.zero), q.scale);
  return v;
}
// eight values at once: u[k] - kRintMagic = code k.  One boundary test for the group (a warp takes the branch when any of its
// 256 values is near a boundary: ~5 %); fully unrolled - a dynamically indexed u[] / x[] would move both arrays to local memory.
__device__ __forceinline__ void quant_round8(const float* x, const ActQuant& q, float* u) {
  float e = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) e = fmaxf(e, quant_round(x[k], q, u[k]));
  if (e > kQuantNear) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float uu;
      if (quant_round(x[k], q, uu) > kQuantNear) u[k] = __fadd_rn(quant_code_exact(__fsub_rn(x[k], q.zero), q.scale), kRintMagic);
    }
  }
}
__device__ __forceinline__ float fake_quant(float x, const ActQuant& q) {
  if (q.mode == 2) return __fadd_rn(__fmul_rn(quant_code(x, q), q.scale), q.zero);
  if (q.mode == 1) {
    float v = rintf(__fdiv_rn(x, q.scale));
    v = fminf(fmaxf(v, -128.f), 127.f);
    return __fmul_rn(v, q.scale);
  }
  return x;
}
#define HDRTV_CUDA_OK(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      snprintf(hdrtv::g_err, sizeof(hdrtv::g_err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(_e));                                                        \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

extern char g_err[512];

}  // namespace hdrtv
