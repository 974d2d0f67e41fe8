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
};
__device__ __forceinline__ float quant_code(float x, const ActQuant& q) {      // mode 2: the uint8 code as a float in [0, 255]
  float v = rintf(__fdiv_rn(__fsub_rn(x, q.zero), q.scale));
  return fminf(fmaxf(v, 0.f), 255.f);
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
// FP16 tensor path: the tensor a W8A8 layer quantises is the fp16 OUTPUT of its producer (the reference calls x.float() on
// it), and the de-quantised value is cast back to the compute dtype, fp16 (:358).
__device__ __forceinline__ float fake_quant_h(float x, const ActQuant& q) {
  if (q.mode == 0) return x;
  return __half2float(__float2half_rn(fake_quant(__half2float(__float2half_rn(x)), q)));
}
__device__ __forceinline__ uint32_t quant_u8_h(float x, const ActQuant& q) {   // uint8 code of the fp16-rounded value
  return static_cast<uint32_t>(quant_code(__half2float(__float2half_rn(x)), q));
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
