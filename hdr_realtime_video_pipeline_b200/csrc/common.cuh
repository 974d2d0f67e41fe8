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
