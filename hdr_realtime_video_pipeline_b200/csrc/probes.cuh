// Micro-probes of the sm_100a tensor path (design evidence, not product code).  Every probe reports cycles per
// iteration measured with clock64 inside one CTA; `blocks` CTAs run concurrently to expose per-SM sharing.
//   probe 0  tcgen05.mma issue rate, A and B from shared memory (SS), M = 128
//   probe 1  tcgen05.mma with the A operand in TMEM (TS), M = 128
//   probe 2  tcgen05.mma SS with M = 64
//   probe 3  tcgen05.ld throughput: `nwarps` warps each draining 64 fp32 columns of their lane quadrant per iteration
//   probe 4  layer-chain round trip: MMA(N=64, `nmma` K-steps) -> commit -> epilogue warps (tcgen05.ld 64 cols,
//            fp16 pack, 8 x st.shared, fence.proxy.async) -> mbarrier -> next MMA; `groups` independent row slots
#pragma once
#include "chain_p8.cuh"

namespace hdrtv {

__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t m, uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct ProbeArgs {
  int kind, n, iters, nwarps, nmma, groups;
  long long* cycles;
  long long* trace;   // kind 4, block 0: [iter < 16][group < 4][event < 4] clock64 stamps
};

template <int KIND>
__global__ void __launch_bounds__(1024) probe_kernel(const ProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int g = 0; g < 4; ++g) {
      mbar_init(smem_u32(&bars[g]), 1);          // tfull[g]
      mbar_init(smem_u32(&bars[4 + g]), 4);      // afull[g]
    }
    mbar_init(smem_u32(&bars[8]), 1);
    for (int g = 9; g < 14; ++g) mbar_init(smem_u32(&bars[g]), 1);
    mbar_fence_init();
  }
  constexpr uint32_t kCols = (KIND == 0 || KIND == 2 || KIND >= 5) ? 256 : 512;   // 256: two CTAs can share an SM
  if (warp == 0) tmem_alloc(smem_u32(&tslot), kCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 48 * 1024;
  const int N = a.n;

  if constexpr (KIND <= 2) {
    if (threadIdx.x == 0) {
      const uint32_t M = KIND == 2 ? 64 : 128;
      const uint32_t idesc = make_idesc_f16(M, N);
      const uint64_t bd = make_smem_desc(b0, N * 16, 128);
      uint64_t adv[4];
      uint32_t atm[4], dv[4];
      for (int k = 0; k < 4; ++k) {
        adv[k] = make_smem_desc(a0 + k * 16, kPlaneBytes, 128);
        atm[k] = tm + 480 + k * 8;                 // TS: A = 128 lanes x 8 columns (16 halves) per K step
        dv[k] = tm + ((2 * N <= 256) ? (k & 1) * N : 0);
      }
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (KIND == 1) tc_mma_f16_ts(dv[k], atm[k], bd, idesc, 1u);
          else tc_mma_f16(dv[k], adv[k], bd, idesc, 1u);
        }
      }
      tc_commit(smem_u32(&bars[8]));
      mbar_wait(smem_u32(&bars[8]), 0, nullptr, 0);
      a.cycles[blockIdx.x] = clock64() - t0;
    }
  } else if constexpr (KIND == 3) {
    if (warp < a.nwarps) {
      const uint32_t tl = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      float acc = 0.f;
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        float v[64];
        const uint32_t col = (i * 64) & 511;
        if (N == 64) {
          tmem_ld32_async(tl + col, reinterpret_cast<uint32_t*>(v));
          tmem_ld32_async(tl + col + 32, reinterpret_cast<uint32_t*>(v) + 32);
        } else if (N == 16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_ld16_async(tl + col + 16 * q, reinterpret_cast<uint32_t*>(v) + 16 * q);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) tmem_ld8_async(tl + col + 8 * q, reinterpret_cast<uint32_t*>(v) + 8 * q);
        }
        tc_wait_ld();
        acc += v[0] + v[63];
      }
      const long long t1 = clock64();
      if (lane == 0 && warp == 0) a.cycles[blockIdx.x] = t1 - t0;
      if (acc == 12345.f) a.cycles[0] = 0;
    }
  } else if constexpr (KIND == 4) {
    const int G = a.groups;
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_f16(128, 64);
        const uint64_t bd = make_smem_desc(b0, 64 * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < a.iters; ++i) {
          for (int g = 0; g < G; ++g) {
            if (i > 0) {
              mbar_wait(smem_u32(&bars[4 + g]), (i - 1) & 1, nullptr, 0);
              tc_fence_after();
            }
            if (a.trace && blockIdx.x == 0 && i < 16) a.trace[(i * 4 + g) * 4 + 0] = clock64();
            const uint64_t ad = make_smem_desc(a0 + g * kTileBytes, kPlaneBytes, 128);
            for (int k = 0; k < a.nmma; ++k) tc_mma_f16(tm + g * 64, ad + ((k & 3) * ((2 * kPlaneBytes) >> 4)), bd, idesc, k > 0);
            tc_commit(smem_u32(&bars[g]));
            if (a.trace && blockIdx.x == 0 && i < 16) a.trace[(i * 4 + g) * 4 + 1] = clock64();
          }
        }
        for (int g = 0; g < G; ++g) mbar_wait(smem_u32(&bars[4 + g]), (a.iters - 1) & 1, nullptr, 0);
        a.cycles[blockIdx.x] = clock64() - t0;
      }
    } else if (warp <= 4 * G) {
      const int g = (warp - 1) >> 2, lg = warp & 3;
      const uint32_t tl = tm + (static_cast<uint32_t>(lg * 32) << 16) + g * 64;
      uint4* tile = reinterpret_cast<uint4*>(smem + g * kTileBytes) + lg * 32 + lane;
      for (int i = 0; i < a.iters; ++i) {
        if (a.nwarps == 1) {                      // parked wait (suspend-time hint)
          while (!mbar_try_wait_hint(smem_u32(&bars[g]), i & 1, 20000)) {}
        } else if (a.nwarps == 2) {               // spin with nanosleep back-off
          while (!mbar_try_wait(smem_u32(&bars[g]), i & 1)) __nanosleep(64);
        } else {
          mbar_wait(smem_u32(&bars[g]), i & 1, nullptr, 0);
        }
        tc_fence_after();
        if (a.trace && blockIdx.x == 0 && i < 16 && lg == 1 && lane == 0) a.trace[(i * 4 + g) * 4 + 2] = clock64();
        if (a.n > 0) {
          float v[64];
          tmem_ld_cols<64>(tl, v);
          tc_fence_before();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(v[c * 8 + k], 0.1f * v[c * 8 + k]);
            tile[c * kPlaneEntries] = pack8(f);
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[4 + g]));
        if (a.trace && blockIdx.x == 0 && i < 16 && lg == 1 && lane == 0) a.trace[(i * 4 + g) * 4 + 3] = clock64();
      }
    }
  } else if constexpr (KIND == 5) {
    // free-running: nmma MMAs (+ one commit) per iteration, no waits until the end.  a.groups = flag bits:
    //   1 commit per iteration, 2 first MMA of an iteration overwrites (accumulate = 0), 4 alternate accumulators,
    //   8 tcgen05.fence::after_thread_sync per iteration
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_f16(128, 64);
      const uint64_t bd = make_smem_desc(b0, 64 * 16, 128);
      const uint64_t ad = make_smem_desc(a0, kPlaneBytes, 128);
      const int f = a.groups;
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t d = tm + ((f & 4) ? (i & 1) * 64 : 0);
        if (f & 8) tc_fence_after();
        for (int k = 0; k < a.nmma; ++k)
          tc_mma_f16(d, ad + ((k & 3) * ((2 * kPlaneBytes) >> 4)), bd, idesc, (f & 2) ? (k > 0) : 1);
        if (f & 1) tc_commit(smem_u32(&bars[9 + (i & 3)]));    // count-1 barriers nobody waits on
        if (f & 16) (void)mbar_try_wait(smem_u32(&bars[13]), 1);   // completes immediately (fresh barrier, parity 1)
      }
      tc_commit(smem_u32(&bars[8]));
      mbar_wait(smem_u32(&bars[8]), 0, nullptr, 0);
      a.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  else if constexpr (KIND == 6) {
    // as probe 5, but the whole warp runs the (warp-uniform) loop and only the tcgen05 instructions are elected
    if (warp == 0) {
      const uint32_t idesc = make_idesc_f16(128, 64);
      const uint64_t bd = make_smem_desc(b0, 64 * 16, 128);
      const uint64_t ad = make_smem_desc(a0, kPlaneBytes, 128);
      const int f = a.groups;
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        const uint32_t d = tm + ((f & 4) ? (i & 1) * 64 : 0);
        for (int k = 0; k < a.nmma; ++k) {
          const uint64_t adk = ad + ((k & 3) * ((2 * kPlaneBytes) >> 4));
          if (elect_one()) tc_mma_f16(d, adk, bd, idesc, (f & 2) ? (k > 0) : 1);
          __syncwarp();
        }
        if (f & 1) {
          if (elect_one()) tc_commit(smem_u32(&bars[9 + (i & 3)]));
          __syncwarp();
        }
        if (f & 16) (void)mbar_try_wait(smem_u32(&bars[13]), 1);
      }
      if (elect_one()) tc_commit(smem_u32(&bars[8]));
      __syncwarp();
      mbar_wait(smem_u32(&bars[8]), 0, nullptr, 0);
      if (lane == 0) a.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  else if constexpr (KIND == 7) {
    // lean batches: exactly 5 MMAs with loop-invariant descriptors + one commit per iteration.  a.groups flag bits:
    //   1 commit, 2 try_wait on a completed barrier, 4 tcgen05.fence::after_thread_sync, 8 alternate accumulator,
    //   16 first MMA overwrites, 32 descriptor low words re-read from shared memory every iteration
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_f16(128, N);
      const uint64_t bd = make_smem_desc(b0, N * 16, 128);
      uint64_t ad[4];
      for (int k = 0; k < 4; ++k)
        ad[k] = make_smem_desc(a0 + (a.nmma == 1 ? 0 : (a.nmma == 2 ? k * 16 : k * 2 * kPlaneBytes)), kPlaneBytes, 128);
      const uint64_t ones = make_smem_desc(a0 + 40960, 16, 128);
      const int f = a.groups;
      volatile uint32_t* lows = reinterpret_cast<volatile uint32_t*>(smem + 90 * 1024);
      for (int k = 0; k < 4; ++k) lows[k] = static_cast<uint32_t>(ad[k]);
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        if (f & 2) (void)mbar_try_wait(smem_u32(&bars[13]), 1);
        if (f & 4) tc_fence_after();
        const uint32_t d = tm + ((f & 8) ? (i & 1) * 64 : 0);
        if (f & 32) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ad[k] = (ad[k] & 0xFFFFFFFF00000000ull) | lows[k];
        }
        const uint32_t d2 = (f & 64) ? d + 128 : d;        // flag 64: consecutive MMAs alternate between two accumulators
        tc_mma_f16(d, ad[0], bd, idesc, (f & 16) ? 0u : 1u);
        tc_mma_f16(d2, ad[1], bd, idesc, 1u);
        tc_mma_f16(d, ad[2], bd, idesc, 1u);
        tc_mma_f16(d2, ad[3], bd, idesc, 1u);
        tc_mma_f16(d, ones, bd, idesc, 1u);
        if (f & 1) tc_commit(smem_u32(&bars[9 + (i & 3)]));
      }
      tc_commit(smem_u32(&bars[8]));
      mbar_wait(smem_u32(&bars[8]), 0, nullptr, 0);
      a.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  else if constexpr (KIND == 9) {
    // operand-stream cost: an unrolled batch of 12 MMAs per iteration whose A tiles (a.nmma & 1) and B tiles
    // (a.nmma & 2) are all DIFFERENT shared-memory regions, as in a real K loop; accumulators alternate when a.nmma & 4.
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_f16(128, N);
      const int f = a.nmma;
      const uint32_t astep = (f & 1) ? ((2 * kPlaneBytes) >> 4) : 0u;          // 12 x 4352 B = 52 KB
      const uint32_t bstep = (f & 2) ? static_cast<uint32_t>((N * 32) >> 4) : 0u;   // 6 distinct tiles of N*32 B (N=128: 24 KB)
      const uint64_t ad0 = make_smem_desc(a0, kPlaneBytes, 128);
      const uint64_t bd0 = make_smem_desc(smem_u32(smem) + 54 * 1024, N * 16, 128);
      const uint32_t dalt = (f & 4) ? 128u : 0u;
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) tc_mma_f16(tm + ((k & 1) ? dalt : 0u), ad0 + k * astep, bd0 + (k % 6) * bstep, idesc, 1u);
      }
      tc_commit(smem_u32(&bars[8]));
      mbar_wait(smem_u32(&bars[8]), 0, nullptr, 0);
      a.cycles[blockIdx.x] = clock64() - t0;
    }
  }
  else if constexpr (KIND == 8) {
    // latency of the synchronisation primitives an issuing / epilogue warp executes per row (whole warp, dependent
    // chain of `iters` ops).  a.groups selects the op.
    if (warp == 0) {
      const uint32_t done = smem_u32(&bars[13]);       // fresh barrier: parity 1 is "already complete"
      const int op = a.groups;
      uint32_t acc = 0;
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < a.iters; ++i) {
        if (op == 1) acc += mbar_try_wait(done, 1);
        else if (op == 2) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(done), "r"(1u) : "memory");
          acc += ok;
        } else if (op == 3) tc_fence_after();
        else if (op == 4) { if (elect_one()) acc += 1; __syncwarp(); }
        else if (op == 5) mbar_wait(done, 1, nullptr, 0);
        else if (op == 6) { if (lane == 0 && a.trace) a.trace[i & 63] = clock64(); }
        else if (op == 7) { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&bars[9 + (i & 3)])); }
        else if (op == 8) { if (elect_one()) tc_commit(smem_u32(&bars[9 + (i & 3)])); __syncwarp(); }
        else if (op == 9) fence_proxy_async_smem();
        else if (op == 10) tc_fence_before();
        else if (op == 11) { if (lane == 0) acc += mbar_try_wait(done, 1); __syncwarp(); }
      }
      const long long t1 = clock64();
      if (lane == 0) a.cycles[blockIdx.x] = t1 - t0;
      if (acc == 0xFFFFFFFFu) a.cycles[0] = 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, kCols);
}

}  // namespace hdrtv
