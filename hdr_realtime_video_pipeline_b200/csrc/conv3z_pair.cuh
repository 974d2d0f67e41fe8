// CondNet{2,3,4}.0 as ONE N = 192 convolution on a CTA pair (tcgen05.mma.cta_group::2, M = 256).
//
// The three stride-2 3x3 64 -> 64 convs of the LE condition pyramid (HDRUNet3T1_arch.py:47-62: CondNet2[0], CondNet3[0],
// CondNet4[0], each LeakyReLU(0.1)) read the same tensor `cond`.  Issued one conv at a time (N = 64, conv_p8_kernel with
// zsplit) an M = 128 x K = 16 MMA holds the tensor pipe for 48 cycles and computes for 32, every 4 KB A tile is fetched
// from shared memory three times (once per conv) and every `cond` row crosses L2 -> SM three times: ncu shows 56 % pipe
// busy, 35 % math, 52 % of the MMA shared-memory operand bandwidth (profiles/r2_ncu_4k.md).  Stacking the three weight
// sets along N gives N = 192, where the pipe runs at the dense rate (N/2 = 96 cycles per K step), but the stacked weights
// are 222 KB.  A CTA pair splits them: with cta_group::2 each CTA supplies its own 128 A rows (its own 128-pixel strip)
// and HALF of the B columns (96 of 192, 111 KB), and receives all 192 accumulator columns for its 128 pixels in its own
// TMEM.  Every `cond` row is fetched once, every A tile is read once per K step.
//
// Dataflow.  A cluster of two CTAs owns two neighbouring strips and walks a contiguous range of the strip-pair-major
// (pair, output row) list, like the chain and two-conv kernels (one balanced wave over all SMs).  Input rows stream
// through a 3-slot ring per CTA (34.8 KB per parity-split 64-channel row) and each row is consumed ONCE, in order:
//   even row 2m   : dy = 2 of output row m-1 (completes it: bias step, commit to the epilogues), then dy = 0 of row m
//   odd row 2m+1  : dy = 1 of output row m
// so a slot is released by the commit right behind its 12 or 24 MMAs and the other two slots are always in flight.
//
// Synchronisation across the pair (leader = cluster rank 0 issues every MMA):
//   full   : each CTA's TMA producer completes its own full barrier; the follower's warp 1 relays that to the leader's
//            peer_full barrier with a remote mbarrier.arrive (the shared::cluster address of the leader's barrier is the
//            local address with the peer bit cleared)
//   empty  : tcgen05.commit.cta_group::2 ... multicast::cluster to the empty barrier of both CTAs
//   tfull  : the same multicast commit to both CTAs' accumulator-full barriers
//   tempty : the 12 epilogue warps of BOTH CTAs arrive on the leader's barrier (local / remote arrive)
#pragma once
#include "conv_p8.cuh"

namespace hdrtv {

constexpr int kPairThreads = 32 * 14;          // TMA producer, MMA issuer / relay, 12 epilogue warps (3 convs x 4 lane quadrants)
constexpr int kPairRing = 3;
constexpr int kPairN = 192, kPairNHalf = 96;
constexpr int kPairSteps = 36;                 // 3 dy x 3 dx x 4 K = 16 steps (64 input channels)
constexpr int kPairSlotBytes = 16 * kPlaneBytes;
constexpr int kPairWBytes = (kPairSteps + 1) * kPairNHalf * 32;      // per CTA: its 96 columns of every K step + the bias step
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                      // clears the CTA-pair peer bit of a shared::cluster address

struct PairParams {
  const uint4* in;             // `cond`: P8, parity-split, 8 channel-chunk planes per row
  long in_row_entries;
  uint32_t in_wp;              // entries per plane (both parities)
  const uint4* wpk[2];         // per cluster rank: [37 steps][96 columns][16 k] fp16, K-major core matrices
  int Ho, Wo;                  // output size
  int strips, pairs;           // 128-pixel output strips, strip pairs
  P8 out[3];                   // CondNet2.0 / CondNet3.0 / CondNet4.0 outputs (64 channels; layouts may differ)
  int pf_rows;                 // > 0: input rows requested into L2 this many rows ahead of the ring (cp.async.bulk.prefetch.L2)
  int* err;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in the LEADER CTA of the pair (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t result_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16_m256(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Contiguous range of (strip pair, output row) items of this cluster, walked as segments that stay inside one pair.
struct PairWalk {
  long lo, hi;
  int Ho;
  __device__ __forceinline__ PairWalk(int pairs, int Ho_) : Ho(Ho_) {
    const long total = static_cast<long>(pairs) * Ho_;
    const long c = cluster_id_x(), nc = cluster_nid_x();
    lo = total * c / nc;
    hi = total * (c + 1) / nc;
  }
  __device__ __forceinline__ bool next(int& pair, int& r0, int& n) {
    if (lo >= hi) return false;
    pair = static_cast<int>(lo / Ho);
    r0 = static_cast<int>(lo - static_cast<long>(pair) * Ho);
    n = static_cast<int>(min(static_cast<long>(Ho - r0), hi - lo));
    lo += n;
    return true;
  }
};

__global__ void __launch_bounds__(kPairThreads, 1) conv3z_pair_kernel(const __grid_constant__ PairParams p) {
  constexpr int SPD = 12;                               // K = 16 steps per input row (3 dx x 4 channel-chunk pairs)
  constexpr uint32_t kTmemCols = 512;                   // 2 accumulator stages x 192 columns
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int i) { return bar0 + 8u * i; };
  auto empty_bar = [&](int i) { return bar0 + 8u * (4 + i); };
  auto peer_full_bar = [&](int i) { return bar0 + 8u * (8 + i); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (12 + i); };
  auto tempty_bar = [&](int i) { return bar0 + 8u * (14 + i); };
  const uint32_t wfull_bar = bar0 + 8u * 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 18);
  uint8_t* ones = smem + 512;
  uint8_t* wsm = smem + kSmemHeader;
  uint8_t* ring = wsm + ((kPairWBytes + 127) & ~127);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPairRing; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
      mbar_init(peer_full_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 24);                     // 12 epilogue warps of each CTA
    }
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kPlaneEntries) {
    reinterpret_cast<uint4*>(ones)[threadIdx.x - 64] = make_uint4(0x3C003C00u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();           // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PairWalk walk(p.pairs, p.Ho);
  int pair, r0, n;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs: own strip, own half of B)
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, kPairWBytes);
      bulk_g2s(smem_u32(wsm), p.wpk[rank], kPairWBytes, wfull_bar);
      grid_dep_wait();          // static weights are on their way; activations need the previous kernel finished
      uint32_t slot = 0, ph = 1;
      const uint32_t wp = p.in_wp, half = p.in_wp >> 1;
      while (walk.next(pair, r0, n)) {
        // an odd strip count leaves the last pair's second CTA without a strip: it re-reads the last real one (in bounds)
        // and stores nothing
        const int strip = min(2 * pair + static_cast<int>(rank), p.strips - 1);
        const uint4* src = p.in + static_cast<long>(2 * r0) * p.in_row_entries + strip * kTileM;
        auto prefetch_row = [&](int q) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            unsigned long long a;
            asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"((c >> 1) * wp + (c & 1) * half), "l"(p.in + (static_cast<long>(2 * r0 + q) * p.in_row_entries + strip * kTileM)));
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(kPlaneBytes) : "memory");
          }
        };
        const int pf = p.pf_rows;
        if (pf > 0)
          for (int q = kPairRing; q < min(kPairRing + pf, 2 * n + 1); ++q) prefetch_row(q);
        for (int q = 0; q < 2 * n + 1; ++q) {
          if (pf > 0 && q + kPairRing + pf < 2 * n + 1) prefetch_row(q + kPairRing + pf);
          mbar_wait(empty_bar(slot), ph, p.err, 41);
          mbar_expect_tx(full_bar(slot), kPairSlotBytes);
          const uint32_t dst = smem_u32(ring) + slot * kPairSlotBytes;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            unsigned long long a;
            asm volatile("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"((c >> 1) * wp + (c & 1) * half), "l"(src));
            bulk_g2s(dst + c * kPlaneBytes, reinterpret_cast<const void*>(a), kPlaneBytes, full_bar(slot));
          }
          src += p.in_row_entries;
          if (++slot == kPairRing) { slot = 0; ph ^= 1; }
        }
      }
      grid_dep_launch();
    }
  } else if (warp == 1 && !leader) {
    // ------------------------------------------------------------------ follower: relay "row landed" to the leader
    // this CTA's half of B first: the first relayed row then also tells the leader that both weight halves are resident
    mbar_wait(wfull_bar, 0, p.err, 48);
    uint32_t slot = 0, ph = 0;
    while (walk.next(pair, r0, n)) {
      for (int q = 0; q < 2 * n + 1; ++q) {
        mbar_wait(full_bar(slot), ph, p.err, 42);
        if (lane == 0) mbar_arrive_leader(peer_full_bar(slot));
        __syncwarp();
        if (++slot == kPairRing) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool lead = elect_one();         // the issuing lane, elected once
    // ------------------------------------------------------------------ leader: MMA issuer for the pair
    mbar_wait(wfull_bar, 0, p.err, 43);      // the follower's half: implied by its first relayed row (see the relay warp)
    constexpr uint32_t idesc = make_idesc_f16_m256(kPairN);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
    auto mkdesc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 128);
    constexpr uint32_t b_lbo = static_cast<uint32_t>(kPairNHalf) << 16;   // (96 x 16 bytes) >> 4 in the LBO field
    constexpr uint32_t b_step = static_cast<uint32_t>(kPairNHalf) * 2;    // (96 x 32 bytes) >> 4
    constexpr uint32_t a_lbo = (kind_a_lbo(IN_PAR3x3S2) >> 4) << 16;
    const uint32_t b_lo0 = (smem_u32(wsm) >> 4) | b_lbo;
    constexpr uint32_t slot16 = kPairSlotBytes >> 4;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    uint32_t slot = 0, ph = 0;
    int R = 0;                                                  // output rows started so far (accumulator stage / phase)
    while (walk.next(pair, r0, n)) {
      for (int q = 0; q < 2 * n + 1; ++q) {
        const int m = q >> 1;
        const bool even = (q & 1) == 0;
        const bool fin = even && m >= 1;                        // dy = 2 of output row m-1
        const bool start = even && m < n;                       // dy = 0 of output row m
        const int row_cur = R + m;                              // global index of output row m of this segment
        if (start) mbar_wait(tempty_bar(row_cur & 1), ((row_cur >> 1) & 1) ^ 1, p.err, 44);
        mbar_wait(full_bar(slot), ph, p.err, 45);
        mbar_wait(peer_full_bar(slot), ph, p.err, 46);
        tc_fence_after();
        const uint32_t a16 = ring16 + slot * slot16;
        if (lead) {
          if (fin) {
            const uint32_t d = tmem_base + static_cast<uint32_t>((row_cur - 1) & 1) * kPairN;
            static_for<0, SPD>([&](auto ic) {
              constexpr int i = decltype(ic)::value;
              constexpr uint32_t a_off16 = kind_a_off(IN_PAR3x3S2, 8, i) >> 4;
              tc_mma_f16_pair(d, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (2 * SPD + i) * b_step), idesc, 1u);
            });
            tc_mma_f16_pair(d, ones_desc, mkdesc(b_lo0 + kPairSteps * b_step), idesc, 1u);      // + bias
            tc_commit_pair(tfull_bar((row_cur - 1) & 1));
          }
          if (start) {
            const uint32_t d = tmem_base + static_cast<uint32_t>(row_cur & 1) * kPairN;
            static_for<0, SPD>([&](auto ic) {
              constexpr int i = decltype(ic)::value;
              constexpr uint32_t a_off16 = kind_a_off(IN_PAR3x3S2, 8, i) >> 4;
              tc_mma_f16_pair(d, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + i * b_step), idesc, i ? 1u : 0u);
            });
          }
          if (!even) {
            const uint32_t d = tmem_base + static_cast<uint32_t>(row_cur & 1) * kPairN;
            static_for<0, SPD>([&](auto ic) {
              constexpr int i = decltype(ic)::value;
              constexpr uint32_t a_off16 = kind_a_off(IN_PAR3x3S2, 8, i) >> 4;
              tc_mma_f16_pair(d, mkdesc((a16 + a_off16) | a_lbo), mkdesc(b_lo0 + (SPD + i) * b_step), idesc, 1u);
            });
          }
          tc_commit_pair(empty_bar(slot));                      // every input row is read exactly once
        }
        __syncwarp();
        if (++slot == kPairRing) { slot = 0; ph ^= 1; }
      }
      R += n;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: conv z, TMEM lane quadrant lg
    grid_dep_wait();            // overwrites buffers the previous kernel may still read
    const int lg = warp & 3;
    const int z = (warp - 2) >> 2;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(z * 64);
    int R = 0;
    while (walk.next(pair, r0, n)) {
      const int x = (2 * pair + static_cast<int>(rank)) * kTileM + lg * 32 + lane;
      const bool xin = x < p.Wo;
      ColRef out;
      out.init(p.out[z], x);
      for (int t = 0; t < n; ++t) {
        const int row = R + t, stage = row & 1, oy = r0 + t;
        mbar_wait(tfull_bar(stage), (row >> 1) & 1, p.err, 47);
        tc_fence_after();
        float v[64];
        tmem_ld_cols<64>(tlane + stage * kPairN, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(stage));
        if (xin) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaxf(v[c * 8 + k], 0.1f * v[c * 8 + k]);      // LeakyReLU(0.1)
            *out.at(oy, c) = pack8(a);
          }
        }
      }
      R += n;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();           // both CTAs are done with the pair's TMEM and with each other's barriers
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

inline size_t pair_smem_bytes() {
  return kSmemHeader + ((kPairWBytes + 127) & ~127) + static_cast<size_t>(kPairRing) * kPairSlotBytes;
}

}  // namespace hdrtv
