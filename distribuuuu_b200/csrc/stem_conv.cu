// Forward of the space-to-depth stem (7x7 / stride 2 / pad 3 over 3 channels == 4x4 / stride 1 over the 16-channel S image,
// extras.cu: stem_s2d_kernel) with the A tile GATHERED by cp.async instead of loaded by im2col TMA.
//
// Status: an experiment kept behind B200_STEM_GATHER=1 (default off) -- measured 324 us against 358 us for the generic
// kernel on the 256 x 3 x 224 x 224 stem, and the reason it cannot do much better is worth recording:
//
// As a 4x1 convolution over the overlapping [N, P+3, Q, 64] view of S (ops/native.py: StemConvFn) the generic kernel pulls
// every 32-byte S pixel four times over the L2 -> SM fabric (4 taps x 16 KB per 128 output pixels).  Here one tile is one
// output row (n, p): thread q of four producer warps copies, for each of the four filter rows, the 128 contiguous bytes
// S[n, p + a, q .. q + 3, :] into row q of a 128B-swizzled K-major sub-tile (a slot of an 8-slot ring, six slots in flight)
// with eight `cp.async.ca` of 16 bytes.  Three producers were measured (profiles/r2/ncu_prof_stem_conv_expand.txt):
// cp.async with two 64 KB stages 325 us, this ring 324 us, and one linear bulk copy of the four S rows per tile expanded
// smem -> smem 358 us.  The time does not depend on how the A tile gets into shared memory, the tensor pipe is 16 % active
// and the epilogue warps spend 33 % of their samples waiting for an accumulator, so the pace is set between "A tile in
// shared memory" and "accumulator complete".  The working hypothesis is shared-memory bandwidth: with N = 64 an SS-mode
// tcgen05.mma (128 x 64 x 16) reads 4 KB of A + 2 KB of B for 32 tensor-core cycles, and a tile costs 96 KB of operand
// reads + 64 KB of A-tile writes + 48 KB of epilogue staging -- 1.6k cycles at 128 B/clk against 3.3k measured.  The halo
// kernel (also N = 64) shows the same picture (35 % tensor activity).  An A operand read straight from the staged rows
// (32-byte-swizzle descriptors, one MMA per (filter row, column) pair, no expanded copy) or N >= 128 would test it.
//
// The packed weights (4 x 8 KB) stay resident; MMA issue, TMEM double buffering, the two-group epilogue with predicated
// coalesced row stores, the BN statistics and the SyncBN flag at the tail are those of conv3x3_halo.cu.
//
// Reference: models/resnet.py:194 (nn.Conv2d(3, 64, 7, 2, 3) -> cuDNN); SURVEY G1.
#include "common.cuh"
#include "stem_conv.h"

namespace b200 {

namespace stemk {
constexpr int BM = 128, BN = 64, BK = 64, kTaps = 4;
constexpr int kTapBytes = BN * BK * 2;            // 8 KB: one filter row's [64 x 64] weight tile
constexpr int kWBytes = kTaps * kTapBytes;        // 32 KB resident weights
constexpr int kSubBytes = BM * 128;               // 16 KB: A sub-tile of one filter row = one slot of the ring
constexpr int kRing = 8;                          // slots (two tiles' worth); the producers run kAhead slots ahead of the MMA
constexpr int kAhead = 6;
constexpr int kEpiWarps = 8, kProdWarps = 4;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = 64 + kEpiWarps * 32 + kProdThreads;
constexpr int kStagingBytes = kEpiWarps * 4096;
constexpr int kSmem = kWBytes + kRing * kSubBytes + kStagingBytes + 1024 + 256;
}  // namespace stemk
using namespace stemk;

__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
stem_conv_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ StemConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + kWBytes;
  uint8_t* s_out = s_a + kRing * kSubBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + kStagingBytes);
  uint64_t* full_bar = bars;                 // [kRing]   kProdThreads arrivals
  uint64_t* empty_bar = bars + kRing;        // [kRing]   tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kRing;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* w_bar = tmem_empty + 2;          // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // rows q >= Q of the A tiles are never written by the producers: zero them once (the MMA reads all 128 rows)
  for (int i = threadIdx.x; i < kRing * kSubBytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < kRing; ++i) { mbar_init(smem_u32(&full_bar[i]), kProdThreads); mbar_init(smem_u32(&empty_bar[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tmem_full[i]), 1); mbar_init(smem_u32(&tmem_empty[i]), kEpiWarps / 2); }
    mbar_init(smem_u32(w_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_ptr), 2 * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncwarp();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== resident weights: four [64 x 64] K-major tiles, once
    if (lane == 0) {
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)kWBytes);
      for (int t = 0; t < kTaps; ++t) tma_load_3d(smem_u32(s_w + t * kTapBytes), &map_w, wb, 0, t, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer: 4 filter rows x 4 K-steps per tile
    if (lane == 0) {
      int acc = 0; uint32_t acc_phase = 0;
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      const uint64_t hi = make_smem_desc_hi_sw128(16, 1024);
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      uint64_t b_desc[kTaps];
#pragma unroll
      for (int t = 0; t < kTaps; ++t) b_desc[t] = hi | (uint64_t)((smem_u32(s_w + t * kTapBytes) >> 4) & 0x3fff);
      int local = 0;
      for (int item = blockIdx.x; item < p.tiles; item += gridDim.x, ++local) {
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        const uint32_t tmem_d = tmem_base + acc * BN;
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
          const int j = local * kTaps + t, slot = j & (kRing - 1);          // (tile, filter row) pair j lives in ring slot j % 8
          mbar_wait(smem_u32(&full_bar[slot]), (uint32_t)((j >> 3) & 1));
          tc_fence_after();
          const uint64_t a_desc = hi | (uint64_t)((smem_u32(s_a + slot * kSubBytes) >> 4) & 0x3fff);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, a_desc + (uint64_t)(2 * k), b_desc[t] + (uint64_t)(2 * k), idesc, (t | k) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[slot]));                           // the slot is free once these four MMAs have read it
        }
        umma_commit(smem_u32(&tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < 2 + kEpiWarps) {
    // =============================== epilogue (conv3x3_halo.cu): TMEM -> bf16 -> swizzled staging -> coalesced row stores
    const int quarter = warp & 3;                 // TMEM lanes [32 q, 32 q + 32)
    const int group = (warp - 2) >> 2;            // 0 / 1: alternate tiles, group g <-> accumulator stage g
    const int acc = group;
    uint32_t acc_phase = 0;
    float st_sum[2] = {0.f, 0.f}, st_sq[2] = {0.f, 0.f};   // lane l: channels 2l, 2l+1
    const bool want_stats = p.stats != nullptr;
    uint8_t* sbuf = s_out + (warp - 2) * 4096;
    const int q = quarter * 32 + lane;            // this lane's output column
    const bool valid = q < p.Q;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    int local = 0;
    for (int item = blockIdx.x; item < p.tiles; item += gridDim.x, ++local) {
      if ((local & 1) != group) continue;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      const long long row_off = valid ? ((long long)item * p.Q + q) * BN : -1ll;   // item = n * P + p: rows of y are contiguous
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + hlf * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
          *reinterpret_cast<uint4*>(sbuf + lane * 128 + (((hlf * 4 + g) ^ (lane & 7)) << 4)) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), g = lane & 7;
        const long long off = __shfl_sync(0xffffffffu, row_off, r);
        const uint4 v = *reinterpret_cast<const uint4*>(sbuf + r * 128 + ((g ^ (r & 7)) << 4));
        if (off >= 0) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + off + g * 8) = v;
      }
      if (want_stats) {
        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const uint32_t wd = *reinterpret_cast<const uint32_t*>(sbuf + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
          const bool ok = (vmask >> r) & 1u;
          const float f0 = ok ? __uint_as_float(wd << 16) : 0.f;
          const float f1 = ok ? __uint_as_float(wd & 0xffff0000u) : 0.f;
          a0 += f0; q0 = fmaf(f0, f0, q0);
          a1 += f1; q1 = fmaf(f1, f1, q1);
        }
        st_sum[0] += a0; st_sum[1] += a1; st_sq[0] += q0; st_sq[1] += q1;
      }
      __syncwarp();
    }
    if (want_stats) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        atomicAdd(p.stats + 2 * lane + h2, st_sum[h2]);
        atomicAdd(p.stats + BN + 2 * lane + h2, st_sq[h2]);
      }
      if (p.peer.world > 1) __threadfence();
    }
  } else {
    // =============================== producers: thread q gathers row q of the four sub-tiles (8 x 16 B per filter row)
    const int q = threadIdx.x - (64 + kEpiWarps * 32);
    const bool has_row = q < p.Q;
    const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(p.s);
    const int my_tiles = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * kTaps;           // (tile, filter row) pairs of this CTA, in MMA order
    const uint32_t dst_row = smem_u32(s_a + q * 128);
    auto issue = [&](int j) {
      const int slot = j & (kRing - 1);
      mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)(((j >> 3) & 1) ^ 1));   // the MMAs of pair j - 8 have read the slot
      if (has_row) {
        const int item = (int)blockIdx.x + (j >> 2) * (int)gridDim.x, a = j & 3;
        const int n = item / p.P, pr = item - n * p.P;
        const __nv_bfloat16* src = s + (((long long)n * p.Hs + pr + a) * p.Ws + q) * 16;
#pragma unroll
        for (int c = 0; c < 8; ++c) cp_async16_ca(dst_row + slot * kSubBytes + ((c ^ (q & 7)) << 4), src + c * 8);
      }
    };
    // kAhead pairs in flight: group g is complete when at most kAhead newer groups are pending
    for (int j = 0; j < kAhead; ++j) {
      if (j < total) issue(j);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int j = 0; j < total; ++j) {
      if (j + kAhead < total) issue(j + kAhead);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");   // pair j has landed
      fence_proxy_async_smem();                   // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(smem_u32(&full_bar[j & (kRing - 1)]));
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
  // SyncBN: the last CTA out announces this layer's statistics exchange (see conv_gemm.cu)
  if (p.peer.world > 1 && threadIdx.x == 0) {
    const int done = atomicAdd(p.peer.ticket + 2, 1);
    if (done == (int)gridDim.x - 1) {
      p.peer.ticket[2] = 0;
      const uint32_t e = *reinterpret_cast<volatile uint32_t*>(p.peer.epoch_dev) + 1u;
      *reinterpret_cast<volatile uint32_t*>(p.peer.epoch_dev) = e;
      __threadfence_system();
      for (int r = 0; r < p.peer.world; ++r)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.peer.signal_pads[r] + p.peer.slot_base + p.peer.rank), "r"(e) : "memory");
    }
  }
}

}  // namespace b200

extern "C" int b200_stem_conv_launch(const CUtensorMap* map_w, const StemConvParams* p, int grid, cudaStream_t stream) {
  using namespace b200;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  stem_conv_kernel<<<grid, kThreads, kSmem, stream>>>(*map_w, *p);
  return (int)cudaGetLastError();
}
