// 3x3 / stride 1 / pad 1 convolution with ONE activation load per tile (halo reuse), for the layers whose whole weight
// tensor fits in shared memory (Cin = Cout = 64: ResNet-18/34/50 layer1) -- forward and data gradient.
//
// Why a second kernel: conv_gemm.cu feeds every filter tap with its own im2col TMA load, i.e. it pulls the activation
// tile nine times from L2.  The B200's L2 -> SM fabric delivers ~7.5 TB/s in total (measured: profiles/r2), and the
// 64 -> 64 @ 56^2 layer moves 9 x 16 KB per 128-pixel tile over it for only 1152 tensor-core cycles of work: it ran at
// 163 us where the tensor cores need 27 us and HBM 31 us.  Here a tile is `rpt` = floor(128 / (W + 2)) full rows of the
// PADDED raster (rows of W + 2 positions): the input of tap (r, s) for position p is padded position p + r (W + 2) + s, so
// ONE TMA load of the (rpt + 2) x (W + 2) patch around the tile (a plain tiled box starting at column -1 / row h0 - 1:
// out-of-bounds zero fill supplies the padding) serves all nine taps, each as an MMA whose A descriptor simply starts
// r (W + 2) + s rows further down.  (A first version loaded the halo in im2col mode: 248 "pixels" per load ran at ~16
// cycles per pixel inside the TMA unit -- 3.9k cycles per tile against 1152 of MMA; tiled boxes move whole rows.)  tools/probes/umma_shift_probe.cu established
// that tcgen05.mma accepts an A tile starting at any 128-byte row of a 128B-swizzled buffer (the swizzle is a function
// of the absolute shared-memory address).  The 2 (W + 2) / (H + 2) garbage positions per row / image cost 7 % extra MMA
// work at 56^2; the epilogue drops them (rows are written from the staging tile with per-row predicates -- a TMA box
// cannot express the variable-length image-row segments of a 32-position chunk -- and the BN statistics mask them).
//
// Replaces the cuDNN call of reference resnet.py:36-54 for this layer; VERDICT r1 "next round" item 1(c).
#include "common.cuh"
#include "conv3x3_halo.h"

namespace b200 {

namespace halo {
constexpr int BM = 128, BN = 64, BK = 64;
constexpr int kTapBytes = BN * BK * 2;          // 8 KB: one tap's [64 x 64] weight tile
constexpr int kWBytes = 9 * kTapBytes;          // 72 KB resident weights
constexpr int kStageBytes = 32 * 1024;          // halo stage (<= 256 rows of 128 B)
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kStagingBytes = kEpiWarps * 4096; // one private 32-row x 64-col tile per epilogue warp
constexpr int kSmem = kWBytes + kStages * kStageBytes + kStagingBytes + 1024 + 256;
}  // namespace halo
using namespace halo;

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_y, const __grid_constant__ Conv3x3HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + kWBytes;
  uint8_t* s_out = s_a + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + kStagingBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* w_bar = tmem_empty + 2;          // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = p.W + 2;
  const int halo_rows = p.halo_rows;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w);
    for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
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
    // =============================== TMA producer: resident weights once, then one halo tile per item
    if (lane == 0) {
      const uint32_t wb = smem_u32(w_bar);
      mbar_expect_tx(wb, (uint32_t)kWBytes);
      for (int t = 0; t < 9; ++t) tma_load_3d(smem_u32(s_w + t * kTapBytes), &map_w, wb, 0, t, 0);
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.tiles; item += gridDim.x) {
        const int n = item / p.tiles_per_img, h0 = (item - n * p.tiles_per_img) * p.rpt;   // first output row of the tile
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t bar = smem_u32(&full_bar[stage]);
        mbar_expect_tx(bar, (uint32_t)(halo_rows * 128));
        // box = (64 channels, W + 2 columns from -1, rpt + 2 rows from h0 - 1, one image); zero fill outside the image
        tma_load_4d(smem_u32(s_a + stage * kStageBytes), &map_x, bar, 0, -1, h0 - 1, n);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer: 9 taps x 4 K-steps per tile, A descriptor shifted by r * Wp + s rows
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      const uint64_t a_hi = make_smem_desc_hi_sw128(16, 1024);
      const uint64_t b_hi = p.dgrad ? make_smem_desc_hi_sw128(8192, 1024) : make_smem_desc_hi_sw128(16, 1024);
      const uint32_t b_step = p.dgrad ? 128u : 2u;
      const uint32_t idesc = make_idesc_bf16(BM, BN, 0, p.dgrad ? 1u : 0u);
      // Per-tap descriptor parts are loop invariants: with N = 64 an MMA is only 32 tensor-core cycles, so the single
      // issuing thread -- not the tensor pipe -- sets the pace of this kernel; keep its per-MMA work to one 64-bit add.
      uint64_t a_off[9], b_desc[9];               // A: start-address advance in 16-byte units (rows are 128 B)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        a_off[t] = (uint64_t)(((t / 3) * Wp + (t % 3)) * 8);
        const int wtap = p.dgrad ? (8 - t) : t;   // data gradient: flipped filter
        b_desc[t] = b_hi | (uint64_t)((smem_u32(s_w + wtap * kTapBytes) >> 4) & 0x3fff);
      }
      for (int item = blockIdx.x; item < p.tiles; item += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const uint64_t a0 = a_hi | (uint64_t)((smem_u32(s_a + stage * kStageBytes) >> 4) & 0x3fff);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_desc = a0 + a_off[t];
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, a_desc + (uint64_t)(2 * k), b_desc[t] + (uint64_t)(b_step * k), idesc, (t | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        umma_commit(smem_u32(&tmem_full[acc]));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> bf16 -> swizzled staging -> predicated row stores (+ statistics)
    // Two groups of four warps take ALTERNATE tiles (group g <-> TMEM accumulator stage g), so two tiles are in the
    // epilogue at any time: with one tile at a time the ~2k-cycle latency chain tcgen05.ld -> pack -> st.shared -> store ->
    // statistics set the pace (ncu: 3.9k cycles per tile for 1152 cycles of MMA).  A warp owns 32 rows x 64 columns.
    const int quarter = warp & 3;                 // TMEM lanes [32 q, 32 q + 32)
    const int group = (warp - 2) >> 2;            // 0 / 1
    const int acc = group;
    uint32_t acc_phase = 0;
    float st_sum[2] = {0.f, 0.f}, st_sq[2] = {0.f, 0.f};   // lane l: columns 2l, 2l+1
    const bool want_stats = p.stats != nullptr;
    uint8_t* sbuf = s_out + (warp - 2) * 4096;    // private 32-row x 128-byte staging tile
    int local = 0;
    for (int item = blockIdx.x; item < p.tiles; item += gridDim.x, ++local) {
      if ((local & 1) != group) continue;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      // this lane's row of the tile: position m -> (row hr of the tile, column wp) and validity
      const int n = item / p.tiles_per_img, h0 = (item - n * p.tiles_per_img) * p.rpt;
      const int m = quarter * 32 + lane;
      const int hr = m / Wp, wp = m - hr * Wp;
      const int hp = h0 + hr;
      const bool valid = (hr < p.rpt) && (hp < p.H) && (wp < p.W);
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      const long long row_off = valid ? (((long long)n * p.H + hp) * p.W + wp) * BN : -1ll;   // element offset in y
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
      // the accumulator stage is free as soon as every lane has its values in registers / shared memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      // 4 rows (4 x 128 B, fully coalesced) per instruction, garbage rows predicated off
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
      __syncwarp();                               // the staging tile is rewritten by the next tile of this warp
    }
    if (want_stats) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        atomicAdd(p.stats + 2 * lane + h2, st_sum[h2]);
        atomicAdd(p.stats + BN + 2 * lane + h2, st_sq[h2]);
      }
      if (p.peer.world > 1) __threadfence();
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
      __threadfence_system();   // ONE fence, then relaxed flag stores: a st.release per peer serialises `world` NVLink round trips
      for (int r = 0; r < p.peer.world; ++r)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.peer.signal_pads[r] + p.peer.slot_base + p.peer.rank), "r"(e) : "memory");
    }
  }
}

}  // namespace b200

extern "C" int b200_conv3x3_halo_launch(const CUtensorMap* map_x, const CUtensorMap* map_w, const CUtensorMap* map_y,
                                        const Conv3x3HaloParams* p, int grid, cudaStream_t stream) {
  using namespace b200;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  conv3x3_halo_kernel<<<grid, kThreads, kSmem, stream>>>(*map_x, *map_w, *map_y, *p);
  return (int)cudaGetLastError();
}
