// Implicit-GEMM convolution / GEMM for sm_100a: TMA (tiled + im2col) -> swizzled smem -> tcgen05.mma with
// fp32 accumulators in TMEM -> tcgen05.ld epilogue.  One persistent, warp-specialised kernel covers
//   fprop : Y[M=pixels, N=Cout]  = sum_taps  X_tap[M, Cin]   * W_tap[Cout, Cin]^T      (+ BN statistics / bias)
//   dgrad : dX[M=pixels, N=Cin]  = sum_taps dY_tap[M, Cout]  * W_flip(tap)[Cout, Cin]
//   wgrad : dW[M=Cout, tap, N=Cin] += sum_pixels dY[pix, Cout]^T * X_tap[pix, Cin]     (split-K, fp32 red.add)
// Activations are NHWC bf16, weights are [Cout][R][S][Cin] bf16 (physical layout of a channels_last OIHW tensor).
//
// Replaces what the reference reaches through cuDNN/cuBLAS (reference resnet.py:36-54 convs, :211 fc; SURVEY G1-G3,G10).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (TMEM lane quarter = warp_idx % 4, two warps per quarter splitting the columns).  smem ring of kStages {A 128x64, B BNx64} bf16 tiles; TMEM holds two
// accumulator stages of BN fp32 columns so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "conv_gemm.h"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;                       // reduction elements per smem stage (128 bytes of bf16)
constexpr int kABytes = BM * BK * 2;         // 16 KB
constexpr int kBoxBytes = 64 * BK * 2;       // one 64x64 MN-major box = 8 KB
constexpr int kEpiWarps = 8;                 // two per TMEM lane quarter: latency of one hides behind the other
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = 64 + kEpiThreads;

// CG = CTAs per MMA (tcgen05 cta_group): with CG == 2 a CTA pair (cluster of 2, same TPC) works on a 256 x BN tile --
// each CTA stages its own 128 A rows but only HALF of the B tile, which the pair's single MMA reads from both shared
// memories.  L2->SM bytes per k-block drop from 16 + BN/8 KB to 16 + BN/16 KB per SM, and that ingest rate (~40
// B/clk/SM measured in round 1, see profiles/) is what bounds every compute-heavy layer.
template <int BN, int CG>
struct Cfg {
  static constexpr int kBBytes = (BN / CG) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (CG == 2) ? (BN == 256 ? 5 : 6) : ((BN == 256) ? 3 : (BN == 128 ? 5 : 6));
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages; 128/256/512 are all legal allocations
  static constexpr int kStagingBytes = kEpiWarps * 2 /*double buffer*/ * 4096;  // 32 rows x 128 B each
  static constexpr int kMaxStages = 12;     // resident-weights mode re-partitions the ring into A-only stages
  static constexpr int kRingBytes = 160 * 1024;  // same for every BN (>= kStages * kStageBytes) so that a resident
                                                 // 128 KB weight slab + 2 A stages fits
  static_assert(kStages * kStageBytes <= kRingBytes, "ring too small");
  static constexpr int kSmemBytes = kRingBytes + kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct PixelCoord { int w, h, n; };

__device__ __forceinline__ PixelCoord decode_pixel(const ConvGemmParams& p, int pix) {
  // linear output-pixel index -> base (w, h, n) coordinate of the im2col traversal
  int pq = p.im_P * p.im_Q;
  int n = pix / pq;
  int rem = pix - n * pq;
  int ph = rem / p.im_Q;
  int q = rem - ph * p.im_Q;
  return {q * p.im_stride + p.im_low_w, ph * p.im_stride + p.im_low_h, n};
}

struct WorkItem { int m0, n0, nb, tap, it_begin, it_end, vb0, nbox, g; };

// out of line on purpose: three warp roles call it once per work item; inlining triples ~150 instructions of
// integer division in an instruction-cache-bound kernel
// m_span = rows covered by one item (BM per CTA of the pair), m_off = this CTA's offset inside it
static __device__ __noinline__ WorkItem decode_item(const ConvGemmParams& p, int item, int bn, int m_span, int m_off) {
  WorkItem w;
  if (p.kind != KIND_WGRAD) {
    int nb, mb;
    if (p.b_resident || !p.m_fastest) {
      // n fastest (resident mode: the grid is a multiple of n_blocks, so a CTA keeps ONE n-block -- its resident weight
      // slab -- for life; streaming mode: the CTAs running at the same time share A tiles and spread over n_blocks B tiles)
      nb = item % p.n_blocks;
      const int rest = item / p.n_blocks;
      mb = rest % p.m_blocks;
      w.g = rest / p.m_blocks;                     // conv group (0 when groups == 1)
    } else {
      // m fastest: a CTA's consecutive items share the n-block (BN statistics flushed a handful of times per launch), but
      // every CTA then pulls the SAME weight tile from L2 at the same time -- measured slower (profiles/r2), kept as an
      // experiment knob (B200_CONV_ORDER=m)
      mb = item % p.m_blocks;
      const int rest = item / p.m_blocks;
      nb = rest % p.n_blocks;
      w.g = rest / p.n_blocks;
    }
    w.m0 = mb * m_span + m_off; w.n0 = nb * bn; w.nb = w.g * p.n_blocks + nb; w.tap = 0; w.vb0 = 0; w.nbox = 0;
    w.it_begin = 0; w.it_end = p.taps * p.kb_per_tap;
  } else {
    // wgrad: an item owns `nbox` consecutive virtual B boxes (tap, 64-channel slice of Cin) -> N = 64*nbox columns
    int split = item % p.splits; int rest = item / p.splits;
    int group = rest % p.n_blocks; rest /= p.n_blocks;
    int mb = rest % p.m_blocks;
    w.g = rest / p.m_blocks;
    w.m0 = mb * m_span + m_off; w.n0 = 0; w.nb = group; w.tap = 0;
    w.vb0 = group * p.vb_per_item;
    w.nbox = min(p.vb_per_item, p.vboxes_total - w.vb0);
    w.it_begin = (int)(((long long)p.k_blocks_total * split) / p.splits);
    w.it_end = (int)(((long long)p.k_blocks_total * (split + 1)) / p.splits);
  }
  return w;
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_add,
                 const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * kABytes;  // (re-pointed below in resident mode)
  uint8_t* smem_stage_out = smem + C::kRingBytes;  // epilogue staging (1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage_out + C::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kMaxStages] TMA -> MMA
  uint64_t* empty_bar = bars + C::kMaxStages;      // [kMaxStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * C::kMaxStages;  // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;            // [2] epilogue -> MMA
  uint64_t* add_bar = tmem_empty + 2;              // [kEpiWarps] addend-tile loads (epilogue only)
  uint64_t* bres_bar = add_bar + kEpiWarps;        // [1] resident weight slab landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);
  // Resident-weights mode (fprop/dgrad with a small weight slab): the CTA's n-block of the weights for ALL k
  // iterations is loaded once to the front of the ring memory; the rest becomes a deeper ring of A-only stages.
  const bool resident = p.b_resident != 0;
  const int n_stages = resident ? p.res_stages : C::kStages;
  const int k_iters_fd = p.taps * p.kb_per_tap;
  if (resident) { smem_a = smem + k_iters_fd * C::kBBytes; smem_b = smem; }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair bookkeeping: both CTAs walk the same item sequence; rank 0 (the leader) issues the MMAs
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int item0 = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_span = BM * CG, m_off = (int)cta_rank * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (EPI != EPI_F32_RED) tma_prefetch_desc(&map_out);
    for (int i = 0; i < C::kMaxStages; ++i) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
    mbar_init(smem_u32(bres_bar), 1);
    // tmem_empty of the leader collects one arrival per epilogue WARP of BOTH CTAs (the peer's arrive remotely; 256
    // per-thread remote arrivals per tile cost ~4k cycles on the short-K layers, measured on 256->1024 @14^2)
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tmem_full[i]), 1); mbar_init(smem_u32(&tmem_empty[i]), kEpiWarps * CG); }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(smem_u32(&add_bar[i]), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_2cta(smem_u32(tmem_ptr), C::kTmemCols); tmem_relinquish_2cta(); }
    else { tmem_alloc(smem_u32(tmem_ptr), C::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncwarp();                      // (lane 0 of warp 0 initialised the barriers in a divergent branch)
  if (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (resident && (int)blockIdx.x < p.total_items) {
        // every item of this CTA has the same n-block (grid is a multiple of n_blocks, groups == 1); CG == 1 only
        const WorkItem w0 = decode_item(p, blockIdx.x, BN, BM, 0);
        const uint32_t bb = smem_u32(bres_bar);
        mbar_expect_tx(bb, (uint32_t)(k_iters_fd * C::kBBytes));
        for (int it = 0; it < k_iters_fd; ++it) {
          const int tap = it / p.kb_per_tap, kb = it - tap * p.kb_per_tap;
          const int btap = p.tap_lut_on ? (int)p.tap_lut[tap] : (p.b_flip_taps ? (p.taps - 1 - tap) : tap);
          const uint32_t dst = smem_u32(smem_b + it * C::kBBytes);
          if (p.kind == KIND_FPROP) {
            tma_load_4d(dst, &map_b, bb, kb * BK, btap, w0.n0, 0);
          } else {
            for (int j = 0; j < p.b_nbox; ++j) tma_load_4d(dst + j * kBoxBytes, &map_b, bb, w0.n0 + 64 * j, btap, kb * BK, 0);
          }
        }
      }
      // All per-iteration address arithmetic below is incremental: this single thread feeds the whole CTA, and an
      // ncu source-level capture of the 3x3 wgrad showed it spending ~900 instructions per stage (eleven integer
      // divisions in decode_pixel / tap / box decoding), i.e. the producer -- not TMA or the tensor pipe -- set the
      // pace.  Divisions now happen once per work item; the reduction-pixel coordinate advances by BK per stage.
      const int step_p = (p.b_im2col && p.kind == KIND_WGRAD) ? BK / p.im_Q : 0;   // BK pixels = step_p rows + step_q
      const int step_q = (p.b_im2col && p.kind == KIND_WGRAD) ? BK - step_p * p.im_Q : 0;
      // CG == 2: TMA completions of both CTAs are counted by the LEADER's full barrier (same offset, peer bit cleared)
      const int n_half = (CG == 2) ? (int)cta_rank * (BN / 2) : 0;   // this CTA's half of the B tile (columns of the output)
      for (int item = item0; item < p.total_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, BN, m_span, m_off);
        if (p.kind != KIND_WGRAD) {
          PixelCoord pa{0, 0, 0};
          if (p.a_im2col) pa = decode_pixel(p, w.m0);
          const int ca0 = w.g * p.a_cg;                      // A channel coordinate of this conv group
          int tap = 0, kb = 0, r = 0, sx = 0;               // it_begin == 0 for fprop / dgrad
          for (int it = w.it_begin; it < w.it_end; ++it) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t bar_local = smem_u32(&full_bar[stage]);
            const uint32_t bar = (CG == 2) ? (bar_local & kPeerBitMask) : bar_local;
            const uint32_t dst_a = smem_u32(smem_a + stage * kABytes);
            const uint32_t dst_b = smem_u32(smem_b + stage * C::kBBytes);
            if (CG == 1) mbar_expect_tx(bar, resident ? (uint32_t)kABytes : (uint32_t)C::kStageBytes);
            else if (leader) mbar_expect_tx(bar_local, 2u * (uint32_t)C::kStageBytes);
            const int ca = ca0 + kb * BK;                    // group slice + K block
            if (CG == 2) {
              if (p.a_im2col)
                tma_load_im2col_4d_2cta(dst_a, &map_a, bar, ca, pa.w, pa.h, pa.n, (uint16_t)(sx * p.dil), (uint16_t)(r * p.dil));
              else
                tma_load_3d_2cta(dst_a, &map_a, bar, ca, 0, w.m0);
            } else if (p.a_im2col)
              tma_load_im2col_4d(dst_a, &map_a, bar, ca, pa.w, pa.h, pa.n, (uint16_t)(sx * p.dil), (uint16_t)(r * p.dil));
            else
              tma_load_3d(dst_a, &map_a, bar, ca, 0, w.m0);
            const int btap = p.tap_lut_on ? (int)p.tap_lut[tap] : (p.b_flip_taps ? (p.taps - 1 - tap) : tap);
            // weights are mapped as (Cin/g, taps, Cout/g, groups): anything past a group's extent is zero-filled
            if (resident) {
              // B slab already in shared memory
            } else if (CG == 2) {                                             // this CTA's half of the weight tile
              if (p.kind == KIND_FPROP) {
                tma_load_4d_2cta(dst_b, &map_b, bar, kb * BK, btap, w.n0 + n_half, w.g);
              } else {
                for (int j = 0; j < p.b_nbox; ++j)
                  tma_load_4d_2cta(dst_b + j * kBoxBytes, &map_b, bar, w.n0 + n_half + 64 * j, btap, kb * BK, w.g);
              }
            } else if (p.kind == KIND_FPROP) {
              tma_load_4d(dst_b, &map_b, bar, kb * BK, btap, w.n0, w.g);      // K-major weights [N][tap][K]
            } else {
              for (int j = 0; j < p.b_nbox; ++j)                              // MN-major weights [K][tap][N]
                tma_load_4d(dst_b + j * kBoxBytes, &map_b, bar, w.n0 + 64 * j, btap, kb * BK, w.g);
            }
            if (++kb == p.kb_per_tap) { kb = 0; ++tap; if (++sx == p.S) { sx = 0; ++r; } }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        } else {
          // wgrad: the item's <= 4 virtual B boxes (tap, 64-channel slice of Cin) are fixed; decode them once
          int box_c[4], box_ow[4], box_oh[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int vb = w.vb0 + (j < w.nbox ? j : 0);
            const int tap = vb / p.cin_boxes, cbox = vb - tap * p.cin_boxes;
            const int r = tap / p.S, sx = tap - r * p.S;
            box_c[j] = w.g * p.a_cg + cbox * 64; box_ow[j] = sx * p.dil; box_oh[j] = r * p.dil;
          }
          const int a_c0 = w.g * p.out_cg + w.m0;
          // reduction-pixel coordinate (q, ph, n) of the item's first block; advanced by BK pixels per stage
          int q = 0, ph = 0, n = 0;
          if (p.b_im2col) {
            const int pix = w.it_begin * BK, pq = p.im_P * p.im_Q;
            n = pix / pq;
            const int rem = pix - n * pq;
            ph = rem / p.im_Q;
            q = rem - ph * p.im_Q;
          }
          for (int it = w.it_begin; it < w.it_end; ++it) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t bar_local = smem_u32(&full_bar[stage]);
            const uint32_t bar = (CG == 2) ? (bar_local & kPeerBitMask) : bar_local;
            const uint32_t dst_a = smem_u32(smem_a + stage * kABytes);
            const uint32_t dst_b = smem_u32(smem_b + stage * C::kBBytes);
            // CG == 2: each CTA stages its own dY^T tile (different Cout rows) and HALF of the item's X boxes
            const int nb_mine = (CG == 2) ? (w.nbox >> 1) : w.nbox;
            const int jb0 = (CG == 2) ? (int)cta_rank * nb_mine : 0;
            if (CG == 1) mbar_expect_tx(bar, (uint32_t)(kABytes + w.nbox * kBoxBytes));
            else if (leader) mbar_expect_tx(bar_local, 2u * (uint32_t)(kABytes + nb_mine * kBoxBytes));
            const int k0 = it * BK;  // first reduction pixel of this block
            for (int j = 0; j < p.a_nbox; ++j) {
              if (CG == 2) tma_load_3d_2cta(dst_a + j * kBoxBytes, &map_a, bar, a_c0 + 64 * j, 0, k0);
              else tma_load_3d(dst_a + j * kBoxBytes, &map_a, bar, a_c0 + 64 * j, 0, k0);
            }
            if (p.b_im2col) {
              const int cw = q * p.im_stride + p.im_low_w, chh = ph * p.im_stride + p.im_low_h;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j >= jb0 && j < jb0 + nb_mine) {
                  if (CG == 2) tma_load_im2col_4d_2cta(dst_b + (j - jb0) * kBoxBytes, &map_b, bar, box_c[j], cw, chh, n, (uint16_t)box_ow[j], (uint16_t)box_oh[j]);
                  else tma_load_im2col_4d(dst_b + j * kBoxBytes, &map_b, bar, box_c[j], cw, chh, n, (uint16_t)box_ow[j], (uint16_t)box_oh[j]);
                }
              }
              q += step_q; ph += step_p;
              if (q >= p.im_Q) { q -= p.im_Q; ++ph; }
              while (ph >= p.im_P) { ph -= p.im_P; ++n; }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j >= jb0 && j < jb0 + nb_mine) {
                  if (CG == 2) tma_load_3d_2cta(dst_b + (j - jb0) * kBoxBytes, &map_b, bar, box_c[j], 0, k0);
                  else tma_load_3d(dst_b + j * kBoxBytes, &map_b, bar, box_c[j], 0, k0);
                }
              }
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && (CG == 1 || leader)) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (resident && (int)blockIdx.x < p.total_items) { mbar_wait(smem_u32(bres_bar), 0); tc_fence_after(); }
      for (int item = item0; item < p.total_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, BN, m_span, m_off);
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const uint32_t idesc = (p.kind == KIND_WGRAD) ? ((p.idesc & ~(0x3fu << 17)) | ((uint32_t)(w.nbox * 64 >> 3) << 17)) : p.idesc;
        for (int it = w.it_begin; it < w.it_end; ++it) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);          // TMA bytes have landed
          tc_fence_after();
          const uint64_t a_desc = p.a_desc_hi | (uint64_t)((smem_u32(smem_a + stage * kABytes) >> 4) & 0x3fff);
          const uint8_t* b_tile = resident ? (smem_b + (it - w.it_begin) * C::kBBytes) : (smem_b + stage * C::kBBytes);
          const uint64_t b_desc = p.b_desc_hi | (uint64_t)((smem_u32(b_tile) >> 4) & 0x3fff);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (CG == 2)
              umma_bf16_2cta(tmem_d, a_desc + (uint64_t)(k * p.a_kstep16), b_desc + (uint64_t)(k * p.b_kstep16), idesc,
                             (it > w.it_begin || k > 0) ? 1u : 0u);
            else
              umma_bf16(tmem_d, a_desc + (uint64_t)(k * p.a_kstep16), b_desc + (uint64_t)(k * p.b_kstep16), idesc,
                        (it > w.it_begin || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when the MMAs retire
          if (CG == 2) umma_commit_2cta(smem_u32(&empty_bar[stage])); else umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) umma_commit_2cta(smem_u32(&tmem_full[acc])); else umma_commit(smem_u32(&tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;

    if constexpr (EPI != EPI_F32_RED) {
      constexpr bool kBias = (EPI == EPI_BF16_BIAS);
      constexpr bool kAdd = (EPI == EPI_BF16_ADD);
      // TMEM -> registers -> bf16 -> 128B-swizzled smem (32 rows x 64 cols) -> TMA store (coalesced, clipped at the
      // M/N edges).  The BN statistics are column sums read back from that same smem tile: lane l owns columns
      // 2l, 2l+1 of the 64-column chunk (conflict-free LDS.32).  Two warps serve each TMEM lane quarter: with
      // >= 2 chunks per tile they take alternate chunks; with a single chunk (BN = 64) they split its columns and
      // its statistics rows and meet on a named barrier.
      constexpr int NCH = BN / 64;
      const int half = (warp - 2) >> 2;
      constexpr int NMINE = (NCH == 1) ? 1 : NCH / 2;   // 64-column chunks this warp owns per tile
      float st_sum[NMINE][2], st_sq[NMINE][2];
#pragma unroll
      for (int c = 0; c < NMINE; ++c) { st_sum[c][0] = st_sum[c][1] = 0.f; st_sq[c][0] = st_sq[c][1] = 0.f; }
      int st_nb = -1;
      const bool want_stats = (p.stats != nullptr);
      int buf = 0;
      // staging: own double buffer per warp; in the cooperative (NCH == 1) case both warps of a quarter use half 0's
      uint8_t* my_stage = smem_stage_out + ((NCH == 1 ? 0 : half) * 4 + quarter) * 8192;
      const uint32_t pair_bar = 1 + quarter;      // named barrier id shared by the two warps of a quarter
      // optional addend (e.g. the residual-branch gradient in a dgrad): its tile is TMA-loaded into the staging
      // buffer first and summed in registers, which removes a separate elementwise add pass
      constexpr bool has_add = kAdd;
      const uint32_t my_add_bar = smem_u32(&add_bar[(NCH == 1 ? 0 : half) * 4 + quarter]);
      uint32_t add_phase = 0;

      auto flush_stats = [&](int nb) {
        if (nb < 0) return;
#pragma unroll
        for (int ci = 0; ci < NMINE; ++ci) {
          const int c = (NCH == 1) ? 0 : (2 * ci + half);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // nb enumerates (group, n-block); p.N is the per-group column count, stats are [2][groups * N]
            const int g = nb / p.n_blocks, col = (nb % p.n_blocks) * BN + c * 64 + 2 * lane + h;
            if (col < p.N) {
              atomicAdd(p.stats + g * p.N + col, st_sum[ci][h]);
              atomicAdd(p.stats + p.groups * p.N + g * p.N + col, st_sq[ci][h]);
            }
            st_sum[ci][h] = 0.f; st_sq[ci][h] = 0.f;
          }
        }
      };
      // 32 accumulator columns -> four swizzled 16-byte pieces (g0..g0+3) of this thread's staging row
      int bias_off = 0;                              // group offset into the bias vector (set per item)
      auto stage_32cols = [&](uint8_t* sbuf, uint32_t taddr, int col_first, int g0) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if constexpr (kBias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col_first + j;
            const float b = (col < p.N) ? __ldg(p.bias + bias_off + col) : 0.f;
            v[j] = __float_as_uint(__uint_as_float(v[j]) + b);
          }
        }
        if constexpr (kAdd) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 ad = *reinterpret_cast<const uint4*>(sbuf + lane * 128 + (((g0 + g) ^ (lane & 7)) << 4));
            const uint32_t aw[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              v[8 * g + 2 * t] = __float_as_uint(__uint_as_float(v[8 * g + 2 * t]) + __uint_as_float(aw[t] << 16));
              v[8 * g + 2 * t + 1] = __float_as_uint(__uint_as_float(v[8 * g + 2 * t + 1]) + __uint_as_float(aw[t] & 0xffff0000u));
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
          *reinterpret_cast<uint4*>(sbuf + lane * 128 + (((g0 + g) ^ (lane & 7)) << 4)) = pk;
        }
      };
      auto column_stats = [&](const uint8_t* sbuf, int r0, int r1, int rows_valid, int ci) {
        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int r = r0; r < r1; ++r) {
          const uint32_t wd = *reinterpret_cast<const uint32_t*>(sbuf + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
          const float f0 = (r < rows_valid) ? __uint_as_float(wd << 16) : 0.f;
          const float f1 = (r < rows_valid) ? __uint_as_float(wd & 0xffff0000u) : 0.f;
          a0 += f0; q0 = fmaf(f0, f0, q0);
          a1 += f1; q1 = fmaf(f1, f1, q1);
        }
        st_sum[ci][0] += a0; st_sum[ci][1] += a1; st_sq[ci][0] += q0; st_sq[ci][1] += q1;
      };

      for (int item = item0; item < p.total_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, BN, m_span, m_off);
        if (want_stats && w.nb != st_nb) { flush_stats(st_nb); st_nb = w.nb; }
        mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
        tc_fence_after();
        const int row_base = w.m0 + quarter * 32;
        const int rows_valid = min(32, p.M - row_base);
        bias_off = w.g * p.N;
        const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
        if constexpr (NCH == 1) {
          uint8_t* sbuf = my_stage + buf * 4096;
          if (half == 0 && lane == 0) {
            bulk_wait_group_read<1>();
            if (has_add) {
              mbar_expect_tx(my_add_bar, 4096);
              tma_load_3d(smem_u32(sbuf), &map_add, my_add_bar, w.n0, row_base, w.g);
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // buffer is free, both warps present
          if (has_add) { mbar_wait(my_add_bar, add_phase); add_phase ^= 1; }
          stage_32cols(sbuf, tacc + half * 32, w.n0 + half * 32, half * 4);
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // whole 32x64 tile staged
          if (half == 0 && lane == 0) {
            tma_store_3d(&map_out, smem_u32(sbuf), w.n0, row_base, w.g);
            bulk_commit_group();
          }
          if (want_stats) column_stats(sbuf, half * 16, half * 16 + 16, rows_valid, 0);
          buf ^= 1;
        } else {
#pragma unroll
          for (int ci = 0; ci < NMINE; ++ci) {
            const int c = 2 * ci + half;                  // the sibling warp of this quarter takes the other chunks
            const int col0 = w.n0 + c * 64;
            if (col0 >= p.N) continue;                   // chunk entirely past the N edge
            uint8_t* sbuf = my_stage + buf * 4096;
            if (lane == 0) {
              bulk_wait_group_read<1>();                  // the TMA store that last read this buffer has drained
              if (has_add) {
                mbar_expect_tx(my_add_bar, 4096);
                tma_load_3d(smem_u32(sbuf), &map_add, my_add_bar, col0, row_base, w.g);
              }
            }
            __syncwarp();
            if (has_add) { mbar_wait(my_add_bar, add_phase); add_phase ^= 1; }
            stage_32cols(sbuf, tacc + c * 64, col0, 0);
            stage_32cols(sbuf, tacc + c * 64 + 32, col0 + 32, 4);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_out, smem_u32(sbuf), col0, row_base, w.g);
              bulk_commit_group();
            }
            if (want_stats) column_stats(sbuf, 0, 32, rows_valid, ci);
            buf ^= 1;
          }
        }
        tc_fence_before();
        __syncwarp();                      // every lane has drained its TMEM loads of this accumulator stage
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(smem_u32(&tmem_empty[acc]) & kPeerBitMask); else mbar_arrive(smem_u32(&tmem_empty[acc]));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (want_stats) { flush_stats(st_nb); if (p.peer.world > 1) __threadfence(); }   // ordered before the tail counter
      if (lane == 0) bulk_wait_group<0>();            // all stores complete before the CTA may exit
      __syncwarp();
    } else {
      // EPI_F32_RED: split-K partial sums into fp32 dW[M][taps][N]
      for (int item = item0; item < p.total_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, BN, m_span, m_off);
        mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
        tc_fence_after();
        const int row = w.m0 + row_in_tile;
        const bool row_ok = row < p.M;
        const bool has_k = w.it_end > w.it_begin;
        const int half = (warp - 2) >> 2;            // the two warps of a lane quarter take alternate 32-col chunks
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if ((c & 1) != half) continue;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, v);
          tmem_ld_wait();
          const int j = c >> 1;                         // which B box of this item
          if (j >= w.nbox) continue;
          const int vb = w.vb0 + j;
          const int tap = vb / p.cin_boxes, cbox = vb - tap * p.cin_boxes;
          const int col0 = cbox * 64 + (c & 1) * 32;    // channel offset inside the tap
          if (row_ok && col0 < p.N && has_k) {
            float* dst = reinterpret_cast<float*>(p.out) + (long long)(w.g * p.out_cg + row) * p.ldo + (long long)tap * p.tap_stride + col0;
            if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g),
                             "f"(__uint_as_float(v[4 * g + 0])), "f"(__uint_as_float(v[4 * g + 1])),
                             "f"(__uint_as_float(v[4 * g + 2])), "f"(__uint_as_float(v[4 * g + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (col0 + jj < p.N) atomicAdd(dst + jj, __uint_as_float(v[jj]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();                      // every lane has drained its TMEM loads of this accumulator stage
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(smem_u32(&tmem_empty[acc]) & kPeerBitMask); else mbar_arrive(smem_u32(&tmem_empty[acc]));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  __syncwarp();                      // the single-lane producer / MMA loops rejoin their warps before the aligned barrier
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer can still signal this CTA's barriers
  else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_2cta(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
  // SyncBN forward: this launch produced the layer's local statistics; the last CTA out raises this rank's flag on every
  // peer, so the flag travels during the kernel boundary instead of inside bn_apply (elementwise.cu: peer_exchange_reduce)
  if (EPI != EPI_F32_RED && p.peer.world > 1 && threadIdx.x == 0) {
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

template <int BN, int EPI, int CG>
static cudaError_t launch_one(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& md,
                              const ConvGemmParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN, CG>;
  auto kern = conv_gemm_kernel<BN, EPI, CG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (CG == 1) {
    kern<<<grid, kNumThreads, C::kSmemBytes, stream>>>(ma, mb, mo, md, p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kNumThreads); cfg.dynamicSmemBytes = C::kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, md, p);
}

}  // namespace b200

extern "C" int b200_conv_gemm_launch(const CUtensorMap* map_a, const CUtensorMap* map_b, const CUtensorMap* map_out,
                                     const CUtensorMap* map_add, const ConvGemmParams* p, int bn, int grid,
                                     cudaStream_t stream) {
  using namespace b200;
  cudaError_t e = cudaErrorInvalidValue;
  const int cg = p->cta_group == 2 ? 2 : 1;
#define B200_DISPATCH_BN(EPI_)                                                                             \
  do {                                                                                                    \
    if (cg == 1) {                                                                                        \
      if (bn == 64) e = launch_one<64, EPI_, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream);     \
      else if (bn == 128) e = launch_one<128, EPI_, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream); \
      else if (bn == 256) e = launch_one<256, EPI_, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream); \
    } else {                                                                                              \
      if (bn == 128) e = launch_one<128, EPI_, 2>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream);   \
      else if (bn == 256) e = launch_one<256, EPI_, 2>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream); \
    }                                                                                                     \
  } while (0)
  switch (p->epi) {
    case EPI_BF16: B200_DISPATCH_BN(EPI_BF16); break;
    case EPI_BF16_BIAS:
      if (cg == 1) {
        if (bn == 64) e = launch_one<64, EPI_BF16_BIAS, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream);
        else if (bn == 128) e = launch_one<128, EPI_BF16_BIAS, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream);
        else if (bn == 256) e = launch_one<256, EPI_BF16_BIAS, 1>(*map_a, *map_b, *map_out, *map_add, *p, grid, stream);
      }
      break;
    case EPI_BF16_ADD: B200_DISPATCH_BN(EPI_BF16_ADD); break;
    case EPI_F32_RED: B200_DISPATCH_BN(EPI_F32_RED); break;
    default: break;
  }
#undef B200_DISPATCH_BN
  return (int)e;
}
