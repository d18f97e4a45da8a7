// Fused multi-head self-attention with 2-D relative-position logits on tcgen05 (BoTNet's MHSA: 14x14 = 196 tokens,
// 128-wide heads; reference botnet.py:193-215 with RelPosEmb 77-98 / rel_to_abs 25-57).
//
// Activations stay in the NHWC layout the 1x1 projections produce: qk [B,196,2*heads*128] (q heads first, then k
// heads), v / out [B,196,heads*128].  No transposes, no per-head copies: every operand is a TMA box of those tensors.
//
//   forward  (CTA = one (128-token tile, head, sample)):
//     S      = Q_t . [K ; rel_w ; rel_h]^T        two MMAs (N = 256 tokens, N = 64 rel rows) -> TMEM (320 fp32 columns)
//     logits = scale * (S[j] + S_rel_w[x_j - x_i + 13] + S_rel_h[y_j - y_i + 13])   -- the "relative -> absolute" shuffle of
//              the reference is an index into the 27 + 27 relative logits of the row, done in registers
//     P      = softmax(logits)  (thread = query row, online max/sum over TMEM chunks) -> bf16, K-major swizzled smem tile
//              (+ saved to global for the backward pass)
//     O      = P . V                              MMA with V as an MN-major operand -> TMEM -> bf16 -> TMA store
//   backward, part 1 (same grid):   dP = dO_t . V^T ; dS = scale * P o (dP - rowsum(P o dP)) ; relative columns of dS
//     are the per-row sums over equal x_j / y_j ; dQ_t = dS' . [K ; rel_w ; rel_h]
//   backward, part 2 (CTA = (128-key tile, head, sample, dV|dK)):  dV = P^T . dO,  dK = dS^T . Q  (MN-major A operands
//     straight from the saved P / dS); d(rel tables) = dS_rel^T . Q over all samples and heads is a grouped wgrad GEMM
//     of the conv kernel followed by rel_grad_reduce.
// Every CTA is single-shot (one set of TMA loads, 2-3 MMAs chains, one epilogue), so synchronisation is a handful of
// one-use mbarriers instead of rings.
#include <algorithm>

#include "attention.h"
#include "common.cuh"

namespace b200 {

namespace attn {
constexpr int S = 196, W = 14, D = 128;
constexpr int PW = 208;                       // row pitch of the saved probabilities (196 rounded up to 16)
constexpr int DSW = 256;                      // row pitch of the saved dS
constexpr int kThreads = 160;                 // warps 0-3: one thread per tile row (TMEM lane), warp 4: TMA + MMA
constexpr uint32_t KB16 = 16384, BOX = 8192;
// smem matrix descriptors (see bindings.cpp): K-major rows of 128 B, 8-row groups 1024 B apart, UMMA_K step = 32 B;
// MN-major [64 k][64 mn] boxes of 8 KB, 8-k groups 1024 B apart, UMMA_K step = 2048 B
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return make_smem_desc_hi_sw128(16, 1024) | (uint64_t)((saddr >> 4) & 0x3fff); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) { return make_smem_desc_hi_sw128(8192, 1024) | (uint64_t)((saddr >> 4) & 0x3fff); }
constexpr uint32_t kStepK = 2, kStepMN = 128;

__device__ __forceinline__ uint32_t ld_tmem_addr(uint32_t base, int warp, int col) { return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col; }
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void unpack8(uint4 u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
// 16-byte chunk `g` (8 bf16 columns) of row `r` in a K-major 128B-swizzled tile whose rows are 128 B
__device__ __forceinline__ uint32_t sw_off(int r, int g) { return (uint32_t)(r * 128 + ((g ^ (r & 7)) << 4)); }

// [128 rows][ncols fp32] accumulator at TMEM column `col0` -> bf16 -> swizzled staging -> TMA store (box 64 cols x 32 rows)
__device__ __forceinline__ void store_rows_bf16(uint32_t tmem_base, int col0, int warp, int lane, uint8_t* stage_warp,
                                                const void* map, int c_first, int row_first, int z) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint8_t* sbuf = stage_warp + c * 4096;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(ld_tmem_addr(tmem_base, warp, col0 + c * 64 + hlf * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[8 * g + t]);
        *reinterpret_cast<uint4*>(sbuf + sw_off(lane, hlf * 4 + g)) = pack8(f);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(map, smem_u32(sbuf), c_first + c * 64, row_first, z);
      bulk_commit_group();
    }
  }
  if (lane == 0) bulk_wait_group<0>();
  __syncwarp();
}
}  // namespace attn

using namespace attn;

// ======================================================================================================= forward
// smem: [0,32K) Q_t | [32K,96K) K tokens (K-major B, 256 rows) | [96K,112K) rel rows (K-major B, 64 rows) |
//       [112K,176K) V (MN-major B, 4 token blocks x 2 boxes) | [176K,~209K) per-row relative logits (fp32, pitch 65)
//       P (64 KB) aliases [0,64K) once S is complete, the output staging (32 KB) aliases [64K,96K)
constexpr int kFwdSmem = 176 * 1024 + 128 * 65 * 4 + 1024 + 256;
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap qk128, const __grid_constant__ CUtensorMap v64,
                const __grid_constant__ CUtensorMap relw, const __grid_constant__ CUtensorMap relh,
                const __grid_constant__ CUtensorMap out32, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_k = smem + 32 * 1024;
  uint8_t* s_rel = smem + 96 * 1024;
  uint8_t* s_v = smem + 112 * 1024;
  float* s_bias = reinterpret_cast<float*>(smem + 176 * 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 176 * 1024 + 128 * 65 * 4);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  uint8_t* s_p = smem;                   // alias
  uint8_t* s_stage = smem + 64 * 1024;   // alias
  const uint32_t bar_load = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_p = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int cq = h * D, ck = p.heads * D + h * D;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&qk128); tma_prefetch_desc(&v64); tma_prefetch_desc(&out32);
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 4) { __syncwarp(); tmem_alloc(smem_u32(tmem_ptr), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  constexpr int kColRel = 256, kColO = 320;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 176u * 1024u);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(smem_u32(s_q + kb * KB16), &qk128, bar_load, cq + kb * 64, i0, b);
        tma_load_3d(smem_u32(s_k + kb * 2 * KB16), &qk128, bar_load, ck + kb * 64, 0, b);
        tma_load_3d(smem_u32(s_k + kb * 2 * KB16 + KB16), &qk128, bar_load, ck + kb * 64, 128, b);
        tma_load_3d(smem_u32(s_rel + kb * BOX), &relw, bar_load, kb * 64, 0, 0);
        tma_load_3d(smem_u32(s_rel + kb * BOX + 4096), &relh, bar_load, kb * 64, 0, 0);
      }
      for (int tb = 0; tb < 4; ++tb)
        for (int nb = 0; nb < 2; ++nb)
          tma_load_3d(smem_u32(s_v + (tb * 2 + nb) * BOX), &v64, bar_load, h * D + nb * 64, tb * 64, b);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t id_tok = make_idesc_bf16(128, 256, 0, 0), id_rel = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t a = desc_k(smem_u32(s_q + kb * KB16)) + (uint64_t)(k * kStepK);
          umma_bf16(tmem, a, desc_k(smem_u32(s_k + kb * 2 * KB16)) + (uint64_t)(k * kStepK), id_tok, (kb | k) ? 1u : 0u);
          umma_bf16(tmem + kColRel, a, desc_k(smem_u32(s_rel + kb * BOX)) + (uint64_t)(k * kStepK), id_rel, (kb | k) ? 1u : 0u);
        }
      umma_commit(bar_s);
      // P is written by the softmax warps; then O = P . V
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t id_pv = make_idesc_bf16(128, 128, 0, 1);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + kColO, desc_k(smem_u32(s_p + kb * KB16)) + (uint64_t)(k * kStepK),
                    desc_mn(smem_u32(s_v + kb * 2 * BOX)) + (uint64_t)(k * kStepMN), id_pv, (kb | k) ? 1u : 0u);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;            // tile row == TMEM lane
    const int t = i0 + row;
    const bool valid = t < S;
    const int tt = valid ? t : S - 1;
    const int yi = tt / W, xi = tt - yi * W;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // relative logits of this row -> private smem row, then the 14 + 14 values that apply to (x_j, y_j)
    float* mybias = s_bias + row * 65;
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, kColRel), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mybias[j] = __uint_as_float(v[j]);
      tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, kColRel + 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mybias[32 + j] = __uint_as_float(v[j]);
    }
    const float c2 = p.scale * 1.4426950408889634f;   // logits in log2 units
    float bw[W], bh[W];
#pragma unroll
    for (int x = 0; x < W; ++x) { bw[x] = mybias[x - xi + 13] * c2; bh[x] = mybias[32 + x - yi + 13] * c2; }
    // pass 1: online max / sum over the 196 keys
    float m = -INFINITY, l = 0.f;
#pragma unroll
    for (int ch = 0; ch < 7; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, ch * 32), v);
      tmem_ld_wait();
      float x[32], cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = ch * 32 + j;
        if (col < S) { x[j] = fmaf(__uint_as_float(v[j]), c2, bw[col % W] + bh[col / W]); cm = fmaxf(cm, x[j]); }
      }
      const float mn = fmaxf(m, cm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) if (ch * 32 + j < S) acc += exp2f(x[j] - mn);
      l = l * exp2f(m - mn) + acc;
      m = mn;
    }
    const float inv_l = 1.f / l;
    // everyone is done reading S before P overwrites Q/K
    // (P aliases the operands of the S MMAs, which have completed: bar_s)
    __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(p.p_save) + ((long long)(b * p.heads + h) * S + tt) * PW;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float pr[32];
      if (ch < 7) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, ch * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = ch * 32 + j;
          pr[j] = (col < S) ? exp2f(fmaf(__uint_as_float(v[j]), c2, bw[col % W] + bh[col / W]) - m) * inv_l : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pr[j] = 0.f;
      }
      uint8_t* kblock = s_p + (ch >> 1) * KB16;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 pk = pack8(pr + 8 * g);
        *reinterpret_cast<uint4*>(kblock + sw_off(row, (ch & 1) * 4 + g)) = pk;
        if (valid && ch * 32 + g * 8 < PW) *reinterpret_cast<uint4*>(prow + ch * 32 + g * 8) = pk;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);
    // epilogue: O -> bf16 -> out[b, t, h*128 + :]
    mbar_wait(bar_o, 0);
    tc_fence_after();
    store_rows_bf16(tmem, kColO, warp, lane, s_stage + warp * 8192, &out32, h * D, i0 + warp * 32, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// ======================================================================================================= backward 1
// smem: [0,32K) dO_t | [32K,96K) V (K-major B, 256 rows) | [96K,176K) [K tokens ; rel] (MN-major B, 5 k-blocks x 2 boxes) |
//       [176K,208K) dQ staging;  dS' (A, 5 k-blocks x 16 KB = 80 KB) aliases [0,80K) once dP is complete
constexpr int kBwd1Smem = 208 * 1024 + 1024 + 256;
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap dout128, const __grid_constant__ CUtensorMap v128,
                   const __grid_constant__ CUtensorMap qk64, const __grid_constant__ CUtensorMap relw,
                   const __grid_constant__ CUtensorMap relh, const __grid_constant__ CUtensorMap dqk32, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_do = smem;
  uint8_t* s_v = smem + 32 * 1024;
  uint8_t* s_kp = smem + 96 * 1024;
  uint8_t* s_stage = smem + 176 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 208 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  uint8_t* s_ds = smem;                  // alias
  const uint32_t bar_load = smem_u32(&bars[0]), bar_dp = smem_u32(&bars[1]), bar_ds = smem_u32(&bars[2]), bar_dq = smem_u32(&bars[3]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int ck = p.heads * D + h * D;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&dout128); tma_prefetch_desc(&v128); tma_prefetch_desc(&qk64); tma_prefetch_desc(&dqk32);
    mbar_init(bar_load, 1); mbar_init(bar_dp, 1); mbar_init(bar_ds, 128); mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 4) { __syncwarp(); tmem_alloc(smem_u32(tmem_ptr), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  constexpr int kColDq = 256;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 176u * 1024u);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_3d(smem_u32(s_do + kb * KB16), &dout128, bar_load, h * D + kb * 64, i0, b);
        tma_load_3d(smem_u32(s_v + kb * 2 * KB16), &v128, bar_load, h * D + kb * 64, 0, b);
        tma_load_3d(smem_u32(s_v + kb * 2 * KB16 + KB16), &v128, bar_load, h * D + kb * 64, 128, b);
      }
      for (int tb = 0; tb < 4; ++tb)
        for (int nb = 0; nb < 2; ++nb)
          tma_load_3d(smem_u32(s_kp + (tb * 2 + nb) * BOX), &qk64, bar_load, ck + nb * 64, tb * 64, b);
      for (int nb = 0; nb < 2; ++nb) {
        tma_load_3d(smem_u32(s_kp + (8 + nb) * BOX), &relw, bar_load, nb * 64, 0, 0);
        tma_load_3d(smem_u32(s_kp + (8 + nb) * BOX + 4096), &relh, bar_load, nb * 64, 0, 0);
      }
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t id_dp = make_idesc_bf16(128, 256, 0, 0);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, desc_k(smem_u32(s_do + kb * KB16)) + (uint64_t)(k * kStepK),
                    desc_k(smem_u32(s_v + kb * 2 * KB16)) + (uint64_t)(k * kStepK), id_dp, (kb | k) ? 1u : 0u);
      umma_commit(bar_dp);
      mbar_wait(bar_ds, 0);
      tc_fence_after();
      const uint32_t id_dq = make_idesc_bf16(128, 128, 0, 1);
#pragma unroll
      for (int kb = 0; kb < 5; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + kColDq, desc_k(smem_u32(s_ds + kb * KB16)) + (uint64_t)(k * kStepK),
                    desc_mn(smem_u32(s_kp + kb * 2 * BOX)) + (uint64_t)(k * kStepMN), id_dq, (kb | k) ? 1u : 0u);
      umma_commit(bar_dq);
    }
    __syncwarp();
  } else {
    const int row = warp * 32 + lane;
    const int t = i0 + row;
    const bool valid = t < S;
    const int tt = valid ? t : S - 1;
    const int yi = tt / W, xi = tt - yi * W;
    const long long bh = (long long)b * p.heads + h;
    const __nv_bfloat16* prow = reinterpret_cast<const __nv_bfloat16*>(p.p_save) + (bh * S + tt) * PW;
    __nv_bfloat16* dsrow = reinterpret_cast<__nv_bfloat16*>(p.ds_save) + (bh * S + tt) * DSW;
    mbar_wait(bar_dp, 0);
    tc_fence_after();
    // pass 1: delta = sum_j P * dP
    float delta = 0.f;
#pragma unroll
    for (int ch = 0; ch < 7; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, ch * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (ch * 32 + g * 8 < PW) {
          float pf[8];
          unpack8(valid ? *reinterpret_cast<const uint4*>(prow + ch * 32 + g * 8) : make_uint4(0, 0, 0, 0), pf);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta = fmaf(pf[e], __uint_as_float(v[8 * g + e]), delta);
        }
      }
    }
    // pass 2: dS = scale * P (dP - delta); relative columns = sums over equal x_j / y_j
    float gw[W], gh[W];
#pragma unroll
    for (int x = 0; x < W; ++x) { gw[x] = 0.f; gh[x] = 0.f; }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ds[j] = 0.f;
      if (ch < 7) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ld_tmem_addr(tmem, warp, ch * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (ch * 32 + g * 8 < PW) {
            float pf[8];
            unpack8(valid ? *reinterpret_cast<const uint4*>(prow + ch * 32 + g * 8) : make_uint4(0, 0, 0, 0), pf);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = ch * 32 + g * 8 + e;
              if (col < S) {
                const float d = p.scale * pf[e] * (__uint_as_float(v[8 * g + e]) - delta);
                ds[8 * g + e] = d;
                gw[col % W] += d;
                gh[col / W] += d;
              }
            }
          }
        }
      }
      uint8_t* kblock = s_ds + (ch >> 1) * KB16;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 pk = pack8(ds + 8 * g);
        *reinterpret_cast<uint4*>(kblock + sw_off(row, (ch & 1) * 4 + g)) = pk;
        if (valid) *reinterpret_cast<uint4*>(dsrow + ch * 32 + g * 8) = pk;
      }
    }
    // relative block (k-block 4 of dS'): zero, then scatter the 14 + 14 sums to their relative indices
    uint8_t* relblock = s_ds + 4 * KB16;
#pragma unroll
    for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(relblock + sw_off(row, g)) = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int x = 0; x < W; ++x) {
      const int cw = x - xi + 13, chh = 32 + x - yi + 13;
      *reinterpret_cast<__nv_bfloat16*>(relblock + sw_off(row, cw >> 3) + (cw & 7) * 2) = __float2bfloat16(valid ? gw[x] : 0.f);
      *reinterpret_cast<__nv_bfloat16*>(relblock + sw_off(row, chh >> 3) + (chh & 7) * 2) = __float2bfloat16(valid ? gh[x] : 0.f);
    }
    if (valid) {
      __nv_bfloat16* rrow = reinterpret_cast<__nv_bfloat16*>(p.dsrel) + ((long long)b * S + t) * (p.heads * 64) + h * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(rrow + g * 8) = *reinterpret_cast<const uint4*>(relblock + sw_off(row, g));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_ds);
    mbar_wait(bar_dq, 0);
    tc_fence_after();
    store_rows_bf16(tmem, kColDq, warp, lane, s_stage + warp * 8192, &dqk32, h * D, i0 + warp * 32, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// ======================================================================================================= backward 2
// dV[j, :] = sum_i P[i, j] dO[i, :]   /   dK[j, :] = sum_i dS[i, j] Q[i, :]      (M = 128 keys, N = 128, K = 256 queries)
// smem: [0,64K) A^T boxes (4 query blocks x 2 key boxes, MN-major) | [64K,128K) B boxes (4 query blocks x 2 d boxes) | [128K,160K) staging
constexpr int kBwd2Smem = 160 * 1024 + 1024 + 256;
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap psave64, const __grid_constant__ CUtensorMap dssave64,
                    const __grid_constant__ CUtensorMap dout64, const __grid_constant__ CUtensorMap qk64,
                    const __grid_constant__ CUtensorMap dv32, const __grid_constant__ CUtensorMap dqk32, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + 64 * 1024;
  uint8_t* s_stage = smem + 128 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_load = smem_u32(&bars[0]), bar_d = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = (blockIdx.x & 1) * 128, which = blockIdx.x >> 1, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + h;
  const CUtensorMap* map_a = which == 0 ? &psave64 : &dssave64;
  const CUtensorMap* map_b = which == 0 ? &dout64 : &qk64;
  const CUtensorMap* map_o = which == 0 ? &dv32 : &dqk32;
  const int c_out = which == 0 ? h * D : p.heads * D + h * D;   // dV -> dv[.., h*128 + :], dK -> dqk[.., heads*128 + h*128 + :]

  if (threadIdx.x == 128) {
    tma_prefetch_desc(map_a); tma_prefetch_desc(map_b); tma_prefetch_desc(map_o);
    mbar_init(bar_load, 1); mbar_init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 4) { __syncwarp(); tmem_alloc(smem_u32(tmem_ptr), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 128u * 1024u);
      for (int kb = 0; kb < 4; ++kb)
        for (int x = 0; x < 2; ++x) {
          tma_load_3d(smem_u32(s_a + (kb * 2 + x) * BOX), map_a, bar_load, j0 + x * 64, kb * 64, bh);
          tma_load_3d(smem_u32(s_b + (kb * 2 + x) * BOX), map_b, bar_load, h * D + x * 64, kb * 64, b);
        }
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t id = make_idesc_bf16(128, 128, 1, 1);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, desc_mn(smem_u32(s_a + kb * 2 * BOX)) + (uint64_t)(k * kStepMN),
                    desc_mn(smem_u32(s_b + kb * 2 * BOX)) + (uint64_t)(k * kStepMN), id, (kb | k) ? 1u : 0u);
      umma_commit(bar_d);
    }
    __syncwarp();
  } else {
    mbar_wait(bar_d, 0);
    tc_fence_after();
    store_rows_bf16(tmem, 0, warp, lane, s_stage + warp * 8192, map_o, c_out, j0 + warp * 32, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

// d(rel_width)[m][:] += sum_heads dw[h*64 + m][:],  d(rel_height)[m][:] += sum_heads dw[h*64 + 32 + m][:]   (m < 27)
__global__ void rel_grad_reduce_kernel(const float* __restrict__ dw, float* __restrict__ gw, float* __restrict__ gh, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 27 * D) return;
  const int m = idx / D, d = idx - m * D;
  float a = 0.f, c = 0.f;
  for (int h = 0; h < heads; ++h) { a += dw[(h * 64 + m) * D + d]; c += dw[(h * 64 + 32 + m) * D + d]; }
  gw[idx] += a;
  gh[idx] += c;
}

template <typename Kern>
static cudaError_t set_smem(Kern kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_fwd(const CUtensorMap* qk128, const CUtensorMap* v64, const CUtensorMap* relw, const CUtensorMap* relh,
                             const CUtensorMap* out32, const AttnParams* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) { cudaError_t e = set_smem(attn_fwd_kernel, kFwdSmem); if (e != cudaSuccess) return (int)e; configured = true; }
  attn_fwd_kernel<<<dim3(2, p->heads, p->B), kThreads, kFwdSmem, s>>>(*qk128, *v64, *relw, *relh, *out32, *p);
  return (int)cudaGetLastError();
}
extern "C" int b200_attn_bwd_dq(const CUtensorMap* dout128, const CUtensorMap* v128, const CUtensorMap* qk64, const CUtensorMap* relw,
                                const CUtensorMap* relh, const CUtensorMap* dqk32, const AttnParams* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) { cudaError_t e = set_smem(attn_bwd_dq_kernel, kBwd1Smem); if (e != cudaSuccess) return (int)e; configured = true; }
  attn_bwd_dq_kernel<<<dim3(2, p->heads, p->B), kThreads, kBwd1Smem, s>>>(*dout128, *v128, *qk64, *relw, *relh, *dqk32, *p);
  return (int)cudaGetLastError();
}
extern "C" int b200_attn_bwd_dkv(const CUtensorMap* psave64, const CUtensorMap* dssave64, const CUtensorMap* dout64, const CUtensorMap* qk64,
                                 const CUtensorMap* dv32, const CUtensorMap* dqk32, const AttnParams* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) { cudaError_t e = set_smem(attn_bwd_dkv_kernel, kBwd2Smem); if (e != cudaSuccess) return (int)e; configured = true; }
  attn_bwd_dkv_kernel<<<dim3(4, p->heads, p->B), kThreads, kBwd2Smem, s>>>(*psave64, *dssave64, *dout64, *qk64, *dv32, *dqk32, *p);
  return (int)cudaGetLastError();
}
extern "C" int b200_rel_grad_reduce(const float* dw, float* grad_w, float* grad_h, int heads, cudaStream_t s) {
  rel_grad_reduce_kernel<<<(27 * 128 + 255) / 256, 256, 0, s>>>(dw, grad_w, grad_h, heads);
  return (int)cudaGetLastError();
}
