// Host/device shared parameter block of the tcgen05 implicit-GEMM kernel (conv_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "elementwise.h"   // PeerCtx

enum ConvGemmKind { KIND_FPROP = 0, KIND_DGRAD = 1, KIND_WGRAD = 2 };
// EPI_BF16: bf16 TMA-store epilogue (+ BN statistics); _BIAS / _ADD add a per-column bias / an addend tile.  They are
// separate instantiations so the common path carries no dead bias/addend code (the epilogue is I-cache sensitive).
enum ConvGemmEpi { EPI_BF16 = 0, EPI_F32_RED = 1, EPI_BF16_BIAS = 2, EPI_BF16_ADD = 3 };

struct ConvGemmParams {
  int kind;            // ConvGemmKind
  int epi;             // ConvGemmEpi
  int M, N;            // GEMM extents used for masking (rows / columns of the output)
  int m_blocks, n_blocks;
  int taps, S;         // R*S filter taps, filter width
  int kb_per_tap;      // fprop/dgrad: ceil(K_channels / 64)
  int dil;
  int a_im2col, a_nbox;
  uint32_t a_kstep16;  // smem-descriptor start-address advance per UMMA_K=16 step (16-byte units)
  uint64_t a_desc_hi;
  int b_im2col, b_nbox;
  uint32_t b_kstep16;
  uint64_t b_desc_hi;
  int b_flip_taps;
  int tap_lut_on;               // dgrad parity classes: weight tap of virtual tap t is tap_lut[t] (overrides b_flip_taps)
  unsigned char tap_lut[28];
  uint32_t idesc;
  int im_P, im_Q, im_stride, im_low_w, im_low_h;  // pixel enumeration of the im2col operand
  int k_blocks_total, splits;                     // wgrad: 64-pixel reduction blocks and split-K factor
  int vb_per_item, cin_boxes, vboxes_total;       // wgrad: B boxes (tap, 64-channel slice) handled by one work item
  int groups, a_cg, out_cg;                       // grouped conv: #groups, A-operand channels per group, output channels per group
  int b_resident;       // fprop/dgrad: the CTA's whole B (weight) slab stays in smem; only A tiles stream through the ring
  int res_stages;       // ring depth (A-only stages) in resident mode
  void* out;
  long long ldo;        // output row pitch in elements
  long long tap_stride; // wgrad: element offset between taps inside one output row
  float* stats;         // optional [2][N] per-column sum / sum of squares (fp32, atomically accumulated)
  const float* bias;    // optional [N]
  long long addend;     // non-zero: add the bf16 tile described by map_add before storing (same geometry as out)
  int total_items;
  PeerCtx peer;         // SyncBN: world > 1 => the last CTA to finish tells the peers that p.stats is final (exchange #epoch)
  int m_fastest;        // fprop/dgrad item order: 0 = n-block fastest, 1 = m-block fastest (ignored in resident mode)
  int cta_group;        // 1, or 2 = CTA pairs (256-row items, B operand split across the pair); see Cfg in conv_gemm.cu
};

// map_out: 2-D tiled map over the bf16 output [M][N] (box 64 cols x 32 rows, 128B swizzle) for the TMA-store
// epilogue; ignored (pass any valid map) for EPI_F32_RED.
extern "C" int b200_conv_gemm_launch(const CUtensorMap* map_a, const CUtensorMap* map_b, const CUtensorMap* map_out,
                                     const CUtensorMap* map_add, const ConvGemmParams* p, int bn, int grid,
                                     cudaStream_t stream);
