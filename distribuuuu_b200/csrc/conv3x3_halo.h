// Parameter block of the halo-reuse 3x3 convolution kernel (conv3x3_halo.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "elementwise.h"   // PeerCtx

struct Conv3x3HaloParams {
  int N, H, W;          // activation [N,H,W,64] -> output [N,H,W,64]
  int rpt;              // image rows per tile = floor(128 / (W+2)): a tile is `rpt` full padded rows (<= 128 positions)
  int tiles_per_img;    // ceil(H / rpt)
  int tiles;            // N * tiles_per_img
  int halo_rows;        // padded positions of one halo load: (rpt + 2) (W+2)  (<= 256)
  int dgrad;            // 0: forward (weights K-major); 1: data gradient (same weight bytes read MN-major, taps flipped)
  void* y;              // output [N,H,W,64] bf16 (rows are written with per-row predicates; map_y is unused)
  float* stats;         // optional [2][64] BN statistics of the output (valid positions only)
  PeerCtx peer;         // SyncBN: world > 1 => the last CTA announces the statistics exchange
};

extern "C" int b200_conv3x3_halo_launch(const CUtensorMap* map_x, const CUtensorMap* map_w, const CUtensorMap* map_y,
                                        const Conv3x3HaloParams* p, int grid, cudaStream_t stream);
