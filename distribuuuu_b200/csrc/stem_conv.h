// Parameter block of the space-to-depth stem forward kernel (stem_conv.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "elementwise.h"   // PeerCtx

struct StemConvParams {
  const void* s;        // space-to-depth image [N][Hs = P + 3][Ws = Q + 3][16] bf16 (extras.cu: stem_s2d_kernel)
  void* y;              // output [N][P][Q][64] bf16
  float* stats;         // optional [2][64] BN statistics of the output
  int N, P, Q, Hs, Ws;
  int tiles;            // N * P: one output row per tile (Q <= 128)
  PeerCtx peer;         // SyncBN: world > 1 => the last CTA announces the statistics exchange
};

extern "C" int b200_stem_conv_launch(const CUtensorMap* map_w, const StemConvParams* p, int grid, cudaStream_t stream);
