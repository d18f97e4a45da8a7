// Depthwise convolution kernels (dwconv.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

struct DwParams {
  const __nv_bfloat16* x;   // [N,H,W,C]   (fprop / wgrad input)
  const __nv_bfloat16* w;   // [C,k,k]     bf16 weights
  __nv_bfloat16* y;         // [N,P,Q,C]   fprop output; dgrad / wgrad read dY through this pointer
  __nv_bfloat16* dx;        // [N,H,W,C]   dgrad output
  float* dw;                // [C,k,k]     fp32 gradient (accumulated)
  float* stats;             // optional [2][C] BN statistics of y
  int N, H, W, C, P, Q, k, stride, pad;
};

extern "C" {
int b200_dw_fprop(const DwParams* p, cudaStream_t s);
int b200_dw_dgrad(const DwParams* p, cudaStream_t s);
int b200_dw_wgrad(const DwParams* p, cudaStream_t s);
}
