// Parameter block of the fused relative-position multi-head self-attention kernels (attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct AttnParams {
  int B, heads;
  float scale;          // logits = scale * (q.k + q.rel_w[dx] + q.rel_h[dy])
  void* p_save;         // bf16 [B*heads][196][208]  softmax probabilities (forward writes, backward reads)
  void* ds_save;        // bf16 [B*heads][196][256]  scale * dLogits (bwd_dq writes, bwd_dkv reads)
  void* dsrel;          // bf16 [B][196][heads*64]   scale * d(rel logits): cols h*64 + [0,27) width, h*64 + 32 + [0,27) height
};

extern "C" {
// maps: qk (C=2*heads*128, 196, B) boxes (64,128,1) and (64,64,1); v / out / dout / dv (C=heads*128, 196, B); rel_w / rel_h (128, 27, 1)
int b200_attn_fwd(const CUtensorMap* qk128, const CUtensorMap* v64, const CUtensorMap* relw, const CUtensorMap* relh,
                  const CUtensorMap* out32, const AttnParams* p, cudaStream_t s);
int b200_attn_bwd_dq(const CUtensorMap* dout128, const CUtensorMap* v128, const CUtensorMap* qk64, const CUtensorMap* relw,
                     const CUtensorMap* relh, const CUtensorMap* dqk32, const AttnParams* p, cudaStream_t s);
int b200_attn_bwd_dkv(const CUtensorMap* psave64, const CUtensorMap* dssave64, const CUtensorMap* dout64, const CUtensorMap* qk64,
                      const CUtensorMap* dv32, const CUtensorMap* dqk32, const AttnParams* p, cudaStream_t s);
int b200_rel_grad_reduce(const float* dw, float* grad_w, float* grad_h, int heads, cudaStream_t s);
}
