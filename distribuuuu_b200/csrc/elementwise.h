// Parameter blocks of the memory-bound kernels (elementwise.cu), shared with the host bindings.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

constexpr int kMaxPeers = 8;

// Peer-memory context for SyncBN-style statistic exchanges (all pointers are device addresses; peers' buffers
// are mapped into this process through symmetric memory, so plain loads/stores on them travel over NVLink).
struct PeerCtx {
  int world;                        // 1 => purely local
  int rank;
  uint32_t epoch;                   // (unused: kept for layout / debugging) host-side value of the exchange counter
  uint32_t* epoch_dev;              // device-resident exchange counter: the kernels read / advance it themselves, so a
                                    // captured CUDA graph of the step can be replayed (no per-step value in kernel arguments)
  int slot_base;                    // first signal-pad slot used by this exchange stream
  uint32_t* signal_pads[kMaxPeers]; // signal pad of every rank
  float* sym_bufs[kMaxPeers];       // statistics buffer of every rank
  int* ticket;                      // [3] local ints: arrival ticket / departure counter / producer-tail counter
  int presignaled;                  // the kernel that PRODUCED the statistics already told the peers (peer_signal_at_tail)
  float* mc_stats;                  // NVLS multicast alias of the statistics buffers (nullptr: P2P loads from every peer)
  float* reduced;                   // local scratch [2][C]: the cross-rank sums published by the designated CTA
  uint32_t* ready;                  // local flag: `reduced` holds exchange #epoch
  unsigned long long* wait_ns;      // optional: accumulated ns the designated CTA spent in the exchange (profiling)
};

// per-channel affine applied to raw uint8 pixels by the stem im2col: v * scale[c] + bias[c]  (= (v/255 - mean) / std)
struct StemNorm { float scale[8]; float bias[8]; };

struct BnApplyParams {
  const __nv_bfloat16* y; long long ldy;        // conv output [rows][C]
  const __nv_bfloat16* residual; long long ldr; // optional
  __nv_bfloat16* out; long long ldo;
  long long rows; int C;
  const float* stats;        // local [2][C] sum / sumsq (== sym_bufs[rank] + sym_offset when world > 1)
  long long sym_offset;      // float offset of this layer's statistics inside the symmetric buffer
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float* save_mean; float* save_invstd;
  float count;               // elements per channel over ALL ranks
  float eps, momentum;
  int act, training;
  uint8_t* relu_mask;        // optional [rows][C/8]: bit i of byte (row, cv) = (output channel cv*8+i > 0); saves the
                             // backward passes from re-reading the residual just to rebuild the ReLU mask
  PeerCtx peer;
};

struct BnBwdParams {
  const __nv_bfloat16* y; long long ldy;
  const __nv_bfloat16* dout; long long ldd;
  const __nv_bfloat16* residual; long long ldr;  // forward residual (needed to recompute act'(z)); optional
  __nv_bfloat16* dy; long long lddy;              // gradient wrt the conv output (own pitch: y may be a channel slice)
  __nv_bfloat16* dresidual;                       // optional gradient wrt the residual input (pitch ldr)
  long long rows; int C;
  float* sums;               // local [2][C]: sum(dz), sum(dz*xhat)
  long long sym_offset;
  const float* gamma; const float* beta;
  const float* save_mean; const float* save_invstd;
  float* dgamma; float* dbeta; // fp32 gradient slots (accumulated)
  float count;
  int act;
  const uint8_t* relu_mask;  // optional, see BnApplyParams
  PeerCtx peer;
};

// Stem tail: BN (batch statistics from the conv epilogue) + ReLU + 3x3 / stride-2 / pad-1 max-pool in ONE pass over the
// conv output, and its two-pass backward straight from the pooled gradient: the normalised 112x112 activation and its
// gradient (411 MB each at batch 256) are never materialised.  H and W even, C % 8 == 0.
struct BnPoolParams {
  const __nv_bfloat16* y;      // conv output [N][H][W][C], contiguous
  __nv_bfloat16* out;          // fwd: pooled activations [N][P][Q][C]
  const __nv_bfloat16* dout;   // bwd: gradient of the pooled activations
  uint8_t* arg;                // [N][P][Q][C]: window tap (0..8) that holds the maximum, 9 = no positive tap (ReLU-dead);
                               // written by fwd (optional), read by bwd
  __nv_bfloat16* dy;           // bwd: gradient wrt the conv output
  int N, H, W, C, P, Q;
  float* stats;                // fwd: local [2][C] sum / sumsq;  bwd: local [2][C] sum(dz), sum(dz*xhat)
  long long sym_offset;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var;
  float* save_mean; float* save_invstd;
  float* dgamma; float* dbeta;
  float count, eps, momentum;
  int training;
  PeerCtx peer;
};

extern "C" {
int b200_bn_apply(const BnApplyParams* p, cudaStream_t s);
int b200_bn_stats(const void* y, long long rows, int C, long long ldy, float* stats, cudaStream_t s);
int b200_bn_bwd_reduce(const BnBwdParams* p, cudaStream_t s);
int b200_bn_bwd_apply(const BnBwdParams* p, cudaStream_t s);
int b200_bn_relu_pool_fwd(const BnPoolParams* p, cudaStream_t s);
int b200_bn_relu_pool_bwd_reduce(const BnPoolParams* p, cudaStream_t s);
int b200_bn_relu_pool_bwd_apply(const BnPoolParams* p, cudaStream_t s);
int b200_maxpool_fwd(const void* x, void* out, void* argmax, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad, cudaStream_t s);
int b200_maxpool_bwd(const void* dout, const void* argmax, void* dx, int N, int H, int W, int C, int P, int Q, int k, int stride, int pad, cudaStream_t s);
int b200_gap_fwd(const void* x, void* out, int N, int HW, int C, cudaStream_t s);
int b200_gap_bwd(const void* dout, void* dx, int N, int HW, int C, cudaStream_t s);
int b200_avgpool2_fwd(const void* x, void* out, int N, int H, int W, int C, cudaStream_t s);
int b200_avgpool2_bwd(const void* dout, void* dx, int N, int H, int W, int C, cudaStream_t s);
int b200_channel_scale_fwd(const void* x, const void* gate, void* out, int N, int HW, int C, cudaStream_t s);
int b200_channel_scale_bwd(const void* dout, const void* x, const void* gate, void* dx, float* dgate, int N, int HW, int C, cudaStream_t s);
int b200_ce_topk(const void* logits, const long long* target, void* dlogits, float* accum, int rows, int ncls, long long ld, int topk, float grad_scale, cudaStream_t s);
int b200_nchw_to_nhwc(const float* x, void* out, int N, int C, int H, int W, cudaStream_t s);
int b200_stem_im2col(const void* x, int x_is_u8, void* patches, int N, int C, int H, int W, int P, int Q, int R, int S, int stride, int pad, int Kpad, const StemNorm* norm, cudaStream_t s);
int b200_pad_rows(const void* src, void* dst, int rows, int cols, int cols_pad, cudaStream_t s);
int b200_unpad_add(const float* src, float* dst, int rows, int cols, int cols_pad, cudaStream_t s);
}
