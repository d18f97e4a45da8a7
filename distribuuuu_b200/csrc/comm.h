// Parameter blocks of the peer-memory gradient all-reduce + fused SGD (comm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kCommMaxPeers = 8;

struct SgdHyper {
  float lr, momentum, dampening, weight_decay;
  int nesterov, first_step;
};

struct CommCtx {
  int world, rank;
  int slot_base;                           // first signal-pad slot of this barrier stream
  uint32_t* signal_pads[kCommMaxPeers];    // every rank's signal pad (symmetric memory)
  __nv_bfloat16* stage[kCommMaxPeers];     // every rank's bf16 gradient staging buffer
  __nv_bfloat16* w16[kCommMaxPeers];       // every rank's bf16 weight buffer
  __nv_bfloat16* mc_stage;                 // NVLS multicast alias of the staging buffers (nullptr: P2P loads)
  __nv_bfloat16* mc_w16;                   // NVLS multicast alias of the weight buffers   (nullptr: P2P stores)
  uint32_t* epoch_dev;                     // device-resident barrier counter (graph-replayable: no per-launch epoch argument)
  int* local_counter;                      // grid barrier arrival counter (device-local)
  uint32_t* local_release;                 // grid barrier release flag (device-local)
};

struct AllreduceSgdParams {
  CommCtx comm;
  float* master; float* mom; float* grad;  // flat fp32 buffers (device-local)
  long long off8, n8;                      // bucket offset / length in units of 8 elements
  SgdHyper hyper;
  uint32_t epoch;                          // (unused; the kernel reads comm.epoch_dev: its barriers are counter+1 and counter+2)
  int one_shot;
};

extern "C" {
int b200_sgd_local(float* master, float* mom, float* grad, void* w16, long long n, const SgdHyper* h, float grad_scale, int zero_grad, cudaStream_t s);
int b200_cast_bf16(const float* src, void* dst, long long n, cudaStream_t s);
int b200_allreduce_sgd(const AllreduceSgdParams* p, int grid, cudaStream_t s);
int b200_rank_barrier(const CommCtx* c, uint32_t epoch, cudaStream_t s);
}
