// Gradient all-reduce fused with the SGD-momentum update, over NVLink peer memory (SURVEY K4 + G13 + G19):
//
//   phase A  local fp32 gradients * (1/world) -> bf16, written into this rank's symmetric staging buffer
//            (and the fp32 gradient slots are zeroed for the next accumulation)
//   barrier  flag exchange on the signal pads (st.release.sys / ld.acquire.sys), no NCCL
//   phase B  two-shot : each rank reduces ITS shard of the bucket -- one multimem.ld_reduce (in-switch NVLS
//                       reduction, fp32 accumulate) per 16 bytes, or P2P loads from every peer -- applies
//                       weight decay + Nesterov momentum to its shard of the fp32 master weights, and
//                       broadcasts the updated bf16 weights to all ranks with multimem.st (or P2P stores);
//            one-shot : every rank reduces the whole (small) bucket and updates its own replica.
//   barrier  so nobody reuses the staging buffer / reads weights before all peers are done
//
// torch.optim.SGD semantics (g += wd*p; buf = mu*buf + (1-damp)*g [buf = g on the first step];
// g = g + mu*buf if nesterov else buf; p -= lr*g), so optimizer state interchanges with the reference.
#include "common.cuh"
#include "comm.h"

namespace b200 {

__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t* p, uint32_t v) {   // ordered by the fence before it
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_16B(void* mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Grid-wide + cross-rank barrier.  All CTAs of the launch must be able to become resident (grid <= #SMs).
__device__ void rank_barrier(const CommCtx& c, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int arrived = atomicAdd(c.local_counter, 1);
    if (arrived == (int)gridDim.x - 1) {
      *c.local_counter = 0;
      if (c.world > 1) {
        __threadfence_system();   // the other CTAs' writes (seen through the counter) before the flags; then relaxed
        for (int r = 0; r < c.world; ++r) st_relaxed_sys_u32(c.signal_pads[r] + c.slot_base + c.rank, epoch);  // stores pipeline
        for (int r = 0; r < c.world; ++r) {
          const uint32_t* flag = c.signal_pads[c.rank] + c.slot_base + r;
          long long t0 = clock64();
          while ((int)(ld_acquire_sys_u32(flag) - epoch) < 0) {
            if (clock64() - t0 > B200_SPIN_LIMIT_CYCLES * 10) {
              printf("b200: rank barrier timed out (rank %d waiting for %d, epoch %u)\n", c.rank, r, epoch);
              __trap();
            }
          }
        }
      }
      __threadfence();
      st_release_gpu_u32(c.local_release, epoch);
    } else {
      long long t0 = clock64();
      while ((int)(ld_acquire_gpu_u32(c.local_release) - epoch) < 0) {
        if (clock64() - t0 > B200_SPIN_LIMIT_CYCLES * 12) { printf("b200: grid barrier timed out\n"); __trap(); }
      }
    }
  }
  __syncthreads();
}

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_f8(const float* p) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  return {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void st_f8(float* p, const F8& f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  return make_uint4(pack_bf16x2(f.v[0], f.v[1]), pack_bf16x2(f.v[2], f.v[3]), pack_bf16x2(f.v[4], f.v[5]),
                    pack_bf16x2(f.v[6], f.v[7]));
}
__device__ __forceinline__ F8 unpack8(uint4 u) {
  F8 f;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f.v[2 * i] = __uint_as_float(w[i] << 16);
    f.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return f;
}

__device__ __forceinline__ void sgd_math(F8& w, F8& m, const F8& g_in, const SgdHyper& h) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float g = fmaf(h.weight_decay, w.v[i], g_in.v[i]);
    float buf = h.first_step ? g : fmaf(h.momentum, m.v[i], (1.f - h.dampening) * g);
    m.v[i] = buf;
    g = h.nesterov ? fmaf(h.momentum, buf, g) : buf;
    w.v[i] = fmaf(-h.lr, g, w.v[i]);
  }
}

// ---- world == 1 (or NCCL-reduced gradients): plain fused optimizer over a flat range ------------
__global__ void sgd_local_kernel(float* __restrict__ master, float* __restrict__ mom, float* __restrict__ grad,
                                 __nv_bfloat16* __restrict__ w16, long long n8, SgdHyper h, float grad_scale, int zero_grad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    F8 g = ld_f8(grad + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] *= grad_scale;
    F8 w = ld_f8(master + i * 8), m = ld_f8(mom + i * 8);
    sgd_math(w, m, g, h);
    st_f8(master + i * 8, w);
    st_f8(mom + i * 8, m);
    if (w16) reinterpret_cast<uint4*>(w16)[i] = pack8(w);
    if (zero_grad) { F8 z = {{0, 0, 0, 0, 0, 0, 0, 0}}; st_f8(grad + i * 8, z); }
  }
}

// fp32 master -> bf16 compute copy (initialisation / after loading a checkpoint)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = pack8(ld_f8(src + i * 8));
}

// ---- fused peer-memory all-reduce + SGD -----------------------------------------------------------
// All offsets / lengths are in units of 8 elements (16 bytes of bf16).
// 256 threads x <= 64 registers: a CTA of this kernel must fit on an SM next to a resident conv / BN CTA of the
// backward pass it overlaps with (the persistent conv kernel keeps ~48k of the 64k registers of every SM).
__global__ void __launch_bounds__(256, 4) allreduce_sgd_kernel(AllreduceSgdParams p) {
  const CommCtx& c = p.comm;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const float scale = 1.f / c.world;
  // every thread reads the barrier counter before the first barrier; thread 0 of CTA 0 advances it after the second one
  const uint32_t epoch0 = *reinterpret_cast<volatile uint32_t*>(c.epoch_dev);

  // phase A: fp32 grad -> scaled bf16 staging (own symmetric buffer), zero the fp32 slots
  __nv_bfloat16* my_stage = c.stage[c.rank];
  for (long long i = tid; i < p.n8; i += nthreads) {
    const long long e = (p.off8 + i) * 8;
    F8 g = ld_f8(p.grad + e);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] *= scale;
    reinterpret_cast<uint4*>(my_stage)[p.off8 + i] = pack8(g);
    F8 z = {{0, 0, 0, 0, 0, 0, 0, 0}};
    st_f8(p.grad + e, z);
  }
  rank_barrier(c, epoch0 + 1);

  // phase B
  long long b0 = 0, b1 = p.n8;  // one-shot: whole bucket on every rank
  if (!p.one_shot) {             // two-shot: my shard only
    const long long per = (p.n8 + c.world - 1) / c.world;
    b0 = per * c.rank < p.n8 ? per * c.rank : p.n8;
    b1 = b0 + per < p.n8 ? b0 + per : p.n8;
  }
  // four 16-byte reductions in flight per thread: one NVLink round trip (~2 us) per multimem.ld_reduce made the
  // single-load loop latency-bound (305 GB/s bus on 12 MiB buckets in round 1)
  constexpr int U = 4;
  for (long long i0 = b0 + tid; i0 < b1; i0 += U * nthreads) {
    uint4 raw[U];
    if (c.mc_stage != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * nthreads;
        if (i < b1) raw[u] = multimem_ld_reduce_bf16x8(reinterpret_cast<const uint4*>(c.mc_stage) + p.off8 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nthreads;
      if (i >= b1) break;
      const long long v = p.off8 + i;
      F8 g;
      if (c.mc_stage != nullptr) {
        g = unpack8(raw[u]);
      } else {
        uint4 t[kCommMaxPeers];
#pragma unroll
        for (int r = 0; r < kCommMaxPeers; ++r)      // P2P loads over NVLink, all issued before the first add
          if (r < c.world) t[r] = reinterpret_cast<const uint4*>(c.stage[r])[v];
#pragma unroll
        for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
#pragma unroll
        for (int r = 0; r < kCommMaxPeers; ++r) {
          if (r < c.world) {
            const F8 f = unpack8(t[r]);
#pragma unroll
            for (int k = 0; k < 8; ++k) g.v[k] += f.v[k];
          }
        }
      }
      F8 w = ld_f8(p.master + v * 8), m = ld_f8(p.mom + v * 8);
      sgd_math(w, m, g, p.hyper);
      st_f8(p.master + v * 8, w);
      st_f8(p.mom + v * 8, m);
      const uint4 wb = pack8(w);
      if (p.one_shot) {
        reinterpret_cast<uint4*>(c.w16[c.rank])[v] = wb;
      } else if (c.mc_w16 != nullptr) {
        multimem_st_16B(reinterpret_cast<uint4*>(c.mc_w16) + v, wb);              // in-switch broadcast
      } else {
        for (int r = 0; r < c.world; ++r) reinterpret_cast<uint4*>(c.w16[r])[v] = wb;  // P2P stores
      }
    }
  }
  rank_barrier(c, epoch0 + 2);
  if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(c.epoch_dev) = epoch0 + 2;
}

__global__ void rank_barrier_kernel(CommCtx c, uint32_t /*unused*/) {
  const uint32_t epoch0 = *reinterpret_cast<volatile uint32_t*>(c.epoch_dev);
  rank_barrier(c, epoch0 + 1);
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(c.epoch_dev) = epoch0 + 1;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sgd_local(float* master, float* mom, float* grad, void* w16, long long n, const SgdHyper* h,
                              float grad_scale, int zero_grad, cudaStream_t s) {
  const long long n8 = n / 8;
  long long g = (n8 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  sgd_local_kernel<<<(int)g, 256, 0, s>>>(master, mom, grad, (__nv_bfloat16*)w16, n8, *h, grad_scale, zero_grad);
  return (int)cudaGetLastError();
}
extern "C" int b200_cast_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
  const long long n8 = n / 8;
  long long g = (n8 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  cast_bf16_kernel<<<(int)g, 256, 0, s>>>(src, (__nv_bfloat16*)dst, n8);
  return (int)cudaGetLastError();
}
extern "C" int b200_allreduce_sgd(const AllreduceSgdParams* p, int grid, cudaStream_t s) {
  allreduce_sgd_kernel<<<grid, 256, 0, s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_rank_barrier(const CommCtx* c, uint32_t epoch, cudaStream_t s) {
  rank_barrier_kernel<<<1, 32, 0, s>>>(*c, epoch);
  return (int)cudaGetLastError();
}
