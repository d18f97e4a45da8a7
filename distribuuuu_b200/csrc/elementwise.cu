// Memory-bound NHWC bf16 kernels: BatchNorm (apply / backward, with optional peer-memory SyncBN reduction),
// pooling, fused softmax-cross-entropy-topk, input layout conversion, stem im2col.
// Replaces the ATen/cuDNN calls the reference makes for BN/ReLU/add/pool/CE/topk.
// All kernels process 8 channels (one 16-byte vector) per thread and keep per-channel state in registers.
#include "common.cuh"
#include "elementwise.h"

namespace b200 {

constexpr int VEC = 8;

struct alignas(16) BF8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  BF8 raw = *reinterpret_cast<const BF8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(raw.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
// raw 16-byte streaming load (kept in registers until all loads of an unrolled batch are issued)
__device__ __forceinline__ uint4 ldg_stream(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// raw 16-byte load through L1 (neighbouring threads re-read the same lines: overlapping pooling windows)
__device__ __forceinline__ uint4 ldg_cached(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
// cp.async ring: global -> shared without holding registers, so many rows per thread can be in flight
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int kBnThreads = 256;
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  BF8 raw;
#pragma unroll
  for (int i = 0; i < 4; ++i) raw.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<BF8*>(p) = raw;
}
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// Peer exchange used by SyncBN: every rank publishes "my statistics for exchange #epoch are final" to all peers'
// signal pads, then waits until every peer has done the same.  Statistics are then read straight out of the peers'
// symmetric buffers (P2P loads over NVLink) -- no NCCL call, no staging copy.  
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// flag store AFTER an explicit __threadfence_system(): one thread raising `world` flags with st.release pays one NVLink
// round trip per peer (each release waits for the previous remote store); relaxed stores behind one fence pipeline
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// in-switch (NVLS) sum over every rank's copy of 4 consecutive floats of the symmetric statistics buffer
__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Producer side of the exchange: called by EVERY thread at the very end of the kernel that accumulates the local
// statistics (conv epilogue atomics, bn_bwd_reduce REDs).  The last CTA to get here tells all peers "my statistics for
// exchange #epoch are final" -- so the flag crosses NVLink during the kernel boundary (and, in backward, during the
// weight-gradient GEMM the engine schedules between the reduce and the apply pass) instead of inside the consumer.
__device__ void peer_signal_at_tail(const PeerCtx& pc, unsigned total_ctas) {
  __threadfence();                    // this thread's atomics / REDs on the statistics are ordered before the counter
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int done = atomicAdd(pc.ticket + 2, 1);
    if (done == (int)total_ctas - 1) {
      pc.ticket[2] = 0;
      const uint32_t e = *reinterpret_cast<volatile uint32_t*>(pc.epoch_dev) + 1u;   // this exchange
      *reinterpret_cast<volatile uint32_t*>(pc.epoch_dev) = e;                       // the consumer kernel reads it back
      __threadfence_system();
      for (int r = 0; r < pc.world; ++r) st_relaxed_sys(pc.signal_pads[r] + pc.slot_base + pc.rank, e);
    }
  }
}

// Cross-rank reduction of one layer's [2][C] statistics, once per launch instead of once per CTA:
//   * the first CTA to take a ticket is the *designated* CTA: it tells every peer "my statistics for exchange
//     #epoch are final" (flags on the signal pads), waits for the peers' flags, sums the statistics of all ranks
//     -- ONE multimem.ld_reduce per 16 bytes (the NVSwitch adds the eight copies), or P2P loads from every peer
//     issued back to back when there is no multicast mapping -- into the local `reduced` scratch and releases a
//     device-local flag;
//   * every other CTA only spins on that local flag (its activation loads are already in flight) and then reads
//     `reduced` from L2.
// Round 1 had every CTA poll all peer flags and then do `world` dependent P2P loads per thread (592 CTAs x 8 peers
// per launch, ~40 us per exchange at 8 GPUs); this is one NVLink round trip issued by 256 threads.
__device__ void peer_exchange_reduce(const PeerCtx& pc, long long sym_offset, int C) {
  __shared__ int s_ticket;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  // exchange number: already advanced by the producer kernel's tail (presignaled), else this launch is exchange counter + 1
  // and its last CTA advances the counter on the way out -- nobody writes it while CTAs of this launch may still read it
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(pc.epoch_dev) + (pc.presignaled ? 0u : 1u);
  if (tid == 0) s_ticket = atomicAdd(pc.ticket, 1);
  __syncthreads();
  if (s_ticket == 0) {
    unsigned long long t_begin = 0;
    if (tid == 0 && pc.wait_ns) t_begin = global_timer_ns();
    if (tid < pc.world) {
      if (!pc.presignaled) {
        __threadfence_system();
        st_release_sys(pc.signal_pads[tid] + pc.slot_base + pc.rank, epoch);
      }
      const uint32_t* flag = pc.signal_pads[pc.rank] + pc.slot_base + tid;
      long long t0 = clock64();
      while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > B200_SPIN_LIMIT_CYCLES * 20) {
          printf("b200: peer exchange timed out (rank %d waiting for %d, epoch %u)\n", pc.rank, tid, epoch);
          __trap();
        }
      }
    }
    __syncthreads();
    const int nvec = C / 2;                        // float4 vectors in [2][C]
    if (pc.mc_stats != nullptr) {
      // up to four reductions in flight per thread: one NVLink round trip for C <= 2048 (a dependent loop paid one
      // round trip per 256 vectors, i.e. four of them on the 2048-channel layers)
      for (int i0 = tid; i0 < nvec; i0 += 4 * nthreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * nthreads < nvec) v[u] = multimem_ld_reduce_f32x4(pc.mc_stats + sym_offset + 4 * (long long)(i0 + u * nthreads));
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * nthreads < nvec) reinterpret_cast<float4*>(pc.reduced)[i0 + u * nthreads] = v[u];
      }
    } else {
      for (int i = tid; i < nvec; i += nthreads) {
        float4 v[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)          // all peers' loads are in flight before the first add
          if (r < pc.world) v[r] = reinterpret_cast<const float4*>(pc.sym_bufs[r] + sym_offset)[i];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
          if (r < pc.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        reinterpret_cast<float4*>(pc.reduced)[i] = acc;
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release_gpu(pc.ready, epoch);
      if (pc.wait_ns) atomicAdd(pc.wait_ns, global_timer_ns() - t_begin);
    }
  } else {
    if (tid == 0) {
      long long t0 = clock64();
      while ((int)(ld_acquire_gpu(pc.ready) - epoch) < 0) {
        if (clock64() - t0 > B200_SPIN_LIMIT_CYCLES * 24) { printf("b200: SyncBN local release timed out\n"); __trap(); }
      }
    }
    __syncthreads();
  }
  // last CTA out resets the ticket for the next launch that uses it
  if (tid == 0) {
    int done = atomicAdd(pc.ticket + 1, 1);
    if (done == (int)(gridDim.x * gridDim.y) - 1) {
      pc.ticket[0] = 0; pc.ticket[1] = 0;
      if (!pc.presignaled) *reinterpret_cast<volatile uint32_t*>(pc.epoch_dev) = epoch;
      __threadfence();
    }
  }
}

// The [2][C] statistics of this thread's 8 channels: the local sums when world == 1, else the cross-rank sums the
// designated CTA published (L2 loads: another CTA of this launch wrote them).
__device__ __forceinline__ void gather_stats(const PeerCtx& pc, const float* local, long long sym_offset, int C, int c0,
                                             float (&s0)[8], float (&s1)[8]) {
  const float* base = (pc.world <= 1) ? local : pc.reduced;
  const float4* a = reinterpret_cast<const float4*>(base + c0);
  const float4* b = reinterpret_cast<const float4*>(base + C + c0);
  const float4 a0 = __ldcg(a), a1 = __ldcg(a + 1), b0 = __ldcg(b), b1 = __ldcg(b + 1);
  s0[0] = a0.x; s0[1] = a0.y; s0[2] = a0.z; s0[3] = a0.w; s0[4] = a1.x; s0[5] = a1.y; s0[6] = a1.z; s0[7] = a1.w;
  s1[0] = b0.x; s1[1] = b0.y; s1[2] = b0.z; s1[3] = b0.w; s1[4] = b1.x; s1[5] = b1.y; s1[6] = b1.z; s1[7] = b1.w;
}

// ------------------------------------------------------------------------------------------------
// BN forward: finalize statistics (optionally across ranks), update running stats, normalise + affine
// (+ residual) + activation in ONE pass over the conv output.
// grid: (ceil(C/8 / blockDim.x), row_chunks); block: (cvx, rows_per_block)
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float z) {
  if (ACT == ACT_RELU) return fmaxf(z, 0.f);
  if (ACT == ACT_SILU) return z / (1.f + __expf(-z));
  return z;
}
template <bool RES> struct FwdRing {
  static constexpr int kArr = RES ? 2 : 1;
  static constexpr int kDepth = RES ? 4 : 8;
  static constexpr int kBytes = kDepth * kArr * kBnThreads * 16;   // 32 KB
};
template <int ACT, bool RES>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(BnApplyParams p) {
  using R = FwdRing<RES>;
  constexpr int D = R::kDepth, A = R::kArr;
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  uint4* ring = reinterpret_cast<uint4*>(dyn_raw);  // [D][A][256]
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const bool active = c0 < p.C;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const long long rstride = (long long)gridDim.y * blockDim.y;
  const long long row0 = (long long)blockIdx.y * blockDim.y + threadIdx.y;
  auto issue = [&](int stage, long long r) {
    if (active && r < p.rows) {
      cp_async16(&ring[(stage * A + 0) * kBnThreads + tid], p.y + r * p.ldy + c0);
      if (RES) cp_async16(&ring[(stage * A + A - 1) * kBnThreads + tid], p.residual + r * p.ldr + c0);
    }
    cp_async_commit();
  };
  // the conv output does not depend on the peers' statistics: start streaming it before waiting on the exchange
  long long r_issue = row0;
#pragma unroll
  for (int st = 0; st < D - 1; ++st) { issue(st, r_issue); r_issue += rstride; }
  if (p.peer.world > 1 && p.training) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
  if (!active) { cp_async_wait<0>(); return; }
  float scale[8], shift[8];
  {
    float g[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = 1.f; b[i] = 0.f; }
    if (p.gamma) ld8f(p.gamma + c0, g);
    if (p.beta) ld8f(p.beta + c0, b);
    if (p.training) {
      float s0[8], s1[8];
      gather_stats(p.peer, p.stats, p.sym_offset, p.C, c0, s0, s1);
      const float inv_n = 1.f / p.count;
      const bool writer = (blockIdx.y == 0 && threadIdx.y == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mean = s0[i] * inv_n;
        const float var = fmaxf(s1[i] * inv_n - mean * mean, 0.f);
        const float invstd = rsqrtf(var + p.eps);
        scale[i] = g[i] * invstd;
        shift[i] = b[i] - mean * scale[i];
        if (writer) {
          p.save_mean[c0 + i] = mean;
          p.save_invstd[c0 + i] = invstd;
          if (p.running_mean) {
            const float unbiased = var * (p.count / fmaxf(p.count - 1.f, 1.f));
            p.running_mean[c0 + i] = (1.f - p.momentum) * p.running_mean[c0 + i] + p.momentum * mean;
            p.running_var[c0 + i] = (1.f - p.momentum) * p.running_var[c0 + i] + p.momentum * unbiased;
          }
        }
      }
    } else {
      float rm[8], rv[8];
      ld8f(p.running_mean + c0, rm);
      ld8f(p.running_var + c0, rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        scale[i] = g[i] * rsqrtf(rv[i] + p.eps);
        shift[i] = b[i] - rm[i] * scale[i];
      }
    }
  }
  int st = 0;
  for (long long r = row0; r < p.rows; r += rstride) {
    issue(st == 0 ? D - 1 : st - 1, r_issue);
    r_issue += rstride;
    cp_async_wait<D - 1>();
    float x[8];
    unpack8f(ring[(st * A + 0) * kBnThreads + tid], x);
    if (RES) {
      float q[8];
      unpack8f(ring[(st * A + A - 1) * kBnThreads + tid], q);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = act_fwd_t<ACT>(fmaf(x[i], scale[i], shift[i]) + q[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = act_fwd_t<ACT>(fmaf(x[i], scale[i], shift[i]));
    }
    store8(p.out + r * p.ldo + c0, x);
    if (p.relu_mask) {
      uint32_t bits = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) bits |= (x[i] > 0.f ? 1u : 0u) << i;
      p.relu_mask[r * (p.C / VEC) + cv] = (uint8_t)bits;
    }
    st = (st + 1 == D) ? 0 : st + 1;
  }
  cp_async_wait<0>();
}

// Per-channel sum / sum-of-squares of a [rows][C] bf16 tensor (used when the producer is not our conv kernel).
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long rows, int C, long long ldy, float* stats) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  if (c0 < C) {
    constexpr int U = 4;
    const long long rstride = (long long)gridDim.y * blockDim.y;
    for (long long row = (long long)blockIdx.y * blockDim.y + threadIdx.y; row < rows; row += U * rstride) {
      uint4 ry[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (row + u * rstride < rows) ry[u] = ldg_stream(y + (row + u * rstride) * ldy + c0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (row + u * rstride < rows) {
          float x[8];
          unpack8f(ry[u], x);
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] += x[i]; b[i] = fmaf(x[i], x[i], b[i]); }
        }
      }
    }
  }
  // reduce over threadIdx.y through shared memory
  const int tx = threadIdx.x, ty = threadIdx.y;
  extern __shared__ float dyn[];
  float* sm = dyn;  // [blockDim.y][blockDim.x][16]
  float* mine = sm + ((size_t)ty * blockDim.x + tx) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = a[i]; mine[8 + i] = b[i]; }
  __syncthreads();
  if (ty == 0 && c0 < C) {
    for (int yy = 1; yy < blockDim.y; ++yy) {
      const float* o = sm + ((size_t)yy * blockDim.x + tx) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += o[i]; b[i] += o[8 + i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(stats + c0 + i, a[i]); atomicAdd(stats + C + c0 + i, b[i]); }
  }
}

// ------------------------------------------------------------------------------------------------
// BN backward.  Both passes are specialised at compile time on the activation and on how act'(z) is obtained:
//   BWD_PLAIN : recomputed from the conv output y (z = y*scale + shift)
//   BWD_MASK  : read from the 1-bit/element ReLU mask the forward pass stored (residual layers)
//   BWD_RES   : recomputed from y and the forward residual (three input streams)
// so the streaming loops are straight-line code.  The cp.async ring is 6 deep for two input streams and 4 deep for
// three (48 KB either way); the mask byte of a row is fetched into a register when that row's cp.async is issued,
// i.e. RING-1 iterations before it is consumed, so no global-load latency is exposed inside the loop.
// ------------------------------------------------------------------------------------------------
enum { BWD_PLAIN = 0, BWD_MASK = 1, BWD_RES = 2 };

__device__ __forceinline__ void red_add_v4(float* dst, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
template <int ACT>
__device__ __forceinline__ float act_bwd_t(float z) {
  if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (ACT == ACT_SILU) { const float s = 1.f / (1.f + __expf(-z)); return s * (1.f + z * (1.f - s)); }
  return 1.f;
}
template <int MODE> struct BwdRing {
  static constexpr int kArr = (MODE == BWD_RES) ? 3 : 2;
  static constexpr int kDepth = (MODE == BWD_RES) ? 4 : 6;
  static constexpr int kBytes = kDepth * kArr * kBnThreads * 16;   // 48 KB
};
// dz = dout * act'(z) for the 8 channels of one row
template <int ACT, int MODE>
__device__ __forceinline__ void bwd_dz(float (&d)[8], const float (&y)[8], const float (&q)[8], uint32_t mbits,
                                       const float (&scale)[8], const float (&shift)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (MODE == BWD_MASK) d[i] = ((mbits >> i) & 1u) ? d[i] : 0.f;
    else if (ACT != ACT_NONE) d[i] *= act_bwd_t<ACT>(fmaf(y[i], scale[i], shift[i]) + (MODE == BWD_RES ? q[i] : 0.f));
  }
}

// pass 1: per-channel sum(dz) and sum(dz * xhat)
template <int ACT, int MODE>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(BnBwdParams p) {
  using R = BwdRing<MODE>;
  constexpr int D = R::kDepth, A = R::kArr;
  extern __shared__ __align__(16) float dyn[];
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  if (c0 < p.C) {
    uint4* ring = reinterpret_cast<uint4*>(dyn);  // [D][A][256]
    const long long rstride = (long long)gridDim.y * blockDim.y;
    const long long row0 = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    const uint8_t* mrow = p.relu_mask + cv;
    const int mpitch = p.C / VEC;
    uint32_t mq[D - 1];
    auto issue = [&](int stage, long long r, uint32_t& m, bool with_mask) {
      if (with_mask) m = 0u;
      if (r < p.rows) {
        cp_async16(&ring[(stage * A + 0) * kBnThreads + tid], p.y + r * p.ldy + c0);
        cp_async16(&ring[(stage * A + 1) * kBnThreads + tid], p.dout + r * p.ldd + c0);
        if (MODE == BWD_RES) cp_async16(&ring[(stage * A + A - 1) * kBnThreads + tid], p.residual + r * p.ldr + c0);
        if (MODE == BWD_MASK && with_mask) m = (uint32_t)__ldg(mrow + r * mpitch);
      }
      cp_async_commit();
    };
    long long r_issue = row0;
#pragma unroll
    for (int st = 0; st < D - 1; ++st) { issue(st, r_issue, mq[st], true); r_issue += rstride; }
    // per-channel constants (loaded after the ring is primed so their latency overlaps the first rows)
    float invstd[8], nmi[8], scale[8], shift[8];
    {
      float mean[8];
      ld8f(p.save_mean + c0, mean);
      ld8f(p.save_invstd + c0, invstd);
      float g[8], be[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { g[i] = 1.f; be[i] = 0.f; }
      if (ACT != ACT_NONE && MODE != BWD_MASK) {
        if (p.gamma) ld8f(p.gamma + c0, g);
        if (p.beta) ld8f(p.beta + c0, be);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        nmi[i] = -mean[i] * invstd[i];
        scale[i] = g[i] * invstd[i];
        shift[i] = be[i] - mean[i] * scale[i];
      }
    }
    // The loop is unrolled by the number of rows in flight so that the prefetched mask bytes keep fixed register
    // names (rotating them with moves would make every iteration wait for the load issued just before it).
    int st = 0;
    long long r = row0;
    while (r < p.rows) {
#pragma unroll
      for (int j = 0; j < D - 1; ++j) {
        if (r >= p.rows) break;
        uint32_t unused;
        issue(st == 0 ? D - 1 : st - 1, r_issue, unused, false);
        cp_async_wait<D - 1>();
        float y[8], d[8], q[8];
        unpack8f(ring[(st * A + 0) * kBnThreads + tid], y);
        unpack8f(ring[(st * A + 1) * kBnThreads + tid], d);
        if (MODE == BWD_RES) unpack8f(ring[(st * A + A - 1) * kBnThreads + tid], q);
        bwd_dz<ACT, MODE>(d, y, q, mq[j], scale, shift);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] += d[i];
          b[i] = fmaf(d[i], fmaf(y[i], invstd[i], nmi[i]), b[i]);
        }
        // reload this slot's mask byte only now that the old one is dead: the load lands in the same register
        if (MODE == BWD_MASK) mq[j] = (r_issue < p.rows) ? (uint32_t)__ldg(mrow + r_issue * mpitch) : 0u;
        r_issue += rstride;
        st = (st + 1 == D) ? 0 : st + 1;
        r += rstride;
      }
    }
    cp_async_wait<0>();
  }
  // CTA-level reduction over threadIdx.y through shared memory (component-major: conflict-free), then one vector
  // RED per 4 channels -- a quarter of the atomic operations of a scalar flush (the L2 atomic unit serialises them)
  __syncthreads();   // the ring memory is reused
#pragma unroll
  for (int i = 0; i < 8; ++i) { dyn[i * kBnThreads + tid] = a[i]; dyn[(8 + i) * kBnThreads + tid] = b[i]; }
  __syncthreads();
  const int bdx = blockDim.x, bdy = blockDim.y;
  if (tid < 4 * bdx) {
    const int quad = tid / bdx, tx = tid - quad * bdx;
    const int c = (blockIdx.x * bdx + tx) * VEC;
    if (c < p.C) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* src = dyn + (quad * 4) * kBnThreads + tx;
      for (int yy = 0; yy < bdy; ++yy) {
        s0 += src[yy * bdx];
        s1 += src[kBnThreads + yy * bdx];
        s2 += src[2 * kBnThreads + yy * bdx];
        s3 += src[3 * kBnThreads + yy * bdx];
      }
      red_add_v4(p.sums + (quad >= 2 ? p.C : 0) + c + (quad & 1) * 4, s0, s1, s2, s3);
    }
  }
  if (p.peer.world > 1) peer_signal_at_tail(p.peer, gridDim.x * gridDim.y);   // SyncBN: the exchange starts here
}

// pass 2: dy = (dz - mean(dz) - xhat * mean(dz*xhat)) * gamma * invstd (means over all ranks for SyncBN);
// also emits d(residual) = dz and the local dgamma / dbeta.
template <int ACT, int MODE>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnBwdParams p) {
  using R = BwdRing<MODE>;
  constexpr int D = R::kDepth, A = R::kArr;
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  uint4* ring = reinterpret_cast<uint4*>(dyn_raw);  // [D][A][256]
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const bool active = c0 < p.C;
  const long long rstride = (long long)gridDim.y * blockDim.y;
  const long long k0 = (long long)blockIdx.y * blockDim.y + threadIdx.y;
  const uint8_t* mrow = p.relu_mask + cv;
  const int mpitch = p.C / VEC;
  uint32_t mq[D - 1];
  // rows are visited from the END of the tensor: the reduce pass finished there, so those lines are the likeliest L2 hits
  auto issue = [&](int stage, long long kk, uint32_t& m, bool with_mask) {
    if (with_mask) m = 0u;
    if (active && kk < p.rows) {
      const long long r = p.rows - 1 - kk;
      cp_async16(&ring[(stage * A + 0) * kBnThreads + tid], p.y + r * p.ldy + c0);
      cp_async16(&ring[(stage * A + 1) * kBnThreads + tid], p.dout + r * p.ldd + c0);
      if (MODE == BWD_RES) cp_async16(&ring[(stage * A + A - 1) * kBnThreads + tid], p.residual + r * p.ldr + c0);
      if (MODE == BWD_MASK && with_mask) m = (uint32_t)__ldg(mrow + r * mpitch);
    }
    cp_async_commit();
  };
  // the activations do not depend on the peers' statistics: get them moving before waiting on the exchange
  long long k_issue = k0;
#pragma unroll
  for (int st = 0; st < D - 1; ++st) { issue(st, k_issue, mq[st], true); k_issue += rstride; }
  if (p.peer.world > 1) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
  if (!active) { cp_async_wait<0>(); return; }
  // dy = (dz - mean(dz) - xhat*mean(dz*xhat)) * gamma*invstd  ==  dz*scale + y*ca + cb
  float scale[8], shift[8], ca[8], cb[8];
  {
    float s0[8], s1[8];
    gather_stats(p.peer, p.sums, p.sym_offset, p.C, c0, s0, s1);
    const float inv_n = 1.f / p.count;
    float mean[8], invstd[8], g[8], be[8];
    ld8f(p.save_mean + c0, mean);
    ld8f(p.save_invstd + c0, invstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = 1.f; be[i] = 0.f; }
    if (p.gamma) ld8f(p.gamma + c0, g);
    if (p.beta) ld8f(p.beta + c0, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m_dz = s0[i] * inv_n, m_dzx = s1[i] * inv_n;
      scale[i] = g[i] * invstd[i];
      shift[i] = be[i] - mean[i] * scale[i];
      ca[i] = -invstd[i] * m_dzx * scale[i];
      cb[i] = -m_dz * scale[i] - mean[i] * ca[i];
    }
    if (blockIdx.y == 0 && threadIdx.y == 0 && p.dgamma) {
      // parameter gradients use the LOCAL sums (the gradient all-reduce averages them afterwards, as DDP does)
      float lg[8], lb[8], dg[8], db[8];
      ld8f(p.sums + p.C + c0, lg);
      ld8f(p.sums + c0, lb);
      ld8f(p.dgamma + c0, dg);
      ld8f(p.dbeta + c0, db);
#pragma unroll
      for (int i = 0; i < 8; ++i) { p.dgamma[c0 + i] = dg[i] + lg[i]; p.dbeta[c0 + i] = db[i] + lb[i]; }
    }
  }
  int st = 0;
  long long kk = k0;
  while (kk < p.rows) {
#pragma unroll
    for (int j = 0; j < D - 1; ++j) {   // unrolled so the prefetched mask bytes keep fixed register names
      if (kk >= p.rows) break;
      uint32_t unused;
      issue(st == 0 ? D - 1 : st - 1, k_issue, unused, false);
      cp_async_wait<D - 1>();
      const long long r = p.rows - 1 - kk;
      float y[8], d[8], q[8], o[8];
      unpack8f(ring[(st * A + 0) * kBnThreads + tid], y);
      unpack8f(ring[(st * A + 1) * kBnThreads + tid], d);
      if (MODE == BWD_RES) unpack8f(ring[(st * A + A - 1) * kBnThreads + tid], q);
      bwd_dz<ACT, MODE>(d, y, q, mq[j], scale, shift);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(d[i], scale[i], fmaf(y[i], ca[i], cb[i]));
      if (p.dresidual) store8(p.dresidual + r * p.ldr + c0, d);
      store8(p.dy + r * p.lddy + c0, o);
      if (MODE == BWD_MASK) mq[j] = (k_issue < p.rows) ? (uint32_t)__ldg(mrow + (p.rows - 1 - k_issue) * mpitch) : 0u;
      k_issue += rstride;
      st = (st + 1 == D) ? 0 : st + 1;
      kk += rstride;
    }
  }
  cp_async_wait<0>();
}

// Plain activation backward when there is no BN (conv + bias + act): dz = dout * act'(z) with z the stored output.
// ------------------------------------------------------------------------------------------------
// Max-pool 3x3 / stride 2 / pad 1 (NHWC), forward stores the argmax tap (0..8) for an atomic-free backward.
// ------------------------------------------------------------------------------------------------
// Index math is 32-bit: one CTA-row loop over (n, output row) and an inner loop over (column, channel vector), so
// the hot loops contain no 64-bit divisions (they made the first version instruction-bound, ~5x off the HBM time).
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                   uint8_t* __restrict__ argmax, int N, int H, int W, int C, int P, int Q, int k,
                                   int stride, int pad) {
  const int cvs = C / VEC;
  const int per_row = Q * cvs;
  for (int row = blockIdx.x; row < N * P; row += gridDim.x) {
    const int n = row / P, ph = row - n * P;
    const int h0 = ph * stride - pad;
    const __nv_bfloat16* xn = x + (long long)n * H * W * C;
    const long long orow = (long long)row * Q * C;
    for (int it = threadIdx.x; it < per_row; it += blockDim.x) {
      const int q = it / cvs, cv = it - q * cvs;
      const int w0 = q * stride - pad;
      float best[8]; int arg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; arg[i] = 0; }
      for (int r = 0; r < k; ++r) {
        const int h = h0 + r;
        if (h < 0 || h >= H) continue;
        for (int s = 0; s < k; ++s) {
          const int w = w0 + s;
          if (w < 0 || w >= W) continue;
          float v[8];
          load8(xn + ((long long)(h * W + w)) * C + cv * VEC, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) if (v[i] > best[i]) { best[i] = v[i]; arg[i] = r * k + s; }
        }
      }
      const long long o = orow + (long long)q * C + cv * VEC;
      store8(out + o, best);
      if (argmax) {
        uint2 packed;
        packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        *reinterpret_cast<uint2*>(argmax + o) = packed;
      }
    }
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ argmax,
                                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int P, int Q, int k,
                                   int stride, int pad) {
  // gather form: each input pixel looks at the (<= ceil(k/stride)^2) windows that contain it
  const int cvs = C / VEC;
  const int per_row = W * cvs;
  for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
    const int n = row / H, h = row - n * H;
    const int ph_lo = max(0, (h + pad - k + stride) / stride);
    const long long obase = (long long)n * P * Q * C;
    for (int it = threadIdx.x; it < per_row; it += blockDim.x) {
      const int w = it / cvs, cv = it - w * cvs;
      const int q_lo = max(0, (w + pad - k + stride) / stride);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int ph = ph_lo; ph < P && ph * stride - pad <= h; ++ph) {
        const int r = h - (ph * stride - pad);
        if (r < 0 || r >= k) continue;
        for (int q = q_lo; q < Q && q * stride - pad <= w; ++q) {
          const int s = w - (q * stride - pad);
          if (s < 0 || s >= k) continue;
          const long long o = obase + ((long long)(ph * Q + q)) * C + cv * VEC;
          const uint2 packed = *reinterpret_cast<const uint2*>(argmax + o);
          float d[8];
          load8(dout + o, d);
          const int tap = r * k + s;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int a = ((i < 4 ? packed.x : packed.y) >> (8 * (i & 3))) & 0xff;
            if (a == tap) acc[i] += d[i];
          }
        }
      }
      store8(dx + ((long long)row * W + w) * C + cv * VEC, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stem tail: BN + ReLU + max-pool 3x3 / 2 / pad 1 fused (BnPoolParams, elementwise.h).
//   forward   out[p,q] = max(0, max over the window of (y * scale + shift)); arg = first tap that holds it, 9 = none > 0.
//             One read of the conv output, one write of the pooled map: the normalised 112x112 map is never stored.
//   backward  every thread owns a 2x2 block of y: exactly the four windows (bp+{0,1}, bq+{0,1}) reach it, through nine
//             (window, element) pairs known at compile time -- dz is rebuilt from the pooled gradient and `arg` with one
//             16-byte + one 8-byte load per window, no atomics and no scatter.  Pass 1 accumulates sum(dz), sum(dz*xhat)
//             (and opens the SyncBN exchange at its tail), pass 2 writes dy.
// grid: (channel-vector groups, pixel chunks); block: (cvx, 256 / cvx) as the other BN kernels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(BnPoolParams p) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const bool active = c0 < p.C;
  if (p.peer.world > 1 && p.training) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
  if (!active) return;
  float scale[8], shift[8];
  {
    float g[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = 1.f; b[i] = 0.f; }
    if (p.gamma) ld8f(p.gamma + c0, g);
    if (p.beta) ld8f(p.beta + c0, b);
    if (p.training) {
      float s0[8], s1[8];
      gather_stats(p.peer, p.stats, p.sym_offset, p.C, c0, s0, s1);
      const float inv_n = 1.f / p.count;
      const bool writer = (blockIdx.y == 0 && threadIdx.y == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mean = s0[i] * inv_n;
        const float var = fmaxf(s1[i] * inv_n - mean * mean, 0.f);
        const float invstd = rsqrtf(var + p.eps);
        scale[i] = g[i] * invstd;
        shift[i] = b[i] - mean * scale[i];
        if (writer) {
          p.save_mean[c0 + i] = mean;
          p.save_invstd[c0 + i] = invstd;
          if (p.running_mean) {
            const float unbiased = var * (p.count / fmaxf(p.count - 1.f, 1.f));
            p.running_mean[c0 + i] = (1.f - p.momentum) * p.running_mean[c0 + i] + p.momentum * mean;
            p.running_var[c0 + i] = (1.f - p.momentum) * p.running_var[c0 + i] + p.momentum * unbiased;
          }
        }
      }
    } else {
      float rm[8], rv[8];
      ld8f(p.running_mean + c0, rm);
      ld8f(p.running_var + c0, rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        scale[i] = g[i] * rsqrtf(rv[i] + p.eps);
        shift[i] = b[i] - rm[i] * scale[i];
      }
    }
  }
  const int total = p.N * p.P * p.Q;
  const int pstride = gridDim.y * blockDim.y;
  for (int px = blockIdx.y * blockDim.y + threadIdx.y; px < total; px += pstride) {
    const int q = px % p.Q, t = px / p.Q;
    const int ph = t % p.P, n = t / p.P;
    const int h0 = 2 * ph - 1, w0 = 2 * q - 1;
    const __nv_bfloat16* base = p.y + (long long)n * p.H * p.W * p.C + c0;
    uint4 raw[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int h = h0 + r, w = w0 + s2;
        ok[r * 3 + s2] = (h >= 0 && h < p.H && w >= 0 && w < p.W);
        raw[r * 3 + s2] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[r * 3 + s2]) raw[r * 3 + s2] = ldg_cached(base + (long long)(h * p.W + w) * p.C);
      }
    }
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = 0.f; arg[i] = 9u; }   // ReLU floor: only a positive tap can win
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (!ok[tap]) continue;
      float v[8];
      unpack8f(raw[tap], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaf(v[i], scale[i], shift[i]);
        const bool gt = z > best[i];
        best[i] = gt ? z : best[i];
        arg[i] = gt ? (uint32_t)tap : arg[i];
      }
    }
    const long long o = (long long)px * p.C + c0;
    store8(p.out + o, best);
    if (p.arg) {
      uint2 packed;
      packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(p.arg + o) = packed;
    }
  }
}

// dz of the 2x2 block (rows 2bp, 2bp+1; columns 2bq, 2bq+1) of image n, channels c0..c0+7: d[br * 2 + bc][i]
__device__ __forceinline__ void pool_bwd_block(const BnPoolParams& p, int n, int bp, int bq, int c0, float (&d)[4][8]) {
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int i = 0; i < 8; ++i) d[e][i] = 0.f;
#pragma unroll
  for (int dp = 0; dp < 2; ++dp) {
#pragma unroll
    for (int dq = 0; dq < 2; ++dq) {
      const int ph = bp + dp, q = bq + dq;
      if (ph >= p.P || q >= p.Q) continue;
      const long long o = (((long long)n * p.P + ph) * p.Q + q) * p.C + c0;
      const uint2 a = *reinterpret_cast<const uint2*>(p.arg + o);
      float g[8];
      load8(p.dout + o, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int tap = (int)(((i < 4 ? a.x : a.y) >> (8 * (i & 3))) & 0xffu);
        // window row ph covers input rows 2ph-1+r: dp=0 -> r=1 is block row 0, r=2 is block row 1; dp=1 -> r=0 is block row 1
#pragma unroll
        for (int br = 0; br < 2; ++br) {
          const int r = (dp == 0) ? br + 1 : (br == 1 ? 0 : -1);
          if (r < 0) continue;
#pragma unroll
          for (int bc = 0; bc < 2; ++bc) {
            const int s2 = (dq == 0) ? bc + 1 : (bc == 1 ? 0 : -1);
            if (s2 < 0) continue;
            d[br * 2 + bc][i] += (tap == r * 3 + s2) ? g[i] : 0.f;   // select, not a (divergent) branch
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) bn_relu_pool_bwd_reduce_kernel(BnPoolParams p) {
  extern __shared__ __align__(16) float dyn[];   // [16][256] for the CTA-level reduction
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  if (c0 < p.C) {
    float invstd[8], nmi[8];
    {
      float mean[8];
      ld8f(p.save_mean + c0, mean);
      ld8f(p.save_invstd + c0, invstd);
#pragma unroll
      for (int i = 0; i < 8; ++i) nmi[i] = -mean[i] * invstd[i];
    }
    const int BP = p.H / 2, BQ = p.W / 2;
    const int total = p.N * BP * BQ;
    const int pstride = gridDim.y * blockDim.y;
    for (int bx = blockIdx.y * blockDim.y + threadIdx.y; bx < total; bx += pstride) {
      const int bq = bx % BQ, t = bx / BQ;
      const int bp = t % BP, n = t / BP;
      const __nv_bfloat16* yb = p.y + (((long long)n * p.H + 2 * bp) * p.W + 2 * bq) * p.C + c0;
      uint4 raw[4];
      raw[0] = ldg_stream(yb);
      raw[1] = ldg_stream(yb + p.C);
      raw[2] = ldg_stream(yb + (long long)p.W * p.C);
      raw[3] = ldg_stream(yb + (long long)p.W * p.C + p.C);
      float d[4][8];
      pool_bwd_block(p, n, bp, bq, c0, d);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float y[8];
        unpack8f(raw[e], y);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] += d[e][i];
          b[i] = fmaf(d[e][i], fmaf(y[i], invstd[i], nmi[i]), b[i]);
        }
      }
    }
  }
  // CTA-level reduction over threadIdx.y, then one vector RED per 4 channels (same scheme as bn_bwd_reduce_kernel)
#pragma unroll
  for (int i = 0; i < 8; ++i) { dyn[i * kBnThreads + tid] = a[i]; dyn[(8 + i) * kBnThreads + tid] = b[i]; }
  __syncthreads();
  const int bdx = blockDim.x, bdy = blockDim.y;
  if (tid < 4 * bdx) {
    const int quad = tid / bdx, tx = tid - quad * bdx;
    const int c = (blockIdx.x * bdx + tx) * VEC;
    if (c < p.C) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* src = dyn + (quad * 4) * kBnThreads + tx;
      for (int yy = 0; yy < bdy; ++yy) {
        s0 += src[yy * bdx];
        s1 += src[kBnThreads + yy * bdx];
        s2 += src[2 * kBnThreads + yy * bdx];
        s3 += src[3 * kBnThreads + yy * bdx];
      }
      red_add_v4(p.stats + (quad >= 2 ? p.C : 0) + c + (quad & 1) * 4, s0, s1, s2, s3);
    }
  }
  if (p.peer.world > 1) peer_signal_at_tail(p.peer, gridDim.x * gridDim.y);
}

__global__ void __launch_bounds__(256) bn_relu_pool_bwd_apply_kernel(BnPoolParams p) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const bool active = c0 < p.C;
  if (p.peer.world > 1) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
  if (!active) return;
  // dy = (dz - mean(dz) - xhat * mean(dz * xhat)) * gamma * invstd  ==  dz * scale + y * ca + cb
  float scale[8], ca[8], cb[8];
  {
    float s0[8], s1[8];
    gather_stats(p.peer, p.stats, p.sym_offset, p.C, c0, s0, s1);
    const float inv_n = 1.f / p.count;
    float mean[8], invstd[8], g[8];
    ld8f(p.save_mean + c0, mean);
    ld8f(p.save_invstd + c0, invstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 1.f;
    if (p.gamma) ld8f(p.gamma + c0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m_dz = s0[i] * inv_n, m_dzx = s1[i] * inv_n;
      scale[i] = g[i] * invstd[i];
      ca[i] = -invstd[i] * m_dzx * scale[i];
      cb[i] = -m_dz * scale[i] - mean[i] * ca[i];
    }
    if (blockIdx.y == 0 && threadIdx.y == 0 && p.dgamma) {
      // parameter gradients use the LOCAL sums (the gradient all-reduce averages them afterwards, as DDP does)
      float lg[8], lb[8], dg[8], db[8];
      ld8f(p.stats + p.C + c0, lg);
      ld8f(p.stats + c0, lb);
      ld8f(p.dgamma + c0, dg);
      ld8f(p.dbeta + c0, db);
#pragma unroll
      for (int i = 0; i < 8; ++i) { p.dgamma[c0 + i] = dg[i] + lg[i]; p.dbeta[c0 + i] = db[i] + lb[i]; }
    }
  }
  const int BP = p.H / 2, BQ = p.W / 2;
  const int total = p.N * BP * BQ;
  const int pstride = gridDim.y * blockDim.y;
  for (int bx = blockIdx.y * blockDim.y + threadIdx.y; bx < total; bx += pstride) {
    const int bq = bx % BQ, t = bx / BQ;
    const int bp = t % BP, n = t / BP;
    const long long e0 = (((long long)n * p.H + 2 * bp) * p.W + 2 * bq) * p.C + c0;
    const long long rowp = (long long)p.W * p.C;
    uint4 raw[4];
    raw[0] = ldg_stream(p.y + e0);
    raw[1] = ldg_stream(p.y + e0 + p.C);
    raw[2] = ldg_stream(p.y + e0 + rowp);
    raw[3] = ldg_stream(p.y + e0 + rowp + p.C);
    float d[4][8];
    pool_bwd_block(p, n, bp, bq, c0, d);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float y[8], o[8];
      unpack8f(raw[e], y);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(d[e][i], scale[i], fmaf(y[i], ca[i], cb[i]));
      store8(p.dy + e0 + (e >> 1) * rowp + (e & 1) * p.C, o);
    }
  }
}

// ---- strip versions of the stem tail: whole image rows are brought into shared memory with linear bulk copies
// (cp.async.bulk + mbarrier, two stages), so the loads of the next strip are in flight while this one is computed.
// The register versions above (loads consumed right after they are issued, ~2 loads in flight per thread) reached only
// 1.3-2.9 TB/s on the 112x112x64 stem; they remain the fallback when a strip does not fit in shared memory.
// block: (cvs = C/8, by); thread (tx, ty) owns channel vector tx and walks the strip's pixels / 2x2 blocks ty, ty+by, ...
__device__ __forceinline__ void unpack8_smem(const unsigned char* base, float (&f)[8]) {
  unpack8f(*reinterpret_cast<const uint4*>(base), f);
}

__global__ void __launch_bounds__(256) bn_relu_pool_fwd_strip_kernel(BnPoolParams p) {
  extern __shared__ __align__(128) unsigned char strip[];    // [2][3 rows][W][C] bf16
  __shared__ __align__(8) uint64_t full[2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * blockDim.x + tx;
  const int c0 = tx * VEC;
  const uint32_t row_bytes = (uint32_t)p.W * p.C * 2u;
  const uint32_t stage_bytes = 3u * row_bytes;
  const int strips = p.N * p.P;
  auto issue = [&](int s, int stage) {        // thread 0 only
    const int n = s / p.P, ph = s - n * p.P;
    const int h_lo = max(2 * ph - 1, 0);                    // rows h_lo .. 2ph+1 are contiguous in memory
    const uint32_t bytes = (uint32_t)(2 * ph + 2 - h_lo) * row_bytes;
    const uint32_t bar = smem_u32(&full[stage]);
    mbar_expect_tx(bar, bytes);
    bulk_load_1d(smem_u32(strip + (size_t)stage * stage_bytes + (size_t)(h_lo - (2 * ph - 1)) * row_bytes),
                 p.y + ((long long)n * p.H + h_lo) * p.W * p.C, bytes, bar);
  };
  if (tid == 0) { mbar_init(smem_u32(&full[0]), 1); mbar_init(smem_u32(&full[1]), 1); fence_barrier_init(); }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < strips) issue(blockIdx.x, 0);
  if (p.peer.world > 1 && p.training) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
  float scale[8], shift[8];
  {
    float g[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = 1.f; b[i] = 0.f; }
    if (p.gamma) ld8f(p.gamma + c0, g);
    if (p.beta) ld8f(p.beta + c0, b);
    if (p.training) {
      float s0[8], s1[8];
      gather_stats(p.peer, p.stats, p.sym_offset, p.C, c0, s0, s1);
      const float inv_n = 1.f / p.count;
      const bool writer = (blockIdx.x == 0 && ty == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mean = s0[i] * inv_n;
        const float var = fmaxf(s1[i] * inv_n - mean * mean, 0.f);
        const float invstd = rsqrtf(var + p.eps);
        scale[i] = g[i] * invstd;
        shift[i] = b[i] - mean * scale[i];
        if (writer) {
          p.save_mean[c0 + i] = mean;
          p.save_invstd[c0 + i] = invstd;
          if (p.running_mean) {
            const float unbiased = var * (p.count / fmaxf(p.count - 1.f, 1.f));
            p.running_mean[c0 + i] = (1.f - p.momentum) * p.running_mean[c0 + i] + p.momentum * mean;
            p.running_var[c0 + i] = (1.f - p.momentum) * p.running_var[c0 + i] + p.momentum * unbiased;
          }
        }
      }
    } else {
      float rm[8], rv[8];
      ld8f(p.running_mean + c0, rm);
      ld8f(p.running_var + c0, rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        scale[i] = g[i] * rsqrtf(rv[i] + p.eps);
        shift[i] = b[i] - rm[i] * scale[i];
      }
    }
  }
  int it = 0;
  for (int s = blockIdx.x; s < strips; s += gridDim.x, ++it) {
    const int stage = it & 1;
    if (tid == 0 && s + (int)gridDim.x < strips) issue(s + gridDim.x, stage ^ 1);   // that stage was drained in iteration it-1
    mbar_wait(smem_u32(&full[stage]), (uint32_t)((it >> 1) & 1));
    const int ph = s % p.P;
    const unsigned char* base = strip + (size_t)stage * stage_bytes + (size_t)c0 * 2;
    const int r_lo = (ph == 0) ? 1 : 0;                    // row 2ph-1 does not exist for the first pooled row
#pragma unroll 2
    for (int q = ty; q < p.Q; q += blockDim.y) {
      float best[8];
      uint32_t arg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { best[i] = 0.f; arg[i] = 9u; }   // ReLU floor: only a positive tap can win
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (r < r_lo) continue;
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const int w = 2 * q - 1 + s2;
          if (w < 0 || w >= p.W) continue;
          float v[8];
          unpack8_smem(base + ((size_t)r * p.W + w) * p.C * 2, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = fmaf(v[i], scale[i], shift[i]);
            const bool gt = z > best[i];
            best[i] = gt ? z : best[i];
            arg[i] = gt ? (uint32_t)(r * 3 + s2) : arg[i];
          }
        }
      }
      const long long o = ((long long)s * p.Q + q) * p.C + c0;
      store8(p.out + o, best);
      if (p.arg) {
        uint2 packed;
        packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        *reinterpret_cast<uint2*>(p.arg + o) = packed;
      }
    }
    __syncthreads();      // everyone is done with this stage before thread 0 refills it in the next iteration
  }
}

// strip = block row bp of image n: y rows 2bp, 2bp+1 | dout rows bp, bp+1 | arg rows bp, bp+1 (the +1 rows absent at the end)
template <bool APPLY>
__global__ void __launch_bounds__(256) bn_relu_pool_bwd_strip_kernel(BnPoolParams p) {
  extern __shared__ __align__(128) unsigned char strip[];    // [2][ 2*W*C*2 | 2*Q*C*2 | 2*Q*C ]
  __shared__ __align__(8) uint64_t full[2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * blockDim.x + tx;
  const int c0 = tx * VEC;
  const int BP = p.H / 2, BQ = p.W / 2;
  const uint32_t y_bytes = 2u * p.W * p.C * 2u, d_row = (uint32_t)p.Q * p.C * 2u, a_row = (uint32_t)p.Q * p.C;
  const uint32_t stage_bytes = y_bytes + 2u * d_row + 2u * a_row;
  const int strips = p.N * BP;
  auto issue = [&](int s, int stage) {        // thread 0 only
    const int n = s / BP, bp = s - n * BP;
    const int rows = (bp + 1 < p.P) ? 2 : 1;
    const uint32_t bar = smem_u32(&full[stage]);
    unsigned char* st = strip + (size_t)stage * stage_bytes;
    mbar_expect_tx(bar, y_bytes + rows * (d_row + a_row));
    bulk_load_1d(smem_u32(st), p.y + ((long long)n * p.H + 2 * bp) * p.W * p.C, y_bytes, bar);
    const long long o = ((long long)n * p.P + bp) * p.Q * p.C;
    bulk_load_1d(smem_u32(st + y_bytes), p.dout + o, rows * d_row, bar);
    bulk_load_1d(smem_u32(st + y_bytes + 2 * d_row), p.arg + o, rows * a_row, bar);
  };
  if (tid == 0) { mbar_init(smem_u32(&full[0]), 1); mbar_init(smem_u32(&full[1]), 1); fence_barrier_init(); }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < strips) issue(blockIdx.x, 0);
  float ka[8], kb[8], kc[8];   // reduce: invstd, -mean*invstd, -   apply: scale, ca, cb
  if (APPLY) {
    if (p.peer.world > 1) peer_exchange_reduce(p.peer, p.sym_offset, p.C);
    float s0[8], s1[8];
    gather_stats(p.peer, p.stats, p.sym_offset, p.C, c0, s0, s1);
    const float inv_n = 1.f / p.count;
    float mean[8], invstd[8], g[8];
    ld8f(p.save_mean + c0, mean);
    ld8f(p.save_invstd + c0, invstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 1.f;
    if (p.gamma) ld8f(p.gamma + c0, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m_dz = s0[i] * inv_n, m_dzx = s1[i] * inv_n;
      ka[i] = g[i] * invstd[i];
      kb[i] = -invstd[i] * m_dzx * ka[i];
      kc[i] = -m_dz * ka[i] - mean[i] * kb[i];
    }
    if (blockIdx.x == 0 && ty == 0 && p.dgamma) {
      float lg[8], lb[8], dg[8], db[8];
      ld8f(p.stats + p.C + c0, lg);
      ld8f(p.stats + c0, lb);
      ld8f(p.dgamma + c0, dg);
      ld8f(p.dbeta + c0, db);
#pragma unroll
      for (int i = 0; i < 8; ++i) { p.dgamma[c0 + i] = dg[i] + lg[i]; p.dbeta[c0 + i] = db[i] + lb[i]; }
    }
  } else {
    float mean[8];
    ld8f(p.save_mean + c0, mean);
    ld8f(p.save_invstd + c0, ka);
#pragma unroll
    for (int i = 0; i < 8; ++i) { kb[i] = -mean[i] * ka[i]; kc[i] = 0.f; }
  }
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  int it = 0;
  for (int s = blockIdx.x; s < strips; s += gridDim.x, ++it) {
    const int stage = it & 1;
    if (tid == 0 && s + (int)gridDim.x < strips) issue(s + gridDim.x, stage ^ 1);
    mbar_wait(smem_u32(&full[stage]), (uint32_t)((it >> 1) & 1));
    const int bp = s % BP;
    const bool row2 = bp + 1 < p.P;
    const unsigned char* ys = strip + (size_t)stage * stage_bytes + (size_t)c0 * 2;
    const unsigned char* ds = strip + (size_t)stage * stage_bytes + y_bytes + (size_t)c0 * 2;
    const unsigned char* as = strip + (size_t)stage * stage_bytes + y_bytes + 2 * d_row + c0;
#pragma unroll 2
    for (int bq = ty; bq < BQ; bq += blockDim.y) {
      float d[4][8];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int i = 0; i < 8; ++i) d[e][i] = 0.f;
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
#pragma unroll
        for (int dq = 0; dq < 2; ++dq) {
          const int q = bq + dq;
          if ((dp == 1 && !row2) || q >= p.Q) continue;
          const size_t o = ((size_t)dp * p.Q + q) * p.C;
          const uint2 av = *reinterpret_cast<const uint2*>(as + o);
          float g[8];
          unpack8_smem(ds + o * 2, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int tap = (int)(((i < 4 ? av.x : av.y) >> (8 * (i & 3))) & 0xffu);
#pragma unroll
            for (int br = 0; br < 2; ++br) {
              const int r = (dp == 0) ? br + 1 : (br == 1 ? 0 : -1);
              if (r < 0) continue;
#pragma unroll
              for (int bc = 0; bc < 2; ++bc) {
                const int s2 = (dq == 0) ? bc + 1 : (bc == 1 ? 0 : -1);
                if (s2 < 0) continue;
                d[br * 2 + bc][i] += (tap == r * 3 + s2) ? g[i] : 0.f;   // select, not a (divergent) branch
              }
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float y[8];
        const size_t yo = ((size_t)(e >> 1) * p.W + 2 * bq + (e & 1)) * p.C;
        unpack8_smem(ys + yo * 2, y);
        if (APPLY) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(d[e][i], ka[i], fmaf(y[i], kb[i], kc[i]));
          store8(p.dy + ((long long)s * 2 * p.W) * p.C + yo + c0, o);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a[i] += d[e][i];
            b[i] = fmaf(d[e][i], fmaf(y[i], ka[i], kb[i]), b[i]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!APPLY) {
    // CTA reduction over ty in the (now idle) strip memory, then one vector RED per 4 channels
    float* red = reinterpret_cast<float*>(strip);            // [by][cvs][16]
    const int cvs = blockDim.x, by = blockDim.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[(ty * cvs + tx) * 16 + i] = a[i]; red[(ty * cvs + tx) * 16 + 8 + i] = b[i]; }
    __syncthreads();
    for (int j = tid; j < cvs * 4; j += cvs * by) {          // j = (cv, quad): quad 0,1 -> sum(dz) halves; 2,3 -> sum(dz*xhat)
      const int cv = j >> 2, quad = j & 3;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int yy = 0; yy < by; ++yy) {
        const float* src = red + (yy * cvs + cv) * 16 + quad * 4;
        s0 += src[0]; s1 += src[1]; s2 += src[2]; s3 += src[3];
      }
      red_add_v4(p.stats + (quad >= 2 ? p.C : 0) + cv * VEC + (quad & 1) * 4, s0, s1, s2, s3);
    }
    if (p.peer.world > 1) peer_signal_at_tail(p.peer, gridDim.x * gridDim.y);
  }
}

// Global average pool [N][HW][C] -> [N][C] and its backward (broadcast / HW).
// CTA = (32 channel vectors) x (8 pixel lanes) of ONE sample: the pixel loop is split over threadIdx.y (4 loads in flight
// per thread) and reduced through shared memory.  The first version had one thread walk all HW pixels of a channel
// vector -- with the squeeze-excite layers of RegNetY / EfficientNet (C = 32..224 at 112^2..56^2, batch 64) that was a
// few thousand threads in total and 6.8 ms of a 20 ms EfficientNet-B0 step.
__global__ void __launch_bounds__(256) gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int HW, int C) {
  const int cv = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.y;
  const bool active = cv * VEC < C;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (active) {
    const __nv_bfloat16* base = x + (long long)n * HW * C + cv * VEC;
    int t = threadIdx.y;
    for (; t + 24 < HW; t += 32) {
      uint4 r0 = ldg_stream(base + (long long)t * C), r1 = ldg_stream(base + (long long)(t + 8) * C);
      uint4 r2 = ldg_stream(base + (long long)(t + 16) * C), r3 = ldg_stream(base + (long long)(t + 24) * C);
      float v[8];
      unpack8f(r0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
      unpack8f(r1, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
      unpack8f(r2, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
      unpack8f(r3, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
    for (; t < HW; t += 8) {
      float v[8];
      load8(base + (long long)t * C, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  __shared__ float sm[8][32][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && active) {
    const float inv = 1.f / HW;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s_ = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s_ += sm[y][threadIdx.x][i];
      o[i] = s_ * inv;
    }
    store8(out + (long long)n * C + cv * VEC, o);
  }
}
// I = index type: 32-bit whenever the element count allows it (64-bit divisions cost ~100 instructions each)
template <typename I>
__global__ void gap_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int N, int HW, int C) {
  const I cvs = (I)(C / VEC);
  const I total = (I)N * (I)HW * cvs;
  const float inv = 1.f / HW;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvs);
    const int n = (int)((idx / cvs) / (I)HW);
    float v[8];
    load8(dout + (long long)n * C + cv * VEC, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= inv;
    store8(dx + (long long)idx * VEC, v);
  }
}

// 2x2 / stride-2 average pool (DenseNet transitions, BoTNet stride-2 block).
template <typename I>
__global__ void avgpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W, int C) {
  const int P = H / 2, Q = W / 2;
  const I cvs = (I)(C / VEC);
  const I total = (I)N * (I)P * (I)Q * cvs;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvs);
    I pix = idx / cvs;
    const int q = (int)(pix % (I)Q); pix /= (I)Q;
    const int ph = (int)(pix % (I)P); const int n = (int)(pix / (I)P);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        float v[8];
        load8(x + (((long long)n * H + 2 * ph + r) * W + 2 * q + s) * C + cv * VEC, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += 0.25f * v[i];
      }
    store8(out + (long long)idx * VEC, acc);
  }
}
template <typename I>
__global__ void avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
  const int P = H / 2, Q = W / 2;
  const I cvs = (I)(C / VEC);
  const I total = (I)N * (I)H * (I)W * cvs;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvs);
    I pix = idx / cvs;
    const int w = (int)(pix % (I)W); pix /= (I)W;
    const int h = (int)(pix % (I)H); const int n = (int)(pix / (I)H);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (h / 2 < P && w / 2 < Q) {
      load8(dout + (((long long)n * P + h / 2) * Q + w / 2) * C + cv * VEC, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= 0.25f;
    }
    store8(dx + (long long)idx * VEC, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excite gating (RegNetY / EfficientNet): out[n,hw,c] = x[n,hw,c] * gate[n,c].
// Backward in ONE pass over (dout, x): dx = dout * gate and dgate[n,c] = sum_hw dout * x (fp32 atomics).
// grid: (channel-vector blocks, hw chunks, N)
// ------------------------------------------------------------------------------------------------
__global__ void channel_scale_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gate,
                                         __nv_bfloat16* __restrict__ out, int HW, int C) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  if (c0 >= C) return;
  const int n = blockIdx.z;
  float g[8];
  load8(gate + (long long)n * C + c0, g);
  for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < HW; t += gridDim.y * blockDim.y) {
    const long long o = ((long long)n * HW + t) * C + c0;
    float v[8];
    load8(x + o, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= g[i];
    store8(out + o, v);
  }
}
__global__ void __launch_bounds__(256) channel_scale_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ dx,
                                         float* __restrict__ dgate, int HW, int C) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * VEC;
  const bool active = c0 < C;
  const int n = blockIdx.z;
  float g[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = 0.f; g[i] = 0.f; }
  if (active) {
    load8(gate + (long long)n * C + c0, g);
    for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < HW; t += gridDim.y * blockDim.y) {
      const long long o = ((long long)n * HW + t) * C + c0;
      float d[8], v[8];
      load8(dout + o, d);
      load8(x + o, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] = fmaf(d[i], v[i], acc[i]); d[i] *= g[i]; }
      store8(dx + o, d);
    }
  }
  // reduce over threadIdx.y in shared memory, then ONE atomic per channel and CTA (the per-thread atomics of the first
  // version -- 16 M per launch on RegNetY-160's stage 3 -- were the whole cost of this kernel)
  __shared__ float sm[256 * 9];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[tid * 9 + i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s_ = 0.f;
      for (int y = 0; y < (int)blockDim.y; ++y) s_ += sm[(y * blockDim.x + threadIdx.x) * 9 + i];
      atomicAdd(dgate + (long long)n * C + c0 + i, s_);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused softmax cross-entropy + top-1/top-k hit counting + dlogits, one CTA per sample.
// loss_sum / hits are accumulated atomically; dlogits = (softmax - onehot) * grad_scale.
// ------------------------------------------------------------------------------------------------
__global__ void ce_topk_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ target,
                               __nv_bfloat16* __restrict__ dlogits, float* __restrict__ accum /*[3]*/, int ncls,
                               long long ld, int topk, float grad_scale) {
  const int row = blockIdx.x;
  const __nv_bfloat16* lg = logits + (long long)row * ld;
  const int tgt = (int)target[row];
  __shared__ float s_red[32];
  __shared__ float s_bcast[2];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) mx = fmaxf(mx, __bfloat162float(lg[c]));
  // block max
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) s_bcast[0] = v;
  }
  __syncthreads();
  mx = s_bcast[0];
  const float tv = __bfloat162float(lg[tgt]);
  float sum = 0.f, rank_gt = 0.f;  // rank_gt: number of classes that beat the target (ties broken by lower index)
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
    const float v = __bfloat162float(lg[c]);
    sum += __expf(v - mx);
    rank_gt += (v > tv || (v == tv && c < tgt)) ? 1.f : 0.f;
  }
  sum = warp_sum(sum); rank_gt = warp_sum(rank_gt);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = sum; s_red[16 + (threadIdx.x >> 5)] = rank_gt; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    float a = threadIdx.x < nw ? s_red[threadIdx.x] : 0.f;
    float b = threadIdx.x < nw ? s_red[16 + threadIdx.x] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) { s_bcast[0] = a; s_bcast[1] = b; }
  }
  __syncthreads();
  sum = s_bcast[0]; rank_gt = s_bcast[1];
  const float lse = mx + __logf(sum);
  if (threadIdx.x == 0) {
    atomicAdd(accum + 0, lse - tv);
    if (rank_gt < 0.5f) atomicAdd(accum + 1, 1.f);
    if (rank_gt < (float)topk - 0.5f) atomicAdd(accum + 2, 1.f);
  }
  if (dlogits) {
    __nv_bfloat16* dl = dlogits + (long long)row * ld;
    const float inv = 1.f / sum;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
      float pr = __expf(__bfloat162float(lg[c]) - mx) * inv;
      if (c == tgt) pr -= 1.f;
      dl[c] = __float2bfloat16_rn(pr * grad_scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Input staging: NCHW fp32 (what the data loader / reference produces) -> NHWC bf16, and the explicit im2col of
// the 7x7/2 stem (Cin=3 makes K=147: the GEMM runs on [pixels][160] patches, columns 147..159 are zero).
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, int H, int W) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % C;
    long long pix = idx / C;
    const int w = pix % W; pix /= W;
    const int h = pix % H; const int n = pix / H;
    out[idx] = __float2bfloat16_rn(x[(((long long)n * C + c) * H + h) * W + w]);
  }
}

// T = float (already normalised images) or uint8_t (raw pixels: v * scale[c] + bias[c] is applied while the rows are
// staged, so real-data batches cross PCIe as bytes and are never materialised in fp32 -- SURVEY G18).
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const T* __restrict__ x /*NCHW*/, __nv_bfloat16* __restrict__ patches,
                                   int N, int C, int H, int W, int P, int Q, int R, int S, int stride, int pad, int Kpad,
                                   StemNorm norm) {
  // One CTA per output row (n, p): the R input rows it needs are staged in shared memory once (coalesced reads of the
  // NCHW image, zero-filled borders), then every thread emits 16-byte patch pieces (row layout [r][s][c] + zero
  // padding to Kpad, matching weights [Cout][R][S][C]) with fully coalesced stores.
  extern __shared__ float tile[];                    // [C][R][W + 2*pad]
  const int Wp = W + 2 * pad;
  const int n = blockIdx.x / P, ph = blockIdx.x % P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // one (channel, filter-row) image row per warp iteration: the divisions are per row, not per element
  for (int cr = warp; cr < C * R; cr += nwarps) {
    const int c = cr / R, r = cr - c * R;
    const int h = ph * stride - pad + r;
    const bool row_ok = (h >= 0 && h < H);
    const T* src = x + (((long long)n * C + c) * H + (row_ok ? h : 0)) * W;
    float* dst = tile + cr * Wp;
    const float sc = norm.scale[c & 7], bi = norm.bias[c & 7];
    for (int wp = lane; wp < Wp; wp += 32) {
      const int w = wp - pad;
      float v = 0.f;   // zero padding lives in the normalised domain, as in Normalize -> Conv2d(padding)
      if (row_ok && w >= 0 && w < W) v = sizeof(T) == 1 ? fmaf((float)__ldg(src + w), sc, bi) : (float)__ldg(src + w);
      dst[wp] = v;
    }
  }
  // per-k lookup of the tile offset (c, r, s) -> (c*R + r)*Wp + s, built once per CTA (no div/mod in the hot loop)
  __shared__ int koff[256];
  const int kdim = R * S * C;
  for (int k = threadIdx.x; k < Kpad && k < 256; k += blockDim.x) {
    if (k < kdim) { const int c = k % C, rs = k / C; koff[k] = (c * R + rs / S) * Wp + rs % S; }
    else koff[k] = -1;
  }
  __syncthreads();
  const int vec_per_row = Kpad / 8;
  __nv_bfloat16* out_row = patches + ((long long)n * P + ph) * Q * Kpad;
  // (q, j8) advance incrementally by blockDim.x items per iteration
  int q = threadIdx.x / vec_per_row, j8 = threadIdx.x - q * vec_per_row;
  const int dq = blockDim.x / vec_per_row, dj = blockDim.x - dq * vec_per_row;
  while (q < Q) {
    const int base = q * stride;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int o = koff[j8 * 8 + i];
      v[i] = o >= 0 ? tile[o + base] : 0.f;
    }
    store8(out_row + (long long)q * Kpad + j8 * 8, v);
    q += dq; j8 += dj;
    if (j8 >= vec_per_row) { j8 -= vec_per_row; ++q; }
  }
}

// rows x cols (bf16) -> rows x cols_pad with zero fill; used to give the stem weight a TMA-legal row pitch.
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols, int cols_pad) {
  const int total = rows * cols_pad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % cols_pad, r = idx / cols_pad;
    dst[idx] = c < cols ? src[r * cols + c] : __float2bfloat16_rn(0.f);
  }
}
// fp32 [rows][cols_pad] -> accumulate the first `cols` columns into fp32 [rows][cols]
__global__ void unpad_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int cols_pad) {
  const int total = rows * cols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % cols, r = idx / cols;
    dst[idx] += src[r * cols_pad + c];
  }
}

}  // namespace b200

// ================================================================================================
// host launchers (C linkage, no torch dependency)
// ================================================================================================
using namespace b200;

static inline void bn_launch_dims(int C, long long rows, dim3& grid, dim3& block, int ctas_per_sm = 4) {
  const int cvs = C / VEC;
  int bx = 1;
  while (bx < 32 && bx < cvs) bx <<= 1;     // power of two <= 32 covering the channel vectors
  const int by = 256 / bx;
  const int gx = (cvs + bx - 1) / bx;
  long long want = (rows + by - 1) / by;
  long long cap = (148 * ctas_per_sm + gx - 1) / gx;   // CTAs per SM in total (2 resident: ring smem + registers)
  int gy = (int)(want < cap ? want : cap);
  if (gy < 1) gy = 1;
  grid = dim3(gx, gy);
  block = dim3(bx, by);
}

// The ring buffers plus the few bytes of static shared memory exceed the default 48 KB limit: opt in once.
template <typename Kern>
static inline cudaError_t allow_big_smem(Kern kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

typedef void (*BnApplyKernel)(BnApplyParams);
struct FwdPick { BnApplyKernel kern; int smem; };
template <int ACT, bool RES>
static FwdPick fwd_pick_one() {
  static cudaError_t once = allow_big_smem(bn_apply_kernel<ACT, RES>, FwdRing<RES>::kBytes);
  return FwdPick{once == cudaSuccess ? bn_apply_kernel<ACT, RES> : nullptr, FwdRing<RES>::kBytes};
}
extern "C" int b200_bn_apply(const BnApplyParams* p, cudaStream_t s) {
  const bool res = p->residual != nullptr;
  FwdPick k;
  if (p->act == ACT_RELU) k = res ? fwd_pick_one<ACT_RELU, true>() : fwd_pick_one<ACT_RELU, false>();
  else if (p->act == ACT_SILU) k = res ? fwd_pick_one<ACT_SILU, true>() : fwd_pick_one<ACT_SILU, false>();
  else k = res ? fwd_pick_one<ACT_NONE, true>() : fwd_pick_one<ACT_NONE, false>();
  if (!k.kern) return (int)cudaErrorInvalidValue;
  dim3 g, b;
  bn_launch_dims(p->C, p->rows, g, b);
  k.kern<<<g, b, k.smem, s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_bn_stats(const void* y, long long rows, int C, long long ldy, float* stats, cudaStream_t s) {
  dim3 g, b;
  bn_launch_dims(C, rows, g, b);
  bn_stats_kernel<<<g, b, b.x * b.y * 16 * sizeof(float), s>>>((const __nv_bfloat16*)y, rows, C, ldy, stats);
  return (int)cudaGetLastError();
}
// (activation, how act' is obtained) -> kernel instance
typedef void (*BnBwdKernel)(BnBwdParams);
struct BwdPick { BnBwdKernel reduce, apply; int smem; };
template <int ACT, int MODE>
static BwdPick bwd_pick_one() {
  static cudaError_t once_r = allow_big_smem(bn_bwd_reduce_kernel<ACT, MODE>, BwdRing<MODE>::kBytes);
  static cudaError_t once_a = allow_big_smem(bn_bwd_apply_kernel<ACT, MODE>, BwdRing<MODE>::kBytes);
  if (once_r != cudaSuccess || once_a != cudaSuccess) return BwdPick{nullptr, nullptr, 0};
  return BwdPick{bn_bwd_reduce_kernel<ACT, MODE>, bn_bwd_apply_kernel<ACT, MODE>, BwdRing<MODE>::kBytes};
}
static BwdPick bwd_pick(const BnBwdParams* p) {
  if (p->act == ACT_NONE) return bwd_pick_one<ACT_NONE, BWD_PLAIN>();
  if (p->relu_mask != nullptr && p->act == ACT_RELU) return bwd_pick_one<ACT_RELU, BWD_MASK>();
  const bool res = p->residual != nullptr;
  if (p->act == ACT_RELU) return res ? bwd_pick_one<ACT_RELU, BWD_RES>() : bwd_pick_one<ACT_RELU, BWD_PLAIN>();
  return res ? bwd_pick_one<ACT_SILU, BWD_RES>() : bwd_pick_one<ACT_SILU, BWD_PLAIN>();
}
extern "C" int b200_bn_bwd_reduce(const BnBwdParams* p, cudaStream_t s) {
  const BwdPick k = bwd_pick(p);
  if (!k.reduce) return (int)cudaErrorInvalidValue;
  dim3 g, b;
  bn_launch_dims(p->C, p->rows, g, b, 2);   // exactly the resident wave: fewer CTAs = fewer contended atomics at the end
  k.reduce<<<g, b, k.smem, s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_bn_bwd_apply(const BnBwdParams* p, cudaStream_t s) {
  const BwdPick k = bwd_pick(p);
  if (!k.apply) return (int)cudaErrorInvalidValue;
  dim3 g, b;
  bn_launch_dims(p->C, p->rows, g, b);
  k.apply<<<g, b, k.smem, s>>>(*p);
  return (int)cudaGetLastError();
}
// strip kernels: block (cvs, by) with by chosen so the strip's items split evenly; two stages of the strip in shared memory
static inline bool pool_strip_cfg(int C, int items, size_t stage_bytes, dim3& grid, dim3& block, size_t& smem, int strips) {
  const int cvs = C / VEC;
  if (cvs > 64) return false;
  smem = 2 * stage_bytes;
  if (smem < (size_t)(256 / cvs) * cvs * 16 * sizeof(float)) smem = (size_t)(256 / cvs) * cvs * 16 * sizeof(float);
  if (smem > 200 * 1024) return false;
  int by = 256 / cvs;
  if (by > items) by = items;
  const int rounds = (items + by - 1) / by;
  by = (items + rounds - 1) / rounds;
  block = dim3(cvs, by);
  const int per_sm = smem <= 100 * 1024 ? 2 : 1;
  int g = 148 * per_sm;
  if (g > strips) g = strips;
  grid = dim3(g);
  return true;
}
template <typename Kern>
static inline cudaError_t pool_strip_attr(Kern kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
extern "C" int b200_bn_relu_pool_fwd(const BnPoolParams* p, cudaStream_t s) {
  dim3 g, b;
  size_t smem;
  static cudaError_t once = pool_strip_attr(bn_relu_pool_fwd_strip_kernel);
  if (once == cudaSuccess && pool_strip_cfg(p->C, p->Q, (size_t)3 * p->W * p->C * 2, g, b, smem, p->N * p->P)) {
    bn_relu_pool_fwd_strip_kernel<<<g, b, smem, s>>>(*p);
    return (int)cudaGetLastError();
  }
  bn_launch_dims(p->C, (long long)p->N * p->P * p->Q, g, b, 8);
  bn_relu_pool_fwd_kernel<<<g, b, 0, s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_bn_relu_pool_bwd_reduce(const BnPoolParams* p, cudaStream_t s) {
  dim3 g, b;
  size_t smem;
  static cudaError_t once = pool_strip_attr(bn_relu_pool_bwd_strip_kernel<false>);
  // the arg rows are copied with cp.async.bulk: Q * C bytes per row must be a multiple of 16
  if (once == cudaSuccess && ((p->W / 2) * p->C) % 16 == 0 &&
      pool_strip_cfg(p->C, p->W / 2, (size_t)7 * p->W * p->C, g, b, smem, p->N * (p->H / 2))) {
    bn_relu_pool_bwd_strip_kernel<false><<<g, b, smem, s>>>(*p);
    return (int)cudaGetLastError();
  }
  bn_launch_dims(p->C, (long long)p->N * (p->H / 2) * (p->W / 2), g, b, 4);
  bn_relu_pool_bwd_reduce_kernel<<<g, b, 16 * kBnThreads * sizeof(float), s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_bn_relu_pool_bwd_apply(const BnPoolParams* p, cudaStream_t s) {
  dim3 g, b;
  size_t smem;
  static cudaError_t once = pool_strip_attr(bn_relu_pool_bwd_strip_kernel<true>);
  // the arg rows are copied with cp.async.bulk: Q * C bytes per row must be a multiple of 16
  if (once == cudaSuccess && ((p->W / 2) * p->C) % 16 == 0 &&
      pool_strip_cfg(p->C, p->W / 2, (size_t)7 * p->W * p->C, g, b, smem, p->N * (p->H / 2))) {
    bn_relu_pool_bwd_strip_kernel<true><<<g, b, smem, s>>>(*p);
    return (int)cudaGetLastError();
  }
  bn_launch_dims(p->C, (long long)p->N * (p->H / 2) * (p->W / 2), g, b, 8);
  bn_relu_pool_bwd_apply_kernel<<<g, b, 0, s>>>(*p);
  return (int)cudaGetLastError();
}
// 32-bit index arithmetic is safe when the element count plus one grid stride cannot wrap
static inline bool fits32(long long total) { return total < (1ll << 31); }
static inline int ew_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = 148 * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
extern "C" int b200_maxpool_fwd(const void* x, void* out, void* argmax, int N, int H, int W, int C, int P, int Q, int k,
                                int stride, int pad, cudaStream_t s) {
  maxpool_fwd_kernel<<<ew_grid((long long)N * P * 256, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, (uint8_t*)argmax,
                                                         N, H, W, C, P, Q, k, stride, pad);
  return (int)cudaGetLastError();
}
extern "C" int b200_maxpool_bwd(const void* dout, const void* argmax, void* dx, int N, int H, int W, int C, int P, int Q,
                                int k, int stride, int pad, cudaStream_t s) {
  maxpool_bwd_kernel<<<ew_grid((long long)N * H * 256, 256), 256, 0, s>>>((const __nv_bfloat16*)dout, (const uint8_t*)argmax,
                                                         (__nv_bfloat16*)dx, N, H, W, C, P, Q, k, stride, pad);
  return (int)cudaGetLastError();
}
extern "C" int b200_gap_fwd(const void* x, void* out, int N, int HW, int C, cudaStream_t s) {
  gap_fwd_kernel<<<dim3((C / VEC + 31) / 32, N), dim3(32, 8), 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, HW, C);
  return (int)cudaGetLastError();
}
extern "C" int b200_gap_bwd(const void* dout, void* dx, int N, int HW, int C, cudaStream_t s) {
  const long long total = (long long)N * HW * (C / VEC);
  if (fits32(total)) gap_bwd_kernel<unsigned><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dx, N, HW, C);
  else gap_bwd_kernel<long long><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dx, N, HW, C);
  return (int)cudaGetLastError();
}
extern "C" int b200_avgpool2_fwd(const void* x, void* out, int N, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / VEC);
  if (fits32(total)) avgpool2_fwd_kernel<unsigned><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, N, H, W, C);
  else avgpool2_fwd_kernel<long long><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, N, H, W, C);
  return (int)cudaGetLastError();
}
extern "C" int b200_avgpool2_bwd(const void* dout, void* dx, int N, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)N * H * W * (C / VEC);
  if (fits32(total)) avgpool2_bwd_kernel<unsigned><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dx, N, H, W, C);
  else avgpool2_bwd_kernel<long long><<<ew_grid(total, 256), 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dx, N, H, W, C);
  return (int)cudaGetLastError();
}
static inline void scale_dims(int HW, int C, int N, dim3& grid, dim3& block) {
  const int cvs = C / VEC;
  int bx = 1;
  while (bx < 32 && bx < cvs) bx <<= 1;
  const int by = 256 / bx;
  int gy = (HW + 4 * by - 1) / (4 * by);   // >= 4 pixel iterations per thread
  if (gy > 32) gy = 32;
  if (gy < 1) gy = 1;
  grid = dim3((cvs + bx - 1) / bx, gy, N);
  block = dim3(bx, by);
}
extern "C" int b200_channel_scale_fwd(const void* x, const void* gate, void* out, int N, int HW, int C, cudaStream_t s) {
  dim3 g, b;
  scale_dims(HW, C, N, g, b);
  channel_scale_fwd_kernel<<<g, b, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)gate, (__nv_bfloat16*)out, HW, C);
  return (int)cudaGetLastError();
}
extern "C" int b200_channel_scale_bwd(const void* dout, const void* x, const void* gate, void* dx, float* dgate, int N, int HW,
                                      int C, cudaStream_t s) {
  dim3 g, b;
  scale_dims(HW, C, N, g, b);
  channel_scale_bwd_kernel<<<g, b, 0, s>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gate,
                                           (__nv_bfloat16*)dx, dgate, HW, C);
  return (int)cudaGetLastError();
}
extern "C" int b200_ce_topk(const void* logits, const long long* target, void* dlogits, float* accum, int rows, int ncls,
                            long long ld, int topk, float grad_scale, cudaStream_t s) {
  ce_topk_kernel<<<rows, 256, 0, s>>>((const __nv_bfloat16*)logits, target, (__nv_bfloat16*)dlogits, accum, ncls, ld, topk, grad_scale);
  return (int)cudaGetLastError();
}
extern "C" int b200_nchw_to_nhwc(const float* x, void* out, int N, int C, int H, int W, cudaStream_t s) {
  nchw_to_nhwc_bf16_kernel<<<ew_grid((long long)N * C * H * W, 256), 256, 0, s>>>(x, (__nv_bfloat16*)out, N, C, H, W);
  return (int)cudaGetLastError();
}
extern "C" int b200_stem_im2col(const void* x, int x_is_u8, void* patches, int N, int C, int H, int W, int P, int Q, int R, int S,
                                int stride, int pad, int Kpad, const StemNorm* norm, cudaStream_t s) {
  if (Kpad > 256 || C > 8) return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)C * R * (W + 2 * pad) * sizeof(float);
  if (x_is_u8)
    stem_im2col_kernel<uint8_t><<<N * P, 256, smem, s>>>((const uint8_t*)x, (__nv_bfloat16*)patches, N, C, H, W, P, Q, R, S, stride, pad, Kpad, *norm);
  else
    stem_im2col_kernel<float><<<N * P, 256, smem, s>>>((const float*)x, (__nv_bfloat16*)patches, N, C, H, W, P, Q, R, S, stride, pad, Kpad, *norm);
  return (int)cudaGetLastError();
}
extern "C" int b200_pad_rows(const void* src, void* dst, int rows, int cols, int cols_pad, cudaStream_t s) {
  pad_rows_kernel<<<ew_grid((long long)rows * cols_pad, 256), 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, rows, cols, cols_pad);
  return (int)cudaGetLastError();
}
extern "C" int b200_unpad_add(const float* src, float* dst, int rows, int cols, int cols_pad, cudaStream_t s) {
  unpad_add_kernel<<<ew_grid((long long)rows * cols, 256), 256, 0, s>>>(src, dst, rows, cols, cols_pad);
  return (int)cudaGetLastError();
}
