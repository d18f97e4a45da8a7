// Depthwise convolution (EfficientNet-B0's 3x3 / 5x5, stride 1 / 2) on NHWC bf16: CUDA-core direct convolution --
// one MAC per element, no tensor-core shape -- with 8-channel (16-byte) vectors.  SURVEY G5.
//   fprop : y[n,p,q,c]  = sum_{r,s} x[n, p*st-pad+r, q*st-pad+s, c] * w[c,r,s]      (+ BN statistics, optional)
//   dgrad : dx[n,h,w,c] = sum_{r,s : (h+pad-r) % st == 0, ...} dy[n,(h+pad-r)/st,(w+pad-s)/st,c] * w[c,r,s]
//   wgrad : dw[c,r,s]  += sum_{n,p,q} dy[n,p,q,c] * x[n, p*st-pad+r, q*st-pad+s, c]  (fp32, atomics across CTAs)
// Weights are the bf16 [C,k,k] block of the flat buffer (physical layout of the channels_last [C,1,k,k] parameter).
#include "common.cuh"
#include "dwconv.h"

namespace b200 {

struct alignas(16) BF8v { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const BF8v raw = *reinterpret_cast<const BF8v*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(raw.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  BF8v raw;
#pragma unroll
  for (int i = 0; i < 4; ++i) raw.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<BF8v*>(p) = raw;
}

constexpr int kCT = 64;  // channels per CTA (8 vector lanes)

// weights of this CTA's channel tile -> smem as [tap][channel] fp32
__device__ __forceinline__ void load_w_tile(const __nv_bfloat16* __restrict__ w, float* sw, int c_base, int C, int taps) {
  for (int i = threadIdx.x; i < taps * kCT; i += blockDim.x) {
    const int t = i / kCT, c = i % kCT;
    sw[i] = (c_base + c < C) ? __bfloat162float(w[(long long)(c_base + c) * taps + t]) : 0.f;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) dw_fprop_kernel(DwParams p) {
  extern __shared__ float sw[];           // [k*k][64]
  __shared__ float s_stat[2][kCT];
  const int taps = p.k * p.k;
  const int c_base = blockIdx.x * kCT;
  load_w_tile(p.w, sw, c_base, p.C, taps);
  if (threadIdx.x < 2 * kCT) (&s_stat[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane8 = threadIdx.x & 7;       // channel vector inside the tile
  const int c0 = c_base + lane8 * 8;
  const long long npix = (long long)p.N * p.P * p.Q;
  float ssum[8], ssq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
  if (c0 < p.C) {
    for (long long pix = (long long)blockIdx.y * (blockDim.x >> 3) + (threadIdx.x >> 3); pix < npix;
         pix += (long long)gridDim.y * (blockDim.x >> 3)) {
      const int q = pix % p.Q; const long long t = pix / p.Q;
      const int ph = t % p.P; const int n = t / p.P;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int r = 0; r < p.k; ++r) {
        const int h = ph * p.stride - p.pad + r;
        if (h < 0 || h >= p.H) continue;
        for (int s = 0; s < p.k; ++s) {
          const int w = q * p.stride - p.pad + s;
          if (w < 0 || w >= p.W) continue;
          float xv[8];
          ld8(p.x + (((long long)n * p.H + h) * p.W + w) * p.C + c0, xv);
          const float* wt = sw + (r * p.k + s) * kCT + lane8 * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv[i], wt[i], acc[i]);
        }
      }
      st8(p.y + pix * p.C + c0, acc);
      if (p.stats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float f = __bfloat162float(__float2bfloat16_rn(acc[i]));
          ssum[i] += f; ssq[i] = fmaf(f, f, ssq[i]);
        }
      }
    }
  }
  if (p.stats) {
    if (c0 < p.C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&s_stat[0][lane8 * 8 + i], ssum[i]); atomicAdd(&s_stat[1][lane8 * 8 + i], ssq[i]); }
    }
    __syncthreads();
    if (threadIdx.x < kCT && c_base + threadIdx.x < p.C) {
      atomicAdd(p.stats + c_base + threadIdx.x, s_stat[0][threadIdx.x]);
      atomicAdd(p.stats + p.C + c_base + threadIdx.x, s_stat[1][threadIdx.x]);
    }
  }
}

__global__ void __launch_bounds__(256) dw_dgrad_kernel(DwParams p) {
  extern __shared__ float sw[];
  const int taps = p.k * p.k;
  const int c_base = blockIdx.x * kCT;
  load_w_tile(p.w, sw, c_base, p.C, taps);
  const int lane8 = threadIdx.x & 7;
  const int c0 = c_base + lane8 * 8;
  if (c0 >= p.C) return;
  const long long npix = (long long)p.N * p.H * p.W;
  for (long long pix = (long long)blockIdx.y * (blockDim.x >> 3) + (threadIdx.x >> 3); pix < npix;
       pix += (long long)gridDim.y * (blockDim.x >> 3)) {
    const int w = pix % p.W; const long long t = pix / p.W;
    const int h = t % p.H; const int n = t / p.H;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r = 0; r < p.k; ++r) {
      const int hh = h + p.pad - r;
      if (hh < 0 || hh % p.stride != 0) continue;
      const int ph = hh / p.stride;
      if (ph >= p.P) continue;
      for (int s = 0; s < p.k; ++s) {
        const int ww = w + p.pad - s;
        if (ww < 0 || ww % p.stride != 0) continue;
        const int q = ww / p.stride;
        if (q >= p.Q) continue;
        float dv[8];
        ld8(p.y + (((long long)n * p.P + ph) * p.Q + q) * p.C + c0, dv);   // p.y holds dY here
        const float* wt = sw + (r * p.k + s) * kCT + lane8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(dv[i], wt[i], acc[i]);
      }
    }
    st8(p.dx + pix * p.C + c0, acc);
  }
}

// one filter row (r) per blockIdx.z keeps the per-thread accumulators at k*8 registers
__global__ void __launch_bounds__(256) dw_wgrad_kernel(DwParams p) {
  __shared__ float s_acc[5][kCT];           // k <= 5
  const int r = blockIdx.z;
  const int c_base = blockIdx.x * kCT;
  for (int i = threadIdx.x; i < 5 * kCT; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  const int lane8 = threadIdx.x & 7;
  const int c0 = c_base + lane8 * 8;
  float acc[5][8];
#pragma unroll
  for (int s = 0; s < 5; ++s)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[s][i] = 0.f;
  if (c0 < p.C) {
    const long long npix = (long long)p.N * p.P * p.Q;
    for (long long pix = (long long)blockIdx.y * (blockDim.x >> 3) + (threadIdx.x >> 3); pix < npix;
         pix += (long long)gridDim.y * (blockDim.x >> 3)) {
      const int q = pix % p.Q; const long long t = pix / p.Q;
      const int ph = t % p.P; const int n = t / p.P;
      const int h = ph * p.stride - p.pad + r;
      if (h < 0 || h >= p.H) continue;
      float dv[8];
      ld8(p.y + pix * p.C + c0, dv);        // dY
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        if (s >= p.k) break;
        const int w = q * p.stride - p.pad + s;
        if (w < 0 || w >= p.W) continue;
        float xv[8];
        ld8(p.x + (((long long)n * p.H + h) * p.W + w) * p.C + c0, xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[s][i] = fmaf(dv[i], xv[i], acc[s][i]);
      }
    }
#pragma unroll
    for (int s = 0; s < 5; ++s)
      if (s < p.k)
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&s_acc[s][lane8 * 8 + i], acc[s][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.k * kCT; i += blockDim.x) {
    const int s = i / kCT, c = i % kCT;
    if (c_base + c < p.C) atomicAdd(p.dw + ((long long)(c_base + c) * p.k + r) * p.k + s, s_acc[s][c]);
  }
}

}  // namespace b200

using namespace b200;

static inline int dw_grid_y(long long npix) {
  long long want = (npix + 31) / 32;
  long long cap = 148 * 6;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" int b200_dw_fprop(const DwParams* p, cudaStream_t s) {
  dim3 grid((p->C + kCT - 1) / kCT, dw_grid_y((long long)p->N * p->P * p->Q) / ((p->C + kCT - 1) / kCT) + 1);
  dw_fprop_kernel<<<grid, 256, p->k * p->k * kCT * sizeof(float), s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_dw_dgrad(const DwParams* p, cudaStream_t s) {
  dim3 grid((p->C + kCT - 1) / kCT, dw_grid_y((long long)p->N * p->H * p->W) / ((p->C + kCT - 1) / kCT) + 1);
  dw_dgrad_kernel<<<grid, 256, p->k * p->k * kCT * sizeof(float), s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_dw_wgrad(const DwParams* p, cudaStream_t s) {
  if (p->k > 5) return (int)cudaErrorInvalidValue;
  int gy = dw_grid_y((long long)p->N * p->P * p->Q) / (((p->C + kCT - 1) / kCT) * p->k) + 1;
  dim3 grid((p->C + kCT - 1) / kCT, gy, p->k);
  dw_wgrad_kernel<<<grid, 256, 0, s>>>(*p);
  return (int)cudaGetLastError();
}
