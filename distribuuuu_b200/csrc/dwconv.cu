// Depthwise convolution (EfficientNet-B0's 3x3 / 5x5, stride 1 / 2) on NHWC bf16: CUDA-core direct convolution --
// one MAC per element, no tensor-core shape -- with 8-channel (16-byte) vectors.  SURVEY G5.
//   fprop : y[n,p,q,c]  = sum_{r,s} x[n, p*st-pad+r, q*st-pad+s, c] * w[c,r,s]      (+ BN statistics, optional)
//   dgrad : dx[n,h,w,c] = sum_{r,s : (h+pad-r) % st == 0, ...} dy[n,(h+pad-r)/st,(w+pad-s)/st,c] * w[c,r,s]
//   wgrad : dw[c,r,s]  += sum_{n,p,q} dy[n,p,q,c] * x[n, p*st-pad+r, q*st-pad+s, c]  (fp32, atomics across CTAs)
// Weights are the bf16 [C,k,k] block of the flat buffer (physical layout of the channels_last [C,1,k,k] parameter).
#include <algorithm>

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

// ------------------------------------------------------------------------------------------------------------------
// Fast path (EfficientNet-B0's layers: k in {3, 5}, stride in {1, 2}, pad = k / 2).  The generic kernels above spend
// their time in 64-bit index arithmetic and per-tap bounds checks and load every input vector k*k times (36 % of an
// EfficientNet-B0 step).  Here everything about the filter is a template parameter: a thread owns 8 channels x TW = 4
// adjacent output pixels of one row, loads each input vector of the (TW-1)*stride + k wide window ONCE per filter row
// and feeds it to every output it contributes to, with all tap indices resolved at compile time; tiles are
// enumerated with 32-bit arithmetic (two divisions per 4-pixel tile instead of three 64-bit ones per pixel).
// ------------------------------------------------------------------------------------------------------------------
constexpr int TW = 4;
constexpr int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

__device__ __forceinline__ void ld8_or_zero(const __nv_bfloat16* p, bool ok, float (&f)[8]) {
  if (ok) {
    ld8(p, f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
  }
}
// this CTA's weights -> smem as [tap][cw] fp32 (cw = channels of the CTA = blockDim.x * 8)
__device__ __forceinline__ void load_w_cta(const __nv_bfloat16* __restrict__ w, float* sw, int c_base, int C, int taps, int cw) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < taps * cw; i += blockDim.x * blockDim.y) {
    const int t = i / cw, c = i - t * cw;
    sw[i] = (c_base + c < C) ? __bfloat162float(w[(long long)(c_base + c) * taps + t]) : 0.f;
  }
  __syncthreads();
}

template <int K, int ST>
__global__ void __launch_bounds__(256, 2) dw_fprop_fast_kernel(DwParams p) {
  constexpr int PAD = K / 2, NCOL = (TW - 1) * ST + K;
  extern __shared__ float sw[];                       // [K*K][cw], then [2][cw] statistics
  const int cw = blockDim.x * 8;
  const int c_base = blockIdx.x * cw;
  load_w_cta(p.w, sw, c_base, p.C, K * K, cw);
  float* s_stat = sw + K * K * cw;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * cw; i += 256) s_stat[i] = 0.f;
  __syncthreads();
  const int c0 = c_base + threadIdx.x * 8;
  const bool active = c0 < p.C;
  const int QT = (p.Q + TW - 1) / TW;
  const int tiles = p.N * p.P * QT;
  float ssum[8], ssq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
  if (active) {
    const float* wt = sw + threadIdx.x * 8;
    for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < tiles; t += gridDim.y * blockDim.y) {
      const int qt = t % QT, rest = t / QT;
      const int ph = rest % p.P, n = rest / p.P;
      const int q0 = qt * TW, w_in0 = q0 * ST - PAD, h_in0 = ph * ST - PAD;
      float acc[TW][8];
#pragma unroll
      for (int tw = 0; tw < TW; ++tw)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[tw][i] = 0.f;
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const int h = h_in0 + r;
        if (h < 0 || h >= p.H) continue;
        float wr[K][8];
#pragma unroll
        for (int sx = 0; sx < K; ++sx)
#pragma unroll
          for (int i = 0; i < 8; ++i) wr[sx][i] = wt[(r * K + sx) * cw + i];
        const __nv_bfloat16* row = p.x + ((long long)(n * p.H + h) * p.W) * p.C + c0;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          const int w = w_in0 + j;
          float xv[8];
          ld8_or_zero(row + (long long)w * p.C, w >= 0 && w < p.W, xv);
#pragma unroll
          for (int tw = 0; tw < TW; ++tw) {
            constexpr int dummy = 0; (void)dummy;
            const int sx = j - tw * ST;                 // compile-time after unrolling
            if (sx >= 0 && sx < K) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[tw][i] = fmaf(xv[i], wr[sx][i], acc[tw][i]);
            }
          }
        }
      }
      __nv_bfloat16* orow = p.y + ((long long)(n * p.P + ph) * p.Q + q0) * p.C + c0;
#pragma unroll
      for (int tw = 0; tw < TW; ++tw) {
        if (q0 + tw < p.Q) {
          st8(orow + (long long)tw * p.C, acc[tw]);
          if (p.stats) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float f = __bfloat162float(__float2bfloat16_rn(acc[tw][i]));
              ssum[i] += f; ssq[i] = fmaf(f, f, ssq[i]);
            }
          }
        }
      }
    }
  }
  if (p.stats) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&s_stat[threadIdx.x * 8 + i], ssum[i]); atomicAdd(&s_stat[cw + threadIdx.x * 8 + i], ssq[i]); }
    }
    __syncthreads();
    for (int i = tid; i < cw; i += 256) {
      if (c_base + i < p.C) {
        atomicAdd(p.stats + c_base + i, s_stat[i]);
        atomicAdd(p.stats + p.C + c_base + i, s_stat[cw + i]);
      }
    }
  }
}

template <int K, int ST>
__global__ void __launch_bounds__(256, 2) dw_dgrad_fast_kernel(DwParams p) {
  constexpr int PAD = K / 2;
  constexpr int FD = floor_div(PAD - (K - 1), ST);                       // first dy column relative to w0 / ST
  constexpr int NQ = floor_div(TW - 1 + PAD, ST) - FD + 1;               // dy columns feeding TW outputs
  extern __shared__ float sw[];
  const int cw = blockDim.x * 8;
  const int c_base = blockIdx.x * cw;
  load_w_cta(p.w, sw, c_base, p.C, K * K, cw);
  const int c0 = c_base + threadIdx.x * 8;
  if (c0 >= p.C) return;
  const float* wt = sw + threadIdx.x * 8;
  const int WT = (p.W + TW - 1) / TW;
  const int tiles = p.N * p.H * WT;
  for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < tiles; t += gridDim.y * blockDim.y) {
    const int wtile = t % WT, rest = t / WT;
    const int h = rest % p.H, n = rest / p.H;
    const int w0 = wtile * TW;
    float acc[TW][8];
#pragma unroll
    for (int tw = 0; tw < TW; ++tw)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[tw][i] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int hh = h + PAD - r;
      if (hh < 0 || (hh % ST) != 0) continue;
      const int ph = hh / ST;
      if (ph >= p.P) continue;
      float wr[K][8];
#pragma unroll
      for (int sx = 0; sx < K; ++sx)
#pragma unroll
        for (int i = 0; i < 8; ++i) wr[sx][i] = wt[(r * K + sx) * cw + i];
      const __nv_bfloat16* row = p.y + ((long long)(n * p.P + ph) * p.Q) * p.C + c0;     // dY
      const int q_lo = w0 / ST + FD;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const int q = q_lo + j;
        float dv[8];
        ld8_or_zero(row + (long long)q * p.C, q >= 0 && q < p.Q, dv);
#pragma unroll
        for (int tw = 0; tw < TW; ++tw) {
          const int sx = tw + PAD - (FD + j) * ST;      // compile-time after unrolling
          if (sx >= 0 && sx < K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[tw][i] = fmaf(dv[i], wr[sx][i], acc[tw][i]);
          }
        }
      }
    }
    __nv_bfloat16* orow = p.dx + ((long long)(n * p.H + h) * p.W + w0) * p.C + c0;
#pragma unroll
    for (int tw = 0; tw < TW; ++tw)
      if (w0 + tw < p.W) st8(orow + (long long)tw * p.C, acc[tw]);
  }
}

// blockIdx.z = filter row r (keeps the accumulators at K x 8 registers)
template <int K, int ST>
__global__ void __launch_bounds__(256, 2) dw_wgrad_fast_kernel(DwParams p) {
  constexpr int PAD = K / 2, NCOL = (TW - 1) * ST + K;
  extern __shared__ float s_acc[];                    // [K][cw]
  const int r = blockIdx.z;
  const int cw = blockDim.x * 8;
  const int c_base = blockIdx.x * cw;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < K * cw; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int c0 = c_base + threadIdx.x * 8;
  float acc[K][8];
#pragma unroll
  for (int sx = 0; sx < K; ++sx)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[sx][i] = 0.f;
  if (c0 < p.C) {
    const int QT = (p.Q + TW - 1) / TW;
    const int tiles = p.N * p.P * QT;
    for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < tiles; t += gridDim.y * blockDim.y) {
      const int qt = t % QT, rest = t / QT;
      const int ph = rest % p.P, n = rest / p.P;
      const int h = ph * ST - PAD + r;
      if (h < 0 || h >= p.H) continue;
      const int q0 = qt * TW, w_in0 = q0 * ST - PAD;
      float dv[TW][8];
      const __nv_bfloat16* drow = p.y + ((long long)(n * p.P + ph) * p.Q + q0) * p.C + c0;   // dY
#pragma unroll
      for (int tw = 0; tw < TW; ++tw) ld8_or_zero(drow + (long long)tw * p.C, q0 + tw < p.Q, dv[tw]);
      const __nv_bfloat16* row = p.x + ((long long)(n * p.H + h) * p.W) * p.C + c0;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int w = w_in0 + j;
        float xv[8];
        ld8_or_zero(row + (long long)w * p.C, w >= 0 && w < p.W, xv);
#pragma unroll
        for (int tw = 0; tw < TW; ++tw) {
          const int sx = j - tw * ST;
          if (sx >= 0 && sx < K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[sx][i] = fmaf(dv[tw][i], xv[i], acc[sx][i]);
          }
        }
      }
    }
#pragma unroll
    for (int sx = 0; sx < K; ++sx)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_acc[sx * cw + threadIdx.x * 8 + i], acc[sx][i]);
  }
  __syncthreads();
  for (int i = tid; i < K * cw; i += 256) {
    const int sx = i / cw, c = i - sx * cw;
    if (c_base + c < p.C) atomicAdd(p.dw + ((long long)(c_base + c) * K + r) * K + sx, s_acc[i]);
  }
}

}  // namespace b200

using namespace b200;

struct DwLaunch { dim3 grid, block; int cw; };
static inline DwLaunch dw_fast_launch(const DwParams* p, long long tiles, int z) {
  const int cvs = p->C / 8;
  int cvx = 1;
  while (cvx < 32 && cvx < cvs) cvx <<= 1;
  const int by = 256 / cvx;
  const int gx = (cvs + cvx - 1) / cvx;
  long long gy = (tiles + by - 1) / by;                 // one tile per thread row ...
  const long long cap = std::max(1, (148 * 6) / (gx * z));   // ... but no more than ~6 CTAs per SM in total
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return DwLaunch{dim3(gx, (unsigned)gy, z), dim3(cvx, by), cvx * 8};
}
static inline bool dw_fast_ok(const DwParams* p) { return (p->k == 3 || p->k == 5) && (p->stride == 1 || p->stride == 2) && p->pad == p->k / 2; }
#define B200_DW_DISPATCH(KERN, L, SMEM)                                                   \
  do {                                                                                    \
    if (p->k == 3 && p->stride == 1) KERN<3, 1><<<L.grid, L.block, SMEM, s>>>(*p);         \
    else if (p->k == 3) KERN<3, 2><<<L.grid, L.block, SMEM, s>>>(*p);                      \
    else if (p->stride == 1) KERN<5, 1><<<L.grid, L.block, SMEM, s>>>(*p);                 \
    else KERN<5, 2><<<L.grid, L.block, SMEM, s>>>(*p);                                     \
  } while (0)

static inline int dw_grid_y(long long npix) {
  long long want = (npix + 31) / 32;
  long long cap = 148 * 6;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" int b200_dw_fprop(const DwParams* p, cudaStream_t s) {
  if (dw_fast_ok(p)) {
    const DwLaunch L = dw_fast_launch(p, (long long)p->N * p->P * ((p->Q + TW - 1) / TW), 1);
    const size_t smem = (size_t)(p->k * p->k + 2) * L.cw * sizeof(float);
    B200_DW_DISPATCH(dw_fprop_fast_kernel, L, smem);
    return (int)cudaGetLastError();
  }
  dim3 grid((p->C + kCT - 1) / kCT, dw_grid_y((long long)p->N * p->P * p->Q) / ((p->C + kCT - 1) / kCT) + 1);
  dw_fprop_kernel<<<grid, 256, p->k * p->k * kCT * sizeof(float), s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_dw_dgrad(const DwParams* p, cudaStream_t s) {
  if (dw_fast_ok(p)) {
    const DwLaunch L = dw_fast_launch(p, (long long)p->N * p->H * ((p->W + TW - 1) / TW), 1);
    const size_t smem = (size_t)(p->k * p->k) * L.cw * sizeof(float);
    B200_DW_DISPATCH(dw_dgrad_fast_kernel, L, smem);
    return (int)cudaGetLastError();
  }
  dim3 grid((p->C + kCT - 1) / kCT, dw_grid_y((long long)p->N * p->H * p->W) / ((p->C + kCT - 1) / kCT) + 1);
  dw_dgrad_kernel<<<grid, 256, p->k * p->k * kCT * sizeof(float), s>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int b200_dw_wgrad(const DwParams* p, cudaStream_t s) {
  if (p->k > 5) return (int)cudaErrorInvalidValue;
  if (dw_fast_ok(p)) {
    const DwLaunch L = dw_fast_launch(p, (long long)p->N * p->P * ((p->Q + TW - 1) / TW), p->k);
    const size_t smem = (size_t)p->k * L.cw * sizeof(float);
    B200_DW_DISPATCH(dw_wgrad_fast_kernel, L, smem);
    return (int)cudaGetLastError();
  }
  int gy = dw_grid_y((long long)p->N * p->P * p->Q) / (((p->C + kCT - 1) / kCT) * p->k) + 1;
  dim3 grid((p->C + kCT - 1) / kCT, gy, p->k);
  dw_wgrad_kernel<<<grid, 256, 0, s>>>(*p);
  return (int)cudaGetLastError();
}
