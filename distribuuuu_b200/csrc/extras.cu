// Small memory-bound kernels that remove the last ATen ops from the native training step:
//   * colsum_add           : bias gradient of the classifier / SE layers  (was `d.float().sum(0)` + add_)
//   * parity_interleave    : assembles the data gradient of a stride-2 convolution from its (<= 4) parity classes,
//                            each computed by a compact stride-1 tcgen05 dgrad (no zero insertion, no ATen scatter)
//   * strided_add_inplace  : dx[:, ::2, ::2] += compact   (1x1 stride-2 projection joining an existing gradient)
//   * blockdiag_pack / blockdiag_unpack_add : thin-group (ResNeXt, 4..16 channels per group) weights <-> 64-channel
//                            block-diagonal weights, so those convolutions run on the grouped tcgen05 path
//   * se_gate_fwd / se_gate_bwd / channel_add_bcast : the squeeze-excite MLP (fc -> act -> fc -> sigmoid) on pooled
//                            vectors, any widths (EfficientNet's 4/6/10/20/28, RegNetY's 308 ...), SURVEY G16
#include <algorithm>

#include "common.cuh"
#include "extras.h"

namespace b200 {

namespace {
struct alignas(16) BF8x { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  BF8x raw = *reinterpret_cast<const BF8x*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(raw.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  BF8x raw;
#pragma unroll
  for (int i = 0; i < 4; ++i) raw.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<BF8x*>(p) = raw;
}
__device__ __forceinline__ float act_f(float z, int act) {
  if (act == 1) return fmaxf(z, 0.f);
  if (act == 2) return z / (1.f + __expf(-z));
  return z;
}
__device__ __forceinline__ float act_df(float z, int act) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) { const float s = 1.f / (1.f + __expf(-z)); return s * (1.f + z * (1.f - s)); }
  return 1.f;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ colsum_add
// out[c] += sum_rows d[row][c];  block (32, 8): thread = 8 channels x strided rows, smem reduce over y, one atomic per channel
__global__ void __launch_bounds__(256) colsum_add_kernel(const __nv_bfloat16* __restrict__ d, long long rows, int C, long long ld,
                                                         float* __restrict__ out) {
  const int cv = blockIdx.x * 32 + threadIdx.x;
  const int c0 = cv * 8;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  if (c0 < C) {
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      float v[8];
      ld8(d + r * ld + c0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += v[i];
    }
  }
  __shared__ float sm[8][32][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.y][threadIdx.x][i] = a[i];
  __syncthreads();
  if (threadIdx.y == 0 && c0 < C) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      for (int y = 0; y < 8; ++y) s += sm[y][threadIdx.x][i];
      atomicAdd(out + c0 + i, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ parity interleave
// dx[n, h, w, :] = src[h & 1][w & 1][n, h >> 1, w >> 1, :] (+ addend[n, h, w, :]);  a null source is all zeros.
// Source (ph, pw) is a contiguous [N, ceil((H - ph) / 2), ceil((W - pw) / 2), C] tensor.
struct ParityParams {
  const __nv_bfloat16* src[4];
  const __nv_bfloat16* addend;
  __nv_bfloat16* dx;
  int N, H, W, C;
};
__global__ void __launch_bounds__(256) parity_interleave_kernel(ParityParams p) {
  const int cvs = p.C / 8;
  const int per_row = p.W * cvs;
  for (int row = blockIdx.x; row < p.N * p.H; row += gridDim.x) {
    const int n = row / p.H, h = row - n * p.H;
    const int ph = h & 1, a = h >> 1;
    const int Hc = (p.H - ph + 1) >> 1;
    for (int it = threadIdx.x; it < per_row; it += blockDim.x) {
      const int w = it / cvs, cv = it - w * cvs;
      const int pw = w & 1, b = w >> 1;
      const int Wc = (p.W - pw + 1) >> 1;
      const long long o = ((long long)row * p.W + w) * p.C + cv * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      const __nv_bfloat16* s = p.src[ph * 2 + pw];
      if (s != nullptr) ld8(s + (((long long)n * Hc + a) * Wc + b) * p.C + cv * 8, v);
      if (p.addend != nullptr) {
        float q[8];
        ld8(p.addend + o, q);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += q[i];
      }
      st8(p.dx + o, v);
    }
  }
}

// dx[n, s*i, s*j, :] += compact[n, i, j, :]   (touches only the sampled pixels)
__global__ void __launch_bounds__(256) strided_add_inplace_kernel(__nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ compact,
                                                                  int N, int H, int W, int C, int P, int Q, int stride) {
  const int cvs = C / 8;
  const long long total = (long long)N * P * Q * cvs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvs);
    long long pix = idx / cvs;
    const int j = (int)(pix % Q); pix /= Q;
    const int i = (int)(pix % P); const int n = (int)(pix / P);
    float a[8], b[8];
    ld8(compact + idx * 8, a);
    __nv_bfloat16* dst = dx + (((long long)n * H + (long long)i * stride) * W + (long long)j * stride) * C + cv * 8;
    ld8(dst, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    st8(dst, a);
  }
}

// ------------------------------------------------------------------------------------------------ block-diagonal weights
// thin [K][taps][cg] (cg = channels per group, cin_g == cout_g == cg, 64 % cg == 0)  <->  dense-in-block [K][taps][64]:
// output channel k lives in 64-block k / 64; its inputs are the cg channels of its own group inside that block.
__global__ void blockdiag_pack_kernel(const __nv_bfloat16* __restrict__ thin, __nv_bfloat16* __restrict__ dense, int K, int taps, int cg) {
  const long long total = (long long)K * taps * 64;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 63);
    const long long kt = idx >> 6;           // k * taps + tap
    const int k = (int)(kt / taps);
    const int g_in_block = (k & 63) / cg;    // which of the 64/cg groups of this block the output channel belongs to
    const int jj = j - g_in_block * cg;
    dense[idx] = (jj >= 0 && jj < cg) ? thin[kt * cg + jj] : __float2bfloat16(0.f);
  }
}
__global__ void blockdiag_unpack_add_kernel(const float* __restrict__ dense, float* __restrict__ thin, int K, int taps, int cg) {
  const long long total = (long long)K * taps * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int jj = (int)(idx % cg);
    const long long kt = idx / cg;
    const int k = (int)(kt / taps);
    const int g_in_block = (k & 63) / cg;
    thin[idx] += dense[kt * 64 + g_in_block * cg + jj];
  }
}

// ------------------------------------------------------------------------------------------------ squeeze-excite MLP
// gate[n][c] = sigmoid(b2[c] + sum_j w2[c][j] * act(b1[j] + sum_c' w1[j][c'] * s[n][c']))   (s = pooled input, bf16)
// The MLP is a pair of skinny GEMMs ([N x C] x [C x r], r = 4 .. 350).  Every kernel below puts the WIDE dimension on the
// grid -- hidden units (fc1) or channels (fc2) in x, groups of kSeS samples in y -- so that even a batch of 64 fills the
// machine; the first version (one CTA per 4 samples doing everything) ran on 16 CTAs and was 40 % of a RegNetY-160 step.
constexpr int kSeS = 4;

// fc1: pre1[n][j] = b1[j] + sum_c w1[j][c] * s[n][c].   grid (ceil(r/8), ceil(N/kSeS)), 8 warps, warp = hidden unit j
__global__ void __launch_bounds__(256) se_fc1_kernel(const __nv_bfloat16* __restrict__ s, const __nv_bfloat16* __restrict__ w1,
                                                     const float* __restrict__ b1, float* __restrict__ pre1, int N, int C, int r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const int n0 = blockIdx.y * kSeS;
  if (j >= r) return;
  float acc[kSeS];
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = 0.f;
  const __nv_bfloat16* wr = w1 + (long long)j * C;
  for (int c = lane * 8; c < C; c += 256) {          // C % 8 == 0: 16-byte vectors
    float w[8];
    ld8(wr + c, w);
#pragma unroll
    for (int q = 0; q < kSeS; ++q) {
      if (n0 + q < N) {
        float x[8];
        ld8(s + (long long)(n0 + q) * C + c, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[q] = fmaf(w[e], x[e], acc[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = warp_sum(acc[q]);
  if (lane == 0) {
    const float bb = b1 ? b1[j] : 0.f;
#pragma unroll
    for (int q = 0; q < kSeS; ++q) if (n0 + q < N) pre1[(long long)(n0 + q) * r + j] = acc[q] + bb;
  }
}

// fc2 + sigmoid: gate[n][c].   grid (ceil(C/256), ceil(N/kSeS)), thread = channel c, hidden activations in smem
__global__ void __launch_bounds__(256) se_fc2_kernel(const float* __restrict__ pre1, const __nv_bfloat16* __restrict__ w2,
                                                     const float* __restrict__ b2, __nv_bfloat16* __restrict__ gate, int N, int C, int r,
                                                     int act) {
  extern __shared__ float sm[];            // [kSeS][r]
  const int n0 = blockIdx.y * kSeS;
  for (int i = threadIdx.x; i < kSeS * r; i += blockDim.x) {
    const int q = i / r, j = i - q * r;
    sm[i] = (n0 + q < N) ? act_f(pre1[(long long)(n0 + q) * r + j], act) : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float acc[kSeS];
  const float bb = b2 ? b2[c] : 0.f;
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = bb;
  const __nv_bfloat16* wr = w2 + (long long)c * r;
  for (int j = 0; j < r; ++j) {
    const float w = __bfloat162float(wr[j]);
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = fmaf(w, sm[q * r + j], acc[q]);
  }
#pragma unroll
  for (int q = 0; q < kSeS; ++q)
    if (n0 + q < N) gate[(long long)(n0 + q) * C + c] = __float2bfloat16(1.f / (1.f + __expf(-acc[q])));
}

// backward 1: dz1[n][j] = (sum_c dz2[n][c] * w2[c][j]) * act'(pre1[n][j]),  h[n][j] = act(pre1[n][j]),
//             with dz2 = dgate * gate * (1 - gate) recomputed on the fly.   grid (ceil(r/8), ceil(N/kSeS)), warp = hidden unit
__global__ void __launch_bounds__(256) se_bwd_hidden_kernel(const float* __restrict__ dgate, const __nv_bfloat16* __restrict__ gate,
                                                            const float* __restrict__ pre1, const __nv_bfloat16* __restrict__ w2,
                                                            float* __restrict__ h_out, float* __restrict__ dz1_out, int N, int C, int r,
                                                            int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const int n0 = blockIdx.y * kSeS;
  if (j >= r) return;
  float acc[kSeS];
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float w = __bfloat162float(w2[(long long)c * r + j]);
#pragma unroll
    for (int q = 0; q < kSeS; ++q) {
      if (n0 + q < N) {
        const long long o = (long long)(n0 + q) * C + c;
        const float g = __bfloat162float(gate[o]);
        acc[q] = fmaf(w, dgate[o] * g * (1.f - g), acc[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = warp_sum(acc[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kSeS; ++q) {
      if (n0 + q < N) {
        const long long o = (long long)(n0 + q) * r + j;
        const float z = pre1[o];
        h_out[o] = act_f(z, act);
        dz1_out[o] = acc[q] * act_df(z, act);
      }
    }
  }
}

// backward 2: ds[n][c] = sum_j dz1[n][j] * w1[j][c]  (gradient wrt the pooled input) and dz2[n][c] written out for the
// weight-gradient kernel.   grid (ceil(C/256), ceil(N/kSeS)), thread = channel
__global__ void __launch_bounds__(256) se_bwd_input_kernel(const float* __restrict__ dgate, const __nv_bfloat16* __restrict__ gate,
                                                           const float* __restrict__ dz1, const __nv_bfloat16* __restrict__ w1,
                                                           float* __restrict__ dz2_out, float* __restrict__ ds, int N, int C, int r) {
  extern __shared__ float sm[];            // [kSeS][r]
  const int n0 = blockIdx.y * kSeS;
  for (int i = threadIdx.x; i < kSeS * r; i += blockDim.x) {
    const int q = i / r, j = i - q * r;
    sm[i] = (n0 + q < N) ? dz1[(long long)(n0 + q) * r + j] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float acc[kSeS];
#pragma unroll
  for (int q = 0; q < kSeS; ++q) acc[q] = 0.f;
  for (int j = 0; j < r; ++j) {
    const float w = __bfloat162float(w1[(long long)j * C + c]);
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = fmaf(sm[q * r + j], w, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < kSeS; ++q) {
    if (n0 + q < N) {
      const long long o = (long long)(n0 + q) * C + c;
      const float g = __bfloat162float(gate[o]);
      dz2_out[o] = dgate[o] * g * (1.f - g);
      ds[o] = acc[q];
    }
  }
}

// backward 3, weight part: every (c, j) pair is owned by exactly one thread, which loops over the batch (no atomics):
//   dw2[c][j] += sum_n dz2[n][c] h[n][j];  dw1[j][c] += sum_n dz1[n][j] s[n][c];  db2, db1 likewise.
// CTA = 32 channels x 8 hidden-unit lanes; blockIdx.y = pass over the hidden width (8 * kSeJ = 128 units per pass).
constexpr int kSeJ = 16, kSeNB = 8;
__global__ void __launch_bounds__(256) se_gate_bwd_weights_kernel(const float* __restrict__ dz2, const float* __restrict__ h,
                                                                  const float* __restrict__ dz1, const __nv_bfloat16* __restrict__ s,
                                                                  float* __restrict__ dw1, float* __restrict__ db1,
                                                                  float* __restrict__ dw2, float* __restrict__ db2, int N, int C, int r) {
  extern __shared__ float sm[];
  const int j0 = blockIdx.y * 8 * kSeJ;
  const int jw = min(8 * kSeJ, r - j0);  // hidden units of this pass
  float* s_h = sm;                       // [kSeNB][jw]
  float* s_d1 = sm + kSeNB * jw;         // [kSeNB][jw]
  const int cl = threadIdx.x & 31, jl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool c_ok = c < C;
  float a2[kSeJ], a1[kSeJ], sb2 = 0.f;
#pragma unroll
  for (int t = 0; t < kSeJ; ++t) { a2[t] = 0.f; a1[t] = 0.f; }
  for (int nb = 0; nb < N; nb += kSeNB) {
    __syncthreads();
    for (int i = threadIdx.x; i < kSeNB * jw; i += blockDim.x) {
      const int q = i / jw, j = i - q * jw;
      const bool ok = nb + q < N;
      s_h[i] = ok ? h[(long long)(nb + q) * r + j0 + j] : 0.f;
      s_d1[i] = ok ? dz1[(long long)(nb + q) * r + j0 + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSeNB; ++q) {
      float z2 = 0.f, sv = 0.f;
      if (c_ok && nb + q < N) {
        z2 = dz2[(long long)(nb + q) * C + c];
        sv = __bfloat162float(s[(long long)(nb + q) * C + c]);
      }
      sb2 += z2;
#pragma unroll
      for (int t = 0; t < kSeJ; ++t) {
        const int j = jl + 8 * t;
        if (j < jw) {
          a2[t] = fmaf(z2, s_h[q * jw + j], a2[t]);
          a1[t] = fmaf(s_d1[q * jw + j], sv, a1[t]);
        }
      }
    }
  }
  if (c_ok) {
#pragma unroll
    for (int t = 0; t < kSeJ; ++t) {
      const int j = jl + 8 * t;
      if (j < jw) {
        dw2[(long long)c * r + j0 + j] += a2[t];
        dw1[(long long)(j0 + j) * C + c] += a1[t];
      }
    }
    if (blockIdx.y == 0 && jl == 0 && db2) db2[c] += sb2;
  }
  if (blockIdx.x == 0 && db1) {                          // db1[j] = sum_n dz1[n][j] for the units of this pass
    for (int j = threadIdx.x; j < jw; j += blockDim.x) {
      float a = 0.f;
      for (int n = 0; n < N; ++n) a += dz1[(long long)n * r + j0 + j];
      db1[j0 + j] += a;
    }
  }
}

// dx[n][hw][c] += ds[n][c] * scale   (gradient of the global average pool folded into the SE backward)
__global__ void __launch_bounds__(256) channel_add_bcast_kernel(__nv_bfloat16* __restrict__ dx, const float* __restrict__ ds, int HW, int C,
                                                                float scale) {
  const int cv = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = cv * 8;
  if (c0 >= C) return;
  const int n = blockIdx.z;
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = ds[(long long)n * C + c0 + i] * scale;
  for (int t = blockIdx.y * blockDim.y + threadIdx.y; t < HW; t += gridDim.y * blockDim.y) {
    __nv_bfloat16* o = dx + ((long long)n * HW + t) * C + c0;
    float v[8];
    ld8(o, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += g[i];
    st8(o, v);
  }
}

// ------------------------------------------------------------------------------------------------ space-to-depth stem
// A 7x7 / stride 2 / pad 3 convolution over 3 channels equals a 4x4 / stride 1 convolution over the 2x2 space-to-depth
// image (12 channels, padded to 16):  r = 2a + u, s = 2b + v  =>  out[p,q] = sum_{a,b<4} W'[a,b,(u,v,c)] . S[p+a, q+b, (u,v,c)],
// S[i,j,(u*2+v)*3+c] = x[c, 2i+u-3, 2j+v-3].  Four horizontally adjacent 16-channel pixels of S are 128 contiguous bytes,
// i.e. ONE 64-channel "pixel" of a virtual NHWC tensor whose W-stride is 16 elements -- so the stem runs on the generic
// im2col tcgen05 kernels as a 4x1 convolution over that overlapping view (ops/native.py: StemConvFn), and the 1 GB
// explicit patch tensor of the im2col path (read again by wgrad) shrinks to the 108 MB S.
template <typename T>
__global__ void __launch_bounds__(256) stem_s2d_kernel(const T* __restrict__ x /*[N][3][H][W]*/, __nv_bfloat16* __restrict__ out /*[N][Hs][Ws][16]*/,
                                                       int N, int H, int W, int Hs, int Ws, float sc0, float sc1, float sc2,
                                                       float bi0, float bi1, float bi2) {
  const long long total = (long long)N * Hs * Ws;
  const float sc[3] = {sc0, sc1, sc2}, bi[3] = {bi0, bi1, bi2};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Ws);
    const long long t = idx / Ws;
    const int i = (int)(t % Hs), n = (int)(t / Hs);
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;      // zero padding lives in the normalised domain (Normalize -> Conv2d(padding))
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int h = 2 * i + u - 3;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int vv = 0; vv < 2; ++vv) {
        const int w = 2 * j + vv - 3;
        if (w < 0 || w >= W) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const T raw = __ldg(x + (((long long)n * 3 + c) * H + h) * W + w);
          v[(u * 2 + vv) * 3 + c] = sizeof(T) == 1 ? fmaf((float)raw, sc[c], bi[c]) : (float)raw;
        }
      }
    }
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo[k] = v[k]; hi[k] = v[8 + k]; }
    st8(out + idx * 16, lo);
    st8(out + idx * 16 + 8, hi);
  }
}
// w [K][7][7][3] bf16 -> w' [K][4][1][64] bf16: w'[k][a][b*16 + (u*2+v)*3 + c] = w[k][2a+u][2b+v][c] (0 where 2a+u or 2b+v is 7)
__global__ void stem_s2d_pack_w_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wp, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * 256) return;
  const int k = idx >> 8, a = (idx >> 6) & 3, b = (idx >> 4) & 3, ch = idx & 15;
  __nv_bfloat16 val = __float2bfloat16_rn(0.f);
  if (ch < 12) {
    const int uv = ch / 3, c = ch - uv * 3, r = 2 * a + (uv >> 1), s2 = 2 * b + (uv & 1);
    if (r < 7 && s2 < 7) val = w[((k * 7 + r) * 7 + s2) * 3 + c];
  }
  wp[idx] = val;
}
// dw [K][7][7][3] fp32 += dw' [K][4][1][64] fp32 (inverse index map of the above)
__global__ void stem_s2d_unpack_dw_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * 147) return;
  const int k = idx / 147, rem = idx - k * 147;
  const int r = rem / 21, s2 = (rem / 3) % 7, c = rem % 3;
  dw[idx] += dwp[k * 256 + (r >> 1) * 64 + (s2 >> 1) * 16 + ((r & 1) * 2 + (s2 & 1)) * 3 + c];
}

}  // namespace b200

using namespace b200;

extern "C" int b200_stem_s2d(const void* x, int x_is_u8, void* out, int N, int H, int W, int Hs, int Ws, const float* scale3,
                             const float* bias3, cudaStream_t s) {
  const long long total = (long long)N * Hs * Ws;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (x_is_u8)
    stem_s2d_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)x, (__nv_bfloat16*)out, N, H, W, Hs, Ws, scale3[0], scale3[1], scale3[2],
                                                  bias3[0], bias3[1], bias3[2]);
  else
    stem_s2d_kernel<float><<<grid, 256, 0, s>>>((const float*)x, (__nv_bfloat16*)out, N, H, W, Hs, Ws, 1.f, 1.f, 1.f, 0.f, 0.f, 0.f);
  return (int)cudaGetLastError();
}
extern "C" int b200_stem_s2d_pack_w(const void* w, void* wp, int K, cudaStream_t s) {
  stem_s2d_pack_w_kernel<<<(K * 256 + 255) / 256, 256, 0, s>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)wp, K);
  return (int)cudaGetLastError();
}
extern "C" int b200_stem_s2d_unpack_dw(const float* dwp, float* dw, int K, cudaStream_t s) {
  stem_s2d_unpack_dw_kernel<<<(K * 147 + 255) / 256, 256, 0, s>>>(dwp, dw, K);
  return (int)cudaGetLastError();
}
extern "C" int b200_colsum_add(const void* d, long long rows, int C, long long ld, float* out, cudaStream_t s) {
  dim3 grid((C / 8 + 31) / 32, (unsigned)std::min<long long>(64, (rows + 7) / 8));
  if (grid.y < 1) grid.y = 1;
  colsum_add_kernel<<<grid, dim3(32, 8), 0, s>>>((const __nv_bfloat16*)d, rows, C, ld, out);
  return (int)cudaGetLastError();
}
extern "C" int b200_parity_interleave(const void* const* src4, const void* addend, void* dx, int N, int H, int W, int C, cudaStream_t s) {
  ParityParams p{};
  for (int i = 0; i < 4; ++i) p.src[i] = (const __nv_bfloat16*)src4[i];
  p.addend = (const __nv_bfloat16*)addend; p.dx = (__nv_bfloat16*)dx; p.N = N; p.H = H; p.W = W; p.C = C;
  const int grid = std::min(N * H, 148 * 8);
  parity_interleave_kernel<<<grid, 256, 0, s>>>(p);
  return (int)cudaGetLastError();
}
extern "C" int b200_strided_add_inplace(void* dx, const void* compact, int N, int H, int W, int C, int P, int Q, int stride, cudaStream_t s) {
  const long long total = (long long)N * P * Q * (C / 8);
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  strided_add_inplace_kernel<<<std::max(grid, 1), 256, 0, s>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)compact, N, H, W, C, P, Q, stride);
  return (int)cudaGetLastError();
}
extern "C" int b200_blockdiag_pack(const void* thin, void* dense, int K, int taps, int cg, cudaStream_t s) {
  const long long total = (long long)K * taps * 64;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  blockdiag_pack_kernel<<<std::max(grid, 1), 256, 0, s>>>((const __nv_bfloat16*)thin, (__nv_bfloat16*)dense, K, taps, cg);
  return (int)cudaGetLastError();
}
extern "C" int b200_blockdiag_unpack_add(const float* dense, float* thin, int K, int taps, int cg, cudaStream_t s) {
  const long long total = (long long)K * taps * cg;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  blockdiag_unpack_add_kernel<<<std::max(grid, 1), 256, 0, s>>>(dense, thin, K, taps, cg);
  return (int)cudaGetLastError();
}
extern "C" int b200_se_gate_fwd(const void* sp, const void* w1, const float* b1, const void* w2, const float* b2, float* pre1, void* gate,
                                int N, int C, int r, int act, cudaStream_t s) {
  const int gy = (N + kSeS - 1) / kSeS;
  se_fc1_kernel<<<dim3((r + 7) / 8, gy), 256, 0, s>>>((const __nv_bfloat16*)sp, (const __nv_bfloat16*)w1, b1, pre1, N, C, r);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  se_fc2_kernel<<<dim3((C + 255) / 256, gy), 256, (size_t)kSeS * r * sizeof(float), s>>>(pre1, (const __nv_bfloat16*)w2, b2, (__nv_bfloat16*)gate,
                                                                                        N, C, r, act);
  return (int)cudaGetLastError();
}
extern "C" int b200_se_gate_bwd(const float* dgate, const void* gate, const void* sp, const float* pre1, const void* w1, const void* w2,
                                float* dw1, float* db1, float* dw2, float* db2, float* ds, float* scratch, int N, int C, int r, int act,
                                cudaStream_t s) {
  // scratch: fp32 [N*C + 2*N*r]  (dz2 | h | dz1)
  float* dz2 = scratch;
  float* h = scratch + (size_t)N * C;
  float* dz1 = h + (size_t)N * r;
  const int gy = (N + kSeS - 1) / kSeS;
  se_bwd_hidden_kernel<<<dim3((r + 7) / 8, gy), 256, 0, s>>>(dgate, (const __nv_bfloat16*)gate, pre1, (const __nv_bfloat16*)w2, h, dz1, N, C, r, act);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  se_bwd_input_kernel<<<dim3((C + 255) / 256, gy), 256, (size_t)kSeS * r * sizeof(float), s>>>(dgate, (const __nv_bfloat16*)gate, dz1,
                                                                                              (const __nv_bfloat16*)w1, dz2, ds, N, C, r);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int jw = std::min(r, 8 * kSeJ);
  se_gate_bwd_weights_kernel<<<dim3((C + 31) / 32, (r + 8 * kSeJ - 1) / (8 * kSeJ)), 256, (size_t)2 * kSeNB * jw * sizeof(float), s>>>(
      dz2, h, dz1, (const __nv_bfloat16*)sp, dw1, db1, dw2, db2, N, C, r);
  return (int)cudaGetLastError();
}
extern "C" int b200_channel_add_bcast(void* dx, const float* ds, int N, int HW, int C, float scale, cudaStream_t s) {
  const int cvs = C / 8;
  const int bx = cvs >= 32 ? 32 : (cvs >= 16 ? 16 : (cvs >= 8 ? 8 : (cvs >= 4 ? 4 : (cvs >= 2 ? 2 : 1))));
  const int by = 256 / bx;
  dim3 grid((cvs + bx - 1) / bx, std::max(1, std::min((HW + by - 1) / by, 8)), N);
  channel_add_bcast_kernel<<<grid, dim3(bx, by), 0, s>>>((__nv_bfloat16*)dx, ds, HW, C, scale);
  return (int)cudaGetLastError();
}
