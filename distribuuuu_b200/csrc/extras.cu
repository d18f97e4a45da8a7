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
// One CTA handles kSeS samples so the (L2-resident) weights are streamed once per kSeS samples.
// pre1 [N][r] fp32 (pre-activation of the hidden layer) is kept for the backward pass.
constexpr int kSeS = 4;
__global__ void __launch_bounds__(256) se_gate_fwd_kernel(const __nv_bfloat16* __restrict__ s, const __nv_bfloat16* __restrict__ w1,
                                                          const float* __restrict__ b1, const __nv_bfloat16* __restrict__ w2,
                                                          const float* __restrict__ b2, float* __restrict__ pre1,
                                                          __nv_bfloat16* __restrict__ gate, int N, int C, int r, int act) {
  extern __shared__ float sm[];
  float* s_in = sm;                 // [kSeS][C]
  float* s_h = sm + kSeS * C;       // [kSeS][r]
  const int n0 = blockIdx.x * kSeS;
  const int ns = min(kSeS, N - n0);
  for (int i = threadIdx.x; i < kSeS * C; i += blockDim.x) {
    const int q = i / C, c = i - q * C;
    s_in[i] = q < ns ? __bfloat162float(s[(long long)(n0 + q) * C + c]) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < r; j += nwarps) {            // hidden unit j: warp-wide dot products over C
    float acc[kSeS];
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = 0.f;
    const __nv_bfloat16* wr = w1 + (long long)j * C;
    for (int c = lane; c < C; c += 32) {
      const float w = __bfloat162float(wr[c]);
#pragma unroll
      for (int q = 0; q < kSeS; ++q) acc[q] = fmaf(w, s_in[q * C + c], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) {
      const float bb = b1 ? b1[j] : 0.f;
#pragma unroll
      for (int q = 0; q < kSeS; ++q) {
        const float z = acc[q] + bb;
        if (q < ns) pre1[(long long)(n0 + q) * r + j] = z;
        s_h[q * r + j] = act_f(z, act);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // output channel c: sequential dot over the (small) hidden width
    float acc[kSeS];
    const float bb = b2 ? b2[c] : 0.f;
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = bb;
    const __nv_bfloat16* wr = w2 + (long long)c * r;
    for (int j = 0; j < r; ++j) {
      const float w = __bfloat162float(wr[j]);
#pragma unroll
      for (int q = 0; q < kSeS; ++q) acc[q] = fmaf(w, s_h[q * r + j], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeS; ++q)
      if (q < ns) gate[(long long)(n0 + q) * C + c] = __float2bfloat16(1.f / (1.f + __expf(-acc[q])));
  }
}

// Backward of the MLP, data part (one CTA per kSeS samples, no weight-gradient atomics):
//   dz2 [N][C] = dgate * gate * (1 - gate);  h [N][r] = act(pre1);  dz1 [N][r] = (dz2 . w2) * act'(pre1);
//   ds [N][C] = dz1 . w1   (gradient wrt the pooled input).   dz2 / h / dz1 are written out for the weight part.
__global__ void __launch_bounds__(256) se_gate_bwd_data_kernel(const float* __restrict__ dgate, const __nv_bfloat16* __restrict__ gate,
                                                               const float* __restrict__ pre1, const __nv_bfloat16* __restrict__ w1,
                                                               const __nv_bfloat16* __restrict__ w2, float* __restrict__ dz2_out,
                                                               float* __restrict__ h_out, float* __restrict__ dz1_out,
                                                               float* __restrict__ ds, int N, int C, int r, int act) {
  extern __shared__ float sm[];
  float* s_dz2 = sm;                     // [kSeS][C]
  float* s_dh = sm + kSeS * C;           // [kSeS][r]  (accumulated with shared atomics, then turned into dz1 in place)
  const int n0 = blockIdx.x * kSeS;
  const int ns = min(kSeS, N - n0);
  for (int i = threadIdx.x; i < kSeS * C; i += blockDim.x) {
    const int q = i / C, c = i - q * C;
    float dz = 0.f;
    if (q < ns) {
      const long long o = (long long)(n0 + q) * C + c;
      const float g = __bfloat162float(gate[o]);
      dz = dgate[o] * g * (1.f - g);
      dz2_out[o] = dz;
    }
    s_dz2[i] = dz;
  }
  for (int i = threadIdx.x; i < kSeS * r; i += blockDim.x) s_dh[i] = 0.f;
  __syncthreads();
  // dh[q][j] = sum_c dz2[q][c] * w2[c][j]: thread = (hidden unit j, slice of c); consecutive threads read consecutive j
  int jt = 1;
  while (jt < r && jt < 256) jt <<= 1;
  const int ct = 256 / jt;
  const int jl = threadIdx.x % jt, cs = threadIdx.x / jt;
  for (int j = jl; j < r; j += jt) {
    float acc[kSeS];
#pragma unroll
    for (int q = 0; q < kSeS; ++q) acc[q] = 0.f;
    for (int c = cs; c < C; c += ct) {
      const float w = __bfloat162float(w2[(long long)c * r + j]);
#pragma unroll
      for (int q = 0; q < kSeS; ++q) acc[q] = fmaf(w, s_dz2[q * C + c], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeS; ++q) atomicAdd(&s_dh[q * r + j], acc[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSeS * r; i += blockDim.x) {
    const int q = i / r, j = i - q * r;
    float d = 0.f;
    if (q < ns) {
      const long long o = (long long)(n0 + q) * r + j;
      const float z = pre1[o];
      d = s_dh[i] * act_df(z, act);
      h_out[o] = act_f(z, act);
      dz1_out[o] = d;
    }
    s_dh[i] = d;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {   // ds[q][c] = sum_j dz1[q][j] * w1[j][c]  (coalesced over c)
    float dsv[kSeS];
#pragma unroll
    for (int q = 0; q < kSeS; ++q) dsv[q] = 0.f;
    for (int j = 0; j < r; ++j) {
      const float w = __bfloat162float(w1[(long long)j * C + c]);
#pragma unroll
      for (int q = 0; q < kSeS; ++q) dsv[q] = fmaf(s_dh[q * r + j], w, dsv[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeS; ++q)
      if (q < ns) ds[(long long)(n0 + q) * C + c] = dsv[q];
  }
}

// Backward of the MLP, weight part: every (c, j) pair is owned by exactly one thread, which loops over the batch
// (no atomics):  dw2[c][j] += sum_n dz2[n][c] h[n][j];  dw1[j][c] += sum_n dz1[n][j] s[n][c];  db2, db1 likewise.
// CTA = 32 channels x 8 hidden-unit lanes; a thread owns hidden units jl, jl + 8, ... (<= kSeJ per pass).
constexpr int kSeJ = 16, kSeNB = 8;
__global__ void __launch_bounds__(256) se_gate_bwd_weights_kernel(const float* __restrict__ dz2, const float* __restrict__ h,
                                                                  const float* __restrict__ dz1, const __nv_bfloat16* __restrict__ s,
                                                                  float* __restrict__ dw1, float* __restrict__ db1,
                                                                  float* __restrict__ dw2, float* __restrict__ db2, int N, int C, int r) {
  extern __shared__ float sm[];
  float* s_h = sm;                       // [kSeNB][r]
  float* s_d1 = sm + kSeNB * r;          // [kSeNB][r]
  const int cl = threadIdx.x & 31, jl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool c_ok = c < C;
  for (int j0 = 0; j0 < r; j0 += 8 * kSeJ) {            // passes over the hidden width (one pass for r <= 128)
    float a2[kSeJ], a1[kSeJ], sb2 = 0.f;
#pragma unroll
    for (int t = 0; t < kSeJ; ++t) { a2[t] = 0.f; a1[t] = 0.f; }
    for (int nb = 0; nb < N; nb += kSeNB) {
      __syncthreads();
      for (int i = threadIdx.x; i < kSeNB * r; i += blockDim.x) {
        const int q = i / r, j = i - q * r;
        const bool ok = nb + q < N;
        s_h[i] = ok ? h[(long long)(nb + q) * r + j] : 0.f;
        s_d1[i] = ok ? dz1[(long long)(nb + q) * r + j] : 0.f;
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
          const int j = j0 + jl + 8 * t;
          if (j < r) {
            a2[t] = fmaf(z2, s_h[q * r + j], a2[t]);
            a1[t] = fmaf(s_d1[q * r + j], sv, a1[t]);
          }
        }
      }
    }
    if (c_ok) {
#pragma unroll
      for (int t = 0; t < kSeJ; ++t) {
        const int j = j0 + jl + 8 * t;
        if (j < r) {
          dw2[(long long)c * r + j] += a2[t];
          dw1[(long long)j * C + c] += a1[t];
        }
      }
      if (j0 == 0 && jl == 0 && db2) db2[c] += sb2;
    }
  }
  if (blockIdx.x == 0 && db1) {                          // db1[j] = sum_n dz1[n][j]
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
      float a = 0.f;
      for (int n = 0; n < N; ++n) a += dz1[(long long)n * r + j];
      db1[j] += a;
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

}  // namespace b200

using namespace b200;

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
  const size_t smem = (size_t)kSeS * (C + r) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(se_gate_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  se_gate_fwd_kernel<<<(N + kSeS - 1) / kSeS, 256, smem, s>>>((const __nv_bfloat16*)sp, (const __nv_bfloat16*)w1, b1, (const __nv_bfloat16*)w2,
                                                             b2, pre1, (__nv_bfloat16*)gate, N, C, r, act);
  return (int)cudaGetLastError();
}
extern "C" int b200_se_gate_bwd(const float* dgate, const void* gate, const void* sp, const float* pre1, const void* w1, const void* w2,
                                float* dw1, float* db1, float* dw2, float* db2, float* ds, float* scratch, int N, int C, int r, int act,
                                cudaStream_t s) {
  // scratch: fp32 [N*C + 2*N*r]  (dz2 | h | dz1)
  float* dz2 = scratch;
  float* h = scratch + (size_t)N * C;
  float* dz1 = h + (size_t)N * r;
  const size_t smem = (size_t)kSeS * (C + r) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(se_gate_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  se_gate_bwd_data_kernel<<<(N + kSeS - 1) / kSeS, 256, smem, s>>>(dgate, (const __nv_bfloat16*)gate, pre1, (const __nv_bfloat16*)w1,
                                                                  (const __nv_bfloat16*)w2, dz2, h, dz1, ds, N, C, r, act);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t smem2 = (size_t)2 * kSeNB * r * sizeof(float);
  static size_t configured2 = 0;
  if (smem2 > 48 * 1024 && smem2 > configured2) {
    e = cudaFuncSetAttribute(se_gate_bwd_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    configured2 = smem2;
  }
  se_gate_bwd_weights_kernel<<<(C + 31) / 32, 256, smem2, s>>>(dz2, h, dz1, (const __nv_bfloat16*)sp, dw1, db1, dw2, db2, N, C, r);
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
