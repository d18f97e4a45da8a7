// Python bindings (pybind11 / torch tensors) for the sm_100a kernels.  Host-side work done here:
// TMA descriptor (CUtensorMap) construction -- tiled and im2col -- with a small cache, UMMA descriptor /
// tile-shape selection, split-K planning, and marshalling of the peer-memory contexts.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <vector>

#include <mutex>
#include <unordered_map>

#include "comm.h"
#include "conv_gemm.h"
#include "dwconv.h"
#include "elementwise.h"
#include "extras.h"
#include "attention.h"
#include "conv3x3_halo.h"
#include "stem_conv.h"

namespace {

#define B200_CUDA_OK(expr)                                                                          \
  do {                                                                                              \
    int _e = (int)(expr);                                                                           \
    TORCH_CHECK(_e == 0, #expr " failed: ", cudaGetErrorString((cudaError_t)_e), " (", _e, ")");   \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename Fn>
Fn driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  TORCH_CHECK(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr, "driver entry point ", name,
              " unavailable");
  return reinterpret_cast<Fn>(fn);
}

struct MapKey {
  uint64_t v[12];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
std::mutex g_map_mutex;

// Rank-3 tiled map over a bf16 tensor: dims (d0 inner, d1, d2), strides in ELEMENTS for d1/d2, 128B swizzle.
CUtensorMap tiled_map_3d(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0,
                         uint32_t b1, uint32_t b2) {
  MapKey key{{(uint64_t)ptr, d0, d1, d2, s1, s2, b0, b1, b2, 0, 0, 1}};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) return it->second;
  static EncodeTiledFn encode = driver_fn<EncodeTiledFn>("cuTensorMapEncodeTiled");
  TORCH_CHECK(((uintptr_t)ptr & 15) == 0, "TMA base address must be 16-byte aligned");
  TORCH_CHECK((s1 * 2) % 16 == 0 && (s2 * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes (", s1, ", ", s2, ")");
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with ", (int)r, " dims=(", d0, ",", d1, ",", d2,
              ") strides=(", s1, ",", s2, ") box=(", b0, ",", b1, ",", b2, ")");
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache.emplace(key, m);
  return m;
}

// Rank-4 tiled map (weights viewed as (Cin/g, taps, Cout/g, groups)); strides in ELEMENTS for d1..d3.
CUtensorMap tiled_map_4d(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t s1, uint64_t s2,
                         uint64_t s3, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  MapKey key{{(uint64_t)ptr, d0, d1, d2, d3, s1, s2, s3, ((uint64_t)b0 << 32) | b1, ((uint64_t)b2 << 32) | b3, 0, 3}};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) return it->second;
  static EncodeTiledFn encode = driver_fn<EncodeTiledFn>("cuTensorMapEncodeTiled");
  TORCH_CHECK(((uintptr_t)ptr & 15) == 0, "TMA base address must be 16-byte aligned");
  TORCH_CHECK((s1 * 2) % 16 == 0 && (s2 * 2) % 16 == 0 && (s3 * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
  CUtensorMap m;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1 * 2, s2 * 2, s3 * 2};
  cuuint32_t box[4] = {b0, b1, b2, b3};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed with ", (int)r);
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache.emplace(key, m);
  return m;
}

// im2col map over an NHWC bf16 activation (dims C, W, H, N).
CUtensorMap im2col_map_4d(const void* ptr, int N, int H, int W, int C, int low_w, int low_h, int up_w, int up_h,
                          int stride, uint32_t channels_per_pixel, uint32_t pixels_per_column, uint64_t sw_el = 0,
                          uint64_t sh_el = 0, uint64_t sn_el = 0) {
  // sw/sh/sn: element strides of the W / H / N dimensions; 0 = dense NHWC.  They may be SMALLER than the extent below
  // them (overlapping pixels): the space-to-depth stem reads 4 neighbouring 16-channel pixels as one 64-channel pixel.
  if (!sw_el) sw_el = (uint64_t)C;
  if (!sh_el) sh_el = (uint64_t)W * sw_el;
  if (!sn_el) sn_el = (uint64_t)H * sh_el;
  auto pk = [](int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
  MapKey key{{(uint64_t)ptr, pk(N, H), pk(W, C), pk(low_w, low_h), pk(up_w, up_h), (uint64_t)stride,
              ((uint64_t)channels_per_pixel << 32) | pixels_per_column, sw_el, sh_el, sn_el, 0, 2}};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) return it->second;
  static EncodeIm2colFn encode = driver_fn<EncodeIm2colFn>("cuTensorMapEncodeIm2col");
  TORCH_CHECK(((uintptr_t)ptr & 15) == 0 && (C % 8) == 0, "im2col TMA needs 16-byte aligned base and C % 8 == 0");
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  TORCH_CHECK(sw_el % 8 == 0 && sh_el % 8 == 0 && sn_el % 8 == 0, "im2col TMA strides must be multiples of 16 bytes");
  cuuint64_t strides[3] = {(cuuint64_t)sw_el * 2, (cuuint64_t)sh_el * 2, (cuuint64_t)sn_el * 2};
  int lower[2] = {low_w, low_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper,
                      channels_per_pixel, pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed with ", (int)r, " NHWC=(", N, ",", H, ",", W, ",", C,
              ") corners=(", low_w, ",", low_h, ",", up_w, ",", up_h, ") stride=", stride);
  // Driver <= 13.1 sets a descriptor bit that breaks im2col loads on tensors smaller than 128 KiB (same
  // workaround CUTLASS applies in make_im2col_tma_copy_desc).
  int drv = 0;
  cudaDriverGetVersion(&drv);
  const uint64_t span_bytes = ((uint64_t)(N - 1) * sn_el + (uint64_t)(H - 1) * sh_el + (uint64_t)(W - 1) * sw_el + (uint64_t)C) * 2;
  if (drv <= 13010 && span_bytes < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache.emplace(key, m);
  return m;
}

inline uint64_t desc_hi_sw128(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
inline uint32_t idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// K-major operand: rows of 128 B, 8-row swizzle atoms 1024 B apart; a UMMA_K=16 step advances 32 B.
constexpr uint32_t kKMajorLbo = 16, kKMajorSbo = 1024, kKMajorStep16 = 2;
// MN-major operand: [64 k-rows][64 mn] boxes of 8 KB; 8-row groups 1024 B apart; UMMA_K=16 step = 2 groups.
constexpr uint32_t kMnMajorLbo = 8192, kMnMajorSbo = 1024, kMnMajorStep16 = 128;

int pick_bn(int n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }

// Tile plan for fprop / dgrad.  Streaming mode re-loads the BN x 64 weight tile for every (M tile, k block); when the
// CTA's whole weight slab (k_iters x BN x 128 B) fits in shared memory next to >= 2 A stages, keeping it resident
// removes that L2->SM traffic (the dominant term for the K <= 256 pointwise layers).  Pick whichever moves fewer
// bytes per 128-row tile: resident = A * ceil(N/bn);  streaming = (A + B) * ceil(N/bn).
// CTA pairs (tcgen05 cta_group::2, conv_gemm.cu): on for every eligible layer unless B200_CONV_CG=1 (A/B measurements).
bool cta_pairs_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_CONV_CG"); v = (e && atoi(e) == 1) ? 0 : 1; }
  return v == 1;
}
bool conv_m_fastest() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_CONV_ORDER"); v = (e && e[0] == 'm') ? 1 : 0; }
  return v == 1;
}
struct TilePlan { int bn; int resident; int res_stages; int cg; };
TilePlan plan_tiles(int n_cols, int k_iters, long long m_tiles, int groups, int sms) {
  constexpr int kRing = 160 * 1024, kA = 16384;
  const int bn0 = pick_bn(n_cols);
  TilePlan best{bn0, 0, 0, 1};
  if (groups != 1) return best;
  auto nblk = [&](int bn) { return (n_cols + bn - 1) / bn; };
  // Measured (profiles/conv_shapes_*): shrinking BN to make the slab fit, or running with fewer than 4 A stages,
  // loses more to A re-reads / exposed latency than the saved weight traffic gains -- so only the natural BN and
  // only slabs that leave a deep A ring qualify.
  {
    const int bn = bn0;
    const long long slab = (long long)k_iters * bn * 128;
    const int stages = slab <= kRing ? (int)((kRing - slab) / kA) : 0;
    if (stages >= 4 && nblk(bn) <= sms && m_tiles * nblk(bn) >= 2LL * sms) best = TilePlan{bn, 1, std::min(stages, 12), 1};
  }
  // streaming layers with a wide N tile: a CTA pair shares the weight tile (each CTA stages half of it)
  // (measured, profiles/r2/conv_ablation: pays from 8 k-blocks per tile on -- 1024->512 @14^2 0.069 -> 0.062 ms, 3x3 layers
  // 5-15 % -- and loses on the short-K wide-N layers such as 256->1024, which stay single-CTA)
  if (!best.resident && best.bn >= 128 && m_tiles >= 2 && k_iters >= 8 && cta_pairs_enabled()) best.cg = 2;
  return best;
}

void check_bf16_contig(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16 && t.is_contiguous(), name, " must be a contiguous CUDA bf16 tensor");
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

struct ConvGeom { int N, H, W, C, K, R, S, P, Q, stride, pad, dil; };

ConvGeom geom(const at::Tensor& x, const at::Tensor& w, int stride, int pad, int dil, int groups = 1) {
  ConvGeom g;
  g.N = x.size(0); g.H = x.size(1); g.W = x.size(2); g.C = x.size(3);
  g.K = w.size(0); g.R = w.size(1); g.S = w.size(2);
  TORCH_CHECK(w.size(3) * groups == g.C && g.K % groups == 0, "weight [", g.K, ",", w.size(3), "] x groups ", groups,
              " does not match activation channels ", g.C);
  g.stride = stride; g.pad = pad; g.dil = dil;
  g.P = (g.H + 2 * pad - dil * (g.R - 1) - 1) / stride + 1;
  g.Q = (g.W + 2 * pad - dil * (g.S - 1) - 1) / stride + 1;
  return g;
}

// ---------------------------------------------------------------------------------------------- conv fprop
// x [N,H,W,C], w [K,R,S,C], out [N,P,Q,K] (all bf16, contiguous).  stats: fp32 [2*K] accumulated, bias fp32 [K].
struct PeerState;
PeerCtx peer_ctx_for_producer(PeerState* peer);
void conv_fprop(const at::Tensor& x, const at::Tensor& w, at::Tensor& out, const c10::optional<at::Tensor>& stats,
                const c10::optional<at::Tensor>& bias, int64_t stride, int64_t pad, int64_t dil, int64_t groups, PeerState* peer) {
  check_bf16_contig(w, "w");
  // x: dense NHWC, or (non-pointwise only) any view with contiguous channels and 16-byte-multiple W/H/N strides -- the
  // im2col tensor map takes the strides as they are, including overlapping pixels (space-to-depth stem, ops/native.py)
  const bool x_dense = x.is_contiguous();
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 4 && x.stride(3) == 1, "x must be an NHWC bf16 CUDA tensor");
  c10::cuda::CUDAGuard guard(x.device());
  const ConvGeom g = geom(x, w, stride, pad, dil, groups);
  // out: [N,P,Q,K] bf16, channels contiguous, uniform pixel pitch >= K (a channel slice of a wider NHWC buffer is fine:
  // the TMA-store epilogue takes the pitch from its tensor map -- concat-free DenseNet blocks)
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kBFloat16 && out.dim() == 4 && out.numel() == (int64_t)g.N * g.P * g.Q * g.K, "bad output shape");
  const int64_t out_pitch = out.stride(2);
  TORCH_CHECK(out.stride(3) == 1 && out_pitch >= g.K && out_pitch % 8 == 0 && out.stride(1) == g.Q * out_pitch &&
              out.stride(0) == (int64_t)g.P * g.Q * out_pitch, "output must be NHWC with a uniform pixel pitch (multiple of 8)");
  TORCH_CHECK(groups == 1 || out_pitch == g.K, "grouped convolutions need a dense output");
  const int G = groups, cin_g = g.C / G, cout_g = g.K / G;
  TORCH_CHECK(cin_g % 8 == 0 && cout_g % 8 == 0, "channels per group must be multiples of 8 (", cin_g, ", ", cout_g, ")");
  const int M = g.N * g.P * g.Q;
  TilePlan plan = plan_tiles(cout_g, g.R * g.S * ((cin_g + 63) / 64), (M + 127) / 128, G, num_sms());
  if (bias.has_value()) plan.cg = 1;    // the bias epilogue (classifier) is only instantiated for single-CTA tiles
  const int bn = plan.bn;
  const bool pointwise = (g.R == 1 && g.S == 1 && stride == 1 && pad == 0);
  TORCH_CHECK(x_dense || !pointwise, "a strided activation view needs the im2col path");
  ConvGemmParams p{};
  p.b_resident = plan.resident; p.res_stages = plan.res_stages;
  p.kind = KIND_FPROP; p.epi = EPI_BF16;
  p.M = M; p.N = cout_g;
  p.groups = G; p.a_cg = cin_g; p.out_cg = cout_g;
  const int cg = plan.cg;
  p.cta_group = cg;
  p.m_fastest = conv_m_fastest() ? 1 : 0;
  p.m_blocks = (M + 128 * cg - 1) / (128 * cg); p.n_blocks = (cout_g + bn - 1) / bn;
  p.taps = g.R * g.S; p.S = g.S; p.kb_per_tap = (cin_g + 63) / 64; p.dil = dil;
  p.a_im2col = pointwise ? 0 : 1; p.a_nbox = 1; p.a_kstep16 = kKMajorStep16; p.a_desc_hi = desc_hi_sw128(kKMajorLbo, kKMajorSbo);
  p.b_im2col = 0; p.b_nbox = 1; p.b_kstep16 = kKMajorStep16; p.b_desc_hi = desc_hi_sw128(kKMajorLbo, kKMajorSbo);
  p.b_flip_taps = 0;
  p.idesc = idesc_bf16(128 * cg, bn, 0, 0);
  p.im_P = g.P; p.im_Q = g.Q; p.im_stride = stride; p.im_low_w = -pad; p.im_low_h = -pad;
  p.k_blocks_total = 0; p.splits = 1;
  p.out = out.data_ptr(); p.ldo = out_pitch; p.tap_stride = 0;
  p.stats = stats.has_value() ? stats->data_ptr<float>() : nullptr;
  p.bias = bias.has_value() ? bias->data_ptr<float>() : nullptr;
  if (bias.has_value()) p.epi = EPI_BF16_BIAS;
  if (stats.has_value()) TORCH_CHECK(stats->numel() >= 2 * g.K && stats->scalar_type() == at::kFloat, "stats must be fp32 [2*K]");
  p.vb_per_item = 0; p.cin_boxes = 1; p.vboxes_total = 0;
  p.total_items = p.m_blocks * p.n_blocks * G;
  p.peer = PeerCtx{}; p.peer.world = 1;
  if (peer != nullptr) { TORCH_CHECK(stats.has_value(), "conv_fprop: a peer context needs the statistics epilogue"); p.peer = peer_ctx_for_producer(peer); }
  CUtensorMap ma = pointwise
                       ? tiled_map_3d(x.data_ptr(), g.C, 1, M, g.C, g.C, 64, 1, 128)
                       : im2col_map_4d(x.data_ptr(), g.N, g.H, g.W, g.C, -pad, -pad, pad - (g.S - 1) * dil,
                                       pad - (g.R - 1) * dil, stride, 64, 128, x.stride(2), x.stride(1), x.stride(0));
  CUtensorMap mb = tiled_map_4d(w.data_ptr(), cin_g, p.taps, cout_g, G, cin_g, (uint64_t)p.taps * cin_g,
                                (uint64_t)p.taps * cin_g * cout_g, 64, 1, bn / cg, 1);   // pair: each CTA loads half the N tile
  CUtensorMap mo = tiled_map_3d(out.data_ptr(), cout_g, M, G, out_pitch, cout_g, 64, 32, 1);
  int grid = cg == 2 ? 2 * std::min(p.total_items, num_sms() / 2) : std::min(p.total_items, num_sms());
  if (p.b_resident) grid = (grid / p.n_blocks) * p.n_blocks;   // every CTA keeps ONE n-block for its whole life
  B200_CUDA_OK(b200_conv_gemm_launch(&ma, &mb, &mo, &mo, &p, bn, grid, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- conv dgrad
// Generic launcher: out [N,Ho,Wo,C] (compact, contiguous) = sum over a virtual Rv x Sv window of dy [N,P,Q,K]:
//   out[n,a,b,:] = sum_{r',s'} dy[n, a + low_h + r'*dil, b + low_w + s'*dil, :] . W[:, tap(r',s'), :]
// where tap(r',s') is either the flipped tap (ordinary stride-1 dgrad) or given by a look-up table (one parity
// class of a strided dgrad).  Reads outside dy are zero-filled by the im2col TMA.
struct DgradGeom {
  int Ho, Wo;            // output grid
  int Rv, Sv;            // virtual window
  int low_h, low_w;      // offset of window tap 0 relative to the output coordinate
  int dil;
  const unsigned char* lut;   // nullptr: flipped taps of the full RxS filter (Rv == R, Sv == S)
};
void dgrad_launch(const at::Tensor& dy, const at::Tensor& w, at::Tensor& out, const DgradGeom& gm,
                  const c10::optional<at::Tensor>& addend, int64_t groups) {
  const int N = dy.size(0), P = dy.size(1), Q = dy.size(2), K = dy.size(3);
  const int R = w.size(1), S = w.size(2);
  const int C = out.size(3);
  const int G = groups, cin_g = C / G, cout_g = K / G;
  TORCH_CHECK(w.size(0) == K && w.size(3) == cin_g && out.size(0) == N && out.size(1) == gm.Ho && out.size(2) == gm.Wo, "dgrad shape mismatch");
  TORCH_CHECK(cin_g % 8 == 0 && cout_g % 8 == 0, "channels per group must be multiples of 8");
  const int taps_v = gm.Rv * gm.Sv;
  TORCH_CHECK(taps_v >= 1 && taps_v <= 25, "virtual window too large");
  const int M = N * gm.Ho * gm.Wo;
  const TilePlan plan = plan_tiles(cin_g, taps_v * ((cout_g + 63) / 64), (M + 127) / 128, G, num_sms());
  const int bn = plan.bn;
  const bool pointwise = (taps_v == 1 && gm.low_h == 0 && gm.low_w == 0 && gm.Ho == P && gm.Wo == Q);
  ConvGemmParams p{};
  p.b_resident = plan.resident; p.res_stages = plan.res_stages;
  p.kind = KIND_DGRAD; p.epi = EPI_BF16;
  p.M = M; p.N = cin_g;
  p.groups = G; p.a_cg = cout_g; p.out_cg = cin_g;
  const int cg = plan.cg;
  p.cta_group = cg;
  p.m_fastest = conv_m_fastest() ? 1 : 0;
  p.m_blocks = (M + 128 * cg - 1) / (128 * cg); p.n_blocks = (cin_g + bn - 1) / bn;
  p.taps = taps_v; p.S = gm.Sv; p.kb_per_tap = (cout_g + 63) / 64; p.dil = gm.dil;
  p.a_im2col = pointwise ? 0 : 1; p.a_nbox = 1; p.a_kstep16 = kKMajorStep16; p.a_desc_hi = desc_hi_sw128(kKMajorLbo, kKMajorSbo);
  p.b_im2col = 0; p.b_nbox = bn / 64 / cg; p.b_kstep16 = kMnMajorStep16; p.b_desc_hi = desc_hi_sw128(kMnMajorLbo, kMnMajorSbo);
  p.b_flip_taps = gm.lut ? 0 : 1;
  if (gm.lut) { p.tap_lut_on = 1; for (int t = 0; t < taps_v; ++t) p.tap_lut[t] = gm.lut[t]; }
  p.idesc = idesc_bf16(128 * cg, bn, 0, 1);
  p.im_P = gm.Ho; p.im_Q = gm.Wo; p.im_stride = 1; p.im_low_w = gm.low_w; p.im_low_h = gm.low_h;
  p.splits = 1;
  p.out = out.data_ptr(); p.ldo = C;
  p.total_items = p.m_blocks * p.n_blocks * G;
  // bounding box of the base pixel: Wo (Ho) positions starting at low  =>  upper corner = Wo - Q + low
  CUtensorMap ma = pointwise ? tiled_map_3d(dy.data_ptr(), K, 1, M, K, K, 64, 1, 128)
                             : im2col_map_4d(dy.data_ptr(), N, P, Q, K, gm.low_w, gm.low_h, gm.Wo - Q + gm.low_w,
                                             gm.Ho - P + gm.low_h, 1, 64, 128);
  // weights viewed as (C inner, taps, K): MN-major B boxes of [64 k-rows (Cout)][64 n (Cin)]
  const int taps_w = R * S;
  CUtensorMap mb = tiled_map_4d(w.data_ptr(), cin_g, taps_w, cout_g, G, cin_g, (uint64_t)taps_w * cin_g,
                                (uint64_t)taps_w * cin_g * cout_g, 64, 1, 64, 1);
  CUtensorMap mo = tiled_map_3d(out.data_ptr(), cin_g, M, G, C, cin_g, 64, 32, 1);
  CUtensorMap md = mo;
  if (addend.has_value()) {
    check_bf16_contig(*addend, "addend");
    TORCH_CHECK(addend->numel() == out.numel(), "addend must have dx's shape");
    p.addend = 1;
    p.epi = EPI_BF16_ADD;
    md = tiled_map_3d(addend->data_ptr(), cin_g, M, G, C, cin_g, 64, 32, 1);
  }
  int grid = cg == 2 ? 2 * std::min(p.total_items, num_sms() / 2) : std::min(p.total_items, num_sms());
  if (p.b_resident) grid = (grid / p.n_blocks) * p.n_blocks;
  B200_CUDA_OK(b200_conv_gemm_launch(&ma, &mb, &mo, &md, &p, bn, grid, cur_stream()));
}

// dy [N,P,Q,K], w [K,R,S,C] -> dx [N,H,W,C], stride 1.
void conv_dgrad(const at::Tensor& dy, const at::Tensor& w, at::Tensor& dx, int64_t stride, int64_t pad, int64_t dil,
                const c10::optional<at::Tensor>& addend, int64_t groups) {
  check_bf16_contig(dy, "dy"); check_bf16_contig(w, "w"); check_bf16_contig(dx, "dx");
  TORCH_CHECK(stride == 1, "conv_dgrad handles stride 1; strided layers go through conv_dgrad_s2");
  c10::cuda::CUDAGuard guard(dy.device());
  const int H = dx.size(1), W = dx.size(2);
  const int R = w.size(1), S = w.size(2);
  const int P = dy.size(1), Q = dy.size(2);
  TORCH_CHECK(P == H + 2 * pad - dil * (R - 1) && Q == W + 2 * pad - dil * (S - 1), "dgrad: dy spatial size inconsistent");
  const int padp_h = dil * (R - 1) - pad, padp_w = dil * (S - 1) - pad;  // padding of the transposed problem
  DgradGeom gm{H, W, R, S, -padp_h, -padp_w, (int)dil, nullptr};
  dgrad_launch(dy, w, dx, gm, addend, groups);
}

// Stride-2 data gradient without zero insertion: dx is split into its four (row parity, column parity) classes; class
// (ph, pw) only receives the taps r = (ph + pad) mod 2, s = (pw + pad) mod 2 (mod 2), so it is a compact stride-1
// dgrad over dy with a 1- or 2-wide virtual window -- exactly the real FLOPs (the zero-insertion path of round 1 ran
// 4x as many) -- and the classes are interleaved into dx by one pass (extras.cu), optionally adding `addend`.
void conv_dgrad_s2(const at::Tensor& dy, const at::Tensor& w, at::Tensor& dx, int64_t pad, const c10::optional<at::Tensor>& addend,
                   int64_t groups) {
  check_bf16_contig(dy, "dy"); check_bf16_contig(w, "w"); check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dy.device());
  const int N = dx.size(0), H = dx.size(1), W = dx.size(2), C = dx.size(3);
  const int R = w.size(1), S = w.size(2);
  const int P = dy.size(1), Q = dy.size(2);
  TORCH_CHECK(P == (H + 2 * pad - R) / 2 + 1 && Q == (W + 2 * pad - S) / 2 + 1, "dgrad_s2: dy spatial size inconsistent");
  TORCH_CHECK(R <= 5 && S <= 5, "dgrad_s2: filter too large");
  at::Tensor parts[4];
  const void* src[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      const int Hc = (H - ph + 1) / 2, Wc = (W - pw + 1) / 2;
      if (Hc <= 0 || Wc <= 0) continue;
      // taps of this class, in order of increasing dy offset: r = r_max, r_max - 2, ...
      int r_max = -1, s_max = -1;
      for (int r = R - 1; r >= 0; --r) if (((ph + pad - r) & 1) == 0) { r_max = r; break; }
      for (int sx = S - 1; sx >= 0; --sx) if (((pw + pad - sx) & 1) == 0) { s_max = sx; break; }
      if (r_max < 0 || s_max < 0) continue;                     // no tap reaches this class: it stays zero
      const int Rv = r_max / 2 + 1, Sv = s_max / 2 + 1;
      unsigned char lut[28];
      for (int rv = 0; rv < Rv; ++rv)
        for (int sv = 0; sv < Sv; ++sv) lut[rv * Sv + sv] = (unsigned char)((r_max - 2 * rv) * S + (s_max - 2 * sv));
      DgradGeom gm{Hc, Wc, Rv, Sv, ((int)(ph + pad) - r_max) / 2, ((int)(pw + pad) - s_max) / 2, 1, lut};
      parts[ph * 2 + pw] = at::empty({N, Hc, Wc, C}, dx.options());
      dgrad_launch(dy, w, parts[ph * 2 + pw], gm, c10::nullopt, groups);
      src[ph * 2 + pw] = parts[ph * 2 + pw].data_ptr();
    }
  }
  if (addend.has_value()) { check_bf16_contig(*addend, "addend"); TORCH_CHECK(addend->numel() == dx.numel(), "addend must have dx's shape"); }
  B200_CUDA_OK(b200_parity_interleave(src, addend.has_value() ? addend->data_ptr() : nullptr, dx.data_ptr(), N, H, W, C, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- conv wgrad
// dy [N,P,Q,K], x [N,H,W,C] -> dw fp32 [K,R,S,C] (accumulated with red.add; caller zeroes).
void conv_wgrad(const at::Tensor& dy, const at::Tensor& x, at::Tensor& dw, int64_t stride, int64_t pad, int64_t dil,
                int64_t groups) {
  check_bf16_contig(dy, "dy");
  TORCH_CHECK(dw.is_cuda() && dw.scalar_type() == at::kFloat && dw.is_contiguous(), "dw must be contiguous fp32");
  c10::cuda::CUDAGuard guard(dy.device());
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3);
  // x: NHWC bf16; a pointwise layer may read a channel slice of a wider buffer (uniform pixel pitch)
  const int64_t x_pitch = x.stride(2);
  const bool x_uniform = x_pitch >= C && x.stride(1) == W * x_pitch && x.stride(0) == (int64_t)H * W * x_pitch;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.stride(3) == 1 && x_pitch % 8 == 0,
              "x must be NHWC bf16 with contiguous channels");
  const int K = dw.size(0), R = dw.size(1), S = dw.size(2);
  const int P = dy.size(1), Q = dy.size(2);
  const int G = groups, cin_g = C / G, cout_g = K / G;
  TORCH_CHECK(dy.size(3) == K && dw.size(3) == cin_g && dy.size(0) == N, "wgrad shape mismatch");
  TORCH_CHECK(cin_g % 8 == 0 && cout_g % 8 == 0, "channels per group must be multiples of 8");
  const long long pixels = (long long)N * P * Q;
  const bool pointwise = (R == 1 && S == 1 && stride == 1 && pad == 0);
  // pointwise: a channel slice of a wider buffer (uniform pixel pitch); otherwise any strided view, overlapping pixels
  // included (the im2col tensor map takes the strides as they are)
  TORCH_CHECK(!pointwise || x_uniform, "1x1 weight gradients need a uniform pixel pitch");
  // B operand = "virtual boxes": (tap, 64-channel slice of Cin).  One work item accumulates up to 4 of them
  // (N = 64..256 TMEM columns) against a single load of the dY^T tile, so dY is re-read taps*Cin/(64*vpi) times
  // instead of taps*Cin/64 times.
  const int cin_boxes = (cin_g + 63) / 64;
  const int vboxes = R * S * cin_boxes;
  int vpi = 1;
  for (int cand = 4; cand >= 1; --cand) if (vboxes % cand == 0) { vpi = cand; break; }
  if (vpi == 1 && vboxes > 4) vpi = 4;            // no nice divisor: groups of 4 with a short tail group
  const int bn = vpi > 2 ? 256 : (vpi == 2 ? 128 : 64);
  // CTA pairs: two 128-row blocks of Cout share the item's X boxes (each CTA stages half of them)
  const int cg = (cta_pairs_enabled() && cout_g >= 256 && (vpi == 2 || vpi == 4) && vboxes % vpi == 0) ? 2 : 1;
  ConvGemmParams p{};
  p.kind = KIND_WGRAD; p.epi = EPI_F32_RED;
  p.cta_group = cg;
  p.M = cout_g; p.N = cin_g;
  p.groups = G; p.a_cg = cin_g; p.out_cg = cout_g;
  p.m_blocks = (cout_g + 128 * cg - 1) / (128 * cg); p.n_blocks = (vboxes + vpi - 1) / vpi;
  p.vb_per_item = vpi; p.cin_boxes = cin_boxes; p.vboxes_total = vboxes;
  p.taps = R * S; p.S = S; p.kb_per_tap = 0; p.dil = dil;
  p.a_im2col = 0; p.a_nbox = 2; p.a_kstep16 = kMnMajorStep16; p.a_desc_hi = desc_hi_sw128(kMnMajorLbo, kMnMajorSbo);
  p.b_im2col = pointwise ? 0 : 1; p.b_nbox = vpi; p.b_kstep16 = kMnMajorStep16; p.b_desc_hi = desc_hi_sw128(kMnMajorLbo, kMnMajorSbo);
  p.b_flip_taps = 0;
  p.idesc = idesc_bf16(128 * cg, 64 * vpi, 1, 1);
  p.im_P = P; p.im_Q = Q; p.im_stride = stride; p.im_low_w = -pad; p.im_low_h = -pad;
  p.k_blocks_total = (int)((pixels + 63) / 64);
  const int base_items = p.m_blocks * p.n_blocks * G;
  // split-K so that the item count is just UNDER a whole number of waves (an extra partial wave costs a full
  // item time): aim at 2 waves, fall back to no split for shapes that already have many items
  const int sms = num_sms() / cg;     // schedulable units: SMs, or SM pairs
  int splits = (2 * sms) / base_items;
  if (splits < 1) splits = 1;
  splits = std::max(1, std::min(splits, std::max(1, p.k_blocks_total / 8)));
  p.splits = splits;
  p.out = dw.data_ptr(); p.ldo = (long long)p.taps * cin_g; p.tap_stride = cin_g;
  p.total_items = base_items * splits;
  CUtensorMap ma = tiled_map_3d(dy.data_ptr(), K, 1, pixels, K, K, 64, 1, 64);
  CUtensorMap mb = pointwise ? tiled_map_3d(x.data_ptr(), C, 1, pixels, x_pitch, x_pitch, 64, 1, 64)
                             : im2col_map_4d(x.data_ptr(), N, H, W, C, -pad, -pad, pad - (S - 1) * dil,
                                             pad - (R - 1) * dil, stride, 64, 64, x.stride(2), x.stride(1), x.stride(0));
  const int grid = cg == 2 ? 2 * std::min(p.total_items, num_sms() / 2) : std::min(p.total_items, num_sms());
  B200_CUDA_OK(b200_conv_gemm_launch(&ma, &mb, &ma, &ma, &p, bn, grid, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- depthwise conv
const float* fptr(const c10::optional<at::Tensor>& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }
float* fptr_mut(c10::optional<at::Tensor>& t) { return t.has_value() ? t->data_ptr<float>() : nullptr; }
const __nv_bfloat16* bptr(const at::Tensor& t) { return reinterpret_cast<const __nv_bfloat16*>(t.data_ptr()); }
__nv_bfloat16* bptr_mut(at::Tensor& t) { return reinterpret_cast<__nv_bfloat16*>(t.data_ptr()); }
DwParams dw_params(const at::Tensor& x_like, int64_t k, int64_t stride, int64_t pad, int P, int Q) {
  DwParams p{};
  p.N = x_like.size(0); p.H = x_like.size(1); p.W = x_like.size(2); p.C = x_like.size(3);
  p.P = P; p.Q = Q; p.k = k; p.stride = stride; p.pad = pad;
  TORCH_CHECK(p.C % 8 == 0 && k <= 5, "depthwise kernels need C % 8 == 0 and k <= 5");
  return p;
}
void dw_fprop(const at::Tensor& x, const at::Tensor& w, at::Tensor& y, const c10::optional<at::Tensor>& stats, int64_t k,
              int64_t stride, int64_t pad) {
  check_bf16_contig(x, "x"); check_bf16_contig(w, "w"); check_bf16_contig(y, "y");
  c10::cuda::CUDAGuard guard(x.device());
  DwParams p = dw_params(x, k, stride, pad, y.size(1), y.size(2));
  p.x = bptr(x); p.w = bptr(w); p.y = bptr_mut(y);
  p.stats = stats.has_value() ? stats->data_ptr<float>() : nullptr;
  B200_CUDA_OK(b200_dw_fprop(&p, cur_stream()));
}
void dw_dgrad(const at::Tensor& dy, const at::Tensor& w, at::Tensor& dx, int64_t k, int64_t stride, int64_t pad) {
  check_bf16_contig(dy, "dy"); check_bf16_contig(w, "w"); check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dy.device());
  DwParams p = dw_params(dx, k, stride, pad, dy.size(1), dy.size(2));
  p.w = bptr(w); p.y = const_cast<__nv_bfloat16*>(bptr(dy)); p.dx = bptr_mut(dx);
  B200_CUDA_OK(b200_dw_dgrad(&p, cur_stream()));
}
void dw_wgrad(const at::Tensor& dy, const at::Tensor& x, at::Tensor& dw, int64_t k, int64_t stride, int64_t pad) {
  check_bf16_contig(dy, "dy"); check_bf16_contig(x, "x");
  TORCH_CHECK(dw.scalar_type() == at::kFloat && dw.is_contiguous(), "dw must be contiguous fp32");
  c10::cuda::CUDAGuard guard(dy.device());
  DwParams p = dw_params(x, k, stride, pad, dy.size(1), dy.size(2));
  p.x = bptr(x); p.y = const_cast<__nv_bfloat16*>(bptr(dy)); p.dw = dw.data_ptr<float>();
  B200_CUDA_OK(b200_dw_wgrad(&p, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- peer contexts
struct PeerState {
  int world = 1, rank = 0, slot_base = 0;
  std::vector<int64_t> signal_pads, sym_bufs;
  int64_t ticket = 0, mc_stats = 0, reduced = 0, ready = 0, wait_ns = 0, epoch_dev = 0;
  uint32_t epoch = 0;
  // make(): context of the NEXT exchange (epoch + 1); current(): the exchange a producer kernel already announced
  PeerCtx current() { return make(false); }
  PeerCtx make(bool advance = true) {
    PeerCtx c{};
    c.world = world; c.rank = rank; c.slot_base = slot_base;
    if (world > 1) {
      TORCH_CHECK((int)signal_pads.size() == world && (int)sym_bufs.size() == world && world <= kMaxPeers, "bad peer state");
      for (int i = 0; i < world; ++i) {
        c.signal_pads[i] = reinterpret_cast<uint32_t*>(signal_pads[i]);
        c.sym_bufs[i] = reinterpret_cast<float*>(sym_bufs[i]);
      }
      c.epoch = advance ? ++epoch : epoch;       // informational only: the kernels use the device-side counter
      TORCH_CHECK(epoch_dev != 0, "PeerState needs the device-side exchange counter (epoch_dev)");
      c.epoch_dev = reinterpret_cast<uint32_t*>(epoch_dev);
    }
    c.ticket = reinterpret_cast<int*>(ticket);
    c.presignaled = 0;
    c.mc_stats = reinterpret_cast<float*>(mc_stats);
    c.reduced = reinterpret_cast<float*>(reduced);
    c.ready = reinterpret_cast<uint32_t*>(ready);
    c.wait_ns = reinterpret_cast<unsigned long long*>(wait_ns);
    if (world > 1) TORCH_CHECK(reduced != 0 && ready != 0, "PeerState needs the `reduced` scratch and the `ready` flag");
    return c;
  }
};


PeerCtx peer_ctx_for_producer(PeerState* peer) { return peer->make(); }

// y/out/residual are [rows, C] views (last dim contiguous, arbitrary row pitch).
void bn_apply(const at::Tensor& y, const c10::optional<at::Tensor>& residual, at::Tensor& out, const at::Tensor& stats,
              int64_t sym_offset, const c10::optional<at::Tensor>& gamma, const c10::optional<at::Tensor>& beta,
              c10::optional<at::Tensor> running_mean, c10::optional<at::Tensor> running_var, at::Tensor& save_mean,
              at::Tensor& save_invstd, double count, double eps, double momentum, int64_t act, bool training,
              PeerState* peer, c10::optional<at::Tensor> relu_mask, bool presignaled) {
  c10::cuda::CUDAGuard guard(y.device());
  TORCH_CHECK(y.dim() == 2 && y.stride(1) == 1 && out.stride(1) == 1 && y.size(1) % 8 == 0, "bn_apply expects [rows, C] with C % 8 == 0");
  BnApplyParams p{};
  p.y = bptr(y); p.ldy = y.stride(0);
  p.residual = residual.has_value() ? bptr(*residual) : nullptr; p.ldr = residual.has_value() ? residual->stride(0) : 0;
  p.out = bptr_mut(out); p.ldo = out.stride(0);
  p.rows = y.size(0); p.C = y.size(1);
  p.stats = stats.data_ptr<float>(); p.sym_offset = sym_offset;
  p.gamma = fptr(gamma); p.beta = fptr(beta);
  p.running_mean = fptr_mut(running_mean); p.running_var = fptr_mut(running_var);
  p.save_mean = save_mean.data_ptr<float>(); p.save_invstd = save_invstd.data_ptr<float>();
  p.count = (float)count; p.eps = (float)eps; p.momentum = (float)momentum; p.act = act; p.training = training;
  p.relu_mask = nullptr;
  if (relu_mask.has_value()) {
    TORCH_CHECK(relu_mask->scalar_type() == at::kByte && relu_mask->is_contiguous() && relu_mask->numel() == y.size(0) * (y.size(1) / 8) && act == ACT_RELU, "relu_mask: uint8 [rows, C/8], relu only");
    p.relu_mask = relu_mask->data_ptr<uint8_t>();
  }
  // presignaled: the conv that produced `stats` already raised this exchange's flag at its tail (conv_fprop(peer=...))
  if (peer && training) { p.peer = presignaled ? peer->current() : peer->make(); p.peer.presignaled = presignaled ? 1 : 0; }
  else { p.peer = PeerCtx{}; p.peer.world = 1; }
  B200_CUDA_OK(b200_bn_apply(&p, cur_stream()));
}

void bn_stats(const at::Tensor& y, at::Tensor& stats) {
  c10::cuda::CUDAGuard guard(y.device());
  TORCH_CHECK(y.dim() == 2 && y.stride(1) == 1 && y.size(1) % 8 == 0, "bn_stats expects [rows, C]");
  B200_CUDA_OK(b200_bn_stats(y.data_ptr(), y.size(0), y.size(1), y.stride(0), stats.data_ptr<float>(), cur_stream()));
}

void bn_backward(const at::Tensor& y, const at::Tensor& dout, const c10::optional<at::Tensor>& residual, at::Tensor& dy,
                 c10::optional<at::Tensor> dresidual, at::Tensor& sums, int64_t sym_offset,
                 const c10::optional<at::Tensor>& gamma, const c10::optional<at::Tensor>& beta, const at::Tensor& save_mean,
                 const at::Tensor& save_invstd, c10::optional<at::Tensor> dgamma, c10::optional<at::Tensor> dbeta,
                 double count, int64_t act, PeerState* peer, const c10::optional<at::Tensor>& relu_mask, int64_t phase) {
  // phase: 3 = reduce + apply (normal); 1 / 2 = only that pass (tools/bench_bn.py times them separately)
  c10::cuda::CUDAGuard guard(y.device());
  TORCH_CHECK(y.dim() == 2 && y.stride(1) == 1 && dout.stride(1) == 1 && dy.stride(1) == 1, "bn_backward expects [rows, C] views");
  BnBwdParams p{};
  p.y = bptr(y); p.ldy = y.stride(0);
  p.dout = bptr(dout); p.ldd = dout.stride(0);
  p.residual = residual.has_value() ? bptr(*residual) : nullptr; p.ldr = residual.has_value() ? residual->stride(0) : 0;
  p.dy = bptr_mut(dy); p.lddy = dy.stride(0);
  p.dresidual = dresidual.has_value() ? bptr_mut(*dresidual) : nullptr;
  if (dresidual.has_value()) { TORCH_CHECK(!residual.has_value() || dresidual->stride(0) == residual->stride(0), "dresidual pitch"); p.ldr = dresidual->stride(0); }
  p.rows = y.size(0); p.C = y.size(1);
  p.sums = sums.data_ptr<float>(); p.sym_offset = sym_offset;
  p.gamma = fptr(gamma); p.beta = fptr(beta);
  p.save_mean = save_mean.data_ptr<float>(); p.save_invstd = save_invstd.data_ptr<float>();
  p.dgamma = fptr_mut(dgamma); p.dbeta = fptr_mut(dbeta);
  p.count = (float)count; p.act = act;
  p.relu_mask = relu_mask.has_value() ? relu_mask->data_ptr<uint8_t>() : nullptr;
  // SyncBN: the reduce pass announces the exchange at its tail (last CTA out), the apply pass only waits for the peers.
  // A phase-2-only call continues the exchange its phase-1 call opened (the engine launches a weight-gradient GEMM
  // in between, which hides the NVLink round trip and the inter-rank skew).
  p.peer = PeerCtx{}; p.peer.world = 1;
  if (phase & 1) {
    if (peer) p.peer = peer->make();
    B200_CUDA_OK(b200_bn_bwd_reduce(&p, cur_stream()));
  }
  if (phase & 2) {
    if (peer) { p.peer = peer->current(); p.peer.presignaled = 1; }
    B200_CUDA_OK(b200_bn_bwd_apply(&p, cur_stream()));
  }
}

// Stem tail (elementwise.h: BnPoolParams): y [N,H,W,C] -> BN -> ReLU -> max-pool 3x3/2/1 -> out [N,H/2,W/2,C] in one pass.
static void bn_pool_geometry(BnPoolParams& p, const at::Tensor& y, const at::Tensor& pooled, const at::Tensor* arg) {
  check_bf16_contig(y, "y"); check_bf16_contig(pooled, "pooled");
  TORCH_CHECK(y.dim() == 4 && pooled.dim() == 4, "bn_relu_pool: NHWC tensors");
  p.N = y.size(0); p.H = y.size(1); p.W = y.size(2); p.C = y.size(3);
  p.P = pooled.size(1); p.Q = pooled.size(2);
  TORCH_CHECK(p.H % 2 == 0 && p.W % 2 == 0 && p.C % 8 == 0 && p.P == p.H / 2 && p.Q == p.W / 2 && pooled.size(0) == p.N && pooled.size(3) == p.C,
              "bn_relu_pool: even H/W, C % 8 == 0, pooled = [N, H/2, W/2, C]");
  TORCH_CHECK(y.numel() < (1ll << 31), "bn_relu_pool: 32-bit pixel indices");
  if (arg) TORCH_CHECK(arg->scalar_type() == at::kByte && arg->is_contiguous() && arg->numel() == pooled.numel(), "arg: uint8, pooled shape");
}
void bn_relu_pool_fwd(const at::Tensor& y, at::Tensor& out, c10::optional<at::Tensor> arg, const at::Tensor& stats, int64_t sym_offset,
                      const c10::optional<at::Tensor>& gamma, const c10::optional<at::Tensor>& beta,
                      c10::optional<at::Tensor> running_mean, c10::optional<at::Tensor> running_var, at::Tensor& save_mean,
                      at::Tensor& save_invstd, double count, double eps, double momentum, bool training, PeerState* peer,
                      bool presignaled) {
  c10::cuda::CUDAGuard guard(y.device());
  BnPoolParams p{};
  bn_pool_geometry(p, y, out, arg.has_value() ? &*arg : nullptr);
  p.y = bptr(y); p.out = bptr_mut(out); p.arg = arg.has_value() ? arg->data_ptr<uint8_t>() : nullptr;
  p.stats = stats.data_ptr<float>(); p.sym_offset = sym_offset;
  p.gamma = fptr(gamma); p.beta = fptr(beta);
  p.running_mean = fptr_mut(running_mean); p.running_var = fptr_mut(running_var);
  TORCH_CHECK(training || (p.running_mean && p.running_var), "bn_relu_pool_fwd: eval mode needs running statistics");
  p.save_mean = save_mean.data_ptr<float>(); p.save_invstd = save_invstd.data_ptr<float>();
  p.count = (float)count; p.eps = (float)eps; p.momentum = (float)momentum; p.training = training;
  if (peer && training) { p.peer = presignaled ? peer->current() : peer->make(); p.peer.presignaled = presignaled ? 1 : 0; }
  else { p.peer = PeerCtx{}; p.peer.world = 1; }
  B200_CUDA_OK(b200_bn_relu_pool_fwd(&p, cur_stream()));
}
// phase: 1 = reduce pass (opens the SyncBN exchange at its tail), 2 = apply pass, 3 = both
void bn_relu_pool_bwd(const at::Tensor& y, const at::Tensor& dout, const at::Tensor& arg, at::Tensor& dy, at::Tensor& sums,
                      int64_t sym_offset, const c10::optional<at::Tensor>& gamma, const at::Tensor& save_mean,
                      const at::Tensor& save_invstd, c10::optional<at::Tensor> dgamma, c10::optional<at::Tensor> dbeta,
                      double count, PeerState* peer, int64_t phase) {
  c10::cuda::CUDAGuard guard(y.device());
  BnPoolParams p{};
  bn_pool_geometry(p, y, dout, &arg);
  check_bf16_contig(dy, "dy");
  TORCH_CHECK(dy.numel() == y.numel(), "bn_relu_pool_bwd: dy has the shape of y");
  p.y = bptr(y); p.dout = bptr(dout); p.arg = const_cast<uint8_t*>(arg.data_ptr<uint8_t>()); p.dy = bptr_mut(dy);
  p.stats = sums.data_ptr<float>(); p.sym_offset = sym_offset;
  p.gamma = fptr(gamma);
  p.save_mean = save_mean.data_ptr<float>(); p.save_invstd = save_invstd.data_ptr<float>();
  p.dgamma = fptr_mut(dgamma); p.dbeta = fptr_mut(dbeta);
  p.count = (float)count; p.training = 1;
  p.peer = PeerCtx{}; p.peer.world = 1;
  if (phase & 1) {
    if (peer) p.peer = peer->make();
    B200_CUDA_OK(b200_bn_relu_pool_bwd_reduce(&p, cur_stream()));
  }
  if (phase & 2) {
    if (peer) { p.peer = peer->current(); p.peer.presignaled = 1; }
    B200_CUDA_OK(b200_bn_relu_pool_bwd_apply(&p, cur_stream()));
  }
}

void maxpool_fwd(const at::Tensor& x, at::Tensor& out, c10::optional<at::Tensor> argmax, int64_t k, int64_t stride, int64_t pad) {
  check_bf16_contig(x, "x"); check_bf16_contig(out, "out");
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.size(3) % 8 == 0 && k * k <= 255, "maxpool: C % 8 == 0");
  B200_CUDA_OK(b200_maxpool_fwd(x.data_ptr(), out.data_ptr(), argmax.has_value() ? argmax->data_ptr() : nullptr, x.size(0),
                                x.size(1), x.size(2), x.size(3), out.size(1), out.size(2), k, stride, pad, cur_stream()));
}
void maxpool_bwd(const at::Tensor& dout, const at::Tensor& argmax, at::Tensor& dx, int64_t k, int64_t stride, int64_t pad) {
  check_bf16_contig(dout, "dout"); check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dout.device());
  B200_CUDA_OK(b200_maxpool_bwd(dout.data_ptr(), argmax.data_ptr(), dx.data_ptr(), dx.size(0), dx.size(1), dx.size(2),
                                dx.size(3), dout.size(1), dout.size(2), k, stride, pad, cur_stream()));
}
void gap_fwd(const at::Tensor& x, at::Tensor& out) {
  check_bf16_contig(x, "x"); check_bf16_contig(out, "out");
  c10::cuda::CUDAGuard guard(x.device());
  B200_CUDA_OK(b200_gap_fwd(x.data_ptr(), out.data_ptr(), x.size(0), x.size(1) * x.size(2), x.size(3), cur_stream()));
}
void gap_bwd(const at::Tensor& dout, at::Tensor& dx) {
  check_bf16_contig(dout, "dout"); check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dout.device());
  B200_CUDA_OK(b200_gap_bwd(dout.data_ptr(), dx.data_ptr(), dx.size(0), dx.size(1) * dx.size(2), dx.size(3), cur_stream()));
}
void avgpool2_fwd(const at::Tensor& x, at::Tensor& out) {
  check_bf16_contig(x, "x"); check_bf16_contig(out, "out");
  c10::cuda::CUDAGuard guard(x.device());
  B200_CUDA_OK(b200_avgpool2_fwd(x.data_ptr(), out.data_ptr(), x.size(0), x.size(1), x.size(2), x.size(3), cur_stream()));
}
void avgpool2_bwd(const at::Tensor& dout, at::Tensor& dx) {
  check_bf16_contig(dout, "dout"); check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dout.device());
  B200_CUDA_OK(b200_avgpool2_bwd(dout.data_ptr(), dx.data_ptr(), dx.size(0), dx.size(1), dx.size(2), dx.size(3), cur_stream()));
}

// x/out/dout/dx: [N,H,W,C] bf16; gate: [N,C] bf16; dgate: [N,C] fp32 (accumulated).
void channel_scale_fwd(const at::Tensor& x, const at::Tensor& gate, at::Tensor& out) {
  check_bf16_contig(x, "x"); check_bf16_contig(gate, "gate"); check_bf16_contig(out, "out");
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.size(3) % 8 == 0 && gate.size(0) == x.size(0) && gate.size(1) == x.size(3), "channel_scale shapes");
  B200_CUDA_OK(b200_channel_scale_fwd(x.data_ptr(), gate.data_ptr(), out.data_ptr(), x.size(0), x.size(1) * x.size(2), x.size(3), cur_stream()));
}
void channel_scale_bwd(const at::Tensor& dout, const at::Tensor& x, const at::Tensor& gate, at::Tensor& dx, at::Tensor& dgate) {
  check_bf16_contig(dout, "dout"); check_bf16_contig(x, "x"); check_bf16_contig(gate, "gate"); check_bf16_contig(dx, "dx");
  TORCH_CHECK(dgate.scalar_type() == at::kFloat && dgate.is_contiguous(), "dgate must be contiguous fp32");
  c10::cuda::CUDAGuard guard(x.device());
  B200_CUDA_OK(b200_channel_scale_bwd(dout.data_ptr(), x.data_ptr(), gate.data_ptr(), dx.data_ptr(), dgate.data_ptr<float>(),
                                      x.size(0), x.size(1) * x.size(2), x.size(3), cur_stream()));
}

// accum: fp32 [3] = (sum of per-sample losses, top-1 hits, top-k hits), accumulated.
void ce_topk(const at::Tensor& logits, const at::Tensor& target, c10::optional<at::Tensor> dlogits, at::Tensor& accum,
             int64_t topk, double grad_scale) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(logits.scalar_type() == at::kBFloat16 && logits.dim() == 2 && logits.stride(1) == 1, "logits: bf16 [B, classes]");
  TORCH_CHECK(target.scalar_type() == at::kLong && target.is_contiguous(), "target: int64");
  if (dlogits.has_value()) TORCH_CHECK(dlogits->stride(0) == logits.stride(0) && dlogits->scalar_type() == at::kBFloat16, "dlogits layout");
  B200_CUDA_OK(b200_ce_topk(logits.data_ptr(), (const long long*)target.data_ptr<int64_t>(),
                            dlogits.has_value() ? dlogits->data_ptr() : nullptr, accum.data_ptr<float>(), logits.size(0),
                            logits.size(1), logits.stride(0), topk, (float)grad_scale, cur_stream()));
}

void nchw_to_nhwc(const at::Tensor& x, at::Tensor& out) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.scalar_type() == at::kFloat && x.is_contiguous() && out.scalar_type() == at::kBFloat16, "nchw_to_nhwc: fp32 -> bf16");
  B200_CUDA_OK(b200_nchw_to_nhwc(x.data_ptr<float>(), out.data_ptr(), x.size(0), x.size(1), x.size(2), x.size(3), cur_stream()));
}
void stem_im2col(const at::Tensor& x, at::Tensor& patches, int64_t R, int64_t S, int64_t stride, int64_t pad, int64_t P, int64_t Q,
                 const std::vector<double>& mean, const std::vector<double>& std_) {
  // x: NCHW fp32 (already normalised) or NCHW uint8 (raw pixels, normalised here with mean/std given in [0,1] units)
  c10::cuda::CUDAGuard guard(x.device());
  const bool u8 = x.scalar_type() == at::kByte;
  TORCH_CHECK((u8 || x.scalar_type() == at::kFloat) && x.is_contiguous() && patches.scalar_type() == at::kBFloat16 && patches.is_contiguous(),
              "stem_im2col dtypes");
  StemNorm norm{};
  for (int c = 0; c < 8; ++c) { norm.scale[c] = 1.f; norm.bias[c] = 0.f; }
  if (u8) {
    TORCH_CHECK((int64_t)mean.size() == x.size(1) && (int64_t)std_.size() == x.size(1), "stem_im2col: uint8 input needs per-channel mean/std");
    for (size_t c = 0; c < mean.size() && c < 8; ++c) { norm.scale[c] = (float)(1.0 / (255.0 * std_[c])); norm.bias[c] = (float)(-mean[c] / std_[c]); }
  }
  B200_CUDA_OK(b200_stem_im2col(x.data_ptr(), u8 ? 1 : 0, patches.data_ptr(), x.size(0), x.size(1), x.size(2), x.size(3), P, Q, R, S,
                                stride, pad, patches.size(-1), &norm, cur_stream()));
}
void pad_rows(const at::Tensor& src, at::Tensor& dst, int64_t rows, int64_t cols, int64_t cols_pad) {
  c10::cuda::CUDAGuard guard(src.device());
  B200_CUDA_OK(b200_pad_rows(src.data_ptr(), dst.data_ptr(), rows, cols, cols_pad, cur_stream()));
}
void unpad_add(const at::Tensor& src, at::Tensor& dst, int64_t rows, int64_t cols, int64_t cols_pad) {
  c10::cuda::CUDAGuard guard(src.device());
  B200_CUDA_OK(b200_unpad_add(src.data_ptr<float>(), dst.data_ptr<float>(), rows, cols, cols_pad, cur_stream()));
}

// Space-to-depth stem (extras.cu): x [N,3,H,W] fp32 / uint8 -> out [N,Hs,Ws,16] bf16 with out[n,i,j,(u*2+v)*3+c] = x[n,c,2i+u-3,2j+v-3]
void stem_s2d(const at::Tensor& x, at::Tensor& out, const std::vector<double>& mean, const std::vector<double>& std_) {
  c10::cuda::CUDAGuard guard(x.device());
  const bool u8 = x.scalar_type() == at::kByte;
  TORCH_CHECK(x.is_cuda() && (u8 || x.scalar_type() == at::kFloat) && x.is_contiguous() && x.dim() == 4 && x.size(1) == 3, "stem_s2d: [N,3,H,W] fp32 / uint8");
  check_bf16_contig(out, "out");
  TORCH_CHECK(out.dim() == 4 && out.size(0) == x.size(0) && out.size(3) == 16, "stem_s2d: out [N,Hs,Ws,16]");
  float sc[3] = {1.f, 1.f, 1.f}, bi[3] = {0.f, 0.f, 0.f};
  if (u8) {
    TORCH_CHECK(mean.size() == 3 && std_.size() == 3, "stem_s2d: uint8 input needs per-channel mean/std");
    for (int c = 0; c < 3; ++c) { sc[c] = (float)(1.0 / (255.0 * std_[c])); bi[c] = (float)(-mean[c] / std_[c]); }
  }
  B200_CUDA_OK(b200_stem_s2d(x.data_ptr(), u8 ? 1 : 0, out.data_ptr(), x.size(0), x.size(2), x.size(3), out.size(1), out.size(2), sc, bi, cur_stream()));
}
void stem_s2d_pack_w(const at::Tensor& w, at::Tensor& wp) {
  check_bf16_contig(w, "w"); check_bf16_contig(wp, "wp");
  c10::cuda::CUDAGuard guard(w.device());
  TORCH_CHECK(w.numel() % 147 == 0 && wp.numel() == w.numel() / 147 * 256, "stem_s2d_pack_w: w [K,7,7,3] -> wp [K,4,1,64]");
  B200_CUDA_OK(b200_stem_s2d_pack_w(w.data_ptr(), wp.data_ptr(), w.numel() / 147, cur_stream()));
}
void stem_s2d_unpack_dw(const at::Tensor& dwp, at::Tensor& dw) {
  c10::cuda::CUDAGuard guard(dwp.device());
  TORCH_CHECK(dwp.scalar_type() == at::kFloat && dw.scalar_type() == at::kFloat && dwp.is_contiguous() && dw.is_contiguous() &&
              dw.numel() % 147 == 0 && dwp.numel() == dw.numel() / 147 * 256, "stem_s2d_unpack_dw: dwp [K,4,1,64] fp32 -> dw [K,7,7,3] fp32");
  B200_CUDA_OK(b200_stem_s2d_unpack_dw(dwp.data_ptr<float>(), dw.data_ptr<float>(), dw.numel() / 147, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- extras
void colsum_add(const at::Tensor& d, at::Tensor& out) {
  c10::cuda::CUDAGuard guard(d.device());
  TORCH_CHECK(d.dim() == 2 && d.stride(1) == 1 && d.scalar_type() == at::kBFloat16 && d.size(1) % 8 == 0, "colsum_add: bf16 [rows, C], C % 8 == 0");
  TORCH_CHECK(out.scalar_type() == at::kFloat && out.is_contiguous() && out.numel() >= d.size(1), "colsum_add: fp32 [C] output");
  B200_CUDA_OK(b200_colsum_add(d.data_ptr(), d.size(0), d.size(1), d.stride(0), out.data_ptr<float>(), cur_stream()));
}
void strided_add_inplace(at::Tensor& dx, const at::Tensor& compact, int64_t stride) {
  check_bf16_contig(dx, "dx"); check_bf16_contig(compact, "compact");
  c10::cuda::CUDAGuard guard(dx.device());
  TORCH_CHECK(dx.size(3) == compact.size(3) && dx.size(3) % 8 == 0 && (compact.size(1) - 1) * stride < dx.size(1) &&
              (compact.size(2) - 1) * stride < dx.size(2), "strided_add_inplace shapes");
  B200_CUDA_OK(b200_strided_add_inplace(dx.data_ptr(), compact.data_ptr(), dx.size(0), dx.size(1), dx.size(2), dx.size(3), compact.size(1),
                                        compact.size(2), stride, cur_stream()));
}
void blockdiag_pack(const at::Tensor& thin, at::Tensor& dense) {
  check_bf16_contig(thin, "thin"); check_bf16_contig(dense, "dense");
  c10::cuda::CUDAGuard guard(thin.device());
  const int K = thin.size(0), taps = thin.size(1) * thin.size(2), cg = thin.size(3);
  TORCH_CHECK(64 % cg == 0 && K % 64 == 0 && dense.numel() == (int64_t)K * taps * 64, "blockdiag_pack: cg must divide 64, K % 64 == 0");
  B200_CUDA_OK(b200_blockdiag_pack(thin.data_ptr(), dense.data_ptr(), K, taps, cg, cur_stream()));
}
void blockdiag_unpack_add(const at::Tensor& dense, at::Tensor& thin) {
  c10::cuda::CUDAGuard guard(thin.device());
  TORCH_CHECK(dense.scalar_type() == at::kFloat && thin.scalar_type() == at::kFloat && dense.is_contiguous() && thin.is_contiguous(), "fp32 contiguous");
  const int K = thin.size(0), taps = thin.size(1) * thin.size(2), cg = thin.size(3);
  TORCH_CHECK(64 % cg == 0 && dense.numel() == (int64_t)K * taps * 64, "blockdiag_unpack_add shapes");
  B200_CUDA_OK(b200_blockdiag_unpack_add(dense.data_ptr<float>(), thin.data_ptr<float>(), K, taps, cg, cur_stream()));
}
// s [N,C] bf16 pooled input, w1 [r,C] bf16, w2 [C,r] bf16, biases fp32 (optional) -> pre1 [N,r] fp32, gate [N,C] bf16
void se_gate_fwd(const at::Tensor& sp, const at::Tensor& w1, const c10::optional<at::Tensor>& b1, const at::Tensor& w2,
                 const c10::optional<at::Tensor>& b2, at::Tensor& pre1, at::Tensor& gate, int64_t act) {
  check_bf16_contig(sp, "s"); check_bf16_contig(w1, "w1"); check_bf16_contig(w2, "w2"); check_bf16_contig(gate, "gate");
  c10::cuda::CUDAGuard guard(sp.device());
  const int N = sp.size(0), C = sp.size(1), r = w1.numel() / C;
  TORCH_CHECK(w2.numel() == (int64_t)C * r && pre1.numel() == (int64_t)N * r && pre1.scalar_type() == at::kFloat && gate.numel() == sp.numel(), "se_gate_fwd shapes");
  TORCH_CHECK((size_t)4 * (C + r) * 4 <= 200 * 1024, "se_gate_fwd: layer too wide for the shared-memory staging");
  B200_CUDA_OK(b200_se_gate_fwd(sp.data_ptr(), w1.data_ptr(), fptr(b1), w2.data_ptr(), fptr(b2), pre1.data_ptr<float>(), gate.data_ptr(), N, C, r,
                                act, cur_stream()));
}
void se_gate_bwd(const at::Tensor& dgate, const at::Tensor& gate, const at::Tensor& sp, const at::Tensor& pre1, const at::Tensor& w1,
                 const at::Tensor& w2, at::Tensor& dw1, c10::optional<at::Tensor> db1, at::Tensor& dw2, c10::optional<at::Tensor> db2,
                 at::Tensor& ds, at::Tensor& scratch, int64_t act) {
  c10::cuda::CUDAGuard guard(sp.device());
  const int N = sp.size(0), C = sp.size(1), r = w1.numel() / C;
  TORCH_CHECK(dgate.scalar_type() == at::kFloat && dgate.is_contiguous() && ds.scalar_type() == at::kFloat && ds.numel() == (int64_t)N * C, "se_gate_bwd: fp32 dgate / ds");
  TORCH_CHECK(dw1.scalar_type() == at::kFloat && dw2.scalar_type() == at::kFloat && dw1.numel() == (int64_t)C * r && dw2.numel() == (int64_t)C * r, "se_gate_bwd: fp32 weight gradients");
  TORCH_CHECK(scratch.scalar_type() == at::kFloat && scratch.numel() >= (int64_t)N * C + 2LL * N * r, "se_gate_bwd: scratch too small");
  B200_CUDA_OK(b200_se_gate_bwd(dgate.data_ptr<float>(), gate.data_ptr(), sp.data_ptr(), pre1.data_ptr<float>(), w1.data_ptr(), w2.data_ptr(),
                                dw1.data_ptr<float>(), fptr_mut(db1), dw2.data_ptr<float>(), fptr_mut(db2), ds.data_ptr<float>(),
                                scratch.data_ptr<float>(), N, C, r, act, cur_stream()));
}
void channel_add_bcast(at::Tensor& dx, const at::Tensor& ds, double scale) {
  check_bf16_contig(dx, "dx");
  c10::cuda::CUDAGuard guard(dx.device());
  TORCH_CHECK(ds.scalar_type() == at::kFloat && ds.is_contiguous() && ds.numel() == dx.size(0) * dx.size(3) && dx.size(3) % 8 == 0, "channel_add_bcast shapes");
  B200_CUDA_OK(b200_channel_add_bcast(dx.data_ptr(), ds.data_ptr<float>(), dx.size(0), dx.size(1) * dx.size(2), dx.size(3), (float)scale, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- 3x3 halo conv (64 -> 64)
// x [N,H,W,64], w [64,3,3,64], y [N,H,W,64]; dgrad: x = dy, y = dx (same weight tensor, read MN-major with flipped taps)
void conv3x3_halo(const at::Tensor& x, const at::Tensor& w, at::Tensor& y, const c10::optional<at::Tensor>& stats, bool dgrad,
                  PeerState* peer) {
  check_bf16_contig(x, "x"); check_bf16_contig(w, "w"); check_bf16_contig(y, "y");
  c10::cuda::CUDAGuard guard(x.device());
  const int N = x.size(0), H = x.size(1), W = x.size(2);
  TORCH_CHECK(x.size(3) == 64 && w.size(0) == 64 && w.size(1) == 3 && w.size(2) == 3 && w.size(3) == 64 && y.numel() == x.numel(),
              "conv3x3_halo handles 64 -> 64 channel 3x3 / stride 1 / pad 1 convolutions");
  const int Wp = W + 2;
  TORCH_CHECK(Wp <= 64, "conv3x3_halo: feature map too wide (W <= 62)");
  const int rpt = 128 / Wp;                 // image rows per tile
  const int halo_rows = (rpt + 2) * Wp;     // <= 256 for Wp <= 64
  TORCH_CHECK(halo_rows <= 256 && 2 * Wp + 2 + 128 <= halo_rows + 128, "conv3x3_halo geometry");
  Conv3x3HaloParams p{};
  p.N = N; p.H = H; p.W = W;
  p.rpt = rpt;
  p.tiles_per_img = (H + rpt - 1) / rpt;
  p.tiles = N * p.tiles_per_img;
  p.halo_rows = halo_rows;
  p.dgrad = dgrad ? 1 : 0;
  p.y = y.data_ptr();
  p.stats = stats.has_value() ? stats->data_ptr<float>() : nullptr;
  if (stats.has_value()) TORCH_CHECK(stats->numel() >= 128 && stats->scalar_type() == at::kFloat, "stats must be fp32 [2*64]");
  p.peer = PeerCtx{}; p.peer.world = 1;
  if (peer != nullptr) { TORCH_CHECK(stats.has_value() && !dgrad, "conv3x3_halo: a peer context needs the statistics epilogue"); p.peer = peer_ctx_for_producer(peer); }
  CUtensorMap mx = tiled_map_4d(x.data_ptr(), 64, W, H, N, 64, (uint64_t)W * 64, (uint64_t)H * W * 64, 64, (uint32_t)Wp, (uint32_t)(rpt + 2), 1);
  CUtensorMap mw = tiled_map_3d(w.data_ptr(), 64, 9, 64, 64, 9 * 64, 64, 1, 64);
  CUtensorMap my = tiled_map_4d(y.data_ptr(), 64, W, H, N, 64, (uint64_t)W * 64, (uint64_t)H * W * 64, 64, 32, 1, 1);
  const int grid = std::min(p.tiles, num_sms());
  B200_CUDA_OK(b200_conv3x3_halo_launch(&mx, &mw, &my, &p, grid, cur_stream()));
}

// Space-to-depth stem forward with a cp.async-gathered A tile (stem_conv.cu): xs [N,P+3,Q+3,16], w2 [64,4,1,64] -> y [N,P,Q,64]
void stem_conv_fprop(const at::Tensor& xs, const at::Tensor& w2, at::Tensor& y, const c10::optional<at::Tensor>& stats, PeerState* peer) {
  check_bf16_contig(xs, "xs"); check_bf16_contig(w2, "w2"); check_bf16_contig(y, "y");
  c10::cuda::CUDAGuard guard(xs.device());
  TORCH_CHECK(xs.dim() == 4 && xs.size(3) == 16 && y.dim() == 4 && y.size(3) == 64 && w2.numel() == 64 * 256, "stem_conv_fprop: xs [N,Hs,Ws,16], w2 [64,4,1,64], y [N,P,Q,64]");
  StemConvParams p{};
  p.N = xs.size(0); p.Hs = xs.size(1); p.Ws = xs.size(2); p.P = y.size(1); p.Q = y.size(2);
  TORCH_CHECK(y.size(0) == p.N && p.Hs == p.P + 3 && p.Ws == p.Q + 3 && p.Q <= 128, "stem_conv_fprop geometry (Q <= 128)");
  p.tiles = p.N * p.P;
  p.s = xs.data_ptr(); p.y = y.data_ptr();
  p.stats = stats.has_value() ? stats->data_ptr<float>() : nullptr;
  if (stats.has_value()) TORCH_CHECK(stats->numel() >= 128 && stats->scalar_type() == at::kFloat, "stats must be fp32 [2*64]");
  p.peer = PeerCtx{}; p.peer.world = 1;
  if (peer != nullptr) { TORCH_CHECK(stats.has_value(), "stem_conv_fprop: a peer context needs the statistics epilogue"); p.peer = peer_ctx_for_producer(peer); }
  CUtensorMap mw = tiled_map_3d(w2.data_ptr(), 64, 4, 64, 64, 4 * 64, 64, 1, 64);
  const int grid = std::min(p.tiles, num_sms());
  B200_CUDA_OK(b200_stem_conv_launch(&mw, &p, grid, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- attention (BoTNet MHSA)
// qk [B,14,14,2*heads*128], v / out [B,14,14,heads*128] (NHWC bf16, contiguous); rel_w / rel_h [27,128] bf16
struct AttnMaps { CUtensorMap qk128, qk64, v128, v64, relw, relh; };
AttnMaps attn_maps(const at::Tensor& qk, const at::Tensor& v, const at::Tensor& relw, const at::Tensor& relh, int heads) {
  check_bf16_contig(qk, "qk"); check_bf16_contig(v, "v"); check_bf16_contig(relw, "rel_w"); check_bf16_contig(relh, "rel_h");
  const int B = qk.size(0);
  TORCH_CHECK(qk.dim() == 4 && qk.size(1) * qk.size(2) == 196 && qk.size(3) == 2 * heads * 128 && v.size(3) == heads * 128 && v.size(0) == B,
              "fused MHSA handles 14x14 maps with 128-wide heads");
  TORCH_CHECK(relw.size(0) == 27 && relw.size(1) == 128 && relh.size(0) == 27 && relh.size(1) == 128, "relative tables must be [27,128]");
  const uint64_t C2 = 2 * heads * 128, Cv = heads * 128;
  AttnMaps m;
  m.qk128 = tiled_map_3d(qk.data_ptr(), C2, 196, B, C2, 196 * C2, 64, 128, 1);
  m.qk64 = tiled_map_3d(qk.data_ptr(), C2, 196, B, C2, 196 * C2, 64, 64, 1);
  m.v128 = tiled_map_3d(v.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 128, 1);
  m.v64 = tiled_map_3d(v.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 64, 1);
  m.relw = tiled_map_3d(relw.data_ptr(), 128, 27, 1, 128, 27 * 128, 64, 32, 1);
  m.relh = tiled_map_3d(relh.data_ptr(), 128, 27, 1, 128, 27 * 128, 64, 32, 1);
  return m;
}
void attn_fwd(const at::Tensor& qk, const at::Tensor& v, const at::Tensor& relw, const at::Tensor& relh, at::Tensor& out,
              at::Tensor& p_save, int64_t heads, double scale) {
  c10::cuda::CUDAGuard guard(qk.device());
  AttnMaps m = attn_maps(qk, v, relw, relh, heads);
  check_bf16_contig(out, "out"); check_bf16_contig(p_save, "p_save");
  const int B = qk.size(0);
  TORCH_CHECK(out.numel() == v.numel() && p_save.numel() == (int64_t)B * heads * 196 * 208, "attn_fwd output shapes");
  const uint64_t Cv = heads * 128;
  CUtensorMap out32 = tiled_map_3d(out.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 32, 1);
  AttnParams p{};
  p.B = B; p.heads = heads; p.scale = (float)scale; p.p_save = p_save.data_ptr();
  B200_CUDA_OK(b200_attn_fwd(&m.qk128, &m.v64, &m.relw, &m.relh, &out32, &p, cur_stream()));
}
void attn_bwd(const at::Tensor& dout, const at::Tensor& qk, const at::Tensor& v, const at::Tensor& relw, const at::Tensor& relh,
              const at::Tensor& p_save, at::Tensor& ds_save, at::Tensor& dsrel, at::Tensor& dqk, at::Tensor& dv, int64_t heads, double scale) {
  c10::cuda::CUDAGuard guard(qk.device());
  AttnMaps m = attn_maps(qk, v, relw, relh, heads);
  check_bf16_contig(dout, "dout"); check_bf16_contig(p_save, "p_save"); check_bf16_contig(ds_save, "ds_save");
  check_bf16_contig(dsrel, "dsrel"); check_bf16_contig(dqk, "dqk"); check_bf16_contig(dv, "dv");
  const int B = qk.size(0);
  const uint64_t C2 = 2 * heads * 128, Cv = heads * 128, BH = (uint64_t)B * heads;
  TORCH_CHECK(dout.numel() == v.numel() && dqk.numel() == qk.numel() && dv.numel() == v.numel() &&
              p_save.numel() == (int64_t)BH * 196 * 208 && ds_save.numel() == (int64_t)BH * 196 * 256 &&
              dsrel.numel() == (int64_t)B * 196 * heads * 64, "attn_bwd shapes");
  CUtensorMap dout128 = tiled_map_3d(dout.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 128, 1);
  CUtensorMap dout64 = tiled_map_3d(dout.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 64, 1);
  CUtensorMap dqk32 = tiled_map_3d(dqk.data_ptr(), C2, 196, B, C2, 196 * C2, 64, 32, 1);
  CUtensorMap dv32 = tiled_map_3d(dv.data_ptr(), Cv, 196, B, Cv, 196 * Cv, 64, 32, 1);
  CUtensorMap psave64 = tiled_map_3d(p_save.data_ptr(), 208, 196, BH, 208, 196 * 208, 64, 64, 1);
  CUtensorMap dssave64 = tiled_map_3d(ds_save.data_ptr(), 256, 196, BH, 256, 196 * 256, 64, 64, 1);
  AttnParams p{};
  p.B = B; p.heads = heads; p.scale = (float)scale;
  p.p_save = p_save.data_ptr(); p.ds_save = ds_save.data_ptr(); p.dsrel = dsrel.data_ptr();
  B200_CUDA_OK(b200_attn_bwd_dq(&dout128, &m.v128, &m.qk64, &m.relw, &m.relh, &dqk32, &p, cur_stream()));
  B200_CUDA_OK(b200_attn_bwd_dkv(&psave64, &dssave64, &dout64, &m.qk64, &dv32, &dqk32, &p, cur_stream()));
}
void rel_grad_reduce(const at::Tensor& dw, at::Tensor& grad_w, at::Tensor& grad_h, int64_t heads) {
  c10::cuda::CUDAGuard guard(dw.device());
  TORCH_CHECK(dw.scalar_type() == at::kFloat && dw.is_contiguous() && dw.numel() == heads * 64 * 128 && grad_w.numel() == 27 * 128 &&
              grad_h.numel() == 27 * 128 && grad_w.scalar_type() == at::kFloat && grad_h.scalar_type() == at::kFloat, "rel_grad_reduce shapes");
  B200_CUDA_OK(b200_rel_grad_reduce(dw.data_ptr<float>(), grad_w.data_ptr<float>(), grad_h.data_ptr<float>(), heads, cur_stream()));
}

// ---------------------------------------------------------------------------------------------- optimizer / comm
SgdHyper hyper(double lr, double momentum, double dampening, double wd, bool nesterov, bool first) {
  SgdHyper h; h.lr = lr; h.momentum = momentum; h.dampening = dampening; h.weight_decay = wd; h.nesterov = nesterov; h.first_step = first;
  return h;
}
void sgd_local(at::Tensor& master, at::Tensor& mom, at::Tensor& grad, c10::optional<at::Tensor> w16, int64_t off, int64_t n,
               double lr, double momentum, double dampening, double wd, bool nesterov, bool first, double grad_scale, bool zero_grad) {
  c10::cuda::CUDAGuard guard(master.device());
  TORCH_CHECK(off % 8 == 0 && n % 8 == 0, "flat ranges must be multiples of 8 elements");
  SgdHyper h = hyper(lr, momentum, dampening, wd, nesterov, first);
  B200_CUDA_OK(b200_sgd_local(master.data_ptr<float>() + off, mom.data_ptr<float>() + off, grad.data_ptr<float>() + off,
                              w16.has_value() ? (void*)(reinterpret_cast<__nv_bfloat16*>(w16->data_ptr()) + off) : nullptr, n, &h,
                              (float)grad_scale, zero_grad, cur_stream()));
}
void cast_bf16(const at::Tensor& src, at::Tensor& dst) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.numel() % 8 == 0 && src.numel() == dst.numel(), "cast_bf16: numel % 8");
  B200_CUDA_OK(b200_cast_bf16(src.data_ptr<float>(), dst.data_ptr(), src.numel(), cur_stream()));
}

struct CommState {
  int world = 1, rank = 0, slot_base = 0;
  std::vector<int64_t> signal_pads, stage, w16;
  int64_t mc_stage = 0, mc_w16 = 0, local_counter = 0, local_release = 0, epoch_dev = 0;
  uint32_t epoch = 0;
  CommCtx make() const {
    CommCtx c{};
    c.world = world; c.rank = rank; c.slot_base = slot_base;
    TORCH_CHECK(world <= kCommMaxPeers && (int)signal_pads.size() == world && (int)stage.size() == world && (int)w16.size() == world, "bad comm state");
    for (int i = 0; i < world; ++i) {
      c.signal_pads[i] = reinterpret_cast<uint32_t*>(signal_pads[i]);
      c.stage[i] = reinterpret_cast<__nv_bfloat16*>(stage[i]);
      c.w16[i] = reinterpret_cast<__nv_bfloat16*>(w16[i]);
    }
    c.mc_stage = reinterpret_cast<__nv_bfloat16*>(mc_stage);
    c.mc_w16 = reinterpret_cast<__nv_bfloat16*>(mc_w16);
    TORCH_CHECK(epoch_dev != 0, "CommState needs the device-side barrier counter (epoch_dev)");
    c.epoch_dev = reinterpret_cast<uint32_t*>(epoch_dev);
    c.local_counter = reinterpret_cast<int*>(local_counter);
    c.local_release = reinterpret_cast<uint32_t*>(local_release);
    return c;
  }
};

void allreduce_sgd(CommState* cs, at::Tensor& master, at::Tensor& mom, at::Tensor& grad, int64_t off, int64_t n, double lr,
                   double momentum, double dampening, double wd, bool nesterov, bool first, bool one_shot, int64_t grid) {
  c10::cuda::CUDAGuard guard(master.device());
  TORCH_CHECK(off % 8 == 0 && n % 8 == 0, "bucket ranges must be multiples of 8 elements");
  AllreduceSgdParams p{};
  p.comm = cs->make();
  p.master = master.data_ptr<float>(); p.mom = mom.data_ptr<float>(); p.grad = grad.data_ptr<float>();
  p.off8 = off / 8; p.n8 = n / 8;
  p.hyper = hyper(lr, momentum, dampening, wd, nesterov, first);
  p.epoch = cs->epoch; cs->epoch += 2;
  p.one_shot = one_shot;
  B200_CUDA_OK(b200_allreduce_sgd(&p, (int)std::max<int64_t>(1, std::min<int64_t>(grid, num_sms())), cur_stream()));
}
void rank_barrier(CommState* cs) {
  CommCtx c = cs->make();
  cs->epoch += 1;
  B200_CUDA_OK(b200_rank_barrier(&c, cs->epoch, cur_stream()));
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "distribuuuu_b200 sm_100a kernels";
  m.def("conv_fprop", &conv_fprop, "tcgen05 implicit-GEMM convolution forward (NHWC bf16)", py::arg("x"), py::arg("w"),
        py::arg("out"), py::arg("stats"), py::arg("bias"), py::arg("stride"), py::arg("pad"), py::arg("dil"), py::arg("groups") = 1,
        py::arg("peer") = py::none());
  m.def("conv_dgrad", &conv_dgrad, "tcgen05 implicit-GEMM data gradient (stride 1), optional fused addend",
        py::arg("dy"), py::arg("w"), py::arg("dx"), py::arg("stride"), py::arg("pad"), py::arg("dil"), py::arg("addend") = py::none(),
        py::arg("groups") = 1);
  m.def("conv_dgrad_s2", &conv_dgrad_s2, "stride-2 data gradient by parity classes (no zero insertion)", py::arg("dy"), py::arg("w"),
        py::arg("dx"), py::arg("pad"), py::arg("addend") = py::none(), py::arg("groups") = 1);
  m.def("attn_fwd", &attn_fwd, "fused relative-position MHSA forward (tcgen05)");
  m.def("attn_bwd", &attn_bwd, "fused relative-position MHSA backward: dQ, dK, dV and the relative-logit gradients");
  m.def("rel_grad_reduce", &rel_grad_reduce);
  m.def("conv3x3_halo", &conv3x3_halo, "3x3/s1/p1 64->64 convolution with one halo load per tile (forward or data gradient)",
        py::arg("x"), py::arg("w"), py::arg("y"), py::arg("stats") = py::none(), py::arg("dgrad") = false, py::arg("peer") = py::none());
  m.def("colsum_add", &colsum_add);
  m.def("stem_s2d", &stem_s2d, py::arg("x"), py::arg("out"), py::arg("mean") = std::vector<double>{}, py::arg("std") = std::vector<double>{});
  m.def("stem_s2d_pack_w", &stem_s2d_pack_w);
  m.def("stem_conv_fprop", &stem_conv_fprop, py::arg("xs"), py::arg("w2"), py::arg("y"), py::arg("stats"), py::arg("peer") = py::none());
  m.def("stem_s2d_unpack_dw", &stem_s2d_unpack_dw);
  m.def("strided_add_inplace", &strided_add_inplace);
  m.def("blockdiag_pack", &blockdiag_pack);
  m.def("blockdiag_unpack_add", &blockdiag_unpack_add);
  m.def("se_gate_fwd", &se_gate_fwd);
  m.def("se_gate_bwd", &se_gate_bwd);
  m.def("channel_add_bcast", &channel_add_bcast);
  m.def("conv_wgrad", &conv_wgrad, "tcgen05 split-K weight gradient (fp32 accumulate)", py::arg("dy"), py::arg("x"), py::arg("dw"),
        py::arg("stride"), py::arg("pad"), py::arg("dil"), py::arg("groups") = 1);
  py::class_<PeerState>(m, "PeerState")
      .def(py::init<>())
      .def_readwrite("world", &PeerState::world).def_readwrite("rank", &PeerState::rank)
      .def_readwrite("slot_base", &PeerState::slot_base).def_readwrite("signal_pads", &PeerState::signal_pads)
      .def_readwrite("sym_bufs", &PeerState::sym_bufs).def_readwrite("ticket", &PeerState::ticket)
      .def_readwrite("mc_stats", &PeerState::mc_stats).def_readwrite("reduced", &PeerState::reduced)
      .def_readwrite("ready", &PeerState::ready).def_readwrite("wait_ns", &PeerState::wait_ns)
      .def_readwrite("epoch_dev", &PeerState::epoch_dev)
      .def_readwrite("epoch", &PeerState::epoch);
  py::class_<CommState>(m, "CommState")
      .def(py::init<>())
      .def_readwrite("world", &CommState::world).def_readwrite("rank", &CommState::rank)
      .def_readwrite("slot_base", &CommState::slot_base).def_readwrite("signal_pads", &CommState::signal_pads)
      .def_readwrite("stage", &CommState::stage).def_readwrite("w16", &CommState::w16)
      .def_readwrite("mc_stage", &CommState::mc_stage).def_readwrite("mc_w16", &CommState::mc_w16)
      .def_readwrite("local_counter", &CommState::local_counter).def_readwrite("local_release", &CommState::local_release)
      .def_readwrite("epoch_dev", &CommState::epoch_dev)
      .def_readwrite("epoch", &CommState::epoch);
  m.def("dw_fprop", &dw_fprop, "depthwise conv forward (+BN statistics)");
  m.def("dw_dgrad", &dw_dgrad);
  m.def("dw_wgrad", &dw_wgrad);
  m.def("bn_apply", &bn_apply, py::arg("y"), py::arg("residual"), py::arg("out"), py::arg("stats"), py::arg("sym_offset"), py::arg("gamma"),
        py::arg("beta"), py::arg("running_mean"), py::arg("running_var"), py::arg("save_mean"), py::arg("save_invstd"), py::arg("count"),
        py::arg("eps"), py::arg("momentum"), py::arg("act"), py::arg("training"), py::arg("peer"), py::arg("relu_mask") = py::none(),
        py::arg("presignaled") = false);
  m.def("bn_stats", &bn_stats);
  m.def("bn_backward", &bn_backward, py::arg("y"), py::arg("dout"), py::arg("residual"), py::arg("dy"), py::arg("dresidual"), py::arg("sums"),
        py::arg("sym_offset"), py::arg("gamma"), py::arg("beta"), py::arg("save_mean"), py::arg("save_invstd"), py::arg("dgamma"), py::arg("dbeta"),
        py::arg("count"), py::arg("act"), py::arg("peer"), py::arg("relu_mask") = py::none(), py::arg("phase") = 3);
  m.def("bn_relu_pool_fwd", &bn_relu_pool_fwd, py::arg("y"), py::arg("out"), py::arg("arg"), py::arg("stats"), py::arg("sym_offset"),
        py::arg("gamma"), py::arg("beta"), py::arg("running_mean"), py::arg("running_var"), py::arg("save_mean"), py::arg("save_invstd"),
        py::arg("count"), py::arg("eps"), py::arg("momentum"), py::arg("training"), py::arg("peer") = py::none(), py::arg("presignaled") = false);
  m.def("bn_relu_pool_bwd", &bn_relu_pool_bwd, py::arg("y"), py::arg("dout"), py::arg("arg"), py::arg("dy"), py::arg("sums"), py::arg("sym_offset"),
        py::arg("gamma"), py::arg("save_mean"), py::arg("save_invstd"), py::arg("dgamma"), py::arg("dbeta"), py::arg("count"),
        py::arg("peer") = py::none(), py::arg("phase") = 3);
  m.def("maxpool_fwd", &maxpool_fwd);
  m.def("maxpool_bwd", &maxpool_bwd);
  m.def("gap_fwd", &gap_fwd);
  m.def("gap_bwd", &gap_bwd);
  m.def("avgpool2_fwd", &avgpool2_fwd);
  m.def("avgpool2_bwd", &avgpool2_bwd);
  m.def("channel_scale_fwd", &channel_scale_fwd);
  m.def("channel_scale_bwd", &channel_scale_bwd);
  m.def("ce_topk", &ce_topk);
  m.def("nchw_to_nhwc", &nchw_to_nhwc);
  m.def("stem_im2col", &stem_im2col, py::arg("x"), py::arg("patches"), py::arg("R"), py::arg("S"), py::arg("stride"), py::arg("pad"),
        py::arg("P"), py::arg("Q"), py::arg("mean") = std::vector<double>{}, py::arg("std") = std::vector<double>{});
  m.def("pad_rows", &pad_rows);
  m.def("unpad_add", &unpad_add);
  m.def("sgd_local", &sgd_local);
  m.def("cast_bf16", &cast_bf16);
  m.def("allreduce_sgd", &allreduce_sgd);
  m.def("rank_barrier", &rank_barrier);
}
