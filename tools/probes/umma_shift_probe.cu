// Probe: can a tcgen05.mma A operand (K-major, 128B swizzle) start at a shared-memory row that is NOT a multiple of 8
// (i.e. not 1024-byte aligned)?  A halo-reuse 3x3 convolution wants exactly that: one activation tile in shared memory,
// nine MMAs whose A descriptors start r*(W+2)+s rows further down.  For every shift r in [0, 16) and both settings of the
// descriptor's base_offset field (0, and (start >> 7) & 7) the kernel multiplies rows [r, r+128) of a 256-row tile by an
// identity B and the host checks D[m][n] == A[r+m][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I distribuuuu_b200/csrc tools/probes/umma_shift_probe.cu -o gpurun_out/umma_shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace b200;

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                       float* out, int shift, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;               // 256 rows x 128 B
  uint8_t* s_b = smem + 32768;       // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_ptr), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), 32768 + 8192);
    tma_load_3d(smem_u32(s_a), &map_a, smem_u32(&bars[0]), 0, 0, 0);
    tma_load_3d(smem_u32(s_a + 16384), &map_a, smem_u32(&bars[0]), 0, 128, 0);
    tma_load_3d(smem_u32(s_b), &map_b, smem_u32(&bars[0]), 0, 0, 0);
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(s_a) + shift * 128;
    uint64_t a_desc = make_smem_desc_hi_sw128(16, 1024) | (uint64_t)((a_addr >> 4) & 0x3fff);
    if (use_base_offset) a_desc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t b_desc = make_smem_desc_hi_sw128(16, 1024) | (uint64_t)((smem_u32(s_b) >> 4) & 0x3fff);
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(smem_u32(&bars[1]));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, int rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<__nv_bfloat16> ha(256 * 64), hb(64 * 64);
  for (int i = 0; i < 256; ++i) for (int k = 0; k < 64; ++k) ha[i * 64 + k] = __float2bfloat16((float)((i * 3 + k * 7) % 61 - 30));
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma = make_map(enc, da, 256, 128), mb = make_map(enc, db, 64, 64);
  const int smem = 32768 + 8192 + 1024 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(128 * 64);
  for (int bo = 0; bo < 2; ++bo) {
    for (int shift = 0; shift < 16; ++shift) {
      cudaMemset(dout, 0xff, 128 * 64 * 4);
      probe_kernel<<<1, 128, smem>>>(ma, mb, dout, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        const float want = __bfloat162float(ha[(shift + m) * 64 + n]);
        if (ho[m * 64 + n] != want) { if (first < 0) first = m * 64 + n; ++bad; }
      }
      printf("shift %2d base_offset_field %s: %s (%d mismatches%s)\n", shift, bo ? "(addr>>7)&7" : "0", bad ? "WRONG" : "exact", bad,
             bad ? "" : "");
      if (bad && first >= 0) printf("    first mismatch at m=%d n=%d: got %.1f want %.1f\n", first / 64, first % 64, ho[first], __bfloat162float(ha[(shift + first / 64) * 64 + first % 64]));
    }
  }
  return 0;
}
