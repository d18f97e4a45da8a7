// Launchers of the small kernels in extras.cu (bias gradient, stride-2 dgrad assembly, thin-group weight packing,
// squeeze-excite MLP).
#pragma once
#include <cuda_runtime.h>

extern "C" {
int b200_colsum_add(const void* d, long long rows, int C, long long ld, float* out, cudaStream_t s);
int b200_stem_s2d(const void* x, int x_is_u8, void* out, int N, int H, int W, int Hs, int Ws, const float* scale3, const float* bias3, cudaStream_t s);
int b200_stem_s2d_pack_w(const void* w, void* wp, int K, cudaStream_t s);
int b200_stem_s2d_unpack_dw(const float* dwp, float* dw, int K, cudaStream_t s);
int b200_parity_interleave(const void* const* src4, const void* addend, void* dx, int N, int H, int W, int C, cudaStream_t s);
int b200_strided_add_inplace(void* dx, const void* compact, int N, int H, int W, int C, int P, int Q, int stride, cudaStream_t s);
int b200_blockdiag_pack(const void* thin, void* dense, int K, int taps, int cg, cudaStream_t s);
int b200_blockdiag_unpack_add(const float* dense, float* thin, int K, int taps, int cg, cudaStream_t s);
int b200_se_gate_fwd(const void* sp, const void* w1, const float* b1, const void* w2, const float* b2, float* pre1, void* gate, int N, int C,
                     int r, int act, cudaStream_t s);
int b200_se_gate_bwd(const float* dgate, const void* gate, const void* sp, const float* pre1, const void* w1, const void* w2, float* dw1,
                     float* db1, float* dw2, float* db2, float* ds, float* scratch, int N, int C, int r, int act, cudaStream_t s);
int b200_channel_add_bcast(void* dx, const float* ds, int N, int HW, int C, float scale, cudaStream_t s);
}
