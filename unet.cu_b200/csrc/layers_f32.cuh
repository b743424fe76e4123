// fp32 NCHW kernels behind the drop-in layer API (reference layouts of dev/*.cuh). The memory-bound operators are
// implemented directly in the reference layout (exact fp32); the contractions convert to the NHWC bf16 operand
// layout of the tcgen05 kernels with the two layout kernels below.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace ub {
namespace f32 {

// layout conversion: x (B, C, HW) fp32 -> y (B, HW, C) bf16 (tiled transpose through smem)
void nchw_to_nhwc_bf16(const float* x, int B, int C, int HW, __nv_bfloat16* y, cudaStream_t st);
void cast_bf16(const float* x, size_t n, __nv_bfloat16* y, cudaStream_t st);
// fp32 layout permutes of the attention block (dev/attention_block.cu:7-62): (B, C, HW) <-> (B, HW, C), tiled
void permute_bchw_to_bhwc(const float* x, float* y, int B, int C, int HW, cudaStream_t st);
void permute_bhwc_to_bchw(const float* x, float* y, int B, int C, int HW, cudaStream_t st);

// GroupNorm (train_unet.cu:1768-1991)
void groupnorm_fwd(const float* x, const float* w, const float* b, float* out, float* mean, float* rstd, int B, int C,
                   int HW, int G, cudaStream_t st);
void groupnorm_bwd(const float* dout, const float* x, const float* mean, const float* rstd, const float* w, float* dx,
                   float* dw, float* db, int B, int C, int HW, int G, cudaStream_t st);

// elementwise (train_unet.cu:187-351)
void silu_fwd(const float* x, float* out, size_t n, cudaStream_t st);
void silu_bwd(const float* dout, const float* x, float* dx, size_t n, cudaStream_t st);
void add(const float* a, const float* b, float* out, size_t n, cudaStream_t st);

// resampling (train_unet.cu:360-533): x (BC, H, W)
void upsample_fwd(float* out, const float* x, size_t BC, int H, int W, cudaStream_t st);
void upsample_bwd(float* dx, const float* dout, size_t BC, int H, int W, cudaStream_t st);
void avgpool_fwd(float* out, const float* x, size_t BC, int H, int W, cudaStream_t st);
void avgpool_bwd(const float* dout, float* dx, size_t BC, int H, int W, cudaStream_t st);

// concat along channels (train_unet.cu:545-627)
void concat_fwd(const float* x1, const float* x2, float* out, int B, int C1, int C2, int HW, cudaStream_t st);
void concat_bwd(const float* dout, float* dx1, float* dx2, int B, int C1, int C2, int HW, cudaStream_t st);

// broadcast over the last two dims (train_unet.cu:187-260)
void broadcast_fwd(const float* x, float* out, size_t N, int HW, cudaStream_t st);
void broadcast_bwd(const float* dout, float* dx, size_t N, int HW, cudaStream_t st);

// loss (train_unet.cu:2981-3030)
void mse_fwd(const float* inp, const float* y, float* loss, size_t n, cudaStream_t st);
void mse_bwd(const float* inp, const float* y, float* dinp, size_t n, cudaStream_t st);

// per-channel sum of an NCHW tensor: out[c] = sum_{b,hw} x[b][c][hw]
void nchw_chansum(const float* x, int B, int C, size_t HW, float* out, cudaStream_t st);
// column sums of a row-major (N, C) matrix
void rows_colsum(const float* x, int N, int C, float* out, cudaStream_t st);

// exact fp32 direct convolution (any channel count; used when the tcgen05 path's divisibility rules do not hold,
// e.g. the 3-channel first / last layer).  KS = 1 or 3, pad KS/2, stride 1.
void conv_direct_fwd(const float* x, const float* w, const float* b, float* out, int B, int Cin, int Cout, int H, int W,
                     int KS, cudaStream_t st);
void conv_direct_dgrad(const float* dout, const float* w, float* dx, int B, int Cin, int Cout, int H, int W, int KS,
                       cudaStream_t st);
void conv_direct_wgrad(const float* dout, const float* x, float* dw, float* db, int B, int Cin, int Cout, int H, int W,
                       int KS, cudaStream_t st);

// multi-head attention in the reference's layouts (train_unet.cu:2389-2760), fp32
void attention_fwd(float* out, float* qkvr, float* preatt, float* att, const float* inp, int B, int T, int C, int NH,
                   cudaStream_t st);
void attention_bwd(float* dinp, float* dqkvr, float* dpreatt, float* datt, const float* dout, const float* qkvr,
                   const float* att, int B, int T, int C, int NH, cudaStream_t st);

// small linear on fp32 rows (fallback of matmul_forward2 / matmul_backward1 for tiny or oddly shaped problems)
void linear_fwd(float* out, const float* inp, const float* w, const float* b, int N, int C, int OC, cudaStream_t st);
void linear_bwd(float* dinp, float* dw, float* db, const float* dout, const float* inp, const float* w, int N, int C,
                int OC, cudaStream_t st);

}  // namespace f32
}  // namespace ub
