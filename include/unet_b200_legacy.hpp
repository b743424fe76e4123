// unet_b200_legacy.hpp -- the reference's C++ launcher signatures (dev/*.cuh) as inline wrappers over the C ABI of
// unet_b200.h.  A maintainer of unet.cu replaces `#include "conv2d_k3.cuh"` (etc.) by this header and links
// libunet_b200.so instead of the dev/*.o objects; call sites stay unchanged.
// Arguments that only existed to drive the reference's implementation (cublasHandle_t, block_size, timing slots
// t1..t6, split-K scratch) are accepted and ignored.  Like the reference (utils.cuh:24-41) these wrappers print the
// error and exit on failure, so existing callers that ignore return values keep their fail-fast behaviour.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "unet_b200.h"

#ifndef UB_LEGACY_NO_CUBLAS
#include <cublas_v2.h>
#else
typedef void* cublasHandle_t;
#endif

#define UB_LEGACY_CHECK(call)                                                                   \
    do {                                                                                        \
        int rc_ = (call);                                                                       \
        if (rc_ != UB_OK) {                                                                     \
            std::fprintf(stderr, "[unet_b200] %s failed (%d): %s\n", #call, rc_, ub_last_error()); \
            std::exit(EXIT_FAILURE);                                                            \
        }                                                                                       \
    } while (0)

// ---- dev/conv2d_k3.cuh
inline void conv2d_k3_forward3(float* x, float* weight, float* bias, float* out, const int B, const int C_in,
                               const int C_out, const int H, const int W) {
    UB_LEGACY_CHECK(ub_conv2d_k3_forward3(x, weight, bias, out, B, C_in, C_out, H, W));
}
inline void conv2d_k3_forward2(cublasHandle_t, const float* x, const float* weight, const float* bias, float* out,
                               int B, int C_in, int C_out, int H, int W, int /*block_size*/, float* = nullptr,
                               float* = nullptr, float* = nullptr) {
    UB_LEGACY_CHECK(ub_conv2d_k3_forward2(x, weight, bias, out, B, C_in, C_out, H, W));
}
inline void conv2d_k3_backward2(float* dout, float* x, float* weight, float* dweight_buf, float* dbias_buf, float* dx,
                                float* dweight, float* dbias, const int B, const int C_in, const int C_out,
                                const int H, const int W) {
    UB_LEGACY_CHECK(ub_conv2d_k3_backward2(dout, x, weight, dweight_buf, dbias_buf, dx, dweight, dbias, B, C_in,
                                           C_out, H, W));
}
inline void conv2d_k3_backward1(cublasHandle_t, const float* dout, const float* x, const float* weight, float* dx,
                                float* dweight, float* dbias, int B, int C_in, int C_out, int H, int W,
                                int /*block_size*/, float* = nullptr, float* = nullptr, float* = nullptr,
                                float* = nullptr, float* = nullptr, float* = nullptr) {
    UB_LEGACY_CHECK(ub_conv2d_k3_backward1(dout, x, weight, dx, dweight, dbias, B, C_in, C_out, H, W));
}
// ---- dev/conv2d_k1.cuh
inline void conv2d_k1_forward1(cublasHandle_t, float* out, const float* x, const float* weight, const float* bias,
                               int B, int C_in, int H, int W, int C_out, const int /*block_size*/, float* = nullptr,
                               float* = nullptr, float* = nullptr) {
    UB_LEGACY_CHECK(ub_conv2d_k1_forward1(out, x, weight, bias, B, C_in, H, W, C_out));
}
inline void conv2d_k1_forward2(const float* x, const float* weight, const float* bias, float* out, int B, int C_in,
                               int H, int W, int C_out) {
    UB_LEGACY_CHECK(ub_conv2d_k1_forward2(x, weight, bias, out, B, C_in, H, W, C_out));
}
inline void conv2d_k1_backward1(cublasHandle_t, const float* dout, const float* x, const float* weight, float* dx,
                                float* dweight, float* dbias, int B, int C_in, int C_out, int H, int W,
                                int /*block_size*/) {
    UB_LEGACY_CHECK(ub_conv2d_k1_backward1(dout, x, weight, dx, dweight, dbias, B, C_in, C_out, H, W));
}
// ---- dev/linear.cuh
inline void matmul_forward2(cublasHandle_t, float* out, const float* inp, const float* weight, const float* bias,
                            int N, int C, int OC, const int /*block_size*/) {
    UB_LEGACY_CHECK(ub_matmul_forward2(out, inp, weight, bias, N, C, OC));
}
inline void matmul_backward1(cublasHandle_t, float* dinp, float* dweight, float* dbias, float* dout, float* inp,
                             float* weight, int N, int C, int OC) {
    UB_LEGACY_CHECK(ub_matmul_backward1(dinp, dweight, dbias, dout, inp, weight, N, C, OC));
}
// ---- dev/groupnorm.cuh
inline void groupnorm_forward(const float* x, const float* weight, const float* bias, float* out, float* mean,
                              float* rstd, int B, int C, int H, int W, int n_groups) {
    UB_LEGACY_CHECK(ub_groupnorm_forward(x, weight, bias, out, mean, rstd, B, C, H, W, n_groups));
}
inline void groupnorm_backward(const float* dout, const float* x, const float* mean, const float* rstd,
                               const float* weight, float* dx, float* dweight, float* dbias, int B, int C, int H,
                               int W, int n_groups) {
    UB_LEGACY_CHECK(ub_groupnorm_backward(dout, x, mean, rstd, weight, dx, dweight, dbias, B, C, H, W, n_groups));
}
// ---- dev/silu.cuh, dev/add.cuh
inline void silu_forward(const float* x, float* out, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_silu_forward(x, out, N));
}
inline void silu_backward(const float* dout, const float* x, float* dx, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_silu_backward(dout, x, dx, N));
}
inline void add_forward(const float* a, const float* b, float* out, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_add_forward(a, b, out, N));
}
inline void add_inplace_forward(const float* a, float* b, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_add_inplace_forward(a, b, N));
}
// ---- dev/upsample.cuh, dev/avgpool.cuh
inline void upsample_forward1(float* out, const float* x, int B, int C, int H, int W, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_upsample_forward1(out, x, B, C, H, W));
}
inline void upsample_backward1(float* dx, const float* dout, int B, int C, int H, int W, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_upsample_backward1(dx, dout, B, C, H, W));
}
inline void avgpool_2d_forward1(float* out, const float* x, int B, int C, int H, int W, const int /*block_size*/) {
    UB_LEGACY_CHECK(ub_avgpool_2d_forward1(out, x, B, C, H, W));
}
inline void avgpool_2d_backward1(const float* dout, float* dx, int B, int C, int H, int W,
                                 const int /*block_size*/) {
    UB_LEGACY_CHECK(ub_avgpool_2d_backward1(dout, dx, B, C, H, W));
}
// ---- dev/concat_channel.cuh, dev/broadcast.cuh
inline void concat_channel_forward(const float* x1, const float* x2, float* out, int B, int C1, int C2, int H, int W,
                                   int /*block_size*/) {
    UB_LEGACY_CHECK(ub_concat_channel_forward(x1, x2, out, B, C1, C2, H, W));
}
inline void concat_channel_backward(const float* dout, float* dx1, float* dx2, int B, int C1, int C2, int H, int W,
                                    int /*block_size*/) {
    UB_LEGACY_CHECK(ub_concat_channel_backward(dout, dx1, dx2, B, C1, C2, H, W));
}
inline void broadcast_last_dims_forward(const float* x, float* out, int N, int H, int W, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_broadcast_last_dims_forward(x, out, N, H, W));
}
inline void broadcast_last_dims_backward(const float* dout, float* dx, int N, int H, int W, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_broadcast_last_dims_backward(dout, dx, N, H, W));
}
// ---- dev/mse.cuh
inline void mse_forward(const float* inp, const float* y, float* loss, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_mse_forward(inp, y, loss, N));
}
inline void mse_backward(const float* inp, const float* y, float* dinp, int N, int /*block_size*/) {
    UB_LEGACY_CHECK(ub_mse_backward(inp, y, dinp, N));
}
// ---- dev/timestep_embedding.cuh (the freqs table of the reference struct is not needed: recomputed on the fly)
typedef struct {
    float* freqs;
    int half_dim;
    int B;
    int max_period;
} TimestepEmbedding;
inline void init_timestep_embedding(TimestepEmbedding* emb, int dim, int B, int max_period = 10000) {
    emb->freqs = nullptr, emb->half_dim = dim / 2, emb->B = B, emb->max_period = max_period;
}
inline void free_timestep_embedding(TimestepEmbedding*) {}
inline void get_timestep_embeddings(TimestepEmbedding* emb, const float* timesteps, float* out) {
    UB_LEGACY_CHECK(ub_get_timestep_embeddings(timesteps, out, emb->B, 2 * emb->half_dim, emb->max_period));
}
// ---- dev/attention.cuh
inline void attention_forward1(cublasHandle_t, float* out, float* qkvr, float* preatt, float* att, float* inp, int B,
                               int T, int C, int NH, const int /*block_size*/) {
    UB_LEGACY_CHECK(ub_attention_forward1(out, qkvr, preatt, att, inp, B, T, C, NH));
}
inline void attention_backward(cublasHandle_t, float* dinp, float* dqkvr, float* dpreatt, float* datt, float* scratch,
                               const float* dout, const float* qkvr, const float* att, int B, int T, int C, int NH) {
    UB_LEGACY_CHECK(ub_attention_backward(dinp, dqkvr, dpreatt, datt, scratch, dout, qkvr, att, B, T, C, NH));
}
