// unet_b200_legacy.hpp -- the reference's C++ launcher signatures, parameter / activation structs and arena helpers
// (dev/*.cuh) as inline wrappers over the C ABI of unet_b200.h.  A maintainer of unet.cu either replaces
// `#include "conv2d_k3.cuh"` (etc.) by this header, or puts include/legacy/ -- same-named one-line shims for every
// dev/*.cuh -- in front of the include path, and links libunet_b200.so instead of the dev/*.o objects; call sites
// stay unchanged (tests/test_legacy_compile.py compiles the reference's own dev/resblock.cu,
// dev/attention_block.cu and dev/unet_test.cu this way).
//
// Two modes:
//   default                  : layer launchers + structs + the composite blocks (resblock_*, attention_block_*) as
//                              inline wrappers over ub_resblock_* / ub_attention_block_* -- for callers of the blocks
//                              (train_unet.cu:3877-4083, dev/unet_test.cu).
//   -DUB_LEGACY_LAYERS_ONLY  : layer launchers + structs + PROTOTYPES of the composites only -- for translation units
//                              that define the composites themselves on top of the layer API (dev/resblock.cu,
//                              dev/attention_block.cu).
// Arguments that only existed to drive the reference's implementation (cublasHandle_t, block_size, timing slots
// t1..t6, split-K scratch) are accepted and ignored.  Like the reference (utils.cuh:24-41) these wrappers print the
// error and exit on failure, so existing callers that ignore return values keep their fail-fast behaviour.
#pragma once
#include <cstddef>
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
// (dev/conv2d_k3.cuh:40-60: parameter / activation pointer pairs and the arena helpers of the U-Net planner)
typedef struct {
    float* w;
    float* b;
} ConvK3Params;
inline void convk3_set_param_ptrs(ConvK3Params* params, float* params_memory, int C_in, int C_out) {
    params->w = params_memory;
    params->b = params->w + C_in * C_out * 9;
}
inline size_t convk3_count_params(int C_in, int C_out) { return C_in * C_out * 9 + C_out; }
typedef struct {
    float* inp;
    float* out;
} ConvK3Acts;
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
// (dev/linear.cuh:20-38)
typedef struct {
    float* w;  // OC, C
    float* b;  // OC
} LinearParams;
typedef struct {
    float* inp;  // N, C
    float* out;  // N, OC
} LinearActs;
inline void linear_set_param_ptrs(LinearParams* params, float* params_memory, int C, int OC) {
    params->w = params_memory;
    params->b = params->w + OC * C;
}
inline size_t linear_count_params(int C, int OC) { return OC * C + OC; }
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
// (dev/groupnorm.cuh:15-49)
typedef struct {
    float* w;
    float* b;
} GroupNormParams;
inline void gn_set_param_ptrs(GroupNormParams* params, float* params_memory, int C) {
    params->w = params_memory;
    params->b = params->w + C;
}
typedef struct {
    float* x;
    float* out;
    float* mean;
    float* rstd;
} GroupNormActs;
inline void gn_set_act_ptrs(GroupNormActs* acts, float* acts_memory, int B, int C, int H, int W, int n_groups) {
    acts->out = acts_memory;
    acts->mean = acts->out + B * C * H * W;
    acts->rstd = acts->mean + B * n_groups;
}
typedef struct {
    float* dx;
    float* dout;
} GroupNormBackActs;
// ---- dev/silu.cuh, dev/add.cuh
typedef struct {
    float* x;
    float* out;
} SiluActs;
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
typedef struct {
    float* x;
    float* out;
} UpsampleActs;
typedef struct {
    float* x;
    float* out;
} AvgpoolActs;
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
typedef struct {
    float* x1;
    float* x2;
    float* out;
} ConcatChannelActs;
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

// ---- dev/resblock.cuh: struct layouts kept field for field (callers fill them, train_unet.cu:3877-4083)
#define NUM_RES_PARAM_TENSORS 12
typedef struct {
    float *gn1_w, *gn1_b, *cv3_1_w, *cv3_1_b, *l_emb_w, *l_emb_b, *gn2_w, *gn2_b, *cv3_2_w, *cv3_2_b, *res_cv1_w,
        *res_cv1_b;
    size_t param_sizes[NUM_RES_PARAM_TENSORS];
    size_t n_params;
} ResBlockParameters;
#define NUM_RES_ACT_TENSORS 18
typedef struct {
    float *gn1, *gn1_mean, *gn1_rstd, *silu1, *ud_h, *ud_x, *cv3_1, *silu_emb, *l_emb, *broad_emb, *add1, *gn2,
        *gn2_mean, *gn2_rstd, *silu2, *cv3_2, *res_cv1, *add2;
    size_t act_sizes[NUM_RES_ACT_TENSORS];
    size_t n_acts;
    float* input;
    float* emb;
} ResBlockActivations;
#define NUM_RES_BACKWARD_TENSORS 7
typedef struct {
    float *buf_BCemb, *buf_BCHoWo, *buf1_BCHW, *buf2_BCHW, *dout, *dweight_buf, *dbias_buf, *dx, *demb;
    size_t back_sizes[NUM_RES_BACKWARD_TENSORS];
    size_t n_backs;
} ResBlockBackwardActivations;

#ifdef UB_LEGACY_LAYERS_ONLY
// prototypes only (dev/resblock.cuh:71-131): the including translation unit defines them on top of the layer API
void set_resblock_params_ptrs(int C, int C_out, ResBlockParameters* params, float* params_memory_gpu);
void set_resblock_acts_ptrs(int C, int C_out, ResBlockActivations* acts, float* acts_memory);
void resblock_forward(cublasHandle_t cublas_handle, int C, int C_emb, int C_out, int B, int H, int W, int block_size,
                      int up, int down, int gn_n_groups, ResBlockParameters* params, ResBlockActivations* acts);
void resblock_backward(cublasHandle_t cublas_handle, int C, int C_emb, int C_out, int B, int H, int W, int block_size,
                       int up, int down, int gn_n_groups, ResBlockParameters* params, ResBlockParameters* grads,
                       ResBlockActivations* acts, ResBlockBackwardActivations* back_acts);
void resblock_count_params(ResBlockParameters* params, int C, int C_emb, int C_out, int B, int H, int W, int up,
                           int down, int gn_n_groups);
void resblock_count_acts(ResBlockActivations* acts, int C, int C_emb, int C_out, int B, int H, int W, int up, int down,
                         int gn_n_groups);
#else
inline UbResBlockParams ub_legacy_params(const ResBlockParameters* p) {
    return UbResBlockParams{p->gn1_w, p->gn1_b, p->cv3_1_w, p->cv3_1_b, p->l_emb_w, p->l_emb_b,
                            p->gn2_w, p->gn2_b, p->cv3_2_w, p->cv3_2_b, p->res_cv1_w, p->res_cv1_b};
}
inline UbResBlockActs ub_legacy_acts(const ResBlockActivations* a) {
    return UbResBlockActs{a->gn1, a->gn1_mean, a->gn1_rstd, a->silu1, a->ud_h, a->ud_x, a->cv3_1, a->silu_emb, a->l_emb,
                          a->broad_emb, a->add1, a->gn2, a->gn2_mean, a->gn2_rstd, a->silu2, a->cv3_2, a->res_cv1,
                          a->add2, a->input, a->emb};
}
// tensor sizes / pointer carving of the reference's arena planner (dev/resblock.cu:204-330): same order, same sizes
inline void resblock_count_params(ResBlockParameters* p, int C, int C_emb, int C_out, int, int, int, int, int, int) {
    const size_t s[NUM_RES_PARAM_TENSORS] = {size_t(C), size_t(C), size_t(C_out) * C * 9, size_t(C_out),
                                             size_t(C_out) * C_emb, size_t(C_out), size_t(C_out), size_t(C_out),
                                             size_t(C_out) * C_out * 9, size_t(C_out),
                                             C == C_out ? 0 : size_t(C_out) * C, C == C_out ? 0 : size_t(C_out)};
    p->n_params = 0;
    for (int i = 0; i < NUM_RES_PARAM_TENSORS; ++i) p->param_sizes[i] = s[i], p->n_params += s[i];
}
inline void set_resblock_params_ptrs(int, int, ResBlockParameters* p, float* mem) {
    float** ptrs[NUM_RES_PARAM_TENSORS] = {&p->gn1_w, &p->gn1_b, &p->cv3_1_w, &p->cv3_1_b, &p->l_emb_w, &p->l_emb_b,
                                          &p->gn2_w, &p->gn2_b, &p->cv3_2_w, &p->cv3_2_b, &p->res_cv1_w, &p->res_cv1_b};
    for (int i = 0; i < NUM_RES_PARAM_TENSORS; ++i) *ptrs[i] = mem, mem += p->param_sizes[i];
}
inline void resblock_count_acts(ResBlockActivations* a, int C, int C_emb, int C_out, int B, int H, int W, int up,
                                int down, int gn_n_groups) {
    const int Ho = up ? 2 * H : (down ? H / 2 : H), Wo = up ? 2 * W : (down ? W / 2 : W);
    const size_t in = size_t(B) * C * H * W, mid = size_t(B) * C * Ho * Wo, out = size_t(B) * C_out * Ho * Wo;
    const size_t g = size_t(B) * gn_n_groups;
    const size_t s[NUM_RES_ACT_TENSORS] = {in, g, g, in, mid, mid, out, size_t(B) * C_emb, size_t(B) * C_out, out, out,
                                           out, g, g, out, out, C == C_out ? 0 : out, out};
    a->n_acts = 0;
    for (int i = 0; i < NUM_RES_ACT_TENSORS; ++i) a->act_sizes[i] = s[i], a->n_acts += s[i];
}
inline void set_resblock_acts_ptrs(int, int, ResBlockActivations* a, float* mem) {
    float** ptrs[NUM_RES_ACT_TENSORS] = {&a->gn1, &a->gn1_mean, &a->gn1_rstd, &a->silu1, &a->ud_h, &a->ud_x,
                                        &a->cv3_1, &a->silu_emb, &a->l_emb, &a->broad_emb, &a->add1, &a->gn2,
                                        &a->gn2_mean, &a->gn2_rstd, &a->silu2, &a->cv3_2, &a->res_cv1, &a->add2};
    for (int i = 0; i < NUM_RES_ACT_TENSORS; ++i) *ptrs[i] = mem, mem += a->act_sizes[i];
}
inline void resblock_forward(cublasHandle_t, int C, int C_emb, int C_out, int B, int H, int W, int /*block_size*/, int up,
                             int down, int gn_n_groups, ResBlockParameters* params, ResBlockActivations* acts) {
    const UbResBlockParams p = ub_legacy_params(params);
    const UbResBlockActs a = ub_legacy_acts(acts);
    UB_LEGACY_CHECK(ub_resblock_forward(C, C_emb, C_out, B, H, W, up, down, gn_n_groups, &p, &a));
}
inline void resblock_backward(cublasHandle_t, int C, int C_emb, int C_out, int B, int H, int W, int /*block_size*/, int up,
                              int down, int gn_n_groups, ResBlockParameters* params, ResBlockParameters* grads,
                              ResBlockActivations* acts, ResBlockBackwardActivations* back_acts) {
    const UbResBlockParams p = ub_legacy_params(params), g = ub_legacy_params(grads);
    const UbResBlockActs a = ub_legacy_acts(acts);
    const UbResBlockBack k{back_acts->buf_BCemb, back_acts->buf_BCHoWo, back_acts->buf1_BCHW, back_acts->buf2_BCHW,
                           back_acts->dout, back_acts->dx, back_acts->demb};
    UB_LEGACY_CHECK(ub_resblock_backward(C, C_emb, C_out, B, H, W, up, down, gn_n_groups, &p, &g, &a, &k));
}

// backward scratch sizes of the reference planner (dev/resblock.cu:206-233; dweight_buf / dbias_buf are the split-K
// scratch of conv2d_k3_backward2, which this library ignores -- sized as the reference does so arenas stay identical)
inline void resblock_count_backs(ResBlockBackwardActivations* b, int C, int C_emb, int C_out, int B, int H, int W, int up,
                                 int down, int /*gn_n_groups*/) {
    const int Ho = up ? 2 * H : (down ? H / 2 : H), Wo = up ? 2 * W : (down ? W / 2 : W);
    const size_t s[NUM_RES_BACKWARD_TENSORS] = {size_t(B) * C_emb, size_t(B) * C * Ho * Wo, size_t(B) * C * H * W,
                                                size_t(B) * C * H * W, size_t(B) * C_out * Ho * Wo,
                                                size_t(C_out) * C * 9 * B * 32, size_t(C_out) * B * 32};
    b->n_backs = 0;
    for (int i = 0; i < NUM_RES_BACKWARD_TENSORS; ++i) b->back_sizes[i] = s[i], b->n_backs += s[i];
}
#endif  // UB_LEGACY_LAYERS_ONLY

// ---- dev/attention_block.cuh
#define NUM_ATT_PARAM_TENSORS 6
typedef struct {
    float *gn_w, *gn_b, *qkv_w, *qkv_b, *proj_w, *proj_b;
    size_t param_sizes[NUM_ATT_PARAM_TENSORS];
    size_t n_params;
} AttentionParams;
#define NUM_ATT_ACT_TENSORS 12
typedef struct {
    float *gn, *gn_mean, *gn_rstd, *perm1, *qkv1, *qkv2, *preatt, *att, *att_out, *proj, *perm2, *add;
    size_t act_sizes[NUM_ATT_ACT_TENSORS];
    size_t n_acts;
    float* input;
} AttentionActs;
#define NUM_ATT_BACKWARD_ACTS_TENSORS 7
typedef struct {
    float *buf1_BCHW, *buf2_BCHW, *buf_B3CHW, *dqkvr, *dpreatt, *datt, *dout, *dinp;
    size_t back_sizes[NUM_ATT_BACKWARD_ACTS_TENSORS];
    size_t n_backs;
} AttentionBackwardActs;

// (dev/attention_block.cuh:47-73: fixture / planner structs the reference's tests use)
#define NUM_ATT_DEBUG_STATES 9
typedef struct {
    float *inp, *gn, *perm1, *qkv, *att, *proj, *out, *dout, *dinp;
} AttentionDebugStates;
typedef struct {
    int C;
    int H;
    int W;
    int HS;
    int B;
    int gn_n_groups;
    size_t param_sizes[NUM_ATT_PARAM_TENSORS];
    size_t act_sizes[NUM_ATT_ACT_TENSORS];
    size_t back_sizes[NUM_ATT_BACKWARD_ACTS_TENSORS];
    size_t n_params;
    size_t n_acts;
    size_t n_backs;
} AttentionConfig;

#ifdef UB_LEGACY_LAYERS_ONLY
// prototypes only (dev/attention_block.cuh:76-112)
void attention_block_count_params(AttentionParams* params, int C);
void attention_block_count_acts(AttentionActs* acts, int B, int C, int H, int W, int gn_n_groups, int HS);
void attention_block_forward(cublasHandle_t cublas_handle, int B, int C, int H, int W, int HS, int gn_n_groups,
                             int block_size, AttentionParams* params, AttentionActs* acts);
void attention_block_backward(cublasHandle_t cublas_handle, int B, int C, int H, int W, int HS, int gn_n_groups,
                              int block_size, AttentionParams* params, AttentionActs* acts,
                              AttentionBackwardActs* back_acts, AttentionParams* grads);
void set_attention_params_pointers(AttentionParams* params, float* params_memory);
void set_attention_acts_pointers(AttentionActs* acts, float* acts_memory);
#else
inline void attention_block_count_params(AttentionParams* p, int C) {
    const size_t s[NUM_ATT_PARAM_TENSORS] = {size_t(C), size_t(C), size_t(3) * C * C, size_t(3) * C, size_t(C) * C,
                                             size_t(C)};
    p->n_params = 0;
    for (int i = 0; i < NUM_ATT_PARAM_TENSORS; ++i) p->param_sizes[i] = s[i], p->n_params += s[i];
}
inline void attention_block_count_acts(AttentionActs* a, int B, int C, int H, int W, int gn_n_groups, int HS) {
    const size_t T = size_t(H) * W, x = size_t(B) * C * T, g = size_t(B) * gn_n_groups, tt = size_t(B) * (C / HS) * T * T;
    const size_t s[NUM_ATT_ACT_TENSORS] = {x, g, g, x, 3 * x, 3 * x, tt, tt, x, x, x, x};
    a->n_acts = 0;
    for (int i = 0; i < NUM_ATT_ACT_TENSORS; ++i) a->act_sizes[i] = s[i], a->n_acts += s[i];
}
inline void set_attention_params_pointers(AttentionParams* p, float* mem) {
    float** ptrs[NUM_ATT_PARAM_TENSORS] = {&p->gn_w, &p->gn_b, &p->qkv_w, &p->qkv_b, &p->proj_w, &p->proj_b};
    for (int i = 0; i < NUM_ATT_PARAM_TENSORS; ++i) *ptrs[i] = mem, mem += p->param_sizes[i];
}
inline void set_attention_acts_pointers(AttentionActs* a, float* mem) {
    float** ptrs[NUM_ATT_ACT_TENSORS] = {&a->gn, &a->gn_mean, &a->gn_rstd, &a->perm1, &a->qkv1, &a->qkv2,
                                        &a->preatt, &a->att, &a->att_out, &a->proj, &a->perm2, &a->add};
    for (int i = 0; i < NUM_ATT_ACT_TENSORS; ++i) *ptrs[i] = mem, mem += a->act_sizes[i];
}
inline void attention_block_forward(cublasHandle_t, int B, int C, int H, int W, int HS, int gn_n_groups,
                                    int /*block_size*/, AttentionParams* params, AttentionActs* acts) {
    const UbAttentionParams p{params->gn_w, params->gn_b, params->qkv_w, params->qkv_b, params->proj_w, params->proj_b};
    const UbAttentionActs a{acts->gn, acts->gn_mean, acts->gn_rstd, acts->perm1, acts->qkv1, acts->qkv2, acts->preatt,
                            acts->att, acts->att_out, acts->proj, acts->perm2, acts->add, acts->input};
    UB_LEGACY_CHECK(ub_attention_block_forward(B, C, H, W, HS, gn_n_groups, &p, &a));
}
inline void attention_block_backward(cublasHandle_t, int B, int C, int H, int W, int HS, int gn_n_groups,
                                     int /*block_size*/, AttentionParams* params, AttentionActs* acts,
                                     AttentionBackwardActs* back_acts, AttentionParams* grads) {
    const UbAttentionParams p{params->gn_w, params->gn_b, params->qkv_w, params->qkv_b, params->proj_w, params->proj_b};
    const UbAttentionParams g{grads->gn_w, grads->gn_b, grads->qkv_w, grads->qkv_b, grads->proj_w, grads->proj_b};
    const UbAttentionActs a{acts->gn, acts->gn_mean, acts->gn_rstd, acts->perm1, acts->qkv1, acts->qkv2, acts->preatt,
                            acts->att, acts->att_out, acts->proj, acts->perm2, acts->add, acts->input};
    const UbAttentionBack k{back_acts->buf1_BCHW, back_acts->buf2_BCHW, back_acts->buf_B3CHW, back_acts->dqkvr,
                            back_acts->dpreatt, back_acts->datt, back_acts->dout, back_acts->dinp};
    UB_LEGACY_CHECK(ub_attention_block_backward(B, C, H, W, HS, gn_n_groups, &p, &a, &k, &g));
}
// backward scratch of the reference planner (dev/attention_block.cu:153-197)
inline void attention_block_count_backs(AttentionBackwardActs* b, int B, int C, int H, int W, int HS, int /*gn_n_groups*/) {
    const size_t T = size_t(H) * W, x = size_t(B) * C * T, tt = size_t(B) * (C / HS) * T * T;
    const size_t s[NUM_ATT_BACKWARD_ACTS_TENSORS] = {x, x, 3 * x, 3 * x, tt, tt, x};
    b->n_backs = 0;
    for (int i = 0; i < NUM_ATT_BACKWARD_ACTS_TENSORS; ++i) b->back_sizes[i] = s[i], b->n_backs += s[i];
}
inline void set_attention_back_acts_pointers(AttentionBackwardActs* b, float* mem, size_t* back_act_sizes) {
    float** ptrs[NUM_ATT_BACKWARD_ACTS_TENSORS] = {&b->buf1_BCHW, &b->buf2_BCHW, &b->buf_B3CHW, &b->dqkvr,
                                                   &b->dpreatt,   &b->datt,      &b->dout};
    for (int i = 0; i < NUM_ATT_BACKWARD_ACTS_TENSORS; ++i) *ptrs[i] = mem, mem += back_act_sizes[i];
}
#endif  // UB_LEGACY_LAYERS_ONLY
