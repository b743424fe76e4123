/* unet_b200.h -- C ABI of the B200-native U-Net diffusion training path.
 *
 * Two groups of entry points:
 *
 *  (1) Layer operators: drop-in replacements of the launchers declared in the reference's dev/\*.cuh
 *      (same argument order and meaning, device pointers to fp32 NCHW / (N,C) buffers, int shapes).  The C symbols
 *      carry a `ub_` prefix so that they can be linked next to the reference objects; include/unet_b200_legacy.hpp
 *      re-exposes the exact reference C++ signatures (including the ignored cublasHandle_t / block_size / t1..t6
 *      arguments) as inline wrappers.  Every function returns 0 on success or a non-zero error code instead of
 *      calling exit() (reference convention: utils.cuh:24-41 cudaCheck -> exit(EXIT_FAILURE)).
 *      All layer calls run on the stream set by ub_set_stream() (default: the legacy default stream, like the
 *      reference which never names a stream).
 *
 *  (2) Trainer: the model graph + optimizer + checkpoint path of train_unet.cu (unet_forward / unet_backward /
 *      unet_update / save_unet_states / load_unet_and_diffusion_states, train_unet.cu:4335-4911) behind an opaque
 *      handle, with batch data parallelism over NCCL.
 *
 * Each declaration cites the reference interface it replaces (paths under /root/reference).
 */
#ifndef UNET_B200_H
#define UNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB_OK 0
#define UB_ERR_SHAPE (-1)   /* unsupported shape / violated precondition (reference: device/host assert) */
#define UB_ERR_CUDA (-2)    /* a CUDA runtime call failed; see ub_last_error() */
#define UB_ERR_IO (-3)      /* file error */
#define UB_ERR_STATE (-4)   /* call sequence error (e.g. update before backward) */
#define UB_ERR_NCCL (-5)

const char* ub_last_error(void);
const char* ub_version(void);
/* stream used by the layer operators; `stream` is a cudaStream_t */
int ub_set_stream(void* stream);
/* Arithmetic of the layer operators' contractions (3x3 / 1x1 convolution, linear):
 *   UB_PRECISION_BF16 (default): bf16 operands on the tcgen05 tensor cores, fp32 accumulation -- the fast path;
 *   UB_PRECISION_FP32          : exact fp32 CUDA-core kernels of this library (slow; a validation mode: the
 *                                reference's own self-tests, dev/resblock.cu:542-630, compare at 1e-5 / 1e-4 absolute).
 * The environment variable UB_LAYER_PRECISION=fp32|bf16 sets the initial value.  Returns the previous mode. */
#define UB_PRECISION_BF16 0
#define UB_PRECISION_FP32 1
int ub_set_layer_precision(int mode);
/* number of kernels launched by this library since process start (bench.py's gpu_launches) */
unsigned long long ub_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * (1) Layer operators
 * ---------------------------------------------------------------------------------------------------------- */

/* dev/conv2d_k3.cuh:17-21  conv2d_k3_forward3(x, weight, bias, out, B, C_in, C_out, H, W)
 * 3x3, pad 1, stride 1.  x (B,C_in,H,W), weight (C_out,C_in,3,3), bias (C_out), out (B,C_out,H,W).
 * BF16 tensor-core path (fp32 accumulate) when C_in % 8 == 0 and C_out % 16 == 0, exact fp32 SIMT otherwise
 * (C_in = 3 or C_out = 3). */
int ub_conv2d_k3_forward3(const float* x, const float* weight, const float* bias, float* out, int B, int C_in,
                          int C_out, int H, int W);
/* dev/conv2d_k3.cuh:8-15  conv2d_k3_forward2 -- superseded generation, same maths; alias of forward3 */
int ub_conv2d_k3_forward2(const float* x, const float* weight, const float* bias, float* out, int B, int C_in,
                          int C_out, int H, int W);
/* dev/conv2d_k3.cuh:33-38  conv2d_k3_backward2(dout, x, weight, dweight_buf, dbias_buf, dx, dweight, dbias, ...)
 * dx, dweight, dbias are overwritten.  dweight_buf / dbias_buf (the reference's split-K scratch) are accepted and
 * ignored: the library keeps its own workspace.  dx may be NULL (first layer). */
int ub_conv2d_k3_backward2(const float* dout, const float* x, const float* weight, float* dweight_buf,
                           float* dbias_buf, float* dx, float* dweight, float* dbias, int B, int C_in, int C_out,
                           int H, int W);
/* dev/conv2d_k3.cuh:23-31  conv2d_k3_backward1 -- superseded generation; alias of backward2 */
int ub_conv2d_k3_backward1(const float* dout, const float* x, const float* weight, float* dx, float* dweight,
                           float* dbias, int B, int C_in, int C_out, int H, int W);

/* dev/conv2d_k1.cuh:15-19  conv2d_k1_forward2(x, weight, bias, out, B, C_in, H, W, C_out) */
int ub_conv2d_k1_forward2(const float* x, const float* weight, const float* bias, float* out, int B, int C_in, int H,
                          int W, int C_out);
/* dev/conv2d_k1.cuh:6-13 conv2d_k1_forward1(cublas, out, x, weight, bias, B, C_in, H, W, C_out, ...) alias */
int ub_conv2d_k1_forward1(float* out, const float* x, const float* weight, const float* bias, int B, int C_in, int H,
                          int W, int C_out);
/* dev/conv2d_k1.cuh:21-27  conv2d_k1_backward1(cublas, dout, x, weight, dx, dweight, dbias, B, C_in, C_out, H, W) */
int ub_conv2d_k1_backward1(const float* dout, const float* x, const float* weight, float* dx, float* dweight,
                           float* dbias, int B, int C_in, int C_out, int H, int W);

/* dev/linear.cuh:5-11  matmul_forward2(cublas, out, inp, weight, bias, N, C, OC, block_size)
 * out (N,OC) = inp (N,C) . weight (OC,C)^T + bias */
int ub_matmul_forward2(float* out, const float* inp, const float* weight, const float* bias, int N, int C, int OC);
/* dev/linear.cuh:13-18  matmul_backward1(cublas, dinp, dweight, dbias, dout, inp, weight, N, C, OC); overwrite */
int ub_matmul_backward1(float* dinp, float* dweight, float* dbias, const float* dout, const float* inp,
                        const float* weight, int N, int C, int OC);

/* dev/groupnorm.cuh:3-7   groupnorm_forward(x, weight, bias, out, mean, rstd, B, C, H, W, n_groups); eps 1e-5 */
int ub_groupnorm_forward(const float* x, const float* weight, const float* bias, float* out, float* mean,
                         float* rstd, int B, int C, int H, int W, int n_groups);
/* dev/groupnorm.cuh:9-13  groupnorm_backward(...): dx overwritten; dweight/dbias ACCUMULATED (reference uses
 * atomicAdd into them, train_unet.cu:1926-1991) */
int ub_groupnorm_backward(const float* dout, const float* x, const float* mean, const float* rstd,
                          const float* weight, float* dx, float* dweight, float* dbias, int B, int C, int H, int W,
                          int n_groups);

/* dev/silu.cuh:3-13 */
int ub_silu_forward(const float* x, float* out, int N);
int ub_silu_backward(const float* dout, const float* x, float* dx, int N);

/* dev/add.cuh:3-13 */
int ub_add_forward(const float* a, const float* b, float* out, int N);
int ub_add_inplace_forward(const float* a, float* b, int N); /* b += a */

/* dev/upsample.cuh:3-15 (x is (B,C,H,W), out (B,C,2H,2W)) ; dev/avgpool.cuh:4-16 (x (B,C,H,W), out (B,C,H/2,W/2)) */
int ub_upsample_forward1(float* out, const float* x, int B, int C, int H, int W);
int ub_upsample_backward1(float* dx, const float* dout, int B, int C, int H, int W);
int ub_avgpool_2d_forward1(float* out, const float* x, int B, int C, int H, int W);
int ub_avgpool_2d_backward1(const float* dout, float* dx, int B, int C, int H, int W);

/* dev/concat_channel.cuh:4-15 */
int ub_concat_channel_forward(const float* x1, const float* x2, float* out, int B, int C1, int C2, int H, int W);
int ub_concat_channel_backward(const float* dout, float* dx1, float* dx2, int B, int C1, int C2, int H, int W);

/* dev/broadcast.cuh:4-14: x (N) -> out (N,H,W); backward sums over the last two dims */
int ub_broadcast_last_dims_forward(const float* x, float* out, int N, int H, int W);
int ub_broadcast_last_dims_backward(const float* dout, float* dx, int N, int H, int W);

/* dev/mse.cuh:1-3: loss is a device pointer to one float (overwritten) */
int ub_mse_forward(const float* inp, const float* y, float* loss, int N);
int ub_mse_backward(const float* inp, const float* y, float* dinp, int N);

/* dev/timestep_embedding.cuh:8-14 (freqs table is recomputed on the fly; no init/free needed) */
int ub_get_timestep_embeddings(const float* timesteps, float* out, int B, int dim, int max_period);

/* dev/attention.cuh:6-13  attention_forward1(cublas, out, qkvr, preatt, att, inp, B, T, C, NH, block_size)
 * inp (B,T,3C) in (B,T,3,NH,HS) order; out (B,T,C).  qkvr (3,B,NH,T,HS), preatt and att (B,NH,T,T) are filled as
 * the reference does (att = softmax probabilities, needed by attention_backward). fp32. */
int ub_attention_forward1(float* out, float* qkvr, float* preatt, float* att, const float* inp, int B, int T, int C,
                          int NH);
/* dev/attention.cuh:15-22  attention_backward(cublas, dinp, dqkvr, dpreatt, datt, scratch, dout, qkvr, att, ...) */
int ub_attention_backward(float* dinp, float* dqkvr, float* dpreatt, float* datt, float* scratch, const float* dout,
                          const float* qkvr, const float* att, int B, int T, int C, int NH);

/* ---- composite blocks (dev/resblock.cuh:71-131, dev/attention_block.cuh:78-112).  The structs carry the same
 * tensors, in the same order and with the same meaning, as ResBlockParameters / ResBlockActivations /
 * ResBlockBackwardActivations and AttentionParams / AttentionActs / AttentionBackwardActs (the size arrays of the
 * reference structs are bookkeeping of its arena planner and are not needed here); include/unet_b200_legacy.hpp maps
 * the reference structs onto them.  All tensors fp32, NCHW / row-major, device memory owned by the caller; every
 * intermediate the reference materialises is written (its unit tests compare them, dev/resblock.cu:542-630).
 * Backward clobbers the forward activations the reference clobbers (cv3_2, add2, ud_x, silu1, l_emb) and writes
 * dx / demb (resblock) or dinp (attention); conv / linear gradients are overwritten, GroupNorm gradients accumulate. */
typedef struct {
    float *gn1_w, *gn1_b, *cv3_1_w, *cv3_1_b, *l_emb_w, *l_emb_b, *gn2_w, *gn2_b, *cv3_2_w, *cv3_2_b, *res_cv1_w,
        *res_cv1_b;
} UbResBlockParams;
typedef struct {
    float *gn1, *gn1_mean, *gn1_rstd, *silu1, *ud_h, *ud_x, *cv3_1, *silu_emb, *l_emb, *broad_emb, *add1, *gn2,
        *gn2_mean, *gn2_rstd, *silu2, *cv3_2, *res_cv1, *add2;
    float *input, *emb;
} UbResBlockActs;
typedef struct {
    float *buf_BCemb, *buf_BCHoWo, *buf1_BCHW, *buf2_BCHW, *dout, *dx, *demb;
} UbResBlockBack;
int ub_resblock_forward(int C, int C_emb, int C_out, int B, int H, int W, int up, int down, int gn_n_groups,
                        const UbResBlockParams* params, const UbResBlockActs* acts);
int ub_resblock_backward(int C, int C_emb, int C_out, int B, int H, int W, int up, int down, int gn_n_groups,
                         const UbResBlockParams* params, const UbResBlockParams* grads, const UbResBlockActs* acts,
                         const UbResBlockBack* back);
typedef struct {
    float *gn_w, *gn_b, *qkv_w, *qkv_b, *proj_w, *proj_b;
} UbAttentionParams;
typedef struct {
    float *gn, *gn_mean, *gn_rstd, *perm1, *qkv1, *qkv2, *preatt, *att, *att_out, *proj, *perm2, *add;
    float* input;
} UbAttentionActs;
typedef struct {
    float *buf1_BCHW, *buf2_BCHW, *buf_B3CHW, *dqkvr, *dpreatt, *datt, *dout, *dinp;
} UbAttentionBack;
int ub_attention_block_forward(int B, int C, int H, int W, int HS, int gn_n_groups, const UbAttentionParams* params,
                               const UbAttentionActs* acts);
int ub_attention_block_backward(int B, int C, int H, int W, int HS, int gn_n_groups, const UbAttentionParams* params,
                                const UbAttentionActs* acts, const UbAttentionBack* back,
                                const UbAttentionParams* grads);

/* ---- native operand layout of the tensor-core path: NHWC bf16 activations, packed bf16 weights.  The reference has
 * no equivalent; a caller that keeps its activations in this layout skips the NCHW fp32 <-> NHWC bf16 conversion the
 * drop-in operators above have to do (this is what the trainer does internally).  ksize = 1 or 3.
 *   wf : bf16 [ksize*ksize][C_out][C_in]  (fprop pack)      wd : bf16 [ksize*ksize][C_in][C_out] (dgrad pack, flipped)
 * Needs C_in % 8 == 0 and C_out % 16 == 0 (wgrad: both % 64 == 0). */
int ub_nchw_to_nhwc_bf16(const float* x, void* y_bf16, int B, int C, int H, int W);
int ub_pack_conv_weight(const float* weight, void* wf_bf16, void* wd_bf16, int C_in, int C_out, int ksize);
int ub_conv2d_nhwc_forward(const void* x_bf16, const void* wf_bf16, const float* bias, void* out_bf16, int B, int H,
                           int W, int C_in, int C_out, int ksize);
int ub_conv2d_nhwc_dgrad(const void* dout_bf16, const void* wd_bf16, void* dx_bf16, int B, int H, int W, int C_in,
                         int C_out, int ksize);
/* dweight (C_out,C_in,k,k) fp32 and dbias (C_out) fp32 are overwritten */
int ub_conv2d_nhwc_wgrad(const void* dout_bf16, const void* x_bf16, float* dweight, float* dbias, int B, int H, int W,
                         int C_in, int C_out, int ksize);

/* GroupNorm (+ optional SiLU) in the native layout (what the trainer runs; replaces groupnorm_forward / groupnorm_backward
 * fused with silu_forward / silu_backward, dev/groupnorm.cuh:3-13, dev/silu.cuh:3-13): x, y, dy, dx are NHWC bf16 [B*H*W][C].
 * Single-pass "slab" kernels when the shape allows (C % 16 == 0; impl = 0), two-pass kernels otherwise / when impl = 1.
 * chsum [B][C][2] fp32 = per-(image, channel) sum and sum of squares: written by forward, read by backward.
 * backward: dy = dL/d act(gn(x)); dx = gn_bwd(dy) (+ add_in if non-NULL); dweight / dbias (C) are ACCUMULATED into, like
 * the reference's groupnorm_backward (train_unet.cu:1926-1991); scratch [B][C][2] fp32 is used by impl 1 only. */
int ub_groupnorm_nhwc_forward(const void* x_bf16, const float* weight, const float* bias, void* y_bf16, float* chsum,
                              int B, int H, int W, int C, int n_groups, int silu, int impl);
int ub_groupnorm_nhwc_backward(const void* x_bf16, const void* dy_bf16, const float* chsum, const float* weight,
                               const float* bias, const void* add_in_bf16, void* dx_bf16, float* dweight, float* dbias,
                               float* scratch, int B, int H, int W, int C, int n_groups, int silu, int impl);

/* Attention core in the native layout (what the trainer runs; replaces attention_forward1 / attention_backward,
 * dev/attention.cuh:6-22, without their permutes and (B,NH,T,T) matrices): qkv is NHWC bf16 [B*T][3C], channel order
 * [Q | K | V], each [NH][32]; out [B*T][C] bf16; lse [B][NH][T] fp32 (log2-domain logsumexp, saved for backward).
 * backward: dqkv [B*T][3C] bf16 from dout [B*T][C]; dsum [B][NH][T] is scratch.  Head size must be 32.
 * impl: 0 = tcgen05 kernels (needs T in {16,32,64,128,256} and an even NH), 1 = SIMT fp32 fallback (any T). */
int ub_attention_nhwc_forward(const void* qkv_bf16, void* out_bf16, float* lse, int B, int T, int C, int NH, int impl);
int ub_attention_nhwc_backward(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse,
                               void* dqkv_bf16, float* dsum, int B, int T, int C, int NH, int impl);

/* ------------------------------------------------------------------------------------------------------------
 * (2) Trainer
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct UbTrainer UbTrainer;

/* UnetConfig of train_unet.cu:3318-3336 (+ the literals hard-coded at :4842-4864), generalised. */
typedef struct {
    int B;               /* per-GPU batch */
    int C_in, C_model, C_out;
    int H, W;
    int max_period;      /* 1000 */
    int n_levels;        /* 4 */
    int channel_mult[8]; /* {1,2,3,4} */
    int n_res_blocks;    /* 2 */
    int att_start_level; /* 2 : attention at levels >= this */
    int head_size;       /* 32 */
    int gn_n_groups;     /* 32 */
    int n_timesteps;     /* 1000, linear beta schedule 1e-4 .. 0.02 (train_unet.cu:3131-3147) */
    unsigned long long seed;
    int use_cuda_graph;  /* capture the whole step in one CUDA graph */
    int compute_dinput;  /* also compute dL/d(x_t) (the reference's unet_backward does, dev/unet_test.cu:2103); 0 */
    int random_flip;     /* on-GPU augmentation: every image of a batch is mirrored left-right with probability 1/2 before
                            q-sample (the PyTorch loader's random_flip, train_unet.py:508-536: arr[:, ::-1]; the reference's
                            C loader has none).  Decisions are Philox draws keyed by (seed, step, image); 0 */
    int num_classes;     /* > 0: class-conditional model (dev/unet.py:142,174-175,301-303 `num_classes`): a label-embedding
                            table (num_classes, 4*C_model) follows the time MLP in the parameter order and
                            emb = time_embed(t) + label_emb[y]; labels come from ub_trainer_set_labels; 0 */
    float ema_rate;      /* > 0: keep an exponential moving average of the parameters, ema = rate*ema + (1-rate)*p
                            after every AdamW update, fused into the AdamW kernel (guided-diffusion's update_ema;
                            the option the reference carries as `ema_rate`, train_unet.py:708); 0 */
    int resblock_updown; /* 1: the Downsample / Upsample layers become ResBlock(down=True) / ResBlock(up=True)
                            (dev/unet.py:147,205-222,271-284 `resblock_updown`, dev/resblock.py:78-86,125-128: average
                            pool / nearest upsample of both the main and the skip path inside the block); each adds the
                            10 parameter tensors of a ResBlock at the place of the parameter-free layer; 0 */
    int use_scale_shift_norm; /* 1: FiLM-like conditioning (dev/unet.py:146 `use_scale_shift_norm`, dev/resblock.py:211,
                            243-247): every ResBlock's embedding projection has 2*Cout outputs [scale | shift] and the second
                            GroupNorm computes gn(h) * (1 + scale) + shift instead of gn(h + emb); 0 */
    float dropout;       /* > 0: dropout with this probability between SiLU and the second conv of every ResBlock
                            (guided-diffusion's out_layers; the reference carries the option commented out,
                            dev/resblock.py:51,61, dev/unet.py:116,140) in training steps; predict / sample run without.
                            Masks are Philox draws keyed by (seed, step, block, element), never stored; 0 */
} UbConfig;

void ub_default_config(UbConfig* cfg);
/* number of parameters for a config (326 tensors / 20 494 211 floats for the default) */
size_t ub_num_params(const UbConfig* cfg);

int ub_trainer_create(UbTrainer** out, const UbConfig* cfg, int device);
void ub_trainer_destroy(UbTrainer* t);

/* checkpoint, train_unet.cu:4762-4911 / train_unet.py:768-795: int32[256] header {12345678, B, C_in, C_model, C_out,
 * H, W, max_period, has_adamw, has_rng} + fp32 params [+ m + v].  The rng flag is always written 0 (the reference's
 * blob is a raw curandState dump); header[10] additionally stores the AdamW step count (the reference forgets it),
 * marked valid by header[11] = 0x55425354.  The reference's C writer leaves words 10..255 uninitialised
 * (train_unet.cu:4764), so ub_trainer_load trusts header[10] only behind that marker and otherwise restarts the count
 * at 0 -- what the reference's own resume does; ub_trainer_set_step supplies a count from elsewhere. */
int ub_trainer_load(UbTrainer* t, const char* path);
int ub_trainer_save(UbTrainer* t, const char* path, int with_adamw);
int ub_trainer_set_step(UbTrainer* t, int step);
/* read the {B, C_in, C_model, C_out, H, W, max_period} header of a checkpoint into cfg (other fields untouched) */
int ub_read_checkpoint_header(const char* path, UbConfig* cfg);

/* host <-> device parameter access, reference parameter order (forward layer order) */
int ub_trainer_set_params(UbTrainer* t, const float* host, size_t n);
int ub_trainer_get_params(UbTrainer* t, float* host, size_t n);
int ub_trainer_get_grads(UbTrainer* t, float* host, size_t n);
/* moving average of the parameters (cfg.ema_rate > 0): starts as a copy of the parameters (set_params / load) */
int ub_trainer_get_ema(UbTrainer* t, float* host, size_t n);
int ub_trainer_set_ema(UbTrainer* t, const float* host, size_t n);
int ub_trainer_save_ema(UbTrainer* t, const char* path); /* parameters-only .bin of the averaged weights */
/* class labels y (B ints in [0, num_classes)) of the batch(es) that follow (cfg.num_classes > 0, dev/unet.py:301-303);
 * used by forward_backward / train_step / predict / sample until set again */
int ub_trainer_set_labels(UbTrainer* t, const int* labels_host, size_t n);
/* the dropout mask (0 / 1 bytes, B x H x W x C in NHWC order) the LAST training step applied in ResBlock `block`
 * (forward order), regenerated from its Philox key -- for parity checks (cfg.dropout > 0) */
int ub_trainer_get_dropout_mask(UbTrainer* t, int block, unsigned char* host, size_t n);
int ub_trainer_get_output(UbTrainer* t, float* host, size_t n); /* last forward's eps prediction (B,C_out,H,W) */
/* the flip decisions (0 / 1 per image) the last step took, when cfg.random_flip is set */
int ub_trainer_get_flips(UbTrainer* t, int* host, size_t n);
int ub_trainer_get_batch(UbTrainer* t, float* host, size_t n);  /* the x0 batch the last step consumed (B,C_in,H,W) */
int ub_trainer_get_dinput(UbTrainer* t, float* host, size_t n); /* dL/d(x_t) (B,C_in,H,W); needs cfg.compute_dinput */

/* One forward + backward (unet_forward + unet_backward, train_unet.cu:4335-4701) on a HOST batch x0 (B,C_in,H,W).
 * t_host (B) and noise_host (B,C_in,H,W) may be NULL: then timesteps / noise are drawn on the device (Philox).
 * Gradients are left in the gradient arena (all-reduced over the data-parallel group if one is attached).
 * loss_out receives the mean-squared error (device->host copy of 4 bytes). */
int ub_trainer_forward_backward(UbTrainer* t, const float* x0_host, const float* t_host, const float* noise_host,
                                float* loss_out);
/* unet_update (train_unet.cu:4738-4757): AdamW with bias correction; zeroes the gradients. */
int ub_trainer_update(UbTrainer* t, float lr, float beta1, float beta2, float eps, float weight_decay);
/* The whole loop body of train_unet.cu:5019-5037 in one call (one CUDA-graph launch when enabled):
 * H2D batch copy, timesteps/noise, q-sample, forward, loss, backward, gradient all-reduce, AdamW.  loss_out may be
 * NULL (then no device->host copy / sync happens; use ub_trainer_sync + ub_trainer_last_loss). */
int ub_trainer_train_step(UbTrainer* t, const float* x0_host, const float* t_host, const float* noise_host, float lr,
                          float beta1, float beta2, float eps, float weight_decay, float* loss_out);
/* Announce the batch the NEXT ub_trainer_train_step will be called with (a page-locked host buffer; NULL cancels).
 * The train_step that follows this call enqueues the announced batch's H2D copy on a copy stream as soon as its own
 * step is launched, so the transfer runs under that step's kernels (the reference copies synchronously before every
 * step, train_unet.cu:5021-5024); the next train_step, called with the same pointer, then starts from the device
 * copy.  The announced buffer must stay unchanged until that next call returns.  A pageable pointer, or a next call
 * with another pointer, silently falls back to the ordinary path. */
int ub_trainer_set_next_batch(UbTrainer* t, const float* next_x0_host);
/* Same step with the batch already resident on the device (bench.py's device-resident `value`). */
int ub_trainer_train_step_device(UbTrainer* t, const float* x0_dev, float lr, float beta1, float beta2, float eps,
                                 float weight_decay);
int ub_trainer_sync(UbTrainer* t);
int ub_trainer_last_loss(UbTrainer* t, float* loss_out);
void* ub_trainer_stream(UbTrainer* t); /* cudaStream_t the trainer launches on */
/* kernels launched per train step (counted while building the step) */
int ub_trainer_launches_per_step(UbTrainer* t);
/* forward only on a device batch that is ALREADY x_t (for sampling, generate.py:29-52): out_dev (B,C_out,H,W) */
int ub_trainer_predict(UbTrainer* t, const float* xt_host, const float* t_host, float* out_host);

/* DDPM ancestral sampling, the loop of generate.py:29-79 on the device: for t = t_start .. t_end (descending)
 *   eps = model(x_t, t);  mu = (x_t - beta_t / sqrt(1 - acp_t) * eps) / sqrt(1 - beta_t);
 *   x_{t-1} = mu + sqrt((1 - acp_{t-1}) / (1 - acp_t) * beta_t) * z,     beta_t = betas[t-1], acp_t = acp[t-1]
 * on all B images of the trainer at once.  generate.py runs t = n_timesteps-1 .. 2 (pass t_start = -1, t_end = -1 for
 * exactly that).  x_init_host (B,C_in,H,W) may be NULL: x_T ~ N(0,1) from `seed`.  noise_host may be NULL (z from the
 * device Philox stream of `seed`), else it holds one (B,C_in,H,W) draw per iteration, first iteration first (parity
 * runs inject the oracle's draws).  out_host (B,C_in,H,W) receives the final x.  One CUDA graph per iteration. */
int ub_trainer_sample(UbTrainer* t, const float* x_init_host, int t_start, int t_end, const float* noise_host,
                      unsigned long long seed, float* out_host);

/* ---- dataset reader: prepare_data.py:20-38 format (int32[256] header {20240620, N, C, H, W} + float32 images),
 * semantics of DataLoader / dataloader_next_batch (train_unet.cu:3035-3099): sequential batches of B images,
 * wrapping to the start when fewer than B images remain.  Batches are prefetched by a background thread into three
 * page-locked buffers, so the returned pointer can be handed to ub_trainer_train_step directly (async H2D), and the
 * batch of one call stays intact through the NEXT call: the caller can hold the current batch and the following one
 * (ub_trainer_set_next_batch) at the same time.
 * Data parallel: rank r of `world` reads global batches r, r + world, ... (each global step consumes world batches). */
typedef struct UbDataLoader UbDataLoader;
int ub_dataloader_open(UbDataLoader** out, const char* path, int B, int rank, int world);
int ub_dataloader_info(UbDataLoader* d, int* n_imgs, int* C, int* H, int* W, long long* batches_per_epoch);
const float* ub_dataloader_next(UbDataLoader* d);  /* valid until the call after the next one; NULL on I/O error */
void ub_dataloader_reset(UbDataLoader* d);
void ub_dataloader_close(UbDataLoader* d);

/* ---- per-kernel-class timing of one training step (CUDA events around every launch, eager replay of the tape).
 * Used by bench.py for the live roofline numbers. */
#define UB_KIND_CONV 0    /* igemm_conv_kernel: 3x3 / 1x1 conv fprop + dgrad, attention qkv/proj GEMMs (tensor) */
#define UB_KIND_WGRAD 1   /* igemm_wgrad_kernel + split-K reduce (tensor) */
#define UB_KIND_NORM 2    /* GroupNorm(+SiLU) forward / backward (HBM) */
#define UB_KIND_ATTN 3    /* attention core forward / backward */
#define UB_KIND_ELTWISE 4 /* concat / pool / add / column sums (HBM) */
#define UB_KIND_SMALL 5   /* embedding MLPs, 3-channel convs, loss */
#define UB_KIND_OPTIM 6   /* diffusion noise + q-sample, AdamW, weight packing (HBM) */
#define UB_NUM_KINDS 7
typedef struct {
    double ms[UB_NUM_KINDS];       /* device time per class, average per step */
    double flops[UB_NUM_KINDS];    /* algorithmic FLOPs per step */
    double bytes[UB_NUM_KINDS];    /* algorithmic bytes per step */
    int launches[UB_NUM_KINDS];    /* kernel launches per step */
    double total_ms;               /* sum over classes */
} UbProfile;
int ub_trainer_profile(UbTrainer* t, int reps, UbProfile* out);

/* ---- data parallel: one process per GPU; rank 0 creates the id, everyone attaches ---- */
#define UB_NCCL_ID_BYTES 128
int ub_nccl_get_unique_id(void* id_out /* UB_NCCL_ID_BYTES */);
/* n_buckets: 0 = the library's bucket count (8 + a tail bucket, fixed when the trainer is built; UB_BUCKETS overrides) */
int ub_trainer_attach_dp(UbTrainer* t, int rank, int world, const void* nccl_id, int n_buckets);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H */
