// Host side of the training path: model graph, arena planner, step builder (eager or one CUDA graph),
// checkpoint I/O, NCCL data parallelism.  Mirrors the responsibilities of train_unet.cu:3318-4911
// (Unet/UnetConfig, unet_make_ptrs_and_count_memory, unet_forward/backward, unet_update, save/load) without
// following its structure: the graph is built once as a tape of launch closures over bump-allocated NHWC bf16
// tensors, every tcgen05 launch carries pre-encoded TMA descriptors, and the whole step replays as one graph.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/unet_b200.h"
#include "../csrc/attn_tc.cuh"
#include "../csrc/igemm.cuh"
#include "../csrc/launch.cuh"
#include "../csrc/misc_ops.cuh"
#include "../csrc/nhwc_ops.cuh"
#include "host_common.h"

namespace ub {
void igemm_init();
void attn_init();
}
extern "C" void ub_count_launches(unsigned long long n);

using namespace ub;

// ======================================================================================================
// NCCL through dlopen (no link-time dependency; inside torchrun the already loaded libnccl.so.2 is reused)
// ======================================================================================================
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.AllReduce =
        (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
    return api;
}
constexpr int kNcclFloat = 7;  // ncclFloat32
constexpr int kNcclSum = 0;    // ncclSum
}  // namespace

// ======================================================================================================
// small utilities
// ======================================================================================================
struct View {  // NHWC bf16 tensor view
    bf16* p = nullptr;
    int ld = 0;
    int C = 0, H = 0, W = 0;
    float* cs = nullptr;  // per-(image, channel) sum / sumsq [B][C][2], filled by the tensor's producer (or null)
};

struct DeviceArena {
    uint8_t* base = nullptr;
    size_t cap = 0, off = 0;
    bool counting = true;  // first pass: only count
    void* alloc(size_t bytes) {
        off = (off + 255) & ~size_t(255);
        void* r = counting ? nullptr : base + off;
        off += bytes;
        return r;
    }
};

struct ParamRef {
    size_t off;
    size_t n;
};

struct UbTrainer {
    UbConfig cfg;
    int device = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    // data parallel: AdamW + weight re-pack of an all-reduced bucket run here, behind the bucket's all-reduce, so that
    // the communication stream is free for the next bucket's all-reduce (the last ones are the step's serial tail)
    cudaStream_t opt_stream = nullptr;
    bool opt_stream_dirty = false;
    // Weight gradients (wgrad GEMMs, bias column sums, embedding-projection backward) feed nothing but the optimizer:
    // they run on a side stream, concurrently with the dgrad -> GroupNorm -> dgrad critical path of backward (in the
    // captured graph this is a parallel branch).  Both kinds of kernel are latency-bound on their own.
    cudaStream_t side_stream = nullptr;
    // Auxiliary main-priority stream for a short fork inside ONE main-stream op (fork and join within the op): the
    // attention backward's dkv kernel beside its dq kernel (attn_tc_bwd).
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
    bool attn_concurrent = true;
    // Micro-batch pipelining: the main-stream ops that work image by image (conv fprop / dgrad, GroupNorm, attention,
    // data movement) are emitted once per micro-batch of B / n_mb images -- batch-major views of the same buffers -- and
    // micro-batch h > 0 runs on mb_streams[h - 1], a parallel branch of the captured graph.  The step is ~40 % batch-
    // independent latency at B = 32 (5.17 ms at B = 32, 3.69 at 16, 2.96 at 8: profiles/r02_microbatch.txt); the idea
    // was that two independent chains hide each other's latency.  They do not (5.30 ms with two chains): an experiment
    // switch (UB_MICROBATCH), off by default.  Ops that need the whole batch (weight gradients on the side stream, loss,
    // optimizer) join the chains first.
    int n_mb = 1;
    bool mb_fwd_only = false;  // UB_MICROBATCH_FWD: the chains exist in the forward pass only
    int mb_max_hw = 0;         // UB_MICROBATCH_LOWRES: > 0 = chains only at levels of at most this many pixels per image
    std::vector<cudaStream_t> mb_streams;
    std::vector<cudaEvent_t> side_events;
    size_t side_ev_next = 0;
    bool use_side = true;
    // flat fp32 arenas in reference order
    size_t nparams = 0;
    float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
    // cfg.ema_rate > 0: exponential moving average of the parameters, updated by the AdamW kernel (guided-diffusion's
    // update_ema; train_unet.py:708 carries the option, the reference trainer ignores it); starts as a copy of params
    float* ema = nullptr;
    // cfg.num_classes > 0: class labels of the current batch (ub_trainer_set_labels) and the label-embedding table
    int* labels = nullptr;
    size_t label_w_off = 0;
    // cfg.dropout > 0: {p bits, step, seed} read by the GroupNorm kernels that regenerate the masks (DropArgs), and the
    // (C, H, W) of every ResBlock's dropped-out tensor (ub_trainer_get_dropout_mask)
    unsigned* drop_ctl = nullptr;
    struct DropInfo {
        int C, H, W;
    };
    std::vector<DropInfo> drop_layers;
    std::vector<ParamRef> tensors;  // every parameter tensor in order
    // everything else lives in one bump arena
    DeviceArena arena;
    // zero-on-every-step region (atomically accumulated small buffers)
    uint8_t* zero_base = nullptr;
    size_t zero_bytes = 0;
    DeviceArena zarena;
    // io buffers
    float *x0 = nullptr, *xt = nullptr, *noise = nullptr, *tsteps = nullptr, *out = nullptr, *dout = nullptr;
    float* dxt = nullptr;  // dL/d(x_t), only with cfg.compute_dinput
    int* flips = nullptr;  // per-image flip decisions of the last step, only with cfg.random_flip
    float *loss = nullptr, *sqrt_ac = nullptr, *sqrt_1mac = nullptr, *betas = nullptr;
    int* samp_state = nullptr;  // {t, iteration} of the sampling loop
    cudaGraphExec_t samp_graph = nullptr;
    bool samp_graph_noise = false;
    float* samp_z = nullptr;  // injected noise of the current iteration (graph reads a fixed address)
    int* step_dev = nullptr;
    float* h_loss = nullptr;  // pinned
    // Host batches are staged through the trainer's OWN page-locked double buffer: the caller's buffer (pageable or
    // pinned, e.g. a data-loader slot that is refilled by a prefetch thread) may be reused as soon as the call returns,
    // like cudaMemcpy from pageable memory -- the asynchronous copy that reads it may still be queued behind the previous
    // step's graph.  A slot is rewritten only after the copies that last read it have completed (event per slot), so the
    // host runs at most two steps ahead of the device.
    float *hs_x0[2] = {nullptr, nullptr}, *hs_noise[2] = {nullptr, nullptr}, *hs_t[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    bool stage_used[2] = {false, false};
    unsigned stage_next = 0;
    // Next-batch prefetch (ub_trainer_set_next_batch): the H2D copy of the batch the NEXT train_step will be called with
    // runs on a copy stream under the current step's kernels; that step then starts with a 1.5 MB device copy instead of
    // a PCIe transfer.  One device buffer: a new prefetch waits (event) for the copy that consumed the previous one.
    const float* next_hint = nullptr;  // announced by the caller, picked up by the train_step that follows
    const float* pref_src = nullptr;   // host pointer whose batch x0_pref holds once ev_pref has completed
    float* x0_pref = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_pref = nullptr, ev_pref_free = nullptr;
    bool pref_free_pending = false;
    float hp_last[5] = {0, 0, 0, 0, 0};  // what hp_dev holds (uploaded again only when a value changes)
    bool hp_uploaded = false;
    // scratch of the 3-channel weight-gradient fallback kernels (the tcgen05 wgrads accumulate with TMA reduce-adds, no workspace)
    float* small_scratch = nullptr;
    size_t small_scratch_floats = size_t(1) << 20;
    // tables (device)
    PackEntry* pack_table = nullptr;
    int n_pack = 0, pack_max_tiles = 0;
    std::vector<PackEntry> h_pack;
    std::vector<size_t> pack_off;  // parameter offset of each pack entry (ascending)
    WgradFinalizeEntry* fin_table = nullptr;
    std::vector<WgradFinalizeEntry> h_fin;  // in backward order
    SmallLinear *emb_table = nullptr, *temb_table = nullptr;
    std::vector<SmallLinear> h_emb, h_temb;
    int emb_max_oc = 0;
    // tapes
    typedef std::function<void(cudaStream_t)> Op;
    std::vector<Op> fwd_ops, bwd_ops;  // executed in push order
    struct OpInfo {
        int kind;       // UB_KIND_*
        int launches;
        double flops;   // algorithmic FLOPs (tensor kernels) ...
        double bytes;   // ... or algorithmic bytes (memory-bound kernels)
        int side;       // 0 = main stream, 1 = weight-gradient branch (side stream), 2 = join marker
        std::string label;
        int half = -1;  // >= 0: this op is micro-batch `half` of a per-image op (runs on that micro-batch's stream)
    };
    std::vector<OpInfo> fwd_info, bwd_info;
    int launches_fwd = 0, launches_bwd = 0, launches_misc = 0;
    // hyper-parameters baked into the captured graph
    float lr = 1e-4f, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f, wd = 0.f;
    bool gen_t = true, gen_noise = true;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_valid = false;
    bool graph_gen_t = false, graph_gen_noise = false, graph_h2d = false;
    float g_lr = 0, g_b1 = 0, g_b2 = 0, g_eps = 0, g_wd = 0;
    bool have_grads = false;
    int host_step = 0;
    // data parallel
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, n_buckets = getenv("UB_BUCKETS") ? atoi(getenv("UB_BUCKETS")) : 8;
    bool comm_off = false;  // rank-local eager replays (profiling) must not enqueue collectives
    // Optimizer per gradient bucket: as soon as a bucket's gradients are final (and all-reduced), AdamW and the weight
    // re-pack of that parameter range run on the branch that finalised them, under the rest of backward; only the
    // last bucket is left for the end of the step.  The tape ops read the current step's hyper-parameters from here.
    bool opt_in_tape = false;  // true while a step WITH an update is being enqueued
    bool opt_overlap_ok = true;
    float o_lr = 0, o_b1 = 0, o_b2 = 0, o_eps = 0, o_wd = 0;
    // {lr, beta1, beta2, eps, weight_decay} in device memory: the captured graph's AdamW nodes read them from here, so a
    // learning-rate schedule changes 20 bytes per step instead of forcing a re-capture of ~450 nodes
    float* hp_dev = nullptr;
    const float* hp_active = nullptr;  // = hp_dev while a graph is being captured / replayed, null for eager steps
    size_t opt_done_lo = 0;   // parameters >= this offset were updated by the tape
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    std::vector<cudaEvent_t> bucket_events;
    std::vector<size_t> bucket_bounds;  // param offsets, descending
    // emb bookkeeping
    float *sin_emb = nullptr, *h0 = nullptr, *emb = nullptr, *semb = nullptr, *d_embact = nullptr, *demb = nullptr,
          *d_h0act = nullptr, *dh0 = nullptr;
    std::string err;
};

static thread_local std::string g_err;
static void set_err(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* ub_host_last_error() { return g_err.c_str(); }
void ub_host_set_error(const char* s) { g_err = s; }

#define CUDA_TRY(x)                                                                    \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            set_err("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return UB_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

// ======================================================================================================
// graph construction: forward tape + per-node backward builders, walked in reverse
// ======================================================================================================
namespace {

constexpr int kMaxEmbEntries = 96;

struct Builder {
    UbTrainer* T;
    const UbConfig& c;
    size_t poff = 0;  // running parameter offset (reference order)
    int B;
    int plan_errors = 0;

    struct Node {
        std::function<View(View)> bwd;  // emits backward ops, returns gradient w.r.t. the node's main input
        bool pushed = false;            // output was pushed on the skip stack
        View out;                       // output view (shape used for the skip-gradient add)
        size_t param_begin = 0;         // parameters of nodes >= this one start here
    };
    std::vector<Node> nodes;
    std::vector<View> skipgrad;  // per node: gradient arriving from the up path
    // UB_MICROBATCH_LOWRES = hw: micro-batch chains only at levels of at most hw pixels per image (the 8x8 / 16x16 levels
    // run on less than half of the SMs and are bound by per-kernel latency); node_nh[i] = chains of node i
    std::vector<int> node_nh;
    void level_mb(int H, int W) {
        if (T->mb_max_hw <= 0) return;
        nh = (H * W <= T->mb_max_hw && T->n_mb > 1 && B % T->n_mb == 0) ? T->n_mb : 1;
        Bh = B / nh;
    }
    void sync_nh() { node_nh.resize(nodes.size(), nh); }

    // micro-batches (see UbTrainer::n_mb): nh chains of Bh images each
    int nh = 1, Bh = 0;
    int cur_half = -1;  // set while the ops of one micro-batch are being emitted (split())
    explicit Builder(UbTrainer* t) : T(t), c(t->cfg), B(t->cfg.B) {
        nh = t->n_mb > 0 && B % t->n_mb == 0 ? t->n_mb : 1;
        Bh = B / nh;
        level_mb(t->cfg.H, t->cfg.W);  // (UB_MICROBATCH_LOWRES: the full-resolution head of the step runs as one chain)
    }
    // emit the same op for every micro-batch: fn(h) pushes the ops of micro-batch h (views offset with mb())
    template <class Fn>
    void split(Fn fn) {
        std::string lbl = next_label;
        for (int h = 0; h < nh; ++h) {
            cur_half = nh > 1 ? h : -1;
            next_label = lbl;
            fn(h);
        }
        cur_half = -1;
        next_label.clear();
    }
    // micro-batch h of a batch-major view / per-image fp32 table with `per_img` floats per image
    View mb(View v, int h) const {
        if (v.p) v.p += size_t(h) * Bh * v.H * v.W * v.ld;
        if (v.cs) v.cs += size_t(h) * Bh * v.C * 2;
        return v;
    }
    float* mbf(float* p, int h, size_t per_img) const { return p ? p + size_t(h) * Bh * per_img : nullptr; }

    size_t take(size_t n) {
        size_t o = poff;
        poff += n;
        if (real()) T->tensors.push_back({o, n});
        return o;
    }
    float* P(size_t off) const { return T->params + off; }
    float* G(size_t off) const { return T->grads + off; }

    View act(int C, int H, int W) {
        View v;
        v.p = (bf16*)T->arena.alloc(size_t(B) * H * W * C * sizeof(bf16));
        v.ld = C, v.C = C, v.H = H, v.W = W;
        return v;
    }
    float* f32(size_t n) { return (float*)T->arena.alloc(n * sizeof(float)); }
    float* zf32(size_t n) { return (float*)T->zarena.alloc(n * sizeof(float)); }
    bf16* packbuf(size_t n) { return (bf16*)T->arena.alloc(n * sizeof(bf16)); }

    bool real() const { return !T->arena.counting; }
    void F(UbTrainer::Op op, int launches = 1, int kind = UB_KIND_SMALL, double flops = 0, double bytes = 0,
           int side = 0) {
        if (!real()) return;
        T->fwd_ops.push_back(std::move(op)), T->launches_fwd += launches;
        T->fwd_info.push_back({kind, launches, flops, bytes, side, next_label, side == 0 ? cur_half : -1}), next_label.clear();
    }
    // forward: the time-embedding chain runs on the side stream until its first consumer (conv1 of the first ResBlock)
    bool fwd_side_pending = false;
    void join_side_fwd() {
        if (fwd_side_pending) F([](cudaStream_t) {}, 0, UB_KIND_SMALL, 0, 0, 2);
        fwd_side_pending = false;
    }
    void Bk(UbTrainer::Op op, int launches = 1, int kind = UB_KIND_SMALL, double flops = 0, double bytes = 0,
            int side = 0) {
        if (!real()) return;
        T->bwd_ops.push_back(std::move(op)), T->launches_bwd += launches;
        T->bwd_info.push_back({kind, launches, flops, bytes, side, next_label, side == 0 ? cur_half : -1}), next_label.clear();
    }
    std::string next_label;  // optional description of the next op pushed (per-op profile dump)
    void label(const char* fmt, ...) {
        char buf[160];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        next_label = buf;
    }
    // the main stream waits for everything the weight-gradient branch has been given so far
    void join_side() { Bk([](cudaStream_t) {}, 0, UB_KIND_SMALL, 0, 0, 2); }
    double act_bytes(int C, int H, int W) const { return 2.0 * B * H * W * C; }

    struct Packed {
        bf16 *wf = nullptr, *wd = nullptr;
    };
    Packed pack(size_t woff, int Cout, int Cin, int ntaps) {
        Packed pk;
        pk.wf = packbuf(size_t(ntaps) * Cout * Cin);
        pk.wd = packbuf(size_t(ntaps) * Cout * Cin);
        if (real()) {
            T->h_pack.push_back({P(woff), pk.wf, pk.wd, Cout, Cin, ntaps});
            T->pack_off.push_back(woff);
            int tiles = ((Cout + 31) / 32) * ((Cin + 31) / 32);
            if (tiles > T->pack_max_tiles) T->pack_max_tiles = tiles;
        }
        return pk;
    }

    void conv_op(bool fwd, std::vector<ConvSegDesc> segs, int H, int W, int Cout, ConvEpilogue ep) {
        if (!real()) return;
        // measured per layer (profiles/r01_ops_rows_vs_basic.txt, r01_ops_rows_vs_basic_final.txt): the persistent
        // kernel wins (a) where the basic one is bound by its per-128-pixel fixed costs -- the 64-pixel-wide levels with
        // <= 64 output channels -- and (b) at wide levels with >= 192 output channels, where the basic kernel's N tile
        // is 160-192 columns: one CTA per SM, i.e. a single MMA issue stream, against the row-tile kernel's two.
        // Elsewhere (N <= 128: two CTAs per SM; 8x8 / 16x16: too few row tiles) the basic kernel is faster.
        // UB_ROWS=0 / 2 forces never / always.
        static const int rows_mode = getenv("UB_ROWS") ? atoi(getenv("UB_ROWS")) : 1;
        const bool rows_pref = (W >= 64 && (Cout <= 64 || Cout >= 192)) || (W >= 32 && Cout >= 192 && segs[0].ntaps == 9);
        // A one-tap conv without a hook (the 1x1 skip-conv dgrads) is all epilogue.  Measured (profiles/r02_plain_1x1.txt):
        // the row-tile kernel stays ahead of the basic kernel there even with the latter's staged tile + TMA store
        // (64 -> 192 at 64x64: 30.6 vs 39.7 us), so the rule above stands; UB_PLAIN_1X1_BASIC=1 routes them to the basic kernel.
        static const bool plain_basic = getenv("UB_PLAIN_1X1_BASIC") && atoi(getenv("UB_PLAIN_1X1_BASIC")) != 0;
        const bool plain_1x1 = segs.size() == 1 && segs[0].ntaps == 1 && !ep.stats && !ep.gn_x && plain_basic;
        const bool no_rows = rows_mode == 0 || (rows_mode == 1 && (!rows_pref || plain_1x1));
        double k = 0, bytes = act_bytes(Cout, H, W);
        for (auto& sg : segs) k += double(sg.ntaps) * sg.Cin, bytes += act_bytes(sg.Cin, H, W) + 2.0 * sg.ntaps * sg.Cin * Cout;
        const double flops = 2.0 * B * H * W * Cout * k;
        // one launch per micro-batch: batch-major views of the same buffers (operands, output, residual, per-image
        // embedding vector / GroupNorm statistics)
        split([&](int h) {
            std::vector<ConvSegDesc> sh = segs;
            for (auto& sg : sh) sg.x += size_t(h) * Bh * H * W * sg.ldx;
            ConvEpilogue eh = ep;
            const size_t opix = size_t(h) * Bh * H * W;
            eh.out = (bf16*)ep.out + opix * (ep.ldo ? ep.ldo : Cout);
            if (ep.residual) eh.residual = ep.residual + opix * (ep.ldr ? ep.ldr : Cout);
            eh.rowvec = mbf(const_cast<float*>(ep.rowvec), h, Cout);
            eh.stats = mbf(ep.stats, h, size_t(Cout) * 2);
            if (ep.gn_x) {
                eh.gn_x = ep.gn_x + opix * ep.gn_ldx;
                eh.gn_chsum = mbf(const_cast<float*>(ep.gn_chsum), h, size_t(Cout) * 2);
                eh.gn_S = mbf(ep.gn_S, h, size_t(Cout) * 2);
            }
            UbTrainer::Op op;
            // persistent row-tile kernel for the 16x16 .. 128x128 levels, one-CTA-per-128-pixels kernel otherwise
            IgemmRowsParams pr;
            if (!no_rows && igemm_rows_eligible(Bh, H, W, Cout) &&
                igemm_rows_plan(&pr, sh.data(), int(sh.size()), Bh, H, W, Cout, eh, sm_count()) == 0) {
                op = [pr](cudaStream_t st) { igemm_rows_launch(pr, st); };
                label("conv%s %s Cin=%d%s Cout=%d %dx%d rows BN=%d stages=%d/%d%s%s", segs[0].ntaps == 9 ? "3x3" : "1x1",
                      fwd ? "fwd" : "dgrad", segs[0].Cin, segs.size() > 1 ? "+skip" : "", Cout, H, W, pr.BN, pr.a_stages,
                      pr.w_stages, ep.stats ? " +stats" : "", ep.gn_x ? " +gnbwd" : "");
            } else {
                // (shared: a GroupNorm that consumes the output may still extend the plan, see gn_finish)
                auto pp = std::make_shared<IgemmConvParams>();
                IgemmConvParams& p = *pp;
                int r = igemm_conv_plan(&p, sh.data(), int(sh.size()), Bh, H, W, Cout, eh);
                if (r) {
                    set_err("igemm_conv_plan failed (%d) for %dx%d Cout=%d Cin0=%d", r, H, W, Cout, segs[0].Cin);
                    plan_errors++;
                    return;
                }
                if (nh == 1 && p.gnf_cluster && (ep.stats || ep.gn_x)) gn_finish[ep.out] = pp;
                op = [pp](cudaStream_t st) { igemm_conv_launch(*pp, st); };
                label("conv%s %s Cin=%d%s Cout=%d %dx%d BN=%d stages=%d%s%s%s", segs[0].ntaps == 9 ? "3x3" : "1x1",
                      fwd ? "fwd" : "dgrad", segs[0].Cin, segs.size() > 1 ? "+skip" : "", Cout, H, W, p.BN, p.stages,
                      ep.stats ? " +stats" : "", ep.gn_x ? " +gnbwd" : "", p.ms == 2 ? " +tmast" : p.ms ? " +ms" : "");
            }
            if (fwd)
                F(op, 1, UB_KIND_CONV, flops / nh, bytes / nh);
            else
                Bk(op, 1, UB_KIND_CONV, flops / nh, bytes / nh);
        });
    }
    // Weight (and bias) gradient of a conv / linear layer on the side stream.  Every CTA adds its tile into an fp32
    // accumulator with TMA reduce-add boxes; 3x3 layers accumulate in [tap][o][c] and are transposed into the reference
    // (Cout, Cin, 3, 3) layout by one finalize launch per gradient bucket (fin_flush), 1x1 layers accumulate straight
    // into the gradient arena.  db (and db2) = sum over pixels of dy, from an extra ones-column MMA in the same kernel.
    int fin_done = 0;
    void wgrad_op(View dy, View x, int Cin, int Cout, int ntaps, float* dw, float* db = nullptr, float* db2 = nullptr) {
        float* acc = ntaps == 9 ? f32(size_t(9) * Cout * Cin) : dw;
        if (!real()) return;
        static const int wg_sms = getenv("UB_WGRAD_SMS") ? atoi(getenv("UB_WGRAD_SMS")) : sm_count();
        IgemmWgradParams p;
        int r = igemm_wgrad_plan_acc(&p, dy.p, dy.ld, x.p, x.ld, B, x.H, x.W, Cin, Cout, ntaps, acc, db, db2, wg_sms);
        if (r) {
            set_err("igemm_wgrad_plan failed (%d) for %dx%d %d->%d", r, x.H, x.W, Cin, Cout);
            plan_errors++;
            return;
        }
        if (ntaps == 9) T->h_fin.push_back({acc, dw, Cout, Cin});
        label("wgrad taps=%d Cin=%d Cout=%d %dx%d MO=%d NC=%d split=%d stages=%d%s", ntaps, Cin, Cout, x.H, x.W, p.MO,
              p.NC, p.nsplit, p.stages, db ? " +bias" : "");
        Bk([p](cudaStream_t st) { igemm_wgrad_launch(p, st); }, 1, UB_KIND_WGRAD,
           2.0 * B * x.H * x.W * double(Cout) * Cin * ntaps,
           act_bytes(Cin, x.H, x.W) + act_bytes(Cout, x.H, x.W) + 4.0 * ntaps * Cin * Cout, 1);
    }
    void fin_flush() {
        if (!real()) return;
        const int lo = fin_done, n = int(T->h_fin.size()) - fin_done;
        if (n <= 0) return;
        size_t mx = 0;
        for (int i = lo; i < lo + n; ++i) mx = std::max(mx, size_t(T->h_fin[i].Cout) * T->h_fin[i].Cin);
        UbTrainer* Tt = T;
        label("wgrad finalize x%d", n);
        Bk([=](cudaStream_t st) { igemm_wgrad_finalize(Tt->fin_table + lo, n, mx, st); }, 1, UB_KIND_WGRAD, 0,
           0, 1);
        fin_done = lo + n;
    }

    struct GN {
        size_t w, b;
        float *chsum, *S;
    };
    // GroupNorm finish (IgemmConvParams::gnf): convs of the low-resolution levels whose epilogue can complete the
    // GroupNorm that consumes their output (forward) / whose gn-bwd hook they carry (backward), keyed by output tensor.
    // gn_fwd / gn_bwd extend the plan and emit no kernel of their own when the output's producer is registered here.
    std::map<const void*, std::shared_ptr<IgemmConvParams>> gn_finish;
    int gn_finished_fwd = 0, gn_finished_bwd = 0;
    // statistics buffer of a tensor whose producer (a conv epilogue) accumulates them
    // GroupNorm scheme (UB_GN_MODE), decided per tensor shape.  Measured round 2 (bench.py, B = 32, ms/step):
    //   hooks 5.09 (default) | auto 5.23 | passes 5.47 | slab 6.05
    //   hooks  statistics / gn-bwd pre-pass accumulated with atomics in the producing conv's epilogue + one apply pass
    //          (costs ~0.5 ms per step inside the conv epilogues, profiles/r02d_ops_nohooks.txt, but every alternative
    //          costs more on the GroupNorm side: the step is bound by kernel count x per-kernel latency, not bytes)
    //   auto   single-pass slab kernels (csrc/gn_slab.cu: statistics + normalisation in one kernel, nothing in the conv
    //          epilogues) wherever they beat the two-pass kernels stand-alone (gn_slab_preferred: below 64x64), hooks
    //          elsewhere
    //   slab   slab kernels wherever they are supported, `passes` elsewhere
    //   passes separate statistics pass + apply pass, no hooks
    static int gn_mode() {
        static const int m = [] {
            const char* e = getenv("UB_GN_MODE");
            if (e && e[0] == 's') return 0;
            if (e && e[0] == 'p') return 2;
            if (e && e[0] == 'a') return 3;
            return 1;
        }();
        return m;
    }
    bool use_slab(int C, int H, int W) const {
        const int m = gn_mode();
        if (m == 0) return gn_slab_supported(Bh, H * W, C, c.gn_n_groups, false) && gn_slab_supported(Bh, H * W, C, c.gn_n_groups, true);
        return m == 3 && gn_slab_preferred(Bh, H * W, C, c.gn_n_groups);
    }
    // does the conv that produces a (C, H, W) tensor accumulate its GroupNorm statistics / run the gn-bwd pre-pass?
    bool use_hooks(int C, int H, int W) const {
        const int m = gn_mode();
        return m == 1 || (m == 3 && !use_slab(C, H, W));
    }
    // The gn-bwd pre-pass hook (x read + silu' + two column sums in the dgrad epilogue) makes the 64x64 dgrad convs
    // epilogue-bound (39 us against 24 us unhooked): above UB_GN_BWD_HOOK_MAX_HW pixels per image the backward uses the
    // separate statistics pass instead (the forward statistics hook stays).
    static int bwd_hook_max_hw() {
        static const int v = getenv("UB_GN_BWD_HOOK_MAX_HW") ? atoi(getenv("UB_GN_BWD_HOOK_MAX_HW")) : (1 << 30);
        return v;
    }
    bool use_bwd_hook(int C, int H, int W) const { return use_hooks(C, H, W) && H * W <= bwd_hook_max_hw(); }
    float* stats_buf(int C, int H, int W) { return use_hooks(C, H, W) ? zf32(size_t(B) * C * 2) : nullptr; }
    // ss (use_scale_shift_norm): per-image [scale | shift] rows applied after the normalisation (gn_scale_shift)
    GN gn_fwd(View x, View y, int silu, const float* ss = nullptr, DropArgs drop = DropArgs()) {
        GN g;
        g.w = take(x.C), g.b = take(x.C);
        const bool have = x.cs != nullptr;  // the producer's epilogue already accumulated the statistics
        g.chsum = have ? x.cs : zf32(size_t(B) * x.C * 2);
        g.S = zf32(size_t(B) * x.C * 2);
        const int HW = x.H * x.W, Gn = c.gn_n_groups;
        float *gw = P(g.w), *gb = P(g.b), *cs = g.chsum;
        const bool slab = !have && !ss && !drop.ctl && use_slab(x.C, x.H, x.W);
        if (have && real() && !ss && !drop.ctl) {
            auto it = gn_finish.find(x.p);
            if (it != gn_finish.end()) {
                auto pp = it->second;
                gn_finish.erase(it);
                if (x.ld == pp->ldo && x.C == pp->Cout && pp->stats == cs &&
                    igemm_conv_gn_finish_fwd(pp.get(), gw, gb, Gn, silu, y.p, y.ld)) {
                    gn_finished_fwd++;
                    return g;  // the producing conv's epilogue writes y
                }
            }
        }
        label("gn_fwd C=%d %dx%d%s", x.C, x.H, x.W, slab ? " slab" : (have ? "" : " +stats pass"));
        split([&](int h) {
            const View xh = mb(x, h), yh = mb(y, h);
            float* csh = mbf(cs, h, size_t(x.C) * 2);
            const int Bq = Bh;
            if (slab)
                F([=](cudaStream_t st) { gn_slab_fwd(xh.p, xh.ld, gw, gb, Bq, HW, xh.C, Gn, silu, yh.p, yh.ld, csh, st); }, 1,
                  UB_KIND_NORM, 0, 2 * act_bytes(x.C, x.H, x.W) / nh);
            else
                F([=](cudaStream_t st) {
                    if (!have) gn_stats(xh.p, xh.ld, Bq, HW, xh.C, csh, st);
                    gn_apply(xh.p, xh.ld, csh, gw, gb, Bq, HW, xh.C, Gn, silu, yh.p, yh.ld, nullptr, st,
                             mbf(const_cast<float*>(ss), h, size_t(2) * xh.C), drop);
                }, have ? 1 : 2, UB_KIND_NORM, 0, (have ? 2 : 3) * act_bytes(x.C, x.H, x.W) / nh);
        });
        return g;
    }
    // Fuse the first half of GroupNorm(+SiLU) backward into the epilogue of the dgrad conv that produces dL/dy:
    // that conv then writes dz = dL/d gn(x) and accumulates S (see epilogue.cuh); call gn_bwd(..., fused = true) after.
    void gn_hook(ConvEpilogue& ep, const GN& g, View x, int silu) {
        if (!use_bwd_hook(x.C, x.H, x.W)) return;
        ep.gn_x = x.p, ep.gn_ldx = x.ld, ep.gn_chsum = g.chsum, ep.gn_gamma = P(g.w), ep.gn_beta = P(g.b);
        ep.gn_S = g.S, ep.gn_silu = silu, ep.gn_groups = c.gn_n_groups;
    }
    void gn_bwd(const GN& g, View x, View dy, int silu, View add_in, View dx, float* colsum_out, bool fused = false,
                const float* ss = nullptr, float* dss = nullptr, DropArgs drop = DropArgs()) {
        fused = fused && use_bwd_hook(x.C, x.H, x.W) && !ss && !drop.ctl;
        const int HW = x.H * x.W, Gn = c.gn_n_groups;
        float *gw = P(g.w), *gb = P(g.b), *cs = g.chsum, *S = g.S, *dgw = G(g.w), *dgb = G(g.b);
        const bool slab = !fused && !ss && !drop.ctl && use_slab(x.C, x.H, x.W);
        const int mode = fused ? (silu ? 2 : 0) : silu;
        if (fused && real()) {
            auto it = gn_finish.find(dy.p);
            if (it != gn_finish.end()) {
                auto pp = it->second;
                gn_finish.erase(it);
                if (dy.ld == pp->ldo && x.C == pp->Cout && pp->gn_S == S &&
                    igemm_conv_gn_finish_bwd(pp.get(), add_in.p, add_in.ld, dx.p, dx.ld, dgw, dgb, colsum_out)) {
                    gn_finished_bwd++;
                    return;  // the dgrad conv's epilogue writes dx, dgamma, dbeta and the column sums
                }
            }
        }
        label("gn_bwd C=%d %dx%d%s%s", x.C, x.H, x.W, slab ? " slab" : (fused ? "" : " +stats pass"), add_in.p ? " +add" : "");
        split([&](int h) {
            const View xh = mb(x, h), dyh = mb(dy, h), ah = mb(add_in, h), dxh = mb(dx, h);
            float *csh = mbf(cs, h, size_t(x.C) * 2), *Sh = mbf(S, h, size_t(x.C) * 2), *colh = mbf(colsum_out, h, x.C);
            const int Bq = Bh;
            if (slab)
                Bk([=](cudaStream_t st) {
                    gn_slab_bwd(xh.p, xh.ld, dyh.p, dyh.ld, csh, gw, gb, Bq, HW, xh.C, Gn, silu, ah.p, ah.ld, dxh.p, dxh.ld,
                                dgw, dgb, colh, st);
                }, 1, UB_KIND_NORM, 0, (3 + (add_in.p ? 1 : 0)) * act_bytes(x.C, x.H, x.W) / nh);
            else
                Bk([=](cudaStream_t st) {
                    const float* ssh = mbf(const_cast<float*>(ss), h, size_t(2) * xh.C);
                    float* dssh = mbf(dss, h, size_t(2) * xh.C);
                    if (!fused) gn_bwd_stats(xh.p, xh.ld, dyh.p, dyh.ld, csh, gw, gb, Bq, HW, xh.C, Gn, silu, Sh, st, ssh, drop);
                    gn_bwd_apply(xh.p, xh.ld, dyh.p, dyh.ld, csh, Sh, gw, gb, Bq, HW, xh.C, Gn, mode, ah.p, ah.ld, dxh.p,
                                 dxh.ld, dgw, dgb, colh, st, ssh, dssh, drop);
                }, fused ? 1 : 2, UB_KIND_NORM, 0, ((fused ? 3 : 5) + (add_in.p ? 1 : 0)) * act_bytes(x.C, x.H, x.W) / nh);
        });
    }

    int res_index = 0;
    // The 22 embedding-projection backward GEMMs (N = B rows each) are latency-bound: they are launched in batches,
    // one batch per gradient bucket (their weight gradients must be final before the bucket is all-reduced).
    int emb_lo = 1 << 30, emb_hi = -1;
    void emb_flush() {
        if (emb_hi <= emb_lo) return;
        UbTrainer* Tt = T;
        const int lo = emb_lo, n = emb_hi - emb_lo, Bn = B, Cemb = 4 * c.C_model;
        Bk([=](cudaStream_t st) { small_linear_bwd(Tt->emb_table + lo, n, Bn, Tt->emb_max_oc, Cemb, st); }, 2,
           UB_KIND_SMALL, 0, 0, 1);
        emb_lo = 1 << 30, emb_hi = -1;
    }

    // Time-MLP backward (needs the complete d_embact, i.e. every embedding projection's backward).  Emitted on the
    // weight-gradient branch as soon as the first ResBlock has produced its embedding gradient, so that it runs under
    // the last dgrad convs instead of after them (it used to be 80 us of the step's serial tail).
    // Step tail (UB_TAIL_AUX=0 restores the single-branch order): the time-MLP backward (six small dependent kernels)
    // forks off the weight-gradient branch into a branch of its own, and the 3-channel input conv's weight gradient --
    // the last gradient of the step -- runs on the main stream, which has nothing left to do: the three chains
    // (last wgrad -> finalize -> bucket optimizer | time MLP | input-conv wgrad) overlap instead of queueing.
    static bool tail_aux() {
        static const bool on = !(getenv("UB_TAIL_AUX") && atoi(getenv("UB_TAIL_AUX")) == 0);
        return on;
    }
    bool time_mlp_bwd_done = false;
    void time_mlp_bwd() {
        if (time_mlp_bwd_done) return;
        time_mlp_bwd_done = true;
        UbTrainer* Tt = T;
        const int Bn = B, Cm = c.C_model, Cemb = 4 * c.C_model;
        const bool cls = c.num_classes > 0;
        Bk([=](cudaStream_t st) {
            dsilu_mul(Tt->d_embact, Tt->emb, Tt->demb, size_t(Bn) * Cemb, st);
            if (cls) label_emb_bwd(Tt->demb, Tt->labels, Bn, Cemb, Tt->grads + Tt->label_w_off, st);
            small_linear_bwd(Tt->temb_table + 1, 1, Bn, Cemb, Cemb, st);
            dsilu_mul(Tt->d_h0act, Tt->h0, Tt->dh0, size_t(Bn) * Cemb, st);
            small_linear_bwd(Tt->temb_table, 1, Bn, Cemb, Cm, st);
        }, cls ? 7 : 6, UB_KIND_SMALL, 0, 0, tail_aux() ? 3 : 1);  // 3: a branch of its own, forked from the weight-gradient branch
    }

    // 2x resampling of a tensor inside an up/down ResBlock (h_upd / x_upd, dev/resblock.py:78-86,125-128): ud = 1 average
    // pool, ud = 2 nearest upsample
    void resample_fwd(View x, View y, int ud) {
        split([&](int hh) {
            const View xm = mb(x, hh), ym = mb(y, hh);
            const int Bq = Bh;
            F([=](cudaStream_t st) {
                if (ud == 1)
                    avgpool2_fwd(xm.p, xm.ld, Bq, xm.H, xm.W, xm.C, ym.p, ym.ld, st);
                else
                    upsample2_fwd(xm.p, xm.ld, Bq, xm.H, xm.W, xm.C, ym.p, ym.ld, st);
            }, 1, UB_KIND_ELTWISE, 0, 1.25 * act_bytes(x.C, std::max(x.H, y.H), std::max(x.W, y.W)) / nh);
        });
    }
    // gradient of resample_fwd: dx (resolution of x) from dy (resolution of y)
    void resample_bwd(View dy, View dx, int ud) {
        split([&](int hh) {
            const View dm = mb(dy, hh), dxm = mb(dx, hh);
            const int Bq = Bh;
            Bk([=](cudaStream_t st) {
                if (ud == 1)
                    avgpool2_bwd(dm.p, dm.ld, Bq, dxm.H, dxm.W, dxm.C, nullptr, 0, dxm.p, dxm.ld, st);
                else
                    upsample2_bwd(dm.p, dm.ld, Bq, dm.H, dm.W, dm.C, dxm.p, dxm.ld, st);
            }, 1, UB_KIND_ELTWISE, 0, 1.25 * act_bytes(dx.C, std::max(dx.H, dy.H), std::max(dx.W, dy.W)) / nh);
        });
    }

    // ResBlock (dev/resblock.py:107-160, train_unet.cu:2213-2384).  ud != 0 (cfg.resblock_updown, dev/unet.py:205-222,
    // 271-284): the block resamples -- 1 = down (average pool), 2 = up (nearest) -- both its main path, between SiLU and
    // conv1, and its skip path (h_upd / x_upd, dev/resblock.py:125-128); x is at the input resolution, everything from
    // conv1 on at the output resolution, and Cout == C.
    View resblock(View x, int Cout, int ud = 0) {
        Node nd;
        nd.param_begin = poff;
        const int C = x.C, Hin = x.H, Win = x.W, Cemb = 4 * c.C_model;
        const int H = ud == 1 ? Hin / 2 : (ud == 2 ? 2 * Hin : Hin), W = ud == 1 ? Win / 2 : (ud == 2 ? 2 * Win : Win);
        const int blk = res_index++;
        View a1 = act(C, Hin, Win);
        GN g1 = gn_fwd(x, a1, 1);
        const View xin = x;  // at the input resolution
        if (ud) {
            View a1r = act(C, H, W), xr = act(C, H, W);
            resample_fwd(a1, a1r, ud);
            resample_fwd(x, xr, ud);
            a1 = a1r, x = xr;  // what conv1 and the (identity) skip connection see
        }
        // use_scale_shift_norm (cfg.use_scale_shift_norm; dev/resblock.py:211,243-247 ResBlockO): the embedding projection
        // has 2 * Cout outputs [scale | shift], nothing is added to conv1's output and GroupNorm 2 becomes
        // gn(h) * (1 + scale) + shift -- a per-image affine inside the GroupNorm kernels (gn_scale_shift); its backward
        // runs unfused (plain dgrad of conv2, statistics pass, apply pass that also writes dscale / dshift)
        const bool ssn = c.use_scale_shift_norm != 0;
        const int OCe = ssn ? 2 * Cout : Cout;
        // dropout (cfg.dropout; guided-diffusion out_layers = GroupNorm, SiLU, Dropout, conv): the mask multiplies
        // silu(gn2(h)) inside the GroupNorm apply kernel and is regenerated by the unfused GroupNorm backward (DropArgs)
        DropArgs drp;
        if (c.dropout > 0.f) {
            drp.ctl = T->drop_ctl, drp.layer = unsigned(blk);
            if (real()) T->drop_layers.push_back({Cout, H, W});
        }
        const bool unfused2 = ssn || c.dropout > 0.f;  // GroupNorm 2 backward without the dgrad hook
        const size_t w1 = take(size_t(Cout) * C * 9), b1 = take(Cout);
        const size_t wl = take(size_t(OCe) * Cemb), bl = take(OCe);
        View h1 = act(Cout, H, W);
        h1.cs = stats_buf(Cout, H, W);
        float* embproj = f32(size_t(B) * OCe);
        float* d_embproj = zf32(size_t(B) * OCe);
        Packed p1 = pack(w1, Cout, C, 9);
        {
            ConvEpilogue ep;
            ep.bias = P(b1), ep.rowvec = ssn ? nullptr : embproj, ep.out = h1.p, ep.ldo = h1.ld, ep.stats = h1.cs;
            join_side_fwd();  // embproj comes from the time-embedding chain
            conv_op(true, {{a1.p, C, a1.ld, p1.wf, 9}}, H, W, Cout, ep);
        }
        View a2 = act(Cout, H, W);
        GN g2 = gn_fwd(h1, a2, 1, ssn ? embproj : nullptr, drp);
        const size_t w2 = take(size_t(Cout) * Cout * 9), b2 = take(Cout);
        Packed p2 = pack(w2, Cout, Cout, 9);
        const bool proj = C != Cout;
        size_t ws = 0, bs = 0;
        Packed ps;
        if (proj) {
            ws = take(size_t(Cout) * C), bs = take(Cout);
            ps = pack(ws, Cout, C, 1);
        }
        View out = act(Cout, H, W);
        out.cs = stats_buf(Cout, H, W);
        {
            ConvEpilogue ep;
            ep.bias = P(b2), ep.out = out.p, ep.ldo = out.ld, ep.stats = out.cs;
            std::vector<ConvSegDesc> segs = {{a2.p, Cout, a2.ld, p2.wf, 9}};
            if (proj) {
                segs.push_back({x.p, C, x.ld, ps.wf, 1});
                ep.bias2 = P(bs);
            } else {
                ep.residual = x.p, ep.ldr = x.ld;
            }
            conv_op(true, segs, H, W, Cout, ep);
        }
        if (real()) {
            SmallLinear e{};
            e.w = P(wl), e.b = P(bl), e.inp = T->semb, e.out = embproj, e.C = Cemb, e.OC = OCe, e.silu_in = 0;
            // (without scale-shift, conv1's bias gradient equals the embedding projection's: h1 = conv1 + b1 + emb)
            e.dout = d_embproj, e.dw = G(wl), e.db = G(bl), e.db2 = ssn ? nullptr : G(b1), e.dinp = T->d_embact;
            T->h_emb.push_back(e);
            if (OCe > T->emb_max_oc) T->emb_max_oc = OCe;
        }
        nd.out = out;
        nd.bwd = [=](View dout) -> View {
            View da2 = act(Cout, H, W), dh1 = act(Cout, H, W), da1 = act(C, H, W), dxg = act(C, H, W);
            View dxo = proj ? act(C, H, W) : dxg;
            const size_t npix = size_t(B) * H * W;
            float *gb2 = G(b2), *gbs = proj ? G(bs) : nullptr;
            UbTrainer* Tt = T;
            const int Bn = B;
            // bias / weight gradients of conv2 (and of the fused 1x1 skip conv)
            wgrad_op(dout, a2, Cout, Cout, 9, G(w2), gb2, gbs);  // + bias gradients of conv2 and of the skip conv
            if (proj) wgrad_op(dout, x, C, Cout, 1, G(ws));
            {
                ConvEpilogue ep;
                ep.out = da2.p, ep.ldo = da2.ld;
                if (!unfused2) gn_hook(ep, g2, h1, 1);  // da2 receives dz = dL/d gn2(h1)
                conv_op(false, {{dout.p, Cout, dout.ld, p2.wd, 9}}, H, W, Cout, ep);
            }
            // GN2 + SiLU backward; per-image column sums of dh1 feed the embedding-projection backward
            // (scale-shift: the apply pass writes [dscale | dshift] there instead)
            if (ssn)
                gn_bwd(g2, h1, da2, 1, View{}, dh1, nullptr, false, embproj, d_embproj, drp);
            else
                gn_bwd(g2, h1, da2, 1, View{}, dh1, d_embproj, !unfused2, nullptr, nullptr, drp);
            emb_lo = blk < emb_lo ? blk : emb_lo, emb_hi = blk + 1 > emb_hi ? blk + 1 : emb_hi;  // batched, see emb_flush
            if (blk == 0) {  // the last embedding gradient of the step: finish the embedding path on the branch now
                emb_flush();
                time_mlp_bwd();
            }
            wgrad_op(dh1, a1, C, Cout, 9, G(w1), ssn ? G(b1) : nullptr);  // (+ conv1's bias gradient when it is its own)
            if (ud) {
                // conv1 ran on the resampled tensor: its dgrad is plain (no gn-bwd hook -- GroupNorm 1 lives at the input
                // resolution), both gradients go back through the resampling, GroupNorm 1 + SiLU backward run unfused
                ConvEpilogue ep;
                ep.out = da1.p, ep.ldo = da1.ld;
                conv_op(false, {{dh1.p, Cout, dh1.ld, p1.wd, 9}}, H, W, C, ep);
                View da1in = act(C, Hin, Win), dskip = act(C, Hin, Win), dxin = act(C, Hin, Win);
                resample_bwd(da1, da1in, ud);
                resample_bwd(dout, dskip, ud);  // identity skip connection behind x_upd (Cout == C)
                gn_bwd(g1, xin, da1in, 1, dskip, dxin, nullptr, false);
                return dxin;
            }
            {
                ConvEpilogue ep;
                ep.out = da1.p, ep.ldo = da1.ld;
                gn_hook(ep, g1, x, 1);
                conv_op(false, {{dh1.p, Cout, dh1.ld, p1.wd, 9}}, H, W, C, ep);
            }
            gn_bwd(g1, x, da1, 1, proj ? View{} : dout, dxg, nullptr, true);
            if (proj) {
                ConvEpilogue ep;
                ep.out = dxo.p, ep.ldo = dxo.ld, ep.residual = dxg.p, ep.ldr = dxg.ld;
                conv_op(false, {{dout.p, Cout, dout.ld, ps.wd, 1}}, H, W, C, ep);
            }
            return dxo;
        };
        nodes.push_back(nd);
        return out;
    }

    // AttentionBlock (dev/unet.py:44-58, train_unet.cu:2933-2976)
    View attnblock(View x) {
        Node nd;
        nd.param_begin = poff;
        const int C = x.C, H = x.H, W = x.W, Tn = H * W, NH = C / c.head_size, HSz = c.head_size;
        View g = act(C, H, W);
        GN gn = gn_fwd(x, g, 0);
        const size_t wq = take(size_t(3) * C * C), bq = take(size_t(3) * C);
        const size_t wp = take(size_t(C) * C), bp = take(C);
        Packed pq = pack(wq, 3 * C, C, 1), pp = pack(wp, C, C, 1);
        View qkv = act(3 * C, H, W), ao = act(C, H, W), out = act(C, H, W);
        out.cs = stats_buf(C, H, W);
        float* lse = f32(size_t(B) * NH * Tn);
        float* dsum = f32(size_t(B) * NH * Tn);
        {
            ConvEpilogue ep;
            ep.bias = P(bq), ep.out = qkv.p, ep.ldo = qkv.ld;
            conv_op(true, {{g.p, C, g.ld, pq.wf, 1}}, H, W, 3 * C, ep);
        }
        // tcgen05 attention core when the shape allows (T <= 256, even head count), SIMT fallback otherwise
        const bool tc = attn_tc_supported(Tn, NH, HSz);
        label("attn fwd T=%d C=%d", Tn, C);
        split([&](int h) {
            const View qh = mb(qkv, h), aoh = mb(ao, h);
            float* lseh = mbf(lse, h, size_t(NH) * Tn);
            const int Bq = Bh;
            AttnTcParams apf;
            if (tc && real()) {
                int r = attn_tc_plan(&apf, qh.p, qh.ld, Bq, Tn, NH, HSz, aoh.p, aoh.ld, lseh, nullptr, 0, nullptr, 0, nullptr);
                if (r) set_err("attn_tc_plan failed (%d)", r), plan_errors++;
            }
            F([=](cudaStream_t st) {
                if (tc)
                    attn_tc_fwd(apf, st);
                else
                    attn_fwd(qh.p, qh.ld, Bq, Tn, NH, HSz, aoh.p, aoh.ld, lseh, st);
            }, 1, UB_KIND_ATTN, 4.0 * Bq * NH * double(Tn) * Tn * HSz, act_bytes(4 * C, H, W) / nh);
        });
        {
            ConvEpilogue ep;
            ep.bias = P(bp), ep.out = out.p, ep.ldo = out.ld, ep.residual = x.p, ep.ldr = x.ld, ep.stats = out.cs;
            conv_op(true, {{ao.p, C, ao.ld, pp.wf, 1}}, H, W, C, ep);
        }
        nd.out = out;
        nd.bwd = [=](View dout) -> View {
            View dao = act(C, H, W), dqkv = act(3 * C, H, W), dg = act(C, H, W), dx = act(C, H, W);
            const size_t npix = size_t(B) * H * W;
            float *gbp = G(bp), *gbq = G(bq);
            wgrad_op(dout, ao, C, C, 1, G(wp), gbp);
            {
                ConvEpilogue ep;
                ep.out = dao.p, ep.ldo = dao.ld;
                conv_op(false, {{dout.p, C, dout.ld, pp.wd, 1}}, H, W, C, ep);
            }
            label("attn bwd T=%d C=%d", Tn, C);
            UbTrainer* Tt = T;
            split([&](int h) {
                const View qh = mb(qkv, h), aoh = mb(ao, h), daoh = mb(dao, h), dqh = mb(dqkv, h);
                float *lseh = mbf(lse, h, size_t(NH) * Tn), *dsh = mbf(dsum, h, size_t(NH) * Tn);
                const int Bq = Bh;
                AttnTcParams apb;
                if (tc && real()) {
                    int r = attn_tc_plan(&apb, qh.p, qh.ld, Bq, Tn, NH, HSz, aoh.p, aoh.ld, lseh, daoh.p, daoh.ld, dqh.p,
                                         dqh.ld, dsh);
                    if (r) set_err("attn_tc_plan (bwd) failed (%d)", r), plan_errors++;
                }
                Bk([=](cudaStream_t st) {
                    if (tc)
                        attn_tc_bwd(apb, st, Tt->attn_concurrent ? Tt->aux_stream : nullptr, Tt->ev_aux_fork, Tt->ev_aux_join);
                    else
                        attn_bwd(qh.p, qh.ld, aoh.p, aoh.ld, daoh.p, daoh.ld, lseh, Bq, Tn, NH, HSz, dqh.p, dqh.ld, dsh, st);
                }, 2, UB_KIND_ATTN, 10.0 * Bq * NH * double(Tn) * Tn * HSz, act_bytes(8 * C, H, W) / nh);
            });
            wgrad_op(dqkv, g, C, 3 * C, 1, G(wq), gbq);
            {
                ConvEpilogue ep;
                ep.out = dg.p, ep.ldo = dg.ld;
                gn_hook(ep, gn, x, 0);
                conv_op(false, {{dqkv.p, 3 * C, dqkv.ld, pq.wd, 1}}, H, W, C, ep);
            }
            gn_bwd(gn, x, dg, 0, dout, dx, nullptr, true);
            return dx;
        };
        nodes.push_back(nd);
        return out;
    }

    int build();
};

int Builder::build() {
    const int Cm = c.C_model, Cemb = 4 * Cm, H0 = c.H, W0 = c.W;
    const size_t img = size_t(c.C_in) * H0 * W0;
    UbTrainer* Tt = T;
    const int Bn = B;
    // io + small fp32 state
    T->x0 = f32(size_t(B) * img), T->xt = f32(size_t(B) * img), T->noise = f32(size_t(B) * img);
    T->tsteps = f32(B);
    T->labels = c.num_classes > 0 ? (int*)T->arena.alloc(size_t(B) * sizeof(int) + 256) : nullptr;
    T->drop_ctl = c.dropout > 0.f ? (unsigned*)T->arena.alloc(256) : nullptr;
    T->dxt = c.compute_dinput ? f32(size_t(B) * img) : nullptr;
    T->flips = c.random_flip ? (int*)T->arena.alloc(size_t(B) * sizeof(int) + 256) : nullptr;
    T->out = f32(size_t(B) * c.C_out * H0 * W0), T->dout = f32(size_t(B) * c.C_out * H0 * W0);
    T->sqrt_ac = f32(c.n_timesteps), T->sqrt_1mac = f32(c.n_timesteps), T->betas = f32(c.n_timesteps);
    T->samp_state = (int*)T->arena.alloc(256);
    T->samp_z = f32(size_t(B) * img);
    T->step_dev = (int*)T->arena.alloc(256);
    T->hp_dev = f32(8);
    T->loss = zf32(64);
    T->sin_emb = f32(size_t(B) * Cm), T->h0 = f32(size_t(B) * Cemb), T->emb = f32(size_t(B) * Cemb);
    T->semb = f32(size_t(B) * Cemb);
    T->d_embact = zf32(size_t(B) * Cemb), T->demb = f32(size_t(B) * Cemb);
    T->d_h0act = zf32(size_t(B) * Cemb), T->dh0 = f32(size_t(B) * Cemb);
    T->small_scratch = f32(T->small_scratch_floats);
    T->emb_table = (SmallLinear*)T->arena.alloc(kMaxEmbEntries * sizeof(SmallLinear));
    T->temb_table = (SmallLinear*)T->arena.alloc(2 * sizeof(SmallLinear));
    T->pack_table = (PackEntry*)T->arena.alloc(256 * sizeof(PackEntry));
    T->fin_table = (WgradFinalizeEntry*)T->arena.alloc(128 * sizeof(WgradFinalizeEntry));

    // ---- time-embedding MLP (dev/unet.py:176-180); its forward and all 22 embedding projections run first
    const size_t tw0 = take(size_t(Cemb) * Cm), tb0 = take(Cemb), tw1 = take(size_t(Cemb) * Cemb), tb1 = take(Cemb);
    if (real()) {
        SmallLinear e0{}, e1{};
        e0.w = P(tw0), e0.b = P(tb0), e0.inp = T->sin_emb, e0.out = T->h0, e0.C = Cm, e0.OC = Cemb, e0.silu_in = 0;
        e0.dout = T->dh0, e0.dw = G(tw0), e0.db = G(tb0);
        e1.w = P(tw1), e1.b = P(tb1), e1.inp = T->h0, e1.out = T->emb, e1.C = Cemb, e1.OC = Cemb, e1.silu_in = 1;
        e1.dout = T->demb, e1.dw = G(tw1), e1.db = G(tb1), e1.dinp = T->d_h0act;
        T->h_temb = {e0, e1};
    }
    // class-conditional model: label_emb.weight (num_classes, 4*C_model) follows the time MLP in named_parameters()
    // order (dev/unet.py:174-175); emb = time_embed(t) + label_emb[y] (dev/unet.py:301-303)
    const size_t lw = c.num_classes > 0 ? take(size_t(c.num_classes) * Cemb) : 0;
    if (real()) T->label_w_off = lw;
    {
        const int mp = c.max_period;
        const bool cls = c.num_classes > 0;
        F([=](cudaStream_t st) {
            const SmallLinear &e0 = Tt->h_temb[0], &e1 = Tt->h_temb[1];
            time_mlp_fwd(Tt->tsteps, Bn, Cm, Cemb, mp, e0.w, e0.b, e1.w, e1.b, Tt->sin_emb, Tt->h0, Tt->emb, Tt->semb,
                         st, cls ? Tt->params + Tt->label_w_off : nullptr,
                         Tt->labels);  // (semb = silu(emb) is the shared input of all embedding projections)
            small_linear_fwd(Tt->emb_table, int(Tt->h_emb.size()), Bn, Tt->emb_max_oc, st, true);  // (Cout % 16 == 0)
        }, 2, UB_KIND_SMALL, 0, 0, 1);  // beside the input conv and the first GroupNorm
        fwd_side_pending = true;
    }
    const size_t time_mlp_end = poff;

    // ---- input conv (dev/unet.py:183-185)
    int ch = c.channel_mult[0] * Cm;
    View h = act(ch, H0, W0);
    {
        Node nd;
        nd.param_begin = poff;
        const size_t wi = take(size_t(ch) * c.C_in * 9), bi = take(ch);
        float *w = P(wi), *b = P(bi), *gw = G(wi), *gb = G(bi);
        const int Cin = c.C_in;
        View hv = h;
        split([&](int hh) {
            const View hm = mb(hv, hh);
            const size_t xoff = size_t(hh) * Bh * img;
            const int Bq = Bh;
            F([=](cudaStream_t st) { conv_in_fwd(Tt->xt + xoff, w, b, Bq, Cin, hm.C, hm.H, hm.W, hm.p, hm.ld, st); });
        });
        nd.out = h, nd.pushed = true;
        nd.bwd = [=](View dout) -> View {
            Bk([=](cudaStream_t st) {
                conv_in_wgrad(Tt->xt, dout.p, dout.ld, Bn, Cin, hv.C, hv.H, hv.W, gw, gb, Tt->small_scratch,
                              Tt->small_scratch_floats, st);
            }, 1, UB_KIND_SMALL, 0, 0, tail_aux() ? 0 : 1);  // (one launch with the row kernel; two on its fallback path)
            if (Tt->cfg.compute_dinput)
                split([&](int hh) {
                    const View dm = mb(dout, hh);
                    const size_t xoff = size_t(hh) * Bh * img;
                    const int Bq = Bh;
                    Bk([=](cudaStream_t st) { conv_in_dgrad(dm.p, dm.ld, w, Bq, Cin, hv.C, hv.H, hv.W, Tt->dxt + xoff, st); }, 1);
                });
            return View{};
        };
        nodes.push_back(nd);
    }
    std::vector<int> skip_stack = {0};  // node ids whose outputs are on the skip stack
    std::vector<View> skip_views = {h};

    const int nlev = c.n_levels;
    level_mb(H0, W0);
    sync_nh();
    for (int level = 0; level < nlev; ++level) {
        const int cout = c.channel_mult[level] * Cm;
        level_mb(H0 >> level, W0 >> level);
        for (int i = 0; i < c.n_res_blocks; ++i) {
            h = resblock(h, cout);
            if (level >= c.att_start_level) h = attnblock(h);
            nodes.back().pushed = true;
            skip_stack.push_back(int(nodes.size()) - 1);
            skip_views.push_back(h);
        }
        if (level != nlev - 1 && c.resblock_updown) {  // ResBlock(down=True) instead of Downsample (dev/unet.py:205-222)
            h = resblock(h, h.C, 1);
            nodes.back().pushed = true;
            skip_stack.push_back(int(nodes.size()) - 1);
            skip_views.push_back(h);
        } else if (level != nlev - 1) {  // Downsample (dev/resblock.py:34-43)
            Node nd;
            nd.param_begin = poff;
            View x = h, y = act(h.C, h.H / 2, h.W / 2);
            split([&](int hh) {
                const View xm = mb(x, hh), ym = mb(y, hh);
                const int Bq = Bh;
                F([=](cudaStream_t st) { avgpool2_fwd(xm.p, xm.ld, Bq, xm.H, xm.W, xm.C, ym.p, ym.ld, st); }, 1,
                  UB_KIND_ELTWISE, 0, 1.25 * act_bytes(x.C, x.H, x.W) / nh);
            });
            nd.out = y, nd.pushed = true;
            nd.bwd = [=](View dout) -> View {
                View dx = act(x.C, x.H, x.W);
                split([&](int hh) {
                    const View dm = mb(dout, hh), dxm = mb(dx, hh);
                    const int Bq = Bh;
                    Bk([=](cudaStream_t st) {
                        avgpool2_bwd(dm.p, dm.ld, Bq, x.H, x.W, x.C, nullptr, 0, dxm.p, dxm.ld, st);
                    }, 1, UB_KIND_ELTWISE, 0, 1.25 * act_bytes(x.C, x.H, x.W) / nh);
                });
                return dx;
            };
            nodes.push_back(nd);
            h = y;
            skip_stack.push_back(int(nodes.size()) - 1);
            skip_views.push_back(h);
        }
    }
    sync_nh();
    // ---- middle (dev/unet.py:224-243)
    h = resblock(h, h.C);
    h = attnblock(h);
    h = resblock(h, h.C);
    // ---- up path (dev/unet.py:246-284)
    bool pending_up = false;
    for (int level = nlev - 1; level >= 0; --level) {
        const int cout = c.channel_mult[level] * Cm;
        sync_nh();
        level_mb(H0 >> level, W0 >> level);
        for (int i = 0; i <= c.n_res_blocks; ++i) {
            const int src = skip_stack.back();
            View sk = skip_views.back();
            skip_stack.pop_back(), skip_views.pop_back();
            // concat (+ fused nearest x2 upsample of the main path, dev/resblock.py:25-32)
            Node nd;
            nd.param_begin = poff;
            const int C1 = h.C, C2 = sk.C, Hc = sk.H, Wc = sk.W, up = pending_up ? 1 : 0;
            View cat = act(C1 + C2, Hc, Wc), a = h;
            if (a.cs && sk.cs) cat.cs = zf32(size_t(B) * (C1 + C2) * 2);  // GroupNorm statistics of the parts, side by side
            split([&](int hh) {
                const View am = mb(a, hh), sm = mb(sk, hh), cm = mb(cat, hh);
                const int Bq = Bh;
                F([=](cudaStream_t st) {
                    concat2(am.p, am.ld, C1, up, sm.p, sm.ld, C2, Bq, Hc, Wc, cm.p, cm.ld, am.cs, sm.cs, cm.cs, st);
                }, 1, UB_KIND_ELTWISE, 0, 2 * act_bytes(C1 + C2, Hc, Wc) / nh);
            });
            nd.out = cat;
            std::vector<View>* sg = &skipgrad;
            nd.bwd = [=](View d) -> View {
                View skv = d;  // gradient of the skip half is a channel-slice view of d(cat): no copy
                skv.p = d.p ? d.p + C1 : nullptr, skv.C = C2;
                (*sg)[src] = skv;
                if (up) {
                    View dlow = act(C1, Hc / 2, Wc / 2);
                    split([&](int hh) {
                        const View dm = mb(d, hh), lm = mb(dlow, hh);
                        const int Bq = Bh;
                        Bk([=](cudaStream_t st) { upsample2_bwd(dm.p, dm.ld, Bq, Hc, Wc, C1, lm.p, lm.ld, st); }, 1,
                           UB_KIND_ELTWISE, 0, 1.25 * act_bytes(C1, Hc, Wc) / nh);
                    });
                    return dlow;
                }
                View mv = d;
                mv.C = C1;
                return mv;
            };
            nodes.push_back(nd);
            pending_up = false;
            h = resblock(cat, cout);
            if (level >= c.att_start_level) h = attnblock(h);
            if (level && i == c.n_res_blocks) {
                if (c.resblock_updown)
                    h = resblock(h, h.C, 2);  // ResBlock(up=True) instead of Upsample (dev/unet.py:271-284)
                else
                    pending_up = true;
            }
        }
    }
    sync_nh();
    // ---- output head: GN + SiLU + conv3x3 (-> C_out) + MSE (dev/unet.py:286-290, train_unet.cu:4408-4418)
    {
        Node nd;
        nd.param_begin = poff;
        View ao = act(h.C, h.H, h.W), hx = h;
        GN g = gn_fwd(h, ao, 1);
        const size_t wo = take(size_t(c.C_out) * h.C * 9), bo = take(c.C_out);
        float *w = P(wo), *b = P(bo), *gw = G(wo), *gb = G(bo);
        const int Cin = h.C, Co = c.C_out, Hh = h.H, Ww = h.W;
        const size_t N = size_t(B) * c.C_out * h.H * h.W;
        F([=](cudaStream_t st) {
            conv_out_fwd_mse(ao.p, ao.ld, w, b, Bn, Cin, Co, Hh, Ww, Tt->out, Tt->noise, Tt->loss, Tt->dout, 1.f, st);
        }, 1);
        nd.out = View{};
        nd.bwd = [=](View) -> View {
            View dao = act(Cin, Hh, Ww), dh = act(Cin, Hh, Ww);
            Bk([=](cudaStream_t st) {
                conv_out_wgrad(ao.p, ao.ld, Tt->dout, Bn, Cin, Co, Hh, Ww, gw, gb, Tt->small_scratch,
                               Tt->small_scratch_floats, st);
            }, 2, UB_KIND_SMALL, 0, 0, 1);  // (row kernel + bias sum; three on the fallback path)
            split([&](int hh) {
                const View dm = mb(dao, hh);
                const size_t ooff = size_t(hh) * Bh * Co * Hh * Ww;
                const int Bq = Bh;
                Bk([=](cudaStream_t st) { conv_out_dgrad(Tt->dout + ooff, w, Bq, Cin, Co, Hh, Ww, dm.p, dm.ld, st); }, 1);
            });
            gn_bwd(g, hx, dao, 1, View{}, dh, nullptr);
            return dh;
        };
        nodes.push_back(nd);
    }
    if (real() && poff != T->nparams) {
        set_err("parameter count mismatch: built %zu, expected %zu", poff, T->nparams);
        return UB_ERR_SHAPE;
    }
    // ---- backward: walk the nodes in reverse.  Gradient buckets (data parallel): parameters live in forward
    //      order, so once node i's backward has been emitted every parameter at offset >= node[i].param_begin is
    //      final (the time MLP at the head of the arena finishes last).
    // UB_MICROBATCH_FWD=n: micro-batch chains in the FORWARD pass only -- there the device runs one kernel at a time and
    // the 8x8 / 16x16 levels fill less than half of it, while backward already shares the device with the weight-gradient
    // branch.  Backward ops are emitted once for the whole batch (the activations are batch-major views either way).
    sync_nh();
    if (T->mb_fwd_only) nh = 1, Bh = B;
    skipgrad.assign(nodes.size(), View{});
    std::vector<size_t> cuts;  // descending offsets at which a bucket is flushed
    {
        const int nb = T->n_buckets < 1 ? 1 : T->n_buckets;
        for (int k = nb - 1; k >= 1; --k) cuts.push_back(T->nparams * size_t(k) / nb);
        // A last cut right above the input conv: the bucket that must wait for the very end of backward (and whose
        // AdamW + weight re-pack are the serial tail of the step) then holds only the time MLP and the 3-channel conv.
        static const bool tail_cut = !(getenv("UB_TAIL_CUT") && atoi(getenv("UB_TAIL_CUT")) == 0);
        // (Measured round 2: further cuts at the level boundaries of the down path, so that only the first level's
        //  small bucket waits for the end of backward, cost more in extra optimizer / re-pack launches on the tail
        //  than the shorter last all-reduce won: 5.08 vs 5.02 ms on one GPU, 5.13 vs 5.11 on two.)
        // UB_TAIL_NODE = n: the last bucket ends above node n (1 (default) = only the time MLP and the input conv; 2 = the
        // first ResBlock too, so that the last full-size all-reduce is complete one ResBlock earlier).  Measured round 2:
        // 2 costs the single GPU 0.08 ms (a longer serial optimizer tail) and changes nothing on 8 GPUs (5.015 vs 5.006).
        static const int tail_node = getenv("UB_TAIL_NODE") ? atoi(getenv("UB_TAIL_NODE")) : 1;
        const size_t tn = size_t(tail_node < 1 ? 1 : tail_node);
        if (tail_cut && nodes.size() > tn && nodes[tn].param_begin > time_mlp_end && (nodes[tn].param_begin % 4) == 0 &&
            (cuts.empty() || nodes[tn].param_begin < cuts.back()))
            cuts.push_back(nodes[tn].param_begin);
        else if (tail_cut && nodes.size() > 1 && nodes[1].param_begin > time_mlp_end &&
                 (cuts.empty() || nodes[1].param_begin < cuts.back()))
            cuts.push_back(nodes[1].param_begin);
    }
    size_t flushed_hi = T->nparams;
    size_t cut_i = 0;
    int bucket_no = 0;
    auto flush_bucket = [&](size_t lo, size_t hi, bool last) {
        if (hi <= lo) return;
        const int k = bucket_no++;
        static const bool dbg_buckets = getenv("UB_DEBUG_BUCKETS") != nullptr;  // (works in the counting pass too)
        if (dbg_buckets) fprintf(stderr, "[unet_b200] bucket %d: [%zu, %zu)%s\n", k, lo, hi, last ? " last" : "");
        // pack entries whose weights lie in [lo, hi)
        int pk_first = 0, pk_count = 0, pk_tiles = 0;
        if (real() && (lo % 4) != 0) Tt->opt_overlap_ok = false;  // AdamW works on float4: ranges must be aligned
        if (real()) {
            const int n = int(Tt->pack_off.size());
            while (pk_first < n && Tt->pack_off[pk_first] < lo) ++pk_first;
            while (pk_first + pk_count < n && Tt->pack_off[pk_first + pk_count] < hi) {
                const PackEntry& e = Tt->h_pack[pk_first + pk_count];
                pk_tiles = std::max(pk_tiles, ((e.Cout + 31) / 32) * ((e.Cin + 31) / 32));
                ++pk_count;
            }
        }
        Bk([=](cudaStream_t st) {
            const bool dp = Tt->world > 1 && !Tt->comm_off;
            if (dp) {
                cudaEvent_t ev = Tt->bucket_events[k % 16];
                cudaEventRecord(ev, st);
                cudaStreamWaitEvent(Tt->comm_stream, ev, 0);
                nccl().AllReduce(Tt->grads + lo, Tt->grads + lo, hi - lo, kNcclFloat, kNcclSum, Tt->comm,
                                 Tt->comm_stream);
            }
            if (!Tt->opt_in_tape || last) return;  // the last bucket is updated at the end of the step
            // optimizer of this bucket behind its final gradients: the producing branch, or (data parallel) a stream
            // of its own that waits for the bucket's all-reduce
            cudaStream_t os = st;
            if (dp) {
                cudaEvent_t ev2 = Tt->bucket_events[16 + k % 16];
                cudaEventRecord(ev2, Tt->comm_stream);
                cudaStreamWaitEvent(Tt->opt_stream, ev2, 0);
                os = Tt->opt_stream, Tt->opt_stream_dirty = true;
            }
            adamw_step(Tt->params + lo, Tt->grads + lo, Tt->m + lo, Tt->v + lo, hi - lo, Tt->o_lr, Tt->o_b1, Tt->o_b2,
                       Tt->o_eps, Tt->o_wd, 1.f / float(Tt->world), Tt->step_dev, os, Tt->hp_active,
                       Tt->ema ? Tt->ema + lo : nullptr, Tt->cfg.ema_rate);
            if (pk_count) pack_weights(Tt->pack_table + pk_first, pk_count, pk_tiles, os);
            Tt->opt_done_lo = lo;
        }, 0, UB_KIND_OPTIM, 0, 0, 1);
    };
    View g{};
    for (int i = int(nodes.size()) - 1; i >= 0; --i) {
        Node& nd = nodes[i];
        if (T->mb_max_hw > 0 && !T->mb_fwd_only) nh = node_nh[size_t(i)], Bh = B / nh;  // the chains of this node's level
        if (nd.pushed) {  // gradient of a tensor that also fed a skip connection: sum both contributions
            View s = skipgrad[i];
            View sum = act(nd.out.C, nd.out.H, nd.out.W);
            View a = g;
            const size_t npix = size_t(Bh) * nd.out.H * nd.out.W;
            const int C = nd.out.C;
            a.H = s.H = nd.out.H, a.W = s.W = nd.out.W;  // (channel-slice views keep the geometry of their buffer)
            split([&](int hh) {
                const View am = mb(a, hh), sm = mb(s, hh), um = mb(sum, hh);
                Bk([=](cudaStream_t st) { add2(am.p, am.ld, sm.p, sm.ld, npix, C, um.p, um.ld, st); }, 1, UB_KIND_ELTWISE, 0,
                   3 * act_bytes(C, nd.out.H, nd.out.W) / nh);
            });
            g = sum;
        }
        if (i == 0) {  // only the 3-channel input conv is left: finalise the tcgen05 weight gradients before it
            emb_flush();
            fin_flush();
        }
        g = nd.bwd(g);
        while (cut_i < cuts.size() && nd.param_begin <= cuts[cut_i] && nd.param_begin > time_mlp_end) {
            // (all of this is on the weight-gradient branch, which has seen everything the main stream did so far:
            //  the main stream does not wait for it here)
            emb_flush();
            fin_flush();
            flush_bucket(nd.param_begin, flushed_hi, false);
            flushed_hi = nd.param_begin;
            while (cut_i < cuts.size() && cuts[cut_i] >= nd.param_begin) ++cut_i;
        }
    }
    emb_flush();
    fin_flush();
    join_side();
    // (a model without ResBlocks: the time-MLP backward has not been emitted yet), then the last bucket
    if (!time_mlp_bwd_done) {
        time_mlp_bwd();
        join_side();
    }
    flush_bucket(0, flushed_hi, true);
    Bk([=](cudaStream_t st) {  // join the communication stream
        if (Tt->world <= 1 || Tt->comm_off) return;
        cudaEventRecord(Tt->ev_join, Tt->comm_stream);
        cudaStreamWaitEvent(st, Tt->ev_join, 0);
        if (Tt->opt_stream_dirty) {
            cudaEventRecord(Tt->ev_join2, Tt->opt_stream);
            cudaStreamWaitEvent(st, Tt->ev_join2, 0);
            Tt->opt_stream_dirty = false;
        }
    }, 0);
    if (real() && getenv("UB_DEBUG_GNF"))
        fprintf(stderr, "[unet_b200] GroupNorm finished inside conv epilogues: %d forward, %d backward\n",
                gn_finished_fwd, gn_finished_bwd);
    return plan_errors ? UB_ERR_SHAPE : UB_OK;
}

}  // namespace

// ======================================================================================================
// trainer lifetime
// ======================================================================================================
extern "C" void ub_default_config(UbConfig* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->B = 32, cfg->C_in = 3, cfg->C_model = 64, cfg->C_out = 3, cfg->H = 64, cfg->W = 64;
    cfg->max_period = 1000, cfg->n_levels = 4;
    cfg->channel_mult[0] = 1, cfg->channel_mult[1] = 2, cfg->channel_mult[2] = 3, cfg->channel_mult[3] = 4;
    cfg->n_res_blocks = 2, cfg->att_start_level = 2, cfg->head_size = 32, cfg->gn_n_groups = 32;
    cfg->n_timesteps = 1000, cfg->seed = 0x5eedULL, cfg->use_cuda_graph = 1;
}

static int validate_config(const UbConfig& c) {
    if (c.B < 1 || c.C_in < 1 || c.C_in > 4 || c.C_out < 1 || c.C_out > 4 || c.n_levels < 1 || c.n_levels > 8) {
        set_err("bad config: B/C_in/C_out/n_levels out of range");
        return UB_ERR_SHAPE;
    }
    if (c.C_model % 64 != 0) {
        set_err("C_model must be a multiple of 64 (tcgen05 K blocks of 64 channels)");
        return UB_ERR_SHAPE;
    }
    if (c.head_size != 32) {
        set_err("head_size must be 32");
        return UB_ERR_SHAPE;
    }
    if ((c.H >> (c.n_levels - 1)) < 1 || c.H % (1 << (c.n_levels - 1)) || c.W % (1 << (c.n_levels - 1))) {
        set_err("H, W must be divisible by 2^(n_levels-1)");
        return UB_ERR_SHAPE;
    }
    if ((c.resblock_updown != 0 && c.resblock_updown != 1) || (c.use_scale_shift_norm != 0 && c.use_scale_shift_norm != 1)) {
        set_err("bad config: resblock_updown / use_scale_shift_norm must be 0 or 1");
        return UB_ERR_SHAPE;
    }
    if (c.num_classes < 0 || !(c.ema_rate >= 0.f && c.ema_rate < 1.f) || !(c.dropout >= 0.f && c.dropout < 1.f)) {
        set_err("bad config: num_classes must be >= 0, ema_rate and dropout in [0, 1)");
        return UB_ERR_SHAPE;
    }
    if (c.random_flip && c.W % 4) {
        set_err("random_flip needs W %% 4 == 0");
        return UB_ERR_SHAPE;
    }
    if ((size_t(c.C_in) * c.H * c.W) % 4) {
        set_err("C_in*H*W must be a multiple of 4");
        return UB_ERR_SHAPE;
    }
    for (int l = 0; l < c.n_levels; ++l)
        if ((c.channel_mult[l] * c.C_model) % c.gn_n_groups) {
            set_err("channels not divisible by gn groups");
            return UB_ERR_SHAPE;
        }
    return UB_OK;
}

extern "C" size_t ub_num_params(const UbConfig* cfg) {
    UbTrainer tmp;
    tmp.cfg = *cfg;
    if (validate_config(tmp.cfg)) return 0;
    tmp.arena.counting = tmp.zarena.counting = true;
    Builder b(&tmp);
    b.build();
    return b.poff;
}

static int upload_tables(UbTrainer* t) {
    if (t->h_emb.size() > size_t(kMaxEmbEntries) || t->h_pack.size() > 256) {
        set_err("table overflow");
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaMemcpy(t->emb_table, t->h_emb.data(), t->h_emb.size() * sizeof(SmallLinear), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t->temb_table, t->h_temb.data(), 2 * sizeof(SmallLinear), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t->pack_table, t->h_pack.data(), t->h_pack.size() * sizeof(PackEntry), cudaMemcpyHostToDevice));
    if (t->h_fin.size() > 128) {
        set_err("finalize table overflow");
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaMemcpy(t->fin_table, t->h_fin.data(), t->h_fin.size() * sizeof(WgradFinalizeEntry),
                        cudaMemcpyHostToDevice));
    t->n_pack = int(t->h_pack.size());
    return UB_OK;
}

static void run_pack(UbTrainer* t, cudaStream_t st) { pack_weights(t->pack_table, t->n_pack, t->pack_max_tiles, st); }

extern "C" int ub_trainer_create(UbTrainer** out, const UbConfig* cfg, int device) {
    *out = nullptr;
    int r = validate_config(*cfg);
    if (r) return r;
    CUDA_TRY(cudaSetDevice(device));
    igemm_init();
    igemm_rows_init();
    nhwc_ops_init();
    attn_init();
    attn_tc_init();
    UbTrainer* t = new UbTrainer();
    t->cfg = *cfg;
    t->device = device;
    {   // micro-batch chains (UbTrainer::n_mb): off by default -- measured 5.18 ms/step with one chain, 5.30 with two, 5.74
        // with four (profiles/r02_microbatch.txt): the latency that B-scaling exposes is per launch, and two chains double
        // the launches.  UB_MICROBATCH=2 / 4 enables (B must be divisible).
        int want = getenv("UB_MICROBATCH") ? atoi(getenv("UB_MICROBATCH")) : 1;
        if (getenv("UB_MICROBATCH_FWD") && atoi(getenv("UB_MICROBATCH_FWD")) >= 2)
            want = atoi(getenv("UB_MICROBATCH_FWD")), t->mb_fwd_only = true;
        t->n_mb = (want >= 2 && want <= 8 && cfg->B % want == 0) ? want : 1;
        if (t->n_mb == 1) t->mb_fwd_only = false;
        t->mb_max_hw = (t->n_mb > 1 && getenv("UB_MICROBATCH_LOWRES")) ? atoi(getenv("UB_MICROBATCH_LOWRES")) : 0;
    }
    // pass 1: count
    t->arena.counting = t->zarena.counting = true;
    {
        Builder b(t);
        b.build();
        t->nparams = b.poff;
    }
    const size_t arena_bytes = t->arena.off + 4096, zero_bytes = t->zarena.off + 4096;
    const size_t pbytes = (t->nparams * sizeof(float) + 255) & ~size_t(255);
    uint8_t* blob = nullptr;
    cudaError_t e = cudaMalloc(&blob, 4 * pbytes + arena_bytes + zero_bytes);
    if (e != cudaSuccess) {
        set_err("cudaMalloc of %zu MiB failed: %s", (4 * pbytes + arena_bytes + zero_bytes) >> 20,
                cudaGetErrorString(e));
        delete t;
        return UB_ERR_CUDA;
    }
    cudaMemset(blob, 0, 4 * pbytes + arena_bytes + zero_bytes);
    t->params = (float*)blob, t->grads = (float*)(blob + pbytes), t->m = (float*)(blob + 2 * pbytes),
    t->v = (float*)(blob + 3 * pbytes);
    t->arena.base = blob + 4 * pbytes, t->arena.cap = arena_bytes, t->arena.off = 0, t->arena.counting = false;
    t->zarena.base = blob + 4 * pbytes + arena_bytes, t->zarena.cap = zero_bytes, t->zarena.off = 0,
    t->zarena.counting = false;
    t->zero_base = t->zarena.base, t->zero_bytes = zero_bytes;
    if (cfg->ema_rate > 0.f) {
        if (cudaMalloc(&t->ema, pbytes) != cudaSuccess) {
            set_err("cudaMalloc of the moving-average arena (%zu MiB) failed", pbytes >> 20);
            t->ema = nullptr;
            ub_trainer_destroy(t);
            return UB_ERR_CUDA;
        }
        cudaMemset(t->ema, 0, pbytes);
    }
    // the main stream carries the critical path (forward, dgrad chain): highest priority; the weight-gradient branch
    // and the collectives fill the gaps (stream priorities become kernel-node priorities in the captured graph)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    static const bool use_prio = !(getenv("UB_NO_PRIO") && atoi(getenv("UB_NO_PRIO")) != 0);
    cudaStreamCreateWithPriority(&t->stream, cudaStreamNonBlocking, use_prio ? prio_hi : 0);
    cudaStreamCreateWithPriority(&t->comm_stream, cudaStreamNonBlocking, use_prio ? prio_hi : 0);
    cudaStreamCreateWithPriority(&t->side_stream, cudaStreamNonBlocking, use_prio ? prio_lo : 0);
    cudaStreamCreateWithPriority(&t->aux_stream, cudaStreamNonBlocking, use_prio ? prio_hi : 0);
    cudaEventCreateWithFlags(&t->ev_aux_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t->ev_aux_join, cudaEventDisableTiming);
    t->attn_concurrent = !(getenv("UB_ATTN_SERIAL") && atoi(getenv("UB_ATTN_SERIAL")) != 0) && (t->n_mb == 1 || t->mb_fwd_only);
    for (int i = 1; i < t->n_mb; ++i) {
        cudaStream_t ms = nullptr;
        cudaStreamCreateWithPriority(&ms, cudaStreamNonBlocking, use_prio ? prio_hi : 0);
        t->mb_streams.push_back(ms);
    }
    t->side_events.resize(4096);
    for (auto& ev : t->side_events) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (const char* e = getenv("UB_NO_SIDE_STREAM")) t->use_side = atoi(e) == 0;
    cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t->ev_join2, cudaEventDisableTiming);
    cudaStreamCreateWithPriority(&t->opt_stream, cudaStreamNonBlocking, use_prio ? prio_lo : 0);
    t->bucket_events.resize(32);  // [0, 16): branch -> communication stream, [16, 32): communication -> optimizer stream
    for (auto& ev : t->bucket_events) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    // pass 2: build
    {
        Builder b(t);
        r = b.build();
        if (r) {
            ub_trainer_destroy(t);
            return r;
        }
    }
    r = upload_tables(t);
    if (r) {
        ub_trainer_destroy(t);
        return r;
    }
    // diffusion tables: float32 cumprod of (1 - beta), beta = float32(linspace_f64(1e-4, 0.02)) scaled
    // (train_unet.py:811-826, 875-892; train_unet.cu:3131-3147)
    {
        const int n = cfg->n_timesteps;
        std::vector<float> sa(n), sb(n), bt(n);
        const double scale = 1000.0 / n, b0 = scale * 0.0001, b1 = scale * 0.02;
        float ac = 1.f;
        for (int i = 0; i < n; ++i) {
            const float beta = float(n > 1 ? b0 + (b1 - b0) * double(i) / double(n - 1) : b0);
            ac = ac * (1.f - beta);
            sa[i] = sqrtf(ac), sb[i] = sqrtf(1.f - ac), bt[i] = beta;
        }
        cudaMemcpy(t->betas, bt.data(), n * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(t->sqrt_ac, sa.data(), n * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(t->sqrt_1mac, sb.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    }
    const size_t img = size_t(cfg->B) * cfg->C_in * cfg->H * cfg->W;
    for (int k = 0; k < 2; ++k) {
        cudaMallocHost(&t->hs_x0[k], img * sizeof(float));
        cudaMallocHost(&t->hs_noise[k], img * sizeof(float));
        cudaMallocHost(&t->hs_t[k], cfg->B * sizeof(float));
        cudaEventCreateWithFlags(&t->ev_stage[k], cudaEventDisableTiming);
    }
    cudaMallocHost(&t->h_loss, 64);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_err("trainer init failed: %s", cudaGetErrorString(e));
        ub_trainer_destroy(t);
        return UB_ERR_CUDA;
    }
    *out = t;
    return UB_OK;
}

extern "C" void ub_trainer_destroy(UbTrainer* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    if (t->graph_exec) cudaGraphExecDestroy(t->graph_exec);
    if (t->samp_graph) cudaGraphExecDestroy(t->samp_graph);
    if (t->comm && nccl().ok) nccl().CommDestroy(t->comm);
    if (t->params) cudaFree(t->params);
    if (t->ema) cudaFree(t->ema);
    for (int k = 0; k < 2; ++k) {
        if (t->hs_x0[k]) cudaFreeHost(t->hs_x0[k]);
        if (t->hs_noise[k]) cudaFreeHost(t->hs_noise[k]);
        if (t->hs_t[k]) cudaFreeHost(t->hs_t[k]);
        if (t->ev_stage[k]) cudaEventDestroy(t->ev_stage[k]);
    }
    if (t->h_loss) cudaFreeHost(t->h_loss);
    if (t->x0_pref) cudaFree(t->x0_pref);
    if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
    if (t->ev_pref) cudaEventDestroy(t->ev_pref);
    if (t->ev_pref_free) cudaEventDestroy(t->ev_pref_free);
    for (auto ev : t->bucket_events) cudaEventDestroy(ev);
    if (t->ev_join) cudaEventDestroy(t->ev_join);
    if (t->ev_join2) cudaEventDestroy(t->ev_join2);
    if (t->opt_stream) cudaStreamDestroy(t->opt_stream);
    if (t->stream) cudaStreamDestroy(t->stream);
    if (t->comm_stream) cudaStreamDestroy(t->comm_stream);
    if (t->side_stream) cudaStreamDestroy(t->side_stream);
    if (t->aux_stream) cudaStreamDestroy(t->aux_stream);
    if (t->ev_aux_fork) cudaEventDestroy(t->ev_aux_fork);
    if (t->ev_aux_join) cudaEventDestroy(t->ev_aux_join);
    for (auto ms : t->mb_streams) cudaStreamDestroy(ms);
    for (auto ev : t->side_events) cudaEventDestroy(ev);
    delete t;
}

// ======================================================================================================
// running the tapes
// ======================================================================================================
struct StepOpts {
    bool gen_t, gen_noise, update;
    float lr, b1, b2, eps, wd;
};

// everything of one step that lives on the device, in stream order (this is what gets captured into the graph)
static void enqueue_step(UbTrainer* t, const StepOpts& o, cudaStream_t st) {
    const UbConfig& c = t->cfg;
    cudaMemsetAsync(t->zero_base, 0, t->zero_bytes, st);
    diffusion_prepare(t->x0, t->sqrt_ac, t->sqrt_1mac, c.B, size_t(c.C_in) * c.H * c.W, c.n_timesteps, c.seed,
                      t->step_dev, o.gen_t ? 1 : 0, o.gen_noise ? 1 : 0, t->tsteps, t->noise, t->xt, st, c.W, t->flips);
    if (t->drop_ctl) dropout_set_ctl(t->drop_ctl, c.dropout, c.seed, t->step_dev, st);  // training: masks of this step
    static const bool debug_sync = getenv("UB_DEBUG_SYNC") != nullptr;  // eager runs only: find the faulting op
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (debug_sync) cudaStreamIsCapturing(st, &cap);
    bool side_dirty = false, aux_dirty = false;
    auto next_event = [&]() { return t->side_events[t->side_ev_next++ % t->side_events.size()]; };
    // micro-batch chains: `forked` while the chains of micro-batches 1.. run beside the main stream
    bool forked = false;
    auto fork_mb = [&]() {
        if (forked || t->mb_streams.empty()) return;
        cudaEvent_t ev = next_event();
        cudaEventRecord(ev, st);
        for (cudaStream_t ms : t->mb_streams) cudaStreamWaitEvent(ms, ev, 0);
        forked = true;
    };
    auto join_mb = [&]() {
        if (!forked) return;
        for (cudaStream_t ms : t->mb_streams) {
            cudaEvent_t ev = next_event();
            cudaEventRecord(ev, ms);
            cudaStreamWaitEvent(st, ev, 0);
        }
        forked = false;
    };
    auto run = [&](std::vector<UbTrainer::Op>& ops, std::vector<UbTrainer::OpInfo>& info, const char* what) {
        for (size_t i = 0; i < ops.size(); ++i) {
            static const bool skip_side = getenv("UB_DEBUG_SKIP_SIDE") != nullptr;  // timing experiments only
            if (info[i].side == 1 && skip_side) continue;
            if (info[i].side == 1 && t->use_side) {  // fork: the branch sees everything enqueued on the main chains so far
                cudaEvent_t ev = next_event();
                cudaEventRecord(ev, st);
                cudaStreamWaitEvent(t->side_stream, ev, 0);
                if (forked)
                    for (cudaStream_t ms : t->mb_streams) {
                        cudaEvent_t e2 = next_event();
                        cudaEventRecord(e2, ms);
                        cudaStreamWaitEvent(t->side_stream, e2, 0);
                    }
                ops[i](t->side_stream);
                side_dirty = true;
            } else if (info[i].side == 3 && t->use_side) {  // a branch off the weight-gradient branch (time-MLP backward)
                cudaEvent_t ev = next_event();
                cudaEventRecord(ev, side_dirty ? t->side_stream : st);
                cudaStreamWaitEvent(t->aux_stream, ev, 0);
                ops[i](t->aux_stream);
                aux_dirty = true;
            } else if (info[i].side == 2) {
                if (aux_dirty) {
                    cudaEvent_t ev = next_event();
                    cudaEventRecord(ev, t->aux_stream);
                    cudaStreamWaitEvent(st, ev, 0);
                    aux_dirty = false;
                }
                if (side_dirty) {
                    cudaEvent_t ev = next_event();
                    cudaEventRecord(ev, t->side_stream);
                    cudaStreamWaitEvent(st, ev, 0);
                    if (forked)
                        for (cudaStream_t ms : t->mb_streams) cudaStreamWaitEvent(ms, ev, 0);
                    side_dirty = false;
                }
            } else if (info[i].half > 0 && size_t(info[i].half) <= t->mb_streams.size()) {
                fork_mb();
                ops[i](t->mb_streams[info[i].half - 1]);
            } else {
                if (info[i].half < 0 || info[i].side == 1 || info[i].side == 3) join_mb();  // whole-batch op: every chain must have arrived
                else fork_mb();
                ops[i](st);
            }
            if (debug_sync && cap == cudaStreamCaptureStatusNone) {
                cudaError_t e = cudaStreamSynchronize(st);
                for (cudaStream_t ms : t->mb_streams)
                    if (e == cudaSuccess) e = cudaStreamSynchronize(ms);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) {
                    fprintf(stderr, "[unet_b200] %s op %zu (kind %d): %s\n", what, i, info[i].kind,
                            cudaGetErrorString(e));
                    break;
                }
            }
        }
    };
    static const bool opt_overlap = !(getenv("UB_NO_OPT_OVERLAP") && atoi(getenv("UB_NO_OPT_OVERLAP")) != 0);
    t->opt_in_tape = o.update && opt_overlap && t->opt_overlap_ok;
    t->o_lr = o.lr, t->o_b1 = o.b1, t->o_b2 = o.b2, t->o_eps = o.eps, t->o_wd = o.wd;
    t->opt_done_lo = t->nparams;
    run(t->fwd_ops, t->fwd_info, "forward");
    run(t->bwd_ops, t->bwd_info, "backward");
    join_mb();
    if (aux_dirty) {
        cudaEvent_t ev = next_event();
        cudaEventRecord(ev, t->aux_stream);
        cudaStreamWaitEvent(st, ev, 0);
    }
    if (side_dirty) {  // (the tape ends with a join; this only guards against a tape that forgot it)
        cudaEvent_t ev = next_event();
        cudaEventRecord(ev, t->side_stream);
        cudaStreamWaitEvent(st, ev, 0);
    }
    if (o.update) {
        // the buckets the tape already updated are [opt_done_lo, nparams); the rest (the last bucket) is updated here
        const size_t n_left = t->opt_done_lo;
        adamw_step(t->params, t->grads, t->m, t->v, n_left, o.lr, o.b1, o.b2, o.eps, o.wd, 1.f / float(t->world),
                   t->step_dev, st, t->hp_active, t->ema, t->cfg.ema_rate);
        int cnt = 0, tiles = 0;
        while (cnt < t->n_pack && t->pack_off[cnt] < n_left) {
            tiles = std::max(tiles, ((t->h_pack[cnt].Cout + 31) / 32) * ((t->h_pack[cnt].Cin + 31) / 32));
            ++cnt;
        }
        if (cnt) pack_weights(t->pack_table, cnt, tiles, st);
        increment_step(t->step_dev, st);
    }
    t->opt_in_tape = false;
}

static int ensure_packed(UbTrainer* t) {
    run_pack(t, t->stream);
    return UB_OK;
}

// `sync_call`: the caller blocks until the step has finished (it asked for the loss), so a page-locked source buffer
// cannot be overwritten while the copy is in flight: it is DMA'd directly, without the pass through the staging slot
// (1.5 MB of host memcpy per step at B = 32, ~0.1 ms of the end-to-end step).
static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
static int stage_inputs(UbTrainer* t, const float* x0_host, const float* t_host, const float* noise_host,
                        bool sync_call = false) {
    const UbConfig& c = t->cfg;
    const size_t img = size_t(c.B) * c.C_in * c.H * c.W;
    if (t->pref_src) {  // a prefetched batch is waiting on the device
        const bool hit = t->pref_src == x0_host && !t_host && !noise_host;
        t->pref_src = nullptr;  // (a miss drops it: the caller went another way)
        if (hit) {
            CUDA_TRY(cudaStreamWaitEvent(t->stream, t->ev_pref, 0));
            CUDA_TRY(cudaMemcpyAsync(t->x0, t->x0_pref, img * sizeof(float), cudaMemcpyDeviceToDevice, t->stream));
            CUDA_TRY(cudaEventRecord(t->ev_pref_free, t->stream));
            t->pref_free_pending = true;
            return UB_OK;
        }
    }
    if (sync_call && !t_host && !noise_host && is_pinned_host(x0_host)) {
        CUDA_TRY(cudaMemcpyAsync(t->x0, x0_host, img * sizeof(float), cudaMemcpyHostToDevice, t->stream));
        return UB_OK;
    }
    const int k = int(t->stage_next++ & 1u);
    if (t->stage_used[k]) CUDA_TRY(cudaEventSynchronize(t->ev_stage[k]));  // the copies that last read this slot are done
    memcpy(t->hs_x0[k], x0_host, img * sizeof(float));
    CUDA_TRY(cudaMemcpyAsync(t->x0, t->hs_x0[k], img * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    if (t_host) {
        memcpy(t->hs_t[k], t_host, c.B * sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(t->tsteps, t->hs_t[k], c.B * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    }
    if (noise_host) {
        memcpy(t->hs_noise[k], noise_host, img * sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(t->noise, t->hs_noise[k], img * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    }
    CUDA_TRY(cudaEventRecord(t->ev_stage[k], t->stream));
    t->stage_used[k] = true;
    return UB_OK;
}

// Enqueue the H2D copy of the announced next batch on the copy stream (page-locked sources only: a pageable source
// would make the copy synchronous with the host).  Called right after the current step has been launched.
static int start_prefetch(UbTrainer* t) {
    const float* src = t->next_hint;
    t->next_hint = nullptr;
    if (!src || !is_pinned_host(src)) return UB_OK;
    const UbConfig& c = t->cfg;
    const size_t img = size_t(c.B) * c.C_in * c.H * c.W;
    if (!t->x0_pref) {
        CUDA_TRY(cudaMalloc(&t->x0_pref, img * sizeof(float)));
        CUDA_TRY(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_pref, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_pref_free, cudaEventDisableTiming));
    }
    if (t->pref_free_pending) CUDA_TRY(cudaStreamWaitEvent(t->copy_stream, t->ev_pref_free, 0));
    t->pref_free_pending = false;
    CUDA_TRY(cudaMemcpyAsync(t->x0_pref, src, img * sizeof(float), cudaMemcpyHostToDevice, t->copy_stream));
    CUDA_TRY(cudaEventRecord(t->ev_pref, t->copy_stream));
    t->pref_src = src;
    return UB_OK;
}

extern "C" int ub_trainer_set_next_batch(UbTrainer* t, const float* next_x0_host) {
    if (!t) return UB_ERR_STATE;
    t->next_hint = next_x0_host;
    return UB_OK;
}

static int fetch_loss(UbTrainer* t, float* loss_out) {
    CUDA_TRY(cudaMemcpyAsync(t->h_loss, t->loss, sizeof(float), cudaMemcpyDeviceToHost, t->stream));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    *loss_out = *t->h_loss;
    return UB_OK;
}

extern "C" int ub_trainer_forward_backward(UbTrainer* t, const float* x0_host, const float* t_host,
                                           const float* noise_host, float* loss_out) {
    CUDA_TRY(cudaSetDevice(t->device));
    int r = stage_inputs(t, x0_host, t_host, noise_host);
    if (r) return r;
    // gradients accumulate in parts of the arena (GroupNorm affine, biases): start from zero like unet_zero_grad
    CUDA_TRY(cudaMemsetAsync(t->grads, 0, t->nparams * sizeof(float), t->stream));
    StepOpts o{t_host == nullptr, noise_host == nullptr, false, 0, 0, 0, 0, 0};
    enqueue_step(t, o, t->stream);
    t->have_grads = true;
    CUDA_TRY(cudaGetLastError());
    if (loss_out) return fetch_loss(t, loss_out);
    return UB_OK;
}

extern "C" int ub_trainer_update(UbTrainer* t, float lr, float beta1, float beta2, float eps, float weight_decay) {
    CUDA_TRY(cudaSetDevice(t->device));
    if (!t->have_grads) {
        set_err("ub_trainer_update called without gradients");
        return UB_ERR_STATE;
    }
    adamw_step(t->params, t->grads, t->m, t->v, t->nparams, lr, beta1, beta2, eps, weight_decay,
               1.f / float(t->world), t->step_dev, t->stream, nullptr, t->ema, t->cfg.ema_rate);
    run_pack(t, t->stream);
    increment_step(t->step_dev, t->stream);
    t->have_grads = false;
    t->host_step++;
    CUDA_TRY(cudaGetLastError());
    return UB_OK;
}

static int launch_step(UbTrainer* t, const StepOpts& o) {
    if (!t->cfg.use_cuda_graph) {
        enqueue_step(t, o, t->stream);
        ub_count_launches((unsigned long long)ub_trainer_launches_per_step(t));
        CUDA_TRY(cudaGetLastError());
        return UB_OK;
    }
    // (hyper-parameters are NOT part of the graph's identity: its AdamW nodes read them from hp_dev)
    const bool same = t->graph_valid && t->graph_gen_t == o.gen_t && t->graph_gen_noise == o.gen_noise;
    {
        const float hp[5] = {o.lr, o.b1, o.b2, o.eps, o.wd};  // pageable source: staged by the driver before the call returns
        if (!same || !t->hp_uploaded || memcmp(hp, t->hp_last, sizeof hp) != 0) {
            CUDA_TRY(cudaMemcpyAsync(t->hp_dev, hp, sizeof hp, cudaMemcpyHostToDevice, t->stream));
            memcpy(t->hp_last, hp, sizeof hp);
            t->hp_uploaded = true;
        }
    }
    t->hp_active = t->hp_dev;
    if (!same) {
        if (t->graph_exec) cudaGraphExecDestroy(t->graph_exec), t->graph_exec = nullptr;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_step(t, o, t->stream);
        t->hp_active = nullptr;
        cudaError_t e = cudaStreamEndCapture(t->stream, &graph);
        if (e != cudaSuccess || !graph) {
            set_err("graph capture failed: %s", cudaGetErrorString(e));
            return UB_ERR_CUDA;
        }
        e = cudaGraphInstantiate(&t->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            set_err("graph instantiate failed: %s", cudaGetErrorString(e));
            return UB_ERR_CUDA;
        }
        t->graph_valid = true, t->graph_gen_t = o.gen_t, t->graph_gen_noise = o.gen_noise;
        t->g_lr = o.lr, t->g_b1 = o.b1, t->g_b2 = o.b2, t->g_eps = o.eps, t->g_wd = o.wd;
    }
    t->hp_active = nullptr;
    CUDA_TRY(cudaGraphLaunch(t->graph_exec, t->stream));
    ub_count_launches((unsigned long long)ub_trainer_launches_per_step(t));
    return UB_OK;
}

extern "C" int ub_trainer_train_step(UbTrainer* t, const float* x0_host, const float* t_host, const float* noise_host,
                                     float lr, float beta1, float beta2, float eps, float weight_decay,
                                     float* loss_out) {
    CUDA_TRY(cudaSetDevice(t->device));
    int r = stage_inputs(t, x0_host, t_host, noise_host, loss_out != nullptr);
    if (r) return r;
    StepOpts o{t_host == nullptr, noise_host == nullptr, true, lr, beta1, beta2, eps, weight_decay};
    r = launch_step(t, o);
    if (r) return r;
    t->host_step++;
    t->have_grads = false;
    if (t->next_hint && (r = start_prefetch(t)) != UB_OK) return r;
    if (loss_out) return fetch_loss(t, loss_out);
    return UB_OK;
}

extern "C" int ub_trainer_train_step_device(UbTrainer* t, const float* x0_dev, float lr, float beta1, float beta2,
                                            float eps, float weight_decay) {
    CUDA_TRY(cudaSetDevice(t->device));
    const UbConfig& c = t->cfg;
    const size_t img = size_t(c.B) * c.C_in * c.H * c.W;
    if (x0_dev != t->x0)
        CUDA_TRY(cudaMemcpyAsync(t->x0, x0_dev, img * sizeof(float), cudaMemcpyDeviceToDevice, t->stream));
    StepOpts o{true, true, true, lr, beta1, beta2, eps, weight_decay};
    int r = launch_step(t, o);
    if (r) return r;
    t->host_step++;
    return UB_OK;
}

// Eager replay of one step with a CUDA event pair around every tape op (device time per kernel class).
// Every op is enqueued R times back to back between its two events and the interval is divided by R (R =
// UB_PROFILE_REPLAY, default 4; 1 = one launch per event pair as in round 1).  With a single launch per pair the
// interval is dominated by the event records and the un-pipelined launch for the ~300 kernels of a step that run for
// less than 10 us (a 256->256 conv at 8x8: 16.7 us between events, 9.6 us in the kernel's own phase trace); replicas
// follow each other through programmatic dependent launch exactly as the different kernels of the captured step do,
// so interval / R is the per-launch device time inside a pipelined stream.  Accumulating ops (statistics / weight-
// gradient REDs) simply accumulate R times: the profile's results are discarded (state is restored below).
extern "C" int ub_trainer_profile(UbTrainer* t, int reps, UbProfile* out) {
    CUDA_TRY(cudaSetDevice(t->device));
    memset(out, 0, sizeof(*out));
    if (reps < 1) reps = 1;
    static const int R = [] {
        const int v = getenv("UB_PROFILE_REPLAY") ? atoi(getenv("UB_PROFILE_REPLAY")) : 4;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    const float invR = 1.f / float(R);
    const UbConfig& c = t->cfg;
    const size_t nops = t->fwd_ops.size() + t->bwd_ops.size() + 2;
    std::vector<cudaEvent_t> ev(nops + 1);
    std::vector<float> op_us(nops, 0.f);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    cudaStream_t st = t->stream;
    t->comm_off = true;  // a profile is rank-local: the peers are not replaying with us
    // The timed replays run REAL optimizer steps (their time is part of the profile): snapshot parameters, AdamW moments
    // and the step counter and restore them afterwards, so that profiling in the middle of a training run changes nothing.
    float* snap = nullptr;
    int snap_step = 0;
    CUDA_TRY(cudaStreamSynchronize(st));
    if (cudaMalloc(&snap, 3 * t->nparams * sizeof(float)) == cudaSuccess) {
        cudaMemcpy(snap, t->params, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(snap + t->nparams, t->m, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(snap + 2 * t->nparams, t->v, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(&snap_step, t->step_dev, sizeof(int), cudaMemcpyDeviceToHost);
    } else {
        cudaGetLastError();
        snap = nullptr;
    }
    for (int rep = 0; rep < reps; ++rep) {
        size_t k = 0;
        // the host needs ~3 us per launch + event, many kernels run for less: park the stream for a few ms so that the
        // whole tape is queued before the first op starts and the event intervals are pure device time
        stream_delay(unsigned(3000 + 3000 * R), st);
        cudaEventRecord(ev[k++], st);
        cudaMemsetAsync(t->zero_base, 0, t->zero_bytes, st);
        diffusion_prepare(t->x0, t->sqrt_ac, t->sqrt_1mac, c.B, size_t(c.C_in) * c.H * c.W, c.n_timesteps, c.seed,
                          t->step_dev, 1, 1, t->tsteps, t->noise, t->xt, st, c.W, t->flips);
        if (t->drop_ctl) dropout_set_ctl(t->drop_ctl, c.dropout, c.seed, t->step_dev, st);
        cudaEventRecord(ev[k++], st);
        for (auto& op : t->fwd_ops) {
            for (int r = 0; r < R; ++r) op(st);
            cudaEventRecord(ev[k++], st);
        }
        for (auto& op : t->bwd_ops) {
            for (int r = 0; r < R; ++r) op(st);
            cudaEventRecord(ev[k++], st);
        }
        adamw_step(t->params, t->grads, t->m, t->v, t->nparams, t->g_lr > 0 ? t->g_lr : 1e-4f, 0.9f, 0.999f, 1e-8f, 0.f,
                   1.f / float(t->world), t->step_dev, st);  // (no moving-average update: it is not snapshotted)
        run_pack(t, st);
        increment_step(t->step_dev, st);
        cudaEventRecord(ev[k++], st);
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0;
        size_t j = 1;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        out->ms[UB_KIND_OPTIM] += ms;
        for (size_t i = 0; i < t->fwd_ops.size(); ++i, ++j) {
            cudaEventElapsedTime(&ms, ev[j], ev[j + 1]);
            ms *= invR;
            out->ms[t->fwd_info[i].kind] += ms;
            op_us[i] = ms * 1e3f;
        }
        for (size_t i = 0; i < t->bwd_ops.size(); ++i, ++j) {
            cudaEventElapsedTime(&ms, ev[j], ev[j + 1]);
            ms *= invR;
            out->ms[t->bwd_info[i].kind] += ms;
            op_us[t->fwd_ops.size() + i] = ms * 1e3f;
        }
        cudaEventElapsedTime(&ms, ev[j], ev[j + 1]);
        out->ms[UB_KIND_OPTIM] += ms;
    }
    if (snap) {
        cudaMemcpy(t->params, snap, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(t->m, snap + t->nparams, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(t->v, snap + 2 * t->nparams, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
        cudaMemcpy(t->step_dev, &snap_step, sizeof(int), cudaMemcpyHostToDevice);
        cudaMemset(t->grads, 0, t->nparams * sizeof(float));
        run_pack(t, st);
        cudaStreamSynchronize(st);
        cudaFree(snap);
    }
    t->comm_off = false;
    for (auto& e : ev) cudaEventDestroy(e);
    if (const char* path = getenv("UB_PROFILE_DUMP")) {  // per-op times of the LAST repetition
        if (FILE* f = fopen(path, "w")) {
            for (size_t i = 0; i < t->fwd_ops.size(); ++i)
                fprintf(f, "fwd\t%zu\t%d\t%.3f\t%.4g\t%.4g\t%s\n", i, t->fwd_info[i].kind, op_us[i],
                        t->fwd_info[i].flops, t->fwd_info[i].bytes, t->fwd_info[i].label.c_str());
            for (size_t i = 0; i < t->bwd_ops.size(); ++i)
                fprintf(f, "bwd\t%zu\t%d\t%.3f\t%.4g\t%.4g\t%s\n", i, t->bwd_info[i].kind,
                        op_us[t->fwd_ops.size() + i], t->bwd_info[i].flops, t->bwd_info[i].bytes,
                        t->bwd_info[i].label.c_str());
            fclose(f);
        }
    }
    for (int k = 0; k < UB_NUM_KINDS; ++k) out->ms[k] /= reps, out->total_ms += out->ms[k];
    auto acc = [&](const std::vector<UbTrainer::OpInfo>& v) {
        for (auto& i : v) out->flops[i.kind] += i.flops, out->bytes[i.kind] += i.bytes, out->launches[i.kind] += i.launches;
    };
    acc(t->fwd_info), acc(t->bwd_info);
    out->launches[UB_KIND_OPTIM] += 5;
    const double img = double(c.B) * c.C_in * c.H * c.W * 4;
    out->bytes[UB_KIND_OPTIM] += 3 * img + 7.0 * 4 * t->nparams + 4.0 * t->nparams + 2.0 * 2 * t->nparams;
    t->host_step += reps;
    return UB_OK;
}

extern "C" int ub_trainer_sync(UbTrainer* t) {
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    return UB_OK;
}
extern "C" int ub_trainer_last_loss(UbTrainer* t, float* loss_out) {
    CUDA_TRY(cudaSetDevice(t->device));
    return fetch_loss(t, loss_out);
}
extern "C" void* ub_trainer_stream(UbTrainer* t) { return (void*)t->stream; }
extern "C" int ub_trainer_launches_per_step(UbTrainer* t) {
    // tapes + memset-free extras: diffusion (2), adamw, pack, step increment
    return t->launches_fwd + t->launches_bwd + 5;
}

extern "C" int ub_trainer_predict(UbTrainer* t, const float* xt_host, const float* t_host, float* out_host) {
    CUDA_TRY(cudaSetDevice(t->device));
    const UbConfig& c = t->cfg;
    const size_t img = size_t(c.B) * c.C_in * c.H * c.W, oimg = size_t(c.B) * c.C_out * c.H * c.W;
    CUDA_TRY(cudaMemcpyAsync(t->xt, xt_host, img * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    CUDA_TRY(cudaMemcpyAsync(t->tsteps, t_host, c.B * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    CUDA_TRY(cudaMemsetAsync(t->zero_base, 0, t->zero_bytes, t->stream));
    CUDA_TRY(cudaMemsetAsync(t->noise, 0, img * sizeof(float), t->stream));
    if (t->drop_ctl) dropout_set_ctl(t->drop_ctl, 0.f, c.seed, t->step_dev, t->stream);  // inference: no dropout
    for (auto& op : t->fwd_ops) op(t->stream);
    CUDA_TRY(cudaMemcpyAsync(out_host, t->out, oimg * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    return UB_OK;
}

// generate.py:29-79 on the device (see include/unet_b200.h)
extern "C" int ub_trainer_sample(UbTrainer* t, const float* x_init_host, int t_start, int t_end,
                                 const float* noise_host, unsigned long long seed, float* out_host) {
    CUDA_TRY(cudaSetDevice(t->device));
    const UbConfig& c = t->cfg;
    if (c.C_in != c.C_out) {
        set_err("sampling needs C_in == C_out");
        return UB_ERR_SHAPE;
    }
    if (t_start < 0) t_start = c.n_timesteps - 1;
    if (t_end < 0) t_end = 2;
    if (t_start >= c.n_timesteps || t_end < 2 || t_end > t_start) {
        set_err("sampling range must satisfy 2 <= t_end <= t_start < n_timesteps");
        return UB_ERR_SHAPE;
    }
    const size_t img = size_t(c.B) * c.C_in * c.H * c.W;
    cudaStream_t st = t->stream;
    if (x_init_host)
        CUDA_TRY(cudaMemcpyAsync(t->xt, x_init_host, img * sizeof(float), cudaMemcpyHostToDevice, st));
    else
        fill_normal(t->xt, img, seed, st);
    const int state[2] = {t_start, 0};
    CUDA_TRY(cudaMemcpyAsync(t->samp_state, state, sizeof state, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(t->noise, 0, img * sizeof(float), st));  // (the forward tape ends with the MSE vs noise)
    const bool inject = noise_host != nullptr;
    if (!t->samp_graph || t->samp_graph_noise != inject) {
        if (t->samp_graph) cudaGraphExecDestroy(t->samp_graph), t->samp_graph = nullptr;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        cudaMemsetAsync(t->zero_base, 0, t->zero_bytes, st);
        sample_set_t(t->samp_state, c.B, t->tsteps, st);
        if (t->drop_ctl) dropout_set_ctl(t->drop_ctl, 0.f, c.seed, t->step_dev, st);  // inference: no dropout
        for (auto& op : t->fwd_ops) op(st);
        ddpm_step(t->xt, t->out, t->betas, t->sqrt_ac, t->sqrt_1mac, inject ? t->samp_z : nullptr, img, seed,
                  t->samp_state, t->samp_state + 1, st);
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (e != cudaSuccess || !graph) {
            set_err("sampling graph capture failed: %s", cudaGetErrorString(e));
            return UB_ERR_CUDA;
        }
        e = cudaGraphInstantiate(&t->samp_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            set_err("sampling graph instantiate failed: %s", cudaGetErrorString(e));
            return UB_ERR_CUDA;
        }
        t->samp_graph_noise = inject;
    }
    for (int tt = t_start, it = 0; tt >= t_end; --tt, ++it) {
        if (inject)
            CUDA_TRY(cudaMemcpyAsync(t->samp_z, noise_host + size_t(it) * img, img * sizeof(float),
                                     cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaGraphLaunch(t->samp_graph, st));
        ub_count_launches((unsigned long long)(t->launches_fwd + 3));
    }
    CUDA_TRY(cudaMemcpyAsync(out_host, t->xt, img * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return UB_OK;
}

// ======================================================================================================
// parameter / gradient access and checkpoints
// ======================================================================================================
static int copy_out(UbTrainer* t, const float* dev, float* host, size_t n, size_t expect) {
    if (n != expect) {
        set_err("size mismatch: got %zu expected %zu", n, expect);
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(host, dev, n * sizeof(float), cudaMemcpyDeviceToHost));
    return UB_OK;
}
extern "C" int ub_trainer_set_params(UbTrainer* t, const float* host, size_t n) {
    if (n != t->nparams) {
        set_err("size mismatch: got %zu expected %zu", n, t->nparams);
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(t->params, host, n * sizeof(float), cudaMemcpyHostToDevice));
    if (t->ema) CUDA_TRY(cudaMemcpy(t->ema, t->params, n * sizeof(float), cudaMemcpyDeviceToDevice));
    ensure_packed(t);
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    return UB_OK;
}
extern "C" int ub_trainer_get_params(UbTrainer* t, float* host, size_t n) {
    return copy_out(t, t->params, host, n, t->nparams);
}
extern "C" int ub_trainer_get_ema(UbTrainer* t, float* host, size_t n) {
    if (!t->ema) {
        set_err("no moving average: the trainer was created with cfg.ema_rate = 0");
        return UB_ERR_STATE;
    }
    return copy_out(t, t->ema, host, n, t->nparams);
}
extern "C" int ub_trainer_get_dropout_mask(UbTrainer* t, int block, unsigned char* host, size_t n) {
    if (!t->drop_ctl) {
        set_err("no dropout: the trainer was created with cfg.dropout = 0");
        return UB_ERR_STATE;
    }
    if (block < 0 || size_t(block) >= t->drop_layers.size()) {
        set_err("dropout mask: block %d outside [0, %zu)", block, t->drop_layers.size());
        return UB_ERR_SHAPE;
    }
    const auto& d = t->drop_layers[size_t(block)];
    const size_t want = size_t(t->cfg.B) * d.C * d.H * d.W;
    if (n != want) {
        set_err("dropout mask of block %d has %zu elements (B x %d x %d x %d, NHWC order), got %zu", block, want, d.H, d.W,
                d.C, n);
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    unsigned char* dev = nullptr;
    CUDA_TRY(cudaMalloc(&dev, want));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    dropout_mask(t->drop_ctl, t->cfg.dropout, unsigned(block), want, dev, t->stream);
    cudaError_t e = cudaMemcpyAsync(host, dev, want, cudaMemcpyDeviceToHost, t->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
    cudaFree(dev);
    if (e != cudaSuccess) {
        set_err("dropout mask readback failed: %s", cudaGetErrorString(e));
        return UB_ERR_CUDA;
    }
    return UB_OK;
}
extern "C" int ub_trainer_set_ema(UbTrainer* t, const float* host, size_t n) {
    if (!t->ema) {
        set_err("no moving average: the trainer was created with cfg.ema_rate = 0");
        return UB_ERR_STATE;
    }
    if (n != t->nparams) {
        set_err("size mismatch: got %zu expected %zu", n, t->nparams);
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(t->ema, host, n * sizeof(float), cudaMemcpyHostToDevice));
    return UB_OK;
}
extern "C" int ub_trainer_set_labels(UbTrainer* t, const int* labels_host, size_t n) {
    if (!t->labels) {
        set_err("not a class-conditional model: the trainer was created with cfg.num_classes = 0");
        return UB_ERR_STATE;
    }
    if (n != size_t(t->cfg.B)) {
        set_err("size mismatch: got %zu labels, batch is %d", n, t->cfg.B);
        return UB_ERR_SHAPE;
    }
    for (size_t i = 0; i < n; ++i)
        if (labels_host[i] < 0 || labels_host[i] >= t->cfg.num_classes) {
            set_err("label %d of image %zu outside [0, %d)", labels_host[i], i, t->cfg.num_classes);
            return UB_ERR_SHAPE;
        }
    CUDA_TRY(cudaSetDevice(t->device));
    // the captured step reads the labels from this fixed buffer: wait for the steps in flight, then overwrite it
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(t->labels, labels_host, n * sizeof(int), cudaMemcpyHostToDevice));
    return UB_OK;
}
extern "C" int ub_trainer_get_grads(UbTrainer* t, float* host, size_t n) {
    return copy_out(t, t->grads, host, n, t->nparams);
}
extern "C" int ub_trainer_get_output(UbTrainer* t, float* host, size_t n) {
    const UbConfig& c = t->cfg;
    return copy_out(t, t->out, host, n, size_t(c.B) * c.C_out * c.H * c.W);
}
extern "C" int ub_trainer_get_batch(UbTrainer* t, float* host, size_t n) {
    const UbConfig& c = t->cfg;
    return copy_out(t, t->x0, host, n, size_t(c.B) * c.C_in * c.H * c.W);
}
extern "C" int ub_trainer_get_dinput(UbTrainer* t, float* host, size_t n) {
    if (!t->dxt) {
        set_err("dL/d(x_t) is only computed when the trainer was created with cfg.compute_dinput = 1");
        return UB_ERR_STATE;
    }
    const UbConfig& c = t->cfg;
    return copy_out(t, t->dxt, host, n, size_t(c.B) * c.C_in * c.H * c.W);
}

extern "C" int ub_trainer_get_flips(UbTrainer* t, int* host, size_t n) {
    if (!t->flips) {
        set_err("flip decisions exist only when the trainer was created with cfg.random_flip = 1");
        return UB_ERR_STATE;
    }
    if (n != size_t(t->cfg.B)) {
        set_err("ub_trainer_get_flips: expected %d entries", t->cfg.B);
        return UB_ERR_SHAPE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(host, t->flips, n * sizeof(int), cudaMemcpyDeviceToHost));
    return UB_OK;
}

static const int kModelMagic = 12345678;  // train_unet.py:781
static const int kStepMagic = 0x55425354;  // "UBST": header[10] holds the AdamW step count of a checkpoint written here

extern "C" int ub_read_checkpoint_header(const char* path, UbConfig* cfg) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_err("cannot open %s", path);
        return UB_ERR_IO;
    }
    int header[256];
    const size_t got = fread(header, sizeof(int), 256, f);
    fclose(f);
    if (got != 256 || header[0] != kModelMagic) {
        set_err("%s: bad header / magic", path);
        return UB_ERR_IO;
    }
    cfg->B = header[1], cfg->C_in = header[2], cfg->C_model = header[3], cfg->C_out = header[4];
    cfg->H = header[5], cfg->W = header[6], cfg->max_period = header[7];
    return UB_OK;
}

extern "C" int ub_trainer_load(UbTrainer* t, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_err("cannot open %s", path);
        return UB_ERR_IO;
    }
    int header[256];
    if (fread(header, sizeof(int), 256, f) != 256 || header[0] != kModelMagic) {
        fclose(f);
        set_err("%s: bad header / magic", path);
        return UB_ERR_IO;
    }
    const UbConfig& c = t->cfg;
    if (header[2] != c.C_in || header[3] != c.C_model || header[4] != c.C_out || header[5] != c.H ||
        header[6] != c.W) {
        fclose(f);
        set_err("%s: checkpoint shape (C_in %d, C_model %d, C_out %d, %dx%d) does not match the trainer", path,
                header[2], header[3], header[4], header[5], header[6]);
        return UB_ERR_SHAPE;
    }
    std::vector<float> buf(t->nparams);
    auto read_arena = [&](float* dev) -> int {
        if (fread(buf.data(), sizeof(float), t->nparams, f) != t->nparams) return UB_ERR_IO;
        return cudaMemcpy(dev, buf.data(), t->nparams * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess
                   ? UB_OK
                   : UB_ERR_CUDA;
    };
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    int r = read_arena(t->params);
    if (!r && header[8] == 1) {  // AdamW moments (train_unet.cu:4874-4886)
        r = read_arena(t->m);
        if (!r) r = read_arena(t->v);
        // The reference writer (save_unet_states, train_unet.cu:4764) fills header words 0..9 of an UNINITIALISED
        // int[256]: words 10.. of a reference-written checkpoint are stack garbage.  The AdamW step count is therefore
        // trusted only behind this library's own marker word; otherwise the count restarts at 0, which is what the
        // reference's own resume does (its loop index restarts, train_unet.cu:5019-5037).  ub_trainer_set_step overrides.
        int step = 0;
        if (header[11] == kStepMagic && header[10] >= 0 && header[10] <= 1000000000) step = header[10];
        cudaMemcpy(t->step_dev, &step, sizeof(int), cudaMemcpyHostToDevice);
        t->host_step = step;
    }
    fclose(f);
    if (r) {
        set_err("%s: truncated checkpoint", path);
        return r;
    }
    // the moving average restarts from the loaded weights (ub_trainer_set_ema restores a saved one)
    if (t->ema) CUDA_TRY(cudaMemcpy(t->ema, t->params, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice));
    ensure_packed(t);
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    return UB_OK;
}

extern "C" int ub_trainer_set_step(UbTrainer* t, int step) {
    if (step < 0) {
        set_err("ub_trainer_set_step: negative step");
        return UB_ERR_STATE;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    CUDA_TRY(cudaMemcpy(t->step_dev, &step, sizeof(int), cudaMemcpyHostToDevice));
    t->host_step = step;
    return UB_OK;
}

static int save_checkpoint(UbTrainer* t, const char* path, const float* weights, int with_adamw) {
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaStreamSynchronize(t->stream));
    FILE* f = fopen(path, "wb");
    if (!f) {
        set_err("cannot open %s for writing", path);
        return UB_ERR_IO;
    }
    const UbConfig& c = t->cfg;
    int header[256];
    memset(header, 0, sizeof header);
    header[0] = kModelMagic, header[1] = c.B, header[2] = c.C_in, header[3] = c.C_model, header[4] = c.C_out;
    header[5] = c.H, header[6] = c.W, header[7] = c.max_period, header[8] = with_adamw ? 1 : 0, header[9] = 0;
    int step = 0;
    cudaMemcpy(&step, t->step_dev, sizeof(int), cudaMemcpyDeviceToHost);
    header[10] = step, header[11] = kStepMagic;
    std::vector<float> buf(t->nparams);
    bool ok = fwrite(header, sizeof(int), 256, f) == 256;
    auto write_arena = [&](const float* dev) {
        ok = ok && cudaMemcpy(buf.data(), dev, t->nparams * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
        ok = ok && fwrite(buf.data(), sizeof(float), t->nparams, f) == t->nparams;
    };
    write_arena(weights);
    if (with_adamw) write_arena(t->m), write_arena(t->v);
    fclose(f);
    if (!ok) {
        set_err("write to %s failed", path);
        return UB_ERR_IO;
    }
    return UB_OK;
}
extern "C" int ub_trainer_save(UbTrainer* t, const char* path, int with_adamw) {
    return save_checkpoint(t, path, t->params, with_adamw);
}
// The moving-average weights as a parameters-only checkpoint of the same layout: what generate.py (or
// ub_trainer_load + ub_trainer_sample) samples from in guided-diffusion practice.
extern "C" int ub_trainer_save_ema(UbTrainer* t, const char* path) {
    if (!t->ema) {
        set_err("no moving average: the trainer was created with cfg.ema_rate = 0");
        return UB_ERR_STATE;
    }
    return save_checkpoint(t, path, t->ema, 0);
}

// ======================================================================================================
// data parallel
// ======================================================================================================
extern "C" int ub_nccl_get_unique_id(void* id_out) {
    if (!nccl().ok) {
        set_err("libnccl.so.2 could not be loaded");
        return UB_ERR_NCCL;
    }
    ncclUniqueId id;
    int r = nccl().GetUniqueId(&id);
    if (r) {
        set_err("ncclGetUniqueId failed (%d)", r);
        return UB_ERR_NCCL;
    }
    memcpy(id_out, &id, sizeof id);
    return UB_OK;
}

extern "C" int ub_trainer_attach_dp(UbTrainer* t, int rank, int world, const void* nccl_id, int n_buckets) {
    if (!nccl().ok) {
        set_err("libnccl.so.2 could not be loaded");
        return UB_ERR_NCCL;
    }
    if (n_buckets != t->n_buckets && n_buckets > 0)
        fprintf(stderr, "[unet_b200] note: bucket count is fixed at build time (%d)\n", t->n_buckets);
    CUDA_TRY(cudaSetDevice(t->device));
    // The all-reduces run beside backward: every NCCL CTA takes an SM slot from the conv / wgrad kernels.  Measured on
    // 8 x B200 (ms per step, B = 32 per GPU): NCCL default 5.089, NCCL_MAX_CTAS = 16: 5.021, 8: 5.004, 4: 5.159
    // (profiles/r02_scaling.txt).  Only a default: an NCCL_MAX_CTAS already in the environment wins.
    setenv("NCCL_MAX_CTAS", "8", /*overwrite=*/0);
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof id);
    int r = nccl().CommInitRank(&t->comm, world, id, rank);
    if (r) {
        set_err("ncclCommInitRank failed (%d: %s)", r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
        return UB_ERR_NCCL;
    }
    t->rank = rank, t->world = world;
    // device-side timestep / noise draws are keyed by (seed, step): give every rank its own stream, otherwise ranks created
    // with the same seed draw identical t and correlated noise for different images
    t->cfg.seed ^= (unsigned long long)rank * 0x9E3779B97F4A7C15ULL;
    t->graph_valid = false;  // the step must be re-captured with the all-reduce nodes
    return UB_OK;
}
