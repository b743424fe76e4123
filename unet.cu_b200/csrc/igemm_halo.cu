// igemm_halo_kernel -- persistent, halo-reuse tcgen05 implicit GEMM for 3x3 / 1x1 convolution fprop and dgrad on
// NHWC bf16 images with W >= 32.
//
// Why a second main loop: igemm_conv_kernel re-fetches the 128-pixel activation box for each of the 9 taps, so
// small-C layers at 64x64 / 32x32 are bound by L2->SMEM traffic (profiles/r01_*: 64->64@64x64 runs at 16 % of the
// bf16 peak).  Here the activation HALO of a tile is loaded once per 64-channel block and the 9 taps are just UMMA
// descriptor start-address offsets into it (the first GPU probe showed that any multiple of 128 B works with
// base_offset 0).  To make a tap shift a uniform smem offset the output tile runs over the image in PADDED
// row-major order (row pitch Wp = W + 2: one zero column on either side, filled by TMA out-of-bounds zero fill);
// outputs that fall on a pad column are computed and dropped (2/(W+2) of the MMA work).
//
//   tile      : M = 256 padded positions of one image (two UMMA M=128 accumulators) x N = BN <= 128 channels
//   A stage   : TMA box (64 ch, Wp, NR rows) = halo of the tile, NR = 3 + ceil(256 / Wp); 2 stages
//   W stage   : one (tap, 64-ch block) weight tile BN x 64, used by BOTH M halves; ring of 3-8 stages
//   TMEM      : 2 tile buffers x 2 halves x BN fp32 columns (<= 512): the epilogue of tile i overlaps tile i+1
//   warps     : 0 = A producer, 1 = W producer, 2 = MMA issuer (+TMEM alloc), 3 = idle, 4-7 = epilogue
//   grid      : persistent, min(#tiles, #SMs) CTAs, static round-robin tile schedule
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <cstring>

namespace ub {

static constexpr int kHaloThreads = 256;
static constexpr int kMaxWStages = 8;
static constexpr int kMarginBytes = 1024;  // keeps tap offsets of pad-column outputs inside the stage

#ifdef UB_HALO_TRACE
// development aid: per-CTA, per-tile clock stamps (8 x long long per tile) -> tools/igemm_test trace mode
__device__ long long* g_halo_trace = nullptr;
#define TRACE(slot, it_) \
    do { if (g_halo_trace && (it_) < 8) g_halo_trace[(size_t(blockIdx.x) * 8 + (it_)) * 8 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot, it_) do { } while (0)
#endif

struct HaloBars {
    uint64_t a_full[2], a_empty[2];
    uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
    uint64_t t_full[2], t_empty[2];
    uint32_t tmem_slot;
    uint32_t pad_[3];
    float comb[2][128];  // staged per-channel addend of the tile in flight (double buffered)
};

__global__ void __launch_bounds__(kHaloThreads, 1) igemm_halo_kernel(const __grid_constant__ IgemmHaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                      // 2 stages of a_stage_bytes
    uint8_t* sW = smem + 2 * size_t(p.a_stage_bytes);        // w_stages of w_stage_bytes
    HaloBars* bars = reinterpret_cast<HaloBars*>(sW + size_t(p.w_stages) * p.w_stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = p.Cout / p.BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nseg; ++s) {
            tma_prefetch_desc(&p.seg[s].tmA);
            tma_prefetch_desc(&p.seg[s].tmW);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], 1);
            mbar_init(&bars->a_empty[i], 1);
            mbar_init(&bars->t_full[i], 1);
            mbar_init(&bars->t_empty[i], 4);  // one arrive per epilogue warp
        }
        for (int i = 0; i < p.w_stages; ++i) {
            mbar_init(&bars->w_full[i], 1);
            mbar_init(&bars->w_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ A (halo) producer
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int mt = (tile / n_tiles_n) % p.tiles_per_img;
                const int b = tile / (n_tiles_n * p.tiles_per_img);
                const int hstart = (mt * 256) / p.Wp - 1;
                TRACE(0, (tile - int(blockIdx.x)) / int(gridDim.x));
                for (int s = 0; s < p.nseg; ++s) {
                    for (int cb = 0; cb < p.seg[s].cblocks; ++cb) {
                        mbar_wait(&bars->a_empty[st], ph ^ 1);
                        mbar_expect_tx(&bars->a_full[st], p.a_bytes);
                        tma_load_4d(sA + size_t(st) * p.a_stage_bytes + kMarginBytes, &p.seg[s].tmA, &bars->a_full[st],
                                    cb * 64, -1, hstart, b);
                        if (++st == 2) st = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ W (weight tile) producer
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n0 = (tile % n_tiles_n) * p.BN;
                for (int s = 0; s < p.nseg; ++s) {
                    for (int cb = 0; cb < p.seg[s].cblocks; ++cb) {
                        for (int tap = 0; tap < p.seg[s].ntaps; ++tap) {
                            mbar_wait(&bars->w_empty[st], ph ^ 1);
                            mbar_expect_tx(&bars->w_full[st], p.w_bytes);
                            tma_load_2d(sW + size_t(st) * p.w_stage_bytes, &p.seg[s].tmW, &bars->w_full[st], cb * 64,
                                        tap * p.Cout + n0);
                            if (++st == p.w_stages) st = 0, ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int mt = (tile / n_tiles_n) % p.tiles_per_img;
                const int buf = it & 1;
                const uint32_t off_tile = uint32_t((mt * 256) % p.Wp + p.Wp);  // first output position in the halo
                mbar_wait(&bars->t_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                TRACE(1, it);
                const uint32_t d0 = tmem_base + uint32_t(buf * 2 * p.BN);
                bool first = true;
                for (int s = 0; s < p.nseg; ++s) {
                    const int ntaps = p.seg[s].ntaps;
                    for (int cb = 0; cb < p.seg[s].cblocks; ++cb) {
                        mbar_wait(&bars->a_full[sa], pa);
                        tc_fence_after();
                        TRACE(2, it);
                        const uint32_t a_base = smem_u32(sA + size_t(sa) * p.a_stage_bytes + kMarginBytes);
                        for (int tap = 0; tap < ntaps; ++tap) {
                            const int dy = ntaps == 9 ? tap / 3 - 1 : 0;
                            const int dx = ntaps == 9 ? tap % 3 - 1 : 0;
                            mbar_wait(&bars->w_full[sw], pw);
                            tc_fence_after();
                            const uint32_t a0 = a_base + uint32_t(int(off_tile) + dy * p.Wp + dx) * 128u;
                            const uint64_t dB = make_smem_desc_sw128(smem_u32(sW + size_t(sw) * p.w_stage_bytes), 16, 1024);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const uint64_t dA = make_smem_desc_sw128(a0 + uint32_t(half) * 128u * 128u, 16, 1024);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16(d0 + uint32_t(half * p.BN), dA + uint64_t(k * 2), dB + uint64_t(k * 2),
                                              idesc, (!first || k != 0) ? 1u : 0u);
                            }
                            first = false;
                            umma_commit(&bars->w_empty[sw]);
                            if (++sw == p.w_stages) sw = 0, pw ^= 1;
                        }
                        umma_commit(&bars->a_empty[sa]);
                        if (++sa == 2) sa = 0, pa ^= 1;
                    }
                }
                umma_commit(&bars->t_full[buf]);
                TRACE(3, it);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;
        const EpiOut eo{p.residual, p.ldr, p.out, p.ldo, p.out_mode, p.Cout, p.H, p.W};
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int n0 = (tile % n_tiles_n) * p.BN;
            const int mt = (tile / n_tiles_n) % p.tiles_per_img;
            const int b = tile / (n_tiles_n * p.tiles_per_img);
            const int buf = it & 1;
            // stage bias + bias2 + embedding vector of this tile while its MMAs are still running
            epi_stage_comb(bars->comb[buf], p.bias, p.bias2, p.rowvec, b, p.Cout, n0, p.BN, threadIdx.x - 128, 128);
            named_bar_sync(1, 128);
            if (threadIdx.x == 128) TRACE(4, it);
            mbar_wait(&bars->t_full[buf], (it >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 128) TRACE(5, it);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int pos = mt * 256 + half * 128 + q * 32 + lane;  // padded position inside the image
                const int h = pos / p.Wp, wp = pos % p.Wp;
                const bool valid = h < p.H && wp >= 1 && wp <= p.W;
                const int w = wp - 1;
                const size_t pix = (size_t(b) * p.H + h) * p.W + w;
                const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * 2 * p.BN + half * p.BN);
                epi_row(eo, trow, bars->comb[buf], p.BN, valid, pix, b, h, w, n0);
            }
            // all TMEM reads of this buffer are complete (tcgen05.wait::ld above): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (threadIdx.x == 128) TRACE(6, it);
            if (lane == 0) mbar_arrive(&bars->t_empty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn igemm_encode_fn();  // igemm.cu

static int halo_act_map(CUtensorMap* m, const __nv_bfloat16* x, int C, int ld, int W, int H, int B, int Wp, int NR) {
    EncodeTiledFn fn = igemm_encode_fn();
    if (!fn) return -10;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (ld % 8) != 0) return -11;
    cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(B)};
    cuuint64_t strides[3] = {cuuint64_t(ld) * 2, cuuint64_t(W) * ld * 2, cuuint64_t(H) * W * ld * 2};
    cuuint32_t box[4] = {64, cuuint32_t(Wp), cuuint32_t(NR), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -12;
}
static int halo_weight_map(CUtensorMap* m, const __nv_bfloat16* wp, int Cin, int rows, int BN) {
    EncodeTiledFn fn = igemm_encode_fn();
    if (!fn) return -10;
    if ((reinterpret_cast<uintptr_t>(wp) & 15) || (Cin % 8) != 0) return -13;
    cuuint64_t dims[2] = {cuuint64_t(Cin), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(Cin) * 2};
    cuuint32_t box[2] = {64, cuuint32_t(BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wp), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -14;
}

bool igemm_halo_eligible(int B, int H, int W, int Cout) {
    (void)B;
    (void)H;
    if (W < 32 || W + 2 > 256) return false;
    if (Cout % 16 != 0) return false;
    return true;
}

int igemm_halo_plan(IgemmHaloParams* p, const ConvSegDesc* segs, int nseg, int B, int H, int W, int Cout,
                    const ConvEpilogue& ep, int sm_count) {
    memset(p, 0, sizeof(*p));
    if (nseg < 1 || nseg > 2 || !igemm_halo_eligible(B, H, W, Cout)) return -1;
    int BN = 0;
    for (int cand = 128; cand >= 16; cand -= 16)
        if (Cout % cand == 0) {
            BN = cand;
            break;
        }
    if (!BN) return -2;
    p->nseg = nseg, p->B = B, p->H = H, p->W = W, p->Cout = Cout, p->BN = BN;
    p->Wp = W + 2;
    p->NR = 3 + (256 + p->Wp - 1) / p->Wp;
    if (p->NR > 256) return -3;
    p->tiles_per_img = (H * p->Wp + 255) / 256;
    p->num_tiles = B * p->tiles_per_img * (Cout / BN);
    p->a_bytes = uint32_t(64 * p->Wp * p->NR * 2);
    p->a_stage_bytes = ((p->a_bytes + 2u * kMarginBytes) + 1023u) & ~1023u;
    p->w_bytes = uint32_t(64 * BN * 2);
    p->w_stage_bytes = (p->w_bytes + 1023u) & ~1023u;
    const size_t budget = size_t(227) * 1024 - 1024 - sizeof(HaloBars) - 2 * size_t(p->a_stage_bytes);
    int ws = int(budget / p->w_stage_bytes);
    if (ws > kMaxWStages) ws = kMaxWStages;
    if (ws < 2) return -3;
    p->w_stages = ws;
    for (int s = 0; s < nseg; ++s) {
        const ConvSegDesc& d = segs[s];
        if ((d.ntaps != 9 && d.ntaps != 1) || d.Cin % 8 != 0) return -4;
        int r = halo_act_map(&p->seg[s].tmA, d.x, d.Cin, d.ldx, W, H, B, p->Wp, p->NR);
        if (r) return r;
        r = halo_weight_map(&p->seg[s].tmW, d.wp, d.Cin, d.ntaps * Cout, BN);
        if (r) return r;
        p->seg[s].cblocks = (d.Cin + 63) / 64;
        p->seg[s].ntaps = d.ntaps;
    }
    p->bias = ep.bias, p->bias2 = ep.bias2, p->rowvec = ep.rowvec, p->residual = ep.residual;
    p->ldr = ep.ldr ? ep.ldr : Cout;
    p->out = ep.out;
    p->ldo = ep.ldo ? ep.ldo : Cout;
    p->out_mode = ep.out_mode;
    if (p->out_mode != OUT_NCHW_F32) {
        const int esz = p->out_mode == OUT_NHWC_BF16 ? 2 : 4;
        if ((p->ldo * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(p->out) & 15)) return -6;
    }
    if (p->residual && ((p->ldr % 8) != 0 || (reinterpret_cast<uintptr_t>(p->residual) & 15))) return -7;
    if ((p->bias && (reinterpret_cast<uintptr_t>(p->bias) & 15)) || (p->bias2 && (reinterpret_cast<uintptr_t>(p->bias2) & 15)) ||
        (p->rowvec && (reinterpret_cast<uintptr_t>(p->rowvec) & 15)))
        return -8;
    p->grid = p->num_tiles < sm_count ? p->num_tiles : sm_count;
    return 0;
}

void igemm_halo_init() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(227 * 1024));
    done = true;
}

#ifdef UB_HALO_TRACE
void igemm_halo_set_trace(long long* dev_buf) { cudaMemcpyToSymbol(g_halo_trace, &dev_buf, sizeof(dev_buf)); }
#endif

int igemm_halo_launch(const IgemmHaloParams& p, cudaStream_t st) {
    igemm_halo_init();
    const size_t smem = 2 * size_t(p.a_stage_bytes) + size_t(p.w_stages) * p.w_stage_bytes + sizeof(HaloBars) + 1024;
    igemm_halo_kernel<<<p.grid, kHaloThreads, smem, st>>>(p);
    return int(cudaGetLastError());
}

}  // namespace ub
