// Shared epilogue of the implicit-GEMM kernels: one thread owns one output pixel (TMEM lane) and NC consecutive
// output channels.  The per-channel addend (bias + second bias + per-image embedding vector) is staged in shared
// memory by the caller BEFORE the accumulator is ready, and the residual row is fetched before the TMEM wait, so
// the only exposed latency per chunk is one tcgen05.wait::ld (the first version waited on dependent global loads per
// 16-column chunk, which made the epilogue -- not the MMA -- the critical path of small-K layers).
//
// Two fused GroupNorm hooks (they remove one full pass over the tensor each, and ~100 launches per step):
//   * stats   : the epilogue of the conv that PRODUCES a tensor accumulates its per-(image, channel) sum and sum of
//               squares, which is all GroupNorm's statistics pass needs (replaces gn_stats).
//   * gn-bwd  : the epilogue of the dgrad conv that produces dL/d(silu(gn(x))) multiplies by silu'(gn(x)) on the
//               spot, writes dz = dL/d(gn(x)) instead, and accumulates the per-(image, channel) sums of dz and
//               dz * xhat (replaces gn_bwd_stats; gn_bwd_apply then needs no sigmoid at all).
// Both reduce over the 32 pixels of a warp through a small smem transpose and add the warp totals of the CTA's
// epilogue warps in shared memory; one red.global.add.v2.f32 per (tile, channel) follows.
#pragma once
#ifndef UB_EPI_SHUFFLE_REDUCE
#define UB_EPI_SHUFFLE_REDUCE 1
#endif
#include "igemm.cuh"
#include "ptx.cuh"

namespace ub {

struct EpiOut {
    const __nv_bfloat16* residual;
    int ldr;
    void* out;
    int ldo;
    int out_mode;
    int Cout, H, W;
    // fused GroupNorm hooks (NHWC bf16 output only)
    float* stats = nullptr;              // [B][Cout][2] += (sum, sumsq) of the stored (bf16-rounded) output
    const __nv_bfloat16* gx = nullptr;   // gn-bwd: the GroupNorm input x, NHWC bf16 [.., ldgx]
    int ldgx = 0;
    float* gS = nullptr;                 // gn-bwd: [B][Cout][2] += (sum dz, sum dz*xhat)
    int gsilu = 0;
    // two MMA streams (igemm_conv2_kernel): the result is the sum of two accumulators acc_stride TMEM columns apart
    int nacc = 1;
    uint32_t acc_stride = 0;
};

// v[0..16) = the 16 accumulator columns at taddr, summed over the kernel's accumulators
__device__ __forceinline__ void epi_tmem_load16(const EpiOut& e, uint32_t taddr, uint32_t (&v)[16]) {
    tmem_ld16(taddr, v);
    if (e.nacc == 2) {
        uint32_t u[16];
        tmem_ld16(taddr + e.acc_stride, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
    }
}

__device__ __forceinline__ float epi_sigmoid(float z) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(0.5f, t, 0.5f);
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float* f) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
        f[2 * i] = __bfloat162float(h2.x);
        f[2 * i + 1] = __bfloat162float(h2.y);
    }
}

// NC = 16 or 32 columns starting at absolute channel n, TMEM address taddr; comb points at the staged addend of
// channel n (shared memory).
template <int NC>
__device__ __forceinline__ void epi_chunk(const EpiOut& e, uint32_t taddr, const float* comb, bool valid, size_t pix,
                                          int b, int h, int w, int n) {
    uint32_t v[NC];
    if constexpr (NC == 32) {
        tmem_ld32(taddr, v);  // (one accumulator only: epi_row)
    } else {
        epi_tmem_load16(e, taddr, v);
    }
    uint4 r[NC / 8];
    if (e.residual && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(e.residual + pix * e.ldr + n);
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) r[j] = rp[j];
    }
    tmem_ld_wait();
    if (!valid) return;
    float f[NC];
    const uint32_t comb_s = smem_u32(comb);
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
        const float4 c = lds_v4(comb_s + 4u * j);
        f[j] = __uint_as_float(v[j]) + c.x, f[j + 1] = __uint_as_float(v[j + 1]) + c.y;
        f[j + 2] = __uint_as_float(v[j + 2]) + c.z, f[j + 3] = __uint_as_float(v[j + 3]) + c.w;
    }
    if (e.residual) {
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            float t[8];
            unpack_bf16x8(r[j], t);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[j * 8 + i] += t[i];
        }
    }
    if (e.out_mode == OUT_NHWC_BF16) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out) + pix * e.ldo + n;
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j * 8 + 2 * i], f[j * 8 + 2 * i + 1]);
                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            reinterpret_cast<uint4*>(op)[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    } else if (e.out_mode == OUT_NHWC_F32) {
        float* op = reinterpret_cast<float*>(e.out) + pix * e.ldo + n;
#pragma unroll
        for (int j = 0; j < NC; j += 4)
            *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {  // OUT_NCHW_F32: a warp writes 32 consecutive pixels of one channel -> coalesced
        float* op = reinterpret_cast<float*>(e.out) + ((size_t(b) * e.Cout + n) * e.H + h) * e.W + w;
        const size_t cs = size_t(e.H) * e.W;
#pragma unroll
        for (int j = 0; j < NC; ++j) op[j * cs] = f[j];
    }
}

// 16-column chunk with one of the GroupNorm hooks (NHWC bf16 output).  Every lane of the warp takes part in the
// reduction; lanes whose pixel is outside the tensor contribute zeros and store nothing.  All pixels of a warp belong
// to one image (the plan guarantees it).  gconst = staged [4][BN] per-channel constants of that image (a = gamma*rstd,
// bb = beta - mean*a, rstd, mean*rstd), offset to this chunk's first column (gn-bwd only); cstride = BN.
// Column sums over the warp's 32 pixels go through a per-warp smem transpose tr[32][36] (f in columns 0..15, q in
// 16..31; 144-byte rows keep both the 16-byte row stores and the column reads bank-conflict free): 8 STS.128 + 32
// independent LDS per lane instead of a 5-stage shuffle butterfly -- the epilogue is latency-bound, not issue-bound.
// xr holds the 16 channels [n, n+16) of the pixel's side input -- the residual row (stats mode) or the GroupNorm input
// row (gn-bwd mode), never both (plan) -- fetched one chunk AHEAD (epi_side_load): a dependent global load per chunk
// was the longest stall of the hooked epilogue.  On return xr holds the chunk at channel n_next (if n_next >= 0).
__device__ __forceinline__ void epi_side_load(const EpiOut& e, bool valid, size_t pix, int n, uint4 (&xr)[2]) {
    if ((e.residual || e.gx) && valid) {
        const uint4* xp = e.gx ? reinterpret_cast<const uint4*>(e.gx + pix * e.ldgx + n)
                               : reinterpret_cast<const uint4*>(e.residual + pix * e.ldr + n);
        xr[0] = xp[0], xr[1] = xp[1];
    }
}
__device__ __forceinline__ void epi_chunk_gn(const EpiOut& e, uint32_t taddr, const float* comb, const float* gconst,
                                             int cstride, bool valid, size_t pix, int n, int lane, float* red,
                                             float* tr, uint4 (&xr)[2], int n_next, uint32_t* keep_pk = nullptr) {
    uint32_t v[16];
    epi_tmem_load16(e, taddr, v);
    uint4 xn[2];
    if (n_next >= 0) epi_side_load(e, valid, pix, n_next, xn);
    tmem_ld_wait();
    float f[16], q[16];
    const uint32_t comb_s = smem_u32(comb);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 c = lds_v4(comb_s + 4u * j);
        f[j] = __uint_as_float(v[j]) + c.x, f[j + 1] = __uint_as_float(v[j + 1]) + c.y;
        f[j + 2] = __uint_as_float(v[j + 2]) + c.z, f[j + 3] = __uint_as_float(v[j + 3]) + c.w;
    }
    if (valid) {
        if (e.gx) {
            // f = dL/d(act(gn(x)))  ->  dz = f * act'(z);  q = dz * xhat
            const uint32_t gc_s = smem_u32(gconst), gstride = 4u * uint32_t(cstride);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float x[8];
                unpack_bf16x8(xr[j], x);
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {  // four channels at a time: one 16-byte read per constant row
                    const uint32_t ga = gc_s + 4u * uint32_t(j * 8 + h4 * 4);
                    const float4 ca = lds_v4(ga), cb = lds_v4(ga + gstride);
                    const float4 cr = lds_v4(ga + 2u * gstride), cm = lds_v4(ga + 3u * gstride);
                    const float a4[4] = {ca.x, ca.y, ca.z, ca.w}, b4[4] = {cb.x, cb.y, cb.z, cb.w};
                    const float r4[4] = {cr.x, cr.y, cr.z, cr.w}, m4[4] = {cm.x, cm.y, cm.z, cm.w};
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const int i = h4 * 4 + i4, c = j * 8 + i;
                        float dz = f[c];
                        if (e.gsilu) {
                            const float z = fmaf(x[i], a4[i4], b4[i4]);
                            const float sg = epi_sigmoid(z);
                            dz *= sg * fmaf(z, 1.f - sg, 1.f);
                        }
                        // the stored (bf16) dz is what gn_bwd_apply will read: sum exactly that
                        dz = __bfloat162float(__float2bfloat16(dz));
                        f[c] = dz;
                        q[c] = dz * fmaf(x[i], r4[i4], -m4[i4]);
                    }
                }
            }
        } else {
            if (e.residual) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float t[8];
                    unpack_bf16x8(xr[j], t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[j * 8 + i] += t[i];
                }
            }
            // round to the stored precision first: the statistics then describe exactly the tensor the consumer reads
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                f[c] = __bfloat162float(__float2bfloat16(f[c]));
                q[c] = f[c] * f[c];
            }
        }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out) + pix * e.ldo + n;
        reinterpret_cast<uint4*>(op)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(op)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (keep_pk) {  // (GroupNorm finish: the stored values stay in registers for pass 2)
#pragma unroll
            for (int i = 0; i < 8; ++i) keep_pk[i] = pk[i];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = 0.f, q[j] = 0.f;
    }
#if UB_EPI_SHUFFLE_REDUCE
    // Column sums over the warp's 32 pixels by a transpose-reduce butterfly: 31 shuffles move half as many bytes
    // through the SM's shuffle/shared-memory crossbar as the 8 STS.128 + 32 LDS of the shared-memory transpose
    // (64 wavefronts per chunk and warp; with 16 epilogue warps per SM that pipe is what bounds the hooked epilogue).
    float cs[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) cs[j] = f[j], cs[16 + j] = q[j];
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
        const bool up = (lane & sft) != 0;
#pragma unroll
        for (int i = 0; i < sft; ++i) {
            const float keep = up ? cs[i + sft] : cs[i];
            const float send = up ? cs[i] : cs[i + sft];
            cs[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
        }
    }
    // lane L holds the warp's total of cs[L]: f column L for L < 16, q column L - 16 above
    sts_f32(smem_u32(red) + 4u * uint32_t(2 * (lane & 15) + (lane >> 4)), cs[0]);
#else
    const uint32_t tr_s = smem_u32(tr);
    const uint32_t row_s = tr_s + uint32_t(lane) * 144u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sts_v4(row_s + 16u * j, f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        sts_v4(row_s + 64u + 16u * j, q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
    }
    __syncwarp();
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const uint32_t col_s = tr_s + 4u * uint32_t(lane);
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
        s0 += lds_f32(col_s + 144u * r), s1 += lds_f32(col_s + 144u * (r + 1));
        s2 += lds_f32(col_s + 144u * (r + 2)), s3 += lds_f32(col_s + 144u * (r + 3));
    }
    __syncwarp();
    // lane j < 16 holds the warp's total of f column j, lane 16 + j that of q column j: park them; epi_flush_stats
    // adds the warps up and issues one vector RED per (tile, channel) (the L2 atomic unit serialises per address)
    sts_f32(smem_u32(red) + 4u * uint32_t(2 * (lane & 15) + (lane >> 4)), (s0 + s1) + (s2 + s3));
#endif
    xr[0] = xn[0], xr[1] = xn[1];
}

// red[4 quadrants][BN][2] -> global [B][Cout][2].  Called by the nthreads epilogue threads after a barrier; quadrant
// q's 32 pixels belong to tile-local image min(32 q / rows_per_img, TB - 1).
__device__ __forceinline__ void epi_flush_stats(const float* red, float* dst, int Cout, int n0, int BN, int b0,
                                                int B, int rows_per_img, int TB, int tid, int nthreads) {
    for (int c = tid; c < BN; c += nthreads) {
        for (int img = 0; img < TB; ++img) {
            float s = 0.f, ss = 0.f;
            bool any = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int iq = min((q * 32) / rows_per_img, TB - 1);
                if (iq == img) {
                    const float2 v = *reinterpret_cast<const float2*>(red + (size_t(q) * BN + c) * 2);
                    s += v.x, ss += v.y, any = true;
                }
            }
            if (any && b0 + img < B) {
                float* d = dst + (size_t(b0 + img) * Cout + n0 + c) * 2;
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(d), "f"(s), "f"(ss) : "memory");
            }
        }
    }
}

// Columns [0, BN) of one accumulator row.  The row is shared by `nhalf` warps of the same TMEM lane quadrant: chunk i
// goes to warp i % nhalf.
__device__ __forceinline__ void epi_row(const EpiOut& e, uint32_t trow, const float* comb, int BN, bool valid,
                                        size_t pix, int b, int h, int w, int n0, const float* gconst = nullptr,
                                        int lane = 0, float* red = nullptr, float* tr = nullptr, int half = 0,
                                        int nhalf = 1, uint4* side = nullptr) {
    if (e.stats || e.gx) {  // BN % 32 == 0 (plan); red = this quadrant's [BN][2] row of the reduction scratch
        // side = the first chunk's side input, loaded by the caller (epi_side_load at channel n0 + 16 * half) before
        // it waited for the accumulator
        uint4 xr[2] = {side[0], side[1]};
        for (int c0 = 16 * half; c0 < BN; c0 += 16 * nhalf) {
            const int cn = c0 + 16 * nhalf;
            epi_chunk_gn(e, trow + uint32_t(c0), comb + c0, gconst + c0, BN, valid, pix, n0 + c0, lane, red + 2 * c0,
                         tr, xr, cn < BN ? n0 + cn : -1);
        }
        return;
    }
    int i = 0, c0 = 0;
    if (e.nacc == 1)  // (two accumulators are added per 16-column chunk)
        for (; c0 + 32 <= BN; c0 += 32, ++i)
            if (i % nhalf == half) epi_chunk<32>(e, trow + uint32_t(c0), comb + c0, valid, pix, b, h, w, n0 + c0);
    for (; c0 < BN; c0 += 16, ++i)
        if (i % nhalf == half) epi_chunk<16>(e, trow + uint32_t(c0), comb + c0, valid, pix, b, h, w, n0 + c0);
}

// comb[c] = bias[n0+c] + bias2[n0+c] + rowvec[b][n0+c] for c in [0, BN), computed by the calling threads (tid in
// [0, nthreads))
__device__ __forceinline__ void epi_stage_comb(float* comb, const float* bias, const float* bias2, const float* rowvec,
                                               int b, int Cout, int n0, int BN, int tid, int nthreads) {
    for (int c = tid; c < BN; c += nthreads) {
        float v = 0.f;
        if (bias) v += bias[n0 + c];
        if (bias2) v += bias2[n0 + c];
        if (rowvec) v += rowvec[size_t(b) * Cout + n0 + c];
        comb[c] = v;
    }
}

// gconst[0..3][BN] for image b, channels [n0, n0+BN): a = gamma*rstd, bb = beta - mean*a, rstd, mean*rstd, from the
// forward statistics chsum[B][C][2] (per-channel sum / sumsq over the HW pixels), groups of cpg channels.
__device__ __forceinline__ void epi_stage_gconst(float* gconst, const float* chsum, const float* gamma,
                                                 const float* beta, int b, int C, int cpg, int HW, int n0, int BN,
                                                 int tid, int nthreads) {
    for (int c = tid; c < BN; c += nthreads) {
        const int ch = n0 + c;
        const int g0 = (ch / cpg) * cpg;
        const float* cs = chsum + (size_t(b) * C + g0) * 2;
        float s = 0.f, ss = 0.f;
        for (int k = 0; k < cpg; ++k) s += cs[2 * k], ss += cs[2 * k + 1];
        const float n = float(cpg) * float(HW);
        const float mean = s / n;
        const float var = fmaxf(ss / n - mean * mean, 0.f);
        const float r = rsqrtf(var + 1e-5f);
        const float a = r * gamma[ch];
        gconst[c] = a;
        gconst[BN + c] = beta[ch] - mean * a;
        gconst[2 * BN + c] = r;
        gconst[3 * BN + c] = mean * r;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm finish (IgemmConvParams::gnf): the second half of a fused GroupNorm, run by the NTHREADS epilogue threads of
// a CTA whose pixel tile holds whole images (gnf_cluster == 1) or one half of an image whose other half belongs to the
// cluster peer (gnf_cluster == 2), after pass 1 (epi_row_keep) and a barrier.
//   red  [4][BN][2]      column sums of the four TMEM lane quadrants (pass 1)
//   tot  [2][TB][BN][2]  per-(image, channel) totals: [0] this CTA's, [1] stored here by the cluster peer (DSMEM)
//   gcf  [TB][3][BN]     per-(image, channel) constants of pass 2
//   gst  [2][BN]         gamma, beta of channels [n0, n0 + BN), staged while the main loop ran
//   xbar                 mbarrier (count BN) the peer's threads arrive on after their remote stores
// forward  (gnf == 1): totals = (sum y, sum y^2) -> a = act(y * A + Bb), A = gamma * rstd, Bb = beta - mean * A
// backward (gnf == 2): totals = (sum dz, sum dz * xhat) -> dx = A * dz - Bc * x + Cc (+ add), the constants of
//   gn_bwd_apply_dz_kernel (nhwc_ops.cu); dgamma / dbeta / the embedding column sums follow from the totals alone:
//   sum_p dx = A * sum dz - Bc * sum x + HW * Cc.
// Pass 2 works on the bf16 values pass 1 stored (kept in registers: keep = y or dz, xkeep = the GroupNorm input), so the
// result is bit-identical to the separate kernel reading those tensors; there is no global-memory round trip between
// the passes, only three CTA barriers (and the peer's arrivals).
template <int MAXCH>
__device__ __forceinline__ void epi_row_keep(const EpiOut& e, uint32_t trow, const float* comb, int BN, bool valid,
                                             size_t pix, int n0, const float* gconst, int lane, float* red, float* tr,
                                             int half, const uint4* side, uint32_t (&keep)[MAXCH][8],
                                             uint4 (&xkeep)[MAXCH][2]) {
    uint4 xr[2] = {side[0], side[1]};
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
        const int c0 = 16 * half + 32 * i;
        if (c0 < BN) {
            xkeep[i][0] = xr[0], xkeep[i][1] = xr[1];
            epi_chunk_gn(e, trow + uint32_t(c0), comb + c0, gconst + c0, BN, valid, pix, n0 + c0, lane, red + 2 * c0, tr,
                         xr, c0 + 32 < BN ? n0 + c0 + 32 : -1, keep[i]);
        }
    }
}

template <int NTHREADS, int MAXCH>
__device__ __forceinline__ void epi_gn_finish(const IgemmConvParams& p, const float* red, float* tot, float* gcf,
                                              const float* gst, const float* gconst, uint64_t* xbar, int n0, int b0,
                                              int et, int lbw, int half, bool valid, size_t pix,
                                              const uint32_t (&keep)[MAXCH][8], const uint4 (&xkeep)[MAXCH][2]) {
    const int BN = p.BN, TB = p.TB, rpi = p.TW * p.TH, cpg = p.gnf_cpg;
    const bool pair = p.gnf_cluster == 2;
    const uint32_t rank = pair ? cluster_ctarank() : 0u;
    float* tot_peer = tot + TB * BN * 2;
    // the residual-gradient rows of pass 2 travel while the totals and constants are formed
    uint4 va[MAXCH][2];
    if (p.gnf == 2 && p.gnf_add && valid) {
#pragma unroll
        for (int i = 0; i < MAXCH; ++i) {
            const int c0 = 16 * half + 32 * i;
            if (c0 < BN) {
                const uint4* ap = reinterpret_cast<const uint4*>(p.gnf_add + pix * p.gnf_ldadd + n0 + c0);
                va[i][0] = ap[0], va[i][1] = ap[1];
            }
        }
    }
    if (pair) cluster_wait();  // the peer has initialised its barrier (every thread of the pair arrived after its prologue)
    for (int idx = et; idx < TB * BN; idx += NTHREADS) {
        const int img = idx / BN, c = idx - img * BN;
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int iq = min((q * 32) / rpi, TB - 1);
            if (iq == img) {
                const float2 v = *reinterpret_cast<const float2*>(red + (size_t(q) * BN + c) * 2);
                s += v.x, ss += v.y;
            }
        }
        tot[idx * 2] = s, tot[idx * 2 + 1] = ss;
        if (pair) {
            st_cluster_v2(mapa_cluster(smem_u32(tot_peer + idx * 2), rank ^ 1u), s, ss);
            mbar_arrive_remote(mapa_cluster(smem_u32(xbar), rank ^ 1u));
        }
    }
    if (pair) mbar_wait_cluster(xbar, 0);
    named_bar_sync(1, NTHREADS);
    const float n = float(cpg) * float(p.H * p.W);
    for (int idx = et; idx < TB * BN; idx += NTHREADS) {
        const int img = idx / BN, c = idx - img * BN, ch = n0 + c;
        const int g0 = (c / cpg) * cpg;
        float* g3 = gcf + size_t(img) * 3 * BN;
        if (p.gnf == 1) {
            float s = 0.f, ss = 0.f;
            for (int k = 0; k < cpg; ++k) {
                const int i2 = (img * BN + g0 + k) * 2;
                s += tot[i2] + (pair ? tot_peer[i2] : 0.f);
                ss += tot[i2 + 1] + (pair ? tot_peer[i2 + 1] : 0.f);
            }
            const float mean = s / n;
            const float var = fmaxf(ss / n - mean * mean, 0.f);
            const float a = rsqrtf(var + 1e-5f) * gst[c];
            g3[c] = a;
            g3[BN + c] = gst[BN + c] - mean * a;
        } else {
            const float* gc = gconst + size_t(img) * 4 * BN;  // a, bb, rstd, mean * rstd of the forward pass
            float m1 = 0.f, m2 = 0.f;
            for (int k = 0; k < cpg; ++k) {
                const int i2 = (img * BN + g0 + k) * 2;
                const float gam = gst[g0 + k];
                m1 += gam * (tot[i2] + (pair ? tot_peer[i2] : 0.f));
                m2 += gam * (tot[i2 + 1] + (pair ? tot_peer[i2 + 1] : 0.f));
            }
            const float r = gc[2 * BN + c], mr = gc[3 * BN + c], A = gc[c];
            const float c1 = r * m1 / n, c2 = r * m2 / n;
            const float Bc = c2 * r, Cc = mr * c2 - c1;
            g3[c] = A, g3[BN + c] = Bc, g3[2 * BN + c] = Cc;
            if (b0 + img < p.B && rank == 0) {  // once per (image, channel)
                const int i2 = (img * BN + c) * 2;
                const float sd = tot[i2] + (pair ? tot_peer[i2] : 0.f);
                const float sq = tot[i2 + 1] + (pair ? tot_peer[i2 + 1] : 0.f);
                atomicAdd(p.gnf_dgamma + ch, sq);
                atomicAdd(p.gnf_dbeta + ch, sd);
                if (p.gnf_colsum) {
                    const size_t bc = size_t(b0 + img) * p.Cout + ch;
                    atomicAdd(p.gnf_colsum + bc, A * sd - Bc * p.gn_chsum[bc * 2] + float(p.H * p.W) * Cc);
                }
            }
        }
    }
    named_bar_sync(1, NTHREADS);
    if (!valid) return;
    const uint32_t g3s = smem_u32(gcf + size_t(lbw) * 3 * BN), row = 4u * uint32_t(BN);
    __nv_bfloat16* orow = p.gnf_out + pix * p.gnf_ldo + n0;
#pragma unroll
    for (int ci = 0; ci < MAXCH; ++ci) {
        const int c0 = 16 * half + 32 * ci;
        if (c0 < BN) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float y[8], x[8], ad[8], o[8];
                unpack_bf16x8(make_uint4(keep[ci][j * 4], keep[ci][j * 4 + 1], keep[ci][j * 4 + 2], keep[ci][j * 4 + 3]), y);
                if (p.gnf == 2) {
                    unpack_bf16x8(xkeep[ci][j], x);
                    if (p.gnf_add) unpack_bf16x8(va[ci][j], ad);
                }
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    const uint32_t ga = g3s + 4u * uint32_t(c0 + j * 8 + h4 * 4);
                    const float4 k0 = lds_v4(ga), k1 = lds_v4(ga + row);
                    const float a4[4] = {k0.x, k0.y, k0.z, k0.w}, b4[4] = {k1.x, k1.y, k1.z, k1.w};
                    if (p.gnf == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float z = fmaf(y[h4 * 4 + i], a4[i], b4[i]);
                            o[h4 * 4 + i] = p.gnf_silu ? z * epi_sigmoid(z) : z;
                        }
                    } else {
                        const float4 k2 = lds_v4(ga + 2u * row);
                        const float c4[4] = {k2.x, k2.y, k2.z, k2.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int e = h4 * 4 + i;
                            const float g = fmaf(a4[i], y[e], fmaf(-b4[i], x[e], c4[i]));
                            o[e] = p.gnf_add ? ad[e] + g : g;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                    pk[j * 4 + i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
            }
            reinterpret_cast<uint4*>(orow + c0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4*>(orow + c0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
    }
}

}  // namespace ub
