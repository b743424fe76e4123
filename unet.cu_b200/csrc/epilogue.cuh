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
// Both reduce over the 32 pixels of a warp with a shuffle butterfly (31 SHFL per quantity per 32 columns) and add the
// warp totals of the CTA's 4 epilogue warps in shared memory; one red.global.add.v2.f32 per (tile, channel) follows.
#pragma once
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
};

__device__ __forceinline__ float epi_sigmoid(float z) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(0.5f, t, 0.5f);
}

// Column sums over the 32 lanes of a warp: on return lane j holds sum_lanes a[j] in a[0] (and likewise b[0]).
__device__ __forceinline__ void warp_colsum2(float (&a)[32], float (&b)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float sa = up ? a[i] : a[i + off], ka = up ? a[i + off] : a[i];
            const float sb = up ? b[i] : b[i + off], kb = up ? b[i + off] : b[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
            b[i] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
        }
    }
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
        tmem_ld32(taddr, v);
    } else {
        tmem_ld16(taddr, v);
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
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
        const float4 c = *reinterpret_cast<const float4*>(comb + j);
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

// 32-column chunk with one of the GroupNorm hooks (NHWC bf16 output).  Every lane of the warp takes part in the
// reduction; lanes whose pixel is outside the tensor contribute zeros and store nothing.  All pixels of a warp belong
// to one image (the plan guarantees it).  gconst = staged [4][BN] per-channel constants of image b (a = gamma*rstd,
// bb = beta - mean*a, rstd, mean*rstd), offset to this chunk's first column (gn-bwd only); cstride = BN.
__device__ __forceinline__ void epi_chunk_gn(const EpiOut& e, uint32_t taddr, const float* comb, const float* gconst,
                                             int cstride, bool valid, size_t pix, int n, int lane, float* red) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    uint4 xr[4];  // the residual row (stats mode) or the GroupNorm input row (gn-bwd mode): never both (plan)
    if ((e.residual || e.gx) && valid) {
        const uint4* xp = e.gx ? reinterpret_cast<const uint4*>(e.gx + pix * e.ldgx + n)
                               : reinterpret_cast<const uint4*>(e.residual + pix * e.ldr + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) xr[j] = xp[j];
    }
    tmem_ld_wait();
    float f[32], q[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 c = *reinterpret_cast<const float4*>(comb + j);
        f[j] = __uint_as_float(v[j]) + c.x, f[j + 1] = __uint_as_float(v[j + 1]) + c.y;
        f[j + 2] = __uint_as_float(v[j + 2]) + c.z, f[j + 3] = __uint_as_float(v[j + 3]) + c.w;
    }
    if (e.residual && !e.gx && valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t[8];
            unpack_bf16x8(xr[j], t);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[j * 8 + i] += t[i];
        }
    }
    if (e.gx) {
        // f = dL/d(act(gn(x)))  ->  dz = f * act'(z);  q = dz * xhat
        if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x[8];
                unpack_bf16x8(xr[j], x);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = j * 8 + i;
                    float dz = f[c];
                    if (e.gsilu) {
                        const float z = fmaf(x[i], gconst[c], gconst[cstride + c]);
                        const float s = epi_sigmoid(z);
                        dz *= s * fmaf(z, 1.f - s, 1.f);
                    }
                    const float xh = fmaf(x[i], gconst[2 * cstride + c], -gconst[3 * cstride + c]);
                    f[c] = dz;
                    q[c] = dz * xh;
                }
            }
        }
    }
    // round to the stored precision first: the statistics then describe exactly the tensor the consumer reads
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        if (!e.gx) {
            f[2 * i] = __bfloat162float(h2.x), f[2 * i + 1] = __bfloat162float(h2.y);
            q[2 * i] = f[2 * i] * f[2 * i], q[2 * i + 1] = f[2 * i + 1] * f[2 * i + 1];
        }
    }
    if (valid) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out) + pix * e.ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            reinterpret_cast<uint4*>(op)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f, q[j] = 0.f;
    }
    warp_colsum2(f, q, lane);
    // lane j now holds the warp's totals of column j: park them; epi_flush_stats adds the 4 warps up and issues
    // one vector RED per (tile, channel) instead of 8 scalar ones (the L2 atomic unit serialises per address)
    *reinterpret_cast<float2*>(red + 2 * lane) = make_float2(f[0], q[0]);
}

// red[4 warps][BN][2] -> global [B][Cout][2].  Called by the 128 epilogue threads after a barrier; img_of_warp(q) is
// the tile-local image index of warp q's 32 pixels.
__device__ __forceinline__ void epi_flush_stats(const float* red, float* dst, int Cout, int n0, int BN, int b0,
                                                int B, int rows_per_img, int TB, int tid) {
    for (int c = tid; c < BN; c += 128) {
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

// all columns [0, BN) of one accumulator row
__device__ __forceinline__ void epi_row(const EpiOut& e, uint32_t trow, const float* comb, int BN, bool valid,
                                        size_t pix, int b, int h, int w, int n0, const float* gconst = nullptr,
                                        int lane = 0, float* red = nullptr) {
    if (e.stats || e.gx) {  // BN % 32 == 0 (plan); red = this warp's [BN][2] row of the reduction scratch
        for (int c0 = 0; c0 < BN; c0 += 32)
            epi_chunk_gn(e, trow + uint32_t(c0), comb + c0, gconst + c0, BN, valid, pix, n0 + c0, lane, red + 2 * c0);
        return;
    }
    int c0 = 0;
    for (; c0 + 32 <= BN; c0 += 32) epi_chunk<32>(e, trow + uint32_t(c0), comb + c0, valid, pix, b, h, w, n0 + c0);
    for (; c0 < BN; c0 += 16) epi_chunk<16>(e, trow + uint32_t(c0), comb + c0, valid, pix, b, h, w, n0 + c0);
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

}  // namespace ub
