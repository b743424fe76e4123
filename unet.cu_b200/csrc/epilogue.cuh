// Shared epilogue of the implicit-GEMM kernels: one thread owns one output pixel (TMEM lane) and NC consecutive
// output channels.  The per-channel addend (bias + second bias + per-image embedding vector) is staged in shared
// memory by the caller BEFORE the accumulator is ready, and the residual row is fetched before the TMEM wait, so
// the only exposed latency per chunk is one tcgen05.wait::ld (the first version waited on dependent global loads per
// 16-column chunk, which made the epilogue -- not the MMA -- the critical path of small-K layers).
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
};

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
            const uint32_t rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
                f[j * 8 + i * 2] += __bfloat162float(h2.x);
                f[j * 8 + i * 2 + 1] += __bfloat162float(h2.y);
            }
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

// all columns [0, BN) of one accumulator row
__device__ __forceinline__ void epi_row(const EpiOut& e, uint32_t trow, const float* comb, int BN, bool valid,
                                        size_t pix, int b, int h, int w, int n0) {
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ub
