// Standalone microbenchmark: how fast can tcgen05.mma be issued / executed on sm_100a?
//
// Round 1 measured "145 cycles per M=128,K=16 instruction whatever N is" inside igemm_conv_kernel and concluded the
// issuing thread's stream was the bound.  This tool pins that finding without TMA, without a pipeline, without an
// epilogue: one CTA per SM (or two), operands resident in shared memory, a long stream of MMAs, cycles per MMA from
// clock64() around (first issue -> completion of the last MMA observed through tcgen05.commit on an mbarrier).
//
//   style 0: the issuing code runs in a lane-divergent region (`if (lane == 0)`), as every kernel of round 1 did.
//            ptxas then wraps EVERY tcgen05.mma / commit in a waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY).
//   style 1: the warp stays converged, only the instruction block is under elect.sync (what CUTLASS does):
//            descriptors stay in uniform registers and the UTCHMMAs are issued back to back.
//
// Sweeps M in {64,128}, N in {32..256}, cta_group 1 / 2, A from shared memory / TMEM, 1-4 issuer warps,
// same / rotating accumulators, commit after every K block (4 MMAs) or only at the end, 1 or 2 CTAs per SM.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/mma_issue_bench tools/mma_issue_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

struct BP {
    uint32_t M, N;      // MMA shape (M is the cta_group-wide M)
    int n_iter;         // K blocks; 4 MMAs (K = 16 each) per K block
    int issuers;        // issuing warps (each its own accumulator set)
    int rotate;         // 0: every MMA of an issuer accumulates into the same D; 1: K blocks rotate over 2 accumulators
    int commit_each;    // 1: tcgen05.commit after every K block (as a pipelined kernel does)
    int stages;         // smem ring the descriptors walk over (1..4)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n.reg .b32 %%rx;\n.reg .pred %%px;\n elect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\n mov.s32 %0, %%rx;\n}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFF));
    return pred;
}
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                     "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                     "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)3)
                     : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128_kmajor(uint32_t saddr) {
    // K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1), version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// smem layout per CTA: [stages][A 16 KiB | B 32 KiB] + barriers
constexpr int kStageBytes = 16384 + 32768;

template <int CG>
__device__ __forceinline__ void issue_block(uint32_t d, uint64_t dA, uint64_t dB, uint32_t a_tmem, bool atmem, uint32_t idesc,
                                            uint32_t acc_first) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (atmem)
            mma_ts<CG>(d, a_tmem + k * 8, dB + k * 2, idesc, k ? 1u : acc_first);
        else
            mma_ss<CG>(d, dA + k * 2, dB + k * 2, idesc, k ? 1u : acc_first);
    }
}

template <int STYLE, int CG, bool ATMEM>
__global__ void __launch_bounds__(192) mma_bench(const BP p, unsigned long long* out, int* err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * kStageBytes);  // [0]=done, [1]=sink
    uint32_t* tmem_slot = (uint32_t*)(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank = 0;
    if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    // fill operands with a harmless bf16 pattern (values do not matter for timing; avoid NaN/denormal)
    for (int i = threadIdx.x; i < p.stages * kStageBytes / 4; i += blockDim.x)
        ((uint32_t*)smem)[i] = 0x3C003C00u + ((i * 2654435761u) >> 28);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(p.issuers));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(0xFFFFF));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const bool issuer_cta = (CG == 1) || cta_rank == 0;
    unsigned long long t0 = 0, t1 = 0;
    if (warp >= 1 && warp <= p.issuers && issuer_cta) {
        const int w = warp - 1;
        const uint32_t idesc = make_idesc(p.M, p.N);
        // accumulators: issuer w owns columns [w * span, (w+1) * span), span = 512 / issuers (A-in-TMEM takes the last 32)
        const uint32_t span = (ATMEM ? 448u : 512u) / (uint32_t)p.issuers;
        const uint32_t d0 = tmem_base + w * span;
        const uint32_t d1 = (p.rotate && 2 * p.N <= span) ? d0 + p.N : d0;
        const uint32_t a_tmem = tmem_base + 480;
        const uint32_t s0 = smem_u32(smem);
        const uint32_t sink = smem_u32(&bars[1]), done = smem_u32(&bars[0]);
        t0 = clock64();
        if constexpr (STYLE == 0) {
            if (lane == 0) {
                int stage = 0;
                for (int it = 0; it < p.n_iter; ++it) {
                    const uint32_t sA = s0 + stage * kStageBytes;
                    const uint64_t dA = desc_sw128_kmajor(sA), dB = desc_sw128_kmajor(sA + 16384);
                    issue_block<CG>((it & 1) ? d1 : d0, dA, dB, a_tmem, ATMEM, idesc, it >= 2 ? 1u : 0u);
                    if (p.commit_each) commit<CG>(sink);
                    if (++stage == p.stages) stage = 0;
                }
                commit<CG>(done);
            }
            __syncwarp();
        } else {
            int stage = 0;
            for (int it = 0; it < p.n_iter; ++it) {
                const uint32_t sA = s0 + stage * kStageBytes;
                const uint64_t dA = desc_sw128_kmajor(sA), dB = desc_sw128_kmajor(sA + 16384);
                if (elect_one_sync()) {
                    issue_block<CG>((it & 1) ? d1 : d0, dA, dB, a_tmem, ATMEM, idesc, it >= 2 ? 1u : 0u);
                    if (p.commit_each) commit<CG>(sink);
                }
                __syncwarp();
                if (++stage == p.stages) stage = 0;
            }
            if (elect_one_sync()) commit<CG>(done);
            __syncwarp();
        }
        // wait for completion (bounded spin)
        uint32_t ok = 0;
        for (long spin = 0; spin < (1L << 26) && !ok; ++spin)
            asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0, 1, 0, P;\n}\n"
                         : "=r"(ok)
                         : "r"(done), "r"(0)
                         : "memory");
        t1 = clock64();
        if (!ok && lane == 0) atomicExch(err, 1);
        if (lane == 0) out[(size_t)blockIdx.x * 4 + w] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 0) {
        if constexpr (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

template <int STYLE, int CG, bool ATMEM>
static void run(const char* tag, BP p, int ctas_per_sm, int nsm) {
    static unsigned long long* d_out = nullptr;
    static int* d_err = nullptr;
    if (!d_out) {
        CK(cudaMalloc(&d_out, 4096 * 4 * sizeof(unsigned long long)));
        CK(cudaMalloc(&d_err, sizeof(int)));
    }
    CK(cudaMemset(d_out, 0, 4096 * 4 * sizeof(unsigned long long)));
    CK(cudaMemset(d_err, 0, sizeof(int)));
    // two CTAs per SM: each takes at most half the shared memory and half the TMEM -- TMEM alloc is 512 columns per CTA
    // here, so "2 per SM" is only possible with 256: keep 1 CTA/SM semantic simple and emulate co-residency through grid.
    const int smem_bytes = p.stages * kStageBytes + 1024 + 256;
    auto kern = mma_bench<STYLE, CG, ATMEM>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int grid = nsm * ctas_per_sm;
    if (CG == 2) grid &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(192), cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, kern, p, d_out, d_err));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h((size_t)grid * 4);
    int herr = 0;
    CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
    double sum = 0, mx = 0;
    int cnt = 0;
    for (int c = 0; c < grid; ++c)
        for (int w = 0; w < p.issuers; ++w) {
            if (CG == 2 && (c & 1)) continue;
            const double cyc = (double)h[(size_t)c * 4 + w] / (p.n_iter * 4.0);
            sum += cyc, mx = std::max(mx, cyc), ++cnt;
        }
    const double avg = sum / std::max(cnt, 1);
    // per-SM tensor floor for this instruction: M*N*16 MACs at 4096 MAC/clk/SM (cta_group::2 spreads M over two SMs)
    const double floor_cyc = (double)(p.M / CG) * p.N * 16.0 / 4096.0;
    const double per_sm_rate = p.issuers / avg;  // MMAs per cycle per issuing CTA
    printf("%-22s st%d cg%d %s M=%3u N=%3u iss=%d rot=%d cmt=%d stg=%d | %7.1f cyc/MMA/issuer (max %7.1f) | tensor floor %5.1f "
           "-> %5.1f %% of peak%s\n",
           tag, STYLE, CG, ATMEM ? "TS" : "SS", p.M, p.N, p.issuers, p.rotate, p.commit_each, p.stages, avg, mx, floor_cyc,
           100.0 * floor_cyc * per_sm_rate, herr ? "  [TIMEOUT]" : "");
    fflush(stdout);
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int nsm = prop.multiProcessorCount;
    printf("# %s, %d SMs, clock %.0f MHz\n", prop.name, nsm, prop.clockRate * 1e-3);
    const int n_iter = 512;
    const uint32_t Ns[] = {32, 64, 96, 128, 192, 256};
    // 1. the round-1 finding: style 0 vs style 1, one issuer, commit per K block, one CTA per SM
    for (uint32_t N : Ns) {
        BP p{128, N, n_iter, 1, 0, 1, 2};
        run<0, 1, false>("lane-divergent", p, 1, nsm);
        run<1, 1, false>("converged+elect", p, 1, nsm);
    }
    // 2. without per-block commits, single stage (pure issue/execute rate)
    for (uint32_t N : Ns) {
        BP p{128, N, n_iter, 1, 0, 0, 1};
        run<0, 1, false>("lane-div nocommit", p, 1, nsm);
        run<1, 1, false>("conv+elect nocommit", p, 1, nsm);
    }
    // 3. rotating accumulators (dependency on D?)
    for (uint32_t N : {32u, 64u, 128u}) {
        BP p{128, N, n_iter, 1, 1, 1, 2};
        run<1, 1, false>("rotate D", p, 1, nsm);
    }
    // 4. M = 64
    for (uint32_t N : Ns) {
        BP p{64, N, n_iter, 1, 0, 1, 2};
        run<1, 1, false>("M=64", p, 1, nsm);
    }
    // 5. several issuer warps in one CTA
    for (int iss : {2, 4})
        for (uint32_t N : {32u, 64u, 128u}) {
            BP p{128, N, n_iter, iss, 0, 1, 2};
            run<0, 1, false>("multi-issuer lane-div", p, 1, nsm);
            run<1, 1, false>("multi-issuer elect", p, 1, nsm);
        }
    // 6. A from TMEM
    for (uint32_t N : Ns) {
        BP p{128, N, n_iter, 1, 0, 1, 2};
        run<1, 1, true>("A in TMEM", p, 1, nsm);
    }
    // 7. one CTA on the whole chip vs all SMs busy (shared-memory / power effects)
    {
        BP p{128, 64, n_iter, 1, 0, 1, 2};
        run<1, 1, false>("single CTA", p, 1, 1);
    }
    if (argc > 1 && atoi(argv[1]) >= 2) {
        // 8. cta_group::2 (M = 256 over a CTA pair, N split between the two CTAs' shared memories)
        for (uint32_t N : Ns) {
            BP p{256, N, n_iter, 1, 0, 1, 2};
            run<1, 2, false>("cta_group::2 M=256", p, 1, nsm);
        }
        for (uint32_t N : {64u, 128u, 256u}) {
            BP p{128, N, n_iter, 1, 0, 1, 2};
            run<1, 2, false>("cta_group::2 M=128", p, 1, nsm);
        }
        for (uint32_t N : {64u, 128u, 256u}) {
            BP p{256, N, n_iter, 1, 0, 1, 2};
            run<1, 2, true>("cta_group::2 TS", p, 1, nsm);
        }
    }
    return 0;
}
