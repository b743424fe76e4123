// Standalone microbenchmark: how many bytes per cycle can one SM pull through TMA (cp.async.bulk.tensor 2D, 128-byte
// swizzled rows of 64 bf16 -- the operand box shape of every tcgen05 kernel here), as a function of the bytes kept in
// flight (stages x stage size), of the working set (L2 resident vs DRAM) and of the number of SMs pulling?
// The conv kernels' operand stream is bounded by this, not by the tensor pipe, for N <= 128 tiles
// (profiles/r02_tma_ingest_bench.txt).
//
// One CTA per SM: warp 0 = producer (waits `empty`, expect_tx, TMA), warp 1 = consumer (waits `full`, arrives `empty`
// immediately: no MMA, no smem reads).  Box = 64 elements x `rows` rows; a stage is `boxes` boxes.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_ingest_bench tools/tma_ingest_bench.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n.reg .b32 %%rx;\n.reg .pred %%px;\n elect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\n mov.s32 %0, %%rx;\n}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFF));
    return pred;
}
// poll = 1: mbarrier.test_wait in a spin loop (never suspends the thread) instead of try_wait (may suspend for a
// system-dependent time before it re-checks)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int poll = 0) {
    uint32_t ok = 0;
    if (poll) {
        while (!ok)
            asm volatile("{\n.reg .pred P;\nmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0, 1, 0, P;\n}\n"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
        return;
    }
    while (!ok)
        asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0, 1, 0, P;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
}

struct IP {
    int stages, boxes, rows;  // stage = boxes x (rows x 128 B)
    int iters;                // stages pulled per CTA
    int total_rows;           // rows of the global tensor (working set = total_rows x 128 B)
    int poll;                 // 1 = test_wait spin loops
    int nprod;                // producer warps: warp w issues the stages it = w (mod nprod) -- is one thread's issue rate the limit?
};

__global__ void __launch_bounds__(160) ingest(const __grid_constant__ CUtensorMap tm, const IP p, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int box_bytes = p.rows * 128, stage_bytes = p.boxes * box_bytes;
    uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + 16;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (warp >= 1 && warp <= p.nprod && p.nprod > 1) {
        // several producers: producer w takes iterations w-1, w-1+nprod, ... (warp 0 is the consumer in this mode)
        const int w = warp - 1;
        const uint32_t mask = (uint32_t)p.total_rows - 1u;
        for (int it = w; it < p.iters; it += p.nprod) {
            const int st = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(smem_u32(&empty[st]), ph ^ 1, p.poll);
            if (elect_one_sync()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(stage_bytes)
                             : "memory");
                const uint32_t row = (blockIdx.x + (uint32_t)it * gridDim.x) * p.boxes * p.rows;
                for (int b = 0; b < p.boxes; ++b) {
                    const int r = (int)((row + (uint32_t)b * p.rows) & mask);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                            smem_u32(smem + (size_t)st * stage_bytes + (size_t)b * box_bytes)),
                        "l"(&tm), "r"(smem_u32(&full[st])), "r"(0), "r"(r)
                        : "memory");
                }
            }
            __syncwarp();
        }
    } else if (warp == 0 && p.nprod > 1) {
        int st = 0;
        uint32_t ph = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(smem_u32(&full[st]), ph, p.poll);
            if (elect_one_sync()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
            __syncwarp();
            if (++st == p.stages) st = 0, ph ^= 1;
        }
        if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    } else if (p.nprod > 1) {
    } else if (warp == 0) {
        int st = 0;
        uint32_t ph = 0;
        // every CTA walks the tensor from its own offset (distinct rows per CTA per iteration, wrapping around)
        // (32-bit, power-of-two wrap: a 64-bit modulo here cost ~220 cycles per box and WAS the measured "limit" in the
        //  first version of this tool)
        uint32_t row = blockIdx.x * p.boxes * p.rows;
        const uint32_t mask = (uint32_t)p.total_rows - 1u;  // total_rows is a power of two, rows divides it
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(smem_u32(&empty[st]), ph ^ 1, p.poll);
            if (elect_one_sync()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(stage_bytes)
                             : "memory");
                for (int b = 0; b < p.boxes; ++b) {
                    const int r = (int)((row + (uint32_t)b * p.rows) & mask);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                            smem_u32(smem + (size_t)st * stage_bytes + (size_t)b * box_bytes)),
                        "l"(&tm), "r"(smem_u32(&full[st])), "r"(0), "r"(r)
                        : "memory");
                }
            }
            __syncwarp();
            row += gridDim.x * p.boxes * p.rows;
            if (++st == p.stages) st = 0, ph ^= 1;
        }
    } else if (warp == 1) {
        int st = 0;
        uint32_t ph = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(smem_u32(&full[st]), ph, p.poll);
            if (elect_one_sync()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
            __syncwarp();
            if (++st == p.stages) st = 0, ph ^= 1;
        }
        if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("# %s, %d SMs\n", prop.name, nsm);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    const size_t big_rows = (size_t)1 << 24;  // 2 GiB of 128-byte rows
    void* buf;
    CK(cudaMalloc(&buf, big_rows * 128));
    CK(cudaMemset(buf, 0, big_rows * 128));
    unsigned long long* d_out;
    CK(cudaMalloc(&d_out, 1024 * 8));
    CK(cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));

    struct Cfg { int stages, boxes, rows; };
    const Cfg cfgs[] = {{2, 1, 128}, {4, 1, 128}, {6, 1, 128}, {8, 1, 128}, {12, 1, 128}, {3, 3, 128}, {4, 3, 128},
                        {4, 1, 256}, {6, 1, 256}, {3, 2, 256}, {8, 1, 64}, {16, 1, 64}, {2, 3, 256}, {6, 2, 128}, {4, 2, 128}};
    const int poll = getenv("POLL") ? atoi(getenv("POLL")) : 0;
    const int only_l2 = getenv("ONLY_L2") ? 1 : 0;
    const int nprod = getenv("NPROD") ? atoi(getenv("NPROD")) : 1;
    printf("# wait = %s, producer warps = %d\n", poll ? "mbarrier.test_wait spin" : "mbarrier.try_wait", nprod);
    for (int ws = 0; ws < (only_l2 ? 1 : 2); ++ws) {
        const size_t rows_total = ws == 0 ? ((size_t)32 << 20) / 128 : big_rows;  // 32 MiB (L2 resident) or 2 GiB (DRAM)
        for (int grid : {nsm, 64, 16}) {
            for (const Cfg& c : cfgs) {
                IP p{c.stages, c.boxes, c.rows, 0, (int)rows_total, poll, nprod};
                const size_t stage_bytes = (size_t)c.boxes * c.rows * 128;
                if (c.stages * stage_bytes > 200 * 1024) continue;
                p.iters = (int)std::max<size_t>(64, ((size_t)8 << 20) / stage_bytes);  // ~8 MiB per CTA
                CUtensorMap tm;
                cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
                cuuint64_t strides[1] = {128};
                cuuint32_t box[2] = {64, (cuuint32_t)c.rows};
                cuuint32_t es[2] = {1, 1};
                CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    printf("encode failed %d\n", (int)r);
                    return 1;
                }
                const size_t smem = c.stages * stage_bytes + 1024 + 512;
                for (int rep = 0; rep < 3; ++rep) ingest<<<grid, 160, smem>>>(tm, p, d_out);
                CK(cudaDeviceSynchronize());
                std::vector<unsigned long long> h(grid);
                CK(cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost));
                double sum = 0, mx = 0;
                for (int i = 0; i < grid; ++i) sum += (double)h[i], mx = std::max(mx, (double)h[i]);
                const double bytes = (double)p.iters * stage_bytes;
                const double bpc = bytes / (sum / grid);
                printf("%-4s grid=%3d stages=%2d x %d box(es) x %3d rows = %6.0f KiB in flight | %6.1f B/clk/SM (slowest %6.1f) | "
                       "stage round trip %6.0f clk | chip %5.2f TB/s @1.9GHz\n",
                       ws == 0 ? "L2" : "DRAM", grid, c.stages, c.boxes, c.rows, c.stages * stage_bytes / 1024.0, bpc, bytes / mx,
                       c.stages * stage_bytes / bpc, bpc * grid * 1.9e9 / 1e12);
                fflush(stdout);
            }
        }
    }
    return 0;
}
