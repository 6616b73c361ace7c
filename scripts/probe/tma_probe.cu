// Microbenchmark: how many bytes per clock can one SM pull from L2 through TMA, and what changes it?
// One CTA per SM, one producer thread keeps a ring of `stages` slots full, one consumer thread frees a slot as soon as it
// lands (no MMA).  Variants: box height (8/16/32 KiB boxes), swizzle on/off, two producer threads, a 2-CTA cluster where
// each CTA issues half of the boxes multicast to both (every SM still RECEIVES all bytes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/tma_probe.bin scripts/probe/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_load_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct P {
    int stages, stage_bytes, boxes_per_stage, box_rows, iters, producers, multicast, k_chunks, rows_total, three_d, hold;
    int lsu_bytes;            // extra bytes per stage fetched by a cp.async (LDGSTS) warp next to the TMA boxes (0 = off)
    const uint8_t* src;
    long long* clocks;
};

// slot s: boxes_per_stage boxes of [box_rows x 64 cols]; box b of iteration it reads rows ((cta*97 + it*boxes + b) * box_rows) % rows_total
__global__ void __launch_bounds__(192, 1) probe(const __grid_constant__ CUtensorMap map, const __grid_constant__ P p) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + (base - smem_u32(raw)) + p.stages * p.stage_bytes);
    const uint32_t full = smem_u32(bars), empty = full + 8 * p.stages, lfull = empty + 8 * p.stages;
    const int tma_bytes = p.stage_bytes - p.lsu_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = p.multicast ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full + 8 * s, p.producers);
            mbar_init(empty + 8 * s, p.multicast ? 2 : 1);
            mbar_init(lfull + 8 * s, 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (p.multicast) cluster_sync();
    const int box_bytes = p.box_rows * 128 * (p.three_d ? 2 : 1);
    const long long t0 = clock64();
    if (warp < p.producers && lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int per = p.boxes_per_stage / p.producers;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(empty + 8 * stage, phase ^ 1);
            const int kx = (it & 7) * 64;
            const uint32_t dst = base + stage * p.stage_bytes;
            if (!p.multicast) {
                mbar_expect(full + 8 * stage, per * box_bytes);
                for (int b = warp * per; b < (warp + 1) * per; ++b) {
                    const int row = (int)(((unsigned)(blockIdx.x * 97 + it * p.boxes_per_stage + b) * (unsigned)p.box_rows) & (unsigned)(p.rows_total - 1));
                    if (p.three_d) tma_load3(dst + b * box_bytes, &map, full + 8 * stage, 0, row, (it & 3) * 2);
                    else tma_load(dst + b * box_bytes, &map, full + 8 * stage, kx, row);
                }
            } else {
                // both CTAs receive every box; this CTA issues the boxes with b % 2 == rank
                mbar_expect(full + 8 * stage, p.boxes_per_stage * box_bytes);
                for (int b = (int)rank; b < p.boxes_per_stage; b += 2) {
                    const int row = (int)(((unsigned)((blockIdx.x >> 1) * 97 + it * p.boxes_per_stage + b) * (unsigned)p.box_rows) & (unsigned)(p.rows_total - 1));
                    tma_load_mc(dst + b * box_bytes, &map, full + 8 * stage, kx, row, 3);
                }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 4 && p.lsu_bytes > 0) {
        // second ingest path: one warp, 16 bytes per lane per cp.async, 128-byte rows (8 lanes per row), completion through
        // cp.async.mbarrier.arrive.noinc on a barrier of its own
        int stage = 0;
        uint32_t phase = 0;
        const int rows = p.lsu_bytes / 128;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(empty + 8 * stage, phase ^ 1);
            const uint32_t dst = base + stage * p.stage_bytes + tma_bytes;
            const unsigned row0 = ((unsigned)(blockIdx.x * 131 + it) * (unsigned)rows) & (unsigned)(p.rows_total - 1);
            const int kx = ((it + 3) & 7) * 128;
            for (int r = lane >> 3; r < rows; r += 4) {
                const uint8_t* g = p.src + (size_t)(row0 + r) * 1536 + kx + (lane & 7) * 16;
                const uint32_t d = dst + r * 128 + (((lane & 7) ^ (r & 7)) << 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(lfull + 8 * stage) : "memory");
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 5 && lane == 0) {
        long long last = 0;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(full + 8 * stage, phase);
            if (p.lsu_bytes > 0) mbar_wait(lfull + 8 * stage, phase);
            if (p.hold) {                       // emulate an in-order consumer that needs `hold` clocks per stage
                const long long until = (it == 0 ? clock64() : last) + p.hold;
                while (clock64() < until) {}
                last = until > clock64() - p.hold ? until : clock64();
            }
            if (!p.multicast) {
                mbar_arrive(empty + 8 * stage);
            } else {       // the slot is free for refill (by EITHER producer) once both CTAs have consumed it
                mbar_arrive_remote(map_to_cta(empty + 8 * stage, 0));
                mbar_arrive_remote(map_to_cta(empty + 8 * stage, 1));
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        p.clocks[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (p.multicast) cluster_sync();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows_total = 32768, cols = 768;
    void* buf;
    CK(cudaMalloc(&buf, (size_t)rows_total * cols * 2));
    CK(cudaMemset(buf, 1, (size_t)rows_total * cols * 2));
    long long* clocks;
    CK(cudaMalloc(&clocks, 148 * 8));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn encode = (EncodeFn)fp;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Case { const char* name; int stages, boxes, box_rows, producers, multicast, swizzle, ctas, three_d, hold, lsu; };
    const Case cases[] = {
        {"8K boxes x4, 6 stages (as the kernels)", 6, 4, 64, 1, 0, 1, 148},
        {"same, 16 CTAs only", 6, 4, 64, 1, 0, 1, 16},
        {"16K boxes x2, 6 stages", 6, 2, 128, 1, 0, 1, 148},
        {"32K box x1, 6 stages", 6, 1, 256, 1, 0, 1, 148},
        {"8K boxes x4, no swizzle", 6, 4, 64, 1, 0, 0, 148},
        {"8K boxes x4, 2 producer threads", 6, 4, 64, 2, 0, 1, 148},
        {"8K boxes x2 (16K stages), 12 stages", 12, 2, 64, 1, 0, 1, 148},
        {"8K boxes x8 (64K stages), 3 stages", 3, 8, 64, 1, 0, 1, 148},
        {"16K boxes x2, 2 producer threads", 6, 2, 128, 2, 0, 1, 148},
        {"16K boxes x4 (64K stages), 3 stages (fwd kernel)", 3, 4, 128, 1, 0, 1, 148},
        {"16K boxes x4 (64K stages), 3 stages, 2 producers", 3, 4, 128, 2, 0, 1, 148},
        {"16K boxes x4 (64K stages), 3 stages, 4 producers", 3, 4, 128, 4, 0, 1, 148},
        {"8K boxes x4, 4 producer threads", 6, 4, 64, 4, 0, 1, 148},
        {"16K boxes x3 (48K stages), 4 stages", 4, 3, 128, 1, 0, 1, 148},
        {"16K x2 + 8K x2 per stage ~ (48K) modelled as 8K x6, 4 stages", 4, 6, 64, 1, 0, 1, 148},
        {"8K x6 per stage, 4 stages, 2 producers", 4, 6, 64, 2, 0, 1, 148},
        {"3-D boxes {64,64 rows,2 chunks} = 16K x2, 6 stages", 6, 2, 64, 1, 0, 1, 148, 1, 0},
        {"8K x4, 6 stages, consumer holds 272 clk/stage", 6, 4, 64, 1, 0, 1, 148, 0, 272},
        {"16K x2, 6 stages, consumer holds 272 clk/stage", 6, 2, 128, 1, 0, 1, 148, 0, 272},
        {"8K x4, 6 stages, consumer holds 400 clk/stage", 6, 4, 64, 1, 0, 1, 148, 0, 400},
        {"16K x2 by TMA + 16K by a cp.async warp (48K stages), 4 stages", 4, 2, 128, 1, 0, 1, 148, 0, 0, 16384},
        {"16K x2 by TMA + 8K by a cp.async warp (40K stages), 4 stages", 4, 2, 128, 1, 0, 1, 148, 0, 0, 8192},
        {"16K x2 by TMA alone (32K stages), 4 stages", 4, 2, 128, 1, 0, 1, 148, 0, 0, 0},
        {"16K x1 by TMA + 16K by a cp.async warp (32K stages), 6 stages", 6, 1, 128, 1, 0, 1, 148, 0, 0, 16384},
        {"multicast pair: each CTA issues 2 of 4 boxes", 6, 4, 64, 1, 1, 1, 148},
        {"multicast pair, 16 CTAs only", 6, 4, 64, 1, 1, 1, 16},
        {"multicast pair, 16K boxes", 6, 2, 128, 1, 1, 1, 148},
    };
    for (const Case& c : cases) {
        CUtensorMap map;
        const cuuint64_t gdim[3] = {(cuuint64_t)(c.three_d ? 64 : cols), (cuuint64_t)rows_total, (cuuint64_t)(cols / 64)};
        const cuuint64_t gstride[2] = {(cuuint64_t)cols * 2, 128};
        const cuuint32_t box[3] = {64, (cuuint32_t)c.box_rows, 2};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, c.three_d ? 3 : 2, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            c.swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        P p{};
        p.stages = c.stages;
        p.boxes_per_stage = c.boxes;
        p.box_rows = c.box_rows;
        p.stage_bytes = c.boxes * c.box_rows * 128 * (c.three_d ? 2 : 1) + c.lsu;
        p.lsu_bytes = c.lsu;
        p.src = static_cast<const uint8_t*>(buf);
        p.three_d = c.three_d;
        p.hold = c.hold;
        p.iters = 2000;
        p.producers = c.producers;
        p.multicast = c.multicast;
        p.k_chunks = cols / 64;
        p.rows_total = rows_total;
        p.clocks = clocks;
        const int smem = 1024 + p.stages * p.stage_bytes + 24 * p.stages + 64;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(c.ctas);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = c.multicast ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaLaunchKernelEx(&cfg, probe, map, p));
            CK(cudaDeviceSynchronize());
        }
        long long h[148];
        CK(cudaMemcpy(h, clocks, c.ctas * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < c.ctas; ++i) avg += (double)h[i];
        avg /= c.ctas;
        const double bytes = (double)p.iters * p.stage_bytes;
        printf("%-50s ring %3d KiB  %6.1f B/clk received per SM (%6.1f issued)   %7.0f clk per 32 KiB\n", c.name,
               p.stages * p.stage_bytes / 1024, bytes / avg, bytes / avg / (c.multicast ? 2 : 1), avg / (bytes / 32768.0));
    }
    return 0;
}
