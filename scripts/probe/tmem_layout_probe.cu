// Which (lane, column) does register k of thread t receive from tcgen05.ld.16x256b.xN?  Writes a known pattern with
// tcgen05.st.32x32b (thread t of warp w <-> lane 32 w + t, registers = consecutive columns) and reads it back with 16x256b.x2
// at lane offsets 0 and 16 of each warp's quadrant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/tmem_layout_probe.bin scripts/probe/tmem_layout_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t row = warp * 32 + lane;
    const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = row * 100 + c;            // value = 100 * lane + column
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int h = 0; h < 2; ++h) {
        uint32_t r[8];
        const uint32_t a = base + ((uint32_t)(warp * 32 + h * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 8; ++k) out[((warp * 2 + h) * 32 + lane) * 8 + k] = r[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}

int main() {
    uint32_t* d;
    CK(cudaMalloc(&d, 4 * 2 * 32 * 8 * 4));
    probe<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    static uint32_t h[4 * 2 * 32 * 8];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int w = 0; w < 4; ++w)
        for (int hh = 0; hh < 2; ++hh)
            for (int t = 0; t < 32; ++t)
                for (int k = 0; k < 8; ++k) {
                    const uint32_t v = h[((w * 2 + hh) * 32 + t) * 8 + k];
                    const int lane = v / 100, col = v % 100;
                    // expectation: register 4 rep + 2 rr + e  <->  lane 32 w + 16 hh + 8 rr + t / 4, column 8 rep + 2 (t % 4) + e
                    const int rep = k >> 2, rr = (k >> 1) & 1, e = k & 1;
                    const int el = 32 * w + 16 * hh + 8 * rr + t / 4, ec = 8 * rep + 2 * (t % 4) + e;
                    if (lane != el || col != ec) {
                        if (bad < 24) printf("warp %d half %d thread %2d reg %d: lane %3d col %2d (expected lane %3d col %2d)\n", w, hh, t, k, lane, col, el, ec);
                        ++bad;
                    }
                }
    printf(bad ? "MISMATCH: %d registers differ from the expected 16x256b mapping\n" : "OK: reg 4 rep + 2 rr + e <-> lane 16 h + 8 rr + t/4, column 8 rep + 2 (t%%4) + e\n", bad);
    return 0;
}
