// Blackwell (sm_100a) tensor-core plumbing shared by the contrastive kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) wrappers, UMMA descriptors and the
// host-side tensor-map encoder.  Inline PTX only -- no CUTLASS dependency.
#pragma once
#include <cuda.h>   // CUtensorMap + enums only; the driver entry point is resolved at run time

#include "common.cuh"

namespace dcb {
namespace tc {

// ------------------------------------------------------------------------------------------
// shared-memory addresses, election
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  tcgen05.mma / commit / TMA are warp-level ("uniform datapath") instructions:
// issued under `if (lane == 0)` the compiler cannot prove that a single lane is active and wraps EVERY such
// instruction in an ELECT / BRA.U.ANY serialisation loop (~36 clk per MMA, measured r01g); under elect.sync it does not.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Wait for the phase with the given parity to complete.  A watchdog turns a protocol bug into a
// launch failure instead of a hung GPU (try_wait suspends in hardware, so the loop is not a hot spin).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s at 2 GHz
    }
}

// ------------------------------------------------------------------------------------------
// TMA: 2-D tiled bulk tensor load, completion on an mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, loads, fences  (cta_group::1)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16/fp16 operands, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (quadrant*32 + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 16 consecutive fp32 columns in the mma fragment layout: thread t = 4 g + m receives, in register 4 rep + 2 rr + e,
// lane (base lane + 8 rr + g), column 8 rep + 2 m + e  (verified on B200: scripts/probe/tmem_layout_probe.cu)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 8 columns: register 2 rr + e <-> lane (base lane + 8 rr + g), column 2 m + e
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, float (&v)[4]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// UMMA descriptors
// ------------------------------------------------------------------------------------------
// K-major operand tile stored as rows of 128 bytes (64 x 16-bit) with the 128-byte swizzle, exactly what
// a TMA box {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes: 8-row groups of 1024 B, SBO = 1024 B.
// The tile base must be 1024-byte aligned; advancing K by 16 elements = +32 bytes on the start address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // layout type: SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16: fp32 accumulator, A/B both `fmt` (0 = fp16, 1 = bf16), K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int fmt) {
    return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// Byte offset of the 16-byte chunk holding elements [8c, 8c+8) of row r inside a K-major SW128 tile.
__device__ __forceinline__ uint32_t sw128_chunk_offset(int r, int c) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------------------------------
// host: tensor map for a row-major [rows, cols] matrix of 16-bit elements, box = {64 cols, box_rows}
// ------------------------------------------------------------------------------------------
int encode_tile_map_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                          uint32_t box_rows);   // tensor_map.cu

}  // namespace tc
}  // namespace dcb

// ------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a 2-cluster (same TPC) cooperate on one MMA.  The leader (cluster
// rank 0) issues; each CTA stages its own half of A (M/2 rows) and B (N/2 rows) in its own shared memory at identical
// offsets, and owns the accumulator rows of its half of M in its own TMEM.
// ------------------------------------------------------------------------------------------
namespace dcb {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// Arrive on a barrier that may live in the peer CTA.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id): an explicit .release.cluster costs ~2200 clk per arrive (measured with the pipeline
// trace, r01h) because it fences the thread's whole memory history at cluster scope.  What the peer's MMA consumes is
// ordered by tcgen05.fence::before_thread_sync (TMEM reads) and fence.proxy.async (G tile in shared memory).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose byte count is credited to a barrier that may live in the peer CTA
// (`bar_cluster_addr` is a shared::cluster address, normally the leader's barrier).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {   // one warp in EACH CTA, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t tmem_addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

}  // namespace tc
}  // namespace dcb
