// Thin wrappers over the sm_90+/sm_100 async-copy PTX used by the tiled batch kernels: mbarrier
// (transaction-count completion), 2-D tensor-map TMA loads and stores (cp.async.bulk.tensor), the
// bulk-group waits of the store side and the generic->async proxy fence.  One elected lane issues;
// the data movement itself costs no LSU (L1/shared data pipe) wavefronts, which is the point: that
// pipe is the binding one for every AES-GCM kernel of this engine.
#pragma once
#include <cuda.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t ag_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ag_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// make the barrier inits visible to the async proxy before the first TMA load signals one
__device__ __forceinline__ void ag_fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void ag_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void ag_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AG_DONE_%=;\n"
        "bra AG_WAIT_%=;\n"
        "AG_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// box (c0 .. c0+box0, c1 .. c1+box1) of the tensor -> dense (swizzled) tile at `dst`; completes on `bar`.
// Elements outside the tensor's extents arrive as zeros.
__device__ __forceinline__ void ag_tma_load_2d(uint32_t dst, const CUtensorMap* tm, int32_t c0, int32_t c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

// tile at `src` -> the same box of the tensor; elements outside the extents are not written.
__device__ __forceinline__ void ag_tma_store_2d(const CUtensorMap* tm, int32_t c0, int32_t c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}

// Blackwell row gather / scatter: four rows y0..y3 of the tensor, box {box0 bytes, 1 row} each, to / from 4 x box0
// contiguous (swizzled) bytes.  A row index outside the tensor loads zeros (still counted in the transaction) and
// stores nothing; columns outside the extent likewise (tools/gather4_probe.cu).
__device__ __forceinline__ void ag_tma_gather4(uint32_t dst, const CUtensorMap* tm, int32_t c0, int32_t y0, int32_t y1, int32_t y2,
                                               int32_t y3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void ag_tma_scatter4(const CUtensorMap* tm, int32_t c0, int32_t y0, int32_t y1, int32_t y2, int32_t y3,
                                                uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(tm), "r"(c0),
                 "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(src)
                 : "memory");
}

__device__ __forceinline__ void ag_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed stores have READ their shared-memory source (the tile may be overwritten)
__device__ __forceinline__ void ag_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed stores are complete (their global writes are visible)
__device__ __forceinline__ void ag_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's generic-proxy shared-memory writes before later async-proxy (TMA) reads
__device__ __forceinline__ void ag_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void ag_prefetch_tmap(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
