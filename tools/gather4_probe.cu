// Probe of the Blackwell-only TMA row gather / scatter (cp.async.bulk.tensor.2d ... tile::gather4 / tile::scatter4):
// which box shape the tensor map needs, where the four rows land in shared memory (with SWIZZLE_32B), what an
// out-of-range row index does on the load (zeros?) and on the store (skipped?), and whether the transaction count is
// always 4 rows.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe tools/gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int x, int y0, int y1,
                        int y2, int y3, uint8_t* dump, uint32_t tx_bytes, uint32_t* status)
{
    __shared__ __align__(1024) uint8_t tile[1024];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = smem_u32(&bar), t = smem_u32(tile);
    for (int i = threadIdx.x; i < 1024; i += 32) tile[i] = 0xEE;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(t),
                     "l"(&tm_in), "r"(x), "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(b)
                     : "memory");
    }
    // bounded wait: a wrong transaction count must not hang the box
    uint32_t done = 0;
    for (int spin = 0; spin < 2000000 && !done; ++spin)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
    if (threadIdx.x == 0) status[0] = done;
    __syncwarp();
    for (int i = threadIdx.x; i < 1024; i += 32) dump[i] = tile[i];
    if (!done) return;
    // scatter the same tile back out to rows of the output tensor
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(&tm_out), "r"(x),
                     "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(t)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char** argv)
{
    const int box1 = argc > 1 ? atoi(argv[1]) : 1;          // box rows in the tensor map
    const uint32_t tx = argc > 2 ? (uint32_t)atoi(argv[2]) : 128;
    const int oob = argc > 3 ? atoi(argv[3]) : 0;           // make the last row index out of range
    const int swz = argc > 4 ? atoi(argv[4]) : 1;
    const uint64_t pitch = 64, rows = 16, width = 48;       // tensor: 16 rows of 48 valid bytes at a 64 B pitch
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 2; }
    std::vector<uint8_t> h(pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(((i / pitch) << 4) | ((i % pitch) >> 2 & 15));   // row in the high nibble
    uint8_t *d_in, *d_out, *d_dump;
    uint32_t* d_status;
    cudaMalloc(&d_in, h.size()); cudaMalloc(&d_out, h.size()); cudaMalloc(&d_dump, 1024); cudaMalloc(&d_status, 4);
    cudaMemcpy(d_in, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xAA, h.size());
    CUtensorMap tm_in, tm_out;
    const cuuint64_t dims[2] = {width, rows}, strides[1] = {pitch};
    const cuuint32_t box[2] = {32, (cuuint32_t)box1}, es[2] = {1, 1};
    for (int k = 0; k < 2; ++k) {
        CUresult r = ((tmap_encode_fn)fn)(k ? &tm_out : &tm_in, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, k ? d_out : d_in, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed: %d (box1 = %d)\n", (int)r, box1); return 3; }
    }
    // x = 32: the box covers columns 32..63, of which 32..47 are inside the tensor
    const int x = argc > 5 ? atoi(argv[5]) : 16;
    k_probe<<<1, 32>>>(tm_in, tm_out, x, 5, 1, 7, oob ? 1000 : 12, d_dump, tx, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t st = 0;
    std::vector<uint8_t> dump(1024), out(h.size());
    cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(dump.data(), d_dump, 1024, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    printf("box1=%d tx=%u oob=%d swizzle=%d x=%d: sync=%s barrier_done=%u\n", box1, tx, oob, swz, x, cudaGetErrorString(e), st);
    for (int r = 0; r < 8; ++r) {
        printf("smem +%3d:", r * 32);
        for (int i = 0; i < 32; ++i) printf(" %02x", dump[r * 32 + i]);
        printf("\n");
    }
    for (uint64_t r = 0; r < rows; ++r) {
        bool touched = false;
        for (uint64_t i = 0; i < pitch; ++i) touched |= out[r * pitch + i] != 0xAA;
        if (!touched) continue;
        printf("out row %2d:", (int)r);
        for (uint64_t i = 0; i < pitch; ++i) printf("%s%02x", (i % 16) ? "" : " ", out[r * pitch + i]);
        printf("\n");
    }
    return 0;
}
