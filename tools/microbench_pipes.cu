// Pipe-rate microbenchmarks behind DESIGN.md 4.1 (formulation choice): how many LOP3 and
// how many conflict-free 32-bit shared-memory lookups one B200 SM retires per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_pipes.cu && /tmp/mb
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(1024, 1) k_lop3(uint32_t* out, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x + 7, c = a ^ 0x9e3779b9u, d = b * 3 + 1;
    uint32_t e = a + 11, f = b ^ 0x55aa55aau, g = c + 5, h = d ^ a;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // 8 chains, each statement is one 3-input LOP3
            a = (a ^ b) | (c & a); b = (b ^ c) | (d & b); c = (c ^ d) | (e & c); d = (d ^ e) | (f & d);
            e = (e ^ f) | (g & e); f = (f ^ g) | (h & f); g = (g ^ h) | (a & g); h = (h ^ a) | (b & h);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

extern __shared__ uint32_t tab[];
__global__ void __launch_bounds__(1024, 1) k_lds(uint32_t* out, int iters)
{
    for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t lane4 = (threadIdx.x & 31) * 4;
    uint32_t s0 = threadIdx.x, s1 = blockIdx.x, s2 = 3, s3 = 5;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(tab);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // data-dependent, lane-private bank: the AES lookup pattern (PRMT + LDS)
            const uint32_t t0 = *reinterpret_cast<const uint32_t*>(base + (__byte_perm(s0, lane4, 0x5504)));
            const uint32_t t1 = *reinterpret_cast<const uint32_t*>(base + (__byte_perm(s1, lane4, 0x5514)));
            const uint32_t t2 = *reinterpret_cast<const uint32_t*>(base + (__byte_perm(s2, lane4, 0x5524)));
            const uint32_t t3 = *reinterpret_cast<const uint32_t*>(base + (__byte_perm(s3, lane4, 0x5534)));
            s0 ^= t1; s1 ^= t2; s2 ^= t3; s3 ^= t0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 ^ s1 ^ s2 ^ s3;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    const int iters = 4000;
    k_lop3<<<sms, 1024>>>(out, 10);
    cudaEventRecord(e0);
    k_lop3<<<sms, 1024>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // 8 statements x 16 = 128 LOP3 per iteration per thread (checked in SASS)
    const double lop = 128.0 * iters * 1024.0 * sms;
    printf("{\"bench\":\"lop3\",\"sms\":%d,\"ms\":%.3f,\"lop3_per_s\":%.3e,\"lop3_per_clk_per_sm_at_%dMHz\":%.1f}\n", sms, ms,
           lop / (ms * 1e-3), clk_khz / 1000, lop / (ms * 1e-3) / sms / (clk_khz * 1e3));
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_lds<<<sms, 1024, 65536>>>(out, 10);
    cudaEventRecord(e0);
    k_lds<<<sms, 1024, 65536>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double lds = 32.0 * iters * 1024.0 * sms;
    printf("{\"bench\":\"lds32_lane_private\",\"sms\":%d,\"ms\":%.3f,\"lookups_per_s\":%.3e,\"lookups_per_clk_per_sm_at_%dMHz\":%.1f}\n", sms,
           ms, lds / (ms * 1e-3), clk_khz / 1000, lds / (ms * 1e-3) / sms / (clk_khz * 1e3));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
