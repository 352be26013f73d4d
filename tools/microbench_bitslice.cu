// How fast can the B200 integer pipe evaluate the AES S-box as LOGIC (bitsliced, 32 bytes per
// call in 8 bit-planes)?  Upper-bounds a bitsliced AES: AES-256 needs 14 x 16 = 224 S-box bytes
// per 16-byte block, before MixColumns / AddRoundKey / the output transpose.
//   python tools/gen_bitslice_sbox.py && nvcc -gencode arch=compute_100a,code=sm_100a -O3 \
//       -o /tmp/mbs tools/microbench_bitslice.cu && /tmp/mbs
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "bitslice_sbox_generated.cuh"

__global__ void __launch_bounds__(256, 4) k_sbox(uint32_t* out, int iters)
{
    // 4 independent byte-groups per thread (ILP), each fed back into itself
    uint32_t a[8], b[8], c[8], d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = threadIdx.x * 2654435761u + k;
        b[k] = blockIdx.x * 40503u + 7 * k;
        c[k] = (threadIdx.x ^ blockIdx.x) * 2246822519u + k;
        d[k] = threadIdx.x + blockIdx.x + 13 * k;
    }
    for (int i = 0; i < iters; ++i) {
        uint32_t s[8];
        bitslice_sbox(a, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = s[k];
        bitslice_sbox(b, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = s[k];
        bitslice_sbox(c, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = s[k];
        bitslice_sbox(d, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = s[k];
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k] ^ b[k] ^ c[k] ^ d[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// correctness on the device too: S-box of bytes 0..255 through the bit planes
__global__ void k_check(uint8_t* sb)
{
    // thread t handles bytes 32t .. 32t+31
    uint32_t x[8], s[8];
    for (int k = 0; k < 8; ++k) {
        x[k] = 0;
        for (int j = 0; j < 32; ++j) x[k] |= (uint32_t)(((32 * threadIdx.x + j) >> k) & 1) << j;
    }
    bitslice_sbox(x, s);
    for (int j = 0; j < 32; ++j) {
        uint8_t v = 0;
        for (int k = 0; k < 8; ++k) v |= (uint8_t)(((s[k] >> j) & 1) << k);
        sb[32 * threadIdx.x + j] = v;
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    uint8_t* d_sb;
    cudaMalloc(&d_sb, 256);
    k_check<<<1, 8>>>(d_sb);
    uint8_t sb[256];
    cudaMemcpy(sb, d_sb, 256, cudaMemcpyDeviceToHost);
    const bool ok = sb[0] == 0x63 && sb[1] == 0x7c && sb[0x53] == 0xed && sb[255] == 0x16;
    uint32_t* out;
    const int blocks = sms * 4;
    cudaMalloc(&out, sizeof(uint32_t) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2000;
    k_sbox<<<blocks, 256>>>(out, 10);
    cudaEventRecord(e0);
    k_sbox<<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double calls = 4.0 * iters * 256.0 * blocks;       // each call = 32 S-box bytes
    const double bytes_per_s = calls * 32.0 / (ms * 1e-3);
    printf("{\"bench\":\"bitsliced sbox\",\"device_check\":%s,\"gates\":%d,\"ms\":%.3f,\"sbox_bytes_per_s\":%.3e,"
           "\"sbox_bytes_per_clk_per_sm\":%.2f,\"aes256_ctr_upper_bound_GBps\":%.1f,\"aes128_ctr_upper_bound_GBps\":%.1f}\n",
           ok ? "true" : "false", BITSLICE_SBOX_GATES, ms, bytes_per_s, bytes_per_s / sms / (clk_khz * 1e3),
           bytes_per_s / 224.0 * 16.0 / 1e9, bytes_per_s / 160.0 * 16.0 / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
