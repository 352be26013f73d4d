// Settles the "T-table + bitsliced warps on one SM" question by
// MEASUREMENT, in the form most favourable to the hybrid: the product's fused kernel k_stream
// (AES-256-GCM, 2^30 B, one persistent 512-thread CTA per SM, bound by the L1/shared data pipe)
// runs back to back on one stream, while on a second stream W extra warps per SM evaluate the
// bitsliced AES S-box (tools/bitslice_sbox_generated.cuh, registers only, ALU pipe only).  Two
// kernels rather than two roles in one kernel, so that the bitsliced side keeps its own register
// budget (all threads of one kernel get the same allocation: a real hybrid would have to fit both
// formulations in one) -- and the bitsliced side is credited with a whole AES-256 block per 224
// S-box bytes, i.e. MixColumns, AddRoundKey, the counter setup, the 128x32 bit transpose of the
// keystream and its GHASH are all counted as FREE.  If even this bound gains less than 5 %, the
// hybrid is not worth a second formulation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o /tmp/hybrid_probe tools/hybrid_probe.cu \
//        -L aes-gcm-128-192-256-bits_b200 -laesgcm_b200 -Xlinker -rpath=$PWD/aes-gcm-128-192-256-bits_b200
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../include/aesgcm_b200.h"
#include "bitslice_sbox_generated.cuh"

// `warps` warps per CTA, one CTA per SM; 4 independent byte groups per thread (ILP), fed back
__global__ void k_sbox(uint32_t* out, int iters)
{
    extern __shared__ uint32_t pad_smem[];   // a few KB of dynamic shared memory: same L1/shared split as k_stream
    if (iters < 0) pad_smem[threadIdx.x] = 0;
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

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define AK(x) do { int rc_ = (x); if (rc_ != 0) { printf("agcm error %d (%s) at %d\n", rc_, agcm_strerror(rc_), __LINE__); return 1; } } while (0)

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t n = 1ull << 30;
    uint8_t *d_in, *d_out, *d_tag, *d_aad;
    CK(cudaMalloc(&d_in, n));
    CK(cudaMalloc(&d_out, n));
    CK(cudaMalloc(&d_tag, 16));
    CK(cudaMalloc(&d_aad, 16));
    CK(cudaMemset(d_in, 0x5a, n));
    CK(cudaMemset(d_aad, 1, 16));
    agcm_ctx* ctx = nullptr;
    AK(agcm_ctx_create(&ctx, 0));
    uint8_t key[32], iv[12] = {0};
    for (int i = 0; i < 32; ++i) key[i] = (uint8_t)i;
    AK(agcm_set_key(ctx, 256, 0, key, 32));
    cudaStream_t sa, sb;
    CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    uint32_t* d_sink;
    CK(cudaMalloc(&d_sink, sizeof(uint32_t) * sms * 1024));
    cudaEvent_t a0, a1, b0, b1;
    cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
    const int reps = 12;   // k_stream launches per measurement (~25 ms)

    auto run_stream = [&](float* ms) -> int {
        CK(cudaEventRecord(a0, sa));
        for (int r = 0; r < reps; ++r) AK(agcm_stream_crypt(ctx, 0, iv, d_aad, 16, d_in, d_out, n, d_tag, nullptr, sa));
        CK(cudaEventRecord(a1, sa));
        return 0;
    };
    // warm up, then the T-table kernel alone
    float t_alone = 0.f;
    for (int w = 0; w < 2; ++w) {
        if (run_stream(nullptr)) return 1;
        CK(cudaStreamSynchronize(sa));
    }
    CK(cudaEventElapsedTime(&t_alone, a0, a1));
    const double gbps_alone = (double)n * reps / (t_alone * 1e-3) / 1e9;
    // the S-box kernel alone: iterations per ms at each width, to size the concurrent runs
    printf("| bitsliced warps per SM | k_stream GB/s | k_stream slowdown | S-box bytes/clk/SM in its shadow | credited AES-256 GB/s (224 S-box bytes = 1 block, all else free) | combined GB/s | gain |\n");
    printf("|---|---|---|---|---|---|---|\n");
    printf("| 0 | %.1f | - | - | - | %.1f | - |\n", gbps_alone, gbps_alone);
    // an SM changes its L1/shared split only when idle: ask for the same split as k_stream (all shared) and take a
    // few KB of dynamic shared memory, or the two kernels never share an SM (first attempts: fully serialised)
    CK(cudaFuncSetAttribute(k_sbox, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int widths[] = {1, 2, 4, 6};   // 7 warps of 48 registers is what fits next to 512 x 106 in the register file
    for (int wi = 0; wi < 4; ++wi) {
        const int warps = widths[wi], threads = 32 * warps;
        // calibrate alone
        k_sbox<<<sms, threads, 4096, sb>>>(d_sink, 200);
        CK(cudaStreamSynchronize(sb));
        CK(cudaEventRecord(b0, sb));
        k_sbox<<<sms, threads, 4096, sb>>>(d_sink, 2000);
        CK(cudaEventRecord(b1, sb));
        CK(cudaStreamSynchronize(sb));
        float ms_cal = 0.f;
        CK(cudaEventElapsedTime(&ms_cal, b0, b1));
        // concurrent: size the S-box kernel for ~1.3x the stream run when alone (it will be slowed a little)
        int iters = (int)(2000.0 * (t_alone * 1.3) / ms_cal);
        if (iters < 100) iters = 100;
        if (run_stream(nullptr)) return 1;   // the persistent T-table CTAs first, one per SM ...
        CK(cudaEventRecord(b0, sb));
        k_sbox<<<sms, threads, 4096, sb>>>(d_sink, iters);   // ... then one bitsliced CTA per SM next to them
        CK(cudaEventRecord(b1, sb));
        CK(cudaStreamSynchronize(sa));
        CK(cudaStreamSynchronize(sb));
        float t_a = 0.f, t_b = 0.f;
        CK(cudaEventElapsedTime(&t_a, a0, a1));
        CK(cudaEventElapsedTime(&t_b, b0, b1));
        const double gbps = (double)n * reps / (t_a * 1e-3) / 1e9;
        const double sbox_bytes = 4.0 * iters * threads * (double)sms * 32.0;
        // the S-box kernel outlasts the stream run: its rate over its own duration mixes shadowed and free time, so
        // use only a lower bound of the shadowed rate?  No: be generous -- credit it with its AVERAGE rate.
        const double rate = sbox_bytes / (t_b * 1e-3);
        const double credited = rate / 224.0 * 16.0 / 1e9;
        printf("| %d | %.1f | %.1f %% | %.2f | %.1f | %.1f | %+.1f %% |\n", warps, gbps, 100.0 * (1.0 - gbps / gbps_alone),
               rate / sms / (clk_khz * 1e3), credited, gbps + credited, 100.0 * ((gbps + credited) / gbps_alone - 1.0));
    }
    printf("\nk_stream alone: %.3f ms per 2^30 B launch; SM clock %d MHz; %d SMs.  %s\n", t_alone / reps, clk_khz / 1000, sms,
           cudaGetErrorString(cudaGetLastError()));
    agcm_ctx_destroy(ctx);
    return 0;
}
