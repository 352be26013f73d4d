// Does the texture path add lookup bandwidth on top of the LSU/shared-memory data pipe?
// Three kernels with the AES access pattern (random byte -> 32-bit table word):
//   lds   : 32 shared-memory lookups per iteration (lane-private replicas)
//   tex   : 32 texture fetches per iteration (256-entry linear texture, L1-resident)
//   mixed : 32 shared + 8 texture lookups per iteration
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

extern __shared__ uint32_t tab[];

template <int NLDS, int NTEX>
__global__ void __launch_bounds__(1024, 1) k_mix(cudaTextureObject_t tex, uint32_t* out, int iters)
{
    for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t lane4 = (threadIdx.x & 31) * 4;
    uint32_t s0 = threadIdx.x * 0x01010101u, s1 = blockIdx.x * 0x9e3779b9u, s2 = 3, s3 = 5;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(tab);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NLDS / 4; ++j) {
            const uint32_t t0 = *reinterpret_cast<const uint32_t*>(base + __byte_perm(s0, lane4, 0x5504));
            const uint32_t t1 = *reinterpret_cast<const uint32_t*>(base + __byte_perm(s1, lane4, 0x5514));
            const uint32_t t2 = *reinterpret_cast<const uint32_t*>(base + __byte_perm(s2, lane4, 0x5524));
            const uint32_t t3 = *reinterpret_cast<const uint32_t*>(base + __byte_perm(s3, lane4, 0x5534));
            s0 ^= t1; s1 ^= t2; s2 ^= t3; s3 ^= t0;
        }
#pragma unroll
        for (int j = 0; j < NTEX / 4; ++j) {
            const uint32_t t0 = tex1Dfetch<uint32_t>(tex, (s0 >> 3) & 0xff);
            const uint32_t t1 = tex1Dfetch<uint32_t>(tex, (s1 >> 11) & 0xff);
            const uint32_t t2 = tex1Dfetch<uint32_t>(tex, (s2 >> 17) & 0xff);
            const uint32_t t3 = tex1Dfetch<uint32_t>(tex, (s3 >> 23) & 0xff);
            s0 ^= t1; s1 ^= t2; s2 ^= t3; s3 ^= t0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 ^ s1 ^ s2 ^ s3;
}

template <int NLDS, int NTEX>
void run(const char* name, cudaTextureObject_t tex, uint32_t* out, int sms, int clk_khz)
{
    const int iters = 2000;
    cudaFuncSetAttribute(k_mix<NLDS, NTEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_mix<NLDS, NTEX><<<sms, 1024, 65536>>>(tex, out, 10);
    cudaEventRecord(e0);
    k_mix<NLDS, NTEX><<<sms, 1024, 65536>>>(tex, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)(NLDS + NTEX) * iters * 1024.0 * sms;
    printf("{\"bench\":\"%s\",\"lds_per_iter\":%d,\"tex_per_iter\":%d,\"ms\":%.3f,\"lookups_per_clk_per_sm\":%.2f}\n", name, NLDS, NTEX,
           ms, n / (ms * 1e-3) / sms / (clk_khz * 1e3));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out, *d_tab;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&d_tab, 1024);
    uint32_t h[256];
    for (int i = 0; i < 256; ++i) h[i] = i * 2654435761u;
    cudaMemcpy(d_tab, h, 1024, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = d_tab;
    rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>();
    rd.res.linear.sizeInBytes = 1024;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    run<32, 0>("lds", tex, out, sms, clk_khz);
    run<0, 32>("tex", tex, out, sms, clk_khz);
    run<32, 8>("mixed 32+8", tex, out, sms, clk_khz);
    run<32, 16>("mixed 32+16", tex, out, sms, clk_khz);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
