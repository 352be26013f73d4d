// sm_100a kernels of the AES-GCM engine.
//
// Shared-memory plan (one persistent 512-thread CTA per SM, 194 KB of the 227 KB):
//
//   [0      , 64 KB)  AES_A : 256 entries x 256 B; entry x = 32 lane-private copies
//                     of Te0[x] (128 B) then 32 copies of Te1[x] (128 B)
//   [64 KB  , 128 KB) AES_B : same for Te2 / Te3
//   [128 KB , 192 KB) GH    : 256 entries x 256 B; entry b = 8 copies of the 16 B row
//                     T_a[b] (128 B) then 8 copies of T_b[b] (128 B)
//   [192 KB , +2 KB)  reduction scratch
//
// Every data-dependent lookup is then bank-conflict free BY CONSTRUCTION: lane l
// reads word l of a 128 B row (32-bit AES lookups), or 16 B slot l%8 of a row
// (128-bit GHASH lookups, served per quarter-warp).  The 256 B entry stride makes
// the address a single PRMT: {0, 0, index byte, lane offset}; the table select is
// an immediate on the LDS.  Per 16 B block that is 16 PRMT + 16 LDS + 8 LOP3 per
// AES round and 16 PRMT + 16 LDS.128 + ~90 LOP3/SHF per GHASH multiply.
//
// Reference blocks replaced: gcm_gctr (aes_icb + aes_ecb + xor, src/gcm_gctr.vhd:150),
// gcm_ghash + ghash_gfmul (src/gcm_ghash.vhd:225-293, src/ghash_gfmul.vhd:42-63),
// aes_kexp (config/config_aes_kexp.py:128-159, tb/key_exp.py:79-114).
#include <cuda_runtime.h>
#include <stdint.h>
#include "gcm_core.cuh"
#include "perkey_core.cuh"
#include "kernels.h"
#include "tma_util.cuh"

namespace {

constexpr uint32_t SM_AES_A = 0;
constexpr uint32_t SM_AES_B = 65536;
constexpr uint32_t SM_GH = 131072;
constexpr uint32_t SM_MISC = 196608;

extern __shared__ __align__(1024) uint8_t ag_smem[];

// Te_tab[(w >> 8k) & 0xff] from the lane-private replicas
struct TeSmem {
    const uint8_t* base;
    uint32_t lane4;  // (lane & 31) * 4
    __device__ __forceinline__ uint32_t operator()(int tab, uint32_t w, int k) const
    {
        const uint32_t off = __byte_perm(w, lane4, 0x5504 | (k << 4));  // (byte_k << 8) | lane4
        return *reinterpret_cast<const uint32_t*>(base + off + (tab & 1) * 128 + (tab >> 1) * 65536);
    }
};

// row T[(w >> 8k) & 0xff] of the GHASH table from the quarter-warp replicas
struct GhSmem {
    const uint8_t* base;  // ag_smem + SM_GH (+128 for T_b)
    uint32_t lane16;      // (lane & 7) * 16
    __device__ __forceinline__ uint4 operator()(uint32_t w, int k) const
    {
        const uint32_t off = __byte_perm(w, lane16, 0x5504 | (k << 4));
        return *reinterpret_cast<const uint4*>(base + off);
    }
};

// slow-path lookups straight from HBM/L2 (setup and finish kernels only)
struct TeGlobal {
    const uint32_t* te0;
    __device__ __forceinline__ uint32_t operator()(int tab, uint32_t w, int k) const
    {
        const uint32_t t = __ldg(te0 + ((w >> (8 * k)) & 0xff));
        return tab ? ag_rotl32(t, 8 * tab) : t;
    }
};

// Te0 (1 KB) is staged once through the reduction scratch with one coalesced load per thread, so
// the 16 expansion passes read shared memory instead of paying a global-load latency each: the
// table fill is most of a short message's kernel time.  Order at the call sites: stage_te0 and
// fill_gh_tables (their global loads overlap), __syncthreads, expand_aes_tables, __syncthreads.
__device__ __forceinline__ void stage_te0(const uint32_t* __restrict__ te0)
{
    uint32_t* stage = reinterpret_cast<uint32_t*>(ag_smem + SM_MISC);
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) stage[i] = __ldg(te0 + i);
}

__device__ __forceinline__ void expand_aes_tables()
{
    const uint32_t* stage = reinterpret_cast<const uint32_t*>(ag_smem + SM_MISC);
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < 256 * 32; idx += blockDim.x) {
        const uint32_t x = idx >> 5, l = idx & 31;
        const uint32_t t = stage[x];
        uint32_t* a = reinterpret_cast<uint32_t*>(ag_smem + SM_AES_A + x * 256 + l * 4);
        uint32_t* b = reinterpret_cast<uint32_t*>(ag_smem + SM_AES_B + x * 256 + l * 4);
        a[0] = t;
        a[32] = ag_rotl32(t, 8);
        b[0] = ag_rotl32(t, 16);
        b[32] = ag_rotl32(t, 24);
    }
}

__device__ __forceinline__ void fill_gh_tables(const uint4* __restrict__ ta, const uint4* __restrict__ tb)
{
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < 256 * 8; idx += blockDim.x) {
        const uint32_t b = idx >> 3, r = idx & 7;
        uint4* d = reinterpret_cast<uint4*>(ag_smem + SM_GH + b * 256 + r * 16);
        d[0] = __ldg(ta + b);
        if (tb) d[8] = __ldg(tb + b);
    }
}

__device__ __forceinline__ gf128 warp_xor(gf128 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.w[0] ^= __shfl_xor_sync(0xffffffffu, v.w[0], o);
        v.w[1] ^= __shfl_xor_sync(0xffffffffu, v.w[1], o);
        v.w[2] ^= __shfl_xor_sync(0xffffffffu, v.w[2], o);
        v.w[3] ^= __shfl_xor_sync(0xffffffffu, v.w[3], o);
    }
    return v;
}

// H^e for a 64-bit exponent, computed by one full warp: lane k contributes
// pow2[k]^(bit k) * pow2[k+32]^(bit k+32); the 32 factors are multiplied by a
// shuffle tree (5 generic products deep).  All lanes return the result.
__device__ gf128 warp_gf_pow(const KeyDev* kd, uint64_t e)
{
    const uint32_t lane = threadIdx.x & 31;
    gf128 f = ((e >> lane) & 1) ? kd->pow2[lane] : gf_one();
    if (e >> 32) {  // uniform
        gf128 f2 = ((e >> (lane + 32)) & 1) ? kd->pow2[lane + 32] : gf_one();
        f = gf_mul(f, f2);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        gf128 g;
        g.w[0] = __shfl_xor_sync(0xffffffffu, f.w[0], o);
        g.w[1] = __shfl_xor_sync(0xffffffffu, f.w[1], o);
        g.w[2] = __shfl_xor_sync(0xffffffffu, f.w[2], o);
        g.w[3] = __shfl_xor_sync(0xffffffffu, f.w[3], o);
        f = gf_mul(f, g);
    }
    return f;
}

}  // namespace

// ===========================================================================
// aes_kexp on the device: one thread per key (tb/key_exp.py:118 semantics;
// output bytes identical to its list, stage r at bytes 16r..16r+15).
// ===========================================================================
__global__ void k_key_expand(const uint8_t* __restrict__ keys, uint64_t n_keys, int key_bytes,
                             const uint32_t* __restrict__ te0, uint8_t* __restrict__ round_keys)
{
    __shared__ uint32_t sbox_s[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sbox_s[i] = (__ldg(te0 + i) >> 8) & 0xff;
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_keys) return;
    uint8_t key[32];
    for (int j = 0; j < key_bytes; ++j) key[j] = keys[i * key_bytes + j];
    uint32_t rk[60];
    auto sb = [&](uint32_t b) { return sbox_s[b & 0xff]; };
    const int nr = aes_key_expand_words(key, key_bytes, sb, rk);
    uint32_t* dst = reinterpret_cast<uint32_t*>(round_keys + i * (uint64_t)(16 * (nr + 1)));
    for (int j = 0; j < 4 * (nr + 1); ++j) dst[j] = rk[j];  // LE words == the byte string
}

// ===========================================================================
// Per-key setup: stage keys, H, H^(2^k), per-thread / per-CTA weights, Shoup
// tables.  One CTA of 256 threads, one launch per agcm_set_key: the key (raw,
// or the Nr+1 user-loaded stages of config/config_aes_kprexp.py:66-95) rides
// in the kernel parameters, so the host does no H2D copy.
// ===========================================================================
__global__ void __launch_bounds__(256) k_key_setup(KeyDev* kd, const __grid_constant__ KeyIn in,
                                                   const uint32_t* __restrict__ te0, uint32_t nt_stream, uint32_t ncta)
{
    const uint32_t tid = threadIdx.x;
    __shared__ gf128 s_pow2[64];
    __shared__ uint32_t s_rk[60];
    if (in.pre_expanded) {
        if (tid < 4 * (in.nr + 1)) s_rk[tid] = in.w[tid];
    } else if (tid == 0) {
        uint8_t key[32];
        for (uint32_t j = 0; j < in.key_bytes; ++j) key[j] = (uint8_t)(in.w[j >> 2] >> (8 * (j & 3)));
        auto sb = [&](uint32_t b) { return (__ldg(te0 + (b & 0xff)) >> 8) & 0xff; };
        aes_key_expand_words(key, (int)in.key_bytes, sb, s_rk);
    }
    __syncthreads();
    if (tid < 60) kd->rk[tid] = tid < 4 * (in.nr + 1) ? s_rk[tid] : 0u;
    if (tid == 0) {
        TeGlobal te{te0};
        uint32_t h[4];
        aes_encrypt_words(s_rk, (int)in.nr, 0, 0, 0, 0, te, h);  // H = E_K(0^128), gcm_gctr.vhd:141-144
        gf128 p = gf_from_le_words(h[0], h[1], h[2], h[3]);
        kd->H = p;
        kd->nr = in.nr;
        kd->nt_stream = nt_stream;
        kd->ncta = ncta;
        for (int k = 0; k < 64; ++k) {
            s_pow2[k] = p;
            kd->pow2[k] = p;
            p = gf_sqr(p);
        }
        kd->hpow_thread[0] = gf_one();
        kd->hpow_thread[1] = s_pow2[0];
        kd->hpow_cta[0] = gf_one();
    }
    __syncthreads();
    // H^k for k = 2..nt_stream by doubling: H^(2^s + j) = H^j * H^(2^s), j = 1..2^s
    for (uint32_t s = 0; (1u << s) < nt_stream; ++s) {
        const uint32_t half = 1u << s;
        for (uint32_t j = tid; j < half; j += blockDim.x) {
            const uint32_t dst = half + 1 + j;
            if (dst <= nt_stream) kd->hpow_thread[dst] = gf_mul(kd->hpow_thread[1 + j], s_pow2[s]);
        }
        __syncthreads();
    }
    // (H^NT)^k for k = 1..ncta, same doubling with base powers pow2[log2(NT) + s]
    uint32_t lg = 0;
    while ((1u << lg) < nt_stream) ++lg;
    if (tid == 0) kd->hpow_cta[1] = s_pow2[lg];
    __syncthreads();
    for (uint32_t s = 0; (1u << s) < ncta; ++s) {
        const uint32_t half = 1u << s;
        for (uint32_t j = tid; j < half; j += blockDim.x) {
            const uint32_t dst = half + 1 + j;
            if (dst <= ncta) kd->hpow_cta[dst] = gf_mul(kd->hpow_cta[1 + j], s_pow2[lg + s]);
        }
        __syncthreads();
    }
    // Shoup tables, row b per thread (blockDim.x == 256)
    for (int j = 0; j < 8; ++j) {
        gf128 c;
        if (j < 6) c = s_pow2[j];
        else if (j == 6) c = s_pow2[lg];
        else c = kd->hpow_cta[ncta];
        gf128 basis[8];
        basis[0] = c;
#pragma unroll
        for (int k = 1; k < 8; ++k) basis[k] = gf_mulx(basis[k - 1]);
        for (uint32_t b = tid; b < 256; b += blockDim.x) kd->tab[j][b] = gf_table_row(basis, b);
    }
}

// One full warp.  s = xor of the shard partials (natural GHASH domain, BE words).
struct FinishArgs {
    const uint32_t* rk;
    uint32_t nr;
    const uint32_t* iv;
    uint32_t j0w;         // counter word of J0 (byte-swapped)
    const KeyDev* key;
    const uint32_t* te0;
    const uint8_t* aad;
    uint64_t aad_len, ct_len;
    uint8_t* tag_calc;
    const uint8_t* tag_expected;
    uint8_t* ok;
    const uint32_t* hn;   // H^(ct blocks) precomputed by k_pow (4 BE words), or null: compute here
    bool smem_tables;     // the CTA's shared T-tables are valid (called from k_stream)
};

__device__ void finish_warp(const FinishArgs& p, gf128 s)
{
    const uint32_t lane = threadIdx.x & 31;
    const KeyDev* kd = p.key;
    const uint64_t n = (p.ct_len + 15) >> 4;
    if (p.aad && p.aad_len) {  // uniform
        // short AAD (host routes long AAD through k_stream<GHASH_ONLY>): the warp
        // runs the same front-padded strided Horner with G = 32 and generic products.
        const uint64_t a = (p.aad_len + 15) >> 4;
        const uint64_t rows = (a + 31) >> 5, pad = rows * 32 - a;
        const gf128 h32 = kd->pow2[5];
        gf128 qa = gf_zero();
        for (uint64_t u = 0; u < rows; ++u) {
            const uint64_t v = u * 32 + lane;
            if (u) qa = gf_mul(qa, h32);
            if (v >= pad) {
                const uint64_t i = v - pad;
                const uint64_t left = p.aad_len - 16 * i;
                uint32_t x[4];
                ag_load_block(p.aad + 16 * i, left < 16 ? (uint32_t)left : 16u, x);
                qa = gf_xor(qa, gf_from_le_words(x[0], x[1], x[2], x[3]));
            }
        }
        qa = gf_mul(qa, kd->hpow_thread[32 - lane]);
        qa = warp_xor(qa);                     // QA = sum A_i H^(a-i)
        gf128 hn;
        if (p.hn) {
            hn.w[0] = __ldcg(p.hn + 0); hn.w[1] = __ldcg(p.hn + 1); hn.w[2] = __ldcg(p.hn + 2); hn.w[3] = __ldcg(p.hn + 3);
        } else {
            hn = warp_gf_pow(kd, n);
        }
        if (lane == 0) s = gf_xor(s, gf_mul(qa, hn));
    }
    if (lane == 0) {
        const uint64_t ab = p.aad_len * 8, cb = p.ct_len * 8;
        s.w[0] ^= (uint32_t)(ab >> 32); s.w[1] ^= (uint32_t)ab; s.w[2] ^= (uint32_t)(cb >> 32); s.w[3] ^= (uint32_t)cb;
        s = gf_mul(s, kd->H);
        uint32_t e[4];
        if (p.smem_tables) {
            TeSmem te{ag_smem, 0};
            aes_encrypt_words(p.rk, (int)p.nr, p.iv[0], p.iv[1], p.iv[2], p.j0w, te, e);  // J0 (= IV || 00000001 for a 96-bit IV)
        } else {
            TeGlobal te{p.te0};
            aes_encrypt_words(p.rk, (int)p.nr, p.iv[0], p.iv[1], p.iv[2], p.j0w, te, e);
        }
        uint32_t t[4] = {ag_bswap32(s.w[0]) ^ e[0], ag_bswap32(s.w[1]) ^ e[1], ag_bswap32(s.w[2]) ^ e[2],
                         ag_bswap32(s.w[3]) ^ e[3]};
        ag_store_block(p.tag_calc, 16, t);
        if (p.tag_expected && p.ok) {
            uint32_t x[4];
            ag_load_block(p.tag_expected, 16, x);
            const uint32_t diff = (x[0] ^ t[0]) | (x[1] ^ t[1]) | (x[2] ^ t[2]) | (x[3] ^ t[3]);
            *p.ok = diff ? 0 : 1;
        }
    }
}

// One tiny all-to-all over NVLink peer memory instead of a collective library call (one warp):
// lane w stores this rank's scaled partial into its slot of rank w's exchange buffer, fences at
// system scope and raises the slot's epoch flag.  Nobody waits here: the bulk kernel of the next
// message may start while the flags of this one are still in flight (k_peer_finish, on a side
// stream, is the only waiter).
__device__ __forceinline__ void ag_peer_post(uint8_t* const* peer_bufs, uint32_t rank, uint32_t world, uint32_t epoch,
                                             const gf128& s)
{
    const uint32_t w = threadIdx.x & 31, slot = epoch % AG_PEER_RING;
    if (w < world) {
        uint8_t* dst = peer_bufs[w];
        __stcg(reinterpret_cast<uint4*>(dst + (slot * AG_PEER_MAX + rank) * 16), make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]));
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(dst + AG_PEER_FLAGS + (slot * AG_PEER_MAX + rank) * 4) = epoch;
    }
}

// Posts a partial computed by earlier launches (the host-buffer pipeline: 16 B in natural GHASH
// byte order) or, with a null pointer, the zero partial of a rank whose counter range is empty
// (fewer blocks than ranks): such a rank still takes part in the exchange.
__global__ void __launch_bounds__(32) k_peer_post(uint8_t* const* peer_bufs, uint32_t rank, uint32_t world, uint32_t epoch,
                                                  const uint8_t* __restrict__ partial16)
{
    gf128 s = gf_zero();
    if (partial16) {
        uint32_t x[4];
        ag_load_block(partial16, 16, x);
        s = gf_from_le_words(x[0], x[1], x[2], x[3]);
    }
    ag_peer_post(peer_bufs, rank, world, epoch, s);
}

// Waits (bounded by the global timer) for the world's flags of `epoch` in this rank's own buffer,
// XORs the slots and finishes the tag.  FAILS CLOSED: if a peer never shows up the tag is
// zeroed, ok = 0, and both status words are raised (the host-mapped one turns every later peer
// call into AGCM_E_PEER_TIMEOUT).
__global__ void __launch_bounds__(32) k_peer_finish(const __grid_constant__ PeerFinishParams p)
{
    const uint32_t w = threadIdx.x, slot = p.epoch % AG_PEER_RING;
    const uint8_t* mine = p.peer_bufs[p.rank];
    bool arrived = true;
    if (w < p.world) {
        const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(mine + AG_PEER_FLAGS + (slot * AG_PEER_MAX + w) * 4);
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (*f != p.epoch) {
            __nanosleep(64);
            if ((++spins & 255u) == 0) {
                uint64_t now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (!t0) t0 = now;
                else if (now - t0 > p.timeout_ns) { arrived = false; break; }
            }
        }
    }
    __threadfence_system();
    gf128 t = gf_zero();
    if (w < p.world) {
        const uint4 q = __ldcv(reinterpret_cast<const uint4*>(mine + (slot * AG_PEER_MAX + w) * 16));
        t.w[0] = q.x; t.w[1] = q.y; t.w[2] = q.z; t.w[3] = q.w;
    }
    const gf128 s = warp_xor(t);
    if (!__all_sync(0xffffffffu, arrived)) {
        if (w == 0) {
            const uint32_t z[4] = {0, 0, 0, 0};
            ag_store_block(p.f.tag_calc, 16, z);
            if (p.f.ok) *p.f.ok = 0;
            if (p.status_dev) *p.status_dev = 1;
            if (p.status_host) *p.status_host = 1;
            __threadfence_system();
        }
        return;
    }
    FinishArgs a{p.f.rk, p.f.nr, p.f.iv, p.f.j0w, p.f.key, p.f.te0, p.f.aad, p.f.aad_len, p.f.ct_len, p.f.tag_calc,
                 p.f.tag_expected, p.f.ok, p.f.hn, false};
    finish_warp(a, s);
}

__global__ void __launch_bounds__(32) k_stream_finish(const __grid_constant__ FinishParams p)
{
    const uint32_t lane = threadIdx.x;
    gf128 s = gf_zero();
    for (uint32_t j = lane; j < p.n_parts; j += 32) {
        uint32_t x[4];
        ag_load_block(p.parts + 16 * j, 16, x);
        s = gf_xor(s, gf_from_le_words(x[0], x[1], x[2], x[3]));
    }
    s = warp_xor(s);
    FinishArgs a{p.rk, p.nr, p.iv, p.j0w, p.key, p.te0, p.aad, p.aad_len, p.ct_len, p.tag_calc, p.tag_expected, p.ok, p.hn, false};
    finish_warp(a, s);
}

// ===========================================================================
// Single stream: fused GCTR + GHASH, grid-wide strided Horner.
// grid = kd->ncta CTAs x kd->nt_stream threads; writes one 16 B partial per CTA.
// ===========================================================================
template <int NR, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(AG_STREAM_NT_MAX, 1) k_stream(const __grid_constant__ StreamParams p)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    if (MODE == AG_MODE_CTR_ONLY && p.gate && __ldcg(p.gate) == 0) return;   // tag did not verify: release nothing
    if (MODE != AG_MODE_GHASH_ONLY) stage_te0(p.te0);
    if (MODE != AG_MODE_CTR_ONLY) fill_gh_tables(p.key->tab[7], nullptr);
    __syncthreads();
    if (MODE != AG_MODE_GHASH_ONLY) expand_aes_tables();
    __syncthreads();

    TeSmem te{ag_smem, lane * 4};
    GhSmem gh{ag_smem + SM_GH, (lane & 7) * 16};
    const uint32_t Gt = gridDim.x * nt;
    const uint32_t g = blockIdx.x * nt + tid;
    gf128 y = ag_stream_lane<NR, MODE, ALIGNED>(p, g, Gt, te, gh);
    if (MODE == AG_MODE_CTR_ONLY) return;

    // lane weight H^(Gt-g) = (H^NT)^(ncta-1-cta) * H^(NT-tid); lanes that absorbed nothing (short
    // messages leave most of the grid idle) skip the bit-serial product
    if (__any_sync(0xffffffffu, (y.w[0] | y.w[1] | y.w[2] | y.w[3]) != 0)) y = gf_mul(y, p.key->hpow_thread[nt - tid]);
    y = warp_xor(y);
    gf128* red = reinterpret_cast<gf128*>(ag_smem + SM_MISC);
    if (lane == 0) red[tid >> 5] = y;
    __syncthreads();
    if (tid < 32) {
        gf128 v = (tid < (nt >> 5)) ? red[tid] : gf_zero();
        v = warp_xor(v);
        if (tid == 0) {
            v = gf_mul(v, p.key->hpow_cta[gridDim.x - 1 - blockIdx.x]);
            uint32_t* dst = p.partials + 4 * blockIdx.x;
            dst[0] = v.w[0]; dst[1] = v.w[1]; dst[2] = v.w[2]; dst[3] = v.w[3];
        }
        // The last CTA to get here folds the per-CTA partials (and, for a single-shard
        // message, finishes the tag) so that one launch does the whole message.
        uint32_t ticket = 0;
        if (p.done_counter) {
            if (tid == 0) {
                __threadfence();
                ticket = atomicAdd(p.done_counter, 1u);
            }
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket == gridDim.x - 1) {
                __threadfence();
                gf128 s = gf_zero();
                for (uint32_t j = tid; j < gridDim.x; j += 32) {
                    const uint4 q = __ldcg(reinterpret_cast<const uint4*>(p.partials) + j);
                    s.w[0] ^= q.x; s.w[1] ^= q.y; s.w[2] ^= q.z; s.w[3] ^= q.w;
                }
                s = warp_xor(s);
                if (p.scale_e) {
                    gf128 he;
                    if (p.scale_pow) {
                        he.w[0] = __ldcg(p.scale_pow + 0); he.w[1] = __ldcg(p.scale_pow + 1);
                        he.w[2] = __ldcg(p.scale_pow + 2); he.w[3] = __ldcg(p.scale_pow + 3);
                    } else {
                        he = warp_gf_pow(p.key, p.scale_e);
                    }
                    s = gf_mul(s, he);
                }
                if (p.out16 && tid == 0) {
                    const uint32_t o[4] = {ag_bswap32(s.w[0]), ag_bswap32(s.w[1]), ag_bswap32(s.w[2]), ag_bswap32(s.w[3])};
                    ag_store_block(p.out16, 16, o);  // natural GHASH byte order
                }
                if (p.peer_world) ag_peer_post(p.peer_bufs, p.peer_rank, p.peer_world, p.peer_epoch, s);
                if (p.fuse_finish) {
                    FinishArgs a{p.rk, (uint32_t)NR, p.iv, p.j0w, p.key, p.te0, p.aad, p.aad_len, p.ct_len, p.tag_calc,
                                 p.tag_expected, p.ok, p.hn, MODE != AG_MODE_GHASH_ONLY};
                    finish_warp(a, s);
                }
                if (tid == 0) *p.done_counter = 0;  // ready for the next launch on this context
            }
        }
    }
}

// out[0..3] = H^e (BE words); cached per (key, e) by the host so that the tag finish
// does not recompute it for every message of the same length
__global__ void __launch_bounds__(32) k_pow(const KeyDev* kd, uint64_t e, uint32_t* __restrict__ out)
{
    const gf128 r = warp_gf_pow(kd, e);
    if (threadIdx.x == 0) { out[0] = r.w[0]; out[1] = r.w[1]; out[2] = r.w[2]; out[3] = r.w[3]; }
}

// J0 per message for IVs of any length (SP 800-38D 7.1 step 2), one thread per IV: a 96-bit IV
// gives IV || 0^31 1 (src/aes_icb.vhd:34,118), any other length GHASH_H(IV || 0^(s+64) || [len(IV)]_64)
// with the serial recurrence of src/gcm_ghash.vhd:269-272.  16 bytes out per message.
__global__ void k_batch_j0(const KeyDev* kd, const uint8_t* __restrict__ iv, const uint64_t* __restrict__ iv_off,
                           uint64_t iv_len, uint64_t n, uint8_t* __restrict__ j0)
{
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const uint64_t off = iv_off ? iv_off[m] : m * iv_len;
    const uint64_t len = iv_off ? iv_off[m + 1] - off : iv_len;
    const uint8_t* p = iv + off;
    uint8_t* dst = j0 + 16 * m;
    if (len == 12) {
        for (int j = 0; j < 12; ++j) dst[j] = p[j];
        dst[12] = dst[13] = dst[14] = 0;
        dst[15] = 1;
        return;
    }
    const gf128 h = kd->H;
    gf128 y = gf_zero();
    for (uint64_t o = 0; o < len; o += 16) {
        uint32_t x[4];
        ag_load_block(p + o, (len - o) < 16 ? (uint32_t)(len - o) : 16u, x);
        y = gf_mul(gf_xor(y, gf_from_le_words(x[0], x[1], x[2], x[3])), h);
    }
    const uint64_t bits = len * 8;
    y.w[2] ^= (uint32_t)(bits >> 32);
    y.w[3] ^= (uint32_t)bits;
    y = gf_mul(y, h);
    const uint32_t o4[4] = {ag_bswap32(y.w[0]), ag_bswap32(y.w[1]), ag_bswap32(y.w[2]), ag_bswap32(y.w[3])};
    ag_store_block(dst, 16, o4);
}

// out16 = xor of n 16-byte partials (natural byte order in and out)
__global__ void __launch_bounds__(32) k_xor_parts(const uint8_t* __restrict__ parts, uint32_t n, uint8_t* __restrict__ out)
{
    const uint32_t lane = threadIdx.x;
    gf128 s = gf_zero();
    for (uint32_t j = lane; j < n; j += 32) {
        uint32_t x[4];
        ag_load_block(parts + 16 * j, 16, x);
        s.w[0] ^= x[0]; s.w[1] ^= x[1]; s.w[2] ^= x[2]; s.w[3] ^= x[3];
    }
    s = warp_xor(s);
    if (lane == 0) ag_store_block(out, 16, s.w);
}

// Tag finish (gcm_ghash.vhd:257,293 + tb/gcm_model.py:33-51):
//   S = xor parts (each already aligned so that the last CT block weighs H^1)
//   QA = sum A_i H^(a-i) over the (short) AAD given here, if any
//   TAG = ((QA * H^n) xor S xor LEN) * H xor E_K(J0)
// decrypt: constant-time compare with the expected tag -> ok flag; the computed
// tag is always written to tag_calc.

// ===========================================================================
// Batched messages under one shared key: G lanes per message, persistent grid.
// Tables: T_a = H^G (row Horner), T_b = H (lane combine).
// ===========================================================================
template <int NR, bool DEC, int G>
__global__ void __launch_bounds__(AG_STREAM_NT_MAX, 1) k_batch(const __grid_constant__ BatchParams p)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    constexpr int LG = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : 5;
    stage_te0(p.te0);
    fill_gh_tables(p.key->tab[LG], p.key->tab[0]);
    __syncthreads();
    expand_aes_tables();
    __syncthreads();

    TeSmem te{ag_smem, lane * 4};
    GhSmem gh_g{ag_smem + SM_GH, (lane & 7) * 16};
    GhSmem gh_1{ag_smem + SM_GH + 128, (lane & 7) * 16};

    const uint32_t t = lane & (G - 1);
    const uint32_t gbase = lane & ~(uint32_t)(G - 1);
    gf128 lane_weight = gf_one();
    if (G >= 16) lane_weight = p.key->hpow_thread[G - t];   // H^(G-t), G <= nt_stream
    const uint64_t groups_per_cta = nt / G;
    const uint64_t n_groups = (uint64_t)gridDim.x * groups_per_cta;
    const uint64_t gid = (uint64_t)blockIdx.x * groups_per_cta + tid / G;
    // every warp runs the same trip count; lanes past the end are predicated off
    const uint64_t warp_first = (uint64_t)blockIdx.x * groups_per_cta + (tid & ~31u) / G;
    for (uint64_t w0 = warp_first, m = gid; w0 < p.n_msgs; w0 += n_groups, m += n_groups) {
        const bool valid = m < p.n_msgs;
        gf128 y = gf_zero();
        AesCtrConst cc;
        AesCtrSeqCache cache;
        cache.key = 0xFFFFFFFFu;  // invalid: the key only ever holds 24 bits
        uint32_t e[4] = {0, 0, 0, 0};  // E_K(J0): produced by lane G-1 (the one that meets the length block)
        if (valid) {
            MsgDesc d = ag_batch_msg(p, m);
            uint32_t iv0, iv1, iv2;
            if (!p.iv_is_j0 && ((uintptr_t)(p.iv + 12 * m) & 3) == 0) {
                const uint32_t* q = reinterpret_cast<const uint32_t*>(p.iv + 12 * m);
                iv0 = q[0]; iv1 = q[1]; iv2 = q[2];
            } else {
                uint32_t ivw[3];
                ag_batch_iv(p, m, ivw, &d.j0ctr);
                iv0 = ivw[0]; iv1 = ivw[1]; iv2 = ivw[2];
            }
            cc = aes_ctr_precompute(p.rk, iv0, iv1, iv2, te);
            y = ag_batch_lane<NR, DEC>(p.rk, cc, cache, d, t, (uint32_t)G, te, gh_g, e);
        }
        __syncwarp();
        // R = sum_t Y_t H^(G-t)
        gf128 r = gf_zero();
        if (G >= 16) {
            // wide groups: every lane applies its own weight with one generic product (integer
            // pipe only), then a butterfly XOR -- G-1 serial table products would keep the lookup
            // pipe, the binding one, busy for G x 64 wavefronts per warp
            r = gf_mul(y, lane_weight);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                r.w[0] ^= __shfl_xor_sync(0xffffffffu, r.w[0], o);
                r.w[1] ^= __shfl_xor_sync(0xffffffffu, r.w[1], o);
                r.w[2] ^= __shfl_xor_sync(0xffffffffu, r.w[2], o);
                r.w[3] ^= __shfl_xor_sync(0xffffffffu, r.w[3], o);
            }
        } else {
            // narrow groups: serial Horner over the group's lanes with T_b = H
#pragma unroll 1
            for (int k = 0; k < G; ++k) {
                gf128 yk;
                yk.w[0] = __shfl_sync(0xffffffffu, y.w[0], gbase + k);
                yk.w[1] = __shfl_sync(0xffffffffu, y.w[1], gbase + k);
                yk.w[2] = __shfl_sync(0xffffffffu, y.w[2], gbase + k);
                yk.w[3] = __shfl_sync(0xffffffffu, y.w[3], gbase + k);
                r = gf_xor(r, yk);
                r = gf_mul_table(r, gh_1);
            }
        }
        if (valid && t == G - 1) {
            uint32_t tg[4] = {ag_bswap32(r.w[0]) ^ e[0], ag_bswap32(r.w[1]) ^ e[1], ag_bswap32(r.w[2]) ^ e[2],
                              ag_bswap32(r.w[3]) ^ e[3]};
            uint8_t* tp = p.tag + 16 * m;
            if (DEC) {
                uint32_t x[4];
                ag_load_block(tp, 16, x);
                const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
                p.ok[m] = diff ? 0 : 1;
            } else {
                ag_store_block(tp, 16, tg);
            }
        }
        __syncwarp();
    }
}

// ===========================================================================
// Batched FIXED-SIZE records under the shared key, one lane per message, records staged through
// shared memory by TMA (BASELINE config 3: 2^20 x 1500 B at a 1504 B stride).
//
// Why: with a message per lane the compute layout is the cheapest there is (no lane combine, no
// front padding, the Horner constant is H itself), but every 128-bit global load/store of a warp
// touches 32 different lines -- ncu: ~48 of ~277 L1/shared data-pipe wavefronts per 32 blocks
// (profiles/r1_ncu_batch.md), on the pipe that binds the kernel.  Here the batch is a 2-D tensor
// [message][byte] (row pitch = the record stride); one elected lane per warp asks the TMA unit for
// the box {32 bytes x 32 messages}: two blocks of each of the warp's 32 messages land as a dense,
// 32B-swizzled 1 KB tile (conflict-free LDS.128/STS.128: 4 + 4 wavefronts per 32 blocks), the lanes
// XOR the keystream in place, and the same box goes back with a TMA store.  The async proxy moves
// the bytes; the LSU pipe only sees the tile accesses.  Both tensors have 16-byte-granular extents:
// the load tensor covers the record rounded UP to whole blocks (it reaches into the caller's
// padding, which the lane masks off: src/gcm_ghash.vhd:228-246), the store tensor the record
// rounded DOWN (a ragged tail leaves by byte stores of its own lane); rows past the last message
// arrive as zeros and are clipped on the way out.
// Two tiles per warp (load of tile t+1 in flight while tile t is processed); groups of 32
// messages are handed out by an atomic ticket, so no warp idles while another still has a queue.
// ===========================================================================
namespace {
constexpr uint32_t SM_TILE_BAR = SM_MISC + 1024;          // 16 warps x 2 mbarriers
constexpr uint32_t SM_TILE = SM_MISC + 2048;              // 16 warps x 2 tiles x 1 KB
constexpr uint32_t TILE_BYTES = 1024;                     // 32 messages x 32 bytes
constexpr size_t kTileSmemBytes = SM_TILE + (AG_STREAM_NT_MAX / 32) * 2 * TILE_BYTES;
}  // namespace

template <int NR, bool DEC>
__global__ void __launch_bounds__(AG_STREAM_NT_MAX, 1) k_batch_tile(const __grid_constant__ TileParams P)
{
    const BatchParams& p = P.b;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_te0(p.te0);
    fill_gh_tables(p.key->tab[0], nullptr);   // Horner constant H (a message per lane)
    const uint32_t bar0 = ag_smem_addr(ag_smem + SM_TILE_BAR + warp * 16), bar1 = bar0 + 8;
    if (lane == 0) {
        ag_mbar_init(bar0, 1);
        ag_mbar_init(bar1, 1);
        ag_fence_barrier_init();
        ag_prefetch_tmap(&P.tm_in);
        ag_prefetch_tmap(&P.tm_out);
        if (P.aad_tiled) ag_prefetch_tmap(&P.tm_aad);
    }
    __syncthreads();
    expand_aes_tables();
    __syncthreads();

    TeSmem te{ag_smem, lane * 4};
    GhSmem gh{ag_smem + SM_GH, (lane & 7) * 16};
    uint8_t* tiles = ag_smem + SM_TILE + warp * (2 * TILE_BYTES);
    const uint32_t tile_sa = ag_smem_addr(tiles);
    // this lane's two 16-byte chunks inside a tile (CU_TENSOR_MAP_SWIZZLE_32B: address bit 4 ^= bit 7)
    const uint32_t sw = (lane >> 2) & 1;
    const uint32_t coff0 = lane * 32 + ((0 ^ sw) << 4), coff1 = lane * 32 + ((1 ^ sw) << 4);
    uint32_t par0 = 0, par1 = 0;

    const uint32_t n_blocks = (uint32_t)((p.len + 15) >> 4), tail = (uint32_t)(p.len & 15), n_full = (uint32_t)(p.len >> 4);
    const uint32_t n_tiles = (n_blocks + 1) >> 1;
    const uint32_t a_blocks = (uint32_t)((p.aad_len + 15) >> 4), atail = (uint32_t)(p.aad_len & 15);
    // the unified sequence AAD | CT (gcm_ghash.vhd:259-272) as ONE stream of tiles through the two buffers
    const uint32_t a_tiles = P.aad_tiled ? (a_blocks + 1) >> 1 : 0;
    const uint32_t tot_tiles = a_tiles + n_tiles;
    const uint32_t n_groups = (uint32_t)((p.n_msgs + 31) >> 5);
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(P.ticket, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_groups) break;
        const int32_t row0 = (int32_t)(g * 32);
        auto issue = [&](uint32_t T, uint32_t buf) {   // lane 0 only
            const uint32_t bar = buf ? bar1 : bar0;
            ag_mbar_expect_tx(bar, TILE_BYTES);
            if (T < a_tiles) ag_tma_load_2d(tile_sa + buf * TILE_BYTES, &P.tm_aad, (int32_t)(T * 32), row0, bar);
            else ag_tma_load_2d(tile_sa + buf * TILE_BYTES, &P.tm_in, (int32_t)((T - a_tiles) * 32), row0, bar);
        };
        if (lane == 0 && tot_tiles) issue(0, 0);
        const uint64_t m_raw = (uint64_t)g * 32 + lane;
        const bool valid = m_raw < p.n_msgs;
        const uint64_t m = valid ? m_raw : p.n_msgs - 1;   // idle lanes of the last group shadow a real message
        uint32_t ivw[3], j0ctr;
        ag_batch_iv(p, m, ivw, &j0ctr);
        const AesCtrConst cc = aes_ctr_precompute(p.rk, ivw[0], ivw[1], ivw[2], te);
        AesCtrSeqCache cache;
        cache.key = 0xFFFFFFFFu;
        gf128 y = gf_zero();
        if (a_blocks && !P.aad_tiled) {   // AAD the TMA cannot address (alignment): read in place
            const uint8_t* ap = p.aad + m * p.aad_stride;
            for (uint32_t i = 0; i < a_blocks; ++i) {
                uint32_t x[4];
                ag_load_block(ap + 16 * (uint64_t)i, (i == a_blocks - 1 && atail) ? atail : 16u, x);
                y.w[0] ^= ag_bswap32(x[0]); y.w[1] ^= ag_bswap32(x[1]); y.w[2] ^= ag_bswap32(x[2]); y.w[3] ^= ag_bswap32(x[3]);
                y = gf_mul_table(y, gh);
            }
        }
        for (uint32_t T = 0; T < tot_tiles; ++T) {
            const uint32_t b = T & 1;
            if (lane == 0 && T + 1 < tot_tiles) {
                ag_bulk_wait_read0();   // the store of tile T-1 has read the buffer tile T+1 lands in
                issue(T + 1, b ^ 1);
            }
            if (b) { ag_mbar_wait(bar1, par1); par1 ^= 1; } else { ag_mbar_wait(bar0, par0); par0 ^= 1; }
            uint8_t* tb = tiles + b * TILE_BYTES;
            if (T < a_tiles) {   // uniform: an AAD tile is absorbed, nothing goes back
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint32_t j = 2 * T + k;
                    if (j < a_blocks) {
                        const uint4 xv = *reinterpret_cast<const uint4*>(tb + (k ? coff1 : coff0));
                        uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                        if (j == a_blocks - 1 && atail) ag_mask_block(x, atail);
                        y.w[0] ^= ag_bswap32(x[0]); y.w[1] ^= ag_bswap32(x[1]); y.w[2] ^= ag_bswap32(x[2]); y.w[3] ^= ag_bswap32(x[3]);
                        y = gf_mul_table(y, gh);
                    }
                }
                __syncwarp();
                continue;
            }
            const uint32_t t = T - a_tiles;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t j = 2 * t + k;
                if (j < n_blocks) {   // uniform
                    uint4* cp = reinterpret_cast<uint4*>(tb + (k ? coff1 : coff0));
                    const uint4 xv = *cp;
                    uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                    const bool ragged = (j == n_full);   // the record's short last block (uniform)
                    if (ragged) ag_mask_block(x, tail);  // the load box reaches into the caller's padding
                    uint32_t ks[4];
                    aes_ctr_block_seq<NR>(p.rk, cc, cache, j0ctr + 1u + j, te, ks);
                    uint32_t o[4] = {x[0] ^ ks[0], x[1] ^ ks[1], x[2] ^ ks[2], x[3] ^ ks[3]};
                    if (!ragged) {
                        *cp = make_uint4(o[0], o[1], o[2], o[3]);
                    } else {
                        // the store tensor ends at the last WHOLE block: the tail goes out byte-wise, once
                        // per message, so that the padding between records is never written
                        if (valid) ag_store_block(p.out + m * p.stride + 16 * (uint64_t)j, tail, o);
                        ag_mask_block(o, tail);
                    }
                    if (DEC) {
                        y.w[0] ^= ag_bswap32(x[0]); y.w[1] ^= ag_bswap32(x[1]); y.w[2] ^= ag_bswap32(x[2]); y.w[3] ^= ag_bswap32(x[3]);
                    } else {
                        y.w[0] ^= ag_bswap32(o[0]); y.w[1] ^= ag_bswap32(o[1]); y.w[2] ^= ag_bswap32(o[2]); y.w[3] ^= ag_bswap32(o[3]);
                    }
                    y = gf_mul_table(y, gh);
                }
            }
            if (2 * t < n_full) {   // uniform: the tile holds at least one whole block
                ag_fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ag_tma_store_2d(&P.tm_out, (int32_t)(t * 32), row0, tile_sa + b * TILE_BYTES);
                    ag_bulk_commit();
                }
            } else {
                __syncwarp();
            }
        }
        // length block (gcm_ghash.vhd:257), last multiply, E_K(J0) (gcm_ghash.vhd:293)
        {
            const uint64_t ab = p.aad_len * 8, cb = p.len * 8;
            y.w[0] ^= (uint32_t)(ab >> 32); y.w[1] ^= (uint32_t)ab; y.w[2] ^= (uint32_t)(cb >> 32); y.w[3] ^= (uint32_t)cb;
            y = gf_mul_table(y, gh);
            uint32_t e[4];
            aes_ctr_block_seq<NR>(p.rk, cc, cache, j0ctr, te, e);
            const uint32_t tg[4] = {ag_bswap32(y.w[0]) ^ e[0], ag_bswap32(y.w[1]) ^ e[1], ag_bswap32(y.w[2]) ^ e[2],
                                    ag_bswap32(y.w[3]) ^ e[3]};
            if (valid) {
                uint8_t* tp = p.tag + 16 * m;
                if (DEC) {
                    uint32_t x[4];
                    ag_load_block(tp, 16, x);
                    const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
                    p.ok[m] = diff ? 0 : 1;
                } else {
                    ag_store_block(tp, 16, tg);
                }
            }
        }
        if (lane == 0) ag_bulk_wait_read0();   // both tiles are free again for the next group
        __syncwarp();
    }
    if (lane == 0) ag_bulk_wait0();
}

// ===========================================================================
// Batched LONG messages under the shared key: one CTA per message (G = blockDim.x
// lanes).  Same front-padded strided Horner as k_batch, constant H^NT (tab[6]); lane
// weights H^(NT-tid) by one bit-serial product per lane per message, then a CTA
// XOR-reduce.  Used when messages are too few to give every warp its own message.
// ===========================================================================
template <int NR, bool DEC>
__global__ void __launch_bounds__(AG_STREAM_NT_MAX, 1) k_batch_cta(const __grid_constant__ BatchParams p)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    stage_te0(p.te0);
    fill_gh_tables(p.key->tab[6], nullptr);
    __syncthreads();
    expand_aes_tables();
    __syncthreads();
    TeSmem te{ag_smem, lane * 4};
    GhSmem gh_g{ag_smem + SM_GH, (lane & 7) * 16};
    gf128* red = reinterpret_cast<gf128*>(ag_smem + SM_MISC);
    const gf128 wgt = p.key->hpow_thread[nt - tid];
    // H^after cache of the split layout, 32 direct-mapped entries (key 0 = empty): with equal-length
    // messages a CTA meets S / gcd(gridDim, S) distinct segment indices (S/4 on 148 CTAs, 4 apart),
    // and one exponentiation by a single warp (seven dependent generic products, the rest of the
    // CTA waiting) costs ~20 us
    uint64_t* pow_keys = reinterpret_cast<uint64_t*>(ag_smem + SM_MISC + 1056);   // 32 x 8 B
    gf128* pow_vals = reinterpret_cast<gf128*>(ag_smem + SM_MISC + 1312);         // 32 x 16 B
    if (tid < 32) pow_keys[tid] = 0;
    __syncthreads();
    // One unit per CTA pass: a whole message, or (split > 1) one counter-range segment of it --
    // a few long messages would otherwise leave the grid idle in the last round (256 messages on
    // 148 CTAs: 2 rounds, 86 % busy).  A segment's partial is scaled by H^(blocks after it), the
    // S partials and E_K(J0) go to seg_parts, and k_batch_split_finish XORs them into the tag.
    const uint32_t S = p.split;
    const uint64_t n_units = p.n_msgs * S;
    // units go to whichever CTA is free next (atomic ticket): with static round-robin every CTA
    // waits for the one that drew the most (or the longest) units
    uint32_t* s_unit = reinterpret_cast<uint32_t*>(ag_smem + SM_MISC + 1824);
    for (uint64_t it = 0;; ++it) {
        uint64_t u;
        if (p.ticket) {
            if (tid == 0) *s_unit = atomicAdd(p.ticket, 1u);
            __syncthreads();
            u = *s_unit;
        } else {
            u = blockIdx.x + it * gridDim.x;
        }
        if (u >= n_units) break;
        const uint64_t m = u / S;
        const uint32_t seg = (uint32_t)(u - m * S);
        uint64_t after = 0;
        MsgDesc d = ag_batch_msg(p, m);
        if (S > 1) d = ag_batch_segment(d, seg, S, &after);
        uint32_t ivw[3];
        ag_batch_iv(p, m, ivw, &d.j0ctr);
        const uint32_t iv0 = ivw[0], iv1 = ivw[1], iv2 = ivw[2];
        const AesCtrConst cc = aes_ctr_precompute(p.rk, iv0, iv1, iv2, te);
        // lanes step their counter by nt (512: a multiple of 256): the stream kernel's cache fits
        AesCtrCache cache;
        cache.key = 0x00FFFF00u;  // invalid: a valid key has those bits clear
        uint32_t e[4] = {0, 0, 0, 0};
        gf128 y = ag_batch_lane<NR, DEC>(p.rk, cc, cache, d, tid, nt, te, gh_g, e);
        if (tid == nt - 1) {   // the lane that met the length block also produced E_K(J0)
            uint32_t* ej = reinterpret_cast<uint32_t*>(ag_smem + SM_MISC + 1024);
            ej[0] = e[0]; ej[1] = e[1]; ej[2] = e[2]; ej[3] = e[3];
        }
        if (__any_sync(0xffffffffu, (y.w[0] | y.w[1] | y.w[2] | y.w[3]) != 0)) y = gf_mul(y, wgt);
        y = warp_xor(y);
        if (lane == 0) red[tid >> 5] = y;
        __syncthreads();
        if (tid < 32) {
            gf128 r = (tid < (nt >> 5)) ? red[tid] : gf_zero();
            r = warp_xor(r);
            const uint32_t* ej = reinterpret_cast<const uint32_t*>(ag_smem + SM_MISC + 1024);
            if (S > 1) {
                if (after) {  // uniform
                    const uint32_t slot = (seg >> 2) & 31u;
                    gf128 ha;
                    if (pow_keys[slot] == after) {
                        ha = pow_vals[slot];
                    } else {
                        ha = warp_gf_pow(p.key, after);
                        __syncwarp();
                        if (tid == 0) {
                            pow_vals[slot] = ha;
                            pow_keys[slot] = after;
                        }
                        __syncwarp();
                    }
                    r = gf_mul(r, ha);
                }
                if (tid == 0) {
                    uint32_t* dst = p.seg_parts + 4 * u;
                    dst[0] = r.w[0]; dst[1] = r.w[1]; dst[2] = r.w[2]; dst[3] = r.w[3];
                    if (d.last) {
                        uint32_t* de = p.seg_parts + 4 * (n_units + m);
                        de[0] = ej[0]; de[1] = ej[1]; de[2] = ej[2]; de[3] = ej[3];
                    }
                }
            } else if (tid == 0) {
                uint32_t tg[4] = {ag_bswap32(r.w[0]) ^ ej[0], ag_bswap32(r.w[1]) ^ ej[1], ag_bswap32(r.w[2]) ^ ej[2],
                                  ag_bswap32(r.w[3]) ^ ej[3]};
                uint8_t* tp = p.tag + 16 * m;
                if (DEC) {
                    uint32_t x[4];
                    ag_load_block(tp, 16, x);
                    const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
                    p.ok[m] = diff ? 0 : 1;
                } else {
                    ag_store_block(tp, 16, tg);
                }
            }
        }
        __syncthreads();
    }
}

// ===========================================================================
// Batched MID-SIZE and LONG messages under the shared key: one WARP per unit, where a unit is a
// message or a counter-range part of it (ag_batch_range, the single-GPU form of the shards of
// parallel.py).  Two ways of cutting:
//   * uniform batches: a static BALANCED partition.  The AAD blocks of all messages are laid end
//     to end on one axis, their payload blocks (+ 8 positions per message for the length block and
//     E_K(J0)) on a second one, and warp w owns the w-th equal share of EACH: every warp gets the
//     same number of GHASH-only rows and the same number of AES rows to within one, whatever the
//     number and size of the messages (a single weighted axis would hand some warps only AAD and
//     others only payload, and those advance at different speeds next to each other: measured).
//     At most 2 x (n_warps + n_msgs) units, so the per-unit overhead is paid a few times per warp;
//   * offset (ragged) batches: `split` equal-work segments per message, handed out by ticket.
// The per-unit epilogue is deferred: a warp DUMPS its 32 raw lane accumulators (512 B, coalesced).
// The lane weights H^(32-t) -- one ~1100-instruction generic product per lane and unit when done in
// place, which is what made fine cuts unaffordable for k_batch / k_batch_cta -- are applied in
// combine rounds with one LANE per unit (a 32-step Horner with the H table, up to 16 units side by
// side), which also scale by H^after and XOR the unit into its message's accumulator; the lane
// that completes a message's last unit writes (or checks) the tag.  One launch.  Linearity of GHASH
// in its input, as in src/gcm_ghash.vhd:317-344.  Tables: T_a = H^32, T_b = H.
// ===========================================================================
template <int NR, bool DEC>
__global__ void __launch_bounds__(AG_STREAM_NT_MAX, 1) k_batch_warp(const __grid_constant__ BatchParams p)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_te0(p.te0);
    fill_gh_tables(p.key->tab[5], p.key->tab[0]);   // T_a = H^32 (rows), T_b = H (unit combine)
    __syncthreads();
    expand_aes_tables();
    __syncthreads();
    TeSmem te{ag_smem, lane * 4};
    GhSmem gh{ag_smem + SM_GH, (lane & 7) * 16};
    GhSmem gh_1{ag_smem + SM_GH + 128, (lane & 7) * 16};
    // units this warp has dumped but not yet combined (ids), 16 per warp
    uint32_t* pend = reinterpret_cast<uint32_t*>(ag_smem + SM_MISC + 1024) + warp * 16;
    uint32_t n_pend = 0;
    const uint32_t S = p.split;
    const uint64_t ax_a = p.aad ? (p.aad_len + 15) >> 4 : 0;               // positions per message on the AAD axis
    const uint64_t ax_p = ((p.len + 15) >> 4) + AG_FINISH_WEIGHT;            // ... and on the payload axis

    // One more unit of message m is done (combined, or a cut that owned no block); whoever completes
    // the count turns the accumulator into the tag.
    auto arrive = [&](uint64_t m) {
        uint32_t units = S;   // how many units the message was cut into
        if (p.quota_pt) {
            units = (uint32_t)(((m + 1) * ax_p - 1) / p.quota_pt - (m * ax_p) / p.quota_pt + 1);
            if (ax_a) units += (uint32_t)(((m + 1) * ax_a - 1) / p.quota_aad - (m * ax_a) / p.quota_aad + 1);
        }
        __threadfence();
        const uint32_t done = atomicAdd(p.msg_cnt + m, 1u) + 1;
        if (done != units) return;
        __threadfence();
        uint32_t* dst = p.msg_acc + 4 * m;
        const uint32_t a0 = atomicOr(dst + 0, 0u), a1 = atomicOr(dst + 1, 0u), a2 = atomicOr(dst + 2, 0u), a3 = atomicOr(dst + 3, 0u);
        const uint4 e = __ldcg(reinterpret_cast<const uint4*>(p.msg_ej0 + 4 * m));
        const uint32_t tg[4] = {ag_bswap32(a0) ^ e.x, ag_bswap32(a1) ^ e.y, ag_bswap32(a2) ^ e.z, ag_bswap32(a3) ^ e.w};
        uint8_t* tp = p.tag + 16 * m;
        if (DEC) {
            uint32_t x[4];
            ag_load_block(tp, 16, x);
            const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
            p.ok[m] = diff ? 0 : 1;
        } else {
            ag_store_block(tp, 16, tg);
        }
    };

    // Combine round: lane i takes pending unit i.  R = sum_t Y_t H^(32-t) by a serial Horner over the
    // dumped accumulators with the H table, times H^(blocks after the unit) (product of the H^(2^k)
    // of the set bits), XOR into the message's accumulator; the lane that completes a message's last
    // unit turns the accumulator into the tag.
    auto combine = [&]() {
        __syncwarp();
        if (lane < n_pend) {
            const uint64_t id = pend[lane];
            const uint64_t m = __ldcg(p.unit_desc + 2 * id) - 1, after = __ldcg(p.unit_desc + 2 * id + 1);
            gf128 r = gf_zero();
            const uint4* acc = p.seg_acc + id * 32;
#pragma unroll 1
            for (int t = 0; t < 32; ++t) {
                const uint4 q = __ldcg(acc + t);
                r.w[0] ^= q.x; r.w[1] ^= q.y; r.w[2] ^= q.z; r.w[3] ^= q.w;
                r = gf_mul_table(r, gh_1);
            }
            if (after) {
                // H^after left to right: a GF(2)-linear squaring (~130 integer ops) per bit and, for a set
                // bit, one product with H through the shared table -- no generic products, no loads
                gf128 f = gf_one();
#pragma unroll 1
                for (int k = 63 - __clzll((long long)after); k >= 0; --k) {
                    f = gf_sqr(f);
                    if ((after >> k) & 1) f = gf_mul_table(f, gh_1);
                }
                r = gf_mul(r, f);
            }
            uint32_t* dst = p.msg_acc + 4 * m;
            atomicXor(dst + 0, r.w[0]); atomicXor(dst + 1, r.w[1]); atomicXor(dst + 2, r.w[2]); atomicXor(dst + 3, r.w[3]);
            arrive(m);
        }
        __syncwarp();
        n_pend = 0;
    };

    auto run_unit = [&](uint64_t id, uint64_t m, const MsgDesc& d, uint64_t after) {
        uint32_t ivw[3];
        MsgDesc du = d;
        ag_batch_iv(p, m, ivw, &du.j0ctr);
        const AesCtrConst cc = aes_ctr_precompute(p.rk, ivw[0], ivw[1], ivw[2], te);
        AesCtrSeqCache cache;
        cache.key = 0xFFFFFFFFu;
        uint32_t e[4] = {0, 0, 0, 0};
        const gf128 y = ag_batch_lane<NR, DEC>(p.rk, cc, cache, du, lane, 32u, te, gh, e);
        __stcg(p.seg_acc + id * 32 + lane, make_uint4(y.w[0], y.w[1], y.w[2], y.w[3]));
        if (lane == 0) {
            __stcg(p.unit_desc + 2 * id, m + 1);
            __stcg(p.unit_desc + 2 * id + 1, after);
            pend[n_pend] = (uint32_t)id;
        }
        if (du.last && lane == 31)   // the lane that met the length block also produced E_K(J0)
            __stcg(reinterpret_cast<uint4*>(p.msg_ej0 + 4 * m), make_uint4(e[0], e[1], e[2], e[3]));
        if (++n_pend == 16) combine();
    };

    if (p.quota_pt) {
        const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
        const uint64_t id_half = (uint64_t)gridDim.x * (blockDim.x >> 5) + p.n_msgs;
        // axis 0: AAD blocks (positions 0 .. a of a message), axis 1: payload blocks + finish (positions a .. )
        for (int axis = ax_a ? 0 : 1; axis < 2; ++axis) {
            const uint64_t per = axis ? ax_p : ax_a, quota = axis ? p.quota_pt : p.quota_aad, total = per * p.n_msgs;
            const uint64_t g0 = w * quota;
            uint64_t g1 = g0 + quota;
            if (g1 > total) g1 = total;
            for (uint64_t m = g0 / per; m * per < g1; ++m) {   // uniform per warp
                const uint64_t lo = m * per, r0 = (g0 > lo ? g0 : lo) - lo, r1 = (g1 < lo + per ? g1 : lo + per) - lo;
                uint64_t after = 0;
                const MsgDesc d = ag_batch_range(ag_batch_msg(p, m), axis ? ax_a + r0 : r0, axis ? ax_a + r1 : r1, &after, 1);
                if (!d.last && d.len == 0 && d.aad_len == 0) {   // a cut inside the finish positions: no block of it is mine
                    if (lane == 0) arrive(m);
                    continue;
                }
                run_unit((axis ? id_half : 0) + w + m, m, d, after);
            }
        }
    } else {
        const uint64_t n_units = p.n_msgs * S;
        for (;;) {
            uint32_t tk = 0;
            if (lane == 0) tk = atomicAdd(p.ticket, 1u);
            const uint64_t u = __shfl_sync(0xffffffffu, tk, 0);
            if (u >= n_units) break;
            const uint64_t m = u / S;
            uint64_t after = 0;
            MsgDesc d = ag_batch_msg(p, m);
            if (S > 1) d = ag_batch_segment(d, (uint32_t)(u - m * S), S, &after);
            run_unit(u, m, d, after);
        }
    }
    if (n_pend) combine();
}

// Tag finish of the split layout: one thread per message XORs its S scaled partials
// (linearity of GHASH in its input, the same algebra as gcm_ghash.vhd:330-332) and E_K(J0).
template <bool DEC>
__global__ void k_batch_split_finish(const __grid_constant__ BatchParams p)
{
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= p.n_msgs) return;
    const uint32_t S = p.split;
    uint32_t r[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < S; ++k) {
        const uint4 q = *reinterpret_cast<const uint4*>(p.seg_parts + 4 * (m * S + k));
        r[0] ^= q.x; r[1] ^= q.y; r[2] ^= q.z; r[3] ^= q.w;
    }
    const uint4 e = *reinterpret_cast<const uint4*>(p.seg_parts + 4 * (p.n_msgs * S + m));
    uint32_t tg[4] = {ag_bswap32(r[0]) ^ e.x, ag_bswap32(r[1]) ^ e.y, ag_bswap32(r[2]) ^ e.z, ag_bswap32(r[3]) ^ e.w};
    uint8_t* tp = p.tag + 16 * m;
    if (DEC) {
        uint32_t x[4];
        ag_load_block(tp, 16, x);
        const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
        p.ok[m] = diff ? 0 : 1;
    } else {
        ag_store_block(tp, 16, tg);
    }
}

// ===========================================================================
// Batched messages, one DISTINCT key per message (BASELINE config 4): one thread
// per message, key schedule on the fly (aes_kexp expand variant), private 4-bit
// GHASH table.  512 threads (128 registers each); shared memory: Te0|Te1 (64 KB,
// Te2/Te3 by a 16-bit rotate) + 512 x 256 B private tables.
// ===========================================================================
namespace {

constexpr uint32_t PK_NT = 512;
constexpr uint32_t PK_GH4 = 65536;          // + up to 2 KB alignment pad
constexpr uint32_t PK_SUBC = 12 * PK_NT * 4;   // SubWord outputs of the schedule: 12 words per thread, [j][tid]
constexpr uint32_t PK_SMEM = PK_GH4 + 2048 + PK_NT * 256 + PK_SUBC;

struct TeSmem2 {
    const uint8_t* base;
    uint32_t lane4;
    __device__ __forceinline__ uint32_t operator()(int tab, uint32_t w, int k) const
    {
        const uint32_t off = __byte_perm(w, lane4, 0x5504 | (k << 4));
        const uint32_t v = *reinterpret_cast<const uint32_t*>(base + off + (tab & 1) * 128);
        return (tab & 2) ? __byte_perm(v, 0, 0x1032) : v;
    }
};

// SubWord through byte 1 of the lane-private Te0 rows (Te0 = {2S, S, S, 3S})
struct SubWordSmem {
    const uint8_t* base;
    uint32_t lane4;
    __device__ __forceinline__ uint32_t operator()(uint32_t w) const
    {
        const uint32_t b0 = *(base + __byte_perm(w, lane4, 0x5504) + 1);
        const uint32_t b1 = *(base + __byte_perm(w, lane4, 0x5514) + 1);
        const uint32_t b2 = *(base + __byte_perm(w, lane4, 0x5524) + 1);
        const uint32_t b3 = *(base + __byte_perm(w, lane4, 0x5534) + 1);
        return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
};

// thread-private column: row n at base + n*128 (8 threads interleave 16 B slots in a
// 128 B row, so the 8 lanes of a quarter-warp never share a bank group)
struct Rows4Smem {
    uint32_t base;  // 32-bit shared address, bits 7..10 clear
    __device__ __forceinline__ void put(int n, uint4 r) const
    {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + n * 128), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
                     : "memory");
    }
    __device__ __forceinline__ uint4 get(uint32_t w, int k) const
    {
        const uint32_t n7 = (4 * k >= 7) ? (w >> (4 * k - 7)) : (w << (7 - 4 * k));
        const uint32_t addr = (n7 & 0x780u) | base;
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
        return r;
    }
};

// thread-private word column: word j at base + j * (4 * PK_NT) (a warp reads 128 consecutive bytes)
template <uint32_t NT>
struct SubCacheT {
    uint32_t base;  // 32-bit shared address of this thread's word 0
    __device__ __forceinline__ void put(int j, uint32_t v) const
    {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + j * (4 * NT)), "r"(v) : "memory");
    }
    __device__ __forceinline__ uint32_t get(int j) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + j * (4 * NT)) : "memory");
        return v;
    }
};
using SubCacheSmem = SubCacheT<PK_NT>;
using SubCacheTile = SubCacheT<448>;

__device__ __forceinline__ void load_words(const uint8_t* p, int n_words, uint32_t* w)
{
    if (((uintptr_t)p & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n_words) w[i] = q[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n_words)
                w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                       ((uint32_t)p[4 * i + 3] << 24);
    }
}

}  // namespace

template <int NK, bool DEC>
__global__ void __launch_bounds__(PK_NT, 1) k_batch_perkey(const __grid_constant__ BatchParams p)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    // Te0 | Te1 only
    for (uint32_t idx = tid; idx < 256 * 32; idx += blockDim.x) {
        const uint32_t x = idx >> 5, l = idx & 31;
        const uint32_t t = __ldg(p.te0 + x);
        uint32_t* a = reinterpret_cast<uint32_t*>(ag_smem + SM_AES_A + x * 256 + l * 4);
        a[0] = t;
        a[32] = ag_rotl32(t, 8);
    }
    __syncthreads();
    TeSmem2 te{ag_smem, lane * 4};
    SubWordSmem sb{ag_smem, lane * 4};
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(ag_smem) + PK_GH4;
    const uint32_t s_al = (s0 + 2047u) & ~2047u;
    Rows4Smem rows{s_al + (tid >> 3) * 2048u + (tid & 7) * 16u};
    SubCacheSmem subc{s_al + PK_NT * 256u + tid * 4u};

    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + tid; m < p.n_msgs; m += (uint64_t)gridDim.x * blockDim.x) {
        const MsgDesc d = ag_batch_msg(p, m);
        uint32_t key[8], iv[3];
        load_words(p.keys + m * (uint64_t)(4 * NK), NK, key);
        load_words(p.iv + 12 * m, 3, iv);
        uint32_t tg[4];
        ag_perkey_message<NK, DEC>(key, iv[0], iv[1], iv[2], d, te, sb, rows, subc, tg);
        uint8_t* tp = p.tag + 16 * m;
        if (DEC) {
            uint32_t x[4];
            ag_load_block(tp, 16, x);
            const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
            p.ok[m] = diff ? 0 : 1;
        } else {
            ag_store_block(tp, 16, tg);
        }
    }
}

// ---------------------------------------------------------------------------
// The same kernel for FIXED-SIZE records, staged by TMA like k_batch_tile: one lane per message,
// box {32 bytes x 32 messages} per warp and tile, two tiles per warp, groups of 32 messages by
// atomic ticket.  Removes the thread-per-message LDG.128 / STG.128 (32 lines per request: ~17 % of
// the binding L1/shared data pipe and 1.29x DRAM over-fetch, profiles/r1_ncu_perkey.md).
// 14 warps instead of 16: the tiles (28 KB) have to fit next to Te0|Te1 (64 KB), the private
// 4-bit GHASH tables (256 B per thread) and the SubWord columns (48 B per thread).
// ---------------------------------------------------------------------------
namespace {
constexpr uint32_t PKT_NT = 448;
constexpr uint32_t PKT_TILES = 65536;                               // 14 warps x 2 x 1 KB
constexpr uint32_t PKT_BARS = PKT_TILES + (PKT_NT / 32) * 2 * TILE_BYTES;   // 14 x 2 mbarriers
constexpr uint32_t PKT_GH4 = PKT_BARS + 256;                        // + up to 2 KB alignment pad
constexpr uint32_t PKT_SMEM = 232448;                               // all of it (227 KB); the layout is checked at run time
}  // namespace

template <int NK, bool DEC>
__global__ void __launch_bounds__(PKT_NT, 1) k_batch_perkey_tile(const __grid_constant__ TileParams P)
{
    const BatchParams& p = P.b;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t idx = tid; idx < 256 * 32; idx += blockDim.x) {   // Te0 | Te1 only (Te2/Te3 by a 16-bit rotate)
        const uint32_t x = idx >> 5, l = idx & 31;
        const uint32_t t = __ldg(p.te0 + x);
        uint32_t* a = reinterpret_cast<uint32_t*>(ag_smem + SM_AES_A + x * 256 + l * 4);
        a[0] = t;
        a[32] = ag_rotl32(t, 8);
    }
    const uint32_t bar0 = ag_smem_addr(ag_smem + PKT_BARS + warp * 16), bar1 = bar0 + 8;
    if (lane == 0) {
        ag_mbar_init(bar0, 1);
        ag_mbar_init(bar1, 1);
        ag_fence_barrier_init();
        ag_prefetch_tmap(&P.tm_in);
        ag_prefetch_tmap(&P.tm_out);
    }
    __syncthreads();
    TeSmem2 te{ag_smem, lane * 4};
    SubWordSmem sb{ag_smem, lane * 4};
    const uint32_t s0 = ag_smem_addr(ag_smem) + PKT_GH4;
    const uint32_t s_al = (s0 + 2047u) & ~2047u;
    Rows4Smem rows{s_al + (tid >> 3) * 2048u + (tid & 7) * 16u};
    SubCacheTile subc{s_al + PKT_NT * 256u + tid * 4u};
    if (s_al + PKT_NT * 256u + 12u * PKT_NT * 4u > ag_smem_addr(ag_smem) + PKT_SMEM) __trap();   // layout does not fit

    uint8_t* tiles = ag_smem + PKT_TILES + warp * (2 * TILE_BYTES);
    const uint32_t tile_sa = ag_smem_addr(tiles);
    const uint32_t sw = (lane >> 2) & 1;
    const uint32_t coff0 = lane * 32 + ((0 ^ sw) << 4), coff1 = lane * 32 + ((1 ^ sw) << 4);
    uint32_t par0 = 0, par1 = 0;
    const uint32_t n_blocks = (uint32_t)((p.len + 15) >> 4), tail = (uint32_t)(p.len & 15), n_full = (uint32_t)(p.len >> 4);
    const uint32_t n_tiles = (n_blocks + 1) >> 1;
    const uint32_t a_blocks = (uint32_t)((p.aad_len + 15) >> 4), atail = (uint32_t)(p.aad_len & 15);
    const uint32_t n_groups = (uint32_t)((p.n_msgs + 31) >> 5);
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(P.ticket, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_groups) break;
        const int32_t row0 = (int32_t)(g * 32);
        if (lane == 0 && n_tiles) {
            ag_mbar_expect_tx(bar0, TILE_BYTES);
            ag_tma_load_2d(tile_sa, &P.tm_in, 0, row0, bar0);
        }
        const uint64_t m_raw = (uint64_t)g * 32 + lane;
        const bool valid = m_raw < p.n_msgs;
        const uint64_t m = valid ? m_raw : p.n_msgs - 1;
        uint32_t key[8], iv[3];
        load_words(p.keys + m * (uint64_t)(4 * NK), NK, key);
        load_words(p.iv + 12 * m, 3, iv);
        uint32_t e[4];
        aes_encrypt_otf<NK>(key, 0, 0, 0, 0, te, sb, e);  // H = E_K(0^128)  (gcm_gctr.vhd:141-144)
        gf_build_table4(gf_from_le_words(e[0], e[1], e[2], e[3]), rows);
        PerKeyCtr<NK> st;
        perkey_ctr_init<NK>(key, iv[0], iv[1], iv[2], te, sb, subc, st);
        perkey_ctr_block<NK>(st, 1u, te, subc, e);  // E_K(J0)
        gf128 y = gf_zero();
        if (a_blocks) {
            const uint8_t* ap = p.aad + m * p.aad_stride;
            for (uint32_t i = 0; i < a_blocks; ++i) {
                uint32_t x[4];
                ag_load_block(ap + 16 * (uint64_t)i, (i == a_blocks - 1 && atail) ? atail : 16u, x);
                y = gf_xor(y, gf_from_le_words(x[0], x[1], x[2], x[3]));
                y = gf_mul_table4(y, rows);
            }
        }
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t b = t & 1;
            if (lane == 0 && t + 1 < n_tiles) {
                ag_bulk_wait_read0();
                ag_mbar_expect_tx(b ? bar0 : bar1, TILE_BYTES);
                ag_tma_load_2d(tile_sa + (b ^ 1) * TILE_BYTES, &P.tm_in, (int32_t)((t + 1) * 32), row0, b ? bar0 : bar1);
            }
            if (b) { ag_mbar_wait(bar1, par1); par1 ^= 1; } else { ag_mbar_wait(bar0, par0); par0 ^= 1; }
            uint8_t* tb = tiles + b * TILE_BYTES;
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const uint32_t j = 2 * t + k;
                if (j < n_blocks) {   // uniform
                    uint4* cp = reinterpret_cast<uint4*>(tb + (k ? coff1 : coff0));
                    const uint4 xv = *cp;
                    uint32_t x[4] = {xv.x, xv.y, xv.z, xv.w};
                    const bool ragged = (j == n_full);
                    if (ragged) ag_mask_block(x, tail);
                    uint32_t ks[4];
                    perkey_ctr_block<NK>(st, 2u + j, te, subc, ks);
                    uint32_t o[4] = {x[0] ^ ks[0], x[1] ^ ks[1], x[2] ^ ks[2], x[3] ^ ks[3]};
                    if (!ragged) {
                        *cp = make_uint4(o[0], o[1], o[2], o[3]);
                    } else {
                        if (valid) ag_store_block(p.out + m * p.stride + 16 * (uint64_t)j, tail, o);
                        ag_mask_block(o, tail);
                    }
                    if (DEC) y = gf_xor(y, gf_from_le_words(x[0], x[1], x[2], x[3]));
                    else y = gf_xor(y, gf_from_le_words(o[0], o[1], o[2], o[3]));
                    y = gf_mul_table4(y, rows);
                }
            }
            if (2 * t < n_full) {
                ag_fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ag_tma_store_2d(&P.tm_out, (int32_t)(t * 32), row0, tile_sa + b * TILE_BYTES);
                    ag_bulk_commit();
                }
            } else {
                __syncwarp();
            }
        }
        const uint64_t ab = p.aad_len * 8, cb = p.len * 8;
        y.w[0] ^= (uint32_t)(ab >> 32); y.w[1] ^= (uint32_t)ab; y.w[2] ^= (uint32_t)(cb >> 32); y.w[3] ^= (uint32_t)cb;
        y = gf_mul_table4(y, rows);
        const uint32_t tg[4] = {ag_bswap32(y.w[0]) ^ e[0], ag_bswap32(y.w[1]) ^ e[1], ag_bswap32(y.w[2]) ^ e[2],
                                ag_bswap32(y.w[3]) ^ e[3]};
        if (valid) {
            uint8_t* tp = p.tag + 16 * m;
            if (DEC) {
                uint32_t x[4];
                ag_load_block(tp, 16, x);
                const uint32_t diff = (x[0] ^ tg[0]) | (x[1] ^ tg[1]) | (x[2] ^ tg[2]) | (x[3] ^ tg[3]);
                p.ok[m] = diff ? 0 : 1;
            } else {
                ag_store_block(tp, 16, tg);
            }
        }
        if (lane == 0) ag_bulk_wait_read0();
        __syncwarp();
    }
    if (lane == 0) ag_bulk_wait0();
}

template <int NK>
static cudaError_t launch_perkey_tile_t(const TileParams& p, int decrypt, int ncta, cudaStream_t st)
{
    cudaError_t e;
    if (decrypt) {
        e = cudaFuncSetAttribute(k_batch_perkey_tile<NK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PKT_SMEM);
        if (e != cudaSuccess) return e;
        k_batch_perkey_tile<NK, true><<<ncta, PKT_NT, PKT_SMEM, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_batch_perkey_tile<NK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PKT_SMEM);
        if (e != cudaSuccess) return e;
        k_batch_perkey_tile<NK, false><<<ncta, PKT_NT, PKT_SMEM, st>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t ag_launch_batch_perkey_tile(const TileParams& p, int nr, int decrypt, int max_cta, cudaStream_t st)
{
    const uint64_t groups = (p.b.n_msgs + 31) / 32, per_cta = PKT_NT / 32;
    const uint64_t need = (groups + per_cta - 1) / per_cta;
    const int ncta = (int)(need < (uint64_t)max_cta ? need : (uint64_t)max_cta);
    switch (nr) {
        case 10: return launch_perkey_tile_t<4>(p, decrypt, ncta, st);
        case 12: return launch_perkey_tile_t<6>(p, decrypt, ncta, st);
        case 14: return launch_perkey_tile_t<8>(p, decrypt, ncta, st);
    }
    return cudaErrorInvalidValue;
}

template <int NK>
static cudaError_t launch_perkey_t(const BatchParams& p, int decrypt, int ncta, cudaStream_t st)
{
    cudaError_t e;
    if (decrypt) {
        e = cudaFuncSetAttribute(k_batch_perkey<NK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM);
        if (e != cudaSuccess) return e;
        k_batch_perkey<NK, true><<<ncta, PK_NT, PK_SMEM, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_batch_perkey<NK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM);
        if (e != cudaSuccess) return e;
        k_batch_perkey<NK, false><<<ncta, PK_NT, PK_SMEM, st>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t ag_launch_batch_perkey(const BatchParams& p, int nr, int decrypt, int max_cta, cudaStream_t st)
{
    const uint64_t need = (p.n_msgs + PK_NT - 1) / PK_NT;
    const int ncta = (int)(need < (uint64_t)max_cta ? need : (uint64_t)max_cta);
    switch (nr) {
        case 10: return launch_perkey_t<4>(p, decrypt, ncta, st);
        case 12: return launch_perkey_t<6>(p, decrypt, ncta, st);
        case 14: return launch_perkey_t<8>(p, decrypt, ncta, st);
    }
    return cudaErrorInvalidValue;
}

// ===========================================================================
// launchers (called from capi.cu)
// ===========================================================================
static const size_t kSmemBytes = SM_MISC + 2048;
size_t ag_smem_bytes() { return kSmemBytes; }

template <int NR, int MODE>
static cudaError_t launch_stream_t(const StreamParams& p, int ncta, int nt, cudaStream_t st)
{
    // 128-bit path when both buffers are 16-byte aligned (GHASH-only has no output buffer)
    const bool aligned = (((uintptr_t)p.in | (uintptr_t)p.out) & 15) == 0;
    cudaError_t e;
    if (aligned) {
        e = cudaFuncSetAttribute(k_stream<NR, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        k_stream<NR, MODE, true><<<ncta, nt, kSmemBytes, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(k_stream<NR, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        k_stream<NR, MODE, false><<<ncta, nt, kSmemBytes, st>>>(p);
    }
    return cudaGetLastError();
}

template <int NR>
static cudaError_t launch_stream_nr(const StreamParams& p, int mode, int ncta, int nt, cudaStream_t st)
{
    switch (mode) {
        case AG_MODE_ENC: return launch_stream_t<NR, AG_MODE_ENC>(p, ncta, nt, st);
        case AG_MODE_DEC: return launch_stream_t<NR, AG_MODE_DEC>(p, ncta, nt, st);
        case AG_MODE_GHASH_ONLY: return launch_stream_t<NR, AG_MODE_GHASH_ONLY>(p, ncta, nt, st);
        case AG_MODE_CTR_ONLY: return launch_stream_t<NR, AG_MODE_CTR_ONLY>(p, ncta, nt, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ag_launch_stream(const StreamParams& p, int nr, int mode, int ncta, int nt, cudaStream_t st)
{
    switch (nr) {
        case 10: return launch_stream_nr<10>(p, mode, ncta, nt, st);
        case 12: return launch_stream_nr<12>(p, mode, ncta, nt, st);
        case 14: return launch_stream_nr<14>(p, mode, ncta, nt, st);
    }
    return cudaErrorInvalidValue;
}

template <int NR, bool DEC, int G>
static cudaError_t launch_batch_t(const BatchParams& p, int ncta, int nt, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_batch<NR, DEC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    k_batch<NR, DEC, G><<<ncta, nt, kSmemBytes, st>>>(p);
    return cudaGetLastError();
}

template <int NR, bool DEC>
static cudaError_t launch_batch_g(const BatchParams& p, int g, int ncta, int nt, cudaStream_t st)
{
    switch (g) {
        case 1: return launch_batch_t<NR, DEC, 1>(p, ncta, nt, st);
        case 2: return launch_batch_t<NR, DEC, 2>(p, ncta, nt, st);
        case 4: return launch_batch_t<NR, DEC, 4>(p, ncta, nt, st);
        case 8: return launch_batch_t<NR, DEC, 8>(p, ncta, nt, st);
        case 16: return launch_batch_t<NR, DEC, 16>(p, ncta, nt, st);
        case 32: return launch_batch_t<NR, DEC, 32>(p, ncta, nt, st);
    }
    return cudaErrorInvalidValue;
}

template <int NR, bool DEC>
static cudaError_t launch_batch_cta_t(const BatchParams& p, int ncta, int nt, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_batch_cta<NR, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    k_batch_cta<NR, DEC><<<ncta, nt, kSmemBytes, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess || p.split <= 1) return e;
    k_batch_split_finish<DEC><<<(unsigned)((p.n_msgs + 127) / 128), 128, 0, st>>>(p);
    return cudaGetLastError();
}

template <int NR, bool DEC>
static cudaError_t launch_batch_tile_t(const TileParams& p, int ncta, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_batch_tile<NR, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
    if (e != cudaSuccess) return e;
    k_batch_tile<NR, DEC><<<ncta, AG_STREAM_NT_MAX, kTileSmemBytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ag_launch_batch_tile(const TileParams& p, int nr, int decrypt, int ncta, cudaStream_t st)
{
    switch (nr) {
        case 10: return decrypt ? launch_batch_tile_t<10, true>(p, ncta, st) : launch_batch_tile_t<10, false>(p, ncta, st);
        case 12: return decrypt ? launch_batch_tile_t<12, true>(p, ncta, st) : launch_batch_tile_t<12, false>(p, ncta, st);
        case 14: return decrypt ? launch_batch_tile_t<14, true>(p, ncta, st) : launch_batch_tile_t<14, false>(p, ncta, st);
    }
    return cudaErrorInvalidValue;
}

template <int NR, bool DEC>
static cudaError_t launch_batch_warp_t(const BatchParams& p, int ncta, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_batch_warp<NR, DEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    k_batch_warp<NR, DEC><<<ncta, AG_STREAM_NT_MAX, kSmemBytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ag_launch_batch_warp(const BatchParams& p, int nr, int decrypt, int ncta, cudaStream_t st)
{
    switch (nr) {
        case 10: return decrypt ? launch_batch_warp_t<10, true>(p, ncta, st) : launch_batch_warp_t<10, false>(p, ncta, st);
        case 12: return decrypt ? launch_batch_warp_t<12, true>(p, ncta, st) : launch_batch_warp_t<12, false>(p, ncta, st);
        case 14: return decrypt ? launch_batch_warp_t<14, true>(p, ncta, st) : launch_batch_warp_t<14, false>(p, ncta, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ag_launch_batch_cta(const BatchParams& p, int nr, int decrypt, int ncta, int nt, cudaStream_t st)
{
    switch (nr) {
        case 10: return decrypt ? launch_batch_cta_t<10, true>(p, ncta, nt, st) : launch_batch_cta_t<10, false>(p, ncta, nt, st);
        case 12: return decrypt ? launch_batch_cta_t<12, true>(p, ncta, nt, st) : launch_batch_cta_t<12, false>(p, ncta, nt, st);
        case 14: return decrypt ? launch_batch_cta_t<14, true>(p, ncta, nt, st) : launch_batch_cta_t<14, false>(p, ncta, nt, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ag_launch_batch(const BatchParams& p, int nr, int decrypt, int g, int ncta, int nt, cudaStream_t st)
{
    switch (nr) {
        case 10: return decrypt ? launch_batch_g<10, true>(p, g, ncta, nt, st) : launch_batch_g<10, false>(p, g, ncta, nt, st);
        case 12: return decrypt ? launch_batch_g<12, true>(p, g, ncta, nt, st) : launch_batch_g<12, false>(p, g, ncta, nt, st);
        case 14: return decrypt ? launch_batch_g<14, true>(p, g, ncta, nt, st) : launch_batch_g<14, false>(p, g, ncta, nt, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t ag_launch_key_expand(const uint8_t* keys, uint64_t n_keys, int key_bytes, const uint32_t* te0,
                                 uint8_t* round_keys, cudaStream_t st)
{
    if (n_keys == 0) return cudaSuccess;
    const int nt = 128;
    const uint64_t nb = (n_keys + nt - 1) / nt;
    k_key_expand<<<(unsigned)nb, nt, 0, st>>>(keys, n_keys, key_bytes, te0, round_keys);
    return cudaGetLastError();
}

cudaError_t ag_launch_key_setup(KeyDev* kd, const KeyIn& in, const uint32_t* te0, int nt_stream, int ncta,
                                cudaStream_t st)
{
    k_key_setup<<<1, 256, 0, st>>>(kd, in, te0, (uint32_t)nt_stream, (uint32_t)ncta);
    return cudaGetLastError();
}

cudaError_t ag_launch_pow(const KeyDev* kd, uint64_t e, uint32_t* out, cudaStream_t st)
{
    k_pow<<<1, 32, 0, st>>>(kd, e, out);
    return cudaGetLastError();
}

cudaError_t ag_launch_batch_j0(const KeyDev* kd, const uint8_t* iv, const uint64_t* iv_off, uint64_t iv_len, uint64_t n,
                               uint8_t* j0, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    k_batch_j0<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kd, iv, iv_off, iv_len, n, j0);
    return cudaGetLastError();
}

cudaError_t ag_launch_xor_parts(const uint8_t* parts, uint32_t n, uint8_t* out, cudaStream_t st)
{
    k_xor_parts<<<1, 32, 0, st>>>(parts, n, out);
    return cudaGetLastError();
}

// CUDA loads a kernel lazily at its first launch, and that load can wait for running kernels to
// drain.  A first launch of k_peer_post / k_peer_finish while a finish kernel of the same process
// spins for a peer flag would then stall until the timeout: load them at agcm_peer_setup instead.
cudaError_t ag_preload_peer_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, k_peer_post);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_peer_finish);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_pow);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_xor_parts);
    return e;
}

cudaError_t ag_launch_peer_finish(const PeerFinishParams& p, cudaStream_t st)
{
    k_peer_finish<<<1, 32, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ag_launch_peer_post(uint8_t* const* peer_bufs, uint32_t rank, uint32_t world, uint32_t epoch,
                                const uint8_t* partial16, cudaStream_t st)
{
    k_peer_post<<<1, 32, 0, st>>>(peer_bufs, rank, world, epoch, partial16);
    return cudaGetLastError();
}

cudaError_t ag_launch_finish(const FinishParams& p, cudaStream_t st)
{
    k_stream_finish<<<1, 32, 0, st>>>(p);
    return cudaGetLastError();
}
