// GF(2^128) arithmetic in the GCM bit order, shared by device kernels and the
// host-side kernel emulator used by the CPU tests (tests/host_emul.cu).
//
// Behavioural spec: src/ghash_gfmul.vhd:42-63 (SP 800-38D Algorithm 1): bit 127
// of the VHDL vector is the MSB of byte 0 and is the coefficient of x^0; the
// multiply-by-x step is "V >> 1, xor 0xE1||0^120 when the dropped bit was 1".
//
// Representation: four BIG-ENDIAN 32-bit words; w[0] bits 31..24 hold byte 0, so
// the whole element reads as one 128-bit string w[0]:w[1]:w[2]:w[3] and
// "multiply by x" is a funnel shift right across the words.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AG_HD __host__ __device__ __forceinline__
#define AG_D __device__ __forceinline__
#else
#define AG_HD inline
#endif

// Only for builds with 1024-thread CTAs (AG_NT_MAX > 512, 64 registers per thread): keeps at
// most 8 of the 16 row lookups of one table product in flight (32 registers) with a
// compiler-level memory fence after the second group of four.  The default 512-thread build has
// 128 registers and lets the compiler hoist all sixteen 128-bit loads.
#if defined(__CUDA_ARCH__) && (!defined(AG_NT_MAX) || AG_NT_MAX > 512)
#define AG_LOOKUP_FENCE(r) do { if ((r) == 1) asm volatile("" ::: "memory"); } while (0)
#else
#define AG_LOOKUP_FENCE(r) do { } while (0)
#endif

struct gf128 {
    uint32_t w[4];
};

AG_HD uint32_t ag_bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

// low 32 bits of (hi:lo) >> sh, 0 <= sh <= 31
AG_HD uint32_t ag_funnel_r(uint32_t lo, uint32_t hi, int sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

AG_HD gf128 gf_zero()
{
    gf128 r;
    r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0;
    return r;
}

// the multiplicative identity: x^0 = MSB of byte 0
AG_HD gf128 gf_one()
{
    gf128 r = gf_zero();
    r.w[0] = 0x80000000u;
    return r;
}

AG_HD gf128 gf_xor(const gf128& a, const gf128& b)
{
    gf128 r;
    r.w[0] = a.w[0] ^ b.w[0];
    r.w[1] = a.w[1] ^ b.w[1];
    r.w[2] = a.w[2] ^ b.w[2];
    r.w[3] = a.w[3] ^ b.w[3];
    return r;
}

// V * x  (ghash_gfmul.vhd:50-57)
AG_HD gf128 gf_mulx(const gf128& v)
{
    gf128 r;
    uint32_t lsb = v.w[3] & 1u;
    r.w[3] = ag_funnel_r(v.w[3], v.w[2], 1);
    r.w[2] = ag_funnel_r(v.w[2], v.w[1], 1);
    r.w[1] = ag_funnel_r(v.w[1], v.w[0], 1);
    r.w[0] = (v.w[0] >> 1) ^ (0xE1000000u & (0u - lsb));
    return r;
}

// Squaring is GF(2)-linear: (sum a_i x^i)^2 = sum a_i x^(2i).  Spread the 128-bit string with
// zeros (bit at string offset k -> offset 2k), then fold degrees 128..254 back (gf_fold256).  ~130 integer
// ops instead of the ~1500 of the bit-serial product; used for the H^(2^k) chain of k_key_setup.
AG_HD uint32_t ag_spread16(uint32_t x)   // bit j of the low 16 bits -> bit 2j
{
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// Reduce an unreduced 256-bit product (words 0..7, word 0 bit 31 = x^0, degrees up to 255) with
// x^128 = 1 + x + x^2 + x^7: one fold of words 4..7, then a fold of the (at most 7) bits the
// first one pushed past degree 255.
AG_HD gf128 gf_fold256(const uint32_t z[8])
{
    uint32_t e[5];
    e[0] = z[4] ^ (z[4] >> 1) ^ (z[4] >> 2) ^ (z[4] >> 7);
#pragma unroll
    for (int m = 1; m < 4; ++m)
        e[m] = z[4 + m] ^ ag_funnel_r(z[4 + m], z[3 + m], 1) ^ ag_funnel_r(z[4 + m], z[3 + m], 2) ^
               ag_funnel_r(z[4 + m], z[3 + m], 7);
    const uint32_t u = (z[7] << 31) ^ (z[7] << 30) ^ (z[7] << 25);   // degrees 128..134 of the first fold
    gf128 o;
    o.w[0] = z[0] ^ e[0] ^ u ^ (u >> 1) ^ (u >> 2) ^ (u >> 7);
    o.w[1] = z[1] ^ e[1];
    o.w[2] = z[2] ^ e[2];
    o.w[3] = z[3] ^ e[3];
    return o;
}

AG_HD gf128 gf_sqr(const gf128& a)
{
    uint32_t z[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // string offset k of word q (k = 0 is bit 31) -> offset 2k of the 64-bit pair (z[2q], z[2q+1])
        z[2 * q] = ag_spread16(a.w[q] >> 16) << 1;
        z[2 * q + 1] = ag_spread16(a.w[q]) << 1;
    }
    return gf_fold256(z);
}

// Generic product X*V (the function of ghash_gfmul.vhd:42-63), off the per-block path (key setup,
// per-thread / per-CTA weights, tag finish).  Instead of the serial "V <- V*x with reduction" chain
// of Algorithm 1 (128 dependent steps), the product is accumulated UNREDUCED: the bit of X at
// degree 32q + j adds V * x^j (V moved j bits along the string, 5 words) at word offset q of an
// 8-word accumulator, and one gf_fold256 reduces at the end.  32 steps of about 33 independent
// integer operations: roughly 2.5x faster than the bit-serial form on one thread.
AG_HD gf128 gf_mul(const gf128& x, const gf128& v)
{
    uint32_t z[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) z[m] = 0;
    uint32_t s0 = v.w[0], s1 = v.w[1], s2 = v.w[2], s3 = v.w[3], s4 = 0;   // V * x^j
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < 32; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t m = 0u - ((x.w[q] >> (31 - j)) & 1u);
            z[q] ^= s0 & m;
            z[q + 1] ^= s1 & m;
            z[q + 2] ^= s2 & m;
            z[q + 3] ^= s3 & m;
            z[q + 4] ^= s4 & m;
        }
        s4 = ag_funnel_r(s4, s3, 1);
        s3 = ag_funnel_r(s3, s2, 1);
        s2 = ag_funnel_r(s2, s1, 1);
        s1 = ag_funnel_r(s1, s0, 1);
        s0 >>= 1;
    }
    return gf_fold256(z);
}

// 16 bytes as loaded little-endian (uint4 of LE words, the AES state layout)
// <-> field element
AG_HD gf128 gf_from_le_words(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    gf128 r;
    r.w[0] = ag_bswap32(a);
    r.w[1] = ag_bswap32(b);
    r.w[2] = ag_bswap32(c);
    r.w[3] = ag_bswap32(d);
    return r;
}

// ---------------------------------------------------------------------------
// Multiply by a FIXED element C through an 8-bit Shoup table T[b] = b*C, where
// the byte b is read as a degree<8 polynomial with its MSB = x^0.
//
//   X = sum_j B_j x^(8j)  (B_j = byte j)   =>   X*C = sum_j T[B_j] x^(8j)
//
// The 16 table rows are XORed, unreduced, into a 31-byte string at byte offset
// j, grouped by r = j mod 4 so that each group only needs whole-word offsets;
// the three groups r=1..3 are then byte-shifted once.  The 120 overflow bits are
// folded back once per product with x^128 = 1 + x + x^2 + x^7.
//
// LOOKUP: functor (uint32_t be_word, int le_byte) -> uint4 {w0,w1,w2,w3} of
// T[(be_word >> 8*le_byte) & 0xff].  On the device it is one PRMT + one LDS.128.
template <class LOOKUP>
AG_HD gf128 gf_mul_table(const gf128& x, LOOKUP&& lookup)
{
    uint32_t z[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) z[m] = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t a[7];
#pragma unroll
        for (int m = 0; m < 7; ++m) a[m] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 t = lookup(x.w[q], 3 - r);
            a[q] ^= t.x;
            a[q + 1] ^= t.y;
            a[q + 2] ^= t.z;
            a[q + 3] ^= t.w;
        }
        if (r == 0) {
#pragma unroll
            for (int m = 0; m < 7; ++m) z[m] ^= a[m];
        } else {
            z[0] ^= a[0] >> (8 * r);
#pragma unroll
            for (int m = 1; m < 7; ++m) z[m] ^= ag_funnel_r(a[m], a[m - 1], 8 * r);
            z[7] ^= a[6] << (32 - 8 * r);
        }
        AG_LOOKUP_FENCE(r);
    }
    // fold bytes 16..30 (degrees 128..247); the last byte of z[7] is always zero,
    // so the shifts by 1, 2 and 7 bits lose nothing.
    gf128 o;
    o.w[0] = z[0] ^ z[4] ^ (z[4] >> 1) ^ (z[4] >> 2) ^ (z[4] >> 7);
#pragma unroll
    for (int m = 1; m < 4; ++m)
        o.w[m] = z[m] ^ z[4 + m] ^ ag_funnel_r(z[4 + m], z[3 + m], 1) ^ ag_funnel_r(z[4 + m], z[3 + m], 2) ^
                 ag_funnel_r(z[4 + m], z[3 + m], 7);
    return o;
}

// Row b of the compact Shoup table for constant C.  basis[k] = C * x^k, k=0..7.
AG_HD uint4 gf_table_row(const gf128 basis[8], uint32_t b)
{
    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t m = 0u - ((b >> (7 - k)) & 1u);
        r.x ^= basis[k].w[0] & m;
        r.y ^= basis[k].w[1] & m;
        r.z ^= basis[k].w[2] & m;
        r.w ^= basis[k].w[3] & m;
    }
    return r;
}
