// AES primitives for the sm_100a engine (host+device so that the CPU tests can
// run the exact per-thread code through tests/host_emul.cu).
//
// Behavioural spec (reference file:line):
//   byte order / state layout      src/aes_func.vhd:85-93  (state(i)(j) = byte 4i+j)
//   S-box                          src/aes_func.vhd:228-301
//   ShiftRows / MixColumns         src/aes_func.vhd:146-169,187-210
//   round phasing, key indexing    config/config_aes_round.py:120-126,142
//   final AddRoundKey              src/aes_last_round.vhd:76
//   key schedule                   tb/key_exp.py:79-114, config/config_aes_kexp.py:128-159
//
// Formulation: a 16-byte block is four LITTLE-ENDIAN words (word c = column c,
// row r in byte r), exactly what a 128-bit global load delivers.  One round is
// 16 table lookups Te_r[byte r of column (c+r)%4] XORed per output column
// (SubBytes+ShiftRows+MixColumns folded into the tables), then the stage key.
// Te0[x] = {2S, S, S, 3S} (rows 0..3), Te_r = Te0 rotated left by 8r bits.
#pragma once
#include <stdint.h>
#include "gf128.cuh"

// ---- S-box and Te0 by definition (host; runs once per context) -------------
static inline uint8_t ag_gf256_xtime(uint8_t a) { return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0)); }

static inline void ag_build_sbox_te0(uint8_t sbox[256], uint32_t te0[256])
{
    // log/antilog over generator 3, then inverse + affine map (FIPS-197 5.1.1)
    uint8_t exp_t[256], log_t[256];
    uint8_t v = 1;
    for (int i = 0; i < 255; i++) {
        exp_t[i] = v;
        log_t[v] = (uint8_t)i;
        v = (uint8_t)(v ^ ag_gf256_xtime(v));  // v *= 3
    }
    for (int x = 0; x < 256; x++) {
        uint8_t inv = x ? exp_t[(255 - log_t[x]) % 255] : 0;
        uint8_t s = inv;
        uint8_t rot = inv;
        for (int i = 0; i < 4; i++) {
            rot = (uint8_t)((rot << 1) | (rot >> 7));
            s ^= rot;
        }
        s ^= 0x63;
        sbox[x] = s;
        uint8_t s2 = ag_gf256_xtime(s), s3 = (uint8_t)(s2 ^ s);
        te0[x] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
    }
}

AG_HD uint32_t ag_rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

// ---- key schedule -----------------------------------------------------------
// SBOX: functor uint32_t(uint32_t byte).  key: key_bytes raw bytes.  rk receives
// 4*(Nr+1) little-endian words (stage r = words 4r..4r+3).  Returns Nr.
template <class SBOX>
AG_HD int aes_key_expand_words(const uint8_t* key, int key_bytes, SBOX&& sbox, uint32_t* rk)
{
    const int nk = key_bytes / 4;
    const int nr = nk + 6;
    const int total = 4 * (nr + 1);
    for (int i = 0; i < nk; ++i)
        rk[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) |
                ((uint32_t)key[4 * i + 3] << 24);
    uint32_t rcon = 1;
    for (int i = nk; i < total; ++i) {
        uint32_t t = rk[i - 1];
        if (i % nk == 0) {
            t = (t >> 8) | (t << 24);  // RotWord on bytes (b0,b1,b2,b3) -> (b1,b2,b3,b0)
            t = sbox(t & 0xff) | (sbox((t >> 8) & 0xff) << 8) | (sbox((t >> 16) & 0xff) << 16) | (sbox(t >> 24) << 24);
            t ^= rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11B : 0);
        } else if (nk == 8 && (i % 8) == 4) {
            t = sbox(t & 0xff) | (sbox((t >> 8) & 0xff) << 8) | (sbox((t >> 16) & 0xff) << 16) | (sbox(t >> 24) << 24);
        }
        rk[i] = rk[i - nk] ^ t;
    }
    return nr;
}

// ---- one block, generic input (used for H = E_K(0) and E_K(J0)) -------------
// TE: functor uint32_t(int table, uint32_t word, int byte_k) = Te_table[(word >> 8k) & 0xff]
template <class TE>
AG_HD void aes_encrypt_words(const uint32_t* rk, int nr, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, TE&& te,
                             uint32_t out[4])
{
    s0 ^= rk[0];
    s1 ^= rk[1];
    s2 ^= rk[2];
    s3 ^= rk[3];
    for (int r = 1; r < nr; ++r) {
        uint32_t t0 = te(0, s0, 0) ^ te(1, s1, 1) ^ te(2, s2, 2) ^ te(3, s3, 3) ^ rk[4 * r + 0];
        uint32_t t1 = te(0, s1, 0) ^ te(1, s2, 1) ^ te(2, s3, 2) ^ te(3, s0, 3) ^ rk[4 * r + 1];
        uint32_t t2 = te(0, s2, 0) ^ te(1, s3, 1) ^ te(2, s0, 2) ^ te(3, s1, 3) ^ rk[4 * r + 2];
        uint32_t t3 = te(0, s3, 0) ^ te(1, s0, 1) ^ te(2, s1, 2) ^ te(3, s2, 3) ^ rk[4 * r + 3];
        s0 = t0;
        s1 = t1;
        s2 = t2;
        s3 = t3;
    }
    // last round: SubBytes + ShiftRows only.  S sits in byte 0 of Te2, byte 1 of
    // Te3, byte 2 of Te0 and byte 3 of Te1.
    out[0] = (te(2, s0, 0) & 0x000000ffu) ^ (te(3, s1, 1) & 0x0000ff00u) ^ (te(0, s2, 2) & 0x00ff0000u) ^
             (te(1, s3, 3) & 0xff000000u) ^ rk[4 * nr + 0];
    out[1] = (te(2, s1, 0) & 0x000000ffu) ^ (te(3, s2, 1) & 0x0000ff00u) ^ (te(0, s3, 2) & 0x00ff0000u) ^
             (te(1, s0, 3) & 0xff000000u) ^ rk[4 * nr + 1];
    out[2] = (te(2, s2, 0) & 0x000000ffu) ^ (te(3, s3, 1) & 0x0000ff00u) ^ (te(0, s0, 2) & 0x00ff0000u) ^
             (te(1, s1, 3) & 0xff000000u) ^ rk[4 * nr + 2];
    out[3] = (te(2, s3, 0) & 0x000000ffu) ^ (te(3, s0, 1) & 0x0000ff00u) ^ (te(0, s1, 2) & 0x00ff0000u) ^
             (te(1, s2, 3) & 0xff000000u) ^ rk[4 * nr + 3];
}

// ---- counter mode, specialised (src/aes_icb.vhd:118: block = IV(96) || cnt(32)) --
// Twelve of the sixteen round-1 lookups depend only on (key, IV): fold them and
// the stage-1 key into four per-message constants.
struct AesCtrConst {
    uint32_t k[4];
};

template <class TE>
AG_HD AesCtrConst aes_ctr_precompute(const uint32_t* rk, uint32_t iv0, uint32_t iv1, uint32_t iv2, TE&& te)
{
    const uint32_t s0 = iv0 ^ rk[0], s1 = iv1 ^ rk[1], s2 = iv2 ^ rk[2];
    AesCtrConst c;
    c.k[0] = te(0, s0, 0) ^ te(1, s1, 1) ^ te(2, s2, 2) ^ rk[4];
    c.k[1] = te(0, s1, 0) ^ te(1, s2, 1) ^ te(3, s0, 3) ^ rk[5];
    c.k[2] = te(0, s2, 0) ^ te(2, s0, 2) ^ te(3, s1, 3) ^ rk[6];
    c.k[3] = te(1, s0, 1) ^ te(2, s1, 2) ^ te(3, s2, 3) ^ rk[7];
    return c;
}

// ---- rounds R0 .. NR-1 (16 lookups each: SubBytes+ShiftRows+MixColumns folded into Te_r, then the
// stage key) and the last round (SubBytes+ShiftRows only: S sits in byte 0 of Te2, byte 1 of Te3,
// byte 2 of Te0 and byte 3 of Te1) on a state that already went through rounds 1 .. R0-1.
// R0 / NR compile-time so that the stage keys (kernel parameters = constant bank) become
// instruction operands.  config/config_aes_round.py:120-126, src/aes_last_round.vhd:76.
template <int R0, int NR, class TE>
AG_HD void aes_rounds_from(const uint32_t* rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, TE&& te, uint32_t out[4])
{
#pragma unroll
    for (int r = R0; r < NR; ++r) {
        uint32_t t0 = te(0, s0, 0) ^ te(1, s1, 1) ^ te(2, s2, 2) ^ te(3, s3, 3) ^ rk[4 * r + 0];
        uint32_t t1 = te(0, s1, 0) ^ te(1, s2, 1) ^ te(2, s3, 2) ^ te(3, s0, 3) ^ rk[4 * r + 1];
        uint32_t t2 = te(0, s2, 0) ^ te(1, s3, 1) ^ te(2, s0, 2) ^ te(3, s1, 3) ^ rk[4 * r + 2];
        uint32_t t3 = te(0, s3, 0) ^ te(1, s0, 1) ^ te(2, s1, 2) ^ te(3, s2, 3) ^ rk[4 * r + 3];
        s0 = t0;
        s1 = t1;
        s2 = t2;
        s3 = t3;
    }
    out[0] = (te(2, s0, 0) & 0x000000ffu) ^ (te(3, s1, 1) & 0x0000ff00u) ^ (te(0, s2, 2) & 0x00ff0000u) ^
             (te(1, s3, 3) & 0xff000000u) ^ rk[4 * NR + 0];
    out[1] = (te(2, s1, 0) & 0x000000ffu) ^ (te(3, s2, 1) & 0x0000ff00u) ^ (te(0, s3, 2) & 0x00ff0000u) ^
             (te(1, s0, 3) & 0xff000000u) ^ rk[4 * NR + 1];
    out[2] = (te(2, s2, 0) & 0x000000ffu) ^ (te(3, s3, 1) & 0x0000ff00u) ^ (te(0, s0, 2) & 0x00ff0000u) ^
             (te(1, s1, 3) & 0xff000000u) ^ rk[4 * NR + 2];
    out[3] = (te(2, s3, 0) & 0x000000ffu) ^ (te(3, s0, 1) & 0x0000ff00u) ^ (te(0, s1, 2) & 0x00ff0000u) ^
             (te(1, s2, 3) & 0xff000000u) ^ rk[4 * NR + 3];
}

// keystream block for counter value ctr (host order); NR compile-time so that the
// stage keys (kernel parameters = constant bank) become instruction operands.
template <int NR, class TE>
AG_HD void aes_ctr_block(const uint32_t* rk, const AesCtrConst& cc, uint32_t ctr, TE&& te, uint32_t out[4])
{
    const uint32_t s3i = ag_bswap32(ctr) ^ rk[3];
    uint32_t s0 = cc.k[0] ^ te(3, s3i, 3);
    uint32_t s1 = cc.k[1] ^ te(2, s3i, 2);
    uint32_t s2 = cc.k[2] ^ te(1, s3i, 1);
    uint32_t s3 = cc.k[3] ^ te(0, s3i, 0);
    aes_rounds_from<2, NR>(rk, s0, s1, s2, s3, te, out);
}

// ---- counter mode for a lane whose counter keeps its low byte (and, almost always, its
// high byte) from one block to the next: the grid-wide stream kernel steps the counter by
// Gt = ncta*1024, so bits 0..7 never change and bits 24..31 change once per 2^24 blocks.
// Round-1 columns 0 and 3 then stay put, and so do the eight round-2 lookups that read them:
// they are folded into q[0..3].  Rounds 1+2 cost 10 lookups instead of 20.  The cache is keyed
// on the two counter bytes; a mismatch (also: any grid whose stride is not a multiple of 256)
// just recomputes it.
struct AesCtrCache {
    uint32_t key;    // (s3i & 0xFF0000FF) the cache was built for; s3i = bswap(ctr) ^ rk[3]
    uint32_t q[4];   // round-2 constants incl. rk[8..11]
};

template <class TE>
AG_HD void aes_ctr_cache_fill(const uint32_t* rk, const AesCtrConst& cc, uint32_t s3i, TE&& te, AesCtrCache& c)
{
    const uint32_t a = cc.k[0] ^ te(3, s3i, 3);  // round-1 column 0
    const uint32_t b = cc.k[3] ^ te(0, s3i, 0);  // round-1 column 3
    c.key = s3i & 0xFF0000FFu;
    c.q[0] = te(0, a, 0) ^ te(3, b, 3) ^ rk[8];
    c.q[1] = te(2, b, 2) ^ te(3, a, 3) ^ rk[9];
    c.q[2] = te(1, b, 1) ^ te(2, a, 2) ^ rk[10];
    c.q[3] = te(0, b, 0) ^ te(1, a, 1) ^ rk[11];
}

template <int NR, class TE>
AG_HD void aes_ctr_block_cached(const uint32_t* rk, const AesCtrConst& cc, AesCtrCache& c, uint32_t ctr, TE&& te,
                                uint32_t out[4])
{
    const uint32_t s3i = ag_bswap32(ctr) ^ rk[3];
    if ((s3i & 0xFF0000FFu) != c.key) aes_ctr_cache_fill(rk, cc, s3i, te, c);
    const uint32_t r1 = cc.k[1] ^ te(2, s3i, 2);  // round-1 column 1
    const uint32_t r2 = cc.k[2] ^ te(1, s3i, 1);  // round-1 column 2
    uint32_t s0 = c.q[0] ^ te(1, r1, 1) ^ te(2, r2, 2);
    uint32_t s1 = c.q[1] ^ te(0, r1, 0) ^ te(1, r2, 1);
    uint32_t s2 = c.q[2] ^ te(0, r2, 0) ^ te(3, r1, 3);
    uint32_t s3 = c.q[3] ^ te(2, r1, 2) ^ te(3, r2, 3);
    aes_rounds_from<3, NR>(rk, s0, s1, s2, s3, te, out);
}

// ---- counter mode for a lane that walks CONSECUTIVE (or small-stride) counters of one
// message (the batched kernels): the three upper counter bytes change once per 256 counters,
// so round-1 columns 1..3 and the twelve round-2 lookups that read them are cached per lane;
// rounds 1+2 cost 5 lookups instead of 20.  Keyed on the three upper bytes (low 24 bits of
// s3i = bswap(ctr) ^ rk[3]); a mismatch recomputes (15 lookups).
struct AesCtrSeqCache {
    uint32_t key;
    uint32_t q[4];
};

template <class TE>
AG_HD void aes_ctr_seq_fill(const uint32_t* rk, const AesCtrConst& cc, uint32_t s3i, TE&& te, AesCtrSeqCache& c)
{
    const uint32_t c1 = cc.k[1] ^ te(2, s3i, 2);  // round-1 column 1
    const uint32_t c2 = cc.k[2] ^ te(1, s3i, 1);  // round-1 column 2
    const uint32_t c3 = cc.k[3] ^ te(0, s3i, 0);  // round-1 column 3
    c.key = s3i & 0x00FFFFFFu;
    c.q[0] = te(1, c1, 1) ^ te(2, c2, 2) ^ te(3, c3, 3) ^ rk[8];
    c.q[1] = te(0, c1, 0) ^ te(1, c2, 1) ^ te(2, c3, 2) ^ rk[9];
    c.q[2] = te(3, c1, 3) ^ te(0, c2, 0) ^ te(1, c3, 1) ^ rk[10];
    c.q[3] = te(2, c1, 2) ^ te(3, c2, 3) ^ te(0, c3, 0) ^ rk[11];
}

template <int NR, class TE>
AG_HD void aes_ctr_block_seq(const uint32_t* rk, const AesCtrConst& cc, AesCtrSeqCache& c, uint32_t ctr, TE&& te,
                             uint32_t out[4])
{
    const uint32_t s3i = ag_bswap32(ctr) ^ rk[3];
    if ((s3i & 0x00FFFFFFu) != c.key) aes_ctr_seq_fill(rk, cc, s3i, te, c);
    const uint32_t a = cc.k[0] ^ te(3, s3i, 3);  // round-1 column 0: the only one that moves
    uint32_t s0 = c.q[0] ^ te(0, a, 0);
    uint32_t s1 = c.q[1] ^ te(3, a, 3);
    uint32_t s2 = c.q[2] ^ te(2, a, 2);
    uint32_t s3 = c.q[3] ^ te(1, a, 1);
    aes_rounds_from<3, NR>(rk, s0, s1, s2, s3, te, out);
}

// One name for both per-lane caches, selected by the cache type the caller keeps.
template <int NR, class TE>
AG_HD void aes_ctr_block_auto(const uint32_t* rk, const AesCtrConst& cc, AesCtrSeqCache& c, uint32_t ctr, TE&& te,
                              uint32_t out[4])
{
    aes_ctr_block_seq<NR>(rk, cc, c, ctr, te, out);
}
template <int NR, class TE>
AG_HD void aes_ctr_block_auto(const uint32_t* rk, const AesCtrConst& cc, AesCtrCache& c, uint32_t ctr, TE&& te,
                              uint32_t out[4])
{
    aes_ctr_block_cached<NR>(rk, cc, c, ctr, te, out);
}
