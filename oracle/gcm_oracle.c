/*
 * oracle/gcm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A slow, plain-C, byte-at-a-time CPU restatement of the AES-GCM datapath of
 * BLu85/AES-GCM-128-192-256-bits.  It is the checker for the CUDA engine: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product library never links it and has
 * no CPU fallback.
 *
 * Every function names the reference file:line it restates (paths relative to
 * the reference checkout).  No table or code is copied from the reference: the
 * S-box is derived from its definition (GF(2^8) inverse + affine map) and
 * checked in tests/ against the fixtures generated from tb/key_exp.py.
 *
 * Parity pinning (see DESIGN.md "Oracle"): the reference stores no expected
 * outputs; its golden model is pycryptodome AES.new(key, MODE_GCM, nonce=iv)
 * (tb/gcm_model.py:1,18), which is not installed here.  This oracle is pinned
 * against (1) tests/golden/key_exp_vectors.json, produced by importing the
 * reference's own tb/key_exp.py, (2) the SP 800-38D / 802.1AE known-answer
 * vectors whose inputs the reference README quotes (README.md:249-258), and
 * (3) OpenSSL (python `cryptography`) on random cases, live in tests/.
 *
 * Conventions: blocks are 16 bytes, big-endian; state(i)(j) = column i, row j
 * = byte 4i+j (src/aes_func.vhd:85-93).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* GF(2^8), polynomial x^8+x^4+x^3+x+1 (0x11B): src/aes_func.vhd:187-200      */
static uint8_t xtime2(uint8_t a) { return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0x00)); }
/* src/aes_func.vhd:205-210 */
static uint8_t xtime3(uint8_t a) { return (uint8_t)(xtime2(a) ^ a); }

static uint8_t gf256_mul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) r ^= a;
        a = xtime2(a);
        b >>= 1;
    }
    return r;
}

/* Rijndael S-box, by definition (FIPS-197 5.1.1): multiplicative inverse then the
 * affine map.  Value-identical to the 256-way case at src/aes_func.vhd:228-301
 * and the list at tb/key_exp.py:23-54 (asserted in tests/test_oracle.py against
 * fixtures generated from tb/key_exp.py). */
static uint8_t SBOX[256];
static int sbox_ready = 0;
static void sbox_init(void)
{
    if (sbox_ready) return;
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) {
            for (int y = 1; y < 256; y++)
                if (gf256_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        }
        uint8_t s = 0;
        for (int i = 0; i < 8; i++) {
            int bit = ((inv >> i) ^ (inv >> ((i + 4) & 7)) ^ (inv >> ((i + 5) & 7)) ^
                       (inv >> ((i + 6) & 7)) ^ (inv >> ((i + 7) & 7)) ^ (0x63 >> i)) & 1;
            s |= (uint8_t)(bit << i);
        }
        SBOX[x] = s;
    }
    sbox_ready = 1;
}

ORACLE_API void oracle_sbox_table(uint8_t out[256])
{
    sbox_init();
    memcpy(out, SBOX, 256);
}

/* ------------------------------------------------------------------------ */
/* Key schedule: tb/key_exp.py:79-114 (and the streamed variant in
 * config/config_aes_kexp.py:128-159).  key_bytes in {16,24,32}; out receives
 * (Nr+1)*16 = 176/208/240 bytes, stage r at bytes 16r..16r+15.  Returns Nr. */
ORACLE_API int oracle_key_expand(const uint8_t *key, int key_bytes, uint8_t *out)
{
    sbox_init();
    int n_stages;
    if (key_bytes == 16) n_stages = 11;
    else if (key_bytes == 24) n_stages = 13;
    else if (key_bytes == 32) n_stages = 15;
    else return -1;
    int total = n_stages * 16;
    memcpy(out, key, (size_t)key_bytes);                 /* key_exp.py:92-95 */
    uint8_t rcon = 0x01;                                 /* key_exp.py:58, doubled per use (config_aes_kexp.py:150) */
    int cur = key_bytes;
    while (cur < total) {
        uint8_t v[4];
        memcpy(v, out + cur - 4, 4);                     /* key_exp.py:100 */
        if (cur % key_bytes == 0) {                      /* key_exp.py:102-104: RotWord, SubWord, Rcon */
            uint8_t t = v[0];
            v[0] = SBOX[v[1]] ^ rcon;
            v[1] = SBOX[v[2]];
            v[2] = SBOX[v[3]];
            v[3] = SBOX[t];
            rcon = xtime2(rcon);
        }
        if (key_bytes == 32 && (cur % key_bytes) == 16) {/* key_exp.py:107-108: extra SubWord */
            for (int i = 0; i < 4; i++) v[i] = SBOX[v[i]];
        }
        for (int i = 0; i < 4; i++) {                    /* key_exp.py:110-112 */
            out[cur] = out[cur - key_bytes] ^ v[i];
            cur++;
        }
    }
    return n_stages - 1;
}

/* ------------------------------------------------------------------------ */
/* One block of AES encryption with an already expanded key.
 * Round phasing follows config/config_aes_round.py:120-126,142:
 *   for cnt in 1..Nr: ARK(stage cnt-1) -> sub_byte -> shift_row -> mix_columns
 *   (mix_columns skipped when cnt == Nr), then src/aes_last_round.vhd:76:
 *   final ARK(stage Nr). */
ORACLE_API void oracle_aes_encrypt_block(const uint8_t *rk, int nr, const uint8_t in[16], uint8_t out[16])
{
    sbox_init();
    uint8_t s[16], t[16];
    memcpy(s, in, 16);
    for (int cnt = 1; cnt <= nr; cnt++) {
        const uint8_t *k = rk + 16 * (cnt - 1);
        for (int i = 0; i < 16; i++) s[i] ^= k[i];                 /* add_round_key: aes_func.vhd:122-131 */
        for (int i = 0; i < 16; i++) s[i] = SBOX[s[i]];            /* sub_byte: aes_func.vhd:108-117 */
        for (int c = 0; c < 4; c++)                                /* shift_row: aes_func.vhd:146-154 */
            for (int r = 0; r < 4; r++)
                t[4 * c + r] = s[4 * ((c + r) & 3) + r];
        if (cnt != nr) {                                           /* mix_columns: aes_func.vhd:159-169 */
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = xtime2(a0) ^ xtime3(a1) ^ a2 ^ a3;
                s[4 * c + 1] = a0 ^ xtime2(a1) ^ xtime3(a2) ^ a3;
                s[4 * c + 2] = a0 ^ a1 ^ xtime2(a2) ^ xtime3(a3);
                s[4 * c + 3] = xtime3(a0) ^ a1 ^ a2 ^ xtime2(a3);
            }
        } else {
            memcpy(s, t, 16);
        }
    }
    const uint8_t *k = rk + 16 * nr;                               /* aes_last_round.vhd:76 */
    for (int i = 0; i < 16; i++) out[i] = s[i] ^ k[i];
}

/* ------------------------------------------------------------------------ */
/* GF(2^128) multiply: src/ghash_gfmul.vhd:42-63 (SP 800-38D Algorithm 1).
 * Bit 127 of the VHDL vector = MSB of byte 0 = coefficient x^0.
 * V starts as H; for each bit of X from the leftmost: Z ^= V if the bit is set;
 * V <- (V >> 1) xor (0xE1 || 0^120 if the dropped bit was 1). */
ORACLE_API void oracle_gfmul(const uint8_t h[16], const uint8_t x[16], uint8_t y[16])
{
    uint8_t v[16], z[16];
    memcpy(v, h, 16);
    memset(z, 0, 16);
    for (int i = 0; i < 128; i++) {
        int xbit = (x[i >> 3] >> (7 - (i & 7))) & 1;
        if (xbit)
            for (int j = 0; j < 16; j++) z[j] ^= v[j];
        int lsb = v[15] & 1;
        for (int j = 15; j > 0; j--) v[j] = (uint8_t)((v[j] >> 1) | (v[j - 1] << 7));
        v[0] >>= 1;
        if (lsb) v[0] ^= 0xE1;
    }
    memcpy(y, z, 16);
}

/* H^e by square-and-multiply (test helper for shard scaling; e >= 0; H^0 = 1 =
 * 0x80 || 0^120 in GCM bit order). */
ORACLE_API void oracle_gf_pow(const uint8_t h[16], uint64_t e, uint8_t out[16])
{
    uint8_t r[16], b[16];
    memset(r, 0, 16);
    r[0] = 0x80;
    memcpy(b, h, 16);
    while (e) {
        if (e & 1) oracle_gfmul(b, r, r);
        oracle_gfmul(b, b, b);
        e >>= 1;
    }
    memcpy(out, r, 16);
}

/* GHASH absorb: src/gcm_ghash.vhd:225-272.  Y <- gfmul(H, (X & mask) xor Y) per
 * 16-byte word; a short last word is left-aligned and zero-padded by the byte
 * mask (gcm_ghash.vhd:228-246,261). */
ORACLE_API void oracle_ghash_absorb(const uint8_t h[16], const uint8_t *data, uint64_t len, uint8_t y[16])
{
    uint8_t x[16];
    while (len) {
        size_t n = len < 16 ? (size_t)len : 16;
        memset(x, 0, 16);
        memcpy(x, data, n);
        for (int j = 0; j < 16; j++) x[j] ^= y[j];
        oracle_gfmul(h, x, y);
        data += n;
        len -= n;
    }
}

static void put_be64(uint8_t *p, uint64_t v)
{
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i));
}

static void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

/* Intermediates for debugging and for shard tests: H = E_K(0^128)
 * (src/gcm_gctr.vhd:141-144, gcm_ghash.vhd:128-139) and E_K(J0) with
 * J0 = IV || 00000001 (src/aes_icb.vhd:34,99,118; gcm_ghash.vhd:158-169). */
ORACLE_API void oracle_h_ej0(const uint8_t *rk, int nr, const uint8_t iv[12], uint8_t h[16], uint8_t ej0[16])
{
    uint8_t blk[16];
    memset(blk, 0, 16);
    oracle_aes_encrypt_block(rk, nr, blk, h);
    memcpy(blk, iv, 12);
    put_be32(blk + 12, 1);
    oracle_aes_encrypt_block(rk, nr, blk, ej0);
}

/* GCTR over a byte range: out[i] = in[i] xor E_K(IV || (first_ctr + i/16 mod 2^32))
 * (src/aes_icb.vhd:100,118: only the low 32 bits count; src/gcm_gctr.vhd:150).
 * The first data block of a message uses first_ctr = 2. */
ORACLE_API void oracle_gctr(const uint8_t *rk, int nr, const uint8_t iv[12], uint32_t first_ctr,
                            const uint8_t *in, uint64_t len, uint8_t *out)
{
    uint8_t cb[16], ks[16];
    memcpy(cb, iv, 12);
    uint32_t ctr = first_ctr;
    uint64_t off = 0;
    while (off < len) {
        put_be32(cb + 12, ctr);
        oracle_aes_encrypt_block(rk, nr, cb, ks);
        size_t n = (len - off) < 16 ? (size_t)(len - off) : 16;
        for (size_t j = 0; j < n; j++) out[off + j] = in[off + j] ^ ks[j];
        off += n;
        ctr++;
    }
}

/* Whole-message AES-GCM with a 96-bit IV.
 *   key_len 16/24/32 = raw key; 176/208/240 = pre-expanded stages
 *   (config/config_aes_kprexp.py:66-95: Nr+1 user-loaded stages).
 *   decrypt = 0: out = CT, tag_out = computed tag.
 *   decrypt = 1: out = PT, tag_out = computed tag (the IP always emits the
 *   computed tag, src/aes_gcm.vhd:207-211; comparing is the caller's job,
 *   tb/gcm_model.py:42-51).
 * GHASH input is the ciphertext in both directions (aes_gcm.vhd:207-211);
 * the length block is [len(A)]64 || [len(C)]64 in bits (gcm_ghash.vhd:257);
 * TAG = Y xor E_K(J0) (gcm_ghash.vhd:293).
 * Returns 0, or -1 for a bad key length, -2 for > 2^32-2 blocks. */
ORACLE_API int oracle_gcm_crypt(const uint8_t *key, int key_len, const uint8_t iv[12],
                                const uint8_t *aad, uint64_t aad_len,
                                const uint8_t *in, uint64_t len, int decrypt,
                                uint8_t *out, uint8_t tag_out[16])
{
    uint8_t rk[240];
    int nr;
    if ((len + 15) / 16 > 0xFFFFFFFEull) return -2;
    if (key_len == 16 || key_len == 24 || key_len == 32) {
        nr = oracle_key_expand(key, key_len, rk);
    } else if (key_len == 176 || key_len == 208 || key_len == 240) {
        nr = key_len / 16 - 1;
        memcpy(rk, key, (size_t)key_len);
    } else {
        return -1;
    }
    uint8_t h[16], ej0[16], y[16], lenblk[16];
    oracle_h_ej0(rk, nr, iv, h, ej0);
    memset(y, 0, 16);
    oracle_ghash_absorb(h, aad, aad_len, y);
    if (decrypt) oracle_ghash_absorb(h, in, len, y);
    oracle_gctr(rk, nr, iv, 2, in, len, out);
    if (!decrypt) oracle_ghash_absorb(h, out, len, y);
    put_be64(lenblk, aad_len * 8);
    put_be64(lenblk + 8, len * 8);
    oracle_ghash_absorb(h, lenblk, 16, y);
    for (int j = 0; j < 16; j++) tag_out[j] = y[j] ^ ej0[j];
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Batched driver used by tests (many small messages) and by bench.py's CPU
 * baseline: messages at in_off[i]..in_off[i+1], AAD likewise, one key per
 * message (key_stride = key_len) or one shared key (key_stride = 0), spread
 * over n_threads POSIX threads. */
typedef struct {
    const uint8_t *keys; int key_len; size_t key_stride;
    const uint8_t *ivs;
    const uint8_t *aad; const uint64_t *aad_off;
    const uint8_t *in; const uint64_t *in_off;
    int decrypt; uint8_t *out; uint8_t *tags;
    size_t lo, hi; int rc;
} batch_job_t;

static void *batch_worker(void *arg)
{
    batch_job_t *j = (batch_job_t *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        uint64_t a0 = j->aad_off ? j->aad_off[i] : 0, a1 = j->aad_off ? j->aad_off[i + 1] : 0;
        uint64_t d0 = j->in_off[i], d1 = j->in_off[i + 1];
        int rc = oracle_gcm_crypt(j->keys + i * j->key_stride, j->key_len, j->ivs + 12 * i,
                                  j->aad ? j->aad + a0 : NULL, a1 - a0,
                                  j->in + d0, d1 - d0, j->decrypt, j->out + d0, j->tags + 16 * i);
        if (rc) j->rc = rc;
    }
    return NULL;
}

ORACLE_API int oracle_gcm_batch(const uint8_t *keys, int key_len, size_t key_stride,
                                const uint8_t *ivs,
                                const uint8_t *aad, const uint64_t *aad_off,
                                const uint8_t *in, const uint64_t *in_off,
                                int decrypt, uint8_t *out, uint8_t *tags,
                                size_t n_msgs, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_msgs && n_msgs > 0) n_threads = (int)n_msgs;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_job_t *jobs = (batch_job_t *)malloc(sizeof(batch_job_t) * (size_t)n_threads);
    size_t per = (n_msgs + (size_t)n_threads - 1) / (size_t)n_threads;
    for (int t = 0; t < n_threads; t++) {
        batch_job_t *j = &jobs[t];
        j->keys = keys; j->key_len = key_len; j->key_stride = key_stride; j->ivs = ivs;
        j->aad = aad; j->aad_off = aad_off; j->in = in; j->in_off = in_off;
        j->decrypt = decrypt; j->out = out; j->tags = tags; j->rc = 0;
        j->lo = (size_t)t * per; j->hi = j->lo + per;
        if (j->lo > n_msgs) j->lo = n_msgs;
        if (j->hi > n_msgs) j->hi = n_msgs;
        pthread_create(&th[t], NULL, batch_worker, j);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(th); free(jobs);
    return rc;
}

/* One large stream split into n_threads counter-range shards (the CPU analogue
 * of SURVEY 8(e) regime 2): each thread runs GCTR on its shard, then GHASH is
 * absorbed serially (it is a serial recurrence in the reference,
 * gcm_ghash.vhd:269-272).  Used only as the multi-threaded CPU baseline. */
typedef struct {
    const uint8_t *rk; int nr; const uint8_t *iv; uint32_t first_ctr;
    const uint8_t *in; uint64_t len; uint8_t *out;
} gctr_job_t;

static void *gctr_worker(void *arg)
{
    gctr_job_t *j = (gctr_job_t *)arg;
    oracle_gctr(j->rk, j->nr, j->iv, j->first_ctr, j->in, j->len, j->out);
    return NULL;
}

typedef struct {
    const uint8_t *h; const uint8_t *data; uint64_t len; uint8_t y[16];
} ghash_job_t;

static void *ghash_worker(void *arg)
{
    ghash_job_t *j = (ghash_job_t *)arg;
    memset(j->y, 0, 16);
    oracle_ghash_absorb(j->h, j->data, j->len, j->y);
    return NULL;
}

ORACLE_API int oracle_gcm_stream_mt(const uint8_t *key, int key_len, const uint8_t iv[12],
                                    const uint8_t *aad, uint64_t aad_len,
                                    const uint8_t *in, uint64_t len, int decrypt,
                                    uint8_t *out, uint8_t tag_out[16], int n_threads)
{
    uint8_t rk[240];
    int nr;
    if (key_len == 16 || key_len == 24 || key_len == 32) nr = oracle_key_expand(key, key_len, rk);
    else if (key_len == 176 || key_len == 208 || key_len == 240) { nr = key_len / 16 - 1; memcpy(rk, key, (size_t)key_len); }
    else return -1;
    uint64_t n_blocks = (len + 15) / 16;
    if (n_blocks > 0xFFFFFFFEull) return -2;
    if (n_threads < 1) n_threads = 1;
    uint8_t h[16], ej0[16];
    oracle_h_ej0(rk, nr, iv, h, ej0);

    uint64_t per = ((n_blocks + (uint64_t)n_threads - 1) / (uint64_t)n_threads) * 16;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    gctr_job_t *gj = (gctr_job_t *)malloc(sizeof(gctr_job_t) * (size_t)n_threads);
    ghash_job_t *hj = (ghash_job_t *)malloc(sizeof(ghash_job_t) * (size_t)n_threads);
    int used = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t lo = (uint64_t)t * per;
        if (lo >= len) break;
        uint64_t n = (len - lo) < per ? (len - lo) : per;
        gj[t].rk = rk; gj[t].nr = nr; gj[t].iv = iv; gj[t].first_ctr = (uint32_t)(2 + lo / 16);
        gj[t].in = in + lo; gj[t].len = n; gj[t].out = out + lo;
        pthread_create(&th[t], NULL, gctr_worker, &gj[t]);
        used++;
    }
    for (int t = 0; t < used; t++) pthread_join(th[t], NULL);

    /* GHASH: per-shard partials in parallel, combined with H^(blocks after the shard)
     * (linearity precedent: gcm_ghash.vhd:317-344). */
    const uint8_t *ct = decrypt ? in : out;
    for (int t = 0; t < used; t++) {
        uint64_t lo = (uint64_t)t * per;
        uint64_t n = (len - lo) < per ? (len - lo) : per;
        hj[t].h = h; hj[t].data = ct + lo; hj[t].len = n;
        pthread_create(&th[t], NULL, ghash_worker, &hj[t]);
    }
    for (int t = 0; t < used; t++) pthread_join(th[t], NULL);

    uint8_t y[16], lenblk[16], hp[16], tmp[16];
    memset(y, 0, 16);
    oracle_ghash_absorb(h, aad, aad_len, y);
    /* y currently weights AAD as if nothing followed; scale by H^(n_blocks) */
    oracle_gf_pow(h, n_blocks, hp);
    oracle_gfmul(hp, y, y);
    for (int t = 0; t < used; t++) {
        uint64_t lo = (uint64_t)t * per;
        uint64_t n = (len - lo) < per ? (len - lo) : per;
        uint64_t blocks_after = n_blocks - (lo / 16 + (n + 15) / 16);
        oracle_gf_pow(h, blocks_after, hp);
        oracle_gfmul(hp, hj[t].y, tmp);
        for (int j = 0; j < 16; j++) y[j] ^= tmp[j];
    }
    put_be64(lenblk, aad_len * 8);
    put_be64(lenblk + 8, len * 8);
    oracle_ghash_absorb(h, lenblk, 16, y);
    for (int j = 0; j < 16; j++) tag_out[j] = y[j] ^ ej0[j];
    free(th); free(gj); free(hj);
    return 0;
}
