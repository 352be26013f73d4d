/* Plain-C client of the C ABI (include/aesgcm_b200.h): the IEEE 802.1AE GCM-AES-128 vector the
 * reference README quotes (README.md:251), encrypt + tag then decrypt + verify through the
 * host-buffer entry points.  Built by tests with gcc only; needs a B200 to run.
 *   gcc -std=c99 -Iinclude tests/abi_example.c -L<pkg> -laesgcm_b200 -Wl,-rpath,<pkg> */
#include <stdio.h>
#include <string.h>
#include "aesgcm_b200.h"

static int unhex(const char* s, uint8_t* out)
{
    int n = 0;
    for (; s[0] && s[1]; s += 2, ++n) {
        unsigned v;
        sscanf(s, "%2x", &v);
        out[n] = (uint8_t)v;
    }
    return n;
}

int main(void)
{
    uint8_t key[16], iv[12], aad[64], pt[64], ct[64], back[64], tag[16], want_tag[16], want_ct[64];
    unhex("AD7A2BD03EAC835A6F620FDCB506B345", key);
    unhex("12153524C0895E81B2C28465", iv);
    const int aad_len = unhex("D609B1F056637A0D46DF998D88E52E00B2C2846512153524C0895E81", aad);
    const int n = unhex("08000F101112131415161718191A1B1C1D1E1F202122232425262728292A2B2C2D2E2F303132333435363738393A0002", pt);
    unhex("701AFA1CC039C0D765128A665DAB69243899BF7318CCDC81C9931DA17FBE8EDD7D17CB8B4C26FC81E3284F2B7FBA713D", want_ct);
    unhex("4F8D55E7D3F06FD5A13C0C29B9D5B880", want_tag);

    agcm_ctx* ctx = NULL;
    int rc = agcm_ctx_create(&ctx, 0);
    if (rc) { printf("ctx: %s\n", agcm_strerror(rc)); return 2; }
    rc = agcm_set_key(ctx, 128, 0, key, sizeof key);
    if (rc) { printf("key: %s\n", agcm_strerror(rc)); return 2; }
    int ok = 0;
    rc = agcm_stream_crypt_host(ctx, 0, iv, aad, (uint64_t)aad_len, pt, ct, (uint64_t)n, tag, &ok);
    if (rc || memcmp(ct, want_ct, (size_t)n) || memcmp(tag, want_tag, 16)) { printf("encrypt mismatch (rc=%d)\n", rc); return 1; }
    rc = agcm_stream_crypt_host(ctx, 1, iv, aad, (uint64_t)aad_len, ct, back, (uint64_t)n, tag, &ok);
    if (rc || !ok || memcmp(back, pt, (size_t)n)) { printf("decrypt mismatch (rc=%d ok=%d)\n", rc, ok); return 1; }
    tag[3] ^= 1;
    rc = agcm_stream_crypt_host(ctx, 1, iv, aad, (uint64_t)aad_len, ct, back, (uint64_t)n, tag, &ok);
    if (rc || ok) { printf("forged tag accepted\n"); return 1; }
    /* an 8-byte IV (McGrew-Viega test case 5): J0 is derived on the device */
    {
        uint8_t k5[16], iv5[8], a5[20], p5[60], c5[60], t5[16], want_t5[16];
        unhex("feffe9928665731c6d6a8f9467308308", k5);
        unhex("cafebabefacedbad", iv5);
        unhex("feedfacedeadbeeffeedfacedeadbeefabaddad2", a5);
        unhex("d9313225f88406e5a55909c5aff5269a86a7a9531534f7da2e4c303d8a318a721c3c0c95956809532fcf0e2449a6b525b16aedf5aa0de657ba637b39", p5);
        unhex("3612d2e79e3b0785561be14aaca2fccb", want_t5);
        rc = agcm_set_key(ctx, 128, 0, k5, sizeof k5);
        if (!rc) rc = agcm_stream_crypt_iv_host(ctx, 0, iv5, sizeof iv5, a5, sizeof a5, p5, c5, sizeof p5, t5, &ok);
        if (rc || memcmp(t5, want_t5, 16) || c5[0] != 0x61 || c5[1] != 0x35) { printf("8-byte IV mismatch (rc=%d)\n", rc); return 1; }
        rc = agcm_set_key(ctx, 128, 0, key, sizeof key);
        if (rc) return 2;
    }
    uint8_t rk[176];
    rc = agcm_key_expand_host(ctx, 128, key, rk);
    if (rc || memcmp(rk, key, 16)) { printf("key_expand\n"); return 1; }
    agcm_ctx_destroy(ctx);
    printf("abi_example ok\n");
    return 0;
}
