// Granule schedule of the host-buffer pipeline (H2D copy -> fused kernel -> D2H copy per granule, three granules in
// flight).  Plain host C++: shared by capi.cu and the CPU test harness (tests/host_emul.cu).
//
// With granules of one size g the first kernel waits for g bytes to arrive and the last g bytes leave after
// everything else is done: 2 x g / link rate of every call is not overlapped (1.3 ms of 23.8 ms for 2^30 B at
// g = 32 MiB).  Small granules shorten that but pay the per-granule launch and copy set-up in the steady state.  So
// the schedule RAMPS: base, 2 base, 4 base, ... up to the steady size, the steady granules, and the mirror image at
// the end.  The replaced interface has no such notion: the model takes one 16-byte block per call
// (tb/gcm_model.py:24-32).
#pragma once
#include <stdint.h>

// Fills sz[0..k) with the granule sizes of an n-byte range (their sum is n; every size but the last is a multiple of
// 16, so the granules are whole counter blocks) and returns k <= cap.  `peak` = steady granule, `base` = first and
// last granule (both multiples of 16; base == 0 or base >= peak: no ramp).  Falls back to equal granules when the
// ramped list would not fit `cap`; returns 0 only for n == 0 (the caller sizes `peak` so that n / peak <= cap).
static inline uint32_t ag_chunk_schedule(uint64_t n, uint64_t peak, uint64_t base, uint64_t* sz, uint32_t cap)
{
    if (!n || !cap) return 0;
    uint32_t k = 0;
    if (base >= 16 && (base & 15) == 0 && base < peak) {
        uint32_t lmax = 0;
        while ((base << (lmax + 1)) <= peak) ++lmax;
        uint32_t l = lmax;
        uint64_t q = peak;
        for (;;) {   // the longest ramp that leaves room for one steady granule
            q = (l == lmax) ? peak : (base << l);
            const uint64_t side = base * ((1ull << l) - 1);
            if (2 * side + q <= n || l == 0) break;
            --l;
        }
        const uint64_t side = base * ((1ull << l) - 1);
        if (l > 0 && 2 * side + q <= n) {
            const uint64_t mid = n - 2 * side;   // >= q
            const uint64_t m = mid / q, rest = mid % q, rest16 = rest & ~15ull;
            if (2ull * l + m + (rest16 ? 1 : 0) <= cap) {
                for (uint32_t i = 0; i < l; ++i) sz[k++] = base << i;
                if (rest16) sz[k++] = rest16;   // the odd-sized granule goes where the pipeline is full
                for (uint64_t i = 0; i < m; ++i) sz[k++] = q;
                for (uint32_t i = l; i-- > 0;) sz[k++] = base << i;
                sz[k - 1] += rest - rest16;     // a ragged last block stays last
                return k;
            }
        }
    }
    uint64_t left = n;
    while (left && k < cap) {
        const uint64_t b = left < peak ? left : peak;
        sz[k++] = b;
        left -= b;
    }
    return left ? 0 : k;
}
