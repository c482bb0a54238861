// TEST INFRASTRUCTURE (host build of the generated header): the literal multiply-add / butterfly / correction
// sequence that k_idct_tile<2,*> executes (build/idct_baked.h, written by csrc/tools/gen_idct_table.cpp) must equal
// the plain sum  ACC[p][q] = sum_{n in parity class p} Li[q][n] * c[n]  (mod 2^32) for every quadrant sample q.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "idct_baked.h"
int main()
{
    unsigned long long seed = 88172645463325252ull; long bad = 0;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (uint32_t)(seed >> 16); };
    for (int it = 0; it < 20000; it++) {
        int c[64];
        for (int i = 0; i < 64; i++) c[i] = (rnd() % 3 == 0) ? 0 : (int)(short)(rnd() & 0xFFFF);
        if (it < 128) for (int i = 0; i < 64; i++) c[i] = (i == it % 64) ? ((it & 64) ? -32768 : 32767) : 0;
        uint32_t acc[4][16] = {};
#define CO(n) ((uint32_t)c[n])
        JS_BAKED_MACS(acc, CO)
        for (int p = 0; p < 4; p++) for (int q = 0; q < 16; q++) {
            uint32_t ref = 0; const int y = q >> 2, x = q & 3;
            for (int n = 1; n < 64; n++) { const int u = n & 7, v = n >> 3; if (((v & 1) * 2 + (u & 1)) != p) continue; ref += (uint32_t)kBakedLi[(y * 8 + x) * 64 + n] * (uint32_t)c[n]; }
            if (acc[p][q] != ref) bad++;
        }
    }
    // the mirror corrections: full 8x8 output from the quadrant accumulators must equal the direct sum as well
    for (int it = 0; it < 2000; it++) {
        int c[64]; for (int i = 0; i < 64; i++) c[i] = (int)(short)(rnd() & 0xFFFF);
        uint32_t acc[4][16] = {};
        JS_BAKED_MACS(acc, CO)
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            const int q = y * 4 + x;
            const uint32_t a00 = acc[0][q], a01 = acc[1][q], a10 = acc[2][q], a11 = acc[3][q];
            const uint32_t A = a00 + a01, B = a00 - a01, C2 = a10 + a11, D = a10 - a11;
            uint32_t s[4] = {A + C2, B + D, A - C2, B - D};
            const int yx[4] = {y * 8 + x, y * 8 + 7 - x, (7 - y) * 8 + x, (7 - y) * 8 + 7 - x};
            for (int k = 0; k < 4; k++) {
                int t = 0;
#define COI(n) c[n]
                JS_BAKED_CORR_TERM(yx[k], COI, t)
                const uint32_t got = s[k] + (uint32_t)t;
                uint32_t ref = 0; for (int n = 1; n < 64; n++) ref += (uint32_t)kBakedLi[yx[k] * 64 + n] * (uint32_t)c[n];
                if (got != ref) bad++;
            }
        }
    }
    printf("bad=%ld ncorr=%d\n", bad, JS_BAKED_NCORR);
    for (int i = 0; i < 4096; i++) printf("%d\n", kBakedLi[i]);
    return bad != 0;
}
