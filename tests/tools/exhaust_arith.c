/* Proof by exhaustion for DESIGN.md section 2, items 1 and 6 (CPU build of the device source, tests/hostsim):
 * the discriminator's atan2f -- wmb_atan2f_bounded: no range escapes, table-driven argument reduction, no branch -- sees
 * its operands only through t = |y| / |x| (one IEEE division), the two sign bits and the zero tests.  Its operands are
 * integers below 2^24 in magnitude (wmb_exact.cuh: wmb_discriminator), so t is 0, inf, NaN or a float in
 * [2^-24, 2^24].  This program walks EVERY float t of that interval (48 binades x 2^23 patterns, both ends included)
 * in all four quadrants as (y, x) = (+-t, +-1) -- the division is then exact and t itself is what the function
 * reduces -- and compares, bit for bit:
 *     the device function (CPU build)  ==  the oracle's restatement of fdlibm  ==  this machine's libm atan2f
 * plus the zero / axis cases.  1.6 G triples; about a minute per process, split by binade across processes:
 *     gcc -O2 -o /tmp/exhaust_arith tests/tools/exhaust_arith.c -ldl -lm
 *     /tmp/exhaust_arith tests/hostsim/_build/libwmbus_hostsim.so oracle/_ref/liboracle.so [first_binade last_binade]
 * With `div` as third argument it walks the run-length tracker's division instead:
 * wmb_div_small(wmb_div_pow2(x, 5), n) == x / (32 * n) in C's truncating division for every |x| <= 2^29 and n = 1..8
 * (the bit length is a x256 fixed-point number of samples: |x| stays below 2^27), 8.6 G quotients, half a minute.
 * TEST INFRASTRUCTURE: loads the oracle and the CPU build, never the product library. */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef float (*f2)(float, float);
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float flt(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s libwmbus_hostsim.so liboracle.so [first_binade last_binade]\n", argv[0]); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW), *o = dlopen(argv[2], RTLD_NOW);
    if (!h || !o) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    f2 dev = (f2)dlsym(h, "hostsim_atan2f_bounded"), gen = (f2)dlsym(h, "hostsim_atan2f_general"), orc = (f2)dlsym(o, "orc_atan2f");
    if (!dev || !gen || !orc) { fprintf(stderr, "missing symbol\n"); return 2; }
    if (argc > 3 && !strcmp(argv[3], "div")) {
        typedef int (*i2)(int, int);
        i2 ds = (i2)dlsym(h, "hostsim_div_small"), dp = (i2)dlsym(h, "hostsim_div_pow2");
        if (!ds || !dp) { fprintf(stderr, "missing symbol\n"); return 2; }
        unsigned long long n = 0, bad = 0;
        for (int x = -(1 << 29); x <= (1 << 29); x++) {
            const int x32 = dp(x, 5);
            if (x32 != x / 32 || dp(x, 4) != x / 16) bad++;
            for (int k = 1; k <= 8; k++, n++)
                if (ds(x32, k) != x / (32 * k)) { if (bad++ < 10) printf("MISMATCH %d / (32 * %d): %d\n", x, k, ds(x32, k)); }
        }
        printf("x / (32 n), |x| <= 2^29, n = 1..8: %llu quotients, %llu mismatches\n", n, bad);
        return bad != 0;
    }
    const int b0 = argc > 3 ? atoi(argv[3]) : -24, b1 = argc > 4 ? atoi(argv[4]) : 24;
    unsigned long long n = 0, bad = 0;
    for (int e = b0; e <= b1; e++) {
        const uint32_t lo = (uint32_t)(e + 127) << 23, hi = e == 24 ? lo : lo + 0x7fffffu;   /* 2^24 itself closes the interval */
        for (uint32_t u = lo; u <= hi; u++)
            for (int q = 0; q < 4; q++) {
                const float y = flt(u | (q & 1 ? 0x80000000u : 0u)), x = q & 2 ? -1.0f : 1.0f;
                const uint32_t a = bits(dev(y, x)), b = bits(orc(y, x)), c = bits(atan2f(y, x)), d = bits(gen(y, x));
                n++;
                if (a != b || b != c || d != b) { if (bad++ < 10) printf("MISMATCH y=%a x=%a dev=%08x gen=%08x oracle=%08x libm=%08x\n", y, x, a, d, b, c); }
            }
    }
    /* zero and axis cases: +-0 / +-x, +-y / +-0, +-0 / +-0 over a few magnitudes */
    const float mags[] = { 1.0f, 3.0f, 16777215.0f, 0.00390625f };
    for (int q = 0; q < 4; q++)
        for (unsigned k = 0; k < 4; k++) {
            const float z = q & 1 ? -0.0f : 0.0f, m = q & 2 ? -mags[k] : mags[k], zz = q & 2 ? -0.0f : 0.0f;
            const float ys[3] = { z, m, z }, xs[3] = { m, z, zz };
            for (int j = 0; j < 3; j++) {
                const uint32_t a = bits(dev(ys[j], xs[j])), b = bits(orc(ys[j], xs[j])), c = bits(atan2f(ys[j], xs[j]));
                n++;
                /* 0 / 0: the reference's cargf(0) is libm's atan2f(+-0, +-0) = +-0 or +-pi like any +-0 / x */
                if (a != b || b != c) { if (bad++ < 10) printf("MISMATCH y=%a x=%a dev=%08x oracle=%08x libm=%08x\n", ys[j], xs[j], a, b, c); }
            }
        }
    printf("binades %d..%d: %llu operand pairs, %llu mismatches\n", b0, b1, n, bad);
    return bad != 0;
}
