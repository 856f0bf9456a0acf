/* How long does the RSSI recurrence (rtl_wmbus.c:475-495: r <- 0.6789f |s| + (1 - 0.6789f) r, consumed as (unsigned)r)
 * take to forget its start value?  The demod kernel starts every 32-sample segment 48 samples early from r = 0 instead
 * of carrying r through the batch.  This program measures, on simulated front-end output -- cu8
 * noise + an FSK carrier through the (int) truncation and the length-8 box sum at decimation 2, the reference's
 * arithmetic -- for EVERY sample position n of a long stream: the number of steps T after which a trajectory started
 * from 0 at n - 48 is bit-identical to the true one (once equal they stay equal), its histogram, the fraction of
 * positions with T > 48 (the float differs at the segment's first output), and how many of those differ in the byte
 * the decoders consume.  Plain C, one thread:
 *     gcc -O2 -o /tmp/rssi_merge tools/rssi_merge.c -lm && /tmp/rssi_merge [positions per scenario, default 2e8]
 * TEST INFRASTRUCTURE (a study, not a test). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2];
static uint64_t rnd(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static double gauss(void)
{
    static int have = 0; static double keep;
    if (have) { have = 0; return keep; }
    double u, v, q;
    do { u = (double)(rnd() >> 11) / 4503599627370496.0 - 1.0; v = (double)(rnd() >> 11) / 4503599627370496.0 - 1.0; q = u * u + v * v; } while (q >= 1.0 || q == 0.0);
    const double f = sqrt(-2.0 * log(q) / q);
    keep = v * f; have = 1;
    return u * f;
}
static int trunc_sample(double x)      /* cu8 byte, then (int)(u - 127.5f) as the reference truncates it (:1310-1352) */
{
    long u = lround(x); if (u < 0) u = 0; if (u > 255) u = 255;
    return (int)((float)u - 127.5f);
}
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

#define WARM 48
#define LOOK 160
int main(int argc, char **argv)
{
    const double want = argc > 1 ? atof(argv[1]) : 2e8;
    const struct { double sigma, amp; const char *what; } sc[] = {
        { 8, 0, "noise sigma 8 (the benchmark captures between telegrams)" }, { 8, 90, "sigma 8 + carrier 90 LSB (inside a telegram)" },
        { 1, 0, "noise sigma 1 (a quiet band)" }, { 20, 0, "noise sigma 20" }, { 3, 25, "sigma 3 + carrier 25 LSB" }, { 0.3, 60, "sigma 0.3 + carrier 60 LSB (a clean CW-like signal)" },
        { 0.3, 0, "sigma 0.3, no carrier (the ADC sits on one code: |s| is 0 most of the time, r decays towards 0 without reaching it)" } };
    const float k1 = 0.6789f, k2 = 1.0f - 0.6789f;
    for (unsigned q = 0; q < sizeof sc / sizeof sc[0]; q++) {
        s[0] = 0x9E3779B97F4A7C15ull + q; s[1] = 0xD1B54A32D192ED03ull;
        static float ring[1 << 16];            /* |s| of the last 65536 decimated samples */
        static float rtrue[1 << 16];
        unsigned long long hist[LOOK + 2] = { 0 }, n = 0, late = 0, byte_diff = 0;
        int bi[8] = { 0 }, bq[8] = { 0 }, si = 0, sq = 0, k = 0, maxT = 0;
        double ph = 0; float r = 0;
        unsigned long long m = 0;
        while ((double)n < want) {
            /* produce 32768 new decimated samples, then evaluate the positions whose look-ahead window is complete */
            for (int j = 0; j < 32768; j++, m++) {
                for (int d = 0; d < 2; d++) {   /* two input samples per decimated sample; box of the last 8 input samples */
                    ph += 2 * M_PI * 50e3 * ((m >> 3) & 1 ? 1 : -1) / 1.6e6;     /* +-50 kHz, 100 kchip/s alternating chips */
                    const int vi = trunc_sample(127.4 + sc[q].amp * cos(ph) + sc[q].sigma * gauss());
                    const int vq = trunc_sample(127.4 + sc[q].amp * sin(ph) + sc[q].sigma * gauss());
                    si += vi - bi[k]; sq += vq - bq[k]; bi[k] = vi; bq[k] = vq; k = (k + 1) & 7;
                }
                const float fi = (float)si / 8, fq = (float)sq / 8;               /* moving_average_filter.h:47-53 */
                const float mag = sqrtf(fi * fi + fq * fq);
                r = k1 * mag + k2 * r;
                ring[m & 65535] = mag; rtrue[m & 65535] = r;
            }
            if (m < 65536) continue;
            for (unsigned long long p = m - 32768 - LOOK; p < m - LOOK && (double)n < want; p++, n++) {
                /* cold trajectory from 0, first input = sample p - WARM; T = steps until it equals the true one */
                float c = 0; int T = LOOK + 1;
                for (int t = 0; t < LOOK; t++) {
                    const unsigned long long x = p - WARM + t;
                    c = k1 * ring[x & 65535] + k2 * c;
                    if (bits(c) == bits(rtrue[x & 65535])) { T = t + 1; break; }
                    if (t >= WARM && t < WARM + 32 && (unsigned)c != (unsigned)rtrue[x & 65535]) byte_diff++;
                }
                hist[T]++;
                if (T > WARM) late++;
                if (T > maxT) maxT = T;
            }
        }
        printf("%s: %llu positions, longest T %d%s, T > %d at %llu positions (%.3g), byte differences inside a segment: %llu\n  T histogram (steps: count):",
               sc[q].what, n, maxT, maxT > LOOK ? " (not merged within the look-ahead)" : "", WARM, late, (double)late / (double)n, byte_diff);
        for (int t = 1; t <= LOOK + 1; t++) if (hist[t]) printf(" %d:%llu", t, hist[t]);
        printf("\n");
        fflush(stdout);
    }
    return 0;
}
