/*
 * wmbus_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see wmbus_oracle.h).
 *
 * Array-at-a-time restatement of rtl-wmbus's DSP + bit-sync + framing path.
 * Compile with -O2 -ffp-contract=off (no FMA contraction: the reference's
 * x86-64 build rounds every multiply and add separately).
 *
 * Parity status: PINNED -- checked line-for-line against the unmodified
 * reference binary (oracle/_ref/rtl_wmbus) on all four sample captures and on
 * synthetic T1 / C1-A / C1-B / S1 captures (tests/test_oracle_vs_ref.py), and
 * stage-for-stage against the reference's own static functions through
 * oracle/ref_stages.c.
 */
#include "wmbus_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* helpers                                                                   */
/* ------------------------------------------------------------------------- */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

void orc_default_opts(orc_opts *o)
{
    memset(o, 0, sizeof(*o));
    o->decimation = 2;          /* rtl_wmbus.c:857 */
    o->accurate_atan = 1;       /* :859 */
    o->rla_enabled = 1;         /* :855 */
    o->t2_enabled = 1;          /* :856 */
    o->t1c1_enabled = 1;        /* :862 */
    o->s1_enabled = 1;          /* :863 */
}

static inline uint32_t eff_decimation(uint32_t d) { return d ? d : 1u; }

size_t orc_num_decimated(size_t n_iq, uint32_t decimation)
{
    /* rtl_wmbus.c:1350-1352: a sample is kept when the running index reaches d;
     * d == 0 or 1 keeps every sample. */
    return n_iq / eff_decimation(decimation);
}

/* ------------------------------------------------------------------------- */
/* atan2f -- fdlibm e_atan2f.c / s_atanf.c as shipped in glibc 2.39          */
/* (reference call site: atan2.h:9  cargf(y) * (float)M_1_PI)                */
/* ------------------------------------------------------------------------- */

static float orc_atanf(float x)
{
    static const uint32_t hi_bits[4] = { 0x3eed6338u, 0x3f490fdau, 0x3f7b985eu, 0x3fc90fdau };
    static const uint32_t lo_bits[4] = { 0x31ac3769u, 0x33222168u, 0x33140fb4u, 0x33a22168u };
    /* NB: aT[0] is 0x3eaaaaab (what the decimal literal 3.3333334327e-01 parses to, and what
     * libm.so.6 holds), not the 0x3eaaaaaa of the fdlibm source comment. */
    static const uint32_t at_bits[11] = {
        0x3eaaaaabu, 0xbe4ccccdu, 0x3e124925u, 0xbde38e38u, 0x3dba2e6eu, 0xbd9d8795u,
        0x3d886b35u, 0xbd6ef16bu, 0x3d4bda59u, 0xbd15a221u, 0x3c8569d7u };
    float aT[11];
    for (int i = 0; i < 11; i++) aT[i] = u2f(at_bits[i]);

    const uint32_t hx = f2u(x);
    const uint32_t ix = hx & 0x7fffffffu;
    int id;

    if (ix >= 0x4c000000u) {                 /* |x| >= 2^25 */
        if (ix > 0x7f800000u) return x + x;  /* NaN */
        const float r = u2f(hi_bits[3]) + u2f(lo_bits[3]);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {                  /* |x| < 0.4375 */
        if (ix < 0x31000000u) return x;      /* |x| < 2^-29 */
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {              /* |x| < 1.1875 */
            if (ix < 0x3f300000u) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else                  { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else                  { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = u2f(hi_bits[id]) - ((x * (s1 + s2) - u2f(lo_bits[id])) - x);
    return (hx >> 31) ? -r : r;
}

float orc_atan2f(float y, float x)
{
    const float pi = u2f(0x40490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi_o_4 = u2f(0x3f490fdbu);
    const float pi_lo = u2f(0xb3bbbd2eu);
    const uint32_t hx = f2u(x), hy = f2u(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;

    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;      /* NaN */
    if (hx == 0x3f800000u) return orc_atanf(y);                  /* x == 1 */
    const unsigned m = ((hy >> 31) & 1u) | ((hx >> 30) & 2u);    /* 2*sign(x)+sign(y) */

    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi;      /* pi + tiny */
        default: return -pi;    /* -pi - tiny */
        }
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 : pi_o_2;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
            case 0: return pi_o_4;
            case 1: return -pi_o_4;
            case 2: return 3.0f * pi_o_4;
            default: return -3.0f * pi_o_4;
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi;
        default: return -pi;
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 : pi_o_2;

    const int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = orc_atanf(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

/* ------------------------------------------------------------------------- */
/* A.1-A.3  convert, optional +-325 kHz mix, integer box filter, decimate     */
/* rtl_wmbus.c:1310-1352, :974-1031 ; moving_average_filter.h:47-54           */
/* ------------------------------------------------------------------------- */

void orc_frontend(const uint8_t *cu8, size_t n_iq, const orc_opts *o, int chain,
                  float *si, float *sq)
{
    const uint32_t d = eff_decimation(o->decimation);
    const int L = (chain == ORC_CHAIN_T1C1) ? 8 : 16;       /* rtl_wmbus.c:167, :183 */
    int ring_i[16] = {0}, ring_q[16] = {0};
    int sum_i = 0, sum_q = 0, pos = 0;
    size_t m = 0;
    uint32_t since = 0;

    /* mixer LUT (only with -s): cosf / -sinf of 2*pi*25*n/fs_kHz, n < fs_kHz/25 */
    const size_t n_max = (size_t)(o->decimation * 800u) / 25u;
    float *lut_c = NULL, *lut_s = NULL;
    size_t n = 0;
    /* -s: both chains step 13 entries per sample, T1/C1 multiplies by the entry, S1 by its conjugate; simultaneous == 2
     * takes both from the chain's carrier offset */
    int32_t off = (chain == ORC_CHAIN_T1C1) ? 13 : -13;
    if (o->simultaneous == 2) off = o->carrier_25khz[chain];
    const int mix_conj = off < 0;
    const size_t mix_step = n_max ? (size_t)(off < 0 ? -(int64_t)off : (int64_t)off) % n_max : 0;
    if (o->simultaneous && n_max) {
        lut_c = malloc(n_max * sizeof(float));
        lut_s = malloc(n_max * sizeof(float));
        const int fs_khz = (int)(o->decimation * 800u);
        for (size_t j = 0; j < n_max; j++) {
            const double phi = (2. * M_PI * (25 * j)) / fs_khz;   /* :989 */
            lut_c[j] = cosf(phi);
            lut_s[j] = -sinf(phi);
        }
    }

    /* the dormant pre-decimation low-passes (rtl_wmbus.c:197-333), one 23-tap design in four arithmetics:
     *   1  lp_fir_butter_1600kHz_160kHz_200kHz_{t1_c1,s1} (:197-233, the same coefficients for both chains) through firf()
     *      (fir.h:49-72): y = sum_j b[j] x[n-j], accumulated from 0 in that order, zero history at the start
     *   2  lp_ppf_butter_1600kHz_160kHz_200kHz (:258-295) through ppf() (ppf.h:44-58): two 12-tap firf()s, the even-phase
     *      samples through the odd coefficients (+ a zero tap), the odd-phase samples through the even ones
     *      ("!inverted indexing of fir!"); sum = 0 at phase 0, sum += firf() at each phase; read after phase 1
     *   3  lp_firfp_butter_1600kHz_160kHz_200kHz (:235-256) through firfp() (fir.h:106-130): 24.8 fixed point
     *      (fixedptc.h: FIXEDPT_BITS 32, FIXEDPT_WBITS 24), sample = fixedpt_fromint(float) = (int64)x << 8 cut to 32 bits,
     *      taps fixedpt_rconst(b) = (int32)(b * 256 + 0.5), products (int64)b * h >> 8, back through fixedpt_tofloat
     *   4  lp_ppffp_butter_1600kHz_160kHz_200kHz (:297-333) through ppffp() (ppf.h:69-83): 2 in the arithmetic of 3 */
    static const float pre_b[23] = {
        0.000140535927, 1.102280392e-05, 0.0001309279731, 0.001356012537, 0.00551787474, 0.01499414005, 0.03160167988,
        0.05525973093, 0.08315031015, 0.1099887688, 0.1295143636, 0.1366692652, 0.1295143636, 0.1099887688, 0.08315031015,
        0.05525973093, 0.03160167988, 0.01499414005, 0.00551787474, 0.001356012537, 0.0001309279731, 1.102280392e-05,
        0.000140535927 };
    static const double pre_bd[23] = {                      /* the literals as the fixedpt_rconst() macro sees them: doubles */
        0.000140535927, 1.102280392e-05, 0.0001309279731, 0.001356012537, 0.00551787474, 0.01499414005, 0.03160167988,
        0.05525973093, 0.08315031015, 0.1099887688, 0.1295143636, 0.1366692652, 0.1295143636, 0.1099887688, 0.08315031015,
        0.05525973093, 0.03160167988, 0.01499414005, 0.00551787474, 0.001356012537, 0.0001309279731, 1.102280392e-05,
        0.000140535927 };
    int32_t pre_bx[24];
    for (int j = 0; j < 23; j++) pre_bx[j] = (int32_t)(pre_bd[j] * 256 + (pre_bd[j] >= 0 ? 0.5 : -0.5));
    pre_bx[23] = (int32_t)(0 * 256 + 0.5);                  /* fixedpt_rconst(0), the polyphase branch's zero tap */
    float pre_i[24] = {0}, pre_q[24] = {0};                 /* [0] = newest */
    int32_t prx_i[24] = {0}, prx_q[24] = {0};

    for (size_t k = 0; k < n_iq; k++) {
        float xi = (float)cu8[2 * k] - 127.5f;              /* :1312 */
        float xq = (float)cu8[2 * k + 1] - 127.5f;          /* :1313 */
        if (lut_c) {
            const float c = lut_c[n], z = lut_s[n];
            n += mix_step;                                  /* 325/25 = 13, :1008 */
            if (n >= n_max) n -= n_max;
            const float ix = xi * c, qx = xq * c, iz = xi * z, qz = xq * z;
            if (!mix_conj) { xi = ix - qz; xq = qx + iz; }  /* :1025-1026 (the T1/C1 chain) */
            else           { xi = ix + qz; xq = qx - iz; }  /* :1029-1030 (the S1 chain)    */
        }
        if (o->prefilter) {
            memmove(pre_i + 1, pre_i, 23 * sizeof(float)); pre_i[0] = xi;
            memmove(pre_q + 1, pre_q, 23 * sizeof(float)); pre_q[0] = xq;
            memmove(prx_i + 1, prx_i, 23 * sizeof(int32_t)); prx_i[0] = (int32_t)((int64_t)xi * 256);   /* fixedpt_fromint */
            memmove(prx_q + 1, prx_q, 23 * sizeof(int32_t)); prx_q[0] = (int32_t)((int64_t)xq * 256);
            if (++since < d) continue;                      /* the filter runs on every sample; only these outputs are used */
            since = 0;
            /* d = 2 (the product's precondition for the 1.6 MS/s designs): the kept sample is the polyphase filter's
             * phase 1, so [0], [2], .. are the odd-phase samples and [1], [3], .. the even-phase ones */
            float yi = 0, yq = 0;
            if (o->prefilter == 1) {
                for (int j = 0; j < 23; j++) { yi += pre_b[j] * pre_i[j]; yq += pre_b[j] * pre_q[j]; }
            } else if (o->prefilter == 2) {
                float ei = 0, eq = 0, oi = 0, oq = 0;
                for (int j = 0; j < 12; j++) {              /* phase 0: fir[0] = b[1][] = taps 1, 3, .., 21 and a zero */
                    const float b = j < 11 ? pre_b[2 * j + 1] : 0.0f;
                    ei += b * pre_i[2 * j + 1]; eq += b * pre_q[2 * j + 1];
                }
                for (int j = 0; j < 12; j++) {              /* phase 1: fir[1] = b[0][] = taps 0, 2, .., 22 */
                    oi += pre_b[2 * j] * pre_i[2 * j]; oq += pre_b[2 * j] * pre_q[2 * j];
                }
                yi = 0; yi += ei; yi += oi;                 /* ppf.h:50, :53 */
                yq = 0; yq += eq; yq += oq;
            } else {
                int32_t ai = 0, aq = 0;
                if (o->prefilter == 3) {
                    for (int j = 0; j < 23; j++) {
                        ai += (int32_t)(((int64_t)pre_bx[j] * (int64_t)prx_i[j]) >> 8);
                        aq += (int32_t)(((int64_t)pre_bx[j] * (int64_t)prx_q[j]) >> 8);
                    }
                } else {
                    int32_t ei = 0, eq = 0, oi = 0, oq = 0;
                    for (int j = 0; j < 12; j++) {
                        const int32_t b = j < 11 ? pre_bx[2 * j + 1] : pre_bx[23];
                        ei += (int32_t)(((int64_t)b * (int64_t)prx_i[2 * j + 1]) >> 8);
                        eq += (int32_t)(((int64_t)b * (int64_t)prx_q[2 * j + 1]) >> 8);
                    }
                    for (int j = 0; j < 12; j++) {
                        oi += (int32_t)(((int64_t)pre_bx[2 * j] * (int64_t)prx_i[2 * j]) >> 8);
                        oq += (int32_t)(((int64_t)pre_bx[2 * j] * (int64_t)prx_q[2 * j]) >> 8);
                    }
                    ai = 0 + ei + oi; aq = 0 + eq + oq;
                }
                yi = (float)(ai * ((float)(1) / (float)(1 << 8)));   /* fixedpt_tofloat */
                yq = (float)(aq * ((float)(1) / (float)(1 << 8)));
            }
            si[m] = yi; sq[m] = yq;
            m++;
            continue;
        }
        const int vi = (int)xi, vq = (int)xq;               /* float -> int param of mavgi() */
        sum_i += vi - ring_i[pos]; ring_i[pos] = vi;
        sum_q += vq - ring_q[pos]; ring_q[pos] = vq;
        if (++pos >= L) pos = 0;
        if (++since < d) continue;                          /* :1350-1352 */
        since = 0;
        si[m] = (float)sum_i / (float)L;
        sq[m] = (float)sum_q / (float)L;
        m++;
    }
    free(lut_c); free(lut_s);
}

/* ------------------------------------------------------------------------- */
/* A.4 polar discriminator  rtl_wmbus.c:517-586, atan2.h:7-10                 */
/* ------------------------------------------------------------------------- */

void orc_discriminator(const float *si, const float *sq, size_t M, int accurate, float *dphi_raw)
{
    float pi_ = 0.0f, pq_ = 0.0f;       /* s_last, zero-initialised static */
    const float inv_pi = (float)M_1_PI;
    for (size_t m = 0; m < M; m++) {
        const float i = si[m], q = sq[m];
        if (accurate) {
            /* s * conjf(s_last): (a+bi)(c+di), c = i', d = -q' */
            const float c = pi_, dd = -pq_;
            const float re = i * c - q * dd;
            const float im = i * dd + q * c;
            dphi_raw[m] = orc_atan2f(im, re) * inv_pi;
        } else {
            dphi_raw[m] = pi_ * q - i * pq_;                /* :545 */
        }
        pi_ = i; pq_ = q;
    }
}

/* ------------------------------------------------------------------------- */
/* A.5 post-demod FIR   fir.h:37-72, coefficients rtl_wmbus.c:372, :384       */
/* ------------------------------------------------------------------------- */

static const float FIR_T1C1[11] = {
    -0.00456638213, -0.002571450348, 0.02689425925, 0.1141330398, 0.2264456422, 0.2793297826,
    0.2264456422, 0.1141330398, 0.02689425925, -0.002571450348, -0.00456638213 };

static const float FIR_S1[46] = {
    -0.000649081282, -0.0009491938209, -0.001361601657, -0.001910785234, -0.002570133495,
    -0.003251218426, -0.003801634695, -0.004012672882, -0.003636803575, -0.002413585945,
    -0.0001013597693, 0.003488892085, 0.008461671287, 0.01481127545, 0.02240598045,
    0.03098477999, 0.0401679839, 0.04948137286, 0.05839197924, 0.06635211627, 0.07284719662,
    0.07744230649, 0.07982251613, 0.07982251613, 0.07744230649, 0.07284719662, 0.06635211627,
    0.05839197924, 0.04948137286, 0.0401679839, 0.03098477999, 0.02240598045, 0.01481127545,
    0.008461671287, 0.003488892085, -0.0001013597693, -0.002413585945, -0.003636803575,
    -0.004012672882, -0.003801634695, -0.003251218426, -0.002570133495, -0.001910785234,
    -0.001361601657, -0.0009491938209, -0.000649081282 };

void orc_fir(const float *x, size_t M, int chain, float *y)
{
    const float *b = (chain == ORC_CHAIN_T1C1) ? FIR_T1C1 : FIR_S1;
    const int T = (chain == ORC_CHAIN_T1C1) ? 11 : 46;
    for (size_t m = 0; m < M; m++) {
        float acc = 0.0f;
        for (int t = 0; t < T; t++) {
            const float xv = ((size_t)t <= m) ? x[m - t] : 0.0f;
            acc += b[t] * xv;           /* newest sample first, then older ones */
        }
        y[m] = acc;
    }
}

/* ------------------------------------------------------------------------- */
/* A.6 DC block (-o), slicer, RSSI   rtl_wmbus.c:475-515, :1059, :1066-1067   */
/* ------------------------------------------------------------------------- */

void orc_dcblock(float *x, size_t M)
{
    const float alpha = 0.999f;                 /* :89-95 */
    float x_old = 0.0f, y_old = 0.0f;
    for (size_t m = 0; m < M; m++) {
        y_old = (1.f + alpha) / 2.f * (x[m] - x_old) + alpha * y_old;   /* :501, :511 */
        x_old = x[m];
        x[m] = y_old;
    }
}

void orc_slicer(const float *dphi, size_t M, uint8_t *bit)
{
    for (size_t m = 0; m < M; m++) bit[m] = (dphi[m] >= 0) ? 1 : 0;     /* :1059 */
}

void orc_rssi(const float *si, const float *sq, size_t M, float *rssi)
{
    float r = 0.0f;
    for (size_t m = 0; m < M; m++) {
        const float mag = sqrtf(si[m] * si[m] + sq[m] * sq[m]);         /* :1066 */
        r = 0.6789f * mag + (1.0f - 0.6789f) * r;                        /* :480 */
        rssi[m] = r;
    }
}

/* ------------------------------------------------------------------------- */
/* A.7 time2 clock recovery   iir.h:36-77, coefficients rtl_wmbus.c:338-341,  */
/* :353-356; lock FSM :1092-1111                                              */
/* ------------------------------------------------------------------------- */

static const float IIR_B_T1C1[9] = { 1, 1.999994649, 0.9999946492, 1, -1.99999482, 0.9999948196, 1, 1.703868036e-07, -1.000010531 };
static const float IIR_A_T1C1[9] = { 1, -1.387139203, 0.9921518712, 1, -1.403492665, 0.9845934971, 1, -1.430055639, 0.9923856172 };
static const float IIR_B_S1[9]   = { 1, 1.999994187, 0.9999941867, 1, -1.999994026, 0.9999940262, 1, -1.605750097e-07, -1.000011787 };
static const float IIR_A_S1[9]   = { 1, -1.92151475, 0.9918135499, 1, -1.922481015, 0.984593497, 1, -1.937432099, 0.9927241336 };
static const float IIR_GAIN = 1.874981046e-06;

void orc_clock_state(const float *dphi, size_t M, int chain, float *h, uint8_t *clk)
{
    const float *b = (chain == ORC_CHAIN_T1C1) ? IIR_B_T1C1 : IIR_B_S1;
    const float *a = (chain == ORC_CHAIN_T1C1) ? IIR_A_T1C1 : IIR_A_S1;
    for (size_t m = 0; m < M; m++) {
        float v = dphi[m] * dphi[m];                        /* :1089 */
        for (int s = 0; s < 3; s++) {
            float *hs = h + 3 * s;
            hs[0] = v - (a[3 * s + 1] * hs[1] + a[3 * s + 2] * hs[2]);
            v = b[3 * s] * hs[0] + b[3 * s + 1] * hs[1] + b[3 * s + 2] * hs[2];
            hs[2] = hs[1];
            hs[1] = hs[0];
        }
        v *= IIR_GAIN;
        clk[m] = (v >= 0) ? 1 : 0;
    }
}

void orc_clock(const float *dphi, size_t M, int chain, uint8_t *clk)
{
    float h[9] = {0};
    orc_clock_state(dphi, M, chain, h, clk);
}

void orc_time2_strobe(const uint8_t *clk, size_t M, uint8_t *strobe)
{
    int old = 0;            /* INT16_MIN: low */
    unsigned lock = 0;
    for (size_t m = 0; m < M; m++) {
        const int c = clk[m];
        strobe[m] = 0;
        if (c > old) lock = 1;
        else if (c) {
            if (lock < 2) lock++;
            else if (lock == 2) { lock++; strobe[m] = 1; }
        }
        old = c;
    }
}

/* ------------------------------------------------------------------------- */
/* CRC-16 (poly 0x3D65, init 0, complemented)  t1_c1_packet_decoder.h:463-469 */
/* ------------------------------------------------------------------------- */

static uint16_t crc_tab[256];
static uint8_t  dec3of6[64];
static uint16_t tlg_len[256];
static int tables_ready;

static void build_tables(void)
{
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) {
        uint16_t c = (uint16_t)(i << 8);
        for (int b = 0; b < 8; b++) c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x3D65) : (uint16_t)(c << 1);
        crc_tab[i] = c;
    }
    /* EN 13757-4 "3 out of 6" code words for nibbles 0..15 */
    static const uint8_t enc[16] = { 0x16, 0x0D, 0x0E, 0x0B, 0x1C, 0x19, 0x1A, 0x13,
                                     0x2C, 0x25, 0x26, 0x23, 0x34, 0x31, 0x32, 0x29 };
    memset(dec3of6, 0xFF, sizeof(dec3of6));
    for (int v = 0; v < 16; v++) dec3of6[enc[v]] = (uint8_t)v;
    /* frame format A total length from the L-field: L-field + L bytes + 2 CRC
     * bytes per block (block 1 = 10 bytes, following blocks = 16 bytes);
     * t1_c1_packet_decoder.h:68-96 */
    for (int L = 0; L < 256; L++) {
        const int blocks = 1 + (L > 9 ? (L - 9 + 15) / 16 : 0);
        tlg_len[L] = (uint16_t)(1 + L + 2 * blocks);
    }
    tables_ready = 1;
}

uint16_t orc_crc16(const uint8_t *data, size_t n)
{
    build_tables();
    uint16_t crc = 0;
    while (n--) crc = (uint16_t)(crc_tab[*data++ ^ (crc >> 8)] ^ (crc << 8));
    return (uint16_t)~crc;
}

/* block-wise CRC check, frame format A   t1_c1_packet_decoder.h:471-506 */
static int crc_ok_format_a(const uint8_t *p, size_t n)
{
    if (n < 12) return 0;
    if (orc_crc16(p, 10) != (uint16_t)((p[10] << 8) | p[11])) return 0;
    p += 12; n -= 12;
    while (n) {
        const size_t blk = (n >= 18) ? 18 : n;
        if (blk < 2) return 0;          /* cannot occur for table lengths */
        if (orc_crc16(p, blk - 2) != (uint16_t)((p[blk - 2] << 8) | p[blk - 1])) return 0;
        p += blk; n -= blk;
    }
    return 1;
}

/* frame format B   t1_c1_packet_decoder.h:508-536 */
static int crc_ok_format_b(const uint8_t *p, size_t n)
{
    if (n < 12) return 0;
    while (n) {
        const size_t blk = (n >= 128) ? 128 : n;
        if (blk < 2) return 0;          /* reference would read out of bounds here */
        if (orc_crc16(p, blk - 2) != (uint16_t)((p[blk - 2] << 8) | p[blk - 1])) return 0;
        p += blk; n -= blk;
    }
    return 1;
}

/* CRC strip, format A   t1_c1_packet_decoder.h:551-592 */
static unsigned strip_format_a(uint8_t *p, unsigned n)
{
    unsigned out = 0;
    if (p[0] > 0 && n >= 12) {
        out = 10;
        unsigned src = 12; n -= 12;
        while (n) {
            const unsigned blk = (n >= 18) ? 18 : n;
            const unsigned keep = blk - 2;
            memmove(p + out, p + src, keep);
            out += keep; src += blk; n -= blk;
        }
    }
    return out;
}

/* CRC strip, format B   t1_c1_packet_decoder.h:595-636 */
static unsigned strip_format_b(uint8_t *p, unsigned n)
{
    unsigned out = 0;
    if (p[0] >= 2 && n >= 12) {
        unsigned src = 0;
        while (n) {
            const unsigned blk = (n >= 128) ? 128 : n;
            if (blk < 2) break;
            const unsigned keep = blk - 2;
            memmove(p + out, p + src, keep);
            out += keep; src += blk; n -= blk;
            p[0] = (uint8_t)(p[0] - 2);
        }
    }
    return out;
}

/* ------------------------------------------------------------------------- */
/* Line sink                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct sink {
    char *buf; size_t cap, len, lines;
    int show_algo, real_ts;
} sink;

static void sink_put(sink *s, const char *txt, size_t n)
{
    if (s->buf && s->len + n < s->cap) memcpy(s->buf + s->len, txt, n);
    s->len += n;
}

static void make_timestamp(char *ts, size_t n, int real)
{
    if (!real) { snprintf(ts, n, "TS"); return; }
    struct timeval tv; struct tm tmv;                   /* rtl_wmbus_util.h:10-39 */
    gettimeofday(&tv, NULL);
    localtime_r(&tv.tv_sec, &tmv);
    char fmt[64];
    strftime(fmt, sizeof(fmt), "%Y-%m-%d %H:%M:%S.%%06u", &tmv);
    snprintf(ts, n, fmt, (unsigned)tv.tv_usec);
}

static void emit_line(sink *s, const char *algo, const char *mode, unsigned crc_ok, unsigned ok3of6,
                      unsigned packet_rssi, unsigned cur_rssi, const uint8_t *pkt, unsigned len)
{
    char head[160], ts[64];
    make_timestamp(ts, sizeof(ts), s->real_ts);
    uint32_t serial; memcpy(&serial, pkt + 4, 4);       /* t1_c1_packet_decoder.h:638-645 */
    int n = snprintf(head, sizeof(head), "%s%s;%u;%u;%s;%u;%u;%08X;0x", s->show_algo ? algo : "", mode,
                     crc_ok, ok3of6, ts, packet_rssi, cur_rssi, serial);
    sink_put(s, head, (size_t)n);
    static const char hexd[] = "0123456789abcdef";
    for (unsigned i = 0; i < len; i++) {
        char h[2] = { hexd[pkt[i] >> 4], hexd[pkt[i] & 15] };
        sink_put(s, h, 2);
    }
    sink_put(s, "\n", 1);
    s->lines++;
}

/* ------------------------------------------------------------------------- */
/* T1 / C1 framer   t1_c1_packet_decoder.h:136-460, :649-712                  */
/* ------------------------------------------------------------------------- */

enum { PH_IDLE = 0, PH_T1_LHI, PH_T1_LLO, PH_T1_DHI, PH_T1_DLO, PH_C1_MODE, PH_C1_L, PH_C1_DATA,
       PH_S1_L, PH_S1_DATA };

typedef struct framer {
    int phase, nbit;
    unsigned packet_rssi, err3of6, c1, bframe, l, L, mode, byte;
    uint8_t packet[292];
} framer;

static void framer_reset(framer *f) { memset(f, 0, sizeof(*f)); }

static void t1c1_finish(framer *f, unsigned rssi, const char *algo, sink *out)
{
    const unsigned crc = f->bframe ? crc_ok_format_b(f->packet, f->L) : crc_ok_format_a(f->packet, f->L);
    /* The serial (bytes 4..7) is printed before the CRC strip (:677 precedes
     * :688); the strip never moves or alters bytes 1..9, so reading it from the
     * stripped copy is equivalent. */
    uint8_t tmp[292]; memcpy(tmp, f->packet, sizeof(tmp));
    const unsigned len = f->bframe ? strip_format_b(tmp, f->L) : strip_format_a(tmp, f->L);
    emit_line(out, algo, f->c1 ? "C1" : "T1", crc, f->err3of6 ^ 1u, f->packet_rssi, rssi, tmp, len);
}

/* returns 1 when a line was completed on this bit */
static int t1c1_push(framer *f, unsigned bit, unsigned sync, unsigned rssi, const char *algo, sink *out)
{
    build_tables();
    int done = 0;
    switch (f->phase) {
    case PH_IDLE:
        if (!sync) { framer_reset(f); return 0; }           /* :272-278 */
        f->phase = PH_T1_LHI; f->nbit = 0;
        break;
    case PH_T1_LHI:
        if (f->nbit == 0) { f->byte = bit; f->packet_rssi = rssi; }   /* :292-296 */
        else f->byte = (f->byte << 1) | bit;
        if (++f->nbit == 6) {                                /* :298-306 */
            f->mode = f->byte;
            f->L = dec3of6[f->byte] == 0xFF ? 0xFFu : (unsigned)dec3of6[f->byte] << 4;
            f->err3of6 = f->c1 = f->bframe = 0;
            f->phase = PH_T1_LLO; f->nbit = 0;
        }
        break;
    case PH_T1_LLO:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 6) {                                /* :313-349 */
            f->mode = (f->mode << 6) | f->byte;
            const unsigned lo = dec3of6[f->byte];
            if (f->L == 0xFFu || lo == 0xFFu) {
                if (f->mode == 0x54Cu)      { f->bframe = 0; f->phase = PH_C1_MODE; f->nbit = 0; }
                else if (f->mode == 0x543u) { f->bframe = 1; f->phase = PH_C1_MODE; f->nbit = 0; }
                else { framer_reset(f); return 0; }
            } else {
                f->bframe = 0; f->c1 = 0;
                f->L |= lo; f->l = 0;
                f->packet[f->l++] = (uint8_t)f->L;
                f->L = tlg_len[f->L];
                f->phase = PH_T1_DHI; f->nbit = 0;
            }
        }
        break;
    case PH_T1_DHI:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 6) {                                /* :356-366 */
            const unsigned hi = dec3of6[f->byte];
            if (hi == 0xFFu) { f->err3of6 = 1; f->packet[f->l] = 0xFF; }
            else f->packet[f->l] = (uint8_t)(hi << 4);
            f->phase = PH_T1_DLO; f->nbit = 0;
        }
        break;
    case PH_T1_DLO:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 6) {                                /* :373-392 */
            const unsigned lo = dec3of6[f->byte];
            if (lo == 0xFFu) f->err3of6 = 1;
            f->packet[f->l++] |= (uint8_t)lo;
            if (f->l < f->L) { f->phase = PH_T1_DHI; f->nbit = 0; }
            else done = 1;
        }
        break;
    case PH_C1_MODE:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 4) {                                /* :399-415 */
            f->mode = (f->mode << 4) | f->byte;
            if (f->byte == 0xDu) { f->c1 = 1; f->phase = PH_C1_L; f->nbit = 0; }
            else { framer_reset(f); return 0; }
        }
        break;
    case PH_C1_L:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 8) {                                /* :422-438 */
            f->L = f->byte; f->l = 0;
            f->packet[f->l++] = (uint8_t)f->L;
            f->L = f->bframe ? 1u + f->L : tlg_len[f->L];
            f->phase = PH_C1_DATA; f->nbit = 0;
        }
        break;
    case PH_C1_DATA:
        f->byte = f->nbit ? ((f->byte << 1) | bit) : bit;
        if (++f->nbit == 8) {                                /* :445-460 */
            f->packet[f->l++] = (uint8_t)f->byte;
            if (f->l < f->L) f->nbit = 0;
            else done = 1;
        }
        break;
    default:
        framer_reset(f); return 0;
    }
    if (done) {                                              /* :659-702 */
        t1c1_finish(f, rssi, algo, out);
        framer_reset(f);
        return 1;
    }
    if (rssi < 5u) framer_reset(f);                          /* :703-710 */
    return 0;
}

/* ------------------------------------------------------------------------- */
/* S1 framer   s1_packet_decoder.h:35-282                                     */
/* ------------------------------------------------------------------------- */

static int s1_push(framer *f, unsigned bit, unsigned sync, unsigned rssi, const char *algo, sink *out)
{
    build_tables();
    int done = 0;
    switch (f->phase) {
    case PH_IDLE:
        if (!sync) { framer_reset(f); return 0; }
        f->phase = PH_S1_L; f->nbit = 0;
        break;
    case PH_S1_L:
    case PH_S1_DATA: {
        if (f->nbit == 0) {
            f->byte = bit;
            if (f->phase == PH_S1_L) f->packet_rssi = rssi;  /* :170-174 */
        } else {
            f->byte = (f->byte << 1) | bit;
            if (f->nbit & 1) {                               /* second chip of a pair: :152-168 */
                const unsigned pair = f->byte & 3u;
                if (pair == 0u || pair == 3u) { framer_reset(f); return 0; }
                const unsigned v = (pair == 1u) ? 1u : 0u;   /* 01 -> one, 10 -> zero */
                f->byte = ((f->byte >> 2) << 1) | v;
            }
        }
        if (++f->nbit == 16) {
            if (f->phase == PH_S1_L) {                       /* :176-197 */
                f->L = f->byte; f->l = 0;
                f->packet[f->l++] = (uint8_t)f->L;
                f->L = tlg_len[f->L & 0xFFu];
                f->phase = PH_S1_DATA; f->nbit = 0;
            } else {                                         /* :204-231 */
                f->packet[f->l++] = (uint8_t)f->byte;
                if (f->l < f->L) f->nbit = 0;
                else done = 1;
            }
        }
        break; }
    default:
        framer_reset(f); return 0;
    }
    if (done) {                                              /* :243-272 */
        const unsigned crc = crc_ok_format_a(f->packet, f->L);
        uint8_t tmp[292]; memcpy(tmp, f->packet, sizeof(tmp));
        const unsigned len = strip_format_a(tmp, f->L);
        /* serial from unstripped bytes 4..7; they never move during the strip */
        emit_line(out, algo, "S1", crc, 1u, f->packet_rssi, rssi, tmp, len);
        framer_reset(f);
        return 1;
    }
    if (rssi < 5u) framer_reset(f);                          /* :273-281 */
    return 0;
}

size_t orc_frame_t1c1(const uint8_t *bits, const uint8_t *rssi, size_t n,
                      const char *algo_prefix, char *out, size_t outcap, int *got_line)
{
    framer f; framer_reset(&f);
    sink s = { out, outcap, 0, 0, 1, 0 };
    *got_line = 0;
    size_t i = 0;
    for (; i < n; i++) {
        const int r = t1c1_push(&f, bits[i] & 1u, i == 0, rssi[i], algo_prefix, &s);
        if (r) { *got_line = 1; i++; break; }
        if (f.phase == PH_IDLE) { i++; break; }
    }
    if (out && s.len < outcap) out[s.len] = 0;
    return i;
}

size_t orc_frame_s1(const uint8_t *bits, const uint8_t *rssi, size_t n,
                    const char *algo_prefix, char *out, size_t outcap, int *got_line)
{
    framer f; framer_reset(&f);
    sink s = { out, outcap, 0, 0, 1, 0 };
    *got_line = 0;
    size_t i = 0;
    for (; i < n; i++) {
        const int r = s1_push(&f, bits[i] & 1u, i == 0, rssi[i], algo_prefix, &s);
        if (r) { *got_line = 1; i++; break; }
        if (f.phase == PH_IDLE) { i++; break; }
    }
    if (out && s.len < outcap) out[s.len] = 0;
    return i;
}

/* ------------------------------------------------------------------------- */
/* A.8 bit sync state machines                                                */
/* ------------------------------------------------------------------------- */

/* time2: shift register + access code compare   rtl_wmbus.c:806-852, :97-103 */
typedef struct t2_state { uint32_t shreg; } t2_state;

static inline unsigned t2_step(t2_state *t, unsigned bit, int chain)
{
    t->shreg = (t->shreg << 1) | bit;
    if (chain == ORC_CHAIN_T1C1) return (t->shreg & 0xFFFFu) == 0x543Du;
    return (t->shreg & 0xFFFFFFu) == 0x547696u;
}

/* run length, T1/C1   rtl_wmbus.c:705-803 */
typedef struct rl_t1_state {
    int run_length, bit_length, cum_err;
    unsigned state;
    uint32_t raw, shreg;
} rl_t1_state;

static void rl_t1_reset(rl_t1_state *r)
{
    r->run_length = 0; r->bit_length = 8 * 256; r->cum_err = 0;
    r->state = 0; r->raw = 0; r->shreg = 0;
}

/* run length, S1   rtl_wmbus.c:617-702 */
typedef struct rl_s1_state {
    int run_length, spb[2];
    unsigned state;
    uint32_t raw, shreg;
} rl_s1_state;

static void rl_s1_reset(rl_s1_state *r)
{
    r->run_length = 0; r->state = 0; r->raw = 0; r->shreg = 0;
    r->spb[0] = 24; r->spb[1] = 24;
}

typedef void (*rl_emit_fn)(void *ctx, size_t m, unsigned bit, unsigned sync, unsigned rssi);
typedef void (*rl_reset_fn)(void *ctx);

static void rl_t1_step(rl_t1_state *r, size_t m, unsigned raw_bit, unsigned rssi,
                       rl_emit_fn emit, rl_reset_fn on_reset, void *ctx)
{
    r->raw = (r->raw << 1) | raw_bit;
    const unsigned st = (__builtin_popcount(r->raw & 0x3Fu) >= 3) ? 1u : 0u;   /* :126-144, :733 */
    if (r->state == st) { r->run_length++; return; }
    if (r->run_length < 5) {                                                     /* :742-748 */
        rl_t1_reset(r); on_reset(ctx); r->state = st; r->run_length = 1; return;
    }
    r->run_length *= 256;
    const int half = r->bit_length / 2;
    if (r->run_length <= half) {                                                 /* :756-762 */
        rl_t1_reset(r); on_reset(ctx); r->state = st; r->run_length = 1; return;
    }
    int n = 0;
    for (; r->run_length > half; n++) {                                          /* :765-779 */
        r->run_length -= r->bit_length;
        r->shreg = (r->shreg << 1) | r->state;
        emit(ctx, m, r->state, (r->shreg & 0xFFFFu) == 0x543Du, rssi);
    }
    r->cum_err += r->run_length;                                                 /* :792 */
    r->bit_length += (r->run_length + r->cum_err / 16) / (32 * n);               /* :796 */
    r->state = st; r->run_length = 1;
}

static void rl_s1_step(rl_s1_state *r, size_t m, unsigned raw_bit, unsigned rssi,
                       rl_emit_fn emit, rl_reset_fn on_reset, void *ctx)
{
    r->raw = (r->raw << 1) | raw_bit;
    const unsigned st = (0xFEEAu >> (r->raw & 0xFu)) & 1u;                      /* :149-154 */
    if (r->state == st) { r->run_length++; return; }
    const int spb = (r->spb[0] + r->spb[1]) / 2;                                 /* :655 */
    if (spb <= 12 || spb >= 36) {                                                /* :659-665 */
        rl_s1_reset(r); on_reset(ctx); r->state = st; r->run_length = 1; return;
    }
    const int half = spb / 2;
    const int run = r->run_length;
    if (run <= half) {                                                           /* :671-677 */
        rl_s1_reset(r); on_reset(ctx); r->state = st; r->run_length = 1; return;
    }
    int n = 0;
    for (; r->run_length > half; n++) {                                          /* :680-694 */
        r->run_length -= spb;
        r->shreg = (r->shreg << 1) | r->state;
        emit(ctx, m, r->state, (r->shreg & 0xFFFFFFu) == 0x547696u, rssi);
    }
    r->spb[r->state] = run / n;                                                  /* :698 */
    r->state = st; r->run_length = 1;
}

/* event collectors */
typedef struct ev_ctx { orc_event *ev; size_t cap, n; uint8_t pending_reset; } ev_ctx;

static void ev_emit(void *c, size_t m, unsigned bit, unsigned sync, unsigned rssi)
{
    ev_ctx *e = c;
    if (e->n < e->cap) {
        e->ev[e->n].m = m; e->ev[e->n].bit = (uint8_t)bit; e->ev[e->n].sync = (uint8_t)sync;
        e->ev[e->n].reset = e->pending_reset; e->ev[e->n].rssi = (uint8_t)rssi;
    }
    e->pending_reset = 0;
    e->n++;
}
static void ev_reset(void *c) { ((ev_ctx *)c)->pending_reset = 1; }

size_t orc_time2_events(const uint8_t *bit, const uint8_t *strobe, const float *rssi,
                        size_t M, int chain, orc_event *ev, size_t cap)
{
    ev_ctx e = { ev, cap, 0, 0 };
    t2_state t = { 0 };
    for (size_t m = 0; m < M; m++) {
        if (!strobe[m]) continue;
        const unsigned sync = t2_step(&t, bit[m], chain);
        ev_emit(&e, m, bit[m], sync, (unsigned)rssi[m]);
    }
    return e.n;
}

size_t orc_runlength_events(const uint8_t *bit, const float *rssi, size_t M, int chain,
                            orc_event *ev, size_t cap)
{
    ev_ctx e = { ev, cap, 0, 0 };
    if (chain == ORC_CHAIN_T1C1) {
        rl_t1_state r; rl_t1_reset(&r);
        for (size_t m = 0; m < M; m++) rl_t1_step(&r, m, bit[m], (unsigned)rssi[m], ev_emit, ev_reset, &e);
    } else {
        rl_s1_state r; rl_s1_reset(&r);
        for (size_t m = 0; m < M; m++) rl_s1_step(&r, m, bit[m], (unsigned)rssi[m], ev_emit, ev_reset, &e);
    }
    return e.n;
}

/* ------------------------------------------------------------------------- */
/* Whole pipeline in the reference's output order                             */
/* (rtl_wmbus.c:1354-1355 -> :1074, :1106, :1166, :1198)                      */
/* ------------------------------------------------------------------------- */

typedef struct chain_run {
    int chain;
    float *dphi, *rssi;
    uint8_t *bit, *strobe;
    framer f_rla, f_t2;
    t2_state t2;
    rl_t1_state rl_t1;
    rl_s1_state rl_s1;
    sink *out;
} chain_run;

static void run_rla_emit(void *c, size_t m, unsigned bit, unsigned sync, unsigned rssi)
{
    (void)m;
    chain_run *cr = c;
    if (cr->chain == ORC_CHAIN_T1C1) t1c1_push(&cr->f_rla, bit, sync, rssi, "rla;", cr->out);
    else                             s1_push(&cr->f_rla, bit, sync, rssi, "rla;", cr->out);
}
static void run_rla_reset(void *c) { framer_reset(&((chain_run *)c)->f_rla); }

static void chain_prepare(chain_run *cr, int chain, const uint8_t *cu8, size_t n_iq, size_t M,
                          const orc_opts *o, sink *out)
{
    memset(cr, 0, sizeof(*cr));
    cr->chain = chain; cr->out = out;
    float *si = malloc(M * sizeof(float) + 4), *sq = malloc(M * sizeof(float) + 4);
    float *raw = malloc(M * sizeof(float) + 4);
    cr->dphi = malloc(M * sizeof(float) + 4);
    cr->rssi = malloc(M * sizeof(float) + 4);
    cr->bit = malloc(M + 4); cr->strobe = malloc(M + 4);
    uint8_t *clk = malloc(M + 4);
    orc_frontend(cu8, n_iq, o, chain, si, sq);
    orc_discriminator(si, sq, M, o->accurate_atan, raw);
    orc_fir(raw, M, chain, cr->dphi);
    if (o->remove_dc) orc_dcblock(cr->dphi, M);
    orc_slicer(cr->dphi, M, cr->bit);
    orc_rssi(si, sq, M, cr->rssi);
    orc_clock(cr->dphi, M, chain, clk);
    orc_time2_strobe(clk, M, cr->strobe);
    free(si); free(sq); free(raw); free(clk);
    rl_t1_reset(&cr->rl_t1); rl_s1_reset(&cr->rl_s1);
}

static void chain_free(chain_run *cr)
{
    free(cr->dphi); free(cr->rssi); free(cr->bit); free(cr->strobe);
}

static void chain_step(chain_run *cr, size_t m, const orc_opts *o)
{
    const unsigned bit = cr->bit[m];
    const unsigned rssi = (unsigned)cr->rssi[m];            /* float -> unsigned parameter */
    if (o->rla_enabled) {
        if (cr->chain == ORC_CHAIN_T1C1) rl_t1_step(&cr->rl_t1, m, bit, rssi, run_rla_emit, run_rla_reset, cr);
        else                             rl_s1_step(&cr->rl_s1, m, bit, rssi, run_rla_emit, run_rla_reset, cr);
    }
    if (o->t2_enabled && cr->strobe[m]) {
        const unsigned sync = t2_step(&cr->t2, bit, cr->chain);
        if (cr->chain == ORC_CHAIN_T1C1) t1c1_push(&cr->f_t2, bit, sync, rssi, "t2a;", cr->out);
        else                             s1_push(&cr->f_t2, bit, sync, rssi, "t2a;", cr->out);
    }
}

size_t orc_run(const uint8_t *cu8, size_t nbytes, const orc_opts *o,
               char *out, size_t outcap, size_t *n_lines)
{
    build_tables();
    nbytes -= nbytes % 4096;                                /* rtl_wmbus.c:1301-1308 */
    const size_t n_iq = nbytes / 2;
    const size_t M = orc_num_decimated(n_iq, o->decimation);
    sink s = { out, outcap, 0, 0, o->show_algorithm, o->real_timestamp };
    chain_run t1, s1;
    if (o->t1c1_enabled) chain_prepare(&t1, ORC_CHAIN_T1C1, cu8, n_iq, M, o, &s);
    if (o->s1_enabled)   chain_prepare(&s1, ORC_CHAIN_S1, cu8, n_iq, M, o, &s);
    for (size_t m = 0; m < M; m++) {
        if (o->t1c1_enabled) chain_step(&t1, m, o);
        if (o->s1_enabled)   chain_step(&s1, m, o);
    }
    if (o->t1c1_enabled) chain_free(&t1);
    if (o->s1_enabled)   chain_free(&s1);
    if (out && s.len < outcap) out[s.len] = 0;
    if (n_lines) *n_lines = s.lines;
    return s.len;
}
