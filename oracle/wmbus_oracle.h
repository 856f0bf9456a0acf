/*
 * wmbus_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C, array-at-a-time restatement of the per-sample DSP hot path of
 * xaelsouth/rtl-wmbus (reference @ b6a7705).  It exists so that the CUDA
 * product path in rtl-wmbus_b200/ can be checked stage by stage and line by
 * line.  Nothing in the product (rtl-wmbus_b200/, include/) may include, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Parity of THIS oracle is pinned against the unmodified reference compiled
 * from /root/reference (oracle/_ref/rtl_wmbus, recipe: oracle/Makefile):
 * tests/test_oracle_vs_ref.py and the committed goldens in tests/golden/.
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef WMBUS_ORACLE_H
#define WMBUS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CHAIN_T1C1 0
#define ORC_CHAIN_S1   1
#define ORC_ALGO_RLA   0
#define ORC_ALGO_T2A   1

/* Command-line options of the reference (rtl_wmbus.c:855-866, :892-967). */
typedef struct orc_opts {
    uint32_t decimation;      /* -d N   (default 2)                      */
    uint8_t  accurate_atan;   /* !-a    (default 1)                      */
    uint8_t  remove_dc;       /* -o     (default 0)                      */
    uint8_t  rla_enabled;     /* -r 0 clears (default 1)                 */
    uint8_t  t2_enabled;      /* -t 0 clears (default 1)                 */
    uint8_t  t1c1_enabled;    /* -p T clears (default 1)                 */
    uint8_t  s1_enabled;      /* -p S clears (default 1)                 */
    uint8_t  simultaneous;    /* -s     (default 0); 2: carriers below   */
    uint8_t  show_algorithm;  /* -v     (default 0)                      */
    uint8_t  real_timestamp;  /* 0: print the literal TS in the TIMESTAMP column */
    /* simultaneous == 2: the -s mixer (rtl_wmbus.c:997-1031) with its two constants as parameters -- table entries per
     * sample (the reference: 13 = 325 kHz / 25 kHz) and entry vs conjugate (the sign) -- per chain.  Pinned by the
     * reference only at {+13, -13}; other carriers are this restatement's own generalisation. */
    int32_t  carrier_25khz[2];
    /* 1..4: one of the reference's dormant pre-decimation low-passes (rtl_wmbus.c:197-333, never called upstream) in
     * place of the moving averages in front of the decimation -- 1 lp_fir_ (23-tap float FIR), 2 lp_ppf_
     * (polyphase, ppf.h), 3 lp_firfp_, 4 lp_ppffp_ (their 24.8 fixed-point twins, fixedptc.h).  Pinned against
     * those functions themselves through ref_stages.c; there is no reference output to compare lines with. */
    uint32_t prefilter;
} orc_opts;

void   orc_default_opts(orc_opts *o);

/* fdlibm single-precision atan2 (what glibc 2.39 libm.so.6 atan2f executes;
 * call site atan2.h:9).  Bit-exact against libm: tests/test_oracle_atan2.py */
float  orc_atan2f(float y, float x);

/* number of decimated samples produced from n_iq input IQ samples
 * (rtl_wmbus.c:1350-1352) */
size_t orc_num_decimated(size_t n_iq, uint32_t decimation);

/* Stage functions; M = number of decimated samples.  All arrays caller-owned. */
void   orc_frontend(const uint8_t *cu8, size_t n_iq, const orc_opts *o, int chain,
                    float *si, float *sq);                         /* A.1-A.3 */
void   orc_discriminator(const float *si, const float *sq, size_t M, int accurate,
                         float *dphi_raw);                         /* A.4 */
void   orc_fir(const float *x, size_t M, int chain, float *y);     /* A.5 */
void   orc_dcblock(float *x, size_t M);                            /* A.6 (-o), in place */
void   orc_slicer(const float *dphi, size_t M, uint8_t *bit);      /* A.6 */
void   orc_rssi(const float *si, const float *sq, size_t M, float *rssi); /* A.6 */
void   orc_clock(const float *dphi, size_t M, int chain, uint8_t *clk);   /* A.7: sign of IIR */
void   orc_clock_state(const float *dphi, size_t M, int chain, float *hist9_inout, uint8_t *clk);
void   orc_time2_strobe(const uint8_t *clk, size_t M, uint8_t *strobe);   /* A.7 lock FSM */

/* Bit events produced by the two bit-sync algorithms (before framing). */
typedef struct orc_event {
    uint64_t m;        /* decimated sample index at which the bit is delivered */
    uint8_t  bit;      /* data bit */
    uint8_t  sync;     /* access code matched on this bit (PACKET_PREAMBLE_DETECTED) */
    uint8_t  reset;    /* run-length FSM was reset since the previous event (rla only) */
    uint8_t  rssi;     /* (unsigned)rssi at sample m */
} orc_event;

size_t orc_time2_events(const uint8_t *bit, const uint8_t *strobe, const float *rssi,
                        size_t M, int chain, orc_event *ev, size_t cap);
size_t orc_runlength_events(const uint8_t *bit, const float *rssi, size_t M, int chain,
                            orc_event *ev, size_t cap);

/* Whole pipeline: cu8 bytes in (only whole 4096-byte blocks are consumed,
 * rtl_wmbus.c:1301-1308), text lines out, in the reference's order.
 * Returns the number of bytes written to out (NUL-terminated if room);
 * *n_lines receives the number of lines.  */
size_t orc_run(const uint8_t *cu8, size_t nbytes, const orc_opts *o,
               char *out, size_t outcap, size_t *n_lines);

/* Framers, exposed for direct tests: feed a bit sequence (first element is the
 * bit carrying the sync flag).  Returns the number of bits consumed until the
 * decoder fell back to idle (abort) or finished a line; writes the line (if
 * any) to out. */
size_t orc_frame_t1c1(const uint8_t *bits, const uint8_t *rssi, size_t n,
                      const char *algo_prefix, char *out, size_t outcap, int *got_line);
size_t orc_frame_s1(const uint8_t *bits, const uint8_t *rssi, size_t n,
                    const char *algo_prefix, char *out, size_t outcap, int *got_line);

uint16_t orc_crc16(const uint8_t *data, size_t n);

#ifdef __cplusplus
}
#endif
#endif
