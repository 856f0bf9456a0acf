/*
 * wmbus_b200.h -- C ABI of libwmbus_b200.so: the B200-native replacement for the
 * per-sample DSP hot path of rtl-wmbus (cu8 IQ -> box filter/decimate -> FM
 * discriminator -> FIR -> slicer/RSSI -> time2 + run-length bit sync -> access-code
 * correlation -> T1/C1/S1 framing -> datagram lines).
 *
 * The reference (xaelsouth/rtl-wmbus @ b6a7705) exposes no library interface: its
 * only boundary is the process contract of main() (rtl_wmbus.c:1217-1372: argv,
 * stdin cu8 in 4096-byte items, one stdout line per telegram).  The entry points
 * below are the seams a maintainer would cut when moving that loop onto a GPU; each
 * one cites the reference code it replaces.  Plain C types only -- no CUDA, torch or
 * C++ types cross this boundary.
 *
 * Threading: one caller per context (the reference is single-threaded and not
 * re-entrant: all filter state is function-local statics, e.g. rtl_wmbus.c:168-175).
 * Several contexts may live in one process (one per GPU / per capture).
 *
 * There is NO CPU fallback: wmb_create() fails with WMB_E_NODEVICE when no CUDA
 * device is usable.
 */
#ifndef WMBUS_B200_H
#define WMBUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WMB_ABI_VERSION 2

/* error codes (negative) */
#define WMB_OK             0
#define WMB_E_INVAL       -1   /* bad argument / unsupported option value          */
#define WMB_E_NODEVICE    -2   /* no usable CUDA device (there is no CPU fallback) */
#define WMB_E_CUDA        -3   /* CUDA runtime error; see wmb_last_error()         */
#define WMB_E_NOMEM       -4
#define WMB_E_OVERFLOW    -5   /* internal event/frame buffer overflow (pathological input) */
#define WMB_E_STATE       -6   /* call order violated                               */

#define WMB_CHAIN_T1C1 0
#define WMB_CHAIN_S1   1
#define WMB_ALGO_RLA   0       /* run-length bit sync   (rtl_wmbus.c:617-803) */
#define WMB_ALGO_T2A   1       /* "time2" clock recovery (rtl_wmbus.c:806-852, :1080-1115) */

/* Options == the reference's command-line switches (rtl_wmbus.c:855-866, getopt
 * string "ofad:p:r:vVst:" at :896). */
typedef struct wmb_opts {
    uint32_t decimation;      /* -d N : input rate = N x 800 kS/s; default 2 (:857); N <= 25 (20 MS/s) */
    uint8_t  accurate_atan;   /* 0 with -a (cross-product discriminator, :536-551)         */
    uint8_t  remove_dc;       /* -o  (:497-515)                                            */
    uint8_t  rla_enabled;     /* 0 with -r 0                                               */
    uint8_t  t2_enabled;      /* 0 with -t 0                                               */
    uint8_t  t1c1_enabled;    /* 0 with -p T                                               */
    uint8_t  s1_enabled;      /* 0 with -p S                                               */
    uint8_t  simultaneous;    /* -s : +-325 kHz translation (:974-1031); 2: the same mixer with the carriers of
                                 carrier_25khz[] below (not a switch of the reference)       */
    uint8_t  show_algorithm;  /* -v : prefix lines with rla; / t2a;                        */
    /* tuning knobs (0 = library default); they never change results, only how the
     * work is cut up on the device */
    uint32_t chunk_samples;   /* decimated samples per bit-sync lane                       */
    uint32_t warmup_samples;  /* speculative warm-up before each lane                      */
    uint32_t max_batch_mib;   /* largest batch handed to the device at once (default 256)  */
    uint32_t manual_frames;   /* 1: wmb_push only gathers candidates; the caller drives the
                                 framers with wmb_poll + wmb_decode_frames.  0 (default):
                                 wmb_push feeds the context's framers itself.            */
    uint32_t reserved[2];     /* test knobs, 0 in production: [0] = 1 forces the monolithic run-length lanes for
                                 T1/C1; [1] bit 0 keeps the clock-sign words for wmb_debug_copy_bits(.., 2, ..),
                                 [1] bit 1 sizes the run-length streams' event rings by the large-batch rule (a
                                 quarter event per sample) also for small batches, [1] >> 8 (if not 0) is the size of
                                 the per-batch candidate tables: both to reach the overflow paths
                                 (wmb_stats.overflow_batches) with a small capture                                 */
    /* simultaneous == 2 (many carriers per capture): offset of the carrier the T1/C1 chain [0] and the
     * S1 chain [1] listen to from the capture's centre frequency, in units of 25 kHz (the grid of the reference's
     * table, rtl_wmbus.c:974-993), |offset| <= fs / 2.  The reference's -s is {+13, -13}.  A capture with more than two
     * carriers is decoded by several contexts over the same input (shard.decode_carriers in the Python mirror). */
    int32_t  carrier_25khz[2];
    /* SURVEY 8f N4 -- one of the reference's dormant pre-decimation low-passes (rtl_wmbus.c:197-333, never called
     * upstream) takes the place of the moving averages in front of the decimation:
     *   1  lp_fir_butter_1600kHz_160kHz_200kHz_{t1_c1,s1}: 23-tap float FIR through firf() (fir.h:49-72)
     *   2  lp_ppf_butter_1600kHz_160kHz_200kHz: the same taps as a two-phase polyphase filter, ppf() (ppf.h:44-58)
     *   3  lp_firfp_butter_1600kHz_160kHz_200kHz: 1 in 24.8 fixed point, firfp() (fir.h:106-130, fixedptc.h)
     *   4  lp_ppffp_butter_1600kHz_160kHz_200kHz: 2 in 24.8 fixed point, ppffp() (ppf.h:69-83)
     * 1.6 MS/s only (decimation 2).  Not a switch of the reference: there are no reference lines to compare with; the
     * stages are checked against those functions themselves (oracle/ref_stages.c). */
    uint32_t prefilter;
} wmb_opts;

typedef struct wmb_ctx wmb_ctx;

/* One bit delivered to a framer, as produced on the device.  Bit layout of `w`:
 *   [31:9] sample offset (decimated samples) relative to wmb_frame.sync_sample
 *   [8:1]  (unsigned)rssi at that sample  (the `unsigned rssi` argument of
 *          t1_c1_packet_decoder(), t1_c1_packet_decoder.h:649)
 *   [0]    data bit                                                                */
typedef uint32_t wmb_bit;
#define WMB_BIT_DATA(w)   ((w) & 1u)
#define WMB_BIT_RSSI(w)   (((w) >> 1) & 0xFFu)
#define WMB_BIT_OFFSET(w) ((w) >> 9)

/* A framing candidate: the bit on which an access code matched (bits[0], the bit the
 * reference flags with PACKET_PREAMBLE_DETECTED, rtl_wmbus.c:773-776 / :822-825)
 * followed by the bits the same bit-sync instance produced after it.  For the
 * run-length algorithm the list stops where that algorithm reset itself (which also
 * resets its decoder, rtl_wmbus.c:717-726).                                         */
typedef struct wmb_frame {
    uint64_t sync_sample;     /* global decimated-sample index of bits[0]                 */
    uint64_t ordinal;         /* index of bits[0] in its (chain, algo) bit stream         */
    uint8_t  chain;           /* WMB_CHAIN_*                                              */
    uint8_t  algo;            /* WMB_ALGO_*                                               */
    uint8_t  truncated;       /* 1: the list ends before the header's length was reached    */
    uint8_t  reserved;        /* 1: partial -- the stream has not produced the remaining bits
                                 yet; the candidate is delivered again by a later poll     */
    uint32_t nbits;           /* number of entries in bits[] (>= 1)                       */
    const wmb_bit *bits;      /* owned by the context until the next wmb_push/wmb_poll    */
} wmb_frame;

/* ---- context ------------------------------------------------------------------ */

void        wmb_default_opts(wmb_opts *o);
int         wmb_abi_version(void);
const char *wmb_last_error(void);
const char *wmb_version_string(void);

/* Replaces the per-process set-up of main() (rtl_wmbus.c:1249-1296: algorithm
 * structs reset, chain/discriminator selection, mixer look-up tables). */
int  wmb_create(const wmb_opts *o, int cuda_device, wmb_ctx **out);
void wmb_destroy(wmb_ctx *c);
/* Start over with a new capture (same as destroy + create, without re-allocating). */
int  wmb_reset(wmb_ctx *c);

/* Page-locked host memory for input buffers (the reference reads stdin into a 4096-byte
 * stack array, rtl_wmbus.c:1249; a GPU pipeline wants to DMA straight out of the read
 * buffer).  wmb_push() accepts any host pointer; pinned ones are copied asynchronously. */
void *wmb_host_alloc(size_t nbytes);
void  wmb_host_free(void *p);

/* ---- sample path (the hot loop, rtl_wmbus.c:1298-1357) ------------------------ */

/* Feed cu8 bytes from HOST memory.  Any length; the library consumes whole
 * 4096-byte items like the reference's fread (rtl_wmbus.c:1301) and keeps a
 * remainder for the next call.  Work is queued asynchronously (pinned staging ring,
 * H2D copy and kernels on the context's streams). */
int  wmb_push(wmb_ctx *c, const uint8_t *cu8, size_t nbytes);

/* Same, for bytes already resident in DEVICE memory of the context's GPU
 * (nbytes % 4096 == 0).  Used by the benchmark's device-resident leg. */
int  wmb_push_device(wmb_ctx *c, const void *dev_cu8, size_t nbytes);

/* Wait for queued work and hand out framing candidates in stream order per
 * (chain, algo).  flush != 0 declares end of input (the reference's EOF): pending
 * candidates are delivered truncated.  Returns the number of frames in *n. */
int  wmb_poll(wmb_ctx *c, wmb_frame *out, size_t cap, size_t *n, int flush);

/* ---- host-side framers (t1_c1_packet_decoder.h, s1_packet_decoder.h) ---------- */

/* Feed frames (as returned by wmb_poll, any interleaving of the four streams) to the
 * context's framers: 3-out-of-6 / NRZ / Manchester decode, L-field, RSSI abort,
 * block CRCs, CRC strip, line formatting -- the work of t1_c1_packet_decoder()
 * (t1_c1_packet_decoder.h:649-712) and s1_packet_decoder() (s1_packet_decoder.h:233-282).
 * Completed lines are appended to the context's line queue in the reference's output
 * order (finish sample, then T1/C1-rla, T1/C1-t2a, S1-rla, S1-t2a; rtl_wmbus.c:1074,
 * :1106, :1166, :1198). */
int  wmb_decode_frames(wmb_ctx *c, const wmb_frame *frames, size_t n);

/* Copy queued lines ('\n'-terminated, concatenated) into buf; returns the number of
 * bytes written (0 when none); *n_lines (optional) receives the line count.  Lines
 * that do not fit stay queued.  timestamp_mode: 0 = wall clock at delivery
 * (rtl_wmbus_util.h:10-39 format), 1 = the literal TS, 2 = the stream position at which the
 * reference would print the line, "@<decimated sample of the telegram's last bit>.<chain*2 + (algo == t2a)>":
 * sorting by it merges lines of several contexts (time-chunk sharding) into the reference's print order. */
size_t wmb_take_lines(wmb_ctx *c, char *buf, size_t cap, size_t *n_lines, int timestamp_mode);

/* Convenience for offline captures: push + flush + decode + take_lines in one call.
 * `flush` as in wmb_poll.  Returns bytes written to out or a negative error. */
long wmb_process(wmb_ctx *c, const uint8_t *cu8, size_t nbytes, int flush,
                 char *out, size_t outcap, size_t *n_lines, int timestamp_mode);
long wmb_process_device(wmb_ctx *c, const void *dev_cu8, size_t nbytes, int flush,
                        char *out, size_t outcap, size_t *n_lines, int timestamp_mode);

/* ---- introspection for tests / benchmark -------------------------------------- */

typedef struct wmb_stats {
    uint64_t input_samples;       /* IQ samples consumed                              */
    uint64_t decimated_samples;
    uint64_t batches;
    uint64_t kernel_launches;     /* launches of this library's own kernels           */
    uint64_t lanes_run;           /* bit-sync lanes executed (incl. re-runs)          */
    uint64_t lanes_rerun;         /* lanes whose speculative start state was refuted  */
    uint64_t candidates[2][2];    /* access-code matches  [chain][algo]               */
    uint64_t lines[2][2];         /* datagram lines       [chain][algo]               */
    uint64_t lines_crc_ok[2][2];
    uint64_t h2d_bytes, d2h_bytes;
    double   demod_kernel_ms;     /* CUDA-event time of the demod kernels of the last push (sum over its batches) */
    double   bitsync_kernel_ms;   /* ... of the bit-sync kernels (clock-recovery lanes + bit streams; they overlap the
                                     next batch's demod kernel, so the two sums can exceed the wall time)      */
    double   batch_device_ms;     /* first demod kernel -> last bit-sync kernel of the last push (device clock) */
    uint64_t rl_fallbacks;        /* T1/C1 batches redone with the monolithic run-length lanes */
    double   host_batch_ms;       /* cumulative wall time in the enqueue+verify part of batches */
    double   host_gather_ms;      /* cumulative wall time gathering candidate frames          */
    double   host_decode_ms;      /* cumulative wall time in the host framers                 */
    uint64_t overflow_batches;    /* batches whose bits or candidates did not fit a device table (a run-length lane's
                                     event buffer, a stream's event ring, frame words, datagram pool, access-code
                                     matches, pending candidates) and may have lost lines; the stream goes on.  Takes an input no receiver produces: sized for one access-code
                                     match per 256 decimated samples, sustained over a whole batch                 */
} wmb_stats;

int wmb_get_stats(wmb_ctx *c, wmb_stats *s);

/* Stage taps for parity tests: after a wmb_push* + wmb_poll, copy the per-decimated-
 * sample outputs of the demod kernel for the LAST batch to host arrays
 * (dphi: post-FIR, pre-DC-block discriminator output; rssi: (unsigned)rssi).
 * Returns the number of samples copied or a negative error. */
long wmb_debug_copy_stage(wmb_ctx *c, int chain, float *dphi, uint8_t *rssi, size_t cap);

/* Bit-sync stage taps of the LAST batch, packed: bit i of word w = decimated sample 32 w + i of the batch.
 *   which 0: data bits   -- the slicer's output, behind the DC block with -o  (rtl_wmbus.c:1059, bits.bin tap :1060-1061)
 *   which 1: time2 strobes -- the samples at which the clock lock delivers a bit (rtl_wmbus.c:1092-1111)
 *   which 2: clock signs -- sign of the clock-recovery band-pass (rtl_wmbus.c:1089-1090, clock.bin tap); kept only by a
 *            context created with opts.reserved[1] = 1
 * Returns the number of words copied or a negative error. */
long wmb_debug_copy_bits(wmb_ctx *c, int chain, int which, uint32_t *words, size_t cap_words);

/* The bit events the LAST batch appended to one (chain, algo) bit stream, i.e. the calls of
 * t1_c1_packet_decoder() / s1_packet_decoder() the reference makes (rtl_wmbus.c:773-781, :822-830, rawbits.bin tap
 * :1107-1108): [63:24] decimated sample (40 bits), [23:16] (unsigned)rssi, [2] run-length reset since the previous
 * event, [1] access code matched on this bit, [0] the bit.  Returns the number of events copied. */
long wmb_debug_copy_events(wmb_ctx *c, int chain, int algo, uint64_t *ev, size_t cap);

/* Test hook: the device's exact-arithmetic building blocks (csrc/wmb_exact.cuh) on caller-made operands, host arrays of
 * n floats.  mode 0: atan2f(y, x) as the discriminator uses it (operands zero or in [2^-8, 2^23))   1: general
 * atan2f(y, x) (glibc 2.39 fdlibm restated)   2: IEEE division y / x for such operands   3: IEEE sqrt(y) for y zero or an
 * integer below 2^23   4: the polar discriminator of (I, Q) = (y[i], x[i]) against (y[i-1], x[i-1]) (atan2.h:7-10) */
int wmb_debug_arith(wmb_ctx *c, int mode, const float *y, const float *x, float *out, size_t n);

/* ---- time-chunk sharding of one capture (several contexts / GPUs on one stream) ----
 * The reference has no counterpart: it is one sequential loop (rtl_wmbus.c:1298-1357).  A worker that
 * owns the decimated samples [lo, hi) of a capture (1) seeks its context to a position a warm-up halo
 * before lo, (2) pushes the halo, takes wmb_boundary_state() and compares it with the state its left
 * neighbour took after pushing up to the same sample -- equal bytes mean that every recurrence, shift
 * register and telegram in flight is bit-identical from there on, so the chunk decodes exactly as in
 * the sequential run; a mismatch means the halo was too short (retry with a longer one; position 0 is
 * exact by definition), (3) pushes its chunk and a right halo of one maximum telegram -- and goes on pushing
 * while wmb_pending_before(hi) says that a telegram matched in the chunk is still in flight -- and keeps
 * only the lines whose access-code match lies in [lo, hi). */

/* wmb_reset() plus: the next byte pushed is IQ sample `first_iq_sample` of the capture (mixer and
 * decimation phases, sample indices).  Must be a multiple of 2048 * decimation. */
int wmb_seek(wmb_ctx *c, uint64_t first_iq_sample);

/* Only telegrams whose access-code match falls on a decimated sample in [sync_lo, sync_hi) produce
 * lines (default: all).  The others are still decoded: they keep the decoders busy as in the reference. */
int wmb_set_line_window(wmb_ctx *c, uint64_t sync_lo, uint64_t sync_hi);

/* Number of telegrams in flight -- access-code matches whose decoder is still waiting for bits -- whose match lies on a
 * decimated sample below sync_hi.  A worker pushes its right halo until this is 0 for its chunk's end: the halo of "one
 * maximum telegram" is a statement about samples with edges in them, and a telegram that runs into a gap in the input
 * (dead air: the run-length tracker emits nothing until the next edge, then all the bits at once) ends arbitrarily
 * late.  Gathers what is enqueued first, like wmb_boundary_state().  Negative: error. */
long wmb_pending_before(wmb_ctx *c, uint64_t sync_hi);

/* Everything that couples the samples pushed so far to the output still to come: the carried filter,
 * clock and run-length states, the shift registers, and the bit events of every telegram in flight
 * (absolute sample indices, so two contexts that started at different positions can be compared).
 * Valid after a push of whole batch granules.  Returns the number of bytes written or a negative error. */
long wmb_boundary_state(wmb_ctx *c, uint8_t *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* WMBUS_B200_H */
