/*
 * wmbus_b200_framer.h -- the host-side framers of libwmbus_b200.so as a C ABI of their own.
 *
 * Frame-at-once replacements for the reference's per-bit decoder state machines
 * t1_c1_packet_decoder() (t1_c1_packet_decoder.h:649-712) and s1_packet_decoder()
 * (s1_packet_decoder.h:233-282): a candidate frame (the bit carrying the access-code flag plus
 * the bits that follow it, as wmb_poll() hands them out) is decoded in one pass -- 3-out-of-6 /
 * NRZ / Manchester, L-field, RSSI abort, block CRCs, CRC strip -- and formatted as the
 * reference prints it (t1_c1_packet_decoder.h:670-699, rtl_wmbus_util.h:10-39).
 *
 * wmb_decode_frames() (wmbus_b200.h) is these functions plus the stream-order bookkeeping; a
 * caller who keeps its own bookkeeping (or only wants the CRC / line format) binds them directly.
 * The default path decodes on the device (kernel K4); wmb_frame_decode() is its host twin and
 * wmb_frame_decode_device() lets a test compare the two candidate by candidate.
 */
#ifndef WMBUS_B200_FRAMER_H
#define WMBUS_B200_FRAMER_H

#include <stddef.h>
#include <stdint.h>
#include "wmbus_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { WMB_DEC_ABORT = 0, WMB_DEC_LINE = 1, WMB_DEC_NEED_MORE = 2 };

typedef struct wmb_decoded {
    int      status;            /* WMB_DEC_*                                            */
    uint32_t consumed;          /* bits consumed incl. the flagged one (>= 1)           */
    uint64_t end_sample;        /* decimated sample of the last consumed bit            */
    char     mode[3];           /* "T1" / "C1" / "S1"                                   */
    uint8_t  crc_ok, ok_3of6;
    uint32_t packet_rssi, current_rssi;
    uint32_t serial;            /* LINK_LAYER_IDENT_NO                                  */
    uint32_t len;               /* datagram bytes after the CRC strip                   */
    uint8_t  datagram[292];
} wmb_decoded;

/* Decode one candidate on the host (t1_c1_packet_decoder.h:272-460, s1_packet_decoder.h:132-282).
 * WMB_DEC_NEED_MORE is returned when the bit list ends while the framer is still receiving. */
void wmb_frame_decode(const wmb_frame *f, wmb_decoded *out);

/* The same decode done by the device framer (kernel K4), n frames at once. */
int wmb_frame_decode_device(wmb_ctx *ctx, const wmb_frame *frames, size_t n, wmb_decoded *out);

/* CRC-16, polynomial 0x3D65, complemented (t1_c1_packet_decoder.h:463-469) */
uint16_t wmb_crc16(const uint8_t *data, size_t n);

/* "MODE;CRC_OK;3OUTOF6OK;TIMESTAMP;PACKET_RSSI;CURRENT_RSSI;IDENT;0xHEX\n"
 * (t1_c1_packet_decoder.h:670-699); returns the length written (excluding NUL). */
size_t wmb_format_line(const wmb_decoded *d, const char *algo_prefix, const char *timestamp,
                       char *buf, size_t cap);

/* YYYY-MM-DD HH:MM:SS.uuuuuu local time (rtl_wmbus_util.h:10-39) */
void wmb_make_time_string(char *ts, size_t n);

#ifdef __cplusplus
}
#endif
#endif
