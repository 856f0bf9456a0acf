/*
 * wmb_framer.h -- host-side Wireless M-Bus framers (internal header).
 *
 * Frame-at-once replacements for the reference's per-bit decoder state machines
 * t1_c1_packet_decoder() (t1_c1_packet_decoder.h:649-712) and s1_packet_decoder()
 * (s1_packet_decoder.h:233-282): a candidate frame (flagged bit + following bits, as
 * gathered on the device) is decoded in one pass -- 3-out-of-6 / NRZ / Manchester,
 * L-field, RSSI abort, block CRCs, CRC strip.
 */
#ifndef WMB_FRAMER_H
#define WMB_FRAMER_H

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

/* Decode one candidate.  WMB_DEC_NEED_MORE is returned when the bit list ends while the
 * framer is still receiving. */
void wmb_frame_decode(const wmb_frame *f, wmb_decoded *out);

/* test hook: the same decode done by the device framer (kernel K4), n frames at once */
struct wmb_ctx;
int wmb_frame_decode_device(struct wmb_ctx *ctx, const wmb_frame *frames, size_t n, wmb_decoded *out);

uint16_t wmb_crc16(const uint8_t *data, size_t n);
unsigned wmb_tlg_length_format_a(unsigned l_field);

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
