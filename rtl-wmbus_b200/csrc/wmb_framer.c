/*
 * wmb_framer.c -- host-side T1 / C1 (frame A, B) / S1 framers of libwmbus_b200.
 *
 * The device hands over, for every access-code match, the flagged bit and the bits that
 * followed it (wmb_frame).  This file turns one such candidate into a datagram, doing in
 * one pass over the bit list what the reference does bit by bit in
 *   t1_c1_packet_decoder.h:272-460 (per-bit handlers), :649-712 (driver, RSSI abort),
 *   :463-536 (block CRCs), :551-636 (CRC strip)   and   s1_packet_decoder.h:132-282.
 */
#include "wmb_framer.h"

#include <stdio.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#define CAPTURE_THRESHOLD 5u           /* PACKET_CAPTURE_THRESHOLD, t1_c1_packet_decoder.h:36 */

/* ---- CRC-16, polynomial 0x3D65 (t1_c1_packet_decoder.h:463-469) ---------------- */

static uint16_t g_crc_tab[256];
static int g_crc_ready;

static void crc_init(void)
{
    for (unsigned i = 0; i < 256; i++) {
        unsigned c = i << 8;
        for (int b = 0; b < 8; b++) c = (c & 0x8000u) ? ((c << 1) ^ 0x3D65u) : (c << 1);
        g_crc_tab[i] = (uint16_t)c;
    }
    g_crc_ready = 1;
}

uint16_t wmb_crc16(const uint8_t *data, size_t n)
{
    if (!g_crc_ready) crc_init();
    unsigned crc = 0;
    for (size_t i = 0; i < n; i++) crc = (g_crc_tab[data[i] ^ (crc >> 8)] ^ (crc << 8)) & 0xFFFFu;
    return (uint16_t)(~crc & 0xFFFFu);
}

static int block_ok(const uint8_t *p, size_t n)      /* n includes the two CRC bytes */
{
    if (n < 2) return 0;
    return wmb_crc16(p, n - 2) == (uint16_t)((p[n - 2] << 8) | p[n - 1]);
}

/* format A: 10-byte first block, 16-byte blocks after it (:471-506) */
static int crc_check_a(const uint8_t *p, size_t n)
{
    if (n < 12 || !block_ok(p, 12)) return 0;
    for (size_t off = 12; off < n;) {
        const size_t blk = (n - off >= 18) ? 18 : n - off;
        if (!block_ok(p + off, blk)) return 0;
        off += blk;
    }
    return 1;
}

/* format B: CRC over the first 126 bytes, then over the rest (:508-536) */
static int crc_check_b(const uint8_t *p, size_t n)
{
    if (n < 12) return 0;
    for (size_t off = 0; off < n;) {
        const size_t blk = (n - off >= 128) ? 128 : n - off;
        if (!block_ok(p + off, blk)) return 0;
        off += blk;
    }
    return 1;
}

/* CRC strip, returns the stripped length (:551-592 format A, :595-636 format B) */
static unsigned strip_a(uint8_t *p, unsigned n)
{
    if (p[0] == 0 || n < 12) return 0;
    unsigned out = 10;
    for (unsigned off = 12; off < n;) {
        const unsigned blk = (n - off >= 18) ? 18 : n - off;
        memmove(p + out, p + off, blk - 2);
        out += blk - 2; off += blk;
    }
    return out;
}

static unsigned strip_b(uint8_t *p, unsigned n)
{
    if (p[0] < 2 || n < 12) return 0;
    unsigned out = 0;
    for (unsigned off = 0; off < n;) {
        const unsigned blk = (n - off >= 128) ? 128 : n - off;
        if (blk < 2) break;                       /* the reference reads out of bounds here */
        memmove(p + out, p + off, blk - 2);
        out += blk - 2; off += blk;
        p[0] = (uint8_t)(p[0] - 2);               /* :618, :630 */
    }
    return out;
}

/* ---- tables --------------------------------------------------------------------- */

/* EN 13757-4 3-out-of-6: code word -> nibble, 0xFF invalid (t1_c1_packet_decoder.h:50-65) */
static uint8_t nibble_3of6(unsigned c)
{
    static const int8_t tab[64] = {
        -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 3, -1, 1, 2, -1,
        -1, -1, -1, 7, -1, -1, 0, -1, -1, 5, 6, -1, 4, -1, -1, -1,
        -1, -1, -1, 11, -1, 9, 10, -1, -1, 15, -1, -1, 8, -1, -1, -1,
        -1, 13, 14, -1, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1 };
    const int v = tab[c & 63u];
    return v < 0 ? 0xFF : (uint8_t)v;
}

static unsigned wmb_tlg_length_format_a(unsigned L)      /* t1_c1_packet_decoder.h:68-96 */
{
    return 1 + L + 2 * (1 + (L > 9 ? (L - 9 + 15) / 16 : 0));
}

/* ---- bit cursor ----------------------------------------------------------------- */

typedef struct cursor {
    const wmb_frame *f;
    uint32_t pos;            /* index of the last consumed bit */
    int stop;                /* 0 running, WMB_DEC_ABORT+10 / WMB_DEC_NEED_MORE+10 */
} cursor;

enum { STOP_ABORT = 1, STOP_MORE = 2 };

/* Consume `n` bits MSB first.  After every bit except the telegram's very last one the
 * reference drops the packet when rssi < 5 (:703-710).  Returns 0 on success. */
static int take(cursor *c, unsigned n, int last_of_telegram, unsigned *value)
{
    unsigned v = 0;
    for (unsigned k = 0; k < n; k++) {
        if (c->pos + 1 >= c->f->nbits) { c->stop = STOP_MORE; return 1; }
        c->pos++;
        const wmb_bit w = c->f->bits[c->pos];
        v = (v << 1) | WMB_BIT_DATA(w);
        const int final_bit = last_of_telegram && k + 1 == n;
        if (!final_bit && WMB_BIT_RSSI(w) < CAPTURE_THRESHOLD) { c->stop = STOP_ABORT; return 1; }
    }
    *value = v;
    return 0;
}

static void finish(const cursor *c, wmb_decoded *d, const char *mode, uint8_t *pkt, unsigned len,
                   int bframe, unsigned err3of6)
{
    const wmb_bit last = c->f->bits[c->pos];
    d->status = WMB_DEC_LINE;
    d->consumed = c->pos + 1;
    d->end_sample = c->f->sync_sample + WMB_BIT_OFFSET(last);
    memcpy(d->mode, mode, 3);
    d->crc_ok = (uint8_t)(bframe ? crc_check_b(pkt, len) : crc_check_a(pkt, len));
    d->ok_3of6 = (uint8_t)(err3of6 ^ 1u);
    d->packet_rssi = WMB_BIT_RSSI(c->f->bits[1]);            /* rssi at the first bit after sync (:295) */
    d->current_rssi = WMB_BIT_RSSI(last);
    memcpy(&d->serial, pkt + 4, 4);                           /* get_serial(), :638-645 (host is LE) */
    d->len = bframe ? strip_b(pkt, len) : strip_a(pkt, len);
    memcpy(d->datagram, pkt, sizeof(d->datagram));
}

static void stopped(const cursor *c, wmb_decoded *d)
{
    d->status = (c->stop == STOP_MORE) ? WMB_DEC_NEED_MORE : WMB_DEC_ABORT;
    d->consumed = c->pos + 1;
    d->end_sample = c->f->sync_sample + WMB_BIT_OFFSET(c->f->bits[c->pos]);
}

/* ---- T1 and C1 ------------------------------------------------------------------ */

static void decode_t1c1(cursor *c, wmb_decoded *d)
{
    uint8_t pkt[292];
    memset(pkt, 0, sizeof(pkt));
    unsigned hi6, lo6, v;

    if (take(c, 6, 0, &hi6) || take(c, 6, 0, &lo6)) { stopped(c, d); return; }
    const unsigned hi = nibble_3of6(hi6), lo = nibble_3of6(lo6);
    const unsigned mode = (hi6 << 6) | lo6;

    if (hi != 0xFF && lo != 0xFF) {
        /* T1: 3-out-of-6 coded L-field and data (:298-392) */
        const unsigned L = (hi << 4) | lo;
        const unsigned len = wmb_tlg_length_format_a(L);
        unsigned err = 0, l = 0;
        pkt[l++] = (uint8_t)L;
        while (l < len) {
            const int last = (l + 1 >= len);
            if (take(c, 6, 0, &hi6) || take(c, 6, last, &lo6)) { stopped(c, d); return; }
            const unsigned h = nibble_3of6(hi6), lw = nibble_3of6(lo6);
            if (h == 0xFF || lw == 0xFF) err = 1;
            pkt[l++] = (uint8_t)((h == 0xFF ? 0xFFu : h << 4) | lw);
        }
        finish(c, d, "T1", pkt, len, 0, err);
        return;
    }
    if (mode != 0x54Cu && mode != 0x543u) {        /* neither L-field nor C1 mode word (:334-337) */
        c->stop = STOP_ABORT; stopped(c, d); return;
    }
    /* C1: 4-bit trailer, 8-bit L, NRZ bytes (:399-460) */
    const int bframe = (mode == 0x543u);
    if (take(c, 4, 0, &v)) { stopped(c, d); return; }
    if (v != 0xDu) { c->stop = STOP_ABORT; stopped(c, d); return; }
    if (take(c, 8, 0, &v)) { stopped(c, d); return; }
    const unsigned len = bframe ? 1 + v : wmb_tlg_length_format_a(v);
    unsigned l = 0;
    pkt[l++] = (uint8_t)v;
    do {
        const int last = (l + 1 >= len);
        if (take(c, 8, last, &v)) { stopped(c, d); return; }
        pkt[l++] = (uint8_t)v;
    } while (l < len);
    finish(c, d, "C1", pkt, len, bframe, 0);
}

/* ---- S1 ------------------------------------------------------------------------- */

/* one Manchester coded byte: 16 chips, "01" = 1, "10" = 0 (s1_packet_decoder.h:35-37, :152-168) */
static int take_manchester_byte(cursor *c, int last_of_telegram, unsigned *value)
{
    unsigned v = 0;
    for (int k = 0; k < 8; k++) {
        unsigned a, b;
        if (take(c, 1, 0, &a)) return 1;
        /* the violation check runs before the rssi check on the second chip */
        if (c->pos + 1 >= c->f->nbits) { c->stop = STOP_MORE; return 1; }
        c->pos++;
        const wmb_bit w = c->f->bits[c->pos];
        b = WMB_BIT_DATA(w);
        if (a == b) { c->stop = STOP_ABORT; return 1; }
        v = (v << 1) | b;
        const int final_bit = last_of_telegram && k == 7;
        if (!final_bit && WMB_BIT_RSSI(w) < CAPTURE_THRESHOLD) { c->stop = STOP_ABORT; return 1; }
    }
    *value = v;
    return 0;
}

static void decode_s1(cursor *c, wmb_decoded *d)
{
    uint8_t pkt[292];
    memset(pkt, 0, sizeof(pkt));
    unsigned v;
    if (take_manchester_byte(c, 0, &v)) { stopped(c, d); return; }
    const unsigned len = wmb_tlg_length_format_a(v);
    unsigned l = 0;
    pkt[l++] = (uint8_t)v;
    while (l < len) {
        const int last = (l + 1 >= len);
        if (take_manchester_byte(c, last, &v)) { stopped(c, d); return; }
        pkt[l++] = (uint8_t)v;
    }
    finish(c, d, "S1", pkt, len, 0, 0);
}

void wmb_frame_decode(const wmb_frame *f, wmb_decoded *d)
{
    memset(d, 0, sizeof(*d));
    cursor c = { f, 0, 0 };
    if (f->nbits == 0) { d->status = WMB_DEC_NEED_MORE; return; }
    /* the flagged bit itself: idle handler keeps the state, then the rssi check (:703-710) */
    if (WMB_BIT_RSSI(f->bits[0]) < CAPTURE_THRESHOLD) { c.stop = STOP_ABORT; stopped(&c, d); return; }
    if (f->chain == WMB_CHAIN_T1C1) decode_t1c1(&c, d);
    else decode_s1(&c, d);
}

/* ---- output --------------------------------------------------------------------- */

void wmb_make_time_string(char *ts, size_t n)
{
    /* the calendar part only changes once a second: localtime_r/strftime are redone when tv_sec moves */
    static __thread time_t cached_sec = (time_t)-1;
    static __thread char cached_fmt[64];
    struct timeval tv;
    if (gettimeofday(&tv, NULL) != 0) { if (n) ts[0] = 0; return; }
    if (tv.tv_sec != cached_sec) {
        struct tm tmv;
        if (localtime_r(&tv.tv_sec, &tmv) == NULL) { if (n) ts[0] = 0; return; }
        strftime(cached_fmt, sizeof(cached_fmt), "%Y-%m-%d %H:%M:%S.", &tmv);
        cached_sec = tv.tv_sec;
    }
    const size_t l = strlen(cached_fmt);
    if (n < l + 7) { if (n) ts[0] = 0; return; }
    memcpy(ts, cached_fmt, l);
    unsigned us = (unsigned)tv.tv_usec;
    for (int i = 5; i >= 0; i--) { ts[l + (size_t)i] = (char)('0' + us % 10u); us /= 10u; }
    ts[l + 6] = 0;
}

static size_t put_str(char *buf, size_t len, size_t cap, const char *s)
{
    while (*s && len + 1 < cap) buf[len++] = *s++;
    return len;
}

static size_t put_u32(char *buf, size_t len, size_t cap, uint32_t v)
{
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (n && len + 1 < cap) buf[len++] = tmp[--n];
    return len;
}

/* hand-rolled (no printf machinery): a batch can carry thousands of lines */
size_t wmb_format_line(const wmb_decoded *d, const char *algo_prefix, const char *timestamp,
                       char *buf, size_t cap)
{
    static const char hexd[] = "0123456789abcdef", hexu[] = "0123456789ABCDEF";
    if (cap < 2) return 0;
    size_t len = 0;
    if (algo_prefix) len = put_str(buf, len, cap, algo_prefix);
    len = put_str(buf, len, cap, d->mode);
    if (len + 1 < cap) buf[len++] = ';';
    len = put_u32(buf, len, cap, d->crc_ok);
    if (len + 1 < cap) buf[len++] = ';';
    len = put_u32(buf, len, cap, d->ok_3of6);
    if (len + 1 < cap) buf[len++] = ';';
    len = put_str(buf, len, cap, timestamp);
    if (len + 1 < cap) buf[len++] = ';';
    len = put_u32(buf, len, cap, d->packet_rssi);
    if (len + 1 < cap) buf[len++] = ';';
    len = put_u32(buf, len, cap, d->current_rssi);
    if (len + 1 < cap) buf[len++] = ';';
    for (int sh = 28; sh >= 0 && len + 1 < cap; sh -= 4) buf[len++] = hexu[(d->serial >> sh) & 15u];   /* %08X */
    len = put_str(buf, len, cap, ";0x");
    for (uint32_t i = 0; i < d->len && len + 3 < cap; i++) {
        buf[len++] = hexd[d->datagram[i] >> 4];
        buf[len++] = hexd[d->datagram[i] & 15];
    }
    if (len + 1 < cap) buf[len++] = '\n';
    buf[len] = 0;
    return len;
}
