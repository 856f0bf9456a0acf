/*
 * rtl_wmbus_b200.c -- drop-in host program: same command line, stdin and stdout contract
 * as the reference's main() (rtl_wmbus.c:855-967 options/usage, :1217-1372 main loop),
 * with the per-sample work done by libwmbus_b200 on a B200.
 *
 *   rtl_sdr -f 868.95M -s 1600000 - 2>/dev/null | rtl_wmbus_b200
 *   cat capture.cu8 | rtl_wmbus_b200 -v
 *
 * stdin : interleaved unsigned 8-bit I/Q at decimation x 800 kS/s, consumed in whole
 *         4096-byte items (a trailing partial item is dropped, rtl_wmbus.c:1301-1308)
 * stdout: MODE;CRC_OK;3OUTOF6OK;TIMESTAMP;PACKET_RSSI;CURRENT_RSSI;LINK_LAYER_IDENT_NO;0xDATAGRAM
 * Extra environment knobs (not options, so the argv surface stays the reference's):
 *   WMBUS_B200_DEVICE=<n>      CUDA device (default 0)
 *   WMBUS_B200_BATCH_MIB=<n>   bytes gathered before a device pass (default 64)
 */
#define _GNU_SOURCE
#include <errno.h>
#include <getopt.h>
#include <poll.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include "wmbus_b200.h"

static void print_usage(const char *program_name)
{
    /* same text as rtl_wmbus.c:869-884 */
    fprintf(stdout, "rtl_wmbus: %s\n\n", wmb_version_string());
    fprintf(stdout, "Usage %s:\n", program_name);
    fprintf(stdout, "\t-o remove DC offset\n");
    fprintf(stdout, "\t-a accelerate (use an inaccurate atan version)\n");
    fprintf(stdout, "\t-r 0 to disable run length algorithm\n");
    fprintf(stdout, "\t-t 0 to disable time2 algorithm\n");
    fprintf(stdout, "\t-d 2 set decimation rate to 2 (defaults to 2 if omitted)\n");
    fprintf(stdout, "\t-v show used algorithm in the output\n");
    fprintf(stdout, "\t-V show version\n");
    fprintf(stdout, "\t-s receive S1 and T1/C1 datagrams simultaneously. rtl_sdr _MUST_ be set to 868.625MHz (-f 868.625M)\n");
    fprintf(stdout, "\t-p [T,S] to disable processing T1/C1 or S1 mode\n");
    fprintf(stdout, "\t-f exit if flow of incoming data stops\n");
    fprintf(stdout, "\t-h print this help\n");
}

static void sig_alarm_handler(int signo)
{
    /* same message and exit status as rtl_wmbus.c:71-78.  The signal may be delivered to one of the CUDA runtime's
     * threads, where exit() would run the runtime's own exit handlers from inside itself: write() and _exit() instead
     * (stdout holds nothing: every line is flushed when it is printed). */
    static const char msg[] = "rtl_wmbus: exiting since incoming data stopped flowing!\n";
    (void)signo;
    if (write(STDERR_FILENO, msg, sizeof(msg) - 1) < 0) { /* nothing left to do about it */ }
    _exit(EXIT_FAILURE);
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* SIGALRM in `seconds` from now (0: cancel) */
static void set_alarm(double seconds)
{
    struct itimerval it;
    memset(&it, 0, sizeof(it));
    if (seconds > 0) {
        it.it_value.tv_sec = (time_t)seconds;
        it.it_value.tv_usec = (suseconds_t)((seconds - (double)(time_t)seconds) * 1e6);
        if (it.it_value.tv_sec == 0 && it.it_value.tv_usec == 0) it.it_value.tv_usec = 1;
    }
    setitimer(ITIMER_REAL, &it, NULL);
}

static int emit_lines(wmb_ctx *ctx, char *out, size_t outcap)
{
    for (;;) {
        size_t nl = 0;
        const size_t n = wmb_take_lines(ctx, out, outcap, &nl, 0);
        if (!nl) return 0;
        fwrite(out, 1, n, stdout);
        fflush(stdout);                                 /* t1_c1_packet_decoder.h:698-699 */
    }
}

int main(int argc, char *argv[])
{
    wmb_opts o;
    int check_flow = 0, option;
    wmb_default_opts(&o);

    if (argc == 1 && isatty(0)) {                       /* rtl_wmbus.c:1223-1228 */
        print_usage(argv[0]);
        exit(0);
    }

    while ((option = getopt(argc, argv, "ofad:p:r:vVst:")) != -1) {   /* rtl_wmbus.c:896 */
        switch (option) {
        case 'o': o.remove_dc = 1; break;
        case 'f': check_flow = 1; break;
        case 'a': o.accurate_atan = 0; break;
        case 'p':
            if (strcmp(optarg, "T") == 0 || strcmp(optarg, "t") == 0) o.t1c1_enabled = 0;
            else if (strcmp(optarg, "S") == 0 || strcmp(optarg, "s") == 0) o.s1_enabled = 0;
            else { print_usage(argv[0]); exit(EXIT_FAILURE); }
            break;
        case 'r':
            if (strcmp(optarg, "0") == 0) o.rla_enabled = 0;
            else { print_usage(argv[0]); exit(EXIT_FAILURE); }
            break;
        case 't':
            if (strcmp(optarg, "0") == 0) o.t2_enabled = 0;
            else { print_usage(argv[0]); exit(EXIT_FAILURE); }
            break;
        case 'd': o.decimation = (uint32_t)strtoul(optarg, NULL, 10); break;
        case 's': o.simultaneous = 1; break;
        case 'v': o.show_algorithm = 1; break;
        case 'V':
            fprintf(stdout, "rtl_wmbus: %s\n", wmb_version_string());
            fprintf(stdout, "libwmbus_b200 ABI %d\n", wmb_abi_version());
            exit(EXIT_SUCCESS);
        default:
            print_usage(argv[0]);
            exit(EXIT_FAILURE);
        }
    }

    if (check_flow) {                                   /* rtl_wmbus.c:1238-1246 */
        struct sigaction new_alarm;
        new_alarm.sa_handler = sig_alarm_handler;
        sigemptyset(&new_alarm.sa_mask);
        new_alarm.sa_flags = 0;
        fprintf(stderr, "rtl_wmbus: monitoring flow\n");
        sigaction(SIGALRM, &new_alarm, NULL);
    }

    const char *e;
    const int device = (e = getenv("WMBUS_B200_DEVICE")) ? atoi(e) : 0;
    unsigned long batch_mib = 64;
    if ((e = getenv("WMBUS_B200_BATCH_MIB")) != NULL) {
        char *end = NULL;
        const unsigned long v = strtoul(e, &end, 10);
        if (end != e && *end == 0 && e[0] != '-') batch_mib = v;
    }
    if (batch_mib < 1) batch_mib = 1;
    if (batch_mib > 1024) batch_mib = 1024;
    const size_t batch = (size_t)batch_mib * 1048576u;
    o.max_batch_mib = (uint32_t)batch_mib;

    wmb_ctx *ctx = NULL;
    if (wmb_create(&o, device, &ctx) != WMB_OK) {
        fprintf(stderr, "rtl_wmbus_b200: %s\n", wmb_last_error());
        return EXIT_FAILURE;
    }
    uint8_t *buf = wmb_host_alloc(batch);
    const size_t outcap = 1u << 20;
    char *out = malloc(outcap);
    if (!buf || !out) {
        fprintf(stderr, "rtl_wmbus_b200: out of memory\n");
        return EXIT_FAILURE;
    }

    /* device buffers are allocated at the first push: do it now, not when the first samples are waiting */
    if (wmb_push(ctx, buf, 0) != WMB_OK) {
        fprintf(stderr, "rtl_wmbus_b200: %s\n", wmb_last_error());
        return EXIT_FAILURE;
    }

    size_t fill = 0;
    /* -f: the reference arms a 2 s alarm around each fread() of one 4096-byte item (rtl_wmbus.c:1300-1302), i.e. it
     * gives up when a whole item does not arrive within 2 s -- a trickle of bytes does not keep it alive.  Here the
     * reads are whatever the pipe holds, so the alarm stays armed until 4096 new bytes have come in since it was set;
     * it is not running while the device works on a hand-over. */
    size_t since_arm = 0;
    int armed = 0;
    double deadline = 0;
    double last_push = now_s();
    int rc = WMB_OK, eof = 0;
    while (!eof) {
        /* wait for input; on a live stream hand over what has arrived every 100 ms so
         * that telegrams are printed promptly */
        struct pollfd pfd = { 0, POLLIN, 0 };
        if (check_flow && !armed) { deadline = now_s() + 2.0; set_alarm(2.0); armed = 1; since_arm = 0; }      /* START_ALARM */
        const int pr = poll(&pfd, 1, fill ? 100 : -1);
        ssize_t n = 0;
        if (pr > 0) {
            n = read(0, buf + fill, batch - fill);
            if (n < 0 && errno == EINTR) n = 0;
            else if (n <= 0) eof = 1;
            else { fill += (size_t)n; since_arm += (size_t)n; }
        } else if (pr < 0 && errno != EINTR) {
            eof = 1;
        }
        if (check_flow && armed && (since_arm >= 4096 || eof)) { set_alarm(0); armed = 0; }     /* STOP_ALARM: an item is in */
        const double t = now_s();
        if (fill == batch || eof || (fill && (pr == 0 || t - last_push > 0.1))) {
            if (check_flow && armed) {                  /* the watchdog times the input, not the device */
                set_alarm(0);
                const double t_in = now_s();
                rc = wmb_push(ctx, buf, fill);
                if (rc == WMB_OK) emit_lines(ctx, out, outcap);
                deadline += now_s() - t_in;             /* the item's two seconds do not run during the hand-over */
                const double left = deadline - now_s();
                set_alarm(left > 1e-3 ? left : 1e-3);
            } else {
                rc = wmb_push(ctx, buf, fill);
                if (rc == WMB_OK) emit_lines(ctx, out, outcap);
            }
            if (rc != WMB_OK) break;
            fill = 0;
            last_push = t;
        }
    }
    if (check_flow) set_alarm(0);
    if (rc == WMB_OK) {
        size_t nframes = 0;
        rc = wmb_poll(ctx, NULL, 0, &nframes, 1);       /* EOF: flush */
        emit_lines(ctx, out, outcap);
    }
    if (rc != WMB_OK) fprintf(stderr, "rtl_wmbus_b200: %s\n", wmb_last_error());
    free(out);
    wmb_host_free(buf);
    wmb_destroy(ctx);
    return rc == WMB_OK ? EXIT_SUCCESS : EXIT_FAILURE;
}
