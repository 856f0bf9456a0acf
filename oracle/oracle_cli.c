/*
 * oracle_cli.c -- TEST INFRASTRUCTURE ONLY.
 * Command-line front end of the CPU oracle with the reference's option surface
 * (rtl_wmbus.c:892-967): reads cu8 from stdin, prints datagram lines.
 * Used by tests (diff against oracle/_ref/rtl_wmbus) and by bench.py's
 * cpu_baseline leg when the compiled reference is not available.
 *   extra option:  -T   print the literal TS instead of a wall-clock timestamp
 */
#include "wmbus_oracle.h"

#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    orc_opts o;
    orc_default_opts(&o);
    o.real_timestamp = 1;
    int c;
    while ((c = getopt(argc, argv, "ofad:p:r:vVst:T")) != -1) {
        switch (c) {
        case 'o': o.remove_dc = 1; break;
        case 'f': break;
        case 'a': o.accurate_atan = 0; break;
        case 'p':
            if (!strcmp(optarg, "T") || !strcmp(optarg, "t")) o.t1c1_enabled = 0;
            else if (!strcmp(optarg, "S") || !strcmp(optarg, "s")) o.s1_enabled = 0;
            else return EXIT_FAILURE;
            break;
        case 'r': if (!strcmp(optarg, "0")) o.rla_enabled = 0; else return EXIT_FAILURE; break;
        case 't': if (!strcmp(optarg, "0")) o.t2_enabled = 0; else return EXIT_FAILURE; break;
        case 'd': o.decimation = (uint32_t)strtoul(optarg, NULL, 10); break;
        case 's': o.simultaneous = 1; break;
        case 'v': o.show_algorithm = 1; break;
        case 'T': o.real_timestamp = 0; break;
        case 'V': printf("wmbus oracle (restatement of rtl-wmbus b6a7705)\n"); return 0;
        default: return EXIT_FAILURE;
        }
    }
    size_t cap = 1 << 24, n = 0;
    uint8_t *buf = malloc(cap);
    for (;;) {
        if (n == cap) { cap *= 2; buf = realloc(buf, cap); }
        size_t r = fread(buf + n, 1, cap - n, stdin);
        if (!r) break;
        n += r;
    }
    size_t outcap = 1 << 26, lines = 0;
    char *out = malloc(outcap);
    size_t len = orc_run(buf, n, &o, out, outcap, &lines);
    fwrite(out, 1, len < outcap ? len : outcap - 1, stdout);
    free(out); free(buf);
    return 0;
}
