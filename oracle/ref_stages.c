/*
 * ref_stages.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Stage-level probe of the UNMODIFIED reference: this file #includes the
 * reference translation unit where it lies under /root/reference (nothing is
 * copied into this repo), renames its main(), and drives the reference's own
 * static leaf functions (moving_average_*, polar_discriminator_*, lp_fir_*,
 * *_remove_dc_offset_demod, rssi_filter_*, bp_iir_cheb1_*) in the order the
 * reference's sample loop does (rtl_wmbus.c:1310-1355, :1038-1116, :1130-1208),
 * writing every intermediate value to stdout as binary records.
 *
 * Built by oracle/Makefile into oracle/_ref/ref_stages (git-ignored).
 *
 *   ref_stages <chain 0|1> <decimation> <accurate 0|1> <remove_dc 0|1> <simultaneous 0|1> [prefilter 0..4] < in.cu8 > out.f32
 *   prefilter 1..4: one of the reference's dormant pre-decimation low-passes instead of the moving averages in front of
 *   the decimation -- 1 lp_fir_butter_1600kHz_160kHz_200kHz_* (rtl_wmbus.c:197-233), 2 lp_ppf_butter_... (:258-295,
 *   ppf.h), 3 lp_firfp_butter_... (:235-256, fixedptc), 4 lp_ppffp_butter_... (:297-333).  2..4 keep ONE filter state
 *   for both chains, so only the requested chain's samples go through them.
 *
 * Output: per decimated sample 6 floats: si, sq, dphi_raw, dphi, rssi, clock(0/1).
 */
#define main rtl_wmbus_reference_main
#include "rtl_wmbus.c"
#undef main

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: ref_stages chain d accurate dc simul\n"); return 2; }
    const int chain = atoi(argv[1]);
    const unsigned d = (unsigned)strtoul(argv[2], NULL, 10);
    const int accurate = atoi(argv[3]);
    const int dc = atoi(argv[4]);
    const int simul = atoi(argv[5]);
    const int pre = argc > 6 ? atoi(argv[6]) : 0;
    const int fs_kHz = (int)(d * 800u);
    uint8_t block[4096];
    unsigned idx = 0;

    setup_lookup_tables_for_frequency_translation(fs_kHz);

    while (fread(block, sizeof(block), 1, stdin) == 1) {
        for (size_t k = 0; k < sizeof(block); k += 2) {
            float it = (float)block[k] - 127.5f, qt = (float)block[k + 1] - 127.5f;
            float is = it, qs = qt;
            if (simul) shift_freq_plus_minus325(&it, &qt, &is, &qs, fs_kHz);
            float i_t1, q_t1, i_s1, q_s1;
            if (pre >= 2) {
                const float xi = chain == 0 ? it : is, xq = chain == 0 ? qt : qs;
                const float yi = pre == 2 ? lp_ppf_butter_1600kHz_160kHz_200kHz(xi, 0)
                               : pre == 3 ? lp_firfp_butter_1600kHz_160kHz_200kHz(xi, 0)
                                          : lp_ppffp_butter_1600kHz_160kHz_200kHz(xi, 0);
                const float yq = pre == 2 ? lp_ppf_butter_1600kHz_160kHz_200kHz(xq, 1)
                               : pre == 3 ? lp_firfp_butter_1600kHz_160kHz_200kHz(xq, 1)
                                          : lp_ppffp_butter_1600kHz_160kHz_200kHz(xq, 1);
                i_t1 = i_s1 = yi; q_t1 = q_s1 = yq;
            } else {
                i_t1 = pre ? lp_fir_butter_1600kHz_160kHz_200kHz_t1_c1(it, 0) : moving_average_t1_c1(it, 0);
                q_t1 = pre ? lp_fir_butter_1600kHz_160kHz_200kHz_t1_c1(qt, 1) : moving_average_t1_c1(qt, 1);
                i_s1 = pre ? lp_fir_butter_1600kHz_160kHz_200kHz_s1(is, 0) : moving_average_s1(is, 0);
                q_s1 = pre ? lp_fir_butter_1600kHz_160kHz_200kHz_s1(qs, 1) : moving_average_s1(qs, 1);
            }
            if (++idx < d) continue;
            idx = 0;
            float rec[6];
            if (chain == 0) {
                rec[0] = i_t1; rec[1] = q_t1;
                rec[2] = accurate ? polar_discriminator_t1_c1(i_t1, q_t1)
                                  : polar_discriminator_t1_c1_inaccurate(i_t1, q_t1);
                rec[3] = lp_fir_butter_800kHz_100kHz_160kHz(rec[2]);
                if (dc) rec[3] = t1_c1_remove_dc_offset_demod(rec[3]);
                rec[4] = rssi_filter_t1_c1(sqrtf(i_t1 * i_t1 + q_t1 * q_t1));
                rec[5] = (bp_iir_cheb1_800kHz_90kHz_98kHz_102kHz_110kHz(rec[3] * rec[3]) >= 0) ? 1.f : 0.f;
            } else {
                rec[0] = i_s1; rec[1] = q_s1;
                rec[2] = accurate ? polar_discriminator_s1(i_s1, q_s1)
                                  : polar_discriminator_s1_inaccurate(i_s1, q_s1);
                rec[3] = lp_fir_butter_800kHz_32kHz_36kHz(rec[2]);
                if (dc) rec[3] = s1_remove_dc_offset_demod(rec[3]);
                rec[4] = rssi_filter_s1(sqrtf(i_s1 * i_s1 + q_s1 * q_s1));
                rec[5] = (bp_iir_cheb1_800kHz_22kHz_30kHz_34kHz_42kHz(rec[3] * rec[3]) >= 0) ? 1.f : 0.f;
            }
            fwrite(rec, sizeof(rec), 1, stdout);
        }
    }
    return 0;
}
