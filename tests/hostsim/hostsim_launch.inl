/*
 * hostsim_launch.inl -- TEST INFRASTRUCTURE ONLY (see hostsim_cuda.h).
 * CPU "launches" of the kernels in wmb_kernels.cuh: every kernel becomes a loop over its
 * (block, thread) indices; barrier-separated phases of the demod kernel run one after
 * another over all threads, which is what __syncthreads() guarantees on the device.
 */

/* Order in which the simulated threads (or lanes, or blocks) of one barrier-separated phase run.  On the device they run
 * concurrently, so nothing may depend on it: WMB_HOSTSIM_ORDER=1 runs every phase backwards, 2 in a scrambled order
 * (i -> (a * i + b) mod n, a coprime to n).  A missing barrier or a lane that reads what another lane of the same launch
 * writes shows up as a result that changes with the order (tests/test_hostsim_pipeline.py runs the suite's core in all three). */
static int hs_order()
{
    static int o = -1;
    if (o < 0) { const char *e = getenv("WMB_HOSTSIM_ORDER"); o = e ? atoi(e) : 0; }
    return o;
}
static void hs_for_impl(uint32_t n, void (*fn)(void *, uint32_t), void *ctx)
{
    const int o = hs_order();
    if (o == 1) { for (uint32_t i = n; i--;) fn(ctx, i); return; }
    if (o == 2 && n > 2) {
        uint64_t a = (uint64_t)(n * 0.6180339887) | 1u;
        auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
        while (gcd(a, n) != 1) a += 2;
        const uint64_t b = n / 3;
        for (uint32_t i = 0; i < n; i++) fn(ctx, (uint32_t)((a * i + b) % n));
        return;
    }
    for (uint32_t i = 0; i < n; i++) fn(ctx, i);
}
/* (one loop body per call site: the phase functions are large and fully unrolled) */
template <class F>
static void hs_for(uint32_t n, F f)
{
    hs_for_impl(n, [](void *c, uint32_t i) { (*(F *)c)(i); }, &f);
}

template <class CH>
static void hostsim_k1_chain(const K1Params &p, K1Smem &sm, const uint8_t *raw, int64_t tile, bool need_convert)
{
    const bool fast = (p.d == 2 && !p.mix);
    if (need_convert) hs_for(K1_THREADS, [&](uint32_t t) {
        if (p.prefilter) k1_convert_float<CH::ID>(p, sm, raw, tile, t);
        else if (fast) k1_convert_fast(p, sm, raw, tile, t); else k1_convert<CH::ID>(p, sm, raw, tile, t);
    });
    if (p.prefilter) {
        hs_for(K1_THREADS, [&](uint32_t t) { k1_prefir(p, sm, t); });
        hs_for(K1_THREADS, [&](uint32_t t) { k1_disc_mag_general(p, sm, t); });
    }
    else if (fast) { hs_for(K1_THREADS, [&](uint32_t t) { k1_box_disc<CH, 1, true>(p, sm, t); }); }
    else if (p.d == 3) { hs_for(K1_THREADS, [&](uint32_t t) { k1_box_disc<CH, 3, false>(p, sm, t); }); }
    else if (p.d == 2) { hs_for(K1_THREADS, [&](uint32_t t) { k1_box_disc<CH, 2, false>(p, sm, t); }); }
    else if (p.d == 1) { hs_for(K1_THREADS, [&](uint32_t t) { k1_box_disc<CH, 1, false>(p, sm, t); }); }
    else {
        hs_for(K1_THREADS, [&](uint32_t t) { k1_box<CH>(p, sm, t); });
        hs_for(K1_THREADS, [&](uint32_t t) { k1_disc_mag(p, sm, t); });
    }
    hs_for(K1_THREADS, [&](uint32_t t) { k1_fir<CH>(p, sm, tile, t); });
    hs_for(32, [&](uint32_t t) { k1_rssi<CH>(p, sm.mag, tile, t); });            /* the block's RSSI warp */
}

static int launch_k1(wmb_ctx *c, const K1Params &p, cudaStream_t)
{
    const int64_t ntiles = (p.M + K1_TILE - 1) / K1_TILE;
    std::vector<uint8_t> smem(k1_smem_bytes(p.d, p.prefilter) + 256, 0xA5);   /* garbage-filled like real smem */
    K1Smem sm;
    uint8_t *base = smem.data();
    base += (128 - ((uintptr_t)base & 127)) & 127;
    k1_carve(sm, base, p.d, p.prefilter);
    hs_for(WMB_ATAN_TAB_ELEMS, [&](uint32_t i) { wmb_atan_tab_fill((WmbAtanTab *)base, i); });
    int rc = WMB_OK;
    hs_for((uint32_t)ntiles, [&](uint32_t tile) {
        const K1Load L = k1_plan_load(p, tile);
        if (L.n0) memcpy(sm.bytes[0], L.src0, (size_t)L.n0);
        if (L.n1) memcpy(sm.bytes[0] + L.off1, L.src1, (size_t)L.n1);
        if ((L.n0 | L.n1 | L.off1) & 15) { rc = set_err(WMB_E_STATE, "hostsim: unaligned bulk copy"); return; }
        const uint8_t *raw = sm.bytes[0];
        if (p.chains & 1u) hostsim_k1_chain<ChainT1C1>(p, sm, raw, tile, true);
        if (p.chains & 2u) hostsim_k1_chain<ChainS1>(p, sm, raw, tile, p.mix || !(p.chains & 1u));
    });
    c->st.kernel_launches++;
    return rc;
}

/* speculative pass, verification, and the fix-up loop of k2*_fixup_kernel (re-run refuted lanes from their
 * predecessors' exact end states until none is left) */
template <class RERUN, class VERIFY>
static int hostsim_fixup(wmb_ctx *c, uint32_t lanes, uint32_t *n_fail, RERUN rerun, VERIFY verify)
{
    for (uint32_t round = 0; *n_fail; round++) {
        if (round > lanes + 2) { *c->d_errors |= 256u; *n_fail = 0; break; }
        c->d_gd->lanes_rerun += *n_fail;
        *n_fail = 0;
        hs_for(lanes, [&](uint32_t lane) { rerun(lane); });
        hs_for(lanes, [&](uint32_t lane) { verify(lane); });
    }
    return WMB_OK;
}

static int launch_k2a_lanes(wmb_ctx *c, int chain, const K2aParams &p, cudaStream_t)
{
    hs_for(p.lanes, [&](uint32_t lane) {
        if (chain == 0) k2a_lane<ChainT1C1>(p, lane);
        else            k2a_lane<ChainS1>(p, lane);
    });
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k2a_verify(wmb_ctx *c, int chain, const K2aParams &p0)
{
    K2aParams p = p0;
    uint32_t *nf = c->d_nfail + chain;
    hs_for(p.lanes, [&](uint32_t lane) { k2a_verify_lane(p, lane, nf); });
    p.mode = 1;
    c->st.kernel_launches += 2;
    return hostsim_fixup(c, p.lanes, nf,
                         [&](uint32_t lane) { if (chain == 0) k2a_lane<ChainT1C1>(p, lane); else k2a_lane<ChainS1>(p, lane); },
                         [&](uint32_t lane) { k2a_verify_lane(p, lane, nf); });
}

static int launch_k2m(wmb_ctx *c, int chain, const K2mParams &p0, cudaStream_t)
{
    K2mParams p = p0;
    uint32_t *nf = c->d_nfail + 2 + chain;
    hs_for(p.lanes, [&](uint32_t lane) {
        if (chain == 0) k2m_lane<ChainT1C1>(p, lane);
        else            k2m_lane<ChainS1>(p, lane);
    });
    hs_for(p.lanes, [&](uint32_t lane) { k2m_verify_lane(p, lane, nf); });
    p.mode = 1;
    c->st.kernel_launches += 3;
    return hostsim_fixup(c, p.lanes, nf,
                         [&](uint32_t lane) { if (chain == 0) k2m_lane<ChainT1C1>(p, lane); else k2m_lane<ChainS1>(p, lane); },
                         [&](uint32_t lane) { k2m_verify_lane(p, lane, nf); });
}

static int launch_k2p1(wmb_ctx *c, const K2p1Params &p0)
{
    K2p1Params p = p0;
    uint32_t *nf = c->d_nfail + 4;
    hs_for(p.lanes, [&](uint32_t lane) { k2p1_lane(p, lane); });
    hs_for(p.lanes, [&](uint32_t lane) { k2p1_verify_lane(p, lane, nf); });
    p.mode = 1;
    c->st.kernel_launches += 3;
    return hostsim_fixup(c, p.lanes, nf, [&](uint32_t lane) { k2p1_lane(p, lane); },
                         [&](uint32_t lane) { k2p1_verify_lane(p, lane, nf); });
}

static void launch_cscan(wmb_ctx *c, const uint32_t *cnt, uint64_t *base, uint32_t n, uint64_t *agg, uint64_t *total,
                         const uint32_t *skip = nullptr, uint32_t *clear = nullptr, uint32_t from_zero = 0, cudaStream_t = nullptr,
                         uint32_t skip_invert = 0)
{
    CountScan s;
    s.cnt = cnt; s.base = base; s.n = n; s.agg = agg; s.total = total; s.skip = skip; s.clear = clear; s.from_zero = from_zero;
    s.skip_invert = skip_invert;
    static uint64_t part[SCAN_BLOCK];
    const uint32_t tiles = scan_tiles(n);
    hs_for(tiles, [&](uint32_t b) {
        hs_for(SCAN_BLOCK, [&](uint32_t t) { cscan_local(s, b, t, part); });
        cscan_a_finish(s, b, part);
    });
    cscan_b(s);
    hs_for(tiles, [&](uint32_t b) {
        hs_for(SCAN_BLOCK, [&](uint32_t t) { cscan_local(s, b, t, part); });
        cscan_c_block(s, b, part);
        hs_for(SCAN_BLOCK, [&](uint32_t t) { cscan_c_write(s, b, t, part); });
    });
    c->st.kernel_launches += 3;
}

static int launch_k2p_rest(wmb_ctx *c, const K2pcParams &pc, K2p2Params p2)
{
    launch_cscan(c, pc.cnt, pc.base, pc.lanes, pc.agg, &pc.pd->n_rec, nullptr, &pc.pd->fallback, 1);
    hs_for(pc.lanes, [&](uint32_t lane) {
        hs_for(4, [&](uint32_t t) { k2pc_compact(pc, lane, t, 4); });
    });
    hs_for(p2.lanes, [&](uint32_t lane) { k2p2_count(p2, lane); });
    {
        static uint32_t part[K2P2W_THREADS];
        hs_for(p2.lanes, [&](uint32_t lane) {
            hs_for(K2P2W_THREADS, [&](uint32_t t) { k2p2w_a(p2, lane, t, part); });
            k2p2_sum_finish(p2, lane, part);
        });
    }
    launch_cscan(c, p2.cnt, p2.base, p2.lanes, p2.agg, &p2.sd->total, &p2.pd->fallback);
    {
        static uint32_t part[K2P2W_THREADS];
        hs_for(p2.lanes, [&](uint32_t lane) {
            hs_for(K2P2W_THREADS, [&](uint32_t t) { k2p2w_a(p2, lane, t, part); });
            k2p2w_b(part, 0);
            hs_for(K2P2W_THREADS, [&](uint32_t t) { k2p2w_c(p2, lane, t, part); });
        });
    }
    c->st.kernel_launches += 4;
    return WMB_OK;
}

static int launch_k2p_fold(wmb_ctx *c, const P1State *p1_end_last, RlState *p2_out, RlState *carry, const K2pDev *pd,
                           const RlState *mono_end)
{
    k2p_fold(p1_end_last, p2_out, carry, pd, mono_end, GD_FIELD(c, rl_fallbacks));
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k2m_carry(wmb_ctx *c, const RlState *end, RlState *carry, const uint32_t *run_if, cudaStream_t)
{
    if (!run_if || *run_if) *carry = *end;
    c->st.kernel_launches += 1;
    return WMB_OK;
}

template <class CH>
static void hostsim_k2t(const K2tParams &p)
{
    hs_for(p.lanes, [&](uint32_t lane) { k2t_count<CH>(p, lane); });
    static T2Fold part[SCAN_BLOCK];
    const uint32_t tiles = scan_tiles(p.lanes);
    hs_for(tiles, [&](uint32_t b) {
        hs_for(SCAN_BLOCK, [&](uint32_t t) { t2scan_local<CH>(p, b, t, part); });
        t2scan_a_finish<CH>(p, b, part);
    });
    t2scan_b<CH>(p);
    hs_for(tiles, [&](uint32_t b) {
        hs_for(SCAN_BLOCK, [&](uint32_t t) { t2scan_local<CH>(p, b, t, part); });
        t2scan_c_block<CH>(p, b, part);
        hs_for(SCAN_BLOCK, [&](uint32_t t) { t2scan_c_write<CH>(p, b, t, part); });
    });
    hs_for(p.lanes, [&](uint32_t lane) { k2t_write<CH>(p, lane); });
}

static int launch_k2t(wmb_ctx *c, int chain, const K2tParams &p)
{
    if (chain == 0) hostsim_k2t<ChainT1C1>(p); else hostsim_k2t<ChainS1>(p);
    c->st.kernel_launches += 5;
    return WMB_OK;
}

static int launch_k2c(wmb_ctx *c, const K2cParams &p, cudaStream_t)
{
    launch_cscan(c, p.cnt, p.base, p.lanes, p.agg, &p.sd->total, p.run_if, nullptr, 0, nullptr, p.run_if ? 1u : 0u);
    hs_for(p.lanes, [&](uint32_t lane) {
        hs_for(4, [&](uint32_t t) { k2c_compact(p, lane, t, 4); });
    });
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k3_k4(wmb_ctx *c, const K3Params &p, const K4Params *q)
{
    k3_plan(p);
    const uint32_t n = p.gd->n;
    hs_for(n, [&](uint32_t i) { k3_fill(p, i, 0, 1); });
    hs_for(n, [&](uint32_t i) { k3_size(p, i); });
    hs_for(n, [&](uint32_t i) {
        hs_for(4, [&](uint32_t t) { k3_cut(p, i, t, 4); });
    });
    hs_for(SCAN_THREADS, [&](uint32_t t) { k3_offsets_a(p, t); });
    k3_offsets_b(p);
    hs_for(SCAN_THREADS, [&](uint32_t t) { k3_offsets_c(p, t); });
    hs_for(n, [&](uint32_t i) {
        hs_for(4, [&](uint32_t t) { k3_copy(p, i, t, 4); });
    });
    hs_for((n > WMB_N_STREAMS ? n : WMB_N_STREAMS), [&](uint32_t i) { k3_carry(p, i); });
    c->st.kernel_launches += 8;
    if (q) {
        static K4Smem sm;                   /* the block's phases need real barriers: one simulated thread */
        hs_for(n, [&](uint32_t i) { k4_decode(*q, i, 0, 1, sm); });
        c->st.kernel_launches += 1;
    }
    k3_publish(p);
    return WMB_OK;
}

static int launch_k4(wmb_ctx *c, const K4Params &p)
{
    static K4Smem sm;                   /* the block's phases need real barriers: one simulated thread */
    hs_for(p.n, [&](uint32_t i) { k4_decode(p, i, 0, 1, sm); });
    c->st.kernel_launches += 1;
    return WMB_OK;
}

/* ---- test hooks into the device arithmetic (CPU build only) ---- */
extern "C" float hostsim_atan2f_bounded(float y, float x)
{
    static WmbAtanTab tab;
    static bool filled = false;
    if (!filled) { for (int i = 0; i < WMB_ATAN_TAB_ELEMS; i++) wmb_atan_tab_fill(&tab, i); filled = true; }
    return wmb_atan2f_bounded(y, x, &tab);
}
extern "C" float hostsim_atan2f_general(float y, float x) { return wmb_atan2f(y, x); }
extern "C" int hostsim_div_small(int x, int n) { return wmb_div_small(x, n); }
extern "C" int hostsim_div_pow2(int x, int s) { return wmb_div_pow2(x, s); }
