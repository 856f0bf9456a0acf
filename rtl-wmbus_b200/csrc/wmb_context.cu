/*
 * wmb_context.cu -- host side of libwmbus_b200.so: the C ABI declared in
 * include/wmbus_b200.h, device buffers, stream orchestration and the stream/batch
 * bookkeeping around the kernels in wmb_kernels.cuh.
 *
 * Per batch of IQ bytes (stream-ordered on the context's compute stream):
 *     H2D (copy stream, double-buffered)                       [host input only]
 *     K1  k1_demod_kernel        cu8 -> dphi (fp32) + rssi (u8), both chains
 *     K2a k2a_lanes_kernel       clock recovery lanes -> packed data bits + time2 strobes
 *         k2a_verify_kernel      compare lane start states with predecessors' end states;
 *                                refuted lanes are re-run until none is left
 *     K2t count / scan / write   time2 bit stream -> stream ring, access-code matches
 *     K2m k2m_lanes_kernel       run-length lanes (+ verify / re-run)
 *     K2c scan + compact         run-length events -> stream ring, access-code matches
 *     K3  size/offsets/copy      candidate frames -> pinned host memory
 * The host then owns 3-out-of-6 / Manchester / CRC (wmb_framer.c), exactly the split
 * BASELINE.json's north_star asks for.
 *
 * With -DWMB_HOSTSIM the same file builds against tests/hostsim/hostsim_cuda.h and runs
 * the kernels' phase functions on the CPU; that build is test infrastructure only.
 */
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <time.h>

#ifdef WMB_HOSTSIM
#include "hostsim_cuda.h"
#else
#include <cuda_runtime.h>
#endif

#include "wmbus_b200.h"
#include "wmb_framer.h"
#include "wmb_kernels.cuh"

#ifndef WMB_VERSION
#define WMB_VERSION "wmbus-b200 0.1 (sm_100a)"
#endif

static thread_local char g_err[512];

static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return set_err(WMB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                \
    } while (0)

/* --------------------------------------------------------------------------- */

struct Stream {                     /* one (chain, algo) bit stream */
    uint32_t *ev = nullptr;         /* run-length: lane-local events              */
    uint32_t *cnt = nullptr;        /* per-lane counts                            */
    uint64_t *base = nullptr;       /* per-lane ordinal base                      */
    uint64_t *ring = nullptr;       /* global event ring                          */
    uint64_t ring_cap = 0;          /* power of two                               */
    StreamDev *sd = nullptr;        /* device bookkeeping                         */
    uint64_t *cand = nullptr;       /* device: new access-code matches (ordinals, unordered) */
    uint64_t *pend = nullptr;       /* device: candidates waiting for more bits   */
    uint64_t *agg = nullptr;        /* scan scratch [tiles]                       */
    uint64_t total = 0;             /* host mirror of sd->total at the last read  */
    uint64_t total_prev = 0;        /* ... before the last batch (stage tap)      */
    /* host framer bookkeeping */
    int64_t busy_until = -1;        /* last ordinal consumed by an accepted packet */
};

/* Buffers that the demod / clock-recovery stage of batch i+1 writes while the bit-stream stage of batch i still
 * reads them: two sets, alternating from batch to batch.  Layout of the sample arrays: [W history | batch]. */
struct SetBuf {
    float *dphi = nullptr;          /* [W | M_max]  post-FIR discriminator output  */
    uint8_t *rssi = nullptr;        /* [W | M_max]                                 */
    uint32_t *dbits = nullptr;      /* [W/32 | M_max/32] data bits                 */
    uint32_t *sbits = nullptr;      /* [W/32 | M_max/32] time2 strobes             */
    uint32_t *cbits = nullptr;      /* [W/32 | M_max/32] clock signs (stage tap, opts.reserved[1] & 1) */
    IirState *ia_start = nullptr, *ia_end = nullptr;
    uint32_t *rerun_a = nullptr;
};

struct ChainBuf {
    SetBuf set[2];
    IirState *ia_carry = nullptr;
    RlState *rl_start = nullptr, *rl_end = nullptr, *rl_carry = nullptr;
    uint32_t *rerun = nullptr;
    uint32_t *lane_err = nullptr;   /* [lanes_max] error bits of the run-length lanes' verified runs */
    /* two-phase run-length path (T1/C1) */
    uint64_t *p1_rec = nullptr; uint32_t *p1_cnt = nullptr; uint64_t *p1_base = nullptr;
    P1State *p1_start = nullptr, *p1_end = nullptr; uint32_t *p1_rerun = nullptr;
    uint32_t *rec_m = nullptr, *rec_v = nullptr; uint16_t *rec_n = nullptr;
    uint32_t *p2_cnt = nullptr; uint64_t *p2_base = nullptr;
    K2pDev *pd = nullptr; RlState *p2_out = nullptr;
    /* time2 lanes */
    uint32_t *t2_tail = nullptr, *t2_len = nullptr, *t2_sr = nullptr, *t2_agg_tail = nullptr, *t2_agg_len = nullptr;
    Stream s[WMB_N_ALGOS];
};

#define WMB_NSLOT 4                      /* batches whose results may be waiting for the host */

static uint32_t g_p2_block = 128u;       /* threads per block of the phase-2 count pass (WMBUS_B200_P2BLK, experiments) */

struct QueuedLine {
    uint64_t end_sample;
    int prio;                       /* chain*2 + (algo == T2A) */
    wmb_decoded d;
    uint8_t algo;
};

struct wmb_ctx {
    wmb_opts o;
    int device = 0;
    uint32_t d = 2;                 /* effective decimation (>= 1) */
    uint32_t chains = 3;
    /* Streams: k1s runs the demod kernels of consecutive batches back to back; as[set] the clock-recovery lanes of
     * a batch (they only need its demod output, so they overlap the next batch's demod and each other); cs everything
     * that is sequential from batch to batch (lane verification, bit streams, gather, framer, result copies), with ts
     * (time2) and s2 (S1 run-length lanes) forked from and joined to it; xs the H2D copies. */
    cudaStream_t cs = nullptr, xs = nullptr, k1s = nullptr, as[2] = {nullptr, nullptr}, as2[2] = {nullptr, nullptr};
    cudaStream_t ts = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k1done[2] = {nullptr, nullptr};
    cudaEvent_t ev_k1[2] = {nullptr, nullptr}, ev_k2a[2] = {nullptr, nullptr}, ev_k2a2[2] = {nullptr, nullptr}, ev_chain[2] = {nullptr, nullptr};
    bool chain_recorded[2] = {false, false};
    cudaEvent_t ev_res[WMB_NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_push_start = nullptr;           /* first demod kernel of the current push (timers) */
    cudaEvent_t ev_reset = nullptr;                /* wmb_reset's device part (the first batch after it waits for it) */
    bool reset_pending = false;
    bool push_started = false;
    cudaEvent_t ev_t[WMB_NSLOT][6];                /* per result slot: demod start/end, bit sync start/end (timers) */
    bool allocated = false;

    /* geometry */
    size_t max_batch_bytes = 0;
    int64_t M_max = 0;
    uint32_t W = 32768;             /* retained history (decimated samples) = max warm-up */
    uint32_t W_a[WMB_N_CHAINS] = {24576, 81920};    /* warm-up of the clock-recovery lanes */
    uint32_t W_m[WMB_N_CHAINS] = {32768, 8192};    /* warm-up of the run-length lanes  */
    uint32_t C_fixed = 0;
    uint32_t lanes_max = 0;
    uint32_t t2_lanes_max = 0;
    uint32_t p1_lanes_max = 0, p2_lanes_max = 0;
    size_t rec_max = 0;
    bool two_phase = true;
    bool taps = false;              /* opts.reserved[1] & 1: keep the clock-sign words for wmb_debug_copy_bits */
    K2pDev *h_pd = nullptr;
    uint32_t cap_words_rl = 0;
    size_t ring_events = 0, ring_events_rl = 0;
    std::vector<void *> dev_allocs, host_allocs;
    uint8_t *d_tmp = nullptr;       /* scratch for the history slides                    */
    uint32_t cand_cap = 1u << 20;
    uint32_t frame_words_cap = 1u << 24;

    /* device buffers */
    uint8_t *d_in[2] = {nullptr, nullptr};
    uint8_t *d_hist = nullptr;
    float *d_lut = nullptr;
    ChainBuf cb[WMB_N_CHAINS];
    uint32_t *d_errors = nullptr, *d_nfail = nullptr;      /* [0] error bits; d_nfail[0..7]: refuted lanes per verified pass */
    GatherDev *d_gd = nullptr;      /* gather bookkeeping + statistics                 */
    /* results: WMB_NSLOT slots, one per batch in flight, each with its part of the candidate / verdict arrays and of
     * the datagram pool; the host mirrors are pinned and filled by copies enqueued right behind the device framer */
    BatchRec *d_rec = nullptr, *h_rec = nullptr;
    FrameHdr *d_hdr = nullptr, *h_hdr = nullptr;
    DecHdr *d_dec = nullptr, *h_dec = nullptr;
    uint8_t *d_pool = nullptr, *h_pool = nullptr;
    uint32_t slot_cap = 0, slot_pool = 0, pend_cap = 0;
    uint32_t *d_words = nullptr, *h_words = nullptr;
    uint32_t *d_cut_n = nullptr;
    uint64_t *d_k3_agg = nullptr;
    uint32_t spec_n = 0, spec_pool = 0;              /* entries / bytes copied before their counts are known */
    struct InFlight { int slot; bool final; bool has_timers; uint64_t m_end; };
    std::vector<InFlight> inflight;                  /* gathered batches whose results the host has not read yet */
    uint64_t stat_rerun_seen = 0, stat_fallback_seen = 0;
    double acc_demod_ms = 0, acc_bitsync_ms = 0, acc_pass_ms = 0;    /* timers of the current push */

    /* stream position */
    uint64_t iq_consumed = 0;       /* input IQ samples handed to the device    */
    uint64_t m_consumed = 0;        /* decimated samples produced               */
    int64_t hist_m = 0;             /* decimated history retained (<= W)        */
    int64_t hist_iq = 0;            /* input history retained (samples)         */
    std::vector<uint8_t> remainder; /* bytes not yet forming a whole batch granule */
    int buf_idx = 0;
    uint64_t batch_no = 0;          /* batches enqueued since create / reset: set = batch_no & 1 */
    uint64_t gather_no = 0;         /* gathers enqueued: result slot = gather_no % WMB_NSLOT    */
    int64_t last_M = 0, prev_M = 0;
    int last_set = 0;

    /* results */
    uint64_t win_lo = 0, win_hi = ~0ull;             /* line window (access-code match sample) */
    /* manual mode (opts.manual_frames): frames wait here for wmb_poll */
    struct Held { wmb_frame f; std::vector<uint32_t> words; };
    std::vector<Held> held, held_prev;
    std::vector<wmb_frame> poll_frames;
    bool manual = false;
    std::vector<QueuedLine> lines;
    wmb_stats st;
};

/* --------------------------------------------------------------------------- */
/* launches                                                                    */
/* --------------------------------------------------------------------------- */

static uint32_t *gd_field(wmb_ctx *c, size_t off) { return (uint32_t *)((uint8_t *)c->d_gd + off); }
#define GD_FIELD(c, f) gd_field(c, offsetof(GatherDev, f))

#ifdef WMB_HOSTSIM
#include "hostsim_launch.inl"
#else
static int g_k1_ctas = 0;                /* WMBUS_B200_K1_CTAS: resident demod blocks per SM (0: as many as fit) */

static int launch_k1(wmb_ctx *c, const K1Params &p, cudaStream_t st)
{
    const int64_t ntiles = (p.M + K1_TILE - 1) / K1_TILE;
    if (ntiles <= 0) return WMB_OK;
    static int sm_count = 0, blocks_per_sm = 0;
    const size_t smem = k1_smem_bytes(p.d, p.prefilter);
    if (!sm_count) {
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->device);
    }
    auto kern = p.prefilter ? (p.chains == 1u ? k1_demod_pre_kernel<1u> : p.chains == 2u ? k1_demod_pre_kernel<2u> : k1_demod_pre_kernel<3u>)
                            : (p.chains == 1u ? k1_demod_kernel<1u> : p.chains == 2u ? k1_demod_kernel<2u> : k1_demod_kernel<3u>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, K1_BLOCK, smem));
    if (blocks_per_sm < 1) return set_err(WMB_E_INVAL, "decimation %u needs %zu B shared memory per CTA", p.d, smem);
    if (g_k1_ctas > 0 && blocks_per_sm > g_k1_ctas) blocks_per_sm = g_k1_ctas;
    int64_t grid = (int64_t)sm_count * blocks_per_sm;      /* persistent: whole waves of resident CTAs */
    if (grid > ntiles) grid = ntiles;
    CUDA_TRY(cudaMemsetAsync(p.tile_ctr, 0, 4, st));
    kern<<<(unsigned)grid, K1_BLOCK, smem, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches++;
    return WMB_OK;
}

/* speculative pass (st: the batch's lane stream), then -- on cs, where batches follow each other in order --
 * verification + on-device fix-up of refuted lanes: no host round trip */
static int g_fixseg = 3;             /* WMBUS_B200_FIXSEG: parallel segment pass in front of the fix-up block for 1: clock lanes, 2: monolithic
                                            run-length lanes, 4: run-length phase 1 (off: its 2304-step lanes are cheap to re-run in one block, and the
                                            4096 idle blocks of the pass cost 0.1 ms per GiB) */
static bool g_k2a_coop = true;           /* WMBUS_B200_K2A=scalar: one lane per thread everywhere (experiments) */

static int launch_k2a_lanes(wmb_ctx *c, int chain, const K2aParams &p, cudaStream_t st)
{
    /* three threads per lane where the lane is the plain case (time2 on, no DC block, whole words); the per-thread
     * kernel otherwise */
    if (g_k2a_coop && p.t2 && !p.dc && p.M % 32 == 0 && p.W % 32 == 0 && p.C % 32 == 0 && p.hist % 32 == 0) {
        const unsigned per = (K2A2_THREADS / 32) * K2A2_LPW;
        const unsigned grid = (p.lanes + per - 1) / per;
        if (chain == 0) k2a2_lanes_kernel<ChainT1C1><<<grid, K2A2_THREADS, 0, st>>>(p);
        else            k2a2_lanes_kernel<ChainS1><<<grid, K2A2_THREADS, 0, st>>>(p);
        CUDA_TRY(cudaGetLastError());
        c->st.kernel_launches += 1;
        return WMB_OK;
    }
    const unsigned grid = (p.lanes + K2_THREADS - 1) / K2_THREADS;
    if (chain == 0) k2a_lanes_kernel<ChainT1C1><<<grid, K2_THREADS, 0, st>>>(p);
    else            k2a_lanes_kernel<ChainS1><<<grid, K2_THREADS, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k2a_verify(wmb_ctx *c, int chain, const K2aParams &p)
{
    uint32_t *nf = c->d_nfail + chain;
    k2a_verify_kernel<<<(p.lanes + 255) / 256, 256, 0, c->cs>>>(p, nf);
    const unsigned segs = (p.lanes + FIX_SEG - 1) / FIX_SEG;
    if (chain == 0) {
        if (g_fixseg & 1) k2a_fixseg_kernel<ChainT1C1><<<segs, FIX_SEG, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun));
        k2a_fixup_kernel<ChainT1C1><<<1, FIX_THREADS, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun), c->d_errors);
    } else {
        if (g_fixseg & 1) k2a_fixseg_kernel<ChainS1><<<segs, FIX_SEG, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun));
        k2a_fixup_kernel<ChainS1><<<1, FIX_THREADS, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun), c->d_errors);
    }
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 3;
    return WMB_OK;
}

static int launch_k2m(wmb_ctx *c, int chain, const K2mParams &p, cudaStream_t st)
{
    const unsigned grid = (p.lanes + K2_THREADS - 1) / K2_THREADS;
    uint32_t *nf = c->d_nfail + 2 + chain;
    if (chain == 0) {
        k2m_lanes_kernel<ChainT1C1><<<grid, K2_THREADS, 0, st>>>(p);
        k2m_verify_kernel<<<(p.lanes + 255) / 256, 256, 0, st>>>(p, nf);
        if (g_fixseg & 2) k2m_fixseg_kernel<ChainT1C1><<<(p.lanes + FIX_SEG - 1) / FIX_SEG, FIX_SEG, 0, st>>>(p, nf, GD_FIELD(c, lanes_rerun));
        k2m_fixup_kernel<ChainT1C1><<<1, FIX_THREADS, 0, st>>>(p, nf, GD_FIELD(c, lanes_rerun), c->d_errors);
    } else {
        k2m_lanes_kernel<ChainS1><<<grid, K2_THREADS, 0, st>>>(p);
        k2m_verify_kernel<<<(p.lanes + 255) / 256, 256, 0, st>>>(p, nf);
        if (g_fixseg & 2) k2m_fixseg_kernel<ChainS1><<<(p.lanes + FIX_SEG - 1) / FIX_SEG, FIX_SEG, 0, st>>>(p, nf, GD_FIELD(c, lanes_rerun));
        k2m_fixup_kernel<ChainS1><<<1, FIX_THREADS, 0, st>>>(p, nf, GD_FIELD(c, lanes_rerun), c->d_errors);
    }
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 4;
    return WMB_OK;
}

static int launch_k2p1(wmb_ctx *c, const K2p1Params &p)
{
    uint32_t *nf = c->d_nfail + 4;
    k2p1_lanes_kernel<<<(p.lanes + 127) / 128, 128, 0, c->cs>>>(p);
    k2p1_verify_kernel<<<(p.lanes + 255) / 256, 256, 0, c->cs>>>(p, nf);
    if (g_fixseg & 4) k2p1_fixseg_kernel<<<(p.lanes + FIX_SEG - 1) / FIX_SEG, FIX_SEG, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun));
    k2p1_fixup_kernel<<<1, FIX_THREADS, 0, c->cs>>>(p, nf, GD_FIELD(c, lanes_rerun), c->d_errors);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 4;
    return WMB_OK;
}

static void launch_cscan(wmb_ctx *c, const uint32_t *cnt, uint64_t *base, uint32_t n, uint64_t *agg, uint64_t *total,
                         const uint32_t *skip = nullptr, uint32_t *clear = nullptr, uint32_t from_zero = 0, cudaStream_t st = nullptr,
                         uint32_t skip_invert = 0)
{
    if (!st) st = c->cs;
    CountScan s;
    s.cnt = cnt; s.base = base; s.n = n; s.agg = agg; s.total = total; s.skip = skip; s.clear = clear; s.from_zero = from_zero;
    s.skip_invert = skip_invert;
    const unsigned tiles = scan_tiles(n);
    cscan_a_kernel<<<tiles, SCAN_BLOCK, 0, st>>>(s);
    cscan_b_kernel<<<1, 32, 0, st>>>(s);
    cscan_c_kernel<<<tiles, SCAN_BLOCK, 0, st>>>(s);
    c->st.kernel_launches += 3;
}

/* two-phase run-length path after phase 1: records -> phase 2 -> ring (the carried state is folded later) */
static int launch_k2p_rest(wmb_ctx *c, const K2pcParams &pc, K2p2Params p2)
{
    launch_cscan(c, pc.cnt, pc.base, pc.lanes, pc.agg, &pc.pd->n_rec, nullptr, &pc.pd->fallback, 1);
    k2pc_compact_kernel<<<pc.lanes, 128, 0, c->cs>>>(pc);
    k2p2_count_kernel<<<(p2.lanes + g_p2_block - 1) / g_p2_block, g_p2_block, 0, c->cs>>>(p2);
    k2p2_sum_kernel<<<p2.lanes, K2P2W_THREADS, 0, c->cs>>>(p2);
    launch_cscan(c, p2.cnt, p2.base, p2.lanes, p2.agg, &p2.sd->total, &p2.pd->fallback);
    k2p2_write_kernel<<<p2.lanes, K2P2W_THREADS, 0, c->cs>>>(p2);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 4;
    return WMB_OK;
}

static int launch_k2p_fold(wmb_ctx *c, const P1State *p1_end_last, RlState *p2_out, RlState *carry, const K2pDev *pd,
                           const RlState *mono_end)
{
    k2p_fold_kernel<<<1, 32, 0, c->cs>>>(p1_end_last, p2_out, carry, pd, mono_end, GD_FIELD(c, rl_fallbacks));
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k2m_carry(wmb_ctx *c, const RlState *end, RlState *carry, const uint32_t *run_if, cudaStream_t st)
{
    k2m_carry_kernel<<<1, 32, 0, st>>>(end, carry, run_if);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 1;
    return WMB_OK;
}

static int launch_k2t(wmb_ctx *c, int chain, const K2tParams &p)
{
    const unsigned grid = (p.lanes + 127) / 128, tiles = scan_tiles(p.lanes);
    if (chain == 0) {
        k2t_count_kernel<ChainT1C1><<<grid, 128, 0, c->ts>>>(p);
        t2scan_a_kernel<ChainT1C1><<<tiles, SCAN_BLOCK, 0, c->ts>>>(p);
        t2scan_b_kernel<ChainT1C1><<<1, 32, 0, c->ts>>>(p);
        t2scan_c_kernel<ChainT1C1><<<tiles, SCAN_BLOCK, 0, c->ts>>>(p);
        k2t_write_kernel<ChainT1C1><<<grid, 128, 0, c->ts>>>(p);
    } else {
        k2t_count_kernel<ChainS1><<<grid, 128, 0, c->ts>>>(p);
        t2scan_a_kernel<ChainS1><<<tiles, SCAN_BLOCK, 0, c->ts>>>(p);
        t2scan_b_kernel<ChainS1><<<1, 32, 0, c->ts>>>(p);
        t2scan_c_kernel<ChainS1><<<tiles, SCAN_BLOCK, 0, c->ts>>>(p);
        k2t_write_kernel<ChainS1><<<grid, 128, 0, c->ts>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 5;
    return WMB_OK;
}

static int launch_k2c(wmb_ctx *c, const K2cParams &p, cudaStream_t st)
{
    launch_cscan(c, p.cnt, p.base, p.lanes, p.agg, &p.sd->total, p.run_if, nullptr, 0, st, p.run_if ? 1u : 0u);
    k2c_compact_kernel<<<p.lanes, 128, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 1;
    return WMB_OK;
}

/* the whole gather + device framer; grids are fixed (the kernels loop over however many candidates there are) */
static int launch_k3_k4(wmb_ctx *c, const K3Params &p, const K4Params *q)
{
    static int sms = 0;
    if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    k3_plan_kernel<<<1, 32, 0, c->cs>>>(p);
    k3_fill_kernel<<<sms * 2, 256, 0, c->cs>>>(p);
    k3_size_kernel<<<sms, 128, 0, c->cs>>>(p);
    k3_cut_kernel<<<sms * 32, 64, 0, c->cs>>>(p);
    k3_offsets_kernel<<<1, SCAN_THREADS, 0, c->cs>>>(p);
    k3_copy_kernel<<<sms * 32, 128, 0, c->cs>>>(p);
    k3_carry_kernel<<<sms, 128, 0, c->cs>>>(p);
    c->st.kernel_launches += 8;
    if (q) {
        k4_decode_kernel<<<sms * 64, K4_THREADS, 0, c->cs>>>(*q);
        c->st.kernel_launches += 1;
    }
    k3_publish_kernel<<<1, 32, 0, c->cs>>>(p);
    CUDA_TRY(cudaGetLastError());
    return WMB_OK;
}

static int launch_k4(wmb_ctx *c, const K4Params &p)
{
    k4_decode_kernel<<<p.n ? p.n : 1, K4_THREADS, 0, c->cs>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->st.kernel_launches += 1;
    return WMB_OK;
}
#endif

/* --------------------------------------------------------------------------- */
/* set-up                                                                      */
/* --------------------------------------------------------------------------- */

extern "C" void wmb_default_opts(wmb_opts *o)
{
    memset(o, 0, sizeof(*o));
    o->decimation = 2;          /* rtl_wmbus.c:857 */
    o->accurate_atan = 1;       /* :859 */
    o->rla_enabled = 1;         /* :855 */
    o->t2_enabled = 1;          /* :856 */
    o->t1c1_enabled = 1;        /* :862 */
    o->s1_enabled = 1;          /* :863 */
}

extern "C" int wmb_abi_version(void) { return WMB_ABI_VERSION; }
extern "C" const char *wmb_last_error(void) { return g_err; }
extern "C" const char *wmb_version_string(void) { return WMB_VERSION; }

extern "C" void *wmb_host_alloc(size_t nbytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, nbytes ? nbytes : 1) != cudaSuccess) {
        set_err(WMB_E_NOMEM, "cannot allocate %zu bytes of pinned host memory", nbytes);
        return nullptr;
    }
    return p;
}

extern "C" void wmb_host_free(void *p) { if (p) cudaFreeHost(p); }

static uint64_t next_pow2(uint64_t v)
{
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename T>
static int dev_alloc(wmb_ctx *c, T **p, size_t count, bool zero = false)
{
    void *q = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    if (cudaMalloc(&q, bytes) != cudaSuccess) return set_err(WMB_E_NOMEM, "cudaMalloc of %zu bytes failed", bytes);
    if (zero && cudaMemset(q, 0, bytes) != cudaSuccess) return set_err(WMB_E_CUDA, "cudaMemset failed");
    c->dev_allocs.push_back(q);
    *p = (T *)q;
    return WMB_OK;
}

template <typename T>
static int host_alloc(wmb_ctx *c, T **p, size_t count)
{
    void *q = nullptr;
    if (cudaMallocHost(&q, std::max<size_t>(count * sizeof(T), 16)) != cudaSuccess)
        return set_err(WMB_E_NOMEM, "cudaMallocHost failed");
    c->host_allocs.push_back(q);
    *p = (T *)q;
    return WMB_OK;
}

#define TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

/* lane geometry of the bit-sync kernels; WMBUS_B200_TUNE="t2words:p1chunk:p2records" overrides it for experiments */
static uint32_t g_t2_words = 128u;       /* time2 lane = 32 * this many decimated samples */
static uint32_t g_p1_chunk = 2048u;      /* decimated samples per phase-1 run-length lane */
static uint32_t g_p2_records = 512u;     /* records per phase-2 lane (nominal) */
#define K2T_WORDS_PER_LANE g_t2_words
#define K2P1_CHUNK g_p1_chunk
#define K2P1_WARM  256u
#define K2P1_CAP   (K2P1_CHUNK / 5 + 2)
#define K2P2_RECORDS g_p2_records

static double wall_ms();
/* WMBUS_B200_TRACE=1: host wall-clock marks of one batch on stderr (debugging aid) */
static bool g_trace = false;
static std::vector<std::pair<const char *, double>> g_marks;
static inline void tr(const char *name) { if (g_trace) g_marks.emplace_back(name, wall_ms()); }
static void tr_dump()
{
    if (!g_trace || g_marks.empty()) return;
    fprintf(stderr, "[trace]");
    for (size_t i = 1; i < g_marks.size(); i++) fprintf(stderr, " %s %.3f", g_marks[i].first, g_marks[i].second - g_marks[i - 1].second);
    fprintf(stderr, " | total %.3f ms\n", g_marks.back().second - g_marks.front().second);
    g_marks.clear();
}

/* WMBUS_B200_PIPE_MIB: cut a long device push into batches of this size that follow each other through the device
 * like through a pipeline (run_batch).  Off by default: measured on the 1 GiB benchmark step, 4 x 256 MiB took 7.7 ms
 * of device time against 5.7 ms for one batch -- the demod kernel is a persistent grid that owns every SM's register
 * file, so the clock-recovery lanes of the batch before (128 registers x 64 threads per block, one serial chain per
 * thread) only get on an SM when demod blocks retire, and whatever issue slots they win stretch their critical path.
 * Host pushes are batched by the H2D copies anyway and go through the same machinery. */
static size_t g_pipe_bytes = ~(size_t)0;

static void read_tuning()
{
#ifndef WMB_HOSTSIM
    if (const char *k = getenv("WMBUS_B200_K2A")) g_k2a_coop = strcmp(k, "scalar") != 0;
    if (const char *k = getenv("WMBUS_B200_K1_CTAS")) g_k1_ctas = atoi(k);
    if (const char *k = getenv("WMBUS_B200_FIXSEG")) g_fixseg = atoi(k);
#endif
    if (const char *b = getenv("WMBUS_B200_PIPE_MIB")) { const unsigned long v = strtoul(b, nullptr, 10); if (v >= 1 && v <= 4096) g_pipe_bytes = (size_t)v << 20; }
    if (const char *b = getenv("WMBUS_B200_P2BLK")) { const unsigned v = (unsigned)atoi(b); if (v == 32 || v == 64 || v == 128) g_p2_block = v; }
    const char *e = getenv("WMBUS_B200_TUNE");
    unsigned a = 0, b = 0, r = 0;
    if (!e || sscanf(e, "%u:%u:%u", &a, &b, &r) != 3) return;
    if (a >= 4 && a <= 4096 && a % 4 == 0) g_t2_words = a;
    if (b >= 512 && b <= 65536 && b % 32 == 0) g_p1_chunk = b;
    if (r >= 32 && r <= K2P2W_THREADS * K2P2W_ITEMS) g_p2_records = r;
}

static int ctx_alloc(wmb_ctx *c)
{
    if (c->allocated) return WMB_OK;
    const uint32_t d = c->d;
    c->M_max = (int64_t)(c->max_batch_bytes / (2 * (size_t)d));
    const uint32_t C_min = c->C_fixed ? c->C_fixed : 8192;
    c->lanes_max = (uint32_t)(c->M_max / C_min + 2);
    c->t2_lanes_max = (uint32_t)(c->M_max / (32 * K2T_WORDS_PER_LANE) + 2);
    const size_t words_rl = (size_t)c->M_max / 4 + (size_t)c->lanes_max * (K2_EDGE_EMIT_CAP + 8) + 1024;
    c->cap_words_rl = (uint32_t)std::min<size_t>(words_rl, 0xFFFFFFFFu);
    c->ring_events = next_pow2((size_t)c->M_max / 4 + 65536 + WMB_MAXBITS);
    /* A run-length stream's ring: in the regime where the tracker's bit length has collapsed (an in-channel carrier) the
     * lanes emit up to their event buffers' capacity, 1.25 events per sample, and a burst of that is as long in a small
     * batch as in a large one.  Up to 2^26 events (128 MiB batches; the CLI's live hand-overs and its 64 MiB default) the
     * ring holds whatever the monolithic lanes of a batch can emit, so it cannot be overrun by them; larger batches keep
     * the size they were measured with, which is at least that */
    c->ring_events_rl = std::max<size_t>(c->ring_events, std::min<size_t>(next_pow2(words_rl + 65536 + WMB_MAXBITS), (size_t)1 << 26));
    if (c->o.reserved[1] & 2u) c->ring_events_rl = c->ring_events;      /* tests: reach the overrun path with a small capture */
    c->p1_lanes_max = (uint32_t)(c->M_max / K2P1_CHUNK + 2);
    c->rec_max = (size_t)c->M_max / 5 + 2 * (size_t)c->p1_lanes_max + 64;
    c->p2_lanes_max = (uint32_t)(c->rec_max / K2P2_RECORDS + 2);
    /* access-code matches and gathered frame bits scale with the batch: one candidate per 256 decimated samples, one
     * frame bit per 16 (dense traffic: a telegram every ~5000 samples; false matches: 2^-16 per bit) */
    c->cand_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(c->M_max >> 8, 1 << 16), 1 << 24);
    if (c->o.reserved[1] >> 8) c->cand_cap = std::max<uint32_t>(c->o.reserved[1] >> 8, 4u);   /* tests: force the overflow path */
    c->frame_words_cap = (uint32_t)std::min<int64_t>(std::max<int64_t>(c->M_max >> 4, 1 << 22), 1 << 28);

    TRY(dev_alloc(c, &c->d_in[0], c->max_batch_bytes + 4096));
    TRY(dev_alloc(c, &c->d_in[1], c->max_batch_bytes + 4096));
    TRY(dev_alloc(c, &c->d_hist, (size_t)k1_hist_bytes(d), true));
    TRY(dev_alloc(c, &c->d_tmp, std::max<size_t>((size_t)c->W * 4, (size_t)k1_hist_bytes(d))));
    TRY(dev_alloc(c, &c->d_lut, 2 * 4096));
    TRY(dev_alloc(c, &c->d_errors, 16, true));
    c->d_nfail = c->d_errors + 4;
    TRY(dev_alloc(c, &c->d_gd, 1, true));
    c->slot_cap = c->cand_cap;                   /* candidates of one batch */
    c->pend_cap = 1u << 16;                      /* candidates younger than one telegram at a batch end */
    c->slot_pool = c->frame_words_cap / 8 + 4 * c->cand_cap;    /* a datagram byte takes >= 8 shipped bit words */
    TRY(dev_alloc(c, &c->d_rec, WMB_NSLOT, true));
    TRY(dev_alloc(c, &c->d_hdr, (size_t)WMB_NSLOT * c->slot_cap));
    TRY(dev_alloc(c, &c->d_dec, (size_t)WMB_NSLOT * c->slot_cap));
    TRY(dev_alloc(c, &c->d_pool, (size_t)WMB_NSLOT * c->slot_pool));
    TRY(dev_alloc(c, &c->d_words, c->frame_words_cap));
    TRY(dev_alloc(c, &c->d_cut_n, c->cand_cap));
    TRY(dev_alloc(c, &c->d_k3_agg, SCAN_THREADS));
    TRY(host_alloc(c, &c->h_rec, WMB_NSLOT));
    TRY(host_alloc(c, &c->h_hdr, (size_t)WMB_NSLOT * c->slot_cap));
    TRY(host_alloc(c, &c->h_dec, (size_t)WMB_NSLOT * c->slot_cap));
    TRY(host_alloc(c, &c->h_pool, (size_t)WMB_NSLOT * c->slot_pool));
    TRY(host_alloc(c, &c->h_words, c->frame_words_cap));
    c->spec_n = std::min<uint32_t>(4096, c->slot_cap);          /* grows with what the batches actually produce (consume_oldest) */
    c->spec_pool = std::min<uint32_t>(1u << 18, c->slot_pool);

    /* mixer look-up tables, built with the host libm exactly like the reference
     * (setup_lookup_tables_for_frequency_translation, rtl_wmbus.c:974-993) */
    if (c->o.simultaneous) {
        const int fs_khz = (int)(c->o.decimation * 800u);
        const size_t n_max = (size_t)fs_khz / 25;
        if (n_max == 0 || n_max > 4096) return set_err(WMB_E_INVAL, "-s needs 1 <= decimation <= 128");
        std::vector<float> lut(2 * 4096, 0.f);
        for (size_t n = 0; n < n_max; n++) {
            const double phi = (2. * M_PI * (25 * n)) / fs_khz;
            lut[n] = cosf(phi);
            lut[4096 + n] = -sinf(phi);
        }
        CUDA_TRY(cudaMemcpy(c->d_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
    }

    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        if (!(c->chains & (1u << ch))) continue;
        ChainBuf &b = c->cb[ch];
        const size_t n = (size_t)c->W + (size_t)c->M_max + 512;    /* slack: block loads may run past M */
        for (int k = 0; k < 2; k++) {
            SetBuf &sb = b.set[k];
            TRY(dev_alloc(c, &sb.dphi, n));
            TRY(dev_alloc(c, &sb.rssi, n));
            TRY(dev_alloc(c, &sb.dbits, n / 32 + 64, true));      /* slack: lanes prefetch 16 words ahead */
            TRY(dev_alloc(c, &sb.sbits, n / 32 + 64, true));
            if (c->taps) TRY(dev_alloc(c, &sb.cbits, n / 32 + 64, true));
            TRY(dev_alloc(c, &sb.ia_start, c->lanes_max));
            TRY(dev_alloc(c, &sb.ia_end, c->lanes_max));
            TRY(dev_alloc(c, &sb.rerun_a, c->lanes_max, true));
        }
        TRY(dev_alloc(c, &b.ia_carry, 1));
        TRY(dev_alloc(c, &b.rl_start, c->lanes_max));
        TRY(dev_alloc(c, &b.rl_end, c->lanes_max));
        TRY(dev_alloc(c, &b.rl_carry, 1));
        TRY(dev_alloc(c, &b.rerun, c->lanes_max, true));
        TRY(dev_alloc(c, &b.lane_err, c->lanes_max, true));
        if (ch == 0 && c->two_phase && c->o.rla_enabled) {
            TRY(dev_alloc(c, &b.p1_rec, (size_t)c->p1_lanes_max * K2P1_CAP));
            TRY(dev_alloc(c, &b.p1_cnt, c->p1_lanes_max, true));
            TRY(dev_alloc(c, &b.p1_base, c->p1_lanes_max));
            TRY(dev_alloc(c, &b.p1_start, c->p1_lanes_max));
            TRY(dev_alloc(c, &b.p1_end, c->p1_lanes_max));
            TRY(dev_alloc(c, &b.p1_rerun, c->p1_lanes_max, true));
            TRY(dev_alloc(c, &b.rec_m, c->rec_max));
            TRY(dev_alloc(c, &b.rec_v, c->rec_max));
            TRY(dev_alloc(c, &b.rec_n, c->rec_max));
            TRY(dev_alloc(c, &b.p2_cnt, c->p2_lanes_max, true));
            TRY(dev_alloc(c, &b.p2_base, c->p2_lanes_max));
            TRY(dev_alloc(c, &b.pd, 1, true));
            TRY(dev_alloc(c, &b.p2_out, 1, true));
        }
        TRY(dev_alloc(c, &b.t2_tail, c->t2_lanes_max));
        TRY(dev_alloc(c, &b.t2_len, c->t2_lanes_max));
        TRY(dev_alloc(c, &b.t2_sr, c->t2_lanes_max));
        uint32_t scan_lanes = c->lanes_max > c->t2_lanes_max ? c->lanes_max : c->t2_lanes_max;
        if (c->p1_lanes_max > scan_lanes) scan_lanes = c->p1_lanes_max;
        if (c->p2_lanes_max > scan_lanes) scan_lanes = c->p2_lanes_max;
        const size_t n_agg = scan_tiles(scan_lanes) + 1;
        TRY(dev_alloc(c, &b.t2_agg_tail, n_agg));
        TRY(dev_alloc(c, &b.t2_agg_len, n_agg));
        IirState ia;
        iir_state_init(ia);
        CUDA_TRY(cudaMemcpy(b.ia_carry, &ia, sizeof(ia), cudaMemcpyHostToDevice));
        RlState rl;
        rl_state_init(rl, ch);
        CUDA_TRY(cudaMemcpy(b.rl_carry, &rl, sizeof(rl), cudaMemcpyHostToDevice));
        for (int a = 0; a < WMB_N_ALGOS; a++) {
            Stream &s = b.s[a];
            const uint32_t nl = a == WMB_ALGO_T2A ? c->t2_lanes_max : c->lanes_max;
            if (a == WMB_ALGO_RLA) TRY(dev_alloc(c, &s.ev, c->cap_words_rl));
            TRY(dev_alloc(c, &s.cnt, nl, true));
            TRY(dev_alloc(c, &s.base, nl));
            s.ring_cap = a == WMB_ALGO_RLA ? c->ring_events_rl : c->ring_events;
            TRY(dev_alloc(c, &s.ring, s.ring_cap));
            TRY(dev_alloc(c, &s.sd, 1, true));
            TRY(dev_alloc(c, &s.cand, c->cand_cap));
            TRY(dev_alloc(c, &s.pend, c->pend_cap));
            TRY(dev_alloc(c, &s.agg, n_agg));
        }
    }
    /* the set-up copies above are plain cudaMemcpy calls from pageable memory: make sure they have landed before any
     * kernel on the context's non-blocking streams can read them */
    CUDA_TRY(cudaDeviceSynchronize());
    c->allocated = true;
    return WMB_OK;
}

extern "C" int wmb_create(const wmb_opts *o, int cuda_device, wmb_ctx **out)
{
    if (!o || !out) return set_err(WMB_E_INVAL, "null argument");
    *out = nullptr;
    read_tuning();
    g_trace = getenv("WMBUS_B200_TRACE") != nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return set_err(WMB_E_NODEVICE, "no CUDA device available (libwmbus_b200 has no CPU fallback)");
    if (cuda_device < 0 || cuda_device >= ndev) return set_err(WMB_E_NODEVICE, "CUDA device %d of %d", cuda_device, ndev);
    /* the demod block stages one tile of raw samples (twice) and its converted words in shared memory: 8 bytes per input
     * sample of a 1024-row tile.  227 KB per block hold that up to decimation 25 (20 MS/s input; an RTL-SDR delivers 3.2) */
    if (o->decimation > WMB_MAX_DECIMATION)
        return set_err(WMB_E_INVAL, "decimation %u not supported (max %u: 20 MS/s input)", o->decimation, (unsigned)WMB_MAX_DECIMATION);
    if (o->simultaneous && o->decimation == 0) return set_err(WMB_E_INVAL, "-s with -d 0 is undefined in the reference");
    if (o->simultaneous > 2) return set_err(WMB_E_INVAL, "simultaneous: 0, 1 (-s) or 2 (explicit carriers)");
    if (o->prefilter > 4) return set_err(WMB_E_INVAL, "prefilter: 0 (moving averages), 1 (23-tap FIR), 2 (polyphase), 3 / 4 (their fixed-point twins)");
    if (o->prefilter && o->decimation != 2) return set_err(WMB_E_INVAL, "the pre-decimation low-passes are 1.6 MS/s designs: decimation must be 2");
    if (o->simultaneous == 2)
        for (int ch = 0; ch < 2; ch++)
            if (o->carrier_25khz[ch] == INT32_MIN || 2 * (o->carrier_25khz[ch] < 0 ? -o->carrier_25khz[ch] : o->carrier_25khz[ch]) > (int32_t)(o->decimation * 32u))
                return set_err(WMB_E_INVAL, "carrier offset outside the sampled band (|offset| <= fs / 2)");
    CUDA_TRY(cudaSetDevice(cuda_device));

    wmb_ctx *c = new wmb_ctx();
    c->o = *o;
    c->device = cuda_device;
    c->d = o->decimation ? o->decimation : 1;               /* rtl_wmbus.c:1350-1352: d==0 keeps every sample */
    c->chains = (o->t1c1_enabled ? 1u : 0u) | (o->s1_enabled ? 2u : 0u);
    /* warm-up (decimated samples) of the speculative lanes: the clock biquads need ~20 k
     * samples to re-join the true trajectory bit for bit, the DC block
     * another ~18 k in front of them (A.6); a run-length lane must reach back past the
     * start of the telegram it may begin in (A.8: <= 28 k samples T1, <= 113 k S1). */
    if (o->warmup_samples) {
        c->W_a[0] = c->W_a[1] = c->W_m[0] = c->W_m[1] = (o->warmup_samples + 255) / 256 * 256;
    } else {
        /* measured re-join times of the biquad state (tools/ + DESIGN.md): T1/C1 filter <= 19 k samples,
         * S1 filter (22-42 kHz band) <= 55 k; the DC block adds its own ~18 k in front */
        c->W_a[0] = o->remove_dc ? 98304u : 24576u;    /* measured re-join <= 19k samples; a miss only costs a re-run */
        c->W_a[1] = o->remove_dc ? 163840u : 81920u;   /* 0 / 32 / 96 re-runs of 0.3 M lanes at 81920 / 65536 / 57344 */
        /* run-length lanes.  T1/C1: the PI bit-length tracker remembers the whole reset-free stretch, so a cold
         * start re-joins at the first reset both trajectories share, at the latest when the telegram it started
         * in is over (<= 28 k samples).  S1: the state is the average run length of the last low and the last
         * high run plus 24 emitted bits, so it re-joins within ~30 runs; measured on the GPU: 0 / 0 / 64 / 248
         * re-runs of 0.3 M lanes at 16384 / 4096 / 2048 / 1024 samples. */
        c->W_m[0] = 32768u; c->W_m[1] = 8192u;
    }
    c->W = 0;
    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        if (!(c->chains & (1u << ch))) continue;
        c->W = std::max(c->W, c->W_a[ch]);
        if (o->rla_enabled) c->W = std::max(c->W, c->W_m[ch]);
    }
    c->manual = o->manual_frames != 0;
    c->taps = (o->reserved[1] & 1u) != 0;
    c->two_phase = o->reserved[0] == 0;                  /* reserved[0] = 1: force the monolithic run-length lanes (tests) */
    c->C_fixed = o->chunk_samples ? (o->chunk_samples + 255) / 256 * 256 : 0;
    if (c->C_fixed && (c->C_fixed < 1024 || c->C_fixed > K2_MAX_CHUNK)) { delete c; return set_err(WMB_E_INVAL, "chunk_samples out of range"); }
    size_t mb = o->max_batch_mib ? (size_t)o->max_batch_mib * 1048576u : (size_t)256 * 1048576u;
    const size_t gran = (size_t)4096 * c->d;
    mb = (mb + gran - 1) / gran * gran;
    c->max_batch_bytes = mb;
    memset(&c->st, 0, sizeof(c->st));
    if (cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->xs, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->k1s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->as[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->as[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->as2[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->as2[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->ts, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s2, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return set_err(WMB_E_CUDA, "cannot create CUDA streams");
    }
    for (int i = 0; i < 2; i++) {
        cudaEventCreate(&c->ev_h2d[i]); cudaEventCreate(&c->ev_k1done[i]);
        cudaEventCreateWithFlags(&c->ev_k1[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_k2a[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_chain[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_k2a2[i], cudaEventDisableTiming);
    }
    cudaEventCreate(&c->ev_push_start);
    cudaEventCreateWithFlags(&c->ev_reset, cudaEventDisableTiming);
    for (int i = 0; i < WMB_NSLOT; i++) {
        cudaEventCreateWithFlags(&c->ev_res[i], cudaEventDisableTiming);
        for (int k = 0; k < 6; k++) cudaEventCreate(&c->ev_t[i][k]);
    }
    cudaEventCreate(&c->ev_fork);
    cudaEventCreate(&c->ev_join);
    cudaEventCreate(&c->ev_fork2);
    cudaEventCreate(&c->ev_join2);
    *out = c;
    return WMB_OK;
}

extern "C" void wmb_destroy(wmb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->cs) cudaStreamSynchronize(c->cs);
    if (c->xs) cudaStreamSynchronize(c->xs);
    for (void *p : c->dev_allocs) cudaFree(p);
    for (void *p : c->host_allocs) cudaFreeHost(p);
    for (cudaStream_t st : { c->k1s, c->as[0], c->as[1], c->as2[0], c->as2[1], c->ts, c->s2 }) if (st) cudaStreamSynchronize(st);
    for (int i = 0; i < 2; i++) {
        for (cudaEvent_t e : { c->ev_h2d[i], c->ev_k1done[i], c->ev_k1[i], c->ev_k2a[i], c->ev_k2a2[i], c->ev_chain[i] }) if (e) cudaEventDestroy(e);
        if (c->as[i]) cudaStreamDestroy(c->as[i]);
        if (c->as2[i]) cudaStreamDestroy(c->as2[i]);
    }
    if (c->k1s) cudaStreamDestroy(c->k1s);
    if (c->ev_push_start) cudaEventDestroy(c->ev_push_start);
    if (c->ev_reset) cudaEventDestroy(c->ev_reset);
    for (int i = 0; i < WMB_NSLOT; i++) {
        if (c->ev_res[i]) cudaEventDestroy(c->ev_res[i]);
        for (int k = 0; k < 6; k++) if (c->ev_t[i][k]) cudaEventDestroy(c->ev_t[i][k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ts) cudaStreamDestroy(c->ts);
    if (c->ev_fork2) cudaEventDestroy(c->ev_fork2);
    if (c->ev_join2) cudaEventDestroy(c->ev_join2);
    if (c->s2) cudaStreamDestroy(c->s2);
    if (c->cs) cudaStreamDestroy(c->cs);
    if (c->xs) cudaStreamDestroy(c->xs);
    delete c;
}

/* --------------------------------------------------------------------------- */
/* one batch on the device                                                     */
/* --------------------------------------------------------------------------- */

/* The history prefix of a [W | batch] array of the set the next batch will use <- the last W elements the previous
 * batch left in its set (element size es, prev_M elements in the previous batch).  Source and destination are different
 * allocations, so one copy does. */
static int copy_history(void *dst, const void *src_set, size_t es, int64_t W, int64_t prev_M, cudaStream_t st)
{
    if (W <= 0) return WMB_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, (const uint8_t *)src_set + (size_t)prev_M * es, (size_t)W * es, cudaMemcpyDeviceToDevice, st));
    return WMB_OK;
}

static uint32_t pick_chunk(const wmb_ctx *c, int64_t M, bool alone)
{
    if (c->C_fixed) return c->C_fixed;
    /* The clock-recovery lanes are bound by the fp32 pipe of the scheduler they sit on, so a batch that has the GPU to
     * itself is cut into about one warp per scheduler (148 SMs x 4 x 32 lanes).  A batch that is followed by another
     * one overlaps its lanes with that batch's demod kernel: their latency is hidden, what counts is the redundant
     * warm-up arithmetic, so the lanes are made longer. */
    int64_t C = (M + 18943) / 18944;
    C = (C + 1023) / 1024 * 1024;
    /* A small batch on its own (a live stream's 100 ms hand-over) is all latency: its lanes take warm-up + C steps
     * however few they are, so they are made as short as the per-lane buffers allow (sized for lanes of 8192 samples at
     * the largest batch). */
    int64_t lo = alone ? 1024 : 16384;
    if (alone && c->lanes_max > 2) {
        const int64_t need = (M + c->lanes_max - 3) / (c->lanes_max - 2);
        lo = std::max<int64_t>(lo, (need + 1023) / 1024 * 1024);
    }
    if (C < lo) C = lo;
    if (C > 65536) C = 65536;
    return (uint32_t)C;
}

/* Lane length of the clock-recovery kernel when three threads share a lane (k2a2_lanes_kernel).  A lone warp takes
 * ~13 cycles per warm-up step and ~21 per live step (a second warp on the same scheduler doubles that: the kernel is
 * issue-bound with two), against 59 for the per-thread kernel -- so the batch is cut into ONE warp (ten lanes) per
 * scheduler, 148 SMs x 4 x 10 lanes, and the 24576-sample warm-up is paid 5920 times instead of 18944 times
 * (measured at 1 GiB: 1.31 ms with 1.5 warps per scheduler, see profiles/). */
static uint32_t pick_chunk_coop(const wmb_ctx *c, int64_t M)
{
    if (c->C_fixed) return c->C_fixed;
    int64_t C = (M + 5919) / 5920;
    C = (C + 255) / 256 * 256;
    if (C < 8192) C = 8192;
    if (C > 131072) C = 131072;
    return (uint32_t)C;
}

/* Enqueue the per-sample device pass for one batch whose bytes are at `src` (device memory): demod on k1s, clock
 * recovery lanes on as[set], everything that is sequential from batch to batch on cs.  Nothing here waits for the
 * device: refuted speculative lanes are re-run by on-device fix-up kernels, and the fallback from the two-phase
 * run-length path to the monolithic lanes is a set of kernels that do nothing unless the device flag asks for them.
 * input_ready: event after which `src` holds the bytes (H2D copy), or null.  alone: no batch follows in this push. */
static int run_batch(wmb_ctx *c, const uint8_t *src, size_t nbytes, cudaEvent_t input_ready, bool alone)
{
    tr("batch-start");
    const uint32_t d = c->d;
    const int64_t n_iq = (int64_t)(nbytes / 2);
    const int64_t M = n_iq / d;
    if (M <= 0) return WMB_OK;
    const int set = (int)(c->batch_no & 1u), pset = set ^ 1;
    const int slot = (int)(c->gather_no % WMB_NSLOT);           /* the gather that follows this batch */
    cudaEvent_t *evt = c->ev_t[slot];
    const bool first = c->batch_no == 0;
    /* a batch with nothing before it in flight and nothing behind it has nobody to overlap with: its demod kernel and
     * clock lanes go on cs too, which saves the cross-stream hand-overs (a few microseconds each) */
    const bool solo = alone && c->inflight.empty();
    cudaStream_t sk = solo ? c->cs : c->k1s, sa = solo ? c->cs : c->as[set];
    for (int ch = 0; ch < WMB_N_CHAINS; ch++)
        for (int a = 0; a < WMB_N_ALGOS; a++) c->cb[ch].s[a].total_prev = c->cb[ch].s[a].total;   /* stage tap: events of this batch */

    if (c->reset_pending) {                            /* the streams that do not follow cs wait for wmb_reset's kernel */
        for (cudaStream_t st : { c->k1s, c->as[0], c->as[1], c->as2[0], c->as2[1] }) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_reset, 0));
        c->reset_pending = false;
    }
    /* ================= stage 1 (k1s): history prefix of this set, demod ================= */
    if (c->chain_recorded[set]) CUDA_TRY(cudaStreamWaitEvent(sk, c->ev_chain[set], 0));    /* batch i-2 is done with the set */
    if (input_ready) CUDA_TRY(cudaStreamWaitEvent(sk, input_ready, 0));
    if (!first)
        for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
            if (!(c->chains & (1u << ch))) continue;
            ChainBuf &b = c->cb[ch];
            TRY(copy_history(b.set[set].dphi, b.set[pset].dphi, 4, c->W, c->prev_M, sk));
            TRY(copy_history(b.set[set].rssi, b.set[pset].rssi, 1, c->W, c->prev_M, sk));
        }
    CUDA_TRY(cudaEventRecord(evt[0], sk));
    if (!c->push_started) { CUDA_TRY(cudaEventRecord(c->ev_push_start, sk)); c->push_started = true; }
    K1Params k1;
    memset(&k1, 0, sizeof(k1));
    k1.in = src; k1.hist = c->d_hist; k1.in_bytes = (int64_t)nbytes;
    k1.n_hist_iq = c->hist_iq;
    k1.M = M; k1.d = d; k1.chains = c->chains;
    k1.accurate = c->o.accurate_atan; k1.mix = c->o.simultaneous ? 1u : 0u;
    k1.prefilter = c->o.prefilter;
    k1.lut_n = c->o.simultaneous ? (c->o.decimation * 800u) / 25u : 1u;
    k1.mix_k0 = (uint32_t)(c->iq_consumed % k1.lut_n);
    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        /* the reference's -s: T1/C1 chain 325 kHz above the centre, S1 chain 325 kHz below (rtl_wmbus.c:1008, :1025-1030) */
        const int32_t off = c->o.simultaneous == 2 ? c->o.carrier_25khz[ch] : (ch == 0 ? 13 : -13);
        const uint32_t mag = (uint32_t)(off < 0 ? -(int64_t)off : (int64_t)off);
        k1.mix_step[ch] = mag % k1.lut_n;
        k1.mix_conj[ch] = off < 0 ? 1u : 0u;
    }
    k1.lut_cos = c->d_lut; k1.lut_msin = c->d_lut + 4096;
    k1.tile_ctr = c->d_errors + 12;
    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        k1.dphi[ch] = c->cb[ch].set[set].dphi ? c->cb[ch].set[set].dphi + c->W : nullptr;
        k1.rssi[ch] = c->cb[ch].set[set].rssi ? c->cb[ch].set[set].rssi + c->W : nullptr;
    }
    TRY(launch_k1(c, k1, sk));
    CUDA_TRY(cudaEventRecord(evt[1], sk));
    CUDA_TRY(cudaEventRecord(c->ev_k1[set], sk));
    /* keep the last k1_hist_bytes() of the stream for the next batch's tile 0 */
    {
        const size_t hb = (size_t)k1_hist_bytes(d);
        if (nbytes >= hb) {
            CUDA_TRY(cudaMemcpyAsync(c->d_hist, src + nbytes - hb, hb, cudaMemcpyDeviceToDevice, sk));
        } else {
            CUDA_TRY(cudaMemcpyAsync(c->d_tmp, c->d_hist + nbytes, hb - nbytes, cudaMemcpyDeviceToDevice, sk));
            CUDA_TRY(cudaMemcpyAsync(c->d_tmp + (hb - nbytes), src, nbytes, cudaMemcpyDeviceToDevice, sk));
            CUDA_TRY(cudaMemcpyAsync(c->d_hist, c->d_tmp, hb, cudaMemcpyDeviceToDevice, sk));
        }
        c->hist_iq = std::min<int64_t>(c->hist_iq + n_iq, (int64_t)hb / 2);
    }
    /* the input buffer may be overwritten by the next H2D from here on */
    CUDA_TRY(cudaEventRecord(c->ev_k1done[c->buf_idx], sk));

    const uint32_t C = pick_chunk(c, M, alone);
    const uint32_t lanes = (uint32_t)((M + C - 1) / C);
    if (lanes > c->lanes_max) return set_err(WMB_E_INVAL, "internal: %u lanes > %u", lanes, c->lanes_max);
    const bool any_sync = c->o.rla_enabled || c->o.t2_enabled;
    const int64_t wofs = c->W / 32;                       /* word offset of batch sample 0 */

    if (any_sync) {
        /* ================= stage 2 (as[set]): clock-recovery lanes, every lane speculative ================= */
        K2aParams ka[WMB_N_CHAINS];
        const bool coop = c->o.t2_enabled && !c->o.remove_dc && M % 32 == 0;
        const uint32_t Ca = coop ? pick_chunk_coop(c, M) : C;
        const uint32_t lanes_a = (uint32_t)((M + Ca - 1) / Ca);
        /* the two chains' lanes side by side when three threads share a lane: one such warp leaves its scheduler half
         * idle (measured IPC 0.5), a second one from the other chain fills it.  (The per-thread kernel already runs at
         * 0.65: side by side it measured slower, 8.3 vs 7.2 ms of bit sync per GiB.) */
        const bool side = coop && c->chains == 3u;
        CUDA_TRY(cudaStreamWaitEvent(sa, c->ev_k1[set], 0));
        if (side) CUDA_TRY(cudaStreamWaitEvent(c->as2[set], c->ev_k1[set], 0));
        CUDA_TRY(cudaEventRecord(evt[4], sa));
        for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
            if (!(c->chains & (1u << ch))) continue;
            ChainBuf &b = c->cb[ch];
            SetBuf &sb = b.set[set];
            K2aParams &p = ka[ch];
            memset(&p, 0, sizeof(p));
            p.dphi = sb.dphi + c->W; p.M = M; p.hist = c->hist_m; p.C = Ca; p.W = c->W_a[ch]; p.lanes = lanes_a;
            p.dbits = sb.dbits + wofs; p.sbits = sb.sbits + wofs;
            p.cbits = sb.cbits ? sb.cbits + wofs : nullptr;
            p.st_start = sb.ia_start; p.st_end = sb.ia_end; p.carry = b.ia_carry; p.rerun = sb.rerun_a;
            p.dc = c->o.remove_dc; p.t2 = c->o.t2_enabled;
            p.mode = 0;
            p.spec0 = first ? 0u : 1u;                   /* the previous batch's lanes may still be running */
            c->st.lanes_run += lanes_a;
            TRY(launch_k2a_lanes(c, ch, p, (side && ch == 0) ? c->as2[set] : sa));     /* S1's lanes are the longer ones: timers on theirs */
        }
        if (side) {
            CUDA_TRY(cudaEventRecord(c->ev_k2a2[set], c->as2[set]));
            CUDA_TRY(cudaStreamWaitEvent(sa, c->ev_k2a2[set], 0));
        }
        CUDA_TRY(cudaEventRecord(evt[5], sa));
        CUDA_TRY(cudaEventRecord(c->ev_k2a[set], sa));
        tr("k1+k2a");

        /* ================= stage 3 (cs): in order from batch to batch ================= */
        CUDA_TRY(cudaStreamWaitEvent(c->cs, c->ev_k2a[set], 0));
        CUDA_TRY(cudaEventRecord(evt[2], c->cs));
        for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
            if (!(c->chains & (1u << ch))) continue;
            ChainBuf &b = c->cb[ch];
            TRY(launch_k2a_verify(c, ch, ka[ch]));
            CUDA_TRY(cudaMemcpyAsync(b.ia_carry, b.set[set].ia_end + (ka[ch].lanes - 1), sizeof(IirState), cudaMemcpyDeviceToDevice, c->cs));
            /* bit history for the run-length warm-ups: the previous batch's last W samples (exact since its fix-up) */
            if (!first && c->prev_M % 32 == 0) {         /* only a final (flush) batch can be ragged */
                TRY(copy_history(b.set[set].dbits, b.set[pset].dbits, 4, c->W / 32, c->prev_M / 32, c->cs));
                TRY(copy_history(b.set[set].sbits, b.set[pset].sbits, 4, c->W / 32, c->prev_M / 32, c->cs));
            }
        }

        /* chain 1's run-length lanes run on their own stream (forked after whatever cs holds, joined at the end) */
        auto fork2 = [&]() -> int {
            CUDA_TRY(cudaEventRecord(c->ev_fork2, c->cs));
            CUDA_TRY(cudaStreamWaitEvent(c->s2, c->ev_fork2, 0));
            return WMB_OK;
        };
        auto join2 = [&]() -> int {
            CUDA_TRY(cudaEventRecord(c->ev_join2, c->s2));
            CUDA_TRY(cudaStreamWaitEvent(c->cs, c->ev_join2, 0));
            return WMB_OK;
        };

        /* ---- K2t: time2 bit streams straight into the rings (own stream: independent of the
         *      run-length kernels below, and both leave most of the GPU idle on their own) ---- */
        if (c->o.t2_enabled) {
            CUDA_TRY(cudaEventRecord(c->ev_fork, c->cs));
            CUDA_TRY(cudaStreamWaitEvent(c->ts, c->ev_fork, 0));
            for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
                if (!(c->chains & (1u << ch))) continue;
                ChainBuf &b = c->cb[ch];
                SetBuf &sb = b.set[set];
                Stream &s = b.s[WMB_ALGO_T2A];
                K2tParams p;
                memset(&p, 0, sizeof(p));
                p.dbits = sb.dbits + wofs; p.sbits = sb.sbits + wofs; p.rssi = sb.rssi + c->W;
                p.M = M; p.Cw = K2T_WORDS_PER_LANE;
                const uint32_t nw = (uint32_t)((M + 31) / 32);
                p.lanes = (nw + p.Cw - 1) / p.Cw;
                if (p.lanes > c->t2_lanes_max) return set_err(WMB_E_INVAL, "internal: time2 lanes");
                p.cnt = s.cnt; p.tail = b.t2_tail; p.tail_len = b.t2_len; p.base = s.base; p.sr_start = b.t2_sr;
                p.agg_cnt = s.agg; p.agg_tail = b.t2_agg_tail; p.agg_len = b.t2_agg_len;
                p.m_base = (int64_t)c->m_consumed;
                p.ring = s.ring; p.ring_mask = s.ring_cap - 1; p.sd = s.sd; p.cand = s.cand; p.cand_cap = c->cand_cap;
                TRY(launch_k2t(c, ch, p));
            }
            CUDA_TRY(cudaEventRecord(c->ev_join, c->ts));
        }

        /* ---- run-length bit sync ---- */
        if (c->o.rla_enabled) {
            const bool two = (c->chains & 1u) && c->two_phase;
            /* chains that take the monolithic lanes unconditionally: S1 always, T1/C1 when forced (tests) */
            const uint32_t mono = (c->chains & 2u) | (((c->chains & 1u) && !two) ? 1u : 0u);
            K2mParams km[WMB_N_CHAINS];
            auto setup_mono = [&](int ch, const uint32_t *run_if) -> int {
                ChainBuf &b = c->cb[ch];
                SetBuf &sb = b.set[set];
                K2mParams &p = km[ch];
                memset(&p, 0, sizeof(p));
                p.dbits = sb.dbits + wofs; p.rssi = sb.rssi + c->W; p.M = M; p.hist = c->hist_m;
                p.C = C; p.W = c->W_m[ch]; p.lanes = lanes;
                p.cap = C / 4 + K2_EDGE_EMIT_CAP + 8;
                if ((uint64_t)lanes * p.cap > c->cap_words_rl) return set_err(WMB_E_INVAL, "internal: event buffers too small for C=%u", C);
                p.ev = b.s[WMB_ALGO_RLA].ev; p.cnt = b.s[WMB_ALGO_RLA].cnt;
                p.st_start = b.rl_start; p.st_end = b.rl_end; p.carry = b.rl_carry; p.rerun = b.rerun;
                p.errors = c->d_errors; p.lane_err = b.lane_err;
                p.mode = 0; p.run_if = run_if;
                if (!run_if) c->st.lanes_run += lanes;
                return WMB_OK;
            };
            auto compact_mono = [&](int ch, cudaStream_t st, const uint32_t *run_if) -> int {       /* lane events -> ring */
                ChainBuf &b = c->cb[ch];
                Stream &s = b.s[WMB_ALGO_RLA];
                K2cParams q;
                memset(&q, 0, sizeof(q));
                q.ev = s.ev; q.cnt = s.cnt; q.base = s.base; q.lanes = lanes; q.cap = km[ch].cap; q.C = C;
                q.m_base = (int64_t)c->m_consumed;
                q.ring = s.ring; q.ring_mask = s.ring_cap - 1; q.sd = s.sd; q.cand = s.cand; q.cand_cap = c->cand_cap;
                q.agg = s.agg; q.rssi = b.set[set].rssi + c->W; q.run_if = run_if;
                q.lane_err = b.lane_err; q.errors = c->d_errors;
                return launch_k2c(c, q, st);
            };
            /* S1 (and a forced T1/C1) beside the two-phase path of T1/C1 */
            const bool s1_beside = two && (mono & 2u);
            if (s1_beside) TRY(fork2());
            for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
                if (!(mono & (1u << ch))) continue;
                cudaStream_t st = (s1_beside && ch == 1) ? c->s2 : c->cs;
                TRY(setup_mono(ch, nullptr));
                TRY(launch_k2m(c, ch, km[ch], st));
                TRY(launch_k2m_carry(c, c->cb[ch].rl_end + (lanes - 1), c->cb[ch].rl_carry, nullptr, st));
                TRY(compact_mono(ch, st, nullptr));
            }
            if (two) {
                /* T1/C1: phase 1 (per-sample, verified) -> records -> phase 2 (per-run) */
                ChainBuf &b = c->cb[0];
                Stream &s = b.s[WMB_ALGO_RLA];
                K2p1Params p1;
                memset(&p1, 0, sizeof(p1));
                p1.dbits = b.set[set].dbits + wofs; p1.M = M; p1.hist = c->hist_m;
                p1.C = K2P1_CHUNK; p1.W = K2P1_WARM; p1.lanes = (uint32_t)((M + K2P1_CHUNK - 1) / K2P1_CHUNK);
                p1.cap = K2P1_CAP; p1.rec = b.p1_rec; p1.cnt = b.p1_cnt;
                p1.st_start = b.p1_start; p1.st_end = b.p1_end; p1.carry = b.rl_carry; p1.rerun = b.p1_rerun;
                p1.mode = 0;
                if (p1.lanes > c->p1_lanes_max) return set_err(WMB_E_INVAL, "internal: phase-1 lanes");
                c->st.lanes_run += p1.lanes;
                TRY(launch_k2p1(c, p1));
                K2pcParams pc;
                memset(&pc, 0, sizeof(pc));
                pc.rec = b.p1_rec; pc.cnt = b.p1_cnt; pc.base = b.p1_base; pc.lanes = p1.lanes; pc.cap = p1.cap; pc.C = p1.C;
                pc.rec_m = b.rec_m; pc.rec_v = b.rec_v; pc.agg = s.agg; pc.pd = b.pd;
                K2p2Params p2;
                memset(&p2, 0, sizeof(p2));
                p2.rec_m = b.rec_m; p2.rec_v = b.rec_v; p2.rec_n = b.rec_n; p2.pd = b.pd; p2.R = K2P2_RECORDS;
                p2.lanes = (uint32_t)(((uint64_t)M / 5 + 2 * (uint64_t)p1.lanes) / K2P2_RECORDS + 2);
                if (p2.lanes > c->p2_lanes_max) return set_err(WMB_E_INVAL, "internal: phase-2 lanes");
                p2.cnt = b.p2_cnt; p2.base = b.p2_base; p2.rssi = b.set[set].rssi + c->W; p2.m_base = (int64_t)c->m_consumed;
                p2.ring = s.ring; p2.ring_mask = s.ring_cap - 1; p2.sd = s.sd; p2.cand = s.cand; p2.cand_cap = c->cand_cap;
                p2.carry = b.rl_carry; p2.p2_out = b.p2_out; p2.agg = s.agg;
                TRY(launch_k2p_rest(c, pc, p2));
                /* the second reset rule (rtl_wmbus.c:756-762) fired somewhere in this batch (pd->fallback, set by phase
                 * 2, which then wrote nothing): redo T1/C1 with the exact monolithic lanes.  The kernels are always
                 * enqueued; without the flag every thread returns at once. */
                const uint32_t *flag = &b.pd->fallback;
                TRY(setup_mono(0, flag));
                TRY(launch_k2m(c, 0, km[0], c->cs));
                TRY(compact_mono(0, c->cs, flag));
                TRY(launch_k2p_fold(c, b.p1_end + (p1.lanes - 1), b.p2_out, b.rl_carry, b.pd, b.rl_end + (lanes - 1)));
            }
            if (s1_beside) TRY(join2());
            tr("k2t+p1+p2");
        }
        if (c->o.t2_enabled) CUDA_TRY(cudaStreamWaitEvent(c->cs, c->ev_join, 0));
        CUDA_TRY(cudaEventRecord(evt[3], c->cs));
    } else {
        CUDA_TRY(cudaStreamWaitEvent(c->cs, c->ev_k1[set], 0));
        CUDA_TRY(cudaEventRecord(evt[2], c->cs));
        CUDA_TRY(cudaEventRecord(evt[3], c->cs));
    }

    c->hist_m = std::min<int64_t>(c->hist_m + M, c->W);
    c->iq_consumed += (uint64_t)n_iq;
    c->m_consumed += (uint64_t)M;
    c->st.input_samples += (uint64_t)n_iq;
    c->st.decimated_samples += (uint64_t)M;
    c->st.batches++;
    c->batch_no++;
    c->prev_M = M; c->last_M = M; c->last_set = set;
    return WMB_OK;
}

static int consume_oldest(wmb_ctx *c);

/* Enqueue the frame gather (K3), the device framer (K4) and the copies of their results into the host mirror of the
 * next result slot, for everything the streams hold: the candidates carried over plus the new access-code matches.
 * after_batch: this gather closes the batch just enqueued (its set may be reused once it is done).  final: end of
 * input, nothing is carried over. */
static int enqueue_gather(wmb_ctx *c, bool final, bool after_batch)
{
    if (!c->allocated) return WMB_OK;
    const bool any_sync = c->o.rla_enabled || c->o.t2_enabled;
    if (c->inflight.size() >= WMB_NSLOT) TRY(consume_oldest(c));          /* the slot's host mirror must be free */
    const int slot = (int)(c->gather_no % WMB_NSLOT);
    const size_t lb = (size_t)slot * c->slot_cap, pb = (size_t)slot * c->slot_pool;
    if (any_sync) {
        K3Params p;
        memset(&p, 0, sizeof(p));
        for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
            if (!(c->chains & (1u << ch))) continue;
            for (int a = 0; a < WMB_N_ALGOS; a++) {
                if ((a == WMB_ALGO_RLA && !c->o.rla_enabled) || (a == WMB_ALGO_T2A && !c->o.t2_enabled)) continue;
                Stream &s = c->cb[ch].s[a];
                const int k = ch * WMB_N_ALGOS + a;
                p.ring[k] = s.ring; p.ring_mask[k] = s.ring_cap - 1; p.sd[k] = s.sd; p.cand[k] = s.cand; p.pend[k] = s.pend;
            }
        }
        p.pend_cap = c->pend_cap; p.cand_cap = c->cand_cap;
        p.gd = c->d_gd; p.rec = c->d_rec + slot;
        p.hdr_log = c->d_hdr; p.dec_log = c->d_dec; p.log_base = (uint32_t)lb; p.log_cap = c->slot_cap;
        p.words = c->d_words; p.words_cap = c->frame_words_cap;
        p.cut_n = c->d_cut_n; p.agg = c->d_k3_agg; p.errors = c->d_errors;
        p.final = final ? 1u : 0u;
        K4Params q;
        memset(&q, 0, sizeof(q));
        q.hdr = c->d_hdr; q.words = c->d_words; q.dec = c->d_dec;
        q.pool = c->d_pool + pb; q.pool_cap = c->slot_pool; q.pool_n = GD_FIELD(c, pool_n); q.errors = c->d_errors; q.gd = c->d_gd;
        TRY(launch_k3_k4(c, p, c->manual ? nullptr : &q));
        /* results -> pinned host mirror: the record and a prefix of the arrays it describes (the rest, if a batch ever
         * produces more, is fetched when the record has been read) */
        CUDA_TRY(cudaMemcpyAsync(c->h_rec + slot, c->d_rec + slot, sizeof(BatchRec), cudaMemcpyDeviceToHost, c->cs));
        CUDA_TRY(cudaMemcpyAsync(c->h_hdr + lb, c->d_hdr + lb, (size_t)c->spec_n * sizeof(FrameHdr), cudaMemcpyDeviceToHost, c->cs));
        if (!c->manual) {
            CUDA_TRY(cudaMemcpyAsync(c->h_dec + lb, c->d_dec + lb, (size_t)c->spec_n * sizeof(DecHdr), cudaMemcpyDeviceToHost, c->cs));
            CUDA_TRY(cudaMemcpyAsync(c->h_pool + pb, c->d_pool + pb, c->spec_pool, cudaMemcpyDeviceToHost, c->cs));
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev_res[slot], c->cs));
    if (after_batch) {
        CUDA_TRY(cudaEventRecord(c->ev_chain[c->last_set], c->cs));
        c->chain_recorded[c->last_set] = true;
    }
    wmb_ctx::InFlight f;
    f.slot = slot; f.final = final; f.has_timers = after_batch; f.m_end = c->m_consumed;
    c->inflight.push_back(f);
    c->gather_no++;
    return WMB_OK;
}

static int book_device_frames(wmb_ctx *c, const FrameHdr *hdr, const DecHdr *dec, const uint8_t *pool, size_t n, bool final);

/* Wait for the oldest gathered batch's results (an event, not a stream: later batches keep running), fetch what the
 * prefix copy did not cover, and run the stream-order bookkeeping over it. */
static int consume_oldest(wmb_ctx *c)
{
    if (c->inflight.empty()) return WMB_OK;
    const wmb_ctx::InFlight f = c->inflight.front();
    c->inflight.erase(c->inflight.begin());
    const bool any_sync = c->o.rla_enabled || c->o.t2_enabled;
    const double t0 = wall_ms();
    CUDA_TRY(cudaEventSynchronize(c->ev_res[f.slot]));
    const double t1 = wall_ms();
    c->st.host_gather_ms += t1 - t0;
    if (f.has_timers) {
        float ms = 0.f;
        cudaEvent_t *evt = c->ev_t[f.slot];
        const bool any = c->o.rla_enabled || c->o.t2_enabled;
        if (cudaEventElapsedTime(&ms, evt[0], evt[1]) == cudaSuccess) c->acc_demod_ms += ms;
        if (any && cudaEventElapsedTime(&ms, evt[4], evt[5]) == cudaSuccess) c->acc_bitsync_ms += ms;     /* clock-recovery lanes */
        if (cudaEventElapsedTime(&ms, evt[2], evt[3]) == cudaSuccess) c->acc_bitsync_ms += ms;            /* bit streams */
        /* the whole per-sample pass of the push so far: first demod kernel -> this batch's last bit-sync kernel */
        if (cudaEventElapsedTime(&ms, c->ev_push_start, evt[3]) == cudaSuccess) c->acc_pass_ms = ms;
    }
    if (!any_sync) return WMB_OK;
    const bool dev_decode = !c->manual;
    const size_t lb = (size_t)f.slot * c->slot_cap, pb = (size_t)f.slot * c->slot_pool;
    const BatchRec r = c->h_rec[f.slot];
    const uint32_t err = r.errors;
    if (err & 2u) return set_err(WMB_E_OVERFLOW, "run-length tracker left its defined range (the reference would spin here)");
    /* lane event buffer (1: a run-length lane emitted more than one bit per four samples plus one capped edge -- the
     * tracker's bit length has collapsed to a fraction of a sample), frame words (4), datagram pool (8), access-code
     * matches (16), pending candidates (64): the device dropped what did not fit and cleared the flags; the reference
     * would have gone on decoding, so does the stream */
    if (err & K3_SOFT_ERRORS) c->st.overflow_batches++;
    if (err & 256u) return set_err(WMB_E_STATE, "internal: lane verification does not converge");
    bool more = false;
    if (r.n > c->spec_n) {
        CUDA_TRY(cudaMemcpyAsync(c->h_hdr + lb + c->spec_n, c->d_hdr + lb + c->spec_n, (size_t)(r.n - c->spec_n) * sizeof(FrameHdr), cudaMemcpyDeviceToHost, c->xs));
        if (dev_decode) CUDA_TRY(cudaMemcpyAsync(c->h_dec + lb + c->spec_n, c->d_dec + lb + c->spec_n, (size_t)(r.n - c->spec_n) * sizeof(DecHdr), cudaMemcpyDeviceToHost, c->xs));
        more = true;
    }
    if (dev_decode && r.pool_n > c->spec_pool) {
        CUDA_TRY(cudaMemcpyAsync(c->h_pool + pb + c->spec_pool, c->d_pool + pb + c->spec_pool, r.pool_n - c->spec_pool, cudaMemcpyDeviceToHost, c->xs));
        more = true;
    }
    if (!dev_decode && r.n_words) {          /* manual mode reads after every batch: the frame words are this batch's */
        CUDA_TRY(cudaMemcpyAsync(c->h_words, c->d_words, (size_t)r.n_words * 4, cudaMemcpyDeviceToHost, c->xs));
        c->st.d2h_bytes += (uint64_t)r.n_words * 4;
        more = true;
    }
    if (more) CUDA_TRY(cudaStreamSynchronize(c->xs));       /* the slot is not written again before it is consumed */
    /* the next batches probably look like this one: let the prefix copy cover them */
    if (r.n > c->spec_n) c->spec_n = std::min<uint32_t>(c->slot_cap, r.n + r.n / 4 + 256);
    if (r.pool_n > c->spec_pool) c->spec_pool = std::min<uint32_t>(c->slot_pool, r.pool_n + r.pool_n / 4 + 4096);
    c->st.d2h_bytes += sizeof(BatchRec) + (size_t)r.n * sizeof(FrameHdr) + (dev_decode ? (size_t)r.n * sizeof(DecHdr) + r.pool_n : 0);
    /* statistics kept on the device */
    c->st.lanes_rerun += r.lanes_rerun - c->stat_rerun_seen; c->st.lanes_run += r.lanes_rerun - c->stat_rerun_seen;
    c->stat_rerun_seen = r.lanes_rerun;
    c->st.rl_fallbacks += r.rl_fallbacks - c->stat_fallback_seen;
    c->stat_fallback_seen = r.rl_fallbacks;
    for (int ch = 0; ch < WMB_N_CHAINS; ch++)
        for (int a = 0; a < WMB_N_ALGOS; a++) {
            const int k = ch * WMB_N_ALGOS + a;
            c->st.candidates[ch][a] = r.n_cand_total[k];
            c->cb[ch].s[a].total = r.total[k];
        }
    if (r.n > c->slot_cap) return set_err(WMB_E_STATE, "internal: batch record beyond its slot");
    FrameHdr *hdr = c->h_hdr + lb;
    /* the device keeps 40 bits of the sample index; widen to the 64-bit stream position: the newest value
     * congruent to it that is not beyond the samples produced when the batch was gathered */
    for (uint32_t i = 0; i < r.n; i++) hdr[i].sync_sample = f.m_end - ((f.m_end - hdr[i].sync_sample) & EVG_M_MASK);
    int rc = WMB_OK;
    if (dev_decode) rc = book_device_frames(c, hdr, c->h_dec + lb, c->h_pool + pb, r.n, f.final);
    else {
        /* manual mode: keep the frames (newest version of a re-delivered partial one wins) for wmb_poll */
        for (uint32_t i = 0; i < r.n; i++) {
            const FrameHdr &h = hdr[i];
            if (h.nbits == 0) continue;
            wmb_frame fr;
            memset(&fr, 0, sizeof(fr));
            fr.sync_sample = h.sync_sample; fr.ordinal = h.ordinal; fr.chain = h.chain; fr.algo = h.algo;
            fr.truncated = (uint8_t)((h.complete && !h.cut) ? 0 : 1);
            fr.reserved = (uint8_t)((!h.complete && !f.final) ? 1 : 0);      /* partial: will be re-delivered */
            fr.nbits = h.nbits;
            wmb_ctx::Held *slot = nullptr;
            for (auto &hh : c->held)
                if (hh.f.chain == fr.chain && hh.f.algo == fr.algo && hh.f.ordinal == fr.ordinal) { slot = &hh; break; }
            if (!slot) { c->held.emplace_back(); slot = &c->held.back(); }
            slot->f = fr;
            slot->words.assign(c->h_words + h.word_off, c->h_words + h.word_off + h.nbits);
        }
    }
    c->st.host_decode_ms += wall_ms() - t1;
    return rc;
}

static int consume_all(wmb_ctx *c)
{
    while (!c->inflight.empty()) TRY(consume_oldest(c));
    c->st.demod_kernel_ms = c->acc_demod_ms; c->st.bitsync_kernel_ms = c->acc_bitsync_ms; c->st.batch_device_ms = c->acc_pass_ms;
    tr("consumed");
    tr_dump();
    return WMB_OK;
}

/* --------------------------------------------------------------------------- */
/* push / poll                                                                 */
/* --------------------------------------------------------------------------- */

static int process_device_batches(wmb_ctx *c, const uint8_t *dev, size_t nbytes, bool final);

static double wall_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

/* gather what the batch just enqueued produced; manual mode reads it at once (the frame words are per batch) */
static int finish_batch(wmb_ctx *c, bool final, bool after_batch)
{
    TRY(enqueue_gather(c, final, after_batch));
    if (c->manual) TRY(consume_all(c));
    return WMB_OK;
}

static size_t batch_granule(const wmb_ctx *c) { return (size_t)4096 * c->d; }

extern "C" int wmb_push_device(wmb_ctx *c, const void *dev_cu8, size_t nbytes)
{
    if (!c || (!dev_cu8 && nbytes)) return set_err(WMB_E_INVAL, "null argument");
    if (nbytes % 4096) return set_err(WMB_E_INVAL, "wmb_push_device needs a multiple of 4096 bytes");
    if (((uintptr_t)dev_cu8) & 15u) return set_err(WMB_E_INVAL, "device buffer must be 16-byte aligned");
    if (!c->remainder.empty()) return set_err(WMB_E_STATE, "wmb_push_device after a partial push");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ctx_alloc(c);
    if (rc) return rc;
    /* whole decimation granules (4096 * d bytes) go to the device as they are; a trailing partial granule (a capture
     * whose length is a multiple of 4096 but not of 4096 * d) waits on the host side like the remainder of a host push */
    const size_t tail = nbytes % batch_granule(c);
    rc = process_device_batches(c, (const uint8_t *)dev_cu8, nbytes - tail, false);
    if (rc) return rc;
    rc = consume_all(c);                                  /* also: the caller's buffer is no longer in use */
    if (rc || !tail) return rc;
    c->remainder.resize(tail);
    CUDA_TRY(cudaMemcpy(c->remainder.data(), (const uint8_t *)dev_cu8 + (nbytes - tail), tail, cudaMemcpyDeviceToHost));
    return WMB_OK;
}

/* dev points to device memory that stays valid until the results have been consumed.  A long push is cut into
 * batches that follow each other through the device like through a pipeline: the demod kernel of batch i+1 runs
 * beside the latency-bound bit-stream kernels of batch i (run_batch), and the host books batch i's results while
 * later batches are still running. */
static int process_device_batches(wmb_ctx *c, const uint8_t *dev, size_t nbytes, bool final)
{
    const size_t gran = batch_granule(c);
    size_t cap = std::min(c->max_batch_bytes, g_pipe_bytes);
    cap -= cap % gran;
    if (!cap) cap = gran;
    size_t off = 0;
    c->acc_demod_ms = c->acc_bitsync_ms = c->acc_pass_ms = 0; c->push_started = false;
    while (off < nbytes) {
        size_t n = std::min(nbytes - off, cap);
        if (nbytes - off - n < cap / 4) n = nbytes - off;                /* no dwarf batch at the end */
        if (n > c->max_batch_bytes) n = c->max_batch_bytes;
        if (n < nbytes - off || !final) {
            /* keep batch boundaries on whole decimation periods */
            if (n % gran) n -= n % gran;
            if (n == 0) break;
        }
        const double tb = wall_ms();
        int rc = run_batch(c, dev + off, n, nullptr, off + n >= nbytes);
        if (rc) return rc;
        c->st.host_batch_ms += wall_ms() - tb;
        rc = finish_batch(c, false, true);
        if (rc) return rc;
        off += n;
    }
    if (off < nbytes) return set_err(WMB_E_INVAL, "device push must be a multiple of %zu bytes unless flushing", gran);
    return WMB_OK;
}

/* host bytes -> device (double-buffered) -> batches */
static int push_host_bytes(wmb_ctx *c, const uint8_t *p, size_t nbytes, bool final)
{
    const size_t gran = batch_granule(c);
    size_t off = 0;
    /* enqueue the first H2D, then for each batch: enqueue the next H2D before computing */
    size_t cur_n = 0;
    /* Batch sizes: the maximum while plenty is left, then a taper (7/16 of what is left, not below half a
     * batch): the copy of batch i+1 hides behind the kernels of batch i, so what a caller waits for after the
     * last byte has crossed PCIe is the LAST batch's kernels -- a smaller last batch shortens that, as long as
     * every batch still computes faster than the next one copies (fixed lane warm-ups: about 1.6 ms + 5.6 us
     * per MiB against 18.8 us per MiB of copy). */
    auto next_size = [&](size_t at) {
        const size_t rem = nbytes - at, mx = c->max_batch_bytes, half = mx / 2;
        size_t n;
        if (rem > 2 * mx) n = mx;
        else if (rem <= half + half / 2) n = rem;
        else if (rem <= 2 * half) n = rem - half;
        else n = std::min(mx, std::max(half, rem / 16 * 7));
        if (!(final && at + n == nbytes)) n -= n % gran;
        if (n == 0 && rem >= gran) n = gran;
        return n;
    };
    cur_n = next_size(0);
    if (cur_n == 0) {
        c->remainder.assign(p, p + nbytes);             /* less than one granule: keep for later */
        return WMB_OK;
    }
    c->acc_demod_ms = c->acc_bitsync_ms = c->acc_pass_ms = 0; c->push_started = false;
    int idx = c->buf_idx;
    CUDA_TRY(cudaStreamWaitEvent(c->xs, c->ev_k1done[idx], 0));
    CUDA_TRY(cudaMemcpyAsync(c->d_in[idx], p, cur_n, cudaMemcpyHostToDevice, c->xs));
    CUDA_TRY(cudaEventRecord(c->ev_h2d[idx], c->xs));
    while (cur_n) {
        const size_t nxt_off = off + cur_n;
        const size_t nxt_n = nxt_off < nbytes ? next_size(nxt_off) : 0;
        if (nxt_n) {
            const int nidx = idx ^ 1;
            CUDA_TRY(cudaStreamWaitEvent(c->xs, c->ev_k1done[nidx], 0));
            CUDA_TRY(cudaMemcpyAsync(c->d_in[nidx], p + nxt_off, nxt_n, cudaMemcpyHostToDevice, c->xs));
            CUDA_TRY(cudaEventRecord(c->ev_h2d[nidx], c->xs));
        }
        c->buf_idx = idx;
        const double tb = wall_ms();
        int rc = run_batch(c, c->d_in[idx], cur_n, c->ev_h2d[idx], nxt_n == 0);
        if (rc) return rc;
        c->st.host_batch_ms += wall_ms() - tb;
        c->st.h2d_bytes += cur_n;
        rc = finish_batch(c, false, true);
        if (rc) return rc;
        off = nxt_off; cur_n = nxt_n; idx ^= 1;
    }
    c->buf_idx = idx;
    CUDA_TRY(cudaStreamSynchronize(c->xs));             /* the caller may reuse its buffer now */
    if (off < nbytes) c->remainder.assign(p + off, p + nbytes);
    return consume_all(c);
}

extern "C" int wmb_push(wmb_ctx *c, const uint8_t *cu8, size_t nbytes)
{
    if (!c || (!cu8 && nbytes)) return set_err(WMB_E_INVAL, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ctx_alloc(c);
    if (rc) return rc;
    if (!c->remainder.empty()) {
        /* complete the pending granule first (remainder is always shorter than one granule) */
        const size_t gran = batch_granule(c);
        const size_t take = std::min(gran - c->remainder.size(), nbytes);
        c->remainder.insert(c->remainder.end(), cu8, cu8 + take);
        cu8 += take; nbytes -= take;
        if (c->remainder.size() < gran) return WMB_OK;
        std::vector<uint8_t> tmp;
        tmp.swap(c->remainder);
        rc = push_host_bytes(c, tmp.data(), tmp.size(), false);
        if (rc) return rc;
        if (!nbytes) return WMB_OK;
    }
    return push_host_bytes(c, cu8, nbytes, false);
}

/* end of input: process what is left in whole 4096-byte items (rtl_wmbus.c:1301-1308) */
static int flush_input(wmb_ctx *c)
{
    int rc = ctx_alloc(c);
    if (rc) return rc;
    if (!c->remainder.empty()) {
        std::vector<uint8_t> tmp;
        tmp.swap(c->remainder);
        const size_t n = tmp.size() - tmp.size() % 4096;
        if (n) {
            rc = push_host_bytes(c, tmp.data(), n, true);
            if (rc) return rc;
            c->remainder.clear();
        }
    }
    rc = finish_batch(c, true, false);
    if (rc) return rc;
    return consume_all(c);
}

extern "C" int wmb_poll(wmb_ctx *c, wmb_frame *out, size_t cap, size_t *n, int flush)
{
    if (!c || !n) return set_err(WMB_E_INVAL, "null argument");
    *n = 0;
    CUDA_TRY(cudaSetDevice(c->device));
    if (flush) {
        int rc = flush_input(c);
        if (rc) return rc;
    }
    if (out) {
        c->poll_frames.clear();
        for (auto &h : c->held) { h.f.bits = h.words.data(); c->poll_frames.push_back(h.f); }
        const size_t k = std::min(cap, c->poll_frames.size());
        memcpy(out, c->poll_frames.data(), k * sizeof(wmb_frame));
        *n = k;
        /* the words stay valid until the next push/poll; forget the frames themselves */
        c->held_prev.swap(c->held);
        c->held.clear();
    }
    return WMB_OK;
}

/* --------------------------------------------------------------------------- */
/* host framers: stream-order bookkeeping around wmb_frame_decode()            */
/* --------------------------------------------------------------------------- */

/* Stream-order bookkeeping over decoded candidates (sorted by chain, algorithm, ordinal): a decoder
 * that is receiving ignores further access-code matches (t1_c1_packet_decoder.h:272-278 honours the
 * flag only in idle), so a candidate inside the telegram of an earlier one is dropped; the rest
 * become lines, queued in the order the reference prints them. */
struct FrameMeta { uint8_t chain, algo, partial, truncated; uint64_t ordinal, sync_sample; };
struct DecLite { int status; uint32_t consumed; uint64_t end_sample; uint8_t crc_ok; };

template <class Meta, class Lite, class Fill>
static int book_frames(wmb_ctx *c, size_t n, Meta meta, Lite lite, Fill fill)
{
    struct Key { uint64_t end_sample; uint32_t prio, seq; size_t fi; uint8_t algo; };
    bool blocked[WMB_N_CHAINS][WMB_N_ALGOS] = {{false, false}, {false, false}};
    std::vector<Key> fresh;
    for (size_t fi = 0; fi < n; fi++) {
        const FrameMeta f = meta(fi);
        Stream &s = c->cb[f.chain].s[f.algo];
        if (blocked[f.chain][f.algo]) continue;
        if ((int64_t)f.ordinal <= s.busy_until) continue;
        const DecLite d = lite(fi);
        if (d.status == WMB_DEC_NEED_MORE) {
            if (f.partial) { blocked[f.chain][f.algo] = true; continue; }   /* comes again */
            if (!f.truncated) return set_err(WMB_E_STATE, "internal: frame shorter than its header demands");
            /* cut by a run-length reset or by the end of input: the reference's decoder is reset too */
            s.busy_until = (int64_t)(f.ordinal + d.consumed - 1);
            continue;
        }
        s.busy_until = (int64_t)(f.ordinal + d.consumed - 1);
        if (d.status == WMB_DEC_LINE && f.sync_sample >= c->win_lo && f.sync_sample < c->win_hi) {
            Key k;
            k.end_sample = d.end_sample;
            k.prio = (uint32_t)(f.chain * 2 + (f.algo == WMB_ALGO_T2A ? 1 : 0));
            k.seq = (uint32_t)fresh.size(); k.fi = fi; k.algo = f.algo;
            fresh.push_back(k);
            c->st.lines[f.chain][f.algo]++;
            if (d.crc_ok) c->st.lines_crc_ok[f.chain][f.algo]++;
        }
    }
    /* the reference prints in the order the per-sample state machines finish:
     * sample index, then T1/C1-rla, T1/C1-t2a, S1-rla, S1-t2a (rtl_wmbus.c:1354-1355).
     * Only the small keys are sorted; the datagrams are materialised once, in print order. */
    auto before = [](const Key &a, const Key &b) {
        if (a.end_sample != b.end_sample) return a.end_sample < b.end_sample;
        if (a.prio != b.prio) return a.prio < b.prio;
        return a.seq < b.seq;
    };
    /* the candidates come stream by stream and, inside a stream, in bit order: the accepted ones of a stream do not
     * overlap, so each stream's lines already are in print order -- merge the (at most four) runs instead of sorting
     * (15 k lines per GiB on dense traffic: 1 ms of std::sort) */
    {
        std::vector<size_t> cut(1, 0);
        bool runs_sorted = true;
        for (size_t i = 1; i < fresh.size(); i++) {
            if (fresh[i].prio != fresh[i - 1].prio) cut.push_back(i);
            else if (before(fresh[i], fresh[i - 1])) runs_sorted = false;
        }
        cut.push_back(fresh.size());
        if (!runs_sorted || cut.size() > 6) std::sort(fresh.begin(), fresh.end(), before);
        else
            for (size_t r = 2; r < cut.size(); r++)
                std::inplace_merge(fresh.begin(), fresh.begin() + (long)cut[r - 1], fresh.begin() + (long)cut[r], before);
    }
    const size_t base = c->lines.size();
    c->lines.resize(base + fresh.size());
    for (size_t i = 0; i < fresh.size(); i++) {
        QueuedLine &q = c->lines[base + i];
        q.end_sample = fresh[i].end_sample; q.prio = (int)fresh[i].prio; q.algo = fresh[i].algo;
        fill(fresh[i].fi, q.d);
    }
    return WMB_OK;
}

/* candidates of one gathered batch, decoded by K4 (already in stream order) */
static int book_device_frames(wmb_ctx *c, const FrameHdr *hdr, const DecHdr *dec, const uint8_t *pool, size_t n, bool final)
{
    static const char modes[3][3] = { "T1", "C1", "S1" };
    /* frames without any bit (candidate at the very end of the stream) are not decoded at all */
    std::vector<uint32_t> idx;
    idx.reserve(n);
    for (size_t i = 0; i < n; i++) if (hdr[i].nbits) idx.push_back((uint32_t)i);
    return book_frames(c, idx.size(),
        [&](size_t k) {
            const FrameHdr &h = hdr[idx[k]];
            FrameMeta m;
            m.chain = h.chain; m.algo = h.algo; m.ordinal = h.ordinal; m.sync_sample = h.sync_sample;
            m.partial = (uint8_t)((!h.complete && !final) ? 1 : 0);
            m.truncated = (uint8_t)((h.complete && !h.cut) ? 0 : 1);
            return m;
        },
        [&](size_t k) {
            const DecHdr &d = dec[idx[k]];
            DecLite l;
            l.status = d.status; l.consumed = d.consumed; l.end_sample = hdr[idx[k]].sync_sample + d.end_off; l.crc_ok = d.crc_ok;
            return l;
        },
        [&](size_t k, wmb_decoded &o) {
            const DecHdr &d = dec[idx[k]];
            memset(&o, 0, sizeof(o));
            o.status = WMB_DEC_LINE; o.consumed = d.consumed; o.end_sample = hdr[idx[k]].sync_sample + d.end_off;
            memcpy(o.mode, modes[d.mode < 3 ? d.mode : 0], 3);
            o.crc_ok = d.crc_ok; o.ok_3of6 = d.ok_3of6; o.packet_rssi = d.packet_rssi; o.current_rssi = d.current_rssi;
            o.serial = d.serial; o.len = d.len;
            memcpy(o.datagram, pool + d.data_off, d.len);
        });
}

/* Test hook (declared in wmb_framer.h, not part of the public ABI): run K4 on caller-made frames so
 * that the device framer can be compared with its host twin candidate by candidate. */
extern "C" int wmb_frame_decode_device(wmb_ctx *c, const wmb_frame *frames, size_t n, wmb_decoded *out)
{
    if (!c || !frames || !out) return set_err(WMB_E_INVAL, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(ctx_alloc(c));
    if (n > c->cand_cap) return set_err(WMB_E_INVAL, "too many frames");
    size_t words = 0;
    for (size_t i = 0; i < n; i++) {
        FrameHdr &h = c->h_hdr[i];
        memset(&h, 0, sizeof(h));
        h.ordinal = frames[i].ordinal; h.sync_sample = frames[i].sync_sample; h.nbits = frames[i].nbits;
        h.word_off = (uint32_t)words; h.chain = frames[i].chain; h.algo = frames[i].algo; h.complete = 1;
        if (words + h.nbits > c->frame_words_cap) return set_err(WMB_E_INVAL, "frames too large");
        memcpy(c->h_words + words, frames[i].bits, (size_t)h.nbits * 4);
        words += h.nbits;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_hdr, c->h_hdr, n * sizeof(FrameHdr), cudaMemcpyHostToDevice, c->cs));
    CUDA_TRY(cudaMemcpyAsync(c->d_words, c->h_words, words * 4, cudaMemcpyHostToDevice, c->cs));
    if (!c->inflight.empty()) return set_err(WMB_E_STATE, "unread results");
    CUDA_TRY(cudaMemsetAsync(GD_FIELD(c, pool_n), 0, 4, c->cs));
    K4Params q;
    memset(&q, 0, sizeof(q));
    q.hdr = c->d_hdr; q.n = (uint32_t)n; q.words = c->d_words; q.dec = c->d_dec;
    q.pool = c->d_pool; q.pool_cap = c->slot_pool; q.pool_n = GD_FIELD(c, pool_n); q.errors = c->d_errors;
    TRY(launch_k4(c, q));
    CUDA_TRY(cudaMemcpyAsync(c->h_dec, c->d_dec, n * sizeof(DecHdr), cudaMemcpyDeviceToHost, c->cs));
    CUDA_TRY(cudaMemcpyAsync(c->h_pool, c->d_pool, c->slot_pool < (1u << 24) ? c->slot_pool : (1u << 24), cudaMemcpyDeviceToHost, c->cs));
    CUDA_TRY(cudaMemsetAsync(GD_FIELD(c, pool_n), 0, 4, c->cs));
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    static const char modes[3][3] = { "T1", "C1", "S1" };
    for (size_t i = 0; i < n; i++) {
        const DecHdr &d = c->h_dec[i];
        wmb_decoded &o = out[i];
        memset(&o, 0, sizeof(o));
        o.status = d.status == K4_SKIP ? (int)WMB_DEC_NEED_MORE : (int)d.status;
        o.consumed = d.consumed;
        o.end_sample = d.status == K4_SKIP ? 0 : frames[i].sync_sample + d.end_off;
        if (d.status != K4_LINE) continue;
        memcpy(o.mode, modes[d.mode < 3 ? d.mode : 0], 3);
        o.crc_ok = d.crc_ok; o.ok_3of6 = d.ok_3of6; o.packet_rssi = d.packet_rssi; o.current_rssi = d.current_rssi;
        o.serial = d.serial; o.len = d.len;
        memcpy(o.datagram, c->h_pool + d.data_off, d.len);
    }
    return WMB_OK;
}

extern "C" int wmb_decode_frames(wmb_ctx *c, const wmb_frame *frames, size_t n)
{
    if (!c || (!frames && n)) return set_err(WMB_E_INVAL, "null argument");
    std::vector<const wmb_frame *> v(n);
    for (size_t i = 0; i < n; i++) v[i] = &frames[i];
    std::sort(v.begin(), v.end(), [](const wmb_frame *a, const wmb_frame *b) {
        if (a->chain != b->chain) return a->chain < b->chain;
        if (a->algo != b->algo) return a->algo < b->algo;
        return a->ordinal < b->ordinal;
    });
    for (const wmb_frame *f : v)
        if (f->chain >= WMB_N_CHAINS || f->algo >= WMB_N_ALGOS) return set_err(WMB_E_INVAL, "bad frame");
    /* frames handed in by the caller are decoded by the host twin of K4 (wmb_framer.c); the
     * per-frame decode is pure, so it is spread over a few host threads */
    std::vector<wmb_decoded> dec(n);
    {
        unsigned nt = n >= 256 ? std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
        auto work = [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) wmb_frame_decode(v[i], &dec[i]); };
        if (nt <= 1) work(0, n);
        else {
            std::vector<std::thread> th;
            const size_t per = (n + nt - 1) / nt;
            for (unsigned t = 0; t < nt; t++) {
                const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
                if (lo < hi) th.emplace_back(work, lo, hi);
            }
            for (auto &t : th) t.join();
        }
    }
    return book_frames(c, n,
        [&](size_t i) {
            FrameMeta m;
            m.chain = v[i]->chain; m.algo = v[i]->algo; m.ordinal = v[i]->ordinal; m.sync_sample = v[i]->sync_sample;
            m.partial = v[i]->reserved; m.truncated = v[i]->truncated;
            return m;
        },
        [&](size_t i) {
            DecLite l;
            l.status = dec[i].status; l.consumed = dec[i].consumed; l.end_sample = dec[i].end_sample; l.crc_ok = dec[i].crc_ok;
            return l;
        },
        [&](size_t i, wmb_decoded &o) { o = dec[i]; });
}

extern "C" size_t wmb_take_lines(wmb_ctx *c, char *buf, size_t cap, size_t *n_lines, int timestamp_mode)
{
    size_t len = 0, taken = 0;
    if (n_lines) *n_lines = 0;
    if (!c || !buf) return 0;
    char ts[64];
    for (; taken < c->lines.size(); taken++) {
        const QueuedLine &q = c->lines[taken];
        if (timestamp_mode == 1) snprintf(ts, sizeof(ts), "TS");
        else if (timestamp_mode == 2) snprintf(ts, sizeof(ts), "@%014llu.%d", (unsigned long long)q.end_sample, q.prio);
        else wmb_make_time_string(ts, sizeof(ts));
        const char *prefix = c->o.show_algorithm ? (q.algo == WMB_ALGO_RLA ? "rla;" : "t2a;") : "";
        char line[1024];
        const size_t l = wmb_format_line(&q.d, prefix, ts, line, sizeof(line));
        if (len + l + 1 > cap) break;
        memcpy(buf + len, line, l);
        len += l;
    }
    if (len < cap) buf[len] = 0;
    c->lines.erase(c->lines.begin(), c->lines.begin() + (long)taken);
    if (n_lines) *n_lines = taken;
    return len;
}

extern "C" long wmb_process(wmb_ctx *c, const uint8_t *cu8, size_t nbytes, int flush,
                            char *out, size_t outcap, size_t *n_lines, int timestamp_mode)
{
    int rc = wmb_push(c, cu8, nbytes);
    if (rc) return rc;
    if (flush) {
        CUDA_TRY(cudaSetDevice(c->device));
        rc = flush_input(c);
        if (rc) return rc;
    }
    return (long)wmb_take_lines(c, out, outcap, n_lines, timestamp_mode);
}

extern "C" long wmb_process_device(wmb_ctx *c, const void *dev_cu8, size_t nbytes, int flush,
                                   char *out, size_t outcap, size_t *n_lines, int timestamp_mode)
{
    if (!c) return set_err(WMB_E_INVAL, "null argument");
    if (nbytes % 4096) return set_err(WMB_E_INVAL, "need a multiple of 4096 bytes");
    if (((uintptr_t)dev_cu8) & 15u) return set_err(WMB_E_INVAL, "device buffer must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ctx_alloc(c);
    if (rc) return rc;
    tr("enter");
    rc = process_device_batches(c, (const uint8_t *)dev_cu8, nbytes, flush != 0);
    if (rc) return rc;
    tr("batches");
    if (flush) rc = flush_input(c);
    else rc = consume_all(c);
    if (rc) return rc;
    tr("flush");
    const long len = (long)wmb_take_lines(c, out, outcap, n_lines, timestamp_mode);
    tr("lines");
    tr_dump();
    return len;
}

/* Back to the state right after wmb_create (a new capture starts): filter memories, stream
 * position, pending candidates, queued lines.  Buffers stay allocated. */
extern "C" int wmb_reset(wmb_ctx *c)
{
    if (!c) return set_err(WMB_E_INVAL, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    /* whatever was enqueued is abandoned: let it finish (a completed push has left the streams idle) */
    for (cudaStream_t st : { c->cs, c->xs, c->k1s, c->as[0], c->as[1], c->as2[0], c->as2[1], c->ts, c->s2 })
        if (st && cudaStreamQuery(st) != cudaSuccess) CUDA_TRY(cudaStreamSynchronize(st));
    c->iq_consumed = 0; c->m_consumed = 0; c->hist_m = 0; c->hist_iq = 0;
    c->remainder.clear(); c->lines.clear(); c->held.clear(); c->held_prev.clear();
    c->batch_no = 0; c->last_M = 0; c->prev_M = 0; c->last_set = 0; c->inflight.clear();
    c->chain_recorded[0] = c->chain_recorded[1] = false;
    c->stat_rerun_seen = 0; c->stat_fallback_seen = 0;
    for (int ch = 0; ch < WMB_N_CHAINS; ch++)
        for (int a = 0; a < WMB_N_ALGOS; a++) { Stream &s = c->cb[ch].s[a]; s.total = 0; s.total_prev = 0; s.busy_until = -1; }
    if (c->allocated) {
        /* device state back to the start of a stream, in stream order on cs: one small kernel, no host copy, no wait;
         * the first batch's kernels on the other streams wait for it (ev_reset) */
        ResetParams r;
        memset(&r, 0, sizeof(r));
        for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
            if (!(c->chains & (1u << ch))) continue;
            r.ia_carry[ch] = c->cb[ch].ia_carry; r.rl_carry[ch] = c->cb[ch].rl_carry;
            for (int a = 0; a < WMB_N_ALGOS; a++) r.sd[ch * WMB_N_ALGOS + a] = c->cb[ch].s[a].sd;
        }
        r.gd = c->d_gd; r.errors = c->d_errors;
#ifdef WMB_HOSTSIM
        wmb_reset_device(r);
#else
        wmb_reset_kernel<<<1, 32, 0, c->cs>>>(r);
        CUDA_TRY(cudaGetLastError());
#endif
        CUDA_TRY(cudaEventRecord(c->ev_reset, c->cs));
        c->reset_pending = true;
    }
    return WMB_OK;
}

extern "C" int wmb_seek(wmb_ctx *c, uint64_t first_iq_sample)
{
    if (!c) return set_err(WMB_E_INVAL, "null argument");
    if (first_iq_sample % (2048ull * c->d)) return set_err(WMB_E_INVAL, "seek position must be a multiple of 2048 * decimation IQ samples");
    int rc = wmb_reset(c);
    if (rc) return rc;
    c->iq_consumed = first_iq_sample;
    c->m_consumed = first_iq_sample / c->d;
    return WMB_OK;
}

extern "C" int wmb_set_line_window(wmb_ctx *c, uint64_t sync_lo, uint64_t sync_hi)
{
    if (!c || sync_lo > sync_hi) return set_err(WMB_E_INVAL, "bad window");
    c->win_lo = sync_lo; c->win_hi = sync_hi;
    return WMB_OK;
}

extern "C" long wmb_boundary_state(wmb_ctx *c, uint8_t *buf, size_t cap)
{
    if (!c || !buf) return set_err(WMB_E_INVAL, "null argument");
    if (!c->remainder.empty()) return set_err(WMB_E_STATE, "boundary state needs whole batch granules");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ctx_alloc(c);
    if (rc) return rc;
    rc = consume_all(c);                                /* every enqueued batch is gathered and booked first */
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    GatherDev gd;
    CUDA_TRY(cudaMemcpy(&gd, c->d_gd, sizeof(gd), cudaMemcpyDeviceToHost));
    std::vector<uint8_t> out;
    auto put = [&](const void *p, size_t n) { const uint8_t *b = (const uint8_t *)p; out.insert(out.end(), b, b + n); };
    const uint64_t pos[2] = { c->iq_consumed, c->m_consumed };
    put(pos, sizeof(pos));
    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        if (!(c->chains & (1u << ch))) continue;
        ChainBuf &b = c->cb[ch];
        IirState ia;
        RlState rl;
        CUDA_TRY(cudaMemcpy(&ia, b.ia_carry, sizeof(ia), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(&rl, b.rl_carry, sizeof(rl), cudaMemcpyDeviceToHost));
        if (!c->o.remove_dc) { ia.dc_x = 0.f; ia.dc_y = 0.f; }        /* unused without -o */
        if (!c->o.t2_enabled) memset(&ia, 0, sizeof(ia));
        if (!c->o.rla_enabled) memset(&rl, 0, sizeof(rl));
        ia.pad = 0;
        put(&ia, sizeof(ia));
        put(&rl, sizeof(rl));
        for (int a = 0; a < WMB_N_ALGOS; a++) {
            if ((a == WMB_ALGO_RLA && !c->o.rla_enabled) || (a == WMB_ALGO_T2A && !c->o.t2_enabled)) continue;
            Stream &s = b.s[a];
            StreamDev sd;
            CUDA_TRY(cudaMemcpy(&sd, s.sd, sizeof(sd), cudaMemcpyDeviceToHost));
            const uint32_t sr = (a == WMB_ALGO_T2A) ? sd.t2_sr : 0u;
            put(&sr, 4);
            /* telegrams in flight: their bit events so far (sample, rssi, flags, bit are all in the word) */
            std::vector<uint64_t> pend(gd.n_pend[ch * WMB_N_ALGOS + a]);
            if (!pend.empty()) CUDA_TRY(cudaMemcpy(pend.data(), s.pend, pend.size() * 8, cudaMemcpyDeviceToHost));
            std::sort(pend.begin(), pend.end());
            const uint32_t np = (uint32_t)pend.size();
            put(&np, 4);
            for (uint64_t ord : pend) {
                if (ord >= sd.total) return set_err(WMB_E_STATE, "internal: pending candidate beyond the stream");
                const uint64_t n = sd.total - ord;
                if (n > WMB_MAXBITS + 64) return set_err(WMB_E_STATE, "internal: pending candidate older than a telegram");
                std::vector<uint64_t> ev((size_t)n);
                for (uint64_t i = 0; i < n;) {                         /* the ring wraps */
                    const uint64_t at = (ord + i) & (s.ring_cap - 1);
                    const uint64_t run = std::min<uint64_t>(n - i, s.ring_cap - at);
                    CUDA_TRY(cudaMemcpy(ev.data() + i, s.ring + at, (size_t)run * 8, cudaMemcpyDeviceToHost));
                    i += run;
                }
                /* a match inside a telegram that is already decoded will be ignored (busy decoder) */
                const uint32_t n32 = (uint32_t)n | (((int64_t)ord <= s.busy_until) ? 0x80000000u : 0u);
                put(&n32, 4);
                put(ev.data(), ev.size() * 8);
            }
        }
    }
    if (out.size() > cap) return set_err(WMB_E_INVAL, "buffer too small (%zu bytes needed)", out.size());
    memcpy(buf, out.data(), out.size());
    return (long)out.size();
}

extern "C" long wmb_pending_before(wmb_ctx *c, uint64_t sync_hi)
{
    if (!c) return set_err(WMB_E_INVAL, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ctx_alloc(c);
    if (rc) return rc;
    rc = consume_all(c);                                /* every enqueued batch is gathered and booked first */
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    GatherDev gd;
    CUDA_TRY(cudaMemcpy(&gd, c->d_gd, sizeof(gd), cudaMemcpyDeviceToHost));
    long n_before = 0;
    std::vector<uint64_t> pend;
    for (int ch = 0; ch < WMB_N_CHAINS; ch++) {
        if (!(c->chains & (1u << ch))) continue;
        for (int a = 0; a < WMB_N_ALGOS; a++) {
            if ((a == WMB_ALGO_RLA && !c->o.rla_enabled) || (a == WMB_ALGO_T2A && !c->o.t2_enabled)) continue;
            Stream &s = c->cb[ch].s[a];
            pend.resize(gd.n_pend[ch * WMB_N_ALGOS + a]);
            if (pend.empty()) continue;
            CUDA_TRY(cudaMemcpy(pend.data(), s.pend, pend.size() * 8, cudaMemcpyDeviceToHost));
            for (uint64_t ord : pend) {
                if ((int64_t)ord <= s.busy_until) continue;              /* inside a telegram already decoded: will be ignored */
                uint64_t ev = 0;                                         /* the flagged bit's event: its sample is the match */
                CUDA_TRY(cudaMemcpy(&ev, s.ring + (ord & (s.ring_cap - 1)), 8, cudaMemcpyDeviceToHost));
                const uint64_t m = c->m_consumed - ((c->m_consumed - EVG_M(ev)) & EVG_M_MASK);   /* 40 bits -> stream position */
                if (m < sync_hi) n_before++;
            }
        }
    }
    return n_before;
}

extern "C" int wmb_get_stats(wmb_ctx *c, wmb_stats *s)
{
    if (!c || !s) return set_err(WMB_E_INVAL, "null argument");
    *s = c->st;
    return WMB_OK;
}

extern "C" long wmb_debug_copy_stage(wmb_ctx *c, int chain, float *dphi, uint8_t *rssi, size_t cap)
{
    if (!c || chain < 0 || chain >= WMB_N_CHAINS || !c->allocated || !c->cb[chain].set[0].dphi)
        return set_err(WMB_E_INVAL, "stage not available");
    const SetBuf &sb = c->cb[chain].set[c->last_set];
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    /* the last batch's outputs still sit at [W, W + last_M) until the next batch overwrites them */
    const size_t n = std::min<size_t>((size_t)c->last_M, cap);
    if (dphi) CUDA_TRY(cudaMemcpy(dphi, sb.dphi + c->W, n * 4, cudaMemcpyDeviceToHost));
    if (rssi) CUDA_TRY(cudaMemcpy(rssi, sb.rssi + c->W, n, cudaMemcpyDeviceToHost));
    return (long)n;
}

extern "C" long wmb_debug_copy_bits(wmb_ctx *c, int chain, int which, uint32_t *words, size_t cap_words)
{
    if (!c || chain < 0 || chain >= WMB_N_CHAINS || !words || !c->allocated || !c->cb[chain].set[0].dbits)
        return set_err(WMB_E_INVAL, "stage not available");
    const SetBuf &b = c->cb[chain].set[c->last_set];
    const uint32_t *src = which == 0 ? b.dbits : which == 1 ? b.sbits : which == 2 ? b.cbits : nullptr;
    if (!src) return set_err(WMB_E_INVAL, which == 2 ? "clock signs are kept only by a context created with opts.reserved[1] = 1" : "no such tap");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    const size_t n = std::min<size_t>(((size_t)c->last_M + 31) / 32, cap_words);
    CUDA_TRY(cudaMemcpy(words, src + c->W / 32, n * 4, cudaMemcpyDeviceToHost));
    return (long)n;
}

extern "C" long wmb_debug_copy_events(wmb_ctx *c, int chain, int algo, uint64_t *ev, size_t cap)
{
    if (!c || chain < 0 || chain >= WMB_N_CHAINS || algo < 0 || algo >= WMB_N_ALGOS || !ev || !c->allocated || !c->cb[chain].s[algo].ring)
        return set_err(WMB_E_INVAL, "stream not available");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->cs));
    const Stream &s = c->cb[chain].s[algo];
    const uint64_t n = std::min<uint64_t>(s.total - s.total_prev, cap);
    if (s.total - s.total_prev > s.ring_cap) return set_err(WMB_E_OVERFLOW, "the batch wrote more events than the ring holds");
    for (uint64_t i = 0; i < n;) {                                   /* the ring wraps */
        const uint64_t at = (s.total_prev + i) & (s.ring_cap - 1);
        const uint64_t run = std::min<uint64_t>(n - i, s.ring_cap - at);
        CUDA_TRY(cudaMemcpy(ev + i, s.ring + at, (size_t)run * 8, cudaMemcpyDeviceToHost));
        i += run;
    }
    return (long)n;
}

/* Test hook: run the device's exact-arithmetic building blocks on caller-made operands (host arrays of n floats).
 * mode 0: bounded atan2f(y, x)   1: general atan2f(y, x)   2: bounded division y / x   3: bounded sqrt(y)
 *      4: discriminator of (y[i], x[i]) against (y[i-1], x[i-1]) taken as (I, Q) */
extern "C" int wmb_debug_arith(wmb_ctx *c, int mode, const float *y, const float *x, float *out, size_t n)
{
    if (!c || !y || !x || !out || mode < 0 || mode > 4) return set_err(WMB_E_INVAL, "bad argument");
#ifdef WMB_HOSTSIM
    WmbAtanTab tab;
    for (int i = 0; i < WMB_ATAN_TAB_ELEMS; i++) wmb_atan_tab_fill(&tab, i);
    for (size_t i = 0; i < n; i++)
        out[i] = mode == 0 ? wmb_atan2f_bounded(y[i], x[i], &tab) : mode == 1 ? wmb_atan2f(y[i], x[i])
               : mode == 2 ? wmb_fdiv_bounded(y[i], x[i]) : mode == 3 ? wmb_fsqrt_pos(y[i])
               : wmb_discriminator(y[i], x[i], y[i ? i - 1 : 0], x[i ? i - 1 : 0], &tab);
    return WMB_OK;
#else
    CUDA_TRY(cudaSetDevice(c->device));
    float *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 3 * n * sizeof(float) + 16));
    /* (everything on the context's stream: a plain cudaMemcpy from pageable memory may still be in flight on the
     * legacy stream when a kernel on a non-blocking stream starts) */
    cudaError_t e = cudaMemcpyAsync(d, y, n * 4, cudaMemcpyHostToDevice, c->cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, x, n * 4, cudaMemcpyHostToDevice, c->cs);
    if (e == cudaSuccess) { dbg_arith_kernel<<<1024, 256, 0, c->cs>>>(d, d + n, d + 2 * n, n, mode); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, c->cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->cs);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(WMB_E_CUDA, "wmb_debug_arith: %s", cudaGetErrorString(e));
    return WMB_OK;
#endif
}
