// capi.cu -- implementation of the C ABI declared in include/b200_ofdm.h (receive side):
// handle objects, device memory, kernel sequencing on one CUDA stream per handle, and the
// device -> host hand-off of decoded frame records.
//
//   b2_mcrx_*     multichannelrx   (lib/multichannelrx.cc:45-195)
//   b2_ofdmsync_* ofdmflexframesync as used by ofdmtxrx (lib/ofdmtxrx.cc:91,482,625), batched
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200_ofdm.h"
#include "design.h"
#include "kernels.h"
#include "capi_util.h"
#include "smpart.h"

using namespace b2;

static_assert(sizeof(FrameRec) == sizeof(b2_frame_rec), "FrameRec must mirror b2_frame_rec");

// ================================================================== synchroniser core
// everything downstream of the channelizer: per-stream state, tables, frame output
struct SyncCore {
    int device = 0;
    cudaStream_t stream = nullptr;
    OfdmPlan plan;
    FftPlan fftM;
    unsigned int streams = 0;
    size_t tmax = 0;                     // max samples per stream per launch
    // tables
    DevBuf t_sctype, t_S0, t_S1, t_data, t_pilot, t_pilotx, t_active, t_seq, t_walk, t_B, t_perm, t_tw, t_rank, t_P, t_arank;
    // state
    DevBuf d_st, d_ring, d_G0, d_R, d_penc, d_ctl;
    unsigned int workers = 1;            // 2: frame-pipelined worker pairs (ofdmsync8.cu)
    // frame-parallel warp-per-worker kernel (ofdmsyncw.cu): wslots workers (slots) per stream
    bool use_w = false;
    unsigned int wslots = 1, wrec_stride = 0;
    DevBuf d_wst, d_wch, d_wRG, d_wrecs, d_waux;
    unsigned int sm_budget = 0;          // SMs the synchroniser kernel may use (0: the whole device); set before init()
    unsigned int launch_id = 0;
    unsigned long long stream_pos = 0;   // samples per stream given to the synchroniser so far
    size_t penc_cap = 0;
    // outputs
    DevBuf d_recs, d_aux, d_arena, d_scratch, d_decoded, d_counters, d_vit, d_crc;
    unsigned int vit_mode = 1;           // conv-coded decode: 0 exact only, 1 speculative first (auto), 2 speculative traceback only (env B2_VIT_MODE)
    unsigned int vit_split = 0;          // frames per launch up to which the decode kernel's 128-thread shape works
    unsigned int vit_grid128 = 0, vit_sms = 148;
    unsigned int vit_ctas = 0, vit_steps = 16384;     // per-CTA Viterbi decision regions of the general decode kernel
    unsigned int recs_cap = 0;
    unsigned long long arena_cap = 0;
    unsigned long long out_cap = 0;      // bytes of decoded payload per batch (= arena_cap for the serial-chain kernels)
    // tap
    DevBuf d_tapX, d_tapc, d_tapi;
    unsigned int tap_cap = 0;
    std::vector<uint32_t> tap_chan; std::vector<uint64_t> tap_index; std::vector<float> tap_X;
    // host mirrors
    unsigned int * h_counters = nullptr; // pinned, 8 uints
    FrameRec * h_recs = nullptr;         // pinned
    uint8_t * h_payload = nullptr;       // pinned
    // results waiting for poll()
    // records in callback order.  payload_offset of record i points into ready_payloads for
    // i < n_compacted, and straight into the pinned copy of the last batch's output (h_payload)
    // for the rest: the payload bytes of a batch are not touched by the host unless it asks
    std::vector<FrameRec> ready;
    std::vector<uint8_t> ready_payloads;
    size_t n_compacted = 0;
    size_t pending_bytes = 0;            // payload bytes of the records beyond n_compacted
    unsigned long long last_used = 0;    // bytes of h_payload the last batch filled
    std::vector<FrameRec> view_recs;     // handed out by poll_view(), alive until the next poll/execute
    std::vector<uint8_t> view_payloads;
    RangeMark * h_range = nullptr;       // pinned copy of d_range
    cudaStream_t xstream = nullptr;      // device -> host copies of finished chunks
    cudaStream_t mstream = nullptr;      // device -> host copies of the chunks' record counts
    std::vector<std::pair<unsigned long long, unsigned int>> order;
    void compact_pending();
    void process_chunk(unsigned int lo, unsigned int hi);
    // kernel config
    int sync_threads = 128;
    size_t sync_smem = 0;
    int decode_grid = 148;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[4] = {0, 0, 0, 0};
    SyncParams sp;
    // a batch = one collect(); its chunks run sync on `stream` and decode on `dstream`
    cudaStream_t dstream = nullptr;      // = dstreams[0]
    static const unsigned int NDS = 3;   // decode launches of successive chunks rotate over NDS streams (a conv-coded
    cudaStream_t dstreams[NDS] = {};     // frame is a long serial recursion: chunks must overlap)
    DevBuf d_range;                      // [chunks+1] record count after each chunk's synchroniser
    unsigned int range_cap = 0, chunk = 0, launches = 0;
    struct ChunkEv { cudaEvent_t s0, s1, d0, d1, x, m; };   // m: the chunk's record / payload counts have reached the host
    std::vector<ChunkEv> cev;            // sync begin/end, decode begin/end of each chunk
    bool timing = true;
    bool host_results = true;            // end_batch brings records + payloads to pinned host memory and orders them

    int init(unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p, unsigned int streams_,
             size_t tmax_, int device_, cudaStream_t st, const cudaStream_t * decode_st = nullptr);
    void destroy();
    int reset_state();                   // fresh object: everything zero
    int reset_streams();                 // ofdmflexframesync_reset on every stream
    // run sync + decode over in[s*stride + t], t < nsamples; appends results to `ready`
    int run(const cf * in, size_t in_stride, unsigned int nsamples, bool record_events);
    int begin_batch();
    // `after`: event on another stream that must complete before the synchroniser may read `in`
    int launch_chunk(const cf * in, size_t in_stride, unsigned int nsamples, cudaEvent_t after);
    int end_batch();
    void fetch_timing();
    bool timing_stale = false;
    int collect();
    int poll(b2_frame_rec * recs, size_t recs_cap, size_t * n_recs, uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes);
    int poll_view(const b2_frame_rec ** recs, size_t * n_recs, const uint8_t ** payloads, size_t * n_payload_bytes);
    int set_tap(int enable, size_t max_symbols);
};

int SyncCore::init(unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p, unsigned int streams_,
                   size_t tmax_, int device_, cudaStream_t st, const cudaStream_t * decode_st)
{
    device = device_; stream = st; streams = streams_; tmax = tmax_;
    if (M < 8 || (M & 1) || cp < 1 || cp > M || taper > cp) return b2_fail(B2_ERR_ARG, "invalid OFDM configuration (M=%u cp=%u taper=%u)", M, cp, taper);
    if (ofdm_plan(plan, M, cp, taper, p) != 0) return b2_fail(B2_ERR_ARG, "invalid subcarrier allocation");
    if (M < 16 || M > 4096 || fft_plan(fftM, M) != 0)
        return b2_fail(B2_ERR_UNSUPPORTED, "the CUDA path needs an even number of subcarriers in [16, 4096] whose prime factors are <= 41 (got %u)", M);
    // tables
    std::vector<cf> B(M);
    {
        float phi = (float)plan.backoff * 2.0f * M_PI / (float)M;
        for (unsigned int i = 0; i < M; i++) { float a = (float)i * phi; B[i] = make_float2(cosf(a), sinf(a)); }
    }
    std::vector<uint16_t> walk(4 * 18);
    {
        unsigned int Mi, Ni;
        interleaver_dims(36, Mi, Ni);
        const unsigned int extra[4] = {0, 2, 4, 8};
        for (int v = 0; v < 4; v++) {
            std::vector<uint16_t> w = interleaver_walk(36, Mi, Ni + extra[v]);
            memcpy(&walk[18 * v], w.data(), 18 * sizeof(uint16_t));
        }
    }
    B2_TRY(t_sctype.upload(plan.p)); B2_TRY(t_S0.upload(plan.S0)); B2_TRY(t_S1.upload(plan.S1));
    B2_TRY(t_data.upload(plan.data_idx)); B2_TRY(t_pilot.upload(plan.pilot_idx)); B2_TRY(t_pilotx.upload(plan.pilot_x));
    B2_TRY(t_active.upload(plan.active_idx)); B2_TRY(t_seq.upload(plan.pilot_seq)); B2_TRY(t_walk.upload(walk));
    std::vector<uint16_t> sc_rank(M, 0xffff);
    for (size_t d = 0; d < plan.data_idx.size(); d++) sc_rank[plan.data_idx[d]] = (uint16_t)d;
    for (size_t n = 0; n < plan.pilot_idx.size(); n++) sc_rank[plan.pilot_idx[n]] = (uint16_t)(0x4000u | n);
    if (plan.M_pilot + plan.M_data < 5) return b2_fail(B2_ERR_UNSUPPORTED, "the CUDA path needs at least 5 active subcarriers");
    B2_TRY(t_P.upload(eqgain_fit_matrix(plan)));
    std::vector<uint16_t> act_rank(M, 0xffff);
    for (size_t n = 0; n < plan.active_idx.size(); n++) act_rank[plan.active_idx[n]] = (uint16_t)n;
    B2_TRY(t_arank.upload(act_rank));
    B2_TRY(t_B.upload(B)); B2_TRY(t_perm.upload(fftM.perm)); B2_TRY(t_tw.upload(fftM.tw)); B2_TRY(t_rank.upload(sc_rank));
    // state
    const size_t W = M + cp;
    // in-progress payload of a stream: one byte per demapped symbol (ofdmsync8.cu; at most 8 symbols
    // per encoded byte, BPSK) or the packed encoded bytes (ofdmsync.cu)
    unsigned int max_payload = 65535;
    if (const char * e = getenv("B2_MAX_PAYLOAD")) {
        unsigned long v = strtoul(e, nullptr, 10);
        if (v >= 1 && v <= 65535) max_payload = (unsigned int)v;
    }
    penc_cap = (8 * (size_t)fec_enc_len(FEC_HAMMING128, fec_enc_len(FEC_CONV_V27, max_payload + 4)) + 64 + 15) & ~(size_t)15;
    // frame-pipelined worker pairs: two CTAs per stream take alternate frames (the second starts its frame
    // search as soon as the first has decoded a header and therefore knows where its frame ends).  Both must be
    // resident at once, so the pair is used only where the SMs given to the synchroniser hold 2 * streams CTAs.
    workers = 1;
    {
        const bool fast = sync8_supported(M) && plan.M_pilot + plan.M_data >= 5 && getenv("B2_SYNC_GENERIC") == nullptr;
        int want = 2;
        if (const char * e = getenv("B2_SYNC_WORKERS")) want = atoi(e);
        if (fast && want == 2) {
            SyncParams probe;
            memset(&probe, 0, sizeof(probe));
            probe.M = M; probe.cp = cp; probe.M_pilot = plan.M_pilot; probe.M_data = plan.M_data;
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            if (sm_budget && (int)sm_budget < sms) sms = (int)sm_budget;
            const long long cap = (long long)sync8_ctas_per_sm(probe) * sms;
            if (2ll * streams <= cap) workers = 2;
        }
    }
    // the frame-parallel kernel takes the shapes it is built for unless told otherwise (B2_SYNC_LEGACY=1: the
    // serial-chain kernels of round 1)
    use_w = syncw_supported(M) && plan.M_pilot + plan.M_data >= 5 && getenv("B2_SYNC_LEGACY") == nullptr && getenv("B2_SYNC_GENERIC") == nullptr;
    if (use_w) {
        workers = 1;
        // slots for about eight waves of 16 warps per SM (a launch uses as many as its samples are worth)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        unsigned int k = (unsigned int)((8u * 16u * (unsigned int)sms + streams - 1) / streams);
        k = std::max(1u, std::min(128u, k));
        if (const char * e = getenv("B2_SYNC_K")) { int v = atoi(e); if (v >= 1 && v <= 64) k = (unsigned int)v; }
        wslots = k;
    }
    const size_t vstreams = (size_t)streams * (use_w ? wslots : workers);
    B2_TRY(d_st.alloc(sizeof(SyncState) * vstreams)); B2_TRY(d_ring.alloc(sizeof(cf) * W * vstreams));
    B2_TRY(d_G0.alloc(sizeof(cf) * M * vstreams)); B2_TRY(d_R.alloc(sizeof(cf) * M * vstreams));
    // (the frame-parallel kernel demaps straight into the arena)
    B2_TRY(d_penc.alloc(use_w ? 16 : penc_cap * vstreams));
    B2_TRY(d_ctl.alloc(sizeof(SyncCtl) * streams));
    if (use_w) {
        wrec_stride = (unsigned int)(tmax / (2 * W)) + 2 * wslots + 4;
        B2_TRY(d_wst.alloc(sizeof(WSync) * vstreams)); B2_TRY(d_wch.alloc(sizeof(WChan) * streams));
        B2_TRY(d_wRG.alloc(sizeof(cf) * M * vstreams));
        B2_TRY(d_wrecs.alloc(sizeof(FrameRec) * (size_t)wrec_stride * streams)); B2_TRY(d_waux.alloc(sizeof(FrameAux) * (size_t)wrec_stride * streams));
    }
    // outputs: a frame needs at least 4 OFDM symbols; payload bits <= 8 per sample
    recs_cap = (unsigned int)(streams * (tmax / (2 * W) + 4));
    // arena of a batch: the symbols demapped inside the batch (<= one byte per sample) plus, per stream, one
    // frame that began in earlier batches and completes in this one (bounded here to 32 KB of symbol bytes;
    // B2_ERR_OVERFLOW reports a frame that does not fit)
    // (up to 64 MB of it: with few streams -- the single-link programs -- any legal frame fits; beyond that a frame
    // that does not fit is reported without its payload)
    const size_t carry = ((unsigned long long)streams * penc_cap <= (64ull << 20)) ? penc_cap : std::min<size_t>(penc_cap, 32768);
    arena_cap = (unsigned long long)streams * (tmax + 64 + carry) + 16ull * recs_cap;
    // (frame-parallel kernel: a stretch whose prediction failed is demapped twice, once speculatively and once by
    // the stitcher, and the speculative copy's arena space is simply left unused)
    out_cap = arena_cap;
    if (use_w) {
        // the arena is a ring of demapped symbols that outlives batches: what a batch demaps (twice: see above), the
        // frames in progress, and room for any single legal frame
        out_cap = arena_cap * (wslots > 1 ? 2 : 1);
        arena_cap = std::max<unsigned long long>(2 * arena_cap, 4ull * penc_cap) & ~15ull;
    }
    B2_TRY(d_recs.alloc(sizeof(FrameRec) * recs_cap)); B2_TRY(d_aux.alloc(sizeof(FrameAux) * recs_cap));
    B2_TRY(d_arena.alloc(arena_cap)); B2_TRY(d_scratch.alloc(arena_cap)); B2_TRY(d_decoded.alloc(out_cap));
    B2_TRY(d_counters.alloc(8 * sizeof(unsigned int)));
    B2_TRY(d_crc.alloc(8 * sizeof(unsigned int)));
    B2_CUDA(cudaMemset(d_crc.p, 0, 8 * sizeof(unsigned int)));
    B2_CUDA(cudaMallocHost(&h_counters, 8 * sizeof(unsigned int)));
    B2_CUDA(cudaMallocHost(&h_recs, sizeof(FrameRec) * recs_cap));
    B2_CUDA(cudaMallocHost(&h_payload, out_cap));
    for (int i = 0; i < 5; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    for (unsigned int i = 0; i < NDS; i++) {
        if (decode_st) dstreams[i] = decode_st[i];
        else B2_CUDA(cudaStreamCreateWithFlags(&dstreams[i], cudaStreamNonBlocking));
    }
    dstream = dstreams[0];
    B2_CUDA(cudaStreamCreateWithFlags(&xstream, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&mstream, cudaStreamNonBlocking));
    range_cap = 4096;
    B2_TRY(d_range.alloc(sizeof(RangeMark) * (range_cap + 1)));
    B2_CUDA(cudaMallocHost(&h_range, sizeof(RangeMark) * (range_cap + 1)));
    memset(h_range, 0, sizeof(RangeMark) * (range_cap + 1));
    timing = getenv("B2_NO_TIMING") == nullptr;

    memset(&sp, 0, sizeof(sp));
    sp.M = M; sp.cp = cp; sp.M2 = M / 2; sp.backoff = plan.backoff;
    sp.M_pilot = plan.M_pilot; sp.M_data = plan.M_data; sp.M_S0 = plan.M_S0; sp.M_S1 = plan.M_S1;
    sp.thresh = plan.thresh; sp.pilot_sx = plan.pilot_sx; sp.pilot_sxx = plan.pilot_sxx;
    {
        const float a = (float)plan.backoff * 2.0f * 3.14159274101257324219f / (float)M;
        sp.b_cos = cosf(a); sp.b_sin = sinf(a);
    }
    for (int i = 0; i < 9; i++) sp.qam_alpha[i] = 1.0f;
    sp.qam_alpha[2] = 1.0f / sqrtf(2.0f); sp.qam_alpha[4] = 1.0f / sqrtf(10.0f);
    sp.qam_alpha[6] = 1.0f / sqrtf(42.0f); sp.qam_alpha[8] = 1.0f / sqrtf(170.0f);
    sp.streams = streams;
    sp.workers = workers; sp.ctl = d_ctl.as<SyncCtl>();
    sp.st = d_st.as<SyncState>(); sp.ring = d_ring.as<cf>(); sp.G0 = d_G0.as<cf>(); sp.R = d_R.as<cf>();
    sp.penc = d_penc.as<uint8_t>(); sp.penc_cap = penc_cap;
    sp.recs = d_recs.as<FrameRec>(); sp.aux = d_aux.as<FrameAux>(); sp.recs_cap = recs_cap;
    sp.arena = d_arena.as<uint8_t>(); sp.arena_cap = arena_cap; sp.decoded_cap = out_cap;
    sp.counters = d_counters.as<unsigned int>();
    sp.tb.sctype = t_sctype.as<uint8_t>(); sp.tb.S0 = t_S0.as<float>(); sp.tb.S1 = t_S1.as<float>();
    sp.tb.data_idx = t_data.as<uint16_t>(); sp.tb.pilot_idx = t_pilot.as<uint16_t>(); sp.tb.pilot_x = t_pilotx.as<float>();
    sp.tb.active_idx = t_active.as<uint16_t>(); sp.tb.pilot_seq = t_seq.as<uint8_t>(); sp.tb.hdr_walk = t_walk.as<uint16_t>();
    sp.tb.B = t_B.as<cf>(); sp.tb.sc_rank = t_rank.as<uint16_t>(); sp.tb.eqfit_P = t_P.as<double>(); sp.tb.act_rank = t_arank.as<uint16_t>();
    sp.fft.n = fftM.n; sp.fft.npass = fftM.npass;
    sp.fft.radices = 0;
    for (unsigned int i = 0; i < fftM.npass; i++) sp.fft.radices |= fft_radix_code(fftM.radix[i]) << (4 * i);
    sp.fft.perm = t_perm.as<uint16_t>(); sp.fft.tw = t_tw.as<cf>();
    sp.wst = d_wst.as<WSync>(); sp.wch = d_wch.as<WChan>(); sp.wRG = d_wRG.as<cf>();
    sp.wrecs = d_wrecs.as<FrameRec>(); sp.waux = d_waux.as<FrameAux>(); sp.wrec_stride = wrec_stride; sp.wslots = wslots;
    sync_smem = sync_smem_bytes(sp);
    const bool fast_sync = sync8_supported(M) && plan.M_pilot + plan.M_data >= 5 && getenv("B2_SYNC_GENERIC") == nullptr;
    if (fast_sync) sync_smem = sync8_smem_bytes(sp);          // the kernel sync_launch() will pick
    if (sync_smem > 227 * 1024) return b2_fail(B2_ERR_UNSUPPORTED, "M=%u needs %zu bytes of shared memory per stream", M, sync_smem);
    B2_CUDA(sync_configure(sync_smem));
    sync_threads = (M >= 512) ? 256 : 128;
    if (const char * e = getenv("B2_SYNC_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) sync_threads = v; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    decode_grid = sms * 2;
    // Viterbi decision regions (16 384 steps = 128 KB each), one per CTA of the general decode kernel.  A launch takes a third
    // of them (the decode streams rotate): up to 16 one-warp CTAs per SM in its 32-thread shape, 4 per SM in the 128-thread
    // one; a handle with few streams never has that many frames in one launch and gets fewer (32 per stream)
    vit_ctas = NDS * std::min((unsigned int)sms * 16u, std::max(32u, 32u * streams));
    vit_sms = (unsigned int)sms;
    if (getenv("B2_VIT_SERIAL")) vit_mode = 0;
    if (const char * e = getenv("B2_VIT_MODE")) { int v = atoi(e); if (v >= 0 && v <= 2) vit_mode = (unsigned int)v; }
    if (const char * e = getenv("B2_VIT_CTAS")) { int v = atoi(e); if (v >= 3 && v <= 96) vit_ctas = (unsigned int)(sms * v) / NDS * NDS; }   // per SM, over the NDS decode streams
    // the 128-thread shape runs 4 CTAs per SM; the one-warp shape takes over only where it has more CTAs than that to offer
    // (with as many or fewer -- a handle with few streams -- it would be the same frames at a quarter of the width)
    vit_grid128 = std::min(vit_ctas / NDS, vit_sms * 4u);
    vit_split = (vit_ctas / NDS > vit_grid128) ? vit_grid128 : 0xffffffffu;
    B2_TRY(d_vit.alloc(sizeof(uint2) * (size_t)vit_ctas * vit_steps));
    if (packet_decode_prepare() != cudaSuccess) return b2_fail(B2_ERR_NOMEM, "could not allocate the Viterbi workspace: %s", cudaGetErrorString(cudaGetLastError()));
    return reset_state();
}

void SyncCore::destroy()
{
    if (h_counters) cudaFreeHost(h_counters);
    if (h_recs) cudaFreeHost(h_recs);
    if (h_payload) cudaFreeHost(h_payload);
    if (h_range) cudaFreeHost(h_range);
    if (xstream) cudaStreamDestroy(xstream);
    if (mstream) cudaStreamDestroy(mstream);
    mstream = nullptr;
    h_counters = nullptr; h_recs = nullptr; h_payload = nullptr; h_range = nullptr; xstream = nullptr;
    for (int i = 0; i < 5; i++) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
    for (auto & e : cev) { cudaEventDestroy(e.s0); cudaEventDestroy(e.s1); cudaEventDestroy(e.d0); cudaEventDestroy(e.d1); cudaEventDestroy(e.x); cudaEventDestroy(e.m); }
    cev.clear();
    for (unsigned int i = 0; i < NDS; i++) if (dstreams[i]) { cudaStreamDestroy(dstreams[i]); dstreams[i] = nullptr; }
    dstream = nullptr;
}

int SyncCore::reset_state()
{
    // ofdmflexframesync_reset on every stream; the sample window is NOT cleared by liquid's
    // reset, but a freshly created object starts from zeros -- reset_state() is also create
    std::vector<SyncState> st((size_t)streams * workers);
    for (size_t i = 0; i < st.size(); i++) {
        sync_state_init(st[i], plan.M, plan.cp);
        st[i].role = (workers == 2 && (i & 1)) ? SW_WAIT : SW_OWNER;
    }
    stream_pos = 0;
    if (use_w) {
        // every slot zero (state SEEK = 0, timer 0, nothing mixed); head slot 0; prediction = the idle seek grid
        std::vector<WChan> wc(streams);
        memset(wc.data(), 0, sizeof(WChan) * streams);
        for (auto & c : wc) { c.pred_next[0] = c.pred_next[1] = plan.M; c.pred_period[0] = c.pred_period[1] = plan.M; }
        B2_CUDA(cudaMemsetAsync(d_counters.p, 0, d_counters.bytes, stream));      // incl. the ring's allocation counter
        B2_CUDA(cudaMemsetAsync(d_wst.p, 0, d_wst.bytes, stream));
        B2_CUDA(cudaMemcpyAsync(d_wch.p, wc.data(), sizeof(WChan) * streams, cudaMemcpyHostToDevice, stream));
        B2_CUDA(cudaMemsetAsync(d_wRG.p, 0, d_wRG.bytes, stream));
        B2_CUDA(cudaMemsetAsync(d_ring.p, 0, d_ring.bytes, stream));
        B2_CUDA(cudaStreamSynchronize(stream));
        return B2_OK;
    }
    B2_CUDA(cudaMemcpyAsync(d_st.p, st.data(), sizeof(SyncState) * st.size(), cudaMemcpyHostToDevice, stream));
    B2_CUDA(cudaMemsetAsync(d_ctl.p, 0, d_ctl.bytes, stream));
    B2_CUDA(cudaMemsetAsync(d_ring.p, 0, d_ring.bytes, stream));
    B2_CUDA(cudaMemsetAsync(d_G0.p, 0, d_G0.bytes, stream));
    B2_CUDA(cudaMemsetAsync(d_R.p, 0, d_R.bytes, stream));
    B2_CUDA(cudaStreamSynchronize(stream));
    return B2_OK;
}

// payloads of the records that still point into h_payload move into ready_payloads (needed
// before the next batch overwrites h_payload, or when a caller wants one compact buffer)
void SyncCore::compact_pending()
{
    if (n_compacted == ready.size()) return;
    size_t o = ready_payloads.size();
    ready_payloads.resize(o + pending_bytes);
    for (size_t i = n_compacted; i < ready.size(); i++) {
        FrameRec & r = ready[i];
        const size_t len = r.header_valid ? r.payload_len : 0;
        if (len) memcpy(ready_payloads.data() + o, h_payload + r.payload_offset, len);
        r.payload_offset = o;
        o += len;
    }
    n_compacted = ready.size();
    pending_bytes = 0;
}

int SyncCore::poll_view(const b2_frame_rec ** recs, size_t * n_recs, const uint8_t ** payloads, size_t * n_payload_bytes)
{
    view_recs.clear(); view_payloads.clear();
    const bool zero_copy = (n_compacted == 0);           // everything waiting comes from the last batch
    if (!zero_copy) compact_pending();
    view_recs.swap(ready);
    if (!zero_copy) view_payloads.swap(ready_payloads);
    if (recs) *recs = (const b2_frame_rec *)view_recs.data();
    if (n_recs) *n_recs = view_recs.size();
    if (payloads) *payloads = zero_copy ? h_payload : view_payloads.data();
    if (n_payload_bytes) *n_payload_bytes = zero_copy ? (view_recs.empty() ? 0 : (size_t)last_used) : view_payloads.size();
    n_compacted = 0; pending_bytes = 0;
    return B2_OK;
}

int SyncCore::reset_streams()
{
    if (use_w) B2_CUDA(syncw_reset_launch(d_wst.as<WSync>(), d_wch.as<WChan>(), streams, wslots, plan.M, stream_pos, stream));
    else B2_CUDA(sync_reset_launch(d_st.as<SyncState>(), streams, workers, d_ctl.as<SyncCtl>(), stream_pos, stream));
    B2_CUDA(cudaStreamSynchronize(stream));
    return B2_OK;
}

int SyncCore::set_tap(int enable, size_t max_symbols)
{
    tap_cap = 0;
    if (!enable) return B2_OK;
    B2_TRY(d_tapX.alloc(sizeof(cf) * plan.M * max_symbols));
    B2_TRY(d_tapc.alloc(sizeof(uint32_t) * max_symbols));
    B2_TRY(d_tapi.alloc(sizeof(uint64_t) * max_symbols));
    tap_cap = (unsigned int)max_symbols;
    return B2_OK;
}

int SyncCore::run(const cf * in, size_t in_stride, unsigned int nsamples, bool record_events)
{
    (void)record_events;
    if (nsamples == 0) return B2_OK;
    B2_TRY(begin_batch());
    B2_TRY(launch_chunk(in, in_stride, nsamples, nullptr));
    return end_batch();
}

int SyncCore::begin_batch()
{
    compact_pending();                       // h_payload is about to be overwritten
    // (frame-parallel kernel: counters[2..3] is the allocation counter of the symbol ring, it keeps counting across batches)
    B2_CUDA(batch_reset_launch(d_counters.as<unsigned int>(), d_range.as<RangeMark>(), use_w ? 1 : 0, stream));
    memset(&h_range[0], 0, sizeof(RangeMark));
    chunk = 0; launches = 0;
    return B2_OK;
}

int SyncCore::launch_chunk(const cf * in, size_t in_stride, unsigned int nsamples, cudaEvent_t after)
{
    if (nsamples == 0) return B2_OK;
    if (nsamples > tmax) return b2_fail(B2_ERR_ARG, "internal: launch of %u samples exceeds tmax %zu", nsamples, tmax);
    if (chunk >= range_cap) return b2_fail(B2_ERR_ARG, "internal: too many chunks in one batch");
    if (cev.size() <= chunk) {
        ChunkEv e;
        B2_CUDA(cudaEventCreate(&e.s0)); B2_CUDA(cudaEventCreate(&e.s1));
        B2_CUDA(cudaEventCreate(&e.d0)); B2_CUDA(cudaEventCreate(&e.d1));
        B2_CUDA(cudaEventCreateWithFlags(&e.x, cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&e.m, cudaEventDisableTiming));
        cev.push_back(e);
    }
    ChunkEv & e = cev[chunk];
    if (after) B2_CUDA(cudaStreamWaitEvent(stream, after, 0));
    SyncParams q = sp;
    q.in = in; q.in_stride = in_stride; q.nsamples = nsamples;
    q.sample_base = stream_pos; q.launch_id = ++launch_id;
    stream_pos += nsamples;
    q.tap_cap = tap_cap;
    q.tap_X = d_tapX.as<cf>(); q.tap_chan = d_tapc.as<uint32_t>(); q.tap_index = d_tapi.as<unsigned long long>();
    if (timing) B2_CUDA(cudaEventRecord(e.s0, stream));
    if (use_w) {
        // workers of this launch: a stretch should hold a few OFDM symbols at least; the debug tap wants the
        // symbols of the serial chain only (speculative workers would tap symbols that are thrown away)
        const unsigned int W = plan.M + plan.cp;
        q.workers = tap_cap ? 1u : std::max(1u, std::min(wslots, nsamples / (4u * W)));
        B2_CUDA(syncw_launch(q, stream));
    } else B2_CUDA(sync_launch(q, sync_threads, sync_smem, stream));
    RangeMark * range = d_range.as<RangeMark>() + chunk;
    B2_CUDA(record_mark_launch(d_counters.as<unsigned int>(), range + 1, stream, use_w ? 6 : 2, h_range + chunk + 1));
    B2_CUDA(cudaEventRecord(e.s1, stream));
    // (the mark kernel has written the counts into the pinned host copy itself: a small D2H copy would be a copy-engine
    // operation and queue behind whatever bulk D2H is in flight -- payloads of earlier chunks, a gather)
    B2_CUDA(cudaEventRecord(e.m, stream));
    // decode of this chunk runs beside the synchroniser of the next one
    cudaStream_t ds = dstreams[chunk % NDS];
    B2_CUDA(cudaStreamWaitEvent(ds, e.s1, 0));
    PacketParams pp;
    pp.recs = d_recs.as<FrameRec>(); pp.aux = d_aux.as<FrameAux>(); pp.range = range;
    pp.arena = d_arena.as<uint8_t>(); pp.scratch = d_scratch.as<uint8_t>(); pp.decoded = d_decoded.as<uint8_t>();
    // every decode stream has its own share of the Viterbi regions (launches on different streams overlap)
    pp.vit_local_ctas = vit_ctas / NDS; pp.vit_local_steps = vit_steps; pp.vit_split = vit_split; pp.vit_grid128 = vit_grid128;
    pp.vit_local = d_vit.as<uint2>() + (size_t)(chunk % NDS) * pp.vit_local_ctas * vit_steps;
    pp.crc_cache = d_crc.as<unsigned int>();
    pp.vit_parallel = vit_mode;
    if (timing) B2_CUDA(cudaEventRecord(e.d0, ds));
    B2_CUDA(packet_decode_launch(pp, decode_grid, ds));
    B2_CUDA(cudaEventRecord(e.d1, ds));
    chunk++; launches += 5;                  // synchroniser, record mark, three decode kernels (one shape of the general one returns at once)
    return B2_OK;
}

// records [lo, hi) of h_recs (one chunk: their completion indices all lie inside the chunk) go to
// `ready` in the order the reference would have fired their callbacks: completion block, then
// channel.  Payload bytes stay where the DMA put them.
void SyncCore::process_chunk(unsigned int lo, unsigned int hi)
{
    order.clear();
    for (unsigned int i = lo; i < hi; i++) order.push_back({(h_recs[i].complete_index << 16) | (h_recs[i].channel & 0xffffu), i});
    std::sort(order.begin(), order.end());
    for (auto & kv : order) {
        const FrameRec & r = h_recs[kv.second];
        ready.push_back(r);
        if (r.header_valid) pending_bytes += r.payload_len;
    }
}

int SyncCore::end_batch()
{
    if (chunk == 0) return B2_OK;
    // chunk by chunk, as soon as its decode is done: records + decoded payloads -> pinned host memory on
    // the copy stream, and the host orders chunk c-1 while chunk c is in flight
    int rc = B2_OK;
    if (!host_results) {
        // the frames stay in device memory (b2_mcrx_shard_pack_results): wait for the last decode, keep the counts
        for (unsigned int i = 0; i < NDS && i < chunk; i++) B2_CUDA(cudaEventSynchronize(cev[chunk - 1 - i].d1));
        B2_CUDA(cudaEventSynchronize(cev[chunk - 1].m));
        last_used = std::min(h_range[chunk].arena_used, out_cap);
        timing_stale = true;
        return collect();
    }
    ready.reserve(ready.size() + 1024);
    for (unsigned int c = 0; c <= chunk; c++) {
        if (c < chunk) {
            B2_CUDA(cudaEventSynchronize(cev[c].d1));
            B2_CUDA(cudaEventSynchronize(cev[c].m));
            const unsigned int lo = std::min(h_range[c].nrec, recs_cap), hi = std::min(h_range[c + 1].nrec, recs_cap);
            const unsigned long long ulo = std::min(h_range[c].arena_used, out_cap), uhi = std::min(h_range[c + 1].arena_used, out_cap);
            if (hi > lo) B2_CUDA(cudaMemcpyAsync(h_recs + lo, d_recs.as<FrameRec>() + lo, sizeof(FrameRec) * (hi - lo), cudaMemcpyDeviceToHost, xstream));
            if (uhi > ulo) B2_CUDA(cudaMemcpyAsync(h_payload + ulo, d_decoded.as<uint8_t>() + ulo, uhi - ulo, cudaMemcpyDeviceToHost, xstream));
            B2_CUDA(cudaEventRecord(cev[c].x, xstream));
        }
        if (c > 0) {
            B2_CUDA(cudaEventSynchronize(cev[c - 1].x));
            const unsigned int lo = std::min(h_range[c - 1].nrec, recs_cap), hi = std::min(h_range[c].nrec, recs_cap);
            process_chunk(lo, hi);

        }
    }

    last_used = std::min(h_range[chunk].arena_used, out_cap);
    rc = collect();
    timing_stale = true;                     // the per-kernel sums are computed when somebody asks (fetch_timing)
    return rc;
}

// per-kernel device times of the last batch, from the chunk events (valid until the next batch)
void SyncCore::fetch_timing()
{
    if (!timing_stale) return;
    timing_stale = false;
    last_ms[1] = 0.f; last_ms[2] = 0.f;
    if (!timing) return;
    for (unsigned int i = 0; i < chunk; i++) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, cev[i].s0, cev[i].s1);
        cudaEventElapsedTime(&b, cev[i].d0, cev[i].d1);
        last_ms[1] += a; last_ms[2] += b;
    }
}

// end of a batch: overflow flag and the debug tap
int SyncCore::collect()
{
    // the overflow flag came back with the last chunk's mark; only the debug tap needs another trip
    // (bit 8: a frame larger than half the symbol arena was reported without its payload -- data, not an error)
    if (h_range[chunk].pad & 7u) return b2_fail(B2_ERR_OVERFLOW, (h_range[chunk].pad & 2u) ? "internal: synchroniser worker hand-off timed out"
                                       : "frame output arena overflow: a frame larger than the per-call buffers completed; create the handle with a larger max_batch");
    if (!tap_cap) return B2_OK;
    B2_CUDA(cudaMemcpyAsync(h_counters, d_counters.p, 8 * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    B2_CUDA(cudaStreamSynchronize(stream));
    const unsigned int ntap = std::min(h_counters[4], tap_cap);
    if (ntap) {
        size_t o = tap_chan.size();
        tap_chan.resize(o + ntap); tap_index.resize(o + ntap); tap_X.resize((o + ntap) * 2 * (size_t)plan.M);
        B2_CUDA(cudaMemcpyAsync(&tap_chan[o], d_tapc.p, sizeof(uint32_t) * ntap, cudaMemcpyDeviceToHost, stream));
        B2_CUDA(cudaMemcpyAsync(&tap_index[o], d_tapi.p, sizeof(uint64_t) * ntap, cudaMemcpyDeviceToHost, stream));
        B2_CUDA(cudaMemcpyAsync(&tap_X[o * 2 * (size_t)plan.M], d_tapX.p, sizeof(cf) * plan.M * ntap, cudaMemcpyDeviceToHost, stream));
        B2_CUDA(cudaStreamSynchronize(stream));
    }
    return B2_OK;
}

int SyncCore::poll(b2_frame_rec * recs, size_t cap, size_t * n_recs, uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes)
{
    const size_t total = ready_payloads.size() + pending_bytes;
    if (n_recs) *n_recs = ready.size();
    if (n_payload_bytes) *n_payload_bytes = total;
    if (!recs) return B2_OK;
    if (cap < ready.size() || (payloads_cap < total)) return b2_fail(B2_ERR_OVERFLOW, "poll buffers too small");
    FrameRec * out = (FrameRec *)recs;
    if (!ready.empty()) memcpy(out, ready.data(), ready.size() * sizeof(FrameRec));
    if (payloads) {
        if (!ready_payloads.empty()) memcpy(payloads, ready_payloads.data(), ready_payloads.size());
        size_t o = ready_payloads.size();
        for (size_t i = n_compacted; i < ready.size(); i++) {       // straight from the pinned batch output
            const size_t len = out[i].header_valid ? out[i].payload_len : 0;
            if (len) memcpy(payloads + o, h_payload + out[i].payload_offset, len);
            out[i].payload_offset = o;
            o += len;
        }
    }
    ready.clear(); ready_payloads.clear();
    n_compacted = 0; pending_bytes = 0;
    return B2_OK;
}

// ================================================================== multichannelrx
struct b2_mcrx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int N = 0, K = 0, lgK = 0, P = 14, TB = 8;
    size_t max_batch = 0;
    FftPlan fftK;
    DevBuf t_taps, t_perm, t_tw;
    DevBuf d_stage;                      // [hist (P-1)K][carry < K][new samples <= max_batch]
    DevBuf d_tail;                       // scratch for moving the stream tail to the front
    DevBuf d_chan;                       // channelizer output [N][tcap]
    size_t tcap = 0;
    size_t hist_len = 0, carry = 0;
    uint32_t nco_theta = 0, nco_dtheta = 0;      // phase of the next incoming sample
    size_t an_smem = 0;
    int an_grid = 148;
    unsigned int an_sms = 148;           // SMs the channelizer may use
    unsigned int last_blocks = 0;
    // pipeline: copies on cstream, channelizer on stream, synchronisers on sstream (core.stream),
    // decode on core.dstream; chunk c of a call flows through all four while chunk c+1 follows
    cudaStream_t cstream = nullptr, sstream = nullptr;
    SmPartition part;                    // SMs split between the synchroniser chains and the throughput kernels
    cudaStream_t spare_stream = nullptr; // the partition stream this handle does not use (destroyed with it)
    unsigned int chunk_blocks = 0;
    struct AnEv { cudaEvent_t copied, a0, a1; };
    std::vector<AnEv> aev;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    unsigned int timed_chunks = 0;       // chunks of the last call whose events await mcrx_fetch_timing()
    SyncCore core;
    // optional rate-matching stage ahead of the NCO (b2_mcrx_set_resampler)
    b2_msresamp * rs = nullptr;
    DevBuf d_rs;
    size_t rs_in_chunk = 0, rs_cap = 0;
};

// device times of the last call, computed from the events when somebody asks for them
static void mcrx_fetch_timing(b2_mcrx_s * q)
{
    q->core.fetch_timing();
    if (!q->timed_chunks) return;
    q->core.last_ms[0] = 0.f;
    if (q->core.timing)
        for (unsigned int i = 0; i < q->timed_chunks; i++) { float a = 0.f; cudaEventElapsedTime(&a, q->aev[i].a0, q->aev[i].a1); q->core.last_ms[0] += a; }
    cudaEventElapsedTime(&q->core.last_ms[3], q->ev_begin, q->ev_end);
    q->timed_chunks = 0;
}

static int mcrx_process(b2_mcrx * q, const float * x, size_t n, bool on_device);

extern "C" int b2_mcrx_create(unsigned int N, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p,
                              int device, size_t max_batch, b2_mcrx ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    // same argument checks as multichannelrx::multichannelrx (lib/multichannelrx.cc:54-66)
    if (N < 1) return b2_fail(B2_ERR_ARG, "must have at least one channel");
    if (M < 8) return b2_fail(B2_ERR_ARG, "number of subcarriers must be at least 8");
    if (cp < 1) return b2_fail(B2_ERR_ARG, "cyclic prefix length must be at least 1");
    if (taper > cp) return b2_fail(B2_ERR_ARG, "taper length cannot exceed cyclic prefix length");
    unsigned int K = 2 * N;
    {
        FftPlan probe;
        if (K > 1024 || fft_plan(probe, K) != 0)
            return b2_fail(B2_ERR_UNSUPPORTED, "the CUDA channelizer needs a channel count <= 512 whose prime factors are <= 41 (got %u)", N);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return b2_fail(B2_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return b2_fail(B2_ERR_ARG, "invalid device ordinal %d", device);
    B2_CUDA(cudaSetDevice(device));
    b2_mcrx * q = new b2_mcrx_s;
    q->device = device;
    q->N = N; q->K = K; q->lgK = ceil_log2(K);
    q->TB = (4096u / K >= 8u) ? (std::min(256u, 4096u / K) & ~7u) : 4u;     // blocks per tile: a multiple of the kernel's JB (8 or 4)
    if (max_batch == 0) max_batch = (size_t)1 << 22;
    max_batch = std::max(max_batch, (size_t)4 * K);
    q->max_batch = max_batch;
    int rc = B2_OK;
    do {
        // many channels: the synchroniser chains get their own SMs (smpart.cu); B2_SYNC_SMS sizes the set
        cudaStream_t decode_streams[SyncCore::NDS] = {};
        bool have_decode_streams = false;
        // (the frame-parallel synchroniser is a throughput kernel like the channelizer: the two simply share the machine)
        const bool wk = syncw_supported(M) && getenv("B2_SYNC_LEGACY") == nullptr && getenv("B2_SYNC_GENERIC") == nullptr;
        if (N >= 32 && sync8_supported(M) && K >= 64 && (!wk || getenv("B2_SM_PARTITION") != nullptr)) {
            // one scheduler per chain warp is all a chain can use (both workers of a pair counted); beyond that the
            // SMs serve the channelizer and the packet decoder better.  72 = measured optimum of the 256 x 512 shape.
            const unsigned int chain_warps = 2u * N * std::max(1u, M / 256u);
            unsigned int want = std::min(72u, std::max(8u, ((chain_warps + 3) / 4 + 7) / 8 * 8));
            if (const char * e = getenv("B2_SYNC_SMS")) { long v = atol(e); if (v >= 8 && v <= 136) want = (unsigned int)v; }
            // every chain must be resident at once (a chain that waits for an SM stalls the pipeline):
            // the synchroniser kernel fits 4 CTAs of M/8 <= 64 threads per SM
            want = std::max(want, std::min(136u, ((N + 3) / 4 + 7) / 8 * 8));
            if (sm_partition_create(q->part, device, want)) {
                q->stream = q->part.big_stream[0];
                q->sstream = q->part.small_stream;
                // packet decode runs beside the channelizer (beside the synchronisers it disturbs the chains: measured)
                for (unsigned int i = 0; i < SyncCore::NDS; i++) decode_streams[i] = q->part.big_stream[1 + i];
                have_decode_streams = true;
                q->spare_stream = q->part.small_stream2;
            }
        }
        if (!q->stream && cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        // firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, 2N, m=7, As=60): lib/multichannelrx.cc:89-91
        std::vector<float> h = firpfbch_prototype(K, 7, 60.0f);
        fft_plan(q->fftK, K);
        if ((rc = q->t_taps.upload(h)) || (rc = q->t_perm.upload(q->fftK.perm)) || (rc = q->t_tw.upload(q->fftK.tw))) break;
        q->hist_len = (size_t)(q->P - 1) * K;
        if ((rc = q->d_stage.alloc(sizeof(cf) * (q->hist_len + K + max_batch + K)))) break;
        if ((rc = q->d_tail.alloc(sizeof(cf) * (q->hist_len + K)))) break;
        q->tcap = ((max_batch + K) / K + 2 + 1) & ~(size_t)1;
        if ((rc = q->d_chan.alloc(sizeof(cf) * q->tcap * N))) break;
        // NCO: lib/multichannelrx.cc:98-100
        float offset = -0.5f * (float)(N - 1) / (float)N * M_PI;
        q->nco_dtheta = nco_constrain(offset);
        q->nco_theta = 0;
        AnalyzerParams ap;
        memset(&ap, 0, sizeof(ap));
        ap.K = K; ap.P = q->P; ap.TB = q->TB;
        q->an_smem = analyzer_smem_bytes(ap);
        if (q->an_smem > 227 * 1024) { rc = b2_fail(B2_ERR_UNSUPPORTED, "channelizer tile does not fit shared memory"); break; }
        if (analyzer_configure(q->an_smem) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (q->part.ok) sms = (int)q->part.big_sms;
        if (const char * e = getenv("B2_AN_SMS")) { long v = atol(e); if (v >= 8 && v < sms) sms = (int)v; }
        int per_sm = std::max(1, (int)((227 * 1024) / (q->an_smem + 1024)));
        q->an_grid = sms * std::min(per_sm, 2);
        q->an_sms = (unsigned int)sms;
        // (experiment knob B2_SYNC_PRIORITY=1: the synchroniser launches, which depend on each other through the carried
        // state, on a high-priority stream -- measured: no gain, the long-lived channelizer CTAs hold their SMs anyway)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const bool prio = getenv("B2_SYNC_PRIORITY") != nullptr && atoi(getenv("B2_SYNC_PRIORITY")) != 0;
        if (cudaStreamCreateWithFlags(&q->cstream, cudaStreamNonBlocking) != cudaSuccess ||
            (!q->sstream && (prio ? cudaStreamCreateWithPriority(&q->sstream, cudaStreamNonBlocking, prio_hi)
                                  : cudaStreamCreateWithFlags(&q->sstream, cudaStreamNonBlocking)) != cudaSuccess) ||
            cudaEventCreate(&q->ev_begin) != cudaSuccess || cudaEventCreate(&q->ev_end) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        // chunk of the pipeline: long enough to amortise launches, short enough to overlap stages
        q->chunk_blocks = std::max(64u, (1u << 25) / K);
        if (const char * e = getenv("B2_CHUNK_BLOCKS")) { long v = atol(e); if (v >= 1) q->chunk_blocks = (unsigned int)v; }
        if (q->part.ok) q->core.sm_budget = (unsigned int)q->part.small_sms;
        if ((rc = q->core.init(M, cp, taper, p, N, q->tcap, device, q->sstream, have_decode_streams ? decode_streams : nullptr))) break;
        B2_CUDA(cudaMemsetAsync(q->d_stage.p, 0, q->d_stage.bytes, q->stream));
        B2_CUDA(cudaStreamSynchronize(q->stream));
    } while (0);
    if (rc) { b2_mcrx_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}

extern "C" int b2_mcrx_destroy(b2_mcrx * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    if (q->stream) cudaStreamSynchronize(q->stream);
    if (q->sstream) cudaStreamSynchronize(q->sstream);
    if (q->cstream) cudaStreamSynchronize(q->cstream);
    q->core.destroy();
    if (q->rs) b2_msresamp_destroy(q->rs);
    for (auto & e : q->aev) { cudaEventDestroy(e.copied); cudaEventDestroy(e.a0); cudaEventDestroy(e.a1); }
    if (q->ev_begin) cudaEventDestroy(q->ev_begin);
    if (q->ev_end) cudaEventDestroy(q->ev_end);
    if (q->stream) cudaStreamDestroy(q->stream);
    if (q->sstream) cudaStreamDestroy(q->sstream);
    if (q->cstream) cudaStreamDestroy(q->cstream);
    if (q->spare_stream) cudaStreamDestroy(q->spare_stream);
    delete q;
    return B2_OK;
}

// multichannelrx::Reset (lib/multichannelrx.cc:135-153): framesyncs + channelizer windows +
// buffer_index; the NCO keeps running (its reset is commented out at :144)
extern "C" int b2_mcrx_reset(b2_mcrx * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    B2_CUDA(cudaMemsetAsync(q->d_stage.p, 0, sizeof(cf) * (q->hist_len + q->K), q->stream));
    B2_CUDA(cudaStreamSynchronize(q->stream));
    q->carry = 0;
    if (q->rs) B2_TRY(b2_msresamp_reset(q->rs));
    return q->core.reset_streams();
}

// msresamp_crcf ahead of multichannelrx::Execute: the rate-matching step the reference's programs
// compute but never apply on this path (src/multichannel_rx.cc:137-138, TODO at lib/multichanneltxrx.cc:605)
extern "C" int b2_mcrx_set_resampler(b2_mcrx * q, float rate, float As)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    if (q->rs) { b2_msresamp_destroy(q->rs); q->rs = nullptr; }
    if (rate == 0.0f) return B2_OK;
    B2_TRY(b2_msresamp_create(rate, As, q->device, &q->rs));
    // one resampler call fills at most one receiver batch
    q->rs_cap = q->max_batch;
    q->rs_in_chunk = std::max<size_t>(1, (size_t)((double)q->max_batch / ((double)rate * 1.001))) ;
    if (q->rs_in_chunk > 8) q->rs_in_chunk -= 4;
    if (q->d_rs.bytes < sizeof(cf) * q->rs_cap) {
        int rc = q->d_rs.alloc(sizeof(cf) * q->rs_cap);
        if (rc) { b2_msresamp_destroy(q->rs); q->rs = nullptr; return rc; }
    }
    return B2_OK;
}

static int mcrx_process(b2_mcrx * q, const float * x, size_t n, bool on_device)
{
    const unsigned int K = q->K;
    cf * stage = q->d_stage.as<cf>();
    const size_t front = q->hist_len + q->carry;             // samples already at the front of the stage
    const size_t total = q->carry + n;
    const size_t T = total / K, leftover = total % K;
    const bool direct = on_device && q->carry == 0 && T > 0 && (((uintptr_t)x) & 15) == 0;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    B2_CUDA(cudaEventRecord(q->ev_begin, q->stream));
    AnalyzerParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.seg0 = stage;
    ap.rows0 = direct ? q->P - 1 : 0xffffffffu;
    ap.seg1 = direct ? (const cf *)x : stage;
    ap.K = K; ap.lgK = q->lgK; ap.N = q->N; ap.P = q->P; ap.TB = q->TB;
    ap.sm_limit = q->an_sms;
    ap.taps = q->t_taps.as<float>();
    ap.dtheta = q->nco_dtheta;
    ap.theta0 = q->nco_theta - (uint32_t)(q->hist_len + q->carry) * q->nco_dtheta;
    ap.out = q->d_chan.as<cf>(); ap.out_stride = q->tcap;
    ap.fft.n = K; ap.fft.npass = q->fftK.npass;
    ap.fft.radices = 0;
    for (unsigned int i = 0; i < q->fftK.npass; i++) ap.fft.radices |= fft_radix_code(q->fftK.radix[i]) << (4 * i);
    ap.fft.perm = q->t_perm.as<uint16_t>(); ap.fft.tw = q->t_tw.as<cf>();

    size_t copied = 0;                                       // samples of x already on their way into the stage
    unsigned int nchunks = 0;
    if (T > 0) B2_TRY(q->core.begin_batch());
    // chunk schedule: short chunks first (the synchronisers start early), doubling up to chunk_blocks, and
    // halving again towards the end (the D2H + host ordering of the last chunk is the tail of the call)
    // (the frame-parallel synchroniser wants launches of many frames per stream and has no chain to start early: equal chunks)
    size_t cb = q->chunk_blocks, cmin = q->core.use_w ? cb : std::max<size_t>(64, cb / 8);
    if (const char * e = getenv("B2_CHUNK_MIN_BLOCKS")) { long v = atol(e); if (v >= 1) cmin = std::min<size_t>(cb, (size_t)v); }
    size_t next_tc = cmin;
    for (size_t b0 = 0, tc = 0; b0 < T; b0 += tc, nchunks++) {
        const size_t left = T - b0;
        tc = next_tc;
        while (tc > cmin && left < 2 * tc - cmin) tc /= 2;   // ramp down: ... cb/2, cb/4, cb/8 fit in what is left
        tc = std::min(tc, left);
        if (left - tc < cmin / 2) tc = left;                  // no tiny tail chunk
        next_tc = std::min(cb, next_tc * 2);
        const bool last = b0 + tc == T;
        if (q->aev.size() <= nchunks) {
            b2_mcrx_s::AnEv e;
            B2_CUDA(cudaEventCreateWithFlags(&e.copied, cudaEventDisableTiming));
            B2_CUDA(cudaEventCreate(&e.a0)); B2_CUDA(cudaEventCreate(&e.a1));
            q->aev.push_back(e);
        }
        b2_mcrx_s::AnEv & e = q->aev[nchunks];
        if (!direct) {
            // samples this chunk's blocks need (the last chunk also brings the leftover)
            const size_t upto = last ? n : (b0 + tc) * K - q->carry;
            if (upto > copied) {
                B2_CUDA(cudaMemcpyAsync(stage + front + copied, (const cf *)x + copied, sizeof(cf) * (upto - copied), kind, q->cstream));
                copied = upto;
            }
            B2_CUDA(cudaEventRecord(e.copied, q->cstream));
            B2_CUDA(cudaStreamWaitEvent(q->stream, e.copied, 0));
        }
        ap.block0 = (unsigned int)b0; ap.nblocks = (unsigned int)tc; ap.out_col0 = b0;
        if (q->core.timing) B2_CUDA(cudaEventRecord(e.a0, q->stream));
        B2_CUDA(analyzer_launch(ap, q->an_grid, q->an_smem, q->stream));
        B2_CUDA(cudaEventRecord(e.a1, q->stream));
        q->core.launches++;
        B2_TRY(q->core.launch_chunk(q->d_chan.as<cf>() + b0, q->tcap, (unsigned int)tc, e.a1));
    }
    if (T == 0 && n > 0) {
        B2_CUDA(cudaMemcpyAsync(stage + front, x, sizeof(cf) * n, kind, q->cstream));
        B2_CUDA(cudaStreamSynchronize(q->cstream));
    }
    q->last_blocks = (unsigned int)T;
    // keep the last (P-1)*K + leftover samples of the stream at the front of the stage
    {
        const size_t keep = q->hist_len + leftover;
        const size_t stream_len = front + n;                  // logical samples in [stage front | new]
        if (direct) {
            // logical stream = stage[0..hist_len) ++ x[0..n)
            if (n >= keep) {
                B2_CUDA(cudaMemcpyAsync(stage, (const cf *)x + (n - keep), sizeof(cf) * keep, cudaMemcpyDeviceToDevice, q->stream));
            } else {
                cf * tail = q->d_tail.as<cf>();
                size_t from_hist = keep - n;
                B2_CUDA(cudaMemcpyAsync(tail, stage + (q->hist_len - from_hist), sizeof(cf) * from_hist, cudaMemcpyDeviceToDevice, q->stream));
                B2_CUDA(cudaMemcpyAsync(tail + from_hist, x, sizeof(cf) * n, cudaMemcpyDeviceToDevice, q->stream));
                B2_CUDA(cudaMemcpyAsync(stage, tail, sizeof(cf) * keep, cudaMemcpyDeviceToDevice, q->stream));
            }
        } else if (T > 0) {
            const cf * src = stage + (stream_len - keep);
            if (stream_len - keep >= keep) {
                B2_CUDA(cudaMemcpyAsync(stage, src, sizeof(cf) * keep, cudaMemcpyDeviceToDevice, q->stream));
            } else {
                cf * tail = q->d_tail.as<cf>();
                B2_CUDA(cudaMemcpyAsync(tail, src, sizeof(cf) * keep, cudaMemcpyDeviceToDevice, q->stream));
                B2_CUDA(cudaMemcpyAsync(stage, tail, sizeof(cf) * keep, cudaMemcpyDeviceToDevice, q->stream));
            }
        }
        q->carry = leftover;
        q->nco_theta += (uint32_t)n * q->nco_dtheta;
    }
    if (T > 0) B2_TRY(q->core.end_batch());
    B2_CUDA(cudaStreamSynchronize(q->stream));
    B2_CUDA(cudaEventRecord(q->ev_end, q->stream));
    B2_CUDA(cudaEventSynchronize(q->ev_end));
    if (T > 0) {
        q->timed_chunks = nchunks;
        if (getenv("B2_DUMP_TIMELINE")) { mcrx_fetch_timing(q); }
        if (q->core.timing && getenv("B2_DUMP_TIMELINE")) {
            // per chunk: begin/end of the channelizer, synchroniser and decode kernels, ms since the call began
            for (unsigned int i = 0; i < nchunks; i++) {
                float t[6] = {0, 0, 0, 0, 0, 0};
                cudaEventElapsedTime(&t[0], q->ev_begin, q->aev[i].a0); cudaEventElapsedTime(&t[1], q->ev_begin, q->aev[i].a1);
                cudaEventElapsedTime(&t[2], q->ev_begin, q->core.cev[i].s0); cudaEventElapsedTime(&t[3], q->ev_begin, q->core.cev[i].s1);
                cudaEventElapsedTime(&t[4], q->ev_begin, q->core.cev[i].d0); cudaEventElapsedTime(&t[5], q->ev_begin, q->core.cev[i].d1);
                fprintf(stderr, "chunk %3u  channelizer %7.3f-%7.3f  sync %7.3f-%7.3f  decode %7.3f-%7.3f\n", i, t[0], t[1], t[2], t[3], t[4], t[5]);
            }
            fprintf(stderr, "call %7.3f ms\n", q->core.last_ms[3]);
        }
    }
    return B2_OK;
}

static int mcrx_execute_any(b2_mcrx * q, const float * x, size_t n, bool on_device)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n == 0) return B2_OK;
    if (!x) return b2_fail(B2_ERR_ARG, "null sample pointer");
    B2_CUDA(cudaSetDevice(q->device));
    size_t done = 0;
    while (done < n) {
        if (q->rs) {
            // resample a piece into device memory, then run the receiver on it in place
            size_t c = std::min(n - done, q->rs_in_chunk), ny = 0;
            float * y = q->d_rs.as<float>();
            B2_TRY(on_device ? b2_msresamp_execute_device(q->rs, x + 2 * done, c, y, q->rs_cap, &ny)
                             : b2_msresamp_execute_to_device(q->rs, x + 2 * done, c, y, q->rs_cap, &ny));
            if (ny) B2_TRY(mcrx_process(q, y, ny, true));
            B2_CUDA(cudaStreamSynchronize(q->stream));      // d_rs is rewritten by the next piece
            done += c;
            continue;
        }
        size_t c = std::min(n - done, q->max_batch);
        int rc = mcrx_process(q, x + 2 * done, c, on_device);
        if (rc) return rc;
        done += c;
    }
    return B2_OK;
}
extern "C" int b2_mcrx_execute(b2_mcrx * q, const float * x_host, size_t n) { return mcrx_execute_any(q, x_host, n, false); }
extern "C" int b2_mcrx_execute_device(b2_mcrx * q, const float * x_dev, size_t n) { return mcrx_execute_any(q, x_dev, n, true); }

extern "C" int b2_mcrx_poll(b2_mcrx * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                            uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll(recs, recs_cap, n_recs, payloads, payloads_cap, n_payload_bytes);
}

extern "C" int b2_mcrx_poll_view(b2_mcrx * q, const b2_frame_rec ** recs, size_t * n_recs,
                                 const uint8_t ** payloads, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll_view(recs, n_recs, payloads, n_payload_bytes);
}

extern "C" int b2_mcrx_tap_symbols(b2_mcrx * q, int enable, size_t max_symbols)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    return q->core.set_tap(enable, max_symbols);
}

static int core_read_symbols(SyncCore & c, uint32_t * channel, uint64_t * index, float * X, size_t cap, size_t * n)
{
    size_t have = c.tap_chan.size();
    if (n) *n = have;
    if (!channel) return B2_OK;
    if (cap < have) return b2_fail(B2_ERR_OVERFLOW, "symbol buffers too small");
    if (have) {
        memcpy(channel, c.tap_chan.data(), have * sizeof(uint32_t));
        memcpy(index, c.tap_index.data(), have * sizeof(uint64_t));
        memcpy(X, c.tap_X.data(), c.tap_X.size() * sizeof(float));
    }
    c.tap_chan.clear(); c.tap_index.clear(); c.tap_X.clear();
    return B2_OK;
}
extern "C" int b2_mcrx_read_symbols(b2_mcrx * q, uint32_t * channel, uint64_t * index, float * X, size_t cap, size_t * n)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return core_read_symbols(q->core, channel, index, X, cap, n);
}

extern "C" int b2_mcrx_last_timing(b2_mcrx * q, float ms[4])
{
    if (!q || !ms) return b2_fail(B2_ERR_ARG, "null argument");
    cudaSetDevice(q->device);
    mcrx_fetch_timing(q);
    for (int i = 0; i < 4; i++) ms[i] = q->core.last_ms[i];
    return B2_OK;
}

extern "C" int b2_mcrx_last_launches(b2_mcrx * q, unsigned int * kernels, unsigned int * chunks)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (kernels) *kernels = q->core.launches;
    if (chunks) *chunks = q->core.chunk;
    return B2_OK;
}

extern "C" int b2_mcrx_read_channelizer(b2_mcrx * q, float * out, size_t cap_samples, size_t * n_blocks)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n_blocks) *n_blocks = q->last_blocks;
    if (!out) return B2_OK;
    if (cap_samples < (size_t)q->last_blocks * q->N) return b2_fail(B2_ERR_OVERFLOW, "buffer too small");
    B2_CUDA(cudaSetDevice(q->device));
    if (q->last_blocks)
        B2_CUDA(cudaMemcpy2D(out, sizeof(cf) * q->last_blocks, q->d_chan.p, sizeof(cf) * q->tcap,
                             sizeof(cf) * q->last_blocks, q->N, cudaMemcpyDeviceToHost));
    return B2_OK;
}

extern "C" void * b2_mcrx_stream(b2_mcrx * q) { return q ? (void *)q->stream : nullptr; }

// stage 1 alone on a time shard; `peers` != nullptr: channel c is written into peers[c / cpp] (multi-GPU split)
static int mcrx_channelize(b2_mcrx * q, const float * x_dev, size_t n_blocks, int64_t sample_offset, float * out_dev,
                           size_t out_stride, size_t out_col0, cf * const * peers, unsigned int n_peer, unsigned int cpp, cudaStream_t st)
{
    AnalyzerParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.seg0 = (const cf *)x_dev; ap.rows0 = 0xffffffffu; ap.seg1 = (const cf *)x_dev;
    ap.K = q->K; ap.lgK = q->lgK; ap.N = q->N; ap.P = q->P; ap.TB = q->TB;
    ap.sm_limit = q->an_sms;
    ap.nblocks = (unsigned int)n_blocks;
    ap.taps = q->t_taps.as<float>();
    ap.dtheta = q->nco_dtheta;
    ap.theta0 = (uint32_t)((uint64_t)sample_offset) * q->nco_dtheta;
    ap.out = (cf *)out_dev; ap.out_stride = out_stride; ap.out_col0 = out_col0;
    if (peers) {
        for (unsigned int i = 0; i < n_peer && i < 8; i++) ap.out_peer[i] = peers[i];
        ap.n_peer = n_peer; ap.chan_per_peer = cpp;
    }
    ap.fft.n = q->K; ap.fft.npass = q->fftK.npass; ap.fft.radices = 0;
    for (unsigned int i = 0; i < q->fftK.npass; i++) ap.fft.radices |= fft_radix_code(q->fftK.radix[i]) << (4 * i);
    ap.fft.perm = q->t_perm.as<uint16_t>(); ap.fft.tw = q->t_tw.as<cf>();
    B2_CUDA(analyzer_launch(ap, q->an_grid, q->an_smem, st));
    return B2_OK;
}

extern "C" int b2_mcrx_channelize_device(b2_mcrx * q, const float * x_dev, size_t n_blocks, int64_t sample_offset,
                                         float * out_dev, size_t out_stride)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n_blocks == 0) return B2_OK;
    if (!x_dev || !out_dev) return b2_fail(B2_ERR_ARG, "null pointer");
    if (((uintptr_t)x_dev) & 15) return b2_fail(B2_ERR_ARG, "input must be 16-byte aligned");
    if (n_blocks > 0x7fffffffu) return b2_fail(B2_ERR_ARG, "too many blocks in one call");
    B2_CUDA(cudaSetDevice(q->device));
    return mcrx_channelize(q, x_dev, n_blocks, sample_offset, out_dev, out_stride, 0, nullptr, 0, 0, q->stream);
}

extern "C" int b2_mcrx_sync_device(b2_mcrx * q, const float * in_dev, size_t n, size_t in_stride)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n == 0) return B2_OK;
    if (!in_dev) return b2_fail(B2_ERR_ARG, "null pointer");
    B2_CUDA(cudaSetDevice(q->device));
    B2_CUDA(cudaStreamSynchronize(q->stream));               // a preceding b2_mcrx_channelize_device
    size_t done = 0;
    while (done < n) {
        size_t c = std::min(n - done, q->core.tmax);
        int rc = q->core.run((const cf *)in_dev + done, in_stride, (unsigned int)c, true);
        if (rc) return rc;
        done += c;
    }
    return B2_OK;
}

// ================================================================== batched single-link synchroniser
struct b2_ofdmsync_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int streams = 0;
    size_t max_batch = 0;
    DevBuf d_in;
    SyncCore core;
};

extern "C" int b2_ofdmsync_create(unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p,
                                  unsigned int streams, int device, size_t max_batch, b2_ofdmsync ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (streams < 1) return b2_fail(B2_ERR_ARG, "need at least one stream");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return b2_fail(B2_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return b2_fail(B2_ERR_ARG, "invalid device ordinal %d", device);
    B2_CUDA(cudaSetDevice(device));
    b2_ofdmsync * q = new b2_ofdmsync_s;
    q->device = device; q->streams = streams;
    if (max_batch == 0) max_batch = (size_t)1 << 20;
    q->max_batch = max_batch;
    int rc = B2_OK;
    do {
        if (cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        if ((rc = q->d_in.alloc(sizeof(cf) * max_batch * streams))) break;
        if ((rc = q->core.init(M, cp, taper, p, streams, max_batch, device, q->stream))) break;
    } while (0);
    if (rc) { b2_ofdmsync_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}
extern "C" int b2_ofdmsync_destroy(b2_ofdmsync * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    if (q->stream) cudaStreamSynchronize(q->stream);
    q->core.destroy();
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
    return B2_OK;
}
extern "C" int b2_ofdmsync_reset(b2_ofdmsync * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    return q->core.reset_streams();
}
static int ofdmsync_run(b2_ofdmsync * q, const cf * in, size_t stride, size_t n)
{
    B2_CUDA(cudaEventRecord(q->core.ev[0], q->stream));
    int rc = q->core.run(in, stride, (unsigned int)n, true);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(q->core.ev[4], q->stream));
    B2_CUDA(cudaEventSynchronize(q->core.ev[4]));
    q->core.last_ms[0] = 0.f;
    cudaEventElapsedTime(&q->core.last_ms[3], q->core.ev[0], q->core.ev[4]);
    return B2_OK;
}
extern "C" int b2_ofdmsync_execute(b2_ofdmsync * q, const float * x_host, size_t n)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n == 0) return B2_OK;
    if (!x_host) return b2_fail(B2_ERR_ARG, "null sample pointer");
    B2_CUDA(cudaSetDevice(q->device));
    size_t done = 0;
    while (done < n) {
        size_t c = std::min(n - done, q->max_batch);
        B2_CUDA(cudaMemcpy2DAsync(q->d_in.p, sizeof(cf) * c, x_host + 2 * done, sizeof(cf) * n, sizeof(cf) * c, q->streams,
                                  cudaMemcpyHostToDevice, q->stream));
        int rc = ofdmsync_run(q, q->d_in.as<cf>(), c, c);
        if (rc) return rc;
        done += c;
    }
    return B2_OK;
}
extern "C" int b2_ofdmsync_execute_device(b2_ofdmsync * q, const float * x_dev, size_t n, size_t stride)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n == 0) return B2_OK;
    if (!x_dev) return b2_fail(B2_ERR_ARG, "null sample pointer");
    B2_CUDA(cudaSetDevice(q->device));
    size_t done = 0;
    while (done < n) {
        size_t c = std::min(n - done, q->max_batch);
        int rc = ofdmsync_run(q, (const cf *)x_dev + done, stride, c);
        if (rc) return rc;
        done += c;
    }
    return B2_OK;
}
extern "C" int b2_ofdmsync_poll(b2_ofdmsync * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                                uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll(recs, recs_cap, n_recs, payloads, payloads_cap, n_payload_bytes);
}
extern "C" int b2_ofdmsync_poll_view(b2_ofdmsync * q, const b2_frame_rec ** recs, size_t * n_recs,
                                     const uint8_t ** payloads, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll_view(recs, n_recs, payloads, n_payload_bytes);
}
extern "C" int b2_ofdmsync_last_timing(b2_ofdmsync * q, float ms[4])
{
    if (!q || !ms) return b2_fail(B2_ERR_ARG, "null argument");
    cudaSetDevice(q->device);
    q->core.fetch_timing();
    for (int i = 0; i < 4; i++) ms[i] = q->core.last_ms[i];
    return B2_OK;
}

// ================================================================== one rank of a multi-GPU multichannelrx
struct b2_mcrx_shard_s {
    int device = 0;
    unsigned int N = 0, K = 0, rank = 0, world = 1, cpp = 0;
    size_t tc = 0, steps = 0, row = 0;   // chunk blocks, steps per call, row length (world * tc) of an exchange slot
    cudaStream_t s1 = nullptr, s2 = nullptr;
    b2_mcrx * chan = nullptr;            // stage 1: tables and launch configuration of a full N-channel receiver
    SyncCore core;                       // stage 2: N / world streams
    DevBuf d_xchg;                       // [B2_SHARD_SLOTS][N / world][world * tc]
    // B2_SHARD_COPY=1: stage 1 writes a local tile and the copy engines carry the slabs to their owners (instead of the
    // channelizer storing into peer memory itself)
    bool copy_mode = false;
    DevBuf d_tile;                       // [N][tc]
    cudaStream_t cpy = nullptr;
    cudaEvent_t ev_tile = nullptr, ev_copied = nullptr;
    cf * peer[8] = {};                   // every rank's exchange buffer as seen from this device
    bool connected = false, in_call = false;
};

extern "C" int b2_mcrx_shard_destroy(b2_mcrx_shard * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    cudaDeviceSynchronize();
    for (unsigned int i = 0; i < q->world && i < 8; i++)
        if (q->connected && i != q->rank && q->peer[i]) cudaIpcCloseMemHandle(q->peer[i]);
    q->core.destroy();
    if (q->chan) b2_mcrx_destroy(q->chan);
    if (q->cpy) cudaStreamDestroy(q->cpy);
    if (q->ev_tile) cudaEventDestroy(q->ev_tile);
    if (q->ev_copied) cudaEventDestroy(q->ev_copied);
    delete q;
    return B2_OK;
}

extern "C" int b2_mcrx_shard_create(unsigned int N, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p,
                                    int device, unsigned int rank, unsigned int world, size_t chunk_blocks, size_t steps_per_call,
                                    void * stream_stage1, void * stream_stage2, b2_mcrx_shard ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (world < 1 || world > 8 || rank >= world) return b2_fail(B2_ERR_ARG, "rank %u of %u: at most 8 ranks", rank, world);
    if (N % world) return b2_fail(B2_ERR_ARG, "the number of channels (%u) must divide by the number of ranks (%u)", N, world);
    if (chunk_blocks < 64 || steps_per_call < 1) return b2_fail(B2_ERR_ARG, "chunk_blocks >= 64 and steps_per_call >= 1 required");
    if ((uint64_t)chunk_blocks * world > 0x3fffffffull) return b2_fail(B2_ERR_ARG, "chunk too long");
    b2_mcrx_shard * q = new b2_mcrx_shard_s;
    q->device = device; q->N = N; q->K = 2 * N; q->rank = rank; q->world = world; q->cpp = N / world;
    q->tc = chunk_blocks; q->steps = steps_per_call; q->row = chunk_blocks * world;
    q->s1 = (cudaStream_t)stream_stage1; q->s2 = (cudaStream_t)stream_stage2;
    int rc = B2_OK;
    do {
        if ((rc = b2_mcrx_create(N, M, cp, taper, p, device, 4 * (size_t)q->K, &q->chan))) break;
        if ((rc = q->d_xchg.alloc(sizeof(cf) * (size_t)B2_SHARD_SLOTS * q->cpp * q->row))) break;
        B2_CUDA(cudaMemset(q->d_xchg.p, 0, q->d_xchg.bytes));
        q->peer[rank] = q->d_xchg.as<cf>();
        // stage 2: one batch = one call = steps_per_call launches of world * chunk_blocks samples per stream
        if ((rc = q->core.init(M, cp, taper, p, q->cpp, q->row * steps_per_call, device, q->s2))) break;
        q->core.sp.chan_base = rank * q->cpp;                           // records carry the global channel index
        if (const char * e = getenv("B2_SHARD_COPY")) q->copy_mode = atoi(e) != 0;
        // stage 1 is bound by NVLink when most of its output goes to peers: it does not need every SM, and the ones
        // it leaves run the synchronisers of the previous step at full occupancy (B2_SHARD_AN_SMS overrides)
        {
            unsigned int an = q->chan->an_sms;
            if (world >= 4 && an > 112) an = 112;
            if (const char * e = getenv("B2_SHARD_AN_SMS")) { long v = atol(e); if (v >= 8 && v <= 1024) an = (unsigned int)v; }
            if (an < q->chan->an_sms) q->chan->an_sms = an;
        }
        if (q->copy_mode) {
            if ((rc = q->d_tile.alloc(sizeof(cf) * (size_t)N * q->tc))) break;
            B2_CUDA(cudaStreamCreateWithFlags(&q->cpy, cudaStreamNonBlocking));
            B2_CUDA(cudaEventCreateWithFlags(&q->ev_tile, cudaEventDisableTiming));
            B2_CUDA(cudaEventCreateWithFlags(&q->ev_copied, cudaEventDisableTiming));
        }
        q->connected = (world == 1);
    } while (0);
    if (rc) { b2_mcrx_shard_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}

extern "C" int b2_mcrx_shard_export(b2_mcrx_shard * q, void * handle64)
{
    if (!q || !handle64) return b2_fail(B2_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    B2_CUDA(cudaSetDevice(q->device));
    cudaIpcMemHandle_t h;
    B2_CUDA(cudaIpcGetMemHandle(&h, q->d_xchg.p));
    memcpy(handle64, &h, 64);
    return B2_OK;
}

extern "C" int b2_mcrx_shard_connect(b2_mcrx_shard * q, const void * handles, size_t n_handles)
{
    if (!q || !handles) return b2_fail(B2_ERR_ARG, "null argument");
    if (n_handles != q->world) return b2_fail(B2_ERR_ARG, "expected %u handles, got %zu", q->world, n_handles);
    B2_CUDA(cudaSetDevice(q->device));
    for (unsigned int i = 0; i < q->world; i++) {
        if (i == q->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)i, 64);
        void * ptr = nullptr;
        B2_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        q->peer[i] = (cf *)ptr;
    }
    q->connected = true;
    return B2_OK;
}

extern "C" int b2_mcrx_shard_begin(b2_mcrx_shard * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (!q->connected) return b2_fail(B2_ERR_ARG, "b2_mcrx_shard_connect has not been called");
    B2_CUDA(cudaSetDevice(q->device));
    B2_TRY(q->core.begin_batch());
    q->in_call = true;
    return B2_OK;
}

extern "C" int b2_mcrx_shard_stage1(b2_mcrx_shard * q, const float * x_dev, uint64_t step)
{
    if (!q || !x_dev) return b2_fail(B2_ERR_ARG, "null argument");
    if (!q->connected) return b2_fail(B2_ERR_ARG, "b2_mcrx_shard_connect has not been called");
    if (((uintptr_t)x_dev) & 15) return b2_fail(B2_ERR_ARG, "input must be 16-byte aligned");
    B2_CUDA(cudaSetDevice(q->device));
    const uint64_t chunk = step * q->world + q->rank;                   // absolute chunk index of the stream
    const int64_t offset = ((int64_t)(chunk * q->tc) - B2_SHARD_HALO_BLOCKS) * (int64_t)q->K;
    // slot base of every peer; this rank's chunk lands in columns [rank * tc, (rank + 1) * tc) of the slot's rows
    cf * dst[8];
    const size_t slot_off = (size_t)(step % B2_SHARD_SLOTS) * q->cpp * q->row;
    for (unsigned int i = 0; i < q->world; i++) dst[i] = q->peer[i] + slot_off;
    if (q->copy_mode) {
        // local tile, then one strided copy per owner on the copy stream; stream_stage1 continues behind the copies
        B2_CUDA(cudaStreamWaitEvent(q->s1, q->ev_copied, 0));            // the previous step's copies have read the tile
        B2_TRY(mcrx_channelize(q->chan, x_dev, q->tc, offset, q->d_tile.as<float>(), q->tc, 0, nullptr, 0, 0, q->s1));
        B2_CUDA(cudaEventRecord(q->ev_tile, q->s1));
        B2_CUDA(cudaStreamWaitEvent(q->cpy, q->ev_tile, 0));
        for (unsigned int k = 0; k < q->world; k++) {
            const unsigned int i = (q->rank + k) % q->world;             // everybody starts with a different owner
            B2_CUDA(cudaMemcpy2DAsync(dst[i] + (size_t)q->rank * q->tc, sizeof(cf) * q->row,
                                      q->d_tile.as<cf>() + (size_t)i * q->cpp * q->tc, sizeof(cf) * q->tc,
                                      sizeof(cf) * q->tc, q->cpp, cudaMemcpyDeviceToDevice, q->cpy));
        }
        B2_CUDA(cudaEventRecord(q->ev_copied, q->cpy));
        B2_CUDA(cudaStreamWaitEvent(q->s1, q->ev_copied, 0));
        return B2_OK;
    }
    return mcrx_channelize(q->chan, x_dev, q->tc, offset, nullptr, q->row, (size_t)q->rank * q->tc, dst, q->world, q->cpp, q->s1);
}

extern "C" int b2_mcrx_shard_stage2(b2_mcrx_shard * q, uint64_t step)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (!q->in_call) return b2_fail(B2_ERR_ARG, "b2_mcrx_shard_begin has not been called");
    B2_CUDA(cudaSetDevice(q->device));
    const cf * in = q->d_xchg.as<cf>() + (size_t)(step % B2_SHARD_SLOTS) * q->cpp * q->row;
    return q->core.launch_chunk(in, q->row, (unsigned int)q->row, nullptr);
}

extern "C" int b2_mcrx_shard_end(b2_mcrx_shard * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (!q->in_call) return B2_OK;
    B2_CUDA(cudaSetDevice(q->device));
    q->in_call = false;
    return q->core.end_batch();
}

extern "C" int b2_mcrx_shard_poll_view(b2_mcrx_shard * q, const b2_frame_rec ** recs, size_t * n_recs,
                                       const uint8_t ** payloads, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll_view(recs, n_recs, payloads, n_payload_bytes);
}

extern "C" int b2_mcrx_shard_pack_results(b2_mcrx_shard * q, void * dst_dev, size_t cap_bytes, uint64_t seq, size_t * n_recs, size_t * n_payload_bytes)
{
    if (!q || !dst_dev) return b2_fail(B2_ERR_ARG, "null argument");
    B2_CUDA(cudaSetDevice(q->device));
    SyncCore & c = q->core;
    const size_t H = B2_SHARD_PACK_HEADER;
    const size_t nr = c.chunk ? std::min<size_t>(c.h_range[c.chunk].nrec, c.recs_cap) : 0;
    const size_t nb = c.chunk ? (size_t)c.last_used : 0;
    if (H + nr * sizeof(FrameRec) + nb > cap_bytes) return b2_fail(B2_ERR_OVERFLOW, "pack buffer too small (%zu records, %zu payload bytes)", nr, nb);
    // header: the sizes travel with the data (no separate size exchange); h_counters is pinned (32 bytes)
    unsigned long long * hdr = (unsigned long long *)c.h_counters;
    B2_CUDA(cudaStreamSynchronize(q->s2));                               // the previous header copy has been consumed
    hdr[0] = nr; hdr[1] = nb; hdr[2] = seq; hdr[3] = 0;
    B2_CUDA(cudaMemcpyAsync(dst_dev, hdr, H, cudaMemcpyHostToDevice, q->s2));
    if (nr) B2_CUDA(pack_sorted_launch(c.d_recs.as<FrameRec>(), (unsigned int)nr, (FrameRec *)((char *)dst_dev + H), q->s2));
    if (nb) B2_CUDA(cudaMemcpyAsync((char *)dst_dev + H + nr * sizeof(FrameRec), c.d_decoded.p, nb, cudaMemcpyDeviceToDevice, q->s2));
    if (n_recs) *n_recs = nr;
    if (n_payload_bytes) *n_payload_bytes = nb;
    return B2_OK;
}

extern "C" int b2_mcrx_shard_poll(b2_mcrx_shard * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                                  uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    return q->core.poll(recs, recs_cap, n_recs, payloads, payloads_cap, n_payload_bytes);
}

extern "C" int b2_mcrx_shard_host_results(b2_mcrx_shard * q, int on)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    q->core.host_results = (on != 0);
    return B2_OK;
}

extern "C" int b2_mcrx_shard_reset(b2_mcrx_shard * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    return q->core.reset_streams();
}

extern "C" int b2_memcpy_async(void * dst, const void * src, size_t bytes, void * stream)
{
    if (bytes == 0) return B2_OK;
    if (!dst || !src) return b2_fail(B2_ERR_ARG, "null pointer");
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return B2_OK;
}
