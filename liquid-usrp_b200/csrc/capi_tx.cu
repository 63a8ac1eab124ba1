// capi_tx.cu -- transmit side and resampler of the C ABI (include/b200_ofdm.h):
//   b2_mctx_*     multichanneltx     (lib/multichanneltx.cc:41-242)
//   b2_ofdmgen_*  ofdmflexframegen as used by ofdmtxrx (lib/ofdmtxrx.cc:79-84,314-328,377-387)
//   b2_msresamp_* msresamp_crcf      (src/flexframe_rx.cc:179,240; src/flexframe_tx.cc:170,237)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200_ofdm.h"
#include "design.h"
#include "kernels.h"
#include "capi_util.h"

using namespace b2;

// ================================================================== frame generator bank
// N frame generators sharing tables; host keeps the frame bookkeeping (which symbol each
// channel is at), the device keeps the taper postfix and the encoded symbols.
struct GenChanHost {
    bool assembled = false;
    bool fresh = false;                  // frame has not produced its first symbol yet
    unsigned int symbol = 0;             // next symbol of the frame
    unsigned int n_hdr = 0, n_pay = 0, mod = 0, bps = 0, payload_mod_len = 0;
    unsigned int total() const { return 3 + n_hdr + n_pay + 1; }
};

struct GenBank {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int N = 0;
    OfdmPlan plan;
    FftPlan fftM;
    DevBuf t_s0, t_s1, t_taper, t_rank, t_seq, t_perm, t_tw;
    DevBuf d_hmod, d_pmod, d_work0, d_work1, d_post, d_desc, d_jobs, d_pay;
    size_t mod_stride = 0, work_stride = 0;
    std::vector<GenChanHost> ch;
    std::vector<EncodeJob> jobs;
    std::vector<uint8_t> job_payloads;
    FramegenParams fp;

    int init(unsigned int N_, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p, int device_, cudaStream_t st);
    int ensure_capacity(unsigned int enc_len, unsigned int mod_len);
    int reset();
    int assemble(unsigned int c, const unsigned char * header, const unsigned char * payload, unsigned int len,
                 int check, int fec0, int fec1, int mod, unsigned int * n_symbols);
    int flush_jobs();
    // generate `nper` symbol periods for every channel into out[c*stride + off + per*(M+cp) + i]
    int generate(cf * out, size_t stride, size_t off, unsigned int nper);
};

int GenBank::init(unsigned int N_, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p, int device_, cudaStream_t st)
{
    device = device_; stream = st; N = N_;
    if (M < 8 || (M & 1) || cp < 1 || cp > M || taper > cp) return b2_fail(B2_ERR_ARG, "invalid OFDM configuration (M=%u cp=%u taper=%u)", M, cp, taper);
    if (ofdm_plan(plan, M, cp, taper, p) != 0) return b2_fail(B2_ERR_ARG, "invalid subcarrier allocation");
    if (M < 16 || M > 4096 || fft_plan(fftM, M) != 0)
        return b2_fail(B2_ERR_UNSUPPORTED, "the CUDA path needs an even number of subcarriers in [16, 4096] whose prime factors are <= 41 (got %u)", M);
    std::vector<float> s0, s1;
    ofdm_training_time(plan, s0, s1);
    std::vector<uint16_t> sc_rank(M, 0xffff);
    for (size_t d = 0; d < plan.data_idx.size(); d++) sc_rank[plan.data_idx[d]] = (uint16_t)d;
    for (size_t n = 0; n < plan.pilot_idx.size(); n++) sc_rank[plan.pilot_idx[n]] = (uint16_t)(0x4000u | n);
    B2_TRY(t_s0.upload(s0)); B2_TRY(t_s1.upload(s1)); B2_TRY(t_taper.upload(ofdm_taper(taper)));
    B2_TRY(t_rank.upload(sc_rank)); B2_TRY(t_seq.upload(plan.pilot_seq));
    B2_TRY(t_perm.upload(fftM.perm)); B2_TRY(t_tw.upload(fftM.tw));
    B2_TRY(d_hmod.alloc((size_t)N * 288)); B2_TRY(d_post.alloc(sizeof(cf) * (size_t)N * (taper + 1)));
    B2_TRY(d_desc.alloc(sizeof(GenDesc) * N));
    ch.assign(N, GenChanHost());
    memset(&fp, 0, sizeof(fp));
    fp.M = M; fp.cp = cp; fp.taper = taper; fp.M_pilot = plan.M_pilot; fp.M_data = plan.M_data;
    fp.g_data = 1.0f / sqrtf((float)(plan.M_pilot + plan.M_data));
    for (int i = 0; i < 9; i++) fp.qam_alpha[i] = 1.0f;
    fp.qam_alpha[2] = 1.0f / sqrtf(2.0f); fp.qam_alpha[4] = 1.0f / sqrtf(10.0f);
    fp.qam_alpha[6] = 1.0f / sqrtf(42.0f); fp.qam_alpha[8] = 1.0f / sqrtf(170.0f);
    fp.nchan = N;
    fp.desc = d_desc.as<GenDesc>(); fp.postfix = d_post.as<cf>();
    fp.header_mod = d_hmod.as<uint8_t>();
    fp.s0 = t_s0.as<cf>(); fp.s1 = t_s1.as<cf>(); fp.taper_w = t_taper.as<float>();
    fp.sc_rank = t_rank.as<uint16_t>(); fp.pilot_seq = t_seq.as<uint8_t>();
    fp.fft.n = M; fp.fft.npass = fftM.npass; fp.fft.radices = 0;
    for (unsigned int i = 0; i < fftM.npass; i++) fp.fft.radices |= fft_radix_code(fftM.radix[i]) << (4 * i);
    fp.fft.perm = t_perm.as<uint16_t>(); fp.fft.tw = t_tw.as<cf>();
    B2_TRY(ensure_capacity(packet_enc_len(1200, CRC_32, FEC_CONV_V27, FEC_HAMMING128), 8 * packet_enc_len(1200, CRC_32, FEC_CONV_V27, FEC_HAMMING128)));
    B2_CUDA(cudaMemsetAsync(d_post.p, 0, d_post.bytes, stream));
    return B2_OK;
}

int GenBank::ensure_capacity(unsigned int enc_len, unsigned int mod_len)
{
    size_t need_work = ((size_t)enc_len + 64 + 15) & ~(size_t)15, need_mod = ((size_t)mod_len + 15) & ~(size_t)15;
    if (need_work <= work_stride && need_mod <= mod_stride) return B2_OK;
    // growing discards the encoded symbols of frames in flight, so finish pending work first
    B2_CUDA(cudaStreamSynchronize(stream));
    size_t new_work = std::max(work_stride, need_work), new_mod = std::max(mod_stride, need_mod);
    DevBuf n_pmod;
    B2_TRY(n_pmod.alloc(new_mod * N));
    if (d_pmod.p && mod_stride)
        B2_CUDA(cudaMemcpy2D(n_pmod.p, new_mod, d_pmod.p, mod_stride, mod_stride, N, cudaMemcpyDeviceToDevice));
    std::swap(d_pmod.p, n_pmod.p); std::swap(d_pmod.bytes, n_pmod.bytes);
    B2_TRY(d_work0.alloc(new_work * N)); B2_TRY(d_work1.alloc(new_work * N));
    work_stride = new_work; mod_stride = new_mod;
    fp.payload_mod = d_pmod.as<uint8_t>(); fp.mod_stride = mod_stride;
    return B2_OK;
}

int GenBank::reset()
{
    // ofdmflexframegen_reset on every generator: nothing assembled, taper postfix cleared
    for (auto & c : ch) c = GenChanHost();
    jobs.clear(); job_payloads.clear();
    B2_CUDA(cudaMemsetAsync(d_post.p, 0, d_post.bytes, stream));
    return B2_OK;
}

int GenBank::assemble(unsigned int c, const unsigned char * header, const unsigned char * payload, unsigned int len,
                      int check, int fec0, int fec1, int mod, unsigned int * n_symbols)
{
    if (c >= N) return b2_fail(B2_ERR_ARG, "invalid channel id %u", c);
    if (!header || (!payload && len)) return b2_fail(B2_ERR_ARG, "null header/payload");
    if (len > 65535) return b2_fail(B2_ERR_ARG, "payload too long (%u)", len);
    int bps = mod_bps((unsigned int)mod);
    if (bps == 0 || !fec_supported((unsigned int)fec0) || !fec_supported((unsigned int)fec1) || (check != CRC_32 && check != CRC_NONE))
        return b2_fail(B2_ERR_UNSUPPORTED, "unsupported frame properties (mod %d, fec %d/%d, check %d)", mod, fec0, fec1, check);
    unsigned int enc = packet_enc_len(len, check, fec0, fec1);
    unsigned int mod_len = (8 * enc + bps - 1) / bps;
    B2_TRY(ensure_capacity(enc, mod_len));
    GenChanHost & g = ch[c];
    g.assembled = true; g.fresh = true; g.symbol = 0;
    g.n_hdr = plan.n_header_syms;
    g.n_pay = (mod_len + plan.M_data - 1) / plan.M_data;
    g.mod = (unsigned int)mod; g.bps = (unsigned int)bps; g.payload_mod_len = mod_len;
    EncodeJob j;
    memset(&j, 0, sizeof(j));
    memcpy(j.header, header, 8);
    j.payload_len = len; j.mod = (unsigned int)mod; j.bps = (unsigned int)bps;
    j.check = (unsigned int)check; j.fec0 = (unsigned int)fec0; j.fec1 = (unsigned int)fec1;
    j.slot = c;
    j.payload_offset = job_payloads.size();
    if (len) job_payloads.insert(job_payloads.end(), payload, payload + len);
    while (job_payloads.size() % 16) job_payloads.push_back(0);
    // a channel re-armed before its previous job was flushed replaces that job
    for (auto & old : jobs) if (old.slot == c) old.slot = 0xffffffffu;
    jobs.push_back(j);
    if (n_symbols) *n_symbols = g.total();
    return B2_OK;
}

int GenBank::flush_jobs()
{
    std::vector<EncodeJob> live;
    for (auto & j : jobs) if (j.slot != 0xffffffffu) live.push_back(j);
    jobs.clear();
    if (live.empty()) { job_payloads.clear(); return B2_OK; }
    if (d_jobs.bytes < live.size() * sizeof(EncodeJob)) B2_TRY(d_jobs.alloc(live.size() * sizeof(EncodeJob) * 2));
    if (d_pay.bytes < job_payloads.size() + 16) B2_TRY(d_pay.alloc(job_payloads.size() * 2 + 64));
    B2_CUDA(cudaMemcpyAsync(d_jobs.p, live.data(), live.size() * sizeof(EncodeJob), cudaMemcpyHostToDevice, stream));
    if (!job_payloads.empty())
        B2_CUDA(cudaMemcpyAsync(d_pay.p, job_payloads.data(), job_payloads.size(), cudaMemcpyHostToDevice, stream));
    EncodeParams ep;
    ep.jobs = d_jobs.as<EncodeJob>(); ep.nframes = (unsigned int)live.size();
    ep.payloads = d_pay.as<uint8_t>();
    ep.header_mod = d_hmod.as<uint8_t>(); ep.payload_mod = d_pmod.as<uint8_t>(); ep.mod_stride = mod_stride;
    ep.work0 = d_work0.as<uint8_t>(); ep.work1 = d_work1.as<uint8_t>(); ep.work_stride = work_stride;
    B2_CUDA(packet_encode_launch(ep, stream));
    // the host vectors are reused by the caller right away: wait for the copies
    B2_CUDA(cudaStreamSynchronize(stream));
    job_payloads.clear();
    return B2_OK;
}

int GenBank::generate(cf * out, size_t stride, size_t off, unsigned int nper)
{
    B2_TRY(flush_jobs());
    std::vector<GenDesc> desc(N);
    for (unsigned int c = 0; c < N; c++) {
        GenChanHost & g = ch[c];
        GenDesc & d = desc[c];
        memset(&d, 0, sizeof(d));
        if (g.assembled) {
            unsigned int left = g.total() - g.symbol;
            d.first_symbol = g.symbol;
            d.n_periods = std::min(left, nper);
            d.n_hdr = g.n_hdr; d.n_pay = g.n_pay; d.mod = g.mod; d.bps = g.bps; d.payload_mod_len = g.payload_mod_len;
            d.fresh = g.fresh ? 1u : 0u;
            g.symbol += d.n_periods;
            g.fresh = false;
            if (g.symbol >= g.total()) g.assembled = false;      // the tail buffer has been written
        } else {
            d.bps = 1;
        }
    }
    B2_CUDA(cudaMemcpyAsync(d_desc.p, desc.data(), sizeof(GenDesc) * N, cudaMemcpyHostToDevice, stream));
    FramegenParams q = fp;
    q.nper = nper;
    q.out = out; q.out_stride = stride; q.out_off = off;
    B2_CUDA(framegen_launch(q, stream));
    B2_CUDA(cudaStreamSynchronize(stream));          // desc is a stack vector
    return B2_OK;
}

// ================================================================== multichanneltx
struct b2_mctx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int N = 0, K = 0, lgK = 0, P = 26, TB = 8, W = 0;
    FftPlan fftK;
    DevBuf t_taps, t_perm, t_tw;
    DevBuf d_sym, d_tmp, d_vhist[2], d_out;
    size_t sym_cap = 0, max_calls = 0;
    size_t sym_avail = 0;                // generated channel samples not yet consumed
    int vh = 0;
    uint32_t nco_theta = 0, nco_dtheta = 0;
    size_t syn_smem = 0;
    int syn_grid = 148;
    GenBank bank;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_ms[4] = {0, 0, 0, 0};
};

extern "C" int b2_mctx_create(unsigned int N, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p,
                              int device, b2_mctx ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    // same argument checks as multichanneltx::multichanneltx (lib/multichanneltx.cc:48-60)
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
    b2_mctx * q = new b2_mctx_s;
    q->device = device; q->N = N; q->K = K; q->lgK = ceil_log2(K); q->W = M + cp;
    q->TB = (K <= 512) ? (std::min(256u, std::max(8u, 4096u / K)) & ~7u) : 2u;   // a multiple of the kernel's JB (8 or 2)
    int rc = B2_OK;
    do {
        if (cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        // firpfbch_crcf_create_kaiser(LIQUID_SYNTHESIZER, 2N, m=13, As=60): lib/multichanneltx.cc:85-87
        std::vector<float> h = firpfbch_prototype(K, 13, 60.0f);
        fft_plan(q->fftK, K);
        if ((rc = q->t_taps.upload(h)) || (rc = q->t_perm.upload(q->fftK.perm)) || (rc = q->t_tw.upload(q->fftK.tw))) break;
        q->max_calls = std::max<size_t>(4 * (size_t)q->W, ((size_t)1 << 22) / K);
        q->sym_cap = (q->max_calls + 2 * (size_t)q->W + 1) & ~(size_t)1;
        if ((rc = q->d_sym.alloc(sizeof(cf) * q->sym_cap * N)) || (rc = q->d_tmp.alloc(sizeof(cf) * (size_t)q->W * N))) break;
        if ((rc = q->d_vhist[0].alloc(sizeof(cf) * (size_t)(q->P - 1) * K)) || (rc = q->d_vhist[1].alloc(sizeof(cf) * (size_t)(q->P - 1) * K))) break;
        if ((rc = q->d_out.alloc(sizeof(cf) * q->max_calls * K))) break;
        float offset = -0.5f * (float)(N - 1) / (float)N * M_PI;       // lib/multichanneltx.cc:94-96
        q->nco_dtheta = nco_constrain(offset);
        SynthParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.K = K; sp.TB = q->TB;
        q->syn_smem = synth_smem_bytes(sp);
        if (q->syn_smem > 227 * 1024) { rc = b2_fail(B2_ERR_UNSUPPORTED, "synthesizer tile does not fit shared memory"); break; }
        if (synth_configure(q->syn_smem) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaFuncSetAttribute failed"); break; }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        q->syn_grid = sms;
        if ((rc = q->bank.init(N, M, cp, taper, p, device, q->stream))) break;
        for (int i = 0; i < 4; i++) cudaEventCreate(&q->ev[i]);
        cudaMemsetAsync(q->d_vhist[0].p, 0, q->d_vhist[0].bytes, q->stream);
        cudaMemsetAsync(q->d_vhist[1].p, 0, q->d_vhist[1].bytes, q->stream);
        cudaMemsetAsync(q->d_sym.p, 0, q->d_sym.bytes, q->stream);
        if (cudaStreamSynchronize(q->stream) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "initialisation failed"); break; }
    } while (0);
    if (rc) { b2_mctx_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}

extern "C" int b2_mctx_destroy(b2_mctx * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    if (q->stream) cudaStreamSynchronize(q->stream);
    for (int i = 0; i < 4; i++) if (q->ev[i]) cudaEventDestroy(q->ev[i]);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
    return B2_OK;
}

// multichanneltx::Reset (lib/multichanneltx.cc:126-149): generators + filterbank windows; the
// symbol buffers are exhausted (fgbuffer_index = fgbuffer_len); the NCO keeps running (:135)
extern "C" int b2_mctx_reset(b2_mctx * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    B2_TRY(q->bank.reset());
    B2_CUDA(cudaMemsetAsync(q->d_vhist[0].p, 0, q->d_vhist[0].bytes, q->stream));
    B2_CUDA(cudaMemsetAsync(q->d_vhist[1].p, 0, q->d_vhist[1].bytes, q->stream));
    q->sym_avail = 0;
    B2_CUDA(cudaStreamSynchronize(q->stream));
    return B2_OK;
}

extern "C" int b2_mctx_nco_advance(b2_mctx * q, int64_t n_samples)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    q->nco_theta += (uint32_t)((uint64_t)n_samples) * q->nco_dtheta;       // uint32 phase: exact, wraps
    return B2_OK;
}

extern "C" int b2_mctx_is_ready(b2_mctx * q, unsigned int channel, int * ready)
{
    if (!q || !ready) return b2_fail(B2_ERR_ARG, "null argument");
    if (channel >= q->N) return b2_fail(B2_ERR_ARG, "invalid channel id %u", channel);
    *ready = q->bank.ch[channel].assembled ? 0 : 1;
    return B2_OK;
}

extern "C" int b2_mctx_update(b2_mctx * q, unsigned int channel, const unsigned char * header, const unsigned char * payload,
                              unsigned int payload_len, int mod, int fec0, int fec1)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (channel >= q->N) return b2_fail(B2_ERR_ARG, "invalid channel id %u", channel);
    if (q->bank.ch[channel].assembled) return b2_fail(B2_ERR_STATE, "channel %u not ready yet", channel);
    B2_CUDA(cudaSetDevice(q->device));
    return q->bank.assemble(channel, header, payload, payload_len, CRC_32, fec0, fec1, mod, nullptr);   // CRC-32 always (:184)
}

extern "C" int b2_mctx_update_many(b2_mctx * q, unsigned int n, const unsigned int * channels, const unsigned char * headers,
                                   const unsigned char * payloads, const unsigned int * payload_lens, int mod, int fec0, int fec1,
                                   unsigned int * n_updated)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n_updated) *n_updated = 0;
    if (n == 0) return B2_OK;
    if (!channels || !headers || !payload_lens) return b2_fail(B2_ERR_ARG, "null argument");
    B2_CUDA(cudaSetDevice(q->device));
    size_t off = 0;
    unsigned int taken = 0;
    for (unsigned int i = 0; i < n; i++) {
        const unsigned int c = channels[i], len = payload_lens[i];
        if (c >= q->N) return b2_fail(B2_ERR_ARG, "invalid channel id %u", c);
        if (!q->bank.ch[c].assembled) {
            B2_TRY(q->bank.assemble(c, headers + 8 * (size_t)i, payloads ? payloads + off : nullptr, len, CRC_32, fec0, fec1, mod, nullptr));
            taken++;
        }
        off += len;
    }
    if (n_updated) *n_updated = taken;
    return B2_OK;
}

extern "C" int b2_mctx_calls_to_boundary(b2_mctx * q, size_t * n_calls)
{
    if (!q || !n_calls) return b2_fail(B2_ERR_ARG, "null argument");
    *n_calls = q->sym_avail ? q->sym_avail : q->W;
    return B2_OK;
}

static int mctx_generate_chunk(b2_mctx * q, cf * out_dev, size_t n_calls)
{
    const unsigned int K = q->K, W = q->W;
    cudaEventRecord(q->ev[0], q->stream);
    if (n_calls > q->sym_avail) {
        unsigned int nper = (unsigned int)((n_calls - q->sym_avail + W - 1) / W);
        B2_TRY(q->bank.generate(q->d_sym.as<cf>(), q->sym_cap, q->sym_avail, nper));
        q->sym_avail += (size_t)nper * W;
    }
    cudaEventRecord(q->ev[1], q->stream);
    SynthParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.in = q->d_sym.as<cf>(); sp.in_stride = q->sym_cap; sp.in_off = 0;
    sp.K = K; sp.lgK = q->lgK; sp.N = q->N; sp.P = q->P; sp.TB = q->TB;
    sp.nblocks = (unsigned int)n_calls;
    sp.taps = q->t_taps.as<float>();
    sp.vhist = q->d_vhist[q->vh].as<cf>(); sp.vhist_out = q->d_vhist[q->vh ^ 1].as<cf>();
    sp.theta0 = q->nco_theta; sp.dtheta = q->nco_dtheta;
    sp.out = out_dev;
    sp.fft.n = K; sp.fft.npass = q->fftK.npass; sp.fft.radices = 0;
    for (unsigned int i = 0; i < q->fftK.npass; i++) sp.fft.radices |= fft_radix_code(q->fftK.radix[i]) << (4 * i);
    sp.fft.perm = q->t_perm.as<uint16_t>(); sp.fft.tw = q->t_tw.as<cf>();
    B2_CUDA(synth_launch(sp, q->syn_grid, q->syn_smem, q->stream));
    cudaEventRecord(q->ev[2], q->stream);
    q->vh ^= 1;
    q->nco_theta += (uint32_t)(n_calls * K) * q->nco_dtheta;
    // unconsumed symbol samples move to the front
    size_t left = q->sym_avail - n_calls;
    if (left) {
        B2_CUDA(cudaMemcpy2DAsync(q->d_tmp.p, sizeof(cf) * W, q->d_sym.as<cf>() + n_calls, sizeof(cf) * q->sym_cap,
                                  sizeof(cf) * left, q->N, cudaMemcpyDeviceToDevice, q->stream));
        B2_CUDA(cudaMemcpy2DAsync(q->d_sym.p, sizeof(cf) * q->sym_cap, q->d_tmp.p, sizeof(cf) * W,
                                  sizeof(cf) * left, q->N, cudaMemcpyDeviceToDevice, q->stream));
    }
    q->sym_avail = left;
    cudaEventRecord(q->ev[3], q->stream);
    return B2_OK;
}

static int mctx_generate_any(b2_mctx * q, float * out, size_t n_calls, bool on_device)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n_calls == 0) return B2_OK;
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    B2_CUDA(cudaSetDevice(q->device));
    size_t done = 0;
    while (done < n_calls) {
        size_t c = std::min(n_calls - done, q->max_calls);
        cf * dst = on_device ? (cf *)out + done * q->K : q->d_out.as<cf>();
        B2_TRY(mctx_generate_chunk(q, dst, c));
        if (!on_device)
            B2_CUDA(cudaMemcpyAsync(out + 2 * done * q->K, dst, sizeof(cf) * c * q->K, cudaMemcpyDeviceToHost, q->stream));
        B2_CUDA(cudaStreamSynchronize(q->stream));
        cudaEventElapsedTime(&q->last_ms[0], q->ev[0], q->ev[1]);
        cudaEventElapsedTime(&q->last_ms[1], q->ev[1], q->ev[2]);
        cudaEventElapsedTime(&q->last_ms[2], q->ev[2], q->ev[3]);
        cudaEventElapsedTime(&q->last_ms[3], q->ev[0], q->ev[3]);
        done += c;
    }
    return B2_OK;
}
extern "C" int b2_mctx_generate(b2_mctx * q, float * out_host, size_t n_calls) { return mctx_generate_any(q, out_host, n_calls, false); }
extern "C" int b2_mctx_generate_device(b2_mctx * q, float * out_dev, size_t n_calls) { return mctx_generate_any(q, out_dev, n_calls, true); }
extern "C" int b2_mctx_last_timing(b2_mctx * q, float ms[4])
{
    if (!q || !ms) return b2_fail(B2_ERR_ARG, "null argument");
    for (int i = 0; i < 4; i++) ms[i] = q->last_ms[i];
    return B2_OK;
}

// ================================================================== single-link frame generator
struct b2_ofdmgen_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int W = 0;
    DevBuf d_out;
    size_t out_cap = 0;                  // symbols
    GenBank bank;
};

extern "C" int b2_ofdmgen_create(unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p, int device, b2_ofdmgen ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return b2_fail(B2_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return b2_fail(B2_ERR_ARG, "invalid device ordinal %d", device);
    B2_CUDA(cudaSetDevice(device));
    b2_ofdmgen * q = new b2_ofdmgen_s;
    q->device = device; q->W = M + cp;
    int rc = B2_OK;
    do {
        if (cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        if ((rc = q->bank.init(1, M, cp, taper, p, device, q->stream))) break;
        q->out_cap = 64;
        if ((rc = q->d_out.alloc(sizeof(cf) * q->out_cap * q->W))) break;
    } while (0);
    if (rc) { b2_ofdmgen_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}
extern "C" int b2_ofdmgen_destroy(b2_ofdmgen * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    if (q->stream) { cudaStreamSynchronize(q->stream); cudaStreamDestroy(q->stream); }
    delete q;
    return B2_OK;
}
extern "C" int b2_ofdmgen_reset(b2_ofdmgen * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    return q->bank.reset();
}
extern "C" int b2_ofdmgen_is_assembled(b2_ofdmgen * q, int * assembled)
{
    if (!q || !assembled) return b2_fail(B2_ERR_ARG, "null argument");
    *assembled = q->bank.ch[0].assembled ? 1 : 0;
    return B2_OK;
}
extern "C" int b2_ofdmgen_assemble(b2_ofdmgen * q, const unsigned char * header, const unsigned char * payload, unsigned int payload_len,
                                   int check, int fec0, int fec1, int mod, unsigned int * n_symbols)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    // ofdmflexframegen_assemble resets the generator first (a frame in flight is dropped)
    B2_TRY(q->bank.reset());
    return q->bank.assemble(0, header, payload, payload_len, check, fec0, fec1, mod, n_symbols);
}
extern "C" int b2_ofdmgen_write(b2_ofdmgen * q, float * out_host, unsigned int n_symbols, int * last)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    if (n_symbols == 0) return B2_OK;
    if (!out_host) return b2_fail(B2_ERR_ARG, "null output pointer");
    B2_CUDA(cudaSetDevice(q->device));
    if (n_symbols > q->out_cap) {
        q->out_cap = n_symbols;
        B2_TRY(q->d_out.alloc(sizeof(cf) * q->out_cap * q->W));
    }
    B2_TRY(q->bank.generate(q->d_out.as<cf>(), (size_t)q->out_cap * q->W, 0, n_symbols));
    B2_CUDA(cudaMemcpyAsync(out_host, q->d_out.p, sizeof(cf) * (size_t)n_symbols * q->W, cudaMemcpyDeviceToHost, q->stream));
    B2_CUDA(cudaStreamSynchronize(q->stream));
    if (last) *last = q->bank.ch[0].assembled ? 0 : 1;
    return B2_OK;
}

// ================================================================== msresamp_crcf
struct b2_msresamp_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    float rate = 1.0f;
    unsigned int m = 7, npfb_bits = 6;
    DevBuf t_h, d_x, d_y, d_hist;       // d_x / d_y stage host buffers; d_hist = the 2m-1 samples before the next call
    size_t x_cap = 0, y_cap = 0;
    unsigned long long tau = 0, step = 0;
};

extern "C" int b2_msresamp_create(float rate, float As, int device, b2_msresamp ** out)
{
    if (!out) return b2_fail(B2_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (!(rate > 0.0f)) return b2_fail(B2_ERR_ARG, "resampling rate must be positive");
    if (!(rate >= 0.5f && rate <= 2.0f)) return b2_fail(B2_ERR_UNSUPPORTED, "rates outside [0.5, 2] need half-band stages (got %g)", rate);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return b2_fail(B2_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return b2_fail(B2_ERR_ARG, "invalid device ordinal %d", device);
    B2_CUDA(cudaSetDevice(device));
    b2_msresamp * q = new b2_msresamp_s;
    q->device = device; q->rate = rate;
    int rc = B2_OK;
    do {
        if (cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = b2_fail(B2_ERR_CUDA, "cudaStreamCreate failed"); break; }
        if ((rc = q->t_h.upload(resamp_prototype(rate, As, q->m, 1u << q->npfb_bits))) != B2_OK) break;
        q->step = (unsigned long long)llrint(4294967296.0 / (double)rate);
        q->x_cap = ((size_t)1 << 20);
        q->y_cap = (size_t)(q->x_cap * 2.1) + 64;
        if ((rc = q->d_x.alloc(sizeof(cf) * q->x_cap)) || (rc = q->d_y.alloc(sizeof(cf) * q->y_cap)) || (rc = q->d_hist.alloc(sizeof(cf) * 32))) break;
        cudaMemsetAsync(q->d_hist.p, 0, q->d_hist.bytes, q->stream);
        cudaStreamSynchronize(q->stream);
    } while (0);
    if (rc) { b2_msresamp_destroy(q); return rc; }
    *out = q;
    return B2_OK;
}
extern "C" int b2_msresamp_destroy(b2_msresamp * q)
{
    if (!q) return B2_OK;
    cudaSetDevice(q->device);
    if (q->stream) { cudaStreamSynchronize(q->stream); cudaStreamDestroy(q->stream); }
    delete q;
    return B2_OK;
}
extern "C" int b2_msresamp_reset(b2_msresamp * q)
{
    if (!q) return b2_fail(B2_ERR_ARG, "null handle");
    B2_CUDA(cudaSetDevice(q->device));
    B2_CUDA(cudaMemsetAsync(q->d_hist.p, 0, q->d_hist.bytes, q->stream));
    q->tau = 0;
    B2_CUDA(cudaStreamSynchronize(q->stream));
    return B2_OK;
}
static int msresamp_any(b2_msresamp * q, const float * x, size_t nx, float * y, size_t y_cap, size_t * ny_out, bool x_dev, bool y_dev)
{
    if (!q || !ny_out) return b2_fail(B2_ERR_ARG, "null argument");
    *ny_out = 0;
    if (nx == 0) return B2_OK;
    if (!x || !y) return b2_fail(B2_ERR_ARG, "null sample pointer");
    if (nx >> 31) return b2_fail(B2_ERR_ARG, "at most 2^31 samples per call");
    B2_CUDA(cudaSetDevice(q->device));
    ResampParams rp;
    rp.hist_buf = q->d_hist.as<cf>(); rp.hist = 2 * q->m - 1;
    rp.h = q->t_h.as<float>(); rp.npfb_bits = q->npfb_bits; rp.m2 = 2 * q->m;
    rp.step = q->step;
    // device input: one launch over the caller's memory; host buffers are staged by chunks of x_cap samples
    size_t done = 0, produced = 0;
    while (done < nx) {
        size_t c = x_dev ? nx : std::min(nx - done, q->x_cap);
        // outputs k with tau + k*step < c * 2^32
        unsigned long long span = (unsigned long long)c << 32;
        unsigned long long ny = (q->tau < span) ? (span - q->tau + q->step - 1) / q->step : 0;
        if (produced + ny > y_cap) return b2_fail(B2_ERR_OVERFLOW, "output buffer too small (%zu needed)", (size_t)(produced + ny));
        if (x_dev) rp.x = (const cf *)x;
        else {
            B2_CUDA(cudaMemcpyAsync(q->d_x.p, x + 2 * done, sizeof(cf) * c, cudaMemcpyHostToDevice, q->stream));
            rp.x = q->d_x.as<cf>();
        }
        rp.y = y_dev ? (cf *)y + produced : q->d_y.as<cf>();
        rp.nx = c; rp.tau0 = q->tau; rp.ny = ny;
        B2_CUDA(resamp_launch(rp, q->stream));
        if (!y_dev && ny)
            B2_CUDA(cudaMemcpyAsync(y + 2 * produced, q->d_y.p, sizeof(cf) * ny, cudaMemcpyDeviceToHost, q->stream));
        q->tau = q->tau + ny * q->step - span;
        produced += ny;
        done += c;
    }
    B2_CUDA(cudaStreamSynchronize(q->stream));
    *ny_out = produced;
    return B2_OK;
}
extern "C" int b2_msresamp_execute(b2_msresamp * q, const float * x_host, size_t nx, float * y_host, size_t y_cap, size_t * ny)
{ return msresamp_any(q, x_host, nx, y_host, y_cap, ny, false, false); }
extern "C" int b2_msresamp_execute_device(b2_msresamp * q, const float * x_dev, size_t nx, float * y_dev, size_t y_cap, size_t * ny)
{ return msresamp_any(q, x_dev, nx, y_dev, y_cap, ny, true, true); }
extern "C" int b2_msresamp_execute_to_device(b2_msresamp * q, const float * x_host, size_t nx, float * y_dev, size_t y_cap, size_t * ny)
{ return msresamp_any(q, x_host, nx, y_dev, y_cap, ny, false, true); }
