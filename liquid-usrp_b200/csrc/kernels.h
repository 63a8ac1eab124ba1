// kernels.h -- launch interfaces between the C ABI (capi.cu) and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dsp.cuh"

namespace b2 {

// ------------------------------------------------------------------ analysis channelizer
// NCO mix-down + firpfbch_crcf analyzer (lib/multichannelrx.cc:163-164,188), keeping output
// channels 0..N-1 (lib/multichannelrx.cc:193-194), written channel-major.
struct AnalyzerParams {
    const cf * seg0;            // logical input rows [0, rows0): seg0 + row*K
    unsigned int rows0;
    const cf * seg1;            // logical input rows [rows0, ...): seg1 + (row-rows0)*K
    unsigned int K, lgK, N, P;  // filterbank size (= 1<<lgK), channels kept, taps per branch
    unsigned int TB;            // blocks per tile (multiple of JB)
    unsigned int sm_limit;      // SMs available to this launch (0 = the whole device)
    unsigned int block0;        // first output block of this launch (chunk of a longer call)
    unsigned int nblocks;       // output blocks this launch; block b reads rows b .. b+P-1
    const float * taps;         // [P][K]: taps[n*K + i] = h[i + n*K]
    uint32_t theta0, dtheta;    // NCO phase of logical sample 0, phase step per sample
    int row_alt;                // set by the launcher: K*dtheta == 0 (+1) or pi (-1) mod 2 pi, else 0
    cf * out;                   // out[c*out_stride + out_col0 + b]
    size_t out_stride, out_col0;
    // multi-GPU split (capi_shard.cu): channel c belongs to GPU c / chan_per_peer and is written straight into that
    // GPU's memory, out_peer[c / chan_per_peer][(c % chan_per_peer)*out_stride + out_col0 + b] (peer mapping over
    // NVLink; n_peer == 0: `out`)
    cf * out_peer[8];
    unsigned int n_peer, chan_per_peer;
    FftDev fft;                 // K-point plan (perm / tw in global memory)
};
size_t analyzer_smem_bytes(const AnalyzerParams & p);
cudaError_t analyzer_configure(size_t smem_bytes);
cudaError_t analyzer_launch(const AnalyzerParams & p, int grid, size_t smem_bytes, cudaStream_t st);
// column-per-thread fast path (channelizer8.cu), K in {64,128,256,512}; analyzer_launch() picks it
bool analyzer8_supported(const AnalyzerParams & p);
cudaError_t analyzer8_launch(const AnalyzerParams & p, cudaStream_t st);

// ------------------------------------------------------------------ per-stream OFDM synchroniser
// One CTA walks one stream (channel) through the ofdmflexframesync state machine
// (lib/multichannelrx.cc:194 -> liquid ofdmframesync/ofdmflexframesync).
enum { ST_SEEK = 0, ST_S0A, ST_S0B, ST_S1, ST_RX };
enum { FS_HEADER = 0, FS_PAYLOAD };

struct SyncState {              // persistent per stream, lives in global memory
    int32_t  state, timer;
    uint32_t num_symbols;
    uint32_t nco_theta, nco_dtheta;
    float    g0;
    float    s_hat0_re, s_hat0_im;
    float    phi_prime, p1_prime;
    uint32_t pilot_pos;         // position in the 255-periodic pilot sequence
    uint32_t ring_head;         // index of the oldest sample in the window ring
    int32_t  fstate;
    uint32_t header_sym_idx, payload_sym_idx;
    float    evm_hat, evm_db;
    uint32_t ms_payload, bps_payload, payload_len, check, fec0, fec1;
    uint32_t payload_enc_len, payload_mod_len;
    uint64_t sample_index;      // index of the next sample to be pushed
    uint64_t detect_index;
    uint8_t  header_bits[36];
    uint8_t  header_dec[20];
    // frame-pipelined workers (ofdmsync8.cu, SyncParams::workers == 2; unused otherwise)
    uint32_t role;              // SW_OWNER / SW_SPEC / SW_WAIT
    uint32_t my_seq;            // hand-off this worker's frame began with (role SW_SPEC)
    uint32_t sent_seq;          // hand-off this worker published for the frame it is in (0: none)
    uint32_t verify;            // frame over, hand-off published: the next seek event settles it
    uint32_t pub_pending, pad0; // header decoded while speculative: publish pub_start once confirmed
    uint64_t pub_start;
};
enum { SW_OWNER = 0, SW_SPEC = 1, SW_WAIT = 2 };
// hand-off word of a chain: (sequence number << 2) | state
enum { HS_NONE = 0, HS_SENT = 1, HS_ACCEPTED = 2, HS_ABORTED = 3 };
struct SyncCtl {                // per chain, global memory
    unsigned long long start;   // absolute index of the first sample of the frame search handed off
    unsigned int hs;            // hand-off word
    unsigned int done[2];       // launch id in which worker 0 / 1 last left the kernel
    unsigned int pad;
};

// ---- frame-parallel synchroniser (ofdmsyncw.cu): one WARP per worker, `workers` workers per stream and launch.
// Worker 0 carries the stream's true state on from the previous launch; workers 1.. start at PREDICTED positions
// of the stream (see WChan) in the canonical state ofdmflexframesync is in right after a seek event that detected
// nothing (state SEEK, timer 0, everything else reset).  A worker runs up to the next worker's start; if it gets
// there in exactly that canonical state the next worker's results are the serial chain's, else the last worker of
// the stream to finish (the "stitcher") redoes the rest serially from the true state.  Nobody waits for anybody.
struct WSync {                  // per worker slot, global memory; the stream's carried state lives in slot WChan::head
    int32_t  state, timer;
    uint32_t num_symbols;
    uint32_t nco_theta, nco_dtheta;      // phase of the next sample / step
    float    g0;
    float    s_hat0_re, s_hat0_im;
    float    phi_prime, p1_prime;
    uint32_t pilot_pos;
    int32_t  fstate;
    uint32_t header_sym_idx, payload_sym_idx;
    float    evm_hat, evm_db;
    uint32_t ms_payload, bps_payload, payload_len, check, fec0, fec1;
    uint32_t payload_enc_len, payload_mod_len;
    // The kernel keeps no sample window: an event's FFT window is re-read from the raw stream and re-mixed.  Samples
    // [mix_start, mix_end) were pushed through the NCO; mix_end == ~0: the segment is open and follows the running
    // NCO (phase of sample i = nco_theta + (i - sample_index) * nco_dtheta), else it is closed (frame over, NCO
    // reset) and sample i had phase q_theta - (mix_end - i) * q_dtheta.  Everything else was pushed unmixed.
    uint32_t q_theta, q_dtheta;
    // payload symbols of the frame in progress go straight into the arena, a RING that outlives launches and
    // batches: sym_abs = value of the allocation counter the frame got (~0: it did not fit, the frame will be
    // reported without payload), sym_off = the same position inside the ring
    uint64_t sym_abs, sym_off;
    uint64_t sample_index;               // index of the next sample to be pushed
    uint64_t detect_index;
    uint64_t mix_start, mix_end;
    // summary of the last launch (read by the stitcher)
    uint64_t b_last, b_prev;             // sample_index right after the last two frame ends (+1: after the seek event
                                         // liquid runs on the first sample behind a frame)
    uint32_t nb;                         // frame ends seen in the launch (saturates at 2)
    uint32_t matched;                    // reached its limit in the canonical state
    uint32_t nrec, pad;                  // records in the worker's private list
    uint8_t  header_bits[36];
    uint8_t  header_dec[20];
};
struct WChan {                  // per stream, global memory.  [launch_id & 1] is read by a launch, the other entry
                                // written by its stitcher (a warp that starts late must still see what its launch began with)
    uint64_t pred_next[2];      // predicted canonical positions: pred_next + j * pred_period
    uint32_t pred_period[2];    // 0: no prediction (one worker)
    uint32_t head[2];           // slot holding the stream's carried state
    uint64_t b_last, b_prev;    // last two frame-end positions of the stream (stitcher only)
    uint32_t done, pad;         // workers of the running launch that have finished
};

struct FrameRec {               // same layout as b2_frame_rec (include/b200_ofdm.h)
    uint32_t channel;
    int32_t  header_valid, payload_valid;
    uint32_t payload_len;
    uint8_t  header[8];
    float    evm, rssi, cfo;
    uint32_t mod_scheme, mod_bps, check, fec0, fec1;
    uint64_t detect_index, complete_index;
    uint64_t payload_offset;    // device: byte offset of the encoded payload in the arena
};

struct FrameAux {               // device-only companion of FrameRec
    uint32_t enc_len;           // encoded payload bytes
    uint32_t sym_bps;           // 0: the arena holds the enc_len packed bytes; else it holds the
                                // ceil(8*enc_len/sym_bps) demapped symbols, one per byte
    uint64_t sym_off;           // where in the arena (FrameRec::payload_offset is where the DECODED payload goes)
};

struct SyncTables {             // read-only, global memory
    const uint8_t * sctype;     // [M]
    const float * S0, * S1;     // [M] +-1/0
    const uint16_t * data_idx;  // [M_data]
    const uint16_t * pilot_idx; // [M_pilot] fft-shifted visiting order
    const float * pilot_x;      // [M_pilot]
    const uint16_t * active_idx;// [M_pilot+M_data] fft-shifted visiting order
    const uint8_t * pilot_seq;  // [255]
    const uint16_t * hdr_walk;  // [4][18] header de-interleaver walks (n = 36)
    const cf * B;               // [M] e^{j 2 pi backoff i / M}
    const double * eqfit_P;     // [M_pilot+M_data][5] least-squares matrix of the S1 gain fit (design.h)
    const uint16_t * act_rank;  // [M] position in the fft-shifted visiting order of the active subcarriers; null: 0xffff
    const uint16_t * sc_rank;   // [M] data: rank among data subcarriers (ascending index);
                                //     pilot: 0x4000 | rank in fft-shifted visiting order; null: 0xffff
};

struct SyncParams {
    unsigned int M, cp, M2, backoff;
    unsigned int M_pilot, M_data, M_S0, M_S1;
    float thresh, pilot_sx, pilot_sxx;
    float b_cos, b_sin;         // e^{j 2 pi backoff / M}: rotation of the S1 metric (ofdmframesync_execute_S1)
    float qam_alpha[9];         // 1/sqrt(2,10,42,170) at index bps = 2,4,6,8
    unsigned int streams;
    unsigned int chan_base;     // added to the stream index in FrameRec::channel (channel-sharded receivers)
    // workers == 2 (ofdmsync8.cu only): two CTAs per stream take alternate frames; st / ring / G0 / R / penc
    // then hold 2*streams entries (index 2*stream + worker)
    unsigned int workers, launch_id;
    unsigned long long sample_base;   // absolute index of in[..][0] (samples given to earlier launches)
    SyncCtl * ctl;              // [streams]
    const cf * in;              // in[s*in_stride + t], t < nsamples
    size_t in_stride;
    unsigned int nsamples;
    SyncState * st;             // [streams]
    cf * ring;                  // [streams][M+cp]
    cf * G0;                    // [streams][M]
    cf * R;                     // [streams][M]
    uint8_t * penc;             // [streams][penc_cap] in-progress encoded payload
    size_t penc_cap;
    // outputs
    FrameRec * recs; FrameAux * aux; unsigned int recs_cap;
    uint8_t * arena; unsigned long long arena_cap;
    unsigned long long decoded_cap;      // bytes of decoded payload a batch may produce (ofdmsyncw.cu)
    unsigned int * counters;    // [0] n_recs, [1] overflow flag, [2..3] arena bytes (u64), [4] n_tap,
                                // [6..7] decoded-payload bytes (u64; ofdmsyncw.cu, where [2..3] is the ring's
                                // allocation counter and is never reset)
    // debug tap
    cf * tap_X; uint32_t * tap_chan; unsigned long long * tap_index; unsigned int tap_cap;
    SyncTables tb;
    FftDev fft;                 // M-point plan
    // frame-parallel kernel (ofdmsyncw.cu); workers = workers of this launch (<= wslots), ring = the M + cp raw
    // samples before in[..][0], penc / wRG indexed by stream * wslots + slot
    WSync * wst;                // [streams][wslots]
    WChan * wch;                // [streams]
    cf * wRG;                   // [streams][wslots][M] equaliser taps (state RX) / S0a gains (state S0B)
    FrameRec * wrecs; FrameAux * waux;   // private record lists of the speculative workers, [streams][wrec_stride]
    unsigned int wrec_stride, wslots;
};
size_t sync_smem_bytes(const SyncParams & p);
cudaError_t sync_configure(size_t smem_bytes);
cudaError_t sync_launch(const SyncParams & p, int threads, size_t smem_bytes, cudaStream_t st);
void sync_state_init(SyncState & s, unsigned int M, unsigned int cp);
cudaError_t sync_reset_launch(SyncState * st, unsigned int streams, unsigned int workers, SyncCtl * ctl,
                              unsigned long long sample_base, cudaStream_t stream);
// register-resident fast path (ofdmsync8.cu), M/8 threads per stream; sync_launch() picks it
bool sync8_supported(unsigned int M);
size_t sync8_smem_bytes(const SyncParams & p);
cudaError_t sync8_launch(const SyncParams & p, cudaStream_t st);
// CTAs of the kernel that fit one SM (both workers of every stream must be resident at once)
int sync8_ctas_per_sm(const SyncParams & p);
// frame-parallel warp-per-worker kernel (ofdmsyncw.cu), M in {256, 512}
bool syncw_supported(unsigned int M);
cudaError_t syncw_launch(const SyncParams & p, cudaStream_t st);
// ofdmflexframesync_reset on every stream (state of slot `head`; prediction restarts from the idle seek grid)
cudaError_t syncw_reset_launch(WSync * wst, WChan * wch, unsigned int streams, unsigned int wslots, unsigned int M,
                               unsigned long long sample_base, cudaStream_t st);

// ------------------------------------------------------------------ packet decode
// de-interleave + FEC decode + CRC of every completed frame (liquid packetizer_decode, called
// from inside ofdmflexframesync_execute in the reference)
struct RangeMark {               // output counters after a chunk's synchroniser
    unsigned int nrec, pad;      // pad carries the overflow flag (counters[1])
    unsigned long long arena_used;
};
struct PacketParams {
    FrameRec * recs; const FrameAux * aux;
    const RangeMark * range;         // device: records [range[0].nrec, range[1].nrec) belong to this launch
    uint8_t * arena;                 // encoded payloads (decoded in place / via scratch)
    uint8_t * scratch;               // same size as arena
    uint8_t * decoded;               // same size as arena; payload bytes end up at payload_offset
    uint2 * vit_local;               // Viterbi decisions: vit_local_ctas regions of vit_local_steps trellis steps,
    unsigned int vit_local_steps;    // one per CTA of the general decode kernel (null: device-wide slots only)
    unsigned int vit_local_ctas;
    unsigned int vit_split;          // launches with at most this many frames take the 128-thread shape of the general kernel
    unsigned int vit_grid128;        // grid of that shape (the one-warp shape gets vit_local_ctas CTAs)
    unsigned int vit_parallel;       // conv-coded frames with a CRC: 0 exact decoder only; 1 speculative decode first (segmented recursion
                                     // where the launch is latency-bound, else thread-parallel traceback); 2 thread-parallel traceback only
    unsigned int * crc_cache;        // [8] device words: [0] = 1 + per of the cached constants (0: empty), [1..5] = x^(8 per 2^l) mod P
};
cudaError_t packet_decode_launch(const PacketParams & p, int grid, cudaStream_t st);
// device-wide Viterbi workspace of the calling thread's device (once; called when a handle is created)
cudaError_t packet_decode_prepare();
// mark_out = {counters[0], counters[2..3]} (records / arena bytes so far), one thread; runs between
// the synchroniser of a chunk and its decode so that chunk c decodes records [mark[c], mark[c+1])
// records of a batch in callback order (completion index, then channel): device sort + permuting copy into dst
cudaError_t pack_sorted_launch(const FrameRec * recs, unsigned int n, FrameRec * dst, cudaStream_t st);
// host_out: the same mark written straight into (mapped) pinned host memory by the kernel -- no copy engine involved
cudaError_t record_mark_launch(const unsigned int * counters, RangeMark * mark_out, cudaStream_t st, int used_at = 2, RangeMark * host_out = nullptr);
// start of a batch: counters[0..1] and [4..7] (and [2..3] unless keep_ring) and the first mark to zero, as a kernel
// (a cudaMemsetAsync of a few bytes is a copy-engine operation and waits behind any bulk copy in flight)
cudaError_t batch_reset_launch(unsigned int * counters, RangeMark * mark0, int keep_ring, cudaStream_t st);

} // namespace b2

namespace b2 {

// ------------------------------------------------------------------ packet encode (transmit)
struct EncodeJob {
    uint8_t  header[8];
    uint32_t payload_len, mod, bps, check, fec0, fec1;
    uint32_t slot;              // output slot (= channel)
    uint64_t payload_offset;    // into EncodeParams::payloads
};
struct EncodeParams {
    const EncodeJob * jobs; unsigned int nframes;
    const uint8_t * payloads;
    uint8_t * header_mod;       // [slots][288] one BPSK bit per byte
    uint8_t * payload_mod;      // [slots][mod_stride] one symbol per byte
    size_t mod_stride;
    uint8_t * work0, * work1;   // [slots][work_stride] scratch
    size_t work_stride;
};
cudaError_t packet_encode_launch(const EncodeParams & p, cudaStream_t st);

// ------------------------------------------------------------------ OFDM frame generator
// ofdmflexframegen_write per channel per symbol period (lib/multichanneltx.cc:230-242,
// lib/ofdmtxrx.cc:328): one CTA per channel walks the symbol periods of this call.
struct GenDesc {                // per channel, per call (host is the source of truth)
    uint32_t first_symbol;      // index of the first symbol to generate (0 = S0a)
    uint32_t n_periods;         // leading periods of this call that carry the frame (rest: zeros)
    uint32_t n_hdr, n_pay;      // header / payload OFDM symbols of the frame
    uint32_t mod, bps, payload_mod_len;
    uint32_t fresh;             // 1: frame starts in this call (clear the taper postfix)
};
struct FramegenParams {
    unsigned int M, cp, taper, M_pilot, M_data;
    float g_data;
    float qam_alpha[9];
    unsigned int nchan, nper;
    const GenDesc * desc;
    cf * postfix;               // [nchan][taper] persistent
    const uint8_t * header_mod; const uint8_t * payload_mod; size_t mod_stride;
    const cf * s0, * s1;        // [M] time-domain training symbols
    const float * taper_w;      // [taper]
    const uint16_t * sc_rank; const uint8_t * pilot_seq;
    cf * out; size_t out_stride, out_off;   // out[c*out_stride + out_off + period*(M+cp) + i]
    FftDev fft;
};
size_t framegen_smem_bytes(const FramegenParams & p);
cudaError_t framegen_launch(const FramegenParams & p, cudaStream_t st);

// ------------------------------------------------------------------ synthesis channelizer
// firpfbch_crcf synthesizer + NCO mix-up (lib/multichanneltx.cc:205-222)
struct SynthParams {
    const cf * in; size_t in_stride, in_off;    // in[c*in_stride + in_off + t], c < N
    unsigned int K, lgK, N, P, TB;
    unsigned int nblocks;
    const float * taps;         // [P][K]
    cf * vhist;                 // [P-1][K] IFFT outputs of the last P-1 blocks (persistent)
    cf * vhist_out;             // where this launch leaves the new history (ping-pong with vhist)
    uint32_t theta0, dtheta;    // NCO phase of output sample 0
    cf * out;                   // [nblocks*K] wideband
    FftDev fft;
};
size_t synth_smem_bytes(const SynthParams & p);
cudaError_t synth_configure(size_t smem_bytes);
cudaError_t synth_launch(const SynthParams & p, int grid, size_t smem_bytes, cudaStream_t st);

// ------------------------------------------------------------------ arbitrary resampler
// msresamp_crcf (arbitrary stage, rate in [0.5, 2]); output k is a closed-form function of k:
// t_k = tau0 + k*step (Q32), input index t_k >> 32 (relative to x[0] = first new sample)
struct ResampParams {
    const cf * x;               // [nx] new samples
    cf * hist_buf;              // [hist] the 2m-1 samples before x[0]; replaced by the launch with the last 2m-1 of this call
    unsigned int hist;          // 2m-1 (<= 31)
    unsigned long long nx;
    const float * h;            // prototype taps [2*m*npfb + 1]
    unsigned int npfb_bits, m2; // log2(npfb), 2m
    unsigned long long tau0, step;
    unsigned long long ny;
    cf * y;
};
cudaError_t resamp_launch(const ResampParams & p, cudaStream_t st);

} // namespace b2
