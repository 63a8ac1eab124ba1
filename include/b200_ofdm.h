/*
 * b200_ofdm.h -- C ABI of the B200 (sm_100a) multichannel OFDM DSP library (libb200ofdm.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, every function
 * returns an int status (0 = ok, < 0 = error, never throws).  Each entry point replaces one
 * call (or one fixed group of calls) the reference makes into liquid-dsp from
 * lib/multichannelrx.cc, lib/multichanneltx.cc and lib/ofdmtxrx.cc; the C++ classes of the same
 * names in liquid-usrp_b200/host/ are thin wrappers over this file (they turn a status into the
 * reference's `throw 0` and replay user callbacks).
 *
 * All sample buffers are interleaved complex float32 (re, im), the layout of
 * std::complex<float> / liquid_float_complex / uhd::io_type_t::COMPLEX_FLOAT32.
 * "host" pointers may be pageable or pinned; "device" pointers are CUDA device pointers on the
 * handle's device.  A handle is driven by one thread at a time (as in the reference, §8b).
 */
#ifndef B200_OFDM_H
#define B200_OFDM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define B2_OK             0
#define B2_ERR_ARG       -1   /* invalid argument (what the reference answers with `throw 0`)     */
#define B2_ERR_UNSUPPORTED -2 /* legal for liquid-dsp, outside what the CUDA path implements      */
#define B2_ERR_CUDA      -3   /* CUDA runtime error; b2_last_error() has the text                 */
#define B2_ERR_NOMEM     -4
#define B2_ERR_STATE     -5   /* call not legal in the current state (e.g. channel not ready)     */
#define B2_ERR_OVERFLOW  -6   /* caller-provided output buffer too small                          */

const char * b2_last_error(void);           /* thread-local text of the last failure             */
const char * b2_version(void);
int  b2_device_count(void);                 /* number of visible CUDA devices (0 = none)          */
/* page-locked host memory for staging buffers handed to the *_execute / *_generate calls */
void * b2_pinned_alloc(size_t bytes);
void   b2_pinned_free(void * p);

/* One decoded frame.  Mirrors the arguments of liquid's framesync_callback
 * (src/multichannel_rx.cc:37-43) plus the sample-index side channel the north star asks for. */
typedef struct {
    uint32_t channel;          /* which ofdmflexframesync produced it                              */
    int32_t  header_valid;
    int32_t  payload_valid;
    uint32_t payload_len;      /* bytes; 0 when the header was invalid                             */
    uint8_t  header[8];        /* user header bytes                                                */
    float    evm, rssi, cfo;   /* framesyncstats_s.{evm,rssi,cfo}                                  */
    uint32_t mod_scheme, mod_bps, check, fec0, fec1;   /* framesyncstats_s fields                  */
    uint64_t detect_index;     /* channel-rate index of the sample that tripped frame detection    */
    uint64_t complete_index;   /* channel-rate index of the sample that completed the frame        */
    uint64_t payload_offset;   /* byte offset of the payload in the payload buffer of that poll    */
} b2_frame_rec;

/* ------------------------------------------------------------------ multichannelrx
 * replaces: multichannelrx::multichannelrx  lib/multichannelrx.cc:45-104
 *             (N x ofdmflexframesync_create, firpfbch_crcf_create_kaiser(ANALYZER,2N,7,60),
 *              nco_crcf_create + set_frequency)
 *           multichannelrx::Execute         lib/multichannelrx.cc:155-182
 *             (nco_crcf_mix_down/step per sample, firpfbch_crcf_analyzer_execute per 2N samples,
 *              N x ofdmflexframesync_execute(.,1))
 *           multichannelrx::Reset           lib/multichannelrx.cc:135-153
 *           multichannelrx::~multichannelrx lib/multichannelrx.cc:107-132                          */
typedef struct b2_mcrx_s b2_mcrx;

/* p: subcarrier allocation (M bytes, OFDMFRAME_SCTYPE_*) or NULL for liquid's default.
 * device: CUDA device ordinal.  max_batch: largest number of wideband samples one
 * b2_mcrx_execute call will be given (0 = default 2^22); larger calls are split internally. */
int b2_mcrx_create(unsigned int num_channels, unsigned int M, unsigned int cp_len, unsigned int taper_len,
                   const unsigned char * p, int device, size_t max_batch, b2_mcrx ** out);
int b2_mcrx_destroy(b2_mcrx * q);
int b2_mcrx_reset(b2_mcrx * q);
/* push n wideband samples (host memory); any n is legal, state carries across calls */
int b2_mcrx_execute(b2_mcrx * q, const float * x_host, size_t n);
/* same, samples already resident in device memory (16-byte aligned) */
int b2_mcrx_execute_device(b2_mcrx * q, const float * x_dev, size_t n);
/* rate-matching stage ahead of the receiver: every later execute call resamples its input by `rate`
 * (msresamp_crcf, stop-band As dB) on the device before the NCO / channelizer -- the step the
 * reference's programs compute (src/multichannel_rx.cc:137-138: usrp rate / wanted rate) but leave
 * as a TODO on this path (lib/multichanneltxrx.cc:605).  rate = 0 removes the stage. */
int b2_mcrx_set_resampler(b2_mcrx * q, float rate, float As);
/* frames completed by the execute calls since the last poll, in the reference's callback order
 * (ascending completion block, then channel).  Pass recs = NULL to get the counts only. */
int b2_mcrx_poll(b2_mcrx * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                 uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes);
/* zero-copy variant: pointers into the handle's own buffers, valid until the next execute, poll
 * or poll_view on this handle */
int b2_mcrx_poll_view(b2_mcrx * q, const b2_frame_rec ** recs, size_t * n_recs,
                      const uint8_t ** payloads, size_t * n_payload_bytes);
/* debug tap: record every equalised OFDM symbol X[0..M) handed to the header/payload layer */
int b2_mcrx_tap_symbols(b2_mcrx * q, int enable, size_t max_symbols);
int b2_mcrx_read_symbols(b2_mcrx * q, uint32_t * channel, uint64_t * index, float * X, size_t cap, size_t * n);
/* device-side timing of the last execute call, milliseconds (CUDA events on the streams the
 * kernels are launched on; the stages of successive chunks overlap, so [0..2] are sums over the
 * chunks of the call and may add up to more than [3]): [0] channelizer kernels, [1] sync kernels,
 * [2] packet-decode kernels, [3] whole call */
int b2_mcrx_last_timing(b2_mcrx * q, float ms[4]);
/* number of kernels launched by the last execute call and number of pipeline chunks it used */
int b2_mcrx_last_launches(b2_mcrx * q, unsigned int * kernels, unsigned int * chunks);
/* raw access for tests: channelizer output of the last execute call, [num_channels][n_blocks] */
int b2_mcrx_read_channelizer(b2_mcrx * q, float * out, size_t cap_samples, size_t * n_blocks);
/* the CUDA stream the kernels of this handle are launched on (cudaStream_t) */
void * b2_mcrx_stream(b2_mcrx * q);
/* Stage 1 alone, stateless, for a time shard of the wideband stream (multi-GPU):
 * x_dev points at the first of (P-1) = 13 halo blocks that precede the shard (zeros at the very
 * start of a stream), followed by n_blocks blocks of 2N samples; sample_offset is the absolute
 * index of x_dev[0] in the stream (the NCO phase is exact in it).  Writes out_dev[c*out_stride + b],
 * c < N, b < n_blocks, on the handle's stream, asynchronously. */
int b2_mcrx_channelize_device(b2_mcrx * q, const float * x_dev, size_t n_blocks, int64_t sample_offset,
                              float * out_dev, size_t out_stride);
/* Stage 2 alone on channelizer output already in device memory: in_dev[c*in_stride + t], c < N,
 * t < n (the handle's per-channel synchroniser state carries over, so a time-ordered sequence of
 * calls equals one call); frames are queued for b2_mcrx_poll. */
int b2_mcrx_sync_device(b2_mcrx * q, const float * in_dev, size_t n, size_t in_stride);

/* ------------------------------------------------------------------ one rank of a multi-GPU multichannelrx
 * The reference runs multichannelrx::Execute (lib/multichannelrx.cc:155-195) on one core and leaves "run each
 * channel in its own thread" as a TODO (:184, :193-194).  Here ONE wideband stream is spread over `world` GPUs of a
 * box, one process per GPU:
 *   stage 1  NCO + analysis bank (lib/multichannelrx.cc:163-164,188), sharded over TIME in round-robin chunks of
 *            chunk_blocks blocks of 2N samples: the rank handles chunks rank, rank + world, rank + 2 world, ...
 *   scatter  fused into stage 1: every channel's run is stored straight into the memory of the GPU that owns the
 *            channel (peer mapping over NVLink, handles exchanged with _export / _connect); no staging, no repack
 *   stage 2  ofdmflexframesync + packet decode (lib/multichannelrx.cc:193-194), sharded over CHANNELS: the rank owns
 *            channels [rank N/world, (rank+1) N/world) for all time and runs them over `world` exchanged chunks per step
 * The caller (one per rank, e.g. liquid-usrp_b200/sharded.py over torch.distributed) provides the two CUDA streams
 * and the only cross-rank ordering the data path needs: a barrier among the ranks behind stage1(step) (a one-word
 * NCCL all-reduce), which stage2(step) must wait for; and, the exchange buffer having B2_SHARD_SLOTS slots,
 * stage1(step + B2_SHARD_SLOTS - 1) must not start before every rank's stage2(step) is done.  A rank that joins
 * barrier(j) only behind its own stage2(j - B2_SHARD_SLOTS + 2) gets that from barrier(step + B2_SHARD_SLOTS - 2)
 * -- one step of slack, so stage 1 never waits for the barrier of its own step. */
typedef struct b2_mcrx_shard_s b2_mcrx_shard;
#define B2_SHARD_SLOTS 4
#define B2_SHARD_HALO_BLOCKS 13     /* taps per branch - 1 of the receive bank (m = 7, lib/multichannelrx.cc:89) */
/* steps_per_call: steps (of world chunks each) between _begin and _end, sizes the frame output */
int b2_mcrx_shard_create(unsigned int num_channels, unsigned int M, unsigned int cp_len, unsigned int taper_len,
                         const unsigned char * p, int device, unsigned int rank, unsigned int world,
                         size_t chunk_blocks, size_t steps_per_call, void * stream_stage1, void * stream_stage2,
                         b2_mcrx_shard ** out);
int b2_mcrx_shard_destroy(b2_mcrx_shard * q);
/* 64 opaque bytes naming this rank's exchange buffer (cudaIpcMemHandle_t); gather them from all ranks ... */
int b2_mcrx_shard_export(b2_mcrx_shard * q, void * handle64);
/* ... and hand every rank the whole list, in rank order (world x 64 bytes) */
int b2_mcrx_shard_connect(b2_mcrx_shard * q, const void * handles, size_t n_handles);
int b2_mcrx_shard_begin(b2_mcrx_shard * q);
/* stage 1 + scatter of this rank's chunk of step `step` (absolute chunk index step*world + rank since the stream
 * began): x_dev points at the 13 halo blocks before the chunk (zeros at the start of a stream) followed by
 * chunk_blocks blocks; asynchronous on stream_stage1 */
int b2_mcrx_shard_stage1(b2_mcrx_shard * q, const float * x_dev, uint64_t step);
/* stage 2 over the world chunks of `step`, asynchronous on stream_stage2 (+ the handle's decode / copy streams) */
int b2_mcrx_shard_stage2(b2_mcrx_shard * q, uint64_t step);
/* end of the call: waits for stage 2 of every step since _begin; frames are queued for _poll (channel = global index) */
int b2_mcrx_shard_end(b2_mcrx_shard * q);
int b2_mcrx_shard_poll(b2_mcrx_shard * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                       uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes);
int b2_mcrx_shard_poll_view(b2_mcrx_shard * q, const b2_frame_rec ** recs, size_t * n_recs,
                            const uint8_t ** payloads, size_t * n_payload_bytes);
/* the frames of the call that just ended, still in DEVICE memory, packed for a collective: dst_dev receives
 * [uint64 n_recs, n_payload_bytes, seq, 0 | n_recs records of sizeof(b2_frame_rec) | payload bytes] (seq: the caller's
 * tag of this pack; a reader of a copy that lands body first, header last can poll it), payload_offset
 * relative to the payload part, records in this rank's callback order (completion index, then channel: sorted on the
 * device); asynchronous on stream_stage2 */
int b2_mcrx_shard_pack_results(b2_mcrx_shard * q, void * dst_dev, size_t cap_bytes, uint64_t seq, size_t * n_recs, size_t * n_payload_bytes);
#define B2_SHARD_PACK_HEADER 32
/* on = 0: _end no longer brings this rank's frames to its own host memory (nor orders them there); they are taken from
 * device memory with _pack_results instead -- the mode of a box where one rank collects everybody's frames.  Default 1. */
int b2_mcrx_shard_host_results(b2_mcrx_shard * q, int on);
int b2_mcrx_shard_reset(b2_mcrx_shard * q);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): for plumbing that moves a pack into registered host memory */
int b2_memcpy_async(void * dst, const void * src, size_t bytes, void * stream);

/* ------------------------------------------------------------------ multichanneltx
 * replaces: multichanneltx::multichanneltx        lib/multichanneltx.cc:41-100
 *           multichanneltx::IsChannelReadyForData lib/multichanneltx.cc:152-162
 *           multichanneltx::UpdateData            lib/multichanneltx.cc:165-189
 *             (ofdmflexframegen_setprops + ofdmflexframegen_assemble)
 *           multichanneltx::GenerateSamples       lib/multichanneltx.cc:192-227
 *             (ofdmflexframegen_write per channel per symbol, firpfbch_crcf_synthesizer_execute,
 *              nco_crcf_mix_up/step)
 *           multichanneltx::Reset                 lib/multichanneltx.cc:126-149                    */
typedef struct b2_mctx_s b2_mctx;

int b2_mctx_create(unsigned int num_channels, unsigned int M, unsigned int cp_len, unsigned int taper_len,
                   const unsigned char * p, int device, b2_mctx ** out);
int b2_mctx_destroy(b2_mctx * q);
int b2_mctx_reset(b2_mctx * q);
/* move the NCO by n_samples output samples (negative: back).  multichanneltx::Reset leaves the NCO alone
 * (lib/multichanneltx.cc:135): a caller that generated ahead and throws samples away at a reset takes them back here. */
int b2_mctx_nco_advance(b2_mctx * q, int64_t n_samples);
int b2_mctx_is_ready(b2_mctx * q, unsigned int channel, int * ready);
int b2_mctx_update(b2_mctx * q, unsigned int channel, const unsigned char * header,
                   const unsigned char * payload, unsigned int payload_len, int mod, int fec0, int fec1);
/* multichanneltx::UpdateData (lib/multichanneltx.cc:165-189) for n channels in one call -- what the loop of
 * src/multichannel_tx.cc:166-185 does at a frame boundary: headers is n x 8 bytes, payloads the n payloads back to
 * back (payload_lens[i] bytes each).  Channels that are not ready are skipped like UpdateData skips them; *n_updated
 * (nullable) counts the ones taken. */
int b2_mctx_update_many(b2_mctx * q, unsigned int n, const unsigned int * channels, const unsigned char * headers,
                        const unsigned char * payloads, const unsigned int * payload_lens, int mod, int fec0, int fec1,
                        unsigned int * n_updated);
/* produce the next n_calls * 2N wideband samples (n_calls consecutive GenerateSamples calls with
 * no UpdateData in between) into host memory */
int b2_mctx_generate(b2_mctx * q, float * out_host, size_t n_calls);
int b2_mctx_generate_device(b2_mctx * q, float * out_dev, size_t n_calls);
/* number of GenerateSamples calls until the next OFDM symbol boundary (where readiness of the
 * channels can change): in [1, M+cp] */
int b2_mctx_calls_to_boundary(b2_mctx * q, size_t * n_calls);
int b2_mctx_last_timing(b2_mctx * q, float ms[4]);

/* ------------------------------------------------------------------ single-link framer
 * replaces, for lib/ofdmtxrx.cc: ofdmflexframegen_{create,setprops,assemble,write,writesymbol,
 * is_assembled,reset,destroy} (lib/ofdmtxrx.cc:79-84,314-328,377-387) and
 * ofdmflexframesync_{create,execute,reset,destroy} (lib/ofdmtxrx.cc:91,482,625).  Implemented as
 * the one-channel, no-channelizer case of the kernels above.                                      */
typedef struct b2_ofdmgen_s b2_ofdmgen;
typedef struct b2_ofdmsync_s b2_ofdmsync;

int b2_ofdmgen_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p,
                      int device, b2_ofdmgen ** out);
int b2_ofdmgen_destroy(b2_ofdmgen * q);
int b2_ofdmgen_reset(b2_ofdmgen * q);
int b2_ofdmgen_is_assembled(b2_ofdmgen * q, int * assembled);
/* number of OFDM symbols (each M+cp samples) of the frame last assembled, including the tail buffer */
int b2_ofdmgen_assemble(b2_ofdmgen * q, const unsigned char * header, const unsigned char * payload,
                        unsigned int payload_len, int check, int fec0, int fec1, int mod,
                        unsigned int * n_symbols);
/* write the next n_symbols * (M+cp) samples of the assembled frame; *last = 1 once the frame is done */
int b2_ofdmgen_write(b2_ofdmgen * q, float * out_host, unsigned int n_symbols, int * last);

/* `streams` independent sample streams, each synchronised by its own ofdmflexframesync state
 * (streams = 1 is the ofdmtxrx receiver; streams > 1 is the batched form used by config 4) */
int b2_ofdmsync_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p,
                       unsigned int streams, int device, size_t max_batch, b2_ofdmsync ** out);
int b2_ofdmsync_destroy(b2_ofdmsync * q);
int b2_ofdmsync_reset(b2_ofdmsync * q);
/* x: [streams][n] samples, stream-major */
int b2_ofdmsync_execute(b2_ofdmsync * q, const float * x_host, size_t n);
int b2_ofdmsync_execute_device(b2_ofdmsync * q, const float * x_dev, size_t n, size_t stride);
int b2_ofdmsync_poll(b2_ofdmsync * q, b2_frame_rec * recs, size_t recs_cap, size_t * n_recs,
                     uint8_t * payloads, size_t payloads_cap, size_t * n_payload_bytes);
int b2_ofdmsync_poll_view(b2_ofdmsync * q, const b2_frame_rec ** recs, size_t * n_recs,
                          const uint8_t ** payloads, size_t * n_payload_bytes);
int b2_ofdmsync_last_timing(b2_ofdmsync * q, float ms[4]);

/* ------------------------------------------------------------------ msresamp_crcf
 * replaces msresamp_crcf_{create,execute,reset,destroy} as used by src/flexframe_rx.cc:179,240
 * (rates in [0.5, 2]: the arbitrary polyphase stage only).                                        */
typedef struct b2_msresamp_s b2_msresamp;
int b2_msresamp_create(float rate, float As, int device, b2_msresamp ** out);
int b2_msresamp_destroy(b2_msresamp * q);
int b2_msresamp_reset(b2_msresamp * q);
int b2_msresamp_execute(b2_msresamp * q, const float * x_host, size_t nx, float * y_host, size_t y_cap, size_t * ny);
/* device pointers: the whole call is one kernel launch over the caller's buffers (nx < 2^31) */
int b2_msresamp_execute_device(b2_msresamp * q, const float * x_dev, size_t nx, float * y_dev, size_t y_cap, size_t * ny);
/* host samples in, device samples out (feeds b2_mcrx_execute_device / b2_ofdmsync_execute_device) */
int b2_msresamp_execute_to_device(b2_msresamp * q, const float * x_host, size_t nx, float * y_dev, size_t y_cap, size_t * ny);

#ifdef __cplusplus
}
#endif
#endif /* B200_OFDM_H */
