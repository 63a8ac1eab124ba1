// multichannelrx.cc -- reference-compatible multichannelrx over the B200 C ABI.
// Interface and error behaviour follow lib/multichannelrx.cc:45-195 of the reference
// (messages on stderr + `throw 0` for bad arguments); the DSP is in libb200ofdm.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>

#include "b200_ofdm.h"
#include "multichannelrx.h"

// sample-index side channel of the frame whose callback is running
static thread_local uint64_t tls_detect_index = 0, tls_complete_index = 0;
extern "C" void b2_callback_indices(uint64_t * detect_index, uint64_t * complete_index)
{
    if (detect_index) *detect_index = tls_detect_index;
    if (complete_index) *complete_index = tls_complete_index;
}
extern "C" void b2_set_callback_indices(uint64_t detect_index, uint64_t complete_index)
{
    tls_detect_index = detect_index;
    tls_complete_index = complete_index;
}

static int env_device()
{
    const char * e = getenv("B2_DEVICE");
    return e ? atoi(e) : 0;
}

multichannelrx::multichannelrx(unsigned int _num_channels, unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                               unsigned char * _p, void ** _userdata, framesync_callback * _callback)
    : num_channels(_num_channels), M(_M), cp_len(_cp_len), taper_len(_taper_len), rx(NULL), stage(NULL), stage_len(0), stage_cap(0)
{
    if (_num_channels < 1) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), must have at least one channel\n");
        throw 0;
    } else if (_M < 8) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), number of subcarriers must be at least 8\n");
        throw 0;
    } else if (_cp_len < 1) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), cyclic prefix length must be at least 1\n");
        throw 0;
    } else if (_taper_len > _cp_len) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), taper length cannot exceed cyclic prefix length\n");
        throw 0;
    }
    // the arrays belong to the caller only during construction (lib/multichannelrx.cc:76-82)
    userdata.assign(_userdata, _userdata + num_channels);
    callback.assign(_callback, _callback + num_channels);

    flush_min = 16384; max_wait_ns = 5000000ll; first_ns = 0;
    if (const char * e = getenv("B2_MCRX_FLUSH_MIN")) { unsigned long v = strtoul(e, NULL, 10); if (v >= 1) flush_min = (unsigned int)v; }
    if (const char * e = getenv("B2_MCRX_MAX_LATENCY_MS")) max_wait_ns = (long long)(atof(e) * 1e6);
    stage_cap = 1u << 20;
    if (const char * e = getenv("B2_MCRX_BATCH")) {
        unsigned long v = strtoul(e, NULL, 10);
        if (v >= 1 && v <= (1ul << 26)) stage_cap = (unsigned int)v;
    }
    int rc = b2_mcrx_create(num_channels, M, cp_len, taper_len, _p, env_device(), stage_cap, &rx);
    if (rc != B2_OK) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), %s\n", b2_last_error());
        throw 0;
    }
    stage = (std::complex<float> *)b2_pinned_alloc(sizeof(std::complex<float>) * stage_cap);
    if (stage == NULL) {
        fprintf(stderr, "error: multichannelrx::multichannelrx(), could not allocate the staging buffer\n");
        b2_mcrx_destroy(rx);
        throw 0;
    }
}

multichannelrx::~multichannelrx()
{
    try { Flush(); } catch (...) { }
    b2_pinned_free(stage);
    b2_mcrx_destroy(rx);
}

void multichannelrx::SetBatchSize(unsigned int _num_samples)
{
    Flush();
    if (_num_samples < 1) _num_samples = 1;
    if (_num_samples > stage_cap) {
        std::complex<float> * n = (std::complex<float> *)b2_pinned_alloc(sizeof(std::complex<float>) * _num_samples);
        if (n == NULL) return;
        b2_pinned_free(stage);
        stage = n;
    }
    stage_cap = _num_samples;
}

void multichannelrx::Reset()
{
    // everything pushed before the reset has been seen by the synchronizers in the reference
    Flush();
    if (b2_mcrx_reset(rx) != B2_OK) {
        fprintf(stderr, "error: multichannelrx::Reset(), %s\n", b2_last_error());
        throw 0;
    }
}

static inline long long now_ns()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

void multichannelrx::Execute(std::complex<float> * _x, unsigned int _num_samples)
{
    if (_num_samples == 0) return;
    if (stage_len == 0 && max_wait_ns > 0) first_ns = now_ns();
    unsigned int i = 0;
    while (i < _num_samples) {
        unsigned int c = _num_samples - i;
        if (c > stage_cap - stage_len) c = stage_cap - stage_len;
        if (c == 1) stage[stage_len] = _x[i];
        else memcpy(stage + stage_len, _x + i, sizeof(std::complex<float>) * c);
        stage_len += c;
        i += c;
        if (stage_len == stage_cap) { Flush(); if (max_wait_ns > 0) first_ns = now_ns(); }
    }
    // bounded callback latency: a call that is a packet by itself is processed now, and so is whatever has been waiting
    // for longer than the bound (the one-sample-per-call pattern of src/multichannel_rx.cc:211 keeps its batching)
    if (stage_len && (_num_samples >= flush_min || (max_wait_ns > 0 && now_ns() - first_ns >= max_wait_ns))) Flush();
}

void multichannelrx::Flush()
{
    if (stage_len) {
        int rc = b2_mcrx_execute(rx, (const float *)stage, stage_len);
        stage_len = 0;
        if (rc != B2_OK) {
            fprintf(stderr, "error: multichannelrx::Execute(), %s\n", b2_last_error());
            throw 0;
        }
    }
    Deliver();
}

// replay the user callbacks: ascending completion block, then channel (lib/multichannelrx.cc:193-194)
void multichannelrx::Deliver()
{
    // zero-copy view of the frames completed so far; pointers stay valid during the callbacks
    // because nothing below touches the handle
    size_t n = 0, nb = 0;
    const b2_frame_rec * recs = NULL;
    const uint8_t * payloads = NULL;
    if (b2_mcrx_poll_view(rx, &recs, &n, &payloads, &nb) != B2_OK) {
        fprintf(stderr, "error: multichannelrx::Execute(), %s\n", b2_last_error());
        throw 0;
    }
    for (size_t i = 0; i < n; i++) {
        const b2_frame_rec & r = recs[i];
        if (r.channel >= num_channels || callback[r.channel] == NULL) continue;
        framesyncstats_s stats;
        stats.evm = r.evm; stats.rssi = r.rssi; stats.cfo = r.cfo;
        stats.framesyms = NULL; stats.num_framesyms = 0;
        stats.mod_scheme = r.mod_scheme; stats.mod_bps = r.mod_bps;
        stats.check = r.check; stats.fec0 = r.fec0; stats.fec1 = r.fec1;
        unsigned char header[8];
        memcpy(header, r.header, 8);
        unsigned char * payload = (r.header_valid && r.payload_len) ? (unsigned char *)payloads + r.payload_offset : NULL;
        tls_detect_index = r.detect_index;
        tls_complete_index = r.complete_index;
        callback[r.channel](header, r.header_valid, payload, r.payload_len, r.payload_valid, stats, userdata[r.channel]);
    }
}
