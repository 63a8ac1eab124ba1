// mc_shim.cc -- test driver: a flat C interface over the reference's C++ class API
// (include/multichannelrx.h:29-83, include/multichanneltx.h:29-93) so that Python/ctypes
// can drive either
//   * the reference's own lib/multichannelrx.cc + lib/multichanneltx.cc compiled over the CPU
//     oracle (oracle/_ref/libref_mc.so, -DMC_SHIM_ORACLE), or
//   * this repo's CUDA-backed classes of the same names (libliquidusrp_b200.so)
// with identical calls.  Frames are recorded from inside the user callback exactly the way
// src/multichannel_rx.cc:37-66 consumes them.
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <liquid/liquid.h>
#include "multichannelrx.h"
#include "multichanneltx.h"

extern "C" {
#ifdef MC_SHIM_ORACLE
void orc_callback_indices(uint64_t * detect_index, uint64_t * complete_index);
typedef void (*orc_symbol_tap_fn)(void * userdata, const liquid_float_complex * X, unsigned int M, uint64_t sample_index);
void orc_set_symbol_tap(orc_symbol_tap_fn fn);
#define CALLBACK_INDICES orc_callback_indices
#else
void b2_callback_indices(uint64_t * detect_index, uint64_t * complete_index);
#define CALLBACK_INDICES b2_callback_indices
#endif
}

struct shim_frame {
    uint32_t channel;
    int32_t  header_valid;
    int32_t  payload_valid;
    uint32_t payload_len;
    uint8_t  header[8];
    float    evm, rssi, cfo;
    uint32_t mod_scheme, mod_bps, check, fec0, fec1;
    uint64_t detect_index, complete_index;
    uint64_t payload_offset;
};

struct shim_rx;
struct chan_ud { shim_rx * owner; unsigned int channel; };
struct shim_sym { uint32_t channel; uint64_t index; };

struct shim_rx {
    multichannelrx * rx = nullptr;
    unsigned int N = 0, M = 0;
    std::vector<chan_ud> ud;
    std::vector<shim_frame> frames;
    std::vector<uint8_t> payloads;
    bool tap = false;
    std::vector<shim_sym> syms;
    std::vector<float> symdata;
};

static int shim_callback(unsigned char * header, int header_valid, unsigned char * payload,
                         unsigned int payload_len, int payload_valid, framesyncstats_s stats, void * userdata)
{
    chan_ud * u = (chan_ud *)userdata;
    shim_rx * s = u->owner;
    shim_frame f;
    memset(&f, 0, sizeof(f));
    f.channel = u->channel;
    f.header_valid = header_valid;
    f.payload_valid = payload_valid;
    f.payload_len = payload_len;
    if (header) memcpy(f.header, header, 8);
    f.evm = stats.evm; f.rssi = stats.rssi; f.cfo = stats.cfo;
    f.mod_scheme = stats.mod_scheme; f.mod_bps = stats.mod_bps;
    f.check = stats.check; f.fec0 = stats.fec0; f.fec1 = stats.fec1;
    CALLBACK_INDICES(&f.detect_index, &f.complete_index);
    f.payload_offset = s->payloads.size();
    if (payload && payload_len) s->payloads.insert(s->payloads.end(), payload, payload + payload_len);
    s->frames.push_back(f);
    return 0;
}

#ifdef MC_SHIM_ORACLE
static void shim_symbol_tap(void * userdata, const liquid_float_complex * X, unsigned int M, uint64_t sample_index)
{
    chan_ud * u = (chan_ud *)userdata;
    shim_rx * s = u->owner;
    if (!s->tap) return;
    shim_sym e = {u->channel, sample_index};
    s->syms.push_back(e);
    const float * xf = (const float *)X;
    s->symdata.insert(s->symdata.end(), xf, xf + 2 * M);
}
#endif

extern "C" {

void * mcshim_rx_create(unsigned int N, unsigned int M, unsigned int cp, unsigned int taper)
{
    shim_rx * s = new shim_rx;
    s->N = N; s->M = M;
    s->ud.resize(N ? N : 1);
    std::vector<void *> udp(N ? N : 1);
    std::vector<framesync_callback> cbs(N ? N : 1);
    for (unsigned int i = 0; i < N; i++) {
        s->ud[i].owner = s; s->ud[i].channel = i;
        udp[i] = &s->ud[i]; cbs[i] = shim_callback;
    }
    try {
        s->rx = new multichannelrx(N, M, cp, taper, NULL, udp.data(), cbs.data());
    } catch (...) {
        delete s;
        return NULL;
    }
    return s;
}

void mcshim_rx_destroy(void * h) { shim_rx * s = (shim_rx *)h; delete s->rx; delete s; }
void mcshim_rx_reset(void * h) { ((shim_rx *)h)->rx->Reset(); }

// push n samples, `chunk` per Execute() call (chunk = 1 is the reference binaries' pattern,
// src/multichannel_rx.cc:211)
void mcshim_rx_execute(void * h, float * x, uint64_t n, uint64_t chunk)
{
    shim_rx * s = (shim_rx *)h;
    if (chunk == 0) chunk = n;
    uint64_t i = 0;
    while (i < n) {
        uint64_t c = (n - i < chunk) ? n - i : chunk;
        s->rx->Execute((std::complex<float> *)(x + 2 * i), (unsigned int)c);
        i += c;
    }
#ifndef MC_SHIM_ORACLE
    s->rx->Flush();
#endif
}

uint64_t mcshim_rx_num_frames(void * h) { return ((shim_rx *)h)->frames.size(); }
uint64_t mcshim_rx_payload_bytes(void * h) { return ((shim_rx *)h)->payloads.size(); }
void mcshim_rx_get_frames(void * h, shim_frame * out)
{
    shim_rx * s = (shim_rx *)h;
    if (!s->frames.empty()) memcpy(out, s->frames.data(), s->frames.size() * sizeof(shim_frame));
}
void mcshim_rx_get_payloads(void * h, uint8_t * out)
{
    shim_rx * s = (shim_rx *)h;
    if (!s->payloads.empty()) memcpy(out, s->payloads.data(), s->payloads.size());
}
void mcshim_rx_clear(void * h)
{
    shim_rx * s = (shim_rx *)h;
    s->frames.clear(); s->payloads.clear(); s->syms.clear(); s->symdata.clear();
}

int mcshim_rx_enable_symbol_tap(void * h, int enable)
{
#ifdef MC_SHIM_ORACLE
    ((shim_rx *)h)->tap = enable != 0;
    orc_set_symbol_tap(shim_symbol_tap);
    return 0;
#else
    (void)h; (void)enable;
    return -1;
#endif
}
uint64_t mcshim_rx_num_symbols(void * h) { return ((shim_rx *)h)->syms.size(); }
void mcshim_rx_get_symbols(void * h, uint32_t * channel, uint64_t * index, float * X)
{
    shim_rx * s = (shim_rx *)h;
    for (size_t i = 0; i < s->syms.size(); i++) { channel[i] = s->syms[i].channel; index[i] = s->syms[i].index; }
    if (!s->symdata.empty()) memcpy(X, s->symdata.data(), s->symdata.size() * sizeof(float));
}

// ------------------------------------------------------------------ transmitter
struct shim_tx {
    multichanneltx * tx = nullptr;
    unsigned int N = 0;
    std::vector<uint32_t> pid;
};

static uint64_t splitmix64(uint64_t * state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// deterministic frame contents shared by every test/bench leg:
// header = {pid_hi, pid_lo, channel, 5 PRNG bytes} (mirrors src/multichannel_tx.cc:172-176),
// payload bytes from SplitMix64(seed, channel, pid)
void mcshim_frame_data(uint64_t seed, unsigned int channel, unsigned int pid,
                       unsigned char * header, unsigned char * payload, unsigned int payload_len)
{
    uint64_t st = seed + 0x1000003ull * channel + 0x7fffffffull * (uint64_t)pid;
    header[0] = (pid >> 8) & 0xff; header[1] = pid & 0xff; header[2] = channel & 0xff;
    uint64_t r = splitmix64(&st);
    for (int i = 0; i < 5; i++) header[3 + i] = (r >> (8 * i)) & 0xff;
    for (unsigned int i = 0; i < payload_len; i += 8) {
        r = splitmix64(&st);
        for (unsigned int j = 0; j < 8 && i + j < payload_len; j++) payload[i + j] = (r >> (8 * j)) & 0xff;
    }
}

void * mcshim_tx_create(unsigned int N, unsigned int M, unsigned int cp, unsigned int taper)
{
    shim_tx * s = new shim_tx;
    s->N = N;
    s->pid.assign(N ? N : 1, 0);
    try {
        s->tx = new multichanneltx(N, M, cp, taper, NULL);
    } catch (...) {
        delete s;
        return NULL;
    }
    return s;
}
void mcshim_tx_destroy(void * h) { shim_tx * s = (shim_tx *)h; delete s->tx; delete s; }
void mcshim_tx_reset(void * h) { ((shim_tx *)h)->tx->Reset(); }
int mcshim_tx_is_ready(void * h, unsigned int c)
{
    try { return ((shim_tx *)h)->tx->IsChannelReadyForData(c); } catch (...) { return -1; }
}
int mcshim_tx_update(void * h, unsigned int c, unsigned char * header, unsigned char * payload,
                     unsigned int len, int mod, int fec0, int fec1)
{
    try { ((shim_tx *)h)->tx->UpdateData(c, header, payload, len, mod, fec0, fec1); } catch (...) { return -1; }
    return 0;
}
// ncalls x GenerateSamples (2N samples each), no new data
void mcshim_tx_generate(void * h, float * out, uint64_t ncalls)
{
    shim_tx * s = (shim_tx *)h;
    for (uint64_t i = 0; i < ncalls; i++)
        s->tx->GenerateSamples((std::complex<float> *)(out + 2 * i * 2 * s->N));
}
// the src/multichannel_tx.cc:163-216 loop: re-arm every ready channel (subject to
// channel_mask bit c%64 and max_frames per channel), then GenerateSamples; ncalls times
void mcshim_tx_run(void * h, float * out, uint64_t ncalls, unsigned int payload_len,
                   int mod, int fec0, int fec1, uint64_t seed, uint64_t channel_mask, unsigned int max_frames,
                   float gain)
{
    shim_tx * s = (shim_tx *)h;
    std::vector<unsigned char> payload(payload_len ? payload_len : 1);
    unsigned char header[8];
    unsigned int K = 2 * s->N;
    for (uint64_t i = 0; i < ncalls; i++) {
        for (unsigned int c = 0; c < s->N; c++) {
            if (!((channel_mask >> (c % 64)) & 1ull)) continue;
            if (max_frames && s->pid[c] >= max_frames) continue;
            if (s->tx->IsChannelReadyForData(c)) {
                mcshim_frame_data(seed, c, s->pid[c], header, payload.data(), payload_len);
                s->tx->UpdateData(c, header, payload.data(), payload_len, mod, fec0, fec1);
                s->pid[c]++;
            }
        }
        float * o = out + 2 * i * K;
        s->tx->GenerateSamples((std::complex<float> *)o);
        if (gain != 1.0f) for (unsigned int j = 0; j < 2 * K; j++) o[j] *= gain;
    }
}

} // extern "C"
