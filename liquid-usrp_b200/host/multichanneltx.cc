// multichanneltx.cc -- reference-compatible multichanneltx over the B200 C ABI.
// Interface and error behaviour follow lib/multichanneltx.cc:41-242 of the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "b200_ofdm.h"
#include "multichanneltx.h"

static int env_device()
{
    const char * e = getenv("B2_DEVICE");
    return e ? atoi(e) : 0;
}

multichanneltx::multichanneltx(unsigned int _num_channels, unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                               unsigned char * _p)
    : num_channels(_num_channels), M(_M), cp_len(_cp_len), taper_len(_taper_len), tx(NULL), fifo_pos(0)
{
    if (_num_channels < 1) {
        fprintf(stderr, "error: multichanneltx::multichanneltx(), must have at least one channel\n");
        throw 0;
    } else if (_M < 8) {
        fprintf(stderr, "error: multichanneltx::multichanneltx(), number of subcarriers must be at least 8\n");
        throw 0;
    } else if (_cp_len < 1) {
        fprintf(stderr, "error: multichanneltx::multichanneltx(), cyclic prefix length must be at least 1\n");
        throw 0;
    } else if (_taper_len > _cp_len) {
        fprintf(stderr, "error: multichanneltx::multichanneltx(), taper length cannot exceed cyclic prefix length\n");
        throw 0;
    }
    int rc = b2_mctx_create(num_channels, M, cp_len, taper_len, _p, env_device(), &tx);
    if (rc != B2_OK) {
        fprintf(stderr, "error: multichanneltx::multichanneltx(), %s\n", b2_last_error());
        throw 0;
    }
}

multichanneltx::~multichanneltx() { b2_mctx_destroy(tx); }

void multichanneltx::Reset()
{
    // samples generated ahead of the caller are dropped: the NCO, which a reset leaves alone (lib/multichanneltx.cc:135),
    // goes back to the phase of the last sample actually handed out
    const long long ahead = (long long)fifo.size() - (long long)fifo_pos;
    fifo.clear();
    fifo_pos = 0;
    if (ahead > 0) b2_mctx_nco_advance(tx, -ahead);
    if (b2_mctx_reset(tx) != B2_OK) {
        fprintf(stderr, "error: multichanneltx::Reset(), %s\n", b2_last_error());
        throw 0;
    }
}

int multichanneltx::IsChannelReadyForData(unsigned int _channel)
{
    if (_channel >= num_channels) {
        fprintf(stderr, "error: multichanneltx:IsChannelReadyForData(%u), invalid channel id\n", _channel);
        throw 0;
    }
    int ready = 0;
    b2_mctx_is_ready(tx, _channel, &ready);
    return ready;
}

void multichanneltx::UpdateData(unsigned int _channel, unsigned char * _header, unsigned char * _payload,
                                unsigned int _payload_len, int _mod, int _fec0, int _fec1)
{
    if (_channel >= num_channels) {
        fprintf(stderr, "error: multichanneltx:UpdateData(%u), invalid channel id\n", _channel);
        throw 0;
    } else if (!IsChannelReadyForData(_channel)) {
        fprintf(stderr, "warning: multichanneltx:UpdateData(%u), channel not ready yet\n", _channel);
        return;
    }
    int rc = b2_mctx_update(tx, _channel, _header, _payload, _payload_len, _mod, _fec0, _fec1);
    if (rc != B2_OK) {
        fprintf(stderr, "error: multichanneltx:UpdateData(%u), %s\n", _channel, b2_last_error());
        throw 0;
    }
}

void multichanneltx::GenerateSamples(std::complex<float> * _buffer)
{
    const size_t K = 2 * (size_t)num_channels;
    if (fifo_pos >= fifo.size()) {
        // everything up to the next symbol boundary; nothing the caller does in between can
        // change these samples
        size_t n_calls = 0;
        b2_mctx_calls_to_boundary(tx, &n_calls);
        fifo.resize(n_calls * K);
        fifo_pos = 0;
        if (b2_mctx_generate(tx, (float *)fifo.data(), n_calls) != B2_OK) {
            fprintf(stderr, "error: multichanneltx::GenerateSamples(), %s\n", b2_last_error());
            throw 0;
        }
    }
    memcpy(_buffer, fifo.data() + fifo_pos, K * sizeof(std::complex<float>));
    fifo_pos += K;
}
