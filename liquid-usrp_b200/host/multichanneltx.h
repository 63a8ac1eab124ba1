// multichanneltx.h -- N-channel OFDM transmitter with the public interface of the reference's
// class (include/multichanneltx.h:29-93).  GenerateSamples() still returns exactly 2N samples
// per call; underneath, the B200 library (b2_mctx_*) produces everything up to the next OFDM
// symbol boundary in one go -- channel readiness can only change at those boundaries
// (lib/multichanneltx.cc:198-201,230-242), so the call sequence seen by the user is unchanged.
#ifndef __MULTICHANNELTX_H__
#define __MULTICHANNELTX_H__

#include <complex>
#include <vector>
#include <liquid/liquid.h>

struct b2_mctx_s;

class multichanneltx {
public:
    // default constructor
    //  _num_channels   :   number of channels
    //  _M              :   OFDM: number of subcarriers
    //  _cp_len         :   OFDM: cyclic prefix length
    //  _taper_len      :   OFDM: taper prefix length
    //  _p              :   OFDM: subcarrier allocation
    multichanneltx(unsigned int    _num_channels,
                   unsigned int    _M,
                   unsigned int    _cp_len,
                   unsigned int    _taper_len,
                   unsigned char * _p);
    ~multichanneltx();

    // reset transmitter (frame generators and channelizer; the NCO keeps running)
    void Reset();

    unsigned int GetNumChannels() { return num_channels; }

    // is channel ready for more data?
    int IsChannelReadyForData(unsigned int _channel);

    // update payload data on a particular channel
    void UpdateData(unsigned int    _channel,
                    unsigned char * _header,
                    unsigned char * _payload,
                    unsigned int    _payload_len,
                    int             _mod,
                    int             _fec0,
                    int             _fec1);

    // generate 2*num_channels samples for transmission
    void GenerateSamples(std::complex<float> * _buffer);

private:
    multichanneltx(const multichanneltx &);
    multichanneltx & operator=(const multichanneltx &);

    unsigned int num_channels;
    unsigned int M, cp_len, taper_len;
    b2_mctx_s * tx;
    std::vector<std::complex<float> > fifo;  // samples generated up to the next symbol boundary
    size_t fifo_pos;
};

#endif // __MULTICHANNELTX_H__
