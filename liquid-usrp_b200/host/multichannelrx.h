// multichannelrx.h -- N-channel OFDM receiver with the public interface of the reference's
// class (include/multichannelrx.h:29-83): same constructor arguments and order, Reset(),
// GetNumChannels(), Execute().  Underneath, samples are staged and handed to the B200 library
// (b2_mcrx_*, include/b200_ofdm.h) in batches; user callbacks are replayed on the calling thread
// in the reference's order (completion block, then channel).
//
// One behavioural difference, by design: callbacks fire when a batch is flushed (every
// `batch` wideband samples, on Flush(), Reset() and destruction), not on the very sample that
// completes a frame.  Results and their order are identical.
#ifndef __MULTICHANNELRX_H__
#define __MULTICHANNELRX_H__

#include <complex>
#include <vector>
#include <liquid/liquid.h>

struct b2_mcrx_s;

class multichannelrx {
public:
    // default constructor
    //  _num_channels   :   number of channels
    //  _M              :   OFDM: number of subcarriers
    //  _cp_len         :   OFDM: cyclic prefix length
    //  _taper_len      :   OFDM: taper prefix length
    //  _p              :   OFDM: subcarrier allocation
    //  _userdata       :   user-defined data structure array
    //  _callback       :   user-defined callback function array
    multichannelrx(unsigned int         _num_channels,
                   unsigned int         _M,
                   unsigned int         _cp_len,
                   unsigned int         _taper_len,
                   unsigned char *      _p,
                   void **              _userdata,
                   framesync_callback * _callback);
    ~multichannelrx();

    // reset multi-channel receiver (frame synchronizers and channelizer; the NCO keeps running)
    void Reset();

    unsigned int GetNumChannels() { return num_channels; }

    // push samples into the receiver
    void Execute(std::complex<float> * _x, unsigned int _num_samples);

    // extension: process everything pushed so far and deliver the pending callbacks
    void Flush();
    // extension: wideband samples accumulated before an automatic flush (default 2^20,
    // or $B2_MCRX_BATCH); Execute also flushes at the end of a call of >= $B2_MCRX_FLUSH_MIN samples and when
    // the oldest staged sample is older than $B2_MCRX_MAX_LATENCY_MS
    void SetBatchSize(unsigned int _num_samples);

private:
    multichannelrx(const multichannelrx &);
    multichannelrx & operator=(const multichannelrx &);
    void Deliver();

    unsigned int num_channels;
    unsigned int M, cp_len, taper_len;
    b2_mcrx_s * rx;
    std::vector<void *> userdata;
    std::vector<framesync_callback> callback;
    std::complex<float> * stage;         // pinned host staging buffer
    unsigned int stage_len, stage_cap;
    // callbacks fire inside Execute in the reference; here samples are staged, so a call flushes when it is long
    // enough to be worth a launch by itself, or when the oldest staged sample has waited max_wait_ns
    unsigned int flush_min;              // $B2_MCRX_FLUSH_MIN, default 16384 samples
    long long max_wait_ns, first_ns;     // $B2_MCRX_MAX_LATENCY_MS, default 5 ms (0: only the batch size counts)
};

#endif // __MULTICHANNELRX_H__
