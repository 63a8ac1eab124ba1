// multichanneltxrx.h -- multichannel transceiver with the public interface of the reference's
// class (include/multichanneltxrx.h:43-157): a multichanneltx and a multichannelrx, each fed by
// its own thread through the (offline) UHD device.  Note the argument order: callbacks BEFORE
// userdata, the opposite of multichannelrx (lib/multichanneltxrx.cc:53-62).
//
// Differences from the reference, on purpose: the transmitter object is shared between the tx
// worker and the caller's thread, so every access goes through tx_mutex (the reference relies on
// usleep() to paper over that race, lib/multichanneltxrx.cc:256-258), and the wait_* calls sleep
// on a condition variable that the tx worker signals at OFDM symbol boundaries instead of polling.
#ifndef __MULTICHANNELTXRX_H__
#define __MULTICHANNELTXRX_H__

#include <complex>
#include <pthread.h>
#include <liquid/liquid.h>
#include <uhd/usrp/multi_usrp.hpp>

#include "multichanneltx.h"
#include "multichannelrx.h"

void * multichanneltxrx_tx_worker(void * _arg);
void * multichanneltxrx_rx_worker(void * _arg);

class multichanneltxrx {
public:
    //  _num_channels   :   number of channels
    //  _M              :   OFDM: number of subcarriers
    //  _cp_len         :   OFDM: cyclic prefix length
    //  _taper_len      :   OFDM: taper prefix length
    //  _p              :   OFDM: subcarrier allocation
    //  _callback       :   frame synchronizer callback function array
    //  _userdata       :   user-defined data structure array
    multichanneltxrx(unsigned int         _num_channels,
                     unsigned int         _M,
                     unsigned int         _cp_len,
                     unsigned int         _taper_len,
                     unsigned char *      _p,
                     framesync_callback * _callback,
                     void **              _userdata);
    ~multichanneltxrx();

    // transmitter methods
    void set_tx_freq(float _tx_freq);
    void set_tx_rate(float _tx_rate);
    void set_tx_gain_soft(float _tx_gain_soft);
    void set_tx_gain_uhd(float _tx_gain_uhd);
    void set_tx_antenna(char * _tx_antenna);
    void reset_tx();
    void start_tx();
    void stop_tx();

    // queue a packet on a channel; returns 0 on success, -1 if the channel is busy
    int transmit_packet(unsigned int    _channel,
                        unsigned char * _header,
                        unsigned char * _payload,
                        unsigned int    _payload_len,
                        int             _mod,
                        int             _fec0,
                        int             _fec1);
    bool is_channel_available(unsigned int _channel);
    unsigned int get_available_channel();       // blocks until some channel is free
    void wait_for_channel(unsigned int _channel);
    void wait_for_tx_to_complete();

    // receiver methods
    void set_rx_freq(float _rx_freq);
    void set_rx_rate(float _rx_rate);
    void set_rx_gain_uhd(float _rx_gain_uhd);
    void set_rx_antenna(char * _rx_antenna);
    void reset_rx();
    void start_rx();
    void stop_rx();

    void debug_enable();
    void debug_disable();

    friend void * multichanneltxrx_tx_worker(void * _arg);
    friend void * multichanneltxrx_rx_worker(void * _arg);

private:
    void set_timespec(struct timespec * _ts, float _timeout);

    unsigned int num_channels;

    multichanneltx mctx;
    float tx_gain;
    pthread_t tx_process;
    pthread_mutex_t tx_mutex;       // guards mctx and the tx flags
    pthread_cond_t  tx_cond;        // start/stop of the transmitter
    pthread_cond_t  tx_ready_cond;  // signalled by the worker when channel readiness may have changed
    bool tx_running;
    bool tx_thread_running;

    multichannelrx mcrx;
    pthread_t rx_process;
    pthread_mutex_t rx_mutex;       // guards mcrx and the rx flags
    pthread_cond_t  rx_cond;
    bool rx_running;
    bool rx_thread_running;
    bool debug_enabled;

    uhd::usrp::multi_usrp::sptr usrp_tx;
    uhd::usrp::multi_usrp::sptr usrp_rx;
    uhd::tx_metadata_t          metadata_tx;
};

#endif // __MULTICHANNELTXRX_H__
