// ofdmtxrx.h -- single-link OFDM transceiver with the public interface of the reference's class
// (include/ofdmtxrx.h:35-174): both constructors, the tx/rx setters, transmit_packet and the
// split assemble_frame / write_symbol / transmit_symbol / end_transmit_frame sequence, rx
// start/stop/reset, debug switches, the two worker entry points and the public data members.
//
// The frame generator and synchroniser handles are liquid-compatible objects backed by the B200
// library (host/liquid_b200.cc); radio I/O goes through the offline UHD stand-in
// (include/uhd/usrp/multi_usrp.hpp).  The rx worker hands whole received buffers to the
// synchroniser instead of one sample per call (lib/ofdmtxrx.cc:620-626).
#ifndef __OFDMTXRX_H__
#define __OFDMTXRX_H__

#include <complex>
#include <vector>
#include <pthread.h>
#include <liquid/liquid.h>
#include <uhd/usrp/multi_usrp.hpp>

// receiver worker thread
void * ofdmtxrx_rx_worker(void * _arg);

// receiver worker thread that publishes each received buffer in *rx_buffer, signals
// rx_buffer_filled_cond and waits for rx_buffer_modified_cond before synchronising, so that
// another thread may edit the samples first
void * ofdmtxrx_rx_worker_blocking(void * _arg);

class ofdmtxrx {
public:
    //  _M              :   OFDM: number of subcarriers
    //  _cp_len         :   OFDM: cyclic prefix length
    //  _taper_len      :   OFDM: taper prefix length
    //  _p              :   OFDM: subcarrier allocation (ignored, as in the reference: default allocation)
    //  _callback       :   frame synchronizer callback function
    //  _userdata       :   user-defined data structure
    ofdmtxrx(unsigned int       _M,
             unsigned int       _cp_len,
             unsigned int       _taper_len,
             unsigned char *    _p,
             framesync_callback _callback,
             void *             _userdata);

    // selects between ofdmtxrx_rx_worker() and ofdmtxrx_rx_worker_blocking()
    ofdmtxrx(unsigned int       _M,
             unsigned int       _cp_len,
             unsigned int       _taper_len,
             unsigned char *    _p,
             framesync_callback _callback,
             void *             _userdata,
             bool               _blocking_rx_worker);

    ~ofdmtxrx();

    // transmitter methods
    void set_tx_freq(float _tx_freq);
    void set_tx_rate(float _tx_rate);
    void set_tx_gain_soft(float _tx_gain_soft);
    void set_tx_gain_uhd(float _tx_gain_uhd);
    void set_tx_antenna(char * _tx_antenna);
    void reset_tx();

    void transmit_packet(unsigned char * _header,
                         unsigned char * _payload,
                         unsigned int    _payload_len,
                         int             _mod,
                         int             _fec0,
                         int             _fec1);

    // transmit_packet() in steps, so the baseband samples in fgbuffer can be edited before
    // they are sent
    void transmit_symbol();
    void assemble_frame(unsigned char * _header,
                        unsigned char * _payload,
                        unsigned int    _payload_len,
                        int             _mod,
                        int             _fec0,
                        int             _fec1);
    bool write_symbol();
    void end_transmit_frame();

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

    friend void * ofdmtxrx_rx_worker(void * _arg);
    friend void * ofdmtxrx_rx_worker_blocking(void * _arg);

    // transmitter objects
    ofdmflexframegen fg;            // frame generator object
    unsigned int fgbuffer_len;      // length of frame generator buffer
    std::complex<float> * fgbuffer; // frame generator output buffer [size: M + cp_len x 1]

    // receiver objects
    std::vector<std::complex<float> > * rx_buffer;
    pthread_mutex_t rx_buffer_mutex;
    pthread_cond_t  rx_buffer_filled_cond;
    pthread_cond_t  rx_buffer_modified_cond;
    pthread_cond_t  esbrs_ready;

private:
    void init(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
              framesync_callback _callback, void * _userdata, bool _blocking);
    void send_fgbuffer();
    void set_timespec(struct timespec * _ts, float _timeout);

    unsigned int M, cp_len, taper_len;
    ofdmflexframegenprops_s fgprops;
    float tx_gain;

    ofdmflexframesync fs;
    pthread_t rx_process;
    pthread_mutex_t rx_mutex;
    pthread_cond_t  rx_cond;
    bool rx_running;
    bool rx_thread_running;
    bool debug_enabled;

    uhd::usrp::multi_usrp::sptr usrp_tx;
    uhd::usrp::multi_usrp::sptr usrp_rx;
    uhd::tx_metadata_t          metadata_tx;
};

#endif // __OFDMTXRX_H__
