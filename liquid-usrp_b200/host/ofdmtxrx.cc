// ofdmtxrx.cc -- reference-compatible single-link transceiver (lib/ofdmtxrx.cc:52-739) over the
// B200-backed liquid handles and the offline UHD stand-in.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/time.h>
#include <unistd.h>

#include "ofdmtxrx.h"

ofdmtxrx::ofdmtxrx(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char * _p,
                   framesync_callback _callback, void * _userdata)
{
    (void)_p;           // the reference always passes NULL to liquid here (lib/ofdmtxrx.cc:78,84,91)
    init(_M, _cp_len, _taper_len, _callback, _userdata, false);
}

ofdmtxrx::ofdmtxrx(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len, unsigned char * _p,
                   framesync_callback _callback, void * _userdata, bool _blocking_rx_worker)
{
    (void)_p;
    init(_M, _cp_len, _taper_len, _callback, _userdata, _blocking_rx_worker);
}

void ofdmtxrx::init(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                    framesync_callback _callback, void * _userdata, bool _blocking)
{
    if (_M < 8) {
        fprintf(stderr, "error: ofdmtxrx::ofdmtxrx(), number of subcarriers must be at least 8\n");
        throw 0;
    } else if (_cp_len < 1) {
        fprintf(stderr, "error: ofdmtxrx::ofdmtxrx(), cyclic prefix length must be at least 1\n");
        throw 0;
    } else if (_taper_len > _cp_len) {
        fprintf(stderr, "error: ofdmtxrx::ofdmtxrx(), taper length cannot exceed cyclic prefix length\n");
        throw 0;
    }
    M = _M; cp_len = _cp_len; taper_len = _taper_len;
    debug_enabled = false;

    // frame generator: default properties CRC-32 / none / h128 / QPSK (lib/ofdmtxrx.cc:79-83)
    ofdmflexframegenprops_init_default(&fgprops);
    fgprops.check = LIQUID_CRC_32;
    fgprops.fec0 = LIQUID_FEC_NONE;
    fgprops.fec1 = LIQUID_FEC_HAMMING128;
    fgprops.mod_scheme = LIQUID_MODEM_QPSK;
    fg = ofdmflexframegen_create(M, cp_len, taper_len, NULL, &fgprops);
    fgbuffer_len = M + cp_len;
    fgbuffer = (std::complex<float> *)malloc(fgbuffer_len * sizeof(std::complex<float>));
    memset((void *)fgbuffer, 0, fgbuffer_len * sizeof(std::complex<float>));

    fs = ofdmflexframesync_create(M, cp_len, taper_len, NULL, _callback, _userdata);

    uhd::device_addr_t dev_addr;
    usrp_tx = uhd::usrp::multi_usrp::make(dev_addr);
    usrp_rx = uhd::usrp::multi_usrp::make(dev_addr);
    // defaults of the reference (lib/ofdmtxrx.cc:99-108)
    set_tx_freq(462.0e6f);
    set_tx_rate(500e3);
    set_tx_gain_soft(-12.0f);
    set_tx_gain_uhd(40.0f);
    set_rx_freq(462.0e6f);
    set_rx_rate(500e3);
    set_rx_gain_uhd(20.0f);

    rx_buffer = new std::vector<std::complex<float> >();
    pthread_mutex_init(&rx_buffer_mutex, NULL);
    pthread_cond_init(&rx_buffer_filled_cond, NULL);
    pthread_cond_init(&rx_buffer_modified_cond, NULL);
    pthread_cond_init(&esbrs_ready, NULL);

    rx_running = false;
    rx_thread_running = true;
    pthread_mutex_init(&rx_mutex, NULL);
    pthread_cond_init(&rx_cond, NULL);
    pthread_create(&rx_process, NULL, _blocking ? ofdmtxrx_rx_worker_blocking : ofdmtxrx_rx_worker, (void *)this);
}

ofdmtxrx::~ofdmtxrx()
{
    // stop the physical receiver, then let the worker leave its wait and exit
    stop_rx();
    pthread_mutex_lock(&rx_mutex);
    rx_thread_running = false;
    pthread_cond_signal(&rx_cond);
    pthread_mutex_unlock(&rx_mutex);
    pthread_mutex_lock(&rx_buffer_mutex);
    pthread_cond_broadcast(&rx_buffer_modified_cond);
    pthread_mutex_unlock(&rx_buffer_mutex);
    void * status;
    pthread_join(rx_process, &status);

    pthread_mutex_destroy(&rx_mutex);
    pthread_cond_destroy(&rx_cond);
    pthread_mutex_destroy(&rx_buffer_mutex);
    pthread_cond_destroy(&rx_buffer_filled_cond);
    pthread_cond_destroy(&rx_buffer_modified_cond);
    pthread_cond_destroy(&esbrs_ready);
    delete rx_buffer;

    ofdmflexframegen_destroy(fg);
    ofdmflexframesync_destroy(fs);
    free(fgbuffer);
}

// ------------------------------------------------------------------ transmitter
void ofdmtxrx::set_tx_freq(float _tx_freq) { usrp_tx->set_tx_freq(_tx_freq); }
void ofdmtxrx::set_tx_rate(float _tx_rate) { usrp_tx->set_tx_rate(_tx_rate); }
void ofdmtxrx::set_tx_gain_soft(float _tx_gain_soft) { tx_gain = powf(10.0f, _tx_gain_soft / 20.0f); }
void ofdmtxrx::set_tx_gain_uhd(float _tx_gain_uhd) { usrp_tx->set_tx_gain(_tx_gain_uhd); }
void ofdmtxrx::set_tx_antenna(char * _tx_antenna) { usrp_tx->set_tx_antenna(_tx_antenna); }
void ofdmtxrx::reset_tx() { ofdmflexframegen_reset(fg); }

// fgbuffer x soft gain -> device
void ofdmtxrx::send_fgbuffer()
{
    std::vector<std::complex<float> > usrp_buffer(fgbuffer_len);
    for (unsigned int i = 0; i < fgbuffer_len; i++) usrp_buffer[i] = fgbuffer[i] * tx_gain;
    usrp_tx->get_device()->send(&usrp_buffer.front(), usrp_buffer.size(), metadata_tx,
                                uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
}

void ofdmtxrx::transmit_packet(unsigned char * _header, unsigned char * _payload, unsigned int _payload_len,
                               int _mod, int _fec0, int _fec1)
{
    metadata_tx.start_of_burst = false;
    metadata_tx.end_of_burst = false;
    metadata_tx.has_time_spec = false;

    assemble_frame(_header, _payload, _payload_len, _mod, _fec0, _fec1);
    bool last_symbol = false;
    while (!last_symbol) {
        last_symbol = ofdmflexframegen_write(fg, fgbuffer, fgbuffer_len);
        send_fgbuffer();
    }
    // the reference repeats the last buffer once more, then closes the burst (lib/ofdmtxrx.cc:344-361)
    send_fgbuffer();
    metadata_tx.start_of_burst = false;
    metadata_tx.end_of_burst = true;
    usrp_tx->get_device()->send("", 0, metadata_tx, uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
}

void ofdmtxrx::assemble_frame(unsigned char * _header, unsigned char * _payload, unsigned int _payload_len,
                              int _mod, int _fec0, int _fec1)
{
    fgprops.mod_scheme = _mod;
    fgprops.fec0 = _fec0;
    fgprops.fec1 = _fec1;
    ofdmflexframegen_setprops(fg, &fgprops);
    ofdmflexframegen_assemble(fg, _header, _payload, _payload_len);
}

bool ofdmtxrx::write_symbol() { return ofdmflexframegen_writesymbol(fg, fgbuffer); }
void ofdmtxrx::transmit_symbol() { send_fgbuffer(); }

void ofdmtxrx::end_transmit_frame()
{
    send_fgbuffer();
    metadata_tx.start_of_burst = false;
    metadata_tx.end_of_burst = true;
    usrp_tx->get_device()->send("", 0, metadata_tx, uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
}

// ------------------------------------------------------------------ receiver
void ofdmtxrx::set_rx_freq(float _rx_freq) { usrp_rx->set_rx_freq(_rx_freq); }
void ofdmtxrx::set_rx_rate(float _rx_rate) { usrp_rx->set_rx_rate(_rx_rate); }
void ofdmtxrx::set_rx_gain_uhd(float _rx_gain_uhd) { usrp_rx->set_rx_gain(_rx_gain_uhd); }
void ofdmtxrx::set_rx_antenna(char * _rx_antenna) { usrp_rx->set_rx_antenna(_rx_antenna); }
void ofdmtxrx::reset_rx() { ofdmflexframesync_reset(fs); }

void ofdmtxrx::start_rx()
{
    pthread_mutex_lock(&rx_mutex);
    rx_running = true;
    usrp_rx->issue_stream_cmd(uhd::stream_cmd_t::STREAM_MODE_START_CONTINUOUS);
    pthread_cond_signal(&rx_cond);
    pthread_mutex_unlock(&rx_mutex);
}

void ofdmtxrx::stop_rx()
{
    pthread_mutex_lock(&rx_mutex);
    rx_running = false;
    usrp_rx->issue_stream_cmd(uhd::stream_cmd_t::STREAM_MODE_STOP_CONTINUOUS);
    pthread_mutex_unlock(&rx_mutex);
}

void ofdmtxrx::debug_enable() { debug_enabled = true; ofdmflexframesync_debug_enable(fs); }
void ofdmtxrx::debug_disable() { debug_enabled = false; ofdmflexframesync_debug_disable(fs); }

void ofdmtxrx::set_timespec(struct timespec * _ts, float _timeout)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    long us = (long)(_timeout * 1e6f) + tp.tv_usec;
    _ts->tv_sec = tp.tv_sec + us / 1000000;
    _ts->tv_nsec = (us % 1000000) * 1000;
}

// ------------------------------------------------------------------ workers
// wait until start_rx() (returns false when the object is being destroyed)
static bool wait_for_start(pthread_mutex_t * m, pthread_cond_t * c, bool * running, bool * alive)
{
    pthread_mutex_lock(m);
    while (*alive && !*running) pthread_cond_wait(c, m);
    bool go = *alive;
    pthread_mutex_unlock(m);
    return go;
}

void * ofdmtxrx_rx_worker(void * _arg)
{
    ofdmtxrx * txcvr = (ofdmtxrx *)_arg;
    const size_t max_samps = txcvr->usrp_rx->get_device()->get_max_recv_samps_per_packet();
    std::vector<std::complex<float> > buffer(max_samps);
    uhd::rx_metadata_t md;
    while (wait_for_start(&txcvr->rx_mutex, &txcvr->rx_cond, &txcvr->rx_running, &txcvr->rx_thread_running)) {
        while (txcvr->rx_running && txcvr->rx_thread_running) {
            size_t n = txcvr->usrp_rx->get_device()->recv(&buffer.front(), buffer.size(), md,
                                                          uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::RECV_MODE_ONE_PACKET);
            if (n) {
                ofdmflexframesync_execute(txcvr->fs, &buffer.front(), (unsigned int)n);
            } else {
                // nothing from the (offline) device: deliver what is pending, then idle
                ofdmflexframesync_flush(txcvr->fs);
                usleep(1000);
            }
        }
        ofdmflexframesync_flush(txcvr->fs);
    }
    pthread_exit(NULL);
}

void * ofdmtxrx_rx_worker_blocking(void * _arg)
{
    ofdmtxrx * txcvr = (ofdmtxrx *)_arg;
    const size_t max_samps = txcvr->usrp_rx->get_device()->get_max_recv_samps_per_packet();
    uhd::rx_metadata_t md;
    while (wait_for_start(&txcvr->rx_mutex, &txcvr->rx_cond, &txcvr->rx_running, &txcvr->rx_thread_running)) {
        while (txcvr->rx_running && txcvr->rx_thread_running) {
            pthread_mutex_lock(&txcvr->rx_buffer_mutex);
            txcvr->rx_buffer->resize(max_samps);
            size_t n = txcvr->usrp_rx->get_device()->recv(&txcvr->rx_buffer->front(), max_samps, md,
                                                          uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::RECV_MODE_ONE_PACKET);
            txcvr->rx_buffer->resize(n);
            if (n) {
                // publish the buffer, give the other thread up to 0.1 s to edit it
                pthread_cond_signal(&txcvr->rx_buffer_filled_cond);
                struct timespec ts;
                txcvr->set_timespec(&ts, 0.1f);
                pthread_cond_timedwait(&txcvr->rx_buffer_modified_cond, &txcvr->rx_buffer_mutex, &ts);
                if (!txcvr->rx_buffer->empty())
                    ofdmflexframesync_execute(txcvr->fs, &txcvr->rx_buffer->front(), (unsigned int)txcvr->rx_buffer->size());
                pthread_mutex_unlock(&txcvr->rx_buffer_mutex);
            } else {
                pthread_mutex_unlock(&txcvr->rx_buffer_mutex);
                ofdmflexframesync_flush(txcvr->fs);
                usleep(1000);
            }
        }
        ofdmflexframesync_flush(txcvr->fs);
    }
    pthread_exit(NULL);
}
