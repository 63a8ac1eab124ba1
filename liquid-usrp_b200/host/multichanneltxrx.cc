// multichanneltxrx.cc -- reference-compatible multichannel transceiver
// (lib/multichanneltxrx.cc:53-624) over the B200-backed multichanneltx / multichannelrx classes
// and the offline UHD stand-in.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sys/time.h>
#include <unistd.h>
#include <vector>

#include "multichanneltxrx.h"

multichanneltxrx::multichanneltxrx(unsigned int _num_channels, unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                                   unsigned char * _p, framesync_callback * _callback, void ** _userdata)
    : num_channels(_num_channels),
      mctx(_num_channels, _M, _cp_len, _taper_len, _p),
      mcrx(_num_channels, _M, _cp_len, _taper_len, _p, _userdata, _callback)       // argument order swaps here
{
    debug_enabled = false;
    uhd::device_addr_t dev_addr;
    usrp_tx = uhd::usrp::multi_usrp::make(dev_addr);
    usrp_rx = uhd::usrp::multi_usrp::make(dev_addr);
    set_tx_freq(462.0e6f);
    set_tx_rate(500e3);
    set_tx_gain_soft(-12.0f);
    set_tx_gain_uhd(40.0f);
    set_rx_freq(462.0e6f);
    set_rx_rate(500e3);
    set_rx_gain_uhd(20.0f);

    tx_running = false; tx_thread_running = true;
    pthread_mutex_init(&tx_mutex, NULL);
    pthread_cond_init(&tx_cond, NULL);
    pthread_cond_init(&tx_ready_cond, NULL);
    pthread_create(&tx_process, NULL, multichanneltxrx_tx_worker, (void *)this);

    rx_running = false; rx_thread_running = true;
    // recursive: the receive worker delivers the frame callbacks with the lock held, and a callback may well call
    // stop_rx() / reset_rx() (the reference takes no lock there at all, lib/multichanneltxrx.cc:613)
    {
        pthread_mutexattr_t attr;
        pthread_mutexattr_init(&attr);
        pthread_mutexattr_settype(&attr, PTHREAD_MUTEX_RECURSIVE);
        pthread_mutex_init(&rx_mutex, &attr);
        pthread_mutexattr_destroy(&attr);
    }
    pthread_cond_init(&rx_cond, NULL);
    pthread_create(&rx_process, NULL, multichanneltxrx_rx_worker, (void *)this);
}

multichanneltxrx::~multichanneltxrx()
{
    stop_rx();
    stop_tx();
    pthread_mutex_lock(&tx_mutex);
    tx_thread_running = false;
    pthread_cond_broadcast(&tx_cond);
    pthread_cond_broadcast(&tx_ready_cond);
    pthread_mutex_unlock(&tx_mutex);
    pthread_mutex_lock(&rx_mutex);
    rx_thread_running = false;
    pthread_cond_broadcast(&rx_cond);
    pthread_mutex_unlock(&rx_mutex);
    void * status;
    pthread_join(tx_process, &status);
    pthread_join(rx_process, &status);
    pthread_mutex_destroy(&tx_mutex);
    pthread_cond_destroy(&tx_cond);
    pthread_cond_destroy(&tx_ready_cond);
    pthread_mutex_destroy(&rx_mutex);
    pthread_cond_destroy(&rx_cond);
}

// ------------------------------------------------------------------ transmitter
void multichanneltxrx::set_tx_freq(float _tx_freq) { usrp_tx->set_tx_freq(_tx_freq); }
void multichanneltxrx::set_tx_rate(float _tx_rate) { usrp_tx->set_tx_rate(_tx_rate); }
void multichanneltxrx::set_tx_gain_soft(float _tx_gain_soft) { tx_gain = powf(10.0f, _tx_gain_soft / 20.0f); }
void multichanneltxrx::set_tx_gain_uhd(float _tx_gain_uhd) { usrp_tx->set_tx_gain(_tx_gain_uhd); }
void multichanneltxrx::set_tx_antenna(char * _tx_antenna) { usrp_tx->set_tx_antenna(_tx_antenna); }

void multichanneltxrx::reset_tx()
{
    pthread_mutex_lock(&tx_mutex);
    mctx.Reset();
    pthread_cond_broadcast(&tx_ready_cond);
    pthread_mutex_unlock(&tx_mutex);
}
void multichanneltxrx::start_tx()
{
    pthread_mutex_lock(&tx_mutex);
    tx_running = true;
    pthread_cond_broadcast(&tx_cond);
    pthread_mutex_unlock(&tx_mutex);
}
void multichanneltxrx::stop_tx()
{
    pthread_mutex_lock(&tx_mutex);
    tx_running = false;
    pthread_cond_broadcast(&tx_ready_cond);
    pthread_mutex_unlock(&tx_mutex);
}

int multichanneltxrx::transmit_packet(unsigned int _channel, unsigned char * _header, unsigned char * _payload,
                                      unsigned int _payload_len, int _mod, int _fec0, int _fec1)
{
    pthread_mutex_lock(&tx_mutex);
    if (!tx_running) {
        pthread_mutex_unlock(&tx_mutex);
        fprintf(stderr, "error: multichanneltxrx:transmit_packet(), transmitter not yet running\n");
        throw 0;
    } else if (_channel >= num_channels) {
        pthread_mutex_unlock(&tx_mutex);
        fprintf(stderr, "error: multichanneltxrx:transmit_packet(), invalid channel %u\n", _channel);
        throw 0;
    } else if (!mctx.IsChannelReadyForData(_channel)) {
        pthread_mutex_unlock(&tx_mutex);
        fprintf(stderr, "warning: multichanneltxrx:transmit_packet(), channel %u not ready for data\n", _channel);
        return -1;
    }
    try {
        mctx.UpdateData(_channel, _header, _payload, _payload_len, _mod, _fec0, _fec1);
    } catch (...) {
        pthread_mutex_unlock(&tx_mutex);
        throw;
    }
    pthread_mutex_unlock(&tx_mutex);
    return 0;
}

bool multichanneltxrx::is_channel_available(unsigned int _channel)
{
    pthread_mutex_lock(&tx_mutex);
    bool ready = false;
    try { ready = mctx.IsChannelReadyForData(_channel) != 0; } catch (...) { pthread_mutex_unlock(&tx_mutex); throw; }
    pthread_mutex_unlock(&tx_mutex);
    return ready;
}

unsigned int multichanneltxrx::get_available_channel()
{
    pthread_mutex_lock(&tx_mutex);
    while (true) {
        for (unsigned int i = 0; i < num_channels; i++) {
            if (mctx.IsChannelReadyForData(i)) {
                pthread_mutex_unlock(&tx_mutex);
                return i;
            }
        }
        struct timespec ts;
        set_timespec(&ts, 0.05f);
        pthread_cond_timedwait(&tx_ready_cond, &tx_mutex, &ts);
    }
}

void multichanneltxrx::wait_for_channel(unsigned int _channel)
{
    pthread_mutex_lock(&tx_mutex);
    while (!mctx.IsChannelReadyForData(_channel)) {
        struct timespec ts;
        set_timespec(&ts, 0.05f);
        pthread_cond_timedwait(&tx_ready_cond, &tx_mutex, &ts);
    }
    pthread_mutex_unlock(&tx_mutex);
}

void multichanneltxrx::wait_for_tx_to_complete()
{
    pthread_mutex_lock(&tx_mutex);
    while (true) {
        bool all_available = true;
        for (unsigned int i = 0; i < num_channels; i++)
            if (!mctx.IsChannelReadyForData(i)) all_available = false;
        if (all_available) break;
        struct timespec ts;
        set_timespec(&ts, 0.05f);
        pthread_cond_timedwait(&tx_ready_cond, &tx_mutex, &ts);
    }
    pthread_mutex_unlock(&tx_mutex);
}

// ------------------------------------------------------------------ receiver
void multichanneltxrx::set_rx_freq(float _rx_freq) { usrp_rx->set_rx_freq(_rx_freq); }
void multichanneltxrx::set_rx_rate(float _rx_rate) { usrp_rx->set_rx_rate(_rx_rate); }
void multichanneltxrx::set_rx_gain_uhd(float _rx_gain_uhd) { usrp_rx->set_rx_gain(_rx_gain_uhd); }
void multichanneltxrx::set_rx_antenna(char * _rx_antenna) { usrp_rx->set_rx_antenna(_rx_antenna); }

void multichanneltxrx::reset_rx()
{
    pthread_mutex_lock(&rx_mutex);
    mcrx.Reset();
    pthread_mutex_unlock(&rx_mutex);
}
void multichanneltxrx::start_rx()
{
    pthread_mutex_lock(&rx_mutex);
    rx_running = true;
    usrp_rx->issue_stream_cmd(uhd::stream_cmd_t::STREAM_MODE_START_CONTINUOUS);
    pthread_cond_broadcast(&rx_cond);
    pthread_mutex_unlock(&rx_mutex);
}
void multichanneltxrx::stop_rx()
{
    pthread_mutex_lock(&rx_mutex);
    rx_running = false;
    usrp_rx->issue_stream_cmd(uhd::stream_cmd_t::STREAM_MODE_STOP_CONTINUOUS);
    pthread_mutex_unlock(&rx_mutex);
}

void multichanneltxrx::debug_enable() { debug_enabled = true; }
void multichanneltxrx::debug_disable() { debug_enabled = false; }

void multichanneltxrx::set_timespec(struct timespec * _ts, float _timeout)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    long us = (long)(_timeout * 1e6f) + tp.tv_usec;
    _ts->tv_sec = tp.tv_sec + us / 1000000;
    _ts->tv_nsec = (us % 1000000) * 1000;
}

// ------------------------------------------------------------------ workers
void * multichanneltxrx_tx_worker(void * _arg)
{
    multichanneltxrx * t = (multichanneltxrx *)_arg;
    const unsigned int K = 2 * t->num_channels;
    std::vector<std::complex<float> > tx_buffer(K);
    std::vector<std::complex<float> > usrp_buffer(256);
    unsigned int count = 0;
    uhd::tx_metadata_t md;
    while (true) {
        pthread_mutex_lock(&t->tx_mutex);
        while (t->tx_thread_running && !t->tx_running) pthread_cond_wait(&t->tx_cond, &t->tx_mutex);
        if (!t->tx_thread_running) { pthread_mutex_unlock(&t->tx_mutex); break; }
        md.start_of_burst = false; md.end_of_burst = false; md.has_time_spec = false;
        t->mctx.Reset();                                   // as the reference does on every start (lib/multichanneltxrx.cc:451-454)
        pthread_mutex_unlock(&t->tx_mutex);
        while (true) {
            pthread_mutex_lock(&t->tx_mutex);
            bool run = t->tx_running && t->tx_thread_running;
            if (run) {
                t->mctx.GenerateSamples(&tx_buffer[0]);
                pthread_cond_broadcast(&t->tx_ready_cond);
            }
            pthread_mutex_unlock(&t->tx_mutex);
            if (!run) break;
            for (unsigned int i = 0; i < K; i++) {
                usrp_buffer[count++] = tx_buffer[i] * t->tx_gain;
                if (count == 256) {
                    count = 0;
                    t->usrp_tx->get_device()->send(&usrp_buffer.front(), usrp_buffer.size(), md,
                                                   uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
                }
            }
        }
        t->usrp_tx->get_device()->send(&usrp_buffer.front(), usrp_buffer.size(), md,
                                       uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
        md.start_of_burst = false; md.end_of_burst = true;
        t->usrp_tx->get_device()->send("", 0, md, uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::SEND_MODE_FULL_BUFF);
    }
    pthread_exit(NULL);
}

void * multichanneltxrx_rx_worker(void * _arg)
{
    multichanneltxrx * t = (multichanneltxrx *)_arg;
    const size_t max_samps = t->usrp_rx->get_device()->get_max_recv_samps_per_packet();
    std::vector<std::complex<float> > buffer(max_samps);
    uhd::rx_metadata_t md;
    while (true) {
        pthread_mutex_lock(&t->rx_mutex);
        while (t->rx_thread_running && !t->rx_running) pthread_cond_wait(&t->rx_cond, &t->rx_mutex);
        bool alive = t->rx_thread_running;
        pthread_mutex_unlock(&t->rx_mutex);
        if (!alive) break;
        while (true) {
            size_t n = t->usrp_rx->get_device()->recv(&buffer.front(), buffer.size(), md,
                                                      uhd::io_type_t::COMPLEX_FLOAT32, uhd::device::RECV_MODE_ONE_PACKET);
            pthread_mutex_lock(&t->rx_mutex);
            bool run = t->rx_running && t->rx_thread_running;
            if (run) {
                if (n) t->mcrx.Execute(&buffer.front(), (unsigned int)n);
                else t->mcrx.Flush();
            }
            pthread_mutex_unlock(&t->rx_mutex);
            if (!run) break;
            if (!n) usleep(1000);
        }
        pthread_mutex_lock(&t->rx_mutex);
        t->mcrx.Flush();
        pthread_mutex_unlock(&t->rx_mutex);
    }
    pthread_exit(NULL);
}
