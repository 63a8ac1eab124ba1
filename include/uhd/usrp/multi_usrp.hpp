// uhd/usrp/multi_usrp.hpp -- offline stand-in for the slice of the (pre-streamer) UHD API that
// jgaeddert/liquid-usrp uses: lib/ofdmtxrx.cc:95-108,263-291,335-361,493,507,560,593-597,
// lib/multichanneltxrx.cc (same calls) and src/multichannel_rx.cc:121-220,
// src/multichannel_tx.cc:101-223.  Written from scratch; header-only.
//
// There is no radio: the "device" is a complex-float32 stream.
//   receive : samples come from the file named by $B2_UHD_RX_FILE (interleaved little-endian
//             cf32, the layout of uhd::io_type_t::COMPLEX_FLOAT32 buffers); when the file is
//             exhausted, or when the variable is unset, recv() returns 0 samples with
//             ERROR_CODE_TIMEOUT (the reference's loops treat that as fatal / end of run).
//             $B2_UHD_RX_LOOP=1 rewinds instead.
//   transmit: samples are appended to $B2_UHD_TX_FILE when set, otherwise dropped.  With
//             $B2_UHD_TX_MAX_SAMPLES=n the "radio" is switched off after n samples: the file is
//             closed and the process exits (the reference's transmit programs loop forever).
//             Next to the samples a SigMF-style sidecar "<file>.sigmf-meta" is written when the
//             stream is closed (datatype cf32_le, sample rate, centre frequency, gain, sample
//             count, and the free-text $B2_UHD_NOTE, e.g. "N=4 M=64 cp=16 taper=4"), so a
//             capture can be replayed later; the receive side reads the sidecar of
//             $B2_UHD_RX_FILE when there is one and warns if its rate differs from set_rx_rate.
// All rate/frequency/gain setters just remember their value.
#ifndef B2_UHD_STUB_MULTI_USRP_HPP
#define B2_UHD_STUB_MULTI_USRP_HPP

#include <complex>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <unistd.h>     // the real UHD headers pull these in; the reference's programs rely on it

namespace uhd {

class device_addr_t : public std::map<std::string, std::string> {
public:
    device_addr_t(const std::string & args = "") { (void)args; }
    std::string to_string() const { return "type=b200-offline-stub"; }
};

struct time_spec_t {
    double secs;
    time_spec_t(double s = 0.0) : secs(s) {}
    double get_real_secs() const { return secs; }
};

struct io_type_t {
    enum tid_t { CUSTOM_TYPE = '?', COMPLEX_FLOAT64 = 'd', COMPLEX_FLOAT32 = 'f', COMPLEX_INT16 = 's', COMPLEX_INT8 = 'b' };
    const size_t size;
    const tid_t tid;
    io_type_t(tid_t t) : size(t == COMPLEX_FLOAT32 ? 8 : (t == COMPLEX_FLOAT64 ? 16 : (t == COMPLEX_INT16 ? 4 : 2))), tid(t) {}
};

struct tx_metadata_t {
    bool has_time_spec;
    time_spec_t time_spec;
    bool start_of_burst;
    bool end_of_burst;
    tx_metadata_t() : has_time_spec(false), time_spec(0.0), start_of_burst(false), end_of_burst(false) {}
};

struct rx_metadata_t {
    bool has_time_spec;
    time_spec_t time_spec;
    bool more_fragments;
    size_t fragment_offset;
    bool start_of_burst;
    bool end_of_burst;
    enum error_code_t {
        ERROR_CODE_NONE = 0x0, ERROR_CODE_TIMEOUT = 0x1, ERROR_CODE_LATE_COMMAND = 0x2, ERROR_CODE_BROKEN_CHAIN = 0x4,
        ERROR_CODE_OVERFLOW = 0x8, ERROR_CODE_ALIGNMENT = 0xc, ERROR_CODE_BAD_PACKET = 0xf
    } error_code;
    rx_metadata_t() : has_time_spec(false), time_spec(0.0), more_fragments(false), fragment_offset(0),
                      start_of_burst(false), end_of_burst(false), error_code(ERROR_CODE_NONE) {}
};

struct stream_cmd_t {
    enum stream_mode_t {
        STREAM_MODE_START_CONTINUOUS = 'a', STREAM_MODE_STOP_CONTINUOUS = 'o',
        STREAM_MODE_NUM_SAMPS_AND_DONE = 'd', STREAM_MODE_NUM_SAMPS_AND_MORE = 'm'
    } stream_mode;
    size_t num_samps;
    bool stream_now;
    time_spec_t time_spec;
    stream_cmd_t(const stream_mode_t & mode) : stream_mode(mode), num_samps(0), stream_now(true), time_spec(0.0) {}
};

class device {
public:
    typedef std::shared_ptr<device> sptr;
    enum send_mode_t { SEND_MODE_FULL_BUFF = 0, SEND_MODE_ONE_PACKET = 1 };
    enum recv_mode_t { RECV_MODE_FULL_BUFF = 0, RECV_MODE_ONE_PACKET = 1 };

    device() : rx_(NULL), tx_(NULL), streaming_(false), tx_sent_(0), tx_max_(0),
               tx_rate_(0), tx_freq_(0), tx_gain_(0), capture_rate_(0)
    {
        const char * mx = getenv("B2_UHD_TX_MAX_SAMPLES");
        if (mx && *mx) tx_max_ = strtoull(mx, NULL, 10);
        const char * r = getenv("B2_UHD_RX_FILE");
        const char * t = getenv("B2_UHD_TX_FILE");
        const char * l = getenv("B2_UHD_RX_LOOP");
        if (r && *r) {
            rx_ = fopen(r, "rb");
            // sidecar of the capture: only the sample rate is needed on this side
            FILE * m = fopen((std::string(r) + ".sigmf-meta").c_str(), "r");
            if (m) {
                char line[512];
                while (fgets(line, sizeof(line), m)) {
                    const char * k = strstr(line, "\"core:sample_rate\"");
                    if (k && (k = strchr(k, ':')) && (k = strchr(k + 1, ':'))) capture_rate_ = strtod(k + 1, NULL);
                }
                fclose(m);
            }
        }
        if (t && *t) { tx_ = fopen(t, "ab"); tx_path_ = t; }
        loop_ = l && *l == '1';
    }
    ~device()
    {
        if (rx_) fclose(rx_);
        close_tx();
    }
    // transmit settings recorded in the sidecar; capture_rate() = rate found in the receive sidecar (0: none)
    void note_tx(double rate, double freq, double gain) { tx_rate_ = rate; tx_freq_ = freq; tx_gain_ = gain; }
    double capture_rate() const { return capture_rate_; }
    void close_tx()
    {
        if (!tx_) return;
        long long bytes = ftell(tx_);
        fclose(tx_);
        tx_ = NULL;
        FILE * m = fopen((tx_path_ + ".sigmf-meta").c_str(), "w");
        if (!m) return;
        const char * note = getenv("B2_UHD_NOTE");
        fprintf(m, "{\n  \"global\": {\n    \"core:datatype\": \"cf32_le\",\n    \"core:version\": \"1.0.0\",\n"
                   "    \"core:sample_rate\": %.17g,\n    \"core:description\": \"%s\",\n    \"b2:tx_gain_db\": %.17g,\n"
                   "    \"b2:samples\": %lld\n  },\n  \"captures\": [ { \"core:sample_start\": 0, \"core:frequency\": %.17g } ],\n"
                   "  \"annotations\": []\n}\n",
                tx_rate_, note ? note : "", tx_gain_, bytes >= 0 ? bytes / 8 : 0LL, tx_freq_);
        fclose(m);
    }
    size_t get_max_send_samps_per_packet() const { return 362; }     // what a USRP1/N2x0 reports at MTU 1500
    size_t get_max_recv_samps_per_packet() const { return 362; }

    size_t send(const void * buff, size_t nsamps, const tx_metadata_t & md, const io_type_t & io,
                send_mode_t mode, double timeout = 0.1)
    {
        (void)md; (void)mode; (void)timeout;
        if (tx_ && nsamps) {
            fwrite(buff, io.size, nsamps, tx_);
            fflush(tx_);
        }
        tx_sent_ += nsamps;
        if (tx_max_ && tx_sent_ >= tx_max_) {
            close_tx();
            printf("uhd stub: %llu samples sent, transmitter off\n", tx_sent_);
            exit(0);
        }
        return nsamps;
    }
    size_t recv(void * buff, size_t nsamps, rx_metadata_t & md, const io_type_t & io,
                recv_mode_t mode, double timeout = 0.1)
    {
        (void)mode; (void)timeout;
        md = rx_metadata_t();
        size_t got = 0;
        if (rx_ && streaming_) {
            got = fread(buff, io.size, nsamps, rx_);
            if (got < nsamps && loop_) {
                rewind(rx_);
                got += fread((char *)buff + got * io.size, io.size, nsamps - got, rx_);
            }
        }
        if (got == 0) md.error_code = rx_metadata_t::ERROR_CODE_TIMEOUT;
        return got;
    }
    void set_streaming(bool on) { streaming_ = on; }

private:
    FILE * rx_;
    FILE * tx_;
    bool loop_;
    bool streaming_;
    unsigned long long tx_sent_, tx_max_;
    std::string tx_path_;
    double tx_rate_, tx_freq_, tx_gain_, capture_rate_;
};

namespace usrp {

class multi_usrp {
public:
    typedef std::shared_ptr<multi_usrp> sptr;
    static sptr make(const device_addr_t & dev_addr)
    {
        (void)dev_addr;
        return sptr(new multi_usrp());
    }
    multi_usrp() : dev_(new device()), rx_rate_(1e6), tx_rate_(1e6), rx_freq_(0), tx_freq_(0), rx_gain_(0), tx_gain_(0) {}
    device::sptr get_device() { return dev_; }
    std::string get_pp_string() { return "offline cf32 stream (UHD stub)\n"; }

    void set_rx_rate(double rate, size_t chan = 0)
    {
        (void)chan; rx_rate_ = rate;
        double c = dev_->capture_rate();
        if (c > 0 && (c > rate * 1.000001 || c < rate * 0.999999))
            fprintf(stderr, "uhd stub: capture was taken at %g S/s, receiver set to %g S/s (resample by %g)\n", c, rate, rate / c);
    }
    double get_rx_rate(size_t chan = 0) { (void)chan; return rx_rate_; }
    void set_tx_rate(double rate, size_t chan = 0) { (void)chan; tx_rate_ = rate; dev_->note_tx(tx_rate_, tx_freq_, tx_gain_); }
    double get_tx_rate(size_t chan = 0) { (void)chan; return tx_rate_; }
    void set_rx_freq(double f, size_t chan = 0) { (void)chan; rx_freq_ = f; }
    double get_rx_freq(size_t chan = 0) { (void)chan; return rx_freq_; }
    void set_tx_freq(double f, size_t chan = 0) { (void)chan; tx_freq_ = f; dev_->note_tx(tx_rate_, tx_freq_, tx_gain_); }
    double get_tx_freq(size_t chan = 0) { (void)chan; return tx_freq_; }
    void set_rx_gain(double g, size_t chan = 0) { (void)chan; rx_gain_ = g; }
    void set_tx_gain(double g, size_t chan = 0) { (void)chan; tx_gain_ = g; dev_->note_tx(tx_rate_, tx_freq_, tx_gain_); }
    void set_rx_antenna(const std::string & a, size_t chan = 0) { (void)chan; rx_ant_ = a; }
    void set_tx_antenna(const std::string & a, size_t chan = 0) { (void)chan; tx_ant_ = a; }
    void set_rx_bandwidth(double bw, size_t chan = 0) { (void)bw; (void)chan; }
    void set_tx_bandwidth(double bw, size_t chan = 0) { (void)bw; (void)chan; }
    void issue_stream_cmd(const stream_cmd_t & cmd, size_t chan = 0)
    {
        (void)chan;
        dev_->set_streaming(cmd.stream_mode == stream_cmd_t::STREAM_MODE_START_CONTINUOUS);
    }

private:
    device::sptr dev_;
    double rx_rate_, tx_rate_, rx_freq_, tx_freq_, rx_gain_, tx_gain_;
    std::string rx_ant_, tx_ant_;
};

} // namespace usrp
} // namespace uhd

#endif // B2_UHD_STUB_MULTI_USRP_HPP
