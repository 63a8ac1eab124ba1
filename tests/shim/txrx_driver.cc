// txrx_driver.cc -- test driver of the ofdmtxrx class (liquid-usrp_b200/host/ofdmtxrx.h, the drop-in for the
// reference's include/ofdmtxrx.h) over the offline UHD stand-in.  Test infrastructure, not part of the product.
//
//   txrx_driver split <a.cf32> <b.cf32>
//       the same packet sent twice: a) ofdmtxrx::transmit_packet (lib/ofdmtxrx.cc:297-363), b) the split-phase calls
//       assemble_frame / write_symbol / transmit_symbol / end_transmit_frame (lib/ofdmtxrx.cc:366-449), with the
//       caller scaling fgbuffer between write_symbol and transmit_symbol (what the split API exists for)
//   txrx_driver blocking            ($B2_UHD_RX_FILE names the capture)
//       ofdmtxrx(..., true): ofdmtxrx_rx_worker_blocking (lib/ofdmtxrx.cc:642-739) publishes every received buffer in
//       *rx_buffer; this program's second thread conjugates it back (the capture holds the conjugate of a valid
//       transmission, so nothing decodes unless the edit reaches the synchroniser) and signals rx_buffer_modified_cond
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <unistd.h>
#include <liquid/liquid.h>
#include "ofdmtxrx.h"

static unsigned int n_frames = 0, n_headers = 0, n_packets = 0, n_bytes = 0;
static int callback(unsigned char * header, int header_valid, unsigned char * payload, unsigned int payload_len,
                    int payload_valid, framesyncstats_s stats, void * userdata)
{
    (void)header; (void)stats; (void)userdata;
    n_frames++;
    if (header_valid) n_headers++;
    if (payload_valid) {
        n_packets++; n_bytes += payload_len;
        for (unsigned int i = 0; i < payload_len; i++) if (payload[i] != (unsigned char)(i * 7 + 3)) { n_packets--; break; }
    }
    return 0;
}

struct editor_arg { ofdmtxrx * t; volatile bool run; unsigned int edited; };
static void * editor(void * p)
{
    editor_arg * a = (editor_arg *)p;
    // the mutex is held whenever this thread is not waiting, so the worker can only publish a buffer while
    // the editor is listening (no lost wake-ups)
    pthread_mutex_lock(&a->t->rx_buffer_mutex);
    while (a->run) {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        ts.tv_nsec += 50000000; if (ts.tv_nsec >= 1000000000) { ts.tv_sec++; ts.tv_nsec -= 1000000000; }
        if (pthread_cond_timedwait(&a->t->rx_buffer_filled_cond, &a->t->rx_buffer_mutex, &ts) != 0) continue;
        std::vector<std::complex<float> > & b = *a->t->rx_buffer;
        for (size_t i = 0; i < b.size(); i++) b[i] = std::conj(b[i]);
        a->edited++;
        pthread_cond_signal(&a->t->rx_buffer_modified_cond);
    }
    pthread_mutex_unlock(&a->t->rx_buffer_mutex);
    return NULL;
}

int main(int argc, char ** argv)
{
    const unsigned int M = 64, cp = 16, taper = 4, plen = 200;
    unsigned char header[8] = {1, 2, 3, 4, 5, 6, 7, 8}, payload[plen];
    for (unsigned int i = 0; i < plen; i++) payload[i] = (unsigned char)(i * 7 + 3);
    if (argc >= 4 && !strcmp(argv[1], "split")) {
        setenv("B2_UHD_TX_FILE", argv[2], 1);
        {
            ofdmtxrx a(M, cp, taper, NULL, callback, NULL);
            a.set_tx_gain_soft(-6.0f);
            a.transmit_packet(header, payload, plen, LIQUID_MODEM_QAM16, LIQUID_FEC_NONE, LIQUID_FEC_HAMMING128);
        }
        setenv("B2_UHD_TX_FILE", argv[3], 1);
        unsigned int nsym = 0;
        {
            ofdmtxrx b(M, cp, taper, NULL, callback, NULL);
            b.set_tx_gain_soft(-6.0f);
            b.assemble_frame(header, payload, plen, LIQUID_MODEM_QAM16, LIQUID_FEC_NONE, LIQUID_FEC_HAMMING128);
            bool last = false;
            while (!last) {
                last = b.write_symbol();
                for (unsigned int i = 0; i < b.fgbuffer_len; i++) b.fgbuffer[i] *= 0.5f;      // the caller's edit
                b.transmit_symbol();
                nsym++;
            }
            b.end_transmit_frame();
        }
        printf("split-phase symbols: %u\n", nsym);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "blocking")) {
        ofdmtxrx t(M, cp, taper, NULL, callback, NULL, true);
        editor_arg ea = {&t, true, 0};
        pthread_t th;
        pthread_create(&th, NULL, editor, &ea);
        usleep(20000);
        t.start_rx();
        for (int i = 0; i < 300 && n_packets < 6; i++) usleep(10000);
        t.stop_rx();
        ea.run = false;
        pthread_join(th, NULL);
        printf("frames detected: %u\nvalid headers: %u\nvalid packets: %u\nbytes: %u\nedited buffers: %u\n", n_frames, n_headers, n_packets, n_bytes, ea.edited);
        return 0;
    }
    fprintf(stderr, "usage: txrx_driver split a.cf32 b.cf32 | blocking\n");
    return 2;
}
