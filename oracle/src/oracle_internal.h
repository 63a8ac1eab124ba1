/*
 * oracle_internal.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Scalar C restatement of the liquid-dsp 1.3.x algorithms that jgaeddert/liquid-usrp's
 * multichannel OFDM path calls.  liquid-dsp is an
 * un-vendored, un-pinned dependency of the reference (configure.ac:56, README.md:21) and is
 * not present in this image, so this is written from the published algorithms, anchored on
 * the reference's call sites (lib/multichannelrx.cc, lib/multichanneltx.cc, lib/ofdmtxrx.cc).
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for this path
 *  and the real library cannot run here.  What pins the oracle is
 * listed in oracle/README.md (external known answers + properties + reference L2 code
 * compiled unmodified on top of it).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call anything in this directory.  The product never does.
 */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H

#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "liquid/liquid.h"

typedef float complex cf32;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* -------------------------------------------------------------- math / fft */
typedef struct orc_fft_s {
    unsigned int n;
    int dir;                 /* -1 forward (e^{-j}), +1 backward (e^{+j}); both unnormalised */
    int pow2;
    unsigned int log2n;
    cf32 * tw;               /* n twiddles  e^{dir*j*2*pi*k/n} (double -> float)             */
    unsigned int * brev;     /* bit-reversal table (pow2 only)                                */
} orc_fft;
#define ORC_FFT_FORWARD  (-1)
#define ORC_FFT_BACKWARD (+1)
orc_fft * orc_fft_create(unsigned int n, int dir);
void      orc_fft_destroy(orc_fft * q);
void      orc_fft_execute(const orc_fft * q, const cf32 * x, cf32 * y);   /* x != y */

float orc_besseli0f(float z);
float orc_kaiser(unsigned int i, unsigned int n, float beta);
float orc_sincf(float x);
float orc_kaiser_beta_As(float As);
void  orc_firdes_kaiser(unsigned int n, float fc, float As, float mu, float * h);
/* least-squares polynomial fit evaluated in double (see oracle/README.md, deviation D3) */
void  orc_polyfit_d(const float * x, const float * y, unsigned int n, double * p, unsigned int k);
double orc_polyval_d(const double * p, unsigned int k, double x);

/* ---------------------------------------------------------------- msequence */
typedef struct { unsigned int m, g, a, n, v, b; } orc_mseq;
void         orc_mseq_init_default(orc_mseq * ms, unsigned int m);
void         orc_mseq_reset(orc_mseq * ms);
unsigned int orc_mseq_advance(orc_mseq * ms);
unsigned int orc_mseq_symbol(orc_mseq * ms, unsigned int bps);

/* ------------------------------------------------------------------- window */
typedef struct {
    unsigned int n;      /* window length                     */
    unsigned int cap;    /* allocated samples (2n)            */
    unsigned int pos;    /* index of oldest sample in buf     */
    cf32 * buf;
} orc_window;
void orc_window_init(orc_window * w, unsigned int n);
void orc_window_free(orc_window * w);
void orc_window_clear(orc_window * w);
static inline void orc_window_push(orc_window * w, cf32 x) {
    if (w->pos + w->n == w->cap) {
        memmove(w->buf, w->buf + w->pos + 1, (w->n - 1) * sizeof(cf32));
        w->pos = 0;
        w->buf[w->n - 1] = x;
        return;
    }
    w->buf[w->pos + w->n] = x;
    w->pos++;
}
static inline const cf32 * orc_window_read(const orc_window * w) { return w->buf + w->pos; }

/* -------------------------------------------------------------------- modem */
typedef struct {
    unsigned int scheme, bps, M;
    unsigned int m_i, m_q;
    float alpha;
    float ref[8];
    cf32 r, x_hat;
} orc_modem;
int          orc_modem_supported(unsigned int scheme);
unsigned int orc_modem_bps(unsigned int scheme);
void         orc_modem_init(orc_modem * q, unsigned int scheme);
cf32         orc_modem_modulate(orc_modem * q, unsigned int sym);
unsigned int orc_modem_demodulate(orc_modem * q, cf32 x);
float        orc_modem_evm2(const orc_modem * q);         /* |x - x_hat|^2 */

/* ---------------------------------------------------------------- fec / crc */
uint32_t     orc_crc32(const unsigned char * msg, unsigned int n);
unsigned int orc_fec_enc_len(unsigned int scheme, unsigned int dec_len);
int          orc_fec_supported(unsigned int scheme);
void         orc_fec_encode(unsigned int scheme, unsigned int dec_len, const unsigned char * dec, unsigned char * enc);
void         orc_fec_decode(unsigned int scheme, unsigned int dec_len, const unsigned char * enc, unsigned char * dec);
unsigned int orc_hamming128_encode_symbol(unsigned int s);
unsigned int orc_hamming128_decode_symbol(unsigned int r);
unsigned int orc_golay2412_encode_symbol(unsigned int s);
unsigned int orc_golay2412_decode_symbol(unsigned int r);
void         orc_interleave(unsigned char * x, unsigned int n, unsigned int depth, int decode);
void         orc_scramble(unsigned char * x, unsigned int n);
void         orc_repack_bytes(const unsigned char * in, unsigned int bps_in, unsigned int n_in,
                              unsigned char * out, unsigned int bps_out, unsigned int n_out,
                              unsigned int * n_written);
void         orc_pack_array(unsigned char * dst, unsigned int n, unsigned int k, unsigned int b, unsigned char sym);

typedef struct {
    unsigned int msg_len, packet_len, check, crc_len;
    unsigned int fs[2], dec_len[2], enc_len[2], depth[2];
    unsigned char * buf0, * buf1;
} orc_packetizer;
unsigned int orc_packetizer_enc_len(unsigned int n, unsigned int check, unsigned int fec0, unsigned int fec1);
void orc_packetizer_init(orc_packetizer * p, unsigned int n, unsigned int check, unsigned int fec0, unsigned int fec1);
void orc_packetizer_free(orc_packetizer * p);
void orc_packetizer_encode(orc_packetizer * p, const unsigned char * msg, unsigned char * pkt);
int  orc_packetizer_decode(orc_packetizer * p, const unsigned char * pkt, unsigned char * msg);

/* ------------------------------------------------------------- ofdm framing */
#define OFDMFLEXFRAME_PROTOCOL  105   /* 104 + packetizer version 1 */
#define OFDMFLEXFRAME_H_USER    8
#define OFDMFLEXFRAME_H_DEC     14
#define OFDMFLEXFRAME_H_ENC     36
#define OFDMFLEXFRAME_H_SYM     288

void ofdmframe_init_S0(const unsigned char * p, unsigned int M, cf32 * S0, cf32 * s0, unsigned int * M_S0);
void ofdmframe_init_S1(const unsigned char * p, unsigned int M, cf32 * S1, cf32 * s1, unsigned int * M_S1);

typedef struct ofdmframegen_s * ofdmframegen;
ofdmframegen ofdmframegen_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p);
void ofdmframegen_destroy(ofdmframegen q);
void ofdmframegen_reset(ofdmframegen q);
void ofdmframegen_write_S0a(ofdmframegen q, cf32 * y);
void ofdmframegen_write_S0b(ofdmframegen q, cf32 * y);
void ofdmframegen_write_S1(ofdmframegen q, cf32 * y);
void ofdmframegen_writesymbol(ofdmframegen q, const cf32 * X, cf32 * y);
void ofdmframegen_writetail(ofdmframegen q, cf32 * y);

typedef int (*ofdmframesync_callback)(cf32 * X, unsigned char * p, unsigned int M, void * userdata);
typedef struct ofdmframesync_s * ofdmframesync;
ofdmframesync ofdmframesync_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p,
                                   ofdmframesync_callback cb, void * userdata);
void  ofdmframesync_destroy(ofdmframesync q);
void  ofdmframesync_reset(ofdmframesync q);
void  ofdmframesync_execute(ofdmframesync q, const cf32 * x, unsigned int n);
float ofdmframesync_get_rssi(ofdmframesync q);
float ofdmframesync_get_cfo(ofdmframesync q);
/* side channel: indices into the stream fed to execute() */
uint64_t ofdmframesync_get_sample_index(ofdmframesync q);   /* index of the sample being processed */
uint64_t ofdmframesync_get_detect_index(ofdmframesync q);   /* index of the sample that tripped SEEK */

/* oracle-only extensions of the liquid API (not in liquid.h) */
void ofdmflexframesync_ext_get_indices(ofdmflexframesync q, uint64_t * detect_index, uint64_t * complete_index);
/* indices of the frame whose user callback is currently running (thread-local) */
void orc_callback_indices(uint64_t * detect_index, uint64_t * complete_index);
/* tap on every equalised symbol X[0..M) handed to the flexframe layer; userdata is the
 * user callback's userdata of that synchroniser */
typedef void (*orc_symbol_tap_fn)(void * userdata, const cf32 * X, unsigned int M, uint64_t sample_index);
void orc_set_symbol_tap(orc_symbol_tap_fn fn);
/* deterministic stand-in for liquid's rand() padding symbols (deviation D5) */
static inline unsigned int orc_pad_symbol(unsigned int slot, unsigned int M_const) {
    return ((slot * 2654435761u) >> 16) % M_const;
}

/* --------------------------------------------------------------- nco (shared) */
struct nco_crcf_s {
    liquid_ncotype type;
    uint32_t theta;     /* phase,     2*pi <-> 2^32 */
    uint32_t d_theta;   /* frequency, 2*pi <-> 2^32 */
};
uint32_t orc_nco_constrain(float theta);
static inline void orc_nco_sincos(uint32_t theta, float * s, float * c) {
    float t = (float)((double)(int32_t)theta * (M_PI / 2147483648.0));
    *s = sinf(t);
    *c = cosf(t);
}

#endif
