/*
 * liquid/liquid.h -- interface declarations for the slice of the liquid-dsp C API
 * that jgaeddert/liquid-usrp's multichannel OFDM path calls.
 *
 * liquid-dsp itself is NOT vendored by the reference (configure.ac:56 only checks for
 * -lliquid) and is not installed in this image.  This header is written from scratch; it
 * declares -- it does not implement -- the handles, enums, structs and functions the
 * reference's lib/ and src/ files use on the hot path, with liquid-dsp 1.3.x names and
 * argument order so that the reference sources compile unmodified:
 *
 *   lib/multichannelrx.cc:82-100,140-194   ofdmflexframesync_*, firpfbch_crcf_*, nco_crcf_*
 *   lib/multichanneltx.cc:70-96,131-236    ofdmflexframegen_*,  firpfbch_crcf_*, nco_crcf_*
 *   lib/ofdmtxrx.cc:79-91,314-328,387,625  ofdmflexframegen_* (write + writesymbol), sync
 *   src/multichannel_tx.cc:46-49,92-94     liquid_getopt_str2*, liquid_print_*_schemes
 *   src/flexframe_rx.cc:179,240            msresamp_crcf_*
 *
 * Two independent implementations sit behind it:
 *   oracle/src/ (C files)          scalar CPU restatement (test infrastructure only)
 *   liquid-usrp_b200/csrc + host/  the B200 product (CUDA), subset used by ofdmtxrx/src
 */
#ifndef __LIQUID_COMPAT_H__
#define __LIQUID_COMPAT_H__

#ifdef __cplusplus
#  include <complex>
typedef std::complex<float> liquid_float_complex;
extern "C" {
#else
#  include <complex.h>
typedef float complex liquid_float_complex;
#endif

#define LIQUID_VERSION          "1.3.2-b200compat"
#define LIQUID_VERSION_NUMBER   1003002
const char * liquid_libversion(void);

/* ------------------------------------------------------------------ enums */

/* error-detection schemes (Appendix A of SURVEY.md) */
typedef enum {
    LIQUID_CRC_UNKNOWN = 0,
    LIQUID_CRC_NONE,
    LIQUID_CRC_CHECKSUM,
    LIQUID_CRC_8,
    LIQUID_CRC_16,
    LIQUID_CRC_24,
    LIQUID_CRC_32
} crc_scheme;
#define LIQUID_CRC_NUM_SCHEMES 7

/* forward error-correction schemes; only NONE, HAMMING128, GOLAY2412 and CONV_V27 are
 * implemented on this path, the remaining names keep the 1.3.x numbering */
typedef enum {
    LIQUID_FEC_UNKNOWN = 0,
    LIQUID_FEC_NONE,
    LIQUID_FEC_REP3,
    LIQUID_FEC_REP5,
    LIQUID_FEC_HAMMING74,
    LIQUID_FEC_HAMMING84,
    LIQUID_FEC_HAMMING128,
    LIQUID_FEC_GOLAY2412,
    LIQUID_FEC_SECDED2216,
    LIQUID_FEC_SECDED3932,
    LIQUID_FEC_SECDED7264,
    LIQUID_FEC_CONV_V27,
    LIQUID_FEC_CONV_V29,
    LIQUID_FEC_CONV_V39,
    LIQUID_FEC_CONV_V615,
    LIQUID_FEC_CONV_V27P23,
    LIQUID_FEC_CONV_V27P34,
    LIQUID_FEC_CONV_V27P45,
    LIQUID_FEC_CONV_V27P56,
    LIQUID_FEC_CONV_V27P67,
    LIQUID_FEC_CONV_V27P78,
    LIQUID_FEC_CONV_V29P23,
    LIQUID_FEC_CONV_V29P34,
    LIQUID_FEC_CONV_V29P45,
    LIQUID_FEC_CONV_V29P56,
    LIQUID_FEC_CONV_V29P67,
    LIQUID_FEC_CONV_V29P78,
    LIQUID_FEC_RS_M8
} fec_scheme;
#define LIQUID_FEC_NUM_SCHEMES 28

/* modulation schemes, 1.3.x numbering; BPSK/QPSK/QAM16/QAM64/QAM256 implemented */
typedef enum {
    LIQUID_MODEM_UNKNOWN = 0,
    LIQUID_MODEM_PSK2,   LIQUID_MODEM_PSK4,   LIQUID_MODEM_PSK8,   LIQUID_MODEM_PSK16,
    LIQUID_MODEM_PSK32,  LIQUID_MODEM_PSK64,  LIQUID_MODEM_PSK128, LIQUID_MODEM_PSK256,
    LIQUID_MODEM_DPSK2,  LIQUID_MODEM_DPSK4,  LIQUID_MODEM_DPSK8,  LIQUID_MODEM_DPSK16,
    LIQUID_MODEM_DPSK32, LIQUID_MODEM_DPSK64, LIQUID_MODEM_DPSK128,LIQUID_MODEM_DPSK256,
    LIQUID_MODEM_ASK2,   LIQUID_MODEM_ASK4,   LIQUID_MODEM_ASK8,   LIQUID_MODEM_ASK16,
    LIQUID_MODEM_ASK32,  LIQUID_MODEM_ASK64,  LIQUID_MODEM_ASK128, LIQUID_MODEM_ASK256,
    LIQUID_MODEM_QAM4,   LIQUID_MODEM_QAM8,   LIQUID_MODEM_QAM16,  LIQUID_MODEM_QAM32,
    LIQUID_MODEM_QAM64,  LIQUID_MODEM_QAM128, LIQUID_MODEM_QAM256,
    LIQUID_MODEM_APSK4,  LIQUID_MODEM_APSK8,  LIQUID_MODEM_APSK16, LIQUID_MODEM_APSK32,
    LIQUID_MODEM_APSK64, LIQUID_MODEM_APSK128,LIQUID_MODEM_APSK256,
    LIQUID_MODEM_BPSK,   LIQUID_MODEM_QPSK,   LIQUID_MODEM_OOK,
    LIQUID_MODEM_SQAM32, LIQUID_MODEM_SQAM128,
    LIQUID_MODEM_V29,    LIQUID_MODEM_ARB16OPT, LIQUID_MODEM_ARB32OPT,
    LIQUID_MODEM_ARB64OPT, LIQUID_MODEM_ARB128OPT, LIQUID_MODEM_ARB256OPT,
    LIQUID_MODEM_ARB64VT, LIQUID_MODEM_ARB
} modulation_scheme;
#define LIQUID_MODEM_NUM_SCHEMES 52

#define LIQUID_ANALYZER     0
#define LIQUID_SYNTHESIZER  1

typedef enum { LIQUID_NCO = 0, LIQUID_VCO } liquid_ncotype;

/* OFDM subcarrier types (ofdmframe.common) */
#define OFDMFRAME_SCTYPE_NULL   0
#define OFDMFRAME_SCTYPE_PILOT  1
#define OFDMFRAME_SCTYPE_DATA   2

/* ---------------------------------------------------------------- structs */

typedef struct {
    float evm;                          /* error vector magnitude [dB]               */
    float rssi;                         /* received signal strength indication [dB]  */
    float cfo;                          /* carrier frequency offset (f/Fs)           */
    liquid_float_complex * framesyms;   /* frame symbols (NULL on this path)         */
    unsigned int num_framesyms;
    unsigned int mod_scheme;
    unsigned int mod_bps;
    unsigned int check;
    unsigned int fec0;
    unsigned int fec1;
} framesyncstats_s;

/* user callback, invoked once per received frame (restated in src/multichannel_rx.cc:37-43) */
typedef int (*framesync_callback)(unsigned char *  _header,
                                  int              _header_valid,
                                  unsigned char *  _payload,
                                  unsigned int     _payload_len,
                                  int              _payload_valid,
                                  framesyncstats_s _stats,
                                  void *           _userdata);

/* brace-initialised in this order at lib/multichanneltx.cc:184 */
typedef struct {
    unsigned int check;
    unsigned int fec0;
    unsigned int fec1;
    unsigned int mod_scheme;
} ofdmflexframegenprops_s;

/* ---------------------------------------------------------------- handles */
typedef struct nco_crcf_s *          nco_crcf;
typedef struct firpfbch_crcf_s *     firpfbch_crcf;
typedef struct ofdmflexframegen_s *  ofdmflexframegen;
typedef struct ofdmflexframesync_s * ofdmflexframesync;
typedef struct msresamp_crcf_s *     msresamp_crcf;

/* --------------------------------------------------------------- utilities */
modulation_scheme liquid_getopt_str2mod(const char * _str);
fec_scheme        liquid_getopt_str2fec(const char * _str);
crc_scheme        liquid_getopt_str2crc(const char * _str);
void liquid_print_modulation_schemes(void);
void liquid_print_fec_schemes(void);
void liquid_print_crc_schemes(void);

/* --------------------------------------------------------------------- nco */
nco_crcf nco_crcf_create(liquid_ncotype _type);
void  nco_crcf_destroy(nco_crcf _q);
void  nco_crcf_reset(nco_crcf _q);
void  nco_crcf_set_frequency(nco_crcf _q, float _dtheta);
void  nco_crcf_adjust_frequency(nco_crcf _q, float _step);
float nco_crcf_get_frequency(nco_crcf _q);
void  nco_crcf_set_phase(nco_crcf _q, float _theta);
float nco_crcf_get_phase(nco_crcf _q);
void  nco_crcf_step(nco_crcf _q);
void  nco_crcf_mix_up(nco_crcf _q, liquid_float_complex _x, liquid_float_complex * _y);
void  nco_crcf_mix_down(nco_crcf _q, liquid_float_complex _x, liquid_float_complex * _y);
void  nco_crcf_mix_block_up(nco_crcf _q, liquid_float_complex * _x, liquid_float_complex * _y, unsigned int _n);
void  nco_crcf_mix_block_down(nco_crcf _q, liquid_float_complex * _x, liquid_float_complex * _y, unsigned int _n);

/* ---------------------------------------------------------------- firpfbch */
firpfbch_crcf firpfbch_crcf_create_kaiser(int _type, unsigned int _M, unsigned int _m, float _As);
void firpfbch_crcf_destroy(firpfbch_crcf _q);
void firpfbch_crcf_reset(firpfbch_crcf _q);
void firpfbch_crcf_synthesizer_execute(firpfbch_crcf _q, liquid_float_complex * _x, liquid_float_complex * _y);
void firpfbch_crcf_analyzer_execute(firpfbch_crcf _q, liquid_float_complex * _x, liquid_float_complex * _y);

/* -------------------------------------------------------- ofdmflexframegen */
void ofdmflexframegenprops_init_default(ofdmflexframegenprops_s * _props);
ofdmflexframegen ofdmflexframegen_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                                         unsigned char * _p, ofdmflexframegenprops_s * _fgprops);
void ofdmflexframegen_destroy(ofdmflexframegen _q);
void ofdmflexframegen_reset(ofdmflexframegen _q);
void ofdmflexframegen_print(ofdmflexframegen _q);
int  ofdmflexframegen_is_assembled(ofdmflexframegen _q);
void ofdmflexframegen_getprops(ofdmflexframegen _q, ofdmflexframegenprops_s * _props);
void ofdmflexframegen_setprops(ofdmflexframegen _q, ofdmflexframegenprops_s * _props);
unsigned int ofdmflexframegen_getframelen(ofdmflexframegen _q);
void ofdmflexframegen_assemble(ofdmflexframegen _q, const unsigned char * _header,
                               const unsigned char * _payload, unsigned int _payload_len);
/* 1.3.x streaming interface (lib/multichanneltx.cc:236, lib/ofdmtxrx.cc:328) */
int  ofdmflexframegen_write(ofdmflexframegen _q, liquid_float_complex * _buf, unsigned int _buf_len);
/* pre-1.3 one-symbol interface (lib/ofdmtxrx.cc:387) */
int  ofdmflexframegen_writesymbol(ofdmflexframegen _q, liquid_float_complex * _buffer);

/* ------------------------------------------------------- ofdmflexframesync */
ofdmflexframesync ofdmflexframesync_create(unsigned int _M, unsigned int _cp_len, unsigned int _taper_len,
                                           unsigned char * _p, framesync_callback _callback, void * _userdata);
void ofdmflexframesync_destroy(ofdmflexframesync _q);
void ofdmflexframesync_print(ofdmflexframesync _q);
void ofdmflexframesync_reset(ofdmflexframesync _q);
void ofdmflexframesync_execute(ofdmflexframesync _q, liquid_float_complex * _x, unsigned int _n);
float ofdmflexframesync_get_rssi(ofdmflexframesync _q);
float ofdmflexframesync_get_cfo(ofdmflexframesync _q);
void ofdmflexframesync_debug_enable(ofdmflexframesync _q);
void ofdmflexframesync_debug_disable(ofdmflexframesync _q);
void ofdmflexframesync_debug_print(ofdmflexframesync _q, const char * _filename);
/* extension of the B200 implementation (not in liquid-dsp): samples given to
 * ofdmflexframesync_execute() are processed in batches; this runs whatever is pending and
 * delivers the callbacks now.  reset/destroy flush implicitly. */
void ofdmflexframesync_flush(ofdmflexframesync _q);

/* ------------------------------------------------------------ ofdmframe.common */
void ofdmframe_init_default_sctype(unsigned int _M, unsigned char * _p);
void ofdmframe_validate_sctype(unsigned char * _p, unsigned int _M,
                               unsigned int * _M_null, unsigned int * _M_pilot, unsigned int * _M_data);

/* ---------------------------------------------------------------- msresamp */
msresamp_crcf msresamp_crcf_create(float _r, float _As);
void  msresamp_crcf_destroy(msresamp_crcf _q);
void  msresamp_crcf_reset(msresamp_crcf _q);
float msresamp_crcf_get_delay(msresamp_crcf _q);
void  msresamp_crcf_execute(msresamp_crcf _q, liquid_float_complex * _x, unsigned int _nx,
                            liquid_float_complex * _y, unsigned int * _ny);

#ifdef __cplusplus
}
#endif
#endif /* __LIQUID_COMPAT_H__ */
