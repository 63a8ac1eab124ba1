/*
 * ofdmflexframe.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * ofdmflexframegen / ofdmflexframesync: header + payload framing on top of the OFDM PHY.
 * liquid-dsp 1.3.x src/framing/src/ofdmflexframegen.c, ofdmflexframesync.c.
 * Reference call sites: lib/multichanneltx.cc:71-80,161,185-188,234-236;
 * lib/multichannelrx.cc:82,140,194; lib/ofdmtxrx.cc:79-91,314-328,387,625.
 *
 * Frame = S0a, S0b, S1, ceil(288/M_data) header symbols (BPSK), ceil(payload_mod_len/M_data)
 * payload symbols, and (write() interface only) one tail buffer of M+cp samples.
 * Normative choices: D5 padding symbols come from orc_pad_symbol(), not rand();
 * D6 stats.evm is the header EVM (payload EVM is not accumulated).
 */
#include "oracle_internal.h"

/* ---------------------------------------------------------------- framegen */
enum { FG_S0a = 0, FG_S0b, FG_S1, FG_HEADER, FG_PAYLOAD, FG_TAIL, FG_NULL };

struct ofdmflexframegen_s {
    unsigned int M, cp_len, taper_len;
    unsigned char * p;
    unsigned int M_null, M_pilot, M_data;
    cf32 * X;
    ofdmframegen fg;
    unsigned int num_symbols_header, num_symbols_payload;
    orc_modem mod_header, mod_payload;
    orc_packetizer p_header, p_payload;
    unsigned char header[OFDMFLEXFRAME_H_DEC];
    unsigned char header_enc[OFDMFLEXFRAME_H_ENC];
    unsigned char header_mod[OFDMFLEXFRAME_H_SYM];
    unsigned int payload_dec_len, payload_enc_len, payload_mod_len;
    unsigned char * payload_enc, * payload_mod;
    unsigned int symbol_number;
    int state, frame_assembled, frame_complete;
    unsigned int header_symbol_index, payload_symbol_index, pad_index;
    ofdmflexframegenprops_s props;
    cf32 * buf_tx;
    unsigned int frame_len, buf_index;
};

static const ofdmflexframegenprops_s props_default = {LIQUID_CRC_32, LIQUID_FEC_NONE, LIQUID_FEC_HAMMING128, LIQUID_MODEM_QPSK};

void ofdmflexframegenprops_init_default(ofdmflexframegenprops_s * props) { *props = props_default; }

static void fg_reconfigure(ofdmflexframegen q)
{
    orc_packetizer_free(&q->p_payload);
    orc_packetizer_init(&q->p_payload, q->payload_dec_len, q->props.check, q->props.fec0, q->props.fec1);
    q->payload_enc_len = q->p_payload.packet_len;
    q->payload_enc = (unsigned char *)realloc(q->payload_enc, q->payload_enc_len + 8);
    orc_modem_init(&q->mod_payload, q->props.mod_scheme);
    unsigned int bps = q->mod_payload.bps;
    q->payload_mod_len = (8 * q->payload_enc_len) / bps + (((8 * q->payload_enc_len) % bps) ? 1 : 0);
    q->payload_mod = (unsigned char *)realloc(q->payload_mod, q->payload_mod_len + 8);
    q->num_symbols_payload = q->payload_mod_len / q->M_data + ((q->payload_mod_len % q->M_data) ? 1 : 0);
}

ofdmflexframegen ofdmflexframegen_create(unsigned int M, unsigned int cp_len, unsigned int taper_len,
                                         unsigned char * p, ofdmflexframegenprops_s * fgprops)
{
    if (M < 8 || (M % 2) || cp_len > M || taper_len > cp_len) {
        fprintf(stderr, "error: ofdmflexframegen_create(), invalid configuration\n");
        exit(1);
    }
    ofdmflexframegen q = (ofdmflexframegen)calloc(1, sizeof(struct ofdmflexframegen_s));
    q->M = M; q->cp_len = cp_len; q->taper_len = taper_len;
    q->p = (unsigned char *)malloc(M);
    if (p == NULL) ofdmframe_init_default_sctype(M, q->p);
    else memmove(q->p, p, M);
    ofdmframe_validate_sctype(q->p, M, &q->M_null, &q->M_pilot, &q->M_data);
    q->X = (cf32 *)calloc(M, sizeof(cf32));
    q->fg = ofdmframegen_create(M, cp_len, taper_len, q->p);
    q->frame_len = M + cp_len;
    q->buf_tx = (cf32 *)calloc(q->frame_len, sizeof(cf32));
    q->num_symbols_header = OFDMFLEXFRAME_H_SYM / q->M_data + ((OFDMFLEXFRAME_H_SYM % q->M_data) ? 1 : 0);
    orc_modem_init(&q->mod_header, LIQUID_MODEM_BPSK);
    orc_packetizer_init(&q->p_header, OFDMFLEXFRAME_H_DEC, LIQUID_CRC_32, LIQUID_FEC_GOLAY2412, LIQUID_FEC_NONE);
    if (q->p_header.packet_len != OFDMFLEXFRAME_H_ENC) {
        fprintf(stderr, "error: ofdmflexframegen_create(), header length mismatch\n");
        exit(1);
    }
    q->payload_dec_len = 1;
    q->props = props_default;
    ofdmflexframegen_setprops(q, fgprops);
    ofdmflexframegen_reset(q);
    return q;
}

void ofdmflexframegen_destroy(ofdmflexframegen q)
{
    ofdmframegen_destroy(q->fg);
    orc_packetizer_free(&q->p_header);
    orc_packetizer_free(&q->p_payload);
    free(q->payload_enc); free(q->payload_mod);
    free(q->X); free(q->p); free(q->buf_tx);
    free(q);
}

void ofdmflexframegen_reset(ofdmflexframegen q)
{
    q->symbol_number = 0;
    q->state = FG_S0a;
    q->frame_assembled = 0;
    q->frame_complete = 0;
    q->header_symbol_index = 0;
    q->payload_symbol_index = 0;
    q->pad_index = 0;
    q->buf_index = q->frame_len;
    ofdmframegen_reset(q->fg);
}

int ofdmflexframegen_is_assembled(ofdmflexframegen q) { return q->frame_assembled; }

void ofdmflexframegen_getprops(ofdmflexframegen q, ofdmflexframegenprops_s * props) { *props = q->props; }

void ofdmflexframegen_setprops(ofdmflexframegen q, ofdmflexframegenprops_s * props)
{
    const ofdmflexframegenprops_s * s = props ? props : &props_default;
    if (s->check == LIQUID_CRC_UNKNOWN || s->check >= LIQUID_CRC_NUM_SCHEMES ||
        s->fec0 == LIQUID_FEC_UNKNOWN || s->fec1 == LIQUID_FEC_UNKNOWN ||
        !orc_fec_supported(s->fec0) || !orc_fec_supported(s->fec1) || !orc_modem_supported(s->mod_scheme)) {
        fprintf(stderr, "error: ofdmflexframegen_setprops(), invalid/unsupported properties\n");
        exit(1);
    }
    q->props = *s;
    fg_reconfigure(q);
}

unsigned int ofdmflexframegen_getframelen(ofdmflexframegen q)
{
    return 3 + q->num_symbols_header + q->num_symbols_payload;
}

void ofdmflexframegen_print(ofdmflexframegen q)
{
    printf("ofdmflexframegen: M=%u cp=%u taper=%u null/pilot/data=%u/%u/%u payload=%u B enc=%u B syms=%u\n",
           q->M, q->cp_len, q->taper_len, q->M_null, q->M_pilot, q->M_data,
           q->payload_dec_len, q->payload_enc_len, q->payload_mod_len);
}

void ofdmflexframegen_assemble(ofdmflexframegen q, const unsigned char * header,
                               const unsigned char * payload, unsigned int payload_len)
{
    ofdmflexframegen_reset(q);
    if (payload_len != q->payload_dec_len) {
        q->payload_dec_len = payload_len;
        fg_reconfigure(q);
    }
    q->frame_assembled = 1;
    memmove(q->header, header, OFDMFLEXFRAME_H_USER);
    unsigned int n = OFDMFLEXFRAME_H_USER;
    q->header[n + 0] = OFDMFLEXFRAME_PROTOCOL;
    q->header[n + 1] = (q->payload_dec_len >> 8) & 0xff;
    q->header[n + 2] = (q->payload_dec_len) & 0xff;
    q->header[n + 3] = (unsigned char)q->props.mod_scheme;
    q->header[n + 4] = (unsigned char)(((q->props.check & 0x07) << 5) | (q->props.fec0 & 0x1f));
    q->header[n + 5] = (unsigned char)(q->props.fec1 & 0x1f);
    orc_packetizer_encode(&q->p_header, q->header, q->header_enc);
    orc_scramble(q->header_enc, OFDMFLEXFRAME_H_ENC);
    orc_repack_bytes(q->header_enc, 8, OFDMFLEXFRAME_H_ENC, q->header_mod, 1, OFDMFLEXFRAME_H_SYM, NULL);
    orc_packetizer_encode(&q->p_payload, payload, q->payload_enc);
    memset(q->payload_mod, 0x00, q->payload_mod_len);
    orc_repack_bytes(q->payload_enc, 8, q->payload_enc_len, q->payload_mod, q->mod_payload.bps, q->payload_mod_len, NULL);
}

static void fg_write_header(ofdmflexframegen q, cf32 * buffer)
{
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] == OFDMFRAME_SCTYPE_DATA) {
            if (q->header_symbol_index < OFDMFLEXFRAME_H_SYM)
                q->X[i] = orc_modem_modulate(&q->mod_header, q->header_mod[q->header_symbol_index++]);
            else
                q->X[i] = orc_modem_modulate(&q->mod_header, orc_pad_symbol(q->pad_index++, q->mod_header.M));
        } else {
            q->X[i] = 0.0f;
        }
    }
    ofdmframegen_writesymbol(q->fg, q->X, buffer);
    if (q->symbol_number == q->num_symbols_header) {
        q->symbol_number = 0;
        q->state = FG_PAYLOAD;
    }
}

static void fg_write_payload(ofdmflexframegen q, cf32 * buffer)
{
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] == OFDMFRAME_SCTYPE_DATA) {
            if (q->payload_symbol_index < q->payload_mod_len)
                q->X[i] = orc_modem_modulate(&q->mod_payload, q->payload_mod[q->payload_symbol_index++]);
            else
                q->X[i] = orc_modem_modulate(&q->mod_payload, orc_pad_symbol(q->pad_index++, q->mod_payload.M));
        } else {
            q->X[i] = 0.0f;
        }
    }
    ofdmframegen_writesymbol(q->fg, q->X, buffer);
    if (q->symbol_number == q->num_symbols_payload) q->state = FG_TAIL;
}

/* produce the next M+cp samples of the frame into buffer */
static void fg_gen_symbol(ofdmflexframegen q, cf32 * buffer)
{
    q->symbol_number++;
    switch (q->state) {
    case FG_S0a: ofdmframegen_write_S0a(q->fg, buffer); q->state = FG_S0b; break;
    case FG_S0b: ofdmframegen_write_S0b(q->fg, buffer); q->state = FG_S1; break;
    case FG_S1:  ofdmframegen_write_S1(q->fg, buffer); q->symbol_number = 0; q->state = FG_HEADER; break;
    case FG_HEADER:  fg_write_header(q, buffer); break;
    case FG_PAYLOAD: fg_write_payload(q, buffer); break;
    case FG_TAIL:
        memset(buffer, 0, q->frame_len * sizeof(cf32));
        ofdmframegen_writetail(q->fg, buffer);
        q->frame_complete = 1;
        q->frame_assembled = 0;
        q->state = FG_NULL;
        break;
    default:
        memset(buffer, 0, q->frame_len * sizeof(cf32));
        break;
    }
}

int ofdmflexframegen_write(ofdmflexframegen q, liquid_float_complex * buf, unsigned int buf_len)
{
    unsigned int i;
    for (i = 0; i < buf_len; i++) {
        if (q->buf_index >= q->frame_len) {
            fg_gen_symbol(q, q->buf_tx);
            q->buf_index = 0;
        }
        buf[i] = q->buf_tx[q->buf_index++];
    }
    return q->frame_complete;
}

/* pre-1.3 interface: one symbol per call, returns 1 with the last payload symbol, no tail */
int ofdmflexframegen_writesymbol(ofdmflexframegen q, liquid_float_complex * buffer)
{
    if (!q->frame_assembled) {
        memset(buffer, 0, q->frame_len * sizeof(cf32));
        return 1;
    }
    fg_gen_symbol(q, buffer);
    if (q->state == FG_TAIL) {
        ofdmflexframegen_reset(q);
        return 1;
    }
    return 0;
}

/* --------------------------------------------------------------- framesync */
enum { FS_HEADER = 0, FS_PAYLOAD };

struct ofdmflexframesync_s {
    unsigned int M, cp_len, taper_len;
    unsigned char * p;
    unsigned int M_null, M_pilot, M_data;
    framesync_callback callback;
    void * userdata;
    framesyncstats_s framestats;
    float evm_hat;
    ofdmframesync fs;
    orc_modem mod_header, mod_payload;
    orc_packetizer p_header, p_payload;
    unsigned char header[OFDMFLEXFRAME_H_DEC];
    unsigned char header_enc[OFDMFLEXFRAME_H_ENC];
    unsigned char header_mod[OFDMFLEXFRAME_H_SYM];
    int header_valid;
    unsigned int ms_payload, bps_payload, payload_len, check, fec0, fec1;
    unsigned int payload_enc_len, payload_mod_len;
    unsigned char * payload_enc, * payload_dec;
    int payload_valid;
    unsigned int symbol_counter, state, header_symbol_index, payload_symbol_index, payload_buffer_index;
    uint64_t complete_index;
};

static int fs_internal_callback(cf32 * X, unsigned char * p, unsigned int M, void * userdata);

/* oracle-only side channels (not in liquid.h): indices of the frame whose user callback is
 * running, and a tap on every equalised OFDM symbol the PHY delivers. */
static __thread uint64_t tls_detect_index, tls_complete_index;
static orc_symbol_tap_fn g_symbol_tap = NULL;
void orc_callback_indices(uint64_t * detect_index, uint64_t * complete_index)
{
    if (detect_index) *detect_index = tls_detect_index;
    if (complete_index) *complete_index = tls_complete_index;
}
void orc_set_symbol_tap(orc_symbol_tap_fn fn) { g_symbol_tap = fn; }

ofdmflexframesync ofdmflexframesync_create(unsigned int M, unsigned int cp_len, unsigned int taper_len,
                                           unsigned char * p, framesync_callback callback, void * userdata)
{
    if (M < 8 || (M % 2) || cp_len > M) {
        fprintf(stderr, "error: ofdmflexframesync_create(), invalid configuration\n");
        exit(1);
    }
    ofdmflexframesync q = (ofdmflexframesync)calloc(1, sizeof(struct ofdmflexframesync_s));
    q->M = M; q->cp_len = cp_len; q->taper_len = taper_len;
    q->callback = callback; q->userdata = userdata;
    q->p = (unsigned char *)malloc(M);
    if (p == NULL) ofdmframe_init_default_sctype(M, q->p);
    else memmove(q->p, p, M);
    ofdmframe_validate_sctype(q->p, M, &q->M_null, &q->M_pilot, &q->M_data);
    q->fs = ofdmframesync_create(M, cp_len, taper_len, q->p, fs_internal_callback, (void *)q);
    orc_modem_init(&q->mod_header, LIQUID_MODEM_BPSK);
    orc_packetizer_init(&q->p_header, OFDMFLEXFRAME_H_DEC, LIQUID_CRC_32, LIQUID_FEC_GOLAY2412, LIQUID_FEC_NONE);
    q->ms_payload = LIQUID_MODEM_QPSK;
    q->bps_payload = 2;
    q->payload_len = 1;
    q->check = LIQUID_CRC_32; q->fec0 = LIQUID_FEC_NONE; q->fec1 = LIQUID_FEC_NONE;
    orc_modem_init(&q->mod_payload, q->ms_payload);
    orc_packetizer_init(&q->p_payload, q->payload_len, q->check, q->fec0, q->fec1);
    q->payload_enc_len = q->p_payload.packet_len;
    q->payload_enc = (unsigned char *)calloc(q->payload_enc_len + 8, 1);
    q->payload_dec = (unsigned char *)calloc(q->payload_len + 8, 1);
    q->payload_mod_len = 0;
    ofdmflexframesync_reset(q);
    return q;
}

void ofdmflexframesync_destroy(ofdmflexframesync q)
{
    ofdmframesync_destroy(q->fs);
    orc_packetizer_free(&q->p_header);
    orc_packetizer_free(&q->p_payload);
    free(q->payload_enc); free(q->payload_dec); free(q->p);
    free(q);
}

void ofdmflexframesync_print(ofdmflexframesync q)
{
    printf("ofdmflexframesync: M=%u cp=%u null/pilot/data=%u/%u/%u\n", q->M, q->cp_len, q->M_null, q->M_pilot, q->M_data);
}

void ofdmflexframesync_reset(ofdmflexframesync q)
{
    q->symbol_counter = 0;
    q->state = FS_HEADER;
    q->header_symbol_index = 0;
    q->payload_symbol_index = 0;
    q->payload_buffer_index = 0;
    q->evm_hat = 0.0f;
    ofdmframesync_reset(q->fs);
}

void ofdmflexframesync_execute(ofdmflexframesync q, liquid_float_complex * x, unsigned int n)
{
    ofdmframesync_execute(q->fs, x, n);
}

float ofdmflexframesync_get_rssi(ofdmflexframesync q) { return ofdmframesync_get_rssi(q->fs); }
float ofdmflexframesync_get_cfo(ofdmflexframesync q) { return ofdmframesync_get_cfo(q->fs); }
void ofdmflexframesync_debug_enable(ofdmflexframesync q) { (void)q; }
void ofdmflexframesync_debug_disable(ofdmflexframesync q) { (void)q; }
void ofdmflexframesync_debug_print(ofdmflexframesync q, const char * filename) { (void)q; (void)filename; }

void ofdmflexframesync_ext_get_indices(ofdmflexframesync q, uint64_t * detect_index, uint64_t * complete_index)
{
    if (detect_index) *detect_index = ofdmframesync_get_detect_index(q->fs);
    if (complete_index) *complete_index = q->complete_index;
}

static void fs_decode_header(ofdmflexframesync q)
{
    unsigned int i;
    memset(q->header_enc, 0, OFDMFLEXFRAME_H_ENC);
    for (i = 0; i < OFDMFLEXFRAME_H_SYM; i++)
        q->header_enc[i >> 3] |= (unsigned char)((q->header_mod[i] & 1u) << (7 - (i & 7)));
    orc_scramble(q->header_enc, OFDMFLEXFRAME_H_ENC);
    q->header_valid = orc_packetizer_decode(&q->p_header, q->header_enc, q->header);
    if (!q->header_valid) return;
    unsigned int n = OFDMFLEXFRAME_H_USER;
    if (q->header[n + 0] != OFDMFLEXFRAME_PROTOCOL) { q->header_valid = 0; return; }
    unsigned int payload_len = ((unsigned int)q->header[n + 1] << 8) | q->header[n + 2];
    unsigned int mod_scheme = q->header[n + 3];
    unsigned int check = (q->header[n + 4] >> 5) & 0x07;
    unsigned int fec0 = q->header[n + 4] & 0x1f;
    unsigned int fec1 = q->header[n + 5] & 0x1f;
    /* range checks as upstream; schemes outside the implemented subset also invalidate */
    if (!orc_modem_supported(mod_scheme) || (check != LIQUID_CRC_32 && check != LIQUID_CRC_NONE) ||
        !orc_fec_supported(fec0) || !orc_fec_supported(fec1)) {
        q->header_valid = 0;
        return;
    }
    if (mod_scheme != q->ms_payload) {
        q->ms_payload = mod_scheme;
        orc_modem_init(&q->mod_payload, mod_scheme);
        q->bps_payload = q->mod_payload.bps;
    }
    q->payload_len = payload_len; q->check = check; q->fec0 = fec0; q->fec1 = fec1;
    orc_packetizer_free(&q->p_payload);
    orc_packetizer_init(&q->p_payload, payload_len, check, fec0, fec1);
    q->payload_enc_len = q->p_payload.packet_len;
    q->payload_enc = (unsigned char *)realloc(q->payload_enc, q->payload_enc_len + 8);
    q->payload_dec = (unsigned char *)realloc(q->payload_dec, q->payload_len + 8);
    memset(q->payload_enc, 0, q->payload_enc_len + 8);
    unsigned int bits = 8 * q->payload_enc_len;
    q->payload_mod_len = bits / q->bps_payload + ((bits % q->bps_payload) ? 1 : 0);
}

static void fs_rxheader(ofdmflexframesync q, cf32 * X)
{
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] != OFDMFRAME_SCTYPE_DATA) continue;
        unsigned int sym = orc_modem_demodulate(&q->mod_header, X[i]);
        q->header_mod[q->header_symbol_index++] = (unsigned char)sym;
        q->evm_hat += orc_modem_evm2(&q->mod_header);
        if (q->header_symbol_index == OFDMFLEXFRAME_H_SYM) {
            fs_decode_header(q);
            q->framestats.evm = 10 * log10f(q->evm_hat / OFDMFLEXFRAME_H_SYM);
            if (q->header_valid) {
                q->state = FS_PAYLOAD;
            } else {
                q->framestats.rssi = ofdmframesync_get_rssi(q->fs);
                q->framestats.cfo = ofdmframesync_get_cfo(q->fs);
                q->framestats.framesyms = NULL;
                q->framestats.num_framesyms = 0;
                q->framestats.mod_scheme = LIQUID_MODEM_UNKNOWN;
                q->framestats.mod_bps = 0;
                q->framestats.check = LIQUID_CRC_UNKNOWN;
                q->framestats.fec0 = LIQUID_FEC_UNKNOWN;
                q->framestats.fec1 = LIQUID_FEC_UNKNOWN;
                q->complete_index = ofdmframesync_get_sample_index(q->fs);
                tls_detect_index = ofdmframesync_get_detect_index(q->fs);
                tls_complete_index = q->complete_index;
                if (q->callback != NULL)
                    q->callback(q->header, q->header_valid, NULL, 0, 0, q->framestats, q->userdata);
                ofdmflexframesync_reset(q);
            }
            break;
        }
    }
}

static void fs_rxpayload(ofdmflexframesync q, cf32 * X)
{
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] != OFDMFRAME_SCTYPE_DATA) continue;
        unsigned int sym = orc_modem_demodulate(&q->mod_payload, X[i]);
        orc_pack_array(q->payload_enc, q->payload_enc_len, q->payload_buffer_index, q->bps_payload, (unsigned char)sym);
        q->payload_buffer_index += q->bps_payload;
        q->payload_symbol_index++;
        if (q->payload_symbol_index == q->payload_mod_len) {
            q->payload_valid = orc_packetizer_decode(&q->p_payload, q->payload_enc, q->payload_dec);
            q->complete_index = ofdmframesync_get_sample_index(q->fs);
            tls_detect_index = ofdmframesync_get_detect_index(q->fs);
            tls_complete_index = q->complete_index;
            if (q->callback != NULL) {
                q->framestats.rssi = ofdmframesync_get_rssi(q->fs);
                q->framestats.cfo = ofdmframesync_get_cfo(q->fs);
                q->framestats.framesyms = NULL;
                q->framestats.num_framesyms = 0;
                q->framestats.mod_scheme = q->ms_payload;
                q->framestats.mod_bps = q->bps_payload;
                q->framestats.check = q->check;
                q->framestats.fec0 = q->fec0;
                q->framestats.fec1 = q->fec1;
                q->callback(q->header, q->header_valid, q->payload_dec, q->payload_len, q->payload_valid,
                            q->framestats, q->userdata);
            }
            ofdmflexframesync_reset(q);
            break;
        }
    }
}

static int fs_internal_callback(cf32 * X, unsigned char * p, unsigned int M, void * userdata)
{
    (void)p; (void)M;
    ofdmflexframesync q = (ofdmflexframesync)userdata;
    q->symbol_counter++;
    if (g_symbol_tap) g_symbol_tap(q->userdata, X, M, ofdmframesync_get_sample_index(q->fs));
    if (q->state == FS_HEADER) fs_rxheader(q, X);
    else fs_rxpayload(q, X);
    return 0;
}
