/*
 * modem.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * Linear modems used on the OFDM path: BPSK (frame header), QPSK, square QAM 4/16/64/256.
 * liquid-dsp 1.3.x src/modem/src/modem_{bpsk,qpsk,qam}.c, modem_common.c
 * (modem_demodulate_linear_array_ref).  Selected per frame by the reference at
 * lib/multichanneltx.cc:184 and lib/ofdmtxrx.cc:314.
 */
#include "oracle_internal.h"

int orc_modem_supported(unsigned int scheme)
{
    switch (scheme) {
    case LIQUID_MODEM_BPSK: case LIQUID_MODEM_QPSK: case LIQUID_MODEM_QAM4:
    case LIQUID_MODEM_QAM16: case LIQUID_MODEM_QAM64: case LIQUID_MODEM_QAM256:
        return 1;
    default:
        return 0;
    }
}

unsigned int orc_modem_bps(unsigned int scheme)
{
    switch (scheme) {
    case LIQUID_MODEM_BPSK:   return 1;
    case LIQUID_MODEM_QPSK:   return 2;
    case LIQUID_MODEM_QAM4:   return 2;
    case LIQUID_MODEM_QAM16:  return 4;
    case LIQUID_MODEM_QAM64:  return 6;
    case LIQUID_MODEM_QAM256: return 8;
    default:                  return 0;
    }
}

void orc_modem_init(orc_modem * q, unsigned int scheme)
{
    memset(q, 0, sizeof(*q));
    q->scheme = scheme;
    q->bps = orc_modem_bps(scheme);
    q->M = 1u << q->bps;
    if (scheme == LIQUID_MODEM_BPSK || scheme == LIQUID_MODEM_QPSK) return;
    q->m_i = q->bps >> 1;
    q->m_q = q->bps >> 1;
    switch (q->M) {
    case 4:   q->alpha = 1.0f / sqrtf(2.0f);   break;
    case 16:  q->alpha = 1.0f / sqrtf(10.0f);  break;
    case 64:  q->alpha = 1.0f / sqrtf(42.0f);  break;
    case 256: q->alpha = 1.0f / sqrtf(170.0f); break;
    default:  q->alpha = 1.0f;
    }
    unsigned int k;
    for (k = 0; k < q->bps && k < 8; k++) q->ref[k] = (float)(1u << k) * q->alpha;
}

static unsigned int gray_encode(unsigned int s) { return s ^ (s >> 1); }
static unsigned int gray_decode(unsigned int s)
{
    unsigned int mask = s >> 1;
    while (mask) { s ^= mask; mask >>= 1; }
    return s;
}

cf32 orc_modem_modulate(orc_modem * q, unsigned int s)
{
    if (q->scheme == LIQUID_MODEM_BPSK) return s ? -1.0f : 1.0f;
    if (q->scheme == LIQUID_MODEM_QPSK)
        return (s & 0x01 ? -(float)M_SQRT1_2 : (float)M_SQRT1_2) +
               (s & 0x02 ? -(float)M_SQRT1_2 : (float)M_SQRT1_2) * _Complex_I;
    unsigned int s_i = gray_decode(s >> q->m_q);
    unsigned int s_q = gray_decode(s & ((1u << q->m_q) - 1));
    float vi = (float)(2 * (int)s_i - (int)(1u << q->m_i) + 1) * q->alpha;
    float vq = (float)(2 * (int)s_q - (int)(1u << q->m_q) + 1) * q->alpha;
    return vi + _Complex_I * vq;
}

static void demod_linear_ref(float v, unsigned int m, const float * ref, unsigned int * s, float * res)
{
    unsigned int sym = 0, i, k = m;
    for (i = 0; i < m; i++) {
        sym <<= 1;
        sym |= (v > 0) ? 1 : 0;
        v += (v > 0) ? -ref[k - 1] : ref[k - 1];
        k--;
    }
    *s = sym;
    *res = v;
}

unsigned int orc_modem_demodulate(orc_modem * q, cf32 x)
{
    unsigned int sym;
    q->r = x;
    if (q->scheme == LIQUID_MODEM_BPSK) {
        sym = (crealf(x) > 0) ? 0 : 1;
        q->x_hat = sym ? -1.0f : 1.0f;
        return sym;
    }
    if (q->scheme == LIQUID_MODEM_QPSK) {
        sym = (crealf(x) > 0 ? 0 : 1) + (cimagf(x) > 0 ? 0 : 2);
        q->x_hat = (sym & 0x01 ? -(float)M_SQRT1_2 : (float)M_SQRT1_2) +
                   (sym & 0x02 ? -(float)M_SQRT1_2 : (float)M_SQRT1_2) * _Complex_I;
        return sym;
    }
    unsigned int s_i, s_q;
    float res_i, res_q;
    demod_linear_ref(crealf(x), q->m_i, q->ref, &s_i, &res_i);
    demod_linear_ref(cimagf(x), q->m_q, q->ref, &s_q, &res_q);
    s_i = gray_encode(s_i);
    s_q = gray_encode(s_q);
    sym = (s_i << q->m_q) + s_q;
    q->x_hat = (crealf(x) - res_i) + _Complex_I * (cimagf(x) - res_q);
    return sym;
}

float orc_modem_evm2(const orc_modem * q)
{
    float dr = crealf(q->r) - crealf(q->x_hat);
    float di = cimagf(q->r) - cimagf(q->x_hat);
    return dr * dr + di * di;
}

/* --------------------------------------------------- string <-> enum helpers */
static const struct { const char * name; unsigned int id; } mod_names[] = {
    {"bpsk", LIQUID_MODEM_BPSK}, {"qpsk", LIQUID_MODEM_QPSK}, {"qam4", LIQUID_MODEM_QAM4},
    {"qam16", LIQUID_MODEM_QAM16}, {"qam64", LIQUID_MODEM_QAM64}, {"qam256", LIQUID_MODEM_QAM256}};
static const struct { const char * name; unsigned int id; } fec_names[] = {
    {"none", LIQUID_FEC_NONE}, {"h128", LIQUID_FEC_HAMMING128}, {"g2412", LIQUID_FEC_GOLAY2412},
    {"v27", LIQUID_FEC_CONV_V27}};
static const struct { const char * name; unsigned int id; } crc_names[] = {
    {"none", LIQUID_CRC_NONE}, {"crc32", LIQUID_CRC_32}};

modulation_scheme liquid_getopt_str2mod(const char * s)
{
    unsigned int i;
    for (i = 0; i < sizeof(mod_names) / sizeof(mod_names[0]); i++)
        if (strcmp(s, mod_names[i].name) == 0) return (modulation_scheme)mod_names[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2mod(), unknown/unsupported mod scheme : %s\n", s);
    return LIQUID_MODEM_UNKNOWN;
}
fec_scheme liquid_getopt_str2fec(const char * s)
{
    unsigned int i;
    for (i = 0; i < sizeof(fec_names) / sizeof(fec_names[0]); i++)
        if (strcmp(s, fec_names[i].name) == 0) return (fec_scheme)fec_names[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2fec(), unknown/unsupported fec scheme : %s\n", s);
    return LIQUID_FEC_UNKNOWN;
}
crc_scheme liquid_getopt_str2crc(const char * s)
{
    unsigned int i;
    for (i = 0; i < sizeof(crc_names) / sizeof(crc_names[0]); i++)
        if (strcmp(s, crc_names[i].name) == 0) return (crc_scheme)crc_names[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2crc(), unknown/unsupported crc scheme : %s\n", s);
    return LIQUID_CRC_UNKNOWN;
}
void liquid_print_modulation_schemes(void)
{
    unsigned int i;
    printf("          ");
    for (i = 0; i < sizeof(mod_names) / sizeof(mod_names[0]); i++) printf("%s%s", i ? ", " : "", mod_names[i].name);
    printf("\n");
}
void liquid_print_fec_schemes(void)
{
    unsigned int i;
    printf("          ");
    for (i = 0; i < sizeof(fec_names) / sizeof(fec_names[0]); i++) printf("%s%s", i ? ", " : "", fec_names[i].name);
    printf("\n");
}
void liquid_print_crc_schemes(void)
{
    unsigned int i;
    printf("          ");
    for (i = 0; i < sizeof(crc_names) / sizeof(crc_names[0]); i++) printf("%s%s", i ? ", " : "", crc_names[i].name);
    printf("\n");
}
