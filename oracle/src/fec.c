/*
 * fec.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * CRC-32, Hamming(12,8), Golay(24,12), convolutional r1/2 K=7 (+ Viterbi), byte interleaver,
 * scrambler, bit repacking and the two-stage packetizer.  liquid-dsp 1.3.x
 * src/fec/src/{crc,fec_hamming128,fec_golay2412,fec_conv,interleaver,packetizer}.c,
 * src/utility/src/{pack_bytes,scramble}.c; libfec viterbi27_port.c for the trellis metric
 * and tie-break.  Implicit in the reference at lib/multichanneltx.cc:184-188 (assemble),
 * lib/ofdmtxrx.cc:314-320 and inside ofdmflexframesync_execute (lib/multichannelrx.cc:194).
 */
#include "oracle_internal.h"

static unsigned int parity32(unsigned int v)
{
    v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
    return v & 1u;
}
static unsigned int count_ones(unsigned int v)
{
    unsigned int c = 0;
    while (v) { c += v & 1u; v >>= 1; }
    return c;
}

/* ------------------------------------------------------------------ crc32 */
uint32_t orc_crc32(const unsigned char * msg, unsigned int n)
{
    uint32_t key = ~0u, poly = 0xEDB88320u;     /* bit-reversed 0x04C11DB7 */
    unsigned int i, j;
    for (i = 0; i < n; i++) {
        key ^= msg[i];
        for (j = 0; j < 8; j++) {
            uint32_t mask = -(key & 1u);
            key = (key >> 1) ^ (poly & mask);
        }
    }
    return ~key;
}

/* ---------------------------------------------------------- hamming(12,8) */
unsigned int orc_hamming128_encode_symbol(unsigned int s)
{
    unsigned int c = ((s & 0x80) << 2) | ((s & 0x70) << 1) | (s & 0x0f);
    c |= parity32(c & 0x2aa) << 11;     /* p1 : positions 3,5,7,9,11  */
    c |= parity32(c & 0x266) << 10;     /* p2 : positions 3,6,7,10,11 */
    c |= parity32(c & 0x0e1) << 8;      /* p4 : positions 5,6,7,12    */
    c |= parity32(c & 0x00f) << 4;      /* p8 : positions 9..12       */
    return c;
}

unsigned int orc_hamming128_decode_symbol(unsigned int r)
{
    unsigned int z = 8 * parity32(r & 0x01f) + 4 * parity32(r & 0x1e1) +
                     2 * parity32(r & 0x666) +     parity32(r & 0xaaa);
    if (z) r ^= 1u << (12 - z);
    return ((r & 0x200) >> 2) | ((r & 0x0e0) >> 1) | (r & 0x00f);
}

static void hamming128_encode(unsigned int n, const unsigned char * dec, unsigned char * enc)
{
    unsigned int i, j = 0, r = n % 2;
    for (i = 0; i < n - r; i += 2) {
        unsigned int m0 = orc_hamming128_encode_symbol(dec[i]);
        unsigned int m1 = orc_hamming128_encode_symbol(dec[i + 1]);
        enc[j + 0] = (m0 >> 4) & 0xff;
        enc[j + 1] = ((m0 << 4) & 0xf0) | ((m1 >> 8) & 0x0f);
        enc[j + 2] = m1 & 0xff;
        j += 3;
    }
    if (r) {
        unsigned int m0 = orc_hamming128_encode_symbol(dec[n - 1]);
        enc[j + 0] = (m0 & 0x0ff0) >> 4;
        enc[j + 1] = (m0 & 0x000f) << 4;
    }
}

static void hamming128_decode(unsigned int n, const unsigned char * enc, unsigned char * dec)
{
    unsigned int i, j = 0, r = n % 2;
    for (i = 0; i < n - r; i += 2) {
        unsigned int m0 = ((unsigned int)enc[j] << 4) | (enc[j + 1] >> 4);
        unsigned int m1 = (((unsigned int)enc[j + 1] & 0x0f) << 8) | enc[j + 2];
        dec[i]     = (unsigned char)orc_hamming128_decode_symbol(m0);
        dec[i + 1] = (unsigned char)orc_hamming128_decode_symbol(m1);
        j += 3;
    }
    if (r) {
        unsigned int m0 = ((unsigned int)enc[j] << 4) | (enc[j + 1] >> 4);
        dec[n - 1] = (unsigned char)orc_hamming128_decode_symbol(m0);
    }
}

/* ----------------------------------------------------------- golay(24,12) */
static const unsigned int golay_P[12] = {
    0x08ed, 0x01db, 0x03b5, 0x0769, 0x0ed1, 0x0da3,
    0x0b47, 0x068f, 0x0d1d, 0x0a3b, 0x0477, 0x0ffe};

static unsigned int golay_mul_P(unsigned int v)
{
    unsigned int x = 0, i;
    for (i = 0; i < 12; i++) { x <<= 1; x |= parity32(golay_P[i] & v); }
    return x;
}

unsigned int orc_golay2412_encode_symbol(unsigned int s)
{
    s &= 0xfff;
    return (golay_mul_P(s) << 12) | s;      /* [ parity(12) | message(12) ] */
}

static int golay_parity_search(unsigned int v)
{
    int i;
    for (i = 0; i < 12; i++)
        if (count_ones(v ^ golay_P[i]) <= 2) return i;
    return -1;
}

unsigned int orc_golay2412_decode_symbol(unsigned int r)
{
    /* syndrome s = H r^T with H = [ I | P ] */
    unsigned int s = ((r >> 12) & 0xfff) ^ golay_mul_P(r & 0xfff);
    unsigned int e_hat = 0;
    if (count_ones(s) <= 3) {
        e_hat = (s << 12) & 0xfff000;
    } else {
        int si = golay_parity_search(s);
        if (si >= 0) {
            e_hat = ((s ^ golay_P[si]) << 12) | (1u << (11 - si));
        } else {
            unsigned int sP = golay_mul_P(s);
            if (count_ones(sP) <= 3) {
                e_hat = sP;
            } else {
                int pi = golay_parity_search(sP);
                if (pi >= 0) e_hat = (1u << (11 - pi + 12)) | (sP ^ golay_P[pi]);
            }
        }
    }
    return (r ^ e_hat) & 0x0fff;
}

/* 3 bytes -> two 12-bit symbols -> two 24-bit codewords -> 6 bytes; remainder bytes are
 * encoded one per codeword (8 message bits, 3 bytes out) */
static void golay2412_encode(unsigned int n, const unsigned char * dec, unsigned char * enc)
{
    unsigned int i = 0, j = 0, r = n % 3;
    for (i = 0; i < n - r; i += 3) {
        unsigned int s0 = ((unsigned int)dec[i] << 4) | (dec[i + 1] >> 4);
        unsigned int s1 = (((unsigned int)dec[i + 1] & 0x0f) << 8) | dec[i + 2];
        unsigned int v0 = orc_golay2412_encode_symbol(s0);
        unsigned int v1 = orc_golay2412_encode_symbol(s1);
        enc[j + 0] = (v0 >> 16) & 0xff; enc[j + 1] = (v0 >> 8) & 0xff; enc[j + 2] = v0 & 0xff;
        enc[j + 3] = (v1 >> 16) & 0xff; enc[j + 4] = (v1 >> 8) & 0xff; enc[j + 5] = v1 & 0xff;
        j += 6;
    }
    for (i = n - r; i < n; i++) {
        unsigned int v0 = orc_golay2412_encode_symbol(dec[i]);
        enc[j + 0] = (v0 >> 16) & 0xff; enc[j + 1] = (v0 >> 8) & 0xff; enc[j + 2] = v0 & 0xff;
        j += 3;
    }
}

static void golay2412_decode(unsigned int n, const unsigned char * enc, unsigned char * dec)
{
    unsigned int i = 0, j = 0, r = n % 3;
    for (i = 0; i < n - r; i += 3) {
        unsigned int v0 = ((unsigned int)enc[j] << 16) | ((unsigned int)enc[j + 1] << 8) | enc[j + 2];
        unsigned int v1 = ((unsigned int)enc[j + 3] << 16) | ((unsigned int)enc[j + 4] << 8) | enc[j + 5];
        unsigned int s0 = orc_golay2412_decode_symbol(v0);
        unsigned int s1 = orc_golay2412_decode_symbol(v1);
        dec[i]     = (s0 >> 4) & 0xff;
        dec[i + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
        dec[i + 2] = s1 & 0xff;
        j += 6;
    }
    for (i = n - r; i < n; i++) {
        unsigned int v0 = ((unsigned int)enc[j] << 16) | ((unsigned int)enc[j + 1] << 8) | enc[j + 2];
        dec[i] = orc_golay2412_decode_symbol(v0) & 0xff;
        j += 3;
    }
}

/* ------------------------------------------------------- conv r1/2, K = 7 */
#define V27_POLYA 0x6d
#define V27_POLYB 0x4f
#define V27_K     7

static void conv27_encode(unsigned int n, const unsigned char * dec, unsigned char * enc)
{
    unsigned int i, j, nbit = 0, sr = 0;
    unsigned char byte_out = 0;
    static const unsigned int poly[2] = {V27_POLYA, V27_POLYB};
    unsigned int total = 8 * n + V27_K - 1;
    for (i = 0; i < total; i++) {
        unsigned int bit = (i < 8 * n) ? (dec[i >> 3] >> (7 - (i & 7))) & 1u : 0u;
        sr = (sr << 1) | bit;
        for (j = 0; j < 2; j++) {
            byte_out = (unsigned char)((byte_out << 1) | parity32(sr & poly[j]));
            enc[nbit / 8] = byte_out;
            nbit++;
        }
    }
    while (nbit % 8) {
        byte_out <<= 1;
        enc[nbit / 8] = byte_out;
        nbit++;
    }
}

/* hard-decision Viterbi: received bit b -> soft symbol 0/255 (liquid fec_conv_decode_hard),
 * libfec metric = sum |expected - received|, start metrics 0 (state 0) / 63 (others), strict
 * "m1 < m0" selects the predecessor with the oldest bit set, chain back from state 0. */
static void conv27_decode(unsigned int n, const unsigned char * enc, unsigned char * dec)
{
    unsigned int nbits = 8 * n + V27_K - 1;
    unsigned char * decisions = (unsigned char *)malloc((size_t)nbits * 64);
    unsigned int metric[64], next[64];
    unsigned int s, t;
    unsigned char exp0[128], exp1[128];     /* expected outputs per 7-bit register */
    for (s = 0; s < 128; s++) { exp0[s] = (unsigned char)parity32(s & V27_POLYA); exp1[s] = (unsigned char)parity32(s & V27_POLYB); }
    for (s = 0; s < 64; s++) metric[s] = 63;
    metric[0] = 0;
    for (t = 0; t < nbits; t++) {
        unsigned int r0 = (enc[(2 * t) >> 3] >> (7 - ((2 * t) & 7))) & 1u;
        unsigned int r1 = (enc[(2 * t + 1) >> 3] >> (7 - ((2 * t + 1) & 7))) & 1u;
        for (s = 0; s < 64; s++) {
            unsigned int p0 = s >> 1, p1 = (s >> 1) | 32, b = s & 1u;
            unsigned int reg0 = (p0 << 1) | b, reg1 = (p1 << 1) | b;
            unsigned int m0 = metric[p0] + 255 * ((exp0[reg0] ^ r0) + (exp1[reg0] ^ r1));
            unsigned int m1 = metric[p1] + 255 * ((exp0[reg1] ^ r0) + (exp1[reg1] ^ r1));
            unsigned int d = (m1 < m0) ? 1u : 0u;
            next[s] = d ? m1 : m0;
            decisions[(size_t)t * 64 + s] = (unsigned char)d;
        }
        memcpy(metric, next, sizeof(metric));
    }
    memset(dec, 0, n);
    s = 0;
    for (t = nbits; t-- > 0;) {
        if (t < 8 * n) dec[t >> 3] |= (unsigned char)((s & 1u) << (7 - (t & 7)));
        s = (s >> 1) | ((unsigned int)decisions[(size_t)t * 64 + s] << 5);
    }
    free(decisions);
}

/* ------------------------------------------------------------ fec generic */
int orc_fec_supported(unsigned int scheme)
{
    return scheme == LIQUID_FEC_NONE || scheme == LIQUID_FEC_HAMMING128 ||
           scheme == LIQUID_FEC_GOLAY2412 || scheme == LIQUID_FEC_CONV_V27;
}

static unsigned int block_enc_len(unsigned int dec_len, unsigned int m, unsigned int k)
{
    unsigned int bits_in = dec_len * 8;
    unsigned int blocks = bits_in / m + ((bits_in % m) ? 1 : 0);
    unsigned int bits_out = blocks * k;
    return bits_out / 8 + ((bits_out % 8) ? 1 : 0);
}

unsigned int orc_fec_enc_len(unsigned int scheme, unsigned int dec_len)
{
    switch (scheme) {
    case LIQUID_FEC_NONE:       return dec_len;
    case LIQUID_FEC_HAMMING128: return block_enc_len(dec_len, 8, 12);
    case LIQUID_FEC_GOLAY2412:  return block_enc_len(dec_len, 12, 24);
    case LIQUID_FEC_CONV_V27: {
        unsigned int bits_out = (dec_len * 8 + V27_K - 1) * 2;
        return bits_out / 8 + ((bits_out % 8) ? 1 : 0);
    }
    default:
        fprintf(stderr, "error: orc_fec_enc_len(), unsupported fec scheme %u\n", scheme);
        exit(1);
    }
}

void orc_fec_encode(unsigned int scheme, unsigned int dec_len, const unsigned char * dec, unsigned char * enc)
{
    switch (scheme) {
    case LIQUID_FEC_NONE:       memmove(enc, dec, dec_len); break;
    case LIQUID_FEC_HAMMING128: hamming128_encode(dec_len, dec, enc); break;
    case LIQUID_FEC_GOLAY2412:  golay2412_encode(dec_len, dec, enc); break;
    case LIQUID_FEC_CONV_V27:   conv27_encode(dec_len, dec, enc); break;
    default: fprintf(stderr, "error: orc_fec_encode(), unsupported fec scheme %u\n", scheme); exit(1);
    }
}

void orc_fec_decode(unsigned int scheme, unsigned int dec_len, const unsigned char * enc, unsigned char * dec)
{
    switch (scheme) {
    case LIQUID_FEC_NONE:       memmove(dec, enc, dec_len); break;
    case LIQUID_FEC_HAMMING128: hamming128_decode(dec_len, enc, dec); break;
    case LIQUID_FEC_GOLAY2412:  golay2412_decode(dec_len, enc, dec); break;
    case LIQUID_FEC_CONV_V27:   conv27_decode(dec_len, enc, dec); break;
    default: fprintf(stderr, "error: orc_fec_decode(), unsupported fec scheme %u\n", scheme); exit(1);
    }
}

/* ------------------------------------------------------------ interleaver */
static void il_permute(unsigned char * x, unsigned int n, unsigned int M, unsigned int N, unsigned char mask)
{
    unsigned int i, j, m = 0, c = n / 3, n2 = n / 2;
    for (i = 0; i < n2; i++) {
        do {
            j = m * N + c;
            m++;
            if (m == M) { c = (c + 1) % N; m = 0; }
        } while (j >= n2);
        unsigned char a = x[2 * i], b = x[2 * j + 1];
        x[2 * i]     = (unsigned char)((a & ~mask) | (b & mask));
        x[2 * j + 1] = (unsigned char)((a & mask) | (b & ~mask));
    }
}

void orc_interleave(unsigned char * x, unsigned int n, unsigned int depth, int decode)
{
    unsigned int M = 1 + (unsigned int)floorf(sqrtf((float)n));
    unsigned int N = n / M;
    while (n >= M * N) N++;
    if (!decode) {
        if (depth > 0) il_permute(x, n, M, N,     0xff);
        if (depth > 1) il_permute(x, n, M, N + 2, 0x0f);
        if (depth > 2) il_permute(x, n, M, N + 4, 0x55);
        if (depth > 3) il_permute(x, n, M, N + 8, 0x33);
    } else {
        if (depth > 3) il_permute(x, n, M, N + 8, 0x33);
        if (depth > 2) il_permute(x, n, M, N + 4, 0x55);
        if (depth > 1) il_permute(x, n, M, N + 2, 0x0f);
        if (depth > 0) il_permute(x, n, M, N,     0xff);
    }
}

/* -------------------------------------------------------- scramble / pack */
void orc_scramble(unsigned char * x, unsigned int n)
{
    static const unsigned char mask[4] = {0xb4, 0x6a, 0x8b, 0x45};
    unsigned int i;
    for (i = 0; i < n; i++) x[i] ^= mask[i & 3];
}

void orc_repack_bytes(const unsigned char * in, unsigned int bps_in, unsigned int n_in,
                      unsigned char * out, unsigned int bps_out, unsigned int n_out,
                      unsigned int * n_written)
{
    unsigned int total_bits = n_in * bps_in;
    unsigned int req = total_bits / bps_out + ((total_bits % bps_out) ? 1 : 0);
    if (n_out < req) {
        fprintf(stderr, "error: orc_repack_bytes(), output too short\n");
        exit(1);
    }
    unsigned int i, i_in = 0, i_out = 0, k = 0, n = 0;
    unsigned char s_in = 0, s_out = 0;
    for (i = 0; i < total_bits; i++) {
        s_out <<= 1;
        if (k == 0) s_in = in[i_in++];
        s_out |= (s_in >> (bps_in - k - 1)) & 0x01;
        if (n == bps_out - 1) { out[i_out++] = s_out; s_out = 0; }
        k = (k + 1) % bps_in;
        n = (n + 1) % bps_out;
    }
    if (i_out != req) {
        for (i = n; i < bps_out; i++) s_out <<= 1;
        out[i_out++] = s_out;
    }
    if (n_written) *n_written = i_out;
}

/* write the b-bit symbol `sym` MSB-first at bit index k of dst[0..n); bits past the end drop */
void orc_pack_array(unsigned char * dst, unsigned int n, unsigned int k, unsigned int b, unsigned char sym)
{
    unsigned int i;
    for (i = 0; i < b; i++) {
        unsigned int bit = (sym >> (b - 1 - i)) & 1u;
        unsigned int pos = k + i;
        if ((pos >> 3) >= n) return;
        unsigned char m = (unsigned char)(0x80u >> (pos & 7));
        if (bit) dst[pos >> 3] |= m; else dst[pos >> 3] &= (unsigned char)~m;
    }
}

/* ------------------------------------------------------------- packetizer */
unsigned int orc_packetizer_enc_len(unsigned int n, unsigned int check, unsigned int fec0, unsigned int fec1)
{
    unsigned int crc_len = (check == LIQUID_CRC_32) ? 4 : 0;
    return orc_fec_enc_len(fec1, orc_fec_enc_len(fec0, n + crc_len));
}

void orc_packetizer_init(orc_packetizer * p, unsigned int n, unsigned int check, unsigned int fec0, unsigned int fec1)
{
    memset(p, 0, sizeof(*p));
    if (check != LIQUID_CRC_32 && check != LIQUID_CRC_NONE) {
        fprintf(stderr, "error: orc_packetizer_init(), unsupported crc scheme %u\n", check);
        exit(1);
    }
    p->msg_len = n;
    p->check = check;
    p->crc_len = (check == LIQUID_CRC_32) ? 4 : 0;
    p->fs[0] = fec0; p->fs[1] = fec1;
    unsigned int n0 = n + p->crc_len, i;
    for (i = 0; i < 2; i++) {
        p->dec_len[i] = n0;
        p->enc_len[i] = orc_fec_enc_len(p->fs[i], n0);
        p->depth[i] = (p->fs[i] == LIQUID_FEC_NONE) ? 0 : 4;
        n0 = p->enc_len[i];
    }
    p->packet_len = n0;
    p->buf0 = (unsigned char *)calloc(p->packet_len + 8, 1);
    p->buf1 = (unsigned char *)calloc(p->packet_len + 8, 1);
}

void orc_packetizer_free(orc_packetizer * p) { free(p->buf0); free(p->buf1); p->buf0 = p->buf1 = NULL; }

void orc_packetizer_encode(orc_packetizer * p, const unsigned char * msg, unsigned char * pkt)
{
    unsigned int i;
    memmove(p->buf0, msg, p->msg_len);
    if (p->crc_len) {
        uint32_t key = orc_crc32(p->buf0, p->msg_len);
        for (i = 0; i < 4; i++) p->buf0[p->msg_len + i] = (key >> (8 * (3 - i))) & 0xff;
    }
    for (i = 0; i < 2; i++) {
        orc_fec_encode(p->fs[i], p->dec_len[i], p->buf0, p->buf1);
        orc_interleave(p->buf1, p->enc_len[i], p->depth[i], 0);
        memmove(p->buf0, p->buf1, p->enc_len[i]);
    }
    memmove(pkt, p->buf0, p->packet_len);
}

int orc_packetizer_decode(orc_packetizer * p, const unsigned char * pkt, unsigned char * msg)
{
    unsigned int i;
    memmove(p->buf0, pkt, p->packet_len);
    for (i = 2; i-- > 0;) {
        orc_interleave(p->buf0, p->enc_len[i], p->depth[i], 1);
        orc_fec_decode(p->fs[i], p->dec_len[i], p->buf0, p->buf1);
        memmove(p->buf0, p->buf1, p->dec_len[i]);
    }
    memmove(msg, p->buf0, p->msg_len);
    if (!p->crc_len) return 1;
    uint32_t key = 0;
    for (i = 0; i < 4; i++) key = (key << 8) | p->buf0[p->msg_len + i];
    return orc_crc32(p->buf0, p->msg_len) == key;
}
