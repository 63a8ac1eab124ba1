/*
 * ofdmframe.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * OFDM PHY framing: default subcarrier allocation, S0/S1 preambles, symbol generator
 * (ofdmframegen) and synchroniser (ofdmframesync).  liquid-dsp 1.3.x
 * src/framing/src/ofdmframe.common.c, ofdmframegen.c, ofdmframesync.c.  These are the
 * objects underneath ofdmflexframe{gen,sync}, which the reference creates at
 * lib/multichannelrx.cc:82, lib/multichanneltx.cc:80 and lib/ofdmtxrx.cc:84,91.
 *
 * Normative choices (unverifiable against upstream here; see oracle/README.md):
 *   D1 nco_rx uses the exact uint32 phase accumulator with direct sinf/cosf
 *   D3 S1 equaliser-gain polynomial fit (order 4) is evaluated in double precision
 *   D4 the rxsymbol phase-error wrap uses 2*pi
 */
#include "oracle_internal.h"

/* ------------------------------------------------------------------ common */
void ofdmframe_init_default_sctype(unsigned int M, unsigned char * p)
{
    unsigned int i, M2 = M / 2;
    unsigned int G = M / 10;
    if (G < 2) G = 2;
    unsigned int P = (M > 34) ? 8 : 4;
    unsigned int P2 = P / 2;
    for (i = 0; i < M; i++) p[i] = OFDMFRAME_SCTYPE_NULL;
    for (i = 1; i < M2 - G; i++) {
        unsigned char t = (((i + P2) % P) == 0) ? OFDMFRAME_SCTYPE_PILOT : OFDMFRAME_SCTYPE_DATA;
        p[i] = t;
        p[M - i] = t;
    }
}

void ofdmframe_validate_sctype(unsigned char * p, unsigned int M,
                               unsigned int * M_null, unsigned int * M_pilot, unsigned int * M_data)
{
    unsigned int i, n0 = 0, n1 = 0, n2 = 0;
    for (i = 0; i < M; i++) {
        if (p[i] == OFDMFRAME_SCTYPE_NULL) n0++;
        else if (p[i] == OFDMFRAME_SCTYPE_PILOT) n1++;
        else if (p[i] == OFDMFRAME_SCTYPE_DATA) n2++;
        else {
            fprintf(stderr, "error: ofdmframe_validate_sctype(), invalid subcarrier type (%u)\n", p[i]);
            exit(1);
        }
    }
    *M_null = n0; *M_pilot = n1; *M_data = n2;
}

static unsigned int nextpow2(unsigned int x)
{
    unsigned int n = 0;
    x--;
    while (x > 0) { x >>= 1; n++; }
    return n;
}

static void init_plcp(const unsigned char * p, unsigned int M, cf32 * S, cf32 * s, unsigned int * M_S, int is_S1)
{
    unsigned int i, m = nextpow2(M);
    if (m < 4) m = 4; else if (m > 8) m = 8;
    if (is_S1) m++;
    orc_mseq ms;
    orc_mseq_init_default(&ms, m);
    unsigned int count = 0;
    for (i = 0; i < M; i++) {
        unsigned int sym = orc_mseq_symbol(&ms, 3) & 0x01;
        if (p[i] == OFDMFRAME_SCTYPE_NULL || (!is_S1 && (i % 2) != 0)) {
            S[i] = 0.0f;
        } else {
            S[i] = sym ? 1.0f : -1.0f;
            count++;
        }
    }
    if (count == 0) {
        fprintf(stderr, "error: ofdmframe_init_S%d(), no subcarriers enabled\n", is_S1);
        exit(1);
    }
    orc_fft * ifft = orc_fft_create(M, ORC_FFT_BACKWARD);
    orc_fft_execute(ifft, S, s);
    orc_fft_destroy(ifft);
    float g = 1.0f / sqrtf((float)count);
    for (i = 0; i < M; i++) s[i] *= g;
    *M_S = count;
}

void ofdmframe_init_S0(const unsigned char * p, unsigned int M, cf32 * S0, cf32 * s0, unsigned int * M_S0)
{ init_plcp(p, M, S0, s0, M_S0, 0); }
void ofdmframe_init_S1(const unsigned char * p, unsigned int M, cf32 * S1, cf32 * s1, unsigned int * M_S1)
{ init_plcp(p, M, S1, s1, M_S1, 1); }

/* ---------------------------------------------------------------- framegen */
struct ofdmframegen_s {
    unsigned int M, cp_len, taper_len;
    unsigned char * p;
    unsigned int M_null, M_pilot, M_data, M_S0, M_S1;
    float g_data;
    float * taper;
    cf32 * postfix;
    cf32 * X, * x;
    orc_fft * ifft;
    cf32 * S0, * s0, * S1, * s1;
    orc_mseq ms_pilot;
};

ofdmframegen ofdmframegen_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p)
{
    if (M < 2 || (M % 2) || cp_len > M || taper_len > cp_len) {
        fprintf(stderr, "error: ofdmframegen_create(), invalid configuration\n");
        exit(1);
    }
    ofdmframegen q = (ofdmframegen)calloc(1, sizeof(struct ofdmframegen_s));
    q->M = M; q->cp_len = cp_len; q->taper_len = taper_len;
    q->p = (unsigned char *)malloc(M);
    if (p == NULL) ofdmframe_init_default_sctype(M, q->p);
    else memmove(q->p, p, M);
    ofdmframe_validate_sctype(q->p, M, &q->M_null, &q->M_pilot, &q->M_data);
    if (q->M_pilot + q->M_data == 0 || q->M_data == 0 || q->M_pilot < 2) {
        fprintf(stderr, "error: ofdmframegen_create(), need at least one data and two pilot subcarriers\n");
        exit(1);
    }
    q->X = (cf32 *)calloc(M, sizeof(cf32));
    q->x = (cf32 *)calloc(M, sizeof(cf32));
    q->ifft = orc_fft_create(M, ORC_FFT_BACKWARD);
    q->S0 = (cf32 *)malloc(M * sizeof(cf32)); q->s0 = (cf32 *)malloc(M * sizeof(cf32));
    q->S1 = (cf32 *)malloc(M * sizeof(cf32)); q->s1 = (cf32 *)malloc(M * sizeof(cf32));
    ofdmframe_init_S0(q->p, M, q->S0, q->s0, &q->M_S0);
    ofdmframe_init_S1(q->p, M, q->S1, q->s1, &q->M_S1);
    q->taper = (float *)malloc((taper_len + 1) * sizeof(float));
    q->postfix = (cf32 *)calloc(taper_len + 1, sizeof(cf32));
    unsigned int i;
    for (i = 0; i < taper_len; i++) {
        float t = ((float)i + 0.5f) / (float)taper_len;
        float g = sinf(M_PI_2 * t);
        q->taper[i] = g * g;
    }
    q->g_data = 1.0f / sqrtf((float)(q->M_pilot + q->M_data));
    orc_mseq_init_default(&q->ms_pilot, 8);
    return q;
}

void ofdmframegen_destroy(ofdmframegen q)
{
    free(q->p); free(q->X); free(q->x); free(q->S0); free(q->s0); free(q->S1); free(q->s1);
    free(q->taper); free(q->postfix);
    orc_fft_destroy(q->ifft);
    free(q);
}

void ofdmframegen_reset(ofdmframegen q)
{
    orc_mseq_reset(&q->ms_pilot);
    unsigned int i;
    for (i = 0; i < q->taper_len; i++) q->postfix[i] = 0.0f;
}

static void gensymbol(ofdmframegen q, cf32 * y)
{
    unsigned int i;
    memmove(y, &q->x[q->M - q->cp_len], q->cp_len * sizeof(cf32));
    memmove(&y[q->cp_len], q->x, q->M * sizeof(cf32));
    for (i = 0; i < q->taper_len; i++) {
        y[i] *= q->taper[i];
        y[i] += q->postfix[i] * q->taper[q->taper_len - i - 1];
    }
    memmove(q->postfix, q->x, q->taper_len * sizeof(cf32));
}

void ofdmframegen_write_S0a(ofdmframegen q, cf32 * y)
{
    unsigned int i;
    for (i = 0; i < q->M + q->cp_len; i++) y[i] = q->s0[(i + q->M - 2 * q->cp_len) % q->M];
    for (i = 0; i < q->taper_len; i++) y[i] *= q->taper[i];
}

void ofdmframegen_write_S0b(ofdmframegen q, cf32 * y)
{
    unsigned int i;
    for (i = 0; i < q->M + q->cp_len; i++) y[i] = q->s0[(i + q->M - q->cp_len) % q->M];
    memmove(q->postfix, q->s0, q->taper_len * sizeof(cf32));
}

void ofdmframegen_write_S1(ofdmframegen q, cf32 * y)
{
    memmove(q->x, q->s1, q->M * sizeof(cf32));
    gensymbol(q, y);
}

void ofdmframegen_writesymbol(ofdmframegen q, const cf32 * X, cf32 * y)
{
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        unsigned int k = (i + q->M / 2) % q->M;
        if (q->p[k] == OFDMFRAME_SCTYPE_NULL) q->X[k] = 0.0f;
        else if (q->p[k] == OFDMFRAME_SCTYPE_PILOT) q->X[k] = (orc_mseq_advance(&q->ms_pilot) ? 1.0f : -1.0f) * q->g_data;
        else q->X[k] = X[k] * q->g_data;
    }
    orc_fft_execute(q->ifft, q->X, q->x);
    gensymbol(q, y);
}

void ofdmframegen_writetail(ofdmframegen q, cf32 * y)
{
    unsigned int i;
    for (i = 0; i < q->taper_len; i++) y[i] = q->postfix[i] * q->taper[q->taper_len - i - 1];
}

/* --------------------------------------------------------------- framesync */
enum { ST_SEEKPLCP = 0, ST_PLCPSHORT0, ST_PLCPSHORT1, ST_PLCPLONG, ST_RXSYMBOLS };

struct ofdmframesync_s {
    unsigned int M, M2, cp_len;
    unsigned char * p;
    unsigned int M_null, M_pilot, M_data, M_S0, M_S1;
    orc_fft * fft;
    cf32 * X, * x;
    orc_window input_buffer;
    cf32 * S0, * s0, * S1, * s1;
    float g0;
    cf32 * G0, * G1, * G, * B, * R;
    int state;
    struct nco_crcf_s nco_rx;
    orc_mseq ms_pilot;
    float phi_prime, p1_prime;
    float plateau_threshold;
    cf32 s_hat_0, s_hat_1;
    int timer;
    unsigned int num_symbols;
    unsigned int backoff;
    ofdmframesync_callback callback;
    void * userdata;
    uint64_t sample_index, detect_index;
};

ofdmframesync ofdmframesync_create(unsigned int M, unsigned int cp_len, unsigned int taper_len, const unsigned char * p,
                                   ofdmframesync_callback cb, void * userdata)
{
    (void)taper_len;
    if (M < 8 || (M % 2) || cp_len > M) {
        fprintf(stderr, "error: ofdmframesync_create(), invalid configuration\n");
        exit(1);
    }
    ofdmframesync q = (ofdmframesync)calloc(1, sizeof(struct ofdmframesync_s));
    q->M = M; q->M2 = M / 2; q->cp_len = cp_len;
    q->p = (unsigned char *)malloc(M);
    if (p == NULL) ofdmframe_init_default_sctype(M, q->p);
    else memmove(q->p, p, M);
    ofdmframe_validate_sctype(q->p, M, &q->M_null, &q->M_pilot, &q->M_data);
    if (q->M_data == 0 || q->M_pilot < 2) {
        fprintf(stderr, "error: ofdmframesync_create(), need at least one data and two pilot subcarriers\n");
        exit(1);
    }
    q->X = (cf32 *)calloc(M, sizeof(cf32));
    q->x = (cf32 *)calloc(M, sizeof(cf32));
    q->fft = orc_fft_create(M, ORC_FFT_FORWARD);
    orc_window_init(&q->input_buffer, M + cp_len);
    q->S0 = (cf32 *)malloc(M * sizeof(cf32)); q->s0 = (cf32 *)malloc(M * sizeof(cf32));
    q->S1 = (cf32 *)malloc(M * sizeof(cf32)); q->s1 = (cf32 *)malloc(M * sizeof(cf32));
    ofdmframe_init_S0(q->p, M, q->S0, q->s0, &q->M_S0);
    ofdmframe_init_S1(q->p, M, q->S1, q->s1, &q->M_S1);
    q->g0 = 1.0f;
    q->G0 = (cf32 *)calloc(M, sizeof(cf32)); q->G1 = (cf32 *)calloc(M, sizeof(cf32));
    q->G = (cf32 *)calloc(M, sizeof(cf32)); q->B = (cf32 *)calloc(M, sizeof(cf32));
    q->R = (cf32 *)calloc(M, sizeof(cf32));
    q->backoff = cp_len < 2 ? cp_len : 2;
    float phi = (float)(q->backoff) * 2.0f * M_PI / (float)M;
    unsigned int i;
    for (i = 0; i < M; i++) {
        float a = (float)i * phi;
        q->B[i] = cosf(a) + _Complex_I * sinf(a);
    }
    q->nco_rx.type = LIQUID_VCO;
    orc_mseq_init_default(&q->ms_pilot, 8);
    q->callback = cb;
    q->userdata = userdata;
    q->sample_index = 0;
    q->detect_index = 0;
    ofdmframesync_reset(q);
    return q;
}

void ofdmframesync_destroy(ofdmframesync q)
{
    free(q->p); free(q->X); free(q->x); free(q->S0); free(q->s0); free(q->S1); free(q->s1);
    free(q->G0); free(q->G1); free(q->G); free(q->B); free(q->R);
    orc_window_free(&q->input_buffer);
    orc_fft_destroy(q->fft);
    free(q);
}

void ofdmframesync_reset(ofdmframesync q)
{
    q->nco_rx.theta = 0; q->nco_rx.d_theta = 0;
    orc_mseq_reset(&q->ms_pilot);
    q->timer = 0;
    q->num_symbols = 0;
    q->s_hat_0 = 0.0f; q->s_hat_1 = 0.0f;
    q->phi_prime = 0.0f; q->p1_prime = 0.0f;
    q->plateau_threshold = (q->M > 44) ? 0.35f : 0.35f + 0.01f * (float)(44 - q->M);
    q->state = ST_SEEKPLCP;
}

float ofdmframesync_get_rssi(ofdmframesync q) { return -10.0f * log10f(q->g0); }
float ofdmframesync_get_cfo(ofdmframesync q) { return nco_crcf_get_frequency(&q->nco_rx); }
uint64_t ofdmframesync_get_sample_index(ofdmframesync q) { return q->sample_index; }
uint64_t ofdmframesync_get_detect_index(ofdmframesync q) { return q->detect_index; }

static inline cf32 cmul(cf32 a, cf32 b)
{
    float ar = crealf(a), ai = cimagf(a), br = crealf(b), bi = cimagf(b);
    return (ar * br - ai * bi) + _Complex_I * (ar * bi + ai * br);
}
static inline cf32 cmulconj(cf32 a, cf32 b)     /* a * conj(b) */
{
    float ar = crealf(a), ai = cimagf(a), br = crealf(b), bi = cimagf(b);
    return (ar * br + ai * bi) + _Complex_I * (ai * br - ar * bi);
}

static void estimate_gain_S0(ofdmframesync q, const cf32 * x, cf32 * G)
{
    orc_fft_execute(q->fft, x, q->X);
    float gain = sqrtf((float)q->M_S0) / (float)q->M;
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] != OFDMFRAME_SCTYPE_NULL && (i % 2) == 0) G[i] = cmulconj(q->X[i], q->S0[i]) * gain;
        else G[i] = 0.0f;
    }
}

static cf32 S0_metrics(ofdmframesync q, const cf32 * G)
{
    float sr = 0.0f, si = 0.0f;
    unsigned int i;
    for (i = 0; i < q->M; i += 2) {
        cf32 t = cmulconj(G[(i + 2) % q->M], G[i]);
        sr += crealf(t); si += cimagf(t);
    }
    return (sr / (float)q->M_S0) + _Complex_I * (si / (float)q->M_S0);
}

static void estimate_gain_S1(ofdmframesync q, const cf32 * x, cf32 * G)
{
    orc_fft_execute(q->fft, x, q->X);
    float gain = sqrtf((float)q->M_S1) / (float)q->M;
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] != OFDMFRAME_SCTYPE_NULL) G[i] = cmulconj(q->X[i], q->S1[i]) * gain;
        else G[i] = 0.0f;
    }
}

static cf32 S1_metrics(ofdmframesync q, const cf32 * G)
{
    float sr = 0.0f, si = 0.0f;
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        cf32 t = cmulconj(G[(i + 1) % q->M], G[i]);
        sr += crealf(t); si += cimagf(t);
    }
    return (sr / (float)q->M_S1) + _Complex_I * (si / (float)q->M_S1);
}

static void execute_seekplcp(ofdmframesync q)
{
    q->timer++;
    if (q->timer < (int)q->M) return;
    q->timer = 0;
    const cf32 * rc = orc_window_read(&q->input_buffer);
    unsigned int i;
    float g = 0.0f;
    for (i = q->cp_len; i < q->M + q->cp_len; i++)
        g += crealf(rc[i]) * crealf(rc[i]) + cimagf(rc[i]) * cimagf(rc[i]);
    g = (float)q->M / g;
    estimate_gain_S0(q, &rc[q->cp_len], q->G0);
    cf32 s_hat = S0_metrics(q, q->G0) * g;
    float tau_hat = cargf(s_hat) * (float)q->M2 / (2 * (float)M_PI);
    q->g0 = g;
    if (cabsf(s_hat) > q->plateau_threshold) {
        int dt = (int)roundf(tau_hat);
        q->timer = (int)((q->M + dt) % q->M2);
        q->timer += (int)q->M;
        q->state = ST_PLCPSHORT0;
        q->detect_index = q->sample_index;
    }
}

static void execute_S0a(ofdmframesync q)
{
    q->timer++;
    if (q->timer < (int)q->M2) return;
    q->timer = 0;
    const cf32 * rc = orc_window_read(&q->input_buffer);
    estimate_gain_S0(q, &rc[q->cp_len], q->G0);
    q->s_hat_0 = S0_metrics(q, q->G0) * q->g0;
    q->state = ST_PLCPSHORT1;
}

static void execute_S0b(ofdmframesync q)
{
    q->timer++;
    if (q->timer < (int)q->M2) return;
    q->timer = (int)(q->M + q->cp_len - q->backoff);
    const cf32 * rc = orc_window_read(&q->input_buffer);
    estimate_gain_S0(q, &rc[q->cp_len], q->G1);
    q->s_hat_1 = S0_metrics(q, q->G1) * q->g0;
    float tau_hat = cargf(q->s_hat_0 + q->s_hat_1) * (float)q->M2 / (2 * (float)M_PI);
    q->timer -= (int)roundf(tau_hat);
    float gr = 0.0f, gi = 0.0f;
    unsigned int i;
    for (i = 0; i < q->M; i++) {
        cf32 t = cmulconj(q->G1[i], q->G0[i]);
        gr += crealf(t); gi += cimagf(t);
    }
    float nu_hat = 2.0f * atan2f(gi, gr) / (float)q->M;
    q->nco_rx.d_theta = orc_nco_constrain(nu_hat);
    q->state = ST_PLCPLONG;
}

static void estimate_eqgain_poly(ofdmframesync q, unsigned int order)
{
    unsigned int i, N = q->M_pilot + q->M_data;
    if (order > N - 1) order = N - 1;
    if (order > 10) order = 10;
    float * x_freq = (float *)malloc(N * sizeof(float));
    float * y_abs = (float *)malloc(N * sizeof(float));
    float * y_arg = (float *)malloc(N * sizeof(float));
    double p_abs[11], p_arg[11];
    unsigned int n = 0;
    for (i = 0; i < q->M; i++) {
        unsigned int k = (i + q->M2) % q->M;
        if (q->p[k] != OFDMFRAME_SCTYPE_NULL) {
            x_freq[n] = (k > q->M2) ? (float)k - (float)q->M : (float)k;
            x_freq[n] = x_freq[n] / (float)q->M;
            y_abs[n] = cabsf(q->G[k]);
            y_arg[n] = cargf(q->G[k]);
            n++;
        }
    }
    for (i = 1; i < N; i++) {
        while ((y_arg[i] - y_arg[i - 1]) >  (float)M_PI) y_arg[i] -= 2 * (float)M_PI;
        while ((y_arg[i] - y_arg[i - 1]) < -(float)M_PI) y_arg[i] += 2 * (float)M_PI;
    }
    orc_polyfit_d(x_freq, y_abs, N, p_abs, order + 1);
    orc_polyfit_d(x_freq, y_arg, N, p_arg, order + 1);
    for (i = 0; i < q->M; i++) {
        float freq = (i > q->M2) ? (float)i - (float)q->M : (float)i;
        freq = freq / (float)q->M;
        float A = (float)orc_polyval_d(p_abs, order + 1, (double)freq);
        float theta = (float)orc_polyval_d(p_arg, order + 1, (double)freq);
        q->G[i] = (q->p[i] == OFDMFRAME_SCTYPE_NULL) ? 0.0f : A * (cosf(theta) + _Complex_I * sinf(theta));
    }
    free(x_freq); free(y_abs); free(y_arg);
}

static void execute_S1(ofdmframesync q)
{
    q->timer--;
    if (q->timer > 0) return;
    q->num_symbols++;
    const cf32 * rc = orc_window_read(&q->input_buffer);
    estimate_gain_S1(q, &rc[q->cp_len], q->G);
    cf32 s_hat = S1_metrics(q, q->G) * q->g0;
    float a = (float)q->backoff * 2.0f * (float)M_PI / (float)q->M;
    s_hat = cmul(s_hat, cosf(a) + _Complex_I * sinf(a));
    if (cabsf(s_hat) > q->plateau_threshold && fabsf(cargf(s_hat)) < 0.1f * (float)M_PI) {
        q->state = ST_RXSYMBOLS;
        q->timer = (int)(q->M + q->cp_len + q->backoff);
        q->num_symbols = 0;
        float g = (float)q->M / sqrtf((float)(q->M_pilot + q->M_data));
        unsigned int i;
        for (i = 0; i < q->M; i++) q->G[i] = cmul(q->G[i] * g, q->B[i]);
        unsigned int poly_order = 4;
        if (poly_order >= q->M_pilot + q->M_data) poly_order = q->M_pilot + q->M_data - 1;
        estimate_eqgain_poly(q, poly_order);
        for (i = 0; i < q->M; i++) {
            if (q->p[i] == OFDMFRAME_SCTYPE_NULL) { q->R[i] = 0.0f; continue; }
            /* R = B / G */
            float gr = crealf(q->G[i]), gi = cimagf(q->G[i]);
            float d = gr * gr + gi * gi;
            cf32 num = cmulconj(q->B[i], q->G[i]);
            q->R[i] = (crealf(num) / d) + _Complex_I * (cimagf(num) / d);
        }
    } else if (q->num_symbols == 16) {
        ofdmframesync_reset(q);
    } else {
        q->timer = (int)q->M2;
    }
}

static void rxsymbol(ofdmframesync q)
{
    unsigned int i, n = 0;
    for (i = 0; i < q->M; i++) q->X[i] = cmul(q->X[i], q->R[i]);
    float x_phase[q->M_pilot], y_phase[q->M_pilot];
    for (i = 0; i < q->M; i++) {
        unsigned int k = (i + q->M2) % q->M;
        if (q->p[k] == OFDMFRAME_SCTYPE_PILOT) {
            float pilot = orc_mseq_advance(&q->ms_pilot) ? 1.0f : -1.0f;
            x_phase[n] = (k > q->M2) ? (float)k - (float)q->M : (float)k;
            y_phase[n] = atan2f(cimagf(q->X[k]) * pilot, crealf(q->X[k]) * pilot);
            n++;
        }
    }
    for (i = 1; i < q->M_pilot; i++) {
        while ((y_phase[i] - y_phase[i - 1]) >  (float)M_PI) y_phase[i] -= 2 * (float)M_PI;
        while ((y_phase[i] - y_phase[i - 1]) < -(float)M_PI) y_phase[i] += 2 * (float)M_PI;
    }
    /* first-order least squares, closed form (polyf_fit with 2 coefficients) */
    float sx = 0.0f, sy = 0.0f, sxx = 0.0f, sxy = 0.0f;
    for (i = 0; i < q->M_pilot; i++) {
        sx += x_phase[i]; sy += y_phase[i];
        sxx += x_phase[i] * x_phase[i]; sxy += x_phase[i] * y_phase[i];
    }
    float np = (float)q->M_pilot;
    float den = np * sxx - sx * sx;
    float p1 = (np * sxy - sx * sy) / den;
    float p0 = (sy - p1 * sx) / np;
    float alpha = 0.3f;
    p1 = alpha * p1 + (1 - alpha) * q->p1_prime;
    q->p1_prime = p1;
    for (i = 0; i < q->M; i++) {
        if (q->p[i] == OFDMFRAME_SCTYPE_NULL) {
            q->X[i] = 0.0f;
        } else {
            float fx = (i > q->M2) ? (float)i - (float)q->M : (float)i;
            float theta = p0 + p1 * fx;
            q->X[i] = cmul(q->X[i], cosf(theta) - _Complex_I * sinf(theta));
        }
    }
    if (q->num_symbols > 0) {
        float dphi = p0 - q->phi_prime;
        while (dphi >  (float)M_PI) dphi -= 2 * (float)M_PI;
        while (dphi < -(float)M_PI) dphi += 2 * (float)M_PI;
        q->nco_rx.d_theta += orc_nco_constrain(1e-3f * dphi);
    }
    q->phi_prime = p0;
    q->num_symbols++;
}

static void execute_rxsymbols(ofdmframesync q)
{
    q->timer--;
    if (q->timer == 0) {
        const cf32 * rc = orc_window_read(&q->input_buffer);
        memmove(q->x, &rc[q->cp_len - q->backoff], q->M * sizeof(cf32));
        orc_fft_execute(q->fft, q->x, q->X);
        rxsymbol(q);
        if (q->callback != NULL) {
            int retval = q->callback(q->X, q->p, q->M, q->userdata);
            if (retval != 0) ofdmframesync_reset(q);
        }
        /* liquid sets this unconditionally, also after a reset from inside the callback */
        q->timer = (int)(q->M + q->cp_len);
    }
}

void ofdmframesync_execute(ofdmframesync q, const cf32 * x_in, unsigned int n)
{
    unsigned int i;
    for (i = 0; i < n; i++) {
        cf32 x = x_in[i];
        if (q->state != ST_SEEKPLCP) {
            float s, c;
            orc_nco_sincos(q->nco_rx.theta, &s, &c);
            float xr = crealf(x), xi = cimagf(x);
            x = (xr * c + xi * s) + _Complex_I * (xi * c - xr * s);
            q->nco_rx.theta += q->nco_rx.d_theta;
        }
        orc_window_push(&q->input_buffer, x);
        switch (q->state) {
        case ST_SEEKPLCP:   execute_seekplcp(q);  break;
        case ST_PLCPSHORT0: execute_S0a(q);       break;
        case ST_PLCPSHORT1: execute_S0b(q);       break;
        case ST_PLCPLONG:   execute_S1(q);        break;
        case ST_RXSYMBOLS:  execute_rxsymbols(q); break;
        }
        q->sample_index++;
    }
}
