/*
 * firpfbch.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * firpfbch_crcf: critically-sampled polyphase filterbank channelizer, liquid-dsp 1.3.x
 * src/multichannel/src/firpfbch.c.  Reference call sites: lib/multichannelrx.cc:91,142,188
 * (analyzer, K = 2N, m = 7) and lib/multichanneltx.cc:87,133,213 (synthesizer, m = 13).
 */
#include "oracle_internal.h"

struct firpfbch_crcf_s {
    int type;
    unsigned int K;          /* number of channels            */
    unsigned int p;          /* taps per polyphase branch     */
    float * hsub;            /* [K][p] reversed branch taps   */
    orc_window * w;          /* [K] branch windows            */
    cf32 * X, * x;           /* fft in / out                  */
    orc_fft * fft;
    unsigned int filter_index;
};

firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned int K, unsigned int m, float As)
{
    if (K == 0 || m == 0) {
        fprintf(stderr, "error: firpfbch_crcf_create_kaiser(), invalid configuration\n");
        exit(1);
    }
    As = fabsf(As);
    unsigned int h_len = 2 * K * m + 1;
    float * h = (float *)malloc(h_len * sizeof(float));
    float fc = 0.5f / (float)K;
    orc_firdes_kaiser(h_len, fc, As, 0.0f, h);

    firpfbch_crcf q = (firpfbch_crcf)calloc(1, sizeof(struct firpfbch_crcf_s));
    q->type = type;
    q->K = K;
    q->p = 2 * m;
    q->hsub = (float *)malloc(K * q->p * sizeof(float));
    q->w = (orc_window *)malloc(K * sizeof(orc_window));
    unsigned int i, n;
    for (i = 0; i < K; i++) {
        for (n = 0; n < q->p; n++)
            q->hsub[i * q->p + (q->p - n - 1)] = h[i + n * K];
        orc_window_init(&q->w[i], q->p);
    }
    free(h);
    q->X = (cf32 *)malloc(K * sizeof(cf32));
    q->x = (cf32 *)malloc(K * sizeof(cf32));
    q->fft = orc_fft_create(K, type == LIQUID_ANALYZER ? ORC_FFT_FORWARD : ORC_FFT_BACKWARD);
    firpfbch_crcf_reset(q);
    return q;
}

void firpfbch_crcf_destroy(firpfbch_crcf q)
{
    unsigned int i;
    for (i = 0; i < q->K; i++) orc_window_free(&q->w[i]);
    free(q->w); free(q->hsub); free(q->X); free(q->x);
    orc_fft_destroy(q->fft);
    free(q);
}

void firpfbch_crcf_reset(firpfbch_crcf q)
{
    unsigned int i;
    for (i = 0; i < q->K; i++) orc_window_clear(&q->w[i]);
    q->filter_index = q->K - 1;
}

static inline cf32 dotprod_crcf(const float * h, const cf32 * r, unsigned int p)
{
    float ar = 0.0f, ai = 0.0f;
    unsigned int i;
    for (i = 0; i < p; i++) { ar += h[i] * crealf(r[i]); ai += h[i] * cimagf(r[i]); }
    return ar + _Complex_I * ai;
}

void firpfbch_crcf_synthesizer_execute(firpfbch_crcf q, liquid_float_complex * X, liquid_float_complex * y)
{
    unsigned int i;
    memmove(q->X, X, q->K * sizeof(cf32));
    orc_fft_execute(q->fft, q->X, q->x);
    for (i = 0; i < q->K; i++) {
        orc_window_push(&q->w[i], q->x[i]);
        y[i] = dotprod_crcf(&q->hsub[i * q->p], orc_window_read(&q->w[i]), q->p);
    }
}

void firpfbch_crcf_analyzer_execute(firpfbch_crcf q, liquid_float_complex * x, liquid_float_complex * y)
{
    unsigned int i;
    for (i = 0; i < q->K; i++) {
        orc_window_push(&q->w[q->filter_index], x[i]);
        q->filter_index = (q->filter_index + q->K - 1) % q->K;
    }
    for (i = 0; i < q->K; i++)
        q->X[q->K - i - 1] = dotprod_crcf(&q->hsub[i * q->p], orc_window_read(&q->w[i]), q->p);
    orc_fft_execute(q->fft, q->X, q->x);
    memmove(y, q->x, q->K * sizeof(cf32));
}
