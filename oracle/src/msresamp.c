/*
 * msresamp.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * msresamp_crcf: multi-stage arbitrary resampler, liquid-dsp 1.3.x src/filter/src/msresamp.c
 * + resamp.c + firpfb.c.  Not used by the reference's lib/; usage pattern from
 * src/flexframe_rx.cc:179,240 and src/flexframe_tx.cc:170,237 (create(rate, As), execute per
 * buffer).  BASELINE config 4 puts it in front of ofdmflexframesync at rate 1.07.
 *
 * Only the arbitrary stage is restated (rates in [0.5, 2] need zero half-band stages).
 * Normative choice D7 (oracle/README.md): the output phase is a 32-bit fixed-point
 * accumulator (2^32 <-> one input sample, step = round(2^32 / rate)), so output k maps to
 * (input index, branch, mu) in closed form; branch npfb is branch 0 delayed by one sample,
 * read from the same prototype, so no output ever waits for the next input.
 */
#include "oracle_internal.h"

#define RS_M     7u      /* filter semi-length            */
#define RS_NPFB  64u     /* number of polyphase branches  */
#define RS_BITS  6u      /* log2(RS_NPFB)                 */

struct msresamp_crcf_s {
    float rate, As;
    unsigned int hlen;       /* 2*m*npfb + 1                       */
    float * h;               /* prototype, sum(h) = npfb           */
    orc_window w;            /* last 2m input samples              */
    uint64_t tau;            /* phase within current input, Q32    */
    uint64_t step;           /* round(2^32 / rate)                 */
};

msresamp_crcf msresamp_crcf_create(float rate, float As)
{
    if (!(rate >= 0.5f && rate <= 2.0f)) {
        fprintf(stderr, "error: msresamp_crcf_create(), oracle restates rates in [0.5,2] only (got %g)\n", rate);
        exit(1);
    }
    msresamp_crcf q = (msresamp_crcf)calloc(1, sizeof(struct msresamp_crcf_s));
    q->rate = rate;
    q->As = fabsf(As);
    q->hlen = 2 * RS_M * RS_NPFB + 1;
    q->h = (float *)malloc(q->hlen * sizeof(float));
    float fc = 0.515f * (rate < 1.0f ? rate : 1.0f);
    if (fc > 0.49f) fc = 0.49f;
    orc_firdes_kaiser(q->hlen, fc / (float)RS_NPFB, q->As, 0.0f, q->h);
    double sum = 0.0;
    unsigned int i;
    for (i = 0; i < q->hlen; i++) sum += (double)q->h[i];
    for (i = 0; i < q->hlen; i++) q->h[i] = (float)((double)q->h[i] * (double)RS_NPFB / sum);
    orc_window_init(&q->w, 2 * RS_M);
    q->step = (uint64_t)llrint(4294967296.0 / (double)rate);
    q->tau = 0;
    return q;
}

void msresamp_crcf_destroy(msresamp_crcf q)
{
    orc_window_free(&q->w);
    free(q->h);
    free(q);
}

void msresamp_crcf_reset(msresamp_crcf q)
{
    orc_window_clear(&q->w);
    q->tau = 0;
}

float msresamp_crcf_get_delay(msresamp_crcf q) { (void)q; return (float)RS_M; }

/* test hook: prototype taps and the Q32 step (the CUDA kernel is built from the same numbers) */
unsigned int orc_msresamp_get_design(msresamp_crcf q, const float ** h, uint64_t * step)
{
    *h = q->h; *step = q->step;
    return q->hlen;
}

void msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex * x, unsigned int nx,
                           liquid_float_complex * y, unsigned int * ny_out)
{
    unsigned int i, n, ny = 0;
    for (i = 0; i < nx; i++) {
        orc_window_push(&q->w, x[i]);
        const cf32 * r = orc_window_read(&q->w);      /* r[2m-1] newest */
        while (q->tau < 4294967296ull) {
            unsigned int f = (unsigned int)q->tau;
            unsigned int b = f >> (32 - RS_BITS);
            float mu = (float)(f & ((1u << (32 - RS_BITS)) - 1u)) * (1.0f / (float)(1u << (32 - RS_BITS)));
            float y0r = 0.0f, y0i = 0.0f, y1r = 0.0f, y1i = 0.0f;
            for (n = 0; n < 2 * RS_M; n++) {
                cf32 s = r[2 * RS_M - 1 - n];
                float h0 = q->h[b + n * RS_NPFB], h1 = q->h[b + 1 + n * RS_NPFB];
                y0r += h0 * crealf(s); y0i += h0 * cimagf(s);
                y1r += h1 * crealf(s); y1i += h1 * cimagf(s);
            }
            y[ny++] = ((1.0f - mu) * y0r + mu * y1r) + _Complex_I * ((1.0f - mu) * y0i + mu * y1i);
            q->tau += q->step;
        }
        q->tau -= 4294967296ull;
    }
    *ny_out = ny;
}
