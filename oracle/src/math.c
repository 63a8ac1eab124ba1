/*
 * math.c -- CPU ORACLE (test infrastructure only; see oracle_internal.h).
 *
 * FFT, Kaiser filter design, polynomial fit, m-sequence, sample window, NCO.
 * Follows liquid-dsp 1.3.x: src/fft/src/fft_radix2.c + fft_dft.c, src/filter/src/firdes.c
 * (liquid_firdes_kaiser, kaiser_beta_As), src/math/src/windows.c (kaiser), math.bessel.c
 * (besseli0f), math.c (sincf), src/sequence/src/msequence.c, src/buffer/src/window.c,
 * src/nco/src/nco.c (uint32 phase accumulator of 1.3.2).
 * Call sites in the reference: lib/multichannelrx.cc:91,98-100,163-164;
 * lib/multichanneltx.cc:87,94-96,219-222.
 */
#include "oracle_internal.h"

const char * liquid_libversion(void) { return LIQUID_VERSION; }

/* ------------------------------------------------------------------- fft */
orc_fft * orc_fft_create(unsigned int n, int dir)
{
    orc_fft * q = (orc_fft *)calloc(1, sizeof(orc_fft));
    q->n = n;
    q->dir = dir;
    q->pow2 = (n & (n - 1)) == 0;
    q->tw = (cf32 *)malloc(n * sizeof(cf32));
    unsigned int i;
    for (i = 0; i < n; i++) {
        double a = (double)dir * 2.0 * M_PI * (double)i / (double)n;
        q->tw[i] = (float)cos(a) + _Complex_I * (float)sin(a);
    }
    if (q->pow2) {
        unsigned int m = 0;
        while ((1u << m) < n) m++;
        q->log2n = m;
        q->brev = (unsigned int *)malloc(n * sizeof(unsigned int));
        for (i = 0; i < n; i++) {
            unsigned int r = 0, b;
            for (b = 0; b < m; b++) if (i & (1u << b)) r |= 1u << (m - 1 - b);
            q->brev[i] = r;
        }
    }
    return q;
}

void orc_fft_destroy(orc_fft * q)
{
    if (!q) return;
    free(q->tw);
    free(q->brev);
    free(q);
}

/* radix-2 decimation-in-time for n = 2^m (liquid: fft_execute_radix2), plain DFT otherwise */
void orc_fft_execute(const orc_fft * q, const cf32 * x, cf32 * y)
{
    unsigned int n = q->n, i, k;
    if (!q->pow2) {
        for (k = 0; k < n; k++) {
            float ar = 0.0f, ai = 0.0f;
            unsigned int idx = 0;
            for (i = 0; i < n; i++) {
                float wr = crealf(q->tw[idx]), wi = cimagf(q->tw[idx]);
                float xr = crealf(x[i]), xi = cimagf(x[i]);
                ar += xr * wr - xi * wi;
                ai += xr * wi + xi * wr;
                idx += k; if (idx >= n) idx -= n;
            }
            y[k] = ar + _Complex_I * ai;
        }
        return;
    }
    for (i = 0; i < n; i++) y[i] = x[q->brev[i]];
    unsigned int half = 1, stride = n >> 1;
    while (half < n) {
        for (k = 0; k < n; k += 2 * half) {
            unsigned int j;
            for (j = 0; j < half; j++) {
                cf32 w = q->tw[j * stride];
                float wr = crealf(w), wi = cimagf(w);
                float br = crealf(y[k + j + half]), bi = cimagf(y[k + j + half]);
                float tr = br * wr - bi * wi;
                float ti = br * wi + bi * wr;
                float ar = crealf(y[k + j]), ai = cimagf(y[k + j]);
                y[k + j]        = (ar + tr) + _Complex_I * (ai + ti);
                y[k + j + half] = (ar - tr) + _Complex_I * (ai - ti);
            }
        }
        half <<= 1;
        stride >>= 1;
    }
}

/* ---------------------------------------------------------- filter design */
#define NUM_BESSELI0_ITERATIONS 32
float orc_besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    unsigned int k;
    float t, y = 0.0f;
    for (k = 0; k < NUM_BESSELI0_ITERATIONS; k++) {
        t = (float)k * logf(0.5f * z) - lgammaf((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}

float orc_kaiser(unsigned int i, unsigned int n, float beta)
{
    float t = (float)i - (float)(n - 1) / 2;
    float r = 2.0f * t / (float)n;
    float a = orc_besseli0f(beta * sqrtf(1 - r * r));
    float b = orc_besseli0f(beta);
    return a / b;
}

float orc_sincf(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
    return sinf(M_PI * x) / (M_PI * x);
}

float orc_kaiser_beta_As(float As)
{
    As = fabsf(As);
    float beta;
    if (As > 50.0f)       beta = 0.1102f * (As - 8.7f);
    else if (As > 21.0f)  beta = 0.5842 * powf(As - 21, 0.4f) + 0.07886f * (As - 21);
    else                  beta = 0.0f;
    return beta;
}

void orc_firdes_kaiser(unsigned int n, float fc, float As, float mu, float * h)
{
    float beta = orc_kaiser_beta_As(As);
    unsigned int i;
    for (i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        float h1 = orc_sincf(2.0f * fc * t);
        float h2 = orc_kaiser(i, n, beta);
        h[i] = h1 * h2;
    }
}

/* normal equations + Gaussian elimination with partial pivoting, all in double */
void orc_polyfit_d(const float * x, const float * y, unsigned int n, double * p, unsigned int k)
{
    double A[11][12];
    unsigned int i, r, c;
    if (k > 11) k = 11;
    for (r = 0; r < k; r++) for (c = 0; c <= k; c++) A[r][c] = 0.0;
    for (i = 0; i < n; i++) {
        double xp[22];
        double xv = (double)x[i];
        xp[0] = 1.0;
        for (r = 1; r < 2 * k; r++) xp[r] = xp[r - 1] * xv;
        for (r = 0; r < k; r++) {
            for (c = 0; c < k; c++) A[r][c] += xp[r + c];
            A[r][k] += xp[r] * (double)y[i];
        }
    }
    for (c = 0; c < k; c++) {
        unsigned int piv = c;
        for (r = c + 1; r < k; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (piv != c) for (i = 0; i <= k; i++) { double t = A[c][i]; A[c][i] = A[piv][i]; A[piv][i] = t; }
        for (r = c + 1; r < k; r++) {
            double f = A[r][c] / A[c][c];
            for (i = c; i <= k; i++) A[r][i] -= f * A[c][i];
        }
    }
    for (r = k; r-- > 0;) {
        double s = A[r][k];
        for (c = r + 1; c < k; c++) s -= A[r][c] * p[c];
        p[r] = s / A[r][r];
    }
}

double orc_polyval_d(const double * p, unsigned int k, double x)
{
    double v = 0.0, xp = 1.0;
    unsigned int i;
    for (i = 0; i < k; i++) { v += p[i] * xp; xp *= x; }
    return v;
}

/* -------------------------------------------------------------- msequence */
static const unsigned int orc_mseq_genpoly[16] = {
    0, 0, 0x0007, 0x000B, 0x0013, 0x0025, 0x0043, 0x0089,
    0x011D, 0x0211, 0x0409, 0x0805, 0x1053, 0x201b, 0x402b, 0x8003};

static unsigned int orc_parity(unsigned int v)
{
    v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
    return v & 1u;
}

void orc_mseq_init_default(orc_mseq * ms, unsigned int m)
{
    ms->m = m;
    ms->g = orc_mseq_genpoly[m] >> 1;
    ms->a = 1;
    ms->n = (1u << m) - 1;
    ms->v = ms->a;
    ms->b = 0;
}

void orc_mseq_reset(orc_mseq * ms) { ms->v = ms->a; ms->b = 0; }

unsigned int orc_mseq_advance(orc_mseq * ms)
{
    ms->b = orc_parity(ms->v & ms->g);
    ms->v <<= 1;
    ms->v |= ms->b;
    ms->v &= ms->n;
    return ms->b;
}

unsigned int orc_mseq_symbol(orc_mseq * ms, unsigned int bps)
{
    unsigned int i, s = 0;
    for (i = 0; i < bps; i++) { s <<= 1; s |= orc_mseq_advance(ms); }
    return s;
}

/* ----------------------------------------------------------------- window */
void orc_window_init(orc_window * w, unsigned int n)
{
    w->n = n;
    w->cap = 2 * n + 16;
    w->pos = 0;
    w->buf = (cf32 *)calloc(w->cap, sizeof(cf32));
}
void orc_window_free(orc_window * w) { free(w->buf); w->buf = NULL; }
void orc_window_clear(orc_window * w) { memset(w->buf, 0, w->cap * sizeof(cf32)); w->pos = 0; }

/* -------------------------------------------------------------------- nco */
/* phase/frequency in radians -> 2^32 fixed point.  liquid 1.3.2 does this product in float;
 * the oracle does it in double with round-to-nearest so that it is exactly reproducible on
 * any IEEE machine (deviation D1 in oracle/README.md). */
uint32_t orc_nco_constrain(float theta)
{
    double p = (double)theta * 0.15915494309189535;   /* 1/(2*pi) */
    double f = p - floor(p);
    double u = rint(f * 4294967296.0);
    return (uint32_t)((uint64_t)u & 0xffffffffu);
}

nco_crcf nco_crcf_create(liquid_ncotype type)
{
    nco_crcf q = (nco_crcf)calloc(1, sizeof(struct nco_crcf_s));
    q->type = type;
    return q;
}
void nco_crcf_destroy(nco_crcf q) { free(q); }
void nco_crcf_reset(nco_crcf q) { q->theta = 0; q->d_theta = 0; }
void nco_crcf_set_frequency(nco_crcf q, float dtheta) { q->d_theta = orc_nco_constrain(dtheta); }
void nco_crcf_adjust_frequency(nco_crcf q, float step) { q->d_theta += orc_nco_constrain(step); }
float nco_crcf_get_frequency(nco_crcf q)
{
    return (float)((double)(int32_t)q->d_theta * (M_PI / 2147483648.0));
}
void nco_crcf_set_phase(nco_crcf q, float theta) { q->theta = orc_nco_constrain(theta); }
float nco_crcf_get_phase(nco_crcf q)
{
    return (float)((double)(int32_t)q->theta * (M_PI / 2147483648.0));
}
void nco_crcf_step(nco_crcf q) { q->theta += q->d_theta; }

void nco_crcf_mix_up(nco_crcf q, liquid_float_complex x, liquid_float_complex * y)
{
    float s, c;
    orc_nco_sincos(q->theta, &s, &c);
    float xr = crealf(x), xi = cimagf(x);
    *y = (xr * c - xi * s) + _Complex_I * (xr * s + xi * c);
}
void nco_crcf_mix_down(nco_crcf q, liquid_float_complex x, liquid_float_complex * y)
{
    float s, c;
    orc_nco_sincos(q->theta, &s, &c);
    float xr = crealf(x), xi = cimagf(x);
    *y = (xr * c + xi * s) + _Complex_I * (xi * c - xr * s);
}
void nco_crcf_mix_block_up(nco_crcf q, liquid_float_complex * x, liquid_float_complex * y, unsigned int n)
{
    unsigned int i;
    for (i = 0; i < n; i++) { nco_crcf_mix_up(q, x[i], &y[i]); nco_crcf_step(q); }
}
void nco_crcf_mix_block_down(nco_crcf q, liquid_float_complex * x, liquid_float_complex * y, unsigned int n)
{
    unsigned int i;
    for (i = 0; i < n; i++) { nco_crcf_mix_down(q, x[i], &y[i]); nco_crcf_step(q); }
}
