// design.h -- host-side (one-off, construction-time) design of everything the kernels read as
// tables: Kaiser prototype for firpfbch_crcf, default OFDM subcarrier allocation, S0/S1 training
// sequences, pilot m-sequence, mixed-radix FFT plans.  These are the parameter computations
// liquid-dsp performs inside the *_create() calls the reference issues at
// lib/multichannelrx.cc:82,91,99-100 and lib/multichanneltx.cc:80,87,95-96.
//
// Plain C++ (no CUDA); evaluated in float with the same formulas liquid uses so that the tables
// agree with a liquid-dsp build to the last bit where libm agrees.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace b2 {

// ---------------------------------------------------------------- subcarrier types / enums
enum { SC_NULL = 0, SC_PILOT = 1, SC_DATA = 2 };
enum { CRC_UNKNOWN = 0, CRC_NONE = 1, CRC_32 = 6 };
enum { FEC_UNKNOWN = 0, FEC_NONE = 1, FEC_HAMMING128 = 6, FEC_GOLAY2412 = 7, FEC_CONV_V27 = 11 };
enum { MOD_QAM4 = 25, MOD_QAM16 = 27, MOD_QAM64 = 29, MOD_QAM256 = 31, MOD_BPSK = 39, MOD_QPSK = 40 };

static inline int mod_bps(unsigned int scheme)
{
    switch (scheme) {
    case MOD_BPSK: return 1;
    case MOD_QPSK: case MOD_QAM4: return 2;
    case MOD_QAM16: return 4;
    case MOD_QAM64: return 6;
    case MOD_QAM256: return 8;
    default: return 0;
    }
}
static inline bool fec_supported(unsigned int s)
{
    return s == FEC_NONE || s == FEC_HAMMING128 || s == FEC_GOLAY2412 || s == FEC_CONV_V27;
}
static inline unsigned int fec_enc_len(unsigned int scheme, unsigned int n)
{
    switch (scheme) {
    case FEC_HAMMING128: return (n * 12 + 7) / 8;                 // 8 -> 12 bits per byte
    case FEC_GOLAY2412: { unsigned int blocks = (n * 8 + 11) / 12; return (blocks * 24 + 7) / 8; }
    case FEC_CONV_V27:  return (2 * (8 * n + 6) + 7) / 8;
    default: return n;
    }
}
static inline unsigned int packet_enc_len(unsigned int n, unsigned int check, unsigned int fec0, unsigned int fec1)
{
    return fec_enc_len(fec1, fec_enc_len(fec0, n + (check == CRC_32 ? 4 : 0)));
}

// ---------------------------------------------------------------- NCO fixed point (2*pi <-> 2^32)
static inline uint32_t nco_constrain(float theta)
{
    double p = (double)theta * 0.15915494309189535;
    double f = p - std::floor(p);
    double u = std::rint(f * 4294967296.0);
    return (uint32_t)((uint64_t)u & 0xffffffffu);
}

// ---------------------------------------------------------------- Kaiser prototype
static inline float besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    for (unsigned int k = 0; k < 32; k++) {
        float t = (float)k * logf(0.5f * z) - lgammaf((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}
static inline float kaiser_beta(float As)
{
    As = fabsf(As);
    if (As > 50.0f) return 0.1102f * (As - 8.7f);
    if (As > 21.0f) return 0.5842 * powf(As - 21, 0.4f) + 0.07886f * (As - 21);
    return 0.0f;
}
static inline float sincf_(float x)
{
    if (fabsf(x) < 0.01f) return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
    return sinf(M_PI * x) / (M_PI * x);
}
static inline std::vector<float> firdes_kaiser(unsigned int n, float fc, float As, float mu = 0.0f)
{
    std::vector<float> h(n);
    float beta = kaiser_beta(As);
    float ib = besseli0f(beta);
    for (unsigned int i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        float r = 2.0f * ((float)i - (float)(n - 1) / 2) / (float)n;
        float w = besseli0f(beta * sqrtf(1 - r * r)) / ib;
        h[i] = sincf_(2.0f * fc * t) * w;
    }
    return h;
}
// firpfbch_crcf_create_kaiser(type, K, m, As): prototype of 2*K*m+1 taps, fc = 0.5/K, the first
// K*2m taps split over K branches.  Returned as taps[n*K + i] = h[i + n*K] (n < 2m, i < K).
static inline std::vector<float> firpfbch_prototype(unsigned int K, unsigned int m, float As)
{
    std::vector<float> h = firdes_kaiser(2 * K * m + 1, 0.5f / (float)K, As);
    h.resize((size_t)2 * K * m);
    return h;
}

// ---------------------------------------------------------------- m-sequence (Fibonacci LFSR)
struct MSeq {
    unsigned int m, g, n, v;
    explicit MSeq(unsigned int m_)
    {
        static const unsigned int genpoly[16] = {0, 0, 0x0007, 0x000B, 0x0013, 0x0025, 0x0043, 0x0089,
                                                 0x011D, 0x0211, 0x0409, 0x0805, 0x1053, 0x201b, 0x402b, 0x8003};
        m = m_; g = genpoly[m] >> 1; n = (1u << m) - 1; v = 1;
    }
    unsigned int advance()
    {
        unsigned int b = __builtin_parity(v & g);
        v = ((v << 1) | b) & n;
        return b;
    }
    unsigned int symbol(unsigned int bps)
    {
        unsigned int s = 0;
        for (unsigned int i = 0; i < bps; i++) s = (s << 1) | advance();
        return s;
    }
};

// ---------------------------------------------------------------- OFDM frame structure
struct OfdmPlan {
    unsigned int M = 0, cp = 0, taper = 0, M2 = 0, backoff = 0;
    std::vector<uint8_t> p;               // subcarrier allocation
    unsigned int M_null = 0, M_pilot = 0, M_data = 0, M_S0 = 0, M_S1 = 0;
    std::vector<float> S0, S1;            // +-1 / 0 per subcarrier (frequency domain)
    std::vector<uint16_t> data_idx;       // data subcarriers, ascending index
    std::vector<uint16_t> pilot_idx;      // pilot subcarriers in fft-shifted visiting order
    std::vector<float> pilot_x;           // signed subcarrier index of each pilot (same order)
    std::vector<uint16_t> active_idx;     // non-null subcarriers in fft-shifted visiting order
    std::vector<uint8_t> pilot_seq;       // one period (255) of the pilot LFSR, bit per advance
    float pilot_sx = 0, pilot_sxx = 0;    // sum x, sum x^2 accumulated in visiting order (float)
    float thresh = 0.35f;
    unsigned int n_header_syms = 0;       // OFDM symbols carrying the 288 header BPSK symbols
};

static inline int ofdm_default_alloc(unsigned int M, uint8_t * p)
{
    unsigned int M2 = M / 2, G = M / 10;
    if (G < 2) G = 2;
    unsigned int P = (M > 34) ? 8 : 4, P2 = P / 2;
    memset(p, SC_NULL, M);
    for (unsigned int i = 1; i < M2 - G; i++) {
        uint8_t t = (((i + P2) % P) == 0) ? SC_PILOT : SC_DATA;
        p[i] = t; p[M - i] = t;
    }
    return 0;
}

static inline unsigned int ceil_log2(unsigned int x)
{
    unsigned int n = 0;
    x--;
    while (x > 0) { x >>= 1; n++; }
    return n;
}

// returns 0 on success, -1 on an allocation liquid would reject
static inline int ofdm_plan(OfdmPlan & o, unsigned int M, unsigned int cp, unsigned int taper, const unsigned char * p)
{
    o.M = M; o.cp = cp; o.taper = taper; o.M2 = M / 2; o.backoff = cp < 2 ? cp : 2;
    o.p.resize(M);
    if (p) memcpy(o.p.data(), p, M); else ofdm_default_alloc(M, o.p.data());
    o.M_null = o.M_pilot = o.M_data = 0;
    for (unsigned int i = 0; i < M; i++) {
        if (o.p[i] == SC_NULL) o.M_null++;
        else if (o.p[i] == SC_PILOT) o.M_pilot++;
        else if (o.p[i] == SC_DATA) o.M_data++;
        else return -1;
    }
    if (o.M_data == 0 || o.M_pilot < 2) return -1;
    unsigned int m = ceil_log2(M);
    if (m < 4) m = 4; else if (m > 8) m = 8;
    o.S0.assign(M, 0.0f); o.S1.assign(M, 0.0f);
    o.M_S0 = o.M_S1 = 0;
    {
        MSeq ms(m);
        for (unsigned int i = 0; i < M; i++) {
            unsigned int s = ms.symbol(3) & 1;
            if (o.p[i] != SC_NULL && (i % 2) == 0) { o.S0[i] = s ? 1.0f : -1.0f; o.M_S0++; }
        }
    }
    {
        MSeq ms(m + 1);
        for (unsigned int i = 0; i < M; i++) {
            unsigned int s = ms.symbol(3) & 1;
            if (o.p[i] != SC_NULL) { o.S1[i] = s ? 1.0f : -1.0f; o.M_S1++; }
        }
    }
    if (o.M_S0 == 0 || o.M_S1 == 0) return -1;
    o.data_idx.clear(); o.pilot_idx.clear(); o.pilot_x.clear(); o.active_idx.clear();
    for (unsigned int i = 0; i < M; i++) if (o.p[i] == SC_DATA) o.data_idx.push_back((uint16_t)i);
    o.pilot_sx = o.pilot_sxx = 0.0f;
    for (unsigned int i = 0; i < M; i++) {
        unsigned int k = (i + o.M2) % M;
        if (o.p[k] != SC_NULL) o.active_idx.push_back((uint16_t)k);
        if (o.p[k] == SC_PILOT) {
            float x = (k > o.M2) ? (float)k - (float)M : (float)k;
            o.pilot_idx.push_back((uint16_t)k);
            o.pilot_x.push_back(x);
            o.pilot_sx += x;
            o.pilot_sxx += x * x;
        }
    }
    o.pilot_seq.resize(255);
    MSeq mp(8);
    for (unsigned int i = 0; i < 255; i++) o.pilot_seq[i] = (uint8_t)mp.advance();
    o.thresh = (M > 44) ? 0.35f : 0.35f + 0.01f * (float)(44 - M);
    o.n_header_syms = (288 + o.M_data - 1) / o.M_data;
    return 0;
}

// ---------------------------------------------------------------- mixed-radix in-place DIT FFT plan
// Passes are applied in order; pass t has radix R[t] and works on sub-transforms of length
// L[t] = R[0]*...*R[t].  Input element n is loaded to position perm[n]; output is in natural order.
struct FftPlan {
    unsigned int n = 0, npass = 0;
    unsigned int radix[12];
    std::vector<uint16_t> perm;
    std::vector<float> tw;                // 2*n floats: cos, -sin (forward e^{-j 2 pi k/n}), from double
};

static inline unsigned int fft_inpos(unsigned int idx, unsigned int L, const unsigned int * radix, int t)
{
    if (t < 0 || L == 1) return 0;
    unsigned int R = radix[t], s = L / R;
    return s * (idx % R) + fft_inpos(idx / R, s, radix, t - 1);
}

// n >= 2 with no prime factor above 41 (liquid's FFT takes any size; the reference programs default
// to M = 48 subcarriers).  The power-of-two part is split into radix-2/4/8 passes, every odd
// prime factor p is a radix-p pass of its own (a p x p DFT per butterfly), at most 8 passes.
static inline int fft_plan(FftPlan & f, unsigned int n)
{
    if (n < 2) return -1;
    f.n = n; f.npass = 0;
    unsigned int odd = n, lg = 0;
    while ((odd & 1u) == 0) { odd >>= 1; lg++; }
    // odd prime factors first (they work on the shortest strides)
    for (unsigned int p = 3; p <= 41 && odd > 1; p += 2) {
        while (odd % p == 0) {
            if (f.npass >= 8) return -1;
            f.radix[f.npass++] = p;
            odd /= p;
        }
    }
    if (odd != 1) return -1;
    unsigned int rem = lg;
    // experiment knob: B2_FFT_MAX_RADIX=4 builds the plan from radix-4 (and one radix-2) passes
    if (const char * e = getenv("B2_FFT_MAX_RADIX")) {
        if (atoi(e) == 4 && (n & (n - 1)) == 0) {
            if (rem & 1) { f.radix[f.npass++] = 2; rem -= 1; }
            while (rem >= 2) { f.radix[f.npass++] = 4; rem -= 2; }
        }
    }
    // small radices first (cheap passes on short strides), radix-8 for the rest
    if (rem == 0) { }
    else if (rem % 3 == 1 && rem >= 4) { f.radix[f.npass++] = 4; f.radix[f.npass++] = 4; rem -= 4; }
    else if (rem % 3 == 1) { f.radix[f.npass++] = 2; rem -= 1; }
    else if (rem % 3 == 2) { f.radix[f.npass++] = 4; rem -= 2; }
    while (rem >= 3) { f.radix[f.npass++] = 8; rem -= 3; }
    if (f.npass > 8) return -1;
    f.perm.resize(n);
    for (unsigned int i = 0; i < n; i++) f.perm[i] = (uint16_t)fft_inpos(i, n, f.radix, (int)f.npass - 1);
    f.tw.resize(2 * (size_t)n);
    for (unsigned int k = 0; k < n; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        f.tw[2 * k] = (float)cos(a);
        f.tw[2 * k + 1] = (float)sin(a);
    }
    return 0;
}

// ---------------------------------------------------------------- S1 equaliser-gain fit
// liquid's ofdmframesync_estimate_eqgain_poly fits order-4 polynomials to |G| and arg G over the
// active subcarriers (abscissa = signed subcarrier index / M).  The abscissae are fixed by the
// allocation, so the least-squares solution is a constant matrix applied to the ordinates:
// coef = P y with P = (X^T X)^-1 X^T, computed here once in long double and returned as
// P[n*5 + r] (one 40-byte row per active subcarrier, fft-shifted visiting order).
static inline std::vector<double> eqgain_fit_matrix(const OfdmPlan & o)
{
    const size_t Na = o.active_idx.size();
    std::vector<double> x(Na);
    for (size_t n = 0; n < Na; n++) {
        unsigned int k = o.active_idx[n];
        float xf = (k > o.M2) ? (float)k - (float)o.M : (float)k;
        x[n] = (double)(xf / (float)o.M);
    }
    long double A[5][10];
    for (int r = 0; r < 5; r++)
        for (int c = 0; c < 10; c++) A[r][c] = (c >= 5 && c - 5 == r) ? 1.0L : 0.0L;
    for (size_t n = 0; n < Na; n++) {
        long double pw[9], xp = 1.0L;
        for (int r = 0; r < 9; r++) { pw[r] = xp; xp *= (long double)x[n]; }
        for (int r = 0; r < 5; r++)
            for (int c = 0; c < 5; c++) A[r][c] += pw[r + c];
    }
    for (int c = 0; c < 5; c++) {                       // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < 5; r++) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        if (piv != c) for (int i = 0; i < 10; i++) std::swap(A[c][i], A[piv][i]);
        long double d = A[c][c];
        if (d == 0.0L) d = 1e-300L;
        for (int i = 0; i < 10; i++) A[c][i] /= d;
        for (int r = 0; r < 5; r++) {
            if (r == c) continue;
            long double f = A[r][c];
            for (int i = 0; i < 10; i++) A[r][i] -= f * A[c][i];
        }
    }
    std::vector<double> P(5 * Na);
    for (size_t n = 0; n < Na; n++) {
        long double pw[5], xp = 1.0L;
        for (int r = 0; r < 5; r++) { pw[r] = xp; xp *= (long double)x[n]; }
        for (int r = 0; r < 5; r++) {
            long double a = 0.0L;
            for (int c = 0; c < 5; c++) a += A[r][5 + c] * pw[c];
            P[n * 5 + (size_t)r] = (double)a;
        }
    }
    return P;
}

// ---------------------------------------------------------------- transmit-side tables
// unnormalised radix-2 transform in float (host, construction time only)
static inline void host_fft(std::vector<float> & re, std::vector<float> & im, int dir)
{
    const unsigned int n = (unsigned int)re.size();
    if (n & (n - 1)) {
        // not a power of two (M = 48, ...): plain DFT, accumulated in double (construction time only)
        std::vector<float> xr(re), xi(im);
        for (unsigned int k = 0; k < n; k++) {
            double sr = 0.0, si = 0.0;
            for (unsigned int i = 0; i < n; i++) {
                double a = (double)dir * 2.0 * M_PI * (double)(((unsigned long long)k * i) % n) / (double)n;
                double c = cos(a), s = sin(a);
                sr += (double)xr[i] * c - (double)xi[i] * s;
                si += (double)xr[i] * s + (double)xi[i] * c;
            }
            re[k] = (float)sr; im[k] = (float)si;
        }
        return;
    }
    unsigned int lg = ceil_log2(n);
    for (unsigned int i = 0; i < n; i++) {
        unsigned int r = 0;
        for (unsigned int b = 0; b < lg; b++) if (i & (1u << b)) r |= 1u << (lg - 1 - b);
        if (r > i) { std::swap(re[i], re[r]); std::swap(im[i], im[r]); }
    }
    for (unsigned int half = 1; half < n; half <<= 1) {
        unsigned int stride = n / (2 * half);
        for (unsigned int k = 0; k < n; k += 2 * half) {
            for (unsigned int j = 0; j < half; j++) {
                double a = (double)dir * 2.0 * M_PI * (double)(j * stride) / (double)n;
                float wr = (float)cos(a), wi = (float)sin(a);
                float br = re[k + j + half], bi = im[k + j + half];
                float tr = br * wr - bi * wi, ti = br * wi + bi * wr;
                float ar = re[k + j], ai = im[k + j];
                re[k + j] = ar + tr; im[k + j] = ai + ti;
                re[k + j + half] = ar - tr; im[k + j + half] = ai - ti;
            }
        }
    }
}
// time-domain training symbols s0, s1 (IFFT of S0/S1 scaled by 1/sqrt(M_S)), interleaved re,im
static inline void ofdm_training_time(const OfdmPlan & o, std::vector<float> & s0, std::vector<float> & s1)
{
    for (int which = 0; which < 2; which++) {
        const std::vector<float> & S = which ? o.S1 : o.S0;
        std::vector<float> re(S), im(o.M, 0.0f);
        host_fft(re, im, +1);
        float g = 1.0f / sqrtf((float)(which ? o.M_S1 : o.M_S0));
        std::vector<float> & out = which ? s1 : s0;
        out.resize(2 * (size_t)o.M);
        for (unsigned int i = 0; i < o.M; i++) { out[2 * i] = re[i] * g; out[2 * i + 1] = im[i] * g; }
    }
}
// raised-sine^2 taper (ofdmframegen)
static inline std::vector<float> ofdm_taper(unsigned int taper)
{
    std::vector<float> w(taper ? taper : 1, 0.0f);
    for (unsigned int i = 0; i < taper; i++) {
        float t = ((float)i + 0.5f) / (float)taper;
        float g = sinf(M_PI_2 * t);
        w[i] = g * g;
    }
    return w;
}
// msresamp_crcf arbitrary stage: prototype (sum = npfb) and Q32 phase step
static inline std::vector<float> resamp_prototype(float rate, float As, unsigned int m, unsigned int npfb)
{
    unsigned int hlen = 2 * m * npfb + 1;
    float fc = 0.515f * (rate < 1.0f ? rate : 1.0f);
    if (fc > 0.49f) fc = 0.49f;
    std::vector<float> h = firdes_kaiser(hlen, fc / (float)npfb, fabsf(As));
    double sum = 0.0;
    for (float v : h) sum += (double)v;
    for (float & v : h) v = (float)((double)v * (double)npfb / sum);
    return h;
}

// ---------------------------------------------------------------- byte interleaver index walk
// liquid interleaver: swap pairs (2i, 2j+1) where j follows a column walk of an M x N grid;
// returns j[i] for i < n/2 (each pass is a set of disjoint swaps)
static inline void interleaver_dims(unsigned int n, unsigned int & Mi, unsigned int & Ni)
{
    Mi = 1 + (unsigned int)floorf(sqrtf((float)n));
    Ni = n / Mi;
    while (n >= Mi * Ni) Ni++;
}
static inline std::vector<uint16_t> interleaver_walk(unsigned int n, unsigned int Mi, unsigned int Ni)
{
    unsigned int n2 = n / 2, m = 0, c = n / 3, j;
    std::vector<uint16_t> out(n2);
    for (unsigned int i = 0; i < n2; i++) {
        do {
            j = m * Ni + c;
            m++;
            if (m == Mi) { c = (c + 1) % Ni; m = 0; }
        } while (j >= n2);
        out[i] = (uint16_t)j;
    }
    return out;
}

} // namespace b2
