// ofdmsync.cu -- per-stream OFDM frame synchroniser + header/payload demodulation.
//
// Replaces what the reference does with one liquid call per channel per sample,
//     ofdmflexframesync_execute(framesync[i], &X[i], 1)        lib/multichannelrx.cc:194
//     ofdmflexframesync_execute(fs, &sample, 1)                lib/ofdmtxrx.cc:625
// i.e. liquid's ofdmframesync state machine (seek S0 -> S0a -> S0b -> S1 -> rxsymbols: FFT,
// S0/S1 cross-correlation metrics, CFO + timing recovery, one-tap equaliser, pilot phase
// tracking) and the ofdmflexframesync layer on top (BPSK header -> Golay/CRC, payload QAM
// demap + bit packing).  The FEC/CRC of the payload is left to packet.cu.
//
// Mapping: ONE CTA PER STREAM.  The recurrence between consecutive OFDM symbols of a stream
// (the NCO is trimmed from each symbol's pilot phase before the next symbol is mixed) is
// sequential, so the CTA walks its stream event by event; inside an event all threads share the
// M-point FFT, the correlator reductions (warp shuffles) and the per-subcarrier work.  The
// sample window (M+cp mixed samples), G0, R and all scalar state persist in global memory
// across launches, so results do not depend on how the stream is chunked.
#include "kernels.h"
#include "fec.cuh"

namespace b2 {

void sync_state_init(SyncState & s, unsigned int M, unsigned int cp)
{
    (void)M; (void)cp;
    memset(&s, 0, sizeof(s));
    s.state = ST_SEEK;
    s.g0 = 1.0f;
    s.fstate = FS_HEADER;
    s.ms_payload = 40;  // QPSK
    s.bps_payload = 2;
    s.payload_len = 1;
    s.check = 6; s.fec0 = 1; s.fec1 = 1;
}

struct SyLayout {
    size_t off_st, off_red, off_dsum, off_ring, off_work, off_G0, off_T, off_R, off_tw, off_perm, off_sym, off_yph, total;
};
__host__ __device__ static inline SyLayout sy_layout(unsigned int M, unsigned int cp, unsigned int M_pilot)
{
    SyLayout L;
    size_t o = 0;
    L.off_st = o;   o += (sizeof(SyncState) + 15) & ~(size_t)15;
    L.off_red = o;  o += 96 * sizeof(float);
    L.off_dsum = o; o += (8 * 19 + 40) * sizeof(double);
    L.off_ring = o; o += (size_t)(M + cp) * sizeof(cf);
    L.off_work = o; o += (size_t)M * sizeof(cf);
    L.off_G0 = o;   o += (size_t)M * sizeof(cf);
    L.off_T = o;    o += (size_t)M * sizeof(cf);
    L.off_R = o;    o += (size_t)M * sizeof(cf);
    L.off_tw = o;   o += (size_t)M * sizeof(cf);
    L.off_perm = o; o += (size_t)M * sizeof(uint16_t);
    L.off_sym = o;  o += ((size_t)(M < 64 ? 64 : M) + 15) & ~(size_t)15;
    L.off_yph = o;  o += (size_t)(M_pilot + 4) * sizeof(float) * 3;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}
size_t sync_smem_bytes(const SyncParams & p) { return sy_layout(p.M, p.cp, p.M_pilot + p.M_data).total; }

// hard demodulation of one symbol (liquid modem_demodulate for BPSK / QPSK / square QAM)
__device__ __forceinline__ unsigned int demod_axis(float v, int m, float alpha, float & res)
{
    unsigned int s = 0;
    for (int k = m - 1; k >= 0; k--) {
        float ref = (float)(1u << k) * alpha;
        s <<= 1;
        if (v > 0) { s |= 1u; v -= ref; } else { v += ref; }
    }
    res = v;
    return s ^ (s >> 1);        // gray encode
}
__device__ __forceinline__ unsigned int demod_symbol(cf x, unsigned int scheme, unsigned int bps, float alpha)
{
    if (scheme == 39) return x.x > 0 ? 0u : 1u;                                   // BPSK
    if (scheme == 40) return (x.x > 0 ? 0u : 1u) + (x.y > 0 ? 0u : 2u);           // QPSK
    int m = (int)bps >> 1;
    float ri, rq;
    unsigned int si = demod_axis(x.x, m, alpha, ri);
    unsigned int sq = demod_axis(x.y, m, alpha, rq);
    return (si << m) + sq;
}

__device__ __forceinline__ unsigned int dev_fec_enc_len(unsigned int scheme, unsigned int n)
{
    switch (scheme) {
    case 6:  return (n * 12 + 7) / 8;
    case 7:  { unsigned int blocks = (n * 8 + 11) / 12; return (blocks * 24 + 7) / 8; }
    case 11: return (2 * (8 * n + 6) + 7) / 8;
    default: return n;
    }
}
__device__ __forceinline__ bool dev_fec_ok(unsigned int s) { return s == 1 || s == 6 || s == 7 || s == 11; }
__device__ __forceinline__ unsigned int dev_mod_bps(unsigned int s)
{
    switch (s) { case 39: return 1; case 40: case 25: return 2; case 27: return 4; case 29: return 6; case 31: return 8; default: return 0; }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(256) sync_kernel(const SyncParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int tid = threadIdx.x, nt = blockDim.x;
    const unsigned int M = p.M, cp = p.cp, W = M + cp, M2 = p.M2;
    const unsigned int sidx = blockIdx.x;
    const SyLayout L = sy_layout(M, cp, p.M_pilot + p.M_data);
    SyncState * S = (SyncState *)(smem + L.off_st);
    float * red = (float *)(smem + L.off_red);
    double * dsum = (double *)(smem + L.off_dsum);
    cf * ring = (cf *)(smem + L.off_ring);
    cf * X = (cf *)(smem + L.off_work);
    cf * G0 = (cf *)(smem + L.off_G0);
    cf * T = (cf *)(smem + L.off_T);
    cf * R = (cf *)(smem + L.off_R);
    cf * tw = (cf *)(smem + L.off_tw);
    uint16_t * perm = (uint16_t *)(smem + L.off_perm);
    uint8_t * sym = (uint8_t *)(smem + L.off_sym);
    float * yph = (float *)(smem + L.off_yph);         // [0..Na): y / y_arg, [Na..2Na): y_abs, [2Na..3Na): x_freq
    const unsigned int Na = p.M_pilot + p.M_data;

    // ---- load persistent state
    {
        const uint32_t * src = (const uint32_t *)(p.st + sidx);
        uint32_t * dst = (uint32_t *)S;
        for (unsigned int i = tid; i < sizeof(SyncState) / 4; i += nt) dst[i] = src[i];
        const cf * gr = p.ring + (size_t)sidx * W;
        for (unsigned int i = tid; i < W; i += nt) ring[i] = gr[i];
        const cf * g0 = p.G0 + (size_t)sidx * M;
        const cf * gR = p.R + (size_t)sidx * M;
        for (unsigned int i = tid; i < M; i += nt) {
            G0[i] = g0[i]; R[i] = gR[i];
            tw[i] = p.fft.tw[i];
            perm[i] = p.fft.perm[i];
        }
    }
    __syncthreads();
    FftDev fft = p.fft;
    fft.tw = tw; fft.perm = perm;

    const cf * in = p.in + (size_t)sidx * p.in_stride;
    uint8_t * penc = p.penc + (size_t)sidx * p.penc_cap;
    unsigned int pos = 0;

    // FFT of M window samples starting `off` samples after the oldest one -> X (natural order)
    auto window_fft = [&](unsigned int off) {
        unsigned int head = S->ring_head;
        for (unsigned int i = tid; i < M; i += nt) {
            unsigned int k = head + off + i;
            if (k >= W) k -= W;
            if (k >= W) k -= W;
            X[perm[i]] = ring[k];
        }
        __syncthreads();
        fft_inplace<-1, 0>(X, M, 1, fft, tid, nt);
    };
    // G[i] = X[i] * ref[i] * gain on the subcarriers where ref != 0
    auto gain_est = [&](const float * __restrict__ ref, float gain, cf * G) {
        for (unsigned int i = tid; i < M; i += nt) {
            float r = ref[i];
            cf x = X[i];
            G[i] = (r != 0.0f) ? make_float2(x.x * r * gain, x.y * r * gain) : make_float2(0.f, 0.f);
        }
        __syncthreads();
    };
    // sum_i G[(i+step)%M] * conj(G[i]), i = 0, step, 2*step, ...
    auto metric = [&](const cf * G, unsigned int step) -> cf {
        cf acc = make_float2(0.f, 0.f);
        for (unsigned int i = tid * step; i < M; i += nt * step) {
            cf t = cmulc(G[(i + step) & (M - 1)], G[i]);
            acc.x += t.x; acc.y += t.y;
        }
        return block_sum_cf(acc, red, tid, nt);
    };
    auto phy_reset = [&]() {                 // ofdmframesync_reset
        S->nco_theta = 0; S->nco_dtheta = 0;
        S->pilot_pos = 0;
        S->timer = 0;
        S->num_symbols = 0;
        S->s_hat0_re = 0.f; S->s_hat0_im = 0.f;
        S->phi_prime = 0.f; S->p1_prime = 0.f;
        S->state = ST_SEEK;
    };
    auto flex_reset = [&]() {                // ofdmflexframesync_reset
        S->fstate = FS_HEADER;
        S->header_sym_idx = 0;
        S->payload_sym_idx = 0;
        S->evm_hat = 0.f;
        phy_reset();
    };

    while (true) {
        // ---- advance to the next event (or to the end of this launch's samples)
        const int state = S->state;
        unsigned int need;
        if (state == ST_SEEK) need = (S->timer < (int)M) ? (unsigned int)((int)M - S->timer) : 1u;
        else if (state == ST_S0A || state == ST_S0B) need = (S->timer < (int)M2) ? (unsigned int)((int)M2 - S->timer) : 1u;
        else need = (S->timer > 1) ? (unsigned int)S->timer : 1u;
        const unsigned int avail = p.nsamples - pos;
        const unsigned int adv = min(need, avail);
        {
            const unsigned int head = S->ring_head;
            const uint32_t th = S->nco_theta, dth = S->nco_dtheta;
            const unsigned int skip = adv > W ? adv - W : 0;        // older than the window: never read
            for (unsigned int i = skip + tid; i < adv; i += nt) {
                cf x = in[pos + i];
                if (state != ST_SEEK) x = mix_down(x, nco_cexp(th + i * dth));
                unsigned int k = head + i;
                while (k >= W) k -= W;
                ring[k] = x;
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int h = S->ring_head + adv;
            while (h >= W) h -= W;
            S->ring_head = h;
            if (state != ST_SEEK) S->nco_theta += adv * S->nco_dtheta;
            S->sample_index += adv;
            if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) S->timer += (int)adv;
            else S->timer -= (int)adv;
        }
        pos += adv;
        __syncthreads();
        if (adv < need) break;               // out of samples; state is consistent, resume next launch

        // ---- event
        if (state == ST_SEEK) {
            // g = M / sum |rc[cp .. cp+M)|^2
            float e = 0.f;
            {
                unsigned int head = S->ring_head;
                for (unsigned int i = tid; i < M; i += nt) {
                    unsigned int k = head + cp + i;
                    if (k >= W) k -= W;
                    if (k >= W) k -= W;
                    cf v = ring[k];
                    e += v.x * v.x + v.y * v.y;
                }
            }
            cf es = block_sum_cf(make_float2(e, 0.f), red, tid, nt);
            float g = (float)M / es.x;
            window_fft(cp);
            gain_est(p.tb.S0, sqrtf((float)p.M_S0) / (float)M, G0);
            cf m = metric(G0, 2);
            if (tid == 0) {
                cf s_hat = make_float2(m.x / (float)p.M_S0 * g, m.y / (float)p.M_S0 * g);
                float tau_hat = atan2f(s_hat.y, s_hat.x) * (float)M2 / (2 * 3.14159274101257324219f);
                S->g0 = g;
                S->timer = 0;
                if (hypotf(s_hat.x, s_hat.y) > p.thresh) {
                    int dt = (int)roundf(tau_hat);
                    S->timer = (int)((M + (unsigned int)dt) % M2) + (int)M;
                    S->state = ST_S0A;
                    S->detect_index = S->sample_index - 1;
                }
            }
            __syncthreads();
        } else if (state == ST_S0A) {
            window_fft(cp);
            gain_est(p.tb.S0, sqrtf((float)p.M_S0) / (float)M, G0);
            cf m = metric(G0, 2);
            if (tid == 0) {
                S->timer = 0;
                S->s_hat0_re = m.x / (float)p.M_S0 * S->g0;
                S->s_hat0_im = m.y / (float)p.M_S0 * S->g0;
                S->state = ST_S0B;
            }
            __syncthreads();
        } else if (state == ST_S0B) {
            window_fft(cp);
            gain_est(p.tb.S0, sqrtf((float)p.M_S0) / (float)M, T);
            cf m = metric(T, 2);
            cf acc = make_float2(0.f, 0.f);
            for (unsigned int i = tid; i < M; i += nt) {
                cf t = cmulc(T[i], G0[i]);
                acc.x += t.x; acc.y += t.y;
            }
            cf gs = block_sum_cf(acc, red, tid, nt);
            if (tid == 0) {
                float s1r = m.x / (float)p.M_S0 * S->g0, s1i = m.y / (float)p.M_S0 * S->g0;
                float tau_hat = atan2f(S->s_hat0_im + s1i, S->s_hat0_re + s1r) * (float)M2 / (2 * 3.14159274101257324219f);
                S->timer = (int)(M + cp - p.backoff) - (int)roundf(tau_hat);
                float nu_hat = 2.0f * atan2f(gs.y, gs.x) / (float)M;
                S->nco_dtheta = nco_constrain_dev(nu_hat);
                S->state = ST_S1;
            }
            __syncthreads();
        } else if (state == ST_S1) {
            window_fft(cp);
            gain_est(p.tb.S1, sqrtf((float)p.M_S1) / (float)M, T);
            cf m = metric(T, 1);
            int accept = 0;
            if (tid == 0) {
                S->num_symbols++;
                cf s_hat = make_float2(m.x / (float)p.M_S1 * S->g0, m.y / (float)p.M_S1 * S->g0);
                float a = (float)p.backoff * 2.0f * 3.14159274101257324219f / (float)M;
                s_hat = cmul(s_hat, make_float2(cosf(a), sinf(a)));
                accept = (hypotf(s_hat.x, s_hat.y) > p.thresh) && (fabsf(atan2f(s_hat.y, s_hat.x)) < 0.1f * 3.14159274101257324219f);
                red[64] = (float)accept;
            }
            __syncthreads();
            accept = red[64] != 0.f;
            if (accept) {
                // G *= M/sqrt(Na) * B ; smooth |G| and arg G with an order-4 polynomial over the
                // active subcarriers (liquid ofdmframesync_estimate_eqgain_poly); R = B / G
                const float gsc = (float)M / sqrtf((float)Na);
                for (unsigned int i = tid; i < M; i += nt) T[i] = cmul(cscale(T[i], gsc), p.tb.B[i]);
                __syncthreads();
                for (unsigned int n = tid; n < Na; n += nt) {
                    unsigned int k = p.tb.active_idx[n];
                    float xf = (k > M2) ? (float)k - (float)M : (float)k;
                    yph[2 * Na + n] = xf / (float)M;
                    yph[Na + n] = hypotf(T[k].x, T[k].y);
                    yph[n] = atan2f(T[k].y, T[k].x);
                }
                __syncthreads();
                if (tid == 0) {
                    const float pi = 3.14159274101257324219f;
                    for (unsigned int i = 1; i < Na; i++) {
                        while ((yph[i] - yph[i - 1]) > pi) yph[i] -= 2 * pi;
                        while ((yph[i] - yph[i - 1]) < -pi) yph[i] += 2 * pi;
                    }
                }
                __syncthreads();
                unsigned int order = 4;
                if (order > Na - 1) order = Na - 1;
                const unsigned int kc = order + 1;                  // coefficients
                // power sums S_0..S_{2k-2}, moments of |G| and arg G; per-thread partials in double
                double ps[9], pa[5], pg[5];
#pragma unroll
                for (int i = 0; i < 9; i++) ps[i] = 0.0;
#pragma unroll
                for (int i = 0; i < 5; i++) { pa[i] = 0.0; pg[i] = 0.0; }
                for (unsigned int n = tid; n < Na; n += nt) {
                    double xv = (double)yph[2 * Na + n], ya = (double)yph[Na + n], yg = (double)yph[n];
                    double xp = 1.0;
#pragma unroll
                    for (int r = 0; r < 9; r++) {
                        ps[r] += xp;
                        if (r < 5) { pa[r] += xp * ya; pg[r] += xp * yg; }
                        xp *= xv;
                    }
                }
#pragma unroll
                for (int i = 0; i < 9; i++) ps[i] = warp_sum_d(ps[i]);
#pragma unroll
                for (int i = 0; i < 5; i++) { pa[i] = warp_sum_d(pa[i]); pg[i] = warp_sum_d(pg[i]); }
                {
                    // deterministic cross-warp sum: per-warp partials, then warp order
                    double * dpart = dsum + 40;
                    const unsigned int wid = tid >> 5, nw = (nt + 31) >> 5;
                    if ((tid & 31) == 0) {
                        for (int i = 0; i < 9; i++) dpart[wid * 19 + i] = ps[i];
                        for (int i = 0; i < 5; i++) { dpart[wid * 19 + 9 + i] = pa[i]; dpart[wid * 19 + 14 + i] = pg[i]; }
                    }
                    __syncthreads();
                    if (tid < 19) {
                        double a = 0.0;
                        for (unsigned int w = 0; w < nw; w++) a += dpart[w * 19 + tid];
                        dsum[tid] = a;
                    }
                    __syncthreads();
                }
                if (tid < 2) {
                    // solve the kc x kc normal equations (Gaussian elimination, partial pivoting)
                    double A[5][6];
                    const double * rhs = dsum + (tid == 0 ? 9 : 14);
                    for (unsigned int r = 0; r < kc; r++) {
                        for (unsigned int c = 0; c < kc; c++) A[r][c] = dsum[r + c];
                        A[r][kc] = rhs[r];
                    }
                    for (unsigned int c = 0; c < kc; c++) {
                        unsigned int piv = c;
                        for (unsigned int r = c + 1; r < kc; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
                        if (piv != c) for (unsigned int i = 0; i <= kc; i++) { double t = A[c][i]; A[c][i] = A[piv][i]; A[piv][i] = t; }
                        for (unsigned int r = c + 1; r < kc; r++) {
                            double f = A[r][c] / A[c][c];
                            for (unsigned int i = c; i <= kc; i++) A[r][i] -= f * A[c][i];
                        }
                    }
                    double coef[5] = {0, 0, 0, 0, 0};
                    for (unsigned int r = kc; r-- > 0;) {
                        double s = A[r][kc];
                        for (unsigned int c = r + 1; c < kc; c++) s -= A[r][c] * coef[c];
                        coef[r] = s / A[r][r];
                    }
                    for (int i = 0; i < 5; i++) dsum[20 + 5 * tid + i] = coef[i];
                }
                __syncthreads();
                for (unsigned int i = tid; i < M; i += nt) {
                    if (p.tb.sctype[i] == 0) { R[i] = make_float2(0.f, 0.f); continue; }
                    float freq = ((i > M2) ? (float)i - (float)M : (float)i) / (float)M;
                    double xv = (double)freq, xp = 1.0, va = 0.0, vg = 0.0;
                    for (unsigned int r = 0; r < kc; r++) { va += dsum[20 + r] * xp; vg += dsum[25 + r] * xp; xp *= xv; }
                    float A = (float)va, th = (float)vg;
                    float sn, cs;
                    sincosf(th, &sn, &cs);
                    cf G = make_float2(A * cs, A * sn);
                    cf B = p.tb.B[i];
                    float d = G.x * G.x + G.y * G.y;
                    cf num = cmulc(B, G);
                    R[i] = make_float2(num.x / d, num.y / d);
                }
                if (tid == 0) {
                    S->state = ST_RX;
                    S->timer = (int)(M + cp + p.backoff);
                    S->num_symbols = 0;
                }
            } else if (tid == 0) {
                if (S->num_symbols == 16) phy_reset();
                else S->timer = (int)M2;
            }
            __syncthreads();
        } else {
            // ---- ST_RX: one OFDM symbol
            window_fft(cp - p.backoff);
            for (unsigned int i = tid; i < M; i += nt) X[i] = cmul(X[i], R[i]);
            __syncthreads();
            // pilot phases in fft-shifted order
            const unsigned int ppos = S->pilot_pos;
            for (unsigned int n = tid; n < p.M_pilot; n += nt) {
                unsigned int k = p.tb.pilot_idx[n];
                unsigned int q = ppos + n;
                q %= 255u;
                float pil = p.tb.pilot_seq[q] ? 1.0f : -1.0f;
                yph[n] = atan2f(X[k].y * pil, X[k].x * pil);
            }
            __syncthreads();
            if (tid == 0) {
                const float pi = 3.14159274101257324219f;
                float sy = 0.f, sxy = 0.f;
                float prev = yph[0];
                sy = __fadd_rn(sy, prev);
                sxy = __fadd_rn(sxy, __fmul_rn(p.tb.pilot_x[0], prev));
                for (unsigned int i = 1; i < p.M_pilot; i++) {
                    float y = yph[i];
                    while ((y - prev) > pi) y -= 2 * pi;
                    while ((y - prev) < -pi) y += 2 * pi;
                    sy = __fadd_rn(sy, y);
                    sxy = __fadd_rn(sxy, __fmul_rn(p.tb.pilot_x[i], y));
                    prev = y;
                }
                const float np = (float)p.M_pilot, sx = p.pilot_sx, sxx = p.pilot_sxx;
                float den = __fsub_rn(__fmul_rn(np, sxx), __fmul_rn(sx, sx));
                float p1 = __fdiv_rn(__fsub_rn(__fmul_rn(np, sxy), __fmul_rn(sx, sy)), den);
                float p0 = __fdiv_rn(__fsub_rn(sy, __fmul_rn(p1, sx)), np);
                const float alpha = 0.3f;
                p1 = __fadd_rn(__fmul_rn(alpha, p1), __fmul_rn(1 - alpha, S->p1_prime));
                S->p1_prime = p1;
                red[65] = p0; red[66] = p1;
                if (S->num_symbols > 0) {
                    float dphi = p0 - S->phi_prime;
                    while (dphi > pi) dphi -= 2 * pi;
                    while (dphi < -pi) dphi += 2 * pi;
                    S->nco_dtheta += nco_constrain_dev(1e-3f * dphi);
                }
                S->phi_prime = p0;
                S->num_symbols++;
                S->pilot_pos = (ppos + p.M_pilot) % 255u;
            }
            __syncthreads();
            {
                const float p0 = red[65], p1 = red[66];
                for (unsigned int i = tid; i < M; i += nt) {
                    if (p.tb.sctype[i] == 0) { X[i] = make_float2(0.f, 0.f); continue; }
                    float fx = (i > M2) ? (float)i - (float)M : (float)i;
                    float th = __fadd_rn(p0, __fmul_rn(p1, fx));
                    float sn, cs;
                    sincosf(th, &sn, &cs);
                    X[i] = cmul(X[i], make_float2(cs, -sn));
                }
            }
            __syncthreads();

            // ---- debug tap
            if (p.tap_cap) {
                if (tid == 0) red[67] = __uint_as_float(atomicAdd(&p.counters[4], 1u));
                __syncthreads();
                unsigned int slot = __float_as_uint(red[67]);
                if (slot < p.tap_cap) {
                    for (unsigned int i = tid; i < M; i += nt) p.tap_X[(size_t)slot * M + i] = X[i];
                    if (tid == 0) { p.tap_chan[slot] = sidx; p.tap_index[slot] = S->sample_index - 1; }
                }
            }

            // ---- ofdmflexframesync layer
            int emit = 0;                       // 1: header invalid, 2: payload complete
            if (S->fstate == FS_HEADER) {
                const unsigned int start = S->header_sym_idx;
                const unsigned int take = min(p.M_data, 288u - start);
                float ev = 0.f;
                for (unsigned int d = tid; d < take; d += nt) {
                    cf x = X[p.tb.data_idx[d]];
                    unsigned int b = x.x > 0 ? 0u : 1u;
                    sym[d] = (uint8_t)b;
                    float dr = x.x - (b ? -1.0f : 1.0f);
                    ev += dr * dr + x.y * x.y;
                }
                cf evs = block_sum_cf(make_float2(ev, 0.f), red, tid, nt);
                {
                    const unsigned int j0 = start >> 3, j1 = (start + take - 1) >> 3;
                    for (unsigned int j = j0 + tid; j <= j1; j += nt) {
                        unsigned int v = 0;
                        for (unsigned int b = 0; b < 8; b++) {
                            unsigned int bit = 8 * j + b;
                            if (bit >= start && bit < start + take) v |= (unsigned int)sym[bit - start] << (7 - b);
                        }
                        if (8 * j < start) v |= S->header_bits[j];
                        S->header_bits[j] = (uint8_t)v;
                    }
                }
                __syncthreads();
                if (tid == 0) { S->evm_hat += evs.x; S->header_sym_idx = start + take; }
                if (start + take == 288u) {
                    // unscramble, de-interleave (n = 36, depth 4), Golay(24,12), CRC-32, parse
                    uint8_t * hb = sym;                        // 36 bytes of scratch
                    uint32_t * gsym = (uint32_t *)yph;         // 12 decoded Golay symbols
                    if (tid == 0) {
                        const uint8_t mask[4] = {0xb4, 0x6a, 0x8b, 0x45};
                        for (int i = 0; i < 36; i++) hb[i] = S->header_bits[i] ^ mask[i & 3];
                        const uint8_t ilmask[4] = {0xff, 0x0f, 0x55, 0x33};
                        for (int v = 3; v >= 0; v--) {
                            const uint16_t * walk = p.tb.hdr_walk + 18 * v;
                            uint8_t mk = ilmask[v];
                            for (int i = 0; i < 18; i++) {
                                unsigned int j = walk[i];
                                uint8_t a = hb[2 * i], b = hb[2 * j + 1];
                                hb[2 * i] = (uint8_t)((a & ~mk) | (b & mk));
                                hb[2 * j + 1] = (uint8_t)((a & mk) | (b & ~mk));
                            }
                        }
                    }
                    __syncthreads();
                    if (tid < 12) {
                        unsigned int v = ((unsigned int)hb[3 * tid] << 16) | ((unsigned int)hb[3 * tid + 1] << 8) | hb[3 * tid + 2];
                        gsym[tid] = golay2412_decode(v);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        uint8_t * hd = S->header_dec;
                        for (int g = 0; g < 6; g++) {
                            unsigned int s0 = gsym[2 * g], s1 = gsym[2 * g + 1];
                            hd[3 * g] = (s0 >> 4) & 0xff;
                            hd[3 * g + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
                            hd[3 * g + 2] = s1 & 0xff;
                        }
                        uint32_t key = ((uint32_t)hd[14] << 24) | ((uint32_t)hd[15] << 16) | ((uint32_t)hd[16] << 8) | hd[17];
                        int valid = crc32_bytes(hd, 14) == key;
                        S->evm_db = 10 * log10f(S->evm_hat / 288.0f);
                        if (valid && hd[8] != 105) valid = 0;          // protocol id
                        unsigned int plen = ((unsigned int)hd[9] << 8) | hd[10];
                        unsigned int ms = hd[11], check = (hd[12] >> 5) & 7, fec0 = hd[12] & 0x1f, fec1 = hd[13] & 0x1f;
                        unsigned int bps = dev_mod_bps(ms);
                        if (valid && (bps == 0 || (check != 6 && check != 1) || !dev_fec_ok(fec0) || !dev_fec_ok(fec1))) valid = 0;
                        unsigned int enc = 0;
                        if (valid) {
                            enc = dev_fec_enc_len(fec1, dev_fec_enc_len(fec0, plen + (check == 6 ? 4 : 0)));
                            if (enc > p.penc_cap) valid = 0;          // cannot happen with penc_cap at its default
                        }
                        if (valid) {
                            S->ms_payload = ms; S->bps_payload = bps; S->payload_len = plen;
                            S->check = check; S->fec0 = fec0; S->fec1 = fec1;
                            S->payload_enc_len = enc;
                            S->payload_mod_len = (8 * enc + bps - 1) / bps;
                            S->fstate = FS_PAYLOAD;
                        }
                        red[68] = (float)valid;
                    }
                    __syncthreads();
                    if (red[68] != 0.f) {
                        const unsigned int enc = S->payload_enc_len;
                        for (unsigned int i = tid; i < (enc + 3) / 4; i += nt) ((uint32_t *)penc)[i] = 0u;
                    } else {
                        emit = 1;
                    }
                }
            } else {
                const unsigned int bps = S->bps_payload, ms = S->ms_payload;
                const unsigned int start = S->payload_sym_idx;
                const unsigned int take = min(p.M_data, S->payload_mod_len - start);
                for (unsigned int d = tid; d < take; d += nt) sym[d] = (uint8_t)demod_symbol(X[p.tb.data_idx[d]], ms, bps, p.qam_alpha[bps]);
                __syncthreads();
                const unsigned int bit0 = start * bps, nbits = take * bps, enc = S->payload_enc_len;
                unsigned int j0 = bit0 >> 3, j1 = (bit0 + nbits - 1) >> 3;
                if (j1 >= enc) j1 = enc - 1;
                for (unsigned int j = j0 + tid; j <= j1; j += nt) {
                    unsigned int v = 0;
                    for (unsigned int b = 0; b < 8; b++) {
                        unsigned int bit = 8 * j + b;
                        if (bit >= bit0 && bit < bit0 + nbits) {
                            unsigned int rel = bit - bit0, d = rel / bps, k = rel - d * bps;
                            v |= (((unsigned int)sym[d] >> (bps - 1 - k)) & 1u) << (7 - b);
                        }
                    }
                    if (8 * j < bit0) v |= penc[j];
                    penc[j] = (uint8_t)v;
                }
                __syncthreads();
                if (tid == 0) S->payload_sym_idx = start + take;
                if (start + take == S->payload_mod_len) emit = 2;
            }
            __syncthreads();

            if (emit) {
                // append a frame record (+ encoded payload) to the output of this launch
                if (tid == 0) {
                    unsigned int slot = atomicAdd(&p.counters[0], 1u);
                    unsigned long long off = 0;
                    unsigned int enc = (emit == 2) ? S->payload_enc_len : 0u;
                    int ok = slot < p.recs_cap;
                    if (ok && enc) {
                        off = atomicAdd((unsigned long long *)(p.counters + 2), (unsigned long long)((enc + 15u) & ~15u));
                        if (off + enc > p.arena_cap) ok = 0;
                    }
                    if (!ok) { atomicExch(&p.counters[1], 1u); red[69] = -1.f; }
                    else {
                        FrameRec r;
                        r.channel = sidx;
                        r.header_valid = (emit == 2);
                        r.payload_valid = 0;
                        r.payload_len = (emit == 2) ? S->payload_len : 0u;
                        for (int i = 0; i < 8; i++) r.header[i] = S->header_dec[i];
                        r.evm = S->evm_db;
                        r.rssi = -10.0f * log10f(S->g0);
                        r.cfo = nco_freq_dev(S->nco_dtheta);
                        r.mod_scheme = (emit == 2) ? S->ms_payload : 0u;
                        r.mod_bps = (emit == 2) ? S->bps_payload : 0u;
                        r.check = (emit == 2) ? S->check : 0u;
                        r.fec0 = (emit == 2) ? S->fec0 : 0u;
                        r.fec1 = (emit == 2) ? S->fec1 : 0u;
                        r.detect_index = S->detect_index;
                        r.complete_index = S->sample_index - 1;
                        r.payload_offset = off;
                        p.recs[slot] = r;
                        FrameAux a; a.enc_len = enc; a.pad = 0;
                        p.aux[slot] = a;
                        red[69] = 1.f;
                        dsum[30] = __longlong_as_double((long long)off);
                    }
                }
                __syncthreads();
                if (emit == 2 && red[69] > 0.f) {
                    const unsigned long long off = (unsigned long long)__double_as_longlong(dsum[30]);
                    const unsigned int enc = S->payload_enc_len;
                    uint32_t * dst = (uint32_t *)(p.arena + off);
                    const uint32_t * src = (const uint32_t *)penc;
                    for (unsigned int i = tid; i < (enc + 3) / 4; i += nt) dst[i] = src[i];
                }
                __syncthreads();
                if (tid == 0) flex_reset();
            }
            if (tid == 0) S->timer = (int)(M + cp);          // set unconditionally, as liquid does
            __syncthreads();
        }
    }

    // ---- store persistent state
    {
        uint32_t * dst = (uint32_t *)(p.st + sidx);
        const uint32_t * src = (const uint32_t *)S;
        for (unsigned int i = tid; i < sizeof(SyncState) / 4; i += nt) dst[i] = src[i];
        cf * gr = p.ring + (size_t)sidx * W;
        for (unsigned int i = tid; i < W; i += nt) gr[i] = ring[i];
        cf * g0 = p.G0 + (size_t)sidx * M;
        cf * gR = p.R + (size_t)sidx * M;
        for (unsigned int i = tid; i < M; i += nt) { g0[i] = G0[i]; gR[i] = R[i]; }
    }
}

cudaError_t sync_configure(size_t smem_bytes)
{
    return cudaFuncSetAttribute(sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

cudaError_t sync_launch(const SyncParams & p, int threads, size_t smem_bytes, cudaStream_t st)
{
    if (p.nsamples == 0 || p.streams == 0) return cudaSuccess;
    sync_kernel<<<p.streams, threads, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

} // namespace b2
