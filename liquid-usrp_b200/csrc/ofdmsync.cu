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
// sequential, so the CTA walks its stream event by event (an event = the sample on which
// liquid's timer fires).  Per event: the samples up to the event were prefetched into shared
// memory with cp.async while the previous event was processed; they are NCO-mixed and pushed
// into the M+cp sample window and, when the FFT window consists of new samples only, straight
// into the FFT buffer in digit-reversed order; the M-point FFT runs in place in shared memory;
// correlator metrics are warp-shuffle reductions; the pilot phase unwrap + line fit is a
// warp-parallel scan in warp 0; equalise / derotate / demap touch each subcarrier once.
// The sample window, G0, R and all scalar state persist in global memory across launches, so
// results do not depend on how the stream is chunked.
#include <cstdlib>
#include "kernels.h"
#include "fec.cuh"
#include "syncdev.cuh"

namespace b2 {

#define PX(i) ((i) + ((i) >> 3))          // skewed index into the FFT buffer (one pad per 8: radix-8 strides)

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
    unsigned int SZ;            // staging ring size (power of two)
    unsigned int PF;            // prefetch distance (samples)
    size_t off_st, off_red, off_dsum, off_ring, off_X, off_G0, off_T, off_R, off_tw, off_perm, off_rank, off_sym, off_yph, off_stg, off_ref, off_px, off_pseq, total;
};
__host__ __device__ static inline SyLayout sy_layout(unsigned int M, unsigned int cp, unsigned int Na)
{
    SyLayout L;
    L.PF = M + cp + M / 4 + 8;
    L.SZ = 64;
    while (L.SZ < L.PF) L.SZ <<= 1;
    size_t o = 0;
    L.off_st = o;   o += (sizeof(SyncState) + 15) & ~(size_t)15;
    L.off_red = o;  o += 128 * sizeof(float);
    L.off_dsum = o; o += (8 * 19 + 40) * sizeof(double);
    L.off_ring = o; o += (size_t)(M + cp) * sizeof(cf);
    L.off_X = o;    o += (size_t)(M + (M >> 3) + 2) * sizeof(cf);      // skewed: one pad element per 8
    L.off_G0 = o;   o += (size_t)M * sizeof(cf);
    L.off_T = o;    o += (size_t)M * sizeof(cf);
    L.off_R = o;    o += (size_t)M * sizeof(cf);
    L.off_tw = o;   o += (size_t)M * sizeof(cf);
    L.off_perm = o; o += (size_t)M * sizeof(uint16_t);
    L.off_rank = o; o += (size_t)M * sizeof(uint16_t);
    L.off_sym = o;  o += ((size_t)(M < 64 ? 64 : M) + 15) & ~(size_t)15;
    L.off_yph = o;  o += (size_t)(Na + 4) * sizeof(float) * 3;
    o = (o + 15) & ~(size_t)15;
    L.off_stg = o;  o += (size_t)L.SZ * sizeof(cf);
    L.off_ref = o;  o += (size_t)2 * M * sizeof(float);            // S0 | S1 training signs
    L.off_px = o;   o += (size_t)(Na + 4) * sizeof(float);         // pilot abscissae (at most Na pilots)
    L.off_pseq = o; o += 256;                                      // one period of the pilot LFSR
    L.total = (o + 15) & ~(size_t)15;
    return L;
}
size_t sync_smem_bytes(const SyncParams & p) { return sy_layout(p.M, p.cp, p.M_pilot + p.M_data).total; }

// optional phase profile (build with -DB2_SYNC_PROF): cycles spent by CTA 0 between barriers
#ifdef B2_SYNC_PROF
__device__ unsigned long long g_sync_prof[16];
#define PH(k) do { if (blockIdx.x == 0 && tid == 0) { long long _t = clock64(); atomicAdd(&g_sync_prof[k], (unsigned long long)(_t - t_last)); t_last = _t; } } while (0)
extern "C" int b2_debug_sync_prof(unsigned long long * out, int reset)
{
    if (out) cudaMemcpyFromSymbol(out, g_sync_prof, sizeof(g_sync_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_sync_prof, z, sizeof(z)); }
    return 0;
}
#else
#define PH(k) do { } while (0)
#endif

// ------------------------------------------------------------------ the kernel
// MT / NT: number of subcarriers / threads known at compile time (0 = take them from the
// launch), so the per-subcarrier loops unroll and the FFT passes are fixed
template <unsigned int MT, unsigned int NT>
__global__ void __launch_bounds__(256) sync_kernel(const SyncParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int tid = threadIdx.x, nt = NT ? NT : blockDim.x, lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
    const unsigned int M = MT ? MT : p.M, cp = p.cp, W = M + cp, M2 = M / 2;
    const unsigned int sidx = blockIdx.x;
    const unsigned int Na = p.M_pilot + p.M_data;
    const SyLayout L = sy_layout(M, cp, Na);
    SyncState * S = (SyncState *)(smem + L.off_st);
    float * red = (float *)(smem + L.off_red);
    double * dsum = (double *)(smem + L.off_dsum);
    cf * ring = (cf *)(smem + L.off_ring);
    cf * X = (cf *)(smem + L.off_X);
    cf * G0 = (cf *)(smem + L.off_G0);
    cf * T = (cf *)(smem + L.off_T);
    cf * R = (cf *)(smem + L.off_R);
    cf * tw = (cf *)(smem + L.off_tw);
    uint16_t * perm = (uint16_t *)(smem + L.off_perm);
    uint16_t * rank = (uint16_t *)(smem + L.off_rank);
    uint8_t * sym = (uint8_t *)(smem + L.off_sym);
    float * yph = (float *)(smem + L.off_yph);         // [0..Na): y / y_arg, [Na..2Na): y_abs, [2Na..3Na): x_freq
    cf * stg = (cf *)(smem + L.off_stg);
    float * refS = (float *)(smem + L.off_ref);
    float * pilot_x = (float *)(smem + L.off_px);
    uint8_t * pilot_seq = (uint8_t *)(smem + L.off_pseq);
    const unsigned int SZM = L.SZ - 1, PF = L.PF;

    const cf * in = p.in + (size_t)sidx * p.in_stride;
    uint8_t * penc = p.penc + (size_t)sidx * p.penc_cap;

    // ---- first prefetch, then load persistent state while it is in flight
    unsigned int fetched = 0;
    auto prefetch = [&](unsigned int upto) {
        for (unsigned int i = fetched + tid; i < upto; i += nt) cp_async8(&stg[i & SZM], in + i);
        fetched = upto;
    };
    prefetch(min(PF, p.nsamples));
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
            perm[i] = (uint16_t)phys<2>(p.fft.perm[i]);         // FFT buffer positions are skewed (bank conflicts)
            rank[i] = p.tb.sc_rank[i];
            refS[i] = p.tb.S0[i];
            refS[M + i] = p.tb.S1[i];
        }
        for (unsigned int i = tid; i < p.M_pilot; i += nt) pilot_x[i] = p.tb.pilot_x[i];
        for (unsigned int i = tid; i < 255; i += nt) pilot_seq[i] = p.tb.pilot_seq[i];
    }
    FftDev fft = p.fft;
    fft.tw = tw; fft.perm = perm;
    unsigned int pos = 0;
#ifdef B2_SYNC_PROF
    long long t_last = clock64();
#endif

    // block-wide sum of up to 4 floats per thread; result in red[100..103] after the call
    auto block_sum4 = [&](float a, float b, float c, float d) {
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
        if (lane == 0) { red[4 * wid] = a; red[4 * wid + 1] = b; red[4 * wid + 2] = c; red[4 * wid + 3] = d; }
        __syncthreads();
        if (tid < 4) {
            float s = 0.f;
            for (unsigned int w = 0; w < nw; w++) s += red[4 * w + tid];
            red[100 + tid] = s;
        }
        __syncthreads();
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
        PH(6);
        cp_async_wait_all();
        __syncthreads();                     // staged samples + state of the previous event visible
        PH(0);
        // ---- advance to the next event (or to the end of this launch's samples)
        const int state = S->state;
        const int timer = S->timer;
        const unsigned int head = S->ring_head;
        const uint32_t th = S->nco_theta, dth = S->nco_dtheta;
        unsigned int need;
        if (state == ST_SEEK) need = (timer < (int)M) ? (unsigned int)((int)M - timer) : 1u;
        else if (state == ST_S0A || state == ST_S0B) need = (timer < (int)M2) ? (unsigned int)((int)M2 - timer) : 1u;
        else need = (timer > 1) ? (unsigned int)timer : 1u;
        const unsigned int avail = p.nsamples - pos;
        const unsigned int adv = min(need, avail);
        const bool fire = (adv == need);
        const unsigned int off = (state == ST_RX) ? cp - p.backoff : cp;    // FFT window offset in the sample window
        const bool fused = fire && (adv >= W - off);                        // FFT window made of new samples only
        const bool mixing = (state != ST_SEEK) && ((th | dth) != 0u);     // e^{-j0} = 1 exactly
        float en = 0.f;
        for (unsigned int j = tid; j < adv; j += nt) {
            cf x = stg[(pos + j) & SZM];
            if (mixing) x = mix_down(x, nco_cexp_fast(th + j * dth));
            if (j + W >= adv) {
                unsigned int k = head + j;
                while (k >= W) k -= W;
                ring[k] = x;
            }
            if (fused) {
                int i = (int)W - (int)adv + (int)j - (int)off;
                if (i >= 0 && i < (int)M) {
                    X[perm[i]] = x;
                    en += x.x * x.x + x.y * x.y;
                }
            }
        }
        pos += adv;
        unsigned int head2 = head + adv;
        while (head2 >= W) head2 -= W;
        __syncthreads();
        PH(1);
        if (tid == 0) {
            S->ring_head = head2;
            if (state != ST_SEEK) S->nco_theta = th + adv * dth;
            S->sample_index += adv;
            if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) S->timer = timer + (int)adv;
            else S->timer = timer - (int)adv;
        }
        if (!fire) break;                    // out of samples; resume in the next launch
        prefetch(min(pos + PF, p.nsamples));
        if (!fused) {
            for (unsigned int i = tid; i < M; i += nt) {
                unsigned int k = head2 + off + i;
                if (k >= W) k -= W;
                if (k >= W) k -= W;
                cf x = ring[k];
                X[perm[i]] = x;
                en += x.x * x.x + x.y * x.y;
            }
            __syncthreads();
        }
        PH(8);
        if (MT) fft_static<MT, -1, 2>(X, tw, tid, nt);
        else fft_inplace<-1, 2>(X, M, 1, fft, tid, nt);
        PH(2);

        if (state != ST_RX) {
            // ---- preamble events.  G[i] = X[i]*ref[i]*gain on the training subcarriers
            const bool long_seq = (state == ST_S1);
            const float * __restrict__ ref = long_seq ? refS + M : refS;
            const unsigned int step = long_seq ? 1u : 2u;
            const float gain = sqrtf((float)(long_seq ? p.M_S1 : p.M_S0)) / (float)M;
            float mr = 0.f, mi = 0.f, cr = 0.f, ci = 0.f;
            for (unsigned int i = tid; i < M; i += nt) {
                float r = ref[i];
                cf g = make_float2(X[PX(i)].x * r * gain, X[PX(i)].y * r * gain);
                if ((i & (step - 1)) == 0) {
                    unsigned int i2 = i + step; if (i2 >= M) i2 -= M;
                    float r2 = ref[i2];
                    cf g2 = make_float2(X[PX(i2)].x * r2 * gain, X[PX(i2)].y * r2 * gain);
                    cf t = cmulc(g2, g);
                    mr += t.x; mi += t.y;
                }
                if (state == ST_S0A) G0[i] = g;
                else if (state == ST_S0B) { cf t = cmulc(g, G0[i]); cr += t.x; ci += t.y; }
                else if (state == ST_S1) T[i] = g;
            }
            if (state == ST_SEEK) cr = en;
            block_sum4(mr, mi, cr, ci);
            if (state == ST_SEEK) {
                if (tid == 0) {
                    float g = (float)M / red[102];
                    cf s_hat = make_float2(red[100] / (float)p.M_S0 * g, red[101] / (float)p.M_S0 * g);
                    float tau_hat = atan2f(s_hat.y, s_hat.x) * (float)M2 / (2 * PI_F);
                    S->g0 = g;
                    S->timer = 0;
                    if (hypotf(s_hat.x, s_hat.y) > p.thresh) {
                        int dt = (int)roundf(tau_hat);
                        S->timer = (int)((M + (unsigned int)dt) % M2) + (int)M;
                        S->state = ST_S0A;
                        S->detect_index = S->sample_index - 1;
                    }
                }
            } else if (state == ST_S0A) {
                if (tid == 0) {
                    S->timer = 0;
                    S->s_hat0_re = red[100] / (float)p.M_S0 * S->g0;
                    S->s_hat0_im = red[101] / (float)p.M_S0 * S->g0;
                    S->state = ST_S0B;
                }
            } else if (state == ST_S0B) {
                if (tid == 0) {
                    float s1r = red[100] / (float)p.M_S0 * S->g0, s1i = red[101] / (float)p.M_S0 * S->g0;
                    float tau_hat = atan2f(S->s_hat0_im + s1i, S->s_hat0_re + s1r) * (float)M2 / (2 * PI_F);
                    S->timer = (int)(M + cp - p.backoff) - (int)roundf(tau_hat);
                    float nu_hat = 2.0f * atan2f(red[103], red[102]) / (float)M;
                    S->nco_dtheta = nco_constrain_dev(nu_hat);
                    S->state = ST_S1;
                }
            } else {
                // ---- S1: accept / retry, and on accept the equaliser
                if (tid == 0) {
                    S->num_symbols++;
                    cf s_hat = make_float2(red[100] / (float)p.M_S1 * S->g0, red[101] / (float)p.M_S1 * S->g0);
                    float a = (float)p.backoff * 2.0f * PI_F / (float)M;
                    s_hat = cmul(s_hat, make_float2(cosf(a), sinf(a)));
                    int accept = (hypotf(s_hat.x, s_hat.y) > p.thresh) && (fabsf(atan2f(s_hat.y, s_hat.x)) < 0.1f * PI_F);
                    red[110] = (float)accept;
                    if (!accept) {
                        if (S->num_symbols == 16) phy_reset();
                        else S->timer = (int)M2;
                    }
                }
                __syncthreads();
                if (red[110] != 0.f) {
                    // G *= M/sqrt(Na) * B ; smooth |G| and arg G with an order-4 polynomial over the
                    // active subcarriers (liquid ofdmframesync_estimate_eqgain_poly); R = B / G
                    const float gsc = (float)M / sqrtf((float)Na);
                    for (unsigned int n = tid; n < Na; n += nt) {
                        unsigned int k = p.tb.active_idx[n];
                        cf g = cmul(cscale(T[k], gsc), p.tb.B[k]);
                        float xf = (k > M2) ? (float)k - (float)M : (float)k;
                        yph[2 * Na + n] = xf / (float)M;
                        yph[Na + n] = hypotf(g.x, g.y);
                        yph[n] = atan2f(g.y, g.x);
                    }
                    __syncthreads();
                    if (wid == 0) {
                        float a, b;
                        warp_unwrap(yph, nullptr, Na, true, lane, a, b);
                    }
                    __syncthreads();
                    // power sums S_0..S_8 and moments of |G|, arg G; per-thread partials in double
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
                    double * dpart = dsum + 40;
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 9; i++) dpart[wid * 19 + i] = ps[i];
#pragma unroll
                        for (int i = 0; i < 5; i++) { dpart[wid * 19 + 9 + i] = pa[i]; dpart[wid * 19 + 14 + i] = pg[i]; }
                    }
                    __syncthreads();
                    if (tid < 19) {
                        double a = 0.0;
                        for (unsigned int w = 0; w < nw; w++) a += dpart[w * 19 + tid];
                        dsum[tid] = a;
                    }
                    __syncthreads();
                    if (tid == 0) solve5(dsum, dsum + 9, dsum + 20);
                    if (tid == 32 || (nt <= 32 && tid == 1)) solve5(dsum, dsum + 14, dsum + 25);
                    __syncthreads();
                    for (unsigned int i = tid; i < M; i += nt) {
                        if (rank[i] == 0xffffu) { R[i] = make_float2(0.f, 0.f); continue; }
                        float freq = ((i > M2) ? (float)i - (float)M : (float)i) / (float)M;
                        double xv = (double)freq, xp = 1.0, va = 0.0, vg = 0.0;
#pragma unroll
                        for (int r = 0; r < 5; r++) { va += dsum[20 + r] * xp; vg += dsum[25 + r] * xp; xp *= xv; }
                        float A = (float)va, thv = (float)vg;
                        float sn, cs;
                        sincosf(thv, &sn, &cs);
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
                }
            }
            PH(7);
            continue;                        // loop top synchronises
        }

        // ---- ST_RX: one OFDM symbol.  Equalise (each thread owns subcarriers tid, tid+nt, ...)
        const unsigned int ppos = S->pilot_pos;
        for (unsigned int i = tid; i < M; i += nt) {
            cf xe = cmul(X[PX(i)], R[i]);
            X[PX(i)] = xe;
            unsigned int rk = rank[i];
            if ((rk & 0xC000u) == 0x4000u) {
                unsigned int n = rk & 0x3fffu;
                unsigned int q = (ppos + n) % 255u;
                float pil = pilot_seq[q] ? 1.0f : -1.0f;
                yph[n] = atan2f(xe.y * pil, xe.x * pil);
            }
        }
        __syncthreads();
        PH(3);
        if (wid == 0) {
            float sy, sxy;
            warp_unwrap(yph, pilot_x, p.M_pilot, false, lane, sy, sxy);
            sy = warp_sum(sy);
            sxy = warp_sum(sxy);
            if (lane == 0) {
                const float np = (float)p.M_pilot, sx = p.pilot_sx, sxx = p.pilot_sxx;
                float den = __fsub_rn(__fmul_rn(np, sxx), __fmul_rn(sx, sx));
                float p1 = __fdiv_rn(__fsub_rn(__fmul_rn(np, sxy), __fmul_rn(sx, sy)), den);
                float p0 = __fdiv_rn(__fsub_rn(sy, __fmul_rn(p1, sx)), np);
                const float alpha = 0.3f;
                p1 = __fadd_rn(__fmul_rn(alpha, p1), __fmul_rn(1 - alpha, S->p1_prime));
                S->p1_prime = p1;
                red[111] = p0; red[112] = p1;
                if (S->num_symbols > 0) {
                    float dphi = p0 - S->phi_prime;
                    while (dphi > PI_F) dphi -= 2 * PI_F;
                    while (dphi < -PI_F) dphi += 2 * PI_F;
                    S->nco_dtheta += nco_constrain_dev(1e-3f * dphi);
                }
                S->phi_prime = p0;
                S->num_symbols++;
                S->pilot_pos = (ppos + p.M_pilot) % 255u;
                S->timer = (int)(M + cp);    // liquid sets this unconditionally (also after a reset below)
            }
        }
        __syncthreads();
        PH(4);

        // ---- derotate own subcarriers, demap the data ones
        const int fstate = S->fstate;
        const unsigned int hstart = S->header_sym_idx, pstart = S->payload_sym_idx;
        const unsigned int bps = S->bps_payload, ms = S->ms_payload, mod_len = S->payload_mod_len, enc = S->payload_enc_len;
        const unsigned int take = (fstate == FS_HEADER) ? min(p.M_data, 288u - hstart) : min(p.M_data, mod_len - pstart);
        float ev = 0.f;
        {
            const float p0 = red[111], p1 = red[112];
            const float alpha = p.qam_alpha[bps];
            for (unsigned int i = tid; i < M; i += nt) {
                unsigned int rk = rank[i];
                if (rk == 0xffffu) { X[PX(i)] = make_float2(0.f, 0.f); continue; }
                float fx = (i > M2) ? (float)i - (float)M : (float)i;
                float thv = __fadd_rn(p0, __fmul_rn(p1, fx));
                float sn, cs;
                __sincosf(thv, &sn, &cs);
                cf x = cmul(X[PX(i)], make_float2(cs, -sn));
                X[PX(i)] = x;
                if (rk < take) {
                    if (fstate == FS_HEADER) {
                        unsigned int b = x.x > 0 ? 0u : 1u;
                        sym[rk] = (uint8_t)b;
                        float dr = x.x - (b ? -1.0f : 1.0f);
                        ev += dr * dr + x.y * x.y;
                    } else {
                        sym[rk] = (uint8_t)demod_symbol(x, ms, bps, alpha);
                    }
                }
            }
        }
        if (fstate == FS_HEADER) block_sum4(ev, 0.f, 0.f, 0.f);      // two barriers inside
        else __syncthreads();
        PH(5);

        // ---- debug tap of the equalised symbol
        if (p.tap_cap) {
            if (tid == 0) red[113] = __uint_as_float(atomicAdd(&p.counters[4], 1u));
            __syncthreads();
            unsigned int slot = __float_as_uint(red[113]);
            if (slot < p.tap_cap) {
                for (unsigned int i = tid; i < M; i += nt) p.tap_X[(size_t)slot * M + i] = X[PX(i)];
                if (tid == 0) { p.tap_chan[slot] = sidx; p.tap_index[slot] = S->sample_index - 1; }
            }
        }

        // ---- ofdmflexframesync layer: pack the demapped symbols
        int emit = 0;                       // 1: header invalid, 2: payload complete
        if (fstate == FS_HEADER) {
            const unsigned int j0 = hstart >> 3, j1 = (hstart + take - 1) >> 3;
            for (unsigned int j = j0 + tid; j <= j1; j += nt) {
                unsigned int v = 0;
                for (unsigned int b = 0; b < 8; b++) {
                    unsigned int bit = 8 * j + b;
                    if (bit >= hstart && bit < hstart + take) v |= (unsigned int)sym[bit - hstart] << (7 - b);
                }
                if (8 * j < hstart) v |= S->header_bits[j];
                S->header_bits[j] = (uint8_t)v;
            }
            if (tid == 0) { S->evm_hat += red[100]; S->header_sym_idx = hstart + take; }
            if (hstart + take == 288u) {
                __syncthreads();
                // unscramble, de-interleave (n = 36, depth 4), Golay(24,12), CRC-32, parse
                uint8_t * hb = sym;                        // 36 bytes of scratch
                uint32_t * gsym = (uint32_t *)yph;         // 12 decoded Golay symbols
                if (wid == 0) {
                    const uint8_t mask[4] = {0xb4, 0x6a, 0x8b, 0x45};
                    for (unsigned int i = lane; i < 36; i += 32) hb[i] = S->header_bits[i] ^ mask[i & 3];
                    __syncwarp();
                    const uint8_t ilmask[4] = {0xff, 0x0f, 0x55, 0x33};
                    for (int v = 3; v >= 0; v--) {
                        if (lane < 18) {
                            unsigned int j = p.tb.hdr_walk[18 * v + lane];
                            uint8_t mk = ilmask[v];
                            uint8_t a = hb[2 * lane], b = hb[2 * j + 1];
                            hb[2 * lane] = (uint8_t)((a & ~mk) | (b & mk));
                            hb[2 * j + 1] = (uint8_t)((a & mk) | (b & ~mk));
                        }
                        __syncwarp();
                    }
                    if (lane < 12) {
                        unsigned int v = ((unsigned int)hb[3 * lane] << 16) | ((unsigned int)hb[3 * lane + 1] << 8) | hb[3 * lane + 2];
                        gsym[lane] = golay2412_decode(v);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        uint8_t * hd = S->header_dec;
                        for (int g = 0; g < 6; g++) {
                            unsigned int s0 = gsym[2 * g], s1 = gsym[2 * g + 1];
                            hd[3 * g] = (s0 >> 4) & 0xff;
                            hd[3 * g + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
                            hd[3 * g + 2] = s1 & 0xff;
                        }
                        uint32_t key = ((uint32_t)hd[14] << 24) | ((uint32_t)hd[15] << 16) | ((uint32_t)hd[16] << 8) | hd[17];
                        int valid = crc32_nibble(hd, 14) == key;
                        S->evm_db = 10 * log10f(S->evm_hat / 288.0f);
                        if (valid && hd[8] != 105) valid = 0;          // protocol id
                        unsigned int plen = ((unsigned int)hd[9] << 8) | hd[10];
                        unsigned int hms = hd[11], check = (hd[12] >> 5) & 7, fec0 = hd[12] & 0x1f, fec1 = hd[13] & 0x1f;
                        unsigned int hbps = dev_mod_bps(hms);
                        if (valid && (hbps == 0 || (check != 6 && check != 1) || !dev_fec_ok(fec0) || !dev_fec_ok(fec1))) valid = 0;
                        unsigned int henc = 0;
                        if (valid) {
                            henc = dev_fec_enc_len(fec1, dev_fec_enc_len(fec0, plen + (check == 6 ? 4 : 0)));
                            if (henc > p.penc_cap) valid = 0;         // cannot happen with penc_cap at its default
                        }
                        if (valid) {
                            S->ms_payload = hms; S->bps_payload = hbps; S->payload_len = plen;
                            S->check = check; S->fec0 = fec0; S->fec1 = fec1;
                            S->payload_enc_len = henc;
                            S->payload_mod_len = (8 * henc + hbps - 1) / hbps;
                            S->fstate = FS_PAYLOAD;
                        }
                        red[114] = (float)valid;
                    }
                }
                __syncthreads();
                if (red[114] != 0.f) {
                    const unsigned int henc = S->payload_enc_len;
                    for (unsigned int i = tid; i < (henc + 3) / 4; i += nt) ((uint32_t *)penc)[i] = 0u;
                } else {
                    emit = 1;
                }
            }
        } else {
            // byte j of the encoded payload collects the bits [8j, 8j+8) of the symbol stream
            // (a frame without payload symbols -- length 0, no check, no FEC -- has nothing to pack: take = 0 or enc = 0
            // would wrap j1 around and the loop below would never end; the frame completes with this symbol)
            const unsigned int bit0 = pstart * bps, nbits = take * bps;
            const bool nothing = (nbits == 0u || enc == 0u);
            unsigned int j0 = bit0 >> 3, j1 = nothing ? 0u : (bit0 + nbits - 1) >> 3;
            if (!nothing && j1 >= enc) j1 = enc - 1;
            for (unsigned int j = j0 + tid; !nothing && j <= j1; j += nt) {
                unsigned int lo = max(8 * j, bit0), hi = min(8 * j + 8, bit0 + nbits);     // bit range from this symbol
                unsigned int d0 = div_bps(lo - bit0, bps), d1 = div_bps(hi - 1 - bit0, bps);
                unsigned long long acc = 0;
                for (unsigned int d = d0; d <= d1; d++) acc = (acc << bps) | sym[d];
                // acc holds stream bits [bit0 + d0*bps, bit0 + (d1+1)*bps); keep [lo, hi)
                unsigned int top = bit0 + (d1 + 1) * bps;
                unsigned int v = (unsigned int)(acc >> (top - hi)) & ((1u << (hi - lo)) - 1u);
                v <<= (8 * j + 8 - hi);
                if (8 * j < bit0) v |= penc[j];
                penc[j] = (uint8_t)v;
            }
            if (tid == 0) S->payload_sym_idx = pstart + take;
            if (pstart + take == mod_len) emit = 2;
        }

        if (emit) {
            __syncthreads();
            // append a frame record (+ encoded payload) to the output of this launch
            if (tid == 0) {
                unsigned int slot = atomicAdd(&p.counters[0], 1u);
                unsigned long long offb = 0;
                unsigned int e2 = (emit == 2) ? S->payload_enc_len : 0u;
                int ok = slot < p.recs_cap;
                bool dropped = false;                   // the payload does not fit what is left of the arena: data, not an error
                if (ok && e2) {
                    offb = atomicAdd((unsigned long long *)(p.counters + 2), (unsigned long long)((e2 + 15u) & ~15u));
                    if (offb + e2 > p.arena_cap) { dropped = true; e2 = 0; offb = 0; atomicOr(&p.counters[1], 8u); }
                }
                if (!ok) { atomicOr(&p.counters[1], 1u); red[115] = -1.f; }
                else {
                    FrameRec r;
                    r.channel = p.chan_base + sidx;
                    r.header_valid = (emit == 2);
                    r.payload_valid = 0;
                    r.payload_len = (emit == 2 && !dropped) ? S->payload_len : 0u;
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
                    r.payload_offset = offb;
                    p.recs[slot] = r;
                    FrameAux a; a.enc_len = e2; a.sym_bps = dropped ? 0xffffffffu : 0u; a.sym_off = offb;
                    p.aux[slot] = a;
                    red[115] = dropped ? 0.f : 1.f;
                    dsum[30] = __longlong_as_double((long long)offb);
                }
            }
            __syncthreads();
            if (emit == 2 && red[115] > 0.f) {
                const unsigned long long offb = (unsigned long long)__double_as_longlong(dsum[30]);
                const unsigned int e2 = S->payload_enc_len;
                uint32_t * dst = (uint32_t *)(p.arena + offb);
                const uint32_t * src = (const uint32_t *)penc;
                for (unsigned int i = tid; i < (e2 + 3) / 4; i += nt) dst[i] = src[i];
            }
            __syncthreads();
            if (tid == 0) {
                flex_reset();
                S->timer = (int)(M + cp);    // survives the reset, as in liquid
            }
        }
    }

    // ---- store persistent state
    __syncthreads();
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

// ofdmflexframesync_reset on every stream: state machine, NCO, pilot generator and header/payload
// progress go back to their initial values; liquid leaves the sample window alone, and the
// sample counter (our side channel) keeps counting.
__global__ void sync_reset_kernel(SyncState * st, unsigned int streams, unsigned int workers, SyncCtl * ctl, unsigned long long sample_base)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= streams * workers) return;
    SyncState * S = st + i;
    S->fstate = FS_HEADER;
    S->header_sym_idx = 0;
    S->payload_sym_idx = 0;
    S->evm_hat = 0.f;
    S->nco_theta = 0; S->nco_dtheta = 0;
    S->pilot_pos = 0;
    S->timer = 0;
    S->num_symbols = 0;
    S->s_hat0_re = 0.f; S->s_hat0_im = 0.f;
    S->phi_prime = 0.f; S->p1_prime = 0.f;
    S->state = ST_SEEK;
    for (int k = 0; k < 36; k++) S->header_bits[k] = 0;      // ofdmsync8.cu ORs the header bits in
    if (workers == 2) {
        // worker 0 carries on alone from the current stream position, worker 1 waits for a hand-off
        S->role = (i & 1u) ? SW_WAIT : SW_OWNER;
        S->my_seq = 0; S->sent_seq = 0; S->verify = 0; S->pub_pending = 0;
        S->sample_index = sample_base;
        if ((i & 1u) == 0) { SyncCtl * c = ctl + (i >> 1); c->hs = (c->hs & ~3u) | HS_NONE; }
    }
}
cudaError_t sync_reset_launch(SyncState * st, unsigned int streams, unsigned int workers, SyncCtl * ctl,
                              unsigned long long sample_base, cudaStream_t stream)
{
    sync_reset_kernel<<<(streams * workers + 127) / 128, 128, 0, stream>>>(st, streams, workers, ctl, sample_base);
    return cudaGetLastError();
}

template <unsigned int MT, unsigned int NT>
static cudaError_t sync_launch_t(const SyncParams & p, int threads, size_t smem_bytes, cudaStream_t st)
{
    static size_t configured_dev[64] = {0};              // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t & configured = configured_dev[(unsigned int)dev & 63u];
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(sync_kernel<MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        configured = smem_bytes;
    }
    sync_kernel<MT, NT><<<p.streams, threads, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t sync_configure(size_t smem_bytes) { (void)smem_bytes; return cudaSuccess; }

cudaError_t sync_launch(const SyncParams & p, int threads, size_t smem_bytes, cudaStream_t st)
{
    if (p.nsamples == 0 || p.streams == 0) return cudaSuccess;
    static const bool force_generic = getenv("B2_SYNC_GENERIC") != nullptr;
    if (!force_generic && sync8_supported(p.M) && p.M_pilot + p.M_data >= 5) return sync8_launch(p, st);
    // the fixed-size instances assume the default pass plan of design.h fft_plan()
    const bool std_plan = (p.M & (p.M - 1)) == 0 && p.fft.radices == fft_static_radices(p.M);
    if (std_plan && threads == 256) {
        switch (p.M) {
        case 64:   return sync_launch_t<64, 256>(p, threads, smem_bytes, st);
        case 128:  return sync_launch_t<128, 256>(p, threads, smem_bytes, st);
        case 256:  return sync_launch_t<256, 256>(p, threads, smem_bytes, st);
        case 512:  return sync_launch_t<512, 256>(p, threads, smem_bytes, st);
        case 1024: return sync_launch_t<1024, 256>(p, threads, smem_bytes, st);
        default: break;
        }
    } else if (std_plan && threads == 128) {
        switch (p.M) {
        case 64:   return sync_launch_t<64, 128>(p, threads, smem_bytes, st);
        case 128:  return sync_launch_t<128, 128>(p, threads, smem_bytes, st);
        case 256:  return sync_launch_t<256, 128>(p, threads, smem_bytes, st);
        default: break;
        }
    }
    return sync_launch_t<0, 0>(p, threads, smem_bytes, st);
}

} // namespace b2
