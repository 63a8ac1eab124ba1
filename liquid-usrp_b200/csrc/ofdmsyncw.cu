// ofdmsyncw.cu -- frame-parallel OFDM frame synchroniser: one WARP per worker, several workers per stream.
//
// Same job as ofdmsync.cu / ofdmsync8.cu (liquid's ofdmframesync + ofdmflexframesync state machines, reached by
// the reference through
//     ofdmflexframesync_execute(framesync[i], &X[i], 1)        lib/multichannelrx.cc:194
//     ofdmflexframesync_execute(fs, &sample, 1)                lib/ofdmtxrx.cc:625 )
// and the same results, but organised for THROUGHPUT instead of for the latency of one serial chain.  The
// reference's own note at lib/multichannelrx.cc:184,193-194 ("TODO: run each channel in its own thread") stops at
// channels; a stream of back-to-back frames is still a serial recurrence there (the NCO is trimmed from each OFDM
// symbol's pilot phase before the next symbol is mixed, the seek grid restarts where a frame ends).  What IS
// independent is everything that happens after a point where the synchroniser is in its canonical idle state:
// state SEEK, timer 0, NCO / pilot generator / header and payload progress reset -- the state right after a seek
// event that detected nothing.  liquid is in that state one sample behind every frame and on every idle seek-grid
// point, and from there on the future depends on the input samples only.  So:
//
//   * a launch gives every stream `workers` workers (warps).  Worker 0 carries the stream's true state on from the
//     previous launch.  Worker w >= 1 starts at a PREDICTED canonical position P_w (frames arriving with the
//     period of the last two frame ends, or the idle seek grid; WChan) in the canonical state, and every worker
//     runs up to the next worker's start;
//   * a worker that arrives at P_{w+1} having just run a seek event there that detected nothing WAS in the
//     canonical state at P_{w+1}: the next worker's records are exactly the serial chain's.  The check is an
//     equality of complete states, so a wrong prediction can never change a result;
//   * the last worker of a stream to finish (atomic counter, nobody waits) stitches: it walks the workers in
//     order, commits the private record lists of the verified ones, and where a worker did not arrive in the
//     canonical state it simply carries on serially from that worker's (true) final state -- until it meets a
//     later worker's start in the canonical state, or the launch's samples end.
//
// The kernel keeps NO sample window: an event's M-sample FFT window is re-read from the raw stream (global memory /
// L2; the previous launch's last M + cp samples are kept in `ring`) and re-mixed with the NCO phase those samples
// had when liquid pushed them (WSync::mix_*), so the cyclic prefix is never even loaded and per-worker state is a
// few scalars plus the equaliser taps.
//
// A worker is ONE warp (M / 32 subcarriers per lane: Stockham radix-8 passes of fft8.cuh, two virtual threads per
// lane at M = 512), all barriers are __syncwarp, all reductions shuffles, the state machine runs redundantly in
// every lane (no broadcasts).  ~8.7 KB of shared memory per worker at M = 512: 16 workers per SM.
#include "kernels.h"
#include "fec.cuh"
#include "syncdev.cuh"
#include "fft8.cuh"
#include <cstdlib>

namespace b2 {

struct SWLayout {
    size_t twt_bytes, per_warp, off_f, off_rg, off_yc, off_ws, off_hb, off_ctl, total;
};
__host__ __device__ static inline SWLayout sw_layout(unsigned int M, unsigned int Mp, unsigned int wpc)
{
    SWLayout L;
    // per-pass twiddle tables of the M-point transform and, for M = 512, of the M/2-point one (S0 events)
    L.twt_bytes = ((size_t)(f8_twt_elems(M, 1) + (M == 512 ? f8_twt_elems(M / 2, 1) : 0)) * sizeof(cf) + 15) & ~(size_t)15;
    size_t o = 0;
    L.off_f = o;  o += (size_t)(M + M / 8) * sizeof(cf);  // FFT exchange buffer (skewed); Gs / yph between transforms
    L.off_rg = o; o += (size_t)M * sizeof(cf);            // equaliser taps R (state RX) / S0a gains (state S0B)
    L.off_yc = o; o += ((size_t)(Mp + 4) * sizeof(cf) + 15) & ~(size_t)15;
    L.off_ws = o; o += (sizeof(WSync) + 15) & ~(size_t)15;
    L.off_hb = o; o += 112;                                // 36 header bytes + 12 Golay symbols + decode results
    L.off_ctl = o; o += 64;                               // stretch / stitch bookkeeping (see WCtl)
    L.per_warp = (o + 15) & ~(size_t)15;
    L.total = L.twt_bytes + L.per_warp * wpc;
    return L;
}

// per-worker bookkeeping that is only touched between stretches (shared memory, one per warp)
struct WCtl {
    unsigned long long chain_b_last, chain_b_prev;   // stitcher: last two frame ends of the verified chain
    unsigned int n_act, J, period, head;
    int rfirst;                                      // first predicted position, relative to the launch's first sample
    unsigned int cur, sc, pbase, pcap;
    unsigned int pad[3];
};
static_assert(sizeof(WCtl) == 64, "WCtl is 64 bytes");

// Exchange between two radix-8 passes through a SKEWED buffer, element x at x + (x >> 3): the scattered stores of a
// pass (element b0 + q Ns, q < 8) and the strided loads of the next one (element j + s N/8) are then both one base
// address per virtual thread plus compile-time offsets, and both are bank-conflict free for a warp of consecutive j
// (Ns = 1: 9 j + q; Ns = 8: 9 (j - k) + k + 9 q; loads: j + (j >> 3) + s (N/8 + N/64)).
template <unsigned int N, unsigned int Ns>
__device__ __forceinline__ void fw_store(const cf (&v)[8], unsigned int j, cf * __restrict__ buf)
{
    static_assert(Ns == 1 || Ns % 8 == 0, "radix-8 passes only");
    const unsigned int k = j & (Ns - 1), b0 = (j - k) * 8 + k;
    cf * dst = buf + (b0 + (b0 >> 3));
#pragma unroll
    for (unsigned int q = 0; q < 8; q++) dst[q * (Ns + Ns / 8)] = v[q];
}
template <unsigned int N>
__device__ __forceinline__ void fw_load(cf (&v)[8], unsigned int j, const cf * __restrict__ buf)
{
    const cf * src = buf + (j + (j >> 3));
#pragma unroll
    for (unsigned int s = 0; s < 8; s++) v[s] = src[s * (N / 8 + N / 64)];
}
// per-pass tables (f8_twt_build) of an N-point transform out of the twiddles of the 2N-point one: e^{-2 pi i k / N} = tw[2k]
template <unsigned int N, unsigned int Ns>
__device__ __forceinline__ void fw_twt_build_half(cf * twt, const cf * __restrict__ tw2, unsigned int tid, unsigned int nthreads)
{
    constexpr unsigned int R = f8_radix(N, Ns);
    if constexpr (Ns > 1) {
        for (unsigned int e = tid; e < (R - 1) * Ns; e += nthreads) {
            const unsigned int q = e / Ns + 1, k = e % Ns;
            twt[e] = tw2[2u * (k * q * (N / (Ns * R)))];
        }
    }
    if constexpr (Ns * R < N) fw_twt_build_half<N, Ns * R>(twt + (Ns > 1 ? (R - 1) * Ns : 0), tw2, tid, nthreads);
}

// passes of the M-point transform from Ns on, VT virtual threads (lane + 32 vt) per lane
template <unsigned int N, unsigned int Ns, unsigned int VT>
__device__ __forceinline__ void fw_run(cf (&v)[VT][8], unsigned int lane, cf * __restrict__ buf, const cf * __restrict__ twt)
{
    constexpr unsigned int R = f8_radix(N, Ns);
#pragma unroll
    for (unsigned int vt = 0; vt < VT; vt++) f8_pass<N, Ns, R, -1>(v[vt], lane + 32 * vt, nullptr, nullptr, twt);
    if constexpr (Ns * R < N) {
        static_assert(R == 8, "only the last pass may have a smaller radix");
#pragma unroll
        for (unsigned int vt = 0; vt < VT; vt++) fw_store<N, Ns>(v[vt], lane + 32 * vt, buf);
        __syncwarp();
#pragma unroll
        for (unsigned int vt = 0; vt < VT; vt++) fw_load<N>(v[vt], lane + 32 * vt, buf);
        __syncwarp();
        fw_run<N, Ns * R, VT>(v, lane, buf, twt + (Ns > 1 ? (R - 1) * Ns : 0));
    }
}

__device__ __forceinline__ float shfl0(float v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ unsigned int shfl0(unsigned int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ unsigned long long shfl0(unsigned long long v) { return __shfl_sync(0xffffffffu, v, 0); }

#define B2W_FORPTS                                                         \
    _Pragma("unroll") for (unsigned int vt = 0; vt < VT; vt++)             \
    _Pragma("unroll") for (unsigned int s = 0; s < 8; s++)
#define B2W_I (lane + 32u * vt + T * s)
// the W samples behind the current window (the next symbol's window, or the next seek window) into L1: one 128-byte
// line per lane and instruction
#define B2W_PREFETCH_NEXT                                                                                     \
    if (rel0 + (long long)(2 * M + cp) <= (long long)p.nsamples) {                                            \
        const char * pf = (const char *)(in + rel0 + M) + 128u * lane;                                        \
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));                                                   \
        if (4096u + 128u * lane < W * 8u) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 4096));          \
    }

template <unsigned int M, unsigned int WPC, unsigned int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) syncw_kernel(const SyncParams p)
{
    constexpr unsigned int VT = M / 256, T = M / 8, M2 = M / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int cp = p.cp, W = M + cp, Mp = p.M_pilot, Na = Mp + p.M_data;
    const SWLayout L = sw_layout(M, Mp, WPC);
    cf * twt = (cf *)smem;
    unsigned char * wb = smem + L.twt_bytes + (size_t)wid * L.per_warp;
    cf * fbuf = (cf *)(wb + L.off_f);
    cf * RG = (cf *)(wb + L.off_rg);
    cf * yc = (cf *)(wb + L.off_yc);
    WSync * S = (WSync *)(wb + L.off_ws);
    uint8_t * hb = (uint8_t *)(wb + L.off_hb);            // 36 bytes
    uint32_t * gsym = (uint32_t *)(wb + L.off_hb + 40);   // 12 words
    uint32_t * hres = (uint32_t *)(wb + L.off_hb + 88);   // header decode verdict
    cf * Gs = fbuf;
    float * yph = (float *)fbuf;

    f8_twt_build<M, 1>(twt, p.fft.tw, threadIdx.x, WPC * 32);
    cf * twt_h = twt + f8_twt_elems(M, 1);
    if constexpr (M == 512) fw_twt_build_half<M / 2, 1>(twt_h, p.fft.tw, threadIdx.x, WPC * 32);
    __syncthreads();

    // ---- which stream, which worker, which stretch.  Positions inside the kernel are 32-bit offsets from the
    //      launch's first sample (E0 = p.sample_base): a launch is shorter than 2^31 samples per stream
    constexpr int RNONE = 0x7fffffff;                 // "open" / "no such position"
    WCtl * ctl = (WCtl *)(wb + L.off_ctl);
    const unsigned int gw = blockIdx.x * WPC + wid;
    const unsigned int ch = gw % p.streams, w = gw / p.streams;
    if (w >= p.workers) return;
    WChan * C = p.wch + ch;
    const unsigned int par = p.launch_id & 1u;
    {
        const unsigned long long E0 = p.sample_base, E1 = E0 + p.nsamples;
        unsigned long long first = 0;
        unsigned int J = 0;
        const unsigned int period = C->pred_period[par];
        if (p.workers > 1 && period >= 64u) {
            first = C->pred_next[par];
            if (first <= E0) first += ((E0 - first) / period + 1ull) * period;
            if (first + M <= E1) J = (unsigned int)min((E1 - M - first) / period + 1ull, (unsigned long long)(1u << 24));
        }
        const unsigned int n_act = min(p.workers, J + 1u);
        if (w >= n_act) return;
        if (lane == 0) {
            ctl->n_act = n_act; ctl->J = J; ctl->period = period; ctl->head = C->head[par];
            ctl->rfirst = J ? (int)(first - E0) : 0;
        }
        __syncwarp();
    }
    // start of worker ww's stretch (ww < 64, J <= 2^24)
    auto rstart = [&](unsigned int ww) -> int {
        if (ww == 0) return 0;
        if (ww >= ctl->n_act) return RNONE;
        return ctl->rfirst + (int)(((ww * (ctl->J + 1u)) / ctl->n_act - 1u) * ctl->period);
    };
    auto rend = [&](unsigned int ww) -> int { return (ww + 1 < ctl->n_act) ? rstart(ww + 1) : (int)p.nsamples; };
    auto slot_of = [&](unsigned int ww) -> unsigned int { return ch * p.wslots + (ctl->head + ww) % p.wslots; };
    // private record list of speculative worker ww: disjoint regions of the stream's list (a record needs > 2 W samples)
    auto priv_base = [&](unsigned int ww) -> unsigned int {
        return ch * p.wrec_stride + (unsigned int)rstart(ww) / (2u * W) + 2u * ww;
    };
    auto priv_cap = [&](unsigned int ww) -> unsigned int { return (unsigned int)(rend(ww) - rstart(ww)) / (2u * W) + 2u; };

    const cf * in = p.in + (size_t)ch * p.in_stride;
    const cf * ring = p.ring + (size_t)ch * W;

    // role / rank of own subcarriers: read where they are used (64-byte rows of an L1-resident table)
#define B2W_RK ((unsigned int)__ldg(&p.tb.sc_rank[B2W_I]))

    // ---- hot state in registers (identical in every lane); everything an event does not touch lives in S (shared
    //      memory; every lane writes the same value, or lane 0 writes and a __syncwarp follows)
    int state = ST_SEEK, timer = 0, fstate = FS_HEADER;
    unsigned int num_symbols = 0, th = 0, dth = 0, pilot_pos = 0, pstart = 0;
    float phi_prime = 0.f, p1_prime = 0.f;
    int rs = 0;                                    // next sample to be pushed
    int rms = 0, rme = 0;                          // mixed segment [rms, rme), rme == RNONE: open
    unsigned long long sym_off = 0;

    auto rel_of = [&](unsigned long long x) -> int {
        const unsigned long long E0 = p.sample_base;
        if (x == ~0ull) return RNONE;
        if (x >= E0) return (int)min(x - E0, (unsigned long long)(1u << 30));
        return -(int)min(E0 - x, (unsigned long long)(1u << 30));
    };
    auto load_state = [&](unsigned int slot) {
        const uint32_t * src = (const uint32_t *)(p.wst + slot);
        uint32_t * dst = (uint32_t *)S;
        __syncwarp();
        for (unsigned int i = lane; i < sizeof(WSync) / 4; i += 32) dst[i] = __ldcg(src + i);
        __syncwarp();
        state = S->state; timer = S->timer; fstate = S->fstate;
        num_symbols = S->num_symbols; th = S->nco_theta; dth = S->nco_dtheta; pilot_pos = S->pilot_pos;
        pstart = S->payload_sym_idx;
        phi_prime = S->phi_prime; p1_prime = S->p1_prime;
        rs = rel_of(S->sample_index); rms = rel_of(S->mix_start); rme = rel_of(S->mix_end); sym_off = S->sym_off;
        __syncwarp();
        if (lane == 0) { S->nb = 0; S->b_last = 0; S->b_prev = 0; S->nrec = 0; }
        if (state == ST_RX || state == ST_S0B) {
            const cf * g = p.wRG + (size_t)slot * M;
            for (unsigned int i = lane; i < M; i += 32) RG[i] = __ldcg(g + i);
        }
        __syncwarp();
    };
    auto fresh_state = [&](int at) {      // (S all zero: S->g0 is set by the seek event that always comes first)
        uint32_t * dst = (uint32_t *)S;
        __syncwarp();
        for (unsigned int i = lane; i < sizeof(WSync) / 4; i += 32) dst[i] = 0u;
        __syncwarp();
        state = ST_SEEK; timer = 0; fstate = FS_HEADER;
        num_symbols = 0; th = 0; dth = 0; pilot_pos = 0; pstart = 0;
        phi_prime = 0.f; p1_prime = 0.f;
        rs = at; rms = 0; rme = 0; sym_off = 0;
    };
    auto save_state = [&](unsigned int slot, unsigned int matched) {
        __syncwarp();
        if (lane == 0) {
            const unsigned long long E0 = p.sample_base;
            S->state = state; S->timer = timer; S->fstate = fstate;
            S->num_symbols = num_symbols; S->nco_theta = th; S->nco_dtheta = dth; S->pilot_pos = pilot_pos;
            S->payload_sym_idx = pstart;
            S->phi_prime = phi_prime; S->p1_prime = p1_prime;
            S->sample_index = E0 + (unsigned long long)rs;
            const bool none = (rme <= rms);
            S->mix_start = none ? 0ull : E0 + (long long)rms;
            S->mix_end = none ? 0ull : (rme == RNONE ? ~0ull : E0 + (long long)rme);
            S->sym_off = sym_off;
            S->matched = matched;
        }
        __syncwarp();
        uint32_t * dst = (uint32_t *)(p.wst + slot);
        const uint32_t * src = (const uint32_t *)S;
        for (unsigned int i = lane; i < sizeof(WSync) / 4; i += 32) dst[i] = src[i];
        if (state == ST_RX || state == ST_S0B) {
            cf * g = p.wRG + (size_t)slot * M;
            for (unsigned int i = lane; i < M; i += 32) g[i] = RG[i];
        }
    };

    // ---- stretch control
    // ctl->cur: the worker whose slot this warp is working in; ctl->sc: stitch cursor
    bool direct = (w == 0);                        // records go straight to the launch's output (known to be the serial chain's)
    bool stitching = false;
    int rlimit = rend(w);
    unsigned int cand = w + 1;                     // next worker whose start may be met in the canonical state
    int rcand = rstart(cand);
    if (lane == 0) { ctl->cur = w; ctl->sc = 0; ctl->pbase = priv_base(w); ctl->pcap = priv_cap(w); }
    if (w == 0) load_state(slot_of(0)); else fresh_state(rstart(w));
    __syncwarp();

    for (;;) {
        unsigned int matched = 0;
        // ================================================================ event loop of one stretch
        for (;;) {
            unsigned int need;
            if (state == ST_SEEK) need = (timer < (int)M) ? (unsigned int)((int)M - timer) : 1u;
            else if (state == ST_S0A || state == ST_S0B) need = (timer < (int)M2) ? (unsigned int)((int)M2 - timer) : 1u;
            else need = (timer > 1) ? (unsigned int)timer : 1u;
            if (rs + (int)need > rlimit) {         // the event lies beyond this stretch: consume what is left
                const unsigned int adv = (unsigned int)(rlimit - rs);
                if (state != ST_SEEK) th += adv * dth;
                if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) timer += (int)adv; else timer -= (int)adv;
                rs = rlimit;
                break;
            }
            const int re = rs + (int)need;
            const unsigned int off = (state == ST_RX) ? cp - p.backoff : cp;
            const int rel0 = re - (int)W + (int)off;                    // first sample of the FFT window (may lie before the launch)
            // ---- the window, re-mixed as liquid pushed it
            cf v[VT][8];
            {
                const bool open = (rme == RNONE);
                const unsigned int r_th = open ? th : S->q_theta, r_dth = open ? dth : S->q_dtheta;
                const bool any_mixed = (rme > rms) && (rms - rel0 < (int)M) && (open || rme - rel0 > 0) && ((r_th | r_dth) != 0u);
                const bool all_mixed = any_mixed && (rms <= rel0) && (open || rme - rel0 >= (int)M);
                const unsigned int d0 = (unsigned int)(rel0 - (open ? rs : rme));
                if (rel0 >= 0 && all_mixed) {
                    const cf * src = in + rel0 + lane;
                    const unsigned int ph0 = r_th + (d0 + lane) * r_dth, ph32 = 32u * r_dth;
                    B2W_FORPTS {
                        const unsigned int q = vt + VT * s;
                        v[vt][s] = mix_down(__ldg(src + 32u * q), nco_cexp_fast(ph0 + q * ph32));
                    }
                    // the samples behind this window (the next symbol's, or the next seek window): into L1 now
                    B2W_PREFETCH_NEXT
                } else if (rel0 >= 0 && !any_mixed) {
                    const cf * src = in + rel0 + lane;
                    B2W_FORPTS { v[vt][s] = __ldg(src + 32u * (vt + VT * s)); }
                    B2W_PREFETCH_NEXT
                } else {
                    const int m0 = rms - rel0, m1 = open ? (int)M : rme - rel0;
#pragma unroll 4
                    for (unsigned int q = 0; q < 8 * VT; q++) {
                        const unsigned int i = lane + 32u * q;
                        const int r = rel0 + (int)i;
                        cf x = (r >= 0) ? __ldg(in + r) : __ldcg(ring + ((int)W + r));
                        if (any_mixed && (int)i >= m0 && (int)i < m1) x = mix_down(x, nco_cexp_fast(r_th + (d0 + i) * r_dth));
                        fbuf[i] = x;
                    }
                    __syncwarp();
                    B2W_FORPTS { v[vt][s] = fbuf[B2W_I]; }
                }
            }
            // ---- advance to the event
            if (state != ST_SEEK) th += need * dth;
            if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) timer += (int)need; else timer -= (int)need;
            rs = re;
            float en = 0.f;
            if (state == ST_SEEK) {
                B2W_FORPTS { en += v[vt][s].x * v[vt][s].x + v[vt][s].y * v[vt][s].y; }
            }
            // ---- M-point forward FFT.  The events on the short training symbol (seek, S0a, S0b) only look at the even
            //      subcarriers, X[2k] = FFT_{M/2}(x[n] + x[n + M/2])[k]: half the transform, half the correlator
            __syncwarp();
            float mr = 0.f, mi = 0.f, cr = 0.f, ci = 0.f;
            if (VT == 2 && state != ST_RX && state != ST_S1) {
                if constexpr (VT == 2) {
                    cf wv[1][8];
#pragma unroll
                    for (unsigned int q = 0; q < 8; q++) wv[0][q] = cadd(v[q & 1][q >> 1], v[q & 1][(q >> 1) + 4]);
                    fw_run<M / 2, 1, 1>(wv, lane, fbuf, twt_h);
                    __syncwarp();
                    // G[2k] = X[2k] ref[2k] gain, k = lane + 32 q
                    const float gain = sqrtf((float)p.M_S0) / (float)M;
#pragma unroll
                    for (unsigned int q = 0; q < 8; q++) {
                        const unsigned int k = lane + 32u * q;
                        const float r = __ldg(p.tb.S0 + 2u * k);
                        wv[0][q] = make_float2(wv[0][q].x * r * gain, wv[0][q].y * r * gain);
                        Gs[k] = wv[0][q];
                    }
                    __syncwarp();
#pragma unroll
                    for (unsigned int q = 0; q < 8; q++) {
                        const cf tt = cmulc(Gs[(lane + 32u * q + 1u) & (M2 - 1)], wv[0][q]);
                        mr += tt.x; mi += tt.y;
                    }
                    if (state == ST_S0A) {
#pragma unroll
                        for (unsigned int q = 0; q < 8; q++) RG[lane + 32u * q] = wv[0][q];
                    } else if (state == ST_S0B) {
#pragma unroll
                        for (unsigned int q = 0; q < 8; q++) { const cf tt = cmulc(wv[0][q], RG[lane + 32u * q]); cr += tt.x; ci += tt.y; }
                    }
                }
            } else {
                fw_run<M, 1, VT>(v, lane, fbuf, twt);
                __syncwarp();                      // fbuf is free again (Gs / yph)
                if (state != ST_RX) {
                    // ---- G[i] = X[i]*ref[i]*gain on the training subcarriers
                    const bool long_seq = (state == ST_S1);
                    const unsigned int step = long_seq ? 1u : 2u;
                    const float gain = sqrtf((float)(long_seq ? p.M_S1 : p.M_S0)) / (float)M;
                    const float * ref = long_seq ? p.tb.S1 : p.tb.S0;
                    B2W_FORPTS {
                        const unsigned int i = B2W_I;
                        const float r = __ldg(ref + i);
                        v[vt][s] = make_float2(v[vt][s].x * r * gain, v[vt][s].y * r * gain);
                        Gs[i] = v[vt][s];
                    }
                    __syncwarp();
                    B2W_FORPTS {
                        const unsigned int i = B2W_I;
                        const cf tt = cmulc(Gs[(i + step) & (M - 1)], v[vt][s]);
                        mr += tt.x; mi += tt.y;
                    }
                    if (state == ST_S0A) {
                        B2W_FORPTS { RG[B2W_I] = v[vt][s]; }
                    } else if (state == ST_S0B) {
                        B2W_FORPTS { const cf tt = cmulc(v[vt][s], RG[B2W_I]); cr += tt.x; ci += tt.y; }
                    }
                }
            }

            if (state != ST_RX) {
                // ---- preamble events: s_hat = sum G[i+step] conj(G[i]) / cross-correlation of the two S0 halves
                if (state == ST_SEEK) cr = en;
                mr = warp_sum(mr); mi = warp_sum(mi); cr = warp_sum(cr); ci = warp_sum(ci);
                __syncwarp();                      // everybody has read Gs (yph aliases it)
                if (state == ST_SEEK) {
                    const float gg = (float)M / cr;
                    const cf s_hat = make_float2(mr / (float)p.M_S0 * gg, mi / (float)p.M_S0 * gg);
                    S->g0 = gg;
                    timer = 0;
                    if (hypotf(s_hat.x, s_hat.y) > p.thresh) {
                        const float tau_hat = atan2f(s_hat.y, s_hat.x) * (float)M2 / (2 * PI_F);
                        const int dt = (int)roundf(tau_hat);
                        timer = (int)((M + (unsigned int)dt) % M2) + (int)M;
                        state = ST_S0A;
                        if (lane == 0) S->detect_index = p.sample_base + (unsigned long long)rs - 1ull;
                    } else {
                        rms = 0; rme = 0;                     // canonical: nothing behind this point matters any more
                    }
                } else if (state == ST_S0A) {
                    timer = 0;
                    S->s_hat0_re = mr / (float)p.M_S0 * S->g0;
                    S->s_hat0_im = mi / (float)p.M_S0 * S->g0;
                    state = ST_S0B;
                } else if (state == ST_S0B) {
                    const float s1r = mr / (float)p.M_S0 * S->g0, s1i = mi / (float)p.M_S0 * S->g0;
                    const float tau_hat = atan2f(S->s_hat0_im + s1i, S->s_hat0_re + s1r) * (float)M2 / (2 * PI_F);
                    timer = (int)(M + cp - p.backoff) - (int)roundf(tau_hat);
                    const float nu_hat = 2.0f * atan2f(ci, cr) / (float)M;
                    dth = nco_constrain_dev(nu_hat);
                    state = ST_S1;
                    rms = rs; rme = RNONE;                    // from here on samples are pushed through the NCO
                } else {
                    // ---- S1: accept / retry, and on accept the equaliser
                    num_symbols++;
                    cf s_hat = make_float2(mr / (float)p.M_S1 * S->g0, mi / (float)p.M_S1 * S->g0);
                    s_hat = cmul(s_hat, make_float2(p.b_cos, p.b_sin));
                    const bool accept = (s_hat.x * s_hat.x + s_hat.y * s_hat.y > p.thresh * p.thresh) &&
                                        (s_hat.x > 0.f) && (fabsf(s_hat.y) < 0.32491969623290632616f * s_hat.x);
                    if (!accept) {
                        if (num_symbols == 16) {              // ofdmframesync_reset
                            th = 0; dth = 0; pilot_pos = 0; timer = 0; num_symbols = 0;
                            S->s_hat0_re = 0.f; S->s_hat0_im = 0.f; phi_prime = 0.f; p1_prime = 0.f;
                            state = ST_SEEK;
                            rms = 0; rme = 0;
                        } else timer = (int)M2;
                    } else {
                        // G *= M/sqrt(Na) * B ; smooth |G| and arg G with an order-4 polynomial over the active
                        // subcarriers (liquid ofdmframesync_estimate_eqgain_poly); R = B / G.  coef = P y with the
                        // constant matrix P of design.h eqgain_fit_matrix().
                        const float gsc = (float)M / sqrtf((float)Na);
                        float ya[VT][8];
                        B2W_FORPTS {
                            const unsigned int i = B2W_I;
                            const unsigned int ar = __ldg(&p.tb.act_rank[i]);
                            ya[vt][s] = 0.f;
                            if (ar != 0xffffu) {
                                const cf gk = cmul(cscale(v[vt][s], gsc), __ldg(&p.tb.B[i]));
                                ya[vt][s] = sqrtf(gk.x * gk.x + gk.y * gk.y);
                                yph[ar] = atan2_fast(gk.y, gk.x);
                            }
                        }
                        __syncwarp();
                        {
                            int wraps = 0;
                            for (unsigned int n = lane + 1; n < Na; n += 32) wraps |= fabsf(yph[n] - yph[n - 1]) > PI_F;
                            if (__any_sync(0xffffffffu, wraps)) {
                                warp_unwrap_seg(yph, Na, lane);
                                __syncwarp();
                            }
                        }
                        double ca[10];
#pragma unroll
                        for (int i = 0; i < 10; i++) ca[i] = 0.0;
                        B2W_FORPTS {
                            const unsigned int ar = __ldg(&p.tb.act_rank[B2W_I]);
                            if (ar != 0xffffu) {
                                const double yad = (double)ya[vt][s], yg = (double)yph[ar];
                                const double * pr = p.tb.eqfit_P + (size_t)ar * 5;
#pragma unroll
                                for (int r = 0; r < 5; r++) {
                                    const double pv = __ldg(pr + r);
                                    ca[r] = fma(pv, yad, ca[r]);
                                    ca[5 + r] = fma(pv, yg, ca[5 + r]);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 10; i++) ca[i] = warp_sum_d(ca[i]);
                        B2W_FORPTS {
                            const unsigned int i = B2W_I;
                            cf Rv = make_float2(0.f, 0.f);
                            if (B2W_RK != 0xffffu) {
                                const float fx = (i > M2) ? (float)i - (float)M : (float)i;
                                const float xv = fx / (float)M;
                                const float A = fmaf(fmaf(fmaf(fmaf((float)ca[4], xv, (float)ca[3]), xv, (float)ca[2]), xv, (float)ca[1]), xv, (float)ca[0]);
                                float thv = fmaf(fmaf(fmaf(fmaf((float)ca[9], xv, (float)ca[8]), xv, (float)ca[7]), xv, (float)ca[6]), xv, (float)ca[5]);
                                thv = fmaf(-6.28318530717958647692f, rintf(thv * 0.15915494309189533577f), thv);
                                float sn, cs;
                                __sincosf(thv, &sn, &cs);
                                const float inv = __frcp_rn(A);
                                const cf num = cmulc(__ldg(&p.tb.B[i]), make_float2(cs, sn));
                                Rv = make_float2(num.x * inv, num.y * inv);
                            }
                            RG[i] = Rv;
                        }
                        state = ST_RX;
                        timer = (int)(M + cp + p.backoff);
                        num_symbols = 0;
                    }
                }
            } else {
                // ================================================================ ST_RX: one OFDM symbol
                B2W_FORPTS {
                    const unsigned int rk = B2W_RK;
                    v[vt][s] = cmul(v[vt][s], RG[B2W_I]);
                    if ((rk & 0xC000u) == 0x4000u) yc[rk & 0x3fffu] = v[vt][s];
                }
                __syncwarp();
                float sy, sxy;
                if (Mp <= 64) {
                    // at most two pilots per lane (n = lane, lane + 32): phases, unwrap and sums stay in registers
                    const bool v0 = lane < Mp, v1 = lane + 32 < Mp;
                    float raw0 = 0.f, raw1 = 0.f, px0 = 0.f, px1 = 0.f;
                    if (v0) {
                        const float pil = __ldg(&p.tb.pilot_seq[(pilot_pos + lane) % 255u]) ? 1.0f : -1.0f;
                        const cf c = yc[lane];
                        raw0 = atan2_fast(c.y * pil, c.x * pil);
                        px0 = __ldg(&p.tb.pilot_x[lane]);
                    }
                    if (v1) {
                        const float pil = __ldg(&p.tb.pilot_seq[(pilot_pos + lane + 32) % 255u]) ? 1.0f : -1.0f;
                        const cf c = yc[lane + 32];
                        raw1 = atan2_fast(c.y * pil, c.x * pil);
                        px1 = __ldg(&p.tb.pilot_x[lane + 32]);
                    }
                    float prev0 = __shfl_up_sync(0xffffffffu, raw0, 1);
                    float prev1 = __shfl_up_sync(0xffffffffu, raw1, 1);
                    const float last0 = __shfl_sync(0xffffffffu, raw0, 31);
                    if (lane == 0) prev1 = last0;
                    int k0 = 0, k1 = 0;
                    if (v0 && lane > 0) { const float d = raw0 - prev0; k0 = (d > PI_F) ? -1 : ((d < -PI_F) ? 1 : 0); }
                    if (v1) { const float d = raw1 - prev1; k1 = (d > PI_F) ? -1 : ((d < -PI_F) ? 1 : 0); }
                    if (__any_sync(0xffffffffu, (k0 | k1) != 0)) {        // rare: most symbols need no unwrapping
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int a = __shfl_up_sync(0xffffffffu, k0, o), b = __shfl_up_sync(0xffffffffu, k1, o);
                            if (lane >= (unsigned int)o) { k0 += a; k1 += b; }
                        }
                        k1 += __shfl_sync(0xffffffffu, k0, 31);
                    }
                    float yy0 = raw0, yy1 = raw1;
                    for (int q = k0; q > 0; q--) yy0 += 2 * PI_F;
                    for (int q = k0; q < 0; q++) yy0 -= 2 * PI_F;
                    for (int q = k1; q > 0; q--) yy1 += 2 * PI_F;
                    for (int q = k1; q < 0; q++) yy1 -= 2 * PI_F;
                    sy = (v0 ? yy0 : 0.f) + (v1 ? yy1 : 0.f);
                    sxy = (v0 ? px0 * yy0 : 0.f) + (v1 ? px1 * yy1 : 0.f);
                } else {
                    for (unsigned int n = lane; n < Mp; n += 32) {
                        const float pil = __ldg(&p.tb.pilot_seq[(pilot_pos + n) % 255u]) ? 1.0f : -1.0f;
                        const cf c = yc[n];
                        yph[n] = atan2_fast(c.y * pil, c.x * pil);
                    }
                    __syncwarp();
                    warp_unwrap(yph, p.tb.pilot_x, Mp, false, lane, sy, sxy);
                    __syncwarp();
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sy += __shfl_xor_sync(0xffffffffu, sy, o);
                    sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
                }
                float fit_p0, p1;
                const unsigned int dth_old = dth;
                {
                    const float np = (float)Mp, sx = p.pilot_sx, sxx = p.pilot_sxx;
                    const float den = __fsub_rn(__fmul_rn(np, sxx), __fmul_rn(sx, sx));
                    p1 = __fdiv_rn(__fsub_rn(__fmul_rn(np, sxy), __fmul_rn(sx, sy)), den);
                    fit_p0 = __fdiv_rn(__fsub_rn(sy, __fmul_rn(p1, sx)), np);
                    const float alpha = 0.3f;
                    p1 = __fadd_rn(__fmul_rn(alpha, p1), __fmul_rn(1 - alpha, p1_prime));
                    // NCO trim (the next symbol is mixed with it)
                    if (num_symbols > 0) {
                        float dphi = fit_p0 - phi_prime;
                        while (dphi > PI_F) dphi -= 2 * PI_F;
                        while (dphi < -PI_F) dphi += 2 * PI_F;
                        dth += nco_constrain_small(1e-3f * dphi);
                    }
                }
                phi_prime = fit_p0; p1_prime = p1;
                num_symbols++;
                pilot_pos = (pilot_pos + Mp) % 255u;
                timer = (int)W;                    // liquid sets this unconditionally (also after a reset below)
                const int seg_start = rms;
                rms = rs; rme = RNONE;             // the next symbol is mixed with the trimmed NCO

                // ---- derotate own subcarriers (null subcarriers carry 0: their equaliser tap is 0)
                B2W_FORPTS {
                    const unsigned int i = B2W_I;
                    const float fx = (i > M2) ? (float)i - (float)M : (float)i;
                    const float thv = __fadd_rn(fit_p0, __fmul_rn(p1, fx));
                    float sn, cs;
                    __sincosf(thv, &sn, &cs);
                    v[vt][s] = cmul(v[vt][s], make_float2(cs, -sn));
                }
                // ---- debug tap of the equalised symbol (the host runs one worker per stream when it is on)
                if (p.tap_cap) {
                    unsigned int slot = 0;
                    if (lane == 0) slot = atomicAdd(&p.counters[4], 1u);
                    slot = shfl0(slot);
                    if (slot < p.tap_cap) {
                        B2W_FORPTS { p.tap_X[(size_t)slot * M + B2W_I] = v[vt][s]; }
                        if (lane == 0) { p.tap_chan[slot] = ch; p.tap_index[slot] = p.sample_base + (unsigned long long)rs - 1ull; }
                    }
                }

                // ---- ofdmflexframesync layer
                int emit = 0;                      // 1: header invalid, 2: payload complete
                if (fstate == FS_PAYLOAD) {
                    // demap; the symbols leave one per byte (packet.cu packs them into the encoded bytes)
                    // (a frame that did not fit the arena is walked through without storing anything: take = 0 below)
                    const unsigned int take_all = min(p.M_data, S->payload_mod_len - pstart);
                    const unsigned int take = (S->sym_abs == ~0ull) ? 0u : take_all;
                    uint8_t * dst = p.arena + sym_off + pstart;
                    const float alpha = p.qam_alpha[S->bps_payload];
#define B2W_DEMAP(EXPR)                                                                  \
                    B2W_FORPTS {                                                         \
                        const unsigned int r = B2W_RK;                                   \
                        if (r < take) { const cf x = v[vt][s]; dst[r] = (uint8_t)(EXPR); } \
                    }
                    if (S->ms_payload == 40) { B2W_DEMAP((x.x > 0 ? 0u : 1u) + (x.y > 0 ? 0u : 2u)) }
                    else if (S->ms_payload == 39) { B2W_DEMAP(x.x > 0 ? 0u : 1u) }
                    else if (S->bps_payload == 6) { B2W_DEMAP(demod_qam_t<3>(x, alpha)) }
                    else if (S->bps_payload == 4) { B2W_DEMAP(demod_qam_t<2>(x, alpha)) }
                    else if (S->bps_payload == 8) { B2W_DEMAP(demod_qam_t<4>(x, alpha)) }
                    else { B2W_DEMAP(demod_qam_t<1>(x, alpha)) }
#undef B2W_DEMAP
                    pstart += take_all;
                    if (pstart == S->payload_mod_len) emit = 2;
                } else {
                    // header: BPSK, 288 symbols; EVM is measured on them (framesyncstats_s.evm)
                    const unsigned int take = min(p.M_data, 288u - S->header_sym_idx);
                    float ev = 0.f;
                    uint32_t * hwords = (uint32_t *)S->header_bits;
                    B2W_FORPTS {
                        const unsigned int r = B2W_RK;
                        if (r < take) {
                            const cf x = v[vt][s];
                            const unsigned int b = x.x > 0 ? 0u : 1u;
                            const unsigned int gb = S->header_sym_idx + r;
                            if (b) atomicOr(&hwords[gb >> 5], 1u << (8u * ((gb >> 3) & 3u) + 7u - (gb & 7u)));
                            const float dr = x.x - (b ? -1.0f : 1.0f);
                            ev += dr * dr + x.y * x.y;
                        }
                    }
                    ev = warp_sum(ev);
                    const float evm_new = S->evm_hat + ev;
                    const unsigned int hnew = S->header_sym_idx + take;
                    __syncwarp();
                    S->evm_hat = evm_new;
                    S->header_sym_idx = hnew;
                    if (hnew == 288u) {
                        __syncwarp();
                        // unscramble, de-interleave (n = 36, depth 4), Golay(24,12), CRC-32, parse
                        const uint8_t mask[4] = {0xb4, 0x6a, 0x8b, 0x45};
                        for (unsigned int i = lane; i < 36; i += 32) hb[i] = S->header_bits[i] ^ mask[i & 3];
                        __syncwarp();
                        const uint8_t ilmask[4] = {0xff, 0x0f, 0x55, 0x33};
#pragma unroll
                        for (int vq = 3; vq >= 0; vq--) {
                            if (lane < 18) {
                                const unsigned int j = __ldg(&p.tb.hdr_walk[18 * vq + lane]);
                                const uint8_t mk = ilmask[vq];
                                const uint8_t a = hb[2 * lane], b = hb[2 * j + 1];
                                hb[2 * lane] = (uint8_t)((a & ~mk) | (b & mk));
                                hb[2 * j + 1] = (uint8_t)((a & mk) | (b & ~mk));
                            }
                            __syncwarp();
                        }
                        if (lane < 12) {
                            const unsigned int vv = ((unsigned int)hb[3 * lane] << 16) | ((unsigned int)hb[3 * lane + 1] << 8) | hb[3 * lane + 2];
                            gsym[lane] = golay2412_decode(vv);
                        }
                        __syncwarp();
                        if (lane == 0) {
                            uint8_t * hd = S->header_dec;
                            for (int gq = 0; gq < 6; gq++) {
                                const unsigned int s0 = gsym[2 * gq], s1 = gsym[2 * gq + 1];
                                hd[3 * gq] = (s0 >> 4) & 0xff;
                                hd[3 * gq + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
                                hd[3 * gq + 2] = s1 & 0xff;
                            }
                            const uint32_t key = ((uint32_t)hd[14] << 24) | ((uint32_t)hd[15] << 16) | ((uint32_t)hd[16] << 8) | hd[17];
                            int valid = crc32_nibble(hd, 14) == key;
                            S->evm_db = 10 * log10f(S->evm_hat / 288.0f);
                            if (valid && hd[8] != 105) valid = 0;          // protocol id
                            const unsigned int plen = ((unsigned int)hd[9] << 8) | hd[10];
                            const unsigned int hms = hd[11], check = (hd[12] >> 5) & 7, fec0 = hd[12] & 0x1f, fec1 = hd[13] & 0x1f;
                            const unsigned int hbps = dev_mod_bps(hms);
                            if (valid && (hbps == 0 || (check != 6 && check != 1) || !dev_fec_ok(fec0) || !dev_fec_ok(fec1))) valid = 0;
                            unsigned int henc = 0, hmod = 0;
                            if (valid) {
                                henc = dev_fec_enc_len(fec1, dev_fec_enc_len(fec0, plen + (check == 6 ? 4 : 0)));
                                hmod = (8 * henc + hbps - 1) / hbps;
                            }
                            if (valid) {
                                S->payload_len = plen; S->check = check; S->fec0 = fec0; S->fec1 = fec1;
                                S->payload_enc_len = henc;
                            }
                            hres[0] = (uint32_t)valid; hres[1] = hms; hres[2] = hbps; hres[3] = hmod;
                        }
                        __syncwarp();
                        if (hres[0]) {
                            S->ms_payload = hres[1]; S->bps_payload = hres[2]; S->payload_mod_len = hres[3];
                            fstate = FS_PAYLOAD;
                            // room for the payload symbols in the arena ring (one byte per symbol, never across the ring's end)
                            {
                                const unsigned long long len = ((unsigned long long)S->payload_mod_len + 15ull) & ~15ull;
                                unsigned long long a = 0, ph = 0;
                                if (lane == 0 && len) {
                                    if (len > p.arena_cap / 2) { a = ~0ull; atomicOr(&p.counters[1], 8u); }
                                    else {
                                        do {
                                            a = atomicAdd((unsigned long long *)(p.counters + 2), len);
                                            ph = a % p.arena_cap;
                                        } while (ph + len > p.arena_cap);
                                    }
                                }
                                a = shfl0(a);
                                S->sym_abs = a; sym_off = shfl0(ph);
                            }
                            // (a frame without payload symbols -- liquid would wait for ever -- completes with the next OFDM
                            // symbol, as in the serial-chain kernels: take = 0 there)
                        } else emit = 1;
                        __syncwarp();
                    }
                }

                if (emit) {
                    // ---- append a frame record; the payload symbols are in the arena already
                    const unsigned int m2 = (emit == 2) ? S->payload_mod_len : 0u;      // symbols, one byte each
                    const bool stored = m2 && S->sym_abs != ~0ull;
                    if (lane == 0) {
                        unsigned int slot = 0, ok = 1;
                        unsigned long long offd = 0;
                        const unsigned int plen = (emit == 2 && (stored || !m2)) ? S->payload_len : 0u;
                        if (stored) {
                            // decoded payload (+ CRC) of this launch's frames: contiguous, in completion order per worker
                            offd = atomicAdd((unsigned long long *)(p.counters + 6), (unsigned long long)((plen + 4u + 15u) & ~15u));
                            if (offd + ((plen + 4u + 15u) & ~15u) > p.decoded_cap) ok = 0;
                            // the ring must not have come round to this frame's symbols
                            const unsigned long long now = *(volatile unsigned long long *)(p.counters + 2);
                            if (now - S->sym_abs > p.arena_cap - (((unsigned long long)m2 + 15ull) & ~15ull)) ok = 0;
                        }
                        if (direct) {
                            slot = atomicAdd(&p.counters[0], 1u);
                            if (slot >= p.recs_cap) ok = 0;
                        } else {
                            slot = S->nrec;
                            if (slot >= ctl->pcap) ok = 0;
                        }
                        if (!ok) atomicOr(&p.counters[1], 1u);
                        else {
                            FrameRec r;
                            r.channel = p.chan_base + ch;
                            r.header_valid = (emit == 2);
                            r.payload_valid = 0;
                            r.payload_len = plen;
                            for (int i = 0; i < 8; i++) r.header[i] = S->header_dec[i];
                            r.evm = S->evm_db;
                            r.rssi = -10.0f * log10f(S->g0);
                            r.cfo = nco_freq_dev(dth);
                            r.mod_scheme = (emit == 2) ? S->ms_payload : 0u;
                            r.mod_bps = (emit == 2) ? S->bps_payload : 0u;
                            r.check = (emit == 2) ? S->check : 0u;
                            r.fec0 = (emit == 2) ? S->fec0 : 0u;
                            r.fec1 = (emit == 2) ? S->fec1 : 0u;
                            r.detect_index = S->detect_index;
                            r.complete_index = p.sample_base + (unsigned long long)rs - 1ull;
                            r.payload_offset = offd;
                            FrameAux a;
                            a.enc_len = stored ? S->payload_enc_len : 0u;
                            a.sym_bps = stored ? S->bps_payload : ((emit == 2 && m2) ? 0xffffffffu : 0u);    // 0xffffffff: payload not stored (too large)
                            a.sym_off = sym_off;
                            if (direct) { p.recs[slot] = r; p.aux[slot] = a; }
                            else { p.wrecs[ctl->pbase + slot] = r; p.waux[ctl->pbase + slot] = a; }
                        }
                        if (!direct && ok) S->nrec = slot + 1u;
                    }
                    // ofdmflexframesync_reset; the symbol timer survives it, as in liquid
                    __syncwarp();
                    if (lane < 9) ((uint32_t *)S->header_bits)[lane] = 0u;
                    __syncwarp();
                    fstate = FS_HEADER; S->header_sym_idx = 0; pstart = 0; S->evm_hat = 0.f;
                    S->q_theta = th; S->q_dtheta = dth_old;     // phase sample `sidx` would have had under the step this symbol was mixed with
                    rms = seg_start; rme = rs;
                    th = 0; dth = 0; pilot_pos = 0; num_symbols = 0;
                    S->s_hat0_re = 0.f; S->s_hat0_im = 0.f; phi_prime = 0.f; p1_prime = 0.f;
                    state = ST_SEEK;
                    timer = (int)W;
                    if (lane == 0) { S->b_prev = S->b_last; S->b_last = p.sample_base + (unsigned long long)rs + 1ull; S->nb = min(S->nb + 1u, 2u); }
                    __syncwarp();
                }
            }

            // ---- canonical state on a later worker's start?  (SEEK with timer 0 is only ever left by a seek event
            //      that detected nothing: NCO, pilot generator, header / payload progress are all reset there)
            while (rcand < rs) { cand++; rcand = rstart(cand); }
            if (state == ST_SEEK && timer == 0 && rcand == rs) { matched = 1; break; }
        }

        // ================================================================ end of a stretch
        // (the stitcher's chain bookkeeping lives in ctl: lane 0 writes, everybody reads after a __syncwarp)
        if (!stitching) {
            save_state(slot_of(ctl->cur), matched);
            unsigned int last = 0;
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                last = (atomicAdd(&C->done, 1u) == ctl->n_act - 1u);
                __threadfence();
            }
            last = shfl0(last);
            if (!last) return;
            stitching = true;
            if (lane == 0) { ctl->sc = 0; ctl->chain_b_last = __ldcg(&C->b_last); ctl->chain_b_prev = __ldcg(&C->b_prev); }
            __syncwarp();
        } else {
            // back from a serial stretch that began in worker cur's final state
            if (lane == 0) {
                if (S->nb >= 2) { ctl->chain_b_last = S->b_last; ctl->chain_b_prev = S->b_prev; }
                else if (S->nb == 1) { ctl->chain_b_prev = ctl->chain_b_last; ctl->chain_b_last = S->b_last; }
                ctl->sc = matched ? cand : ctl->n_act;      // matched: met worker cand's start in the canonical state, its results stand
            }
            __syncwarp();
            if (!matched) save_state(slot_of(ctl->cur), 0);
            __syncwarp();
        }
        // ---- walk the workers in order
        unsigned int final_w = ctl->cur;
        bool again = false;
        for (;;) {
            const unsigned int sc = ctl->sc, n_act = ctl->n_act;
            if (sc >= n_act) break;
            const WSync * Q = p.wst + slot_of(sc);
            const unsigned int q_nb = __ldcg(&Q->nb), q_matched = __ldcg(&Q->matched), q_nrec = __ldcg(&Q->nrec);
            __syncwarp();
            if (lane == 0) {
                if (q_nb >= 2) { ctl->chain_b_last = __ldcg(&Q->b_last); ctl->chain_b_prev = __ldcg(&Q->b_prev); }
                else if (q_nb == 1) { ctl->chain_b_prev = ctl->chain_b_last; ctl->chain_b_last = __ldcg(&Q->b_last); }
            }
            if (sc != 0 && q_nrec) {
                // commit the private record list
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&p.counters[0], q_nrec);
                base = shfl0(base);
                if (base + q_nrec > p.recs_cap) { if (lane == 0) atomicOr(&p.counters[1], 1u); }
                else {
                    const unsigned int pb = priv_base(sc);
                    const uint32_t * src = (const uint32_t *)(p.wrecs + pb);
                    uint32_t * dst = (uint32_t *)(p.recs + base);
                    for (unsigned int i = lane; i < q_nrec * (sizeof(FrameRec) / 4); i += 32) dst[i] = __ldcg(src + i);
                    const uint32_t * asrc = (const uint32_t *)(p.waux + pb);
                    uint32_t * adst = (uint32_t *)(p.aux + base);
                    for (unsigned int i = lane; i < q_nrec * (sizeof(FrameAux) / 4); i += 32) adst[i] = __ldcg(asrc + i);
                }
            }
            final_w = sc;
            if (sc == n_act - 1) break;
            if (q_matched) {
                __syncwarp();
                if (lane == 0) ctl->sc = sc + 1;
                __syncwarp();
                continue;
            }
            // worker sc did not arrive in the canonical state: carry on serially from its final (true) state
            load_state(slot_of(sc));
            if (lane == 0) ctl->cur = sc;
            __syncwarp();
            direct = true; rlimit = (int)p.nsamples;
            cand = sc + 2; rcand = rstart(cand);
            again = true;
            break;
        }
        if (again) continue;

        // ================================================================ the stream's launch is complete
        if (ctl->sc >= ctl->n_act) final_w = ctl->cur;
        {
            const WSync * F = p.wst + slot_of(final_w);
            __syncwarp();
            const int f_state = __ldcg(&F->state), f_timer = __ldcg(&F->timer);
            const unsigned long long E1 = p.sample_base + p.nsamples;
            const unsigned long long chain_b_last = ctl->chain_b_last, chain_b_prev = ctl->chain_b_prev;
            unsigned long long nx = 0;
            unsigned int per = 0;
            const bool have = chain_b_last && chain_b_prev && chain_b_last > chain_b_prev && (chain_b_last - chain_b_prev) < (1ull << 31);
            if (have && (E1 - min(E1, chain_b_last)) < 2ull * (chain_b_last - chain_b_prev)) {
                per = (unsigned int)(chain_b_last - chain_b_prev); nx = chain_b_last;          // frames keep arriving
            } else if (f_state == ST_SEEK) {
                per = M; nx = E1 + (f_timer < (int)M ? (unsigned long long)((int)M - f_timer) : 1ull);   // idle seek grid
            } else if (have) {
                per = (unsigned int)(chain_b_last - chain_b_prev); nx = chain_b_last;
            }
            // the launch's last M + cp raw samples, for the windows of the next launch that reach back
            cf * ringw = p.ring + (size_t)ch * W;
            if (p.nsamples >= W) {
                for (unsigned int j = lane; j < W; j += 32) ringw[j] = __ldg(in + (p.nsamples - W + j));
            } else {
                cf * tmp = fbuf;                       // fbuf and RG are contiguous: 2 M >= W samples
                const unsigned int n = p.nsamples;
                for (unsigned int j = lane; j < W; j += 32) tmp[j] = (j + n < W) ? __ldcg(ringw + j + n) : __ldg(in + (j + n - W));
                __syncwarp();
                for (unsigned int j = lane; j < W; j += 32) ringw[j] = tmp[j];
            }
            if (lane == 0) {
                C->pred_next[par ^ 1u] = nx; C->pred_period[par ^ 1u] = per;
                C->head[par ^ 1u] = (ctl->head + final_w) % p.wslots;
                C->b_last = chain_b_last; C->b_prev = chain_b_prev;
                C->done = 0;
            }
        }
        return;
    }
}

// ofdmflexframesync_reset on every stream (lib/multichannelrx.cc:140): the state of the stream's head slot
__global__ void syncw_reset_kernel(WSync * wst, WChan * wch, unsigned int streams, unsigned int wslots, unsigned int M, unsigned long long sample_base)
{
    const unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= streams) return;
    WChan * C = wch + c;
    const unsigned int head = C->head[0];
    C->head[1] = head;
    WSync * S = wst + (size_t)c * wslots + head;
    S->fstate = FS_HEADER;
    S->header_sym_idx = 0; S->payload_sym_idx = 0;
    S->evm_hat = 0.f;
    S->nco_theta = 0; S->nco_dtheta = 0;
    S->pilot_pos = 0;
    S->timer = 0;
    S->num_symbols = 0;
    S->s_hat0_re = 0.f; S->s_hat0_im = 0.f;
    S->phi_prime = 0.f; S->p1_prime = 0.f;
    S->state = ST_SEEK;
    S->mix_start = 0; S->mix_end = 0;
    S->sample_index = sample_base;
    for (int k = 0; k < 36; k++) S->header_bits[k] = 0;
    C->pred_next[0] = C->pred_next[1] = sample_base + M; C->pred_period[0] = C->pred_period[1] = M;
    C->b_last = 0; C->b_prev = 0; C->done = 0;
}
cudaError_t syncw_reset_launch(WSync * wst, WChan * wch, unsigned int streams, unsigned int wslots, unsigned int M,
                               unsigned long long sample_base, cudaStream_t st)
{
    syncw_reset_kernel<<<(streams + 127) / 128, 128, 0, st>>>(wst, wch, streams, wslots, M, sample_base);
    return cudaGetLastError();
}

bool syncw_supported(unsigned int M) { return M == 256 || M == 512; }

template <unsigned int M, unsigned int WPC, unsigned int MINB>
static cudaError_t syncw_launch_t(const SyncParams & p, cudaStream_t st)
{
    static size_t configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = sw_layout(M, p.M_pilot, WPC).total;
    size_t & conf = configured[(unsigned int)dev & 63u];
    if (smem > conf) {
        cudaError_t e = cudaFuncSetAttribute(syncw_kernel<M, WPC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        conf = smem;
    }
    const unsigned int warps = p.streams * p.workers;
    syncw_kernel<M, WPC, MINB><<<(warps + WPC - 1) / WPC, WPC * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t syncw_launch(const SyncParams & p, cudaStream_t st)
{
    if (p.nsamples == 0 || p.streams == 0) return cudaSuccess;
    // (experiment knob: warps per CTA / CTAs per SM, i.e. the register budget)
    static const int variant = getenv("B2_SYNCW_VARIANT") ? atoi(getenv("B2_SYNCW_VARIANT")) : 0;
    switch (p.M) {
    case 256: return syncw_launch_t<256, 4, 4>(p, st);
    case 512:
        switch (variant) {
        case 2:  return syncw_launch_t<512, 2, 8>(p, st);
        case 3:  return syncw_launch_t<512, 2, 10>(p, st);
        case 4:  return syncw_launch_t<512, 1, 16>(p, st);
        case 5:  return syncw_launch_t<512, 1, 20>(p, st);
        case 6:  return syncw_launch_t<512, 4, 4>(p, st);
        default: return syncw_launch_t<512, 4, 5>(p, st);      // 96 registers, 20 workers per SM
        }
    default:  return cudaErrorInvalidValue;
    }
}

} // namespace b2
