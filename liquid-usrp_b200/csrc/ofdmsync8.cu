// ofdmsync8.cu -- latency-optimised OFDM frame synchroniser for M >= 256 subcarriers.
//
// Same job and same persistent state as ofdmsync.cu (liquid's ofdmframesync +
// ofdmflexframesync state machines, reached by the reference through
//     ofdmflexframesync_execute(framesync[i], &X[i], 1)        lib/multichannelrx.cc:194
//     ofdmflexframesync_execute(fs, &sample, 1)                lib/ofdmtxrx.cc:625 )
// but organised around the fact that one stream is a SERIAL chain of events (the NCO is
// trimmed from each OFDM symbol's pilot phase before the next symbol is mixed), so what counts
// is the latency of one event, not the throughput of one CTA:
//   * a stream is carried by M/8 threads (64 for M = 512); every thread keeps 8 subcarriers in
//     registers from the first FFT pass to the demapper (Stockham radix-8, natural-order output,
//     fft8.cuh), together with their equaliser taps R[k] and subcarrier roles;
//   * the NCO mix-down happens on the way from the staging ring into the first FFT pass;
//   * a steady-state payload symbol costs 4 CTA barriers of 2 warps: 2 FFT exchanges, pilots -> warp 0
//     (polynomial atan2 / unwrap / line fit / NCO trim, all in registers), fit -> all.  Its tail is
//     PIPELINED: the next symbol's samples are mixed and taken through the first FFT pass in the same
//     stretch of code that derotates and demaps this symbol, so there is no loop-top barrier and no
//     state reload between payload symbols, and two dependency chains interleave;
//   * the demapped symbols leave as one byte each (bit packing is throughput work, done by packet.cu
//     off the chain); header bits are packed with shared-memory atomicOr;
//   * the S1 equaliser-gain polynomial fit is a constant 5 x Na matrix (design.h) applied to the
//     measured |G| / arg G instead of a per-frame normal-equation solve, and unwraps only if needed;
//   * samples are prefetched two events ahead with 16-byte cp.async into a ring addressed by stream
//     position, by the LAST warp while warp 0 fits the pilots; cp.async.wait_group 1 leaves the
//     newest group in flight;
//   * complex arithmetic is packed FP32x2 (FADD2 / FMUL2 / FFMA2, dsp.cuh); FFT twiddles, equaliser
//     taps, training signs, S1 tables and header de-interleaver walks live in registers.
// A chain is 2 warps, <= 255 registers and ~39 KB of shared memory: 4 chains per SM, 256 chains on the
// 72-SM partition capi.cu / smpart.cu give the synchronisers, the channelizer of the next chunk
// running on the other SMs.
#include "kernels.h"
#include "fec.cuh"
#include "syncdev.cuh"
#include "fft8.cuh"

namespace b2 {

struct S8Layout {
    unsigned int SZ;
    size_t off_st, off_red, off_dsum, off_stg, off_hist, off_fa, off_fb, off_G0, off_yc, off_px, off_sym, off_pseq, total;
};
// ~39 KB for M = 512 (16 KB of it the staging ring).
// Gs (training-symbol gains of the current event) aliases FFT buffer B and yph (phases handed to
// warp 0) aliases FFT buffer A: both are only touched between the last FFT pass of an event and
// the top barrier of the next one (which of the two is safe for Gs depends on the pass count).
__host__ __device__ static inline S8Layout s8_layout(unsigned int M, unsigned int cp, unsigned int Na, unsigned int Mp)
{
    S8Layout L;
    const unsigned int W = M + cp;
    L.SZ = 256;
    while (L.SZ < W + M / 2 + 64) L.SZ <<= 1;
    // two events deep where it is cheap (<= 32 KB): the newest prefetch group may then stay in flight
#ifndef B2_S8_SHALLOW_RING
    if (L.SZ < 2 * W + M / 2 + 64 && L.SZ * 2 * sizeof(cf) <= 32768) L.SZ <<= 1;
#endif
    size_t o = 0;
    L.off_st = o;   o += (sizeof(SyncState) + 15) & ~(size_t)15;
    L.off_red = o;  o += 160 * sizeof(float);
    L.off_dsum = o; o += (16 * 10 + 32) * sizeof(double);
    L.off_stg = o;  o += (size_t)L.SZ * sizeof(cf);
    L.off_hist = o; o += (size_t)W * sizeof(cf);
    L.off_fa = o;   o += (size_t)f8_buf_elems(M) * sizeof(cf);          // >= 2*(Na+4) floats
    L.off_fb = o;   o += (size_t)f8_buf_elems(M) * sizeof(cf);          // >= M cf
    L.off_G0 = o;   o += (size_t)M * sizeof(cf);
    L.off_yc = o;   o += (size_t)(Mp + 4) * sizeof(cf);
    L.off_px = o;   o += (size_t)(Mp + 4) * sizeof(float);
    L.off_sym = o;  o += 304;                                           // 288 header bits
    L.off_pseq = o; o += 256;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

__device__ __forceinline__ void cp_async16(void * dst_smem, const void * src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory"); }

// optional phase profile (build with -DB2_SYNC_PROF): cycles spent by CTA 0 between marks
#ifdef B2_SYNC_PROF
__device__ unsigned long long g_sync8_prof[16];
#define PH(k) do { if (blockIdx.x == 0 && t == 0) { long long _t = clock64(); atomicAdd(&g_sync8_prof[k], (unsigned long long)(_t - t_last)); t_last = _t; } } while (0)
extern "C" int b2_debug_sync8_prof(unsigned long long * out, int reset)
{
    if (out) cudaMemcpyFromSymbol(out, g_sync8_prof, sizeof(g_sync8_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_sync8_prof, z, sizeof(z)); }
    return 0;
}
#else
#define PH(k) do { } while (0)
#endif

template <unsigned int M>
#ifndef B2_S8_THREADS_PER_SM
#define B2_S8_THREADS_PER_SM 256        // experiment knob: 448 = seven 64-thread chains per SM (<= 144 registers)
#endif
__global__ void __launch_bounds__(M / 8, (M <= 1024 ? B2_S8_THREADS_PER_SM : 512) / (M / 8)) sync8_kernel(const SyncParams p)   // <= 255 registers at M = 512: 4 streams per SM
{
    constexpr unsigned int T = M / 8, NW = T / 32, M2 = M / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const unsigned int cp = p.cp, W = M + cp;
    // frame-pipelined worker pairs (p.workers == 2): CTA 2c / 2c + 1 are the two workers of stream c; see the
    // protocol note above sync8_gate below.  vs indexes the per-worker state, sidx the stream (channel)
    const bool duo = (p.workers == 2);
    const unsigned int vs = blockIdx.x;
    const unsigned int sidx = duo ? (vs >> 1) : vs;
    const unsigned int wk = duo ? (vs & 1u) : 0u;
    const unsigned int Na = p.M_pilot + p.M_data, Mp = p.M_pilot;
    const S8Layout L = s8_layout(M, cp, Na, Mp);
    SyncState * S = (SyncState *)(smem + L.off_st);
    float * red = (float *)(smem + L.off_red);
    double * dsum = (double *)(smem + L.off_dsum);
    cf * stg = (cf *)(smem + L.off_stg);
    cf * hist = (cf *)(smem + L.off_hist);
    cf * fa = (cf *)(smem + L.off_fa);
    cf * fb = (cf *)(smem + L.off_fb);
    cf * G0 = (cf *)(smem + L.off_G0);
    // aliases (see s8_layout): Gs must be the buffer the LAST exchange of the FFT did not use (a thread
    // leaving the last pass may not overwrite what a slower thread is still loading)
    constexpr bool last_load_from_b = (M == 256 || M == 512);          // 3 passes: A, B; 4 passes: A, B, A
    cf * Gs = last_load_from_b ? fa : fb;
    float * yph = (float *)(last_load_from_b ? fb : fa);               // [0..Na) y / y_arg, [Na..2Na) y_abs
    cf * yc = (cf *)(smem + L.off_yc);                  // pilots of the current symbol, sign removed
    float * pilot_x = (float *)(smem + L.off_px);
    uint8_t * sym = (uint8_t *)(smem + L.off_sym);
    uint8_t * pilot_seq = (uint8_t *)(smem + L.off_pseq);
    const unsigned int SZM = L.SZ - 1, PF = L.SZ - 2;
    const unsigned int NEED_MAX = W + M2;

    const cf * in = p.in + (size_t)sidx * p.in_stride;
    uint8_t * penc = p.penc + (size_t)vs * p.penc_cap;
    const bool al16 = (((size_t)in) & 15) == 0;

    // ---- sample prefetch: ring slot = stream position & SZM.  Issued by the LAST warp only, at the
    //      points of an event where that warp would otherwise wait for warp 0 (pilot fit, metric maths);
    //      every thread tracks the frontiers, the issuing warp owns the cp.async groups.
    unsigned int fetched = 0, done_frontier = 0;
    const bool pf_warp = (wid == NW - 1);
    const unsigned int tp = lane;
    auto prefetch = [&](unsigned int upto) {
        unsigned int hi = upto;
        if (al16 && hi < p.nsamples) hi &= ~1u;
        done_frontier = fetched;
        if (hi > fetched) {
            if (pf_warp) {
                if (al16) {
                    const unsigned int even_hi = hi & ~1u;
                    for (unsigned int i = fetched + 2 * tp; i < even_hi; i += 64) cp_async16(&stg[i & SZM], in + i);
                    if ((hi & 1u) && tp == 0) cp_async8(&stg[(hi - 1) & SZM], in + hi - 1);
                } else {
                    for (unsigned int i = fetched + tp; i < hi; i += 32) cp_async8(&stg[i & SZM], in + i);
                }
            }
            fetched = hi;
        }
        if (pf_warp) cp_async_commit();
    };
    if (!duo) prefetch(min(PF, p.nsamples));     // a worker of a pair first has to learn where it stands

    // ---- persistent state and tables
    cf Rr[8];
    unsigned int rk[8];
    float ref_s0[8], ref_s1[8];              // training symbols (+-1 / 0) of own subcarriers
    {
        const uint32_t * src = (const uint32_t *)(p.st + vs);
        uint32_t * dst = (uint32_t *)S;
        for (unsigned int i = t; i < sizeof(SyncState) / 4; i += T) dst[i] = src[i];
        const cf * gr = p.ring + (size_t)vs * W;
        for (unsigned int i = t; i < W; i += T) hist[i] = gr[i];
        const cf * g0 = p.G0 + (size_t)vs * M;
        const cf * gR = p.R + (size_t)vs * M;
#pragma unroll
        for (unsigned int s = 0; s < 8; s++) {
            const unsigned int i = t + s * T;
            G0[i] = g0[i];
            Rr[s] = gR[i];
            rk[s] = p.tb.sc_rank[i];
            ref_s0[s] = p.tb.S0[i];
            ref_s1[s] = p.tb.S1[i];
        }
        for (unsigned int i = t; i < Mp; i += T) pilot_x[i] = p.tb.pilot_x[i];
        for (unsigned int i = t; i < 255; i += T) pilot_seq[i] = p.tb.pilot_seq[i];
    }
    cf twr[f8_tw_count(M, 8) + 1];           // this thread's twiddles of the passes after the first
    f8_tw_init<M, 8>(twr, t, p.fft.tw);
    const cf * tw = nullptr;                 // no table at run time
    unsigned int ar[8];                      // rank of own subcarriers among the active ones (fft-shifted order)
    cf Bq[8];                                // B[i] = e^{j 2 pi backoff i / M}: timing back-off of the S1 gain estimate
#pragma unroll
    for (unsigned int s = 0; s < 8; s++) {
        const unsigned int i = t + s * T;
        ar[s] = p.tb.act_rank[i];
        Bq[s] = p.tb.B[i];
    }
    // header de-interleaver walks (n = 36): lane l < 18 swaps bytes 2l <-> 2 walk[v][l] + 1 in pass v
    unsigned int hwalk[4] = {0, 0, 0, 0};
    if (lane < 18) {
#pragma unroll
        for (int vq = 0; vq < 4; vq++) hwalk[vq] = p.tb.hdr_walk[18 * vq + lane];
    }
    const float px0 = lane < Mp ? p.tb.pilot_x[lane] : 0.f, px1 = lane + 32 < Mp ? p.tb.pilot_x[lane + 32] : 0.f;
    float fxs[8];                            // signed subcarrier index of own subcarriers
    unsigned int pilot_mask = 0;
#pragma unroll
    for (unsigned int s = 0; s < 8; s++) {
        const unsigned int i = t + s * T;
        fxs[s] = (i > M2) ? (float)i - (float)M : (float)i;
        if ((rk[s] & 0xC000u) == 0x4000u) pilot_mask |= 1u << s;
    }
    unsigned int pos = 0;
    // ---- worker pair: role and starting point of this worker in this launch
    volatile SyncCtl * ctl = duo ? (volatile SyncCtl *)(p.ctl + sidx) : nullptr;
    unsigned int role = SW_OWNER;
    if (duo) {
        __syncthreads();                     // state copy visible
        role = S->role;
        if (role != SW_WAIT) {
            const unsigned long long rel = S->sample_index - p.sample_base;
            pos = rel >= (unsigned long long)p.nsamples ? p.nsamples : (unsigned int)rel;
        }
        fetched = pos & ~1u;
        done_frontier = fetched;
        if (role != SW_WAIT) prefetch(min(pos + PF, p.nsamples));     // a waiting worker learns its position with the hand-off
    }
#ifdef B2_SYNC_PROF
    long long t_last = clock64();
#endif

    // block-wide sum of 4 floats per thread, result in every thread (one barrier)
    auto block_sum4 = [&](float & a, float & b, float & c, float & d) {
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
        if (NW > 1) {
            if (lane == 0) { red[4 * wid] = a; red[4 * wid + 1] = b; red[4 * wid + 2] = c; red[4 * wid + 3] = d; }
            __syncthreads();
            a = 0.f; b = 0.f; c = 0.f; d = 0.f;
#pragma unroll
            for (unsigned int w = 0; w < NW; w++) { a += red[4 * w]; b += red[4 * w + 1]; c += red[4 * w + 2]; d += red[4 * w + 3]; }
        }
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
        for (int i = 0; i < 9; i++) ((uint32_t *)S->header_bits)[i] = 0u;
        S->fstate = FS_HEADER;
        S->header_sym_idx = 0;
        S->payload_sym_idx = 0;
        S->evm_hat = 0.f;
        phy_reset();
    };
    auto bsync = [] { __syncthreads(); };

    // ---- worker pair protocol (thread 0 talks to global memory; decisions reach the CTA through red[120..])
    // One worker OWNS the stream position; when it has decoded a valid header it knows where its frame ends and
    // hands the search for the NEXT frame to its partner (ctl->start, ctl->hs = seq|SENT), which starts from a
    // fresh ofdmflexframesync_reset one sample past the frame end -- the state the owner itself would be in
    // there, PROVIDED the seek event liquid runs on that first sample (timer = M + cp survives the reset) finds
    // nothing.  The owner runs that event on its own window when its frame is over and settles the hand-off:
    // ACCEPTED (it now waits for a hand-off itself) or ABORTED (it detected something: it simply carries on and
    // the partner throws its speculative work away).  Until then the partner is SPECULATIVE: it computes, but
    // neither emits records nor hands off; where it cannot wait (an invalid header wants a record, the launch
    // runs out of samples) it waits at a resumable point or returns the hand-off (hs -> NONE) and the owner
    // carries on serially.  Results are those of the serial chain in every case.
    auto publish = [&](unsigned long long start) {         // thread 0
        const unsigned int seq = (ctl->hs >> 2) + 1u;
        ctl->start = start;
        __threadfence();
        ctl->hs = (seq << 2) | HS_SENT;
        S->sent_seq = seq;
    };
    // loop-top gate; true: this worker leaves the launch (nothing more it can do with these samples)
    auto gate = [&]() -> bool {
        for (;;) {
            if (role == SW_OWNER) {
                const bool settle = S->verify && !(S->state == ST_SEEK && S->timer == (int)(M + cp));
                if (settle) {
                    // the seek event after my frame has run: settle the hand-off I published
                    __syncthreads();         // everybody has read the flags thread 0 is about to change
                    if (t == 0) {
                        const unsigned int w0 = (S->sent_seq << 2) | HS_SENT;
                        unsigned int nr = SW_OWNER;
                        if (S->state == ST_SEEK) {
                            if (atomicCAS((unsigned int *)&ctl->hs, w0, (S->sent_seq << 2) | HS_ACCEPTED) == w0) nr = SW_WAIT;
                        } else {
                            atomicCAS((unsigned int *)&ctl->hs, w0, (S->sent_seq << 2) | HS_ABORTED);
                        }
                        S->verify = 0; S->sent_seq = 0; S->role = nr;
                        red[120] = __uint_as_float(nr);
                    }
                    __syncthreads();
                    role = __float_as_uint(red[120]);
                    __syncthreads();
                    if (role == SW_OWNER) return false;
                    continue;
                }
                return false;
            }
            if (role == SW_SPEC) {
                // may the coming event have side effects?  (last payload symbol: record + arena reservation)
                const bool blocking = (S->state == ST_RX) && (S->fstate == FS_PAYLOAD) &&
                                      (S->payload_sym_idx + min(p.M_data, S->payload_mod_len - S->payload_sym_idx) == S->payload_mod_len);
                if (t == 0) {
                    const unsigned int mine = S->my_seq << 2;
                    unsigned int res = 0;                   // 0 pending, 1 accepted, 2 aborted, 3 leave the launch
                    const long long t0 = clock64();
                    for (;;) {
                        const unsigned int hs = ctl->hs;
                        if (hs == (mine | HS_ACCEPTED)) { res = 1; break; }
                        if (hs != (mine | HS_SENT)) { res = 2; break; }
                        if (!blocking) break;
                        if (ctl->done[wk ^ 1u] == p.launch_id) {
                            const unsigned int h2 = ctl->hs;
                            if (h2 == (mine | HS_ACCEPTED)) res = 1; else if (h2 != (mine | HS_SENT)) res = 2; else res = 3;
                            break;
                        }
                        if (clock64() - t0 > (1ll << 32)) { atomicOr(&p.counters[1], 2u); res = 3; break; }
                    }
                    if (res == 1) {
                        S->role = SW_OWNER;
                        if (S->pub_pending) { if (S->fstate == FS_PAYLOAD) publish(S->pub_start); S->pub_pending = 0; }
                    } else if (res == 2) { S->role = SW_WAIT; S->pub_pending = 0; }
                    red[120] = __uint_as_float(res);
                }
                __syncthreads();
                const unsigned int res = __float_as_uint(red[120]);
                __syncthreads();
                if (res == 3) return true;
                if (res == 1) role = SW_OWNER;
                if (res == 2) { role = SW_WAIT; continue; }
                return false;
            }
            // SW_WAIT: a hand-off, or the partner leaving the launch without one
            if (t == 0) {
                unsigned int res = 0;
                const long long t0 = clock64();
                for (;;) {
                    unsigned int hs = ctl->hs;
                    if ((hs & 3u) == HS_SENT) { res = hs >> 2; break; }
                    if (ctl->done[wk ^ 1u] == p.launch_id) {
                        hs = ctl->hs;
                        if ((hs & 3u) == HS_SENT) res = hs >> 2;
                        break;
                    }
                    if (clock64() - t0 > (1ll << 32)) { atomicOr(&p.counters[1], 2u); break; }
                }
                if (res) {
                    __threadfence();
                    const unsigned long long start = ctl->start;
                    flex_reset();
                    S->ring_head = 0;
                    S->sample_index = start;
                    S->detect_index = 0;
                    S->role = SW_SPEC; S->my_seq = res; S->sent_seq = 0; S->verify = 0; S->pub_pending = 0;
                    const unsigned long long rel = start - p.sample_base;
                    red[121] = __uint_as_float(rel >= (unsigned long long)p.nsamples ? p.nsamples : (unsigned int)rel);
                }
                red[120] = __uint_as_float(res);
            }
            __syncthreads();
            const unsigned int res = __float_as_uint(red[120]);
            const unsigned int npos = __float_as_uint(red[121]);
            __syncthreads();
            if (!res) return true;
            role = SW_SPEC;
            pos = npos;
            cp_async_wait_group<0>();        // restart the sample prefetch at the new position
            __syncthreads();                 // (ring slots change hands between the lanes of the issuing warp)
            fetched = pos & ~1u;
            done_frontier = fetched;
            prefetch(min(pos + PF, p.nsamples));
            cp_async_wait_group<0>();
            __syncthreads();
            return false;
        }
    };

    // per-event registers; `pre` = this event's samples are already consumed, mixed and through the
    // first FFT pass (done by the previous payload event, see the pipelined tail of the RX path)
    bool pre = false;
    int state = 0, timer = 0, fstate = 0;
    unsigned int head = 0, ppos = 0, hstart = 0, pstart = 0, bps = 0, ms = 0, mod_len = 0, adv = 0, head2 = 0, off = 0;
    uint32_t th = 0, dth = 0;
    float en = 0.f;
    float r_p1p = 0.f, r_phip = 0.f;         // pilot-fit memory (p1_prime, phi_prime) and symbol count of the frame,
    unsigned int r_nsym = 0;                 // register copies so that the fit does not wait on shared memory
    cf v[8];

    PH(15);                                   // (profile build) launch set-up
    while (true) {
      if (!pre) {
        PH(6);
        if (pos + NEED_MAX <= done_frontier) cp_async_wait_group<1>();
        else cp_async_wait_group<0>();
        __syncthreads();                     // staged samples + state of the previous event visible
        if (duo && gate()) break;
        PH(0);
        // ---- advance to the next event (or to the end of this launch's samples)
        state = S->state;
        timer = S->timer;
        head = S->ring_head;
        th = S->nco_theta; dth = S->nco_dtheta;
        ppos = S->pilot_pos;
        fstate = S->fstate;
        hstart = S->header_sym_idx; pstart = S->payload_sym_idx;
        bps = S->bps_payload; ms = S->ms_payload; mod_len = S->payload_mod_len;
        r_p1p = S->p1_prime; r_phip = S->phi_prime; r_nsym = S->num_symbols;
        unsigned int need;
        if (state == ST_SEEK) need = (timer < (int)M) ? (unsigned int)((int)M - timer) : 1u;
        else if (state == ST_S0A || state == ST_S0B) need = (timer < (int)M2) ? (unsigned int)((int)M2 - timer) : 1u;
        else need = (timer > 1) ? (unsigned int)timer : 1u;
        const unsigned int avail = p.nsamples - pos;
        adv = min(need, avail);
        const bool fire = (adv == need);
        off = (state == ST_RX) ? cp - p.backoff : cp;                     // FFT window offset in the sample window
        const bool mixing = (state != ST_SEEK) && ((th | dth) != 0u);     // e^{-j0} = 1 exactly
        head2 = head + adv;
        while (head2 >= W) head2 -= W;

        if (state == ST_RX && adv == W) {
            // steady state of a frame: the whole window is replaced, so it is rewritten from slot 0
            // (head2 = 0) and the FFT window starts `off` samples in
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const unsigned int j = off + t + s * T;
                const cf x = mix_down(stg[(pos + j) & SZM], nco_cexp_fast(th + j * dth));     // e^{-j0} = 1 exactly
                v[s] = x;
                hist[j] = x;
            }
            for (unsigned int jj = t; jj < cp; jj += T) {
                const unsigned int j = (jj < off) ? jj : jj + M;
                hist[j] = mix_down(stg[(pos + j) & SZM], nco_cexp_fast(th + j * dth));
            }
            head2 = 0;
        } else if (state == ST_SEEK && adv == M) {
            // idle seek: the M new samples ARE the FFT window (no NCO while seeking); push and gather at once
            unsigned int k = head + t;
            while (k >= W) k -= W;
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const cf x = stg[(pos + t + s * T) & SZM];
                v[s] = x;
                hist[k] = x;
                k += T;
                if (k >= W) k -= W;
            }
        } else {
            // only the last W of the new samples can survive in the window
            const unsigned int jlo = adv > W ? adv - W : 0u;
            unsigned int k = head + jlo + t;
            while (k >= W) k -= W;
            const unsigned int kstep = T % W;            // T < W for every supported shape
            if (mixing) {
                for (unsigned int j = jlo + t; j < adv; j += T) {
                    hist[k] = mix_down(stg[(pos + j) & SZM], nco_cexp_fast(th + j * dth));
                    k += kstep;
                    if (k >= W) k -= W;
                }
            } else {
                for (unsigned int j = jlo + t; j < adv; j += T) {
                    hist[k] = stg[(pos + j) & SZM];
                    k += kstep;
                    if (k >= W) k -= W;
                }
            }
            __syncthreads();
            if (fire) {
#pragma unroll
                for (unsigned int s = 0; s < 8; s++) {
                    unsigned int k = head2 + off + t + s * T;
                    if (k >= W) k -= W;
                    if (k >= W) k -= W;
                    v[s] = hist[k];
                }
            }
        }
        pos += adv;
        if (!fire) {                         // out of samples; resume in the next launch
            if (t == 0) {
                S->ring_head = head2;
                if (state != ST_SEEK) S->nco_theta = th + adv * dth;
                S->sample_index += adv;
                if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) S->timer = timer + (int)adv;
                else S->timer = timer - (int)adv;
            }
            break;
        }
        en = 0.f;
        if (state == ST_SEEK) {
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) en += v[s].x * v[s].x + v[s].y * v[s].y;
        }

        // ---- M-point forward FFT, 8 points per thread; the first exchange also orders the
        //      state reads above against the position update below
        f8_pass<M, 1, 8, -1>(v, t, tw);
        f8_store<M, 1, 8>(v, t, fa);
        __syncthreads();
        PH(1);
      }
      pre = false;
        // last payload symbol of a frame: reserve the record slot and the arena space now, so that the
        // round trip of the two atomics hides behind this symbol's FFT instead of sitting in the emit path
        const bool last_payload = (state == ST_RX) && (fstate == FS_PAYLOAD) && (pstart + min(p.M_data, mod_len - pstart) == mod_len);
        unsigned int pre_slot = 0;
        unsigned long long pre_off = 0;
        if (t == 0 && last_payload) {
            pre_slot = atomicAdd(&p.counters[0], 1u);
            pre_off = atomicAdd((unsigned long long *)(p.counters + 2), (unsigned long long)((mod_len + 15u) & ~15u));
        }
        if (t == 0) {
            S->ring_head = head2;
            if (state != ST_SEEK) S->nco_theta = th + adv * dth;
            S->sample_index += adv;
            if (state == ST_SEEK || state == ST_S0A || state == ST_S0B) S->timer = timer + (int)adv;
            else S->timer = timer - (int)adv;
        }
        f8_load<M>(v, t, fa);
        f8_run<M, 8, -1>(v, t, fb, fa, tw, bsync, twr);
        PH(2);

        if (state != ST_RX) {
            prefetch(min(pos + PF, p.nsamples));
            // ---- preamble events.  G[i] = X[i]*ref[i]*gain on the training subcarriers
            const bool long_seq = (state == ST_S1);
            const unsigned int step = long_seq ? 1u : 2u;
            const float gain = sqrtf((float)(long_seq ? p.M_S1 : p.M_S0)) / (float)M;
            cf g[8];
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const unsigned int i = t + s * T;
                const float r = long_seq ? ref_s1[s] : ref_s0[s];
                g[s] = make_float2(v[s].x * r * gain, v[s].y * r * gain);
                Gs[i] = g[s];
            }
            __syncthreads();
            float mr = 0.f, mi = 0.f, cr = 0.f, ci = 0.f;
            // s_hat = sum G[i+step] conj(G[i]) over the training subcarriers (ref = 0 elsewhere, so the
            // products of the odd subcarriers of S0 vanish by themselves)
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const unsigned int i = t + s * T;
                const cf g2 = Gs[(i + step) & (M - 1)];
                const cf tt = cmulc(g2, g[s]);
                mr += tt.x; mi += tt.y;
            }
            if (state == ST_S0A) {
#pragma unroll
                for (unsigned int s = 0; s < 8; s++) G0[t + s * T] = g[s];
            } else if (state == ST_S0B) {
#pragma unroll
                for (unsigned int s = 0; s < 8; s++) { const cf tt = cmulc(g[s], G0[t + s * T]); cr += tt.x; ci += tt.y; }
            }
            if (state == ST_SEEK) cr = en;
            block_sum4(mr, mi, cr, ci);
            if (state == ST_SEEK) {
                if (t == 0) {
                    float gg = (float)M / cr;
                    cf s_hat = make_float2(mr / (float)p.M_S0 * gg, mi / (float)p.M_S0 * gg);
                    S->g0 = gg;
                    S->timer = 0;
                    if (hypotf(s_hat.x, s_hat.y) > p.thresh) {
                        float tau_hat = atan2f(s_hat.y, s_hat.x) * (float)M2 / (2 * PI_F);
                        int dt = (int)roundf(tau_hat);
                        S->timer = (int)((M + (unsigned int)dt) % M2) + (int)M;
                        S->state = ST_S0A;
                        S->detect_index = S->sample_index - 1;
                    }
                }
            } else if (state == ST_S0A) {
                if (t == 0) {
                    S->timer = 0;
                    S->s_hat0_re = mr / (float)p.M_S0 * S->g0;
                    S->s_hat0_im = mi / (float)p.M_S0 * S->g0;
                    S->state = ST_S0B;
                }
            } else if (state == ST_S0B) {
                if (t == 0) {
                    float s1r = mr / (float)p.M_S0 * S->g0, s1i = mi / (float)p.M_S0 * S->g0;
                    float tau_hat = atan2f(S->s_hat0_im + s1i, S->s_hat0_re + s1r) * (float)M2 / (2 * PI_F);
                    S->timer = (int)(M + cp - p.backoff) - (int)roundf(tau_hat);
                    float nu_hat = 2.0f * atan2f(ci, cr) / (float)M;
                    S->nco_dtheta = nco_constrain_dev(nu_hat);
                    S->state = ST_S1;
                }
            } else {
                // ---- S1: accept / retry, and on accept the equaliser
                if (t == 0) {
                    S->num_symbols++;
                    cf s_hat = make_float2(mr / (float)p.M_S1 * S->g0, mi / (float)p.M_S1 * S->g0);
                    s_hat = cmul(s_hat, make_float2(p.b_cos, p.b_sin));
                    // |s_hat| > thresh and |arg s_hat| < 0.1 pi, without hypotf / atan2f
                    int accept = (s_hat.x * s_hat.x + s_hat.y * s_hat.y > p.thresh * p.thresh) &&
                                 (s_hat.x > 0.f) && (fabsf(s_hat.y) < 0.32491969623290632616f * s_hat.x);
                    red[110] = (float)accept;
                    if (!accept) {
                        if (S->num_symbols == 16) phy_reset();
                        else S->timer = (int)M2;
                    }
                }
                __syncthreads();
                if (red[110] != 0.f) {
                    // G *= M/sqrt(Na) * B ; smooth |G| and arg G with an order-4 polynomial over the
                    // active subcarriers (liquid ofdmframesync_estimate_eqgain_poly); R = B / G.
                    // coef = P y with the constant matrix P of design.h eqgain_fit_matrix(); every
                    // thread works on its own 8 subcarriers, only the phase unwrap (sequential in the
                    // fft-shifted visiting order) goes through shared memory.
                    const float gsc = (float)M / sqrtf((float)Na);
                    float ya[8];
                    double pv[8][5];                 // own rows of the fit matrix, fetched while the phases unwrap
#pragma unroll
                    for (unsigned int s = 0; s < 8; s++) {
                        const double * pr = p.tb.eqfit_P + (size_t)(ar[s] == 0xffffu ? 0u : ar[s]) * 5;
#pragma unroll
                        for (int r = 0; r < 5; r++) pv[s][r] = __ldg(pr + r);
                    }
#pragma unroll
                    for (unsigned int s = 0; s < 8; s++) {
                        if (ar[s] == 0xffffu) continue;
                        const cf gk = cmul(cscale(g[s], gsc), Bq[s]);
                        ya[s] = sqrtf(gk.x * gk.x + gk.y * gk.y);
                        yph[ar[s]] = atan2_fast(gk.y, gk.x);
                    }
                    __syncthreads();
                    {
                        // unwrap only if some neighbouring pair is more than pi apart (smooth channels: never)
                        int wraps = 0;
                        for (unsigned int n = t + 1; n < Na; n += T) wraps |= fabsf(yph[n] - yph[n - 1]) > PI_F;
                        if (__syncthreads_or(wraps)) {
                            if (wid == 0) warp_unwrap_seg(yph, Na, lane);
                            __syncthreads();
                        }
                    }
                    double ca[10];
#pragma unroll
                    for (int i = 0; i < 10; i++) ca[i] = 0.0;
#pragma unroll
                    for (unsigned int s = 0; s < 8; s++) {
                        if (ar[s] == 0xffffu) continue;
                        const double yav = (double)ya[s], yg = (double)yph[ar[s]];
#pragma unroll
                        for (int r = 0; r < 5; r++) {
                            ca[r] = fma(pv[s][r], yav, ca[r]);
                            ca[5 + r] = fma(pv[s][r], yg, ca[5 + r]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 10; i++) ca[i] = warp_sum_d(ca[i]);
                    if (NW > 1) {
                        if (lane == 0) {
#pragma unroll
                            for (int i = 0; i < 10; i++) dsum[wid * 10 + i] = ca[i];
                        }
                        __syncthreads();
#pragma unroll
                        for (int i = 0; i < 10; i++) {
                            double a = 0.0;
                            for (unsigned int w = 0; w < NW; w++) a += dsum[w * 10 + i];
                            ca[i] = a;
                        }
                    }
#pragma unroll
                    for (unsigned int s = 0; s < 8; s++) {
                        if (rk[s] == 0xffffu) { Rr[s] = make_float2(0.f, 0.f); continue; }
                        // order-4 polynomials in x in [-1/2, 1/2): float Horner is good to ~1e-7 relative
                        const float xv = fxs[s] / (float)M;
                        const float A = fmaf(fmaf(fmaf(fmaf((float)ca[4], xv, (float)ca[3]), xv, (float)ca[2]), xv, (float)ca[1]), xv, (float)ca[0]);
                        float thv = fmaf(fmaf(fmaf(fmaf((float)ca[9], xv, (float)ca[8]), xv, (float)ca[7]), xv, (float)ca[6]), xv, (float)ca[5]);
                        thv = fmaf(-6.28318530717958647692f, rintf(thv * 0.15915494309189533577f), thv);
                        float sn, cs;
                        __sincosf(thv, &sn, &cs);
                        // R = B / G = B conj(G) / |G|^2 with G = A e^{j thv}
                        const float inv = __frcp_rn(A);
                        const cf num = cmulc(Bq[s], make_float2(cs, sn));
                        Rr[s] = make_float2(num.x * inv, num.y * inv);
                    }
                    if (t == 0) {
                        S->state = ST_RX;
                        S->timer = (int)(M + cp + p.backoff);
                        S->num_symbols = 0;
                    }
                }
            }
            PH(12);
            continue;                        // loop top synchronises
        }

        // ---- ST_RX: one OFDM symbol.  Equalise in registers; the pilots go to warp 0
#pragma unroll
        for (unsigned int s = 0; s < 8; s++) {
            v[s] = cmul(v[s], Rr[s]);
            if (pilot_mask & (1u << s)) yc[rk[s] & 0x3fffu] = v[s];
        }
        __syncthreads();
        PH(3);
        prefetch(min(pos + PF, p.nsamples));     // last warp, while warp 0 fits the pilots
        float fit_p0 = 0.f;
        if (wid == 0) {
            float sy, sxy;
            if (Mp <= 64) {
                // at most two pilots per lane (n = lane, lane + 32): phases, unwrap and sums stay in registers
                const bool v0 = lane < Mp, v1 = lane + 32 < Mp;
                float raw0 = 0.f, raw1 = 0.f;
                if (v0) {
                    const float pil = pilot_seq[(ppos + lane) % 255u] ? 1.0f : -1.0f;
                    const cf c = yc[lane];
                    raw0 = atan2_fast(c.y * pil, c.x * pil);
                }
                if (v1) {
                    const float pil = pilot_seq[(ppos + lane + 32) % 255u] ? 1.0f : -1.0f;
                    const cf c = yc[lane + 32];
                    raw1 = atan2_fast(c.y * pil, c.x * pil);
                }
                PH(8);
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
                    const unsigned int q = (ppos + n) % 255u;
                    const float pil = pilot_seq[q] ? 1.0f : -1.0f;
                    const cf c = yc[n];
                    yph[n] = atan2_fast(c.y * pil, c.x * pil);
                }
                __syncwarp();
                warp_unwrap(yph, pilot_x, Mp, false, lane, sy, sxy);
            }
            PH(9);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
            }
            PH(10);
            if (lane == 0) {
                const float np = (float)Mp, sx = p.pilot_sx, sxx = p.pilot_sxx;
                float den = __fsub_rn(__fmul_rn(np, sxx), __fmul_rn(sx, sx));
                float p1 = __fdiv_rn(__fsub_rn(__fmul_rn(np, sxy), __fmul_rn(sx, sy)), den);
                fit_p0 = __fdiv_rn(__fsub_rn(sy, __fmul_rn(p1, sx)), np);
                const float alpha = 0.3f;
                p1 = __fadd_rn(__fmul_rn(alpha, p1), __fmul_rn(1 - alpha, r_p1p));
                red[111] = fit_p0; red[112] = p1;
                // NCO trim (the next symbol is mixed with it)
                uint32_t nd = dth;
                if (r_nsym > 0) {
                    float dphi = fit_p0 - r_phip;
                    while (dphi > PI_F) dphi -= 2 * PI_F;
                    while (dphi < -PI_F) dphi += 2 * PI_F;
                    nd += nco_constrain_small(1e-3f * dphi);
                }
                S->nco_dtheta = nd;
                red[116] = __uint_as_float(nd);
            }
        }
        // can the next event be started right behind this barrier?  (steady state of a payload: the
        // frame goes on, a whole symbol is available, no debug tap)  Its samples must have landed.
        PH(11);
        const unsigned int take_now = (fstate == FS_PAYLOAD) ? min(p.M_data, mod_len - pstart) : 0u;
        // (a speculative worker must pass the loop-top gate before its last payload symbol)
        const bool pipe = (fstate == FS_PAYLOAD) && (pstart + take_now < mod_len) && (p.nsamples - pos >= W) && (p.tap_cap == 0) &&
                          !(role == SW_SPEC && pstart + take_now + p.M_data >= mod_len);
        if (pipe) {
            if (pos + W <= done_frontier) cp_async_wait_group<1>();
            else cp_async_wait_group<0>();
        }
        __syncthreads();
        PH(4);
        if (t == 0) {
            // symbol bookkeeping: off the path of the other threads, who only need p0 / p1 / the NCO step
            S->phi_prime = red[111];
            S->p1_prime = red[112];
            S->num_symbols = r_nsym + 1;
            S->pilot_pos = (ppos + Mp) % 255u;
            S->timer = (int)(M + cp);    // liquid sets this unconditionally (also after a reset below)
        }

        if (pipe) {
            // ---- pipelined tail: the next symbol's samples are mixed and taken through the first FFT
            //      pass in the same stretch of code that derotates and demaps this symbol, so the two
            //      dependency chains interleave and the loop-top barrier + state reload disappear
            // speculative worker: look at the hand-off word now, act on it at the end of the event
            unsigned int hs_peek = 0;
            if (duo && role == SW_SPEC && t == 0) hs_peek = ctl->hs;
            const uint32_t dth2 = __float_as_uint(red[116]);
            const uint32_t th2 = th + adv * dth;             // phase after this event's samples
            const unsigned int off2 = cp - p.backoff;
            cf v2[8];
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const unsigned int j = off2 + t + s * T;
                const cf x = mix_down(stg[(pos + j) & SZM], nco_cexp_fast(th2 + j * dth2));
                v2[s] = x;
                hist[j] = x;
            }
            for (unsigned int jj = t; jj < cp; jj += T) {
                const unsigned int j = (jj < off2) ? jj : jj + M;
                hist[j] = mix_down(stg[(pos + j) & SZM], nco_cexp_fast(th2 + j * dth2));
            }
            {
                const float p0 = red[111], p1 = red[112];
#pragma unroll
                for (unsigned int s = 0; s < 8; s++) {
                    float thv = __fadd_rn(p0, __fmul_rn(p1, fxs[s]));
                    float sn, cs;
                    __sincosf(thv, &sn, &cs);
                    v[s] = cmul(v[s], make_float2(cs, -sn));
                }
            }
            f8_pass<M, 1, 8, -1>(v2, t, tw);
            f8_store<M, 1, 8>(v2, t, fa);
            {
                uint8_t * dst = penc + pstart;
                const float alpha = p.qam_alpha[bps];
#define B2_DEMAP(EXPR)                                                              \
                _Pragma("unroll")                                                   \
                for (unsigned int s = 0; s < 8; s++) {                              \
                    const unsigned int r = rk[s];                                   \
                    if (r < take_now) { const cf x = v[s]; dst[r] = (uint8_t)(EXPR); } \
                }
                if (ms == 40) { B2_DEMAP((x.x > 0 ? 0u : 1u) + (x.y > 0 ? 0u : 2u)) }
                else if (ms == 39) { B2_DEMAP(x.x > 0 ? 0u : 1u) }
                else if (bps == 6) { B2_DEMAP(demod_qam_t<3>(x, alpha)) }
                else if (bps == 4) { B2_DEMAP(demod_qam_t<2>(x, alpha)) }
                else if (bps == 8) { B2_DEMAP(demod_qam_t<4>(x, alpha)) }
                else { B2_DEMAP(demod_qam_t<1>(x, alpha)) }
#undef B2_DEMAP
            }
            if (t == 0) S->payload_sym_idx = pstart + take_now;
            if (duo && role == SW_SPEC && t == 0) {
                unsigned int nr = SW_SPEC;
                if (hs_peek == ((S->my_seq << 2) | HS_ACCEPTED)) {
                    nr = SW_OWNER;
                    S->role = SW_OWNER;
                    if (S->pub_pending) { publish(S->pub_start); S->pub_pending = 0; }
                }
                red[122] = __uint_as_float(nr);
            }
            __syncthreads();                 // first FFT exchange of the next event
            if (duo && role == SW_SPEC) role = __float_as_uint(red[122]);
            PH(5);
            // the next event, as the loop top would have found it
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) v[s] = v2[s];
            pstart += take_now;
            ppos = (ppos + Mp) % 255u;
            th = th2; dth = dth2;
            r_phip = red[111]; r_p1p = red[112]; r_nsym += 1;
            timer = (int)W; adv = W; head = 0; head2 = 0; off = off2;
            pos += W;
            pre = true;
            continue;
        }

        // ---- derotate own subcarriers
        {
            const float p0 = red[111], p1 = red[112];
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {       // null subcarriers carry 0 (their equaliser tap is 0)
                float thv = __fadd_rn(p0, __fmul_rn(p1, fxs[s]));
                float sn, cs;
                __sincosf(thv, &sn, &cs);
                v[s] = cmul(v[s], make_float2(cs, -sn));
            }
        }

        // ---- debug tap of the equalised symbol
        if (p.tap_cap) {
            if (t == 0) red[113] = __uint_as_float(atomicAdd(&p.counters[4], 1u));
            __syncthreads();
            unsigned int slot = __float_as_uint(red[113]);
            if (slot < p.tap_cap) {
#pragma unroll
                for (unsigned int s = 0; s < 8; s++) p.tap_X[(size_t)slot * M + t + s * T] = v[s];
                if (t == 0) { p.tap_chan[slot] = sidx; p.tap_index[slot] = S->sample_index - 1; }
            }
        }

        // ---- ofdmflexframesync layer
        int emit = 0;                       // 1: header invalid, 2: payload complete
        if (fstate == FS_PAYLOAD) {
            // demap; the symbols leave one per byte (packet.cu packs them into the encoded bytes)
            const unsigned int take = min(p.M_data, mod_len - pstart);
            uint8_t * dst = penc + pstart;
            const float alpha = p.qam_alpha[bps];
#define B2_DEMAP(EXPR)                                                          \
            _Pragma("unroll")                                                   \
            for (unsigned int s = 0; s < 8; s++) {                              \
                const unsigned int r = rk[s];                                   \
                if (r < take) { const cf x = v[s]; dst[r] = (uint8_t)(EXPR); } \
            }
            if (ms == 40) { B2_DEMAP((x.x > 0 ? 0u : 1u) + (x.y > 0 ? 0u : 2u)) }
            else if (ms == 39) { B2_DEMAP(x.x > 0 ? 0u : 1u) }
            else if (bps == 6) { B2_DEMAP(demod_qam_t<3>(x, alpha)) }
            else if (bps == 4) { B2_DEMAP(demod_qam_t<2>(x, alpha)) }
            else if (bps == 8) { B2_DEMAP(demod_qam_t<4>(x, alpha)) }
            else { B2_DEMAP(demod_qam_t<1>(x, alpha)) }
#undef B2_DEMAP
            if (t == 0) S->payload_sym_idx = pstart + take;
            if (pstart + take == mod_len) emit = 2;
            PH(5);
        } else {
            // header: BPSK, 288 symbols; EVM is measured on them (framesyncstats_s.evm)
            const unsigned int take = min(p.M_data, 288u - hstart);
            float ev = 0.f;
            // hard bits go straight into the 288-bit header word array (MSB first inside every byte; the
            // array is cleared when a frame starts), one shared-memory atomicOr per 1-bit
            uint32_t * hwords = (uint32_t *)S->header_bits;
#pragma unroll
            for (unsigned int s = 0; s < 8; s++) {
                const unsigned int r = rk[s];
                if (r < take) {
                    const cf x = v[s];
                    const unsigned int b = x.x > 0 ? 0u : 1u;
                    const unsigned int gb = hstart + r;
                    if (b) atomicOr(&hwords[gb >> 5], 1u << (8u * ((gb >> 3) & 3u) + 7u - (gb & 7u)));
                    const float dr = x.x - (b ? -1.0f : 1.0f);
                    ev += dr * dr + x.y * x.y;
                }
            }
            {
                float z0 = 0.f, z1 = 0.f, z2 = 0.f;
                block_sum4(ev, z0, z1, z2);
                if (NW == 1) __syncthreads();
            }
            if (t == 0) { S->evm_hat += ev; S->header_sym_idx = hstart + take; }
            PH(13);
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
#pragma unroll
                    for (int vq = 3; vq >= 0; vq--) {
                        if (lane < 18) {
                            unsigned int j = hwalk[vq];
                            uint8_t mk = ilmask[vq];
                            uint8_t a = hb[2 * lane], b = hb[2 * j + 1];
                            hb[2 * lane] = (uint8_t)((a & ~mk) | (b & mk));
                            hb[2 * j + 1] = (uint8_t)((a & mk) | (b & ~mk));
                        }
                        __syncwarp();
                    }
                    if (lane < 12) {
                        unsigned int vv = ((unsigned int)hb[3 * lane] << 16) | ((unsigned int)hb[3 * lane + 1] << 8) | hb[3 * lane + 2];
                        gsym[lane] = golay2412_decode(vv);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        uint8_t * hd = S->header_dec;
                        for (int gq = 0; gq < 6; gq++) {
                            unsigned int s0 = gsym[2 * gq], s1 = gsym[2 * gq + 1];
                            hd[3 * gq] = (s0 >> 4) & 0xff;
                            hd[3 * gq + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
                            hd[3 * gq + 2] = s1 & 0xff;
                        }
                        uint32_t key = ((uint32_t)hd[14] << 24) | ((uint32_t)hd[15] << 16) | ((uint32_t)hd[16] << 8) | hd[17];
                        int valid = crc32_nibble(hd, 14) == key;
                        S->evm_db = 10 * log10f(S->evm_hat / 288.0f);
                        if (valid && hd[8] != 105) valid = 0;          // protocol id
                        unsigned int plen = ((unsigned int)hd[9] << 8) | hd[10];
                        unsigned int hms = hd[11], check = (hd[12] >> 5) & 7, fec0 = hd[12] & 0x1f, fec1 = hd[13] & 0x1f;
                        unsigned int hbps = dev_mod_bps(hms);
                        if (valid && (hbps == 0 || (check != 6 && check != 1) || !dev_fec_ok(fec0) || !dev_fec_ok(fec1))) valid = 0;
                        unsigned int henc = 0, hmod = 0;
                        if (valid) {
                            henc = dev_fec_enc_len(fec1, dev_fec_enc_len(fec0, plen + (check == 6 ? 4 : 0)));
                            hmod = (8 * henc + hbps - 1) / hbps;
                            if (hmod > p.penc_cap) valid = 0;         // cannot happen with penc_cap at its default
                        }
                        if (valid) {
                            S->ms_payload = hms; S->bps_payload = hbps; S->payload_len = plen;
                            S->check = check; S->fec0 = fec0; S->fec1 = fec1;
                            S->payload_enc_len = henc;
                            S->payload_mod_len = hmod;
                            S->fstate = FS_PAYLOAD;
                            if (duo) {
                                // the frame ends nps symbols from here; the partner's search starts one sample later
                                const unsigned int nps = (hmod + p.M_data - 1) / p.M_data;
                                if (nps >= 2) {
                                    const unsigned long long start = S->sample_index + (unsigned long long)nps * W + 1ull;
                                    if (role == SW_OWNER) publish(start);
                                    else { S->pub_pending = 1; S->pub_start = start; }
                                }
                            }
                        }
                        red[114] = (float)valid;
                    }
                }
                __syncthreads();
                if (red[114] == 0.f) emit = 1;
            }
        }

        PH(14);
        if (emit && duo && role == SW_SPEC) {
            // only an invalid header gets here (the gate settles the role before a last payload symbol): the record
            // cannot wait, so either the hand-off has been settled by now or it is returned to the owner
            if (t == 0) {
                const unsigned int mine = S->my_seq << 2;
                const unsigned int old = atomicCAS((unsigned int *)&ctl->hs, mine | HS_SENT, mine | HS_NONE);
                const unsigned int nr = (old == (mine | HS_ACCEPTED)) ? SW_OWNER : SW_WAIT;
                S->role = nr; S->pub_pending = 0;
                red[123] = __uint_as_float(nr);
            }
            __syncthreads();
            role = __float_as_uint(red[123]);
            if (role == SW_WAIT) continue;   // loop top: barrier, then the gate waits for the next hand-off
        }
        if (emit) {
            // append a frame record (+ the payload symbols) to the output of this launch; thread 0 reserves
            // the slots, one barrier publishes them together with every thread's symbol stores
            if (t == 0) {
                unsigned int e2 = (emit == 2) ? S->payload_enc_len : 0u;
                unsigned int m2 = (emit == 2) ? S->payload_mod_len : 0u;      // symbols, one byte each
                // reserved at the start of this symbol when it is the last payload symbol (emit == 2);
                // a frame dropped for an invalid header (emit == 1) reserves its record slot here
                unsigned int slot = pre_slot;
                unsigned long long offb = pre_off;
                if (!last_payload) {
                    slot = atomicAdd(&p.counters[0], 1u);
                    offb = m2 ? atomicAdd((unsigned long long *)(p.counters + 2), (unsigned long long)((m2 + 15u) & ~15u)) : 0ull;
                }
                int ok = slot < p.recs_cap;
                if (m2 && offb + ((m2 + 15u) & ~15u) > p.arena_cap) ok = 0;
                if (!ok) { atomicOr(&p.counters[1], 1u); red[115] = -1.f; }
                else {
                    FrameRec r;
                    r.channel = p.chan_base + sidx;
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
                    r.payload_offset = offb;
                    p.recs[slot] = r;
                    FrameAux a; a.enc_len = e2; a.sym_bps = (emit == 2) ? S->bps_payload : 0u; a.sym_off = offb;
                    p.aux[slot] = a;
                    red[115] = 1.f;
                    dsum[16 * 10 + 2] = __longlong_as_double((long long)offb);
                }
            }
            __syncthreads();
            if (emit == 2 && red[115] > 0.f) {
                const unsigned long long offb = (unsigned long long)__double_as_longlong(dsum[16 * 10 + 2]);
                const unsigned int m2 = S->payload_mod_len;
                uint4 * dst = (uint4 *)(p.arena + offb);
                const uint4 * src = (const uint4 *)penc;
                for (unsigned int i = t; i < (m2 + 15) / 16; i += T) dst[i] = src[i];
            }
            if (t == 0) {                    // the loop-top barrier orders this against everybody's next reads
                flex_reset();
                S->timer = (int)(M + cp);    // survives the reset, as in liquid
                if (duo) { S->verify = S->sent_seq ? 1u : 0u; S->pub_pending = 0; }
            }
        }
    }

    // ---- store persistent state
    cp_async_wait_group<0>();
    __syncthreads();
    {
        uint32_t * dst = (uint32_t *)(p.st + vs);
        const uint32_t * src = (const uint32_t *)S;
        for (unsigned int i = t; i < sizeof(SyncState) / 4; i += T) dst[i] = src[i];
        cf * gr = p.ring + (size_t)vs * W;
        for (unsigned int i = t; i < W; i += T) gr[i] = hist[i];
        cf * g0 = p.G0 + (size_t)vs * M;
        cf * gR = p.R + (size_t)vs * M;
#pragma unroll
        for (unsigned int s = 0; s < 8; s++) { g0[t + s * T] = G0[t + s * T]; gR[t + s * T] = Rr[s]; }
    }
    if (duo && t == 0) {                     // the partner stops waiting for this worker
        __threadfence();
        ctl->done[wk] = p.launch_id;
    }
}

template <unsigned int M>
static cudaError_t sync8_launch_t(const SyncParams & p, size_t smem_bytes, cudaStream_t st)
{
    static size_t configured[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    if (smem_bytes > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sync8_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        configured[dev] = smem_bytes;
    }
    sync8_kernel<M><<<p.streams * (p.workers == 2 ? 2u : 1u), M / 8, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

size_t sync8_smem_bytes(const SyncParams & p) { return s8_layout(p.M, p.cp, p.M_pilot + p.M_data, p.M_pilot).total; }

bool sync8_supported(unsigned int M) { return M == 256 || M == 512 || M == 1024 || M == 2048 || M == 4096; }

template <unsigned int M>
static int sync8_ctas_t(size_t smem)
{
    int nb = 0;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(sync8_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sync8_kernel<M>, (int)(M / 8), smem) != cudaSuccess) return 0;
    return nb;
}
int sync8_ctas_per_sm(const SyncParams & p)
{
    const size_t smem = s8_layout(p.M, p.cp, p.M_pilot + p.M_data, p.M_pilot).total;
    switch (p.M) {
    case 256:  return sync8_ctas_t<256>(smem);
    case 512:  return sync8_ctas_t<512>(smem);
    case 1024: return sync8_ctas_t<1024>(smem);
    case 2048: return sync8_ctas_t<2048>(smem);
    case 4096: return sync8_ctas_t<4096>(smem);
    default:   return 0;
    }
}

cudaError_t sync8_launch(const SyncParams & p, cudaStream_t st)
{
    const size_t smem = s8_layout(p.M, p.cp, p.M_pilot + p.M_data, p.M_pilot).total;
    switch (p.M) {
    case 256:  return sync8_launch_t<256>(p, smem, st);
    case 512:  return sync8_launch_t<512>(p, smem, st);
    case 1024: return sync8_launch_t<1024>(p, smem, st);
    case 2048: return sync8_launch_t<2048>(p, smem, st);
    case 4096: return sync8_launch_t<4096>(p, smem, st);
    default:   return cudaErrorInvalidValue;
    }
}

} // namespace b2
