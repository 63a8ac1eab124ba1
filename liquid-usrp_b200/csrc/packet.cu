// packet.cu -- payload packet decode: 2-stage (de-interleave -> FEC decode), then CRC-32.
//
// This is liquid's packetizer_decode(), which the reference reaches from inside
// ofdmflexframesync_execute (lib/multichannelrx.cc:194, lib/ofdmtxrx.cc:625) once the last
// payload symbol of a frame has been demodulated.  Integer/bitwise work, HBM/L2 bound:
// one CTA per completed frame, grid-stride over the frames of a launch.
//   interleaver : liquid's 4-pass byte/bit-mask permutation.  Each pass is a set of DISJOINT
//                 swaps (2i <-> 2j+1) whose partner index j follows a column walk of an M x N
//                 grid; the walk is turned into a parallel prefix count so all swaps of a pass
//                 run concurrently.
//   FEC         : none, Hamming(12,8) (thread per codeword pair), Golay(24,12) (thread per
//                 codeword pair), convolutional r1/2 K=7 (hard-decision Viterbi, one warp per
//                 frame: lane L owns states 2L, 2L+1; decisions ballot-packed, 64 bit per step).
//   CRC-32      : 32 lanes x bytewise CRC of a slice, slices merged with x^(8n) mod P products.
#include <mutex>
#include <cub/device/device_radix_sort.cuh>
#include "kernels.h"
#include "fec.cuh"

namespace b2 {

constexpr int PK_THREADS = 128;
constexpr unsigned int PKF_MAX = 4096;              // longest packet the uncoded fast path stages (bytes)
constexpr int PKF_WARPS = 8;
constexpr unsigned int PK_TB_STEPS = 2048;          // traceback staging chunk (steps)

__device__ __forceinline__ unsigned int pk_fec_enc_len(unsigned int scheme, unsigned int n)
{
    switch (scheme) {
    case 6:  return (n * 12 + 7) / 8;
    case 7:  { unsigned int blocks = (n * 8 + 11) / 12; return (blocks * 24 + 7) / 8; }
    case 11: return (2 * (8 * n + 6) + 7) / 8;
    default: return n;
    }
}

// ------------------------------------------------------------------ symbols -> encoded bytes
// ofdmsync8.cu leaves a payload as one demapped symbol per byte; liquid packs them MSB first into
// the encoded message (liquid_repack_bytes in ofdmflexframesync's payload path).  Thread `tid` of
// `nthreads` builds 32-bit words of the byte stream: word w = stream bits [32w, 32w+32).
__device__ __forceinline__ void pack_symbols(const uint8_t * __restrict__ sym, unsigned int mod_len, unsigned int bps,
                                             uint32_t * out0, uint32_t * out1, unsigned int nbytes,
                                             unsigned int tid, unsigned int nthreads)
{
    const unsigned int nwords = (nbytes + 3) / 4;
    for (unsigned int w = tid; w < nwords; w += nthreads) {
        const unsigned int lo = 32 * w, hi = lo + 32;
        const unsigned int d0 = lo / bps, d1 = (hi - 1) / bps;
        unsigned long long acc = 0;
        for (unsigned int d = d0; d <= d1; d++) acc = (acc << bps) | (d < mod_len ? (unsigned long long)sym[d] : 0ull);
        const unsigned int top = (d1 + 1) * bps;
        const uint32_t be = (uint32_t)(acc >> (top - hi));
        const uint32_t le = __byte_perm(be, 0, 0x0123);           // first stream byte at the lowest address
        out0[w] = le;
        if (out1) out1[w] = le;
    }
}

// ------------------------------------------------------------------ block-wide exclusive scan of small counts
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, unsigned int * scratch, unsigned int tid,
                                                        unsigned int * total)
{
    unsigned int lane = tid & 31, wid = tid >> 5;
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned int)o) x += y;
    }
    __syncthreads();
    if (lane == 31) scratch[wid] = x;
    __syncthreads();
    unsigned int base = 0, tot = 0;
    for (unsigned int w = 0; w < blockDim.x / 32; w++) {
        unsigned int s = scratch[w];
        if (w < wid) base += s;
        tot += s;
    }
    *total = tot;
    return base + x - v;
}

// one de-interleaver pass over x[0..n): grid Mi x Ni, bit mask `mask`
__device__ void deinterleave_pass(uint8_t * x, unsigned int n, unsigned int Mi, unsigned int Ni, unsigned int mask,
                                  unsigned int * scratch, unsigned int tid)
{
    const unsigned int n2 = n / 2, c0 = n / 3;
    unsigned int count = 0, base = 0;
    while (count < n2) {
        unsigned int q0 = base + tid * 4;
        unsigned int jv[4], f = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned int q = q0 + k;
            unsigned int t = q / Mi, m = q - t * Mi;
            unsigned int c = (t == 0) ? c0 : (c0 + t) % Ni;
            unsigned int j = m * Ni + c;
            jv[k] = j;
            f += (j < n2) ? 1u : 0u;
        }
        unsigned int tot;
        unsigned int idx = count + block_excl_scan(f, scratch, tid, &tot);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (jv[k] < n2) {
                if (idx < n2) {
                    unsigned int a = x[2 * idx], b = x[2 * jv[k] + 1];
                    x[2 * idx] = (uint8_t)((a & ~mask) | (b & mask));
                    x[2 * jv[k] + 1] = (uint8_t)((a & mask) | (b & ~mask));
                }
                idx++;
            }
        }
        count += tot;
        base += blockDim.x * 4;
        __syncthreads();
    }
}

__device__ void deinterleave(uint8_t * x, unsigned int n, unsigned int * scratch, unsigned int tid)
{
    if (n < 2) return;
    unsigned int Mi = 1 + (unsigned int)floorf(sqrtf((float)n));
    unsigned int Ni = n / Mi;
    while (n >= Mi * Ni) Ni++;
    deinterleave_pass(x, n, Mi, Ni + 8, 0x33, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni + 4, 0x55, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni + 2, 0x0f, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni, 0xff, scratch, tid);
}

// ------------------------------------------------------------------ block codes
__device__ __forceinline__ unsigned int hamming128_decode(unsigned int r)
{
    unsigned int z = 8 * (__popc(r & 0x01f) & 1) + 4 * (__popc(r & 0x1e1) & 1) + 2 * (__popc(r & 0x666) & 1) + (__popc(r & 0xaaa) & 1);
    if (z) r ^= 1u << (12 - z);
    return ((r & 0x200) >> 2) | ((r & 0x0e0) >> 1) | (r & 0x00f);
}

__device__ void hamming128_decode_block(const uint8_t * enc, uint8_t * dec, unsigned int n, unsigned int tid)
{
    unsigned int pairs = n / 2;
    for (unsigned int k = tid; k < pairs; k += blockDim.x) {
        unsigned int e0 = enc[3 * k], e1 = enc[3 * k + 1], e2 = enc[3 * k + 2];
        dec[2 * k] = (uint8_t)hamming128_decode((e0 << 4) | (e1 >> 4));
        dec[2 * k + 1] = (uint8_t)hamming128_decode(((e1 & 0x0f) << 8) | e2);
    }
    if ((n & 1) && tid == 0) {
        unsigned int j = 3 * pairs;
        dec[n - 1] = (uint8_t)hamming128_decode(((unsigned int)enc[j] << 4) | (enc[j + 1] >> 4));
    }
}
__device__ void golay2412_decode_block(const uint8_t * enc, uint8_t * dec, unsigned int n, unsigned int tid)
{
    unsigned int groups = n / 3, r = n % 3;
    for (unsigned int k = tid; k < groups; k += blockDim.x) {
        const uint8_t * e = enc + 6 * k;
        unsigned int s0 = golay2412_decode(((unsigned int)e[0] << 16) | ((unsigned int)e[1] << 8) | e[2]);
        unsigned int s1 = golay2412_decode(((unsigned int)e[3] << 16) | ((unsigned int)e[4] << 8) | e[5]);
        dec[3 * k] = (s0 >> 4) & 0xff;
        dec[3 * k + 1] = ((s0 << 4) & 0xf0) | ((s1 >> 8) & 0x0f);
        dec[3 * k + 2] = s1 & 0xff;
    }
    if (tid < r) {
        const uint8_t * e = enc + 6 * groups + 3 * tid;
        dec[3 * groups + tid] = golay2412_decode(((unsigned int)e[0] << 16) | ((unsigned int)e[1] << 8) | e[2]) & 0xff;
    }
}

// ------------------------------------------------------------------ conv r1/2 K=7 Viterbi (warp 0 of the CTA)
// libfec metric (|expected - received| with hard bits as 0/255), start metrics 0 / 63,
// predecessor with the oldest bit set wins only on strictly smaller metric, chain back from 0.
//
// Layout of the 64 path metrics over the warp: a lane keeps two states that differ in the newest bit
// (slot = bit 0); the other five state bits select the lane through a rotation that advances with the
// trellis step, lane = rotr5(state >> 1, t mod 5) after t steps.  With it the two predecessors j and
// j + 32 of a pair of new states 2j, 2j + 1 sit in the same slot of two lanes that differ in one lane
// bit, so a step costs ONE shuffle (the lanes swap the slot they do not work on) and the new pair stays
// in the lane that computed it.  The traceback walks in the same rotated domain (see below).

// add-compare-select recursion of the whole frame by warp 0: decisions[t] = the 64 survivor choices of step t
__device__ void viterbi27_acs(const uint8_t * enc, unsigned int n, uint2 * decisions, unsigned int tid)
{
    const unsigned int nbits = 8 * n + 6;
    if (tid < 32) {
        const unsigned int L = tid;
        // expected output pair of branch j -> 2j of this lane's butterfly in each of the five phases, as
        // (c1 << 1) | c0 (c0 from 0x6d, c1 from 0x4f: the order the streamed received pairs have below).  Both
        // generators tap the newest and the oldest register bit, so the other three branches expect the
        // complement (j -> 2j+1, j+32 -> 2j) or the same pair (j+32 -> 2j+1).
        unsigned int exw[5];
#pragma unroll
        for (unsigned int tm = 0; tm < 5; tm++) {
            const unsigned int lbv = (L >> (4u - tm)) & 1u;
            unsigned int j = lbv;                                   // state bit 0 = the slot this lane works on
#pragma unroll
            for (unsigned int k = 1; k <= 4; k++) j |= ((L >> ((k - 1u + 5u - tm) % 5u)) & 1u) << k;
            const unsigned int reg = j << 1;
            exw[tm] = (((unsigned int)__popc(reg & 0x4f) & 1u) << 1) | ((unsigned int)__popc(reg & 0x6d) & 1u);
        }
        unsigned int me = (L == 0) ? 0u : 63u, mo = 63u;      // metrics of states 2L, 2L+1 (phase 0)
        // received bit pairs come 16 per 32-bit word (the buffers are 16-byte aligned and padded).  A word is
        // byte-swapped and bit-reversed once, after which pair q sits at bits 2q, 2q+1 as (r1 << 1) | r0 and the
        // stream is consumed by shifting; the next word is fetched while the current one is consumed.
        const uint32_t * e32 = (const uint32_t *)enc;
        const unsigned int wmax = (2u * nbits + 31u) / 32u - 1u;             // last word that holds pairs
        unsigned long long buf = __brev(__byte_perm(e32[0], 0u, 0x0123u));   // pairs not yet consumed, oldest at bit 0
        unsigned int avail = 16, widx = 1;
        uint32_t nxt = e32[min(1u, wmax)];
        auto step = [&](const unsigned int tm, const unsigned int tt) {
            const unsigned int r = (unsigned int)buf & 3u;
            buf >>= 2;
            const unsigned int lbv = (L >> (4u - tm)) & 1u;
            const unsigned int recv = __shfl_xor_sync(0xffffffffu, lbv ? me : mo, 16u >> tm);
            const unsigned int m_lo = lbv ? recv : me, m_hi = lbv ? mo : recv;
            const unsigned int x = 255u * __popc(exw[tm] ^ r), y = 510u - x;
            unsigned int m0 = m_lo + x, m1 = m_hi + y;
            const unsigned int d_e = m1 < m0;
            me = min(m0, m1);
            m0 = m_lo + y; m1 = m_hi + x;
            const unsigned int d_o = m1 < m0;
            mo = min(m0, m1);
            const unsigned int be = __ballot_sync(0xffffffffu, d_e), bo = __ballot_sync(0xffffffffu, d_o);
            if (L == 0) decisions[tt] = make_uint2(be, bo);
        };
        unsigned int t0 = 0;
        for (; t0 + 5 <= nbits; t0 += 5) {
            // top the pair buffer up (branch-free: the trellis below is one dependent chain per warp)
            const bool ref = avail < 5;
            const unsigned long long add = (unsigned long long)__brev(__byte_perm(nxt, 0u, 0x0123u)) << (2u * avail);
            buf |= ref ? add : 0ull;
            avail += ref ? 16u : 0u;
            widx += ref ? 1u : 0u;
            const uint32_t ld = e32[min(widx, wmax)];
            nxt = ref ? ld : nxt;
            avail -= 5;
#pragma unroll
            for (unsigned int tm = 0; tm < 5; tm++) step(tm, t0 + tm);
        }
        if (t0 < nbits) {                    // fewer than five steps left
            if (avail < 5) buf |= (unsigned long long)__brev(__byte_perm(nxt, 0u, 0x0123u)) << (2u * avail);
#pragma unroll
            for (unsigned int tm = 0; tm < 4; tm++)
                if (t0 + tm < nbits) step(tm, t0 + tm);
        }
    }
}

// exact traceback from the terminated end state (one thread: the chain is dependent from step to step)
__device__ void viterbi27_traceback(uint8_t * dec, unsigned int n, const uint2 * decisions, uint2 * stage, unsigned int tid, bool in_smem,
                                    unsigned int tb_steps = PK_TB_STEPS)
{
    const unsigned int nbits = 8 * n + 6;
    for (unsigned int i = tid; i < n; i += blockDim.x) dec[i] = 0;
    __syncthreads();
    // traceback, in the rotated lane domain: the lane y of the current state stays in place from one step to
    // the next except for ONE bit (position pb, advancing with the step) that is replaced by the decision bit,
    // and the bit it replaces is the slot of the next (older) state.  Decoded bit of step t = slot.
    // Decisions come in chunks staged through shared memory (or sit there already).
    __shared__ unsigned int tb_state;                        // y | slot << 5 | pb << 6
    if (tid == 0) tb_state = ((5u - nbits % 5u) % 5u) << 6;
    auto tb_walk = [&](const uint2 * dsrc, unsigned int base, unsigned int hi_, unsigned int lo_) {
        unsigned int y = tb_state & 31u, slot = (tb_state >> 5) & 1u, pb = tb_state >> 6, acc = 0;
        auto one = [&](const uint2 d) {
            const unsigned int bit = ((slot ? d.y : d.x) >> y) & 1u;
            const unsigned int old = (y >> pb) & 1u;
            y = (y & ~(1u << pb)) | (bit << pb);
            slot = old;
            pb = (pb == 4u) ? 0u : pb + 1u;
        };
        unsigned int t = hi_;
        // down to a byte boundary (only the flush bits at the very end of a frame start off one)
        while (t > lo_ && (t & 7u)) {
            --t;
            acc |= slot << (7u - (t & 7u));
            one(dsrc[t - base]);
        }
        if (hi_ & 7u) { if (t < 8 * n && (t & 7u) == 0) dec[t >> 3] = (uint8_t)acc; }
        // whole bytes: the eight decision words are fetched before the dependent walk through them
        while (t > lo_) {
            uint2 d[8];
#pragma unroll
            for (unsigned int k = 0; k < 8; k++) d[k] = dsrc[t - 1u - k - base];
            acc = 0;
#pragma unroll
            for (unsigned int k = 0; k < 8; k++) {
                acc |= slot << k;               // step t-1-k carries bit 7 - ((t-1-k) & 7) = k of its byte
                one(d[k]);
            }
            t -= 8;
            if (t < 8 * n) dec[t >> 3] = (uint8_t)acc;
        }
        tb_state = y | (slot << 5) | (pb << 6);
    };
    unsigned int hi = nbits;
    if (in_smem) {
        __syncthreads();
        if (tid == 0) tb_walk(decisions, 0, nbits, 0);       // one walk, no staging
        hi = 0;
    }
    while (hi > 0) {
        unsigned int lo = hi > tb_steps ? ((hi - tb_steps + 7u) & ~7u) : 0;     // chunks end on byte boundaries
        __syncthreads();
        for (unsigned int i = lo + tid; i < hi; i += blockDim.x) stage[i - lo] = decisions[i];
        __syncthreads();
        // every `lo` is a multiple of 8, so a decoded byte never straddles two chunks; the flush bits
        // (t >= 8n) of the first chunk carry no data
        if (tid == 0) tb_walk(stage, lo, hi, lo);
        hi = lo;
    }
    __syncthreads();
}

__device__ void viterbi27_decode(const uint8_t * enc, uint8_t * dec, unsigned int n, uint2 * decisions,
                                 uint2 * stage, unsigned int tid, bool in_smem, unsigned int tb_steps = PK_TB_STEPS)
{
    viterbi27_acs(enc, n, decisions, tid);
    viterbi27_traceback(dec, n, decisions, stage, tid, in_smem, tb_steps);
}

// ------------------------------------------------------------------ speculative traceback, one short walk per thread
// The exact traceback is ONE thread following 8n + 6 dependent steps (a quarter of the decoder's instructions, at 1/32 of a
// warp's width).  Survivor paths of a K = 7 code merge within a few constraint lengths, so `nth` threads each take a
// byte-aligned slice of the steps [lo, keep_hi): thread th starts VIT_TB_OV steps above its slice from an arbitrary state
// (state 0; the true end state where that point is the terminated end of the frame), walks down discarding bits until it
// reaches its slice -- by then it is on the maximum-likelihood path with overwhelming probability -- and keeps the bits of
// its slice.  SPECULATIVE like the segmented recursion below: the caller keeps the result only if the CRC-32 passes.
// Decision of step t at dsrc[t - base] (global memory, read through L2: other threads of the CTA wrote them).
constexpr unsigned int VIT_TB_OV = 128;
__device__ void viterbi27_traceback_spec(const uint2 * dsrc, unsigned int base, unsigned int lo, unsigned int keep_hi, unsigned int hi,
                                         uint8_t * dec, unsigned int wlimit, unsigned int th, unsigned int nth)
{
    const unsigned int sl = (((keep_hi - lo) + nth - 1u) / nth + 7u) & ~7u;
    const unsigned int a = lo + th * sl;
    if (a >= keep_hi) return;
    const unsigned int b = min(keep_hi, a + sl);
    const unsigned int S = min(hi, b + VIT_TB_OV);
    unsigned int y = 0, slot = 0, pb = (5u - S % 5u) % 5u, acc = 0;
    auto one = [&](const uint2 d) {
        const unsigned int bit = ((slot ? d.y : d.x) >> y) & 1u;
        const unsigned int old = (y >> pb) & 1u;
        y = (y & ~(1u << pb)) | (bit << pb);
        slot = old;
        pb = (pb == 4u) ? 0u : pb + 1u;
    };
    unsigned int t = S;
    while (t & 7u) {                             // only from the unaligned end of the frame: flush bits, no data
        --t;
        one(__ldcg(dsrc + (t - base)));
    }
    while (t > a) {
        uint2 d[8];
#pragma unroll
        for (unsigned int k = 0; k < 8; k++) d[k] = __ldcg(dsrc + (t - 1u - k - base));
        acc = 0;
#pragma unroll
        for (unsigned int k = 0; k < 8; k++) {
            acc |= slot << k;
            one(d[k]);
        }
        t -= 8;
        if (t < b && t < wlimit) dec[t >> 3] = (uint8_t)acc;
    }
}

// ------------------------------------------------------------------ the same decoder, four trellis segments at once
// A conv-coded 1200-byte frame is a 9 638-step dependent chain (0.68 ms on one warp), and since the frame-parallel
// synchroniser it is what bounds the conv-coded configuration.  Survivor paths of a K = 7 code merge within a few
// constraint lengths, so the four warps of the CTA each take a quarter of the trellis: warp p runs the add-compare-select
// recursion from VIT_OV steps before its segment (all path metrics equal: after the overlap the metric DIFFERENCES, and
// with them the decisions, are those of the full recursion) to VIT_OV steps behind it, and traces back from an arbitrary
// state at the end of that overlap (after which the path has merged with the maximum-likelihood one), keeping only the
// bits of its own segment.  "Merged" is overwhelmingly likely, not certain -- so the result is SPECULATIVE: the caller keeps
// it only if the packet's CRC-32 passes and otherwise runs the exact full-frame decoder above, which makes the output
// identical to libfec's except for a CRC collision.  Packets without a CRC always take the exact decoder.
constexpr unsigned int VIT_OV = 160;                 // overlap on either side of a segment: 23 constraint lengths, a multiple of 80

// true: dec holds the speculative decode; false: the frame does not fit this formulation (too short / workspace too small)
__device__ bool viterbi27_decode_par(const uint8_t * enc, uint8_t * dec, unsigned int n, uint2 * decisions, size_t cap,
                                     unsigned int tid)
{
    constexpr unsigned int NWP = PK_THREADS / 32;
    const unsigned int nbits = 8 * n + 6;
    // segments are multiples of 80 steps: whole groups of five trellis phases, whole 32-bit words of received pairs, whole bytes
    const unsigned int seg = ((nbits + NWP - 1) / NWP + 79u) / 80u * 80u;
    if (seg < 4 * VIT_OV || (size_t)NWP * (seg + 2 * VIT_OV) > cap) return false;
    const unsigned int wp = tid >> 5, L = tid & 31u;
    const unsigned int a = wp * seg, b = min(nbits, a + seg);
    const bool active = a < nbits;
    const unsigned int s0 = (wp == 0 || !active) ? 0u : a - VIT_OV;              // first / one-past-last step of this warp's recursion
    const unsigned int e0 = !active ? 0u : ((b == nbits || b + VIT_OV >= nbits) ? nbits : b + VIT_OV);
    uint2 * dloc = decisions + (size_t)wp * (seg + 2 * VIT_OV);                   // decisions of step t at dloc[t - s0]
    if (active) {
        unsigned int exw[5];
#pragma unroll
        for (unsigned int tm = 0; tm < 5; tm++) {
            const unsigned int lbv = (L >> (4u - tm)) & 1u;
            unsigned int j = lbv;
#pragma unroll
            for (unsigned int k = 1; k <= 4; k++) j |= ((L >> ((k - 1u + 5u - tm) % 5u)) & 1u) << k;
            const unsigned int reg = j << 1;
            exw[tm] = (((unsigned int)__popc(reg & 0x4f) & 1u) << 1) | ((unsigned int)__popc(reg & 0x6d) & 1u);
        }
        // the encoder starts in state 0; a segment that starts inside the frame knows nothing: all metrics equal
        unsigned int me = (wp == 0) ? ((L == 0) ? 0u : 63u) : 0u, mo = (wp == 0) ? 63u : 0u;
        const uint32_t * e32 = (const uint32_t *)enc;
        const unsigned int wmax = (2u * nbits + 31u) / 32u - 1u;
        unsigned int widx = s0 / 16u;                                            // s0 is a multiple of 16 pairs
        unsigned long long buf = __brev(__byte_perm(e32[min(widx, wmax)], 0u, 0x0123u));
        unsigned int avail = 16;
        widx++;
        uint32_t nxt = e32[min(widx, wmax)];
        auto step = [&](const unsigned int tm, const unsigned int tt) {
            const unsigned int r = (unsigned int)buf & 3u;
            buf >>= 2;
            const unsigned int lbv = (L >> (4u - tm)) & 1u;
            const unsigned int recv = __shfl_xor_sync(0xffffffffu, lbv ? me : mo, 16u >> tm);
            const unsigned int m_lo = lbv ? recv : me, m_hi = lbv ? mo : recv;
            const unsigned int x = 255u * __popc(exw[tm] ^ r), y = 510u - x;
            unsigned int m0 = m_lo + x, m1 = m_hi + y;
            const unsigned int d_e = m1 < m0;
            me = min(m0, m1);
            m0 = m_lo + y; m1 = m_hi + x;
            const unsigned int d_o = m1 < m0;
            mo = min(m0, m1);
            const unsigned int be = __ballot_sync(0xffffffffu, d_e), bo = __ballot_sync(0xffffffffu, d_o);
            if (L == 0) dloc[tt - s0] = make_uint2(be, bo);
        };
        unsigned int t0 = s0;
        for (; t0 + 5 <= e0; t0 += 5) {
            const bool ref = avail < 5;
            const unsigned long long add = (unsigned long long)__brev(__byte_perm(nxt, 0u, 0x0123u)) << (2u * avail);
            buf |= ref ? add : 0ull;
            avail += ref ? 16u : 0u;
            widx += ref ? 1u : 0u;
            const uint32_t ld = e32[min(widx, wmax)];
            nxt = ref ? ld : nxt;
            avail -= 5;
#pragma unroll
            for (unsigned int tm = 0; tm < 5; tm++) step(tm, t0 + tm);
        }
        if (t0 < e0) {                       // fewer than five steps left: only at the very end of the frame
            if (avail < 5) buf |= (unsigned long long)__brev(__byte_perm(nxt, 0u, 0x0123u)) << (2u * avail);
#pragma unroll
            for (unsigned int tm = 0; tm < 4; tm++)
                if (t0 + tm < e0) step(tm, t0 + tm);
        }
    }
    for (unsigned int i = tid; i < n; i += blockDim.x) dec[i] = 0;
    __syncthreads();
    if (active) {
        // traceback of [a, e0), bits of [a, b) kept: the warp's 32 lanes each walk a slice (see above)
        __syncwarp();
        viterbi27_traceback_spec(dloc, s0, a, min(b, nbits), e0, dec, min(b, 8u * n), L, 32u);
    }
    __syncthreads();
    return true;
}

// ------------------------------------------------------------------ CRC-32 (reflected 0xEDB88320)
__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// x^(8*nbytes) mod P
__device__ uint32_t crc_x8n(unsigned int nbytes)
{
    uint32_t p = 1u << 31;                   // x^0
    uint32_t sq = crc_multmodp(1u << 30, 1u << 30);   // x^2
    sq = crc_multmodp(sq, sq);               // x^4
    sq = crc_multmodp(sq, sq);               // x^8
    while (nbytes) {
        if (nbytes & 1u) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
        nbytes >>= 1;
    }
    return p;
}
// CRC-32 of m[0..n) computed by warp 0; result valid in lane 0
__device__ uint32_t crc32_warp(const uint8_t * m, unsigned int n, unsigned int lane)
{
    unsigned int per = (n + 31) / 32;
    unsigned int lo = min(n, lane * per), hi = min(n, lo + per);
    uint32_t key = (lane == 0) ? ~0u : 0u;                 // only the first slice carries the preset
    for (unsigned int i = lo; i < hi; i++) {
        key ^= m[i];
#pragma unroll
        for (int j = 0; j < 8; j++) key = (key >> 1) ^ (0xEDB88320u & (0u - (key & 1u)));
    }
    // merge: reg(A||B) = reg(A) * x^(8|B|) + reg0(B)
    uint32_t xp = crc_x8n(per);
    uint32_t acc = 0;
    for (unsigned int l = 0; l < 32; l++) {
        uint32_t k = __shfl_sync(0xffffffffu, key, l);
        unsigned int llo = min(n, l * per), lhi = min(n, llo + per);
        unsigned int len = lhi - llo;
        if (lane == 0 && (len || l == 0)) {
            uint32_t sh = (len == per) ? xp : crc_x8n(len);
            acc = (l == 0) ? k : (crc_multmodp(sh, acc) ^ k);
        }
    }
    return ~acc;
}

// ------------------------------------------------------------------ kernel
// Two shapes of the same kernel are launched per chunk, and the device-side frame count of the launch decides which one works
// (the other returns at once):
//   NT = 128  a launch with at most one frame per CTA is bound by the LATENCY of a frame: four warps per frame, the
//             segmented recursion for conv-coded frames.  64 registers, 8 CTAs per SM (tighter budgets measured slower).
//   NT = 32   a launch with more frames than that is bound by how many trellis recursions run side by side -- each is one
//             warp's dependent chain at a quarter of an instruction per cycle, and in the 128-thread shape the other three
//             warps of the CTA only hold registers meanwhile.  One warp per frame, up to 32 CTAs per SM.
template <unsigned int NT>
__global__ void __launch_bounds__(NT, NT == 32 ? 24 : 8) packet_decode_kernel(const PacketParams p, uint2 * vit_ws, size_t vit_ws_stride,
                                                                               int * vit_locks, unsigned int vit_slots)
{
    constexpr unsigned int TB_STEPS = NT == 32 ? 512u : PK_TB_STEPS;         // traceback staging chunk of the exact decoder
    __shared__ unsigned int scratch[NT / 32];
    __shared__ uint2 stage[TB_STEPS];
    __shared__ int s_valid;
    const unsigned int tid = threadIdx.x;
    const unsigned int nrec = p.range[1].nrec;
    {
        const unsigned int nfr = nrec - p.range[0].nrec;
        const bool wide = p.vit_parallel == 2 || (p.vit_parallel == 1 && nfr > p.vit_split);     // the NT = 32 shape works
        if (wide != (NT == 32)) return;
    }
    for (unsigned int ri = p.range[0].nrec + blockIdx.x; ri < nrec; ri += gridDim.x) {
        FrameRec * rec = p.recs + ri;
        if (!rec->header_valid) continue;
        const unsigned int plen = rec->payload_len, check = rec->check, fec0 = rec->fec0, fec1 = rec->fec1;
        if (fec0 == 1 && fec1 == 1 && plen + 4 <= PKF_MAX) continue;      // done by packet_plain_kernel
        if (p.aux[ri].sym_bps == 0xffffffffu) continue;                  // payload was not stored (did not fit the arena)
        const unsigned long long off = p.aux[ri].sym_off;        // encoded payload: arena + work buffer
        const unsigned int crc_len = (check == 6) ? 4u : 0u;
        const unsigned int n0 = plen + crc_len;
        const unsigned int e0 = pk_fec_enc_len(fec0, n0);
        const unsigned int e1 = pk_fec_enc_len(fec1, e0);
        uint8_t * A = p.arena + off;
        uint8_t * Bf = p.scratch + off;
        uint8_t * D = p.decoded + rec->payload_offset;           // n0 bytes
        // The Viterbi decision workspace is shared by every handle of the device (kernels of different
        // handles may run concurrently): a CTA that needs it claims a free slot and releases it after
        // the frame.  Frames without a convolutional stage never touch it.
        __shared__ unsigned int s_slot;
        // typical frames use this CTA's own region of the handle's workspace (p.vit_local); only frames with a
        // longer trellis take a slot of the device-wide one
        uint2 * ws = nullptr;
        const unsigned long long max_steps = 8ull * ((fec1 == 11) ? e0 : n0) + 6;      // longest trellis of this frame
        const bool conv = (fec0 == 11 || fec1 == 11);
        const bool ws_local = conv && p.vit_local && max_steps <= p.vit_local_steps;
        const bool need_ws = conv && !ws_local;
        const size_t ws_cap = ws_local ? p.vit_local_steps : vit_ws_stride;
        if (ws_local) ws = p.vit_local + (size_t)blockIdx.x * p.vit_local_steps;
        if (need_ws) {
            if (tid == 0) {
                unsigned int sl = blockIdx.x % vit_slots;
                while (atomicCAS(&vit_locks[sl], 0, 1) != 0) sl = (sl + 1) % vit_slots;
                s_slot = sl;
            }
            __syncthreads();
            ws = vit_ws + (size_t)s_slot * vit_ws_stride;
        }
        int ok = 1;
        const unsigned int sym_bps = p.aux[ri].sym_bps;
        if (sym_bps) {
            // the arena holds demapped symbols: pack them, then work in place as before
            const unsigned int mod_len = (8 * e1 + sym_bps - 1) / sym_bps;
            pack_symbols(A, mod_len, sym_bps, (uint32_t *)Bf, nullptr, e1, tid, blockDim.x);
            __syncthreads();
            for (unsigned int i = tid; i < (e1 + 3) / 4; i += blockDim.x) ((uint32_t *)A)[i] = ((const uint32_t *)Bf)[i];
            __syncthreads();
        }
        // stage 1 (outer code)
        uint8_t * s1out = A;
        if (fec1 != 1) {
            deinterleave(A, e1, scratch, tid);
            __syncthreads();
            if (fec1 == 6) hamming128_decode_block(A, Bf, e0, tid);
            else if (fec1 == 7) golay2412_decode_block(A, Bf, e0, tid);
            else if (8ull * e0 + 6 <= ws_cap) viterbi27_decode(A, Bf, e0, ws, stage, tid, false, TB_STEPS);
            else ok = 0;
            s1out = Bf;
            __syncthreads();
        }
        // stage 0 (inner code)
        if (fec0 != 1) {
            deinterleave(s1out, e0, scratch, tid);
            __syncthreads();
            if (fec0 == 6) hamming128_decode_block(s1out, D, n0, tid);
            else if (fec0 == 7) golay2412_decode_block(s1out, D, n0, tid);
            else if (8ull * n0 + 6 <= ws_cap) {
                // speculative decode first, kept only if the CRC confirms it.  The 128-thread shape (a launch bound by the
                // latency of one frame) runs four trellis segments at once (13 % more work); the 32-thread shape the exact
                // recursion; both with the traceback spread over all threads.
                // If the CRC fails, the exact traceback (over the decisions already there) or the exact decoder follows.
                bool done = false, have_acs = false;
                if (crc_len && p.vit_parallel) {
                    bool segmented = false;
                    if constexpr (NT == 128) {
                        if (p.vit_parallel == 1) segmented = viterbi27_decode_par(s1out, D, n0, ws, ws_cap, tid);
                    }
                    if (!segmented) {
                        viterbi27_acs(s1out, n0, ws, tid);
                        __syncthreads();
                        viterbi27_traceback_spec(ws, 0, 0, 8u * n0 + 6u, 8u * n0 + 6u, D, 8u * n0, tid, blockDim.x);
                        __syncthreads();
                        have_acs = true;
                    }
                    if (tid < 32) {
                        const uint32_t c = crc32_warp(D, plen, tid);
                        if (tid == 0) {
                            const uint32_t key = ((uint32_t)D[plen] << 24) | ((uint32_t)D[plen + 1] << 16) | ((uint32_t)D[plen + 2] << 8) | D[plen + 3];
                            s_valid = (c == key);
                        }
                    }
                    __syncthreads();
                    done = (s_valid != 0);
                    __syncthreads();
                }
                if (!done && have_acs) viterbi27_traceback(D, n0, ws, stage, tid, false, TB_STEPS);
                else if (!done) viterbi27_decode(s1out, D, n0, ws, stage, tid, false, TB_STEPS);
            }            else ok = 0;
        } else {
            for (unsigned int i = tid; i < n0; i += blockDim.x) D[i] = s1out[i];
        }
        __syncthreads();
        if (tid < 32) {
            int valid = ok;
            if (ok && crc_len) {
                uint32_t c = crc32_warp(D, plen, tid);
                if (tid == 0) {
                    uint32_t key = ((uint32_t)D[plen] << 24) | ((uint32_t)D[plen + 1] << 16) | ((uint32_t)D[plen + 2] << 8) | D[plen + 3];
                    valid = (c == key);
                }
            }
            if (tid == 0) s_valid = valid;
        }
        __syncthreads();
        if (tid == 0) {
            rec->payload_valid = s_valid;
            if (need_ws) { __threadfence(); atomicExch(&vit_locks[s_slot], 0); }
        }
        __syncthreads();
    }
}

// ================================================================== packet ENCODE (transmit side)
// liquid packetizer_encode + the ofdmflexframegen header/payload symbol mapping, as reached from
// ofdmflexframegen_assemble (lib/multichanneltx.cc:188, lib/ofdmtxrx.cc:320,380):
//   payload: msg || CRC-32 -> fec0 -> interleave -> fec1 -> interleave -> bps-bit symbols
//   header : 8 user + 6 internal bytes || CRC-32 -> Golay(24,12) -> interleave -> scramble -> 288 bits
__device__ void interleave(uint8_t * x, unsigned int n, unsigned int * scratch, unsigned int tid)
{
    if (n < 2) return;
    unsigned int Mi = 1 + (unsigned int)floorf(sqrtf((float)n));
    unsigned int Ni = n / Mi;
    while (n >= Mi * Ni) Ni++;
    deinterleave_pass(x, n, Mi, Ni, 0xff, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni + 2, 0x0f, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni + 4, 0x55, scratch, tid);
    deinterleave_pass(x, n, Mi, Ni + 8, 0x33, scratch, tid);
}

__device__ __forceinline__ unsigned int hamming128_encode(unsigned int s)
{
    unsigned int c = ((s & 0x80) << 2) | ((s & 0x70) << 1) | (s & 0x0f);
    c |= (__popc(c & 0x2aa) & 1) << 11;
    c |= (__popc(c & 0x266) & 1) << 10;
    c |= (__popc(c & 0x0e1) & 1) << 8;
    c |= (__popc(c & 0x00f) & 1) << 4;
    return c;
}
__device__ __forceinline__ unsigned int golay2412_encode(unsigned int s)
{
    s &= 0xfff;
    return (golay_mul_P(s) << 12) | s;
}

__device__ void fec_encode_block(unsigned int scheme, const uint8_t * dec, uint8_t * enc, unsigned int n, unsigned int tid)
{
    if (scheme == 6) {                                  // Hamming(12,8): 2 bytes -> 3 bytes
        unsigned int pairs = n / 2;
        for (unsigned int k = tid; k < pairs; k += blockDim.x) {
            unsigned int m0 = hamming128_encode(dec[2 * k]), m1 = hamming128_encode(dec[2 * k + 1]);
            enc[3 * k] = (m0 >> 4) & 0xff;
            enc[3 * k + 1] = ((m0 << 4) & 0xf0) | ((m1 >> 8) & 0x0f);
            enc[3 * k + 2] = m1 & 0xff;
        }
        if ((n & 1) && tid == 0) {
            unsigned int m0 = hamming128_encode(dec[n - 1]);
            enc[3 * pairs] = (m0 & 0x0ff0) >> 4;
            enc[3 * pairs + 1] = (m0 & 0x000f) << 4;
        }
    } else if (scheme == 7) {                           // Golay(24,12): 3 bytes -> 6 bytes
        unsigned int groups = n / 3, r = n % 3;
        for (unsigned int k = tid; k < groups; k += blockDim.x) {
            unsigned int s0 = ((unsigned int)dec[3 * k] << 4) | (dec[3 * k + 1] >> 4);
            unsigned int s1 = (((unsigned int)dec[3 * k + 1] & 0x0f) << 8) | dec[3 * k + 2];
            unsigned int v0 = golay2412_encode(s0), v1 = golay2412_encode(s1);
            uint8_t * e = enc + 6 * k;
            e[0] = (v0 >> 16) & 0xff; e[1] = (v0 >> 8) & 0xff; e[2] = v0 & 0xff;
            e[3] = (v1 >> 16) & 0xff; e[4] = (v1 >> 8) & 0xff; e[5] = v1 & 0xff;
        }
        if (tid < r) {
            unsigned int v0 = golay2412_encode(dec[3 * groups + tid]);
            uint8_t * e = enc + 6 * groups + 3 * tid;
            e[0] = (v0 >> 16) & 0xff; e[1] = (v0 >> 8) & 0xff; e[2] = v0 & 0xff;
        }
    } else if (scheme == 11) {                          // conv r1/2 K=7: output byte j <- input bits 4j .. 4j+3
        unsigned int nbits = 8 * n + 6, nout = (2 * nbits + 7) / 8;
        for (unsigned int j = tid; j < nout; j += blockDim.x) {
            unsigned int v = 0;
            for (unsigned int q = 0; q < 4; q++) {
                unsigned int t = 4 * j + q;
                unsigned int o = 0;
                if (t < nbits) {
                    // shift register after input bit t: bits t-6 .. t (bit t in the LSB)
                    unsigned int sr = 0;
                    for (int b = 6; b >= 0; b--) {
                        int ti = (int)t - b;
                        unsigned int bit = (ti >= 0 && (unsigned int)ti < 8 * n) ? (dec[ti >> 3] >> (7 - (ti & 7))) & 1u : 0u;
                        sr = (sr << 1) | bit;
                    }
                    o = ((__popc(sr & 0x6d) & 1) << 1) | (__popc(sr & 0x4f) & 1);
                }
                v = (v << 2) | o;
            }
            enc[j] = (uint8_t)v;
        }
    } else {
        for (unsigned int i = tid; i < n; i += blockDim.x) enc[i] = dec[i];
    }
}

__global__ void __launch_bounds__(PK_THREADS) packet_encode_kernel(const EncodeParams p)
{
    __shared__ unsigned int scratch[PK_THREADS / 32];
    __shared__ uint8_t hbuf[2][40];
    const unsigned int tid = threadIdx.x;
    for (unsigned int fi = blockIdx.x; fi < p.nframes; fi += gridDim.x) {
        const EncodeJob job = p.jobs[fi];
        // ---- header: 14 bytes + CRC -> Golay -> interleave -> scramble -> 288 bits
        if (tid < 8) hbuf[0][tid] = job.header[tid];
        if (tid == 8) {
            hbuf[0][8] = 105;                                       // protocol id (104 + packetizer version)
            hbuf[0][9] = (job.payload_len >> 8) & 0xff;
            hbuf[0][10] = job.payload_len & 0xff;
            hbuf[0][11] = (uint8_t)job.mod;
            hbuf[0][12] = (uint8_t)(((job.check & 7) << 5) | (job.fec0 & 0x1f));
            hbuf[0][13] = (uint8_t)(job.fec1 & 0x1f);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t c = crc32_warp(hbuf[0], 14, tid);
            if (tid == 0) { hbuf[0][14] = c >> 24; hbuf[0][15] = (c >> 16) & 0xff; hbuf[0][16] = (c >> 8) & 0xff; hbuf[0][17] = c & 0xff; }
        }
        __syncthreads();
        fec_encode_block(7, hbuf[0], hbuf[1], 18, tid);
        __syncthreads();
        interleave(hbuf[1], 36, scratch, tid);
        __syncthreads();
        {
            const uint8_t mask[4] = {0xb4, 0x6a, 0x8b, 0x45};
            uint8_t * hm = p.header_mod + (size_t)job.slot * 288;
            for (unsigned int i = tid; i < 288; i += blockDim.x) {
                unsigned int byte = hbuf[1][i >> 3] ^ mask[(i >> 3) & 3];
                hm[i] = (byte >> (7 - (i & 7))) & 1u;
            }
        }
        // ---- payload
        const unsigned int plen = job.payload_len, crc_len = (job.check == 6) ? 4u : 0u;
        const unsigned int n0 = plen + crc_len;
        const unsigned int e0 = pk_fec_enc_len(job.fec0, n0), e1 = pk_fec_enc_len(job.fec1, e0);
        uint8_t * A = p.work0 + (size_t)job.slot * p.work_stride;
        uint8_t * B = p.work1 + (size_t)job.slot * p.work_stride;
        const uint8_t * msg = p.payloads + job.payload_offset;
        for (unsigned int i = tid; i < plen; i += blockDim.x) A[i] = msg[i];
        __syncthreads();
        if (crc_len && tid < 32) {
            uint32_t c = crc32_warp(A, plen, tid);
            if (tid == 0) { A[plen] = c >> 24; A[plen + 1] = (c >> 16) & 0xff; A[plen + 2] = (c >> 8) & 0xff; A[plen + 3] = c & 0xff; }
        }
        __syncthreads();
        fec_encode_block(job.fec0, A, B, n0, tid);
        __syncthreads();
        if (job.fec0 != 1) interleave(B, e0, scratch, tid);
        __syncthreads();
        fec_encode_block(job.fec1, B, A, e0, tid);
        __syncthreads();
        if (job.fec1 != 1) interleave(A, e1, scratch, tid);
        __syncthreads();
        // repack e1 bytes into bps-bit symbols, MSB first, zero padded
        const unsigned int bps = job.bps, nsym = (8 * e1 + bps - 1) / bps;
        uint8_t * pm = p.payload_mod + (size_t)job.slot * p.mod_stride;
        for (unsigned int s = tid; s < nsym; s += blockDim.x) {
            unsigned int v = 0;
            for (unsigned int b = 0; b < bps; b++) {
                unsigned int bit = s * bps + b;
                unsigned int bv = (bit < 8 * e1) ? (A[bit >> 3] >> (7 - (bit & 7))) & 1u : 0u;
                v = (v << 1) | bv;
            }
            pm[s] = (uint8_t)v;
        }
        __syncthreads();
    }
}

cudaError_t packet_encode_launch(const EncodeParams & p, cudaStream_t st)
{
    if (p.nframes == 0) return cudaSuccess;
    packet_encode_kernel<<<p.nframes < 1024 ? p.nframes : 1024, PK_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

// ================================================================== uncoded frames: one WARP per frame
// fec0 = fec1 = none (BASELINE configs 1, 4, 5): the packet is the payload followed by its CRC.
// Each warp copies its frame through shared memory and checks the CRC with a 256-entry table:
// lane 0 takes the first n - 31*per bytes, lanes 1..31 take `per` bytes each, and the partial
// registers are merged in a 5-level tree with x^(8*per*2^l) mod P products.



__global__ void __launch_bounds__(PKF_WARPS * 32) packet_plain_kernel(const PacketParams p)
{
    __shared__ uint32_t table[256];
    __shared__ __align__(16) uint8_t stage[PKF_WARPS][PKF_MAX];
    const unsigned int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        uint32_t c = tid;
#pragma unroll
        for (int j = 0; j < 8; j++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
        table[tid] = c;                          // blockDim.x == 256
    }
    // merge constants x^(8 per 2^l) mod P of the launch's first frame length (frames of a launch nearly always share
    // their length; the GF(2) products that build them cost more than the CRC of a whole frame)
    __shared__ uint32_t xps[5];
    __shared__ unsigned int xps_per;
    const unsigned int nrec = p.range[1].nrec;
    if (wid == 0) {
        unsigned int per0 = 0;
        const unsigned int r0 = p.range[0].nrec;
        if (r0 < nrec) per0 = p.recs[r0].payload_len / 32;
        // ... and kept in device memory from one launch to the next (every CTA that misses computes the same words)
        const bool hit = p.crc_cache && __ldcg(p.crc_cache) == per0 + 1u;
        if (hit) {
            if (lane < 5) xps[lane] = __ldcg(p.crc_cache + 1 + lane);
        } else {
            uint32_t xp = crc_x8n(per0);
#pragma unroll
            for (int l = 0; l < 5; l++) {
                if (lane == 0) { xps[l] = xp; if (p.crc_cache) p.crc_cache[1 + l] = xp; }
                xp = crc_multmodp(xp, xp);
            }
            if (lane == 0 && p.crc_cache) { __threadfence(); p.crc_cache[0] = per0 + 1u; }
        }
        if (lane == 0) xps_per = per0;
    }
    __syncthreads();
    for (unsigned int ri = p.range[0].nrec + blockIdx.x * PKF_WARPS + wid; ri < nrec; ri += gridDim.x * PKF_WARPS) {
        FrameRec * rec = p.recs + ri;
        if (!rec->header_valid || rec->fec0 != 1 || rec->fec1 != 1) continue;
        const unsigned int plen = rec->payload_len, crc_len = (rec->check == 6) ? 4u : 0u, n0 = plen + crc_len;
        if (n0 > PKF_MAX) continue;              // left to the general kernel
        if (p.aux[ri].sym_bps == 0xffffffffu) continue;                  // payload was not stored
        const unsigned long long off = p.aux[ri].sym_off;
        const uint32_t * src = (const uint32_t *)(p.arena + off);       // offsets are 16-byte aligned
        uint32_t * dst = (uint32_t *)(p.decoded + rec->payload_offset);
        uint32_t * st = (uint32_t *)stage[wid];
        const unsigned int sym_bps = p.aux[ri].sym_bps;
        if (sym_bps) pack_symbols(p.arena + off, (8 * n0 + sym_bps - 1) / sym_bps, sym_bps, st, dst, n0, lane, 32);
        else for (unsigned int i = lane; i < (n0 + 3) / 4; i += 32) { uint32_t v = src[i]; st[i] = v; dst[i] = v; }
        __syncwarp();
        int valid = 1;
        if (crc_len) {
            const uint8_t * m = stage[wid];
            const unsigned int per = plen / 32, first = plen - 31 * per;
            const unsigned int lo = lane ? first + (lane - 1) * per : 0, hi = lane ? lo + per : first;
            uint32_t key = lane ? 0u : ~0u;
            for (unsigned int i = lo; i < hi; i++) key = (key >> 8) ^ table[(key ^ m[i]) & 0xffu];
            // tree merge; the right operand of every merge spans a multiple of `per` bytes
            if (per == xps_per) {
#pragma unroll
                for (int l = 0; l < 5; l++) {
                    const int o = 1 << l;
                    uint32_t right = __shfl_down_sync(0xffffffffu, key, o);
                    if ((lane & (2 * o - 1)) == 0) key = crc_multmodp(xps[l], key) ^ right;
                }
            } else {
                uint32_t xp = crc_x8n(per);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t right = __shfl_down_sync(0xffffffffu, key, o);
                    if ((lane & (2 * o - 1)) == 0) key = crc_multmodp(xp, key) ^ right;
                    xp = crc_multmodp(xp, xp);
                }
            }
            if (lane == 0) {
                uint32_t want = ((uint32_t)m[plen] << 24) | ((uint32_t)m[plen + 1] << 16) | ((uint32_t)m[plen + 2] << 8) | m[plen + 3];
                valid = (~key == want);
            }
        }
        if (lane == 0) rec->payload_valid = valid;
        __syncwarp();
    }
}

__global__ void record_mark_kernel(const unsigned int * counters, RangeMark * mark_out, int used_at, RangeMark * host_out)
{
    RangeMark m;
    m.nrec = counters[0]; m.pad = counters[1];           // pad: overflow flag so far
    m.arena_used = *(const unsigned long long *)(counters + used_at);   // bytes of decoded payload so far
    *mark_out = m;
    if (host_out) { *host_out = m; __threadfence_system(); }
}
cudaError_t record_mark_launch(const unsigned int * counters, RangeMark * mark_out, cudaStream_t st, int used_at, RangeMark * host_out)
{
    record_mark_kernel<<<1, 1, 0, st>>>(counters, mark_out, used_at, host_out);
    return cudaGetLastError();
}
__global__ void batch_reset_kernel(unsigned int * counters, RangeMark * mark0, int keep_ring)
{
    const unsigned int i = threadIdx.x;
    if (i < 8 && !(keep_ring && (i == 2 || i == 3))) counters[i] = 0u;
    if (i == 8) { RangeMark z; z.nrec = 0; z.pad = 0; z.arena_used = 0; *mark0 = z; }
}
cudaError_t batch_reset_launch(unsigned int * counters, RangeMark * mark0, int keep_ring, cudaStream_t st)
{
    batch_reset_kernel<<<1, 32, 0, st>>>(counters, mark0, keep_ring);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ records in callback order, on the device
__global__ void rec_keys_kernel(const FrameRec * recs, unsigned int n, unsigned long long * keys, unsigned int * idx)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (recs[i].complete_index << 16) | (unsigned long long)(recs[i].channel & 0xffffu);
    idx[i] = i;
}
__global__ void rec_permute_kernel(const FrameRec * recs, const unsigned int * idx, unsigned int n, FrameRec * dst)
{
    // 22 words per record, one thread per word
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int r = e / (sizeof(FrameRec) / 4), wd = e % (sizeof(FrameRec) / 4);
    if (r >= n) return;
    ((uint32_t *)(dst + r))[wd] = ((const uint32_t *)(recs + idx[r]))[wd];
}
cudaError_t pack_sorted_launch(const FrameRec * recs, unsigned int n, FrameRec * dst, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    // scratch of the calling thread's device, grown on demand and never freed on the data path (no allocation, and
    // therefore no device-wide synchronisation, once it has its size)
    struct Scratch { unsigned long long * keys[2] = {nullptr, nullptr}; unsigned int * idx[2] = {nullptr, nullptr}; void * tmp = nullptr;
                     size_t tmp_bytes = 0; unsigned int cap = 0; int dev = -1; };
    static thread_local Scratch sc;
    int dev = 0;
    cudaGetDevice(&dev);
    if (n > sc.cap || dev != sc.dev) {
        for (int i = 0; i < 2; i++) { if (sc.keys[i]) cudaFree(sc.keys[i]); if (sc.idx[i]) cudaFree(sc.idx[i]); sc.keys[i] = nullptr; sc.idx[i] = nullptr; }
        if (sc.tmp) cudaFree(sc.tmp);
        sc.tmp = nullptr; sc.cap = 0; sc.dev = dev;
        const unsigned int cap = n + n / 4 + 1024;
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaMalloc(&sc.keys[i], sizeof(unsigned long long) * cap);
            if (e == cudaSuccess) e = cudaMalloc(&sc.idx[i], sizeof(unsigned int) * cap);
        }
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, sc.tmp_bytes, sc.keys[0], sc.keys[1], sc.idx[0], sc.idx[1], (int)cap, 0, 64, st);
        if (e == cudaSuccess) e = cudaMalloc(&sc.tmp, sc.tmp_bytes);
        if (e != cudaSuccess) return e;
        sc.cap = cap;
    }
    rec_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(recs, n, sc.keys[0], sc.idx[0]);
    size_t tb = sc.tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(sc.tmp, tb, sc.keys[0], sc.keys[1], sc.idx[0], sc.idx[1], (int)n, 0, 64, st);
    if (e != cudaSuccess) return e;
    const unsigned int words = n * (unsigned int)(sizeof(FrameRec) / 4);
    rec_permute_kernel<<<(words + 255) / 256, 256, 0, st>>>(recs, sc.idx[1], n, dst);
    return cudaGetLastError();
}

struct VitWorkspace { uint2 * ws = nullptr; int * locks = nullptr; size_t stride = 0; unsigned int slots = 0; };
static VitWorkspace g_vit[16];
static std::mutex g_vit_mutex;

// Viterbi decision workspace for frames whose trellis does not fit a CTA's own region (> 2 KB conv-coded payloads):
// 8 bytes per trellis step per slot, sized for the largest packet (also covers v27 as outer code over an inner code),
// 16 slots shared by the handles of a device (a CTA claims one for the duration of a frame).  Allocated when the first
// handle of the device is created -- not on the data path.
cudaError_t packet_decode_prepare()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    std::lock_guard<std::mutex> guard(g_vit_mutex);
    VitWorkspace & g = g_vit[dev];
    if (g.ws) return cudaSuccess;
    const size_t steps = 8ull * (65535 + 4 + 2) * 2 + 64;
    const unsigned int slots = 16;
    cudaError_t e = cudaMalloc(&g.ws, (size_t)slots * steps * sizeof(uint2));
    if (e != cudaSuccess) { g.ws = nullptr; return e; }
    e = cudaMalloc(&g.locks, slots * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(g.locks, 0, slots * sizeof(int));
    if (e != cudaSuccess) { cudaFree(g.ws); g.ws = nullptr; return e; }
    g.stride = steps; g.slots = slots;
    return cudaSuccess;
}

cudaError_t packet_decode_launch(const PacketParams & p, int grid, cudaStream_t st)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    VitWorkspace w;
    {
        cudaError_t pe = packet_decode_prepare();        // (a no-op after the handle's creation)
        if (pe != cudaSuccess) return pe;
        std::lock_guard<std::mutex> guard(g_vit_mutex);
        VitWorkspace & g = g_vit[dev];
        w = g;
    }
    packet_plain_kernel<<<grid, PKF_WARPS * 32, 0, st>>>(p);
    // the general kernel in its two shapes (one of them returns at once, see above): a region of the workspace per CTA
    const int ggrid = p.vit_local ? (int)p.vit_local_ctas : grid;
    packet_decode_kernel<128><<<p.vit_local ? (int)p.vit_grid128 : grid, 128, 0, st>>>(p, w.ws, w.stride, w.locks, w.slots);
    packet_decode_kernel<32><<<ggrid, 32, 0, st>>>(p, w.ws, w.stride, w.locks, w.slots);
    return cudaGetLastError();
}

} // namespace b2
