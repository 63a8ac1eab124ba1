// channelizer.cu -- polyphase filterbank channelizers (firpfbch_crcf, K = 2N channels) fused with
// the spectrum-centring NCO.
//
// analyzer_kernel replaces, per block of K wideband samples, the reference's
//     nco_crcf_mix_down + nco_crcf_step            lib/multichannelrx.cc:163-164   (x K)
//     firpfbch_crcf_analyzer_execute(x -> X)       lib/multichannelrx.cc:188
// and keeps X[0..N-1] only, as lib/multichannelrx.cc:193-194 does, written channel-major
// ([channel][time]) so the per-channel synchroniser reads contiguous memory.
//
// Math (liquid firpfbch.c, analyzer):  with xm[n] = x[n] * e^{-j theta_n},
//     V_b[r] = sum_{n=0}^{P-1} h[(K-1-r) + n*K] * xm[(b-n)*K + r],    r < K
//     y_b    = FFT_K(V_b)                                             (forward, unnormalised)
//
// Mapping to the SM: a persistent CTA owns a contiguous range of blocks.  Rows of K samples are
// pulled HBM -> shared memory by the TMA unit (1-D cp.async.bulk, mbarrier completion), two
// tiles in flight; the NCO is applied once per sample on arrival (row phasor x column phasor,
// uint32 phase so theta_n = theta_0 + n*dtheta exactly); each thread slides a P-deep register
// window down one column for JB consecutive blocks (each staged sample is read once per JB
// outputs); the K-point FFTs run in place in shared memory; the N kept channels leave through a
// transposed read so each channel's TB outputs are one contiguous run in HBM.
#include <cstdlib>
#include "kernels.h"

namespace b2 {

constexpr int AN_THREADS = 256;
constexpr int AN_P = 14;            // taps per branch of the receive bank (m = 7, lib/multichannelrx.cc:89)

struct AnSmem {
    unsigned int RR;                // ring rows
    unsigned int ldx;               // xbuf row stride (elements)
    size_t off_colw, off_roww, off_ring, off_xbuf, total;
};

__host__ __device__ static inline AnSmem an_layout(unsigned int K, unsigned int P, unsigned int TB)
{
    AnSmem s;
    s.RR = P - 1 + 2 * TB;
    s.ldx = K + (K >> 5) + 1;
    size_t o = 16;                                  // two mbarriers
    s.off_colw = o; o += (size_t)K * sizeof(cf);
    s.off_roww = o; o += (size_t)s.RR * sizeof(cf);
    o = (o + 127) & ~(size_t)127;
    s.off_ring = o; o += (size_t)s.RR * K * sizeof(cf);
    s.off_xbuf = o; o += (size_t)TB * s.ldx * sizeof(cf);
    s.total = o;
    return s;
}

size_t analyzer_smem_bytes(const AnalyzerParams & p) { return an_layout(p.K, p.P, p.TB).total; }

template <int JB>
__global__ void __launch_bounds__(AN_THREADS, 1) analyzer_kernel(const AnalyzerParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned int tid = threadIdx.x;
    const unsigned int K = p.K, N = p.N, TB = p.TB;
    const AnSmem L = an_layout(K, AN_P, TB);
    uint64_t * bar = (uint64_t *)smem;
    cf * colw = (cf *)(smem + L.off_colw);
    cf * roww = (cf *)(smem + L.off_roww);
    cf * ring = (cf *)(smem + L.off_ring);
    cf * xbuf = (cf *)(smem + L.off_xbuf);
    const unsigned int RR = L.RR, ldx = L.ldx;

    // contiguous block range of this CTA, whole tiles
    unsigned int tiles_total = (p.nblocks + TB - 1) / TB;
    unsigned int tiles_per = (tiles_total + gridDim.x - 1) / gridDim.x;
    unsigned int b_begin = blockIdx.x * tiles_per * TB;
    if (b_begin >= p.nblocks) return;
    unsigned int b_end = min(p.nblocks, b_begin + tiles_per * TB);
    unsigned int ntiles = (b_end - b_begin + TB - 1) / TB;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (unsigned int r = tid; r < K; r += AN_THREADS) colw[r] = nco_cexp_pi(r * p.dtheta);
    __syncthreads();

    // rows are addressed relative to b_begin: rel row g <-> logical row b_begin + g, slot g % RR
    auto issue = [&](unsigned int lo, unsigned int hi, uint64_t * mb) {
        // one elected thread: expect the byte count, then one bulk copy per contiguous run
        mbar_expect_tx(mb, (hi - lo) * K * (unsigned int)sizeof(cf));
        while (lo < hi) {
            unsigned int slot = lo % RR;
            unsigned int run = min(hi - lo, RR - slot);
            unsigned int g = p.block0 + b_begin + lo;
            const cf * src;
            if (g < p.rows0) { run = min(run, p.rows0 - g); src = p.seg0 + (size_t)g * K; }
            else src = p.seg1 + (size_t)(g - p.rows0) * K;
            // a single bulk copy moves at most 2^20-16 bytes
            unsigned int maxrows = max(1u, (1u << 19) / (K * (unsigned int)sizeof(cf)));
            run = min(run, maxrows);
            bulk_g2s(ring + (size_t)slot * K, src, run * K * (unsigned int)sizeof(cf), mb);
            lo += run;
        }
    };

    unsigned int nb0 = min(TB, b_end - b_begin);
    if (tid == 0) issue(0, AN_P - 1 + nb0, &bar[0]);

    const unsigned int groups = TB / JB;
    for (unsigned int t = 0; t < ntiles; t++) {
        const unsigned int tb0 = t * TB;                               // first block of tile (relative)
        const unsigned int nb = min(TB, b_end - b_begin - tb0);
        const unsigned int new_lo = (t == 0) ? 0 : tb0 + AN_P - 1;     // rows that arrive with this tile
        const unsigned int new_hi = tb0 + AN_P - 1 + nb;

        // prefetch the next tile's rows while this one is processed
        if (t + 1 < ntiles && tid == 0) {
            unsigned int nb1 = min(TB, b_end - b_begin - (tb0 + TB));
            fence_proxy_async();
            issue(tb0 + TB + AN_P - 1, tb0 + TB + AN_P - 1 + nb1, &bar[(t + 1) & 1]);
        }

        // row phasors of the arriving rows
        for (unsigned int g = new_lo + tid; g < new_hi; g += AN_THREADS)
            roww[g % RR] = nco_cexp_pi(p.theta0 + (p.block0 + b_begin + g) * K * p.dtheta);
        mbar_wait(&bar[t & 1], (t >> 1) & 1);
        __syncthreads();

        // NCO mix-down in place: x * conj(roww * colw)
        {
            const unsigned int total = (new_hi - new_lo) * K;
            for (unsigned int e = tid; e < total; e += AN_THREADS) {
                unsigned int g = new_lo + e / K, r = e % K;
                unsigned int slot = g % RR;
                cf w = cmul(roww[slot], colw[r]);
                cf * px = ring + (size_t)slot * K + r;
                *px = mix_down(*px, w);
            }
        }
        __syncthreads();

        // polyphase FIR: thread owns column r for JB consecutive blocks
        for (unsigned int it = tid; it < groups * K; it += AN_THREADS) {
            const unsigned int jb = it / K, r = it % K;
            if (jb * JB >= nb) continue;
            float h[AN_P];
#pragma unroll
            for (int n = 0; n < AN_P; n++) h[n] = __ldg(p.taps + (size_t)n * K + (K - 1 - r));
            cf v[JB + AN_P - 1];
            const unsigned int row0 = tb0 + jb * JB;
#pragma unroll
            for (int q = 0; q < JB + AN_P - 1; q++) v[q] = ring[(size_t)((row0 + q) % RR) * K + r];
            const unsigned int pr = phys<1>(p.fft.perm[r]);
#pragma unroll
            for (int bl = 0; bl < JB; bl++) {
                float ar = 0.f, ai = 0.f;
#pragma unroll
                for (int n = AN_P - 1; n >= 0; n--) {          // oldest sample first, as dotprod_crcf
                    ar = fmaf(h[n], v[bl + AN_P - 1 - n].x, ar);
                    ai = fmaf(h[n], v[bl + AN_P - 1 - n].y, ai);
                }
                xbuf[(size_t)(jb * JB + bl) * ldx + pr] = make_float2(ar, ai);
            }
        }
        __syncthreads();

        // K-point forward FFT of each block, in place
        fft_inplace<-1, 1>(xbuf, ldx, nb, p.fft, tid, AN_THREADS);

        // channels 0..N-1 -> out[c][col0 + b], runs of nb contiguous samples per channel
        {
            const size_t ocol = p.out_col0 + b_begin + tb0;
            const unsigned int total = N * nb;
            for (unsigned int e = tid; e < total; e += AN_THREADS) {
                unsigned int c = e / nb, b = e - c * nb;
                (p.n_peer ? p.out_peer[c / p.chan_per_peer] + (size_t)(c % p.chan_per_peer) * p.out_stride
                          : p.out + (size_t)c * p.out_stride)[ocol + b] = xbuf[(size_t)b * ldx + phys<1>(c)];
            }
        }
        __syncthreads();
    }
}

cudaError_t analyzer_configure(size_t smem_bytes)
{
    cudaError_t e = cudaFuncSetAttribute(analyzer_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(analyzer_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

cudaError_t analyzer_launch(const AnalyzerParams & p, int grid, size_t smem_bytes, cudaStream_t st)
{
    if (p.nblocks == 0) return cudaSuccess;
    static const bool force_generic = getenv("B2_ANALYZER_GENERIC") != nullptr;
    if (!force_generic && analyzer8_supported(p)) return analyzer8_launch(p, st);
    if (p.TB % 8 == 0) analyzer_kernel<8><<<grid, AN_THREADS, smem_bytes, st>>>(p);
    else analyzer_kernel<4><<<grid, AN_THREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

} // namespace b2

// ================================================================== synthesis channelizer
// synth_kernel replaces, per call of multichanneltx::GenerateSamples (lib/multichanneltx.cc:192-227),
//     X[i] = fgbuffer[i][fgbuffer_index]  (i < N; X[N..2N) stay 0)    :205-210
//     firpfbch_crcf_synthesizer_execute(channelizer, X, buffer)      :213
//     nco_crcf_mix_up + nco_crcf_step per output sample              :219-222
// Math (liquid firpfbch.c, synthesizer):  V_t = IFFT_K(X_t) (backward, unnormalised),
//     y_t[i] = sum_{n=0}^{P-1} h[i + n*K] * V_{t-n}[i],   out[t*K + i] = y_t[i] * e^{+j theta}
// A CTA owns a contiguous range of blocks; it keeps the last P-1+TB IFFT outputs in a shared
// memory ring (the P-1 rows before its range are recomputed from the input, or come from the
// persistent history for the first blocks of a call), slides a P-deep register window down each
// column like the analyzer, and writes whole rows of K wideband samples.
namespace b2 {

constexpr int SY_THREADS = 256;
constexpr int SY_P = 26;            // taps per branch of the transmit bank (m = 13, lib/multichanneltx.cc:85)

struct SynSmem { unsigned int RR; size_t off_colw, off_roww, off_ring, total; };
__host__ __device__ static inline SynSmem syn_layout(unsigned int K, unsigned int TB)
{
    SynSmem s;
    s.RR = SY_P - 1 + TB;
    size_t o = 0;
    s.off_colw = o; o += (size_t)K * sizeof(cf);
    s.off_roww = o; o += (size_t)(TB + 1) * sizeof(cf);
    o = (o + 15) & ~(size_t)15;
    s.off_ring = o; o += (size_t)s.RR * K * sizeof(cf);
    s.total = o;
    return s;
}
size_t synth_smem_bytes(const SynthParams & p) { return syn_layout(p.K, p.TB).total; }

template <int JB>
__global__ void __launch_bounds__(SY_THREADS, 1) synth_kernel(const SynthParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned int tid = threadIdx.x;
    const unsigned int K = p.K, N = p.N, TB = p.TB;
    const SynSmem L = syn_layout(K, TB);
    cf * colw = (cf *)(smem + L.off_colw);
    cf * roww = (cf *)(smem + L.off_roww);
    cf * ring = (cf *)(smem + L.off_ring);
    const unsigned int RR = L.RR;

    unsigned int tiles_total = (p.nblocks + TB - 1) / TB;
    unsigned int tiles_per = (tiles_total + gridDim.x - 1) / gridDim.x;
    unsigned int b_begin = blockIdx.x * tiles_per * TB;
    if (b_begin >= p.nblocks) return;
    unsigned int b_end = min(p.nblocks, b_begin + tiles_per * TB);
    unsigned int ntiles = (b_end - b_begin + TB - 1) / TB;

    for (unsigned int r = tid; r < K; r += SY_THREADS) colw[r] = nco_cexp_pi(r * p.dtheta);

    // V row of block t (t may be negative: history) -> ring slot (t - (b_begin - (P-1))) % RR.
    // rows with t < 0 come from vhist[P-1+t]; others are IFFTs of the gathered channel samples.
    auto fill_rows = [&](int t_lo, int t_hi) {
        const int org = (int)b_begin - (SY_P - 1);
        // history rows (t < 0) are copied, the others gathered in FFT-permuted order; the gather
        // runs with time fastest so each channel contributes one contiguous run
        const int h_hi = min(t_hi, 0);
        for (int t = t_lo; t < h_hi; t++) {
            cf * row = ring + (size_t)((unsigned int)(t - org) % RR) * K;
            const cf * src = p.vhist + (size_t)(SY_P - 1 + t) * K;
            for (unsigned int i = tid; i < K; i += SY_THREADS) row[i] = src[i];
        }
        const int g_lo = max(t_lo, 0);
        if (g_lo < t_hi) {
            const unsigned int nr = (unsigned int)(t_hi - g_lo);
            for (unsigned int e = tid; e < nr * (K - N); e += SY_THREADS) {
                unsigned int tq = e / (K - N), i = N + (e - tq * (K - N));
                ring[(size_t)((unsigned int)(g_lo + (int)tq - org) % RR) * K + p.fft.perm[i]] = make_float2(0.f, 0.f);
            }
            for (unsigned int e = tid; e < nr * N; e += SY_THREADS) {
                unsigned int i = e / nr, tq = e - i * nr;
                cf v = p.in[(size_t)i * p.in_stride + p.in_off + (unsigned int)(g_lo + (int)tq)];
                ring[(size_t)((unsigned int)(g_lo + (int)tq - org) % RR) * K + p.fft.perm[i]] = v;
            }
        }
        __syncthreads();
        for (int t = g_lo; t < t_hi;) {              // IFFT the rows, one contiguous run of ring slots at a time
            unsigned int slot = (unsigned int)(t - org) % RR;
            unsigned int run = min((unsigned int)(t_hi - t), RR - slot);
            fft_inplace<+1, 0>(ring + (size_t)slot * K, K, run, p.fft, tid, SY_THREADS);
            t += (int)run;
        }
    };

    // halo: the P-1 rows before the range
    fill_rows((int)b_begin - (SY_P - 1), (int)b_begin);

    const unsigned int groups = TB / JB;
    for (unsigned int tt = 0; tt < ntiles; tt++) {
        const unsigned int tb0 = b_begin + tt * TB;
        const unsigned int nb = min(TB, b_end - tb0);
        fill_rows((int)tb0, (int)(tb0 + nb));
        for (unsigned int g = tid; g < nb; g += SY_THREADS) roww[g] = nco_cexp_pi(p.theta0 + (tb0 + g) * K * p.dtheta);
        __syncthreads();
        for (unsigned int it = tid; it < groups * K; it += SY_THREADS) {
            const unsigned int jb = it / K, r = it % K;
            if (jb * JB >= nb) continue;
            float h[SY_P];
#pragma unroll
            for (int n = 0; n < SY_P; n++) h[n] = __ldg(p.taps + (size_t)n * K + r);
            cf v[JB + SY_P - 1];
            // v[q] = V_{tb0 + jb*JB - (P-1) + q}[r]
            const unsigned int rel0 = tt * TB + jb * JB;                   // ring-relative row of v[0]
#pragma unroll
            for (int q = 0; q < JB + SY_P - 1; q++) v[q] = ring[(size_t)((rel0 + q) % RR) * K + r];
            const cf cw = colw[r];
#pragma unroll
            for (int bl = 0; bl < JB; bl++) {
                if (jb * JB + bl >= nb) break;
                float ar = 0.f, ai = 0.f;
#pragma unroll
                for (int n = SY_P - 1; n >= 0; n--) {                      // oldest first, as dotprod_crcf
                    ar = fmaf(h[n], v[bl + SY_P - 1 - n].x, ar);
                    ai = fmaf(h[n], v[bl + SY_P - 1 - n].y, ai);
                }
                cf w = cmul(roww[jb * JB + bl], cw);
                p.out[(size_t)(tb0 + jb * JB + bl) * K + r] = mix_up(make_float2(ar, ai), w);
            }
        }
        __syncthreads();
    }
    // the CTA that owns the last blocks leaves the new history (last P-1 IFFT rows of the call)
    if (b_end == p.nblocks) {
        for (int q = 0; q < SY_P - 1; q++) {
            int t = (int)p.nblocks - (SY_P - 1) + q;
            cf * dst = p.vhist_out + (size_t)q * K;
            if (t < (int)b_begin - (SY_P - 1)) {
                // call shorter than the history: older rows shift down from the previous history
                const cf * src = p.vhist + (size_t)(SY_P - 1 + t) * K;
                for (unsigned int i = tid; i < K; i += SY_THREADS) dst[i] = src[i];
            } else {
                unsigned int slot = (unsigned int)(t - ((int)b_begin - (SY_P - 1))) % RR;
                const cf * src = ring + (size_t)slot * K;
                for (unsigned int i = tid; i < K; i += SY_THREADS) dst[i] = src[i];
            }
        }
    }
}

cudaError_t synth_configure(size_t smem_bytes)
{
    cudaError_t e = cudaFuncSetAttribute(synth_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(synth_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}

cudaError_t synth_launch(const SynthParams & p, int grid, size_t smem_bytes, cudaStream_t st)
{
    if (p.nblocks == 0) return cudaSuccess;
    if (p.TB % 8 == 0) synth_kernel<8><<<grid, SY_THREADS, smem_bytes, st>>>(p);
    else synth_kernel<2><<<grid, SY_THREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

} // namespace b2
