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
            unsigned int g = b_begin + lo;
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
            roww[g % RR] = nco_cexp_pi(p.theta0 + (b_begin + g) * K * p.dtheta);
        mbar_wait(&bar[t & 1], (t >> 1) & 1);
        __syncthreads();

        // NCO mix-down in place: x * conj(roww * colw)
        {
            const unsigned int total = (new_hi - new_lo) * K;
            for (unsigned int e = tid; e < total; e += AN_THREADS) {
                unsigned int g = new_lo + (e >> p.lgK), r = e & (K - 1);
                unsigned int slot = g % RR;
                cf w = cmul(roww[slot], colw[r]);
                cf * px = ring + (size_t)slot * K + r;
                *px = mix_down(*px, w);
            }
        }
        __syncthreads();

        // polyphase FIR: thread owns column r for JB consecutive blocks
        for (unsigned int it = tid; it < groups * K; it += AN_THREADS) {
            const unsigned int jb = it >> p.lgK, r = it & (K - 1);
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
            cf * out = p.out + p.out_col0 + b_begin + tb0;
            const unsigned int total = N * nb;
            for (unsigned int e = tid; e < total; e += AN_THREADS) {
                unsigned int c = e / nb, b = e - c * nb;
                out[(size_t)c * p.out_stride + b] = xbuf[(size_t)b * ldx + phys<1>(c)];
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
    if (p.TB % 8 == 0) analyzer_kernel<8><<<grid, AN_THREADS, smem_bytes, st>>>(p);
    else analyzer_kernel<4><<<grid, AN_THREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

} // namespace b2
