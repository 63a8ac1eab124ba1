// fft8.cuh -- register-resident Stockham FFT: a transform of N points is carried by N/8
// threads, every thread holding 8 points (v[s] <-> element j + s*N/8 for thread j) before the
// first pass and after the last one.  Passes are radix 8 as far as N allows, the last one
// radix 2/4 where log2 N is not a multiple of 3; between passes the points are exchanged
// through two ping-pong shared-memory buffers (XOR-swizzled, p ^ ((p >> 3) & 15), which makes the
// scattered stores of every pass and the strided loads of the next one bank-conflict free for
// every N from 64 to 4096 -- modelled exhaustively).  The output is in NATURAL
// order, so the transform needs neither a bit-reversal pass nor a permutation table.
//
// This is the K-point transform inside firpfbch_crcf_analyzer_execute
// (lib/multichannelrx.cc:188) and the M-point transform inside ofdmframesync_execute
// (lib/multichannelrx.cc:194, lib/ofdmtxrx.cc:625); unnormalised like liquid's.
#pragma once
#include "dsp.cuh"

namespace b2 {

__host__ __device__ constexpr unsigned int f8_pad(unsigned int i) { return i ^ ((i >> 3) & 15u); }
__host__ __device__ constexpr unsigned int f8_buf_elems(unsigned int n) { return n; }

template <int DIR> __device__ __forceinline__ void f8_dft(cf & a0, cf & a1) { dft2<DIR>(a0, a1); }
template <int DIR> __device__ __forceinline__ void f8_dft(cf & a0, cf & a1, cf & a2, cf & a3)
{
    cf v[4] = {a0, a1, a2, a3};
    dft4<DIR>(v);
    a0 = v[0]; a1 = v[1]; a2 = v[2]; a3 = v[3];
}

// twiddle + butterflies of one pass; Ns = product of the radices of the earlier passes.
// The twiddles of a thread do not change from one transform to the next, so a persistent
// kernel may keep them in registers: twr != nullptr -> twr[u*(R-1) + q-1] (see f8_tw_init).
// A third source, for kernels short of registers: twt = per-pass table in the order the lanes read
// it, twt[(q-1)*Ns + k] (conflict-free; built once per CTA by f8_twt_build).
template <unsigned int N, unsigned int Ns, unsigned int R, int DIR>
__device__ __forceinline__ void f8_pass(cf (&v)[8], unsigned int j, const cf * __restrict__ tw, const cf * twr = nullptr,
                                        const cf * __restrict__ twt = nullptr)
{
    constexpr unsigned int T = N / 8, U = 8 / R;
#pragma unroll
    for (unsigned int u = 0; u < U; u++) {
        if (Ns > 1) {
            const unsigned int k = (j + u * T) & (Ns - 1);
#pragma unroll
            for (unsigned int q = 1; q < R; q++) {
                cf w = twr ? twr[u * (R - 1) + q - 1] : (twt ? twt[(q - 1) * Ns + k] : tw[k * q * (N / (Ns * R))]);
                if (DIR > 0) w.y = -w.y;
                v[u + q * U] = cmul(v[u + q * U], w);
            }
        }
        if (R == 8) dft8<DIR>(v);
        else if (R == 4) f8_dft<DIR>(v[u], v[u + 2], v[u + 4], v[u + 6]);
        else f8_dft<DIR>(v[u], v[u + 4]);
    }
}

// scatter the outputs of a pass (not the last) / gather the inputs of the next one
template <unsigned int N, unsigned int Ns, unsigned int R>
__device__ __forceinline__ void f8_store(const cf (&v)[8], unsigned int j, cf * __restrict__ buf)
{
    constexpr unsigned int T = N / 8, U = 8 / R;
#pragma unroll
    for (unsigned int u = 0; u < U; u++) {
        const unsigned int b = j + u * T, k = b & (Ns - 1);
        const unsigned int b0 = (b - k) * R + k;
#pragma unroll
        for (unsigned int q = 0; q < R; q++) buf[f8_pad(b0 + q * Ns)] = v[u + q * U];
    }
}
template <unsigned int N>
__device__ __forceinline__ void f8_load(cf (&v)[8], unsigned int j, const cf * __restrict__ buf)
{
#pragma unroll
    for (unsigned int s = 0; s < 8; s++) v[s] = buf[f8_pad(j + s * (N / 8))];
}

__host__ __device__ constexpr unsigned int f8_radix(unsigned int n, unsigned int ns) { return (n / ns >= 8) ? 8u : n / ns; }

// number of per-thread twiddles of the passes from Ns on (7 per radix-8 pass, 6 / 4 for a final
// radix-4 / radix-2 pass), and their one-off computation from the table
__host__ __device__ constexpr unsigned int f8_tw_count(unsigned int n, unsigned int ns)
{
    return ns >= n ? 0u : (ns > 1 ? (8 / f8_radix(n, ns)) * (f8_radix(n, ns) - 1) : 0u) + f8_tw_count(n, ns * f8_radix(n, ns));
}
template <unsigned int N, unsigned int Ns>
__device__ __forceinline__ void f8_tw_init(cf * twr, unsigned int j, const cf * __restrict__ tw)
{
    constexpr unsigned int R = f8_radix(N, Ns), T = N / 8, U = 8 / R;
    if constexpr (Ns > 1) {
#pragma unroll
        for (unsigned int u = 0; u < U; u++) {
            const unsigned int k = (j + u * T) & (Ns - 1);
#pragma unroll
            for (unsigned int q = 1; q < R; q++) twr[u * (R - 1) + q - 1] = tw[k * q * (N / (Ns * R))];
        }
    }
    if constexpr (Ns * R < N) f8_tw_init<N, Ns * R>(twr + (Ns > 1 ? U * (R - 1) : 0), j, tw);
}

// ordered per-pass tables (see f8_pass): total size and construction by `nthreads` threads
__host__ __device__ constexpr unsigned int f8_twt_elems(unsigned int n, unsigned int ns)
{
    return ns >= n ? 0u : (ns > 1 ? (f8_radix(n, ns) - 1) * ns : 0u) + f8_twt_elems(n, ns * f8_radix(n, ns));
}
template <unsigned int N, unsigned int Ns>
__device__ __forceinline__ void f8_twt_build(cf * twt, const cf * __restrict__ tw, unsigned int tid, unsigned int nthreads)
{
    constexpr unsigned int R = f8_radix(N, Ns);
    if constexpr (Ns > 1) {
        for (unsigned int e = tid; e < (R - 1) * Ns; e += nthreads) {
            const unsigned int q = e / Ns + 1, k = e % Ns;
            twt[e] = tw[k * q * (N / (Ns * R))];
        }
    }
    if constexpr (Ns * R < N) f8_twt_build<N, Ns * R>(twt + (Ns > 1 ? (R - 1) * Ns : 0), tw, tid, nthreads);
}

// passes from Ns on; `sync` is the barrier of the N/8 threads that carry this transform
template <unsigned int N, unsigned int Ns, int DIR, typename Sync>
__device__ __forceinline__ void f8_run(cf (&v)[8], unsigned int j, cf * __restrict__ bufA, cf * __restrict__ bufB,
                                       const cf * __restrict__ tw, Sync sync, const cf * twr = nullptr,
                                       const cf * __restrict__ twt = nullptr)
{
    constexpr unsigned int R = f8_radix(N, Ns), U = 8 / R;
    f8_pass<N, Ns, R, DIR>(v, j, tw, twr, twt);
    if constexpr (Ns * R < N) {
        f8_store<N, Ns, R>(v, j, bufA);
        sync();
        f8_load<N>(v, j, bufA);
        f8_run<N, Ns * R, DIR, Sync>(v, j, bufB, bufA, tw, sync, twr ? twr + (Ns > 1 ? U * (R - 1) : 0) : nullptr,
                                     twt ? twt + (Ns > 1 ? (R - 1) * Ns : 0) : nullptr);
    }
}

// in: v[s] = x[j + s*N/8]; out: v[s] = X[j + s*N/8].  tw[i] = e^{-2 pi i/N} (any memory space).
// bufA / bufB: f8_buf_elems(N) elements each; the caller keeps a barrier between two transforms
// that share them.
template <unsigned int N, int DIR, typename Sync>
__device__ __forceinline__ void fft8(cf (&v)[8], unsigned int j, cf * __restrict__ bufA, cf * __restrict__ bufB,
                                     const cf * __restrict__ tw, Sync sync)
{
    static_assert(N >= 64 && (N & (N - 1)) == 0, "fft8 needs a power of two >= 64");
    f8_run<N, 1, DIR, Sync>(v, j, bufA, bufB, tw, sync);
}

} // namespace b2
