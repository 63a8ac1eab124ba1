// framegen.cu -- OFDM frame generator (ofdmflexframegen + ofdmframegen) and the arbitrary
// resampler.
//
// framegen_kernel replaces the per-channel, per-symbol-period call
//     ofdmflexframegen_write(framegen[i], fgbuffer[i], M+cp)   lib/multichanneltx.cc:236
//     ofdmflexframegen_write(fg, fgbuffer, fgbuffer_len)        lib/ofdmtxrx.cc:328
// (liquid: S0a, S0b, S1, header symbols, payload symbols, tail; data subcarriers in natural
// index order, pilots +-1 from an 8-bit LFSR visited in fft-shifted order, unnormalised IFFT,
// cyclic prefix, raised-sine^2 taper overlapped with the previous symbol's postfix).
// One CTA per channel walks the symbol periods of the call (consecutive symbols are coupled
// through the taper postfix); idle channels/periods produce zeros (lib/multichanneltx.cc:239).
//
// resamp_kernel is msresamp_crcf's arbitrary polyphase stage (src/flexframe_rx.cc:179,240) with
// a Q32 fixed-point output phase: every output is an independent dot product.
#include "kernels.h"

namespace b2 {

constexpr int FG_THREADS = 128;

struct FgLayout { size_t off_X, off_tw, off_perm, off_rank, off_post, total; };
__host__ __device__ static inline FgLayout fg_layout(unsigned int M, unsigned int taper)
{
    FgLayout L;
    size_t o = 0;
    L.off_X = o;    o += (size_t)M * sizeof(cf);
    L.off_tw = o;   o += (size_t)M * sizeof(cf);
    L.off_post = o; o += (size_t)(taper + 2) * sizeof(cf);
    L.off_perm = o; o += (size_t)M * sizeof(uint16_t);
    L.off_rank = o; o += (size_t)M * sizeof(uint16_t);
    L.total = (o + 15) & ~(size_t)15;
    return L;
}
size_t framegen_smem_bytes(const FramegenParams & p) { return fg_layout(p.M, p.taper).total; }

__device__ __forceinline__ unsigned int gray_decode(unsigned int s)
{
    unsigned int m = s >> 1;
    while (m) { s ^= m; m >>= 1; }
    return s;
}
// liquid modem_modulate for BPSK / QPSK / square QAM
__device__ __forceinline__ cf modulate(unsigned int s, unsigned int scheme, unsigned int bps, float alpha)
{
    if (scheme == 39) return make_float2(s ? -1.0f : 1.0f, 0.f);
    if (scheme == 40) {
        const float h = 0.70710678118654752440f;
        return make_float2((s & 1) ? -h : h, (s & 2) ? -h : h);
    }
    unsigned int m = bps >> 1;
    unsigned int si = gray_decode(s >> m), sq = gray_decode(s & ((1u << m) - 1u));
    float vi = (float)(2 * (int)si - (int)(1u << m) + 1) * alpha;
    float vq = (float)(2 * (int)sq - (int)(1u << m) + 1) * alpha;
    return make_float2(vi, vq);
}
// deterministic stand-in for liquid's rand() padding symbols (oracle normative choice D5)
__device__ __forceinline__ unsigned int pad_symbol(unsigned int slot, unsigned int Mconst)
{
    return ((slot * 2654435761u) >> 16) % Mconst;
}

__global__ void __launch_bounds__(FG_THREADS) framegen_kernel(const FramegenParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int tid = threadIdx.x, nt = FG_THREADS;
    const unsigned int M = p.M, cp = p.cp, W = M + cp, taper = p.taper;
    const unsigned int c = blockIdx.x;
    const FgLayout L = fg_layout(M, taper);
    cf * X = (cf *)(smem + L.off_X);
    cf * tw = (cf *)(smem + L.off_tw);
    cf * post = (cf *)(smem + L.off_post);
    uint16_t * perm = (uint16_t *)(smem + L.off_perm);
    uint16_t * rank = (uint16_t *)(smem + L.off_rank);
    const GenDesc d = p.desc[c];
    cf * out = p.out + (size_t)c * p.out_stride + p.out_off;
    cf * gpost = p.postfix + (size_t)c * taper;

    for (unsigned int i = tid; i < M; i += nt) { tw[i] = p.fft.tw[i]; perm[i] = p.fft.perm[i]; rank[i] = p.sc_rank[i]; }
    for (unsigned int i = tid; i < taper; i += nt) post[i] = d.fresh ? make_float2(0.f, 0.f) : gpost[i];
    __syncthreads();
    FftDev fft = p.fft;
    fft.tw = tw; fft.perm = perm;

    const unsigned int total_syms = 3 + d.n_hdr + d.n_pay + 1;      // incl. the tail buffer
    const unsigned int hdr_pad = d.n_hdr * p.M_data - 288u;
    const unsigned int Mconst = 1u << d.bps;
    for (unsigned int per = 0; per < p.nper; per++) {
        cf * y = out + (size_t)per * W;
        const unsigned int s = d.first_symbol + per;
        if (per >= d.n_periods || s >= total_syms) {
            for (unsigned int i = tid; i < W; i += nt) y[i] = make_float2(0.f, 0.f);
            continue;
        }
        if (s == 0) {                    // S0a: no cyclic prefix logic, taper up
            for (unsigned int i = tid; i < W; i += nt) {
                cf v = p.s0[(i + M - 2 * cp) % M];
                if (i < taper) v = cscale(v, p.taper_w[i]);
                y[i] = v;
            }
            continue;
        }
        if (s == 1) {                    // S0b; its postfix is s0[0..taper)
            for (unsigned int i = tid; i < W; i += nt) y[i] = p.s0[(i + M - cp) % M];
            __syncthreads();
            for (unsigned int i = tid; i < taper; i += nt) post[i] = p.s0[i];
            __syncthreads();
            continue;
        }
        if (s == total_syms - 1) {       // tail: ramp the last postfix down, then silence
            for (unsigned int i = tid; i < W; i += nt)
                y[i] = (i < taper) ? cscale(post[i], p.taper_w[taper - 1 - i]) : make_float2(0.f, 0.f);
            continue;
        }
        // time-domain symbol x[0..M) into X (natural order)
        if (s == 2) {
            for (unsigned int i = tid; i < M; i += nt) X[i] = p.s1[i];
            __syncthreads();
        } else {
            const bool is_hdr = s < 3 + d.n_hdr;
            const unsigned int dsym = s - 3;                                     // data-symbol counter (pilot phase)
            const unsigned int base = is_hdr ? dsym * p.M_data : (s - 3 - d.n_hdr) * p.M_data;
            const unsigned int limit = is_hdr ? 288u : d.payload_mod_len;
            const uint8_t * src = is_hdr ? p.header_mod + (size_t)c * 288 : p.payload_mod + (size_t)c * p.mod_stride;
            const unsigned int scheme = is_hdr ? 39u : d.mod;
            const unsigned int bps = is_hdr ? 1u : d.bps;
            const float alpha = p.qam_alpha[bps];
            const unsigned int pad_base = is_hdr ? 0u : hdr_pad;
            const unsigned int ppos = (dsym * p.M_pilot) % 255u;
            for (unsigned int k = tid; k < M; k += nt) {
                unsigned int rk = rank[k];
                cf v = make_float2(0.f, 0.f);
                if (rk == 0xffffu) {
                } else if (rk & 0x4000u) {
                    unsigned int q = (ppos + (rk & 0x3fffu)) % 255u;
                    v = make_float2(p.pilot_seq[q] ? p.g_data : -p.g_data, 0.f);
                } else {
                    unsigned int idx = base + rk;
                    unsigned int sym = (idx < limit) ? src[idx] : pad_symbol(pad_base + (idx - limit), is_hdr ? 2u : Mconst);
                    v = cscale(modulate(sym, scheme, bps, alpha), p.g_data);
                }
                X[perm[k]] = v;
            }
            __syncthreads();
            fft_inplace<+1, 0>(X, M, 1, fft, tid, nt);
        }
        // cyclic prefix + symbol, taper overlapped with the previous symbol's postfix
        for (unsigned int i = tid; i < W; i += nt) {
            cf v = (i < cp) ? X[M - cp + i] : X[i - cp];
            if (i < taper) {
                float w0 = p.taper_w[i], w1 = p.taper_w[taper - 1 - i];
                v = make_float2(v.x * w0 + post[i].x * w1, v.y * w0 + post[i].y * w1);
            }
            y[i] = v;
        }
        __syncthreads();
        for (unsigned int i = tid; i < taper; i += nt) post[i] = X[i];
        __syncthreads();
    }
    for (unsigned int i = tid; i < taper; i += nt) gpost[i] = post[i];
}

cudaError_t framegen_launch(const FramegenParams & p, cudaStream_t st)
{
    if (p.nchan == 0 || p.nper == 0) return cudaSuccess;
    size_t smem = framegen_smem_bytes(p);
    static size_t configured_dev[64] = {0};              // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t & configured = configured_dev[(unsigned int)dev & 63u];
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(framegen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    framegen_kernel<<<p.nchan, FG_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ msresamp_crcf (arbitrary stage)
// One CTA produces RS_OUT consecutive outputs.  Their input window (RS_OUT*step/2^32 + 2m samples,
// at most 2*RS_OUT + 2m for rates >= 0.5) and the prototype, regrouped as (h[b + n*npfb], h[b + 1 + n*npfb])
// pairs, are staged in shared memory with coalesced loads; samples older than x[0] come from the
// 2m-1 sample history of the previous call.  Arithmetic per output is the oracle's: two
// accumulations over n ascending (newest sample first), then the linear blend.
#define RS_THREADS 256
#define RS_PER     4
#define RS_OUT     (RS_THREADS * RS_PER)
#define RS_MAXTAPS 16
__global__ void __launch_bounds__(RS_THREADS) resamp_kernel(const ResampParams p)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const unsigned int npfb = 1u << p.npfb_bits;
    float2 * hp = reinterpret_cast<float2 *>(rs_smem);               // [m2][npfb] pairs
    cf * xs = reinterpret_cast<cf *>(hp + p.m2 * npfb);              // input window
    const unsigned long long k0 = (unsigned long long)blockIdx.x * RS_OUT;
    const unsigned long long kn = (p.ny - k0 < RS_OUT) ? p.ny - k0 : RS_OUT;
    const long long i_lo = (long long)((p.tau0 + k0 * p.step) >> 32) - (long long)(p.m2 - 1);
    const long long i_hi = (long long)((p.tau0 + (k0 + kn - 1) * p.step) >> 32);
    const unsigned int nwin = (unsigned int)(i_hi - i_lo + 1);
    for (unsigned int j = threadIdx.x; j < p.m2 * npfb; j += RS_THREADS)
        hp[j] = make_float2(__ldg(p.h + j), __ldg(p.h + j + 1));
    for (unsigned int j = threadIdx.x; j < nwin; j += RS_THREADS) {
        long long i = i_lo + j;
        xs[j] = (i >= 0) ? __ldg(p.x + i) : p.hist_buf[(long long)p.hist + i];
    }
    __syncthreads();
    const unsigned int fb = 32 - p.npfb_bits;
    const float inv = 1.0f / (float)(1u << fb);
#pragma unroll
    for (unsigned int r = 0; r < RS_PER; r++) {
        const unsigned int kk = threadIdx.x + r * RS_THREADS;
        if (kk >= kn) break;
        const unsigned long long t = p.tau0 + (k0 + kk) * p.step;
        const unsigned int f = (unsigned int)t;
        const unsigned int bnk = f >> fb;
        const float mu = (float)(f & ((1u << fb) - 1u)) * inv;
        const cf * x = xs + ((long long)(t >> 32) - i_lo);          // x[-n] = sample i-n
        const float2 * h = hp + bnk;
        float y0r = 0.f, y0i = 0.f, y1r = 0.f, y1i = 0.f;
        for (unsigned int n = 0; n < p.m2; n++) {
            const cf s = x[-(int)n];
            const float2 c = h[n * npfb];
            y0r = fmaf(c.x, s.x, y0r); y0i = fmaf(c.x, s.y, y0i);
            y1r = fmaf(c.y, s.x, y1r); y1i = fmaf(c.y, s.y, y1i);
        }
        p.y[k0 + kk] = make_float2((1.0f - mu) * y0r + mu * y1r, (1.0f - mu) * y0i + mu * y1i);
    }
}

// history of the next call: the last `hist` samples of (previous history ++ x[0..nx))
__global__ void resamp_hist_kernel(cf * hist_buf, unsigned int hist, const cf * x, unsigned long long nx)
{
    const unsigned int j = threadIdx.x;
    cf v = make_float2(0.f, 0.f);
    if (j < hist) {
        long long i = (long long)nx - (long long)hist + j;           // index into x; negative = old history
        v = (i >= 0) ? x[i] : hist_buf[(long long)hist + i];
    }
    __syncthreads();
    if (j < hist) hist_buf[j] = v;
}

size_t resamp_smem_bytes(const ResampParams & p)
{
    // window: ceil(RS_OUT * step / 2^32) + m2 + 1 samples
    unsigned long long span = ((unsigned long long)RS_OUT * p.step >> 32) + p.m2 + 2;
    return sizeof(float2) * p.m2 * ((size_t)1 << p.npfb_bits) + sizeof(cf) * span;
}

cudaError_t resamp_launch(const ResampParams & p, cudaStream_t st)
{
    if (p.ny) {
        const size_t smem = resamp_smem_bytes(p);
        static size_t configured_dev[64] = {0};          // the attribute is per device
        int dev = 0;
        cudaGetDevice(&dev);
        size_t & configured = configured_dev[(unsigned int)dev & 63u];
        if (smem > 48 * 1024 && smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(resamp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        unsigned long long blocks = (p.ny + RS_OUT - 1) / RS_OUT;
        resamp_kernel<<<(unsigned int)blocks, RS_THREADS, smem, st>>>(p);
    }
    resamp_hist_kernel<<<1, 32, 0, st>>>(p.hist_buf, p.hist, p.x, p.nx);
    return cudaGetLastError();
}

} // namespace b2
