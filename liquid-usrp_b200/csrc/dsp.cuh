// dsp.cuh -- device-side building blocks shared by the kernels: complex helpers, the uint32-phase
// NCO (nco_crcf, LIQUID_VCO; lib/multichannelrx.cc:99-100,163-164), a mixed-radix in-place
// shared-memory FFT (the K-point transform inside firpfbch_crcf_*_execute and the M-point
// transform inside ofdmframe{gen,sync}), warp/block reductions for the S0/S1 correlators.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

typedef float2 cf;

// Packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: one instruction per complex add, scale or
// multiply-accumulate instead of two; ptxas folds the re/im swaps and per-half sign flips of the complex
// products and of the +-j rotations into operand modifiers).  Same IEEE rounding per component as the scalar ops.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ cf f2_unpack(unsigned long long v)
{
    cf r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ cf f2_add(cf a, cf b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(d);
}
__device__ __forceinline__ cf f2_sub(cf a, cf b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(d);
}
__device__ __forceinline__ cf f2_mul(cf a, cf b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)));
    return f2_unpack(d);
}
__device__ __forceinline__ cf f2_fma(cf a, cf b, cf c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a.x, a.y)), "l"(f2_pack(b.x, b.y)), "l"(f2_pack(c.x, c.y)));
    return f2_unpack(d);
}

// a*b = a.x*(b.x, b.y) + (-t.x, t.y) with t = a.y*(b.y, b.x): FMUL2 + FFMA2, the swap and the per-half
// sign are operand modifiers
__device__ __forceinline__ cf cmul(cf a, cf b)
{
    const cf t = f2_mul(make_float2(a.y, a.y), make_float2(b.y, b.x));
    return f2_fma(make_float2(a.x, a.x), b, make_float2(-t.x, t.y));
}
// a*conj(b) = b.x*(a.x, a.y) + (t.x, -t.y) with t = b.y*(a.y, a.x)
__device__ __forceinline__ cf cmulc(cf a, cf b)
{
    const cf t = f2_mul(make_float2(b.y, b.y), make_float2(a.y, a.x));
    return f2_fma(make_float2(b.x, b.x), a, make_float2(t.x, -t.y));
}
__device__ __forceinline__ cf cadd(cf a, cf b) { return f2_add(a, b); }
__device__ __forceinline__ cf csub(cf a, cf b) { return f2_sub(a, b); }
__device__ __forceinline__ cf cscale(cf a, float s) { return f2_mul(a, make_float2(s, s)); }
// acc + h*x, real tap times complex sample (the polyphase FIR inner step)
__device__ __forceinline__ cf cfma_real(float h, cf x, cf acc) { return f2_fma(make_float2(h, h), x, acc); }

// e^{+j theta} for a uint32 phase (2*pi <-> 2^32): cos/sin of (int32)theta * pi / 2^31
__device__ __forceinline__ cf nco_cexp(uint32_t theta)
{
    float t = (float)((double)(int32_t)theta * (3.14159265358979323846 / 2147483648.0));
    float s, c;
    sincosf(t, &s, &c);
    return make_float2(c, s);
}
// cheaper variant: same phase in half-turns, evaluated by sincospif (exact argument reduction)
__device__ __forceinline__ cf nco_cexp_pi(uint32_t theta)
{
    float s, c;
    sincospif((float)((double)(int32_t)theta * (1.0 / 2147483648.0)), &s, &c);
    return make_float2(c, s);
}
// fast variant for the per-sample mixers: float phase in (-pi, pi], SFU sine/cosine
// (absolute error ~4e-7 on this range, one order below the 1e-5 budget on equalised symbols)
__device__ __forceinline__ cf nco_cexp_fast(uint32_t theta)
{
    float t = (float)(int32_t)theta * (3.14159265358979323846f / 2147483648.0f);
    float s, c;
    __sincosf(t, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ cf mix_down(cf x, cf w) { return cmulc(x, w); } // x*conj(w)
__device__ __forceinline__ cf mix_up(cf x, cf w) { return cmul(x, w); }

// radians -> uint32 phase, same rounding as the host (design.h nco_constrain)
__device__ __forceinline__ uint32_t nco_constrain_dev(float theta)
{
    double p = (double)theta * 0.15915494309189535;
    double f = p - floor(p);
    double u = rint(f * 4294967296.0);
    return (uint32_t)((unsigned long long)u & 0xffffffffull);
}
__device__ __forceinline__ float nco_freq_dev(uint32_t dtheta)
{
    return (float)((double)(int32_t)dtheta * (3.14159265358979323846 / 2147483648.0));
}

// ------------------------------------------------------------------ small DFT butterflies
// DIR = -1 forward (e^{-j}), +1 backward (e^{+j}); all unnormalised
template <int DIR> __device__ __forceinline__ cf mul_j(cf a)   // multiply by DIR*j
{
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
template <int DIR> __device__ __forceinline__ void dft2(cf & a, cf & b)
{
    cf t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <int DIR> __device__ __forceinline__ void dft4(cf * v)
{
    cf a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    cf a2 = cadd(v[1], v[3]), a3 = mul_j<DIR>(csub(v[1], v[3]));
    v[0] = cadd(a0, a2); v[2] = csub(a0, a2);
    v[1] = cadd(a1, a3); v[3] = csub(a1, a3);
}
template <int DIR> __device__ __forceinline__ void dft8(cf * v)
{
    const float h = 0.70710678118654752440f;
    // two interleaved radix-4 on even / odd elements, then combine with w8^k
    cf e[4] = {v[0], v[2], v[4], v[6]};
    cf o[4] = {v[1], v[3], v[5], v[7]};
    dft4<DIR>(e);
    dft4<DIR>(o);
    // w8^1 = (1 + DIR*j)/sqrt2, w8^2 = DIR*j, w8^3 = (-1 + DIR*j)/sqrt2
    // w8^1 x = (x + DIR*j*x) * h,  w8^3 x = (-x + DIR*j*x) * h
    cf t1 = cscale(cadd(o[1], mul_j<DIR>(o[1])), h);
    cf t2 = mul_j<DIR>(o[2]);
    cf t3 = cscale(csub(mul_j<DIR>(o[3]), o[3]), h);
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], t1);   v[5] = csub(e[1], t1);
    v[2] = cadd(e[2], t2);   v[6] = csub(e[2], t2);
    v[3] = cadd(e[3], t3);   v[7] = csub(e[3], t3);
}

// skewed shared-memory index (one pad element per 32) used where the access stride is a
// multiple of the bank count (channelizer FFT rows)
template <int PAD> __device__ __forceinline__ unsigned int phys(unsigned int i) { return PAD == 2 ? i + (i >> 3) : (PAD ? i + (i >> 5) : i); }

// 4-bit code of a pass radix: 2, 4, 8 and the odd primes up to 13 stand for themselves, the spare codes
// carry the primes 17 .. 41
__host__ __device__ constexpr unsigned int fft_radix_code(unsigned int R)
{
    return R <= 13 ? R : (R == 17 ? 1u : R == 19 ? 6u : R == 23 ? 9u : R == 29 ? 10u : R == 31 ? 12u : R == 37 ? 14u : 15u);
}
__host__ __device__ constexpr unsigned int fft_radix_of(unsigned int code)
{
    return code == 1 ? 17u : code == 6 ? 19u : code == 9 ? 23u : code == 10 ? 29u : code == 12 ? 31u : code == 14 ? 37u : code == 15 ? 41u : code;
}
constexpr unsigned int FFT_MAX_RADIX = 41;

struct FftDev {
    unsigned int n, npass;
    unsigned int radices;      // fft_radix_code of pass t in bits [4t, 4t+4)
    const uint16_t * perm;     // input permutation (device or shared)
    const cf * tw;             // forward twiddles e^{-j 2 pi k / n}
};

// one radix-R pass over `nfft` transforms of length n = 2^lgn laid out back to back in `buf` (row
// stride `ld` elements); 2^lgL = sub-transform length after this pass.  All sizes are powers of
// two, so the butterfly index splits with shifts and masks only.
template <int R, int DIR, int PAD>
__device__ __forceinline__ void fft_pass(cf * buf, unsigned int ld, unsigned int nfft, unsigned int lgn, unsigned int lgL,
                                         const cf * __restrict__ tw, unsigned int tid, unsigned int nthreads)
{
    constexpr unsigned int lgR = (R == 8) ? 3 : ((R == 4) ? 2 : 1);
    const unsigned int lgs = lgL - lgR;           // log2 of the stride between butterfly inputs
    const unsigned int lgper = lgn - lgR;         // log2 of the butterflies per transform
    const unsigned int s = 1u << lgs;
    const unsigned int tshift = lgn - lgL;        // twiddle index step: w_L^j = tw[j << tshift]
    const unsigned int total = nfft << lgper;
    for (unsigned int w = tid; w < total; w += nthreads) {
        const unsigned int f = w >> lgper, b = w & ((1u << lgper) - 1u);
        const unsigned int blk = b >> lgs, j = b & (s - 1u);
        cf * x = buf + (size_t)f * ld;
        const unsigned int i0 = (blk << lgL) + j;
        cf v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = x[phys<PAD>(i0 + (r << lgs))];
        if (lgs > 0) {
#pragma unroll
            for (int r = 1; r < R; r++) {
                cf t = tw[(j * r) << tshift];
                if (DIR > 0) t.y = -t.y;
                v[r] = cmul(v[r], t);
            }
        }
        if (R == 2) dft2<DIR>(v[0], v[1]);
        else if (R == 4) dft4<DIR>(v);
        else dft8<DIR>(v);
#pragma unroll
        for (int r = 0; r < R; r++) x[phys<PAD>(i0 + (r << lgs))] = v[r];
    }
}

// the same pass for ANY transform length and radix (sizes that are not powers of two: M = 48, K = 6, ...):
// integer division instead of shifts, a direct R x R DFT per butterfly.  L = sub-transform length
// after this pass.
template <int DIR, int PAD>
__device__ __noinline__ void fft_pass_any(cf * buf, unsigned int ld, unsigned int nfft, unsigned int n, unsigned int L, unsigned int R,
                                          const cf * __restrict__ tw, unsigned int tid, unsigned int nthreads)
{
    const unsigned int s = L / R, per = n / R, tstep = n / L, rstep = n / R;
    const unsigned int total = nfft * per;
    for (unsigned int w = tid; w < total; w += nthreads) {
        const unsigned int f = w / per, b = w - f * per;
        const unsigned int blk = b / s, j = b - blk * s;
        cf * x = buf + (size_t)f * ld;
        const unsigned int i0 = blk * L + j;
        cf v[FFT_MAX_RADIX], y[FFT_MAX_RADIX];
        for (unsigned int r = 0; r < R; r++) {
            cf a = x[phys<PAD>(i0 + r * s)];
            if (s > 1 && r > 0) {
                cf t = tw[j * r * tstep];
                if (DIR > 0) t.y = -t.y;
                a = cmul(a, t);
            }
            v[r] = a;
        }
        for (unsigned int q = 0; q < R; q++) {
            cf acc = v[0];
            unsigned int m = 0;
            for (unsigned int r = 1; r < R; r++) {
                m += q;
                if (m >= R) m -= R;
                cf t = tw[m * rstep];
                if (DIR > 0) t.y = -t.y;
                acc = cadd(acc, cmul(v[r], t));
            }
            y[q] = acc;
        }
        for (unsigned int q = 0; q < R; q++) x[phys<PAD>(i0 + q * s)] = y[q];
    }
}

// in-place DIT over data already stored in permuted order; natural-order output.
// All threads of the CTA must call; ends with a __syncthreads().
template <int DIR, int PAD>
__device__ __forceinline__ void fft_inplace(cf * buf, unsigned int ld, unsigned int nfft, const FftDev & f,
                                            unsigned int tid, unsigned int nthreads)
{
    if (f.n & (f.n - 1)) {                       // not a power of two
        unsigned int L = 1;
        for (unsigned int t = 0; t < f.npass; t++) {
            const unsigned int R = fft_radix_of((f.radices >> (4 * t)) & 15u);
            L *= R;
            fft_pass_any<DIR, PAD>(buf, ld, nfft, f.n, L, R, f.tw, tid, nthreads);
            __syncthreads();
        }
        return;
    }
    const unsigned int lgn = 31u - (unsigned int)__clz((int)f.n);
    unsigned int lgL = 0;
    for (unsigned int t = 0; t < f.npass; t++) {
        unsigned int R = (f.radices >> (4 * t)) & 15u;
        if (R == 8) { lgL += 3; fft_pass<8, DIR, PAD>(buf, ld, nfft, lgn, lgL, f.tw, tid, nthreads); }
        else if (R == 4) { lgL += 2; fft_pass<4, DIR, PAD>(buf, ld, nfft, lgn, lgL, f.tw, tid, nthreads); }
        else { lgL += 1; fft_pass<2, DIR, PAD>(buf, ld, nfft, lgn, lgL, f.tw, tid, nthreads); }
        __syncthreads();
    }
}

// the same pass plan as design.h fft_plan() (radix-4 passes first when log2 n is not a multiple
// of 3, radix-8 for the rest), with every size known at compile time
__host__ __device__ constexpr unsigned int fft_static_radices(unsigned int n)
{
    unsigned int lg = 0;
    while ((1u << lg) < n) lg++;
    unsigned int r = 0, np = 0, rem = lg;
    if (rem % 3 == 1 && rem >= 4) { r |= 4u << (4 * np++); r |= 4u << (4 * np++); rem -= 4; }
    else if (rem % 3 == 1) { r |= 2u << (4 * np++); rem -= 1; }
    else if (rem % 3 == 2) { r |= 4u << (4 * np++); rem -= 2; }
    while (rem >= 3) { r |= 8u << (4 * np++); rem -= 3; }
    return r;
}
template <unsigned int N, int DIR, int PAD>
__device__ __forceinline__ void fft_static(cf * buf, const cf * __restrict__ tw, unsigned int tid, unsigned int nthreads)
{
    constexpr unsigned int radices = fft_static_radices(N);
    constexpr unsigned int lgn = (N <= 2) ? 1 : (N <= 4) ? 2 : (N <= 8) ? 3 : (N <= 16) ? 4 : (N <= 32) ? 5 : (N <= 64) ? 6 : (N <= 128) ? 7 :
                                 (N <= 256) ? 8 : (N <= 512) ? 9 : (N <= 1024) ? 10 : (N <= 2048) ? 11 : 12;
    unsigned int lgL = 0;
#pragma unroll
    for (unsigned int t = 0; t < 6; t++) {
        const unsigned int R = (radices >> (4 * t)) & 15u;
        if (R == 0) break;
        if (R == 8) { lgL += 3; fft_pass<8, DIR, PAD>(buf, N, 1, lgn, lgL, tw, tid, nthreads); }
        else if (R == 4) { lgL += 2; fft_pass<4, DIR, PAD>(buf, N, 1, lgn, lgL, tw, tid, nthreads); }
        else { lgL += 1; fft_pass<2, DIR, PAD>(buf, N, 1, lgn, lgL, tw, tid, nthreads); }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of a complex value; scratch >= 2*32 floats; result broadcast to all threads.
// fixed tree: lane butterflies, then warp partials summed in warp order.
__device__ __forceinline__ cf block_sum_cf(cf v, float * scratch, unsigned int tid, unsigned int nthreads)
{
    v.x = warp_sum(v.x);
    v.y = warp_sum(v.y);
    unsigned int nw = (nthreads + 31) >> 5;
    __syncthreads();
    if ((tid & 31) == 0) { scratch[2 * (tid >> 5)] = v.x; scratch[2 * (tid >> 5) + 1] = v.y; }
    __syncthreads();
    cf r = make_float2(0.f, 0.f);
    for (unsigned int w = 0; w < nw; w++) { r.x += scratch[2 * w]; r.y += scratch[2 * w + 1]; }
    return r;
}

// ------------------------------------------------------------------ mbarrier + bulk async copy (TMA unit, 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, unsigned int parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, unsigned int parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void * dst_smem, const void * src_gmem, unsigned int bytes, uint64_t * bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 8-byte asynchronous global -> shared copy (LDGSTS), completion via cp.async.wait_all
__device__ __forceinline__ void cp_async8(void * dst_smem, const void * src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace b2
