// syncdev.cuh -- device helpers shared by the two synchroniser kernels (ofdmsync.cu: generic,
// any power-of-two M; ofdmsync8.cu: register-resident fast path for M >= 256): hard demapper,
// packet length arithmetic, warp-parallel phase unwrap, 5x5 solver for the S1 equaliser-gain
// polynomial fit, header CRC.
#pragma once
#include "kernels.h"
#include "fec.cuh"

namespace b2 {

#define PI_F 3.14159274101257324219f

// hard demodulation of one symbol (liquid modem_demodulate for BPSK / QPSK / square QAM)
__device__ __forceinline__ unsigned int demod_axis(float v, int m, float alpha)
{
    unsigned int s = 0;
    for (int k = m - 1; k >= 0; k--) {
        float ref = (float)(1u << k) * alpha;
        s <<= 1;
        if (v > 0) { s |= 1u; v -= ref; } else { v += ref; }
    }
    return s ^ (s >> 1);        // gray encode
}
__device__ __forceinline__ unsigned int demod_symbol(cf x, unsigned int scheme, unsigned int bps, float alpha)
{
    if (scheme == 39) return x.x > 0 ? 0u : 1u;                                   // BPSK
    if (scheme == 40) return (x.x > 0 ? 0u : 1u) + (x.y > 0 ? 0u : 2u);           // QPSK
    int m = (int)bps >> 1;
    unsigned int si = demod_axis(x.x, m, alpha);
    unsigned int sq = demod_axis(x.y, m, alpha);
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
__device__ __forceinline__ unsigned int div_bps(unsigned int x, unsigned int bps)
{
    switch (bps) {
    case 1: return x;
    case 2: return x >> 1;
    case 4: return x >> 2;
    case 8: return x >> 3;
    default: return (x * 0xAAABu) >> 18;        // bps = 6, x < 2^15
    }
}

// warp-parallel phase unwrap of y[0..n) (liquid: "while (y[i]-y[i-1] > pi) y[i] -= 2pi" ...),
// in place when `store`, returning per-lane partial sums of y' and x*y' (x may be null).
// The number of 2*pi steps of element i is the running sum of the steps implied by the RAW
// neighbour differences; the value itself is then built by repeated float adds like liquid does.
__device__ __forceinline__ void warp_unwrap(float * y, const float * __restrict__ x, unsigned int n, bool store,
                                            unsigned int lane, float & sy, float & sxy)
{
    int carry = 0;
    float last_raw = 0.f;
    sy = 0.f; sxy = 0.f;
    for (unsigned int base = 0; base < n; base += 32) {
        const unsigned int i = base + lane;
        const bool valid = i < n;
        const float raw = valid ? y[i] : 0.f;
        float prev = __shfl_up_sync(0xffffffffu, raw, 1);
        if (lane == 0) prev = last_raw;
        int k = 0;
        if (valid && i > 0) {
            float d = raw - prev;
            k = (d > PI_F) ? -1 : ((d < -PI_F) ? 1 : 0);
        }
        if (__any_sync(0xffffffffu, k != 0)) {      // rare: most symbols need no unwrapping at all
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, k, o);
                if (lane >= (unsigned int)o) k += t;
            }
        }
        k += carry;
        carry = __shfl_sync(0xffffffffu, k, 31);
        last_raw = __shfl_sync(0xffffffffu, raw, 31);
        float yy = raw;
        for (int q = k; q > 0; q--) yy += 2 * PI_F;
        for (int q = k; q < 0; q++) yy -= 2 * PI_F;
        if (valid) {
            if (store) y[i] = yy;
            sy = __fadd_rn(sy, yy);
            if (x) sxy = __fadd_rn(sxy, __fmul_rn(x[i], yy));
        }
    }
}

// the same unwrap for long arrays (the S1 gain phases, one per active subcarrier), in place: each
// lane walks one contiguous segment twice -- first to count its 2*pi steps, then, after a warp
// scan of the counts, to apply them -- instead of 32 elements per warp-wide iteration
__device__ __forceinline__ void warp_unwrap_seg(float * y, unsigned int n, unsigned int lane)
{
    const unsigned int L = (n + 31) / 32;
    const unsigned int lo = min(n, lane * L), hi = min(n, lo + L);
    const float before = (lo > 0 && lo < n) ? y[lo - 1] : 0.f;      // raw value left of the segment
    int total = 0;
    {
        float prev = before;
        for (unsigned int i = lo; i < hi; i++) {
            const float raw = y[i];
            if (i > 0) {
                const float d = raw - prev;
                total += (d > PI_F) ? -1 : ((d < -PI_F) ? 1 : 0);
            }
            prev = raw;
        }
    }
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned int)o) incl += t;
    }
    __syncwarp();                                                    // every lane has read its `before`
    int k = incl - total;
    float prev = before;
    for (unsigned int i = lo; i < hi; i++) {
        const float raw = y[i];
        if (i > 0) {
            const float d = raw - prev;
            k += (d > PI_F) ? -1 : ((d < -PI_F) ? 1 : 0);
        }
        prev = raw;
        float yy = raw;
        for (int q = k; q > 0; q--) yy += 2 * PI_F;
        for (int q = k; q < 0; q++) yy -= 2 * PI_F;
        y[i] = yy;
    }
}

// solve the 5x5 normal equations sum_c S[r+c] p[c] = b[r] (Gaussian elimination, partial pivoting)
__device__ __forceinline__ void solve5(const double * __restrict__ S, const double * __restrict__ b, double * __restrict__ coef)
{
    double A[5][6];
#pragma unroll
    for (int r = 0; r < 5; r++) {
#pragma unroll
        for (int c = 0; c < 5; c++) A[r][c] = S[r + c];
        A[r][5] = b[r];
    }
#pragma unroll
    for (int c = 0; c < 5; c++) {
#pragma unroll
        for (int r = c + 1; r < 5; r++) {
            if (fabs(A[r][c]) > fabs(A[c][c])) {
#pragma unroll
                for (int i = 0; i < 6; i++) { double t = A[c][i]; A[c][i] = A[r][i]; A[r][i] = t; }
            }
        }
        const double inv = 1.0 / A[c][c];          // one reciprocal per pivot (double division is slow)
        A[c][c] = inv;
#pragma unroll
        for (int r = c + 1; r < 5; r++) {
            double f = A[r][c] * inv;
#pragma unroll
            for (int i = c + 1; i < 6; i++) A[r][i] -= f * A[c][i];
        }
    }
    double p[5];
#pragma unroll
    for (int r = 4; r >= 0; r--) {
        double s = A[r][5];
#pragma unroll
        for (int c = r + 1; c < 5; c++) s -= A[r][c] * p[c];
        p[r] = s * A[r][r];                        // diagonal holds the reciprocal pivot
    }
#pragma unroll
    for (int i = 0; i < 5; i++) coef[i] = p[i];
}

// atan2 for the pilot phases: minimax odd polynomial on [0, 1] (|err| < 1e-7 rad), one
// approximate division; quadrant handling as atan2f (the arguments are never both zero here)
__device__ __forceinline__ float atan2_fast(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = __fdividef(mn, mx);
    const float s = a * a;
    float r = fmaf(s, 0.002456609858199954f, -0.01440086867660284f);
    r = fmaf(r, s, 0.03978036344051361f);
    r = fmaf(r, s, -0.07234777510166168f);
    r = fmaf(r, s, 0.10498903691768646f);
    r = fmaf(r, s, -0.14161217212677002f);
    r = fmaf(r, s, 0.19985905289649963f);
    r = fmaf(r, s, -0.33332598209381104f);
    r = fmaf(r, s, 0.9999998807907104f);
    r *= a;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.f) r = 3.14159274f - r;
    return copysignf(r, y);
}

// nco_constrain_dev for |theta| < 2 pi (the NCO trims): p - floor(p) is p or p + 1 there
__device__ __forceinline__ uint32_t nco_constrain_small(float theta)
{
    const double p = (double)theta * 0.15915494309189535;
    const double f = p < 0.0 ? p + 1.0 : p;
    return (uint32_t)(__double2ull_rn(f * 4294967296.0) & 0xffffffffull);
}

// hard demapper with the constellation known at compile time (same decisions as demod_symbol).
// liquid walks the levels by successive approximation, v_{k+1} = v_k -+ 2^k alpha with the sign of v_k, and Gray-codes
// the binary index.  With u_k = |v_k| the same float operations read u_{k+1} = |u_k| - 2^k alpha (v_{k+1} = +-u_{k+1},
// negation is exact), and the Gray bits are simply [v > 0, u_1 < 0, u_2 < 0, ...]: a bit of s ^ (s >> 1) is set
// where two consecutive decisions differ, i.e. where the magnitude fell below the level that was subtracted.
template <int MB> __device__ __forceinline__ unsigned int demod_axis_t(float v, float alpha)
{
    unsigned int g = (v > 0) ? (1u << (MB - 1)) : 0u;
    float u = v;
#pragma unroll
    for (int k = MB - 1; k >= 1; k--) {
        u = fabsf(u) - (float)(1u << k) * alpha;
        g |= (u < 0) ? (1u << (k - 1)) : 0u;
    }
    return g;
}
template <int MB> __device__ __forceinline__ unsigned int demod_qam_t(cf x, float alpha)
{
    return (demod_axis_t<MB>(x.x, alpha) << MB) + demod_axis_t<MB>(x.y, alpha);
}

// CRC-32 with a 16-entry nibble table (frame header: 14 bytes)
static __constant__ uint32_t b2_crc_nibble_table[16] = {
    0x00000000u, 0x1DB71064u, 0x3B6E20C8u, 0x26D930ACu, 0x76DC4190u, 0x6B6B51F4u, 0x4DB26158u, 0x5005713Cu,
    0xEDB88320u, 0xF00F9344u, 0xD6D6A3E8u, 0xCB61B38Cu, 0x9B64C2B0u, 0x86D3D2D4u, 0xA00AE278u, 0xBDBDF21Cu};
__device__ __forceinline__ uint32_t crc32_nibble(const uint8_t * m, unsigned int n)
{
    const uint32_t * T = b2_crc_nibble_table;
    uint32_t key = ~0u;
    for (unsigned int i = 0; i < n; i++) {
        key ^= m[i];
        key = (key >> 4) ^ T[key & 15u];
        key = (key >> 4) ^ T[key & 15u];
    }
    return ~key;
}

} // namespace b2
