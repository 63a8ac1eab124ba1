// channelizer8.cu -- analysis channelizer for K = 2N in {64, 128, 256, 512}: the NCO mix-down,
// the polyphase FIR and the K-point FFT of
//     nco_crcf_mix_down + nco_crcf_step            lib/multichannelrx.cc:163-164   (x K)
//     firpfbch_crcf_analyzer_execute(x -> X)       lib/multichannelrx.cc:188
// fused in one pass over the wideband stream, keeping X[0..N-1] (lib/multichannelrx.cc:193-194)
// channel-major.  Math as in channelizer.cu:
//     V_b[r] = sum_{n<P} h[(K-1-r) + n*K] * xm[(b-n)*K + r],   y_b = FFT_K(V_b)
//
// One CTA = K threads = one filterbank; it walks a contiguous range of blocks 8 at a time:
//   * TMA (1-D cp.async.bulk, mbarrier completion) stages the 8 new rows of K samples while the
//     previous 8 are being transformed;
//   * thread r OWNS COLUMN r: the last P-1 mixed samples of its column live in registers and
//     slide by 8 per round, so a staged sample is read from shared memory exactly once and the
//     FIR is 14 packed FFMA2 (real tap x complex sample) per output out of registers;
//   * the 8 FIR rows are transformed by 8 groups of K/8 threads, one row each, with the
//     register-resident Stockham radix-8 FFT of fft8.cuh (8 points per thread, two exchanges
//     for K = 512); only the N kept channels are written, transposed through shared memory so
//     that each channel receives one contiguous 64-byte run per round.
// Shared memory is 122 KB and the register file is full at K = 512 (one 512-thread CTA per SM), which is
// why the synchroniser chains get their own SMs (smpart.cu) instead of sharing these.
#include "kernels.h"
#include "fft8.cuh"

namespace b2 {

constexpr unsigned int A8_JB = 8;       // rows per round
constexpr unsigned int A8_P = 14;       // taps per branch (m = 7, lib/multichannelrx.cc:89)
constexpr unsigned int A8_OLD = 2 * A8_JB + 1;   // row length of the output staging tile: two rounds of columns + 1 (bank skew)

struct A8Layout { size_t off_bar, off_rw, off_tw, off_stage, off_a, off_b, off_out, total; };
__host__ __device__ constexpr A8Layout a8_layout(unsigned int K)
{
    A8Layout L{};
    size_t o = 0;
    L.off_bar = o;   o += 16;
    L.off_rw = o;    o += (2 * A8_JB + 16) * sizeof(cf);
    L.off_tw = o;    o += (size_t)K * sizeof(cf);
    o = (o + 127) & ~(size_t)127;
    L.off_stage = o; o += (size_t)A8_JB * K * sizeof(cf);
    L.off_a = o;     o += (size_t)A8_JB * f8_buf_elems(K) * sizeof(cf);
    L.off_b = o;     o += (size_t)A8_JB * f8_buf_elems(K) * sizeof(cf);
    L.off_out = o;   o += (size_t)(K / 2) * A8_OLD * sizeof(cf);
    L.total = o;
    return L;
}

template <unsigned int K>
__global__ void __launch_bounds__(K, 1) analyzer8_kernel(const AnalyzerParams p)
{
    constexpr unsigned int T8 = K / 8, JB = A8_JB, P = A8_P, BUF = f8_buf_elems(K), N = K / 2;
    constexpr A8Layout L = a8_layout(K);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t * bar = (uint64_t *)(smem + L.off_bar);
    cf * rw = (cf *)(smem + L.off_rw);                  // [2][JB] row phasors of a round, then [16] for the warm-up
    cf * tw = (cf *)(smem + L.off_tw);
    cf * stage = (cf *)(smem + L.off_stage);
    cf * bufA = (cf *)(smem + L.off_a);
    cf * bufB = (cf *)(smem + L.off_b);
    cf * outst = (cf *)(smem + L.off_out);
    const unsigned int r = threadIdx.x, g = r / T8, j = r % T8;

    // contiguous range of rounds of this CTA
    const unsigned int rounds_total = (p.nblocks + JB - 1) / JB;
    const unsigned int rounds_per = (rounds_total + gridDim.x - 1) / gridDim.x;
    const unsigned int B0 = blockIdx.x * rounds_per * JB;
    if (B0 >= p.nblocks) return;
    const unsigned int B1 = min(p.nblocks, B0 + rounds_per * JB);
    const unsigned int nrounds = (B1 - B0 + JB - 1) / JB;
    const unsigned int L0 = p.block0 + B0;              // logical row of the oldest history row

    auto row_ptr = [&](unsigned int row) -> const cf * {
        return row < p.rows0 ? p.seg0 + (size_t)row * K : p.seg1 + (size_t)(row - p.rows0) * K;
    };
    // one elected thread: rows [row, row + n) -> stage, one bulk copy per contiguous run
    auto issue = [&](unsigned int row, unsigned int n) {
        mbar_expect_tx(bar, n * K * (unsigned int)sizeof(cf));
        unsigned int done = 0;
        while (done < n) {
            unsigned int run = n - done;
            if (row + done < p.rows0) run = min(run, p.rows0 - (row + done));
            bulk_g2s(stage + (size_t)done * K, row_ptr(row + done), run * K * (unsigned int)sizeof(cf), bar);
            done += run;
        }
    };

    if (r == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    f8_twt_build<K, 1>(tw, p.fft.tw, r, K);          // per-pass twiddle tables in lane order
    if (r < P - 1) rw[2 * JB + r] = nco_cexp_pi(p.theta0 + (L0 + r) * K * p.dtheta);
    if (r >= 32 && r < 32 + JB) rw[r - 32] = nco_cexp_pi(p.theta0 + (L0 + P - 1 + (r - 32)) * K * p.dtheta);
    __syncthreads();
    if (r == 0) issue(L0 + P - 1, min(JB, B1 - B0));

    float h[P];
#pragma unroll
    for (unsigned int n = 0; n < P; n++) h[n] = __ldg(p.taps + (size_t)n * K + (K - 1 - r));
    const cf colw = nco_cexp_pi(r * p.dtheta);

    // warm-up: the P-1 rows before the first block, straight from global memory
    cf a[P - 1 + JB];
#pragma unroll
    for (unsigned int q = 0; q < P - 1; q++) a[q] = mix_down(row_ptr(L0 + q)[r], cmul(rw[2 * JB + q], colw));

    for (unsigned int k = 0; k < nrounds; k++) {
        const unsigned int nb = min(JB, B1 - (B0 + k * JB));
        const cf * rwk = rw + (k & 1) * JB;
        mbar_wait(bar, k & 1);
        if (p.row_alt) {
            // the NCO of multichannelrx advances by a multiple of pi per block (K*dtheta = -(N-1) pi,
            // lib/multichannelrx.cc:98): one phasor per round, the sign alternates or stays
            const cf w0 = cmul(rwk[0], colw);
            const cf w1 = p.row_alt < 0 ? make_float2(-w0.x, -w0.y) : w0;
#pragma unroll
            for (unsigned int i = 0; i < JB; i++) a[P - 1 + i] = mix_down(stage[i * K + r], (i & 1) ? w1 : w0);
        } else {
#pragma unroll
            for (unsigned int i = 0; i < JB; i++) a[P - 1 + i] = mix_down(stage[i * K + r], cmul(rwk[i], colw));
        }
        __syncthreads();                                 // stage consumed
        if (k + 1 < nrounds) {
            if (r == 0) {
                fence_proxy_async();
                issue(L0 + P - 1 + (k + 1) * JB, min(JB, B1 - (B0 + (k + 1) * JB)));
            }
            if (r >= 32 && r < 32 + JB)
                rw[((k + 1) & 1) * JB + (r - 32)] = nco_cexp_pi(p.theta0 + (L0 + P - 1 + (k + 1) * JB + (r - 32)) * K * p.dtheta);
        }
        // polyphase FIR out of registers: block i of the round uses a[i .. i+P-1], newest first
#pragma unroll
        for (unsigned int i = 0; i < JB; i++) {
            cf acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = P - 1; n >= 0; n--) acc = cfma_real(h[n], a[i + P - 1 - n], acc);   // oldest sample first, as dotprod_crcf
            bufA[i * BUF + f8_pad(r)] = acc;
        }
#pragma unroll
        for (unsigned int q = 0; q < P - 1; q++) a[q] = a[q + JB];
        __syncthreads();

        // K-point forward FFT of row g by group g
        cf v[8];
        f8_load<K>(v, j, bufA + g * BUF);
        // (the exchanges of a row concern its own K/8 threads only: a named barrier per group lets the eight groups
        // drift apart, one in its butterflies while another waits for shared memory)
        if constexpr (T8 >= 32 && T8 % 32 == 0)
            f8_run<K, 1, -1>(v, j, bufB + g * BUF, bufA + g * BUF, nullptr, [g] { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T8) : "memory"); }, nullptr, tw);
        else
            f8_run<K, 1, -1>(v, j, bufB + g * BUF, bufA + g * BUF, nullptr, [] { __syncthreads(); }, nullptr, tw);

        // channels 0..N-1 -> out[c][col0 + block], via a transposed tile that collects TWO rounds: every channel then
        // receives one contiguous 128-byte run per flush (full lines in HBM, and full-size packets when the run goes to a
        // peer GPU over NVLink in the multi-GPU split)
        const unsigned int half = k & 1u;
#pragma unroll
        for (unsigned int s = 0; s < 4; s++) outst[(j + s * T8) * A8_OLD + half * JB + g] = v[s];
        if (half == 1u || k + 1 == nrounds) {
            __syncthreads();
            const unsigned int ncols = half * JB + nb;
            const size_t col = p.out_col0 + B0 + (k - half) * JB;
            for (unsigned int e = r; e < N * JB; e += K) {
                const unsigned int c = e / JB, q2 = (e % JB) * 2;
                if (q2 >= ncols) continue;
                // (multi-GPU split: the run goes to the GPU that owns channel c)
                cf * row = p.n_peer ? p.out_peer[c / p.chan_per_peer] + (size_t)(c % p.chan_per_peer) * p.out_stride
                                    : p.out + (size_t)c * p.out_stride;
                cf * dst = row + col + q2;
                const cf x0 = outst[c * A8_OLD + q2], x1 = outst[c * A8_OLD + q2 + 1];
                if (q2 + 1 < ncols && ((((size_t)dst) & 15) == 0)) *(float4 *)dst = make_float4(x0.x, x0.y, x1.x, x1.y);
                else { dst[0] = x0; if (q2 + 1 < ncols) dst[1] = x1; }
            }
        }
        // the next round's outst / bufA writes come after its own barriers
    }
}

template <unsigned int K>
static cudaError_t analyzer8_launch_t(const AnalyzerParams & p, cudaStream_t st)
{
    static int configured[16] = {0};
    static int per_sm[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    const size_t smem = a8_layout(K).total;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(analyzer8_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int nb = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, analyzer8_kernel<K>, (int)K, smem);
        per_sm[dev] = nb < 1 ? 1 : nb;
        configured[dev] = 1;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (p.sm_limit && (int)p.sm_limit < sms) sms = (int)p.sm_limit;
    const unsigned int rounds = (p.nblocks + A8_JB - 1) / A8_JB;
    unsigned int grid = (unsigned int)(sms * per_sm[dev]);
    if (grid > rounds) grid = rounds;
    analyzer8_kernel<K><<<grid, K, smem, st>>>(p);
    return cudaGetLastError();
}

bool analyzer8_supported(const AnalyzerParams & p)
{
    return p.P == A8_P && (p.K == 64 || p.K == 128 || p.K == 256 || p.K == 512) && p.N * 2 == p.K;
}

cudaError_t analyzer8_launch(const AnalyzerParams & p0, cudaStream_t st)
{
    if (p0.nblocks == 0) return cudaSuccess;
    AnalyzerParams p = p0;
    const uint32_t kd = p.K * p.dtheta;
    p.row_alt = (kd == 0u) ? 1 : ((kd == 0x80000000u) ? -1 : 0);
    switch (p.K) {
    case 64:  return analyzer8_launch_t<64>(p, st);
    case 128: return analyzer8_launch_t<128>(p, st);
    case 256: return analyzer8_launch_t<256>(p, st);
    case 512: return analyzer8_launch_t<512>(p, st);
    default:  return cudaErrorInvalidValue;
    }
}

} // namespace b2
