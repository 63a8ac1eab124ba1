// smpart.cu -- spatial partition of the GPU's SMs between the two kinds of work on the receive
// path, through CUDA green contexts (driver API, CUDA >= 12.4):
//   * the per-channel synchronisers are 256 SERIAL chains of short events: they need few SMs but
//     need them all the time (a chain that waits for an SM stalls the whole pipeline);
//   * the channelizer / packet decode are throughput kernels that fill whatever they are given
//     (a 512-thread channelizer CTA takes the whole register file of an SM, so the two kinds
//     cannot simply share SMs).
// Streams created on the two green contexts confine their kernels to disjoint SM sets, so the
// synchroniser of chunk c runs undisturbed beside the channelizer of chunk c+1.
// The driver entry points are fetched with cudaGetDriverEntryPoint, so the library has no link
// dependency on libcuda (it must load on machines without a driver, for the ABI tests).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "smpart.h"

namespace b2 {

template <typename F> static bool entry(const char * name, F & fn)
{
    void * p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = (F)p;
    return true;
}

struct GreenPair { bool tried = false, ok = false; unsigned int req = 0, small_sms = 0, big_sms = 0; CUgreenCtx g_small = nullptr, g_big = nullptr; };
static GreenPair g_pairs[16][12];         // per device, one pair per requested size (sizes are multiples of 8)

bool sm_partition_create(SmPartition & sp, int device, unsigned int small_sms)
{
    sp = SmPartition();
    if (getenv("B2_NO_SM_PARTITION") || device < 0 || device >= 16) return false;
    CUresult (*p_stream)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    if (!entry("cuGreenCtxStreamCreate", p_stream)) return false;
    GreenPair * slot = nullptr;
    for (auto & g : g_pairs[device]) if (g.tried && g.req == small_sms) { slot = &g; break; }
    if (!slot) for (auto & g : g_pairs[device]) if (!g.tried) { slot = &g; break; }
    if (!slot) return false;                             // more distinct sizes than this table holds: no partition
    GreenPair & gp = *slot;
    if (!gp.tried) {
        // (re)build the pair of green contexts of this device; they live as long as the process
        gp = GreenPair();
        gp.tried = true; gp.req = small_sms;
        CUresult (*p_cuDeviceGet)(CUdevice *, int) = nullptr;
        CUresult (*p_getres)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
        CUresult (*p_split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
        CUresult (*p_desc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
        CUresult (*p_create)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
        if (!entry("cuDeviceGet", p_cuDeviceGet) || !entry("cuDeviceGetDevResource", p_getres) ||
            !entry("cuDevSmResourceSplitByCount", p_split) || !entry("cuDevResourceGenerateDesc", p_desc) ||
            !entry("cuGreenCtxCreate", p_create))
            return false;
        cudaFree(0);                                     // make sure the primary context exists
        CUdevice dev;
        if (p_cuDeviceGet(&dev, device) != CUDA_SUCCESS) return false;
        CUdevResource all, grp, rem;
        if (p_getres(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
        if (small_sms < 8 || small_sms + 8 > all.sm.smCount) return false;
        unsigned int nb = 1;
        if (p_split(&grp, &nb, &all, &rem, 0, small_sms) != CUDA_SUCCESS || nb != 1) return false;
        if (rem.sm.smCount == 0) return false;
        CUdevResourceDesc d_small, d_big;
        if (p_desc(&d_small, &grp, 1) != CUDA_SUCCESS || p_desc(&d_big, &rem, 1) != CUDA_SUCCESS) return false;
        if (p_create(&gp.g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
        if (p_create(&gp.g_big, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
        gp.small_sms = grp.sm.smCount; gp.big_sms = rem.sm.smCount;
        gp.ok = true;
        if (getenv("B2_VERBOSE")) fprintf(stderr, "b200ofdm: SM partition %u (synchronisers) + %u (channelizer, decode)\n", gp.small_sms, gp.big_sms);
    }
    if (!gp.ok) return false;
    CUstream s0 = nullptr, s3 = nullptr;
    if (p_stream(&s0, gp.g_small, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    if (p_stream(&s3, gp.g_small, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    for (unsigned int i = 0; i < SmPartition::NBIG; i++) {
        CUstream sb = nullptr;
        if (p_stream(&sb, gp.g_big, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
        sp.big_stream[i] = (cudaStream_t)sb;
    }
    sp.ok = true;
    sp.small_sms = gp.small_sms;
    sp.big_sms = gp.big_sms;
    sp.small_stream = (cudaStream_t)s0;
    sp.small_stream2 = (cudaStream_t)s3;
    return true;
}

} // namespace b2
