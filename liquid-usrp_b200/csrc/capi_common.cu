// capi_common.cu -- error text, version and device discovery for the C ABI (include/b200_ofdm.h).
#include <cstdarg>
#include <cstdio>
#include "capi_util.h"

static thread_local char g_err[512] = "";

int b2_fail(int code, const char * fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char * b2_last_error(void) { return g_err; }
extern "C" const char * b2_version(void) { return "b200ofdm 0.1 (sm_100a)"; }
extern "C" int b2_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void * b2_pinned_alloc(size_t bytes)
{
    void * p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void b2_pinned_free(void * p)
{
    if (p) cudaFreeHost(p);
}
