// capi_util.h -- error reporting and device-buffer helpers shared by the C ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <vector>
#include "b200_ofdm.h"

int b2_fail(int code, const char * fmt, ...);      // records the message (thread-local), returns code

#define B2_TRY(expr) do { int _rc = (expr); if (_rc != B2_OK) return _rc; } while (0)
#define B2_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    return b2_fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)

struct DevBuf {
    void * p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf & operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n)
    {
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) { p = nullptr; return b2_fail(B2_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e)); }
        bytes = n;
        return B2_OK;
    }
    template <typename T> int upload(const std::vector<T> & v)
    {
        int rc = alloc(v.size() * sizeof(T));
        if (rc) return rc;
        if (!v.empty()) {
            cudaError_t e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return b2_fail(B2_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
        }
        return B2_OK;
    }
    template <typename T> T * as() const { return (T *)p; }
};
