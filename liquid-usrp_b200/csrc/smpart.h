// smpart.h -- SM partition (green contexts) for the receive pipeline, see smpart.cu
#pragma once
#include <cuda_runtime.h>

namespace b2 {

struct SmPartition {
    bool ok = false;
    unsigned int small_sms = 0, big_sms = 0;
    cudaStream_t small_stream = nullptr;         // synchroniser chains
    cudaStream_t small_stream2 = nullptr;        // a second stream on the same SM set
    static const unsigned int NBIG = 7;
    cudaStream_t big_stream[NBIG] = {};          // channelizer, 6 x packet decode
};
// streams of a (cached, per device) partition with `small_sms` SMs in the small set; false when the
// driver cannot partition (the caller then uses ordinary streams).  The caller destroys the streams.
bool sm_partition_create(SmPartition & sp, int device, unsigned int small_sms);

} // namespace b2
