// fec.cuh -- small integer codecs shared by the synchroniser (frame header) and the packet
// decoder: Golay(24,12) and bytewise CRC-32, as liquid's fec_golay2412.c / crc.c.
#pragma once
#include <stdint.h>

namespace b2 {

__device__ __forceinline__ unsigned int golay_mul_P(unsigned int v)
{
    const unsigned int P[12] = {0x08ed, 0x01db, 0x03b5, 0x0769, 0x0ed1, 0x0da3, 0x0b47, 0x068f, 0x0d1d, 0x0a3b, 0x0477, 0x0ffe};
    unsigned int x = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) x = (x << 1) | (__popc(P[i] & v) & 1);
    return x;
}
__device__ __forceinline__ int golay_search(unsigned int v)
{
    const unsigned int P[12] = {0x08ed, 0x01db, 0x03b5, 0x0769, 0x0ed1, 0x0da3, 0x0b47, 0x068f, 0x0d1d, 0x0a3b, 0x0477, 0x0ffe};
    for (int i = 0; i < 12; i++)
        if (__popc(v ^ P[i]) <= 2) return i;
    return -1;
}
__device__ __forceinline__ unsigned int golay2412_decode(unsigned int r)
{
    const unsigned int P[12] = {0x08ed, 0x01db, 0x03b5, 0x0769, 0x0ed1, 0x0da3, 0x0b47, 0x068f, 0x0d1d, 0x0a3b, 0x0477, 0x0ffe};
    unsigned int s = ((r >> 12) & 0xfff) ^ golay_mul_P(r & 0xfff);
    unsigned int e = 0;
    if (__popc(s) <= 3) {
        e = (s << 12) & 0xfff000;
    } else {
        int si = golay_search(s);
        if (si >= 0) {
            e = ((s ^ P[si]) << 12) | (1u << (11 - si));
        } else {
            unsigned int sP = golay_mul_P(s);
            if (__popc(sP) <= 3) {
                e = sP;
            } else {
                int pi = golay_search(sP);
                if (pi >= 0) e = (1u << (11 - pi + 12)) | (sP ^ P[pi]);
            }
        }
    }
    return (r ^ e) & 0x0fff;
}
__device__ __forceinline__ uint32_t crc32_bytes(const uint8_t * m, unsigned int n)
{
    uint32_t key = ~0u;
    for (unsigned int i = 0; i < n; i++) {
        key ^= m[i];
#pragma unroll
        for (int j = 0; j < 8; j++) key = (key >> 1) ^ (0xEDB88320u & (0u - (key & 1u)));
    }
    return ~key;
}


} // namespace b2
