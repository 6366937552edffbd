// scan_i64.cuh -- exclusive prefix sum of int32 lengths into int64 offsets.
// cub::DeviceScan::ExclusiveSum takes its accumulator type from the INPUT type, so an int32 input wraps at 2^31 even when the
// output array is 64-bit (found with 25 M BAM records: 2.5e9 sequence bytes).  ExclusiveScan with a 64-bit initial value
// accumulates in 64 bits.
#pragma once
#include <cub/device/device_scan.cuh>

#include <stdint.h>

namespace exon {

struct SumI64 {
    __host__ __device__ __forceinline__ long long operator()(long long a, long long b) const { return a + b; }
};

// tmp == nullptr: size query (bytes out), as in cub
inline cudaError_t exclusive_sum_i32_i64(void *tmp, size_t &bytes, const int32_t *in, long long *out, int n, cudaStream_t st) {
    return cub::DeviceScan::ExclusiveScan(tmp, bytes, in, out, SumI64(), (long long)0, n, st);
}

}  // namespace exon
