// cub_baseline.cu -- LIBRARY BASELINE for the sort micro-benchmark (benchmarks/sort_microbench.py), not product code:
// cub::DeviceRadixSort::SortPairs<uint64_t, uint32_t> from the CUDA toolkit, on the same box and the same keys as
// splatapult_b200's own onesweep sorter (SURVEY.md 8d "Sort micro-bench").
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <cstdint>

extern "C" __attribute__((visibility("default"))) size_t cub_sort_temp_bytes(uint64_t n, int begin_bit, int end_bit)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int64_t)n, begin_bit, end_bit);
    return bytes;
}

extern "C" __attribute__((visibility("default"))) int cub_sort_pairs(void* temp, size_t temp_bytes, const void* keys_in, void* keys_out,
                                                                      const void* vals_in, void* vals_out, uint64_t n, int begin_bit, int end_bit,
                                                                      void* stream)
{
    return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const unsigned long long*)keys_in, (unsigned long long*)keys_out,
                                                (const uint32_t*)vals_in, (uint32_t*)vals_out, (int64_t)n, begin_bit, end_bit, (cudaStream_t)stream);
}
