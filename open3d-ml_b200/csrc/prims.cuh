// prims.cuh -- internal device-wide primitives (scan, radix sort).
#pragma once
#include "common.cuh"

namespace o3dml {

size_t scan_temp_bytes(int64_t n);
cudaError_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out,
                               void* temp, cudaStream_t st);

size_t radix_sort_temp_bytes(int64_t n);
cudaError_t radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                             bool vals_are_iota, int64_t n, int num_bits, void* temp,
                             cudaStream_t st, int* result_in_b);

}  // namespace o3dml
