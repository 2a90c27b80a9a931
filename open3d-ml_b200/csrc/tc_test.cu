// tc_test.cu -- bring-up / regression kernel for the tcgen05 primitives of tc.cuh:
// D[128, N] = A[128, K] * B[N, K]^T with the 3xFP16 split, one CTA, operands converted and
// laid out in shared memory by the CTA itself (exactly what the fused kernels do).
// Exposed through the C ABI (o3dml_tc_gemm_test) so that tests/test_gpu_tc.py can pin the
// descriptor encodings against torch.
#include "../../include/o3dml_b200_bringup.h"
#include "common.cuh"
#include "tc.cuh"

namespace o3dml {

__global__ void __launch_bounds__(128, 1)
tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N,
                    int K, int terms) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KC = K / 8;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + (size_t)KC * 128 * 16;
    uint8_t* b_hi = a_lo + (size_t)KC * 128 * 16;
    uint8_t* b_lo = b_hi + (size_t)KC * N * 16;
    for (int i = tid; i < KC * 128; i += 128) {
        const int r = i % 128, kc = i / 128;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = A[(size_t)r * K + kc * 8 + e];
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + tc::op_off(128, r, kc)) = hi;
        *reinterpret_cast<uint4*>(a_lo + tc::op_off(128, r, kc)) = lo;
    }
    for (int i = tid; i < KC * N; i += 128) {
        const int r = i % N, kc = i / N;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = B[(size_t)r * K + kc * 8 + e];
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        *reinterpret_cast<uint4*>(b_hi + tc::op_off(N, r, kc)) = hi;
        *reinterpret_cast<uint4*>(b_lo + tc::op_off(N, r, kc)) = lo;
    }
    tc::fence_async_smem();
    if (tid == 0) {
        tc::mbar_init(&mbar, 1);
        tc::fence_mbar_init();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16(128, N);
        const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16;
        const int variant = terms >> 8;  // bring-up only: 1 = LBO/SBO fields swapped
        terms &= 0xff;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t aL = variant ? 128u : a_lbo, aS = variant ? a_lbo : 128u;
            const uint32_t bL = variant ? 128u : b_lbo, bS = variant ? b_lbo : 128u;
            const uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi) + ks * 2 * a_lbo, aL, aS);
            const uint64_t al = tc::smem_desc(tc::smem_u32(a_lo) + ks * 2 * a_lbo, aL, aS);
            const uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 2 * b_lbo, bL, bS);
            const uint64_t bl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 2 * b_lbo, bL, bS);
            tc::umma_f16(taddr, ah, bh, idesc, ks > 0);
            if (terms >= 3) {
                tc::umma_f16(taddr, ah, bl, idesc, 1);
                tc::umma_f16(taddr, al, bh, idesc, 1);
            }
        }
        tc::umma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(taddr);
}


// Issue-rate probe: `reps` x 6 MMAs (M=128, N, K=16 each) on resident (zeroed) operands, one commit;
// out[0] = cycles from first issue to completion, out[1] = cycles spent issuing.
__global__ void __launch_bounds__(128, 1)
tc_mma_rate_kernel(int N, int reps, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_bytes = 4 * 128 * 16, b_bytes = 4 * N * 16;     // 32 channels each
    for (int i = tid; i < (2 * a_bytes + 2 * b_bytes) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    __syncthreads();
    if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16(128, N);
        const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16;
        const uint32_t ah0 = tc::smem_u32(smem), al0 = ah0 + a_bytes, bh0 = al0 + a_bytes, bl0 = bh0 + b_bytes;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ah = tc::smem_desc(ah0 + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t al = tc::smem_desc(al0 + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t bh = tc::smem_desc(bh0 + ks * 2 * b_lbo, b_lbo, 128);
                const uint64_t bl = tc::smem_desc(bl0 + ks * 2 * b_lbo, b_lbo, 128);
                tc::umma_f16(taddr, ah, bh, idesc, 1);
                tc::umma_f16(taddr, ah, bl, idesc, 1);
                tc::umma_f16(taddr, al, bh, idesc, 1);
            }
        }
        const long long t1 = clock64();
        tc::umma_commit(&mbar);
        tc::mbar_wait(&mbar, 0);
        const long long t2 = clock64();
        out[0] = t2 - t0;
        out[1] = t1 - t0;
    }
    __syncthreads();
    tc::tc_fence_before(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(taddr);
}

}  // namespace o3dml

using namespace o3dml;

extern "C" int o3dml_tc_mma_rate(int n, int reps, long long* out, void* stream) {
    O3DML_CHECK(n >= 16 && n <= 256 && (n % 16) == 0, "tc_mma_rate: bad N");
    const size_t smem = 2 * 4 * 128 * 16 + 2 * 4 * (size_t)n * 16;
    O3DML_CUDA(cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n, reps, out);
    O3DML_LAUNCH_CHECK();
    return O3DML_OK;
}


extern "C" int o3dml_tc_gemm_test(const float* a, const float* b, float* d, int n, int k, int terms,
                                  void* stream) {
    O3DML_CHECK(n >= 16 && n <= 256 && (n % 16) == 0, "tc_gemm_test: N must be a multiple of 16 in [16,256]");
    O3DML_CHECK(k >= 16 && (k % 16) == 0, "tc_gemm_test: K must be a multiple of 16");
    const size_t smem = (size_t)(k / 8) * 16 * 2 * (128 + n);
    O3DML_CHECK(smem <= 200 * 1024, "tc_gemm_test: operands do not fit in shared memory");
    O3DML_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b, d, n, k, terms);
    O3DML_LAUNCH_CHECK();
    o3dml_count_launches(1);
    return O3DML_OK;
}
