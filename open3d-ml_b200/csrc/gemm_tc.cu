// gemm_tc.cu -- the gathered GEMM / implicit-GEMM convolution of gemm.cu on the tcgen05 tensor
// cores.  Same operand model (A rows gathered on the fly: identity / index gather with shadow rows /
// batch-relative index / 3x3 conv taps; epilogue = folded BN + residual + activation, row-major /
// NCHW / pixel-shuffle output), but
//   * A k-slices (32 channels) are converted by the CTA to fp16 hi/lo pairs and written straight into
//     the UMMA chunk-major shared-memory layout (tc.cuh), 3 MMAs per k-step (3xFP16, ~2^-21 relative),
//   * B comes from a host-packed hi/lo operand image ([K/8][Cout_pad][8 halves], zero padded),
//   * the accumulator [128 x BN] lives in TMEM; a 3-stage ring lets the conversion of slice s+1
//     overlap the MMAs of slice s (global loads of slice s+1 are in flight across the barrier).
// fp16 has a narrow exponent range (the lo parts go subnormal for |x| < 0.125 and precision
// decays to 1e-4 for |x| ~ 1e-4), so every CTA first takes the max |A| of its own tile (one extra
// pass over data that is read again right after, i.e. L2 hits) and rescales by an exact power of two
// to [2^13, 2^14); the weight image is normalised the same way by the host; the epilogue undoes both.
// Serves SharedMLP / UnaryBlock / KPConv [15*Cin, Cout] / SECOND + FPN + head convolutions whenever
// every source has a multiple of 8 channels; the FP32 SIMT kernel (gemm.cu) takes the rest.
#include "../../include/o3dml_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace o3dml {

constexpr int GT_THREADS = 256;
constexpr int GT_ROWS = 128;
constexpr int GT_KS = 32;       // channels per k-slice
constexpr int GT_CH = GT_KS / 8;
constexpr int GT_STAGES = 5;
constexpr int GT_FLUSH = 8;      // k-slices (256 channels) per TMEM accumulation chunk
constexpr int GT_MAX_SRC = 3;

struct GtSrc {
    const float* data;
    const void* index;
    int64_t rows, out_rows_per_batch, src_rows_per_batch;
    int32_t channels, ld, index_is64, index_ld;
};

struct GemmTcParams {
    int64_t N;
    int K, Kpad, Cout, Npad;
    int mode;  // 0 rows, 1 conv3x3
    int nsrc;
    GtSrc src[GT_MAX_SRC];
    int koff[GT_MAX_SRC + 1];
    int H, W, OH, OW, stride, C;
    const uint4* wimg;  // hi image then lo image, each [Kpad/8][Npad] uint4
    int wexp;           // the image holds weight * 2^wexp (host-side range normalisation)
    const float* scale;
    const float* shift;
    const float* residual;
    int res_ld;
    int act;
    float slope;
    float* out;
    int out_ld;
    int out_mode;   // 0 rows, 1 NCHW, 2 deconv pixel shuffle
    int64_t plane;
    int ds, dIH, dIW, dC;
};

__device__ __forceinline__ const float* gt_src_ptr(const GemmTcParams& p, int s, int64_t n) {
    const GtSrc& S = p.src[s];
    int64_t r = n;
    if (S.index) {
        r = load_index(S.index, n * S.index_ld, S.index_is64);
        if (r < 0) return nullptr;
        if (S.out_rows_per_batch > 0) {
            if (r >= S.src_rows_per_batch) return nullptr;
            r += (n / S.out_rows_per_batch) * S.src_rows_per_batch;
        }
        if (r >= S.rows) return nullptr;
    }
    return S.data + (size_t)r * S.ld;
}

// ---- cp.async (16-byte, zero-fill when src_bytes == 0) -----------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// the mbarrier gets one arrival from this thread once all of its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN>
struct GtCfg {
    // chunk stride of the A operand padded by 16 B (UMMA LBO is free): lanes that differ in the chunk
    // index hit different banks when they convert their slots in place
    static constexpr int A_LBO = GT_ROWS * 16 + 16;
    static constexpr int A_BYTES = GT_CH * A_LBO;         // one of hi/lo per stage
    static constexpr int B_BYTES = GT_CH * BN * 16;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int B_U4 = 2 * GT_CH * BN;                     // uint4 of B per slice (hi + lo)
    static constexpr int B_PER_THREAD = (B_U4 + GT_THREADS - 1) / GT_THREADS;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulator buffers
    static constexpr size_t SMEM = (size_t)GT_STAGES * STAGE + GT_MAX_SRC * GT_ROWS * 8 + GT_ROWS * 12 + 256;
};

// Pipeline (per CTA, one [128 x BN] output tile, k-slices of 32 channels through a GT_STAGES ring):
//   cp.async   raw fp32 A pieces are parked IN the slots where their fp16 hi / lo words will live
//              (floats 0-3 of a (row, 8-channel chunk) in the hi slot, floats 4-7 in the lo slot), the
//              weight image slices go straight to their final place; GT_STAGES-1 slices are in flight
//   convert    each thread rewrites its own two 16-byte slots in place: x*2^e -> (hi, lo) halves
//   MMA        one thread issues 3 x 2 tcgen05.mma per slice; its commit frees the stage
//   flush      every GT_FLUSH slices the TMEM accumulator is added (RN) into registers and the next
//              chunk starts a fresh accumulator in the other TMEM buffer: the tensor core adds with
//              truncation (measured -3e-8 relative per accumulation, -5e-5 at K = 7680 otherwise)
template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
    using C = GtCfg<BN>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stages = smem;
    const float** rowptr = reinterpret_cast<const float**>(stages + GT_STAGES * C::STAGE);  // [src][row]
    int* rowinfo = reinterpret_cast<int*>(rowptr + GT_MAX_SRC * GT_ROWS);                   // [row][3]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(rowinfo + GT_ROWS * 3);   // full[S], empty[S], chunk[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * GT_STAGES + 2);
    __shared__ unsigned amax_warp[GT_THREADS / 32];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * GT_ROWS;
    const int col0 = blockIdx.y * BN;

    // ---- per-row gather bookkeeping
    if (p.mode == 0) {
        for (int i = tid; i < p.nsrc * GT_ROWS; i += GT_THREADS) {
            const int s = i / GT_ROWS, m = i % GT_ROWS;
            const int64_t n = row0 + m;
            rowptr[s * GT_ROWS + m] = (n < p.N) ? gt_src_ptr(p, s, n) : nullptr;
        }
    } else {
        for (int m = tid; m < GT_ROWS; m += GT_THREADS) {
            const int64_t n = row0 + m;
            if (n < p.N) {
                const int64_t per = (int64_t)p.OH * p.OW;
                const int b = (int)(n / per), r = (int)(n % per);
                rowinfo[m * 3 + 0] = b * p.H * p.W;
                rowinfo[m * 3 + 1] = (r / p.OW) * p.stride - 1;
                rowinfo[m * 3 + 2] = (r % p.OW) * p.stride - 1;
            } else {
                rowinfo[m * 3 + 0] = -1;
                rowinfo[m * 3 + 1] = rowinfo[m * 3 + 2] = 0;
            }
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < GT_STAGES; ++i) tc::mbar_init(&mbar[i], GT_THREADS / 2);        // full: loaders
#pragma unroll
        for (int i = GT_STAGES; i < 2 * GT_STAGES + 2; ++i) tc::mbar_init(&mbar[i], 1);       // empty, chunk
        tc::fence_mbar_init();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int nsl = p.Kpad / GT_KS;
    const size_t img_u4 = (size_t)(p.Kpad / 8) * p.Npad;  // uint4 per image

    // ---- per-thread gather state.  The two (row, 8-channel chunk) items of a thread are the same in
    // every slice, so everything that does not depend on the slice index is resolved here once.
    int it_m[2], it_c[2];
    const float* it_base[2];      // conv: pixel (iy0, ix0) of the row, rows mode with one source: the row
    unsigned it_taps[2];          // conv: bit t set when tap t lies inside the image
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int item = tid + it * GT_THREADS;
        it_m[it] = item >> 2;          // 4 consecutive lanes read the 4 x 32 B of one row's 128-byte slice:
        it_c[it] = item & 3;           // 8 cache lines per warp request instead of 32
        it_base[it] = nullptr;
        it_taps[it] = 0;
        const int m = it_m[it];
        if (p.mode == 1) {
            if (rowinfo[m * 3] >= 0) {
                const int iy0 = rowinfo[m * 3 + 1], ix0 = rowinfo[m * 3 + 2];
                it_base[it] = p.src[0].data + ((int64_t)rowinfo[m * 3] + (int64_t)iy0 * p.W + ix0) * p.C;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int iy = iy0 + t / 3, ix = ix0 + t % 3;
                    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) it_taps[it] |= 1u << t;
                }
            }
        } else if (p.nsrc == 1) {
            it_base[it] = rowptr[m];
        }
    }
    const bool conv_fast = p.mode == 1 && (p.C % GT_KS) == 0;   // a 32-channel slice never straddles taps
    const int slices_per_tap = conv_fast ? p.C / GT_KS : 1;
    auto a_src = [&](int s, int it) -> const float* {
        const int m = it_m[it], c = it_c[it];
        const int k = s * GT_KS + c * 8;
        if (k >= p.K) return nullptr;
        if (p.mode == 1) {
            int tap, cc;
            if (conv_fast) {
                tap = s / slices_per_tap;
                cc = (s - tap * slices_per_tap) * GT_KS + c * 8;
            } else {
                tap = k / p.C;
                cc = k - tap * p.C;
            }
            if (!((it_taps[it] >> tap) & 1u)) return nullptr;
            return it_base[it] + ((int64_t)(tap / 3) * p.W + (tap % 3)) * p.C + cc;
        }
        if (p.nsrc == 1) return it_base[it] ? it_base[it] + k : nullptr;
        int sidx = 0;
        while (sidx + 1 < p.nsrc && k >= p.koff[sidx + 1]) ++sidx;
        const float* base = rowptr[sidx * GT_ROWS + m];
        return base ? base + (k - p.koff[sidx]) : nullptr;
    };

    float a_scale, out_scale;
    {   // ---- range pass: max |A| over this CTA's tile -> exact power-of-two scale
        float mx4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s0 = 0; s0 < nsl; s0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {       // four slices in flight: independent loads and maxima
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const float* src = (s0 + u < nsl) ? a_src(s0 + u, it) : nullptr;
                    if (src) {
                        const float4 v0 = *reinterpret_cast<const float4*>(src);
                        const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
                        mx4[u] = fmaxf(mx4[u], fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
                        mx4[u] = fmaxf(mx4[u], fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
                    }
                }
            }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const unsigned wmx = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));  // mx >= 0: bits are ordered
        if ((tid & 31) == 0) amax_warp[warp] = wmx;
        __syncthreads();
        unsigned bm = 0;
#pragma unroll
        for (int i = 0; i < GT_THREADS / 32; ++i) bm = max(bm, amax_warp[i]);
        const float amax = __uint_as_float(bm);
        int e = 0;
        if (amax > 0.f && amax < 3.0e38f) e = 13 - ilogbf(amax);   // amax * 2^e in [2^13, 2^14)
        e = max(-100, min(100, e));
        a_scale = ldexpf(1.f, e);
        out_scale = ldexpf(1.f, -e - p.wexp);
    }

    // ---- warp-specialised pipeline.  Warps 0-3 only LOAD (cp.async, completion signalled through
    // full[stage] by cp.async.mbarrier.arrive); warps 4-7 CONVERT in place, fence, and one of their
    // threads issues the MMAs whose commit frees the stage (empty[stage]).  Keeping the two roles in
    // different threads matters: fence.proxy.async drains the issuing thread's outstanding cp.async,
    // so a thread that both prefetches and fences never has a load in flight across the fence
    // (measured: 2.5 us per k-slice regardless of the prefetch depth).
    uint64_t* full_bar = mbar;                       // [GT_STAGES], 128 loader arrivals
    uint64_t* empty_bar = mbar + GT_STAGES;          // [GT_STAGES], tcgen05.commit
    uint64_t* chunk_bar = mbar + 2 * GT_STAGES;      // [2]
    constexpr int HALF = GT_THREADS / 2;
    const bool is_loader = tid < HALF;
    const int rt = tid & (HALF - 1);                 // thread index inside its role
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float racc[BN];                                  // converters: RN-accumulated chunk sums of their row
#pragma unroll
    for (int i = 0; i < BN; ++i) racc[i] = 0.f;
    const int last_chunk = (nsl - 1) / GT_FLUSH;

    if (is_loader) {
        // ------------------------------------------------------------------ loaders
        int lm[4], lc[4];
        const float* lbase[4];
        unsigned ltaps[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                // 512 (row, chunk) items, 4 per loader thread
            const int item = rt + j * HALF;
            lm[j] = item >> 2;
            lc[j] = item & 3;
            lbase[j] = nullptr;
            ltaps[j] = 0;
            const int m = lm[j];
            if (p.mode == 1) {
                if (rowinfo[m * 3] >= 0) {
                    const int iy0 = rowinfo[m * 3 + 1], ix0 = rowinfo[m * 3 + 2];
                    lbase[j] = p.src[0].data + ((int64_t)rowinfo[m * 3] + (int64_t)iy0 * p.W + ix0) * p.C;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int iy = iy0 + t / 3, ix = ix0 + t % 3;
                        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) ltaps[j] |= 1u << t;
                    }
                }
            } else if (p.nsrc == 1) {
                lbase[j] = rowptr[m];
            }
        }
        constexpr int LB = (C::B_U4 + HALF - 1) / HALF;   // weight-image words per loader thread
        const uint4* b_src[LB];
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int idx = rt + j * HALF;
            const int img = idx / (GT_CH * BN), rem = idx % (GT_CH * BN);
            b_src[j] = p.wimg + img * img_u4 + (size_t)(rem / BN) * p.Npad + col0 + (rem % BN);
        }
        const size_t b_step = (size_t)GT_CH * p.Npad;
        for (int s = 0; s < nsl; ++s) {
            const int stage = s % GT_STAGES, use = s / GT_STAGES;
            if (use >= 1) tc::mbar_wait(&empty_bar[stage], (use - 1) & 1);   // MMAs of slice s-STAGES done
            uint8_t* a_hi = stages + (size_t)stage * C::STAGE;
            uint8_t* a_lo = a_hi + C::A_BYTES;
            uint4* bdst = reinterpret_cast<uint4*>(a_lo + C::A_BYTES);
            int tap = 0, cc0 = s * GT_KS;
            if (p.mode == 1) {
                if (conv_fast) {
                    tap = s / slices_per_tap;
                    cc0 = (s - tap * slices_per_tap) * GT_KS;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = s * GT_KS + lc[j] * 8;
                const float* src = nullptr;
                if (k < p.K) {
                    if (p.mode == 1) {
                        int t = tap, cc = cc0 + lc[j] * 8;
                        if (!conv_fast) {
                            t = k / p.C;
                            cc = k - t * p.C;
                        }
                        if ((ltaps[j] >> t) & 1u) src = lbase[j] + ((int64_t)(t / 3) * p.W + (t % 3)) * p.C + cc;
                    } else if (p.nsrc == 1) {
                        if (lbase[j]) src = lbase[j] + k;
                    } else {
                        int sidx = 0;
                        while (sidx + 1 < p.nsrc && k >= p.koff[sidx + 1]) ++sidx;
                        const float* base = rowptr[sidx * GT_ROWS + lm[j]];
                        if (base) src = base + (k - p.koff[sidx]);
                    }
                }
                const void* g0 = src ? (const void*)src : (const void*)p.wimg;
                const void* g1 = src ? (const void*)(src + 4) : (const void*)p.wimg;
                const uint32_t off = (uint32_t)lc[j] * C::A_LBO + (uint32_t)lm[j] * 16u;
                cp_async16(a_hi + off, g0, src ? 16 : 0);
                cp_async16(a_lo + off, g1, src ? 16 : 0);
            }
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int idx = rt + j * HALF;
                if (idx < C::B_U4) cp_async16(&bdst[idx], b_src[j] + (size_t)s * b_step, 16);
            }
            cp_async_arrive(&full_bar[stage]);
        }
        cp_async_wait<0>();
    } else {
        // ---------------------------------------------------------------- converters + MMA + flush
        auto flush = [&](int buf) {       // racc += TMEM accumulator buffer `buf` (round to nearest)
#pragma unroll
            for (int q = 0; q < BN / 16; ++q) {
                float v[16];
                tc::tmem_ld16(tmem_lane + buf * BN + q * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) racc[q * 16 + j] += v[j];
            }
        };
        uint32_t ph_chunk[2] = {0, 0};
        int pending_chunk = -1;
        for (int s = 0; s < nsl; ++s) {
            const int stage = s % GT_STAGES, use = s / GT_STAGES;
            const int chunk = s / GT_FLUSH;
            tc::mbar_wait(&full_bar[stage], use & 1);        // every piece of slice s has landed
            uint8_t* a_hi = stages + (size_t)stage * C::STAGE;
            uint8_t* a_lo = a_hi + C::A_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {                    // 512 items, 4 per converter thread, in place
                const int item = rt + j * HALF;
                const uint32_t off = (uint32_t)(item & 3) * C::A_LBO + (uint32_t)(item >> 2) * 16u;
                uint4* ph = reinterpret_cast<uint4*>(a_hi + off);
                uint4* pl = reinterpret_cast<uint4*>(a_lo + off);
                const float4 v0 = *reinterpret_cast<const float4*>(ph);
                const float4 v1 = *reinterpret_cast<const float4*>(pl);
                const float x[8] = {v0.x * a_scale, v0.y * a_scale, v0.z * a_scale, v0.w * a_scale,
                                    v1.x * a_scale, v1.y * a_scale, v1.z * a_scale, v1.w * a_scale};
                uint4 hi, lo;
                tc::split8(x, hi, lo);
                *ph = hi;
                *pl = lo;
            }
            tc::fence_async_smem();
            tc::tc_fence_before();
            named_bar_sync(1, HALF);                          // all 128 converters
            tc::tc_fence_after();
            if (rt == 0) {
                constexpr uint32_t idesc = tc::idesc_f16(GT_ROWS, BN);
                constexpr uint32_t A_LBO = C::A_LBO, B_LBO = BN * 16;
                const uint32_t ah0 = tc::smem_u32(a_hi);
                const uint32_t al0 = ah0 + C::A_BYTES;
                const uint32_t bh0 = al0 + C::A_BYTES;
                const uint32_t bl0 = bh0 + C::B_BYTES;
                const uint32_t acc = tmem + (uint32_t)((chunk & 1) * BN);
                const bool first = (s % GT_FLUSH) == 0;
#pragma unroll
                for (int ks = 0; ks < GT_KS / 16; ++ks) {
                    const uint64_t ah = tc::smem_desc(ah0 + ks * 2 * A_LBO, A_LBO, 128);
                    const uint64_t al = tc::smem_desc(al0 + ks * 2 * A_LBO, A_LBO, 128);
                    const uint64_t bh = tc::smem_desc(bh0 + ks * 2 * B_LBO, B_LBO, 128);
                    const uint64_t bl = tc::smem_desc(bl0 + ks * 2 * B_LBO, B_LBO, 128);
                    tc::umma_f16(acc, ah, bh, idesc, !(first && ks == 0));
                    tc::umma_f16(acc, ah, bl, idesc, 1);
                    tc::umma_f16(acc, al, bh, idesc, 1);
                }
                tc::umma_commit(&empty_bar[stage]);
                if ((s % GT_FLUSH) == GT_FLUSH - 1 || s == nsl - 1) tc::umma_commit(&chunk_bar[chunk & 1]);
            }
            if (pending_chunk >= 0) {     // the chunk that ended one slice ago has drained by now
                tc::mbar_wait(&chunk_bar[pending_chunk & 1], ph_chunk[pending_chunk & 1]);
                ph_chunk[pending_chunk & 1] ^= 1;
                tc::tc_fence_after();
                flush(pending_chunk & 1);
                pending_chunk = -1;
            }
            if ((s % GT_FLUSH) == GT_FLUSH - 1 && s != nsl - 1) pending_chunk = chunk;
        }
        tc::mbar_wait(&chunk_bar[last_chunk & 1], ph_chunk[last_chunk & 1]);
        tc::tc_fence_after();
    }

    // ---- epilogue: converter thread = output row; 16 columns at a time
    if (!is_loader) {
    const int row = rt;
    const int64_t n = row0 + row;
#pragma unroll
    for (int q = 0; q < BN / 16; ++q) {
        const int c0 = 16 * q;
        float v[16];
        tc::tmem_ld16(tmem_lane + (last_chunk & 1) * BN + c0, v);   // warp-collective: every lane takes part
        const int cbase = col0 + c0;
        if (n >= p.N || cbase >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = cbase + j;
            if (c < p.Cout) {
                float x = (v[j] + racc[q * 16 + j]) * out_scale;   // exact: undoes the power-of-two scalings
                x = fmaf(x, p.scale ? p.scale[c] : 1.f, p.shift ? p.shift[c] : 0.f);
                if (p.residual) x += p.residual[(size_t)n * p.res_ld + c];
                v[j] = apply_act(x, p.act, p.slope);
            }
        }
        if (p.out_mode == 0) {
            float* o = p.out + (size_t)n * p.out_ld + cbase;
            if (cbase + 15 < p.Cout && (p.out_ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<float4*>(o + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (cbase + j < p.Cout) o[j] = v[j];
            }
        } else if (p.out_mode == 1) {
            const int64_t b = n / p.plane, pix = n % p.plane;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (cbase + j < p.Cout) p.out[((size_t)b * p.Cout + cbase + j) * p.plane + pix] = v[j];
        } else {
            const int64_t per = (int64_t)p.dIH * p.dIW;
            const int64_t b = n / per;
            const int r = (int)(n % per);
            const int iy = r / p.dIW, ix = r % p.dIW;
            const int OWd = p.dIW * p.ds;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = cbase + j;
                if (c < p.Cout) {
                    const int sub = c / p.dC, co = c - sub * p.dC;
                    const int dy = sub / p.ds, dx = sub - dy * p.ds;
                    const size_t opix = ((size_t)b * p.dIH * p.ds + (size_t)iy * p.ds + dy) * OWd +
                                        (size_t)ix * p.ds + dx;
                    p.out[opix * p.out_ld + co] = v[j];
                }
            }
        }
    }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int BN>
static int gemm_tc_launch_bn(const GemmTcParams& p, cudaStream_t st) {
    using C = GtCfg<BN>;
    static bool configured = false;
    if (!configured) {
        O3DML_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C::SMEM));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div<int64_t>(p.N, GT_ROWS), (unsigned)(p.Npad / BN));
    gemm_tc_kernel<BN><<<grid, GT_THREADS, C::SMEM, st>>>(p);
    O3DML_LAUNCH_CHECK();
    o3dml_count_launches(1);
    return O3DML_OK;
}

static int gemm_tc_launch(const GemmTcParams& p, cudaStream_t st) {
    if (p.N <= 0 || p.Cout <= 0) return O3DML_OK;
    O3DML_CHECK(p.Kpad % GT_KS == 0 && p.Kpad >= p.K, "linear_tc: weight image K padding must be a multiple of 32");
    if (p.Npad == 32) return gemm_tc_launch_bn<32>(p, st);
    if (p.Npad == 64) return gemm_tc_launch_bn<64>(p, st);
    O3DML_CHECK(p.Npad % 128 == 0, "linear_tc: weight image rows must be padded to 32, 64 or a multiple of 128");
    return gemm_tc_launch_bn<128>(p, st);
}

}  // namespace o3dml

using namespace o3dml;

static int gt_common(GemmTcParams& p, const void* wimg, int k_pad, int n_pad, int w_exp, const float* scale,
                     const float* shift, const float* residual, int residual_ld, int act, float slope,
                     float* out, int out_ld, int out_channels) {
    p.wimg = (const uint4*)wimg;
    p.Kpad = k_pad;
    p.Npad = n_pad;
    p.wexp = w_exp;
    p.scale = scale; p.shift = shift; p.residual = residual; p.res_ld = residual_ld;
    p.act = act; p.slope = slope; p.out = out; p.out_ld = out_ld; p.Cout = out_channels;
    O3DML_CHECK(act >= 0 && act <= 2, "linear_tc: unknown activation %d", act);
    O3DML_CHECK(wimg && out, "linear_tc: null weight image / out");
    O3DML_CHECK(n_pad >= out_channels, "linear_tc: weight image has fewer rows than out_channels");
    return O3DML_OK;
}

extern "C" int o3dml_linear_tc(int64_t num_rows, const o3dml_src_t* srcs, int num_srcs,
                               const void* weight_image, int k_pad, int n_pad, int weight_exp,
                               const float* scale,
                               const float* shift, const float* residual, int residual_ld, int act,
                               float slope, float* out, int out_ld, int out_channels,
                               int out_nchw_plane, void* stream) {
    O3DML_CHECK(num_srcs >= 1 && num_srcs <= GT_MAX_SRC, "linear_tc: 1..3 sources");
    GemmTcParams p = {};
    p.N = num_rows;
    p.mode = 0;
    p.nsrc = num_srcs;
    int k = 0;
    for (int s = 0; s < num_srcs; ++s) {
        const o3dml_src_t& S = srcs[s];
        O3DML_CHECK(S.data && S.channels > 0 && S.ld >= S.channels, "linear_tc: bad source %d", s);
        O3DML_CHECK((S.channels & 7) == 0 && (S.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(S.data) & 15) == 0,
                    "linear_tc: sources need a multiple of 8 channels and 16-byte aligned rows");
        p.src[s].data = S.data; p.src[s].index = S.index; p.src[s].rows = S.rows;
        p.src[s].out_rows_per_batch = S.out_rows_per_batch;
        p.src[s].src_rows_per_batch = S.src_rows_per_batch;
        p.src[s].channels = S.channels; p.src[s].ld = S.ld; p.src[s].index_is64 = S.index_is64;
        p.src[s].index_ld = S.index ? (S.index_ld > 0 ? S.index_ld : 1) : 0;
        p.koff[s] = k;
        k += S.channels;
    }
    for (int s = num_srcs; s <= GT_MAX_SRC; ++s) p.koff[s] = k;
    p.K = k;
    int rc = gt_common(p, weight_image, k_pad, n_pad, weight_exp, scale, shift, residual, residual_ld, act, slope, out,
                       out_ld, out_channels);
    if (rc) return rc;
    if (out_nchw_plane > 0) {
        p.out_mode = 1;
        p.plane = out_nchw_plane;
    }
    return gemm_tc_launch(p, (cudaStream_t)stream);
}

extern "C" int o3dml_conv3x3_nhwc_tc(const float* in, int batch, int H, int W, int C, int stride,
                                     const void* weight_image, int k_pad, int n_pad, int weight_exp,
                               const float* scale,
                                     const float* shift, int act, float slope, float* out,
                                     int out_channels, void* stream) {
    O3DML_CHECK(in && batch > 0 && H > 0 && W > 0, "conv3x3_tc: bad input");
    O3DML_CHECK((C % 8) == 0, "conv3x3_tc: input channels must be a multiple of 8");
    O3DML_CHECK(stride == 1 || stride == 2, "conv3x3_tc: stride 1 or 2");
    O3DML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv3x3_tc: input must be 16-byte aligned");
    GemmTcParams p = {};
    p.mode = 1;
    p.nsrc = 1;
    p.src[0].data = in;
    p.H = H; p.W = W; p.C = C; p.stride = stride;
    p.OH = (H + 2 - 3) / stride + 1;
    p.OW = (W + 2 - 3) / stride + 1;
    p.N = (int64_t)batch * p.OH * p.OW;
    p.K = 9 * C;
    int rc = gt_common(p, weight_image, k_pad, n_pad, weight_exp, scale, shift, nullptr, 0, act, slope, out,
                       out_channels, out_channels);
    if (rc) return rc;
    return gemm_tc_launch(p, (cudaStream_t)stream);
}

extern "C" int o3dml_deconv_nhwc_tc(const float* in, int batch, int H, int W, int C, int stride,
                                    const void* weight_image, int k_pad, int n_pad, int weight_exp,
                               const float* scale,
                                    const float* shift, int act, float slope, float* out, int out_ld,
                                    int out_channels, void* stream) {
    O3DML_CHECK(in && batch > 0 && H > 0 && W > 0 && stride >= 1, "deconv_tc: bad input");
    O3DML_CHECK((C % 8) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0, "deconv_tc: C % 8, aligned input");
    GemmTcParams p = {};
    p.N = (int64_t)batch * H * W;
    p.mode = 0;
    p.nsrc = 1;
    p.src[0].data = in; p.src[0].rows = p.N; p.src[0].channels = C; p.src[0].ld = C;
    p.koff[0] = 0;
    for (int i = 1; i <= GT_MAX_SRC; ++i) p.koff[i] = C;
    p.K = C;
    int rc = gt_common(p, weight_image, k_pad, n_pad, weight_exp, scale, shift, nullptr, 0, act, slope, out, out_ld,
                       stride * stride * out_channels);
    if (rc) return rc;
    p.out_mode = 2;
    p.ds = stride; p.dIH = H; p.dIW = W; p.dC = out_channels;
    return gemm_tc_launch(p, (cudaStream_t)stream);
}
