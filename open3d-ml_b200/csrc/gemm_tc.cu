// gemm_tc.cu -- the gathered GEMM / implicit-GEMM convolution of gemm.cu on the tcgen05 tensor
// cores.  Same operand model (A rows gathered on the fly: identity / index gather with shadow rows /
// batch-relative index / 3x3 conv taps; epilogue = folded BN + residual + activation, row-major /
// NCHW / pixel-shuffle output), but
//   * A k-slices (32 channels) are converted by the CTA to fp16 hi/lo pairs and written straight into
//     the UMMA chunk-major shared-memory layout (tc.cuh), 3 MMAs per k-step (3xFP16, ~2^-21 relative),
//   * B comes from a host-packed hi/lo operand image ([K/8][Cout_pad][8 halves], zero padded),
//   * the accumulator [128 x BN] lives in TMEM (two column ranges, chunked accumulation, see "flush"
//     below); a GT_STAGES-deep ring of (A, B) slices decouples the loader, converter and MMA warps.
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

#ifdef O3DML_DEBUG_TIMING
__device__ long long g_gt_dbg[8192];
#endif

constexpr int GT_THREADS = 416;   // warp 0 MMA, warps 1-4 loaders, warps 5-8 / 9-12 converter groups
constexpr int GT_LOADERS = 128;
constexpr int GT_CONV = 128;
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
// 1-D bulk async copy (TMA engine) global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN>
struct GtCfg {
    // chunk stride of the A operand padded by 32 B (UMMA LBO is free).  A quarter-warp of the loaders /
    // converters touches (row m, chunk c) for m in {m0, m0+1}, c in 0..3: bank = 8c + 4(m - m0) mod 32 is then
    // distinct for all eight 16-byte accesses (a 16 B pad made (c, m0+1) collide with (c+1, m0))
    static constexpr int A_LBO = GT_ROWS * 16 + 32;
    static constexpr int A_BYTES = GT_CH * A_LBO;         // one of hi/lo per stage
    static constexpr int B_BYTES = GT_CH * BN * 16;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int B_U4 = 2 * GT_CH * BN;                     // uint4 of B per slice (hi + lo)
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulator buffers
    static constexpr size_t SMEM = (size_t)GT_STAGES * STAGE + GT_MAX_SRC * GT_ROWS * 8 + GT_ROWS * 12 + 512;
};

// Pipeline (per CTA, one [128 x BN] output tile, k-slices of 32 channels through a GT_STAGES ring),
// 13 warps with fixed roles -- the roles are kept in different threads on purpose: the proxy fence that
// publishes converted operands drains the issuing thread's outstanding memory operations, so a thread
// that prefetches AND fences never has a load in flight (measured 2.5 us per slice at any depth), and a
// thread that converts AND issues MMAs serialises 700 + 740 cycles per slice:
//   warp 0        MMA issuer: hands the weight-image slices to the bulk-copy engine (cp.async.bulk,
//                 bytes counted on full[stage]), waits conv[stage], one elected lane issues 6 tcgen05.mma,
//                 tcgen05.commit -> empty[stage] (+ chunk[] at chunk ends)
//   warps 1-4     loaders: raw fp32 A pieces global -> registers -> shared (LDG.128 / STS.128, two slices
//                 in flight) straight INTO the slots where their fp16 hi / lo words will live;
//                 mbarrier.arrive -> full[stage]
//   warps 5-8 /   two converter groups taking alternate slices: each thread rewrites its own 16-byte
//   warps 9-12    slots in place (x * 2^e -> hi, lo), fences, arrives on conv[stage]; group g also owns
//                 column half g of the accumulator for the flushes and the epilogue; the same 8 warps
//                 run the range pass (max |A| of the tile) on their own named barrier at the start and
//                 the shared-memory-staged, row-coalesced output stores at the end
//   flush         every GT_FLUSH slices the TMEM accumulator is added (RN) into registers and the next
//                 chunk starts fresh in the other TMEM buffer: the tensor core accumulates with
//                 truncation (measured -3e-8 relative per accumulation, -5e-5 at K = 7680 otherwise)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
    using C = GtCfg<BN>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stages = smem;
    const float** rowptr = reinterpret_cast<const float**>(stages + GT_STAGES * C::STAGE);  // [src][row]
    int* rowinfo = reinterpret_cast<int*>(rowptr + GT_MAX_SRC * GT_ROWS);                   // [row][3]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(rowinfo + GT_ROWS * 3);
    uint64_t* full_bar = mbar;                       // [S] loaders' cp.async landed
    uint64_t* conv_bar = mbar + GT_STAGES;           // [S] operands converted + fenced
    uint64_t* empty_bar = mbar + 2 * GT_STAGES;      // [S] MMAs that read the stage are done
    uint64_t* chunk_bar = mbar + 3 * GT_STAGES;      // [2] accumulation chunk complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3 * GT_STAGES + 2);
    __shared__ unsigned amax_warp[GT_THREADS / 32];

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const int64_t row0 = (int64_t)blockIdx.x * GT_ROWS;
    const int col0 = blockIdx.y * BN;
#ifdef O3DML_DEBUG_TIMING
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0;
    if (dbg && tid == 0) g_gt_dbg[4000] = clock64();
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0 && cta_lin < 1000) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_gt_dbg[5000 + cta_lin] = (long long)gt;
    }
#endif

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
        for (int i = 0; i < GT_STAGES; ++i) {
            tc::mbar_init(&full_bar[i], GT_LOADERS + 1);   // loaders + the weight-slice issuer
            tc::mbar_init(&conv_bar[i], GT_CONV);
            tc::mbar_init(&empty_bar[i], 1);
        }
        tc::mbar_init(&chunk_bar[0], 1);
        tc::mbar_init(&chunk_bar[1], 1);
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
    const bool conv_fast = p.mode == 1 && (p.C % GT_KS) == 0;   // a 32-channel slice never straddles taps
    const int slices_per_tap = conv_fast ? p.C / GT_KS : 1;

    // source address of the 8 floats of (row m, chunk c) in slice s; nullptr = zeros
    auto a_src = [&](int s, int m, int c) -> const float* {
        const int k = s * GT_KS + c * 8;
        if (k >= p.K) return nullptr;
        if (p.mode == 1) {
            if (rowinfo[m * 3] < 0) return nullptr;
            int tap, cc;
            if (conv_fast) {
                tap = s / slices_per_tap;
                cc = (s - tap * slices_per_tap) * GT_KS + c * 8;
            } else {
                tap = k / p.C;
                cc = k - tap * p.C;
            }
            const int iy = rowinfo[m * 3 + 1] + tap / 3, ix = rowinfo[m * 3 + 2] + tap % 3;
            if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return nullptr;
            return p.src[0].data + ((int64_t)rowinfo[m * 3] + (int64_t)iy * p.W + ix) * p.C + cc;
        }
        int sidx = 0;
        while (sidx + 1 < p.nsrc && k >= p.koff[sidx + 1]) ++sidx;
        const float* base = rowptr[sidx * GT_ROWS + m];
        return base ? base + (k - p.koff[sidx]) : nullptr;
    };

    float a_scale = 1.f, out_scale = 1.f;
    // ---- range pass: max |A| over this CTA's tile -> exact power-of-two scale.  Only the converter
    // warps (the consumers of a_scale / out_scale) take part, on their own named barrier: the MMA warp
    // starts the weight-slice bulk copies and the loader warps the first GT_STAGES slices of A meanwhile.
    constexpr int RT = 2 * GT_CONV;                       // range threads
    if (tid >= GT_THREADS - RT) {
        const int rtid = tid - (GT_THREADS - RT);
        float mx = 0.f;
        if (p.mode == 1) {
            // conv: every tap of every row lies in ONE contiguous pixel span of the NHWC input (a superset
            // is fine for an upper bound) -> a single coalesced sweep instead of 9 gathers per row
            const int64_t per = (int64_t)p.OH * p.OW;
            const int64_t n_last = min(row0 + GT_ROWS, p.N) - 1;
            auto in_pix = [&](int64_t n, int dy, int dx) {
                const int64_t b = n / per, r = n % per;
                const int64_t iy = (r / p.OW) * p.stride - 1 + dy, ix = (r % p.OW) * p.stride - 1 + dx;
                return b * p.H * p.W + iy * p.W + ix;
            };
            const int64_t tot_pix = (p.N / per) * (int64_t)p.H * p.W;
            int64_t lo = max((int64_t)0, in_pix(row0, 0, 0));
            int64_t hi = min(tot_pix - 1, in_pix(n_last, 2, 2));
            const float4* base = reinterpret_cast<const float4*>(p.src[0].data + lo * p.C);
            const int64_t n4 = (hi - lo + 1) * p.C / 4;
            for (int64_t i = rtid; i < n4; i += 4 * RT) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = (i + u * RT < n4) ? base[i + u * RT] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
            }
        } else {
            const int total = nsl * (GT_ROWS * GT_CH);            // (slice, row, chunk) items
            for (int i0 = rtid; i0 < total; i0 += 4 * RT) {
                float4 v[4][2];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                      // 4 independent gathers in flight per thread
                    const int i = i0 + u * RT;
                    const float* src = nullptr;
                    if (i < total) {
                        const int s = i / (GT_ROWS * GT_CH), rem = i - s * (GT_ROWS * GT_CH);
                        src = a_src(s, rem >> 2, rem & 3);
                    }
                    v[u][0] = src ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u][1] = src ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u][h].x), fabsf(v[u][h].y)),
                                             fmaxf(fabsf(v[u][h].z), fabsf(v[u][h].w))));
            }
        }
        const unsigned wmx = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));  // mx >= 0: bits are ordered
        if ((rtid & 31) == 0) amax_warp[rtid >> 5] = wmx;
        asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory");
        unsigned bm = 0;
#pragma unroll
        for (int i = 0; i < RT / 32; ++i) bm = max(bm, amax_warp[i]);
        const float amax = __uint_as_float(bm);
        int e = 0;
        if (amax > 0.f && amax < 3.0e38f) e = 13 - ilogbf(amax);   // amax * 2^e in [2^13, 2^14)
        e = max(-100, min(100, e));
        a_scale = ldexpf(1.f, e);
        out_scale = ldexpf(1.f, -e - p.wexp);
    }
#ifdef O3DML_DEBUG_TIMING
    if (dbg && tid == 0) g_gt_dbg[4001] = clock64();
#endif
    const int last_chunk = (nsl - 1) / GT_FLUSH;

    if (warp == 0) {
        // ================================================================= MMA issuer
        constexpr uint32_t idesc = tc::idesc_f16(GT_ROWS, BN);
        constexpr uint32_t A_LBO = C::A_LBO, B_LBO = BN * 16;
        // this warp also feeds the weight-image slices: 2 x GT_CH contiguous [BN x 16 B] segments per
        // slice go to the bulk-copy engine (8 copies instead of BN*8 cp.async), counted in bytes on full[]
        const uint4* b_img0 = p.wimg + col0;
        const size_t b_step = (size_t)GT_CH * p.Npad;
        auto issue_b = [&](int sl) {
            const int st = sl % GT_STAGES;
            uint4* bdst = reinterpret_cast<uint4*>(stages + (size_t)st * C::STAGE + 2 * C::A_BYTES);
            mbar_arrive_expect_tx(&full_bar[st], 2 * C::B_BYTES);
#pragma unroll
            for (int img = 0; img < 2; ++img)
#pragma unroll
                for (int cc = 0; cc < GT_CH; ++cc)
                    bulk_copy_g2s(bdst + (img * GT_CH + cc) * BN,
                                  b_img0 + img * img_u4 + (size_t)sl * b_step + (size_t)cc * p.Npad, BN * 16,
                                  &full_bar[st]);
        };
        if (elect_one()) {
            for (int sl = 0; sl < GT_STAGES && sl < nsl; ++sl) issue_b(sl);
        }
        __syncwarp();
        for (int s = 0; s < nsl; ++s) {
            const int stage = s % GT_STAGES, use = s / GT_STAGES, chunk = s / GT_FLUSH;
#ifdef O3DML_DEBUG_TIMING
            const long long tm0 = clock64();
#endif
            tc::mbar_wait(&conv_bar[stage], use & 1);
            tc::tc_fence_after();
#ifdef O3DML_DEBUG_TIMING
            const long long tm1 = clock64();
#endif
            if (elect_one()) {
                const uint32_t ah0 = tc::smem_u32(stages + (size_t)stage * C::STAGE);
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
#ifdef O3DML_DEBUG_TIMING
                if (dbg && s < 300) { g_gt_dbg[2000 + 3 * s] = tm0; g_gt_dbg[2001 + 3 * s] = tm1; g_gt_dbg[2002 + 3 * s] = clock64(); }
#endif
            }
            __syncwarp();
            // refill the weight half of the stage that slice s-1 used (its MMAs were issued one slice ago)
            const int sn = s + GT_STAGES - 1;
            if (s >= 1 && sn < nsl) {
                tc::mbar_wait(&empty_bar[(s - 1) % GT_STAGES], ((s - 1) / GT_STAGES) & 1);
                if (elect_one()) issue_b(sn);
                __syncwarp();
            }
        }
    } else if (warp < 5) {
        // ================================================================= loaders (GT_LOADERS threads)
        const int rt = tid - 32;
        constexpr int LA = (GT_ROWS * GT_CH) / GT_LOADERS;                    // (row, chunk) items per thread
        // everything that does not depend on the slice index is resolved once
        int lm[LA], lc[LA];
        const float* lbase[LA];      // conv: pixel (iy0, ix0) of the row; rows mode with one source: the row
        unsigned ltaps[LA];          // conv: bit t set when tap t lies inside the image
#pragma unroll
        for (int j = 0; j < LA; ++j) {
            const int item = rt + j * GT_LOADERS;
            lm[j] = item >> 2;       // 4 consecutive lanes read the 4 x 32 B of one row's 128-byte slice
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
        // the weight-image slice is 2 x GT_CH contiguous [BN x 16 B] segments: one loader thread hands them
        // to the bulk-copy engine (8 copies of BN*16 bytes instead of BN*8 cp.async instructions)
        // A pieces travel global -> registers -> shared (LDG.128 / STS.128): cp.async issued from here ran
        // at ~30 cycles per warp instruction (1000 cycles per slice).  Two slices stay in flight in
        // registers; these threads never execute a proxy fence, so nothing drains the loads early.
        // a single warp retires ~1 dependent instruction per 4-6 cycles, so the per-slice work of a loader
        // thread has to stay well under 100 instructions: slices are visited in order and the source
        // address of every piece is advanced incrementally (no division in the loop)
        int ld_tap = 0, ld_in_tap = 0, ld_sl = 0;     // state of the NEXT slice to load (conv_fast)
        int64_t ld_toff = 0;                          // float offset of the current tap + channel block
        constexpr int LOOK = 2;
        float4 rg[LOOK + 1][LA][2];
        auto load_slice = [&](int slot) {             // loads slice ld_sl into register slot `slot`
            const int sl = ld_sl;
#pragma unroll
            for (int j = 0; j < LA; ++j) {
                const float* src = nullptr;
                if (sl < nsl) {
                    const int k = sl * GT_KS + lc[j] * 8;
                    if (k < p.K) {
                        if (conv_fast) {
                            if ((ltaps[j] >> ld_tap) & 1u) src = lbase[j] + ld_toff + lc[j] * 8;
                        } else if (p.mode == 1) {
                            const int t = k / p.C, cc = k - t * p.C;
                            if ((ltaps[j] >> t) & 1u) src = lbase[j] + ((int64_t)(t / 3) * p.W + (t % 3)) * p.C + cc;
                        } else if (p.nsrc == 1) {
                            if (lbase[j]) src = lbase[j] + k;
                        } else {
                            src = a_src(sl, lm[j], lc[j]);
                        }
                    }
                }
                rg[slot][j][0] = src ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                rg[slot][j][1] = src ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ++ld_sl;
            if (conv_fast) {
                ld_toff += GT_KS;
                if (++ld_in_tap == slices_per_tap) {
                    ld_in_tap = 0;
                    ++ld_tap;
                    ld_toff = ((int64_t)(ld_tap / 3) * p.W + (ld_tap % 3)) * p.C;
                }
            }
        };
#pragma unroll
        for (int l = 0; l < LOOK; ++l) load_slice(l);
        for (int s0 = 0; s0 < nsl; s0 += LOOK + 1) {
#pragma unroll
            for (int u = 0; u < LOOK + 1; ++u) {       // register slots are compile-time indices
                const int s = s0 + u;
                if (s < nsl) {
                    load_slice((u + LOOK) % (LOOK + 1));
                    const int stage = s % GT_STAGES, use = s / GT_STAGES;
#ifdef O3DML_DEBUG_TIMING
                    const long long tl0 = clock64();
#endif
                    if (use >= 1) tc::mbar_wait(&empty_bar[stage], (use - 1) & 1);   // MMAs of slice s-STAGES done
#ifdef O3DML_DEBUG_TIMING
                    const long long tl1 = clock64();
#endif
                    uint8_t* a_hi = stages + (size_t)stage * C::STAGE;
                    uint8_t* a_lo = a_hi + C::A_BYTES;
#pragma unroll
                    for (int j = 0; j < LA; ++j) {
                        const uint32_t off = (uint32_t)lc[j] * C::A_LBO + (uint32_t)lm[j] * 16u;
                        *reinterpret_cast<float4*>(a_hi + off) = rg[u][j][0];
                        *reinterpret_cast<float4*>(a_lo + off) = rg[u][j][1];
                    }
#ifdef O3DML_DEBUG_TIMING
                    const long long tl2 = clock64();
#endif
                    mbar_arrive(&full_bar[stage]);                                   // release: stores visible
#ifdef O3DML_DEBUG_TIMING
                    if (dbg && rt == 0 && s < 700) {
                        g_gt_dbg[1000 + 4 * s] = tl0; g_gt_dbg[1001 + 4 * s] = tl1; g_gt_dbg[1002 + 4 * s] = tl2;
                        g_gt_dbg[1003 + 4 * s] = clock64();
                    }
#endif
                }
            }
        }
    } else {
        // ================================================================= converters + flush + epilogue
        const int grp = (warp - 5) >> 2;               // converter group = column half
        const int rt = (warp & 3) * 32 + (tid & 31);   // output row = TMEM lane this thread can read
        const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int HB = BN / 2 < 16 ? 16 : BN / 2;  // columns owned by a group
        constexpr int NQ = HB / 16;
        float racc[HB];
#pragma unroll
        for (int i = 0; i < HB; ++i) racc[i] = 0.f;
        const bool owns_cols = (BN >= 32) || grp == 0;
        const int colbase = (BN >= 32) ? grp * HB : 0;
        uint32_t ph_chunk[2] = {0, 0};
        int next_flush = 0;                            // first chunk this group has not folded yet
        auto fold = [&](int chunk) {                   // racc += accumulator of `chunk` (RN adds)
            tc::mbar_wait(&chunk_bar[chunk & 1], ph_chunk[chunk & 1]);
            ph_chunk[chunk & 1] ^= 1;
            tc::tc_fence_after();
            if (owns_cols) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float v[16];
                    tc::tmem_ld16(tmem_lane + (chunk & 1) * BN + colbase + q * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) racc[q * 16 + j] += v[j];
                }
            }
            tc::tc_fence_before();
        };
        for (int s = grp; s < nsl; s += 2) {
            const int stage = s % GT_STAGES, use = s / GT_STAGES;
            // chunks that ended at least one slice ago have drained: fold them before converting on
            while (next_flush < last_chunk && s >= (next_flush + 1) * GT_FLUSH + 1) fold(next_flush++);
#ifdef O3DML_DEBUG_TIMING
            const long long tq0 = clock64();
#endif
            tc::mbar_wait(&full_bar[stage], use & 1);         // every piece of slice s has landed
#ifdef O3DML_DEBUG_TIMING
            const long long tq1 = clock64();
#endif
            uint8_t* a_hi = stages + (size_t)stage * C::STAGE;
            uint8_t* a_lo = a_hi + C::A_BYTES;
#pragma unroll
            for (int j = 0; j < (GT_ROWS * GT_CH) / GT_CONV; ++j) {   // 512 items, 4 per thread, in place
                const int item = rt + j * GT_CONV;
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
            mbar_arrive(&conv_bar[stage]);
#ifdef O3DML_DEBUG_TIMING
            if (dbg && rt == 0 && s < 900) {
                g_gt_dbg[4 * s + 0] = tq0; g_gt_dbg[4 * s + 1] = tq1; g_gt_dbg[4 * s + 2] = clock64();
            }
#endif
        }
        while (next_flush < last_chunk) fold(next_flush++);
        tc::mbar_wait(&chunk_bar[last_chunk & 1], ph_chunk[last_chunk & 1]);
        tc::tc_fence_after();
#ifdef O3DML_DEBUG_TIMING
        if (dbg && rt == 0 && grp == 0) g_gt_dbg[4002] = clock64();
#endif
        // ---- epilogue: thread = output row, this group's column half, 16 columns at a time.
        // Row-major and pixel-shuffle outputs are staged through the (now dead) pipeline stages so that the
        // global stores are whole rows written by consecutive lanes: with thread = row, one store
        // instruction touched 32 different rows (32 half-filled sectors; 154 us for the 53 k-row K = 64
        // deconvolution of the PointPillars neck).
        const int64_t n = row0 + rt;
        constexpr int SLD = BN + 4;                                  // staging row stride (floats)
        float* stg = reinterpret_cast<float*>(stages);
        const bool staged = p.out_mode == 0 || (p.out_mode == 2 && (p.dC & 3) == 0);
        if (owns_cols) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int c0 = colbase + 16 * q;
                float v[16];
                tc::tmem_ld16(tmem_lane + (last_chunk & 1) * BN + c0, v);   // warp-collective
                const int cbase = col0 + c0;
                if (n >= p.N || cbase >= p.Cout) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = cbase + j;
                    if (c < p.Cout) {
                        float x = (v[j] + racc[q * 16 + j]) * out_scale;   // exact: undoes the 2^e scalings
                        x = fmaf(x, p.scale ? p.scale[c] : 1.f, p.shift ? p.shift[c] : 0.f);
                        if (p.residual) x += p.residual[(size_t)n * p.res_ld + c];
                        v[j] = apply_act(x, p.act, p.slope);
                    }
                }
                if (staged) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<float4*>(stg + rt * SLD + c0 + 4 * u) =
                            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                } else if (p.out_mode == 0) {
                    float* o = p.out + (size_t)n * p.out_ld + cbase;
                    if (cbase + 15 < p.Cout && (p.out_ld & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            *reinterpret_cast<float4*>(o + 4 * u) =
                                make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
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
        if (staged) {
            named_bar_sync(2, 2 * GT_CONV);                            // the 8 epilogue warps
            constexpr int LPR = BN / 4;                                // lanes per output row
            const int et = grp * GT_CONV + rt;                         // 0 .. 255
            const int c = (et % LPR) * 4, cg = col0 + c;
            const bool vec = (p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
            if (cg < p.Cout) {
                for (int r = et / LPR; r < GT_ROWS; r += (2 * GT_CONV) / LPR) {
                    const int64_t nr = row0 + r;
                    if (nr >= p.N) break;
                    const float4 v = *reinterpret_cast<const float4*>(stg + r * SLD + c);
                    float* o;
                    if (p.out_mode == 0) {
                        o = p.out + (size_t)nr * p.out_ld + cg;
                    } else {
                        const int64_t per = (int64_t)p.dIH * p.dIW;
                        const int64_t b = nr / per;
                        const int rr = (int)(nr % per);
                        const int iy = rr / p.dIW, ix = rr % p.dIW;
                        const int sub = cg / p.dC, co = cg - sub * p.dC;
                        const int dy = sub / p.ds, dx = sub - dy * p.ds;
                        const size_t opix = ((size_t)b * p.dIH * p.ds + (size_t)iy * p.ds + dy) * (p.dIW * p.ds) +
                                            (size_t)ix * p.ds + dx;
                        o = p.out + opix * p.out_ld + co;
                    }
                    if (vec && cg + 3 < p.Cout) {
                        *reinterpret_cast<float4*>(o) = v;
                    } else {
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (cg + j < p.Cout) o[j] = e[j];
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
#ifdef O3DML_DEBUG_TIMING
    if (dbg && tid == 0) g_gt_dbg[4003] = clock64();
    if (tid == 0 && cta_lin < 1000) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_gt_dbg[6000 + cta_lin] = (long long)gt;
    }
#endif
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

#ifdef O3DML_DEBUG_TIMING
extern "C" __attribute__((visibility("default"))) int o3dml_gt_debug_read(long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, o3dml::g_gt_dbg, sizeof(long long) * n);
}
#endif
