// 3xBF16 tcgen05 engine for the MLP GEMMs on sm_100a (the default engine).
//
//   C[M,N] = sum_k A[M,k] * B[N,k]      fp32 values, both operands K-contiguous, fp32 accumulation in TMEM
//
// Every fp32 operand exists in HBM as two bf16 copies, hi = bf16_rn(x) and lo = bf16_rn(x - hi), written by the kernel that
// PRODUCED the operand (forward epilogue, act_bn_bwd_kernel, x0_split_kernel, dense_apply_kernel), so this kernel moves and
// multiplies bf16 only: per 64-element k-block four TMA tiles (A_hi, A_lo, B_hi, B_lo; 128-byte swizzle) and
// 4 k-steps x 3 tcgen05.mma.kind::f16 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) into one fp32 accumulator.  hi + lo represents x to
// 2^-17 and the dropped a_lo*b_lo term is below 2^-16, so one product carries ~1e-5 relative error (rms ~8e-6; random signs, so a
// long dot product does better): measured 5e-6 on the logits of the benchmark shape, up to ~1e-4 on small ill-conditioned
// towers (tests/test_gpu_parity.py, scratch/engine_err.py).  That is a FAST mode: it meets the 1e-4 logit bar on the
// benchmarked configuration (bench.py re-checks it against the oracle in the same run) but is not fp32-faithful the way the
// 3xTF32 engine (gemm_tc.cu, 2^-21) is; it costs half the shared-memory bytes per flop, runs at twice the tensor-pipe rate and
// needs no in-kernel splitting pass.  (Keeping the residual in fp16 would give 19 bits, but tcgen05.mma.kind::f16 traps with an
// illegal-instruction fault on sm_100a when the A and B formats differ — tried, reverted.)
//
// Operand majors.  Forward: A = activations [m][k] (K-major), B = weights W[k][n] (MN-major: n contiguous) — no transposed weight
// copy exists.  Data gradient: A = dZ [m][n] and B = W[k_in][n], both K-major over n.  Weight gradient: A = layer input [b][k_in]
// and B = dZ [b][n], both MN-major with the batch as the reduction dimension — no transposed activation copies exist either,
// and TMA zero-fills the batch tail.  An MN-major operand is fetched as 64-column boxes (64 k-rows x 128 B, 128-byte swizzle);
// its UMMA descriptor uses LBO = one box (8 KB) between 64-wide column groups and SBO = 1 KB between 8-row k groups.
//
// One persistent CTA per SM, six warps: 0 = TMA producer, 1 = MMA issuer (one elected lane), 2-5 = epilogue (TMEM lane quarter
// = warp % 4).  The TMEM accumulator is double buffered (2 x TBN columns) so the epilogue's global stores of tile i overlap the
// MMAs of tile i+1.  Forward epilogue = bias + activation + BN-affine, stores the fp32 outputs and the bf16 hi/lo copies (row
// major and transposed) that the next layer / the weight gradient will read.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace wd {

int tc_make_map_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows);   // gemm_tc.cu
int tc_make_map_out(CUtensorMap* map, const void* ptr, int esize, int rows, int cols, int64_t ld, int nz, int64_t zstride,
                    int box_cols, int box_rows, int swizzle);                                        // gemm_tc.cu

namespace {

constexpr int QBM = 128;         // UMMA M
constexpr int QBK = 64;          // bf16 elements per k-block = one 128-byte swizzle row
constexpr int Q_THREADS = 192;

struct QMaps {
    CUtensorMap a_hi[kMaxSegs], a_lo[kMaxSegs];
    CUtensorMap b_hi, b_lo;
    // TMA-store targets: STORE / WGRAD output C; FWD outputs A (post-activation), H (optional), bf16 hi / lo copies of H
    CUtensorMap o_c, o_a, o_h, o_hs_hi, o_hs_lo;
};
struct QSegs { int n; int k[kMaxSegs]; int koff[kMaxSegs]; };

template <int TBN, int MODE>
__global__ void __launch_bounds__(Q_THREADS, 1) tc_gemm_bf16_kernel(const __grid_constant__ QMaps maps, const QSegs segs, int M, int N, int ktot,
                                                                   int ksplit_len, int nsplit, Epi ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr bool A_MN = MODE == EPI_WGRAD;                       // operand stored with its M / N index contiguous
    constexpr bool B_MN = MODE != EPI_STORE;
    constexpr int A_BYTES = QBM * 128, B_BYTES = TBN * 128;
    constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    constexpr int NST = TBN == 256 ? 2 : 3;
    auto a_hi = [&](int s) { return base + s * STAGE_BYTES; };
    auto a_lo = [&](int s) { return base + s * STAGE_BYTES + A_BYTES; };
    auto b_hi = [&](int s) { return base + s * STAGE_BYTES + 2 * A_BYTES; };
    auto b_lo = [&](int s) { return base + s * STAGE_BYTES + 2 * A_BYTES + B_BYTES; };
    uint8_t* stage_out = base + NST * STAGE_BYTES;                 // 4 epilogue warps x 2 x 4 KB store staging (1024-byte aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 32768);
    uint64_t* full = bars; uint64_t* empty = bars + NST;
    uint64_t* tmem_full = bars + 2 * NST; uint64_t* tmem_empty = bars + 2 * NST + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);
    float* epi_params = reinterpret_cast<float*>(bars + 16);      // [4 epilogue warps][bias | scale | shift][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + TBN - 1) / TBN, tiles_m = (M + QBM - 1) / QBM;
    const int ntiles = tiles_n * tiles_m * nsplit;
    // k-blocks of one tile: FWD / STORE walk the segments (each padded to whole k-blocks by TMA zero fill), WGRAD walks its split
    int nkb_all = 0;
    for (int s = 0; s < segs.n; ++s) nkb_all += (segs.k[s] + QBK - 1) / QBK;
    auto tile_range = [&](int tile, int& m0, int& n0, int& z, int& kbeg, int& nkb) {
        z = tile / (tiles_n * tiles_m);
        int r = tile % (tiles_n * tiles_m);
        m0 = (r / tiles_n) * QBM;
        n0 = (r % tiles_n) * TBN;
        kbeg = 0;
        nkb = nkb_all;
        if (MODE == EPI_WGRAD) {
            kbeg = z * ksplit_len;
            int kend = min(ktot, kbeg + ksplit_len);
            nkb = kend > kbeg ? (kend - kbeg + QBK - 1) / QBK : 0;
        }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * TBN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int m0, n0, z, kbeg, nkb;
                tile_range(tile, m0, n0, z, kbeg, nkb);
                int seg = 0, kin = 0;                            // current segment, k offset inside it
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST, it = g / NST;
                    if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
                    int ka, kbcoord;
                    if (MODE == EPI_WGRAD) { ka = kbeg + kb * QBK; kbcoord = ka; }
                    else {
                        if (kin >= segs.k[seg]) { ++seg; kin = 0; }
                        ka = kin; kbcoord = segs.koff[seg] + kin;
                        kin += QBK;
                    }
                    mbar_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
                    if (!A_MN) {
                        tma_load_2d(a_hi(s), &maps.a_hi[seg], &full[s], ka, m0);
                        tma_load_2d(a_lo(s), &maps.a_lo[seg], &full[s], ka, m0);
                    } else {
#pragma unroll
                        for (int i = 0; i < QBM / 64; ++i) {
                            tma_load_2d(a_hi(s) + i * 8192, &maps.a_hi[seg], &full[s], m0 + 64 * i, ka);
                            tma_load_2d(a_lo(s) + i * 8192, &maps.a_lo[seg], &full[s], m0 + 64 * i, ka);
                        }
                    }
                    if (!B_MN) {
                        tma_load_2d(b_hi(s), &maps.b_hi, &full[s], kbcoord, n0);
                        tma_load_2d(b_lo(s), &maps.b_lo, &full[s], kbcoord, n0);
                    } else {
#pragma unroll
                        for (int i = 0; i < TBN / 64; ++i) {
                            tma_load_2d(b_hi(s) + i * 8192, &maps.b_hi, &full[s], n0 + 64 * i, kbcoord);
                            tma_load_2d(b_lo(s) + i * 8192, &maps.b_lo, &full[s], n0 + 64 * i, kbcoord);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(QBM >> 4) << 24);
        constexpr uint32_t a_step = A_MN ? 2048u : 32u, b_step = B_MN ? 2048u : 32u;   // bytes per 16-element k-step
        int g = 0, use = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int m0, n0, z, kbeg, nkb;
            tile_range(tile, m0, n0, z, kbeg, nkb);
            if (nkb == 0) continue;
            const int a = use & 1, au = use >> 1;
            if (au > 0) mbar_wait(&tmem_empty[a], (au - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(a * TBN);
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const int s = g % NST, it = g / NST;
                mbar_wait(&full[s], it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sa_hi = smem_u32(a_hi(s)), sb_hi = smem_u32(b_hi(s));
                    const uint32_t sa_lo = smem_u32(a_lo(s)), sb_lo = smem_u32(b_lo(s));
#pragma unroll
                    for (int k = 0; k < QBK / 16; ++k) {
                        const uint64_t da_hi = A_MN ? make_desc_mn(sa_hi + k * a_step) : make_desc(sa_hi + k * a_step);
                        const uint64_t da_lo = A_MN ? make_desc_mn(sa_lo + k * a_step) : make_desc(sa_lo + k * a_step);
                        const uint64_t db_hi = B_MN ? make_desc_mn(sb_hi + k * b_step) : make_desc(sb_hi + k * b_step);
                        const uint64_t db_lo = B_MN ? make_desc_mn(sb_lo + k * b_step) : make_desc(sb_lo + k * b_step);
                        umma_bf16(tacc, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(tacc, da_hi, db_lo, idesc, 1u);
                        umma_bf16(tacc, da_hi, db_hi, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == nkb - 1) umma_commit(&tmem_full[a]);
                }
                __syncwarp();
            }
            ++use;
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        // A thread owns one accumulator row (tcgen05.ld 32x32b), so direct global stores would touch 32 different lines per
        // instruction.  Each warp instead stages its 32 x 32 chunk in shared memory (swizzled, conflict-free) and one lane hands
        // it to the TMA store unit; two 4 KB staging buffers per warp alternate so a store drains while the next chunk is built.
        const int q = warp & 3;
        uint8_t* stg = stage_out + (warp - 2) * 8192;
        int nuse = 0, use = 0;
        auto acquire = [&]() -> uint8_t* {
            if (lane == 0) bulk_wait_read<1>();                   // the store issued two uses ago has finished reading its buffer
            __syncwarp();
            uint8_t* p = stg + (nuse & 1) * 4096;
            ++nuse;
            return p;
        };
        auto put_f32 = [&](uint8_t* p, const uint32_t (&x)[32]) {     // [32 rows][32 floats], 128-byte swizzle, lane = row
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(p + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        };
        auto put_bf16 = [&](uint8_t* p, const uint32_t (&x)[16]) {    // [32 rows][32 bf16], 64-byte swizzle, lane = row
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(p + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        };
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int m0, n0, z, kbeg, nkb;
            tile_range(tile, m0, n0, z, kbeg, nkb);
            const int mw = m0 + q * 32;                        // first row of this warp's chunk
            const int m = mw + lane;
            const int a = use & 1, au = use >> 1;
            // bias / BN scale / BN shift of column nb + lane: fetched one chunk ahead (the first chunk's before the wait for the
            // accumulator), so their global-load latency never sits on the epilogue's critical path
            float pb = 0.f, pg = 1.f, pe = 0.f;
            auto load_params = [&](int nbx, float& b_, float& g_, float& e_) {
                const int gn = nbx + lane;
                const bool in = gn < ep.n_logical;
                b_ = in ? ep.bias[gn] : 0.f;
                g_ = (in && ep.bn) ? ep.gamma[gn] * 0.99950037468777f : 1.f;
                e_ = (in && ep.bn) ? ep.beta[gn] : 0.f;
            };
            if (MODE == EPI_FWD) load_params(n0, pb, pg, pe);
            if (nkb > 0) {
                mbar_wait(&tmem_full[a], au & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#pragma unroll 1
            for (int c = 0; c < TBN / 32; ++c) {
                const int nb = n0 + c * 32;
                if (nb >= N) break;
                float qb = 0.f, qg = 1.f, qe = 0.f;
                if (MODE == EPI_FWD && c + 1 < TBN / 32 && nb + 32 < N) load_params(nb + 32, qb, qg, qe);
                uint32_t v[32];
                if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TBN + c * 32), v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (MODE == EPI_FWD) {
                    float* wp = epi_params + (warp - 2) * 96;
                    wp[lane] = pb; wp[32 + lane] = pg; wp[64 + lane] = pe;
                    pb = qb; pg = qg; pe = qe;
                    __syncwarp();
                    uint32_t h[32];
                    const bool rv = m < ep.m_valid;
                    if (ep.act == WD_ACT_RELU) {                      // warp-uniform fast path
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const bool ok = rv && (nb + j) < ep.n_logical;
                            const float av = ok ? fmaxf(__uint_as_float(v[j]) + wp[j], 0.f) : 0.f;
                            v[j] = __float_as_uint(av);
                            h[j] = __float_as_uint(ok ? fmaf(av, wp[32 + j], wp[64 + j]) : 0.f);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const bool ok = rv && (nb + j) < ep.n_logical;
                            const float av = ok ? act_fwd(ep.act, __uint_as_float(v[j]) + wp[j]) : 0.f;
                            v[j] = __float_as_uint(av);
                            h[j] = __float_as_uint(ok ? fmaf(av, wp[32 + j], wp[64 + j]) : 0.f);
                        }
                    }
                    __syncwarp();
                    if (ep.A_out != ep.H_out) {
                        uint8_t* p = acquire();
                        put_f32(p, v);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&maps.o_a, p, nb, mw, 0); bulk_commit(); }
                    }
                    if (ep.H_out) {                                    // fp32 copy only where something reads it (logits layer)
                        uint8_t* p = acquire();
                        put_f32(p, h);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&maps.o_h, p, nb, mw, 0); bulk_commit(); }
                    }
                    uint32_t hh[16], hl[16];                           // packed bf16 pairs of the hi / lo copies
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        __nv_bfloat16 h0, l0, h1, l1;
                        split_bf16(__uint_as_float(h[2 * j]), h0, l0);
                        split_bf16(__uint_as_float(h[2 * j + 1]), h1, l1);
                        hh[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        hl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    {
                        uint8_t* p = acquire();
                        put_bf16(p, hh);
                        put_bf16(p + 2048, hl);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_3d(&maps.o_hs_hi, p, nb, mw, 0); tma_store_3d(&maps.o_hs_lo, p + 2048, nb, mw, 0); bulk_commit(); }
                    }
                } else {
                    uint8_t* p = acquire();
                    put_f32(p, v);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (MODE == EPI_STORE && ep.accumulate) tma_reduce_add_3d(&maps.o_c, p, nb, mw, 0);
                        else tma_store_3d(&maps.o_c, p, nb, mw, MODE == EPI_WGRAD ? z : 0);
                        bulk_commit();
                    }
                }
            }
            if (nkb > 0) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&tmem_empty[a]);                   // accumulator buffer a may be overwritten
                ++use;
            }
        }
        if (lane == 0) bulk_wait_all();                        // every store has landed before the CTA retires
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * TBN)));
}

template <int TBN, int MODE>
int launch_q(WdModel* m, const QMaps& maps, const QSegs& segs, int M, int N, int ktot, int splits, int ksplit_len, const Epi& ep) {
    constexpr int NST = TBN == 256 ? 2 : 3;
    constexpr int smem = NST * 2 * (QBM * 128 + TBN * 128) + 32768 + 1024 + 128 + 1536;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        WD_CUDA(cudaFuncSetAttribute(tc_gemm_bf16_kernel<TBN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        WD_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, m->device));
        configured = true;
    }
    const int nsplit = MODE == EPI_WGRAD ? splits : 1;
    const int ntiles = ((N + TBN - 1) / TBN) * ((M + QBM - 1) / QBM) * nsplit;
    tc_gemm_bf16_kernel<TBN, MODE><<<ntiles < num_sms ? ntiles : num_sms, Q_THREADS, smem, m->stream>>>(maps, segs, M, N, ktot, ksplit_len, nsplit, ep);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

}  // namespace

// whether the engine uses 128 x 256 output tiles for this problem (also consulted when the split-K factor is chosen)
bool tc_bf16_wide_tile(int mode, int M, int N, int splits, int num_sms) {
    static const bool no_wide = getenv("WD_TC_N128") != nullptr;
    if (no_wide || N <= 128) return false;
    // rounds of the persistent grid x cost of one k-block (shared-memory bytes per k-block: 160 KB vs 240 KB)
    const int64_t tm = (M + QBM - 1) / QBM, sp = mode == EPI_WGRAD ? splits : 1;
    const int64_t t128 = tm * ((N + 127) / 128) * sp, t256 = tm * ((N + 255) / 256) * sp;
    const int64_t c128 = (t128 + num_sms - 1) / num_sms * 1250, c256 = (t256 + num_sms - 1) / num_sms * 1920;
    return c256 < c128;
}

int tc_gemm_bf16(WdModel* m, int mode, const GemmA& A, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, int M, int N,
                 const Epi& ep, int splits, int ksplit_len) {
    if (N % 32 != 0 || A.n > kMaxSegs || !B_hi || !B_lo) { set_error("bf16 GEMM engine: unsupported operands"); return WD_EUNSUPPORTED; }
    static int num_sms = 0;
    if (!num_sms) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, m->device);
    QMaps maps;
    QSegs segs{};
    segs.n = A.n;
    int rc, ktot = 0;
    const bool a_mn = mode == EPI_WGRAD, b_mn = mode != EPI_STORE;
    for (int s = 0; s < A.n; ++s) {
        if (!A.hi[s] || !A.lo[s]) { set_error("bf16 GEMM engine: operand without hi/lo copies"); return WD_EINVAL; }
        // K-major: [M rows][k contiguous], box 64 k x 128 rows.  MN-major: [k rows][M contiguous], box 64 columns x 64 k-rows
        if (!a_mn) {
            if ((rc = tc_make_map_bf16(&maps.a_hi[s], A.hi[s], M, A.k[s], A.ld[s], QBM))) return rc;
            if ((rc = tc_make_map_bf16(&maps.a_lo[s], A.lo[s], M, A.k[s], A.ld[s], QBM))) return rc;
        } else {
            if ((rc = tc_make_map_bf16(&maps.a_hi[s], A.hi[s], A.k[s], M, A.ld[s], 64))) return rc;
            if ((rc = tc_make_map_bf16(&maps.a_lo[s], A.lo[s], A.k[s], M, A.ld[s], 64))) return rc;
        }
        segs.k[s] = A.k[s]; segs.koff[s] = ktot;
        ktot += A.k[s];
    }
    for (int s = A.n; s < kMaxSegs; ++s) { maps.a_hi[s] = maps.a_hi[0]; maps.a_lo[s] = maps.a_lo[0]; }
    const bool wide = tc_bf16_wide_tile(mode, M, N, splits, num_sms);
    const int tbn = wide ? 256 : 128;
    if (!b_mn) {
        if ((rc = tc_make_map_bf16(&maps.b_hi, B_hi, N, ktot, ldb, tbn))) return rc;
        if ((rc = tc_make_map_bf16(&maps.b_lo, B_lo, N, ktot, ldb, tbn))) return rc;
    } else {
        if ((rc = tc_make_map_bf16(&maps.b_hi, B_hi, ktot, N, ldb, 64))) return rc;
        if ((rc = tc_make_map_bf16(&maps.b_lo, B_lo, ktot, N, ldb, 64))) return rc;
    }
    if (mode == EPI_WGRAD) ksplit_len = (ksplit_len + QBK - 1) / QBK * QBK;
    if (mode == EPI_FWD) {
        if (!ep.Hs_hi || !ep.Hs_lo) { set_error("bf16 GEMM engine: forward without hi/lo outputs"); return WD_EINVAL; }
        if ((rc = tc_make_map_out(&maps.o_hs_hi, ep.Hs_hi, 2, M, N, ep.ldh, 1, 0, 32, 32, 1))) return rc;
        if ((rc = tc_make_map_out(&maps.o_hs_lo, ep.Hs_lo, 2, M, N, ep.ldh, 1, 0, 32, 32, 1))) return rc;
        if (ep.H_out) { if ((rc = tc_make_map_out(&maps.o_h, ep.H_out, 4, M, N, ep.ldh, 1, 0, 32, 32, 2))) return rc; }
        else maps.o_h = maps.o_hs_hi;
        if (ep.A_out != ep.H_out) { if ((rc = tc_make_map_out(&maps.o_a, ep.A_out, 4, M, N, ep.ldh, 1, 0, 32, 32, 2))) return rc; }
        else maps.o_a = maps.o_h;
        maps.o_c = maps.o_hs_hi;
    } else {
        if ((rc = tc_make_map_out(&maps.o_c, ep.C, 4, M, N, ep.ldc, mode == EPI_WGRAD ? splits : 1, ep.split_stride, 32, 32, 2))) return rc;
        maps.o_a = maps.o_h = maps.o_hs_hi = maps.o_hs_lo = maps.o_c;
    }
#define WD_Q_LAUNCH(MODE_) \
    return wide ? launch_q<256, MODE_>(m, maps, segs, M, N, ktot, splits, ksplit_len, ep) : launch_q<128, MODE_>(m, maps, segs, M, N, ktot, splits, ksplit_len, ep)
    if (mode == EPI_FWD) { WD_Q_LAUNCH(EPI_FWD); }
    if (mode == EPI_STORE) { WD_Q_LAUNCH(EPI_STORE); }
    WD_Q_LAUNCH(EPI_WGRAD);
#undef WD_Q_LAUNCH
}

}  // namespace wd
