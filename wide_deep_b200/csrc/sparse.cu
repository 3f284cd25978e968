// Sparse half of the step: wide linear logit, multihot embedding gather + mean-pool, and the backward
// scatter of both fused with their per-row optimizers.
//
//   wide_fwd_kernel        tf.feature_column.linear_model(sparse_combiner='sum')   (reference linear.py:29-36)
//   emb_pool_fwd_kernel    embedding_column(combiner='mean') inside input_layer     (reference dnn.py:88-90;
//                          safe_embedding_lookup_sparse: empty bag -> zeros)
//   seg_* / *_apply        SparseSegmentMeanGrad + SparseApplyAdagrad / SparseApplyFtrl (reference
//                          joint.py:234-247 minimize()): one update per touched row from the
//                          ordered sum of its gradients.
// HBM-bound kernels: 16-byte vector loads through the read-only path, several rows in flight per lane
// group, rows are 16..256 B records {w | optimizer slots} so forward touches one line and backward one
// contiguous record.
#include <stdlib.h>

#include "common.cuh"
#include "sparse_dev.cuh"

namespace wd {


// ------------------------------------------------------------------------------------------- wide forward
// one warp per example: sum of w[row] over the example's wide ids (+ bias)
__global__ void __launch_bounds__(256) wide_fwd_kernel(int B, int C, const int32_t* __restrict__ offs,
                                                       const uint32_t* __restrict__ e_wide, const float4* __restrict__ wide,
                                                       const float* __restrict__ bias, float* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    float b0 = bias[0];
    for (int b = warp; b < B; b += nwarps) {
        int s = offs[(int64_t)b * C], e = offs[(int64_t)(b + 1) * C];
        float acc = 0.f;
        for (int j = s + lane; j < e; j += 32) {
            uint32_t r = e_wide[j];
            if (r != kInvalidRow) acc += __ldg(&wide[r].x);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[b] = acc + b0;
    }
}

// --------------------------------------------------------------------------------------- embedding forward
// All tables of one width D (G = D/4 lanes per bag, float4 per lane).  Bag = (example b, table k).
// NARROW: one G-lane group per bag, 32/G bags per warp in flight (single-valued / short bags).
// WIDE (full warp per bag): the 32/G groups take interleaved ids, then a shuffle tree combines them
// (long multihot bags).
template <int G, bool WIDEBAG>
__global__ void __launch_bounds__(256) emb_pool_fwd_kernel(int B, int C, int ntab, const int32_t* __restrict__ tab_ids,
                                                           float* const* __restrict__ tab_data,
                                                           const int32_t* __restrict__ tab_stride,
                                                           const int32_t* __restrict__ tab_x0, const int32_t* __restrict__ tab_col,
                                                           const int64_t* __restrict__ tab_row_base,
                                                           const int32_t* __restrict__ offs, const uint32_t* __restrict__ e_emb,
                                                           float* __restrict__ X0, int ld) {
    constexpr int GROUPS = 32 / G;
    const int lane = threadIdx.x & 31, lig = lane % G, grp = lane / G;
    const int64_t nbags = (int64_t)B * ntab;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (!WIDEBAG) {
        for (int64_t bag = gwarp * GROUPS + grp; bag < nbags; bag += nwarp * GROUPS) {
            int b = (int)(bag / ntab), t = tab_ids[bag % ntab];
            int c = tab_col[t];
            int s = offs[(int64_t)b * C + c], e = offs[(int64_t)b * C + c + 1];
            const float* base = tab_data[t] + lig * 4;
            const int stride = tab_stride[t];
            const int64_t rb = tab_row_base[t];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int j = s;
            for (; j + 4 <= e; j += 4) {                         // 4 rows in flight per group
                float4 v0 = ldg_nc_f4(base + (int64_t)(e_emb[j] - rb) * stride);
                float4 v1 = ldg_nc_f4(base + (int64_t)(e_emb[j + 1] - rb) * stride);
                float4 v2 = ldg_nc_f4(base + (int64_t)(e_emb[j + 2] - rb) * stride);
                float4 v3 = ldg_nc_f4(base + (int64_t)(e_emb[j + 3] - rb) * stride);
                acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
            }
            for (; j < e; ++j) {
                float4 v = ldg_nc_f4(base + (int64_t)(e_emb[j] - rb) * stride);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            int n = e - s;
            if (n > 1) {
                float inv = 1.f / (float)n;                      // combiner='mean'
                acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            }
            *reinterpret_cast<float4*>(X0 + (int64_t)b * ld + tab_x0[t] + lig * 4) = acc;
        }
    } else {
        for (int64_t bag = gwarp; bag < nbags; bag += nwarp) {
            int b = (int)(bag / ntab), t = tab_ids[bag % ntab];
            int c = tab_col[t];
            int s = offs[(int64_t)b * C + c], e = offs[(int64_t)b * C + c + 1];
            const float* base = tab_data[t] + lig * 4;
            const int stride = tab_stride[t];
            const int64_t rb = tab_row_base[t];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            // 32 ids of the bag per round trip (one per lane), then the rows of the chunk that belong to this lane group eight at a
            // time back to back: the chain offsets -> ids -> rows is 2 + ceil(rows / (8 GROUPS)) dependent round trips per 32 ids
            // instead of one per four rows.  (Sixteen in flight was measured slower: 120 registers -> 2 blocks per SM -> 3.5 waves
            // of the 8192 bags, 25.6 us against 20 us.)  Group grp still sums rows grp, grp + GROUPS, ... in that order (the result
            // is bit-identical to the sequential walk).
            constexpr int STEPS = 32 / GROUPS, RND = STEPS < 8 ? STEPS : 8;
            for (int j0 = s; j0 < e; j0 += 32) {
                const int cnt = min(32, e - j0);
                const uint32_t my = lane < cnt ? e_emb[j0 + lane] : 0u;
#pragma unroll
                for (int k0 = 0; k0 < STEPS; k0 += RND) {
                    if (k0 * GROUPS >= cnt) break;
                    float4 v[RND];
#pragma unroll
                    for (int u = 0; u < RND; ++u) {
                        const int r = (k0 + u) * GROUPS + grp;
                        const uint32_t id = __shfl_sync(0xffffffffu, my, r & 31);
                        v[u] = r < cnt ? ldg_nc_f4(base + (int64_t)(id - rb) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < RND; ++u) {
                        if ((k0 + u) * GROUPS + grp < cnt) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
                    }
                }
            }
#pragma unroll
            for (int d = G; d < 32; d <<= 1) {                  // segmented (per lane-in-group) shuffle reduction
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, d);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, d);
            }
            int n = e - s;
            if (n > 1) {
                float inv = 1.f / (float)n;
                acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            }
            if (grp == 0) *reinterpret_cast<float4*>(X0 + (int64_t)b * ld + tab_x0[t] + lig * 4) = acc;
        }
    }
}

// Warp per example (short bags).  The warp walks the example's bags of this width in rounds of 32/G bags and
// issues the loads of RMAX rounds back to back: offsets, then first ids, then first rows — every lane group keeps
// RMAX independent 16-byte row loads in flight instead of one dependent chain at a time.
template <int G, int RMAX>
__global__ void __launch_bounds__(256, 4) emb_pool_fwd_rows_kernel(int B, int C, int ntab, const TabDesc* __restrict__ desc,
                                                                const int32_t* __restrict__ offs, const uint32_t* __restrict__ e_emb,
                                                                float* __restrict__ X0, int ld) {
    constexpr int GROUPS = 32 / G;
    const int lane = threadIdx.x & 31, lig = lane % G, grp = lane / G;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int b = gwarp; b < B; b += nwarp) {
        const int32_t* orow = offs + (int64_t)b * C;
        float* xrow = X0 + (int64_t)b * ld;
        for (int k0 = 0; k0 < ntab; k0 += GROUPS * RMAX) {
            TabDesc d[RMAX];
            int s[RMAX], n[RMAX];
            float4 acc[RMAX];
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int k = k0 + r * GROUPS + grp;
                n[r] = -1;
                if (k < ntab) {
                    d[r] = desc[k];
                    s[r] = orow[d[r].col];
                    n[r] = orow[d[r].col + 1] - s[r];
                }
            }
            uint32_t id0[RMAX];
#pragma unroll
            for (int r = 0; r < RMAX; ++r) id0[r] = n[r] > 0 ? e_emb[s[r]] : 0u;
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n[r] > 0) acc[r] = ldg_nc_f4(d[r].data + (int64_t)(id0[r] - d[r].row_base) * d[r].stride + lig * 4);
            }
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                if (n[r] > 1) {                                  // multihot bag: remaining rows, then the mean
                    for (int j = 1; j < n[r]; ++j) {
                        float4 v = ldg_nc_f4(d[r].data + (int64_t)(e_emb[s[r] + j] - d[r].row_base) * d[r].stride + lig * 4);
                        acc[r].x += v.x; acc[r].y += v.y; acc[r].z += v.z; acc[r].w += v.w;
                    }
                    const float inv = 1.f / (float)n[r];
                    acc[r].x *= inv; acc[r].y *= inv; acc[r].z *= inv; acc[r].w *= inv;
                }
                if (n[r] >= 0) *reinterpret_cast<float4*>(xrow + d[r].x0 + lig * 4) = acc[r];
            }
        }
    }
}

// --------------------------------------------------------------------------------- TMA-staged gather (short bags)
// Warp per example, rows staged in shared memory by 1-D TMA bulk copies (cp.async.bulk.shared::cluster.global with
// mbarrier complete_tx): every lane issues the copies of "its" table's rows, so a warp has the whole example (e.g.
// 26 x 128 B) in flight with a handful of instructions and no registers tied up; a ring of NBUF staging buffers per
// warp keeps several examples in flight.  Single-valued bags whose columns are adjacent in the deep input (the
// Criteo shape) are written back with ONE bulk store (cp.async.bulk.global.shared::cta) straight from the staging
// buffer; otherwise lane groups pool the staged rows (mean) and store float4s.
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sm_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra W_DONE;\n\t"
        "bra W_LOOP;\n\t"
        "W_DONE:\n\t}" ::"r"(sm_u32(bar)), "r"(parity) : "memory");
}

template <int G>
__global__ void __launch_bounds__(256) emb_pool_fwd_tma_kernel(int B, int C, int ntab, const TabDesc* __restrict__ desc,
                                                               const int32_t* __restrict__ offs, const uint32_t* __restrict__ e_emb,
                                                               float* __restrict__ X0, int ld, int buf_bytes, int nbuf, int contiguous) {
    extern __shared__ __align__(128) uint8_t tsm[];
    constexpr int ROWB = G * 16;                         // bytes per embedding row
    constexpr int GROUPS = 32 / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slots = buf_bytes / ROWB;
    uint8_t* wbase = tsm + (size_t)warp * nbuf * buf_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + (size_t)8 * nbuf * buf_bytes) + warp * 8;      // up to 8 buffers per warp
    if (lane == 0)
        for (int i = 0; i < nbuf; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int nex = gwarp < B ? (B - gwarp + nwarp - 1) / nwarp : 0;   // examples of this warp

    // per-lane state of an in-flight example (tables lane, lane+32, ... only the first 32 tables use the TMA path)
    auto issue = [&](int i) -> int {                 // stage example i of this warp; returns total entries (or -1: too many)
        const int b = gwarp + i * nwarp;
        const int buf = i % nbuf;
        const int32_t* orow = offs + (int64_t)b * C;
        int s = 0, n = 0;
        TabDesc d{};
        if (lane < ntab) { d = desc[lane]; s = orow[d.col]; n = orow[d.col + 1] - s; }
        int pos = n;                                   // exclusive prefix of n over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, pos, o); if (lane >= o) pos += t; }
        const int total = __shfl_sync(0xffffffffu, pos, 31);
        pos -= n;
        if (total > slots) {                           // does not fit the staging buffer: keep the barrier phase in step, drain with direct loads
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(&bars[buf])), "r"(0u) : "memory");
            return -1;
        }
        uint8_t* bbase = wbase + (size_t)buf * buf_bytes;
        for (int j = 0; j < n; ++j) {
            const float* src = d.data + (int64_t)(e_emb[s + j] - d.row_base) * d.stride;
            bulk_g2s(bbase + (size_t)(pos + j) * ROWB, src, ROWB, &bars[buf]);
        }
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(&bars[buf])), "r"((uint32_t)(total * ROWB)) : "memory");
        return total;
    };

    const int depth = nbuf - 1 > 0 ? nbuf - 1 : 1;      // examples in flight ahead of the one being drained
    for (int i = 0; i < nex && i < depth; ++i) {
        if (issue(i) < 0) { /* handled when drained (falls back to direct loads) */ }
    }
    for (int i = 0; i < nex; ++i) {
        const int b = gwarp + i * nwarp;
        const int buf = i % nbuf;
        const int32_t* orow = offs + (int64_t)b * C;
        float* xrow = X0 + (int64_t)b * ld;
        // recompute this example's layout (cheap, L1/L2 hits) instead of carrying it across the pipeline
        int s = 0, n = 0;
        TabDesc d{};
        if (lane < ntab) { d = desc[lane]; s = orow[d.col]; n = orow[d.col + 1] - s; }
        int pos = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, pos, o); if (lane >= o) pos += t; }
        const int total = __shfl_sync(0xffffffffu, pos, 31);
        pos -= n;
        const bool staged = total <= slots;
        uint8_t* bbase = wbase + (size_t)buf * buf_bytes;
        mbar_wait_parity(&bars[buf], (uint32_t)((i / nbuf) & 1));
        const bool all_single = __all_sync(0xffffffffu, lane >= ntab || n == 1);
        if (staged && all_single && contiguous) {
            // the staging buffer already is the example's slice of the deep input
            if (lane == 0) {
                bulk_s2g(xrow + desc[0].x0, bbase, (uint32_t)(ntab * ROWB));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            const int lig = lane % G, grp = lane / G;
            for (int k0 = 0; k0 < ntab; k0 += GROUPS) {
                const int k = k0 + grp;
                const int kk = k < ntab ? k : 0;
                const int pk = __shfl_sync(0xffffffffu, pos, kk), nk = __shfl_sync(0xffffffffu, n, kk), sk = __shfl_sync(0xffffffffu, s, kk);
                if (k < ntab) {
                    const TabDesc dk = desc[k];
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = 0; j < nk; ++j) {
                        float4 v = staged ? *reinterpret_cast<const float4*>(bbase + (size_t)(pk + j) * ROWB + lig * 16)
                                          : ldg_nc_f4(dk.data + (int64_t)(e_emb[sk + j] - dk.row_base) * dk.stride + lig * 4);
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    }
                    if (nk > 1) { const float inv = 1.f / (float)nk; acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
                    *reinterpret_cast<float4*>(xrow + dk.x0 + lig * 4) = acc;
                }
            }
        }
        __syncwarp();
        // refill: example i + depth goes into buffer (i + depth) % nbuf == the buffer drained one iteration ago (or this
        // one when nbuf == depth + 1 ... ) -> wait until its bulk store has finished reading shared memory
        const int nxt = i + depth;
        if (nxt < nex) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            issue(nxt);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int G>
static void launch_emb_fwd(WdModel* m, int di, bool widebag) {
    int ntab = m->dim_ntables[di];
    static const int gather_mode = getenv("WD_GATHER") ? atoi(getenv("WD_GATHER")) : 0;     // 0: LDG kernel (default: measured faster), 1: TMA-staged kernel
    if (!widebag && gather_mode == 1 && ntab <= 32) {
        const int rowb = G * 16;
        int buf_bytes = ((ntab * rowb * 5 / 4 + 1023) / 1024) * 1024;             // 25% head-room for multihot bags (larger ones use direct loads)
        if (buf_bytes < 1024) buf_bytes = 1024;
        int nbuf = 2;                                  // two staging buffers per warp: ~3 blocks (24 warps) per SM at 5 KB buffers
        if (nbuf >= 2) {
            // tables of this width adjacent in the deep input, in descriptor order?  (host check once per model would do; cheap here)
            int contiguous = 1;
            int prev = -1;
            for (auto& tb : m->tables)
                if (tb.dim == m->dims[di]) {
                    if (prev >= 0 && tb.x0_off != prev + tb.dim) contiguous = 0;
                    prev = tb.x0_off;
                }
            size_t smem = (size_t)8 * nbuf * buf_bytes + 8 * 8 * sizeof(uint64_t);
            static bool configured[wd::kMaxDims] = {false};
            if (!configured[di]) {
                cudaFuncSetAttribute(emb_pool_fwd_tma_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                configured[di] = true;
            }
            int grid = grid_for((int64_t)m->dbatch.B * 32, 256, 148 * 3);
            emb_pool_fwd_tma_kernel<G><<<grid, 256, smem, m->stream>>>(m->dbatch.B, m->n_columns, ntab, m->d_dim_desc[di], m->d_col_offs, m->d_e_emb,
                                                                   m->d_X0, m->d0_phys, buf_bytes, nbuf, contiguous);
            m->launches++;
            return;
        }
    }
    if (!widebag) {
        constexpr int RMAX = 4;                       // 4 rounds in flight per lane group at <= 64 registers: 32 warps per SM
        int grid = grid_for((int64_t)m->dbatch.B * 32, 256, 148 * 8);
        emb_pool_fwd_rows_kernel<G, RMAX><<<grid, 256, 0, m->stream>>>(m->dbatch.B, m->n_columns, ntab, m->d_dim_desc[di], m->d_col_offs,
                                                                        m->d_e_emb, m->d_X0, m->d0_phys);
        m->launches++;
        return;
    }
    int64_t nbags = (int64_t)m->dbatch.B * ntab;
    int64_t warps = widebag ? nbags : (nbags + (32 / G) - 1) / (32 / G);
    int grid = grid_for(warps * 32, 256, 148 * 8);
    if (widebag)
        emb_pool_fwd_kernel<G, true><<<grid, 256, 0, m->stream>>>(m->dbatch.B, m->n_columns, ntab, m->d_dim_tables[di],
            m->d_tab_data, m->d_tab_stride, m->d_tab_x0, m->d_tab_col, m->d_tab_row_base, m->d_col_offs, m->d_e_emb, m->d_X0, m->d0_phys);
    else
        emb_pool_fwd_kernel<G, false><<<grid, 256, 0, m->stream>>>(m->dbatch.B, m->n_columns, ntab, m->d_dim_tables[di],
            m->d_tab_data, m->d_tab_stride, m->d_tab_x0, m->d_tab_col, m->d_tab_row_base, m->d_col_offs, m->d_e_emb, m->d_X0, m->d0_phys);
    m->launches++;
}

// the wide half of the forward (independent of the deep half until the head adds the logits)
int sparse_forward_wide(WdModel* m) {
    const int B = m->dbatch.B;
    if (m->use_wide) {
        wide_fwd_kernel<<<grid_for((int64_t)B * 32, 256), 256, 0, m->stream>>>(B, m->n_columns, m->d_col_offs, m->d_e_wide,
                                                                              m->d_wide, m->d_P + m->dense[0].off, m->d_wide_logit);
        m->launches++;
        mark(m, "wide_fwd");
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
int sparse_forward_emb(WdModel* m);
int sparse_forward(WdModel* m) {
    int rc = sparse_forward_wide(m);
    return rc ? rc : sparse_forward_emb(m);
}
int sparse_forward_emb(WdModel* m) {
    if (m->use_deep) {
        // heuristic: average bag length from the key count of the batch decides narrow vs full-warp bags
        bool widebag = m->max_nnz > 0 && (m->keys_cap / (int64_t)(m->max_batch * (m->n_cat_fields > 0 ? m->n_cat_fields : 1))) >= 8;
        for (int di = 0; di < m->n_dims; ++di) {
            switch (m->dims[di] / 4) {
                case 1: launch_emb_fwd<1>(m, di, false); break;
                case 2: launch_emb_fwd<2>(m, di, widebag); break;
                case 4: launch_emb_fwd<4>(m, di, widebag); break;
                case 8: launch_emb_fwd<8>(m, di, widebag); break;
                case 16: launch_emb_fwd<16>(m, di, widebag); break;
                case 32: launch_emb_fwd<32>(m, di, true); break;
                default: set_error("unsupported embedding width %d (supported: 4,8,16,32,64,128)", m->dims[di]); return WD_EUNSUPPORTED;
            }
        }
        mark(m, "emb_fwd");
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ------------------------------------------------------------------------------------ backward: grouping
// (key, value) pairs for the two sorts (which = 0 embedding rows, 1 wide rows): key = row — non-participating entries get the key
// `invalid` (= 1 << bits), so they sort behind every real row —, value = val_src[i], or the entry's index i without val_src.
// The batch's own lists carry the entry's cell index bc = b * C + c (e_bc) as the value: that is all the gradient sums need to
// find an occurrence's gradient, so they read it straight from the sorted list instead of chasing e_bc[index] per occurrence.
__global__ void sort_keys_kernel(const int32_t* __restrict__ d_nnz, const uint32_t* __restrict__ e_row, uint32_t invalid,
                                 uint32_t* keys, uint32_t* vals, const int32_t* __restrict__ val_src) {
    int n = *d_nnz;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t r = e_row[i];
        keys[i] = r == kInvalidRow ? invalid : r;
        vals[i] = val_src ? (uint32_t)val_src[i] : (uint32_t)i;
    }
}


// ---- per unique row: g = ordered sum of its occurrences' gradients.
// Rows touched at most kChunk times are summed by one lane group directly.  Hotter rows (small tables,
// skewed ids) are split into chunks of kChunk occurrences that are summed in parallel and then combined in
// chunk order, so the result stays deterministic and no single group walks thousands of occurrences.

// embedding rows: contribution of occurrence j = dX0[b, x0_off : x0_off + dim] / bag_size(b, column)
// 8 lanes per work item, each lane covers float4 chunks lig, lig+8, ... of the row.  One launch covers both kinds of work item:
// items [0, nu) are the unique rows (summed directly into ugrad unless they are hot), items [nu, nu + nchunks) are the chunks of
// the hot rows (summed into cpart, combined afterwards by chunk_combine_kernel).
// APPLY (single-GPU step, row-local optimizer): the optimizer update of a directly summed row follows its sum in the same thread —
// the summed gradient never reaches memory; the hot rows are updated by chunk_combine_kernel<1>.
struct RowApply { const uint32_t* urow; float* const* tab_data; const int32_t* tab_stride; const int64_t* tab_row_base; OptParams o; };
template <bool APPLY>
__global__ void __launch_bounds__(256) emb_grad_sum_kernel(const int32_t* __restrict__ d_nuniq, const int32_t* __restrict__ d_nchunks,
                                                           const int32_t* __restrict__ ustart, const int32_t* __restrict__ choff,
                                                           const uint32_t* __restrict__ sbc /* sorted cell indices bc = b * C + c */,
                                                           const int32_t* __restrict__ offs,
                                                           int C, const int32_t* __restrict__ col_table,
                                                           const int32_t* __restrict__ tab_dim, const int32_t* __restrict__ tab_x0,
                                                           const float* __restrict__ dX0, int ld, float* __restrict__ ugrad,
                                                           float* __restrict__ cpart, int width, RowApply ra) {
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    const int nu = *d_nuniq;
    const int64_t nitems = (int64_t)nu + *d_nchunks;
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t it = g0; it < nitems; it += gstep) {
        int s, e;
        float* out;
        const bool direct = it < nu;
        if (direct) {
            s = ustart[it]; e = ustart[it + 1];
            if (e - s > kChunk) continue;                     // hot row: summed chunk by chunk below
            out = ugrad + it * width;
        } else {
            const int c = (int)(it - nu);
            const int u = chunk_owner(choff, nu, c);
            s = ustart[u] + (c - choff[u]) * kChunk;
            e = min(ustart[u + 1], s + kChunk);
            out = cpart + (int64_t)c * width;
        }
        int bc0 = (int)sbc[s];
        int t = col_table[bc0 % C];
        int dim = tab_dim[t], x0 = tab_x0[t];
        for (int q = lig; q * 4 < width; q += 8) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < dim) {
                int j = s;
                for (; j + 4 <= e; j += 4) {                  // 4 gradient rows in flight
                    int bcs[4]; float4 v[4]; float inv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) bcs[r] = (int)sbc[j + r];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        v[r] = *reinterpret_cast<const float4*>(dX0 + (int64_t)(bcs[r] / C) * ld + x0 + q * 4);
                        inv[r] = 1.f / (float)(offs[bcs[r] + 1] - offs[bcs[r]]);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) { acc.x += v[r].x * inv[r]; acc.y += v[r].y * inv[r]; acc.z += v[r].z * inv[r]; acc.w += v[r].w * inv[r]; }
                }
                for (; j < e; ++j) {
                    int bc = (int)sbc[j];
                    float4 v = *reinterpret_cast<const float4*>(dX0 + (int64_t)(bc / C) * ld + x0 + q * 4);
                    float inv = 1.f / (float)(offs[bc + 1] - offs[bc]);
                    acc.x += v.x * inv; acc.y += v.y * inv; acc.z += v.z * inv; acc.w += v.w * inv;
                }
            }
            if (APPLY && direct) {
                if (q * 4 < dim) {
                    const int stride = ra.tab_stride[t];
                    float* rec = ra.tab_data[t] + ((int64_t)ra.urow[it] - ra.tab_row_base[t]) * stride;
                    const int nslots = stride / dim - 1;
                    float4 w = *reinterpret_cast<float4*>(rec + q * 4);
                    float4 s1 = nslots >= 1 ? *reinterpret_cast<float4*>(rec + dim + q * 4) : make_float4(0, 0, 0, 0);
                    float4 s2 = nslots >= 2 ? *reinterpret_cast<float4*>(rec + 2 * dim + q * 4) : make_float4(0, 0, 0, 0);
                    opt_update(ra.o, acc.x, w.x, s1.x, s2.x);
                    opt_update(ra.o, acc.y, w.y, s1.y, s2.y);
                    opt_update(ra.o, acc.z, w.z, s1.z, s2.z);
                    opt_update(ra.o, acc.w, w.w, s1.w, s2.w);
                    *reinterpret_cast<float4*>(rec + q * 4) = w;
                    if (nslots >= 1) *reinterpret_cast<float4*>(rec + dim + q * 4) = s1;
                    if (nslots >= 2) *reinterpret_cast<float4*>(rec + 2 * dim + q * 4) = s2;
                }
            } else {
                *reinterpret_cast<float4*>(out + q * 4) = acc;
            }
        }
    }
}

// ugrad[u] = sum of the row's chunk partials (multi-chunk rows only).  Each lane checks one unique row; the (rare)
// multi-chunk rows of a warp are combined by the whole warp: lane groups add chunks g, g+NG, ... and a fixed-order shuffle
// tree adds the group sums, so the result does not depend on scheduling.  Hot rows are neighbours in row order (they are the
// rows of the small tables), so a warp takes every NW-th group of rows (lane l of warp w checks row (it * 32 + l) * NW + w):
// neighbouring hot rows land in different warps and their long chunk lists are walked concurrently, not one after the other.
// KIND 0: the sum goes to ugrad.  KIND 1 / 2 (single-GPU step, row-local optimizer): the hot row's optimizer update follows its
// sum here (1: embedding record, width a power of two in [4, 128]; 2: wide record, width 1) — together with the APPLY variants of
// the gradient-sum kernels this leaves no separate optimizer launch for the list.
struct HotApply {
    const uint32_t* urow; int ntab; const int64_t* tab_row_base; float* const* tab_data; const int32_t* tab_dim; const int32_t* tab_stride;   // KIND 1 (tables in row order)
    float4* wide;                                                                                                                          // KIND 2
    OptParams o;
};
template <int KIND>
__global__ void __launch_bounds__(256) chunk_combine_kernel(const int32_t* __restrict__ d_nuniq, const int32_t* __restrict__ choff,
                                                            const float* __restrict__ cpart, float* __restrict__ ugrad, int width, HotApply ha) {
    const int nu = *d_nuniq;
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = 0; it * 32 * nw < nu; ++it) {
        const int64_t u = (it * 32 + lane) * nw + w0;
        int c0 = 0, c1 = 0;
        if (u < nu) { c0 = choff[u]; c1 = choff[u + 1]; }
        unsigned multi = __ballot_sync(0xffffffffu, c1 > c0);
        while (multi) {
            const int src = __ffs(multi) - 1;
            multi &= multi - 1;
            const int b0 = __shfl_sync(0xffffffffu, c0, src), b1 = __shfl_sync(0xffffffffu, c1, src);
            const int64_t uu = (it * 32 + src) * nw + w0;          // the unique row being combined
            const int G = width >> 2;
            if (width >= 4 && (G & (G - 1)) == 0 && G <= 32) {
                // G lanes cover one chunk's row (float4 each), 32/G chunks in flight per step, 4 steps unrolled;
                // fixed-order tree over the chunk groups
                const int lq = lane % G, cg = lane / G, NG = 32 / G;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int c = b0 + cg;
                for (; c + 3 * NG < b1; c += 4 * NG) {
                    const float4 v0 = *reinterpret_cast<const float4*>(cpart + (int64_t)c * width + lq * 4);
                    const float4 v1 = *reinterpret_cast<const float4*>(cpart + (int64_t)(c + NG) * width + lq * 4);
                    const float4 v2 = *reinterpret_cast<const float4*>(cpart + (int64_t)(c + 2 * NG) * width + lq * 4);
                    const float4 v3 = *reinterpret_cast<const float4*>(cpart + (int64_t)(c + 3 * NG) * width + lq * 4);
                    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                    acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                    acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
                }
                for (; c < b1; c += NG) {
                    const float4 v = *reinterpret_cast<const float4*>(cpart + (int64_t)c * width + lq * 4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                for (int d = G; d < 32; d <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, d); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, d);
                }
                if (KIND == 1) {
                    if (cg == 0) {
                        const int64_t row = ha.urow[uu];
                        int lo = 0, hi = ha.ntab - 1;               // table of this global row (tables are few: binary search)
                        while (lo < hi) {
                            int mid = (lo + hi + 1) >> 1;
                            if (ha.tab_row_base[mid] <= row) lo = mid; else hi = mid - 1;
                        }
                        const int dim = ha.tab_dim[lo], stride = ha.tab_stride[lo];
                        if (lq * 4 < dim) {
                            float* rec = ha.tab_data[lo] + (row - ha.tab_row_base[lo]) * stride;
                            const int nslots = stride / dim - 1;
                            float4 w = *reinterpret_cast<float4*>(rec + lq * 4);
                            float4 s1 = nslots >= 1 ? *reinterpret_cast<float4*>(rec + dim + lq * 4) : make_float4(0, 0, 0, 0);
                            float4 s2 = nslots >= 2 ? *reinterpret_cast<float4*>(rec + 2 * dim + lq * 4) : make_float4(0, 0, 0, 0);
                            opt_update(ha.o, acc.x, w.x, s1.x, s2.x);
                            opt_update(ha.o, acc.y, w.y, s1.y, s2.y);
                            opt_update(ha.o, acc.z, w.z, s1.z, s2.z);
                            opt_update(ha.o, acc.w, w.w, s1.w, s2.w);
                            *reinterpret_cast<float4*>(rec + lq * 4) = w;
                            if (nslots >= 1) *reinterpret_cast<float4*>(rec + dim + lq * 4) = s1;
                            if (nslots >= 2) *reinterpret_cast<float4*>(rec + 2 * dim + lq * 4) = s2;
                        }
                    }
                } else if (cg == 0) *reinterpret_cast<float4*>(ugrad + uu * width + lq * 4) = acc;
            } else {
                for (int q = 0; q < width; ++q) {
                    float acc = 0.f;
                    for (int c = b0 + lane; c < b1; c += 32) acc += cpart[(int64_t)c * width + q];
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    if (lane == 0) {
                        if (KIND == 2) {                            // width == 1
                            float4 r = ha.wide[ha.urow[uu]];
                            opt_update(ha.o, acc, r.x, r.y, r.z);
                            ha.wide[ha.urow[uu]] = r;
                        } else ugrad[uu * width + q] = acc;
                    }
                }
            }
        }
    }
}

// wide rows: contribution of occurrence j = dlogit[b]; one thread per work item.  Items [0, nu) = unique rows (hot ones skipped),
// items [nu, nu + nchunks) = chunks of the hot rows, as in emb_grad_sum_kernel; APPLY: record {w, s1, s2, -} updated in place.
template <bool APPLY>
__global__ void wide_grad_sum_kernel(const int32_t* __restrict__ d_nuniq, const int32_t* __restrict__ d_nchunks,
                                     const int32_t* __restrict__ ustart, const int32_t* __restrict__ choff,
                                     const uint32_t* __restrict__ sbc /* sorted cell indices bc = b * C + c */, int C,
                                     const float* __restrict__ dlogit, float* __restrict__ ugrad, float* __restrict__ cpart,
                                     const uint32_t* __restrict__ urow, float4* __restrict__ wide, OptParams o) {
    const int nu = *d_nuniq;
    const int64_t nitems = (int64_t)nu + *d_nchunks;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += (int64_t)gridDim.x * blockDim.x) {
        int s, e;
        const bool direct = it < nu;
        if (direct) {
            s = ustart[it]; e = ustart[it + 1];
            if (e - s > kChunk) continue;
        } else {
            const int c = (int)(it - nu);
            const int u = chunk_owner(choff, nu, c);
            s = ustart[u] + (c - choff[u]) * kChunk;
            e = min(ustart[u + 1], s + kChunk);
        }
        float acc = 0.f;
        for (int j = s; j < e; ++j) acc += dlogit[sbc[j] / (uint32_t)C];
        if (!direct) cpart[it - nu] = acc;
        else if (APPLY) {
            float4 r = wide[urow[it]];
            opt_update(o, acc, r.x, r.y, r.z);
            wide[urow[it]] = r;
        } else ugrad[it] = acc;
    }
}

// --------------------------------------------------------------------------------------------- optimizers

// embedding rows: record = [w[dim] | s1[dim] | s2[dim]]
__global__ void __launch_bounds__(256) emb_apply_kernel(const int32_t* __restrict__ d_nuniq, const uint32_t* __restrict__ urow,
                                                        const float* __restrict__ ugrad, int width, int ntab,
                                                        const int64_t* __restrict__ tab_row_base, float* const* __restrict__ tab_data,
                                                        const int32_t* __restrict__ tab_dim, const int32_t* __restrict__ tab_stride,
                                                        OptParams o) {
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    const int nu = *d_nuniq;
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t u = g0; u < nu; u += gstep) {
        int64_t row = urow[u];
        int lo = 0, hi = ntab - 1;                  // table of this global row (tables are few: binary search)
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (tab_row_base[mid] <= row) lo = mid; else hi = mid - 1;
        }
        const int dim = tab_dim[lo], stride = tab_stride[lo];
        float* rec = tab_data[lo] + (row - tab_row_base[lo]) * stride;
        const int nslots = stride / dim - 1;
        for (int q = lig; q * 4 < dim; q += 8) {
            float4 g = *reinterpret_cast<const float4*>(ugrad + (int64_t)u * width + q * 4);
            float4 w = *reinterpret_cast<float4*>(rec + q * 4);
            float4 s1 = nslots >= 1 ? *reinterpret_cast<float4*>(rec + dim + q * 4) : make_float4(0, 0, 0, 0);
            float4 s2 = nslots >= 2 ? *reinterpret_cast<float4*>(rec + 2 * dim + q * 4) : make_float4(0, 0, 0, 0);
            opt_update(o, g.x, w.x, s1.x, s2.x);
            opt_update(o, g.y, w.y, s1.y, s2.y);
            opt_update(o, g.z, w.z, s1.z, s2.z);
            opt_update(o, g.w, w.w, s1.w, s2.w);
            *reinterpret_cast<float4*>(rec + q * 4) = w;
            if (nslots >= 1) *reinterpret_cast<float4*>(rec + dim + q * 4) = s1;
            if (nslots >= 2) *reinterpret_cast<float4*>(rec + 2 * dim + q * 4) = s2;
        }
    }
}

// wide rows: record {w, s1, s2, -}; also the bias record (dense gradient = sum of dlogit)
__global__ void wide_apply_kernel(const int32_t* __restrict__ d_nuniq, const uint32_t* __restrict__ urow,
                                  const float* __restrict__ ugrad, float4* __restrict__ wide, OptParams o) {
    const int nu = *d_nuniq;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
        float4 r = wide[urow[u]];
        opt_update(o, ugrad[u], r.x, r.y, r.z);
        wide[urow[u]] = r;
    }
}


// sort (row, occurrence) pairs by row and find the unique rows; e_row: per-occurrence row ids (kInvalidRow = skip)
static int group_tail(WdModel* m, int which, const int32_t* d_n);
static int group_rows(WdModel* m, int which, const int32_t* d_n, const uint32_t* e_row, const int32_t* val_src = nullptr) {
    const uint32_t invalid = 1u << m->sort_bits[which];
    int g = grid_for(m->max_nnz, 256);
    sort_keys_kernel<<<g, 256, 0, m->stream>>>(d_n, e_row, invalid, m->d_sk[which], m->d_sv[which], val_src);
    m->launches++;
    int rc = radix_sort_pairs(m, which, m->sort_bits[which] + 1, d_n);
    if (rc) return rc;
    return group_tail(m, which, d_n);
}
// unique rows + segment starts of the sorted (row, occurrence) pairs in d_sk / d_sv
static int group_tail(WdModel* m, int which, const int32_t* d_n) {
    const uint32_t invalid = 1u << m->sort_bits[which];
    int32_t* pos = (int32_t*)m->d_sk2[which];                 // ping-pong buffer is free after the sort
    return seg_heads(m, d_n, m->d_sk[which], invalid, pos, m->max_nnz, m->d_ustart[which], m->d_urow[which], m->d_nuniq[which]);
}

// out[u] = sum over the segment of in[sv[j]] (rows of `width` floats); one thread per (unique row, float4 chunk)
__global__ void merged_sum_kernel(const int32_t* __restrict__ d_nuniq, const int32_t* __restrict__ ustart, const uint32_t* __restrict__ svals,
                                  const float* __restrict__ in, float* __restrict__ out, int width) {
    const int nu = *d_nuniq;
    const int64_t total = (int64_t)nu * width;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int u = (int)(t / width), q = (int)(t % width);
        float acc = 0.f;
        for (int j = ustart[u]; j < ustart[u + 1]; ++j) acc += in[(int64_t)svals[j] * width + q];
        out[t] = acc;
    }
}

__global__ void set_count_kernel(int32_t* dst, int32_t v) { *dst = v; }

// Replace the gradient list `which` by the row-wise sum of an external (rows, grads) list, e.g. the
// all-gathered lists of every rank: keeps "sum duplicates, apply once" across data-parallel replicas.
int merge_sparse(WdModel* m, int which, const void* rows, const void* grads, int64_t n) {
    if (n > m->max_nnz) { set_error("merged sparse list has %lld rows, capacity %lld", (long long)n, (long long)m->max_nnz); return WD_EINVAL; }
    set_count_kernel<<<1, 1, 0, m->stream>>>(m->d_nvalid[which], (int32_t)n);     // (a kernel, not a memcpy: capturable into a graph)
    m->launches++;
    int rc = group_rows(m, which, m->d_nvalid[which], (const uint32_t*)rows);
    if (rc) return rc;
    const int width = which == 0 ? m->emb_max_dim : 1;
    merged_sum_kernel<<<grid_for(std::max<int64_t>(n, 1) * width, 256), 256, 0, m->stream>>>(m->d_nuniq[which], m->d_ustart[which], m->d_sv[which],
                                                                                           (const float*)grads, m->d_ugrad[which], width);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// Merge of G lists that are each sorted ascending, duplicate-free and padded with kInvalidRow (what wd_sparse_grads hands out, so
// what a fixed-size all-gather of it yields): no sort — every element finds its position in the stable merged order with one
// binary search per list (own list: its index), one kernel; then the usual unique-row / segment pass and the ordered sums.
__global__ void __launch_bounds__(256) merge_rank_kernel(const uint32_t* __restrict__ rows, int G, int K, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals, int32_t* __restrict__ d_nvalid) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int n = 0;
        for (int r = 0; r < G; ++r) n += lower_bound_u32(rows + (int64_t)r * K, K, kInvalidRow);
        *d_nvalid = n;
    }
    const int64_t total = (int64_t)G * K;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t x = rows[e];
        if (x == kInvalidRow) continue;
        const int r = (int)(e / K);
        int pos = (int)(e - (int64_t)r * K);                        // earlier entries of the own list are all smaller
        for (int q = 0; q < G; ++q) {
            if (q == r) continue;
            const uint32_t* lst = rows + (int64_t)q * K;
            pos += q < r ? upper_bound_u32(lst, K, x) : lower_bound_u32(lst, K, x);   // ties: lower list index first (stable)
        }
        keys[pos] = x;
        vals[pos] = (uint32_t)e;
    }
}
// this rank touched more unique rows than the fixed exchange length: its tail was NOT exchanged -> fail the step loudly (flag 8)
__global__ void list_len_check_kernel(const int32_t* __restrict__ d_nuniq, int64_t list_len, int32_t* __restrict__ flags) {
    if ((int64_t)*d_nuniq > list_len) atomicOr(flags, 8);
}
int merge_sparse_sorted(WdModel* m, int which, const void* rows, const void* grads, int n_lists, int64_t list_len) {
    const int64_t n = (int64_t)n_lists * list_len;
    if (n_lists < 1 || list_len < 1 || list_len > 0x7fffffff) { set_error("merge: bad list shape"); return WD_EINVAL; }
    if (n > m->max_nnz) { set_error("merged sparse list has %lld rows, capacity %lld", (long long)n, (long long)m->max_nnz); return WD_EINVAL; }
    list_len_check_kernel<<<1, 1, 0, m->stream>>>(m->d_nuniq[which], list_len, m->d_flags);   // (d_nuniq still holds the local count)
    m->launches++;
    merge_rank_kernel<<<grid_for(n, 256), 256, 0, m->stream>>>((const uint32_t*)rows, n_lists, (int)list_len, m->d_sk[which], m->d_sv[which], m->d_nvalid[which]);
    m->launches++;
    int rc = group_tail(m, which, m->d_nvalid[which]);
    if (rc) return rc;
    const int width = which == 0 ? m->emb_max_dim : 1;
    merged_sum_kernel<<<grid_for(std::max<int64_t>(n, 1) * width, 256), 256, 0, m->stream>>>(m->d_nuniq[which], m->d_ustart[which], m->d_sv[which],
                                                                                           (const float*)grads, m->d_ugrad[which], width);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// sort + per-row gradient sums for both tables; leaves (urow, ugrad, nuniq) ready for exchange / apply
// Stage 1 of the sparse backward: group the step's (row, occurrence) pairs by row for both table spaces and lay out
// the hot-row chunks.  Depends only on the ids of the batch (not on any gradient), so api.cu runs it on a side stream
// concurrently with the towers' forward/backward.
int sparse_group_which(WdModel* m, int which) {
    int rc;
    const int g = grid_for(m->max_nnz, 256);
    const bool present = which == 0 ? (m->use_deep && !m->tables.empty()) : m->use_wide;
    if (present) {
        if ((rc = group_rows(m, which, m->d_nnz, which == 0 ? m->d_e_emb : m->d_e_wide, m->d_e_bc))) return rc;     // values = cell indices
        if ((rc = chunk_offsets(m, m->d_nuniq[which], m->d_ustart[which], m->d_urow[which], m->d_choff[which], m->max_nnz, kChunk, m->d_nchunks[which]))) return rc;
        mark(m, which == 0 ? "emb_group" : "wide_group");
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
int sparse_group(WdModel* m) {
    int rc = sparse_group_which(m, 0);
    return rc ? rc : sparse_group_which(m, 1);
}

// Stage 2: per-row gradient sums; leaves (urow, ugrad, nuniq) ready for exchange / apply.
// Single-GPU step (train_eager: nothing exchanges the sums between backward and optimizer) with a row-local optimizer (everything
// but Adam, whose moments decay over whole tables): the rows are updated inside these two launches and sparse_apply_which has
// nothing left to launch for the list (m->list_apply_fused).
static bool fuse_row_apply(const WdModel* m, const WdOptimizer& o) {
    static const bool no_fuse = getenv("WD_NO_FUSED_ROW_APPLY") != nullptr;              // A/B switch (bench only)
    return m->fuse_dense && o.kind != WD_OPT_ADAM && m->gs_count == 0 && !no_fuse;
}
int sparse_reduce_emb(WdModel* m) {
    const int g = grid_for(m->max_nnz, 256);
    if (m->use_deep && !m->tables.empty()) {
        const int width = m->emb_max_dim, G4 = width >> 2;
        const int ge = grid_for((m->max_nnz + m->cpart_cap) * 8, 256);
        // the hot rows' update lives in the lane-group branch of chunk_combine_kernel: widths 4, 8, ..., 128
        const bool fused = fuse_row_apply(m, m->dnn_opt) && width >= 4 && (G4 & (G4 - 1)) == 0 && G4 <= 32;
        const RowApply ra{m->d_urow[0], m->d_tab_data, m->d_tab_stride, m->d_tab_row_base, make_opt(m->dnn_opt)};
        const HotApply ha{m->d_urow[0], m->n_rtab, m->d_rtab_row_base, m->d_rtab_data, m->d_rtab_dim, m->d_rtab_stride, nullptr, make_opt(m->dnn_opt)};
        if (fused) {
            emb_grad_sum_kernel<true><<<ge, 256, 0, m->stream>>>(m->d_nuniq[0], m->d_nchunks[0], m->d_ustart[0], m->d_choff[0], m->d_sv[0],
                m->d_col_offs, m->n_columns, m->dplan.col_emb_table, m->d_tab_dim, m->d_tab_x0, m->d_dX0, m->d0_phys, m->d_ugrad[0], m->d_cpart[0], width, ra);
            chunk_combine_kernel<1><<<g, 256, 0, m->stream>>>(m->d_nuniq[0], m->d_choff[0], m->d_cpart[0], m->d_ugrad[0], width, ha);
        } else {
            emb_grad_sum_kernel<false><<<ge, 256, 0, m->stream>>>(m->d_nuniq[0], m->d_nchunks[0], m->d_ustart[0], m->d_choff[0], m->d_sv[0],
                m->d_col_offs, m->n_columns, m->dplan.col_emb_table, m->d_tab_dim, m->d_tab_x0, m->d_dX0, m->d0_phys, m->d_ugrad[0], m->d_cpart[0], width, ra);
            chunk_combine_kernel<0><<<g, 256, 0, m->stream>>>(m->d_nuniq[0], m->d_choff[0], m->d_cpart[0], m->d_ugrad[0], width, ha);
        }
        m->list_apply_fused[0] = fused;
        m->launches += 2;
        mark(m, "emb_grad_sum");
        m->sparse_overridden[0] = false;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int sparse_reduce_wide(WdModel* m) {
    const int g = grid_for(m->max_nnz, 256);
    if (m->use_wide) {
        const bool fused = fuse_row_apply(m, m->lin_opt);
        const HotApply ha{m->d_urow[1], 0, nullptr, nullptr, nullptr, nullptr, m->d_wide, make_opt(m->lin_opt)};
        const int gw = grid_for(m->max_nnz + m->cpart_cap, 256);
        if (fused) {
            wide_grad_sum_kernel<true><<<gw, 256, 0, m->stream>>>(m->d_nuniq[1], m->d_nchunks[1], m->d_ustart[1], m->d_choff[1], m->d_sv[1],
                                                                 m->n_columns, m->d_dlogit, m->d_ugrad[1], m->d_cpart[1], m->d_urow[1], m->d_wide, ha.o);
            chunk_combine_kernel<2><<<g, 256, 0, m->stream>>>(m->d_nuniq[1], m->d_choff[1], m->d_cpart[1], m->d_ugrad[1], 1, ha);
        } else {
            wide_grad_sum_kernel<false><<<gw, 256, 0, m->stream>>>(m->d_nuniq[1], m->d_nchunks[1], m->d_ustart[1], m->d_choff[1], m->d_sv[1],
                                                                  m->n_columns, m->d_dlogit, m->d_ugrad[1], m->d_cpart[1], m->d_urow[1], m->d_wide, ha.o);
            chunk_combine_kernel<0><<<g, 256, 0, m->stream>>>(m->d_nuniq[1], m->d_choff[1], m->d_cpart[1], m->d_ugrad[1], 1, ha);
        }
        m->list_apply_fused[1] = fused;
        m->launches += 2;
        mark(m, "wide_grad_sum");
        m->sparse_overridden[1] = false;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ------------------------------------------------------------------------- dense exchange of small tables
// Rows of the small tables sit at the end of the global row space, so they are the tail of the sorted unique-row list.  Their summed
// gradients move into the dense block behind the dense gradient arena (all-reduced with it in data-parallel runs) and leave the
// list: the tail is overwritten with kInvalidRow and the list length becomes the number of large-table rows.
__global__ void __launch_bounds__(256) small_scatter_emb_kernel(const int32_t* __restrict__ d_nuniq, uint32_t* __restrict__ urow,
                                                                const float* __restrict__ ugrad, int width, uint32_t small_base, int ntab,
                                                                const int64_t* __restrict__ rtab_row_base, const int32_t* __restrict__ rtab_dim,
                                                                const int64_t* __restrict__ rtab_gs_off, float* __restrict__ Gs,
                                                                float* __restrict__ touched, int32_t* __restrict__ d_nubig) {
    const int nu = *d_nuniq;
    const int lb = lower_bound_u32(urow, nu, small_base);        // (a concurrently invalidated tail entry still compares >= small_base)
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_nubig = lb;
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t u = lb + g0; u < nu; u += gstep) {
        const int64_t row = urow[u];
        int lo = 0, hi = ntab - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (rtab_row_base[mid] <= row) lo = mid; else hi = mid - 1; }
        const int dim = rtab_dim[lo];
        float* dst = Gs + rtab_gs_off[lo] + (row - rtab_row_base[lo]) * dim;
        for (int q = lig; q * 4 < dim; q += 8)
            *reinterpret_cast<float4*>(dst + q * 4) = *reinterpret_cast<const float4*>(ugrad + u * width + q * 4);
        __syncwarp();
        if (lig == 0) {
            touched[row - small_base] = 1.f;                      // "this rank touched the row" (summed over ranks with the block)
            urow[u] = kInvalidRow;
        }
    }
}
__global__ void __launch_bounds__(256) small_scatter_wide_kernel(const int32_t* __restrict__ d_nuniq, uint32_t* __restrict__ urow,
                                                                 const float* __restrict__ ugrad, uint32_t small_base, float* __restrict__ Gs,
                                                                 float* __restrict__ touched, int32_t* __restrict__ d_nubig) {
    const int nu = *d_nuniq;
    const int lb = lower_bound_u32(urow, nu, small_base);
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_nubig = lb;
    for (int64_t u = lb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += (int64_t)gridDim.x * blockDim.x) {
        Gs[urow[u] - small_base] = ugrad[u];
        touched[urow[u] - small_base] = 1.f;
        urow[u] = kInvalidRow;
    }
}
__global__ void copy_count_kernel(int32_t* dst, const int32_t* src) { *dst = *src; }

int small_scatter(WdModel* m, int which) {
    if (m->gs_count == 0) return WD_OK;
    float* block = m->d_G + m->dense_count;
    if (which == 0) {
        if (m->n_small_tab == 0 || !(m->use_deep && !m->tables.empty())) return WD_OK;
        small_scatter_emb_kernel<<<grid_for(m->max_nnz * 8, 256, 148 * 8), 256, 0, m->stream>>>(m->d_nuniq[0], m->d_urow[0], m->d_ugrad[0], m->emb_max_dim,
            (uint32_t)m->small_base[0], m->n_rtab, m->d_rtab_row_base, m->d_rtab_dim, m->d_rtab_gs_off, block,
            block + m->gs_touch_off[0], m->d_nubig[0]);
    } else {
        if (!m->use_wide || m->small_base[1] >= m->wide_rows) return WD_OK;
        small_scatter_wide_kernel<<<grid_for(m->max_nnz, 256, 148 * 8), 256, 0, m->stream>>>(m->d_nuniq[1], m->d_urow[1], m->d_ugrad[1],
            (uint32_t)m->small_base[1], block + m->gs_emb_floats, block + m->gs_touch_off[1], m->d_nubig[1]);
    }
    copy_count_kernel<<<1, 1, 0, m->stream>>>(m->d_nuniq[which], m->d_nubig[which]);
    m->launches += 2;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// optimizer over the dense block: one update per small-table row that ANY rank touched this step.  "Touched" travels with the
// block as a per-row count (summed by the same all-reduce), not as "gradient != 0": a touched row whose summed gradient is
// exactly zero (saturated sigmoids give dlogit == 0.0) still takes its optimizer step in TensorFlow — FTRL then rebuilds w from
// (z, n) and RMSProp decays its mean square.
__global__ void __launch_bounds__(256) small_apply_emb_kernel(const float* __restrict__ Gs, const float* __restrict__ touched, int64_t n4, int ntab,
                                                              int first_small, int64_t small_base, const int64_t* __restrict__ rtab_row_base,
                                                              const int64_t* __restrict__ rtab_gs_off, float* const* __restrict__ rtab_data,
                                                              const int32_t* __restrict__ rtab_dim, const int32_t* __restrict__ rtab_stride, OptParams o) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        int lo = first_small, hi = ntab - 1;                       // small tables are the last entries, offsets ascending
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (rtab_gs_off[mid] <= i * 4) lo = mid; else hi = mid - 1; }
        const int dim = rtab_dim[lo], stride = rtab_stride[lo];
        const int64_t local = i * 4 - rtab_gs_off[lo];
        if (touched[rtab_row_base[lo] - small_base + local / dim] == 0.f) continue;
        const float4 g = *reinterpret_cast<const float4*>(Gs + i * 4);
        float* rec = rtab_data[lo] + (local / dim) * stride + (local % dim);
        const int nslots = stride / dim - 1;
        float4 w = *reinterpret_cast<float4*>(rec);
        float4 s1 = nslots >= 1 ? *reinterpret_cast<float4*>(rec + dim) : make_float4(0, 0, 0, 0);
        float4 s2 = nslots >= 2 ? *reinterpret_cast<float4*>(rec + 2 * dim) : make_float4(0, 0, 0, 0);
        opt_update(o, g.x, w.x, s1.x, s2.x);
        opt_update(o, g.y, w.y, s1.y, s2.y);
        opt_update(o, g.z, w.z, s1.z, s2.z);
        opt_update(o, g.w, w.w, s1.w, s2.w);
        *reinterpret_cast<float4*>(rec) = w;
        if (nslots >= 1) *reinterpret_cast<float4*>(rec + dim) = s1;
        if (nslots >= 2) *reinterpret_cast<float4*>(rec + 2 * dim) = s2;
    }
}
__global__ void __launch_bounds__(256) small_apply_wide_kernel(const float* __restrict__ Gs, const float* __restrict__ touched, int64_t n,
                                                               float4* __restrict__ wide, OptParams o) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (touched[i] == 0.f) continue;
        const float g = Gs[i];
        float4 r = wide[i];
        opt_update(o, g, r.x, r.y, r.z);
        wide[i] = r;
    }
}
int small_apply(WdModel* m) {
    if (m->gs_count == 0) return WD_OK;
    const float* block = m->d_G + m->dense_count;
    if (m->gs_emb_floats > 0) {
        small_apply_emb_kernel<<<grid_for(m->gs_emb_floats / 4, 256), 256, 0, m->stream>>>(block, block + m->gs_touch_off[0], m->gs_emb_floats / 4, m->n_rtab,
            m->n_rtab - m->n_small_tab, m->small_base[0], m->d_rtab_row_base, m->d_rtab_gs_off, m->d_rtab_data, m->d_rtab_dim, m->d_rtab_stride,
            make_opt(m->dnn_opt));
        m->launches++;
    }
    const int64_t nw = m->use_wide ? m->wide_rows - m->small_base[1] : 0;
    if (nw > 0) {
        small_apply_wide_kernel<<<grid_for(nw, 256), 256, 0, m->stream>>>(block + m->gs_emb_floats, block + m->gs_touch_off[1], nw,
            reinterpret_cast<float4*>(m->d_wide) + m->small_base[1], make_opt(m->lin_opt));
        m->launches++;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- sparse Adam (tf.train.AdamOptimizer on IndexedSlices, AdamOptimizer._apply_sparse_shared): m *= beta1, v *= beta2 over the
// WHOLE variable, scatter-add of the summed gradients (opt_update's Adam branch, touched rows), then every row moves by
// lr_t * m / (sqrt(v) + eps) with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t).  Record layout [w | m | v]; bpow = {beta1^t, beta2^t}.
__global__ void __launch_bounds__(256) adam_decay_kernel(float* __restrict__ data, int64_t rows, int dim, int stride, float b1, float b2) {
    const int64_t total = rows * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float* rec = data + (i / dim) * stride + (i % dim);
        rec[dim] *= b1;
        rec[2 * dim] *= b2;
    }
}
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ data, int64_t rows, int dim, int stride, float lr, float eps,
                                                       const float* __restrict__ bpow) {
    const int64_t total = rows * dim;
    const float lr_t = lr * sqrtf(1.f - bpow[1]) / (1.f - bpow[0]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float* rec = data + (i / dim) * stride + (i % dim);
        rec[0] -= lr_t * rec[dim] / (sqrtf(rec[2 * dim]) + eps);
    }
}
static int adam_dense_pass(WdModel* m, int which, bool step) {
    const WdOptimizer& o = which == 0 ? m->dnn_opt : m->lin_opt;
    const float* bpow = m->d_bpow + (which == 0 ? 2 : 0);
    auto run = [&](float* data, int64_t rows, int dim, int stride) {
        if (rows <= 0) return;
        if (!step) adam_decay_kernel<<<grid_for(rows * dim, 256), 256, 0, m->stream>>>(data, rows, dim, stride, o.beta1, o.beta2);
        else adam_step_kernel<<<grid_for(rows * dim, 256), 256, 0, m->stream>>>(data, rows, dim, stride, o.lr, o.epsilon, bpow);
        m->launches++;
    };
    if (which == 0) { for (auto& tb : m->tables) run(tb.data, tb.arows, tb.dim, tb.stride); }
    else run(reinterpret_cast<float*>(m->d_wide), m->wide_rows, 1, 4);          // wide record {w, m, v, -}
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int sparse_apply_which(WdModel* m, int which) {
    int rc;
    // rows already updated by the gradient-sum / combine launches of this step (single-GPU step), unless a caller replaced the list since
    const bool done = m->list_apply_fused[which] && !m->sparse_overridden[which];
    m->list_apply_fused[which] = false;
    if (done) return WD_OK;
    if (which == 0 && m->use_deep && !m->tables.empty()) {
        const bool adam = m->dnn_opt.kind == WD_OPT_ADAM;
        if (adam && (rc = adam_dense_pass(m, 0, false))) return rc;
        emb_apply_kernel<<<grid_for(m->max_nnz * 8, 256), 256, 0, m->stream>>>(
            m->d_nuniq[0], m->d_urow[0], m->d_ugrad[0], m->emb_max_dim, m->n_rtab, m->d_rtab_row_base, m->d_rtab_data,
            m->d_rtab_dim, m->d_rtab_stride, make_opt(m->dnn_opt));          // tables in row order (binary search by row)
        m->launches++;
        if (adam && (rc = adam_dense_pass(m, 0, true))) return rc;
    }
    if (which == 1 && m->use_wide) {
        const bool adam = m->lin_opt.kind == WD_OPT_ADAM;
        if (adam && (rc = adam_dense_pass(m, 1, false))) return rc;
        wide_apply_kernel<<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(m->d_nuniq[1], m->d_urow[1], m->d_ugrad[1], m->d_wide,
                                                                          make_opt(m->lin_opt));
        m->launches++;
        if (adam && (rc = adam_dense_pass(m, 1, true))) return rc;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
// ---- wrappers used by shard.cu (rows this rank owns in a row-sharded table space)
// stable sort of (e_key[i], i) pairs, i < *d_n, by key; keys equal to kInvalidRow sort last; result in d_sk / d_sv of list `which`
int list_sort_by_key(WdModel* m, int which, const int32_t* d_n, const uint32_t* e_key) {
    sort_keys_kernel<<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(d_n, e_key, 1u << m->sort_bits[which], m->d_sk[which], m->d_sv[which], nullptr);
    m->launches++;
    return radix_sort_pairs(m, which, m->sort_bits[which] + 1, d_n);
}
int list_group(WdModel* m, int which, const int32_t* d_n, const uint32_t* e_row) {
    int rc = group_rows(m, which, d_n, e_row);
    if (rc) return rc;
    if ((rc = chunk_offsets(m, m->d_nuniq[which], m->d_ustart[which], m->d_urow[which], m->d_choff[which], m->max_nnz, kChunk, m->d_nchunks[which]))) return rc;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
int list_chunk_combine(WdModel* m, int which, int width) {
    chunk_combine_kernel<0><<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(m->d_nuniq[which], m->d_choff[which], m->d_cpart[which], m->d_ugrad[which], width, HotApply{});
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
int list_apply_emb(WdModel* m, int which, int width, int ntab, const int64_t* d_row_base, float* const* d_data, const int32_t* d_dim,
                   const int32_t* d_stride, const WdOptimizer& o) {
    emb_apply_kernel<<<grid_for(m->max_nnz * 8, 256), 256, 0, m->stream>>>(m->d_nuniq[which], m->d_urow[which], m->d_ugrad[which], width, ntab,
                                                                            d_row_base, d_data, d_dim, d_stride, make_opt(o));
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
int list_apply_wide(WdModel* m, int which, float4* wide, const WdOptimizer& o) {
    wide_apply_kernel<<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(m->d_nuniq[which], m->d_urow[which], m->d_ugrad[which], wide, make_opt(o));
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

int sparse_apply(WdModel* m) {
    int rc = sparse_apply_which(m, 0);
    return rc ? rc : sparse_apply_which(m, 1);
}

}  // namespace wd
