// Row-sharded tables: embedding tables and wide columns too large to replicate are split by row over the G ranks of one box
// (row id -> owner rank id mod G, local row id / G) and reached through PEER MEMORY over NVLink instead of a collective library.
//
// The reference partitions its large variables over parameter-server tasks with tf.min_max_variable_partitioner (reference
// python/lib/joint.py:141-143) and lets every worker pull rows / push sparse updates asynchronously (python/train.py:197-217).
// Here the same partitioning is synchronous and exact — G ranks on G batch shards compute what one rank computes on the
// concatenated batch — and every transfer is fused into the kernel that produces or consumes the data:
//
//   route + send   (requester) ids of sharded columns are grouped by owner (one stable radix pass) and written as
//                  {local row, bag} pairs straight into the owner's inbox (P2P stores)
//   serve          (owner)     gathers the rows of each received bag from its shard, pools them (partial sum) and writes the
//                  pooled vector straight into the requester's receive buffer (P2P stores): gather + all-to-all in one kernel
//   combine        (requester) sums the <= G partials of a bag in rank order, applies the mean, writes the deep-input slice /
//                  adds the wide partial logits
//   backward       (owner)     sorts the received rows, then PULLS each occurrence's gradient (the requester's dX0 slice or
//                  dlogit, P2P loads) while summing per row in a fixed order, and applies Adagrad / FTRL to its shard: the
//                  all-to-all of gradients is fused into the segmented reduction, the optimizer runs once per touched row
//   dense          gradients of the MLP / wide bias / small replicated tables: two-shot all-reduce over peer memory (each rank
//                  reduces one slice in rank order, then every rank gathers the slices) — deterministic, identical on all ranks
//
// Ranks synchronise with flag barriers in peer memory (st.release.sys / ld.acquire.sys); when all ranks live in ONE process
// (tests on a single GPU) the caller orders the phases with events instead (wd_shard_phase + wd_shard_local_sync).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "sparse_dev.cuh"

namespace wd {

constexpr uint32_t kTagBagBits = 27;
constexpr uint32_t kTagBagMask = (1u << kTagBagBits) - 1;
enum { BAR_A = 0, BAR_B = 1, BAR_CW = 2, BAR_CE = 3, BAR_G = 4, BAR_R = 5, BAR_END = 6 };

// ------------------------------------------------------------------------------------------------- flag barrier
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// One warp: lane r signals rank r ("rank `me` reached barrier k for the e-th time") and waits for rank r's signal.  The epoch
// lives in device memory so the kernel can be replayed from a CUDA graph.  A rank that never arrives trips a 20 s timeout that
// raises an error flag instead of hanging the device.
// `trace` (optional): [kBarriers][2] globaltimer stamps {barrier entered, barrier left} of the last step (wd_debug_shard_trace).
__global__ void shard_barrier_kernel(uint32_t* const* __restrict__ peer_flags, uint32_t* __restrict__ my_flags, uint32_t* __restrict__ epoch,
                                     int k, int G, int me, int32_t* __restrict__ err, unsigned long long* __restrict__ trace) {
    __shared__ uint32_t e_sh;
    if (threadIdx.x == 0) { e_sh = epoch[k] + 1u; epoch[k] = e_sh; if (trace) trace[2 * k] = gtimer_ns(); }
    __syncthreads();
    const uint32_t e = e_sh;
    const int r = threadIdx.x;
    if (r < G) {
        __threadfence_system();                                   // everything this device wrote before (peer stores included)
        st_release_sys(peer_flags[r] + k * kMaxRanks + me, e);
        const unsigned long long t0 = gtimer_ns();
        while ((int32_t)(ld_acquire_sys(my_flags + k * kMaxRanks + r) - e) < 0) {
            __nanosleep(100);
            if (gtimer_ns() - t0 > 20000000000ull) { atomicOr(err, 4); break; }
        }
    }
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0 && trace) trace[2 * k + 1] = gtimer_ns();
}

__global__ void shard_stamp_kernel(unsigned long long* __restrict__ t) { *t = gtimer_ns(); }

// --------------------------------------------------------------------------------------------- requester: route + send
// starts[o] = first position of owner o in the owner-sorted key list (keys >= G are "not a sharded column"), o = 0 .. G
__global__ void shard_starts_kernel(const int32_t* __restrict__ d_n, const uint32_t* __restrict__ keys, int G, int32_t* __restrict__ ostart) {
    const int o = threadIdx.x;
    if (o <= G) ostart[o] = lower_bound_u32(keys, *d_n, (uint32_t)o);
}

template <bool EMB>
__device__ __forceinline__ int shard_bag(int bc, int C, int n_slots, const int32_t* __restrict__ col_slot) {
    const int b = bc / C;
    return EMB ? b * n_slots + col_slot[bc - b * C] : b;
}

template <bool EMB>
__global__ void __launch_bounds__(256) shard_send_kernel(const int32_t* __restrict__ ostart, const uint32_t* __restrict__ sk, const uint32_t* __restrict__ sv,
                                                         const uint32_t* __restrict__ lrow, const int32_t* __restrict__ e_bc,
                                                         const int32_t* __restrict__ offs, int C, int n_slots, const int32_t* __restrict__ col_slot,
                                                         const ShardPeer* __restrict__ peers, int G, int me, int64_t pair_cap,
                                                         int32_t* __restrict__ bagmask, float* __restrict__ bagscale, int32_t* __restrict__ err) {
    const int total = ostart[G];
    if (blockIdx.x == 0 && threadIdx.x < G) {
        const int o = threadIdx.x;
        int c = ostart[o + 1] - ostart[o];
        if ((int64_t)c > pair_cap) { c = (int)pair_cap; atomicOr(err, 2); }
        peers[o].inbox_cnt[0][me] = c;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int o = (int)sk[p];
        const int e = (int)sv[p];
        const int i = p - ostart[o];
        if ((int64_t)i >= pair_cap) continue;
        const int bc = e_bc[e];
        const int bag = shard_bag<EMB>(bc, C, n_slots, col_slot);
        peers[o].inbox[0][(int64_t)me * pair_cap + i] = make_uint2(lrow[e], (uint32_t)bag);
        const bool head = i == 0 || shard_bag<EMB>(e_bc[sv[p - 1]], C, n_slots, col_slot) != bag;
        if (head) {
            atomicOr(&bagmask[bag], 1 << o);
            if (EMB) bagscale[bag] = 1.f / (float)(offs[bc + 1] - offs[bc]);      // combiner = mean (same value from every owner's head)
        }
    }
}

// ------------------------------------------------------------------------------------------------------- owner: serve
// flat index f over the received entries of all sources -> (source rank r, position i)
__device__ __forceinline__ bool shard_locate(int64_t f, const int32_t* __restrict__ cnt, int G, int& r, int& i) {
    for (r = 0; r < G; ++r) {
        const int c = __ldcg(cnt + r);
        if (f < c) { i = (int)f; return true; }
        f -= c;
    }
    return false;
}

// embedding space: 8 lanes per received entry; the lanes of a bag's first entry pool the whole run and store the partial sum
__global__ void __launch_bounds__(256) shard_serve_emb_kernel(const uint2* __restrict__ inbox, const int32_t* __restrict__ cnt, int G, int me,
                                                              int64_t pair_cap, int n_slots, const int64_t* __restrict__ slot_base,
                                                              float* const* __restrict__ slot_data, const int32_t* __restrict__ slot_dim,
                                                              const int32_t* __restrict__ slot_stride, const ShardPeer* __restrict__ peers,
                                                              int64_t nbags_cap, int width) {
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    int64_t total = 0;
    for (int r = 0; r < G; ++r) total += __ldcg(cnt + r);
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t f = g0; f < total; f += gstep) {
        int r, i;
        if (!shard_locate(f, cnt, G, r, i)) break;
        const uint2* box = inbox + (int64_t)r * pair_cap;
        const uint2 en = __ldcg(box + i);
        if (i > 0 && __ldcg(box + i - 1).y == en.y) continue;     // not the head of its bag
        int lo = 0, hi = n_slots - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (slot_base[mid] <= (int64_t)en.x) lo = mid; else hi = mid - 1; }
        const int dim = slot_dim[lo], stride = slot_stride[lo];
        const float* data = slot_data[lo];
        const int64_t base = slot_base[lo];
        const int n = __ldcg(cnt + r);
        float* dst = peers[r].recv + ((int64_t)me * nbags_cap + en.y) * width;
        for (int q = lig; q * 4 < dim; q += 8) {
            float4 acc = ldg_nc_f4(data + ((int64_t)en.x - base) * stride + q * 4);
            for (int j = i + 1; j < n; ++j) {
                const uint2 e2 = __ldcg(box + j);
                if (e2.y != en.y) break;
                const float4 v = ldg_nc_f4(data + ((int64_t)e2.x - base) * stride + q * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            *reinterpret_cast<float4*>(dst + q * 4) = acc;
        }
    }
}

// wide space: thread per received entry; partial logit of the example = sum of the weights of its run
__global__ void __launch_bounds__(256) shard_serve_wide_kernel(const uint2* __restrict__ inbox, const int32_t* __restrict__ cnt, int G, int me,
                                                               int64_t pair_cap, const float4* __restrict__ wide, const ShardPeer* __restrict__ peers,
                                                               int64_t nbags_cap) {
    int64_t total = 0;
    for (int r = 0; r < G; ++r) total += __ldcg(cnt + r);
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        int r, i;
        if (!shard_locate(f, cnt, G, r, i)) break;
        const uint2* box = inbox + (int64_t)r * pair_cap;
        const uint2 en = __ldcg(box + i);
        if (i > 0 && __ldcg(box + i - 1).y == en.y) continue;
        const int n = __ldcg(cnt + r);
        float acc = __ldg(&wide[en.x].x);
        for (int j = i + 1; j < n; ++j) {
            const uint2 e2 = __ldcg(box + j);
            if (e2.y != en.y) break;
            acc += __ldg(&wide[e2.x].x);
        }
        peers[r].recv[(int64_t)me * nbags_cap + en.y] = acc;
    }
}

// received entries of all sources, flattened in (source rank, position) order = global (example, column, id) order of the
// concatenated batch, so the per-row gradient sums run in the order a single rank would use
__global__ void __launch_bounds__(256) shard_flatten_kernel(const uint2* __restrict__ inbox, const int32_t* __restrict__ cnt, int G, int64_t pair_cap,
                                                            uint32_t* __restrict__ rrow, uint32_t* __restrict__ rtag, int32_t* __restrict__ d_nrecv,
                                                            int64_t cap, int32_t* __restrict__ err) {
    int64_t total = 0;
    for (int r = 0; r < G; ++r) total += __ldcg(cnt + r);
    if (total > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, 2);
        total = cap;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_nrecv = (int)total;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        int r, i;
        if (!shard_locate(f, cnt, G, r, i)) break;
        const uint2 en = __ldcg(inbox + (int64_t)r * pair_cap + i);
        rrow[f] = en.x;
        rtag[f] = ((uint32_t)r << kTagBagBits) | (en.y & kTagBagMask);
    }
}

// --------------------------------------------------------------------------------------------------- requester: combine
__global__ void __launch_bounds__(256) shard_combine_emb_kernel(int B, int n_slots, const int32_t* __restrict__ slot_dim, const int32_t* __restrict__ slot_x0,
                                                                const int32_t* __restrict__ bagmask, const float* __restrict__ bagscale,
                                                                const float* __restrict__ recv, int G, int64_t nbags_cap, int width,
                                                                float* __restrict__ X0, int ld) {
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    const int64_t nb = (int64_t)B * n_slots;
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t bag = g0; bag < nb; bag += gstep) {
        const int b = (int)(bag / n_slots), slot = (int)(bag - (int64_t)b * n_slots);
        const int mask = bagmask[bag];
        const float scale = mask ? bagscale[bag] : 0.f;
        const int dim = slot_dim[slot];
        for (int q = lig; q * 4 < dim; q += 8) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int o = 0; o < G; ++o)
                if (mask >> o & 1) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(recv + ((int64_t)o * nbags_cap + bag) * width + q * 4));
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
            *reinterpret_cast<float4*>(X0 + (int64_t)b * ld + slot_x0[slot] + q * 4) = acc;
        }
    }
}
__global__ void __launch_bounds__(256) shard_combine_wide_kernel(int B, const int32_t* __restrict__ bagmask, const float* __restrict__ recv, int G,
                                                                 int64_t nbags_cap, float* __restrict__ wide_logit) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const int mask = bagmask[b];
        float acc = 0.f;
        for (int o = 0; o < G; ++o)
            if (mask >> o & 1) acc += __ldcg(recv + (int64_t)o * nbags_cap + b);
        wide_logit[b] += acc;
    }
}

// -------------------------------------------------------------------------------------- owner: gradient sums (P2P pull)
// Per unique owned row: ordered sum over its occurrences of (requester's dX0 slice of the bag) / (ids in the bag), read from the
// requester's memory.  Same decomposition as emb_grad_sum_kernel (8 lanes per item, 4 rows in flight, hot rows in chunks).
template <bool CHUNKED>
__global__ void __launch_bounds__(256) shard_emb_grad_sum_kernel(const int32_t* __restrict__ d_nitems, const int32_t* __restrict__ d_nuniq,
                                                                 const int32_t* __restrict__ ustart, const int32_t* __restrict__ choff,
                                                                 const uint32_t* __restrict__ svals, const uint32_t* __restrict__ rtag,
                                                                 const ShardPeer* __restrict__ peers, int n_slots, const int32_t* __restrict__ slot_dim,
                                                                 const int32_t* __restrict__ slot_x0, int ld, float* __restrict__ out, int width) {
    const int lane = threadIdx.x & 31, lig = lane & 7, grp = lane >> 3;
    const int nitems = *d_nitems, nu = *d_nuniq;
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4 + grp;
    const int64_t gstep = (((int64_t)gridDim.x * blockDim.x) >> 5) * 4;
    for (int64_t it = g0; it < nitems; it += gstep) {
        int s, e;
        if (!CHUNKED) {
            s = ustart[it]; e = ustart[it + 1];
            if (e - s > kChunk) continue;
        } else {
            int u = chunk_owner(choff, nu, (int)it);
            s = ustart[u] + ((int)it - choff[u]) * kChunk;
            e = min(ustart[u + 1], s + kChunk);
        }
        const int slot = (int)((rtag[svals[s]] & kTagBagMask) % (uint32_t)n_slots);
        const int dim = slot_dim[slot], x0 = slot_x0[slot];
        for (int q = lig; q * 4 < width; q += 8) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < dim) {
                int j = s;
                for (; j + 4 <= e; j += 4) {
                    uint32_t tg[4]; float4 v[4]; float inv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) tg[r] = rtag[svals[j + r]];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const ShardPeer& pr = peers[tg[r] >> kTagBagBits];
                        const uint32_t bag = tg[r] & kTagBagMask;
                        v[r] = __ldcg(reinterpret_cast<const float4*>(pr.gradbase + (int64_t)(bag / (uint32_t)n_slots) * ld + x0 + q * 4));
                        inv[r] = __ldcg(pr.bagscale + bag);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) { acc.x += v[r].x * inv[r]; acc.y += v[r].y * inv[r]; acc.z += v[r].z * inv[r]; acc.w += v[r].w * inv[r]; }
                }
                for (; j < e; ++j) {
                    const uint32_t tg = rtag[svals[j]];
                    const ShardPeer& pr = peers[tg >> kTagBagBits];
                    const uint32_t bag = tg & kTagBagMask;
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(pr.gradbase + (int64_t)(bag / (uint32_t)n_slots) * ld + x0 + q * 4));
                    const float inv = __ldcg(pr.bagscale + bag);
                    acc.x += v.x * inv; acc.y += v.y * inv; acc.z += v.z * inv; acc.w += v.w * inv;
                }
            }
            *reinterpret_cast<float4*>(out + (int64_t)it * width + q * 4) = acc;
        }
    }
}

template <bool CHUNKED>
__global__ void shard_wide_grad_sum_kernel(const int32_t* __restrict__ d_nitems, const int32_t* __restrict__ d_nuniq,
                                           const int32_t* __restrict__ ustart, const int32_t* __restrict__ choff,
                                           const uint32_t* __restrict__ svals, const uint32_t* __restrict__ rtag,
                                           const ShardPeer* __restrict__ peers, float* __restrict__ out) {
    const int nitems = *d_nitems, nu = *d_nuniq;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
        int s, e;
        if (!CHUNKED) {
            s = ustart[it]; e = ustart[it + 1];
            if (e - s > kChunk) continue;
        } else {
            int u = chunk_owner(choff, nu, it);
            s = ustart[u] + (it - choff[u]) * kChunk;
            e = min(ustart[u + 1], s + kChunk);
        }
        float acc = 0.f;
        for (int j = s; j < e; ++j) {
            const uint32_t tg = rtag[svals[j]];
            acc += __ldcg(peers[tg >> kTagBagBits].gradbase + (tg & kTagBagMask));       // the example's dlogit on its own rank
        }
        out[it] = acc;
    }
}

// ----------------------------------------------------------------------------------------- dense gradients: all-reduce
// two-shot over peer memory: rank `me` sums slice `me` of every rank's arena in rank order, then every rank copies all slices
__global__ void __launch_bounds__(256) shard_ar_reduce_kernel(float* const* __restrict__ peer_G, float* __restrict__ gred, int64_t n4, int64_t slice4,
                                                              int G, int me) {
    const int64_t lo = (int64_t)me * slice4, hi = min(n4, lo + slice4);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        // every rank's value in flight before the first add (a peer load is ~2 us: one round trip instead of G), summed in rank order
        float4 v[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            v[r] = r < G ? __ldcg(reinterpret_cast<const float4*>(peer_G[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = v[0];
#pragma unroll
        for (int r = 1; r < kMaxRanks; ++r)
            if (r < G) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        reinterpret_cast<float4*>(gred)[i] = acc;
    }
}
__global__ void __launch_bounds__(256) shard_ar_gather_kernel(float* const* __restrict__ peer_gred, float* __restrict__ Gout, int64_t n4, int64_t slice4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int owner = (int)(i / slice4);
        reinterpret_cast<float4*>(Gout)[i] = __ldcg(reinterpret_cast<const float4*>(peer_gred[owner]) + i);
    }
}

// ================================================================================================== host side
static int bits_for64(int64_t n) { int b = 1; while ((1ll << b) < n) ++b; return b; }
static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

template <typename T>
static int upload_arr(WdModel* m, const std::vector<T>& h, T** out) {
    T* p = nullptr;
    int rc = dev_alloc(m, &p, (int64_t)std::max<size_t>(h.size(), 1), true);
    if (rc) return rc;
    if (!h.empty()) WD_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, m->stream));
    *out = p;
    return WD_OK;
}

// Lays out the sharded spaces and the exchange segment.  Called at the end of build_model (world > 1): by then the dense arena
// size is known; d_dX0 / d_dlogit / d_G live INSIDE the segment so that peers can read them.
int shard_build(WdModel* m, const WdPlanDesc* d) {
    ShardState& S = m->shard;
    S.world = d->shard_world; S.rank = d->shard_rank;
    const int G = S.world, C = m->n_columns;
    if (G > kMaxRanks) { set_error("shard_world %d > %d", G, kMaxRanks); return WD_EUNSUPPORTED; }
    int rc;
    const int64_t route_cap = m->max_nnz;                                  // ids one rank can route per step and space
    // ---- embedding space
    {
        ShardSpace& sp = S.sp[0];
        std::vector<int32_t> col_slot(C, -1), dim, x0, stride;
        std::vector<int64_t> base;
        std::vector<float*> data;
        int64_t rows = 0;
        for (size_t t = 0; t < m->tables.size(); ++t) {
            EmbTable& tb = m->tables[t];
            if (!tb.sharded) continue;
            col_slot[tb.col] = sp.n_slots++;
            base.push_back(rows); dim.push_back(tb.dim); x0.push_back(tb.x0_off); stride.push_back(tb.stride); data.push_back(tb.data);
            tb.row_base = rows;
            rows += (tb.rows + G - 1) / G;            // the SAME layout on every rank (a requester computes the owner's local row): ceil(rows / G) per table
            sp.width = std::max(sp.width, tb.dim);
        }
        sp.on = sp.n_slots > 0;
        sp.local_rows = rows;
        sp.bags_per_row = sp.n_slots;
        sp.h_col_slot = col_slot; sp.h_slot_base = base;
        if (sp.on) {
            if ((rc = upload_arr(m, col_slot, &sp.d_col_slot))) return rc;
            if ((rc = upload_arr(m, base, &sp.d_slot_base))) return rc;
            if ((rc = upload_arr(m, dim, &sp.d_slot_dim))) return rc;
            if ((rc = upload_arr(m, x0, &sp.d_slot_x0))) return rc;
            if ((rc = upload_arr(m, stride, &sp.d_slot_stride))) return rc;
            if ((rc = upload_arr(m, data, &sp.d_slot_data))) return rc;
        }
    }
    // ---- wide space
    {
        ShardSpace& sp = S.sp[1];
        std::vector<int32_t> col_slot(C, -1);
        std::vector<int64_t> base;
        int64_t rows = 0;
        if (m->use_wide && d->col_wide_sharded)
            for (int c = 0; c < C; ++c)
                if (d->col_wide_sharded[c]) {
                    col_slot[c] = sp.n_slots++;
                    base.push_back(rows);
                    rows += (d->col_buckets[c] + G - 1) / G;     // rank-independent layout: ceil(buckets / G) rows per column
                }
        sp.h_col_slot = col_slot; sp.h_slot_base = base;
        sp.on = sp.n_slots > 0;
        sp.local_rows = rows;
        sp.width = 1;
        sp.bags_per_row = 1;
        if (sp.on) {
            if ((rc = upload_arr(m, col_slot, &sp.d_col_slot))) return rc;
            if ((rc = upload_arr(m, base, &sp.d_slot_base))) return rc;
            if ((rc = dev_alloc(m, &sp.d_wide, rows))) return rc;
        }
    }
    for (int s = 0; s < 2; ++s)
        if (S.sp[s].local_rows >= (1ll << 30)) { set_error("more than 2^30 sharded rows per rank in one table space"); return WD_EUNSUPPORTED; }
    // ---- per-space scratch (requester + owner) and the sort lists 2 + s (owned rows) / 4 + s (routing)
    for (int s = 0; s < 2; ++s) {
        ShardSpace& sp = S.sp[s];
        if (!sp.on) continue;
        sp.nbags_cap = (int64_t)m->max_batch * sp.bags_per_row;
        if (sp.nbags_cap > (int64_t)kTagBagMask) { set_error("too many sharded bags per step for the tag encoding"); return WD_EUNSUPPORTED; }
        sp.pair_cap = route_cap;
        if ((rc = dev_alloc(m, &sp.d_own, m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &sp.d_lrow, m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &sp.d_ostart, kMaxRanks + 2))) return rc;
        if ((rc = dev_alloc(m, &sp.d_bagmask, sp.nbags_cap + 8))) return rc;
        if ((rc = dev_alloc(m, &sp.d_rtag, m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &sp.d_rrow, m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &sp.d_nrecv, 4))) return rc;
        if ((rc = dev_alloc(m, &sp.d_peers, kMaxRanks))) return rc;
        const int Lo = 2 + s, Lr = 4 + s, w = s == 0 ? std::max(sp.width, 4) : 1;
        for (int L : {Lo, Lr}) {
            if ((rc = dev_alloc(m, &m->d_sk[L], m->max_nnz + 8))) return rc;
            if ((rc = dev_alloc(m, &m->d_sv[L], m->max_nnz + 8))) return rc;
            if ((rc = dev_alloc(m, &m->d_sk2[L], m->max_nnz + 8))) return rc;
            if ((rc = dev_alloc(m, &m->d_sv2[L], m->max_nnz + 8))) return rc;
        }
        if ((rc = dev_alloc(m, &m->d_urow[Lo], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_ustart[Lo], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_ugrad[Lo], (m->max_nnz + 8) * w))) return rc;
        if ((rc = dev_alloc(m, &m->d_nuniq[Lo], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_nvalid[Lo], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_choff[Lo], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_nchunks[Lo], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_cpart[Lo], m->cpart_cap * w))) return rc;
        m->sort_bits[Lo] = bits_for64(std::max<int64_t>(sp.local_rows, 2));
        m->sort_bits[Lr] = bits_for64(std::max(G, 2));
    }
    // ---- exchange segment layout (identical on every rank)
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 256); return o; };
    S.off_flags = take((int64_t)kBarriers * kMaxRanks * 4);
    for (int s = 0; s < 2; ++s) {
        ShardSpace& sp = S.sp[s];
        if (!sp.on) continue;
        sp.off_inbox[0] = sp.off_inbox[1] = take((int64_t)G * sp.pair_cap * (int64_t)sizeof(uint2));   // (one buffer: see the hazard note below)
        sp.off_cnt[0] = sp.off_cnt[1] = take(kMaxRanks * 4);
        sp.off_recv = take((int64_t)G * sp.nbags_cap * sp.width * 4);
        sp.off_bagscale = take(sp.nbags_cap * 4);
    }
    const int64_t x0n = m->use_deep ? (int64_t)m->max_batch_pad * std::max(m->d0_phys, 1) : 4;
    S.sp[0].off_grad = take(x0n * 4);                                      // dX0
    S.sp[1].off_grad = take((int64_t)m->max_batch * 4 + 64);               // dlogit
    S.ar_count = align_up(m->dense_count + m->gs_count, 4);
    S.off_G = take(std::max<int64_t>(S.ar_count, 4) * 4);
    S.off_gred = take(std::max<int64_t>(S.ar_count, 4) * 4);
    S.seg_bytes = off;
    void* seg = nullptr;
    cudaError_t e = cudaMalloc(&seg, (size_t)S.seg_bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc of the %lld-byte exchange segment failed: %s", (long long)S.seg_bytes, cudaGetErrorString(e)); return WD_ENOMEM; }
    m->allocs.push_back(seg);
    m->bytes_allocated += S.seg_bytes;
    WD_CUDA(cudaMemsetAsync(seg, 0, (size_t)S.seg_bytes, m->stream));
    S.seg = (uint8_t*)seg;
    m->d_dX0 = reinterpret_cast<float*>(S.seg + S.sp[0].off_grad);
    m->d_dlogit = reinterpret_cast<float*>(S.seg + S.sp[1].off_grad);
    m->d_G = reinterpret_cast<float*>(S.seg + S.off_G);
    S.gred = reinterpret_cast<float*>(S.seg + S.off_gred);
    if ((rc = dev_alloc(m, &S.d_peer_flags, kMaxRanks))) return rc;
    if ((rc = dev_alloc(m, &S.d_epoch, kBarriers))) return rc;
    if (getenv("WD_SHARD_TRACE") && (rc = dev_alloc(m, &S.d_trace, 2 * kBarriers))) return rc;
    if ((rc = dev_alloc(m, &S.d_peer_G, kMaxRanks))) return rc;
    if ((rc = dev_alloc(m, &S.d_peer_gred, kMaxRanks))) return rc;
    WD_CUDA(cudaEventCreateWithFlags(&S.ev_a, cudaEventDisableTiming));
    WD_CUDA(cudaEventCreateWithFlags(&S.ev_ids2, cudaEventDisableTiming));
    WD_CUDA(cudaEventCreateWithFlags(&S.ev_routed1, cudaEventDisableTiming));
    WD_CUDA(cudaEventCreateWithFlags(&S.ev_a2, cudaEventDisableTiming));
    WD_CUDA(cudaEventCreateWithFlags(&S.ev_aux_done, cudaEventDisableTiming));
    WD_CUDA(cudaStreamCreateWithFlags(&S.aux, cudaStreamNonBlocking));
    return WD_OK;
}

// peer segment bases known: fill the per-peer pointer tables
static int shard_finish_connect(WdModel* m) {
    ShardState& S = m->shard;
    const int G = S.world;
    std::vector<uint32_t*> pf(kMaxRanks, nullptr);
    std::vector<float*> pg(kMaxRanks, nullptr), pr(kMaxRanks, nullptr);
    for (int r = 0; r < G; ++r) {
        uint8_t* b = S.peer_seg[r];
        pf[r] = reinterpret_cast<uint32_t*>(b + S.off_flags);
        pg[r] = reinterpret_cast<float*>(b + S.off_G);
        pr[r] = reinterpret_cast<float*>(b + S.off_gred);
        for (int s = 0; s < 2; ++s) {
            ShardSpace& sp = S.sp[s];
            if (!sp.on) continue;
            ShardPeer& p = sp.peers[r];
            for (int k = 0; k < 2; ++k) {
                p.inbox[k] = reinterpret_cast<uint2*>(b + sp.off_inbox[k]);
                p.inbox_cnt[k] = reinterpret_cast<int32_t*>(b + sp.off_cnt[k]);
            }
            p.recv = reinterpret_cast<float*>(b + sp.off_recv);
            p.bagscale = reinterpret_cast<const float*>(b + sp.off_bagscale);
            p.gradbase = reinterpret_cast<const float*>(b + sp.off_grad);
        }
    }
    WD_CUDA(cudaMemcpyAsync(S.d_peer_flags, pf.data(), kMaxRanks * sizeof(void*), cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaMemcpyAsync(S.d_peer_G, pg.data(), kMaxRanks * sizeof(void*), cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaMemcpyAsync(S.d_peer_gred, pr.data(), kMaxRanks * sizeof(void*), cudaMemcpyHostToDevice, m->stream));
    for (int s = 0; s < 2; ++s)
        if (S.sp[s].on) WD_CUDA(cudaMemcpyAsync(S.sp[s].d_peers, S.sp[s].peers, sizeof(ShardPeer) * kMaxRanks, cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaStreamSynchronize(m->stream));
    S.connected = true;
    return WD_OK;
}

static int barrier(WdModel* m, int k) {
    ShardState& S = m->shard;
    if (!S.ipc) return WD_OK;                                  // ranks of one process: the caller orders the phases with events
    shard_barrier_kernel<<<1, 32, 0, m->stream>>>(S.d_peer_flags, reinterpret_cast<uint32_t*>(S.seg + S.off_flags), S.d_epoch, k, S.world, S.rank,
                                                  m->d_flags, S.d_trace);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- requester: group the step's sharded ids by owner and store them into the owners' inboxes
static int shard_route_send(WdModel* m, int s) {
    ShardState& S = m->shard;
    ShardSpace& sp = S.sp[s];
    if (!sp.on) return WD_OK;
    const int G = S.world, L = 4 + s;
    int rc;
    WD_CUDA(cudaMemsetAsync(sp.d_bagmask, 0, (size_t)(m->dbatch.B * (int64_t)sp.bags_per_row) * 4, m->stream));
    // keys = owner (or "not sharded" = 1 << bits, sorts last), values = entry index; one stable radix pass
    if ((rc = list_sort_by_key(m, L, m->d_nnz, sp.d_own))) return rc;
    shard_starts_kernel<<<1, 32, 0, m->stream>>>(m->d_nnz, m->d_sk[L], G, sp.d_ostart);
    float* bagscale = const_cast<float*>(sp.peers[S.rank].bagscale);
    const int g = grid_for(m->max_nnz, 256);
    if (s == 0)
        shard_send_kernel<true><<<g, 256, 0, m->stream>>>(sp.d_ostart, m->d_sk[L], m->d_sv[L], sp.d_lrow, m->d_e_bc, m->d_col_offs, m->n_columns,
            sp.n_slots, sp.d_col_slot, sp.d_peers, G, S.rank, sp.pair_cap, sp.d_bagmask, bagscale, m->d_flags);
    else
        shard_send_kernel<false><<<g, 256, 0, m->stream>>>(sp.d_ostart, m->d_sk[L], m->d_sv[L], sp.d_lrow, m->d_e_bc, m->d_col_offs, m->n_columns,
            sp.n_slots, sp.d_col_slot, sp.d_peers, G, S.rank, sp.pair_cap, sp.d_bagmask, bagscale, m->d_flags);
    m->launches += 2;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- owner: pooled partial sums of the received bags -> requesters' receive buffers
static int shard_serve(WdModel* m, int s) {
    ShardState& S = m->shard;
    ShardSpace& sp = S.sp[s];
    if (!sp.on) return WD_OK;
    const ShardPeer& me = sp.peers[S.rank];
    if (s == 0)
        shard_serve_emb_kernel<<<grid_for(m->max_nnz * 8, 256, 148 * 8), 256, 0, m->stream>>>(me.inbox[0], me.inbox_cnt[0], S.world, S.rank, sp.pair_cap,
            sp.n_slots, sp.d_slot_base, sp.d_slot_data, sp.d_slot_dim, sp.d_slot_stride, sp.d_peers, sp.nbags_cap, sp.width);
    else
        shard_serve_wide_kernel<<<grid_for(m->max_nnz, 256, 148 * 8), 256, 0, m->stream>>>(me.inbox[0], me.inbox_cnt[0], S.world, S.rank, sp.pair_cap,
            sp.d_wide, sp.d_peers, sp.nbags_cap);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- owner: sort the received rows (depends only on ids: runs beside the towers)
static int shard_owner_group(WdModel* m, int s) {
    ShardState& S = m->shard;
    ShardSpace& sp = S.sp[s];
    if (!sp.on) return WD_OK;
    const ShardPeer& me = sp.peers[S.rank];
    shard_flatten_kernel<<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(me.inbox[0], me.inbox_cnt[0], S.world, sp.pair_cap, sp.d_rrow, sp.d_rtag,
                                                                            sp.d_nrecv, m->max_nnz, m->d_flags);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return list_group(m, 2 + s, sp.d_nrecv, sp.d_rrow);
}

static int shard_combine(WdModel* m, int s) {
    ShardState& S = m->shard;
    ShardSpace& sp = S.sp[s];
    if (!sp.on) return WD_OK;
    const int B = m->dbatch.B;
    const ShardPeer& me = sp.peers[S.rank];
    if (s == 0)
        shard_combine_emb_kernel<<<grid_for((int64_t)B * sp.n_slots * 8, 256, 148 * 8), 256, 0, m->stream>>>(B, sp.n_slots, sp.d_slot_dim, sp.d_slot_x0,
            sp.d_bagmask, me.bagscale, me.recv, S.world, sp.nbags_cap, sp.width, m->d_X0, m->d0_phys);
    else
        shard_combine_wide_kernel<<<grid_for(B, 256), 256, 0, m->stream>>>(B, sp.d_bagmask, me.recv, S.world, sp.nbags_cap, m->d_wide_logit);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- owner: per-row gradient sums (pulled from the requesters) + optimizer on the shard
static int shard_owner_reduce_apply(WdModel* m, int s) {
    ShardState& S = m->shard;
    ShardSpace& sp = S.sp[s];
    if (!sp.on) return WD_OK;
    const int L = 2 + s;
    int rc;
    if (s == 0) {
        shard_emb_grad_sum_kernel<false><<<grid_for(m->max_nnz * 8, 256), 256, 0, m->stream>>>(m->d_nuniq[L], m->d_nuniq[L], m->d_ustart[L], m->d_choff[L],
            m->d_sv[L], sp.d_rtag, sp.d_peers, sp.n_slots, sp.d_slot_dim, sp.d_slot_x0, m->d0_phys, m->d_ugrad[L], sp.width);
        shard_emb_grad_sum_kernel<true><<<grid_for(m->cpart_cap * 8, 256), 256, 0, m->stream>>>(m->d_nchunks[L], m->d_nuniq[L], m->d_ustart[L], m->d_choff[L],
            m->d_sv[L], sp.d_rtag, sp.d_peers, sp.n_slots, sp.d_slot_dim, sp.d_slot_x0, m->d0_phys, m->d_cpart[L], sp.width);
        m->launches += 2;
        if ((rc = list_chunk_combine(m, L, sp.width))) return rc;
        if ((rc = list_apply_emb(m, L, sp.width, sp.n_slots, sp.d_slot_base, sp.d_slot_data, sp.d_slot_dim, sp.d_slot_stride, m->dnn_opt))) return rc;
    } else {
        shard_wide_grad_sum_kernel<false><<<grid_for(m->max_nnz, 256), 256, 0, m->stream>>>(m->d_nuniq[L], m->d_nuniq[L], m->d_ustart[L], m->d_choff[L],
            m->d_sv[L], sp.d_rtag, sp.d_peers, m->d_ugrad[L]);
        shard_wide_grad_sum_kernel<true><<<grid_for(m->cpart_cap, 256), 256, 0, m->stream>>>(m->d_nchunks[L], m->d_nuniq[L], m->d_ustart[L], m->d_choff[L],
            m->d_sv[L], sp.d_rtag, sp.d_peers, m->d_cpart[L]);
        m->launches += 2;
        if ((rc = list_chunk_combine(m, L, 1))) return rc;
        if ((rc = list_apply_wide(m, L, sp.d_wide, m->lin_opt))) return rc;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

static int shard_ar_reduce(WdModel* m) {
    ShardState& S = m->shard;
    if (S.ar_count == 0) return WD_OK;
    const int64_t n4 = S.ar_count / 4, slice4 = (n4 + S.world - 1) / S.world;
    shard_ar_reduce_kernel<<<grid_for(slice4, 256), 256, 0, m->stream>>>(S.d_peer_G, S.gred, n4, slice4, S.world, S.rank);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}
static int shard_ar_gather(WdModel* m) {
    ShardState& S = m->shard;
    if (S.ar_count == 0) return WD_OK;
    const int64_t n4 = S.ar_count / 4, slice4 = (n4 + S.world - 1) / S.world;
    shard_ar_gather_kernel<<<grid_for(n4, 256), 256, 0, m->stream>>>(S.d_peer_gred, m->d_G, n4, slice4);
    m->launches++;
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---------------------------------------------------------------------------------------------------------- the step
int sparse_forward(WdModel* m);
int mlp_forward(WdModel* m, bool train);
int loss_forward(WdModel* m, bool need_grad);
int ids_prepare(WdModel* m);
int shard_backward_local(WdModel* m, bool overlap);     // api.cu: towers' backward + replicated lists + dense gradient arena
int shard_apply_local(WdModel* m);                     // api.cu: dense optimizer + small-table block + joins
int shard_group_async(WdModel* m);                     // api.cu: replicated lists' grouping on the side streams

// run `fn` on the side stream of sparse list `w` (its scratch set), as api.cu does for the replicated lists
template <typename F>
static int on_side(WdModel* m, int w, F fn) {
    cudaStream_t main_stream = m->stream;
    m->stream = m->sstream[w]; m->scratch_sel = 1 + w;
    int rc = fn();
    m->stream = main_stream; m->scratch_sel = 0;
    return rc;
}

// run `fn` on the auxiliary stream (scratch set 3)
template <typename F>
static int on_aux(WdModel* m, F fn) {
    cudaStream_t main_stream = m->stream;
    m->stream = m->shard.aux; m->scratch_sel = 3;
    int rc = fn();
    m->stream = main_stream; m->scratch_sel = 0;
    return rc;
}

// The critical chain in front of the towers is   ids -> route + send (embedding space) -> [A] -> serve (embedding space) -> [B].
// Everything else that must exist before the towers — the wide space's routing, its serve, and this rank's local gathers
// (replicated tables, wide bias) — is independent of that chain, so it runs beside it on an auxiliary stream: routing right after
// the ids, serving + local gathers once barrier A has passed, joined before barrier B.  (The two side streams cannot take it: they
// already hold the grouping of the replicated lists, ~100 us of sort launches enqueued before barrier A.)
static bool aux_split(WdModel* m) { return m->shard.sp[0].on && m->shard.sp[1].on; }

// phase 0: ids, routing
int shard_phase0(WdModel* m, bool train) {
    ShardState& S = m->shard;
    int rc;
    if ((rc = ids_prepare(m))) return rc;
    const bool split = aux_split(m);
    if (split) {
        WD_CUDA(cudaEventRecord(S.ev_ids2, m->stream));
        WD_CUDA(cudaStreamWaitEvent(S.aux, S.ev_ids2, 0));
        if ((rc = on_aux(m, [&] { return shard_route_send(m, 1); }))) return rc;
        WD_CUDA(cudaEventRecord(S.ev_routed1, S.aux));
    }
    if (train && (rc = shard_group_async(m))) return rc;
    if ((rc = shard_route_send(m, 0))) return rc;
    if (!split && (rc = shard_route_send(m, 1))) return rc;
    if (split) WD_CUDA(cudaStreamWaitEvent(m->stream, S.ev_routed1, 0));
    return WD_OK;
}
// after barrier A: serve both spaces and run the local part of the forward (needed only by this rank's towers, while every peer's
// barrier B waits for the serves)
static int shard_serve_both(WdModel* m) {
    ShardState& S = m->shard;
    int rc;
    if (aux_split(m)) {
        WD_CUDA(cudaEventRecord(S.ev_a2, m->stream));
        WD_CUDA(cudaStreamWaitEvent(S.aux, S.ev_a2, 0));
        if ((rc = on_aux(m, [&] { int r = shard_serve(m, 1); return r ? r : sparse_forward(m); }))) return rc;
        WD_CUDA(cudaEventRecord(S.ev_aux_done, S.aux));
        if ((rc = shard_serve(m, 0))) return rc;
        WD_CUDA(cudaStreamWaitEvent(m->stream, S.ev_aux_done, 0));
        return WD_OK;
    }
    for (int s = 0; s < 2; ++s) if ((rc = shard_serve(m, s))) return rc;
    return sparse_forward(m);
}
// phase 1: serve the peers, local gathers; sort what was received
int shard_phase1(WdModel* m, bool train) {
    int rc;
    if ((rc = shard_serve_both(m))) return rc;
    if (train)
        for (int s = 0; s < 2; ++s) if ((rc = shard_owner_group(m, s))) return rc;
    return WD_OK;
}
// phase 2: combine, towers forward / backward, dense gradient arena
int shard_phase2(WdModel* m, bool train) {
    int rc;
    for (int s = 0; s < 2; ++s) if ((rc = shard_combine(m, s))) return rc;
    if ((rc = mlp_forward(m, train))) return rc;
    if ((rc = loss_forward(m, train))) return rc;
    if (!train) return WD_OK;
    if (m->side_pending[0] || m->side_pending[1]) WD_CUDA(cudaEventRecord(m->ev_head, m->stream));   // dlogit exists (as forward_core does)
    return shard_backward_local(m, false);
}
// phase 3: owners pull gradients and update their shards; first half of the all-reduce
int shard_phase3(WdModel* m) {
    int rc;
    for (int s = 1; s >= 0; --s) if ((rc = shard_owner_reduce_apply(m, s))) return rc;
    return shard_ar_reduce(m);
}
// phase 4: second half of the all-reduce, dense optimizers
int shard_phase4(WdModel* m) {
    int rc;
    if ((rc = shard_ar_gather(m))) return rc;
    if ((rc = shard_apply_local(m))) return rc;
    m->shard.step++;
    return WD_OK;
}

// The whole step of one rank of a multi-process job.  Main stream: ids, routing, serve, combine, towers, dense all-reduce and
// optimizers, with flag barriers A (ids delivered), B (pooled sums delivered), G (gradient arenas final), R (slices reduced) and
// END.  Side stream of each table space: the owner-side grouping of the received rows (needs only ids: runs beside the towers)
// and, once every rank's dlogit / dX0 exists (barriers Cw / Ce, on the side streams), the owners' pulled gradient sums and
// optimizer — hidden behind the remaining weight gradients, the dense all-reduce and the dense optimizer.
// Buffer hazards: within a step every producer / consumer pair is separated by one of the barriers; the END barrier keeps a fast
// rank from starting the next step's sends while a slow owner still pulls this step's gradients and bag scales.
int shard_step_ipc(WdModel* m, bool train) {
    ShardState& S = m->shard;
    int rc;
    if (S.d_trace) { shard_stamp_kernel<<<1, 1, 0, m->stream>>>(S.d_trace + 2 * (kBarriers - 1)); m->launches++; }   // step start
    if ((rc = shard_phase0(m, train))) return rc;
    if ((rc = barrier(m, BAR_A))) return rc;
    if ((rc = shard_serve_both(m))) return rc;
    if (train) {
        WD_CUDA(cudaEventRecord(S.ev_a, m->stream));
        for (int s = 0; s < 2; ++s) {
            if (!S.sp[s].on) continue;
            if (m->side_pending[s]) {
                WD_CUDA(cudaStreamWaitEvent(m->sstream[s], S.ev_a, 0));
                if ((rc = on_side(m, s, [&] { return shard_owner_group(m, s); }))) return rc;
            } else if ((rc = shard_owner_group(m, s))) return rc;
        }
    }
    if ((rc = barrier(m, BAR_B))) return rc;
    if ((rc = shard_phase2(m, train))) return rc;               // combine, towers forward (+ backward, replicated lists, dense gradient arena)
    if (!train) { S.step++; return barrier(m, BAR_END); }
    for (int s = 1; s >= 0; --s) {
        if (!S.sp[s].on) continue;
        auto owner = [&]() -> int { int r = barrier(m, s == 1 ? BAR_CW : BAR_CE); return r ? r : shard_owner_reduce_apply(m, s); };
        if (m->side_active[s]) { if ((rc = on_side(m, s, owner))) return rc; }
        else if ((rc = owner())) return rc;
    }
    if ((rc = barrier(m, BAR_G))) return rc;                    // every rank's gradient arena (dense + small-table block) is final
    if ((rc = shard_ar_reduce(m))) return rc;
    if ((rc = barrier(m, BAR_R))) return rc;
    if ((rc = shard_phase4(m))) return rc;                      // gather the reduced slices, dense optimizers, join the side streams
    if (S.d_trace) { shard_stamp_kernel<<<1, 1, 0, m->stream>>>(S.d_trace + 2 * (kBarriers - 1) + 1); m->launches++; }   // before END
    return barrier(m, BAR_END);
}

}  // namespace wd

using namespace wd;

// ================================================================================================== C-ABI
extern "C" int wd_shard_info(WdModel* m, int32_t* world, int32_t* rank, int64_t* seg_bytes) {
    if (!m) { set_error("null model"); return WD_EINVAL; }
    if (world) *world = m->shard.world;
    if (rank) *rank = m->shard.rank;
    if (seg_bytes) *seg_bytes = m->shard.seg_bytes;
    return WD_OK;
}

// debugging aid (not part of the public header): enter / leave stamps (ns, globaltimer) of the last step's flag barriers, WD_SHARD_TRACE=1
extern "C" int wd_debug_shard_trace(WdModel* m, unsigned long long* out) {
    if (!m || !m->shard.d_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, m->shard.d_trace, sizeof(unsigned long long) * 2 * kBarriers, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

extern "C" int wd_shard_ipc_handle(WdModel* m, void* handle_out64) {
    if (!m || !handle_out64) { set_error("null argument"); return WD_EINVAL; }
    if (m->shard.world <= 1 || !m->shard.seg) { set_error("model has no sharded tables (shard_world <= 1)"); return WD_ESTATE; }
    WD_CUDA(cudaSetDevice(m->device));
    cudaIpcMemHandle_t h;
    WD_CUDA(cudaIpcGetMemHandle(&h, m->shard.seg));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out64, &h, 64);
    return WD_OK;
}

extern "C" int wd_shard_connect_ipc(WdModel* m, const void* handles, int32_t n_ranks) {
    if (!m || !handles) { set_error("null argument"); return WD_EINVAL; }
    ShardState& S = m->shard;
    if (n_ranks != S.world) { set_error("wd_shard_connect_ipc: %d handles for shard_world %d", n_ranks, S.world); return WD_EINVAL; }
    WD_CUDA(cudaSetDevice(m->device));
    for (int r = 0; r < S.world; ++r) {
        if (r == S.rank) { S.peer_seg[r] = S.seg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e)); return WD_ECUDA; }
        S.peer_seg[r] = (uint8_t*)p;
    }
    S.ipc = true;
    return shard_finish_connect(m);
}

extern "C" int wd_shard_connect_local(WdModel** models, int32_t n_ranks) {
    if (!models || n_ranks < 1) { set_error("bad arguments"); return WD_EINVAL; }
    for (int r = 0; r < n_ranks; ++r) {
        WdModel* m = models[r];
        if (!m || m->shard.world != n_ranks || m->shard.rank != r) { set_error("wd_shard_connect_local: handle %d is not rank %d of %d", r, r, n_ranks); return WD_EINVAL; }
        if (m->shard.seg_bytes != models[0]->shard.seg_bytes) { set_error("exchange segments differ between ranks (plans differ)"); return WD_EINVAL; }
    }
    for (int r = 0; r < n_ranks; ++r) {
        WdModel* m = models[r];
        WD_CUDA(cudaSetDevice(m->device));
        for (int q = 0; q < n_ranks; ++q) {
            m->shard.peer_seg[q] = models[q]->shard.seg;
            if (models[q]->device != m->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(models[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return WD_ECUDA; }
                cudaGetLastError();
            }
        }
        m->shard.ipc = false;
        int rc = shard_finish_connect(m);
        if (rc) return rc;
    }
    return WD_OK;
}

// every stream of every handle waits for everything enqueued so far on all of them (device side only): the "barrier" between
// phases when all ranks are driven by one process
extern "C" int wd_shard_local_sync(WdModel** models, int32_t n_ranks) {
    if (!models) { set_error("null argument"); return WD_EINVAL; }
    std::vector<cudaEvent_t> evs;
    for (int r = 0; r < n_ranks; ++r) {
        WdModel* m = models[r];
        WD_CUDA(cudaSetDevice(m->device));
        for (cudaStream_t st : {m->stream, m->sstream[0], m->sstream[1], m->shard.aux}) {
            if (!st) continue;
            cudaEvent_t ev;
            WD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            WD_CUDA(cudaEventRecord(ev, st));
            evs.push_back(ev);
        }
    }
    for (int r = 0; r < n_ranks; ++r) {
        WdModel* m = models[r];
        WD_CUDA(cudaSetDevice(m->device));
        for (cudaStream_t st : {m->stream, m->sstream[0], m->sstream[1], m->shard.aux})
            if (st) for (cudaEvent_t ev : evs) WD_CUDA(cudaStreamWaitEvent(st, ev, 0));
    }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
    return WD_OK;
}
