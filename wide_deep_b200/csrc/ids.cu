// Stage K1/K2: batch keys -> categorical-column ids (hash-bucket, vocabulary, identity, bucketize, hashed cross).
//
// Replaces the per-step feature-column transforms of the reference:
//   categorical_column_with_hash_bucket   Fingerprint64(s) mod n            (build_estimator.py:86-88)
//   categorical_column_with_vocabulary_list  index in list, OOV dropped      (build_estimator.py:102-106; A.3)
//   categorical_column_with_identity      out of range -> 0, -1 dropped      (build_estimator.py:114-116; A.4)
//   bucketized_column                     #boundaries <= x, fp32             (build_estimator.py:133,145; A.6)
//   crossed_column                        FingerprintCat64 chain, seed 0xDECAFCAFFE, Cartesian product with
//                                         the last key innermost             (build_estimator.py:153; A.5)
// All integer results are bit-exact with TensorFlow's CPU kernels.  Output: CSR over (row, column) plus, per
// entry, the global wide-table row and the global embedding row it addresses.
#include "common.cuh"
#include "farmhash.cuh"

namespace wd {

__device__ __forceinline__ void field_range(const DevBatch& bt, int F, int b, int f, int& s, int& e) {
    if (bt.cat_offsets) {
        s = bt.cat_offsets[b * F + f];
        e = bt.cat_offsets[b * F + f + 1];
    } else {
        s = b * F + f;
        e = s + 1;
    }
}

__device__ __forceinline__ float normalise(int kind, float a, float bb, float x) {
    // fp32 IEEE arithmetic exactly as the reference's normalizer lambdas (build_estimator.py:61-68)
    if (kind == WD_NORM_MINMAX || kind == WD_NORM_STANDARD) return __fdiv_rn(__fsub_rn(x, a), bb);
    if (kind == WD_NORM_LOG) return logf(x);
    return x;
}

__device__ __forceinline__ int bucketize(const float* bounds, int nb, float x) {
    int lo = 0, hi = nb;                       // upper_bound: number of boundaries <= x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (bounds[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int vocab_lookup(const uint64_t* v, int n, uint64_t key) {
    for (int i = 0; i < n; ++i)
        if (v[i] == key) return i;
    return -1;
}

// id of one key under a simple (non-cross) column; returns false if the key is dropped
__device__ __forceinline__ bool simple_id(const DevPlan& p, int c, int kind, uint64_t key, int64_t* id) {
    if (kind == WD_COL_HASH) {
        if (key == kFpEmpty) return false;                     // '' is a missing value (A.1)
        *id = (int64_t)(key % (uint64_t)p.col_buckets[c]);
        return true;
    }
    if (kind == WD_COL_VOCAB) {
        int i = vocab_lookup(p.vocab_fp + p.col_aux_off[c], p.col_aux_n[c], key);
        *id = i;
        return i >= 0;
    }
    // identity
    int64_t v = (int64_t)key;
    if (v == -1) return false;
    *id = (v < 0 || v >= p.col_buckets[c]) ? 0 : v;
    return true;
}

__device__ int simple_count(const DevPlan& p, const DevBatch& bt, int b, int c) {
    int kind = p.col_kind[c];
    if (kind == WD_COL_BUCKET) return 1;
    int s, e;
    field_range(bt, p.n_cat_fields, b, p.col_field[c], s, e);
    int n = 0;
    int64_t id;
    for (int j = s; j < e; ++j) n += simple_id(p, c, kind, bt.cat_keys[j], &id) ? 1 : 0;
    return n;
}

// k-th id of a simple column for row b (k < simple_count)
__device__ int64_t simple_kth(const DevPlan& p, const DevBatch& bt, int b, int c, int k) {
    int kind = p.col_kind[c];
    if (kind == WD_COL_BUCKET) {
        float x = bt.dense[b * p.n_dense_fields + p.col_field[c]];
        x = normalise(p.col_norm_kind[c], p.col_norm_a[c], p.col_norm_b[c], x);
        return bucketize(p.boundaries + p.col_aux_off[c], p.col_aux_n[c], x);
    }
    int s, e;
    field_range(bt, p.n_cat_fields, b, p.col_field[c], s, e);
    int64_t id = 0;
    for (int j = s; j < e; ++j) {
        if (simple_id(p, c, kind, bt.cat_keys[j], &id)) {
            if (k == 0) return id;
            --k;
        }
    }
    return id;
}

constexpr int kMaxCrossKeys = 8;

__device__ int cross_key_count(const DevPlan& p, const DevBatch& bt, int b, int type, int idx) {
    if (type == WD_KEY_FIELD) {
        int s, e;
        field_range(bt, p.n_cat_fields, b, idx, s, e);
        return e - s;                                          // dense string input: every entry incl. '' pads (Q2)
    }
    return simple_count(p, bt, b, idx);
}

__device__ uint64_t cross_key_value(const DevPlan& p, const DevBatch& bt, int b, int type, int idx, int k) {
    if (type == WD_KEY_FIELD) {
        int s, e;
        field_range(bt, p.n_cat_fields, b, idx, s, e);
        return bt.cat_keys[s + k];                             // Fingerprint64 of the string
    }
    return (uint64_t)simple_kth(p, bt, b, idx, k);             // integer ids enter the chain raw
}

__device__ int column_count(const DevPlan& p, const DevBatch& bt, int b, int c) {
    if (p.col_kind[c] != WD_COL_CROSS) return simple_count(p, bt, b, c);
    int off = p.col_aux_off[c], nk = p.col_aux_n[c];
    int total = 1;
    for (int i = 0; i < nk; ++i) total *= cross_key_count(p, bt, b, p.cross_key_type[off + i], p.cross_key_idx[off + i]);
    return total;
}

__global__ void ids_count_kernel(DevPlan p, DevBatch bt, int32_t* counts) {
    int64_t total = (int64_t)bt.B * p.n_columns;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(t / p.n_columns), c = (int)(t % p.n_columns);
        counts[t] = column_count(p, bt, b, c);
    }
}

// numeric deep columns (X0[b, off] = normalise(dense[b, field])) ride along in the same launch
struct NumericArgs { int n; const int32_t *field, *kind, *x0_off; const float *a, *bb; };

__global__ void ids_fill_kernel(DevPlan p, DevBatch bt, const int32_t* __restrict__ offs, int64_t cap,
                                uint32_t* e_wide, uint32_t* e_emb, int32_t* e_bc, int32_t* e_id, float* X0,
                                int32_t* flags, NumericArgs num) {
    if (X0 && num.n > 0) {
        const int64_t tn = (int64_t)bt.B * num.n;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tn; t += (int64_t)gridDim.x * blockDim.x) {
            const int b = (int)(t / num.n), i = (int)(t % num.n);
            const float x = bt.dense[(int64_t)b * p.n_dense_fields + num.field[i]];
            X0[(int64_t)b * p.d0_phys + num.x0_off[i]] = normalise(num.kind[i], num.a[i], num.bb[i], x);
        }
    }
    int64_t total = (int64_t)bt.B * p.n_columns;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(t / p.n_columns), c = (int)(t % p.n_columns);
        int base = offs[t], cnt = offs[t + 1] - base;
        const int64_t wbase = p.col_wide_base[c];
        const int tab = p.col_emb_table[c];
        const int ind = p.col_ind_off[c];
        float* xrow = X0 ? X0 + (int64_t)b * p.d0_phys : nullptr;
        if (ind >= 0 && xrow) {
            int nb = (int)p.col_buckets[c];
            for (int i = 0; i < nb; ++i) xrow[ind + i] = 0.f;
        }
        if ((int64_t)base + cnt > cap) {
            atomicOr(&flags[0], 1);                            // nnz capacity exceeded: fail the step loudly
            continue;
        }
        const int64_t ebase = tab >= 0 ? p.table_row_base[tab] : 0;
        // row-sharded tables: the entry addresses row id / G of rank id mod G instead of a replicated row
        const int she = p.sh_col_emb ? p.sh_col_emb[c] : -1, shw = p.sh_col_wide ? p.sh_col_wide[c] : -1;
        const int64_t she_base = she >= 0 ? p.sh_base_emb[she] : 0, shw_base = shw >= 0 ? p.sh_base_wide[shw] : 0;
        const uint32_t G = (uint32_t)p.sh_world;
        auto emit = [&](int j, int64_t id) {
            e_id[j] = (int32_t)id;
            e_bc[j] = (int32_t)t;
            e_wide[j] = (wbase >= 0 && shw < 0) ? (uint32_t)(wbase + id) : kInvalidRow;
            e_emb[j] = (tab >= 0 && she < 0) ? (uint32_t)(ebase + id) : kInvalidRow;
            if (p.sh_own_emb) {
                p.sh_own_emb[j] = she >= 0 ? (uint32_t)id % G : kInvalidRow;
                p.sh_lrow_emb[j] = (uint32_t)(she_base + (uint32_t)id / G);
            }
            if (p.sh_own_wide) {
                p.sh_own_wide[j] = shw >= 0 ? (uint32_t)id % G : kInvalidRow;
                p.sh_lrow_wide[j] = (uint32_t)(shw_base + (uint32_t)id / G);
            }
            if (ind >= 0 && xrow) xrow[ind + id] += 1.f;       // indicator_column: multi-hot counts (A.7)
        };
        const int kind = p.col_kind[c];
        if (kind == WD_COL_BUCKET) {
            if (cnt > 0) emit(base, simple_kth(p, bt, b, c, 0));
            continue;
        }
        if (kind != WD_COL_CROSS) {
            // one pass over the field's keys (a long multihot bag would otherwise be rescanned from its start for every id)
            int s, e, k = 0;
            field_range(bt, p.n_cat_fields, b, p.col_field[c], s, e);
            int64_t id;
            for (int j = s; j < e && k < cnt; ++j)
                if (simple_id(p, c, kind, bt.cat_keys[j], &id)) emit(base + k++, id);
            continue;
        }
        if (cnt == 0) continue;
        int off = p.col_aux_off[c], nk = p.col_aux_n[c];
        int idx[kMaxCrossKeys], kc[kMaxCrossKeys];
        for (int i = 0; i < nk; ++i) {
            idx[i] = 0;
            kc[i] = cross_key_count(p, bt, b, p.cross_key_type[off + i], p.cross_key_idx[off + i]);
        }
        const uint64_t nbuckets = (uint64_t)p.col_buckets[c];
        for (int j = 0; j < cnt; ++j) {
            uint64_t h = kCrossHashKey;
            for (int i = 0; i < nk; ++i)
                h = fingerprint_cat64(h, cross_key_value(p, bt, b, p.cross_key_type[off + i], p.cross_key_idx[off + i], idx[i]));
            emit(base + j, (int64_t)(h % nbuckets));
            for (int i = nk - 1; i >= 0; --i) {                // odometer: last key innermost
                if (++idx[i] < kc[i]) break;
                idx[i] = 0;
            }
        }
    }
}

// numeric deep columns: X0[b, off] = normalise(dense[b, field])
__global__ void numeric_kernel(DevBatch bt, int n_dense, int n_numeric, const int32_t* __restrict__ field,
                               const int32_t* __restrict__ kind, const float* __restrict__ a, const float* __restrict__ bb,
                               const int32_t* __restrict__ x0_off, float* X0, int ld) {
    int64_t total = (int64_t)bt.B * n_numeric;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(t / n_numeric), i = (int)(t % n_numeric);
        float x = bt.dense[(int64_t)b * n_dense + field[i]];
        X0[(int64_t)b * ld + x0_off[i]] = normalise(kind[i], a[i], bb[i], x);
    }
}

int ids_prepare(WdModel* m) {
    const DevBatch& bt = m->dbatch;
    int64_t total = (int64_t)bt.B * m->n_columns;
    if (total > 0) {
        ids_count_kernel<<<grid_for(total, 256), 256, 0, m->stream>>>(m->dplan, bt, m->d_col_offs);
        m->launches++;
    }
    int rc = exclusive_scan_i32(m, m->d_col_offs, total, m->d_nnz);
    if (rc) return rc;
    if (total > 0) {
        NumericArgs num{m->use_deep ? m->n_numeric : 0, m->d_num_field, m->d_num_norm_kind, m->d_num_x0_off, m->d_num_a, m->d_num_b};
        ids_fill_kernel<<<grid_for(total, 256), 256, 0, m->stream>>>(m->dplan, bt, m->d_col_offs, m->max_nnz, m->d_e_wide,
                                                                    m->d_e_emb, m->d_e_bc, m->d_e_id,
                                                                    m->use_deep ? m->d_X0 : nullptr, m->d_flags, num);
        m->launches++;
    }
    if (m->use_deep && m->n_numeric > 0 && total <= 0) {
        int64_t tn = (int64_t)bt.B * m->n_numeric;
        numeric_kernel<<<grid_for(tn, 256), 256, 0, m->stream>>>(bt, m->n_dense_fields, m->n_numeric, m->d_num_field,
                                                                m->d_num_norm_kind, m->d_num_a, m->d_num_b,
                                                                m->d_num_x0_off, m->d_X0, m->d0_phys);
        m->launches++;
    }
    WD_CUDA(cudaGetLastError());
    return WD_OK;
}

// ---- stand-alone Fingerprint64 over byte strings on the device (parity tests; loaders hash on the host)
__global__ void fingerprint_kernel(const uint8_t* bytes, const int64_t* offs, int64_t n, uint64_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fingerprint64(bytes + offs[i], (size_t)(offs[i + 1] - offs[i]));
}

}  // namespace wd

extern "C" int wd_fingerprint64_device(const uint8_t* bytes, const int64_t* offsets, int64_t n, uint64_t* out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        wd::set_error("no CUDA device");
        return WD_ENODEVICE;
    }
    if (n <= 0) return WD_OK;
    int64_t nbytes = offsets[n];
    uint8_t* d_b = nullptr;
    int64_t* d_o = nullptr;
    uint64_t* d_out = nullptr;
    WD_CUDA(cudaMalloc(&d_b, nbytes > 0 ? nbytes : 1));
    WD_CUDA(cudaMalloc(&d_o, (n + 1) * sizeof(int64_t)));
    WD_CUDA(cudaMalloc(&d_out, n * sizeof(uint64_t)));
    WD_CUDA(cudaMemcpy(d_b, bytes, nbytes, cudaMemcpyHostToDevice));
    WD_CUDA(cudaMemcpy(d_o, offsets, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    wd::fingerprint_kernel<<<wd::grid_for(n, 128), 128>>>(d_b, d_o, n, d_out);
    WD_CUDA(cudaGetLastError());
    WD_CUDA(cudaMemcpy(out, d_out, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    cudaFree(d_b); cudaFree(d_o); cudaFree(d_out);
    return WD_OK;
}
