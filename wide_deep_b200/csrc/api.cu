// C-ABI of libwd_b200 (include/wd_b200.h): model construction from the compiled plan, parameter IO,
// and the orchestration of one train / forward / eval step on the model's stream.
//
// A step replaces one sess.run(train_op) of the reference's Estimator.train loop (reference
// python/train.py:128-133 -> joint.py:81-269): ids -> sparse forward -> towers -> head -> backward ->
// optimizers.  Nothing here falls back to the CPU: without a CUDA device every entry point returns
// WD_ENODEVICE.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <random>

#include "common.cuh"
#include "farmhash.cuh"
#include "sparse_dev.cuh"

namespace wd {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int init_sparse_tables(WdModel* m, uint64_t seed, int random_w);
int dense_refresh_transposes(WdModel* m);
int wide_bias_grad(WdModel* m);
int metrics_setup();
int merge_sparse(WdModel* m, int which, const void* rows, const void* grads, int64_t n);
int shard_build(WdModel* m, const WdPlanDesc* d);
int shard_phase0(WdModel* m, bool train);
int shard_phase1(WdModel* m, bool train);
int shard_phase2(WdModel* m, bool train);
int shard_phase3(WdModel* m);
int shard_phase4(WdModel* m);
int shard_step_ipc(WdModel* m, bool train);

static int pad_to(int n, int k) { return (n + k - 1) / k * k; }
static int bits_for(int64_t n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

template <typename T>
static int upload_vec(WdModel* m, const T* src, int64_t n, const T** dst) {
    T* p = nullptr;
    int rc = dev_alloc(m, &p, n, true);
    if (rc) return rc;
    if (n > 0) WD_CUDA(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, m->stream));
    *dst = p;
    return WD_OK;
}

// sources of each layer input, in concat order (reference dnn.py:92-193); -1 = deep input x
static std::vector<std::vector<int>> layer_sources(int mode, int L) {
    std::vector<std::vector<int>> out;
    for (int l = 0; l < L; ++l) {
        std::vector<int> s;
        if (l == 0) s = {-1};
        else if (mode == WD_MODE_SIMPLE || mode == WD_MODE_LAST_DENSE) s = {l - 1};
        else if (mode == WD_MODE_FIRST_DENSE) s = {l - 1, -1};
        else if (mode == WD_MODE_DENSE) { s.push_back(-1); for (int j = 0; j < l; ++j) s.push_back(j); }
        else { for (int j = l - 1; j >= 0; --j) s.push_back(j); s.push_back(-1); }
        out.push_back(s);
    }
    std::vector<int> last;
    if (L == 0) last = {-1};
    else if (mode == WD_MODE_SIMPLE) last = {L - 1};
    else if (mode == WD_MODE_FIRST_DENSE) last = {L - 1, -1};
    else if (mode == WD_MODE_LAST_DENSE || mode == WD_MODE_DENSE) { last.push_back(-1); for (int j = 0; j < L; ++j) last.push_back(j); }
    else { for (int j = L - 1; j >= 0; --j) last.push_back(j); last.push_back(-1); }
    out.push_back(last);
    return out;
}
}  // namespace wd

using namespace wd;

struct WdModelExtra {   // host-only bookkeeping kept beside WdModel
    std::vector<uint8_t> x0_real;            // [d0_phys] 1 where a physical deep-input column is a real feature
    int d0_logical = 0;
    std::vector<std::vector<int>> dense_index;   // [tower-layer id][sub] -> index into WdModel::dense, -1
    std::vector<int> did_tower, did_layer;
};
static std::vector<std::pair<WdModel*, WdModelExtra*>> g_extra;
static WdModelExtra* extra_of(WdModel* m) {
    for (auto& p : g_extra) if (p.first == m) return p.second;
    return nullptr;
}

extern "C" const char* wd_last_error(void) { return wd::g_err; }
extern "C" int wd_version(void) { return WD_API_VERSION; }
extern "C" int wd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
extern "C" uint64_t wd_fingerprint64(const uint8_t* bytes, size_t n) { return wd::fingerprint64(bytes, n); }
extern "C" uint64_t wd_fingerprint_cat64(uint64_t a, uint64_t b) { return wd::fingerprint_cat64(a, b); }

namespace wd { void tc_map_cache_clear(); }

extern "C" int wd_model_destroy(WdModel* m) {
    if (!m) return WD_OK;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    tc_map_cache_clear();
    for (auto& sl : m->slots) {
        if (sl.graph) cudaGraphExecDestroy(sl.graph);
        if (sl.graph_bwd) cudaGraphExecDestroy(sl.graph_bwd);
        if (sl.ev_up) cudaEventDestroy(sl.ev_up);
        if (sl.ev_used) cudaEventDestroy(sl.ev_used);
    }
    if (m->stream_up) { cudaStreamSynchronize(m->stream_up); cudaStreamDestroy(m->stream_up); }
    for (auto& g : m->merge_graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (m->ev_bwd_done) cudaEventDestroy(m->ev_bwd_done);
    for (void* p : m->allocs) cudaFree(p);
    if (m->h_loss_pinned) cudaFreeHost(m->h_loss_pinned);
    for (auto& e : m->timer.ev) if (e) cudaEventDestroy(e);
    if (m->shard.aux) { cudaStreamSynchronize(m->shard.aux); cudaStreamDestroy(m->shard.aux); }
    for (cudaEvent_t ev : {m->shard.ev_a, m->shard.ev_ids2, m->shard.ev_routed1, m->shard.ev_a2, m->shard.ev_aux_done}) if (ev) cudaEventDestroy(ev);
    for (int w = 0; w < 2; ++w) {
        if (m->sstream[w]) { cudaStreamSynchronize(m->sstream[w]); cudaStreamDestroy(m->sstream[w]); }
        if (m->ev_grouped[w]) cudaEventDestroy(m->ev_grouped[w]);
        if (m->ev_done[w]) cudaEventDestroy(m->ev_done[w]);
    }
    if (m->ev_ids) cudaEventDestroy(m->ev_ids);
    if (m->ev_wide_fwd) cudaEventDestroy(m->ev_wide_fwd);
    if (m->ev_wgrad_rest) cudaEventDestroy(m->ev_wgrad_rest);
    if (m->ev_head) cudaEventDestroy(m->ev_head);
    if (m->ev_dx0) cudaEventDestroy(m->ev_dx0);
    if (m->stream) cudaStreamDestroy(m->stream);
    for (size_t i = 0; i < g_extra.size(); ++i)
        if (g_extra[i].first == m) { delete g_extra[i].second; g_extra.erase(g_extra.begin() + i); break; }
    delete m;
    return WD_OK;
}

// optimizer slots of the dense arena start at the initial accumulator value everywhere (padding included)
static int init_dense_slots(WdModel* m) {
    if (m->dense_count == 0) return WD_OK;
    std::vector<float> S1(m->dense_count, 0.f);
    for (size_t i = 0; i < m->dense.size(); ++i) {
        const WdOptimizer& o = (m->use_wide && i == 0) ? m->lin_opt : m->dnn_opt;
        float s1 = slot1_init(o);
        int64_t end = i + 1 < m->dense.size() ? m->dense[i + 1].off : m->dense_count;
        for (int64_t j = m->dense[i].off; j < end; ++j) S1[j] = s1;
    }
    WD_CUDA(cudaMemcpyAsync(m->d_S1, S1.data(), S1.size() * 4, cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaStreamSynchronize(m->stream));
    return WD_OK;
}

static int build_model(const WdPlanDesc* d, WdModel* m, WdModelExtra* x) {
    int rc;
    m->use_wide = d->model_type & 1;
    m->use_deep = (d->model_type & 2) != 0;
    m->n_cat_fields = d->n_cat_fields; m->n_dense_fields = d->n_dense_fields; m->n_columns = d->n_columns;
    const int C = d->n_columns;
    m->col_kind.assign(d->col_kind, d->col_kind + C);
    m->col_field.assign(d->col_field, d->col_field + C);
    m->col_buckets.assign(d->col_buckets, d->col_buckets + C);
    m->col_wide_base.assign(d->col_wide_base, d->col_wide_base + C);
    m->col_emb_table.assign(d->col_emb_table, d->col_emb_table + C);
    m->col_ind_off.assign(d->col_ind_off, d->col_ind_off + C);
    m->n_numeric = d->n_numeric;
    m->d0_phys = d->d0_phys; m->wide_rows = d->wide_rows;
    // crelu(z) = [relu(z) | relu(-z)]: the layer is built twice as wide with relu, its kernel / bias columns n + u tied to minus
    // columns n (exact: every product and sum just changes sign); see crelu_fold / crelu_mirror in mlp.cu
    m->crelu = d->activation == WD_ACT_CRELU;
    m->activation = m->crelu ? WD_ACT_RELU : d->activation; m->batch_norm = d->batch_norm;
    m->dropout_rate = d->dropout_rate; m->dropout_seed = d->dropout_seed;
    if (!(m->dropout_rate >= 0.f && m->dropout_rate < 1.f)) { set_error("dnn_dropout %g outside [0, 1)", m->dropout_rate); return WD_EINVAL; }
    m->lin_opt = d->lin_opt; m->dnn_opt = d->dnn_opt;
    m->max_batch = d->max_batch; m->max_batch_pad = pad_to(d->max_batch, 128);
    m->ldt = m->max_batch_pad;
    m->row_tiles = m->max_batch_pad / 128;
    m->gemm_engine = d->gemm_engine == WD_GEMM_AUTO ? WD_GEMM_TC3X : d->gemm_engine;
    m->dense_exchange_max_rows = d->dense_exchange_max_rows > 0 ? d->dense_exchange_max_rows : 0;
    m->small_base[1] = (m->dense_exchange_max_rows > 0 && d->wide_small_base >= 0 && d->wide_small_base <= d->wide_rows) ? d->wide_small_base : d->wide_rows;
    m->max_nnz = d->max_nnz > 0 ? d->max_nnz : (int64_t)d->max_batch * std::max(C, 1) * 2;
    const int G = d->shard_world > 1 ? d->shard_world : 1;
    m->shard.world = G; m->shard.rank = G > 1 ? d->shard_rank : 0;
    if (G > 1) {
        if (d->shard_rank < 0 || d->shard_rank >= G) { set_error("shard_rank %d outside [0, %d)", d->shard_rank, G); return WD_EINVAL; }
        // as an owner a rank can receive more ids than it routes itself (skewed ids): all per-entry scratch is sized for that
        const int64_t route = d->shard_capacity > 0 ? d->shard_capacity : m->max_nnz;
        const double slack = d->shard_slack >= 1.f ? d->shard_slack : 2.0;
        m->max_nnz = std::max<int64_t>(m->max_nnz, (int64_t)(route * slack));
    }
    m->keys_cap = d->max_keys > 0 ? d->max_keys : (int64_t)d->max_batch * std::max(d->n_cat_fields, 1) * 4;
    if (m->lin_opt.kind == WD_OPT_FTRL && m->lin_opt.lr_power != -0.5f) { set_error("FTRL: only learning_rate_power=-0.5 is supported"); return WD_EUNSUPPORTED; }
    if (m->dnn_opt.kind == WD_OPT_FTRL && m->dnn_opt.lr_power != -0.5f) { set_error("FTRL: only learning_rate_power=-0.5 is supported"); return WD_EUNSUPPORTED; }
    for (int c = 0; c < C; ++c)
        if (d->col_kind[c] == WD_COL_CROSS && d->col_aux_n[c] > 8) { set_error("cross with more than 8 keys"); return WD_EUNSUPPORTED; }

    // ---- plan tables on the device
    DevPlan& p = m->dplan;
    p.n_cat_fields = d->n_cat_fields; p.n_dense_fields = d->n_dense_fields; p.n_columns = C; p.d0_phys = d->d0_phys;
    if ((rc = upload_vec(m, d->cat_field_is_string, d->n_cat_fields, &p.field_is_string))) return rc;
    if ((rc = upload_vec(m, d->col_kind, C, &p.col_kind))) return rc;
    if ((rc = upload_vec(m, d->col_field, C, &p.col_field))) return rc;
    if ((rc = upload_vec(m, d->col_buckets, C, &p.col_buckets))) return rc;
    if ((rc = upload_vec(m, d->col_aux_off, C, &p.col_aux_off))) return rc;
    if ((rc = upload_vec(m, d->col_aux_n, C, &p.col_aux_n))) return rc;
    if ((rc = upload_vec(m, d->col_norm_kind, C, &p.col_norm_kind))) return rc;
    if ((rc = upload_vec(m, d->col_norm_a, C, &p.col_norm_a))) return rc;
    if ((rc = upload_vec(m, d->col_norm_b, C, &p.col_norm_b))) return rc;
    if ((rc = upload_vec(m, d->col_wide_base, C, &p.col_wide_base))) return rc;
    if ((rc = upload_vec(m, d->col_emb_table, C, &p.col_emb_table))) return rc;
    if ((rc = upload_vec(m, d->col_ind_off, C, &p.col_ind_off))) return rc;
    if ((rc = upload_vec(m, d->vocab_fp, d->n_vocab_fp, &p.vocab_fp))) return rc;
    if ((rc = upload_vec(m, d->boundaries, d->n_boundaries, &p.boundaries))) return rc;
    if ((rc = upload_vec(m, d->cross_key_type, d->n_cross_keys, &p.cross_key_type))) return rc;
    if ((rc = upload_vec(m, d->cross_key_idx, d->n_cross_keys, &p.cross_key_idx))) return rc;
    {
        const int32_t* t;
        const float* f;
        if ((rc = upload_vec(m, d->num_field, d->n_numeric, &t))) return rc; m->d_num_field = (int32_t*)t;
        if ((rc = upload_vec(m, d->num_norm_kind, d->n_numeric, &t))) return rc; m->d_num_norm_kind = (int32_t*)t;
        if ((rc = upload_vec(m, d->num_x0_off, d->n_numeric, &t))) return rc; m->d_num_x0_off = (int32_t*)t;
        if ((rc = upload_vec(m, d->num_norm_a, d->n_numeric, &f))) return rc; m->d_num_a = (float*)f;
        if ((rc = upload_vec(m, d->num_norm_b, d->n_numeric, &f))) return rc; m->d_num_b = (float*)f;
    }

    // ---- batch buffers
    const int64_t Bm = m->max_batch;
    if ((rc = dev_alloc(m, &m->d_cat_offsets, Bm * std::max(d->n_cat_fields, 1) + 1))) return rc;
    if ((rc = dev_alloc(m, &m->d_cat_keys, m->keys_cap))) return rc;
    if ((rc = dev_alloc(m, &m->d_dense, Bm * std::max(d->n_dense_fields, 1)))) return rc;
    if ((rc = dev_alloc(m, &m->d_label, Bm))) return rc;
    if ((rc = dev_alloc(m, &m->d_weight, Bm))) return rc;
    if ((rc = dev_alloc(m, &m->d_col_offs, Bm * std::max(C, 1) + 2))) return rc;
    if ((rc = dev_alloc(m, &m->d_e_wide, m->max_nnz))) return rc;
    if ((rc = dev_alloc(m, &m->d_e_emb, m->max_nnz))) return rc;
    if ((rc = dev_alloc(m, &m->d_e_bc, m->max_nnz))) return rc;
    if ((rc = dev_alloc(m, &m->d_e_id, m->max_nnz))) return rc;
    if ((rc = dev_alloc(m, &m->d_nnz, 4))) return rc;
    if ((rc = dev_alloc(m, &m->d_flags, 4))) return rc;
    for (int k = 0; k < 4; ++k) if ((rc = dev_alloc(m, &m->d_sort_counter_s[k], 4))) return rc;
    {
        int64_t n = std::max<int64_t>(Bm * std::max(C, 1) + 2, m->max_nnz + 2);
        for (int k = 0; k < 4; ++k) {
            int32_t* t;
            if ((rc = dev_alloc(m, &t, n / 4096 + 8))) return rc;
            m->d_scan_tmp_s[k] = t;
        }
    }
    if ((rc = dev_alloc(m, &m->d_logits, Bm))) return rc;
    if (G == 1 && (rc = dev_alloc(m, &m->d_dlogit, Bm))) return rc;
    if ((rc = dev_alloc(m, &m->d_loss_part, 512))) return rc;
    if ((rc = dev_alloc(m, &m->d_loss, 4))) return rc;
    if ((rc = dev_alloc(m, &m->d_head_counter, 4))) return rc;
    if (getenv("WD_STEP_TRACE") && (rc = dev_alloc(m, &m->d_step_trace, 16))) return rc;
    if ((rc = dev_alloc(m, &m->d_step, 4))) return rc;
    if ((rc = dev_alloc(m, &m->d_bpow, 4))) return rc;
    {
        const float bp[4] = {m->lin_opt.beta1, m->lin_opt.beta2, m->dnn_opt.beta1, m->dnn_opt.beta2};     // beta^1: state before the first step
        WD_CUDA(cudaMemcpyAsync(m->d_bpow, bp, sizeof(bp), cudaMemcpyHostToDevice, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
    }
    if ((rc = dev_alloc(m, &m->d_metrics, 512))) return rc;
    WD_CUDA(cudaMallocHost(&m->h_loss_pinned, 64));

    // ---- dense tensors
    auto add_dense = [&](int rows, int cols, int gparts, int g_rowtiles, int64_t gstride, bool transpose) {
        DenseTensor t{};
        t.off = m->dense_count; t.count = (int64_t)rows * cols; t.rows = rows; t.cols = cols;
        t.gpart_off = m->gpart_count; t.gparts = gparts; t.g_rowtiles = g_rowtiles; t.gstride = gstride;
        t.wt_off = transpose ? m->wt_count : -1;
        m->dense_count += (t.count + 3) / 4 * 4;
        m->gpart_count += (int64_t)gparts * gstride;
        if (transpose) m->wt_count += t.count;
        m->dense.push_back(t);
        return (int)m->dense.size() - 1;
    };
    if (m->use_wide) {
        add_dense(1, 1, m->row_tiles, 1, 4, false);            // [0] wide bias
        if ((rc = dev_alloc(m, &m->d_wide, m->wide_rows))) return rc;
        if ((rc = dev_alloc(m, &m->d_wide_logit, Bm))) return rc;
    }

    // ---- deep part
    x->x0_real.assign(std::max(d->d0_phys, 1), 0);
    if (m->use_deep) {
        int64_t row_base = 0;
        std::vector<int64_t> h_row_base;
        const int nslots = opt_nslots(m->dnn_opt);
        for (int t = 0; t < d->n_tables; ++t) {
            EmbTable tb{};
            tb.rows = d->table_rows[t]; tb.dim = d->table_dim[t]; tb.x0_off = d->table_x0_off[t];
            tb.dim_logical = d->table_dim_logical[t];
            if (tb.dim_logical < 1 || tb.dim_logical > tb.dim) { set_error("table %d: bad logical width", t); return WD_EINVAL; }
            tb.row_base = 0; tb.stride = tb.dim * (1 + nslots); tb.col = -1;
            tb.sharded = G > 1 && d->table_sharded && d->table_sharded[t];
            tb.arows = tb.sharded ? (tb.rows - m->shard.rank + G - 1) / G : tb.rows;
            if (tb.dim % 4 || tb.x0_off % 4) { set_error("table %d: dim and deep-input offset must be multiples of 4", t); return WD_EINVAL; }
            for (int c = 0; c < C; ++c) if (d->col_emb_table[c] == t) tb.col = c;
            if (tb.col < 0) { set_error("table %d has no producing column", t); return WD_EINVAL; }
            if ((rc = dev_alloc(m, &tb.data, tb.arows * tb.stride, true))) return rc;
            for (int i = 0; i < tb.dim_logical; ++i) x->x0_real[tb.x0_off + i] = 1;
            m->emb_max_dim = std::max(m->emb_max_dim, tb.dim);
            m->tables.push_back(tb);
        }
        // global row space: large tables first, then the small (dense-exchanged) ones, each group in table order
        std::vector<int> row_order;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) m->small_base[0] = row_base;
            for (int t = 0; t < d->n_tables; ++t) {
                if (m->tables[t].sharded) continue;                  // rows live in the sharded space (shard.cu), not here
                const bool small = m->dense_exchange_max_rows > 0 && m->tables[t].rows <= m->dense_exchange_max_rows;
                if (small != (pass == 1)) continue;
                if (small) m->n_small_tab++;
                m->tables[t].row_base = row_base;
                row_base += m->tables[t].rows;
                row_order.push_back(t);
            }
        }
        for (int t = 0; t < d->n_tables; ++t) h_row_base.push_back(m->tables[t].row_base);
        {
            std::vector<int64_t> rb, go; std::vector<float*> dt; std::vector<int32_t> dm, st;
            int64_t off = 0;
            for (int t : row_order) {
                const EmbTable& tb = m->tables[t];
                const bool small = m->dense_exchange_max_rows > 0 && tb.rows <= m->dense_exchange_max_rows;
                rb.push_back(tb.row_base); dt.push_back(tb.data); dm.push_back(tb.dim); st.push_back(tb.stride);
                go.push_back(small ? off : -1);
                if (small) off += tb.rows * tb.dim;
            }
            m->gs_emb_floats = off;
            m->n_rtab = (int)row_order.size();
            { const int64_t* t_; if ((rc = upload_vec(m, rb.data(), m->n_rtab, &t_))) return rc; m->d_rtab_row_base = (int64_t*)t_; }
            { const int64_t* t_; if ((rc = upload_vec(m, go.data(), m->n_rtab, &t_))) return rc; m->d_rtab_gs_off = (int64_t*)t_; }
            { float* const* t_; if ((rc = upload_vec<float*>(m, dt.data(), m->n_rtab, (float* const**)&t_))) return rc; m->d_rtab_data = (float**)t_; }
            { const int32_t* t_;
              if ((rc = upload_vec(m, dm.data(), m->n_rtab, &t_))) return rc; m->d_rtab_dim = (int32_t*)t_;
              if ((rc = upload_vec(m, st.data(), m->n_rtab, &t_))) return rc; m->d_rtab_stride = (int32_t*)t_; }
        }
        m->emb_total_rows = row_base;
        if (row_base >= (1ll << 31)) { set_error("more than 2^31 embedding rows on one device"); return WD_EUNSUPPORTED; }
        for (int i = 0; i < d->n_numeric; ++i) x->x0_real[d->num_x0_off[i]] = 1;
        for (int c = 0; c < C; ++c)
            if (d->col_ind_off[c] >= 0) for (int64_t i = 0; i < d->col_buckets[c]; ++i) x->x0_real[d->col_ind_off[c] + i] = 1;
        for (auto v : x->x0_real) x->d0_logical += v;
        // device table descriptors
        const int nt = (int)m->tables.size();
        std::vector<float*> h_data; std::vector<int32_t> h_dim, h_stride, h_x0, h_col;
        for (auto& tb : m->tables) { h_data.push_back(tb.data); h_dim.push_back(tb.dim); h_stride.push_back(tb.stride); h_x0.push_back(tb.x0_off); h_col.push_back(tb.col); }
        { float* const* t; if ((rc = upload_vec<float*>(m, h_data.data(), nt, (float* const**)&t))) return rc; m->d_tab_data = (float**)t; }
        { const int32_t* t;
          if ((rc = upload_vec(m, h_dim.data(), nt, &t))) return rc; m->d_tab_dim = (int32_t*)t;
          if ((rc = upload_vec(m, h_stride.data(), nt, &t))) return rc; m->d_tab_stride = (int32_t*)t;
          if ((rc = upload_vec(m, h_x0.data(), nt, &t))) return rc; m->d_tab_x0 = (int32_t*)t;
          if ((rc = upload_vec(m, h_col.data(), nt, &t))) return rc; m->d_tab_col = (int32_t*)t; }
        { const int64_t* t; if ((rc = upload_vec(m, h_row_base.data(), nt, &t))) return rc; m->d_tab_row_base = (int64_t*)t; p.table_row_base = t; }
        // group tables by width
        for (int t = 0; t < nt; ++t) {
            if (m->tables[t].sharded) continue;                      // gathered by their owners, not by the local gather kernels
            int di = -1;
            for (int i = 0; i < m->n_dims; ++i) if (m->dims[i] == m->tables[t].dim) di = i;
            if (di < 0) {
                if (m->n_dims == kMaxDims) { set_error("more than %d distinct embedding widths", kMaxDims); return WD_EUNSUPPORTED; }
                di = m->n_dims++; m->dims[di] = m->tables[t].dim; m->dim_ntables[di] = 0;
            }
            m->dim_ntables[di]++;
        }
        for (int i = 0; i < m->n_dims; ++i) {
            std::vector<int32_t> ids;
            for (int t = 0; t < nt; ++t) if (m->tables[t].dim == m->dims[i] && !m->tables[t].sharded) ids.push_back(t);
            const int32_t* dp;
            if ((rc = upload_vec(m, ids.data(), (int64_t)ids.size(), &dp))) return rc;
            m->d_dim_tables[i] = (int32_t*)dp;
            std::vector<TabDesc> descs;
            for (int t : ids) descs.push_back(TabDesc{m->tables[t].data, m->tables[t].row_base, m->tables[t].stride, m->tables[t].x0_off, m->tables[t].col, m->tables[t].dim});
            const TabDesc* dd;
            if ((rc = upload_vec(m, descs.data(), (int64_t)descs.size(), &dd))) return rc;
            m->d_dim_desc[i] = (TabDesc*)dd;
        }
        const int64_t actn = (int64_t)m->max_batch_pad * d->d0_phys;
        if ((rc = dev_alloc(m, &m->d_X0, actn))) return rc;
        if ((rc = dev_alloc(m, &m->d_X0T, actn))) return rc;
        if (G == 1 && (rc = dev_alloc(m, &m->d_dX0, actn))) return rc;      // (sharded runs: inside the exchange segment, peers read it)
        if (m->gemm_engine == WD_GEMM_BF16X3)
            for (int part = 0; part < 2; ++part) {
                if ((rc = dev_alloc(m, &m->d_X0s[part], actn))) return rc;
            }

        // towers
        int hu_off = 0, did = 0;
        for (int t = 0; t < d->n_towers; ++t) {
            Tower tw{};
            tw.n_hidden = d->tower_nlayers[t]; tw.mode = d->tower_mode[t];
            std::vector<int> hu(d->hidden_units + hu_off, d->hidden_units + hu_off + tw.n_hidden);
            hu_off += tw.n_hidden;
            const std::vector<int> units = hu;                         // the conf's units per layer
            if (m->crelu) for (int& h : hu) h *= 2;                    // features a layer hands on
            auto srcs = layer_sources(tw.mode, tw.n_hidden);
            for (int l = 0; l <= tw.n_hidden; ++l) {
                Layer L{};
                if ((int)srcs[l].size() > kMaxSegs) { set_error("layer with more than %d concatenated inputs", kMaxSegs); return WD_EUNSUPPORTED; }
                L.n_in_segs = (int)srcs[l].size();
                int koff = 0, klog = 0;
                for (int s = 0; s < L.n_in_segs; ++s) {
                    int src = srcs[l][s];
                    Seg sg{};
                    sg.src = src;
                    sg.width = src < 0 ? x->d0_logical : hu[src];
                    sg.width_phys = src < 0 ? d->d0_phys : pad_to(hu[src], 32);
                    sg.k_off = koff;
                    koff += sg.width_phys; klog += sg.width;
                    L.segs[s] = sg;
                }
                L.K = klog; L.K_phys = koff;
                const bool hidden = l < tw.n_hidden;
                L.N = hidden ? hu[l] : 1;
                L.N_phys = hidden ? pad_to(hu[l], 32) : 1;
                L.N_param = hidden ? units[l] : 1;
                L.t_gamma = L.t_beta = -1;
                L.h_fp32 = false;
                for (int src : srcs[tw.n_hidden]) if (src == l) L.h_fp32 = true;      // read by the logits layer
                std::vector<int> idx(4, -1);
                if (hidden) {
                    // split-K factor of the weight gradient: enough (tile x split) work items to fill one wave of SMs,
                    // each split still at least 512 batch rows long
                    {
                        // (the 3xBF16 engine covers N in 256-wide tiles when N allows it)
                        const int tn = (m->gemm_engine == WD_GEMM_BF16X3 && L.N_phys % 256 == 0) ? 256 : 128;
                        const int tiles = ((L.K_phys + 127) / 128) * ((L.N_phys + tn - 1) / tn);
                        int sp = m->wgrad_splits;
                        while (sp < 32 && tiles * sp * 2 <= 148 && m->max_batch_pad / (sp * 2) >= 512) sp *= 2;   // double only while one wave still holds it
                        L.wgrad_splits = sp;
                    }
                    L.t_kernel = add_dense(L.K_phys, L.N_phys, L.wgrad_splits, 0, (int64_t)L.K_phys * L.N_phys, true);
                    L.t_bias = add_dense(1, L.N_phys, m->row_tiles, 1, L.N_phys, false);
                    if (m->crelu) m->dense[L.t_kernel].mirror_u = m->dense[L.t_bias].mirror_u = L.N_param;
                    if (m->batch_norm) {
                        L.t_gamma = add_dense(1, L.N_phys, m->row_tiles, 1, L.N_phys, false);
                        L.t_beta = add_dense(1, L.N_phys, m->row_tiles, 1, L.N_phys, false);
                    }
                    const int64_t n = (int64_t)m->max_batch_pad * L.N_phys;
                    if ((rc = dev_alloc(m, &L.H, n))) return rc;
                    // post-activation values are kept apart from the layer output whenever something sits between them (BN affine, dropout)
                    if (m->batch_norm || m->dropout_rate > 0.f) { if ((rc = dev_alloc(m, &L.A, n))) return rc; } else L.A = L.H;
                    if ((rc = dev_alloc(m, &L.HT, n))) return rc;
                    if ((rc = dev_alloc(m, &L.dH, n))) return rc;
                    if ((rc = dev_alloc(m, &L.dZ, n))) return rc;
                    if ((rc = dev_alloc(m, &L.dZT, n))) return rc;
                    if (m->gemm_engine == WD_GEMM_BF16X3)
                        for (int part = 0; part < 2; ++part) {
                            if ((rc = dev_alloc(m, &L.Hs[part], n))) return rc;
                            if ((rc = dev_alloc(m, &L.dZs[part], n))) return rc;
                        }
                } else {
                    L.t_kernel = add_dense(L.K_phys, 1, m->row_tiles, 1, L.K_phys, false);
                    L.t_bias = add_dense(1, 1, m->row_tiles, 1, 4, false);
                }
                idx[WD_D_KERNEL] = L.t_kernel; idx[WD_D_BIAS] = L.t_bias; idx[WD_D_GAMMA] = L.t_gamma; idx[WD_D_BETA] = L.t_beta;
                x->dense_index.push_back(idx);
                x->did_tower.push_back(t); x->did_layer.push_back(l);
                ++did;
                tw.layers.push_back(L);
            }
            if ((rc = dev_alloc(m, &tw.logit, Bm))) return rc;
            m->towers.push_back(tw);
        }
    }
    {
        // block layout: [embedding gradients | wide gradients | touched counts of the small embedding rows | ... of the small wide rows]
        const int64_t nw_small = m->use_wide ? m->wide_rows - m->small_base[1] : 0;
        const int64_t ne_small = (m->use_deep && m->n_small_tab > 0) ? m->emb_total_rows - m->small_base[0] : 0;
        const int64_t grads = m->gs_emb_floats + nw_small;
        m->gs_touch_off[0] = grads;
        m->gs_touch_off[1] = grads + ne_small;
        m->gs_count = grads > 0 ? grads + ne_small + nw_small : 0;
    }
    if (m->dense_count > 0) {
        if ((rc = dev_alloc(m, &m->d_P, m->dense_count))) return rc;
        if ((rc = dev_alloc(m, &m->d_S1, m->dense_count))) return rc;
        if ((rc = dev_alloc(m, &m->d_S2, m->dense_count))) return rc;
        if (G == 1 && (rc = dev_alloc(m, &m->d_G, m->dense_count + m->gs_count))) return rc;
        if ((rc = dev_alloc(m, &m->d_gpart, m->gpart_count))) return rc;
        if ((rc = dev_alloc(m, &m->d_Wt, std::max<int64_t>(m->wt_count, 1)))) return rc;
        if ((rc = dev_alloc(m, &m->d_Wsplit, std::max<int64_t>(4 * m->wt_count, 1)))) return rc;
        const DenseTensor* dp;
        if ((rc = upload_vec(m, m->dense.data(), (int64_t)m->dense.size(), &dp))) return rc;
        m->d_dense_desc = (DenseTensor*)dp;
    }

    // ---- sparse backward scratch
    m->sort_bits[0] = bits_for(std::max<int64_t>(m->emb_total_rows, 2));
    m->sort_bits[1] = bits_for(std::max<int64_t>(m->wide_rows, 2));
    if (m->sort_bits[0] > 30 || m->sort_bits[1] > 30) { set_error("more than 2^30 rows in one table space on one device"); return WD_EUNSUPPORTED; }
    const int which_lo = (m->use_deep && !m->tables.empty()) ? 0 : 1, which_hi = m->use_wide ? 1 : 0;
    for (int w = which_lo; w <= which_hi; ++w) {
        if ((rc = dev_alloc(m, &m->d_sk[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_sv[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_sk2[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_sv2[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_urow[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_ustart[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_ugrad[w], (m->max_nnz + 8) * (w == 0 ? std::max(m->emb_max_dim, 4) : 1)))) return rc;
        if ((rc = dev_alloc(m, &m->d_nuniq[w], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_nubig[w], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_nvalid[w], 4))) return rc;
        m->cpart_cap = 2 * (m->max_nnz / 16) + 64;        // kChunk = 16 (sparse.cu)
        if ((rc = dev_alloc(m, &m->d_choff[w], m->max_nnz + 8))) return rc;
        if ((rc = dev_alloc(m, &m->d_nchunks[w], 4))) return rc;
        if ((rc = dev_alloc(m, &m->d_cpart[w], m->cpart_cap * (w == 0 ? std::max(m->emb_max_dim, 4) : 1)))) return rc;
        m->sparse_cap[w] = m->max_nnz;
    }
    m->sort_hist_cap = 1024 * ((m->max_nnz + kSortTile - 1) / kSortTile + 1) + 4 * 1024 + 64;
    for (int k = 0; k < 4; ++k) if ((rc = dev_alloc(m, &m->d_sort_hist_s[k], m->sort_hist_cap))) return rc;
    if (G > 1) {
        if ((rc = shard_build(m, d))) return rc;
        DevPlan& dp = m->dplan;
        dp.sh_world = G;
        for (int sx = 0; sx < 2; ++sx) {
            ShardSpace& sp = m->shard.sp[sx];
            if (!sp.on) continue;
            if (sx == 0) { dp.sh_col_emb = sp.d_col_slot; dp.sh_base_emb = sp.d_slot_base; dp.sh_own_emb = sp.d_own; dp.sh_lrow_emb = sp.d_lrow; }
            else { dp.sh_col_wide = sp.d_col_slot; dp.sh_base_wide = sp.d_slot_base; dp.sh_own_wide = sp.d_own; dp.sh_lrow_wide = sp.d_lrow; }
        }
    }
    if ((rc = metrics_setup())) return rc;
    if ((rc = init_sparse_tables(m, 0, 0))) return rc;           // slots = initial accumulator, weights 0
    if ((rc = init_dense_slots(m))) return rc;
    for (auto& e : m->timer.ev) WD_CUDA(cudaEventCreate(&e));
    WD_CUDA(cudaStreamSynchronize(m->stream));
    return WD_OK;
}

extern "C" int wd_model_create(const WdPlanDesc* d, int device, WdModel** out) {
    if (!d || !out) { set_error("null argument"); return WD_EINVAL; }
    if (d->api_version != WD_API_VERSION) { set_error("plan api_version %d != library %d", d->api_version, WD_API_VERSION); return WD_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libwd_b200 has no CPU fallback");
        return WD_ENODEVICE;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return WD_EINVAL; }
    WD_CUDA(cudaSetDevice(device));
    WdModel* m = new WdModel();
    WdModelExtra* x = new WdModelExtra();
    g_extra.push_back({m, x});
    m->device = device;
    memset(m->timer.ev, 0, sizeof(m->timer.ev));
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream_up, cudaStreamNonBlocking);
    for (int w = 0; w < 2; ++w) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->sstream[w], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_grouped[w], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_done[w], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_ids, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_wide_fwd, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_wgrad_rest, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_head, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_dx0, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_bwd_done, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); wd_model_destroy(m); return WD_ECUDA; }
    m->graphs_enabled = getenv("WD_NO_GRAPH") == nullptr;
    int rc = build_model(d, m, x);
    if (rc) { wd_model_destroy(m); return rc; }
    *out = m;
    return WD_OK;
}

// glorot-uniform kernels / zero biases / gamma 1 / beta 0 built on the host (dense part is ~1.5M floats)
static int init_dense(WdModel* m, WdModelExtra* x, uint64_t seed) {
    if (m->dense_count == 0) return WD_OK;
    std::vector<float> P(m->dense_count, 0.f), S1(m->dense_count, 0.f), S2(m->dense_count, 0.f);
    std::mt19937_64 rng(seed * 7919 + 17);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    for (size_t ti = 0; ti < m->towers.size(); ++ti) {
        Tower& tw = m->towers[ti];
        for (int l = 0; l <= tw.n_hidden; ++l) {
            Layer& L = tw.layers[l];
            const DenseTensor& tk = m->dense[L.t_kernel];
            const float lim = std::sqrt(6.f / (float)(L.K + L.N_param));          // glorot_uniform over the variable's shape [K, N_param]
            for (int s = 0; s < L.n_in_segs; ++s) {
                const Seg& sg = L.segs[s];
                for (int j = 0; j < sg.width_phys; ++j) {
                    bool real = sg.src < 0 ? (x->x0_real[j] != 0) : (j < sg.width);
                    if (!real) continue;
                    float* prow = &P[tk.off + (int64_t)(sg.k_off + j) * L.N_phys];
                    for (int n = 0; n < L.N_param; ++n) {
                        prow[n] = lim * U(rng);
                        if (tk.mirror_u) prow[n + tk.mirror_u] = -prow[n];
                    }
                }
            }
            if (L.t_gamma >= 0) for (int n = 0; n < L.N; ++n) P[m->dense[L.t_gamma].off + n] = 1.f;
        }
    }
    for (size_t i = 0; i < m->dense.size(); ++i) {
        const WdOptimizer& o = (m->use_wide && i == 0) ? m->lin_opt : m->dnn_opt;
        float s1 = slot1_init(o);
        for (int64_t j = 0; j < m->dense[i].count; ++j) S1[m->dense[i].off + j] = s1;
    }
    WD_CUDA(cudaMemcpyAsync(m->d_P, P.data(), P.size() * 4, cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaMemcpyAsync(m->d_S1, S1.data(), S1.size() * 4, cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaMemcpyAsync(m->d_S2, S2.data(), S2.size() * 4, cudaMemcpyHostToDevice, m->stream));
    WD_CUDA(cudaStreamSynchronize(m->stream));
    return dense_refresh_transposes(m);
}

extern "C" int wd_model_init(WdModel* m, uint64_t seed) {
    if (!m) { set_error("null model"); return WD_EINVAL; }
    WD_CUDA(cudaSetDevice(m->device));
    int rc = init_sparse_tables(m, seed, 1);
    if (rc) return rc;
    rc = init_dense(m, extra_of(m), seed);
    if (rc) return rc;
    WD_CUDA(cudaStreamSynchronize(m->stream));
    m->initialized = true;
    return WD_OK;
}

// ---------------------------------------------------------------------------------------------- tensor IO
static int resolve_dense(WdModel* m, WdModelExtra* x, int did, int sub, int* out) {
    if (did < 0 || did >= (int)x->dense_index.size() || sub < 0 || sub > 3 || x->dense_index[did][sub] < 0) {
        set_error("no dense tensor (%d, %d)", did, sub);
        return WD_EINVAL;
    }
    *out = x->dense_index[did][sub];
    return WD_OK;
}

extern "C" int64_t wd_tensor_size(WdModel* m, int kind, int index, int sub) {
    if (!m) return WD_EINVAL;
    WdModelExtra* x = extra_of(m);
    if (kind == WD_T_WIDE_COL) {
        if (index < 0 || index >= m->n_columns) return WD_EINVAL;
        const ShardSpace& sw = m->shard.sp[1];
        if (sw.on && sw.h_col_slot[index] >= 0) return (m->col_buckets[index] - m->shard.rank + m->shard.world - 1) / m->shard.world;   // this rank's rows
        return m->col_wide_base[index] >= 0 ? m->col_buckets[index] : WD_EINVAL;
    }
    if (kind == WD_T_WIDE_BIAS) return m->use_wide ? 1 : WD_EINVAL;
    if (kind == WD_T_EMB_TABLE) return (index >= 0 && index < (int)m->tables.size()) ? m->tables[index].arows * m->tables[index].dim_logical : WD_EINVAL;
    if (kind == WD_T_DENSE) {
        int di;
        if (resolve_dense(m, x, index, sub, &di)) return WD_EINVAL;
        Layer& L = m->towers[x->did_tower[index]].layers[x->did_layer[index]];
        return sub == WD_D_KERNEL ? (int64_t)L.K * L.N_param : (sub == WD_D_BIAS ? L.N_param : L.N);
    }
    return WD_EINVAL;
}

extern "C" int wd_tensor_io(WdModel* m, int kind, int index, int sub, int slot, void* host, int64_t count, int to_device) {
    if (!m || !host) { set_error("null argument"); return WD_EINVAL; }
    WD_CUDA(cudaSetDevice(m->device));
    WD_CUDA(cudaStreamSynchronize(m->stream));
    WdModelExtra* x = extra_of(m);
    int64_t want = wd_tensor_size(m, kind, index, sub);
    if (want < 0 || want != count) { set_error("tensor (%d,%d,%d): size %lld, caller passed %lld", kind, index, sub, (long long)want, (long long)count); return WD_EINVAL; }
    if (slot < 0 || slot > 2) { set_error("slot out of range"); return WD_EINVAL; }
    const cudaMemcpyKind dir = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (kind == WD_T_WIDE_COL) {
        const ShardSpace& sw = m->shard.sp[1];
        float* dev = (sw.on && sw.h_col_slot[index] >= 0)
                         ? reinterpret_cast<float*>(sw.d_wide + sw.h_slot_base[sw.h_col_slot[index]]) + slot     // rows r = rank, rank + G, ...
                         : reinterpret_cast<float*>(m->d_wide + m->col_wide_base[index]) + slot;
        if (to_device) WD_CUDA(cudaMemcpy2DAsync(dev, 16, host, 4, 4, count, dir, m->stream));
        else WD_CUDA(cudaMemcpy2DAsync(host, 4, dev, 16, 4, count, dir, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
        return WD_OK;
    }
    if (kind == WD_T_EMB_TABLE) {
        EmbTable& tb = m->tables[index];
        if (slot * tb.dim >= tb.stride) { set_error("table has no optimizer slot %d", slot); return WD_EINVAL; }
        float* dev = tb.data + slot * tb.dim;
        const size_t lw = (size_t)tb.dim_logical * 4;
        if (to_device) WD_CUDA(cudaMemcpy2DAsync(dev, (size_t)tb.stride * 4, host, lw, lw, tb.arows, dir, m->stream));
        else WD_CUDA(cudaMemcpy2DAsync(host, lw, dev, (size_t)tb.stride * 4, lw, tb.arows, dir, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
        return WD_OK;
    }
    float* arena = slot == 0 ? m->d_P : (slot == 1 ? m->d_S1 : m->d_S2);
    if (kind == WD_T_WIDE_BIAS) {
        WD_CUDA(cudaMemcpyAsync(to_device ? (void*)(arena + m->dense[0].off) : host, to_device ? host : (void*)(arena + m->dense[0].off), 4, dir, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
        return WD_OK;
    }
    int di;
    int rc = resolve_dense(m, x, index, sub, &di);
    if (rc) return rc;
    const DenseTensor& t = m->dense[di];
    Layer& L = m->towers[x->did_tower[index]].layers[x->did_layer[index]];
    std::vector<float> phys(t.count, 0.f);
    float* h = (float*)host;
    if (!to_device || slot > 0) {
        WD_CUDA(cudaMemcpyAsync(phys.data(), arena + t.off, t.count * 4, cudaMemcpyDeviceToHost, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
    }
    if (sub == WD_D_KERNEL) {
        int64_t lk = 0;
        for (int s = 0; s < L.n_in_segs; ++s) {
            const Seg& sg = L.segs[s];
            for (int j = 0; j < sg.width_phys; ++j) {
                bool real = sg.src < 0 ? (x->x0_real[j] != 0) : (j < sg.width);
                if (!real) continue;
                for (int n = 0; n < L.N_param; ++n) {
                    float& pv = phys[(int64_t)(sg.k_off + j) * L.N_phys + n];
                    if (to_device) {
                        pv = h[lk * L.N_param + n];
                        if (t.mirror_u) (&pv)[t.mirror_u] = slot == 0 ? -pv : pv;       // tied half: minus the weight, same slot value
                    } else h[lk * L.N_param + n] = pv;
                }
                ++lk;
            }
        }
        if (lk != L.K) { set_error("internal: logical K mismatch %lld vs %d", (long long)lk, L.K); return WD_ESTATE; }
    } else {
        const int nn = sub == WD_D_BIAS ? L.N_param : L.N;
        for (int n = 0; n < nn; ++n) {
            if (to_device) {
                phys[n] = h[n];
                if (t.mirror_u) phys[n + t.mirror_u] = slot == 0 ? -h[n] : h[n];
            } else h[n] = phys[n];
        }
    }
    if (to_device) {
        // all copies go through the model stream: a pageable cudaMemcpy on the NULL stream may still be in flight when a
        // kernel on this (non-blocking) stream starts
        WD_CUDA(cudaMemcpyAsync(arena + t.off, phys.data(), t.count * 4, cudaMemcpyHostToDevice, m->stream));
        WD_CUDA(cudaStreamSynchronize(m->stream));
        if (slot == 0 && t.wt_off >= 0) return dense_refresh_transposes(m);
    }
    return WD_OK;
}

// -------------------------------------------------------------------------------------------------- steps
static int check_ready(WdModel* m) {
    if (!m) { set_error("null model"); return WD_EINVAL; }
    WD_CUDA(cudaSetDevice(m->device));
    return WD_OK;
}
static void timer_begin(WdModel* m) {
    if (!m->timer.enabled) return;
    m->timer.n = 0;
    mark(m, "start");
}

// make batch slot `s` current (allocating its buffers on first use); slot 0 aliases the model's own buffers
static int ensure_slot(WdModel* m, int s) {
    if (s < 0 || s >= 64) { set_error("batch slot %d out of range [0, 64)", s); return WD_EINVAL; }
    if (m->slots.empty()) {
        BatchSlot b0;
        b0.off = m->d_cat_offsets; b0.keys = m->d_cat_keys; b0.dense = m->d_dense; b0.label = m->d_label; b0.weight = m->d_weight;
        m->slots.push_back(b0);
    }
    while ((int)m->slots.size() <= s) {
        BatchSlot b;
        int rc;
        const int64_t Bm = m->max_batch;
        if ((rc = dev_alloc(m, &b.off, Bm * std::max(m->n_cat_fields, 1) + 1))) return rc;
        if ((rc = dev_alloc(m, &b.keys, m->keys_cap))) return rc;
        if ((rc = dev_alloc(m, &b.dense, Bm * std::max(m->n_dense_fields, 1)))) return rc;
        if ((rc = dev_alloc(m, &b.label, Bm))) return rc;
        if ((rc = dev_alloc(m, &b.weight, Bm))) return rc;
        m->slots.push_back(b);
    }
    return WD_OK;
}
static int select_slot(WdModel* m, int s) {
    int rc = ensure_slot(m, s);
    if (rc) return rc;
    BatchSlot& b = m->slots[s];
    m->cur_slot = s;
    m->d_cat_offsets = b.off; m->d_cat_keys = b.keys; m->d_dense = b.dense; m->d_label = b.label; m->d_weight = b.weight;
    if (b.filled) { m->dbatch = b.view; m->batch_has_label = b.has_label; }
    if (b.up_pending) {                                     // a prefetch refilled this slot on the upload stream
        WD_CUDA(cudaStreamWaitEvent(m->stream, b.ev_up, 0));
        b.up_pending = false;
    }
    return WD_OK;
}
// the model stream has consumed the current slot up to here (a later prefetch into it must wait for this point)
static int mark_slot_used(WdModel* m) {
    if (m->slots.empty()) return WD_OK;
    BatchSlot& b = m->slots[m->cur_slot];
    if (!b.ev_used) WD_CUDA(cudaEventCreateWithFlags(&b.ev_used, cudaEventDisableTiming));
    WD_CUDA(cudaEventRecord(b.ev_used, m->stream));
    b.used_recorded = true;
    return WD_OK;
}

// copies a host batch into the buffers of slot `sl` on stream `st`; fills the device view of the batch
static int upload_into(WdModel* m, BatchSlot& sl, const WdBatch* b, cudaStream_t st, DevBatch* view, bool* has_label) {
    if (!b || b->batch_size <= 0 || b->batch_size > m->max_batch) { set_error("batch_size %d outside (0, %d]", b ? b->batch_size : -1, m->max_batch); return WD_EINVAL; }
    const int B = b->batch_size, F = m->n_cat_fields, Nd = m->n_dense_fields;
    int64_t nnz = b->cat_offsets ? b->nnz : (int64_t)B * F;
    if (nnz > m->keys_cap) { set_error("batch has %lld keys, capacity %lld (raise max_keys)", (long long)nnz, (long long)m->keys_cap); return WD_EINVAL; }
    if (F > 0) {
        if (b->cat_offsets) WD_CUDA(cudaMemcpyAsync(sl.off, b->cat_offsets, ((int64_t)B * F + 1) * 4, cudaMemcpyHostToDevice, st));
        if (nnz > 0) WD_CUDA(cudaMemcpyAsync(sl.keys, b->cat_keys, nnz * 8, cudaMemcpyHostToDevice, st));
    }
    if (Nd > 0) WD_CUDA(cudaMemcpyAsync(sl.dense, b->dense, (int64_t)B * Nd * 4, cudaMemcpyHostToDevice, st));
    if (b->label) WD_CUDA(cudaMemcpyAsync(sl.label, b->label, (int64_t)B * 4, cudaMemcpyHostToDevice, st));
    if (b->weight) WD_CUDA(cudaMemcpyAsync(sl.weight, b->weight, (int64_t)B * 4, cudaMemcpyHostToDevice, st));
    view->B = B;
    view->cat_offsets = (F > 0 && b->cat_offsets) ? sl.off : nullptr;
    view->cat_keys = sl.keys;
    view->dense = sl.dense;
    view->label = b->label ? sl.label : nullptr;
    view->weight = b->weight ? sl.weight : nullptr;
    *has_label = b->label != nullptr;
    return WD_OK;
}

static int upload_current(WdModel* m, const WdBatch* b) {
    BatchSlot& sl = m->slots[m->cur_slot];
    int rc = upload_into(m, sl, b, m->stream, &m->dbatch, &m->batch_has_label);
    if (rc) return rc;
    mark(m, "h2d");
    return WD_OK;
}

// Asynchronous refill of a batch slot: the copies run on the library's upload stream, behind the last step that read the slot
// and concurrently with whatever the model stream is doing (normally: the step on another slot).  The next step on this slot
// waits for them on the device.  Host buffers must stay untouched until that step has been issued and returned.
extern "C" int wd_batch_prefetch_slot(WdModel* m, int slot, const WdBatch* b) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = ensure_slot(m, slot))) return rc;
    BatchSlot& sl = m->slots[slot];
    if (!sl.ev_up) WD_CUDA(cudaEventCreateWithFlags(&sl.ev_up, cudaEventDisableTiming));
    if (sl.used_recorded) WD_CUDA(cudaStreamWaitEvent(m->stream_up, sl.ev_used, 0));
    DevBatch view{};
    bool has_label = false;
    if ((rc = upload_into(m, sl, b, m->stream_up, &view, &has_label))) return rc;
    WD_CUDA(cudaEventRecord(sl.ev_up, m->stream_up));
    sl.view = view; sl.has_label = has_label; sl.filled = true; sl.up_pending = true;
    if (slot == m->cur_slot) { m->dbatch = view; m->batch_has_label = has_label; }
    return WD_OK;
}

extern "C" int wd_batch_upload_slot(WdModel* m, int slot, const WdBatch* b) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, slot))) return rc;
    if ((rc = upload_current(m, b))) return rc;
    m->slots[slot].view = m->dbatch;
    m->slots[slot].has_label = m->batch_has_label;
    m->slots[slot].filled = true;
    return WD_OK;
}
extern "C" int wd_batch_upload(WdModel* m, const WdBatch* b) { return wd_batch_upload_slot(m, 0, b); }

static int finish_step(WdModel* m, float* loss_out, float* logits_out) {
    int32_t* flags_host = reinterpret_cast<int32_t*>(m->h_loss_pinned + 4);
    WD_CUDA(cudaMemcpyAsync(m->h_loss_pinned, m->d_loss, 4, cudaMemcpyDeviceToHost, m->stream));
    WD_CUDA(cudaMemcpyAsync(flags_host, m->d_flags, 4, cudaMemcpyDeviceToHost, m->stream));
    if (logits_out) WD_CUDA(cudaMemcpyAsync(logits_out, m->d_logits, (int64_t)m->dbatch.B * 4, cudaMemcpyDeviceToHost, m->stream));
    mark(m, "d2h");
    WD_CUDA(cudaStreamSynchronize(m->stream));
    if (flags_host[0] & 1) {
        cudaMemsetAsync(m->d_flags, 0, 16, m->stream);
        set_error("categorical-column id capacity exceeded (max_nnz=%lld): recreate the model with a larger max_nnz", (long long)m->max_nnz);
        return WD_EINVAL;
    }
    if (flags_host[0] & 2) {
        cudaMemsetAsync(m->d_flags, 0, 16, m->stream);
        set_error("row-sharded exchange capacity exceeded (a rank received more than %lld ids): raise shard_slack / shard_capacity", (long long)m->max_nnz);
        return WD_EINVAL;
    }
    if (flags_host[0] & 8) {
        cudaMemsetAsync(m->d_flags, 0, 16, m->stream);
        set_error("data-parallel list exchange: this rank touched more unique rows than the fixed list length it exchanges (wd_sparse_set_sorted list_len); raise fixed_rows");
        return WD_EINVAL;
    }
    if (flags_host[0] & 4) {
        cudaMemsetAsync(m->d_flags, 0, 16, m->stream);
        set_error("row-sharded exchange: a peer rank did not reach a barrier within 20 s (ranks out of step, or a rank failed)");
        return WD_ESTATE;
    }
    if (loss_out) *loss_out = m->batch_has_label ? m->h_loss_pinned[0] : 0.f;
    if (m->timer.enabled) {
        PhaseTimer& t = m->timer;
        for (int i = 1; i < t.n; ++i) cudaEventElapsedTime(&t.ms[i], t.ev[i - 1], t.ev[i]);
        t.ms[0] = 0.f;
        if (t.n > 1) cudaEventElapsedTime(&t.ms[0], t.ev[0], t.ev[t.n - 1]);   // [0] = total
        t.n_last = t.n;
    }
    return WD_OK;
}

static bool list_present(const WdModel* m, int which) { return which == 0 ? (m->use_deep && !m->tables.empty()) : m->use_wide; }

// run `fn` on the side stream of sparse list `which` with that stream's scratch set
template <typename F>
static int on_side(WdModel* m, int which, F fn) {
    cudaStream_t main_stream = m->stream;
    m->stream = m->sstream[which]; m->scratch_sel = 1 + which;
    int rc = fn();
    m->stream = main_stream; m->scratch_sel = 0;
    return rc;
}

// launch the id-only grouping of both sparse lists on their side streams (overlaps forward + backward of the towers)
// WD_STEP_TRACE=1: stamp i of the step timeline on whatever stream is current (captured into the step's graph like any kernel)
__global__ void step_stamp_kernel(unsigned long long* t) { unsigned long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); *t = v; }
enum { ST_START = 0, ST_IDS, ST_GATHER, ST_HEAD, ST_BWD, ST_MAIN_END, ST_G0, ST_G1, ST_R0_BEG, ST_R0_END, ST_R1_BEG, ST_R1_END, ST_A0_END, ST_A1_END };
static void stamp(WdModel* m, int i) {
    if (!m->d_step_trace) return;
    step_stamp_kernel<<<1, 1, 0, m->stream>>>(m->d_step_trace + i);
    m->launches++;
}

static int group_async(WdModel* m) {
    if (m->timer.enabled) return WD_OK;                      // profiling: keep everything on one stream (done before the reduce)
    WD_CUDA(cudaEventRecord(m->ev_ids, m->stream));
    for (int w = 0; w < 2; ++w) {
        if (!list_present(m, w)) continue;
        WD_CUDA(cudaStreamWaitEvent(m->sstream[w], m->ev_ids, 0));
        int rc = on_side(m, w, [&] { int r = sparse_group_which(m, w); stamp(m, ST_G0 + w); return r; });
        if (rc) return rc;
        WD_CUDA(cudaEventRecord(m->ev_grouped[w], m->sstream[w]));
        m->side_pending[w] = true;
    }
    return WD_OK;
}

static int forward_core(WdModel* m, bool train) {
    int rc;
    stamp(m, ST_START);
    if ((rc = ids_prepare(m))) return rc;
    mark(m, "ids");
    stamp(m, ST_IDS);
    // training: the wide logit is needed only by the head, so it is computed on side stream 1, ahead of that stream's grouping work
    const bool wide_aside = train && !m->timer.enabled && m->use_wide && m->use_deep && list_present(m, 1);
    if (wide_aside) {
        WD_CUDA(cudaEventRecord(m->ev_ids, m->stream));
        WD_CUDA(cudaStreamWaitEvent(m->sstream[1], m->ev_ids, 0));
        if ((rc = on_side(m, 1, [&] { return sparse_forward_wide(m); }))) return rc;
        WD_CUDA(cudaEventRecord(m->ev_wide_fwd, m->sstream[1]));
    }
    if (train && (rc = group_async(m))) return rc;
    if (!wide_aside && (rc = sparse_forward_wide(m))) return rc;
    if ((rc = sparse_forward_emb(m))) return rc;
    stamp(m, ST_GATHER);
    if ((rc = mlp_forward(m, train))) return rc;
    mark(m, "mlp_other");
    if (wide_aside) WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_wide_fwd, 0));
    if ((rc = loss_forward(m, train))) return rc;
    mark(m, "head");
    stamp(m, ST_HEAD);
    if (train && (m->side_pending[0] || m->side_pending[1])) WD_CUDA(cudaEventRecord(m->ev_head, m->stream));
    return WD_OK;
}

// Backward.  Main stream: towers (dgrad before wgrad per layer), dense gradient reduction.  Side streams (when the
// grouping already lives there): wide gradient sums as soon as dlogit exists, embedding gradient sums as soon as dX0
// exists — i.e. under the remaining weight-gradient GEMMs.
static int backward_core(WdModel* m) {
    int rc;
    m->side_active[0] = m->side_active[1] = false;
    if (m->gs_count > 0) WD_CUDA(cudaMemsetAsync(m->d_G + m->dense_count, 0, (size_t)m->gs_count * sizeof(float), m->stream));
    if (m->side_pending[1]) {
        if (m->gs_count > 0) { WD_CUDA(cudaEventRecord(m->ev_head, m->stream)); }     // (re-recorded: also orders the memset above)
        WD_CUDA(cudaStreamWaitEvent(m->sstream[1], m->ev_head, 0));
        if ((rc = on_side(m, 1, [&] { stamp(m, ST_R1_BEG); int r = sparse_reduce_wide(m); if (!r) r = small_scatter(m, 1); stamp(m, ST_R1_END); return r; }))) return rc;
        m->side_active[1] = true;
    }
    m->record_dx0 = m->side_pending[0];
    m->dx0_recorded = false;
    // single-GPU fused step with the wide list on its side stream: the dense optimizer of everything but the first layer's kernel
    // runs there (after the wide rows' updates), under that kernel's weight-gradient GEMM
    m->dense_split_tensor = -1;
    m->record_wgrad_rest = m->fuse_dense && m->side_active[1] && m->gemm_engine == WD_GEMM_BF16X3 && !m->timer.enabled && !m->crelu;
    if ((rc = mlp_backward(m))) return rc;
    m->record_dx0 = false;
    m->record_wgrad_rest = false;
    mark(m, "mlp_other");
    stamp(m, ST_BWD);
    if (m->side_pending[0] && m->dx0_recorded) {
        WD_CUDA(cudaStreamWaitEvent(m->sstream[0], m->ev_dx0, 0));
        if ((rc = on_side(m, 0, [&] { stamp(m, ST_R0_BEG); int r = sparse_reduce_emb(m); if (!r) r = small_scatter(m, 0); stamp(m, ST_R0_END); return r; }))) return rc;
        m->side_active[0] = true;
    }
    if ((rc = wide_bias_grad(m))) return rc;
    if ((rc = dense_reduce_grads(m))) return rc;
    mark(m, "dense_reduce");
    for (int w = 0; w < 2; ++w) {                           // lists that stay on the main stream
        if (m->side_active[w] || !list_present(m, w)) { m->side_pending[w] = false; continue; }
        if (m->side_pending[w]) WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_grouped[w], 0));
        else if ((rc = sparse_group_which(m, w))) return rc;
        m->side_pending[w] = false;
        if ((rc = (w == 0 ? sparse_reduce_emb(m) : sparse_reduce_wide(m)))) return rc;
        if ((rc = small_scatter(m, w))) return rc;
    }
    if (m->gs_count > 0)                                    // the dense block is read (all-reduced) on the main stream
        for (int w = 0; w < 2; ++w)
            if (m->side_active[w]) {
                WD_CUDA(cudaEventRecord(m->ev_done[w], m->sstream[w]));
                WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_done[w], 0));
            }
    m->grads_pending = true;
    return WD_OK;
}

// Optimizer.  A list whose sums live on its side stream is applied there (after its merge in data-parallel runs); the dense
// optimizer runs on the main stream meanwhile and the streams join at the end of the step.
static int apply_core(WdModel* m) {
    int rc;
    const bool split_dense = m->fuse_dense && m->dense_split_tensor >= 0 && m->side_active[1];
    for (int w = 0; w < 2; ++w) {
        if (m->side_active[w]) {
            if ((rc = on_side(m, w, [&]() -> int {
                    int r = sparse_apply_which(m, w);
                    stamp(m, ST_A0_END + w);
                    if (!r && w == 1 && split_dense) {
                        WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_wgrad_rest, 0));
                        m->dense_part = 2;
                        r = dense_apply(m);
                        m->dense_part = 0;
                    }
                    return r;
                }))) return rc;
            WD_CUDA(cudaEventRecord(m->ev_done[w], m->sstream[w]));
        } else if ((rc = sparse_apply_which(m, w))) return rc;
    }
    mark(m, "sparse_apply");
    m->dense_part = split_dense ? 1 : 0;
    rc = dense_apply(m);
    m->dense_part = 0;
    if (rc) return rc;
    if ((rc = small_apply(m))) return rc;
    mark(m, "dense_apply");
    stamp(m, ST_MAIN_END);
    if (m->dropout_rate > 0.f && (rc = step_tick(m))) return rc;          // the dropout counter advances once per train step
    if (m->lin_opt.kind == WD_OPT_ADAM || m->dnn_opt.kind == WD_OPT_ADAM) {
        // AdamOptimizer._finish: beta powers advance once per step, after every variable of the optimizer has been updated (the
        // sparse lists may still be running on their side streams and read the powers: join them first)
        for (int w = 0; w < 2; ++w)
            if (m->side_active[w]) WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_done[w], 0));
        if ((rc = adam_tick(m))) return rc;
    }
    for (int w = 0; w < 2; ++w)
        if (m->side_active[w]) { WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_done[w], 0)); m->side_active[w] = false; }
    m->grads_pending = false;
    return WD_OK;
}

static int train_eager(WdModel* m) {
    int rc;
    // the whole step runs here, nothing exchanges the dense gradient arena between backward and optimizer: reduce the gradient
    // partials inside the optimizer kernel (mlp.cu dense_vec_kernel<2>)
    m->fuse_dense = true;
    rc = forward_core(m, true);
    if (!rc) rc = backward_core(m);
    if (!rc) rc = apply_core(m);
    m->fuse_dense = false;
    return rc;
}

static bool same_view(const DevBatch& a, const DevBatch& b) {
    return a.B == b.B && a.cat_offsets == b.cat_offsets && a.cat_keys == b.cat_keys && a.dense == b.dense && a.label == b.label && a.weight == b.weight;
}

// One whole train step on the current slot.  After two eager steps the step (three streams, ~55 kernels, no host sync)
// is captured once into a CUDA graph per batch slot and replayed: launch gaps between the many small kernels
// disappear and the host cost of a step becomes one cudaGraphLaunch.
static int train_current(WdModel* m, float* loss_out) {
    int rc;
    if (!m->batch_has_label) { set_error("training needs labels"); return WD_EINVAL; }
    BatchSlot& sl = m->slots[m->cur_slot];
    const bool can_graph = m->graphs_enabled && !m->timer.enabled;
    if (can_graph && sl.graph && same_view(sl.graph_view, m->dbatch)) {
        WD_CUDA(cudaGraphLaunch(sl.graph, m->stream));
        m->launches += sl.graph_launches;
        m->grads_pending = false;
    } else if (can_graph && sl.eager_steps >= 2) {
        if (sl.graph) { cudaGraphExecDestroy(sl.graph); sl.graph = nullptr; }
        const int64_t l0 = m->launches;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = train_eager(m);
            cudaError_t e2 = cudaStreamEndCapture(m->stream, &g);
            if (rc == WD_OK && e2 == cudaSuccess && g) e = cudaGraphInstantiate(&sl.graph, g, 0);
            else e = e2 != cudaSuccess ? e2 : cudaErrorUnknown;
            if (g) cudaGraphDestroy(g);
        }
        if (e != cudaSuccess || !sl.graph) {            // capture not possible here: stay eager for good
            cudaGetLastError();
            m->graphs_enabled = false;
            sl.graph = nullptr;
            m->side_pending[0] = m->side_pending[1] = m->side_active[0] = m->side_active[1] = false;
            if ((rc = train_eager(m))) return rc;
        } else {
            sl.graph_view = m->dbatch;
            sl.graph_launches = m->launches - l0;
            m->launches = l0;
            m->side_pending[0] = m->side_pending[1] = m->side_active[0] = m->side_active[1] = false; m->grads_pending = false;
            WD_CUDA(cudaGraphLaunch(sl.graph, m->stream));
            m->launches += sl.graph_launches;
        }
    } else {
        if ((rc = train_eager(m))) return rc;
        sl.eager_steps++;
    }
    if ((rc = mark_slot_used(m))) return rc;
    if (loss_out) return finish_step(m, loss_out, nullptr);
    return WD_OK;
}

extern "C" int wd_train_step_slot(WdModel* m, int slot, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, slot))) return rc;
    if (!m->slots[slot].filled) { set_error("batch slot %d was never uploaded", slot); return WD_ESTATE; }
    timer_begin(m);
    return train_current(m, loss_out);
}
extern "C" int wd_train_step_resident(WdModel* m, float* loss_out) { return wd_train_step_slot(m, 0, loss_out); }

extern "C" int wd_train_step(WdModel* m, const WdBatch* b, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, 0))) return rc;
    timer_begin(m);
    if ((rc = upload_current(m, b))) return rc;
    m->slots[0].view = m->dbatch; m->slots[0].has_label = m->batch_has_label; m->slots[0].filled = true;
    float dummy;
    return train_current(m, loss_out ? loss_out : &dummy);
}

extern "C" int wd_forward_resident(WdModel* m, float* logits_out, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, 0))) return rc;
    timer_begin(m);
    if ((rc = forward_core(m, false))) return rc;
    return finish_step(m, loss_out, logits_out);
}

extern "C" int wd_forward(WdModel* m, const WdBatch* b, float* logits_out, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, 0))) return rc;
    timer_begin(m);
    if ((rc = upload_current(m, b))) return rc;
    m->slots[0].view = m->dbatch; m->slots[0].has_label = m->batch_has_label; m->slots[0].filled = true;
    if ((rc = forward_core(m, false))) return rc;
    return finish_step(m, loss_out, logits_out);
}


// forward + backward of the current batch; `join` brings the side streams back into the main stream (needed to end a capture)
static int backward_eager(WdModel* m, bool join) {
    int rc;
    if ((rc = forward_core(m, true))) return rc;
    if ((rc = backward_core(m))) return rc;
    if (join)
        for (int w = 0; w < 2; ++w)
            if (m->side_active[w]) {
                WD_CUDA(cudaEventRecord(m->ev_done[w], m->sstream[w]));
                WD_CUDA(cudaStreamWaitEvent(m->stream, m->ev_done[w], 0));
            }
    return WD_OK;
}

// Data-parallel steps run forward + backward, then the exchange (NCCL, outside the library), then merge + apply.  The first part
// is ~65 launches on three streams: like the full step it is captured per batch slot after two eager runs and replayed.  Inside
// the graph the streams overlap as in the eager schedule; after it the sparse lists continue on their side streams, which wait for
// the graph through ev_bwd_done.
extern "C" int wd_step_backward_slot(WdModel* m, int slot, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if ((rc = select_slot(m, slot))) return rc;
    if (!m->slots[slot].filled) { set_error("batch slot %d was never uploaded", slot); return WD_ESTATE; }
    if (!m->batch_has_label) { set_error("training needs labels"); return WD_EINVAL; }
    timer_begin(m);
    BatchSlot& sl = m->slots[slot];
    const bool can_graph = m->graphs_enabled && !m->timer.enabled;
    auto after_graph = [&]() -> int {
        WD_CUDA(cudaGraphLaunch(sl.graph_bwd, m->stream));
        m->launches += sl.graph_bwd_launches;
        WD_CUDA(cudaEventRecord(m->ev_bwd_done, m->stream));
        for (int w = 0; w < 2; ++w) {
            m->side_pending[w] = false;
            m->side_active[w] = sl.bwd_side_active[w];
            if (m->side_active[w]) WD_CUDA(cudaStreamWaitEvent(m->sstream[w], m->ev_bwd_done, 0));
        }
        m->grads_pending = true;
        return WD_OK;
    };
    if (can_graph && sl.graph_bwd && same_view(sl.graph_bwd_view, m->dbatch)) {
        if ((rc = after_graph())) return rc;
    } else if (can_graph && sl.bwd_eager_steps >= 2) {
        if (sl.graph_bwd) { cudaGraphExecDestroy(sl.graph_bwd); sl.graph_bwd = nullptr; }
        const int64_t l0 = m->launches;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = backward_eager(m, true);
            cudaError_t e2 = cudaStreamEndCapture(m->stream, &g);
            if (rc == WD_OK && e2 == cudaSuccess && g) e = cudaGraphInstantiate(&sl.graph_bwd, g, 0);
            else e = e2 != cudaSuccess ? e2 : cudaErrorUnknown;
            if (g) cudaGraphDestroy(g);
        }
        if (e != cudaSuccess || !sl.graph_bwd) {             // capture not possible here: stay eager for good
            cudaGetLastError();
            m->graphs_enabled = false;
            sl.graph_bwd = nullptr;
            m->side_pending[0] = m->side_pending[1] = m->side_active[0] = m->side_active[1] = false;
            if ((rc = backward_eager(m, false))) return rc;
        } else {
            sl.graph_bwd_view = m->dbatch;
            sl.graph_bwd_launches = m->launches - l0;
            m->launches = l0;
            sl.bwd_side_active[0] = m->side_active[0]; sl.bwd_side_active[1] = m->side_active[1];
            if ((rc = after_graph())) return rc;
        }
    } else {
        if ((rc = backward_eager(m, false))) return rc;
        sl.bwd_eager_steps++;
    }
    if ((rc = mark_slot_used(m))) return rc;
    if (loss_out) return finish_step(m, loss_out, nullptr);
    return WD_OK;
}

// loss of the most recent forward / train step (synchronises the model stream)
extern "C" int wd_last_loss(WdModel* m, float* loss_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (!loss_out) { set_error("wd_last_loss: null output"); return WD_EINVAL; }
    return finish_step(m, loss_out, nullptr);
}

extern "C" int wd_step_backward(WdModel* m, const WdBatch* b, float* loss_out) {
    int rc = b ? wd_batch_upload(m, b) : check_ready(m);
    if (rc) return rc;
    timer_begin(m);
    if (!m->batch_has_label) { set_error("training needs labels"); return WD_EINVAL; }
    if ((rc = forward_core(m, true))) return rc;
    if ((rc = backward_core(m))) return rc;
    if (loss_out) return finish_step(m, loss_out, nullptr);
    return WD_OK;
}

extern "C" int wd_step_apply(WdModel* m) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (!m->grads_pending) { set_error("wd_step_apply without wd_step_backward"); return WD_ESTATE; }
    return apply_core(m);
}

extern "C" int64_t wd_dense_grad_count(WdModel* m) { return m ? m->dense_count + m->gs_count : 0; }
extern "C" void* wd_dense_grad_ptr(WdModel* m) { return m ? m->d_G : nullptr; }

extern "C" int wd_sparse_grads(WdModel* m, int which, void** rows, void** grads, int64_t* n, int32_t* width, int64_t* capacity) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (which < 0 || which > 1 || !m->d_urow[which]) { set_error("no sparse gradient list %d", which); return WD_EINVAL; }
    if (n) {                                   // the count needs a sync; pass n = NULL for the asynchronous fixed-size exchange
        int32_t nu = 0;
        WD_CUDA(cudaStreamSynchronize(m->stream));
        WD_CUDA(cudaStreamSynchronize(m->sstream[which]));
        WD_CUDA(cudaMemcpy(&nu, m->d_nuniq[which], 4, cudaMemcpyDeviceToHost));
        *n = nu;
    }
    if (rows) *rows = m->d_urow[which];
    if (grads) *grads = m->d_ugrad[which];
    if (width) *width = which == 0 ? m->emb_max_dim : 1;
    if (capacity) *capacity = m->sparse_cap[which];
    return WD_OK;
}

// n_lists = 0: general (unsorted) list of n rows; n_lists > 0: n_lists sorted, duplicate-free lists of n / n_lists rows each
static int sparse_set_impl(WdModel* m, int which, const void* rows_dev, const void* grads_dev, int64_t n, int n_lists) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (which < 0 || which > 1 || !m->d_urow[which]) { set_error("no sparse gradient list %d", which); return WD_EINVAL; }
    const bool side = m->side_active[which];
    auto merge = [&]() -> int {
        return n_lists > 0 ? merge_sparse_sorted(m, which, rows_dev, grads_dev, n_lists, n / n_lists) : merge_sparse(m, which, rows_dev, grads_dev, n);
    };
    auto run = [&]() -> int {
        if (side) return on_side(m, which, merge);
        return merge();
    };
    // ~25 small launches on one stream; with a fixed-size exchange the arguments never change, so after two eager runs the merge
    // is captured and replayed (the data-parallel tail is otherwise bound by the launching thread, not by the GPU)
    WdModel::MergeGraph& g = m->merge_graph[which];
    cudaStream_t st = side ? m->sstream[which] : m->stream;
    const bool same = g.rows == rows_dev && g.grads == grads_dev && g.n == n && g.on_side == side && g.n_lists == n_lists;
    if (!m->graphs_enabled || m->timer.enabled) return run();
    if (g.exec && same) {
        WD_CUDA(cudaGraphLaunch(g.exec, st));
        m->launches += g.launches;
        return WD_OK;
    }
    if (!same) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        g.rows = rows_dev; g.grads = grads_dev; g.n = n; g.on_side = side; g.n_lists = n_lists; g.eager = 0;
    }
    if (g.eager < 2) { g.eager++; return run(); }
    const int64_t l0 = m->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = run();
        cudaError_t e2 = cudaStreamEndCapture(st, &graph);
        if (rc == WD_OK && e2 == cudaSuccess && graph) e = cudaGraphInstantiate(&g.exec, graph, 0);
        else e = e2 != cudaSuccess ? e2 : cudaErrorUnknown;
        if (graph) cudaGraphDestroy(graph);
    }
    if (e != cudaSuccess || !g.exec) {                        // not capturable here: stay eager
        cudaGetLastError();
        g.exec = nullptr; g.eager = -1000000;
        return run();
    }
    g.launches = m->launches - l0;
    m->launches = l0;
    WD_CUDA(cudaGraphLaunch(g.exec, st));
    m->launches += g.launches;
    return WD_OK;
}

extern "C" int wd_sparse_set(WdModel* m, int which, const void* rows_dev, const void* grads_dev, int64_t n) {
    return sparse_set_impl(m, which, rows_dev, grads_dev, n, 0);
}
extern "C" int wd_sparse_set_sorted(WdModel* m, int which, const void* rows_dev, const void* grads_dev, int32_t n_lists, int64_t list_len) {
    if (n_lists < 1 || list_len < 1) { set_error("wd_sparse_set_sorted: bad list shape"); return WD_EINVAL; }
    return sparse_set_impl(m, which, rows_dev, grads_dev, (int64_t)n_lists * list_len, n_lists);
}

// ------------------------------------------------------------------------------------- row-sharded tables
namespace wd {
int shard_group_async(WdModel* m) { return ::group_async(m); }
int shard_backward_local(WdModel* m, bool) { return ::backward_core(m); }
int shard_apply_local(WdModel* m) { return ::apply_core(m); }
}

static int shard_ready(WdModel* m, int slot) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (m->shard.world <= 1) { set_error("model has no row-sharded tables (shard_world <= 1)"); return WD_ESTATE; }
    if (!m->shard.connected) { set_error("row-sharded model is not connected to its peers (wd_shard_connect_ipc / wd_shard_connect_local)"); return WD_ESTATE; }
    if ((rc = select_slot(m, slot))) return rc;
    if (!m->slots[slot].filled) { set_error("batch slot %d was never uploaded", slot); return WD_ESTATE; }
    return WD_OK;
}

// One phase of a sharded step (ranks driven by ONE process: the caller runs phase k on every rank, then wd_shard_local_sync).
extern "C" int wd_shard_phase(WdModel* m, int slot, int phase, int train) {
    int rc = shard_ready(m, slot);
    if (rc) return rc;
    if (m->shard.ipc) { set_error("wd_shard_phase is for ranks of one process; multi-process ranks call wd_shard_train_step_slot"); return WD_ESTATE; }
    if (train && !m->batch_has_label) { set_error("training needs labels"); return WD_EINVAL; }
    switch (phase) {
        case 0: return shard_phase0(m, train != 0);
        case 1: return shard_phase1(m, train != 0);
        case 2: rc = shard_phase2(m, train != 0); if (!train && rc == WD_OK) m->shard.step++; return rc;
        case 3: return train ? shard_phase3(m) : WD_OK;
        case 4: if (!train) return WD_OK; rc = shard_phase4(m); return rc ? rc : mark_slot_used(m);
    }
    set_error("wd_shard_phase: phase %d outside [0, 4]", phase);
    return WD_EINVAL;
}

// Loss (and optionally logits) of the step / forward just issued; synchronises the model stream.
extern "C" int wd_shard_finish(WdModel* m, float* loss_out, float* logits_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    return finish_step(m, loss_out, logits_out);
}

// The whole step of one rank of a multi-process job: ids, routing, serve, combine, towers, owners' updates, dense all-reduce and
// optimizers, with flag barriers in peer memory between the phases.  Every rank must call it once per step (it is a collective).
// After two eager steps per batch slot the step is captured into one CUDA graph (barrier kernels included) and replayed.
extern "C" int wd_shard_train_step_slot(WdModel* m, int slot, float* loss_out) {
    int rc = shard_ready(m, slot);
    if (rc) return rc;
    if (!m->shard.ipc) { set_error("wd_shard_train_step_slot needs wd_shard_connect_ipc (ranks of one process use wd_shard_phase)"); return WD_ESTATE; }
    if (!m->batch_has_label) { set_error("training needs labels"); return WD_EINVAL; }
    if (slot >= 64) { set_error("slot out of range"); return WD_EINVAL; }
    ShardState& S = m->shard;
    const bool can_graph = m->graphs_enabled && !m->timer.enabled;
    if (can_graph && S.graph[slot] && same_view(S.graph_view[slot], m->dbatch)) {
        WD_CUDA(cudaGraphLaunch(S.graph[slot], m->stream));
        m->launches += S.graph_launches[slot];
        S.step++;
    } else if (can_graph && S.eager_steps[slot] >= 2) {
        if (S.graph[slot]) { cudaGraphExecDestroy(S.graph[slot]); S.graph[slot] = nullptr; }
        const int64_t l0 = m->launches;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = shard_step_ipc(m, true);
            cudaError_t e2 = cudaStreamEndCapture(m->stream, &g);
            if (rc == WD_OK && e2 == cudaSuccess && g) e = cudaGraphInstantiate(&S.graph[slot], g, 0);
            else e = e2 != cudaSuccess ? e2 : cudaErrorUnknown;
            if (g) cudaGraphDestroy(g);
        }
        m->side_pending[0] = m->side_pending[1] = m->side_active[0] = m->side_active[1] = false; m->grads_pending = false;
        if (e != cudaSuccess || !S.graph[slot]) {              // not capturable here: stay eager for good
            cudaGetLastError();
            m->graphs_enabled = false;
            S.graph[slot] = nullptr;
            if ((rc = shard_step_ipc(m, true))) return rc;
        } else {
            S.graph_view[slot] = m->dbatch;
            S.graph_launches[slot] = m->launches - l0;
            m->launches = l0;
            WD_CUDA(cudaGraphLaunch(S.graph[slot], m->stream));
            m->launches += S.graph_launches[slot];
        }
    } else {
        if ((rc = shard_step_ipc(m, true))) return rc;
        S.eager_steps[slot]++;
    }
    if ((rc = mark_slot_used(m))) return rc;
    if (loss_out) return finish_step(m, loss_out, nullptr);
    return WD_OK;
}

// Forward only on a sharded model (collective, multi-process ranks): logits of this rank's batch shard.
extern "C" int wd_shard_forward_slot(WdModel* m, int slot, float* logits_out, float* loss_out) {
    int rc = shard_ready(m, slot);
    if (rc) return rc;
    if (!m->shard.ipc) { set_error("wd_shard_forward_slot needs wd_shard_connect_ipc"); return WD_ESTATE; }
    if ((rc = shard_step_ipc(m, false))) return rc;
    return finish_step(m, loss_out, logits_out);
}

// ------------------------------------------------------------------------------------------------- eval
extern "C" int wd_eval_reset(WdModel* m) {
    int rc = check_ready(m);
    if (rc) return rc;
    WD_CUDA(cudaMemsetAsync(m->d_metrics, 0, 512 * sizeof(double), m->stream));
    m->eval_batches = 0;
    return WD_OK;
}
extern "C" int wd_eval_accumulate(WdModel* m, const WdBatch* b) {
    int rc = wd_batch_upload(m, b);
    if (rc) return rc;
    if (!m->batch_has_label) { set_error("evaluation needs labels"); return WD_EINVAL; }
    if ((rc = forward_core(m, false))) return rc;
    if ((rc = metrics_accumulate(m))) return rc;
    return finish_step(m, nullptr, nullptr);
}
extern "C" int wd_eval_finish(WdModel* m, double* out10) {
    int rc = check_ready(m);
    if (rc) return rc;
    return metrics_finish(m, out10);
}

// ------------------------------------------------------------------------------------------ debug / misc
extern "C" int wd_debug_column_ids(WdModel* m, int32_t* offsets_out, int64_t offsets_cap, int64_t* ids_out, int64_t ids_cap, int64_t* nnz_out) {
    int rc = check_ready(m);
    if (rc) return rc;
    WD_CUDA(cudaStreamSynchronize(m->stream));
    int64_t no = (int64_t)m->dbatch.B * m->n_columns + 1;
    int32_t nnz = 0;
    WD_CUDA(cudaMemcpy(&nnz, m->d_nnz, 4, cudaMemcpyDeviceToHost));
    if (nnz_out) *nnz_out = nnz;
    if (offsets_out) {
        if (offsets_cap < no) { set_error("offsets buffer too small"); return WD_EINVAL; }
        WD_CUDA(cudaMemcpy(offsets_out, m->d_col_offs, no * 4, cudaMemcpyDeviceToHost));
    }
    if (ids_out) {
        if (ids_cap < nnz) { set_error("ids buffer too small"); return WD_EINVAL; }
        std::vector<int32_t> tmp(nnz);
        WD_CUDA(cudaMemcpy(tmp.data(), m->d_e_id, (int64_t)nnz * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nnz; ++i) ids_out[i] = tmp[i];
    }
    return WD_OK;
}
extern "C" int wd_debug_deep_input(WdModel* m, float* out, int64_t cap) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (!m->use_deep) { set_error("model has no deep part"); return WD_EINVAL; }
    int64_t n = (int64_t)m->dbatch.B * m->d0_phys;
    if (cap < n) { set_error("buffer too small"); return WD_EINVAL; }
    WD_CUDA(cudaStreamSynchronize(m->stream));
    WD_CUDA(cudaMemcpy(out, m->d_X0, n * 4, cudaMemcpyDeviceToHost));
    return WD_OK;
}
extern "C" int wd_debug_hidden(WdModel* m, int tower, int layer, float* out, int64_t cap) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (tower < 0 || tower >= (int)m->towers.size() || layer < 0 || layer >= m->towers[tower].n_hidden) { set_error("no such hidden layer"); return WD_EINVAL; }
    Layer& L = m->towers[tower].layers[layer];
    int64_t n = (int64_t)m->dbatch.B * L.N_phys;
    if (cap < n) { set_error("buffer too small"); return WD_EINVAL; }
    if (m->gemm_engine == WD_GEMM_BF16X3 && !L.h_fp32) {
        set_error("hidden layer %d is not materialised in fp32 by the bf16x3 engine (use gemm_engine ffma / tc3x to inspect it)", layer);
        return WD_EUNSUPPORTED;
    }
    WD_CUDA(cudaStreamSynchronize(m->stream));
    WD_CUDA(cudaMemcpy(out, L.H, n * 4, cudaMemcpyDeviceToHost));
    return L.N_phys;
}
extern "C" int64_t wd_launch_count(WdModel* m) { return m ? m->launches : 0; }
extern "C" int64_t wd_gemm_fallback_count(WdModel* m) { return m ? m->gemm_fallbacks : 0; }
extern "C" int wd_last_timings(WdModel* m, float* ms_out, int cap) {
    if (!m) return WD_EINVAL;
    int n = m->timer.n_last;
    for (int i = 0; i < n && i < cap; ++i) ms_out[i] = m->timer.ms[i];
    return n;
}
extern "C" const char* wd_timing_name(WdModel* m, int i) {
    if (!m || i < 0 || i >= m->timer.n_last) return "";
    return i == 0 ? "total" : m->timer.name[i];
}
extern "C" int wd_set_profile(WdModel* m, int enable) {
    if (!m) return WD_EINVAL;
    m->timer.enabled = enable != 0;
    return WD_OK;
}
extern "C" void* wd_stream(WdModel* m) { return m ? (void*)m->stream : nullptr; }
extern "C" void* wd_stream_sparse(WdModel* m, int which) {
    if (!m) return nullptr;
    return (which >= 0 && which < 2 && m->side_active[which]) ? (void*)m->sstream[which] : (void*)m->stream;
}
extern "C" int wd_sync(WdModel* m) {
    int rc = check_ready(m);
    if (rc) return rc;
    WD_CUDA(cudaStreamSynchronize(m->stream));
    return WD_OK;
}

// debugging aid (not part of the public header): stamps of the last train step's timeline (ns, globaltimer), WD_STEP_TRACE=1:
// [start, ids, gather, head, towers' backward, main stream end, grouping 0 / 1 end, reduce 0 begin / end, reduce 1 begin / end,
// apply 0 / 1 end] (0 = embedding rows, 1 = wide rows; side streams)
extern "C" int wd_debug_step_trace(WdModel* m, unsigned long long* out) {
    if (!m || !m->d_step_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, m->d_step_trace, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
