// Shared declarations of libwd_b200: host-side model object, device plan tables, launch helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/wd_b200.h"

namespace wd {

constexpr int kSortTile = 1024;          // keys per radix-sort tile (sort.cu); sizes the per-tile histogram scratch

void set_error(const char* fmt, ...);

#define WD_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            wd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return WD_ECUDA;                                                                  \
        }                                                                                     \
    } while (0)

constexpr uint32_t kInvalidRow = 0xFFFFFFFFu;
constexpr int kMaxSegs = 8;      // sources concatenated into one MLP layer input
constexpr int kMaxDims = 8;      // distinct embedding widths per model
// sort / unique-row scratch sets ("lists"): 0 embedding rows and 1 wide rows of the replicated tables; row-sharded tables add
// 2 / 3 = rows this rank OWNS that were touched by any rank this step (embedding / wide) and 4 / 5 = this rank's own ids
// grouped by owner rank (routing; only the key / value ping-pong buffers exist)
constexpr int kLists = 6;
constexpr int kMaxRanks = 16;    // ranks of one box a row-sharded table can be split over

// ---- device image of the categorical-column plan (all pointers are device pointers)
struct DevPlan {
    int n_cat_fields, n_dense_fields, n_columns;
    const uint8_t* field_is_string;
    const int32_t* col_kind;
    const int32_t* col_field;
    const int64_t* col_buckets;
    const int32_t* col_aux_off;
    const int32_t* col_aux_n;
    const int32_t* col_norm_kind;
    const float* col_norm_a;
    const float* col_norm_b;
    const int64_t* col_wide_base;
    const int32_t* col_emb_table;
    const int32_t* col_ind_off;
    const uint64_t* vocab_fp;
    const float* boundaries;
    const int32_t* cross_key_type;
    const int32_t* cross_key_idx;
    const int64_t* table_row_base;   // global embedding row id of row 0 of each table
    int d0_phys;
    // row-sharded tables (shard_world > 1): per column the slot in the sharded embedding / wide space (-1: replicated), the first
    // local row of each slot's shard, and per entry the owner rank / local row arrays the id stage fills (null: nothing sharded)
    int sh_world;
    const int32_t *sh_col_emb, *sh_col_wide;
    const int64_t *sh_base_emb, *sh_base_wide;
    uint32_t *sh_own_emb, *sh_lrow_emb, *sh_own_wide, *sh_lrow_wide;
};

// ---- device view of one batch
struct DevBatch {
    int B;
    const int32_t* cat_offsets;      // may be null: one key per (row, field)
    const uint64_t* cat_keys;
    const float* dense;
    const float* label;              // may be null
    const float* weight;             // may be null
};

struct EmbTable {
    int64_t rows;
    int dim;                 // physical width (multiple of 4)
    int dim_logical;         // embedding_column dimension; columns [dim_logical, dim) are zero padding
    int x0_off;
    int col;                 // producing column
    int64_t row_base;        // global row id base
    float* data;             // arows * stride floats; record = [w[dim] | slot1[dim] | slot2[dim]]
    int stride;              // dim * (1 + nslots)
    bool sharded;            // row-sharded over the ranks: this rank holds rows r with r mod G == rank at local index r / G
    int64_t arows;           // rows allocated on this rank (= rows unless sharded); row_base then counts in the shard's row space
};

struct TabDesc {            // per-table descriptor grouped by width (one 32-byte load in the gather kernels)
    float* data;
    int64_t row_base;
    int stride, x0, col, dim;
};

struct DenseTensor {         // one trainable dense tensor inside the dense arena
    int64_t off;             // offset in arena (floats)
    int64_t count;           // physical element count
    int rows, cols;          // physical shape (rows = K_phys for kernels, 1 for vectors)
    // gradient partials: grad[i] = sum_p gpart[p * gstride + i]
    int64_t gpart_off;       // offset in the partial buffer
    int gparts;
    int g_rowtiles;          // 1: partials are per 128-row batch tile (only the tiles of the current batch are live)
    int64_t gstride;
    int64_t wt_off;          // offset of the transposed copy in the wt buffer, -1 if none
    int mirror_u;            // crelu layers (kernel, bias): column n + mirror_u holds minus column n, n < mirror_u (0: plain tensor)
};

struct Seg {                 // one source of a layer input
    int src;                 // -1: deep input x, >=0: hidden layer index of the same tower
    int width;               // logical width
    int width_phys;          // padded to 32
    int k_off;               // physical row offset inside the layer's kernel
};

struct Layer {
    int n_in_segs;
    Seg segs[kMaxSegs];
    int K, K_phys;           // logical / physical input width
    int N, N_phys;           // logical / physical output width (1 for logits)
    int N_param;             // logical columns of kernel / bias: N, or N / 2 for a crelu layer (the other half mirrors them)
    int t_kernel, t_bias, t_gamma, t_beta;   // indices into Model::dense (-1 if absent)
    int wgrad_splits;        // split-K factor of this layer's weight gradient (= gparts of its kernel tensor)
    // activations (hidden layers only), all [max_batch_pad, N_phys] unless noted
    float* A;                // post-activation (pre-BN); == H when no batch norm
    float* H;                // layer output
    float* HT;               // transposed output [N_phys, ldt]
    float* dH;               // gradient w.r.t. H (accumulated over consumers)
    float* dZ;               // gradient w.r.t. pre-activation
    float* dZT;              // transposed [N_phys, ldt]
    // 3xBF16 engine: bf16 hi / lo copies of H and dZ (what the tensor-core GEMMs read; no transposed copies)
    __nv_bfloat16 *Hs[2], *dZs[2];
    bool h_fp32;             // 3xBF16 engine: H is also stored in fp32 (only layers the logits layer reads)
    float* colpart;          // [3][row_tiles][N_phys] partial column sums: dbias, dgamma, dbeta
};

struct Tower {
    int n_hidden;
    int mode;
    std::vector<Layer> layers;   // n_hidden hidden layers + 1 logits layer
    float* logit;                // [max_batch]
};

struct PhaseTimer {          // named CUDA-event marks on the model stream (wd_set_profile)
    static constexpr int kMax = 160;
    cudaEvent_t ev[kMax];
    const char* name[kMax];
    float ms[kMax];
    int n = 0, n_last = 0;
    bool enabled = false;
};

struct BatchSlot {           // one device-resident batch (ring used by benchmarks / prefetch)
    int32_t* off = nullptr;
    uint64_t* keys = nullptr;
    float *dense = nullptr, *label = nullptr, *weight = nullptr;
    DevBatch view{};
    bool has_label = false, filled = false;
    // asynchronous refill (wd_batch_prefetch_slot): copies run on the upload stream between these two events
    cudaEvent_t ev_up = nullptr;             // recorded on the upload stream after the slot's copies
    cudaEvent_t ev_used = nullptr;           // recorded on the model stream after the last step that read the slot
    bool up_pending = false, used_recorded = false;
    // CUDA graph of one whole train step on this slot (captured after a few eager steps; keyed by the batch view)
    cudaGraphExec_t graph = nullptr;
    DevBatch graph_view{};
    int64_t graph_launches = 0;
    int eager_steps = 0;
    // CUDA graph of forward + backward only (data-parallel steps: wd_step_backward_slot), side streams joined at its end
    cudaGraphExec_t graph_bwd = nullptr;
    DevBatch graph_bwd_view{};
    int64_t graph_bwd_launches = 0;
    int bwd_eager_steps = 0;
    bool bwd_side_active[2] = {false, false};
};


// ---- row-sharded tables (shard.cu) ------------------------------------------------------------------------------------
// Pointers into ONE rank's exchange segment (a single cudaMalloc block that peers map through CUDA IPC, or address directly when
// all ranks live in one process).  Every rank lays its segment out identically, so a peer pointer = peer base + own offset.
struct ShardPeer {
    uint2* inbox[2];          // [G][pair_cap] entries {local row, bag} written by requester ranks (double buffered by step parity)
    int32_t* inbox_cnt[2];    // [G] entries each requester sent this step
    float* recv;              // [G][nbags][width] pooled partial sums written by owner ranks
    const float* bagscale;    // [nbags] 1 / (ids in the bag)   (embedding space: combiner = mean)
    const float* gradbase;    // dX0 (embedding space) / dlogit (wide space) of that rank: owners pull gradients from here
};
struct ShardSpace {           // one sharded table space on this rank: 0 = embedding tables, 1 = wide columns
    bool on = false;
    int width = 0;            // floats per pooled vector: max width of the sharded embedding tables / 1
    int n_slots = 0;          // sharded tables (embedding) / sharded wide columns
    int bags_per_row = 0;     // embedding: n_slots (bag = example * n_slots + slot); wide: 1 (bag = example)
    int64_t nbags_cap = 0;    // max_batch * bags_per_row
    int64_t local_rows = 0;   // rows of this rank's shard of the space
    int64_t pair_cap = 0;     // entries one rank may send to one owner per step
    std::vector<int32_t> h_col_slot;   // host copies (tensor IO)
    std::vector<int64_t> h_slot_base;
    int32_t* d_col_slot = nullptr;     // [n_columns] slot fed by column c, -1
    int64_t* d_slot_base = nullptr;    // [n_slots] first local row of the slot's shard
    int32_t *d_slot_dim = nullptr, *d_slot_x0 = nullptr, *d_slot_stride = nullptr;
    float** d_slot_data = nullptr;     // [n_slots] shard of the table (embedding space)
    float4* d_wide = nullptr;          // wide space: {w, n, z, -} per local row
    uint32_t* d_own = nullptr;         // [max_nnz] owner rank of entry j or kInvalidRow (not a sharded column)
    uint32_t* d_lrow = nullptr;        // [max_nnz] local row at the owner
    int32_t* d_ostart = nullptr;       // [G + 1] start of each owner's run in the routed (owner-sorted) list
    int32_t* d_bagmask = nullptr;      // [nbags] bit o: owner o holds a partial sum of the bag
    uint32_t* d_rtag = nullptr;        // owner side, per received entry: (source rank << 27) | bag
    uint32_t* d_rrow = nullptr;        // owner side, per received entry: local row
    int32_t* d_nrecv = nullptr;        // device scalar
    ShardPeer* d_peers = nullptr;      // [G] device copy
    ShardPeer peers[kMaxRanks];        // host copy (pointers are device addresses)
    int64_t off_inbox[2] = {0, 0}, off_cnt[2] = {0, 0}, off_recv = 0, off_bagscale = 0, off_grad = 0;   // offsets in the segment
};
struct ShardState {
    int world = 1, rank = 0;
    bool connected = false, ipc = false;
    uint8_t* seg = nullptr;            // this rank's exchange segment
    int64_t seg_bytes = 0;
    uint8_t* peer_seg[kMaxRanks] = {};
    ShardSpace sp[2];
    // flag barriers: flags[k][r] = last epoch rank r signalled on barrier k (written by rank r, through peer memory)
    int64_t off_flags = 0, off_gred = 0, off_G = 0;
    uint32_t** d_peer_flags = nullptr;  // [G] device array of peers' flag blocks
    uint32_t* d_epoch = nullptr;        // [kBarriers] local epoch counters
    unsigned long long* d_trace = nullptr;   // WD_SHARD_TRACE=1: [kBarriers][2] enter / leave stamps of the last step's barriers
    float** d_peer_G = nullptr;         // [G] peers' dense gradient arenas
    float** d_peer_gred = nullptr;      // [G] peers' reduced slices
    float* gred = nullptr;              // this rank's reduced slice buffer (whole-arena sized; only the own slice is written)
    int64_t ar_count = 0;               // floats all-reduced per step (dense gradients + small-table block)
    uint64_t step = 0;                  // steps issued (inbox double buffering)
    cudaEvent_t ev_a = nullptr;         // after barrier A on the main stream (owner-side grouping may start)
    cudaStream_t aux = nullptr;         // the wide space's routing / serving and the local gathers, beside the embedding space's chain
    cudaEvent_t ev_ids2 = nullptr, ev_routed1 = nullptr, ev_a2 = nullptr, ev_aux_done = nullptr;
    cudaGraphExec_t graph[64] = {};     // whole sharded step per batch slot
    DevBatch graph_view[64];
    int64_t graph_launches[64] = {};
    int eager_steps[64] = {};
};
constexpr int kBarriers = 8;

}  // namespace wd

struct WdModel {
    // ---- host copy of the plan
    int device = 0;
    bool use_wide = false, use_deep = false;
    int n_cat_fields = 0, n_dense_fields = 0, n_columns = 0;
    std::vector<int32_t> col_kind, col_field, col_aux_off, col_aux_n, col_norm_kind, col_emb_table, col_ind_off;
    std::vector<int64_t> col_buckets, col_wide_base;
    std::vector<float> col_norm_a, col_norm_b;
    int n_numeric = 0;
    int d0_phys = 0;
    int64_t wide_rows = 0;
    int activation = 0, batch_norm = 0;
    WdOptimizer lin_opt{}, dnn_opt{};
    int max_batch = 0, max_batch_pad = 0;   // pad: multiple of 128 (leading dim of transposed buffers)
    int64_t max_nnz = 0;
    int gemm_engine = 0;

    cudaStream_t stream = nullptr;
    // side streams, one per sparse gradient list (0 = embedding rows, 1 = wide rows): the id-only grouping, the gradient sums,
    // the data-parallel merge and the row updates of a list all run there, overlapping the towers on the main stream
    cudaStream_t sstream[2] = {nullptr, nullptr};
    cudaStream_t stream_up = nullptr;        // host->device refills of batch slots (wd_batch_prefetch_slot), overlapping the running step
    cudaEvent_t ev_ids = nullptr, ev_head = nullptr, ev_dx0 = nullptr, ev_bwd_done = nullptr, ev_wide_fwd = nullptr, ev_wgrad_rest = nullptr;
    bool sort_smem_opt_in = false;               // rs_scatter_kernel<RS_BIG_TILE, true> was granted its dynamic shared memory on this device
    bool crelu = false;                          // dnn_activation_function crelu: relu on mirrored kernels (mlp.cu crelu_fold / crelu_mirror)
    bool list_apply_fused[2] = {false, false};   // this step's rows of the list were updated by its gradient-sum / combine launches (sparse.cu)
    bool record_wgrad_rest = false;           // mlp_backward: record ev_wgrad_rest before the first layer's weight gradient
    int dense_split_tensor = -1, dense_part = 0;   // dense_apply: see mlp.cu (single-GPU step, split dense optimizer)
    cudaEvent_t ev_grouped[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    bool side_pending[2] = {false, false};   // the list's grouping of this step was issued on its side stream
    bool side_active[2] = {false, false};    // the list's sums live on its side stream (merge / apply follow there)
    bool record_dx0 = false, dx0_recorded = false;
    // CUDA graph of the data-parallel merge of list w (wd_sparse_set with the same buffers every step: fixed-size exchange)
    struct MergeGraph { cudaGraphExec_t exec = nullptr; const void* rows = nullptr; const void* grads = nullptr; int64_t n = 0; int n_lists = 0;
                        bool on_side = false; int eager = 0; int64_t launches = 0; } merge_graph[2];
    // ---- dense exchange of small tables (WdPlanDesc::dense_exchange_max_rows)
    int64_t dense_exchange_max_rows = 0;
    int64_t small_base[2] = {0, 0};          // first global row of the small embedding tables / small wide columns
    int64_t gs_count = 0, gs_emb_floats = 0; // floats of the small-table gradient block behind d_G[dense_count]; its embedding part
    int64_t gs_touch_off[2] = {0, 0};        // offsets (floats, inside the block) of the per-row "touched" counts: small embedding rows / small wide rows
    int n_rtab = 0, n_small_tab = 0;         // tables in row order (large first, then small); how many of them are small
    int64_t* d_rtab_row_base = nullptr;      // [n_rtab] row base, ascending
    float** d_rtab_data = nullptr;
    int32_t *d_rtab_dim = nullptr, *d_rtab_stride = nullptr;
    int64_t* d_rtab_gs_off = nullptr;        // [n_rtab] float offset inside the block (-1: large table)
    int32_t* d_nubig[2] = {nullptr, nullptr};   // unique rows below small_base (what stays in the list)
    wd::ShardState shard;                    // row-sharded tables (world == 1: unused)
    wd::DevPlan dplan{};
    std::vector<void*> allocs;               // everything cudaMalloc'ed (freed in destroy)
    int64_t bytes_allocated = 0;

    // numeric deep columns (device arrays)
    int32_t *d_num_field = nullptr, *d_num_norm_kind = nullptr, *d_num_x0_off = nullptr;
    float *d_num_a = nullptr, *d_num_b = nullptr;

    // ---- batch buffers
    int32_t* d_cat_offsets = nullptr;
    uint64_t* d_cat_keys = nullptr;
    int64_t keys_cap = 0;
    float *d_dense = nullptr, *d_label = nullptr, *d_weight = nullptr;
    wd::DevBatch dbatch{};
    bool batch_has_label = false;

    // ---- column ids (CSR over (row, column)) and per-entry arrays
    int32_t* d_col_offs = nullptr;           // [max_batch * n_columns + 1]
    uint32_t* d_e_wide = nullptr;            // [max_nnz] global wide row or kInvalidRow
    uint32_t* d_e_emb = nullptr;             // [max_nnz] global embedding row or kInvalidRow
    int32_t* d_e_bc = nullptr;               // [max_nnz] row * n_columns + column
    int32_t* d_e_id = nullptr;               // [max_nnz] column-local id (debug / parity)
    int32_t* d_nnz = nullptr;                // device scalar: entries this step
    int32_t* d_flags = nullptr;              // device error flags [4]
    void* d_scan_tmp_s[4] = {nullptr, nullptr, nullptr, nullptr};   // scan / sort scratch, one set per stream (0 main, 1 + list for the side streams, 3 aux)
    int scratch_sel = 0;

    // ---- wide part: record {w, n, z, 0} per row
    float4* d_wide = nullptr;
    float* d_wide_logit = nullptr;           // [max_batch]

    // ---- deep part
    std::vector<wd::EmbTable> tables;
    int64_t emb_total_rows = 0;
    int emb_max_dim = 0;
    int n_dims = 0;
    int dims[wd::kMaxDims];                  // distinct widths
    int32_t* d_dim_tables[wd::kMaxDims];     // table ids per width (device)
    wd::TabDesc* d_dim_desc[wd::kMaxDims];    // descriptors of the same tables
    int dim_ntables[wd::kMaxDims];
    // device table descriptors
    float** d_tab_data = nullptr;
    int32_t *d_tab_dim = nullptr, *d_tab_stride = nullptr, *d_tab_x0 = nullptr, *d_tab_col = nullptr;
    int64_t* d_tab_row_base = nullptr;
    float *d_X0 = nullptr, *d_X0T = nullptr, *d_dX0 = nullptr;
    __nv_bfloat16* d_X0s[2] = {nullptr, nullptr};   // bf16 hi / lo copy of X0 (3xBF16 engine)
    int ldt = 0;                             // leading dim of transposed activations (= max_batch_pad)
    std::vector<wd::Tower> towers;

    // ---- dense parameter arena
    std::vector<wd::DenseTensor> dense;
    int64_t dense_count = 0, gpart_count = 0, wt_count = 0;
    float *d_P = nullptr, *d_S1 = nullptr, *d_S2 = nullptr, *d_G = nullptr, *d_gpart = nullptr, *d_Wt = nullptr;
    float* d_Wsplit = nullptr;               // [W_hi | W_lo | Wt_hi | Wt_lo], each wt_count floats (3xTF32 pre-split weights)
    wd::DenseTensor* d_dense_desc = nullptr;
    int row_tiles = 0;                       // max_batch_pad / 128
    int wgrad_splits = 4;

    // ---- loss / logits
    float* d_logits = nullptr;               // [max_batch]
    float* d_dlogit = nullptr;               // [max_batch]
    float* d_loss_part = nullptr;            // per-block partials
    float* d_loss = nullptr;                 // scalar
    unsigned long long* d_step_trace = nullptr;   // WD_STEP_TRACE=1: globaltimer stamps of the last step (wd_debug_step_trace)
    int32_t* d_head_counter = nullptr;       // blocks of the fused head kernel that have finished (last one sums the loss)
    float* d_bpow = nullptr;                 // Adam: {linear beta1^t, linear beta2^t, dnn beta1^t, dnn beta2^t}, multiplied in fp32 after every step (AdamOptimizer._finish)
    float* h_loss_pinned = nullptr;

    // ---- sparse backward scratch (two sorts: 0 = embedding rows, 1 = wide rows)
    uint32_t *d_sk[wd::kLists] = {}, *d_sv[wd::kLists] = {};     // ping-pong keys / values
    uint32_t *d_sk2[wd::kLists] = {}, *d_sv2[wd::kLists] = {};
    int32_t* d_sort_hist_s[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t sort_hist_cap = 0;
    int32_t* d_sort_counter_s[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* d_urow[wd::kLists] = {};   // unique rows
    int32_t* d_ustart[wd::kLists] = {};  // segment starts in the sorted list (+1 sentinel)
    float* d_ugrad[wd::kLists] = {};     // [cap, width]
    int32_t* d_nuniq[wd::kLists] = {};   // device scalars
    int32_t* d_nvalid[wd::kLists] = {};
    int32_t* d_choff[wd::kLists] = {};   // chunk offsets of hot rows (exclusive scan)
    int32_t* d_nchunks[wd::kLists] = {};
    float* d_cpart[wd::kLists] = {};     // chunk partial sums
    int64_t cpart_cap = 0;
    int64_t sparse_cap[2] = {0, 0};
    bool sparse_overridden[2] = {false, false};
    int64_t sparse_override_n[2] = {0, 0};
    int sort_bits[wd::kLists] = {};

    // ---- eval metrics
    double* d_metrics = nullptr;             // accumulators
    int64_t eval_batches = 0;

    int64_t launches = 0;
    float dropout_rate = 0.f;               // dnn_dropout: tf.layers.dropout after every hidden layer's activation, train steps only
    unsigned long long dropout_seed = 0;
    unsigned int* d_step = nullptr;         // device: train steps completed (dropout counter; advances at the end of every step)
    bool fuse_dense = false;                // whole local step (train_eager): dense gradient reduction fused into the optimizer kernel
    int64_t gemm_fallbacks = 0;            // tensor-core engine GEMMs that ran on the FFMA kernel instead (tests assert 0)
    int cur_slot = 0;
    bool graphs_enabled = true;              // WD_NO_GRAPH=1 disables step graphs
    int cur_layer = 0;                       // layer being launched (names the profiling marks)
    wd::PhaseTimer timer;
    std::vector<wd::BatchSlot> slots;        // slot 0 aliases the d_cat_* buffers above
    bool initialized = false;
    bool grads_pending = false;
};

namespace wd {
// ---- kernels / stages implemented in the .cu files (all enqueue on m->stream)
int ids_prepare(WdModel* m);                                     // ids.cu
int sparse_forward(WdModel* m);                                  // sparse.cu: wide logit + embedding pooling + numerics
int sparse_forward_wide(WdModel* m);                             //   the wide half alone
int sparse_forward_emb(WdModel* m);                              //   the deep half alone
int sparse_group(WdModel* m);                                    // sparse.cu: sort (row, occurrence) pairs, unique rows, chunks
int sparse_reduce_emb(WdModel* m);                               // sparse.cu: per-row gradient sums (needs dX0)
int sparse_reduce_wide(WdModel* m);                              // sparse.cu: per-row gradient sums (needs dlogit only)
int sparse_apply(WdModel* m);                                    // sparse.cu: Adagrad / FTRL / SGD on touched rows (both lists)
int sparse_apply_which(WdModel* m, int which);                   // sparse.cu: one list (0 = embedding rows, 1 = wide rows)
int sparse_group_which(WdModel* m, int which);                   // sparse.cu: grouping of one list
int merge_sparse_sorted(WdModel* m, int which, const void* rows, const void* grads, int n_lists, int64_t list_len);   // sparse.cu
int small_scatter(WdModel* m, int which);                        // sparse.cu: small-table rows of list `which` -> dense block
int small_apply(WdModel* m);                                     // sparse.cu: optimizer over the dense block (after its all-reduce)
int mlp_forward(WdModel* m, bool want_transposes);               // mlp.cu: towers -> logits, loss
int mlp_backward(WdModel* m);                                    // mlp.cu: grads of dense params, dX0
int dense_reduce_grads(WdModel* m);                              // mlp.cu
int dense_apply(WdModel* m);                                     // mlp.cu
int loss_forward(WdModel* m, bool need_grad);                    // mlp.cu: logits = wide + deep, loss, dlogit
int model_init_params(WdModel* m, uint64_t seed);                // init.cu
int step_tick(WdModel* m);                                       // misc.cu: train-step counter on the device (dropout)
int adam_tick(WdModel* m);                                       // misc.cu: beta powers advance (after every optimizer of the step)
int metrics_accumulate(WdModel* m);                              // metrics.cu
int metrics_finish(WdModel* m, double* out10);

int radix_sort_pairs(WdModel* m, int which, int bits, const int32_t* d_n);   // sort.cu
int exclusive_scan_i32(WdModel* m, int32_t* data, int64_t n, int32_t* total_out);   // sort.cu (in place, n known on host)
int seg_heads(WdModel* m, const int32_t* d_n, const uint32_t* keys, uint32_t invalid, int32_t* pos, int64_t cap, int32_t* ustart, uint32_t* urow,
              int32_t* d_nuniq);                                                    // sort.cu: unique rows of sorted keys (2 launches)
int chunk_offsets(WdModel* m, const int32_t* d_nuniq, const int32_t* ustart, uint32_t* urow, int32_t* choff, int64_t cap, int chunk,
                  int32_t* d_nchunks);                                              // sort.cu: hot-row chunk layout (2 launches)

template <typename T>
int dev_alloc(WdModel* m, T** p, int64_t count, bool zero = true) {
    if (count <= 0) count = 1;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (size_t)count * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %lld bytes failed: %s", (long long)(count * sizeof(T)), cudaGetErrorString(e));
        return WD_ENOMEM;
    }
    if (zero) {
        e = cudaMemsetAsync(q, 0, (size_t)count * sizeof(T), m->stream);
        if (e != cudaSuccess) { set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return WD_ECUDA; }
    }
    m->allocs.push_back(q);
    m->bytes_allocated += count * (int64_t)sizeof(T);
    *p = (T*)q;
    return WD_OK;
}

// record a named mark on the model stream (no-op unless profiling is on)
inline void mark(WdModel* m, const char* name) {
    PhaseTimer& t = m->timer;
    if (!t.enabled || t.n >= PhaseTimer::kMax) return;
    t.name[t.n] = name;
    cudaEventRecord(t.ev[t.n], m->stream);
    t.n++;
}

inline int grid_for(int64_t n, int block, int cap = 148 * 16) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
}  // namespace wd
