// TSV -> CSR batch loader (host, multi-threaded).
// Replaces _CsvDataset._parse_csv (reference python/lib/dataset.py:107-165): fields split on TAB
// only, no quoting; an empty field or the NA token "-" takes its default ('' / 0 / 0.0); multi-valued
// string fields split on ',' with empty tokens dropped; strings leave the loader as Fingerprint64 values
// (the same function the GPU stage uses), so no string ever crosses PCIe.  Rows are parsed in contiguous blocks, one block per
// thread, without per-row heap allocations; keys are assembled with a parallel copy.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "farmhash.cuh"

namespace {
struct LineRef { const char* p; int len; };

// Per-thread parse state: nothing is allocated per row.  Tokens of a row are collected in file order as (field, key) pairs and
// written field-major with a counting sort (fields may appear in any file column order).
struct ThreadOut {
    std::vector<uint64_t> keys;         // keys of this thread's rows, row after row, field-major inside a row
    std::vector<uint64_t> tok_key;      // scratch: tokens of the current row
    std::vector<int32_t> tok_field;
    std::vector<int32_t> pos;           // scratch: per-field write cursor
    std::string err;
};

inline bool parse_int_field(const char* f, int flen, long long* out) {
    char buf[48];
    if (flen >= (int)sizeof(buf)) return false;
    memcpy(buf, f, flen); buf[flen] = 0;
    char* ep = nullptr;
    *out = strtoll(buf, &ep, 10);
    return *ep == 0;
}
inline bool parse_float_field(const char* f, int flen, float* out) {
    char buf[64];
    if (flen >= (int)sizeof(buf)) return false;
    memcpy(buf, f, flen); buf[flen] = 0;
    char* ep = nullptr;
    *out = strtof(buf, &ep);
    return *ep == 0;
}

// one record: appends its keys to st.keys (field-major), writes counts[F], dense[Nd], label, weight
bool parse_row(const WdTsvSpec* sp, const char* p, int len, ThreadOut& st, int32_t* counts, float* dense, float* label, float* weight, char* err) {
    const int F = sp->n_cat_fields;
    st.tok_key.clear(); st.tok_field.clear();
    for (int i = 0; i < F; ++i) counts[i] = 0;
    int col = 0;
    const char* end = p + len;
    const char* f = p;
    for (int i = 0; i < sp->n_dense_fields; ++i) dense[i] = 0.f;
    float lab = 0.f;
    while (true) {
        const char* q = (const char*)memchr(f, '\t', end - f);
        const char* fe = q ? q : end;
        if (col >= sp->n_columns) { snprintf(err, 256, "Expect %d fields but have more in record", sp->n_columns); return false; }
        const int role = sp->col_role[col], tgt = sp->col_target[col];
        const int flen = (int)(fe - f);
        const bool na = flen == 0 || (flen == 1 && f[0] == '-');
        if (role == 0) {
            long long v = 0;
            lab = (!na && ((flen == 1 && f[0] == '1') || (parse_int_field(f, flen, &v) && v == 1))) ? 1.f : 0.f;
        } else if (role == 1) {
            if (!na) {
                if (sp->multivalue) {
                    const char* t = f;
                    while (t <= fe) {
                        const char* c = (const char*)memchr(t, ',', fe - t);
                        const char* te = c ? c : fe;
                        if (te > t) { st.tok_key.push_back(wd::fingerprint64((const uint8_t*)t, te - t)); st.tok_field.push_back(tgt); counts[tgt]++; }
                        if (!c) break;
                        t = c + 1;
                    }
                } else {
                    st.tok_key.push_back(wd::fingerprint64((const uint8_t*)f, flen)); st.tok_field.push_back(tgt); counts[tgt]++;
                }
            }
        } else if (role == 2) {
            long long v = 0;
            if (!na && !parse_int_field(f, flen, &v)) { snprintf(err, 256, "Field %d in record is not a valid int32: %.*s", col, flen < 40 ? flen : 40, f); return false; }
            st.tok_key.push_back((uint64_t)v); st.tok_field.push_back(tgt); counts[tgt]++;
        } else if (role == 3) {
            float v = 0.f;
            if (!na && !parse_float_field(f, flen, &v)) { snprintf(err, 256, "Field %d in record is not a valid float: %.*s", col, flen < 40 ? flen : 40, f); return false; }
            dense[tgt] = v;
        }
        ++col;
        if (!q) break;
        f = q + 1;
    }
    if (col != sp->n_columns) { snprintf(err, 256, "Expect %d fields but have %d in record", sp->n_columns, col); return false; }
    // field-major write-out (stable inside a field)
    const size_t base = st.keys.size(), nt = st.tok_key.size();
    st.keys.resize(base + nt);
    int32_t run = 0;
    for (int i = 0; i < F; ++i) { st.pos[i] = run; run += counts[i]; }
    for (size_t j = 0; j < nt; ++j) st.keys[base + st.pos[st.tok_field[j]]++] = st.tok_key[j];
    if (label) *label = lab;
    if (weight) *weight = sp->use_weight ? (lab > 0.5f ? sp->pos_weight : sp->neg_weight) : 1.f;
    return true;
}
}  // namespace

extern "C" int64_t wd_tsv_parse(const WdTsvSpec* sp, const char* text, int64_t text_len, int32_t n_lines,
                                int32_t* offsets_out, uint64_t* keys_out, int64_t keys_cap,
                                float* dense_out, float* label_out, float* weight_out, int32_t n_threads) {
    if (!sp || !text || n_lines < 0) { wd::set_error("wd_tsv_parse: bad arguments"); return WD_EINVAL; }
    static const bool timing = getenv("WD_TSV_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t0 = now();
    // scratch kept per calling thread across calls: a batch needs tens of MB of it, and mapping / faulting / unmapping that much
    // on every call costs more than the parsing itself (and serialises the worker threads on the kernel's address-space lock)
    static thread_local std::vector<LineRef> tl_lines;
    static thread_local std::vector<int32_t> tl_counts;
    static thread_local std::vector<float> tl_dense_tmp;
    static thread_local std::vector<ThreadOut> tl_outs;
    // (local references: a lambda run on a worker thread would otherwise name the WORKER's thread_local instances)
    std::vector<LineRef>& lines = tl_lines;
    std::vector<int32_t>& counts = tl_counts;
    std::vector<float>& dense_tmp = tl_dense_tmp;
    std::vector<ThreadOut>& outs = tl_outs;
    lines.clear();
    lines.reserve(n_lines);
    const char* p = text;
    const char* end = text + text_len;
    while (p < end && (int)lines.size() < n_lines) {
        const char* q = (const char*)memchr(p, '\n', end - p);
        const char* le = q ? q : end;
        int len = (int)(le - p);
        if (len > 0 && p[len - 1] == '\r') --len;
        lines.push_back({p, len});
        p = q ? q + 1 : end;
    }
    if ((int)lines.size() != n_lines) { wd::set_error("wd_tsv_parse: text holds %d lines, %d requested", (int)lines.size(), n_lines); return WD_EINVAL; }
    const int F = sp->n_cat_fields, Nd = sp->n_dense_fields;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_threads > n_lines / 256 + 1) n_threads = n_lines / 256 + 1;       // a thread is not worth less than a few hundred rows
    counts.resize((size_t)n_lines * (F > 0 ? F : 1));
    dense_tmp.resize(dense_out ? 0 : (size_t)n_lines * (Nd > 0 ? Nd : 1));
    if ((int)outs.size() < n_threads) outs.resize(n_threads);
    for (auto& o : outs) { o.keys.clear(); o.err.clear(); }
    // contiguous blocks of rows per thread
    auto row_lo = [&](int t) { return (int)((int64_t)n_lines * t / n_threads); };
    auto work = [&](int t) {
        ThreadOut& st = outs[t];
        const int lo = row_lo(t), hi = row_lo(t + 1);
        st.pos.assign(F > 0 ? F : 1, 0);
        st.tok_key.reserve(256); st.tok_field.reserve(256);
        st.keys.reserve((size_t)(hi - lo) * (F > 0 ? F : 1));
        char err[256];
        for (int i = lo; i < hi; ++i) {
            float* d = dense_out ? dense_out + (size_t)i * Nd : dense_tmp.data() + (size_t)i * Nd;
            if (!parse_row(sp, lines[i].p, lines[i].len, st, counts.data() + (size_t)i * F, d, (sp->has_label && label_out) ? label_out + i : nullptr,
                           weight_out ? weight_out + i : nullptr, err)) {
                st.err = err;
                return;
            }
        }
    };
    auto run_threads = [&](auto&& fn) {
        if (n_threads == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    auto t1 = now();
    run_threads(work);
    auto t2 = now();
    for (auto& o : outs) if (!o.err.empty()) { wd::set_error("%s", o.err.c_str()); return WD_EINVAL; }
    // quirk Q2 (tf_compat_pad): string fields behave like dense padded tensors -> pad every row of a string
    // field to the batch max length with Fingerprint64("")
    std::vector<int32_t> maxlen(F, 0);
    if (sp->tf_compat_pad)
        for (int i = 0; i < n_lines; ++i)
            for (int f = 0; f < F; ++f) maxlen[f] = std::max(maxlen[f], counts[(size_t)i * F + f]);
    std::vector<uint8_t> is_string(F, 0);
    for (int c = 0; c < sp->n_columns; ++c) if (sp->col_role[c] == 1) is_string[sp->col_target[c]] = 1;
    // output offsets: one serial pass of adds; per-thread starting offsets for the parallel key copy
    std::vector<int64_t> out_start(n_threads + 1, 0);
    int64_t nnz = 0;
    {
        int t = 0;
        for (int i = 0; i < n_lines; ++i) {
            while (t < n_threads && i == row_lo(t)) out_start[t++] = nnz;
            for (int f = 0; f < F; ++f) {
                const int cnt = counts[(size_t)i * F + f];
                if (offsets_out) offsets_out[(int64_t)i * F + f] = (int32_t)nnz;
                nnz += (sp->tf_compat_pad && is_string[f]) ? maxlen[f] : cnt;
            }
        }
        while (t <= n_threads) out_start[t++] = nnz;
    }
    if (offsets_out) offsets_out[(int64_t)n_lines * F] = (int32_t)nnz;
    if (nnz > 0x7fffffffLL) { wd::set_error("wd_tsv_parse: more than 2^31 keys in one batch"); return WD_EINVAL; }
    if (keys_out && keys_cap > 0) {
        if (nnz > keys_cap) { wd::set_error("wd_tsv_parse: key capacity %lld too small (need %lld)", (long long)keys_cap, (long long)nnz); return WD_EINVAL; }
        auto copy = [&](int t) {
            const ThreadOut& st = outs[t];
            const int lo = row_lo(t), hi = row_lo(t + 1);
            const uint64_t* k = st.keys.data();
            int64_t o = out_start[t];
            for (int i = lo; i < hi; ++i)
                for (int f = 0; f < F; ++f) {
                    const int cnt = counts[(size_t)i * F + f];
                    const int outc = (sp->tf_compat_pad && is_string[f]) ? maxlen[f] : cnt;
                    for (int j = 0; j < cnt; ++j) keys_out[o + j] = k[j];
                    for (int j = cnt; j < outc; ++j) keys_out[o + j] = wd::kFpEmpty;
                    o += outc;
                    k += cnt;
                }
        };
        auto t3 = now();
        run_threads(copy);
        if (timing) fprintf(stderr, "wd_tsv_parse: split %.1f ms, parse %.1f ms (%d threads), offsets %.1f ms, copy %.1f ms\n", ms(t0, t1), ms(t1, t2), n_threads, ms(t2, t3), ms(t3, now()));
    }
    return nnz;
}

// Page-locked host memory for the input pipeline (dataset.py parses TSV text straight into a ring of these buffers, so the
// asynchronous refill of a batch slot, wd_batch_prefetch_slot, really is asynchronous).  Counterpart of the buffers tf.data's
// prefetch owns in the reference's input_fn (python/lib/dataset.py:181-184).
extern "C" int wd_host_alloc(size_t bytes, void** out) {
    if (!out) { wd::set_error("wd_host_alloc: null output"); return WD_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { wd::set_error("no CUDA device: cannot page-lock host memory"); return WD_ENODEVICE; }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes > 0 ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { wd::set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return WD_ENOMEM; }
    *out = p;
    return WD_OK;
}
extern "C" int wd_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return WD_OK;
}
