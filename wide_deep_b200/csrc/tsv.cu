// TSV -> CSR batch loader (host, multi-threaded).
// Replaces _CsvDataset._parse_csv (reference python/lib/dataset.py:107-165): fields split on TAB
// only, no quoting; an empty field or the NA token "-" takes its default ('' / 0 / 0.0); multi-valued
// string fields split on ',' with empty tokens dropped; strings leave the loader as Fingerprint64 values
// (the same function the GPU stage uses), so no string ever crosses PCIe.  Rows are parsed in contiguous blocks, one block per
// thread, without per-row heap allocations; keys are assembled with a parallel copy.
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "farmhash.cuh"

namespace {
struct LineRef { const char* p; int len; };

// Persistent worker threads for the per-batch parse.  Spawning threads per call looked free (~20 us each) but a freshly created
// thread starts on its creator's core and a millisecond of work is over before the scheduler spreads the threads out: measured
// here, the eight workers of a 2048-line batch ran ONE AFTER THE OTHER (same wall time as one thread).  Workers that already
// sit on their own cores are woken instead; the calling thread takes tasks too.  One dispatch at a time (callers queue up).
class WorkerPool {
    // one dispatch: its own task counter, so a worker that wakes up late (or is still leaving the previous dispatch) can only ever
    // see "no task left" of the job it holds — never a task of a newer dispatch through stale state
    struct Job {
        const std::function<void(int)>* fn;
        int n;
        std::atomic<int> next{0}, done{0};
    };

public:
    void run(int n_tasks, const std::function<void(int)>& fn) {
        if (n_tasks <= 1) { if (n_tasks == 1) fn(0); return; }
        std::lock_guard<std::mutex> serial(dispatch_);
        ensure(n_tasks - 1);
        auto job = std::make_shared<Job>();
        job->fn = &fn; job->n = n_tasks;
        {
            std::lock_guard<std::mutex> lk(m_);
            cur_ = job;
            ++gen_;
        }
        cv_work_.notify_all();
        drain(*job);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return job->done.load() == job->n; });      // every task has FINISHED: fn may go out of scope
        cur_.reset();
    }

private:
    void drain(Job& j) {                             // take tasks until none is left; the last finisher wakes the dispatcher
        for (;;) {
            const int t = j.next.fetch_add(1);
            if (t >= j.n) return;
            (*j.fn)(t);
            if (j.done.fetch_add(1) + 1 == j.n) {
                std::lock_guard<std::mutex> lk(m_);  // (under the lock: the dispatcher is either before its wait or inside it)
                cv_done_.notify_all();
            }
        }
    }
    void ensure(int n) {                             // create workers: first use, or more threads asked for than exist
        while (n_workers_ < n && n_workers_ < 63) {
            std::thread([this] {
                uint64_t seen = 0;
                for (;;) {
                    std::shared_ptr<Job> job;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_work_.wait(lk, [&] { return gen_ != seen; });
                        seen = gen_;
                        job = cur_;
                    }
                    if (job) drain(*job);
                }
            }).detach();                             // detached: the workers live as long as the process (no joins at exit)
            ++n_workers_;
        }
    }
    int n_workers_ = 0;
    std::mutex dispatch_, m_;
    std::condition_variable cv_work_, cv_done_;
    std::shared_ptr<Job> cur_;
    uint64_t gen_ = 0;
};
// One pool per process, leaked on purpose (workers may outlive static destructors).  A forked child gets a NEW pool: the parent's
// workers do not exist there, and its condition variables still count them as waiters.
WorkerPool& pool() {
    static std::mutex guard;
    static WorkerPool* p = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(guard);
    if (!p || owner != getpid()) { p = new WorkerPool(); owner = getpid(); }
    return *p;
}

// Per-thread parse state: nothing is allocated per row.  Tokens of a row are collected in file order as (field, key) pairs and
// written field-major with a counting sort (fields may appear in any file column order).
// (one cache-line pair per thread: the vectors' end pointers are written on every token, and neighbouring ThreadOut headers in one
// line made the worker threads bounce it between their cores — eight threads parsed no faster than one)
struct alignas(128) ThreadOut {
    std::vector<uint64_t> keys;         // keys of this thread's rows, row after row, field-major inside a row
    std::vector<uint64_t> tok_key;      // scratch: tokens of the current row
    std::vector<int32_t> tok_field;
    std::vector<int32_t> pos;           // scratch: per-field write cursor
    std::string err;
};

// Integers and floats are what strtoll / strtof make of the field (the reference's decode_csv), but the common shapes are decoded
// here: glibc's converters cost 30-100 ns per field (locale lookup, general rounding machinery) and a record holds a dozen of them.
inline bool parse_int_slow(const char* f, int flen, long long* out) {
    char buf[48];
    if (flen >= (int)sizeof(buf)) return false;
    memcpy(buf, f, flen); buf[flen] = 0;
    char* ep = nullptr;
    *out = strtoll(buf, &ep, 10);
    return *ep == 0;
}
inline bool parse_int_field(const char* f, int flen, long long* out) {
    // [-]digits, at most 18 of them (no overflow possible); anything else (spaces, '+', longer) takes the strtoll path
    int i = 0;
    const bool neg = flen > 0 && f[0] == '-';
    if (neg) i = 1;
    const int nd = flen - i;
    if (nd < 1 || nd > 18) return parse_int_slow(f, flen, out);
    long long v = 0;
    for (; i < flen; ++i) {
        const unsigned d = (unsigned)(f[i] - '0');
        if (d > 9) return parse_int_slow(f, flen, out);
        v = v * 10 + d;
    }
    *out = neg ? -v : v;
    return true;
}
inline bool parse_float_slow(const char* f, int flen, float* out) {
    char buf[64];
    if (flen >= (int)sizeof(buf)) return false;
    memcpy(buf, f, flen); buf[flen] = 0;
    char* ep = nullptr;
    *out = strtof(buf, &ep);
    return *ep == 0;
}
inline bool parse_float_field(const char* f, int flen, float* out) {
    // [-]digits[.digits] with at most 15 significant digits: mantissa and power of ten are exact doubles, so their quotient is the
    // correctly rounded DOUBLE of the decimal (Clinger's fast path).  Rounding that double to float equals rounding the decimal to
    // float unless the double sits on (or next to) a midpoint between two floats — its low 29 mantissa bits then read 0x0FFFFFFF,
    // 0x10000000 or 0x10000001 — and exactly those cases, like every other shape (exponents, inf / nan, spaces), go to strtof.
    static const double kPow10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    int i = 0;
    const bool neg = flen > 0 && f[0] == '-';
    if (neg) i = 1;
    unsigned long long m = 0;
    int nd = 0, frac = 0;
    bool dot = false, any = false;
    for (; i < flen; ++i) {
        const char c = f[i];
        if (c == '.') {
            if (dot) return parse_float_slow(f, flen, out);
            dot = true;
            continue;
        }
        const unsigned d = (unsigned)(c - '0');
        if (d > 9) return parse_float_slow(f, flen, out);
        any = true;
        if (nd > 0 || d != 0) {                             // significant digits (leading zeros do not count)
            if (++nd > 15) return parse_float_slow(f, flen, out);
            m = m * 10 + d;
        }
        if (dot && ++frac > 15) return parse_float_slow(f, flen, out);
    }
    if (!any) return parse_float_slow(f, flen, out);
    if (m == 0) { *out = neg ? -0.f : 0.f; return true; }
    const double d = (double)m / kPow10[frac];
    if (!(d > 1e-30 && d < 1e30)) return parse_float_slow(f, flen, out);    // far from float's subnormal / overflow ranges
    unsigned long long bits;
    memcpy(&bits, &d, 8);
    const unsigned low = (unsigned)(bits & 0x1FFFFFFFull);
    if (low >= 0x0FFFFFFFu && low <= 0x10000001u) return parse_float_slow(f, flen, out);
    *out = neg ? -(float)d : (float)d;
    return true;
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

namespace {
// What the last counting call (keys_out == NULL or a capacity that turned out too small) parsed on this thread: a following call for
// the SAME lines and output arrays only has to copy the keys out — the two-call protocol (size, then fill) parses once.
struct Pending {
    bool valid = false;
    const WdTsvSpec* sp = nullptr;
    const char* first = nullptr; const char* last = nullptr;
    int32_t n_lines = 0, n_threads = 0;
    const void *offsets = nullptr, *dense = nullptr, *label = nullptr, *weight = nullptr;
    int64_t nnz = 0, bytes = 0;
    uint64_t fp_first = 0, fp_last = 0;      // the pointers alone could be a recycled allocation holding other text
};

// parse `lines` (already split) into the CSR batch; see wd_tsv_parse for the contract
int64_t parse_lines(const WdTsvSpec* sp, std::vector<LineRef>& lines, int32_t n_lines, int32_t* offsets_out, uint64_t* keys_out, int64_t keys_cap,
                    float* dense_out, float* label_out, float* weight_out, int32_t n_threads, double split_ms) {
    static const bool timing = getenv("WD_TSV_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    // scratch kept per calling thread across calls: a batch needs tens of MB of it, and mapping / faulting / unmapping that much
    // on every call costs more than the parsing itself (and serialises the worker threads on the kernel's address-space lock)
    static thread_local std::vector<int32_t> tl_counts;
    static thread_local std::vector<float> tl_dense_tmp;
    static thread_local std::vector<ThreadOut> tl_outs;
    static thread_local std::vector<int32_t> tl_maxlen;
    static thread_local Pending tl_pending;
    // (local references: a lambda run on a worker thread would otherwise name the WORKER's thread_local instances)
    std::vector<int32_t>& counts = tl_counts;
    std::vector<float>& dense_tmp = tl_dense_tmp;
    std::vector<ThreadOut>& outs = tl_outs;
    std::vector<int32_t>& maxlen = tl_maxlen;
    Pending& pend = tl_pending;
    const int F = sp->n_cat_fields, Nd = sp->n_dense_fields;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_threads > n_lines / 256 + 1) n_threads = n_lines / 256 + 1;       // a thread is not worth less than a few hundred rows
    auto row_lo = [&](int t) { return (int)((int64_t)n_lines * t / n_threads); };
    auto run_threads = [&](const std::function<void(int)>& fn) { pool().run(n_threads, fn); };
    const char* first = n_lines > 0 ? lines[0].p : nullptr;
    const char* last = n_lines > 0 ? lines[n_lines - 1].p : nullptr;
    int64_t bytes = 0;
    for (int i = 0; i < n_lines; ++i) bytes += lines[i].len;
    const uint64_t fp_first = n_lines > 0 ? wd::fingerprint64((const uint8_t*)lines[0].p, lines[0].len) : 0;
    const uint64_t fp_last = n_lines > 0 ? wd::fingerprint64((const uint8_t*)lines[n_lines - 1].p, lines[n_lines - 1].len) : 0;
    const bool reuse = pend.valid && pend.sp == sp && pend.first == first && pend.last == last && pend.n_lines == n_lines && pend.n_threads == n_threads &&
                       pend.offsets == offsets_out && pend.dense == dense_out && pend.label == label_out && pend.weight == weight_out && dense_out &&
                       pend.bytes == bytes && pend.fp_first == fp_first && pend.fp_last == fp_last;
    pend.valid = false;
    auto t1 = now();
    if (!reuse) {
        counts.resize((size_t)n_lines * (F > 0 ? F : 1));
        dense_tmp.resize(dense_out ? 0 : (size_t)n_lines * (Nd > 0 ? Nd : 1));
        if ((int)outs.size() < n_threads) outs.resize(n_threads);
        for (auto& o : outs) { o.keys.clear(); o.err.clear(); }
        // contiguous blocks of rows per thread
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
        run_threads(work);
        for (auto& o : outs) if (!o.err.empty()) { wd::set_error("%s", o.err.c_str()); return WD_EINVAL; }
        // quirk Q2 (tf_compat_pad): string fields behave like dense padded tensors -> pad every row of a string
        // field to the batch max length with Fingerprint64("")
        maxlen.assign(F > 0 ? F : 1, 0);
        if (sp->tf_compat_pad)
            for (int i = 0; i < n_lines; ++i)
                for (int f = 0; f < F; ++f) maxlen[f] = std::max(maxlen[f], counts[(size_t)i * F + f]);
    }
    auto t2 = now();
    std::vector<uint8_t> is_string(F > 0 ? F : 1, 0);
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
    if (!(keys_out && keys_cap > 0) || nnz > keys_cap) {
        // counting call, or the caller's buffer is too small: nothing is copied, the parsed state stays for the follow-up call
        pend.valid = true; pend.sp = sp; pend.first = first; pend.last = last; pend.n_lines = n_lines; pend.n_threads = n_threads;
        pend.offsets = offsets_out; pend.dense = dense_out; pend.label = label_out; pend.weight = weight_out; pend.nnz = nnz;
        pend.bytes = bytes; pend.fp_first = fp_first; pend.fp_last = fp_last;
        return nnz;
    }
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
    if (timing) fprintf(stderr, "wd_tsv_parse: split %.1f ms, parse %.1f ms%s (%d threads), offsets %.1f ms, copy %.1f ms\n", split_ms, ms(t1, t2),
                        reuse ? " (reused)" : "", n_threads, ms(t2, t3), ms(t3, now()));
    return nnz;
}
thread_local std::vector<LineRef> tl_lines;
}  // namespace

extern "C" int64_t wd_tsv_parse(const WdTsvSpec* sp, const char* text, int64_t text_len, int32_t n_lines,
                                int32_t* offsets_out, uint64_t* keys_out, int64_t keys_cap,
                                float* dense_out, float* label_out, float* weight_out, int32_t n_threads) {
    if (!sp || !text || n_lines < 0) { wd::set_error("wd_tsv_parse: bad arguments"); return WD_EINVAL; }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<LineRef>& lines = tl_lines;
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
    const double split_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return parse_lines(sp, lines, n_lines, offsets_out, keys_out, keys_cap, dense_out, label_out, weight_out, n_threads, split_ms);
}

// Line index of a whole file image: starts / lengths of its non-empty lines (a trailing '\r' is not part of a line), so a shuffled
// or sharded pass over the file is a permutation of indices — the text itself is never split, copied or joined again.
extern "C" int64_t wd_tsv_index_lines(const char* text, int64_t text_len, int64_t* starts_out, int32_t* lens_out, int64_t cap) {
    if (!text || text_len < 0) { wd::set_error("wd_tsv_index_lines: bad arguments"); return WD_EINVAL; }
    int64_t n = 0;
    const char* p = text;
    const char* end = text + text_len;
    while (p < end) {
        const char* q = (const char*)memchr(p, '\n', end - p);
        const char* le = q ? q : end;
        int64_t len = le - p;
        if (len > 0) {                                          // dataset.py keeps every non-empty line ("\r" alone counts as one)
            if (len > 0x7fffffffLL) { wd::set_error("wd_tsv_index_lines: line longer than 2^31 bytes"); return WD_EINVAL; }
            if (starts_out && lens_out && n < cap) {
                starts_out[n] = p - text;
                lens_out[n] = (int32_t)((p[len - 1] == '\r') ? len - 1 : len);
            }
            ++n;
        }
        p = q ? q + 1 : end;
    }
    return n;
}

// wd_tsv_parse over lines picked by index: line i of the batch = text[starts[idx[i]] .. + lens[idx[i]]) (idx NULL: i itself).
extern "C" int64_t wd_tsv_parse_lines(const WdTsvSpec* sp, const char* text, const int64_t* starts, const int32_t* lens, const int64_t* idx,
                                      int32_t n_lines, int32_t* offsets_out, uint64_t* keys_out, int64_t keys_cap,
                                      float* dense_out, float* label_out, float* weight_out, int32_t n_threads) {
    if (!sp || !text || !starts || !lens || n_lines < 0) { wd::set_error("wd_tsv_parse_lines: bad arguments"); return WD_EINVAL; }
    std::vector<LineRef>& lines = tl_lines;
    lines.resize(n_lines);
    for (int i = 0; i < n_lines; ++i) {
        const int64_t j = idx ? idx[i] : i;
        lines[i] = {text + starts[j], lens[j]};
    }
    return parse_lines(sp, lines, n_lines, offsets_out, keys_out, keys_cap, dense_out, label_out, weight_out, n_threads, 0.0);
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
