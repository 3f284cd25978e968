// TSV -> CSR batch loader (host, multi-threaded).
// Replaces _CsvDataset._parse_csv (reference python/lib/dataset.py:107-165): fields split on TAB
// only, no quoting; an empty field or the NA token "-" takes its default ('' / 0 / 0.0); multi-valued
// string fields split on ',' with empty tokens dropped; strings leave the loader as Fingerprint64 values
// (the same function the GPU stage uses), so no string ever crosses PCIe.
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"
#include "farmhash.cuh"

namespace {
struct LineRef { const char* p; int len; };

struct RowOut {
    std::vector<uint64_t> keys;        // all fields of the row, field-major
    std::vector<int32_t> counts;       // per cat field
};

bool parse_row(const WdTsvSpec* sp, const char* p, int len, RowOut& out, float* dense, float* label, float* weight, char* err) {
    out.keys.clear();
    out.counts.assign(sp->n_cat_fields, 0);
    // per-field temporary lists (fields may appear in any file column order)
    std::vector<std::vector<uint64_t>> fk(sp->n_cat_fields);
    int col = 0;
    const char* end = p + len;
    const char* f = p;
    for (int i = 0; i < sp->n_dense_fields; ++i) dense[i] = 0.f;
    float lab = 0.f;
    while (true) {
        const char* q = (const char*)memchr(f, '\t', end - f);
        const char* fe = q ? q : end;
        if (col >= sp->n_columns) { snprintf(err, 256, "Expect %d fields but have more in record", sp->n_columns); return false; }
        int role = sp->col_role[col], tgt = sp->col_target[col];
        int flen = (int)(fe - f);
        bool na = flen == 0 || (flen == 1 && f[0] == '-');
        if (role == 0) {
            lab = na ? 0.f : ((flen == 1 && f[0] == '1') || strtol(std::string(f, flen).c_str(), nullptr, 10) == 1 ? 1.f : 0.f);
        } else if (role == 1) {
            if (!na) {
                if (sp->multivalue) {
                    const char* t = f;
                    while (t <= fe) {
                        const char* c = (const char*)memchr(t, ',', fe - t);
                        const char* te = c ? c : fe;
                        if (te > t) fk[tgt].push_back(wd::fingerprint64((const uint8_t*)t, te - t));
                        if (!c) break;
                        t = c + 1;
                    }
                } else {
                    fk[tgt].push_back(wd::fingerprint64((const uint8_t*)f, flen));
                }
            }
        } else if (role == 2) {
            long long v = 0;
            if (!na) {
                char* ep = nullptr;
                std::string s(f, flen);
                v = strtoll(s.c_str(), &ep, 10);
                if (*ep != 0) { snprintf(err, 256, "Field %d in record is not a valid int32: %s", col, s.c_str()); return false; }
            }
            fk[tgt].push_back((uint64_t)v);
        } else if (role == 3) {
            float v = 0.f;
            if (!na) {
                char* ep = nullptr;
                std::string s(f, flen);
                v = strtof(s.c_str(), &ep);
                if (*ep != 0) { snprintf(err, 256, "Field %d in record is not a valid float: %s", col, s.c_str()); return false; }
            }
            dense[tgt] = v;
        }
        ++col;
        if (!q) break;
        f = q + 1;
    }
    if (col != sp->n_columns) { snprintf(err, 256, "Expect %d fields but have %d in record", sp->n_columns, col); return false; }
    for (int i = 0; i < sp->n_cat_fields; ++i) {
        out.counts[i] = (int32_t)fk[i].size();
        out.keys.insert(out.keys.end(), fk[i].begin(), fk[i].end());
    }
    if (label) *label = lab;
    if (weight) *weight = sp->use_weight ? (lab > 0.5f ? sp->pos_weight : sp->neg_weight) : 1.f;
    return true;
}
}  // namespace

extern "C" int64_t wd_tsv_parse(const WdTsvSpec* sp, const char* text, int64_t text_len, int32_t n_lines,
                                int32_t* offsets_out, uint64_t* keys_out, int64_t keys_cap,
                                float* dense_out, float* label_out, float* weight_out, int32_t n_threads) {
    if (!sp || !text || n_lines < 0) { wd::set_error("wd_tsv_parse: bad arguments"); return WD_EINVAL; }
    std::vector<LineRef> lines;
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
    std::vector<RowOut> rows(n_lines);
    std::vector<float> dense_tmp((size_t)n_lines * (Nd > 0 ? Nd : 1));
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    std::vector<std::string> errs(n_threads);
    auto work = [&](int t) {
        char err[256];
        for (int i = t; i < n_lines; i += n_threads) {
            float* d = dense_out ? dense_out + (size_t)i * Nd : dense_tmp.data() + (size_t)i * Nd;
            if (!parse_row(sp, lines[i].p, lines[i].len, rows[i], d, (sp->has_label && label_out) ? label_out + i : nullptr,
                           weight_out ? weight_out + i : nullptr, err)) {
                if (errs[t].empty()) errs[t] = err;
                return;
            }
        }
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (auto& e : errs) if (!e.empty()) { wd::set_error("%s", e.c_str()); return WD_EINVAL; }
    // quirk Q2 (tf_compat_pad): string fields behave like dense padded tensors -> pad every row of a string
    // field to the batch max length with Fingerprint64("")
    std::vector<int32_t> maxlen(F, 0);
    if (sp->tf_compat_pad)
        for (int i = 0; i < n_lines; ++i)
            for (int f = 0; f < F; ++f) maxlen[f] = std::max(maxlen[f], rows[i].counts[f]);
    std::vector<uint8_t> is_string(F, 0);
    for (int c = 0; c < sp->n_columns; ++c) if (sp->col_role[c] == 1) is_string[sp->col_target[c]] = 1;
    int64_t nnz = 0;
    for (int i = 0; i < n_lines; ++i) {
        const uint64_t* k = rows[i].keys.data();
        for (int f = 0; f < F; ++f) {
            int cnt = rows[i].counts[f];
            int outc = (sp->tf_compat_pad && is_string[f]) ? maxlen[f] : cnt;
            if (offsets_out) offsets_out[(int64_t)i * F + f] = (int32_t)nnz;
            if (keys_out && keys_cap > 0) {
                if (nnz + outc > keys_cap) { wd::set_error("wd_tsv_parse: key capacity %lld too small", (long long)keys_cap); return WD_EINVAL; }
                for (int j = 0; j < cnt; ++j) keys_out[nnz + j] = k[j];
                for (int j = cnt; j < outc; ++j) keys_out[nnz + j] = wd::kFpEmpty;
            }
            nnz += outc;
            k += cnt;
        }
    }
    if (offsets_out) offsets_out[(int64_t)n_lines * F] = (int32_t)nnz;
    return nnz;
}
