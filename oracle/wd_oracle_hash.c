/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * CPU restatement (plain C) of the integer arithmetic the reference reaches through its
 * un-vendored dependency `tensorflow` (requirements.txt:1, pin ">=1.4"):
 *
 *   - farmhash::Fingerprint64 (= farmhashna::Hash64), used by
 *       categorical_column_with_hash_bucket      (reference python/lib/build_estimator.py:86-88)
 *       and by SparseCross for string keys       (reference python/lib/build_estimator.py:153)
 *   - tensorflow::FingerprintCat64 + the SparseCross hashed chain with
 *       hash_key 0xDECAFCAFFE                    (reference python/lib/build_estimator.py:153)
 *   - Bucketize (upper_bound over fp32 boundaries)(reference python/lib/build_estimator.py:133,145)
 *
 * The algorithm is restated from the published FarmHash (farmhashna) and TensorFlow
 * (core/platform/fingerprint.h, core/kernels/sparse_cross_op.cc) sources; SURVEY.md
 * Appendix B.1 holds the language-neutral spec this file follows line by line.
 *
 * Parity pin: the reference's own tests hold no golden values for this path
 * (python/lib/wide_deep_test.py:81-85 only asserts monotone improvement) => the oracle is
 * pinned on TensorFlow-upstream known-answer vectors (tests/golden/hash_kat.json); the
 * reference itself cannot run here (no TensorFlow) -> "parity unpinned" w.r.t. a live
 * reference run, see DESIGN.md.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

static const uint64_t k0 = 0xc3a5c85c97cb3127ULL;
static const uint64_t k1 = 0xb492b66fbe98f273ULL;
static const uint64_t k2 = 0x9ae16a3b2f90404fULL;

static inline uint64_t fetch64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t fetch32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
static inline uint64_t smix(uint64_t v) { return v ^ (v >> 47); }

static inline uint64_t h16(uint64_t u, uint64_t v, uint64_t mul) {
    uint64_t a = (u ^ v) * mul;
    a ^= (a >> 47);
    uint64_t b = (v ^ a) * mul;
    b ^= (b >> 47);
    b *= mul;
    return b;
}

static uint64_t len0to16(const uint8_t *s, size_t n) {
    if (n >= 8) {
        uint64_t mul = k2 + n * 2;
        uint64_t a = fetch64(s) + k2;
        uint64_t b = fetch64(s + n - 8);
        uint64_t c = rot(b, 37) * mul + a;
        uint64_t d = (rot(a, 25) + b) * mul;
        return h16(c, d, mul);
    }
    if (n >= 4) {
        uint64_t mul = k2 + n * 2;
        uint64_t a = fetch32(s);
        return h16(n + (a << 3), fetch32(s + n - 4), mul);
    }
    if (n > 0) {
        uint8_t a = s[0], b = s[n >> 1], c = s[n - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
        uint32_t z = (uint32_t)n + ((uint32_t)c << 2);
        return smix(y * k2 ^ z * k0) * k2;
    }
    return k2;
}

static uint64_t len17to32(const uint8_t *s, size_t n) {
    uint64_t mul = k2 + n * 2;
    uint64_t a = fetch64(s) * k1;
    uint64_t b = fetch64(s + 8);
    uint64_t c = fetch64(s + n - 8) * mul;
    uint64_t d = fetch64(s + n - 16) * k2;
    return h16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + k2, 18) + c, mul);
}

static uint64_t len33to64(const uint8_t *s, size_t n) {
    uint64_t mul = k2 + n * 2;
    uint64_t a = fetch64(s) * k2;
    uint64_t b = fetch64(s + 8);
    uint64_t c = fetch64(s + n - 8) * mul;
    uint64_t d = fetch64(s + n - 16) * k2;
    uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
    uint64_t z = h16(y, a + rot(b + k2, 18) + c, mul);
    uint64_t e = fetch64(s + 16) * mul;
    uint64_t f = fetch64(s + 24);
    uint64_t g = (y + fetch64(s + n - 32)) * mul;
    uint64_t h = (z + fetch64(s + n - 24)) * mul;
    return h16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
}

static inline void weak32(const uint8_t *p, uint64_t a, uint64_t b, uint64_t *o0, uint64_t *o1) {
    uint64_t w = fetch64(p), x = fetch64(p + 8), y = fetch64(p + 16), z = fetch64(p + 24);
    a += w;
    b = rot(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += rot(a, 44);
    *o0 = a + z;
    *o1 = b + c;
}

uint64_t wdo_fingerprint64(const uint8_t *s, size_t n) {
    if (n <= 32) return n <= 16 ? len0to16(s, n) : len17to32(s, n);
    if (n <= 64) return len33to64(s, n);
    uint64_t x = 81;
    uint64_t y = 81 * k1 + 113;
    uint64_t z = smix(y * k2 + 113) * k2;
    uint64_t v0 = 0, v1 = 0, w0 = 0, w1 = 0;
    x = x * k2 + fetch64(s);
    const uint8_t *end = s + ((n - 1) / 64) * 64;
    const uint8_t *last64 = end + ((n - 1) & 63) - 63;
    do {
        x = rot(x + y + v0 + fetch64(s + 8), 37) * k1;
        y = rot(y + v1 + fetch64(s + 48), 42) * k1;
        x ^= w1;
        y += v0 + fetch64(s + 40);
        z = rot(z + w0, 33) * k1;
        weak32(s, v1 * k1, x + w0, &v0, &v1);
        weak32(s + 32, z + w1, y + fetch64(s + 16), &w0, &w1);
        uint64_t t = z; z = x; x = t;
        s += 64;
    } while (s != end);
    uint64_t mul = k1 + ((z & 0xff) << 1);
    s = last64;
    w0 += ((n - 1) & 63);
    v0 += w0;
    w0 += v0;
    x = rot(x + y + v0 + fetch64(s + 8), 37) * mul;
    y = rot(y + v1 + fetch64(s + 48), 42) * mul;
    x ^= w1 * 9;
    y += v0 * 9 + fetch64(s + 40);
    z = rot(z + w0, 33) * mul;
    weak32(s, v1 * mul, x + w0, &v0, &v1);
    weak32(s + 32, z + w1, y + fetch64(s + 16), &w0, &w1);
    { uint64_t t = z; z = x; x = t; }
    return h16(h16(v0, w0, mul) + smix(y) * k0 + z, h16(v1, w1, mul) + x, mul);
}

uint64_t wdo_fingerprint_cat64(uint64_t fp1, uint64_t fp2) {
    const uint64_t kMul = 0xc6a4a7935bd1e995ULL;
    uint64_t r = fp1 ^ kMul;
    r ^= smix(fp2 * kMul) * kMul;
    r *= kMul;
    r = smix(r) * kMul;
    r = smix(r);
    return r;
}

/* batch: fingerprints of n byte strings stored back-to-back, offs[n+1] */
void wdo_fingerprint64_batch(const uint8_t *bytes, const int64_t *offs, int64_t n, uint64_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = wdo_fingerprint64(bytes + offs[i], (size_t)(offs[i + 1] - offs[i]));
}

/* SparseCross hashed chain over one tuple of keys (already in op order). */
uint64_t wdo_cross_chain(const uint64_t *keys, int nkeys, uint64_t hash_key) {
    uint64_t h = hash_key;
    for (int i = 0; i < nkeys; ++i) h = wdo_fingerprint_cat64(h, keys[i]);
    return h;
}

/*
 * Row-wise Cartesian-product cross (sparse_cross_op.cc ProductIterator: last column innermost).
 * ncols key columns in op order; column c holds CSR (offs[c][B+1], vals[c][...]).
 * Writes ids (h % num_buckets) and per-row output offsets; returns total count.
 * If out_ids == NULL only counts.
 */
int64_t wdo_cross_rows(int ncols, const int64_t *const *offs, const uint64_t *const *vals, int64_t B,
                       uint64_t num_buckets, uint64_t hash_key, int64_t *out_offs, int64_t *out_ids) {
    int64_t total = 0;
    int idx[16];
    for (int64_t b = 0; b < B; ++b) {
        if (out_offs) out_offs[b] = total;
        int empty = 0;
        for (int c = 0; c < ncols; ++c) { idx[c] = 0; if (offs[c][b + 1] == offs[c][b]) empty = 1; }
        if (empty) continue;
        for (;;) {
            uint64_t h = hash_key;
            for (int c = 0; c < ncols; ++c) h = wdo_fingerprint_cat64(h, vals[c][offs[c][b] + idx[c]]);
            if (out_ids) out_ids[total] = (int64_t)(h % num_buckets);
            ++total;
            int c = ncols - 1;
            for (; c >= 0; --c) {
                if (++idx[c] < offs[c][b + 1] - offs[c][b]) break;
                idx[c] = 0;
            }
            if (c < 0) break;
        }
    }
    if (out_offs) out_offs[B] = total;
    return total;
}

/* Bucketize: id = number of boundaries <= x (upper_bound), boundaries fp32 ascending. */
void wdo_bucketize(const float *x, int64_t n, const float *bounds, int nb, int64_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        int lo = 0, hi = nb;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (bounds[mid] <= x[i]) lo = mid + 1; else hi = mid; }
        out[i] = lo;
    }
}
