// FarmHash Fingerprint64 (farmhashna::Hash64) and TensorFlow's FingerprintCat64 for host and device.
// These are the hash functions behind tf.feature_column.categorical_column_with_hash_bucket and
// crossed_column (reference python/lib/build_estimator.py:86-88, 153); results must be bit-exact.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define WD_HD __host__ __device__ __forceinline__
#else
#define WD_HD inline
#endif

namespace wd {

constexpr uint64_t kK0 = 0xc3a5c85c97cb3127ULL;
constexpr uint64_t kK1 = 0xb492b66fbe98f273ULL;
constexpr uint64_t kK2 = 0x9ae16a3b2f90404fULL;     // also Fingerprint64("")
constexpr uint64_t kFpEmpty = kK2;
constexpr uint64_t kCrossHashKey = 0xDECAFCAFFEULL;  // crossed_column default hash_key

WD_HD uint64_t ld64(const uint8_t* p) {             // unaligned little-endian loads
    uint64_t v = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
WD_HD uint32_t ld32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
WD_HD uint64_t rotr(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
WD_HD uint64_t smix(uint64_t v) { return v ^ (v >> 47); }
WD_HD uint64_t hash16(uint64_t u, uint64_t v, uint64_t mul) {
    uint64_t a = (u ^ v) * mul;
    a ^= (a >> 47);
    uint64_t b = (v ^ a) * mul;
    b ^= (b >> 47);
    return b * mul;
}
struct U128 { uint64_t lo, hi; };
WD_HD U128 weak32(const uint8_t* p, uint64_t a, uint64_t b) {
    uint64_t w = ld64(p), x = ld64(p + 8), y = ld64(p + 16), z = ld64(p + 24);
    a += w;
    b = rotr(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += rotr(a, 44);
    return U128{a + z, b + c};
}

WD_HD uint64_t fingerprint64(const uint8_t* s, size_t n) {
    if (n <= 16) {
        if (n >= 8) {
            uint64_t mul = kK2 + n * 2, a = ld64(s) + kK2, b = ld64(s + n - 8);
            return hash16(rotr(b, 37) * mul + a, (rotr(a, 25) + b) * mul, mul);
        }
        if (n >= 4) {
            uint64_t mul = kK2 + n * 2, a = ld32(s);
            return hash16(n + (a << 3), ld32(s + n - 4), mul);
        }
        if (n > 0) {
            uint32_t y = (uint32_t)s[0] + ((uint32_t)s[n >> 1] << 8);
            uint32_t z = (uint32_t)n + ((uint32_t)s[n - 1] << 2);
            return smix(y * kK2 ^ z * kK0) * kK2;
        }
        return kK2;
    }
    if (n <= 32) {
        uint64_t mul = kK2 + n * 2, a = ld64(s) * kK1, b = ld64(s + 8);
        uint64_t c = ld64(s + n - 8) * mul, d = ld64(s + n - 16) * kK2;
        return hash16(rotr(a + b, 43) + rotr(c, 30) + d, a + rotr(b + kK2, 18) + c, mul);
    }
    if (n <= 64) {
        uint64_t mul = kK2 + n * 2, a = ld64(s) * kK2, b = ld64(s + 8);
        uint64_t c = ld64(s + n - 8) * mul, d = ld64(s + n - 16) * kK2;
        uint64_t y = rotr(a + b, 43) + rotr(c, 30) + d;
        uint64_t z = hash16(y, a + rotr(b + kK2, 18) + c, mul);
        uint64_t e = ld64(s + 16) * mul, f = ld64(s + 24);
        uint64_t g = (y + ld64(s + n - 32)) * mul, h = (z + ld64(s + n - 24)) * mul;
        return hash16(rotr(e + f, 43) + rotr(g, 30) + h, e + rotr(f + a, 18) + g, mul);
    }
    uint64_t x = 81, y = 81 * kK1 + 113, z = smix(y * kK2 + 113) * kK2;
    U128 v{0, 0}, w{0, 0};
    x = x * kK2 + ld64(s);
    const uint8_t* end = s + ((n - 1) / 64) * 64;
    const uint8_t* last64 = end + ((n - 1) & 63) - 63;
    do {
        x = rotr(x + y + v.lo + ld64(s + 8), 37) * kK1;
        y = rotr(y + v.hi + ld64(s + 48), 42) * kK1;
        x ^= w.hi;
        y += v.lo + ld64(s + 40);
        z = rotr(z + w.lo, 33) * kK1;
        v = weak32(s, v.hi * kK1, x + w.lo);
        w = weak32(s + 32, z + w.hi, y + ld64(s + 16));
        uint64_t t = z; z = x; x = t;
        s += 64;
    } while (s != end);
    uint64_t mul = kK1 + ((z & 0xff) << 1);
    s = last64;
    w.lo += ((n - 1) & 63);
    v.lo += w.lo;
    w.lo += v.lo;
    x = rotr(x + y + v.lo + ld64(s + 8), 37) * mul;
    y = rotr(y + v.hi + ld64(s + 48), 42) * mul;
    x ^= w.hi * 9;
    y += v.lo * 9 + ld64(s + 40);
    z = rotr(z + w.lo, 33) * mul;
    v = weak32(s, v.hi * mul, x + w.lo);
    w = weak32(s + 32, z + w.hi, y + ld64(s + 16));
    { uint64_t t = z; z = x; x = t; }
    return hash16(hash16(v.lo, w.lo, mul) + smix(y) * kK0 + z, hash16(v.hi, w.hi, mul) + x, mul);
}

WD_HD uint64_t fingerprint_cat64(uint64_t fp1, uint64_t fp2) {
    const uint64_t kMul = 0xc6a4a7935bd1e995ULL;
    uint64_t r = fp1 ^ kMul;
    r ^= smix(fp2 * kMul) * kMul;
    r *= kMul;
    r = smix(r) * kMul;
    return smix(r);
}

}  // namespace wd
