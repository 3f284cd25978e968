// Shared GEMM interface of the MLP engines (FFMA in mlp.cu, tcgen05 in gemm_tc.cu and gemm_bf16.cu): operand descriptions,
// epilogue description and the activation functions (reference python/lib/utils/model_util.py:28-59).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace wd {

// --------------------------------------------------------------------------------------------- activations
__device__ __forceinline__ float act_fwd(int kind, float z) {
    switch (kind) {
        case WD_ACT_RELU: return fmaxf(z, 0.f);
        case WD_ACT_RELU6: return fminf(fmaxf(z, 0.f), 6.f);
        case WD_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
        case WD_ACT_TANH: return tanhf(z);
        case WD_ACT_LEAKY_RELU: return z > 0.f ? z : 0.2f * z;
        case WD_ACT_ELU: return z > 0.f ? z : expm1f(z);
        case WD_ACT_SELU: return 1.0507009873554805f * (z > 0.f ? z : 1.6732632423543772f * expm1f(z));
        case WD_ACT_SOFTPLUS: return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
        case WD_ACT_SOFTSIGN: return z / (1.f + fabsf(z));
    }
    return z;
}
// derivative expressed through the stored post-activation value a
__device__ __forceinline__ float act_bwd(int kind, float a) {
    switch (kind) {
        case WD_ACT_RELU: return a > 0.f ? 1.f : 0.f;
        case WD_ACT_RELU6: return (a > 0.f && a < 6.f) ? 1.f : 0.f;
        case WD_ACT_SIGMOID: return a * (1.f - a);
        case WD_ACT_TANH: return 1.f - a * a;
        case WD_ACT_LEAKY_RELU: return a > 0.f ? 1.f : 0.2f;
        case WD_ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
        case WD_ACT_SELU: return a > 0.f ? 1.0507009873554805f : a + 1.0507009873554805f * 1.6732632423543772f;
        case WD_ACT_SOFTPLUS: return 1.f - expf(-a);
        case WD_ACT_SOFTSIGN: { float t = 1.f - fabsf(a); return t * t; }
    }
    return 1.f;
}

// ------------------------------------------------------------------------------------------- FFMA GEMM
struct GemmA {                         // A operand: up to kMaxSegs K-contiguous segments
    int n;
    const float* ptr[kMaxSegs];
    int ld[kMaxSegs];
    int k[kMaxSegs];                   // multiple of 16
    const __nv_bfloat16* hi[kMaxSegs]; // 3xBF16 engine: the same segments pre-split into bf16 hi / lo copies (same ld)
    const __nv_bfloat16* lo[kMaxSegs];
};
// EPI_DACT (3xBF16 pair kernel only): a data-gradient GEMM whose epilogue is the activation / batch-norm backward of the layer
// it feeds — dH never reaches memory: dZ = dH * gamma' * act'(A) leaves as bf16 hi / lo copies and the 128-row column partials of
// the bias / gamma / beta gradients go to the partial arena (what act_bn_bwd_q_kernel does in a separate pass otherwise)
enum { EPI_FWD = 0, EPI_STORE = 1, EPI_WGRAD = 2, EPI_DACT = 3 };
struct Epi {
    float* C; int ldc;                 // STORE / WGRAD target
    int accumulate;                    // STORE: C += acc
    float* A_out; float* H_out; int ldh;   // FWD outputs
    float* HT; int ldt;                // FWD transposed output (nullable)
    const float *bias, *gamma, *beta;
    int n_logical, act, bn;
    int m_valid;                       // rows >= m_valid are written as zero (transposed padding)
    int64_t split_stride;              // WGRAD: floats between split partials
    // 3xBF16 engine, FWD: the layer output leaves as bf16 hi / lo copies, row-major [M, ldh] (H_out may then be NULL)
    __nv_bfloat16 *Hs_hi, *Hs_lo;
    // EPI_DACT: stored post-activation values of the fed layer [M, ldh], its partial arenas ([128-row tile][pstride]); dZ hi / lo
    // copies go to Hs_hi / Hs_lo (row-major [M, ldh]); gamma / act / bn / n_logical as in FWD
    const float* Aact;
    float *p_bias, *p_gamma, *p_beta;
    int64_t pstride;
};

// ---- dropout (tf.layers.dropout after a hidden layer's activation, TRAIN only; reference dnn.py:111-112).  TensorFlow's random
// stream cannot be reproduced, so the keep mask is DEFINED by a counter-based generator shared bit for bit with the oracle
// (oracle/model.py drop_keep): element (m, n) of layer `layer_id` in train step `step` is kept iff
//   u >= rate,  u = top 24 bits of splitmix64(key ^ (m * 65536 + n)) / 2^24,  key = splitmix64(seed ^ step * GOLDEN ^ layer_id << 48)
// and kept elements are scaled by 1 / (1 - rate).  Nothing is stored: the backward regenerates the mask.
struct DropArgs { float rate; unsigned long long seed; const unsigned int* step; int layer_id; };
__device__ __forceinline__ unsigned long long splitmix64_dev(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long drop_key(const DropArgs& d) {
    return splitmix64_dev(d.seed ^ ((unsigned long long)(*d.step) * 0x9E3779B97F4A7C15ULL) ^ ((unsigned long long)d.layer_id << 48));
}
__device__ __forceinline__ float drop_mult(unsigned long long key, unsigned int m, unsigned int n, float rate, float inv_keep) {
    const unsigned long long r = splitmix64_dev(key ^ ((unsigned long long)m * 65536ULL + n));
    const float u = (float)(unsigned int)(r >> 40) * (1.f / 16777216.f);
    return u >= rate ? inv_keep : 0.f;
}

// hi = bf16(x) (round to nearest), lo = bf16(x - hi): x = hi + lo up to 2^-17 relative; the product a*b is rebuilt as
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi on the bf16 tensor pipe with fp32 accumulation (dropped term ~2^-18)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}


}  // namespace wd
