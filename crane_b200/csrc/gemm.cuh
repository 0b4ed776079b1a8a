// Prefill / ViT GEMM:  C[M,N] (+)= A[M,K] (bf16, row-major) x W[N,K]^T (bf16, row-major = K-major)
//
// B200 design: a persistent kernel, one CTA per SM, stream-K over the (128 x BN output tile, 64-wide k-step) iteration space; A and W
// tiles stream HBM/L2 -> shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a multi-stage mbarrier ring; one elected
// thread issues tcgen05.mma (UMMA 128 x BN x 16, bf16 x bf16 -> f32) into a double-buffered TMEM accumulator; four epilogue warps
// pull it back with tcgen05.ld and apply the fused epilogue (bias / residual add / SiLU*up / GELU / dtype cast) straight into
// global memory, or park a partial tile for the tile's last contributor to sum in a fixed order (see gemm.cu).
//
// Replaces the cuBLAS GEMMs Candle launches for every `Linear` on the reference's GPU path and the
// f32 gemm on its CPU path (crane-core/src/models/qwen3/modeling.rs:318-329,532,608-642;
// qwen3_5/vision.rs:46-58,76-80,129,174,275-277).
#pragma once

#include <cuda.h>
#include "common.cuh"

namespace cb {

enum GemmEpiMode : int {
    EPI_STORE_F32 = 0,     // out f32 [M, ldo] = acc (+ bias)
    EPI_STORE_BF16 = 1,    // out bf16 [M, ldo] = acc (+ bias)
    EPI_RESID_F32 = 2,     // out f32 [M, ldo] += acc (+ bias)           (residual stream, in place)
    EPI_SILU_MUL_BF16 = 3, // W rows interleaved (gate_j, up_j): out bf16 [M, ldo][j] = silu(acc[2j]) * acc[2j+1]
    EPI_GELU_ERF_BF16 = 4, // out bf16 = gelu_erf(acc + bias)
    EPI_GELU_TANH_BF16 = 5 // out bf16 = gelu_tanh(acc + bias)
};

struct GemmEpi {
    void* out;
    void* out_lo;       // bf16-output modes: optional second plane receiving bf16(value - bf16(value)) (split-precision operand)
    int ldo;            // leading dimension of `out` in elements
    const float* bias;  // [N] or nullptr
    int mode;
    bool w_dynamic = false;   // W was written by an earlier kernel of the chain (dequantised scratch): no weight prefetch before the dependency wait
    unsigned long long* prof = nullptr;   // optional: 8 globaltimer stamps per CTA (tile timeline, tools/gemm_probe.py)
};

struct TmaEncoder;  // host: resolves cuTensorMapEncodeTiled once

// Host API.  `use_simt` selects the slow debugging kernel (bring-up A/B only).
// Returns cudaError_t (as int) or -1000 for unsupported shapes.
// `A_lo` (nullable): low-order plane of a split-precision activation operand, same shape/stride as A (see gemm.cu).
int gemm_bf16_launch(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K,
                     const GemmEpi& epi, bool use_simt);
// Frees the stream-K scratch (partial tiles, arrival counters) kept for `stream`; call before destroying the stream.
void gemm_release_stream(cudaStream_t stream);

}  // namespace cb
