// Device-side sampler behind the lm_head: penalties, exact top-k order, top-p, Gumbel-max.  See sampler.cu.
#pragma once

#include "common.cuh"

namespace cb {

constexpr int TOPK_MAX = 512;          // as the reference's kernel (crane-core/tests/rocm_kernels.rs:117-131); the sampler itself caps k at 64

struct SamplerRow {                    // one logits row
    float temperature;                 // <= 0: greedy
    float top_p;                       // (0, 1): nucleus over the top-k candidates
    int top_k;                         // candidates (1 .. 64), or 0: draw over the whole vocabulary
    float rep_mul_pos, rep_mul_neg;    // repetition penalty as the two factors candle's affine ops apply (1, 1 = off)
    float frequency_penalty, presence_penalty;
    int ctx_off, ctx_len;              // this row's slice of the context array
    int uni_off;                       // this row's slice of the uniforms array (-1: draw on the device from seed)
    unsigned long long seed;
};

// penalties in place on logits[rows, V]
int sampler_penalties_launch(cudaStream_t st, float* logits, int V, int rows, const SamplerRow* rows_dev, const uint32_t* ctx_dev);
// top-k order (value descending, index ascending) of each row: idx_out / val_out [rows, k]
int sampler_topk_launch(cudaStream_t st, const float* logits, int V, int rows, int k, uint32_t* idx_out, float* val_out);
// greedy / top-k(+top-p) Gumbel-max / full-vocabulary Gumbel-max -> tokens_out[rows]; topk_* hold >= row.top_k candidates per row (stride k_stride)
int sampler_draw_launch(cudaStream_t st, const float* logits, int V, int rows, const SamplerRow* rows_dev, const uint32_t* topk_idx,
                        const float* topk_val, int k_stride, const float* uniforms_dev, uint32_t* tokens_out);

}  // namespace cb
