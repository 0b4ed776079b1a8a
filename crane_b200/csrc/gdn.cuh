// Gated-Delta-Net (Qwen3.5 linear-attention layers) kernels.  See gdn.cu.
#pragma once

#include "common.cuh"

namespace cb {

struct GdnArgs {
    // in_proj output rows: [S, ldp] f32 laid out as [ q|k|v (conv_dim) | z (value_dim) | b (nv) | a (nv) ]
    const float* proj;
    int ldp;
    int S;
    int nk, nv, dk, dv, ck;      // key heads, value heads, head dims, conv kernel (4)
    const float* conv_w;         // [conv_dim, ck]
    float* conv_state;           // [conv_dim, ck] last ck inputs per channel (oldest first), updated in place
    const float* neg_exp_a;      // [nv]  -exp(A_log)
    const float* dt_bias;        // [nv]
    const float* norm_w;         // [dv]
    float eps;
    float* rec_state;            // [nv, dk, dv] f32, updated in place
    // workspaces
    float* conv_out;             // [S, conv_dim]  silu(conv)
    float* qn;                   // [S, nk, dk]    l2norm(q) / sqrt(dk)
    float* kn;                   // [S, nk, dk]    l2norm(k)
    float* gb;                   // [S, nv, 2]     (exp(g), beta)
    float* y;                    // [S, nv, dv]    recurrence output
    // outputs (one of)
    bf16* out_bf16;              // [S, value_dim]  gated-norm output as the out_proj GEMM operand (prefill)
    float* out_f32;              // [S, value_dim]  (decode: GEMV input)
    long long out_lo_off;        // split precision: element offset of out_bf16's low-order plane (0 = none)
};

int gdn_forward_launch(cudaStream_t st, const GdnArgs& a);

}  // namespace cb
