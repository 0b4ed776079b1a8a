// Persistent single-launch decode loop.  See decode_persistent.cu.
#pragma once

#include <algorithm>

#include "decode.cuh"

namespace cb {

constexpr int PK_WARPS_C = 24;   // warps per CTA of the persistent kernel (host-side sizing of the staging buffer)

struct PLayer {
    const bf16 *wqkv, *wo, *wgu, *wdown;
    const float *ln1, *ln2, *qn, *kn;
    bf16 *k_pool, *v_pool;
};

struct PersistArgs {
    int L, H, I, V, nh, nkv, qkv_dim, q_dim;
    float eps, scale;
    const PLayer* layers;          // device array [L]
    const bf16* lm_head;
    const float* final_norm;
    const bf16* embed;
    const float* cos_tab;          // [max_pos, 64]
    const float* sin_tab;
    const unsigned char* axis_of;  // [64]
    SeqState* state;               // [1]
    const int* block_table;
    float* x;                      // [H]   residual stream (in: embedding of the first token)
    float* qkv;                    // [qkv_dim]
    float* act;                    // [I]
    float* logits;                 // [V]
    float* part_o;                 // [nh, 8, 128]
    float* part_ml;                // [nh, 8, 2]
    float* part_val;               // [grid]
    int* part_idx;                 // [grid]
    uint32_t* out_tokens;          // greedy tokens, index state.step + t
    unsigned int* barrier;         // zeroed before the launch
    int n_steps;
    int advance;                   // 1: feed each argmax back as the next input (on-device greedy loop)
    int xs_floats;                 // staging buffer size (floats)
    unsigned long long* prof;      // optional [8] ns accumulators (CTA 0): attention, attn barrier, staging, stream, epilogue, barrier, token
};

bool decode_persistent_supported(int D, int nrep, int H, int I, int q_dim, int nkv, int num_sms);
size_t decode_persistent_smem(const PersistArgs& a, int num_sms);
int decode_persistent_launch(cudaStream_t st, const PersistArgs& a, int num_sms);

}  // namespace cb
