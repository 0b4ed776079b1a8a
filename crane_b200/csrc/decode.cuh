// Decode-path kernels (S = 1 per sequence): bandwidth-bound weight streaming + paged-KV GQA attention.
//
// Replaces, per decoded token, what the reference runs as ~10 Candle ops + cuBLAS gemv calls per layer
// (crane-core/src/models/qwen3/modeling.rs:698-716 -> :307-533, :608-642, :1024-1035) and the two-phase
// `gpu_argmax` (crane-core/kernels/cuda/fused_ops.cu:251-382) with 4 fused kernels per layer:
//   gemv<NORM, STORE>      input RMSNorm + merged QKV projection
//   attn_decode            QK-norm + RoPE/MRoPE + KV-page append + split-KV GQA attention + split merge
//   gemv<RESID>            O-projection + residual add
//   gemv<NORM, SILU_MUL>   post-attention RMSNorm + merged gate/up projection + SiLU*up
//   gemv<RESID>            down projection + residual add
// and  gemv<NORM, LOGITS_ARGMAX>  final RMSNorm + lm_head + greedy argmax + next-token embedding gather.
// All kernels are chained with programmatic dependent launch: each one starts streaming its (immutable)
// weights before `griddepcontrol.wait`, so the HBM pipe stays busy across kernel boundaries.
#pragma once

#include "common.cuh"

namespace cb {

constexpr int KV_PAGE = 64;        // tokens per KV page
constexpr int ATTN_NSPLIT = 8;     // split-KV factor == thread-block cluster size (static grid: the step lives in a CUDA graph)
constexpr int MAX_BATCH = 8;

// Device-resident per-sequence decode state (lets a whole decode step replay as a CUDA graph).
struct SeqState {
    int kv_len;        // tokens already in the cache (the step appends position kv_len)
    int pos[3];        // MRoPE (t, h, w) position of the token being decoded
    uint32_t token;    // input token of the step
    int step;          // decode steps taken since the state was armed
    int slot;          // sequence slot: its page table is block_table + slot * max_pages
    int pad;
};

enum GemvEpi : int { GEMV_STORE = 0, GEMV_RESID = 1, GEMV_SILU_MUL = 2, GEMV_LOGITS_ARGMAX = 3 };

struct GemvArgs {
    const bf16* W;          // [N, K] row-major (gate/up rows interleaved for GEMV_SILU_MUL)
    int N, K;
    const float* x;         // [B, ldx] f32
    int ldx;
    const float* norm_w;    // [K] RMSNorm weight (NORM) else nullptr
    float eps;
    float* y;               // [B, ldy] f32 (STORE: = ; RESID: += ; SILU_MUL: [B, N/2]; LOGITS_ARGMAX: logits)
    int ldy;
    // --- GEMV_LOGITS_ARGMAX only ---
    float* part_val;        // [B, gridDim.x]
    int* part_idx;          // [B, gridDim.x]
    unsigned int* ticket;   // 1 counter
    SeqState* state;        // [B]
    uint32_t* out_tokens;   // [B, out_stride] greedy tokens, index state.step
    int out_stride;
    const bf16* embed;      // [V, H] embedding table for the next-step gather
    float* x_next;          // [B, H] residual-stream input of the next step
    int H;
    int advance;            // 0: report the argmax only; 1: feed it back (token, next embedding, kv_len/pos + 1);
                            // 2: advance kv_len/pos and gather embed[state.token] WITHOUT replacing the token (caller-chosen input)
    float* norm_out;        // optional [B, K]: the RMS-normalised input (x * w * rstd), written by CTA 0 (NORM kernels)
    const uint32_t* force_tokens;  // optional [out_stride]: teacher forcing -- token fed back at step s is force_tokens[s] (argmax still reported)
};

struct AttnDecArgs {
    const float* qkv;        // [B, nh*q_stride + 2*kv_dim] f32 (pre-norm, pre-rope); head h's query at h*q_stride
    int q_stride;            // D, or 2*D when q_proj emits per-head [query | gate] (Qwen3.5, qwen3_5/modeling.rs:428-455)
    int gated;               // 1: out *= sigmoid(gate)
    int rot_half;            // rotary pairs (i, i + rot_half), i < rot_half; D/2 for full rotary, 32 for Qwen3.5's 64-of-256
    const float* q_norm_w;   // [D]
    const float* k_norm_w;   // [D]
    float eps;
    const float* cos_tab;    // [max_pos, rot_half]
    const float* sin_tab;
    const unsigned char* axis_of;  // [rot_half] MRoPE axis per rotary column (all 0 for 1-D RoPE)
    const SeqState* state;   // [B]
    const int* block_table;  // [n_slots, max_pages], row = SeqState.slot
    int max_pages;
    bf16* k_pool;            // this layer's pages: [n_pages, nkv, KV_PAGE, D]
    bf16* v_pool;
    int nh, nkv;
    float scale;
    float* out;              // [B, nh*D] f32
    long long kv_lo_off;     // split precision: element offset of the pools' low-order planes (0 = plain bf16 pages)
    // quantised KV pages (QuantKvCache, qwen3_5/kv_cache.rs:209-342): kv_bits 8 / 4 -> k_pool / v_pool are unused, the pages are
    // u8 codes [n_pages, nkv, KV_PAGE, D * bits / 8] + f32 scales [n_pages, nkv, KV_PAGE]; dequantised while the tile is staged
    int kv_bits = 0;
    int kv_split = 0;        // kv_bits != 0: stage the dequantised value as hi + lo bf16 planes (the parity mode) or hi only
    unsigned char *k_codes = nullptr, *v_codes = nullptr;
    float *k_scale = nullptr, *v_scale = nullptr;
    int groups = 1;          // set by attn_decode_launch: CTA clusters per KV head (query group width / kernel sub-group width)
};

int gemv_max_group(int K, int N, int num_sms);
int gemv_launch(cudaStream_t st, int B, int epi, bool norm, const GemvArgs& a, int num_sms, bool pdl);
int attn_decode_launch(cudaStream_t st, int B, int D, const AttnDecArgs& a, bool pdl);
int embed_decode_launch(cudaStream_t st, int B, const bf16* embed, int H, const SeqState* state, float* x, bool pdl);

}  // namespace cb
