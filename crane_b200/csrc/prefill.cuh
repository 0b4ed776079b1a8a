// Prefill / ViT kernels around the tcgen05 GEMMs: row norms, RoPE + KV-page append, flash attention,
// image-feature splice / DeepStack adds.  See prefill.cu.
#pragma once

#include "common.cuh"
#include "decode.cuh"

namespace cb {

struct RopeAppendArgs {
    const float* qkv;        // [S, nh*q_stride + 2 kv_dim] f32
    int q_stride;            // D or 2*D (per-head [query | gate])
    int rot_half;            // rotary pairs (i, i + rot_half)
    const float* q_norm_w;   // [D] (nullptr: no QK-norm)
    const float* k_norm_w;
    float eps;
    const float* cos_tab;    // [max_pos, rot_half]
    const float* sin_tab;
    const unsigned char* axis_of;   // [rot_half]
    const int* pos3;         // [3, S] rotary positions per token
    int S, start_pos;        // cache positions start_pos .. start_pos + S - 1
    const int* block_table;  // [max_pages] of this sequence
    bf16* k_pool;
    bf16* v_pool;
    int nh, nkv;
    bf16* q_out;             // [S, nh * D]
    long long q_lo_off;      // split precision: element offset of the low-order plane of q_out (0 = none)
    long long kv_lo_off;     // ... and of the K / V pools
    // quantised KV pages (engine.kv_cache = int8 / int4): k_pool / v_pool are the dequantised scratch (block_table = identity); the
    // rows are quantised here, from their f32 values, into the int pages addressed through code_bt -- the scratch receives code * scale
    int kv_bits = 0;
    unsigned char *k_codes = nullptr, *v_codes = nullptr;
    float *k_scale = nullptr, *v_scale = nullptr;
    const int* code_bt = nullptr;
};

struct FlashArgs {
    const bf16* q;   int q_stride;      // q[(row) * q_stride + head * D + d]
    const bf16* k;   const bf16* v;     // contiguous mode: k[(row) * kv_stride + kv_head * D + d]
    int kv_stride;
    const bf16* k_pool; const bf16* v_pool;   // paged mode: [n_pages, nkv, KV_PAGE, D]
    const int* block_table;
    int nh, nkv;
    bf16* out;       int o_stride;      // out[(row) * o_stride + head * D + d]
    const int* seq_start;               // [nseq] first row of each sequence (nullptr: single sequence at row 0)
    const int* seq_len;                 // [nseq] query rows per sequence
    int S;                              // single-sequence mode: query rows
    int kv_offset;                      // causal: query row i sees keys j <= kv_offset + i; keys total = kv_offset + S
    float scale;
    int nseq;
    int max_len;                        // max query rows over sequences (grid sizing)
    // split precision (x = hi + lo, both bf16): element offsets of the low-order planes, 0 = plain bf16
    long long q_lo_off, kv_lo_off, out_lo_off;
};

// quantised KV pages <-> the bf16 (hi + lo) scratch the prefill kernels read and write (see prefill.cu)
struct KvQuantArgs {
    unsigned char *k_codes, *v_codes;   // [n_pages, nkv, KV_PAGE, D * bits / 8]
    float *k_scale, *v_scale;           // [n_pages, nkv, KV_PAGE]
    const int* bt;                      // the sequence's block table into the int pages
    int nkv, D, bits;
    bf16 *sk, *sv;                      // scratch K / V in page layout, page p of the sequence at scratch page p
    long long s_lo;                     // element offset of the scratch's low-order planes (0: plain bf16)
};
int kv_quant_rows_launch(cudaStream_t st, const KvQuantArgs& a, int t0, int S);
int kv_dequant_pages_launch(cudaStream_t st, const KvQuantArgs& a, int T);

int embed_rows_launch(cudaStream_t st, const uint32_t* ids, int S, const bf16* embed, int H, float* x);
// bf16 outputs take `lo_off`: element offset (from `out`) of a second plane receiving bf16(v - bf16(v)); 0 = none
int rmsnorm_rows_launch(cudaStream_t st, const float* x, int S, int H, const float* w, float eps, bf16* out, long long lo_off = 0);
int layernorm_rows_launch(cudaStream_t st, const float* x, int rows, int W, const float* w, const float* b, float eps, bf16* out, long long lo_off = 0);
int rope_append_launch(cudaStream_t st, int D, const RopeAppendArgs& a);
int flash_prefill_launch(cudaStream_t st, int D, bool causal, bool paged, const FlashArgs& a);
int gate_mul_launch(cudaStream_t st, bf16* attn, const float* qkv, int S, int nh, int D, int q_stride, int row_width, long long lo_off = 0);
int set_rows_launch(cudaStream_t st, float* x, int H, const int* rows, int n, const float* src, bool add);
int cast_f32_bf16_launch(cudaStream_t st, const float* src, bf16* dst, size_t n, long long lo_off = 0);
int planes_to_f32_launch(cudaStream_t st, const bf16* src, long long lo_off, size_t n, float* dst);
int kv_pages_to_rows_launch(cudaStream_t st, const bf16* pool, long long lo_off, const int* bt, int nkv, int D, int T, float* out);
int kv_rows_to_pages_launch(cudaStream_t st, bf16* pool, long long lo_off, const int* bt, int nkv, int D, int T, const float* in);
int vit_pos_embed_add_launch(cudaStream_t st, float* x, int N, int Hv, const float* table, const int* idx4, const float* w4);
int vit_rope_launch(cudaStream_t st, const float* qkv, int N, int nh, int hd, const float* cos, const float* sin, bf16* out, long long lo_off = 0);

}  // namespace cb
