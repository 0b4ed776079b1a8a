// Persistent decode kernel with flag-carrying activations ("LL" exchange): n greedy decode steps of one sequence in ONE launch,
// no grid barriers.  See decode_ll.cu for the design.
#pragma once

#include "decode.cuh"

namespace cb {

constexpr int LL_WARPS = 16;
constexpr int LL_THREADS = LL_WARPS * 32;
constexpr int LL_MAX_NCC = 4;        // 256-column chunks of a weight row one warp may own (activation slice held in registers)
constexpr int LL_PART_STRIDE = 132;  // pairs per (attention item, head): 128 outputs, running max, running sum, 2 pad
constexpr int LL_MAX_LAYERS = 64;

struct LLLayer {
    const bf16 *wqkv, *wo, *wgu, *wdown;
    const float *ln1, *ln2, *qn, *kn;
    bf16 *k_pool, *v_pool;
};

// How the 16 warps of a CTA share a [rows, K] row block: G column groups x RL row lanes; a warp owns `ncc` 256-column chunks of
// every RL-th row, so the activation columns it needs never change during a phase (8 * ncc registers per lane).
struct LLGeom { int G, RL, ncc; };
__host__ __device__ inline bool ll_geom(int K, LLGeom& g) {
    if (K <= 0 || (K & 255)) return false;
    const int cpr = K >> 8;
    int G = 1;
    while (G < LL_WARPS && (cpr % (G * 2)) == 0) G *= 2;
    g.G = G; g.RL = LL_WARPS / G; g.ncc = cpr / G;
    return g.ncc <= LL_MAX_NCC;
}

struct LLArgs {
    int L, H, I, V, nh, nkv, qkv_dim, q_dim;
    float eps, scale;
    LLLayer layers[LL_MAX_LAYERS]; // by value: the whole table sits in the kernel's constant parameter bank
    const bf16* lm_head;
    const float* final_norm;
    const bf16* embed;
    const float* cos_tab;          // [max_pos, 64]
    const float* sin_tab;
    const unsigned char* axis_of;  // [64]
    SeqState* state;               // [1]
    const int* block_table;        // [n_slots, max_pages]
    int max_pages;
    long long kv_lo_off;           // split precision: element offset of the low-order KV planes (0 = plain bf16 pages)
    float* x_io;                   // [H] residual-stream input of the first step; next step's embedding on exit (advance)
    float* logits;                 // [V]
    uint32_t* out_tokens;          // greedy tokens, index state.step + s
    // (value, tag) exchange buffers, 8 bytes per element, zero-initialised once
    unsigned long long *xa, *xb;   // [H] residual before the attention block / before the MLP block
    unsigned long long* qkv;       // [qkv_dim]
    unsigned long long* att;       // [q_dim]
    unsigned long long* act;       // [I]
    unsigned long long* part;      // [grid, nrep, LL_PART_STRIDE] attention partials
    unsigned long long* amax;      // [grid, 2] per-CTA (max logit, index)
    unsigned int tag_base;         // tags of this launch are tag_base + 1 .. tag_base + n_steps * (L + 2)
    int n_steps;
    int advance;                   // 1: feed each argmax back as the next input
    unsigned int* err;             // [1] set when a wait timed out (the launch then drains without waiting)
    unsigned long long* prof;      // optional [16] ns accumulators of CTA 0
};

// tags consumed by one launch (host advances its counter by this)
inline unsigned int ll_tags_per_launch(int L, int n_steps) { return (unsigned int)n_steps * (unsigned int)(L + 2); }
bool decode_ll_supported(int D, int rot_half, int nh, int nkv, int H, int I, int q_dim, int qkv_dim, int V, int num_sms);
size_t decode_ll_part_pairs(int num_sms, int nh, int nkv);
int decode_ll_launch(cudaStream_t st, const LLArgs& a, int num_sms);

}  // namespace cb
