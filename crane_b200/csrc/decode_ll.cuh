// Persistent decode kernel with flag-carrying activations ("LL" exchange): n greedy decode steps of one sequence in ONE launch,
// no grid barriers.  See decode_ll.cu for the design.
#pragma once

#include "decode.cuh"

namespace cb {

constexpr int LL_WARPS = 8;           // 8 x 32 threads x 255 registers: nothing of the streaming loop may spill (no L1 is left beside 227 KB of shared memory)
constexpr int LL_THREADS = LL_WARPS * 32;
constexpr int LL_PART_STRIDE = 132;  // pairs per (attention item, head): 128 outputs, running max, running sum, 2 pad
constexpr int LL_MAX_LAYERS = 64;
constexpr int LL_TRACE_CAP = 4096;

struct LLLayer {
    const bf16 *wqkv, *wo, *wgu, *wdown;
    const float *ln1, *ln2, *qn, *kn;
    bf16 *k_pool, *v_pool;
};

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
    unsigned long long* amax;      // [2 (step parity), grid, 2] per-CTA (max logit, index)
    unsigned int tag_base;         // tags of this launch are tag_base + 1 .. tag_base + n_steps * (L + 2)
    int n_steps;
    int advance;                   // 1: feed each argmax back as the next input
    int l2_ahead;                  // > 0: every phase pulls its CTA's weight slab of the phase l2_ahead later into L2
    unsigned int* err;             // [1] set when a wait timed out (the launch then drains without waiting)
    unsigned long long* prof;      // optional [16] cycle accumulators of CTA 0
    unsigned long long* trace;     // optional [6][LL_TRACE_CAP][2] (event | phase << 8, globaltimer ns) of CTAs 0 / 73 / 140, warps 0 / 9, during step trace_step
    int trace_step;
};

// tags consumed by one launch (host advances its counter by this)
inline unsigned int ll_tags_per_launch(int L, int n_steps) { return (unsigned int)n_steps * (unsigned int)(L + 2); }
bool decode_ll_supported(int D, int rot_half, int nh, int nkv, int H, int I, int q_dim, int qkv_dim, int V, int num_sms);
size_t decode_ll_part_pairs(int num_sms, int nh, int nkv);
int decode_ll_launch(cudaStream_t st, const LLArgs& a, int num_sms);

}  // namespace cb
