// Argument blocks of the Gated-Delta-Net kernels (gdn.cu, gdn_chunk.cu).  Plain structs, no device code: the chunk kernels are
// also compiled by the host-side emulation test (tests/emu), which includes this file without the CUDA runtime.
#pragma once

#include <cuda_bf16.h>
#include <cstddef>
#include <cstdint>

namespace cb {

typedef __nv_bfloat16 bf16;

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
    // chunkwise prefill (gdn_chunk.cu); both null = token-by-token recurrence
    float* glog;                 // [S, nv]  g itself (the log of gb's decay), written by the prep kernel when non-null
    void* chunk_ws;              // >= gdn_chunk_ws_bytes(S, nv, dk, dv) bytes of scratch
};

constexpr int GDN_CHUNK = 64;    // tokens per chunk of the chunkwise form

// Scratch of the chunkwise path.  Every bf16 operand is a (hi, lo) plane pair -- hi = bf16(x), lo = bf16(x - hi) -- laid out
// plane-major inside its (chunk, head) block, exactly as the consuming kernel copies it into shared memory.
struct GdnChunkWs {
    bf16* w;     // [n_chunks, nv, 2, 64, dk]   W  = (I + A)^-1 diag(beta e^b) K
    bf16* kt;    // [n_chunks, nv, 2, 64, dk]   K~ = diag(e^{b_C - b}) K
    bf16* qt;    // [n_chunks, nv, 2, 64, dk]   Q~ = diag(e^b) Q
    bf16* p;     // [n_chunks, nv, 2, 64, 64]   P  = tril(e^{b_i - b_j} q_i.k_j)
    float* ut;   // [n_chunks, nv, dv, 64]      U^T, U = (I + A)^-1 diag(beta) V
    float* gc;   // [n_chunks, nv]              e^{b_C}
    bf16* st;    // [n_chunks, nv, 2, dv, dk]   S^T at the start of each chunk
    bf16* dt;    // [n_chunks, nv, 2, dv, 64]   D^T, D = U - W S0
};

inline size_t gdn_ws_align(size_t b) { return (b + 255) / 256 * 256; }
inline int gdn_n_chunks(int S) { return (S + GDN_CHUNK - 1) / GDN_CHUNK; }
inline size_t gdn_chunk_ws_bytes(int S, int nv, int dk, int dv) {
    const size_t n = (size_t)gdn_n_chunks(S) * nv;
    return 3 * gdn_ws_align(n * 2 * GDN_CHUNK * dk * 2) + gdn_ws_align(n * 2 * GDN_CHUNK * GDN_CHUNK * 2) +
           gdn_ws_align(n * dv * GDN_CHUNK * 4) + gdn_ws_align(n * 4) + gdn_ws_align(n * 2 * dv * dk * 2) +
           gdn_ws_align(n * 2 * dv * GDN_CHUNK * 2);
}
inline GdnChunkWs gdn_chunk_ws_carve(void* base, int S, int nv, int dk, int dv) {
    const size_t n = (size_t)gdn_n_chunks(S) * nv;
    unsigned char* p = static_cast<unsigned char*>(base);
    GdnChunkWs w;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += gdn_ws_align(bytes); return r; };
    w.w = reinterpret_cast<bf16*>(take(n * 2 * GDN_CHUNK * dk * 2));
    w.kt = reinterpret_cast<bf16*>(take(n * 2 * GDN_CHUNK * dk * 2));
    w.qt = reinterpret_cast<bf16*>(take(n * 2 * GDN_CHUNK * dk * 2));
    w.p = reinterpret_cast<bf16*>(take(n * 2 * GDN_CHUNK * GDN_CHUNK * 2));
    w.ut = reinterpret_cast<float*>(take(n * dv * GDN_CHUNK * 4));
    w.gc = reinterpret_cast<float*>(take(n * 4));
    w.st = reinterpret_cast<bf16*>(take(n * 2 * dv * dk * 2));
    w.dt = reinterpret_cast<bf16*>(take(n * 2 * dv * GDN_CHUNK * 2));
    return w;
}

}  // namespace cb
