// GGUF-quantised linear layers (Q4_K / Q6_K / Q8_0): decode GEMV on the quantised bytes + dequantise-to-bf16 for prefill.
// See quant.cu.
#pragma once

#include "decode.cuh"

namespace cb {

// ggml_type values of the formats handled (crane-core/src/ops/linear.rs:55-78 `parse_ggml_dtype`)
enum QType : int { QT_NONE = 0, QT_Q8_0 = 8, QT_Q4_K = 12, QT_Q6_K = 14 };

// Device layout per 256 elements ("super-block"); Q4_K is ggml's own 144-byte block, the other two are repacked at load
// so every field is 16-byte aligned:
//   Q4_K : [f16 d][f16 dmin][12 B scales][128 B nibbles]                                  144 B (unchanged)
//   Q6_K : [128 B ql][64 B qh][16 x i8 scales][f16 d][14 B pad]                           224 B (ggml: 210 B)
//   Q8_0 : [256 x i8][8 x f16 d]                                                          272 B (ggml: 8 x 34 B)
__host__ __device__ inline int q_sb_bytes(int qt) { return qt == QT_Q4_K ? 144 : qt == QT_Q6_K ? 224 : qt == QT_Q8_0 ? 272 : 0; }
inline int q_src_block_bytes(int qt) { return qt == QT_Q4_K ? 144 : qt == QT_Q6_K ? 210 : qt == QT_Q8_0 ? 34 : 0; }
inline int q_src_block_elems(int qt) { return qt == QT_Q8_0 ? 32 : 256; }

// Host: repack `rows` rows of K elements from the ggml byte layout into the device layout (dst sized rows * K/256 * q_sb_bytes).
void q_repack_rows(int qt, const unsigned char* src, unsigned char* dst, size_t rows, int K);

struct QGemvArgs {
    GemvArgs g;        // W is reinterpreted as the quantised bytes; epilogue fields as for the bf16 GEMV; g.x / g.norm_w unused
    int qtype;
    int epi;           // GemvEpi
    int norm;          // (kept for the epilogue contract) xquant_launch already normalised the rows: the stored scale is 1
    const unsigned char* xq;   // activations of the B sequences, quantised by xquant_launch (xquant_bytes(B, K) bytes)
};

// Quantise B activation rows (f32, row stride ldx; RMS-normalised with norm_w first when non-null) for qgemv_launch, with the
// activation block rule candle pairs with the weight type: XQ_Q8_K for Q4_K / Q6_K, XQ_Q8_0 for Q8_0 (see xquant_kernel).
enum XQMode : int { XQ_Q8_0 = 0, XQ_Q8_K = 1 };
inline int xq_mode_for(int qt) { return qt == QT_Q8_0 ? XQ_Q8_0 : XQ_Q8_K; }
size_t xquant_bytes(int B, int K);
int xquant_launch(cudaStream_t st, int B, const float* x, int ldx, int K, const float* norm_w, float eps, int mode, unsigned char* out, bool pdl);

int qgemv_launch(cudaStream_t st, int B, const QGemvArgs& a, int num_sms, bool pdl);
// rows x K quantised -> bf16 row-major
int q_dequant_bf16_launch(cudaStream_t st, int qt, const unsigned char* w, size_t rows, int K, bf16* out);
// gather + dequantise rows of a quantised embedding table to f32: ids[rows] (prefill) or, with ids == nullptr, state[r].token
int embed_rows_q_launch(cudaStream_t st, int qt, const unsigned char* table, int H, const uint32_t* ids, const SeqState* state, int rows,
                        float* x, bool pdl);

}  // namespace cb
