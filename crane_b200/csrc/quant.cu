// GGUF-quantised linears on the decode path: stream the QUANTISED bytes (0.56 / 0.875 / 1.06 B per weight) and do the
// dot products with dp4a on int8 activations, dequantising per block in registers -- never materialising bf16 weights.
//
// Replaces candle's `QMatMul::forward` behind `LinearLayer::Quantized` (crane-core/src/ops/linear.rs:23-48) /
// `Gguf::linear` (crane-core/src/models/hunyuan_dense/modeling.rs:37-41).  Like candle's own paths (CPU: activations
// quantised to Q8_K blocks; CUDA: mmvq with q8_1 activations) the activation vector is quantised to int8 per 32-element
// block (d = amax/127) before the integer dot; block formats follow ggml exactly (oracle/ggml_quant.py is pinned to the
// `gguf` package byte for byte).  Parity status: "unpinned" in SURVEY.md section 8c -- the oracle is f32 x . dequant(W).
//
// Kernel shape is the bf16 GEMV's (decode.cu): grid = #SMs, contiguous row block per CTA, one private cp.async ring per
// warp, weights primed before griddepcontrol.wait, fused epilogues.  Two (Q4_K, Q6_K) or four (Q8_0) lanes share one
// 256-element super-block.
#include "quant.cuh"

#include <cstring>
#include <vector>

namespace cb {

// ------------------------------------------------------------------------------------------------ host repack
void q_repack_rows(int qt, const unsigned char* src, unsigned char* dst, size_t rows, int K) {
    const size_t nsb = (size_t)rows * (K / 256);
    if (qt == QT_Q4_K) { std::memcpy(dst, src, nsb * 144); return; }
    if (qt == QT_Q6_K) {
        for (size_t i = 0; i < nsb; ++i) {
            const unsigned char* s = src + i * 210;
            unsigned char* d = dst + i * 224;
            std::memcpy(d, s, 208);              // ql (128) | qh (64) | scales (16) keep their offsets
            std::memcpy(d + 208, s + 208, 2);    // f16 d
            std::memset(d + 210, 0, 14);
        }
        return;
    }
    for (size_t i = 0; i < nsb; ++i) {           // Q8_0: 8 blocks of [f16 d | 32 x i8] -> [256 x i8 | 8 x f16 d]
        const unsigned char* s = src + i * 8 * 34;
        unsigned char* d = dst + i * 272;
        for (int b = 0; b < 8; ++b) {
            std::memcpy(d + 32 * b, s + 34 * b + 2, 32);
            std::memcpy(d + 256 + 2 * b, s + 34 * b, 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ float f16_bits_to_float(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
__device__ __forceinline__ int dp4a_s(int a, int b, int c) { return __dp4a(a, b, c); }

template <int QT> struct QTraits;
template <> struct QTraits<QT_Q4_K> { static constexpr int SB = 144, LPS = 2, DEPTH = 3; };
template <> struct QTraits<QT_Q6_K> { static constexpr int SB = 224, LPS = 2, DEPTH = 2; };
template <> struct QTraits<QT_Q8_0> { static constexpr int SB = 272, LPS = 4, DEPTH = 3; };

// Quantised activation of one sequence in shared memory.
struct XQ {
    const int* q;       // [K/4]   int8 x 4
    const float* dx;    // [K/32]  block scale
    const float* sx;    // [K/32]  dx * sum(q) over the block (for the Q4_K minima)
    const int* isum16;  // [K/16]  sum(q) over 16 elements (for the Q6_K -32 offset)
};

// Partial dot of one lane's share of a super-block (elements [e0, e0 + 256/LPS) of the row) with the activation.
template <int QT>
__device__ __forceinline__ float sb_part_dot(const unsigned char* sb, int part, int e0, const XQ& x) {
    if constexpr (QT == QT_Q4_K) {
        // half `part` = sub-blocks 4p .. 4p+3
        const uint4 hdr = *reinterpret_cast<const uint4*>(sb);
        const float d = f16_bits_to_float(hdr.x & 0xffffu), dmin = f16_bits_to_float(hdr.x >> 16);
        const unsigned char* sc8 = reinterpret_cast<const unsigned char*>(&hdr) + 4;     // 12 packed bytes
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {            // 32 qs bytes -> sub-blocks (ja, ja+1)
            const int ja = 4 * part + 2 * c;
            int sca, ma, scb, mb;
            {   // ggml get_scale_min_k4
                auto get = [&](int j, int& sc, int& m) {
                    if (j < 4) { sc = sc8[j] & 63; m = sc8[j + 4] & 63; }
                    else { sc = (sc8[j + 4] & 0xF) | ((sc8[j - 4] >> 6) << 4); m = (sc8[j + 4] >> 4) | ((sc8[j] >> 6) << 4); }
                };
                get(ja, sca, ma);
                get(ja + 1, scb, mb);
            }
            const uint4* qp = reinterpret_cast<const uint4*>(sb + 16 + 64 * part + 32 * c);
            const uint4 w0 = qp[0], w1 = qp[1];
            const int ea = e0 + 64 * c;          // first element of sub-block ja in the row
            const int4* xa = reinterpret_cast<const int4*>(x.q + (ea >> 2));
            const int4* xb = reinterpret_cast<const int4*>(x.q + ((ea + 32) >> 2));
            const int4 xa0 = xa[0], xa1 = xa[1], xb0 = xb[0], xb1 = xb[1];
            int sa = 0, sb2 = 0;
            sa = dp4a_s(w0.x & 0x0F0F0F0F, xa0.x, sa); sb2 = dp4a_s((w0.x >> 4) & 0x0F0F0F0F, xb0.x, sb2);
            sa = dp4a_s(w0.y & 0x0F0F0F0F, xa0.y, sa); sb2 = dp4a_s((w0.y >> 4) & 0x0F0F0F0F, xb0.y, sb2);
            sa = dp4a_s(w0.z & 0x0F0F0F0F, xa0.z, sa); sb2 = dp4a_s((w0.z >> 4) & 0x0F0F0F0F, xb0.z, sb2);
            sa = dp4a_s(w0.w & 0x0F0F0F0F, xa0.w, sa); sb2 = dp4a_s((w0.w >> 4) & 0x0F0F0F0F, xb0.w, sb2);
            sa = dp4a_s(w1.x & 0x0F0F0F0F, xa1.x, sa); sb2 = dp4a_s((w1.x >> 4) & 0x0F0F0F0F, xb1.x, sb2);
            sa = dp4a_s(w1.y & 0x0F0F0F0F, xa1.y, sa); sb2 = dp4a_s((w1.y >> 4) & 0x0F0F0F0F, xb1.y, sb2);
            sa = dp4a_s(w1.z & 0x0F0F0F0F, xa1.z, sa); sb2 = dp4a_s((w1.z >> 4) & 0x0F0F0F0F, xb1.z, sb2);
            sa = dp4a_s(w1.w & 0x0F0F0F0F, xa1.w, sa); sb2 = dp4a_s((w1.w >> 4) & 0x0F0F0F0F, xb1.w, sb2);
            const int ba = ea >> 5;
            acc += d * ((float)sca * x.dx[ba] * (float)sa + (float)scb * x.dx[ba + 1] * (float)sb2)
                 - dmin * ((float)ma * x.sx[ba] + (float)mb * x.sx[ba + 1]);
        }
        return acc;
    } else if constexpr (QT == QT_Q6_K) {
        // half `part` = elements [128p, 128p + 128): ql[64p..], qh[32p..], scales[8p..]
        const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 208));
        const signed char* sc = reinterpret_cast<const signed char*>(sb + 192 + 8 * part);
        float acc = 0.f;
#pragma unroll
        for (int l4 = 0; l4 < 8; ++l4) {         // 4 consecutive l per step
            const uint32_t A = *reinterpret_cast<const uint32_t*>(sb + 64 * part + 4 * l4);
            const uint32_t Bq = *reinterpret_cast<const uint32_t*>(sb + 64 * part + 32 + 4 * l4);
            const uint32_t Hh = *reinterpret_cast<const uint32_t*>(sb + 128 + 32 * part + 4 * l4);
            const int q1 = (A & 0x0F0F0F0F) | ((Hh & 0x03030303) << 4);
            const int q2 = (Bq & 0x0F0F0F0F) | (((Hh >> 2) & 0x03030303) << 4);
            const int q3 = ((A >> 4) & 0x0F0F0F0F) | (((Hh >> 4) & 0x03030303) << 4);
            const int q4 = ((Bq >> 4) & 0x0F0F0F0F) | (((Hh >> 6) & 0x03030303) << 4);
            const int e = e0 + 4 * l4;           // element of q1; q2/q3/q4 at +32/+64/+96
            const int g = l4 >> 2;               // l / 16
            const int x1 = x.q[e >> 2], x2 = x.q[(e + 32) >> 2], x3 = x.q[(e + 64) >> 2], x4 = x.q[(e + 96) >> 2];
            // (q - 32) . x = q . x - 32 * sum(x) : the -32 term is applied once per 16-element group below
            acc += d * ((float)sc[g] * x.dx[e >> 5] * (float)dp4a_s(q1, x1, 0) + (float)sc[2 + g] * x.dx[(e + 32) >> 5] * (float)dp4a_s(q2, x2, 0) +
                        (float)sc[4 + g] * x.dx[(e + 64) >> 5] * (float)dp4a_s(q3, x3, 0) + (float)sc[6 + g] * x.dx[(e + 96) >> 5] * (float)dp4a_s(q4, x4, 0));
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {            // 8 groups of 16 elements in this half
            const int e = e0 + 16 * g;
            acc -= d * (float)sc[g] * x.dx[e >> 5] * 32.f * (float)x.isum16[e >> 4];
        }
        return acc;
    } else {
        // Q8_0, quarter `part` = 64 elements = blocks 2p, 2p+1
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int blk = 2 * part + c;
            const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 256 + 2 * blk));
            const int4* wq = reinterpret_cast<const int4*>(sb + 32 * blk);
            const int e = e0 + 32 * c;
            const int4* xq = reinterpret_cast<const int4*>(x.q + (e >> 2));
            const int4 w0 = wq[0], w1 = wq[1], x0 = xq[0], x1 = xq[1];
            int s = 0;
            s = dp4a_s(w0.x, x0.x, s); s = dp4a_s(w0.y, x0.y, s); s = dp4a_s(w0.z, x0.z, s); s = dp4a_s(w0.w, x0.w, s);
            s = dp4a_s(w1.x, x1.x, s); s = dp4a_s(w1.y, x1.y, s); s = dp4a_s(w1.z, x1.z, s); s = dp4a_s(w1.w, x1.w, s);
            acc += d * x.dx[e >> 5] * (float)s;
        }
        return acc;
    }
}

constexpr int QG_WARPS = 16;
constexpr int QG_THREADS = QG_WARPS * 32;

template <int B, int QT>
__global__ void __launch_bounds__(QG_THREADS, 1)
qgemv_kernel(QGemvArgs qa) {
    using TR = QTraits<QT>;
    constexpr int SBP = 32 / TR::LPS;                  // super-blocks per warp pass
    constexpr int SLOT = SBP * TR::SB;                 // bytes per ring slot
    constexpr int CPS = SLOT / 16;                     // 16-byte chunks per slot
    const GemvArgs& a = qa.g;
    extern __shared__ __align__(1024) unsigned char qsm[];
    // layout: [rings: QG_WARPS * DEPTH * SLOT][per b: xq K | dx K/32 f32 | sx K/32 f32 | isum16 K/16 i32][acc B * rpc f32]
    const int K = a.K;
    const size_t xb = (size_t)K + (size_t)(K / 32) * 8 + (size_t)(K / 16) * 4;       // bytes of quantised activation per sequence
    unsigned char* xbase = qsm + (size_t)QG_WARPS * TR::DEPTH * SLOT;
    const int rows_per_unit = (qa.epi == GEMV_SILU_MUL) ? 2 : 1;
    const int units = a.N / rows_per_unit;
    const int upc = (units + gridDim.x - 1) / gridDim.x;
    const int rpc = upc * rows_per_unit;
    const int r0 = blockIdx.x * rpc;
    const int nrows = max(0, min(a.N, r0 + rpc) - r0);
    float* acc_s = reinterpret_cast<float*>(xbase + (size_t)B * xb);
    __shared__ float red[32];
    __shared__ float rstd_s[B];
    __shared__ float wbest_v[QG_WARPS][B];
    __shared__ int wbest_i[QG_WARPS][B];
    __shared__ int is_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t spr = (uint32_t)(K >> 8);                                          // super-blocks per row
    const uint32_t total_sb = (uint32_t)nrows * spr;
    const uint32_t total_pass = (total_sb + SBP - 1) / SBP;
    const uint32_t ppw = (total_pass + QG_WARPS - 1) / QG_WARPS;                      // passes per warp
    const uint32_t sb_begin = min(total_sb, (uint32_t)warp * ppw * SBP);
    const uint32_t sb_end = min(total_sb, ((uint32_t)warp + 1) * ppw * SBP);
    const uint32_t npass = (sb_end - sb_begin + SBP - 1) / SBP;
    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.W) + ((size_t)r0 * spr + sb_begin) * TR::SB;
    unsigned char* ring = qsm + (size_t)warp * TR::DEPTH * SLOT;
    const uint32_t ring_u32 = smem_u32(ring);

    auto issue = [&](uint32_t p) {
        if (p < npass) {
            const uint32_t nsb = min((uint32_t)SBP, sb_end - sb_begin - p * SBP);
            const uint32_t nchunk = nsb * TR::SB / 16;
            const unsigned char* src = gsrc + (size_t)p * SLOT;
            const uint32_t dst = ring_u32 + (p % TR::DEPTH) * SLOT;
            for (uint32_t c = lane; c < nchunk; c += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + (size_t)c * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int p = 0; p < TR::DEPTH; ++p) issue((uint32_t)p);
    for (int i = tid; i < B * rpc; i += QG_THREADS) acc_s[i] = 0.f;

    pdl_wait();
    pdl_launch_dependents();

    // ---- quantise the activation(s): int8 per 32-element block, RMSNorm weight folded in, 1/rms kept aside ----
#pragma unroll
    for (int b = 0; b < B; ++b) {
        unsigned char* xq = xbase + (size_t)b * xb;
        float* dxp = reinterpret_cast<float*>(xq + K);
        float* sxp = dxp + K / 32;
        int* isp = reinterpret_cast<int*>(sxp + K / 32);
        float ssq = 0.f;
        for (int blk = warp; blk < K / 32; blk += QG_WARPS) {
            float v = a.x[(size_t)b * a.ldx + blk * 32 + lane];
            if (qa.norm) { ssq += v * v; v *= a.norm_w[blk * 32 + lane]; }
            const float amax = warp_max(fabsf(v));
            const float d = amax / 127.f;
            const int q = (d > 0.f) ? __float2int_rn(v / d) : 0;
            xq[blk * 32 + lane] = (unsigned char)(signed char)q;
            int s16 = q;
            s16 += __shfl_xor_sync(0xffffffffu, s16, 1); s16 += __shfl_xor_sync(0xffffffffu, s16, 2);
            s16 += __shfl_xor_sync(0xffffffffu, s16, 4); s16 += __shfl_xor_sync(0xffffffffu, s16, 8);
            const int s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 16);
            if ((lane & 15) == 0) isp[blk * 2 + (lane >> 4)] = s16;
            if (lane == 0) { dxp[blk] = d; sxp[blk] = d * (float)s32; }
        }
        if (qa.norm) {
            const float tot = block_sum(ssq, red);
            if (tid == 0) rstd_s[b] = rsqrtf(tot / (float)K + a.eps);
        }
    }
    __syncthreads();

    // ---- stream this warp's super-blocks ----
    {
        const int part = lane % TR::LPS;
        const int sl = lane / TR::LPS;                         // super-block slot inside the pass
        const bool whole_row_pass = (spr % SBP) == 0;          // a pass never straddles rows: one reduction per pass
        for (uint32_t p = 0; p < npass; ++p) {
            asm volatile("cp.async.wait_group %0;" ::"n"(TR::DEPTH - 1) : "memory");
            __syncwarp();
            const uint32_t sbi = sb_begin + p * SBP + sl;      // super-block index inside the CTA's row block
            const bool valid = sbi < sb_end;
            const uint32_t row = valid ? sbi / spr : 0;
            const uint32_t sbk = sbi - row * spr;
            const unsigned char* sb = ring + (p % TR::DEPTH) * SLOT + sl * TR::SB;
            const int e0 = (int)sbk * 256 + part * (256 / TR::LPS);
            float part_acc[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                part_acc[b] = 0.f;
                if (valid) {
                    const unsigned char* xq = xbase + (size_t)b * xb;
                    XQ x;
                    x.q = reinterpret_cast<const int*>(xq);
                    x.dx = reinterpret_cast<const float*>(xq + K);
                    x.sx = x.dx + K / 32;
                    x.isum16 = reinterpret_cast<const int*>(x.sx + K / 32);
                    part_acc[b] = sb_part_dot<QT>(sb, part, e0, x);
                }
            }
            if (whole_row_pass) {
                const uint32_t prow = (sb_begin + p * SBP) / spr;
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float v = warp_sum(part_acc[b]);
                    if (lane == 0) atomicAdd(&acc_s[(size_t)b * rpc + prow], v);
                }
            } else if (valid) {
#pragma unroll
                for (int b = 0; b < B; ++b) atomicAdd(&acc_s[(size_t)b * rpc + row], part_acc[b]);
            }
            __syncwarp();
            issue(p + TR::DEPTH);
        }
    }
    __syncthreads();

    // ---- epilogue (same contracts as the bf16 GEMV) ----
    float bestv[B];
    int besti[B];
#pragma unroll
    for (int b = 0; b < B; ++b) { bestv[b] = -INFINITY; besti[b] = 0x7fffffff; }
    if (qa.epi == GEMV_SILU_MUL) {
        for (int u = tid; u < nrows / 2; u += QG_THREADS)
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float r = qa.norm ? rstd_s[b] : 1.f;
                a.y[(size_t)b * a.ldy + (r0 / 2 + u)] = silu_f(acc_s[(size_t)b * rpc + 2 * u] * r) * (acc_s[(size_t)b * rpc + 2 * u + 1] * r);
            }
    } else {
        for (int i = tid; i < nrows; i += QG_THREADS)
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float v = acc_s[(size_t)b * rpc + i] * (qa.norm ? rstd_s[b] : 1.f);
                float* yp = a.y + (size_t)b * a.ldy + r0 + i;
                if (qa.epi == GEMV_RESID) *yp += v; else *yp = v;
                if (qa.epi == GEMV_LOGITS_ARGMAX && v > bestv[b]) { bestv[b] = v; besti[b] = r0 + i; }
            }
    }
    if (qa.epi == GEMV_LOGITS_ARGMAX) {
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bestv[b], o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti[b], o);
                if (ov > bestv[b] || (ov == bestv[b] && oi < besti[b])) { bestv[b] = ov; besti[b] = oi; }
            }
        if (lane == 0)
#pragma unroll
            for (int b = 0; b < B; ++b) { wbest_v[warp][b] = bestv[b]; wbest_i[warp][b] = besti[b]; }
        __syncthreads();
        if (tid < B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int w = 0; w < QG_WARPS; ++w) {
                const float v = wbest_v[w][tid]; const int i = wbest_i[w][tid];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            a.part_val[tid * gridDim.x + blockIdx.x] = bv;
            a.part_idx[tid * gridDim.x + blockIdx.x] = bi;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last_s = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (is_last_s) {
            __threadfence();
            __shared__ uint32_t tok_s[B];
            if (warp < B) {
                const int b = warp;
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int c = lane; c < (int)gridDim.x; c += 32) {
                    const float v = __ldcg(a.part_val + b * gridDim.x + c);
                    const int i = __ldcg(a.part_idx + b * gridDim.x + c);
                    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) {
                    tok_s[b] = (uint32_t)bi;
                    SeqState* s = a.state + b;
                    if (a.out_tokens) a.out_tokens[(size_t)b * a.out_stride + s->step] = (uint32_t)bi;
                    if (a.advance) { s->token = (uint32_t)bi; s->kv_len += 1; s->pos[0] += 1; s->pos[1] += 1; s->pos[2] += 1; }
                    s->step += 1;
                }
            }
            if (tid == 0) *a.ticket = 0u;
            __syncthreads();
            if (a.advance && a.embed != nullptr)
                for (int b = 0; b < B; ++b) {
                    const bf16* rowp = a.embed + (size_t)tok_s[b] * a.H;
                    for (int i = tid; i < a.H; i += QG_THREADS) a.x_next[(size_t)b * a.H + i] = __bfloat162float(rowp[i]);
                }
        }
    }
}

template <int B, int QT>
static int qgemv_launch_t(cudaStream_t st, const QGemvArgs& qa, int num_sms, bool pdl) {
    using TR = QTraits<QT>;
    const GemvArgs& a = qa.g;
    const int rpu = qa.epi == GEMV_SILU_MUL ? 2 : 1;
    const int rpc = ((a.N / rpu) + num_sms - 1) / num_sms * rpu;
    const size_t xb = (size_t)a.K + (size_t)(a.K / 32) * 8 + (size_t)(a.K / 16) * 4;
    const size_t smem = (size_t)QG_WARPS * TR::DEPTH * (32 / TR::LPS) * TR::SB + (size_t)B * xb + (size_t)B * rpc * 4 + 16;
    if (smem > 227 * 1024) return -1000;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(qgemv_kernel<B, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms);
    cfg.blockDim = dim3(QG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, qgemv_kernel<B, QT>, qa);
}

template <int B>
static int qgemv_launch_b(cudaStream_t st, const QGemvArgs& qa, int num_sms, bool pdl) {
    switch (qa.qtype) {
        case QT_Q4_K: return qgemv_launch_t<B, QT_Q4_K>(st, qa, num_sms, pdl);
        case QT_Q6_K: return qgemv_launch_t<B, QT_Q6_K>(st, qa, num_sms, pdl);
        case QT_Q8_0: return qgemv_launch_t<B, QT_Q8_0>(st, qa, num_sms, pdl);
        default: return -1000;
    }
}

int qgemv_launch(cudaStream_t st, int B, const QGemvArgs& qa, int num_sms, bool pdl) {
    if ((qa.g.K % 256) != 0 || qa.g.N <= 0) return -1000;
    if (qa.epi == GEMV_SILU_MUL && (qa.g.N % 2) != 0) return -1000;
    switch (B) {
        case 1: return qgemv_launch_b<1>(st, qa, num_sms, pdl);
        case 2: return qgemv_launch_b<2>(st, qa, num_sms, pdl);
        case 4: return qgemv_launch_b<4>(st, qa, num_sms, pdl);
        default: return -1000;
    }
}

// ------------------------------------------------------------------------------------------------ dequantise to bf16
// One thread per (super-block, 8-element group) -- the prefill GEMMs consume the bf16 copy of one layer at a time.
template <int QT>
__global__ void __launch_bounds__(256)
q_dequant_kernel(const unsigned char* __restrict__ w, size_t nsb, bf16* __restrict__ out) {
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t sbi = gid >> 5;
    const int g = (int)(gid & 31);               // elements [8g, 8g+8)
    if (sbi >= nsb) return;
    const unsigned char* sb = w + sbi * QTraits<QT>::SB;
    float v[8];
    if constexpr (QT == QT_Q4_K) {
        const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb));
        const float dmin = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 2));
        const unsigned char* sc8 = sb + 4;
        const int j = g >> 2;                    // sub-block
        int sc, m;
        if (j < 4) { sc = sc8[j] & 63; m = sc8[j + 4] & 63; }
        else { sc = (sc8[j + 4] & 0xF) | ((sc8[j - 4] >> 6) << 4); m = (sc8[j + 4] >> 4) | ((sc8[j] >> 6) << 4); }
        const unsigned char* qs = sb + 16 + 32 * (j >> 1) + 8 * (g & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = (j & 1) ? (qs[i] >> 4) : (qs[i] & 0xF);
            v[i] = d * (float)sc * (float)q - dmin * (float)m;
        }
    } else if constexpr (QT == QT_Q6_K) {
        const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 208));
        const signed char* sc = reinterpret_cast<const signed char*>(sb + 192);
        const int half = g >> 4, gi = g & 15;    // 16 groups of 8 per half
        const int quarter = gi >> 2;             // which of q1..q4 (elements +0, +32, +64, +96)
        const int l0 = (gi & 3) * 8;             // l of the first element
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int l = l0 + i;
            const unsigned char ql = sb[64 * half + l + ((quarter & 1) ? 32 : 0)];
            const unsigned char qh = sb[128 + 32 * half + l];
            const int lo = (quarter & 2) ? (ql >> 4) : (ql & 0xF);
            const int hi = (qh >> (2 * quarter)) & 3;
            const int q = (lo | (hi << 4)) - 32;
            v[i] = d * (float)sc[8 * half + 2 * quarter + l / 16] * (float)q;
        }
    } else {
        const int blk = g >> 2;
        const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 256 + 2 * blk));
        const signed char* q = reinterpret_cast<const signed char*>(sb + 8 * g);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = d * (float)q[i];
    }
    uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + sbi * 256 + 8 * g) = o;
}

int q_dequant_bf16_launch(cudaStream_t st, int qt, const unsigned char* w, size_t rows, int K, bf16* out) {
    const size_t nsb = rows * (K / 256);
    const size_t threads = nsb * 32;
    const unsigned int grid = (unsigned int)((threads + 255) / 256);
    switch (qt) {
        case QT_Q4_K: q_dequant_kernel<QT_Q4_K><<<grid, 256, 0, st>>>(w, nsb, out); break;
        case QT_Q6_K: q_dequant_kernel<QT_Q6_K><<<grid, 256, 0, st>>>(w, nsb, out); break;
        case QT_Q8_0: q_dequant_kernel<QT_Q8_0><<<grid, 256, 0, st>>>(w, nsb, out); break;
        default: return -1000;
    }
    return (int)cudaGetLastError();
}

}  // namespace cb
