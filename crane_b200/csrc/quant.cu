// GGUF-quantised linears on the decode path: stream the QUANTISED bytes (0.56 / 0.875 / 1.06 B per weight) and do the
// dot products with dp4a on int8 activations, dequantising per block in registers -- never materialising bf16 weights.
//
// Replaces candle's `QMatMul::forward` behind `LinearLayer::Quantized` (crane-core/src/ops/linear.rs:23-48) /
// `Gguf::linear` (crane-core/src/models/hunyuan_dense/modeling.rs:37-41).  Like candle's own paths (CPU: activations
// quantised to Q8_K blocks) the activation vector is quantised to int8 blocks -- Q8_K (256 elements) in front of K-quant
// weights, Q8_0 (32 elements) in front of Q8_0 weights, with candle's own rounding rules (xquant_kernel) -- before the
// integer dot; block formats follow ggml exactly (oracle/ggml_quant.py is pinned to the `gguf` package byte for byte, and
// its `qmatmul` restates the ggml integer dots this kernel reproduces).
//
// Kernel shape is the bf16 GEMV's (decode.cu): grid = #SMs, contiguous row block per CTA, one private cp.async ring per
// warp, weights primed before griddepcontrol.wait, fused epilogues.  Two (Q4_K, Q6_K) or four (Q8_0) lanes share one
// 256-element super-block.
#include "quant.cuh"

#include <cuda_fp16.h>

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
template <> struct QTraits<QT_Q4_K> { static constexpr int SB = 144; };
template <> struct QTraits<QT_Q6_K> { static constexpr int SB = 224; };
template <> struct QTraits<QT_Q8_0> { static constexpr int SB = 272; };

// Quantised activations in shared memory, B sequences back to back, each:
//   [K int8 | K/32 float2 {d, d * sum(q)} | K/16 i32 sum(q) per 16 elements]
// The int8 part is stored in 16-byte granules XOR-swizzled by the 128-element group index, so the 8 lanes of an LDS.128
// phase (2 super-blocks x 4 quarters) hit 8 different bank groups.
__device__ __forceinline__ int xq_swz(int byte_off) {
    const int g = byte_off >> 4;
    return ((g ^ ((g >> 3) & 7)) << 4) | (byte_off & 15);
}

// One lane's slice of the activations for a k-tile: 64 elements of one super-block, held in registers while the warp
// walks the rows of the tile (the weights are read once from shared memory, the activations once per tile).
//   Q4_K / Q8_0: quarter q = elements [64q, 64q + 64)                       -> blocks 2q, 2q + 1
//   Q6_K       : half h = q >> 1, lq = q & 1: elements 128h + 32t + 16lq + [0, 16), t = 0..3 (ggml's q1..q4 interleave)
template <int QT, int B>
struct XSlice {
    int4 q[B][4];
    float d[B][4];      // Q4_K: {d0, d0*sum0, d1, d1*sum1}; Q6_K: d of blocks 4h + t; Q8_0: {d0, d1, -, -}
    float s[B][4];      // Q6_K only: 32 * d_t * isum16_t
    __device__ __forceinline__ void load(const unsigned char* xbase, size_t xb, int K, int sbk_abs, int qq) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const unsigned char* xq = xbase + (size_t)b * xb;
            const float2* ds = reinterpret_cast<const float2*>(xq + K);
            const int* is16 = reinterpret_cast<const int*>(xq + K + (K / 32) * 8);
            if constexpr (QT == QT_Q6_K) {
                const int h = qq >> 1, lq = qq & 1;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e = sbk_abs * 256 + 128 * h + 32 * t + 16 * lq;
                    q[b][t] = *reinterpret_cast<const int4*>(xq + xq_swz(e));
                    d[b][t] = ds[e >> 5].x;
                    s[b][t] = 32.f * d[b][t] * (float)is16[e >> 4];
                }
            } else {
                const int e = sbk_abs * 256 + 64 * qq;
#pragma unroll
                for (int t = 0; t < 4; ++t) q[b][t] = *reinterpret_cast<const int4*>(xq + xq_swz(e + 16 * t));
                const float4 v = *reinterpret_cast<const float4*>(ds + (e >> 5));
                d[b][0] = v.x; d[b][1] = v.y; d[b][2] = v.z; d[b][3] = v.w;
            }
        }
    }
};

__device__ __forceinline__ int dot16(const int4& w0, const int4& w1, const int4& x0, const int4& x1) {
    int s = 0;
    s = dp4a_s(w0.x, x0.x, s); s = dp4a_s(w0.y, x0.y, s); s = dp4a_s(w0.z, x0.z, s); s = dp4a_s(w0.w, x0.w, s);
    s = dp4a_s(w1.x, x1.x, s); s = dp4a_s(w1.y, x1.y, s); s = dp4a_s(w1.z, x1.z, s); s = dp4a_s(w1.w, x1.w, s);
    return s;
}
__device__ __forceinline__ int dot8(const int4& w, const int4& x) {
    int s = 0;
    s = dp4a_s(w.x, x.x, s); s = dp4a_s(w.y, x.y, s); s = dp4a_s(w.z, x.z, s); s = dp4a_s(w.w, x.w, s);
    return s;
}

// acc[b] += (this lane's 64 elements of super-block `sb`) . x[b]; weight bits are unpacked once for all B sequences.
template <int QT, int B>
__device__ __forceinline__ void sb_quarter_dot(const unsigned char* sb, int qq, const XSlice<QT, B>& x, float* acc) {
    if constexpr (QT == QT_Q4_K) {
        const uint4 hdr = *reinterpret_cast<const uint4*>(sb);         // f16 d | f16 dmin | 12 packed 6-bit scales / mins
        const float d = f16_bits_to_float(hdr.x & 0xffffu), dmin = f16_bits_to_float(hdr.x >> 16);
        // ggml get_scale_min_k4 for sub-blocks ja = 2q, ja + 1
        const int sh = 16 * (qq & 1);
        const uint32_t ya = hdr.y >> sh, za = hdr.z >> sh, wa = hdr.w >> sh;
        int sca, ma, scb, mb;
        if (qq < 2) {
            sca = ya & 63; ma = za & 63; scb = (ya >> 8) & 63; mb = (za >> 8) & 63;
        } else {
            sca = (wa & 0xF) | (((ya >> 6) & 3) << 4);        ma = ((wa >> 4) & 0xF) | (((za >> 6) & 3) << 4);
            scb = ((wa >> 8) & 0xF) | (((ya >> 14) & 3) << 4); mb = ((wa >> 12) & 0xF) | (((za >> 14) & 3) << 4);
        }
        const uint4* qp = reinterpret_cast<const uint4*>(sb + 16 + 32 * qq);
        const uint4 w0 = qp[0], w1 = qp[1];
        int4 lo0, lo1, hi0, hi1;
        lo0.x = w0.x & 0x0F0F0F0F; lo0.y = w0.y & 0x0F0F0F0F; lo0.z = w0.z & 0x0F0F0F0F; lo0.w = w0.w & 0x0F0F0F0F;
        lo1.x = w1.x & 0x0F0F0F0F; lo1.y = w1.y & 0x0F0F0F0F; lo1.z = w1.z & 0x0F0F0F0F; lo1.w = w1.w & 0x0F0F0F0F;
        hi0.x = (w0.x >> 4) & 0x0F0F0F0F; hi0.y = (w0.y >> 4) & 0x0F0F0F0F; hi0.z = (w0.z >> 4) & 0x0F0F0F0F; hi0.w = (w0.w >> 4) & 0x0F0F0F0F;
        hi1.x = (w1.x >> 4) & 0x0F0F0F0F; hi1.y = (w1.y >> 4) & 0x0F0F0F0F; hi1.z = (w1.z >> 4) & 0x0F0F0F0F; hi1.w = (w1.w >> 4) & 0x0F0F0F0F;
        const float dsa = d * (float)sca, dsb = d * (float)scb, dma = dmin * (float)ma, dmb = dmin * (float)mb;
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int sa = dot16(lo0, lo1, x.q[b][0], x.q[b][1]);
            const int sb2 = dot16(hi0, hi1, x.q[b][2], x.q[b][3]);
            acc[b] += dsa * x.d[b][0] * (float)sa + dsb * x.d[b][2] * (float)sb2 - dma * x.d[b][1] - dmb * x.d[b][3];
        }
    } else if constexpr (QT == QT_Q6_K) {
        const int h = qq >> 1, lq = qq & 1;
        const float d = f16_bits_to_float(*reinterpret_cast<const unsigned short*>(sb + 208));
        const uint2 scw = *reinterpret_cast<const uint2*>(sb + 192 + 8 * h);        // 8 int8 scales of this half
        const uint4 A = *reinterpret_cast<const uint4*>(sb + 64 * h + 16 * lq);
        const uint4 Bq = *reinterpret_cast<const uint4*>(sb + 64 * h + 32 + 16 * lq);
        const uint4 Hh = *reinterpret_cast<const uint4*>(sb + 128 + 32 * h + 16 * lq);
        int4 q1, q2, q3, q4;
#define CB_Q6(c) \
        q1.c = (A.c & 0x0F0F0F0F) | ((Hh.c & 0x03030303) << 4); \
        q2.c = (Bq.c & 0x0F0F0F0F) | (((Hh.c >> 2) & 0x03030303) << 4); \
        q3.c = ((A.c >> 4) & 0x0F0F0F0F) | (((Hh.c >> 4) & 0x03030303) << 4); \
        q4.c = ((Bq.c >> 4) & 0x0F0F0F0F) | (((Hh.c >> 6) & 0x03030303) << 4);
        CB_Q6(x) CB_Q6(y) CB_Q6(z) CB_Q6(w)
#undef CB_Q6
        // scale of group t is byte 2t + lq of the half's 8 scales
        const float s1 = d * (float)(signed char)((scw.x >> (8 * lq)) & 0xff), s2 = d * (float)(signed char)((scw.x >> (16 + 8 * lq)) & 0xff);
        const float s3 = d * (float)(signed char)((scw.y >> (8 * lq)) & 0xff), s4 = d * (float)(signed char)((scw.y >> (16 + 8 * lq)) & 0xff);
#pragma unroll
        for (int b = 0; b < B; ++b)      // (q - 32) . x = q . x - 32 sum(x)
            acc[b] += s1 * (x.d[b][0] * (float)dot8(q1, x.q[b][0]) - x.s[b][0]) + s2 * (x.d[b][1] * (float)dot8(q2, x.q[b][1]) - x.s[b][1]) +
                      s3 * (x.d[b][2] * (float)dot8(q3, x.q[b][2]) - x.s[b][2]) + s4 * (x.d[b][3] * (float)dot8(q4, x.q[b][3]) - x.s[b][3]);
    } else {
        const int4* wq = reinterpret_cast<const int4*>(sb + 64 * qq);
        const int4 w0 = wq[0], w1 = wq[1], w2 = wq[2], w3 = wq[3];
        const uint32_t dd = *reinterpret_cast<const uint32_t*>(sb + 256 + 4 * qq);
        const float d0 = f16_bits_to_float(dd & 0xffffu), d1 = f16_bits_to_float(dd >> 16);
#pragma unroll
        for (int b = 0; b < B; ++b)
            acc[b] += d0 * x.d[b][0] * (float)dot16(w0, w1, x.q[b][0], x.q[b][1]) + d1 * x.d[b][2] * (float)dot16(w2, w3, x.q[b][2], x.q[b][3]);
    }
}

// Sum N per-lane values over the warp in N + log2(32/N) - 1 shuffles (recursive halving): on return every lane holds in
// v[0] the warp total of value `idx`.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane, int& idx) {
    idx = 0;
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        if (up) idx += n / 2;
    }
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

// ------------------------------------------------------------------------------------------------ activation quantiser
// One CTA per sequence: the (optionally RMS-normalised) activation row -> int8 blocks in exactly the shared-memory image the
// GEMV consumes (see xq_swz), written once to global so the 148 GEMV CTAs fetch 1.5 B/element with one bulk copy each instead
// of re-reading and re-quantising 4 B/element.  The block rule is the one candle's CPU `QMatMul::forward` applies to the
// activations before its integer dots (`LinearLayer::Quantized`, crane-core/src/ops/linear.rs:23-48):
//   XQ_Q8_K (Q4_K / Q6_K weights): BlockQ8K::from_float -- 256 elements; max = first element of largest magnitude, sign kept;
//                                  iscale = -128 / max; q = min(127, round(iscale * x)); d = 1 / iscale
//   XQ_Q8_0 (Q8_0 weights)       : BlockQ8_0::from_float -- 32 elements; d = amax / 127; q = round(x * (1 / d)); the dot uses f16(d)
// (round = half away from zero, Rust's f32::round).  RMSNorm is applied BEFORE quantising, as the reference does
// (rms_norm -> linear), so the GEMV epilogue has nothing left to scale.
constexpr int XQ_THREADS = 1024;
__host__ __device__ inline size_t xq_seq_bytes(int K) { return (size_t)K + (size_t)(K / 32) * 8 + (size_t)(K / 16) * 4; }

__global__ void __launch_bounds__(XQ_THREADS, 1)
xquant_kernel(const float* __restrict__ x, int ldx, int K, const float* __restrict__ norm_w, float eps, int B, int mode, unsigned char* __restrict__ out) {
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* xq = out + (size_t)b * xq_seq_bytes(K);
    float2* dsp = reinterpret_cast<float2*>(xq + K);
    int* isp = reinterpret_cast<int*>(xq + K + (K / 32) * 8);
    pdl_wait();
    pdl_launch_dependents();
    const float* xrow = x + (size_t)b * ldx;
    float rstd = 1.f;
    if (norm_w) {
        float ssq = 0.f;
        for (int i = tid * 4; i < K; i += XQ_THREADS * 4) {
            const float4 v = *reinterpret_cast<const float4*>(xrow + i);
            ssq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        rstd = rsqrtf(block_sum(ssq, red) / (float)K + eps);
    }
    for (int g = warp; g < K / 256; g += XQ_THREADS / 32) {           // one 256-element super-block per warp pass, 8 elements per lane
        const int e = g * 256 + lane * 8;
        float v[8];
        {
            const float4 a0 = *reinterpret_cast<const float4*>(xrow + e), a1 = *reinterpret_cast<const float4*>(xrow + e + 4);
            v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        }
        if (norm_w) {
            const float4 w0 = *reinterpret_cast<const float4*>(norm_w + e), w1 = *reinterpret_cast<const float4*>(norm_w + e + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (v[j] * rstd) * wv[j];
        }
        float am = 0.f, mx = 0.f;
        int ai = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (fabsf(v[j]) > am) { am = fabsf(v[j]); mx = v[j]; ai = j; }       // strict: the first occurrence wins
        int q[8];
        float d;
        if (mode == XQ_Q8_K) {
            int idx = lane * 8 + ai;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float oa = __shfl_xor_sync(0xffffffffu, am, o), om = __shfl_xor_sync(0xffffffffu, mx, o);
                const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (oa > am || (oa == am && oi < idx)) { am = oa; mx = om; idx = oi; }
            }
            if (am > 0.f) {
                const float iscale = -128.f / mx;
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = (int)fminf(roundf(iscale * v[j]), 127.f);
                d = 1.f / iscale;
            } else {
                d = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = 0;
            }
        } else {
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
            am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
            const float df = am / 127.f;
            const float id = df != 0.f ? 1.f / df : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = (int)roundf(v[j] * id);
            d = __half2float(__float2half_rn(df));
        }
        const uint32_t p0 = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
        const uint32_t p1 = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
        *reinterpret_cast<uint2*>(xq + xq_swz(e)) = make_uint2(p0, p1);
        const int s8 = q[0] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + q[7];
        const int s16 = s8 + __shfl_xor_sync(0xffffffffu, s8, 1);
        const int s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 2);
        if ((lane & 1) == 0) isp[e >> 4] = s16;
        if ((lane & 3) == 0) dsp[e >> 5] = make_float2(d, d * (float)s32);
    }
    if (tid == 0) reinterpret_cast<float*>(out + (size_t)B * xq_seq_bytes(K))[b] = 1.f;      // the rows are already normalised
}

size_t xquant_bytes(int B, int K) { return (size_t)B * xq_seq_bytes(K) + 16; }

int xquant_launch(cudaStream_t st, int B, const float* x, int ldx, int K, const float* norm_w, float eps, int mode, unsigned char* out, bool pdl) {
    if ((K % 256) != 0 || B < 1 || B > 4 || (ldx % 4) != 0) return -1000;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B);
    cfg.blockDim = dim3(XQ_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, xquant_kernel, x, ldx, K, norm_w, eps, B, mode, out);
}

#ifndef QG_R_DEF                    // tools/qgemv_bench.cu sweeps these
#define QG_R_DEF 4
#define QG_DEPTH_DEF 2
#define QG_W1_DEF 16
#define QG_W4_DEF 12
#endif
constexpr int QG_R = QG_R_DEF;     // rows per tile
constexpr int QG_KT = 8;           // super-blocks per tile row (32 lanes x 64 elements)
constexpr int QG_DEPTH = QG_DEPTH_DEF;   // tiles in flight per warp
__host__ __device__ constexpr int qg_warps(int B) { return B >= 4 ? QG_W4_DEF : QG_W1_DEF; }

// One CTA per SM owns a contiguous block of output rows.  Work unit = tile of QG_R rows x QG_KT super-blocks; a warp takes
// a contiguous range of tiles (k fastest), streams each through its private cp.async ring, keeps its activation slice in
// registers and one f32 accumulator per (row, sequence), and flushes them to the CTA's shared row accumulators when it
// moves on to another row group.
template <int B, int QT>
__global__ void __launch_bounds__(qg_warps(B) * 32, 1)
qgemv_kernel(QGemvArgs qa) {
    using TR = QTraits<QT>;
    constexpr int QG_MAXW = qg_warps(B);
    const int QG_WARPS = (int)(blockDim.x >> 5);       // chosen at launch so rings + activations fit in shared memory
    const int QG_THREADS = (int)blockDim.x;
    constexpr int ROWB = QG_KT * TR::SB;               // bytes of one tile row
    constexpr int SLOT = QG_R * ROWB;                  // bytes per ring slot
    const GemvArgs& a = qa.g;
    extern __shared__ __align__(1024) unsigned char qsm[];
    // layout: [rings: QG_WARPS * DEPTH * SLOT][per b: xq K | {d, d*sum} K/32 float2 | isum16 K/16 i32][1/rms B f32, 16 B][acc B * rpc f32]
    const int K = a.K;
    const size_t xb = (size_t)K + (size_t)(K / 32) * 8 + (size_t)(K / 16) * 4;       // bytes of quantised activation per sequence
    unsigned char* xbase = qsm + (size_t)QG_WARPS * QG_DEPTH * SLOT;
    const int rows_per_unit = (qa.epi == GEMV_SILU_MUL) ? 2 : 1;
    const int units = a.N / rows_per_unit;
    const int upc = (units + gridDim.x - 1) / gridDim.x;
    const int rpc = upc * rows_per_unit;
    const int r0 = blockIdx.x * rpc;
    const int nrows = max(0, min(a.N, r0 + rpc) - r0);
    float* acc_s = reinterpret_cast<float*>(xbase + (size_t)B * xb + 16);
    __shared__ uint64_t xbar;
    __shared__ float wbest_v[QG_MAXW][B];
    __shared__ int wbest_i[QG_MAXW][B];
    __shared__ int is_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t spr = (uint32_t)(K >> 8);                                          // super-blocks per row
    const uint32_t nkt = (spr + QG_KT - 1) / QG_KT;                                   // k-tiles per row
    // rows per tile: QG_R when that leaves every warp at least two tiles, fewer for short row blocks (q/k/v/o projections)
    uint32_t R = QG_R;
    while (R > 1 && (((uint32_t)nrows + R - 1) / R) * nkt < 2u * (uint32_t)QG_WARPS) R >>= 1;
    const uint32_t nrg = ((uint32_t)nrows + R - 1) / R;                               // row groups
    const uint32_t ntile = nrg * nkt;
    const uint32_t tpw = (ntile + QG_WARPS - 1) / QG_WARPS;
    const uint32_t t_begin = min(ntile, (uint32_t)warp * tpw), t_end = min(ntile, ((uint32_t)warp + 1) * tpw);
    const uint32_t nt = t_end - t_begin;
    const unsigned char* wbase = reinterpret_cast<const unsigned char*>(a.W) + (size_t)r0 * spr * TR::SB;
    unsigned char* ring = qsm + (size_t)warp * QG_DEPTH * SLOT;
    const uint32_t ring_u32 = smem_u32(ring);

    auto issue = [&](uint32_t p) {
        if (p < nt) {
            const uint32_t t = t_begin + p, rg = t / nkt, kt = t - rg * nkt;
            const uint32_t nr = min(R, (uint32_t)nrows - rg * R);
            const uint32_t cpr = min((uint32_t)QG_KT, spr - kt * QG_KT) * (TR::SB / 16);      // 16-byte chunks per tile row
            const uint32_t dst = ring_u32 + (p % QG_DEPTH) * SLOT;
            for (uint32_t r = 0; r < nr; ++r) {
                const unsigned char* src = wbase + ((size_t)(rg * R + r) * spr + (size_t)kt * QG_KT) * TR::SB;
                for (uint32_t c = lane; c < cpr; c += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * ROWB + c * 16), "l"(src + (size_t)c * 16) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int p = 0; p < QG_DEPTH; ++p) issue((uint32_t)p);
    for (int i = tid; i < B * rpc; i += QG_THREADS) acc_s[i] = 0.f;
    if (tid == 0) { mbar_init(&xbar, 1); fence_barrier_init(); }
    __syncthreads();

    pdl_wait();
    pdl_launch_dependents();

    // ---- fetch the activations quantised by xquant_kernel: one bulk copy of the ready-made shared-memory image ----
    const uint32_t xbytes = (uint32_t)(B * xb) + 16u;
    if (tid == 0) {
        mbar_arrive_expect_tx(&xbar, xbytes);
        bulk_load(xbase, qa.xq, xbytes, &xbar);
    }
    mbar_wait(&xbar, 0);
    const float* rstd_s = reinterpret_cast<const float*>(xbase + (size_t)B * xb);

    // ---- stream this warp's tiles ----
    {
        const int sbk = lane >> 2, qq = lane & 3;              // this lane's super-block within the tile row and its quarter
        XSlice<QT, B> xs;
        float acc[QG_R * B];
#pragma unroll
        for (int i = 0; i < QG_R * B; ++i) acc[i] = 0.f;
        uint32_t cur_rg = 0xffffffffu, cur_kt = 0xffffffffu;
        auto flush = [&]() {
            int idx;
            warp_reduce_scatter<QG_R * B>(acc, lane, idx);
            const int r = idx / B, b = idx % B;
            const uint32_t row = cur_rg * R + r;
            if ((lane & (32 / (QG_R * B) - 1)) == 0 && (uint32_t)r < R && row < (uint32_t)nrows) atomicAdd(&acc_s[(size_t)b * rpc + row], acc[0]);
#pragma unroll
            for (int i = 0; i < QG_R * B; ++i) acc[i] = 0.f;
        };
        for (uint32_t p = 0; p < nt; ++p) {
            asm volatile("cp.async.wait_group %0;" ::"n"(QG_DEPTH - 1) : "memory");
            __syncwarp();
            const uint32_t t = t_begin + p, rg = t / nkt, kt = t - rg * nkt;
            if (rg != cur_rg) {
                if (cur_rg != 0xffffffffu) flush();
                cur_rg = rg;
            }
            const bool active = kt * QG_KT + sbk < spr;
            if (kt != cur_kt) {
                if (active) xs.load(xbase, xb, K, (int)(kt * QG_KT) + sbk, qq);
                cur_kt = kt;
            }
            const uint32_t nr = min(R, (uint32_t)nrows - rg * R);
            const unsigned char* sb = ring + (p % QG_DEPTH) * SLOT + sbk * TR::SB;
            if (active) {
#pragma unroll
                for (int r = 0; r < QG_R; ++r)
                    if ((uint32_t)r < nr) sb_quarter_dot<QT, B>(sb + r * ROWB, qq, xs, acc + r * B);
            }
            __syncwarp();
            issue(p + QG_DEPTH);
        }
        if (cur_rg != 0xffffffffu) flush();
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
                if ((unsigned)bi >= (unsigned)a.N) bi = 0;       // an all-NaN row has no maximum: never gather from an out-of-range id
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
    const size_t fixed = (size_t)B * xb + 16 + (size_t)B * rpc * 4 + 16, slot = (size_t)QG_DEPTH * QG_R * QG_KT * TR::SB;
    int QG_WARPS = qg_warps(B);
    while (QG_WARPS > 4 && fixed + QG_WARPS * slot > 226 * 1024) --QG_WARPS;
    const int QG_THREADS = QG_WARPS * 32;
    const size_t smem = fixed + QG_WARPS * slot;
    if (smem > 226 * 1024) return -1000;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(qgemv_kernel<B, QT>, smem, seen)) return e;
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

// ------------------------------------------------------------------------------------------------ dequantise
// elements [8g, 8g + 8) of one 256-element super-block (device layout) -> f32, exactly d * sc * q - dmin * m / d * sc * (q - 32) / d * q
template <int QT>
__device__ __forceinline__ void dequant8(const unsigned char* sb, int g, float (&v)[8]) {
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
}

// One thread per (super-block, 8-element group): a whole quantised matrix -> bf16
template <int QT>
__global__ void __launch_bounds__(256)
q_dequant_kernel(const unsigned char* __restrict__ w, size_t nsb, bf16* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t sbi = gid >> 5;
    const int g = (int)(gid & 31);               // elements [8g, 8g+8)
    if (sbi >= nsb) return;
    float v[8];
    dequant8<QT>(w + sbi * QTraits<QT>::SB, g, v);
    uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + sbi * 256 + 8 * g) = o;
}

int q_dequant_bf16_launch(cudaStream_t st, int qt, const unsigned char* w, size_t rows, int K, bf16* out) {
    const size_t nsb = rows * (K / 256);
    const size_t threads = nsb * 32;
    const unsigned int grid = (unsigned int)((threads + 255) / 256);
    switch (qt) {
        case QT_Q4_K: return launch_k(q_dequant_kernel<QT_Q4_K>, dim3(grid), dim3(256), 0, st, prefill_pdl(), w, nsb, out);
        case QT_Q6_K: return launch_k(q_dequant_kernel<QT_Q6_K>, dim3(grid), dim3(256), 0, st, prefill_pdl(), w, nsb, out);
        case QT_Q8_0: return launch_k(q_dequant_kernel<QT_Q8_0>, dim3(grid), dim3(256), 0, st, prefill_pdl(), w, nsb, out);
        default: return -1000;
    }
}

// Quantised embedding: the table stays in its ggml blocks and only the gathered rows are dequantised, to f32
// (`QuantizedEmbedding::forward`, crane-core/src/models/modules/embedding.rs:31-105: index_select of Q-blocks + dequantize).
// Row ids come from `ids` (prefill) or, when ids == nullptr, from state[row].token (the decode step's input).
template <int QT>
__global__ void __launch_bounds__(256)
embed_rows_q_kernel(const unsigned char* __restrict__ table, int H, const uint32_t* __restrict__ ids, const SeqState* __restrict__ state,
                    float* __restrict__ x) {
    pdl_wait();
    pdl_launch_dependents();
    const int r = blockIdx.x;
    const uint32_t tok = ids ? ids[r] : state[r].token;
    const unsigned char* row = table + (size_t)tok * (H / 256) * QTraits<QT>::SB;
    for (int it = threadIdx.x; it < H / 8; it += 256) {
        float v[8];
        dequant8<QT>(row + (size_t)(it >> 5) * QTraits<QT>::SB, it & 31, v);
        float4* o = reinterpret_cast<float4*>(x + (size_t)r * H + it * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

int embed_rows_q_launch(cudaStream_t st, int qt, const unsigned char* table, int H, const uint32_t* ids, const SeqState* state, int rows,
                        float* x, bool pdl) {
    if (rows <= 0 || (H % 256) != 0) return -1000;
    switch (qt) {
        case QT_Q4_K: return launch_k(embed_rows_q_kernel<QT_Q4_K>, dim3(rows), dim3(256), 0, st, pdl, table, H, ids, state, x);
        case QT_Q6_K: return launch_k(embed_rows_q_kernel<QT_Q6_K>, dim3(rows), dim3(256), 0, st, pdl, table, H, ids, state, x);
        case QT_Q8_0: return launch_k(embed_rows_q_kernel<QT_Q8_0>, dim3(rows), dim3(256), 0, st, pdl, table, H, ids, state, x);
        default: return -1000;
    }
}

}  // namespace cb
