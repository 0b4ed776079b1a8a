// Decode-path kernels.  See decode.cuh for the design.
#include "decode.cuh"

#include <cooperative_groups.h>

namespace cb {

// =====================================================================================
// Weight-streaming GEMV:  y[b, n] = epilogue( sum_k W[n, k] * x[b, k] )          (B <= 4 sequences)
//
// grid = #SMs; CTA c owns the contiguous row block [c*rpc, (c+1)*rpc) of W -- one contiguous byte slab -- and splits it
// into GV_WARPS contiguous sub-slabs, one per warp.  Every warp runs its own private, barrier-free pipeline:
// GV_DEPTH x 2 KB shared-memory slots filled by cp.async (LDGSTS, 16 B per lane, L1-bypassing) and drained with 16-byte
// LDS (conflict-free) into f32 FMAs -- 80 KB of weight bytes in flight per SM, no cross-warp synchronisation in the
// streaming loop.  Weights do not depend on the previous kernel, so the pipelines are primed BEFORE griddepcontrol.wait:
// with programmatic dependent launch the next kernel's CTAs sit next to this kernel's on the SM (2 x 104 KB shared
// memory) and the HBM pipe stays busy across the kernel boundary.  Activations are staged once per CTA as f32 (RMSNorm
// folded in).  Rows are flushed to a shared accumulator (shuffle reduce + one shared atomic per row change per warp).
// Epilogue (all threads): store / residual add / SiLU*up on interleaved rows / logits + CTA argmax.
// =====================================================================================
#ifndef GV_WARPS
#define GV_WARPS 32
#endif
#ifndef GV_DEPTH
#define GV_DEPTH 2
#endif
#ifndef GV_MINB                    // CTAs per SM the GEMV is compiled for (2 lets the next launch's CTA prime its ring early)
#define GV_MINB 1
#endif
constexpr int GV_SEG_CHUNKS = 4;                       // 512-byte chunks per pipeline slot (2 KB)
constexpr int GV_SEG_BYTES = GV_SEG_CHUNKS * 512;
constexpr int GV_THREADS = GV_WARPS * 32;
__host__ __device__ constexpr int gv_depth(int B) { return B == 1 ? GV_DEPTH : 1; }   // batched activations take the ring's shared memory

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_g() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_g() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int B, int EPI, bool NORM>
__global__ void __launch_bounds__(GV_THREADS, GV_MINB)
gemv_kernel(GemvArgs a) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    // layout: [rings: GV_WARPS * GV_DEPTH * 2 KB][xs: B * K f32][acc: B * rpc f32]
    const int K8 = a.K >> 3;
    constexpr int ROWS_PER_UNIT = (EPI == GEMV_SILU_MUL) ? 2 : 1;
    const int units = a.N / ROWS_PER_UNIT;
    const int upc = (units + gridDim.x - 1) / gridDim.x;
    const int rpc = upc * ROWS_PER_UNIT;                               // rows per CTA (capacity)
    const int r0 = blockIdx.x * rpc;
    const int r1 = min(a.N, r0 + rpc);
    const int nrows = max(0, r1 - r0);
    constexpr int DEPTH = gv_depth(B);
    float4* xs = reinterpret_cast<float4*>(gsm + GV_WARPS * DEPTH * GV_SEG_BYTES);
    float* acc_s = reinterpret_cast<float*>(xs) + (size_t)B * a.K;
    __shared__ float red[32];
    __shared__ float rstd_s[B];
    __shared__ float wbest_v[GV_WARPS][B];
    __shared__ int wbest_i[GV_WARPS][B];
    __shared__ int is_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 32-bit chunk bookkeeping (chunk = 512 B = 256 elements; never straddles a row since K % 256 == 0)
    const uint32_t cpr = (uint32_t)(a.K >> 8);                                     // chunks per row
    const uint32_t total_chunks = (uint32_t)nrows * cpr;
    const uint32_t total_segs = (total_chunks + GV_SEG_CHUNKS - 1) / GV_SEG_CHUNKS;
    const uint32_t spw = (total_segs + GV_WARPS - 1) / GV_WARPS;                   // segments per warp
    const uint32_t c_begin = min(total_chunks, (uint32_t)warp * spw * GV_SEG_CHUNKS);
    const uint32_t c_end = min(total_chunks, ((uint32_t)warp + 1) * spw * GV_SEG_CHUNKS);
    const uint32_t nseg = (c_end - c_begin + GV_SEG_CHUNKS - 1) / GV_SEG_CHUNKS;
    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.W) + (size_t)r0 * a.K * 2 + (size_t)c_begin * 512 + lane * 16;
    unsigned char* ring = gsm + (size_t)warp * DEPTH * GV_SEG_BYTES + lane * 16;
    const uint32_t ring_u32 = smem_u32(ring);

    auto issue = [&](uint32_t g) {                      // segment g of this warp -> slot g % GV_DEPTH
        if (g < nseg) {
            const uint32_t cb = g * GV_SEG_CHUNKS;
            const uint32_t slot = g % DEPTH;
#pragma unroll
            for (int c = 0; c < GV_SEG_CHUNKS; ++c)
                if (c_begin + cb + c < c_end)
                    cp_async16_cg(ring_u32 + slot * GV_SEG_BYTES + c * 512, gsrc + (size_t)(cb + c) * 512);
        }
        cp_async_commit_g();
    };
    // prime the pipeline: weights do not depend on the previous kernel
#pragma unroll
    for (int g = 0; g < DEPTH; ++g) issue((uint32_t)g);
    for (int i = tid; i < B * rpc; i += GV_THREADS) acc_s[i] = 0.f;
    // RMSNorm weights are immutable (and cold in every cache): fetch this thread's first chunk before the dependency wait
    float4 nw0 = make_float4(0.f, 0.f, 0.f, 0.f), nw1 = nw0;
    if (NORM && tid < K8) { nw0 = reinterpret_cast<const float4*>(a.norm_w)[2 * tid]; nw1 = reinterpret_cast<const float4*>(a.norm_w)[2 * tid + 1]; }

    pdl_wait();
    pdl_launch_dependents();

    // --- stage activations (f32) and, if NORM, fold the RMSNorm weight in and get 1/rms ---
#pragma unroll
    for (int b = 0; b < B; ++b) {
        float ssq = 0.f;
        for (int i = tid; i < K8; i += GV_THREADS) {
            const float4* xp = reinterpret_cast<const float4*>(a.x + (size_t)b * a.ldx) + 2 * i;
            float4 lo = xp[0], hi = xp[1];
            if (NORM) {
                ssq += lo.x * lo.x + lo.y * lo.y + lo.z * lo.z + lo.w * lo.w + hi.x * hi.x + hi.y * hi.y + hi.z * hi.z + hi.w * hi.w;
                const float4 w0 = (i == tid) ? nw0 : reinterpret_cast<const float4*>(a.norm_w)[2 * i];
                const float4 w1 = (i == tid) ? nw1 : reinterpret_cast<const float4*>(a.norm_w)[2 * i + 1];
                lo.x *= w0.x; lo.y *= w0.y; lo.z *= w0.z; lo.w *= w0.w;
                hi.x *= w1.x; hi.y *= w1.y; hi.z *= w1.z; hi.w *= w1.w;
            }
            xs[(size_t)(b * 2 + 0) * K8 + i] = lo;
            xs[(size_t)(b * 2 + 1) * K8 + i] = hi;
        }
        if (NORM) {
            const float tot = block_sum(ssq, red);
            if (tid == 0) rstd_s[b] = rsqrtf(tot / (float)a.K + a.eps);
        }
    }
    __syncthreads();
    if (NORM && a.norm_out != nullptr && blockIdx.x == 0) {      // post-norm hidden state for callers that need it (TTS code predictor input)
#pragma unroll
        for (int b = 0; b < B; ++b)
            for (int i = tid; i < K8; i += GV_THREADS) {
                const float r = rstd_s[b];
                const float4 lo = xs[(size_t)(b * 2 + 0) * K8 + i], hi = xs[(size_t)(b * 2 + 1) * K8 + i];
                float4* o = reinterpret_cast<float4*>(a.norm_out + (size_t)b * a.K) + 2 * i;
                o[0] = make_float4(lo.x * r, lo.y * r, lo.z * r, lo.w * r);
                o[1] = make_float4(hi.x * r, hi.y * r, hi.z * r, hi.w * r);
            }
    }

    {
        float accum[B], accum2[B];
#pragma unroll
        for (int b = 0; b < B; ++b) { accum[b] = 0.f; accum2[b] = 0.f; }
        int cur_row = -1;
        auto flush = [&]() {
            if (cur_row >= 0) {
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float v = warp_sum(accum[b] + accum2[b]);
                    if (lane == 0) atomicAdd(&acc_s[(size_t)b * rpc + cur_row], v);
                    accum[b] = 0.f; accum2[b] = 0.f;
                }
            }
        };
        uint32_t row = c_begin / cpr, rem = c_begin - row * cpr;
        for (uint32_t g = 0; g < nseg; ++g) {
            cp_async_wait_g<DEPTH - 1>();               // segment g has landed (groups complete in order)
            __syncwarp();
            const unsigned char* sp = ring + (g % DEPTH) * GV_SEG_BYTES;
            const uint32_t cb = c_begin + g * GV_SEG_CHUNKS;
            uint4 w[GV_SEG_CHUNKS];
#pragma unroll
            for (int c = 0; c < GV_SEG_CHUNKS; ++c)
                if (cb + c < c_end) w[c] = *reinterpret_cast<const uint4*>(sp + c * 512);
#pragma unroll
            for (int c = 0; c < GV_SEG_CHUNKS; ++c) {
                if (cb + c < c_end) {
                    const int idx = (int)(rem << 5) + lane;                     // 8-element chunk index in the row
                    if ((int)row != cur_row) { flush(); cur_row = (int)row; }
                    const float w0 = bf16lo(w[c].x), w1 = bf16hi(w[c].x), w2 = bf16lo(w[c].y), w3 = bf16hi(w[c].y);
                    const float w4 = bf16lo(w[c].z), w5 = bf16hi(w[c].z), w6 = bf16lo(w[c].w), w7 = bf16hi(w[c].w);
#pragma unroll
                    for (int b = 0; b < B; ++b) {
                        const float4 xl = xs[(size_t)(b * 2 + 0) * K8 + idx];
                        const float4 xh = xs[(size_t)(b * 2 + 1) * K8 + idx];
                        float t = accum[b], u = accum2[b];
                        t = fmaf(w0, xl.x, t); u = fmaf(w4, xh.x, u);
                        t = fmaf(w1, xl.y, t); u = fmaf(w5, xh.y, u);
                        t = fmaf(w2, xl.z, t); u = fmaf(w6, xh.z, u);
                        t = fmaf(w3, xl.w, t); u = fmaf(w7, xh.w, u);
                        accum[b] = t; accum2[b] = u;
                    }
                }
                if (++rem == cpr) { rem = 0; ++row; }
            }
            __syncwarp();                               // every lane has its slot data in registers
            issue(g + DEPTH);                           // refill the slot just drained
        }
        flush();
    }
    __syncthreads();

    // ------------------------------- epilogue (all threads) -------------------------------
    float bestv[B];
    int besti[B];
#pragma unroll
    for (int b = 0; b < B; ++b) { bestv[b] = -INFINITY; besti[b] = 0x7fffffff; }
    if constexpr (EPI == GEMV_SILU_MUL) {
        for (int u = tid; u < nrows / 2; u += GV_THREADS) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float r = NORM ? rstd_s[b] : 1.f;
                const float g = acc_s[(size_t)b * rpc + 2 * u] * r, up = acc_s[(size_t)b * rpc + 2 * u + 1] * r;
                a.y[(size_t)b * a.ldy + (r0 / 2 + u)] = silu_f(g) * up;
            }
        }
    } else {
        for (int i = tid; i < nrows; i += GV_THREADS) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float v = acc_s[(size_t)b * rpc + i] * (NORM ? rstd_s[b] : 1.f);
                float* yp = a.y + (size_t)b * a.ldy + r0 + i;
                if constexpr (EPI == GEMV_RESID) *yp += v; else *yp = v;
                if constexpr (EPI == GEMV_LOGITS_ARGMAX) {
                    if (v > bestv[b]) { bestv[b] = v; besti[b] = r0 + i; }   // a thread's rows ascend: first maximum wins
                }
            }
        }
    }

    if constexpr (EPI == GEMV_LOGITS_ARGMAX) {
        // CTA-level (value, lowest index) reduction, then last-CTA-done finalisation.
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bestv[b], o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti[b], o);
                if (ov > bestv[b] || (ov == bestv[b] && oi < besti[b])) { bestv[b] = ov; besti[b] = oi; }
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < B; ++b) { wbest_v[warp][b] = bestv[b]; wbest_i[warp][b] = besti[b]; }
        }
        __syncthreads();
        if (tid < B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int w = 0; w < GV_WARPS; ++w) {
                const float v = wbest_v[w][tid]; const int i = wbest_i[w][tid];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            a.part_val[tid * gridDim.x + blockIdx.x] = bv;
            a.part_idx[tid * gridDim.x + blockIdx.x] = bi;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int t = atomicAdd(a.ticket, 1u);
            is_last_s = (t == gridDim.x - 1);
        }
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
                    SeqState* s = a.state + b;
                    if (a.out_tokens) a.out_tokens[(size_t)b * a.out_stride + s->step] = (uint32_t)bi;
                    uint32_t fed = (uint32_t)bi;
                    if (a.force_tokens) fed = a.force_tokens[s->step];
                    if (a.advance == 2) fed = s->token;
                    tok_s[b] = fed;
                    if (a.advance) {
                        s->token = fed;
                        s->kv_len += 1;
                        s->pos[0] += 1; s->pos[1] += 1; s->pos[2] += 1;
                    }
                    s->step += 1;
                }
            }
            if (tid == 0) *a.ticket = 0u;
            __syncthreads();
            if (a.advance && a.embed != nullptr) {   // gather the next step's input embedding (bf16 -> f32 residual stream)
                for (int b = 0; b < B; ++b) {
                    const bf16* row = a.embed + (size_t)tok_s[b] * a.H;
                    for (int i = tid; i < a.H; i += GV_THREADS) a.x_next[(size_t)b * a.H + i] = __bfloat162float(row[i]);
                }
            }
        }
    }
}

static size_t gemv_smem_bytes(int B, int K, int N, int rows_per_unit, int grid) {
    const int units = N / rows_per_unit;
    const int rpc = (units + grid - 1) / grid * rows_per_unit;
    return (size_t)GV_WARPS * gv_depth(B) * GV_SEG_BYTES + (size_t)B * K * 4 + (size_t)B * rpc * 4 + 16;
}

template <int B, int EPI, bool NORM>
static int gemv_launch_t(cudaStream_t st, const GemvArgs& a, int num_sms, bool pdl) {
    const size_t smem = gemv_smem_bytes(B, a.K, a.N, EPI == GEMV_SILU_MUL ? 2 : 1, num_sms);
    if (smem > 227 * 1024) return -1000;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(gemv_kernel<B, EPI, NORM>, smem, seen)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms);
    cfg.blockDim = dim3(GV_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, gemv_kernel<B, EPI, NORM>, a);
}

template <int B>
static int gemv_launch_b(cudaStream_t st, int epi, bool norm, const GemvArgs& a, int num_sms, bool pdl) {
    switch (epi) {
        case GEMV_STORE:
            return norm ? gemv_launch_t<B, GEMV_STORE, true>(st, a, num_sms, pdl) : gemv_launch_t<B, GEMV_STORE, false>(st, a, num_sms, pdl);
        case GEMV_RESID: return gemv_launch_t<B, GEMV_RESID, false>(st, a, num_sms, pdl);
        case GEMV_SILU_MUL:
            return norm ? gemv_launch_t<B, GEMV_SILU_MUL, true>(st, a, num_sms, pdl) : gemv_launch_t<B, GEMV_SILU_MUL, false>(st, a, num_sms, pdl);
        case GEMV_LOGITS_ARGMAX: return gemv_launch_t<B, GEMV_LOGITS_ARGMAX, true>(st, a, num_sms, pdl);
        default: return -1000;
    }
}

// Largest sequence group (4, 2 or 1) whose f32 activations fit next to the weight rings for an in_dim of K.
int gemv_max_group(int K, int N, int num_sms) {
    for (int B : {4, 2})
        if (gemv_smem_bytes(B, K, N, 1, num_sms) <= 227 * 1024) return B;
    return 1;
}

int gemv_launch(cudaStream_t st, int B, int epi, bool norm, const GemvArgs& a, int num_sms, bool pdl) {
    if ((a.K % 256) != 0 || a.N <= 0) return -1000;      // a 512-byte chunk must not straddle rows
    if (epi == GEMV_SILU_MUL && (a.N % 2) != 0) return -1000;
    switch (B) {
        case 1: return gemv_launch_b<1>(st, epi, norm, a, num_sms, pdl);
        case 2: return gemv_launch_b<2>(st, epi, norm, a, num_sms, pdl);
        case 4: return gemv_launch_b<4>(st, epi, norm, a, num_sms, pdl);
        default: return -1000;
    }
}

// =====================================================================================
// Next-token embedding gather for the host-driven path (token id arrives from the host)
// =====================================================================================
__global__ void __launch_bounds__(256)
embed_decode_kernel(const bf16* __restrict__ embed, int H, const SeqState* __restrict__ state, float* __restrict__ x) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x;
    const bf16* row = embed + (size_t)state[b].token * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) x[(size_t)b * H + i] = __bfloat162float(row[i]);
}

int embed_decode_launch(cudaStream_t st, int B, const bf16* embed, int H, const SeqState* state, float* x, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, embed_decode_kernel, embed, H, state, x);
}

// =====================================================================================
// Decode attention: QK-norm + (M)RoPE + KV-page append + split-KV GQA attention + in-cluster split merge
//   grid (nkv * ATTN_NSPLIT, B), 256 threads, thread-block CLUSTER of ATTN_NSPLIT CTAs per KV head.
//   * every CTA first pulls its token range of the K/V pages into shared memory with cp.async -- BEFORE
//     griddepcontrol.wait: cached K/V and the sequence state were written at least two kernels ago, so the whole
//     KV read overlaps the QKV GEMV that precedes this kernel;
//   * all NREP query heads of the KV group share every K/V byte (read once per group, 16-byte vectors);
//   * the split partials are merged through distributed shared memory (no global round trip, no atomics).
// =====================================================================================
template <int D, int NREP>
__global__ void __launch_bounds__(256)
attn_decode_kernel(AttnDecArgs a) {
    namespace cg = cooperative_groups;
    constexpr int EPL = 8;               // bf16 elements per 16-byte lane load
    constexpr int LPT = D / EPL;         // lanes per token (16 for D=128, 32 for D=256)
    constexpr int TPW = 32 / LPT;        // tokens per warp pass
    constexpr int NW = 8;
    static_assert(NREP < NW, "one q/k vector per warp");
    constexpr int TILE = 32768 / (D * 2);   // tokens per shared-memory tile (K and V: 32 KB each)
    constexpr int CPR = D / 8;              // 16-byte chunks per token row
    extern __shared__ __align__(16) unsigned char asm_[];
    bf16* k_t = reinterpret_cast<bf16*>(asm_);
    bf16* v_t = k_t + TILE * D;
    bf16* kl_t = v_t + TILE * D;          // split precision only: low-order planes of the same tile
    bf16* vl_t = kl_t + TILE * D;
    const bool split_kv = a.kv_bits ? (a.kv_split != 0) : (a.kv_lo_off != 0);
    float* mo_s = reinterpret_cast<float*>(asm_);          // [NW][NREP][D], aliases the tiles after the token loop
    __shared__ float q_s[NREP][D];
    __shared__ float knew_s[D], vnew_s[D];
    __shared__ float mm_s[NW][NREP], ml_s[NW][NREP];
    __shared__ float cta_o[NREP][D];
    __shared__ float cta_ml[NREP][2];

    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a KV head whose query group is wider than NREP is served by `groups` CTA clusters of NREP query heads each (6 = 2 x 3,
    // 8 = 2 x 4, ...): kvv numbers those sub-groups, query head = kvv * NREP + h; only sub-group 0 appends the new K/V row
    const int kvv = blockIdx.x / ATTN_NSPLIT, split = blockIdx.x % ATTN_NSPLIT;
    const int kvh = kvv / a.groups, sub = kvv % a.groups;
    const int b = blockIdx.y;
    const int q_dim = a.nh * D, kv_dim = a.nkv * D;
    const int q_span = a.nh * a.q_stride;

    // ---- token range of this split, then start the K/V copy (independent of the previous kernel) ----
    const SeqState st = a.state[b];
    const int T = st.kv_len + 1;                      // including the token being decoded
    int chunk = (T + ATTN_NSPLIT - 1) / ATTN_NSPLIT;
    chunk = (chunk + 7) & ~7;
    const int t0 = split * chunk;
    const int t1 = min(T, t0 + chunk);
    const int s_last = (T - 1) / chunk;
    const int t_end = min(t1, T - 1);                 // cached tokens only; position T-1 comes from this step's qkv
    const int* bt = a.block_table + (size_t)st.slot * a.max_pages;
    auto load_tile = [&](int tb) {
        const int n = min(TILE, t_end - tb);
        if (a.kv_bits) {
            // codes -> (code - offset) * scale -> bf16 hi (+ lo) rows of the tile: 16 elements per thread and pass
            const int off = a.kv_bits == 8 ? 128 : 8;
            for (int c = tid; c < n * (D / 16) * 2; c += 256) {
                const int which = c & 1, cc = c >> 1;
                const int r = cc / (D / 16), ch = cc % (D / 16);
                const int t = tb + r;
                const size_t prow = ((size_t)bt[t / KV_PAGE] * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE);
                const unsigned char* codes = which ? a.v_codes : a.k_codes;
                const float scale = (which ? a.v_scale : a.k_scale)[prow];
                int q[16];
                if (a.kv_bits == 8) {
                    const uint4 u = *reinterpret_cast<const uint4*>(codes + prow * D + ch * 16);
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e = 0; e < 16; ++e) q[e] = (int)((w[e >> 2] >> (8 * (e & 3))) & 255u);
                } else {
                    const uint2 u = *reinterpret_cast<const uint2*>(codes + prow * (D / 2) + ch * 8);
                    const uint32_t w[2] = {u.x, u.y};
#pragma unroll
                    for (int e = 0; e < 16; ++e) q[e] = (int)((w[e >> 3] >> (4 * (e & 7))) & 15u);
                }
                bf16* hi_t = (which ? v_t : k_t) + r * D + ch * 16;
                bf16* lo_t = (which ? vl_t : kl_t) + r * D + ch * 16;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float y = (float)(q[e] - off) * scale;
                    const bf16 hb = __float2bfloat16_rn(y);
                    hi_t[e] = hb;
                    if (split_kv) lo_t[e] = __float2bfloat16_rn(y - __bfloat162float(hb));
                }
            }
            return;
        }
        for (int c = tid; c < n * CPR; c += 256) {
            const int r = c / CPR, ch = c % CPR;
            const int t = tb + r;
            const int page = bt[t / KV_PAGE];
            const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D + ch * 8;
            const uint32_t kd = smem_u32(k_t + r * D + ch * 8), vd = smem_u32(v_t + r * D + ch * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kd), "l"(a.k_pool + off) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vd), "l"(a.v_pool + off) : "memory");
            if (split_kv) {
                const uint32_t kld = smem_u32(kl_t + r * D + ch * 8), vld = smem_u32(vl_t + r * D + ch * 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kld), "l"(a.k_pool + a.kv_lo_off + off) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vld), "l"(a.v_pool + a.kv_lo_off + off) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (t0 < t_end) load_tile(t0);

    // norm weights and this position's cos/sin depend only on immutable data and the sequence state: fetch them before
    // the dependency wait too (lane l owns elements l + 32 j, so a rotary pair is lane-local)
    constexpr int NE = D / 32;
    const int RJ = a.rot_half >> 5;                    // lane-local rotary pairs (j, j + RJ), j < RJ
    const bool vec_is_k = (warp == NREP);
    float nwv[NE], cs_c[NE], cs_s[NE];
    if (warp <= NREP) {
        const float* nw = vec_is_k ? a.k_norm_w : a.q_norm_w;
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            nwv[j] = nw[lane + 32 * j];
            cs_c[j] = 1.f; cs_s[j] = 0.f;
            if (j < 2 * RJ) {
                const int i = lane + 32 * (j < RJ ? j : j - RJ);
                const int p = st.pos[a.axis_of[i]];
                cs_c[j] = a.cos_tab[(size_t)p * a.rot_half + i];
                cs_s[j] = a.sin_tab[(size_t)p * a.rot_half + i];
            }
        }
    }

    pdl_wait();
    pdl_launch_dependents();
    const float* qkv = a.qkv + (size_t)b * (q_span + 2 * kv_dim);

    // ---- q (NREP heads) and, on the split that owns position T-1, the new k: RMSNorm then rotate ----
    for (int vec = warp; vec < NREP + 1; vec += NW) {      // NW > NREP: one vector per warp, vec == warp
        const bool is_k = (vec == NREP);
        if (is_k && split != s_last) continue;
        const float* src = is_k ? (qkv + q_span + kvh * D) : (qkv + (kvv * NREP + vec) * a.q_stride);
        float e[NE];
        float ssq = 0.f;
#pragma unroll
        for (int j = 0; j < NE; ++j) { e[j] = src[lane + 32 * j]; ssq += e[j] * e[j]; }
        ssq = warp_sum(ssq);
        const float rstd = rsqrtf(ssq / (float)D + a.eps);
#pragma unroll
        for (int j = 0; j < NE; ++j) e[j] = e[j] * rstd * nwv[j];
        float* dst = is_k ? knew_s : q_s[vec];
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            float r = e[j];
            if (j < 2 * RJ) {
                const bool lo = j < RJ;
                const float c = cs_c[j], s = cs_s[j];
                float other = 0.f;
#pragma unroll
                for (int jj = 0; jj < NE; ++jj) if (jj == (lo ? j + RJ : j - RJ)) other = e[jj];
                r = lo ? (e[j] * c - other * s) : (other * s + e[j] * c);
            }
            dst[lane + 32 * j] = (is_k && !a.kv_bits) ? (split_kv ? round_bf16_split(r) : round_bf16(r)) : r;
        }
    }
    if (split == s_last && warp == NW - 1) {
        const float* vsrc = qkv + q_span + kv_dim + kvh * D;
        for (int i = lane; i < D; i += 32) vnew_s[i] = a.kv_bits ? vsrc[i] : (split_kv ? round_bf16_split(vsrc[i]) : round_bf16(vsrc[i]));
    }
    __syncthreads();
    if (a.kv_bits && split == s_last && warp < 2) {
        // quantise the new K (warp 0) / V (warp 1) row per token and head; attention -- here and in every later step -- sees code * scale
        float* row = warp ? vnew_s : knew_s;
        const int t = T - 1;
        const size_t prow = ((size_t)bt[t / KV_PAGE] * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE);
        float am = 0.f;
        for (int i = lane; i < D; i += 32) am = fmaxf(am, fabsf(row[i]));
        am = warp_max(am);
        const float inv = a.kv_bits == 8 ? (float)(1.0 / 127.0) : (float)(1.0 / 7.0);
        const float scale = __fadd_rn(__fmul_rn(am, inv), 1e-8f);
        const int off = a.kv_bits == 8 ? 128 : 8;
        unsigned char* codes = warp ? a.v_codes : a.k_codes;
        for (int i = 2 * lane; i < D; i += 64) {          // an (even, odd) pair per lane and pass
            const int q0 = (int)roundf(__fdiv_rn(row[i], scale)), q1 = (int)roundf(__fdiv_rn(row[i + 1], scale));
            row[i] = (float)q0 * scale; row[i + 1] = (float)q1 * scale;
            if (sub == 0) {
                if (a.kv_bits == 8) { codes[prow * D + i] = (unsigned char)(q0 + off); codes[prow * D + i + 1] = (unsigned char)(q1 + off); }
                else codes[prow * (D / 2) + (i >> 1)] = (unsigned char)((q0 + off) + 16 * (q1 + off));
            }
        }
        if (sub == 0 && lane == 0) (warp ? a.v_scale : a.k_scale)[prow] = scale;
    }
    if (a.kv_bits) __syncthreads();
    if (!a.kv_bits && split == s_last && sub == 0) {   // append the new token to its page
        const int t = T - 1;
        const int page = bt[t / KV_PAGE];
        const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D;
        for (int i = tid; i < D; i += 256) {
            const bf16 kh = __float2bfloat16_rn(knew_s[i]), vh = __float2bfloat16_rn(vnew_s[i]);
            a.k_pool[off + i] = kh;
            a.v_pool[off + i] = vh;
            if (split_kv) {
                a.k_pool[a.kv_lo_off + off + i] = __float2bfloat16_rn(knew_s[i] - __bfloat162float(kh));
                a.v_pool[a.kv_lo_off + off + i] = __float2bfloat16_rn(vnew_s[i] - __bfloat162float(vh));
            }
        }
    }

    // ---- the cached tokens of this split, tile by tile out of shared memory ----
    const int grp = lane / LPT, gl = lane % LPT;        // token group inside the warp, lane inside the group
    float qr[NREP][EPL];
#pragma unroll
    for (int h = 0; h < NREP; ++h)
#pragma unroll
        for (int j = 0; j < EPL; ++j) qr[h][j] = q_s[h][gl * EPL + j] * a.scale;
    float m[NREP], l[NREP], o[NREP][EPL];
#pragma unroll
    for (int h = 0; h < NREP; ++h) {
        m[h] = -INFINITY; l[h] = 0.f;
#pragma unroll
        for (int j = 0; j < EPL; ++j) o[h][j] = 0.f;
    }
    for (int tb = t0; tb < t_end; tb += TILE) {
        if (tb != t0) { __syncthreads(); load_tile(tb); }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int n = min(TILE, t_end - tb);
        for (int r0 = warp * TPW; r0 < n; r0 += NW * TPW) {
            const int r = r0 + grp;
            const bool valid = r < n;
            float kf[EPL], vf[EPL];
            if (valid) {
                const uint4 kr = *reinterpret_cast<const uint4*>(k_t + r * D + gl * EPL);
                const uint4 vr = *reinterpret_cast<const uint4*>(v_t + r * D + gl * EPL);
                kf[0] = bf16lo(kr.x); kf[1] = bf16hi(kr.x); kf[2] = bf16lo(kr.y); kf[3] = bf16hi(kr.y);
                kf[4] = bf16lo(kr.z); kf[5] = bf16hi(kr.z); kf[6] = bf16lo(kr.w); kf[7] = bf16hi(kr.w);
                vf[0] = bf16lo(vr.x); vf[1] = bf16hi(vr.x); vf[2] = bf16lo(vr.y); vf[3] = bf16hi(vr.y);
                vf[4] = bf16lo(vr.z); vf[5] = bf16hi(vr.z); vf[6] = bf16lo(vr.w); vf[7] = bf16hi(vr.w);
                if (split_kv) {
                    const uint4 kl = *reinterpret_cast<const uint4*>(kl_t + r * D + gl * EPL);
                    const uint4 vl = *reinterpret_cast<const uint4*>(vl_t + r * D + gl * EPL);
                    kf[0] += bf16lo(kl.x); kf[1] += bf16hi(kl.x); kf[2] += bf16lo(kl.y); kf[3] += bf16hi(kl.y);
                    kf[4] += bf16lo(kl.z); kf[5] += bf16hi(kl.z); kf[6] += bf16lo(kl.w); kf[7] += bf16hi(kl.w);
                    vf[0] += bf16lo(vl.x); vf[1] += bf16hi(vl.x); vf[2] += bf16lo(vl.y); vf[3] += bf16hi(vl.y);
                    vf[4] += bf16lo(vl.z); vf[5] += bf16hi(vl.z); vf[6] += bf16lo(vl.w); vf[7] += bf16hi(vl.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPL; ++j) { kf[j] = 0.f; vf[j] = 0.f; }
            }
#pragma unroll
            for (int h = 0; h < NREP; ++h) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < EPL; ++j) s = fmaf(qr[h][j], kf[j], s);
#pragma unroll
                for (int ofs = LPT / 2; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
                if (valid) {
                    const float mn = fmaxf(m[h], s);
                    const float corr = __expf(m[h] - mn), p = __expf(s - mn);
                    l[h] = l[h] * corr + p;
#pragma unroll
                    for (int j = 0; j < EPL; ++j) o[h][j] = o[h][j] * corr + p * vf[j];
                    m[h] = mn;
                }
            }
        }
    }
    // the token being decoded (K/V still in shared memory): group 0 of warp 0 on the owning split
    if (split == s_last && warp == 0 && grp == 0) {
#pragma unroll
        for (int h = 0; h < NREP; ++h) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < EPL; ++j) s = fmaf(qr[h][j], knew_s[gl * EPL + j], s);
#pragma unroll
            for (int ofs = LPT / 2; ofs > 0; ofs >>= 1) s += __shfl_xor_sync((TPW == 1) ? 0xffffffffu : 0x0000ffffu, s, ofs);
            const float mn = fmaxf(m[h], s);
            const float corr = __expf(m[h] - mn), p = __expf(s - mn);
            l[h] = l[h] * corr + p;
#pragma unroll
            for (int j = 0; j < EPL; ++j) o[h][j] = o[h][j] * corr + p * vnew_s[gl * EPL + j];
            m[h] = mn;
        }
    }
    // merge the token groups of a warp
    if (TPW == 2) {
#pragma unroll
        for (int h = 0; h < NREP; ++h) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[h], 16);
            const float lo = __shfl_xor_sync(0xffffffffu, l[h], 16);
            const float mn = fmaxf(m[h], mo);
            const float c0 = (m[h] == -INFINITY) ? 0.f : __expf(m[h] - mn);
            const float c1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
            l[h] = l[h] * c0 + lo * c1;
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const float oo = __shfl_xor_sync(0xffffffffu, o[h][j], 16);
                o[h][j] = o[h][j] * c0 + oo * c1;
            }
            m[h] = mn;
        }
    }
    __syncthreads();                                     // tiles are dead: reuse them as the warp-merge buffer
    if (grp == 0) {
#pragma unroll
        for (int h = 0; h < NREP; ++h) {
            if (gl == 0) { mm_s[warp][h] = m[h]; ml_s[warp][h] = l[h]; }
#pragma unroll
            for (int j = 0; j < EPL; ++j) mo_s[((size_t)warp * NREP + h) * D + gl * EPL + j] = o[h][j];
        }
    }
    __syncthreads();
    // merge the warps -> this CTA's partial in shared memory
    for (int idx = tid; idx < NREP * D; idx += 256) {
        const int h = idx / D, i = idx % D;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, mm_s[w][h]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float c = (mm_s[w][h] == -INFINITY) ? 0.f : __expf(mm_s[w][h] - M);
            L += ml_s[w][h] * c;
            O += mo_s[((size_t)w * NREP + h) * D + i] * c;
        }
        cta_o[h][i] = O;
        if (i == 0) { cta_ml[h][0] = M; cta_ml[h][1] = L; }
    }
    // ---- merge the splits of this KV group through distributed shared memory ----
    cluster.sync();
    for (int idx = split * 256 + tid; idx < NREP * D; idx += ATTN_NSPLIT * 256) {
        const int h = idx / D, i = idx % D;
        float ms[ATTN_NSPLIT];
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < ATTN_NSPLIT; ++s) {
            const float* rml = cluster.map_shared_rank(&cta_ml[0][0], s);
            ms[s] = rml[h * 2];
            M = fmaxf(M, ms[s]);
        }
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int s = 0; s < ATTN_NSPLIT; ++s) {
            const float c = (ms[s] == -INFINITY) ? 0.f : __expf(ms[s] - M);
            const float* rml = cluster.map_shared_rank(&cta_ml[0][0], s);
            const float* ro = cluster.map_shared_rank(&cta_o[0][0], s);
            L += rml[h * 2 + 1] * c;
            O += ro[h * D + i] * c;
        }
        const int head = kvv * NREP + h;
        float r = O / L;
        if (a.gated) r *= 1.0f / (1.0f + expf(-qkv[head * a.q_stride + D + i]));   // y * sigmoid(gate), modeling.rs:516-523
        a.out[(size_t)b * q_dim + head * D + i] = r;
    }
    cluster.sync();                                      // nobody exits while its shared memory is still being read
}

template <int D, int NREP>
static int attn_decode_launch_t(cudaStream_t st, int B, const AttnDecArgs& a, bool pdl) {
    const int SMEM = (a.kv_bits ? a.kv_split : (a.kv_lo_off != 0)) ? 131072 : 65536;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(attn_decode_kernel<D, NREP>, (size_t)SMEM, seen)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.nkv * a.groups * ATTN_NSPLIT, B);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ATTN_NSPLIT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    return (int)cudaLaunchKernelEx(&cfg, attn_decode_kernel<D, NREP>, a);
}

int attn_decode_launch(cudaStream_t st, int B, int D, const AttnDecArgs& a_in, bool pdl) {
    AttnDecArgs a = a_in;
    if (a.nkv <= 0 || a.nh <= 0 || a.nh % a.nkv) return -1000;
    const int nrep = a.nh / a.nkv;
    if (a.rot_half < 32 || (a.rot_half % 32) != 0 || 2 * a.rot_half > D) return -1000;
    // any group width: the widest instantiated sub-group (4, 3, 2, 1 query heads) that divides it
    const int sub = nrep % 4 == 0 ? 4 : nrep % 3 == 0 ? 3 : nrep % 2 == 0 ? 2 : 1;
    a.groups = nrep / sub;
    if (D == 128) {
        switch (sub) {
            case 1: return attn_decode_launch_t<128, 1>(st, B, a, pdl);
            case 2: return attn_decode_launch_t<128, 2>(st, B, a, pdl);
            case 3: return attn_decode_launch_t<128, 3>(st, B, a, pdl);
            case 4: return attn_decode_launch_t<128, 4>(st, B, a, pdl);
        }
    } else if (D == 256) {
        switch (sub) {
            case 1: return attn_decode_launch_t<256, 1>(st, B, a, pdl);
            case 2: return attn_decode_launch_t<256, 2>(st, B, a, pdl);
            case 3: return attn_decode_launch_t<256, 3>(st, B, a, pdl);
            case 4: return attn_decode_launch_t<256, 4>(st, B, a, pdl);
        }
    }
    return -1000;
}

}  // namespace cb
