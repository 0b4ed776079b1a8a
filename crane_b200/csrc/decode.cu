// Decode-path kernels.  See decode.cuh for the design.
#include "decode.cuh"

namespace cb {

constexpr int GEMV_THREADS = 512;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;
constexpr int GEMV_U = 4;   // 16-byte weight loads in flight per lane per batch

// Dot products of one weight row with the B staged activation vectors.
// xs layout: [B][2][K8] float4 -- for 8-element chunk i of vector b, xs[(b*2+0)*K8+i] holds elements
// 8i..8i+3 and xs[(b*2+1)*K8+i] elements 8i+4..8i+7 (lane-contiguous float4 reads: no bank conflicts).
template <int B>
__device__ __forceinline__ void row_dot(const uint4* __restrict__ wrow, int K8, const float4* __restrict__ xs,
                                        int lane, const uint4* pre, bool use_pre, float* acc) {
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = 0.f;
    for (int c0 = lane; c0 < K8; c0 += 32 * GEMV_U) {
        uint4 w[GEMV_U];
        if (use_pre && c0 == lane) {
#pragma unroll
            for (int u = 0; u < GEMV_U; ++u) w[u] = pre[u];
        } else {
#pragma unroll
            for (int u = 0; u < GEMV_U; ++u) {
                const int idx = c0 + 32 * u;
                w[u] = (idx < K8) ? ldg_stream(wrow + idx) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < GEMV_U; ++u) {
            const int idx = c0 + 32 * u;
            if (idx < K8) {
                const float w0 = bf16lo(w[u].x), w1 = bf16hi(w[u].x), w2 = bf16lo(w[u].y), w3 = bf16hi(w[u].y);
                const float w4 = bf16lo(w[u].z), w5 = bf16hi(w[u].z), w6 = bf16lo(w[u].w), w7 = bf16hi(w[u].w);
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float4 xl = xs[(size_t)(b * 2 + 0) * K8 + idx];
                    const float4 xh = xs[(size_t)(b * 2 + 1) * K8 + idx];
                    float s = acc[b];
                    s = fmaf(w0, xl.x, s); s = fmaf(w1, xl.y, s); s = fmaf(w2, xl.z, s); s = fmaf(w3, xl.w, s);
                    s = fmaf(w4, xh.x, s); s = fmaf(w5, xh.y, s); s = fmaf(w6, xh.z, s); s = fmaf(w7, xh.w, s);
                    acc[b] = s;
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = warp_sum(acc[b]);
}

template <int B, int EPI, bool NORM>
__global__ void __launch_bounds__(GEMV_THREADS, (B <= 2) ? 2 : 1)
gemv_kernel(GemvArgs a) {
    extern __shared__ float4 xs[];            // [B][2][K8]
    __shared__ float red[32];
    __shared__ float rstd_s[B];
    __shared__ float wbest_v[GEMV_WARPS][B];
    __shared__ int wbest_i[GEMV_WARPS][B];
    __shared__ int is_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K8 = a.K >> 3;
    constexpr int ROWS_PER_UNIT = (EPI == GEMV_SILU_MUL) ? 2 : 1;
    const int units = a.N / ROWS_PER_UNIT;
    const int upc = (units + gridDim.x - 1) / gridDim.x;          // units per CTA
    const int u0 = blockIdx.x * upc;
    const int u1 = min(units, u0 + upc);

    // --- weights do not depend on the previous kernel: start streaming before the PDL wait ---
    uint4 pre[GEMV_U];
    const int ufirst = u0 + warp;
    const bool have_first = ufirst < u1;
    if (have_first) {
        const uint4* wrow = reinterpret_cast<const uint4*>(a.W + (size_t)(ufirst * ROWS_PER_UNIT) * a.K);
#pragma unroll
        for (int u = 0; u < GEMV_U; ++u) {
            const int idx = lane + 32 * u;
            pre[u] = (idx < K8) ? ldg_stream(wrow + idx) : make_uint4(0, 0, 0, 0);
        }
    }
    pdl_wait();

    // --- stage activations (f32) and, if NORM, fold the RMSNorm weight in and get 1/rms ---
#pragma unroll
    for (int b = 0; b < B; ++b) {
        float ssq = 0.f;
        for (int i = tid; i < K8; i += GEMV_THREADS) {
            const float4* xp = reinterpret_cast<const float4*>(a.x + (size_t)b * a.ldx) + 2 * i;
            float4 lo = xp[0], hi = xp[1];
            if (NORM) {
                ssq += lo.x * lo.x + lo.y * lo.y + lo.z * lo.z + lo.w * lo.w + hi.x * hi.x + hi.y * hi.y + hi.z * hi.z + hi.w * hi.w;
                const float4* gp = reinterpret_cast<const float4*>(a.norm_w) + 2 * i;
                const float4 g0 = gp[0], g1 = gp[1];
                lo.x *= g0.x; lo.y *= g0.y; lo.z *= g0.z; lo.w *= g0.w;
                hi.x *= g1.x; hi.y *= g1.y; hi.z *= g1.z; hi.w *= g1.w;
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
    pdl_launch_dependents();

    float bestv[B];
    int besti[B];
#pragma unroll
    for (int b = 0; b < B; ++b) { bestv[b] = -INFINITY; besti[b] = 0x7fffffff; }

    bool first = true;
    for (int un = u0 + warp; un < u1; un += GEMV_WARPS) {
        float acc[B];
        if constexpr (EPI == GEMV_SILU_MUL) {
            float g[B];
            row_dot<B>(reinterpret_cast<const uint4*>(a.W + (size_t)(2 * un) * a.K), K8, xs, lane, pre, first, g);
            row_dot<B>(reinterpret_cast<const uint4*>(a.W + (size_t)(2 * un + 1) * a.K), K8, xs, lane, pre, false, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float r = NORM ? rstd_s[b] : 1.f;
                    a.y[(size_t)b * a.ldy + un] = silu_f(g[b] * r) * (acc[b] * r);
                }
            }
        } else {
            row_dot<B>(reinterpret_cast<const uint4*>(a.W + (size_t)un * a.K), K8, xs, lane, pre, first, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float v = acc[b] * (NORM ? rstd_s[b] : 1.f);
                    float* yp = a.y + (size_t)b * a.ldy + un;
                    if constexpr (EPI == GEMV_RESID) *yp += v; else *yp = v;
                    if constexpr (EPI == GEMV_LOGITS_ARGMAX) {
                        if (v > bestv[b]) { bestv[b] = v; besti[b] = un; }   // rows ascend: first maximum wins
                    }
                }
            }
        }
        first = false;
    }

    if constexpr (EPI == GEMV_LOGITS_ARGMAX) {
        // CTA-level (value, lowest index) reduction, then last-CTA-done finalisation.
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < B; ++b) { wbest_v[warp][b] = bestv[b]; wbest_i[warp][b] = besti[b]; }
        }
        __syncthreads();
        if (tid < B) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int w = 0; w < GEMV_WARPS; ++w) {
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
                if (lane == 0) {
                    tok_s[b] = (uint32_t)bi;
                    SeqState* s = a.state + b;
                    if (a.out_tokens) a.out_tokens[(size_t)b * a.out_stride + s->step] = (uint32_t)bi;
                    if (a.advance) {
                        s->token = (uint32_t)bi;
                        s->kv_len += 1;
                        s->pos[0] += 1; s->pos[1] += 1; s->pos[2] += 1;
                    }
                    s->step += 1;
                }
            }
            if (tid == 0) *a.ticket = 0u;
            __syncthreads();
            if (a.advance) {   // gather the next step's input embedding (bf16 -> f32 residual stream)
                for (int b = 0; b < B; ++b) {
                    const bf16* row = a.embed + (size_t)tok_s[b] * a.H;
                    for (int i = tid; i < a.H; i += GEMV_THREADS) a.x_next[(size_t)b * a.H + i] = __bfloat162float(row[i]);
                }
            }
        }
    }
}

template <int B, int EPI, bool NORM>
static int gemv_launch_t(cudaStream_t st, const GemvArgs& a, int num_sms, bool pdl) {
    const size_t smem = (size_t)B * a.K * sizeof(float);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(gemv_kernel<B, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms);
    cfg.blockDim = dim3(GEMV_THREADS);
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

int gemv_launch(cudaStream_t st, int B, int epi, bool norm, const GemvArgs& a, int num_sms, bool pdl) {
    if ((a.K % 8) != 0 || a.N <= 0) return -1000;
    if (epi == GEMV_SILU_MUL && (a.N % 2) != 0) return -1000;
    if ((size_t)B * a.K * sizeof(float) > 200 * 1024) return -1000;
    switch (B) {
        case 1: return gemv_launch_b<1>(st, epi, norm, a, num_sms, pdl);
        case 2: return gemv_launch_b<2>(st, epi, norm, a, num_sms, pdl);
        case 4: return gemv_launch_b<4>(st, epi, norm, a, num_sms, pdl);
        case 8: return gemv_launch_b<8>(st, epi, norm, a, num_sms, pdl);
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
// Decode attention: QK-norm + (M)RoPE + KV-page append + split-KV GQA attention + split merge
//   grid (nkv * ATTN_NSPLIT, B), 128 threads.  One CTA = one KV head x one token range, all NREP
//   query heads of the group share every K/V byte it reads (read once per group, 16-byte loads).
// =====================================================================================
template <int D, int NREP>
__global__ void __launch_bounds__(128)
attn_decode_kernel(AttnDecArgs a) {
    constexpr int EPL = 8;               // bf16 elements per 16-byte lane load
    constexpr int LPT = D / EPL;         // lanes per token (16 for D=128, 32 for D=256)
    constexpr int TPW = 32 / LPT;        // tokens per warp pass
    constexpr int NW = 4;
    constexpr int HALF = D / 2;
    __shared__ float q_s[NREP][D];
    __shared__ float knew_s[D], vnew_s[D];
    __shared__ float mo_s[NW][NREP][D];
    __shared__ float mm_s[NW][NREP], ml_s[NW][NREP];
    __shared__ int is_last_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kvh = blockIdx.x / ATTN_NSPLIT, split = blockIdx.x % ATTN_NSPLIT;
    const int b = blockIdx.y;
    const int q_dim = a.nh * D, kv_dim = a.nkv * D;

    pdl_wait();
    pdl_launch_dependents();

    const SeqState st = a.state[b];
    const int T = st.kv_len + 1;                      // including the token being decoded
    int chunk = (T + ATTN_NSPLIT - 1) / ATTN_NSPLIT;
    chunk = (chunk + 7) & ~7;
    const int t0 = split * chunk;
    const int t1 = min(T, t0 + chunk);
    const int s_last = (T - 1) / chunk;
    const float* qkv = a.qkv + (size_t)b * (q_dim + 2 * kv_dim);

    // ---- q (NREP heads) and, on the split that owns position T-1, the new k: RMSNorm then rotate ----
    // Each warp takes vectors round-robin; lane l owns elements l + 32 j so a rotary pair (i, i + D/2) is lane-local.
    constexpr int NE = D / 32;
    for (int vec = warp; vec < NREP + 1; vec += NW) {
        const bool is_k = (vec == NREP);
        if (is_k && split != s_last) continue;
        const float* src = is_k ? (qkv + q_dim + kvh * D) : (qkv + (kvh * NREP + vec) * D);
        const float* nw = is_k ? a.k_norm_w : a.q_norm_w;
        float e[NE];
        float ssq = 0.f;
#pragma unroll
        for (int j = 0; j < NE; ++j) { e[j] = src[lane + 32 * j]; ssq += e[j] * e[j]; }
        ssq = warp_sum(ssq);
        const float rstd = rsqrtf(ssq / (float)D + a.eps);
#pragma unroll
        for (int j = 0; j < NE; ++j) e[j] = e[j] * rstd * nw[lane + 32 * j];
        float* dst = is_k ? knew_s : q_s[vec];
#pragma unroll
        for (int j = 0; j < NE / 2; ++j) {
            const int i = lane + 32 * j;               // rotary column in [0, D/2)
            const int p = st.pos[a.axis_of[i]];
            const float c = a.cos_tab[(size_t)p * HALF + i], s = a.sin_tab[(size_t)p * HALF + i];
            const float x1 = e[j], x2 = e[j + NE / 2];
            float r1 = x1 * c - x2 * s, r2 = x1 * s + x2 * c;
            if (is_k) { r1 = round_bf16(r1); r2 = round_bf16(r2); }
            dst[i] = r1;
            dst[i + HALF] = r2;
        }
    }
    if (split == s_last && warp == NW - 1) {
        const float* vsrc = qkv + q_dim + kv_dim + kvh * D;
        for (int i = lane; i < D; i += 32) vnew_s[i] = round_bf16(vsrc[i]);
    }
    __syncthreads();
    if (split == s_last) {   // append the new token to its page
        const int t = T - 1;
        const int page = a.block_table[(size_t)b * a.max_pages + t / KV_PAGE];
        const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D;
        for (int i = tid; i < D; i += 128) {
            a.k_pool[off + i] = __float2bfloat16_rn(knew_s[i]);
            a.v_pool[off + i] = __float2bfloat16_rn(vnew_s[i]);
        }
    }

    // ---- stream the cached tokens of this split ----
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
    const int t_end = min(t1, T - 1);                   // cached tokens only; T-1 comes from shared memory
    const int* bt = a.block_table + (size_t)b * a.max_pages;
    for (int tb = t0 + warp * TPW; tb < t_end; tb += NW * TPW) {
        const int t = tb + grp;
        const bool valid = t < t_end;
        float kf[EPL], vf[EPL];
        if (valid) {
            const int page = bt[t / KV_PAGE];
            const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D + gl * EPL;
            const uint4 kr = ldg_stream(a.k_pool + off);
            const uint4 vr = ldg_stream(a.v_pool + off);
            kf[0] = bf16lo(kr.x); kf[1] = bf16hi(kr.x); kf[2] = bf16lo(kr.y); kf[3] = bf16hi(kr.y);
            kf[4] = bf16lo(kr.z); kf[5] = bf16hi(kr.z); kf[6] = bf16lo(kr.w); kf[7] = bf16hi(kr.w);
            vf[0] = bf16lo(vr.x); vf[1] = bf16hi(vr.x); vf[2] = bf16lo(vr.y); vf[3] = bf16hi(vr.y);
            vf[4] = bf16lo(vr.z); vf[5] = bf16hi(vr.z); vf[6] = bf16lo(vr.w); vf[7] = bf16hi(vr.w);
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
    if (grp == 0) {
#pragma unroll
        for (int h = 0; h < NREP; ++h) {
            if (gl == 0) { mm_s[warp][h] = m[h]; ml_s[warp][h] = l[h]; }
#pragma unroll
            for (int j = 0; j < EPL; ++j) mo_s[warp][h][gl * EPL + j] = o[h][j];
        }
    }
    __syncthreads();
    // merge the warps, write this split's partial
    for (int idx = tid; idx < NREP * D; idx += 128) {
        const int h = idx / D, i = idx % D;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, mm_s[w][h]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float c = (mm_s[w][h] == -INFINITY) ? 0.f : __expf(mm_s[w][h] - M);
            L += ml_s[w][h] * c;
            O += mo_s[w][h][i] * c;
        }
        const int head = kvh * NREP + h;
        const size_t pbase = ((size_t)b * a.nh + head) * ATTN_NSPLIT + split;
        a.part_o[pbase * D + i] = O;
        if (i == 0) { a.part_ml[pbase * 2 + 0] = M; a.part_ml[pbase * 2 + 1] = L; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int tk = atomicAdd(a.counters + b * a.nkv + kvh, 1u);
        is_last_s = (tk == ATTN_NSPLIT - 1);
        if (is_last_s) a.counters[b * a.nkv + kvh] = 0u;
    }
    __syncthreads();
    if (is_last_s) {   // last split to finish combines all splits of this KV group
        __threadfence();
        for (int idx = tid; idx < NREP * D; idx += 128) {
            const int h = idx / D, i = idx % D;
            const int head = kvh * NREP + h;
            const size_t pbase = ((size_t)b * a.nh + head) * ATTN_NSPLIT;
            float M = -INFINITY;
#pragma unroll
            for (int s = 0; s < ATTN_NSPLIT; ++s) M = fmaxf(M, __ldcg(a.part_ml + (pbase + s) * 2));
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int s = 0; s < ATTN_NSPLIT; ++s) {
                const float ms = __ldcg(a.part_ml + (pbase + s) * 2);
                const float c = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                L += __ldcg(a.part_ml + (pbase + s) * 2 + 1) * c;
                O += __ldcg(a.part_o + (pbase + s) * D + i) * c;
            }
            a.out[(size_t)b * q_dim + head * D + i] = O / L;
        }
    }
}

template <int D, int NREP>
static int attn_decode_launch_t(cudaStream_t st, int B, const AttnDecArgs& a, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.nkv * ATTN_NSPLIT, B);
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, attn_decode_kernel<D, NREP>, a);
}

int attn_decode_launch(cudaStream_t st, int B, int D, const AttnDecArgs& a, bool pdl) {
    const int nrep = a.nh / a.nkv;
    if (D == 128) {
        switch (nrep) {
            case 1: return attn_decode_launch_t<128, 1>(st, B, a, pdl);
            case 2: return attn_decode_launch_t<128, 2>(st, B, a, pdl);
            case 4: return attn_decode_launch_t<128, 4>(st, B, a, pdl);
            case 8: return attn_decode_launch_t<128, 8>(st, B, a, pdl);
        }
    } else if (D == 256) {
        switch (nrep) {
            case 1: return attn_decode_launch_t<256, 1>(st, B, a, pdl);
            case 2: return attn_decode_launch_t<256, 2>(st, B, a, pdl);
            case 4: return attn_decode_launch_t<256, 4>(st, B, a, pdl);
        }
    }
    return -1000;
}

}  // namespace cb
