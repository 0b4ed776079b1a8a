// Prefill / ViT kernels around the tcgen05 GEMMs.
//
// Reference arithmetic being replaced (crane-core/src/models/):
//   embedding gather ................ qwen3/modeling.rs:951
//   RMSNorm rows .................... qwen3/modeling.rs:706,713,1024 (candle_nn::rms_norm, f32 accumulate)
//   QK-norm + rope_thd + KV append .. qwen3/modeling.rs:335-366, modules/kv_cache.rs:38-101
//   causal attention (flash, f32) ... qwen3/modeling.rs:422-456 ; ViT non-causal: qwen3_5/vision.rs:144-172
//   LayerNorm / ViT RoPE / pos-embed  qwen3_5/vision.rs:90-102,215-218,382-489
//   splice / DeepStack .............. qwen3_5/vlm.rs:433-468 ; qwen3_vl/text.rs:280-333
#include "prefill.cuh"

namespace cb {

// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_rows_kernel(const uint32_t* __restrict__ ids, const bf16* __restrict__ embed, int H, float* __restrict__ x) {
    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.x;
    const bf16* row = embed + (size_t)ids[s] * H;
    for (int i = threadIdx.x * 2; i < H; i += 512) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(row + i);
        *reinterpret_cast<float2*>(x + (size_t)s * H + i) = make_float2(bf16lo(u), bf16hi(u));
    }
}
int embed_rows_launch(cudaStream_t st, const uint32_t* ids, int S, const bf16* embed, int H, float* x) {
    return launch_k(embed_rows_kernel, dim3(S), dim3(256), 0, st, prefill_pdl(), ids, embed, H, x);
}

// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsnorm_rows_kernel(const float* __restrict__ x, int H, const float* __restrict__ w, float eps, bf16* __restrict__ out, long long lo_off) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * H;
    float ssq = 0.f;
    for (int i = threadIdx.x * 4; i < H; i += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        ssq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float rstd = rsqrtf(block_sum(ssq, red) / (float)H + eps);
    bf16* o = out + (size_t)blockIdx.x * H;
    for (int i = threadIdx.x * 4; i < H; i += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float4 g = *reinterpret_cast<const float4*>(w + i);
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v.x * rstd * g.x, v.y * rstd * g.y, h0, l0);
        split_bf16x2(v.z * rstd * g.z, v.w * rstd * g.w, h1, l1);
        *reinterpret_cast<uint2*>(o + i) = make_uint2(h0, h1);
        if (lo_off) *reinterpret_cast<uint2*>(o + lo_off + i) = make_uint2(l0, l1);
    }
}
int rmsnorm_rows_launch(cudaStream_t st, const float* x, int S, int H, const float* w, float eps, bf16* out, long long lo_off) {
    if (H % 4) return -1000;
    return launch_k(rmsnorm_rows_kernel, dim3(S), dim3(256), 0, st, prefill_pdl(), x, H, w, eps, out, lo_off);
}

__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, int W, const float* __restrict__ w, const float* __restrict__ b,
                      float eps, bf16* __restrict__ out, long long lo_off) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * W;
    float s = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        s += v.x + v.y + v.z + v.w;
    }
    const float mean = block_sum(s, red) / (float)W;
    float ss = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
        ss += a * a + c * c + d * d + e * e;
    }
    const float rstd = rsqrtf(block_sum(ss, red) / (float)W + eps);
    bf16* o = out + (size_t)blockIdx.x * W;
    for (int i = threadIdx.x * 4; i < W; i += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float4 g = *reinterpret_cast<const float4*>(w + i);
        const float4 bb = *reinterpret_cast<const float4*>(b + i);
        uint32_t h0, l0, h1, l1;
        split_bf16x2((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y, h0, l0);
        split_bf16x2((v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w, h1, l1);
        *reinterpret_cast<uint2*>(o + i) = make_uint2(h0, h1);
        if (lo_off) *reinterpret_cast<uint2*>(o + lo_off + i) = make_uint2(l0, l1);
    }
}
int layernorm_rows_launch(cudaStream_t st, const float* x, int rows, int W, const float* w, const float* b, float eps, bf16* out, long long lo_off) {
    if (W % 4) return -1000;
    return launch_k(layernorm_rows_kernel, dim3(rows), dim3(256), 0, st, prefill_pdl(), x, W, w, b, eps, out, lo_off);
}

// -------------------------------------------------------------------------------------
// One warp per (token, vector): nh query heads, nkv key heads, nkv value heads.
// per-(token, head) quantisation of one K or V row held as e[j] = element lane + 32 j (see the QuantKvCache notes further down):
// writes codes + scale to the int page row `prow`, replaces e[] by code * scale
template <int NE>
__device__ __forceinline__ void kvq_row(float (&e)[NE], int bits, unsigned char* codes, float* scales, size_t prow, int D, int lane) {
    float am = 0.f;
#pragma unroll
    for (int j = 0; j < NE; ++j) am = fmaxf(am, fabsf(e[j]));
    am = warp_max(am);
    const float inv = bits == 8 ? (float)(1.0 / 127.0) : (float)(1.0 / 7.0);
    const float scale = __fadd_rn(__fmul_rn(am, inv), 1e-8f);
    const int off = bits == 8 ? 128 : 8;
#pragma unroll
    for (int j = 0; j < NE; ++j) {
        const int q = (int)roundf(__fdiv_rn(e[j], scale));
        e[j] = (float)q * scale;
        const int i = lane + 32 * j;
        if (bits == 8) codes[prow * D + i] = (unsigned char)(q + off);
        else {
            const int qo = __shfl_down_sync(0xffffffffu, q, 1);          // the odd neighbour's code
            if (!(lane & 1)) codes[prow * (D / 2) + (i >> 1)] = (unsigned char)((q + off) + 16 * (qo + off));
        }
    }
    if (lane == 0) scales[prow] = scale;
}

template <int D>
__global__ void __launch_bounds__(128)
rope_append_kernel(RopeAppendArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int NE = D / 32;
    const int lane = threadIdx.x & 31;
    const int nvec = a.nh + 2 * a.nkv;
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gw >= a.S * nvec) return;
    const int s = gw / nvec, vec = gw % nvec;
    const int q_dim = a.nh * D, kv_dim = a.nkv * D, q_span = a.nh * a.q_stride;
    const float* row = a.qkv + (size_t)s * (q_span + 2 * kv_dim);
    const int t = a.start_pos + s;
    const int page = a.block_table[t / KV_PAGE];
    if (vec >= a.nh + a.nkv) {   // value: cast + append
        const int kvh = vec - a.nh - a.nkv;
        const float* src = row + q_span + kv_dim + kvh * D;
        bf16* dst = a.v_pool + (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D;
        float vv[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) vv[j] = src[lane + 32 * j];
        if (a.kv_bits) kvq_row<NE>(vv, a.kv_bits, a.v_codes, a.v_scale, ((size_t)a.code_bt[t / KV_PAGE] * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE), D, lane);
#pragma unroll
        for (int j = 0; j < NE; ++j) {
            const float v = vv[j];
            const bf16 h = __float2bfloat16_rn(v);
            dst[lane + 32 * j] = h;
            if (a.kv_lo_off) dst[a.kv_lo_off + lane + 32 * j] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        return;
    }
    const bool is_k = vec >= a.nh;
    const int head = is_k ? vec - a.nh : vec;
    const float* src = is_k ? row + q_span + head * D : row + head * a.q_stride;
    const float* nw = is_k ? a.k_norm_w : a.q_norm_w;
    float e[NE];
    float ssq = 0.f;
#pragma unroll
    for (int j = 0; j < NE; ++j) { e[j] = src[lane + 32 * j]; ssq += e[j] * e[j]; }
    if (nw != nullptr) {
        ssq = warp_sum(ssq);
        const float rstd = rsqrtf(ssq / (float)D + a.eps);
#pragma unroll
        for (int j = 0; j < NE; ++j) e[j] = e[j] * rstd * nw[lane + 32 * j];
    }
    bf16* dst = is_k ? a.k_pool + (((size_t)page * a.nkv + head) * KV_PAGE + (t % KV_PAGE)) * D
                     : a.q_out + (size_t)s * q_dim + head * D;
    const int RJ = a.rot_half >> 5;
    float rr[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
        float r = e[j];
        if (j < 2 * RJ) {
            const bool lo = j < RJ;
            const int i = lane + 32 * (lo ? j : j - RJ);
            const int p = a.pos3[a.axis_of[i] * a.S + s];
            const float c = a.cos_tab[(size_t)p * a.rot_half + i], sn = a.sin_tab[(size_t)p * a.rot_half + i];
            float other = 0.f;
#pragma unroll
            for (int jj = 0; jj < NE; ++jj) if (jj == (lo ? j + RJ : j - RJ)) other = e[jj];
            r = lo ? (e[j] * c - other * sn) : (other * sn + e[j] * c);
        }
        rr[j] = r;
    }
    if (is_k && a.kv_bits) kvq_row<NE>(rr, a.kv_bits, a.k_codes, a.k_scale, ((size_t)a.code_bt[t / KV_PAGE] * a.nkv + head) * KV_PAGE + (t % KV_PAGE), D, lane);
#pragma unroll
    for (int j = 0; j < NE; ++j) {
        const float r = rr[j];
        const bf16 h = __float2bfloat16_rn(r);
        dst[lane + 32 * j] = h;
        const long long lo_off = is_k ? a.kv_lo_off : a.q_lo_off;
        if (lo_off) dst[lo_off + lane + 32 * j] = __float2bfloat16_rn(r - __bfloat162float(h));
    }
}
int rope_append_launch(cudaStream_t st, int D, const RopeAppendArgs& a) {
    if (a.rot_half < 32 || (a.rot_half % 32) != 0 || 2 * a.rot_half > D) return -1000;
    const int nwarps = a.S * (a.nh + 2 * a.nkv);
    const int grid = (nwarps + 3) / 4;
    if (D == 128) return launch_k(rope_append_kernel<128>, dim3(grid), dim3(128), 0, st, prefill_pdl(), a);
    if (D == 256) return launch_k(rope_append_kernel<256>, dim3(grid), dim3(128), 0, st, prefill_pdl(), a);
    if (D == 64) return launch_k(rope_append_kernel<64>, dim3(grid), dim3(128), 0, st, prefill_pdl(), a);
    return -1000;
}

// -------------------------------------------------------------------------------------
// Flash attention forward (prefill + ViT): 64 query rows x one head per CTA, 64-key tiles,
// bf16 mma.sync m16n8k16 with f32 accumulation and an online (base-2) softmax.
// Keys/values come either from KV pages (one tile == one page) or from a strided buffer.
// -------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// SPLIT: q, k, v (and the probabilities) are hi + lo bf16 pairs; every product keeps its three leading terms
//   S = qh.kh + qh.kl + ql.kh          O += ph.vh + ph.vl + pl.vh
// so the result carries ~16 mantissa bits instead of 8 (the f32-oracle parity mode; 3x the tensor work of 4 % of the flops).
template <int D, bool CAUSAL, bool PAGED, bool SPLIT>
__global__ void __launch_bounds__(256, D == 64 ? 2 : 1)
flash_prefill_kernel(FlashArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int BM = 64;
    constexpr int BN = (SPLIT || D > 128) ? 32 : 64;     // half-page key tiles wherever a 64-key tile x 2 groups x 2 stages would not fit
    constexpr int STG = (SPLIT && D > 128) ? 1 : 2;      // stages per group (D = 256 with hi + lo planes: one)
    constexpr int LDS = D + 8;                 // padded row (elements): conflict-free ldmatrix
    constexpr int TILE = BN * LDS;             // elements per K or V tile plane
    constexpr int CPR = D / 8;                 // 16-byte chunks per row
    constexpr int P = SPLIT ? 2 : 1;           // planes per operand
    extern __shared__ __align__(16) unsigned char fsm[];
    bf16* q_s = reinterpret_cast<bf16*>(fsm);  // [P][BM x LDS]
    bf16* kv_s = q_s + P * BM * LDS;           // [2 groups][STG][K hi | K lo | V hi | V lo]

    // Two groups of four warps work on the SAME 64 query rows: group 0 takes the even key tiles, group 1 the odd ones, each through
    // its own double-buffered tile ring, and the two online-softmax states are merged at the end.
    // A causal CTA's critical path -- q tile i needs i + 1 key tiles, the last one 8 at 454 rows -- is halved that way.
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 7, gtid = tid & 127, warp = (tid >> 5) & 3;
    const int g = lane >> 2, tq = lane & 3;
    const int head = blockIdx.y, z = blockIdx.z;
    const int row0 = a.seq_start ? a.seq_start[z] : 0;
    const int S = a.seq_len ? a.seq_len[z] : a.S;
    const int q0 = blockIdx.x * BM;
    if (q0 >= S) return;
    const int kv_off = CAUSAL ? a.kv_offset : 0;
    const int T = CAUSAL ? (kv_off + S) : S;   // keys available
    const int kvh = head / (a.nh / a.nkv);
    const int n_tiles = CAUSAL ? min((T + BN - 1) / BN, (kv_off + q0 + BM - 1) / BN + 1) : (T + BN - 1) / BN;

    // ---- Q tile -> smem (zero-fill rows past S) ----
    for (int c = tid; c < BM * CPR; c += 256) {
        const int r = c / CPR, ch = c % CPR;
        const int qr = q0 + r;
        const bf16* src = a.q + (size_t)(row0 + min(qr, S - 1)) * a.q_stride + head * D + ch * 8;
        cp_async16(smem_u32(q_s + r * LDS + ch * 8), src, qr < S ? 16 : 0);
        if (SPLIT) cp_async16(smem_u32(q_s + BM * LDS + r * LDS + ch * 8), src + a.q_lo_off, qr < S ? 16 : 0);
    }
    auto load_kv = [&](int tile, int stage) {
        bf16* ks = kv_s + (grp * STG + stage) * 2 * P * TILE;
        bf16* vs = ks + P * TILE;
        const int kv0 = tile * BN;
        for (int c = gtid; c < BN * CPR; c += 128) {
            const int r = c / CPR, ch = c % CPR;
            const int t = kv0 + r;
            const bool ok = t < T;
            const bf16 *ksrc, *vsrc;
            if (PAGED) {
                const int page = a.block_table[kv0 / KV_PAGE];     // a tile is a whole page or an aligned half of one
                const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (kv0 % KV_PAGE) + r) * D + ch * 8;
                ksrc = a.k_pool + off; vsrc = a.v_pool + off;
            } else {
                const size_t off = (size_t)(row0 + (ok ? t : 0)) * a.kv_stride + kvh * D + ch * 8;
                ksrc = a.k + off; vsrc = a.v + off;
            }
            cp_async16(smem_u32(ks + r * LDS + ch * 8), ksrc, ok ? 16 : 0);
            cp_async16(smem_u32(vs + r * LDS + ch * 8), vsrc, ok ? 16 : 0);
            if (SPLIT) {
                cp_async16(smem_u32(ks + TILE + r * LDS + ch * 8), ksrc + a.kv_lo_off, ok ? 16 : 0);
                cp_async16(smem_u32(vs + TILE + r * LDS + ch * 8), vsrc + a.kv_lo_off, ok ? 16 : 0);
            }
        }
    };
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();                           // the Q tile is complete for both groups
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };

    float o_acc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_row[2] = {-INFINITY, -INFINITY}, l_row[2] = {0.f, 0.f};
    const float sl2 = a.scale * 1.4426950408889634f;
    uint32_t qf[SPLIT ? 1 : D / 16][4];        // plain mode keeps the Q fragments in registers; SPLIT re-reads them from smem

    bool first = true;
    if (grp < n_tiles) load_kv(grp, 0);
    cp_async_commit();
    int it = 0;
    for (int tile = grp; tile < n_tiles; tile += 2, ++it) {
        const int stage = STG == 2 ? (it & 1) : 0;
        if (STG == 2) {
            if (tile + 2 < n_tiles) load_kv(tile + 2, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        group_sync();
        if (!SPLIT && first) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qf[SPLIT ? 0 : kk][0], qf[SPLIT ? 0 : kk][1], qf[SPLIT ? 0 : kk][2], qf[SPLIT ? 0 : kk][3], smem_u32(q_s + r * LDS + c));
            }
        }
        const bf16* ks = kv_s + (grp * STG + stage) * 2 * P * TILE;
        const bf16* vs = ks + P * TILE;
        // ---- S = Q K^T ----
        float s_acc[BN / 8][4];
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) { s_acc[j][0] = s_acc[j][1] = s_acc[j][2] = s_acc[j][3] = 0.f; }
        if constexpr (!SPLIT) {
#pragma unroll
            for (int j = 0; j < BN / 8; ++j) {
#pragma unroll
                for (int kk = 0; kk < D / 16; kk += 2) {
                    uint32_t b0, b1, b2, b3;
                    const int r = j * 8 + (lane & 7);
                    const int c = kk * 16 + (lane >> 3) * 8;
                    ldmatrix_x4(b0, b1, b2, b3, smem_u32(ks + r * LDS + c));
                    mma_bf16_16816(s_acc[j], qf[SPLIT ? 0 : kk], b0, b1);
                    mma_bf16_16816(s_acc[j], qf[SPLIT ? 0 : kk + 1], b2, b3);
                }
            }
        } else {
#pragma unroll 1
            for (int kk = 0; kk < D / 16; kk += 2) {
                uint32_t qh0[4], qh1[4], ql0[4], ql1[4];
                {
                    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int c = kk * 16 + (lane >> 4) * 8;
                    ldmatrix_x4(qh0[0], qh0[1], qh0[2], qh0[3], smem_u32(q_s + r * LDS + c));
                    ldmatrix_x4(qh1[0], qh1[1], qh1[2], qh1[3], smem_u32(q_s + r * LDS + c + 16));
                    ldmatrix_x4(ql0[0], ql0[1], ql0[2], ql0[3], smem_u32(q_s + BM * LDS + r * LDS + c));
                    ldmatrix_x4(ql1[0], ql1[1], ql1[2], ql1[3], smem_u32(q_s + BM * LDS + r * LDS + c + 16));
                }
#pragma unroll
                for (int j = 0; j < BN / 8; ++j) {
                    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                    const int r = j * 8 + (lane & 7);
                    const int c = kk * 16 + (lane >> 3) * 8;
                    ldmatrix_x4(h0, h1, h2, h3, smem_u32(ks + r * LDS + c));
                    ldmatrix_x4(l0, l1, l2, l3, smem_u32(ks + TILE + r * LDS + c));
                    mma_bf16_16816(s_acc[j], ql0, h0, h1);      // small terms first
                    mma_bf16_16816(s_acc[j], ql1, h2, h3);
                    mma_bf16_16816(s_acc[j], qh0, l0, l1);
                    mma_bf16_16816(s_acc[j], qh1, l2, l3);
                    mma_bf16_16816(s_acc[j], qh0, h0, h1);
                    mma_bf16_16816(s_acc[j], qh1, h2, h3);
                }
            }
        }
        // ---- mask + online softmax (rows g and g+8 of this warp's 16) ----
        const int kv0 = tile * BN;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = kv0 + j * 8 + tq * 2 + (e & 1);
                const int qrow = q0 + warp * 16 + g + (e >> 1) * 8;
                const bool vis = (col < T) && (!CAUSAL || col <= kv_off + qrow);
                const float v = vis ? s_acc[j][e] * sl2 : -INFINITY;
                s_acc[j][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float corr[2], muse[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mn = fmaxf(m_row[r], mx[r]);
            muse[r] = (mn == -INFINITY) ? 0.f : mn;
            corr[r] = exp2f(m_row[r] - muse[r]);      // m_row = -inf -> 0
            m_row[r] = mn;
            l_row[r] *= corr[r];
        }
        float ps[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < BN / 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = exp2f(s_acc[j][e] - muse[e >> 1]);
                s_acc[j][e] = p;
                ps[e >> 1] += p;
            }
        }
        l_row[0] += ps[0];
        l_row[1] += ps[1];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
            o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
        }
        // ---- O += P V ----
#pragma unroll
        for (int kt = 0; kt < BN / 16; ++kt) {
            uint32_t pa[4], pl[4];
            split_bf16x2(s_acc[2 * kt][0], s_acc[2 * kt][1], pa[0], pl[0]);
            split_bf16x2(s_acc[2 * kt][2], s_acc[2 * kt][3], pa[1], pl[1]);
            split_bf16x2(s_acc[2 * kt + 1][0], s_acc[2 * kt + 1][1], pa[2], pl[2]);
            split_bf16x2(s_acc[2 * kt + 1][2], s_acc[2 * kt + 1][3], pa[3], pl[3]);
#pragma unroll
            for (int dj = 0; dj < D / 8; dj += 2) {
                uint32_t b0, b1, b2, b3;
                const int r = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = dj * 8 + (lane >> 4) * 8;
                ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(vs + r * LDS + c));
                if constexpr (SPLIT) {
                    uint32_t c0, c1, c2, c3;
                    ldmatrix_x4_trans(c0, c1, c2, c3, smem_u32(vs + TILE + r * LDS + c));
                    mma_bf16_16816(o_acc[dj], pl, b0, b1);
                    mma_bf16_16816(o_acc[dj + 1], pl, b2, b3);
                    mma_bf16_16816(o_acc[dj], pa, c0, c1);
                    mma_bf16_16816(o_acc[dj + 1], pa, c2, c3);
                }
                mma_bf16_16816(o_acc[dj], pa, b0, b1);
                mma_bf16_16816(o_acc[dj + 1], pa, b2, b3);
            }
        }
        first = false;
        group_sync();      // the group is done with this stage before it is refilled
        if (STG == 1 && tile + 2 < n_tiles) { load_kv(tile + 2, 0); cp_async_commit(); }
    }
    cp_async_wait<0>();
    // ---- merge the two groups' states (group 1 -> group 0 through its own, now idle, tile buffer) ----
    {
        float* mg = reinterpret_cast<float*>(kv_s + STG * 2 * P * TILE);        // group 1's ring: [128 threads][D / 2 + 4] floats
        constexpr int MW = D / 2 + 4;
        __syncthreads();
        if (grp == 1) {
            float* d = mg + (size_t)gtid * MW;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) { d[4 * i] = o_acc[i][0]; d[4 * i + 1] = o_acc[i][1]; d[4 * i + 2] = o_acc[i][2]; d[4 * i + 3] = o_acc[i][3]; }
            d[D / 2] = m_row[0]; d[D / 2 + 1] = m_row[1]; d[D / 2 + 2] = l_row[0]; d[D / 2 + 3] = l_row[1];
        }
        __syncthreads();
        if (grp == 1) return;
        const float* d = mg + (size_t)gtid * MW;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float mo = d[D / 2 + r], lo = d[D / 2 + 2 + r];
            const float mn = fmaxf(m_row[r], mo);
            const float c0 = (m_row[r] == -INFINITY) ? 0.f : exp2f(m_row[r] - mn), c1 = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
            l_row[r] = l_row[r] * c0 + lo * c1;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                o_acc[i][2 * r] = o_acc[i][2 * r] * c0 + d[4 * i + 2 * r] * c1;
                o_acc[i][2 * r + 1] = o_acc[i][2 * r + 1] * c0 + d[4 * i + 2 * r + 1] * c1;
            }
        }
    }
    // ---- normalise and store ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 1);
        l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qrow = q0 + warp * 16 + g + r * 8;
        if (qrow < S) {
            const float inv = 1.f / l_row[r];
            bf16* o = a.out + (size_t)(row0 + qrow) * a.o_stride + head * D;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                uint32_t hi, lo;
                split_bf16x2(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv, hi, lo);
                *reinterpret_cast<uint32_t*>(o + i * 8 + tq * 2) = hi;
                if (SPLIT) *reinterpret_cast<uint32_t*>(o + a.out_lo_off + i * 8 + tq * 2) = lo;
            }
        }
    }
}

template <int D, bool CAUSAL, bool PAGED, bool SPLIT>
static int flash_launch_t(cudaStream_t st, const FlashArgs& a) {
    constexpr int P = SPLIT ? 2 : 1, GROUPS = 2, BN = (SPLIT || D > 128) ? 32 : 64, STG = (SPLIT && D > 128) ? 1 : 2;
    constexpr int SMEM = (P * 64 * (D + 8) + GROUPS * STG * 2 * P * BN * (D + 8)) * 2;
    static_assert(STG * 2 * P * BN * (D + 8) * 2 >= 128 * (D / 2 + 4) * 4, "the merge buffer lives in group 1's ring");
    static_assert(SMEM <= 227 * 1024, "flash tile does not fit");
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(flash_prefill_kernel<D, CAUSAL, PAGED, SPLIT>, (size_t)SMEM, seen)) return e;
    const int max_len = a.seq_len ? a.max_len : a.S;
    dim3 grid((max_len + 63) / 64, a.nh, a.seq_len ? a.nseq : 1);
    return launch_k(flash_prefill_kernel<D, CAUSAL, PAGED, SPLIT>, grid, dim3(256), SMEM, st, prefill_pdl(), a);
}

int flash_prefill_launch(cudaStream_t st, int D, bool causal, bool paged, const FlashArgs& a) {
    const bool split = a.q_lo_off != 0;
    if (split && (a.kv_lo_off == 0 || a.out_lo_off == 0)) return -1000;
    if (D == 128 && causal && paged) return split ? flash_launch_t<128, true, true, true>(st, a) : flash_launch_t<128, true, true, false>(st, a);
    if (D == 256 && causal && paged) return split ? flash_launch_t<256, true, true, true>(st, a) : flash_launch_t<256, true, true, false>(st, a);
    if (D == 64 && !causal && !paged) return split ? flash_launch_t<64, false, false, true>(st, a) : flash_launch_t<64, false, false, false>(st, a);
    if (D == 128 && !causal && !paged) return split ? flash_launch_t<128, false, false, true>(st, a) : flash_launch_t<128, false, false, false>(st, a);
    return -1000;
}

// attn[s, h*D + i] *= sigmoid(gate), gate = qkv[s, h*q_stride + D + i]   (qwen3_5/modeling.rs:556-563)
__global__ void __launch_bounds__(256)
gate_mul_kernel(bf16* __restrict__ attn, const float* __restrict__ qkv, int nh, int D, int q_stride, int row_width, long long lo_off) {
    pdl_wait();
    pdl_launch_dependents();
    const int s = blockIdx.x;
    for (int i = threadIdx.x; i < nh * D; i += blockDim.x) {
        const int h = i / D, d = i % D;
        const float g = qkv[(size_t)s * row_width + h * q_stride + D + d];
        bf16* p = attn + (size_t)s * nh * D + i;
        float v = __bfloat162float(*p);
        if (lo_off) v += __bfloat162float(p[lo_off]);
        v = v / (1.0f + expf(-g));
        const bf16 hi = __float2bfloat16_rn(v);
        *p = hi;
        if (lo_off) p[lo_off] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}
int gate_mul_launch(cudaStream_t st, bf16* attn, const float* qkv, int S, int nh, int D, int q_stride, int row_width, long long lo_off) {
    return launch_k(gate_mul_kernel, dim3(S), dim3(256), 0, st, prefill_pdl(), attn, qkv, nh, D, q_stride, row_width, lo_off);
}

// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
set_rows_kernel(float* __restrict__ x, int H, const int* __restrict__ rows, const float* __restrict__ src, int add) {
    pdl_wait();
    pdl_launch_dependents();
    float* dst = x + (size_t)rows[blockIdx.x] * H;
    const float* s = src + (size_t)blockIdx.x * H;
    for (int i = threadIdx.x * 4; i < H; i += 1024) {
        float4 v = *reinterpret_cast<const float4*>(s + i);
        if (add) {
            const float4 d = *reinterpret_cast<const float4*>(dst + i);
            v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
        }
        *reinterpret_cast<float4*>(dst + i) = v;
    }
}
int set_rows_launch(cudaStream_t st, float* x, int H, const int* rows, int n, const float* src, bool add) {
    if (n <= 0) return 0;
    return launch_k(set_rows_kernel, dim3(n), dim3(256), 0, st, prefill_pdl(), x, H, rows, src, add ? 1 : 0);
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4, long long lo_off) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v.x, v.y, h0, l0);
        split_bf16x2(v.z, v.w, h1, l1);
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(h0, h1);
        if (lo_off) reinterpret_cast<uint2*>(dst + lo_off)[i] = make_uint2(l0, l1);
    }
}
int cast_f32_bf16_launch(cudaStream_t st, const float* src, bf16* dst, size_t n, long long lo_off) {
    if (n % 4) return -1000;
    const size_t n4 = n / 4;
    const int grid = (int)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
    return launch_k(cast_f32_bf16_kernel, dim3(grid), dim3(256), 0, st, prefill_pdl(), src, dst, n4, lo_off);
}

// Quantised KV pages (QuantKvCache, crane-core/src/models/qwen3_5/kv_cache.rs:209-342): per (token, KV head) symmetric int8 / int4
// codes + one f32 scale.  scale = amax * f32(1 / qmax) + 1e-8, code = round_half_away(x / scale) + offset (128 / 8); int4 packs
// (even, odd) elements as lo + 16 * hi.  What attention sees is (code - offset) * scale.
// The prefill side works through a bf16 (hi + lo) scratch in page layout (one sequence, one layer at a time, identity block table):
//   kv_dequant_pages: the cached prefix [0, T) of the sequence -> scratch       (the reference dequantises the whole cache per step)
//   rope_append writes the new rows into the scratch as for a lossless cache
//   kv_quant_rows: new rows [t0, t0 + S) scratch -> codes + scales in the int pages, and the scratch rows are replaced by their
//                  dequantised values, so the flash kernel reads exactly what the reference's attention reads.
__device__ __forceinline__ float kvq_scale(float amax, int bits) {
    const float inv = bits == 8 ? (float)(1.0 / 127.0) : (float)(1.0 / 7.0);
    return __fadd_rn(__fmul_rn(amax, inv), 1e-8f);
}
__global__ void __launch_bounds__(128)
kv_quant_rows_kernel(KvQuantArgs a, int t0) {
    __shared__ float red[4];
    const int t = t0 + blockIdx.x, h = blockIdx.y, which = blockIdx.z, tid = threadIdx.x;
    bf16* sp = which ? a.sv : a.sk;
    const size_t srow = (((size_t)(t / KV_PAGE) * a.nkv + h) * KV_PAGE + (t % KV_PAGE)) * a.D;
    const size_t prow = ((size_t)a.bt[t / KV_PAGE] * a.nkv + h) * KV_PAGE + (t % KV_PAGE);
    float x[2];
    float am = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid * 2 + j;                   // (even, odd) pair per thread: D <= 256
        x[j] = 0.f;
        if (i < a.D) {
            x[j] = __bfloat162float(sp[srow + i]);
            if (a.s_lo) x[j] += __bfloat162float(sp[a.s_lo + srow + i]);
        }
        am = fmaxf(am, fabsf(x[j]));
    }
    am = warp_max(am);
    if ((tid & 31) == 0) red[tid >> 5] = am;
    __syncthreads();
    am = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const float scale = kvq_scale(am, a.bits);
    const int off = a.bits == 8 ? 128 : 8;
    int q[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) q[j] = (int)roundf(__fdiv_rn(x[j], scale));
    if (tid * 2 < a.D) {
        unsigned char* codes = which ? a.v_codes : a.k_codes;
        if (a.bits == 8) {
            codes[prow * a.D + tid * 2] = (unsigned char)(q[0] + off);
            codes[prow * a.D + tid * 2 + 1] = (unsigned char)(q[1] + off);
        } else {
            codes[prow * (a.D / 2) + tid] = (unsigned char)((q[0] + off) + 16 * (q[1] + off));
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float y = (float)q[j] * scale;
            const bf16 hi = __float2bfloat16_rn(y);
            sp[srow + tid * 2 + j] = hi;
            if (a.s_lo) sp[a.s_lo + srow + tid * 2 + j] = __float2bfloat16_rn(y - __bfloat162float(hi));
        }
    }
    if (tid == 0) (which ? a.v_scale : a.k_scale)[prow] = scale;
}
__global__ void __launch_bounds__(128)
kv_dequant_pages_kernel(KvQuantArgs a) {
    const int t = blockIdx.x, h = blockIdx.y, which = blockIdx.z, tid = threadIdx.x;
    bf16* sp = which ? a.sv : a.sk;
    const size_t srow = (((size_t)(t / KV_PAGE) * a.nkv + h) * KV_PAGE + (t % KV_PAGE)) * a.D;
    const size_t prow = ((size_t)a.bt[t / KV_PAGE] * a.nkv + h) * KV_PAGE + (t % KV_PAGE);
    const unsigned char* codes = which ? a.v_codes : a.k_codes;
    const float scale = (which ? a.v_scale : a.k_scale)[prow];
    const int off = a.bits == 8 ? 128 : 8;
    if (tid * 2 >= a.D) return;
    int q[2];
    if (a.bits == 8) { q[0] = codes[prow * a.D + tid * 2]; q[1] = codes[prow * a.D + tid * 2 + 1]; }
    else { const int b = codes[prow * (a.D / 2) + tid]; q[0] = b & 15; q[1] = b >> 4; }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float y = (float)(q[j] - off) * scale;
        const bf16 hi = __float2bfloat16_rn(y);
        sp[srow + tid * 2 + j] = hi;
        if (a.s_lo) sp[a.s_lo + srow + tid * 2 + j] = __float2bfloat16_rn(y - __bfloat162float(hi));
    }
}
int kv_quant_rows_launch(cudaStream_t st, const KvQuantArgs& a, int t0, int S) {
    if (S <= 0) return 0;
    if (a.D > 256 || (a.D & 1) || (a.bits != 8 && a.bits != 4)) return -1000;
    kv_quant_rows_kernel<<<dim3(S, a.nkv, 2), 128, 0, st>>>(a, t0);
    return (int)cudaGetLastError();
}
int kv_dequant_pages_launch(cudaStream_t st, const KvQuantArgs& a, int T) {
    if (T <= 0) return 0;
    if (a.D > 256 || (a.D & 1) || (a.bits != 8 && a.bits != 4)) return -1000;
    kv_dequant_pages_kernel<<<dim3(T, a.nkv, 2), 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}

// KV swap (ModelBackend::get_kv_caches / set_kv_caches, crane-serve/src/engine/backend.rs:65-84): one layer's pages of one sequence
// <-> the reference's contiguous cache tensors [n_kv, T, D] (f32 here: the two bf16 planes of the split mode summed / re-split).
__global__ void __launch_bounds__(128)
kv_pages_to_rows_kernel(const bf16* __restrict__ pool, long long lo_off, const int* __restrict__ bt, int nkv, int D, int T, float* __restrict__ out) {
    const int t = blockIdx.x, h = blockIdx.y;
    const size_t off = (((size_t)bt[t / KV_PAGE] * nkv + h) * KV_PAGE + (t % KV_PAGE)) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        float v = __bfloat162float(pool[off + i]);
        if (lo_off) v += __bfloat162float(pool[lo_off + off + i]);
        out[((size_t)h * T + t) * D + i] = v;
    }
}
__global__ void __launch_bounds__(128)
kv_rows_to_pages_kernel(bf16* __restrict__ pool, long long lo_off, const int* __restrict__ bt, int nkv, int D, int T, const float* __restrict__ in) {
    const int t = blockIdx.x, h = blockIdx.y;
    const size_t off = (((size_t)bt[t / KV_PAGE] * nkv + h) * KV_PAGE + (t % KV_PAGE)) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float v = in[((size_t)h * T + t) * D + i];
        const bf16 hi = __float2bfloat16_rn(v);
        pool[off + i] = hi;
        if (lo_off) pool[lo_off + off + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}
int kv_pages_to_rows_launch(cudaStream_t st, const bf16* pool, long long lo_off, const int* bt, int nkv, int D, int T, float* out) {
    if (T <= 0) return 0;
    kv_pages_to_rows_kernel<<<dim3(T, nkv), 128, 0, st>>>(pool, lo_off, bt, nkv, D, T, out);
    return (int)cudaGetLastError();
}
int kv_rows_to_pages_launch(cudaStream_t st, bf16* pool, long long lo_off, const int* bt, int nkv, int D, int T, const float* in) {
    if (T <= 0) return 0;
    kv_rows_to_pages_kernel<<<dim3(T, nkv), 128, 0, st>>>(pool, lo_off, bt, nkv, D, T, in);
    return (int)cudaGetLastError();
}

// split-precision planes (hi [+ lo]) -> f32: the input of a quantised linear that follows a tensor-core kernel
__global__ void __launch_bounds__(256)
planes_to_f32_kernel(const bf16* __restrict__ src, long long lo_off, size_t n4, float* __restrict__ dst) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 h = reinterpret_cast<const uint2*>(src)[i];
        float4 v = make_float4(bf16lo(h.x), bf16hi(h.x), bf16lo(h.y), bf16hi(h.y));
        if (lo_off) {
            const uint2 l = reinterpret_cast<const uint2*>(src + lo_off)[i];
            v.x += bf16lo(l.x); v.y += bf16hi(l.x); v.z += bf16lo(l.y); v.w += bf16hi(l.y);
        }
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}
int planes_to_f32_launch(cudaStream_t st, const bf16* src, long long lo_off, size_t n, float* dst) {
    if (n % 4 || (lo_off % 4)) return -1000;
    const size_t n4 = n / 4;
    const int grid = (int)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
    return launch_k(planes_to_f32_kernel, dim3(grid), dim3(256), 0, st, prefill_pdl(), src, lo_off, n4, dst);
}

// x[p] += sum_c w4[c][p] * table[idx4[c][p]]      (bilinear pos-embed, qwen3_5/vision.rs:445-459)
__global__ void __launch_bounds__(256)
vit_pos_embed_add_kernel(float* __restrict__ x, int N, int Hv, const float* __restrict__ table,
                         const int* __restrict__ idx4, const float* __restrict__ w4) {
    pdl_wait();
    pdl_launch_dependents();
    const int p = blockIdx.x;
    int id[4]; float w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { id[c] = idx4[c * N + p]; w[c] = w4[c * N + p]; }
    for (int i = threadIdx.x; i < Hv; i += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += table[(size_t)id[c] * Hv + i] * w[c];
        x[(size_t)p * Hv + i] += acc;
    }
}
int vit_pos_embed_add_launch(cudaStream_t st, float* x, int N, int Hv, const float* table, const int* idx4, const float* w4) {
    return launch_k(vit_pos_embed_add_kernel, dim3(N), dim3(256), 0, st, prefill_pdl(), x, N, Hv, table, idx4, w4);
}

// qkv f32 [N, 3, nh, hd] -> bf16 same layout; q and k rotated in f32 with the 2-D (row, col) table
// cos/sin [N, hd/2] (full-width cat(emb, emb) of the reference collapses to a half-split rotation).
__global__ void __launch_bounds__(256)
vit_rope_kernel(const float* __restrict__ qkv, int nh, int hd, const float* __restrict__ cs, const float* __restrict__ sn,
                bf16* __restrict__ out, long long lo_off) {
    pdl_wait();
    pdl_launch_dependents();
    auto put = [&](bf16* p, float v) {
        const bf16 h = __float2bfloat16_rn(v);
        *p = h;
        if (lo_off) p[lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
    };
    const int p = blockIdx.x;
    const int half = hd / 2;
    const int Hv = nh * hd;
    const float* row = qkv + (size_t)p * 3 * Hv;
    bf16* orow = out + (size_t)p * 3 * Hv;
    for (int i = threadIdx.x; i < 2 * nh * half; i += blockDim.x) {   // q and k pairs
        const int which = i / (nh * half), rem = i % (nh * half);
        const int h = rem / half, j = rem % half;
        const float c = cs[(size_t)p * half + j], s = sn[(size_t)p * half + j];
        const float* v = row + which * Hv + h * hd;
        const float x1 = v[j], x2 = v[j + half];
        bf16* o = orow + which * Hv + h * hd;
        put(o + j, x1 * c - x2 * s);
        put(o + j + half, x2 * c + x1 * s);
    }
    for (int i = threadIdx.x; i < Hv; i += blockDim.x) put(orow + 2 * Hv + i, row[2 * Hv + i]);
}
int vit_rope_launch(cudaStream_t st, const float* qkv, int N, int nh, int hd, const float* cos, const float* sin, bf16* out, long long lo_off) {
    return launch_k(vit_rope_kernel, dim3(N), dim3(256), 0, st, prefill_pdl(), qkv, nh, hd, cos, sin, out, lo_off);
}

}  // namespace cb
