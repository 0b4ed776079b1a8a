// Persistent decode kernel: ALL phases of n_steps greedy decode steps of one sequence in ONE launch.
//
// Why: a batch-1 decode step of Qwen3-VL-2B is 141 dependent bandwidth-bound phases of 1.3-96 us of HBM time each.
// As separate kernels every phase pays launch / drain / refill latency comparable to its streaming time (measured:
// 47 % of the HBM roofline with PDL-chained kernels).  Here one CTA per SM stays resident for the whole run:
//   * every warp owns a fixed slice of every weight matrix and streams it through a private cp.async ring
//     (PK_DEPTH x 2 KB per warp, 144 KB per SM).  The prefetch cursor runs AHEAD across phase, layer and token
//     boundaries -- while a CTA waits at a grid barrier or stages activations, the next phase's weights keep
//     arriving, so the HBM pipe does not drain between phases;
//   * phases are separated by a grid barrier (one atomic + acquire spin per CTA);
//   * attention runs on nkv x 8 CTAs (one KV head x one token range each, K/V straight from the pages), the
//     split merge is folded into the activation staging of the O-projection phase.
// Arithmetic is identical to the multi-kernel path (decode.cu): f32 activations / residual, bf16 weights and KV,
// RMSNorm folded into the GEMV, lowest-index argmax.  Reference being replaced: the per-token loop of
// `Model::generate` (crane-core/src/models/qwen3/model.rs:298-331) over `Qwen3Model::forward`
// (qwen3/modeling.rs:942-1036) and the server's decode rounds (crane-serve/src/engine/mod.rs:898-1008).
#include "decode_persistent.cuh"

namespace cb {

constexpr int PK_WARPS = PK_WARPS_C;
constexpr int PK_THREADS = PK_WARPS * 32;
#ifndef PK_DEPTH_C
#define PK_DEPTH_C 1
#endif
constexpr int PK_DEPTH = PK_DEPTH_C;
constexpr int PK_SEG_CHUNKS = 4;
constexpr int PK_SEG_BYTES = PK_SEG_CHUNKS * 512;
constexpr int PK_NSPLIT = 8;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Grid barrier: arrivals are counted on bar[0]; the last arriver publishes the generation on bar[32] (its own cache
// line), which is what everybody else polls -- arrivals never queue behind pollers at the L2 slice.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int old = atomicAdd(bar, 1u);
        if (old == target - 1) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"(target) : "memory");
        } else {
            while (ld_acquire_u32(bar + 32) < target) { __nanosleep(20); }
        }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PK_PROF(slot)                                                                   \
    do {                                                                                \
        if (a.prof && blockIdx.x == 0 && tid == 0) {                                    \
            const unsigned long long now__ = gtime_ns();                                \
            a.prof[slot] += now__ - tprof;                                              \
            tprof = now__;                                                              \
        }                                                                               \
    } while (0)
__device__ __forceinline__ void pk_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct Geom { const bf16* W; int N, K, rpu; };
__device__ __forceinline__ Geom phase_geom(const PersistArgs& a, int gp) {
    Geom g;
    if (gp >= 4 * a.L) { g.W = a.lm_head; g.N = a.V; g.K = a.H; g.rpu = 1; return g; }
    const PLayer& l = a.layers[gp >> 2];
    switch (gp & 3) {
        case 0: g.W = l.wqkv; g.N = a.qkv_dim; g.K = a.H; g.rpu = 1; break;
        case 1: g.W = l.wo; g.N = a.H; g.K = a.q_dim; g.rpu = 1; break;
        case 2: g.W = l.wgu; g.N = 2 * a.I; g.K = a.H; g.rpu = 2; break;
        default: g.W = l.wdown; g.N = a.H; g.K = a.I; g.rpu = 1; break;
    }
    return g;
}
struct Slab { const unsigned char* gsrc; uint32_t c_begin, c_end, nseg, cpr; int rpc, r0, nrows; };
__device__ __forceinline__ Slab warp_slab(const Geom& g, int warp, int lane) {
    Slab s;
    const int units = g.N / g.rpu;
    const int upc = (units + gridDim.x - 1) / gridDim.x;
    s.rpc = upc * g.rpu;
    s.r0 = blockIdx.x * s.rpc;
    s.nrows = max(0, min(g.N, s.r0 + s.rpc) - s.r0);
    s.cpr = (uint32_t)(g.K >> 8);
    const uint32_t total_chunks = (uint32_t)s.nrows * s.cpr;
    const uint32_t total_segs = (total_chunks + PK_SEG_CHUNKS - 1) / PK_SEG_CHUNKS;
    const uint32_t spw = (total_segs + PK_WARPS - 1) / PK_WARPS;
    s.c_begin = min(total_chunks, (uint32_t)warp * spw * PK_SEG_CHUNKS);
    s.c_end = min(total_chunks, ((uint32_t)warp + 1) * spw * PK_SEG_CHUNKS);
    s.nseg = (s.c_end - s.c_begin + PK_SEG_CHUNKS - 1) / PK_SEG_CHUNKS;
    s.gsrc = reinterpret_cast<const unsigned char*>(g.W) + (size_t)s.r0 * g.K * 2 + (size_t)s.c_begin * 512 + lane * 16;
    return s;
}

struct AttnShared {
    float* q_s;       // [NREP][128]
    float* knew_s;    // [128]
    float* vnew_s;    // [128]
    float* mm_s;      // [PK_WARPS][NREP]
    float* ml_s;      // [PK_WARPS][NREP]
    float* scratch;   // [PK_WARPS][NREP][128]  (the activation staging buffer, free during this phase)
};

// Attention partials of one decode step for layer `ly`: CTA (kvh, split) handles one KV head x one token range.
template <int NREP>
__device__ __noinline__ void pk_attention(const PersistArgs& a, const PLayer& ly, const AttnShared sh, int T, int p0, int p1, int p2) {
    constexpr int D = 128, EPL = 8, LPT = D / EPL, TPW = 32 / LPT, NE = D / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* bt = a.block_table;
    const int q_dim = a.q_dim, kv_dim = a.nkv * D;
    float (*q_s)[D] = reinterpret_cast<float (*)[D]>(sh.q_s);
    float* knew_s = sh.knew_s;
    float* vnew_s = sh.vnew_s;
    float (*mm_s)[NREP] = reinterpret_cast<float (*)[NREP]>(sh.mm_s);
    float (*ml_s)[NREP] = reinterpret_cast<float (*)[NREP]>(sh.ml_s);
    float* xs_f = sh.scratch;
    const int pos3[3] = {p0, p1, p2};
    {
        const int kvh = blockIdx.x / PK_NSPLIT, split = blockIdx.x % PK_NSPLIT;
        int chunk = (T + PK_NSPLIT - 1) / PK_NSPLIT;
        chunk = (chunk + 7) & ~7;
        const int t0 = split * chunk;
        const int t1 = min(T, t0 + chunk);
        const int s_last = (T - 1) / chunk;
        const int t_end = min(t1, T - 1);
        for (int vec = warp; vec < NREP + 1; vec += PK_WARPS) {
            const bool is_k = (vec == NREP);
            if (is_k && split != s_last) continue;
            const float* src = is_k ? (a.qkv + q_dim + kvh * D) : (a.qkv + (kvh * NREP + vec) * D);
            const float* nw = is_k ? ly.kn : ly.qn;
            float e[NE];
            float ssq = 0.f;
#pragma unroll
            for (int j = 0; j < NE; ++j) { e[j] = __ldcg(src + lane + 32 * j); ssq += e[j] * e[j]; }
            ssq = warp_sum(ssq);
            const float rstd = rsqrtf(ssq / (float)D + a.eps);
#pragma unroll
            for (int j = 0; j < NE; ++j) e[j] = e[j] * rstd * nw[lane + 32 * j];
            float* dst = is_k ? knew_s : q_s[vec];
#pragma unroll
            for (int j = 0; j < NE / 2; ++j) {       // full rotary, D = 128: pairs (i, i + 64)
                const int i = lane + 32 * j;
                const int p = pos3[a.axis_of[i]];
                const float c = a.cos_tab[(size_t)p * (D / 2) + i], s = a.sin_tab[(size_t)p * (D / 2) + i];
                const float x1 = e[j], x2 = e[j + NE / 2];
                float r1 = x1 * c - x2 * s, r2 = x1 * s + x2 * c;
                if (is_k) { r1 = round_bf16(r1); r2 = round_bf16(r2); }
                dst[i] = r1;
                dst[i + D / 2] = r2;
            }
        }
        if (split == s_last && warp == PK_WARPS - 1) {
            const float* vsrc = a.qkv + q_dim + kv_dim + kvh * D;
            for (int i = lane; i < D; i += 32) vnew_s[i] = round_bf16(__ldcg(vsrc + i));
        }
        __syncthreads();
        if (split == s_last) {
            const int tt = T - 1;
            const int page = bt[tt / KV_PAGE];
            const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (tt % KV_PAGE)) * D;
            for (int i = tid; i < D; i += PK_THREADS) {
                ly.k_pool[off + i] = __float2bfloat16_rn(knew_s[i]);
                ly.v_pool[off + i] = __float2bfloat16_rn(vnew_s[i]);
            }
        }
        const int grp = lane / LPT, gl = lane % LPT;
        float qr[NREP][EPL];
#pragma unroll
        for (int h = 0; h < NREP; ++h)
#pragma unroll
            for (int j = 0; j < EPL; ++j) qr[h][j] = q_s[h][gl * EPL + j] * a.scale;
        float m[NREP], lsum[NREP], o[NREP][EPL];
#pragma unroll
        for (int h = 0; h < NREP; ++h) {
            m[h] = -INFINITY; lsum[h] = 0.f;
#pragma unroll
            for (int j = 0; j < EPL; ++j) o[h][j] = 0.f;
        }
        // two tokens per group in flight: 4 x 16-byte loads per lane before the first use
        for (int tb = t0 + warp * TPW; tb < t_end; tb += 2 * PK_WARPS * TPW) {
            uint4 kr[2], vr[2];
            bool valid[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int tt = tb + u * PK_WARPS * TPW + grp;
                valid[u] = tt < t_end;
                if (valid[u]) {
                    const int page = bt[tt / KV_PAGE];
                    const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (tt % KV_PAGE)) * D + gl * EPL;
                    kr[u] = ldg_stream(ly.k_pool + off);
                    vr[u] = ldg_stream(ly.v_pool + off);
                } else { kr[u] = make_uint4(0, 0, 0, 0); vr[u] = make_uint4(0, 0, 0, 0); }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float kf[EPL], vf[EPL];
                kf[0] = bf16lo(kr[u].x); kf[1] = bf16hi(kr[u].x); kf[2] = bf16lo(kr[u].y); kf[3] = bf16hi(kr[u].y);
                kf[4] = bf16lo(kr[u].z); kf[5] = bf16hi(kr[u].z); kf[6] = bf16lo(kr[u].w); kf[7] = bf16hi(kr[u].w);
                vf[0] = bf16lo(vr[u].x); vf[1] = bf16hi(vr[u].x); vf[2] = bf16lo(vr[u].y); vf[3] = bf16hi(vr[u].y);
                vf[4] = bf16lo(vr[u].z); vf[5] = bf16hi(vr[u].z); vf[6] = bf16lo(vr[u].w); vf[7] = bf16hi(vr[u].w);
#pragma unroll
                for (int h = 0; h < NREP; ++h) {
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < EPL; ++j) s = fmaf(qr[h][j], kf[j], s);
#pragma unroll
                    for (int ofs = LPT / 2; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
                    if (valid[u]) {
                        const float mn = fmaxf(m[h], s);
                        const float corr = __expf(m[h] - mn), p = __expf(s - mn);
                        lsum[h] = lsum[h] * corr + p;
#pragma unroll
                        for (int j = 0; j < EPL; ++j) o[h][j] = o[h][j] * corr + p * vf[j];
                        m[h] = mn;
                    }
                }
            }
        }
        if (split == s_last && warp == 0 && grp == 0) {
#pragma unroll
            for (int h = 0; h < NREP; ++h) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < EPL; ++j) s = fmaf(qr[h][j], knew_s[gl * EPL + j], s);
#pragma unroll
                for (int ofs = LPT / 2; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0x0000ffffu, s, ofs);
                const float mn = fmaxf(m[h], s);
                const float corr = __expf(m[h] - mn), p = __expf(s - mn);
                lsum[h] = lsum[h] * corr + p;
#pragma unroll
                for (int j = 0; j < EPL; ++j) o[h][j] = o[h][j] * corr + p * vnew_s[gl * EPL + j];
                m[h] = mn;
            }
        }
#pragma unroll
        for (int h = 0; h < NREP; ++h) {             // merge the two token groups of a warp
            const float mo = __shfl_xor_sync(0xffffffffu, m[h], 16);
            const float lo = __shfl_xor_sync(0xffffffffu, lsum[h], 16);
            const float mn = fmaxf(m[h], mo);
            const float c0 = (m[h] == -INFINITY) ? 0.f : __expf(m[h] - mn);
            const float c1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
            lsum[h] = lsum[h] * c0 + lo * c1;
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const float oo = __shfl_xor_sync(0xffffffffu, o[h][j], 16);
                o[h][j] = o[h][j] * c0 + oo * c1;
            }
            m[h] = mn;
        }
        float* mo_s = xs_f;                           // [PK_WARPS][NREP][D] scratch (xs is free during this phase)
        if (grp == 0) {
#pragma unroll
            for (int h = 0; h < NREP; ++h) {
                if (gl == 0) { mm_s[warp][h] = m[h]; ml_s[warp][h] = lsum[h]; }
#pragma unroll
                for (int j = 0; j < EPL; ++j) mo_s[((size_t)warp * NREP + h) * D + gl * EPL + j] = o[h][j];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < NREP * D; idx += PK_THREADS) {
            const int h = idx / D, i = idx % D;
            float M = -INFINITY;
            for (int w = 0; w < PK_WARPS; ++w) M = fmaxf(M, mm_s[w][h]);
            float Ls = 0.f, O = 0.f;
            for (int w = 0; w < PK_WARPS; ++w) {
                const float c = (mm_s[w][h] == -INFINITY) ? 0.f : __expf(mm_s[w][h] - M);
                Ls += ml_s[w][h] * c;
                O += mo_s[((size_t)w * NREP + h) * D + i] * c;
            }
            const int head = kvh * NREP + h;
            const size_t pbase = (size_t)head * PK_NSPLIT + split;
            a.part_o[pbase * D + i] = O;
            if (i == 0) { a.part_ml[pbase * 2 + 0] = M; a.part_ml[pbase * 2 + 1] = Ls; }
        }
    }
}

template <int NREP>
__global__ void __launch_bounds__(PK_THREADS, 1)
decode_persistent_kernel(PersistArgs a) {
    constexpr int D = 128;
    extern __shared__ __align__(1024) unsigned char psm[];
    // layout: [rings: PK_WARPS * PK_DEPTH * 2 KB][xs: staging f32 (also attention merge scratch)][acc: max rows/CTA f32]
    float4* xs = reinterpret_cast<float4*>(psm + PK_WARPS * PK_DEPTH * PK_SEG_BYTES);
    float* xs_f = reinterpret_cast<float*>(xs);
    float* acc_s = xs_f + a.xs_floats;
    __shared__ float red[32];
    __shared__ float rstd_s;
    __shared__ float wbest_v[PK_WARPS];
    __shared__ int wbest_i[PK_WARPS];
    __shared__ float q_s[NREP][D];
    __shared__ float knew_s[D], vnew_s[D];
    __shared__ float mm_s[PK_WARPS][NREP], ml_s[PK_WARPS][NREP];
    __shared__ uint32_t tok_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_gp = 4 * a.L + 1;                       // GEMV phases per token
    unsigned int bar_target = 0;
    const unsigned int nct = gridDim.x;
    AttnShared ash{&q_s[0][0], knew_s, vnew_s, &mm_s[0][0], &ml_s[0][0], xs_f};

    // ---------------- prefetch cursor (runs ahead of consumption across phases and tokens) ----------------
    unsigned char* ring = psm + (size_t)warp * PK_DEPTH * PK_SEG_BYTES + lane * 16;
    const uint32_t ring_u32 = smem_u32(ring);
    int pf_tok = 0, pf_gp = 0;
    uint32_t pf_g = 0, pf_seq = 0;
    Slab pf = warp_slab(phase_geom(a, 0), warp, lane);
    auto issue_next = [&]() {
        while (pf_tok < a.n_steps && pf_g >= pf.nseg) {
            pf_g = 0;
            if (++pf_gp == n_gp) { pf_gp = 0; ++pf_tok; }
            if (pf_tok < a.n_steps) pf = warp_slab(phase_geom(a, pf_gp), warp, lane);
        }
        if (pf_tok < a.n_steps) {
            const uint32_t cb = pf_g * PK_SEG_CHUNKS;
            const uint32_t slot = pf_seq % PK_DEPTH;
#pragma unroll
            for (int c = 0; c < PK_SEG_CHUNKS; ++c)
                if (pf.c_begin + cb + c < pf.c_end)
                    pk_cp_async16(ring_u32 + slot * PK_SEG_BYTES + c * 512, pf.gsrc + (size_t)(cb + c) * 512);
            ++pf_g;
            ++pf_seq;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int i = 0; i < PK_DEPTH; ++i) issue_next();
    uint32_t c_seq = 0;                                  // segments consumed so far (ring slot = c_seq % PK_DEPTH)

    const SeqState st0 = a.state[0];
    const int q_dim = a.q_dim;
    unsigned long long tprof = gtime_ns();

#pragma unroll 1
    for (int t = 0; t < a.n_steps; ++t) {
        const int T = st0.kv_len + t + 1;
        float bestv = -INFINITY;
        int besti = 0x7fffffff;
#pragma unroll 1
        for (int gp = 0; gp < n_gp; ++gp) {
            const int kind = (gp == 4 * a.L) ? 4 : (gp & 3);      // 0 qkv, 1 o-proj, 2 gate/up, 3 down, 4 lm_head
            const PLayer& ly = a.layers[min(gp >> 2, a.L - 1)];
            if (kind == 1) {
                // ---- attention partials (nkv x PK_NSPLIT CTAs), then the grid barrier that publishes them ----
                if ((int)blockIdx.x < a.nkv * PK_NSPLIT) pk_attention<NREP>(a, ly, ash, T, st0.pos[0] + t, st0.pos[1] + t, st0.pos[2] + t);
                PK_PROF(0);
                bar_target += nct; grid_barrier(a.barrier, bar_target);
                PK_PROF(1);
            }
            const Geom g = phase_geom(a, gp);
            const Slab sl = warp_slab(g, warp, lane);
            const int K8 = g.K >> 3;
            PK_PROF(7);
            // ---- activation staging: xs <- input (* RMSNorm weight), 1/rms ----
            const float* nw = (kind == 0) ? ly.ln1 : (kind == 2) ? ly.ln2 : (kind == 4) ? a.final_norm : nullptr;
            if (kind == 1) {
                // split merge folded into the staging of the O-projection input
                for (int idx = tid; idx < q_dim; idx += PK_THREADS) {
                    const int head = idx / D, i = idx % D;
                    const size_t pbase = (size_t)head * PK_NSPLIT;
                    float ms[PK_NSPLIT];
                    float M = -INFINITY;
#pragma unroll
                    for (int s = 0; s < PK_NSPLIT; ++s) { ms[s] = __ldcg(a.part_ml + (pbase + s) * 2); M = fmaxf(M, ms[s]); }
                    float Ls = 0.f, O = 0.f;
#pragma unroll
                    for (int s = 0; s < PK_NSPLIT; ++s) {
                        const float c = (ms[s] == -INFINITY) ? 0.f : __expf(ms[s] - M);
                        Ls += __ldcg(a.part_ml + (pbase + s) * 2 + 1) * c;
                        O += __ldcg(a.part_o + (pbase + s) * D + i) * c;
                    }
                    // xs layout: 8-element chunk ci -> xs[ci] (elements 0-3) and xs[K8 + ci] (elements 4-7)
                    const int ci = idx >> 3, e = idx & 7;
                    xs_f[(size_t)((e < 4) ? ci : K8 + ci) * 4 + (e & 3)] = O / Ls;
                }
            } else {
                const float* src = (kind == 3) ? a.act : a.x;
                float ssq = 0.f;
                for (int i = tid; i < K8; i += PK_THREADS) {
                    float4 lo = __ldcg(reinterpret_cast<const float4*>(src) + 2 * i), hi = __ldcg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
                    if (nw) {
                        ssq += lo.x * lo.x + lo.y * lo.y + lo.z * lo.z + lo.w * lo.w + hi.x * hi.x + hi.y * hi.y + hi.z * hi.z + hi.w * hi.w;
                        const float4 w0 = reinterpret_cast<const float4*>(nw)[2 * i], w1 = reinterpret_cast<const float4*>(nw)[2 * i + 1];
                        lo.x *= w0.x; lo.y *= w0.y; lo.z *= w0.z; lo.w *= w0.w;
                        hi.x *= w1.x; hi.y *= w1.y; hi.z *= w1.z; hi.w *= w1.w;
                    }
                    xs[i] = lo;
                    xs[K8 + i] = hi;
                }
                PK_PROF(8);
                if (nw) {
                    const float tot = block_sum(ssq, red);
                    if (tid == 0) rstd_s = rsqrtf(tot / (float)g.K + a.eps);
                }
                PK_PROF(9);
            }
            for (int i = tid; i < sl.rpc; i += PK_THREADS) acc_s[i] = 0.f;
            __syncthreads();
            PK_PROF(2);

            // ---- stream this warp's slice of the weight matrix out of its ring ----
            {
                float accum = 0.f, accum2 = 0.f;
                int cur_row = -1;
                uint32_t row = sl.c_begin / sl.cpr, rem = sl.c_begin - row * sl.cpr;
#pragma unroll 1
                for (uint32_t sg = 0; sg < sl.nseg; ++sg) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(PK_DEPTH - 1) : "memory");
                    __syncwarp();
                    const unsigned char* sp = ring + (c_seq % PK_DEPTH) * PK_SEG_BYTES;
                    const uint32_t cb = sl.c_begin + sg * PK_SEG_CHUNKS;
                    uint4 w[PK_SEG_CHUNKS];
#pragma unroll
                    for (int c = 0; c < PK_SEG_CHUNKS; ++c)
                        if (cb + c < sl.c_end) w[c] = *reinterpret_cast<const uint4*>(sp + c * 512);
#pragma unroll
                    for (int c = 0; c < PK_SEG_CHUNKS; ++c) {
                        if (cb + c < sl.c_end) {
                            const int idx = (int)(rem << 5) + lane;
                            if ((int)row != cur_row) {
                                if (cur_row >= 0) {
                                    const float v = warp_sum(accum + accum2);
                                    if (lane == 0) atomicAdd(&acc_s[cur_row], v);
                                    accum = 0.f; accum2 = 0.f;
                                }
                                cur_row = (int)row;
                            }
                            const float4 xl = xs[idx], xh = xs[K8 + idx];
                            accum = fmaf(bf16lo(w[c].x), xl.x, accum); accum2 = fmaf(bf16lo(w[c].z), xh.x, accum2);
                            accum = fmaf(bf16hi(w[c].x), xl.y, accum); accum2 = fmaf(bf16hi(w[c].z), xh.y, accum2);
                            accum = fmaf(bf16lo(w[c].y), xl.z, accum); accum2 = fmaf(bf16lo(w[c].w), xh.z, accum2);
                            accum = fmaf(bf16hi(w[c].y), xl.w, accum); accum2 = fmaf(bf16hi(w[c].w), xh.w, accum2);
                        }
                        if (++rem == sl.cpr) { rem = 0; ++row; }
                    }
                    __syncwarp();
                    issue_next();
                    ++c_seq;
                }
                if (cur_row >= 0) {
                    const float v = warp_sum(accum + accum2);
                    if (lane == 0) atomicAdd(&acc_s[cur_row], v);
                }
            }
            __syncthreads();
            PK_PROF(3);

            // ---- fused epilogue ----
            const float r = nw ? rstd_s : 1.f;
            if (kind == 2) {          // SiLU(gate) * up on interleaved rows
                for (int u = tid; u < sl.nrows / 2; u += PK_THREADS)
                    a.act[sl.r0 / 2 + u] = silu_f(acc_s[2 * u] * r) * (acc_s[2 * u + 1] * r);
            } else if (kind == 1 || kind == 3) {   // residual add (L2 read: x is also written by CTA 0)
                for (int i = tid; i < sl.nrows; i += PK_THREADS) a.x[sl.r0 + i] = __ldcg(a.x + sl.r0 + i) + acc_s[i];
            } else if (kind == 0) {
                for (int i = tid; i < sl.nrows; i += PK_THREADS) a.qkv[sl.r0 + i] = acc_s[i] * r;
            } else {                  // logits + running argmax (a thread's rows ascend: first maximum wins)
                for (int i = tid; i < sl.nrows; i += PK_THREADS) {
                    const float v = acc_s[i] * r;
                    a.logits[sl.r0 + i] = v;
                    if (v > bestv) { bestv = v; besti = sl.r0 + i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
                }
                if (lane == 0) { wbest_v[warp] = bestv; wbest_i[warp] = besti; }
                __syncthreads();
                if (tid == 0) {
                    float bv = -INFINITY; int bi = 0x7fffffff;
                    for (int w = 0; w < PK_WARPS; ++w)
                        if (wbest_v[w] > bv || (wbest_v[w] == bv && wbest_i[w] < bi)) { bv = wbest_v[w]; bi = wbest_i[w]; }
                    a.part_val[blockIdx.x] = bv;
                    a.part_idx[blockIdx.x] = bi;
                }
            }
            PK_PROF(4);
            bar_target += nct; grid_barrier(a.barrier, bar_target);
            PK_PROF(5);
        }

        // ---- token: every CTA's warp 0 reduces the per-CTA maxima (lowest index wins); CTA 0 publishes ----
        if (blockIdx.x == 0) {
            if (warp == 0) {
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int c = lane; c < (int)nct; c += 32) {
                    const float v = __ldcg(a.part_val + c);
                    const int i = __ldcg(a.part_idx + c);
                    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) tok_s = (uint32_t)bi;
            }
            __syncthreads();
            if (tid == 0 && a.out_tokens) a.out_tokens[st0.step + t] = tok_s;
            if (a.advance) {                               // next step's input embedding (bf16 -> f32 residual stream)
                const bf16* rowp = a.embed + (size_t)tok_s * a.H;
                for (int i = tid; i < a.H; i += PK_THREADS) a.x[i] = __bfloat162float(rowp[i]);
            }
        }
        if (a.advance && t + 1 < a.n_steps) { bar_target += nct; grid_barrier(a.barrier, bar_target); }
        PK_PROF(6);
    }
    if (blockIdx.x == 0 && tid == 0) {
        SeqState* s = a.state;
        s->step = st0.step + a.n_steps;
        if (a.advance) {
            s->kv_len = st0.kv_len + a.n_steps;
            s->pos[0] = st0.pos[0] + a.n_steps; s->pos[1] = st0.pos[1] + a.n_steps; s->pos[2] = st0.pos[2] + a.n_steps;
            s->token = tok_s;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

size_t decode_persistent_smem(const PersistArgs& a, int num_sms) {
    const int max_rows = std::max((a.V + num_sms - 1) / num_sms, (2 * a.I + num_sms - 1) / num_sms + 2) + 2;
    return (size_t)PK_WARPS * PK_DEPTH * PK_SEG_BYTES + (size_t)a.xs_floats * 4 + (size_t)max_rows * 4 + 64;
}

bool decode_persistent_supported(int D, int nrep, int H, int I, int q_dim, int nkv, int num_sms) {
    return D == 128 && (nrep == 1 || nrep == 2 || nrep == 4 || nrep == 8) && H % 256 == 0 && I % 256 == 0 && q_dim % 256 == 0 &&
           nkv * PK_NSPLIT <= num_sms;
}

template <int NREP>
static int launch_t(cudaStream_t st, const PersistArgs& a, int num_sms) {
    const size_t smem = decode_persistent_smem(a, num_sms);
    if (smem > 227 * 1024) return -1000;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(decode_persistent_kernel<NREP>, smem, seen)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms);
    cfg.blockDim = dim3(PK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs must be co-resident (grid barriers)
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, decode_persistent_kernel<NREP>, a);
}

int decode_persistent_launch(cudaStream_t st, const PersistArgs& a, int num_sms) {
    const int nrep = a.nh / a.nkv;
    switch (nrep) {
        case 1: return launch_t<1>(st, a, num_sms);
        case 2: return launch_t<2>(st, a, num_sms);
        case 4: return launch_t<4>(st, a, num_sms);
        case 8: return launch_t<8>(st, a, num_sms);
        default: return -1000;
    }
}

}  // namespace cb
