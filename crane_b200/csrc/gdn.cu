// Gated-Delta-Net layer body between the in_proj and out_proj GEMMs (Qwen3.5 linear-attention layers).
//
// Reference arithmetic (crane-core/src/ops/gdn/): causal depthwise conv1d + SiLU with a carried conv state
// (conv.rs:23-133), split into per-head q/k/v (layer.rs:194-238, Interleaved v-head order), L2-norm of q and k
// (backend.rs:26-56), beta = sigmoid(b), g = -exp(A_log) * softplus(a + dt_bias) (backend.rs:197-211), the gated
// delta rule  S <- S*exp(g); kv = S^T k; d = (v - kv)*beta; S <- S + k (x) d; y = S^T (q/sqrt(K))
// (backend.rs:90-156, kernels/cuda/gdn.cu:29-34), gated RMSNorm y*rsqrt(mean y^2+eps)*w*silu(z) (norm.rs:39-45).
// Everything is f32, as in the reference's recurrence.
//
// B200 mapping of the recurrence (the reference kernel runs 2048 threads with two __syncthreads per timestep):
// one CTA per (value head, 32 value columns); a state column's 128 K-entries are split over 4 adjacent lanes
// (32 registers each), so a timestep needs two 2-step shuffle reductions and NO block barrier; q/k/v/gates of 16
// timesteps are staged in shared memory per barrier pair.  State is read once and written once per call.
#include "gdn.cuh"
#include "prof.h"

namespace cb {

constexpr int GDN_MAX_CK = 8;

__global__ void __launch_bounds__(128)
gdn_conv_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int t = blockIdx.y;
    if (c >= conv_dim) return;
    float acc = 0.f;
    for (int j = 0; j < a.ck; ++j) {
        const int m = t + 1 + j;   // index into [state(ck) | x(S)]
        const float h = (m < a.ck) ? a.conv_state[(size_t)c * a.ck + m] : a.proj[(size_t)(m - a.ck) * a.ldp + c];
        acc = fmaf(a.conv_w[(size_t)c * a.ck + j], h, acc);
    }
    a.conv_out[(size_t)t * conv_dim + c] = silu_f(acc);
}

// The same convolution for the usual 4-tap kernel over many rows: one thread per channel walks GDN_CONV_TT consecutive timesteps with
// the tap window in registers, so every input element is loaded once (the per-element kernel above loads it ck times and, at
// 4 096 rows, spends its time on 200 K tiny CTAs: 205 us per layer measured, profiles/r02_launches_gdn_chunk.csv).
constexpr int GDN_CONV_TT = 32;
__global__ void __launch_bounds__(128)
gdn_conv4_rows_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int t0 = blockIdx.y * GDN_CONV_TT;
    if (c >= conv_dim) return;
    const int t1 = min(t0 + GDN_CONV_TT, a.S);
    const float w0 = a.conv_w[(size_t)c * 4 + 0], w1 = a.conv_w[(size_t)c * 4 + 1], w2 = a.conv_w[(size_t)c * 4 + 2], w3 = a.conv_w[(size_t)c * 4 + 3];
    // history index m into [state(4) | x(S)]: the window of timestep t is m = t + 1 .. t + 4
    auto hist = [&](int m) { return (m < 4) ? a.conv_state[(size_t)c * 4 + m] : a.proj[(size_t)(m - 4) * a.ldp + c]; };
    float h0 = hist(t0 + 1), h1 = hist(t0 + 2), h2 = hist(t0 + 3);
    for (int tb = t0; tb < t1; tb += 8) {                 // eight rows in flight before the first store (see gdn_conv4_qkv_kernel)
        float hx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) hx[u] = (tb + u < t1) ? a.proj[(size_t)(tb + u) * a.ldp + c] : 0.f;      // m = t + 4 -> x[t]
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (tb + u >= t1) break;
            const float h3 = hx[u];
            float acc = w0 * h0;
            acc = fmaf(w1, h1, acc);
            acc = fmaf(w2, h2, acc);
            acc = fmaf(w3, h3, acc);
            a.conv_out[(size_t)(tb + u) * conv_dim + c] = silu_f(acc);
            h0 = h1; h1 = h2; h2 = h3;
        }
    }
}

// Conv + SiLU + the per-head L2 norms + the gates in one pass, for 128-wide key heads: a thread owns 4 adjacent channels (16-byte
// loads), a warp therefore owns exactly one q head, one k head or 128 value channels, and walks GDN_CONV_TT timesteps with the tap
// window in registers.  q / k leave as l2norm(q) / sqrt(dk) and l2norm(k) (backend.rs:26-56) straight into qn / kn -- the q / k part
// of conv_out is never materialised -- v goes to conv_out; the CTAs of the first channel block also write beta and g
// (backend.rs:197-211).  Replaces gdn_conv4_rows_kernel + gdn_prep_kernel (75 + 30 us per layer at 4 096 rows).
// SiLU with the hardware exponential and reciprocal (2 ulp each): the kernel is bound by its instruction count -- 4 SiLUs per thread
// and timestep were 173 instructions per warp-timestep with expf and an IEEE division (ncu: 34 M instructions, SFU pipe 30 % busy)
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }

__global__ void __launch_bounds__(128)
gdn_conv4_qkv_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = blockIdx.y * GDN_CONV_TT;
    const int t1 = min(t0 + GDN_CONV_TT, a.S);
    const int cb = (blockIdx.x * 4 + warp) * 128;          // this warp's 128 channels
    if (cb < conv_dim) {
        const int c = cb + lane * 4;
        const int grp = cb / 128;                          // < nk: q head; < 2 nk: k head; else value channels
        const float4* wp = reinterpret_cast<const float4*>(a.conv_w + (size_t)c * 4);
        const float4 wa = wp[0], wb = wp[1], wc = wp[2], wd = wp[3];      // taps of channels c, c+1, c+2, c+3
        auto hist = [&](int m) {                           // history index m into [state(4) | x(S)], 4 channels at once
            if (m < 4) return make_float4(a.conv_state[(size_t)c * 4 + m], a.conv_state[(size_t)(c + 1) * 4 + m],
                                          a.conv_state[(size_t)(c + 2) * 4 + m], a.conv_state[(size_t)(c + 3) * 4 + m]);
            return *reinterpret_cast<const float4*>(a.proj + (size_t)(m - 4) * a.ldp + c);
        };
        float4 h0 = hist(t0 + 1), h1 = hist(t0 + 2), h2 = hist(t0 + 3);
        const float qscale = 1.0f / sqrtf((float)a.dk);
        // rows are fetched eight at a time BEFORE any of them is consumed: the stores below may alias the loads as far as the compiler
        // knows, so a plain loop serialises one DRAM round trip per timestep (74 us per layer measured that way)
        for (int tb = t0; tb < t1; tb += 8) {
            float4 hx[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                hx[u] = (tb + u < t1) ? *reinterpret_cast<const float4*>(a.proj + (size_t)(tb + u) * a.ldp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = tb + u;
                if (t >= t1) break;
                const float4 h3 = hx[u];
                float4 y;
                y.x = silu_fast(fmaf(wa.w, h3.x, fmaf(wa.z, h2.x, fmaf(wa.y, h1.x, wa.x * h0.x))));
                y.y = silu_fast(fmaf(wb.w, h3.y, fmaf(wb.z, h2.y, fmaf(wb.y, h1.y, wb.x * h0.y))));
                y.z = silu_fast(fmaf(wc.w, h3.z, fmaf(wc.z, h2.z, fmaf(wc.y, h1.z, wc.x * h0.z))));
                y.w = silu_fast(fmaf(wd.w, h3.w, fmaf(wd.z, h2.w, fmaf(wd.y, h1.w, wd.x * h0.w))));
                if (grp < 2 * a.nk) {
                    const float ssq = warp_sum((y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w));
                    float sc = rsqrtf(ssq + 1e-6f);
                    const bool is_k = grp >= a.nk;
                    if (!is_k) sc *= qscale;
                    float* dst = (is_k ? a.kn : a.qn) + ((size_t)t * a.nk + (is_k ? grp - a.nk : grp)) * 128 + lane * 4;
                    *reinterpret_cast<float4*>(dst) = make_float4(y.x * sc, y.y * sc, y.z * sc, y.w * sc);
                } else {
                    *reinterpret_cast<float4*>(a.conv_out + (size_t)t * conv_dim + c) = y;
                }
                h0 = h1; h1 = h2; h2 = h3;
            }
        }
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < (t1 - t0) * a.nv; i += 128) {
            const int t = t0 + i / a.nv, h = i % a.nv;
            const float* pr = a.proj + (size_t)t * a.ldp + conv_dim + a.nv * a.dv;
            const float b = pr[h], av = pr[a.nv + h];
            const float g = a.neg_exp_a[h] * logf(1.0f + expf(av + a.dt_bias[h]));
            a.gb[((size_t)t * a.nv + h) * 2 + 0] = expf(g);
            a.gb[((size_t)t * a.nv + h) * 2 + 1] = 1.0f / (1.0f + expf(-b));
            if (a.glog) a.glog[(size_t)t * a.nv + h] = g;
        }
    }
}

__global__ void __launch_bounds__(128)
gdn_conv_state_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= conv_dim) return;
    float ns[GDN_MAX_CK];
    for (int j = 0; j < a.ck; ++j) {
        const int m = a.S + j;
        ns[j] = (m < a.ck) ? a.conv_state[(size_t)c * a.ck + m] : a.proj[(size_t)(m - a.ck) * a.ldp + c];
    }
    for (int j = 0; j < a.ck; ++j) a.conv_state[(size_t)c * a.ck + j] = ns[j];
}

// grid (S), 256 threads: warps normalise the 2*nk q/k head vectors of timestep t; threads < nv compute the gates.
__global__ void __launch_bounds__(256)
gdn_prep_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int t = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const int key_dim = a.nk * a.dk;
    const float* row = a.conv_out + (size_t)t * conv_dim;
    for (int vec = warp; vec < 2 * a.nk; vec += 8) {
        const bool is_k = vec >= a.nk;
        const int head = is_k ? vec - a.nk : vec;
        const float* src = row + (is_k ? key_dim : 0) + head * a.dk;
        float ssq = 0.f;
        for (int i = lane; i < a.dk; i += 32) ssq += src[i] * src[i];
        ssq = warp_sum(ssq);
        float sc = 1.0f / sqrtf(ssq + 1e-6f);
        if (!is_k) sc *= 1.0f / sqrtf((float)a.dk);
        float* dst = (is_k ? a.kn : a.qn) + ((size_t)t * a.nk + head) * a.dk;
        for (int i = lane; i < a.dk; i += 32) dst[i] = src[i] * sc;
    }
    if (threadIdx.x < a.nv) {
        const int h = threadIdx.x;
        const float* pr = a.proj + (size_t)t * a.ldp + conv_dim + a.nv * a.dv;
        const float b = pr[h], av = pr[a.nv + h];
        const float beta = 1.0f / (1.0f + expf(-b));
        const float g = a.neg_exp_a[h] * logf(1.0f + expf(av + a.dt_bias[h]));
        a.gb[((size_t)t * a.nv + h) * 2 + 0] = expf(g);
        a.gb[((size_t)t * a.nv + h) * 2 + 1] = beta;
        if (a.glog) a.glog[(size_t)t * a.nv + h] = g;      // the chunkwise path sums g inside a chunk instead of multiplying decays
    }
}

// grid (nv * dv/32), 128 threads.  dk == 128.
// The S steps are strictly sequential, so a step's latency is everything: the two 32-term dots run as 4 independent FMA chains
// (one per float4 component), and the per-step operands (q, k, v, decay, beta) stream through a double-buffered cp.async ring of
// GDN_TC-step chunks so no step ever waits on global memory.
constexpr int GDN_TC = 32;
constexpr int GDN_KP = 128 + 16;                   // padded row: 4 floats of skew per 32 (conflict-free float4 reads by the 4 lanes of a column)
__device__ __forceinline__ int gdn_pad(int k) { return k + (k >> 5) * 4; }
constexpr size_t GDN_RECUR_SMEM = (size_t)2 * GDN_TC * (2 * GDN_KP + 32 + 2) * sizeof(float);

__global__ void __launch_bounds__(128)
gdn_recur_kernel(GdnArgs a) {
    pdl_wait();
    constexpr int DK = 128, KP = GDN_KP;
    extern __shared__ __align__(16) float gsm[];
    float* q_s = gsm;                               // [2][TC][KP]
    float* k_s = q_s + 2 * GDN_TC * KP;             // [2][TC][KP]
    float* v_s = k_s + 2 * GDN_TC * KP;             // [2][TC][32]
    float* gb_s = v_s + 2 * GDN_TC * 32;            // [2][TC][2]
    const int tiles = a.dv / 32;
    const int h = blockIdx.x / tiles, vt = blockIdx.x % tiles;
    const int kh = h / (a.nv / a.nk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c4 = lane & 3;
    const int col_local = warp * 8 + (lane >> 2);
    const int col = vt * 32 + col_local;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    float s[32];
    float* sp = a.rec_state + ((size_t)h * DK + 32 * c4) * a.dv + col;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = sp[(size_t)i * a.dv];

    auto prefetch = [&](int t0, int buf) {
        const int nt = min(GDN_TC, a.S - t0);
        if (nt > 0) {
            for (int i = tid; i < nt * 32; i += 128) {       // 32 16-byte pieces of q and of k per step
                const int tt = i >> 5, c = i & 31;
                const size_t src = ((size_t)(t0 + tt) * a.nk + kh) * DK + c * 4;
                const int dst = (buf * GDN_TC + tt) * KP + gdn_pad(c * 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(q_s + dst)), "l"(a.qn + src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(k_s + dst)), "l"(a.kn + src) : "memory");
            }
            for (int i = tid; i < nt * 8; i += 128) {
                const int tt = i >> 3, c = i & 7;
                const float* src = a.conv_out + (size_t)(t0 + tt) * conv_dim + 2 * a.nk * a.dk + h * a.dv + vt * 32 + c * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(v_s + (buf * GDN_TC + tt) * 32 + c * 4)), "l"(src) : "memory");
            }
            if (tid < nt) {
                const float* src = a.gb + ((size_t)(t0 + tid) * a.nv + h) * 2;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(gb_s + (buf * GDN_TC + tid) * 2)), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < a.S; t0 += GDN_TC, buf ^= 1) {
        const int nt = min(GDN_TC, a.S - t0);
        prefetch(t0 + GDN_TC, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        for (int tt = 0; tt < nt; ++tt) {
            const float decay = gb_s[(buf * GDN_TC + tt) * 2], beta = gb_s[(buf * GDN_TC + tt) * 2 + 1];
            const float4* kp = reinterpret_cast<const float4*>(k_s + (buf * GDN_TC + tt) * KP + gdn_pad(32 * c4));
            const float4* qp = reinterpret_cast<const float4*>(q_s + (buf * GDN_TC + tt) * KP + gdn_pad(32 * c4));
            float4 kr[8];
            float kx = 0.f, ky = 0.f, kz = 0.f, kw = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                kr[i] = kp[i];
                s[4 * i + 0] *= decay; s[4 * i + 1] *= decay; s[4 * i + 2] *= decay; s[4 * i + 3] *= decay;
                kx = fmaf(s[4 * i + 0], kr[i].x, kx); ky = fmaf(s[4 * i + 1], kr[i].y, ky);
                kz = fmaf(s[4 * i + 2], kr[i].z, kz); kw = fmaf(s[4 * i + 3], kr[i].w, kw);
            }
            float kv = (kx + ky) + (kz + kw);
            kv += __shfl_xor_sync(0xffffffffu, kv, 1);
            kv += __shfl_xor_sync(0xffffffffu, kv, 2);
            const float delta = (v_s[(buf * GDN_TC + tt) * 32 + col_local] - kv) * beta;
            float yx = 0.f, yy = 0.f, yz = 0.f, yw = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 q4 = qp[i];
                s[4 * i + 0] = fmaf(kr[i].x, delta, s[4 * i + 0]); yx = fmaf(s[4 * i + 0], q4.x, yx);
                s[4 * i + 1] = fmaf(kr[i].y, delta, s[4 * i + 1]); yy = fmaf(s[4 * i + 1], q4.y, yy);
                s[4 * i + 2] = fmaf(kr[i].z, delta, s[4 * i + 2]); yz = fmaf(s[4 * i + 2], q4.z, yz);
                s[4 * i + 3] = fmaf(kr[i].w, delta, s[4 * i + 3]); yw = fmaf(s[4 * i + 3], q4.w, yw);
            }
            float y = (yx + yy) + (yz + yw);
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (c4 == 0) a.y[((size_t)(t0 + tt) * a.nv + h) * a.dv + col] = y;
        }
        __syncthreads();                               // this buffer is refilled by the prefetch of the next iteration
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) sp[(size_t)i * a.dv] = s[i];
    pdl_launch_dependents();
}

// The same recurrence for the other key widths the reference accepts (K <= 256, kernels/cuda/gdn.cu:45-153): DK = 64 or 256.
// Identical mapping -- 4 lanes per state column, DK / 4 state entries per lane, operands through the double-buffered ring --
// with the row padding computed per 16-byte piece (a lane's slice crosses the 32-float padding groups when DK = 256).
template <int DK>
__global__ void __launch_bounds__(128)
gdn_recur_any_kernel(GdnArgs a) {
    pdl_wait();
    constexpr int KP = DK + DK / 8, KPL = DK / 4;   // padded row, state entries per lane
    extern __shared__ __align__(16) float gsm[];
    float* q_s = gsm;                               // [2][TC][KP]
    float* k_s = q_s + 2 * GDN_TC * KP;             // [2][TC][KP]
    float* v_s = k_s + 2 * GDN_TC * KP;             // [2][TC][32]
    float* gb_s = v_s + 2 * GDN_TC * 32;            // [2][TC][2]
    const int tiles = a.dv / 32;
    const int h = blockIdx.x / tiles, vt = blockIdx.x % tiles;
    const int kh = h / (a.nv / a.nk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c4 = lane & 3;
    const int col_local = warp * 8 + (lane >> 2);
    const int col = vt * 32 + col_local;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    float s[KPL];
    float* sp = a.rec_state + ((size_t)h * DK + KPL * c4) * a.dv + col;
#pragma unroll
    for (int i = 0; i < KPL; ++i) s[i] = sp[(size_t)i * a.dv];

    auto prefetch = [&](int t0, int buf) {
        const int nt = min(GDN_TC, a.S - t0);
        if (nt > 0) {
            for (int i = tid; i < nt * (DK / 4); i += 128) {   // DK / 4 16-byte pieces of q and of k per step
                const int tt = i / (DK / 4), c = i % (DK / 4);
                const size_t src = ((size_t)(t0 + tt) * a.nk + kh) * DK + c * 4;
                const int dst = (buf * GDN_TC + tt) * KP + gdn_pad(c * 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(q_s + dst)), "l"(a.qn + src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(k_s + dst)), "l"(a.kn + src) : "memory");
            }
            for (int i = tid; i < nt * 8; i += 128) {
                const int tt = i >> 3, c = i & 7;
                const float* src = a.conv_out + (size_t)(t0 + tt) * conv_dim + 2 * a.nk * a.dk + h * a.dv + vt * 32 + c * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(v_s + (buf * GDN_TC + tt) * 32 + c * 4)), "l"(src) : "memory");
            }
            if (tid < nt) {
                const float* src = a.gb + ((size_t)(t0 + tid) * a.nv + h) * 2;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(gb_s + (buf * GDN_TC + tid) * 2)), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < a.S; t0 += GDN_TC, buf ^= 1) {
        const int nt = min(GDN_TC, a.S - t0);
        prefetch(t0 + GDN_TC, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        for (int tt = 0; tt < nt; ++tt) {
            const float decay = gb_s[(buf * GDN_TC + tt) * 2], beta = gb_s[(buf * GDN_TC + tt) * 2 + 1];
            const float* krow = k_s + (buf * GDN_TC + tt) * KP;
            const float* qrow = q_s + (buf * GDN_TC + tt) * KP;
            float kx = 0.f, ky = 0.f, kz = 0.f, kw = 0.f;
#pragma unroll
            for (int i = 0; i < KPL / 4; ++i) {
                const float4 k4 = *reinterpret_cast<const float4*>(krow + gdn_pad(KPL * c4 + 4 * i));
                s[4 * i + 0] *= decay; s[4 * i + 1] *= decay; s[4 * i + 2] *= decay; s[4 * i + 3] *= decay;
                kx = fmaf(s[4 * i + 0], k4.x, kx); ky = fmaf(s[4 * i + 1], k4.y, ky);
                kz = fmaf(s[4 * i + 2], k4.z, kz); kw = fmaf(s[4 * i + 3], k4.w, kw);
            }
            float kv = (kx + ky) + (kz + kw);
            kv += __shfl_xor_sync(0xffffffffu, kv, 1);
            kv += __shfl_xor_sync(0xffffffffu, kv, 2);
            const float delta = (v_s[(buf * GDN_TC + tt) * 32 + col_local] - kv) * beta;
            float yx = 0.f, yy = 0.f, yz = 0.f, yw = 0.f;
#pragma unroll
            for (int i = 0; i < KPL / 4; ++i) {
                const float4 k4 = *reinterpret_cast<const float4*>(krow + gdn_pad(KPL * c4 + 4 * i));
                const float4 q4 = *reinterpret_cast<const float4*>(qrow + gdn_pad(KPL * c4 + 4 * i));
                s[4 * i + 0] = fmaf(k4.x, delta, s[4 * i + 0]); yx = fmaf(s[4 * i + 0], q4.x, yx);
                s[4 * i + 1] = fmaf(k4.y, delta, s[4 * i + 1]); yy = fmaf(s[4 * i + 1], q4.y, yy);
                s[4 * i + 2] = fmaf(k4.z, delta, s[4 * i + 2]); yz = fmaf(s[4 * i + 2], q4.z, yz);
                s[4 * i + 3] = fmaf(k4.w, delta, s[4 * i + 3]); yw = fmaf(s[4 * i + 3], q4.w, yw);
            }
            float y = (yx + yy) + (yz + yw);
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (c4 == 0) a.y[((size_t)(t0 + tt) * a.nv + h) * a.dv + col] = y;
        }
        __syncthreads();                               // this buffer is refilled by the prefetch of the next iteration
    }
#pragma unroll
    for (int i = 0; i < KPL; ++i) sp[(size_t)i * a.dv] = s[i];
    pdl_launch_dependents();
}
template <int DK> constexpr size_t gdn_recur_any_smem() { return (size_t)2 * GDN_TC * (2 * (DK + DK / 8) + 32 + 2) * sizeof(float); }

// The gated norm for 128-wide value heads: a lane owns 4 adjacent elements (one 16-byte load of y, one of z, one 8-byte store per
// output plane) instead of four strided scalars -- 100 MB of traffic per layer at 4 096 rows that the scalar version moved in 38 us.
__global__ void __launch_bounds__(128)
gdn_gated_norm128_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gw >= a.S * a.nv) return;
    const int t = gw / a.nv, h = gw % a.nv;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const size_t idx = ((size_t)t * a.nv + h) * 128 + lane * 4;
    const float4 y = *reinterpret_cast<const float4*>(a.y + idx);
    const float4 z = *reinterpret_cast<const float4*>(a.proj + (size_t)t * a.ldp + conv_dim + h * 128 + lane * 4);
    const float4 w = *reinterpret_cast<const float4*>(a.norm_w + lane * 4);
    const float ssq = warp_sum((y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w));
    const float rstd = rsqrtf(ssq / 128.0f + a.eps);
    float4 o;
    o.x = y.x * rstd * w.x * silu_f(z.x); o.y = y.y * rstd * w.y * silu_f(z.y);
    o.z = y.z * rstd * w.z * silu_f(z.z); o.w = y.w * rstd * w.w * silu_f(z.w);
    if (a.out_bf16) {
        uint2 hi, lo;
        split_bf16x2(o.x, o.y, hi.x, lo.x);
        split_bf16x2(o.z, o.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(a.out_bf16 + idx) = hi;
        if (a.out_lo_off) *reinterpret_cast<uint2*>(a.out_bf16 + a.out_lo_off + idx) = lo;
    }
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + idx) = o;
}

// One warp per (t, value head): y * rsqrt(mean(y^2) + eps) * w * silu(z)
__global__ void __launch_bounds__(128)
gdn_gated_norm_kernel(GdnArgs a) {
    pdl_wait();
    pdl_launch_dependents();
    const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gw >= a.S * a.nv) return;
    const int t = gw / a.nv, h = gw % a.nv;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const float* yr = a.y + ((size_t)t * a.nv + h) * a.dv;
    const float* zr = a.proj + (size_t)t * a.ldp + conv_dim + h * a.dv;
    float ssq = 0.f;
    for (int i = lane; i < a.dv; i += 32) ssq += yr[i] * yr[i];
    ssq = warp_sum(ssq);
    const float rstd = rsqrtf(ssq / (float)a.dv + a.eps);
    for (int i = lane; i < a.dv; i += 32) {
        const float o = yr[i] * rstd * a.norm_w[i] * silu_f(zr[i]);
        const size_t idx = ((size_t)t * a.nv + h) * a.dv + i;
        if (a.out_bf16) {
            const bf16 hi = __float2bfloat16_rn(o);
            a.out_bf16[idx] = hi;
            if (a.out_lo_off) a.out_bf16[a.out_lo_off + idx] = __float2bfloat16_rn(o - __bfloat162float(hi));
        }
        if (a.out_f32) a.out_f32[idx] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Decode (S == 1) fast path: the five stages above for ONE token in ONE launch.  Everything a value head needs is head-local when
// nk == nv (its own q / k / v channels of the conv, its 128 x 128 state, its gates), so one 512-thread CTA per head does
// conv + state shift -> l2norm / gates -> one recurrence step -> gated RMSNorm.  Same arithmetic, four launches fewer per layer.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
gdn_decode_kernel(GdnArgs a) {
    constexpr int DK = 128, DV = 128;
    __shared__ __align__(16) float q_s[DK], k_s[DK], v_s[DV], y_s[DV];
    __shared__ float red[32];
    __shared__ float gate_s[2];
    pdl_wait();
    pdl_launch_dependents();
    const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int key_dim = a.nk * DK, conv_dim = 2 * key_dim + a.nv * DV;
    // ---- depthwise causal conv (window = last ck-1 cached inputs + this token) + SiLU, and the state shift ----
    if (tid < 3 * DK) {
        const int which = tid / DK, i = tid % DK;                 // 0: q, 1: k, 2: v
        const int c = which == 0 ? h * DK + i : which == 1 ? key_dim + h * DK + i : 2 * key_dim + h * DV + i;
        float* st = a.conv_state + (size_t)c * a.ck;
        const float xin = a.proj[c];
        float acc = 0.f;
        for (int j = 0; j < a.ck; ++j) {
            const float hv = (j + 1 < a.ck) ? st[j + 1] : xin;
            acc = fmaf(a.conv_w[(size_t)c * a.ck + j], hv, acc);
        }
        for (int j = 0; j + 1 < a.ck; ++j) st[j] = st[j + 1];      // one thread per channel: in-order shift is safe
        st[a.ck - 1] = xin;
        (which == 0 ? q_s : which == 1 ? k_s : v_s)[i] = silu_f(acc);
    }
    if (tid == 3 * DK) {
        const float* pr = a.proj + conv_dim + a.nv * DV;
        const float b = pr[h], av = pr[a.nv + h];
        gate_s[0] = expf(a.neg_exp_a[h] * logf(1.0f + expf(av + a.dt_bias[h])));
        gate_s[1] = 1.0f / (1.0f + expf(-b));
    }
    __syncthreads();
    // ---- l2norm(q) / sqrt(dk), l2norm(k) ----
    if (warp < 2) {
        float* vsrc = warp == 0 ? q_s : k_s;
        float e[4], ssq = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { e[j] = vsrc[lane + 32 * j]; ssq += e[j] * e[j]; }
        ssq = warp_sum(ssq);
        float sc = 1.0f / sqrtf(ssq + 1e-6f);
        if (warp == 0) sc *= 1.0f / sqrtf((float)DK);
#pragma unroll
        for (int j = 0; j < 4; ++j) vsrc[lane + 32 * j] = e[j] * sc;
    }
    __syncthreads();
    // ---- one step of the gated delta rule: 4 lanes per state column ----
    const int col = tid >> 2, c4 = tid & 3;
    float* sp = a.rec_state + ((size_t)h * DK + 32 * c4) * DV + col;
    const float decay = gate_s[0], beta = gate_s[1];
    float s[32];
    float kx = 0.f, ky = 0.f, kz = 0.f, kw = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = sp[(size_t)i * DV] * decay;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 k4 = *reinterpret_cast<const float4*>(k_s + 32 * c4 + 4 * i);
        kx = fmaf(s[4 * i + 0], k4.x, kx); ky = fmaf(s[4 * i + 1], k4.y, ky);
        kz = fmaf(s[4 * i + 2], k4.z, kz); kw = fmaf(s[4 * i + 3], k4.w, kw);
    }
    float kv = (kx + ky) + (kz + kw);
    kv += __shfl_xor_sync(0xffffffffu, kv, 1);
    kv += __shfl_xor_sync(0xffffffffu, kv, 2);
    const float delta = (v_s[col] - kv) * beta;
    float yx = 0.f, yy = 0.f, yz = 0.f, yw = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 k4 = *reinterpret_cast<const float4*>(k_s + 32 * c4 + 4 * i);
        const float4 q4 = *reinterpret_cast<const float4*>(q_s + 32 * c4 + 4 * i);
        s[4 * i + 0] = fmaf(k4.x, delta, s[4 * i + 0]); yx = fmaf(s[4 * i + 0], q4.x, yx);
        s[4 * i + 1] = fmaf(k4.y, delta, s[4 * i + 1]); yy = fmaf(s[4 * i + 1], q4.y, yy);
        s[4 * i + 2] = fmaf(k4.z, delta, s[4 * i + 2]); yz = fmaf(s[4 * i + 2], q4.z, yz);
        s[4 * i + 3] = fmaf(k4.w, delta, s[4 * i + 3]); yw = fmaf(s[4 * i + 3], q4.w, yw);
    }
    float y = (yx + yy) + (yz + yw);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
#pragma unroll
    for (int i = 0; i < 32; ++i) sp[(size_t)i * DV] = s[i];
    if (c4 == 0) y_s[col] = y;
    __syncthreads();
    // ---- rmsnorm(y) * w * silu(z) ----
    const float yv = tid < DV ? y_s[tid] : 0.f;
    const float tot = block_sum(yv * yv, red);
    if (tid < DV) {
        const float rstd = rsqrtf(tot / (float)DV + a.eps);
        const float z = a.proj[conv_dim + h * DV + tid];
        a.out_f32[(size_t)h * DV + tid] = yv * rstd * a.norm_w[tid] * silu_f(z);
    }
}

template <int DK>
static int recur_any_launch(cudaStream_t st, const GdnArgs& a) {
    static SmemOptIn seen;
    int r = ensure_dyn_smem(gdn_recur_any_kernel<DK>, gdn_recur_any_smem<DK>(), seen);
    if (!r) r = launch_k(gdn_recur_any_kernel<DK>, dim3(a.nv * (a.dv / 32)), dim3(128), gdn_recur_any_smem<DK>(), st, false, a);
    return r;
}

bool gdn_shape_supported(const GdnArgs& a) {
    return (a.dk == 64 || a.dk == 128 || a.dk == 256) && a.dv > 0 && (a.dv % 32) == 0 && a.ck >= 1 && a.ck <= GDN_MAX_CK && a.nv <= 256 &&
           a.nk > 0 && (a.nv % a.nk) == 0;
}

// kernels gdn_forward_launch enqueues for these arguments (the engine's launch counter)
int gdn_forward_launch_count(const GdnArgs& a) {
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    if (a.S == 1 && a.nk == a.nv && a.dk == 128 && a.dv == 128 && a.ck >= 2 && a.out_f32 != nullptr && a.out_bf16 == nullptr) return 1;
    const bool fused_qkv = a.ck == 4 && a.dk == 128 && a.S >= GDN_CONV_TT && conv_dim % 128 == 0 && a.ldp % 4 == 0;
    return 2 + (fused_qkv ? 0 : 1) + (gdn_chunk_supported(a) ? 3 : 1) + 1;
}

int gdn_forward_launch(cudaStream_t st, const GdnArgs& a) {
    if (!gdn_shape_supported(a)) return -1000;
    const int conv_dim = 2 * a.nk * a.dk + a.nv * a.dv;
    const bool pdl = prefill_pdl();
    if (a.S == 1 && a.nk == a.nv && a.dk == 128 && a.dv == 128 && a.ck >= 2 && a.out_f32 != nullptr && a.out_bf16 == nullptr)
        return launch_k(gdn_decode_kernel, dim3(a.nv), dim3(512), 0, st, pdl, a);
    span_mark(SP_GDN_CONV);
    const bool fused_qkv = a.ck == 4 && a.dk == 128 && a.S >= GDN_CONV_TT && conv_dim % 128 == 0 && a.ldp % 4 == 0;
    int r = fused_qkv
                ? launch_k(gdn_conv4_qkv_kernel, dim3((conv_dim + 511) / 512, (a.S + GDN_CONV_TT - 1) / GDN_CONV_TT), dim3(128), 0, st, pdl, a)
            : (a.ck == 4 && a.S >= GDN_CONV_TT)
                ? launch_k(gdn_conv4_rows_kernel, dim3((conv_dim + 127) / 128, (a.S + GDN_CONV_TT - 1) / GDN_CONV_TT), dim3(128), 0, st, pdl, a)
                : launch_k(gdn_conv_kernel, dim3((conv_dim + 127) / 128, a.S), dim3(128), 0, st, pdl, a);
    if (!r) r = launch_k(gdn_conv_state_kernel, dim3((conv_dim + 127) / 128), dim3(128), 0, st, pdl, a);
    span_mark(SP_GDN_QKV);
    if (!r && !fused_qkv) r = launch_k(gdn_prep_kernel, dim3(a.S), dim3(256), 0, st, pdl, a);
    span_mark(SP_GDN_RECUR);
    if (!r && gdn_chunk_supported(a)) {
        // prefill: 64 tokens per serial step on the tensor cores (gdn_chunk.cu)
        r = gdn_chunk_recur_launch(st, a);
    } else if (!r && a.dk == 128) {
        // NOT a programmatic dependent: launched early, its 64 long-running CTAs land wherever the previous kernel leaves room and
        // pile up several to an SM; launched after it, they spread one per SM (measured: 4.4 vs 2.4 ms per layer at 4096 tokens)
        static SmemOptIn seen;
        r = ensure_dyn_smem(gdn_recur_kernel, GDN_RECUR_SMEM, seen);
        if (!r) r = launch_k(gdn_recur_kernel, dim3(a.nv * (a.dv / 32)), dim3(128), GDN_RECUR_SMEM, st, false, a);
    } else if (!r) {
        r = a.dk == 64 ? recur_any_launch<64>(st, a) : recur_any_launch<256>(st, a);
    }
    span_mark(SP_GDN_FINISH);
    const bool norm128 = a.dv == 128 && a.ldp % 4 == 0 && (2 * a.nk * a.dk) % 4 == 0 && (a.out_lo_off % 4) == 0;
    if (!r) r = norm128 ? launch_k(gdn_gated_norm128_kernel, dim3((a.S * a.nv + 3) / 4), dim3(128), 0, st, pdl, a)
                        : launch_k(gdn_gated_norm_kernel, dim3((a.S * a.nv + 3) / 4), dim3(128), 0, st, pdl, a);
    return r;
}

}  // namespace cb
