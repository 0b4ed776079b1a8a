// Persistent decode kernel: ALL phases of n greedy decode steps of one sequence in ONE cooperative launch, synchronised by
// flag-carrying activations instead of grid barriers.
//
// Why.  A batch-1 decode step of Qwen3-VL-2B is 141 dependent bandwidth-bound phases (113 weight matrices + 28 attentions) of
// 1.3-96 us of HBM time each.  As separate kernels every phase boundary costs ~2.5 us during which the HBM pipe is idle
// (round 1: 50 % of the HBM roofline); a grid barrier per phase costs about the same.  Here:
//   * one CTA per SM stays resident; each of its 16 warps owns a fixed (column group, row lane) of every weight matrix and
//     streams it through a private cp.async ring (DEPTH x 2 KB per warp, 192 KB per SM).  The prefetch cursor is independent of
//     the consumption cursor and runs ahead across phase, layer and token boundaries: while a warp waits for its input
//     activations the next matrices keep arriving, so a dependency stall shorter than the ring (~4 us of HBM time) costs no
//     bandwidth;
//   * activations cross SMs as 8-byte (value, tag) pairs written with one store and polled by their consumers (the NCCL
//     "LL" protocol applied to a GEMV chain): no separate flag, no fence, no barrier -- one L2 round trip per dependency.  A
//     warp polls only the 256*ncc columns it owns and keeps them in registers for the whole phase (no shared-memory staging,
//     no block barrier at phase start); RMSNorm is folded in (weights multiplied at load, 1/rms applied in the epilogue);
//   * attention runs as nkv x nsplit CTA items (K/V straight from the pages into registers, loads issued before the query is
//     polled), partials are published as pairs and merged by the CTA that owns the head (one split per lane, shuffles).
// Arithmetic is identical to the multi-kernel path (decode.cu): f32 activations / residual, bf16 weights, split-precision KV
// pages, lowest-index argmax.  Reference being replaced: the per-token loop of `Model::generate`
// (crane-core/src/models/qwen3/model.rs:298-331) over `Qwen3Model::forward` (qwen3/modeling.rs:942-1036) and the server's
// decode rounds (crane-serve/src/engine/mod.rs:898-1008).
#include "decode_ll.cuh"

#include <algorithm>
#include <type_traits>

namespace cb {

#ifndef LL_POLL_LIMIT
#define LL_POLL_LIMIT (1u << 22)        // failed polls before a wait is declared dead (~seconds); the launch then drains
#endif
constexpr int LL_SEG_BYTES = 2048;      // ring slot = 4 chunks of 512 B (one chunk = 256 bf16 of one row)
constexpr int LL_MAX_SPLIT = 32;        // attention splits per KV head (one per lane in the merge)
constexpr int LL_MAX_NREP = 8;

__device__ __forceinline__ void ll_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ll_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ll_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- (value, tag) pairs ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_pair(unsigned long long* p, float v, uint32_t tag) {
    const unsigned long long u = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(u) : "memory");
}
__device__ __forceinline__ void ld_pair2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ld_pair1(const unsigned long long* p) {
    unsigned long long a;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ uint32_t pair_tag(unsigned long long u) { return (uint32_t)(u >> 32); }
__device__ __forceinline__ float pair_val(unsigned long long u) { return __uint_as_float((uint32_t)u); }

// Bounded waiting: a wait that never completes (a bug, or a peer that died) must not hang the GPU.  After LL_POLL_LIMIT failed
// polls the thread raises *err; every waiter gives up as soon as it sees the flag, and the launch drains with garbage results
// that the host discards (CRANE_B200_CUDA_ERROR).
struct Waiter {
    unsigned int* err;
    bool dead;
    __device__ __forceinline__ bool again(uint32_t& it) {
        if (dead) return false;
        if ((++it & 255u) == 0u) {
            unsigned int e = *reinterpret_cast<volatile unsigned int*>(err);
            if (it >= LL_POLL_LIMIT) { atomicExch(err, 1u); e = 1u; }
            if (__any_sync(0xffffffffu, e != 0u)) { dead = true; return false; }
        }
        return true;
    }
};

// Sum 8 per-lane values over the warp in 7 + 2 shuffles (recursive halving): on return v[0] of every lane holds the warp total
// of value `idx` (lanes with equal idx hold the same total; idx = (lane >> 2) bit-reversed over 3 bits, see below).
__device__ __forceinline__ void ll_reduce_scatter8(float (&v)[8], int lane, int& idx) {
    idx = 0;
    int off = 16;
#pragma unroll
    for (int n = 8; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        if (up) idx += n / 2;
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- phase geometry --------------------------------------------------------------------------------------------------
struct PDesc {                 // one weight matrix as seen by this CTA
    const unsigned char* W;
    int N, K, rpc, r0, nrows;
    LLGeom g;
};
__device__ __forceinline__ PDesc phase_desc(const LLArgs& a, int p) {      // p = 0 .. 4L (4L = lm_head)
    PDesc d;
    int rpu = 1;
    if (p >= 4 * a.L) { d.W = reinterpret_cast<const unsigned char*>(a.lm_head); d.N = a.V; d.K = a.H; }
    else {
        const LLLayer& l = a.layers[p >> 2];
        switch (p & 3) {
            case 0: d.W = reinterpret_cast<const unsigned char*>(l.wqkv); d.N = a.qkv_dim; d.K = a.H; break;
            case 1: d.W = reinterpret_cast<const unsigned char*>(l.wo); d.N = a.H; d.K = a.q_dim; break;
            case 2: d.W = reinterpret_cast<const unsigned char*>(l.wgu); d.N = 2 * a.I; d.K = a.H; rpu = 2; break;
            default: d.W = reinterpret_cast<const unsigned char*>(l.wdown); d.N = a.H; d.K = a.I; break;
        }
    }
    const int units = d.N / rpu;
    const int upc = (units + (int)gridDim.x - 1) / (int)gridDim.x;
    d.rpc = upc * rpu;
    d.r0 = (int)blockIdx.x * d.rpc;
    d.nrows = max(0, min(d.N, d.r0 + d.rpc) - d.r0);
    ll_geom(d.K, d.g);
    return d;
}
struct WDesc {                 // ... and by one warp: chunk q of the warp = row (rl + RL * (q / ncc)), column chunk (cg * ncc + q % ncc)
    const unsigned char* base;
    uint32_t nq, nmy;
    size_t row_stride;
    int cg, rl;
};
__device__ __forceinline__ WDesc warp_desc(const PDesc& d, int warp, int lane) {
    WDesc w;
    w.cg = warp % d.g.G;
    w.rl = warp / d.g.G;
    w.nmy = d.nrows > w.rl ? (uint32_t)((d.nrows - w.rl + d.g.RL - 1) / d.g.RL) : 0u;
    w.nq = w.nmy * (uint32_t)d.g.ncc;
    w.base = d.W + ((size_t)(d.r0 + w.rl) * d.K + (size_t)w.cg * d.g.ncc * 256) * 2 + lane * 16;
    w.row_stride = (size_t)d.g.RL * d.K * 2;
    return w;
}

// prefetch cursor of one warp: runs over the phases of all steps, independent of what the warp is consuming
struct Cursor {
    const unsigned char* base;
    size_t row_stride, row_off;
    uint32_t nq, q, cc, ncc;
    uint32_t gp, gp_end;        // global phase index (step * n_phase + p) and its end
    uint32_t seq;               // non-empty slots issued so far
};

// Position the cursor on the first phase (from pf.gp on) in which this warp owns chunks.  Out of line: it runs once per phase
// per warp, and the streaming loops that call it stay small.
__device__ __noinline__ void cursor_enter(const LLArgs& a, Cursor& pf, int n_phase, int warp, int lane) {
    for (;;) {
        if (pf.gp >= pf.gp_end) { pf.nq = 0; pf.q = 0; return; }
        const PDesc d = phase_desc(a, (int)(pf.gp % (uint32_t)n_phase));
        const WDesc w = warp_desc(d, warp, lane);
        if (w.nq != 0) {
            pf.base = w.base; pf.row_stride = w.row_stride; pf.row_off = 0; pf.nq = w.nq; pf.q = 0; pf.cc = 0;
            pf.ncc = (uint32_t)d.g.ncc;
            return;
        }
        ++pf.gp;
    }
}

// ---- activation slice of a warp: poll the pairs of its 256 * NCC columns into registers -------------------------------
template <int NCC>
__device__ __forceinline__ void load_x(const unsigned long long* buf, uint32_t tag, int cg, int lane, const float* norm_w,
                                       float (&xr)[NCC][8], float& ssq, Waiter& wt) {
    float g[NCC][8];
    if (norm_w != nullptr) {          // immutable: in flight while the activations are still being produced
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
            const float4* gp = reinterpret_cast<const float4*>(norm_w + (size_t)(cg * NCC + cc) * 256 + lane * 8);
            const float4 g0 = gp[0], g1 = gp[1];
            g[cc][0] = g0.x; g[cc][1] = g0.y; g[cc][2] = g0.z; g[cc][3] = g0.w;
            g[cc][4] = g1.x; g[cc][5] = g1.y; g[cc][6] = g1.z; g[cc][7] = g1.w;
        }
    }
    uint32_t it = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
            const unsigned long long* p = buf + (size_t)(cg * NCC + cc) * 256 + lane * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned long long u0, u1;
                ld_pair2(p + 2 * j, u0, u1);
                ok = ok && pair_tag(u0) == tag && pair_tag(u1) == tag;
                xr[cc][2 * j] = pair_val(u0);
                xr[cc][2 * j + 1] = pair_val(u1);
            }
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (!wt.again(it)) break;
    }
    ssq = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ssq = fmaf(xr[cc][j], xr[cc][j], ssq);
            if (norm_w != nullptr) xr[cc][j] *= g[cc][j];
        }
}

// =====================================================================================================================
template <int DEPTH>
__global__ void __launch_bounds__(LL_THREADS, 1)
decode_ll_kernel(const __grid_constant__ LLArgs a, int max_rows) {
    constexpr int D = 128;
    extern __shared__ __align__(1024) unsigned char lsm[];
    // dynamic: [rings: LL_WARPS * DEPTH * 2 KB][acc: 2 x max_rows f32][xres: residual rows owned by this CTA]
    float* acc_s = reinterpret_cast<float*>(lsm + (size_t)LL_WARPS * DEPTH * LL_SEG_BYTES);
    float* xres_s = acc_s + 2 * max_rows;
    __shared__ float ssq_s[4];
    __shared__ __align__(16) float q_s[LL_MAX_NREP][D];
    __shared__ __align__(16) float knew_s[D];
    __shared__ __align__(16) float vnew_s[D];
    __shared__ __align__(16) float wp_s[LL_WARPS][2][LL_PART_STRIDE];
    __shared__ float wbest_v[LL_WARPS];
    __shared__ int wbest_i[LL_WARPS];
    __shared__ uint32_t tok_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const int n_phase = 4 * a.L + 1;
    const int NREP = a.nh / a.nkv;
    const uint32_t tag0 = a.tag_base;
    const uint32_t tstride = (uint32_t)(a.L + 2);
    Waiter wt{a.err, false};

    long long tprof = clock64();
    auto prof = [&](int slot) {
        if (a.prof != nullptr && cta == 0 && tid == 0) {
            const long long now = clock64();
            a.prof[slot] += (unsigned long long)(now - tprof);
            tprof = now;
        }
    };

    // ---- prefetch cursor ---------------------------------------------------------------------------------------------
    unsigned char* ring = lsm + (size_t)warp * DEPTH * LL_SEG_BYTES + lane * 16;
    const uint32_t ring_u32 = smem_u32(ring);
    Cursor pf;
    pf.gp = 0; pf.gp_end = (uint32_t)a.n_steps * (uint32_t)n_phase; pf.seq = 0;
    auto pf_enter = [&]() { cursor_enter(a, pf, n_phase, warp, lane); };
    auto issue_next = [&]() {         // one ring slot (<= 4 chunks, never across a phase boundary) + exactly one commit
        if (pf.q < pf.nq) {
            const uint32_t dst = ring_u32 + (pf.seq % DEPTH) * LL_SEG_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (pf.q < pf.nq) {
                    ll_cp16(dst + c * 512, pf.base + pf.row_off + (size_t)pf.cc * 512);
                    ++pf.q;
                    if (++pf.cc == pf.ncc) { pf.cc = 0; pf.row_off += pf.row_stride; }
                }
            }
            ++pf.seq;
            if (pf.q >= pf.nq) { ++pf.gp; pf_enter(); }
        }
        ll_commit();
    };
    pf_enter();
#pragma unroll 1
    for (int i = 0; i < DEPTH; ++i) issue_next();
    uint32_t cseq = 0;                // slots consumed so far

    // ---- sequence state, residual slice, first input ----------------------------------------------------------------------
    const SeqState st0 = a.state[0];
    const int* bt = a.block_table + (size_t)st0.slot * a.max_pages;
    const int res_rpc = (a.H + grid - 1) / grid;
    const int res_r0 = cta * res_rpc;
    const int res_n = max(0, min(a.H, res_r0 + res_rpc) - res_r0);
    for (int i = tid; i < 2 * max_rows; i += LL_THREADS) acc_s[i] = 0.f;
    if (tid < 4) ssq_s[tid] = 0.f;
    for (int i = tid; i < res_n; i += LL_THREADS) {
        const float v = a.x_io[res_r0 + i];
        xres_s[i] = v;
        st_pair(a.xa + res_r0 + i, v, tag0 + 1u);          // input of layer 0 of step 0
    }
    __syncthreads();
    prof(0);

    uint32_t gphase = 0;              // global phase counter (parity of the shared accumulators)
#pragma unroll 1
    for (int s = 0; s < a.n_steps; ++s) {
        const int T = st0.kv_len + s + 1;                  // keys including the token being decoded
        const int pos3[3] = {st0.pos[0] + s, st0.pos[1] + s, st0.pos[2] + s};
        const uint32_t tstep = tag0 + (uint32_t)s * tstride;
#pragma unroll 1
        for (int p = 0; p < n_phase; ++p, ++gphase) {
            const int kind = (p == 4 * a.L) ? 4 : (p & 3);  // 0 qkv, 1 o-proj, 2 gate/up, 3 down, 4 lm_head
            const int l = min(p >> 2, a.L - 1);
            const uint32_t tagl = tstep + (uint32_t)(p >> 2) + 1u;      // tag of layer l's buffers; lm_head input = tag (s, L)

            // ============================ attention of layer l (between the QKV and the O projections) ============================
            if (kind == 1) {
                const LLLayer& ly = a.layers[l];
                const int nsplit = max(1, min(min(grid / a.nkv, LL_MAX_SPLIT), (T + 31) / 32));
                if (cta < a.nkv * nsplit) {
                    const int kvh = cta / nsplit, split = cta % nsplit;
                    const int chunk = (T + nsplit - 1) / nsplit;
                    const int t0 = split * chunk;
                    const int t_end = min(min(T, t0 + chunk), T - 1);   // cached tokens of this split (position T-1 comes from this step's qkv)
                    const bool owner = (split == (T - 1) / chunk);
                    const bool split_kv = a.kv_lo_off != 0;
                    const int half = lane >> 4, hl = lane & 15;
                    const int npass = t_end > t0 ? (t_end - t0 + 31) / 32 : 0;
                    uint4 kh = make_uint4(0, 0, 0, 0), kl = kh, vh = kh, vl = kh;
                    auto load_pass = [&](int ps, uint4& rkh, uint4& rkl, uint4& rvh, uint4& rvl) {
                        const int t = t0 + ps * 32 + warp * 2 + half;
                        rkh = make_uint4(0, 0, 0, 0); rkl = rkh; rvh = rkh; rvl = rkh;
                        if (t < t_end) {
                            const int page = __ldg(bt + t / KV_PAGE);
                            const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D + hl * 8;
                            rkh = ldg_stream(ly.k_pool + off);
                            rvh = ldg_stream(ly.v_pool + off);
                            if (split_kv) { rkl = ldg_stream(ly.k_pool + a.kv_lo_off + off); rvl = ldg_stream(ly.v_pool + a.kv_lo_off + off); }
                        }
                    };
                    if (npass > 0) load_pass(0, kh, kl, vh, vl);        // in flight while the query is still being produced
                    // ---- q (NREP heads), and on the owning split the new k / v: poll, RMSNorm, rotate ----
                    if (warp < NREP + 2 && (warp < NREP || owner)) {
                        const bool is_k = warp == NREP, is_v = warp == NREP + 1;
                        const unsigned long long* src = a.qkv + (is_v ? a.q_dim + a.nkv * D + kvh * D : is_k ? a.q_dim + kvh * D : (kvh * NREP + warp) * D);
                        const float* nw = is_k ? ly.kn : ly.qn;
                        float nwv[4], cs_c[4], cs_s[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            nwv[j] = is_v ? 1.f : nw[lane + 32 * j];
                            const int i = lane + 32 * (j & 1);           // rotary pair (j, j + 2): column i = lane + 32 * (j mod 2)
                            const int pp = pos3[a.axis_of[i]];
                            cs_c[j] = a.cos_tab[(size_t)pp * (D / 2) + i];
                            cs_s[j] = a.sin_tab[(size_t)pp * (D / 2) + i];
                        }
                        float e[4];
                        uint32_t it = 0;
                        for (;;) {
                            bool ok = true;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const unsigned long long u = ld_pair1(src + lane + 32 * j);
                                ok = ok && pair_tag(u) == tagl;
                                e[j] = pair_val(u);
                            }
                            if (__all_sync(0xffffffffu, ok)) break;
                            if (!wt.again(it)) break;
                        }
                        if (is_v) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) vnew_s[lane + 32 * j] = split_kv ? round_bf16_split(e[j]) : round_bf16(e[j]);
                        } else {
                            float ssq = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
                            ssq = warp_sum(ssq);
                            const float rstd = rsqrtf(ssq / (float)D + a.eps);
#pragma unroll
                            for (int j = 0; j < 4; ++j) e[j] = e[j] * rstd * nwv[j];
                            float r[4];
                            r[0] = e[0] * cs_c[0] - e[2] * cs_s[0];
                            r[1] = e[1] * cs_c[1] - e[3] * cs_s[1];
                            r[2] = e[0] * cs_s[2] + e[2] * cs_c[2];
                            r[3] = e[1] * cs_s[3] + e[3] * cs_c[3];
                            float* dst = is_k ? knew_s : q_s[warp];
#pragma unroll
                            for (int j = 0; j < 4; ++j) dst[lane + 32 * j] = is_k ? (split_kv ? round_bf16_split(r[j]) : round_bf16(r[j])) : r[j];
                        }
                    }
                    __syncthreads();
                    // ---- head groups of <= 2 query heads share every K/V register ----
                    const int HG = (NREP & 1) ? 1 : 2;
#pragma unroll 1
                    for (int hg0 = 0; hg0 < NREP; hg0 += HG) {
                        float qr[2][8], m[2], ls[2], o[2][8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            m[h] = -INFINITY; ls[h] = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) { o[h][j] = 0.f; qr[h][j] = (h < HG) ? q_s[hg0 + h][hl * 8 + j] * a.scale : 0.f; }
                        }
                        if (hg0 != 0 && npass > 0) load_pass(0, kh, kl, vh, vl);
#pragma unroll 1
                        for (int ps = 0; ps < npass; ++ps) {
                            uint4 nkh = make_uint4(0, 0, 0, 0), nkl = nkh, nvh = nkh, nvl = nkh;
                            if (ps + 1 < npass) load_pass(ps + 1, nkh, nkl, nvh, nvl);
                            const bool valid = (t0 + ps * 32 + warp * 2 + half) < t_end;
                            float kf[8], vf[8];
                            kf[0] = bf16lo(kh.x); kf[1] = bf16hi(kh.x); kf[2] = bf16lo(kh.y); kf[3] = bf16hi(kh.y);
                            kf[4] = bf16lo(kh.z); kf[5] = bf16hi(kh.z); kf[6] = bf16lo(kh.w); kf[7] = bf16hi(kh.w);
                            vf[0] = bf16lo(vh.x); vf[1] = bf16hi(vh.x); vf[2] = bf16lo(vh.y); vf[3] = bf16hi(vh.y);
                            vf[4] = bf16lo(vh.z); vf[5] = bf16hi(vh.z); vf[6] = bf16lo(vh.w); vf[7] = bf16hi(vh.w);
                            if (split_kv) {
                                kf[0] += bf16lo(kl.x); kf[1] += bf16hi(kl.x); kf[2] += bf16lo(kl.y); kf[3] += bf16hi(kl.y);
                                kf[4] += bf16lo(kl.z); kf[5] += bf16hi(kl.z); kf[6] += bf16lo(kl.w); kf[7] += bf16hi(kl.w);
                                vf[0] += bf16lo(vl.x); vf[1] += bf16hi(vl.x); vf[2] += bf16lo(vl.y); vf[3] += bf16hi(vl.y);
                                vf[4] += bf16lo(vl.z); vf[5] += bf16hi(vl.z); vf[6] += bf16lo(vl.w); vf[7] += bf16hi(vl.w);
                            }
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h < HG) {
                                    float sc = 0.f;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) sc = fmaf(qr[h][j], kf[j], sc);
#pragma unroll
                                    for (int ofs = 8; ofs > 0; ofs >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, ofs);
                                    if (valid) {
                                        const float mn = fmaxf(m[h], sc);
                                        const float corr = __expf(m[h] - mn), pr = __expf(sc - mn);
                                        ls[h] = ls[h] * corr + pr;
#pragma unroll
                                        for (int j = 0; j < 8; ++j) o[h][j] = o[h][j] * corr + pr * vf[j];
                                        m[h] = mn;
                                    }
                                }
                            }
                            kh = nkh; kl = nkl; vh = nvh; vl = nvl;
                        }
                        // the token being decoded: half 0 of warp 0 on the owning split
                        if (owner && warp == 0) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h < HG) {
                                    float sc = 0.f;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) sc = fmaf(qr[h][j], knew_s[hl * 8 + j], sc);
#pragma unroll
                                    for (int ofs = 8; ofs > 0; ofs >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, ofs);
                                    if (half == 0) {
                                        const float mn = fmaxf(m[h], sc);
                                        const float corr = __expf(m[h] - mn), pr = __expf(sc - mn);
                                        ls[h] = ls[h] * corr + pr;
#pragma unroll
                                        for (int j = 0; j < 8; ++j) o[h][j] = o[h][j] * corr + pr * vnew_s[hl * 8 + j];
                                        m[h] = mn;
                                    }
                                }
                            }
                        }
                        // merge the two token halves of the warp, then the warps of the CTA through shared memory
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float mo = __shfl_xor_sync(0xffffffffu, m[h], 16);
                            const float lo = __shfl_xor_sync(0xffffffffu, ls[h], 16);
                            const float mn = fmaxf(m[h], mo);
                            const float c0 = (m[h] == -INFINITY) ? 0.f : __expf(m[h] - mn);
                            const float c1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
                            ls[h] = ls[h] * c0 + lo * c1;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float oo = __shfl_xor_sync(0xffffffffu, o[h][j], 16);
                                o[h][j] = o[h][j] * c0 + oo * c1;
                            }
                            m[h] = mn;
                            if (h < HG && half == 0) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) wp_s[warp][h][hl * 8 + j] = o[h][j];
                                if (hl == 0) { wp_s[warp][h][128] = m[h]; wp_s[warp][h][129] = ls[h]; }
                            }
                        }
                        __syncthreads();
                        for (int idx = tid; idx < HG * D; idx += LL_THREADS) {
                            const int h = idx / D, i = idx % D;
                            float M = -INFINITY;
#pragma unroll
                            for (int w = 0; w < LL_WARPS; ++w) M = fmaxf(M, wp_s[w][h][128]);
                            float Ls = 0.f, O = 0.f;
#pragma unroll
                            for (int w = 0; w < LL_WARPS; ++w) {
                                const float mw = wp_s[w][h][128];
                                const float c = (mw == -INFINITY) ? 0.f : __expf(mw - M);
                                Ls += wp_s[w][h][129] * c;
                                O += wp_s[w][h][i] * c;
                            }
                            unsigned long long* dst = a.part + ((size_t)cta * NREP + hg0 + h) * LL_PART_STRIDE;
                            st_pair(dst + i, O, tagl);
                            if (i == 0) { st_pair(dst + 128, M, tagl); st_pair(dst + 129, Ls, tagl); }
                        }
                        __syncthreads();
                    }
                    if (owner) {          // append the new token to its page (read from the next step on)
                        const int t = T - 1;
                        const int page = bt[t / KV_PAGE];
                        const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D;
                        for (int i = tid; i < D; i += LL_THREADS) {
                            const bf16 kb = __float2bfloat16_rn(knew_s[i]), vb = __float2bfloat16_rn(vnew_s[i]);
                            ly.k_pool[off + i] = kb;
                            ly.v_pool[off + i] = vb;
                            if (split_kv) {
                                ly.k_pool[a.kv_lo_off + off + i] = __float2bfloat16_rn(knew_s[i] - __bfloat162float(kb));
                                ly.v_pool[a.kv_lo_off + off + i] = __float2bfloat16_rn(vnew_s[i] - __bfloat162float(vb));
                            }
                        }
                        __threadfence();
                        __syncthreads();  // knew_s / vnew_s are rewritten by the next layer's item
                    }
                }
                prof(4);
                // ---- split merge of head `cta`: one split per lane, 8 output columns per warp ----
                if (cta < a.nh) {
                    const int h = cta, kvh = h / NREP, hh = h % NREP;
                    const bool have = lane < nsplit;
                    const unsigned long long* src = a.part + ((size_t)(kvh * nsplit + (have ? lane : 0)) * NREP + hh) * LL_PART_STRIDE;
                    float ov[8], ms = -INFINITY, lsum = 0.f;
                    uint32_t it = 0;
                    for (;;) {
                        bool ok = true;
                        if (have) {
                            unsigned long long u0, u1;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                ld_pair2(src + warp * 8 + 2 * j, u0, u1);
                                ok = ok && pair_tag(u0) == tagl && pair_tag(u1) == tagl;
                                ov[2 * j] = pair_val(u0); ov[2 * j + 1] = pair_val(u1);
                            }
                            ld_pair2(src + 128, u0, u1);
                            ok = ok && pair_tag(u0) == tagl && pair_tag(u1) == tagl;
                            ms = pair_val(u0); lsum = pair_val(u1);
                        }
                        if (__all_sync(0xffffffffu, ok)) break;
                        if (!wt.again(it)) break;
                    }
                    if (!have) {
                        ms = -INFINITY; lsum = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) ov[j] = 0.f;
                    }
                    const float M = warp_max(ms);
                    const float c = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                    const float Ls = warp_sum(lsum * c);
                    float outv = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float t = warp_sum(ov[j] * c);
                        if (lane == j) outv = t;
                    }
                    if (lane < 8) st_pair(a.att + (size_t)h * D + warp * 8 + lane, outv / Ls, tagl);
                }
                prof(5);
            }

            // ============================ weight phase: y = W . x over this CTA's row block ============================
            const PDesc d = phase_desc(a, p);
            const WDesc w = warp_desc(d, warp, lane);
            const LLLayer& ly = a.layers[l];
            const float* nw = (kind == 0) ? ly.ln1 : (kind == 2) ? ly.ln2 : (kind == 4) ? a.final_norm : nullptr;
            const unsigned long long* xin = (kind == 0 || kind == 4) ? a.xa : (kind == 1) ? a.att : (kind == 2) ? a.xb : a.act;
            float* acc = acc_s + (gphase & 1u) * max_rows;
            float* ssq_p = &ssq_s[gphase & 3u];

            auto run = [&](auto ncc_tag) {
                constexpr int NCC = decltype(ncc_tag)::value;
                float xr[NCC][8];
                float ssq = 0.f;
                const bool stat = nw != nullptr && w.rl == 0;     // this warp's columns count towards the RMSNorm statistic
                if (w.nq == 0 && !stat) return;                    // no rows of this matrix for this warp
                load_x<NCC>(xin, tagl, w.cg, lane, nw, xr, ssq, wt);
                if (stat) {
                    ssq = warp_sum(ssq);
                    if (lane == 0) atomicAdd(ssq_p, ssq);
                }
                prof(1);
                uint32_t q = 0;
#pragma unroll 1
                for (uint32_t b = 0; q < w.nq; ++b) {            // batches of 8 rows of this warp = 2 * NCC ring slots
                    float av[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) av[i] = 0.f;
#pragma unroll
                    for (int sl = 0; sl < 2 * NCC; ++sl) {
                        if (q < w.nq) {                            // warp-uniform
                            ll_wait<DEPTH - 1>();
                            __syncwarp();
                            const unsigned char* sp = ring + (cseq % DEPTH) * LL_SEG_BYTES;
                            uint4 wv[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (q + c < w.nq) wv[c] = *reinterpret_cast<const uint4*>(sp + c * 512);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int i = sl * 4 + c;          // compile-time: row i / NCC of the batch, column chunk i % NCC
                                if (q + c < w.nq) {
                                    const float* xv = xr[i % NCC];
                                    float t = av[i / NCC];
                                    t = fmaf(bf16lo(wv[c].x), xv[0], t); t = fmaf(bf16hi(wv[c].x), xv[1], t);
                                    t = fmaf(bf16lo(wv[c].y), xv[2], t); t = fmaf(bf16hi(wv[c].y), xv[3], t);
                                    t = fmaf(bf16lo(wv[c].z), xv[4], t); t = fmaf(bf16hi(wv[c].z), xv[5], t);
                                    t = fmaf(bf16lo(wv[c].w), xv[6], t); t = fmaf(bf16hi(wv[c].w), xv[7], t);
                                    av[i / NCC] = t;
                                }
                            }
                            q += 4;
                            __syncwarp();                          // every lane has its slot data in registers
                            issue_next();
                            ++cseq;
                        }
                    }
                    int idx;
                    ll_reduce_scatter8(av, lane, idx);
                    const uint32_t rloc = b * 8u + (uint32_t)idx;
                    if ((lane & 3) == 0 && rloc < w.nmy) atomicAdd(&acc[w.rl + d.g.RL * (int)rloc], av[0]);
                }
            };
            switch (d.g.ncc) {
                case 1: run(std::integral_constant<int, 1>{}); break;
                case 2: run(std::integral_constant<int, 2>{}); break;
                case 3: run(std::integral_constant<int, 3>{}); break;
                default: run(std::integral_constant<int, 4>{}); break;
            }
            prof(2);
            __syncthreads();
            prof(3);

            // ============================ epilogue: publish this CTA's rows ============================
            const float rstd = (nw != nullptr) ? rsqrtf(*ssq_p / (float)d.K + a.eps) : 1.f;
            if (tid == 0) ssq_s[(gphase + 2u) & 3u] = 0.f;          // last read two phases ago, next written two phases from now
            if (kind == 0) {
                for (int i = tid; i < d.nrows; i += LL_THREADS) {
                    st_pair(a.qkv + d.r0 + i, acc[i] * rstd, tagl);
                    acc[i] = 0.f;
                }
            } else if (kind == 1 || kind == 3) {
                unsigned long long* dst = (kind == 1) ? a.xb : a.xa;
                const uint32_t tg = (kind == 1) ? tagl : tagl + 1u;   // down-proj output = input of the next layer (or of the lm_head)
                for (int i = tid; i < d.nrows; i += LL_THREADS) {
                    const float v = xres_s[i] + acc[i];
                    xres_s[i] = v;
                    st_pair(dst + d.r0 + i, v, tg);
                    acc[i] = 0.f;
                }
            } else if (kind == 2) {
                for (int u = tid; u < d.nrows / 2; u += LL_THREADS) {
                    const float g = acc[2 * u] * rstd, up = acc[2 * u + 1] * rstd;
                    st_pair(a.act + d.r0 / 2 + u, silu_f(g) * up, tagl);
                    acc[2 * u] = 0.f; acc[2 * u + 1] = 0.f;
                }
            } else {
                // logits + (value, lowest index) argmax: per thread rows ascend, so the first maximum wins
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int i = tid; i < d.nrows; i += LL_THREADS) {
                    const float v = acc[i] * rstd;
                    a.logits[d.r0 + i] = v;
                    acc[i] = 0.f;
                    if (v > bv) { bv = v; bi = d.r0 + i; }
                }
#pragma unroll
                for (int ofs = 16; ofs > 0; ofs >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, ofs);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, ofs);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) { wbest_v[warp] = bv; wbest_i[warp] = bi; }
                __syncthreads();
                if (tid == 0) {
                    for (int ww = 1; ww < LL_WARPS; ++ww)
                        if (wbest_v[ww] > bv || (wbest_v[ww] == bv && wbest_i[ww] < bi)) { bv = wbest_v[ww]; bi = wbest_i[ww]; }
                    const uint32_t tg = tstep + (uint32_t)a.L + 2u;
                    st_pair(a.amax + 2 * cta, bv, tg);
                    st_pair(a.amax + 2 * cta + 1, __int_as_float(bi), tg);
                }
            }
            prof(6);
        }

        // ============================ token: every CTA reduces the per-CTA maxima itself (no broadcast hop) ============================
        if (warp == 0) {
            const uint32_t tg = tstep + (uint32_t)a.L + 2u;
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int c0 = 0; c0 < grid; c0 += 32) {
                const int c = c0 + lane;
                float v = -INFINITY; int ix = 0x7fffffff;
                uint32_t it = 0;
                for (;;) {
                    bool ok = true;
                    if (c < grid) {
                        unsigned long long u0, u1;
                        ld_pair2(a.amax + 2 * c, u0, u1);
                        ok = pair_tag(u0) == tg && pair_tag(u1) == tg;
                        v = pair_val(u0); ix = __float_as_int(pair_val(u1));
                    }
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (!wt.again(it)) break;
                }
                if (c < grid && (v > bv || (v == bv && ix < bi))) { bv = v; bi = ix; }
            }
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, ofs);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, ofs);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if ((unsigned)bi >= (unsigned)a.V) bi = 0;              // all-NaN logits must not become an out-of-range gather
            if (lane == 0) tok_s = (uint32_t)bi;
        }
        __syncthreads();
        const uint32_t tok = tok_s;
        if (cta == 0 && tid == 0 && a.out_tokens != nullptr) a.out_tokens[st0.step + s] = tok;
        if (a.advance) {                                             // next input: this CTA's slice of the embedding row
            const bf16* rowp = a.embed + (size_t)tok * a.H + res_r0;
            const bool more = s + 1 < a.n_steps;
            for (int i = tid; i < res_n; i += LL_THREADS) {
                const float v = __bfloat162float(rowp[i]);
                xres_s[i] = v;
                if (more) st_pair(a.xa + res_r0 + i, v, tstep + tstride + 1u);
                else a.x_io[res_r0 + i] = v;
            }
        }
        __syncthreads();                                             // tok_s is rewritten by the next step
        prof(7);
    }
    if (cta == 0 && tid == 0) {
        SeqState* sp = a.state;
        sp->step = st0.step + a.n_steps;
        if (a.advance) {
            sp->kv_len = st0.kv_len + a.n_steps;
            sp->pos[0] = st0.pos[0] + a.n_steps; sp->pos[1] = st0.pos[1] + a.n_steps; sp->pos[2] = st0.pos[2] + a.n_steps;
            sp->token = tok_s;
        }
    }
    ll_wait<0>();
}

// ---- host ------------------------------------------------------------------------------------------------------------
static int ll_max_rows(const LLArgs& a, int grid) {
    auto rpc = [&](int N, int rpu) { return ((N / rpu) + grid - 1) / grid * rpu; };
    return std::max(std::max(rpc(a.qkv_dim, 1), rpc(a.H, 1)), std::max(rpc(2 * a.I, 2), rpc(a.V, 1)));
}
static size_t ll_smem(const LLArgs& a, int grid, int depth) {
    const int mr = ll_max_rows(a, grid);
    return (size_t)LL_WARPS * depth * LL_SEG_BYTES + (size_t)2 * mr * 4 + (size_t)((a.H + grid - 1) / grid) * 4 + 16;
}

bool decode_ll_supported(int D, int rot_half, int nh, int nkv, int H, int I, int q_dim, int qkv_dim, int V, int num_sms) {
    LLGeom g;
    if (D != 128 || rot_half != 64 || nkv <= 0 || nh % nkv || nh / nkv > LL_MAX_NREP || nh > num_sms || nkv > num_sms) return false;
    if (!ll_geom(H, g) || !ll_geom(I, g) || !ll_geom(q_dim, g)) return false;
    if ((2 * I) % 2 || qkv_dim <= 0 || V <= 0) return false;
    LLArgs a = {};
    a.H = H; a.I = I; a.V = V; a.qkv_dim = qkv_dim;
    return ll_smem(a, num_sms, 4) + 24 * 1024 <= 227 * 1024;       // static shared memory: ~22.5 KB
}

size_t decode_ll_part_pairs(int num_sms, int nh, int nkv) { return (size_t)num_sms * (size_t)(nh / nkv) * LL_PART_STRIDE; }

template <int DEPTH>
static int ll_launch_t(cudaStream_t st, const LLArgs& a, int grid, size_t smem, int max_rows) {
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(decode_ll_kernel<DEPTH>, smem, seen)) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(LL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // every CTA must be resident: consumers spin on their producers
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, decode_ll_kernel<DEPTH>, a, max_rows);
}

int decode_ll_launch(cudaStream_t st, const LLArgs& a, int num_sms) {
    const int mr = ll_max_rows(a, num_sms);
    static const int forced = [] { const char* e = getenv("CRANE_B200_LL_DEPTH"); return e ? atoi(e) : 0; }();
    const size_t budget = 227 * 1024 - 23 * 1024;      // dynamic part: the kernel's static shared memory is ~22.5 KB
    for (int depth : {6, 5, 4}) {
        if (forced && depth != forced) continue;
        const size_t smem = ll_smem(a, num_sms, depth);
        if (smem > budget) continue;
        switch (depth) {
            case 6: return ll_launch_t<6>(st, a, num_sms, smem, mr);
            case 5: return ll_launch_t<5>(st, a, num_sms, smem, mr);
            default: return ll_launch_t<4>(st, a, num_sms, smem, mr);
        }
    }
    return -1000;
}

}  // namespace cb
