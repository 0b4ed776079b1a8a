// Persistent decode kernel: ALL phases of n greedy decode steps of one sequence in ONE cooperative launch, synchronised by
// flag-carrying activations instead of grid barriers.
//
// Why.  A batch-1 decode step of Qwen3-VL-2B is 141 dependent bandwidth-bound phases (113 weight matrices + 28 attentions) of
// 1.3-96 us of HBM time each.  As separate kernels every phase boundary costs ~2.5 us during which the HBM pipe is idle
// (round 1: 50 % of the HBM roofline); a grid barrier per phase costs about the same.  Here:
//   * one CTA per SM stays resident; each of its 8 warps owns a fixed column range of every weight matrix and streams it, in
//     tiles of 16 rows x 64 columns, through a private cp.async ring (12 x 2 KB per warp, 192 KB per SM).  The prefetch cursor
//     is independent of the consumption cursor and runs ahead across phase, layer and token boundaries: while a warp waits
//     for its input activations the next matrices keep arriving, so a dependency stall shorter than the ring (~4 us of HBM
//     time) costs no bandwidth;
//   * the products run on the tensor cores -- not for flops (the step is HBM-bound) but for instruction count: a 2 KB tile is
//     4 ldmatrix + 4 mma.sync (m16n8k16, weights = A from the ring, activations = the B columns (hi, lo) of a split-bf16
//     pair, f32 accumulate) where FMA lanes need ~110 instructions, so whatever landed during a wait is consumed at
//     shared-memory speed and the stream catches up with its prefetch;
//   * activations cross SMs as 8-byte (value, tag) pairs written with one store and polled by their consumers (the NCCL
//     "LL" protocol applied to a GEMV chain): no separate flag, no fence, no barrier -- one L2 round trip per dependency.  A
//     warp polls only the columns it owns and keeps them in registers as B fragments for the whole phase (no shared-memory
//     staging, no block barrier at phase start); RMSNorm is folded in (weights multiplied at load, 1/rms applied in the
//     epilogue);
//   * attention runs as nkv x nsplit CTA items (K/V straight from the pages into registers, loads issued before the query is
//     polled), partials are published as pairs and merged by the CTA that owns the head (one split per lane, shuffles).
// Arithmetic: f32 residual stream and accumulation, bf16 weights (exact), activations enter the products as hi + lo bf16 pairs
// (~16 mantissa bits, the engine's split precision -- the multi-kernel path of decode.cu multiplies by the f32 value itself),
// split-precision KV pages, lowest-index argmax.  Reference being replaced: the per-token loop of `Model::generate`
// (crane-core/src/models/qwen3/model.rs:298-331) over `Qwen3Model::forward` (qwen3/modeling.rs:942-1036) and the server's
// decode rounds (crane-serve/src/engine/mod.rs:898-1008).
#include "decode_ll.cuh"

#include <algorithm>
#include <type_traits>

namespace cb {

#ifndef LL_WAIT_CYCLES
#define LL_WAIT_CYCLES (1ll << 31)      // SM cycles (~1 s) before a wait is declared dead; the launch then drains
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

// Bounded waiting: a wait that never completes (a bug, or a peer that died) must not hang the GPU.  After LL_WAIT_CYCLES
// cycles of waiting the thread raises *err; every waiter gives up as soon as it sees the flag, and the launch drains with garbage results
// that the host discards (CRANE_B200_CUDA_ERROR).
struct Waiter {
    unsigned int* err;         // [0] flag, [1..7] who gave up first: site, cta, warp, step, phase, tag wanted, tag seen;
                               // [16 + cta * LL_WARPS + warp] what each warp was last seen waiting for: step << 20 | phase << 8 | site
    bool dead;
    uint32_t where;            // step << 12 | phase (diagnostics only)
    __device__ __forceinline__ bool again(uint32_t& it, long long& t0, uint32_t site, uint32_t want, uint32_t seen) {
        if (dead) return false;
        if ((++it & 255u) == 0u) {
            if (it == 256u) {
                t0 = clock64();
                if ((threadIdx.x & 31) == 0) err[16 + blockIdx.x * LL_WARPS + (threadIdx.x >> 5)] = ((where >> 12) << 20) | ((where & 0xfffu) << 8) | site;
            }
            unsigned int e = *reinterpret_cast<volatile unsigned int*>(err);
            if (e == 0u && clock64() - t0 > LL_WAIT_CYCLES) {
                if (atomicCAS(err, 0u, 1u) == 0u) {
                    err[1] = site; err[2] = blockIdx.x; err[3] = threadIdx.x >> 5; err[4] = where >> 12; err[5] = where & 0xfffu; err[6] = want; err[7] = seen;
                }
                e = 1u;
            }
            if (__any_sync(0xffffffffu, e != 0u)) { dead = true; return false; }
        }
        return true;
    }
};

// Shared row accumulators.  The eight warps of a CTA own different columns of the same rows and finish a tile at the same moment;
// atomicAdd(float) -- and the 64-bit integer add -- on shared memory are CAS spin loops, only the 32-bit integer add is native.  So a
// row sum is kept as two independent 32-bit fixed-point words, v = A * 2^-8 + B * 2^-32 with A = round(v * 2^8) and B the exact
// remainder: each addend keeps its full f32 mantissa (down to 2^-32), no carry links the two words, integer addition is exact and
// therefore independent of arrival order.  |sum| < 2^23 (A) and at most 2^8 addends (B): far beyond any activation.
struct FixAcc { int a, b; };
__device__ __forceinline__ void fix_add(FixAcc* p, float v) {
    const int A = __float2int_rn(v * 256.0f);
    const int B = __float2int_rn((v - (float)A * (1.0f / 256.0f)) * 4294967296.0f);
    atomicAdd(&p->a, A);
    atomicAdd(&p->b, B);
}
__device__ __forceinline__ float fix_take(FixAcc* p) {       // read and clear (one reader per row, after the CTA barrier)
    const FixAcc v = *p;
    *p = FixAcc{0, 0};
    return (float)v.a * (1.0f / 256.0f) + (float)v.b * (1.0f / 4294967296.0f);      // (float)a is exact below 2^24 (|v| < 65 536)
}

// ---- phase geometry --------------------------------------------------------------------------------------------------
// A weight matrix [N, K] as seen by one CTA (row block [r0, r0 + nrows)) and one warp (columns [col0, col0 + 64 * ncs)): the warp
// walks its block in tiles of 16 rows x 64 columns = one 2 KB ring slot = four m16n8k16 tensor-core steps.  Only five distinct
// shapes exist (qkv, o, gate/up, down, lm_head): computed once per launch into shared memory, so switching phase costs a table read.
struct WGeo {
    unsigned long long base_off;   // bytes from the matrix start to (row r0, column col0)
    uint32_t nslots;               // ring slots of this warp in the phase: row tiles x ncs (0: the warp idles -- K < 64 * warps)
    int r0, nrows, ntiles;         // row block of the CTA, in tiles of 16 rows
    int K, ncs, col0;              // ncs = 64-column steps per warp
    int pad;
};
__host__ __device__ inline int ll_active_warps(int K) {        // warps that share the columns of a row: K = 64 * ncs * active
    int aw = LL_WARPS;
    while (aw > 1 && (K % (64 * aw)) != 0) aw >>= 1;
    return aw;
}
__device__ __forceinline__ int phase_kind(int p, int L) { return (p == 4 * L) ? 4 : (p & 3); }   // 0 qkv, 1 o, 2 gate/up, 3 down, 4 lm_head
__device__ __forceinline__ const unsigned char* phase_weights(const LLArgs& a, int p) {
    if (p >= 4 * a.L) return reinterpret_cast<const unsigned char*>(a.lm_head);
    const LLLayer& l = a.layers[p >> 2];
    const bf16* w = (p & 3) == 0 ? l.wqkv : (p & 3) == 1 ? l.wo : (p & 3) == 2 ? l.wgu : l.wdown;
    return reinterpret_cast<const unsigned char*>(w);
}
__device__ void fill_geometry(const LLArgs& a, WGeo* wg /* [5] of this warp */, int warp) {
    for (int kind = 0; kind < 5; ++kind) {
        int N, K, rpu = 1;
        switch (kind) {
            case 0: N = a.qkv_dim; K = a.H; break;
            case 1: N = a.H; K = a.q_dim; break;
            case 2: N = 2 * a.I; K = a.H; rpu = 2; break;
            case 3: N = a.H; K = a.I; break;
            default: N = a.V; K = a.H; break;
        }
        const int aw = ll_active_warps(K);
        const int units = N / rpu;
        const int upc = (units + (int)gridDim.x - 1) / (int)gridDim.x;
        const int rpc = upc * rpu;
        WGeo w;
        w.r0 = (int)blockIdx.x * rpc;
        w.nrows = max(0, min(N, w.r0 + rpc) - w.r0);
        w.ntiles = (w.nrows + 15) / 16;
        w.K = K; w.pad = 0;
        w.ncs = K / (64 * aw);
        w.col0 = warp * 64 * w.ncs;
        w.nslots = warp < aw ? (uint32_t)(w.ntiles * w.ncs) : 0u;
        w.base_off = ((unsigned long long)w.r0 * K + (unsigned long long)w.col0) * 2;
        wg[kind] = w;
    }
}

// prefetch cursor of one warp: runs over the phases of all steps, independent of what the warp is consuming
struct Cursor {
    const unsigned char* tile;  // (first row of the current row tile, first column of the warp) of the matrix being prefetched
    uint32_t row_bytes;         // K * 2
    int rows_left;              // rows of the CTA's block from the current tile on
    int cs, ncs;                // 64-column step inside the tile
    uint32_t left;              // slots of this phase still to issue
    int step, p;                // phase being prefetched; step == n_steps: past the end
    uint32_t slot;              // byte offset of the next ring slot to fill
};

// Position the cursor on the first phase (from (pf.step, pf.p) on) in which this warp owns slots.  Inlined on purpose: the
// cursor must stay in registers -- with 227 KB of shared memory carved out there is no L1 left, so anything that lives in local
// memory (a struct whose address is passed to a real call, a spilled register) costs an L2 round trip per access.
__device__ __forceinline__ void cursor_enter(const LLArgs& a, Cursor& pf, const WGeo* wg) {
    const int n_phase = 4 * a.L + 1;
    for (;;) {
        if (pf.p >= n_phase) { pf.p = 0; ++pf.step; }
        if (pf.step >= a.n_steps) { pf.left = 0; return; }
        const WGeo& w = wg[phase_kind(pf.p, a.L)];
        if (w.nslots != 0) {
            pf.tile = phase_weights(a, pf.p) + w.base_off;
            pf.row_bytes = (uint32_t)w.K * 2u; pf.rows_left = w.nrows; pf.cs = 0; pf.ncs = w.ncs; pf.left = w.nslots;
            return;
        }
        ++pf.p;
    }
}

// HBM -> L2 prefetch of a contiguous range (one thread).  The shared-memory rings can only run ~4 us ahead; while the CTAs sit in a
// dependency wait the rings are full and HBM would idle.  The weights of a phase a little further ahead are therefore pulled into
// the 126 MB L2 in the meantime, so that the ring later refills at L2 speed.
__device__ __forceinline__ void ll_prefetch_l2(const unsigned char* p, size_t bytes) {
    while (bytes != 0) {
        const uint32_t n = (uint32_t)(bytes < (size_t)(1u << 20) ? bytes : (size_t)(1u << 20));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(n) : "memory");
        p += n; bytes -= n;
    }
}

// ---- tensor-core pieces: A = 16 weight rows x 16 columns (ldmatrix from the ring slot), B = the activations as the columns
// (hi, lo, 0, ...) of a 16 x 8 operand, D = f32.  One mma.sync does 16 rows x 16 columns for both halves of the split activation.
__device__ __forceinline__ void ll_ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ll_mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- activation slice of a warp: poll the pairs of its 64 * NCS columns and shape them into tensor-core B fragments --------------
// B (k x n, "col") of m16n8k16: lane l holds rows k = 2 (l % 4) + {0, 1} (b0) and + {8, 9} (b1) of column n = l / 4.  Column 0 carries
// hi = bf16(x), column 1 lo = bf16(x - hi) (~16 mantissa bits together, the weights are exact in bf16), the other six are zero.
// Every lane fetches two adjacent pairs of each 64-column step (one 16-byte load, the whole step = 512 contiguous bytes per
// warp), folds the RMSNorm weight in, splits, and the fragments are dealt out with shuffles: lanes 0-3 receive the hi halves,
// lanes 4-7 the lo halves, the rest hold zeros.  The sum of squares of the raw values comes back through `ssq` (per lane).
template <int NCS, bool NORM>
__device__ __forceinline__ void load_xb(const unsigned long long* buf, uint32_t tag, int col0, int lane, const float* norm_w,
                                        uint32_t (&xb)[NCS * 4][2], float& ssq, Waiter& wt) {
    const int c_lane = col0 + 2 * lane;                   // this lane's two columns inside each 64-column step
    float2 g[NORM ? NCS : 1];
    if (NORM) {                                           // immutable: in flight while the activations are still being produced
#pragma unroll
        for (int cs = 0; cs < NCS; ++cs) g[cs] = *reinterpret_cast<const float2*>(norm_w + c_lane + 64 * cs);
    }
    float v0[NCS], v1[NCS];
    uint32_t it = 0;
    long long wt0 = 0;
    for (;;) {
        bool ok = true;
        uint32_t seen = tag;
#pragma unroll
        for (int cs = 0; cs < NCS; ++cs) {
            unsigned long long u0, u1;
            ld_pair2(buf + c_lane + 64 * cs, u0, u1);
            if (pair_tag(u0) != tag) { ok = false; seen = pair_tag(u0); }
            if (pair_tag(u1) != tag) { ok = false; seen = pair_tag(u1); }
            v0[cs] = pair_val(u0); v1[cs] = pair_val(u1);
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (!wt.again(it, wt0, 1u, tag, seen)) break;
    }
    ssq = 0.f;
    const int kq = lane & 3;
    const bool is_lo = (lane & 4) != 0, act = lane < 8;
#pragma unroll
    for (int cs = 0; cs < NCS; ++cs) {
        float a0 = v0[cs], a1 = v1[cs];
        if (NORM) { ssq = fmaf(a0, a0, fmaf(a1, a1, ssq)); a0 *= g[cs].x; a1 *= g[cs].y; }
        uint32_t hp, lp;
        split_bf16x2(a0, a1, hp, lp);                     // (hi, lo) packed pairs of columns (2 lane, 2 lane + 1)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            // columns 16 ks + 2 kq + {0, 1} live in lane 8 ks + kq, columns + {8, 9} in lane 8 ks + kq + 4
            const uint32_t h0 = __shfl_sync(0xffffffffu, hp, 8 * ks + kq), l0 = __shfl_sync(0xffffffffu, lp, 8 * ks + kq);
            const uint32_t h1 = __shfl_sync(0xffffffffu, hp, 8 * ks + kq + 4), l1 = __shfl_sync(0xffffffffu, lp, 8 * ks + kq + 4);
            xb[cs * 4 + ks][0] = act ? (is_lo ? l0 : h0) : 0u;
            xb[cs * 4 + ks][1] = act ? (is_lo ? l1 : h1) : 0u;
        }
    }
}

// =====================================================================================================================
template <int DEPTH, bool DIAG>
__global__ void __launch_bounds__(LL_THREADS, 1)
decode_ll_kernel(const __grid_constant__ LLArgs a, int max_rows) {
    constexpr int D = 128;
    extern __shared__ __align__(1024) unsigned char lsm[];
    // dynamic: [rings: LL_WARPS * DEPTH * 2 KB][acc: 2 x max_rows fixed-point pairs][xres: residual rows owned by this CTA]
    FixAcc* acc_s = reinterpret_cast<FixAcc*>(lsm + (size_t)LL_WARPS * DEPTH * LL_SEG_BYTES);
    float* xres_s = reinterpret_cast<float*>(acc_s + 2 * max_rows);
    __shared__ float ssq_s[4][LL_WARPS];                       // RMSNorm statistic: one slot per warp and phase parity (no float atomics:
                                                               // atomicAdd(float) on shared memory is a CAS spin loop)
    __shared__ __align__(16) float q_s[LL_MAX_NREP][D];
    __shared__ __align__(16) float knew_s[D];
    __shared__ __align__(16) float vnew_s[D];
    __shared__ __align__(16) float wp_s[LL_WARPS][2][LL_PART_STRIDE];
    __shared__ float wbest_v[LL_WARPS];
    __shared__ int wbest_i[LL_WARPS];
    __shared__ uint32_t tok_s;
    __shared__ WGeo wgeo_s[LL_WARPS][5];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const int n_phase = 4 * a.L + 1;
    const int NREP = a.nh / a.nkv;
    const uint32_t tag0 = a.tag_base;
    const uint32_t tstride = (uint32_t)(a.L + 2);
    Waiter wt{a.err, false, 0u};

    long long tprof = DIAG ? clock64() : 0ll;
    auto prof = [&](int slot) {
        if (DIAG && a.prof != nullptr && cta == 0 && tid == 0) {
            const long long now = clock64();
            a.prof[slot] += (unsigned long long)(now - tprof);
            tprof = now;
        }
    };

    // optional event trace of a few warps (tools/ll_trace.py): where the time of a phase goes, on a common clock
    int tr_s = -1, tr_p = 0, tr_n = 0;
    const int tr_cta = cta == 0 ? 0 : cta == 73 ? 1 : cta == 140 ? 2 : -1, tr_w = warp == 0 ? 0 : warp == 5 ? 1 : -1;
    const bool tr_on = DIAG && a.trace != nullptr && lane == 0 && tr_cta >= 0 && tr_w >= 0;
    unsigned long long* tr = a.trace + (size_t)((tr_cta < 0 ? 0 : tr_cta) * 2 + (tr_w < 0 ? 0 : tr_w)) * LL_TRACE_CAP * 2;
    auto trace = [&](int ev) {
        if (DIAG && tr_on && tr_s == a.trace_step && tr_n < LL_TRACE_CAP) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tr[2 * tr_n] = (unsigned long long)(ev | (tr_p << 8));
            tr[2 * tr_n + 1] = t;
            ++tr_n;
        }
    };

    // ---- prefetch cursor ---------------------------------------------------------------------------------------------
    const uint32_t ring_u32 = smem_u32(lsm + (size_t)warp * DEPTH * LL_SEG_BYTES);
    if (lane == 0) fill_geometry(a, wgeo_s[warp], warp);
    __syncwarp();
    const WGeo* wg = wgeo_s[warp];
    Cursor pf;
    pf.step = 0; pf.p = 0; pf.slot = 0;
    auto pf_enter = [&]() { cursor_enter(a, pf, wg); };
    // a slot = 16 rows x 128 bytes; 16-byte segment j of row r sits at r * 128 + ((j ^ (r & 7)) << 4) so that the eight rows an
    // ldmatrix phase reads fall into eight different bank groups.  Lane l copies segment (l & 7) of rows (l >> 3) + 4 i.
    const uint32_t cp_row = (uint32_t)(lane >> 3), cp_seg = (uint32_t)(lane & 7);
    auto issue_next = [&]() {         // one ring slot (never across a phase boundary) + exactly one commit
        if (pf.left != 0) {
            const uint32_t dst = ring_u32 + pf.slot;
            const unsigned char* src = pf.tile + (size_t)pf.cs * 128 + cp_seg * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t r = cp_row + 4u * i;
                if ((int)r < pf.rows_left) ll_cp16(dst + r * 128u + ((cp_seg ^ (r & 7u)) << 4), src + (size_t)r * pf.row_bytes);
            }
            pf.slot = (pf.slot + LL_SEG_BYTES == DEPTH * LL_SEG_BYTES) ? 0u : pf.slot + LL_SEG_BYTES;
            if (++pf.cs == pf.ncs) { pf.cs = 0; pf.tile += (size_t)16 * pf.row_bytes; pf.rows_left -= 16; }
            if (--pf.left == 0) { ++pf.p; pf_enter(); }
        }
        ll_commit();
    };
    pf_enter();
#pragma unroll 1
    for (int i = 0; i < DEPTH; ++i) issue_next();
    uint32_t cslot = 0;               // byte offset of the next ring slot to consume

    // ---- sequence state, residual slice, first input ----------------------------------------------------------------------
    const SeqState st0 = a.state[0];
    const int* bt = a.block_table + (size_t)st0.slot * a.max_pages;
    const int res_rpc = (a.H + grid - 1) / grid;
    const int res_r0 = cta * res_rpc;
    const int res_n = max(0, min(a.H, res_r0 + res_rpc) - res_r0);
    for (int i = tid; i < 2 * max_rows; i += LL_THREADS) acc_s[i] = FixAcc{0, 0};
    if (tid < 4 * LL_WARPS) (&ssq_s[0][0])[tid] = 0.f;
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
            wt.where = ((uint32_t)s << 12) | (uint32_t)p;
            if (DIAG) { tr_s = s; tr_p = p; }

            // ============================ attention of layer l (between the QKV and the O projections) ============================
            if (kind == 1) {
                const LLLayer& ly = a.layers[l];
                const int nsplit = max(1, min(min(grid / a.nkv, LL_MAX_SPLIT), (T + 31) / 32));
                trace(7);
                if (cta < a.nkv * nsplit) {
                    const int kvh = cta / nsplit, split = cta % nsplit;
                    const int chunk = (T + nsplit - 1) / nsplit;
                    const int t0 = split * chunk;
                    const int t_end = min(min(T, t0 + chunk), T - 1);   // cached tokens of this split (position T-1 comes from this step's qkv)
                    const bool owner = (split == (T - 1) / chunk);
                    const bool split_kv = a.kv_lo_off != 0;
                    const int half = lane >> 4, hl = lane & 15;
                    constexpr int TPP = 2 * LL_WARPS;                   // tokens per pass: two per warp (16 lanes each)
                    const int npass = t_end > t0 ? (t_end - t0 + TPP - 1) / TPP : 0;
                    uint4 kh = make_uint4(0, 0, 0, 0), kl = kh, vh = kh, vl = kh;
                    auto load_pass = [&](int ps, uint4& rkh, uint4& rkl, uint4& rvh, uint4& rvl) {
                        const int t = t0 + ps * TPP + warp * 2 + half;
                        rkh = make_uint4(0, 0, 0, 0); rkl = rkh; rvh = rkh; rvl = rkh;
                        if (t < t_end) {
                            const int page = __ldg(bt + t / KV_PAGE);
                            const size_t off = (((size_t)page * a.nkv + kvh) * KV_PAGE + (t % KV_PAGE)) * D + hl * 8;
                            rkh = ldg_stream(ly.k_pool + off);
                            rvh = ldg_stream(ly.v_pool + off);
                            if (split_kv) { rkl = ldg_stream(ly.k_pool + a.kv_lo_off + off); rvl = ldg_stream(ly.v_pool + a.kv_lo_off + off); }
                        }
                    };
                    uint4 nkh = make_uint4(0, 0, 0, 0), nkl = nkh, nvh = nkh, nvl = nkh;
                    if (npass > 0) load_pass(0, kh, kl, vh, vl);        // two passes in flight while the query is still being produced
                    if (npass > 1) load_pass(1, nkh, nkl, nvh, nvl);
                    // ---- q (NREP heads), and on the owning split the new k / v: poll, RMSNorm, rotate ----
                    for (int vec = warp; vec < NREP + 2; vec += LL_WARPS) {
                        if (vec >= NREP && !owner) continue;
                        const bool is_k = vec == NREP, is_v = vec == NREP + 1;
                        const unsigned long long* src = a.qkv + (is_v ? a.q_dim + a.nkv * D + kvh * D : is_k ? a.q_dim + kvh * D : (kvh * NREP + vec) * D);
                        const float* nw = is_k ? ly.kn : ly.qn;
                        float nwv[4], cs_c[4], cs_s[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            nwv[j] = is_v ? 1.f : nw[lane + 32 * j];
                            const int i = lane + 32 * (j & 1);           // rotary pair (j, j + 2): column i = lane + 32 * (j mod 2)
                            const int ax = a.axis_of[i];
                            const int pp = ax == 0 ? pos3[0] : ax == 1 ? pos3[1] : pos3[2];   // no dynamically indexed local array: there is no L1
                            cs_c[j] = a.cos_tab[(size_t)pp * (D / 2) + i];
                            cs_s[j] = a.sin_tab[(size_t)pp * (D / 2) + i];
                        }
                        float e[4];
                        uint32_t it = 0;
    long long wt0 = 0;
                        for (;;) {
                            bool ok = true;
                            uint32_t seen = tagl;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const unsigned long long u = ld_pair1(src + lane + 32 * j);
                                if (pair_tag(u) != tagl) { ok = false; seen = pair_tag(u); }
                                e[j] = pair_val(u);
                            }
                            if (__all_sync(0xffffffffu, ok)) break;
                            if (!wt.again(it, wt0, 2u, tagl, seen)) break;
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
                            float* dst = is_k ? knew_s : q_s[vec];
#pragma unroll
                            for (int j = 0; j < 4; ++j) dst[lane + 32 * j] = is_k ? (split_kv ? round_bf16_split(r[j]) : round_bf16(r[j])) : r[j];
                        }
                    }
                    __syncthreads();
                    trace(8);
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
                        if (hg0 != 0) {
                            if (npass > 0) load_pass(0, kh, kl, vh, vl);
                            if (npass > 1) load_pass(1, nkh, nkl, nvh, nvl);
                        }
#pragma unroll 1
                        for (int ps = 0; ps < npass; ++ps) {
                            uint4 fkh = make_uint4(0, 0, 0, 0), fkl = fkh, fvh = fkh, fvl = fkh;      // pass ps + 2
                            if (ps + 2 < npass) load_pass(ps + 2, fkh, fkl, fvh, fvl);
                            const bool valid = (t0 + ps * TPP + warp * 2 + half) < t_end;
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
                            nkh = fkh; nkl = fkl; nvh = fvh; nvl = fvl;
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
                        // no fence: the page row is first read a whole decode step later, through L2 (the loads bypass L1), and a fence
                        // here would sit on the critical path of every layer (this CTA's O-projection rows wait behind it)
                        __syncthreads();  // knew_s / vnew_s are rewritten by the next layer's item
                    }
                }
                prof(4);
                trace(9);
                // ---- split merge of head `cta`: one split per lane, 8 output columns per warp ----
                if (cta < a.nh) {
                    const int h = cta, kvh = h / NREP, hh = h % NREP;
                    const bool have = lane < nsplit;
                    const unsigned long long* src = a.part + ((size_t)(kvh * nsplit + (have ? lane : 0)) * NREP + hh) * LL_PART_STRIDE;
                    constexpr int CPW = D / LL_WARPS;               // output columns per warp
                    float ov[CPW], ms = -INFINITY, lsum = 0.f;
                    uint32_t it = 0;
    long long wt0 = 0;
                    for (;;) {
                        bool ok = true;
                        uint32_t seen = tagl;
                        if (have) {
                            unsigned long long u0, u1;
#pragma unroll
                            for (int j = 0; j < CPW / 2; ++j) {
                                ld_pair2(src + warp * CPW + 2 * j, u0, u1);
                                if (pair_tag(u0) != tagl) { ok = false; seen = pair_tag(u0); }
                                if (pair_tag(u1) != tagl) { ok = false; seen = pair_tag(u1); }
                                ov[2 * j] = pair_val(u0); ov[2 * j + 1] = pair_val(u1);
                            }
                            ld_pair2(src + 128, u0, u1);
                            if (pair_tag(u0) != tagl) { ok = false; seen = pair_tag(u0); }
                            if (pair_tag(u1) != tagl) { ok = false; seen = pair_tag(u1); }
                            ms = pair_val(u0); lsum = pair_val(u1);
                        }
                        if (__all_sync(0xffffffffu, ok)) break;
                        if (!wt.again(it, wt0, 3u, tagl, seen)) break;
                    }
                    if (!have) {
                        ms = -INFINITY; lsum = 0.f;
#pragma unroll
                        for (int j = 0; j < CPW; ++j) ov[j] = 0.f;
                    }
                    const float M = warp_max(ms);
                    const float c = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                    const float Ls = warp_sum(lsum * c);
                    float outv = 0.f;
#pragma unroll
                    for (int j = 0; j < CPW; ++j) {
                        const float t = warp_sum(ov[j] * c);
                        if (lane == j) outv = t;
                    }
                    if (lane < CPW) st_pair(a.att + (size_t)h * D + warp * CPW + lane, outv / Ls, tagl);
                }
                prof(5);
                trace(10);
            }

            // ============================ weight phase: y = W . x over this CTA's row block ============================
            if (a.l2_ahead > 0 && tid == 0) {          // this CTA's slab of a phase `l2_ahead` phases from now -> L2 (never the lm_head: 4x the L2)
                int pp = p + a.l2_ahead, ss = s;
                if (pp >= n_phase) { pp -= n_phase; ++ss; }
                if (ss < a.n_steps && pp < 4 * a.L) {
                    const WGeo& wf = wgeo_s[0][phase_kind(pp, a.L)];
                    if (wf.nrows > 0) ll_prefetch_l2(phase_weights(a, pp) + (size_t)wf.r0 * wf.K * 2, (size_t)wf.nrows * wf.K * 2);
                }
            }
            const WGeo w = wg[kind];
            const LLLayer& ly = a.layers[l];
            const float* nw = (kind == 0) ? ly.ln1 : (kind == 2) ? ly.ln2 : (kind == 4) ? a.final_norm : nullptr;
            const unsigned long long* xin = (kind == 0 || kind == 4) ? a.xa : (kind == 1) ? a.att : (kind == 2) ? a.xb : a.act;
            FixAcc* acc = acc_s + (gphase & 1u) * max_rows;
            float* ssq_p = ssq_s[gphase & 3u];
            const unsigned int aw = (unsigned int)ll_active_warps(w.K);

            auto run = [&](auto ncs_tag, auto norm_tag) {
                constexpr int NCS = decltype(ncs_tag)::value;
                constexpr bool NORM = decltype(norm_tag)::value;
                // A CTA reads an exchange buffer only in phases where it owns rows: every such read is waited for, transitively, by
                // whoever overwrites the buffer next (each writer needs the outputs of all row owners of the phase before).  A read
                // nobody depends on could be overtaken by the next layer's values and would then never see its tag.
                if (w.nslots == 0) return;
                trace(1);
                uint32_t xb[NCS * 4][2];
                float ssq = 0.f;
                load_xb<NCS, NORM>(xin, tagl, w.col0, lane, nw, xb, ssq, wt);
                if (NORM) {
                    ssq = warp_sum(ssq);
                    if (lane == 0) ssq_p[warp] = ssq;
                }
                prof(1);
                trace(2);
                // ldmatrix addresses of this lane inside a slot: row (lane & 7) + 8 * ((lane >> 3) & 1), 16-byte segment 2 ks + (lane >> 4)
                const uint32_t lrow = (uint32_t)(lane & 7) + (((uint32_t)lane >> 3) & 1u) * 8u;
                const uint32_t lbase = ring_u32 + lrow * 128u;
                const uint32_t lsw = lrow & 7u, lhalf = (uint32_t)lane >> 4;
#pragma unroll 1
                for (int rt = 0; rt < w.ntiles; ++rt) {
                    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int cs = 0; cs < NCS; ++cs) {
                        ll_wait<DEPTH - 1>();
                        __syncwarp();
                        trace(3);
                        const uint32_t sb = lbase + cslot;
                        uint32_t af[4][4];
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            ll_ldmatrix_x4(af[ks][0], af[ks][1], af[ks][2], af[ks][3], sb + (((2u * ks + lhalf) ^ lsw) << 4));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (ks & 1) ll_mma_16816(d1, af[ks][0], af[ks][1], af[ks][2], af[ks][3], xb[cs * 4 + ks][0], xb[cs * 4 + ks][1]);
                            else ll_mma_16816(d0, af[ks][0], af[ks][1], af[ks][2], af[ks][3], xb[cs * 4 + ks][0], xb[cs * 4 + ks][1]);
                        }
                        __syncwarp();                              // every lane has its fragments: the slot may be refilled
                        issue_next();
                        cslot = (cslot + LL_SEG_BYTES == DEPTH * LL_SEG_BYTES) ? 0u : cslot + LL_SEG_BYTES;
                    }
                    // D: lane l holds rows l / 4 and l / 4 + 8, columns 2 (l % 4) + {0, 1}; columns 0 and 1 are the hi and lo products.
                    // The warps own different columns of the same 16 rows: exact fixed-point accumulation in shared memory (fix_add).
                    if ((lane & 3) == 0) {
                        const int r = rt * 16 + (lane >> 2);
                        if (r < w.nrows) fix_add(&acc[r], (d0[0] + d1[0]) + (d0[1] + d1[1]));
                        if (r + 8 < w.nrows) fix_add(&acc[r + 8], (d0[2] + d1[2]) + (d0[3] + d1[3]));
                    }
                }
            };
            // (columns per warp, with / without RMSNorm): the norm phases read K = hidden_size, the widest slices are MLP down-projections
            const bool norm = nw != nullptr;
            using TT = std::true_type; using FF = std::false_type;
            switch (w.ncs) {
                case 1: if (norm) run(std::integral_constant<int, 1>{}, TT{}); else run(std::integral_constant<int, 1>{}, FF{}); break;
                case 2: if (norm) run(std::integral_constant<int, 2>{}, TT{}); else run(std::integral_constant<int, 2>{}, FF{}); break;
                case 4: if (norm) run(std::integral_constant<int, 4>{}, TT{}); else run(std::integral_constant<int, 4>{}, FF{}); break;
                case 8: if (norm) run(std::integral_constant<int, 8>{}, TT{}); else run(std::integral_constant<int, 8>{}, FF{}); break;
                default: run(std::integral_constant<int, 12>{}, FF{}); break;
            }
            prof(2);
            trace(4);
            __syncthreads();
            prof(3);
            trace(5);

            // ============================ epilogue ============================
            float rstd = 1.f;
            if (nw != nullptr) {
                float tot = 0.f;
                for (unsigned int ww = 0; ww < aw; ++ww) tot += ssq_p[ww];
                rstd = rsqrtf(tot / (float)w.K + a.eps);
            }
            if (kind == 0) {
                for (int i = tid; i < w.nrows; i += LL_THREADS) {
                    st_pair(a.qkv + w.r0 + i, fix_take(&acc[i]) * rstd, tagl);
                }
            } else if (kind == 1 || kind == 3) {
                unsigned long long* dst = (kind == 1) ? a.xb : a.xa;
                const uint32_t tg = (kind == 1) ? tagl : tagl + 1u;   // down-proj output = input of the next layer (or of the lm_head)
                for (int i = tid; i < w.nrows; i += LL_THREADS) {
                    const float v = xres_s[i] + fix_take(&acc[i]);
                    xres_s[i] = v;
                    st_pair(dst + w.r0 + i, v, tg);
                }
            } else if (kind == 2) {
                for (int u = tid; u < w.nrows / 2; u += LL_THREADS) {
                    const float g = fix_take(&acc[2 * u]) * rstd, up = fix_take(&acc[2 * u + 1]) * rstd;
                    st_pair(a.act + w.r0 / 2 + u, silu_f(g) * up, tagl);
                }
            } else {
                // logits + (value, lowest index) argmax: per thread rows ascend, so the first maximum wins
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int i = tid; i < w.nrows; i += LL_THREADS) {
                    const float v = fix_take(&acc[i]) * rstd;
                    a.logits[w.r0 + i] = v;
                    if (v > bv) { bv = v; bi = w.r0 + i; }
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
                    unsigned long long* am = a.amax + (size_t)(s & 1) * 2 * grid;     // two copies: a CTA nobody waits for may still be reading the last step's
                    st_pair(am + 2 * cta, bv, tg);
                    st_pair(am + 2 * cta + 1, __int_as_float(bi), tg);
                }
            }
            prof(6);
            trace(6);
        }

        // ============================ token: every CTA reduces the per-CTA maxima itself (no broadcast hop) ============================
        wt.where = ((uint32_t)s << 12) | (uint32_t)n_phase;
        if (warp == 0) {
            const uint32_t tg = tstep + (uint32_t)a.L + 2u;
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int c0 = 0; c0 < grid; c0 += 32) {
                const int c = c0 + lane;
                float v = -INFINITY; int ix = 0x7fffffff;
                uint32_t it = 0;
    long long wt0 = 0;
                for (;;) {
                    bool ok = true;
                    uint32_t seen = tg;
                    if (c < grid) {
                        unsigned long long u0, u1;
                        ld_pair2(a.amax + (size_t)(s & 1) * 2 * grid + 2 * c, u0, u1);
                        ok = pair_tag(u0) == tg && pair_tag(u1) == tg;
                        if (!ok) seen = pair_tag(u0) != tg ? pair_tag(u0) : pair_tag(u1);
                        v = pair_val(u0); ix = __float_as_int(pair_val(u1));
                    }
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (!wt.again(it, wt0, 4u, tg, seen)) break;
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
        if (DIAG) tr_p = n_phase;
        trace(11);
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
    return (size_t)LL_WARPS * depth * LL_SEG_BYTES + (size_t)2 * mr * 8 + (size_t)((a.H + grid - 1) / grid) * 4 + 16;
}

bool decode_ll_supported(int D, int rot_half, int nh, int nkv, int H, int I, int q_dim, int qkv_dim, int V, int num_sms) {
    if (D != 128 || rot_half != 64 || nkv <= 0 || nh % nkv || nh / nkv > LL_MAX_NREP || nh > num_sms || nkv > num_sms) return false;
    auto cols_ok = [](int K, bool norm) { // every warp owns 64 * ncs columns, ncs one of the instantiated widths
        if (K <= 0 || (K % 64) != 0) return false;
        const int ncs = K / (64 * ll_active_warps(K));
        return ncs == 1 || ncs == 2 || ncs == 4 || ncs == 8 || (!norm && ncs == 12);
    };
    if (!cols_ok(H, true) || !cols_ok(I, false) || !cols_ok(q_dim, false)) return false;
    if ((2 * I) % 2 || qkv_dim <= 0 || V <= 0) return false;
    LLArgs a = {};
    a.H = H; a.I = I; a.V = V; a.qkv_dim = qkv_dim;
    return ll_smem(a, num_sms, 64 / LL_WARPS) + 20 * 1024 <= 227 * 1024;     // + the static shared memory (~16 KB)
}

size_t decode_ll_part_pairs(int num_sms, int nh, int nkv) { return (size_t)num_sms * (size_t)(nh / nkv) * LL_PART_STRIDE; }

template <int DEPTH, bool DIAG>
static int ll_launch_t(cudaStream_t st, const LLArgs& a, int grid, size_t smem, int max_rows) {
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(decode_ll_kernel<DEPTH, DIAG>, smem, seen)) return e;
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
    return (int)cudaLaunchKernelEx(&cfg, decode_ll_kernel<DEPTH, DIAG>, a, max_rows);
}

int decode_ll_launch(cudaStream_t st, const LLArgs& a, int num_sms) {
    const int mr = ll_max_rows(a, num_sms);
    static const int forced = [] { const char* e = getenv("CRANE_B200_LL_DEPTH"); return e ? atoi(e) : 0; }();
    cudaFuncAttributes fa;
    if (const cudaError_t e = cudaFuncGetAttributes(&fa, decode_ll_kernel<64 / LL_WARPS, true>)) return (int)e;
    const size_t budget = 227 * 1024 - fa.sharedSizeBytes;      // dynamic part next to the kernel's static shared memory
    for (int depth : {96 / LL_WARPS, 80 / LL_WARPS, 64 / LL_WARPS}) {
        if (forced && depth != forced) continue;
        const size_t smem = ll_smem(a, num_sms, depth);
        if (smem > budget) continue;
        const bool diag = a.prof != nullptr || a.trace != nullptr;       // the instrumented build of the kernel only when asked for
        if (depth == 96 / LL_WARPS) return diag ? ll_launch_t<96 / LL_WARPS, true>(st, a, num_sms, smem, mr) : ll_launch_t<96 / LL_WARPS, false>(st, a, num_sms, smem, mr);
        if (depth == 80 / LL_WARPS) return diag ? ll_launch_t<80 / LL_WARPS, true>(st, a, num_sms, smem, mr) : ll_launch_t<80 / LL_WARPS, false>(st, a, num_sms, smem, mr);
        return diag ? ll_launch_t<64 / LL_WARPS, true>(st, a, num_sms, smem, mr) : ll_launch_t<64 / LL_WARPS, false>(st, a, num_sms, smem, mr);
    }
    return -1000;
}

}  // namespace cb
