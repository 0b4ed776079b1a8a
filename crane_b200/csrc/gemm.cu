// tcgen05 + TMA GEMM for the prefill / ViT path.  See gemm.cuh for the design notes.
#include "gemm.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace cb {

// CRANE_B200_GEMM_BN=64|128|256 pins the tile width (sweeps).  (Round 1 also had 2x2 TMA-multicast clusters; measured on B200 they
// did not pay -- the L2 already merges concurrent requests for a line from a few CTAs -- and they are gone.)
static const int g_gemm_bn = [] { const char* e = getenv("CRANE_B200_GEMM_BN"); return e ? atoi(e) : 0; }();

// =====================================================================================
// PTX wrappers (sm_100a)
// =====================================================================================
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on `bar` when every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in [0,14), LBO(=1, ignored for swizzled K-major) in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// =====================================================================================
// Fused epilogue of one warp: a 32-row x 32-column accumulator block (lane = row, as tcgen05.ld delivers it).
// The arithmetic (bias / activation / SiLU*up / bf16 split) runs on the lane's own row; the results are then transposed
// through a 32 x 144 B shared-memory patch so that every global store (and the residual read) covers whole 128-byte rows --
// row-per-lane stores cost 32 LSU wavefronts per instruction and made the epilogue as long as the main loop.
// =====================================================================================
// Row-per-lane stores of 32 finished columns (bias already added): the bf16 modes of the tcgen05 epilogue and every mode of the
// SIMT debugging kernel.
template <int MODE>
__device__ __forceinline__ void epi_store32(const GemmEpi& ep, int m, int n, float* v, bool ok) {
    if (!ok) return;
    if constexpr (MODE == EPI_STORE_F32 || MODE == EPI_RESID_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = (MODE == EPI_RESID_F32 ? o[j] : 0.f) + v[j];
    } else if constexpr (MODE == EPI_SILU_MUL_BF16) {
        const size_t off = (size_t)m * ep.ldo + (n >> 1);
        uint32_t p[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16x2(silu_f(v[4 * j]) * v[4 * j + 1], silu_f(v[4 * j + 2]) * v[4 * j + 3], p[j], q[j]);
        bf16* o = reinterpret_cast<bf16*>(ep.out) + off;
        *reinterpret_cast<uint4*>(o) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4*>(o + 8) = make_uint4(p[4], p[5], p[6], p[7]);
        if (ep.out_lo != nullptr) {
            bf16* ol = reinterpret_cast<bf16*>(ep.out_lo) + off;
            *reinterpret_cast<uint4*>(ol) = make_uint4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<uint4*>(ol + 8) = make_uint4(q[4], q[5], q[6], q[7]);
        }
    } else {
        const size_t off = (size_t)m * ep.ldo + n;
        bf16* o = reinterpret_cast<bf16*>(ep.out) + off;
        bf16* ol = ep.out_lo != nullptr ? reinterpret_cast<bf16*>(ep.out_lo) + off : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            uint32_t p[4], q[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float a = v[j + 2 * t], b = v[j + 2 * t + 1];
                if constexpr (MODE == EPI_GELU_ERF_BF16) { a = gelu_erf_f(a); b = gelu_erf_f(b); }
                else if constexpr (MODE == EPI_GELU_TANH_BF16) { a = gelu_tanh_f(a); b = gelu_tanh_f(b); }
                split_bf16x2(a, b, p[t], q[t]);
            }
            *reinterpret_cast<uint4*>(o + j) = make_uint4(p[0], p[1], p[2], p[3]);
            if (ol != nullptr) *reinterpret_cast<uint4*>(ol + j) = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }
}

constexpr int EPI_PITCH = 144;                     // bytes per staged row (128 B payload + 16 B skew)
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH;

template <int MODE>
__device__ __forceinline__ void epi_block(const GemmEpi& ep, uint8_t* stg, int m_base, int M, int n, float* v, int lane) {
    if (ep.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(ep.bias + n + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
    }
    if constexpr (MODE == EPI_STORE_F32 || MODE == EPI_RESID_F32) {
        uint8_t* mine = stg + lane * EPI_PITCH;
        // the residual rows first (independent loads, all in flight together), then the transpose
        float4 y[8];
        if constexpr (MODE == EPI_RESID_F32) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int m = m_base + it * 4 + (lane >> 3);
                y[it] = (m < M) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.out) + (size_t)m * ep.ldo + n + (lane & 7) * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(mine + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {           // 4 rows x 128 B per instruction
            const int r = it * 4 + (lane >> 3), c4 = lane & 7;
            const int m = m_base + r;
            if (m < M) {
                float4 x = *reinterpret_cast<const float4*>(stg + r * EPI_PITCH + c4 * 16);
                if constexpr (MODE == EPI_RESID_F32) { x.x += y[it].x; x.y += y[it].y; x.z += y[it].z; x.w += y[it].w; }
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n + c4 * 4) = x;
            }
        }
        __syncwarp();                              // the patch is rewritten by the next block
    } else {
        // bf16 outputs: a lane's 32 (16) values are 64 (32) contiguous bytes = whole sectors already -- stored straight from registers
        epi_store32<MODE>(ep, m_base + lane, n, v, m_base + lane < M);
    }
}

// =====================================================================================
// tcgen05 kernel: persistent, stream-K over (tile, k-step) iterations, 256 threads per CTA, one CTA per SM
//   warp 0 : TMA producer (one elected lane)       warp 1 : MMA issuer (one elected lane)
//   warp 2 : TMEM allocator / deallocator           warps 4-7 : epilogue (TMEM lane quadrant = warp % 4)
//
// Why stream-K.  At the prefill shapes of this engine (M = 454 text rows, 784 patches) a 128 x BN grid of output tiles is either too
// few CTAs (BN = 256: 32-64 tiles for N = 2048-4096) or, with narrow tiles, bound by L2 -> SM operand traffic: measured (tools/
// gemm_probe.py) a k-step costs ~325 ns whether BN is 64, 128 or 256, i.e. the ~12 TB/s the L2 can hand out, while the tensor pipe
// needs 32 / 64 / 128 cycles per MMA.  Only BN = 256 tiles are MMA-bound, so the work is cut the other way: the iteration space
// tiles x k-steps is dealt out evenly, in order, to min(#SMs, ...) CTAs; a CTA whose range starts or ends inside a tile owns a
// PARTIAL sum of that tile.  Partials go to a workspace; the contributor that arrives last adds them IN CONTRIBUTOR ORDER (so the
// result does not depend on who was last: bitwise reproducible) and runs the fused epilogue.  Nobody ever waits for another CTA.
// The accumulator is double-buffered in TMEM (2 x BN columns): the MMAs of a CTA's next segment run under the epilogue of the last.
// =====================================================================================
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGE_BUDGET = 192 * 1024;
constexpr int GEMM_MAX_CTAS = 160;                 // workspace slots: two partial tiles per CTA

struct GemmSched {
    int tiles_m, tiles_n, KB;      // output tiles and k-steps per tile; tile t = n_tile * tiles_m + m_tile (the CTAs working at the
                                   // same time share their weight rows: the L2 merges those requests)
    int G, unit, base, rem;        // G CTAs; CTA c owns iterations [unit * (c * base + min(c, rem)), ...): base or base + 1 units of
                                   // `unit` k-steps -- unit = KB deals out whole tiles (nothing is split), unit = 1 is stream-K
    float* ws;                     // [2 * G][128 * BN] partial tiles
    int* counters;                 // [tiles] arrivals per tile, zero between launches (the finisher resets its tile)
};
__host__ __device__ inline int sched_start(const GemmSched& s, int c) { return s.unit * (c * s.base + (c < s.rem ? c : s.rem)); }
__host__ __device__ inline int sched_owner(const GemmSched& s, int it) {       // CTA whose range holds iteration `it`
    const int u = it / s.unit, big = s.rem * (s.base + 1);
    return u < big ? u / (s.base + 1) : s.rem + (u - big) / s.base;
}

template <int BN, bool SPLIT>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES * (SPLIT ? 2 : 1) + B_BYTES;     // [A hi | A lo (SPLIT) | B]
    static constexpr int STAGES = (GEMM_STAGE_BUDGET / STAGE_BYTES) < 8 ? (GEMM_STAGE_BUDGET / STAGE_BYTES) : 8;
    static constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;                        // transpose patches of the four epilogue warps
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;
};

// SPLIT: the activation operand comes as two bf16 planes A = hi + lo (lo = bf16(x - hi), ~16 mantissa bits together); the
// weight tile is fetched once and multiplied by both, accumulating into the same TMEM columns: D = hi.W^T + lo.W^T.
template <int BN, int MODE, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmB,
               GemmEpi ep, int M, int N, int K, GemmSched sc) {
    using Cfg = GemmCfg<BN, SPLIT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2] accumulator b complete
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] accumulator b drained by the four epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int KB = sc.KB;
    const int it_begin = sched_start(sc, cta), it_end = sched_start(sc, cta + 1);
    // optional timeline of this CTA (tools/gemm_probe.py): 8 globaltimer stamps per CTA
    unsigned long long* prof = ep.prof ? ep.prof + 8 * (size_t)cta : nullptr;
    auto stamp = [&](int i) {
        if (prof) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); prof[i] = t; }
    };
    if (threadIdx.x == 0) stamp(0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (SPLIT) tma_prefetch_desc(&tmAlo);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(1);                 // setup done

    pdl_launch_dependents();                        // the next kernel may begin its own setup / weight prefetch now
    auto tile_origin = [&](int it, int& m0, int& n0) {
        const int t = it / KB;
        m0 = (t % sc.tiles_m) * GEMM_BM;
        n0 = (t / sc.tiles_m) * BN;
    };
    if (warp == 0) {
        if (lane == 0) {
            // weights do not depend on the predecessor kernel: their first tiles go out before the dependency wait
            const int npre = ep.w_dynamic ? 0 : min(it_end - it_begin, STAGES);
            for (int i = 0; i < npre; ++i) {
                int m0, n0;
                tile_origin(it_begin + i, m0, n0);
                uint8_t* st = smem + i * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
                tma_load_2d(st + Cfg::A_BYTES * (SPLIT ? 2 : 1), &tmB, &full_bar[i], ((it_begin + i) % KB) * GEMM_BK, n0);
            }
            pdl_wait();
            int s = 0; uint32_t ph = 0;
            for (int it = it_begin; it < it_end; ++it) {
                int m0, n0;
                tile_origin(it, m0, n0);
                const int kc = (it % KB) * GEMM_BK;
                uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
                const bool pre = it - it_begin < npre;
                if (!pre) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    tma_load_2d(a_dst + Cfg::A_BYTES * (SPLIT ? 2 : 1), &tmB, &full_bar[s], kc, n0);
                }
                tma_load_2d(a_dst, &tmA, &full_bar[s], kc, m0);
                if (SPLIT) tma_load_2d(a_dst + Cfg::A_BYTES, &tmAlo, &full_bar[s], kc, m0);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
            int s = 0; uint32_t ph = 0;
            int seg = 0;
            for (int it = it_begin; it < it_end; ++seg) {
                const int seg_end = min(it_end, (it / KB + 1) * KB);
                const int b = seg & 1;
                mbar_wait(&tmem_empty_bar[b], ((seg >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator (first two: free)
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(b * BN);
                for (int i = it; i < seg_end; ++i) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    if (i == it_begin) stamp(2);        // first operands landed
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc_sw128(a_addr);
                    const uint64_t aldesc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
                    const uint64_t bdesc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES * (SPLIT ? 2 : 1));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // +32 B per K=16 step inside the 128 B swizzle atom -> +2 in the >>4-encoded start address
                        umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((i != it) | (k != 0)));
                        if (SPLIT) umma_bf16(acc, aldesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);         // frees the smem stage once these MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full_bar[b]);         // this segment's accumulator is complete
                it = seg_end;
            }
            stamp(3);                                   // last MMA issued
        }
    } else if (warp >= 4) {
        pdl_wait();                                 // the epilogue reads / overwrites activations of the predecessor too
        const int q = warp & 3;
        uint8_t* stg = epi_smem + q * EPI_WARP_BYTES;
        constexpr int CHUNKS = BN / 32;
        constexpr size_t TILE_F = (size_t)GEMM_BM * BN;
        int seg = 0;
        for (int it = it_begin; it < it_end; ++seg) {
            const int seg_end = min(it_end, (it / KB + 1) * KB);
            const int t = it / KB, kb0 = it % KB, kb1 = kb0 + (seg_end - it);
            const int m0 = (t % sc.tiles_m) * GEMM_BM, n0 = (t / sc.tiles_m) * BN;
            const int b = seg & 1;
            const uint32_t acc = tmem_base + (uint32_t)(b * BN) + ((uint32_t)(q * 32) << 16);
            const bool sole = kb0 == 0 && kb1 == KB;
            mbar_wait(&tmem_full_bar[b], (seg >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 128 && seg == 0) stamp(4);   // first accumulator ready
            if (sole) {
#pragma unroll 1
                for (int c = 0; c < CHUNKS; ++c) {
                    uint32_t r[32];
                    tmem_ld32(acc + (uint32_t)(c * 32), r);
                    const int n = n0 + c * 32;
                    if (n < N) {                            // warp-uniform
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        epi_block<MODE>(ep, stg, m0 + q * 32, M, n, v, lane);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[b]);
            } else {
                // partial sum of tile t: park it.  A 32 x 32 block of (row = lane, column j) is stored as 8 rows of 32 float4: the float4
                // of lane l in row k holds columns 4k .. 4k+3 of row l, so every store / reload instruction moves 512 contiguous bytes
                float* mine = sc.ws + (size_t)(2 * cta + (it == it_begin ? 0 : 1)) * TILE_F + (size_t)q * CHUNKS * 1024;
#pragma unroll 1
                for (int c = 0; c < CHUNKS; ++c) {
                    uint32_t r[32];
                    tmem_ld32(acc + (uint32_t)(c * 32), r);
                    float4* dst = reinterpret_cast<float4*>(mine + c * 1024) + lane;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        dst[k * 32] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[b]);
                // arrive at the tile; the last of its contributors finishes it
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int c_first = sched_owner(sc, t * KB), c_last = sched_owner(sc, t * KB + KB - 1);
                if (threadIdx.x == 128) *last_flag = atomicAdd(&sc.counters[t], 1) == c_last - c_first;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (*last_flag) {
                    __threadfence();
#pragma unroll 1
                    for (int c = 0; c < CHUNKS; ++c) {
                        const int n = n0 + c * 32;
                        if (n >= N) break;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                        auto part = [&](int cc) {
                            return reinterpret_cast<const float4*>(sc.ws + (size_t)(2 * cc + (sched_start(sc, cc) < t * KB ? 1 : 0)) * TILE_F +
                                                                   (size_t)(q * CHUNKS + c) * 1024) + lane;
                        };
                        // contributor order, whoever came last; two contributors' loads in flight together (the additions stay in order)
                        for (int cc = c_first; cc <= c_last; cc += 2) {
                            const bool two = cc + 1 <= c_last;
                            const float4* p0 = part(cc);
                            const float4* p1 = part(two ? cc + 1 : cc);
                            float4 x0[8], x1[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) x0[k] = __ldcg(p0 + k * 32);
#pragma unroll
                            for (int k = 0; k < 8; ++k) x1[k] = two ? __ldcg(p1 + k * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int k = 0; k < 8; ++k) { v[4 * k] += x0[k].x; v[4 * k + 1] += x0[k].y; v[4 * k + 2] += x0[k].z; v[4 * k + 3] += x0[k].w; }
                            if (two) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) { v[4 * k] += x1[k].x; v[4 * k + 1] += x1[k].y; v[4 * k + 2] += x1[k].z; v[4 * k + 3] += x1[k].w; }
                            }
                        }
                        epi_block<MODE>(ep, stg, m0 + q * 32, M, n, v, lane);
                    }
                    if (threadIdx.x == 128) sc.counters[t] = 0;       // clean for the next launch on this stream
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");           // last_flag is rewritten by the next segment
            }
            it = seg_end;
        }
        if (threadIdx.x == 128) stamp(5);               // this CTA's rows stored
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(6);                 // whole CTA done
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); if (lane == 0) stamp(7); }
}

// =====================================================================================
// SIMT debugging kernel (bring-up A/B for the tcgen05 path; never on the product path)
// =====================================================================================
template <int MODE>
__global__ void __launch_bounds__(128)
gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ A_lo, int lda, const bf16* __restrict__ W, GemmEpi ep, int M, int N, int K) {
    __shared__ float ws[32][65];
    const int m = blockIdx.y * 128 + threadIdx.x;
    const int n0 = blockIdx.x * 32;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += 128) {
            const int r = i >> 6, c = i & 63;
            ws[r][c] = (n0 + r < N) ? __bfloat162float(W[(size_t)(n0 + r) * K + k0 + c]) : 0.f;
        }
        __syncthreads();
        if (m < M) {
            for (int c = 0; c < 64; ++c) {
                float a = __bfloat162float(A[(size_t)m * lda + k0 + c]);
                if (A_lo != nullptr) a += __bfloat162float(A_lo[(size_t)m * lda + k0 + c]);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, ws[j][c], acc[j]);
            }
        }
    }
    if (m < M && n0 < N) {
        if (ep.bias != nullptr)
            for (int j = 0; j < 32; ++j) acc[j] += ep.bias[n0 + j];
        epi_store32<MODE>(ep, m, n0, acc, true);
    }
}

// =====================================================================================
// Host side
// =====================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// 2-D bf16 row-major [rows, cols] with `ld` elements between rows; box = [box_rows, 64 cols], SWIZZLE_128B.
// Encoding a map costs a few microseconds of host time -- as much as a small GEMM runs -- and a prefill pass issues the same
// few hundred (buffer, shape) pairs on every request, so encoded maps are cached by their defining tuple.
struct TmapKey {
    const void* base; uint64_t rows, cols, ld; uint32_t box_rows;
    bool operator==(const TmapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = std::hash<const void*>()(k.base);
        for (uint64_t v : {k.rows, k.cols, k.ld, (uint64_t)k.box_rows}) h = h * 1000003u ^ std::hash<uint64_t>()(v);
        return h;
    }
};
static bool make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    static std::mutex mu;
    const TmapKey key{base, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *map = it->second; return true; }
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *map);
    return true;
}

// Per-stream stream-K scratch: partial tiles + arrival counters.  GEMMs on one stream are ordered (and a programmatic dependent only
// touches the scratch after its dependency wait), so one set per stream is enough; gemm_release_stream frees it with the stream.
struct GemmScratch { float* ws = nullptr; int* counters = nullptr; size_t ws_floats = 0; int n_counters = 0; };
static std::mutex g_scratch_mu;
static std::unordered_map<uint64_t, GemmScratch> g_scratch;
static uint64_t scratch_key(cudaStream_t st) { int dev = 0; cudaGetDevice(&dev); return ((uint64_t)(uintptr_t)st << 6) ^ (uint64_t)dev; }

static int scratch_get(cudaStream_t st, size_t ws_floats, int n_counters, GemmScratch* out) {
    std::lock_guard<std::mutex> g(g_scratch_mu);
    GemmScratch& sc = g_scratch[scratch_key(st)];
    if (sc.ws_floats < ws_floats || sc.n_counters < n_counters) {
        // (growing: the old buffers may still be in use by kernels in flight on this stream)
        if (sc.ws || sc.counters) { cudaStreamSynchronize(st); cudaFree(sc.ws); cudaFree(sc.counters); }
        sc.ws_floats = std::max(ws_floats, sc.ws_floats);
        sc.n_counters = std::max(n_counters, std::max(sc.n_counters, 4096));
        if (cudaMalloc(&sc.ws, sc.ws_floats * sizeof(float)) != cudaSuccess) { sc = GemmScratch{}; return (int)cudaErrorMemoryAllocation; }
        if (cudaMalloc(&sc.counters, (size_t)sc.n_counters * sizeof(int)) != cudaSuccess) { cudaFree(sc.ws); sc = GemmScratch{}; return (int)cudaErrorMemoryAllocation; }
        if (cudaMemsetAsync(sc.counters, 0, (size_t)sc.n_counters * sizeof(int), st) != cudaSuccess) return (int)cudaGetLastError();
    }
    *out = sc;
    return 0;
}
void gemm_release_stream(cudaStream_t st) {
    std::lock_guard<std::mutex> g(g_scratch_mu);
    auto it = g_scratch.find(scratch_key(st));
    if (it == g_scratch.end()) return;
    cudaFree(it->second.ws); cudaFree(it->second.counters);
    g_scratch.erase(it);
}
static int device_sms() {
    static int sms[CB_MAX_DEV] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= CB_MAX_DEV) return 148;
    if (!sms[dev]) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); sms[dev] = v > 0 ? v : 148; }
    return sms[dev];
}
// CRANE_B200_GEMM_MINIT: fewest k-steps worth giving a CTA (below that the fixed cost of a CTA -- setup, first operand latency,
// parking and re-reading partial tiles -- outweighs the parallelism)
static const int g_gemm_minit = [] { const char* e = getenv("CRANE_B200_GEMM_MINIT"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();

template <int BN, int MODE, bool SPLIT>
static int launch_tc(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K, const GemmEpi& epi,
                     bool stream_k) {
    using Cfg = GemmCfg<BN, SPLIT>;
    CUtensorMap tmA, tmAlo, tmB;
    if (!make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1001;
    if (!make_tmap_bf16(&tmAlo, SPLIT ? A_lo : A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1001;
    if (!make_tmap_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN)) return -1001;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(gemm_tc_kernel<BN, MODE, SPLIT>, (size_t)Cfg::SMEM_BYTES, seen)) return e;
    GemmSched sc = {};
    sc.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
    sc.tiles_n = (N + BN - 1) / BN;
    sc.KB = K / GEMM_BK;
    const long long tiles = (long long)sc.tiles_m * sc.tiles_n, iters = tiles * sc.KB;
    if (iters > 0x7fffffffLL) return -1000;
    const int cap = std::min(device_sms(), GEMM_MAX_CTAS);
    sc.unit = stream_k ? 1 : sc.KB;
    const long long units = iters / sc.unit;
    sc.G = (int)std::max<long long>(1, std::min<long long>(cap, stream_k ? iters / g_gemm_minit : tiles));
    sc.base = (int)(units / sc.G);
    sc.rem = (int)(units % sc.G);
    GemmScratch scratch;
    if (const int e = scratch_get(stream, (size_t)2 * GEMM_MAX_CTAS * GEMM_BM * 256, (int)std::max<long long>(tiles, 1), &scratch)) return e;
    sc.ws = scratch.ws;
    sc.counters = scratch.counters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sc.G);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (prefill_pdl()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return (int)cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, SPLIT>, tmA, tmAlo, tmB, epi, M, N, K, sc);
}

// Tile width and schedule from a cost model fitted to tools/gemm_probe.py on B200 (microseconds):
//   a k-step costs t_k(BN): the L2 hands a CTA its operands at ~0.33 us per k-step whatever the tile width (0.46 with the second
//   activation plane), only the 128 x 256 tile is slower than that because its MMAs are (0.34 / 0.60);
//   whole tiles: each CTA runs ceil(tiles / CTAs) tiles, the epilogue of all but the last hidden under the next tile's MMAs;
//   stream-K:    tiles x k-steps dealt evenly to all SMs, plus parking a partial tile and the last contributor re-reading c of them.
struct GemmPlan { int bn; bool stream_k; double cost; };
static GemmPlan gemm_plan(int M, int N, int K, bool split, int sms) {
    static const int g_mode = [] { const char* e = getenv("CRANE_B200_GEMM_SCHED"); return e ? atoi(e) : 0; }();   // 1: whole tiles, 2: stream-K
    const int tm = (M + GEMM_BM - 1) / GEMM_BM, KB = K / GEMM_BK;
    GemmPlan best{64, false, 1e30};
    for (int bn : {256, 128, 64}) {
        if (g_gemm_bn && bn != g_gemm_bn) continue;
        if (!g_gemm_bn && bn > 64 && N < bn) continue;
        const long long tiles = (long long)tm * ((N + bn - 1) / bn);
        const bool busy = tiles >= sms - 20;                                  // (nearly) every SM pulling operands: the L2 is the limit
        const double tk = split ? (bn == 256 ? (busy ? 0.68 : 0.60) : 0.46) : (bn == 256 ? 0.34 : 0.33);
        const double epi = 6.0 * bn / 256.0;
        const long long per = (tiles + std::min<long long>(tiles, sms) - 1) / std::min<long long>(tiles, sms);
        const double whole = 1.5 + (double)per * KB * tk + epi;
        const double it = (double)tiles * KB / sms;
        const double contrib = std::ceil((double)KB / std::max(it, 1.0)) + 1.0;
        const double tks = split ? (bn == 256 ? 0.68 : 0.46) : (bn == 256 ? 0.36 : 0.33);
        const double sk = it < g_gemm_minit ? 1e30 : 1.5 + it * tks + 1.6 * epi + contrib * 6.0 * bn / 256.0;
        if (g_mode != 2 && whole < best.cost) best = GemmPlan{bn, false, whole};
        // stream-K only below one wave of whole tiles: above it the extra partial-tile traffic costs more than the idle SMs
        if (g_mode != 1 && (tiles < sms || g_mode == 2) && tiles * KB >= 2LL * sms && sk < best.cost) best = GemmPlan{bn, true, sk};
    }
    return best;
}

template <int MODE>
static int launch_mode(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K,
                       const GemmEpi& epi, bool use_simt) {
    if (use_simt) {
        dim3 grid(N / 32, (M + 127) / 128);
        gemm_simt_kernel<MODE><<<grid, 128, 0, stream>>>(A, A_lo, lda, W, epi, M, N, K);
        return (int)cudaGetLastError();
    }
    const GemmPlan p = gemm_plan(M, N, K, A_lo != nullptr, std::min(device_sms(), GEMM_MAX_CTAS));
    if (A_lo != nullptr) {
        switch (p.bn) {
            case 256: return launch_tc<256, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, p.stream_k);
            case 128: return launch_tc<128, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, p.stream_k);
            default:  return launch_tc<64, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, p.stream_k);
        }
    }
    switch (p.bn) {
        case 256: return launch_tc<256, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, p.stream_k);
        case 128: return launch_tc<128, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, p.stream_k);
        default:  return launch_tc<64, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, p.stream_k);
    }
}

int gemm_bf16_launch(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K,
                     const GemmEpi& epi, bool use_simt) {
    if (M <= 0 || N <= 0 || K <= 0 || (K % GEMM_BK) != 0 || (N % 32) != 0 || (lda % 8) != 0) return -1000;
    switch (epi.mode) {
        case EPI_STORE_F32:      return launch_mode<EPI_STORE_F32>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        case EPI_STORE_BF16:     return launch_mode<EPI_STORE_BF16>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        case EPI_RESID_F32:      return launch_mode<EPI_RESID_F32>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        case EPI_SILU_MUL_BF16:  return launch_mode<EPI_SILU_MUL_BF16>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        case EPI_GELU_ERF_BF16:  return launch_mode<EPI_GELU_ERF_BF16>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        case EPI_GELU_TANH_BF16: return launch_mode<EPI_GELU_TANH_BF16>(stream, A, A_lo, lda, W, M, N, K, epi, use_simt);
        default: return -1000;
    }
}

}  // namespace cb
