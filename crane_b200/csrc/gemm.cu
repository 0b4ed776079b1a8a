// tcgen05 + TMA GEMM for the prefill / ViT path.  See gemm.cuh for the design notes.
#include "gemm.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace cb {

// CRANE_B200_GEMM_CLUSTERS=1 turns the 2x2 TMA-multicast clusters on.  Measured on B200 (tools/gemm_probe.py) they do not pay:
// the L2 already de-duplicates concurrent requests for a line from a few CTAs, and what bounds the small-M GEMMs is per-CTA
// latency, not L2 -> SM bytes.  CRANE_B200_GEMM_BN=64|128|256 pins the tile width (sweeps).
static const bool g_gemm_clusters = [] { const char* e = getenv("CRANE_B200_GEMM_CLUSTERS"); return e && e[0] == '1'; }();
static const int g_gemm_bn = [] { const char* e = getenv("CRANE_B200_GEMM_BN"); return e ? atoi(e) : 0; }();

// =====================================================================================
// PTX wrappers (sm_100a)
// =====================================================================================
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// Same, delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster whose rank bit is set in `mask`.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
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
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
// tcgen05 kernel: one 128 x BN tile per CTA, 256 threads
//   warp 0 : TMA producer (one elected lane)       warp 1 : MMA issuer (one elected lane)
//   warp 2 : TMEM allocator / deallocator           warps 4-7 : epilogue (TMEM lane quadrant = warp % 4)
// =====================================================================================
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_SMEM_BUDGET = 192 * 1024;

template <int BN, bool SPLIT>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES * (SPLIT ? 2 : 1) + B_BYTES;     // [A hi | A lo (SPLIT) | B]
    static constexpr int STAGES = GEMM_SMEM_BUDGET / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// SPLIT: the activation operand comes as two bf16 planes A = hi + lo (lo = bf16(x - hi), ~16 mantissa bits together); the
// weight tile is fetched once and multiplied by both, accumulating into the same TMEM columns: D = hi.W^T + lo.W^T.
template <int BN, int MODE, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmB,
               GemmEpi ep, int M, int N, int K, int cx, int cy) {
    // (cx, cy) = thread-block cluster shape (1 or 2 each).  The cx CTAs of a cluster row work on the same 128 activation rows:
    // each fetches 128/cx of them and TMA-multicasts its part to the others; likewise the cy CTAs of a cluster column share the
    // BN weight rows.  At M ~ 450 these GEMMs are bound by L2->SM traffic, which this divides by up to 2.
    using Cfg = GemmCfg<BN, SPLIT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BN;
    const int KB = K / GEMM_BK;
    // optional timeline of this CTA (tools/gemm_probe.py): 8 globaltimer stamps per CTA
    unsigned long long* prof = ep.prof ? ep.prof + 8 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    auto stamp = [&](int i) {
        if (prof) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); prof[i] = t; }
    };
    if (threadIdx.x == 0) stamp(0);
    const bool clustered = cx * cy > 1;
    const int rx = blockIdx.x % cx, ry = blockIdx.y % cy;                 // position inside the cluster; rank = rx + ry * cx
    const uint16_t mask_a = (uint16_t)(((1u << cx) - 1u) << (ry * cx));   // CTAs sharing my activation rows
    const uint16_t mask_b = (uint16_t)((cy == 2 ? ((1u << cx) | 1u) : 1u) << rx);   // CTAs sharing my weight rows
    const uint16_t mask_all = mask_a | mask_b;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (SPLIT) tma_prefetch_desc(&tmAlo);
        tma_prefetch_desc(&tmB);
        // a stage is refilled by every CTA that multicasts into it: it is free once all of their consumers have drained it
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], cx + cy - 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    if (clustered) cluster_sync_all();          // peers' barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(1);                 // setup done

    pdl_launch_dependents();                        // the next kernel may begin its own setup / weight prefetch now
    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            int kb0 = 0;
            if (!clustered && !ep.w_dynamic) {
                // weights do not depend on the predecessor kernel: their first tiles go out before the dependency wait
                const int npre = min(KB, STAGES);
                for (int kb = 0; kb < npre; ++kb) {
                    uint8_t* a_dst = smem + kb * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[kb], Cfg::STAGE_BYTES);
                    tma_load_2d(a_dst + Cfg::A_BYTES * (SPLIT ? 2 : 1), &tmB, &full_bar[kb], kb * GEMM_BK, n0);
                }
                pdl_wait();
                for (int kb = 0; kb < npre; ++kb) {
                    uint8_t* a_dst = smem + kb * Cfg::STAGE_BYTES;
                    tma_load_2d(a_dst, &tmA, &full_bar[kb], kb * GEMM_BK, m0);
                    if (SPLIT) tma_load_2d(a_dst + Cfg::A_BYTES, &tmAlo, &full_bar[kb], kb * GEMM_BK, m0);
                }
                kb0 = npre;
                if (npre == STAGES) ph = 1; else s = npre;
            } else {
                pdl_wait();
            }
            for (int kb = kb0; kb < KB; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
                uint8_t* b_dst = a_dst + Cfg::A_BYTES * (SPLIT ? 2 : 1);
                mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);     // own parts + the parts the peers multicast here
                if (!clustered) {
                    tma_load_2d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m0);
                    if (SPLIT) tma_load_2d(a_dst + Cfg::A_BYTES, &tmAlo, &full_bar[s], kb * GEMM_BK, m0);
                    tma_load_2d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n0);
                } else {
                    const int ar = rx * (GEMM_BM / cx), br = ry * (BN / cy);      // my slice of the shared tiles (rows)
                    tma_load_2d_mc(a_dst + ar * 128, &tmA, &full_bar[s], kb * GEMM_BK, m0 + ar, mask_a);
                    if (SPLIT) tma_load_2d_mc(a_dst + Cfg::A_BYTES + ar * 128, &tmAlo, &full_bar[s], kb * GEMM_BK, m0 + ar, mask_a);
                    tma_load_2d_mc(b_dst + br * 128, &tmB, &full_bar[s], kb * GEMM_BK, n0 + br, mask_b);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (kb == 0) stamp(2);              // first operands landed
                const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t adesc = make_smem_desc_sw128(a_addr);
                const uint64_t aldesc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
                const uint64_t bdesc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES * (SPLIT ? 2 : 1));
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    // +32 B per K=16 step inside the 128 B swizzle atom -> +2 in the >>4-encoded start address
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    if (SPLIT) umma_bf16(tmem_base, aldesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                }
                // frees the smem stage (here and in every CTA that refills it) once these MMAs retire
                if (clustered) umma_commit_mc(&empty_bar[s], mask_all); else umma_commit(&empty_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(tmem_full_bar);                 // accumulator complete
            stamp(3);                                   // last MMA issued
        }
    } else if (warp >= 4) {
        pdl_wait();                                 // the epilogue reads / overwrites activations of the predecessor too
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 128) stamp(4);               // accumulator ready
        const int q = warp & 3;
        uint8_t* stg = smem + q * EPI_WARP_BYTES;       // the pipeline stages are dead once the accumulator is complete
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            const int n = n0 + c * 32;
            if (n < N) {                                // warp-uniform
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                epi_block<MODE>(ep, stg, m0 + q * 32, M, n, v, lane);
            }
        }
        if (threadIdx.x == 128) stamp(5);               // this warp's rows stored
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(6);                 // whole CTA done
    if (clustered) cluster_sync_all();          // nobody leaves while a peer may still signal its barriers
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, BN); }
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

template <int BN, int MODE, bool SPLIT>
static int launch_tc(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K, const GemmEpi& epi,
                     int cx, int cy) {
    using Cfg = GemmCfg<BN, SPLIT>;
    CUtensorMap tmA, tmAlo, tmB;
    if (!make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM / cx)) return -1001;
    if (!make_tmap_bf16(&tmAlo, SPLIT ? A_lo : A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM / cx)) return -1001;
    if (!make_tmap_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN / cy)) return -1001;
    static SmemOptIn seen;
    if (const int e = ensure_dyn_smem(gemm_tc_kernel<BN, MODE, SPLIT>, (size_t)Cfg::SMEM_BYTES, seen)) return e;
    dim3 grid((N + BN - 1) / BN, ((M + GEMM_BM - 1) / GEMM_BM + cy - 1) / cy * cy);   // whole clusters; surplus tiles are all-OOB
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (prefill_pdl()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cx * cy > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cx; attr[na].val.clusterDim.y = cy; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return (int)cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, SPLIT>, tmA, tmAlo, tmB, epi, M, N, K, cx, cy);
}

template <int MODE>
static int launch_mode(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K,
                       const GemmEpi& epi, bool use_simt) {
    if (use_simt) {
        dim3 grid(N / 32, (M + 127) / 128);
        gemm_simt_kernel<MODE><<<grid, 128, 0, stream>>>(A, A_lo, lda, W, epi, M, N, K);
        return (int)cudaGetLastError();
    }
    // Tile width and cluster shape from a two-term cost model (times in "A-tile loads"): the tensor pipe needs
    // waves x k-steps x (BN / 64) x (split ? 2 : 1); the L2 -> SM fabric moves, per tile and k-step, the A planes / cx plus the
    // B tile / cy.  At M ~ 450 the second term dominates, which is what the clusters attack.
    const int mt = (M + GEMM_BM - 1) / GEMM_BM;
    const bool split = A_lo != nullptr;
    int best_bn = 64, best_cx = 1, best_cy = 1;
    double best = 1e30;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        if (g_gemm_bn && bn != g_gemm_bn) continue;
        const int nt = (N + bn - 1) / bn;
        for (int cy = 1; cy <= (mt > 1 && g_gemm_clusters ? 2 : 1); ++cy)
            for (int cx = 1; cx <= ((nt % 2) == 0 && g_gemm_clusters ? 2 : 1); ++cx) {
                const int mtp = (mt + cy - 1) / cy * cy;
                const int tiles = mtp * nt;
                const int waves = (tiles + 147) / 148;
                const double mma = (double)waves * (bn / 64.0) * (split ? 2.0 : 1.0) * 0.55;     // 128x64x64 MMA vs one 16 KB tile over the fabric
                const double bytes = (double)tiles * ((split ? 2.0 : 1.0) / cx + (bn / 128.0) / cy) / 148.0;
                const double cost = std::max(mma, bytes) + 0.15 * std::min(mma, bytes) + (cx * cy > 1 ? 0.02 : 0.0);
                if (cost < best) { best = cost; best_bn = bn; best_cx = cx; best_cy = cy; }
            }
    }
    if (A_lo != nullptr) {
        switch (best_bn) {
            case 256: return launch_tc<256, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, best_cx, best_cy);
            case 128: return launch_tc<128, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, best_cx, best_cy);
            default:  return launch_tc<64, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi, best_cx, best_cy);
        }
    }
    switch (best_bn) {
        case 256: return launch_tc<256, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, best_cx, best_cy);
        case 128: return launch_tc<128, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, best_cx, best_cy);
        default:  return launch_tc<64, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi, best_cx, best_cy);
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
