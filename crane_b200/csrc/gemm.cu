// tcgen05 + TMA GEMM for the prefill / ViT path.  See gemm.cuh for the design notes.
#include "gemm.cuh"

#include <mutex>

namespace cb {

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
// Fused epilogue: 32 consecutive output columns of one row
// =====================================================================================
template <int MODE>
__device__ __forceinline__ void epi_store32(const GemmEpi& ep, int m, int n, float* v) {
    if (ep.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(ep.bias + n + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
    }
    if constexpr (MODE == EPI_STORE_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if constexpr (MODE == EPI_RESID_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float4 r = *reinterpret_cast<float4*>(o + j);
            r.x += v[j]; r.y += v[j + 1]; r.z += v[j + 2]; r.w += v[j + 3];
            *reinterpret_cast<float4*>(o + j) = r;
        }
    } else if constexpr (MODE == EPI_SILU_MUL_BF16) {
        const size_t off = (size_t)m * ep.ldo + (n >> 1);
        uint32_t p[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            split_bf16x2(silu_f(v[4 * j]) * v[4 * j + 1], silu_f(v[4 * j + 2]) * v[4 * j + 3], p[j], q[j]);
        bf16* o = reinterpret_cast<bf16*>(ep.out) + off;
        *reinterpret_cast<uint4*>(o) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4*>(o + 8) = make_uint4(p[4], p[5], p[6], p[7]);
        if (ep.out_lo != nullptr) {
            bf16* ol = reinterpret_cast<bf16*>(ep.out_lo) + off;
            *reinterpret_cast<uint4*>(ol) = make_uint4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<uint4*>(ol + 8) = make_uint4(q[4], q[5], q[6], q[7]);
        }
    } else {
        if constexpr (MODE == EPI_GELU_ERF_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
        } else if constexpr (MODE == EPI_GELU_TANH_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
        }
        const size_t off = (size_t)m * ep.ldo + n;
        bf16* o = reinterpret_cast<bf16*>(ep.out) + off;
        bf16* ol = ep.out_lo != nullptr ? reinterpret_cast<bf16*>(ep.out_lo) + off : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            uint32_t p[4], q[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) split_bf16x2(v[j + 2 * t], v[j + 2 * t + 1], p[t], q[t]);
            *reinterpret_cast<uint4*>(o + j) = make_uint4(p[0], p[1], p[2], p[3]);
            if (ol != nullptr) *reinterpret_cast<uint4*>(ol + j) = make_uint4(q[0], q[1], q[2], q[3]);
        }
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
               GemmEpi ep, int M, int N, int K) {
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

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (SPLIT) tma_prefetch_desc(&tmAlo);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
                uint8_t* b_dst = a_dst + Cfg::A_BYTES * (SPLIT ? 2 : 1);
                mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                tma_load_2d(a_dst, &tmA, &full_bar[s], kb * GEMM_BK, m0);
                if (SPLIT) tma_load_2d(a_dst + Cfg::A_BYTES, &tmAlo, &full_bar[s], kb * GEMM_BK, m0);
                tma_load_2d(b_dst, &tmB, &full_bar[s], kb * GEMM_BK, n0);
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
                umma_commit(&empty_bar[s]);             // frees the smem stage once these MMAs retire
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(tmem_full_bar);                 // accumulator complete
        }
    } else if (warp >= 4) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            const int n = n0 + c * 32;
            if (m < M && n < N) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                epi_store32<MODE>(ep, m, n, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
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
    if (m < M && n0 < N) epi_store32<MODE>(ep, m, n0, acc);
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
static bool make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int MODE, bool SPLIT>
static int launch_tc(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K, const GemmEpi& epi) {
    using Cfg = GemmCfg<BN, SPLIT>;
    CUtensorMap tmA, tmAlo, tmB;
    if (!make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1001;
    if (!make_tmap_bf16(&tmAlo, SPLIT ? A_lo : A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM)) return -1001;
    if (!make_tmap_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN)) return -1001;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((N + BN - 1) / BN, (M + GEMM_BM - 1) / GEMM_BM);
    gemm_tc_kernel<BN, MODE, SPLIT><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmAlo, tmB, epi, M, N, K);
    return (int)cudaGetLastError();
}

template <int MODE>
static int launch_mode(cudaStream_t stream, const bf16* A, const bf16* A_lo, int lda, const bf16* W, int M, int N, int K,
                       const GemmEpi& epi, bool use_simt) {
    if (use_simt) {
        dim3 grid(N / 32, (M + 127) / 128);
        gemm_simt_kernel<MODE><<<grid, 128, 0, stream>>>(A, A_lo, lda, W, epi, M, N, K);
        return (int)cudaGetLastError();
    }
    // Tile width: fill the 148 SMs in as few waves as possible, wider tiles on ties.
    const int mt = (M + GEMM_BM - 1) / GEMM_BM;
    int best_bn = 64; double best = -1.0;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const int tiles = mt * ((N + bn - 1) / bn);
        const int waves = (tiles + 147) / 148;
        double eff = (double)tiles / (waves * 148.0);
        eff *= (bn == 256 ? 1.0 : bn == 128 ? 0.95 : 0.80);   // narrower tiles re-read A from smem more often
        if (eff > best) { best = eff; best_bn = bn; }
    }
    if (A_lo != nullptr) {
        switch (best_bn) {
            case 256: return launch_tc<256, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi);
            case 128: return launch_tc<128, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi);
            default:  return launch_tc<64, MODE, true>(stream, A, A_lo, lda, W, M, N, K, epi);
        }
    }
    switch (best_bn) {
        case 256: return launch_tc<256, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi);
        case 128: return launch_tc<128, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi);
        default:  return launch_tc<64, MODE, false>(stream, A, nullptr, lda, W, M, N, K, epi);
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
