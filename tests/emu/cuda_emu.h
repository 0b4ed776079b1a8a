// TEST INFRASTRUCTURE.  A minimal host-side stand-in for the CUDA execution model, just enough to run the chunkwise
// Gated-Delta-Net kernels (crane_b200/csrc/gdn_chunk_kernels.inc) on CPU cores: one OS thread per CUDA thread, one CTA at a time,
// pthread barriers for __syncthreads and for the warp-collective instructions (ldmatrix, mma.sync), whose register <-> matrix
// element mappings are written here from the PTX ISA's fragment tables, independently of the kernels that use them.
// It exists so the index arithmetic of those kernels is checked by the CPU test suite against the numpy oracle; the product
// never links it (libcrane_b200.so has no CPU path), and no timing taken here means anything.
#pragma once

#include <cuda_bf16.h>
#include <vector_types.h>
#include <vector_functions.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __forceinline__
#undef __launch_bounds__
#undef __restrict__
#undef __host__
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __host__

namespace cuda_emu {

struct Dim3 { int x = 1, y = 1, z = 1; };
struct WarpCtx {
    pthread_barrier_t bar;
    uint32_t a[32][4], b[32][2];
    uint32_t addr[32];
};
struct BlockCtx {
    pthread_barrier_t bar;
    std::vector<WarpCtx> warps;
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
};
inline thread_local Dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_block = nullptr;
inline thread_local int t_lane = 0, t_warp = 0;

inline float bf16_bits_to_float(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

template <typename F>
void launch(Dim3 grid, Dim3 block, size_t smem_bytes, F&& body) {
    const int nthreads = block.x * block.y * block.z;
    assert(nthreads % 32 == 0);
    assert(smem_bytes <= 227 * 1024 && "more dynamic shared memory than an sm_100 CTA may have");
    for (int bz = 0; bz < grid.z; ++bz)
        for (int by = 0; by < grid.y; ++by)
            for (int bx = 0; bx < grid.x; ++bx) {
                BlockCtx ctx;
                pthread_barrier_init(&ctx.bar, nullptr, nthreads);
                ctx.warps = std::vector<WarpCtx>(nthreads / 32);
                for (auto& w : ctx.warps) pthread_barrier_init(&w.bar, nullptr, 32);
                void* mem = nullptr;
                if (posix_memalign(&mem, 1024, smem_bytes + 1024)) abort();
                memset(mem, 0xA5, smem_bytes + 1024);            // shared memory starts out as garbage
                ctx.smem = static_cast<unsigned char*>(mem);
                ctx.smem_bytes = smem_bytes;
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (int tix = 0; tix < nthreads; ++tix)
                    ts.emplace_back([&, tix] {
                        t_threadIdx.x = tix; t_threadIdx.y = 0; t_threadIdx.z = 0;
                        t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
                        t_blockDim = block; t_gridDim = grid;
                        t_block = &ctx; t_lane = tix & 31; t_warp = tix >> 5;
                        body();
                    });
                for (auto& th : ts) th.join();
                for (size_t i = smem_bytes; i < smem_bytes + 1024; ++i)
                    if (ctx.smem[i] != 0xA5) { fprintf(stderr, "cuda_emu: write past the end of shared memory\n"); abort(); }
                free(mem);
                for (auto& w : ctx.warps) pthread_barrier_destroy(&w.bar);
                pthread_barrier_destroy(&ctx.bar);
            }
}

inline void warp_sync() { pthread_barrier_wait(&t_block->warps[t_warp].bar); }

}  // namespace cuda_emu

#define threadIdx cuda_emu::t_threadIdx
#define blockIdx cuda_emu::t_blockIdx
#define blockDim cuda_emu::t_blockDim
#define gridDim cuda_emu::t_gridDim
#define CB_DYN_SMEM(name) unsigned char* name = cuda_emu::t_block->smem

using std::min;
using std::max;

inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::t_block->bar); }
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

inline uint32_t smem_u32(const void* p) {
    const ptrdiff_t off = static_cast<const unsigned char*>(p) - cuda_emu::t_block->smem;
    assert(off >= 0 && (size_t)off < cuda_emu::t_block->smem_bytes && "pointer is not inside this CTA's shared memory");
    return (uint32_t)off;
}

inline void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    assert(dst % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 && "cp.async 16-byte alignment");
    assert((size_t)dst + 16 <= cuda_emu::t_block->smem_bytes);
    unsigned char* d = cuda_emu::t_block->smem + dst;
    memset(d, 0, 16);
    memcpy(d, src, src_bytes);
}
// mbarrier + bulk (TMA 1-D) copy.  The copy itself is performed by the issuing thread at issue, but its completion is published
// through the barrier word exactly as the hardware does it (expect_tx arms a byte count, each copy retires its bytes, the phase
// completes when the count reaches zero) and mbar_wait really waits, with acquire / release ordering -- so a consumer that skips
// the wait, or a stage that is overwritten while it is still being read, is a data race ThreadSanitizer reports (the TSan build of
// this harness is the CPU stand-in for compute-sanitizer's racecheck).  Word layout: low 32 bits = completed phases, high 32 = pending bytes.
inline std::atomic<uint64_t>& mbar_word(uint64_t* bar) { return *reinterpret_cast<std::atomic<uint64_t>*>(bar); }
inline void mbar_init(uint64_t* bar, uint32_t) { mbar_word(bar).store(0, std::memory_order_release); }
inline void fence_barrier_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { mbar_word(bar).fetch_add((uint64_t)bytes << 32, std::memory_order_acq_rel); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (((uint32_t)mbar_word(bar).load(std::memory_order_acquire) & 1u) == parity) sched_yield();
}
inline void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    assert(bytes % 16 == 0 && reinterpret_cast<uintptr_t>(gsrc) % 16 == 0 && smem_u32(smem_dst) % 16 == 0 && "cp.async.bulk alignment");
    assert((size_t)smem_u32(smem_dst) + bytes <= cuda_emu::t_block->smem_bytes);
    memcpy(smem_dst, gsrc, bytes);
    uint64_t old = mbar_word(bar).load(std::memory_order_relaxed), next;
    do {                                               // retire the bytes; the copy that brings the count to zero completes the phase
        const uint64_t pending = (old >> 32) - bytes;
        next = (pending << 32) | (uint32_t)((uint32_t)old + (pending == 0 ? 1u : 0u));
    } while (!mbar_word(bar).compare_exchange_weak(old, next, std::memory_order_acq_rel));
}
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}

inline uint32_t pack_bf16(float a, float b) {
    __nv_bfloat16 x = __float2bfloat16_rn(a), y = __float2bfloat16_rn(b);
    uint16_t xb, yb;
    memcpy(&xb, &x, 2); memcpy(&yb, &y, 2);
    return (uint32_t)xb | ((uint32_t)yb << 16);
}
inline void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16(a, b);
    lo = pack_bf16(a - cuda_emu::bf16_bits_to_float(hi & 0xffff), b - cuda_emu::bf16_bits_to_float(hi >> 16));
}

// ldmatrix.sync.aligned.m8n8.x4[.trans].shared.b16: lanes 8m .. 8m+7 supply the eight row addresses of matrix m; lane (g = lane/4,
// t = lane%4) receives in register m the elements [g][2t], [g][2t+1] of matrix m (plain) or [2t][g], [2t+1][g] (.trans), low half first.
inline void ldmatrix_impl(uint32_t (&r)[4], uint32_t addr, bool trans) {
    auto& w = cuda_emu::t_block->warps[cuda_emu::t_warp];
    const int lane = cuda_emu::t_lane, g = lane >> 2, t = lane & 3;
    assert(addr % 16 == 0 && "ldmatrix row address must be 16-byte aligned");
    assert((size_t)addr + 16 <= cuda_emu::t_block->smem_bytes);
    w.addr[lane] = addr;
    cuda_emu::warp_sync();
    const unsigned char* sm = cuda_emu::t_block->smem;
    for (int m = 0; m < 4; ++m) {
        uint16_t e0, e1;
        if (!trans) {
            memcpy(&e0, sm + w.addr[8 * m + g] + 2 * (2 * t), 2);
            memcpy(&e1, sm + w.addr[8 * m + g] + 2 * (2 * t + 1), 2);
        } else {
            memcpy(&e0, sm + w.addr[8 * m + 2 * t] + 2 * g, 2);
            memcpy(&e1, sm + w.addr[8 * m + 2 * t + 1] + 2 * g, 2);
        }
        r[m] = (uint32_t)e0 | ((uint32_t)e1 << 16);
    }
    cuda_emu::warp_sync();
}
inline void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    uint32_t r[4]; ldmatrix_impl(r, addr, false); r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}
inline void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    uint32_t r[4]; ldmatrix_impl(r, addr, true); r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}

// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32, D = A B + D.  Fragments (PTX ISA, "Matrix Fragments for mma.m16n8k16"):
//   A (16 x 16, row): register i of lane (g, t) holds A[g + 8 (i & 1)][2 t + 8 (i >> 1) + {0, 1}]
//   B (16 x 8,  col): register i holds B[2 t + 8 i + {0, 1}][g]
//   C / D (16 x 8):   element i holds C[g + 8 (i >> 1)][2 t + (i & 1)]
inline void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    auto& w = cuda_emu::t_block->warps[cuda_emu::t_warp];
    const int lane = cuda_emu::t_lane, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; ++i) w.a[lane][i] = a[i];
    w.b[lane][0] = b0; w.b[lane][1] = b1;
    cuda_emu::warp_sync();
    auto A = [&](int row, int k) {
        const int ln = (row & 7) * 4 + ((k & 7) >> 1), reg = (row >> 3) + 2 * (k >> 3);
        const uint32_t v = w.a[ln][reg];
        return cuda_emu::bf16_bits_to_float((k & 1) ? (uint16_t)(v >> 16) : (uint16_t)(v & 0xffff));
    };
    auto B = [&](int k, int n) {
        const int ln = n * 4 + ((k & 7) >> 1), reg = k >> 3;
        const uint32_t v = w.b[ln][reg];
        return cuda_emu::bf16_bits_to_float((k & 1) ? (uint16_t)(v >> 16) : (uint16_t)(v & 0xffff));
    };
    float d[4];
    for (int i = 0; i < 4; ++i) {
        const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 16; ++k) acc += A(row, k) * B(k, col);
        d[i] = acc;
    }
    cuda_emu::warp_sync();
    for (int i = 0; i < 4; ++i) c[i] = d[i];
}
