// Chunkwise (WY / UT-transform) evaluation of the Gated-Delta-Net recurrence for prefill: 64 tokens per step instead of one.
//
// The reference evaluates  S <- S e^{g_t};  d = beta_t (v_t - S^T k_t);  S <- S + k_t d^T;  y_t = S^T q_t  token by token
// (crane-core/src/ops/gdn/backend.rs:90-156, kernels/cuda/gdn.cu:29-153 -- 2 x __syncthreads per token) and so did round 1 of this
// engine (gdn.cu: gdn_recur_kernel, 1.26 ms per layer at 4 096 tokens, latency-bound on 64 CTAs).  The same recurrence in chunk
// form is three launches, two of them parallel over all chunks and the serial one 64x shorter; kernels and algebra are in
// gdn_chunk_kernels.inc, the numpy restatement the tests use is oracle/gdn_chunked.py.
#include "gdn.cuh"
#include "prof.h"

namespace cb {

namespace {

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

#define CB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]

#include "gdn_chunk_kernels.inc"

template <int DK>
int launch_all(cudaStream_t st, const GdnArgs& a) {
    using Cfg = GdnChunkCfg<DK>;
    const int n_chunks = gdn_n_chunks(a.S);
    const GdnChunkWs w = gdn_chunk_ws_carve(a.chunk_ws, a.S, a.nv, a.dk, a.dv);
    const size_t prep_smem = gdn_chunk_prep_smem(DK, a.dv);
    static SmemOptIn s1, s2, s3;
    int r = ensure_dyn_smem(gdn_chunk_prep_kernel<DK>, prep_smem, s1);
    // warps of the serial kernel: 4 = one per scheduler.  8 (two per scheduler, each with half the columns) was measured slower at
    // DK = 128 -- 118 vs 112 us per layer: the A fragments of S^T are loaded by every warp and the hand-over barriers get wider
    constexpr int NW = 4;
    if (!r) r = ensure_dyn_smem(gdn_chunk_state_kernel<DK, NW>, Cfg::STATE_SMEM, s2);
    if (!r) r = ensure_dyn_smem(gdn_chunk_out_kernel<DK>, Cfg::OUT_SMEM, s3);
    // plain stream order (no programmatic early start): each kernel reads what the one before it wrote in full
    if (!r) r = launch_k(gdn_chunk_prep_kernel<DK>, dim3(n_chunks, a.nv), dim3(256), prep_smem, st, false, a, w);
    if (!r) r = launch_k(gdn_chunk_state_kernel<DK, NW>, dim3(a.nv * (a.dv / 16)), dim3(32 * NW), Cfg::STATE_SMEM, st, false, a, w, n_chunks);
    if (!r) r = launch_k(gdn_chunk_out_kernel<DK>, dim3(n_chunks, a.nv, a.dv / 64), dim3(256), Cfg::OUT_SMEM, st, false, a, w);
    return r;
}

}  // namespace

bool gdn_chunk_supported(const GdnArgs& a) {
    return a.glog != nullptr && a.chunk_ws != nullptr && a.S >= GDN_CHUNK && (a.dk == 64 || a.dk == 128 || a.dk == 256) &&
           a.dv % 64 == 0 && a.dv <= 256 && a.nk > 0 && a.nv % a.nk == 0;
}

int gdn_chunk_recur_launch(cudaStream_t st, const GdnArgs& a) {
    if (!gdn_chunk_supported(a)) return -1000;
    switch (a.dk) {
        case 64: return launch_all<64>(st, a);
        case 128: return launch_all<128>(st, a);
        default: return launch_all<256>(st, a);
    }
}

}  // namespace cb
