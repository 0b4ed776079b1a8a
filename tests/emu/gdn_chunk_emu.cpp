// TEST INFRASTRUCTURE.  Runs the three chunkwise Gated-Delta-Net kernels of crane_b200/csrc/gdn_chunk_kernels.inc on CPU threads
// (cuda_emu.h) so tests/test_gdn_chunk_emu.py can compare every intermediate with the numpy oracle.  Not part of the product.
#include "cuda_emu.h"

#include "../../crane_b200/csrc/gdn_args.h"

namespace cb {
#include "../../crane_b200/csrc/gdn_chunk_kernels.inc"
}  // namespace cb

using namespace cb;

template <int DK, int NW = 4>
static int run(const GdnArgs& a, const GdnChunkWs& w) {
    using Cfg = GdnChunkCfg<DK>;
    const int n_chunks = gdn_n_chunks(a.S);
    cuda_emu::Dim3 g1{n_chunks, a.nv, 1}, b1{256, 1, 1};
    cuda_emu::launch(g1, b1, gdn_chunk_prep_smem(DK, a.dv), [&] { gdn_chunk_prep_kernel<DK>(a, w); });
    cuda_emu::Dim3 g2{a.nv * (a.dv / 16), 1, 1}, b2{32 * NW, 1, 1};
    cuda_emu::launch(g2, b2, Cfg::STATE_SMEM, [&] { gdn_chunk_state_kernel<DK, NW>(a, w, n_chunks); });
    cuda_emu::Dim3 g3{n_chunks, a.nv, a.dv / 64}, b3{256, 1, 1};
    cuda_emu::launch(g3, b3, Cfg::OUT_SMEM, [&] { gdn_chunk_out_kernel<DK>(a, w); });
    return 0;
}

extern "C" {

size_t gdn_chunk_emu_ws_bytes(int S, int nv, int dk, int dv) { return gdn_chunk_ws_bytes(S, nv, dk, dv); }

// Byte offsets of the scratch sub-buffers, in GdnChunkWs order (w, kt, qt, p, ut, gc, st, dt).
void gdn_chunk_emu_ws_offsets(int S, int nv, int dk, int dv, size_t* out) {
    unsigned char* base = reinterpret_cast<unsigned char*>(uintptr_t(4096));
    const GdnChunkWs w = gdn_chunk_ws_carve(base, S, nv, dk, dv);
    const void* p[8] = {w.w, w.kt, w.qt, w.p, w.ut, w.gc, w.st, w.dt};
    for (int i = 0; i < 8; ++i) out[i] = (size_t)(static_cast<const unsigned char*>(p[i]) - base);
}

// qn, kn [S, nk, dk]; conv_out [S, conv_dim] (only the v part is read); gb [S, nv, 2]; glog [S, nv]; rec_state [nv, dk, dv] in/out;
// y [S, nv, dv] out; ws = gdn_chunk_emu_ws_bytes bytes, 256-byte aligned.
int gdn_chunk_emu_run(const float* qn, const float* kn, const float* conv_out, const float* gb, const float* glog, float* rec_state,
                      float* y, void* ws, int S, int nk, int nv, int dk, int dv, int state_warps) {
    GdnArgs a = {};
    a.S = S; a.nk = nk; a.nv = nv; a.dk = dk; a.dv = dv; a.ck = 4;
    a.qn = const_cast<float*>(qn); a.kn = const_cast<float*>(kn); a.conv_out = const_cast<float*>(conv_out);
    a.gb = const_cast<float*>(gb); a.glog = const_cast<float*>(glog); a.rec_state = rec_state; a.y = y; a.chunk_ws = ws;
    if (dv % 64 != 0 || nv % nk != 0) return -1;
    const GdnChunkWs w = gdn_chunk_ws_carve(ws, S, nv, dk, dv);
    switch (dk) {
        case 64: return run<64>(a, w);
        case 128: return state_warps == 8 ? run<128, 8>(a, w) : run<128>(a, w);
        case 256: return run<256>(a, w);
        default: return -1;
    }
}

}  // extern "C"

#ifdef GDN_EMU_MAIN
// Stand-alone run on random inputs (no numeric check: tests/test_gdn_chunk_emu.py does that through the shared library).  Built with
// -fsanitize=thread this is the race check of the three kernels: CUDA threads are OS threads, __syncthreads / warp collectives are
// pthread barriers, mbarriers are acquire / release atomics -- every shared- or global-memory hazard between the threads of a CTA that
// is not ordered by one of those is a data race TSan reports.
#include <random>
int main(int argc, char** argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 130, dk = argc > 2 ? atoi(argv[2]) : 128, dv = 64, nk = 1, nv = 1;
    const int state_warps = argc > 3 ? atoi(argv[3]) : 4;
    const int conv_dim = 2 * nk * dk + nv * dv;
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> qn((size_t)S * nk * dk), kn(qn.size()), conv((size_t)S * conv_dim), gb((size_t)S * nv * 2), gl((size_t)S * nv),
        st((size_t)nv * dk * dv), y((size_t)S * nv * dv);
    for (auto& x : qn) x = nd(rng) * 0.05f;
    for (auto& x : kn) x = nd(rng) * 0.09f;
    for (auto& x : conv) x = nd(rng);
    for (auto& x : st) x = nd(rng) * 0.3f;
    for (int i = 0; i < S * nv; ++i) { gl[i] = -0.5f; gb[2 * i] = expf(-0.5f); gb[2 * i + 1] = 0.5f; }
    void* ws = nullptr;
    if (posix_memalign(&ws, 256, gdn_chunk_emu_ws_bytes(S, nv, dk, dv))) return 2;
    memset(ws, 0, gdn_chunk_emu_ws_bytes(S, nv, dk, dv));
    const int rc = gdn_chunk_emu_run(qn.data(), kn.data(), conv.data(), gb.data(), gl.data(), st.data(), y.data(), ws, S, nk, nv, dk, dv, state_warps);
    double acc = 0;
    for (float v : y) acc += v;
    printf("gdn chunk emu: rc %d, sum(y) %.6f\n", rc, acc);
    return rc;
}
#endif
