// Standalone GEMV micro-benchmark (compiled ON the GPU box so pipeline parameters can be swept in one call):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGV_WARPS=8 -DGV_DEPTH=5 tools/gemv_bench.cu -o /tmp/gb && /tmp/gb
// 1) isolated kernels per shape (weights rotated through > L2 worth of copies), 2) the PDL-chained weight stream of a whole
// Qwen3-VL-2B decode step (28 x [qkv, o, gate_up, down] + lm_head, attention omitted).
#include "../crane_b200/csrc/decode.cu"
#include <vector>
using namespace cb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

static int run(cudaStream_t st, int epi, bool norm, const bf16* W, int N, int K, const float* x, const float* nw, float* y, GemvArgs extra, bool pdl) {
    GemvArgs g = extra;
    g.W = W; g.N = N; g.K = K; g.x = x; g.ldx = K; g.norm_w = nw; g.eps = 1e-6f; g.y = y; g.ldy = (epi == GEMV_SILU_MUL) ? N / 2 : N;
    return gemv_launch(st, 1, epi, norm, g, 148, pdl);
}

int main() {
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int H = 2048, I = 6144, V = 151936, QKV = 4096, L = 28;
    const size_t per_layer = (size_t)QKV * H + (size_t)H * H + (size_t)2 * I * H + (size_t)H * I;
    const size_t total = per_layer * L + (size_t)V * H;
    bf16* W; CK(cudaMalloc(&W, total * 2)); CK(cudaMemset(W, 0x11, total * 2));
    float *x, *nw, *y, *act; CK(cudaMalloc(&x, 65536 * 4)); CK(cudaMalloc(&nw, 65536 * 4)); CK(cudaMalloc(&y, (size_t)V * 4)); CK(cudaMalloc(&act, 65536 * 4));
    CK(cudaMemset(x, 0, 65536 * 4)); CK(cudaMemset(nw, 0, 65536 * 4)); CK(cudaMemset(act, 0, 65536 * 4));
    float* pv; int* pi; unsigned int* tk; SeqState* ss; uint32_t* ot;
    CK(cudaMalloc(&pv, 4096)); CK(cudaMalloc(&pi, 4096)); CK(cudaMalloc(&tk, 4)); CK(cudaMemset(tk, 0, 4)); CK(cudaMalloc(&ss, sizeof(SeqState))); CK(cudaMemset(ss, 0, sizeof(SeqState)));
    CK(cudaMalloc(&ot, 4 * 100000));
    GemvArgs ex = {}; ex.part_val = pv; ex.part_idx = pi; ex.ticket = tk; ex.state = ss; ex.out_tokens = nullptr; ex.out_stride = 0; ex.embed = W; ex.x_next = x; ex.H = H; ex.advance = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Shape { const char* name; int epi; bool norm; int N, K; };
    Shape shapes[] = {{"qkv", GEMV_STORE, true, QKV, H}, {"o", GEMV_RESID, false, H, H}, {"gate_up", GEMV_SILU_MUL, true, 2 * I, H},
                      {"down", GEMV_RESID, false, H, I}, {"lm_head", GEMV_LOGITS_ARGMAX, true, V, H}};
    printf("GV_WARPS=%d GV_DEPTH=%d\n", GV_WARPS, GV_DEPTH);
    for (auto& s : shapes) {
        const size_t wsz = (size_t)s.N * s.K;
        const int copies = (int)std::max<size_t>(1, std::min<size_t>(total / wsz, 64));
        const int iters = 200;
        for (int it = 0; it < 20; ++it) if (run(st, s.epi, s.norm, W + (size_t)(it % copies) * wsz, s.N, s.K, s.K == H ? x : act, nw, y, ex, false)) return 2;
        CK(cudaStreamSynchronize(st));
        cudaEventRecord(e0, st);
        for (int it = 0; it < iters; ++it) run(st, s.epi, s.norm, W + (size_t)(it % copies) * wsz, s.N, s.K, s.K == H ? x : act, nw, y, ex, false);
        cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / iters;
        printf("  isolated %-8s N=%6d K=%5d : %7.2f us  %7.1f GB/s\n", s.name, s.N, s.K, us, wsz * 2 / us / 1e3);
    }
    for (int pdl = 0; pdl <= 1; ++pdl) {
        auto step = [&]() {
            const bf16* w = W;
            for (int l = 0; l < L; ++l) {
                run(st, GEMV_STORE, true, w, QKV, H, x, nw, y, ex, pdl); w += (size_t)QKV * H;
                run(st, GEMV_RESID, false, w, H, H, x, nullptr, act, ex, pdl); w += (size_t)H * H;
                run(st, GEMV_SILU_MUL, true, w, 2 * I, H, x, nw, act, ex, pdl); w += (size_t)2 * I * H;
                run(st, GEMV_RESID, false, w, H, I, act, nullptr, x, ex, pdl); w += (size_t)H * I;
            }
            run(st, GEMV_LOGITS_ARGMAX, true, w, V, H, x, nw, y, ex, pdl);
        };
        for (int i = 0; i < 3; ++i) step();
        CK(cudaStreamSynchronize(st));
        cudaEventRecord(e0, st);
        const int iters = 20;
        for (int i = 0; i < iters; ++i) step();
        cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / iters;
        printf("  chained step (113 GEMVs, no attention) pdl=%d : %8.1f us  %7.1f GB/s\n", pdl, us, total * 2 / us / 1e3);
    }
    return 0;
}
