// Standalone quantised-GEMV micro-benchmark (compiled ON the GPU box so parameters can be swept in one call):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DQG_R_DEF=4 -DQG_DEPTH_DEF=2 -DQG_W1_DEF=16 -DQG_W4_DEF=12] tools/qgemv_bench.cu -o /tmp/qb && /tmp/qb
// Shapes: the linears of a Qwen3-8B Q4_K_M-like decode step, B = 1 and 4; isolated (weights rotated through > L2) and as a
// PDL-chained step (attention omitted).
#include "../crane_b200/csrc/quant.cu"
#include <algorithm>
#include <cstdio>
using namespace cb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Shape { const char* name; int qt, epi; bool norm; int N, K; };
static GemvArgs ex;
static float *x, *nw, *y, *act;
static unsigned char* xqb;

static int run(cudaStream_t st, int B, const Shape& s, const unsigned char* W, bool pdl) {
    QGemvArgs q = {};
    q.g = ex;
    q.g.W = reinterpret_cast<const bf16*>(W); q.g.N = s.N; q.g.K = s.K; q.g.x = s.K == 4096 ? x : act; q.g.ldx = s.K; q.g.norm_w = nw; q.g.eps = 1e-6f;
    q.g.y = s.epi == GEMV_SILU_MUL ? act : y; q.g.ldy = (s.epi == GEMV_SILU_MUL) ? s.N / 2 : s.N;
    q.qtype = s.qt; q.epi = s.epi; q.norm = s.norm; q.xq = xqb;
    if (s.name[0] != 'k' && s.name[0] != 'v') { int r = xquant_launch(st, B, q.g.x, s.K, s.K, s.norm ? nw : nullptr, 1e-6f, xq_mode_for(q.qtype), xqb, pdl); if (r) return r; }
    return qgemv_launch(st, B, q, 148, pdl);
}

int main() {
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int H = 4096, I = 12288, V = 151936, L = 36;
    Shape q{"q", QT_Q4_K, GEMV_STORE, true, 4096, H}, k{"k", QT_Q4_K, GEMV_STORE, true, 1024, H}, v{"v", QT_Q6_K, GEMV_STORE, true, 1024, H},
        o{"o", QT_Q4_K, GEMV_RESID, false, H, 4096}, gu{"gate_up", QT_Q4_K, GEMV_SILU_MUL, true, 2 * I, H}, dn{"down", QT_Q6_K, GEMV_RESID, false, H, I},
        head{"lm_head", QT_Q6_K, GEMV_LOGITS_ARGMAX, true, V, H};
    Shape layer[] = {q, k, v, o, gu, dn};
    auto wbytes = [](const Shape& s) { return (size_t)s.N * (s.K / 256) * q_sb_bytes(s.qt); };
    size_t per_layer = 0;
    for (auto& s : layer) per_layer += wbytes(s);
    const size_t total = per_layer * L + wbytes(head);
    unsigned char* W; CK(cudaMalloc(&W, total)); CK(cudaMemset(W, 0x11, total));
    CK(cudaMalloc(&x, 4 * 65536 * 4)); CK(cudaMalloc(&nw, 65536 * 4)); CK(cudaMalloc(&y, (size_t)4 * V * 4)); CK(cudaMalloc(&act, 4 * 65536 * 4));
    CK(cudaMemset(x, 0, 4 * 65536 * 4)); CK(cudaMemset(nw, 0, 65536 * 4)); CK(cudaMemset(act, 0, 4 * 65536 * 4));
    CK(cudaMalloc(&xqb, xquant_bytes(4, 12288)));
    float* pv; int* pi; unsigned int* tk; SeqState* ss;
    CK(cudaMalloc(&pv, 16384)); CK(cudaMalloc(&pi, 16384)); CK(cudaMalloc(&tk, 4)); CK(cudaMemset(tk, 0, 4)); CK(cudaMalloc(&ss, 4 * sizeof(SeqState))); CK(cudaMemset(ss, 0, 4 * sizeof(SeqState)));
    ex = GemvArgs{}; ex.part_val = pv; ex.part_idx = pi; ex.ticket = tk; ex.state = ss; ex.out_tokens = nullptr; ex.out_stride = 0; ex.embed = nullptr; ex.x_next = x; ex.H = H; ex.advance = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("QG_R=%d QG_DEPTH=%d warps(B=1)=%d warps(B=4)=%d\n", QG_R, QG_DEPTH, qg_warps(1), qg_warps(4));
    Shape all[] = {q, k, v, o, gu, dn, head};
    for (int B : {1, 4}) {
        for (auto& s : all) {
            const size_t wsz = wbytes(s);
            const int copies = (int)std::max<size_t>(1, std::min<size_t>(total / wsz, 64));
            const int iters = 100;
            for (int it = 0; it < 10; ++it) { int r = run(st, B, s, W + (size_t)(it % copies) * wsz, false); if (r) { printf("launch %s B=%d failed %d\n", s.name, B, r); return 2; } }
            CK(cudaStreamSynchronize(st));
            cudaEventRecord(e0, st);
            for (int it = 0; it < iters; ++it) run(st, B, s, W + (size_t)(it % copies) * wsz, false);
            cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1e3 / iters;
            printf("  B=%d isolated %-8s N=%6d K=%5d : %7.2f us  %7.1f GB/s\n", B, s.name, s.N, s.K, us, wsz / us / 1e3);
        }
        for (int pdl = 0; pdl <= 1; ++pdl) {
            auto step = [&]() {
                const unsigned char* w = W;
                for (int l = 0; l < L; ++l)
                    for (auto& s : layer) { run(st, B, s, w, pdl); w += wbytes(s); }
                run(st, B, head, w, pdl);
            };
            for (int i = 0; i < 2; ++i) step();
            CK(cudaStreamSynchronize(st));
            cudaEventRecord(e0, st);
            const int iters = 10;
            for (int i = 0; i < iters; ++i) step();
            cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1e3 / iters;
            printf("  B=%d chained step (%d GEMVs, no attention) pdl=%d : %8.1f us  %7.1f GB/s\n", B, L * 6 + 1, pdl, us, total / us / 1e3);
        }
    }
    return 0;
}
