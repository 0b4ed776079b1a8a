// crane_b200 engine: owns weights / KV pages / workspaces for one model on one GPU and orchestrates the
// kernels of gemm.cu, decode.cu and prefill.cu behind the C ABI of include/crane_b200.h.
//
// Host-side structure mirrors the reference (crane-core/src/models/):
//   Qwen3Model::{forward, forward_embeds, decode}      qwen3/modeling.rs:942-1036
//   Model::{forward_step, generate, clear_kv_cache}    qwen3/model.rs:34-349
//   Qwen3_5VLModel::{encode_images, forward, decode_step, build_position_ids}   qwen3_5/vlm.rs:150-301
//   Qwen3_5VisionModel::forward                         qwen3_5/vision.rs:558-584
#include "../../include/crane_b200.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "decode.cuh"
#include "decode_ll.cuh"
#include "gdn.cuh"
#include "gemm.cuh"
#include "json_min.h"
#include "prefill.cuh"
#include "quant.cuh"
#include "sampler.cuh"
#include "prof.h"

using namespace cb;

namespace {

struct EngineError {
    int code;
    std::string msg;
};
[[noreturn]] void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw EngineError{code, buf};
}
#define CUDA_OK(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            fail(e__ == cudaErrorMemoryAllocation ? CRANE_B200_OOM : CRANE_B200_CUDA_ERROR, "%s: %s (%s:%d)", \
                 #expr, cudaGetErrorString(e__), __FILE__, __LINE__);                                         \
    } while (0)
#define LAUNCH_OK(expr)                                                                                       \
    do {                                                                                                      \
        int r__ = (expr);                                                                                     \
        if (r__ == -1000) fail(CRANE_B200_UNSUPPORTED, "unsupported shape in %s (%s:%d)", #expr, __FILE__, __LINE__); \
        if (r__ != 0) fail(CRANE_B200_CUDA_ERROR, "%s failed: %d %s (%s:%d)", #expr, r__,                     \
                           r__ > 0 ? cudaGetErrorString((cudaError_t)r__) : "", __FILE__, __LINE__);          \
    } while (0)

std::string g_create_error;

// ---- host dtype conversion ---------------------------------------------------------------------
inline uint16_t f32_to_bf16_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf16_bits_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, u;
    if (exp == 0) {
        if (man == 0) u = sign;
        else {
            exp = 127 - 15 + 1;
            while (!(man & 0x400u)) { man <<= 1; --exp; }
            u = sign | (exp << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
    else u = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

void to_bf16(const void* src, int dt, size_t n, std::vector<uint16_t>& out) {
    out.resize(n);
    if (dt == CRANE_B200_BF16) std::memcpy(out.data(), src, n * 2);
    else if (dt == CRANE_B200_F32) { const float* s = (const float*)src; for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16_bits(s[i]); }
    else { const uint16_t* s = (const uint16_t*)src; for (size_t i = 0; i < n; ++i) out[i] = f32_to_bf16_bits(f16_bits_to_f32(s[i])); }
}
void to_f32(const void* src, int dt, size_t n, std::vector<float>& out) {
    out.resize(n);
    if (dt == CRANE_B200_F32) std::memcpy(out.data(), src, n * 4);
    else if (dt == CRANE_B200_BF16) { const uint16_t* s = (const uint16_t*)src; for (size_t i = 0; i < n; ++i) out[i] = bf16_bits_to_f32(s[i]); }
    else { const uint16_t* s = (const uint16_t*)src; for (size_t i = 0; i < n; ++i) out[i] = f16_bits_to_f32(s[i]); }
}

// ggml block formats -> f32 on the host (dequantize_row_q8_0 / q4_K / q6_K of ggml-quants.c; restated in oracle/ggml_quant.py, which
// is byte-exact against the `gguf` package).  Used where quantised weights are not run as such: the Qwen3.5 hybrid's GGUF tensors.
void ggml_dequant_rows(int qt, const unsigned char* src, size_t rows, size_t K, std::vector<float>& out) {
    out.resize(rows * K);
    auto f16 = [](const unsigned char* p) { uint16_t h; std::memcpy(&h, p, 2); return f16_bits_to_f32(h); };
    float* y = out.data();
    if (qt == 8) {
        for (size_t b = 0; b < rows * K / 32; ++b, src += 34, y += 32) {
            const float d = f16(src);
            for (int i = 0; i < 32; ++i) y[i] = d * (float)(int8_t)src[2 + i];
        }
    } else if (qt == 12) {
        for (size_t b = 0; b < rows * K / 256; ++b, src += 144, y += 256) {
            const float d = f16(src), dmin = f16(src + 2);
            const unsigned char *sc = src + 4, *q = src + 16;
            auto scale_min = [&](int j, int& s_, int& m_) {
                if (j < 4) { s_ = sc[j] & 63; m_ = sc[j + 4] & 63; }
                else { s_ = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m_ = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
            };
            for (int j = 0; j < 4; ++j) {
                int s1, m1, s2, m2;
                scale_min(2 * j, s1, m1); scale_min(2 * j + 1, s2, m2);
                const float d1 = d * s1, n1 = dmin * m1, d2 = d * s2, n2 = dmin * m2;
                for (int l = 0; l < 32; ++l) { y[64 * j + l] = d1 * (q[32 * j + l] & 0xF) - n1; y[64 * j + 32 + l] = d2 * (q[32 * j + l] >> 4) - n2; }
            }
        }
    } else {
        for (size_t b = 0; b < rows * K / 256; ++b, src += 210, y += 256) {
            const unsigned char *ql = src, *qh = src + 128;
            const int8_t* sc = (const int8_t*)(src + 192);
            const float d = f16(src + 208);
            for (int n = 0; n < 2; ++n, ql += 64, qh += 32, sc += 8)
                for (int l = 0; l < 32; ++l) {
                    const int is = l / 16;
                    const int q1 = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32, q2 = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int q3 = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32, q4 = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    float* yy = y + 128 * n;
                    yy[l] = d * sc[is] * q1; yy[l + 32] = d * sc[is + 2] * q2; yy[l + 64] = d * sc[is + 4] * q3; yy[l + 96] = d * sc[is + 6] * q4;
                }
        }
    }
}

struct LayerW {
    bf16 *wqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wdown = nullptr;
    float *ln1 = nullptr, *ln2 = nullptr, *qn = nullptr, *kn = nullptr;
    bf16 *k_pool = nullptr, *v_pool = nullptr;
    unsigned char *k_codes = nullptr, *v_codes = nullptr;   // engine.kv_cache = int8 / int4: quantised pages (codes + per-token scales)
    float *k_scale = nullptr, *v_scale = nullptr;
    int loaded = 0;   // bit per tensor
    // GGUF-quantised linears (ggml bytes repacked by q_repack_rows): qt_* == 0 means the bf16 buffer above is used
    unsigned char *q_wq = nullptr, *q_wk = nullptr, *q_wv = nullptr, *q_wo = nullptr, *q_wgu = nullptr, *q_wdown = nullptr;
    int qt_q = 0, qt_k = 0, qt_v = 0, qt_o = 0, qt_gu = 0, qt_down = 0, qt_gate_seen = 0, qt_up_seen = 0;
    // Gated-Delta-Net layers (Qwen3.5 `linear_attn.*`)
    bool full = true;                       // softmax-attention layer (else GDN)
    bf16 *w_in = nullptr, *w_out = nullptr; // [in_pad, H] rows = q|k|v|z|b|a ; [H, value_dim]
    float *conv_w = nullptr, *neg_exp_a = nullptr, *dt_bias = nullptr, *gnorm = nullptr;
    float *conv_state = nullptr, *rec_state = nullptr;   // working state of the CURRENT sequence slot
    float *conv_slots = nullptr, *rec_slots = nullptr;   // [max_batch] parked states (max_batch > 1): swapped in by select_seq
};
struct VitBlockW {
    float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;
    bf16* wqkv = nullptr; float* bqkv = nullptr;
    bf16* wproj = nullptr; float* bproj = nullptr;
    bf16* wfc1 = nullptr; float* bfc1 = nullptr;
    bf16* wfc2 = nullptr; float* bfc2 = nullptr;
    int loaded = 0;
};
struct MergerW {
    float *nw = nullptr, *nb = nullptr;
    bf16* w1 = nullptr; float* b1 = nullptr;
    bf16* w2 = nullptr; float* b2 = nullptr;
    bool post = false;
    int loaded = 0;
};

}  // namespace

#include "engine_sampler.inc"

struct crane_b200_model {
    // ---- configuration ----
    int device = 0, num_sms = 148;
    bool is_vl = false;
    int V = 0, H = 0, I = 0, L = 0, nh = 0, nkv = 0, D = 0;
    float eps = 1e-6f;
    double theta = 1e6;
    bool tied = true;
    std::vector<int> mrope_section;
    // Qwen3.5 hybrid
    bool hybrid = false;
    int nk = 0, nv = 0, dk = 0, dv = 0, ck = 4, gdn_in = 0, gdn_in_pad = 0, rot_half = 0, full_interval = 4;
    std::vector<int> layer_is_full;
    int max_seq = 4096, max_batch = 1, max_pages = 0;
    bool use_simt = false, use_graphs = true, use_pdl = true, use_persistent = true;
    bool tensor_gguf_named = false;         // the tensor being loaded came under its llama.cpp name (Qwen3.5: folded norms, Chunked v-heads)
    struct crane_b200_comm* comm = nullptr;   // multi-GPU: NCCL communicator + gather buffers (engine_comm.inc)
    PassProfiler spans;                     // CRANE_PROF=1 / crane_b200_prof_enable: enqueue vs wall per pass + stage spans (ops/prof.rs)
    // persistent decode kernel (decode_ll.cu): exchange buffers of (value, tag) pairs, the tag counter, the timeout flag
    unsigned long long *ll_xa = nullptr, *ll_xb = nullptr, *ll_qkv = nullptr, *ll_att = nullptr, *ll_act = nullptr, *ll_part = nullptr, *ll_amax = nullptr;
    unsigned int ll_tag = 0;
    unsigned int* ll_err = nullptr;
    unsigned int* h_ll_err = nullptr;      // pinned
    unsigned long long* ll_prof = nullptr;
    unsigned long long* ll_trace = nullptr;   // CRANE_B200_LL_TRACE=<file>: event trace of one step (tools/ll_trace.py)
    void ll_err_fetch();                   // enqueue the D2H copy of the flag (before a stream sync the caller does anyway)
    void ll_err_verify();                  // after that sync
    // vision
    int v_depth = 0, v_H = 0, v_I = 0, v_nh = 0, v_hd = 0, v_patch = 16, v_merge = 2, v_tpatch = 2, v_in = 3, v_out = 0,
        v_npos = 0, v_side = 0;
    std::vector<int> v_deepstack;
    int vit_gelu_mode = EPI_GELU_ERF_BF16, merger_gelu_mode = EPI_GELU_TANH_BF16;
    uint32_t image_token_id = 0;

    // ---- device state ----
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    bf16 *embed = nullptr, *lm_head = nullptr;
    unsigned char* q_lm_head = nullptr;   // quantised output head (GGUF `output.weight`, or the tied quantised `token_embd.weight`)
    int qt_lm = 0;
    unsigned char* q_embed = nullptr;     // quantised embedding table: rows are dequantised as they are gathered (modules/embedding.rs:31-105)
    int qt_embed = 0;
    void embed_step_input(int B);         // x_dec <- embedding rows of state[b].token
    bool any_quant = false;
    float* final_norm = nullptr;
    std::vector<LayerW> layers;
    bool got_embed = false, got_lm_head = false, got_final_norm = false, finalized = false;
    float *cos_tab = nullptr, *sin_tab = nullptr;
    unsigned char* axis_of = nullptr;
    // vision weights
    bf16* v_wpatch = nullptr; float* v_bpatch = nullptr; float* v_pos = nullptr;
    std::vector<VitBlockW> vblocks;
    MergerW v_merger;
    std::vector<MergerW> v_ds_mergers;
    int v_loaded = 0;

    // decode buffers
    float *x_dec = nullptr, *qkv_dec = nullptr, *attn_dec = nullptr, *act_dec = nullptr, *logits = nullptr;
    float *part_o = nullptr, *part_ml = nullptr, *part_val = nullptr;
    int* part_idx = nullptr;
    unsigned int *counters = nullptr, *ticket = nullptr;
    SeqState* state = nullptr;
    uint32_t* out_tokens = nullptr;
    int out_cap = 8192;
    // multiple output heads / extra embedding tables (Qwen3-TTS code predictor: lm_head.{g}, codec_embedding.{g})
    std::vector<bf16*> heads;          // heads[g]; the single-head models only use lm_head
    std::vector<bf16*> emb_tables;     // emb_tables[g]: [table_rows, H]
    int table_rows = 0;
    float* hidden_out = nullptr;       // [B, H] post-final-norm hidden state of the last pass (norm_out of the head GEMV)
    struct HeadOverride {              // per-pass head / gather / forcing selection used by lm_head_last_row
        const bf16* head = nullptr; const bf16* gather = nullptr; bool gather_set = false;
        const uint32_t* force = nullptr; float* logits = nullptr;
    } hov;
    bool owns_stream = true;
    // Qwen3-TTS composite: this model is the talker, `cp` the code predictor
    bool is_tts = false;
    crane_b200_model* cp = nullptr;
    int tts_groups = 0, tts_Ht = 0, tts_Vt = 0, tts_eos = 0;
    bf16 *tts_text_emb = nullptr, *tts_fc1 = nullptr, *tts_fc2 = nullptr;
    float *tts_b1 = nullptr, *tts_b2 = nullptr;
    int tts_loaded = 0;
    float* tts_contrib = nullptr;      // [n_trailing + 1, H] trailing-text rows then the tts_pad row
    int tts_n_trailing = 0, tts_max_frames = 0, tts_prefill_len = 0, tts_frames_done = 0;
    uint32_t *tts_frames = nullptr, *tts_force = nullptr;   // [max_frames, groups]
    int* tts_ctrl = nullptr;           // [4]: step, done_step (-1 while running), -, -
    unsigned char* tts_seen = nullptr; // [V] first codes generated so far (repetition penalty is applied once per distinct token)
    float *tts_first_logits = nullptr; // [V] penalised / masked copy used by the selector
    std::string cp_json;
    void tts_setup();
    void tts_load(const std::string& name, int dt, const int64_t* shape, int ndim, const void* data);
    void tts_text_project(const uint32_t* ids, size_t n, float* out_host);
    void tts_prefill(const float* embeds, size_t P, const float* trailing, size_t nt, const float* pad);
    void tts_frame(float rep_penalty, bool forced, int frame);
    const bf16** tts_tables_dev = nullptr;
    int* block_table = nullptr;
    // Sequence state travels host -> device through a ring of pinned snapshots, one per call: the ABI hands out stream-ordered
    // results, so a caller may issue the next call before the previous upload has run, and a single staging slot would be
    // overwritten under it.  A slot is reused only after its own upload has completed (its event).
    static constexpr int STATE_RING = 16, STATE_GROUP = 4;
    SeqState* h_state_ring = nullptr; // pinned [STATE_RING][STATE_GROUP]
    cudaEvent_t h_state_ev[STATE_RING] = {};
    bool h_state_used[STATE_RING] = {};
    int h_state_next = 0;
    SeqState* h_state = nullptr;      // the slot being filled (stage_state)
    SeqState* stage_state() {
        const int i = h_state_next;
        if (h_state_used[i]) CUDA_OK(cudaEventSynchronize(h_state_ev[i]));
        h_state = h_state_ring + (size_t)i * STATE_GROUP;
        return h_state;
    }
    void push_state(int nseq) {       // upload the staged snapshot(s) to the device-resident state
        CUDA_OK(cudaMemcpyAsync(state, h_state, sizeof(SeqState) * nseq, cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaEventRecord(h_state_ev[h_state_next], stream));
        h_state_used[h_state_next] = true;
        h_state_next = (h_state_next + 1) % STATE_RING;
    }
    uint32_t* h_tokens = nullptr;     // pinned
    // prefill workspaces (grown on demand)
    int ws_S = 0;
    float *x = nullptr, *qkv = nullptr;
    bf16 *xn = nullptr, *q_bf = nullptr, *attn_bf = nullptr, *act_bf = nullptr;
    float* rows_f32 = nullptr;         // [S, max(I, q_dim)] f32 rows in front of / behind a quantised linear in prefill
    float *g_proj = nullptr, *g_conv = nullptr, *g_qn = nullptr, *g_kn = nullptr, *g_gb = nullptr, *g_y = nullptr;   // GDN prefill workspaces
    float* g_gl = nullptr;                 // [S, nv] log decays for the chunkwise recurrence
    unsigned char* g_chunk = nullptr;      // its scratch (gdn_chunk_ws_bytes)
    int gdn_mode = 0;                      // engine.gdn: 0 = auto (chunkwise from 64 rows per call), 1 = sequential, 2 = chunked
    float *gd_proj = nullptr, *gd_conv = nullptr, *gd_qn = nullptr, *gd_kn = nullptr, *gd_gb = nullptr, *gd_y = nullptr, *gd_out = nullptr;  // decode
    uint32_t* ids_dev = nullptr;
    int* h_stage = nullptr;                 // pinned staging of a prefill's positions / ids / splice rows (5 x capacity ints)
    cudaEvent_t stage_ev = nullptr;         // ... free again once this event has passed
    cudaEvent_t pix_ev = nullptr;           // the caller's pixel buffer has been read (vl_forward returns no earlier)
    bool stage_busy = false;
    int* pos3_dev = nullptr;
    int* rows_dev = nullptr;
    float* embeds_in = nullptr;
    // vision workspaces
    int vws_N = 0;
    float *v_x = nullptr, *v_qkv = nullptr, *v_pv = nullptr, *v_cos = nullptr, *v_sin = nullptr, *v_w4 = nullptr;
    bf16 *v_pvb = nullptr, *v_xn = nullptr, *v_qkvb = nullptr, *v_attn = nullptr, *v_act = nullptr, *v_m1 = nullptr;
    int *v_idx4 = nullptr, *v_seq_start = nullptr, *v_seq_len = nullptr;
    float* img_embeds = nullptr;      // [n_img_tok, v_out]
    float* ds_embeds = nullptr;       // [n_ds, n_img_tok, v_out]
    int img_tokens = 0;
    std::vector<uint32_t> v_tab_key;        // grids the device-resident ViT index / rotary tables were built for
    int v_tab_nseq = 0, v_tab_max_len = 0;

    // host state: `kv_len` / `next_mrope_pos` are the view of the CURRENT sequence slot (`cur`); the others are parked in seq_*
    size_t kv_len = 0;
    uint32_t next_mrope_pos = 0;
    int cur = 0;
    std::vector<size_t> seq_kv;
    std::vector<uint32_t> seq_pos;
    std::vector<char> seq_used;
    void select_seq(int s);               // parks the current slot's view (and GDN state) and brings slot s in
    void park_gdn_state(int slot);         // working recurrent / conv state -> the slot's store (hybrid models)
    void fork_seq(int src, int dst);
    const int* bt_cur() const { return block_table + (size_t)cur * max_pages; }
    cudaGraphExec_t graph_step[2][5] = {};   // [advance][sequences in the group: 1, 2, 4]
    bool graph_failed = false;
    cudaEvent_t pev0 = nullptr, pev1 = nullptr, dev0 = nullptr, dev1 = nullptr;
    bool pev0_armed = false;
    float last_prefill_ms = 0.f, last_decode_ms = 0.f;
    size_t last_decode_steps = 0;
    uint64_t launches = 0;
    uint64_t graph_launches[2][5] = {};
    unsigned char* xq_buf = nullptr;   // activations quantised for the quantised GEMVs (xquant_launch)
    int xq_mode_last = -1;
    // Split precision (default): every bf16 activation operand / KV page has a low-order plane (x = hi + lo) so the prefill
    // tensor-core path carries ~16 mantissa bits; lo_* = element offset of that plane from the buffer base, 0 when off.
    bool split = true;
    long long lo_xn = 0, lo_q = 0, lo_attn = 0, lo_act = 0, lo_kv = 0;
    int kv_bits = 0;                        // 0: bf16 (+lo) pages; 8 / 4: QuantKvCache pages (qwen3_5/kv_cache.rs:209-342)
    bf16 *kq_sk = nullptr, *kq_sv = nullptr; // quantised mode: one layer's K / V of one sequence, dequantised, page layout (prefill scratch)
    int* bt_identity = nullptr;             // block table of that scratch: page p at p
    KvQuantArgs kvq_args(LayerW& l) {
        KvQuantArgs a = {};
        a.k_codes = l.k_codes; a.v_codes = l.v_codes; a.k_scale = l.k_scale; a.v_scale = l.v_scale; a.bt = bt_cur();
        a.nkv = nkv; a.D = D; a.bits = kv_bits; a.sk = kq_sk; a.sv = kq_sv; a.s_lo = lo_kv;
        return a;
    }
    long long lo_vpvb = 0, lo_vxn = 0, lo_vqkvb = 0, lo_vattn = 0, lo_vact = 0, lo_vm1 = 0;
    bf16* dalloc_act(size_t n, long long& lo_off) {     // bf16 activation buffer (+ its lo plane)
        lo_off = split ? (long long)n : 0;
        return dalloc<bf16>(split ? 2 * n : n);
    }
    SamplerScratch sampler;
    std::string last_error;

    // ---------------------------------------------------------------------------------------------
    template <typename T>
    T* dalloc(size_t n) {
        void* p = nullptr;
        CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        allocs.push_back(p);
        return (T*)p;
    }
    void dfree(void* p) {
        if (!p) return;
        auto it = std::find(allocs.begin(), allocs.end(), p);
        if (it != allocs.end()) allocs.erase(it);
        cudaFree(p);
    }
    int q_stride() const { return hybrid ? 2 * D : D; }     // Qwen3.5 q_proj emits per-head [query | gate]
    int qkv_dim() const { return nh * q_stride() + 2 * nkv * D; }
    int conv_dim() const { return 2 * nk * dk + nv * dv; }
    int value_dim() const { return nv * dv; }
    int q_dim() const { return nh * D; }

    void parse_config(const char* json);
    void alloc_weights();
    void load_tensor(const std::string& name, int dt, const int64_t* shape, int ndim, const void* data);
    void load_tensor_ggml(const std::string& name, int qt, const int64_t* shape, int ndim, const void* data, size_t nbytes);
    void load_safetensors_file(const char* path, size_t* n_loaded, size_t* n_skipped);
    void load_gguf_file(const char* path, size_t* n_loaded, size_t* n_skipped);
    unsigned char* up_quant(int qt, const void* data, size_t rows, int K, unsigned char* dst = nullptr, size_t dst_row_pitch = 0);
    void linear_decode(int epi, bool norm, const bf16* w, const unsigned char* qw, int qt, int N, int K, const float* xin, int ldx,
                       const float* norm_w, float* y, int ldy, const GemvArgs* extra = nullptr, int B = 1, bool reuse_xq = false);
    void qlinear_rows(int epi, const unsigned char* qw, int qt, int N, int K, const float* xin, int ldx, const float* norm_w, float* y, int ldy,
                      int S, bool reuse_xq = false);
    bool load_text_tensor(const std::string& n, int dt, const int64_t* shape, int ndim, const void* data);
    bool load_vision_tensor(const std::string& n, int dt, const int64_t* shape, int ndim, const void* data);
    void finalize();
    void ensure_prefill_ws(int S);
    void ensure_vision_ws(int N);

    void up_bf16(bf16* dst, const void* data, int dt, size_t n) {
        std::vector<uint16_t> tmp;
        to_bf16(data, dt, n, tmp);
        CUDA_OK(cudaMemcpy(dst, tmp.data(), n * 2, cudaMemcpyHostToDevice));
    }
    void up_f32(float* dst, const void* data, int dt, size_t n) {
        std::vector<float> tmp;
        to_f32(data, dt, n, tmp);
        CUDA_OK(cudaMemcpy(dst, tmp.data(), n * 4, cudaMemcpyHostToDevice));
    }
    // RMSNorm weights: Qwen3.5 applies (1 + w); folded once at load like the reference (qwen3_5/modeling.rs:45-79)
    void up_norm(float* dst, const void* data, int dt, size_t n) {
        std::vector<float> tmp;
        to_f32(data, dt, n, tmp);
        if (hybrid && !tensor_gguf_named) for (auto& v : tmp) v += 1.0f;   // (llama.cpp's converter stores the +1 folded in already)
        CUDA_OK(cudaMemcpy(dst, tmp.data(), n * 4, cudaMemcpyHostToDevice));
    }

    // forward paths
    void arm_state(uint32_t token, size_t start_pos, int p0, int p1, int p2);
    void enqueue_decode_step(int advance, bool with_embed, int B = 1);
    void decode_step_graphed(int advance, int B = 1);
    void decode_steps(int n_steps, int advance);
    void prefill(const uint32_t* ids, const float* embeds, size_t S, const uint32_t* pos3_host, size_t start_pos,
                 const int* vis_rows, int n_vis, int advance);
    void lm_head_last_row(const float* xrow, int advance, int B = 1);
    void gdn_args(GdnArgs& g, const LayerW& l, int S, const float* proj, float* conv, float* qn, float* kn, float* gb, float* y) const;
    void reset_recurrent_state();
    void encode_images(const float* pv, const uint32_t* grid, size_t n_images);
    // a_lo / out_lo: element offsets of the low-order planes of A and of a bf16 `out` (0 = plain bf16)
    void gemm(const bf16* A, long long a_lo, int lda, const bf16* W, int M, int N, int K, int mode, void* out, int ldo, const float* bias,
              long long out_lo = 0) {
        GemmEpi ep{out, out_lo ? (void*)((bf16*)out + out_lo) : nullptr, ldo, bias, mode};
        LAUNCH_OK(gemm_bf16_launch(stream, A, a_lo ? A + a_lo : nullptr, lda, W, M, N, K, ep, use_simt));
        ++launches;
    }
};

// =================================================================================================
// configuration
// =================================================================================================
void crane_b200_model::parse_config(const char* json) {
    cbjson::Value root = cbjson::Parser(json).parse();
    if (root.type != cbjson::Value::OBJ) fail(CRANE_B200_INVALID_ARG, "config must be a JSON object");
    is_tts = root.has("talker_config");
    const cbjson::Value& tc = root.has("text_config") ? root.at("text_config") : is_tts ? root.at("talker_config") : root;
    is_vl = root.has("vision_config");
    V = (int)tc.integer("vocab_size");
    H = (int)tc.integer("hidden_size");
    I = (int)tc.integer("intermediate_size");
    L = (int)tc.integer("num_hidden_layers");
    nh = (int)tc.integer("num_attention_heads");
    nkv = (int)tc.integer("num_key_value_heads");
    D = (int)tc.integer("head_dim", H / nh);
    eps = (float)tc.number("rms_norm_eps", 1e-6);
    theta = tc.number("rope_theta", 1000000.0);
    if (tc.has("rope_parameters") && tc.at("rope_parameters").has("rope_theta")) theta = tc.at("rope_parameters").number("rope_theta", theta);
    tied = root.boolean("tie_word_embeddings", tc.boolean("tie_word_embeddings", true));
    if (is_tts) {
        tied = false;                    // codec_embedding and codec_head are separate tensors
        tts_groups = (int)tc.integer("num_code_groups", 16);
        tts_Ht = (int)tc.integer("text_hidden_size", 2048);
        tts_Vt = (int)tc.integer("text_vocab_size", 151936);
        tts_eos = (int)tc.integer("codec_eos_token_id", 0);
        if (tts_Ht % 64 || tts_Ht % 32) fail(CRANE_B200_UNSUPPORTED, "text_hidden_size %d", tts_Ht);
        const cbjson::Value& cc = tc.at("code_predictor_config");
        if (cc.integer("hidden_size") != tc.integer("hidden_size"))
            fail(CRANE_B200_UNSUPPORTED, "code predictor hidden_size != talker hidden_size (small_to_mtp_projection) is not supported");
        char buf[1024];
        snprintf(buf, sizeof buf,
                 "{\"vocab_size\": %lld, \"hidden_size\": %lld, \"intermediate_size\": %lld, \"num_hidden_layers\": %lld, "
                 "\"num_attention_heads\": %lld, \"num_key_value_heads\": %lld, \"head_dim\": %lld, \"rms_norm_eps\": %g, "
                 "\"rope_theta\": %.1f, \"tie_word_embeddings\": false, \"engine\": {\"max_seq_len\": 64, \"persistent\": false}}",
                 cc.integer("vocab_size", 2048), cc.integer("hidden_size"), cc.integer("intermediate_size"), cc.integer("num_hidden_layers"),
                 cc.integer("num_attention_heads"), cc.integer("num_key_value_heads"), cc.integer("head_dim", 128), cc.number("rms_norm_eps", 1e-6),
                 cc.number("rope_theta", 1000000.0));
        cp_json = buf;
        if ((int)cc.integer("num_code_groups", tts_groups) != tts_groups) fail(CRANE_B200_INVALID_ARG, "num_code_groups differs between talker and code predictor");
    }
    const cbjson::Value* rs = tc.has("rope_scaling") ? &tc.at("rope_scaling") : (tc.has("rope_parameters") ? &tc.at("rope_parameters") : nullptr);
    if (rs && rs->has("mrope_section"))
        for (const auto& v : rs->at("mrope_section").arr) mrope_section.push_back((int)v.num);
    hybrid = tc.has("linear_num_value_heads");
    rot_half = D / 2;
    if (hybrid) {
        nk = (int)tc.integer("linear_num_key_heads");
        nv = (int)tc.integer("linear_num_value_heads");
        dk = (int)tc.integer("linear_key_head_dim", 128);
        dv = (int)tc.integer("linear_value_head_dim", 128);
        ck = (int)tc.integer("linear_conv_kernel_dim", 4);
        full_interval = (int)tc.integer("full_attention_interval", 4);
        double prf = tc.number("partial_rotary_factor", 0.25);
        if (rs && rs->has("partial_rotary_factor")) prf = rs->number("partial_rotary_factor", prf);
        rot_half = (int)(D * prf) / 2;
        gdn_in = conv_dim() + value_dim() + 2 * nv;
        gdn_in_pad = (gdn_in + 31) / 32 * 32;
        // key widths: the reference's kernel takes K <= 256 (kernels/cuda/gdn.cu:45-153); 64 / 128 / 256 are instantiated here
        if ((dk != 64 && dk != 128 && dk != 256) || dv % 32 || dv <= 0 || nk <= 0 || nv % nk || nv > 256 || ck < 1 || ck > 8)
            fail(CRANE_B200_UNSUPPORTED, "Gated-Delta-Net geometry key_head_dim=%d (64, 128 or 256) value_head_dim=%d (multiple of 32) key_heads=%d value_heads=%d (multiple of key_heads, <= 256) conv_kernel=%d (1..8)", dk, dv, nk, nv, ck);
        if (rot_half < 32 || rot_half % 32) fail(CRANE_B200_UNSUPPORTED, "rotary width %d", 2 * rot_half);
    }
    layer_is_full.assign(L, 1);
    if (hybrid) {
        for (int i = 0; i < L; ++i) layer_is_full[i] = ((i + 1) % full_interval == 0) ? 1 : 0;
        if (tc.has("layer_types")) {
            const auto& lt = tc.at("layer_types").arr;
            for (int i = 0; i < L && i < (int)lt.size(); ++i) layer_is_full[i] = lt[i].str == "full_attention";
        }
    }
    if (tc.boolean("attention_bias", false)) fail(CRANE_B200_UNSUPPORTED, "attention_bias=true is not supported");
    if (!tc.boolean("use_qk_norm", true)) fail(CRANE_B200_UNSUPPORTED, "use_qk_norm=false is not supported");
    if (D != 128 && D != 256) fail(CRANE_B200_UNSUPPORTED, "head_dim %d (supported: 128, 256)", D);
    if (nh % nkv) fail(CRANE_B200_INVALID_ARG, "num_attention_heads %% num_key_value_heads != 0");
    if (H % 256 || I % 256) fail(CRANE_B200_UNSUPPORTED, "hidden/intermediate sizes must be multiples of 256");
    if (root.has("engine")) {
        const cbjson::Value& e = root.at("engine");
        max_seq = (int)e.integer("max_seq_len", max_seq);
        max_batch = (int)e.integer("max_batch", max_batch);
        use_simt = e.string("gemm", "tcgen05") == "simt";
        use_graphs = e.boolean("graphs", true);
        use_pdl = e.boolean("pdl", true);
        const std::string prec = e.string("precision", "split");
        if (prec == "bf16") split = false;
        else if (prec != "split") fail(CRANE_B200_INVALID_ARG, "engine.precision must be \"split\" or \"bf16\"");
        const std::string kvc = e.string("kv_cache", "fp");
        if (kvc == "int8") kv_bits = 8;
        else if (kvc == "int4") kv_bits = 4;
        else if (kvc != "fp") fail(CRANE_B200_INVALID_ARG, "engine.kv_cache must be \"fp\", \"int8\" or \"int4\"");
        if (e.string("vit_act", "erf") == "tanh") vit_gelu_mode = EPI_GELU_TANH_BF16;
        const std::string gm = e.string("gdn", "auto");
        if (gm == "sequential") gdn_mode = 1;
        else if (gm == "chunked") gdn_mode = 2;
        else if (gm != "auto") fail(CRANE_B200_INVALID_ARG, "engine.gdn must be \"auto\", \"chunked\" or \"sequential\"");
        if (e.string("merger_act", "tanh") == "erf") merger_gelu_mode = EPI_GELU_ERF_BF16;
    }
    if (const char* g = getenv("CRANE_B200_GEMM")) use_simt = std::string(g) == "simt";
    if (const char* g = getenv("CRANE_B200_GRAPHS")) use_graphs = std::string(g) != "0";
    if (const char* g = getenv("CRANE_B200_PDL")) use_pdl = std::string(g) != "0";
    if (const char* g = getenv("CRANE_B200_GDN")) gdn_mode = std::string(g) == "sequential" ? 1 : std::string(g) == "chunked" ? 2 : 0;
    cb::prefill_pdl() = use_pdl;          // process-wide: the prefill-side launchers read it
    if (const char* g = getenv("CRANE_B200_PRECISION")) split = std::string(g) != "bf16";
    if (root.has("engine")) use_persistent = root.at("engine").boolean("persistent", true);
    if (const char* g = getenv("CRANE_B200_PERSISTENT")) use_persistent = std::string(g) != "0";
    spans.init_from_env();
    if (max_batch < 1 || max_batch > 64) fail(CRANE_B200_INVALID_ARG, "max_batch %d (1..64 sequence slots)", max_batch);
    max_seq = (max_seq + KV_PAGE - 1) / KV_PAGE * KV_PAGE;
    max_pages = max_seq / KV_PAGE;
    if (is_vl) {
        const cbjson::Value& vc = root.at("vision_config");
        v_depth = (int)vc.integer("depth");
        v_H = (int)vc.integer("hidden_size");
        v_I = (int)vc.integer("intermediate_size");
        v_nh = (int)vc.integer("num_heads");
        v_hd = v_H / v_nh;
        v_patch = (int)vc.integer("patch_size", 16);
        v_merge = (int)vc.integer("spatial_merge_size", 2);
        v_tpatch = (int)vc.integer("temporal_patch_size", 2);
        v_in = (int)vc.integer("in_channels", 3);
        v_out = (int)vc.integer("out_hidden_size");
        v_npos = (int)vc.integer("num_position_embeddings", 2304);
        v_side = (int)std::lround(std::sqrt((double)v_npos));
        if (v_side * v_side != v_npos) fail(CRANE_B200_INVALID_ARG, "num_position_embeddings %d is not a perfect square", v_npos);
        if (vc.has("deepstack_visual_indexes"))
            for (const auto& v : vc.at("deepstack_visual_indexes").arr) v_deepstack.push_back((int)v.num);
        if (v_out != H) fail(CRANE_B200_INVALID_ARG, "vision out_hidden_size %d != text hidden_size %d", v_out, H);
        if (v_hd != 64 && v_hd != 128) fail(CRANE_B200_UNSUPPORTED, "vision head_dim %d", v_hd);
        image_token_id = (uint32_t)root.integer("image_token_id", 151655);
        if ((int)v_deepstack.size() > L) fail(CRANE_B200_INVALID_ARG, "more deepstack levels than decoder layers");
    }
}

void crane_b200_model::alloc_weights() {
    embed = dalloc<bf16>((size_t)V * H);
    lm_head = tied ? embed : nullptr;      // untied head: allocated when `lm_head.weight` arrives (bf16 or quantised)
    final_norm = dalloc<float>(H);
    layers.resize(L);
    for (int li = 0; li < L; ++li) {
        LayerW& l = layers[li];
        l.full = layer_is_full[li] != 0;
        if (l.full) {
            l.qn = dalloc<float>(D); l.kn = dalloc<float>(D);   // big matrices are allocated when their tensors arrive (bf16 or quantised)
        } else {
            l.w_in = dalloc<bf16>((size_t)gdn_in_pad * H);
            CUDA_OK(cudaMemset(l.w_in, 0, (size_t)gdn_in_pad * H * 2));
            l.w_out = dalloc<bf16>((size_t)H * value_dim());
            l.conv_w = dalloc<float>((size_t)conv_dim() * ck);
            l.neg_exp_a = dalloc<float>(nv); l.dt_bias = dalloc<float>(nv); l.gnorm = dalloc<float>(dv);
        }
        l.ln1 = dalloc<float>(H); l.ln2 = dalloc<float>(H);
    }
    if (is_vl) {
        const int pk = v_in * v_tpatch * v_patch * v_patch, mh = v_H * v_merge * v_merge;
        v_wpatch = dalloc<bf16>((size_t)v_H * pk);
        v_bpatch = dalloc<float>(v_H);
        v_pos = dalloc<float>((size_t)v_npos * v_H);
        vblocks.resize(v_depth);
        for (auto& b : vblocks) {
            b.n1w = dalloc<float>(v_H); b.n1b = dalloc<float>(v_H); b.n2w = dalloc<float>(v_H); b.n2b = dalloc<float>(v_H);
            b.wqkv = dalloc<bf16>((size_t)3 * v_H * v_H); b.bqkv = dalloc<float>(3 * v_H);
            b.wproj = dalloc<bf16>((size_t)v_H * v_H); b.bproj = dalloc<float>(v_H);
            b.wfc1 = dalloc<bf16>((size_t)v_I * v_H); b.bfc1 = dalloc<float>(v_I);
            b.wfc2 = dalloc<bf16>((size_t)v_H * v_I); b.bfc2 = dalloc<float>(v_H);
        }
        auto mk = [&](MergerW& m, bool post) {
            m.post = post;
            const int nd = post ? mh : v_H;
            m.nw = dalloc<float>(nd); m.nb = dalloc<float>(nd);
            m.w1 = dalloc<bf16>((size_t)mh * mh); m.b1 = dalloc<float>(mh);
            m.w2 = dalloc<bf16>((size_t)v_out * mh); m.b2 = dalloc<float>(v_out);
        };
        mk(v_merger, false);
        v_ds_mergers.resize(v_deepstack.size());
        for (auto& m : v_ds_mergers) mk(m, true);
    }
}

// =================================================================================================
// weight registration
// =================================================================================================
static std::string gguf_to_hf(const std::string& n);
static void want_shape(const std::string& name, const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
    size_t n = 1, w = 1;
    for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
    for (auto v : want) w *= (size_t)v;
    bool ok = (n == w);
    if (ok && ndim == (int)want.size()) {
        int i = 0;
        for (auto v : want) ok = ok && (shape[i++] == v);
    }
    if (!ok) {
        std::string got;
        for (int i = 0; i < ndim; ++i) got += (i ? "x" : "") + std::to_string(shape[i]);
        std::string exp;
        for (auto v : want) exp += (exp.empty() ? "" : "x") + std::to_string(v);
        fail(CRANE_B200_INVALID_ARG, "tensor %s: shape %s, expected %s", name.c_str(), got.c_str(), exp.c_str());
    }
}

bool crane_b200_model::load_text_tensor(const std::string& n, int dt, const int64_t* shape, int ndim, const void* data) {
    if (n == "embed_tokens.weight") {
        want_shape(n, shape, ndim, {V, H});
        up_bf16(embed, data, dt, (size_t)V * H);
        got_embed = true;
        return true;
    }
    if (n == "norm.weight") {
        want_shape(n, shape, ndim, {H});
        up_norm(final_norm, data, dt, H);
        got_final_norm = true;
        return true;
    }
    if (n.rfind("layers.", 0) != 0) return false;
    const size_t dot = n.find('.', 7);
    if (dot == std::string::npos) return false;
    const int li = std::atoi(n.substr(7, dot - 7).c_str());
    if (li < 0 || li >= L) fail(CRANE_B200_INVALID_ARG, "tensor %s: layer index out of range", n.c_str());
    const std::string t = n.substr(dot + 1);
    LayerW& l = layers[li];
    const int qd = q_dim(), kvd = nkv * D, qs = nh * q_stride();
    auto rows_bf16 = [&](bf16* dst, int rows, int cols) { up_bf16(dst, data, dt, (size_t)rows * cols); };
    if (l.full && !l.wqkv && (t == "self_attn.q_proj.weight" || t == "self_attn.k_proj.weight" || t == "self_attn.v_proj.weight"))
        l.wqkv = dalloc<bf16>((size_t)qkv_dim() * H);
    if (!l.wo && t == "self_attn.o_proj.weight") l.wo = dalloc<bf16>((size_t)H * q_dim());
    if (!l.wgu && (t == "mlp.gate_proj.weight" || t == "mlp.up_proj.weight")) l.wgu = dalloc<bf16>((size_t)2 * I * H);
    if (!l.wdown && t == "mlp.down_proj.weight") l.wdown = dalloc<bf16>((size_t)H * I);
    const bool attn_t = t.rfind("self_attn.", 0) == 0, gdn_t = t.rfind("linear_attn.", 0) == 0;
    if (attn_t && !l.full) fail(CRANE_B200_INVALID_ARG, "tensor %s: layer %d is a linear-attention layer", n.c_str(), li);
    if (gdn_t && l.full) fail(CRANE_B200_INVALID_ARG, "tensor %s: layer %d is a full-attention layer", n.c_str(), li);
    if (t == "self_attn.q_proj.weight") { want_shape(n, shape, ndim, {qs, H}); rows_bf16(l.wqkv, qs, H); l.loaded |= 1; }
    else if (t == "self_attn.k_proj.weight") { want_shape(n, shape, ndim, {kvd, H}); rows_bf16(l.wqkv + (size_t)qs * H, kvd, H); l.loaded |= 2; }
    else if (t == "self_attn.v_proj.weight") { want_shape(n, shape, ndim, {kvd, H}); rows_bf16(l.wqkv + (size_t)(qs + kvd) * H, kvd, H); l.loaded |= 4; }
    else if (t == "self_attn.o_proj.weight") { want_shape(n, shape, ndim, {H, qd}); rows_bf16(l.wo, H, qd); l.loaded |= 8; }
    else if (t == "mlp.gate_proj.weight" || t == "mlp.up_proj.weight") {
        // merged gate/up with rows interleaved (gate_j, up_j) so SiLU(gate)*up fuses into the GEMV/GEMM epilogue
        want_shape(n, shape, ndim, {I, H});
        std::vector<uint16_t> tmp;
        to_bf16(data, dt, (size_t)I * H, tmp);
        const bool up = (t == "mlp.up_proj.weight");
        CUDA_OK(cudaMemcpy2D(l.wgu + (up ? H : 0), (size_t)2 * H * 2, tmp.data(), (size_t)H * 2, (size_t)H * 2, I, cudaMemcpyHostToDevice));
        l.loaded |= up ? 32 : 16;
    }
    else if (t == "mlp.down_proj.weight") { want_shape(n, shape, ndim, {H, I}); rows_bf16(l.wdown, H, I); l.loaded |= 64; }
    else if (t == "input_layernorm.weight") { want_shape(n, shape, ndim, {H}); up_norm(l.ln1, data, dt, H); l.loaded |= 128; }
    else if (t == "post_attention_layernorm.weight") { want_shape(n, shape, ndim, {H}); up_norm(l.ln2, data, dt, H); l.loaded |= 256; }
    else if (t == "self_attn.q_norm.weight") { want_shape(n, shape, ndim, {D}); up_norm(l.qn, data, dt, D); l.loaded |= 512; }
    else if (t == "self_attn.k_norm.weight") { want_shape(n, shape, ndim, {D}); up_norm(l.kn, data, dt, D); l.loaded |= 1024; }
    // ---- Gated-Delta-Net (names: ops/gdn/layer.rs:55-67, projection.rs:75-83); in_proj rows merged as q|k|v, z, b, a ----
    // A llama.cpp-named tensor orders the value heads `Chunked` (index = replica * nk + key_head) where HF -- and every kernel here
    // -- has them `Interleaved` (key_head * (nv / nk) + replica) (ops/gdn/config.rs:13-22).  The reference adapts its Q/K expansion;
    // the tensors are dequantised here anyway, so the value-head blocks are simply put back into HF order: vhead() returns the f32
    // tensor with the nv blocks of `blk` elements that start at element `first` of every `period`-element group permuted.
    else if (t.rfind("linear_attn.", 0) == 0) {
        size_t numel = 1;
        for (int i = 0; i < ndim; ++i) numel *= (size_t)shape[i];
        const bool chunked = tensor_gguf_named && nv != nk;
        std::vector<float> f;
        auto vhead = [&](size_t period, size_t first, size_t blk) -> const void* {
            if (!chunked) return data;
            to_f32(data, dt, numel, f);
            std::vector<float> o(f);
            const int vpg = nv / nk;
            for (size_t g0 = 0; g0 + period <= numel; g0 += period)
                for (int kh = 0; kh < nk; ++kh)
                    for (int r = 0; r < vpg; ++r)
                        std::memcpy(&o[g0 + first + (size_t)(kh * vpg + r) * blk], &f[g0 + first + (size_t)(r * nk + kh) * blk], blk * sizeof(float));
            f.swap(o);
            return f.data();
        };
        const int dtp = chunked ? (int)CRANE_B200_F32 : dt;
        const size_t key2 = (size_t)2 * nk * dk;
        if (t == "linear_attn.in_proj_qkv.weight") {
            want_shape(n, shape, ndim, {conv_dim(), H});
            up_bf16(l.w_in, vhead(numel, key2 * H, (size_t)dv * H), dtp, (size_t)conv_dim() * H); l.loaded |= 1;
        } else if (t == "linear_attn.in_proj_z.weight") {
            want_shape(n, shape, ndim, {value_dim(), H});
            up_bf16(l.w_in + (size_t)conv_dim() * H, vhead(numel, 0, (size_t)dv * H), dtp, (size_t)value_dim() * H); l.loaded |= 2;
        } else if (t == "linear_attn.in_proj_b.weight") {
            want_shape(n, shape, ndim, {nv, H});
            up_bf16(l.w_in + (size_t)(conv_dim() + value_dim()) * H, vhead(numel, 0, H), dtp, (size_t)nv * H); l.loaded |= 4;
        } else if (t == "linear_attn.in_proj_a.weight") {
            want_shape(n, shape, ndim, {nv, H});
            up_bf16(l.w_in + (size_t)(conv_dim() + value_dim() + nv) * H, vhead(numel, 0, H), dtp, (size_t)nv * H); l.loaded |= 8;
        } else if (t == "linear_attn.conv1d.weight") {      // HF [conv_dim, 1, ck]; GGUF stores it 2-D [conv_dim, ck]
            want_shape(n, shape, ndim, {conv_dim(), 1, ck});
            up_f32(l.conv_w, vhead(numel, key2 * ck, (size_t)dv * ck), dtp, (size_t)conv_dim() * ck); l.loaded |= 512;
        } else if (t == "linear_attn.dt_bias") {
            want_shape(n, shape, ndim, {nv});
            up_f32(l.dt_bias, vhead(numel, 0, 1), dtp, nv); l.loaded |= 1024;
        } else if (t == "linear_attn.A_log" || t == "linear_attn.neg_exp_a") {
            want_shape(n, shape, ndim, {nv});
            std::vector<float> tmp;
            to_f32(vhead(numel, 0, 1), dtp, nv, tmp);
            // HF stores A_log, the kernels want -exp(A_log) (GdnGateConsts::new, ops/gdn/backend.rs:176-190); llama.cpp's `ssm_a` already
            // is -exp(A_log) (the reference takes (-ssm_a).ln() to get A_log back, qwen3_5/modeling.rs:747-750)
            if (t == "linear_attn.A_log") for (auto& v : tmp) v = -expf(v);
            CUDA_OK(cudaMemcpy(l.neg_exp_a, tmp.data(), nv * 4, cudaMemcpyHostToDevice));
            l.loaded |= 2048;
        } else if (t == "linear_attn.norm.weight") {
            want_shape(n, shape, ndim, {dv}); up_f32(l.gnorm, data, dt, dv); l.loaded |= 4096;
        } else if (t == "linear_attn.out_proj.weight") {     // value heads are COLUMN blocks here
            want_shape(n, shape, ndim, {H, value_dim()});
            up_bf16(l.w_out, vhead((size_t)value_dim(), 0, dv), dtp, (size_t)H * value_dim()); l.loaded |= 8192;
        } else return false;
    }
    else return false;
    return true;
}

bool crane_b200_model::load_vision_tensor(const std::string& n, int dt, const int64_t* shape, int ndim, const void* data) {
    if (!is_vl) fail(CRANE_B200_INVALID_ARG, "vision tensor %s for a text-only config", n.c_str());
    const int pk = v_in * v_tpatch * v_patch * v_patch, mh = v_H * v_merge * v_merge;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= (size_t)shape[i];
    if (n == "patch_embed.proj.weight") { want_shape(n, shape, ndim, {v_H, v_in, v_tpatch, v_patch, v_patch}); up_bf16(v_wpatch, data, dt, (size_t)v_H * pk); v_loaded |= 1; return true; }
    if (n == "patch_embed.proj.bias") { want_shape(n, shape, ndim, {v_H}); up_f32(v_bpatch, data, dt, v_H); v_loaded |= 2; return true; }
    if (n == "pos_embed.weight") { want_shape(n, shape, ndim, {v_npos, v_H}); up_f32(v_pos, data, dt, (size_t)v_npos * v_H); v_loaded |= 4; return true; }
    auto merger_tensor = [&](MergerW& m, const std::string& t) -> bool {
        const int nd = m.post ? mh : v_H;
        if (t == "norm.weight") { want_shape(n, shape, ndim, {nd}); up_f32(m.nw, data, dt, nd); m.loaded |= 1; }
        else if (t == "norm.bias") { want_shape(n, shape, ndim, {nd}); up_f32(m.nb, data, dt, nd); m.loaded |= 2; }
        else if (t == "linear_fc1.weight") { want_shape(n, shape, ndim, {mh, mh}); up_bf16(m.w1, data, dt, (size_t)mh * mh); m.loaded |= 4; }
        else if (t == "linear_fc1.bias") { want_shape(n, shape, ndim, {mh}); up_f32(m.b1, data, dt, mh); m.loaded |= 8; }
        else if (t == "linear_fc2.weight") { want_shape(n, shape, ndim, {v_out, mh}); up_bf16(m.w2, data, dt, (size_t)v_out * mh); m.loaded |= 16; }
        else if (t == "linear_fc2.bias") { want_shape(n, shape, ndim, {v_out}); up_f32(m.b2, data, dt, v_out); m.loaded |= 32; }
        else return false;
        return true;
    };
    if (n.rfind("merger.", 0) == 0) return merger_tensor(v_merger, n.substr(7));
    if (n.rfind("deepstack_merger_list.", 0) == 0) {
        const size_t dot = n.find('.', 22);
        const int j = std::atoi(n.substr(22, dot - 22).c_str());
        if (j < 0 || j >= (int)v_ds_mergers.size()) fail(CRANE_B200_INVALID_ARG, "tensor %s: merger index out of range", n.c_str());
        return merger_tensor(v_ds_mergers[j], n.substr(dot + 1));
    }
    if (n.rfind("blocks.", 0) != 0) return false;
    const size_t dot = n.find('.', 7);
    const int bi = std::atoi(n.substr(7, dot - 7).c_str());
    if (bi < 0 || bi >= v_depth) fail(CRANE_B200_INVALID_ARG, "tensor %s: block index out of range", n.c_str());
    const std::string t = n.substr(dot + 1);
    VitBlockW& b = vblocks[bi];
    if (t == "norm1.weight") { want_shape(n, shape, ndim, {v_H}); up_f32(b.n1w, data, dt, v_H); b.loaded |= 1; }
    else if (t == "norm1.bias") { want_shape(n, shape, ndim, {v_H}); up_f32(b.n1b, data, dt, v_H); b.loaded |= 2; }
    else if (t == "norm2.weight") { want_shape(n, shape, ndim, {v_H}); up_f32(b.n2w, data, dt, v_H); b.loaded |= 4; }
    else if (t == "norm2.bias") { want_shape(n, shape, ndim, {v_H}); up_f32(b.n2b, data, dt, v_H); b.loaded |= 8; }
    else if (t == "attn.qkv.weight") { want_shape(n, shape, ndim, {3 * v_H, v_H}); up_bf16(b.wqkv, data, dt, (size_t)3 * v_H * v_H); b.loaded |= 16; }
    else if (t == "attn.qkv.bias") { want_shape(n, shape, ndim, {3 * v_H}); up_f32(b.bqkv, data, dt, 3 * v_H); b.loaded |= 32; }
    else if (t == "attn.proj.weight") { want_shape(n, shape, ndim, {v_H, v_H}); up_bf16(b.wproj, data, dt, (size_t)v_H * v_H); b.loaded |= 64; }
    else if (t == "attn.proj.bias") { want_shape(n, shape, ndim, {v_H}); up_f32(b.bproj, data, dt, v_H); b.loaded |= 128; }
    else if (t == "mlp.linear_fc1.weight") { want_shape(n, shape, ndim, {v_I, v_H}); up_bf16(b.wfc1, data, dt, (size_t)v_I * v_H); b.loaded |= 256; }
    else if (t == "mlp.linear_fc1.bias") { want_shape(n, shape, ndim, {v_I}); up_f32(b.bfc1, data, dt, v_I); b.loaded |= 512; }
    else if (t == "mlp.linear_fc2.weight") { want_shape(n, shape, ndim, {v_H, v_I}); up_bf16(b.wfc2, data, dt, (size_t)v_H * v_I); b.loaded |= 1024; }
    else if (t == "mlp.linear_fc2.bias") { want_shape(n, shape, ndim, {v_H}); up_f32(b.bfc2, data, dt, v_H); b.loaded |= 2048; }
    else return false;
    (void)numel;
    return true;
}

void crane_b200_model::load_tensor(const std::string& name_in, int dt, const int64_t* shape, int ndim, const void* data) {
    if (is_tts && name_in.rfind("talker.", 0) == 0) { tts_load(name_in, dt, shape, ndim, data); return; }
    const std::string name = gguf_to_hf(name_in);
    struct Flag { bool& f; ~Flag() { f = false; } } flag_reset{tensor_gguf_named};
    tensor_gguf_named = name != name_in;
    if (finalized) fail(CRANE_B200_INVALID_ARG, "load_tensor after finalize");
    if (dt < 0 || dt > 2) fail(CRANE_B200_INVALID_ARG, "tensor %s: unknown dtype %d", name.c_str(), dt);
    CUDA_OK(cudaSetDevice(device));
    static const char* vis_prefix[] = {"model.visual.", "visual."};
    for (const char* p : vis_prefix)
        if (name.rfind(p, 0) == 0) {
            if (!load_vision_tensor(name.substr(std::strlen(p)), dt, shape, ndim, data))
                fail(CRANE_B200_INVALID_ARG, "unknown vision tensor %s", name.c_str());
            return;
        }
    if (name.rfind("lm_head.", 0) == 0 && name != "lm_head.weight") {       // lm_head.{g}.weight: extra output heads
        const int g = std::atoi(name.substr(8).c_str());
        if (g < 0 || g >= 64) fail(CRANE_B200_INVALID_ARG, "tensor %s: head index", name.c_str());
        want_shape(name, shape, ndim, {V, H});
        if ((int)heads.size() <= g) heads.resize(g + 1, nullptr);
        if (!heads[g]) heads[g] = dalloc<bf16>((size_t)V * H);
        up_bf16(heads[g], data, dt, (size_t)V * H);
        if (g == 0) { lm_head = heads[0]; got_lm_head = true; }
        return;
    }
    if (name.rfind("embed_tables.", 0) == 0) {                                 // embed_tables.{g}.weight: extra [rows, H] tables
        const int g = std::atoi(name.substr(13).c_str());
        if (g < 0 || g >= 64 || ndim != 2 || shape[1] != H) fail(CRANE_B200_INVALID_ARG, "tensor %s: bad table", name.c_str());
        if (table_rows && table_rows != (int)shape[0]) fail(CRANE_B200_INVALID_ARG, "tensor %s: table rows differ", name.c_str());
        table_rows = (int)shape[0];
        if ((int)emb_tables.size() <= g) emb_tables.resize(g + 1, nullptr);
        if (!emb_tables[g]) emb_tables[g] = dalloc<bf16>((size_t)table_rows * H);
        up_bf16(emb_tables[g], data, dt, (size_t)table_rows * H);
        return;
    }
    if (name == "lm_head.weight") {
        want_shape(name, shape, ndim, {V, H});
        if (!tied) {
            if (!lm_head) lm_head = dalloc<bf16>((size_t)V * H);
            up_bf16(lm_head, data, dt, (size_t)V * H);
            got_lm_head = true;
        }
        return;   // tied checkpoints may still carry a copy: ignored, as the reference does (qwen3/modeling.rs:786-794)
    }
    static const char* txt_prefix[] = {"model.language_model.", "language_model.model.", "language_model.", "model."};
    for (const char* p : txt_prefix)
        if (name.rfind(p, 0) == 0) {
            if (!load_text_tensor(name.substr(std::strlen(p)), dt, shape, ndim, data))
                fail(CRANE_B200_INVALID_ARG, "unknown tensor %s", name.c_str());
            return;
        }
    fail(CRANE_B200_INVALID_ARG, "unknown tensor %s", name.c_str());
}

// GGUF tensor names as read by the reference's GGUF loaders (qwen3/modeling.rs:252-282,598-606,686-689,898-916)
static std::string gguf_to_hf(const std::string& n) {
    if (n == "token_embd.weight") return "model.embed_tokens.weight";
    if (n == "output_norm.weight") return "model.norm.weight";
    if (n == "output.weight") return "lm_head.weight";
    if (n.rfind("blk.", 0) != 0) return n;
    const size_t dot = n.find('.', 4);
    if (dot == std::string::npos) return n;
    const std::string idx = n.substr(4, dot - 4), t = n.substr(dot + 1);
    static const char* map[][2] = {
        {"attn_norm.weight", "input_layernorm.weight"}, {"ffn_norm.weight", "post_attention_layernorm.weight"},
        {"attn_q.weight", "self_attn.q_proj.weight"}, {"attn_k.weight", "self_attn.k_proj.weight"}, {"attn_v.weight", "self_attn.v_proj.weight"},
        {"attn_output.weight", "self_attn.o_proj.weight"}, {"attn_q_norm.weight", "self_attn.q_norm.weight"},
        {"attn_k_norm.weight", "self_attn.k_norm.weight"}, {"ffn_gate.weight", "mlp.gate_proj.weight"}, {"ffn_up.weight", "mlp.up_proj.weight"},
        {"ffn_down.weight", "mlp.down_proj.weight"},
        // llama.cpp `qwen35` (qwen3_5/modeling.rs:688-775): the split Gated-Delta-Net projections and their side tensors
        {"post_attention_norm.weight", "post_attention_layernorm.weight"}, {"attn_qkv.weight", "linear_attn.in_proj_qkv.weight"},
        {"attn_gate.weight", "linear_attn.in_proj_z.weight"}, {"ssm_beta.weight", "linear_attn.in_proj_b.weight"},
        {"ssm_alpha.weight", "linear_attn.in_proj_a.weight"}, {"ssm_conv1d.weight", "linear_attn.conv1d.weight"},
        {"ssm_dt.bias", "linear_attn.dt_bias"}, {"ssm_a", "linear_attn.neg_exp_a"}, {"ssm_norm.weight", "linear_attn.norm.weight"},
        {"ssm_out.weight", "linear_attn.out_proj.weight"}};
    for (auto& m : map)
        if (t == m[0]) return "model.layers." + idx + "." + m[1];
    return n;
}

// Upload `rows` x K quantised weights: repack on the host, copy row-wise (dst_row_pitch lets gate/up rows interleave).
unsigned char* crane_b200_model::up_quant(int qt, const void* data, size_t rows, int K, unsigned char* dst, size_t dst_row_pitch) {
    const size_t rb = (size_t)(K / 256) * q_sb_bytes(qt);
    std::vector<unsigned char> tmp(rows * rb);
    q_repack_rows(qt, (const unsigned char*)data, tmp.data(), rows, K);
    if (!dst) { dst = dalloc<unsigned char>(rows * rb); dst_row_pitch = rb; }
    CUDA_OK(cudaMemcpy2D(dst, dst_row_pitch, tmp.data(), rb, rb, rows, cudaMemcpyHostToDevice));
    return dst;
}

void crane_b200_model::load_tensor_ggml(const std::string& name_in, int qt, const int64_t* shape, int ndim, const void* data, size_t nbytes) {
    if (finalized) fail(CRANE_B200_INVALID_ARG, "load_tensor after finalize");
    if (qt != QT_Q4_K && qt != QT_Q6_K && qt != QT_Q8_0) fail(CRANE_B200_UNSUPPORTED, "ggml type %d (supported: Q8_0=8, Q4_K=12, Q6_K=14)", qt);
    if (is_vl) fail(CRANE_B200_UNSUPPORTED, "quantised tensors are not supported for the VL model");
    if (ndim != 2) fail(CRANE_B200_INVALID_ARG, "tensor %s: quantised tensors are 2-D [rows, cols]", name_in.c_str());
    if (hybrid) {
        // The hybrid's kernels stream bf16: its GGUF tensors are dequantised at load (the reference keeps QTensor blocks for the
        // linears -- qwen3_5/modeling.rs:379-411,704-775 -- so logits follow `x . dequant(W)^T`, not candle's int8-activation QMatMul)
        const int64_t r_ = shape[0], k_ = shape[1];
        const int be = q_src_block_elems(qt);
        if (r_ <= 0 || k_ <= 0 || k_ % be) fail(CRANE_B200_INVALID_ARG, "tensor %s: bad quantised shape", name_in.c_str());
        if (nbytes != (size_t)r_ * (k_ / be) * q_src_block_bytes(qt)) fail(CRANE_B200_INVALID_ARG, "tensor %s: byte count", name_in.c_str());
        std::vector<float> f;
        ggml_dequant_rows(qt, (const unsigned char*)data, (size_t)r_, (size_t)k_, f);
        load_tensor(name_in, CRANE_B200_F32, shape, ndim, f.data());
        return;
    }
    CUDA_OK(cudaSetDevice(device));
    std::string name = gguf_to_hf(name_in);
    const int64_t rows = shape[0], K = shape[1];
    if (K % 256) fail(CRANE_B200_UNSUPPORTED, "tensor %s: in_dim %lld is not a multiple of 256 (ops/linear.rs:86-94 falls back to Q8_0 there)", name.c_str(), (long long)K);
    const size_t want = (size_t)rows * (K / q_src_block_elems(qt)) * q_src_block_bytes(qt);
    if (nbytes != want) fail(CRANE_B200_INVALID_ARG, "tensor %s: %zu bytes, expected %zu", name.c_str(), nbytes, want);
    any_quant = true;
    if (name == "model.embed_tokens.weight") {
        // the table keeps its blocks; gathered rows are dequantised to f32 (QuantizedEmbedding, modules/embedding.rs:31-105);
        // a tied head streams the same bytes
        want_shape(name, shape, ndim, {V, H});
        q_embed = up_quant(qt, data, (size_t)V, H);
        qt_embed = qt;
        dfree(embed);
        embed = nullptr;
        got_embed = true;
        if (tied) { q_lm_head = q_embed; qt_lm = qt; lm_head = nullptr; }
        return;
    }
    if (name == "lm_head.weight") {
        want_shape(name, shape, ndim, {V, H});
        if (!tied) { q_lm_head = up_quant(qt, data, (size_t)V, H); qt_lm = qt; got_lm_head = true; }
        return;
    }
    static const char* txt_prefix[] = {"model.language_model.", "model."};
    std::string n;
    for (const char* p : txt_prefix)
        if (name.rfind(p, 0) == 0) { n = name.substr(std::strlen(p)); break; }
    if (n.rfind("layers.", 0) != 0) fail(CRANE_B200_INVALID_ARG, "unknown quantised tensor %s", name_in.c_str());
    const size_t dot = n.find('.', 7);
    const int li = std::atoi(n.substr(7, dot - 7).c_str());
    if (li < 0 || li >= L) fail(CRANE_B200_INVALID_ARG, "tensor %s: layer index out of range", name.c_str());
    const std::string t = n.substr(dot + 1);
    LayerW& l = layers[li];
    const int qd = q_dim(), kvd = nkv * D;
    if (t == "self_attn.q_proj.weight") { want_shape(name, shape, ndim, {qd, H}); l.q_wq = up_quant(qt, data, qd, H); l.qt_q = qt; l.loaded |= 1; }
    else if (t == "self_attn.k_proj.weight") { want_shape(name, shape, ndim, {kvd, H}); l.q_wk = up_quant(qt, data, kvd, H); l.qt_k = qt; l.loaded |= 2; }
    else if (t == "self_attn.v_proj.weight") { want_shape(name, shape, ndim, {kvd, H}); l.q_wv = up_quant(qt, data, kvd, H); l.qt_v = qt; l.loaded |= 4; }
    else if (t == "self_attn.o_proj.weight") { want_shape(name, shape, ndim, {H, qd}); l.q_wo = up_quant(qt, data, H, qd); l.qt_o = qt; l.loaded |= 8; }
    else if (t == "mlp.gate_proj.weight" || t == "mlp.up_proj.weight") {
        want_shape(name, shape, ndim, {I, H});
        const bool up = (t == "mlp.up_proj.weight");
        if (l.qt_gu && l.qt_gu != qt) fail(CRANE_B200_UNSUPPORTED, "layer %d: gate_proj and up_proj use different ggml types (%d vs %d)", li, l.qt_gu, qt);
        const size_t rb = (size_t)(H / 256) * q_sb_bytes(qt);
        if (!l.q_wgu) l.q_wgu = dalloc<unsigned char>((size_t)2 * I * rb);
        l.qt_gu = qt;
        (up ? l.qt_up_seen : l.qt_gate_seen) = qt;
        up_quant(qt, data, I, H, l.q_wgu + (up ? rb : 0), 2 * rb);     // rows interleaved (gate_j, up_j)
        l.loaded |= up ? 32 : 16;
    }
    else if (t == "mlp.down_proj.weight") { want_shape(name, shape, ndim, {H, I}); l.q_wdown = up_quant(qt, data, H, I); l.qt_down = qt; l.loaded |= 64; }
    else fail(CRANE_B200_INVALID_ARG, "tensor %s cannot be quantised here", name.c_str());
}

// One decode-path linear: bf16 GEMV or quantised GEMV with the same fused epilogue.
void crane_b200_model::linear_decode(int epi, bool norm, const bf16* w, const unsigned char* qw, int qt, int N, int K, const float* xin, int ldx,
                                     const float* norm_w, float* y, int ldy, const GemvArgs* extra, int B, bool reuse_xq) {
    GemvArgs g = extra ? *extra : GemvArgs{};
    g.N = N; g.K = K; g.x = xin; g.ldx = ldx; g.norm_w = norm_w; g.eps = eps; g.y = y; g.ldy = ldy;
    if (qt) {
        QGemvArgs qa;
        g.W = reinterpret_cast<const bf16*>(qw);
        qa.g = g; qa.qtype = qt; qa.epi = epi; qa.norm = norm ? 1 : 0; qa.xq = xq_buf;
        const int xmode = xq_mode_for(qt);
        if (!reuse_xq || xmode != xq_mode_last) {     // same input (norm, block rule) as the previous quantised linear: k and v after q
            LAUNCH_OK(xquant_launch(stream, B, xin, ldx, K, norm ? norm_w : nullptr, eps, xmode, xq_buf, use_pdl));
            xq_mode_last = xmode;
            ++launches;
        }
        LAUNCH_OK(qgemv_launch(stream, B, qa, num_sms, use_pdl));
    } else {
        g.W = w;
        LAUNCH_OK(gemv_launch(stream, B, epi, norm, g, num_sms, use_pdl));
    }
    ++launches;
}

// Prefill through a quantised linear: candle's `QMatMul::forward` quantises every activation ROW to Q8_K / Q8_0 blocks and takes
// ggml integer dots whatever the number of rows (crane-core/src/ops/linear.rs:23-48), so S > 1 runs the same xquant + qgemv
// kernels as decode, four rows per pass over the quantised weights.  (A tensor-core formulation would have to reproduce the
// per-block integer dots exactly; the decode kernels already do.)
void crane_b200_model::qlinear_rows(int epi, const unsigned char* qw, int qt, int N, int K, const float* xin, int ldx, const float* norm_w,
                                    float* y, int ldy, int S, bool reuse_xq) {
    for (int s0 = 0; s0 < S;) {
        const int B = (S - s0 >= 4) ? 4 : (S - s0 >= 2) ? 2 : 1;
        linear_decode(epi, norm_w != nullptr, nullptr, qw, qt, N, K, xin + (size_t)s0 * ldx, ldx, norm_w, y + (size_t)s0 * ldy, ldy, nullptr, B,
                      reuse_xq && S <= 4);
        s0 += B;
    }
}

// =================================================================================================
// finalize: tables, KV pages, decode buffers
// =================================================================================================
void crane_b200_model::finalize() {
    if (finalized) return;
    CUDA_OK(cudaSetDevice(device));
    if (!got_embed) fail(CRANE_B200_NOT_LOADED, "missing tensor embed_tokens.weight");
    if (!got_final_norm) fail(CRANE_B200_NOT_LOADED, "missing tensor norm.weight");
    if (!tied && !got_lm_head) fail(CRANE_B200_NOT_LOADED, "missing tensor lm_head.weight");
    if (any_quant) {
        size_t mx = 0;
        for (auto& l : layers) {
            if (l.qt_q || l.qt_k || l.qt_v) {
                if (!(l.qt_q && l.qt_k && l.qt_v)) fail(CRANE_B200_UNSUPPORTED, "q/k/v must be all quantised or all bf16 within a layer");
                mx = std::max(mx, (size_t)qkv_dim() * H);
            }
            if ((l.qt_gate_seen != 0) != (l.qt_up_seen != 0) || (l.qt_gu && ((l.loaded & 48) != 48 || l.wgu != nullptr)))
                fail(CRANE_B200_UNSUPPORTED, "gate_proj and up_proj of a layer must both be quantised (same ggml type) or both bf16");
            if (l.qt_o) mx = std::max(mx, (size_t)H * q_dim());
            if (l.qt_gu) mx = std::max(mx, (size_t)2 * I * H);
            if (l.qt_down) mx = std::max(mx, (size_t)H * I);
        }
        (void)mx;
        xq_buf = dalloc<unsigned char>(xquant_bytes(4, std::max(std::max(H, I), q_dim())));      // groups of <= 4 rows / sequences
    }
    for (int i = 0; i < L; ++i) {
        // full: q,k,v,o (1|2|4|8) gate,up,down (16|32|64) ln1,ln2 (128|256) q_norm,k_norm (512|1024)
        // GDN : qkv,z,b,a (1|2|4|8) gate,up,down, ln1,ln2, conv (512) dt_bias (1024) A_log (2048) norm (4096) out_proj (8192)
        const int want = layers[i].full ? 2047 : (15 | 16 | 32 | 64 | 128 | 256 | 512 | 1024 | 2048 | 4096 | 8192);
        if (layers[i].loaded != want)
            fail(CRANE_B200_NOT_LOADED, "layer %d: missing tensors (mask 0x%x, want 0x%x)", i, layers[i].loaded, want);
    }
    if (is_vl) {
        if (v_loaded != 7) fail(CRANE_B200_NOT_LOADED, "vision stem: missing tensors (mask 0x%x)", v_loaded);
        for (int i = 0; i < v_depth; ++i)
            if (vblocks[i].loaded != 4095) fail(CRANE_B200_NOT_LOADED, "vision block %d: missing tensors (mask 0x%x)", i, vblocks[i].loaded);
        if (v_merger.loaded != 63) fail(CRANE_B200_NOT_LOADED, "vision merger: missing tensors");
        for (auto& m : v_ds_mergers)
            if (m.loaded != 63) fail(CRANE_B200_NOT_LOADED, "deepstack merger: missing tensors");
    }
    // RotaryEmbedding::new (modules/rotary.rs:29-46): inv_freq f64 -> f32, freqs = pos_f32 * inv_freq (f32), cos/sin f32
    // Qwen3.5: MRotaryEmbedding::new (qwen3_5/modeling.rs:110-132) -- table over the rotary slice only, inv_freq in F32
    const int half = rot_half, rows = max_seq + 1;
    std::vector<float> inv(half), ct((size_t)rows * half), st((size_t)rows * half);
    for (int i = 0; i < half; ++i)
        inv[i] = hybrid ? 1.0f / powf((float)theta, (float)i * 2.0f / (float)(2 * rot_half))
                        : (float)(1.0 / std::pow(theta, (double)(2 * i) / (double)D));
    for (int p = 0; p < rows; ++p)
        for (int i = 0; i < half; ++i) {
            const float f = (float)p * inv[i];
            ct[(size_t)p * half + i] = cosf(f);
            st[(size_t)p * half + i] = sinf(f);
        }
    cos_tab = dalloc<float>(ct.size());
    sin_tab = dalloc<float>(st.size());
    CUDA_OK(cudaMemcpy(cos_tab, ct.data(), ct.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(sin_tab, st.data(), st.size() * 4, cudaMemcpyHostToDevice));
    // interleaved-MRoPE column ownership (qwen3_5/modeling.rs:203-233)
    std::vector<unsigned char> ax(half, 0);
    for (int dim = 1; dim <= 2; ++dim) {
        const int sec = dim < (int)mrope_section.size() ? mrope_section[dim] : 0;
        for (int i = dim; i < std::min(sec * 3, half); i += 3) ax[i] = (unsigned char)dim;
    }
    axis_of = dalloc<unsigned char>(half);
    CUDA_OK(cudaMemcpy(axis_of, ax.data(), half, cudaMemcpyHostToDevice));

    const size_t page_elems = (size_t)nkv * KV_PAGE * D;
    for (auto& l : layers) {
        if (l.full && kv_bits) {
            const size_t rows = (size_t)max_batch * max_pages * nkv * KV_PAGE;
            l.k_codes = dalloc<unsigned char>(rows * D * kv_bits / 8); l.v_codes = dalloc<unsigned char>(rows * D * kv_bits / 8);
            l.k_scale = dalloc<float>(rows); l.v_scale = dalloc<float>(rows);
        } else if (l.full) {
            l.k_pool = dalloc_act((size_t)max_batch * max_pages * page_elems, lo_kv);
            l.v_pool = dalloc_act((size_t)max_batch * max_pages * page_elems, lo_kv);
        } else {   // GdnLayerCache (ops/gdn/cache.rs:15-45): conv window + [Hv, K, V] f32 state, zero-initialised
            l.conv_state = dalloc<float>((size_t)conv_dim() * ck);
            l.rec_state = dalloc<float>((size_t)nv * dk * dv);
            if (max_batch > 1) {   // one parked copy per sequence slot (the reference keeps a GdnLayerCache per sequence, cache.rs:15-45)
                l.conv_slots = dalloc<float>((size_t)max_batch * conv_dim() * ck);
                l.rec_slots = dalloc<float>((size_t)max_batch * nv * dk * dv);
                CUDA_OK(cudaMemset(l.conv_slots, 0, (size_t)max_batch * conv_dim() * ck * sizeof(float)));
                CUDA_OK(cudaMemset(l.rec_slots, 0, (size_t)max_batch * nv * dk * dv * sizeof(float)));
            }
        }
    }
    if (hybrid) {
        gd_proj = dalloc<float>(gdn_in_pad); gd_conv = dalloc<float>(conv_dim());
        gd_qn = dalloc<float>((size_t)nk * dk); gd_kn = dalloc<float>((size_t)nk * dk);
        gd_gb = dalloc<float>((size_t)nv * 2); gd_y = dalloc<float>(value_dim()); gd_out = dalloc<float>(value_dim());
        reset_recurrent_state();
    }
    if (kv_bits) {
        if (is_tts) fail(CRANE_B200_UNSUPPORTED, "engine.kv_cache = int8 / int4 is not wired into the TTS handle");
        kq_sk = dalloc_act((size_t)max_pages * page_elems, lo_kv);       // (lo_kv now names the scratch's low-order plane)
        kq_sv = dalloc_act((size_t)max_pages * page_elems, lo_kv);
        std::vector<int> idn(max_pages);
        for (int i = 0; i < max_pages; ++i) idn[i] = i;
        bt_identity = dalloc<int>(max_pages);
        CUDA_OK(cudaMemcpy(bt_identity, idn.data(), idn.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    // page tables: slot s owns pages [s * max_pages, (s+1) * max_pages) -- static assignment, 64-token pages
    std::vector<int> bt((size_t)max_batch * max_pages);
    for (size_t i = 0; i < bt.size(); ++i) bt[i] = (int)i;
    block_table = dalloc<int>(bt.size());
    CUDA_OK(cudaMemcpy(block_table, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice));
    seq_kv.assign(max_batch, 0); seq_pos.assign(max_batch, 0); seq_used.assign(max_batch, 0);
    seq_used[0] = 1;

    const int B = std::min(max_batch, 4);   // decode batches run in groups of <= 4 sequences per weight pass
    x_dec = dalloc<float>((size_t)B * H);
    qkv_dec = dalloc<float>((size_t)B * qkv_dim());
    attn_dec = dalloc<float>((size_t)B * q_dim());
    act_dec = dalloc<float>((size_t)B * I);
    logits = dalloc<float>((size_t)B * V);
    hidden_out = dalloc<float>((size_t)B * H);
    part_o = dalloc<float>((size_t)B * nh * ATTN_NSPLIT * D);
    part_ml = dalloc<float>((size_t)B * nh * ATTN_NSPLIT * 2);
    part_val = dalloc<float>((size_t)B * num_sms);
    part_idx = dalloc<int>((size_t)B * num_sms);
    counters = dalloc<unsigned int>((size_t)B * nkv);
    ticket = dalloc<unsigned int>(1);
    state = dalloc<SeqState>(B);
    out_tokens = dalloc<uint32_t>((size_t)B * out_cap);
    CUDA_OK(cudaMemset(counters, 0, (size_t)B * nkv * sizeof(unsigned int)));
    CUDA_OK(cudaMemset(ticket, 0, sizeof(unsigned int)));
    CUDA_OK(cudaMemset(state, 0, B * sizeof(SeqState)));
    // single-sequence bf16 decode runs as ONE persistent launch per call (decode_ll.cu) when the geometry allows it
    use_persistent = use_persistent && !hybrid && !any_quant && !is_tts && !kv_bits && owns_stream && L <= LL_MAX_LAYERS &&
                     decode_ll_supported(D, rot_half, nh, nkv, H, I, q_dim(), qkv_dim(), V, num_sms);
    if (use_persistent) {
        int coop = 0;
        CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        use_persistent = coop != 0;
    }
    if (use_persistent) {
        auto pairs = [&](size_t n) { auto* p = dalloc<unsigned long long>(n); CUDA_OK(cudaMemset(p, 0, n * 8)); return p; };
        ll_xa = pairs(H); ll_xb = pairs(H); ll_qkv = pairs(qkv_dim()); ll_att = pairs(q_dim()); ll_act = pairs(I);
        ll_part = pairs(decode_ll_part_pairs(num_sms, nh, nkv));
        ll_amax = pairs((size_t)4 * num_sms);
        ll_err = dalloc<unsigned int>(16 + num_sms * LL_WARPS);       // flag, first failing wait, per-CTA heartbeats (decode_ll.cu `Waiter`)
        CUDA_OK(cudaMemset(ll_err, 0, sizeof(unsigned int) * (16 + num_sms * LL_WARPS)));
        CUDA_OK(cudaMallocHost((void**)&h_ll_err, sizeof(unsigned int)));
        *h_ll_err = 0;
        if (getenv("CRANE_B200_PROF")) { ll_prof = dalloc<unsigned long long>(16); CUDA_OK(cudaMemset(ll_prof, 0, 128)); }
        if (getenv("CRANE_B200_LL_TRACE")) ll_trace = dalloc<unsigned long long>((size_t)6 * LL_TRACE_CAP * 2);
    }
    CUDA_OK(cudaMallocHost((void**)&h_state_ring, sizeof(SeqState) * STATE_RING * STATE_GROUP));
    for (auto& e : h_state_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_OK(cudaMallocHost((void**)&h_tokens, sizeof(uint32_t) * out_cap));
    CUDA_OK(cudaEventCreate(&pev0));
    CUDA_OK(cudaEventCreate(&pev1));
    CUDA_OK(cudaEventCreate(&dev0));
    CUDA_OK(cudaEventCreate(&dev1));
    if (is_tts) tts_setup();
    CUDA_OK(cudaDeviceSynchronize());
    finalized = true;
}

void crane_b200_model::ensure_prefill_ws(int S) {
    if (S <= ws_S) return;
    CUDA_OK(cudaStreamSynchronize(stream));
    for (void* p : {(void*)x, (void*)qkv, (void*)xn, (void*)q_bf, (void*)attn_bf, (void*)act_bf, (void*)ids_dev, (void*)pos3_dev,
                    (void*)rows_dev, (void*)embeds_in, (void*)rows_f32, (void*)g_proj, (void*)g_conv, (void*)g_qn, (void*)g_kn, (void*)g_gb, (void*)g_y, (void*)g_gl, (void*)g_chunk})
        dfree(p);
    const int cap = (S + 127) / 128 * 128;
    x = dalloc<float>((size_t)cap * H);
    qkv = dalloc<float>((size_t)cap * qkv_dim());
    xn = dalloc_act((size_t)cap * H, lo_xn);
    q_bf = dalloc_act((size_t)cap * q_dim(), lo_q);
    attn_bf = dalloc_act((size_t)cap * std::max(q_dim(), value_dim()), lo_attn);
    if (hybrid) {
        g_proj = dalloc<float>((size_t)cap * gdn_in_pad); g_conv = dalloc<float>((size_t)cap * conv_dim());
        g_qn = dalloc<float>((size_t)cap * nk * dk); g_kn = dalloc<float>((size_t)cap * nk * dk);
        g_gb = dalloc<float>((size_t)cap * nv * 2); g_y = dalloc<float>((size_t)cap * value_dim());
        g_gl = nullptr; g_chunk = nullptr;
        if (gdn_mode != 1 && cap >= cb::GDN_CHUNK && dv % 64 == 0) {     // chunkwise recurrence: 3.8 KB of scratch per (token, value head) at dk = dv = 128
            g_gl = dalloc<float>((size_t)cap * nv);
            g_chunk = dalloc<unsigned char>(cb::gdn_chunk_ws_bytes(cap, nv, dk, dv));
        }
    }
    act_bf = dalloc_act((size_t)cap * I, lo_act);
    if (h_stage) { cudaFreeHost(h_stage); h_stage = nullptr; }
    CUDA_OK(cudaMallocHost((void**)&h_stage, (size_t)5 * cap * sizeof(int)));
    if (!stage_ev) CUDA_OK(cudaEventCreateWithFlags(&stage_ev, cudaEventDisableTiming));
    stage_busy = false;
    ids_dev = dalloc<uint32_t>(cap);
    pos3_dev = dalloc<int>((size_t)3 * cap);
    rows_dev = dalloc<int>(cap);
    embeds_in = dalloc<float>((size_t)cap * H);
    rows_f32 = any_quant ? dalloc<float>((size_t)cap * std::max(I, q_dim())) : nullptr;
    ws_S = cap;
}

// =================================================================================================
// decode step
// =================================================================================================
void crane_b200_model::arm_state(uint32_t token, size_t start_pos, int p0, int p1, int p2) {
    stage_state();
    h_state[0].kv_len = (int)start_pos;
    h_state[0].pos[0] = p0; h_state[0].pos[1] = p1; h_state[0].pos[2] = p2;
    h_state[0].token = token;
    h_state[0].step = 0;
    h_state[0].slot = cur;
    push_state(1);
}

// x_dec <- the embedding rows of state[b].token (bf16 table, or rows dequantised out of a quantised table)
void crane_b200_model::embed_step_input(int B) {
    if (q_embed) LAUNCH_OK(embed_rows_q_launch(stream, qt_embed, q_embed, H, nullptr, state, B, x_dec, false));
    else LAUNCH_OK(embed_decode_launch(stream, B, embed, H, state, x_dec, false));
    ++launches;
}

void crane_b200_model::enqueue_decode_step(int advance, bool with_embed, int B) {
    const bool pdl = use_pdl;
    if (with_embed) embed_step_input(B);
    for (int li = 0; li < L; ++li) {
        LayerW& l = layers[li];
        if (l.full) {
            if (l.qt_q) {   // GGUF keeps q/k/v as separate quantised matrices (qwen3/modeling.rs:252-255): three GEMVs into one qkv row
                const int qs = nh * q_stride(), kvd = nkv * D;
                linear_decode(GEMV_STORE, true, nullptr, l.q_wq, l.qt_q, qs, H, x_dec, H, l.ln1, qkv_dec, qkv_dim(), nullptr, B);
                linear_decode(GEMV_STORE, true, nullptr, l.q_wk, l.qt_k, kvd, H, x_dec, H, l.ln1, qkv_dec + qs, qkv_dim(), nullptr, B, true);
                linear_decode(GEMV_STORE, true, nullptr, l.q_wv, l.qt_v, kvd, H, x_dec, H, l.ln1, qkv_dec + qs + kvd, qkv_dim(), nullptr, B, true);
            } else {
                linear_decode(GEMV_STORE, true, l.wqkv, nullptr, 0, qkv_dim(), H, x_dec, H, l.ln1, qkv_dec, qkv_dim(), nullptr, B);
            }
            AttnDecArgs a = {};
            a.qkv = qkv_dec; a.q_stride = q_stride(); a.gated = hybrid ? 1 : 0; a.rot_half = rot_half;
            a.q_norm_w = l.qn; a.k_norm_w = l.kn; a.eps = eps; a.cos_tab = cos_tab; a.sin_tab = sin_tab; a.axis_of = axis_of;
            a.state = state; a.block_table = block_table; a.max_pages = max_pages; a.k_pool = l.k_pool; a.v_pool = l.v_pool;
            a.nh = nh; a.nkv = nkv; a.scale = 1.0f / std::sqrt((float)D);
            a.out = attn_dec; a.kv_lo_off = kv_bits ? 0 : lo_kv;
            a.kv_bits = kv_bits; a.kv_split = split ? 1 : 0; a.k_codes = l.k_codes; a.v_codes = l.v_codes; a.k_scale = l.k_scale; a.v_scale = l.v_scale;
            LAUNCH_OK(attn_decode_launch(stream, B, D, a, pdl));
            ++launches;   // attention
            linear_decode(GEMV_RESID, false, l.wo, l.q_wo, l.qt_o, H, q_dim(), attn_dec, q_dim(), nullptr, x_dec, H, nullptr, B);
        } else {   // Gated-Delta-Net token mixer (ops/gdn/layer.rs:122-163), S = 1
            GemvArgs g = {};
            g.W = l.w_in; g.N = gdn_in_pad; g.K = H; g.x = x_dec; g.ldx = H; g.norm_w = l.ln1; g.eps = eps; g.y = gd_proj; g.ldy = gdn_in_pad;
            LAUNCH_OK(gemv_launch(stream, B, GEMV_STORE, true, g, num_sms, pdl));
            GdnArgs ga;
            gdn_args(ga, l, 1, gd_proj, gd_conv, gd_qn, gd_kn, gd_gb, gd_y);
            ga.out_f32 = gd_out;
            LAUNCH_OK(gdn_forward_launch(stream, ga));
            GemvArgs o = {};
            o.W = l.w_out; o.N = H; o.K = value_dim(); o.x = gd_out; o.ldx = value_dim(); o.y = x_dec; o.ldy = H;
            LAUNCH_OK(gemv_launch(stream, B, GEMV_RESID, false, o, num_sms, pdl));
            launches += 2 + gdn_forward_launch_count(ga);     // the two GEMVs around it + one fused kernel or five
        }
        linear_decode(GEMV_SILU_MUL, true, l.wgu, l.q_wgu, l.qt_gu, 2 * I, H, x_dec, H, l.ln2, act_dec, I, nullptr, B);
        linear_decode(GEMV_RESID, false, l.wdown, l.q_wdown, l.qt_down, H, I, act_dec, I, nullptr, x_dec, H, nullptr, B);
    }
    lm_head_last_row(x_dec, advance, B);
}

void crane_b200_model::gdn_args(GdnArgs& g, const LayerW& l, int S, const float* proj, float* conv, float* qn, float* kn, float* gb,
                                float* y) const {
    g = GdnArgs{};
    g.proj = proj; g.ldp = gdn_in_pad; g.S = S; g.nk = nk; g.nv = nv; g.dk = dk; g.dv = dv; g.ck = ck;
    g.conv_w = l.conv_w; g.conv_state = l.conv_state; g.neg_exp_a = l.neg_exp_a; g.dt_bias = l.dt_bias; g.norm_w = l.gnorm; g.eps = eps;
    g.rec_state = l.rec_state; g.conv_out = conv; g.qn = qn; g.kn = kn; g.gb = gb; g.y = y;
}

void crane_b200_model::park_gdn_state(int slot) {
    if (!hybrid || max_batch <= 1) return;
    const size_t cs = (size_t)conv_dim() * ck, rs = (size_t)nv * dk * dv;
    for (auto& l : layers)
        if (!l.full) {
            CUDA_OK(cudaMemcpyAsync(l.conv_slots + slot * cs, l.conv_state, cs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
            CUDA_OK(cudaMemcpyAsync(l.rec_slots + slot * rs, l.rec_state, rs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        }
}

void crane_b200_model::select_seq(int s) {
    if (s == cur) return;
    seq_kv[cur] = kv_len; seq_pos[cur] = next_mrope_pos;
    park_gdn_state(cur);
    cur = s;
    kv_len = seq_kv[s]; next_mrope_pos = seq_pos[s];
    if (hybrid && max_batch > 1) {
        const size_t cs = (size_t)conv_dim() * ck, rs = (size_t)nv * dk * dv;
        for (auto& l : layers)
            if (!l.full) {
                CUDA_OK(cudaMemcpyAsync(l.conv_state, l.conv_slots + s * cs, cs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
                CUDA_OK(cudaMemcpyAsync(l.rec_state, l.rec_slots + s * rs, rs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
            }
    }
}

// dst becomes a copy of src: the used KV pages of every attention layer (both precision planes) and the parked GDN state.
// Pages are statically owned by their slot, so a fork copies bytes; it is O(context), device to device, on the engine stream.
void crane_b200_model::fork_seq(int src, int dst) {
    seq_kv[cur] = kv_len; seq_pos[cur] = next_mrope_pos;
    if (src == cur) park_gdn_state(cur);
    const size_t len = seq_kv[src];
    const size_t pages = (len + KV_PAGE - 1) / KV_PAGE;
    const size_t page_elems = (size_t)nkv * KV_PAGE * D;
    const size_t so = (size_t)src * max_pages * page_elems, dofs = (size_t)dst * max_pages * page_elems, n = pages * page_elems;
    const size_t cs = hybrid ? (size_t)conv_dim() * ck : 0, rs = hybrid ? (size_t)nv * dk * dv : 0;
    for (auto& l : layers) {
        if (l.full && kv_bits) {
            if (!pages) continue;
            const size_t rows = pages * nkv * KV_PAGE, rs0 = (size_t)src * max_pages * nkv * KV_PAGE, rd0 = (size_t)dst * max_pages * nkv * KV_PAGE;
            const size_t cb = (size_t)D * kv_bits / 8;
            for (unsigned char* c : {l.k_codes, l.v_codes})
                CUDA_OK(cudaMemcpyAsync(c + rd0 * cb, c + rs0 * cb, rows * cb, cudaMemcpyDeviceToDevice, stream));
            for (float* sc : {l.k_scale, l.v_scale})
                CUDA_OK(cudaMemcpyAsync(sc + rd0, sc + rs0, rows * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        } else if (l.full) {
            if (!n) continue;
            for (bf16* pool : {l.k_pool, l.v_pool}) {
                CUDA_OK(cudaMemcpyAsync(pool + dofs, pool + so, n * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
                if (lo_kv) CUDA_OK(cudaMemcpyAsync(pool + lo_kv + dofs, pool + lo_kv + so, n * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
            }
        } else if (max_batch > 1) {
            CUDA_OK(cudaMemcpyAsync(l.conv_slots + dst * cs, l.conv_slots + src * cs, cs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
            CUDA_OK(cudaMemcpyAsync(l.rec_slots + dst * rs, l.rec_slots + src * rs, rs * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        }
    }
    seq_kv[dst] = len; seq_pos[dst] = seq_pos[src];
}

void crane_b200_model::reset_recurrent_state() {
    for (auto& l : layers)
        if (!l.full) {
            CUDA_OK(cudaMemsetAsync(l.conv_state, 0, (size_t)conv_dim() * ck * sizeof(float), stream));
            CUDA_OK(cudaMemsetAsync(l.rec_state, 0, (size_t)nv * dk * dv * sizeof(float), stream));
        }
}

void crane_b200_model::lm_head_last_row(const float* xrow, int advance, int B) {
    GemvArgs h = {};
    h.part_val = part_val; h.part_idx = part_idx; h.ticket = ticket; h.state = state; h.out_tokens = out_tokens; h.out_stride = out_cap;
    h.embed = hov.gather_set ? hov.gather : embed; h.x_next = x_dec; h.H = H; h.advance = advance;
    h.norm_out = hidden_out; h.force_tokens = hov.force;
    linear_decode(GEMV_LOGITS_ARGMAX, true, hov.head ? hov.head : lm_head, q_lm_head, hov.head ? 0 : qt_lm, V, H, xrow, H, final_norm,
                  hov.logits ? hov.logits : logits, V, &h, B);
    // a quantised table cannot be gathered by the head's epilogue: the next input is dequantised by its own launch
    if (advance && q_embed && !hov.gather_set) embed_step_input(B);
}

// One decode step (layers + lm_head), replayed from a CUDA graph when capture is available.
void crane_b200_model::decode_step_graphed(int advance, int B) {
    cudaGraphExec_t& ge = graph_step[advance][B];
    if (use_graphs && !graph_failed && ge == nullptr) {
        cudaGraph_t g = nullptr;
        const uint64_t before = launches;
        cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed);
        bool ok = (e == cudaSuccess);
        if (ok) {
            try { enqueue_decode_step(advance, false, B); } catch (const EngineError&) { ok = false; }
            e = cudaStreamEndCapture(stream, &g);
            ok = ok && e == cudaSuccess && g != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&ge, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        graph_launches[advance][B] = launches - before;     // kernels one replay stands for
        launches = before;
        if (!ok) { graph_failed = true; ge = nullptr; cudaGetLastError(); }
    }
    if (ge) {
        CUDA_OK(cudaGraphLaunch(ge, stream));
        launches += graph_launches[advance][B];
    } else {
        enqueue_decode_step(advance, false, B);
    }
}

// n dependent decode steps: one persistent launch when available, else n graph replays / kernel chains.
void crane_b200_model::decode_steps(int n_steps, int advance) {
    struct DecodePass {                          // one profiled pass per call: n_steps decode passes of one position each
        PassProfiler& p; int n; bool mine;
        DecodePass(PassProfiler& pp, cudaStream_t st, int n_) : p(pp), n(n_), mine(pp.on && !pp.active()) { if (mine) { p.begin(st); p.mark(SP_DECODE); } }
        ~DecodePass() { if (mine) p.end((size_t)n, 0, (uint64_t)n); }
    } decode_pass(spans, stream, n_steps);
    if (use_persistent && (advance == 1 || (advance == 0 && n_steps == 1))) {
        LLArgs p = {};
        p.L = L; p.H = H; p.I = I; p.V = V; p.nh = nh; p.nkv = nkv; p.qkv_dim = qkv_dim(); p.q_dim = q_dim();
        p.eps = eps; p.scale = 1.0f / std::sqrt((float)D);
        for (int i = 0; i < L; ++i) {
            const LayerW& l = layers[i];
            p.layers[i] = LLLayer{l.wqkv, l.wo, l.wgu, l.wdown, l.ln1, l.ln2, l.qn, l.kn, l.k_pool, l.v_pool};
        }
        p.lm_head = lm_head; p.final_norm = final_norm; p.embed = embed;
        p.cos_tab = cos_tab; p.sin_tab = sin_tab; p.axis_of = axis_of; p.state = state; p.block_table = block_table; p.max_pages = max_pages;
        p.kv_lo_off = lo_kv; p.x_io = x_dec; p.logits = logits; p.out_tokens = out_tokens;
        p.xa = ll_xa; p.xb = ll_xb; p.qkv = ll_qkv; p.att = ll_att; p.act = ll_act; p.part = ll_part; p.amax = ll_amax;
        const unsigned int need = ll_tags_per_launch(L, n_steps);
        if (ll_tag > 0xffffffffu - need - 1u) {       // tag space exhausted (once per ~10^8 tokens): start over from clean buffers
            for (auto* b : {ll_xa, ll_xb}) CUDA_OK(cudaMemsetAsync(b, 0, (size_t)H * 8, stream));
            CUDA_OK(cudaMemsetAsync(ll_qkv, 0, (size_t)qkv_dim() * 8, stream));
            CUDA_OK(cudaMemsetAsync(ll_att, 0, (size_t)q_dim() * 8, stream));
            CUDA_OK(cudaMemsetAsync(ll_act, 0, (size_t)I * 8, stream));
            CUDA_OK(cudaMemsetAsync(ll_part, 0, decode_ll_part_pairs(num_sms, nh, nkv) * 8, stream));
            CUDA_OK(cudaMemsetAsync(ll_amax, 0, (size_t)4 * num_sms * 8, stream));
            ll_tag = 0;
        }
        p.tag_base = ll_tag;
        ll_tag += need;
        p.n_steps = n_steps; p.advance = advance; p.err = ll_err; p.prof = ll_prof;
        { static const int la = [] { const char* e = getenv("CRANE_B200_LL_L2AHEAD"); return e ? atoi(e) : 0; }(); p.l2_ahead = la; }
        if (ll_trace && n_steps > 8) { CUDA_OK(cudaMemsetAsync(ll_trace, 0, (size_t)6 * LL_TRACE_CAP * 16, stream)); p.trace = ll_trace; p.trace_step = 6; }
        LAUNCH_OK(decode_ll_launch(stream, p, num_sms));
        ++launches;
        if (p.trace) {
            std::vector<unsigned long long> h((size_t)6 * LL_TRACE_CAP * 2);
            CUDA_OK(cudaStreamSynchronize(stream));
            CUDA_OK(cudaMemcpy(h.data(), ll_trace, h.size() * 8, cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(getenv("CRANE_B200_LL_TRACE"), "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
        }
        if (ll_prof && n_steps > 8) {   // CRANE_B200_PROF=1: where CTA 0's first thread spent its cycles (the CRANE_PROF spans of ops/prof.rs:37-61)
            unsigned long long h[16];
            CUDA_OK(cudaStreamSynchronize(stream));
            CUDA_OK(cudaMemcpy(h, ll_prof, 128, cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemset(ll_prof, 0, 128));
            const char* names[8] = {"init", "wait_x", "stream", "cta_barrier", "attention", "merge", "epilogue", "token"};
            fprintf(stderr, "[crane_b200 prof] cycles per step:");
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / n_steps);
            fprintf(stderr, "\n");
        }
        return;
    }
    for (int i = 0; i < n_steps; ++i) decode_step_graphed(advance);
}

// The persistent kernel bounds every wait; a wait that gave up leaves garbage behind and raises this flag (read at sync points).
void crane_b200_model::ll_err_fetch() {
    if (ll_err) CUDA_OK(cudaMemcpyAsync(h_ll_err, ll_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
}
void crane_b200_model::ll_err_verify() {
    if (!ll_err || *h_ll_err == 0) return;
    std::vector<unsigned int> d(16 + num_sms * LL_WARPS);
    CUDA_OK(cudaMemcpy(d.data(), ll_err, d.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemsetAsync(ll_err, 0, sizeof(unsigned int) * d.size(), stream));
    *h_ll_err = 0;
    static const char* sites[5] = {"?", "activation slice", "attention query", "attention split merge", "token argmax"};
    // the warps that were waiting furthest behind are where the chain broke
    unsigned int lo = ~0u;
    for (size_t i = 16; i < d.size(); ++i) if (d[i] != 0 && d[i] < lo) lo = d[i];
    std::string who;
    int shown = 0;
    for (size_t i = 16; i < d.size() && shown < 12; ++i)
        if (d[i] == lo) { who += " cta" + std::to_string((i - 16) / LL_WARPS) + ".w" + std::to_string((i - 16) % LL_WARPS); ++shown; }
    fail(CRANE_B200_CUDA_ERROR,
         "persistent decode kernel: a dependency wait timed out (results discarded): %s wait in CTA %u warp %u at step %u phase %u wanted tag %u, "
         "saw %u; earliest unfinished wait: step %u phase %u site %u in%s",
         sites[d[1] < 5 ? d[1] : 0], d[2], d[3], d[4], d[5], d[6], d[7], lo >> 20, (lo >> 8) & 0xfff, lo & 0xff, who.c_str());
}

// =================================================================================================
// prefill (S > 1) -- `Qwen3Model::decode` body for a chunk of S tokens
// =================================================================================================
void crane_b200_model::prefill(const uint32_t* ids, const float* embeds, size_t S_, const uint32_t* pos3_host, size_t start_pos,
                               const int* vis_rows, int n_vis, int advance) {
    const int S = (int)S_;
    ensure_prefill_ws(S);
    spans.begin(stream);                          // (a VL request began its pass before the vision tower)
    spans.mark(SP_EMBED);
    if (!pev0_armed) CUDA_OK(cudaEventRecord(pev0, stream));
    pev0_armed = false;
    // positions [3, S], ids and splice rows go through pinned staging: nothing here makes the host wait for the GPU (a VL request is
    // still running its vision tower at this point, and the text layers are enqueued underneath it)
    if (stage_busy) { CUDA_OK(cudaEventSynchronize(stage_ev)); stage_busy = false; }
    int* p3 = h_stage;                               // [3 S] | ids [S] | rows [n_vis]
    for (int a = 0; a < 3; ++a)
        for (int s = 0; s < S; ++s) p3[(size_t)a * S + s] = pos3_host ? (int)pos3_host[(size_t)a * S + s] : (int)(start_pos + s);
    for (size_t i = 0; i < (size_t)3 * S; ++i)
        if (p3[i] < 0 || p3[i] > max_seq) fail(CRANE_B200_INVALID_ARG, "rotary position %d outside the table (max_seq_len %d)", p3[i], max_seq);
    CUDA_OK(cudaMemcpyAsync(pos3_dev, p3, (size_t)3 * S * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (ids) {
        uint32_t* hid = reinterpret_cast<uint32_t*>(h_stage + (size_t)3 * S);
        std::memcpy(hid, ids, S * sizeof(uint32_t));
        CUDA_OK(cudaMemcpyAsync(ids_dev, hid, S * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        if (q_embed) LAUNCH_OK(embed_rows_q_launch(stream, qt_embed, q_embed, H, ids_dev, nullptr, S, x, prefill_pdl()));
        else LAUNCH_OK(embed_rows_launch(stream, ids_dev, S, embed, H, x));
        ++launches;
    } else {
        CUDA_OK(cudaMemcpyAsync(x, embeds, (size_t)S * H * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    if (n_vis > 0) {   // splice the image features over the placeholder rows (qwen3_5/vlm.rs:433-468)
        spans.mark(SP_SPLICE);
        if (n_vis > S) fail(CRANE_B200_INVALID_ARG, "more image rows (%d) than positions (%d)", n_vis, S);
        int* hr = h_stage + (size_t)4 * S;
        std::memcpy(hr, vis_rows, n_vis * sizeof(int));
        CUDA_OK(cudaMemcpyAsync(rows_dev, hr, n_vis * sizeof(int), cudaMemcpyHostToDevice, stream));
        LAUNCH_OK(set_rows_launch(stream, x, H, rows_dev, n_vis, img_embeds, false));
        ++launches;
    }
    CUDA_OK(cudaEventRecord(stage_ev, stream));
    stage_busy = true;
    if (!ids) CUDA_OK(cudaStreamSynchronize(stream));   // caller-owned embeddings were copied from pageable memory

    const int qd = q_dim();
    for (int li = 0; li < L; ++li) {
        LayerW& l = layers[li];
        const bool q_qkv = l.full && l.qt_q != 0;
        spans.mark(SP_NORM);
        if (!q_qkv) LAUNCH_OK(rmsnorm_rows_launch(stream, x, S, H, l.ln1, eps, xn, lo_xn));
        if (l.full) {
            spans.mark(SP_ATTN_QKV);
            if (q_qkv) {   // GGUF keeps q / k / v separate and quantised (qwen3/modeling.rs:252-255): integer-dot rows, 4 at a time
                const int qs = nh * q_stride(), kvd = nkv * D;
                for (int s0 = 0; s0 < S;) {
                    const int B = (S - s0 >= 4) ? 4 : (S - s0 >= 2) ? 2 : 1;
                    const float* xr = x + (size_t)s0 * H;
                    float* yr = qkv + (size_t)s0 * qkv_dim();
                    linear_decode(GEMV_STORE, true, nullptr, l.q_wq, l.qt_q, qs, H, xr, H, l.ln1, yr, qkv_dim(), nullptr, B);
                    linear_decode(GEMV_STORE, true, nullptr, l.q_wk, l.qt_k, kvd, H, xr, H, l.ln1, yr + qs, qkv_dim(), nullptr, B, true);
                    linear_decode(GEMV_STORE, true, nullptr, l.q_wv, l.qt_v, kvd, H, xr, H, l.ln1, yr + qs + kvd, qkv_dim(), nullptr, B, true);
                    s0 += B;
                }
            } else {
                gemm(xn, lo_xn, H, l.wqkv, S, qkv_dim(), H, EPI_STORE_F32, qkv, qkv_dim(), nullptr);
            }
            spans.mark(SP_ATTN_ROPE);
            RopeAppendArgs ra = {};
            ra.qkv = qkv; ra.q_stride = q_stride(); ra.rot_half = rot_half;
            ra.q_norm_w = l.qn; ra.k_norm_w = l.kn; ra.eps = eps; ra.cos_tab = cos_tab; ra.sin_tab = sin_tab; ra.axis_of = axis_of;
            // quantised pages: the prefix is dequantised into the scratch, the new rows go through it (prefill.cu, KvQuantArgs)
            bf16 *kp = kv_bits ? kq_sk : l.k_pool, *vp = kv_bits ? kq_sv : l.v_pool;
            const int* btp = kv_bits ? bt_identity : bt_cur();
            if (kv_bits && start_pos > 0) { LAUNCH_OK(kv_dequant_pages_launch(stream, kvq_args(l), (int)start_pos)); ++launches; }
            ra.pos3 = pos3_dev; ra.S = S; ra.start_pos = (int)start_pos; ra.block_table = btp; ra.k_pool = kp; ra.v_pool = vp;
            ra.nh = nh; ra.nkv = nkv; ra.q_out = q_bf; ra.q_lo_off = lo_q; ra.kv_lo_off = lo_kv;
            ra.kv_bits = kv_bits; ra.k_codes = l.k_codes; ra.v_codes = l.v_codes; ra.k_scale = l.k_scale; ra.v_scale = l.v_scale; ra.code_bt = bt_cur();
            LAUNCH_OK(rope_append_launch(stream, D, ra));
            spans.mark(SP_ATTN_FLASH);
            FlashArgs fa = {};
            fa.q = q_bf; fa.q_stride = qd; fa.k_pool = kp; fa.v_pool = vp; fa.block_table = btp; fa.nh = nh; fa.nkv = nkv;
            fa.out = attn_bf; fa.o_stride = qd; fa.S = S; fa.kv_offset = (int)start_pos; fa.scale = 1.0f / std::sqrt((float)D); fa.nseq = 1;
            fa.q_lo_off = lo_q; fa.kv_lo_off = lo_kv; fa.out_lo_off = lo_attn;
            LAUNCH_OK(flash_prefill_launch(stream, D, true, true, fa));
            if (hybrid) { LAUNCH_OK(gate_mul_launch(stream, attn_bf, qkv, S, nh, D, q_stride(), qkv_dim(), lo_attn)); ++launches; }
            spans.mark(SP_ATTN_O);
            if (l.qt_o) {
                LAUNCH_OK(planes_to_f32_launch(stream, attn_bf, lo_attn, (size_t)S * qd, rows_f32));
                ++launches;
                qlinear_rows(GEMV_RESID, l.q_wo, l.qt_o, H, qd, rows_f32, qd, nullptr, x, H, S);
            } else {
                gemm(attn_bf, lo_attn, qd, l.wo, S, H, qd, EPI_RESID_F32, x, H, nullptr);
            }
        } else {
            spans.mark(SP_GDN_PROJ);
            gemm(xn, lo_xn, H, l.w_in, S, gdn_in_pad, H, EPI_STORE_F32, g_proj, gdn_in_pad, nullptr);
            GdnArgs ga;
            gdn_args(ga, l, S, g_proj, g_conv, g_qn, g_kn, g_gb, g_y);
            ga.out_bf16 = attn_bf; ga.out_lo_off = lo_attn;
            // "auto": chunkwise from two chunks on (at one chunk its three launches cost what 64 sequential steps do); "chunked": from one
            if (g_chunk && S >= (gdn_mode == 2 ? cb::GDN_CHUNK : 2 * cb::GDN_CHUNK)) { ga.glog = g_gl; ga.chunk_ws = g_chunk; }
            LAUNCH_OK(gdn_forward_launch(stream, ga));      // marks conv / qkv / recur / finish itself
            gemm(attn_bf, lo_attn, value_dim(), l.w_out, S, H, value_dim(), EPI_RESID_F32, x, H, nullptr);
            launches += gdn_forward_launch_count(ga);     // conv (+ norms + gates), conv state, recurrence (1 kernel, or 3 chunkwise), gated norm
        }
        // MLP: either linear may be quantised on its own (Q4_K_M keeps ffn_down in Q6_K, the others in Q4_K)
        spans.mark(l.qt_gu ? SP_MLP_GATE_UP : SP_NORM);
        if (l.qt_gu) {
            qlinear_rows(GEMV_SILU_MUL, l.q_wgu, l.qt_gu, 2 * I, H, x, H, l.ln2, rows_f32, I, S);
            if (!l.qt_down) { LAUNCH_OK(cast_f32_bf16_launch(stream, rows_f32, act_bf, (size_t)S * I, lo_act)); ++launches; }
        } else {
            LAUNCH_OK(rmsnorm_rows_launch(stream, x, S, H, l.ln2, eps, xn, lo_xn));
            spans.mark(SP_MLP_GATE_UP);
            gemm(xn, lo_xn, H, l.wgu, S, 2 * I, H, EPI_SILU_MUL_BF16, act_bf, I, nullptr, lo_act);
            if (l.qt_down) { LAUNCH_OK(planes_to_f32_launch(stream, act_bf, lo_act, (size_t)S * I, rows_f32)); ++launches; }
        }
        spans.mark(SP_MLP_DOWN);
        if (l.qt_down) qlinear_rows(GEMV_RESID, l.q_wdown, l.qt_down, H, I, rows_f32, I, nullptr, x, H, S);
        else gemm(act_bf, lo_act, I, l.wdown, S, H, I, EPI_RESID_F32, x, H, nullptr);
        launches += 4;
        if (n_vis > 0 && li < (int)v_deepstack.size()) {   // DeepStack (qwen3_vl/text.rs:262-268)
            spans.mark(SP_SPLICE);
            LAUNCH_OK(set_rows_launch(stream, x, H, rows_dev, n_vis, ds_embeds + (size_t)li * n_vis * H, true));
            ++launches;
        }
    }
    // state for the last-row lm_head / a following on-device decode loop
    spans.mark(SP_HEAD);
    stage_state();
    h_state[0].kv_len = (int)(start_pos + S - 1);   // lm_head(advance) bumps it to start_pos + S
    h_state[0].pos[0] = h_state[0].pos[1] = h_state[0].pos[2] = (int)next_mrope_pos - 1;
    h_state[0].token = 0;
    h_state[0].step = 0;
    h_state[0].slot = cur;
    push_state(1);
    lm_head_last_row(x + (size_t)(S - 1) * H, advance);
    CUDA_OK(cudaEventRecord(pev1, stream));
    kv_len = start_pos + S;
    spans.end(S_, 1);
}

// =================================================================================================
// vision tower
// =================================================================================================
void crane_b200_model::ensure_vision_ws(int N) {
    if (N <= vws_N) return;
    CUDA_OK(cudaStreamSynchronize(stream));
    for (void* p : {(void*)v_x, (void*)v_qkv, (void*)v_pv, (void*)v_cos, (void*)v_sin, (void*)v_w4, (void*)v_pvb, (void*)v_xn, (void*)v_qkvb,
                    (void*)v_attn, (void*)v_act, (void*)v_m1, (void*)v_idx4, (void*)v_seq_start, (void*)v_seq_len, (void*)img_embeds,
                    (void*)ds_embeds})
        dfree(p);
    v_tab_key.clear();
    const int cap = (N + 127) / 128 * 128;
    const int pk = v_in * v_tpatch * v_patch * v_patch, mh = v_H * v_merge * v_merge, m2 = v_merge * v_merge;
    v_x = dalloc<float>((size_t)cap * v_H);
    v_qkv = dalloc<float>((size_t)cap * 3 * v_H);
    v_pv = dalloc<float>((size_t)cap * pk);
    v_cos = dalloc<float>((size_t)cap * (v_hd / 2));
    v_sin = dalloc<float>((size_t)cap * (v_hd / 2));
    v_w4 = dalloc<float>((size_t)4 * cap);
    v_pvb = dalloc_act((size_t)cap * pk, lo_vpvb);
    v_xn = dalloc_act((size_t)cap * v_H, lo_vxn);
    v_qkvb = dalloc_act((size_t)cap * 3 * v_H, lo_vqkvb);
    v_attn = dalloc_act((size_t)cap * v_H, lo_vattn);
    v_act = dalloc_act((size_t)cap * v_I, lo_vact);
    v_m1 = dalloc_act((size_t)(cap / m2) * mh, lo_vm1);
    v_idx4 = dalloc<int>((size_t)4 * cap);
    v_seq_start = dalloc<int>(cap);
    v_seq_len = dalloc<int>(cap);
    img_embeds = dalloc<float>((size_t)(cap / m2) * v_out);
    ds_embeds = dalloc<float>((size_t)std::max<size_t>(v_deepstack.size(), 1) * (cap / m2) * v_out);
    vws_N = cap;
}

void crane_b200_model::encode_images(const float* pv, const uint32_t* grid, size_t n_images) {
    if (!is_vl) fail(CRANE_B200_UNSUPPORTED, "encode_images on a text-only model");
    const int pk = v_in * v_tpatch * v_patch * v_patch, mh = v_H * v_merge * v_merge, m2 = v_merge * v_merge, half = v_hd / 2;
    int N = 0;
    for (size_t i = 0; i < n_images; ++i) {
        const int t = grid[3 * i], h = grid[3 * i + 1], w = grid[3 * i + 2];
        if (t < 1 || h < 1 || w < 1 || h % v_merge || w % v_merge) fail(CRANE_B200_INVALID_ARG, "bad grid_thw for image %zu", i);
        N += t * h * w;
    }
    if (N == 0) fail(CRANE_B200_INVALID_ARG, "no patches");
    ensure_vision_ws(N);
    spans.begin(stream);
    spans.mark(SP_VIT_STAGE);
    // ---- host index arithmetic: bilinear pos-embed corners (vision.rs:382-489), 2-D rotary table (:491-541),
    //      per-frame sequence bounds (:543-556) ----
    // The tables depend on the grids only: a request with the grids of the previous one (every request of a fixed-resolution
    // pipeline) reuses what is already on the device and pays neither the ~0.4 ms of host trigonometry nor the uploads.
    const std::vector<uint32_t> grid_key(grid, grid + 3 * n_images);
    const bool tables_cached = grid_key == v_tab_key;
    std::vector<int> idx4(tables_cached ? 0 : (size_t)4 * N), sstart, slen;
    std::vector<float> w4(tables_cached ? 0 : (size_t)4 * N), cs(tables_cached ? 0 : (size_t)N * half), sn(tables_cached ? 0 : (size_t)N * half);
    const int qdim = half / 2;   // rotary_pos_emb dim = head_dim/2 -> head_dim/4 frequencies per axis
    std::vector<float> inv(qdim);
    for (int i = 0; i < qdim; ++i) inv[i] = 1.0f / powf(10000.0f, (float)(2 * i) / (float)half);
    int base = 0;
    for (size_t im = 0; im < n_images && !tables_cached; ++im) {
        const int t = grid[3 * im], h = grid[3 * im + 1], w = grid[3 * im + 2];
        auto lin = [&](int steps, std::vector<float>& out) {
            out.resize(steps);
            if (steps == 1) { out[0] = 0.f; return; }
            const float step = (float)(v_side - 1) / (float)(steps - 1);
            for (int i = 0; i < steps; ++i) out[i] = (float)i * step;
        };
        std::vector<float> hv, wv;
        lin(h, hv); lin(w, wv);
        int p = base;
        for (int f = 0; f < t; ++f) {
            sstart.push_back(base + f * h * w);
            slen.push_back(h * w);
            for (int br = 0; br < h / v_merge; ++br)
                for (int bc = 0; bc < w / v_merge; ++bc)
                    for (int ir = 0; ir < v_merge; ++ir)
                        for (int ic = 0; ic < v_merge; ++ic, ++p) {
                            const int r = br * v_merge + ir, c = bc * v_merge + ic;
                            const int hf = (int)floorf(hv[r]), wf = (int)floorf(wv[c]);
                            const int hc = std::min((int)ceilf(hv[r]), v_side - 1), wc = std::min((int)ceilf(wv[c]), v_side - 1);
                            const float dh = hv[r] - (float)hf, dw = wv[c] - (float)wf;
                            idx4[(size_t)0 * N + p] = hf * v_side + wf; w4[(size_t)0 * N + p] = (1.f - dh) * (1.f - dw);
                            idx4[(size_t)1 * N + p] = hf * v_side + wc; w4[(size_t)1 * N + p] = (1.f - dh) * dw;
                            idx4[(size_t)2 * N + p] = hc * v_side + wf; w4[(size_t)2 * N + p] = dh * (1.f - dw);
                            idx4[(size_t)3 * N + p] = hc * v_side + wc; w4[(size_t)3 * N + p] = dh * dw;
                            for (int i = 0; i < qdim; ++i) {
                                const float fr = (float)r * inv[i], fc = (float)c * inv[i];
                                cs[(size_t)p * half + i] = cosf(fr); sn[(size_t)p * half + i] = sinf(fr);
                                cs[(size_t)p * half + qdim + i] = cosf(fc); sn[(size_t)p * half + qdim + i] = sinf(fc);
                            }
                        }
        }
        base += t * h * w;
    }
    if (!tables_cached) {
        v_tab_nseq = (int)sstart.size();
        v_tab_max_len = 0;
        for (int v : slen) v_tab_max_len = std::max(v_tab_max_len, v);
    }
    const int nseq = v_tab_nseq, max_len = v_tab_max_len;
    CUDA_OK(cudaMemcpyAsync(v_pv, pv, (size_t)N * pk * sizeof(float), cudaMemcpyHostToDevice, stream));
    if (!tables_cached) {
        v_tab_key.clear();
        CUDA_OK(cudaMemcpyAsync(v_idx4, idx4.data(), idx4.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(v_w4, w4.data(), w4.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(v_cos, cs.data(), cs.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(v_sin, sn.data(), sn.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(v_seq_start, sstart.data(), nseq * sizeof(int), cudaMemcpyHostToDevice, stream));
        CUDA_OK(cudaMemcpyAsync(v_seq_len, slen.data(), nseq * sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    if (!pix_ev) CUDA_OK(cudaEventCreateWithFlags(&pix_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(pix_ev, stream));
    LAUNCH_OK(cast_f32_bf16_launch(stream, v_pv, v_pvb, (size_t)N * pk, lo_vpvb));
    if (!tables_cached) {
        CUDA_OK(cudaStreamSynchronize(stream));   // host staging vectors go out of scope below
        v_tab_key = grid_key;                     // (only now: a failed upload must not leave a key behind)
    }
    // (cached tables: the caller's pixel buffer may still be being read -- vl_forward waits for pix_ev before it returns)

    // patch embed: Conv3d(kernel == stride) == GEMM [N, C*T*P*P] x [Hv, C*T*P*P]^T + bias (vision.rs:46-58)
    spans.mark(SP_VIT_PATCH);
    gemm(v_pvb, lo_vpvb, pk, v_wpatch, N, v_H, pk, EPI_STORE_F32, v_x, v_H, v_bpatch);
    LAUNCH_OK(vit_pos_embed_add_launch(stream, v_x, N, v_H, v_pos, v_idx4, v_w4));
    launches += 2;
    const int Ng = N / m2;
    auto run_merger = [&](MergerW& m, float* out) {
        spans.mark(SP_VIT_MERGER);
        if (m.post) LAUNCH_OK(layernorm_rows_launch(stream, v_x, Ng, mh, m.nw, m.nb, 1e-6f, v_xn, lo_vxn));
        else LAUNCH_OK(layernorm_rows_launch(stream, v_x, N, v_H, m.nw, m.nb, 1e-6f, v_xn, lo_vxn));
        gemm(v_xn, lo_vxn, mh, m.w1, Ng, mh, mh, merger_gelu_mode, v_m1, mh, m.b1, lo_vm1);
        gemm(v_m1, lo_vm1, mh, m.w2, Ng, v_out, mh, EPI_STORE_F32, out, v_out, m.b2);
        ++launches;
    };
    for (int bi = 0; bi < v_depth; ++bi) {
        VitBlockW& b = vblocks[bi];
        spans.mark(SP_VIT_NORM);
        LAUNCH_OK(layernorm_rows_launch(stream, v_x, N, v_H, b.n1w, b.n1b, 1e-6f, v_xn, lo_vxn));
        spans.mark(SP_VIT_QKV);
        gemm(v_xn, lo_vxn, v_H, b.wqkv, N, 3 * v_H, v_H, EPI_STORE_F32, v_qkv, 3 * v_H, b.bqkv);
        spans.mark(SP_VIT_ROPE);
        LAUNCH_OK(vit_rope_launch(stream, v_qkv, N, v_nh, v_hd, v_cos, v_sin, v_qkvb, lo_vqkvb));
        FlashArgs fa = {};
        fa.q = v_qkvb; fa.q_stride = 3 * v_H; fa.k = v_qkvb + v_H; fa.v = v_qkvb + 2 * v_H; fa.kv_stride = 3 * v_H;
        fa.nh = v_nh; fa.nkv = v_nh; fa.out = v_attn; fa.o_stride = v_H; fa.seq_start = v_seq_start; fa.seq_len = v_seq_len;
        fa.scale = 1.0f / std::sqrt((float)v_hd); fa.nseq = nseq; fa.max_len = max_len;
        fa.q_lo_off = lo_vqkvb; fa.kv_lo_off = lo_vqkvb; fa.out_lo_off = lo_vattn;
        spans.mark(SP_VIT_FLASH);
        LAUNCH_OK(flash_prefill_launch(stream, v_hd, false, false, fa));
        spans.mark(SP_VIT_PROJ);
        gemm(v_attn, lo_vattn, v_H, b.wproj, N, v_H, v_H, EPI_RESID_F32, v_x, v_H, b.bproj);
        spans.mark(SP_VIT_NORM);
        LAUNCH_OK(layernorm_rows_launch(stream, v_x, N, v_H, b.n2w, b.n2b, 1e-6f, v_xn, lo_vxn));
        spans.mark(SP_VIT_FC1);
        gemm(v_xn, lo_vxn, v_H, b.wfc1, N, v_I, v_H, vit_gelu_mode, v_act, v_I, b.bfc1, lo_vact);
        spans.mark(SP_VIT_FC2);
        gemm(v_act, lo_vact, v_I, b.wfc2, N, v_H, v_I, EPI_RESID_F32, v_x, v_H, b.bfc2);
        launches += 4;
        for (size_t j = 0; j < v_deepstack.size(); ++j)
            if (v_deepstack[j] == bi) run_merger(v_ds_mergers[j], ds_embeds + j * (size_t)Ng * v_out);
    }
    run_merger(v_merger, img_embeds);
    img_tokens = Ng;
    spans.mark(SP_OTHER);                          // the pass goes on into prefill(); a bare encode_images call closes it itself
}

// =================================================================================================
// C ABI
// =================================================================================================
#define API_BEGIN(m)                                   \
    if (!(m)) return CRANE_B200_INVALID_ARG;           \
    try {                                              \
        cudaSetDevice((m)->device);
#define API_END(m)                                     \
    }                                                  \
    catch (const EngineError& e) {                     \
        (m)->last_error = e.msg;                       \
        return e.code;                                 \
    }                                                  \
    catch (const std::exception& e) {                  \
        (m)->last_error = e.what();                    \
        return CRANE_B200_INVALID_ARG;                 \
    }                                                  \
    return CRANE_B200_OK;

static void need_ready(crane_b200_model* m) {
    if (!m->finalized) fail(CRANE_B200_NOT_LOADED, "model is not finalized");
}
static void fill_logits(crane_b200_model* m, crane_b200_logits* out) {
    if (out) { out->device_ptr = m->logits; out->rows = 1; out->vocab = (size_t)m->V; out->stream = (void*)m->stream; }
}
static void check_room(crane_b200_model* m, size_t start_pos, size_t n) {
    if (start_pos != m->kv_len) fail(CRANE_B200_INVALID_ARG, "start_pos %zu != cached length %zu", start_pos, m->kv_len);
    if (start_pos + n > (size_t)m->max_seq) fail(CRANE_B200_OOM, "sequence length %zu exceeds max_seq_len %d", start_pos + n, m->max_seq);
}

extern "C" {

int crane_b200_create(const char* config_json, int device_ordinal, crane_b200_model** out) {
    if (!config_json || !out) { g_create_error = "null argument"; return CRANE_B200_INVALID_ARG; }
    *out = nullptr;
    std::unique_ptr<crane_b200_model> m(new crane_b200_model());
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail(CRANE_B200_CUDA_ERROR, "no CUDA device: %s (crane_b200 has no CPU fallback)", cudaGetErrorString(e));
        if (device_ordinal < 0 || device_ordinal >= ndev) fail(CRANE_B200_INVALID_ARG, "device ordinal %d of %d", device_ordinal, ndev);
        m->device = device_ordinal;
        CUDA_OK(cudaSetDevice(device_ordinal));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device_ordinal));
        if (prop.major != 10) fail(CRANE_B200_UNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
        m->num_sms = prop.multiProcessorCount;
        m->parse_config(config_json);
        CUDA_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->alloc_weights();
        if (m->is_tts) {                 // code predictor = a second dense decoder on the same stream
            std::unique_ptr<crane_b200_model> c(new crane_b200_model());
            c->device = device_ordinal; c->num_sms = m->num_sms;
            c->parse_config(m->cp_json.c_str());
            c->split = m->split;
            c->stream = m->stream; c->owns_stream = false;
            c->alloc_weights();
            c->got_embed = true;         // its inputs come from the codec embedding tables, not embed_tokens
            m->cp = c.release();
        }
    } catch (const EngineError& e) {
        g_create_error = e.msg;
        for (void* p : m->allocs) cudaFree(p);
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        for (void* p : m->allocs) cudaFree(p);
        return CRANE_B200_INVALID_ARG;
    }
    *out = m.release();
    return CRANE_B200_OK;
}

static void comm_release(crane_b200_model* m);

void crane_b200_destroy(crane_b200_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    comm_release(m);
    if (m->cp) { crane_b200_destroy(m->cp); m->cp = nullptr; }
    for (auto& ga : m->graph_step) for (auto& g : ga) if (g) cudaGraphExecDestroy(g);
    for (void* p : m->allocs) cudaFree(p);
    if (m->h_state_ring) cudaFreeHost(m->h_state_ring);
    for (cudaEvent_t e : m->h_state_ev) if (e) cudaEventDestroy(e);
    if (m->h_tokens) cudaFreeHost(m->h_tokens);
    if (m->h_ll_err) cudaFreeHost(m->h_ll_err);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->stage_ev) cudaEventDestroy(m->stage_ev);
    if (m->pix_ev) cudaEventDestroy(m->pix_ev);
    m->sampler.release();
    m->spans.release();
    for (cudaEvent_t e : {m->pev0, m->pev1, m->dev0, m->dev1}) if (e) cudaEventDestroy(e);
    if (m->stream) gemm_release_stream(m->stream);
    if (m->stream && m->owns_stream) cudaStreamDestroy(m->stream);
    delete m;
}

const char* crane_b200_last_error(const crane_b200_model* m) { return m ? m->last_error.c_str() : g_create_error.c_str(); }

int crane_b200_load_tensor(crane_b200_model* m, const char* name, int dtype, const int64_t* shape, int ndim, const void* data) {
    API_BEGIN(m)
    if (!name || !shape || !data || ndim < 1 || ndim > 8) fail(CRANE_B200_INVALID_ARG, "load_tensor: bad arguments");
    m->load_tensor(name, dtype, shape, ndim, data);
    API_END(m)
}

int crane_b200_load_tensor_ggml(crane_b200_model* m, const char* name, int ggml_type, const int64_t* shape, int ndim, const void* data,
                                size_t nbytes) {
    API_BEGIN(m)
    if (!name || !shape || !data) fail(CRANE_B200_INVALID_ARG, "load_tensor_ggml: bad arguments");
    m->load_tensor_ggml(name, ggml_type, shape, ndim, data, nbytes);
    API_END(m)
}

int crane_b200_finalize(crane_b200_model* m) {
    API_BEGIN(m)
    m->finalize();
    API_END(m)
}

int crane_b200_forward_step(crane_b200_model* m, const uint32_t* ids, size_t n, size_t start_pos, crane_b200_logits* out) {
    API_BEGIN(m)
    need_ready(m);
    if (!ids || n == 0) fail(CRANE_B200_INVALID_ARG, "forward_step: empty input");
    check_room(m, start_pos, n);
    for (size_t i = 0; i < n; ++i)
        if (ids[i] >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", ids[i], m->V);
    if (n == 1) {
        CUDA_OK(cudaEventRecord(m->dev0, m->stream));
        m->arm_state(ids[0], start_pos, (int)start_pos, (int)start_pos, (int)start_pos);
        m->embed_step_input(1);
        m->decode_steps(1, 0);
        CUDA_OK(cudaEventRecord(m->dev1, m->stream));
        m->last_decode_steps = 1;
        m->kv_len = start_pos + 1;
        m->next_mrope_pos = (uint32_t)(start_pos + 1);
    } else {
        m->next_mrope_pos = (uint32_t)(start_pos + n);
        m->prefill(ids, nullptr, n, nullptr, start_pos, nullptr, 0, 0);
    }
    fill_logits(m, out);
    API_END(m)
}

int crane_b200_forward_step_argmax(crane_b200_model* m, const uint32_t* ids, size_t n, size_t start_pos, uint32_t* token_out) {
    if (!token_out) return CRANE_B200_INVALID_ARG;
    int r = crane_b200_forward_step(m, ids, n, start_pos, nullptr);
    if (r != CRANE_B200_OK) return r;
    API_BEGIN(m)
    CUDA_OK(cudaMemcpyAsync(m->h_tokens, m->out_tokens, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    m->ll_err_fetch();
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->ll_err_verify();
    *token_out = m->h_tokens[0];
    API_END(m)
}

int crane_b200_clear_kv_cache(crane_b200_model* m) {
    API_BEGIN(m)
    m->kv_len = 0;
    m->next_mrope_pos = 0;
    if (m->finalized && m->hybrid) m->reset_recurrent_state();
    API_END(m)
}

int crane_b200_num_layers(const crane_b200_model* m) { return m ? m->L : 0; }
int crane_b200_vocab_size(const crane_b200_model* m) { return m ? m->V : 0; }
int crane_b200_hidden_size(const crane_b200_model* m) { return m ? m->H : 0; }
size_t crane_b200_kv_len(const crane_b200_model* m) { return m ? m->kv_len : 0; }
uint32_t crane_b200_next_mrope_pos(const crane_b200_model* m) { return m ? m->next_mrope_pos : 0; }
uint64_t crane_b200_active_kv_cache_bytes(const crane_b200_model* m) {
    if (!m) return 0;
    uint64_t full = 0;
    for (int f : m->layer_is_full) full += f;
    const uint64_t per_row = m->kv_bits ? (uint64_t)m->D * m->kv_bits / 8 + 4 /*codes + f32 scale*/ : (uint64_t)m->D * (m->split ? 4 : 2) /*bf16 (+ lo plane)*/;
    return (uint64_t)m->kv_len * m->nkv * 2 /*K,V*/ * per_row * full +
           (m->hybrid ? (uint64_t)(m->L - full) * ((uint64_t)m->nv * m->dk * m->dv + (uint64_t)m->conv_dim() * m->ck) * 4 : 0);
}
uint64_t crane_b200_kernel_launches(const crane_b200_model* m) { return m ? m->launches : 0; }
int crane_b200_prof_enable(crane_b200_model* m, int on) {
    if (!m) return CRANE_B200_INVALID_ARG;
    m->spans.on = on != 0;
    if (!on) { m->spans.window[0] = m->spans.window[1] = m->spans.total[0] = m->spans.total[1] = cb::PassTotals{}; }
    return CRANE_B200_OK;
}
int crane_b200_prof_report(const crane_b200_model* m, char* buf, size_t cap, size_t* needed) {
    if (!m || !needed) return CRANE_B200_INVALID_ARG;
    const std::string s = cb::PassProfiler::json(m->spans.total);
    *needed = s.size() + 1;
    if (buf && cap >= s.size() + 1) memcpy(buf, s.c_str(), s.size() + 1);
    return CRANE_B200_OK;
}
int crane_b200_decode_path(const crane_b200_model* m) { return m && m->finalized && m->use_persistent ? 1 : 0; }

int crane_b200_warmup(crane_b200_model* m) {
    API_BEGIN(m)
    need_ready(m);
    const size_t saved = m->kv_len;
    const uint32_t saved_pos = m->next_mrope_pos;
    if (saved != 0) fail(CRANE_B200_INVALID_ARG, "warmup requires an empty KV cache");
    uint32_t ids[4] = {0, 1 % (uint32_t)m->V, 2 % (uint32_t)m->V, 3 % (uint32_t)m->V};
    m->prefill(ids, nullptr, 4, nullptr, 0, nullptr, 0, 1);
    m->decode_steps(2, 1);
    m->decode_steps(1, 0);
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->kv_len = 0;
    m->next_mrope_pos = saved_pos;
    if (m->hybrid) m->reset_recurrent_state();
    API_END(m)
}

int crane_b200_copy_logits(crane_b200_model* m, float* host_out, size_t n_floats) {
    API_BEGIN(m)
    need_ready(m);
    if (!host_out || n_floats > (size_t)m->V) fail(CRANE_B200_INVALID_ARG, "copy_logits: bad arguments");
    CUDA_OK(cudaMemcpyAsync(host_out, m->logits, n_floats * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    m->ll_err_fetch();
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->ll_err_verify();
    API_END(m)
}

int crane_b200_forward_embeds(crane_b200_model* m, const float* embeds, size_t s, const uint32_t* pos3, size_t start_pos,
                              crane_b200_logits* out) {
    API_BEGIN(m)
    need_ready(m);
    if (!embeds || s == 0) fail(CRANE_B200_INVALID_ARG, "forward_embeds: empty input");
    check_room(m, start_pos, s);
    uint32_t nxt = (uint32_t)(start_pos + s);
    if (pos3) {
        nxt = 0;
        for (size_t i = 0; i < 3 * s; ++i) nxt = std::max(nxt, pos3[i] + 1);
    }
    m->next_mrope_pos = nxt;
    // the single-row case also goes through the prefill kernels: embeddings come from the caller, not the table
    m->prefill(nullptr, embeds, s, pos3, start_pos, nullptr, 0, 0);
    fill_logits(m, out);
    API_END(m)
}

int crane_b200_decode_greedy(crane_b200_model* m, uint32_t first_token, size_t start_pos, size_t n_steps, const uint32_t* eos_ids,
                             size_t n_eos, uint32_t* tokens_out, size_t* n_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!tokens_out || !n_out || n_steps == 0) fail(CRANE_B200_INVALID_ARG, "decode_greedy: bad arguments");
    if (n_steps > (size_t)m->out_cap) fail(CRANE_B200_INVALID_ARG, "decode_greedy: n_steps %zu > %d", n_steps, m->out_cap);
    if (first_token >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", first_token, m->V);
    check_room(m, start_pos, n_steps);
    CUDA_OK(cudaEventRecord(m->dev0, m->stream));
    const int p = (int)m->next_mrope_pos;
    m->arm_state(first_token, start_pos, p, p, p);
    m->embed_step_input(1);
    m->decode_steps((int)n_steps, 1);
    CUDA_OK(cudaEventRecord(m->dev1, m->stream));
    CUDA_OK(cudaMemcpyAsync(m->h_tokens, m->out_tokens, n_steps * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    m->ll_err_fetch();
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->ll_err_verify();
    m->last_decode_steps = n_steps;
    m->kv_len = start_pos + n_steps;
    m->next_mrope_pos += (uint32_t)n_steps;
    size_t cnt = 0;
    for (size_t i = 0; i < n_steps; ++i) {
        tokens_out[cnt++] = m->h_tokens[i];
        bool stop = false;
        for (size_t e = 0; e < n_eos; ++e) stop = stop || (eos_ids && m->h_tokens[i] == eos_ids[e]);
        if (stop) break;
    }
    *n_out = cnt;
    API_END(m)
}

int crane_b200_generate_greedy(crane_b200_model* m, const uint32_t* prompt, size_t n_prompt, size_t max_new_tokens, const uint32_t* eos_ids,
                               size_t n_eos, uint32_t* tokens_out, size_t* n_out) {
    if (!m || !prompt || n_prompt == 0 || !tokens_out || !n_out || max_new_tokens == 0) return CRANE_B200_INVALID_ARG;
    int r = crane_b200_clear_kv_cache(m);
    if (r) return r;
    uint32_t tok = 0;
    r = crane_b200_forward_step_argmax(m, prompt, n_prompt, 0, &tok);
    if (r) return r;
    tokens_out[0] = tok;
    *n_out = 1;
    for (size_t e = 0; e < n_eos; ++e)
        if (eos_ids && tok == eos_ids[e]) return CRANE_B200_OK;
    if (max_new_tokens == 1) return CRANE_B200_OK;
    size_t got = 0;
    r = crane_b200_decode_greedy(m, tok, n_prompt, max_new_tokens - 1, eos_ids, n_eos, tokens_out + 1, &got);
    if (r) return r;
    *n_out = 1 + got;
    return CRANE_B200_OK;
}

int crane_b200_encode_images(crane_b200_model* m, const float* pixel_values, const uint32_t* grid_thw, size_t n_images, float* image_embeds_out,
                             float* deepstack_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!pixel_values || !grid_thw || n_images == 0) fail(CRANE_B200_INVALID_ARG, "encode_images: bad arguments");
    m->encode_images(pixel_values, grid_thw, n_images);
    m->spans.end((size_t)m->img_tokens * m->v_merge * m->v_merge, 1);
    const size_t n = (size_t)m->img_tokens * m->v_out;
    if (image_embeds_out) CUDA_OK(cudaMemcpyAsync(image_embeds_out, m->img_embeds, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    if (deepstack_out && !m->v_deepstack.empty())
        CUDA_OK(cudaMemcpyAsync(deepstack_out, m->ds_embeds, m->v_deepstack.size() * n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    API_END(m)
}

int crane_b200_vl_forward(crane_b200_model* m, const uint32_t* ids, size_t n, const float* pixel_values, const uint32_t* grid_thw,
                          size_t n_images, size_t start_pos, crane_b200_logits* out) {
    API_BEGIN(m)
    need_ready(m);
    if (!m->is_vl) fail(CRANE_B200_UNSUPPORTED, "vl_forward on a text-only model");
    if (!ids || n == 0) fail(CRANE_B200_INVALID_ARG, "vl_forward: empty input");
    check_room(m, start_pos, n);
    for (size_t i = 0; i < n; ++i)
        if (ids[i] >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", ids[i], m->V);
    std::vector<int> vis_rows;
    std::vector<uint32_t> pos3((size_t)3 * n);
    uint32_t next_pos = (uint32_t)start_pos;
    if (pixel_values && grid_thw && n_images > 0) {
        CUDA_OK(cudaEventRecord(m->pev0, m->stream));   // the prefill span of a VL request includes the ViT
        m->pev0_armed = true;
        m->encode_images(pixel_values, grid_thw, n_images);
        // build_position_ids (qwen3_5/vlm.rs:190-241)
        size_t i = 0, img = 0;
        while (i < n) {
            if (ids[i] != m->image_token_id) {
                pos3[0 * n + i] = pos3[1 * n + i] = pos3[2 * n + i] = next_pos++;
                ++i;
                continue;
            }
            if (img >= n_images) fail(CRANE_B200_INVALID_ARG, "image span %zu has no image_grid_thw entry", img);
            const uint32_t gt = grid_thw[3 * img], gh = grid_thw[3 * img + 1] / m->v_merge, gw = grid_thw[3 * img + 2] / m->v_merge;
            const size_t span = (size_t)gt * gh * gw;
            if (i + span > n) fail(CRANE_B200_INVALID_ARG, "image %zu needs %zu placeholder tokens but only %zu remain", img, span, n - i);
            const uint32_t base = next_pos, hw = gh * gw;
            for (size_t k = 0; k < span; ++k) {
                if (ids[i + k] != m->image_token_id) fail(CRANE_B200_INVALID_ARG, "image span %zu interrupted at token %zu", img, i + k);
                pos3[0 * n + i + k] = base + (uint32_t)k / hw;
                pos3[1 * n + i + k] = base + ((uint32_t)k % hw) / gw;
                pos3[2 * n + i + k] = base + ((uint32_t)k % hw) % gw;
                vis_rows.push_back((int)(i + k));
            }
            next_pos = base + std::max(gt, std::max(gh, gw));
            i += span;
            ++img;
        }
        if ((int)vis_rows.size() != m->img_tokens)
            fail(CRANE_B200_INVALID_ARG, "placeholder positions (%zu) != image embeddings (%d)", vis_rows.size(), m->img_tokens);
    } else {
        for (size_t i = 0; i < n; ++i) pos3[0 * n + i] = pos3[1 * n + i] = pos3[2 * n + i] = (uint32_t)(start_pos + i);
        next_pos = (uint32_t)(start_pos + n);
    }
    m->next_mrope_pos = next_pos;
    if (n == 1 && vis_rows.empty()) {
        m->arm_state(ids[0], start_pos, (int)pos3[0], (int)pos3[1], (int)pos3[2]);
        m->embed_step_input(1);
        m->decode_steps(1, 0);
        m->kv_len = start_pos + 1;
    } else {
        m->prefill(ids, nullptr, n, pos3.data(), start_pos, vis_rows.data(), (int)vis_rows.size(), 0);
    }
    if (pixel_values && m->pix_ev) CUDA_OK(cudaEventSynchronize(m->pix_ev));   // the caller may reuse its pixel buffer on return
    fill_logits(m, out);
    API_END(m)
}

int crane_b200_vl_decode_step(crane_b200_model* m, uint32_t token, size_t start_pos, crane_b200_logits* out) {
    API_BEGIN(m)
    need_ready(m);
    if (token >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", token, m->V);
    check_room(m, start_pos, 1);
    const int p = (int)m->next_mrope_pos;
    if (p > m->max_seq) fail(CRANE_B200_OOM, "rotary position exceeds max_seq_len");
    m->arm_state(token, start_pos, p, p, p);
    m->embed_step_input(1);
    m->decode_steps(1, 0);
    m->next_mrope_pos = (uint32_t)(p + 1);
    m->kv_len = start_pos + 1;
    fill_logits(m, out);
    API_END(m)
}

int crane_b200_vl_decode_step_argmax(crane_b200_model* m, uint32_t token, size_t start_pos, uint32_t* token_out) {
    if (!token_out) return CRANE_B200_INVALID_ARG;
    int r = crane_b200_vl_decode_step(m, token, start_pos, nullptr);
    if (r != CRANE_B200_OK) return r;
    API_BEGIN(m)
    CUDA_OK(cudaMemcpyAsync(m->h_tokens, m->out_tokens, sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    m->ll_err_fetch();
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->ll_err_verify();
    *token_out = m->h_tokens[0];
    API_END(m)
}

// ---- sequence slots + batched decode (ModelBackend::{setup,step}_batch_decode / extract_batch_kv, backend.rs:86-150) ----
int crane_b200_seq_create(crane_b200_model* m, int* seq_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!seq_out) fail(CRANE_B200_INVALID_ARG, "seq_create: null output");
    int s = -1;
    for (int i = 0; i < m->max_batch; ++i)
        if (!m->seq_used[i]) { s = i; break; }
    if (s < 0) fail(CRANE_B200_OOM, "all %d sequence slots are in use (engine.max_batch)", m->max_batch);
    m->seq_used[s] = 1;
    m->seq_kv[s] = 0; m->seq_pos[s] = 0;
    if (s == m->cur) { m->kv_len = 0; m->next_mrope_pos = 0; }
    if (m->hybrid && m->max_batch > 1) {
        const size_t cs = (size_t)m->conv_dim() * m->ck, rs = (size_t)m->nv * m->dk * m->dv;
        for (auto& l : m->layers)
            if (!l.full) {
                CUDA_OK(cudaMemsetAsync(l.conv_slots + s * cs, 0, cs * sizeof(float), m->stream));
                CUDA_OK(cudaMemsetAsync(l.rec_slots + s * rs, 0, rs * sizeof(float), m->stream));
            }
    }
    *seq_out = s;
    API_END(m)
}

// ---- KV swap: get_kv_caches / set_kv_caches (backend.rs:65-84) of the CURRENT sequence, one layer at a time ----
int crane_b200_kv_export(crane_b200_model* m, int layer, float* k_out, float* v_out, size_t capacity_floats, size_t* n_tokens) {
    API_BEGIN(m)
    need_ready(m);
    if (layer < 0 || layer >= m->L || !n_tokens) fail(CRANE_B200_INVALID_ARG, "kv_export: bad layer %d", layer);
    LayerW& l = m->layers[layer];
    const size_t T = m->kv_len;
    if (!l.full) {   // Gated-Delta-Net layer: k_out = conv window [conv_dim, ck], v_out = recurrent state [nv, dk, dv]; *n_tokens = 0
        const size_t cs = (size_t)m->conv_dim() * m->ck, rs = (size_t)m->nv * m->dk * m->dv;
        *n_tokens = 0;
        if (k_out || v_out) {
            if (capacity_floats < std::max(cs, rs)) fail(CRANE_B200_INVALID_ARG, "kv_export: GDN layer needs %zu floats", std::max(cs, rs));
            if (k_out) CUDA_OK(cudaMemcpyAsync(k_out, l.conv_state, cs * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
            if (v_out) CUDA_OK(cudaMemcpyAsync(v_out, l.rec_state, rs * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
            CUDA_OK(cudaStreamSynchronize(m->stream));
        }
    } else {
        *n_tokens = T;
        const size_t n = (size_t)m->nkv * T * m->D;
        if ((k_out || v_out) && T) {
            if (capacity_floats < n) fail(CRANE_B200_INVALID_ARG, "kv_export: layer holds %zu floats per tensor, capacity %zu", n, capacity_floats);
            float* tmp = nullptr;
            CUDA_OK(cudaMalloc((void**)&tmp, n * sizeof(float)));
            for (int which = 0; which < 2; ++which) {
                float* dst = which ? v_out : k_out;
                if (!dst) continue;
                if (m->kv_bits && which == 0) { const int rq = kv_dequant_pages_launch(m->stream, m->kvq_args(l), (int)T); if (rq) { cudaFree(tmp); fail(CRANE_B200_CUDA_ERROR, "kv_export dequant: %d", rq); } }
                const int r = m->kv_bits ? kv_pages_to_rows_launch(m->stream, which ? m->kq_sv : m->kq_sk, m->lo_kv, m->bt_identity, m->nkv, m->D, (int)T, tmp)
                                         : kv_pages_to_rows_launch(m->stream, which ? l.v_pool : l.k_pool, m->lo_kv, m->bt_cur(), m->nkv, m->D, (int)T, tmp);
                if (r) { cudaFree(tmp); fail(CRANE_B200_CUDA_ERROR, "kv_export kernel: %d", r); }
                cudaMemcpyAsync(dst, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream);
                cudaStreamSynchronize(m->stream);
            }
            cudaFree(tmp);
        }
    }
    API_END(m)
}

int crane_b200_kv_import(crane_b200_model* m, int layer, const float* k, const float* v, size_t n_tokens) {
    API_BEGIN(m)
    need_ready(m);
    if (layer < 0 || layer >= m->L || !k || !v) fail(CRANE_B200_INVALID_ARG, "kv_import: bad arguments (layer %d)", layer);
    LayerW& l = m->layers[layer];
    if (!l.full) {
        const size_t cs = (size_t)m->conv_dim() * m->ck, rs = (size_t)m->nv * m->dk * m->dv;
        CUDA_OK(cudaMemcpyAsync(l.conv_state, k, cs * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        CUDA_OK(cudaMemcpyAsync(l.rec_state, v, rs * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));
    } else if (n_tokens) {
        if (n_tokens > (size_t)m->max_seq) fail(CRANE_B200_OOM, "kv_import: %zu tokens exceed max_seq_len %d", n_tokens, m->max_seq);
        const size_t n = (size_t)m->nkv * n_tokens * m->D;
        float* tmp = nullptr;
        CUDA_OK(cudaMalloc((void**)&tmp, n * sizeof(float)));
        for (int which = 0; which < 2; ++which) {
            cudaMemcpyAsync(tmp, which ? v : k, n * sizeof(float), cudaMemcpyHostToDevice, m->stream);
            const int r = m->kv_bits ? kv_rows_to_pages_launch(m->stream, which ? m->kq_sv : m->kq_sk, m->lo_kv, m->bt_identity, m->nkv, m->D, (int)n_tokens, tmp)
                                     : kv_rows_to_pages_launch(m->stream, which ? l.v_pool : l.k_pool, m->lo_kv, m->bt_cur(), m->nkv, m->D, (int)n_tokens, tmp);
            cudaStreamSynchronize(m->stream);
            if (r) { cudaFree(tmp); fail(CRANE_B200_CUDA_ERROR, "kv_import kernel: %d", r); }
        }
        cudaFree(tmp);
        if (m->kv_bits) {   // the imported rows are re-quantised: codes of code * scale values reproduce themselves
            LAUNCH_OK(kv_quant_rows_launch(m->stream, m->kvq_args(l), 0, (int)n_tokens));
            CUDA_OK(cudaStreamSynchronize(m->stream));
        }
    }
    API_END(m)
}

// After importing every layer: the cached length and next rotary position of the current sequence (set_kv_caches leaves them to
// the caller in the reference too -- `forward_step(.., start_pos)` carries them).
int crane_b200_kv_set_len(crane_b200_model* m, size_t n_tokens, uint32_t next_rotary_pos) {
    API_BEGIN(m)
    need_ready(m);
    if (n_tokens > (size_t)m->max_seq) fail(CRANE_B200_OOM, "kv_set_len: %zu exceeds max_seq_len %d", n_tokens, m->max_seq);
    m->kv_len = n_tokens;
    m->next_mrope_pos = next_rotary_pos;
    API_END(m)
}

int crane_b200_seq_fork(crane_b200_model* m, int src, int* seq_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!seq_out) fail(CRANE_B200_INVALID_ARG, "seq_fork: null output");
    if (src < 0 || src >= m->max_batch || !m->seq_used[src]) fail(CRANE_B200_INVALID_ARG, "seq_fork: bad sequence %d", src);
    int s = -1;
    for (int i = 0; i < m->max_batch; ++i)
        if (!m->seq_used[i]) { s = i; break; }
    if (s < 0) fail(CRANE_B200_OOM, "all %d sequence slots are in use (engine.max_batch)", m->max_batch);
    m->seq_used[s] = 1;
    m->fork_seq(src, s);
    *seq_out = s;
    API_END(m)
}

int crane_b200_seq_free(crane_b200_model* m, int seq) {
    API_BEGIN(m)
    need_ready(m);
    if (seq < 0 || seq >= m->max_batch || !m->seq_used[seq]) fail(CRANE_B200_INVALID_ARG, "seq_free: bad sequence %d", seq);
    if (seq == 0) fail(CRANE_B200_INVALID_ARG, "sequence 0 is the handle's implicit sequence (use clear_kv_cache)");
    m->select_seq(0);
    m->seq_used[seq] = 0;
    m->seq_kv[seq] = 0; m->seq_pos[seq] = 0;
    API_END(m)
}

int crane_b200_seq_select(crane_b200_model* m, int seq) {
    API_BEGIN(m)
    need_ready(m);
    if (seq < 0 || seq >= m->max_batch || !m->seq_used[seq]) fail(CRANE_B200_INVALID_ARG, "seq_select: bad sequence %d", seq);
    m->select_seq(seq);
    API_END(m)
}

int crane_b200_decode_batch(crane_b200_model* m, const int* seqs, const uint32_t* tokens, size_t n, size_t n_steps, uint32_t* tokens_out,
                            float* logits_host) {
    API_BEGIN(m)
    need_ready(m);
    if (!seqs || !tokens || !tokens_out || n == 0 || n_steps == 0) fail(CRANE_B200_INVALID_ARG, "decode_batch: bad arguments");
    if (m->hybrid) fail(CRANE_B200_UNSUPPORTED, "decode_batch on the hybrid model");
    if (n_steps > (size_t)m->out_cap) fail(CRANE_B200_INVALID_ARG, "decode_batch: n_steps %zu > %d", n_steps, m->out_cap);
    m->seq_kv[m->cur] = m->kv_len; m->seq_pos[m->cur] = m->next_mrope_pos;
    for (size_t i = 0; i < n; ++i) {
        const int s = seqs[i];
        if (s < 0 || s >= m->max_batch || !m->seq_used[s]) fail(CRANE_B200_INVALID_ARG, "decode_batch: bad sequence %d", s);
        for (size_t j = 0; j < i; ++j) if (seqs[j] == s) fail(CRANE_B200_INVALID_ARG, "decode_batch: sequence %d listed twice", s);
        if (tokens[i] >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", tokens[i], m->V);
        if (m->seq_kv[s] + n_steps > (size_t)m->max_seq) fail(CRANE_B200_OOM, "sequence %d would exceed max_seq_len %d", s, m->max_seq);
    }
    CUDA_OK(cudaEventRecord(m->dev0, m->stream));
    // groups of 4 / 2 / 1 sequences share one pass over the weights (cb::gemv_kernel<B>)
    int maxg = 4;
    if (!m->any_quant)
        for (int K : {m->H, m->I, m->q_dim()}) maxg = std::min(maxg, gemv_max_group(K, m->V, m->num_sms));
    size_t done = 0;
    while (done < n) {
        const int B = std::min(maxg, (n - done >= 4) ? 4 : (n - done >= 2) ? 2 : 1);
        m->stage_state();
        for (int b = 0; b < B; ++b) {
            const int s = seqs[done + b];
            SeqState& st = m->h_state[b];
            st.kv_len = (int)m->seq_kv[s];
            st.pos[0] = st.pos[1] = st.pos[2] = (int)m->seq_pos[s];
            st.token = tokens[done + b];
            st.step = 0;
            st.slot = s;
        }
        m->push_state(B);
        m->embed_step_input(B);
        for (size_t t = 0; t < n_steps; ++t) m->decode_step_graphed(1, B);
        for (int b = 0; b < B; ++b)
            CUDA_OK(cudaMemcpyAsync(tokens_out + (done + b) * n_steps, m->out_tokens + (size_t)b * m->out_cap, n_steps * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, m->stream));
        if (logits_host)
            CUDA_OK(cudaMemcpyAsync(logits_host + done * (size_t)m->V, m->logits, (size_t)B * m->V * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
        CUDA_OK(cudaStreamSynchronize(m->stream));     // tokens_out / logits_host are the caller's memory: complete before returning
        for (int b = 0; b < B; ++b) {
            const int s = seqs[done + b];
            m->seq_kv[s] += n_steps;
            m->seq_pos[s] += (uint32_t)n_steps;
        }
        done += B;
    }
    CUDA_OK(cudaEventRecord(m->dev1, m->stream));
    m->last_decode_steps = n_steps;
    m->kv_len = m->seq_kv[m->cur]; m->next_mrope_pos = m->seq_pos[m->cur];
    API_END(m)
}

// ---- device-side sampling (crane-serve/src/engine/sampling.rs:169-480) ----
int crane_b200_sample(crane_b200_model* m, const crane_b200_sampling* p, uint32_t* token_out) {
    API_BEGIN(m)
    need_ready(m);
    sampler_run(m->stream, m->sampler, m->logits, m->V, 1, p, token_out);
    m->ll_err_fetch();
    CUDA_OK(cudaStreamSynchronize(m->stream));
    m->ll_err_verify();
    API_END(m)
}

int crane_b200_forward_step_sample(crane_b200_model* m, const uint32_t* ids, size_t n, size_t start_pos, const crane_b200_sampling* p,
                                   uint32_t* token_out) {
    if (!p || !token_out) return CRANE_B200_INVALID_ARG;
    const int r = crane_b200_forward_step(m, ids, n, start_pos, nullptr);
    if (r != CRANE_B200_OK) return r;
    return crane_b200_sample(m, p, token_out);
}

int crane_b200_topk(crane_b200_model* m, size_t k, uint32_t* idx_out, float* vals_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!idx_out || k == 0 || k > (size_t)TOPK_MAX || k > (size_t)m->V) fail(CRANE_B200_INVALID_ARG, "topk: k must be in 1..min(%d, vocab)", TOPK_MAX);
    m->sampler.reserve(1, 0, 0);
    LAUNCH_OK(sampler_topk_launch(m->stream, m->logits, m->V, 1, (int)k, m->sampler.tk_idx, m->sampler.tk_val));
    CUDA_OK(cudaMemcpyAsync(idx_out, m->sampler.tk_idx, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
    if (vals_out) CUDA_OK(cudaMemcpyAsync(vals_out, m->sampler.tk_val, k * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(cudaStreamSynchronize(m->stream));
    API_END(m)
}

int crane_b200_decode_batch_sample(crane_b200_model* m, const int* seqs, const uint32_t* tokens, size_t n, const crane_b200_sampling* params,
                                   uint32_t* tokens_out) {
    API_BEGIN(m)
    need_ready(m);
    if (!seqs || !tokens || !tokens_out || !params || n == 0) fail(CRANE_B200_INVALID_ARG, "decode_batch_sample: bad arguments");
    if (m->hybrid) fail(CRANE_B200_UNSUPPORTED, "decode_batch_sample on the hybrid model");
    m->seq_kv[m->cur] = m->kv_len; m->seq_pos[m->cur] = m->next_mrope_pos;
    for (size_t i = 0; i < n; ++i) {
        const int s = seqs[i];
        if (s < 0 || s >= m->max_batch || !m->seq_used[s]) fail(CRANE_B200_INVALID_ARG, "decode_batch_sample: bad sequence %d", s);
        for (size_t j = 0; j < i; ++j) if (seqs[j] == s) fail(CRANE_B200_INVALID_ARG, "decode_batch_sample: sequence %d listed twice", s);
        if (tokens[i] >= (uint32_t)m->V) fail(CRANE_B200_INVALID_ARG, "token id %u >= vocab %d", tokens[i], m->V);
        if (m->seq_kv[s] + 1 > (size_t)m->max_seq) fail(CRANE_B200_OOM, "sequence %d would exceed max_seq_len %d", s, m->max_seq);
    }
    int maxg = 4;
    if (!m->any_quant)
        for (int K : {m->H, m->I, m->q_dim()}) maxg = std::min(maxg, gemv_max_group(K, m->V, m->num_sms));
    size_t done = 0;
    while (done < n) {
        const int B = std::min(maxg, (n - done >= 4) ? 4 : (n - done >= 2) ? 2 : 1);
        m->stage_state();
        for (int b = 0; b < B; ++b) {
            const int s = seqs[done + b];
            SeqState& st = m->h_state[b];
            st.kv_len = (int)m->seq_kv[s];
            st.pos[0] = st.pos[1] = st.pos[2] = (int)m->seq_pos[s];
            st.token = tokens[done + b];
            st.step = 0;
            st.slot = s;
        }
        m->push_state(B);
        m->embed_step_input(B);
        m->decode_step_graphed(0, B);                // logits [B, V] stay on the device; the host picks the next inputs from the sampled ids
        sampler_run(m->stream, m->sampler, m->logits, m->V, (size_t)B, params + done, tokens_out + done);
        for (int b = 0; b < B; ++b) {
            const int s = seqs[done + b];
            m->seq_kv[s] += 1;
            m->seq_pos[s] += 1;
        }
        done += B;
    }
    m->kv_len = m->seq_kv[m->cur]; m->next_mrope_pos = m->seq_pos[m->cur];
    API_END(m)
}

int crane_b200_op_qlinear(int device, const float* x, size_t m, size_t k, const void* raw, size_t raw_bytes, int ggml_type, size_t n,
                          const float* norm_w, float eps, float* y) {
    if (!x || !raw || !y || m == 0 || n == 0 || k == 0 || (k % 256) != 0) return CRANE_B200_INVALID_ARG;
    if (ggml_type != QT_Q4_K && ggml_type != QT_Q6_K && ggml_type != QT_Q8_0) return CRANE_B200_UNSUPPORTED;
    if (raw_bytes != n * (k / q_src_block_elems(ggml_type)) * q_src_block_bytes(ggml_type)) return CRANE_B200_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CRANE_B200_CUDA_ERROR;
    int rc = CRANE_B200_OK;
    float *dx = nullptr, *dy = nullptr, *dn = nullptr;
    unsigned char *dw = nullptr, *dxq = nullptr;
    try {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        const size_t rb = (k / 256) * (size_t)q_sb_bytes(ggml_type);
        std::vector<unsigned char> packed(n * rb);
        q_repack_rows(ggml_type, (const unsigned char*)raw, packed.data(), n, (int)k);
        CUDA_OK(cudaMalloc((void**)&dx, m * k * 4)); CUDA_OK(cudaMalloc((void**)&dy, m * n * 4)); CUDA_OK(cudaMalloc((void**)&dw, packed.size()));
        CUDA_OK(cudaMalloc((void**)&dxq, xquant_bytes(4, (int)k)));
        CUDA_OK(cudaMemcpy(dx, x, m * k * 4, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dw, packed.data(), packed.size(), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemset(dy, 0, m * n * 4));
        if (norm_w) { CUDA_OK(cudaMalloc((void**)&dn, k * 4)); CUDA_OK(cudaMemcpy(dn, norm_w, k * 4, cudaMemcpyHostToDevice)); }
        for (size_t s0 = 0; s0 < m;) {
            const int B = (m - s0 >= 4) ? 4 : (m - s0 >= 2) ? 2 : 1;
            LAUNCH_OK(xquant_launch(nullptr, B, dx + s0 * k, (int)k, (int)k, dn, eps, xq_mode_for(ggml_type), dxq, false));
            QGemvArgs qa = {};
            qa.g.W = reinterpret_cast<const bf16*>(dw); qa.g.N = (int)n; qa.g.K = (int)k; qa.g.y = dy + s0 * n; qa.g.ldy = (int)n; qa.g.eps = eps;
            qa.qtype = ggml_type; qa.epi = GEMV_STORE; qa.norm = dn ? 1 : 0; qa.xq = dxq;
            LAUNCH_OK(qgemv_launch(nullptr, B, qa, prop.multiProcessorCount, false));
            s0 += B;
        }
        CUDA_OK(cudaDeviceSynchronize());
        CUDA_OK(cudaMemcpy(y, dy, m * n * 4, cudaMemcpyDeviceToHost));
    } catch (const EngineError& e) {
        g_create_error = e.msg;
        rc = e.code;
    }
    for (void* p : {(void*)dx, (void*)dy, (void*)dn, (void*)dw, (void*)dxq}) if (p) cudaFree(p);
    return rc;
}

/* debugging aid (not in the public header): copy a prefill workspace to the host -- 0: residual stream x [S, H], 1: qkv [S, qkv_dim],
 * 2: the f32 rows in front of / behind a quantised linear [S, max(I, q_dim)] */
extern "C" CRANE_B200_API int crane_b200_debug_peek(crane_b200_model* m, int which, float* out, size_t n_floats) {
    API_BEGIN(m)
    need_ready(m);
    const float* src = which == 0 ? m->x : which == 1 ? m->qkv : m->rows_f32;
    if (!src || !out) fail(CRANE_B200_INVALID_ARG, "debug_peek: buffer %d does not exist", which);
    CUDA_OK(cudaStreamSynchronize(m->stream));
    CUDA_OK(cudaMemcpy(out, src, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    API_END(m)
}

static int op_sampler(int device, const float* logits, size_t vocab, size_t k, const crane_b200_sampling* p, uint32_t* out, float* logits_after) {
    if (!logits || !out || vocab == 0 || vocab > (size_t)INT32_MAX) return CRANE_B200_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CRANE_B200_CUDA_ERROR;
    SamplerScratch sc;
    float* dl = nullptr;
    int rc = CRANE_B200_OK;
    try {
        CUDA_OK(cudaMalloc((void**)&dl, vocab * sizeof(float)));
        CUDA_OK(cudaMemcpy(dl, logits, vocab * sizeof(float), cudaMemcpyHostToDevice));
        if (p) {
            sampler_run(nullptr, sc, dl, (int)vocab, 1, p, out);
            if (logits_after) CUDA_OK(cudaMemcpy(logits_after, dl, vocab * sizeof(float), cudaMemcpyDeviceToHost));
        } else {
            if (k == 0 || k > (size_t)TOPK_MAX || k > vocab) fail(CRANE_B200_INVALID_ARG, "op_topk: k must be in 1..min(%d, vocab)", TOPK_MAX);
            sc.reserve(1, 0, 0);
            LAUNCH_OK(sampler_topk_launch(nullptr, dl, (int)vocab, 1, (int)k, sc.tk_idx, nullptr));
            CUDA_OK(cudaMemcpy(out, sc.tk_idx, k * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        }
    } catch (const EngineError& e) {
        g_create_error = e.msg;
        rc = e.code;
    }
    if (dl) cudaFree(dl);
    sc.release();
    return rc;
}
int crane_b200_op_topk(int device, const float* logits, size_t vocab, size_t k, uint32_t* idx_out) {
    return op_sampler(device, logits, vocab, k, nullptr, idx_out, nullptr);
}
int crane_b200_op_sample(int device, const float* logits, size_t vocab, const crane_b200_sampling* p, uint32_t* token_out, float* logits_after) {
    if (!p) return CRANE_B200_INVALID_ARG;
    return op_sampler(device, logits, vocab, 0, p, token_out, logits_after);
}

int crane_b200_last_timing(const crane_b200_model* m, float* prefill_ms, float* decode_ms, size_t* decode_steps) {
    if (!m || !m->finalized) return CRANE_B200_INVALID_ARG;
    cudaSetDevice(m->device);
    float ms = 0.f;
    crane_b200_model* mm = const_cast<crane_b200_model*>(m);
    if (cudaEventSynchronize(m->pev1) == cudaSuccess && cudaEventElapsedTime(&ms, m->pev0, m->pev1) == cudaSuccess) mm->last_prefill_ms = ms;
    if (cudaEventSynchronize(m->dev1) == cudaSuccess && cudaEventElapsedTime(&ms, m->dev0, m->dev1) == cudaSuccess) mm->last_decode_ms = ms;
    cudaGetLastError();
    if (prefill_ms) *prefill_ms = m->last_prefill_ms;
    if (decode_ms) *decode_ms = m->last_decode_ms;
    if (decode_steps) *decode_steps = m->last_decode_steps;
    return CRANE_B200_OK;
}

int crane_b200_op_gemm(int device, const uint16_t* a, const uint16_t* a_lo, const uint16_t* w, int M, int N, int K, int mode, const float* bias,
                       void* out_inout, int use_simt) {
    if (!a || !w || !out_inout) return CRANE_B200_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CRANE_B200_CUDA_ERROR;
    const bool half_out = (mode == EPI_STORE_BF16 || mode == EPI_SILU_MUL_BF16 || mode == EPI_GELU_ERF_BF16 || mode == EPI_GELU_TANH_BF16);
    const int out_cols = (mode == EPI_SILU_MUL_BF16) ? N / 2 : N;
    const size_t out_bytes = (size_t)M * out_cols * (half_out ? 2 : 4);
    bf16 *da = nullptr, *dw = nullptr;
    float* db = nullptr;
    void* dout = nullptr;
    int rc = CRANE_B200_CUDA_ERROR;
    if (cudaMalloc(&da, (size_t)M * K * 2 * 2) == cudaSuccess && cudaMalloc(&dw, (size_t)N * K * 2) == cudaSuccess &&
        cudaMalloc(&dout, out_bytes) == cudaSuccess && (!bias || cudaMalloc(&db, (size_t)N * 4) == cudaSuccess)) {
        cudaMemcpy(da, a, (size_t)M * K * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dw, w, (size_t)N * K * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dout, out_inout, out_bytes, cudaMemcpyHostToDevice);
        if (bias) cudaMemcpy(db, bias, (size_t)N * 4, cudaMemcpyHostToDevice);
        if (a_lo) cudaMemcpy(da + (size_t)M * K, a_lo, (size_t)M * K * 2, cudaMemcpyHostToDevice);
        GemmEpi ep{dout, nullptr, out_cols, db, mode};
        const int r = gemm_bf16_launch(nullptr, da, a_lo ? da + (size_t)M * K : nullptr, K, dw, M, N, K, ep, use_simt != 0);
        if (r == -1000) rc = CRANE_B200_UNSUPPORTED;
        else if (r == 0 && cudaDeviceSynchronize() == cudaSuccess) {
            cudaMemcpy(out_inout, dout, out_bytes, cudaMemcpyDeviceToHost);
            rc = CRANE_B200_OK;
            if (getenv("CRANE_B200_GEMM_PROF") && !use_simt) {      // tile timeline + event timing of this shape (tools/gemm_probe.py)
                const size_t max_tiles = 8192;
                unsigned long long* dprof = nullptr;
                cudaMalloc(&dprof, max_tiles * 8 * sizeof(unsigned long long));
                GemmEpi ep2 = ep;
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                for (int i = 0; i < 5; ++i) gemm_bf16_launch(nullptr, da, a_lo ? da + (size_t)M * K : nullptr, K, dw, M, N, K, ep2, false);
                cudaEventRecord(e0, nullptr);
                for (int i = 0; i < 20; ++i) gemm_bf16_launch(nullptr, da, a_lo ? da + (size_t)M * K : nullptr, K, dw, M, N, K, ep2, false);
                cudaEventRecord(e1, nullptr);
                cudaDeviceSynchronize();
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaMemset(dprof, 0, max_tiles * 8 * sizeof(unsigned long long));
                ep2.prof = dprof;
                gemm_bf16_launch(nullptr, da, a_lo ? da + (size_t)M * K : nullptr, K, dw, M, N, K, ep2, false);
                cudaDeviceSynchronize();
                std::vector<unsigned long long> hp(max_tiles * 8);
                cudaMemcpy(hp.data(), dprof, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                unsigned long long t0 = ~0ull, t1 = 0, t7 = 0;
                size_t n = 0;
                for (size_t i = 0; i < max_tiles; ++i) if (hp[8 * i]) { t0 = std::min(t0, hp[8 * i]); t1 = std::max(t1, hp[8 * i + 6]); t7 = std::max(t7, hp[8 * i + 7]); ++n; }
                // (stamp 6 is taken by thread 0 right after the CTA barrier: BAR.SYNC defers its blocking to the next dependent instruction,
                //  so that stamp is the ARRIVAL of the producer warp; the TMEM dealloc stamp 7 is the real end of the CTA)
                fprintf(stderr, "   isolated launch: first CTA start -> last CTA end %.2f us\n", ((double)t7 - (double)t0) / 1e3);
                double acc[6] = {0, 0, 0, 0, 0, 0}, start_max = 0;
                for (size_t i = 0; i < max_tiles; ++i) if (hp[8 * i]) {
                    for (int j = 0; j < 6; ++j) acc[j] += (double)((long long)hp[8 * i + j + 1] - (long long)hp[8 * i + j]);
                    start_max = std::max(start_max, (double)(hp[8 * i] - t0));
                }
                fprintf(stderr, "gemm_prof M=%d N=%d K=%d split=%d: %.2f us/launch back-to-back (%.1f TFLOP/s); one launch: %zu CTAs, span %.2f us, last CTA start +%.2f us; "
                        "per-CTA mean ns: setup %.0f | first operands %.0f | mainloop issue %.0f | accumulator wait %.0f | epilogue %.0f\n",
                        M, N, K, a_lo ? 1 : 0, ms * 1e3 / 20, 2.0 * M * N * K / (ms * 1e-3 / 20) / 1e12, n, (t1 - t0) / 1e3, start_max / 1e3,
                        acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n);
                {   // the same launch 8 times back to back, each with its own stamp block: where the time BETWEEN kernels goes
                    const int NB = 8;
                    cudaMemset(dprof, 0, max_tiles * 8 * sizeof(unsigned long long));
                    const size_t per = max_tiles / NB;
                    for (int i = 0; i < NB; ++i) {
                        ep2.prof = dprof + (size_t)i * per * 8;
                        gemm_bf16_launch(nullptr, da, a_lo ? da + (size_t)M * K : nullptr, K, dw, M, N, K, ep2, false);
                    }
                    cudaDeviceSynchronize();
                    cudaMemcpy(hp.data(), dprof, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                    unsigned long long prev_end = 0;
                    for (int i = 0; i < NB; ++i) {
                        unsigned long long s0 = ~0ull, s0max = 0, e6 = 0, e6min = ~0ull, e7 = 0;
                        for (size_t c = 0; c < per; ++c) {
                            const unsigned long long* h = &hp[((size_t)i * per + c) * 8];
                            if (!h[0]) continue;
                            s0 = std::min(s0, h[0]); s0max = std::max(s0max, h[0]); e6 = std::max(e6, h[6]); e6min = std::min(e6min, h[6]); e7 = std::max(e7, h[7]);
                        }
                        if (i >= 4)
                            fprintf(stderr, "   b2b launch %d: first CTA starts %+.2f us after the previous launch's last CTA ended; CTA starts spread %.2f us; "
                                    "first start -> last end %.2f us\n", i, ((double)s0 - (double)prev_end) / 1e3, (s0max - s0) / 1e3, ((double)e7 - (double)s0) / 1e3);
                        prev_end = e7;
                    }
                }
                cudaFree(dprof);
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
        } else {
            g_create_error = std::string("op_gemm: ") + cudaGetErrorString(cudaGetLastError()) + " rc=" + std::to_string(r);
        }
    }
    cudaFree(da); cudaFree(dw); cudaFree(db); cudaFree(dout);
    return rc;
}

}  // extern "C"

#include "engine_tts.inc"
#include "engine_comm.inc"
#include "engine_loaders.inc"
