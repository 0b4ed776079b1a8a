// C++ host mirror test (built by tests/test_cpp_mirror.py with g++ against include/crane_b200.hpp and libcrane_b200.so).
//   ./test_backend symbols   no GPU needed: the library loads, create() on a machine without a usable GPU fails with a status + message
//   ./test_backend gpu       tiny random-weight Qwen3 through the ModelBackend surface: start_pos contract, chunked == single prefill,
//                            forward_step argmax == forward_step_argmax == generate, batched decode == per-sequence decode
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "crane_b200.hpp"

using namespace crane_b200;

static const char* CFG =
    "{\"model_type\": \"qwen3\", \"hidden_size\": 256, \"intermediate_size\": 512, \"num_hidden_layers\": 2, \"num_attention_heads\": 4,"
    " \"num_key_value_heads\": 2, \"head_dim\": 128, \"vocab_size\": 512, \"rms_norm_eps\": 1e-6, \"rope_theta\": 1000000.0,"
    " \"tie_word_embeddings\": true, \"max_position_embeddings\": 512, \"engine\": {\"max_seq_len\": 256, \"max_batch\": 4}}";

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<float> randn(std::mt19937& g, size_t n, float scale, float mean = 0.f) {
    std::normal_distribution<float> d(0.f, 1.f);
    std::vector<float> v(n);
    for (auto& x : v) x = mean + scale * d(g);
    return v;
}

static int run_gpu() {
    B200Backend m(CFG, 0);
    std::mt19937 g(7);
    const int H = 256, I = 512, V = 512, D = 128, nh = 4, nkv = 2, L = 2;
    auto put = [&](const std::string& name, std::vector<int64_t> shape, float scale, float mean = 0.f) {
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        auto w = randn(g, n, scale, mean);
        m.load_tensor(name, CRANE_B200_F32, shape, w.data());
    };
    put("model.embed_tokens.weight", {V, H}, 1.0f);
    put("model.norm.weight", {H}, 0.02f, 1.0f);
    for (int l = 0; l < L; ++l) {
        const std::string p = "model.layers." + std::to_string(l) + ".";
        put(p + "input_layernorm.weight", {H}, 0.02f, 1.0f);
        put(p + "post_attention_layernorm.weight", {H}, 0.02f, 1.0f);
        put(p + "self_attn.q_proj.weight", {nh * D, H}, 1.f / std::sqrt((float)H));
        put(p + "self_attn.k_proj.weight", {nkv * D, H}, 1.f / std::sqrt((float)H));
        put(p + "self_attn.v_proj.weight", {nkv * D, H}, 1.f / std::sqrt((float)H));
        put(p + "self_attn.o_proj.weight", {H, nh * D}, 1.f / std::sqrt((float)(nh * D)));
        put(p + "self_attn.q_norm.weight", {D}, 0.02f, 1.0f);
        put(p + "self_attn.k_norm.weight", {D}, 0.02f, 1.0f);
        put(p + "mlp.gate_proj.weight", {I, H}, 1.f / std::sqrt((float)H));
        put(p + "mlp.up_proj.weight", {I, H}, 1.f / std::sqrt((float)H));
        put(p + "mlp.down_proj.weight", {H, I}, 1.f / std::sqrt((float)I));
    }
    m.finalize();
    CHECK(m.num_layers() == 2 && m.vocab_size() == V && m.supports_batch_decode());

    std::vector<uint32_t> prompt(37);
    for (size_t i = 0; i < prompt.size(); ++i) prompt[i] = (uint32_t)((i * 37 + 11) % V);
    // single-pass prefill vs chunked prefill (tests/qwen3_5_chunked_prefill.rs in the reference): identical last-position logits
    auto full = m.forward_step(prompt, 0);
    CHECK(full.size() == (size_t)V && m.kv_len() == prompt.size());
    m.clear_kv_cache();
    m.forward_step(std::vector<uint32_t>(prompt.begin(), prompt.begin() + 20), 0);
    auto part = m.forward_step(std::vector<uint32_t>(prompt.begin() + 20, prompt.end()), 20);
    float mx = 0.f, dmax = 0.f;
    for (int i = 0; i < V; ++i) { mx = std::fmax(mx, std::fabs(full[i])); dmax = std::fmax(dmax, std::fabs(full[i] - part[i])); }
    std::printf("chunked vs single prefill: rel %.3e\n", dmax / mx);
    CHECK(dmax / mx < 1e-4f);
    // start_pos must equal the cached length; the handle stays usable after the error
    bool threw = false;
    try { m.forward_step({1, 2, 3}, 5); } catch (const Error& e) { threw = e.code() == CRANE_B200_INVALID_ARG; }
    CHECK(threw);
    // argmax of the logits (lowest index among maxima) == forward_step_argmax == first generated token
    int am = 0;
    for (int i = 1; i < V; ++i) if (full[i] > full[am]) am = i;
    m.clear_kv_cache();
    CHECK(m.forward_step_argmax(prompt, 0) == (uint32_t)am);
    auto gen = m.generate(prompt, 6);
    CHECK(gen.size() == 6 && gen[0] == (uint32_t)am);
    // batched decode == each sequence alone
    std::vector<std::vector<uint32_t>> prompts = {prompt, std::vector<uint32_t>(prompt.begin(), prompt.begin() + 9), {5, 6, 7}};
    std::vector<std::vector<uint32_t>> solo;
    for (auto& p : prompts) solo.push_back(m.generate(p, 5));
    m.clear_kv_cache();
    std::vector<int> seqs;
    std::vector<uint32_t> first;
    for (size_t i = 0; i < prompts.size(); ++i) {
        const int s = i == 0 ? 0 : m.seq_create();
        m.seq_select(s);
        if (i == 0) m.clear_kv_cache();
        first.push_back(m.forward_step_argmax(prompts[i], 0));
        seqs.push_back(s);
    }
    auto toks = m.decode_batch(seqs, first, 4);
    for (size_t i = 0; i < prompts.size(); ++i) {
        CHECK(first[i] == solo[i][0]);
        for (int j = 0; j < 4; ++j) CHECK(toks[i][j] == solo[i][j + 1]);
    }
    // sampler: temperature 0 == argmax; top-k order starts at the argmax; a fork continues like its source
    m.seq_select(0);
    m.clear_kv_cache();
    auto lg = m.forward_step(prompt, 0);
    CHECK(m.sample(B200Backend::greedy()) == (uint32_t)am);
    auto tk = m.topk(8);
    CHECK(tk[0] == (uint32_t)am);
    for (int j = 0; j + 1 < 8; ++j) CHECK(lg[tk[j]] > lg[tk[j + 1]] || (lg[tk[j]] == lg[tk[j + 1]] && tk[j] < tk[j + 1]));
    crane_b200_sampling sp = B200Backend::greedy();
    sp.temperature = 0.8f; sp.top_k = 8; sp.top_p = 0.9f; sp.seed = 42;
    const uint32_t drawn = m.sample(sp);
    bool in_topk = false;
    for (uint32_t t : tk) in_topk = in_topk || t == drawn;
    CHECK(in_topk);
    const int f = m.seq_fork(0);
    const uint32_t a1 = m.forward_step_argmax({(uint32_t)am}, prompt.size());
    m.seq_select(f);
    CHECK(m.kv_len() == prompt.size());
    CHECK(m.forward_step_argmax({(uint32_t)am}, prompt.size()) == a1);
    // KV swap: export layer 0 of the fork, import it back, nothing changes
    const size_t per = (size_t)2 /*n_kv*/ * m.kv_len() * 128;
    auto c0 = m.kv_export(0, per);
    CHECK(c0.n_tokens == m.kv_len());
    m.kv_import(0, c0);
    m.kv_set_len(m.kv_len(), (uint32_t)m.kv_len());
    auto again = m.kv_export(0, per);
    CHECK(again.k == c0.k && again.v == c0.v);
    std::printf("C++ mirror: ok (%llu kernel launches, %llu KV bytes)\n", (unsigned long long)m.kernel_launches(), (unsigned long long)m.active_kv_cache_bytes());
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "gpu") == 0) {
        try { return run_gpu(); } catch (const Error& e) { std::printf("FAILED: %s\n", e.what()); return 1; }
    }
    // symbols mode: linking proved the exports; without a usable GPU create() must fail loudly, not fall back to anything
    try {
        B200Backend m(CFG, 0);
        std::printf("create() succeeded: a GPU is present\n");
    } catch (const Error& e) {
        std::printf("create() failed as it must without a B200: %s\n", e.what());
        if (e.code() == CRANE_B200_OK) return 1;
    }
    std::printf("C++ mirror: symbols ok\n");
    return 0;
}
