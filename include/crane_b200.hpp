// C++ host-side mirror of the reference's backend interface over the C ABI (include/crane_b200.h).
//
// The reference's host language is Rust; this image has no Rust toolchain, so the compiled host mirror is C++ (header-only,
// nothing but the C ABI underneath).  Names, argument meaning and error behaviour follow
//   trait ModelBackend ................ crane-serve/src/engine/backend.rs:30-151
//   Qwen3Backend / Qwen3_5Backend ..... crane-serve/src/engine/backend.rs:505-748
//   Model::generate ................... crane-core/src/models/qwen3/model.rs:275-349
// so that a port of the reference's own backend tests reads the same: `forward_step(ids, start_pos)` returns the logits of the
// last position, `start_pos` must equal the cached length, errors surface as exceptions carrying the C status code and the
// handle's message (anyhow::Result in the reference), and a failed call leaves the backend usable.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "crane_b200.h"

namespace crane_b200 {

class Error : public std::runtime_error {
   public:
    Error(int code, const std::string& msg) : std::runtime_error("crane_b200 error " + std::to_string(code) + ": " + msg), code_(code) {}
    int code() const { return code_; }

   private:
    int code_;
};

// `trait ModelBackend: Send + 'static` (backend.rs:30-151).  Tokenizer / EOS lookup stay with the caller, as in INTEGRATION.md.
class ModelBackend {
   public:
    virtual ~ModelBackend() = default;
    virtual std::vector<float> forward_step(const std::vector<uint32_t>& input_ids, size_t start_pos) = 0;   // backend.rs:42
    virtual void clear_kv_cache() = 0;                                                                     // :45
    virtual size_t num_layers() const = 0;                                                                 // :48
    virtual void warmup() = 0;                                                                             // :61
    virtual bool supports_kv_swap() const { return false; }                                                // :65
    virtual uint64_t active_kv_cache_bytes() const { return 0; }                                           // :82
    virtual bool supports_batch_decode() const { return false; }                                           // :86
};

// The one backend this library provides: dense Qwen3 / Qwen3.5 hybrid / GGUF-quantised Qwen3 behind the same handle.
class B200Backend : public ModelBackend {
   public:
    // `Qwen3Backend::new(model_path, device, dtype)` (backend.rs:615-625) minus the file I/O: the caller streams the checkpoint in
    // with load_tensor / load_tensor_ggml (names exactly as in the safetensors / GGUF file), then finalize().
    B200Backend(const std::string& config_json, int device_ordinal) {
        const int rc = crane_b200_create(config_json.c_str(), device_ordinal, &h_);
        if (rc != CRANE_B200_OK) throw Error(rc, crane_b200_last_error(nullptr));
    }
    ~B200Backend() override { crane_b200_destroy(h_); }
    B200Backend(const B200Backend&) = delete;
    B200Backend& operator=(const B200Backend&) = delete;

    void load_tensor(const std::string& name, crane_b200_dtype dt, const std::vector<int64_t>& shape, const void* data) {
        ck(crane_b200_load_tensor(h_, name.c_str(), (int)dt, shape.data(), (int)shape.size(), data));
    }
    void load_tensor_ggml(const std::string& name, int ggml_type, const std::vector<int64_t>& shape, const void* blocks, size_t nbytes) {
        ck(crane_b200_load_tensor_ggml(h_, name.c_str(), ggml_type, shape.data(), (int)shape.size(), blocks, nbytes));
    }
    void finalize() { ck(crane_b200_finalize(h_)); }

    // ---- ModelBackend ----
    std::vector<float> forward_step(const std::vector<uint32_t>& input_ids, size_t start_pos) override {
        crane_b200_logits lg{};
        ck(crane_b200_forward_step(h_, input_ids.data(), input_ids.size(), start_pos, &lg));
        std::vector<float> out(lg.rows * lg.vocab);
        ck(crane_b200_copy_logits(h_, out.data(), out.size()));
        return out;
    }
    void clear_kv_cache() override { ck(crane_b200_clear_kv_cache(h_)); }
    size_t num_layers() const override { return (size_t)crane_b200_num_layers(h_); }
    void warmup() override { ck(crane_b200_warmup(h_)); }
    bool supports_kv_swap() const override { return true; }           // a swap is seq_select(slot): the pages never move
    uint64_t active_kv_cache_bytes() const override { return crane_b200_active_kv_cache_bytes(h_); }
    bool supports_batch_decode() const override { return true; }

    // ---- greedy fast paths (sampling.rs:189-218 `gpu_argmax`; qwen3/model.rs:298-331) ----
    uint32_t forward_step_argmax(const std::vector<uint32_t>& input_ids, size_t start_pos) {
        uint32_t tok = 0;
        ck(crane_b200_forward_step_argmax(h_, input_ids.data(), input_ids.size(), start_pos, &tok));
        return tok;
    }
    std::vector<uint32_t> generate(const std::vector<uint32_t>& prompt, size_t max_new_tokens, const std::vector<uint32_t>& eos = {}) {
        std::vector<uint32_t> out(max_new_tokens);
        size_t n = 0;
        ck(crane_b200_generate_greedy(h_, prompt.data(), prompt.size(), max_new_tokens, eos.data(), eos.size(), out.data(), &n));
        out.resize(n);
        return out;
    }

    // ---- sequence slots + batched decode (setup_batch_decode / step_batch_decode / extract_batch_kv, backend.rs:86-150) ----
    int seq_create() { int s = -1; ck(crane_b200_seq_create(h_, &s)); return s; }
    void seq_free(int s) { ck(crane_b200_seq_free(h_, s)); }
    void seq_select(int s) { ck(crane_b200_seq_select(h_, s)); }
    size_t kv_len() const { return crane_b200_kv_len(h_); }
    // tokens[i] is consumed by seqs[i] at its cached length; returns [n][n_steps] greedy ids
    std::vector<std::vector<uint32_t>> decode_batch(const std::vector<int>& seqs, const std::vector<uint32_t>& tokens, size_t n_steps) {
        std::vector<uint32_t> flat(seqs.size() * n_steps);
        ck(crane_b200_decode_batch(h_, seqs.data(), tokens.data(), seqs.size(), n_steps, flat.data(), nullptr));
        std::vector<std::vector<uint32_t>> out(seqs.size());
        for (size_t i = 0; i < seqs.size(); ++i) out[i].assign(flat.begin() + i * n_steps, flat.begin() + (i + 1) * n_steps);
        return out;
    }

    // ---- prefix sharing and the legacy tensor swap (backend.rs:65-84 get_kv_caches / set_kv_caches) ----
    int seq_fork(int src) { int s = -1; ck(crane_b200_seq_fork(h_, src, &s)); return s; }
    struct LayerCache { std::vector<float> k, v; size_t n_tokens = 0; };      // attention layer: [n_kv, T, D] each
    LayerCache kv_export(int layer, size_t floats_per_tensor) {
        LayerCache c;
        c.k.resize(floats_per_tensor); c.v.resize(floats_per_tensor);
        ck(crane_b200_kv_export(h_, layer, c.k.data(), c.v.data(), floats_per_tensor, &c.n_tokens));
        return c;
    }
    void kv_import(int layer, const LayerCache& c) { ck(crane_b200_kv_import(h_, layer, c.k.data(), c.v.data(), c.n_tokens)); }
    void kv_set_len(size_t n_tokens, uint32_t next_rotary_pos) { ck(crane_b200_kv_set_len(h_, n_tokens, next_rotary_pos)); }

    // ---- the server's sampler on the device (crane-serve/src/engine/sampling.rs:169-480) ----
    static crane_b200_sampling greedy() { crane_b200_sampling p{}; p.repetition_penalty = 1.f; return p; }
    uint32_t sample(const crane_b200_sampling& p) { uint32_t t = 0; ck(crane_b200_sample(h_, &p, &t)); return t; }
    uint32_t forward_step_sample(const std::vector<uint32_t>& input_ids, size_t start_pos, const crane_b200_sampling& p) {
        uint32_t t = 0;
        ck(crane_b200_forward_step_sample(h_, input_ids.data(), input_ids.size(), start_pos, &p, &t));
        return t;
    }
    std::vector<uint32_t> topk(size_t k) { std::vector<uint32_t> idx(k); ck(crane_b200_topk(h_, k, idx.data(), nullptr)); return idx; }

    int vocab_size() const { return crane_b200_vocab_size(h_); }
    int hidden_size() const { return crane_b200_hidden_size(h_); }
    uint64_t kernel_launches() const { return crane_b200_kernel_launches(h_); }
    crane_b200_model* raw() { return h_; }

   private:
    void ck(int rc) const {
        if (rc != CRANE_B200_OK) throw Error(rc, crane_b200_last_error(h_));
    }
    crane_b200_model* h_ = nullptr;
};

}  // namespace crane_b200
