/*
 * crane_b200 -- C ABI of the B200-native forward-pass engine that sits behind crane-core's model API.
 *
 * Every entry point below is what the reference's Rust side would bind through `extern "C"` for this
 * path; the reference interface each one replaces is cited as file:line under /root/reference.
 * The Rust-side binding a Crane maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary;
 *   - every function returns 0 on success or a negative crane_b200_status; it never throws/aborts;
 *     `crane_b200_last_error()` gives the message (per handle; handle NULL = last create() failure);
 *   - a handle is externally synchronised (one caller at a time, the engine thread of
 *     crane-serve/src/engine/mod.rs:169-271) but may migrate between threads;
 *   - the library owns weights, KV pages, workspaces and the logits buffer (valid until the next call
 *     on the same handle); the caller owns every input array (consumed before return);
 *   - there is no CPU fallback: without a CUDA device create() fails with CRANE_B200_CUDA_ERROR.
 */
#ifndef CRANE_B200_H
#define CRANE_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CRANE_B200_API __attribute__((visibility("default")))
#else
#define CRANE_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crane_b200_model crane_b200_model;

typedef enum {
    CRANE_B200_OK = 0,
    CRANE_B200_INVALID_ARG = -1,
    CRANE_B200_OOM = -2,
    CRANE_B200_CUDA_ERROR = -3,
    CRANE_B200_UNSUPPORTED = -4,
    CRANE_B200_NOT_LOADED = -5
} crane_b200_status;

typedef enum {
    CRANE_B200_F32 = 0,
    CRANE_B200_BF16 = 1,
    CRANE_B200_F16 = 2
} crane_b200_dtype;

/* Logits of the LAST position of the most recent forward call (f32, device memory, `stream` ordered).
 * Shape [rows, vocab]; rows == 1 for the single-sequence calls.
 * Mirrors the `Tensor` returned by `Model::forward_step` (crane-core/src/models/qwen3/model.rs:34-268;
 * logits [1,1,V], qwen3/modeling.rs:1032-1035). */
typedef struct {
    const float* device_ptr;
    size_t rows;
    size_t vocab;
    void* stream; /* cudaStream_t the result is ordered on */
} crane_b200_logits;

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* Create an engine for one model on one GPU.  `config_json` is the checkpoint's HF config.json text
 * (qwen3: crane-core/src/models/qwen3/modeling.rs:94-130 `Config`; qwen3_vl: text_config + vision_config
 * as crane-core/src/models/qwen3_5/config.rs:112-137) optionally extended with an "engine" object:
 *   {"max_seq_len": 4096, "max_batch": 1, "gemm": "tcgen05"|"simt", "graphs": true, "pdl": true,
 *    "precision": "split"|"bf16", "kv_cache": "fp"|"int8"|"int4", "persistent": true, "vit_act": "erf"|"tanh", "merger_act": "tanh"|"erf",
 *    "gdn": "auto"|"chunked"|"sequential"}
 * "gdn" (Qwen3.5 linear-attention layers): how a prefill call evaluates the gated delta rule (crane-core/src/ops/gdn/backend.rs:90-156).
 * "sequential" = token by token, as the reference does; "auto" (default: calls of 128 rows or more) and "chunked" (64 or more) run the
 * chunkwise form (64 tokens per serial step, tensor cores; same recurrence, results within 1e-4 of the sequential kernel --
 * the reference's own bar for that pair, crane-core/tests/rocm_kernels.rs:39-84); shorter calls and decode stay sequential.
 * linear_key_head_dim may be 64, 128 or 256 (the reference's kernel accepts K <= 256, kernels/cuda/gdn.cu:45-153).
 * "kv_cache": "int8" / "int4" = the reference's QuantKvCache (crane-core/src/models/qwen3_5/kv_cache.rs:209-342): per (token, KV
 * head) symmetric codes + an f32 scale, 8.25 / 4.25 bits per cached element instead of 32 (split) or 16; attention reads
 * code * scale, dequantised while the decode kernel stages its tile.
 * "precision": "split" (default) carries every bf16 tensor-core operand and KV page as a hi + lo pair (~16 mantissa bits:
 * logits within ~3e-5 of the f32 CPU path); "bf16" is the plain-bf16 fast mode (~1e-2, what the reference's GPU path does).
 * Replaces `model_factory::create_backend` / `Qwen3Backend::new`
 * (crane-serve/src/engine/model_factory.rs:471-560, backend.rs:615-625). */
CRANE_B200_API int crane_b200_create(const char* config_json, int device_ordinal, crane_b200_model** out);
CRANE_B200_API void crane_b200_destroy(crane_b200_model* m);
CRANE_B200_API const char* crane_b200_last_error(const crane_b200_model* m);

/* Register one checkpoint tensor by its safetensors name (`model.layers.{i}.self_attn.q_proj.weight`,
 * `model.language_model...`, `model.visual.blocks.{i}.attn.qkv.weight`, `lm_head.weight`, ... exactly
 * the names read at crane-core/src/models/qwen3/modeling.rs:160-282,598-606,771-813 and
 * qwen3_5/vision.rs:23-35,70-71,114-115,245-250,321-339).  `data` is host memory. */
CRANE_B200_API int crane_b200_load_tensor(crane_b200_model* m, const char* name, int dtype, const int64_t* shape, int ndim,
                           const void* data);
/* Register a GGUF-quantised 2-D tensor as raw ggml blocks (ggml_type: Q8_0 = 8, Q4_K = 12, Q6_K = 14), shape = [rows, cols]
 * (cols % 256 == 0).  Accepts the HF names above or the GGUF names the reference's GGUF loaders read
 * (`blk.{i}.attn_q.weight`, `token_embd.weight`, `output.weight`, ...: crane-core/src/models/qwen3/modeling.rs:252-282,598-606,
 * 898-916).  Replaces `Gguf::linear` -> `LinearLayer::Quantized(QMatMul)` (crane-core/src/models/hunyuan_dense/modeling.rs:37-41,
 * crane-core/src/ops/linear.rs:23-48): decode streams the quantised bytes, prefill dequantises one matrix at a time. */
CRANE_B200_API int crane_b200_load_tensor_ggml(crane_b200_model* m, const char* name, int ggml_type, const int64_t* shape, int ndim,
                                               const void* data, size_t nbytes);
/* Whole-file readers: every tensor of a .safetensors file (`VarBuilder::from_mmaped_safetensors`,
 * crane-core/src/models/qwen3/model.rs:91-92) or of a GGUF v2/v3 file (`gguf_file::Content::read` + `Qwen3Model::from_gguf`,
 * crane-core/src/models/qwen3/modeling.rs:821-935) is registered under its own name; quantised GGUF tensors keep their blocks.
 * Tensors the engine has no use for are skipped and counted in *n_skipped (both counters optional).  The model config still
 * comes from crane_b200_create (GGUF metadata is not interpreted).  Call once per shard, then crane_b200_finalize. */
CRANE_B200_API int crane_b200_load_safetensors(crane_b200_model* m, const char* path, size_t* n_loaded, size_t* n_skipped);
CRANE_B200_API int crane_b200_load_gguf(crane_b200_model* m, const char* path, size_t* n_loaded, size_t* n_skipped);
/* The config.json text of a GGUF checkpoint, derived from its metadata with the reference's recipe (`Qwen3Model::from_gguf`,
 * crane-core/src/models/qwen3/modeling.rs:821-905: head counts, key_length default 128, block_count, embedding_length,
 * feed_forward_length, context_length, rms epsilon, rope base; vocab from token_embd.weight; tied iff no output.weight).
 * Pure host code (no GPU, no handle): feed the text to crane_b200_create, then crane_b200_load_gguf the same file.
 * *needed receives the byte count including the terminator; errors are reported through crane_b200_last_error(NULL). */
CRANE_B200_API int crane_b200_gguf_config(const char* path, char* json_out, size_t capacity, size_t* needed);
/* All tensors registered: merge QKV / gate-up, build rotary tables, allocate KV pages + workspaces. */
CRANE_B200_API int crane_b200_finalize(crane_b200_model* m);

/* ---- ModelBackend surface (crane-serve/src/engine/backend.rs:30-151) ---------------------------- */

/* `forward_step(&mut self, input_ids: &[u32], start_pos: usize) -> Result<Tensor>` (backend.rs:42). */
CRANE_B200_API int crane_b200_forward_step(crane_b200_model* m, const uint32_t* input_ids, size_t n, size_t start_pos,
                            crane_b200_logits* out);
/* Same pass + device-side greedy argmax; only 4 bytes return to the host.  The server's greedy fast
 * path: `gpu_argmax` (crane-serve/src/engine/sampling.rs:189-218,
 * crane-core/src/ops/fused_ops/cuda_impl.rs:204-282).  Ties resolve to the LOWEST index. */
CRANE_B200_API int crane_b200_forward_step_argmax(crane_b200_model* m, const uint32_t* input_ids, size_t n, size_t start_pos,
                                   uint32_t* token_out);
/* `clear_kv_cache` (backend.rs:45), `num_layers` (:48), `warmup` (:61). */
CRANE_B200_API int crane_b200_clear_kv_cache(crane_b200_model* m);
CRANE_B200_API int crane_b200_num_layers(const crane_b200_model* m);
CRANE_B200_API int crane_b200_warmup(crane_b200_model* m);
/* `active_kv_cache_bytes` (backend.rs:82-84) and the cached length (`Qwen3Model::kv_cache_len`,
 * qwen3/modeling.rs:1068). */
CRANE_B200_API uint64_t crane_b200_active_kv_cache_bytes(const crane_b200_model* m);
CRANE_B200_API size_t crane_b200_kv_len(const crane_b200_model* m);
CRANE_B200_API int crane_b200_vocab_size(const crane_b200_model* m);
CRANE_B200_API int crane_b200_hidden_size(const crane_b200_model* m);

/* Copy the logits of the last call to host memory (`logits.to_dtype(F32)?.to_vec1()`, model.rs:305). */
CRANE_B200_API int crane_b200_copy_logits(crane_b200_model* m, float* host_out, size_t n_floats);

/* ---- library surface (crane-core/src/models/qwen3/model.rs) ------------------------------------- */

/* `Qwen3Model::forward_embeds` (qwen3/modeling.rs:964-978): caller-supplied embeddings [s, hidden] f32
 * (host).  `position_ids_3xs` is NULL for 1-D positions start_pos..start_pos+s-1, else the [3, s] MRoPE
 * ids of `MRotaryEmbedding::cos_sin_with_position_ids` (qwen3_5/modeling.rs:172-245). */
CRANE_B200_API int crane_b200_forward_embeds(crane_b200_model* m, const float* embeds, size_t s, const uint32_t* position_ids_3xs,
                              size_t start_pos, crane_b200_logits* out);

/* On-device greedy decode loop: `n_steps` dependent single-token passes without returning to the host
 * (the inner loop of `Model::generate`, qwen3/model.rs:298-331 with temperature None, and of
 * `step_decode_batch`'s `decode_tokens_per_seq` rounds, crane-serve/src/engine/mod.rs:898-1008).
 * `first_token` is consumed at cache position `start_pos` (== kv_len).  Writes min(n_steps, up to and
 * including the first EOS) tokens to `tokens_out`, the count to `n_out`.  The KV cache afterwards holds
 * start_pos + n_steps positions. */
CRANE_B200_API int crane_b200_decode_greedy(crane_b200_model* m, uint32_t first_token, size_t start_pos, size_t n_steps,
                             const uint32_t* eos_ids, size_t n_eos, uint32_t* tokens_out, size_t* n_out);

/* `ModelForCausalLM::generate` greedy branch (crane-core/src/generation/based.rs:5-34,
 * qwen3/model.rs:275-349): clear cache, prefill `prompt`, then decode until EOS / max_new_tokens. */
CRANE_B200_API int crane_b200_generate_greedy(crane_b200_model* m, const uint32_t* prompt, size_t n_prompt, size_t max_new_tokens,
                               const uint32_t* eos_ids, size_t n_eos, uint32_t* tokens_out, size_t* n_out);

/* ---- sequence slots + batched decode (crane-serve continuous batching) ------------------------------- */

/* A handle holds `engine.max_batch` sequence slots, each with its own KV pages (slot 0 is the implicit sequence of all the calls
 * above).  `seq_select` makes a slot current: forward_step / forward_embeds / clear_kv_cache / kv_len then act on it -- the
 * handle-level counterpart of the engine's per-Sequence swap-in (`set_kv_caches`, crane-serve/src/engine/mod.rs:1172). */
CRANE_B200_API int crane_b200_seq_create(crane_b200_model* m, int* seq_out);
/* KV swap, the legacy surface (ModelBackend::get_kv_caches / set_kv_caches, crane-serve/src/engine/backend.rs:65-84; per-layer
 * `(Tensor, Tensor)` of shape [1, n_kv, T, d]): the CURRENT sequence's cache of one layer as contiguous host tensors [n_kv, T, D]
 * f32.  export with k_out == v_out == NULL only reports *n_tokens.  For a Gated-Delta-Net layer of the hybrid model the pair is
 * (conv window [conv_dim, 4], recurrent state [n_v, d_k, d_v]) (ops/gdn/cache.rs:15-55) and *n_tokens is 0.  After importing every
 * layer, kv_set_len states the cached length and the next rotary position.  Synchronous; O(context) per call. */
CRANE_B200_API int crane_b200_kv_export(crane_b200_model* m, int layer, float* k_out, float* v_out, size_t capacity_floats, size_t* n_tokens);
CRANE_B200_API int crane_b200_kv_import(crane_b200_model* m, int layer, const float* k, const float* v, size_t n_tokens);
CRANE_B200_API int crane_b200_kv_set_len(crane_b200_model* m, size_t n_tokens, uint32_t next_rotary_pos);
/* A new sequence that starts as a copy of `src` (KV pages of every attention layer, the Gated-Delta-Net conv / recurrent state of
 * a hybrid model, cached length and rotary position): prefix sharing for the server's scheduler, what the reference does by cloning
 * its per-sequence caches (backend.rs:65-84 get_kv_caches / set_kv_caches).  Device-to-device, O(context), on the engine stream. */
CRANE_B200_API int crane_b200_seq_fork(crane_b200_model* m, int src, int* seq_out);
CRANE_B200_API int crane_b200_seq_free(crane_b200_model* m, int seq);
CRANE_B200_API int crane_b200_seq_select(crane_b200_model* m, int seq);
/* `n_steps` greedy decode rounds for `n` sequences at once: replaces setup_batch_decode + step_batch_decode x rounds +
 * extract_batch_kv (crane-core/src/models/qwen3/modeling.rs:1141-1277, crane-serve/src/engine/mod.rs:822-1062).  No left-padding,
 * mask or KV copy: each sequence attends over its own pages with its own length and position; groups of up to 4 sequences share
 * one pass over the weights.  tokens[i] is consumed by seqs[i] at its cached length; tokens_out is [n, n_steps];
 * logits_host (optional) receives the [n, vocab] f32 logits of the LAST round. */
CRANE_B200_API int crane_b200_decode_batch(crane_b200_model* m, const int* seqs, const uint32_t* tokens, size_t n, size_t n_steps,
                                           uint32_t* tokens_out, float* logits_host);

/* ---- multi-GPU batch decode (BASELINE.json config 4: 32 sequences over 8 x B200) ----------------------------------------------
 * One process and one handle per GPU, every rank holding the full weights and its share of the sequences; the only data-path
 * exchange is the all-gather of each round's results (NCCL over NVLink, enqueued on the engine stream behind the lm_head).
 * NCCL (libnccl.so.2) is bound at run time by the first of these calls; a single-GPU deployment never loads it.
 *   rank 0: crane_b200_comm_unique_id(id, 128) -> ship the 128 bytes to the other ranks (the launcher's job: file, TCP store, MPI)
 *   every rank: crane_b200_comm_init(h, id, 128, rank, world)            (collective: returns when all ranks have joined)
 *   every round: crane_b200_decode_batch_gather(h, seqs, tokens, n, tokens_all, &logits_all)   (collective, the same n everywhere)
 * tokens_all [world * n]: greedy token of every sequence of every rank, rank-major; logits_all (nullable): [world * n, V] f32 in
 * device memory on every rank -- what `step_batch_decode` (crane-serve/src/engine/backend.rs:86-150) returns for the whole batch. */
CRANE_B200_API int crane_b200_comm_unique_id(uint8_t* id_out, size_t capacity);
CRANE_B200_API int crane_b200_comm_init(crane_b200_model* m, const uint8_t* id, size_t id_bytes, int rank, int world);
CRANE_B200_API int crane_b200_comm_world(const crane_b200_model* m, int* rank_out, int* world_out);
CRANE_B200_API int crane_b200_decode_batch_gather(crane_b200_model* m, const int* seqs, const uint32_t* tokens, size_t n,
                                                  uint32_t* tokens_all_out, crane_b200_logits* logits_all);
/* Host copy of the rows gathered by the last decode_batch_gather that asked for logits ([world * n, V] f32; tests, host samplers). */
CRANE_B200_API int crane_b200_copy_gathered_logits(crane_b200_model* m, float* host_out, size_t n_floats);

/* ---- device-side sampling (crane-serve/src/engine/sampling.rs:169-480) ------------------------------------------ */
/* Per-sequence sampling request = the fields of `Sequence` that `sampling::sample` reads.  The draw needs uniforms in
 * (1e-7, 0.999) (`rand_like(1e-7, 0.999)`, sampling.rs:387): the reference takes them from candle's device RNG, whose stream no
 * test pins, so they are an input here -- `uniforms` (host, top_k values, or vocab values when neither top-k nor top-p applies), or
 * NULL to have them derived on the device from `seed`. */
typedef struct {
    float temperature;          /* <= 0: greedy (argmax after penalties, lowest index among maxima) */
    float top_p;                /* in (0, 1): nucleus over the top-k candidates; anything else: off */
    int32_t top_k;              /* 0: unset -- 64 when top_p is active (CRANE_TOPP_FALLBACK_TOPK), else the whole vocabulary; capped at 64 */
    float repetition_penalty;   /* 1 = off: logit >= 0 ? logit / p : logit * p, once per distinct context token */
    float frequency_penalty;    /* 0 = off: minus count * frequency_penalty */
    float presence_penalty;     /* 0 = off: minus presence_penalty for any token present */
    const uint32_t* context;    /* host: the trailing `repeat_last_n` tokens of the sequence */
    size_t n_context;
    const float* uniforms;      /* host or NULL */
    uint64_t seed;
} crane_b200_sampling;
/* `sampling::sample` on the logits of the LAST forward call of the current sequence (they stay on the device; 4 bytes return). */
CRANE_B200_API int crane_b200_sample(crane_b200_model* m, const crane_b200_sampling* p, uint32_t* token_out);
/* forward_step + sample in one call: the server's non-greedy decode step (crane-serve/src/engine/mod.rs:938-1062). */
CRANE_B200_API int crane_b200_forward_step_sample(crane_b200_model* m, const uint32_t* input_ids, size_t n, size_t start_pos,
                                                  const crane_b200_sampling* p, uint32_t* token_out);
/* One decode round for `n` sequences with per-sequence sampling (params[n]): step_batch_decode + sample per row, logits never
 * leave the device; tokens_out [n]. */
CRANE_B200_API int crane_b200_decode_batch_sample(crane_b200_model* m, const int* seqs, const uint32_t* tokens, size_t n,
                                                  const crane_b200_sampling* params, uint32_t* tokens_out);
/* `crane_core::ops::topk_indices` (ops/fused_ops/portable.rs:28-66, kernels/cuda/topk.cu:213-259) on the logits of the last call:
 * the k largest in the total order (value descending, index ascending), k <= 512.  vals_out is optional. */
CRANE_B200_API int crane_b200_topk(crane_b200_model* m, size_t k, uint32_t* idx_out, float* vals_out);

/* ---- vision-language surface (crane-core/src/models/qwen3_5/vlm.rs) ------------------------------ */

/* `Qwen3_5VLModel::encode_images` -> `Qwen3_5VisionModel::forward` (vlm.rs:150-170, vision.rs:558-584).
 * pixel_values [sum(t*h*w), C*T*P*P] f32 host, grid_thw [n_images, 3].  Optional host outputs:
 * image_embeds_out [sum(t*h*w)/merge^2, out_hidden], deepstack_out [n_deepstack, same rows, out_hidden]. */
CRANE_B200_API int crane_b200_encode_images(crane_b200_model* m, const float* pixel_values, const uint32_t* grid_thw, size_t n_images,
                             float* image_embeds_out, float* deepstack_out);
/* `Qwen3_5VLModel::forward` (vlm.rs:250-285): ViT + embed + splice + 3-axis position ids + decoder
 * (+ DeepStack, qwen3_vl/text.rs:252-270).  `pixel_values` may be NULL (text-only prefill). */
CRANE_B200_API int crane_b200_vl_forward(crane_b200_model* m, const uint32_t* input_ids, size_t n, const float* pixel_values,
                          const uint32_t* grid_thw, size_t n_images, size_t start_pos, crane_b200_logits* out);
/* `Qwen3_5VLModel::decode_step` (vlm.rs:294-301): one token at cache position start_pos with the scalar
 * MRoPE counter seeded by vl_forward. */
CRANE_B200_API int crane_b200_vl_decode_step(crane_b200_model* m, uint32_t token, size_t start_pos, crane_b200_logits* out);
/* The same step returning only the greedy token (4 bytes back instead of the logits row): the serving loop's call. */
CRANE_B200_API int crane_b200_vl_decode_step_argmax(crane_b200_model* m, uint32_t token, size_t start_pos, uint32_t* token_out);
/* The MRoPE counter (`next_mrope_pos`, vlm.rs:272-282). */
CRANE_B200_API uint32_t crane_b200_next_mrope_pos(const crane_b200_model* m);

/* ---- Qwen3-TTS codec-LM surface (crane-core/src/models/qwen3_tts/modeling.rs) ------------------------------------ */
/* A `qwen3_tts` config (talker_config + code_predictor_config) makes the handle a talker with an embedded code predictor;
 * tensors are registered under their checkpoint names (`talker.model.*`, `talker.codec_head.*`, `talker.text_projection.*`,
 * `talker.code_predictor.*`, modeling.rs:297-345,513-575).
 * text_embedding -> ResizeMlp (fc1+bias, SiLU, fc2+bias) for `n` text ids -> [n, hidden] f32 host (modeling.rs:244-268,612-614). */
CRANE_B200_API int crane_b200_tts_text_project(crane_b200_model* m, const uint32_t* text_ids, size_t n, float* out_host);
/* Rows of the talker codec_embedding (group = -1) or of code-predictor codec_embedding[group] -> [n, hidden] f32 host. */
CRANE_B200_API int crane_b200_tts_codec_embed(crane_b200_model* m, int group, const uint32_t* ids, size_t n, float* out_host);
/* Talker prefill over caller-assembled embeddings (`build_prefill_embeds`, modeling.rs:597-726, stays host glue) and the per-step
 * text contributions of the frame loop: trailing_text [n_trailing, hidden] then the tts_pad row (modeling.rs:1546-1552). */
CRANE_B200_API int crane_b200_tts_prefill(crane_b200_model* m, const float* embeds, size_t prefill_len, const float* trailing_text,
                                          size_t n_trailing, const float* tts_pad_embed);
/* The frame loop of `generate_speech_codes` / `generate_one_frame` (modeling.rs:1492-1568,1677-1749): per frame the first
 * code from codec_head (repetition penalty over previous first codes, suppress window [V-1024, V) \ {EOS}, EOS held back for
 * 2 frames), 15 code-predictor passes (`CodePredictor::predict`, :373-479), the summed-embedding next input and one talker
 * pass -- all on the device, greedy (the reference's seeded host sampler is not pinned) or teacher-forced from
 * `forced_frames` [max_frames, groups].  frames_out [max_frames, groups]; optional logits for parity:
 * first_logits_out [max_frames, codec_vocab] (raw codec_head logits), group_logits_out [max_frames, groups-1, cp_vocab]. */
CRANE_B200_API int crane_b200_tts_generate(crane_b200_model* m, size_t max_frames, float repetition_penalty, const uint32_t* forced_frames,
                                           uint32_t* frames_out, size_t* n_frames_out, float* first_logits_out, float* group_logits_out);

/* ---- profiling (crane-core/src/ops/prof.rs:37-243 `CRANE_PROF`) ----------------------------------- */
/* Device time (ms, CUDA events on the engine stream) of the last prefill pass and last decode run. */
CRANE_B200_API int crane_b200_last_timing(const crane_b200_model* m, float* prefill_ms, float* decode_ms, size_t* decode_steps);
/* Kernels launched by this handle since creation (graph replays count their nodes). */
CRANE_B200_API uint64_t crane_b200_kernel_launches(const crane_b200_model* m);
/* Forward-pass profiling, the counterpart of crane-core/src/ops/prof.rs:1-61,99-197: `CRANE_PROF=1` in the environment (summary
 * line on stderr every `CRANE_PROF_EVERY` passes, default 64) or this switch.  Every pass is timed twice on the host (`enqueue`:
 * last launch submitted, `wall`: after a stream sync) and split into the reference's stage names (embed norm attn gdn mlp resid
 * head | proj conv qkv recur finish | prep launch post) with DEVICE time from cudaEvents beside the host submission time.
 * prof_report writes the running totals as JSON ({"decode": {...}, "prefill": {...}}, per pass); *needed = bytes incl. the NUL.
 * A profiled pass syncs and loses kernel overlap: never quote its times as benchmark numbers. */
CRANE_B200_API int crane_b200_prof_enable(crane_b200_model* m, int on);
CRANE_B200_API int crane_b200_prof_report(const crane_b200_model* m, char* buf, size_t cap, size_t* needed);
/* Which decode path single-sequence greedy decode takes after finalize: 1 = the persistent kernel (one launch for n steps),
 * 0 = the kernel chain (one graph replay per step: hybrid / quantised / TTS models, batched sequences, borrowed streams). */
CRANE_B200_API int crane_b200_decode_path(const crane_b200_model* m);

/* ---- kernel-level test hooks (used by tests/ only; host buffers in, host buffers out) ------------ */
/* C[M,N] = epilogue((A + A_lo)[M,K] bf16 x W[N,K]^T bf16); a_lo = optional low-order plane of a split-precision A
 * (NULL: plain bf16); mode = cb::GemmEpiMode; out dtype follows the mode. */
CRANE_B200_API int crane_b200_op_gemm(int device, const uint16_t* a, const uint16_t* a_lo, const uint16_t* w, int M, int N, int K, int mode,
                       const float* bias, void* out_inout, int use_simt);

/* y[m, n] = quantised linear of x[m, k] (host f32) through the decode kernels (xquant + qgemv, rows in groups of <= 4): `raw` = n rows
 * of ggml blocks (ggml_type 8 / 12 / 14), optional RMSNorm weight norm_w[k] applied first (eps).  = candle's CPU `QMatMul::forward`
 * (crane-core/src/ops/linear.rs:23-48). */
CRANE_B200_API int crane_b200_op_qlinear(int device, const float* x, size_t m, size_t k, const void* raw, size_t raw_bytes, int ggml_type, size_t n,
                                         const float* norm_w, float eps, float* y);
/* top-k order / sampler on caller-supplied logits [rows = 1, vocab] (host): the reference's own known-answer vectors run through these
 * (crane-core/tests/rocm_kernels.rs:96-198, crane-serve/src/engine/sampling.rs:489-640).  logits_after (optional, host) receives the
 * row after the penalties. */
CRANE_B200_API int crane_b200_op_topk(int device, const float* logits, size_t vocab, size_t k, uint32_t* idx_out);
CRANE_B200_API int crane_b200_op_sample(int device, const float* logits, size_t vocab, const crane_b200_sampling* p, uint32_t* token_out,
                                        float* logits_after);

#ifdef __cplusplus
}
#endif
#endif /* CRANE_B200_H */
