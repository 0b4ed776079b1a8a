//! Raw FFI for libcrane_b200.so -- GENERATED from include/crane_b200.h by tools/gen_rust_sys.py; do not edit.
//! The safe wrapper implementing crane-serve's `ModelBackend` on top of this is sketched in INTEGRATION.md section 2.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct crane_b200_model {
    _private: [u8; 0],
}

pub const CRANE_B200_OK: c_int = 0;
pub const CRANE_B200_INVALID_ARG: c_int = -1;
pub const CRANE_B200_OOM: c_int = -2;
pub const CRANE_B200_CUDA_ERROR: c_int = -3;
pub const CRANE_B200_UNSUPPORTED: c_int = -4;
pub const CRANE_B200_NOT_LOADED: c_int = -5;

pub const CRANE_B200_F32: c_int = 0;
pub const CRANE_B200_BF16: c_int = 1;
pub const CRANE_B200_F16: c_int = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct crane_b200_logits {
    pub device_ptr: *const f32,
    pub rows: usize,
    pub vocab: usize,
    pub stream: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct crane_b200_sampling {
    pub temperature: f32,
    pub top_p: f32,
    pub top_k: i32,
    pub repetition_penalty: f32,
    pub frequency_penalty: f32,
    pub presence_penalty: f32,
    pub context: *const u32,
    pub n_context: usize,
    pub uniforms: *const f32,
    pub seed: u64,
}

#[link(name = "crane_b200")]
extern "C" {
    pub fn crane_b200_create(config_json: *const c_char, device_ordinal: c_int, out: *mut *mut crane_b200_model) -> c_int;
    pub fn crane_b200_destroy(m: *mut crane_b200_model);
    pub fn crane_b200_last_error(m: *const crane_b200_model) -> *const c_char;
    pub fn crane_b200_load_tensor(m: *mut crane_b200_model, name: *const c_char, dtype: c_int, shape: *const i64, ndim: c_int, data: *const c_void) -> c_int;
    pub fn crane_b200_load_tensor_ggml(m: *mut crane_b200_model, name: *const c_char, ggml_type: c_int, shape: *const i64, ndim: c_int, data: *const c_void, nbytes: usize) -> c_int;
    pub fn crane_b200_load_safetensors(m: *mut crane_b200_model, path: *const c_char, n_loaded: *mut usize, n_skipped: *mut usize) -> c_int;
    pub fn crane_b200_load_gguf(m: *mut crane_b200_model, path: *const c_char, n_loaded: *mut usize, n_skipped: *mut usize) -> c_int;
    pub fn crane_b200_gguf_config(path: *const c_char, json_out: *mut c_char, capacity: usize, needed: *mut usize) -> c_int;
    pub fn crane_b200_finalize(m: *mut crane_b200_model) -> c_int;
    pub fn crane_b200_forward_step(m: *mut crane_b200_model, input_ids: *const u32, n: usize, start_pos: usize, out: *mut crane_b200_logits) -> c_int;
    pub fn crane_b200_forward_step_argmax(m: *mut crane_b200_model, input_ids: *const u32, n: usize, start_pos: usize, token_out: *mut u32) -> c_int;
    pub fn crane_b200_clear_kv_cache(m: *mut crane_b200_model) -> c_int;
    pub fn crane_b200_num_layers(m: *const crane_b200_model) -> c_int;
    pub fn crane_b200_warmup(m: *mut crane_b200_model) -> c_int;
    pub fn crane_b200_active_kv_cache_bytes(m: *const crane_b200_model) -> u64;
    pub fn crane_b200_kv_len(m: *const crane_b200_model) -> usize;
    pub fn crane_b200_vocab_size(m: *const crane_b200_model) -> c_int;
    pub fn crane_b200_hidden_size(m: *const crane_b200_model) -> c_int;
    pub fn crane_b200_copy_logits(m: *mut crane_b200_model, host_out: *mut f32, n_floats: usize) -> c_int;
    pub fn crane_b200_forward_embeds(m: *mut crane_b200_model, embeds: *const f32, s: usize, position_ids_3xs: *const u32, start_pos: usize, out: *mut crane_b200_logits) -> c_int;
    pub fn crane_b200_decode_greedy(m: *mut crane_b200_model, first_token: u32, start_pos: usize, n_steps: usize, eos_ids: *const u32, n_eos: usize, tokens_out: *mut u32, n_out: *mut usize) -> c_int;
    pub fn crane_b200_generate_greedy(m: *mut crane_b200_model, prompt: *const u32, n_prompt: usize, max_new_tokens: usize, eos_ids: *const u32, n_eos: usize, tokens_out: *mut u32, n_out: *mut usize) -> c_int;
    pub fn crane_b200_seq_create(m: *mut crane_b200_model, seq_out: *mut c_int) -> c_int;
    pub fn crane_b200_kv_export(m: *mut crane_b200_model, layer: c_int, k_out: *mut f32, v_out: *mut f32, capacity_floats: usize, n_tokens: *mut usize) -> c_int;
    pub fn crane_b200_kv_import(m: *mut crane_b200_model, layer: c_int, k: *const f32, v: *const f32, n_tokens: usize) -> c_int;
    pub fn crane_b200_kv_set_len(m: *mut crane_b200_model, n_tokens: usize, next_rotary_pos: u32) -> c_int;
    pub fn crane_b200_seq_fork(m: *mut crane_b200_model, src: c_int, seq_out: *mut c_int) -> c_int;
    pub fn crane_b200_seq_free(m: *mut crane_b200_model, seq: c_int) -> c_int;
    pub fn crane_b200_seq_select(m: *mut crane_b200_model, seq: c_int) -> c_int;
    pub fn crane_b200_decode_batch(m: *mut crane_b200_model, seqs: *const c_int, tokens: *const u32, n: usize, n_steps: usize, tokens_out: *mut u32, logits_host: *mut f32) -> c_int;
    pub fn crane_b200_comm_unique_id(id_out: *mut u8, capacity: usize) -> c_int;
    pub fn crane_b200_comm_init(m: *mut crane_b200_model, id: *const u8, id_bytes: usize, rank: c_int, world: c_int) -> c_int;
    pub fn crane_b200_comm_world(m: *const crane_b200_model, rank_out: *mut c_int, world_out: *mut c_int) -> c_int;
    pub fn crane_b200_decode_batch_gather(m: *mut crane_b200_model, seqs: *const c_int, tokens: *const u32, n: usize, tokens_all_out: *mut u32, logits_all: *mut crane_b200_logits) -> c_int;
    pub fn crane_b200_copy_gathered_logits(m: *mut crane_b200_model, host_out: *mut f32, n_floats: usize) -> c_int;
    pub fn crane_b200_sample(m: *mut crane_b200_model, p: *const crane_b200_sampling, token_out: *mut u32) -> c_int;
    pub fn crane_b200_forward_step_sample(m: *mut crane_b200_model, input_ids: *const u32, n: usize, start_pos: usize, p: *const crane_b200_sampling, token_out: *mut u32) -> c_int;
    pub fn crane_b200_decode_batch_sample(m: *mut crane_b200_model, seqs: *const c_int, tokens: *const u32, n: usize, params: *const crane_b200_sampling, tokens_out: *mut u32) -> c_int;
    pub fn crane_b200_topk(m: *mut crane_b200_model, k: usize, idx_out: *mut u32, vals_out: *mut f32) -> c_int;
    pub fn crane_b200_encode_images(m: *mut crane_b200_model, pixel_values: *const f32, grid_thw: *const u32, n_images: usize, image_embeds_out: *mut f32, deepstack_out: *mut f32) -> c_int;
    pub fn crane_b200_vl_forward(m: *mut crane_b200_model, input_ids: *const u32, n: usize, pixel_values: *const f32, grid_thw: *const u32, n_images: usize, start_pos: usize, out: *mut crane_b200_logits) -> c_int;
    pub fn crane_b200_vl_decode_step(m: *mut crane_b200_model, token: u32, start_pos: usize, out: *mut crane_b200_logits) -> c_int;
    pub fn crane_b200_vl_decode_step_argmax(m: *mut crane_b200_model, token: u32, start_pos: usize, token_out: *mut u32) -> c_int;
    pub fn crane_b200_next_mrope_pos(m: *const crane_b200_model) -> u32;
    pub fn crane_b200_tts_text_project(m: *mut crane_b200_model, text_ids: *const u32, n: usize, out_host: *mut f32) -> c_int;
    pub fn crane_b200_tts_codec_embed(m: *mut crane_b200_model, group: c_int, ids: *const u32, n: usize, out_host: *mut f32) -> c_int;
    pub fn crane_b200_tts_prefill(m: *mut crane_b200_model, embeds: *const f32, prefill_len: usize, trailing_text: *const f32, n_trailing: usize, tts_pad_embed: *const f32) -> c_int;
    pub fn crane_b200_tts_generate(m: *mut crane_b200_model, max_frames: usize, repetition_penalty: f32, forced_frames: *const u32, frames_out: *mut u32, n_frames_out: *mut usize, first_logits_out: *mut f32, group_logits_out: *mut f32) -> c_int;
    pub fn crane_b200_last_timing(m: *const crane_b200_model, prefill_ms: *mut f32, decode_ms: *mut f32, decode_steps: *mut usize) -> c_int;
    pub fn crane_b200_kernel_launches(m: *const crane_b200_model) -> u64;
    pub fn crane_b200_prof_enable(m: *mut crane_b200_model, on: c_int) -> c_int;
    pub fn crane_b200_prof_report(m: *const crane_b200_model, buf: *mut c_char, cap: usize, needed: *mut usize) -> c_int;
    pub fn crane_b200_decode_path(m: *const crane_b200_model) -> c_int;
    pub fn crane_b200_op_gemm(device: c_int, a: *const u16, a_lo: *const u16, w: *const u16, M: c_int, N: c_int, K: c_int, mode: c_int, bias: *const f32, out_inout: *mut c_void, use_simt: c_int) -> c_int;
    pub fn crane_b200_op_qlinear(device: c_int, x: *const f32, m: usize, k: usize, raw: *const c_void, raw_bytes: usize, ggml_type: c_int, n: usize, norm_w: *const f32, eps: f32, y: *mut f32) -> c_int;
    pub fn crane_b200_op_topk(device: c_int, logits: *const f32, vocab: usize, k: usize, idx_out: *mut u32) -> c_int;
    pub fn crane_b200_op_sample(device: c_int, logits: *const f32, vocab: usize, p: *const crane_b200_sampling, token_out: *mut u32, logits_after: *mut f32) -> c_int;
}
