"""CPU: the C-ABI library loads, exports every symbol include/crane_b200.h declares, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest

import crane_b200
from crane_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "crane_b200.h")).read()
    return sorted(set(re.findall(r"CRANE_B200_API[^;(]*?\b(crane_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(crane_b200.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/crane_b200.h but not exported"
    assert sorted(crane_b200._SIGNATURES) == names, "python binding and header disagree"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(crane_b200.CraneB200Error) as e:
        crane_b200.Qwen3Model(synth.TINY_QWEN3)
    assert e.value.code == crane_b200.CUDA_ERROR and "no CPU fallback" in str(e.value)


def test_null_and_bad_config_are_errors_not_crashes():
    lib = crane_b200.load_library()
    h = ctypes.c_void_p()
    assert lib.crane_b200_create(None, 0, ctypes.byref(h)) == crane_b200.INVALID_ARG
    assert lib.crane_b200_forward_step(None, None, 0, 0, None) == crane_b200.INVALID_ARG
    assert lib.crane_b200_num_layers(None) == 0
    lib.crane_b200_destroy(None)


def test_gguf_metadata_to_config_is_pure_host_code(tmp_path):
    """`crane_b200_gguf_config` follows Qwen3Model::from_gguf's recipe (qwen3/modeling.rs:821-905) and needs no GPU."""
    import gguf
    import numpy as np
    path = str(tmp_path / "meta.gguf")
    wr = gguf.GGUFWriter(path, "qwen3")
    wr.add_uint32("qwen3.block_count", 3)
    wr.add_uint32("qwen3.embedding_length", 256)
    wr.add_uint32("qwen3.feed_forward_length", 512)
    wr.add_uint32("qwen3.attention.head_count", 8)
    wr.add_uint32("qwen3.attention.head_count_kv", 2)
    wr.add_float32("qwen3.rope.freq_base", 10000.0)
    wr.add_float32("qwen3.attention.layer_norm_rms_epsilon", 1e-5)
    wr.add_array("tokenizer.ggml.tokens", ["a", "b", "c"])
    wr.add_tensor("token_embd.weight", np.zeros((1000, 256), np.float32))
    wr.add_tensor("blk.0.attn_q_norm.weight", np.ones(128, np.float32))
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()
    cfg = crane_b200.gguf_config(path)
    assert cfg["model_type"] == "qwen3" and cfg["vocab_size"] == 1000 and cfg["hidden_size"] == 256 and cfg["intermediate_size"] == 512
    assert cfg["num_hidden_layers"] == 3 and cfg["num_attention_heads"] == 8 and cfg["num_key_value_heads"] == 2
    assert cfg["head_dim"] == 128 and cfg["max_position_embeddings"] == 32768           # the reference's defaults
    assert abs(cfg["rope_theta"] - 1e4) < 1e-3 and abs(cfg["rms_norm_eps"] - 1e-5) < 1e-9
    assert cfg["tie_word_embeddings"] is True and cfg["use_qk_norm"] is True
    # missing required key / not a GGUF file: error code + message, no crash
    bad = str(tmp_path / "bad.gguf")
    open(bad, "wb").write(b"not a gguf file at all")
    import pytest
    with pytest.raises(crane_b200.CraneB200Error):
        crane_b200.gguf_config(bad)


def test_every_entry_point_survives_null_arguments():
    """No entry point may crash, abort or throw across the ABI: all-NULL / zero arguments give an error status (or a neutral 0)."""
    lib = crane_b200.load_library()
    for name, (res, args) in crane_b200._SIGNATURES.items():
        call = []
        for a in args:
            if a in (ctypes.c_int, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64):
                call.append(0)
            elif a in (ctypes.c_float, ctypes.c_double):
                call.append(0.0)
            else:
                call.append(None)
        r = getattr(lib, name)(*call)
        if res is ctypes.c_int:
            assert r <= 0, f"{name}(NULL...) returned {r}"
        elif res in (ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint32):
            assert r == 0, f"{name}(NULL...) returned {r}"


def test_rust_sys_crate_declares_every_entry_point():
    """rust/crane-b200-sys/src/lib.rs is generated from include/crane_b200.h (tools/gen_rust_sys.py): regenerating reproduces the
    committed file, and every symbol the header declares has exactly one `pub fn` there (no Rust toolchain in this image, so the
    crate is checked textually)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0
    hdr = open(os.path.join(root, "include", "crane_b200.h")).read()
    rs = open(os.path.join(root, "rust", "crane-b200-sys", "src", "lib.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"CRANE_B200_API[^;(]*?\b(crane_b200_\w+)\s*\(", hdr)
    assert len(names) >= 49 and len(set(names)) == len(names)
    for n in names:
        assert len(re.findall(r"pub fn " + n + r"\(", rs)) == 1, n


def test_gguf_config_for_the_qwen35_architecture(tmp_path):
    """crane_b200_gguf_config on llama.cpp `qwen35` metadata (Qwen3_5Model::from_gguf, qwen3_5/model.rs:155-325) -- pure host code:
    layer layout from `blk.{i}.ssm_a` presence, value-head size from ssm.inner_size / ssm.time_step_rank, partial rotary factor
    from rope.dimension_count, mrope sections from an ARRAY value, tie_word_embeddings from the absence of output.weight."""
    import gguf
    import numpy as np
    import crane_b200
    path = str(tmp_path / "meta.gguf")
    wr = gguf.GGUFWriter(path, "qwen35")
    for k, v in (("block_count", 4), ("embedding_length", 256), ("feed_forward_length", 512), ("attention.head_count", 4),
                 ("attention.head_count_kv", 2), ("attention.key_length", 256), ("rope.dimension_count", 64), ("ssm.conv_kernel", 4),
                 ("ssm.state_size", 128), ("ssm.group_count", 2), ("ssm.time_step_rank", 4), ("ssm.inner_size", 512)):
        wr.add_uint32("qwen35." + k, v)
    wr.add_float32("qwen35.rope.freq_base", 1e7)
    wr.add_array("qwen35.rope.dimension_sections", [11, 11, 10, 0])
    wr.add_array("tokenizer.ggml.tokens", ["a", "b", "c"])
    wr.add_tensor("token_embd.weight", np.zeros((1024, 256), np.float32))
    for i in (0, 1, 2):
        wr.add_tensor(f"blk.{i}.ssm_a", np.zeros(4, np.float32))
    wr.add_tensor("blk.3.attn_q.weight", np.zeros((2 * 4 * 256, 256), np.float32))
    wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
    c = crane_b200.gguf_config(path)
    assert c["model_type"] == "qwen3_5_text" and c["vocab_size"] == 1024 and c["tie_word_embeddings"] is True
    assert c["layer_types"] == ["linear_attention"] * 3 + ["full_attention"]
    assert (c["linear_num_key_heads"], c["linear_num_value_heads"], c["linear_key_head_dim"], c["linear_value_head_dim"]) == (2, 4, 128, 128)
    assert c["head_dim"] == 256 and abs(c["partial_rotary_factor"] - 0.25) < 1e-9
    assert c["rope_parameters"]["mrope_section"] == [11, 11, 10, 0] and c["rope_parameters"]["rope_theta"] == 1e7
