"""Hostile inputs to the two host-side parsers that read files (round-1 advisor findings: unchecked offsets in the GGUF reader,
over-reads in the JSON string scanner).  Both are pure host code, so they are exercised here without a GPU:
  * `crane_b200/csrc/json_min.h` compiled alone with g++ -fsanitize=address,undefined and driven with mutated documents
    (tests/cpp/fuzz_json.cpp): every input parses or throws, nothing reads past the terminator;
  * `crane_b200_gguf_config` (metadata -> config.json text) on mutated GGUF files, in a child process so that a crash is a test
    failure and not the end of the test run: every call returns a config or an error status."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_json_parser_survives_mutated_documents(tmp_path):
    exe = str(tmp_path / "fuzz_json")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                    os.path.join(ROOT, "tests", "cpp", "fuzz_json.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "30000"], capture_output=True, text=True, timeout=300)
    print(r.stdout[-300:], r.stderr[-2000:])
    if r.returncode != 0 and "json fuzz ok" not in r.stdout and ("Shadow memory" in r.stderr or "ReserveShadowMemoryRange" in r.stderr):
        pytest.skip("AddressSanitizer cannot run here: " + r.stderr.strip().splitlines()[0][:200])     # sandbox address-space layout, not a finding
    assert r.returncode == 0 and "json fuzz ok" in r.stdout


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import crane_b200
data = bytearray(open(sys.argv[2], "rb").read())
rng = np.random.default_rng(int(sys.argv[4]))
ok = bad = 0
for it in range(int(sys.argv[3])):
    d = bytearray(data)
    kind = it % 4
    if kind == 0:                                   # flip a few bytes in the header / metadata / tensor-info region
        for _ in range(1 + int(rng.integers(4))):
            d[int(rng.integers(min(len(d), 2048)))] = int(rng.integers(256))
    elif kind == 1:                                 # truncate anywhere
        d = d[: int(rng.integers(len(d)))]
    elif kind == 2:                                 # a huge little-endian count / length / offset somewhere in the first 2 KB
        p = int(rng.integers(min(len(d), 2048) - 8))
        d[p:p + 8] = [2**63 - 1, 2**64 - 1, 2**32, 2**40 + 7][int(rng.integers(4))].to_bytes(8, "little")
    else:                                           # both
        d = d[: max(16, int(rng.integers(len(d))))]
        d[int(rng.integers(len(d)))] = int(rng.integers(256))
    path = sys.argv[2] + ".mut"
    open(path, "wb").write(bytes(d))
    try:
        crane_b200.gguf_config(path)
        ok += 1
    except crane_b200.CraneB200Error:
        bad += 1
print(f"gguf fuzz ok: {ok} accepted, {bad} rejected")
"""


def _write_gguf(path, arch):
    import gguf
    wr = gguf.GGUFWriter(path, arch)
    keys = [("block_count", 4), ("embedding_length", 256), ("feed_forward_length", 512), ("attention.head_count", 4), ("attention.head_count_kv", 2)]
    if arch == "qwen35":                            # the hybrid's metadata: layer_types is one entry per block_count, sections are an array
        keys += [("attention.key_length", 256), ("ssm.conv_kernel", 4), ("ssm.state_size", 128), ("ssm.group_count", 2), ("ssm.time_step_rank", 4),
                 ("ssm.inner_size", 512), ("full_attention_interval", 4)]
    for k, v in keys:
        wr.add_uint32(f"{arch}.{k}", v)
    wr.add_float32(f"{arch}.rope.freq_base", 10000.0)
    wr.add_float32(f"{arch}.attention.layer_norm_rms_epsilon", 1e-5)
    if arch == "qwen35":
        wr.add_array("qwen35.rope.dimension_sections", [11, 11, 10, 0])
    wr.add_array("tokenizer.ggml.tokens", ["a", "b", "c"])
    wr.add_tensor("token_embd.weight", np.zeros((64, 256), np.float32))
    wr.add_tensor("blk.0.attn_q_norm.weight", np.ones(128, np.float32))
    if arch == "qwen35":
        wr.add_tensor("blk.0.ssm_a", np.ones(4, np.float32))
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()


@pytest.mark.parametrize("arch", ["qwen3", "qwen35"])
def test_gguf_config_survives_mutated_files(tmp_path, arch):
    """Found by this test in round 2 and fixed: a mutated `block_count` drove a 4-billion-entry `layer_types` loop (hang) or overflowed
    the fixed 12 KB text buffer (truncated JSON returned as success); an architecture name with a quote in it produced invalid JSON;
    64-bit counts were cast to signed integers unchecked."""
    path = str(tmp_path / "meta.gguf")
    _write_gguf(path, arch)
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, path, "1500", "7"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-300:], r.stderr[-2000:])
    assert r.returncode == 0 and "gguf fuzz ok" in r.stdout, "the GGUF metadata reader crashed, hung or returned invalid JSON on a mutated file"
