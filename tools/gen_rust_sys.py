#!/usr/bin/env python
"""Generate rust/crane-b200-sys/src/lib.rs (raw `extern "C"` declarations) from include/crane_b200.h, so the binding a Crane
maintainer links cannot drift from the header.  No Rust toolchain exists in this image: the output is
checked for completeness by tests/test_abi.py (every exported symbol declared once, regenerating gives the committed file).
    python tools/gen_rust_sys.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "crane_b200.h")
OUT = os.path.join(ROOT, "rust", "crane-b200-sys", "src", "lib.rs")

BASE = {"int": "c_int", "void": "c_void", "char": "c_char", "float": "f32", "double": "f64", "size_t": "usize", "uint8_t": "u8",
        "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int64_t": "i64", "unsigned char": "u8",
        "crane_b200_model": "crane_b200_model", "crane_b200_logits": "crane_b200_logits", "crane_b200_sampling": "crane_b200_sampling"}


def rust_type(c: str) -> str:
    c = c.strip()
    const = False
    ptr = c.count("*")
    c = c.replace("*", " ").strip()
    if c.startswith("const "):
        const, c = True, c[6:].strip()
    c = c.replace(" const", "").strip()
    base = BASE[c]
    if ptr == 0:
        return base
    t = base
    for i in range(ptr):
        t = ("*const " if (const and i == 0) else "*mut ") + t
    return t


def parse_functions(text: str):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for m in re.finditer(r"CRANE_B200_API\s+([^;(]+?)\s*\b(crane_b200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.+?)(\w+)$", a)
                params.append((mm.group(2), rust_type(mm.group(1))))
        yield name, rust_type(ret) if ret != "void" else None, params


def parse_struct(text: str, name: str):
    m = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", text, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    for line in body.split(";"):
        line = " ".join(line.split())
        if not line:
            continue
        mm = re.match(r"(.+?)(\w+)$", line)
        yield mm.group(2), rust_type(mm.group(1))


def generate() -> str:
    text = open(HDR).read()
    out = ["//! Raw FFI for libcrane_b200.so -- GENERATED from include/crane_b200.h by tools/gen_rust_sys.py; do not edit.",
           "//! The safe wrapper implementing crane-serve's `ModelBackend` on top of this is sketched in INTEGRATION.md section 2.",
           "#![allow(non_camel_case_types)]", "", "use std::os::raw::{c_char, c_int, c_void};", "",
           "#[repr(C)]", "pub struct crane_b200_model {", "    _private: [u8; 0],", "}", ""]
    for en in ("crane_b200_status", "crane_b200_dtype"):
        m = re.search(r"typedef enum \{([^}]*)\}\s*" + en + r"\s*;", text, flags=re.S)
        for item in re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S).split(","):
            item = " ".join(item.split())
            if item:
                k, v = [x.strip() for x in item.split("=")]
                out.append(f"pub const {k}: c_int = {v};")
        out.append("")
    for st in ("crane_b200_logits", "crane_b200_sampling"):
        out += ["#[repr(C)]", "#[derive(Clone, Copy)]", f"pub struct {st} {{"]
        out += [f"    pub {n}: {t}," for n, t in parse_struct(text, st)]
        out += ["}", ""]
    out.append('#[link(name = "crane_b200")]')
    out.append('extern "C" {')
    for name, ret, params in parse_functions(text):
        ps = ", ".join(f"{'r#' + n if n in ('type', 'ref', 'in', 'box', 'move') else n}: {t}" for n, t in params)
        out.append(f"    pub fn {name}({ps}){' -> ' + ret if ret else ''};")
    out.append("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    src = generate()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == src else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(src)
    print(f"wrote {OUT}: {src.count('pub fn ')} functions")
