#!/usr/bin/env python3
"""Writes the Rust shim of INTEGRATION.md out as files under rust/ (build.rs, src/ffi.rs, src/gpu.rs), so that a
maintainer can copy them into the reference crate and so that tests/test_abi.py can check the extern "C" declarations
against include/jjschnorr_b200.h.  The markdown stays the single source; nothing here is compiled in this image (no
Rust toolchain)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = "// Extracted from INTEGRATION.md by tools/extract_rust_shim.py -- edit the markdown, not this file.\n// NOT compiled in this repository's environment (no cargo/rustc in the image).\n\n"


def blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```rust\n(.*?)```", text, flags=re.S)


def files():
    b = blocks()
    assert len(b) >= 4, "INTEGRATION.md lost a rust block"
    return {"build.rs": HEADER + b[0], os.path.join("src", "ffi.rs"): HEADER + b[1],
            os.path.join("src", "gpu.rs"): HEADER + b[2] + "\n// ---- typed inputs (jjs_verify_ext) ----\n" + b[3]}


def main():
    check = "--check" in sys.argv
    ok = True
    for rel, content in files().items():
        path = os.path.join(ROOT, "rust", rel)
        if check:
            ok = ok and os.path.exists(path) and open(path).read() == content
        else:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            open(path, "w").write(content)
    if check and not ok:
        sys.exit("rust/ is out of date: run tools/extract_rust_shim.py")


if __name__ == "__main__":
    main()
