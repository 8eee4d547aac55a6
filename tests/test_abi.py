"""The C-ABI library loads on a CPU-only machine, exports every symbol include/*.h declares, and refuses to work
without a CUDA device (no CPU fallback).  Also checks the generated constants and the synthetic workload classes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import c_oracle as co
from oracle import jjs_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"\b(jjs_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    entry.build_cuda()
    from jubjub_schnorr_b200 import _native
    lib = _native.lib()
    declared = _declared()
    assert len(declared) >= 15
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_native.EXPORTS) == declared


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from jubjub_schnorr_b200 import BatchVerifier, JjsError
    with pytest.raises(JjsError, match="no CUDA device"):
        BatchVerifier([0])


def test_failed_init_leaves_an_inert_context():
    """jjs_init without a usable device hands back a context for jjs_last_error only: every entry point must answer
    JJS_ERR_CUDA on it instead of touching device state that was never built (round-1 advisor finding: SIGFPE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    entry.build_cuda()
    from jubjub_schnorr_b200 import _native
    lib = _native.lib()
    ctx = C.c_void_p()
    assert lib.jjs_init(None, 1, C.byref(ctx)) == -2 and ctx
    assert b"no CUDA device" in lib.jjs_last_error(ctx)
    buf = (C.c_uint8 * 4096)()
    off = (C.c_uint32 * 5)(0, 1, 2, 3, 4)
    n = 4
    calls = [
        lambda: lib.jjs_verify_single(ctx, buf, buf, buf, n, buf, None),
        lambda: lib.jjs_verify_double(ctx, buf, buf, buf, n, buf, None),
        lambda: lib.jjs_verify_vargen(ctx, buf, buf, buf, n, buf, buf),
        lambda: lib.jjs_verify_aggregate(ctx, buf, off, buf, buf, n, buf, None, None),
        lambda: lib.jjs_verify_batch(ctx, buf, buf, buf, n, buf),
        lambda: lib.jjs_verify_batch_double(ctx, buf, buf, buf, n, buf),
        lambda: lib.jjs_verify_batch_vargen(ctx, buf, buf, buf, n, buf),
        lambda: lib.jjs_verify_batch_aggregate(ctx, buf, off, buf, buf, n, buf),
        lambda: lib.jjs_verify_mixed(ctx, (_native.Part * 1)(_native.Part(0, C.addressof(buf), None, C.addressof(buf), C.addressof(buf), n,
                                                                          C.addressof(buf), None, None, None)), 1),
        lambda: lib.jjs_verify_single_device(ctx, 0, buf, buf, buf, n, buf, None, None),
        lambda: lib.jjs_verify_double_device(ctx, 0, buf, buf, buf, n, buf, None, None),
        lambda: lib.jjs_verify_vargen_device(ctx, 0, buf, buf, buf, n, buf, None, None),
        lambda: lib.jjs_verify_aggregate_device(ctx, 0, buf, off, off, buf, buf, n, buf, None, None, None),
        lambda: lib.jjs_status_bitmap_device(ctx, 0, buf, n, buf, None),
        lambda: lib.jjs_verify_ext(ctx, 0, buf, buf, buf, n, buf, None),
        lambda: lib.jjs_points_to_ext(ctx, buf, buf, n, buf),
        lambda: lib.jjs_challenge_only(ctx, 0, buf, buf, buf, n, buf),
        lambda: lib.jjs_subgroup_check(ctx, buf, n, 0, buf),
        lambda: lib.jjs_fb_table_check(ctx, 0, buf, n, buf),
        lambda: lib.jjs_sign_batch(ctx, 0, buf, buf, None, buf, n, buf, buf),
        lambda: lib.jjs_sign_aggregate_batch(ctx, buf, off, buf, buf, n, buf, buf),
        lambda: lib.jjs_multisig_combine(ctx, buf, buf, buf, buf, off, buf, n, None, buf, None, None),
    ]
    for i, call in enumerate(calls):
        assert call() == -2, i
    assert b"no CUDA device" in lib.jjs_last_error(ctx)      # the reason survives the refused calls
    assert lib.jjs_device_count(ctx) == 0
    lib.jjs_destroy(ctx)


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "jubjub_schnorr_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "workload.py" and "No oracle" in text, f


def test_generated_constants_match_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_device_constants.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.round_constants() == o.ROUND_CONSTANTS and gen.mds_matrix() == o.MDS
    assert gen.safe_tag(5) == o.safe_tag(5) and gen.safe_tag(10) == o.safe_tag(10)
    sch = gen.hades_schedule()
    for st in ([1, 2, 3, 4, 5], [o.Q - 1, 0, 5, 1 << 250, 77]):
        assert gen.hades_scaled_model(st, sch) == o.hades_permute(st)
    hdr = open(os.path.join(ROOT, "jubjub_schnorr_b200", "csrc", "jjs_constants_uniform.h")).read()
    first = "0x%08xu" % (sch["first_ark"][0] & 0xFFFFFFFF)
    assert first in hdr, "jjs_constants_uniform.h is stale: run tools/gen_device_constants.py"


@pytest.mark.parametrize("variant,gen,ver", [(0, co.gen_single, co.verify_single), (1, co.gen_double, co.verify_double),
                                             (2, co.gen_vargen, co.verify_vargen)])
def test_workload_classes_have_the_stated_status(variant, gen, ver):
    from jubjub_schnorr_b200 import workload as wl
    pk, sig, msg = gen(1, 640)
    pk, sig, msg, exp, cls = wl.invalidate(variant, pk, sig, msg, 0.25, seed=0x5A0CE)
    st, _ = ver(pk, sig, msg)
    assert np.array_equal(st, exp)
    assert set(np.unique(cls).tolist()) == set(range(-1, len(wl.CLASSES)))


def test_mixed_order_constants_of_the_workload():
    """workload.MIXED_ORDER: on the curve, not of small order, not in the prime-order subgroup; the torsion part has order 2, 4 or 8."""
    from jubjub_schnorr_b200 import workload as wl
    orders = set()
    for enc in wl.MIXED_ORDER:
        p = o.point_from_bytes(enc.tobytes())
        assert p is not None and not o.point_is_valid(p)
        t = o.pmul(p, o.R_ORDER)
        assert t != o.IDENTITY and o.pmul(p, 8) != o.IDENTITY
        orders.add(next(k for k in (2, 4, 8) if o.pmul(t, k) == o.IDENTITY))
    assert orders == {2, 4, 8}


def test_base58_codec_matches_reference_strings():
    """The host mirror's base58 (src/serde_support.rs text form) against the pinned strings of tests/serde.rs (CPU only)."""
    import json
    from jubjub_schnorr_b200 import api
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))["serde_kat"]
    for key, size in (("serde_public_key", 32), ("serde_signature", 64), ("serde_public_key_double", 64), ("serde_signature_double", 96),
                      ("serde_secret_key", 32), ("serde_signature_var_gen", 64)):
        raw = api.b58decode(s[key])
        assert len(raw) == size and api.b58encode(raw) == s[key] and raw == o.b58decode(s[key], size)
    assert api.b58encode(b"\x00\x00\x01") == "112" and api.b58decode("112") == b"\x00\x00\x01"


def test_header_is_plain_c(tmp_path):
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-c", os.path.join(ROOT, "tests", "cpp", "abi_c_check.c"),
                           "-o", str(tmp_path / "abi_c_check.o")])


def _c_prototypes():
    """{name: [parameter type strings]} of every jjs_* function declared in include/jjschnorr_b200.h"""
    src = open(os.path.join(ROOT, "include", "jjschnorr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"\b(int|void|const char\*|uint64_t)\s+(jjs_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", src):
        out[name] = [" ".join(p.split()) for p in params.split(",")]
    return out


def test_rust_shim_matches_the_header():
    """rust/ holds the reference-side binding (build.rs, src/ffi.rs, src/gpu.rs, patches/).  It cannot be compiled here (no Rust
    toolchain), so it is held against the C header instead: EVERY exported function is declared in ffi.rs with the same number of
    parameters, pointer-ness, constness and integer widths; jjs_part has the header's fields in the header's order; and gpu.rs wraps
    every host-buffer verify entry point, is brace-balanced and keeps the reference's scalar signatures."""
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    ffi = re.sub(r"//[^\n]*", "", ffi)
    protos = _c_prototypes()
    rust = dict(re.findall(r"pub fn (jjs_[a-z_0-9]+)\s*\(([^)]*)\)", ffi))
    assert sorted(rust) == sorted(protos), (sorted(set(protos) - set(rust)), sorted(set(rust) - set(protos)))
    widths = {"size_t": "usize", "int": "c_int", "uint64_t": "u64"}
    for name, params in rust.items():
        rparams = [p.strip() for p in params.split(",") if p.strip()]
        cparams = protos[name]
        assert len(rparams) == len(cparams), (name, rparams, cparams)
        for rp, cp in zip(rparams, cparams):
            rtype = rp.split(":", 1)[1].strip()
            is_ptr = "*" in cp
            assert rtype.startswith("*") == is_ptr, (name, rp, cp)
            if is_ptr and "**" not in cp:
                assert rtype.startswith("*const") == cp.startswith("const"), (name, rp, cp)
                pointee = {"uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "int": "c_int", "double": "f64", "void": "c_void", "jjs_ctx": "jjs_ctx",
                           "jjs_part": "jjs_part", "char": "c_char"}[cp.replace("const", "").replace("*", "").split()[0]]
                assert rtype.split()[-1] == pointee, (name, rp, cp)
            if not is_ptr:
                assert widths[cp.split()[0]] == rtype, (name, rp, cp)
    # struct jjs_part: same fields, same order, same pointer-ness
    hdr = open(os.path.join(ROOT, "include", "jjschnorr_b200.h")).read()
    c_fields = re.findall(r"^\s+(const\s+)?(\w+)(\*?)\s+(\w+);", re.search(r"typedef struct jjs_part \{(.*?)\} jjs_part;", hdr, flags=re.S).group(1), flags=re.M)
    r_fields = re.findall(r"pub (\w+): ([^,]+),", re.search(r"pub struct jjs_part \{(.*?)\}", ffi, flags=re.S).group(1))
    assert [f[3] for f in c_fields] == [f[0] for f in r_fields]
    for (const, ctype, star, _), (_, rtype) in zip(c_fields, r_fields):
        assert bool(star) == rtype.strip().startswith("*") and (not star or bool(const) == rtype.strip().startswith("*const")), (ctype, rtype)
    # the wrappers
    gpu = open(os.path.join(ROOT, "rust", "src", "gpu.rs")).read()
    code = re.sub(r"//[^\n]*", "", gpu)
    for a_, b_ in ("{}", "()", "[]"):
        assert code.count(a_) == code.count(b_), (a_, code.count(a_), code.count(b_))
    for fn in ("jjs_init", "jjs_destroy", "jjs_last_error", "jjs_verify_single", "jjs_verify_double", "jjs_verify_vargen", "jjs_verify_aggregate", "jjs_verify_batch",
               "jjs_verify_batch_double", "jjs_verify_batch_vargen", "jjs_verify_batch_aggregate", "jjs_verify_ext", "jjs_verify_mixed"):
        assert "ffi::" + fn + "(" in code, fn
    for sig_ in ("pub fn verify_single(pk: &PublicKey, sig: &Signature, message: BlsScalar) -> Result<(), Error>",
                 "pub fn verify_double(pk: &PublicKeyDouble, sig: &SignatureDouble, message: BlsScalar) -> Result<(), Error>",
                 "pub fn verify_var_gen(pk: &PublicKeyVarGen, sig: &SignatureVarGen, message: BlsScalar) -> Result<(), Error>",
                 "pub fn verify_batch(items: &[(PublicKey, Signature, BlsScalar)]) -> Vec<bool>"):
        assert sig_ in gpu, sig_
    patch = open(os.path.join(ROOT, "rust", "patches", "verify_methods.patch")).read()
    for call in ("crate::gpu::verify_single(self, sig, message)", "crate::gpu::verify_double(self, sig_double, message)",
                 "crate::gpu::verify_var_gen(self, sig_var_gen, message)"):
        assert call in patch, call
    # the status codes are the header's
    hdr = open(os.path.join(ROOT, "include", "jjschnorr_b200.h")).read()
    for cname in ("JJS_OK", "JJS_INVALID_SIGNATURE", "JJS_INVALID_POINT", "JJS_BYTES_ERROR", "JJS_INVALID_MULTISIG_TRANSCRIPT", "JJS_INVALID_MULTISIG_SHARE"):
        c_val = int(re.search(r"#define\s+%s\s+(\d+)" % cname, hdr).group(1))
        r_val = int(re.search(r"pub const %s: u8 = (\d+);" % cname, ffi).group(1))
        assert c_val == r_val, cname
