"""ctypes binding of oracle/libjjs_oracle.so (the C restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never from jubjub_schnorr_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libjjs_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "jjs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s", "libjjs_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        u8p, u32p = C.c_void_p, C.c_void_p
        _lib.jjo_verify_single.argtypes = [u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_int]
        _lib.jjo_verify_double.argtypes = [u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_int]
        _lib.jjo_verify_vargen.argtypes = [u8p, u8p, u8p, C.c_size_t, u8p, u8p, C.c_int]
        _lib.jjo_verify_aggregate.argtypes = [u8p, u32p, u8p, u8p, C.c_size_t, u8p, u8p, u8p, C.c_int]
        for name in ("jjo_gen_single", "jjo_gen_double", "jjo_gen_vargen"):
            getattr(_lib, name).argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, u8p, u8p, u8p, C.c_int]
        _lib.jjo_gen_aggregate.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, u32p, u8p, u8p, u8p, C.c_int]
        for name in ("jjo_verify_single", "jjo_verify_double", "jjo_verify_vargen", "jjo_verify_aggregate",
                     "jjo_gen_single", "jjo_gen_double", "jjo_gen_vargen", "jjo_gen_aggregate",
                     "jjo_hades_permute", "jjo_fq_mul"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(-1, width)


def default_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _verify(fn, pk, pkw, sig, sigw, msg, threads):
    pk, sig, msg = _u8(pk, pkw), _u8(sig, sigw), _u8(msg, 32)
    n = msg.shape[0]
    assert pk.shape[0] == n and sig.shape[0] == n
    status = np.zeros(n, dtype=np.uint8)
    c = np.zeros((n, 32), dtype=np.uint8)
    fn(_p(pk), _p(sig), _p(msg), n, _p(status), _p(c), threads or default_threads())
    return status, c


def verify_single(pk, sig, msg, threads=0):
    return _verify(lib().jjo_verify_single, pk, 32, sig, 64, msg, threads)


def verify_double(pk, sig, msg, threads=0):
    return _verify(lib().jjo_verify_double, pk, 64, sig, 96, msg, threads)


def verify_vargen(pk, sig, msg, threads=0):
    return _verify(lib().jjo_verify_vargen, pk, 64, sig, 64, msg, threads)


def verify_aggregate(pks, offsets, sig, msg, threads=0):
    pks, sig, msg = _u8(pks, 32), _u8(sig, 64), _u8(msg, 32)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    n = msg.shape[0]
    assert offsets.shape[0] == n + 1 and sig.shape[0] == n and int(offsets[-1]) == pks.shape[0]
    status = np.zeros(n, dtype=np.uint8)
    c = np.zeros((n, 32), dtype=np.uint8)
    agg = np.zeros((n, 32), dtype=np.uint8)
    lib().jjo_verify_aggregate(_p(pks), _p(offsets), _p(sig), _p(msg), n, _p(status), _p(c), _p(agg),
                               threads or default_threads())
    return status, c, agg


def _gen(fn, pkw, sigw, seed, first, n, threads):
    pk = np.zeros((n, pkw), dtype=np.uint8)
    sig = np.zeros((n, sigw), dtype=np.uint8)
    msg = np.zeros((n, 32), dtype=np.uint8)
    fn(seed, first, n, _p(pk), _p(sig), _p(msg), threads or default_threads())
    return pk, sig, msg


def gen_single(seed, n, first=0, threads=0):
    return _gen(lib().jjo_gen_single, 32, 64, seed, first, n, threads)


def gen_double(seed, n, first=0, threads=0):
    return _gen(lib().jjo_gen_double, 64, 96, seed, first, n, threads)


def gen_vargen(seed, n, first=0, threads=0):
    return _gen(lib().jjo_gen_vargen, 64, 64, seed, first, n, threads)


def gen_aggregate(seed, signers, first=0, threads=0):
    """signers: array of per-item signer counts.  Returns (pks[total,32], offsets[n+1], sig, msg)."""
    signers = np.asarray(signers, dtype=np.uint32)
    n = signers.shape[0]
    offsets = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(signers, out=offsets[1:])
    pks = np.zeros((int(offsets[-1]), 32), dtype=np.uint8)
    sig = np.zeros((n, 64), dtype=np.uint8)
    msg = np.zeros((n, 32), dtype=np.uint8)
    lib().jjo_gen_aggregate(seed, first, n, _p(offsets), _p(pks), _p(sig), _p(msg), threads or default_threads())
    return pks, offsets, sig, msg


def multisig_combine(pks, Rs, Ss, zs, offsets, msg, threads=0):
    """multisig::combine over ragged sessions.  Returns (status[n], bad_index[n], sig[n,64], share_ok[K])."""
    pks, Rs, Ss, zs, msg = _u8(pks, 32), _u8(Rs, 32), _u8(Ss, 32), _u8(zs, 32), _u8(msg, 32)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    n, K = msg.shape[0], pks.shape[0]
    assert offsets.shape[0] == n + 1 and int(offsets[-1]) == K == Rs.shape[0] == Ss.shape[0] == zs.shape[0]
    status = np.zeros(n, dtype=np.uint8)
    bad = np.zeros(n, dtype=np.uint32)
    sig = np.zeros((n, 64), dtype=np.uint8)
    ok = np.zeros(K, dtype=np.uint8)
    f = lib().jjo_multisig_combine
    f.argtypes = [C.c_void_p] * 6 + [C.c_size_t] + [C.c_void_p] * 4 + [C.c_int]
    f.restype = None
    f(_p(pks), _p(Rs), _p(Ss), _p(zs), _p(offsets), _p(msg), n, _p(ok), _p(status), _p(bad), _p(sig), threads or default_threads())
    return status, bad, sig, ok


def gen_multisig(seed, signers, first=0, threads=0):
    """Valid SpeedyMuSig sessions.  Returns (pks, Rs, Ss, zs [K,32 each], offsets[n+1], msg[n,32])."""
    signers = np.asarray(signers, dtype=np.uint32)
    n = signers.shape[0]
    offsets = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(signers, out=offsets[1:])
    K = int(offsets[-1])
    pks, Rs, Ss, zs = (np.zeros((K, 32), dtype=np.uint8) for _ in range(4))
    msg = np.zeros((n, 32), dtype=np.uint8)
    f = lib().jjo_gen_multisig
    f.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t] + [C.c_void_p] * 6 + [C.c_int]
    f.restype = None
    f(seed, first, n, _p(offsets), _p(pks), _p(Rs), _p(Ss), _p(zs), _p(msg), threads or default_threads())
    return pks, Rs, Ss, zs, offsets, msg


def verify_ext(variant, pts160, u32, msg32):
    """Typed inputs (JubJubExtended Montgomery coordinates, 160 bytes per point, item-major); single-threaded."""
    slots = {0: 2, 1: 4, 2: 3}[variant]
    pts, u, msg = _u8(pts160, 160 * slots), _u8(u32, 32), _u8(msg32, 32)
    n = msg.shape[0]
    status = np.zeros(n, dtype=np.uint8)
    c = np.zeros((n, 32), dtype=np.uint8)
    f = lib().jjo_verify_ext
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    f.restype = None
    f(variant, _p(pts), _p(u), _p(msg), n, _p(status), _p(c))
    return status, c


def point_to_ext(p32, z_mont: int):
    """Compressed point -> 160-byte JubJubExtended coordinates (u z, v z, z, u z, v) with Montgomery scale z_mont."""
    out = (C.c_uint8 * 160)()
    ok = lib().jjo_point_to_ext(_b32(p32), _b32(z_mont.to_bytes(32, "little")), out)
    return bytes(out) if ok else None


def _b32(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b))


def point_decode(p32):
    out = (C.c_uint8 * 64)()
    if not lib().jjo_point_decode(_b32(p32), out):
        return None
    raw = bytes(out)
    return int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:], "little")


def point_from_uv(u: int, v: int):
    out = (C.c_uint8 * 32)()
    ok = lib().jjo_point_from_uv(_b32(u.to_bytes(32, "little") + v.to_bytes(32, "little")), out)
    return bytes(out) if ok else None


def point_add(a32, b32):
    out = (C.c_uint8 * 32)()
    return bytes(out) if lib().jjo_point_add(_b32(a32), _b32(b32), out) else None


def point_mul(p32, k: int):
    out = (C.c_uint8 * 32)()
    return bytes(out) if lib().jjo_point_mul(_b32(p32), _b32(k.to_bytes(32, "little")), out) else None


def point_is_valid(p32) -> int:
    return lib().jjo_point_is_valid(_b32(p32))


def hades_permute(state):
    buf = (C.c_uint8 * 160).from_buffer_copy(b"".join(int(x).to_bytes(32, "little") for x in state))
    lib().jjo_hades_permute(buf)
    raw = bytes(buf)
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(5)]


def poseidon_hash(inputs, truncated=True):
    out = (C.c_uint8 * 32)()
    data = b"".join(int(x).to_bytes(32, "little") for x in inputs)
    ok = lib().jjo_poseidon_hash(_b32(data), len(inputs), int(truncated), out)
    return int.from_bytes(bytes(out), "little") if ok else None


def fq_mul(a: int, b: int) -> int:
    out = (C.c_uint8 * 32)()
    lib().jjo_fq_mul(_b32(a.to_bytes(32, "little")), _b32(b.to_bytes(32, "little")), out)
    return int.from_bytes(bytes(out), "little")
