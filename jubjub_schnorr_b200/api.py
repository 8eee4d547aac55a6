"""Host-side mirror of the reference's verify-path types (same names, argument meaning and error behaviour), on top of
the C ABI.  Values are held as the reference's wire encodings; every check runs on the GPU.

  reference                                   here
  PublicKey::from_bytes / to_bytes / verify   PublicKey.from_bytes / to_bytes / verify      (src/keys/public.rs:80-135)
  PublicKeyDouble, SignatureDouble            same names                                    (src/keys/public/double.rs, src/signatures/double.rs)
  PublicKeyVarGen, SignatureVarGen            same names                                    (src/keys/public/var_gen.rs, src/signatures/var_gen.rs)
  multisig::aggregate_pk                      multisig_aggregate_pk                         (src/multisig.rs:154-156)
  Error::{InvalidSignature, InvalidPoint, BytesError}   Error enum                          (src/error.rs:13-26)
  serde (base58 strings of to_bytes())         to_base58 / from_base58                       (src/serde_support.rs)
  (new) verify_batch(&[(PublicKey, Signature, BlsScalar)]) -> Vec<bool>   verify_batch

`verify` returns None for Ok(()) and an Error member for Err(..), so tests read like the reference's:
    assert pk.verify(sig, msg) is None
    assert wrong_pk.verify(sig, msg) == Error.InvalidSignature
"""
from __future__ import annotations

import enum
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .batch import DOUBLE, SINGLE, VARGEN, BatchVerifier

_default: Optional[BatchVerifier] = None


def default_verifier() -> BatchVerifier:
    global _default
    if _default is None:
        _default = BatchVerifier([0])
    return _default


def set_default_verifier(bv: Optional[BatchVerifier]) -> None:
    global _default
    _default = bv


class Error(enum.IntEnum):
    InvalidSignature = 1
    InvalidPoint = 2
    BytesError = 3


class BytesError(ValueError):
    """dusk_bytes::Error::InvalidData raised by from_bytes of a malformed encoding."""


def _status_to_result(st: int) -> Optional[Error]:
    return None if st == 0 else Error(int(st))


def _msg_bytes(message) -> bytes:
    """BlsScalar: an int (canonical value) or its 32-byte little-endian encoding."""
    if isinstance(message, int):
        return message.to_bytes(32, "little")
    b = bytes(message)
    if len(b) != 32:
        raise BytesError("a BlsScalar is 32 bytes")
    return b


def _check_points(encodings: Sequence[bytes]) -> None:
    """JubJubAffine::from_bytes on every 32-byte encoding (decode only, no subgroup check)."""
    arr = np.frombuffer(b"".join(encodings), dtype=np.uint8)
    res = default_verifier().subgroup_check(arr, method=0)
    if (res == 0xFF).any():
        raise BytesError("InvalidData: not a canonical JubJub point encoding")


def _check_scalar(u: bytes) -> None:
    if int.from_bytes(u, "little") >= 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7:
        raise BytesError("InvalidData: scalar is not canonical")


_B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def b58encode(b: bytes) -> str:
    n, s = int.from_bytes(b, "big"), ""
    while n:
        n, rem = divmod(n, 58)
        s = _B58[rem] + s
    return "1" * (len(b) - len(b.lstrip(b"\0"))) + s


def b58decode(s: str) -> bytes:
    n = 0
    for ch in s:
        if ch not in _B58:
            raise BytesError("invalid base58 character")
        n = n * 58 + _B58.index(ch)
    pad = len(s) - len(s.lstrip("1"))
    return b"\0" * pad + (n.to_bytes((n.bit_length() + 7) // 8, "big") if n else b"")


class _Wire:
    SIZE = 0
    _points: Tuple[Tuple[int, int], ...] = ()
    _scalar = False

    def __init__(self, raw: bytes):
        self._raw = bytes(raw)

    @classmethod
    def from_bytes(cls, b):
        b = bytes(b)
        if len(b) != cls.SIZE:
            raise BytesError(f"{cls.__name__} is {cls.SIZE} bytes")
        if cls._scalar:
            _check_scalar(b[:32])
        _check_points([b[lo:hi] for lo, hi in cls._points])
        return cls(b)

    @classmethod
    def from_raw_unchecked(cls, b):
        """Keep any bytes (the reference's from_raw_unchecked keeps any point); verify() reports what is wrong."""
        return cls(b)

    def to_bytes(self) -> bytes:
        return self._raw

    # serde text form of the reference: a base58 (Bitcoin alphabet) string of to_bytes()  (src/serde_support.rs)
    def to_base58(self) -> str:
        return b58encode(self._raw)

    @classmethod
    def from_base58(cls, s: str):
        b = b58decode(s)
        if len(b) != cls.SIZE:
            raise BytesError(f"{cls.__name__}: expected {cls.SIZE} bytes, got {len(b)}")
        return cls.from_bytes(b)

    def __eq__(self, other):
        return type(self) is type(other) and self._raw == other._raw

    def __hash__(self):
        return hash((type(self).__name__, self._raw))

    def __repr__(self):
        return f"{type(self).__name__}({self._raw.hex()})"


class Signature(_Wire):
    SIZE, _points, _scalar = 64, ((32, 64),), True

    def u(self) -> bytes:
        return self._raw[:32]

    def R(self) -> bytes:
        return self._raw[32:]


class SignatureDouble(_Wire):
    SIZE, _points, _scalar = 96, ((32, 64), (64, 96)), True

    def u(self) -> bytes:
        return self._raw[:32]

    def R(self) -> bytes:
        return self._raw[32:64]

    def R_prime(self) -> bytes:
        return self._raw[64:]


class SignatureVarGen(Signature):
    pass


class _Key(_Wire):
    _variant = SINGLE

    def is_valid(self) -> bool:
        """is_torsion_free && is_on_curve && !is_identity for every point of the key (src/keys/public.rs:159-164)."""
        parts = [self._raw[lo:hi] for lo, hi in self._points]
        res = default_verifier().subgroup_check(np.frombuffer(b"".join(parts), dtype=np.uint8), method=0)
        identity = (1).to_bytes(32, "little")
        return all(r == 1 and p != identity for r, p in zip(res.tolist(), parts))

    def verify(self, sig, message) -> Optional[Error]:
        bv = default_verifier()
        fn = {SINGLE: bv.verify_single, DOUBLE: bv.verify_double, VARGEN: bv.verify_vargen}[self._variant]
        st = fn(np.frombuffer(self._raw, dtype=np.uint8), np.frombuffer(sig.to_bytes(), dtype=np.uint8),
                np.frombuffer(_msg_bytes(message), dtype=np.uint8))
        return _status_to_result(int(st[0]))


class PublicKey(_Key):
    SIZE, _points, _variant = 32, ((0, 32),), SINGLE


class PublicKeyDouble(_Key):
    SIZE, _points, _variant = 64, ((0, 32), (32, 64)), DOUBLE

    def pk(self) -> bytes:
        return self._raw[:32]

    def pk_prime(self) -> bytes:
        return self._raw[32:]


class PublicKeyVarGen(_Key):
    SIZE, _points, _variant = 64, ((0, 32), (32, 64)), VARGEN

    def public_key(self) -> bytes:
        return self._raw[:32]

    def generator(self) -> bytes:
        return self._raw[32:]


def verify_batch(items: Iterable[Tuple[PublicKey, Signature, object]]) -> List[bool]:
    """verify_batch(&[(PublicKey, Signature, BlsScalar)]) -> Vec<bool>: one GPU pass, true iff PublicKey::verify is Ok."""
    items = list(items)
    if not items:
        return []
    pk = np.frombuffer(b"".join(k.to_bytes() for k, _, _ in items), dtype=np.uint8)
    sig = np.frombuffer(b"".join(s.to_bytes() for _, s, _ in items), dtype=np.uint8)
    msg = np.frombuffer(b"".join(_msg_bytes(m) for _, _, m in items), dtype=np.uint8)
    bv = default_verifier()
    return bv.unpack_bitmap(bv.verify_batch(pk, sig, msg), len(items)).tolist()


def multisig_aggregate_pk(pk_vec: Sequence[PublicKey]) -> PublicKey:
    """multisig::aggregate_pk: sum of d_i * pk_i with d_i = H(pk_i || pk_1 .. pk_n); inputs are not validated."""
    pks = np.frombuffer(b"".join(k.to_bytes() for k in pk_vec), dtype=np.uint8)
    dummy_sig = np.zeros(64, dtype=np.uint8)
    dummy_msg = np.zeros(32, dtype=np.uint8)
    _, agg = default_verifier().verify_aggregate(pks, [0, len(pk_vec)], dummy_sig, dummy_msg, want_aggregate_key=True)
    return PublicKey.from_raw_unchecked(agg.tobytes())
