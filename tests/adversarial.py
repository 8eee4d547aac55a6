"""Adversarial batch builder for the parity tests (SURVEY.md section 8(c) classes).

Test infrastructure: uses the oracle for point arithmetic.  Every class states the status the
reference semantics give (0 Ok, 1 InvalidSignature, 2 InvalidPoint, 3 BytesError); the tests still
compare the CUDA path against the oracle item by item, the expected code is a second check.
"""
from __future__ import annotations

import numpy as np

from oracle import c_oracle as co
from oracle import jjs_oracle as o

Q_BYTES = np.frombuffer(o.le32(o.Q), dtype=np.uint8)
IDENTITY_ENC = o.point_to_bytes(o.IDENTITY)
ORDER2_ENC = o.point_to_bytes((0, o.Q - 1))


def torsion_points():
    """Encodings of points of exact order 2, 4 and 8 (T = [r] * random curve point)."""
    out = {}
    v = 2
    while len(out) < 3:
        v += 1
        enc = bytearray(o.le32(v))
        p = co.point_decode(bytes(enc))
        if p is None:
            continue
        t = co.point_mul(bytes(enc), o.R_ORDER)
        order = 1
        cur = t
        while cur != IDENTITY_ENC:
            cur = co.point_add(cur, t)
            order += 1
            assert order <= 8
        if order == 8:
            t8 = t
            out[8] = t8
            out[4] = co.point_add(t8, t8)
            out[2] = co.point_add(out[4], out[4])
    assert out[2] == ORDER2_ENC
    return out


_TORSION = None


def torsion():
    global _TORSION
    if _TORSION is None:
        _TORSION = torsion_points()
    return _TORSION


def off_curve_encoding(rng) -> bytes:
    while True:
        v = int.from_bytes(rng.bytes(32), "little") % o.Q
        b = o.le32(v)
        if co.point_decode(b) is None:
            return b


def noncanonical_v_encoding(rng) -> bytes:
    """v in [q, 2^255): rejected by JubJubAffine::from_bytes."""
    v = o.Q + int(rng.integers(0, 1 << 62))
    assert v < (1 << 255)
    b = bytearray(o.le32(v))
    b[31] |= int(rng.integers(0, 2)) << 7
    return bytes(b)


# (name, expected status, which field it touches)
POINT_CLASSES = [
    ("off_curve", 3), ("v_ge_q", 3), ("identity", 2), ("order2", 2), ("plus_t2", 2), ("plus_t4", 2),
    ("plus_t8", 2), ("zip216_identity", 3), ("zip216_order2", 3),
]


def tamper_point(enc: bytes, cls: str, rng) -> bytes:
    t = torsion()
    if cls == "off_curve":
        return off_curve_encoding(rng)
    if cls == "v_ge_q":
        return noncanonical_v_encoding(rng)
    if cls == "identity":
        return IDENTITY_ENC
    if cls == "order2":
        return ORDER2_ENC
    if cls in ("plus_t2", "plus_t4", "plus_t8"):
        return co.point_add(enc, t[int(cls[-1])])
    if cls == "zip216_identity":
        b = bytearray(IDENTITY_ENC); b[31] |= 0x80; return bytes(b)
    if cls == "zip216_order2":
        b = bytearray(ORDER2_ENC); b[31] |= 0x80; return bytes(b)
    raise ValueError(cls)


def make_adversarial(kind: str, pk, sig, msg, seed: int, frac: float = 0.5):
    """Tamper a fraction of a valid batch.  kind in {"single", "double", "vargen"}.

    Returns (pk, sig, msg, expected_status, class_names).  Point fields: single pk[0:32], sig[32:64];
    double pk[0:32], pk[32:64], sig[32:64], sig[64:96]; vargen pk[0:32] (key), pk[32:64] (generator),
    sig[32:64].
    """
    rng = np.random.default_rng(seed)
    pk, sig, msg = pk.copy(), sig.copy(), msg.copy()
    n = msg.shape[0]
    expected = np.zeros(n, dtype=np.uint8)
    names = ["valid"] * n
    pk_fields = {"single": [0], "double": [0, 32], "vargen": [0, 32]}[kind]
    sig_fields = {"single": [32], "double": [32, 64], "vargen": [32]}[kind]
    classes = []
    for cname, st in POINT_CLASSES:
        for f in pk_fields:
            classes.append(("pk%d_%s" % (f, cname), st, ("pk", f, cname)))
        for f in sig_fields:
            classes.append(("sig%d_%s" % (f, cname), st, ("sig", f, cname)))
    classes += [
        ("u_tampered", 1, None), ("u_ge_r", 3, None), ("u_all_ff", 3, None), ("m_tampered", 1, None), ("m_ge_q", 3, None),
        ("R_other_point", 1, None), ("R_sign_flip", 1, None), ("pk_other_key", 1, None),
    ]
    idx = rng.permutation(n)[: int(n * frac)]
    for j, i in enumerate(idx):
        name, st, spec = classes[j % len(classes)]
        if spec is not None:
            arr = pk if spec[0] == "pk" else sig
            f = spec[1]
            arr[i, f:f + 32] = np.frombuffer(tamper_point(arr[i, f:f + 32].tobytes(), spec[2], rng), dtype=np.uint8)
        elif name == "u_tampered":
            u = (int.from_bytes(sig[i, :32].tobytes(), "little") + 1 + int(rng.integers(0, 1 << 60))) % o.R_ORDER
            sig[i, :32] = np.frombuffer(o.le32(u), dtype=np.uint8)
        elif name == "u_ge_r":
            u = o.R_ORDER + int(rng.integers(0, 1 << 60))
            sig[i, :32] = np.frombuffer(o.le32(u), dtype=np.uint8)
        elif name == "u_all_ff":
            sig[i, :32] = 0xFF
        elif name == "m_tampered":
            m = (int.from_bytes(msg[i].tobytes(), "little") + 1) % o.Q
            msg[i] = np.frombuffer(o.le32(m), dtype=np.uint8)
        elif name == "m_ge_q":
            msg[i] = np.frombuffer(o.le32(o.Q + int(rng.integers(0, 1 << 60))), dtype=np.uint8)
        elif name == "R_other_point":
            other = (i + 1) % n
            sig[i, 32:64] = sig[other, 32:64] if names[other] == "valid" else pk[other, :32]
            if not co.point_is_valid(sig[i, 32:64].tobytes()) == 1:
                sig[i, 32:64] = np.frombuffer(o.point_to_bytes(o.G), dtype=np.uint8)
        elif name == "R_sign_flip":
            sig[i, 63] ^= 0x80
        elif name == "pk_other_key":
            pk[i, :32] = np.frombuffer(o.point_to_bytes(o.pmul(o.G, 1 + int(rng.integers(1, 1 << 30)))), dtype=np.uint8)
        expected[i] = st
        names[i] = name
    return pk, sig, msg, expected, names


def torsion_shifted_signatures(kind: str, n: int, seed: int):
    """Signatures made by a signer who knows the secret key but shifts the commitment by a small-order point:
    R = r*B + T with T of order 2, 4 or 8, c = H(R, ...), u = r - c*sk.  Then u*B + c*PK == R - T, so the equation
    fails only by the torsion part; the reference rejects them in Signature::is_valid (InvalidPoint, status 2).
    These are the inputs a verifier that skips the subgroup test of R would get wrong: any multiple rho of the
    equation with rho * T == O holds.  Item i uses T of order 2, 4, 8 in turn; for the double variant R and R' are
    shifted alternately (i // 3 even: R, odd: R').  Returns (pk, sig, msg, expected_status)."""
    rng = np.random.default_rng(seed)
    t = torsion()
    G_ENC, GN_ENC = o.point_to_bytes(o.G), o.point_to_bytes(o.G_NUMS)
    pkw = 32 if kind == "single" else 64
    sigw = 96 if kind == "double" else 64
    pk = np.zeros((n, pkw), np.uint8)
    sig = np.zeros((n, sigw), np.uint8)
    msg = np.zeros((n, 32), np.uint8)
    for i in range(n):
        sk = 1 + int.from_bytes(rng.bytes(32), "little") % (o.R_ORDER - 1)
        r = 1 + int.from_bytes(rng.bytes(32), "little") % (o.R_ORDER - 1)
        m = int.from_bytes(rng.bytes(64), "little") % o.Q
        T = t[(2, 4, 8)[i % 3]]
        if kind == "single":
            pk_e = co.point_mul(G_ENC, sk)
            R_e = co.point_add(co.point_mul(G_ENC, r), T)
            c = co.poseidon_hash(list(co.point_decode(R_e)) + list(co.point_decode(pk_e)) + [m])
            pts = pk_e, R_e
        elif kind == "double":
            pk_e = co.point_mul(G_ENC, sk) + co.point_mul(GN_ENC, sk)
            R0, R1 = co.point_mul(G_ENC, r), co.point_mul(GN_ENC, r)
            if (i // 3) % 2 == 0:
                R0 = co.point_add(R0, T)
            else:
                R1 = co.point_add(R1, T)
            R_e = R0 + R1
            c = co.poseidon_hash([o.DOUBLE_CHALLENGE_DOMAIN] + list(co.point_decode(R0)) + list(co.point_decode(R1))
                                 + list(co.point_decode(pk_e[:32])) + list(co.point_decode(pk_e[32:])) + [m])
        else:
            gen_e = co.point_mul(G_ENC, 1 + int.from_bytes(rng.bytes(32), "little") % (o.R_ORDER - 1))
            key_e = co.point_mul(gen_e, sk)
            pk_e = key_e + gen_e
            R_e = co.point_add(co.point_mul(gen_e, r), T)
            c = co.poseidon_hash(list(co.point_decode(R_e)) + list(co.point_decode(key_e)) + list(co.point_decode(gen_e)) + [m])
        u = (r - c * sk) % o.R_ORDER
        pk[i] = np.frombuffer(pk_e, np.uint8)
        sig[i] = np.frombuffer(o.le32(u) + R_e, np.uint8)
        msg[i] = np.frombuffer(o.le32(m), np.uint8)
    return pk, sig, msg, np.full(n, 2, np.uint8)


def bitflip_fuzz(pk, sig, msg, seed: int, frac: float = 0.75, max_flips: int = 3):
    """Differential-fuzz input: `frac` of the items get 1..max_flips random single-bit flips anywhere in their key,
    signature or message bytes (so any field can become non-canonical, leave the curve, change sign, leave the subgroup
    or simply stop matching).  No expectation is attached: the GPU is compared with the oracle item by item."""
    rng = np.random.default_rng(seed)
    pk, sig, msg = pk.copy(), sig.copy(), msg.copy()
    n = msg.shape[0]
    widths = (pk.shape[1], sig.shape[1], 32)
    total = sum(widths) * 8
    for i in np.nonzero(rng.random(n) < frac)[0]:
        for _ in range(int(rng.integers(1, max_flips + 1))):
            # bias half of the flips towards the top byte of a 32-byte field, where the range and sign rules live
            if rng.random() < 0.5:
                field = int(rng.integers(0, sum(widths) // 32))
                bit = field * 256 + 248 + int(rng.integers(0, 8))
            else:
                bit = int(rng.integers(0, total))
            byte, b = divmod(bit, 8)
            arr, off = (pk, 0) if byte < widths[0] else ((sig, widths[0]) if byte < widths[0] + widths[1] else (msg, widths[0] + widths[1]))
            arr[i, byte - off] ^= np.uint8(1 << b)
    return pk, sig, msg
