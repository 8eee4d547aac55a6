"""Builds typed (JubJubExtended-coordinate) batches out of wire batches for the jjs_verify_ext parity tests."""
import numpy as np

from oracle import c_oracle as co
from oracle import jjs_oracle as o

SLOT_FIELDS = {0: [("pk", 0), ("sig", 32)], 1: [("pk", 0), ("pk", 32), ("sig", 32), ("sig", 64)], 2: [("pk", 0), ("pk", 32), ("sig", 32)]}


def to_typed(variant, pk, sig, msg, seed):
    """Returns (pts[n, slots*160], u[n,32], keep) -- items whose point fields do not decode cannot be typed and are dropped."""
    rng = np.random.default_rng(seed)
    n = msg.shape[0]
    fields = SLOT_FIELDS[variant]
    pts = np.zeros((n, 160 * len(fields)), dtype=np.uint8)
    keep = np.ones(n, dtype=bool)
    for i in range(n):
        for s, (which, off) in enumerate(fields):
            src = pk if which == "pk" else sig
            z = int.from_bytes(rng.bytes(32), "little") % (o.Q - 1) + 1
            e = co.point_to_ext(src[i, off:off + 32].tobytes(), z)
            if e is None:
                keep[i] = False
                break
            pts[i, 160 * s:160 * (s + 1)] = np.frombuffer(e, dtype=np.uint8)
    return pts, sig[:, :32].copy(), keep


def corrupt_typed(pts, status, seed):
    """Typed-only failure modes on a copy: z = 0, inconsistent t1, off-curve u, unreduced coordinate.  Returns
    (pts, expected) where expected is None for untouched items."""
    rng = np.random.default_rng(seed)
    pts = pts.copy()
    n = pts.shape[0]
    expected = [None] * n
    slots = pts.shape[1] // 160
    kinds = ["z_zero", "t1_bad", "u_bad", "unreduced", "identity_projective"]
    for j, i in enumerate(rng.permutation(n)[: n // 3]):
        kind = kinds[j % len(kinds)]
        s = int(rng.integers(0, slots))
        base = 160 * s
        if kind == "z_zero":
            pts[i, base + 64: base + 96] = 0
            expected[i] = 2
        elif kind == "t1_bad":
            pts[i, base + 96] ^= 1
            expected[i] = 2
        elif kind == "u_bad":
            pts[i, base] ^= 1
            expected[i] = 2
        elif kind == "unreduced":
            pts[i, base + 32: base + 64] = 0xFF
            expected[i] = 3
        else:  # (0, z, z, 0, z): the identity in projective form
            z = pts[i, base + 64: base + 96].copy()
            pts[i, base: base + 32] = 0
            pts[i, base + 32: base + 64] = z
            pts[i, base + 96: base + 128] = 0
            pts[i, base + 128: base + 160] = z
            expected[i] = 2 if status[i] != 3 else 3
    return pts, expected
