"""Big-integer CPU oracle for the jubjub-schnorr verify path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain Python integers, the algorithm the reference
(dusk-network/jubjub-schnorr 0.7.0-rc.0) runs for signature verification, plus the sign-side
functions needed to regenerate the reference's pinned vectors.  Nothing in the product path
(`jubjub_schnorr_b200/`) may import it; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may.

Parity status: PINNED.  The arithmetic itself lives in crates that are not vendored in the
reference checkout (dusk-poseidon 0.42.0-rc.0 + dusk-safe, dusk-jubjub 0.15, dusk-bls12_381 0.14,
dusk-bytes 0.1; reference Cargo.toml:21-30), so their published algorithms are restated here and the
restatement is pinned against every known-answer vector the reference carries for this path
(tests/test_oracle_kat.py): src/multisig.rs:544-735, tests/serde.rs:34-142,
tests/common/mod.rs:23-66 + tests/schnorr_double.rs:72-82.

Each function cites the reference call site (file:line under /root/reference) it follows.
"""
from __future__ import annotations

import hashlib
import struct

# --------------------------------------------------------------------------------------------
# Fields (dusk-bls12_381 BlsScalar = Fq, dusk-jubjub JubJubScalar = Fr)
# --------------------------------------------------------------------------------------------
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # BLS12-381 scalar field
R_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7  # JubJub subgroup order
MONT_R = (1 << 256) % Q
EDWARDS_D = (-10240 * pow(10241, -1, Q)) % Q

STATUS_OK = 0                 # Ok(())
STATUS_INVALID_SIGNATURE = 1  # Error::InvalidSignature   (src/error.rs:17)
STATUS_INVALID_POINT = 2      # Error::InvalidPoint       (src/error.rs:19)
STATUS_BYTES_ERROR = 3        # Error::BytesError(..)     (src/error.rs:15) -- from_bytes failed


def fq_from_le(b: bytes):
    """BlsScalar::from_bytes: canonical 32-byte LE, None if >= q."""
    x = int.from_bytes(b, "little")
    return x if x < Q else None


def fr_from_le(b: bytes):
    """JubJubScalar::from_slice/from_bytes (src/signatures.rs:113): canonical, None if >= r."""
    x = int.from_bytes(b, "little")
    return x if x < R_ORDER else None


def le32(x: int) -> bytes:
    return x.to_bytes(32, "little")


def fq_sqrt(a: int):
    """Any square root of a mod q or None (Tonelli-Shanks, q-1 = 2^32 * t)."""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 32, (Q - 1) >> 32
    z = 7  # a generator-ish non residue; verified below
    while pow(z, (Q - 1) // 2, Q) == 1:
        z += 1
    c = pow(z, t, Q)
    x = pow(a, (t + 1) // 2, Q)
    b = pow(a, t, Q)
    m = s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % Q
            i += 1
        g = pow(c, 1 << (m - i - 1), Q)
        x = x * g % Q
        c = g * g % Q
        b = b * c % Q
        m = i
    return x


# --------------------------------------------------------------------------------------------
# JubJub (dusk-jubjub): -u^2 + v^2 = 1 + d u^2 v^2, affine big-int arithmetic
# --------------------------------------------------------------------------------------------
IDENTITY = (0, 1)
G = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 0x12)  # GENERATOR_EXTENDED
G_NUMS = (  # GENERATOR_NUMS_EXTENDED, pinned through tests/serde.rs:75-86
    0x5E67B8F316F414F7BD9514C773FD4456931E316A39FE4541921710179DF76377,
    0x43D80EB3B2F3EB1B7B162DBEEB3B34FD9949BA0F82A5507A6705B707162E3EF8,
)


def on_curve(p) -> bool:
    u, v = p
    return (v * v - u * u - 1 - EDWARDS_D * u * u % Q * v * v) % Q == 0


def padd(p, q_):
    """Complete twisted-Edwards addition (a = -1)."""
    u1, v1 = p
    u2, v2 = q_
    k = EDWARDS_D * u1 * u2 % Q * v1 * v2 % Q
    u3 = (u1 * v2 + v1 * u2) * pow(1 + k, -1, Q) % Q
    v3 = (v1 * v2 + u1 * u2) * pow(1 - k, -1, Q) % Q
    return (u3, v3)


def pneg(p):
    return ((-p[0]) % Q, p[1])


def pmul(p, k: int):
    """Scalar multiplication; group-level result is all parity needs."""
    acc = IDENTITY
    for bit in bin(k)[2:] if k else "":
        acc = padd(acc, acc)
        if bit == "1":
            acc = padd(acc, p)
    return acc


def is_torsion_free(p) -> bool:
    """dusk_jubjub::JubJubExtended::is_torsion_free: [r]P == identity."""
    return pmul(p, R_ORDER) == IDENTITY


def point_is_valid(p) -> bool:
    """PublicKey::is_valid / Signature::is_valid (src/keys/public.rs:159-164,
    src/signatures.rs:93-98): torsion free, on curve, not the identity."""
    return is_torsion_free(p) and on_curve(p) and p != IDENTITY


def point_to_bytes(p) -> bytes:
    """JubJubAffine::to_bytes: v little-endian, bit 255 = parity of u."""
    u, v = p
    b = bytearray(le32(v))
    b[31] |= (u & 1) << 7
    return bytes(b)


def point_from_bytes(b: bytes):
    """JubJubAffine::from_bytes / from_slice (src/keys/public.rs:88, src/signatures.rs:114).

    Returns the affine point or None.  Rejects v >= q, non-square u^2, and the u == 0 encoding
    that carries a sign bit (ZIP-216 rule).  No subgroup check here.
    """
    assert len(b) == 32
    sign = b[31] >> 7
    v = int.from_bytes(b, "little") & ((1 << 255) - 1)
    if v >= Q:
        return None
    v2 = v * v % Q
    u2 = (v2 - 1) * pow(1 + EDWARDS_D * v2, -1, Q) % Q
    u = fq_sqrt(u2)
    if u is None:
        return None
    if (u & 1) != sign:
        u = (-u) % Q
    if u == 0 and sign == 1:
        return None
    return (u, v)


# --------------------------------------------------------------------------------------------
# Poseidon (dusk-poseidon 0.42 + dusk-safe): Hades width 5, x^5, 4 + 60 + 4 rounds
# --------------------------------------------------------------------------------------------
WIDTH = 5
FULL_ROUNDS = 8
PARTIAL_ROUNDS = 60
N_ROUND_CONSTANTS = (FULL_ROUNDS + PARTIAL_ROUNDS) * WIDTH  # 340 consumed of a stream of 960


def _gen_round_constants():
    out = []
    b = b"poseidon-for-plonk"
    c = 1
    for _ in range(N_ROUND_CONSTANTS):
        b = hashlib.sha512(b).digest()
        c = (int.from_bytes(b, "little") % Q + c) % Q
        # the upstream asset file stores Montgomery limbs that are loaded as raw integers,
        # which multiplies every constant by R = 2^256 mod q
        out.append(c * MONT_R % Q)
    return out


def _gen_mds():
    return [[MONT_R * pow(i + j + 5, -1, Q) % Q for j in range(WIDTH)] for i in range(WIDTH)]


ROUND_CONSTANTS = _gen_round_constants()
MDS = _gen_mds()


def hades_permute(state):
    """dusk_poseidon Hades permutation over 5 Fq lanes."""
    s = list(state)
    rc = iter(ROUND_CONSTANTS)

    def mds(s):
        return [sum(MDS[i][k] * s[k] for k in range(WIDTH)) % Q for i in range(WIDTH)]

    for rnd in range(FULL_ROUNDS + PARTIAL_ROUNDS):
        s = [(x + next(rc)) % Q for x in s]
        if rnd < FULL_ROUNDS // 2 or rnd >= FULL_ROUNDS // 2 + PARTIAL_ROUNDS:
            s = [pow(x, 5, Q) for x in s]
        else:
            s[WIDTH - 1] = pow(s[WIDTH - 1], 5, Q)
        s = mds(s)
    return s


def safe_tag(n_absorb: int, n_squeeze: int = 1, domain: int = 0) -> int:
    """dusk-safe tag for IO pattern [Absorb(n), Squeeze(k)], Domain::Other."""
    data = struct.pack(">II", 0x80000000 | n_absorb, n_squeeze) + struct.pack(">Q", domain)
    return int.from_bytes(hashlib.blake2b(data, digest_size=64).digest(), "little") % Q


def poseidon_hash(inputs) -> int:
    """Hash::digest(Domain::Other, inputs)[0]: rate 4, capacity lane 0 holds the tag."""
    state = [safe_tag(len(inputs)), 0, 0, 0, 0]
    pos = 0
    for x in inputs:
        if pos == 4:
            state = hades_permute(state)
            pos = 0
        state[1 + pos] = (state[1 + pos] + x) % Q
        pos += 1
    state = hades_permute(state)
    return state[1]


def poseidon_hash_truncated(inputs) -> int:
    """Hash::digest_truncated(Domain::Other, inputs)[0] (src/signatures.rs:130): low 250 bits."""
    return poseidon_hash(inputs) & ((1 << 250) - 1)


# --------------------------------------------------------------------------------------------
# Challenge transcripts
# --------------------------------------------------------------------------------------------
DOUBLE_CHALLENGE_DOMAIN = 0x4A4A53434844424C  # "JJSCHDBL", src/signatures/double.rs:24-25


def challenge_single(R, pk, m: int) -> int:
    """signatures::challenge_hash (src/signatures.rs:122-141)."""
    return poseidon_hash_truncated([R[0], R[1], pk[0], pk[1], m])


def challenge_double(R, Rp, pk, pkp, m: int) -> int:
    """signatures::double::challenge_hash (src/signatures/double.rs:151-177)."""
    return poseidon_hash_truncated(
        [DOUBLE_CHALLENGE_DOMAIN, R[0], R[1], Rp[0], Rp[1], pk[0], pk[1], pkp[0], pkp[1], m])


def challenge_var_gen(R, pk, gen, m: int) -> int:
    """signatures::var_gen::challenge_hash (src/signatures/var_gen.rs:121-142)."""
    return poseidon_hash_truncated([R[0], R[1], pk[0], pk[1], gen[0], gen[1], m])


def delinearization_coeff(pk_i, pks) -> int:
    """multisig::delinearization_coeff (src/multisig.rs:393-409)."""
    pre = [pk_i[0], pk_i[1]]
    for p in pks:
        pre += [p[0], p[1]]
    return poseidon_hash_truncated(pre)


def aggregate_pk(pks):
    """multisig::aggregate_pk / aggregate_key (src/multisig.rs:154-156, 416-429).
    No validation of the individual keys, exactly like the reference."""
    acc = IDENTITY
    for p in pks:
        acc = padd(acc, pmul(p, delinearization_coeff(p, pks)))
    return acc


# --------------------------------------------------------------------------------------------
# verify: typed level (points already decoded) and wire level (bytes -> status)
# --------------------------------------------------------------------------------------------
def verify_single_points(pk, u: int, R, m: int):
    """PublicKey::verify (src/keys/public.rs:114-135).  Returns (status, c or None)."""
    if not point_is_valid(pk) or not point_is_valid(R):
        return STATUS_INVALID_POINT, None
    c = challenge_single(R, pk, m)
    lhs = padd(pmul(G, u), pmul(pk, c))
    return (STATUS_OK if lhs == R else STATUS_INVALID_SIGNATURE), c


def verify_double_points(pk, pkp, u: int, R, Rp, m: int):
    """PublicKeyDouble::verify (src/keys/public/double.rs:86-117)."""
    if not (point_is_valid(pk) and point_is_valid(pkp)) or not (point_is_valid(R) and point_is_valid(Rp)):
        return STATUS_INVALID_POINT, None
    c = challenge_double(R, Rp, pk, pkp, m)
    p1 = padd(pmul(G, u), pmul(pk, c))
    p2 = padd(pmul(G_NUMS, u), pmul(pkp, c))
    return (STATUS_OK if (p1 == R and p2 == Rp) else STATUS_INVALID_SIGNATURE), c


def verify_var_gen_points(pk, gen, u: int, R, m: int):
    """PublicKeyVarGen::verify (src/keys/public/var_gen.rs:107-133)."""
    if not (point_is_valid(pk) and point_is_valid(gen)) or not point_is_valid(R):
        return STATUS_INVALID_POINT, None
    c = challenge_var_gen(R, pk, gen, m)
    lhs = padd(pmul(gen, u), pmul(pk, c))
    return (STATUS_OK if lhs == R else STATUS_INVALID_SIGNATURE), c


def _msg(msg32: bytes):
    return fq_from_le(msg32)


def verify_single(pk32: bytes, sig64: bytes, msg32: bytes):
    """PublicKey::from_bytes + Signature::from_bytes + BlsScalar::from_bytes, then verify.
    Returns (status, c_bytes or None).  Status 3 if any decode fails."""
    pk = point_from_bytes(pk32)
    u = fr_from_le(sig64[:32])
    R = point_from_bytes(sig64[32:])
    m = _msg(msg32)
    if pk is None or u is None or R is None or m is None:
        return STATUS_BYTES_ERROR, None
    st, c = verify_single_points(pk, u, R, m)
    return st, (le32(c) if c is not None else None)


def verify_double(pk64: bytes, sig96: bytes, msg32: bytes):
    pk = point_from_bytes(pk64[:32])
    pkp = point_from_bytes(pk64[32:])
    u = fr_from_le(sig96[:32])
    R = point_from_bytes(sig96[32:64])
    Rp = point_from_bytes(sig96[64:])
    m = _msg(msg32)
    if None in (pk, pkp, u, R, Rp, m):
        return STATUS_BYTES_ERROR, None
    st, c = verify_double_points(pk, pkp, u, R, Rp, m)
    return st, (le32(c) if c is not None else None)


def verify_var_gen(pk64: bytes, sig64: bytes, msg32: bytes):
    pk = point_from_bytes(pk64[:32])
    gen = point_from_bytes(pk64[32:])
    u = fr_from_le(sig64[:32])
    R = point_from_bytes(sig64[32:])
    m = _msg(msg32)
    if None in (pk, gen, u, R, m):
        return STATUS_BYTES_ERROR, None
    st, c = verify_var_gen_points(pk, gen, u, R, m)
    return st, (le32(c) if c is not None else None)


def verify_aggregate(pks32, sig64: bytes, msg32: bytes):
    """aggregate_pk(&pks).verify(&sig, m): returns (status, c_bytes, aggregate_pk_bytes)."""
    pks = [point_from_bytes(b) for b in pks32]
    u = fr_from_le(sig64[:32])
    R = point_from_bytes(sig64[32:])
    m = _msg(msg32)
    if None in pks or u is None or R is None or m is None:
        return STATUS_BYTES_ERROR, None, None
    agg = aggregate_pk(pks)
    st, c = verify_single_points(agg, u, R, m)
    return st, (le32(c) if c is not None else None), point_to_bytes(agg)


# --------------------------------------------------------------------------------------------
# sign side (only to regenerate the reference's pinned vectors and to make test inputs)
# --------------------------------------------------------------------------------------------
def sign_single(sk: int, rnd: int, m: int):
    """SecretKey::sign (src/keys/secret.rs:174-194) with hedged_nonce (src/nonce.rs:32-44);
    `rnd` is the JubJubScalar the RNG produced."""
    r = poseidon_hash_truncated([rnd, sk, 1, m])
    R = pmul(G, r)
    c = challenge_single(R, pmul(G, sk), m)
    return (r - c * sk) % R_ORDER, R


def sign_double(sk: int, rnd: int, m: int):
    """SecretKey::sign_double (src/keys/secret/double.rs:57-85), nonce tag 2 (src/nonce.rs:49-61)."""
    r = poseidon_hash_truncated([rnd, sk, 2, m])
    R, Rp = pmul(G, r), pmul(G_NUMS, r)
    c = challenge_double(R, Rp, pmul(G, sk), pmul(G_NUMS, sk), m)
    return (r - c * sk) % R_ORDER, R, Rp


def sign_var_gen(sk: int, gen, rnd: int, m: int):
    """SecretKeyVarGen::sign (src/keys/secret/var_gen.rs) with hedged_nonce_var_gen (src/nonce.rs:69-86)."""
    r = poseidon_hash_truncated([rnd, sk, gen[0], gen[1], m])
    R = pmul(gen, r)
    c = challenge_var_gen(R, pmul(gen, sk), gen, m)
    return (r - c * sk) % R_ORDER, R


# --------------------------------------------------------------------------------------------
# rand 0.8 StdRng (ChaCha12) and base58, needed only for the tests/serde.rs vectors
# --------------------------------------------------------------------------------------------
class StdRng:
    """rand::rngs::StdRng::seed_from_u64: PCG32-expanded seed into ChaCha12."""

    def __init__(self, seed_u64: int):
        mul, inc, mask = 6364136223846793005, 11634580027462260723, (1 << 64) - 1
        state, seed = seed_u64, b""
        for _ in range(8):
            state = (state * mul + inc) & mask
            xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            seed += struct.pack("<I", ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf = b""

    def _block(self):
        def rotl(x, n):
            return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF

        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *self.key,
                self.counter & 0xFFFFFFFF, self.counter >> 32, 0, 0]
        x = list(init)

        def qr(a, b, c, d):
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
            x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
            x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)

        for _ in range(6):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.counter += 1
        return struct.pack("<16I", *[(a + b) & 0xFFFFFFFF for a, b in zip(x, init)])

    def fill_bytes(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += self._block()
        out, self.buf = self.buf[:n], self.buf[n:]
        return out

    def random_fr(self) -> int:
        """JubJubScalar::random: 64 bytes, wide-reduced."""
        return int.from_bytes(self.fill_bytes(64), "little") % R_ORDER

    def random_fq(self) -> int:
        """BlsScalar::random: 64 bytes, wide-reduced."""
        return int.from_bytes(self.fill_bytes(64), "little") % Q


_B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def b58encode(b: bytes) -> str:
    n = int.from_bytes(b, "big")
    s = ""
    while n:
        n, rem = divmod(n, 58)
        s = _B58[rem] + s
    return "1" * (len(b) - len(b.lstrip(b"\0"))) + s


def b58decode(s: str, length: int) -> bytes:
    n = 0
    for ch in s:
        n = n * 58 + _B58.index(ch)
    return n.to_bytes(length, "big")


# --------------------------------------------------------------------------------------------
# multisig: share verification and combine
# --------------------------------------------------------------------------------------------
STATUS_INVALID_MULTISIG_TRANSCRIPT = 4  # Error::InvalidMultisigTranscript  (src/error.rs)
STATUS_INVALID_MULTISIG_SHARE = 5       # Error::InvalidMultisigShare(index)


def multisig_common(pks, Rs, Ss, m):
    """multisig::multisig_common (src/multisig.rs:440-500): (d_i list, aggregate key, a, RSa, c)."""
    ds = [delinearization_coeff(p, pks) for p in pks]
    agg = IDENTITY
    for p, d in zip(pks, ds):
        agg = padd(agg, pmul(p, d))
    pre = [agg[0], agg[1], m]
    for R, S in zip(Rs, Ss):
        pre += [R[0], R[1], S[0], S[1]]
    a = poseidon_hash_truncated(pre)
    RSa = IDENTITY
    for R, S in zip(Rs, Ss):
        RSa = padd(padd(RSa, R), pmul(S, a))
    c = challenge_single(RSa, agg, m)
    return ds, agg, a, RSa, c


def multisig_share_ok(z, i, pks, Rs, Ss, coeffs) -> bool:
    """verify_share_with_coefficients (src/multisig.rs:366-387): z_i G + (c d_i) pk_i == R_i + a S_i."""
    ds, _, a, _, c = coeffs
    lhs = padd(pmul(G, z), pmul(pks[i], c * ds[i] % R_ORDER))
    return lhs == padd(Rs[i], pmul(Ss[i], a))


def multisig_combine(zs32, pks32, Rs32, Ss32, msg32):
    """multisig::combine (src/multisig.rs:311-347) on wire encodings.
    Returns (status, first_bad_index or None, signature bytes or None, [share_ok])."""
    n = len(pks32)
    if n == 0 or not (len(zs32) == len(Rs32) == len(Ss32) == n):
        return STATUS_INVALID_MULTISIG_TRANSCRIPT, None, None, []
    zs = [fr_from_le(z) for z in zs32]
    pks, Rs, Ss = ([point_from_bytes(b) for b in arr] for arr in (pks32, Rs32, Ss32))
    m = fq_from_le(msg32)
    if None in zs or None in pks or None in Rs or None in Ss or m is None:
        return STATUS_BYTES_ERROR, None, None, []
    coeffs = multisig_common(pks, Rs, Ss, m)
    oks = [multisig_share_ok(zs[i], i, pks, Rs, Ss, coeffs) for i in range(n)]
    if not all(oks):
        return STATUS_INVALID_MULTISIG_SHARE, oks.index(False), None, oks
    return STATUS_OK, None, le32(sum(zs) % R_ORDER) + point_to_bytes(coeffs[3]), oks
