"""Pin the C oracle (oracle/jjs_oracle.c): reference KATs directly, then item-by-item agreement with
the big-integer oracle on valid and adversarial batches (CPU only)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import jjs_oracle as o
from tests import adversarial as adv
from tests.test_oracle_kat import legacy_double_fixture

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))


def _a(b):
    return np.frombuffer(bytes(b), dtype=np.uint8)


def test_c_oracle_multisig_kat():
    k = KAT["multisig_kat"]
    pks = b"".join(bytes.fromhex(x) for x in k["PUBLIC_KEYS"])
    sig, m = bytes.fromhex(k["SIGNATURE"]), o.le32(31)
    st, c = co.verify_single(_a(bytes.fromhex(k["AGGREGATE_PUBLIC_KEY"])), _a(sig), _a(m))
    assert st[0] == 0 and c.tobytes().hex() == k["CHALLENGE"]
    st, c, agg = co.verify_aggregate(_a(pks), [0, 3], _a(sig), _a(m))
    assert st[0] == 0 and c.tobytes().hex() == k["CHALLENGE"] and agg.tobytes().hex() == k["AGGREGATE_PUBLIC_KEY"]
    for sk, enc in zip(k["_inputs"]["sk"], k["PUBLIC_KEYS"]):
        assert co.point_mul(o.point_to_bytes(o.G), sk).hex() == enc
    pts = [o.point_from_bytes(bytes.fromhex(x)) for x in k["PUBLIC_KEYS"]]
    for p, d in zip(pts, k["DELINEARIZATION"]):
        pre = [p[0], p[1]] + [c_ for q in pts for c_ in q]
        assert o.le32(co.poseidon_hash(pre)).hex() == d


def test_c_oracle_serde_signatures():
    s = KAT["serde_kat"]
    m = bytes.fromhex("6dfe107145b1cba63d5f5ed0c410c09441fbc0d70c9bfea970949499aa128214")
    st, c = co.verify_single(_a(o.b58decode(s["serde_public_key"], 32)), _a(o.b58decode(s["serde_signature"], 64)), _a(m))
    assert st[0] == 0 and c.tobytes().hex() == "7ad531e479fe4f1d2c1858c180e18f57e549d7eda93c85f46344f4716de67e02"
    st, c = co.verify_double(_a(o.b58decode(s["serde_public_key_double"], 64)),
                             _a(o.b58decode(s["serde_signature_double"], 96)), _a(m))
    assert st[0] == 0 and c.tobytes().hex() == "b706ff0423cbdff51e73ee23985e66a5f827343f03ecf4e9ecd7f32fc9791901"
    # var-gen: message is the third draw after sk and the generator scalar
    rng = o.StdRng(s["_seed"]); rng.random_fr(); rng.random_fr()
    mv = o.le32(rng.random_fq())
    st, c = co.verify_vargen(_a(o.b58decode(s["serde_public_key_var_gen"], 64)),
                             _a(o.b58decode(s["serde_signature_var_gen"], 64)), _a(mv))
    assert st[0] == 0 and c.tobytes().hex() == "648cf37f901b93870bec5cb3d7339934efd8307d7c667a3718cd14643a677603"


def test_c_oracle_adaptive_secondary_key_rejected():
    pkb, sig, mb = legacy_double_fixture()
    st, _ = co.verify_double(_a(pkb), _a(sig), _a(mb))
    assert st[0] == o.STATUS_INVALID_SIGNATURE


def test_c_oracle_primitives_match_python():
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = int.from_bytes(rng.bytes(32), "little") % o.Q
        b = int.from_bytes(rng.bytes(32), "little") % o.Q
        assert co.fq_mul(a, b) == a * b % o.Q
    st = [int.from_bytes(rng.bytes(32), "little") % o.Q for _ in range(5)]
    assert co.hades_permute(st) == o.hades_permute(st)
    # 131+: beyond the generated tag table, the C port derives the SAFE tag itself (BLAKE2b), as hashlib does for the Python one
    for n in (1, 4, 5, 7, 8, 10, 15, 16, 17, 130, 131, 132, 402, 803):
        ins = [int.from_bytes(rng.bytes(32), "little") % o.Q for _ in range(n)]
        assert co.poseidon_hash(ins) == o.poseidon_hash_truncated(ins)
        assert co.poseidon_hash(ins, truncated=False) == o.poseidon_hash(ins)
    for _ in range(40):
        enc = rng.bytes(32)
        p = o.point_from_bytes(enc)
        assert co.point_decode(enc) == p
        if p is not None:
            assert co.point_is_valid(enc) == int(o.point_is_valid(p))
    t = adv.torsion()
    for order, enc in t.items():
        p = o.point_from_bytes(enc)
        assert o.pmul(p, order) == o.IDENTITY and o.pmul(p, order // 2) != o.IDENTITY
        assert co.point_is_valid(enc) == 0


@pytest.mark.parametrize("kind,n", [("single", 124), ("double", 96), ("vargen", 80)])
def test_c_oracle_matches_python_adversarial(kind, n):
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    cver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    pver = {"single": o.verify_single, "double": o.verify_double, "vargen": o.verify_var_gen}[kind]
    pk, sig, msg = gen(0xB200, n)
    pk, sig, msg, expected, names = adv.make_adversarial(kind, pk, sig, msg, seed=5, frac=0.85)
    st, c = cver(pk, sig, msg)
    for i in range(n):
        ps, pc = pver(pk[i].tobytes(), sig[i].tobytes(), msg[i].tobytes())
        assert ps == st[i] == expected[i], (i, names[i], ps, st[i], expected[i])
        assert (pc or bytes(32)) == c[i].tobytes(), (i, names[i])
    assert set(st.tolist()) == {0, 1, 2, 3}


@pytest.mark.parametrize("kind,n", [("single", 64), ("double", 40), ("vargen", 40)])
def test_c_oracle_matches_python_on_bitflip_fuzz(kind, n):
    """The two oracles (big-int Python pinned on the reference's vectors, C used at scale) agree on random bit flips,
    and so do the torsion-shifted forgeries that only Signature::is_valid rejects."""
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    cver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    pver = {"single": o.verify_single, "double": o.verify_double, "vargen": o.verify_var_gen}[kind]
    pk, sig, msg = gen(0xF022, n)
    pk, sig, msg = adv.bitflip_fuzz(pk, sig, msg, seed=17)
    fpk, fsig, fmsg, fexp = adv.torsion_shifted_signatures(kind, 6, seed=2)
    pk, sig, msg = np.concatenate([pk, fpk]), np.concatenate([sig, fsig]), np.concatenate([msg, fmsg])
    st, c = cver(pk, sig, msg)
    assert (st[n:] == 2).all()
    for i in range(n + 6):
        ps, pc = pver(pk[i].tobytes(), sig[i].tobytes(), msg[i].tobytes())
        assert ps == st[i], (i, ps, st[i])
        assert (pc or bytes(32)) == c[i].tobytes(), i


def test_c_oracle_aggregate_matches_python():
    signers = [1, 2, 3, 4, 2, 3]
    pks, off, sig, msg = co.gen_aggregate(11, signers)
    sig[2, 0] ^= 1  # tamper u of one item
    pks[off[3]] = _a(adv.torsion()[4])  # a small-order signer key (aggregate_pk does not validate inputs)
    st, c, agg = co.verify_aggregate(pks, off, sig, msg)
    for i in range(len(signers)):
        ps, pc, pa = o.verify_aggregate([pks[j].tobytes() for j in range(off[i], off[i + 1])], sig[i].tobytes(), msg[i].tobytes())
        assert ps == st[i] and (pc or bytes(32)) == c[i].tobytes() and pa == agg[i].tobytes()
    assert st.tolist()[:3] == [0, 0, 1]


def test_c_oracle_aggregate_has_no_signer_limit():
    """multisig::aggregate_pk takes any number of signers (reference src/multisig.rs:393-429); 65 and 200 lie beyond the 64 of round 1."""
    signers = [65, 200]
    pks, off, sig, msg = co.gen_aggregate(23, signers)
    st, c, agg = co.verify_aggregate(pks, off, sig, msg)
    assert st.tolist() == [0, 0]
    ps, pc, pa = o.verify_aggregate([pks[j].tobytes() for j in range(off[0], off[1])], sig[0].tobytes(), msg[0].tobytes())
    assert ps == 0 and pc == c[0].tobytes() and pa == agg[0].tobytes()


def test_generated_batches_are_deterministic_and_shardable():
    a = co.gen_single(42, 64)
    b0 = co.gen_single(42, 32, first=0)
    b1 = co.gen_single(42, 32, first=32)
    for x, y0, y1 in zip(a, b0, b1):
        assert np.array_equal(x, np.concatenate([y0, y1]))


def test_multisig_combine_oracles_and_kat():
    """multisig::combine: both oracles reproduce the pinned shares -> signature step of reference src/multisig.rs:544-735
    and agree with each other on generated sessions (incl. a tampered share)."""
    k = KAT["multisig_kat"]
    arr = lambda xs: _a(b"".join(bytes.fromhex(x) for x in xs))
    st, bad, sig, ok = co.multisig_combine(arr(k["PUBLIC_KEYS"]), arr(k["R_POINTS"]), arr(k["S_POINTS"]), arr(k["INDIVIDUAL_SHARES"]), [0, 3], _a(o.le32(31)))
    assert st[0] == 0 and sig.tobytes().hex() == k["SIGNATURE"] and ok.tolist() == [1, 1, 1]
    ps, bi, psig, oks = o.multisig_combine([bytes.fromhex(x) for x in k["INDIVIDUAL_SHARES"]], [bytes.fromhex(x) for x in k["PUBLIC_KEYS"]],
                                           [bytes.fromhex(x) for x in k["R_POINTS"]], [bytes.fromhex(x) for x in k["S_POINTS"]], o.le32(31))
    assert ps == 0 and psig.hex() == k["SIGNATURE"] and all(oks)
    pks, Rs, Ss, zs, off, msg = co.gen_multisig(5, [1, 2, 3, 4, 2])
    zs[off[2] + 1, 0] ^= 1
    st, bad, sig, ok = co.multisig_combine(pks, Rs, Ss, zs, off, msg)
    for i in range(5):
        a, b = off[i], off[i + 1]
        r = o.multisig_combine([zs[j].tobytes() for j in range(a, b)], [pks[j].tobytes() for j in range(a, b)],
                               [Rs[j].tobytes() for j in range(a, b)], [Ss[j].tobytes() for j in range(a, b)], msg[i].tobytes())
        assert r[0] == st[i] and (r[2] or bytes(64)) == sig[i].tobytes() and [int(x) for x in r[3]] == ok[a:b].tolist()
        if st[i] == 0:  # the combined signature verifies under the aggregate key
            agg = o.point_to_bytes(o.aggregate_pk([o.point_from_bytes(pks[j].tobytes()) for j in range(a, b)]))
            assert co.verify_single(_a(agg), sig[i], msg[i])[0][0] == 0
    assert st.tolist() == [0, 0, 5, 0, 0] and bad[2] == 1
