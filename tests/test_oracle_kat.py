"""Pin the Python oracle against every known-answer vector the reference holds for the verify path.

Vectors: tests/golden/reference_kat.json (extracted by tools/extract_golden.py from
reference src/multisig.rs:544-735 and tests/serde.rs:34-142) and the deterministic reject fixture of
reference tests/common/mod.rs:23-66 / tests/schnorr_double.rs:72-82 (rebuilt here from its integers).
"""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import jjs_oracle as o  # noqa: E402

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))


def test_constants_anchors():
    # SURVEY Appendix A.5 anchors (canonical integers)
    assert o.ROUND_CONSTANTS[0] == 0x6D67DFB07C22C6FD0B22407B580659556E7C8F8B712CAB9E973D2BB834DE71C5
    assert o.ROUND_CONSTANTS[339] == 0x33FAD9B52943648E1098BBFFA93E758268B8EB565DA6C8390BF6E77A839652ED
    assert o.MDS[0][0] == 0x04D4237855C1011651E8DCC995BF433111B424CB999A419A0000000066666666
    assert o.on_curve(o.G) and o.on_curve(o.G_NUMS)
    assert o.is_torsion_free(o.G) and o.is_torsion_free(o.G_NUMS)
    assert o.point_to_bytes(o.G).hex() == "12" + "00" * 31
    assert o.safe_tag(5) == 0x4A160E2860BF61DBE4F2307D562BC8B987B234208740A3C8DA695AA49D726B0E
    assert o.poseidon_hash_truncated([1, 2, 3, 4, 5]) == \
        0x001D08FA186DD1B0071C141EFAA53F49EC4BAB6A1F6D189D5963282CC69FBB5A


def test_multisig_transcript_known_answer():
    """reference src/multisig.rs:544-735, same assertion order."""
    k = KAT["multisig_kat"]
    sks, rs, ss, m = k["_inputs"]["sk"], k["_inputs"]["r"], k["_inputs"]["s"], k["_inputs"]["m"]
    pks = [o.pmul(o.G, s) for s in sks]
    Rs = [o.pmul(o.G, s) for s in rs]
    Ss = [o.pmul(o.G, s) for s in ss]
    assert [o.point_to_bytes(p).hex() for p in pks] == k["PUBLIC_KEYS"]
    assert [o.point_to_bytes(p).hex() for p in Rs] == k["R_POINTS"]
    assert [o.point_to_bytes(p).hex() for p in Ss] == k["S_POINTS"]
    ds = [o.delinearization_coeff(p, pks) for p in pks]
    assert [o.le32(d).hex() for d in ds] == k["DELINEARIZATION"]
    agg = o.aggregate_pk(pks)
    assert o.point_to_bytes(agg).hex() == k["AGGREGATE_PUBLIC_KEY"]
    pre = [agg[0], agg[1], m]
    for R, S in zip(Rs, Ss):
        pre += [R[0], R[1], S[0], S[1]]
    a = o.poseidon_hash_truncated(pre)
    assert o.le32(a).hex() == k["BINDING_COEFFICIENT"]
    RSa = o.IDENTITY
    for R, S in zip(Rs, Ss):
        RSa = o.padd(RSa, o.padd(R, o.pmul(S, a)))
    assert o.point_to_bytes(RSa).hex() == k["AGGREGATE_COMMITMENT"]
    c = o.challenge_single(RSa, agg, m)
    assert o.le32(c).hex() == k["CHALLENGE"]
    zs = [(r + a * s - c * d * sk) % o.R_ORDER for r, s, d, sk in zip(rs, ss, ds, sks)]
    assert [o.le32(z).hex() for z in zs] == k["INDIVIDUAL_SHARES"]
    sig = o.le32(sum(zs) % o.R_ORDER) + o.point_to_bytes(RSa)
    assert sig.hex() == k["SIGNATURE"]
    # decode the pinned bytes and run the wire-level verify paths
    st, cb = o.verify_single(bytes.fromhex(k["AGGREGATE_PUBLIC_KEY"]), bytes.fromhex(k["SIGNATURE"]), o.le32(m))
    assert st == o.STATUS_OK and cb.hex() == k["CHALLENGE"]
    st, cb, aggb = o.verify_aggregate([bytes.fromhex(x) for x in k["PUBLIC_KEYS"]],
                                      bytes.fromhex(k["SIGNATURE"]), o.le32(m))
    assert st == o.STATUS_OK and cb.hex() == k["CHALLENGE"] and aggb.hex() == k["AGGREGATE_PUBLIC_KEY"]


def test_serde_vectors_single_and_double():
    """reference tests/serde.rs:34-101 (StdRng seed 2321: sk, then msg, then the nonce scalar)."""
    s = KAT["serde_kat"]
    rng = o.StdRng(s["_seed"])
    sk = rng.random_fr()
    assert o.b58encode(o.le32(sk)) == s["serde_secret_key"]
    pk = o.pmul(o.G, sk)
    assert o.b58encode(o.point_to_bytes(pk)) == s["serde_public_key"]
    pkp = o.pmul(o.G_NUMS, sk)
    assert o.b58encode(o.point_to_bytes(pk) + o.point_to_bytes(pkp)) == s["serde_public_key_double"]
    m = rng.random_fq()
    rnd = rng.random_fr()
    u, R = o.sign_single(sk, rnd, m)
    sig = o.le32(u) + o.point_to_bytes(R)
    assert o.b58encode(sig) == s["serde_signature"]
    st, c = o.verify_single(o.point_to_bytes(pk), sig, o.le32(m))
    assert st == o.STATUS_OK
    # SURVEY Appendix B.2 derived anchors
    assert o.le32(m).hex() == "6dfe107145b1cba63d5f5ed0c410c09441fbc0d70c9bfea970949499aa128214"
    assert c.hex() == "7ad531e479fe4f1d2c1858c180e18f57e549d7eda93c85f46344f4716de67e02"
    u2, R2, Rp2 = o.sign_double(sk, rnd, m)
    sigd = o.le32(u2) + o.point_to_bytes(R2) + o.point_to_bytes(Rp2)
    assert o.b58encode(sigd) == s["serde_signature_double"]
    st, c = o.verify_double(o.point_to_bytes(pk) + o.point_to_bytes(pkp), sigd, o.le32(m))
    assert st == o.STATUS_OK
    assert c.hex() == "b706ff0423cbdff51e73ee23985e66a5f827343f03ecf4e9ecd7f32fc9791901"


def test_serde_vectors_var_gen():
    """reference tests/serde.rs:104-142 (draw order: sk, generator scalar, msg, nonce scalar)."""
    s = KAT["serde_kat"]
    rng = o.StdRng(s["_seed"])
    sk = rng.random_fr()
    gen = o.pmul(o.G, rng.random_fr())
    assert o.b58encode(o.le32(sk) + o.point_to_bytes(gen)) == s["serde_secret_key_var_gen"]
    pk = o.pmul(gen, sk)
    pkb = o.point_to_bytes(pk) + o.point_to_bytes(gen)
    assert o.b58encode(pkb) == s["serde_public_key_var_gen"]
    m = rng.random_fq()
    rnd = rng.random_fr()
    u, R = o.sign_var_gen(sk, gen, rnd, m)
    sig = o.le32(u) + o.point_to_bytes(R)
    assert o.b58encode(sig) == s["serde_signature_var_gen"]
    st, c = o.verify_var_gen(pkb, sig, o.le32(m))
    assert st == o.STATUS_OK
    assert c.hex() == "648cf37f901b93870bec5cb3d7339934efd8307d7c667a3718cd14643a677603"


def legacy_double_fixture():
    """reference tests/common/mod.rs:23-66: adaptive secondary key against the legacy transcript."""
    sk, m, nonce = 17, 23, 31
    pk = o.pmul(o.G, sk)
    R = o.pmul(o.G, nonce)
    Rp = o.pmul(o.G_NUMS, 37)
    legacy_c = o.poseidon_hash_truncated([R[0], R[1], Rp[0], Rp[1], pk[0], pk[1], m])
    u = (nonce - legacy_c * sk) % o.R_ORDER
    pkp = o.pmul(o.padd(Rp, o.pneg(o.pmul(o.G_NUMS, u))), pow(legacy_c, -1, o.R_ORDER))
    sig = o.le32(u) + o.point_to_bytes(R) + o.point_to_bytes(Rp)
    return o.point_to_bytes(pk) + o.point_to_bytes(pkp), sig, o.le32(m)


def test_adaptive_secondary_key_is_rejected():
    """reference tests/schnorr_double.rs:72-82: key valid, verify -> InvalidSignature."""
    pkb, sig, mb = legacy_double_fixture()
    # SURVEY Appendix B.3 bytes
    assert sig.hex() == ("b9f3dc4321caaca2c32c6b33718bae703c3845a1889be18b9535f1ec12465f0a"
                         "80da3101692e3ed7208f88ddca700ab6acb35fa21e8e83e50d212bde3568af3f"
                         "2859a98c0dbcd98bf578e1602627b0c107be24a05ac5474edea512b44f06d901")
    assert pkb.hex() == ("030fa17156c36af83deba1a959c722cef3ae9c5c23a1ee36aa5f22175cd4b3b2"
                         "3e1dd46bae1f94b2ab35e3f89c48a2fe8f54f2d3497a10d25f77bf2b8555fb91")
    assert o.point_is_valid(o.point_from_bytes(pkb[:32])) and o.point_is_valid(o.point_from_bytes(pkb[32:]))
    st, _ = o.verify_double(pkb, sig, mb)
    assert st == o.STATUS_INVALID_SIGNATURE


@pytest.mark.parametrize("variant", ["single", "double", "var_gen"])
def test_identity_key_is_invalid_point(variant):
    """reference tests/schnorr.rs:58-66, tests/schnorr_double.rs:61-69,
    tests/schnorr_var_generator.rs:116-124: sk = 0 -> InvalidPoint although the equation holds."""
    rng = o.StdRng(0xBEEF)
    m = rng.random_fq()
    rnd = rng.random_fr()
    ident = o.point_to_bytes(o.IDENTITY)
    if variant == "single":
        u, R = o.sign_single(0, rnd, m)
        st, _ = o.verify_single(ident, o.le32(u) + o.point_to_bytes(R), o.le32(m))
    elif variant == "double":
        u, R, Rp = o.sign_double(0, rnd, m)
        st, _ = o.verify_double(ident + ident, o.le32(u) + o.point_to_bytes(R) + o.point_to_bytes(Rp), o.le32(m))
    else:
        gen = o.pmul(o.G, 12345)
        u, R = o.sign_var_gen(0, gen, rnd, m)
        st, _ = o.verify_var_gen(ident + o.point_to_bytes(gen), o.le32(u) + o.point_to_bytes(R), o.le32(m))
    assert st == o.STATUS_INVALID_POINT


def test_wrong_key_is_invalid_signature():
    """reference tests/schnorr.rs:30-43."""
    rng = o.StdRng(0xBEEF)
    sk, wrong, m, rnd = rng.random_fr(), rng.random_fr(), rng.random_fq(), rng.random_fr()
    u, R = o.sign_single(sk, rnd, m)
    sig = o.le32(u) + o.point_to_bytes(R)
    assert o.verify_single(o.point_to_bytes(o.pmul(o.G, sk)), sig, o.le32(m))[0] == o.STATUS_OK
    assert o.verify_single(o.point_to_bytes(o.pmul(o.G, wrong)), sig, o.le32(m))[0] == o.STATUS_INVALID_SIGNATURE
