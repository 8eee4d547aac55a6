"""The reference's own integration tests, restated against the host-side mirror of its API (jubjub_schnorr_b200.api):
tests/schnorr.rs, tests/schnorr_double.rs, tests/schnorr_var_generator.rs, src/multisig.rs KAT, plus verify_batch."""
import json
import os

import numpy as np
import pytest

from oracle import jjs_oracle as o
from tests.test_oracle_kat import legacy_double_fixture

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))


@pytest.fixture(scope="module")
def api():
    from jubjub_schnorr_b200 import BatchVerifier
    from jubjub_schnorr_b200 import api as a
    bv = BatchVerifier([0])
    a.set_default_verifier(bv)
    yield a
    a.set_default_verifier(None)
    bv.close()


def _keys(api, variant, seed, n=1, sk=None):
    """(secret scalars, messages, public keys, signatures) signed on the GPU like SecretKey::sign."""
    rng = o.StdRng(seed)
    sks = [rng.random_fr() if sk is None else sk for _ in range(n)]
    msgs = [rng.random_fq() for _ in range(n)]
    rnds = [rng.random_fr() for _ in range(n)]
    gs = [rng.random_fr() for _ in range(n)]
    arr = lambda xs: np.frombuffer(b"".join(o.le32(x) for x in xs), dtype=np.uint8)
    pk, sig = api.default_verifier().sign_batch(variant, arr(sks), arr(rnds), arr(msgs), arr(gs) if variant == 2 else None)
    return sks, msgs, pk, sig


def test_sign_verify(api):  # tests/schnorr.rs:16-28
    _, msgs, pk, sig = _keys(api, 0, 0xBEEF)
    pk, sig = api.PublicKey.from_bytes(pk[0].tobytes()), api.Signature.from_bytes(sig[0].tobytes())
    assert pk.is_valid()
    assert pk.verify(sig, msgs[0]) is None


def test_test_wrong_keys(api):  # tests/schnorr.rs:30-43
    _, msgs, pk, sig = _keys(api, 0, 0xBEEF)
    _, _, wrong, _ = _keys(api, 0, 0xBEE0)
    sig = api.Signature.from_bytes(sig[0].tobytes())
    assert api.PublicKey.from_bytes(wrong[0].tobytes()).verify(sig, msgs[0]) == api.Error.InvalidSignature


def test_to_from_bytes(api):  # tests/schnorr.rs:45-56
    _, _, pk, sig = _keys(api, 0, 0xBEEF)
    assert api.Signature.from_bytes(sig[0].tobytes()).to_bytes() == sig[0].tobytes()
    assert api.PublicKey.from_bytes(pk[0].tobytes()).to_bytes() == pk[0].tobytes()
    with pytest.raises(api.BytesError):
        api.PublicKey.from_bytes(o.le32(o.Q + 1))                      # v >= q
    with pytest.raises(api.BytesError):
        api.Signature.from_bytes(o.le32(o.R_ORDER) + pk[0].tobytes())  # u >= r
    with pytest.raises(api.BytesError):
        api.PublicKey.from_bytes(b"\x01" * 31)                          # wrong length


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_sign_verify_identity_fails(api, variant):  # tests/schnorr.rs:58-66, schnorr_double.rs:61-69, schnorr_var_generator.rs:116-124
    _, msgs, pk, sig = _keys(api, variant, 0xBEEF, sk=0)
    K = [api.PublicKey, api.PublicKeyDouble, api.PublicKeyVarGen][variant]
    S = [api.Signature, api.SignatureDouble, api.SignatureVarGen][variant]
    key = K.from_bytes(pk[0].tobytes())
    assert not key.is_valid()
    assert key.verify(S.from_bytes(sig[0].tobytes()), msgs[0]) == api.Error.InvalidPoint


def test_double_and_var_gen_round_trip(api):  # tests/schnorr_double.rs:18-32, tests/schnorr_var_generator.rs:19-33
    _, msgs, pk, sig = _keys(api, 1, 0xBEEF)
    assert api.PublicKeyDouble.from_bytes(pk[0].tobytes()).verify(api.SignatureDouble.from_bytes(sig[0].tobytes()), msgs[0]) is None
    _, msgs, pk, sig = _keys(api, 2, 0xBEEF)
    key = api.PublicKeyVarGen.from_bytes(pk[0].tobytes())
    assert key.verify(api.SignatureVarGen.from_bytes(sig[0].tobytes()), msgs[0]) is None
    # cross-generator forgery (tests/schnorr_var_generator.rs:61-113): same pk bytes under another generator must fail
    _, _, pk2, _ = _keys(api, 2, 0xBEE1)
    other = api.PublicKeyVarGen.from_bytes(pk[0, :32].tobytes() + pk2[0, 32:].tobytes())
    assert other.verify(api.SignatureVarGen.from_bytes(sig[0].tobytes()), msgs[0]) == api.Error.InvalidSignature


def test_adaptive_secondary_key_is_rejected(api):  # tests/schnorr_double.rs:72-82 with tests/common/mod.rs:23-66
    pkb, sig, mb = legacy_double_fixture()
    key = api.PublicKeyDouble.from_bytes(pkb)
    assert key.is_valid()
    assert key.verify(api.SignatureDouble.from_bytes(sig), mb) == api.Error.InvalidSignature


def test_multisig_aggregate_key_known_answer(api):  # src/multisig.rs:544-735
    k = KAT["multisig_kat"]
    keys = [api.PublicKey.from_bytes(bytes.fromhex(x)) for x in k["PUBLIC_KEYS"]]
    agg = api.multisig_aggregate_pk(keys)
    assert agg.to_bytes().hex() == k["AGGREGATE_PUBLIC_KEY"]
    assert agg.verify(api.Signature.from_bytes(bytes.fromhex(k["SIGNATURE"])), 31) is None
    # rogue-key style check (tests/schnorr_multisig.rs:284-323): the signature does not verify under the plain key sum
    plain = o.IDENTITY
    for x in k["PUBLIC_KEYS"]:
        plain = o.padd(plain, o.point_from_bytes(bytes.fromhex(x)))
    assert api.PublicKey.from_bytes(o.point_to_bytes(plain)).verify(api.Signature.from_bytes(bytes.fromhex(k["SIGNATURE"])), 31) \
        == api.Error.InvalidSignature


def test_verify_batch(api):
    _, msgs, pk, sig = _keys(api, 0, 7, n=64)
    items = [(api.PublicKey.from_raw_unchecked(pk[i].tobytes()), api.Signature.from_raw_unchecked(sig[i].tobytes()), msgs[i]) for i in range(64)]
    items[3] = (items[4][0], items[3][1], items[3][2])              # wrong key
    items[9] = (items[9][0], items[9][1], (msgs[9] + 1) % o.Q)      # wrong message
    got = api.verify_batch(items)
    assert got == [i not in (3, 9) for i in range(64)]
    assert api.verify_batch([]) == []


def test_serde_base58_round_trip_and_verify(api):  # tests/serde.rs:34-142 (the pinned strings, decoded and verified on the GPU)
    s = KAT["serde_kat"]
    pk = api.PublicKey.from_base58(s["serde_public_key"])
    sig = api.Signature.from_base58(s["serde_signature"])
    assert pk.to_base58() == s["serde_public_key"] and sig.to_base58() == s["serde_signature"]
    m = bytes.fromhex("6dfe107145b1cba63d5f5ed0c410c09441fbc0d70c9bfea970949499aa128214")
    assert pk.verify(sig, m) is None
    pkd = api.PublicKeyDouble.from_base58(s["serde_public_key_double"])
    assert pkd.verify(api.SignatureDouble.from_base58(s["serde_signature_double"]), m) is None
    rng = o.StdRng(s["_seed"]); rng.random_fr(); rng.random_fr()
    pkv = api.PublicKeyVarGen.from_base58(s["serde_public_key_var_gen"])
    assert pkv.verify(api.SignatureVarGen.from_base58(s["serde_signature_var_gen"]), rng.random_fq()) is None
    for bad in (s["serde_too_long_encoded"], s["serde_too_short_encoded"]):  # tests/serde.rs:145-175
        with pytest.raises(api.BytesError):
            api.PublicKey.from_base58(bad)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4097, 300_001])
def test_verify_batch_bitmap(n):
    """jjs_verify_batch: the accept bitmap packed on the GPU equals status == 0 of jjs_verify_single, for ragged sizes
    (last word partly filled, several pipeline slices); jjs_status_bitmap_device packs a device status array the same."""
    import torch
    from jubjub_schnorr_b200 import BatchVerifier
    from jubjub_schnorr_b200 import workload as wl
    with BatchVerifier([0]) as bv:
        pk, sig, msg, expected, _ = wl.make_batch(bv, 0, n, 0.3, seed=n)
        words = bv.verify_batch(pk, sig, msg)
        assert words.shape == ((n + 31) // 32,)
        bits = bv.unpack_bitmap(words, n)
        assert np.array_equal(bits, expected == 0)
        if n % 32:
            assert int(words[-1]) >> (n % 32) == 0
        d_st = torch.from_numpy(expected).cuda()
        d_w = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
        bv.status_bitmap_device(d_st.data_ptr(), n, d_w.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_w.cpu().numpy().view(np.uint32), words)
