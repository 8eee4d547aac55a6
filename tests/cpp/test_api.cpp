// The reference's integration tests restated against include/jubjub_schnorr.hpp (run by tests/test_gpu_cpp_api.py,
// which generates kat.h from tests/golden/reference_kat.json).  Exit code 0 = all assertions held.
#include <cstdio>
#include <cstdlib>

#include "../../include/jubjub_schnorr.hpp"
#include "kat.h"

using namespace jubjub_schnorr;
#define CHECK(x) do { if (!(x)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); std::exit(1); } } while (0)

int main() {
    Gpu gpu({0});
    const BlsScalar m31 = bls_scalar_from_u64(31);
    // src/multisig.rs:544-735: aggregate key, then the pinned signature verifies under it
    std::vector<PublicKey> keys;
    for (int i = 0; i < 3; i++) keys.push_back(PublicKey::from_bytes(gpu, KAT_PUBLIC_KEYS[i]));
    PublicKey agg = multisig::aggregate_pk(gpu, keys);
    CHECK(std::memcmp(agg.to_bytes().data(), KAT_AGGREGATE_PUBLIC_KEY, 32) == 0);
    Signature sig = Signature::from_bytes(gpu, KAT_SIGNATURE);
    CHECK(agg.is_valid(gpu));
    CHECK(!agg.verify(gpu, sig, m31).has_value());                                   // tests/schnorr.rs:16-28
    CHECK(keys[0].verify(gpu, sig, m31) == Error::InvalidSignature);                 // tests/schnorr.rs:30-43 (wrong key)
    CHECK(agg.verify(gpu, sig, bls_scalar_from_u64(32)) == Error::InvalidSignature); // wrong message
    // identity key -> InvalidPoint                                                    tests/schnorr.rs:58-66
    uint8_t ident[32] = {1};
    PublicKey id = PublicKey::from_bytes(gpu, ident);
    CHECK(!id.is_valid(gpu));
    CHECK(id.verify(gpu, sig, m31) == Error::InvalidPoint);
    // from_bytes rejects malformed encodings                                          tests/schnorr.rs:45-56
    uint8_t bad[32];
    std::memset(bad, 0xff, 32);
    bool threw = false;
    try { PublicKey::from_bytes(gpu, bad); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    // a key that cannot be decoded surfaces as BytesError from verify when built unchecked
    CHECK(PublicKey::from_raw_unchecked(bad).verify(gpu, sig, m31) == Error::BytesError);
    // double: adaptive secondary key fixture -> InvalidSignature                     tests/schnorr_double.rs:72-82
    PublicKeyDouble pkd = PublicKeyDouble::from_bytes(gpu, KAT_LEGACY_DOUBLE_PK);
    SignatureDouble sgd = SignatureDouble::from_bytes(gpu, KAT_LEGACY_DOUBLE_SIG);
    CHECK(pkd.verify(gpu, sgd, bls_scalar_from_u64(23)) == Error::InvalidSignature);
    // verify_batch
    std::vector<std::tuple<PublicKey, Signature, BlsScalar>> items = {{agg, sig, m31}, {keys[1], sig, m31}, {agg, sig, m31}};
    std::vector<bool> ok = verify_batch(gpu, items);
    CHECK(ok.size() == 3 && ok[0] && !ok[1] && ok[2]);
    CHECK(verify_batch(gpu, {}).empty());
    std::puts("cpp api ok");
    return 0;
}
