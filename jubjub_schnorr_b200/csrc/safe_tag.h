// Host-side computation of the dusk-safe sponge tag for an arbitrary IO pattern [Absorb(n), Squeeze(1)], Domain::Other
// (SURVEY.md appendix A.6): tag = BLAKE2b-512(be32(0x80000000 | n) || be32(1) || be64(0)) read little-endian, reduced
// mod q, delivered in the Montgomery form the kernels keep field elements in.
//
// The kernels carry a constant table for the transcripts of fixed length (5, 7 and 10 inputs).  multisig::aggregate_pk
// hashes 2 + 2 n elements for n signers and multisig::combine 3 + 4 n, with no upper limit on n in the reference
// (src/multisig.rs:393-429, 440-500), so those tags are computed here at call time and handed to the device as a table
// indexed by the number of absorbed elements.  Plain C++ (no CUDA): included by kernels.cu and by tests/hostsim.
#pragma once
#include <stdint.h>
#include <string.h>

namespace jjs {
namespace safe_tag_detail {

inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

// unkeyed BLAKE2b, 64-byte digest, message of at most one block (RFC 7693)
inline void blake2b_512_short(uint8_t out[64], const uint8_t* msg, size_t len) {
    static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    static const uint8_t SIGMA[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t h[8], m[16], v[16];
    uint8_t block[128];
    memset(block, 0, sizeof(block));
    memcpy(block, msg, len < 128 ? len : 128);
    for (int i = 0; i < 8; i++) h[i] = IV[i];
    h[0] ^= 0x01010000ull ^ 64ull;  // digest length 64, no key, fanout = depth = 1
    for (int i = 0; i < 16; i++) {
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | block[8 * i + b];
        m[i] = w;
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = IV[i]; }
    v[12] ^= (uint64_t)len;   // byte counter (low word); the high word stays 0
    v[14] = ~v[14];           // final block
    auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
        v[a] = v[a] + v[b] + x; v[d] = rotr64(v[d] ^ v[a], 32);
        v[c] = v[c] + v[d];     v[b] = rotr64(v[b] ^ v[c], 24);
        v[a] = v[a] + v[b] + y; v[d] = rotr64(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];     v[b] = rotr64(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) {
        uint64_t w = h[i] ^ v[i] ^ v[8 + i];
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(w >> (8 * b));
    }
}

// acc = (2 acc + bit) mod q on 8 x 32-bit limbs (acc < q on entry)
inline void shl1_mod_q(uint32_t acc[8], uint32_t bit) {
    static const uint32_t Q[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    uint32_t carry = bit;
    for (int i = 0; i < 8; i++) {
        uint32_t nx = acc[i] >> 31;
        acc[i] = (acc[i] << 1) | carry;
        carry = nx;
    }
    // 2 acc + bit < 2 q < 2^256: one conditional subtraction
    uint32_t s[8];
    int64_t br = 0;
    for (int i = 0; i < 8; i++) {
        br += (int64_t)acc[i] - (int64_t)Q[i];
        s[i] = (uint32_t)br;
        br >>= 32;
    }
    if (br == 0)
        for (int i = 0; i < 8; i++) acc[i] = s[i];
}

}  // namespace safe_tag_detail

// tag for [Absorb(n_absorb), Squeeze(1)], Domain::Other, as 8 little-endian Montgomery limbs (value * 2^256 mod q)
inline void safe_tag_mont(uint32_t out[8], uint32_t n_absorb) {
    using namespace safe_tag_detail;
    uint8_t msg[16], digest[64];
    const uint32_t w0 = 0x80000000u | n_absorb, w1 = 1u;
    msg[0] = (uint8_t)(w0 >> 24); msg[1] = (uint8_t)(w0 >> 16); msg[2] = (uint8_t)(w0 >> 8); msg[3] = (uint8_t)w0;
    msg[4] = (uint8_t)(w1 >> 24); msg[5] = (uint8_t)(w1 >> 16); msg[6] = (uint8_t)(w1 >> 8); msg[7] = (uint8_t)w1;
    memset(msg + 8, 0, 8);
    blake2b_512_short(digest, msg, 16);
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int bit = 511; bit >= 0; bit--) shl1_mod_q(acc, (digest[bit >> 3] >> (bit & 7)) & 1u);   // digest mod q (little-endian integer)
    for (int k = 0; k < 256; k++) shl1_mod_q(acc, 0);                                             // times 2^256
    for (int i = 0; i < 8; i++) out[i] = acc[i];
}

}  // namespace jjs
