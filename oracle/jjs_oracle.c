/*
 * CPU oracle (plain C) for the jubjub-schnorr verify path.  TEST INFRASTRUCTURE ONLY.
 *
 * A restatement of the algorithm the reference (dusk-network/jubjub-schnorr 0.7.0-rc.0) executes on
 * the CPU for PublicKey::verify and its double / var-generator / aggregate-key siblings, written on
 * 4x64-bit Montgomery limbs.  It deliberately keeps the reference's *algorithms* (four 252-step
 * double-and-add scalar multiplications per single verify, dense-MDS Hades rounds, per-point
 * inversions, same check order) so that it can also serve as the timed "port" CPU baseline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (jubjub_schnorr_b200/) never does.
 *
 * Parity status: PINNED -- cross-checked against oracle/jjs_oracle.py, which itself reproduces every
 * known-answer vector of the reference (tests/test_oracle_kat.py), and directly against those
 * vectors (tests/test_c_oracle.py).  The field/curve/hash primitives come from crates that are not
 * vendored in the reference tree (dusk-bls12_381 0.14, dusk-jubjub 0.15, dusk-poseidon 0.42.0-rc.0,
 * dusk-safe); their published algorithms are restated from SURVEY.md Appendix A.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "jjs_oracle_constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;                 /* 256-bit residue, Montgomery form */
typedef struct { const uint64_t *m; uint64_t inv; const uint64_t *r1, *r2; fe r3; } field;

static field FQ, FR;

/* ------------------------------------------------------------------ generic Montgomery field */
static int ge256(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static void sub256(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)br;
        o[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static void f_mul(const field *F, fe *o, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * F->inv;
        c = ((u128)k * F->m[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)k * F->m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || ge256(t, F->m)) sub256(t, t, F->m);
    memcpy(o->l, t, 32);
}
static void f_sqr(const field *F, fe *o, const fe *a) { f_mul(F, o, a, a); }
static void f_add(const field *F, fe *o, const fe *a, const fe *b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || ge256(t, F->m)) sub256(t, t, F->m);
    memcpy(o->l, t, 32);
}
static void f_sub(const field *F, fe *o, const fe *a, const fe *b) {
    uint64_t t[4];
    if (ge256(a->l, b->l)) {
        sub256(t, a->l, b->l);
    } else {
        uint64_t s[4];
        sub256(s, b->l, a->l);
        sub256(t, F->m, s);
    }
    memcpy(o->l, t, 32);
}
static int f_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int f_eq(const fe *a, const fe *b) { return memcmp(a->l, b->l, 32) == 0; }
static void f_from_raw(const field *F, fe *o, const uint64_t raw[4]) { /* raw < m */
    fe a, r2;
    memcpy(a.l, raw, 32);
    memcpy(r2.l, F->r2, 32);
    f_mul(F, o, &a, &r2);
}
static void f_to_raw(const field *F, uint64_t raw[4], const fe *a) {
    fe one = {{1, 0, 0, 0}}, t;
    f_mul(F, &t, a, &one);
    memcpy(raw, t.l, 32);
}
static void f_one(const field *F, fe *o) { memcpy(o->l, F->r1, 32); }
static void f_zero(fe *o) { memset(o->l, 0, 32); }
/* canonical 32-byte little-endian -> Montgomery; returns 0 if value >= modulus */
static int f_from_bytes(const field *F, fe *o, const uint8_t b[32]) {
    uint64_t raw[4];
    memcpy(raw, b, 32);
    if (ge256(raw, F->m)) return 0;
    f_from_raw(F, o, raw);
    return 1;
}
static void f_to_bytes(const field *F, uint8_t b[32], const fe *a) {
    uint64_t raw[4];
    f_to_raw(F, raw, a);
    memcpy(b, raw, 32);
}
/* 64 little-endian bytes reduced mod m (BlsScalar::from_bytes_wide / JubJubScalar::random) */
static void f_from_wide(const field *F, fe *o, const uint8_t b[64]) {
    fe lo, hi, r2;
    memcpy(lo.l, b, 32);
    memcpy(hi.l, b + 32, 32);
    memcpy(r2.l, F->r2, 32);
    f_mul(F, &lo, &lo, &r2);      /* lo * R      (inputs may exceed m; CIOS still reduces below 2m -> one sub) */
    f_mul(F, &hi, &hi, &F->r3);   /* hi * R^2 = (hi * 2^256) * R */
    f_add(F, o, &lo, &hi);
}
static void f_pow(const field *F, fe *o, const fe *a, const uint64_t e[4]) {
    fe acc;
    f_one(F, &acc);
    for (int i = 255; i >= 0; i--) {
        f_sqr(F, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
    }
    *o = acc;
}
static void f_inv(const field *F, fe *o, const fe *a) { /* Fermat; inv(0) = 0 */
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub256(e, F->m, two);
    f_pow(F, o, a, e);
}

#define Q_MUL(o, a, b) f_mul(&FQ, o, a, b)
#define Q_SQR(o, a) f_sqr(&FQ, o, a)
#define Q_ADD(o, a, b) f_add(&FQ, o, a, b)
#define Q_SUB(o, a, b) f_sub(&FQ, o, a, b)

/* Tonelli-Shanks, q - 1 = 2^32 t.  Returns 0 if a is a non-residue. */
static int q_sqrt(fe *o, const fe *a) {
    if (f_is_zero(a)) { f_zero(o); return 1; }
    fe one, w, x, b, c, g;
    f_one(&FQ, &one);
    uint64_t e[4]; /* (t-1)/2 */
    memcpy(e, JJO_TS_T, 32);
    e[0] -= 1; /* t odd */
    for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
    e[3] >>= 1;
    f_pow(&FQ, &w, a, e);
    Q_MUL(&x, a, &w);   /* a^((t+1)/2) */
    Q_MUL(&b, &x, &w);  /* a^t */
    memcpy(c.l, JJO_TS_ROOT, 32);
    int m = 32;
    while (!f_eq(&b, &one)) {
        int i = 0;
        fe b2 = b;
        while (!f_eq(&b2, &one)) {
            Q_SQR(&b2, &b2);
            if (++i == m) return 0;
        }
        g = c;
        for (int k = 0; k < m - i - 1; k++) Q_SQR(&g, &g);
        Q_MUL(&x, &x, &g);
        Q_SQR(&c, &g);
        Q_MUL(&b, &b, &c);
        m = i;
    }
    *o = x;
    return 1;
}

/* ------------------------------------------------------------------ JubJub, extended coordinates */
typedef struct { fe X, Y, Z, T; } pt;
static fe D_, D2_;
static pt G_PT, GN_PT;

static void pt_identity(pt *p) { f_zero(&p->X); f_one(&FQ, &p->Y); f_one(&FQ, &p->Z); f_zero(&p->T); }
static void pt_from_affine(pt *p, const fe *u, const fe *v) {
    p->X = *u; p->Y = *v; f_one(&FQ, &p->Z); Q_MUL(&p->T, u, v);
}
static void pt_add(pt *o, const pt *p, const pt *q) { /* add-2008-hwcd-3, a = -1, complete */
    fe a, b, c, d, e, f, g, h, t0, t1;
    Q_SUB(&t0, &p->Y, &p->X); Q_SUB(&t1, &q->Y, &q->X); Q_MUL(&a, &t0, &t1);
    Q_ADD(&t0, &p->Y, &p->X); Q_ADD(&t1, &q->Y, &q->X); Q_MUL(&b, &t0, &t1);
    Q_MUL(&c, &p->T, &q->T); Q_MUL(&c, &c, &D2_);
    Q_MUL(&d, &p->Z, &q->Z); Q_ADD(&d, &d, &d);
    Q_SUB(&e, &b, &a); Q_SUB(&f, &d, &c); Q_ADD(&g, &d, &c); Q_ADD(&h, &b, &a);
    Q_MUL(&o->X, &e, &f); Q_MUL(&o->Y, &g, &h); Q_MUL(&o->T, &e, &h); Q_MUL(&o->Z, &f, &g);
}
static void pt_dbl(pt *o, const pt *p) { /* dbl-2008-hwcd, a = -1 */
    fe a, b, c, e, f, g, h, t0;
    Q_SQR(&a, &p->X); Q_SQR(&b, &p->Y); Q_SQR(&c, &p->Z); Q_ADD(&c, &c, &c);
    Q_ADD(&t0, &p->X, &p->Y); Q_SQR(&e, &t0); Q_SUB(&e, &e, &a); Q_SUB(&e, &e, &b);
    Q_SUB(&g, &b, &a);          /* D + B with D = -A */
    Q_SUB(&f, &g, &c);
    Q_ADD(&h, &a, &b); f_zero(&t0); Q_SUB(&h, &t0, &h); /* D - B = -(A + B) */
    Q_MUL(&o->X, &e, &f); Q_MUL(&o->Y, &g, &h); Q_MUL(&o->T, &e, &h); Q_MUL(&o->Z, &f, &g);
}
static void pt_neg(pt *o, const pt *p) {
    fe z; f_zero(&z);
    Q_SUB(&o->X, &z, &p->X); o->Y = p->Y; o->Z = p->Z; Q_SUB(&o->T, &z, &p->T);
}
static int pt_eq(const pt *p, const pt *q) { /* projective equality, as JubJubExtended::eq */
    fe a, b;
    Q_MUL(&a, &p->X, &q->Z); Q_MUL(&b, &q->X, &p->Z);
    if (!f_eq(&a, &b)) return 0;
    Q_MUL(&a, &p->Y, &q->Z); Q_MUL(&b, &q->Y, &p->Z);
    return f_eq(&a, &b);
}
static int pt_is_identity(const pt *p) { return f_is_zero(&p->X) && f_eq(&p->Y, &p->Z); }
static void pt_to_affine(fe *u, fe *v, const pt *p) { /* one inversion, as to_hash_inputs / JubJubAffine::from */
    fe zi;
    f_inv(&FQ, &zi, &p->Z);
    Q_MUL(u, &p->X, &zi); Q_MUL(v, &p->Y, &zi);
}
/* 252-step MSB-first double-and-always-add over a plain (non-Montgomery) scalar, as the dependency
 * does for `point * scalar` and for is_torsion_free. */
static void pt_mul_raw(pt *o, const pt *p, const uint64_t k[4]) {
    pt acc, tmp;
    pt_identity(&acc);
    for (int i = 251; i >= 0; i--) {
        pt_dbl(&acc, &acc);
        pt_add(&tmp, &acc, p);
        if ((k[i >> 6] >> (i & 63)) & 1) acc = tmp;
    }
    *o = acc;
}
static void pt_mul(pt *o, const pt *p, const fe *k_mont_fr) {
    uint64_t raw[4];
    f_to_raw(&FR, raw, k_mont_fr);
    pt_mul_raw(o, p, raw);
}
static int pt_is_torsion_free(const pt *p) {
    pt t;
    pt_mul_raw(&t, p, JJO_R);
    return pt_is_identity(&t);
}
static int pt_is_on_curve(const pt *p) { /* z != 0 and the affine equation */
    if (f_is_zero(&p->Z)) return 0;
    fe u, v, u2, v2, l, r, one;
    pt_to_affine(&u, &v, p);
    Q_SQR(&u2, &u); Q_SQR(&v2, &v);
    Q_SUB(&l, &v2, &u2);
    Q_MUL(&r, &u2, &v2); Q_MUL(&r, &r, &D_); f_one(&FQ, &one); Q_ADD(&r, &r, &one);
    return f_eq(&l, &r);
}
/* PublicKey::is_valid / Signature::is_valid: reference src/keys/public.rs:159-164, src/signatures.rs:93-98 */
static int pt_is_valid(const pt *p) {
    int ident = pt_is_identity(p);
    return pt_is_torsion_free(p) && pt_is_on_curve(p) && !ident;
}
/* JubJubAffine::from_bytes (reference src/keys/public.rs:88, src/signatures.rs:114); SURVEY A.3/A.4 */
static int pt_decode(pt *p, const uint8_t in[32]) {
    uint8_t b[32];
    memcpy(b, in, 32);
    int sign = b[31] >> 7;
    b[31] &= 0x7f;
    fe v, v2, num, den, u2, u, one;
    if (!f_from_bytes(&FQ, &v, b)) return 0;
    f_one(&FQ, &one);
    Q_SQR(&v2, &v);
    Q_SUB(&num, &v2, &one);
    Q_MUL(&den, &v2, &D_); Q_ADD(&den, &den, &one);
    f_inv(&FQ, &den, &den);
    Q_MUL(&u2, &num, &den);
    if (!q_sqrt(&u, &u2)) return 0;
    uint64_t raw[4];
    f_to_raw(&FQ, raw, &u);
    if ((int)(raw[0] & 1) != sign) { fe z; f_zero(&z); Q_SUB(&u, &z, &u); }
    if (f_is_zero(&u) && sign) return 0;
    pt_from_affine(p, &u, &v);
    return 1;
}
static void pt_encode(uint8_t out[32], const pt *p) {
    fe u, v;
    pt_to_affine(&u, &v, p);
    uint64_t raw[4];
    f_to_bytes(&FQ, out, &v);
    f_to_raw(&FQ, raw, &u);
    out[31] |= (uint8_t)((raw[0] & 1) << 7);
}

/* fixed-base tables (generation only): 63 windows of 4 bits, entries 1..15 */
typedef struct { pt e[63][15]; } fbtable;
static fbtable FB_G, FB_GN;
static void fb_build(fbtable *t, const pt *base) {
    pt b = *base;
    for (int w = 0; w < 63; w++) {
        t->e[w][0] = b;
        for (int j = 1; j < 15; j++) pt_add(&t->e[w][j], &t->e[w][j - 1], &b);
        pt_add(&b, &t->e[w][14], &b); /* 16 * b */
    }
}
static void fb_mul(pt *o, const fbtable *t, const fe *k_mont_fr) {
    uint64_t raw[4];
    f_to_raw(&FR, raw, k_mont_fr);
    pt acc;
    pt_identity(&acc);
    for (int w = 0; w < 63; w++) {
        unsigned d = (unsigned)(raw[w >> 4] >> ((w & 15) * 4)) & 15;
        if (d) pt_add(&acc, &acc, &t->e[w][d - 1]);
    }
    *o = acc;
}

/* ------------------------------------------------------------------ Poseidon (dense reference form) */
static void hades_permute(fe s[5]) {
    const fe *rc = (const fe *)JJO_RC;
    for (int rnd = 0; rnd < 68; rnd++) {
        for (int i = 0; i < 5; i++) Q_ADD(&s[i], &s[i], rc++);
        int full = rnd < 4 || rnd >= 64;
        for (int i = full ? 0 : 4; i < 5; i++) {
            fe x2, x4;
            Q_SQR(&x2, &s[i]); Q_SQR(&x4, &x2); Q_MUL(&s[i], &x4, &s[i]);
        }
        fe o[5];
        for (int i = 0; i < 5; i++) {
            fe acc, t;
            f_zero(&acc);
            for (int k = 0; k < 5; k++) {
                Q_MUL(&t, (const fe *)JJO_MDS[i][k], &s[k]);
                Q_ADD(&acc, &acc, &t);
            }
            o[i] = acc;
        }
        memcpy(s, o, sizeof(o));
    }
}
/* Hash::digest(Domain::Other, in)[0]; SAFE sponge rate 4 */
/* dusk-safe tag of the IO pattern [Absorb(n), Squeeze(1)], Domain::Other for a transcript longer than the
 * pre-generated table: BLAKE2b-512 over be32(0x80000000 | n) || be32(1) || be64(0), the digest read as a little-endian
 * integer mod q.  multisig::aggregate_pk hashes 2 + 2 n elements for n signers (reference src/multisig.rs:393-409) and
 * the reference puts no limit on n. */
#define B2B_ROR(x, r) (((x) >> (r)) | ((x) << (64 - (r))))
#define B2B_MIX(a, b, c, d, x, y) do { a += b + (x); d = B2B_ROR(d ^ a, 32); c += d; b = B2B_ROR(b ^ c, 24); \
                                       a += b + (y); d = B2B_ROR(d ^ a, 16); c += d; b = B2B_ROR(b ^ c, 63); } while (0)
static void blake2b512_one_block(uint8_t digest[64], const uint8_t *data, size_t len) { /* len <= 128, no key */
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    static const uint8_t perm[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    uint8_t buf[128] = {0};
    uint64_t h[8], w[16], v[16];
    memcpy(buf, data, len);
    memcpy(w, buf, 128); /* little-endian host */
    memcpy(h, iv, sizeof h);
    h[0] ^= 0x01010040ULL;
    memcpy(v, h, sizeof h);
    memcpy(v + 8, iv, sizeof iv);
    v[12] ^= (uint64_t)len;
    v[14] ^= ~0ULL;
    for (int round = 0; round < 12; round++) {
        const uint8_t *p = perm[round % 10];
        B2B_MIX(v[0], v[4], v[8], v[12], w[p[0]], w[p[1]]);
        B2B_MIX(v[1], v[5], v[9], v[13], w[p[2]], w[p[3]]);
        B2B_MIX(v[2], v[6], v[10], v[14], w[p[4]], w[p[5]]);
        B2B_MIX(v[3], v[7], v[11], v[15], w[p[6]], w[p[7]]);
        B2B_MIX(v[0], v[5], v[10], v[15], w[p[8]], w[p[9]]);
        B2B_MIX(v[1], v[6], v[11], v[12], w[p[10]], w[p[11]]);
        B2B_MIX(v[2], v[7], v[8], v[13], w[p[12]], w[p[13]]);
        B2B_MIX(v[3], v[4], v[9], v[14], w[p[14]], w[p[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    memcpy(digest, h, 64);
}
static void half_mod_q(fe *o, const uint8_t bytes[32]) { /* a 256-bit little-endian value -> Fq (Montgomery) */
    uint64_t raw[4];
    memcpy(raw, bytes, 32);
    while (ge256(raw, JJO_Q)) { /* at most twice: 2^256 < 3 q */
        unsigned __int128 br = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 d = (unsigned __int128)raw[i] - JJO_Q[i] - (uint64_t)br;
            raw[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }
    f_from_raw(&FQ, o, raw);
}
static void safe_tag_compute(fe *tag, uint32_t n_absorb) {
    uint8_t in[16] = {0}, digest[64];
    uint32_t w0 = 0x80000000u | n_absorb;
    in[0] = (uint8_t)(w0 >> 24); in[1] = (uint8_t)(w0 >> 16); in[2] = (uint8_t)(w0 >> 8); in[3] = (uint8_t)w0;
    in[7] = 1;
    blake2b512_one_block(digest, in, 16);
    fe lo, hi, two256;
    half_mod_q(&lo, digest);
    half_mod_q(&hi, digest + 32);
    memcpy(two256.l, JJO_Q_R2, 32); /* 2^256 mod q as a Montgomery element is R^2 mod q */
    Q_MUL(&hi, &hi, &two256);
    Q_ADD(tag, &lo, &hi);
}
static void poseidon_hash(fe *out, const fe *in, int n) {
    fe s[5];
    if (n <= JJO_MAX_ABSORB) memcpy(s[0].l, JJO_TAG[n], 32);
    else safe_tag_compute(&s[0], (uint32_t)n);
    for (int i = 1; i < 5; i++) f_zero(&s[i]);
    int pos = 0;
    for (int i = 0; i < n; i++) {
        if (pos == 4) { hades_permute(s); pos = 0; }
        Q_ADD(&s[1 + pos], &s[1 + pos], &in[i]);
        pos++;
    }
    hades_permute(s);
    *out = s[1];
}
/* digest_truncated: low 250 bits of the canonical value, read as a JubJubScalar.
 * Output: c as an Fr Montgomery element and (optionally) its canonical bytes. */
static void poseidon_hash_truncated(fe *c_fr, uint8_t c_bytes[32], const fe *in, int n) {
    fe h;
    uint8_t b[32];
    poseidon_hash(&h, in, n);
    f_to_bytes(&FQ, b, &h);
    b[31] &= 0x03;
    if (c_bytes) memcpy(c_bytes, b, 32);
    f_from_bytes(&FR, c_fr, b);
}

/* ------------------------------------------------------------------ init */
static pthread_once_t once = PTHREAD_ONCE_INIT;
static void init_impl(void) {
    FQ.m = JJO_Q; FQ.inv = JJO_Q_INV; FQ.r1 = JJO_Q_R1; FQ.r2 = JJO_Q_R2;
    FR.m = JJO_R; FR.inv = JJO_R_INV; FR.r1 = JJO_R_R1; FR.r2 = JJO_R_R2;
    fe r2;
    memcpy(r2.l, JJO_Q_R2, 32); f_mul(&FQ, &FQ.r3, &r2, &r2);
    memcpy(r2.l, JJO_R_R2, 32); f_mul(&FR, &FR.r3, &r2, &r2);
    memcpy(D_.l, JJO_D, 32);
    memcpy(D2_.l, JJO_D2, 32);
    pt_from_affine(&G_PT, (const fe *)JJO_G[0], (const fe *)JJO_G[1]);
    pt_from_affine(&GN_PT, (const fe *)JJO_G_NUMS[0], (const fe *)JJO_G_NUMS[1]);
    fb_build(&FB_G, &G_PT);
    fb_build(&FB_GN, &GN_PT);
}
static void init(void) { pthread_once(&once, init_impl); }

/* ------------------------------------------------------------------ verify (reference semantics) */
enum { ST_OK = 0, ST_INVALID_SIGNATURE = 1, ST_INVALID_POINT = 2, ST_BYTES_ERROR = 3 };

/* PublicKey::verify, reference src/keys/public.rs:114-135 */
static int verify_single_pts(const pt *pk, const fe *u, const pt *R, const fe *m, uint8_t c_out[32]) {
    if (!pt_is_valid(pk) || !pt_is_valid(R)) return ST_INVALID_POINT;
    fe in[5], c;
    pt_to_affine(&in[0], &in[1], R);   /* to_hash_inputs, src/signatures.rs:127 */
    pt_to_affine(&in[2], &in[3], pk);  /* src/signatures.rs:128 */
    in[4] = *m;
    poseidon_hash_truncated(&c, c_out, in, 5);
    pt a, b, s;
    pt_mul(&a, &G_PT, u);
    pt_mul(&b, pk, &c);
    pt_add(&s, &a, &b);
    return pt_eq(&s, R) ? ST_OK : ST_INVALID_SIGNATURE;
}
static int verify_single_one(const uint8_t pk32[32], const uint8_t sig64[64], const uint8_t msg32[32], uint8_t c_out[32]) {
    pt pk, R;
    fe u, m;
    if (c_out) memset(c_out, 0, 32);
    int ok = pt_decode(&pk, pk32);
    ok &= f_from_bytes(&FR, &u, sig64);
    ok &= pt_decode(&R, sig64 + 32);
    ok &= f_from_bytes(&FQ, &m, msg32);
    if (!ok) return ST_BYTES_ERROR;
    return verify_single_pts(&pk, &u, &R, &m, c_out);
}
/* PublicKeyDouble::verify, reference src/keys/public/double.rs:86-117 */
static int verify_double_one(const uint8_t pk64[64], const uint8_t sig96[96], const uint8_t msg32[32], uint8_t c_out[32]) {
    pt pk, pkp, R, Rp;
    fe u, m;
    if (c_out) memset(c_out, 0, 32);
    int ok = pt_decode(&pk, pk64);
    ok &= pt_decode(&pkp, pk64 + 32);
    ok &= f_from_bytes(&FR, &u, sig96);
    ok &= pt_decode(&R, sig96 + 32);
    ok &= pt_decode(&Rp, sig96 + 64);
    ok &= f_from_bytes(&FQ, &m, msg32);
    if (!ok) return ST_BYTES_ERROR;
    int pk_ok = pt_is_valid(&pk) & pt_is_valid(&pkp);
    int sig_ok = pt_is_valid(&R) & pt_is_valid(&Rp);
    if (!pk_ok || !sig_ok) return ST_INVALID_POINT;
    fe in[10], c;
    uint64_t tag[4] = {0x4a4a53434844424cULL, 0, 0, 0}; /* DOUBLE_CHALLENGE_DOMAIN, src/signatures/double.rs:24-25 */
    f_from_raw(&FQ, &in[0], tag);
    pt_to_affine(&in[1], &in[2], &R);
    pt_to_affine(&in[3], &in[4], &Rp);
    pt_to_affine(&in[5], &in[6], &pk);
    pt_to_affine(&in[7], &in[8], &pkp);
    in[9] = m;
    poseidon_hash_truncated(&c, c_out, in, 10);
    pt a, b, p1, p2;
    pt_mul(&a, &G_PT, &u); pt_mul(&b, &pk, &c); pt_add(&p1, &a, &b);
    pt_mul(&a, &GN_PT, &u); pt_mul(&b, &pkp, &c); pt_add(&p2, &a, &b);
    return (pt_eq(&p1, &R) && pt_eq(&p2, &Rp)) ? ST_OK : ST_INVALID_SIGNATURE;
}
/* PublicKeyVarGen::verify, reference src/keys/public/var_gen.rs:107-133 */
static int verify_vargen_one(const uint8_t pk64[64], const uint8_t sig64[64], const uint8_t msg32[32], uint8_t c_out[32]) {
    pt pk, gen, R;
    fe u, m;
    if (c_out) memset(c_out, 0, 32);
    int ok = pt_decode(&pk, pk64);
    ok &= pt_decode(&gen, pk64 + 32);
    ok &= f_from_bytes(&FR, &u, sig64);
    ok &= pt_decode(&R, sig64 + 32);
    ok &= f_from_bytes(&FQ, &m, msg32);
    if (!ok) return ST_BYTES_ERROR;
    int pk_ok = pt_is_valid(&pk) & pt_is_valid(&gen);
    if (!pk_ok || !pt_is_valid(&R)) return ST_INVALID_POINT;
    fe in[7], c;
    pt_to_affine(&in[0], &in[1], &R);
    pt_to_affine(&in[2], &in[3], &pk);
    pt_to_affine(&in[4], &in[5], &gen);
    in[6] = m;
    poseidon_hash_truncated(&c, c_out, in, 7);
    pt a, b, s;
    pt_mul(&a, &gen, &u); pt_mul(&b, &pk, &c); pt_add(&s, &a, &b);
    return pt_eq(&s, &R) ? ST_OK : ST_INVALID_SIGNATURE;
}
/* multisig::aggregate_pk (reference src/multisig.rs:154-156, 393-429) then PublicKey::verify */
static int aggregate_pts(pt *agg, const pt *pks, int n) {
    fe *pre = (fe *)malloc(sizeof(fe) * (size_t)(2 + 2 * n));
    for (int i = 0; i < n; i++) pt_to_affine(&pre[2 + 2 * i], &pre[3 + 2 * i], &pks[i]);
    pt acc, t;
    pt_identity(&acc);
    for (int i = 0; i < n; i++) {
        fe d;
        pre[0] = pre[2 + 2 * i];
        pre[1] = pre[3 + 2 * i];
        poseidon_hash_truncated(&d, NULL, pre, 2 + 2 * n);
        pt_mul(&t, &pks[i], &d);
        pt_add(&acc, &acc, &t);
    }
    free(pre);
    *agg = acc;
    return 0;
}
static int verify_aggregate_one(const uint8_t *pks32, int n, const uint8_t sig64[64], const uint8_t msg32[32],
                                uint8_t c_out[32], uint8_t agg_out[32]) {
    pt *pks = (pt *)malloc(sizeof(pt) * (size_t)(n > 0 ? n : 1));
    pt R, agg;
    fe u, m;
    if (c_out) memset(c_out, 0, 32);
    if (agg_out) memset(agg_out, 0, 32);
    int ok = 1;
    for (int i = 0; i < n; i++) ok &= pt_decode(&pks[i], pks32 + 32 * i);
    ok &= f_from_bytes(&FR, &u, sig64);
    ok &= pt_decode(&R, sig64 + 32);
    ok &= f_from_bytes(&FQ, &m, msg32);
    int st = ST_BYTES_ERROR;
    if (ok && aggregate_pts(&agg, pks, n) == 0) {
        if (agg_out) pt_encode(agg_out, &agg);
        st = verify_single_pts(&agg, &u, &R, &m, c_out);
    }
    free(pks);
    return st;
}

/* Typed inputs: a point as JubJubExtended coordinates (u, v, z, t1, t2), each a 32-byte little-endian Montgomery value.
 * is_valid() as verify() evaluates it on typed values (dusk-jubjub semantics). */
static int ext5_load(pt *p, int *valid, const uint8_t in[160]) {
    fe c[5];
    for (int k = 0; k < 5; k++) {
        memcpy(c[k].l, in + 32 * k, 32);
        if (ge256(c[k].l, JJO_Q)) return 0; /* not a reduced field element */
    }
    p->X = c[0]; p->Y = c[1]; p->Z = c[2];
    Q_MUL(&p->T, &c[3], &c[4]);
    /* is_on_curve: z != 0, affine equation, (u/z)(v/z) z == t1 t2 */
    int on = !f_is_zero(&c[2]);
    fe u, v, t;
    pt_to_affine(&u, &v, p);
    pt a;
    pt_from_affine(&a, &u, &v);
    on = on && pt_is_on_curve(&a);
    Q_MUL(&t, &u, &v); Q_MUL(&t, &t, &c[2]);
    on = on && f_eq(&t, &p->T);
    int ident = pt_is_identity(p);
    *valid = on && !ident && pt_is_torsion_free(&a);
    return 1;
}
static int verify_ext_one(int variant, const uint8_t *pts, const uint8_t u32[32], const uint8_t msg32[32], uint8_t c_out[32]) {
    const int slots = variant == 0 ? 2 : (variant == 1 ? 4 : 3);
    pt p[4];
    int valid[4], ok = 1, all_valid = 1;
    fe u, m;
    if (c_out) memset(c_out, 0, 32);
    for (int s = 0; s < slots; s++) ok &= ext5_load(&p[s], &valid[s], pts + 160 * s);
    ok &= f_from_bytes(&FR, &u, u32);
    ok &= f_from_bytes(&FQ, &m, msg32);
    if (!ok) return ST_BYTES_ERROR;
    for (int s = 0; s < slots; s++) all_valid &= valid[s];
    if (!all_valid) return ST_INVALID_POINT;
    fe in[10], c;
    pt a, b, s1, s2;
    if (variant == 0) {
        pt_to_affine(&in[0], &in[1], &p[1]); pt_to_affine(&in[2], &in[3], &p[0]); in[4] = m;
        poseidon_hash_truncated(&c, c_out, in, 5);
        pt_mul(&a, &G_PT, &u); pt_mul(&b, &p[0], &c); pt_add(&s1, &a, &b);
        return pt_eq(&s1, &p[1]) ? ST_OK : ST_INVALID_SIGNATURE;
    }
    if (variant == 1) {
        uint64_t tag[4] = {0x4a4a53434844424cULL, 0, 0, 0};
        f_from_raw(&FQ, &in[0], tag);
        pt_to_affine(&in[1], &in[2], &p[2]); pt_to_affine(&in[3], &in[4], &p[3]);
        pt_to_affine(&in[5], &in[6], &p[0]); pt_to_affine(&in[7], &in[8], &p[1]); in[9] = m;
        poseidon_hash_truncated(&c, c_out, in, 10);
        pt_mul(&a, &G_PT, &u); pt_mul(&b, &p[0], &c); pt_add(&s1, &a, &b);
        pt_mul(&a, &GN_PT, &u); pt_mul(&b, &p[1], &c); pt_add(&s2, &a, &b);
        return (pt_eq(&s1, &p[2]) && pt_eq(&s2, &p[3])) ? ST_OK : ST_INVALID_SIGNATURE;
    }
    pt_to_affine(&in[0], &in[1], &p[2]); pt_to_affine(&in[2], &in[3], &p[0]); pt_to_affine(&in[4], &in[5], &p[1]); in[6] = m;
    poseidon_hash_truncated(&c, c_out, in, 7);
    pt_mul(&a, &p[1], &u); pt_mul(&b, &p[0], &c); pt_add(&s1, &a, &b);
    return pt_eq(&s1, &p[2]) ? ST_OK : ST_INVALID_SIGNATURE;
}

/* ------------------------------------------------------------------ deterministic synthetic batches */
static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void item_rng(uint64_t *state, uint64_t seed, uint64_t index, uint64_t stream) {
    uint64_t s = seed ^ (index * 0xd1342543de82ef95ULL) ^ (stream << 56);
    splitmix64(&s);
    *state = s;
}
static void rand_wide(uint64_t *state, uint8_t out[64]) {
    for (int i = 0; i < 8; i++) {
        uint64_t w = splitmix64(state);
        memcpy(out + 8 * i, &w, 8);
    }
}
static void fr_as_fq(fe *o, const fe *a_fr) { /* src/nonce.rs:89-106: bytes of an Fr element read as Fq */
    uint64_t raw[4];
    f_to_raw(&FR, raw, a_fr);
    f_from_raw(&FQ, o, raw);
}
static void fq_small(fe *o, uint64_t x) {
    uint64_t raw[4] = {x, 0, 0, 0};
    f_from_raw(&FQ, o, raw);
}
/* u = r - c * sk */
static void sig_scalar(uint8_t out[32], const fe *r, const fe *c, const fe *sk) {
    fe t;
    f_mul(&FR, &t, c, sk);
    f_sub(&FR, &t, r, &t);
    f_to_bytes(&FR, out, &t);
}
/* SecretKey::sign (reference src/keys/secret.rs:174-194) with hedged_nonce (src/nonce.rs:32-44) */
static void gen_single(uint64_t seed, uint64_t idx, uint8_t pk32[32], uint8_t sig64[64], uint8_t msg32[32]) {
    uint64_t st;
    uint8_t w[64];
    fe sk, m, rnd, in[5], r, c;
    item_rng(&st, seed, idx, 1);
    rand_wide(&st, w); f_from_wide(&FR, &sk, w);
    rand_wide(&st, w); f_from_wide(&FQ, &m, w);
    rand_wide(&st, w); f_from_wide(&FR, &rnd, w);
    fr_as_fq(&in[0], &rnd); fr_as_fq(&in[1], &sk); fq_small(&in[2], 1); in[3] = m;
    poseidon_hash_truncated(&r, NULL, in, 4);
    pt PK, R;
    fb_mul(&PK, &FB_G, &sk);
    fb_mul(&R, &FB_G, &r);
    pt_to_affine(&in[0], &in[1], &R);
    pt_to_affine(&in[2], &in[3], &PK);
    in[4] = m;
    poseidon_hash_truncated(&c, NULL, in, 5);
    sig_scalar(sig64, &r, &c, &sk);
    pt_encode(sig64 + 32, &R);
    pt_encode(pk32, &PK);
    f_to_bytes(&FQ, msg32, &m);
}
/* SecretKey::sign_double (reference src/keys/secret/double.rs:57-85) */
static void gen_double(uint64_t seed, uint64_t idx, uint8_t pk64[64], uint8_t sig96[96], uint8_t msg32[32]) {
    uint64_t st;
    uint8_t w[64];
    fe sk, m, rnd, in[10], r, c;
    item_rng(&st, seed, idx, 2);
    rand_wide(&st, w); f_from_wide(&FR, &sk, w);
    rand_wide(&st, w); f_from_wide(&FQ, &m, w);
    rand_wide(&st, w); f_from_wide(&FR, &rnd, w);
    fr_as_fq(&in[0], &rnd); fr_as_fq(&in[1], &sk); fq_small(&in[2], 2); in[3] = m;
    poseidon_hash_truncated(&r, NULL, in, 4);
    pt PK, PKp, R, Rp;
    fb_mul(&PK, &FB_G, &sk); fb_mul(&PKp, &FB_GN, &sk);
    fb_mul(&R, &FB_G, &r); fb_mul(&Rp, &FB_GN, &r);
    fq_small(&in[0], 0x4a4a53434844424cULL);
    pt_to_affine(&in[1], &in[2], &R);
    pt_to_affine(&in[3], &in[4], &Rp);
    pt_to_affine(&in[5], &in[6], &PK);
    pt_to_affine(&in[7], &in[8], &PKp);
    in[9] = m;
    poseidon_hash_truncated(&c, NULL, in, 10);
    sig_scalar(sig96, &r, &c, &sk);
    pt_encode(sig96 + 32, &R);
    pt_encode(sig96 + 64, &Rp);
    pt_encode(pk64, &PK);
    pt_encode(pk64 + 32, &PKp);
    f_to_bytes(&FQ, msg32, &m);
}
/* SecretKeyVarGen::random + sign (reference src/keys/secret/var_gen.rs) */
static void gen_vargen(uint64_t seed, uint64_t idx, uint8_t pk64[64], uint8_t sig64[64], uint8_t msg32[32]) {
    uint64_t st;
    uint8_t w[64];
    fe sk, g, m, rnd, in[7], r, c, t;
    item_rng(&st, seed, idx, 3);
    rand_wide(&st, w); f_from_wide(&FR, &sk, w);
    rand_wide(&st, w); f_from_wide(&FR, &g, w);
    rand_wide(&st, w); f_from_wide(&FQ, &m, w);
    rand_wide(&st, w); f_from_wide(&FR, &rnd, w);
    pt GEN, PK, R;
    fb_mul(&GEN, &FB_G, &g);
    f_mul(&FR, &t, &sk, &g); fb_mul(&PK, &FB_G, &t);   /* sk * (g G) */
    fe gen_u, gen_v;
    pt_to_affine(&gen_u, &gen_v, &GEN);
    fr_as_fq(&in[0], &rnd); fr_as_fq(&in[1], &sk); in[2] = gen_u; in[3] = gen_v; in[4] = m;
    poseidon_hash_truncated(&r, NULL, in, 5);
    f_mul(&FR, &t, &r, &g); fb_mul(&R, &FB_G, &t);     /* r * (g G) */
    pt_to_affine(&in[0], &in[1], &R);
    pt_to_affine(&in[2], &in[3], &PK);
    in[4] = gen_u; in[5] = gen_v; in[6] = m;
    poseidon_hash_truncated(&c, NULL, in, 7);
    sig_scalar(sig64, &r, &c, &sk);
    pt_encode(sig64 + 32, &R);
    pt_encode(pk64, &PK);
    pt_encode(pk64 + 32, &GEN);
    f_to_bytes(&FQ, msg32, &m);
}
/* n_signers keys, the signature is a plain Schnorr signature under sum d_i sk_i (what a completed
 * SpeedyMuSig session yields; reference src/multisig.rs:416-429 for the key) */
static void gen_aggregate(uint64_t seed, uint64_t idx, int n, uint8_t *pks32, uint8_t sig64[64], uint8_t msg32[32]) {
    uint64_t st;
    uint8_t w[64];
    item_rng(&st, seed, idx, 4);
    fe *sk = (fe *)malloc(sizeof(fe) * (size_t)n);
    fe *pre = (fe *)malloc(sizeof(fe) * (size_t)(2 + 2 * n));
    for (int i = 0; i < n; i++) {
        pt P;
        rand_wide(&st, w); f_from_wide(&FR, &sk[i], w);
        fb_mul(&P, &FB_G, &sk[i]);
        pt_to_affine(&pre[2 + 2 * i], &pre[3 + 2 * i], &P);
        pt_encode(pks32 + 32 * i, &P);
    }
    fe agg_sk, d, t, m, r, c, in[5];
    f_zero(&agg_sk);
    for (int i = 0; i < n; i++) {
        pre[0] = pre[2 + 2 * i]; pre[1] = pre[3 + 2 * i];
        poseidon_hash_truncated(&d, NULL, pre, 2 + 2 * n);
        f_mul(&FR, &t, &d, &sk[i]);
        f_add(&FR, &agg_sk, &agg_sk, &t);
    }
    rand_wide(&st, w); f_from_wide(&FQ, &m, w);
    rand_wide(&st, w); f_from_wide(&FR, &r, w);
    pt PK, R;
    fb_mul(&PK, &FB_G, &agg_sk);
    fb_mul(&R, &FB_G, &r);
    pt_to_affine(&in[0], &in[1], &R);
    pt_to_affine(&in[2], &in[3], &PK);
    in[4] = m;
    poseidon_hash_truncated(&c, NULL, in, 5);
    sig_scalar(sig64, &r, &c, &agg_sk);
    pt_encode(sig64 + 32, &R);
    f_to_bytes(&FQ, msg32, &m);
    free(sk); free(pre);
}

/* ------------------------------------------------------------------ multisig: combine / verify_share */
enum { ST_INVALID_MULTISIG_TRANSCRIPT = 4, ST_INVALID_MULTISIG_SHARE = 5 };
typedef struct { fe *d; pt agg; fe agg_u, agg_v; fe a; pt RSa; fe c; } msig_coeffs;

/* multisig_common (reference src/multisig.rs:440-500); d must hold n elements (Fr Montgomery) */
static void msig_common(msig_coeffs *k, const pt *pks, const pt *Rs, const pt *Ss, int n, const fe *m) {
    fe *pre = (fe *)malloc(sizeof(fe) * (size_t)(3 + 4 * n + 2));
    for (int i = 0; i < n; i++) pt_to_affine(&pre[2 + 2 * i], &pre[3 + 2 * i], &pks[i]);
    pt acc, t;
    pt_identity(&acc);
    for (int i = 0; i < n; i++) {
        pre[0] = pre[2 + 2 * i];
        pre[1] = pre[3 + 2 * i];
        poseidon_hash_truncated(&k->d[i], NULL, pre, 2 + 2 * n);
        pt_mul(&t, &pks[i], &k->d[i]);
        pt_add(&acc, &acc, &t);
    }
    k->agg = acc;
    pt_to_affine(&k->agg_u, &k->agg_v, &acc);
    pre[0] = k->agg_u; pre[1] = k->agg_v; pre[2] = *m;
    for (int i = 0; i < n; i++) {
        pt_to_affine(&pre[3 + 4 * i], &pre[4 + 4 * i], &Rs[i]);
        pt_to_affine(&pre[5 + 4 * i], &pre[6 + 4 * i], &Ss[i]);
    }
    poseidon_hash_truncated(&k->a, NULL, pre, 3 + 4 * n);
    pt_identity(&acc);
    for (int i = 0; i < n; i++) {
        pt_add(&acc, &acc, &Rs[i]);
        pt_mul(&t, &Ss[i], &k->a);
        pt_add(&acc, &acc, &t);
    }
    k->RSa = acc;
    fe in[5];
    pt_to_affine(&in[0], &in[1], &acc);
    in[2] = k->agg_u; in[3] = k->agg_v; in[4] = *m;
    poseidon_hash_truncated(&k->c, NULL, in, 5);
    free(pre);
}
/* combine (reference src/multisig.rs:311-347) on wire encodings */
static int msig_combine_one(const uint8_t *pks32, const uint8_t *R32, const uint8_t *S32, const uint8_t *z32, int n, const uint8_t msg32[32],
                            uint8_t *share_ok, uint32_t *bad, uint8_t sig64[64]) {
    *bad = 0xffffffffu;
    memset(sig64, 0, 64);
    for (int i = 0; i < n; i++) share_ok[i] = 0;
    if (n <= 0) return ST_INVALID_MULTISIG_TRANSCRIPT;
    pt *P = (pt *)malloc(sizeof(pt) * 3 * (size_t)n);
    fe *z = (fe *)malloc(sizeof(fe) * 2 * (size_t)n), m;
    pt *pks = P, *Rs = P + n, *Ss = P + 2 * n;
    int ok = f_from_bytes(&FQ, &m, msg32);
    for (int i = 0; i < n; i++) {
        ok &= pt_decode(&pks[i], pks32 + 32 * i);
        ok &= pt_decode(&Rs[i], R32 + 32 * i);
        ok &= pt_decode(&Ss[i], S32 + 32 * i);
        ok &= f_from_bytes(&FR, &z[i], z32 + 32 * i);
    }
    int st = ST_BYTES_ERROR;
    if (ok) {
        msig_coeffs k;
        k.d = z + n;
        msig_common(&k, pks, Rs, Ss, n, &m);
        fe sum;
        f_zero(&sum);
        st = ST_OK;
        for (int i = 0; i < n; i++) {
            fe cd;
            pt a, b, lhs, rhs;
            f_mul(&FR, &cd, &k.c, &k.d[i]);
            pt_mul(&a, &G_PT, &z[i]); pt_mul(&b, &pks[i], &cd); pt_add(&lhs, &a, &b);
            pt_mul(&a, &Ss[i], &k.a); pt_add(&rhs, &Rs[i], &a);
            share_ok[i] = (uint8_t)pt_eq(&lhs, &rhs);
            if (!share_ok[i] && st == ST_OK) { st = ST_INVALID_MULTISIG_SHARE; *bad = (uint32_t)i; }
            f_add(&FR, &sum, &sum, &z[i]);
        }
        if (st == ST_OK) {
            f_to_bytes(&FR, sig64, &sum);
            pt_encode(sig64 + 32, &k.RSa);
        }
    }
    free(P); free(z);
    return st;
}
/* a complete, valid session: keys, commitments R_i = r_i G, S_i = s_i G and shares z_i = r_i + a s_i - c d_i sk_i
 * (sign_round_1 / sign_round_2, reference src/multisig.rs:172-253) */
static void gen_multisig(uint64_t seed, uint64_t idx, int n, uint8_t *pks32, uint8_t *R32, uint8_t *S32, uint8_t *z32, uint8_t msg32[32]) {
    uint64_t st;
    uint8_t w[64];
    item_rng(&st, seed, idx, 5);
    fe *sc = (fe *)malloc(sizeof(fe) * 4 * (size_t)n), m;
    pt *P = (pt *)malloc(sizeof(pt) * 3 * (size_t)n);
    fe *sk = sc, *r = sc + n, *s = sc + 2 * n;
    for (int i = 0; i < n; i++) {
        rand_wide(&st, w); f_from_wide(&FR, &sk[i], w);
        rand_wide(&st, w); f_from_wide(&FR, &r[i], w);
        rand_wide(&st, w); f_from_wide(&FR, &s[i], w);
        fb_mul(&P[i], &FB_G, &sk[i]); fb_mul(&P[n + i], &FB_G, &r[i]); fb_mul(&P[2 * n + i], &FB_G, &s[i]);
        pt_encode(pks32 + 32 * i, &P[i]); pt_encode(R32 + 32 * i, &P[n + i]); pt_encode(S32 + 32 * i, &P[2 * n + i]);
    }
    rand_wide(&st, w); f_from_wide(&FQ, &m, w);
    f_to_bytes(&FQ, msg32, &m);
    msig_coeffs k;
    k.d = sc + 3 * n;
    msig_common(&k, P, P + n, P + 2 * n, n, &m);
    for (int i = 0; i < n; i++) {
        fe t, zz;
        f_mul(&FR, &t, &k.a, &s[i]); f_add(&FR, &zz, &r[i], &t);
        f_mul(&FR, &t, &k.c, &k.d[i]); f_mul(&FR, &t, &t, &sk[i]); f_sub(&FR, &zz, &zz, &t);
        f_to_bytes(&FR, z32 + 32 * i, &zz);
    }
    free(sc); free(P);
}

/* ------------------------------------------------------------------ threaded batch drivers */
typedef struct {
    int kind; /* 0 verify single, 1 double, 2 vargen, 3 aggregate; 10.. generate */
    const uint8_t *pk, *sig, *msg;
    uint8_t *opk, *osig, *omsg;
    const uint32_t *offsets;
    const uint8_t *R, *S, *z;
    uint8_t *oR, *oS, *oz, *share_ok;
    uint32_t *bad;
    uint8_t *status, *c, *agg;
    uint64_t seed, first;
    size_t lo, hi;
} job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        uint8_t *c = j->c ? j->c + 32 * i : NULL;
        switch (j->kind) {
        case 0: j->status[i] = (uint8_t)verify_single_one(j->pk + 32 * i, j->sig + 64 * i, j->msg + 32 * i, c); break;
        case 1: j->status[i] = (uint8_t)verify_double_one(j->pk + 64 * i, j->sig + 96 * i, j->msg + 32 * i, c); break;
        case 2: j->status[i] = (uint8_t)verify_vargen_one(j->pk + 64 * i, j->sig + 64 * i, j->msg + 32 * i, c); break;
        case 3:
            j->status[i] = (uint8_t)verify_aggregate_one(j->pk + 32 * (size_t)j->offsets[i], (int)(j->offsets[i + 1] - j->offsets[i]),
                                                         j->sig + 64 * i, j->msg + 32 * i, c, j->agg ? j->agg + 32 * i : NULL);
            break;
        case 4: {
            size_t lo = j->offsets[i];
            int cnt = (int)(j->offsets[i + 1] - j->offsets[i]);
            j->status[i] = (uint8_t)msig_combine_one(j->pk + 32 * lo, j->R + 32 * lo, j->S + 32 * lo, j->z + 32 * lo, cnt, j->msg + 32 * i,
                                                     j->share_ok + lo, &j->bad[i], j->osig + 64 * i);
            break;
        }
        case 14: {
            size_t lo = j->offsets[i];
            gen_multisig(j->seed, j->first + i, (int)(j->offsets[i + 1] - j->offsets[i]), j->opk + 32 * lo, j->oR + 32 * lo, j->oS + 32 * lo,
                         j->oz + 32 * lo, j->omsg + 32 * i);
            break;
        }
        case 10: gen_single(j->seed, j->first + i, j->opk + 32 * i, j->osig + 64 * i, j->omsg + 32 * i); break;
        case 11: gen_double(j->seed, j->first + i, j->opk + 64 * i, j->osig + 96 * i, j->omsg + 32 * i); break;
        case 12: gen_vargen(j->seed, j->first + i, j->opk + 64 * i, j->osig + 64 * i, j->omsg + 32 * i); break;
        case 13:
            gen_aggregate(j->seed, j->first + i, (int)(j->offsets[i + 1] - j->offsets[i]), j->opk + 32 * (size_t)j->offsets[i],
                          j->osig + 64 * i, j->omsg + 32 * i);
            break;
        }
    }
    return NULL;
}
static void run(job *tmpl, size_t n, int threads) {
    init();
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job *jobs = (job *)malloc(sizeof(job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = *tmpl;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        if (threads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (threads > 1)
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------ exported C ABI (ctypes) */
#define EXPORT __attribute__((visibility("default")))

EXPORT void jjo_verify_single(const uint8_t *pk32, const uint8_t *sig64, const uint8_t *msg32, size_t n,
                              uint8_t *status, uint8_t *c32_or_null, int threads) {
    job j = {0}; j.kind = 0; j.pk = pk32; j.sig = sig64; j.msg = msg32; j.status = status; j.c = c32_or_null;
    run(&j, n, threads);
}
EXPORT void jjo_verify_double(const uint8_t *pk64, const uint8_t *sig96, const uint8_t *msg32, size_t n,
                              uint8_t *status, uint8_t *c32_or_null, int threads) {
    job j = {0}; j.kind = 1; j.pk = pk64; j.sig = sig96; j.msg = msg32; j.status = status; j.c = c32_or_null;
    run(&j, n, threads);
}
EXPORT void jjo_verify_vargen(const uint8_t *pk64, const uint8_t *sig64, const uint8_t *msg32, size_t n,
                              uint8_t *status, uint8_t *c32_or_null, int threads) {
    job j = {0}; j.kind = 2; j.pk = pk64; j.sig = sig64; j.msg = msg32; j.status = status; j.c = c32_or_null;
    run(&j, n, threads);
}
EXPORT void jjo_verify_aggregate(const uint8_t *pks32, const uint32_t *offsets, const uint8_t *sig64, const uint8_t *msg32,
                                 size_t n, uint8_t *status, uint8_t *c32_or_null, uint8_t *aggpk32_or_null, int threads) {
    job j = {0}; j.kind = 3; j.pk = pks32; j.offsets = offsets; j.sig = sig64; j.msg = msg32; j.status = status;
    j.c = c32_or_null; j.agg = aggpk32_or_null;
    run(&j, n, threads);
}
EXPORT void jjo_gen_single(uint64_t seed, uint64_t first, size_t n, uint8_t *pk32, uint8_t *sig64, uint8_t *msg32, int threads) {
    job j = {0}; j.kind = 10; j.seed = seed; j.first = first; j.opk = pk32; j.osig = sig64; j.omsg = msg32;
    run(&j, n, threads);
}
EXPORT void jjo_gen_double(uint64_t seed, uint64_t first, size_t n, uint8_t *pk64, uint8_t *sig96, uint8_t *msg32, int threads) {
    job j = {0}; j.kind = 11; j.seed = seed; j.first = first; j.opk = pk64; j.osig = sig96; j.omsg = msg32;
    run(&j, n, threads);
}
EXPORT void jjo_gen_vargen(uint64_t seed, uint64_t first, size_t n, uint8_t *pk64, uint8_t *sig64, uint8_t *msg32, int threads) {
    job j = {0}; j.kind = 12; j.seed = seed; j.first = first; j.opk = pk64; j.osig = sig64; j.omsg = msg32;
    run(&j, n, threads);
}
EXPORT void jjo_gen_aggregate(uint64_t seed, uint64_t first, size_t n, const uint32_t *offsets, uint8_t *pks32, uint8_t *sig64,
                              uint8_t *msg32, int threads) {
    job j = {0}; j.kind = 13; j.seed = seed; j.first = first; j.offsets = offsets; j.opk = pks32; j.osig = sig64; j.omsg = msg32;
    run(&j, n, threads);
}

EXPORT void jjo_verify_ext(int variant, const uint8_t *pts160, const uint8_t *u32, const uint8_t *msg32, size_t n, uint8_t *status,
                           uint8_t *c32_or_null) {
    init();
    const size_t slots = variant == 0 ? 2 : (variant == 1 ? 4 : 3);
    for (size_t i = 0; i < n; i++)
        status[i] = (uint8_t)verify_ext_one(variant, pts160 + 160 * slots * i, u32 + 32 * i, msg32 + 32 * i, c32_or_null ? c32_or_null + 32 * i : NULL);
}
/* compressed point -> JubJubExtended coordinates scaled by the Montgomery value z_mont (any non-zero field element),
 * with t1 = u z, t2 = v (so that t1 t2 / z = u v z as required); used to build typed test inputs */
EXPORT int jjo_point_to_ext(const uint8_t in32[32], const uint8_t z_mont32[32], uint8_t out160[160]) {
    init();
    pt p;
    if (!pt_decode(&p, in32)) return 0;
    fe z, c[5];
    memcpy(z.l, z_mont32, 32);
    if (ge256(z.l, JJO_Q)) return 0;
    Q_MUL(&c[0], &p.X, &z); Q_MUL(&c[1], &p.Y, &z); c[2] = z;
    c[3] = c[0]; c[4] = p.Y;
    for (int k = 0; k < 5; k++) memcpy(out160 + 32 * k, c[k].l, 32);
    return 1;
}

EXPORT void jjo_multisig_combine(const uint8_t *pks32, const uint8_t *R32, const uint8_t *S32, const uint8_t *z32, const uint32_t *offsets,
                                 const uint8_t *msg32, size_t n, uint8_t *share_ok, uint8_t *status, uint32_t *bad_index, uint8_t *sig64, int threads) {
    job j = {0}; j.kind = 4; j.pk = pks32; j.R = R32; j.S = S32; j.z = z32; j.offsets = offsets; j.msg = msg32; j.share_ok = share_ok;
    j.status = status; j.bad = bad_index; j.osig = sig64;
    run(&j, n, threads);
}
EXPORT void jjo_gen_multisig(uint64_t seed, uint64_t first, size_t n, const uint32_t *offsets, uint8_t *pks32, uint8_t *R32, uint8_t *S32,
                             uint8_t *z32, uint8_t *msg32, int threads) {
    job j = {0}; j.kind = 14; j.seed = seed; j.first = first; j.offsets = offsets; j.opk = pks32; j.oR = R32; j.oS = S32; j.oz = z32; j.omsg = msg32;
    run(&j, n, threads);
}

/* small helpers for building adversarial inputs and for unit parity checks */
EXPORT int jjo_point_decode(const uint8_t in32[32], uint8_t uv64[64]) { /* canonical affine (u, v) */
    init();
    pt p;
    if (!pt_decode(&p, in32)) return 0;
    f_to_bytes(&FQ, uv64, &p.X);
    f_to_bytes(&FQ, uv64 + 32, &p.Y);
    return 1;
}
EXPORT int jjo_point_from_uv(const uint8_t uv64[64], uint8_t out32[32]) { /* compress an arbitrary curve point */
    init();
    fe u, v;
    if (!f_from_bytes(&FQ, &u, uv64) || !f_from_bytes(&FQ, &v, uv64 + 32)) return 0;
    pt p;
    pt_from_affine(&p, &u, &v);
    if (!pt_is_on_curve(&p)) return 0;
    pt_encode(out32, &p);
    return 1;
}
EXPORT int jjo_point_add(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]) {
    init();
    pt a, b, s;
    if (!pt_decode(&a, a32) || !pt_decode(&b, b32)) return 0;
    pt_add(&s, &a, &b);
    pt_encode(out32, &s);
    return 1;
}
EXPORT int jjo_point_mul(const uint8_t p32[32], const uint8_t k32[32], uint8_t out32[32]) { /* k: any 256-bit LE integer */
    init();
    pt p, acc, tmp;
    if (!pt_decode(&p, p32)) return 0;
    pt_identity(&acc);
    for (int i = 255; i >= 0; i--) {
        pt_dbl(&acc, &acc);
        if ((k32[i >> 3] >> (i & 7)) & 1) { pt_add(&tmp, &acc, &p); acc = tmp; }
    }
    pt_encode(out32, &acc);
    return 1;
}
EXPORT int jjo_point_is_valid(const uint8_t p32[32]) { /* -1 decode failure, else is_valid() */
    init();
    pt p;
    if (!pt_decode(&p, p32)) return -1;
    return pt_is_valid(&p);
}
EXPORT void jjo_hades_permute(uint8_t state160[160]) { /* 5 canonical LE lanes in/out */
    init();
    fe s[5];
    for (int i = 0; i < 5; i++) {
        uint64_t raw[4];
        memcpy(raw, state160 + 32 * i, 32);
        f_from_raw(&FQ, &s[i], raw);
    }
    hades_permute(s);
    for (int i = 0; i < 5; i++) f_to_bytes(&FQ, state160 + 32 * i, &s[i]);
}
EXPORT int jjo_poseidon_hash(const uint8_t *in32, int n, int truncated, uint8_t out32[32]) {
    init();
    if (n < 1) return 0;
    fe *in = (fe *)malloc(sizeof(fe) * (size_t)n), h;
    int ok = 1;
    for (int i = 0; i < n; i++) ok &= f_from_bytes(&FQ, &in[i], in32 + 32 * i);
    if (ok) {
        poseidon_hash(&h, in, n);
        f_to_bytes(&FQ, out32, &h);
        if (truncated) out32[31] &= 0x03;
    }
    free(in);
    return ok;
}
EXPORT void jjo_fq_mul(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]) { /* canonical in/out */
    init();
    fe a, b;
    uint64_t ra[4], rb[4];
    memcpy(ra, a32, 32); memcpy(rb, b32, 32);
    f_from_raw(&FQ, &a, ra); f_from_raw(&FQ, &b, rb);
    Q_MUL(&a, &a, &b);
    f_to_bytes(&FQ, out32, &a);
}
