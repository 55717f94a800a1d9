/* oracle/c/ofield.h — 256-bit Montgomery prime-field arithmetic on 4x64-bit limbs.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/zkp_oracle.py header). PARITY UNPINNED.
 * Restates what ark-ff ^0.5 `Fp256<MontBackend<_, 4>>` computes [UPSTREAM]
 * (reference call sites: src/backend/snark.rs:194,203-208): R = 2^256, CIOS
 * multiplication, values kept fully reduced in [0, p).
 * One code path serves Fr and Fq; the modulus lives in an `fctx`.
 */
#ifndef ORA_FIELD_H
#define ORA_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } u256;

typedef struct {
    u256 p;        /* modulus */
    uint64_t inv;  /* -p^{-1} mod 2^64 */
    u256 one;      /* R mod p */
    u256 r2;       /* R^2 mod p */
    u256 pm2;      /* p - 2 (Fermat inverse exponent) */
} fctx;

static inline int u256_is_zero(const u256 *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int u256_eq(const u256 *a, const u256 *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int u256_cmp(const u256 *a, const u256 *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return -1;
    }
    return 0;
}
static inline uint64_t u256_add(u256 *o, const u256 *a, const u256 *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; o->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t u256_sub(u256 *o, const u256 *a, const u256 *b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - br;
        o->l[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}
static inline int u256_bit(const u256 *a, int i) { return (int)((a->l[i >> 6] >> (i & 63)) & 1); }

static inline void f_add(u256 *o, const u256 *a, const u256 *b, const fctx *c) {
    u256 t;
    uint64_t cy = u256_add(&t, a, b);
    if (cy || u256_cmp(&t, &c->p) >= 0) u256_sub(&t, &t, &c->p);
    *o = t;
}
static inline void f_sub(u256 *o, const u256 *a, const u256 *b, const fctx *c) {
    u256 t;
    if (u256_sub(&t, a, b)) u256_add(&t, &t, &c->p);
    *o = t;
}
static inline void f_neg(u256 *o, const u256 *a, const fctx *c) {
    if (u256_is_zero(a)) { *o = *a; return; }
    u256_sub(o, &c->p, a);
}
static inline void f_dbl(u256 *o, const u256 *a, const fctx *c) { f_add(o, a, a, c); }

/* CIOS Montgomery product a*b*R^{-1} mod p. */
static inline void f_mul(u256 *o, const u256 *a, const u256 *b, const fctx *c) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 x, cy = 0;
        uint64_t bi = b->l[i];
        for (int j = 0; j < 4; j++) {
            x = (u128)a->l[j] * bi + t[j] + cy;
            t[j] = (uint64_t)x; cy = x >> 64;
        }
        x = (u128)t[4] + cy; t[4] = (uint64_t)x; t[5] = (uint64_t)(x >> 64);
        uint64_t m = t[0] * c->inv;
        x = (u128)m * c->p.l[0] + t[0]; cy = x >> 64;
        for (int j = 1; j < 4; j++) {
            x = (u128)m * c->p.l[j] + t[j] + cy;
            t[j - 1] = (uint64_t)x; cy = x >> 64;
        }
        x = (u128)t[4] + cy; t[3] = (uint64_t)x; t[4] = t[5] + (uint64_t)(x >> 64);
    }
    u256 r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || u256_cmp(&r, &c->p) >= 0) u256_sub(&r, &r, &c->p);
    *o = r;
}
static inline void f_sqr(u256 *o, const u256 *a, const fctx *c) { f_mul(o, a, a, c); }
static inline void f_from_canon(u256 *o, const u256 *a, const fctx *c) { f_mul(o, a, &c->r2, c); }
static inline void f_to_canon(u256 *o, const u256 *a, const fctx *c) {
    u256 one = {{1, 0, 0, 0}};
    f_mul(o, a, &one, c);
}
static inline void f_pow(u256 *o, const u256 *a, const u256 *e, const fctx *c) {
    u256 r = c->one, b = *a;
    int top = 255;
    while (top >= 0 && !u256_bit(e, top)) top--;
    for (int i = top; i >= 0; i--) {
        f_sqr(&r, &r, c);
        if (u256_bit(e, i)) f_mul(&r, &r, &b, c);
    }
    *o = r;
}
static inline void f_inv(u256 *o, const u256 *a, const fctx *c) { f_pow(o, a, &c->pm2, c); }
static inline void f_from_u64(u256 *o, uint64_t v, const fctx *c) {
    u256 t = {{v, 0, 0, 0}};
    f_from_canon(o, &t, c);
}

static inline void fctx_init(fctx *c, const uint64_t p[4]) {
    memcpy(c->p.l, p, 32);
    uint64_t inv = 1;                               /* Newton: inv = p^{-1} mod 2^64 */
    for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;
    c->inv = (uint64_t)0 - inv;
    u256 two = {{2, 0, 0, 0}};
    u256_sub(&c->pm2, &c->p, &two);
    /* R mod p and R^2 mod p by repeated doubling of 1 (256 resp. 512 times) */
    u256 x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        uint64_t cy = u256_add(&x, &x, &x);
        if (cy || u256_cmp(&x, &c->p) >= 0) u256_sub(&x, &x, &c->p);
        if (i == 255) c->one = x;
    }
    c->r2 = x;
}

static inline void u256_from_le(u256 *o, const uint8_t *b) { memcpy(o->l, b, 32); }
static inline void u256_to_le(uint8_t *b, const u256 *a) { memcpy(b, a->l, 32); }

#endif
