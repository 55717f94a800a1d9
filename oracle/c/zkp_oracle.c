/* oracle/c/zkp_oracle.c — C restatement of the Groth16/BN254 prover path behind
 * libzkp's SNARK backend, for parity checks at sizes the Python oracle cannot
 * reach and as the timed CPU baseline ("port") in bench.py.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 * PARITY UNPINNED: /root/reference holds no golden vectors for this path and its
 * arithmetic sits in un-vendored arkworks ^0.5 crates (Cargo.toml:15-27) that
 * cannot be built here.  This file follows
 *   src/backend/snark.rs:182-211  (MiMC-5, constants)
 *   src/backend/snark.rs:232-291  (mimc_hash_circuit, EqualityCircuit)
 *   src/backend/snark.rs:505-585  (MembershipCircuit)
 *   src/backend/snark.rs:343-374, 405-452 (prove wrappers, 256-byte output)
 * and the [UPSTREAM] recipe of ark-groth16 / ark-poly / ark-ec / ark-serialize
 * listed in SURVEY.md §8a (a3-a16).  It is pinned against oracle/zkp_oracle.py
 * (independent big-int restatement) by tests/test_oracle_c.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>
#include "ofield.h"
#include "osha256.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ fields */
static fctx FR, FQ;
static const uint64_t FR_P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t FQ_P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static u256 FR_ROOT28;          /* 5^((r-1)/2^28), Montgomery */
static u256 FR_GEN;             /* 5, Montgomery */
static u256 MIMC_C[110];        /* Montgomery */
static int g_init_done = 0;

typedef struct { u256 c0, c1; } fq2;
static inline void fq2_add(fq2 *o, const fq2 *a, const fq2 *b) { f_add(&o->c0, &a->c0, &b->c0, &FQ); f_add(&o->c1, &a->c1, &b->c1, &FQ); }
static inline void fq2_sub(fq2 *o, const fq2 *a, const fq2 *b) { f_sub(&o->c0, &a->c0, &b->c0, &FQ); f_sub(&o->c1, &a->c1, &b->c1, &FQ); }
static inline void fq2_neg(fq2 *o, const fq2 *a) { f_neg(&o->c0, &a->c0, &FQ); f_neg(&o->c1, &a->c1, &FQ); }
static inline void fq2_mul(fq2 *o, const fq2 *a, const fq2 *b) {
    u256 v0, v1, s, t;
    f_mul(&v0, &a->c0, &b->c0, &FQ); f_mul(&v1, &a->c1, &b->c1, &FQ);
    f_add(&s, &a->c0, &a->c1, &FQ); f_add(&t, &b->c0, &b->c1, &FQ);
    f_mul(&s, &s, &t, &FQ); f_sub(&s, &s, &v0, &FQ); f_sub(&s, &s, &v1, &FQ);
    f_sub(&o->c0, &v0, &v1, &FQ); o->c1 = s;
}
static inline void fq2_sqr(fq2 *o, const fq2 *a) {
    u256 s, d, m;
    f_add(&s, &a->c0, &a->c1, &FQ); f_sub(&d, &a->c0, &a->c1, &FQ);
    f_mul(&m, &a->c0, &a->c1, &FQ);
    f_mul(&o->c0, &s, &d, &FQ); f_add(&o->c1, &m, &m, &FQ);
}
static inline void fq2_inv(fq2 *o, const fq2 *a) {
    u256 n, t;
    f_sqr(&n, &a->c0, &FQ); f_sqr(&t, &a->c1, &FQ); f_add(&n, &n, &t, &FQ);
    f_inv(&n, &n, &FQ);
    f_mul(&o->c0, &a->c0, &n, &FQ);
    f_mul(&t, &a->c1, &n, &FQ); f_neg(&o->c1, &t, &FQ);
}
static inline int fq2_is_zero(const fq2 *a) { return u256_is_zero(&a->c0) && u256_is_zero(&a->c1); }
static inline int fq2_eq(const fq2 *a, const fq2 *b) { return u256_eq(&a->c0, &b->c0) && u256_eq(&a->c1, &b->c1); }

/* --------------------------------------------------- curve instantiations */
#define FE u256
#define PFX(n) g1_##n
#define FE_ADD(o, a, b) f_add(o, a, b, &FQ)
#define FE_SUB(o, a, b) f_sub(o, a, b, &FQ)
#define FE_NEG(o, a) f_neg(o, a, &FQ)
#define FE_MUL(o, a, b) f_mul(o, a, b, &FQ)
#define FE_SQR(o, a) f_sqr(o, a, &FQ)
#define FE_INV(o, a) f_inv(o, a, &FQ)
#define FE_IS_ZERO(a) u256_is_zero(a)
#define FE_EQ(a, b) u256_eq(a, b)
#define FE_SET_ONE(a) (*(a) = FQ.one)
#define FE_SET_ZERO(a) memset(a, 0, sizeof(u256))
#define ORA_TMPL_FIRST
#include "ocurve_tmpl.h"
#undef ORA_TMPL_FIRST
#undef FE
#undef PFX
#undef FE_ADD
#undef FE_SUB
#undef FE_NEG
#undef FE_MUL
#undef FE_SQR
#undef FE_INV
#undef FE_IS_ZERO
#undef FE_EQ
#undef FE_SET_ONE
#undef FE_SET_ZERO

#define FE fq2
#define PFX(n) g2_##n
#define FE_ADD(o, a, b) fq2_add(o, a, b)
#define FE_SUB(o, a, b) fq2_sub(o, a, b)
#define FE_NEG(o, a) fq2_neg(o, a)
#define FE_MUL(o, a, b) fq2_mul(o, a, b)
#define FE_SQR(o, a) fq2_sqr(o, a)
#define FE_INV(o, a) fq2_inv(o, a)
#define FE_IS_ZERO(a) fq2_is_zero(a)
#define FE_EQ(a, b) fq2_eq(a, b)
#define FE_SET_ONE(a) do { (a)->c0 = FQ.one; memset(&(a)->c1, 0, sizeof(u256)); } while (0)
#define FE_SET_ZERO(a) memset(a, 0, sizeof(fq2))
#include "ocurve_tmpl.h"

/* --------------------------------------------------------------- init */
static void ora_init(void) {
    if (g_init_done) return;
#pragma omp critical(ora_init_lock)
    {
        if (!g_init_done) {
            fctx_init(&FR, FR_P);
            fctx_init(&FQ, FQ_P);
            f_from_u64(&FR_GEN, 5, &FR);
            u256 e = FR.p;                              /* (r-1) >> 28 */
            e.l[0] -= 1;
            for (int i = 0; i < 28; i++) {
                for (int k = 0; k < 3; k++) e.l[k] = (e.l[k] >> 1) | (e.l[k + 1] << 63);
                e.l[3] >>= 1;
            }
            f_pow(&FR_ROOT28, &FR_GEN, &e, &FR);
            /* snark.rs:186-198 */
            for (uint64_t i = 0; i < 110; i++) {
                uint8_t msg[23], dig[32];
                memcpy(msg, "libzkp_mimc_v1:", 15);
                memcpy(msg + 15, &i, 8);
                ora_sha256(msg, 23, dig);
                /* from_le_bytes_mod_order: 256-bit LE value mod r (value < 2^256 < 6r) */
                u256 v; u256_from_le(&v, dig);
                while (u256_cmp(&v, &FR.p) >= 0) u256_sub(&v, &v, &FR.p);
                f_from_canon(&MIMC_C[i], &v, &FR);
            }
            g_init_done = 1;
        }
    }
}

/* -------------------------------------------- ark-serialize (uncompressed) */
static int fq_read(u256 *o, const uint8_t *b, int flags, uint8_t *fl) {
    uint8_t tmp[32];
    memcpy(tmp, b, 32);
    if (flags) { *fl = tmp[31] & 0xC0; tmp[31] &= 0x3F; }
    u256 v; u256_from_le(&v, tmp);
    if (u256_cmp(&v, &FQ.p) >= 0) return -1;
    f_from_canon(o, &v, &FQ);
    return 0;
}
static int g1_read(g1_aff *p, const uint8_t *b) {
    uint8_t fl = 0;
    if (fq_read(&p->x, b, 0, NULL) || fq_read(&p->y, b + 32, 1, &fl)) return -1;
    p->inf = (fl & 0x40) != 0;
    if (p->inf) { memset(&p->x, 0, 32); memset(&p->y, 0, 32); }
    return 0;
}
static int g2_read(g2_aff *p, const uint8_t *b) {
    uint8_t fl = 0;
    if (fq_read(&p->x.c0, b, 0, NULL) || fq_read(&p->x.c1, b + 32, 0, NULL) ||
        fq_read(&p->y.c0, b + 64, 0, NULL) || fq_read(&p->y.c1, b + 96, 1, &fl)) return -1;
    p->inf = (fl & 0x40) != 0;
    if (p->inf) { memset(&p->x, 0, 64); memset(&p->y, 0, 64); }
    return 0;
}
/* "y is negative" <=> y > -y as canonical integers */
static int fq_is_neg(const u256 *y_canon) {
    u256 ny; u256_sub(&ny, &FQ.p, y_canon);
    if (u256_is_zero(y_canon)) return 0;
    return u256_cmp(y_canon, &ny) > 0;
}
static void g1_write(uint8_t *b, const g1_aff *p) {
    memset(b, 0, 64);
    if (p->inf) { b[63] = 0x40; return; }
    u256 x, y;
    f_to_canon(&x, &p->x, &FQ); f_to_canon(&y, &p->y, &FQ);
    u256_to_le(b, &x); u256_to_le(b + 32, &y);
    if (fq_is_neg(&y)) b[63] |= 0x80;
}
static void g2_write(uint8_t *b, const g2_aff *p) {
    memset(b, 0, 128);
    if (p->inf) { b[127] = 0x40; return; }
    u256 x0, x1, y0, y1;
    f_to_canon(&x0, &p->x.c0, &FQ); f_to_canon(&x1, &p->x.c1, &FQ);
    f_to_canon(&y0, &p->y.c0, &FQ); f_to_canon(&y1, &p->y.c1, &FQ);
    u256_to_le(b, &x0); u256_to_le(b + 32, &x1); u256_to_le(b + 64, &y0); u256_to_le(b + 96, &y1);
    /* Fq2 order: c1 first, then c0 */
    int neg;
    if (!u256_is_zero(&y1)) neg = fq_is_neg(&y1);
    else neg = fq_is_neg(&y0);
    if (neg) b[127] |= 0x80;
}

/* ------------------------------------------------------------------ MiMC */
static void mimc_native(u256 *out, uint64_t v) {       /* snark.rs:201-211 */
    u256 x, t, t2, t4;
    f_from_u64(&x, v, &FR);
    for (int i = 0; i < 110; i++) {
        f_add(&t, &x, &MIMC_C[i], &FR);
        f_sqr(&t2, &t, &FR); f_sqr(&t4, &t2, &FR); f_mul(&x, &t4, &t, &FR);
    }
    *out = x;
}
EXPORT int ora_mimc_hash(uint64_t v, uint8_t out[32]) {
    ora_init();
    u256 h, c; mimc_native(&h, v); f_to_canon(&c, &h, &FR); u256_to_le(out, &c);
    return 0;
}
EXPORT int ora_mimc_constant(uint32_t i, uint8_t out[32]) {
    ora_init();
    if (i >= 110) return -1;
    u256 c; f_to_canon(&c, &MIMC_C[i], &FR); u256_to_le(out, &c);
    return 0;
}

/* ------------------------------------------------------------------- NTT
 * ark-poly Radix2EvaluationDomain [UPSTREAM]: omega_n = ROOT28^(2^(28-log n)),
 * in-order in/out, inverse scales by n^{-1}; coset = distribute powers of g=5. */
static void fr_domain_root(u256 *w, unsigned log_n, int inverse) {
    u256 r = FR_ROOT28;
    for (unsigned i = log_n; i < 28; i++) f_sqr(&r, &r, &FR);
    if (inverse) f_inv(&r, &r, &FR);
    *w = r;
}
static void ntt_core(u256 *a, unsigned log_n, const u256 *root, int threads) {
    size_t n = (size_t)1 << log_n;
    int nt = threads > 0 ? threads : omp_get_max_threads();
    if (n < 4096) nt = 1;
    for (size_t i = 0; i < n; i++) {                     /* bit reversal */
        size_t j = 0;
        for (unsigned b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < j) { u256 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    u256 *tw = (u256 *)malloc(sizeof(u256) * (n / 2 ? n / 2 : 1));
    tw[0] = FR.one;
    for (size_t k = 1; k < n / 2; k++) f_mul(&tw[k], &tw[k - 1], root, &FR);
    for (unsigned s = 1; s <= log_n; s++) {
        size_t len = (size_t)1 << s, half = len >> 1, stride = n / len;
#pragma omp parallel for schedule(static) num_threads(nt)
        for (size_t bf = 0; bf < n / 2; bf++) {
            size_t blk = bf / half, k = bf % half;
            u256 *x = &a[blk * len + k], *y = x + half, v;
            f_mul(&v, y, &tw[k * stride], &FR);
            f_sub(y, x, &v, &FR);
            f_add(x, x, &v, &FR);
        }
    }
    free(tw);
}
static void distribute_powers(u256 *a, size_t n, const u256 *g, const u256 *scale, int threads) {
    int nt = threads > 0 ? threads : omp_get_max_threads();
    if (n < 4096) nt = 1;
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) continue;
        u256 e = {{lo, 0, 0, 0}}, p;
        f_pow(&p, g, &e, &FR);
        if (lo == 0) p = FR.one;
        if (scale) f_mul(&p, &p, scale, &FR);
        for (size_t i = lo; i < hi; i++) { f_mul(&a[i], &a[i], &p, &FR); f_mul(&p, &p, g, &FR); }
    }
}
static void fr_ntt(u256 *a, unsigned log_n, int inverse, int coset, int threads) {
    size_t n = (size_t)1 << log_n;
    u256 w; fr_domain_root(&w, log_n, inverse);
    if (!inverse) {
        if (coset) distribute_powers(a, n, &FR_GEN, NULL, threads);
        ntt_core(a, log_n, &w, threads);
    } else {
        ntt_core(a, log_n, &w, threads);
        u256 nn, ninv, ginv;
        f_from_u64(&nn, (uint64_t)n, &FR); f_inv(&ninv, &nn, &FR);
        if (coset) { f_inv(&ginv, &FR_GEN, &FR); distribute_powers(a, n, &ginv, &ninv, threads); }
        else {
            int nt = threads > 0 ? threads : omp_get_max_threads();
            if (n < 4096) nt = 1;
#pragma omp parallel for schedule(static) num_threads(nt)
            for (size_t i = 0; i < n; i++) f_mul(&a[i], &a[i], &ninv, &FR);
        }
    }
}
/* data: n x 32 B canonical LE, transformed in place. */
EXPORT int ora_ntt(uint8_t *data, uint32_t log_n, int inverse, int coset, int threads) {
    ora_init();
    if (log_n > 28) return -1;
    size_t n = (size_t)1 << log_n;
    u256 *a = (u256 *)malloc(sizeof(u256) * n);
    for (size_t i = 0; i < n; i++) {
        u256 v; u256_from_le(&v, data + 32 * i);
        if (u256_cmp(&v, &FR.p) >= 0) { free(a); return -2; }
        f_from_canon(&a[i], &v, &FR);
    }
    fr_ntt(a, log_n, inverse, coset, threads);
    for (size_t i = 0; i < n; i++) { u256 v; f_to_canon(&v, &a[i], &FR); u256_to_le(data + 32 * i, &v); }
    free(a);
    return 0;
}

/* ------------------------------------------------------------------- MSM */
EXPORT int ora_msm_g1(const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t out[64], int threads) {
    ora_init();
    g1_aff *b = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
    u256 *s = (u256 *)malloc(sizeof(u256) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        if (g1_read(&b[i], bases + 64 * i)) { free(b); free(s); return -1; }
        u256_from_le(&s[i], scalars + 32 * i);
        if (u256_cmp(&s[i], &FR.p) >= 0) { free(b); free(s); return -2; }
    }
    g1_jac r; g1_msm(&r, b, s, n, threads);
    g1_aff ra; g1_jac_to_aff(&ra, &r); g1_write(out, &ra);
    free(b); free(s);
    return 0;
}
EXPORT int ora_msm_g2(const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t out[128], int threads) {
    ora_init();
    g2_aff *b = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
    u256 *s = (u256 *)malloc(sizeof(u256) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        if (g2_read(&b[i], bases + 128 * i)) { free(b); free(s); return -1; }
        u256_from_le(&s[i], scalars + 32 * i);
        if (u256_cmp(&s[i], &FR.p) >= 0) { free(b); free(s); return -2; }
    }
    g2_jac r; g2_msm(&r, b, s, n, threads);
    g2_aff ra; g2_jac_to_aff(&ra, &r); g2_write(out, &ra);
    free(b); free(s);
    return 0;
}
/* k*G for the standard generators (fixtures: bases with known discrete logs). */
EXPORT int ora_g1_gen_mul(const uint8_t *scalars, size_t n, uint8_t *out /* n x 64 */) {
    ora_init();
    g1_aff g; f_from_u64(&g.x, 1, &FQ); f_from_u64(&g.y, 2, &FQ); g.inf = 0;
    g1_aff *tbl = g1_fixed_table(&g);
    g1_jac *j = (g1_jac *)malloc(sizeof(g1_jac) * (n ? n : 1));
    g1_aff *a = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { u256 k; u256_from_le(&k, scalars + 32 * i); g1_fixed_mul(&j[i], tbl, &k); }
    g1_jac_batch_to_aff(a, j, n);
    for (size_t i = 0; i < n; i++) g1_write(out + 64 * i, &a[i]);
    free(tbl); free(j); free(a);
    return 0;
}
static void g2_generator(g2_aff *g) {
    static const char *hex[4] = {
        "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
        "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2",
        "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa",
        "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"};
    u256 v[4];
    for (int k = 0; k < 4; k++) {
        memset(&v[k], 0, 32);
        for (int i = 0; i < 64; i++) {
            char ch = hex[k][i];
            uint64_t d = (ch <= '9') ? (uint64_t)(ch - '0') : (uint64_t)(ch - 'a' + 10);
            int bit = (63 - i) * 4;
            v[k].l[bit >> 6] |= d << (bit & 63);
        }
        f_from_canon(&v[k], &v[k], &FQ);
    }
    g->x.c0 = v[0]; g->x.c1 = v[1]; g->y.c0 = v[2]; g->y.c1 = v[3]; g->inf = 0;
}
EXPORT int ora_g2_gen_mul(const uint8_t *scalars, size_t n, uint8_t *out /* n x 128 */) {
    ora_init();
    g2_aff g; g2_generator(&g);
    g2_aff *tbl = g2_fixed_table(&g);
    g2_jac *j = (g2_jac *)malloc(sizeof(g2_jac) * (n ? n : 1));
    g2_aff *a = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { u256 k; u256_from_le(&k, scalars + 32 * i); g2_fixed_mul(&j[i], tbl, &k); }
    g2_jac_batch_to_aff(a, j, n);
    for (size_t i = 0; i < n; i++) g2_write(out + 128 * i, &a[i]);
    free(tbl); free(j); free(a);
    return 0;
}

/* ------------------------------------------------------------------ R1CS
 * Rows are sparse linear combinations over z = instance || witness
 * (ark-relations with OptimizationGoal::Constraints, all LCs inlined) [UPSTREAM].
 * While building, witness j is column WIT|j; finalize maps it to n_inst + j. */
#define WIT 0x80000000u
typedef struct { uint32_t col; u256 c; } term;
typedef struct { term *t; uint32_t n, cap; } lc_t;
static void lc_init(lc_t *l) { l->t = NULL; l->n = l->cap = 0; }
static void lc_free(lc_t *l) { free(l->t); lc_init(l); }
static void lc_add_term(lc_t *l, uint32_t col, const u256 *c) {
    for (uint32_t i = 0; i < l->n; i++)
        if (l->t[i].col == col) {
            f_add(&l->t[i].c, &l->t[i].c, c, &FR);
            if (u256_is_zero(&l->t[i].c)) { l->t[i] = l->t[l->n - 1]; l->n--; }
            return;
        }
    if (u256_is_zero(c)) return;
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 4; l->t = (term *)realloc(l->t, sizeof(term) * l->cap); }
    l->t[l->n].col = col; l->t[l->n].c = *c; l->n++;
}
static void lc_copy(lc_t *d, const lc_t *s) {
    lc_init(d);
    if (s->n) { d->t = (term *)malloc(sizeof(term) * s->n); memcpy(d->t, s->t, sizeof(term) * s->n); d->n = d->cap = s->n; }
}
static void lc_acc(lc_t *d, const lc_t *s, int negate) {
    for (uint32_t i = 0; i < s->n; i++) {
        u256 c = s->t[i].c;
        if (negate) f_neg(&c, &c, &FR);
        lc_add_term(d, s->t[i].col, &c);
    }
}
typedef struct { u256 val; lc_t lc; } fpvar;

typedef struct { uint32_t *rowptr; uint32_t *col; u256 *val; size_t nnz, cap; uint32_t rows, rcap; } csr;
typedef struct {
    csr M[3];
    u256 *inst, *wit;
    uint32_t n_inst, n_wit, cap_inst, cap_wit;
} cs_t;
static void csr_push_row(csr *m, const lc_t *l) {
    if (m->rows + 2 > m->rcap) { m->rcap = m->rcap ? m->rcap * 2 : 1024; m->rowptr = (uint32_t *)realloc(m->rowptr, sizeof(uint32_t) * m->rcap); }
    if (m->rows == 0) m->rowptr[0] = 0;
    if (m->nnz + l->n > m->cap) {
        m->cap = (m->cap ? m->cap * 2 : 4096) + l->n;
        m->col = (uint32_t *)realloc(m->col, sizeof(uint32_t) * m->cap);
        m->val = (u256 *)realloc(m->val, sizeof(u256) * m->cap);
    }
    for (uint32_t i = 0; i < l->n; i++) { m->col[m->nnz] = l->t[i].col; m->val[m->nnz] = l->t[i].c; m->nnz++; }
    m->rows++;
    m->rowptr[m->rows] = (uint32_t)m->nnz;
}
static void cs_init(cs_t *cs) {
    memset(cs, 0, sizeof(*cs));
    cs->cap_inst = 16; cs->inst = (u256 *)malloc(sizeof(u256) * 16);
    cs->cap_wit = 1024; cs->wit = (u256 *)malloc(sizeof(u256) * 1024);
    cs->inst[0] = FR.one; cs->n_inst = 1;
}
static void cs_new_input(cs_t *cs, fpvar *v, const u256 *val) {
    if (cs->n_inst == cs->cap_inst) { cs->cap_inst *= 2; cs->inst = (u256 *)realloc(cs->inst, sizeof(u256) * cs->cap_inst); }
    cs->inst[cs->n_inst] = *val;
    v->val = *val; lc_init(&v->lc); lc_add_term(&v->lc, cs->n_inst, &FR.one);
    cs->n_inst++;
}
static void cs_new_witness(cs_t *cs, fpvar *v, const u256 *val) {
    if (cs->n_wit == cs->cap_wit) { cs->cap_wit *= 2; cs->wit = (u256 *)realloc(cs->wit, sizeof(u256) * cs->cap_wit); }
    cs->wit[cs->n_wit] = *val;
    v->val = *val; lc_init(&v->lc); lc_add_term(&v->lc, WIT | cs->n_wit, &FR.one);
    cs->n_wit++;
}
static void cs_enforce(cs_t *cs, const lc_t *a, const lc_t *b, const lc_t *c) {
    csr_push_row(&cs->M[0], a); csr_push_row(&cs->M[1], b); csr_push_row(&cs->M[2], c);
}
/* AllocatedFp::mul: p = new witness(x*y); enforce x * y = p */
static void cs_mul(cs_t *cs, fpvar *p, const fpvar *x, const fpvar *y) {
    u256 v; f_mul(&v, &x->val, &y->val, &FR);
    cs_new_witness(cs, p, &v);
    cs_enforce(cs, &x->lc, &y->lc, &p->lc);
}
/* conditional_enforce_equal(TRUE): (x - y) * 1 = 0 */
static void cs_enforce_equal(cs_t *cs, const lc_t *x, const lc_t *y) {
    lc_t a, one, zero;
    lc_copy(&a, x); lc_acc(&a, y, 1);
    lc_init(&one); lc_add_term(&one, 0, &FR.one); lc_init(&zero);
    cs_enforce(cs, &a, &one, &zero);
    lc_free(&a); lc_free(&one);
}
/* AllocatedBool::new_variable: (1 - a) * a = 0 */
static void cs_new_bool(cs_t *cs, fpvar *v, int bit, int is_input) {
    u256 val; f_from_u64(&val, (uint64_t)(bit != 0), &FR);
    if (is_input) cs_new_input(cs, v, &val); else cs_new_witness(cs, v, &val);
    lc_t a, zero;
    lc_init(&a); lc_add_term(&a, 0, &FR.one); lc_acc(&a, &v->lc, 1); lc_init(&zero);
    cs_enforce(cs, &a, &v->lc, &zero);
    lc_free(&a);
}
/* snark.rs:232-247; rounds > 110 cycles the constants (synthetic config 4). */
static void cs_mimc(cs_t *cs, fpvar *x, uint32_t rounds) {
    for (uint32_t i = 0; i < rounds; i++) {
        fpvar t, t2, t4, x5;
        f_add(&t.val, &x->val, &MIMC_C[i % 110], &FR);
        lc_copy(&t.lc, &x->lc); lc_add_term(&t.lc, 0, &MIMC_C[i % 110]);
        cs_mul(cs, &t2, &t, &t);
        cs_mul(cs, &t4, &t2, &t2);
        cs_mul(cs, &x5, &t4, &t);
        lc_free(&t.lc); lc_free(&t2.lc); lc_free(&t4.lc); lc_free(&x->lc);
        *x = x5;
    }
}

typedef struct {
    uint32_t kind;              /* 0 = equality / MiMC chain, 1 = membership */
    uint32_t param;             /* rounds (kind 0) or slots (kind 1) */
    uint32_t m, n_inst, n_wit, log_n;
    csr M[3];                   /* final columns, rows sorted by column */
} ora_circuit;

static int term_cmp(const void *a, const void *b) {
    uint32_t x = ((const term *)a)->col, y = ((const term *)b)->col;
    return x < y ? -1 : x > y;
}
static void circuit_finalize(ora_circuit *c, cs_t *cs) {
    c->m = cs->M[0].rows; c->n_inst = cs->n_inst; c->n_wit = cs->n_wit;
    unsigned lg = 0;
    while (((uint64_t)1 << lg) < (uint64_t)c->m + c->n_inst) lg++;
    c->log_n = lg;
    for (int k = 0; k < 3; k++) {
        csr *m = &cs->M[k];
        for (size_t i = 0; i < m->nnz; i++) if (m->col[i] & WIT) m->col[i] = c->n_inst + (m->col[i] & ~WIT);
        for (uint32_t r = 0; r < m->rows; r++) {
            uint32_t lo = m->rowptr[r], hi = m->rowptr[r + 1];
            if (hi - lo > 1) {
                term *tmp = (term *)malloc(sizeof(term) * (hi - lo));
                for (uint32_t i = lo; i < hi; i++) { tmp[i - lo].col = m->col[i]; tmp[i - lo].c = m->val[i]; }
                qsort(tmp, hi - lo, sizeof(term), term_cmp);
                for (uint32_t i = lo; i < hi; i++) { m->col[i] = tmp[i - lo].col; m->val[i] = tmp[i - lo].c; }
                free(tmp);
            }
        }
        c->M[k] = *m;
    }
}

/* snark.rs:263-290 (rounds = 110), or the MiMC-chain variant for config 4. */
static void synth_equality(cs_t *cs, uint32_t rounds, uint64_t a, uint64_t b, const u256 *commitment) {
    fpvar av, bv, cv;
    u256 t;
    f_from_u64(&t, a, &FR); cs_new_witness(cs, &av, &t);
    f_from_u64(&t, b, &FR); cs_new_witness(cs, &bv, &t);
    cs_enforce_equal(cs, &av.lc, &bv.lc);
    cs_mimc(cs, &av, rounds);
    cs_new_input(cs, &cv, commitment);
    cs_enforce_equal(cs, &av.lc, &cv.lc);
    lc_free(&av.lc); lc_free(&bv.lc); lc_free(&cv.lc);
}
/* snark.rs:515-584 with `slots` set positions (the reference fixes 64). */
static void synth_membership(cs_t *cs, uint32_t slots, uint64_t value, const uint8_t *sel, const uint64_t *set_values,
                             const uint8_t *is_real, const u256 *commitment) {
    fpvar vv, hv, cv;
    u256 t;
    f_from_u64(&t, value, &FR); cs_new_witness(cs, &vv, &t);
    hv.val = vv.val; lc_copy(&hv.lc, &vv.lc);
    cs_mimc(cs, &hv, 110);
    cs_new_input(cs, &cv, commitment);
    cs_enforce_equal(cs, &hv.lc, &cv.lc);
    fpvar *setv = (fpvar *)malloc(sizeof(fpvar) * slots), *real = (fpvar *)malloc(sizeof(fpvar) * slots),
          *sels = (fpvar *)malloc(sizeof(fpvar) * slots);
    for (uint32_t i = 0; i < slots; i++) { f_from_u64(&t, set_values[i], &FR); cs_new_input(cs, &setv[i], &t); }
    for (uint32_t i = 0; i < slots; i++) cs_new_bool(cs, &real[i], is_real[i], 1);
    for (uint32_t i = 0; i < slots; i++) cs_new_bool(cs, &sels[i], sel[i], 0);
    lc_t sum, zero, one;
    lc_init(&sum); lc_init(&zero); lc_init(&one); lc_add_term(&one, 0, &FR.one);
    for (uint32_t i = 0; i < slots; i++) {
        lc_acc(&sum, &sels[i].lc, 0);
        fpvar om, prod;
        f_sub(&om.val, &FR.one, &real[i].val, &FR);
        lc_copy(&om.lc, &one); lc_acc(&om.lc, &real[i].lc, 1);
        cs_mul(cs, &prod, &sels[i], &om);
        cs_enforce_equal(cs, &zero, &prod.lc);            /* (const 0 - prod) * 1 = 0 */
        lc_free(&om.lc); lc_free(&prod.lc);
    }
    cs_enforce_equal(cs, &one, &sum);                     /* (const 1 - sum) * 1 = 0 */
    lc_t acc; lc_init(&acc);
    for (uint32_t i = 0; i < slots; i++) {
        fpvar diff, p;
        f_sub(&diff.val, &vv.val, &setv[i].val, &FR);
        lc_copy(&diff.lc, &vv.lc); lc_acc(&diff.lc, &setv[i].lc, 1);
        cs_mul(cs, &p, &sels[i], &diff);
        lc_acc(&acc, &p.lc, 0);
        lc_free(&diff.lc); lc_free(&p.lc);
    }
    cs_enforce_equal(cs, &zero, &acc);                    /* Var == Constant(0): ark-r1cs-std enforces (const - var) * 1 = 0 */
    for (uint32_t i = 0; i < slots; i++) { lc_free(&setv[i].lc); lc_free(&real[i].lc); lc_free(&sels[i].lc); }
    free(setv); free(real); free(sels);
    lc_free(&sum); lc_free(&one); lc_free(&acc); lc_free(&vv.lc); lc_free(&hv.lc); lc_free(&cv.lc);
}

EXPORT ora_circuit *ora_circuit_equality(uint32_t rounds) {
    ora_init();
    cs_t cs; cs_init(&cs);
    u256 zero; memset(&zero, 0, 32);
    synth_equality(&cs, rounds, 0, 0, &zero);             /* dummy assignment, snark.rs:332-336 */
    ora_circuit *c = (ora_circuit *)calloc(1, sizeof(*c));
    c->kind = 0; c->param = rounds;
    circuit_finalize(c, &cs);
    free(cs.inst); free(cs.wit);
    return c;
}
EXPORT ora_circuit *ora_circuit_membership(uint32_t slots) {
    ora_init();
    cs_t cs; cs_init(&cs);
    u256 zero; memset(&zero, 0, 32);
    uint8_t *f = (uint8_t *)calloc(slots, 1);
    uint64_t *sv = (uint64_t *)calloc(slots, 8);
    synth_membership(&cs, slots, 0, f, sv, f, &zero);     /* dummy, snark.rs:311-317 */
    ora_circuit *c = (ora_circuit *)calloc(1, sizeof(*c));
    c->kind = 1; c->param = slots;
    circuit_finalize(c, &cs);
    free(cs.inst); free(cs.wit); free(f); free(sv);
    return c;
}
EXPORT void ora_circuit_free(ora_circuit *c) {
    if (!c) return;
    for (int k = 0; k < 3; k++) { free(c->M[k].rowptr); free(c->M[k].col); free(c->M[k].val); }
    free(c);
}
/* shape[0..5) = m, n_inst, n_wit, log_n, nnzA, nnzB, nnzC */
EXPORT void ora_circuit_shape(const ora_circuit *c, uint64_t shape[7]) {
    shape[0] = c->m; shape[1] = c->n_inst; shape[2] = c->n_wit; shape[3] = c->log_n;
    shape[4] = c->M[0].nnz; shape[5] = c->M[1].nnz; shape[6] = c->M[2].nnz;
}
/* Export one matrix as CSR: rowptr[m+1], col[nnz], val[nnz x 32 B canonical LE]. */
EXPORT void ora_circuit_matrix(const ora_circuit *c, int which, uint32_t *rowptr, uint32_t *col, uint8_t *val) {
    const csr *m = &c->M[which];
    memcpy(rowptr, m->rowptr, sizeof(uint32_t) * (c->m + 1));
    memcpy(col, m->col, sizeof(uint32_t) * m->nnz);
    for (size_t i = 0; i < m->nnz; i++) { u256 v; f_to_canon(&v, &m->val[i], &FR); u256_to_le(val + 32 * i, &v); }
}

/* Full assignment z = instance || witness by running the circuit's own synthesis
 * (what the reference does on every prove, snark.rs:353-364). z_out: n_vars x 32 B. */
static int circuit_assign(const ora_circuit *c, u256 *z /* Montgomery */, uint64_t a, uint64_t b,
                          const uint64_t *set, uint32_t set_len, const u256 *commitment) {
    cs_t cs; cs_init(&cs);
    if (c->kind == 0) synth_equality(&cs, c->param, a, b, commitment);
    else {
        uint32_t slots = c->param;
        if (set_len == 0 || set_len > slots) { free(cs.inst); free(cs.wit); return -1; }   /* snark.rs:406 */
        uint32_t pos = set_len;
        for (uint32_t i = 0; i < set_len; i++) if (set[i] == a) { pos = i; break; }     /* snark.rs:415 */
        if (pos == set_len) { free(cs.inst); free(cs.wit); return -2; }
        uint8_t *sel = (uint8_t *)calloc(slots, 1), *real = (uint8_t *)calloc(slots, 1);
        uint64_t *sv = (uint64_t *)calloc(slots, 8);
        for (uint32_t i = 0; i < set_len; i++) { sv[i] = set[i]; real[i] = 1; }
        sel[pos] = 1;
        synth_membership(&cs, slots, a, sel, sv, real, commitment);
        free(sel); free(real); free(sv);
    }
    memcpy(z, cs.inst, sizeof(u256) * cs.n_inst);
    memcpy(z + cs.n_inst, cs.wit, sizeof(u256) * cs.n_wit);
    for (int k = 0; k < 3; k++) { free(cs.M[k].rowptr); free(cs.M[k].col); free(cs.M[k].val); }
    free(cs.inst); free(cs.wit);
    return 0;
}
EXPORT int ora_assign(const ora_circuit *c, uint64_t a, uint64_t b, const uint64_t *set, uint32_t set_len,
                      const uint8_t commitment[32], uint8_t *z_out) {
    ora_init();
    u256 cm; u256_from_le(&cm, commitment);
    if (u256_cmp(&cm, &FR.p) >= 0) return -3;
    f_from_canon(&cm, &cm, &FR);
    size_t nv = (size_t)c->n_inst + c->n_wit;
    u256 *z = (u256 *)malloc(sizeof(u256) * nv);
    int rc = circuit_assign(c, z, a, b, set, set_len, &cm);
    if (rc == 0) for (size_t i = 0; i < nv; i++) { u256 v; f_to_canon(&v, &z[i], &FR); u256_to_le(z_out + 32 * i, &v); }
    free(z);
    return rc;
}

/* ----------------------------------------------------------- witness map
 * LibsnarkReduction::witness_map_from_matrices [UPSTREAM]. */
static void spmv_row(u256 *o, const csr *m, uint32_t r, const u256 *z) {
    u256 acc, t; memset(&acc, 0, 32);
    for (uint32_t i = m->rowptr[r]; i < m->rowptr[r + 1]; i++) { f_mul(&t, &m->val[i], &z[m->col[i]], &FR); f_add(&acc, &acc, &t, &FR); }
    *o = acc;
}
static void witness_map(const ora_circuit *c, const u256 *z, u256 *h /* n */, int threads) {
    size_t n = (size_t)1 << c->log_n;
    u256 *a = (u256 *)calloc(n, sizeof(u256)), *b = (u256 *)calloc(n, sizeof(u256)), *cc = (u256 *)calloc(n, sizeof(u256));
    int nt = threads > 0 ? threads : omp_get_max_threads();
    if (n < 4096) nt = 1;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (uint32_t i = 0; i < c->m; i++) { spmv_row(&a[i], &c->M[0], i, z); spmv_row(&b[i], &c->M[1], i, z); spmv_row(&cc[i], &c->M[2], i, z); }
    for (uint32_t j = 0; j < c->n_inst; j++) a[c->m + j] = z[j];
    fr_ntt(a, c->log_n, 1, 0, threads); fr_ntt(b, c->log_n, 1, 0, threads);
    fr_ntt(a, c->log_n, 0, 1, threads); fr_ntt(b, c->log_n, 0, 1, threads);
    fr_ntt(cc, c->log_n, 1, 0, threads); fr_ntt(cc, c->log_n, 0, 1, threads);
    u256 zinv, e = {{n, 0, 0, 0}};
    f_pow(&zinv, &FR_GEN, &e, &FR); f_sub(&zinv, &zinv, &FR.one, &FR); f_inv(&zinv, &zinv, &FR);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t i = 0; i < n; i++) {
        u256 t; f_mul(&t, &a[i], &b[i], &FR); f_sub(&t, &t, &cc[i], &FR); f_mul(&h[i], &t, &zinv, &FR);
    }
    fr_ntt(h, c->log_n, 1, 1, threads);
    free(a); free(b); free(cc);
}
EXPORT int ora_witness_map(const ora_circuit *c, const uint8_t *z_in, uint8_t *h_out, int threads) {
    ora_init();
    size_t nv = (size_t)c->n_inst + c->n_wit, n = (size_t)1 << c->log_n;
    u256 *z = (u256 *)malloc(sizeof(u256) * nv), *h = (u256 *)malloc(sizeof(u256) * n);
    for (size_t i = 0; i < nv; i++) { u256 v; u256_from_le(&v, z_in + 32 * i); f_from_canon(&z[i], &v, &FR); }
    witness_map(c, z, h, threads);
    for (size_t i = 0; i < n; i++) { u256 v; f_to_canon(&v, &h[i], &FR); u256_to_le(h_out + 32 * i, &v); }
    free(z); free(h);
    return 0;
}

/* ------------------------------------------------------------ proving key */
typedef struct {
    g1_aff alpha_g1, beta_g1, delta_g1;
    g2_aff beta_g2, gamma_g2, delta_g2;
    g1_aff *gamma_abc; size_t n_gamma_abc;
    g1_aff *a_query, *b_g1_query, *h_query, *l_query;
    g2_aff *b_g2_query;
    size_t n_a, n_b1, n_b2, n_h, n_l;
} ora_pk;

static int rd_vec_g1(const uint8_t **p, const uint8_t *end, g1_aff **out, size_t *n) {
    if (*p + 8 > end) return -1;
    uint64_t len; memcpy(&len, *p, 8); *p += 8;
    if ((uint64_t)(end - *p) / 64 < len) return -1;
    *out = (g1_aff *)malloc(sizeof(g1_aff) * (len ? len : 1)); *n = len;
    for (uint64_t i = 0; i < len; i++, *p += 64) if (g1_read(&(*out)[i], *p)) return -2;
    return 0;
}
static int rd_vec_g2(const uint8_t **p, const uint8_t *end, g2_aff **out, size_t *n) {
    if (*p + 8 > end) return -1;
    uint64_t len; memcpy(&len, *p, 8); *p += 8;
    if ((uint64_t)(end - *p) / 128 < len) return -1;
    *out = (g2_aff *)malloc(sizeof(g2_aff) * (len ? len : 1)); *n = len;
    for (uint64_t i = 0; i < len; i++, *p += 128) if (g2_read(&(*out)[i], *p)) return -2;
    return 0;
}
EXPORT void ora_pk_free(ora_pk *pk) {
    if (!pk) return;
    free(pk->gamma_abc); free(pk->a_query); free(pk->b_g1_query); free(pk->h_query); free(pk->l_query); free(pk->b_g2_query);
    free(pk);
}
/* ProvingKey<Bn254>::deserialize_uncompressed layout (snark.rs:64) without the
 * curve/subgroup validation pass. */
EXPORT ora_pk *ora_pk_parse(const uint8_t *b, size_t len) {
    ora_init();
    ora_pk *pk = (ora_pk *)calloc(1, sizeof(*pk));
    const uint8_t *p = b, *end = b + len;
    int bad = 0;
    if (len < 64 + 3 * 128) { free(pk); return NULL; }
    bad |= g1_read(&pk->alpha_g1, p); p += 64;
    bad |= g2_read(&pk->beta_g2, p); p += 128;
    bad |= g2_read(&pk->gamma_g2, p); p += 128;
    bad |= g2_read(&pk->delta_g2, p); p += 128;
    bad |= rd_vec_g1(&p, end, &pk->gamma_abc, &pk->n_gamma_abc);
    if (!bad && p + 128 <= end) { bad |= g1_read(&pk->beta_g1, p); p += 64; bad |= g1_read(&pk->delta_g1, p); p += 64; } else bad = 1;
    if (!bad) bad |= rd_vec_g1(&p, end, &pk->a_query, &pk->n_a);
    if (!bad) bad |= rd_vec_g1(&p, end, &pk->b_g1_query, &pk->n_b1);
    if (!bad) bad |= rd_vec_g2(&p, end, &pk->b_g2_query, &pk->n_b2);
    if (!bad) bad |= rd_vec_g1(&p, end, &pk->h_query, &pk->n_h);
    if (!bad) bad |= rd_vec_g1(&p, end, &pk->l_query, &pk->n_l);
    if (bad || p != end) { ora_pk_free(pk); return NULL; }
    return pk;
}

/* ------------------------------------------------------------------ setup
 * generate_parameters_with_qap [UPSTREAM] with explicit toxic waste
 * (alpha, beta, gamma, delta, tau: 5 x 32 B canonical LE) and the standard
 * generators.  Writes pk and vk in ark-serialize uncompressed layout. */
static void wr_u64(uint8_t **p, uint64_t v) { memcpy(*p, &v, 8); *p += 8; }
EXPORT size_t ora_pk_size(const ora_circuit *c) {
    size_t nv = (size_t)c->n_inst + c->n_wit, n = (size_t)1 << c->log_n;
    return 64 + 3 * 128 + 8 + 64 * (size_t)c->n_inst + 128 + (8 + 64 * nv) * 2 + 8 + 128 * nv + 8 + 64 * (n - 1) + 8 + 64 * (size_t)c->n_wit;
}
EXPORT size_t ora_vk_size(const ora_circuit *c) { return 64 + 3 * 128 + 8 + 64 * (size_t)c->n_inst; }

EXPORT int ora_setup(const ora_circuit *c, const uint8_t trapdoor[160], uint8_t *pk_out, uint8_t *vk_out) {
    ora_init();
    u256 td[5];
    for (int i = 0; i < 5; i++) { u256 v; u256_from_le(&v, trapdoor + 32 * i); if (u256_cmp(&v, &FR.p) >= 0) return -1; f_from_canon(&td[i], &v, &FR); }
    const u256 *alpha = &td[0], *beta = &td[1], *gamma = &td[2], *delta = &td[3], *tau = &td[4];
    size_t nv = (size_t)c->n_inst + c->n_wit, n = (size_t)1 << c->log_n, m = c->m, ni = c->n_inst;
    /* evaluate_all_lagrange_coefficients(tau): u_i = Z(tau)/n * w^i / (tau - w^i) */
    u256 w, zt, e = {{n, 0, 0, 0}}, nn, ninv;
    fr_domain_root(&w, c->log_n, 0);
    f_pow(&zt, tau, &e, &FR); f_sub(&zt, &zt, &FR.one, &FR);
    if (u256_is_zero(&zt)) return -2;                     /* tau inside the domain: resample */
    f_from_u64(&nn, n, &FR); f_inv(&ninv, &nn, &FR);
    u256 *u = (u256 *)malloc(sizeof(u256) * n), *den = (u256 *)malloc(sizeof(u256) * n), *pre = (u256 *)malloc(sizeof(u256) * n);
    u256 wi = FR.one, acc = FR.one, k0;
    f_mul(&k0, &zt, &ninv, &FR);
    for (size_t i = 0; i < n; i++) {
        f_sub(&den[i], tau, &wi, &FR);
        f_mul(&u[i], &k0, &wi, &FR);
        pre[i] = acc; f_mul(&acc, &acc, &den[i], &FR);
        f_mul(&wi, &wi, &w, &FR);
    }
    f_inv(&acc, &acc, &FR);
    for (size_t i = n; i-- > 0;) { u256 di; f_mul(&di, &acc, &pre[i], &FR); f_mul(&acc, &acc, &den[i], &FR); f_mul(&u[i], &u[i], &di, &FR); }
    free(den); free(pre);
    /* instance_map_with_evaluation */
    u256 *qa = (u256 *)calloc(nv, sizeof(u256)), *qb = (u256 *)calloc(nv, sizeof(u256)), *qc = (u256 *)calloc(nv, sizeof(u256));
    for (size_t j = 0; j < ni; j++) qa[j] = u[m + j];
    u256 *q[3] = {qa, qb, qc};
    for (int k = 0; k < 3; k++)
        for (size_t i = 0; i < m; i++)
            for (uint32_t t = c->M[k].rowptr[i]; t < c->M[k].rowptr[i + 1]; t++) {
                u256 x; f_mul(&x, &u[i], &c->M[k].val[t], &FR);
                f_add(&q[k][c->M[k].col[t]], &q[k][c->M[k].col[t]], &x, &FR);
            }
    free(u);
    u256 ginv, dinv;
    f_inv(&ginv, gamma, &FR); f_inv(&dinv, delta, &FR);
    /* scalars for every query, canonical */
    size_t n_g1 = 3 + ni + nv * 2 + (n - 1) + c->n_wit, n_g2 = 3 + nv;
    u256 *s1 = (u256 *)malloc(sizeof(u256) * n_g1), *s2 = (u256 *)malloc(sizeof(u256) * n_g2);
    size_t o = 0;
    s1[o++] = *alpha; s1[o++] = *beta; s1[o++] = *delta;
    size_t off_abc = o;
    for (size_t j = 0; j < nv; j++) {
        u256 t, x;
        f_mul(&t, beta, &qa[j], &FR); f_mul(&x, alpha, &qb[j], &FR); f_add(&t, &t, &x, &FR); f_add(&t, &t, &qc[j], &FR);
        qc[j] = t;                                       /* beta*a + alpha*b + c */
    }
    for (size_t j = 0; j < ni; j++) f_mul(&s1[o++], &qc[j], &ginv, &FR);
    size_t off_a = o; for (size_t j = 0; j < nv; j++) s1[o++] = qa[j];
    size_t off_b1 = o; for (size_t j = 0; j < nv; j++) s1[o++] = qb[j];
    size_t off_h = o;
    { u256 p; f_mul(&p, &zt, &dinv, &FR); for (size_t i = 0; i + 1 < n; i++) { s1[o++] = p; f_mul(&p, &p, tau, &FR); } }
    size_t off_l = o; for (size_t j = ni; j < nv; j++) f_mul(&s1[o++], &qc[j], &dinv, &FR);
    s2[0] = *beta; s2[1] = *gamma; s2[2] = *delta;
    for (size_t j = 0; j < nv; j++) s2[3 + j] = qb[j];
    free(qa); free(qb); free(qc);
    for (size_t i = 0; i < n_g1; i++) f_to_canon(&s1[i], &s1[i], &FR);
    for (size_t i = 0; i < n_g2; i++) f_to_canon(&s2[i], &s2[i], &FR);
    g1_aff g1; f_from_u64(&g1.x, 1, &FQ); f_from_u64(&g1.y, 2, &FQ); g1.inf = 0;
    g2_aff g2; g2_generator(&g2);
    g1_aff *t1 = g1_fixed_table(&g1); g2_aff *t2 = g2_fixed_table(&g2);
    g1_jac *j1 = (g1_jac *)malloc(sizeof(g1_jac) * n_g1); g2_jac *j2 = (g2_jac *)malloc(sizeof(g2_jac) * n_g2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_g1; i++) g1_fixed_mul(&j1[i], t1, &s1[i]);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_g2; i++) g2_fixed_mul(&j2[i], t2, &s2[i]);
    g1_aff *p1 = (g1_aff *)malloc(sizeof(g1_aff) * n_g1); g2_aff *p2 = (g2_aff *)malloc(sizeof(g2_aff) * n_g2);
    g1_jac_batch_to_aff(p1, j1, n_g1); g2_jac_batch_to_aff(p2, j2, n_g2);
    free(j1); free(j2); free(t1); free(t2); free(s1); free(s2);
    /* vk: alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1 */
    for (int pass = 0; pass < 2; pass++) {
        uint8_t *p = pass == 0 ? pk_out : vk_out;
        if (!p) continue;
        g1_write(p, &p1[0]); p += 64;
        g2_write(p, &p2[0]); p += 128; g2_write(p, &p2[1]); p += 128; g2_write(p, &p2[2]); p += 128;
        wr_u64(&p, ni); for (size_t j = 0; j < ni; j++, p += 64) g1_write(p, &p1[off_abc + j]);
        if (pass == 1) break;
        g1_write(p, &p1[1]); p += 64; g1_write(p, &p1[2]); p += 64;
        wr_u64(&p, nv); for (size_t j = 0; j < nv; j++, p += 64) g1_write(p, &p1[off_a + j]);
        wr_u64(&p, nv); for (size_t j = 0; j < nv; j++, p += 64) g1_write(p, &p1[off_b1 + j]);
        wr_u64(&p, nv); for (size_t j = 0; j < nv; j++, p += 128) g2_write(p, &p2[3 + j]);
        wr_u64(&p, n - 1); for (size_t j = 0; j + 1 < n; j++, p += 64) g1_write(p, &p1[off_h + j]);
        wr_u64(&p, c->n_wit); for (size_t j = 0; j < c->n_wit; j++, p += 64) g1_write(p, &p1[off_l + j]);
    }
    free(p1); free(p2);
    return 0;
}

/* ----------------------------------------------------------------- prover
 * create_proof_with_assignment + calculate_coeff [UPSTREAM]. */
static void prove_core(const ora_circuit *c, const ora_pk *pk, const u256 *z /* Montgomery */,
                       const u256 *r, const u256 *s /* Montgomery */, uint8_t out[256], int threads) {
    size_t nv = (size_t)c->n_inst + c->n_wit, n = (size_t)1 << c->log_n, ni = c->n_inst;
    u256 *h = (u256 *)malloc(sizeof(u256) * n);
    witness_map(c, z, h, threads);
    u256 *hc = (u256 *)malloc(sizeof(u256) * n), *zc = (u256 *)malloc(sizeof(u256) * nv);
    for (size_t i = 0; i < n; i++) f_to_canon(&hc[i], &h[i], &FR);
    for (size_t i = 0; i < nv; i++) f_to_canon(&zc[i], &z[i], &FR);
    free(h);
    u256 rc, sc, rs, rsc;
    f_to_canon(&rc, r, &FR); f_to_canon(&sc, s, &FR); f_mul(&rs, r, s, &FR); f_to_canon(&rsc, &rs, &FR);
    g1_jac h_acc, l_acc, t, ga, gb1, gc, dj;
    size_t nh = pk->n_h < n ? pk->n_h : n, nl = pk->n_l < nv - ni ? pk->n_l : nv - ni;
    g1_msm(&h_acc, pk->h_query, hc, nh, threads);
    g1_msm(&l_acc, pk->l_query, zc + ni, nl, threads);
    g1_jac_from_aff(&dj, &pk->delta_g1);
    size_t na = (pk->n_a ? pk->n_a - 1 : 0) < nv - 1 ? (pk->n_a ? pk->n_a - 1 : 0) : nv - 1;
    /* A = r*delta + a_query[0] + msm(a_query[1..], z[1..]) + alpha */
    g1_jac_mul(&ga, &dj, &rc);
    g1_jac_add_mixed(&ga, &ga, &pk->a_query[0]);
    g1_msm(&t, pk->a_query + 1, zc + 1, na, threads); g1_jac_add(&ga, &ga, &t);
    g1_jac_add_mixed(&ga, &ga, &pk->alpha_g1);
    /* B1 (skipped iff r == 0) */
    if (!u256_is_zero(&rc)) {
        g1_jac_mul(&gb1, &dj, &sc);
        g1_jac_add_mixed(&gb1, &gb1, &pk->b_g1_query[0]);
        g1_msm(&t, pk->b_g1_query + 1, zc + 1, na, threads); g1_jac_add(&gb1, &gb1, &t);
        g1_jac_add_mixed(&gb1, &gb1, &pk->beta_g1);
    } else g1_jac_set_inf(&gb1);
    /* B2 */
    g2_jac gb2, t2, d2;
    g2_jac_from_aff(&d2, &pk->delta_g2);
    g2_jac_mul(&gb2, &d2, &sc);
    g2_jac_add_mixed(&gb2, &gb2, &pk->b_g2_query[0]);
    g2_msm(&t2, pk->b_g2_query + 1, zc + 1, na, threads); g2_jac_add(&gb2, &gb2, &t2);
    g2_jac_add_mixed(&gb2, &gb2, &pk->beta_g2);
    /* C = s*A + r*B1 - rs*delta + L + H */
    g1_jac_mul(&gc, &ga, &sc);
    g1_jac_mul(&t, &gb1, &rc); g1_jac_add(&gc, &gc, &t);
    g1_jac_mul(&t, &dj, &rsc); g1_jac_neg(&t, &t); g1_jac_add(&gc, &gc, &t);
    g1_jac_add(&gc, &gc, &l_acc); g1_jac_add(&gc, &gc, &h_acc);
    g1_aff aa, ca; g2_aff ba;
    g1_jac_to_aff(&aa, &ga); g2_jac_to_aff(&ba, &gb2); g1_jac_to_aff(&ca, &gc);
    g1_write(out, &aa); g2_write(out + 64, &ba); g1_write(out + 192, &ca);
    free(hc); free(zc);
}

/* One proof from a full assignment z (n_vars x 32 B canonical LE), r, s canonical. */
EXPORT int ora_prove(const ora_circuit *c, const ora_pk *pk, const uint8_t *z_in, const uint8_t r_in[32],
                     const uint8_t s_in[32], uint8_t out[256], int threads) {
    ora_init();
    size_t nv = (size_t)c->n_inst + c->n_wit;
    if (pk->n_a != nv || pk->n_b1 != nv || pk->n_b2 != nv || pk->n_l != c->n_wit) return -1;
    u256 *z = (u256 *)malloc(sizeof(u256) * nv), r, s;
    for (size_t i = 0; i < nv; i++) { u256 v; u256_from_le(&v, z_in + 32 * i); if (u256_cmp(&v, &FR.p) >= 0) { free(z); return -2; } f_from_canon(&z[i], &v, &FR); }
    u256_from_le(&r, r_in); u256_from_le(&s, s_in);
    if (u256_cmp(&r, &FR.p) >= 0 || u256_cmp(&s, &FR.p) >= 0) { free(z); return -2; }
    f_from_canon(&r, &r, &FR); f_from_canon(&s, &s, &FR);
    prove_core(c, pk, z, &r, &s, out, threads);
    free(z);
    return 0;
}

/* The reference's process_batch shape (src/advanced/batch.rs:123-131): a
 * proof-parallel map over operations, each running the whole single-proof path
 * (synthesis -> witness map -> 5 MSMs -> serialize) on one worker.
 * ops: a[i], b[i] (equality) or value a[i] + set rows (membership: sets is
 * n x set_stride u64, set_len[i] used).  r, s: n x 32 B.  out: n x 256 B.
 * status[i] != 0 marks a failed proof (empty Vec in the reference). */
EXPORT int ora_prove_batch(const ora_circuit *c, const ora_pk *pk, size_t n, const uint64_t *a, const uint64_t *b,
                           const uint64_t *sets, const uint32_t *set_len, uint32_t set_stride,
                           const uint8_t *r_in, const uint8_t *s_in, uint8_t *out, int32_t *status, int threads) {
    ora_init();
    size_t nv = (size_t)c->n_inst + c->n_wit;
    int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (size_t i = 0; i < n; i++) {
        u256 *z = (u256 *)malloc(sizeof(u256) * nv), cm, r, s;
        status[i] = 0;
        if (c->kind == 0 && a[i] != b[i]) status[i] = 1;                 /* snark.rs:344 */
        mimc_native(&cm, a[i]);                                           /* commit_value_snark, equality_proof.rs:13 */
        if (!status[i]) {
            int rc = circuit_assign(c, z, a[i], c->kind == 0 ? b[i] : 0, sets ? sets + i * set_stride : NULL,
                                    set_len ? set_len[i] : 0, &cm);
            if (rc) status[i] = 2;
        }
        if (!status[i]) {
            u256_from_le(&r, r_in + 32 * i); u256_from_le(&s, s_in + 32 * i);
            f_from_canon(&r, &r, &FR); f_from_canon(&s, &s, &FR);
            prove_core(c, pk, z, &r, &s, out + 256 * i, 1);
        } else memset(out + 256 * i, 0, 256);
        free(z);
    }
    return 0;
}

EXPORT int ora_num_threads(void) { return omp_get_max_threads(); }
