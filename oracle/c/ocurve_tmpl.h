/* oracle/c/ocurve_tmpl.h — short-Weierstrass (a = 0) group law + Pippenger MSM,
 * written once and instantiated for G1 (over Fq) and G2 (over Fq2).
 *
 * TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (oracle/zkp_oracle.py header).
 * Restates ark-ec ^0.5 [UPSTREAM]: `short_weierstrass::Projective` (Jacobian
 * coordinates, mixed addition) and `VariableBaseMSM::msm_bigint` in its
 * signed-digit form (`make_digits`, one task per window, running-sum bucket
 * reduction, windows folded high-to-low with c doublings).
 *
 * Instantiate with: FE, PFX(name), and FE_* macros (add, sub, neg, mul, sqr,
 * inv, is_zero, eq, set_one, set_zero).
 */

typedef struct { FE x, y; int inf; } PFX(aff);
typedef struct { FE x, y, z; } PFX(jac);

static inline void PFX(jac_set_inf)(PFX(jac) *p) { FE_SET_ONE(&p->x); FE_SET_ONE(&p->y); FE_SET_ZERO(&p->z); }
static inline int PFX(jac_is_inf)(const PFX(jac) *p) { return FE_IS_ZERO(&p->z); }
static inline void PFX(jac_from_aff)(PFX(jac) *o, const PFX(aff) *a) {
    if (a->inf) { PFX(jac_set_inf)(o); return; }
    o->x = a->x; o->y = a->y; FE_SET_ONE(&o->z);
}
static inline void PFX(jac_neg)(PFX(jac) *o, const PFX(jac) *p) { o->x = p->x; FE_NEG(&o->y, &p->y); o->z = p->z; }
static inline void PFX(aff_neg)(PFX(aff) *o, const PFX(aff) *p) { o->x = p->x; FE_NEG(&o->y, &p->y); o->inf = p->inf; }

static void PFX(jac_dbl)(PFX(jac) *o, const PFX(jac) *p) {
    if (FE_IS_ZERO(&p->z) || FE_IS_ZERO(&p->y)) { PFX(jac_set_inf)(o); return; }
    FE A, B, C, D, E, F, t, X3, Y3, Z3;
    FE_SQR(&A, &p->x); FE_SQR(&B, &p->y); FE_SQR(&C, &B);
    FE_ADD(&t, &p->x, &B); FE_SQR(&t, &t); FE_SUB(&t, &t, &A); FE_SUB(&t, &t, &C);
    FE_ADD(&D, &t, &t);
    FE_ADD(&E, &A, &A); FE_ADD(&E, &E, &A);
    FE_SQR(&F, &E);
    FE_SUB(&X3, &F, &D); FE_SUB(&X3, &X3, &D);
    FE_ADD(&C, &C, &C); FE_ADD(&C, &C, &C); FE_ADD(&C, &C, &C);
    FE_SUB(&t, &D, &X3); FE_MUL(&Y3, &E, &t); FE_SUB(&Y3, &Y3, &C);
    FE_ADD(&t, &p->y, &p->y); FE_MUL(&Z3, &t, &p->z);
    o->x = X3; o->y = Y3; o->z = Z3;
}

static void PFX(jac_add_mixed)(PFX(jac) *o, const PFX(jac) *p, const PFX(aff) *q) {
    if (q->inf) { *o = *p; return; }
    if (FE_IS_ZERO(&p->z)) { PFX(jac_from_aff)(o, q); return; }
    FE Z1Z1, U2, S2, H, R, HH, HHH, V, t, X3, Y3, Z3;
    FE_SQR(&Z1Z1, &p->z);
    FE_MUL(&U2, &q->x, &Z1Z1);
    FE_MUL(&S2, &q->y, &p->z); FE_MUL(&S2, &S2, &Z1Z1);
    if (FE_EQ(&U2, &p->x)) {
        if (FE_EQ(&S2, &p->y)) { PFX(jac_dbl)(o, p); return; }
        PFX(jac_set_inf)(o); return;
    }
    FE_SUB(&H, &U2, &p->x); FE_SUB(&R, &S2, &p->y);
    FE_SQR(&HH, &H); FE_MUL(&HHH, &H, &HH); FE_MUL(&V, &p->x, &HH);
    FE_SQR(&X3, &R); FE_SUB(&X3, &X3, &HHH); FE_SUB(&X3, &X3, &V); FE_SUB(&X3, &X3, &V);
    FE_SUB(&t, &V, &X3); FE_MUL(&Y3, &R, &t); FE_MUL(&t, &p->y, &HHH); FE_SUB(&Y3, &Y3, &t);
    FE_MUL(&Z3, &p->z, &H);
    o->x = X3; o->y = Y3; o->z = Z3;
}

static void PFX(jac_add)(PFX(jac) *o, const PFX(jac) *p, const PFX(jac) *q) {
    if (FE_IS_ZERO(&p->z)) { *o = *q; return; }
    if (FE_IS_ZERO(&q->z)) { *o = *p; return; }
    FE Z1Z1, Z2Z2, U1, U2, S1, S2, H, R, HH, HHH, V, t, X3, Y3, Z3;
    FE_SQR(&Z1Z1, &p->z); FE_SQR(&Z2Z2, &q->z);
    FE_MUL(&U1, &p->x, &Z2Z2); FE_MUL(&U2, &q->x, &Z1Z1);
    FE_MUL(&S1, &p->y, &q->z); FE_MUL(&S1, &S1, &Z2Z2);
    FE_MUL(&S2, &q->y, &p->z); FE_MUL(&S2, &S2, &Z1Z1);
    if (FE_EQ(&U1, &U2)) {
        if (FE_EQ(&S1, &S2)) { PFX(jac_dbl)(o, p); return; }
        PFX(jac_set_inf)(o); return;
    }
    FE_SUB(&H, &U2, &U1); FE_SUB(&R, &S2, &S1);
    FE_SQR(&HH, &H); FE_MUL(&HHH, &H, &HH); FE_MUL(&V, &U1, &HH);
    FE_SQR(&X3, &R); FE_SUB(&X3, &X3, &HHH); FE_SUB(&X3, &X3, &V); FE_SUB(&X3, &X3, &V);
    FE_SUB(&t, &V, &X3); FE_MUL(&Y3, &R, &t); FE_MUL(&t, &S1, &HHH); FE_SUB(&Y3, &Y3, &t);
    FE_MUL(&Z3, &p->z, &q->z); FE_MUL(&Z3, &Z3, &H);
    o->x = X3; o->y = Y3; o->z = Z3;
}

static void PFX(jac_to_aff)(PFX(aff) *o, const PFX(jac) *p) {
    if (FE_IS_ZERO(&p->z)) { FE_SET_ZERO(&o->x); FE_SET_ZERO(&o->y); o->inf = 1; return; }
    FE zi, zi2;
    FE_INV(&zi, &p->z); FE_SQR(&zi2, &zi);
    FE_MUL(&o->x, &p->x, &zi2);
    FE_MUL(&zi2, &zi2, &zi); FE_MUL(&o->y, &p->y, &zi2);
    o->inf = 0;
}

/* Montgomery-trick batch normalisation (setup only). */
static void PFX(jac_batch_to_aff)(PFX(aff) *o, const PFX(jac) *p, size_t n) {
    FE *pre = (FE *)malloc(sizeof(FE) * (n + 1));
    FE acc; FE_SET_ONE(&acc);
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!FE_IS_ZERO(&p[i].z)) FE_MUL(&acc, &acc, &p[i].z);
    }
    FE inv; FE_INV(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (FE_IS_ZERO(&p[i].z)) { FE_SET_ZERO(&o[i].x); FE_SET_ZERO(&o[i].y); o[i].inf = 1; continue; }
        FE zi, zi2;
        FE_MUL(&zi, &inv, &pre[i]);
        FE_MUL(&inv, &inv, &p[i].z);
        FE_SQR(&zi2, &zi);
        FE_MUL(&o[i].x, &p[i].x, &zi2);
        FE_MUL(&zi2, &zi2, &zi); FE_MUL(&o[i].y, &p[i].y, &zi2);
        o[i].inf = 0;
    }
    free(pre);
}

/* k * P, k canonical (non-Montgomery) integer, MSB-first double-and-add. */
static void PFX(jac_mul)(PFX(jac) *o, const PFX(jac) *p, const u256 *k) {
    PFX(jac) acc; PFX(jac_set_inf)(&acc);
    int top = 255;
    while (top >= 0 && !u256_bit(k, top)) top--;
    for (int i = top; i >= 0; i--) {
        PFX(jac_dbl)(&acc, &acc);
        if (u256_bit(k, i)) PFX(jac_add)(&acc, &acc, p);
    }
    *o = acc;
}

/* ark-ec make_digits [UPSTREAM]: signed radix-2^w digits of a canonical scalar. */
static void PFX(make_digits)(int64_t *digits, const u256 *a, unsigned w, unsigned num_bits) {
    const uint64_t radix = 1ull << w, mask = radix - 1;
    uint64_t carry = 0;
    unsigned count = (num_bits + w - 1) / w;
    for (unsigned i = 0; i < count; i++) {
        unsigned off = i * w, wi = off / 64, bi = off % 64;
        uint64_t buf;
        if (bi < 64 - w || wi == 3) buf = a->l[wi] >> bi;
        else buf = (a->l[wi] >> bi) | (a->l[wi + 1] << (64 - bi));
        uint64_t coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        int64_t d = (int64_t)coef - (int64_t)(carry << w);
        if (i == count - 1) d += (int64_t)(carry << w);
        digits[i] = d;
    }
}

#ifdef ORA_TMPL_FIRST
static unsigned ora_msm_window(size_t n) {
    if (n < 32) return 3;
    unsigned lg = 0;                               /* ark_std::log2 = ceil(log2 n) */
    while (((size_t)1 << lg) < n) lg++;
    return lg * 69 / 100 + 2;                      /* ln_without_floats(n) + 2 */
}
#endif

/* Sum scalars[i] * bases[i] (scalars canonical). threads <= 0 => OpenMP default. */
static void PFX(msm)(PFX(jac) *out, const PFX(aff) *bases, const u256 *scalars, size_t n, int threads) {
    PFX(jac_set_inf)(out);
    if (n == 0) return;
    const unsigned c = ora_msm_window(n), num_bits = 254;
    const unsigned W = (num_bits + c - 1) / c;
    int64_t *digits = (int64_t *)malloc(sizeof(int64_t) * n * W);
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : omp_get_max_threads())
    for (size_t i = 0; i < n; i++) PFX(make_digits)(digits + i * W, &scalars[i], c, num_bits);
    PFX(jac) *sums = (PFX(jac) *)malloc(sizeof(PFX(jac)) * W);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
    for (unsigned w = 0; w < W; w++) {
        size_t nb = (size_t)1 << c;
        PFX(jac) *buckets = (PFX(jac) *)malloc(sizeof(PFX(jac)) * nb);
        for (size_t b = 0; b < nb; b++) PFX(jac_set_inf)(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int64_t d = digits[i * W + w];
            if (d > 0) PFX(jac_add_mixed)(&buckets[d - 1], &buckets[d - 1], &bases[i]);
            else if (d < 0) {
                PFX(aff) nb_; PFX(aff_neg)(&nb_, &bases[i]);
                PFX(jac_add_mixed)(&buckets[-d - 1], &buckets[-d - 1], &nb_);
            }
        }
        PFX(jac) run, res; PFX(jac_set_inf)(&run); PFX(jac_set_inf)(&res);
        for (size_t b = nb; b-- > 0;) {
            PFX(jac_add)(&run, &run, &buckets[b]);
            PFX(jac_add)(&res, &res, &run);
        }
        sums[w] = res;
        free(buckets);
    }
    PFX(jac) total; PFX(jac_set_inf)(&total);
    for (unsigned w = W; w-- > 1;) {
        PFX(jac_add)(&total, &total, &sums[w]);
        for (unsigned k = 0; k < c; k++) PFX(jac_dbl)(&total, &total);
    }
    PFX(jac_add)(out, &sums[0], &total);
    free(sums); free(digits);
}

/* Fixed-base comb for trapdoor setup: tbl[w][d-1] = d * 2^(8w) * P (affine). */
static PFX(aff) *PFX(fixed_table)(const PFX(aff) *P) {
    const int W = 32, D = 255;
    PFX(jac) *j = (PFX(jac) *)malloc(sizeof(PFX(jac)) * W * D);
    PFX(jac) base; PFX(jac_from_aff)(&base, P);
    for (int w = 0; w < W; w++) {
        j[w * D] = base;
        for (int d = 1; d < D; d++) PFX(jac_add)(&j[w * D + d], &j[w * D + d - 1], &base);
        PFX(jac_add)(&base, &j[w * D + D - 1], &base);
    }
    PFX(aff) *t = (PFX(aff) *)malloc(sizeof(PFX(aff)) * W * D);
    PFX(jac_batch_to_aff)(t, j, (size_t)W * D);
    free(j);
    return t;
}
static void PFX(fixed_mul)(PFX(jac) *o, const PFX(aff) *tbl, const u256 *k) {
    PFX(jac) acc; PFX(jac_set_inf)(&acc);
    const uint8_t *kb = (const uint8_t *)k->l;
    for (int w = 0; w < 32; w++)
        if (kb[w]) PFX(jac_add_mixed)(&acc, &acc, &tbl[w * 255 + kb[w] - 1]);
    *o = acc;
}
