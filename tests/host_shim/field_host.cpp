// Host build of libzkp_b200/csrc/field.cuh (row primitives take their portable C++
// bodies) so the Montgomery composition logic can be checked without a GPU.
#include "../../libzkp_b200/csrc/ec.cuh"
#include <cstring>
using namespace lzkp;
template <class F> static void ld(F &f, const uint8_t *p) { std::memcpy(f.l, p, 32); }
template <class F> static void st(uint8_t *p, const F &f) { std::memcpy(p, f.l, 32); }
extern "C" {
// op: 0 mul(raw Montgomery product), 1 add, 2 sub, 3 neg, 4 from_canonical, 5 to_canonical, 6 inverse, 7 sqr
void fr_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *o) {
    Fr x, y, r; ld(x, a); ld(y, b);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
        case 4: r = Fr::from_canonical(x); break; case 5: r = x.to_canonical(); break; case 6: r = x.inverse(); break; default: r = x.sqr(); }
    st(o, r);
}
void fq_op(int op, const uint8_t *a, const uint8_t *b, uint8_t *o) {
    Fq x, y, r; ld(x, a); ld(y, b);
    switch (op) { case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
        case 4: r = Fq::from_canonical(x); break; case 5: r = x.to_canonical(); break; case 6: r = x.inverse(); break; default: r = x.sqr(); }
    st(o, r);
}
// Fq2 product with lazy reduction (field.cuh fq2_mul_lazy) and the wide primitives under it; a, b, o: c0 || c1 (64 B)
void fq2_mul_lazy_host(const uint8_t *a, const uint8_t *b, uint8_t *o) {
    Fq2 x, y; ld(x.c0, a); ld(x.c1, a + 32); ld(y.c0, b); ld(y.c1, b + 32);
    Fq2 r = fq2_mul_lazy(x, y);
    st(o, r.c0); st(o + 32, r.c1);
}
// a*b - c*d with one reduction (Fp::msub)
void fq_msub_host(const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d, uint8_t *o) {
    Fq x, y, z, w; ld(x, a); ld(y, b); ld(z, c); ld(w, d);
    st(o, Fq::msub(x, y, z, w));
}
void fq_mul_wide_host(const uint8_t *a, const uint8_t *b, uint8_t *o64) {
    Fq x, y; ld(x, a); ld(y, b);
    uint32_t T[16];
    mul_wide(T, x.l, y.l);
    std::memcpy(o64, T, 64);
}
void fq_mul_wide_k_host(const uint8_t *a, const uint8_t *b, uint8_t *o64) {      // Karatsuba form
    Fq x, y; ld(x, a); ld(y, b);
    uint32_t T[16];
    mul_wide_k(T, x.l, y.l);
    std::memcpy(o64, T, 64);
}
void fq_reduce_wide_host(const uint8_t *t64, uint8_t *o) {
    uint32_t T[16];
    std::memcpy(T, t64, 64);
    st(o, Fq::reduce_wide(T));
}
int fq_gt(const uint8_t *a, const uint8_t *b) { Fq x, y; ld(x, a); ld(y, b); return Fq::gt_canonical(x, y); }
// k * P on G1 by ec.cuh's windowed scalar multiplication (host build of the same code): p = x || y canonical (64 B),
// k canonical (32 B), nbits 128 or 254 picks scalar_mul_u128 / scalar_mul; out = affine x || y canonical, all zero = infinity
void g1_scalar_mul_host(const uint8_t *p, const uint8_t *k, int nbits, uint8_t *out) {
    Fq x, y; ld(x, p); ld(y, p + 32);
    G1XYZZ P = G1XYZZ::from_affine(G1Affine{Fq::from_canonical(x), Fq::from_canonical(y)});
    Fr s; ld(s, k);
    G1XYZZ R;
    if (nbits == 128) { uint32_t k4[4] = {s.l[0], s.l[1], s.l[2], s.l[3]}; R = scalar_mul_u128(P, k4); }
    else R = scalar_mul(P, s);
    G1Affine a = R.to_affine();
    std::memset(out, 0, 64);
    if (!a.is_inf()) { st(out, a.x.to_canonical()); st(out + 32, a.y.to_canonical()); }
}
// GLV split (ec.cuh): out = k1 (16 B) || k2 (16 B); returns neg1 | neg2 << 1 | ok << 2
int glv_split_host(const uint8_t *k, uint8_t *out) {
    Fr x; ld(x, k);
    GlvSplit sp = glv_split(x);
    std::memcpy(out, sp.k1, 16);
    std::memcpy(out + 16, sp.k2, 16);
    return (sp.neg1 ? 1 : 0) | (sp.neg2 ? 2 : 0) | (sp.ok ? 4 : 0);
}
}
