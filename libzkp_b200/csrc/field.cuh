// field.cuh — 256-bit Montgomery prime-field arithmetic for sm_100a on 8 x 32-bit limbs.
//
// Replaces what ark-ff's Fp256<MontBackend<_,4>> does for the reference's prover
// (reference call sites: src/backend/snark.rs:194,203-208 and everything inside
// Groth16::prove at snark.rs:364,442).  Same field, same Montgomery radix R = 2^256,
// values always fully reduced, so canonical outputs are bit-identical.
//
// Multiplication is a CIOS Montgomery product built from carry-chained
// mad.lo.cc / madc.hi.cc rows; ptxas fuses each lo/hi pair into one
// IMAD.WIDE.U32(.X), so one product costs 128 wide multiply-adds + 8 IMADs on the
// fma pipe.  Two partial accumulators ("aligned" limbs k <-> column k, "offset"
// limbs k <-> column k+1) let every row be one uninterrupted carry chain.
//
// Every carry chain lives inside ONE asm statement, so the PTX condition-code
// register never has to survive between statements.  The same row primitives have
// a portable C++ body (used when this header is compiled for the host), which lets
// tests/test_field_host.py check the composition logic without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LZ_HD __host__ __device__ __forceinline__
#define LZ_COLD __host__ __device__ __noinline__
#else
#define LZ_HD inline
#define LZ_COLD inline
#endif
#define LZ_CONST_ARRAY(NAME, ...)                                   \
    LZ_HD static constexpr uint32_t NAME(int i) {                   \
        const uint32_t v[8] = {__VA_ARGS__};                        \
        return v[i];                                                \
    }

#include "bn254_constants.cuh"

namespace lzkp {

// --------------------------------------------------------------------------
// Row primitives.  D is an 8-limb accumulator; (s0..s3) are four multiplicand
// limbs, k the multiplier limb.  Pair t of D receives the 64-bit product s_t*k.
// --------------------------------------------------------------------------

// D = (s0,s1,s2,s3) * k
LZ_HD void row_mul(uint32_t (&D)[8], uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t k) {
#ifdef __CUDA_ARCH__
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(D[0]), "=r"(D[1]), "=r"(D[2]), "=r"(D[3]), "=r"(D[4]), "=r"(D[5]), "=r"(D[6]), "=r"(D[7])
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(k));
#else
    const uint32_t s[4] = {s0, s1, s2, s3};
    for (int t = 0; t < 4; t++) {
        uint64_t p = (uint64_t)s[t] * k;
        D[2 * t] = (uint32_t)p;
        D[2 * t + 1] = (uint32_t)(p >> 32);
    }
#endif
}

// D += (s0,s1,s2,s3) * k as one carry chain; returns the carry out of D[7].
LZ_HD uint32_t row_mad(uint32_t (&D)[8], uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t k) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7]), "=r"(cy)
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(k));
#else
    const uint32_t s[4] = {s0, s1, s2, s3};
    uint64_t c = 0;
    for (int t = 0; t < 4; t++) {
        uint64_t p = (uint64_t)s[t] * k;
        uint64_t lo = (uint64_t)D[2 * t] + (uint32_t)p + c;
        D[2 * t] = (uint32_t)lo;
        uint64_t hi = (uint64_t)D[2 * t + 1] + (uint32_t)(p >> 32) + (lo >> 32);
        D[2 * t + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    cy = (uint32_t)c;
#endif
    return cy;
}

// The per-iteration column shift of CIOS, fused with the next offset row:
//   X0 += D[1]                     (limb that drops to column 0 joins the aligned array)
//   D[j] = s*k + D[j+2] + carry    (D moves down two limbs and becomes the offset array)
// D[6..7] receive the last product plus carry only.
LZ_HD void row_mad_shift(uint32_t &X0, uint32_t (&D)[8], uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                         uint32_t k) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(X0), "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7])
        : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(k));
#else
    const uint32_t s[4] = {s0, s1, s2, s3};
    uint64_t c = (uint64_t)X0 + D[1];
    X0 = (uint32_t)c;
    c >>= 32;
    for (int t = 0; t < 4; t++) {
        uint64_t p = (uint64_t)s[t] * k;
        uint32_t alo = t < 3 ? D[2 * t + 2] : 0, ahi = t < 3 ? D[2 * t + 3] : 0;
        uint64_t lo = (uint64_t)alo + (uint32_t)p + c;
        uint64_t hi = (uint64_t)ahi + (uint32_t)(p >> 32) + (lo >> 32);
        D[2 * t] = (uint32_t)lo;
        D[2 * t + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
#endif
}

// r = a + b, returns carry.  r = a - b, returns borrow (1 if a < b).
LZ_HD uint32_t add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t cy;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(cy)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    cy = (uint32_t)c;
#endif
    return cy;
}

LZ_HD uint32_t sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t bw;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    bw &= 1u;   // subc of 0-0-borrow yields 0xffffffff on borrow
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - c;
        r[i] = (uint32_t)d;
        c = (d >> 32) & 1;
    }
    bw = (uint32_t)c;
#endif
    return bw;
}

// --------------------------------------------------------------------------
// Field element.  P supplies INV, MOD(i), ONE(i), R2(i), ...
// --------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
    uint32_t l[8];

    LZ_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    LZ_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::ONE(i);
        return r;
    }
    LZ_HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::MOD(i);
        return r;
    }
    LZ_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::R2(i);
        return r;
    }
    LZ_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= l[i];
        return o == 0;
    }
    LZ_HD bool operator==(const Fp &b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    LZ_HD bool operator!=(const Fp &b) const { return !(*this == b); }

    // r = x - p if x >= p else x   (x < 2p)
    LZ_HD static Fp reduce_once(const uint32_t (&x)[8]) {
        Fp m = modulus(), d, r;
        uint32_t bw = sub8(d.l, x, m.l);
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = bw ? x[i] : d.l[i];
        return r;
    }
    LZ_HD friend Fp operator+(const Fp &a, const Fp &b) {
        uint32_t s[8];
        add8(s, a.l, b.l);      // p < 2^254: no carry out
        return reduce_once(s);
    }
    LZ_HD friend Fp operator-(const Fp &a, const Fp &b) {
        Fp d, e, m = modulus(), r;
        uint32_t bw = sub8(d.l, a.l, b.l);
        add8(e.l, d.l, m.l);
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = bw ? e.l[i] : d.l[i];
        return r;
    }
    LZ_HD Fp neg() const {
        Fp m = modulus(), d, r;
        sub8(d.l, m.l, l);
        bool z = is_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : d.l[i];
        return r;
    }
    LZ_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod p.
    LZ_HD friend Fp operator*(const Fp &a, const Fp &b) {
        uint32_t X[8], Y[8];      // aligned / offset accumulators (roles swap every iteration)
        uint32_t m, c;
        // i = 0
        row_mul(X, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
        row_mul(Y, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
        m = X[0] * P::INV;
        row_mad(Y, P::MOD(1), P::MOD(3), P::MOD(5), P::MOD(7), m);
        c = row_mad(X, P::MOD(0), P::MOD(2), P::MOD(4), P::MOD(6), m);
        Y[7] += c;
#pragma unroll
        for (int i = 1; i < 8; i++) {
            if (i & 1) {
                // X was just reduced (X[0] == 0): Y becomes aligned, X shifts into the offset role.
                row_mad_shift(Y[0], X, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
                c = row_mad(Y, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
                X[7] += c;
                m = Y[0] * P::INV;
                row_mad(X, P::MOD(1), P::MOD(3), P::MOD(5), P::MOD(7), m);
                c = row_mad(Y, P::MOD(0), P::MOD(2), P::MOD(4), P::MOD(6), m);
                X[7] += c;
            } else {
                row_mad_shift(X[0], Y, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
                c = row_mad(X, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
                Y[7] += c;
                m = X[0] * P::INV;
                row_mad(Y, P::MOD(1), P::MOD(3), P::MOD(5), P::MOD(7), m);
                c = row_mad(X, P::MOD(0), P::MOD(2), P::MOD(4), P::MOD(6), m);
                Y[7] += c;
            }
        }
        // After i = 7 (odd): Y aligned with Y[0] == 0, X offset.  result = X + (Y >> 32).
        uint32_t hi[8], s[8];
#pragma unroll
        for (int j = 0; j < 7; j++) hi[j] = Y[j + 1];
        hi[7] = 0;
        add8(s, X, hi);
        return reduce_once(s);
    }
    LZ_HD Fp sqr() const { return *this * *this; }

    // canonical (plain integer, < p) <-> Montgomery
    LZ_HD static Fp from_canonical(const Fp &c) { return c * r2(); }
    LZ_HD Fp to_canonical() const {
        Fp o = zero();
        o.l[0] = 1;
        return *this * o;
    }
    // Inverse by the binary extended Euclidean algorithm (HAC 14.61 for an odd modulus); 0 -> 0.
    // ~750 shift / add / subtract steps on 256-bit integers: about 5x shorter as a serial chain than the
    // Fermat ladder (254 squarings + 127 products), which matters because every use is latency-bound
    // (3 per proof in k_assemble, 1 per MSM result, 1 per table entry group).  Works on the limbs as a
    // plain integer, so for a Montgomery input aR it yields (aR)^-1; one product with R^3 restores a^-1 R.
    LZ_COLD Fp inverse() const {
        if (is_zero()) return zero();
        uint32_t u[8], v[8], x1[8], x2[8];
        const Fp m = modulus();
#pragma unroll
        for (int i = 0; i < 8; i++) { u[i] = l[i]; v[i] = m.l[i]; x1[i] = 0; x2[i] = 0; }
        x1[0] = 1;
        auto is_one = [](const uint32_t (&a)[8]) {
            uint32_t o = a[0] ^ 1u;
#pragma unroll
            for (int i = 1; i < 8; i++) o |= a[i];
            return o == 0;
        };
        auto shr1 = [](uint32_t (&a)[8]) {
#pragma unroll
            for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
            a[7] >>= 1;
        };
        auto halve_mod = [&](uint32_t (&x)[8]) {          // x / 2 mod p, x < p
            if (x[0] & 1u) { uint32_t t[8]; add8(t, x, m.l); 
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = t[i]; }
            shr1(x);
        };
        auto sub_mod = [&](uint32_t (&x)[8], const uint32_t (&y)[8]) {   // x = x - y mod p
            uint32_t d[8], e[8];
            uint32_t bw = sub8(d, x, y);
            add8(e, d, m.l);
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = bw ? e[i] : d[i];
        };
#pragma unroll 1
        while (!is_one(u) && !is_one(v)) {
#pragma unroll 1
            while (!(u[0] & 1u)) { shr1(u); halve_mod(x1); }
#pragma unroll 1
            while (!(v[0] & 1u)) { shr1(v); halve_mod(x2); }
            uint32_t d[8];
            if (sub8(d, u, v) == 0) {                       // u >= v
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = d[i];
                sub_mod(x1, x2);
            } else {
                sub8(d, v, u);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = d[i];
                sub_mod(x2, x1);
            }
        }
        Fp r, r3;
        const bool uu = is_one(u);
#pragma unroll
        for (int i = 0; i < 8; i++) { r.l[i] = uu ? x1[i] : x2[i]; r3.l[i] = P::R3(i); }
        return r * r3;
    }
    // canonical a > b ?
    LZ_HD static bool gt_canonical(const Fp &a, const Fp &b) {
        uint32_t d[8];
        return sub8(d, b.l, a.l) != 0;   // b - a borrows  <=>  a > b
    }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

// --------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2 + 1)
// --------------------------------------------------------------------------
struct Fq2;
#if !defined(LZKP_FQ2_INLINE) && defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq2 fq2_mul_call(Fq2 a, Fq2 b);
static __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 a);
#endif
struct Fq2 {
    Fq c0, c1;
    LZ_HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    LZ_HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    LZ_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    LZ_HD bool operator==(const Fq2 &b) const { return c0 == b.c0 && c1 == b.c1; }
    LZ_HD bool operator!=(const Fq2 &b) const { return !(*this == b); }
    LZ_HD friend Fq2 operator+(const Fq2 &a, const Fq2 &b) { return Fq2{a.c0 + b.c0, a.c1 + b.c1}; }
    LZ_HD friend Fq2 operator-(const Fq2 &a, const Fq2 &b) { return Fq2{a.c0 - b.c0, a.c1 - b.c1}; }
    LZ_HD Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
    LZ_HD Fq2 dbl() const { return Fq2{c0.dbl(), c1.dbl()}; }
    LZ_HD friend Fq2 operator*(const Fq2 &a, const Fq2 &b) {
#if !defined(LZKP_FQ2_INLINE) && defined(__CUDA_ARCH__)
        return fq2_mul_call(a, b);
#elif defined(LZKP_FQ2_SCHOOLBOOK)
        // four independent products, two additions: more multiplies than Karatsuba but no dependent add chains
        return Fq2{a.c0 * b.c0 - a.c1 * b.c1, a.c0 * b.c1 + a.c1 * b.c0};
#else
        Fq v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;                 // Karatsuba, 3 Fq products
        Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
        return Fq2{v0 - v1, s - v0 - v1};
#endif
    }
    LZ_HD Fq2 sqr() const {                                     // 2 Fq products
#if !defined(LZKP_FQ2_INLINE) && defined(__CUDA_ARCH__)
        return fq2_sqr_call(*this);
#endif
        Fq m = c0 * c1;
        return Fq2{(c0 + c1) * (c0 - c1), m.dbl()};
    }
    LZ_HD Fq2 inverse() const {
        Fq n = (c0.sqr() + c1.sqr()).inverse();
        return Fq2{c0 * n, (c1 * n).neg()};
    }
};
#if !defined(LZKP_FQ2_INLINE) && defined(__CUDA_ARCH__)
// Out-of-line Fq2 product / square (operands and result in registers).  A G2 mixed addition with its 28 Montgomery
// products inlined is ~110 KB of SASS, more than the instruction cache holds: ncu showed 11 % of the G2 gather
// kernel's stall samples as no_instruction.  As calls the loop body is ~15 KB; measured -19 % on the G2 bucket
// accumulation and -5 % on the G2 table-gather kernel.  -DLZKP_FQ2_INLINE restores the inlined form.
static __device__ __noinline__ Fq2 fq2_mul_call(Fq2 a, Fq2 b) {
    Fq v0 = a.c0 * b.c0, v1 = a.c1 * b.c1;
    Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
    return Fq2{v0 - v1, s - v0 - v1};
}
static __device__ __noinline__ Fq2 fq2_sqr_call(Fq2 a) {
    Fq m = a.c0 * a.c1;
    return Fq2{(a.c0 + a.c1) * (a.c0 - a.c1), m.dbl()};
}
#endif

}  // namespace lzkp
