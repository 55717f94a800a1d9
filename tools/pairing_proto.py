#!/usr/bin/env python3
"""Big-int prototype of the tower-field optimal-ate pairing the GPU verifier implements (libzkp_b200/csrc/
pairing.cuh), checked against the oracle's textbook pairing over Fq[w]/(w^12 - 18 w^6 + 82).  It also derives the
Frobenius / twist constants the device code embeds.  Test infrastructure: run  python tools/pairing_proto.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zkp_oracle as O  # noqa: E402

P = O.Q_MOD
R = O.R_MOD
XI = (9, 1)


# ---- Fq2
def f2add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2sqr(a): return f2mul(a, a)
def f2conj(a): return (a[0], (-a[1]) % P)
def f2scal(a, k): return (a[0] * k % P, a[1] * k % P)
def f2inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)
def f2pow(a, e):
    r = (1, 0)
    for bit in bin(e)[2:]:
        r = f2sqr(r)
        if bit == '1':
            r = f2mul(r, a)
    return r
F2_ZERO, F2_ONE = (0, 0), (1, 0)


# ---- Fq6 = Fq2[v]/(v^3 - xi): (c0, c1, c2)
def f6add(a, b): return tuple(f2add(x, y) for x, y in zip(a, b))
def f6sub(a, b): return tuple(f2sub(x, y) for x, y in zip(a, b))
def f6neg(a): return tuple(f2neg(x) for x in a)
def f6mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    t0, t1, t2 = f2mul(a0, b0), f2mul(a1, b1), f2mul(a2, b2)
    c0 = f2add(t0, f2mul(XI, f2sub(f2sub(f2mul(f2add(a1, a2), f2add(b1, b2)), t1), t2)))
    c1 = f2add(f2sub(f2sub(f2mul(f2add(a0, a1), f2add(b0, b1)), t0), t1), f2mul(XI, t2))
    c2 = f2add(f2sub(f2sub(f2mul(f2add(a0, a2), f2add(b0, b2)), t0), t2), t1)
    return (c0, c1, c2)
def f6mulv(a): return (f2mul(XI, a[2]), a[0], a[1])          # times v
def f6inv(a):
    a0, a1, a2 = a
    c0 = f2sub(f2sqr(a0), f2mul(XI, f2mul(a1, a2)))
    c1 = f2sub(f2mul(XI, f2sqr(a2)), f2mul(a0, a1))
    c2 = f2sub(f2sqr(a1), f2mul(a0, a2))
    t = f2inv(f2add(f2mul(a0, c0), f2mul(XI, f2add(f2mul(a2, c1), f2mul(a1, c2)))))
    return (f2mul(c0, t), f2mul(c1, t), f2mul(c2, t))
F6_ZERO, F6_ONE = (F2_ZERO,) * 3, (F2_ONE, F2_ZERO, F2_ZERO)


# ---- Fq12 = Fq6[w]/(w^2 - v): (c0, c1)
def f12mul(a, b):
    t0, t1 = f6mul(a[0], b[0]), f6mul(a[1], b[1])
    c1 = f6sub(f6sub(f6mul(f6add(a[0], a[1]), f6add(b[0], b[1])), t0), t1)
    return (f6add(t0, f6mulv(t1)), c1)
def f12sqr(a): return f12mul(a, a)
def f12conj(a): return (a[0], f6neg(a[1]))
def f12inv(a):
    t = f6inv(f6sub(f6mul(a[0], a[0]), f6mulv(f6mul(a[1], a[1]))))
    return (f6mul(a[0], t), f6neg(f6mul(a[1], t)))
def f12pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12sqr(r)
        if bit == '1':
            r = f12mul(r, a)
    return r
F12_ONE = (F6_ONE, F6_ZERO)

# Frobenius^2 on Fq12: coefficient of v^i w^j is multiplied by xi^((p^2 - 1) (2i + j) / 6), an element of Fq
G2C = [f2pow(XI, (P * P - 1) * k // 6) for k in range(6)]
assert all(c[1] == 0 for c in G2C)
def f12frob2(a):
    return tuple(tuple(f2scal(a[j][i], G2C[2 * i + j][0]) for i in range(3)) for j in range(2))

# twist Frobenius constants (mul_by_char): pi(x, y) = (conj(x) * xi^((p-1)/3), conj(y) * xi^((p-1)/2))
TW_X, TW_Y = f2pow(XI, (P - 1) // 3), f2pow(XI, (P - 1) // 2)
B_TWIST = f2mul((3, 0), f2inv(XI))
TWO_INV = pow(2, P - 2, P)


def sparse(c0, c3, c4):
    """c0 + c3 * w + c4 * v w as a full Fq12 element (D-twist line)."""
    return ((c0, F2_ZERO, F2_ZERO), (c3, c4, F2_ZERO))


def dbl_step(Rp):
    X, Y, Z = Rp
    a = f2scal(f2mul(X, Y), TWO_INV)
    b, c = f2sqr(Y), f2sqr(Z)
    e = f2mul(B_TWIST, f2add(f2add(c, c), c))
    f = f2add(f2add(e, e), e)
    g = f2scal(f2add(b, f), TWO_INV)
    h = f2sub(f2sqr(f2add(Y, Z)), f2add(b, c))
    i = f2sub(e, b)
    j = f2sqr(X)
    e2 = f2sqr(e)
    Xn = f2mul(a, f2sub(b, f))
    Yn = f2sub(f2sqr(g), f2add(f2add(e2, e2), e2))
    Zn = f2mul(b, h)
    return (Xn, Yn, Zn), (f2neg(h), f2add(f2add(j, j), j), i)


def add_step(Rp, Q):
    X, Y, Z = Rp
    theta = f2sub(Y, f2mul(Q[1], Z))
    lam = f2sub(X, f2mul(Q[0], Z))
    c, d = f2sqr(theta), f2sqr(lam)
    e = f2mul(lam, d)
    f = f2mul(Z, c)
    g = f2mul(X, d)
    h = f2sub(f2add(e, f), f2add(g, g))
    Xn = f2mul(lam, h)
    Yn = f2sub(f2mul(theta, f2sub(g, h)), f2mul(e, Y))
    Zn = f2mul(Z, e)
    j = f2sub(f2mul(theta, Q[0]), f2mul(lam, Q[1]))
    return (Xn, Yn, Zn), (lam, f2neg(theta), j)


def ell(f, coeffs, Pt):
    c0 = f2scal(coeffs[0], Pt[1])
    c1 = f2scal(coeffs[1], Pt[0])
    return f12mul(f, sparse(c0, c1, coeffs[2]))


def miller(Pt, Q):
    Rp = (Q[0], Q[1], F2_ONE)
    f = F12_ONE
    for i in range(O.ATE_LOOP.bit_length() - 2, -1, -1):
        f = f12sqr(f)
        Rp, co = dbl_step(Rp)
        f = ell(f, co, Pt)
        if (O.ATE_LOOP >> i) & 1:
            Rp, co = add_step(Rp, Q)
            f = ell(f, co, Pt)
    q1 = (f2mul(f2conj(Q[0]), TW_X), f2mul(f2conj(Q[1]), TW_Y))
    q2 = (f2mul(f2conj(q1[0]), TW_X), f2neg(f2mul(f2conj(q1[1]), TW_Y)))
    Rp, co = add_step(Rp, q1)
    f = ell(f, co, Pt)
    Rp, co = add_step(Rp, q2)
    f = ell(f, co, Pt)
    return f


HARD = (P ** 4 - P ** 2 + 1) // R


def final_exp(f):
    f1 = f12mul(f12conj(f), f12inv(f))          # f^(p^6 - 1)
    f2 = f12mul(f12frob2(f1), f1)               # ^(p^2 + 1)
    return f12pow(f2, HARD)


def to_poly(a):
    out = [0] * 12
    for j in range(2):
        for i in range(3):
            out = O._f12_add(out, O._f12_from_f2(a[j][i], 2 * i + j))
    return out


def limbs(v):
    return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


if __name__ == "__main__":
    # Frobenius^2 really is x -> x^(p^2)
    x = (((3, 5), (7, 11), (13, 17)), ((19, 23), (29, 31), (37, 41)))
    assert f12frob2(x) == f12pow(x, P * P)
    assert f12mul(x, f12inv(x)) == F12_ONE
    Pt, Q = O.G1.mul(O.G1_GEN, 12345), O.G2.mul(O.G2_GEN, 67890)
    mine = final_exp(miller(Pt, Q))
    ref = O.final_exp(O.miller_loop(Q, Pt))
    print("matches oracle pairing:", to_poly(mine) == ref)
    # bilinearity: e(aP, bQ) == e(P, Q)^(ab)
    e1 = final_exp(miller(O.G1_GEN, O.G2_GEN))
    assert mine == f12pow(e1, 12345 * 67890 % R)
    print("bilinear: True")
    Rm = 1 << 256
    print("// generated by tools/pairing_proto.py")
    for name, c in (("TW_X", TW_X), ("TW_Y", TW_Y), ("B_TWIST", B_TWIST)):
        print(f"{name}_C0 {limbs(c[0] * Rm % P)}\n{name}_C1 {limbs(c[1] * Rm % P)}")
    for k in range(1, 6):
        print(f"FROB2_{k} {limbs(G2C[k][0] * Rm % P)}")
    print("TWO_INV", limbs(TWO_INV * Rm % P))
    print("HARD bits", HARD.bit_length())
    print("HARD limbs", ", ".join("0x%08xu" % ((HARD >> (32 * i)) & 0xFFFFFFFF) for i in range((HARD.bit_length() + 31) // 32)))


# ---------------------------------------------------------------- fast hard part (BN addition chain, x > 0)
BN_X = O.BN_X
G1C = [f2pow(XI, (P - 1) * k // 6) for k in range(6)]          # Frobenius^1 coefficients (Fq2)
G3C = [f2pow(XI, (P ** 3 - 1) * k // 6) for k in range(6)]      # Frobenius^3 coefficients (Fq2)


def f12frob_odd(a, C):
    """x -> x^(p^k) for odd k: conjugate every Fq2 coefficient, scale the v^i w^j one by xi^((p^k-1)(2i+j)/6)."""
    return tuple(tuple(f2mul(f2conj(a[j][i]), C[2 * i + j]) for i in range(3)) for j in range(2))


def exp_by_neg_x(f):
    return f12conj(f12pow(f, BN_X))         # f^(-x) in the cyclotomic subgroup: inverse == conjugate


def hard_part_chain(r):
    y0 = exp_by_neg_x(r)
    y1 = f12sqr(y0)
    y2 = f12sqr(y1)
    y3 = f12mul(y2, y1)
    y4 = exp_by_neg_x(y3)
    y5 = f12sqr(y4)
    y6 = exp_by_neg_x(y5)
    y3 = f12conj(y3)
    y6 = f12conj(y6)
    y7 = f12mul(y6, y4)
    y8 = f12mul(y7, y3)
    y9 = f12mul(y8, y1)
    y10 = f12mul(y8, y4)
    y11 = f12mul(y10, r)
    y12 = f12frob_odd(y9, G1C)
    y13 = f12mul(y12, y11)
    y8 = f12frob2(y8)
    y14 = f12mul(y8, y13)
    r = f12conj(r)
    y15 = f12mul(r, y9)
    y15 = f12frob_odd(y15, G3C)
    return f12mul(y15, y14)


def proto_fast():
    x = (((3, 5), (7, 11), (13, 17)), ((19, 23), (29, 31), (37, 41)))
    assert f12frob_odd(x, G1C) == f12pow(x, P) and f12frob_odd(x, G3C) == f12pow(x, P ** 3)
    Pt, Q = O.G1.mul(O.G1_GEN, 12345), O.G2.mul(O.G2_GEN, 67890)
    f = miller(Pt, Q)
    f1 = f12mul(f12conj(f), f12inv(f))
    f2 = f12mul(f12frob2(f1), f1)
    slow = f12pow(f2, HARD)
    fast = hard_part_chain(f2)
    k = next((k for k in range(1, 40) if f12pow(slow, k) == fast), None)
    print("fast hard part == slow hard part ^", k)
    Rm = 1 << 256
    for name, C in (("FROB1", G1C), ("FROB3", G3C)):
        for i in range(1, 6):
            print(f"{name}_{i}_C0 {limbs(C[i][0] * Rm % P)}\n{name}_{i}_C1 {limbs(C[i][1] * Rm % P)}")
    print("BN_X", hex(BN_X), bin(BN_X).count("1"))
    # subgroup test used on the device: psi(P) == [6 x^2] P  for P in G2 and not for a point outside it
    psi = lambda T: (f2mul(f2conj(T[0]), TW_X), f2mul(f2conj(T[1]), TW_Y))
    assert psi(Q) == O.G2.mul(Q, 6 * BN_X * BN_X)
    # a curve point outside the r-torsion (the twist has a large cofactor): the psi test must reject it
    def f2sqrt(a):                      # p = 3 mod 4 (Adj / Rodriguez-Henriquez, Alg. 9)
        a1 = f2pow(a, (P - 3) // 4)
        alpha = f2mul(a1, f2mul(a1, a))
        x0 = f2mul(a1, a)
        if alpha == ((-1) % P, 0):
            x = f2mul((0, 1), x0)
        else:
            x = f2mul(f2pow(f2add(F2_ONE, alpha), (P - 1) // 2), x0)
        return x if f2sqr(x) == a else None
    for t in range(1, 50):
        xx = (t, 1)
        rhs = f2add(f2mul(f2sqr(xx), xx), B_TWIST)
        y = f2sqrt(rhs)
        if y is not None:
            T = (xx, y)
            assert O.G2.on_curve(T)
            in_sub = O.G2.mul(T, R) is None
            print("off-subgroup point found:", not in_sub, "psi test rejects it:", psi(T) != O.G2.mul(T, 6 * BN_X * BN_X))
            print("OFFSUB", O.g2_to_bytes(T).hex())
            break


if __name__ == "__main__":
    proto_fast()


# ---------------------------------------------------------------- Granger-Scott squaring in the cyclotomic subgroup
def f12cyclo_sqr(a):
    r0, r4, r3 = a[0]
    r2, r1, r5 = a[1]
    def sq(x, y):                 # (x + y*s)^2 over Fq2[s]/(s^2 - xi): returns (t0, t1)
        tmp = f2mul(x, y)
        t0 = f2sub(f2sub(f2mul(f2add(x, y), f2add(f2mul(XI, y), x)), tmp), f2mul(XI, tmp))
        return t0, f2add(tmp, tmp)
    t0, t1 = sq(r0, r1)
    t2, t3 = sq(r2, r3)
    t4, t5 = sq(r4, r5)
    three_minus = lambda t, z: f2add(f2add(f2sub(t, z), f2sub(t, z)), t)     # 3t - 2z
    three_plus = lambda t, z: f2add(f2add(f2add(t, z), f2add(t, z)), t)      # 3t + 2z
    z0 = three_minus(t0, r0)
    z1 = three_plus(t1, r1)
    z2 = three_plus(f2mul(XI, t5), r2)
    z3 = three_minus(t4, r3)
    z4 = three_minus(t2, r4)
    z5 = three_plus(t3, r5)
    return ((z0, z4, z3), (z2, z1, z5))


def proto_cyclo():
    Pt, Q = O.G1.mul(O.G1_GEN, 777), O.G2.mul(O.G2_GEN, 999)
    f = miller(Pt, Q)
    f1 = f12mul(f12conj(f), f12inv(f))
    g = f12mul(f12frob2(f1), f1)             # in the cyclotomic subgroup after the easy part
    print("cyclotomic squaring == squaring:", f12cyclo_sqr(g) == f12sqr(g), f12cyclo_sqr(f12sqr(g)) == f12sqr(f12sqr(g)))


if __name__ == "__main__":
    proto_cyclo()
