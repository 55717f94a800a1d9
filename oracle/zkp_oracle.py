"""CPU oracle (Python big-int) for the libzkp Groth16/BN254 prover path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``libzkp_b200/`` may import this; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg do.

PARITY UNPINNED: the reference (/root/reference, pure Rust) holds no golden
vectors for this path and its arithmetic lives in
un-vendored crates (ark-groth16/ark-poly/ark-ec/ark-ff/ark-serialize ^0.5,
Cargo.toml:15-27) that cannot be built here (no Rust toolchain).  This file
restates the published algorithms of those crates and anchors on the
reference's call sites:

  * MiMC-5 + constants ................ src/backend/snark.rs:182-211
  * fr_to_commitment .................. src/backend/snark.rs:214-221
  * mimc_hash_circuit ................. src/backend/snark.rs:232-247
  * EqualityCircuit ................... src/backend/snark.rs:255-291
  * MembershipCircuit ................. src/backend/snark.rs:505-585
  * prove_equality_zk / _membership_zk  src/backend/snark.rs:343-374, 405-452
  * verify public-input order ......... src/backend/snark.rs:398, 482-492
  * envelope .......................... src/proof/mod.rs:23-36
  * membership payload prefix ......... src/proof/set_membership.rs:29-34
  * [UPSTREAM] ark-groth16 prover: create_proof_with_reduction / calculate_coeff,
    LibsnarkReduction::{witness_map_from_matrices, instance_map_with_evaluation},
    generator::generate_parameters_with_qap; ark-poly Radix2EvaluationDomain;
    ark-serialize uncompressed SW-affine layout (flags in last byte).

What pins it instead (tests/test_oracle_*.py): the MiMC known answers of
SURVEY.md §8c, public BN254 constants (2-adic root shipped by arkworks, EIP-196
2*G1), naive O(n^2) polynomial evaluation vs the FFT, a trapdoor-exponent check
of every proof element, and an independent optimal-ate pairing verifier.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# Fields
# --------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
FR_GENERATOR = 5                     # ark-bn254 Fr::GENERATOR [UPSTREAM]
FR_TWO_ADICITY = 28
FR_TWO_ADIC_ROOT = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)
BN_X = 4965661367192848881
ATE_LOOP = 6 * BN_X + 2              # 29793968203157093288


def inv_mod(a: int, m: int) -> int:
    return pow(a, m - 2, m)


# Fq2 = Fq[u]/(u^2+1), elements are (c0, c1)
def f2_add(a, b): return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)
def f2_sub(a, b): return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)
def f2_neg(a): return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)
def f2_sqr(a): return f2_mul(a, a)
def f2_inv(a):
    d = inv_mod((a[0] * a[0] + a[1] * a[1]) % Q_MOD, Q_MOD)
    return (a[0] * d % Q_MOD, (-a[1]) * d % Q_MOD)
def f2_scalar(a, k): return (a[0] * k % Q_MOD, a[1] * k % Q_MOD)
F2_ZERO = (0, 0)
F2_ONE = (1, 0)


class _Fq:
    """Field-ops vtable so the curve code below is written once for G1 and G2."""
    zero = 0
    one = 1
    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def neg(a): return (-a) % Q_MOD
    @staticmethod
    def mul(a, b): return a * b % Q_MOD
    @staticmethod
    def sqr(a): return a * a % Q_MOD
    @staticmethod
    def inv(a): return inv_mod(a, Q_MOD)
    @staticmethod
    def is_zero(a): return a == 0


class _Fq2:
    zero = F2_ZERO
    one = F2_ONE
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    neg = staticmethod(f2_neg)
    mul = staticmethod(f2_mul)
    sqr = staticmethod(f2_sqr)
    inv = staticmethod(f2_inv)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0


# --------------------------------------------------------------------------
# Curves.  Affine points are (x, y) or None (= infinity); Jacobian (X, Y, Z).
# G1: y^2 = x^3 + 3 over Fq.   G2: y^2 = x^3 + 3/(9+u) over Fq2.
# --------------------------------------------------------------------------
G1_B = 3
G2_B = f2_mul((3, 0), f2_inv((9, 1)))
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


class Curve:
    def __init__(self, F, b, gen):
        self.F, self.b, self.gen = F, b, gen

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)) == F.zero

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    # --- Jacobian internals (a = 0 curves) ---
    def _jdbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd(self, P, Q):
        F = self.F
        if F.is_zero(P[2]): return Q
        if F.is_zero(Q[2]): return P
        Z1Z1 = F.sqr(P[2]); Z2Z2 = F.sqr(Q[2])
        U1 = F.mul(P[0], Z2Z2); U2 = F.mul(Q[0], Z1Z1)
        S1 = F.mul(F.mul(P[1], Q[2]), Z2Z2); S2 = F.mul(F.mul(Q[1], P[2]), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self._jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1); Rr = F.sub(S2, S1)
        HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(P[2], Q[2]), H)
        return (X3, Y3, Z3)

    def _to_j(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def _to_a(self, P):
        F = self.F
        if F.is_zero(P[2]):
            return None
        zi = F.inv(P[2]); zi2 = F.sqr(zi)
        return (F.mul(P[0], zi2), F.mul(P[1], F.mul(zi2, zi)))

    # --- affine API ---
    def add(self, P, Q):
        return self._to_a(self._jadd(self._to_j(P), self._to_j(Q)))

    def mul(self, P, k: int):
        """k*P with k taken as a non-negative integer (callers reduce mod r)."""
        if P is None or k == 0:
            return None
        acc = (self.F.one, self.F.one, self.F.zero)
        base = self._to_j(P)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == '1':
                acc = self._jadd(acc, base)
        return self._to_a(acc)

    def sum(self, pts):
        acc = (self.F.one, self.F.one, self.F.zero)
        for P in pts:
            acc = self._jadd(acc, self._to_j(P))
        return self._to_a(acc)

    def msm(self, bases, scalars):
        """Naive Sum scalars[i]*bases[i]; truncates to the shorter input like
        ark-ec's msm_bigint [UPSTREAM]."""
        n = min(len(bases), len(scalars))
        acc = (self.F.one, self.F.one, self.F.zero)
        for i in range(n):
            s = scalars[i] % R_MOD
            if s == 0 or bases[i] is None:
                continue
            acc = self._jadd(acc, self._to_j(self.mul(bases[i], s)))
        return self._to_a(acc)

    def fixed_base_table(self, P, window=8):
        """table[w][d] = d * 2^(window*w) * P, Jacobian, for fast trapdoor setup."""
        tbl = []
        base = self._to_j(P)
        for _ in range((256 + window - 1) // window):
            row = [(self.F.one, self.F.one, self.F.zero)]
            for _d in range(1, 1 << window):
                row.append(self._jadd(row[-1], base))
            tbl.append(row)
            base = self._jadd(row[-1], base)
        return tbl

    def fixed_base_mul(self, tbl, k: int, window=8):
        acc = (self.F.one, self.F.one, self.F.zero)
        w = 0
        while k:
            d = k & ((1 << window) - 1)
            if d:
                acc = self._jadd(acc, tbl[w][d])
            k >>= window
            w += 1
        return self._to_a(acc)


G1 = Curve(_Fq, G1_B, G1_GEN)
G2 = Curve(_Fq2, G2_B, G2_GEN)


# --------------------------------------------------------------------------
# MiMC-5 commitment (src/backend/snark.rs:182-221)
# --------------------------------------------------------------------------
MIMC_ROUNDS = 110
_MIMC_C: Optional[List[int]] = None


def mimc_constants() -> List[int]:
    """snark.rs:186-198 — c_i = SHA-256("libzkp_mimc_v1:" || u64_le(i)) as LE int mod r."""
    global _MIMC_C
    if _MIMC_C is None:
        _MIMC_C = [
            int.from_bytes(hashlib.sha256(b"libzkp_mimc_v1:" + i.to_bytes(8, "little")).digest(),
                           "little") % R_MOD
            for i in range(MIMC_ROUNDS)
        ]
    return _MIMC_C


def mimc_hash_native(value: int) -> int:
    """snark.rs:201-211."""
    x = value % R_MOD
    for c in mimc_constants():
        t = (x + c) % R_MOD
        t2 = t * t % R_MOD
        t4 = t2 * t2 % R_MOD
        x = t4 * t % R_MOD
    return x


def fr_to_bytes(f: int) -> bytes:
    """snark.rs:214-221 — canonical 32-byte LE."""
    return (f % R_MOD).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> Optional[int]:
    """snark.rs:224-229 — rejects len != 32 and non-canonical values (>= r)."""
    if len(b) != 32:
        return None
    v = int.from_bytes(b, "little")
    return v if v < R_MOD else None


def commit_value_snark(value: int) -> bytes:
    """src/utils/commitment.rs:14-16."""
    return fr_to_bytes(mimc_hash_native(value))


# --------------------------------------------------------------------------
# R1CS (ark-relations semantics restated; OptimizationGoal::Constraints => all
# symbolic LCs are inlined, so a constraint row is three sparse rows over
# z = instance || witness) [UPSTREAM]
# --------------------------------------------------------------------------
class LC(dict):
    """Sparse linear combination {column: coeff}; column c < n_inst is instance c
    (column 0 = the constant One), witness j is stored as ('w', j) until finalize."""

    def copy(self):
        return LC(self)

    def add_term(self, col, coeff):
        v = (self.get(col, 0) + coeff) % R_MOD
        if v:
            self[col] = v
        elif col in self:
            del self[col]
        return self

    def __add__(self, o):
        r = self.copy()
        for c, v in o.items():
            r.add_term(c, v)
        return r

    def __sub__(self, o):
        r = self.copy()
        for c, v in o.items():
            r.add_term(c, -v)
        return r

    def __neg__(self):
        return LC({c: (-v) % R_MOD for c, v in self.items()})


def lc_const(k: int) -> LC:
    return LC({0: k % R_MOD}) if k % R_MOD else LC()


class FpVar:
    """Value + LC pair, the shape of ark-r1cs-std's FpVar::Var after inlining."""
    __slots__ = ("val", "lc")

    def __init__(self, val, lc):
        self.val, self.lc = val % R_MOD, lc


class ConstraintSystem:
    def __init__(self):
        self.instance = [1]          # Variable::One
        self.witness: List[int] = []
        self.rows: List[Tuple[LC, LC, LC]] = []

    # allocation ------------------------------------------------------------
    def new_input(self, v: int) -> FpVar:
        self.instance.append(v % R_MOD)
        return FpVar(v, LC({len(self.instance) - 1: 1}))

    def new_witness(self, v: int) -> FpVar:
        self.witness.append(v % R_MOD)
        return FpVar(v, LC({("w", len(self.witness) - 1): 1}))

    def enforce(self, a: LC, b: LC, c: LC):
        self.rows.append((a, b, c))

    # gadgets (ark-r1cs-std AllocatedFp / AllocatedBool) [UPSTREAM] ------------
    def mul(self, x: FpVar, y: FpVar) -> FpVar:
        """AllocatedFp::mul — one new witness, one constraint x*y = p."""
        p = self.new_witness(x.val * y.val)
        self.enforce(x.lc, y.lc, p.lc)
        return p

    def enforce_equal(self, x: FpVar, y: FpVar):
        """AllocatedFp::conditional_enforce_equal with TRUE: (x - y) * 1 = 0."""
        self.enforce(x.lc - y.lc, lc_const(1), LC())

    def new_bool(self, bit: int, is_input: bool) -> FpVar:
        """AllocatedBool::new_variable: allocate, then (1 - a) * a = 0."""
        v = self.new_input(bit) if is_input else self.new_witness(bit)
        self.enforce(lc_const(1) - v.lc, v.lc, LC())
        return v

    # finalize ----------------------------------------------------------------
    def num_instance(self): return len(self.instance)
    def num_witness(self): return len(self.witness)
    def num_constraints(self): return len(self.rows)

    def assignment(self) -> List[int]:
        return self.instance + self.witness

    def matrices(self):
        """Three row-lists of [(coeff, column)] with witness j at n_inst + j
        (ark-relations to_matrices) [UPSTREAM]."""
        ni = len(self.instance)

        def conv(lc):
            return sorted(((v, (c if isinstance(c, int) else ni + c[1])) for c, v in lc.items()),
                          key=lambda t: t[1])
        A = [conv(r[0]) for r in self.rows]
        B = [conv(r[1]) for r in self.rows]
        C = [conv(r[2]) for r in self.rows]
        return A, B, C

    def is_satisfied(self) -> bool:
        z = self.assignment()
        A, B, C = self.matrices()
        ev = lambda row: sum(v * z[c] for v, c in row) % R_MOD
        return all(ev(a) * ev(b) % R_MOD == ev(c) for a, b, c in zip(A, B, C))


def fp_add_const(x: FpVar, k: int) -> FpVar:
    return FpVar(x.val + k, x.lc + lc_const(k))


def mimc_hash_circuit(cs: ConstraintSystem, x: FpVar, rounds: Optional[int] = None) -> FpVar:
    """snark.rs:232-247 — 3 constraints + 3 witnesses per round.
    rounds > 110 (synthetic config 4) cycles the 110 constants."""
    cst = mimc_constants()
    n = MIMC_ROUNDS if rounds is None else rounds
    for i in range(n):
        t = fp_add_const(x, cst[i % MIMC_ROUNDS])
        t2 = cs.mul(t, t)
        t4 = cs.mul(t2, t2)
        x = cs.mul(t4, t)
    return x


def mimc_chain_native(value: int, rounds: int) -> int:
    cst = mimc_constants()
    x = value % R_MOD
    for i in range(rounds):
        t = (x + cst[i % MIMC_ROUNDS]) % R_MOD
        x = pow(t, 5, R_MOD)
    return x


def equality_circuit(a: int, b: int, commitment: int, rounds: Optional[int] = None) -> ConstraintSystem:
    """snark.rs:263-290.  rounds=None is the reference circuit (110 rounds);
    rounds=R gives the synthetic MiMC-chain circuit of BASELINE config 4."""
    cs = ConstraintSystem()
    a_var = cs.new_witness(a)
    b_var = cs.new_witness(b)
    cs.enforce_equal(a_var, b_var)
    h = mimc_hash_circuit(cs, a_var, rounds)
    c_var = cs.new_input(commitment)
    cs.enforce_equal(h, c_var)
    return cs


MAX_SET_SIZE = 64  # snark.rs:503


def membership_circuit(value: int, sel: Sequence[int], set_values: Sequence[int],
                       is_real: Sequence[int], commitment: int) -> ConstraintSystem:
    """snark.rs:515-584 (slot count = len(set_values); the reference fixes 64)."""
    n = len(set_values)
    assert len(is_real) == n and len(sel) == n
    cs = ConstraintSystem()
    value_var = cs.new_witness(value)
    h = mimc_hash_circuit(cs, value_var)
    c_var = cs.new_input(commitment)
    cs.enforce_equal(h, c_var)
    set_vars = [cs.new_input(v) for v in set_values]
    real = [cs.new_bool(int(b), True) for b in is_real]
    sels = [cs.new_bool(int(b), False) for b in sel]
    sum_sel = FpVar(0, LC())
    for i in range(n):
        sum_sel = FpVar(sum_sel.val + sels[i].val, sum_sel.lc + sels[i].lc)
        one_minus = FpVar(1 - real[i].val, lc_const(1) - real[i].lc)
        prod = cs.mul(sels[i], one_minus)
        cs.enforce_equal(FpVar(0, LC()), prod)          # (c - v) * 1 = 0, c const
    cs.enforce_equal(FpVar(1, lc_const(1)), sum_sel)
    acc = FpVar(0, LC())
    for i in range(n):
        diff = FpVar(value_var.val - set_vars[i].val, value_var.lc - set_vars[i].lc)
        p = cs.mul(sels[i], diff)
        acc = FpVar(acc.val + p.val, acc.lc + p.lc)
    # FpVar::Var == FpVar::Constant takes the branch c.conditional_enforce_equal(v): (c - v) * 1 = 0
    cs.enforce_equal(FpVar(0, LC()), acc)
    return cs


def membership_inputs(value: int, set_: Sequence[int], slots: int = MAX_SET_SIZE):
    """snark.rs:406-427 — pad to `slots`, one-hot selector at first match."""
    if len(set_) == 0 or len(set_) > slots:
        return None
    if value not in set_:
        return None
    pos = list(set_).index(value)
    set_values = list(set_) + [0] * (slots - len(set_))
    is_real = [1] * len(set_) + [0] * (slots - len(set_))
    sel = [0] * slots
    sel[pos] = 1
    return sel, set_values, is_real


# --------------------------------------------------------------------------
# Radix-2 evaluation domain (ark-poly Radix2EvaluationDomain) [UPSTREAM]
# --------------------------------------------------------------------------
class Domain:
    def __init__(self, min_size: int):
        n = 1
        while n < min_size:
            n <<= 1
        self.n = n
        self.log_n = n.bit_length() - 1
        assert self.log_n <= FR_TWO_ADICITY
        self.omega = pow(FR_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - self.log_n), R_MOD)
        self.omega_inv = inv_mod(self.omega, R_MOD)
        self.n_inv = inv_mod(n, R_MOD)

    def _fft(self, a: List[int], w: int) -> List[int]:
        n = self.n
        a = list(a) + [0] * (n - len(a))
        j = 0
        for i in range(1, n):                       # bit reversal
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, R_MOD)
            half = length >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * wl % R_MOD
            for s in range(0, n, length):
                for k in range(half):
                    u = a[s + k]
                    v = a[s + k + half] * tw[k] % R_MOD
                    a[s + k] = (u + v) % R_MOD
                    a[s + k + half] = (u - v) % R_MOD
            length <<= 1
        return a

    def fft(self, a): return self._fft(a, self.omega)

    def ifft(self, a):
        return [x * self.n_inv % R_MOD for x in self._fft(a, self.omega_inv)]

    def coset_fft(self, a, g=FR_GENERATOR):
        p = 1
        out = []
        for x in list(a) + [0] * (self.n - len(a)):
            out.append(x * p % R_MOD)
            p = p * g % R_MOD
        return self.fft(out)

    def coset_ifft(self, a, g=FR_GENERATOR):
        gi = inv_mod(g, R_MOD)
        p = 1
        out = []
        for x in self.ifft(a):
            out.append(x * p % R_MOD)
            p = p * gi % R_MOD
        return out

    def vanishing_at(self, t: int) -> int:
        return (pow(t, self.n, R_MOD) - 1) % R_MOD

    def lagrange_at(self, t: int) -> List[int]:
        """evaluate_all_lagrange_coefficients(t) [UPSTREAM]."""
        n = self.n
        zt = self.vanishing_at(t)
        if zt == 0:
            out = [0] * n
            w = 1
            for i in range(n):
                if w == t % R_MOD:
                    out[i] = 1
                w = w * self.omega % R_MOD
            return out
        out = []
        w = 1
        c = zt * self.n_inv % R_MOD
        for _ in range(n):
            out.append(c * w % R_MOD * inv_mod((t - w) % R_MOD, R_MOD) % R_MOD)
            w = w * self.omega % R_MOD
        return out


def _eval_row(row, z):
    return sum(v * z[c] for v, c in row) % R_MOD


def witness_map(A, B, C, n_inst: int, z: Sequence[int]) -> List[int]:
    """LibsnarkReduction::witness_map_from_matrices [UPSTREAM].
    Returns h[0..n) (h[n-1] == 0 for a satisfying assignment)."""
    m = len(A)
    dom = Domain(m + n_inst)
    n = dom.n
    a = [0] * n
    b = [0] * n
    c = [0] * n
    for i in range(m):
        a[i] = _eval_row(A[i], z)
        b[i] = _eval_row(B[i], z)
        c[i] = _eval_row(C[i], z)
    for j in range(n_inst):
        a[m + j] = z[j] % R_MOD
    a = dom.coset_fft(dom.ifft(a))
    b = dom.coset_fft(dom.ifft(b))
    c = dom.coset_fft(dom.ifft(c))
    zinv = inv_mod(dom.vanishing_at(FR_GENERATOR), R_MOD)
    ab = [((a[i] * b[i] - c[i]) % R_MOD) * zinv % R_MOD for i in range(n)]
    return dom.coset_ifft(ab)


# --------------------------------------------------------------------------
# Deterministic PRNG for synthetic inputs: SplitMix64
# --------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def next_fr(self) -> int:
        """4 words LE, mask to 254 bits, reject >= r."""
        while True:
            v = 0
            for i in range(4):
                v |= self.next_u64() << (64 * i)
            v &= (1 << 254) - 1
            if v < R_MOD:
                return v


# --------------------------------------------------------------------------
# Groth16 keys, setup, prover (ark-groth16 ^0.5) [UPSTREAM]
# --------------------------------------------------------------------------
class VerifyingKey:
    def __init__(self, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1):
        self.alpha_g1, self.beta_g2, self.gamma_g2 = alpha_g1, beta_g2, gamma_g2
        self.delta_g2, self.gamma_abc_g1 = delta_g2, gamma_abc_g1


class ProvingKey:
    def __init__(self, vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query):
        self.vk, self.beta_g1, self.delta_g1 = vk, beta_g1, delta_g1
        self.a_query, self.b_g1_query, self.b_g2_query = a_query, b_g1_query, b_g2_query
        self.h_query, self.l_query = h_query, l_query


class Trapdoor:
    def __init__(self, alpha, beta, gamma, delta, tau):
        self.alpha, self.beta, self.gamma, self.delta, self.tau = alpha, beta, gamma, delta, tau

    @staticmethod
    def from_seed(seed: int) -> "Trapdoor":
        rng = SplitMix64(seed)
        return Trapdoor(*(rng.next_fr() for _ in range(5)))


def qap_at_tau(A, B, C, n_inst: int, n_vars: int, tau: int):
    """LibsnarkReduction::instance_map_with_evaluation [UPSTREAM]:
    a_j(tau), b_j(tau), c_j(tau) for every variable j, Z(tau), domain size."""
    m = len(A)
    dom = Domain(m + n_inst)
    u = dom.lagrange_at(tau)
    a = [0] * n_vars
    b = [0] * n_vars
    c = [0] * n_vars
    for j in range(n_inst):
        a[j] = u[m + j]
    for i in range(m):
        for v, col in A[i]:
            a[col] = (a[col] + u[i] * v) % R_MOD
        for v, col in B[i]:
            b[col] = (b[col] + u[i] * v) % R_MOD
        for v, col in C[i]:
            c[col] = (c[col] + u[i] * v) % R_MOD
    return a, b, c, dom.vanishing_at(tau), dom.n


def setup(cs: ConstraintSystem, td: Trapdoor, g1=G1_GEN, g2=G2_GEN) -> ProvingKey:
    """generate_parameters_with_qap [UPSTREAM] with the toxic waste as explicit
    input (the reference draws it from OsRng, snark.rs:310,331) and the standard
    generators (arkworks draws random ones)."""
    A, B, C = cs.matrices()
    ni, nv = cs.num_instance(), cs.num_instance() + cs.num_witness()
    a, b, c, zt, n = qap_at_tau(A, B, C, ni, nv, td.tau)
    gi = inv_mod(td.gamma, R_MOD)
    di = inv_mod(td.delta, R_MOD)
    t1 = G1.fixed_base_table(g1)
    t2 = G2.fixed_base_table(g2)
    m1 = lambda k: G1.fixed_base_mul(t1, k % R_MOD)
    m2 = lambda k: G2.fixed_base_mul(t2, k % R_MOD)
    gamma_abc = [m1((td.beta * a[j] + td.alpha * b[j] + c[j]) * gi) for j in range(ni)]
    l_query = [m1((td.beta * a[j] + td.alpha * b[j] + c[j]) * di) for j in range(ni, nv)]
    h_query = []
    p = zt * di % R_MOD
    for _ in range(n - 1):
        h_query.append(m1(p))
        p = p * td.tau % R_MOD
    vk = VerifyingKey(m1(td.alpha), m2(td.beta), m2(td.gamma), m2(td.delta), gamma_abc)
    return ProvingKey(vk, m1(td.beta), m1(td.delta),
                      [m1(x) for x in a], [m1(x) for x in b], [m2(x) for x in b],
                      h_query, l_query)


def prove_with_assignment(pk: ProvingKey, r: int, s: int, h: Sequence[int],
                          z: Sequence[int], n_inst: int):
    """create_proof_with_assignment + calculate_coeff [UPSTREAM]."""
    r %= R_MOD
    s %= R_MOD
    aux = z[n_inst:]
    h_acc = G1.msm(pk.h_query, h)
    l_acc = G1.msm(pk.l_query, aux)
    rs_delta = G1.mul(pk.delta_g1, r * s % R_MOD)

    def coeff(curve, initial, query, vk_param):
        acc = curve.msm(query[1:], z[1:])
        return curve.sum([initial, query[0], acc, vk_param])

    g_a = coeff(G1, G1.mul(pk.delta_g1, r), pk.a_query, pk.vk.alpha_g1)
    g1_b = coeff(G1, G1.mul(pk.delta_g1, s), pk.b_g1_query, pk.beta_g1) if r != 0 else None
    g2_b = coeff(G2, G2.mul(pk.vk.delta_g2, s), pk.b_g2_query, pk.vk.beta_g2)
    g_c = G1.sum([G1.mul(g_a, s), G1.mul(g1_b, r), G1.neg(rs_delta), l_acc, h_acc])
    return g_a, g2_b, g_c


def prove(pk: ProvingKey, cs: ConstraintSystem, r: int, s: int):
    """Groth16::create_proof_with_reduction(circuit, pk, r, s) [UPSTREAM]."""
    A, B, C = cs.matrices()
    z = cs.assignment()
    h = witness_map(A, B, C, cs.num_instance(), z)
    return prove_with_assignment(pk, r, s, h, z, cs.num_instance())


# --------------------------------------------------------------------------
# ark-serialize uncompressed layout [UPSTREAM]
# --------------------------------------------------------------------------
def _fq_bytes(x: int, flags: int = 0) -> bytes:
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= flags
    return bytes(b)


def g1_to_bytes(P) -> bytes:
    if P is None:
        return bytes(32) + _fq_bytes(0, 0x40)
    x, y = P
    neg = y > (Q_MOD - y) % Q_MOD
    return _fq_bytes(x) + _fq_bytes(y, 0x80 if neg else 0)


def _f2_gt(a, b) -> bool:
    return (a[1], a[0]) > (b[1], b[0])          # c1 first, then c0


def g2_to_bytes(P) -> bytes:
    if P is None:
        return bytes(96) + _fq_bytes(0, 0x40)
    x, y = P
    neg = _f2_gt(y, f2_neg(y))
    return _fq_bytes(x[0]) + _fq_bytes(x[1]) + _fq_bytes(y[0]) + _fq_bytes(y[1], 0x80 if neg else 0)


def _fq_from(b: bytes, with_flags: bool):
    v = int.from_bytes(b, "little")
    flags = 0
    if with_flags:
        flags = b[31] & 0xC0
        v &= (1 << 254) - 1
    if v >= Q_MOD:
        raise ValueError("non-canonical Fq")
    return v, flags


def g1_from_bytes(b: bytes, validate=True):
    x, _ = _fq_from(b[0:32], False)
    y, fl = _fq_from(b[32:64], True)
    if fl & 0x40:
        return None
    P = (x, y)
    if validate and not G1.on_curve(P):
        raise ValueError("G1 point not on curve")
    return P


def g2_from_bytes(b: bytes, validate=True):
    x0, _ = _fq_from(b[0:32], False)
    x1, _ = _fq_from(b[32:64], False)
    y0, _ = _fq_from(b[64:96], False)
    y1, fl = _fq_from(b[96:128], True)
    if fl & 0x40:
        return None
    P = ((x0, x1), (y0, y1))
    if validate:
        if not G2.on_curve(P):
            raise ValueError("G2 point not on curve")
        if G2.mul(P, R_MOD) is not None:
            raise ValueError("G2 point not in the r-torsion subgroup")
    return P


def proof_to_bytes(proof) -> bytes:
    """Proof<Bn254> = A(G1) || B(G2) || C(G1) = 256 B (snark.rs:369-373)."""
    a, b, c = proof
    return g1_to_bytes(a) + g2_to_bytes(b) + g1_to_bytes(c)


def proof_from_bytes(b: bytes, validate=True):
    if len(b) != 256:
        raise ValueError("bad proof length")
    return (g1_from_bytes(b[0:64], validate), g2_from_bytes(b[64:192], validate),
            g1_from_bytes(b[192:256], validate))


def _vec(items, f) -> bytes:
    return len(items).to_bytes(8, "little") + b"".join(f(p) for p in items)


def vk_to_bytes(vk: VerifyingKey) -> bytes:
    return (g1_to_bytes(vk.alpha_g1) + g2_to_bytes(vk.beta_g2) + g2_to_bytes(vk.gamma_g2)
            + g2_to_bytes(vk.delta_g2) + _vec(vk.gamma_abc_g1, g1_to_bytes))


def pk_to_bytes(pk: ProvingKey) -> bytes:
    """ProvingKey<Bn254>::serialize_uncompressed field order (snark.rs:97-101)."""
    return (vk_to_bytes(pk.vk) + g1_to_bytes(pk.beta_g1) + g1_to_bytes(pk.delta_g1)
            + _vec(pk.a_query, g1_to_bytes) + _vec(pk.b_g1_query, g1_to_bytes)
            + _vec(pk.b_g2_query, g2_to_bytes) + _vec(pk.h_query, g1_to_bytes)
            + _vec(pk.l_query, g1_to_bytes))


class _Reader:
    def __init__(self, b): self.b, self.o = b, 0
    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated")
        r = self.b[self.o:self.o + n]; self.o += n
        return r
    def g1(self, v): return g1_from_bytes(self.take(64), v)
    def g2(self, v): return g2_from_bytes(self.take(128), v)
    def vec(self, f, v):
        n = int.from_bytes(self.take(8), "little")
        return [f(v) for _ in range(n)]


def _read_vk(rd: _Reader, validate) -> VerifyingKey:
    return VerifyingKey(rd.g1(validate), rd.g2(validate), rd.g2(validate), rd.g2(validate),
                        rd.vec(rd.g1, validate))


def vk_from_bytes(b: bytes, validate=False) -> VerifyingKey:
    return _read_vk(_Reader(b), validate)


def pk_from_bytes(b: bytes, validate=False) -> ProvingKey:
    rd = _Reader(b)
    vk = _read_vk(rd, validate)
    pk = ProvingKey(vk, rd.g1(validate), rd.g1(validate), rd.vec(rd.g1, validate),
                    rd.vec(rd.g1, validate), rd.vec(rd.g2, validate), rd.vec(rd.g1, validate),
                    rd.vec(rd.g1, validate))
    if rd.o != len(b):
        raise ValueError("trailing bytes")
    return pk


# --------------------------------------------------------------------------
# libzkp framing (src/proof/mod.rs:23-36, set_membership.rs:29-34)
# --------------------------------------------------------------------------
def envelope(scheme: int, proof: bytes, commitment: bytes) -> bytes:
    return (bytes([2, scheme]) + len(proof).to_bytes(4, "little")
            + len(commitment).to_bytes(4, "little") + proof + commitment)


def membership_payload(set_: Sequence[int], snark_proof: bytes) -> bytes:
    return (len(set_).to_bytes(4, "little") + b"".join(int(v).to_bytes(8, "little") for v in set_)
            + snark_proof)


# --------------------------------------------------------------------------
# Independent checks: trapdoor-exponent check and pairing verifier
# --------------------------------------------------------------------------
def trapdoor_expected_proof(cs: ConstraintSystem, td: Trapdoor, r: int, s: int, g1=G1_GEN, g2=G2_GEN):
    """Recompute A, B, C as single scalar multiplications of the generators from
    the toxic waste — independent of the MSM / pk / serialization code."""
    A, B, C = cs.matrices()
    z = cs.assignment()
    ni, nv = cs.num_instance(), len(z)
    a, b, c, zt, n = qap_at_tau(A, B, C, ni, nv, td.tau)
    h = witness_map(A, B, C, ni, z)
    di = inv_mod(td.delta, R_MOD)
    az = sum(z[j] * a[j] for j in range(nv)) % R_MOD
    bz = sum(z[j] * b[j] for j in range(nv)) % R_MOD
    ea = (td.alpha + az + r * td.delta) % R_MOD
    eb = (td.beta + bz + s * td.delta) % R_MOD
    lz = sum(z[j] * (td.beta * a[j] + td.alpha * b[j] + c[j]) for j in range(ni, nv)) % R_MOD * di % R_MOD
    ht = 0
    p = zt * di % R_MOD
    for i in range(n - 1):
        ht = (ht + h[i] * p) % R_MOD
        p = p * td.tau % R_MOD
    ec = (s * ea + r * eb - r * s % R_MOD * td.delta + lz + ht) % R_MOD
    return G1.mul(g1, ea), G2.mul(g2, eb), G1.mul(g1, ec)


# ---- Fq12 as Fq[w]/(w^12 - 18 w^6 + 82); w^6 = 9 + u -----------------------
_F12_ONE = [1] + [0] * 11


def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):
        v = t[k]
        if v:
            t[k - 6] += 18 * v
            t[k - 12] -= 82 * v
    return [x % Q_MOD for x in t[:12]]


def f12_pow(a, e: int):
    r = _F12_ONE
    for bit in bin(e)[2:]:
        r = f12_mul(r, r)
        if bit == '1':
            r = f12_mul(r, a)
    return r


def _f12_inv_deg(a):
    # polynomial extended Euclid over Fq on degree-12 modulus
    def deg(p):
        d = len(p) - 1
        while d and p[d] == 0:
            d -= 1
        return d

    def pdiv(a_, b_):
        da, db = deg(a_), deg(b_)
        t = list(a_)
        o = [0] * len(a_)
        inv = inv_mod(b_[db], Q_MOD)
        for i in range(da - db, -1, -1):
            o[i] = t[db + i] * inv % Q_MOD
            for c in range(db + 1):
                t[c + i] = (t[c + i] - o[i] * b_[c]) % Q_MOD
        return o[: deg(o) + 1]

    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, (-18) % Q_MOD, 0, 0, 0, 0, 0, 1]
    while deg(low):
        r = pdiv(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q_MOD
                new[i + j] = (new[i + j] - low[i] * r[j]) % Q_MOD
        lm, low, hm, high = nm, new, lm, low
    li = inv_mod(low[0], Q_MOD)
    return [x * li % Q_MOD for x in lm[:12]]


def _f12_from_f2(c, shift: int):
    """Embed c0 + c1*u (u = w^6 - 9) into Fq12 and multiply by w^shift."""
    e = [0] * 12
    e[0] = (c[0] - 9 * c[1]) % Q_MOD
    e[6] = c[1] % Q_MOD
    sh = [0] * 12
    sh[shift] = 1
    return f12_mul(e, sh)


def _twist(Q):
    return (_f12_from_f2(Q[0], 2), _f12_from_f2(Q[1], 3))


def _f12_add(a, b): return [(x + y) % Q_MOD for x, y in zip(a, b)]
def _f12_sub(a, b): return [(x - y) % Q_MOD for x, y in zip(a, b)]
def _f12_scal(a, k): return [x * k % Q_MOD for x in a]


def _aff12_double(P):
    x, y = P
    m = f12_mul(_f12_scal(f12_mul(x, x), 3), _f12_inv_deg(_f12_scal(y, 2)))
    nx = _f12_sub(f12_mul(m, m), _f12_scal(x, 2))
    ny = _f12_sub(f12_mul(m, _f12_sub(x, nx)), y)
    return (nx, ny), m


def _aff12_add(P, Q):
    (x1, y1), (x2, y2) = P, Q
    m = f12_mul(_f12_sub(y2, y1), _f12_inv_deg(_f12_sub(x2, x1)))
    nx = _f12_sub(_f12_sub(f12_mul(m, m), x1), x2)
    ny = _f12_sub(f12_mul(m, _f12_sub(x1, nx)), y1)
    return (nx, ny), m


def _line(m, P1, T):
    # m*(xt - x1) - (yt - y1)
    return _f12_sub(f12_mul(m, _f12_sub(T[0], P1[0])), _f12_sub(T[1], P1[1]))


def _frob12(a, k=1):
    return f12_pow(a, Q_MOD ** k)


def miller_loop(Q2, P1):
    """Optimal-ate Miller loop on BN254, textbook affine form over Fq12."""
    if Q2 is None or P1 is None:
        return _F12_ONE
    Q = _twist(Q2)
    T = ([P1[0]] + [0] * 11, [P1[1]] + [0] * 11)
    R = Q
    f = _F12_ONE
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        R2, m = _aff12_double(R)
        f = f12_mul(f12_mul(f, f), _line(m, R, T))
        R = R2
        if (ATE_LOOP >> i) & 1:
            Rn, m = _aff12_add(R, Q)
            f = f12_mul(f, _line(m, R, T))
            R = Rn
    Q1 = (_frob12(Q[0]), _frob12(Q[1]))
    nQ2 = (_frob12(Q1[0]), [(-x) % Q_MOD for x in _frob12(Q1[1])])
    Rn, m = _aff12_add(R, Q1)
    f = f12_mul(f, _line(m, R, T))
    R = Rn
    _, m = _aff12_add(R, nQ2)
    f = f12_mul(f, _line(m, R, T))
    return f


def final_exp(f):
    return f12_pow(f, (Q_MOD ** 12 - 1) // R_MOD)


def pairing_product_is_one(pairs) -> bool:
    f = _F12_ONE
    for P1, Q2 in pairs:
        f = f12_mul(f, miller_loop(Q2, P1))
    return final_exp(f) == _F12_ONE


def verify(vk: VerifyingKey, public_inputs: Sequence[int], proof) -> bool:
    """Groth16 check e(A,B) = e(alpha,beta) e(sum x_i gamma_abc_i, gamma) e(C,delta)
    (what verify_with_processed_vk decides, snark.rs:400,494)."""
    a, b, c = proof
    if len(public_inputs) + 1 != len(vk.gamma_abc_g1):
        return False
    if not (G1.on_curve(a) and G2.on_curve(b) and G1.on_curve(c)):
        return False
    acc = G1.sum([vk.gamma_abc_g1[0]] + [G1.mul(g, x % R_MOD)
                                         for g, x in zip(vk.gamma_abc_g1[1:], public_inputs)])
    return pairing_product_is_one([
        (G1.neg(a), b), (vk.alpha_g1, vk.beta_g2), (acc, vk.gamma_g2), (c, vk.delta_g2)])


def equality_public_inputs(commitment: int):
    return [commitment]                                        # snark.rs:398


def membership_public_inputs(commitment: int, set_: Sequence[int], slots: int = MAX_SET_SIZE):
    return ([commitment] + [(set_[i] if i < len(set_) else 0) for i in range(slots)]
            + [(1 if i < len(set_) else 0) for i in range(slots)])   # snark.rs:482-492
