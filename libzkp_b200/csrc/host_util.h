// host_util.h — host-side helpers of the engine: SHA-256 (MiMC constants), 256-bit
// compares, ark-serialize point parsing, and the native R1CS synthesis of libzkp's two
// circuits.  Product code (not the oracle): it may not include anything under oracle/.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "field.cuh"

namespace lzkp {
namespace host {

// ---- SHA-256 (FIPS 180-4), one-shot --------------------------------------------------
inline void sha256(const uint8_t *msg, size_t len, uint8_t out[32]) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
        0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
        0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
        0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
        0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
        0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
        0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
        0xc67178f2};
    auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<uint8_t> m(msg, msg + len);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(bits >> (8 * i)));
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)m[off + 4 * i] << 24) | ((uint32_t)m[off + 4 * i + 1] << 16) |
                   ((uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
}

// ---- field helpers on the host (field.cuh's portable bodies) ---------------------------
template <class F>
inline bool lt_modulus(const F &a) {
    uint32_t d[8];
    F m = F::modulus();
    return sub8(d, a.l, m.l) != 0;
}
// MiMC round constant i (snark.rs:186-198), Montgomery form
inline Fr mimc_constant(uint32_t i) {
    uint8_t msg[23], dig[32];
    memcpy(msg, "libzkp_mimc_v1:", 15);
    uint64_t le = i;
    memcpy(msg + 15, &le, 8);
    sha256(msg, 23, dig);
    Fr v;
    memcpy(v.l, dig, 32);
    while (!lt_modulus(v)) {        // from_le_bytes_mod_order: value < 2^256 < 6r
        uint32_t d[8];
        Fr m = Fr::modulus();
        sub8(d, v.l, m.l);
        memcpy(v.l, d, 32);
    }
    return Fr::from_canonical(v);
}
inline Fr fr_from_u64(uint64_t x) {
    Fr c = Fr::zero();
    c.l[0] = (uint32_t)x;
    c.l[1] = (uint32_t)(x >> 32);
    return Fr::from_canonical(c);
}
// mimc_hash_native (snark.rs:201-211) -> canonical bytes (fr_to_commitment, snark.rs:214-221)
inline void commit_value_snark(uint64_t value, uint8_t out[32]) {
    static Fr cst[110];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 110; i++) cst[i] = mimc_constant(i);
        init = true;
    }
    Fr x = fr_from_u64(value);
    for (int i = 0; i < 110; i++) {
        Fr t = x + cst[i], t2 = t.sqr(), t4 = t2.sqr();
        x = t4 * t;
    }
    Fr c = x.to_canonical();
    memcpy(out, c.l, 32);
}

// ---- ark-serialize point parsing --------------------------------------------------------
// Reads one 32-byte base-field element; `flags` (may be null) receives the top two bits.
// Returns false for a non-canonical value.
inline bool read_fq(const uint8_t *b, Fq &out, uint8_t *flags) {
    uint8_t t[32];
    memcpy(t, b, 32);
    if (flags) { *flags = t[31] & 0xC0; t[31] &= 0x3F; }
    memcpy(out.l, t, 32);
    return lt_modulus(out);
}
struct G1Canon { Fq x, y; };            // canonical limbs; (0,0) = infinity
struct G2Canon { Fq x0, x1, y0, y1; };
inline bool read_g1(const uint8_t *b, G1Canon &p) {
    uint8_t fl = 0;
    if (!read_fq(b, p.x, nullptr) || !read_fq(b + 32, p.y, &fl)) return false;
    if (fl & 0x40) { p.x = Fq::zero(); p.y = Fq::zero(); }
    return true;
}
inline bool read_g2(const uint8_t *b, G2Canon &p) {
    uint8_t fl = 0;
    if (!read_fq(b, p.x0, nullptr) || !read_fq(b + 32, p.x1, nullptr) || !read_fq(b + 64, p.y0, nullptr) ||
        !read_fq(b + 96, p.y1, &fl))
        return false;
    if (fl & 0x40) { p.x0 = p.x1 = p.y0 = p.y1 = Fq::zero(); }
    return true;
}
inline bool is_inf(const G1Canon &p) { return p.x.is_zero() && p.y.is_zero(); }
inline bool is_inf(const G2Canon &p) { return p.x0.is_zero() && p.x1.is_zero() && p.y0.is_zero() && p.y1.is_zero(); }

// ---- R1CS of the two libzkp circuits -------------------------------------------------------
struct Csr {
    std::vector<uint32_t> rowptr{0};
    std::vector<uint32_t> col;
    std::vector<Fr> val;                 // canonical
    void push_row(std::vector<std::pair<uint32_t, Fr>> terms) {   // terms: (column, Montgomery coeff)
        std::sort(terms.begin(), terms.end(), [](auto &a, auto &b) { return a.first < b.first; });
        for (auto &t : terms) {
            if (t.second.is_zero()) continue;
            col.push_back(t.first);
            val.push_back(t.second.to_canonical());
        }
        rowptr.push_back((uint32_t)col.size());
    }
};
struct R1cs {
    uint32_t m = 0, n_inst = 0, n_wit = 0;
    Csr A, B, C;
    void row(std::vector<std::pair<uint32_t, Fr>> a, std::vector<std::pair<uint32_t, Fr>> b,
             std::vector<std::pair<uint32_t, Fr>> c) {
        A.push_row(std::move(a)); B.push_row(std::move(b)); C.push_row(std::move(c));
        m++;
    }
};
// mimc_hash_circuit (snark.rs:232-247): 3 rows per round on wires starting at column `w`;
// x0 is the input column.  Returns the column of the hash output.
inline uint32_t synth_mimc(R1cs &cs, uint32_t x0, uint32_t w, uint32_t rounds) {
    const Fr one = Fr::one();
    uint32_t x = x0;
    for (uint32_t i = 0; i < rounds; i++) {
        Fr c = mimc_constant(i % 110);
        uint32_t t2 = w + 3 * i, t4 = t2 + 1, x5 = t2 + 2;
        cs.row({{0, c}, {x, one}}, {{0, c}, {x, one}}, {{t2, one}});
        cs.row({{t2, one}}, {{t2, one}}, {{t4, one}});
        cs.row({{t4, one}}, {{0, c}, {x, one}}, {{x5, one}});
        x = x5;
    }
    return x;
}
// EqualityCircuit (snark.rs:263-290); rounds != 110 gives the synthetic MiMC-chain circuit.
inline R1cs synth_equality(uint32_t rounds) {
    R1cs cs;
    cs.n_inst = 2;                       // One, commitment
    cs.n_wit = 2 + 3 * rounds;           // a, b, wires
    const Fr one = Fr::one(), mone = Fr::one().neg();
    const uint32_t a = 2, b = 3;
    cs.row({{a, one}, {b, mone}}, {{0, one}}, {});                       // a == b
    uint32_t h = synth_mimc(cs, a, 4, rounds);
    cs.row({{h, one}, {1, mone}}, {{0, one}}, {});                       // hash == commitment
    return cs;
}
// MembershipCircuit (snark.rs:515-584) with S set slots (the reference fixes S = 64).
inline R1cs synth_membership(uint32_t S) {
    R1cs cs;
    cs.n_inst = 2 + 2 * S;               // One, commitment, set[S], is_real[S]
    cs.n_wit = 1 + 330 + 3 * S;          // value, MiMC wires, sel[S], sel*(1-real)[S], sel*(value-set)[S]
    const Fr one = Fr::one(), mone = Fr::one().neg();
    const uint32_t set0 = 2, real0 = 2 + S, value = 2 + 2 * S, wires = value + 1, sel0 = wires + 330, p10 = sel0 + S,
                   p20 = p10 + S;
    uint32_t h = synth_mimc(cs, value, wires, 110);
    cs.row({{h, one}, {1, mone}}, {{0, one}}, {});
    for (uint32_t i = 0; i < S; i++) cs.row({{0, one}, {real0 + i, mone}}, {{real0 + i, one}}, {});   // Boolean input
    for (uint32_t i = 0; i < S; i++) cs.row({{0, one}, {sel0 + i, mone}}, {{sel0 + i, one}}, {});     // Boolean witness
    for (uint32_t i = 0; i < S; i++) {
        cs.row({{sel0 + i, one}}, {{0, one}, {real0 + i, mone}}, {{p10 + i, one}});   // sel * (1 - is_real)
        cs.row({{p10 + i, mone}}, {{0, one}}, {});                                     // == 0
    }
    {
        std::vector<std::pair<uint32_t, Fr>> a{{0, one}};
        for (uint32_t i = 0; i < S; i++) a.push_back({sel0 + i, mone});
        cs.row(a, {{0, one}}, {});                                                     // sum sel == 1
    }
    for (uint32_t i = 0; i < S; i++)
        cs.row({{sel0 + i, one}}, {{value, one}, {set0 + i, mone}}, {{p20 + i, one}}); // sel * (value - set)
    {
        std::vector<std::pair<uint32_t, Fr>> a;
        for (uint32_t i = 0; i < S; i++) a.push_back({p20 + i, mone});                 // Var == Constant(0): (0 - sum) * 1 = 0
        cs.row(a, {{0, one}}, {});                                                     // sum == 0
    }
    return cs;
}

// ---- full assignments z = instance || witness of the builtin circuits (canonical limbs) ------------
// Same values generate_constraints assigns (snark.rs:263-290, 515-584); used by callers that prove from
// an explicit z (lzkp_prove_batch) and by the large synthetic MiMC-chain circuit.
inline Fr canon_u64(uint64_t x) {
    Fr c = Fr::zero();
    c.l[0] = (uint32_t)x;
    c.l[1] = (uint32_t)(x >> 32);
    return c;
}
// appends t^2, t^4, t^5 per round to z; returns the chain output (Montgomery)
inline Fr assign_mimc(std::vector<Fr> &z, uint64_t input, uint32_t rounds) {
    static Fr cst[110];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 110; i++) cst[i] = mimc_constant(i);
        init = true;
    }
    Fr x = fr_from_u64(input);
    for (uint32_t i = 0; i < rounds; i++) {
        Fr t = x + cst[i % 110], t2 = t.sqr(), t4 = t2.sqr();
        x = t4 * t;
        z.push_back(t2.to_canonical()); z.push_back(t4.to_canonical()); z.push_back(x.to_canonical());
    }
    return x;
}
inline std::vector<Fr> assign_equality(uint32_t rounds, uint64_t a, uint64_t b, const uint8_t *commitment) {
    std::vector<Fr> z;
    z.reserve(4 + 3 * (size_t)rounds);
    z.push_back(canon_u64(1));
    z.push_back(Fr::zero());
    z.push_back(canon_u64(a));
    z.push_back(canon_u64(b));
    Fr h = assign_mimc(z, a, rounds);
    if (commitment) memcpy(z[1].l, commitment, 32);
    else z[1] = h.to_canonical();
    return z;
}
// returns an empty vector when the set is empty / too long / does not contain value (snark.rs:406,415-418)
inline std::vector<Fr> assign_membership(uint32_t S, uint64_t value, const uint64_t *set, uint32_t len,
                                         const uint8_t *commitment) {
    std::vector<Fr> z;
    if (len < 1 || len > S) return z;
    uint32_t pos = S;
    for (uint32_t i = 0; i < len; i++) if (set[i] == value) { pos = i; break; }
    if (pos == S) return z;
    z.push_back(canon_u64(1));
    z.push_back(Fr::zero());
    for (uint32_t i = 0; i < S; i++) z.push_back(i < len ? canon_u64(set[i]) : Fr::zero());
    for (uint32_t i = 0; i < S; i++) z.push_back(i < len ? canon_u64(1) : Fr::zero());
    z.push_back(canon_u64(value));
    Fr h = assign_mimc(z, value, 110);
    if (commitment) memcpy(z[1].l, commitment, 32);
    else z[1] = h.to_canonical();
    for (uint32_t i = 0; i < S; i++) z.push_back(i == pos ? canon_u64(1) : Fr::zero());
    for (uint32_t i = 0; i < 2 * S; i++) z.push_back(Fr::zero());      // sel*(1-is_real), sel*(value-set): 0 when honest
    return z;
}

}  // namespace host
}  // namespace lzkp
