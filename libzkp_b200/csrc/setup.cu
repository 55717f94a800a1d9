// setup.cu — Groth16 circuit-specific setup on the device (lzkp_setup / lzkp_setup_builtin).
//
// Replaces Groth16::<Bn254>::circuit_specific_setup (ark-groth16 generate_parameters_with_qap,
// un-vendored) which the reference calls at src/backend/snark.rs:318 and :337 when no key files
// exist.  The toxic waste (alpha, beta, gamma, delta, tau) is an INPUT here (the reference draws
// it from OsRng, snark.rs:310,331); the generators are the standard ones (arkworks draws random
// ones — any generator pair gives a valid key).  Output: ark-serialize uncompressed
// ProvingKey<Bn254> / VerifyingKey<Bn254>, the exact layout snark.rs:97-112 persists.
//
// Host: the QAP polynomials evaluated at tau (Lagrange coefficients by one batched inversion, then
// a sparse accumulation over the R1CS rows).  Device: every key element is one fixed-base scalar
// multiplication of a generator, done as W table gathers + XYZZ mixed adds per scalar.
#include <algorithm>
#include <cstring>
#include <mutex>

#include "host_util.h"
#include "setup.h"
#include "tables.h"
#include "dev_util.cuh"

namespace lzkp {

// out[i] = scalars[i] * G as ark-serialize uncompressed bytes; table = window table of G (1 row).
template <class F, int BYTES>
__global__ void __launch_bounds__(128) k_fixed_mul(const Affine<F> *__restrict__ table, uint32_t N, uint32_t c,
                                                   uint32_t W, const Fr *__restrict__ scalars, size_t count,
                                                   uint8_t *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t v[10];
    recode_offset(ld_vec(scalars + i), c, W, v);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t w = 0; w < W; w++) {
        int d = recoded_digit(v, c, w);
        if (d) acc.madd(gather_point(table, N, w, d));
    }
    Affine<F> a = acc.to_affine();
    if constexpr (BYTES == 64) write_g1(out + i * 64, a);
    else write_g2(out + i * 128, a);
}

namespace eng {

namespace {
struct HostCsr { const uint32_t *rowptr, *col; const uint8_t *val; };

Fr fr_load_canonical(const uint8_t *b, bool *ok) {
    Fr c;
    memcpy(c.l, b, 32);
    if (!host::lt_modulus(c)) *ok = false;
    return Fr::from_canonical(c);
}
void put_u64(std::vector<uint8_t> &o, uint64_t v) {
    for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
}
}  // namespace

namespace {
struct GenTables {
    DBuf t1, t2;
    static constexpr int c = 16;
    static constexpr uint32_t W = 16, N = 32768;
};
// window tables of the standard generators (built on first use, kept for the process lifetime)
GenTables *g_gen = nullptr;
bool g_gen_have2 = false;
std::mutex g_gen_mu;
int gen_tables(GenTables **out, int need_g2) {
    GenTables *&T = g_gen;
    bool &have2 = g_gen_have2;
    std::lock_guard<std::mutex> lk(g_gen_mu);
    if (!T) {
        GenTables *t = new GenTables();
        host::G1Canon g1c;
        g1c.x = Fq::zero(); g1c.y = Fq::zero();
        g1c.x.l[0] = 1; g1c.y.l[0] = 2;                                 // G1 generator (1, 2)
        G1Affine g1m{Fq::from_canonical(g1c.x), Fq::from_canonical(g1c.y)};
        DBuf d_g1;
        TRY(d_g1.alloc(sizeof(g1m)));
        CUDA_TRY(cudaMemcpy(d_g1.p, &g1m, sizeof(g1m), cudaMemcpyHostToDevice));
        TRY(t->t1.alloc((size_t)GenTables::W * GenTables::N * sizeof(G1Affine)));
        TRY(build_table_g1(d_g1.p, 1, GenTables::c, GenTables::W, GenTables::N, t->t1.p, nullptr));
        T = t;
    }
    if (need_g2 && !have2) {
        host::G2Canon g2c;
        for (int i = 0; i < 8; i++) {
            g2c.x0.l[i] = FqParams::G2X0(i); g2c.x1.l[i] = FqParams::G2X1(i);
            g2c.y0.l[i] = FqParams::G2Y0(i); g2c.y1.l[i] = FqParams::G2Y1(i);
        }
        G2Affine g2m{Fq2{Fq::from_canonical(g2c.x0), Fq::from_canonical(g2c.x1)},
                     Fq2{Fq::from_canonical(g2c.y0), Fq::from_canonical(g2c.y1)}};
        DBuf d_g2;
        TRY(d_g2.alloc(sizeof(g2m)));
        CUDA_TRY(cudaMemcpy(d_g2.p, &g2m, sizeof(g2m), cudaMemcpyHostToDevice));
        TRY(T->t2.alloc((size_t)GenTables::W * GenTables::N * sizeof(G2Affine)));
        TRY(build_table_g2(d_g2.p, 1, GenTables::c, GenTables::W, GenTables::N, T->t2.p, nullptr));
        have2 = true;
    }
    *out = T;
    return LZKP_OK;
}
}  // namespace

void generator_tables_free() {
    std::lock_guard<std::mutex> lk(g_gen_mu);
    delete g_gen;
    g_gen = nullptr;
    g_gen_have2 = false;
}

// out[i] = scalars[i] * generator (ark-serialize affine bytes); scalars canonical, host buffers
int generator_mul(int group, const uint8_t *scalars, size_t n, uint8_t *out) {
    TRY(ensure_device());
    GenTables *T;
    TRY(gen_tables(&T, group == 2));
    if (n == 0) return LZKP_OK;
    const size_t pb = group == 1 ? 64 : 128;
    DBuf d_s, d_o;
    TRY(d_s.alloc(n * 32)); TRY(d_o.alloc(n * pb));
    CUDA_TRY(cudaMemcpy(d_s.p, scalars, n * 32, cudaMemcpyHostToDevice));
    if (group == 1)
        LAUNCH((k_fixed_mul<Fq, 64>), (unsigned)((n + 127) / 128), 128, 0, 0, T->t1.as<G1Affine>(), GenTables::N,
               (uint32_t)GenTables::c, GenTables::W, d_s.as<Fr>(), n, d_o.as<uint8_t>());
    else
        LAUNCH((k_fixed_mul<Fq2, 128>), (unsigned)((n + 127) / 128), 128, 0, 0, T->t2.as<G2Affine>(), GenTables::N,
               (uint32_t)GenTables::c, GenTables::W, d_s.as<Fr>(), n, d_o.as<uint8_t>());
    CUDA_TRY(cudaMemcpy(out, d_o.p, n * pb, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaGetLastError());
    return LZKP_OK;
}

int setup_run(uint32_t m, uint32_t n_inst, uint32_t n_wit, const uint32_t *const rowptr[3],
              const uint32_t *const col[3], const uint8_t *const val[3], const uint8_t *toxic,
              std::vector<uint8_t> &pk_out, std::vector<uint8_t> &vk_out) {
    const uint32_t nv = n_inst + n_wit;
    if (n_inst < 1 || nv < 2) return fail(LZKP_E_INVALID, "setup: empty circuit");
    uint64_t need = (uint64_t)m + n_inst;
    uint32_t log_n = 0;
    while ((1ull << log_n) < need) log_n++;
    if (log_n < 1) log_n = 1;
    if (log_n > 28) return fail(LZKP_E_INVALID, "setup: domain larger than 2^28");
    const size_t n = (size_t)1 << log_n;
    bool ok = true;
    Fr alpha = fr_load_canonical(toxic, &ok), beta = fr_load_canonical(toxic + 32, &ok),
       gamma = fr_load_canonical(toxic + 64, &ok), delta = fr_load_canonical(toxic + 96, &ok),
       tau = fr_load_canonical(toxic + 128, &ok);
    if (!ok || alpha.is_zero() || beta.is_zero() || gamma.is_zero() || delta.is_zero() || tau.is_zero())
        return fail(LZKP_E_INVALID, "setup: toxic-waste scalars must be canonical and non-zero");

    // ---- Lagrange coefficients u_i = L_i(tau) = Z(tau)/n * w^i / (tau - w^i)
    Fr w;
    for (int i = 0; i < 8; i++) w.l[i] = FrParams::ROOT28(i);
    for (uint32_t i = log_n; i < 28; i++) w = w.sqr();
    Fr tn = tau;
    for (uint32_t i = 0; i < log_n; i++) tn = tn.sqr();
    Fr zt = tn - Fr::one();
    if (zt.is_zero()) return fail(LZKP_E_INVALID, "setup: tau lies in the evaluation domain");
    Fr v = zt * host::fr_from_u64(n).inverse();
    std::vector<Fr> u(n), pre(n);
    {
        Fr wi = Fr::one(), acc = Fr::one();
        for (size_t i = 0; i < n; i++) {
            u[i] = tau - wi;               // non-zero: tau is outside the domain
            pre[i] = acc;
            acc = acc * u[i];
            wi = wi * w;
        }
        Fr inv = acc.inverse();
        // walk back: 1/d_i = inv * pre[i]; need w^i again -> recompute backwards with w^-1
        Fr winv = w.inverse(), wcur = wi * winv;    // w^(n-1)
        for (size_t i = n; i-- > 0;) {
            Fr di = inv * pre[i];
            inv = inv * u[i];
            u[i] = v * wcur * di;
            wcur = wcur * winv;
        }
    }
    // ---- QAP at tau
    std::vector<Fr> qa(nv, Fr::zero()), qb(nv, Fr::zero()), qc(nv, Fr::zero());
    for (uint32_t j = 0; j < n_inst; j++) qa[j] = u[m + j];
    std::vector<Fr> *q[3] = {&qa, &qb, &qc};
    for (int k = 0; k < 3; k++) {
        for (uint32_t i = 0; i < m; i++)
            for (uint32_t t = rowptr[k][i]; t < rowptr[k][i + 1]; t++) {
                uint32_t cj = col[k][t];
                if (cj >= nv) return fail(LZKP_E_INVALID, "setup: matrix column out of range");
                bool okv = true;
                Fr cv = fr_load_canonical(val[k] + 32 * (size_t)t, &okv);
                if (!okv) return fail(LZKP_E_INVALID, "setup: non-canonical matrix coefficient");
                (*q[k])[cj] = (*q[k])[cj] + u[i] * cv;
            }
    }
    // ---- scalars of every key element (canonical)
    const Fr gi = gamma.inverse(), di = delta.inverse();
    std::vector<Fr> s1, s2;                     // G1 / G2 scalar lists
    s1.reserve(3 + n_inst + 2 * (size_t)nv + (n - 1) + n_wit);
    auto push1 = [&](const Fr &x) { s1.push_back(x.to_canonical()); };
    push1(alpha); push1(beta); push1(delta);
    for (uint32_t j = 0; j < nv; j++) {
        Fr t = beta * qa[j] + alpha * qb[j] + qc[j];
        qc[j] = t;                              // reuse: combined numerator
    }
    for (uint32_t j = 0; j < n_inst; j++) push1(qc[j] * gi);       // gamma_abc_g1
    for (uint32_t j = 0; j < nv; j++) push1(qa[j]);                 // a_query
    for (uint32_t j = 0; j < nv; j++) push1(qb[j]);                 // b_g1_query
    {
        Fr p = zt * di;                                             // h_query: tau^i * Z(tau) / delta
        for (size_t i = 0; i + 1 < n; i++) { push1(p); p = p * tau; }
    }
    for (uint32_t j = n_inst; j < nv; j++) push1(qc[j] * di);       // l_query
    s2.push_back(beta.to_canonical()); s2.push_back(gamma.to_canonical()); s2.push_back(delta.to_canonical());
    for (uint32_t j = 0; j < nv; j++) s2.push_back(qb[j].to_canonical());   // b_g2_query

    // ---- device: generator tables + fixed-base multiplications
    TRY(ensure_device());
    GenTables *GT;
    TRY(gen_tables(&GT, 1));
    const int c = GenTables::c;
    const uint32_t W = GenTables::W, N = GenTables::N;
    DBuf &t1 = GT->t1, &t2 = GT->t2;
    DBuf d_s1, d_s2, d_o1, d_o2;
    cudaStream_t st = nullptr;
    TRY(upload(d_s1, s1)); TRY(upload(d_s2, s2));
    TRY(d_o1.alloc(s1.size() * 64)); TRY(d_o2.alloc(s2.size() * 128));
    LAUNCH((k_fixed_mul<Fq, 64>), (unsigned)((s1.size() + 127) / 128), 128, 0, st, t1.as<G1Affine>(), N, (uint32_t)c, W,
           d_s1.as<Fr>(), s1.size(), d_o1.as<uint8_t>());
    LAUNCH((k_fixed_mul<Fq2, 128>), (unsigned)((s2.size() + 127) / 128), 128, 0, st, t2.as<G2Affine>(), N, (uint32_t)c, W,
           d_s2.as<Fr>(), s2.size(), d_o2.as<uint8_t>());
    std::vector<uint8_t> o1(s1.size() * 64), o2(s2.size() * 128);
    CUDA_TRY(cudaMemcpy(o1.data(), d_o1.p, o1.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(o2.data(), d_o2.p, o2.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaGetLastError());

    // ---- ark-serialize layout
    const uint8_t *p1 = o1.data(), *p2 = o2.data();
    const uint8_t *alpha_g1 = p1, *beta_g1 = p1 + 64, *delta_g1 = p1 + 128, *gabc = p1 + 192;
    const uint8_t *aq = gabc + 64 * (size_t)n_inst, *b1q = aq + 64 * (size_t)nv, *hq = b1q + 64 * (size_t)nv,
                  *lq = hq + 64 * (n - 1);
    const uint8_t *beta_g2 = p2, *gamma_g2 = p2 + 128, *delta_g2 = p2 + 256, *b2q = p2 + 384;
    auto app = [](std::vector<uint8_t> &o, const uint8_t *b, size_t len) { o.insert(o.end(), b, b + len); };
    vk_out.clear();
    app(vk_out, alpha_g1, 64); app(vk_out, beta_g2, 128); app(vk_out, gamma_g2, 128); app(vk_out, delta_g2, 128);
    put_u64(vk_out, n_inst); app(vk_out, gabc, 64 * (size_t)n_inst);
    pk_out.clear();
    pk_out.reserve(vk_out.size() + 128 + 40 + 64 * (2 * (size_t)nv + n + n_wit) + 128 * (size_t)nv);
    app(pk_out, vk_out.data(), vk_out.size());
    app(pk_out, beta_g1, 64); app(pk_out, delta_g1, 64);
    put_u64(pk_out, nv); app(pk_out, aq, 64 * (size_t)nv);
    put_u64(pk_out, nv); app(pk_out, b1q, 64 * (size_t)nv);
    put_u64(pk_out, nv); app(pk_out, b2q, 128 * (size_t)nv);
    put_u64(pk_out, n - 1); app(pk_out, hq, 64 * (n - 1));
    put_u64(pk_out, n_wit); app(pk_out, lq, 64 * (size_t)n_wit);
    return LZKP_OK;
}

}  // namespace eng
}  // namespace lzkp
