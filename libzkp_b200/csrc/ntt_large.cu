// ntt_large.cu — radix-2 (i)NTT over BN254 Fr for sizes up to 2^28 (lzkp_ntt, lzkp_ntt_device).
//
// Replaces ark-poly's Radix2EvaluationDomain::{fft,ifft}_in_place and their coset variants
// (get_coset(Fr::GENERATOR)), which ark-groth16's witness map calls seven times per proof
// (reached from src/backend/snark.rs:364,442).  Results are the unique
// field elements, so they are bit-identical to arkworks' for the same domain constants.
//
// Decomposition (Cooley-Tukey, decimation in frequency over digit groups): n = n_0 * n_1 * ... with
// every n_q <= 2^11.  Pass q runs, for each tile, a complete n_q-point NTT in shared memory on
// elements that differ only in digit q (stride M_q = n / (n_0 ... n_q), `cols` adjacent columns per
// CTA so global accesses are 32*cols contiguous bytes), then multiplies by the cross twiddle
// w^(P_q * r * k_q).  Passes 0..p-2 work in place; the last pass writes the natural-order result to
// the output buffer (digit-reversed scatter, one 32-byte sector per element).
//
// Data stays in the caller's representation: only data x constant products occur, and a Montgomery
// product with a Montgomery-form constant leaves the scale of the data unchanged, so canonical in
// gives canonical out with no conversion pass.
//
// Roofline: (n/2) log2 n butterflies x 1 Montgomery product (272 IMAD) + ~2n twiddle products per
// pass boundary; 64 B/element/pass of HBM traffic.  IMAD-bound by ~8x.
#include <cstdlib>
#include <map>
#include <mutex>

#include "dev_util.cuh"
#include "large.h"

namespace lzkp {

namespace {

constexpr uint32_t kMaxTileBits = 11;   // 2^11 x 32 B = 64 KB of shared memory per column

struct SmemVec {                          // two 128-bit planes: conflict-free LDS.128 / STS.128
    uint4 *lo, *hi;
    __device__ __forceinline__ Fr get(uint32_t i) const {
        Fr r;
        uint4 a = lo[i], b = hi[i];
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void put(uint32_t i, const Fr &v) const {
        lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};

struct PowTable {                         // base^e = hi[e >> lb] * lo[e & mask]   (Montgomery)
    const Fr *lo, *hi;
    uint32_t lb;
    __device__ __forceinline__ Fr pow(uint32_t e) const {
        return ldg_vec(hi + (e >> lb)) * ldg_vec(lo + (e & ((1u << lb) - 1u)));
    }
};

struct PassArgs {
    uint32_t log_n;         // m
    uint32_t bits[3];       // digit widths b_0 .. b_{p-1}
    uint32_t n_pass, q;     // number of passes, this pass
    uint32_t log_stride;    // log2 M_q
    uint32_t log_cols;      // columns per tile (1 << log_cols <= M_q)
    uint32_t tw_shift;      // tile twiddle table is for 2^tile_max points: index shift for smaller tiles
    const Fr *tile_tw;      // rho^k, k < 2^(tile_max-1)
    size_t batch_stride;    // elements between consecutive polynomials of a batch (blockIdx.y)
    const Fr *cross;        // cross twiddle of this pass by output position (n entries), or null: two-level powers
    PowTable w;             // powers of the n-th root (forward or inverse)
    PowTable g;             // powers of the coset generator (forward) or its inverse times n^-1 (inverse)
    Fr ninv;                // n^-1 (Montgomery), plain inverse only
    int inverse, coset;
};

// out[k] = scale * base^(k * step), k < count
__global__ void k_ntt_pow_table(Fr *out, Fr base, Fr scale, uint32_t count, uint32_t step_log) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint64_t e = (uint64_t)k << step_log;
    Fr acc = Fr::one(), b = base;
    while (e) {
        if (e & 1u) acc = acc * b;
        b = b.sqr();
        e >>= 1;
    }
    st_vec(out + k, acc * scale);
}

// cross[pos] = w^(P_q * r * k) for the element that pass q stores at `pos` (digit q holds k, r = pos mod M_q)
__global__ void k_ntt_cross_table(Fr *cross, PowTable w, uint32_t log_n, uint32_t b, uint32_t log_stride) {
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >> log_n) return;
    const uint32_t r = (uint32_t)pos & ((1u << log_stride) - 1u), k = ((uint32_t)pos >> log_stride) & ((1u << b) - 1u);
    const uint32_t e = (r * k) << (log_n - b - log_stride);
    st_vec(cross + pos, e ? w.pow(e) : Fr::one());
}

__global__ void __launch_bounds__(1024) k_ntt_pass(const Fr *__restrict__ in, Fr *__restrict__ out, PassArgs A) {
    extern __shared__ uint4 smem[];
    in += (size_t)blockIdx.y * A.batch_stride;
    out += (size_t)blockIdx.y * A.batch_stride;
    const uint32_t b = A.bits[A.q], nq = 1u << b, cols = 1u << A.log_cols, E = nq << A.log_cols;
    SmemVec s{smem, smem + E};
    const uint32_t tiles_per_hi = 1u << (A.log_stride - A.log_cols);
    const uint32_t hi = blockIdx.x >> (A.log_stride - A.log_cols), lo0 = (blockIdx.x & (tiles_per_hi - 1)) << A.log_cols;
    const size_t base = ((size_t)hi << (b + A.log_stride)) + lo0;
    const bool first = A.q == 0, last = A.q + 1 == A.n_pass;

    // ---- load (natural digit order), column-major in shared memory
    for (uint32_t idx = threadIdx.x; idx < E; idx += blockDim.x) {
        uint32_t j = idx >> A.log_cols, c = idx & (cols - 1);
        size_t pos = base + ((size_t)j << A.log_stride) + c;
        Fr v = ld_vec(in + pos);
        if (first && A.coset && !A.inverse) v = v * A.g.pow((uint32_t)pos);
        s.put((c << b) + j, v);
    }
    __syncthreads();

    // ---- n_q-point decimation-in-frequency NTT per column: natural in, bit-reversed out.
    // Two stages per shared-memory round trip (radix-4 in registers): a thread owns the four elements
    // i, i+Q, i+2Q, i+3Q (Q = quarter span), runs the span-2Q butterflies then the span-Q ones.
    int lh = (int)b;
    while (lh >= 2) {
        const uint32_t H = 1u << (lh - 1), Q = H >> 1;                 // half-distances of the two stages
        for (uint32_t t = threadIdx.x; t < E / 4; t += blockDim.x) {
            const uint32_t c = t >> (b - 2), tt = t & ((nq >> 2) - 1);
            const uint32_t low = tt & (Q - 1), i = (c << b) + (((tt >> (lh - 2)) << lh) | low);
            Fr x0 = s.get(i), x1 = s.get(i + Q), x2 = s.get(i + H), x3 = s.get(i + H + Q);
            // stage with half-distance H: twiddle exponent (index mod H) << (b - lh)
            const uint32_t k0 = low, k1 = low + Q;
            Fr y0 = x0 + x2, y2 = x0 - x2, y1 = x1 + x3, y3 = x1 - x3;
            if (k0) y2 = y2 * ldg_vec(A.tile_tw + ((k0 << (b - lh)) << A.tw_shift));
            y3 = y3 * ldg_vec(A.tile_tw + ((k1 << (b - lh)) << A.tw_shift));
            // stage with half-distance Q: twiddle exponent (index mod Q) << (b - lh + 1), shared by both pairs
            Fr z0 = y0 + y1, z1 = y0 - y1, z2 = y2 + y3, z3 = y2 - y3;
            if (lh > 2 && low) {
                Fr w = ldg_vec(A.tile_tw + ((low << (b - lh + 1)) << A.tw_shift));
                z1 = z1 * w;
                z3 = z3 * w;
            }
            s.put(i, z0); s.put(i + Q, z1); s.put(i + H, z2); s.put(i + H + Q, z3);
        }
        __syncthreads();
        lh -= 2;
    }
    if (lh == 1) {                                                       // odd number of stages: last radix-2, no twiddle
        for (uint32_t bf = threadIdx.x; bf < E / 2; bf += blockDim.x) {
            Fr x = s.get(2 * bf), y = s.get(2 * bf + 1);
            s.put(2 * bf, x + y);
            s.put(2 * bf + 1, x - y);
        }
        __syncthreads();
    }

    // ---- store
    if (!last) {
        // cross twiddle w^(P_q * r * k): P_q = n_0 ... n_{q-1} = 2^(log_n - b - log_stride), r = lo0 + c
        const uint32_t p_log = A.log_n - b - A.log_stride;
        for (uint32_t idx = threadIdx.x; idx < E; idx += blockDim.x) {
            uint32_t k = idx >> A.log_cols, c = idx & (cols - 1);
            Fr v = s.get((c << b) + bitrev(k, b));
            const size_t pos = base + ((size_t)k << A.log_stride) + c;
            uint32_t e = ((lo0 + c) * k) << p_log;
            if (e) v = v * (A.cross ? ldg_vec(A.cross + pos) : A.w.pow(e));
            st_vec(out + pos, v);
        }
    } else {
        // final index: digits of `hi` (k_0 most significant) reversed, k_{p-1} on top
        uint32_t rev = 0, shift = 0, rest = hi, rem_bits = A.log_n - b;
        for (uint32_t d = 0; d + 1 < A.n_pass; d++) {
            rem_bits -= A.bits[d];
            uint32_t digit = rest >> rem_bits;
            rest &= (1u << rem_bits) - 1u;
            rev |= digit << shift;
            shift += A.bits[d];
        }
        for (uint32_t k = threadIdx.x; k < nq; k += blockDim.x) {
            Fr v = s.get(bitrev(k, b));
            uint32_t K = rev | (k << shift);
            if (A.inverse) v = v * (A.coset ? A.g.pow(K) : A.ninv);
            st_vec(out + K, v);
        }
    }
}

// ---------------------------------------------------------------- plans (device tables per size/direction)
struct NttPlan {
    uint32_t log_n = 0, n_pass = 0, bits[3] = {0, 0, 0}, tile_max = 0, lb = 0;
    int device = 0;                                        // the tables live on this device
    eng::DBuf tile_tw, w_lo, w_hi, g_lo, g_hi, cross[2];   // cross[q]: per-position twiddles after pass q (optional)
    Fr ninv;
};
std::mutex g_plan_mu;
std::map<uint32_t, NttPlan *> g_plans;     // key = device * 256 + log_n * 2 + inverse

Fr host_pow2k(Fr x, uint32_t k) {           // x^(2^k)
    for (uint32_t i = 0; i < k; i++) x = x.sqr();
    return x;
}

}  // namespace

namespace eng {

static int get_plan(uint32_t log_n, int inverse, cudaStream_t st, NttPlan **out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    const int dev = std::max(0, current_device());
    uint32_t key = (uint32_t)dev * 256u + log_n * 2 + (inverse ? 1 : 0);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return LZKP_OK; }
    NttPlan *P = new NttPlan();
    P->log_n = log_n;
    P->device = dev;
    // Measured on B200: tiles of <= 2^10 points run at the full 64 G products/s (several CTAs per SM hide the
    // barriers), 2^11-point tiles at ~78 % of it; an extra pass costs one product per element (cross twiddle table).
    static const uint32_t tile_bits = getenv("LZKP_NTT_TILE_BITS") ? (uint32_t)atoi(getenv("LZKP_NTT_TILE_BITS")) : 10u;
    const uint32_t tb = std::min(std::max(tile_bits, 4u), kMaxTileBits);
    P->n_pass = std::max(1u, (log_n + tb - 1) / tb);
    if (P->n_pass > 3) P->n_pass = 3;
    for (uint32_t q = 0; q < P->n_pass; q++) P->bits[q] = log_n / P->n_pass + (q < log_n % P->n_pass ? 1 : 0);
    P->tile_max = P->bits[0];
    P->lb = (log_n + 1) / 2;
    Fr w, g;
    for (int i = 0; i < 8; i++) {
        w.l[i] = inverse ? FrParams::ROOT28_INV(i) : FrParams::ROOT28(i);
        g.l[i] = inverse ? FrParams::GEN_INV(i) : FrParams::GEN(i);
    }
    w = host_pow2k(w, 28 - log_n);                       // primitive n-th root (or its inverse)
    Fr nn = Fr::zero();
    nn.l[log_n >> 5] = 1u << (log_n & 31);               // n = 2^log_n (canonical limbs)
    P->ninv = Fr::from_canonical(nn).inverse();
    const uint32_t n_lo = 1u << P->lb, n_hi = 1u << (log_n - P->lb), n_tw = std::max(1u, (1u << P->tile_max) >> 1);
    TRY(P->tile_tw.alloc(sizeof(Fr) * n_tw));
    TRY(P->w_lo.alloc(sizeof(Fr) * n_lo)); TRY(P->w_hi.alloc(sizeof(Fr) * n_hi));
    TRY(P->g_lo.alloc(sizeof(Fr) * n_lo)); TRY(P->g_hi.alloc(sizeof(Fr) * n_hi));
    Fr rho = host_pow2k(w, log_n - P->tile_max);         // root of the largest tile transform
    LAUNCH(k_ntt_pow_table, (n_tw + 127) / 128, 128, 0, st, P->tile_tw.as<Fr>(), rho, Fr::one(), n_tw, 0u);
    LAUNCH(k_ntt_pow_table, (n_lo + 127) / 128, 128, 0, st, P->w_lo.as<Fr>(), w, Fr::one(), n_lo, 0u);
    LAUNCH(k_ntt_pow_table, (n_hi + 127) / 128, 128, 0, st, P->w_hi.as<Fr>(), w, Fr::one(), n_hi, P->lb);
    LAUNCH(k_ntt_pow_table, (n_lo + 127) / 128, 128, 0, st, P->g_lo.as<Fr>(), g, Fr::one(), n_lo, 0u);
    LAUNCH(k_ntt_pow_table, (n_hi + 127) / 128, 128, 0, st, P->g_hi.as<Fr>(), g, inverse ? P->ninv : Fr::one(), n_hi, P->lb);
    // per-position cross twiddles (one product per element instead of two) while they stay below 1 GiB
    static const bool cross_off = getenv("LZKP_NTT_NO_CROSS_TABLE") != nullptr;
    if (!cross_off && P->n_pass > 1 && ((size_t)32 << log_n) * (P->n_pass - 1) <= ((size_t)1 << 30)) {
        uint32_t log_stride = log_n;
        for (uint32_t q = 0; q + 1 < P->n_pass; q++) {
            log_stride -= P->bits[q];
            TRY(P->cross[q].alloc((size_t)32 << log_n));
            LAUNCH(k_ntt_cross_table, (unsigned)((((size_t)1 << log_n) + 255) / 256), 256, 0, st, P->cross[q].as<Fr>(),
                   PowTable{P->w_lo.as<Fr>(), P->w_hi.as<Fr>(), P->lb}, log_n, P->bits[q], log_stride);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    static uint64_t attr_set = 0;              // function attributes are per device
    if (!(attr_set >> (dev & 63) & 1)) {
        CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set |= 1ull << (dev & 63);
    }
    g_plans[key] = P;
    *out = P;
    return LZKP_OK;
}

// Releases the cached per-size tables (lzkp_shutdown).
void ntt_plans_free() {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (auto &kv : g_plans) { DeviceScope ds(kv.second->device); delete kv.second; }
    g_plans.clear();
}

// d_in is clobbered when the transform needs more than one pass; the result is written to d_out
// (d_out may equal d_in only for single-pass sizes, log_n <= 11).
int large_ntt_device(void *d_in, void *d_out, uint32_t log_n, int inverse, int coset, cudaStream_t st, uint32_t batch,
                     size_t batch_stride) {
    if (log_n > 28) return fail(LZKP_E_INVALID, "NTT size above 2^28 (the two-adicity of Fr)");
    NttPlan *P;
    TRY(get_plan(log_n, inverse, st, &P));
    PassArgs A;
    A.log_n = log_n;
    A.n_pass = P->n_pass;
    for (int i = 0; i < 3; i++) A.bits[i] = P->bits[i];
    A.tile_tw = P->tile_tw.as<Fr>();
    A.w = PowTable{P->w_lo.as<Fr>(), P->w_hi.as<Fr>(), P->lb};
    A.g = PowTable{P->g_lo.as<Fr>(), P->g_hi.as<Fr>(), P->lb};
    A.ninv = P->ninv;
    A.inverse = inverse;
    A.coset = coset;
    A.batch_stride = batch_stride;
    if (batch == 0) return LZKP_OK;
    if (batch > 65535) {                     // the batch index is the grid's y dimension: run it in slices
        for (uint32_t b0 = 0; b0 < batch; b0 += 65535) {
            const uint32_t nb = std::min(65535u, batch - b0);
            TRY(large_ntt_device((Fr *)d_in + (size_t)b0 * batch_stride, (Fr *)d_out + (size_t)b0 * batch_stride, log_n, inverse,
                                 coset, st, nb, batch_stride));
        }
        return LZKP_OK;
    }
    uint32_t log_stride = log_n;
    for (uint32_t q = 0; q < P->n_pass; q++) {
        const uint32_t b = P->bits[q];
        log_stride -= b;
        A.q = q;
        A.log_stride = log_stride;
        // two columns per tile (64 contiguous bytes) while the tile still fits 128 KB of shared memory
        // two adjacent columns per CTA (64 contiguous bytes) for small tiles; one column for 2^10+ tiles so that
        // several CTAs share an SM and their load / compute / store phases overlap
        A.log_cols = (log_stride >= 1 && b <= 9) ? 1 : 0;
        A.cross = (q + 1 < P->n_pass && q < 2) ? P->cross[q].as<Fr>() : nullptr;
        A.tw_shift = P->tile_max - b;
        const uint32_t E = 1u << (b + A.log_cols);
        const uint32_t threads = std::max(32u, std::min(1024u, E / 4));
        const size_t tiles = ((size_t)1 << log_n) >> (b + A.log_cols);
        const bool last = q + 1 == P->n_pass;
        LAUNCH(k_ntt_pass, dim3((unsigned)tiles, batch), threads, (size_t)32 * E, st, (const Fr *)d_in,
               (Fr *)(last ? d_out : d_in), A);
    }
    CUDA_TRY(cudaGetLastError());
    return LZKP_OK;
}

int large_ntt_host(uint8_t *data, uint32_t log_n, int inverse, int coset) {
    const size_t bytes = (size_t)32 << log_n;
    DBuf a, b;
    TRY(a.alloc(bytes)); TRY(b.alloc(bytes));
    CUDA_TRY(cudaMemcpy(a.p, data, bytes, cudaMemcpyHostToDevice));
    TRY(large_ntt_device(a.p, b.p, log_n, inverse, coset, nullptr));
    CUDA_TRY(cudaMemcpy(data, b.p, bytes, cudaMemcpyDeviceToHost));
    return LZKP_OK;
}

}  // namespace eng
}  // namespace lzkp
