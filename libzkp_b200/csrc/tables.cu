// tables.cu — fixed-base window tables of the proving key, built once at lzkp_pk_load and kept
// resident in HBM (the device form of ProvingKey<Bn254>'s query vectors).
#include "tables.h"
#include "dev_util.cuh"

namespace lzkp {

// ---------------------------------------------------------------- fixed-base window tables
// table[(row * W + w) * N + (k-1)] = k * 2^(c*w) * base[row], affine, k = 1..N, N = 2^(c-1).
// Built as k = hi * 2^LB + lo from two small affine tables per (row, w) so that every entry
// is one affine addition with a per-thread batched inversion.

// wb[row * W + w] = 2^(c*w) * base[row]
template <class F>
__global__ void k_tb_window_bases(const Affine<F> *bases, uint32_t nb, uint32_t c, uint32_t W, Affine<F> *wb) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    Affine<F> p = ld_vec(bases + b);
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    st_vec(wb + (size_t)b * W, p);
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t i = 0; i < c; i++) acc.dbl_cold();
        st_vec(wb + (size_t)b * W + w, acc.to_affine());
    }
}
// small[(u * SM) + j]: j < NLO-1 -> (j+1) * B_u ; j >= NLO-1 -> (j - (NLO-1) + 1) * NLO * B_u   (NHI entries)
template <class F>
__global__ void k_tb_small(const Affine<F> *wb, uint32_t n_units, uint32_t lo_bits, uint32_t n_hi, Affine<F> *small) {
    const uint32_t n_lo = 1u << lo_bits, SM = n_lo - 1 + n_hi;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_units * SM) return;
    uint32_t u = (uint32_t)(t / SM), j = (uint32_t)(t % SM);
    uint32_t k = j < n_lo - 1 ? j + 1 : (j - (n_lo - 1) + 1) << lo_bits;
    XYZZ<F> base = XYZZ<F>::from_affine(ld_vec(wb + u)), acc = XYZZ<F>::inf();
    for (int i = 31 - __clz(k); i >= 0; i--) {
        acc.dbl_cold();
        if ((k >> i) & 1u) acc.add_cold(base);
    }
    st_vec(small + t, acc.to_affine());
}
// One thread fills G consecutive `lo` values of one (unit, hi).
template <class F, int G>
__global__ void __launch_bounds__(128) k_tb_fill(const Affine<F> *small, uint32_t n_units, uint32_t lo_bits,
                                                 uint32_t n_hi, uint32_t N, Affine<F> *table, size_t unit0) {
    const uint32_t n_lo = 1u << lo_bits, SM = n_lo - 1 + n_hi;
    const uint32_t groups_per_hi = n_lo / G, hi_count = (N >> lo_bits) + 1;   // hi = 0..N/n_lo
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_units * hi_count * groups_per_hi) return;
    uint32_t g = (uint32_t)(t % groups_per_hi);
    uint32_t hi = (uint32_t)((t / groups_per_hi) % hi_count);
    uint32_t u = (uint32_t)(t / ((size_t)groups_per_hi * hi_count));
    const Affine<F> *sm = small + (size_t)u * SM;
    Affine<F> *out = table + (unit0 + u) * (size_t)N;
    Affine<F> H = Affine<F>::inf();
    if (hi) H = ld_vec(sm + (n_lo - 1) + (hi - 1));
    // batched inversion of (x_L - x_H) over the group
    F pre[G], acc = F::one();
#pragma unroll
    for (int e = 0; e < G; e++) {
        uint32_t lo = g * G + e;
        pre[e] = acc;
        if (hi && lo) {
            F d = ld_vec(&sm[lo - 1].x) - H.x;
            if (!d.is_zero()) acc = acc * d;
        }
    }
    F inv = acc.inverse();
#pragma unroll
    for (int e = G - 1; e >= 0; e--) {
        uint32_t lo = g * G + e;
        uint32_t k = (hi << lo_bits) + lo;
        Affine<F> R = Affine<F>::inf();
        if (!hi) {
            if (lo) R = ld_vec(sm + lo - 1);
        } else if (!lo) {
            R = H;
        } else {
            Affine<F> L = ld_vec(sm + lo - 1);
            F d = L.x - H.x;
            if (d.is_zero()) {                       // never for prime-order bases; kept for totality
                XYZZ<F> s = XYZZ<F>::from_affine(H);
                s.madd_cold(L);
                R = s.to_affine();
            } else {
                F di = inv * pre[e];
                inv = inv * d;
                F lam = (L.y - H.y) * di;
                R.x = lam.sqr() - H.x - L.x;
                R.y = lam * (H.x - R.x) - H.y;
            }
        }
        if (k >= 1 && k <= N) st_vec(out + (k - 1), R);
    }
}


namespace eng {

template <class F, int G>
static int build_table_t(const Affine<F> *d_bases, uint32_t rows, int c, uint32_t W, uint32_t N, Affine<F> *d_table,
                         cudaStream_t st) {
    const uint32_t lo_bits = std::min<uint32_t>(8, (uint32_t)c - 1), n_lo = 1u << lo_bits, n_hi = N >> lo_bits;
    const uint32_t SM = n_lo - 1 + n_hi, ROWS_PER_PASS = 32;
    DBuf wb, small;
    TRY(wb.alloc((size_t)rows * W * sizeof(Affine<F>)));
    TRY(small.alloc((size_t)ROWS_PER_PASS * W * SM * sizeof(Affine<F>)));
    LAUNCH((k_tb_window_bases<F>), (rows + 63) / 64, 64, 0, st, d_bases, rows, (uint32_t)c, W, wb.as<Affine<F>>());
    for (uint32_t r0 = 0; r0 < rows; r0 += ROWS_PER_PASS) {
        uint32_t nr = std::min(ROWS_PER_PASS, rows - r0), units = nr * W;
        size_t t_small = (size_t)units * SM;
        LAUNCH((k_tb_small<F>), (unsigned)((t_small + 127) / 128), 128, 0, st, wb.as<Affine<F>>() + (size_t)r0 * W,
               units, lo_bits, n_hi, small.as<Affine<F>>());
        size_t t_fill = (size_t)units * (n_hi + 1) * (n_lo / G);
        LAUNCH((k_tb_fill<F, G>), (unsigned)((t_fill + 127) / 128), 128, 0, st, small.as<Affine<F>>(), units, lo_bits,
               n_hi, N, d_table, (size_t)r0 * W);
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return LZKP_OK;
}

int build_table_g1(const void *d_bases, uint32_t rows, int c, uint32_t W, uint32_t N, void *d_table, cudaStream_t st) {
    return build_table_t<Fq, 16>((const G1Affine *)d_bases, rows, c, W, N, (G1Affine *)d_table, st);
}
int build_table_g2(const void *d_bases, uint32_t rows, int c, uint32_t W, uint32_t N, void *d_table, cudaStream_t st) {
    return build_table_t<Fq2, 8>((const G2Affine *)d_bases, rows, c, W, N, (G2Affine *)d_table, st);
}

}  // namespace eng
}  // namespace lzkp
