// large.h — stand-alone large-size transforms behind lzkp_ntt / lzkp_msm_g1 / lzkp_msm_g2.
#pragma once
#include "common.h"

namespace lzkp {
namespace eng {

// In-place radix-2 (i)NTT over Fr on 2^log_n canonical elements in HOST memory (ntt_large.cu).
int large_ntt_host(uint8_t *data, uint32_t log_n, int inverse, int coset);
// Device buffers: d_in is clobbered for multi-pass sizes (log_n > 11), the result lands in d_out.
// `batch` polynomials, `batch_stride` elements apart, are transformed by the same launches.
int large_ntt_device(void *d_in, void *d_out, uint32_t log_n, int inverse, int coset, cudaStream_t st, uint32_t batch = 1,
                     size_t batch_stride = 0);
void ntt_plans_free();
// Pippenger MSM over host buffers in ark-serialize layout (msm_large.cu).
int large_msm_g1_host(const uint8_t *bases_affine, const uint8_t *scalars, size_t n, uint8_t *out_affine);
int large_msm_g2_host(const uint8_t *bases_affine, const uint8_t *scalars, size_t n, uint8_t *out_affine);

// Resident MSM bases (msm_large.cu): group 1 = G1, 2 = G2; `resident` keeps 2^(c*w) * P for every window.
struct MsmBases;
// canon_input != 0: `bases` holds parsed points (canonical limbs, x||y, (0,0) = infinity) instead of ark bytes
int msm_bases_load(int group, const uint8_t *bases, size_t n, int window_bits, int resident, int validate, MsmBases **out,
                   int canon_input = 0);
void msm_bases_free(MsmBases *b);
uint32_t msm_bases_size(const MsmBases *b);
int msm_bases_group(const MsmBases *b);
int msm_host(MsmBases *b, const uint8_t *scalars, size_t n_used, uint8_t *out);
int msm_device(MsmBases *b, const void *d_scalars, size_t n_used, void *d_out, cudaStream_t st);
int msm_device_raw(MsmBases *b, const void *d_scalars, size_t n_used, void *d_out_xyzz, cudaStream_t st);
int msm_take_bad(MsmBases *b, cudaStream_t st, int *bad);

}  // namespace eng
}  // namespace lzkp
