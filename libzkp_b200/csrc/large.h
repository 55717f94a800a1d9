// large.h — stand-alone large-size transforms behind lzkp_ntt / lzkp_msm_g1 / lzkp_msm_g2.
#pragma once
#include "common.h"

namespace lzkp {
namespace eng {

// In-place radix-2 (i)NTT over Fr on 2^log_n canonical elements in HOST memory (ntt_large.cu).
int large_ntt_host(uint8_t *data, uint32_t log_n, int inverse, int coset);
// Pippenger MSM over host buffers in ark-serialize layout (msm_large.cu).
int large_msm_g1_host(const uint8_t *bases_affine, const uint8_t *scalars, size_t n, uint8_t *out_affine);
int large_msm_g2_host(const uint8_t *bases_affine, const uint8_t *scalars, size_t n, uint8_t *out_affine);

}  // namespace eng
}  // namespace lzkp
