// verify.h — batched device-side Groth16 verification (verify.cu).
#pragma once
#include <mutex>
#include "common.h"

namespace lzkp {
namespace eng {
struct VerifyingKeyDev;
int vk_load(const uint8_t *bytes, size_t len, VerifyingKeyDev **out);
void vk_free(VerifyingKeyDev *v);
uint32_t vk_num_inputs(const VerifyingKeyDev *v);
int verify_batch(VerifyingKeyDev *v, size_t n, const uint8_t *proofs, const uint8_t *inputs, size_t n_pub, uint8_t *ok_out);
}  // namespace eng
}  // namespace lzkp
