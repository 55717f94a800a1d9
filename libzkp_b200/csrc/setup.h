// setup.h — device-side Groth16 setup (setup.cu).
#pragma once
#include <vector>
#include "common.h"

namespace lzkp {
namespace eng {
int setup_run(uint32_t m, uint32_t n_inst, uint32_t n_wit, const uint32_t *const rowptr[3],
              const uint32_t *const col[3], const uint8_t *const val[3], const uint8_t *toxic,
              std::vector<uint8_t> &pk_out, std::vector<uint8_t> &vk_out);
int generator_mul(int group, const uint8_t *scalars, size_t n, uint8_t *out);
void generator_tables_free();
}
}  // namespace lzkp
