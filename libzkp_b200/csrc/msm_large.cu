#include "large.h"
namespace lzkp { namespace eng {
int large_msm_g1_host(const uint8_t *, const uint8_t *, size_t, uint8_t *) { return fail(LZKP_E_UNSUPPORTED, "lzkp_msm_g1 not built yet"); }
int large_msm_g2_host(const uint8_t *, const uint8_t *, size_t, uint8_t *) { return fail(LZKP_E_UNSUPPORTED, "lzkp_msm_g2 not built yet"); }
}}
