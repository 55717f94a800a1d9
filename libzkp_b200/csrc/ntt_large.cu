#include "large.h"
namespace lzkp { namespace eng {
int large_ntt_host(uint8_t *, uint32_t, int, int) { return fail(LZKP_E_UNSUPPORTED, "lzkp_ntt: log_n > 12 not built yet"); }
}}
