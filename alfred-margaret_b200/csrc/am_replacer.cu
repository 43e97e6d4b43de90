// am_replacer.cu -- Replacer.build / run / runWithLimit (placeholder until the device passes land).
#include "am_api_internal.h"

using namespace am;

struct am_replacer { int dummy; };

extern "C" {
int am_replacer_build(const am_u8slice*, const am_u8slice*, size_t, int, const am_lower_table*, const am_options*, am_replacer** out) {
  if (out) *out = nullptr;
  return fail(AM_E_UNSUPPORTED, "replacer not built yet");
}
void am_replacer_free(am_replacer* r) { delete r; }
int am_replacer_run(const am_replacer*, am_u8slice, uint64_t, uint8_t**, uint64_t*, int*) { return fail(AM_E_UNSUPPORTED, "replacer not built yet"); }
}
