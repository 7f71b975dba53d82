// Error state, version and launch accounting of the C ABI (include/gd_b200.h).
#include "common.cuh"
#include "../../include/gd_b200.h"
#include <stdarg.h>
#include <stdlib.h>

namespace gd {
namespace {
thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("GD_B200_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;  // measured slower: off unless asked for
  }
  return g_pdl != 0;
}
// the same attribute for the TINY kernels between two convs only (statistics finalize / affine table): their launch
// latency is a visible share of a small-batch step (GD_B200_PDL_SMALL=1; measured in profiles/pdl_small_r02.log)
static int g_pdl_small = -1;
bool pdl_small_enabled() {
  if (g_pdl_small < 0) {
    const char* e = getenv("GD_B200_PDL_SMALL");
    g_pdl_small = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl_small != 0 || pdl_enabled();
}
void conv_debug_set(int key, int value);
void attn_debug_set(int value);
}  // namespace gd

#ifdef GD_B200_DEVTOOLS
extern "C" void gd_debug_set(int key, int value) {
  if (key == 6) gd::g_pdl = value ? 1 : 0;
  else if (key == 5) gd::attn_debug_set(value);
  else gd::conv_debug_set(key, value);
}
#endif

extern "C" const char* gd_last_error(void) { return gd::g_err; }
extern "C" int gd_version(void) { return GD_B200_ABI_VERSION; }
extern "C" int64_t gd_launch_count(void) { return gd::g_launches; }
extern "C" void gd_launch_count_reset(void) { gd::g_launches = 0; }
