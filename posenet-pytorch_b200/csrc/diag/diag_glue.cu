// Glue for libposenet_b200_diag.so (include/posenet_b200_diag.h): the diagnostics library links the probe kernels with the
// tensor-map encoder only, so it carries its own copy of the error plumbing of api.cu.
#include <stdarg.h>
#include <stdlib.h>

#include "../common.cuh"

namespace pn {

static thread_local char g_diag_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_diag_err, sizeof(g_diag_err), fmt, ap);
    va_end(ap);
}

}  // namespace pn

extern "C" const char *pn_diag_last_error_string(void) { return pn::g_diag_err; }
