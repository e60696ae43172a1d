// C-ABI plumbing shared by every entry point of include/hologan_b200.h: version + thread-local errors.
#include <stdarg.h>

#include "hg_common.cuh"

namespace hg {

char *error_buffer()
{
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace hg

extern "C" int hg_abi_version(void) { return HG_ABI_VERSION; }

extern "C" const char *hg_last_error(void) { return hg::error_buffer(); }
