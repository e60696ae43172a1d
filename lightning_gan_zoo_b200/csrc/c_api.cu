// C-ABI plumbing shared by every entry point of include/hologan_b200.h: version + thread-local errors.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

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

struct OptionDef { const char *name; int dflt; };
static const OptionDef kOptionDefs[kOptCount] = {
    {"ADAIN_CL_NO_CLUSTER", 0}, {"ADAIN_CL_CLUSTER_BWD", 0}, {"TAPGEMM_DUAL", -1},
    {"FINAL_CONV_MMA", 7},      {"ROTATE_SLAB32", 1},        {"ROTATE_GATHER_BWD", 1},
    {"ADAIN_GEMM_STATS", 0},    {"TAPGEMM_PERSISTENT", -1},  {"TAPGEMM_MSUB", 0},
    {"TAPGEMM_SHARE_A", 2},     {"ADAIN_CL_RING", 1},
    {"ADAIN_CL_SMALL_REGS", 1},
};
static std::atomic<int> g_options[kOptCount];
static std::once_flag g_options_once;

static void options_init()
{
    for (int i = 0; i < kOptCount; ++i) {
        char env[64];
        snprintf(env, sizeof env, "HG_%s", kOptionDefs[i].name);
        const char *e = getenv(env);
        g_options[i].store((e && e[0]) ? atoi(e) : kOptionDefs[i].dflt, std::memory_order_relaxed);
    }
}

int option(Option o)
{
    std::call_once(g_options_once, options_init);
    return g_options[o].load(std::memory_order_relaxed);
}

static int option_index(const char *name)
{
    if (!name) return -1;
    if (strncmp(name, "HG_", 3) == 0) name += 3;
    for (int i = 0; i < kOptCount; ++i)
        if (strcmp(name, kOptionDefs[i].name) == 0) return i;
    return -1;
}

}  // namespace hg

extern "C" int hg_set_option(const char *name, int value)
{
    const int i = hg::option_index(name);
    HG_REQUIRE(i >= 0, HG_ERR_INVALID_ARG, "hg_set_option: unknown option '%s'", name ? name : "(null)");
    hg::option(static_cast<hg::Option>(i));             // make sure the environment was read first
    hg::g_options[i].store(value, std::memory_order_relaxed);
    return HG_OK;
}

extern "C" int hg_get_option(const char *name, int *value)
{
    const int i = hg::option_index(name);
    HG_REQUIRE(i >= 0 && value, HG_ERR_INVALID_ARG, "hg_get_option: unknown option '%s' or null result pointer", name ? name : "(null)");
    *value = hg::option(static_cast<hg::Option>(i));
    return HG_OK;
}

extern "C" int hg_abi_version(void) { return HG_ABI_VERSION; }

extern "C" const char *hg_last_error(void) { return hg::error_buffer(); }
