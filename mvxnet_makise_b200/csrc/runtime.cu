// Error text, launch counter and version of the mvx_b200 C-ABI library.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace mvx {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mvx

extern "C" const char *mvx_last_error(void) { return mvx::g_err; }
extern "C" int mvx_version(void) { return 100; }
extern "C" int64_t mvx_launch_count(void) { return mvx::g_launches.load(std::memory_order_relaxed); }
