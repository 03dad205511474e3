// Error plumbing, version and launch accounting shared by every entry point of libyvb200.so.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

namespace {
thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void yv_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool yv_pdl_enabled() {
    static const bool on = []() { const char* e = getenv("YVB200_PDL"); return e && e[0] == '1'; }();   // opt-in: measured no gain inside the captured step
    return on;
}
void yv_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" const char* yv_last_error(void) { return g_err; }
extern "C" int yv_version(void) { return 100; }
extern "C" uint64_t yv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
