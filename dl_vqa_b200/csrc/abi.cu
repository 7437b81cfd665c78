// ABI bookkeeping: version and per-thread error string.
#include <stdlib.h>
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void vqa_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
int vqa_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("VQA_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on;
}
void vqa_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" uint64_t vqa_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" const char* vqa_last_error_string(void) { return g_err; }
extern "C" int vqa_abi_version(void) { return VQA_ABI_VERSION; }

// host-side diagnostic: how the vector dropout scheme quantises p (see dropout_bf16_threshold in common.cuh)
extern "C" int vqa_dropout_threshold_pattern(float p, uint32_t* pattern, uint32_t* count16) {
    if (!pattern || !count16 || !(p >= 0.f) || !(p < 1.f)) { vqa_set_error("dropout_threshold_pattern: bad arguments"); return -1; }
    const Dropout d = make_dropout(0, p);
    *pattern = d.thr2 & 0xFFFFu;
    *count16 = d.threshold;
    return 0;
}
