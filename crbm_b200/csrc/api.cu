// Error plumbing, device queries and the trivial entry points of the C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"

namespace crbm {

static thread_local char g_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("CRBM_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;   // B200
    }
    return sms;
}

}  // namespace crbm

extern "C" {

int crbm_abi_version(void) { return 1; }
const char *crbm_last_error(void) { return crbm::g_error; }
int crbm_packed_row_words(int L) { return L > 0 ? crbm::packed_row_words(L) : CRBM_ERR_ARG; }
int crbm_hidden_words(int K) { return K > 0 ? crbm::ceil_div(K, 32) : CRBM_ERR_ARG; }
int crbm_fantasy_words(void) { return crbm::FANTASY_WORDS; }
unsigned long long crbm_launch_count(void) { return crbm::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
