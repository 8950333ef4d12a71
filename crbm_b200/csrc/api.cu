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

// ---- host side of the streaming ingest: one pass over a block of complete FASTA records -> base codes --------------------------
int crbm_fasta_parse_block_host(const unsigned char *buf, size_t n, unsigned char *codes, size_t capacity_records, int *L_io,
                                long long *n_records, long long *n_skipped, int *bad_char) {
    if (!buf || !codes || !L_io || !n_records || !n_skipped || !bad_char) { crbm::set_error("crbm_fasta_parse_block_host: null argument"); return CRBM_ERR_ARG; }
    static unsigned char lut[256];
    static bool lut_ready = false;
    if (!lut_ready) {
        for (int i = 0; i < 256; ++i) lut[i] = 255;
        lut[(int)'A'] = lut[(int)'a'] = 0; lut[(int)'C'] = lut[(int)'c'] = 1; lut[(int)'G'] = lut[(int)'g'] = 2; lut[(int)'T'] = lut[(int)'t'] = 3;
        lut[(int)'N'] = lut[(int)'n'] = 4;
        lut[10] = lut[13] = lut[32] = lut[9] = 254;          // line ends and blanks inside a record are dropped
        lut_ready = true;
    }
    long long L = *L_io, out = 0, skipped = 0;
    size_t i = 0;
    while (i < n && !(buf[i] == '>' && (i == 0 || buf[i - 1] == 10))) ++i;      // anything before the first header is not a record
    while (i < n) {
        while (i < n && buf[i] != 10) ++i;                    // header line
        if (i < n) ++i;
        if ((size_t)out >= capacity_records) { crbm::set_error("crbm_fasta_parse_block_host: more records than the output holds"); return CRBM_ERR_ARG; }
        unsigned char *row = codes + (size_t)out * (size_t)(L > 0 ? L : 0);
        long long len = 0;
        bool has_n = false;
        // sequence lines until the next header at a line start; before L is known the first record is written from `codes` on
        while (i < n && !(buf[i] == '>' && buf[i - 1] == 10)) {
            const unsigned char v = lut[buf[i]];
            if (v <= 4) {
                if (L > 0 && len >= L) return -11;            // ragged: longer than the first record
                row[len++] = v;
                has_n |= v == 4;
            } else if (v == 255) { *bad_char = buf[i]; return -12; }
            ++i;
        }
        if (len == 0) continue;                               // a header without sequence: dropped
        if (L <= 0) L = len;
        else if (len != L) return -11;
        if (has_n) ++skipped; else ++out;
    }
    *L_io = (int)L; *n_records = out; *n_skipped = skipped;
    return CRBM_OK;
}

}  // extern "C"
