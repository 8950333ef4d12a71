// Host side of the streaming ingest (secomo/sequences.py:33-52 readSeqsFromFasta, :20-31 the base table, :101-117
// seqToOneHot): a block of complete multi-FASTA records -> rows of 2-bit bases in the device layout (common.cuh), written where the
// caller says — a page-locked staging buffer, so that one cudaMemcpyAsync of 0.25 byte per base is all that crosses PCIe and no packing
// kernel runs. No device code in this file.
//
// The block is cut into segments at record starts, one host thread per segment:
//   pass 1  every thread counts the headers of its segment (memchr): an upper bound on the rows it will write, hence its first row;
//   pass 2  every thread parses its records: a sequence line is found with memchr, translated 16 bytes at a time
//           (code = ((c >> 1) ^ (c >> 2)) & 3 maps A/a C/c G/g T/t to 0 1 2 3; (c | 0x20) is compared with the four letters
//           for validity, both loops auto-vectorised), a line holding anything else goes through the byte-wise table of
//           crbm_fasta_parse_block_host (N: record skipped, blank / CR: dropped, other letters: error), and a complete record of L
//           codes is packed 8 bases per 64-bit word operation into its row;
//   pass 3  rows of segments behind a segment that dropped records (N, or a header without sequence) are moved down.
// The first error in file order is the one reported, as the one-pass parser does.
#include <stdlib.h>
#include <string.h>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <vector>
#include "common.cuh"

namespace {

struct Lut {
    unsigned char v[256];
    Lut() {
        for (int i = 0; i < 256; ++i) v[i] = 255;
        v[(int)'A'] = v[(int)'a'] = 0; v[(int)'C'] = v[(int)'c'] = 1; v[(int)'G'] = v[(int)'g'] = 2; v[(int)'T'] = v[(int)'t'] = 3;
        v[(int)'N'] = v[(int)'n'] = 4;
        v[10] = v[13] = v[32] = v[9] = 254;                   // line ends and blanks inside a record are dropped
    }
};
const Lut g_lut;

inline bool header_at(const unsigned char *buf, size_t i) { return buf[i] == '>' && (i == 0 || buf[i - 1] == 10); }

// first header at or after `from` (n if none)
size_t next_header(const unsigned char *buf, size_t n, size_t from) {
    size_t i = from;
    while (i < n) {
        const void *p = memchr(buf + i, '>', n - i);
        if (!p) return n;
        i = (const unsigned char *)p - buf;
        if (header_at(buf, i)) return i;
        ++i;
    }
    return n;
}

long long count_headers(const unsigned char *buf, size_t begin, size_t end) {
    long long c = 0;
    for (size_t i = next_header(buf, end, begin); i < end; i = next_header(buf, end, i + 1)) ++c;
    return c;
}

// 8 base codes (one per byte, little endian) -> 16 bits
inline uint32_t pack8(uint64_t x) {
    x = (x | (x >> 6)) & 0x000F000F000F000FULL;
    x = (x | (x >> 12)) & 0x000000FF000000FFULL;
    x = (x | (x >> 24)) & 0xFFFFULL;
    return (uint32_t)x;
}

// nw * 16 base codes (one per byte, zero beyond the record) -> one row in the device layout, pad words zeroed
inline void pack_row(const unsigned char *codes, int nw, uint32_t *row, int row_words) {
    for (int w = 0; w < crbm::FRONT_PAD_WORDS; ++w) row[w] = 0;
    for (int w = 0; w < nw; ++w) {
        uint64_t a, b;
        memcpy(&a, codes + 16 * w, 8);
        memcpy(&b, codes + 16 * w + 8, 8);
        row[crbm::FRONT_PAD_WORDS + w] = pack8(a) | (pack8(b) << 16);
    }
    for (int w = crbm::FRONT_PAD_WORDS + nw; w < row_words; ++w) row[w] = 0;
}

// fn(0) here, fn(1..T-1) on threads of their own; a thread that cannot be started (process limit) has its share run here instead
template <typename F>
void run_segments(int T, F fn) {
    std::vector<std::thread> th;
    int started = 1;
    try {
        for (int t = 1; t < T; ++t) { th.emplace_back(fn, t); ++started; }
    } catch (...) {
    }
    fn(0);
    for (int t = started; t < T; ++t) fn(t);
    for (auto &x : th) x.join();
}

struct alignas(64) Segment {
    size_t begin = 0, end = 0;          // [begin, end) of the block; begin is at a header
    long long first_row = 0;            // row index of this segment's first record (from the header counts)
    long long headers = 0, n_out = 0, n_skip = 0;
    int rc = CRBM_OK, bad = 0;
};

// translate a sequence line; returns false if it holds anything but ACGT / acgt
inline bool translate_line(const unsigned char *src, size_t len, unsigned char *dst) {
#if defined(__SSE2__)
    if (len >= 16) {
        const __m128i x20 = _mm_set1_epi8(0x20), three = _mm_set1_epi8(3);
        const __m128i la = _mm_set1_epi8('a'), lc = _mm_set1_epi8('c'), lg = _mm_set1_epi8('g'), lt = _mm_set1_epi8('t');
        __m128i ok = _mm_set1_epi8((char)0xff);
        auto step = [&](size_t j) {
            const __m128i c = _mm_loadu_si128((const __m128i *)(src + j));
            const __m128i u = _mm_or_si128(c, x20);
            ok = _mm_and_si128(ok, _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(u, la), _mm_cmpeq_epi8(u, lc)),
                                                _mm_or_si128(_mm_cmpeq_epi8(u, lg), _mm_cmpeq_epi8(u, lt))));
            _mm_storeu_si128((__m128i *)(dst + j), _mm_and_si128(_mm_xor_si128(_mm_srli_epi16(c, 1), _mm_srli_epi16(c, 2)), three));
        };
        size_t j = 0;
        for (; j + 16 <= len; j += 16) step(j);
        if (j < len) step(len - 16);                              // the last 16 bytes again (same values)
        return _mm_movemask_epi8(ok) == 0xffff;
    }
#endif
    unsigned char bad = 0;
    for (size_t j = 0; j < len; ++j) {
        const unsigned char c = src[j], u = c | 0x20;
        bad |= (unsigned char)!((u == 'a') | (u == 'c') | (u == 'g') | (u == 't'));
        dst[j] = ((c >> 1) ^ (c >> 2)) & 3;
    }
    return !bad;
}

// One segment of complete records. A record's length is the number of its non-blank characters (the reference: len(record.seq));
// records without any are dropped; a record containing N / n is skipped whatever else it holds (the reference drops it before it
// looks at anything else, sequences.py:47-50); then another letter outside ACGT is error -12 (KeyError of sequences.py:30) and a
// length other than L error -11 (the one-hot arrays cannot be concatenated). Output: packed rows (`rows`) or L base codes (`codes`).
void parse_segment(const unsigned char *buf, Segment &sg, long long L, uint32_t *rows, int row_words, unsigned char *codes,
                   size_t capacity_rows) {
    const int nw = (int)((L + 15) / 16);
    std::vector<unsigned char> tmp((size_t)nw * 16 + 16, 0);      // codes of the record being read; the tail beyond L stays zero
    size_t i = sg.begin;
    const size_t end = sg.end;
    while (i < end) {
        const void *p = memchr(buf + i, 10, end - i);              // header line
        i = p ? (size_t)((const unsigned char *)p - buf) + 1 : end;
        long long len = 0;
        bool has_n = false;
        int badc = 0;
        while (i < end && buf[i] != '>') {                         // sequence lines (i is at a line start)
            p = memchr(buf + i, 10, end - i);
            const size_t e = p ? (size_t)((const unsigned char *)p - buf) : end;
            size_t line = e - i;
            if (line && buf[e - 1] == 13) --line;                  // CRLF file: the CR is dropped, as the table does
            if (len + (long long)line <= L && translate_line(buf + i, line, tmp.data() + len)) {
                len += (long long)line;
            } else {
                for (size_t j = i; j < e; ++j) {
                    const unsigned char v = g_lut.v[buf[j]];
                    if (v == 254) continue;
                    if (v <= 3) { if (len < L) tmp[len] = v; }
                    else if (v == 4) has_n = true;
                    else if (!badc) badc = buf[j];
                    ++len;
                }
            }
            i = p ? e + 1 : end;
        }
        if (len == 0) continue;                                    // a header without sequence: dropped
        if (has_n) { ++sg.n_skip; continue; }
        if (badc) { sg.bad = badc; sg.rc = -12; return; }
        if (len != L) { sg.rc = -11; return; }
        const size_t r = (size_t)(sg.first_row + sg.n_out);
        if (r >= capacity_rows) { sg.rc = -13; return; }
        if (codes) memcpy(codes + r * (size_t)L, tmp.data(), (size_t)L);
        else pack_row(tmp.data(), nw, rows + r * (size_t)row_words, row_words);
        ++sg.n_out;
    }
}

// ---- one-hot (N,1,4,L) float32 in host memory -> packed rows (A0: the layout secomo/sequences.py:20-31 builds, read back) -------------
// Same rule as pack_onehot_kernel (ops.cu): a position is valid iff exactly one of its four values == 1.0f and the other three
// == 0.0f; the base is the plane holding the one. Returns the number of invalid positions of sequences [n0, n1).
long long pack_onehot_range(const float *oh, long long n0, long long n1, int L, long long seq_stride, long long plane_stride,
                            uint32_t *rows, int row_words) {
    const int nw = (L + 15) / 16;
    std::vector<unsigned char> tmp((size_t)nw * 16 + 16, 0);
    long long bad = 0;
    for (long long n = n0; n < n1; ++n) {
        const float *pa = oh + n * seq_stride, *pc = pa + plane_stride, *pg = pc + plane_stride, *pt = pg + plane_stride;
        int i = 0;
#if defined(__SSE2__)
        const __m128 one = _mm_set1_ps(1.f), zero = _mm_setzero_ps();
        const __m128i i1 = _mm_set1_epi32(1), i2 = _mm_set1_epi32(2), m1 = _mm_set1_epi32(-1), m3 = _mm_set1_epi32(-3);
        __m128i badv = _mm_setzero_si128();
        for (; i + 4 <= L; i += 4) {
            const __m128 a = _mm_loadu_ps(pa + i), c = _mm_loadu_ps(pc + i), g = _mm_loadu_ps(pg + i), t = _mm_loadu_ps(pt + i);
            const __m128i ea = _mm_castps_si128(_mm_cmpeq_ps(a, one)), ec = _mm_castps_si128(_mm_cmpeq_ps(c, one));
            const __m128i eg = _mm_castps_si128(_mm_cmpeq_ps(g, one)), et = _mm_castps_si128(_mm_cmpeq_ps(t, one));
            const __m128i za = _mm_castps_si128(_mm_cmpeq_ps(a, zero)), zc = _mm_castps_si128(_mm_cmpeq_ps(c, zero));
            const __m128i zg = _mm_castps_si128(_mm_cmpeq_ps(g, zero)), zt = _mm_castps_si128(_mm_cmpeq_ps(t, zero));
            const __m128i ones = _mm_add_epi32(_mm_add_epi32(ea, ec), _mm_add_epi32(eg, et));           // -(number of ones)
            const __m128i zeros = _mm_add_epi32(_mm_add_epi32(za, zc), _mm_add_epi32(zg, zt));          // -(number of zeros)
            const __m128i ok = _mm_and_si128(_mm_cmpeq_epi32(ones, m1), _mm_cmpeq_epi32(zeros, m3));
            badv = _mm_add_epi32(badv, _mm_add_epi32(ok, i1));                                          // + (ok ? 0 : 1)
            const __m128i code = _mm_or_si128(_mm_and_si128(_mm_or_si128(ec, et), i1), _mm_and_si128(_mm_or_si128(eg, et), i2));
            const __m128i b16 = _mm_packs_epi32(code, code);
            const int four = _mm_cvtsi128_si32(_mm_packus_epi16(b16, b16));
            memcpy(tmp.data() + i, &four, 4);
        }
        {
            int lanes[4];
            _mm_storeu_si128((__m128i *)lanes, badv);
            bad += (long long)lanes[0] + lanes[1] + lanes[2] + lanes[3];
        }
#endif
        for (; i < L; ++i) {
            const float a = pa[i], c = pc[i], g = pg[i], t = pt[i];
            const int ones = (a == 1.f) + (c == 1.f) + (g == 1.f) + (t == 1.f);
            const int zeros = (a == 0.f) + (c == 0.f) + (g == 0.f) + (t == 0.f);
            bad += ones != 1 || zeros != 3;
            tmp[i] = (unsigned char)(((c == 1.f) | (t == 1.f)) | (((g == 1.f) | (t == 1.f)) << 1));
        }
        pack_row(tmp.data(), nw, rows + (size_t)n * row_words, row_words);
    }
    return bad;
}

}  // namespace

extern "C" int crbm_pack_onehot_host(const float *onehot, long long N, int L, long long seq_stride, long long plane_stride,
                                     uint32_t *rows, int row_words, long long *bad_count, int n_threads) {
    if (!onehot || !rows || !bad_count || N <= 0 || L <= 0) { crbm::set_error("crbm_pack_onehot_host: bad arguments"); return CRBM_ERR_ARG; }
    if (row_words < crbm::packed_row_words(L)) { crbm::set_error("crbm_pack_onehot_host: row_words %d < %d", row_words, crbm::packed_row_words(L)); return CRBM_ERR_ARG; }
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const long long min_positions = 1 << 16;                       // a thread is not worth starting for less
    if ((long long)T > N * L / min_positions + 1) T = (int)(N * L / min_positions + 1);
    if ((long long)T > N) T = (int)N;
    std::vector<long long> bad((size_t)T * 8, 0);                  // (one cache line per thread)
    auto work = [&](int t) {
        bad[(size_t)t * 8] = pack_onehot_range(onehot, N * t / T, N * (t + 1) / T, L, seq_stride, plane_stride, rows, row_words);
    };
    run_segments(T, work);
    long long total = 0;
    for (int t = 0; t < T; ++t) total += bad[(size_t)t * 8];
    *bad_count = total;
    return CRBM_OK;
}

static int fasta_block(const unsigned char *buf, size_t n, uint32_t *rows, int row_words, unsigned char *codes, size_t capacity_rows, int *L_io,
                       long long *n_records, long long *n_skipped, int *bad_char, int n_threads) {
    *n_records = 0; *n_skipped = 0;
    long long L = *L_io;
    const size_t first = next_header(buf, n, 0);                   // anything before the first header is not a record
    if (L <= 0) {                          // record length = that of the first record that has bases and no N (N records are skipped)
        long long skipped = 0;
        for (size_t i = first; i < n && L <= 0;) {
            const void *p = memchr(buf + i, 10, n - i);
            i = p ? (size_t)((const unsigned char *)p - buf) + 1 : n;
            const size_t e = next_header(buf, n, i);
            long long len = 0;
            bool has_n = false;
            for (size_t j = i; j < e; ++j) { const unsigned char v = g_lut.v[buf[j]]; len += v != 254; has_n |= v == 4; }
            if (len > 0 && !has_n) L = len;
            else if (len > 0) ++skipped;
            i = e;
        }
        if (L <= 0) { *n_skipped = skipped; return CRBM_OK; }      // nothing but N records / empty records in this block
        if (L > 0x7fffffff) { crbm::set_error("crbm_fasta_pack_block_host: record longer than 2^31 bases"); return CRBM_ERR_ARG; }
        *L_io = (int)L;
    }
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    size_t min_segment = 1 << 20;                                  // a thread is not worth starting for less
    if (const char *e = getenv("CRBM_INGEST_MIN_SEGMENT")) {       // (tests: many segments in a small file)
        const long long v = atoll(e);
        if (v > 0) min_segment = (size_t)v;
    }
    if ((size_t)T > (n - first) / min_segment + 1) T = (int)((n - first) / min_segment + 1);

    std::vector<Segment> seg((size_t)T);
    size_t at = first;
    for (int t = 0; t < T; ++t) {
        seg[t].begin = at;
        at = t + 1 == T ? n : next_header(buf, n, first + (n - first) / T * (size_t)(t + 1));
        if (at < seg[t].begin) at = seg[t].begin;
        seg[t].end = at;
    }
    auto run = [&](auto fn) { run_segments(T, fn); };
    // (one thread: no first rows to find; the headers are counted only if the parser runs out of rows)
    const bool have_out = rows || codes;
    const bool counted = T > 1 || !have_out || capacity_rows == 0;
    if (counted) run([&](int t) { seg[t].headers = count_headers(buf, seg[t].begin, seg[t].end); });
    long long need = 0;
    for (int t = 0; t < T; ++t) { seg[t].first_row = need; need += seg[t].headers; }
    if (counted && ((size_t)need > capacity_rows || !have_out)) {      // the caller sizes its buffer from this and calls again
        *n_records = need;
        crbm::set_error("crbm_fasta_pack_block_host: %lld rows needed, %zu given", need, capacity_rows);
        return -13;
    }
    if (rows && row_words < crbm::packed_row_words((int)L)) { crbm::set_error("crbm_fasta_pack_block_host: row_words %d < %d", row_words, crbm::packed_row_words((int)L)); return CRBM_ERR_ARG; }
    run([&](int t) { parse_segment(buf, seg[t], L, rows, row_words, codes, capacity_rows); });
    long long out = 0, skipped = 0;
    for (int t = 0; t < T; ++t) {
        if (seg[t].rc == -13) {                                    // (single thread) out of rows: now count
            *n_records = count_headers(buf, first, n);
            crbm::set_error("crbm_fasta_pack_block_host: %lld rows needed, %zu given", *n_records, capacity_rows);
            return -13;
        }
        if (seg[t].rc != CRBM_OK) {                                // first error in file order
            if (seg[t].rc == -12) *bad_char = seg[t].bad;
            return seg[t].rc;
        }
        if (seg[t].first_row != out && seg[t].n_out) {
            if (codes) memmove(codes + (size_t)out * L, codes + (size_t)seg[t].first_row * L, (size_t)seg[t].n_out * L);
            else memmove(rows + (size_t)out * row_words, rows + (size_t)seg[t].first_row * row_words, (size_t)seg[t].n_out * row_words * sizeof(uint32_t));
        }
        out += seg[t].n_out;
        skipped += seg[t].n_skip;
    }
    *n_records = out; *n_skipped = skipped;
    return CRBM_OK;
}

extern "C" int crbm_fasta_pack_block_host(const unsigned char *buf, size_t n, uint32_t *rows, int row_words, size_t capacity_rows, int *L_io,
                                          long long *n_records, long long *n_skipped, int *bad_char, int n_threads) {
    if (!buf || !L_io || !n_records || !n_skipped || !bad_char) { crbm::set_error("crbm_fasta_pack_block_host: null argument"); return CRBM_ERR_ARG; }
    return fasta_block(buf, n, rows, row_words, nullptr, capacity_rows, L_io, n_records, n_skipped, bad_char, n_threads);
}

// the same pass with one byte per base as its output (iterFastaCodes)
extern "C" int crbm_fasta_parse_block_host(const unsigned char *buf, size_t n, unsigned char *codes, size_t capacity_records, int *L_io,
                                           long long *n_records, long long *n_skipped, int *bad_char) {
    if (!buf || !codes || !L_io || !n_records || !n_skipped || !bad_char) { crbm::set_error("crbm_fasta_parse_block_host: null argument"); return CRBM_ERR_ARG; }
    const int rc = fasta_block(buf, n, nullptr, 0, codes, capacity_records, L_io, n_records, n_skipped, bad_char, 1);
    if (rc == -13) { crbm::set_error("crbm_fasta_parse_block_host: more records than the output holds"); return CRBM_ERR_ARG; }
    return rc;
}
