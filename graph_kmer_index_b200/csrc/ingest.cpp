// ingest.cpp -- CPU 2-bit packing of read rows (see ingest.h).  The wide path handles 64 bases per step with AVX-512BW
// (one byte shuffle maps letters to codes, two multiply-adds fold four 2-bit codes into a byte, vpmovdb gathers the 16 bytes); the scalar path is a 256-entry
// table.  Which one runs is decided once from cpuid; both produce identical words.
#include "ingest.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <thread>

#if defined(__x86_64__)
#include <immintrin.h>
#define GKI_X86 1
#endif

namespace gki {

namespace {

struct Lut {
    uint8_t v[256];   // bits 0-1 code, bit 2 valid
    Lut() {
        for (int i = 0; i < 256; i++) v[i] = 0;
        const char *letters = "acgt";
        for (int c = 0; c < 4; c++) {
            v[(uint8_t)letters[c]] = (uint8_t)(c | 4);
            v[(uint8_t)(letters[c] - 32)] = (uint8_t)(c | 4);
        }
    }
};
const Lut lut;

bool pack_row_scalar(const uint8_t *row, int32_t read_len, int words, uint64_t *out) {
    unsigned ok = 4;
    for (int w = 0; w < words; w++) {
        const int first = w * 32, nb = read_len - first < 32 ? read_len - first : 32;
        uint64_t cw = 0;
        for (int i = 0; i < nb; i++) {
            const unsigned e = lut.v[row[first + i]];
            ok &= e;
            cw |= (uint64_t)(e & 3u) << (2 * i);
        }
        out[w] = cw;
    }
    return ok != 0;
}

#if GKI_X86
__attribute__((target("avx512f,avx512bw"))) inline bool pack_row_avx512(const uint8_t *row, int32_t read_len, int words, uint64_t *out) {
    // The low nibble tells the four letters apart in either case (A 1, C 3, G 7, T 4): one byte shuffle gives the code, a second one
    // the lower-case letter that nibble stands for -- a byte is valid iff it, lower-cased, is that letter (bytes >= 128 shuffle to 0).
    const __m512i lower = _mm512_set1_epi8(0x20);
    const __m512i code_of = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i letter_of = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i w8 = _mm512_set1_epi16(0x0401);        // bytes (1, 4): code pairs -> 4 bits per 16-bit lane
    const __m512i w16 = _mm512_set1_epi32(0x00100001);   // words (1, 16): -> 8 bits (4 bases) per 32-bit lane
    __mmask64 bad = 0;
    for (int c = 0; c < read_len; c += 64) {
        const int n = read_len - c < 64 ? read_len - c : 64;
        const __mmask64 in = n == 64 ? ~0ull : ((1ull << n) - 1ull);
        const __m512i raw = n == 64 ? _mm512_loadu_si512(row + c) : _mm512_maskz_loadu_epi8(in, row + c);
        bad |= _mm512_cmpneq_epi8_mask(_mm512_or_si512(raw, lower), _mm512_shuffle_epi8(letter_of, raw)) & in;
        const __m512i code = _mm512_shuffle_epi8(code_of, raw);                      // bytes past the read are 0 and pack to zero
        const __m128i bits = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(code, w8), w16));   // 64 bases -> 128 bits
        const int w = c >> 5;
        if (w + 1 < words) _mm_storeu_si128((__m128i *)(out + w), bits);
        else _mm_storel_epi64((__m128i *)(out + w), bits);
    }
    return bad == 0;
}

bool have_avx512() {
    static const bool have = !getenv("GKI_PACK_SCALAR") && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    return have;
}
#else
bool have_avx512() { return false; }
#endif

}  // namespace

const char *pack_rows_isa() { return have_avx512() ? "avx512" : "scalar"; }

namespace {
// The row loop, once per instruction set: inside the AVX-512 instantiation the row packer inlines and its constants stay in
// registers across rows (a call per 150-byte row through the dispatch cost about a fifth of the time).
template <bool WIDE>
static inline __attribute__((always_inline)) int64_t pack_rows_loop(const uint8_t *reads, int64_t row_stride, const int64_t *row_offsets, int32_t read_len,
                                                                    int64_t r0, int64_t r1, uint64_t *packed, uint8_t *dirty_rows, int64_t *dirty_index,
                                                                    int64_t dirty_cap, int64_t *n_dirty, int64_t prefetch_rows) {
    const int words = (read_len + 31) / 32;
    int64_t clean = 0, dirty = 0;
    for (int64_t r = r0; r < r1; r++) {
        const uint8_t *row = reads + (row_offsets ? row_offsets[r] : r * row_stride);
        uint64_t *out = packed + clean * words;
#if GKI_X86
        if (prefetch_rows && r + prefetch_rows < r1) {   // hardware prefetchers do not run far enough ahead on short rows
            const uint8_t *ahead = reads + (row_offsets ? row_offsets[r + prefetch_rows] : (r + prefetch_rows) * row_stride);
            for (int32_t o = 0; o < read_len; o += 64) _mm_prefetch((const char *)(ahead + o), _MM_HINT_T0);
        }
        const bool ok = WIDE ? pack_row_avx512(row, read_len, words, out) : pack_row_scalar(row, read_len, words, out);
#else
        const bool ok = pack_row_scalar(row, read_len, words, out);
#endif
        if (ok) {
            clean++;
        } else {
            if (dirty_rows && dirty < dirty_cap) memcpy(dirty_rows + dirty * read_len, row, (size_t)read_len);
            if (dirty_index && dirty < dirty_cap) dirty_index[dirty] = r;
            dirty++;
        }
    }
    if (n_dirty) *n_dirty = dirty;
    return clean;
}
#if GKI_X86
__attribute__((target("avx512f,avx512bw"))) int64_t pack_rows_wide(const uint8_t *reads, int64_t row_stride, const int64_t *row_offsets, int32_t read_len,
                                                                   int64_t r0, int64_t r1, uint64_t *packed, uint8_t *dirty_rows, int64_t *dirty_index,
                                                                   int64_t dirty_cap, int64_t *n_dirty, int64_t prefetch_rows) {
    return pack_rows_loop<true>(reads, row_stride, row_offsets, read_len, r0, r1, packed, dirty_rows, dirty_index, dirty_cap, n_dirty, prefetch_rows);
}
#endif
}  // namespace

// plain read sweep of a byte range (the host-memory ceiling the packers run against): sum of its 64-bit words
#if GKI_X86
__attribute__((target("avx512f"))) static uint64_t sweep_words_avx512(const uint64_t *w, int64_t n) {
    __m512i a0 = _mm512_setzero_si512(), a1 = a0, a2 = a0, a3 = a0;
    int64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        _mm_prefetch((const char *)(w + i + 256), _MM_HINT_T0);
        _mm_prefetch((const char *)(w + i + 272), _MM_HINT_T0);
        a0 = _mm512_add_epi64(a0, _mm512_loadu_si512((const void *)(w + i)));
        a1 = _mm512_add_epi64(a1, _mm512_loadu_si512((const void *)(w + i + 8)));
        a2 = _mm512_add_epi64(a2, _mm512_loadu_si512((const void *)(w + i + 16)));
        a3 = _mm512_add_epi64(a3, _mm512_loadu_si512((const void *)(w + i + 24)));
    }
    uint64_t s = (uint64_t)_mm512_reduce_add_epi64(_mm512_add_epi64(_mm512_add_epi64(a0, a1), _mm512_add_epi64(a2, a3)));
    for (; i < n; i++) s += w[i];
    return s;
}
#endif
uint64_t sweep_words(const uint64_t *w, int64_t n) {
#if GKI_X86
    if (have_avx512()) return sweep_words_avx512(w, n);
#endif
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        a0 += w[i];
        a1 += w[i + 1];
        a2 += w[i + 2];
        a3 += w[i + 3];
    }
    for (; i < n; i++) a0 += w[i];
    return a0 + a1 + a2 + a3;
}

int64_t pack_rows(const uint8_t *reads, int64_t row_stride, const int64_t *row_offsets, int32_t read_len, int64_t r0, int64_t r1,
                  uint64_t *packed, uint8_t *dirty_rows, int64_t *dirty_index, int64_t dirty_cap, int64_t *n_dirty, int force_scalar) {
    static const int64_t prefetch_rows = 32;   // measured: 57 -> 78 GB/s with 14 threads
#if GKI_X86
    if (have_avx512() && !force_scalar)
        return pack_rows_wide(reads, row_stride, row_offsets, read_len, r0, r1, packed, dirty_rows, dirty_index, dirty_cap, n_dirty, prefetch_rows);
#endif
    return pack_rows_loop<false>(reads, row_stride, row_offsets, read_len, r0, r1, packed, dirty_rows, dirty_index, dirty_cap, n_dirty, prefetch_rows);
}

// ---- FASTA / FASTQ ----
static inline bool is_blank(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// One entry per newline of data[lo, hi), ascending: its position, and -- looked at while the bytes around it are in cache -- how
// many blanks precede it on its line, how many blanks open the next line, and whether the next line starts with '>'.
// With these the sequence lines can be cut out later without touching the file again.
constexpr int NL_POS_BITS = 48, NL_COUNT_MAX = 127;   // counts of 127 mean "127 or more": the emitter then looks at the file
constexpr int64_t NL_POS_MASK = (1ll << NL_POS_BITS) - 1;
static inline int64_t nl_pos(int64_t e) { return e & NL_POS_MASK; }
static inline int nl_tail(int64_t e) { return (int)((e >> NL_POS_BITS) & 127); }
static inline int nl_head(int64_t e) { return (int)((e >> (NL_POS_BITS + 7)) & 127); }
static inline bool nl_next_is_header(int64_t e) { return (e >> (NL_POS_BITS + 14)) & 1; }
static inline bool is_inline_blank(uint8_t c) { return c <= ' ' && c != '\n' && is_blank(c); }

static inline int64_t newline_entry(const uint8_t *data, size_t bytes, size_t p) {
    int64_t tail = 0, head = 0;
    if (p > 0 && data[p - 1] <= ' ')
        while (tail < NL_COUNT_MAX && (size_t)tail < p && is_inline_blank(data[p - 1 - (size_t)tail])) tail++;
    const size_t s = p + 1;
    if (s < bytes && data[s] <= ' ')
        while (head < NL_COUNT_MAX && s + (size_t)head < bytes && is_inline_blank(data[s + (size_t)head])) head++;
    const int64_t header = (s < bytes && data[s] == '>') ? 1 : 0;
    return (int64_t)p | (tail << NL_POS_BITS) | (head << (NL_POS_BITS + 7)) | (header << (NL_POS_BITS + 14));
}
static void find_newlines_scalar(const uint8_t *data, size_t bytes, size_t lo, size_t hi, std::vector<int64_t> &out) {
    const uint8_t *p = data + lo, *end = data + hi;
    while (p < end) {
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        out.push_back(newline_entry(data, bytes, (size_t)(q - data)));
        p = q + 1;
    }
}
#if GKI_X86
__attribute__((target("avx512f,avx512bw,bmi"))) static void find_newlines_avx512(const uint8_t *data, size_t bytes, size_t lo, size_t hi,
                                                                                std::vector<int64_t> &out) {
    const __m512i nl = _mm512_set1_epi8('\n');
    size_t i = lo;
    for (; i + 64 <= hi; i += 64) {
        uint64_t m = _mm512_cmpeq_epi8_mask(_mm512_loadu_si512(data + i), nl);
        while (m) {
            out.push_back(newline_entry(data, bytes, i + (size_t)__builtin_ctzll(m)));
            m &= m - 1;
        }
    }
    find_newlines_scalar(data, bytes, i, hi, out);
}
#endif
static void find_newlines(const uint8_t *data, size_t bytes, size_t lo, size_t hi, std::vector<int64_t> &out) {
#if GKI_X86
    if (have_avx512()) return find_newlines_avx512(data, bytes, lo, hi, out);
#endif
    find_newlines_scalar(data, bytes, lo, hi, out);
}

FastxFile *fastx_open(const char *path, int n_threads, std::string &err) {
    FastxFile *f = new FastxFile();
    f->fd = open(path, O_RDONLY);
    struct stat st;
    if (f->fd < 0 || fstat(f->fd, &st) != 0) {
        err = std::string("cannot open ") + path + ": " + strerror(errno);
        fastx_close(f);
        return nullptr;
    }
    f->bytes = (size_t)st.st_size;
    if (f->bytes == 0) return f;
    void *m = mmap(nullptr, f->bytes, PROT_READ, MAP_PRIVATE, f->fd, 0);
    if (m == MAP_FAILED) {
        err = std::string("cannot map ") + path + ": " + strerror(errno);
        fastx_close(f);
        return nullptr;
    }
    f->data = (const uint8_t *)m;
    madvise(m, f->bytes, MADV_SEQUENTIAL);
    f->format = f->data[0] == '@' ? 1 : 0;
    // One parallel pass over byte ranges records every newline (64 bytes per step with AVX-512BW) together with the blanks around
    // it; the prefix sums of the per-range counts then number the lines -- FASTQ needs that: the sequence is line 1 of every 4 --
    // and every thread cuts out the sequence lines that start right after one of ITS newlines (thread 0 also owns line 0) from
    // those entries alone.  The file is read once: a second look at it misses the cache for every line (measured: 3x the scan).
    int T = n_threads > 0 ? n_threads : default_pack_threads() + 1;
    if ((size_t)T > f->bytes / (1 << 20) + 1) T = (int)(f->bytes / (1 << 20) + 1);
    auto range = [&](int t) { return f->bytes * (size_t)t / (size_t)T; };
    auto run_parallel = [&](auto &&body) {
        std::vector<std::thread> threads;
        for (int t = 1; t < T; t++) threads.emplace_back(body, t);
        body(0);
        for (auto &th : threads) th.join();
    };
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now_ms();
    std::vector<std::vector<int64_t>> newline_at((size_t)T);
    run_parallel([&](int t) {
        std::vector<int64_t> v;   // local: the shared vector headers sit side by side, and every push_back writes one
        v.reserve((range(t + 1) - range(t)) / 64 + 16);
        find_newlines(f->data, f->bytes, range(t), range(t + 1), v);
        newline_at[(size_t)t] = std::move(v);
    });
    const double t_scanned = now_ms();
    std::vector<int64_t> newlines((size_t)T + 1, 0);   // newlines[t]: newlines before range t
    for (int t = 0; t < T; t++) newlines[(size_t)t + 1] = newlines[(size_t)t] + (int64_t)newline_at[(size_t)t].size();
    // entry of the first newline at or after range t; "the end of the file" when there is none (its tail count is found the slow way)
    const int64_t at_file_end = (int64_t)f->bytes | ((int64_t)NL_COUNT_MAX << NL_POS_BITS);
    std::vector<int64_t> next_newline((size_t)T + 1, at_file_end);
    for (int t = T - 1; t >= 0; t--) next_newline[(size_t)t] = newline_at[(size_t)t].empty() ? next_newline[(size_t)t + 1] : newline_at[(size_t)t][0];
    std::vector<std::vector<int64_t>> offs((size_t)T);
    std::vector<std::vector<int32_t>> lens((size_t)T);
    run_parallel([&](int t) {
        std::vector<int64_t> vo;
        std::vector<int32_t> vl;
        const std::vector<int64_t> &nl = newline_at[(size_t)t];
        const size_t guess = (nl.size() + 1) / (f->format ? 4 : 2) + 16;
        vo.reserve(guess);
        vl.reserve(guess);
        // the line numbered `line` is [begin, stop) without its newline; head: blanks it opens with, tail: blanks it ends with
        auto emit = [&](int64_t line, int64_t begin, bool is_header, int head, int64_t closing) {
            if (begin >= (int64_t)f->bytes) return;                    // nothing follows the file's last newline
            if (f->format ? (line % 4 != 1) : is_header) return;
            const int64_t stop = nl_pos(closing);
            int64_t a, b;
            if (head < NL_COUNT_MAX && nl_tail(closing) < NL_COUNT_MAX) {
                b = stop - nl_tail(closing);
                a = begin + head < b ? begin + head : b;
                if (b < begin) a = b = begin;                          // cannot happen: the tail count stops at the line's start
            } else {                                                   // long runs of blanks, or a last line without newline
                const uint8_t *pa = f->data + begin, *pb = f->data + stop;
                while (pb > pa && is_blank(pb[-1])) pb--;
                while (pa < pb && is_blank(*pa)) pa++;
                a = pa - f->data;
                b = pb - f->data;
            }
            vo.push_back(a);
            vl.push_back((int32_t)(b - a));
        };
        if (t == 0) emit(0, 0, f->data[0] == '>', NL_COUNT_MAX, next_newline[0]);
        for (size_t i = 0; i < nl.size(); i++)
            emit(newlines[(size_t)t] + (int64_t)i + 1, nl_pos(nl[i]) + 1, nl_next_is_header(nl[i]), nl_head(nl[i]),
                 i + 1 < nl.size() ? nl[i + 1] : next_newline[(size_t)t + 1]);
        offs[(size_t)t] = std::move(vo);
        lens[(size_t)t] = std::move(vl);
    });
    const double t_emitted = now_ms();
    std::vector<size_t> first((size_t)T + 1, 0);
    for (int t = 0; t < T; t++) first[(size_t)t + 1] = first[(size_t)t] + offs[(size_t)t].size();
    f->offsets.resize(first[(size_t)T]);
    f->lengths.resize(first[(size_t)T]);
    std::vector<int32_t> maxima((size_t)T, 0), minima((size_t)T, INT_MAX);
    run_parallel([&](int t) {
        const size_t n = offs[(size_t)t].size();
        if (n) {
            memcpy(f->offsets.data() + first[(size_t)t], offs[(size_t)t].data(), n * 8);
            memcpy(f->lengths.data() + first[(size_t)t], lens[(size_t)t].data(), n * 4);
        }
        for (int32_t v : lens[(size_t)t]) {
            if (v > maxima[(size_t)t]) maxima[(size_t)t] = v;
            if (v < minima[(size_t)t]) minima[(size_t)t] = v;
        }
    });
    f->min_len = INT_MAX;
    for (int t = 0; t < T; t++) {
        if (maxima[(size_t)t] > f->max_len) f->max_len = maxima[(size_t)t];
        if (minima[(size_t)t] < f->min_len) f->min_len = minima[(size_t)t];
    }
    if (f->offsets.empty()) f->min_len = 0;
    if (getenv("GKI_FASTX_DEBUG"))
        fprintf(stderr, "[gki fastx] %d threads: scan %.1f ms, lines %.1f ms, gather %.1f ms\n", T, t_scanned - t_start, t_emitted - t_scanned, now_ms() - t_emitted);
    return f;
}

void fastx_close(FastxFile *f) {
    if (!f) return;
    if (f->data) munmap((void *)f->data, f->bytes);
    if (f->fd >= 0) close(f->fd);
    delete f;
}

int default_pack_threads() {
    if (const char *e = getenv("GKI_PACK_THREADS")) {
        int v = atoi(e);
        return v < 0 ? 0 : (v > 64 ? 64 : v);
    }
    int hw = (int)std::thread::hardware_concurrency();
    int ranks = 1;   // one process per GPU under torchrun: the ranks of a node share its cores
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    int v = hw / ranks - (ranks == 1 ? 2 : 1);
    return v < 0 ? 0 : (v > 30 ? 30 : v);
}

}  // namespace gki
