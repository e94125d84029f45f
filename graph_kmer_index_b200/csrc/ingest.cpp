// ingest.cpp -- CPU 2-bit packing of read rows (see ingest.h).  The wide path handles 64 bases per step with AVX-512BW
// (two multiply-adds fold four 2-bit codes into a byte, vpmovdb gathers the 16 bytes); the scalar path is a 256-entry
// table.  Which one runs is decided once from cpuid; both produce identical words.
#include "ingest.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
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
__attribute__((target("avx512f,avx512bw"))) bool pack_row_avx512(const uint8_t *row, int32_t read_len, int words, uint64_t *out) {
    const __m512i lower = _mm512_set1_epi8(0x20), three = _mm512_set1_epi8(3);
    const __m512i letters = _mm512_broadcast_i32x4(_mm_setr_epi8('a', 'c', 'g', 't', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i w8 = _mm512_set1_epi16(0x0401);        // bytes (1, 4): code pairs -> 4 bits per 16-bit lane
    const __m512i w16 = _mm512_set1_epi32(0x00100001);   // words (1, 16): -> 8 bits (4 bases) per 32-bit lane
    __mmask64 bad = 0;
    for (int c = 0; c < read_len; c += 64) {
        const int n = read_len - c < 64 ? read_len - c : 64;
        const __mmask64 in = n == 64 ? ~0ull : ((1ull << n) - 1ull);
        const __m512i lc = _mm512_or_si512(n == 64 ? _mm512_loadu_si512(row + c) : _mm512_maskz_loadu_epi8(in, row + c), lower);
        // (bit1 ^ bit2, bit2 ^ bit3) of the lower-case letter is a0 c1 g2 t3; the 16-bit shifts only leak into bits that are masked off
        const __m512i code = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(lc, 1), _mm512_srli_epi16(lc, 2)), three);
        bad |= _mm512_cmpneq_epi8_mask(lc, _mm512_shuffle_epi8(letters, code)) & in;
        const __m512i clean = _mm512_maskz_mov_epi8(in, code);                       // bytes past the read pack to zero
        const __m128i bits = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(clean, w8), w16));   // 64 bases -> 128 bits
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

int64_t pack_rows(const uint8_t *reads, int64_t row_stride, const int64_t *row_offsets, int32_t read_len, int64_t r0, int64_t r1,
                  uint64_t *packed, uint8_t *dirty_rows, int64_t *dirty_index, int64_t dirty_cap, int64_t *n_dirty, int force_scalar) {
    const int words = (read_len + 31) / 32;
    const bool wide = have_avx512() && !force_scalar;
    static const int64_t prefetch_rows = getenv("GKI_PACK_PREFETCH") ? atoll(getenv("GKI_PACK_PREFETCH")) : 32;   // measured: 57 -> 78 GB/s with 14 threads
    int64_t clean = 0, dirty = 0;
    for (int64_t r = r0; r < r1; r++) {
        const uint8_t *row = reads + (row_offsets ? row_offsets[r] : r * row_stride);
        uint64_t *out = packed + clean * words;
#if GKI_X86
        if (prefetch_rows && r + prefetch_rows < r1) {   // hardware prefetchers do not run far enough ahead on short rows
            const uint8_t *ahead = reads + (row_offsets ? row_offsets[r + prefetch_rows] : (r + prefetch_rows) * row_stride);
            for (int32_t o = 0; o < read_len; o += 64) _mm_prefetch((const char *)(ahead + o), _MM_HINT_T0);
        }
#endif
        bool ok;
#if GKI_X86
        ok = wide ? pack_row_avx512(row, read_len, words, out) : pack_row_scalar(row, read_len, words, out);
#else
        ok = pack_row_scalar(row, read_len, words, out);
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

// ---- FASTA / FASTQ ----
static inline bool is_blank(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

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
    // Two parallel passes over byte ranges: count the newlines of every range (their prefix sums number the lines, which FASTQ
    // needs: the sequence is line 1 of every 4), then every thread emits the sequence lines that START in its range.
    int T = n_threads > 0 ? n_threads : default_pack_threads() + 1;
    if ((size_t)T > f->bytes / (1 << 20) + 1) T = (int)(f->bytes / (1 << 20) + 1);
    auto range = [&](int t) { return f->bytes * (size_t)t / (size_t)T; };
    auto run_parallel = [&](auto &&body) {
        std::vector<std::thread> threads;
        for (int t = 1; t < T; t++) threads.emplace_back(body, t);
        body(0);
        for (auto &th : threads) th.join();
    };
    std::vector<int64_t> newlines((size_t)T + 1, 0);
    run_parallel([&](int t) {
        int64_t n = 0;
        const uint8_t *p = f->data + range(t), *end = f->data + range(t + 1);
        while (p < end) {
            const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
            if (!q) break;
            n++;
            p = q + 1;
        }
        newlines[(size_t)t + 1] = n;
    });
    for (int t = 0; t < T; t++) newlines[(size_t)t + 1] += newlines[(size_t)t];   // lines that start before range t: newlines[t] (+1 for line 0)
    std::vector<std::vector<int64_t>> offs((size_t)T);
    std::vector<std::vector<int32_t>> lens((size_t)T);
    const uint8_t *file_end = f->data + f->bytes;
    run_parallel([&](int t) {
        std::vector<int64_t> &vo = offs[(size_t)t];
        std::vector<int32_t> &vl = lens[(size_t)t];
        const size_t guess = (range(t + 1) - range(t)) / (f->format ? 200 : 100) + 16;
        vo.reserve(guess);
        vl.reserve(guess);
        int64_t line = newlines[(size_t)t] + (t == 0 ? 0 : 1);   // index of the first line that starts in this range
        const uint8_t *p = f->data + range(t), *end = f->data + range(t + 1);
        if (t != 0) {   // the first line start in the range follows the first newline at or after range(t) - 1
            const uint8_t *from = f->data + range(t) - 1;
            const uint8_t *q = (const uint8_t *)memchr(from, '\n', (size_t)(file_end - from));
            p = q ? q + 1 : file_end;
            // newlines before range(t) number newlines[t]; if data[range(t) - 1] is itself a newline the line starts exactly at range(t)
            line = newlines[(size_t)t] + (from[0] == '\n' ? 0 : 1);
        }
        while (p < end && p < file_end) {   // p: start of a line inside the range
            const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(file_end - p));
            const uint8_t *stop = q ? q : file_end;   // [p, stop): the line without its newline
            if (f->format ? (line % 4 == 1) : (*p != '>')) {
                const uint8_t *a = p, *b = stop;
                while (b > a && is_blank(b[-1])) b--;
                while (a < b && is_blank(*a)) a++;
                vo.push_back((int64_t)(a - f->data));
                vl.push_back((int32_t)(b - a));
            }
            line++;
            p = stop + 1;
        }
    });
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
