// ingest.h -- host side of read ingestion: 2-bit packing of ASCII read rows on the CPU.
// Plain C++ (compiled by the host compiler, no CUDA), shared by count.cu's host-input pipeline and gki_pack_reads.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace gki {

// Pack rows [r0, r1) of an ASCII read matrix.  A clean row (every byte one of ACGTacgt) becomes `words` = ceil(read_len/32)
// 64-bit words -- base i at bits 2*(i%32) of word i/32, a0 c1 g2 t3 (flat_kmers.py:134-145), unused high bits zero --
// appended to `packed`; a row with any other byte is appended to `dirty_rows` (read_len ASCII bytes, dense) when that is
// non-NULL and its index to `dirty_index` when that is non-NULL -- the first dirty_cap of them; the rest are only counted.
// Returns the number of clean rows; *n_dirty the others.  force_scalar != 0 selects the table-driven path (tests).
// row_offsets (may be NULL): row r starts at reads + row_offsets[r] instead of reads + r * row_stride (sequence lines of a file).
uint64_t sweep_words(const uint64_t *words, int64_t n);   // sum of n 64-bit words (a plain read sweep, AVX-512 where available)
int64_t pack_rows(const uint8_t *reads, int64_t row_stride, const int64_t *row_offsets, int32_t read_len, int64_t r0, int64_t r1,
                  uint64_t *packed, uint8_t *dirty_rows, int64_t *dirty_index, int64_t dirty_cap, int64_t *n_dirty, int force_scalar);

// "avx512" or "scalar": which implementation pack_rows dispatches to on this CPU
const char *pack_rows_isa();

// std::vector whose resize() leaves new elements uninitialised: the line arrays of a large file are filled by all threads at once,
// and a zero-fill by the resizing thread would touch (and page-fault) every page first
template <typename T> struct DefaultInitAllocator : std::allocator<T> {
    template <typename U> struct rebind { using other = DefaultInitAllocator<U>; };
    template <typename U, typename... Args> void construct(U *p, Args &&...args) {
        if constexpr (sizeof...(Args) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<Args>(args)...);
    }
};
using OffsetVector = std::vector<int64_t, DefaultInitAllocator<int64_t>>;
using LengthVector = std::vector<int32_t, DefaultInitAllocator<int32_t>>;

// A FASTA / FASTQ file mapped into memory with the positions of its sequence lines: FASTA -- every line that does not start
// with '>' (what read_kmers.py:16-18 treats as a read), FASTQ (first byte '@') -- the second line of every four.  Lines are
// stripped of surrounding blanks like the reference's line.strip().
struct FastxFile {
    int fd = -1;
    const uint8_t *data = nullptr;
    size_t bytes = 0;
    int format = 0;                 // 0 FASTA, 1 FASTQ
    OffsetVector offsets;           // start of every sequence line
    LengthVector lengths;           // its length after stripping
    int32_t max_len = 0, min_len = 0;
};
// NULL on failure (message in err)
FastxFile *fastx_open(const char *path, int n_threads, std::string &err);
void fastx_close(FastxFile *f);

// number of packing threads the host-input pipeline uses: GKI_PACK_THREADS, else hardware threads - 2 (at most 30);
// under torchrun (LOCAL_WORLD_SIZE ranks on the node) each rank takes its share of the cores minus one
int default_pack_threads();

}  // namespace gki
