// index.cuh -- the device-resident index object shared by index.cu (reference layout + lookups) and
// count.cu (probe table + counting).
#pragma once
#include "common.cuh"

namespace gki {

// Reference layout on the device: cells[b] = {hashes_to_index[b], n_kmers[b]} (one 8-byte load per probe instead
// of two loads from two tables), kmers / nodes (+ optional payload columns) in bucket order.
struct IndexView {
    const uint2 *cells;
    const uint64_t *kmers;
    FastMod fm;
};

// Counting structure (count.cu): a bucketised open-addressing table over the DISTINCT index k-mers.
//   bucket = 64 bytes = key[4] (one 32-byte sector, read by a probe with ONE 256-bit load) | cnt[4][2] (the other sector);
//   key    = canonical form min(x, revcomp_k(x)) when the k-mer length k is known (k > 0), else x itself;
//   cnt[0] = number of counted queries q with canonical(q) == key and orientation o = (q != key) equal to 0;
//   cnt[1] = (the same for orientation 1) - cnt[0]; both modulo 2^32 (count.cu: add_both_orientations / read_orientation).
// Both strands of a read position share one canonical key, so a position costs ONE filter access and at most
// ONE table access, and a hit is ONE 32-bit RED on the line that was just fetched.
//   filter = register-blocked Bloom filter (32-bit words, filter_k bits per key) sized to stay resident in L2.
struct Bucket {
    unsigned long long key[4];
    uint32_t cnt[4][2];
};
constexpr unsigned long long SLOT_EMPTY = ~0ull;
constexpr int SLOTS_PER_BUCKET = 4;

struct TableView {
    Bucket *buckets;          // n_buckets, + 1 special bucket for the key that equals SLOT_EMPTY (raw mode only)
    const uint32_t *filter;   // filter_words u32 (NULL: no filter)
    uint32_t n_buckets;
    uint32_t filter_words;
    int32_t filter_k;         // bits set per key (1..3)
    int32_t filter_m;         // 0: the filter word is chosen by the key's hash; m > 0: by the k-mer's minimizer (canonical m-mer with
                              // the smallest hash), so consecutive windows of a read share 32-byte sectors (filters larger than L2)
    int32_t k;                // canonical k-mer length, 0 = raw keys
};

}  // namespace gki

struct gki_index {
    int device = 0;
    int64_t n = 0;
    uint64_t modulo = 0;
    gki::FastMod fm{};
    uint2 *cells = nullptr;
    uint64_t *kmers = nullptr;
    uint32_t *nodes = nullptr;
    uint64_t *ref_offsets = nullptr;
    uint16_t *freq = nullptr;
    float *af = nullptr;
    int64_t max_node = -1;
    int64_t nonempty = 0;
    uint64_t max_kmer = 0;
    size_t device_bytes = 0;
    // counting structure, built by the first counting call (count.cu)
    gki::TableView table{};
    // entries regrouped by count-table slot (count.cu): slot index and node | orientation << 31, so that get_node_counts
    // streams the table instead of looking every entry up
    uint32_t *cs_slot = nullptr, *cs_node = nullptr;
    // get_node_counts with more nodes than L2 keeps (count.cu): where every tile of entries writes its (node, weight) words, grouped
    // by node range (nb_seg[range * nb_tiles + tile], one more element at the end); built by the first such call
    uint32_t *nb_seg = nullptr;
    uint32_t nb_bounds[65] = {};       // first word of every range, and the total
    int32_t nb_ranges = 0, nb_shift = 0;
    int64_t nb_tiles = 0;
    int64_t n_distinct = 0;
    size_t table_bytes = 0, filter_bytes = 0;
    // host-buffer streaming (gki_count_reads / gki_count_kmers with host pointers)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // host-input pipeline with CPU packing lanes (count.cu, gki_count_reads): persistent worker threads, per lane a stream and buffers
    void *pipeline = nullptr;

    gki::IndexView view() const { return gki::IndexView{cells, kmers, fm}; }
};

namespace gki {
void destroy_count_table(gki_index *ix);   // count.cu (also frees the packing lanes)
}
