/* oracle/gki_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's (ivargr/graph_kmer_index) counting hot path, used to
 * check the CUDA product path and as the timed CPU baseline ("port") in bench.py.  Nothing in
 * graph_kmer_index_b200/ links or loads this file.  Parity status: pinned -- checked against the
 * numpy oracle, the reference's known-answer vectors and tests/golden/ by tests/test_oracle_golden.py.
 *
 * Citations are file:line in the reference repository.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* flat_kmers.py:134-145: a/A->0 c/C->1 g/G->2 t/T->3, n, m and everything else -> 0 */
static inline int base_code(uint8_t c, int *valid) {
    switch (c) {
        case 'a': case 'A': *valid = 1; return 0;
        case 'c': case 'C': *valid = 1; return 1;
        case 'g': case 'G': *valid = 1; return 2;
        case 't': case 'T': *valid = 1; return 3;
        default: *valid = 0; return 0;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the bench's reference arm asks for the cores it may use instead */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* flat_kmers.py:134-145 */
void orc_encode(const uint8_t *seq, int64_t n, uint64_t *out) {
    for (int64_t i = 0; i < n; i++) { int v; out[i] = (uint64_t)base_code(seq[i], &v); }
}

/* read_kmers.py:67-70 (np.convolve(read, power_array(k), 'valid')): out[i] = sum_j read[i+j]*4^j.
 * read_kmers.py:21-26: the reverse strand is the hash of the reverse-complemented read string:
 * rcread[m] = complement(read[L-1-m]) for ACGT, every other character stays non-ACGT (code 0). */
static void hash_one_read(const uint8_t *read, int L, int k, uint64_t *fwd, uint64_t *rc) {
    int nk = L - k + 1;
    if (nk <= 0) return;
    for (int i = 0; i < nk; i++) {
        uint64_t f = 0, r = 0, p = 1;
        for (int j = 0; j < k; j++) {
            int v; int c = base_code(read[i + j], &v);
            f += (uint64_t)c * p;
            if (rc) {
                int v2; int c2 = base_code(read[L - 1 - (i + j)], &v2);
                r += (uint64_t)(v2 ? 3 - c2 : 0) * p;
            }
            p *= 4;
        }
        fwd[i] = f;
        if (rc) rc[i] = r;
    }
}

void orc_hash_reads(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride,
                    int32_t k, uint64_t *fwd, uint64_t *rc) {
    int64_t nk = read_len - k + 1;
    if (nk <= 0) return;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_reads; r++)
        hash_one_read(reads + r * row_stride, read_len, k, fwd + r * nk, rc ? rc + r * nk : NULL);
}

/* kmer_hashing.py:24-28: sum_j (3-base[j]) * 4^(k-1-j) */
void orc_revcomp_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = in[i], r = 0;
        for (int j = 0; j < k; j++) { r = (r << 2) | (3 - (x & 3)); x >>= 2; }
        out[i] = r;
    }
}

/* kmer_hashing.py:31-36 */
void orc_complement_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = in[i], r = 0;
        for (int j = 0; j < k; j++) r |= (uint64_t)(3 - ((x >> (2 * j)) & 3)) << (2 * j);
        out[i] = r;
    }
}

/* collision_free_kmer_index.py:422-457 with a stable sort by bucket (counting sort):
 * hashes = kmers % modulo; entries ordered by bucket keeping input order inside a bucket;
 * hashes_to_index[b] = first position of bucket b (0 when empty), n_kmers[b] = its size.
 * set_frequencies (cfki:267-293): per entry, number of distinct ref_offsets among the entries of
 * the same k-mer (they all sit in one bucket). */
int orc_build(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af,
              int64_t n, uint64_t modulo, int32_t skip_frequencies,
              int32_t *hashes_to_index, uint32_t *n_kmers,
              uint64_t *kmers_o, uint32_t *nodes_o, uint64_t *ref_o, float *af_o, uint16_t *freq_o) {
    memset(hashes_to_index, 0, sizeof(int32_t) * modulo);
    memset(n_kmers, 0, sizeof(uint32_t) * modulo);
    for (int64_t i = 0; i < n; i++) n_kmers[kmers[i] % modulo]++;
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * modulo);
    if (!cursor) return -1;
    int64_t run = 0;
    for (uint64_t b = 0; b < modulo; b++) {
        cursor[b] = run;
        if (n_kmers[b]) hashes_to_index[b] = (int32_t)run;
        run += n_kmers[b];
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t p = cursor[kmers[i] % modulo]++;
        kmers_o[p] = kmers[i]; nodes_o[p] = nodes[i]; ref_o[p] = ref_offsets[i]; af_o[p] = af[i];
    }
    free(cursor);
    memset(freq_o, 0, sizeof(uint16_t) * n);
    if (!skip_frequencies) {
        /* first[a] = 1 iff no earlier entry of the bucket has the same (k-mer, ref_offset) pair */
        uint8_t *first = (uint8_t *)malloc(n);
        if (!first) return -1;
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t a = 0; a < n; a++) {
            int64_t s = hashes_to_index[kmers_o[a] % modulo];
            uint8_t f = 1;
            for (int64_t c = a - 1; c >= s; c--)
                if (kmers_o[c] == kmers_o[a] && ref_o[c] == ref_o[a]) { f = 0; break; }
            first[a] = f;
        }
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t e = 0; e < n; e++) {
            uint64_t b = kmers_o[e] % modulo;
            int64_t s = hashes_to_index[b], t = s + n_kmers[b];
            uint32_t cnt = 0;
            for (int64_t a = s; a < t; a++) cnt += (kmers_o[a] == kmers_o[e]) & first[a];
            freq_o[e] = (uint16_t)cnt;
        }
        free(first);
    }
    return 0;
}

/* The k-mer and node columns and both tables of orc_build (same stable order, same values), built by all host threads: thread t
 * owns a contiguous range of buckets, scans the bucket ids of every entry and places those of its range in input order.  Used by
 * the bench's CPU arm for the human-scale index (1 B entries), where the serial counting sort above takes minutes; checked
 * against orc_build in tests/test_oracle_golden.py. */
int orc_build_kmers_nodes(const uint64_t *kmers, const uint32_t *nodes, int64_t n, uint64_t modulo,
                          int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_o, uint32_t *nodes_o) {
    uint32_t *bucket = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    if (!bucket) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) bucket[i] = (uint32_t)(kmers[i] % modulo);
    int T = orc_num_threads();
    int64_t *range_total = (int64_t *)calloc((size_t)T + 1, sizeof(int64_t));
    if (!range_total) { free(bucket); return -1; }
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int t = 0, nt = 1;
#endif
        uint64_t lo = modulo * (uint64_t)t / (uint64_t)nt, hi = modulo * (uint64_t)(t + 1) / (uint64_t)nt;
        memset(n_kmers + lo, 0, sizeof(uint32_t) * (hi - lo));
        int64_t mine = 0;
        for (int64_t i = 0; i < n; i++) {
            uint32_t b = bucket[i];
            if (b >= lo && b < hi) { n_kmers[b]++; mine++; }
        }
        range_total[t + 1] = mine;
#pragma omp barrier
#pragma omp single
        for (int q = 0; q < nt; q++) range_total[q + 1] += range_total[q];
        int64_t run = range_total[t];
        for (uint64_t b = lo; b < hi; b++) {          /* hashes_to_index doubles as the bucket cursor while placing */
            hashes_to_index[b] = (int32_t)run;
            run += n_kmers[b];
        }
        for (int64_t i = 0; i < n; i++) {
            uint32_t b = bucket[i];
            if (b >= lo && b < hi) {
                int64_t p = hashes_to_index[b]++;
                kmers_o[p] = kmers[i];
                nodes_o[p] = nodes[i];
            }
        }
        for (uint64_t b = lo; b < hi; b++) hashes_to_index[b] = n_kmers[b] ? hashes_to_index[b] - (int32_t)n_kmers[b] : 0;
    }
    free(range_total);
    free(bucket);
    return 0;
}

/* Probe of one query, cython_kmer_index.pyx:57-72 / collision_free_kmer_index.py:303-309, without
 * the .pyx's three gates: every entry of the bucket whose k-mer equals the query gets +1.
 * Afterwards counts[e] == counter[kmers[e]] of CounterKmerIndex (cfki:33-40). */
static inline void probe_count(const int32_t *h2i, const uint32_t *nk, const uint64_t *ikm, uint64_t modulo,
                               uint64_t q, uint32_t *entry_counts) {
    uint64_t h = q % modulo;
    uint32_t n = nk[h];
    if (!n) return;
    int64_t pos = h2i[h];
    for (uint32_t j = 0; j < n; j++)
        if (ikm[pos + j] == q) {
#pragma omp atomic
            entry_counts[pos + j]++;
        }
}

void orc_count_kmers(const int32_t *h2i, const uint32_t *nk, const uint64_t *ikm, uint64_t modulo,
                     const uint64_t *queries, int64_t nq, uint32_t *entry_counts) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) probe_count(h2i, nk, ikm, modulo, queries[i], entry_counts);
}

/* read_kmers.py:14-26 + cfki:33-37 fused: hash forward (and reverse-complement) k-mers of every
 * read and count them. */
void orc_count_reads(const int32_t *h2i, const uint32_t *nk, const uint64_t *ikm, uint64_t modulo,
                     const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, int32_t k,
                     int32_t both_strands, uint32_t *entry_counts) {
    int nkm = read_len - k + 1;
    if (nkm <= 0) return;
#pragma omp parallel
    {
        uint64_t *f = (uint64_t *)malloc(sizeof(uint64_t) * nkm * 2);
        uint64_t *r = f + nkm;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_reads; i++) {
            hash_one_read(reads + i * row_stride, read_len, k, f, both_strands ? r : NULL);
            for (int j = 0; j < nkm; j++) probe_count(h2i, nk, ikm, modulo, f[j], entry_counts);
            if (both_strands)
                for (int j = 0; j < nkm; j++) probe_count(h2i, nk, ikm, modulo, r[j], entry_counts);
        }
        free(f);
    }
}

/* cfki:39-40: np.bincount(nodes, weights=counter[kmers], minlength) */
void orc_node_counts(const uint32_t *nodes, const uint32_t *entry_counts, int64_t n, double *out) {
    for (int64_t e = 0; e < n; e++) out[nodes[e]] += (double)entry_counts[e];
}
/* the same on all host threads (integer-valued additions below 2^53 are exact in any order) */
void orc_node_counts_parallel(const uint32_t *nodes, const uint32_t *entry_counts, int64_t n, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; e++) {
        if (!entry_counts[e]) continue;
        double w = (double)entry_counts[e];
#pragma omp atomic
        out[nodes[e]] += w;
    }
}

/* cython_kmer_index.pyx:47-109, two passes, gates optional. out may be NULL (count only).
 * out is (5, n_hits) row-major: node, ref_offset, query index, frequency, uint64(1000*af). */
int64_t orc_lookup_hits(const int32_t *h2i, const uint32_t *nk, const uint64_t *ikm, const uint32_t *nodes,
                        const uint64_t *ref, const uint16_t *freq, const float *af, uint64_t modulo,
                        const uint64_t *queries, int64_t nq, int32_t skip_bucket0, int64_t max_bucket,
                        int32_t max_frequency, uint64_t *out, int64_t n_hits_alloc) {
    int64_t c = 0;
    for (int64_t i = 0; i < nq; i++) {
        uint64_t h = queries[i] % modulo;
        if (skip_bucket0 && h == 0) continue;
        uint32_t n = nk[h];
        if (max_bucket >= 0 && (int64_t)n > max_bucket) continue;
        int64_t pos = h2i[h];
        for (uint32_t j = 0; j < n; j++) {
            if (ikm[pos + j] != queries[i]) continue;
            if (max_frequency >= 0 && freq[pos + j] > max_frequency) continue;
            if (out && c < n_hits_alloc) {
                out[0 * n_hits_alloc + c] = nodes[pos + j];
                out[1 * n_hits_alloc + c] = ref[pos + j];
                out[2 * n_hits_alloc + c] = (uint64_t)i;
                out[3 * n_hits_alloc + c] = freq[pos + j];
                out[4 * n_hits_alloc + c] = (uint64_t)(1000 * af[pos + j]);
            }
            c++;
        }
    }
    return c;
}
