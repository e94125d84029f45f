/* gki.h -- C ABI of libgki.so: the B200 (sm_100a) implementation of graph_kmer_index's
 * read-k-mer counting hot path.
 *
 * The reference (ivargr/graph_kmer_index) is pure Python/numpy and has no FFI of its own; its
 * boundary is the Python API.  Each entry point below names the reference
 * function(s) it replaces (file:line relative to the reference repository);
 * graph_kmer_index_b200/*.py binds these through ctypes behind the reference's own class and
 * function names, and INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative gki_status; gki_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - Every array argument is a plain pointer that may live in DEVICE memory or in HOST memory
 *     (pageable or pinned); the library detects which (cudaPointerGetAttributes).  Host inputs are
 *     staged to the device (chunked and double-buffered for the streaming read batches), host
 *     outputs are copied back before the call returns; pageable host arrays of 32 MB or more
 *     cross the bus through pinned double buffers filled / drained by several host threads
 *     (GKI_HOST_COPY_THREADS, csrc/runtime.cu).  Calls whose arguments are all device
 *     pointers are asynchronous on `stream` unless stated otherwise.
 *   - Buffers are caller-owned and never freed or resized by the library; inputs are not modified.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - There is no CPU fallback: without a CUDA device every call fails with GKI_ERR_CUDA.
 */
#ifndef GKI_H
#define GKI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GKI_OK = 0,
    GKI_ERR_INVALID = -1,     /* bad argument (k > 31, modulo == 0, NULL pointer ...)            */
    GKI_ERR_CUDA = -2,        /* CUDA runtime error; text in gki_last_error()                    */
    GKI_ERR_UNSUPPORTED = -3, /* outside the supported envelope (modulo >= 2^32, n >= 2^31 ...)   */
    GKI_ERR_OVERFLOW = -4     /* caller-provided output capacity too small                       */
} gki_status;

typedef struct gki_index gki_index_t; /* opaque, device-resident CollisionFreeKmerIndex */
typedef void *gki_stream_t;

const char *gki_last_error(void);
int gki_version(void);
/* number of CUDA devices / current device selection (one process per GPU). */
int gki_device_count(int *count);
int gki_set_device(int device);
/* number of kernel launches issued by this library so far (bench.py's gpu_launches). */
int64_t gki_launch_count(void);

/* return the library's cached device scratch (stream-ordered pool) to the driver; synchronises the device */
int gki_release_scratch(void);

/* ------------------------------------------------------------------ K1: encoding + hashing */

/* flat_kmers.py:134-145 letter_sequence_to_numeric: ASCII (case-insensitive) -> a0 c1 g2 t3,
 * every other byte -> 0.  out[n] uint64 (the reference's dtype). */
int gki_encode_bases(const uint8_t *seq, int64_t n, uint64_t *out, gki_stream_t stream);

/* read_kmers.py:67-70 ReadKmers.get_kmers_from_read_dynamic (np.convolve(read, power_array(k),
 * 'valid')) for a batch of equal-length reads, plus read_kmers.py:21-26 (hashes of the
 * reverse-complemented read).  reads: n_reads rows of read_len ASCII bytes, row_stride bytes apart.
 * fwd / rc: (n_reads, read_len-k+1) uint64 row-major; either may be NULL.
 * fwd[r][i] = sum_j code(read[i+j]) * 4^j;  rc[r][i] = the same over the reverse-complemented read.
 * 1 <= k <= 31.  read_len < k yields no k-mers (documented deviation). */
int gki_hash_reads(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, int32_t k,
                   uint64_t *fwd, uint64_t *rc, gki_stream_t stream);

/* Ragged form of the above (read_kmers.py:14-49 from_fasta_file iterates reads of any length):
 * read r occupies seq[offsets[r] .. offsets[r+1]); its max(len-k+1,0) hashes are written at
 * fwd[out_offsets[r] ..) and rc[out_offsets[r] ..).  offsets/out_offsets have n_reads+1 entries. */
int gki_hash_reads_ragged(const uint8_t *seq, const int64_t *offsets, const int64_t *out_offsets, int64_t n_reads,
                          int32_t k, uint64_t *fwd, uint64_t *rc, gki_stream_t stream);

/* kmer_hashing.py:24-28 kmer_hashes_to_reverse_complement_hash (+ :12-22 scalar/chunked wrappers). */
int gki_revcomp_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream);
/* kmer_hashing.py:31-36 kmer_hashes_to_complement_hashes. */
int gki_complement_hashes(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream);
/* kmer_hashing.py:53-65 kmer_hashes_to_bases: out is (n, k) uint64 row-major, column j = base j. */
int gki_hashes_to_bases(const uint64_t *in, int64_t n, int32_t k, uint64_t *out, gki_stream_t stream);

/* ------------------------------------------------------------------ K2: index construction */

#define GKI_BUILD_SKIP_FREQUENCIES 1 /* from_flat_kmers(skip_frequencies=True): frequencies all 0 */

/* collision_free_kmer_index.py:422-467 CollisionFreeKmerIndex.from_flat_kmers (+ :267-293
 * set_frequencies): bucket = kmer % modulo, STABLE radix sort by bucket, run heads/lengths scattered
 * into the dense tables, payload columns permuted by the sort.
 *   in : kmers[n] u64, nodes[n] u32, ref_offsets[n] 8-byte items (moved bit-exactly), af[n] f32
 *   out: hashes_to_index[modulo] i32, n_kmers[modulo] u32 (0 for empty buckets),
 *        kmers_out/nodes_out/ref_out/af_out[n] permuted, freq_out[n] u16.
 * Any of nodes/ref_offsets/af (and the matching output) may be NULL.  perm_out (optional, u32[n])
 * receives the sort permutation (the reference's `sorting`, cfki:435) for columns of other dtypes.
 * Requires 1 <= n < 2^31 and 1 <= modulo < 2^32. */
int gki_index_build(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af,
                    int64_t n, uint64_t modulo, int32_t flags, int32_t *hashes_to_index, uint32_t *n_kmers,
                    uint64_t *kmers_out, uint32_t *nodes_out, uint64_t *ref_out, float *af_out, uint16_t *freq_out,
                    uint32_t *perm_out, gki_stream_t stream);

/* Hash-range partitioned build (the reference's closest analogue is the per-chunk index + merge of
 * command_line_interface.py:588-620).  Rank p of n_parts owns buckets [p*S, min((p+1)*S, modulo)), S = ceil(modulo/n_parts).
 * Step 1, on every rank: order the local FlatKmers by owner (stable; input order kept inside a part).
 *   perm_out[n] u32, counts_out[n_parts] i64.  The caller gathers its columns through perm_out (gki_gather) and
 *   exchanges them (all-to-all by counts).
 * Step 2, on every rank: build the slice of the index for its bucket range from the received entries (every k-mer must
 *   fall in the range).  Tables have bucket_hi-bucket_lo entries; hashes_to_index of non-empty buckets gets
 *   position_offset (= entries owned by lower ranks) added, so the concatenation over ranks IS the global index. */
int gki_partition_by_bucket_range(const uint64_t *kmers, int64_t n, uint64_t modulo, int32_t n_parts, uint32_t *perm_out,
                                  int64_t *counts_out, gki_stream_t stream);
/* The same two steps with ONE exchange: gki_partition_pack orders the local FlatKmers by owner (stable) and emits them as packed
 * 32-byte records {kmer, ref_offset, node | allele_frequency bits << 32, bucket} into records_out (device memory, n records,
 * n_parts <= 32; absent columns travel as zeros); the caller sends counts_out[p] records to rank p with a single all-to-all
 * and hands what it received, in source-rank order, to gki_index_build_records (arguments as gki_index_build_range). */
int gki_partition_pack(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af, int64_t n,
                       uint64_t modulo, int32_t n_parts, void *records_out, int64_t *counts_out, gki_stream_t stream);
int gki_index_build_records(const void *records, int64_t n, uint64_t modulo, uint64_t bucket_lo, uint64_t bucket_hi,
                            int64_t position_offset, int32_t flags, int32_t *hashes_to_index, uint32_t *n_kmers,
                            uint64_t *kmers_out, uint32_t *nodes_out, uint64_t *ref_out, float *af_out, uint16_t *freq_out,
                            gki_stream_t stream);
int gki_index_build_range(const uint64_t *kmers, const uint32_t *nodes, const uint64_t *ref_offsets, const float *af,
                          int64_t n, uint64_t modulo, uint64_t bucket_lo, uint64_t bucket_hi, int64_t position_offset,
                          int32_t flags, int32_t *hashes_to_index, uint32_t *n_kmers, uint64_t *kmers_out,
                          uint32_t *nodes_out, uint64_t *ref_out, float *af_out, uint16_t *freq_out, gki_stream_t stream);

/* Side indexes keyed by node or reference position: ReverseKmerIndex.from_flat_kmers (reverse_kmer_index.py:59-84,
 * key = node) and ReferenceKmerIndex.from_flat_kmers (reference_kmer_index.py:81-121, key = ref_offset).
 * Stable grouping of n entries by an integer key < n_keys <= 2^32 (keys: 4- or 8-byte unsigned, key_size says which):
 *   perm_out[n] u32      stable argsort of the keys; the caller moves its columns with gki_gather
 *   first_out[n_keys] u32, count_out[n_keys] u32   first sorted position and run length of every key, 0 when absent
 * flags GKI_GROUP_REFERENCE_INDEX: first_out is ReferenceKmerIndex.ref_position_to_index instead -- the smallest key
 *   is left unmarked (np.ediff1d(..., to_begin=0), :92) and every unmarked slot takes the value of the next marked
 *   slot to its right (fill_zeros_from_end, :16-21); count_out must be NULL.
 * Any output may be NULL. */
#define GKI_GROUP_REFERENCE_INDEX 1
int gki_group_by_key(const void *keys, int32_t key_size, int64_t n, uint64_t n_keys, int32_t flags, uint32_t *perm_out,
                     uint32_t *first_out, uint32_t *count_out, gki_stream_t stream);

/* out[i] = src[perm[i]] for items of item_size in {1,2,4,8} bytes (cfki:436-440 for any dtype). */
int gki_gather(const void *src, int32_t item_size, const uint32_t *perm, int64_t n, void *out, gki_stream_t stream);

/* flat_kmers.py:98-125 FlatKmers.get_new_without_singletons: keep[i] = 0 for the first occurrence
 * of every hash, 1 otherwise (the caller compacts).  keep[n] u8. */
int gki_mark_non_first_occurrences(const uint64_t *hashes, int64_t n, uint8_t *keep, gki_stream_t stream);

/* ------------------------------------------------------------------ K3: resident index, lookup, counting */

/* Upload a CollisionFreeKmerIndex (cfki:176-189 attribute arrays; npz layout cfki:393-402) and lay it
 * out for probing: {hashes_to_index, n_kmers} interleaved into one 8-byte cell per bucket.
 * ref_offsets/frequencies/af may be NULL (only gki_lookup_hits and the frequency gates need them).
 * flags: reserved, pass 0. */
int gki_index_create(const int32_t *hashes_to_index, const uint32_t *n_kmers, const uint64_t *kmers,
                     const uint32_t *nodes, const uint64_t *ref_offsets, const uint16_t *frequencies,
                     const float *af, int64_t n, uint64_t modulo, int32_t flags, gki_index_t **out,
                     gki_stream_t stream);
int gki_index_destroy(gki_index_t *index);
/* introspection: n entries, modulo, max node id (cfki:237-238), device bytes, L2 Bloom filter in use, number of
 * non-empty buckets.  Any output pointer may be NULL. */
int gki_index_info(const gki_index_t *index, int64_t *n, uint64_t *modulo, int64_t *max_node,
                   int64_t *device_bytes, int32_t *has_filter, int64_t *nonempty_buckets);

/* Build the counting structure now (otherwise the first counting call builds it): a Bloom filter sized to stay in
 * L2 + a bucketised table over the distinct k-mers.  k = k-mer length of the reads that will be counted (lets both
 * strands of a read position share one canonical key); k = 0 if unknown.  The first call fixes the mode; results
 * never depend on it.  Synchronises `stream`. */
int gki_prepare_counting(gki_index_t *index, int32_t k, gki_stream_t stream);

/* cfki:30-31 CounterKmerIndex.reset (intended meaning: zero every counter). */
int gki_reset_counts(gki_index_t *index, gki_stream_t stream);

/* cfki:33-37 CounterKmerIndex.count_kmers: for every query k-mer present in the index add one to
 * that k-mer's counter; absent k-mers are ignored.  queries[nq] u64. */
int gki_count_kmers(gki_index_t *index, const uint64_t *queries, int64_t nq, gki_stream_t stream);

/* Fused K1 -> K3 (read_kmers.py:14-26 then cfki:33-37): hash every k-mer of every read (forward and,
 * if both_strands, the reverse-complemented read) and count it; hashes never touch HBM. */
int gki_count_reads(gki_index_t *index, const uint8_t *reads, int64_t n_reads, int32_t read_len,
                    int64_t row_stride, int32_t k, int32_t both_strands, gki_stream_t stream);

#define GKI_COUNTS_WRAP_UINT16 1 /* weight = counter mod 2^16 (the reference's uint16 Counter, cfki:27) */

/* Read ingestion ahead of K1 (the reference's own form is the per-line Python loop of read_kmers.py:14-49).
 * gki_pack_reads (host only, no device needed): 2-bit packing of an ASCII read matrix on the CPU with n_threads threads
 * (<= 0: all but one hardware thread).  Clean rows -- every byte one of ACGTacgt -- are appended to `packed` in input order,
 * ceil(read_len/32) 64-bit words each, base i at bits 2*(i%32) of word i/32, a0 c1 g2 t3 (flat_kmers.py:134-145), unused
 * bits zero; rows holding any other byte are skipped and their indices written to dirty_index (first dirty_cap of them, in
 * order; may be NULL).  packed must hold n_reads rows.  flags: GKI_PACK_FORCE_SCALAR selects the table-driven path.
 * gki_count_packed_reads: gki_count_reads for such clean packed rows (host or device pointer).
 * gki_count_reads on a large host batch uses the same packing internally: GKI_PACK_THREADS lanes (default hardware
 * threads - 2; 0 disables) pack chunks into pinned buffers; whether the copy engine moves other chunks as ASCII at the same
 * time is decided per handle from the measured rate of the first calls (GKI_PIPELINE_DMA=0/1 overrides). */
#define GKI_PACK_FORCE_SCALAR 1
int gki_pack_reads(const uint8_t *reads, int64_t n_reads, int32_t read_len, int64_t row_stride, uint64_t *packed,
                   int64_t *dirty_index, int64_t dirty_cap, int64_t *n_clean, int64_t *n_dirty, int32_t n_threads, int32_t flags);
int gki_count_packed_reads(gki_index_t *index, const uint64_t *packed, int64_t n_reads, int32_t read_len, int32_t k,
                           int32_t both_strands, gki_stream_t stream);
/* host DRAM read bandwidth (GB/s) of n_threads threads (<= 0: as many as the packing lanes) summing `bytes` bytes of a host
 * buffer: the ceiling of the host side of gki_count_reads, measured on the caller's own read batch */
int gki_host_read_bandwidth(const void *host, int64_t bytes, int32_t n_threads, double *gb_per_s);

/* FASTA / FASTQ in front of the packer.  gki_fastx_open maps the file and finds its sequence lines with all host threads:
 * FASTA -- every line that does not start with '>' (exactly what ReadKmers.from_fasta_file treats as a read,
 * read_kmers.py:16-18, multi-line records included), FASTQ (first byte '@') -- the second line of every four; lines are
 * stripped of surrounding blanks (line.strip(), read_kmers.py:21).  gki_fastx_lines copies out the n_reads (offset, length)
 * pairs.  gki_count_fastx counts every sequence line of at least k bases like gki_count_reads would (lines are grouped by
 * length; large groups are packed to 2 bits straight from the mapping by the packing lanes); *n_kmers receives the number
 * of k-mers looked up. */
typedef struct gki_fastx gki_fastx_t;
int gki_fastx_open(const char *path, gki_fastx_t **out, int64_t *n_reads, int32_t *max_len, int32_t *format);
int gki_fastx_lines(const gki_fastx_t *file, int64_t *offsets, int32_t *lengths);
int gki_count_fastx(gki_index_t *index, const gki_fastx_t *file, int32_t k, int32_t both_strands, int64_t *n_kmers, gki_stream_t stream);
int gki_fastx_close(gki_fastx_t *file);

/* cfki:39-40 CounterKmerIndex.get_node_counts: out[node] = sum over entries e with nodes[e]==node of
 * counter[kmers[e]], as float64.  n_out must be >= max(min_nodes, max_node+1); out is overwritten. */
int gki_node_counts(gki_index_t *index, double *out, int64_t n_out, int32_t flags, gki_stream_t stream);
/* counter[kmers[e]] for every entry e (what `self.counter[self.kmers]` evaluates to, cfki:40). */
int gki_entry_counts(gki_index_t *index, uint32_t *out, gki_stream_t stream);

#define GKI_PROBE_SKIP_BUCKET0 1 /* cython_kmer_index.pyx:59-60, 84-85 */
/* cfki:210-212 map_kmers (kmer_mapper.map_kmers_to_graph_index): node_counts[nodes[e]] += 1 for every
 * (query, entry e) with kmers[e]==query [, frequencies[e] <= max_frequency if max_frequency >= 0].
 * node_counts[n_nodes] u64 is ACCUMULATED into (caller zeroes it); entries whose node >= n_nodes are skipped. */
int gki_map_kmers(gki_index_t *index, const uint64_t *queries, int64_t nq, uint64_t *node_counts, int64_t n_nodes,
                  int32_t flags, int32_t max_frequency, gki_stream_t stream);
/* cfki:214-216 has_kmers (kmer_mapper.in_graph_index): out[i] = 1 if queries[i] is in the index. */
int gki_has_kmers(gki_index_t *index, const uint64_t *queries, int64_t nq, uint8_t *out, int32_t flags,
                  gki_stream_t stream);
/* cython_kmer_index.pyx:47-109 CythonKmerIndex.get: rows [node, ref_offset, query index, frequency,
 * uint64(1000*af)] for every hit, ordered by query then entry.  out is (5, capacity) u64 row-major;
 * *n_hits receives the total (call with out=NULL to size).  max_bucket / max_frequency < 0 disable
 * the .pyx gates (pyx:62-63, 70-71); pass 10000 / 20 and GKI_PROBE_SKIP_BUCKET0 to reproduce them.
 * Synchronises `stream`. */
int gki_lookup_hits(gki_index_t *index, const uint64_t *queries, int64_t nq, int32_t flags, int64_t max_bucket,
                    int32_t max_frequency, uint64_t *out, int64_t capacity, int64_t *n_hits, gki_stream_t stream);

/* Positions form of the hit list (cfki:303-315 `hit_positions + start`, cfki:354-391 the *_from_multiple_kmers
 * loops): entries[c] = index position of hit c, query_index[c] = its query; same order and gates as
 * gki_lookup_hits.  The caller gathers whatever columns it needs with their own dtypes. */
int gki_lookup_entries(gki_index_t *index, const uint64_t *queries, int64_t nq, int32_t flags, int64_t max_bucket,
                       int32_t max_frequency, int64_t *entries, int64_t *query_index, int64_t capacity, int64_t *n_hits,
                       gki_stream_t stream);
/* `counter[keys]` for arbitrary keys (cfki:40): out[i] = current counter of queries[i], 0 when absent. */
int gki_query_counts(gki_index_t *index, const uint64_t *queries, int64_t nq, uint32_t *out, gki_stream_t stream);

/* ------------------------------------------------------------------ multi-GPU
 * Reads shard over the ranks, the index is replicated, the ranks' node-count vectors are summed by ONE
 * NCCL all-reduce.  Reference analogue: the parent summing / concatenating its workers' results (shared_mem.py:164-171, cfki:222-232).
 * gki_allreduce_counts takes the host application's ncclComm_t (as void*: this header does not include nccl.h) and sums `counts`
 * (device memory, n elements of float64 -- exact below 2^53, the dtype of get_node_counts, cfki:39-40 -- or uint64) in place over its
 * ranks, on `stream`.  NCCL is bound at run time (libnccl.so.2 of the process, or GKI_NCCL_LIBRARY).  A host without a communicator
 * of its own gets one from gki_nccl_unique_id (rank 0; send the 128 bytes to every rank by any means) + gki_nccl_comm_create. */
#define GKI_COUNTS_FLOAT64 0
#define GKI_COUNTS_UINT64 1
int gki_nccl_unique_id(void *id128);
int gki_nccl_comm_create(const void *id128, int32_t rank, int32_t world_size, void **nccl_comm_out);
int gki_nccl_comm_destroy(void *nccl_comm);
int gki_allreduce_counts(void *nccl_comm, void *counts, int64_t n, int32_t dtype, gki_stream_t stream);


/* ------------------------------------------------------------------ DenseKmerFinder (BASELINE config 5)
 * The variant graph is passed as flat CSR arrays: seq_offsets[n_nodes+1] / seq (base codes 0..3; an empty node is a
 * dummy node), edge_offsets[n_nodes+1] / edges, is_linear[n_nodes] (linear-ref node or linear-ref dummy node),
 * allele_frequencies[n_nodes] f64, n_in_edges[n_nodes]. */
typedef struct gki_finder gki_finder_t;

/* critical_graph_paths.py:42-104 CriticalGraphPaths.from_graph.  nodes_out u32 / offsets_out u16 of `capacity`
 * entries (n_nodes is always enough); *n_out = number of critical positions. */
int gki_critical_paths(const int64_t *seq_offsets, const int64_t *edge_offsets, const int32_t *edges, const uint8_t *is_linear,
                       const int32_t *n_in_edges, int64_t n_nodes, const int64_t *chromosome_start_nodes, int32_t n_chromosomes,
                       int32_t k, uint32_t *nodes_out, uint16_t *offsets_out, int64_t capacity, int64_t *n_out,
                       gki_stream_t stream);

/* kmer_finder.py:179-434 DenseKmerFinder.find / find_only_kmers_starting_at_position, in two calls: prepare uploads
 * the graph, walks every chain of starting points once to count the rows (*n_rows); fill walks again and writes the
 * rows in the reference's order: kmers i64, nodes i32, start_nodes i32, start_offsets i16, allele_frequencies f64
 * (kf:54-58).  crit_index[crit_len] is CriticalGraphPaths._index (critical_graph_paths.py:11-19); start_nodes /
 * start_offsets list the starting points in processing order (kf:192-214), chain_first[n_chains+1] groups the ones a
 * single walker must handle one after the other; store_flags (optional) is only_store_nodes as a per-node byte;
 * force_follow_flags (optional) marks, per node, `force_follow` of kf:385-388: only_follow_nodes is applied by the
 * caller to the edge lists (a node with successors in the set keeps only those), and the successors of a flagged node
 * are all followed even when max_variant_nodes is reached;
 * treated_slots (power of two) sizes the set behind `_positions_treated` (kf:311-319). */
int gki_finder_prepare(const int64_t *seq_offsets, const uint8_t *seq, const int64_t *edge_offsets, const int32_t *edges,
                       const uint8_t *is_linear, const double *allele_frequencies, int64_t n_nodes, const uint16_t *crit_index,
                       int64_t crit_len, const uint8_t *store_flags, const uint8_t *force_follow_flags, const int32_t *start_nodes,
                       const int32_t *start_offsets, int64_t n_starts, const int64_t *chain_first, int64_t n_chains, int32_t k,
                       int32_t max_variant_nodes, int32_t one_node_per_kmer, int32_t early_stop, int64_t treated_slots,
                       gki_finder_t **out, int64_t *n_rows, gki_stream_t stream);
int gki_finder_fill(gki_finder_t *finder, int64_t *kmers, int32_t *nodes, int32_t *start_nodes, int16_t *start_offsets,
                    double *allele_frequencies, gki_stream_t stream);
int gki_finder_destroy(gki_finder_t *finder);

/* ------------------------------------------------------------------ synthetic workloads + calibration
 * (bench / test support; graph_kmer_index_b200/synthetic.py is the bit-identical host mirror) */
int gki_synth_genome(uint8_t *codes, int64_t length, gki_stream_t stream);
int gki_synth_flat_kmers(const uint8_t *genome_codes, int64_t n_entries, int64_t n_nodes, int32_t k, uint64_t *hashes,
                         uint32_t *nodes, uint64_t *ref_offsets, float *af, gki_stream_t stream);
int gki_synth_reads(const uint8_t *genome_codes, int64_t genome_len, int64_t first_read, int64_t n_reads,
                    int32_t read_len, int32_t p_hit_permille, int32_t n_permille, uint8_t *reads, gki_stream_t stream);
/* random 8-byte gathers over a table of table_bytes (<= 32 GB, power of two not required): the measured
 * random-access ceiling K3 is compared with.  dependent_loads bit 0: every gather is followed by a dependent second
 * one; bit 1: loads ask for a 64-byte L2 fill (ld.global.nc.L2::64B) instead of the default whole line.
 * *ms receives the kernel time. Synchronous. */
int gki_calibrate_random_gather(int64_t table_bytes, int64_t n_gathers, int32_t dependent_loads, float *ms);
/* n random stores or atomics (the measurements the index-build design rests on).  mode 0/1/2: 32/16/8-byte stores to random
 * slots of an n-slot array; 3: returning atomicAdd on n_bins random counters; 4: the same without a return value; 5: returning
 * atomicAdd picks a slot inside the counter's own bin of a (n_bins x n/n_bins) array and a 32-byte record is stored there;
 * 6: 32-byte records to random slots with one 256-bit store each. */
int gki_calibrate_scatter(int64_t n, int32_t mode, int64_t n_bins, float *ms);
/* scattered 32-byte stores where `group` consecutive lanes write consecutive records, into n_slots slots, from `ctas` SMs */
int gki_calibrate_store_groups(int64_t n, int32_t group, int64_t n_slots, int32_t ctas, float *ms);
int gki_calibrate_copy(int64_t bytes, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* GKI_H */
