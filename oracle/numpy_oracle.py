"""numpy restatement of the reference hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line (relative to ivargr/graph_kmer_index) it follows.
The restatement uses the same numpy primitives as the reference wherever the reference's result
is order-dependent, except that the bucket sort is *stable* (canonical payload order,
SURVEY.md section 8c(ii)).  Pinned by tests/test_oracle_golden.py.
"""
import numpy as np

DEFAULT_MODULO = 452930477  # collision_free_kmer_index.py:423


# --------------------------------------------------------------------------- encoding / hashing
def power_array(k):
    """kmer_hashing.py:4-5."""
    return np.power(4, np.arange(k - 1, -1, -1)).astype(np.uint64)


def reverse_power_array(k):
    """kmer_hashing.py:8-9."""
    return np.power(4, np.arange(k)).astype(np.uint64)


_CODE = np.zeros(256, dtype=np.uint64)      # flat_kmers.py:134-145: a/n/m/other->0 c->1 g->2 t->3
_VALID = np.zeros(256, dtype=bool)
for _ch, _v in (("a", 0), ("c", 1), ("g", 2), ("t", 3)):
    for _c in (_ch, _ch.upper()):
        _CODE[ord(_c)] = _v
        _VALID[ord(_c)] = True


def _as_ascii(seq):
    if isinstance(seq, str):
        return np.frombuffer(seq.encode("latin-1", "replace"), dtype=np.uint8)
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    seq = np.asarray(seq)
    assert seq.dtype == np.uint8
    return seq


def letter_sequence_to_numeric(seq):
    """flat_kmers.py:134-145 for str / ASCII-byte input (case-insensitive)."""
    return _CODE[_as_ascii(seq)]


def numeric_to_letter_sequence(numeric):
    """flat_kmers.py:147-154: 0..3 -> 'a','c','g','t' (lower case)."""
    return np.array(list("acgt"), dtype=object)[np.asarray(numeric).astype(np.int64)]


def sequence_to_kmer_hash(sequence):
    """snp_kmer_finder.py:19-26: sum(base[j] * 4**j), python int."""
    num = letter_sequence_to_numeric(sequence)
    return int(np.sum(num * reverse_power_array(len(num))))


def kmer_hashes_to_bases(hashes, k):
    """kmer_hashing.py:53-65: (N,k) uint64, column j = base j of the sequence."""
    h = np.asarray(hashes).astype(np.uint64)
    shifts = (2 * np.arange(k)).astype(np.uint64)
    return (h[:, None] >> shifts[None, :]) & np.uint64(3)


def kmer_hash_to_sequence(hash_, k):
    """snp_kmer_finder.py:14-16."""
    return "".join(numeric_to_letter_sequence(kmer_hashes_to_bases(np.array([hash_], dtype=np.uint64), k)[0]))


def read_kmer_hashes(read, k):
    """read_kmers.py:67-70: np.convolve(numeric, power_array(k), 'valid') == sum_j read[i+j]*4**j.
    Reads shorter than k give no k-mers (the reference's np.convolve swaps its arguments there and
    returns k-len+1 meaningless values; documented deviation)."""
    num = letter_sequence_to_numeric(read)
    if len(num) < k:
        return np.zeros(0, dtype=np.uint64)
    return np.convolve(num, power_array(k), mode="valid").astype(np.uint64)


def reverse_complement_ascii(read_u8):
    """read_kmers.py:24 (Bio.Seq reverse_complement): ACGT<->TGCA, case kept; every other character
    maps to another non-ACGT character, i.e. still encodes to 0."""
    table = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        table[a] = b
    return table[_as_ascii(read_u8)][::-1]


def hash_reads(reads_u8, k):
    """Batch form of read_kmers.py:14-26: for a (n_reads, L) uint8 ASCII matrix return
    (fwd, rc), each (n_reads, L-k+1) uint64; rc[r] = hashes of the reverse-complemented read."""
    reads_u8 = np.asarray(reads_u8, dtype=np.uint8)
    n, L = reads_u8.shape
    nk = max(L - k + 1, 0)
    fwd = np.zeros((n, nk), dtype=np.uint64)
    rc = np.zeros((n, nk), dtype=np.uint64)
    if nk == 0:
        return fwd, rc
    code = _CODE[reads_u8]
    ccode = np.where(_VALID[reads_u8], np.uint64(3) - code, np.uint64(0))[:, ::-1]
    for j in range(k):
        w = np.uint64(4) ** np.uint64(j)
        fwd += code[:, j:j + nk] * w
        rc += ccode[:, j:j + nk] * w
    return fwd, rc


def revcomp_hashes(hashes, k):
    """kmer_hashing.py:24-28: sum_j (3-base[j]) * 4**(k-1-j)."""
    assert k <= 31
    bases = kmer_hashes_to_bases(hashes, k)
    return np.sum((np.uint64(3) - bases) * power_array(k), axis=1).astype(np.uint64)


def complement_hashes(hashes, k):
    """kmer_hashing.py:31-36: complement without reversal."""
    assert k <= 31
    bases = kmer_hashes_to_bases(hashes, k)
    return np.sum((np.uint64(3) - bases) * reverse_power_array(k), axis=1).astype(np.uint64)


# --------------------------------------------------------------------------- index build
def frequencies(kmers_sorted, ref_offsets_sorted):
    """collision_free_kmer_index.py:267-293 (set_frequencies): for each entry the number of distinct
    ref_offsets among all entries holding the same k-mer; uint16 (wraps like a numpy-1 cast)."""
    n = len(kmers_sorted)
    if n == 0:
        return np.zeros(0, dtype=np.uint16)
    ro = np.asarray(ref_offsets_sorted)
    pairs = np.stack([np.asarray(kmers_sorted).astype(np.uint64), ro.astype(np.float64).view(np.uint64)
                      if ro.dtype.kind == "f" else ro.astype(np.uint64)], axis=1)
    uniq_pairs = np.unique(pairs, axis=0)
    uk, cnt = np.unique(uniq_pairs[:, 0], return_counts=True)
    pos = np.searchsorted(uk, pairs[:, 0])
    return cnt[pos].astype(np.uint64).astype(np.uint16)


def build_index(kmers, nodes, ref_offsets, allele_frequencies, modulo=DEFAULT_MODULO, skip_frequencies=False):
    """collision_free_kmer_index.py:422-467 with a STABLE bucket sort.  Returns a dict with the
    attribute names of CollisionFreeKmerIndex (cfki:176-189)."""
    kmers = np.asarray(kmers)
    n = len(kmers)
    if n == 0:
        raise IndexError("empty FlatKmers (reference raises IndexError at cfki:455)")
    hashes = kmers % np.uint64(modulo) if kmers.dtype == np.uint64 else kmers % modulo   # cfki:433
    sorting = np.argsort(hashes, kind="stable")                                         # cfki:435
    hashes = hashes[sorting]
    out_kmers = kmers[sorting]
    out_nodes = np.asarray(nodes)[sorting]
    out_ref = np.asarray(ref_offsets)[sorting]
    out_af = np.asarray(allele_frequencies)[sorting]
    head = np.ones(n, dtype=bool)                                                       # cfki:444-445
    head[1:] = hashes[1:] != hashes[:-1]
    unique_entry_positions = np.nonzero(head)[0]
    unique_hashes = hashes[unique_entry_positions].astype(np.int64)
    lookup = np.zeros(modulo, dtype=np.int32)                                           # cfki:453-454
    lookup[unique_hashes] = unique_entry_positions
    n_entries = np.diff(np.append(unique_entry_positions, n))                           # cfki:455
    n_kmers = np.zeros(modulo, dtype=np.uint32)                                         # cfki:456-457
    n_kmers[unique_hashes] = n_entries
    if skip_frequencies:
        freq = np.zeros(n, dtype=np.uint16)                                             # cfki:270-274
    else:
        freq = frequencies(out_kmers, out_ref)
    return dict(_hashes_to_index=lookup, _n_kmers=n_kmers, _nodes=out_nodes, _ref_offsets=out_ref,
                _kmers=out_kmers, _modulo=int(modulo), _frequencies=freq, _allele_frequencies=out_af)


def without_singletons(hashes, nodes, ref_offsets, allele_frequencies):
    """flat_kmers.py:98-125: drop the FIRST occurrence of every hash, keep the rest in order."""
    hashes = np.asarray(hashes)
    _, first = np.unique(hashes, return_index=True)
    keep = np.ones(len(hashes), dtype=bool)
    keep[first] = False
    return hashes[keep], np.asarray(nodes)[keep], np.asarray(ref_offsets)[keep], np.asarray(allele_frequencies)[keep]


# --------------------------------------------------------------------------- lookup / count
def index_get(index, kmer, max_hits=10):
    """collision_free_kmer_index.py:303-315."""
    h = int(kmer) % index["_modulo"]
    start = int(index["_hashes_to_index"][h])
    end = start + int(index["_n_kmers"][h])
    hit = np.where(index["_kmers"][start:end] == kmer)[0] + start
    freq = index["_frequencies"][hit]
    if len(hit) == 0 or freq[0] > max_hits:
        return None, None, None, None
    return index["_nodes"][hit], index["_ref_offsets"][hit], freq, index["_allele_frequencies"][hit]


def _probe(index, queries):
    """Vectorised restatement of the probe loop (cython_kmer_index.pyx:57-72, cfki:303-309):
    yields (query_idx, entry_idx) for every index entry whose k-mer equals the query and which
    lies in the query's bucket, ordered by query then by entry position."""
    q = np.asarray(queries).astype(np.uint64)
    mod = index["_modulo"]
    h = (q % np.uint64(mod)).astype(np.int64)
    n = index["_n_kmers"][h].astype(np.int64)
    start = index["_hashes_to_index"][h].astype(np.int64)
    qi = np.repeat(np.arange(len(q), dtype=np.int64), n)
    if len(qi) == 0:
        return qi, qi, h
    first = np.cumsum(n) - n
    ei = np.arange(len(qi), dtype=np.int64) - np.repeat(first, n) + np.repeat(start, n)
    match = index["_kmers"][ei].astype(np.uint64) == q[qi]
    return qi[match], ei[match], h


def lookup_hits(index, queries, skip_bucket0=True, max_bucket=10000, max_frequency=20):
    """cython_kmer_index.pyx:47-109 (CythonKmerIndex.get): (5, n_hits) uint64 rows
    [node, ref_offset, query_idx, frequency, uint64(1000*allele_frequency)].  The three gates are
    the .pyx's own (pyx:59-63, 70-71); pass skip_bucket0=False, max_bucket=None, max_frequency=None
    to disable them."""
    qi, ei, h = _probe(index, queries)
    keep = np.ones(len(qi), dtype=bool)
    if skip_bucket0:
        keep &= h[qi] != 0
    if max_bucket is not None:
        keep &= index["_n_kmers"][h[qi]] <= max_bucket
    if max_frequency is not None:
        keep &= index["_frequencies"][ei] <= max_frequency
    qi, ei = qi[keep], ei[keep]
    out = np.zeros((5, len(qi)), dtype=np.uint64)
    out[0] = index["_nodes"][ei]
    out[1] = index["_ref_offsets"][ei]
    out[2] = qi
    out[3] = index["_frequencies"][ei]
    out[4] = (np.float32(1000) * index["_allele_frequencies"][ei].astype(np.float32)).astype(np.uint64)
    return out


def has_kmers(index, queries):
    """collision_free_kmer_index.py:214-216 (kmer_mapper.in_graph_index): membership per query."""
    qi, _, _ = _probe(index, queries)
    out = np.zeros(len(np.asarray(queries)), dtype=bool)
    out[qi] = True
    return out


def kmer_counts(index, queries):
    """collision_free_kmer_index.py:20-37 (CounterKmerIndex.from_kmer_index + count_kmers): returns
    (unique index k-mers, count of query occurrences per unique k-mer). Absent queries are ignored."""
    uniq = np.unique(index["_kmers"].astype(np.uint64))
    q = np.asarray(queries).astype(np.uint64)
    pos = np.searchsorted(uniq, q)
    pos[pos >= len(uniq)] = 0
    found = uniq[pos] == q if len(uniq) else np.zeros(len(q), dtype=bool)
    return uniq, np.bincount(pos[found], minlength=len(uniq)).astype(np.int64)


def node_counts(index, queries, min_nodes=0):
    """collision_free_kmer_index.py:39-40 (get_node_counts): bincount(nodes, weights=counter[kmers],
    minlength=min_nodes) -> float64.  counter[kmer] = number of query occurrences of that k-mer."""
    uniq, counts = kmer_counts(index, queries)
    w = counts[np.searchsorted(uniq, index["_kmers"].astype(np.uint64))]
    return np.bincount(index["_nodes"].astype(np.int64), weights=w.astype(np.float64), minlength=min_nodes)


def map_kmers(index, queries, n_nodes, skip_bucket0=False, max_frequency=None):
    """collision_free_kmer_index.py:210-212 (kmer_mapper.map_kmers_to_graph_index restated from its
    call site): one count per (query, matching entry) on the entry's node."""
    qi, ei, h = _probe(index, queries)
    keep = np.ones(len(qi), dtype=bool)
    if skip_bucket0:
        keep &= h[qi] != 0
    if max_frequency is not None:
        keep &= index["_frequencies"][ei] <= max_frequency
    return np.bincount(index["_nodes"][ei[keep]].astype(np.int64), minlength=n_nodes).astype(np.uint64)


def read_node_counts(index, reads_u8, k, min_nodes=0, both_strands=True):
    """The metric path: hash every read (forward + reverse complement, read_kmers.py:21-26),
    count (cfki:33-37), bincount per node (cfki:39-40)."""
    fwd, rc = hash_reads(reads_u8, k)
    q = np.concatenate([fwd.ravel(), rc.ravel()]) if both_strands else fwd.ravel()
    return node_counts(index, q, min_nodes)


# ---- side indexes keyed by node / reference position ----------------------------------------------------
def reverse_index(hashes, nodes, ref_offsets):
    """reverse_kmer_index.py:59-84 with a stable argsort -> (nodes_to_index_positions u32, nodes_to_n_hashes u16 (wraps),
    hashes, ref_positions)."""
    nodes = np.asarray(nodes)
    order = np.argsort(nodes, kind="stable")
    s_nodes = nodes[order]
    first = np.zeros(int(s_nodes[-1]) + 1, dtype=np.uint32)
    n_kmers = np.zeros(int(s_nodes[-1]) + 1, dtype=np.uint16)
    heads = np.flatnonzero(np.concatenate([[True], s_nodes[1:] != s_nodes[:-1]]))
    first[s_nodes[heads]] = heads
    n_kmers[s_nodes[heads]] = (np.diff(np.concatenate([heads, [len(nodes)]])) & 0xFFFF).astype(np.uint16)
    return first, n_kmers, np.asarray(hashes)[order], np.asarray(ref_offsets)[order]


def reference_index(hashes, nodes, ref_offsets):
    """reference_kmer_index.py:81-121 with a stable argsort -> (ref_position_to_index u32, kmers (u32 if all < 2^32),
    ref_positions, nodes).  The first run is left unmarked (:92, to_begin=0) and unmarked slots take the next marked value
    to their right (:16-21)."""
    ref = np.asarray(ref_offsets)
    order = np.argsort(ref, kind="stable")
    s_ref = ref[order]
    kmers = np.asarray(hashes)[order]
    if kmers.max() < 2 ** 32:
        kmers = kmers.astype(np.uint32)
    heads = np.flatnonzero(s_ref[1:] != s_ref[:-1]) + 1            # run heads except the first run
    table = np.zeros(int(s_ref[-1]) + 1, dtype=np.uint32)
    table[s_ref[heads].astype(np.int64)] = heads
    out = np.zeros_like(table)
    nxt = 0
    for p in range(len(table) - 1, -1, -1):                        # small cases only
        if table[p]:
            nxt = table[p]
        out[p] = nxt
    return out, kmers, s_ref, np.asarray(nodes)[order]
