"""Deterministic synthetic workloads -- counter-based splitmix64, so the same index,
reads and "genome" can be regenerated bit-exactly on the host (numpy, this file) and on the device
(csrc/synth.cu via gki_synth_*), without shipping files.

Model: a random "genome" G of U+k-1 bases; the index holds the U = ceil(N/2) windows of G, each emitted
twice with two different nodes (variant k-mers have >= 2 nodes) in a pseudo-random order; a fraction
`p_hit` of the reads are length-L substrings of G (random strand), the rest are random ACGT.  Both read
strands are counted, so ~p_hit/2 of the read k-mers hit the index.
"""
import numpy as np

SEED_GENOME, SEED_NODES, SEED_AF, SEED_READS, SEED_BASES, SEED_NS = 1, 2, 3, 4, 5, 6
PERM_MULT = 2654435761          # prime > 2^31 > N: j -> (j*PERM_MULT + PERM_ADD) % N is a bijection
PERM_ADD = 12345
_M64 = (1 << 64) - 1
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rnd(seed, i):
    with np.errstate(over="ignore"):
        return splitmix64(np.asarray(i, dtype=np.uint64) + np.uint64((seed * 0x632BE59BD9B4E019) & _M64))


def _packed_bases(seed, idx):
    """base code of element idx of stream `seed`: 32 bases per 64-bit draw."""
    idx = np.asarray(idx, dtype=np.uint64)
    return (rnd(seed, idx >> np.uint64(5)) >> ((idx & np.uint64(31)) << np.uint64(1))) & np.uint64(3)


def genome_codes(length):
    return _packed_bases(SEED_GENOME, np.arange(length, dtype=np.uint64)).astype(np.uint8)


def n_unique(n_entries):
    return (n_entries + 1) // 2


def genome_length(n_entries, k):
    return n_unique(n_entries) + k - 1


def flat_kmers(n_entries, n_nodes, k=31, codes=None):
    """-> (hashes uint64, nodes uint32, ref_offsets uint64, allele_frequencies float32)."""
    if codes is None:
        codes = genome_codes(genome_length(n_entries, k))
    u_count = n_unique(n_entries)
    win = np.zeros(u_count, dtype=np.uint64)
    c64 = codes.astype(np.uint64)
    for j in range(k):
        win |= c64[j:j + u_count] << np.uint64(2 * j)
    j = np.arange(n_entries, dtype=np.uint64)
    p = (j * np.uint64(PERM_MULT) + np.uint64(PERM_ADD)) % np.uint64(n_entries)
    u = p >> np.uint64(1)
    hashes = win[u.astype(np.int64)]
    nodes = (rnd(SEED_NODES, p) % np.uint64(n_nodes)).astype(np.uint32)
    af = (((rnd(SEED_AF, p) & np.uint64(1023)) + np.uint64(1)).astype(np.float32) * np.float32(1.0 / 1024.0))
    return hashes, nodes, u.astype(np.uint64), af.astype(np.float32)


def reads(n_reads, read_len, n_entries, k=31, p_hit_permille=100, n_permille=0, first_read=0, codes=None):
    """-> (n_reads, read_len) uint8 ASCII.  `first_read` offsets the read counter (sharding)."""
    glen = genome_length(n_entries, k)
    if codes is None:
        codes = genome_codes(glen)
    r = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    h = rnd(SEED_READS, r)
    from_genome = (h % np.uint64(1000)) < np.uint64(p_hit_permille)
    if glen < read_len:
        from_genome[:] = False
    span = np.uint64(max(glen - read_len + 1, 1))
    start = ((h >> np.uint64(16)) % span).astype(np.int64)
    strand = ((h >> np.uint64(12)) & np.uint64(1)).astype(bool)
    m = np.arange(read_len, dtype=np.int64)
    flat_idx = (r[:, None] * np.uint64(read_len) + m[None, :].astype(np.uint64))
    out = _packed_bases(SEED_BASES, flat_idx).astype(np.uint8)
    if from_genome.any():
        pos_f = start[:, None] + m[None, :]
        pos_r = start[:, None] + (read_len - 1 - m)[None, :]
        g_f = codes[np.minimum(pos_f, glen - 1)]
        g_r = 3 - codes[np.minimum(pos_r, glen - 1)]
        g = np.where(strand[:, None], g_r, g_f).astype(np.uint8)
        out = np.where(from_genome[:, None], g, out)
    ascii_ = _ACGT[out]
    if n_permille:
        is_n = (rnd(SEED_NS, flat_idx) % np.uint64(1000)) < np.uint64(n_permille)
        ascii_ = np.where(is_n, np.uint8(ord("N")), ascii_)
    return np.ascontiguousarray(ascii_, dtype=np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# Synthetic variant graph (BASELINE config 5): a linear backbone with a bubble every `spacing` bp on
# average: SNPs (two 1-bp alleles), deletions (ref allele of 1-5 bp vs an empty dummy node) and, optionally, insertions
# and bubbles nested inside an alternative allele.  Returned in the flat CSR form the finder takes
# (same keys as oracle/obgraph_standin.Graph.to_arrays()).
def variant_graph(n_variants, spacing=300, seed=0, p_deletion=0.2, p_insertion=0.0, p_nested=0.0, min_gap=1, tail=40):
    rng = np.random.default_rng(seed)
    seqs, edges, linear, af = {}, {}, [], {}
    nid = 0

    def new(seq, is_lin, freq=1.0):
        nonlocal nid
        n = nid
        nid += 1
        seqs[n] = seq
        edges[n] = []
        af[n] = freq
        if is_lin:
            linear.append(n)
        return n

    def rand_seq(length):
        return "".join("ACGT"[c] for c in rng.integers(0, 4, length))

    prev = [new(rand_seq(int(rng.integers(max(min_gap, 1), 2 * spacing))), True)]
    for _ in range(n_variants):
        f = float(np.round(rng.random() * 0.98 + 0.01, 3))
        r = rng.random()
        if r < p_deletion:                               # deletion: ref allele vs dummy
            ref = new(rand_seq(int(rng.integers(1, 6))), True, 1 - f)
            alts = [new("", False, f)]
        elif r < p_deletion + p_insertion:               # insertion: dummy on the linear ref vs inserted sequence
            ref = new("", False, 1 - f)
            alts = [new(rand_seq(int(rng.integers(1, 6))), False, f)]
        else:                                            # SNP (sometimes tri-allelic)
            ref = new(rand_seq(1), True, 1 - f)
            alts = [new(rand_seq(1), False, f)]
            if rng.random() < 0.1:
                alts.append(new(rand_seq(1), False, f / 2))
        heads, tails = [ref] + alts, [ref] + alts
        if rng.random() < p_nested:                      # a SNP nested inside a longer alternative allele
            a0 = new(rand_seq(int(rng.integers(1, 8))), False, f)
            s1, s2 = new(rand_seq(1), False, f), new(rand_seq(1), False, f / 3)
            a1 = new(rand_seq(int(rng.integers(1, 8))), False, f)
            edges[a0] += [s1, s2]
            edges[s1].append(a1)
            edges[s2].append(a1)
            heads.append(a0)
            tails.append(a1)
        for p in prev:
            edges[p] += heads
        nxt = new(rand_seq(int(rng.integers(max(min_gap, 1), 2 * spacing))), True)
        for t in tails:
            edges[t].append(nxt)
        prev = [nxt]
    last = new(rand_seq(tail), True)
    edges[prev[0]].append(last)
    # an insertion's dummy bridges two linear-ref nodes: it is a "linear-ref dummy node"
    return seqs, edges, linear, af
