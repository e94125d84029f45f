"""Stand-in for the parts of obgraph (third party, absent) that graph_kmer_index's DenseKmerFinder /
CriticalGraphPaths touch (TEST INFRASTRUCTURE).  Call sites in the reference: kmer_finder.py:50,62,83,117,138,143,
209,259,279,350,374,384; critical_graph_paths.py:46,52-53,62,74,85,95.

Graph.from_dicts(node_sequences, edges, linear_ref_nodes) is the constructor every reference test uses.  A node with an
empty sequence is a dummy node; it counts as "linear-ref dummy" when it bridges two consecutive linear-ref nodes.
The same class (plain numpy arrays) is what graph_kmer_index_b200.kmer_finder consumes, see `to_arrays`."""
import numpy as np

_CODE = {"a": 0, "c": 1, "g": 2, "t": 3}


class Graph:
    def __init__(self, node_sequences, edges, linear_ref_nodes, allele_frequencies=None):
        self._seqs = {int(n): s for n, s in node_sequences.items()}
        self._edges = {int(n): [int(x) for x in e] for n, e in edges.items()}
        self._linear = [int(n) for n in linear_ref_nodes]
        self._linear_set = set(self._linear)
        max_id = max(self._seqs)
        self.nodes = np.zeros(max_id + 1, dtype=np.int64)
        for n, s in self._seqs.items():
            self.nodes[n] = len(s)
        self._numeric = {n: np.array([_CODE.get(c, 0) for c in s.lower()], dtype=np.uint8) for n, s in self._seqs.items()}
        self._reverse = {n: [] for n in range(max_id + 1)}
        for a, outs in self._edges.items():
            for b in outs:
                self._reverse[b].append(a)
        self._linear_dummy = set()
        for a, b in zip(self._linear[:-1], self._linear[1:]):
            for d in self._edges.get(a, []):
                if self.nodes[d] == 0 and b in self._edges.get(d, []):
                    self._linear_dummy.add(d)
        self.chromosome_start_nodes = {1: self._linear[0]}
        self.node_to_ref_offset = np.zeros(max_id + 2, dtype=np.int64)
        off = 0
        for n in self._linear:
            self.node_to_ref_offset[n] = off
            off += self.nodes[n]
        af = np.ones(max_id + 1, dtype=float)
        if allele_frequencies is not None:
            for n, v in allele_frequencies.items():
                af[int(n)] = v
        self._af = af

    @classmethod
    def from_dicts(cls, node_sequences, edges, linear_ref_nodes, allele_frequencies=None):
        return cls(node_sequences, edges, linear_ref_nodes, allele_frequencies)

    def make_linear_ref_node_and_ref_dummy_node_index(self):
        """obgraph builds this index lazily; here it is computed in the constructor."""

    def linear_ref_nodes(self):
        return self._linear_set

    def get_first_node(self):
        return self._linear[0]

    def get_node_size(self, node):
        return int(self.nodes[node])

    def get_numeric_base_sequence(self, node, offset):
        return int(self._numeric[node][offset])

    def get_numeric_node_sequence(self, node):
        return self._numeric[node]

    def get_edges(self, node):
        return list(self._edges.get(int(node), []))

    def is_linear_ref_node_or_linear_ref_dummy_node(self, node):
        return int(node) in self._linear_set or int(node) in self._linear_dummy

    def get_node_allele_frequencies(self, nodes):
        return self._af[np.asarray(nodes).astype(np.int64)]

    def get_node_allele_frequency(self, node):
        return float(self._af[int(node)])

    def max_node_id(self):
        return len(self.nodes) - 1

    def get_reverse_edges_hashtable(self):
        return self._reverse

    # ---- flat arrays (CSR), the form the CUDA finder takes -------------------------------------------------
    def to_arrays(self):
        n = len(self.nodes)
        seq_offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.nodes, out=seq_offsets[1:])
        seq = np.zeros(int(seq_offsets[-1]), dtype=np.uint8)
        for node, arr in self._numeric.items():
            seq[seq_offsets[node]:seq_offsets[node + 1]] = arr
        edge_offsets = np.zeros(n + 1, dtype=np.int64)
        flat = []
        for node in range(n):
            outs = self._edges.get(node, [])
            flat.extend(outs)
            edge_offsets[node + 1] = len(flat)
        rev_counts = np.array([len(self._reverse[node]) for node in range(n)], dtype=np.int32)
        is_linear = np.array([self.is_linear_ref_node_or_linear_ref_dummy_node(node) for node in range(n)], dtype=np.uint8)
        return dict(seq_offsets=seq_offsets, seq=seq, edge_offsets=edge_offsets, edges=np.array(flat, dtype=np.int32),
                    is_linear=is_linear, allele_frequencies=self._af.astype(np.float64), n_in_edges=rev_counts,
                    first_node=np.int64(self.get_first_node()), node_to_ref_offset=self.node_to_ref_offset,
                    chromosome_start_nodes=np.array(list(self.chromosome_start_nodes.values()), dtype=np.int64))


class PositionId:
    """obgraph.position_id.PositionId stand-in: a unique id per (node, offset): cumulative node start + offset."""

    def __init__(self, starts):
        self._starts = starts

    @classmethod
    def from_graph(cls, graph):
        starts = np.zeros(len(graph.nodes) + 1, dtype=np.int64)
        np.cumsum(np.maximum(graph.nodes, 1), out=starts[1:])
        return cls(starts)

    def get(self, nodes, offsets):
        return self._starts[np.asarray(nodes).astype(np.int64)] + np.asarray(offsets).astype(np.int64)
