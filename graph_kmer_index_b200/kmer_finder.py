"""Drop-in for graph_kmer_index/kmer_finder.py (DenseKmerFinder) and critical_graph_paths.py (CriticalGraphPaths);
the walks run in libgki.so (csrc/finder.cu).  Output rows, dtypes and row ORDER are the reference's.

The graph argument is an obgraph.Graph-like object.  The finder needs it as flat CSR arrays
(`graph_arrays`): objects that offer ``to_arrays()`` are used directly, anything else is converted through the same
methods the reference calls (kf:50,62,138,143,259,279,350,374,384; cgp:46-95).

``whitelist`` (kf:95-104, 130-132, 362-365) only decides which rows are stored, never the walk, so it is applied to the rows
the device returns.  ``only_follow_nodes`` (kf:385-388) redirects the walk: a node with successors in the set is left with
only those -- in the order the reference's ``set.intersection`` iterates them -- and ignores ``max_variant_nodes``; that is a
property of the node, so it is applied to the edge lists here (`follow_only`) and the device gets a per-node flag."""
import ctypes
import logging

import numpy as np

from . import _lib
from .flat_kmers import FlatKmers, FlatKmers2


def graph_arrays(graph):
    """-> dict(seq_offsets i64[n+1], seq u8 (codes 0..3), edge_offsets i64[n+1], edges i32, is_linear u8, allele_frequencies f64,
    n_in_edges i32, first_node, chromosome_start_nodes i64, node_to_ref_offset)"""
    if isinstance(graph, dict):
        return graph
    if hasattr(graph, "to_arrays"):
        return graph.to_arrays()
    n = int(graph.max_node_id()) + 1
    sizes = np.array([graph.get_node_size(i) for i in range(n)], dtype=np.int64)
    seq_offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=seq_offsets[1:])
    seq = np.zeros(int(seq_offsets[-1]), dtype=np.uint8)
    edge_offsets = np.zeros(n + 1, dtype=np.int64)
    edges = []
    for i in range(n):
        if sizes[i]:
            seq[seq_offsets[i]:seq_offsets[i + 1]] = np.asarray(graph.get_numeric_node_sequence(i)).astype(np.uint8)
        edges.extend(int(x) for x in graph.get_edges(i))
        edge_offsets[i + 1] = len(edges)
    edges = np.array(edges, dtype=np.int32)
    n_in = np.bincount(edges, minlength=n).astype(np.int32)
    return dict(seq_offsets=seq_offsets, seq=seq, edge_offsets=edge_offsets, edges=edges,
                is_linear=np.array([bool(graph.is_linear_ref_node_or_linear_ref_dummy_node(i)) for i in range(n)], dtype=np.uint8),
                allele_frequencies=np.asarray(graph.get_node_allele_frequencies(np.arange(n)), dtype=np.float64), n_in_edges=n_in,
                first_node=np.int64(graph.get_first_node()), node_to_ref_offset=np.asarray(graph.node_to_ref_offset),
                chromosome_start_nodes=np.array(list(graph.chromosome_start_nodes.values()), dtype=np.int64))


def follow_only(edge_offsets, edges, only_follow_nodes):
    """kf:385-388 as a graph rewrite -> (edge_offsets, edges, force_follow u8[n]).  Nodes with a successor in
    `only_follow_nodes` keep `only_follow_nodes.intersection(successors)` in that set's own iteration order (the order the
    reference's loop at kf:406 sees) and are flagged `force_follow`."""
    edge_offsets = np.asarray(edge_offsets, dtype=np.int64)
    edges = np.asarray(edges, dtype=np.int32)
    n = len(edge_offsets) - 1
    follow = set(int(x) for x in only_follow_nodes)
    force = np.zeros(n, dtype=np.uint8)
    hits = np.flatnonzero(np.isin(edges, np.fromiter(follow, dtype=np.int64, count=len(follow))))
    if len(hits) == 0:
        return edge_offsets, edges, force
    sources = np.unique(np.searchsorted(edge_offsets, hits, side="right") - 1)
    keep = np.ones(len(edges), dtype=bool)
    chosen = {}
    for u in sources:
        e0, e1 = int(edge_offsets[u]), int(edge_offsets[u + 1])
        chosen[int(u)] = list(follow.intersection(int(x) for x in edges[e0:e1]))
        keep[e0 + len(chosen[int(u)]):e1] = False
        force[u] = 1
    degrees = np.diff(edge_offsets)
    degrees[sources] = [len(chosen[int(u)]) for u in sources]
    new_offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(degrees, out=new_offsets[1:])
    new_edges = edges[keep].copy()
    for u, nodes in chosen.items():
        new_edges[new_offsets[u]:new_offsets[u + 1]] = nodes
    return new_offsets, new_edges, force


def starting_points_in_order(crit_nodes, crit_offsets, first_node, start_at, stop_at):
    """The starting points of DenseKmerFinder.find in processing order, as arrays.  The reference pops them off the end of the
    reversed list of critical paths (kf:192, 223-226): they come in the order of the critical paths, minus the first `start_at`,
    up to the first one on the node of path number `stop_at`; the beginning of the graph (`first_node`, when its node is too short to
    be critical, kf:212-214) comes first."""
    nodes = np.asarray(crit_nodes).astype(np.int64)
    offsets = np.asarray(crit_offsets).astype(np.int64)
    stop_at_node = int(nodes[stop_at]) if stop_at is not None and stop_at < len(nodes) else None
    if start_at is not None and start_at > 0:
        nodes, offsets = nodes[start_at:], offsets[start_at:]
    if first_node is not None:
        nodes, offsets = np.append(first_node, nodes), np.append(0, offsets)
    if stop_at_node is not None:
        at_stop = np.flatnonzero(nodes == stop_at_node)
        if len(at_stop):
            nodes, offsets = nodes[:at_stop[0]], offsets[:at_stop[0]]
    return nodes, offsets


def _c(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


class CriticalGraphPaths:
    """critical_graph_paths.py:5-104."""

    def __init__(self, nodes, offsets, index=None):
        self.nodes = nodes
        self.offsets = offsets
        self._index = index

    def _make_index(self):
        if len(self.nodes) == 0:
            self._index = np.zeros(0)
            return
        self._index = np.zeros(np.max(self.nodes) + 1, dtype=np.uint16)
        self._index[self.nodes] = self.offsets

    @classmethod
    def empty(cls):
        return cls(np.array([]), np.array([]), np.array([]))

    def is_critical(self, node, offset):
        if self._index is None:
            self._make_index()
        if node >= len(self._index):
            return False
        return self._index[node] == offset

    def __len__(self):
        return len(self.nodes)

    def __iter__(self):
        return ((node, offset) for node, offset in zip(self.nodes, self.offsets))

    @classmethod
    def from_graph(cls, graph, k):
        """critical_graph_paths.py:42-104 on the device (gki_critical_paths)."""
        a = graph_arrays(graph)
        n = len(a["seq_offsets"]) - 1
        nodes = np.empty(n, dtype=np.uint32)
        offsets = np.empty(n, dtype=np.uint16)
        count = ctypes.c_int64()
        chrom = _c(a["chromosome_start_nodes"], np.int64)
        _lib.call("gki_critical_paths", _lib.ptr(_c(a["seq_offsets"], np.int64)), _lib.ptr(_c(a["edge_offsets"], np.int64)),
                  _lib.ptr(_c(a["edges"], np.int32)), _lib.ptr(_c(a["is_linear"], np.uint8)), _lib.ptr(_c(a["n_in_edges"], np.int32)), n,
                  _lib.ptr(chrom), len(chrom), int(k), _lib.ptr(nodes), _lib.ptr(offsets), n, ctypes.byref(count), _lib.current_stream())
        return cls(nodes[:count.value].copy(), offsets[:count.value].copy())


class DenseKmerFinder:
    """kmer_finder.py:37-434: finds all possible k-mers in the graph."""

    def __init__(self, graph, k, critical_graph_paths=None, position_id=None, only_save_one_node_per_kmer=False, max_variant_nodes=4,
                 only_store_variant_nodes=False, start_at_critical_path_number=None, stop_at_critical_path_number=None, whitelist=None,
                 only_store_nodes=None, only_follow_nodes=None):
        self._only_follow_nodes = only_follow_nodes
        self._whitelist = None if whitelist is None else np.fromiter((int(x) for x in whitelist), dtype=np.int64)
        assert 1 <= k <= 31
        self._graph = graph
        self._arrays = graph_arrays(graph)
        self._k = k
        self._only_save_one_node_per_kmer = only_save_one_node_per_kmer
        self._max_variant_nodes = max_variant_nodes
        self._critical_graph_paths = critical_graph_paths
        self._position_id = position_id
        self._stop_at_critical_path_number = stop_at_critical_path_number
        self._start_at_critical_path_number = start_at_critical_path_number
        self._only_store_nodes = only_store_nodes
        self.kmers_found = []
        self._results = dict(kmers=np.zeros(0, np.int64), nodes=np.zeros(0, np.int32), start_nodes=np.zeros(0, np.int32),
                             start_offsets=np.zeros(0, np.int16), allele_frequencies=np.zeros(0, np.float64))

    # ---- results (kf:106-126) ----
    def get_found_kmers_and_nodes(self):
        return self._results["kmers"], self._results["nodes"]

    def get_flat_kmers(self, v="2"):
        r = self._results
        if v == "0" or v == "1":
            if v == "1":
                if self._position_id is None:
                    raise ValueError("get_flat_kmers(v='1') needs a position_id (obgraph.position_id.PositionId)")
                ref_offsets = self._position_id.get(r["start_nodes"], r["start_offsets"])
            else:
                ref_offsets = np.asarray(self._arrays["node_to_ref_offset"])[r["start_nodes"]] + r["start_offsets"]
            return FlatKmers(r["kmers"], r["nodes"], ref_offsets, r["allele_frequencies"])
        return FlatKmers2(r["kmers"], r["start_nodes"], r["start_offsets"], r["nodes"], r["allele_frequencies"])

    # ---- searches ----
    def _run(self, start_nodes, start_offsets, crit_index, early_stop):
        a = self._arrays
        n = len(a["seq_offsets"]) - 1
        k = self._k
        start_nodes = _c(start_nodes, np.int32)
        start_offsets = _c(start_offsets, np.int32)
        n_starts = len(start_nodes)
        # a start at offset 0 is overrun by the search before it (kf:333 only tests offset + 1), so it shares the
        # walker -- and the `_positions_treated` history -- of that search
        is_first = start_offsets != 0
        is_first[:1] = True
        first = np.flatnonzero(is_first)
        chain_first = _c(np.append(first, n_starts), np.int64)
        store = None
        if self._only_store_nodes is not None:
            store = np.zeros(n, dtype=np.uint8)
            keep = np.array([x for x in self._only_store_nodes if 0 <= x < n], dtype=np.int64)
            store[keep] = 1
        crit_index = _c(crit_index, np.uint16)
        total_bases = int(a["seq_offsets"][-1]) + n
        slots = 1 << max(10, int(np.ceil(np.log2(max(4 * total_bases, 1024)))))
        handle = ctypes.c_void_p()
        n_rows = ctypes.c_int64()
        edge_offsets, edges, force = a["edge_offsets"], a["edges"], None
        if self._only_follow_nodes is not None:
            edge_offsets, edges, force = follow_only(edge_offsets, edges, self._only_follow_nodes)
        keep_alive = [_c(a["seq_offsets"], np.int64), _c(a["seq"], np.uint8), _c(edge_offsets, np.int64), _c(edges, np.int32),
                      _c(a["is_linear"], np.uint8), _c(a["allele_frequencies"], np.float64)]
        _lib.call("gki_finder_prepare", *[_lib.ptr(x) for x in keep_alive], n, _lib.ptr(crit_index) if len(crit_index) else None,
                  len(crit_index), _lib.ptr(store), _lib.ptr(force), _lib.ptr(start_nodes), _lib.ptr(start_offsets), n_starts,
                  _lib.ptr(chain_first), len(first), k, int(self._max_variant_nodes), int(bool(self._only_save_one_node_per_kmer)), int(early_stop), slots,
                  ctypes.byref(handle), ctypes.byref(n_rows), _lib.current_stream())
        try:
            m = n_rows.value
            out = dict(kmers=np.empty(m, np.int64), nodes=np.empty(m, np.int32), start_nodes=np.empty(m, np.int32),
                       start_offsets=np.empty(m, np.int16), allele_frequencies=np.empty(m, np.float64))
            if m:
                _lib.call("gki_finder_fill", handle, _lib.ptr(out["kmers"]), _lib.ptr(out["nodes"]), _lib.ptr(out["start_nodes"]),
                          _lib.ptr(out["start_offsets"]), _lib.ptr(out["allele_frequencies"]), _lib.current_stream())
        finally:
            _lib.load().gki_finder_destroy(handle)
        if self._whitelist is not None:              # kf:130-132, 362-365: rows of other k-mers are not stored
            keep = np.isin(out["kmers"], self._whitelist)
            out = {key: v[keep] for key, v in out.items()}
        # results accumulate over calls like the reference's NpLists do
        self._results = out if len(self._results["kmers"]) == 0 else {key: np.concatenate([self._results[key], out[key]]) for key in out}

    def find_only_kmers_starting_at_position(self, node, offset):
        """kf:170-177."""
        self._run([int(node)], [int(offset)], np.zeros(0, dtype=np.uint16), early_stop=True)

    def find(self):
        """kf:179-244."""
        k = self._k
        if self._critical_graph_paths is None:
            self._critical_graph_paths = CriticalGraphPaths.from_graph(self._arrays, k)
        crit = self._critical_graph_paths
        a = self._arrays
        first_node = None
        if self._start_at_critical_path_number is None or self._start_at_critical_path_number == 0:
            if int(a["seq_offsets"][int(a["first_node"]) + 1] - a["seq_offsets"][int(a["first_node"])]) <= k:      # kf:212-214
                first_node = int(a["first_node"])
        nodes, offsets = starting_points_in_order(crit.nodes, crit.offsets, first_node, self._start_at_critical_path_number,
                                                  self._stop_at_critical_path_number)
        if crit._index is None:
            crit._make_index()
        self._run(nodes, offsets, np.asarray(crit._index), early_stop=False)
