"""CPU restatement of CriticalGraphPaths.from_graph and DenseKmerFinder (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows graph_kmer_index/critical_graph_paths.py:42-104 and graph_kmer_index/kmer_finder.py:15-434 line by line,
but iteratively (explicit stack of window snapshots instead of recursion over two shared growing arrays), which is
the form the CUDA kernel (csrc/finder.cu) takes.  Pinned by tests/test_oracle_finder.py against the unmodified
reference run on the reference's own test graphs and on random SNP/indel graphs (tests/golden/finder_*.npz).

The graph is the flat-array form of oracle/obgraph_standin.Graph.to_arrays():
  seq_offsets[n+1], seq (codes 0..3), edge_offsets[n+1], edges, is_linear (linear-ref node or linear-ref dummy),
  allele_frequencies (float64), n_in_edges, first_node, chromosome_start_nodes.
"""
import numpy as np


class G:
    def __init__(self, arrays):
        self.__dict__.update(arrays)
        self.n_nodes = len(self.seq_offsets) - 1

    def size(self, node):
        return int(self.seq_offsets[node + 1] - self.seq_offsets[node])

    def base(self, node, offset):
        return int(self.seq[self.seq_offsets[node] + offset])

    def out(self, node):
        return [int(x) for x in self.edges[self.edge_offsets[node]:self.edge_offsets[node + 1]]]


def critical_paths(arrays, k):
    """critical_graph_paths.py:42-104 -> (nodes uint32, offsets uint16)"""
    g = G(arrays)
    nodes, offsets = [], []
    for start_node in g.chromosome_start_nodes:
        current, depth, bp = int(start_node), 0, 0
        while True:
            prev_depth = depth
            depth -= int(g.n_in_edges[current])
            if prev_depth > 1 and depth == 0:
                bp = 0
            size = g.size(current)
            if depth == 0 and size != 0:
                if bp <= k and bp + size >= k:
                    nodes.append(current)
                    offsets.append(k - bp - 1)
            nxt = g.out(current)
            depth += len(nxt)
            if len(nxt) == 0:
                break
            elif len(nxt) == 1:
                bp += size
                current = nxt[0]
            else:
                lin = [n for n in nxt if g.is_linear[n]]
                if len(lin) != 1:
                    raise Exception("Did not find 1 next node from node %d" % current)
                current = lin[0]
    return np.array(nodes, dtype=np.uint32), np.array(offsets, dtype=np.uint16)


def _critical_index(crit_nodes, crit_offsets):
    """critical_graph_paths.py:11-19 (note: nodes that are not critical keep index 0)"""
    if len(crit_nodes) == 0:
        return np.zeros(0, dtype=np.uint16)
    index = np.zeros(int(np.max(crit_nodes)) + 1, dtype=np.uint16)
    index[crit_nodes] = crit_offsets
    return index


def dense_kmer_finder(arrays, k, critical=None, max_variant_nodes=4, only_save_one_node_per_kmer=False, only_store_nodes=None,
                      only_position=None, only_follow_nodes=None):
    """kmer_finder.py:37-434.  only_position=(node, offset): find_only_kmers_starting_at_position (kf:170-177).
    Returns dict(kmers int64, nodes int32, start_nodes int32, start_offsets int16, allele_frequencies float64)."""
    g = G(arrays)
    out = {"kmers": [], "nodes": [], "start_nodes": [], "start_offsets": [], "af": []}
    treated = set()
    early_stop = only_position is not None
    if early_stop:
        crit_index = np.zeros(0, dtype=np.uint16)
        starting_points = [tuple(only_position)]
    else:
        if critical is None:
            critical = critical_paths(arrays, k)
        crit_index = _critical_index(*critical)
        starting_points = [(int(n), int(o)) for n, o in zip(*critical)][::-1]            # kf:192
        if g.size(int(g.first_node)) <= k:                                               # kf:212-214
            starting_points.append((int(g.first_node), 0))
    starting_set = set(starting_points)

    def is_critical(node, offset):
        return node < len(crit_index) and int(crit_index[node]) == offset

    def add_kmer(kmer, node, offset, wnodes):                                            # kf:128-168
        nodes = sorted(set(wnodes))
        af = min(float(g.allele_frequencies[n]) for n in nodes)
        if only_save_one_node_per_kmer:
            nodes = nodes[:1]
        for n in nodes:
            if only_store_nodes is not None and n not in only_store_nodes:
                continue
            out["kmers"].append(kmer)
            out["nodes"].append(n)
            out["start_nodes"].append(node)
            out["start_offsets"].append(offset)
            out["af"].append(af)

    while starting_points:
        crit_node, crit_offset = starting_points.pop()
        wbase, wnode = [], []          # current_bases / current_nodes from _current_path_start_position on
        nonempty = 0
        start_offset = crit_offset
        if not early_stop and start_offset >= k - 1:                                     # kf:229-230
            start_offset -= k - 1
        stack = []                     # frames: [children, next child index, hash, wbase, wnode, nonempty]
        node, offset, cur_hash = crit_node, start_offset, 0
        while True:
            # ---------------- search_from(node, offset, cur_hash) (kf:254-347) ----------------
            size = g.size(node)
            stopped = False
            if offset == 0 and size == 0:
                wbase.append(-1)
                wnode.append(node)
            while offset < size:
                if offset == k + 2 and size > offset + k + 1 and not early_stop:
                    # _process_whole_node (kf:349-381): windows inside one node, stored whatever only_store_nodes says
                    for off in range(offset, size - 1):
                        h = 0
                        for j in range(k):
                            h += g.base(node, off - k + 1 + j) * 4 ** j
                        out["kmers"].append(h)
                        out["nodes"].append(node)
                        out["start_nodes"].append(node)
                        out["start_offsets"].append(off)
                        out["af"].append(float(g.allele_frequencies[node]))
                    cur_hash = out["kmers"][-1]
                    first = size - 2 - (k - 1)          # window = the k bases ending at size-2
                    wbase = [g.base(node, first + j) for j in range(k)]
                    wnode = [node] * k
                    offset = size - 1
                # _get_first_base_in_path (kf:419-434)
                if nonempty >= k:
                    first_base = wbase[0]
                    if len(wbase) > 1:
                        while wbase[1] == -1:
                            wbase.pop(0)
                            wnode.pop(0)
                else:
                    first_base = 0
                base = g.base(node, offset)
                if nonempty >= k:
                    wbase.pop(0)
                    wnode.pop(0)
                    cur_hash = (cur_hash - first_base) // 4 + base * 4 ** (k - 1)        # update_hash, kf:31
                else:
                    cur_hash = cur_hash + 4 ** nonempty * base                          # kf:27
                wbase.append(base)
                wnode.append(node)
                nonempty += 1
                desc = (node, offset, frozenset(wnode))
                if (node != crit_node or offset != crit_offset) and desc in treated and len(wnode) >= k:
                    stopped = True
                    break
                treated.add(desc)
                if nonempty >= k:
                    add_kmer(cur_hash, node, offset, wnode)
                    if early_stop:
                        stopped = True
                        break
                if (node != crit_node or offset + 1 != crit_offset) and is_critical(node, offset + 1):
                    if (node, offset + 1) not in starting_set:
                        starting_points.append((node, offset + 1))
                        starting_set.add((node, offset + 1))
                    stopped = True
                    break
                offset += 1
            # ---------------- _search_next_nodes (kf:383-417) ----------------
            children = []
            if not stopped:
                children = g.out(node)
                force_follow = False
                if only_follow_nodes is not None and len(only_follow_nodes.intersection(children)) > 0:                   # kf:385-388
                    children = list(only_follow_nodes.intersection(children))      # a set in the reference: its iteration order
                    force_follow = True
                if children:
                    n_var = len(set(n for n in wnode if not g.is_linear[n]))
                    if not force_follow and n_var >= max_variant_nodes:
                        children = [n for n in children if g.is_linear[n]]
                        assert len(children) == 1, "Not 1 linear ref next nodes from node %d: %s" % (node, children)
            if children:
                stack.append([children, 0, cur_hash, list(wbase), list(wnode), nonempty])
                node, offset = children[0], 0
                continue
            # return to the closest frame that still has an unexplored child
            while stack:
                frame = stack[-1]
                frame[1] += 1
                if frame[1] < len(frame[0]):
                    node, offset, cur_hash = frame[0][frame[1]], 0, frame[2]
                    wbase, wnode, nonempty = list(frame[3]), list(frame[4]), frame[5]
                    break
                stack.pop()
            else:
                break
    return dict(kmers=np.array(out["kmers"], dtype=np.int64), nodes=np.array(out["nodes"], dtype=np.int32),
                start_nodes=np.array(out["start_nodes"], dtype=np.int32),
                start_offsets=np.array(out["start_offsets"]).astype(np.int16),
                allele_frequencies=np.array(out["af"], dtype=np.float64))
