"""Drop-in for graph_kmer_index/structural_variants.py:6-43, the caller of ``bionumpy_hash`` (row a10): for every
variant whose node is longer than k + 5, k-mers inside the node that are rare in ``kmer_index_with_frequencies``,
taken greedily so that no two chosen k-mers overlap."""
import numpy as np

from .bionumpy_wrapper import bionumpy_hash
from .flat_kmers import FlatKmers


def sample_kmers_from_structural_variants(graph, variant_to_nodes, kmer_index_with_frequencies, k, max_frequency=2):
    kmers, nodes = [np.zeros(0, dtype=np.uint64)], [np.zeros(0, dtype=np.uint32)]
    for pair in variant_to_nodes:
        for node in pair:
            if graph.get_node_size(node) <= k + 5:
                continue
            node_kmers = bionumpy_hash(graph.get_numeric_node_sequence(node), k)
            frequencies = np.array([kmer_index_with_frequencies.get_frequency(x) for x in node_kmers])
            chosen, free_from = [], -1
            for position in np.flatnonzero(frequencies < max_frequency):          # leftmost first, k apart (structural_variants.py:24-30)
                if position >= free_from:
                    chosen.append(position)
                    free_from = position + k
            kmers.append(node_kmers[chosen])
            nodes.append(np.full(len(chosen), node, dtype=np.uint32))
    kmers = np.concatenate(kmers)
    return FlatKmers(kmers.astype(np.uint64), np.concatenate(nodes), np.zeros(len(kmers), dtype=np.uint32))
