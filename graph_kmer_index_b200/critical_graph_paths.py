"""Import path of graph_kmer_index/critical_graph_paths.py:5-104; the class is built next to the finder that consumes it
(kmer_finder.py, device pass gki_critical_paths in csrc/finder.cu)."""
from .kmer_finder import CriticalGraphPaths  # noqa: F401
