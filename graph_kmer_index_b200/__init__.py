"""graph_kmer_index_b200 -- B200 (sm_100a) implementation of graph_kmer_index's read-k-mer counting hot path.

Same public names as ``graph_kmer_index/__init__.py:1-12`` for the components on that path; everything
numeric runs in the in-tree CUDA library ``libgki.so`` (C ABI: include/gki.h).  Importing the package does not
touch CUDA; the first call that needs the library loads it and fails loudly if it is missing.
"""
from .kmer_hashing import (letter_sequence_to_numeric, numeric_to_letter_sequence, kmer_to_hash_fast,  # noqa: F401
                           sequence_to_kmer_hash, kmer_hash_to_sequence)
from .flat_kmers import FlatKmers, FlatKmers2  # noqa: F401
from .collision_free_kmer_index import CollisionFreeKmerIndex, CounterKmerIndex, MinimalKmerIndex, DeviceIndex, KmerIndex2  # noqa: F401
from .collision_free_kmer_index import CollisionFreeKmerIndex as KmerIndex  # noqa: F401
from .cython_kmer_index import CythonKmerIndex  # noqa: F401
from .read_kmers import ReadKmers  # noqa: F401
from .reverse_kmer_index import ReverseKmerIndex  # noqa: F401
from .reference_kmer_index import ReferenceKmerIndex  # noqa: F401

__version__ = "0.1.0"
from .kmer_finder import DenseKmerFinder, CriticalGraphPaths  # noqa: F401,E402
