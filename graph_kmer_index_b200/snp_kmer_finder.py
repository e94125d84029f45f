"""Import path of graph_kmer_index/snp_kmer_finder.py for the scalar hash helpers on the hot path (snp_kmer_finder.py:14-26):
KAGE and the reference's own tests import ``kmer_to_hash_fast`` / ``sequence_to_kmer_hash`` / ``kmer_hash_to_sequence`` from here.
They live in kmer_hashing.py (one-read calls of K1).  The legacy ``SnpKmerFinder`` class of the same file is out of scope
(SURVEY.md section 2)."""
from .kmer_hashing import kmer_hash_to_sequence, kmer_to_hash_fast, sequence_to_kmer_hash  # noqa: F401
