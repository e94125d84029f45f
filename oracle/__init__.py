"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's (ivargr/graph_kmer_index) algorithm for the
read-k-mer counting hot path: hashing, CollisionFreeKmerIndex construction,
bucket probe, per-node counting.  It exists to *check* the CUDA product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here -- and only as the checker,
never as the thing shipped.  Nothing under ``graph_kmer_index_b200/`` imports it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here
against (i) the known-answer vectors of the reference's own tests
(``tests/test_kmer_hashing.py``, ``tests/test_collision_free_kmer_index.py``) and
(ii) fixtures under ``tests/golden/`` produced by running the unmodified reference
(imported from /root/reference with the shims in ``oracle/ref_shims.py``) through
``tests/golden/make_golden.py``.  The counting step (``CounterKmerIndex.count_kmers`` /
``get_node_counts``) delegates to the third-party ``npstructures.Counter`` which is not
vendored in the reference; for that step the oracle restates the published semantics
(increment the counter of every query key that is present, ignore absent keys) and is
anchored on the reference's in-repo probe loops (``CollisionFreeKmerIndex.get``,
``cython_kmer_index.pyx``), see ``numpy_oracle.node_counts``.
"""
