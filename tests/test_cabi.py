"""CPU-side checks of the boundary: libgki.so builds/loads, exports every symbol include/gki.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from graph_kmer_index_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gki.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gki_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.gki_version() >= 100


def test_ctypes_table_matches_header():
    names = set(declared_symbols())
    bound = {n for n, sig in _lib._SIGNATURES.items() if sig is not None} | {"gki_last_error", "gki_version", "gki_launch_count"}
    assert names <= bound | {n for n, sig in _lib._SIGNATURES.items()}, names - bound


def test_no_cpu_fallback(have_cuda):
    if have_cuda:
        pytest.skip("CUDA present: the loud failure is only observable without a device")
    from graph_kmer_index_b200 import kmer_hashing
    with pytest.raises(_lib.GkiError):
        kmer_hashing.kmer_hashes_to_reverse_complement_hash(np.arange(4, dtype=np.uint64), 5)
    with pytest.raises(_lib.GkiError):
        kmer_hashing.sequence_to_kmer_hash("ACGT")


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: no product source imports, loads or executes anything under oracle/"""
    import re
    pkg = os.path.join(ROOT, "graph_kmer_index_b200")
    forbidden = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|oracle[./](_ref|gki_oracle|c_oracle|numpy_oracle|finder_oracle)|liboracle", re.M)
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not forbidden.search(src), f
                checked += 1
    assert checked > 15
