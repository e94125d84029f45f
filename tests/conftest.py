import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_index(g, prefix="stable_"):
    """Golden npz -> oracle-style index dict (attribute names of CollisionFreeKmerIndex)."""
    return dict(_hashes_to_index=g[prefix + "hashes_to_index"], _n_kmers=g[prefix + "n_kmers"],
                _nodes=g[prefix + "nodes"], _ref_offsets=g[prefix + "ref_offsets"], _kmers=g[prefix + "kmers"],
                _modulo=int(g[prefix + "modulo"]), _frequencies=g[prefix + "frequencies"],
                _allele_frequencies=g[prefix + "allele_frequencies"])


@pytest.fixture(scope="session")
def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
