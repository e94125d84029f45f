"""Import path of graph_kmer_index/nplist.py: an append-only list over a numpy buffer.  The device finder returns whole arrays and
does not need it; it is kept for callers that fill or read ``DenseKmerFinder``-style result lists through this interface
(append / extend / get_nparray / set_n_elements / copy / len / indexing)."""
import numpy as np


class NpList:
    def __init__(self, dtype=None):
        self._dtype = dtype
        self._data = np.empty(0, dtype=dtype) if dtype is not None else np.empty(0)
        self._n_elements = 0

    def _reserve(self, wanted):
        if wanted <= len(self._data):
            return
        grown = np.zeros(max(wanted, 100, len(self._data) * 3 // 2), dtype=self._data.dtype)
        grown[:self._n_elements] = self._data[:self._n_elements]
        self._data = grown

    def _adopt_dtype(self, sample):
        if self._dtype is None:
            self._dtype = np.asarray(sample).dtype if not isinstance(sample, (int, float, bool)) else type(sample)
            self._data = self._data.astype(self._dtype)

    def append(self, element):
        self._adopt_dtype(element)
        self._reserve(self._n_elements + 1)
        self._data[self._n_elements] = element
        self._n_elements += 1

    def extend(self, elements):
        elements = np.asarray(elements)
        if len(elements) == 0:
            return
        self._adopt_dtype(elements[0])
        self._reserve(self._n_elements + len(elements))
        self._data[self._n_elements:self._n_elements + len(elements)] = elements
        self._n_elements += len(elements)

    def get_nparray(self):
        return self._data[:self._n_elements]

    def set_n_elements(self, n):
        self._n_elements = n

    def copy(self):
        other = NpList(dtype=self._dtype)
        other.extend(self.get_nparray())
        return other

    def __getitem__(self, item):
        return self.get_nparray()[item]

    def __len__(self):
        return self._n_elements

    def __eq__(self, other):
        return bool(np.all(self.get_nparray() == other.get_nparray()))

    def __repr__(self):
        return "NpList(%s)" % self.get_nparray()

    __str__ = lambda self: str(self.get_nparray())  # noqa: E731
