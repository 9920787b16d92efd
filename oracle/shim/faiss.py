"""ORACLE (test infrastructure only): stand-in for `faiss` so the Python reference can be imported
in the dev container (`PYTHONPATH=oracle/shim:/root/reference/src`).  faiss-cpu is not installed
in this image and there is no network.

Implements exactly the protocol the reference uses:
    IndexFlatL2(d), .add(x[n,d]), .search(x[nq,d], k) -> (D fp32 asc, I int64, -1 padded),
    .remove_ids(ids) (compacting: later rows shift down), .ntotal
Distance arithmetic = oracle.knn_oracle (restated fvec_L2sqr lane order).  Never imported by the
product package.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.knn_oracle import knn_l2_numpy  # noqa: E402

__version__ = "0.0-oracle-shim"


class IndexFlatL2:
    def __init__(self, d: int):
        self.d = int(d)
        self._x = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return int(self._x.shape[0])

    def add(self, x):
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32).reshape(-1, self.d)
        self._x = np.concatenate([self._x, x], axis=0)

    def remove_ids(self, ids):
        ids = np.asarray(ids).reshape(-1).astype(np.int64)
        keep = np.ones(self.ntotal, dtype=bool)
        ids = ids[(ids >= 0) & (ids < self.ntotal)]
        keep[ids] = False
        removed = int((~keep).sum())
        self._x = self._x[keep]
        return removed

    def search(self, x, k):
        x = np.ascontiguousarray(np.asarray(x), dtype=np.float32).reshape(-1, self.d)
        return knn_l2_numpy(x, self._x, int(k))

    def reset(self):
        self._x = np.zeros((0, self.d), dtype=np.float32)
