"""ORACLE (test infrastructure only): Python face of oracle/knn_oracle.c plus a numpy twin.

Follows /root/reference/src/adaptive_classifier/memory.py:85-136 and the restated
faiss.IndexFlatL2 semantics documented in knn_oracle.c.  PARITY UNPINNED against real FAISS
(absent offline); pinned against float64 ground truth and the reference's tests/test_memory.py.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = _build.SO if os.path.exists(_build.SO) else _build.build()
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.oracle_l2sqr.restype = ctypes.c_float
        L.oracle_l2sqr.argtypes = [fp, fp, ctypes.c_int]
        L.oracle_l2sqr_ny.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, fp]
        L.oracle_knn_l2.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp, ip, ctypes.c_int64]
        L.oracle_proto_scores.argtypes = [fp, ip, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_topk_merge.argtypes = [fp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ip]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def knn_l2(Q: np.ndarray, P: np.ndarray, k: int, row_offset: int = 0):
    """IndexFlatL2.search restatement -> (D[nq,k] fp32 ascending, I[nq,k] int64)."""
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    P = np.ascontiguousarray(P, dtype=np.float32)
    nq, d = Q.shape
    n = P.shape[0]
    out_d = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    lib().oracle_knn_l2(_f(Q), _f(P), nq, n, d, k, _f(out_d), _i(out_i), row_offset)
    return out_d, out_i


def all_dist(q: np.ndarray, P: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32)
    P = np.ascontiguousarray(P, dtype=np.float32)
    out = np.empty((P.shape[0],), dtype=np.float32)
    lib().oracle_l2sqr_ny(_f(q), _f(P), P.shape[0], P.shape[1], _f(out))
    return out


def proto_scores(d: np.ndarray, idx: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(d, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.zeros_like(d)
    lib().oracle_proto_scores(_f(d), _i(idx), d.shape[0], d.shape[1], _f(out))
    return out


def topk_merge(d: np.ndarray, idx: np.ndarray):
    d = np.ascontiguousarray(d, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    G, nq, k = d.shape
    od = np.empty((nq, k), dtype=np.float32)
    oi = np.empty((nq, k), dtype=np.int64)
    lib().oracle_topk_merge(_f(d), _i(idx), G, nq, k, _f(od), _i(oi))
    return od, oi


# ---- numpy twin (same lane order; used by oracle/shim/faiss.py so the shim needs no compiler) ----

def all_dist_numpy(Q: np.ndarray, P: np.ndarray) -> np.ndarray:
    """[nq, n] fp32 distances with exactly the 8-lane order of oracle_l2sqr."""
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    P = np.ascontiguousarray(P, dtype=np.float32)
    nq, d = Q.shape
    n = P.shape[0]
    dp = (d + 7) // 8 * 8
    out = np.empty((nq, n), dtype=np.float32)
    Pp = np.zeros((n, dp), dtype=np.float32)
    Pp[:, :d] = P
    for b in range(nq):
        qp = np.zeros((dp,), dtype=np.float32)
        qp[:d] = Q[b]
        t = (qp[None, :] - Pp)            # fp32
        sq = (t * t).reshape(n, dp // 8, 8)
        lanes = np.zeros((n, 8), dtype=np.float32)
        for c in range(dp // 8):
            lanes = lanes + sq[:, c, :]
        a = lanes[:, 0] + lanes[:, 4]
        bb = lanes[:, 1] + lanes[:, 5]
        c2 = lanes[:, 2] + lanes[:, 6]
        e = lanes[:, 3] + lanes[:, 7]
        out[b] = (a + bb) + (c2 + e)
    return out


def knn_l2_numpy(Q, P, k, row_offset=0):
    D = all_dist_numpy(Q, P)
    nq, n = D.shape
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    ids = np.arange(n, dtype=np.int64)
    for b in range(nq):
        order = np.lexsort((ids, D[b]))[:k]       # primary key d, ties -> lower id
        out_d[b, : len(order)] = D[b, order]
        out_i[b, : len(order)] = order + row_offset
    return out_d, out_i
