"""Upstream's SHA-256 dictionary keys for whole arrays of nodes / edges at once.

``hashlib_hash(x) = int(sha256(pickle.dumps(x)).hexdigest(), 16)`` (amira/construct_gene.py:5-10) per node
and (twice) per edge is a third of the time the drop-in class spends turning exported arrays into
upstream-shaped dictionaries.  ``libamira_gmg.so`` restates pickle's protocol-4 byte stream for tuples of
ints and SHA-256 in C (csrc/host_keys.cpp); this module feeds it numpy arrays and -- because the values
must equal what upstream's interpreter would compute -- checks a few results of every call against
hashlib/pickle, falling back to the plain Python formula if they ever differ (another pickle default
protocol, for instance)."""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np

from . import _lib
from .construct_edge import edge_key
from .construct_gene import hashlib_hash

_enabled = pickle.DEFAULT_PROTOCOL == 4


def _ints(digests: np.ndarray) -> list:
    raw = digests.tobytes()
    return [int.from_bytes(raw[i:i + 32], "big") for i in range(0, len(raw), 32)]


def _spot(n: int):
    return sorted({0, n // 2, n - 1}) if n else []


def node_keys(key: np.ndarray, vocab):
    """key: (n, k) signed SHA ranks of the canonical gene-mers -> ([node key ints], (n, 32) uint8 digests)"""
    global _enabled
    n, k = key.shape
    H = vocab.signed_hashes()
    V = len(vocab)
    if _enabled and n:
        mags = np.ascontiguousarray(vocab.sha_bytes()[np.abs(key) - 1])           # (n, k, 32)
        neg = np.ascontiguousarray((key < 0).astype(np.int8))
        out = np.empty((n, 32), np.uint8)
        rc = _lib.load().amira_host_tuple_sha(mags.ctypes.data_as(C.c_void_p), neg.ctypes.data_as(C.c_void_p), n, k,
                                              out.ctypes.data_as(C.c_void_p))
        keys = _ints(out) if rc == 0 else None
        if keys is not None and all(keys[i] == hashlib_hash(tuple(H[key[i] + V].tolist())) for i in _spot(n)):
            return keys, out
        _enabled = False
    keys = [hashlib_hash(tuple(row)) for row in H[key + V].tolist()] if n else []
    out = np.frombuffer(b"".join(x.to_bytes(32, "big") for x in keys), np.uint8).reshape(n, 32) if n else np.zeros((0, 32), np.uint8)
    return keys, out


def edge_keys(node_keys_: list, node_sha: np.ndarray, src: np.ndarray, tgt: np.ndarray, sd: np.ndarray, td: np.ndarray) -> list:
    """keys of the exported edges (construct_edge.py:104-124), in edge order"""
    global _enabled
    m = len(src)
    slow = lambda j: edge_key(node_keys_[src[j]], node_keys_[tgt[j]], int(sd[j]), int(td[j]))
    if _enabled and m:
        a = [np.ascontiguousarray(x) for x in (node_sha, src.astype(np.int32), tgt.astype(np.int32), sd.astype(np.int8),
                                               td.astype(np.int8))]
        out = np.empty((m, 32), np.uint8)
        rc = _lib.load().amira_host_edge_keys(*[x.ctypes.data_as(C.c_void_p) for x in a], m, out.ctypes.data_as(C.c_void_p))
        keys = _ints(out) if rc == 0 else None
        if keys is not None and all(keys[j] == slow(j) for j in _spot(m)):
            return keys
        _enabled = False
    return [slow(j) for j in range(m)]
