"""Gene-call encoding: ``{read_id: ['+geneA', '-geneB', ...]}`` -> signed int32 CSR for the CUDA path.

Replaces, for the whole read set at once, what upstream does one object at a time in
``Gene.__init__`` (amira/construct_gene.py:48-67) and ``convert_genes`` (amira/construct_read.py:5-8).

Gene id = strand * rank, rank in 1..V = position of the gene name in the vocabulary sorted by
``int(sha256(pickle.dumps(name)).hexdigest(), 16)`` ascending.  Upstream orders the genes of a
gene-mer by that integer times the strand (construct_gene.py:91-93, construct_gene_mer.py:15-39);
rank is a monotone map of it, so signed-rank order reproduces the canonical choice exactly.
"""
from __future__ import annotations

import ctypes as C
from itertools import chain

import numpy as np

from . import _lib
from .construct_gene import name_hash


class Vocabulary:
    """gene names in SHA-rank order"""

    def __init__(self, names):
        self.names = sorted(names, key=name_hash)
        self._blob = None

    def __len__(self):
        return len(self.names)

    def blob(self):
        """(utf-8 bytes of all names, int64 byte offsets) for amira_vocab_encode"""
        if self._blob is None:
            enc = [n.encode("utf-8") for n in self.names]
            off = np.zeros(len(enc) + 1, np.int64)
            np.cumsum([len(e) for e in enc], out=off[1:])
            self._blob = (b"".join(enc), off)
        return self._blob

    def signed_hashes(self):
        """object array H with H[id + V] = signed SHA int of gene id (index V unused)"""
        V = len(self.names)
        H = np.empty(2 * V + 1, object)
        for r, n in enumerate(self.names, 1):
            h = name_hash(n)
            H[V + r], H[V - r] = h, -h
        return H


def collect_names(reads: dict) -> set:
    """unique gene names of a read dict (strand stripped, spaces -> '_')"""
    toks = set(chain.from_iterable(reads.values()))
    names = set()
    for t in toks:
        if not isinstance(t, str) or t.replace(" ", "") == "" or t[0] not in "+-" or len(t) < 2:
            continue                      # the encoder reports these with upstream's assertion message
        names.add(t[1:].replace(" ", "_"))
    return names


def encode_reads(reads: dict, vocab: Vocabulary, positions: dict | None = None):
    """-> ids int32[G], off int64[R+1], pos_start, pos_end (or None, None)

    Raises AssertionError with upstream's messages for blank tokens, bad strand characters and
    empty names (through the C ABI's status codes), KeyError if a read has no positions entry."""
    lens = np.fromiter((len(v) for v in reads.values()), np.int64, len(reads))
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    G = int(off[-1])
    toks = list(chain.from_iterable(reads.values()))
    try:
        blob = "".join(toks).encode("utf-8")
    except TypeError:
        _bad_token(next(t for t in toks if not isinstance(t, str)))
    tok_off = np.zeros(G + 1, np.int64)
    np.cumsum(np.fromiter(map(len, toks), np.int64, G), out=tok_off[1:])
    if len(blob) != int(tok_off[-1]):          # non-ASCII gene names: byte lengths differ from str lengths
        enc = [t.encode("utf-8") for t in toks]
        np.cumsum(np.fromiter(map(len, enc), np.int64, G), out=tok_off[1:])
        blob = b"".join(enc)
    vblob, voff = vocab.blob()
    ids = np.empty(G, np.int32)
    bad = C.c_int64(-1)
    lib = _lib.load()
    status = lib.amira_vocab_encode(blob, tok_off.ctypes.data_as(C.c_void_p), G, vblob,
                                    voff.ctypes.data_as(C.c_void_p), len(vocab),
                                    ids.ctypes.data_as(C.c_void_p), C.byref(bad))
    _lib.check(status)
    ps = pe = None
    if positions:
        ps = np.empty(G, np.int32)
        pe = np.empty(G, np.int32)
        i = 0
        for rid, calls in reads.items():
            p = positions[rid]
            n = len(calls)
            if n:
                if not p:
                    ps[i:i + n] = -1
                    pe[i:i + n] = -1
                else:
                    arr = np.asarray(p[:n], np.int64).reshape(-1, 2)
                    ps[i:i + n] = arr[:, 0]
                    pe[i:i + n] = arr[:, 1]
            i += n
    return ids, off, ps, pe


def _bad_token(t):
    raise AttributeError("gene calls must be strings, got %r" % (t,))
