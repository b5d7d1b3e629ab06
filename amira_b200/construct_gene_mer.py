"""GeneMer: k consecutive gene calls, stored as its canonical orientation plus the reverse
complement.  Host-side mirror of upstream amira/construct_gene_mer.py.

Canonical = the smaller of (signed gene hashes of the window) and (signed gene hashes of its reverse
complement) under Python list order (construct_gene_mer.py:15-39); direction = +1 when the window as
read is the canonical one.  The CUDA path makes the same choice on SHA-rank integers."""
from __future__ import annotations

from .construct_gene import Gene, hashlib_hash


def define_rc_geneMer(geneMer):
    assert all(isinstance(g, Gene) for g in geneMer)
    return [g.reverse_gene() for g in geneMer[::-1]]


def sort_geneMers(geneMer, rcGeneMer):
    fwd = [g.__hash__() for g in geneMer]
    rev = [g.__hash__() for g in rcGeneMer]
    assert fwd != rev, "Gene-mer and reverse complement gene-mer are identical"
    return fwd, rev, sorted((fwd, rev))


def choose_canonical_geneMer(geneMer, geneMerHashes, rcGeneMer, rcGeneMerHashes, sortedGeneMerhashes):
    if sortedGeneMerhashes[0] == geneMerHashes:
        return geneMer, rcGeneMer
    return rcGeneMer, geneMer


def define_geneMer(geneMer):
    assert isinstance(geneMer, list), "Gene-mer is not a list of Gene objects"
    assert geneMer != [], "Gene-mer is empty"
    rc = define_rc_geneMer(geneMer)
    fwd_h, rc_h, ordered = sort_geneMers(geneMer, rc)
    return choose_canonical_geneMer(geneMer, fwd_h, rc, rc_h, ordered)


class GeneMer:
    __slots__ = ("canonicalGeneMer", "rcGeneMer", "geneMerSize", "geneMerDirection", "_hash")

    def __init__(self, geneMer: list):
        self.canonicalGeneMer, self.rcGeneMer = define_geneMer(geneMer)
        self.geneMerSize = len(self.canonicalGeneMer)
        self.geneMerDirection = 1 if self.canonicalGeneMer == geneMer else -1
        self._hash = None

    @classmethod
    def from_canonical(cls, canonical: list, rc: list, direction: int, node_hash: int | None = None) -> "GeneMer":
        """assemble a GeneMer whose orientation was already decided (by the CUDA path)"""
        gm = object.__new__(cls)
        gm.canonicalGeneMer, gm.rcGeneMer = canonical, rc
        gm.geneMerSize = len(canonical)
        gm.geneMerDirection = direction
        gm._hash = node_hash
        return gm

    def get_canonical_geneMer(self):
        return self.canonicalGeneMer

    def get_rc_geneMer(self):
        return self.rcGeneMer

    def get_geneMerDirection(self):
        return self.geneMerDirection

    def get_geneMer_size(self) -> int:
        return self.geneMerSize

    def __eq__(self, otherGeneMer):
        return (self.canonicalGeneMer == otherGeneMer.get_canonical_geneMer()
                and self.rcGeneMer == otherGeneMer.get_rc_geneMer())

    def __hash__(self):
        """SHA-256 of the tuple of canonical signed gene hashes -- the node key (construct_gene_mer.py:94-97)"""
        if self._hash is None:
            self._hash = hashlib_hash(tuple(g.__hash__() for g in self.canonicalGeneMer))
        return self._hash
