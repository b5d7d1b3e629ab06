"""Gene: one signed gene call.  Host-side mirror of upstream amira/construct_gene.py (same names,
argument meaning and error behaviour) so code written against upstream runs unchanged.

The integer a Gene hashes to, ``int(sha256(pickle.dumps(name)).hexdigest(), 16) * strand``
(upstream construct_gene.py:5-10, 91-93), orders genes inside a gene-mer and therefore decides which
orientation is canonical; the CUDA path works on the rank of that integer instead (encode.py)."""
from __future__ import annotations

import hashlib
import pickle

_STRANDS = {"+": 1, "-": -1}
_SYMBOLS = {1: "+", -1: "-"}
_NAME_HASH: dict = {}


def hashlib_hash(value) -> int:
    """SHA-256 of the pickled value as a (256-bit) Python int -- construct_gene.py:5-10"""
    return int.from_bytes(hashlib.sha256(pickle.dumps(value)).digest(), "big")


def name_hash(name: str) -> int:
    """hashlib_hash(name), memoised per gene name (upstream recomputes it on every __hash__ call)"""
    h = _NAME_HASH.get(name)
    if h is None:
        if len(_NAME_HASH) > 1 << 20:
            _NAME_HASH.clear()
        h = _NAME_HASH[name] = hashlib_hash(name)
    return h


def convert_string_strand_to_int(stringStrand: str) -> int:
    assert stringStrand in _STRANDS
    return _STRANDS[stringStrand]


def reverse_strand(geneStrand: int) -> int:
    assert geneStrand in _SYMBOLS
    return -geneStrand


def convert_int_strand_to_string(intStrand: int) -> str:
    assert intStrand in _SYMBOLS
    return _SYMBOLS[intStrand]


def split_call(gene: str):
    """'+name' -> (name, +1); the three assertion messages are upstream's (construct_gene.py:52-62)"""
    assert gene.replace(" ", "") != "", "Gene information is missing"
    strand, name = gene[0], gene[1:].replace(" ", "_")
    assert strand in _STRANDS, "Strand information missing for: " + gene
    assert name != "", "Gene name information missing for: " + gene
    return name, _STRANDS[strand]


class Gene:
    __slots__ = ("name", "strand")

    def __init__(self, gene: str):
        self.name, self.strand = split_call(gene)

    @classmethod
    def from_parts(cls, name: str, strand: int) -> "Gene":
        g = object.__new__(cls)
        g.name, g.strand = name, strand
        return g

    def get_name(self) -> str:
        return self.name

    def get_strand(self) -> int:
        return self.strand

    def reverse_gene(self) -> "Gene":
        return Gene.from_parts(self.name, reverse_strand(self.strand))

    def __eq__(self, otherGene) -> bool:
        return self.strand == otherGene.get_strand() and self.name == otherGene.get_name()

    def __hash__(self) -> int:
        return name_hash(self.name) * self.strand

    def __repr__(self) -> str:
        return "Gene(%s%s)" % (_SYMBOLS[self.strand], self.name)
