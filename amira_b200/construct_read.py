"""Read: an ordered list of gene calls and its k-length windows.  Mirror of upstream
amira/construct_read.py; the CUDA path enumerates the same windows from the CSR layout."""
from __future__ import annotations

from ._surface import expose
from .construct_gene import Gene
from .construct_gene_mer import GeneMer


def convert_genes(annotatedGenes):
    return list(map(Gene, annotatedGenes))


@expose(getters=[("get_readId", "readId"), ("get_genes", "listOfGenes"), ("get_number_of_genes", "numberOfGenes"),
                 ("get_annotatedGenes", "_annotatedGenes"), ("get_annotatedGenePositions", "_annotatedGenePositions")])
class Read:
    def __init__(self, readId: str, annotatedGenes, annotatedGenePositions=None):
        self.readId = readId
        self._annotatedGenes, self._annotatedGenePositions = annotatedGenes, annotatedGenePositions
        self.listOfGenes = convert_genes(annotatedGenes)
        self.numberOfGenes = len(self.listOfGenes)

    def get_geneMers(self, kmerSize: int):
        """every window of kmerSize consecutive calls, in order (construct_read.py:37-59), and per window the
        (start of its first gene, end of its last gene) on the read, or None without positions"""
        starts = range(max(0, self.numberOfGenes - kmerSize + 1))
        windows = [GeneMer(self.listOfGenes[i:i + kmerSize]) for i in starts]
        where = self._annotatedGenePositions
        if not where:
            return windows, [None] * len(windows)
        return windows, [(where[i][0], where[i + kmerSize - 1][1]) for i in starts]
