"""Read: an ordered list of gene calls and its k-length windows.  Mirror of upstream
amira/construct_read.py; the CUDA path enumerates the same windows from the CSR layout."""
from __future__ import annotations

from .construct_gene import Gene
from .construct_gene_mer import GeneMer


def convert_genes(annotatedGenes):
    return [Gene(g) for g in annotatedGenes]


class Read:
    def __init__(self, readId: str, annotatedGenes, annotatedGenePositions=None):
        self.readId = readId
        self.numberOfGenes = len(annotatedGenes)
        self.listOfGenes = convert_genes(annotatedGenes)
        self._annotatedGenes = annotatedGenes
        self._annotatedGenePositions = annotatedGenePositions

    def get_readId(self) -> str:
        return self.readId

    def get_genes(self) -> list:
        return self.listOfGenes

    def get_number_of_genes(self) -> int:
        return self.numberOfGenes

    def get_annotatedGenes(self) -> list:
        return self._annotatedGenes

    def get_annotatedGenePositions(self) -> list:
        return self._annotatedGenePositions

    def get_geneMers(self, kmerSize: int):
        """all L-k+1 windows in order, and (first gene start, last gene end) per window or None"""
        n_windows = self.numberOfGenes - (kmerSize - 1) if self.numberOfGenes > kmerSize - 1 else 0
        genes, pos = self.listOfGenes, self._annotatedGenePositions
        geneMers = [GeneMer(genes[i:i + kmerSize]) for i in range(n_windows)]
        if pos:
            spans = [(pos[i][0], pos[i + kmerSize - 1][1]) for i in range(n_windows)]
        else:
            spans = [None] * n_windows
        return geneMers, spans
