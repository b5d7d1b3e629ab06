"""Entry points every upstream caller uses to build a graph (amira/graph_utils.py:12-14, 105-124).

Upstream's ``build_multiprocessed_graph`` splits the reads into ``cores`` strided batches, builds
sub-graphs in joblib workers and merges them; every call site passes ``cores=1`` (the merge of more
than one sub-graph mis-counts edge coverage, graph_utils.py:75) so the single-build result is the
contract.  Here ``cores`` is accepted and ignored: the build is one GPU pass over all reads."""
from __future__ import annotations

from .construct_graph import GeneMerGraph


def build_graph(read_dict, kmer_size, gene_positions=None):
    return GeneMerGraph(read_dict, kmer_size, gene_positions)


def build_multiprocessed_graph(annotatedReads, geneMer_size, cores=1, gene_positions=None):
    graph = GeneMerGraph(annotatedReads, geneMer_size, gene_positions)
    return graph
