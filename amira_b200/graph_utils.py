"""Entry points every upstream caller uses to build a graph (amira/graph_utils.py:12-14, 105-124).

Upstream's ``build_multiprocessed_graph`` splits the reads into ``cores`` strided batches, builds
sub-graphs in joblib workers and merges them; every call site passes ``cores=1`` (the merge of more
than one sub-graph mis-counts edge coverage, graph_utils.py:75) so the single-build result is the
contract.  Here ``cores`` is accepted and ignored: the build is one GPU pass over all reads."""
from __future__ import annotations

import numpy as np

from .construct_graph import GeneMerGraph


def build_graph(read_dict, kmer_size, gene_positions=None):
    return GeneMerGraph(read_dict, kmer_size, gene_positions)


def build_multiprocessed_graph(annotatedReads, geneMer_size, cores=1, gene_positions=None):
    graph = GeneMerGraph(annotatedReads, geneMer_size, gene_positions)
    return graph


def get_overall_mean_node_coverages(graph):
    """upstream amira/graph_utils.py:299-313: for k = 3, 5, ..., 15 the mean over nodes of the number of reads on
    the node that have at least k gene calls.  Called right after the first build of every run.

    Upstream walks every (node, read) incidence seven times in Python; a graph that still matches its device
    build answers from the exported incidence array in one numpy pass per k (same values, same types:
    statistics.mean of ints is an int when exact and the correctly rounded quotient otherwise)."""
    inc = getattr(graph, "_incidence_arrays", None)
    if inc is not None and getattr(graph, "_device_synced", False) and not getattr(graph, "_device_ops", ()):
        node_reads, read_len, n_nodes = inc
        out = {}
        for k in range(3, 16, 2):
            if n_nodes == 0:
                out[k] = 0
                continue
            total = int(np.count_nonzero(read_len[node_reads] >= k))
            out[k] = total // n_nodes if total % n_nodes == 0 else total / n_nodes
        return out
    import statistics
    reads = graph.get_reads()
    out = {}
    for k in range(3, 16, 2):
        long_enough = {r for r, calls in reads.items() if len(calls) >= k}
        coverages = [sum(1 for r in node.get_reads() if r in long_enough) for node in graph.all_nodes()]
        out[k] = statistics.mean(coverages) if coverages else 0
    return out
