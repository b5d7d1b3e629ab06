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
    build answers with one pass of k_read_length_coverages over the node -> reads CSR on the device (same values, same
    types: statistics.mean of ints is an int when exact and the correctly rounded quotient otherwise).  Upstream reads
    len(graph.get_reads()[r]) at call time: the device path is only taken while the read lengths are those of the
    build."""
    ks = list(range(3, 16, 2))
    on_device = getattr(graph, "_on_device", None)
    if on_device is not None and on_device() and _read_lengths_unchanged(graph):
        h = graph._require_device_state()
        n_nodes = graph.get_total_number_of_nodes()
        if n_nodes == 0:
            return {k: 0 for k in ks}
        sums = h.read_length_coverages(ks).tolist()
        return {k: (t // n_nodes if t % n_nodes == 0 else t / n_nodes) for k, t in zip(ks, sums)}
    import statistics
    reads = graph.get_reads()
    out = {}
    for k in ks:
        long_enough = {r for r, calls in reads.items() if len(calls) >= k}
        coverages = [sum(1 for r in node.get_reads() if r in long_enough) for node in graph.all_nodes()]
        out[k] = statistics.mean(coverages) if coverages else 0
    return out


def _read_lengths_unchanged(graph) -> bool:
    reads = graph.get_reads()
    lens = getattr(graph, "_read_len", None)
    if lens is None or len(lens) != len(reads):
        return False
    return bool(np.array_equal(np.fromiter((len(v) for v in reads.values()), np.int64, len(reads)), lens))


def build_k_sweep(reads, ks, gene_positions=None, device=None):
    """The k sweep of choose_kmer_size (amira/graph_utils.py:258-296: k = 3, 5, ..., 15 over the SAME reads) in one go:
    the reads are parsed once (EncodedReads), the CSR is uploaded once and stays resident, and one build per k is
    enqueued back to back; nothing is copied to the host until a graph is looked at.  Returns {k: GeneMerGraph}."""
    from .encode import EncodedReads
    enc = reads if isinstance(reads, EncodedReads) else EncodedReads(reads, gene_positions)
    return {k: GeneMerGraph(enc, k, device=device) for k in ks}
