"""Edge of the gene-mer graph.  Mirror of upstream amira/construct_edge.py.

An edge's key is min(SHA((s*sd, t*td)), SHA((-s*sd, -t*td))) over the node keys s, t and the
traversal directions sd, td (construct_edge.py:104-124): it identifies (source, target, sd*td), which
is exactly the key of the CUDA edge table."""
from __future__ import annotations

from ._surface import expose
from .construct_gene import hashlib_hash


def edge_key(source_hash: int, target_hash: int, sd: int, td: int) -> int:
    a, b = source_hash * sd, target_hash * td
    return min(hashlib_hash((a, b)), hashlib_hash((-a, -b)))


def extract_node_hashes(firstNode, secondNode):  # the full SHA-256 integers: hash() would truncate them to 64 bits
    return firstNode.__hash__(), secondNode.__hash__()


def sort_node_hashes(firstNodeHash, secondNodeHash):
    return tuple(sorted((firstNodeHash, secondNodeHash)))


def define_source_and_target(firstNode, secondNode):
    return sort_node_hashes(*extract_node_hashes(firstNode, secondNode))


_FIELDS = ("sourceNode", "targetNode", "sourceNodeDirection", "targetNodeDirection")


@expose(getters=[("get_" + f, f) for f in _FIELDS] + [("get_edge_coverage", "edgeCoverage")],
        setters=[("set_" + f, f) for f in _FIELDS],
        steppers=[("increment_edge_coverage", "edgeCoverage", 1), ("reduce_edge_coverage", "edgeCoverage", -1),
                  ("extend_edge_coverage", "edgeCoverage", None)])
class Edge:
    def __init__(self, sourceNode, targetNode, sourceNodeDirection, targetNodeDirection):
        self.sourceNode, self.targetNode = sourceNode, targetNode
        self.sourceNodeDirection, self.targetNodeDirection = sourceNodeDirection, targetNodeDirection
        self.edgeCoverage = 0

    def _end_hashes(self):
        return sorted((self.sourceNode.__hash__(), self.targetNode.__hash__()))

    def __eq__(self, otherEdge) -> bool:
        return self._end_hashes() == sorted((otherEdge.get_sourceNode().__hash__(), otherEdge.get_targetNode().__hash__()))

    def __hash__(self):
        return edge_key(self.sourceNode.__hash__(), self.targetNode.__hash__(), self.sourceNodeDirection,
                        self.targetNodeDirection)
