"""Edge of the gene-mer graph.  Mirror of upstream amira/construct_edge.py.

An edge's key is min(SHA((s*sd, t*td)), SHA((-s*sd, -t*td))) over the node keys s, t and the
traversal directions sd, td (construct_edge.py:104-124): it identifies (source, target, sd*td), which
is exactly the key of the CUDA edge table."""
from __future__ import annotations

from .construct_gene import hashlib_hash


def extract_node_hashes(firstNode, secondNode):
    return firstNode.__hash__(), secondNode.__hash__()


def sort_node_hashes(firstNodeHash, secondNodeHash):
    lo, hi = sorted((firstNodeHash, secondNodeHash))
    return lo, hi


def define_source_and_target(firstNode, secondNode):
    return sort_node_hashes(*extract_node_hashes(firstNode, secondNode))


def edge_key(source_hash: int, target_hash: int, sd: int, td: int) -> int:
    a, b = source_hash * sd, target_hash * td
    return min(hashlib_hash((a, b)), hashlib_hash((-a, -b)))


class Edge:
    def __init__(self, sourceNode, targetNode, sourceNodeDirection, targetNodeDirection):
        self.sourceNode = sourceNode
        self.targetNode = targetNode
        self.edgeCoverage = 0
        self.sourceNodeDirection = sourceNodeDirection
        self.targetNodeDirection = targetNodeDirection

    def get_sourceNode(self):
        return self.sourceNode

    def get_targetNode(self):
        return self.targetNode

    def set_sourceNode(self, new_sourceNode):
        self.sourceNode = new_sourceNode
        return self.sourceNode

    def set_targetNode(self, new_targetNode):
        self.targetNode = new_targetNode
        return self.targetNode

    def set_sourceNodeDirection(self, sourceDirection) -> int:
        self.sourceNodeDirection = sourceDirection
        return self.sourceNodeDirection

    def get_sourceNodeDirection(self) -> int:
        return self.sourceNodeDirection

    def set_targetNodeDirection(self, targetDirection) -> int:
        self.targetNodeDirection = targetDirection
        return self.targetNodeDirection

    def get_targetNodeDirection(self) -> int:
        return self.targetNodeDirection

    def get_edge_coverage(self) -> int:
        return self.edgeCoverage

    def increment_edge_coverage(self) -> int:
        self.edgeCoverage += 1
        return self.edgeCoverage

    def extend_edge_coverage(self, value) -> int:
        self.edgeCoverage += value
        return self.edgeCoverage

    def reduce_edge_coverage(self):
        self.edgeCoverage -= 1
        return self.edgeCoverage

    def __eq__(self, otherEdge) -> bool:
        mine = sorted((self.sourceNode.__hash__(), self.targetNode.__hash__()))
        theirs = sorted((otherEdge.get_sourceNode().__hash__(), otherEdge.get_targetNode().__hash__()))
        return mine == theirs

    def __hash__(self):
        return edge_key(self.sourceNode.__hash__(), self.targetNode.__hash__(), self.sourceNodeDirection,
                        self.targetNodeDirection)
