"""Node of the gene-mer graph.  Mirror of upstream amira/construct_node.py: downstream correction and
path-finding code reaches into these attributes, so their names and list semantics are kept."""
from __future__ import annotations

from ._surface import expose
from .construct_gene_mer import GeneMer


def _append_once(lst: list, item):
    if item not in lst:
        lst.append(item)


def _take_out(lst: list, item, message: str):
    assert item in lst, message
    lst.remove(item)


@expose(getters=[("get_geneMer", "geneMer"), ("get_canonical_geneMer", "canonicalGeneMer"),
                 ("get_reverse_geneMer", "reverseGeneMer"), ("get_node_coverage", "nodeCoverage"),
                 ("get_list_of_reads", "listOfReads"), ("get_color", "_color"), ("get_component", "_component_ID"),
                 ("get_forward_edge_hashes", "forwardEdgeHashes"), ("get_backward_edge_hashes", "backwardEdgeHashes"),
                 ("get_node_Id", "_nodeId")],
        setters=[("set_component", "_component_ID", int), ("assign_node_Id", "_nodeId")],
        steppers=[("increment_node_coverage", "nodeCoverage", 1), ("extend_node_coverage", "nodeCoverage", None)])
class Node:
    def __init__(self, geneMer: GeneMer):
        self._fill(geneMer, geneMer.__hash__(), 0, [], None)

    @classmethod
    def from_arrays(cls, geneMer: GeneMer, node_hash: int, coverage: int, reads: list, component: int) -> "Node":
        """assemble a node from the exported device arrays (no hashing of the gene-mer)"""
        node = object.__new__(cls)
        node._fill(geneMer, node_hash, coverage, reads, component)
        return node

    def _fill(self, geneMer, node_hash, coverage, reads, component):
        self.geneMer = geneMer
        self.canonicalGeneMer, self.reverseGeneMer = geneMer.canonicalGeneMer, geneMer.rcGeneMer
        self.geneMerHash = node_hash
        self.nodeCoverage, self.listOfReads = coverage, reads
        self.forwardEdgeHashes, self.backwardEdgeHashes = [], []
        self._color, self._component_ID = None, component

    def get_reads(self):
        return iter(self.listOfReads)

    # reads and edge lists keep first-touch order and hold each entry once (construct_node.py:64-101)
    def add_read(self, read):
        _append_once(self.listOfReads, read)

    def remove_read(self, read):
        _take_out(self.listOfReads, read, "This node does not contain the read: " + str(read))

    def add_forward_edge_hash(self, forwardEdgeHash):
        _append_once(self.forwardEdgeHashes, forwardEdgeHash)
        return self

    def add_backward_edge_hash(self, backwardEdgeHash):
        _append_once(self.backwardEdgeHashes, backwardEdgeHash)
        return self

    def remove_forward_edge_hash(self, edgeHash):
        _take_out(self.forwardEdgeHashes, edgeHash, "This edge hash is not in the list of forward edge hashes")

    def remove_backward_edge_hash(self, edgeHash):
        _take_out(self.backwardEdgeHashes, edgeHash, "This edge hash is not in the list of backward edge hashes")

    def color_node(self, listOfAMRGenes):
        """0: no AMR gene; 1: AMR gene, degree <= 2; 2: AMR gene at a junction"""
        names = {g.get_name() for g in self.canonicalGeneMer}
        if names.isdisjoint(listOfAMRGenes):
            self._color = 0
        else:
            self._color = 2 if len(self.forwardEdgeHashes) + len(self.backwardEdgeHashes) > 2 else 1

    def __eq__(self, otherNode):
        return self.__hash__() == otherNode.__hash__() and self.nodeCoverage == otherNode.get_node_coverage()

    def __hash__(self):
        return self.geneMerHash
