"""Node of the gene-mer graph.  Mirror of upstream amira/construct_node.py: downstream correction and
path-finding code reaches into these attributes, so their names and list semantics are kept."""
from __future__ import annotations

from .construct_gene_mer import GeneMer


def _drop(lst: list, item, message: str):
    assert item in lst, message
    del lst[lst.index(item)]


class Node:
    def __init__(self, geneMer: GeneMer):
        self.geneMer = geneMer
        self.canonicalGeneMer = geneMer.get_canonical_geneMer()
        self.reverseGeneMer = geneMer.get_rc_geneMer()
        self.geneMerHash = geneMer.__hash__()
        self.nodeCoverage = 0
        self.listOfReads = []
        self.forwardEdgeHashes = []
        self.backwardEdgeHashes = []
        self._color = None
        self._component_ID = None

    @classmethod
    def from_arrays(cls, geneMer: GeneMer, node_hash: int, coverage: int, reads: list, component: int) -> "Node":
        """assemble a node from the exported device arrays (no hashing of the gene-mer)"""
        n = object.__new__(cls)
        n.geneMer = geneMer
        n.canonicalGeneMer = geneMer.canonicalGeneMer
        n.reverseGeneMer = geneMer.rcGeneMer
        n.geneMerHash = node_hash
        n.nodeCoverage = coverage
        n.listOfReads = reads
        n.forwardEdgeHashes = []
        n.backwardEdgeHashes = []
        n._color = None
        n._component_ID = component
        return n

    # -- accessors -------------------------------------------------------------------------
    def get_geneMer(self):
        return self.geneMer

    def get_canonical_geneMer(self):
        return self.canonicalGeneMer

    def get_reverse_geneMer(self):
        return self.reverseGeneMer

    def get_node_coverage(self) -> int:
        return self.nodeCoverage

    def get_list_of_reads(self) -> list:
        return self.listOfReads

    def get_reads(self):
        yield from self.listOfReads

    def get_color(self):
        return self._color

    def get_component(self):
        return self._component_ID

    def get_forward_edge_hashes(self):
        return self.forwardEdgeHashes

    def get_backward_edge_hashes(self):
        return self.backwardEdgeHashes

    def get_node_Id(self):
        return self._nodeId

    # -- mutators --------------------------------------------------------------------------
    def increment_node_coverage(self) -> int:
        self.nodeCoverage += 1
        return self.nodeCoverage

    def extend_node_coverage(self, value) -> int:
        self.nodeCoverage += value
        return self.nodeCoverage

    def set_component(self, new_component_ID) -> int:
        self._component_ID = int(new_component_ID)
        return self._component_ID

    def add_read(self, read):
        if read not in self.listOfReads:
            self.listOfReads.append(read)

    def remove_read(self, read):
        _drop(self.listOfReads, read, "This node does not contain the read: " + str(read))

    def add_forward_edge_hash(self, forwardEdgeHash):
        if forwardEdgeHash not in self.forwardEdgeHashes:
            self.forwardEdgeHashes.append(forwardEdgeHash)
        return self

    def remove_forward_edge_hash(self, edgeHash):
        _drop(self.forwardEdgeHashes, edgeHash, "This edge hash is not in the list of forward edge hashes")

    def add_backward_edge_hash(self, backwardEdgeHash):
        if backwardEdgeHash not in self.backwardEdgeHashes:
            self.backwardEdgeHashes.append(backwardEdgeHash)
        return self

    def remove_backward_edge_hash(self, edgeHash):
        _drop(self.backwardEdgeHashes, edgeHash, "This edge hash is not in the list of backward edge hashes")

    def assign_node_Id(self, nodeId):
        self._nodeId = nodeId
        return self._nodeId

    def color_node(self, listOfAMRGenes):
        """0: no AMR gene; 1: AMR gene, degree <= 2; 2: AMR gene at a junction"""
        if not any(g.get_name() in listOfAMRGenes for g in self.canonicalGeneMer):
            self._color = 0
        else:
            degree = len(self.forwardEdgeHashes) + len(self.backwardEdgeHashes)
            self._color = 2 if degree > 2 else 1

    def __eq__(self, otherNode):
        return self.__hash__() == otherNode.__hash__() and self.nodeCoverage == otherNode.get_node_coverage()

    def __hash__(self):
        return self.geneMerHash
