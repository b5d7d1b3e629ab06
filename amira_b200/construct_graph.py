"""GeneMerGraph: drop-in for upstream ``amira.construct_graph.GeneMerGraph`` on the graph-build path.

The constructor (upstream construct_graph.py:31-102), ``filter_graph`` (:523-540) and
``remove_low_coverage_components`` (:950-958) run on the GPU through the C ABI in
``include/amira_gmg.h``; this module encodes the read dict, calls the library and materialises the
exported arrays into the dictionaries and ``Node`` / ``Edge`` objects downstream code reaches into:
``_nodes`` / ``_edges`` keyed by upstream's SHA-256 integers in upstream's insertion order,
``_readNodes`` / ``_readNodeDirections`` / ``_readNodePositions`` per read, ``_shortReads``,
``_readsToCorrect``.  The single-object accessors and mutators of the same surface
(``add_node``, ``remove_node``, ...) are plain host code over those objects, as upstream's are.

There is no CPU build: without ``libamira_gmg.so`` or without a CUDA device construction fails.

To run upstream's own correction / path-finding methods on top of the GPU build, see
``bind_upstream`` below and INTEGRATION.md.
"""
from __future__ import annotations

import os
import statistics
import weakref

import numpy as np

from . import _keys, _lib, encode
from .construct_edge import Edge, edge_key
from .construct_gene import Gene, convert_int_strand_to_string, hashlib_hash
from .construct_gene_mer import GeneMer
from .construct_node import Node
from .device_graph import DeviceGraph

_HANDLES: dict = {}      # device index -> DeviceGraph shared by the graphs built on that device


def _device_index(device) -> int:
    if device is None:
        return int(os.environ.get("AMIRA_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return int(device)


def _handle(device: int) -> DeviceGraph:
    h = _HANDLES.get(device)
    if h is None:
        h = _HANDLES[device] = DeviceGraph(device)
        h.owner = None
    return h


class _Classes:
    """the element classes a graph is materialised with (ours, or upstream's when bound)"""
    Gene, GeneMer, Node, Edge = Gene, GeneMer, Node, Edge

    @staticmethod
    def make_gene(name, strand):
        return Gene.from_parts(name, strand)

    @staticmethod
    def make_genemer(canonical, rc, direction, node_hash):
        return GeneMer.from_canonical(canonical, rc, direction, node_hash)

    @staticmethod
    def make_node(genemer, node_hash, cov, reads, comp):
        return Node.from_arrays(genemer, node_hash, cov, reads, comp)

    @staticmethod
    def make_edge(src, tgt, sd, td, cov):
        e = Edge(src, tgt, sd, td)
        e.edgeCoverage = cov
        return e


class GeneMerGraph:
    _cls = _Classes

    def __init__(self, readDict, kmerSize, gene_positions=None, device=None):
        # an encode.EncodedReads in place of the dict: the strings were parsed once, reuse the CSR
        self._encoded = readDict if isinstance(readDict, encode.EncodedReads) else None
        if self._encoded is not None:
            gene_positions = self._encoded.positions if gene_positions is None else gene_positions
            readDict = self._encoded.reads
        self._reads = readDict
        self._kmerSize = kmerSize
        self._minNodeCoverage = 1
        self._minEdgeCoverage = 1
        self._genePositions = gene_positions
        self._nodes = {}
        self._edges = {}
        self._readNodes = {}
        self._readNodeDirections = {}
        self._readNodePositions = {}
        self._shortReads = {}
        self._readsToCorrect = set()
        self._device = _device_index(device)
        self._device_synced = False      # device state == host objects
        self._node_order = []            # node hashes in device index order
        self._edge_order = []
        self._read_ids = []
        self._win_off = None
        self.timings = {}
        _lib.load()                      # fail loudly if the CUDA library is missing
        if len(readDict) > 0:
            self._build_on_device()

    # ------------------------------------------------------------------ device build
    def _encode(self):
        if self._encoded is not None:
            e = self._encoded
            return e.vocab, e.ids, e.off, e.pos_start, e.pos_end
        reads = self._reads
        vocab = encode.Vocabulary(encode.collect_names(reads))
        positions = self._genePositions if self._genePositions else None
        ids, off, ps, pe = encode.encode_reads(reads, vocab, positions)
        return vocab, ids, off, ps, pe

    def _build_on_device(self):
        vocab, ids, off, ps, pe = self._encode()
        self._vocab = vocab
        h = _handle(self._device)
        h.owner = None
        if self._encoded is not None and hasattr(h, "build_resident"):
            h.build_resident(self._encoded, int(self._kmerSize), self._device)
        else:
            h.build(ids, off, int(self._kmerSize), ps, pe)
        arrays = h.arrays()
        h.owner = weakref.ref(self)
        self._materialise(arrays, vocab)
        # kept for array-level statistics (graph_utils.get_overall_mean_node_coverages)
        self._incidence_arrays = (arrays["node_reads"], np.diff(off), len(arrays["node_cov"]))
        self._device_synced = True

    def _materialise(self, a, vocab):
        C = self._cls
        reads = self._reads
        k = int(self._kmerSize)
        read_ids = list(reads)
        self._read_ids = read_ids
        V = len(vocab)
        # one Gene object per signed id, and its signed SHA integer
        genes = np.empty(2 * V + 1, object)
        for r, name in enumerate(vocab.names, 1):
            genes[V + r] = C.make_gene(name, 1)
            genes[V - r] = C.make_gene(name, -1)
        key = a["node_key"].astype(np.int64)
        n_nodes = key.shape[0]
        canon_rows = genes[key + V].tolist() if n_nodes else []
        rc_rows = genes[V - key[:, ::-1]].tolist() if n_nodes else []
        node_hashes, node_sha = _keys.node_keys(key, vocab)
        rid_arr = np.empty(len(read_ids), object)
        rid_arr[:] = read_ids
        nr_off = a["node_reads_off"].tolist()
        nr = rid_arr[a["node_reads"]].tolist() if len(a["node_reads"]) else []
        cov = a["node_cov"].tolist()
        ndir = a["node_dir"].tolist()
        comp = a["node_comp"].tolist()
        nodes = self._nodes
        node_objs = []
        for i in range(n_nodes):
            nh = node_hashes[i]
            gm = C.make_genemer(canon_rows[i], rc_rows[i], ndir[i], nh)
            node = C.make_node(gm, nh, cov[i], nr[nr_off[i]:nr_off[i + 1]], comp[i])
            nodes[nh] = node
            node_objs.append(node)
        self._node_order = node_hashes
        # edges
        src, tgt = a["edge_src"].tolist(), a["edge_tgt"].tolist()
        sd, td, ecov = a["edge_sd"].tolist(), a["edge_td"].tolist(), a["edge_cov"].tolist()
        edges = self._edges
        edge_hashes = _keys.edge_keys(node_hashes, node_sha, a["edge_src"], a["edge_tgt"], a["edge_sd"], a["edge_td"])
        for j, eh in enumerate(edge_hashes):
            edges[eh] = C.make_edge(node_objs[src[j]], node_objs[tgt[j]], sd[j], td[j], ecov[j])
        self._edge_order = edge_hashes
        if edge_hashes:
            eh_arr = np.empty(len(edge_hashes), object)
            eh_arr[:] = edge_hashes
            fw = eh_arr[a["fw_edges"]].tolist()
            bw = eh_arr[a["bw_edges"]].tolist()
            fo, bo = a["fw_off"].tolist(), a["bw_off"].tolist()
            for i, node in enumerate(node_objs):
                node.forwardEdgeHashes = fw[fo[i]:fo[i + 1]]
                node.backwardEdgeHashes = bw[bo[i]:bo[i + 1]]
        # per-read lists
        woff = a["win_off"].tolist()
        self._win_off = woff
        if n_nodes:
            nh_arr = np.empty(n_nodes, object)
            nh_arr[:] = node_hashes
            wn = nh_arr[a["win_node"]].tolist()
        else:
            wn = []
        wd = a["win_dir"].tolist()
        has_pos = bool(self._genePositions)
        if has_pos:
            wp = list(zip(a["win_start"].tolist(), a["win_end"].tolist()))
        short = a["is_short"].tolist()
        rn, rd, rp = self._readNodes, self._readNodeDirections, self._readNodePositions
        for i, rid in enumerate(read_ids):
            if short[i]:
                self._shortReads[rid] = reads[rid]
                continue
            lo, hi = woff[i], woff[i + 1]
            rn[rid] = wn[lo:hi]
            rd[rid] = wd[lo:hi]
            if has_pos and self._genePositions[rid]:
                rp[rid] = wp[lo:hi]
            else:
                rp[rid] = [None] * (hi - lo)

    def _require_device_state(self):
        """the device copy of this graph, rebuilt if another graph has used the handle since"""
        if not self._device_synced:
            raise RuntimeError(
                "this graph was modified on the host after its device build; the GPU filter works on the "
                "device copy (there is no CPU filter path). Rebuild it with GeneMerGraph(reads, k) first.")
        h = _handle(self._device)
        if h.owner is None or h.owner() is not self:
            vocab, ids, off, ps, pe = self._encode()
            if vocab.names != self._vocab.names:
                raise RuntimeError("the read dict of this graph changed since it was built; rebuild it first")
            h.owner = None
            h.build(ids, off, int(self._kmerSize), ps, pe)
            for op, args in self._device_ops:
                getattr(h, op)(*args)
            h.owner = weakref.ref(self)
        return h

    _device_ops: tuple = ()

    def _apply_device_removal(self, h):
        """mirror the device's last removal on the host objects (same deletions upstream performs)"""
        node_keep, edge_keep = h.filter_masks()
        if node_keep.all() and edge_keep.all():
            return
        edges, nodes = self._edges, self._nodes
        for j in np.flatnonzero(~edge_keep).tolist():
            eh = self._edge_order[j]
            e = edges.pop(eh)
            src = e.get_sourceNode()
            lst = src.forwardEdgeHashes if e.get_sourceNodeDirection() == 1 else src.backwardEdgeHashes
            lst.remove(eh)
        self._edge_order = [eh for eh, kp in zip(self._edge_order, edge_keep.tolist()) if kp]
        gone = np.flatnonzero(~node_keep).tolist()
        if not gone:
            return
        affected = set()
        for i in gone:
            node = nodes.pop(self._node_order[i])
            affected.update(node.get_list_of_reads())
        self._node_order = [nh for nh, kp in zip(self._node_order, node_keep.tolist()) if kp]
        # per-read lists: windows of removed nodes become None (remove_node_from_reads, :442-461)
        a = h.arrays_reads_only()
        wn = a["win_node"]
        woff = self._win_off
        rn, rd, rp = self._readNodes, self._readNodeDirections, self._readNodePositions
        ridx = {r: i for i, r in enumerate(self._read_ids)}
        for rid in affected:
            i = ridx[rid]
            mask = (wn[woff[i]:woff[i + 1]] < 0).tolist()
            rn[rid] = [None if m else v for v, m in zip(rn[rid], mask)]
            rd[rid] = [None if m else v for v, m in zip(rd[rid], mask)]
            rp[rid] = [None if m else v for v, m in zip(rp[rid], mask)]
        self._readsToCorrect.update(affected)

    # ------------------------------------------------------------------ GPU-backed operations
    def filter_graph(self, minNodeCoverage: int, minEdgeCoverage: int):
        """upstream construct_graph.py:523-540"""
        minNodeCoverage = self.set_minNodeCoverage(minNodeCoverage)
        minEdgeCoverage = self.set_minEdgeCoverage(minEdgeCoverage)
        if not self._nodes:
            return self
        h = self._require_device_state()
        h.filter_graph(max(int(minNodeCoverage), 0), max(int(minEdgeCoverage), 0))
        self._device_ops = self._device_ops + (("filter_graph", (max(int(minNodeCoverage), 0), max(int(minEdgeCoverage), 0))),)
        self._apply_device_removal(h)
        return self

    def remove_low_coverage_components(self, min_component_coverage):
        """upstream construct_graph.py:950-958"""
        if not self._nodes:
            return
        h = self._require_device_state()
        h.remove_low_coverage_components(max(int(min_component_coverage), 0))
        self._device_ops = self._device_ops + (("remove_low_coverage_components", (max(int(min_component_coverage), 0),)),)
        self._apply_device_removal(h)

    # ------------------------------------------------------------------ accessors
    def get_reads(self):
        return self._reads

    def get_short_read_annotations(self):
        return self._shortReads

    def get_gene_positions(self):
        return self._genePositions

    def get_short_read_gene_positions(self):
        return {r: self._genePositions[r] for r in self._shortReads}

    def get_readNodes(self):
        return self._readNodes

    def get_readNodeDirections(self):
        return self._readNodeDirections

    def get_readNodePositions(self):
        return self._readNodePositions

    def get_kmerSize(self) -> int:
        return self._kmerSize

    def get_minEdgeCoverage(self) -> int:
        return self._minEdgeCoverage

    def get_minNodeCoverage(self) -> int:
        return self._minNodeCoverage

    def get_nodes(self):
        return self._nodes

    def get_edges(self):
        return self._edges

    def get_reads_to_correct(self) -> set:
        return self._readsToCorrect

    def all_nodes(self):
        yield from self._nodes.values()

    def get_reads_for_nodes(self, list_of_nodes) -> set:
        reads = set()
        for node_hash in list_of_nodes:
            reads.update(self._nodes[node_hash].get_list_of_reads())
        return reads

    def get_nodes_containing_read(self, readId: str):
        return [self._nodes[h] for h in self._readNodes[readId] if h in self._nodes]

    def get_node_by_hash(self, nodeHash: int):
        return self._nodes[nodeHash]

    def get_edge_by_hash(self, edgeHash: int):
        return self._edges[edgeHash]

    def get_node(self, geneMer):
        nodeHash = geneMer.__hash__()
        assert nodeHash in self._nodes, "This gene-mer is not in the graph"
        return self._nodes[nodeHash]

    def get_nodes_containing(self, geneOfInterest: str):
        assert not (geneOfInterest[0] == "+" or geneOfInterest[0] == "-"), \
            "Strand information cannot be present for any specified genes"
        assert isinstance(geneOfInterest, str), "Gene of interest is the wrong type"
        return [n for n in self._nodes.values()
                if geneOfInterest in [g.get_name() for g in n.get_canonical_geneMer()]]

    def get_total_number_of_nodes(self) -> int:
        return len(self._nodes)

    def get_total_number_of_edges(self) -> int:
        return len(self._edges)

    def get_total_number_of_reads(self) -> int:
        return len(self._reads)

    def get_degree(self, node) -> int:
        return len(node.get_forward_edge_hashes()) + len(node.get_backward_edge_hashes())

    def get_forward_edges(self, node):
        return [self._edges[h] for h in node.get_forward_edge_hashes()]

    def get_backward_edges(self, node):
        return [self._edges[h] for h in node.get_backward_edge_hashes()]

    def get_forward_neighbors(self, node):
        return [e.get_targetNode() for e in self.get_forward_edges(node)]

    def get_backward_neighbors(self, node):
        return [e.get_targetNode() for e in self.get_backward_edges(node)]

    def get_all_neighbors(self, node):
        return self.get_forward_neighbors(node) + self.get_backward_neighbors(node)

    def get_all_neighbor_hashes(self, node) -> set:
        return {n.__hash__() for n in self.get_all_neighbors(node)}

    def check_if_nodes_are_adjacent(self, sourceNode, targetNode) -> bool:
        return (targetNode.__hash__() in self.get_all_neighbor_hashes(sourceNode)
                and sourceNode.__hash__() in self.get_all_neighbor_hashes(targetNode))

    def get_edge_hashes_between_nodes(self, sourceNode, targetNode):
        assert self.check_if_nodes_are_adjacent(sourceNode, targetNode)
        out = [e.__hash__() for e in self.get_forward_edges(sourceNode) + self.get_backward_edges(sourceNode)
               if e.get_targetNode() == targetNode]
        back = [e.__hash__() for e in self.get_forward_edges(targetNode) + self.get_backward_edges(targetNode)
                if e.get_targetNode() == sourceNode]
        if len(out) > 1 or len(back) > 1:
            return (out, back)
        return (out[0], back[0])

    def get_edges_between_nodes(self, sourceNode, targetNode):
        s2t, t2s = self.get_edge_hashes_between_nodes(sourceNode, targetNode)
        if isinstance(s2t, list) or isinstance(t2s, list):
            return [self._edges[h] for h in s2t], [self._edges[h] for h in t2s]
        return self._edges[s2t], self._edges[t2s]

    def remove_junk_reads(self, error_rate):
        """upstream construct_graph.py:1398-1420: reads with more than round(n * (1 - error_rate)) filtered
        (None) nodes are rejected"""
        reads, positions = self._reads, self._genePositions
        kept, kept_pos, rejected, rejected_pos = {}, {}, {}, {}
        for read_id, nodes in self._readNodes.items():
            ok = nodes.count(None) <= round(len(nodes) * (1 - error_rate))
            (kept if ok else rejected)[read_id] = reads[read_id]
            (kept_pos if ok else rejected_pos)[read_id] = positions[read_id]
        return kept, kept_pos, rejected, rejected_pos

    def get_valid_reads_only(self):
        """upstream construct_graph.py:1422-1427"""
        bad = self._readsToCorrect
        return {r: calls for r, calls in self._reads.items() if r not in bad}

    def get_all_node_coverages(self):
        return [n.get_node_coverage() for n in self._nodes.values()]

    def get_mean_node_coverage(self):
        return statistics.mean(self.get_all_node_coverages())

    # ------------------------------------------------------------------ single-object mutators (host)
    def _touch(self):
        self._device_synced = False

    def add_node_to_read(self, node, readId: str, node_direction: int, node_position=None):
        self._touch()
        if readId not in self._readNodes:
            self._readNodes[readId] = []
            self._readNodeDirections[readId] = []
            self._readNodePositions[readId] = []
        self._readNodes[readId].append(node.__hash__())
        self._readNodeDirections[readId].append(node_direction)
        self._readNodePositions[readId].append(node_position)
        return self._readNodes[readId]

    def add_node_to_nodes(self, node, nodeHash: int) -> None:
        self._touch()
        self._nodes[nodeHash] = node

    def add_node(self, geneMer, reads: list):
        self._touch()
        nodeHash = geneMer.__hash__()
        node = self._nodes.get(nodeHash)
        if node is None:
            node = self._cls.Node(geneMer)
            self._nodes[nodeHash] = node
        for r in reads:
            node.add_read(r)
        return node

    def create_edges(self, sourceNode, targetNode, sourceGeneMerDirection: int, targetGeneMerDirection: int):
        E = self._cls.Edge
        return (E(sourceNode, targetNode, sourceGeneMerDirection, targetGeneMerDirection),
                E(targetNode, sourceNode, targetGeneMerDirection * -1, sourceGeneMerDirection * -1))

    def add_edge_to_edges(self, edge):
        self._touch()
        return self._edges.setdefault(edge.__hash__(), edge)

    def add_edges_to_graph(self, sourceToTargetEdge, reverseTargetToSourceEdge):
        return self.add_edge_to_edges(sourceToTargetEdge), self.add_edge_to_edges(reverseTargetToSourceEdge)

    def add_edge_to_node(self, node, edge):
        self._touch()
        if edge.get_sourceNodeDirection() == 1:
            node.add_forward_edge_hash(edge.__hash__())
        if edge.get_sourceNodeDirection() == -1:
            node.add_backward_edge_hash(edge.__hash__())
        return node

    def add_edge(self, sourceGeneMer, targetGeneMer):
        sourceNode = self.add_node(sourceGeneMer, [])
        targetNode = self.add_node(targetGeneMer, [])
        fwd, rev = self.create_edges(sourceNode, targetNode, sourceGeneMer.get_geneMerDirection(),
                                     targetGeneMer.get_geneMerDirection())
        fwd, rev = self.add_edges_to_graph(fwd, rev)
        self.add_edge_to_node(sourceNode, fwd)
        self.add_edge_to_node(targetNode, rev)
        return fwd, rev

    def remove_edge_from_edges(self, edgeHash: int) -> None:
        self._touch()
        del self._edges[edgeHash]
        assert edgeHash not in self._edges, "This edge was not removed from the graph successfully"

    def remove_edge(self, edgeHash: int) -> None:
        edge = self._edges.get(edgeHash)
        if edge is None:
            return
        self._touch()
        if edge.get_sourceNodeDirection() == 1:
            edge.get_sourceNode().remove_forward_edge_hash(edgeHash)
        if edge.get_sourceNodeDirection() == -1:
            edge.get_sourceNode().remove_backward_edge_hash(edgeHash)
        del self._edges[edgeHash]

    def remove_node_from_reads(self, node_to_remove) -> None:
        self._touch()
        gone = node_to_remove.__hash__()
        for readId in node_to_remove.get_reads():
            keep = [h != gone for h in self._readNodes[readId]]
            for table in (self._readNodes, self._readNodeDirections, self._readNodePositions):
                table[readId] = [v if kp else None for v, kp in zip(table[readId], keep)]
            self._readsToCorrect.add(readId)

    def remove_node(self, node):
        nodeHash = node.__hash__()
        assert nodeHash in self._nodes, "This node is not in the graph"
        assert node == self._nodes[nodeHash]
        self.remove_node_from_reads(node)
        for edgeHash in set(node.get_forward_edge_hashes() + node.get_backward_edge_hashes()):
            targetNode = self._edges[edgeHash].get_targetNode()
            for e in self.get_edge_hashes_between_nodes(node, targetNode):
                self.remove_edge(e)
        del self._nodes[nodeHash]

    def set_minNodeCoverage(self, minNodeCoverage: int):
        self._minNodeCoverage = minNodeCoverage
        return self._minNodeCoverage

    def set_minEdgeCoverage(self, minEdgeCoverage: int):
        self._minEdgeCoverage = minEdgeCoverage
        return self._minEdgeCoverage

    def list_nodes_to_remove(self, minNodeCoverage: int):
        return {n for n in self._nodes.values() if not n.get_node_coverage() > minNodeCoverage - 1}

    def list_edges_to_remove(self, minEdgeCoverage: int, nodesToRemove):
        doomed = set()
        for eh, e in self._edges.items():
            if not e.get_edge_coverage() > minEdgeCoverage - 1:
                doomed.add(eh)
            if e.get_sourceNode() in nodesToRemove or e.get_targetNode() in nodesToRemove:
                doomed.add(eh)
        return doomed

    # ------------------------------------------------------------------ components
    def dfs_component(self, start, component_id, visited=None):
        if visited is None:
            visited = set()
        stack = [start]
        visited.add(start.__hash__())
        while stack:
            node = stack.pop()
            node.set_component(component_id)
            for nb in self.get_all_neighbors(node):
                if nb.__hash__() not in visited:
                    visited.add(nb.__hash__())
                    stack.append(nb)

    def assign_component_ids(self):
        """component ids 1, 2, ... in order of each component's first node (upstream :920-927)"""
        visited = set()
        component_id = 1
        for nodeHash, node in self._nodes.items():
            if nodeHash not in visited:
                self.dfs_component(node, component_id, visited)
                component_id += 1

    def get_nodes_in_component(self, component):
        return [n for n in self._nodes.values() if n.get_component() == int(component)]

    def components(self) -> list:
        return sorted({n.get_component() for n in self._nodes.values()})

    def get_number_of_component(self) -> int:
        return len(self.components())

    # ------------------------------------------------------------------ GML (upstream :542-654, 873-909)
    def write_node_entry(self, node_id, node_string, node_coverage, reads, component_ID, nodeColor):
        lines = ["\tnode\t[", "\t\tid\t%s" % node_id, '\t\tlabel\t"%s"' % node_string,
                 "\t\tcoverage\t%s" % node_coverage]
        if component_ID:
            lines.append("\t\tcomponent\t%s" % component_ID)
        lines.append('\t\treads\t"%s"' % ",".join(reads))
        if nodeColor:
            lines.append('\t\tcolor\t"%s"' % nodeColor)
        lines.append("\t]")
        return "\n".join(lines)

    def write_edge_entry(self, source_node, target_node, source_edge_direction, target_edge_direction, edge_coverage):
        return "\n".join(["\tedge\t[", "\t\tsource\t%s" % source_node, "\t\ttarget\t%s" % target_node,
                          "\t\tsource_direction\t%s" % source_edge_direction,
                          "\t\ttarget_direction\t%s" % target_edge_direction,
                          "\t\tweight\t%s" % edge_coverage, "\t]"])

    def assign_Id_to_nodes(self):
        for i, node in enumerate(self._nodes.values()):
            assert node.assign_node_Id(i) == i, "This node was assigned an incorrect ID"

    def write_gml_to_file(self, output_file, gml_content):
        d = os.path.dirname(output_file)
        if d != "" and not os.path.exists(d):
            os.mkdir(d)
        with open(output_file + ".gml", "w") as f:
            f.write("\n".join(gml_content))

    def get_gene_mer_genes(self, sourceNode) -> list:
        return [convert_int_strand_to_string(g.get_strand()) + g.get_name() for g in sourceNode.get_canonical_geneMer()]

    def get_reverse_gene_mer_genes(self, sourceNode) -> list:
        return [convert_int_strand_to_string(g.get_strand()) + g.get_name() for g in sourceNode.get_reverse_geneMer()]

    def get_gene_mer_label(self, sourceNode) -> str:
        return "~~~".join(self.get_gene_mer_genes(sourceNode))

    def generate_gml(self, output_file: str, geneMerSize: int, min_node_coverage: int, min_edge_coverage: int):
        graph_data = ["graph\t[", "multigraph 1"]
        self.assign_Id_to_nodes()
        for node in self._nodes.values():
            graph_data.append(self.write_node_entry(node.get_node_Id(), self.get_gene_mer_label(node),
                                                    node.get_node_coverage(), list(node.get_reads()),
                                                    node.get_component(), node.get_color()))
            for edge in self.get_forward_edges(node) + self.get_backward_edges(node):
                if edge.get_edge_coverage() == 0:
                    continue
                graph_data.append(self.write_edge_entry(node.get_node_Id(), edge.get_targetNode().get_node_Id(),
                                                        edge.get_sourceNodeDirection(), edge.get_targetNodeDirection(),
                                                        edge.get_edge_coverage()))
        graph_data.append("]")
        self.write_gml_to_file(".".join([output_file, str(geneMerSize), str(min_node_coverage), str(min_edge_coverage)]),
                               graph_data)
        return graph_data


def bind_upstream(upstream_construct_graph):
    """Subclass upstream's own GeneMerGraph so that every correction / path-finding method it defines
    runs unchanged on a graph built by the CUDA path.

        import amira.construct_graph as cg
        import amira_b200
        cg.GeneMerGraph = amira_b200.bind_upstream(cg)       # before amira.graph_utils is imported

    The element objects are upstream's own Gene / GeneMer / Node / Edge classes."""
    up = upstream_construct_graph
    import importlib
    gene_mod = importlib.import_module(up.GeneMer.__module__.rsplit(".", 1)[0] + ".construct_gene")
    UpGene, UpGeneMer, UpNode, UpEdge = gene_mod.Gene, up.GeneMer, up.Node, up.Edge

    class _Up:
        Gene, GeneMer, Node, Edge = UpGene, UpGeneMer, UpNode, UpEdge

        @staticmethod
        def make_gene(name, strand):
            g = object.__new__(UpGene)
            g.name, g.strand = name, strand
            return g

        @staticmethod
        def make_genemer(canonical, rc, direction, node_hash):
            gm = object.__new__(UpGeneMer)
            gm.canonicalGeneMer, gm.rcGeneMer = canonical, rc
            gm.geneMerSize = len(canonical)
            gm.geneMerDirection = direction
            return gm

        @staticmethod
        def make_node(genemer, node_hash, cov, reads, comp):
            n = object.__new__(UpNode)
            n.geneMer = genemer
            n.canonicalGeneMer = genemer.canonicalGeneMer
            n.reverseGeneMer = genemer.rcGeneMer
            n.geneMerHash = node_hash
            n.nodeCoverage = cov
            n.listOfReads = reads
            n.forwardEdgeHashes = []
            n.backwardEdgeHashes = []
            n._color = None
            n._component_ID = comp
            return n

        @staticmethod
        def make_edge(src, tgt, sd, td, cov):
            e = UpEdge(src, tgt, sd, td)
            e.edgeCoverage = cov
            return e

    ours = GeneMerGraph
    gpu_methods = ("__init__", "_encode", "_build_on_device", "_materialise", "_require_device_state",
                   "_apply_device_removal", "filter_graph", "remove_low_coverage_components", "_touch",
                   "remove_junk_reads", "get_valid_reads_only")
    ns = {name: ours.__dict__[name] for name in gpu_methods}
    ns["_cls"] = _Up
    ns["_device_ops"] = ()
    # host mutators must mark the device copy stale
    for name in ("add_node", "add_node_to_read", "add_node_to_nodes", "add_edge_to_edges", "add_edge_to_node",
                 "remove_edge", "remove_edge_from_edges", "remove_node_from_reads"):
        base = getattr(up.GeneMerGraph, name)

        def wrapped(self, *a, _base=base, **kw):
            self._device_synced = False
            return _base(self, *a, **kw)

        wrapped.__name__ = name
        ns[name] = wrapped
    return type("GeneMerGraph", (up.GeneMerGraph,), ns)
