"""GeneMerGraph: drop-in for upstream ``amira.construct_graph.GeneMerGraph`` on the graph-build path.

The constructor (upstream construct_graph.py:31-102), ``filter_graph`` (:523-540) and
``remove_low_coverage_components`` (:950-958) run on the GPU through the C ABI in
``include/amira_gmg.h``; this module encodes the read dict, calls the library and materialises the
exported arrays into the dictionaries and ``Node`` / ``Edge`` objects downstream code reaches into:
``_nodes`` / ``_edges`` keyed by upstream's SHA-256 integers in upstream's insertion order,
``_readNodes`` / ``_readNodeDirections`` / ``_readNodePositions`` per read, ``_shortReads``,
``_readsToCorrect``.  The single-object accessors and mutators of the same surface
(``add_node``, ``remove_node``, ...) are plain host code over those objects, as upstream's are.

The host objects are materialised LAZILY: the constructor only encodes the reads and enqueues the device
build; the dictionaries and ``Node`` / ``Edge`` objects are created the first time something reaches for
``_nodes`` / ``_edges`` / the per-read lists.  Filters, component removal, the coverage statistics, the junk-read
and valid-read selections and the GML writer work on the device arrays and never need them.

There is no CPU build: without ``libamira_gmg.so`` or without a CUDA device construction fails.

To run upstream's own correction / path-finding methods on top of the GPU build, see
``bind_upstream`` below and INTEGRATION.md.
"""
from __future__ import annotations

import os
import statistics
import weakref
import zlib

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


_LAZY_ATTRS = ("_nodes", "_edges", "_readNodes", "_readNodeDirections", "_readNodePositions", "_shortReads",
               "_readsToCorrect")


def _lazy_property(name):
    slot = "_m" + name

    def get(self):
        if self.__dict__.get("_lazy"):
            self._materialise_now()
        return self.__dict__[slot]

    def set_(self, value):
        self.__dict__[slot] = value

    return property(get, set_)


class GeneMerGraph:
    _cls = _Classes
    for _a in _LAZY_ATTRS:
        locals()[_a] = _lazy_property(_a)
    del _a

    def __init__(self, readDict, kmerSize, gene_positions=None, device=None):
        # an encode.EncodedReads in place of the dict: the strings were parsed once, reuse the CSR
        self._lazy = False
        self._encoded = readDict if isinstance(readDict, encode.EncodedReads) else None
        if self._encoded is not None:
            if gene_positions is not None and gene_positions is not self._encoded.positions:
                # explicit positions win: encode them against the same reads (the cached CSR has none / others)
                self._encoded = encode.EncodedReads(self._encoded.reads, gene_positions)
            gene_positions = self._encoded.positions
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
        self._node_objs = []             # Node objects in device index order
        self._read_ids = []
        self._win_off = None
        self._device_ops = ()
        self._steps_cache = None
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
        self._csr_fingerprint = (len(ids), int(off[-1]) if len(off) else 0, zlib.crc32(np.ascontiguousarray(ids).tobytes()))
        self._read_len = np.diff(off)
        h = _handle(self._device)
        h.owner = None
        if self._encoded is not None and hasattr(h, "build_resident"):
            h.build_resident(self._encoded, int(self._kmerSize), self._device)
        else:
            h.build(ids, off, int(self._kmerSize), ps, pe)
        h.owner = weakref.ref(self)
        self._read_ids = list(self._reads)
        self._device_synced = True
        self._lazy = True                # host objects on first use (_materialise_now)

    def _materialise_now(self):
        """create the host dictionaries and objects from the device arrays (first access to _nodes / _edges / ...)"""
        if not self.__dict__.get("_lazy"):
            return
        self._lazy = False
        h = self._require_device_state()
        # 10^5..10^6 long-lived objects are created in one go: the cyclic collector would re-scan them generation by
        # generation on the way (measured on C2: 1.3 s -> 0.8 s without it); nothing here can form garbage cycles
        import gc
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            self._materialise(h.arrays(), self._vocab)
        finally:
            if was_enabled:
                gc.enable()

    def _arrays(self, *fields):
        """current device arrays of this graph (no host objects involved)"""
        h = self._require_device_state()
        a = h.arrays()
        return a if not fields else tuple(a[f] for f in fields)

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
        self._node_objs = node_objs
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
        # per-read lists (windows of removed nodes are None: the arrays may come from a filtered device graph)
        woff = a["win_off"].tolist()
        self._win_off = woff
        win = a["win_node"]
        gone = win < 0
        any_gone = bool(gone.any())
        if n_nodes:
            nh_arr = np.empty(n_nodes + 1, object)
            nh_arr[:n_nodes] = node_hashes
            nh_arr[n_nodes] = None
            wn = nh_arr[np.where(gone, n_nodes, win)].tolist()
        else:
            wn = [None] * len(win)
        wd = a["win_dir"].tolist()
        if any_gone:
            wd = [None if g else d for d, g in zip(wd, gone.tolist())]
        has_pos = bool(self._genePositions)
        if has_pos:
            wp = list(zip(a["win_start"].tolist(), a["win_end"].tolist()))
            if any_gone:
                wp = [None if g else p for p, g in zip(wp, gone.tolist())]
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
        self._readsToCorrect.update(rid for rid, c in zip(read_ids, a["to_correct"].tolist()) if c)

    def _require_device_state(self):
        """the device copy of this graph, rebuilt if another graph has used the handle since"""
        if not self._device_synced:
            raise RuntimeError(
                "this graph was modified on the host after its device build; the GPU filter works on the "
                "device copy (there is no CPU filter path). Rebuild it with GeneMerGraph(reads, k) first.")
        h = _handle(self._device)
        if h.owner is None or h.owner() is not self:
            vocab, ids, off, ps, pe = self._encode()
            fp = (len(ids), int(off[-1]) if len(off) else 0, zlib.crc32(np.ascontiguousarray(ids).tobytes()))
            if vocab.names != self._vocab.names or fp != self._csr_fingerprint:
                raise RuntimeError("the read dict of this graph changed since it was built; rebuild it first")
            h.owner = None
            h.build(ids, off, int(self._kmerSize), ps, pe)
            for op, args in self._device_ops:
                getattr(h, op)(*args)
            h.owner = weakref.ref(self)
        return h

    def _apply_device_removal(self, h):
        """mirror the device's last removal on the host objects (same deletions upstream performs)"""
        self._steps_cache = None
        if self.__dict__.get("_lazy"):
            return                       # nothing materialised yet: the arrays are the graph
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
        self._node_objs = [n for n, kp in zip(self._node_objs, node_keep.tolist()) if kp]
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
        if self.get_total_number_of_nodes() == 0:
            return self
        h = self._require_device_state()
        h.filter_graph(max(int(minNodeCoverage), 0), max(int(minEdgeCoverage), 0))
        self._device_ops = self._device_ops + (("filter_graph", (max(int(minNodeCoverage), 0), max(int(minEdgeCoverage), 0))),)
        self._apply_device_removal(h)
        return self

    def remove_low_coverage_components(self, min_component_coverage):
        """upstream construct_graph.py:950-958"""
        if self.get_total_number_of_nodes() == 0:
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
        if self._on_device():
            nodes = self._nodes                                   # materialises; keeps _node_objs in device order
            rank = self._gene_ranks([geneOfInterest])
            if not rank:
                return []
            flags = self._require_device_state().nodes_containing(rank)
            return [self._node_objs[i] for i in np.flatnonzero(flags).tolist()]
        return [n for n in self._nodes.values()
                if geneOfInterest in [g.get_name() for g in n.get_canonical_geneMer()]]

    def _gene_ranks(self, names) -> list:
        """SHA ranks (1..V) of the gene names that occur in this graph's vocabulary"""
        if getattr(self, "_rank_of", None) is None:
            self._rank_of = {n: i + 1 for i, n in enumerate(self._vocab.names)}
        return [self._rank_of[n] for n in names if n in self._rank_of]

    def get_AMR_nodes(self, listOfGenes):
        """upstream construct_graph.py:963-972: {node hash: node} of the nodes that contain any of the genes"""
        if self._on_device():
            nodes = self._nodes
            h = self._require_device_state()
            out = {}
            for g in listOfGenes:                  # upstream's dict order: by gene, then by node
                for r in self._gene_ranks([g]):
                    for i in np.flatnonzero(h.nodes_containing([r])).tolist():
                        out[self._node_order[i]] = self._node_objs[i]
            return out
        out = {}
        for g in listOfGenes:
            for node in self.get_nodes_containing(g):
                out[node.__hash__()] = node
        return out

    def remove_non_AMR_associated_nodes(self, genesOfInterest):
        """upstream construct_graph.py:2941-2959: drop every node that shares no read with a node holding one of the
        genes.  On the device: flag nodes, mark their reads, keep the nodes that touch a marked read, compact."""
        if self._on_device():
            for g in genesOfInterest:
                assert not (g[0] == "+" or g[0] == "-"), "Strand information cannot be present for any specified genes"
            h = self._require_device_state()
            ranks = self._gene_ranks(list(genesOfInterest))
            h.remove_nodes_without_reads_of(ranks)
            self._device_ops = self._device_ops + (("remove_nodes_without_reads_of", (tuple(ranks),)),)
            self._apply_device_removal(h)
            return
        readsOfInterest = set()
        for g in genesOfInterest:
            for node in self.get_nodes_containing(g):
                readsOfInterest.update(node.get_reads())
        doomed = [n for n in self._nodes.values() if not readsOfInterest.intersection(n.get_list_of_reads())]
        for node in doomed:
            self.remove_node(node)

    # ------------------------------------------------------------------ linear paths (upstream :722-861, 679-720)
    def _linear_steps(self) -> dict:
        """per node and side: upstream's one step of a linear-path walk (next node index, entry direction, extend flag)
        and the node degrees -- from the device adjacency CSR (k_linear_steps), or from the host objects when the
        graph was edited on the host since"""
        if self._steps_cache is not None and self._on_device():
            return self._steps_cache
        nodes = self._nodes
        if self._on_device():
            st = {k: v.tolist() for k, v in self._require_device_state().linear_steps().items()}
            st["index"] = {h: i for i, h in enumerate(self._node_order)}
            st["hash"] = self._node_order
            self._steps_cache = st
            return st
        order = list(nodes)
        index = {h: i for i, h in enumerate(order)}
        st = {"index": index, "hash": order, "degree": [self.get_degree(n) for n in nodes.values()]}
        for side, getter in (("fw", "get_forward_edge_hashes"), ("bw", "get_backward_edge_hashes")):
            nxt, dr, ext = [], [], []
            for h, node in nodes.items():
                eh = getattr(node, getter)()
                take = len(eh) == 1 if side == "fw" else len(eh) > 0
                if not take:
                    nxt.append(-1); dr.append(0); ext.append(0)
                    continue
                e = self._edges[eh[0]]
                t = e.get_targetNode()
                nxt.append(index[t.__hash__()])
                dr.append(e.get_targetNodeDirection())
                ext.append(int(self.get_degree(t) in (1, 2) and t != node))
            st[side + "_next"], st[side + "_dir"], st[side + "_ext"] = nxt, dr, ext
        return st

    def _walk(self, st, start: int, first_side: str, backward: bool, want_branched: bool) -> list:
        """get_forward_path_from_node / get_backward_path_from_node on node indices"""
        path = [start]
        ext, nx, d = st[first_side + "_ext"][start], st[first_side + "_next"][start], st[first_side + "_dir"][start]
        while ext:
            if path[-1 if backward else 0] == nx:
                break                                   # (upstream compares with the far end of the list and stops)
            if backward:
                path.insert(0, nx)
            else:
                path.append(nx)
            side = ("bw" if d == -1 else "fw") if backward else ("fw" if d == 1 else "bw")
            ext, nx, d = st[side + "_ext"][nx], st[side + "_next"][nx], st[side + "_dir"][nx]
        if want_branched and nx is not None and nx >= 0:
            if backward:
                path.insert(0, nx)
            else:
                path.append(nx)
        return path

    def get_forward_path_from_node(self, node, startDirection, wantBranchedNode=False) -> list:
        st = self._linear_steps()
        i = st["index"][node.__hash__()]
        return [st["hash"][j] for j in self._walk(st, i, "fw" if startDirection == 1 else "bw", False, wantBranchedNode)]

    def get_backward_path_from_node(self, node, startDirection, wantBranchedNode=False) -> list:
        st = self._linear_steps()
        i = st["index"][node.__hash__()]
        return [st["hash"][j] for j in self._walk(st, i, "bw" if startDirection == -1 else "fw", True, wantBranchedNode)]

    def get_linear_path_for_node(self, node, wantBranchedNode=False) -> list:
        """upstream construct_graph.py:849-861"""
        d = node.get_geneMer().get_geneMerDirection()
        back = self.get_backward_path_from_node(node, -1 * d, wantBranchedNode)
        assert back[-1] == node.__hash__()
        fwd = self.get_forward_path_from_node(node, d, wantBranchedNode)
        assert fwd[0] == node.__hash__()
        return back[:-1] + [node.__hash__()] + fwd[1:]

    def remove_short_linear_paths(self, min_length, sample_genesOfInterest={}):
        """upstream construct_graph.py:679-720: dead ends (degree-1 nodes) whose linear path is shorter than min_length
        are removed unless the whole path is well covered, holds an AMR gene or is its whole component.  The paths
        come from the device step table; the removal is one device pass (amira_gmg_remove_nodes)."""
        st = self._linear_steps()
        nodes = self._nodes
        mean15 = self.get_mean_node_coverage() * 1.5 if len(nodes) else 0
        paths_to_remove = {}
        for i, node in enumerate(list(nodes.values())):
            if st["degree"][st["index"][node.__hash__()]] != 1:
                continue
            path = self.get_linear_path_for_node(node)
            if 0 < len(path) < min_length:
                if all(nodes[n].get_node_coverage() > mean15 for n in path):
                    continue
                paths_to_remove.setdefault(node.get_component(), []).append(path)
        AMR_nodes = self.get_AMR_nodes(sample_genesOfInterest)
        removed = []
        seen = set()
        for component, paths in paths_to_remove.items():
            in_comp = {n.__hash__() for n in self.get_nodes_in_component(component)} if component is not None else set()
            for path in paths:
                if component is not None and len(in_comp.intersection(path)) == len(in_comp):
                    continue
                for nh in path:
                    if nh in AMR_nodes or nh in seen:
                        continue
                    seen.add(nh)
                    removed.append(nh)
        if not removed:
            return []
        if self._on_device():
            h = self._require_device_state()
            flags = np.zeros(len(self._node_order), np.uint8)
            index = st["index"]
            flags[[index[nh] for nh in removed]] = 1
            h.remove_nodes(flags)
            self._device_ops = self._device_ops + (("remove_nodes", (flags,)),)
            self._apply_device_removal(h)
        else:
            for nh in removed:
                self.remove_node(nodes[nh])
        return removed

    def _on_device(self) -> bool:
        """the device copy still is this graph (no host-side mutation since the build / last device operation)"""
        return bool(self._device_synced) and len(self._reads) > 0

    def get_total_number_of_nodes(self) -> int:
        if self.__dict__.get("_lazy"):
            return self._require_device_state().sizes_early()["nodes"]
        return len(self._nodes)

    def get_total_number_of_edges(self) -> int:
        if self.__dict__.get("_lazy"):
            return self._require_device_state().sizes_early()["edges"]
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
        (None) nodes are rejected.  On the device: one thread per read counts its None windows (k_junk_read_mask)."""
        reads, positions = self._reads, self._genePositions
        kept, kept_pos, rejected, rejected_pos = {}, {}, {}, {}
        if self._on_device():
            mask = self._require_device_state().junk_read_mask(error_rate).tolist()
            for read_id, m in zip(self._read_ids, mask):
                if m == 2:
                    continue             # short read: not in _readNodes
                (kept if m else rejected)[read_id] = reads[read_id]
                (kept_pos if m else rejected_pos)[read_id] = positions[read_id]
            return kept, kept_pos, rejected, rejected_pos
        for read_id, nodes in self._readNodes.items():
            ok = nodes.count(None) <= round(len(nodes) * (1 - error_rate))
            (kept if ok else rejected)[read_id] = reads[read_id]
            (kept_pos if ok else rejected_pos)[read_id] = positions[read_id]
        return kept, kept_pos, rejected, rejected_pos

    def get_valid_reads_only(self):
        """upstream construct_graph.py:1422-1427"""
        if self.__dict__.get("_lazy"):
            to_correct = self._arrays("to_correct")[0].tolist()
            return {r: self._reads[r] for r, bad in zip(self._read_ids, to_correct) if not bad}
        bad = self._readsToCorrect
        return {r: calls for r, calls in self._reads.items() if r not in bad}

    def get_all_node_coverages(self):
        if self.__dict__.get("_lazy"):
            return self._arrays("node_cov")[0].tolist()
        return [n.get_node_coverage() for n in self._nodes.values()]

    def get_mean_node_coverage(self):
        """upstream construct_graph.py:868-871 (statistics.mean of the coverages: exact, int when it divides)"""
        if self._on_device():
            n = self.get_total_number_of_nodes()
            if n == 0:
                raise statistics.StatisticsError("mean requires at least one data point")
            total, _ = self._require_device_state().node_coverage_stats()
            return total // n if total % n == 0 else total / n
        return statistics.mean(self.get_all_node_coverages())

    # ------------------------------------------------------------------ single-object mutators (host)
    def _touch(self):
        self._nodes                          # host edits need the host objects (materialise before going stale)
        self._device_synced = False
        self._steps_cache = None

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
        if self.__dict__.get("_lazy"):
            return sorted(set(self._arrays("node_comp")[0].tolist()))
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

    def _gml_from_arrays(self) -> list:
        """the GML entries of upstream generate_gml (construct_graph.py:873-909) straight from the device arrays: node id =
        insertion index, label from the canonical gene-mer, reads by id, edges in forward-then-backward list order"""
        a = self._arrays()
        names = self._vocab.names
        signed = [None] * (2 * len(names) + 1)
        V = len(names)
        for r, n in enumerate(names, 1):
            signed[V + r], signed[V - r] = "+" + n, "-" + n
        key = (a["node_key"].astype(np.int64) + V).tolist()
        cov, comp = a["node_cov"].tolist(), a["node_comp"].tolist()
        roff, reads = a["node_reads_off"].tolist(), a["node_reads"].tolist()
        fo, fe, bo, be = a["fw_off"].tolist(), a["fw_edges"].tolist(), a["bw_off"].tolist(), a["bw_edges"].tolist()
        tgt, sd, td, ecov = a["edge_tgt"].tolist(), a["edge_sd"].tolist(), a["edge_td"].tolist(), a["edge_cov"].tolist()
        rid = self._read_ids
        out = ["graph\t[", "multigraph 1"]
        for i in range(len(cov)):
            out.append(self.write_node_entry(i, "~~~".join(signed[g] for g in key[i]), cov[i],
                                             [rid[r] for r in reads[roff[i]:roff[i + 1]]], comp[i], None))
            for e in fe[fo[i]:fo[i + 1]] + be[bo[i]:bo[i + 1]]:
                if ecov[e] == 0:
                    continue
                out.append(self.write_edge_entry(i, tgt[e], sd[e], td[e], ecov[e]))
        out.append("]")
        return out

    def generate_gml(self, output_file: str, geneMerSize: int, min_node_coverage: int, min_edge_coverage: int):
        if self.__dict__.get("_lazy") or (self._on_device() and not any(n.get_color() for n in self._node_objs)):
            graph_data = self._gml_from_arrays()
            if not self.__dict__.get("_lazy"):
                self.assign_Id_to_nodes()
            self.write_gml_to_file(".".join([output_file, str(geneMerSize), str(min_node_coverage), str(min_edge_coverage)]),
                                   graph_data)
            return graph_data
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
    gpu_methods = ("__init__", "_encode", "_build_on_device", "_materialise", "_materialise_now", "_arrays", "_on_device",
                   "_require_device_state", "_apply_device_removal", "filter_graph", "remove_low_coverage_components",
                   "_touch", "remove_junk_reads", "get_valid_reads_only", "get_total_number_of_nodes",
                   "get_total_number_of_edges", "get_all_node_coverages", "get_mean_node_coverage", "components",
                   "get_nodes_containing", "_gene_ranks", "get_AMR_nodes", "remove_non_AMR_associated_nodes", "_linear_steps",
                   "_walk", "get_forward_path_from_node", "get_backward_path_from_node", "get_linear_path_for_node",
                   "remove_short_linear_paths", "_gml_from_arrays", "generate_gml") + _LAZY_ATTRS
    ns = {name: ours.__dict__[name] for name in gpu_methods}
    ns["_cls"] = _Up
    # host mutators must mark the device copy stale
    for name in ("add_node", "add_node_to_read", "add_node_to_nodes", "add_edge_to_edges", "add_edge_to_node",
                 "remove_edge", "remove_edge_from_edges", "remove_node_from_reads"):
        base = getattr(up.GeneMerGraph, name)

        def wrapped(self, *a, _base=base, **kw):
            self._nodes                      # host edits need the host objects
            self._device_synced = False
            return _base(self, *a, **kw)

        wrapped.__name__ = name
        ns[name] = wrapped
    return type("GeneMerGraph", (up.GeneMerGraph,), ns)
