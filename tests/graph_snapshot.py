"""snapshot of a GeneMerGraph-like object in the layout of tests/golden/small_cases.json"""


def snapshot(g):
    hx = lambda h: None if h is None else hex(h)
    nodes = list(g.get_nodes().values())
    edges = list(g.get_edges().values())
    return {
        "node_hashes": [hex(h) for h in g.get_nodes()],
        "node_cov": [n.get_node_coverage() for n in nodes],
        "node_canonical": [[("+" if x.get_strand() == 1 else "-") + x.get_name() for x in n.get_canonical_geneMer()]
                           for n in nodes],
        "node_first_dir": [n.get_geneMer().get_geneMerDirection() for n in nodes],
        "node_reads": [list(n.get_list_of_reads()) for n in nodes],
        "node_fw": [[hex(h) for h in n.get_forward_edge_hashes()] for n in nodes],
        "node_bw": [[hex(h) for h in n.get_backward_edge_hashes()] for n in nodes],
        "node_comp": [n.get_component() for n in nodes],
        "edge_hashes": [hex(h) for h in g.get_edges()],
        "edge_src": [hex(e.get_sourceNode().__hash__()) for e in edges],
        "edge_tgt": [hex(e.get_targetNode().__hash__()) for e in edges],
        "edge_sd": [e.get_sourceNodeDirection() for e in edges],
        "edge_td": [e.get_targetNodeDirection() for e in edges],
        "edge_cov": [e.get_edge_coverage() for e in edges],
        "read_nodes": {r: [hx(h) for h in v] for r, v in g.get_readNodes().items()},
        "read_dirs": {r: list(v) for r, v in g.get_readNodeDirections().items()},
        "read_pos": {r: [None if p is None else list(p) for p in v] for r, v in g.get_readNodePositions().items()},
        "short_reads": dict(g.get_short_read_annotations()),
        "reads_to_correct": sorted(g.get_reads_to_correct()),
        "min_node_cov": g.get_minNodeCoverage(), "min_edge_cov": g.get_minEdgeCoverage(),
    }


def check_small_cases(GeneMerGraph, cases, pytest):
    for case in cases:
        reads, k, pos = case["reads"], case["k"], case["positions"]
        if case.get("raises"):
            with pytest.raises(AssertionError) as ei:
                GeneMerGraph(reads, k, pos)
            assert str(ei.value) == case["message"], case["name"]
            continue
        got = snapshot(GeneMerGraph(reads, k, pos))
        assert got == case["build"], (case["name"], [f for f in got if got[f] != case["build"][f]])
        g = GeneMerGraph(reads, k, pos)
        if "raises" in case["rlcc5"]:
            with pytest.raises(TypeError):
                g.remove_low_coverage_components(5)
        else:
            g.remove_low_coverage_components(5)
            got = snapshot(g)
            assert got == case["rlcc5"], (case["name"], "rlcc5", [f for f in got if got[f] != case["rlcc5"][f]])
        g = GeneMerGraph(reads, k, pos)
        g.filter_graph(2, 2)
        got = snapshot(g)
        assert got == case["filter2_2"], (case["name"], "filter2_2", [f for f in got if got[f] != case["filter2_2"][f]])
