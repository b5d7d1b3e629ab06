"""Pins both oracles (oracle/gmg_oracle.py, oracle/gmg_oracle.c) to the golden vectors that
oracle/make_golden.py produced from the unmodified upstream GeneMerGraph."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import gmg_oracle as O
from tests.helpers import EXPECTED, GOLDEN_KEYS, STAGE_ORDER, apply_stage, load_input, read_dict


@pytest.mark.parametrize("key", GOLDEN_KEYS)
def test_c_oracle_matches_upstream_golden(key):
    exp = EXPECTED[key]
    vocab, ids, off, ps, pe = load_input(exp["input"])
    g = c_oracle.COracleGraph(ids, off, exp["k"], ps, pe)
    for stage in STAGE_ORDER:
        want = exp["stages"].get(stage)
        if want is None:
            break
        if "raises" in want:
            with pytest.raises(TypeError):
                apply_stage(g, stage)
            break
        apply_stage(g, stage)
        a = g.arrays()
        assert O.summary(a) == want["summary"], (key, stage)
        got = O.digest_arrays(a)
        assert got == want["digest"], (key, stage, [f for f in got if got[f] != want["digest"][f]])


PY_KEYS = [k for k in GOLDEN_KEYS if EXPECTED[k]["windows"] <= 90000]


@pytest.mark.parametrize("key", PY_KEYS)
def test_python_oracle_matches_upstream_golden(key):
    exp = EXPECTED[key]
    vocab, ids, off, ps, pe = load_input(exp["input"])
    reads, pos = read_dict(vocab, ids, off, ps, pe)
    g = O.OracleGraph(reads, exp["k"], pos)
    used = set(O.build_vocabulary(reads))
    assert O.build_vocabulary(reads) == [v for v in vocab if v in used]
    for stage in STAGE_ORDER:
        want = exp["stages"].get(stage)
        if want is None:
            break
        if "raises" in want:
            with pytest.raises(TypeError):
                apply_stage(g, stage)
            break
        apply_stage(g, stage)
        a = g.arrays(vocab)
        assert O.summary(a) == want["summary"], (key, stage)
        assert O.digest_arrays(a) == want["digest"], (key, stage)


def _small_snapshot(g):
    hx = lambda h: None if h is None else hex(h)
    nodes = list(g.nodes.values())
    edges = list(g.edges.values())
    return {
        "node_hashes": [hex(n.hash) for n in nodes],
        "node_cov": [n.cov for n in nodes],
        "node_canonical": [[("+" if s == 1 else "-") + nm for nm, s in n.names] for n in nodes],
        "node_first_dir": [n.first_dir for n in nodes],
        "node_reads": [list(n.reads) for n in nodes],
        "node_fw": [[hex(h) for h in n.fw] for n in nodes],
        "node_bw": [[hex(h) for h in n.bw] for n in nodes],
        "node_comp": [n.comp for n in nodes],
        "edge_hashes": [hex(e.hash) for e in edges],
        "edge_src": [hex(e.src.hash) for e in edges],
        "edge_tgt": [hex(e.tgt.hash) for e in edges],
        "edge_sd": [e.sd for e in edges], "edge_td": [e.td for e in edges], "edge_cov": [e.cov for e in edges],
        "read_nodes": {r: [hx(h) for h in v] for r, v in g.read_nodes.items()},
        "read_dirs": {r: list(v) for r, v in g.read_dirs.items()},
        "read_pos": {r: [None if p is None else list(p) for p in v] for r, v in g.read_pos.items()},
        "short_reads": dict(g.short_reads),
        "reads_to_correct": sorted(g.to_correct),
        "min_node_cov": g.min_node_cov, "min_edge_cov": g.min_edge_cov,
    }


def test_python_oracle_small_cases(golden_small):
    """string-level cases incl. the upstream SHA-256 node / edge keys and error behaviour"""
    assert len(golden_small) >= 20
    for case in golden_small:
        reads, k, pos = case["reads"], case["k"], case["positions"]
        if case.get("raises"):
            with pytest.raises(AssertionError) as ei:
                O.OracleGraph(reads, k, pos)
            assert str(ei.value) == case["message"], case["name"]
            continue
        assert _small_snapshot(O.OracleGraph(reads, k, pos)) == case["build"], case["name"]
        g = O.OracleGraph(reads, k, pos)
        if "raises" in case["rlcc5"]:
            with pytest.raises(TypeError):
                g.remove_low_coverage_components(5)
        else:
            g.remove_low_coverage_components(5)
            assert _small_snapshot(g) == case["rlcc5"], case["name"]
        g = O.OracleGraph(reads, k, pos)
        g.filter_graph(2, 2)
        assert _small_snapshot(g) == case["filter2_2"], case["name"]


def test_c_oracle_small_cases(golden_small):
    for case in golden_small:
        reads, k, pos = case["reads"], case["k"], case["positions"]
        if case.get("raises") and "identical" not in case["message"]:
            continue  # token-level errors are the encoder's business, not the integer oracle's
        vocab = O.build_vocabulary(reads)
        ids, off, ps, pe = O.encode_reads(reads, vocab, pos)
        if case.get("raises"):
            with pytest.raises(AssertionError):
                c_oracle.COracleGraph(ids, off, k, ps, pe)
            continue
        for stage, op in (("build", None), ("rlcc5", "rlcc"), ("filter2_2", "filter")):
            c = c_oracle.COracleGraph(ids, off, k, ps, pe)
            p = O.OracleGraph(reads, k, pos)
            if "raises" in case[stage]:
                with pytest.raises(TypeError):
                    c.remove_low_coverage_components(5)
                continue
            if op == "rlcc":
                c.remove_low_coverage_components(5), p.remove_low_coverage_components(5)
            elif op == "filter":
                c.filter_graph(2, 2), p.filter_graph(2, 2)
            assert O.diff_arrays(c.arrays(), p.arrays(vocab)) == [], (case["name"], stage)
