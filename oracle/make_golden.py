"""Generate tests/golden/ by running the UNMODIFIED upstream GeneMerGraph -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs the read-only upstream checkout):

    python oracle/make_golden.py

Writes
  tests/golden/inputs/<name>.npz   integer CSR form of each input (vocabulary names in SHA-rank order,
                                   signed ids, read offsets, optional per-call positions)
  tests/golden/expected.json       per (input, k): summary counts + SHA-256 digests of every graph
                                   array after build, after remove_low_coverage_components(5) and
                                   after a following filter_graph(3, 1)
  tests/golden/small_cases.json    string-level cases with the upstream SHA-256 node / edge keys,
                                   per-read lists, coverages and error behaviour, in full

Sources of the inputs: the upstream test fixtures tests/complex_gene_calls_*.json (+ positions),
the constructor / filter cases of tests/test_gene_mer_graph.py:14-175, 1971-2212, the special
cases listed in SURVEY.md section 8c, and seeded synthetic reads from amira_b200.synth.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from amira_b200 import synth  # noqa: E402
from oracle import gmg_oracle as O  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TESTS = os.path.join(ref_harness.REFERENCE_ROOT, "tests")

FIXTURES = [  # (fixture, ks, keep positions)
    ("one", (3, 5, 7), False), ("three", (3,), False), ("four", (5, 3), True), ("five", (3,), False),
    ("six", (3,), False), ("seven", (3, 1, 2), True), ("eight", (3,), True), ("nine", (3, 5), False),
]

STAGES = (("build", None), ("rlcc5", ("rlcc", 5)), ("rlcc5_filter3_1", ("filter", 3, 1)))


def run_stages(cg, reads, k, positions, vocab):
    """build + staged removals on the upstream class -> {stage: {summary, digest}}, timings"""
    out = {}
    read_ids = list(reads)
    t0 = time.perf_counter()
    try:
        g = cg.GeneMerGraph(reads, k, positions)
    except AssertionError as e:
        return {"build": {"raises": "AssertionError", "message": str(e)}}, 0.0
    dt = time.perf_counter() - t0
    for stage, op in STAGES:
        if op is not None:
            try:
                if op[0] == "rlcc":
                    g.remove_low_coverage_components(op[1])
                else:
                    g.filter_graph(op[1], op[2])
            except TypeError as e:
                out[stage] = {"raises": "TypeError", "message": str(e)}
                break
        a = ref_harness.reference_arrays(g, read_ids, vocab)
        out[stage] = {"summary": O.summary(a), "digest": O.digest_arrays(a)}
    return out, dt


def save_input(name, vocab, ids, off, ps=None, pe=None):
    os.makedirs(os.path.join(GOLD, "inputs"), exist_ok=True)
    small = np.int16 if len(vocab) < 32000 else np.int32
    arrs = {"vocab": np.asarray(vocab), "ids": ids.astype(small), "off": off.astype(np.int64)}
    if ps is not None:
        arrs["pos_start"], arrs["pos_end"] = ps.astype(np.int32), pe.astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "inputs", name + ".npz"), **arrs)


SMALL_CASES = [
    # upstream tests/test_gene_mer_graph.py:38-77, 79-135, 137-175
    ("ctor_two_reads_shared_prefix", {"read1": ["+gene1", "-gene2", "+gene3", "-gene4"],
                                      "read2": ["+gene1", "-gene2", "+gene3", "-gene6"]}, 3, None),
    ("ctor_duplicate_nodes", {"read1": ["+gene1", "-gene2", "+gene3", "-gene4", "+gene1", "-gene2", "+gene3", "+gene8"],
                              "read2": ["+gene1", "-gene2", "+gene3", "-gene6", "+gene1", "-gene2", "+gene3"]}, 3, None),
    ("ctor_one_read_two_genemers", {"read1": ["+gene1", "-gene2", "+gene3", "-gene4"]}, 3, None),
    ("ctor_empty", {}, 0, None),
    ("ctor_empty_k3", {}, 3, None),
    # SURVEY.md 8c special cases
    ("homopolymer_self_loop", {"r1": ["+a", "+a", "+a", "+a", "+a"]}, 3, None),
    ("tandem_repeat_r1_first", {"r1": ["+g1", "+g2", "+g1", "+g2"], "r2": ["-g1", "-g2", "-g1", "-g2"]}, 3, None),
    ("tandem_repeat_r2_first", {"r2": ["-g1", "-g2", "-g1", "-g2"], "r1": ["+g1", "+g2", "+g1", "+g2"]}, 3, None),
    ("palindrome_k2", {"r1": ["+a", "-a", "+b"]}, 2, None),
    ("palindrome_k1_ok", {"r1": ["+a", "-a", "+b"]}, 1, None),
    ("hairpin", {"r1": ["+p", "+x", "+y", "-y", "-x", "-p"]}, 3, None),
    ("homopolymer_filter", {"r1": ["+a", "+a", "+a", "+a"]}, 3, None),
    ("short_and_single_window", {"r1": ["+a", "+b"], "r2": ["+a", "+b", "+c"], "r3": [], "r4": ["-c", "-b", "-a"]}, 3, None),
    ("spaces_in_names", {"r1": ["+gene 1", "-gene_1", "+gene 2", "+x y z"]}, 2, None),
    ("with_positions", {"r1": ["+a", "+b", "+c", "+d"], "r2": ["-d", "-c", "-b"]}, 3,
     {"r1": [[0, 10], [12, 20], [25, 40], [41, 50]], "r2": [[5, 9], [10, 30], [33, 60]]}),
    ("k1_three_reads", {"r1": ["+a", "+b", "+c", "+a"], "r2": ["-a", "+b", "-c"], "r3": ["+c", "+c", "-b"]}, 1, None),
    ("k2_even", {"r1": ["+a", "+b", "+c", "+a", "+b"], "r2": ["-b", "-a", "-c"]}, 2, None),
    ("k4_even", {"r1": ["+a", "+b", "+c", "+d", "+e", "+f"], "r2": ["-f", "-e", "-d", "-c", "-b"]}, 4, None),
    ("bad_strand", {"r1": ["+a", "b", "+c"]}, 3, None),
    ("blank_gene", {"r1": ["+a", " ", "+c"]}, 3, None),
    ("missing_name", {"r1": ["+a", "+", "+c"]}, 3, None),
]


def small_case(cg, name, reads, k, positions):
    rec = {"name": name, "reads": reads, "k": k, "positions": positions}
    try:
        g = cg.GeneMerGraph(reads, k, positions)
    except AssertionError as e:
        rec["raises"] = "AssertionError"
        rec["message"] = str(e)
        return rec

    def snap(g):
        return {
            "node_hashes": [hex(h) for h in g.get_nodes()],
            "node_cov": [n.get_node_coverage() for n in g.get_nodes().values()],
            "node_canonical": [[("+" if x.get_strand() == 1 else "-") + x.get_name() for x in n.get_canonical_geneMer()]
                               for n in g.get_nodes().values()],
            "node_first_dir": [n.get_geneMer().get_geneMerDirection() for n in g.get_nodes().values()],
            "node_reads": [list(n.get_list_of_reads()) for n in g.get_nodes().values()],
            "node_fw": [[hex(h) for h in n.get_forward_edge_hashes()] for n in g.get_nodes().values()],
            "node_bw": [[hex(h) for h in n.get_backward_edge_hashes()] for n in g.get_nodes().values()],
            "node_comp": [n.get_component() for n in g.get_nodes().values()],
            "edge_hashes": [hex(h) for h in g.get_edges()],
            "edge_src": [hex(e.get_sourceNode().__hash__()) for e in g.get_edges().values()],
            "edge_tgt": [hex(e.get_targetNode().__hash__()) for e in g.get_edges().values()],
            "edge_sd": [e.get_sourceNodeDirection() for e in g.get_edges().values()],
            "edge_td": [e.get_targetNodeDirection() for e in g.get_edges().values()],
            "edge_cov": [e.get_edge_coverage() for e in g.get_edges().values()],
            "read_nodes": {r: [None if h is None else hex(h) for h in v] for r, v in g.get_readNodes().items()},
            "read_dirs": {r: list(v) for r, v in g.get_readNodeDirections().items()},
            "read_pos": {r: [None if p is None else list(p) for p in v] for r, v in g.get_readNodePositions().items()},
            "short_reads": dict(g.get_short_read_annotations()),
            "reads_to_correct": sorted(g.get_reads_to_correct()),
            "min_node_cov": g.get_minNodeCoverage(), "min_edge_cov": g.get_minEdgeCoverage(),
        }

    rec["build"] = snap(g)
    g2 = cg.GeneMerGraph(reads, k, positions)
    try:
        g2.remove_low_coverage_components(5)
        rec["rlcc5"] = snap(g2)
    except TypeError as e:
        rec["rlcc5"] = {"raises": "TypeError", "message": str(e)}
    g3 = cg.GeneMerGraph(reads, k, positions)
    g3.filter_graph(2, 2)
    rec["filter2_2"] = snap(g3)
    return rec


def main():
    cg = ref_harness.load()
    os.makedirs(GOLD, exist_ok=True)
    expected = {"_meta": {"python": sys.version.split()[0], "pickle_protocol": __import__("pickle").DEFAULT_PROTOCOL,
                          "generator": "oracle/make_golden.py", "upstream": "Danderson123/Amira v0.11.0"}}
    for fx, ks, keep_pos in FIXTURES:
        reads = json.load(open(os.path.join(REF_TESTS, "complex_gene_calls_%s.json" % fx)))
        ppath = os.path.join(REF_TESTS, "complex_gene_positions_%s.json" % fx)
        positions = json.load(open(ppath)) if (keep_pos and os.path.exists(ppath)) else None
        vocab = O.build_vocabulary(reads)
        ids, off, ps, pe = O.encode_reads(reads, vocab, positions)
        save_input("fixture_" + fx, vocab, ids, off, ps, pe)
        for k in ks:
            res, dt = run_stages(cg, reads, k, positions, vocab)
            W = synth.count_windows(off, k)
            key = "fixture_%s/k%d" % (fx, k)
            expected[key] = {"input": "fixture_" + fx, "k": k, "stages": res,
                             "upstream_build_seconds": round(dt, 3), "windows": W}
            print(key, res.get("build", {}).get("summary"), "%.1fs" % dt, flush=True)
    # seeded synthetic reads through the upstream class
    for name, cfg, n, ks, with_pos in [
        ("synth_c2_2000", synth.replace(synth.CONFIGS["c2"], error_rate=0.05), 2000, (3, 5, 7, 9), True),
        ("synth_c4_3000", synth.CONFIGS["c4"], 3000, (3, 5), False),
        ("synth_c5_2000", synth.CONFIGS["c5"], 2000, (5, 15), False),
    ]:
        ids, off = synth.generate(cfg, 0, n)
        names = synth.vocabulary_names(cfg.vocab)
        used = sorted({abs(int(x)) for x in ids.tolist()})
        # compact the vocabulary to the genes that occur (keeps SHA order, keeps the file small)
        remap = np.zeros(cfg.vocab + 1, np.int32)
        remap[used] = np.arange(1, len(used) + 1)
        ids = (np.sign(ids) * remap[np.abs(ids)]).astype(np.int32)
        vocab = [names[u - 1] for u in used]
        ps = pe = None
        if with_pos:
            ps, pe = synth.positions_for(off, cfg.seed)
        save_input(name, vocab, ids, off, ps, pe)
        reads = synth.to_read_dict(ids, off, vocab)
        positions = None
        if with_pos:
            positions = {r: [[int(ps[j]), int(pe[j])] for j in range(off[i], off[i + 1])]
                         for i, r in enumerate(reads)}
        for k in ks:
            res, dt = run_stages(cg, reads, k, positions, vocab)
            key = "%s/k%d" % (name, k)
            expected[key] = {"input": name, "k": k, "stages": res, "upstream_build_seconds": round(dt, 3),
                             "windows": synth.count_windows(off, k)}
            print(key, res.get("build", {}).get("summary"), "%.1fs" % dt, flush=True)
    with open(os.path.join(GOLD, "expected.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    small = [small_case(cg, *c) for c in SMALL_CASES]
    with open(os.path.join(GOLD, "small_cases.json"), "w") as f:
        json.dump(small, f, indent=1)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
