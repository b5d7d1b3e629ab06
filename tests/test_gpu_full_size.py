"""Full-size runs of the BASELINE.json workloads on the GPU: size-independent properties, and bit-exact parity
against the C oracle through committed digests of its builds of the same inputs (the oracle itself would take too
long here; live oracle runs at oracle-sized inputs are in test_gpu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("c5", 1_250_000, 5),      # one rank's shard of the 10M-read set (the bench workload)
         ("c3", 500_000, 3),        # high-error reads
         ("c3", 500_000, 7)]        # ... with 32-byte node slots (7 x 14 bits > 85)


def canonical_windows(ids, off, k, sample):
    """canonical key and direction of the windows starting at call positions `sample` (numpy restatement of
    construct_gene_mer.py:4-39 on signed ranks)"""
    w = ids[sample[:, None] + np.arange(k)[None, :]].astype(np.int64)
    rc = -w[:, ::-1]
    diff = w != rc
    first = diff.argmax(axis=1)
    rows = np.arange(len(sample))
    fwd = w[rows, first] < rc[rows, first]
    return np.where(fwd[:, None], w, rc), np.where(fwd, 1, -1)


@pytest.mark.parametrize("cfg_name,n_reads,k", CASES)
def test_full_size_invariants(cfg_name, n_reads, k):
    from amira_b200 import synth
    from amira_b200.device_graph import DeviceGraph
    from oracle import gmg_oracle as O
    ids, off = synth.generate(synth.CONFIGS[cfg_name], 0, n_reads)
    L = np.diff(off)
    nwin = np.maximum(L - (k - 1), 0)
    W, pairs = int(nwin.sum()), int(np.maximum(nwin - 1, 0).sum())
    dg = DeviceGraph(0)
    dg.build(ids, off, k)
    a = dg.arrays()
    N, E = len(a["node_cov"]), len(a["edge_cov"])
    # counting identities: one coverage unit per window; one forward + one reverse unit per adjacent pair
    assert len(a["win_node"]) == W and int(a["node_cov"].sum(dtype=np.int64)) == W
    assert int(a["edge_cov"].sum(dtype=np.int64)) == 2 * pairs
    assert np.array_equal(np.diff(a["win_off"]), nwin) and np.array_equal(a["is_short"].astype(bool), nwin == 0)
    assert a["win_node"].min() >= 0 and a["win_node"].max() == N - 1
    assert np.array_equal(np.bincount(a["win_node"], minlength=N), a["node_cov"])
    # insertion order: node i's first window comes before node i+1's (dict order of upstream's _nodes)
    first_win = np.full(N, W, np.int64)
    np.minimum.at(first_win, a["win_node"], np.arange(W))
    assert np.all(np.diff(first_win) > 0)
    assert np.array_equal(a["node_dir"], a["win_dir"][first_win])
    # canonicalisation on a sample of windows: exported key of the window's node, and its direction
    rng = np.random.default_rng(0)
    ws = np.sort(rng.choice(W, 200_000, replace=False))
    read_of = np.searchsorted(a["win_off"], ws, side="right") - 1
    pos = off[read_of] + (ws - a["win_off"][read_of])
    key, d = canonical_windows(ids, off, k, pos)
    assert np.array_equal(a["node_key"][a["win_node"][ws]], key) and np.array_equal(a["win_dir"][ws], d)
    # node -> reads: ascending, unique, and exactly the reads whose windows hit the node
    ro, rd = a["node_reads_off"], a["node_reads"]
    seg = np.repeat(np.arange(N), np.diff(ro))
    same = seg[1:] == seg[:-1]
    assert np.all(rd[1:][same] > rd[:-1][same])
    win_read = np.repeat(np.arange(len(L)), nwin)
    pairs_nr = np.unique(a["win_node"].astype(np.int64) * len(L) + win_read)
    assert len(pairs_nr) == len(rd) and np.array_equal(pairs_nr, seg.astype(np.int64) * len(L) + rd)
    # edges: forward / reverse come in pairs with equal coverage; adjacency lists partition the edges
    src, tgt, sd, td = a["edge_src"], a["edge_tgt"], a["edge_sd"], a["edge_td"]
    non_self = np.flatnonzero(src != tgt)
    assert len(non_self) % 2 == 0
    f, r = non_self[0::2], non_self[1::2]
    assert np.array_equal(src[f], tgt[r]) and np.array_equal(tgt[f], src[r])
    assert np.array_equal(sd[f], -td[r]) and np.array_equal(td[f], -sd[r]) and np.array_equal(a["edge_cov"][f], a["edge_cov"][r])
    assert len(a["fw_edges"]) + len(a["bw_edges"]) == E
    assert np.all(sd[a["fw_edges"]] == 1) and np.all(sd[a["bw_edges"]] == -1)
    assert np.array_equal(src[a["fw_edges"]], np.repeat(np.arange(N), np.diff(a["fw_off"])))
    # components: every edge inside one component; ids numbered by first node
    comp = a["node_comp"].astype(np.int64)
    assert np.array_equal(comp[src], comp[tgt])
    first_node = np.full(comp.max() + 1, N, np.int64)
    np.minimum.at(first_node, comp, np.arange(N))
    assert np.all(np.diff(first_node[1:]) > 0) and first_node[1] == 0
    # idempotence: the same input builds the same arrays again (tables sized differently the second time)
    digest = O.digest_arrays(a)
    dg.build(ids, off, k)
    assert O.digest_arrays(dg.arrays()) == digest
    # filters: thresholds hold afterwards, masked windows are exactly those of removed nodes
    cov_before, win_before = a["node_cov"], a["win_node"]
    dg.remove_low_coverage_components(5)
    dg.filter_graph(3, 1)
    b = dg.arrays()
    assert b["node_cov"].min() >= 3 and b["edge_cov"].min() >= 1
    gone = b["win_node"] < 0
    comp_max = np.zeros(comp.max() + 1, np.int64)
    np.maximum.at(comp_max, comp, cov_before)
    removed = (cov_before < 3) | (comp_max[comp] < 5)
    assert np.array_equal(gone, removed[win_before])
    assert np.array_equal(b["node_cov"], cov_before[~removed])
    assert np.array_equal(np.flatnonzero(b["to_correct"]), np.unique(win_read[gone]))
    dg.close()


def _golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_digests.json")) as f:
        return json.load(f)


def test_full_size_c5_shard_bit_exact_against_oracle_digests():
    """the whole C5 shard (1.25M reads, 32.4M gene-mers, k=5): every exported array, digested, against the digests of
    the C oracle's build of the same input (tests/golden/bench_digests.json, scripts/make_bench_digests.py)"""
    from amira_b200 import synth
    from amira_b200.device_graph import DeviceGraph
    from oracle import digests
    from dataclasses import replace
    gold = _golden()["c5_n1"]
    cfg = replace(synth.CONFIGS["c5"], n_reads=gold["reads"])   # as scripts/make_bench_digests.py and bench.py do
    ids, off = synth.generate(cfg, 0, gold["reads"])
    dg = DeviceGraph(0)
    for _ in range(2):                      # cold tables, then tables sized from the first build: same graph
        dg.build(ids, off, gold["k"])
        a = dg.arrays()
        assert digests.diff_digests(digests.global_digest(a), gold["global"]) == []
        assert digests.diff_digests(digests.rank_digest(a), gold["ranks"][0]) == []
    dg.close()


@pytest.mark.parametrize("k", [3, 5, 7])
def test_full_size_c3_through_the_filters_bit_exact_against_oracle_digests(k):
    """C3 (500k reads, 10 % bad calls): build, remove_low_coverage_components(5), filter_graph(3, 1) -- every stage
    against the C oracle's digests"""
    from amira_b200 import synth
    from amira_b200.device_graph import DeviceGraph
    from oracle import digests
    gold = _golden()["c3_k%d" % k]
    ids, off = synth.generate(synth.CONFIGS["c3"], 0, gold["reads"])
    dg = DeviceGraph(0)
    dg.build(ids, off, k)
    stages = (("build", lambda: None), ("rlcc5", lambda: dg.remove_low_coverage_components(5)),
              ("rlcc5_filter3_1", lambda: dg.filter_graph(3, 1)))
    for stage, step in stages:
        step()
        a = dg.arrays()
        want = gold["stages"][stage]
        assert digests.diff_digests(digests.global_digest(a), want["global"]) == [], stage
        assert digests.diff_digests(digests.rank_digest(a), want["rank"]) == [], stage
    dg.close()
