"""Randomised small read sets: the two independent CPU restatements of upstream's algorithm (the Python one,
which hashes with SHA-256 exactly as upstream does, and the plain-C one on SHA ranks) must agree on every
array, through both filters -- tiny vocabularies make tandem repeats, self edges, hairpins, palindromic
windows, multi-edges and short / empty reads all common."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import gmg_oracle as O


def random_reads(rng, n_reads, vocab_size, max_len):
    vocab_names = ["g%d" % i for i in range(vocab_size)]
    reads = {}
    for r in range(n_reads):
        n = int(rng.integers(0, max_len + 1))
        if rng.random() < 0.3 and n >= 2:                     # a read that repeats itself
            half = ["%s%s" % ("+-"[int(rng.integers(2))], vocab_names[int(rng.integers(vocab_size))]) for _ in range(n // 2)]
            calls = (half * 2)[:n]
        else:
            calls = ["%s%s" % ("+-"[int(rng.integers(2))], vocab_names[int(rng.integers(vocab_size))]) for _ in range(n)]
        if rng.random() < 0.2:                                # ... or is the reverse complement of an earlier one
            prev = reads.get("r%03d" % int(rng.integers(0, max(r, 1))))
            if prev:
                calls = [("-" if c[0] == "+" else "+") + c[1:] for c in reversed(prev)]
        reads["r%03d" % r] = calls
    return reads


def apply_both(py, c, op, *args):
    """apply a filter to both oracles; upstream's remove_node raises TypeError on a multi-edge, so must both"""
    errs = []
    for g in (py, c):
        try:
            getattr(g, op)(*args)
            errs.append(None)
        except TypeError:
            errs.append(TypeError)
    assert errs[0] == errs[1], (op, errs)
    return errs[0] is None


@pytest.mark.parametrize("seed", range(60))
def test_python_and_c_oracles_agree_on_random_small_inputs(seed):
    rng = np.random.default_rng(seed)
    reads = random_reads(rng, int(rng.integers(1, 25)), int(rng.integers(2, 9)), int(rng.integers(1, 14)))
    k = int(rng.integers(1, 6))
    vocab = O.build_vocabulary(reads)
    ids, off, _, _ = O.encode_reads(reads, vocab)
    try:
        py = O.OracleGraph(reads, k)
    except AssertionError as e:                               # palindromic even-k window
        assert "identical" in str(e)
        with pytest.raises(AssertionError):
            c_oracle.COracleGraph(ids, off, k)
        return
    c = c_oracle.COracleGraph(ids, off, k)
    assert O.diff_arrays(py.arrays(vocab), c.arrays()) == []
    if not apply_both(py, c, "remove_low_coverage_components", int(rng.integers(1, 5))):
        return
    assert O.diff_arrays(py.arrays(vocab), c.arrays()) == []
    apply_both(py, c, "filter_graph", int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    assert O.diff_arrays(py.arrays(vocab), c.arrays()) == []
