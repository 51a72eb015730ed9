"""N2: mpgpu_split_table against the restatement of MTreeSet::convertSplits (oracle/splits_oracle.py):
distinct splits in first-seen order, summed weights, the row of every emitted split."""
import random

import numpy as np
import pytest

from oracle import splits_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ntaxa,ntrees,seed", [(12, 5, 1), (33, 40, 2), (64, 60, 3), (100, 200, 4), (257, 30, 5), (4, 3, 6)])
def test_split_table_matches_restatement(ntaxa, ntrees, seed):
    from mpboot_b200.engine import Engine
    rng = random.Random(seed)
    base = [so.random_tree(range(ntaxa), rng) for _ in range(max(2, ntrees // 3))]
    trees, weights = [], []
    for i in range(ntrees):
        root, sub = base[rng.randrange(len(base))] if rng.random() < 0.6 else so.random_tree(range(ntaxa), rng)
        trees.append(so.tokens_of(sub))
        weights.append(rng.randrange(1, 9))
    weights[-1] = 0                                           # a tree whose supports are read, not counted
    order, wsum, emit = so.split_table(ntaxa, trees, weights)
    tokens = np.concatenate([np.asarray(t, dtype=np.int32) for t in trees])
    begin = np.concatenate([[0], np.cumsum([len(t) for t in trees])]).astype(np.int64)
    eng = Engine()
    bits, wgt, eu = eng.split_table(ntaxa, tokens, begin, np.asarray(weights, dtype=np.int32))
    assert len(wgt) == len(order)
    assert np.array_equal(wgt, np.asarray(wsum, dtype=np.int32))
    assert np.array_equal(eu, np.asarray(emit, dtype=np.int32))
    want = np.asarray([so.bits_of(s, ntaxa) for s in order], dtype=np.uint32)
    assert np.array_equal(bits, want)
    # every tree has 2n-3 edges; the trivial splits all have the total weight
    tot = sum(weights)
    for row, s in enumerate(order):
        if len(s) == 1:
            assert wgt[row] == tot


def test_split_table_rejects_malformed_streams():
    from mpboot_b200.engine import Engine, MpGpuError
    eng = Engine()
    with pytest.raises(MpGpuError):
        eng.split_table(8, np.array([0, 1, -3], dtype=np.int32), np.array([0, 3], dtype=np.int64), np.array([1], dtype=np.int32))
    with pytest.raises(MpGpuError):
        eng.split_table(8, np.array([0, 9, -2], dtype=np.int32), np.array([0, 3], dtype=np.int64), np.array([1], dtype=np.int32))
