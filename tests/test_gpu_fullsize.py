"""Full-size parity against the reference itself (oracle/_ref: /root/reference/sprparsimony.cpp compiled in place).

BASELINE.json's configurations at the sizes the bench runs them -- C3 (protein, 500 taxa x 50 000 sites), C5 (32-state
morphology, 300 x 20 000) -- and a 100 000-site slice of C4 (DNA, 1000 taxa): the tree score, the per-pattern vector and
every insertion score of >= 20 sampled node visits (first / middle / last of the sweep; tips and inner nodes, i.e. both
the p side and the q side of rearrangeParsimony, sprparsimony.cpp:2259-2376) must equal what the reference's own
kernels compute on the same alignment and tree; for C3 and C5 also under -cost (Sankoff kernels :477-551, :880-961).
The C port stands in only where oracle/_ref was not built."""
import numpy as np
import pytest

from oracle import portlib, reflib

pytestmark = pytest.mark.gpu


def _case(name, sites=None):
    import bench
    from mpboot_b200 import hostprep, synth
    n, full_sites, dt, mu, seed, tseed = bench.WORKLOADS[name]
    L = full_sites if sites is None else sites
    gen = synth.evolve_alignment if n * L <= (1 << 28) else synth.evolve_alignment_blocked
    chars = gen(n, L, dt, mu, seed)
    prep = hostprep.prepare(chars, dt, compress=False)
    bn, bs = synth.random_tree_rings(n, np.random.default_rng(tseed))
    prep.update(n=n, datatype=dt, bn=bn, bs=bs)
    return prep


def _reference(c):
    if reflib.available():
        return reflib.RefEngine(c["chars"], c["weights"], c["datatype"], n_informative=c["n_inf"])
    return portlib.OracleEngine(c["codes"], c["weights"], c["datatype"])


def _sampled_visits(n, k=7):
    """>= 3k visits of the sweep order[1 .. 2n-2]: the first, middle and last k (tips come first in the order, the inner
    nodes in DFS order after them, so both kinds are covered)."""
    nv = 2 * n - 2
    mid = nv // 2
    return sorted(set(list(range(1, k + 1)) + list(range(mid - k // 2, mid + k - k // 2)) + list(range(nv - k + 1, nv + 1))))


def _engine(c):
    from mpboot_b200.engine import Engine
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], c["datatype"])
    return eng


def _compare(c, eng, ref, maxtrav=6):
    n = c["n"]
    eng.set_tree(c["bn"], c["bs"])
    ref.set_ring(c["bn"], c["bs"])
    ref.allocate(per_site=True)
    s_ref = ref.evaluate_full(per_site=True)
    assert eng.tree_score() == s_ref                                         # R3 + R4
    pp_ref, sum_ref = ref.pattern_parsimony(c["n_inf"])
    pp, sm = eng.pattern_parsimony()                                         # R5
    assert sm == sum_ref and np.array_equal(pp[: c["n_inf"]], pp_ref)
    order = eng.visit_order()
    visits = _sampled_visits(n)
    assert len(visits) >= 20
    checked = 0
    for i in visits:                                                         # R6: every testInsertParsimony score of the visit
        vb, mp, _, _ = eng.scan_visits(order, i, 1, 1, maxtrav)
        ref.record(False)
        ref.rearrange(i, 1, maxtrav, True, s_ref)
        want = ref.saved()[1:]
        assert np.array_equal(want, mp.astype(np.int32)), "visit %d" % i
        checked += len(want)
    assert checked > 20 * 8
    return s_ref


@pytest.mark.parametrize("name,sites", [("c3", None), ("c5", None), ("c4", 100000)], ids=["c3-full", "c5-full", "c4-100k-slice"])
def test_full_size_fitch_vs_reference(name, sites):
    c = _case(name, sites)
    eng = _engine(c)
    _compare(c, eng, _reference(c))


@pytest.mark.parametrize("name", ["c3", "c5"], ids=["c3-full-cost", "c5-full-cost"])
def test_full_size_sankoff_vs_reference(name):
    c = _case(name)
    S = {2: 20, 6: 32}[c["datatype"]]
    rng = np.random.default_rng(17)
    cost = rng.integers(1, 4, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
    cost = cost.astype(np.uint32)
    # segments as IQTree::doSegmenting draws them from the scores of a tree (iqtree.cpp:3793): here from the Fitch scores of
    # the test tree, which only have to give bounds at multiples of 16 that keep the 16-bit segment sums meaningful
    eng = _engine(c)
    eng.set_tree(c["bn"], c["bs"])
    pp, _ = eng.pattern_parsimony()
    import bench
    seg = bench.do_segmenting(3 * pp[: c["n_inf"]].astype(np.int64), c["weights"], c["n_inf"])
    ref = _reference(c)
    hi = ref.set_cost_matrix(cost, seg)
    assert eng.set_cost_matrix(cost, seg) == hi
    _compare(c, eng, ref)
