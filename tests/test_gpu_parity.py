"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the
committed golden vectors (produced by the reference itself) and against the C oracle on
further seeded inputs.  Bit-exact: everything on this path is integer/bit work."""
import glob
import os

import numpy as np
import pytest

from oracle import portlib
from tests.helpers import make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
                    if not f.endswith("tables.npz") and not os.path.basename(f).startswith(("sankoff_", "mulhits", "skbb")))
IDS = [os.path.basename(p)[:-4] for p in CASE_FILES]


def _engine(codes, weights, dt, bn=None, bs=None):
    from mpboot_b200.engine import Engine
    eng = Engine()
    eng.load_alignment(codes, weights, dt)
    if bn is not None:
        eng.set_tree(bn, bs)
    return eng


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_golden_planes_scores_patterns(path):
    g = dict(np.load(path))
    n, dt = int(g["n"]), int(g["datatype"])
    eng = _engine(g["codes"], g["weights"], dt, g["bn"], g["bs"])
    assert eng.ref_words == int(g["W"]) and eng.n_inf == int(g["n_inf"])
    for t in range(1, n + 1):
        assert np.array_equal(eng.tip_planes(t), g["tip_planes"][t - 1])       # R1
    assert eng.tree_score() == int(g["score"])                                 # R3 + R4
    pp, sm = eng.pattern_parsimony()                                           # R5
    assert sm == int(g["ptn_sum"]) and np.array_equal(pp[: int(g["n_inf"])], g["ptn_pars"])
    assert np.array_equal(eng.visit_order()[1:], g["order"][1:])               # nodeRectifierPars


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_golden_insertion_scores(path):
    g = dict(np.load(path))
    n, dt, mt = int(g["n"]), int(g["datatype"]), int(g["maxtrav"])
    eng = _engine(g["codes"], g["weights"], dt, g["bn"], g["bs"])
    vb, mp, cref, cprune = eng.scan_visits(g["order"], 1, 2 * n - 2, 1, mt)
    assert np.array_equal(vb, g["visit_begin"])
    assert np.array_equal(mp.astype(np.int32), g["visit_mp"])                  # every testInsertParsimony score
    # batching must not matter: visit-by-visit gives the same numbers
    for first in (1, n, 2 * n - 2):
        vb1, mp1, _, _ = eng.scan_visits(g["order"], first, 1, 1, mt)
        assert np.array_equal(mp1.astype(np.int32), g["visit_mp"][vb[first - 1]: vb[first]])


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_golden_spr_search_same_moves(path):
    g = dict(np.load(path))
    dt, mt = int(g["datatype"]), int(g["maxtrav"])
    eng = _engine(g["codes"], g["weights"], dt)
    portlib.seed_rng(2024)
    ret, bn, bs, nins = eng.optimize_spr(g["bn"], g["bs"], portlib.rng_fn_address(), 1, mt)
    assert ret == int(g["opt_plain_ret"])
    assert portlib.rng_draws() == int(g["opt_plain_draws"])                    # same RNG draws, same order
    assert np.array_equal(bn[3:], g["opt_plain_bn"][3:]) and np.array_equal(bs[3:], g["opt_plain_bs"][3:])
    eng.set_tree(bn, bs)
    assert eng.tree_score() == int(g["opt_plain_score"])


@pytest.mark.parametrize("n,L,dt,seed", [(64, 6000, 1, 41), (48, 1500, 2, 42), (33, 700, 6, 43), (21, 500, 0, 44),
                                          (100, 5000, 1, 45)])
def test_oracle_seeded_cases(n, L, dt, seed):
    c = make_case(n, L, dt, seed)
    eng = _engine(c["codes"], c["weights"], dt, c["bn"], c["bs"])
    o = portlib.OracleEngine(c["codes"], c["weights"], dt)
    o.set_ring(c["bn"], c["bs"])
    assert o.allocate(True) == eng.ref_words
    s0 = o.evaluate_full(True)
    assert eng.tree_score() == s0
    a, sa = o.pattern_parsimony(c["n_inf"]); b, sb = eng.pattern_parsimony()
    assert sa == sb == s0 and np.array_equal(a, b[: c["n_inf"]])
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 6)
    portlib.seed_rng(1)
    for i in range(1, 2 * n - 1):
        o.record(False)
        o.rearrange(i, 1, 6, True, s0)
        assert np.array_equal(o.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), "visit %d" % i
    # view lengths = tr->parsimonyScore of the oriented node
    for node in (n + 1, n + 2, 2 * n - 2):
        for slot in range(3):
            assert eng.view_length(node, slot) >= 0
    portlib.seed_rng(7)
    o.set_ring(c["bn"], c["bs"])
    want = o.optimize_spr(1, 6, bb=False)
    draws = portlib.rng_draws()
    want_ring = o.get_ring()
    portlib.seed_rng(7)
    ret, bn, bs, nins = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 6)
    assert ret == want and portlib.rng_draws() == draws
    assert np.array_equal(bn[3:], want_ring[0][3:]) and np.array_equal(bs[3:], want_ring[1][3:])


def test_edge_cases_weights_and_reweighting():
    """Ragged inputs: zero-weight patterns, a heavy pattern, all-uninformative tail, then the
    ratchet-style re-weighting over resident codes (mpgpu_set_weights)."""
    c = make_case(16, 400, 1, 51)
    w = c["weights"].copy()
    w[0] = 977; w[1] = 0; w[5] = 33
    eng = _engine(c["codes"], w, 1, c["bn"], c["bs"])
    o = portlib.OracleEngine(c["codes"], w, 1); o.set_ring(c["bn"], c["bs"])
    assert o.allocate(True) == eng.ref_words
    for t in (1, 7, 16):
        assert np.array_equal(o.parsvect(t), eng.tip_planes(t))
    s0 = o.evaluate_full(True)
    assert eng.tree_score() == s0
    a, sa = o.pattern_parsimony(c["n_inf"]); b, sb = eng.pattern_parsimony()
    assert sa == sb and np.array_equal(a, b[: c["n_inf"]])
    rng = np.random.default_rng(3)
    w2 = c["weights"] + (rng.random(len(w)) < 0.5)
    eng.set_weights(w2)
    o.set_weights(w2); o.set_ring(c["bn"], c["bs"]); assert o.allocate(True) == eng.ref_words
    assert eng.tree_score() == o.evaluate_full(True)
    a, sa = o.pattern_parsimony(c["n_inf"]); b, sb = eng.pattern_parsimony()
    assert sa == sb and np.array_equal(a, b[: c["n_inf"]])


def test_minimum_tree_and_small_radius():
    c = make_case(4, 120, 1, 61)
    eng = _engine(c["codes"], c["weights"], 1, c["bn"], c["bs"])
    o = portlib.OracleEngine(c["codes"], c["weights"], 1); o.set_ring(c["bn"], c["bs"]); o.allocate(True)
    s0 = o.evaluate_full(True)
    assert eng.tree_score() == s0
    for mt in (1, 2, 6):
        vb, mp, _, _ = eng.scan_visits(eng.visit_order(), 1, 6, 1, mt)
        got = []
        for i in range(1, 7):
            o.record(False); o.rearrange(i, 1, mt, True, s0); got.append(o.saved()[1:])
        assert np.array_equal(np.concatenate(got), mp.astype(np.int32))


def test_full_size_properties_c2_slice():
    """At a BASELINE-sized width (200 taxa x 100k sites is ~3128 words; here 200 x 20k to keep the
    oracle quick) use size-independent properties: score is invariant under the edge it is
    evaluated on, equals the sum of pattern scores x weights, and every insertion score is
    >= the pruned tree's length + the subtree's length."""
    c = make_case(200, 20000, 1, 2, amb=0.001)
    eng = _engine(c["codes"], c["weights"], 1, c["bn"], c["bs"])
    s = eng.tree_score()
    pp, sm = eng.pattern_parsimony()
    assert sm == s
    n = 200
    for node, slot in ((1, 0), (17, 0), (n + 5, 1), (2 * n - 2, 2)):
        bnode = int(c["bn"][3 * node + slot]); bslot = int(c["bs"][3 * node + slot])
        tot = eng.edge_mismatch_partial(node, slot) + eng.view_length(node, slot) + eng.view_length(bnode, bslot)
        assert tot == s
    o = portlib.OracleEngine(c["codes"], c["weights"], 1); o.set_ring(c["bn"], c["bs"]); o.allocate(False)
    assert o.evaluate_full(False) == s
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 6)
    assert len(mp) > 10000 and mp.min() >= 0.9 * s
    # spot-check three visits against the oracle (needs per-site mode for the recorder)
    o2 = portlib.OracleEngine(c["codes"], c["weights"], 1); o2.set_ring(c["bn"], c["bs"]); o2.allocate(True)
    s2 = o2.evaluate_full(True)
    for i in (2, 250, 397):
        o2.record(False); o2.rearrange(i, 1, 6, True, s2)
        assert np.array_equal(o2.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32))


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_golden_stepwise_addition(path):
    """R7: _pllComputeRandomizedStepwiseAdditionParsimonyTree -- same taxon order (PLL's randum), same
    insertion branches, same tie-break draws, same SPR rounds, same tree in the reference's numbering."""
    g = dict(np.load(path))
    dt, mt = int(g["datatype"]), int(g["maxtrav"])
    eng = _engine(g["codes"], g["weights"], dt)
    portlib.seed_rng(77)
    ret, bn, bs, nins, seed_after = eng.stepwise_addition(int(g["ras_seed"]), mt, portlib.rng_fn_address())
    assert ret == int(g["ras_ret"]) and portlib.rng_draws() == int(g["ras_draws"])
    assert np.array_equal(bn[3:], g["ras_bn"][3:]) and np.array_equal(bs[3:], g["ras_bs"][3:])
    eng.set_tree(bn, bs)
    assert eng.tree_score() == ret and nins > 0


@pytest.mark.parametrize("n,L,dt,seed", [(50, 3000, 1, 71), (30, 600, 2, 72), (18, 400, 6, 73), (5, 200, 1, 74), (4, 100, 1, 75)])
def test_stepwise_addition_matches_oracle(n, L, dt, seed):
    c = make_case(n, L, dt, seed)
    o = portlib.OracleEngine(c["codes"], c["weights"], dt)
    for ras_seed in (12345, 987):
        portlib.seed_rng(5)
        want = o.ras(ras_seed, 6)
        draws = portlib.rng_draws()
        wbn, wbs = o.get_ring()
        eng = _engine(c["codes"], c["weights"], dt)
        portlib.seed_rng(5)
        ret, bn, bs, nins, _ = eng.stepwise_addition(ras_seed, 6, portlib.rng_fn_address())
        assert ret == want and portlib.rng_draws() == draws
        assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])


def test_replicate_reweighting_search_matches_oracle():
    """R12 building block (IQTree::optimizeBootTrees, iqtree.cpp:2475-2915): a bootstrap replicate is the
    same resident codes under new pattern frequencies -- mpgpu_set_weights + mpgpu_optimize_spr must
    give what the reference gets after re-creating its data structures for the re-weighted alignment
    (zero-weight patterns included)."""
    c = make_case(36, 2500, 1, 81)
    eng = _engine(c["codes"], c["weights"], 1)
    o = portlib.OracleEngine(c["codes"], c["weights"], 1)
    rng = np.random.default_rng(8)
    L = int(c["weights"].sum())
    for rep in range(3):
        w = rng.multinomial(L, c["weights"] / L).astype(np.int32)
        o.set_weights(w); o.set_ring(c["bn"], c["bs"]); o.allocate(False)
        portlib.seed_rng(100 + rep)
        want = o.optimize_spr(1, 6, bb=False)
        draws = portlib.rng_draws()
        wring = o.get_ring()
        eng.set_weights(w)
        portlib.seed_rng(100 + rep)
        ret, bn, bs, nins = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 6)
        assert ret == want and portlib.rng_draws() == draws
        assert np.array_equal(bn[3:], wring[0][3:]) and np.array_equal(bs[3:], wring[1][3:])


@pytest.mark.parametrize("n,L,dt,seed", [(60, 5000, 1, 51), (40, 1200, 2, 52), (30, 700, 6, 53), (24, 600, 0, 54)])
def test_incremental_view_update_equals_full_recompute(n, L, dt, seed):
    """After every applied move the search recomputes only the stale directed views (k_fitch_wave, one
    launch).  What it leaves in the context must be, word for word and length for length, what a fresh
    mpgpu_set_tree of the final tree computes (k_fitch_level, every view)."""
    c = make_case(n, L, dt, seed)
    eng = _engine(c["codes"], c["weights"], dt)
    portlib.seed_rng(seed)
    ret, bn, bs, nins = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 6)
    refs = [(node, slot) for node in range(n + 1, 2 * n - 1) for slot in range(3)]
    inc_len = [eng.view_length(node, slot) for node, slot in refs]
    inc_planes = [eng.view_planes(node, slot) for node, slot in refs]
    inc_score = eng.tree_score()
    eng.set_tree(bn, bs)
    assert eng.tree_score() == inc_score == ret
    assert inc_len == [eng.view_length(node, slot) for node, slot in refs]
    for (node, slot), p in zip(refs, inc_planes):
        assert np.array_equal(p, eng.view_planes(node, slot)), (node, slot)


@pytest.mark.parametrize("n,L,dt,seed", [(36, 2500, 1, 81), (28, 900, 2, 82)])
def test_refine_replicates_matches_sequential_reference_loop(n, L, dt, seed):
    """N1 (IQTree::optimizeBootTrees default policy, iqtree.cpp:2795-2862): B replicates refined in one
    call over the resident codes == the reference's sequential loop (re-weight, re-allocate, hill-climb),
    one RNG stream running through all replicates."""
    from tests.helpers import make_boot
    c = make_case(n, L, dt, seed)
    B = 5
    boot = make_boot(c, B, seed)
    boot[1, :7] = 0                                           # zero-frequency patterns drop out of the planes
    o = portlib.OracleEngine(c["codes"], c["weights"], dt)
    o.set_ring(c["bn"], c["bs"]); o.allocate(False)
    portlib.seed_rng(5)
    o.optimize_spr(1, 6, bb=False)
    start = o.get_ring()                                      # the replicates' trees: one good tree and the raw one
    trees_bn = np.stack([start[0] if b % 2 == 0 else c["bn"] for b in range(B)]).astype(np.int32)
    trees_bs = np.stack([start[1] if b % 2 == 0 else c["bs"] for b in range(B)]).astype(np.int32)
    portlib.seed_rng(77)
    want_scores, want_rings = [], []
    for b in range(B):
        o.set_weights(boot[b].astype(np.int32)); o.set_ring(trees_bn[b], trees_bs[b]); o.allocate(False)
        want_scores.append(o.optimize_spr(1, 6, bb=False))
        want_rings.append(o.get_ring())
    draws = portlib.rng_draws()
    eng = _engine(c["codes"], c["weights"], dt, c["bn"], c["bs"])
    s0 = eng.tree_score()
    portlib.seed_rng(77)
    scores, tbn, tbs, nins = eng.refine_replicates(boot, trees_bn, trees_bs, portlib.rng_fn_address(), 1, 6)
    assert list(scores) == want_scores and portlib.rng_draws() == draws and nins > 0
    for b in range(B):
        assert np.array_equal(tbn[b][3:], want_rings[b][0][3:]) and np.array_equal(tbs[b][3:], want_rings[b][1][3:]), b
    eng.set_tree(c["bn"], c["bs"])                            # original frequencies are back
    assert eng.tree_score() == s0


def _c2_parts():
    """BASELINE's C2 shape (200 taxa x 100 000 sites, every site its own pattern, the bench's alignment) and its two column
    halves; the cut sits where the number of informative sites before it is a multiple of 16."""
    import bench
    from mpboot_b200 import hostprep, synth
    n, sites, dt, mu, seed, tseed = bench.WORKLOADS["c2"]
    chars = synth.evolve_alignment(n, sites, dt, mu, seed)
    full = hostprep.prepare(chars, dt, compress=False)
    inf = hostprep.informative_mask(__import__("mpboot_b200.encoding", fromlist=["encode"]).encode(chars, dt), dt)
    csum = np.cumsum(inf)
    cut = int(np.nonzero((csum % 16 == 0) & (np.arange(sites) >= sites // 2))[0][0]) + 1
    a = hostprep.prepare(chars[:, :cut], dt, compress=False)
    b = hostprep.prepare(chars[:, cut:], dt, compress=False)
    bn, bs = synth.random_tree_rings(n, np.random.default_rng(tseed))
    return n, dt, full, a, b, bn, bs


def test_full_size_c2_additivity_over_sites():
    """Full BASELINE size (C2: 200 x 100 000).  Parsimony is a sum over sites, so every number the path produces on the
    whole alignment must equal the sum of the numbers on its two column halves: tree score, every view length, every
    insertion score of a whole sweep (14 476 of them) and the per-pattern vector's total."""
    n, dt, full, a, b, bn, bs = _c2_parts()
    engs = [_engine(p["codes"], p["weights"], dt, bn, bs) for p in (full, a, b)]
    assert engs[0].n_inf == engs[1].n_inf + engs[2].n_inf
    s = [e.tree_score() for e in engs]
    assert s[0] == s[1] + s[2]
    for node, slot in ((n + 1, 0), (n + 77, 2), (2 * n - 2, 1)):
        assert engs[0].view_length(node, slot) == engs[1].view_length(node, slot) + engs[2].view_length(node, slot)
    order = engs[0].visit_order()
    res = [e.scan_visits(order, 1, 2 * n - 2, 1, 6) for e in engs]
    assert len(res[0][1]) == 14476
    for k in (0, 2, 3):
        assert np.array_equal(res[0][k], res[1][k]) and np.array_equal(res[0][k], res[2][k])      # same candidates
    assert np.array_equal(res[0][1].astype(np.int64), res[1][1].astype(np.int64) + res[2][1].astype(np.int64))
    pp = [e.pattern_parsimony() for e in engs]
    assert pp[0][1] == pp[1][1] + pp[2][1] == s[0]
    assert np.array_equal(pp[0][0][: engs[0].n_inf], np.concatenate([pp[1][0][: engs[1].n_inf], pp[2][0][: engs[2].n_inf]]))
