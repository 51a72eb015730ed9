"""GPU parity, -cost (Sankoff weighted parsimony): the CUDA path through the C-ABI against
the golden vectors tools/make_golden_sankoff.py produced by running the reference itself (tip vectors, node
vectors, node scores, remainder bounds, every insertion score of a sweep, the decisions of the same sweep in
plain mode with the lower-bound early exit, whole searches in both modes, RAS trees) and against the C oracle
on further seeded cases.  Bit-exact."""
import glob
import os

import numpy as np
import pytest

from oracle import portlib
from tests.helpers import make_case

pytestmark = pytest.mark.gpu


def _reference(chars, weights, dt, n_inf, codes):
    from oracle import portlib, reflib
    if reflib.available():
        return reflib.RefEngine(chars, weights, dt, n_informative=n_inf)
    return portlib.OracleEngine(codes, weights, dt)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "sankoff_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FILES]


def _engine(g, tree=True):
    from mpboot_b200.engine import Engine
    eng = Engine()
    eng.load_alignment(g["codes"], g["weights"], int(g["datatype"]))
    if tree:
        eng.set_tree(g["bn"], g["bs"])
    assert eng.set_cost_matrix(g["cost"], g["seg"]) == int(g["highest"])
    return eng


def _replay_visit(mp, est, cref, cprune, best_in, early):
    """testInsertParsimony's bookkeeping (:2168-2176) over one visit, with the reference's early exit (:951-956)"""
    best, hits, ins, rem = int(best_in), 1, 0, 0
    for j in range(len(mp)):
        if early and int(est[j]) > best:
            continue
        m = int(mp[j])
        if m < best:
            hits = 1
        elif m == best:
            hits += 1
        if m < best or (m == best and portlib.lib().mporacle_random_double(None) <= 1.0 / hits):
            best, ins, rem = m, int(cref[j]), int(cprune[j])
    return np.array([best, rem // 3, rem % 3 if rem else 0, ins // 3, ins % 3 if ins else 0, hits], dtype=np.uint32)


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden_vectors_scores_bounds(path):
    g = dict(np.load(path))
    n, ninf = int(g["n"]), int(g["n_inf"])
    eng = _engine(g)
    L, lb = eng.sankoff_layout()
    assert L == int(g["L"]) and np.array_equal(lb, g["lower_bounds"])
    for t in range(1, n + 1):
        assert np.array_equal(eng.sankoff_view(t), g["tips"][t - 1])
    assert eng.tree_score() == int(g["score"])
    order = eng.visit_order()
    assert np.array_equal(order[1:], g["order"][1:])
    for k in range(n + 1, 2 * n - 1):                    # the reference's vectors face tr->start after evaluate(start, full)
        node, slot = int(order[k]) // 3, int(order[k]) % 3
        assert np.array_equal(eng.sankoff_view(node, slot), g["node_vect"][node - n - 1]), node
        assert eng.view_length(node, slot) == int(g["node_score"][node - n - 1])
    pp, sm = eng.pattern_parsimony()
    assert sm == int(g["ptn_sum"]) and np.array_equal(pp[:ninf], g["ptn_pars"][:ninf])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden_insertion_scores_and_early_exit(path):
    g = dict(np.load(path))
    n, mt = int(g["n"]), int(g["maxtrav"])
    eng = _engine(g)
    vb, mp, cref, cprune = eng.scan_visits(g["order"], 1, 2 * n - 2, 1, mt)
    est = eng.scan_bounds(len(mp))
    assert np.array_equal(vb, g["visit_begin"])
    assert np.array_equal(mp.astype(np.int32), g["visit_mp"])                  # exact score of every insertion
    for early, key, draws in ((True, "plain_visit_out", int(g["plain_draws"])), (False, "visit_out", None)):
        portlib.seed_rng(31337)
        for i in range(1, 2 * n - 1):
            a, b = vb[i - 1], vb[i]
            out = _replay_visit(mp[a:b], est[a:b], cref[a:b], cprune[a:b], int(g["score"]), early and len(g["seg"]) > 1)
            assert np.array_equal(out, g[key][i - 1]), (key, i)
        if draws is not None:
            assert portlib.rng_draws() == draws
    for first in (1, n, 2 * n - 2):                                            # batching must not matter
        vb1, mp1, _, _ = eng.scan_visits(g["order"], first, 1, 1, mt)
        assert np.array_equal(mp1.astype(np.int32), g["visit_mp"][vb[first - 1]: vb[first]])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden_searches_and_ras(path):
    g = dict(np.load(path))
    mt = int(g["maxtrav"])
    eng = _engine(g, tree=False)
    for tag, exact in (("plain", 0), ("bb", 1)):
        eng.set_option("sankoff_exact", exact)
        portlib.seed_rng(2024)
        ret, bn, bs, nins = eng.optimize_spr(g["bn"], g["bs"], portlib.rng_fn_address(), 1, mt)
        assert ret == int(g["opt_%s_ret" % tag])
        assert portlib.rng_draws() == int(g["opt_%s_draws" % tag])
        assert np.array_equal(bn[3:], g["opt_%s_bn" % tag][3:]) and np.array_equal(bs[3:], g["opt_%s_bs" % tag][3:])
        eng.set_tree(bn, bs)
        assert eng.tree_score() == ret
    eng.set_option("sankoff_exact", 0)
    portlib.seed_rng(77)
    best, bn, bs, nins, _ = eng.stepwise_addition(int(g["ras_seed"]), mt, portlib.rng_fn_address())
    assert best == int(g["ras_ret"])
    assert portlib.rng_draws() == int(g["ras_draws"])
    assert np.array_equal(bn[3:], g["ras_bn"][3:]) and np.array_equal(bs[3:], g["ras_bs"][3:])


@pytest.mark.parametrize("n,L,dt,seed,hi", [(60, 4000, 1, 51, 6), (40, 1200, 2, 52, 5), (25, 600, 6, 53, 4), (18, 400, 0, 54, 7),
                                             (90, 3000, 1, 55, 40)])
def test_oracle_seeded_cases(n, L, dt, seed, hi):
    from mpboot_b200.engine import Engine
    c = make_case(n, L, dt, seed)
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, hi, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
    ninf = c["n_inf"]
    seg = np.array([s for s in range(80, ninf, 80)] + [ninf], dtype=np.int32)
    ora = portlib.OracleEngine(c["codes"], c["weights"], dt)
    hi_o = ora.set_cost_matrix(cost.astype(np.uint32), seg)
    ora.set_ring(c["bn"], c["bs"])
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], dt)
    assert eng.set_cost_matrix(cost, seg) == hi_o
    eng.set_tree(c["bn"], c["bs"])
    ora.allocate(per_site=False)
    s0 = ora.evaluate_full(per_site=False)
    assert eng.tree_score() == s0
    assert np.array_equal(eng.sankoff_layout()[1], ora.remainder_bounds())
    ora.allocate(per_site=True)
    assert ora.evaluate_full(per_site=True) == s0
    pp, sm = eng.pattern_parsimony()
    opp, osm = ora.pattern_parsimony(ninf)
    assert sm == osm and np.array_equal(pp[:ninf], opp[:ninf])
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 5)
    portlib.seed_rng(5)
    for i in range(1, 2 * n - 1):
        ora.record(False)
        ora.rearrange(i, 1, 5, True, s0)
        assert np.array_equal(ora.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), i
    for exact in (0, 1):                                  # whole searches, both modes, then RAS
        eng.set_option("sankoff_exact", exact)
        portlib.seed_rng(99)
        ora.set_ring(c["bn"], c["bs"]); ora.allocate(bool(exact))
        want = ora.optimize_spr(1, 5, bb=bool(exact))
        wd = portlib.rng_draws()
        wbn, wbs = ora.get_ring()
        portlib.seed_rng(99)
        ret, bn, bs, _ = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 5)
        assert ret == want and portlib.rng_draws() == wd
        assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])
    eng.set_option("sankoff_exact", 0)
    portlib.seed_rng(11)
    want = ora.ras(777 + seed, 4); wd = portlib.rng_draws(); wbn, wbs = ora.get_ring()
    portlib.seed_rng(11)
    best, bn, bs, _, _ = eng.stepwise_addition(777 + seed, 4, portlib.rng_fn_address())
    assert best == want and portlib.rng_draws() == wd
    assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])


@pytest.mark.parametrize("n,L,dt,seed,hi", [(40, 2500, 1, 151, 6), (24, 900, 2, 152, 5), (20, 500, 6, 153, 4), (18, 400, 0, 154, 7),
                                             (70, 2000, 1, 155, 30)])
def test_asymmetric_cost_matrix(n, L, dt, seed, hi):
    """cost[i][j] != cost[j][i] is legal in the reference (parstree.cpp:31-90 only repairs the triangle inequality): scores then
    depend on the root, so the device takes the reference's rooted forms -- an insertion is evaluated at r, the node above the
    insertion point (testInsertParsimony :2160 on p->next->next), a stepwise insertion at the new tip (:2994-2998), the tree at
    tr->start's neighbour.  Everything the symmetric cases check, against the oracle (pinned to the live reference on asymmetric
    matrices by tests/test_oracle_cpu.py)."""
    from mpboot_b200.engine import Engine
    c = make_case(n, L, dt, seed)
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, hi, size=(S, S)); np.fill_diagonal(cost, 0)
    if np.array_equal(cost, cost.T):
        cost[0, 1] += 1
    assert not np.array_equal(cost, cost.T)
    ninf = c["n_inf"]
    seg = np.array([s for s in range(80, ninf, 80)] + [ninf], dtype=np.int32)
    ora = portlib.OracleEngine(c["codes"], c["weights"], dt)
    hi_o = ora.set_cost_matrix(cost.astype(np.uint32), seg)
    ora.set_ring(c["bn"], c["bs"])
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], dt)
    assert eng.set_cost_matrix(cost, seg) == hi_o
    eng.set_tree(c["bn"], c["bs"])
    ora.allocate(per_site=True)
    s0 = ora.evaluate_full(per_site=True)
    assert eng.tree_score() == s0
    pp, sm = eng.pattern_parsimony()
    opp, osm = ora.pattern_parsimony(ninf)
    assert sm == osm and np.array_equal(pp[:ninf], opp[:ninf])
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 5)
    for i in range(1, 2 * n - 1):
        ora.record(False)
        ora.rearrange(i, 1, 5, True, s0)
        assert np.array_equal(ora.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), i
    for exact in (0, 1):                                  # whole searches, both modes, then RAS
        eng.set_option("sankoff_exact", exact)
        portlib.seed_rng(99)
        ora.set_ring(c["bn"], c["bs"]); ora.allocate(bool(exact))
        want = ora.optimize_spr(1, 5, bb=bool(exact))
        wd = portlib.rng_draws()
        wbn, wbs = ora.get_ring()
        portlib.seed_rng(99)
        ret, bn, bs, _ = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 5)
        assert ret == want and portlib.rng_draws() == wd
        assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])
    eng.set_option("sankoff_exact", 0)
    portlib.seed_rng(11)
    want = ora.ras(777 + seed, 4); wd = portlib.rng_draws(); wbn, wbs = ora.get_ring()
    portlib.seed_rng(11)
    best, bn, bs, _, _ = eng.stepwise_addition(777 + seed, 4, portlib.rng_fn_address())
    assert best == want and portlib.rng_draws() == wd
    assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])


@pytest.mark.parametrize("n,L,dt,seed,scale", [(20, 600, 1, 251, 700), (14, 300, 2, 252, 900), (30, 1500, 1, 253, 70000)])
def test_short_off_u32_sums(n, L, dt, seed, scale):
    """-short_off (tools.cpp:2365; option "sankoff_u32"): the reference's 32-bit vectors -- pattern weights are not cut to 16 bits
    and the per-segment weighted sums do not wrap at 2^16.  Pattern weights scaled up so that every segment sum exceeds 2^16 (and,
    in the last case, single weights exceed it): tree score, pattern vector, every insertion score of a sweep, whole searches in
    both modes and a RAS tree against the reference itself (oracle/_ref with sankoff_short_int = false)."""
    import ctypes as C
    from oracle import reflib
    from mpboot_b200.engine import Engine
    if not reflib.available():
        pytest.skip("oracle/_ref not built")
    c = make_case(n, L, dt, seed)
    w = (c["weights"].astype(np.int64) * scale).astype(np.int32)
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, 5, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
    ninf = c["n_inf"]
    seg = np.array([s for s in range(80, ninf, 80)] + [ninf], dtype=np.int32)
    L_ = reflib.lib()
    fn = C.cast(L_.mpref_random_double, C.c_void_p).value
    ref = reflib.RefEngine(c["chars"], w, dt, n_informative=ninf)
    ref.set_sankoff_short(False)
    hi = ref.set_cost_matrix(cost.astype(np.uint32), seg)
    ref.set_ring(c["bn"], c["bs"]); ref.allocate(per_site=True)
    s0 = ref.evaluate_full(per_site=True)
    eng = Engine()
    eng.load_alignment(c["codes"], w, dt)
    eng.set_option("sankoff_u32", 1)
    assert eng.set_cost_matrix(cost, seg) == hi
    eng.set_tree(c["bn"], c["bs"])
    assert eng.tree_score() == s0
    assert s0 > 65535 * len(seg) // 2                      # the 16-bit wrap would have changed it
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 5)
    for i in range(1, 2 * n - 1):
        ref.record(False)
        ref.rearrange(i, 1, 5, True, s0)
        assert np.array_equal(ref.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), i
    for exact in (0, 1):
        eng.set_option("sankoff_exact", exact)
        L_.mpref_seed_rng(99)
        ref.set_ring(c["bn"], c["bs"]); ref.allocate(bool(exact))
        want = ref.optimize_spr(1, 5, bb=bool(exact))
        wd = L_.mpref_rng_draws()
        wbn, wbs = ref.get_ring()
        L_.mpref_seed_rng(99)
        ret, bn, bs, _ = eng.optimize_spr(c["bn"], c["bs"], fn, 1, 5)
        assert ret == want and L_.mpref_rng_draws() == wd
        assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])


def test_unit_costs_equal_fitch_and_switch_back():
    from mpboot_b200.engine import Engine
    c = make_case(30, 2000, 1, 61)
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], 1)
    eng.set_tree(c["bn"], c["bs"])
    fitch = eng.tree_score()
    order = eng.visit_order()
    vb, mp_f, _, _ = eng.scan_visits(order, 1, 58, 1, 4)
    eng.set_cost_matrix((1 - np.eye(4)).astype(np.uint32), np.array([c["n_inf"]], dtype=np.int32))
    assert eng.tree_score() == fitch
    vb2, mp_s, _, _ = eng.scan_visits(order, 1, 58, 1, 4)
    assert np.array_equal(vb, vb2) and np.array_equal(mp_f, mp_s)
    eng.set_cost_matrix(None, None)
    assert eng.tree_score() == fitch


def test_preconditions_fail_loudly():
    from mpboot_b200.engine import Engine, MpGpuError
    c = make_case(12, 300, 1, 62)
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], 1)
    seg = np.array([c["n_inf"]], dtype=np.int32)
    with pytest.raises(MpGpuError, match="too large"):
        eng.set_cost_matrix((6000 * (1 - np.eye(4))).astype(np.uint32), seg)
    with pytest.raises(MpGpuError, match="segment_upper"):
        eng.set_cost_matrix((1 - np.eye(4)).astype(np.uint32), np.array([c["n_inf"] + 1], dtype=np.int32))
    with pytest.raises(MpGpuError, match="segment_upper"):
        eng.set_cost_matrix((1 - np.eye(4)).astype(np.uint32), np.array([24, c["n_inf"]], dtype=np.int32))      # interior bound not a multiple of 16


def test_last_segment_bound_below_pll_informative_count():
    """IQ-TREE's informative count (ras_pars_score != 0, the last segment bound, iqtree.cpp:3814) can be smaller than PLL's
    (two distinct codes, sprparsimony.cpp:2488-2495): the patterns in between (e.g. A next to R) cost nothing on any tree and
    the reference never sums them (:944-948 stops at pllSegmentUpper).  Scores must equal the reference's with those bounds."""
    from mpboot_b200.engine import Engine
    from mpboot_b200 import hostprep, synth
    n = 14
    chars = synth.evolve_alignment(n, 500, 1, 0.05, 65, gap=0.0, amb=0.0)
    extra = np.full((n, 3), ord("A"), dtype=np.uint8)
    extra[3, 0] = ord("R"); extra[5, 1] = ord("M"); extra[7, 2] = ord("N"); extra[8, 2] = ord("R")     # zero-cost, two codes each
    prep = hostprep.prepare(np.concatenate([chars, extra], axis=1), 1)
    bn, bs = synth.random_tree_rings(n, np.random.default_rng(3))
    cost = np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.uint32)
    eng = Engine()
    eng.load_alignment(prep["codes"], prep["weights"], 1)
    eng.set_tree(bn, bs)
    pp, _ = eng.pattern_parsimony()
    positive = int(np.count_nonzero(pp[: prep["n_inf"]]))
    order = np.argsort(-(pp[: prep["n_inf"]].astype(np.int64) * prep["weights"][: prep["n_inf"]]), kind="stable")   # score x freq, like optimizeAlignment
    idx = np.concatenate([order, np.arange(prep["n_inf"], prep["codes"].shape[1])])
    codes = np.ascontiguousarray(prep["codes"][:, idx]); chars2 = np.ascontiguousarray(prep["chars"][:, idx]); w = np.ascontiguousarray(prep["weights"][idx])
    assert positive < prep["n_inf"]
    seg = np.array([16, 32, positive], dtype=np.int32)
    ora = _reference(chars2, w, 1, prep["n_inf"], codes)
    hi = ora.set_cost_matrix(cost, seg)
    ora.set_ring(bn, bs); ora.allocate(per_site=True)
    want = ora.evaluate_full(per_site=True)
    eng2 = Engine()
    eng2.load_alignment(codes, w, 1)
    assert eng2.set_cost_matrix(cost, seg) == hi
    eng2.set_tree(bn, bs)
    assert eng2.tree_score() == want
    order_v = eng2.visit_order()
    for i in (1, n, 2 * n - 2):
        vb, mp, _, _ = eng2.scan_visits(order_v, i, 1, 1, 6)
        ora.record(False); ora.rearrange(i, 1, 6, True, want)
        assert np.array_equal(ora.saved()[1:], mp.astype(np.int32))


@pytest.mark.parametrize("n,L,dt,seed", [(4, 60, 1, 71), (5, 40, 1, 72), (7, 90, 0, 73), (9, 30, 2, 74)])
def test_small_and_ragged_inputs(n, L, dt, seed):
    """Minimum taxon counts, fewer informative patterns than one 16-lane vector, a single segment."""
    from mpboot_b200.engine import Engine
    c = make_case(n, L, dt, seed, mu=0.3)
    if c["n_inf"] == 0:
        pytest.skip("no informative pattern in this draw")
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, 7, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
    seg = np.array([c["n_inf"]], dtype=np.int32)
    ora = portlib.OracleEngine(c["codes"], c["weights"], dt)
    ora.set_cost_matrix(cost.astype(np.uint32), seg)
    ora.set_ring(c["bn"], c["bs"])
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], dt)
    eng.set_cost_matrix(cost, seg)
    eng.set_tree(c["bn"], c["bs"])
    ora.allocate(per_site=True)
    s0 = ora.evaluate_full(per_site=True)
    assert eng.tree_score() == s0
    order = eng.visit_order()
    mt = 3
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, mt)
    for i in range(1, 2 * n - 1):
        ora.record(False)
        ora.rearrange(i, 1, mt, True, s0)
        assert np.array_equal(ora.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), i
    portlib.seed_rng(3)
    ora.set_ring(c["bn"], c["bs"]); ora.allocate(False)
    want = ora.optimize_spr(1, mt, bb=False); wd = portlib.rng_draws(); wbn, wbs = ora.get_ring()
    portlib.seed_rng(3)
    ret, bn, bs, _ = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, mt)
    assert ret == want and portlib.rng_draws() == wd and np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:])


def test_reweighting_under_cost():
    """Ratchet / bootstrap re-weighting (mpgpu_set_weights) with a cost matrix set: zero weights, weights above 255, the
    16-bit truncation of informativePtnWgt (:2755) and segment sums that wrap."""
    from mpboot_b200.engine import Engine
    n, dt = 26, 1
    c = make_case(n, 900, dt, 81)
    ninf = c["n_inf"]
    cost = np.array([[0, 3, 1, 3], [3, 0, 3, 1], [1, 3, 0, 3], [3, 1, 3, 0]], dtype=np.uint32)
    seg = np.array([s for s in range(64, ninf, 64)] + [ninf], dtype=np.int32)
    rng = np.random.default_rng(5)
    w2 = c["weights"].copy()
    w2[:ninf] = rng.integers(0, 4, size=ninf)
    w2[3] = 300; w2[10] = 9000; w2[11] = 70000        # a segment sum that wraps; a weight that loses its high bits as u16
    ora = portlib.OracleEngine(c["codes"], c["weights"], dt)
    ora.set_cost_matrix(cost, seg)
    eng = Engine()
    eng.load_alignment(c["codes"], c["weights"], dt)
    eng.set_cost_matrix(cost, seg)
    eng.set_tree(c["bn"], c["bs"])
    for w in (w2, c["weights"]):
        ora.set_weights(w); ora.set_ring(c["bn"], c["bs"]); ora.allocate(per_site=True)
        s0 = ora.evaluate_full(per_site=True)
        eng.set_weights(w)
        assert eng.tree_score() == s0
        order = eng.visit_order()
        vb, mp, _, _ = eng.scan_visits(order, 1, 2 * n - 2, 1, 4)
        for i in (1, 7, n + 3, 2 * n - 2):
            ora.record(False)
            ora.rearrange(i, 1, 4, True, s0)
            assert np.array_equal(ora.saved()[1:], mp[vb[i - 1]: vb[i]].astype(np.int32)), i


def test_full_size_c2_additivity_over_sites_under_cost():
    """Full BASELINE size (C2: 200 x 100 000) under -cost: with segment bounds that respect the cut, the weighted score
    (per-segment 16-bit sums included), every insertion score of a whole sweep and the node scores (mod 2^16) on the whole
    alignment equal the sums over its two column halves; the early-exit bounds stay consistent with the totals."""
    from tests.test_gpu_parity import _c2_parts
    from mpboot_b200.engine import Engine
    n, dt, full, a, b, bn, bs = _c2_parts()
    cost = np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.uint32)
    na, nb = a["n_inf"], b["n_inf"]
    assert na % 16 == 0
    seg_a = np.array(list(range(96, na, 96)) + [na], dtype=np.int32)
    seg_b = np.array(list(range(96, nb, 96)) + [nb], dtype=np.int32)
    seg_f = np.concatenate([seg_a, na + seg_b]).astype(np.int32)
    engs = []
    for p, seg in ((full, seg_f), (a, seg_a), (b, seg_b)):
        e = Engine()
        e.load_alignment(p["codes"], p["weights"], dt)
        e.set_cost_matrix(cost, seg)
        e.set_tree(bn, bs)
        engs.append(e)
    s = [e.tree_score() for e in engs]
    assert s[0] == s[1] + s[2]
    for node, slot in ((n + 1, 0), (n + 77, 2), (2 * n - 2, 1)):
        assert engs[0].view_length(node, slot) == (engs[1].view_length(node, slot) + engs[2].view_length(node, slot)) & 0xFFFF
    order = engs[0].visit_order()
    res = [e.scan_visits(order, 1, 2 * n - 2, 1, 6) for e in engs]
    assert len(res[0][1]) == 14476
    assert np.array_equal(res[0][1].astype(np.int64), res[1][1].astype(np.int64) + res[2][1].astype(np.int64))
    est = engs[0].scan_bounds(14476)
    lb = engs[0].sankoff_layout()[1]
    assert (est.astype(np.int64) >= int(lb[0])).all()          # est_max >= first prefix + its remainder bound >= the bound alone
