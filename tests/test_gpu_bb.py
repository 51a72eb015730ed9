"""GPU parity tests of the -bb path (R8 + the -bb search): replicate scores from the device
(int8 tensor-core contraction + exact exception path) against the oracle's u16-lane REPS on the
same pattern vectors, and whole -bb SPR searches against the oracle's restatement of
IQTree::saveCurrentTree (itself pinned to the reference driver by tests/test_bb_cpu.py).
Bit-exact: integers only."""
import numpy as np
import pytest

from oracle import portlib
from tests.helpers import fingerprint_ring
from tests.test_bb_cpu import bb_setup, run_bb

pytestmark = pytest.mark.gpu

CASES = [(12, 300, 1, 7, 50, 0.05), (24, 400, 2, 5, 40, 0.05), (30, 800, 1, 21, 64, 0.01), (20, 300, 6, 9, 30, 0.05),
         (40, 1500, 1, 11, 300, 0.05), (16, 300, 0, 3, 20, 0.05)]


def _engine(c, boot, seg, tensor, ratchet=None, cost=None):
    from mpboot_b200.engine import Engine
    eng = Engine()
    eng.set_option("reps_tensor", tensor)
    eng.load_alignment(c["codes"], c["weights"], c["datatype"])
    if ratchet is not None:
        eng.set_weights(ratchet[0])                 # the search runs on the perturbed frequencies
    if cost is not None:
        eng.set_cost_matrix(cost, seg)              # -cost: before the replicates are loaded
    eng.set_tree(c["bn"], c["bs"])
    eng.load_replicates(boot, seg, original_sample=None if ratchet is None else ratchet[1])
    return eng


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES)
def test_reps_current_tree_and_candidates(n, L, dt, seed, B, mu, tensor):
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    eng = _engine(c, boot, seg, tensor)
    want = portlib.reps(pp, boot[:, : c["n_inf"]], seg)
    assert np.array_equal(eng.reps_current_tree(), want)
    groups, exc, t = eng.reps_info()
    assert t == tensor and groups >= 2 and exc > 0            # the heavy replicates make segments wrap-prone
    # every saveCurrentTree call of three node visits: current tree first, then each insertion
    order = eng.visit_order()
    for i in (1, n + 1, 2 * n - 2):
        o.set_ring(c["bn"], c["bs"]); o.allocate(True); o.evaluate_full(True)
        o.record(True)
        o.rearrange(i, 1, 6, True, s0)
        mps, ptn = o.saved(True)
        vb, mp, cref, cprune = eng.scan_visits(order, i, 1, 1, 6)
        assert np.array_equal(mps[1:], mp.astype(np.int32))
        got = eng.reps_candidates(np.arange(-1, len(mp), dtype=np.int32))
        for k in range(len(mps)):
            assert np.array_equal(got[k], portlib.reps(ptn[k, : c["n_inf"]], boot[:, : c["n_inf"]], seg)), (i, k)


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:3])
def test_reps_nowrap_is_the_plain_int_dot_product(n, L, dt, seed, B, mu, tensor):
    """-autovec (option reps_nowrap, iqtree.cpp:3418-3423): res = sum_ptn pattern_pars * boot_sample as plain ints -- on replicates
    whose 16-bit segment sums do wrap, so the two semantics differ."""
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    ninf = c["n_inf"]
    w = boot[:, :ninf].astype(np.int64)
    eng = _engine(c, boot, seg, tensor)
    wrapped = eng.reps_current_tree()
    eng.set_option("reps_nowrap", 1)
    plain = eng.reps_current_tree()
    assert np.array_equal(plain, w @ pp[:ninf].astype(np.int64)) and not np.array_equal(plain, wrapped)
    order = eng.visit_order()
    o.set_ring(c["bn"], c["bs"]); o.allocate(True); o.evaluate_full(True)
    o.record(True)
    o.rearrange(n + 1, 1, 6, True, s0)
    mps, ptn = o.saved(True)
    vb, mp, cref, cprune = eng.scan_visits(order, n + 1, 1, 1, 6)
    got = eng.reps_candidates(np.arange(-1, len(mp), dtype=np.int32))
    assert np.array_equal(got, (w @ ptn[:, :ninf].astype(np.int64).T).T)
    eng.set_option("reps_nowrap", 0)
    assert np.array_equal(eng.reps_current_tree(), wrapped)


def _same_as_oracle(g, w):
    assert g["ret"] == w["ret"] and g["draws"] == w["draws"]
    assert np.array_equal(g["ring"][0][3:], w["ring"][0][3:]) and np.array_equal(g["ring"][1][3:], w["ring"][1][3:])
    assert all(np.array_equal(x, y) for x, y in zip(g["state"], w["state"]))
    assert g["ncalls"] == w["counters"][0] and g["nreps"] == w["counters"][2]
    assert np.array_equal(g["treels"], w["treels"])
    assert np.array_equal(g["mats"][:, :3], w["mats"][:, 1:4]) and np.array_equal(g["mats"][:, 3], w["mats"][:, 4])


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:4])
def test_bb_search_ratchet_iteration_matches_oracle(n, L, dt, seed, B, mu, tensor):
    """on_ratchet_hclimb1 (iqtree.cpp:3283-3294): search on perturbed frequencies, cur_logl of every call =
    original-frequency score of the previous passing call's tree; the chain breaks for good once such a
    score fails the cutoff."""
    from tests.test_bb_cpu import ratchet_setup
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    rt = ratchet_setup(c, pp, seed)
    a = run_bb(o, c, boot, seg, 0.0, None, False, ratchet=rt)
    _same_as_oracle(_run_gpu_bb(c, boot, seg, 0.0, tensor, ratchet=rt), a)
    top = np.unique(-a["treels"])[::-1]
    for worst in top[:2]:
        cutoff = -(worst - 0.5)
        w = run_bb(o, c, boot, seg, cutoff, None, False, ratchet=rt)
        _same_as_oracle(_run_gpu_bb(c, boot, seg, cutoff, tensor, ratchet=rt), w)


def test_reps_wrap_free_bulk_uses_no_exceptions():
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(40, 1500, 1, 11, 100, heavy=False)
    eng = _engine(c, boot, seg, 1)
    assert np.array_equal(eng.reps_current_tree(), portlib.reps(pp, boot[:, : c["n_inf"]], seg))
    groups, exc, t = eng.reps_info()
    assert groups == 1 and exc == 0 and t == 1                # everything on the tensor path


def _run_gpu_bb(c, boot, seg, cutoff, tensor, seed=2024, mt=6, ratchet=None, mulhits=False, cost=None, topboot=0):
    from mpboot_b200.engine import Treels
    eng = _engine(c, boot, seg, tensor, ratchet, cost)
    B = boot.shape[0]
    bl = np.full(B, -float(np.iinfo(np.int64).max), dtype=np.float64)   # -LONG_MAX, iqtree.cpp:248
    bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
    tl = Treels(c["n"])
    portlib.seed_rng(seed)
    ret, bn, bs, nins, ncalls, nreps = eng.optimize_spr_bb(c["bn"], c["bs"], tl.hooks(portlib.rng_fn_address()),
                                                           bl, bc, bt, cutoff, 0.5, 1, mt,
                                                           ratchet_pattern_pars=None if ratchet is None else ratchet[2],
                                                           mulhits=mulhits, topboot=topboot)
    top = tl.toplists(B)
    thr = eng._boot_threshold.copy() if topboot else np.full(B, -(2 ** 31 - 1), dtype=np.int32)
    return dict(ret=ret, draws=portlib.rng_draws(), ring=(bn, bs), state=(bl, bc, bt), ncalls=ncalls, nreps=nreps,
                treels=tl.logl(), mats=tl.materialized(), nins=nins, mulhits=tl.mulhits(B), toplists=(top[0], thr, top[1]))


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:5])
def test_bb_search_matches_oracle(n, L, dt, seed, B, mu, tensor):
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    want = run_bb(o, c, boot, seg, 0.0, None, False)
    for cutoff in (0.0, -(want["ret"] + 4.0)):
        w = run_bb(o, c, boot, seg, cutoff, None, False)
        g = _run_gpu_bb(c, boot, seg, cutoff, tensor)
        assert g["ret"] == w["ret"] and g["draws"] == w["draws"]
        assert np.array_equal(g["ring"][0][3:], w["ring"][0][3:]) and np.array_equal(g["ring"][1][3:], w["ring"][1][3:])
        assert all(np.array_equal(x, y) for x, y in zip(g["state"], w["state"]))          # boot_logl, boot_counts, boot_trees
        assert g["ncalls"] == w["counters"][0] and g["nreps"] == w["counters"][2]
        assert np.array_equal(g["treels"], w["treels"])
        wm = w["mats"]
        assert len(g["mats"]) == len(wm)
        assert np.array_equal(g["mats"][:, :3], wm[:, 1:4])                              # pruned ref, insertion ref, tree_index
        assert np.array_equal(g["mats"][:, 3], wm[:, 4])                                 # topology fingerprint


# ---- the committed golden vectors: what the reference driver produced for the same -bb search ----
from tests.test_bb_cpu import CASE_FILES, IDS, check_against_golden  # noqa: E402


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_bb_search_matches_golden(path):
    g = dict(np.load(path))
    n, dt, mt = int(g["n"]), int(g["datatype"]), int(g["maxtrav"])
    c = dict(n=n, datatype=dt, codes=g["codes"], weights=g["weights"], n_inf=int(g["n_inf"]), bn=g["bn"], bs=g["bs"])
    for tag in ("all", "cut", "rall", "rcut"):
        rt = (g["bb_ratchet_weights"], g["bb_ratchet_orig"], g["bb_ratchet_init"]) if tag[0] == "r" else None
        r = _run_gpu_bb(c, g["bb_boot"], g["bb_seg"], float(g["bb_%s_cutoff" % tag]), 1, mt=mt, ratchet=rt)
        r["counters"] = (r["ncalls"], len(r["treels"]), r["nreps"])
        r["mats"] = r["mats"][:, [2, 3]]
        check_against_golden(g, tag, r)


# ---- -mulhits (params->multiple_hits, iqtree.cpp:3498-3531): policy MPGPU_BB_MULHITS ----
from tests.test_bb_cpu import MULHITS_CASES, MULHITS_GOLD, check_mulhits_golden, mulhits_golden_case  # noqa: E402


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("k", range(len(MULHITS_CASES)))
def test_bb_mulhits_matches_golden_and_oracle(k, tensor):
    g = dict(np.load(MULHITS_GOLD))
    c, o, seg, boot, bound = mulhits_golden_case(g, k)
    for tag in ("all", "cut"):
        cutoff = float(g["c%d_%s_cutoff" % (k, tag)])
        r = _run_gpu_bb(c, boot, seg, cutoff, tensor, mulhits=True)
        r["mats_tf"] = r["mats"][:, [2, 3]]
        check_mulhits_golden(g, k, tag, r)                       # what the reference driver produced
        w = run_bb(o, c, boot, seg, cutoff, None, False, mulhits=True)
        assert r["ncalls"] == w["counters"][0] and r["nreps"] == w["counters"][2]
        assert np.array_equal(r["mats"][:, :2], w["mats"][:, 1:3])   # pruned ref, insertion ref of every materialised tree
        assert np.array_equal(r["state"][1], w["state"][1]) and np.array_equal(r["state"][2], w["state"][2])   # untouched


def test_bb_mulhits_ratchet_iteration_matches_oracle():
    from tests.test_bb_cpu import ratchet_setup
    n, L, dt, seed, B, mu = MULHITS_CASES[2]
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    rt = ratchet_setup(c, pp, seed)
    w = run_bb(o, c, boot, seg, 0.0, None, False, ratchet=rt, mulhits=True)
    r = _run_gpu_bb(c, boot, seg, 0.0, 1, ratchet=rt, mulhits=True)
    assert r["ret"] == w["ret"] and r["draws"] == w["draws"]
    assert np.array_equal(r["state"][0], w["state"][0])
    assert all(np.array_equal(x, y) for x, y in zip(r["mulhits"], w["mulhits"]))
    assert np.array_equal(r["treels"], w["treels"])


# ---- -cost together with -bb: REPS on Sankoff pattern vectors (k_sk_scan<ROWS> + k_sk_reps) ----
from tests.test_bb_cpu import (SANKOFF_BB_CASES, SANKOFF_BB_GOLD, check_sankoff_bb_golden, sankoff_bb_cost,  # noqa: E402
                               sankoff_bb_golden_case)


@pytest.mark.parametrize("k", range(len(SANKOFF_BB_CASES)))
def test_sankoff_reps_vectors_match_oracle(k):
    n, L, dt, seed, B, mu = SANKOFF_BB_CASES[k]
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    cost = sankoff_bb_cost(dt, seed)
    ninf = c["n_inf"]
    o.set_cost_matrix(cost, seg)
    try:
        eng = _engine(c, boot, seg, 1, cost=cost)
        o.set_ring(c["bn"], c["bs"]); o.allocate(True)
        s1 = o.evaluate_full(True)
        spp, _ = o.pattern_parsimony(ninf)
        assert np.array_equal(eng.reps_current_tree(), portlib.reps(spp, boot[:, :ninf], seg))
        order = eng.visit_order()
        for i in (1, n + 1, 2 * n - 2):
            o.set_ring(c["bn"], c["bs"]); o.allocate(True); o.evaluate_full(True)
            o.record(True)
            o.rearrange(i, 1, 6, True, s1)
            mps, ptn = o.saved(True)
            vb, mp, cref, cprune = eng.scan_visits(order, i, 1, 1, 6)
            assert np.array_equal(mps[1:], mp.astype(np.int32))
            got = eng.reps_candidates(np.arange(-1, len(mp), dtype=np.int32))
            for q in range(len(mps)):
                assert np.array_equal(got[q], portlib.reps(ptn[q, :ninf], boot[:, :ninf], seg)), (i, q)
    finally:
        o.set_cost_matrix(None, None)


@pytest.mark.parametrize("k", range(len(SANKOFF_BB_CASES)))
def test_sankoff_bb_search_matches_golden_and_oracle(k):
    g = dict(np.load(SANKOFF_BB_GOLD))
    c, o, seg, boot, cost = sankoff_bb_golden_case(g, k)
    try:
        for tag in ("all", "cut"):
            cutoff = float(g["c%d_%s_cutoff" % (k, tag)])
            r = _run_gpu_bb(c, boot, seg, cutoff, 1, cost=cost)
            r["mats_tf"] = r["mats"][:, [2, 3]]
            check_sankoff_bb_golden(g, k, tag, r)                # what the reference driver produced
            w = run_bb(o, c, boot, seg, cutoff, None, False, cost=cost)
            assert r["ncalls"] == w["counters"][0] and r["nreps"] == w["counters"][2]
            assert np.array_equal(r["mats"][:, :2], w["mats"][:, 1:3])
        w = run_bb(o, c, boot, seg, 0.0, None, False, cost=cost, mulhits=True)
        r = _run_gpu_bb(c, boot, seg, 0.0, 1, cost=cost, mulhits=True)
        assert r["ret"] == w["ret"] and r["draws"] == w["draws"] and np.array_equal(r["state"][0], w["state"][0])
        assert all(np.array_equal(x, y) for x, y in zip(r["mulhits"], w["mulhits"]))
    finally:
        o.set_cost_matrix(None, None)


@pytest.mark.parametrize("asym", [False, True], ids=["symmetric", "asymmetric"])
@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("k", [0, 2])
def test_sankoff_reps_tensor_path_equals_exact(k, tensor, asym):
    """Light replicate weights, costs below 256, short segments: every chunk qualifies for the tcgen05 path; the same
    vectors through the exact kernel (reps_tensor = 0) and through the oracle's u16 lanes.  asymmetric: the insertion's vector
    in its rooted form (k_sk_scan<ROWS, ASYM>) and the current tree's vector rooted at every visited edge in turn (the call
    -(2 + v), rearrangeParsimony :2286-2289), then a whole -bb search."""
    n, L, dt, seed, B, mu = SANKOFF_BB_CASES[k]
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu, heavy=False)
    cost = sankoff_bb_cost(dt, seed)
    if asym:
        cost = cost.copy(); cost[0, 1] += 1; cost[2 % cost.shape[0], 0] += 2
    ninf = c["n_inf"]
    o.set_cost_matrix(cost, seg)
    try:
        eng = _engine(c, boot, seg, tensor, cost=cost)
        o.set_ring(c["bn"], c["bs"]); o.allocate(True)
        s1 = o.evaluate_full(True)
        order = eng.visit_order()
        vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 6)
        calls = []
        for v in range(2 * n - 2):
            calls.append(-(2 + v) if asym else -1); calls.extend(range(vb[v], vb[v + 1]))
        got = eng.reps_candidates(np.array(calls, dtype=np.int32))
        row = 0
        for i in range(1, 2 * n - 1):
            o.set_ring(c["bn"], c["bs"]); o.allocate(True); o.evaluate_full(True)
            o.record(True)
            o.rearrange(i, 1, 6, True, s1)
            mps, ptn = o.saved(True)
            for q in range(len(mps)):
                assert np.array_equal(got[row], portlib.reps(ptn[q, :ninf], boot[:, :ninf], seg)), (i, q)
                row += 1
        assert row == len(calls)
        t, e = eng.sankoff_reps_stats()
        assert (t > 0 and e == 0) if tensor else (t == 0 and e > 0)
        w = run_bb(o, c, boot, seg, 0.0, None, False, cost=cost)
        r = _run_gpu_bb(c, boot, seg, 0.0, tensor, cost=cost)
        assert r["ret"] == w["ret"] and r["draws"] == w["draws"]
        assert all(np.array_equal(x, y) for x, y in zip(r["state"], w["state"]))
        assert np.array_equal(r["treels"], w["treels"])
    finally:
        o.set_cost_matrix(None, None)


@pytest.mark.parametrize("k", range(len(MULHITS_CASES)))
def test_bb_topboot_matches_golden_and_oracle(k):
    """-mulhits -topboot N (policy MPGPU_BB_MULHITS_TOP, iqtree.cpp:3536-3583)"""
    from tests.test_bb_cpu import TOPBOOT_NS
    g = dict(np.load(MULHITS_GOLD))
    c, o, seg, boot, bound = mulhits_golden_case(g, k)
    for N in TOPBOOT_NS:
        r = _run_gpu_bb(c, boot, seg, 0.0, 1, topboot=N)
        r["mats_tf"] = r["mats"][:, [2, 3]]
        check_mulhits_golden(g, k, "top%d" % N, r)
    w = run_bb(o, c, boot, seg, -(int(g["c%d_all_ret" % k]) + 4.0), None, False, topboot=3)
    r = _run_gpu_bb(c, boot, seg, -(int(g["c%d_all_ret" % k]) + 4.0), 0, topboot=3)
    assert r["ret"] == w["ret"] and r["draws"] == w["draws"] and np.array_equal(r["treels"], w["treels"])
    assert all(np.array_equal(x, y) for x, y in zip(r["toplists"], w["toplists"]))
    assert np.array_equal(r["mats"][:, :2], w["mats"][:, 1:3])


# ---- -distinct_iter_top_boot (policy MPGPU_BB_DISTINCT_ITER, iqtree.cpp:3587-3685) --------------------------------------------
def _remain_bounds(boot, seg, bound, upper):
    """boot_samples_pars_remain_bounds (IQTree::pllComputeRellRemainBound, iqtree.cpp:3841-3856) from the per-pattern lower bounds"""
    w = boot[:, :upper].astype(np.int64) * bound[:upper].astype(np.int64)
    return np.stack([w[:, seg[s]:].sum(axis=1) for s in range(len(seg) - 1)], axis=1).astype(np.int32)


def _prefix_max_numpy(ptn, boot, seg, remain):
    """max over nseg/4 < s < nseg-1 of (sum of the 16-bit segment sums up to s + remain[b][s]), the loop at iqtree.cpp:3424-3445"""
    nseg = len(seg)
    res = np.zeros(boot.shape[0], dtype=np.int64)
    best = np.full(boot.shape[0], -(2 ** 31), dtype=np.int64)
    lo = 0
    for s in range(nseg):
        hi = min(int(seg[s]), len(ptn))
        res += (boot[:, lo:hi].astype(np.int64) * ptn[lo:hi].astype(np.int64)).sum(axis=1) & 0xFFFF
        if nseg // 4 < s < nseg - 1:
            best = np.maximum(best, res + remain[:, s])
        lo = hi
    return best


@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:5])
def test_reps_prefix_max_matches_numpy(n, L, dt, seed, B, mu):
    """The left side of the remain-bound skip test for the current tree and every insertion of three node visits, all replicates,
    on data whose segments wrap (heavy replicates) -- from the oracle's pattern vector of each saveCurrentTree call."""
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    ninf = c["n_inf"]
    remain = _remain_bounds(boot, seg, bound, ninf)
    eng = _engine(c, boot, seg, 1)
    eng.set_remain_bounds(remain)
    allb = np.arange(B, dtype=np.int32)
    assert np.array_equal(eng.reps_prefix_max(-1, allb), _prefix_max_numpy(pp[:ninf], boot, seg, remain))
    order = eng.visit_order()
    for i in (1, n + 1, 2 * n - 2):
        o.set_ring(c["bn"], c["bs"]); o.allocate(True); o.evaluate_full(True)
        o.record(True)
        o.rearrange(i, 1, 6, True, s0)
        mps, ptn = o.saved(True)
        vb, mp, cref, cprune = eng.scan_visits(order, i, 1, 1, 6)
        for k in range(len(mps)):
            want = _prefix_max_numpy(ptn[k, :ninf], boot, seg, remain)
            assert np.array_equal(eng.reps_prefix_max(k - 1, allb), want), (i, k)
        sub = np.array([B - 1, 0, B // 2], dtype=np.int32)                     # a short, unordered list
        assert np.array_equal(eng.reps_prefix_max(len(mps) - 2, sub), _prefix_max_numpy(ptn[len(mps) - 1, :ninf], boot, seg, remain)[sub])


def _run_gpu_bb_distinct(c, boot, seg, tensor, K, remain, iters=3, seed=2024, mt=6, cutoff=0.0):
    from mpboot_b200 import synth
    from mpboot_b200.engine import Treels
    eng = _engine(c, boot, seg, tensor)
    if remain is not None:
        eng.set_remain_bounds(remain)
    B = boot.shape[0]
    bl = np.full(B, -float(np.iinfo(np.int64).max), dtype=np.float64)
    bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
    thr = np.full(B, -(2 ** 31 - 1), dtype=np.int32)
    tl = Treels(c["n"])
    portlib.seed_rng(seed)
    rets, ncalls, nreps = [], 0, 0
    for it in range(1, iters + 1):
        bn, bs = (c["bn"], c["bs"]) if it == 1 else synth.random_tree_rings(c["n"], np.random.default_rng(5000 + it))
        ret, bn, bs, nins, nc, nr = eng.optimize_spr_bb(bn, bs, tl.hooks(portlib.rng_fn_address()), bl, bc, bt, cutoff, 0.5, 1, mt,
                                                        distinct=K, cur_it=it, boot_threshold=thr)
        rets.append(ret); ncalls += nc; nreps += nr
    top = tl.toplists(B)
    return dict(ret=rets, draws=portlib.rng_draws(), ring=(bn, bs), state=(bl, bc, bt), ncalls=ncalls, nreps=nreps, treels=tl.logl(),
                mats=tl.materialized(), toplists=(top[0], thr, top[1]), topiters=tl.topiters(B))


@pytest.mark.parametrize("tensor", [0, 1], ids=["exact-cuda-core", "tensor"])
@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:3])
@pytest.mark.parametrize("K", [1, 3])
def test_bb_distinct_iter_matches_oracle(n, L, dt, seed, B, mu, K, tensor):
    """Three iterations of -bb SPR searches under -distinct_iter_top_boot K on one bookkeeping state: lists, iterations, thresholds,
    boot_logl / counts / trees, draws, treels and materialised trees equal the oracle's -- without remain bounds and with them
    (where the skip of replicates changes decisions, tests/test_bb_cpu.py)."""
    from tests.test_bb_cpu import run_bb_distinct
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    for bnd in (None, bound):
        w = run_bb_distinct(o, c, boot, seg, bnd, False, K)
        g = _run_gpu_bb_distinct(c, boot, seg, tensor, K, None if bnd is None else _remain_bounds(boot, seg, bnd, c["n_inf"]))
        assert g["ret"] == w["ret"] and g["draws"] == w["draws"]
        assert np.array_equal(g["ring"][0][3:], w["ring"][0][3:]) and np.array_equal(g["ring"][1][3:], w["ring"][1][3:])
        assert all(np.array_equal(x, y) for x, y in zip(g["state"], w["state"]))
        assert all(np.array_equal(x, y) for x, y in zip(g["toplists"], w["toplists"])) and np.array_equal(g["topiters"], w["topiters"])
        assert g["ncalls"] == w["counters"][0] and g["nreps"] == w["counters"][2]
        assert np.array_equal(g["treels"], w["treels"])
        assert np.array_equal(g["mats"][:, :3], w["mats"][:, 1:4]) and np.array_equal(g["mats"][:, 3], w["mats"][:, 4])


@pytest.mark.parametrize("n,L,dt,seed,B,mu", CASES[:3])
def test_bb_min_iter1_cand_and_cutoff_from_btrees(n, L, dt, seed, B, mu):
    """-min_iter1_cand in iteration 1 (state.updates_off, iqtree.cpp:3404): every call that passes the cutoff lands in treels_logl and
    nothing else happens -- the search is the plain search (same draws, same tree), replicates untouched, nothing materialised.
    -cutoff_from_btrees (state.boot_tree_orig_logl, :3717-3718 / :3524-3527): each replicate ends up with the score, on the original
    alignment, of the tree it holds.  (The patched program's outputs under both options equal the stock binary's:
    tests/test_gpu_dropin.py.)"""
    from mpboot_b200.engine import Engine, Treels
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    full = _run_gpu_bb(c, boot, seg, 0.0, 1)

    def run(**kw):
        eng = _engine(c, boot, seg, 1)
        bl = np.full(B, -float(np.iinfo(np.int64).max), dtype=np.float64)
        bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
        tl = Treels(c["n"])
        portlib.seed_rng(2024)
        ret, bn, bs, nins, ncalls, nreps = eng.optimize_spr_bb(c["bn"], c["bs"], tl.hooks(portlib.rng_fn_address()), bl, bc, bt, 0.0, 0.5, 1, 6, **kw)
        return dict(ret=ret, draws=portlib.rng_draws(), ring=(bn, bs), state=(bl, bc, bt), treels=tl.logl(), mats=tl.materialized(), tl=tl,
                    ncalls=ncalls, nins=nins)

    portlib.seed_rng(2024)
    eng = Engine(); eng.load_alignment(c["codes"], c["weights"], c["datatype"])
    plain_ret, pbn, pbs, plain_nins = eng.optimize_spr(c["bn"], c["bs"], portlib.rng_fn_address(), 1, 6)
    plain_draws = portlib.rng_draws()

    off = run(updates_off=True)
    assert off["ret"] == plain_ret and off["draws"] == plain_draws and off["nins"] == plain_nins      # no saveCurrentTree draw
    assert np.array_equal(off["ring"][0][3:], pbn[3:]) and np.array_equal(off["ring"][1][3:], pbs[3:])
    assert len(off["mats"]) == 0 and (off["state"][2] == -1).all() and (off["state"][1] == 0).all()
    assert len(off["treels"]) == off["ncalls"] > plain_nins                                # cutoff off: every call is appended (insertions + one per visit)
    assert off["treels"].max() == -float(plain_ret)                                        # the best tree the search saw

    orig = np.zeros(B, dtype=np.int32)
    r = run(boot_tree_orig_logl=orig)
    assert r["ret"] == full["ret"] and r["draws"] == full["draws"] and all(np.array_equal(x, y) for x, y in zip(r["state"], full["state"]))
    assert np.array_equal(orig.astype(np.float64), r["treels"][r["state"][2]])             # cur_logl of the tree each replicate holds
    # -mulhits only ever raises the entry (:3524-3527): from the reference's initial 0 (iqtree.cpp:254) a negative score never does ...
    orig2 = np.zeros(B, dtype=np.int32)
    m = run(mulhits=True, boot_tree_orig_logl=orig2)
    assert (orig2 == 0).all()
    # ... from below it climbs to the best original-alignment score among the trees that ever tied or beat the replicate's best
    orig3 = np.full(B, -10 ** 6, dtype=np.int32)
    m = run(mulhits=True, boot_tree_orig_logl=orig3)
    sizes, flat = m["tl"].mulhits(B)
    best = np.array([m["treels"][flat[sizes[:b].sum(): sizes[:b + 1].sum()]].max() for b in range(B)])
    assert (orig3 >= best).all() and np.isin(orig3.astype(np.float64), m["treels"]).all() and (orig3 < 0).all()


def test_full_size_c2_reps_linearity_and_dot_product():
    """Full BASELINE size (C2: 200 x 100 000 patterns, B = 999 replicates in three blocks A, B, A + B): with MPBoot's own
    segmentation no 16-bit segment sum wraps, so REPS is linear in the replicate frequencies -- res(A + B) = res(A) + res(B)
    for the current tree and for every insertion of 40 node visits -- and, for the current tree, equals the plain integer dot
    product of the per-pattern scores with the frequencies (numpy int64), all through the tensor-core path."""
    import bench
    from tests.test_gpu_parity import _c2_parts
    from mpboot_b200.engine import Engine
    n, dt, full, a, b, bn, bs = _c2_parts()
    ninf = full["n_inf"]
    eng = Engine()
    eng.load_alignment(full["codes"], full["weights"], dt)
    eng.set_tree(bn, bs)
    pp, sm = eng.pattern_parsimony()
    seg = bench.do_segmenting(pp[:ninf], full["weights"], ninf)
    rng = np.random.default_rng(12)
    K = 333
    wa = rng.multinomial(ninf, np.full(ninf, 1.0 / ninf), size=K).astype(np.uint16)
    wb = rng.multinomial(ninf, np.full(ninf, 1.0 / ninf), size=K).astype(np.uint16)
    boot = np.concatenate([wa, wb, wa + wb]).astype(np.uint16)
    eng.load_replicates(boot, seg)
    groups, exc, tensor = eng.reps_info()
    assert tensor == 1 and groups == 1 and exc == 0
    cur = eng.reps_current_tree().astype(np.int64)
    want = boot.astype(np.int64) @ pp[:ninf].astype(np.int64)
    assert np.array_equal(cur, want)
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 150, 40, 1, 6)
    got = eng.reps_candidates(np.arange(-1, len(mp), dtype=np.int32)).astype(np.int64)
    assert len(mp) > 1000
    assert np.array_equal(got[:, 2 * K: 3 * K], got[:, :K] + got[:, K: 2 * K])
    # a candidate's replicate score under the all-ones replicate would be its parsimony score: A + B has column sums 2 * ninf / ninf
    ones = np.ones((1, ninf), dtype=np.uint16)
    eng2 = Engine()
    eng2.load_alignment(full["codes"], full["weights"], dt)
    eng2.set_tree(bn, bs)
    eng2.load_replicates(ones, seg)
    eng2.scan_visits(order, 150, 40, 1, 6)
    r1 = eng2.reps_candidates(np.arange(-1, len(mp), dtype=np.int32))[:, 0]
    assert r1[0] == sm and np.array_equal(r1[1:].astype(np.int64), mp.astype(np.int64))
